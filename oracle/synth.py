"""Seeded synthetic inputs for the cutoffvgg16 hot path (test infrastructure, see oracle/__init__.py).

Frames mimic masked lung-ultrasound sectors; weights are Glorot-uniform in
Keras layout (HWIO conv kernels, (in,out) dense) as ``VGG16(weights=None)`` + ``Dense`` would
initialise them (reference graph: src/models/models.py:407-417).
"""
import numpy as np

from ab_line_classifier_b200.synth import VGG_MEAN_BGR  # noqa: F401

# Keras VGG16(include_top=False).layers[1:] -- (name, kind, out_channels)
VGG16_LAYERS = [
    ("block1_conv1", "conv", 64), ("block1_conv2", "conv", 64), ("block1_pool", "pool", 0),
    ("block2_conv1", "conv", 128), ("block2_conv2", "conv", 128), ("block2_pool", "pool", 0),
    ("block3_conv1", "conv", 256), ("block3_conv2", "conv", 256), ("block3_conv3", "conv", 256),
    ("block3_pool", "pool", 0),
    ("block4_conv1", "conv", 512), ("block4_conv2", "conv", 512), ("block4_conv3", "conv", 512),
    ("block4_pool", "pool", 0),
    ("block5_conv1", "conv", 512), ("block5_conv2", "conv", 512), ("block5_conv3", "conv", 512),
    ("block5_pool", "pool", 0),
]


def layer_plan(cutoff_layer=10):
    """Layers kept by ``vgg16.layers[1:cutoff_layer]`` (models.py:410)."""
    return VGG16_LAYERS[: cutoff_layer - 1]


from ab_line_classifier_b200.synth import synth_frames, vgg16_preprocess, one_hot  # noqa: E402,F401


def glorot_weights(seed=10001, cutoff_layer=10, n_classes=2, output_bias=None, cin=3):
    """Dict of Keras-named fp32 arrays: '<layer>/kernel' HWIO, '<layer>/bias', 'output/kernel' (C,n)."""
    rng = np.random.default_rng(seed)
    params = {}
    c = cin
    for name, kind, cout in layer_plan(cutoff_layer):
        if kind != "conv":
            continue
        lim = np.sqrt(6.0 / (9 * c + 9 * cout))
        params[name + "/kernel"] = rng.uniform(-lim, lim, size=(3, 3, c, cout)).astype(np.float32)
        params[name + "/bias"] = np.zeros((cout,), dtype=np.float32)
        c = cout
    lim = np.sqrt(6.0 / (c + n_classes))
    params["output/kernel"] = rng.uniform(-lim, lim, size=(c, n_classes)).astype(np.float32)
    params["output/bias"] = (np.zeros((n_classes,), dtype=np.float32) if output_bias is None
                             else np.asarray(output_bias, dtype=np.float32))
    return params


def randomize_biases(params, seed=7, scale=0.05):
    """Non-zero biases for parity tests (Glorot init leaves them zero, which hides bias bugs)."""
    rng = np.random.default_rng(seed)
    out = dict(params)
    for k in params:
        if k.endswith("/bias"):
            out[k] = rng.uniform(-scale, scale, size=params[k].shape).astype(np.float32)
    return out


def dropout_mask(batch, channels, rate, seed=10001, step=0, rank=0):
    """Keep-mask {0,1} float32 [batch,channels]; counter-based so ranks/steps are reproducible."""
    rng = np.random.default_rng([seed, step, rank])
    return (rng.random((batch, channels)) >= rate).astype(np.float32)

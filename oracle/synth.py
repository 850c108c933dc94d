"""Seeded synthetic inputs for the cutoffvgg16 hot path (test infrastructure, see oracle/__init__.py).

Frames mimic masked lung-ultrasound sectors; weights are Glorot-uniform in
Keras layout (HWIO conv kernels, (in,out) dense) as ``VGG16(weights=None)`` + ``Dense`` would
initialise them (reference graph: src/models/models.py:407-417).
"""
import numpy as np

VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], dtype=np.float32)

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


def synth_frames(n, seed=1234, h=128, w=128):
    """uint8 [n,h,w,3] grey (R=G=B) fan-sector frames; labels y = idx % 2 (0 A-lines, 1 B-lines)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ang = np.arctan2(xx - (w - 1) / 2.0, yy + 1.0)
    rad = np.hypot(xx - (w - 1) / 2.0, yy + 1.0)
    sector = (np.abs(ang) < np.deg2rad(35.0)) & (rad < 0.98 * h)
    frames = np.zeros((n, h, w), dtype=np.float32)
    labels = (np.arange(n) % 2).astype(np.int64)
    for i in range(n):
        img = rng.rayleigh(40.0, size=(h, w)).astype(np.float32)
        pleura = 0.15 * h + rng.uniform(-3, 3)
        img += 120.0 * np.exp(-0.5 * ((rad - pleura) / 1.5) ** 2)
        if labels[i] == 0:  # A-lines: horizontal arcs at equal depth spacing
            k = rng.integers(3, 6)
            for j in range(1, k + 1):
                img += (90.0 / j) * np.exp(-0.5 * ((rad - pleura * (j + 1)) / 1.5) ** 2)
        else:  # B-lines: vertical rays from the pleural line
            k = rng.integers(1, 5)
            for _ in range(k):
                a0 = rng.uniform(-0.5, 0.5)
                img += 110.0 * np.exp(-0.5 * ((ang - a0) / 0.03) ** 2) * (rad > pleura)
        frames[i] = np.where(sector, img, 0.0)
    frames = np.clip(frames, 0, 255).astype(np.uint8)
    return np.repeat(frames[..., None], 3, axis=-1), labels


def vgg16_preprocess(frames_rgb):
    """Keras ``vgg16.preprocess_input`` (caffe mode): RGB->BGR, subtract the ImageNet BGR means
    (reference binds it at models.py:13,44; applied in preprocessor.py:60-63)."""
    x = np.asarray(frames_rgb, dtype=np.float32)[..., ::-1]
    return np.ascontiguousarray(x - VGG_MEAN_BGR)


def one_hot(labels, n_classes=2):
    y = np.zeros((len(labels), n_classes), dtype=np.float32)
    y[np.arange(len(labels)), labels] = 1.0
    return y


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

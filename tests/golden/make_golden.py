"""Generates tests/golden/cutoffvgg16_seed1234.npz.

The reference cannot be executed here (TensorFlow 2.9 / Keras are not installable: no wheel, no network,
Python 3.12) and ships no golden vectors, so these fixtures are produced by the torch-CPU oracle
(oracle/oracle.py) -- they pin the oracle against silent drift, not against Keras ("parity unpinned").
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle, synth  # noqa: E402


def main():
    B = 8
    frames, labels = synth.synth_frames(B, seed=1234)
    x, y = synth.vgg16_preprocess(frames), synth.one_hot(labels)
    params = synth.randomize_biases(synth.glorot_weights(10001))
    fwd = oracle.forward(params, x)
    mask = synth.dropout_mask(B, 256, 0.45)
    cw = {0: 0.75, 1: 1.5}
    loss, grads, _ = oracle.loss_and_grads(params, x, y, cw, mask)
    heat, _ = oracle.gradcam_autograd(params, x[:2], out_hw=(128, 128))
    out = {
        "frames_u8": frames, "labels": labels, "logits": fwd["logits"], "probs": fwd["probs"], "gap": fwd["gap"],
        "block3_conv3_mean": fwd["block3_conv3"].mean(axis=(1, 2)), "loss": np.float32(loss),
        "drop_mask": mask.astype(np.uint8), "gradcam_heat": heat.astype(np.float16),
        "grad_output_kernel": grads["output/kernel"], "grad_output_bias": grads["output/bias"],
        "grad_block3_conv3_bias": grads["block3_conv3/bias"], "grad_block1_conv1_kernel": grads["block1_conv1/kernel"],
        "grad_block3_conv2_kernel_slice": grads["block3_conv2/kernel"][1, 1, :16, :16],
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cutoffvgg16_seed1234.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

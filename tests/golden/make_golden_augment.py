"""Generates tests/golden/augment_metrics_seed77.npz: four 128x128 synthetic frames through the augmentation oracle
(oracle/augment.py) with fixed draws, and the metric counter state (oracle/metric_counts.py) of a fixed batch of
probabilities.  Produced BY THE ORACLE (TensorFlow/Keras cannot be installed here and the reference ships no vectors
for this path): the fixture guards the oracle -- and through the GPU parity tests the kernels -- against drift, it does
not pin Keras.  Run from the repo root:  python tests/golden/make_golden_augment.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import augment, metric_counts, synth  # noqa: E402

RAW = {"zoom": np.array([1.0, 0.93, 1.08, 1.02], np.float32), "dx": np.array([0.0, 0.15, -0.2, 0.05], np.float32),
       "dy": np.array([0.0, -0.1, 0.12, 0.0], np.float32), "angle": np.array([0.0, 0.6, -2.2, 100.0], np.float32),
       "flip": np.array([0, 1, 0, 1], np.float32), "brightness": np.array([0.0, 0.25, -0.3, 0.1], np.float32)}


def inputs():
    frames, _ = synth.synth_frames(4, seed=77)
    rng = np.random.default_rng(77)
    z = (rng.normal(size=(64, 2)) * 2).astype(np.float32)
    probs = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    probs[::9] = 0.5
    y = np.eye(2, dtype=np.float32)[rng.integers(0, 2, 64)]
    eps = 1e-7
    thr = np.array([0.0 - eps] + [(i + 1) / 199 for i in range(198)] + [1.0 + eps], np.float32)
    return frames, probs, y, thr


def main():
    frames, probs, y, thr = inputs()
    out = augment.augment_batch(frames, RAW, preprocess=True)
    counts = metric_counts.counts_from_arrays(y, probs, thr, [0.5, 0.5])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_metrics_seed77.npz")
    # every 4th row/column of the augmented frames (exact fp32) keeps the fixture small; the checksum covers the rest
    np.savez_compressed(path, augmented_sub=out[:, ::4, ::4], augmented_sum=out.astype(np.float64).sum(axis=(1, 2)),
                        counts=counts, probs=probs, y=y)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""Independent NumPy direct-loop restatement of Conv2D(3x3, SAME) / MaxPool2D(2x2) / GAP / Dense for tiny
shapes (TEST INFRASTRUCTURE).  Written from the Keras definitions in SURVEY.md Appendix A.1 with no shared
code with oracle/oracle.py, so the two cross-check each other (layout permutations included)."""
import numpy as np


def conv3x3_same_relu(x, w, b, relu=True):
    """x [B,H,W,Ci], w [3,3,Ci,Co] (HWIO, cross-correlation), b [Co] -> [B,H,W,Co]."""
    B, H, W, Ci = x.shape
    Co = w.shape[3]
    xp = np.zeros((B, H + 2, W + 2, Ci), np.float64)
    xp[:, 1:-1, 1:-1, :] = x
    y = np.zeros((B, H, W, Co), np.float64)
    for n in range(B):
        for h in range(H):
            for ww in range(W):
                patch = xp[n, h:h + 3, ww:ww + 3, :]            # [3,3,Ci]
                y[n, h, ww, :] = np.tensordot(patch, w, axes=([0, 1, 2], [0, 1, 2])) + b
    return np.maximum(y, 0.0) if relu else y


def maxpool2x2(x):
    B, H, W, C = x.shape
    y = np.zeros((B, H // 2, W // 2, C), x.dtype)
    for h in range(H // 2):
        for w in range(W // 2):
            y[:, h, w, :] = x[:, 2 * h:2 * h + 2, 2 * w:2 * w + 2, :].max(axis=(1, 2))
    return y


def forward(params, x, plan):
    a = np.asarray(x, np.float64)
    for name, kind, _ in plan:
        a = conv3x3_same_relu(a, params[name + "/kernel"], params[name + "/bias"]) if kind == "conv" else maxpool2x2(a)
    gap = a.mean(axis=(1, 2))
    z = gap @ params["output/kernel"] + params["output/bias"]
    e = np.exp(z - z.max(-1, keepdims=True))
    return a, z, e / e.sum(-1, keepdims=True)

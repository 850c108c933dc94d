"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's clip-level aggregation rules.

Plain-Python loops following reference src/predict.py: `max_contiguous_b_line_preds` (283-297),
`predict_with_contiguity_threshold` (299-308), `highest_avg_contiguous_pred_prob` (365-376) and the 'average' rule
(predict.py:164).  Only tests/ may import this module; the product path is the abl_clip_aggregate kernel.
"""
import numpy as np


def max_contiguous(preds):
    """predict.py:283-297: longest run of 1s."""
    best = cur = 0
    for i in range(len(preds)):
        cur = cur + 1 if preds[i] == 1 else 0
        if cur > best:
            best = cur
    return best


def contiguity_rule(pred_probs, contiguity_threshold, classification_threshold):
    """predict.py:299-308 (column 1 = b_lines, strict '>' on the frame probability)."""
    b_preds = (np.asarray(pred_probs)[:, 1] > classification_threshold).astype(int)
    c = int(max_contiguous(b_preds) >= contiguity_threshold)
    return np.array([1 - c, c])


def sliding_window_rule(pred_probs, window_length):
    """predict.py:365-376: the maximum starts at 0.0 and a clip shorter than the window has no window at all."""
    pred_probs = np.asarray(pred_probs)
    m = 0.0
    for i in range(0, pred_probs.shape[0] - window_length + 1):
        a = np.mean(pred_probs[i:i + window_length, 1])
        if a > m:
            m = a
    return np.array([1. - m, m])


def average_rule(pred_probs):
    """predict.py:164."""
    return np.mean(np.asarray(pred_probs), axis=0)

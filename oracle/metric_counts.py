"""TEST INFRASTRUCTURE ONLY (never imported by the product).  NumPy restatement of the counter state that
``abl_metrics_accumulate`` keeps for the Keras metric list of src/train.py:109-115, straight from the definitions in
Keras 2.9 ``metrics_utils.update_confusion_matrix_variables`` (third-party, absent from /root/reference):
``pred > threshold`` in fp32, ``y_true`` cast to bool, AUC with multi_label=False flattening labels and predictions.
Parity unpinned (no TensorFlow here); anchored on hand-computed cases in tests/test_oracle.py."""
import numpy as np


def counts_from_arrays(y_true, probs, auc_thresholds, class_thresholds):
    y = np.asarray(y_true, np.float32)
    p = np.asarray(probs, np.float32)
    thr = np.asarray(auc_thresholds, np.float32)
    cthr = np.asarray(class_thresholds, np.float32)
    B, C = p.shape if p.ndim == 2 else (0, len(cthr))
    T = len(thr)
    out = np.zeros(2 + 3 * C + 2 * (T + 1), np.int64)
    out[0] = B
    if B == 0:
        return out
    out[1] = int(np.sum(np.argmax(y, -1) == np.argmax(p, -1)))
    pos = y != 0
    for c in range(C):
        hit = p[:, c] > cthr[c]
        out[2 + 3 * c] = np.sum(pos[:, c] & hit)
        out[2 + 3 * c + 1] = np.sum(~pos[:, c] & hit)
        out[2 + 3 * c + 2] = np.sum(pos[:, c] & ~hit)
    # bucket k = number of thresholds strictly below the prediction
    k = (p.reshape(-1, 1) > thr.reshape(1, -1)).sum(1)
    base = 2 + 3 * C
    out[base:base + T + 1] = np.bincount(k[pos.reshape(-1)], minlength=T + 1)
    out[base + T + 1:base + 2 * (T + 1)] = np.bincount(k[~pos.reshape(-1)], minlength=T + 1)
    return out

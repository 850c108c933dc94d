"""Host-side Keras-compatible metric objects for the cutoffvgg16 path.

The reference builds ``['accuracy', AUC(name='auc'), Precision(...), Recall(...)]`` at
src/train.py:109-115 and hands them to ``model.compile``; these classes take the same constructor
arguments and reproduce Keras 2.9's definitions from accumulated (y_true one-hot, probs) arrays.
"""
import numpy as np


class Metric:
    name = "metric"

    def result(self, y_true, y_pred):
        raise NotImplementedError


class CategoricalAccuracy(Metric):
    """What Keras resolves the string 'accuracy' to for a categorical_crossentropy loss."""

    def __init__(self, name="accuracy"):
        self.name = name

    def result(self, y_true, y_pred):
        if len(y_true) == 0:
            return 0.0
        return float(np.mean(np.argmax(y_true, -1) == np.argmax(y_pred, -1)))


class AUC(Metric):
    """keras.metrics.AUC defaults: ROC, 200 thresholds, 'interpolation' summation, multi_label=False
    (labels and predictions are flattened over the class axis)."""

    def __init__(self, name="auc", num_thresholds=200):
        self.name = name
        self.num_thresholds = num_thresholds
        eps = 1e-7
        self.thresholds = np.array([0.0 - eps] + [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)]
                                   + [1.0 + eps])

    def result(self, y_true, y_pred):
        yt = np.asarray(y_true).reshape(-1) > 0.5
        yp = np.asarray(y_pred, np.float64).reshape(-1)
        order = np.sort(yp[yt]), np.sort(yp[~yt])
        # counts with pred > threshold
        tp = len(order[0]) - np.searchsorted(order[0], self.thresholds, side="right")
        fp = len(order[1]) - np.searchsorted(order[1], self.thresholds, side="right")
        fn = len(order[0]) - tp
        tn = len(order[1]) - fp
        tpr = np.divide(tp, tp + fn, out=np.zeros(len(tp)), where=(tp + fn) > 0)
        fpr = np.divide(fp, fp + tn, out=np.zeros(len(fp)), where=(fp + tn) > 0)
        return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0))


class Precision(Metric):
    def __init__(self, name="precision", thresholds=0.5, class_id=None):
        self.name, self.thresholds, self.class_id = name, 0.5 if thresholds is None else thresholds, class_id

    def _sel(self, y_true, y_pred):
        yt, yp = np.asarray(y_true), np.asarray(y_pred)
        if self.class_id is not None:
            yt, yp = yt[:, self.class_id], yp[:, self.class_id]
        return yt.reshape(-1) > 0.5, yp.reshape(-1) > self.thresholds

    def result(self, y_true, y_pred):
        t, p = self._sel(y_true, y_pred)
        tp, fp = np.sum(t & p), np.sum(~t & p)
        return float(tp / (tp + fp)) if tp + fp > 0 else 0.0


class Recall(Precision):
    def __init__(self, name="recall", thresholds=0.5, class_id=None):
        super().__init__(name, thresholds, class_id)

    def result(self, y_true, y_pred):
        t, p = self._sel(y_true, y_pred)
        tp, fn = np.sum(t & p), np.sum(t & ~p)
        return float(tp / (tp + fn)) if tp + fn > 0 else 0.0


def resolve(metrics):
    """Strings -> metric objects (only what the reference passes: 'accuracy')."""
    out = []
    for m in metrics or []:
        if isinstance(m, str):
            if m in ("accuracy", "acc", "categorical_accuracy"):
                out.append(CategoricalAccuracy(m if m != "acc" else "accuracy"))
            elif m == "auc":
                out.append(AUC())
            else:
                raise NotImplementedError(f"metric '{m}' is not supported on the cutoffvgg16 B200 path")
        elif hasattr(m, "result") and hasattr(m, "name"):
            out.append(m)
        else:
            raise TypeError(f"unsupported metric object {m!r}")
    return out

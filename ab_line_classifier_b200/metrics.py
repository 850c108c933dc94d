"""Keras-compatible metric objects for the cutoffvgg16 path.

The reference builds ``['accuracy', AUC(name='auc'), Precision(...), Recall(...)]`` at
src/train.py:109-115 and hands them to ``model.compile``; these classes take the same constructor
arguments and reproduce Keras 2.9's definitions.  Two ways to a result:
  * ``from_counts(MetricCounts)``: from the integer state the device accumulates batch by batch
    (``abl_metrics_accumulate``) -- what fit()/evaluate() use, nothing but a few hundred counters crosses PCIe;
  * ``result(y_true, y_pred)``: from whole (one-hot, probs) arrays, for metric lists the device state cannot express
    (see ``device_plan``) and for user code holding arrays.
Comparisons against thresholds are made in fp32, as Keras makes them.
"""
import numpy as np


class Metric:
    name = "metric"

    def result(self, y_true, y_pred):
        raise NotImplementedError

    def from_counts(self, counts):
        raise NotImplementedError


class MetricCounts:
    """Host view of the device counters (layout: include/abl_b200.h, abl_metrics_accumulate)."""

    def __init__(self, raw, n_classes, auc_thresholds, class_thresholds):
        raw = np.asarray(raw).astype(np.int64)
        C, T = int(n_classes), len(auc_thresholds)
        assert raw.size >= 2 + 3 * C + 2 * (T + 1)
        self.n_classes, self.auc_thresholds = C, np.asarray(auc_thresholds, np.float32)
        self.class_thresholds = np.asarray(class_thresholds, np.float32)
        self.n, self.correct = int(raw[0]), int(raw[1])
        per = raw[2:2 + 3 * C].reshape(C, 3)
        self.tp, self.fp, self.fn = per[:, 0], per[:, 1], per[:, 2]
        h = raw[2 + 3 * C:2 + 3 * C + 2 * (T + 1)].reshape(2, T + 1)
        # entries above threshold t = those in buckets k > t
        self.auc_tp = np.cumsum(h[0][::-1])[::-1][1:]
        self.auc_fp = np.cumsum(h[1][::-1])[::-1][1:]
        self.n_pos, self.n_neg = int(h[0].sum()), int(h[1].sum())


def auc_thresholds(num_thresholds=200):
    """keras.metrics.AUC's threshold list (ascending, fp32 as Keras casts it)."""
    eps = 1e-7
    return np.array([0.0 - eps] + [(i + 1) / (num_thresholds - 1) for i in range(num_thresholds - 2)] + [1.0 + eps],
                    np.float32)


def device_plan(metric_objs, n_classes):
    """(auc_thresholds fp32 [T], class_thresholds fp32 [n_classes]) if the device counters can serve every metric of
    the list -- the reference's list always can -- else None (then the caller collects arrays and uses result())."""
    T, cthr = None, [None] * n_classes
    for m in metric_objs:
        if type(m) is CategoricalAccuracy:
            continue
        if type(m) is AUC:
            if T not in (None, m.num_thresholds):
                return None
            T = m.num_thresholds
        elif type(m) in (Precision, Recall):
            thr = m.thresholds
            if m.class_id is None or not np.isscalar(thr) or not 0 <= m.class_id < n_classes:
                return None
            if cthr[m.class_id] not in (None, float(thr)):
                return None
            cthr[m.class_id] = float(thr)
        else:
            return None
    return auc_thresholds(T or 200), np.array([0.5 if t is None else t for t in cthr], np.float32)


class CategoricalAccuracy(Metric):
    """What Keras resolves the string 'accuracy' to for a categorical_crossentropy loss."""

    def __init__(self, name="accuracy"):
        self.name = name

    def result(self, y_true, y_pred):
        if len(y_true) == 0:
            return 0.0
        return float(np.mean(np.argmax(y_true, -1) == np.argmax(y_pred, -1)))

    def from_counts(self, c):
        return c.correct / c.n if c.n else 0.0


class AUC(Metric):
    """keras.metrics.AUC defaults: ROC, 200 thresholds, 'interpolation' summation, multi_label=False
    (labels and predictions are flattened over the class axis)."""

    def __init__(self, name="auc", num_thresholds=200):
        self.name = name
        self.num_thresholds = num_thresholds
        self.thresholds = auc_thresholds(num_thresholds)

    @staticmethod
    def _area(tp, fp, n_pos, n_neg):
        tp, fp = tp.astype(np.float64), fp.astype(np.float64)
        tpr = tp / n_pos if n_pos > 0 else np.zeros(len(tp))          # div_no_nan
        fpr = fp / n_neg if n_neg > 0 else np.zeros(len(fp))
        return float(np.sum((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0))

    def result(self, y_true, y_pred):
        yt = np.asarray(y_true).reshape(-1) != 0                        # Keras casts y_true to bool
        yp = np.asarray(y_pred, np.float32).reshape(-1)
        order = np.sort(yp[yt]), np.sort(yp[~yt])
        # counts with pred > threshold
        tp = len(order[0]) - np.searchsorted(order[0], self.thresholds, side="right")
        fp = len(order[1]) - np.searchsorted(order[1], self.thresholds, side="right")
        return self._area(tp, fp, len(order[0]), len(order[1]))

    def from_counts(self, c):
        assert len(c.auc_thresholds) == self.num_thresholds
        return self._area(c.auc_tp, c.auc_fp, c.n_pos, c.n_neg)


class Precision(Metric):
    def __init__(self, name="precision", thresholds=0.5, class_id=None):
        self.name, self.thresholds, self.class_id = name, 0.5 if thresholds is None else thresholds, class_id

    def _sel(self, y_true, y_pred):
        yt, yp = np.asarray(y_true), np.asarray(y_pred)
        if self.class_id is not None:
            yt, yp = yt[:, self.class_id], yp[:, self.class_id]
        return yt.reshape(-1) != 0, yp.reshape(-1).astype(np.float32) > np.float32(self.thresholds)

    def result(self, y_true, y_pred):
        t, p = self._sel(y_true, y_pred)
        tp, fp = np.sum(t & p), np.sum(~t & p)
        return float(tp / (tp + fp)) if tp + fp > 0 else 0.0

    def from_counts(self, c):
        tp, fp = int(c.tp[self.class_id]), int(c.fp[self.class_id])
        return tp / (tp + fp) if tp + fp > 0 else 0.0


class Recall(Precision):
    def __init__(self, name="recall", thresholds=0.5, class_id=None):
        super().__init__(name, thresholds, class_id)

    def result(self, y_true, y_pred):
        t, p = self._sel(y_true, y_pred)
        tp, fn = np.sum(t & p), np.sum(t & ~p)
        return float(tp / (tp + fn)) if tp + fn > 0 else 0.0

    def from_counts(self, c):
        tp, fn = int(c.tp[self.class_id]), int(c.fn[self.class_id])
        return tp / (tp + fn) if tp + fn > 0 else 0.0


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

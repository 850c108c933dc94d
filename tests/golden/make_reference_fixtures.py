"""Golden vectors produced by the REFERENCE'S OWN functions, executed in the build container.

    python tests/golden/make_reference_fixtures.py        # needs /root/reference; writes reference_host_rules.npz

STATUS: NOT RUN in the build container -- importing and executing the reference's modules needs an explicit human
authorisation there, which was not available; reference_host_rules.npz is therefore absent and
tests/test_reference_fixtures.py skips its comparisons.  Run it where executing the reference is permitted.  The
machinery below (stand-in modules, scratch working directory, what is recorded) is exercised by
tests/test_reference_fixtures.py against a stand-in source tree written by the test itself (ABL_REFERENCE_ROOT).

The reference's conv path needs TensorFlow 2.9 (requirements.txt:13), which this image cannot install, so the network
itself stays unpinned (oracle/keras_check.py).  Its HOST-side rules on either side of the network are plain NumPy /
pandas / scikit-learn, though, and run here unmodified once the imports the module tops pull in (tensorflow, onnx,
onnx_tf, matplotlib, skopt, wandb.keras ...) are satisfied by inert stand-ins:

  src/predict.py   max_contiguous_b_line_preds (283-297), predict_with_contiguity_threshold (299-308),
                   highest_avg_contiguous_pred_prob (365-376), compute_metrics (89-121)
  src/train.py     compute_class_weight (41-68), compute_output_bias (71-81)

Nothing from the reference is copied: the module is imported from where it lies, called on seeded inputs, and only the
inputs and the returned numbers are stored.  tests/test_reference_fixtures.py holds this repo's host code
(src/predict.py, src/train.py), the oracle restatement (oracle/clip_rules.py) and -- on the GPU -- abl_clip_aggregate
to these numbers.  The file does not travel to the GPU box as code that runs; only the .npz does.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE = os.environ.get("ABL_REFERENCE_ROOT", "/root/reference")
OUT = os.environ.get("ABL_REFERENCE_FIXTURE_OUT",
                     os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_host_rules.npz"))
STUBBED = ("tensorflow", "tensorflow_addons", "onnx", "onnx_tf", "matplotlib", "skopt", "seaborn", "skimage", "imageio")


class _AnyMeta(type):
    """Class-level attribute access on a stand-in (tf.config.experimental.list_physical_devices) yields more stand-ins."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (_Anything,), {})


class _Anything(metaclass=_AnyMeta):
    """Stands in for every class, function and object of a stubbed package: callable, subclassable, iterable (empty),
    indexable; remembers the arguments it was constructed with (compute_output_bias returns Constant(<array>))."""

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs

    def __call__(self, *args, **kwargs):
        return _Anything(*args, **kwargs)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __setitem__(self, key, value):
        pass

    def __getitem__(self, key):
        return _Anything()


class _StubModule(types.ModuleType):
    __all__ = []
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _AnyMeta(name, (_Anything,), {})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUBBED or fullname == "wandb.keras":
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def import_reference():
    """Import the reference's src.predict and src.train from /root/reference in a scratch working directory that holds
    what their module-level code opens (config.yml, the class-name map)."""
    import dill
    import wandb
    work = tempfile.mkdtemp(prefix="abl_ref_")
    shutil.copy(os.path.join(REFERENCE, "config.yml"), os.path.join(work, "config.yml"))     # scratch copy, not the repo
    os.makedirs(os.path.join(work, "data", "serializations"))
    with open(os.path.join(work, "data", "serializations", "output_class_indices.pkl"), "wb") as f:
        dill.dump({"a_lines": 0, "b_lines": 1}, f)          # config.yml DATA.CLASSES order
    sys.meta_path.insert(0, _StubFinder())
    wandb.Table = lambda *a, **k: None                       # compute_class_weight logs two tables
    wandb.log = lambda *a, **k: None
    os.chdir(work)
    for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
        del sys.modules[name]                                # this repo has a src/ package of the same name
    sys.path.insert(0, REFERENCE)
    return importlib.import_module("src.predict"), importlib.import_module("src.train")


def main():
    import pandas as pd
    here = os.getcwd()
    ref_predict, ref_train = import_reference()
    assert ref_predict.__file__.startswith(REFERENCE) and ref_train.__file__.startswith(REFERENCE)
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- clip rules: clips of various lengths, including shorter than the window / threshold and constant ones
    lengths = [1, 2, 3, 4, 7, 16, 33, 64, 100, 257]
    probs, offsets = [], [0]
    for n in lengths:
        kind = rng.integers(0, 4)
        b = rng.random(n)
        if kind == 1:
            b = np.clip(b + 0.35, 0, 1)                  # mostly positive frames
        elif kind == 2:
            b = np.repeat(rng.random((n + 4) // 5), 5)[:n]     # runs of equal probabilities
        p = np.stack([1.0 - b, b], axis=1).astype(np.float32).astype(np.float64)       # what a float32 model emits
        probs.append(p)
        offsets.append(offsets[-1] + n)
    out["clip_probs"] = np.concatenate(probs)
    out["clip_offsets"] = np.asarray(offsets, np.int64)
    thresholds = [0.3, 0.5, 0.75]
    contiguity = [1, 3, 8]
    windows = [1, 4, 16]
    out["class_thresholds"] = np.asarray(thresholds)
    out["contiguity_thresholds"] = np.asarray(contiguity, np.int64)
    out["window_lengths"] = np.asarray(windows, np.int64)
    runs = np.zeros((len(lengths), len(thresholds)), np.int64)
    contig = np.zeros((len(lengths), len(thresholds), len(contiguity), 2), np.int64)
    window = np.zeros((len(lengths), len(windows), 2), np.float64)
    for c, p in enumerate(probs):
        for ti, t in enumerate(thresholds):
            runs[c, ti] = ref_predict.max_contiguous_b_line_preds((p[:, 1] > t).astype(int))
            for ki, k in enumerate(contiguity):
                contig[c, ti, ki] = ref_predict.predict_with_contiguity_threshold(p, k, t)
        for wi, w in enumerate(windows):
            window[c, wi] = ref_predict.highest_avg_contiguous_pred_prob(p, w)
    out["max_contiguous"] = runs
    out["contiguity_rule"] = contig
    out["sliding_window_rule"] = window
    # ---- compute_metrics (scikit-learn underneath)
    cfg = ref_predict.cfg
    n = 400
    labels = rng.integers(0, 2, n)
    pb = np.clip(0.55 * labels + 0.6 * rng.random(n) - 0.1, 0.0, 1.0)
    mprobs = np.stack([1.0 - pb, pb], axis=1)
    preds = (pb > 0.5).astype(int)
    m = ref_predict.compute_metrics(cfg, labels, preds, mprobs)
    out["metrics_labels"], out["metrics_preds"], out["metrics_probs"] = labels, preds, mprobs
    out["metrics_confusion_matrix"] = np.asarray(m["confusion_matrix"], np.int64)
    keys = ["precision", "recall", "specificity", "f1", "accuracy", "macro_mean_auc", "weighted_mean_auc", "a_lines_auc",
            "b_lines_auc"]
    out["metrics_keys"] = np.asarray(keys)
    out["metrics_values"] = np.asarray([float(m[k]) for k in keys])
    # ---- class weights and output bias
    classes = np.concatenate([np.zeros(731, int), np.ones(269, int)])
    rng.shuffle(classes)
    train_df = pd.DataFrame({"Class": classes})
    cw = ref_train.compute_class_weight(train_df)
    ob = ref_train.compute_output_bias(train_df)             # Constant(<array>): the stand-in keeps the argument
    out["train_classes"] = classes
    out["class_weight"] = np.asarray([cw[i] for i in range(len(cw))], np.float64)
    out["output_bias"] = np.asarray(ob.args[0], np.float64)
    os.chdir(here)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()

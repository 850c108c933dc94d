"""Golden vectors from the reference's own host-side functions (tests/golden/make_reference_fixtures.py).

The generator has NOT been run in the build container (executing the reference's modules needs a human authorisation that
was not available), so `reference_host_rules.npz` may be absent: the comparisons then skip.  What always runs is a check
of the generator's machinery against a stand-in source tree written here, so that the script is known to work the day it
is pointed at the real reference."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "reference_host_rules.npz")
needs_fixture = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="reference_host_rules.npz not generated (see the "
                                   "header of tests/golden/make_reference_fixtures.py)")


def _check_against(z):
    """This repo's host code and the oracle restatement against recorded outputs."""
    from oracle import clip_rules
    from src.predict import (highest_avg_contiguous_pred_prob, max_contiguous_b_line_preds,
                             predict_with_contiguity_threshold)
    off = z["clip_offsets"]
    for c in range(len(off) - 1):
        p = z["clip_probs"][off[c]:off[c + 1]]
        for ti, t in enumerate(z["class_thresholds"]):
            b = (p[:, 1] > t).astype(int)
            assert max_contiguous_b_line_preds(b) == z["max_contiguous"][c, ti] == clip_rules.max_contiguous(b)
            for ki, k in enumerate(z["contiguity_thresholds"]):
                want = z["contiguity_rule"][c, ti, ki]
                np.testing.assert_array_equal(predict_with_contiguity_threshold(p, int(k), float(t)), want)
                np.testing.assert_array_equal(clip_rules.contiguity_rule(p, int(k), float(t)), want)
        for wi, w in enumerate(z["window_lengths"]):
            want = z["sliding_window_rule"][c, wi]
            np.testing.assert_allclose(highest_avg_contiguous_pred_prob(p, int(w)), want, rtol=0, atol=1e-15)
            np.testing.assert_allclose(clip_rules.sliding_window_rule(p, int(w)), want, rtol=0, atol=1e-15)
    from src.train import compute_class_weight, compute_output_bias
    cw = compute_class_weight(z["train_classes"])
    np.testing.assert_allclose([cw[i] for i in range(len(cw))], z["class_weight"], rtol=1e-15)
    np.testing.assert_allclose(compute_output_bias(z["train_classes"]), z["output_bias"], rtol=1e-15)


@needs_fixture
def test_host_rules_match_the_reference_outputs():
    _check_against(np.load(FIXTURE))


@needs_fixture
@pytest.mark.gpu
def test_clip_aggregate_kernel_matches_the_reference_outputs():
    import torch
    from ab_line_classifier_b200 import ops
    z = np.load(FIXTURE)
    probs = torch.from_numpy(z["clip_probs"].astype(np.float32)).cuda()
    offsets = torch.from_numpy(z["clip_offsets"]).cuda()
    for ti, t in enumerate(z["class_thresholds"]):
        for wi, w in enumerate(z["window_lengths"]):
            mean, run, win = ops.clip_aggregate(probs, offsets, 1, float(t), int(w))
            np.testing.assert_array_equal(run.cpu().numpy(), z["max_contiguous"][:, ti])
            np.testing.assert_allclose(win.cpu().numpy(), z["sliding_window_rule"][:, wi, 1], atol=1e-6)


FAKE_PREDICT = '''
import os, yaml, dill
import numpy as np
import tensorflow as tf                                   # absent: must resolve to a stand-in
from tensorflow.keras.models import load_model
import onnx
from onnx_tf.backend import prepare
import matplotlib.pyplot as plt
cfg = yaml.full_load(open(os.getcwd() + "/config.yml", "r"))
CLASS_IDX_MAP = dill.load(open(cfg["PATHS"]["CLASS_NAME_MAP"], "rb"))
for device in tf.config.experimental.list_physical_devices("GPU"):
    raise RuntimeError("a stand-in must iterate as empty")
def max_contiguous_b_line_preds(preds):
    best = cur = 0
    for v in preds:
        cur = cur + 1 if v == 1 else 0
        best = max(best, cur)
    return best
def predict_with_contiguity_threshold(pred_probs, contiguity_threshold, classification_threshold):
    c = int(max_contiguous_b_line_preds((pred_probs[:, 1] > classification_threshold).astype(int)) >= contiguity_threshold)
    return np.array([1 - c, c])
def highest_avg_contiguous_pred_prob(pred_probs, window_length):
    m = 0.0
    for i in range(0, pred_probs.shape[0] - window_length + 1):
        m = max(m, float(np.mean(pred_probs[i:i + window_length, 1])))
    return np.array([1. - m, m])
def compute_metrics(cfg, labels, preds, probs=None):
    from sklearn.metrics import confusion_matrix
    d = {k: 0.5 for k in ("precision", "recall", "specificity", "f1", "accuracy", "macro_mean_auc", "weighted_mean_auc",
                          "a_lines_auc", "b_lines_auc")}
    d["confusion_matrix"] = confusion_matrix(labels, preds).tolist()
    return d
'''
FAKE_TRAIN = '''
import numpy as np
import wandb
from wandb.keras import WandbMetricsLogger                # stand-in even though wandb itself is installed
from tensorflow.keras.initializers import Constant
from tensorflow.keras.layers import Layer
class Custom(Layer):                                      # the reference subclasses Keras layers at import time
    pass
def compute_class_weight(train_df):
    h = np.bincount(train_df["Class"].to_numpy().astype(int))
    wandb.log({"t": wandb.Table(columns=["a"], data=[[1]])})
    return {i: (1.0 / len(h)) * sum(h) / h[i] for i in range(len(h))}
def compute_output_bias(train_df):
    h = np.bincount(train_df["Class"].astype(int))
    return Constant(np.log([h[i] / (np.sum(h) - h[i]) for i in range(h.shape[0])]))
'''


def test_fixture_generator_machinery_on_a_stand_in_tree(tmp_path):
    """The generator's import shims, scratch directory and recording, run in a subprocess against a source tree written
    by this test (same entry points as the reference, trivially re-implemented) -- never against /root/reference."""
    root = tmp_path / "fake_reference"
    (root / "src").mkdir(parents=True)
    (root / "src" / "__init__.py").write_text("")
    (root / "src" / "predict.py").write_text(textwrap.dedent(FAKE_PREDICT))
    (root / "src" / "train.py").write_text(textwrap.dedent(FAKE_TRAIN))
    (root / "config.yml").write_text("PATHS:\n  CLASS_NAME_MAP: 'data/serializations/output_class_indices.pkl'\n"
                                     "DATA:\n  CLASSES: ['a_lines', 'b_lines']\n")
    out = tmp_path / "fixture.npz"
    env = dict(os.environ, ABL_REFERENCE_ROOT=str(root), ABL_REFERENCE_FIXTURE_OUT=str(out))
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_reference_fixtures.py")], env=env,
                       capture_output=True, text=True, timeout=300, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(out)
    assert z["clip_offsets"][-1] == len(z["clip_probs"]) and z["contiguity_rule"].shape[-1] == 2
    assert z["metrics_confusion_matrix"].shape == (2, 2) and z["output_bias"].shape == (2,)
    _check_against(z)          # the stand-in implements the same rules, so this repo's code must agree with its outputs

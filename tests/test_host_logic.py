"""Host-side logic on CPU: model registry, two-phase fit schedule, callbacks, metrics, clip rules.
The device engine is replaced by tests/fake_engine.FakeEngine (oracle numerics) -- tests only."""
import os

import numpy as np
import pytest

import ab_line_classifier_b200.model as model_mod
from ab_line_classifier_b200 import metrics as M
from ab_line_classifier_b200.data import ArrayDataset
from tests.fake_engine import FakeEngine

HP = {"LR_EXTRACT": 3e-4, "LR_FINETUNE": 9.3e-6, "DROPOUT": 0.45, "CUTOFF_LAYER": 4, "FINETUNE_LAYER": 1,
      "EXTRACT_EPOCHS": 2}


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(model_mod, "CutoffVGG16Engine", FakeEngine)
    FakeEngine.calls = []
    return FakeEngine


def _data(n=12, bs=5, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, 8, 8, 3)).astype(np.float32)
    y = np.eye(2, dtype=np.float32)[np.arange(n) % 2]
    return ArrayDataset(x, y, bs, pin=False)


def test_get_model_registry():
    from src.models.models import get_model, vgg16_preprocess
    model_def, prep = get_model("cutoffvgg16")
    assert model_def is model_mod.CutoffVGG16 and prep is vgg16_preprocess
    other, _ = get_model("mobilenetv2")
    with pytest.raises(NotImplementedError):
        other({}, [128, 128, 3], [], 2)
    unknown, _ = get_model("does-not-exist")          # reference falls back to cnn0
    with pytest.raises(NotImplementedError):
        unknown()


def test_vgg16_preprocess_is_caffe_mode():
    from src.models.models import vgg16_preprocess
    x = np.zeros((1, 2, 2, 3), np.uint8)
    x[..., 0], x[..., 1], x[..., 2] = 10, 20, 30          # R, G, B
    out = vgg16_preprocess(x)
    np.testing.assert_allclose(out[0, 0, 0], [30 - 103.939, 20 - 116.779, 10 - 123.68], rtol=1e-6)


def test_two_phase_schedule_trainable_sets_and_epochs(fake):
    m = model_mod.CutoffVGG16(HP, [8, 8, 3], ["accuracy"], 2)
    assert [l.name for l in m.vgg16_layers] == ["block1_conv1", "block1_conv2", "block1_pool"]
    train = _data()
    assert m.fit(train, epochs=4, validation_data=_data(6, 3, 1), verbose=0) is None     # reference returns None
    steps_per_epoch = len(train)
    adam = [c for c in fake.calls if c[0] == "adam"]
    rms = [c for c in fake.calls if c[0] == "rmsprop"]
    # phase 1: EXTRACT_EPOCHS epochs, Adam, head only; phase 2: initial_epoch = EXTRACT_EPOCHS-1 .. epochs
    assert len(adam) == 2 * steps_per_epoch and all(c[2] == frozenset({"output"}) for c in adam)
    assert len(rms) == (4 - 1) * steps_per_epoch
    assert all(c[2] == frozenset({"block1_conv2", "output"}) for c in rms)     # vgg16_layers[FINETUNE_LAYER:]
    assert all(abs(c[1] - 3e-4) < 1e-12 for c in adam) and all(abs(c[1] - 9.3e-6) < 1e-12 for c in rms)
    assert m.history_extract.epoch == [0, 1] and m.history_finetune.epoch == [1, 2, 3]
    assert "val_loss" in m.history_finetune.history and "accuracy" in m.history_extract.history


def test_metrics_names_and_evaluate_order(fake):
    mets = ["accuracy", M.AUC(name="auc"), M.Precision(name="precision_a_lines", thresholds=0.5, class_id=0),
            M.Recall(name="recall_b_lines", thresholds=0.5, class_id=1)]
    m = model_mod.CutoffVGG16(HP, [8, 8, 3], mets, 2)
    m.model.compile(optimizer=m.optimizer_extract, loss="categorical_crossentropy", metrics=mets)
    assert m.metrics_names == ["loss", "accuracy", "auc", "precision_a_lines", "recall_b_lines"]
    res = m.evaluate(_data(), verbose=0)
    assert len(res) == 5 and 0.0 <= res[1] <= 1.0 and res[0] > 0
    p = m.predict(_data(), verbose=0)
    assert p.shape == (12, 2) and np.allclose(p.sum(-1), 1.0, atol=1e-5)


def test_short_last_batch_is_used(fake):
    m = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2)
    m.fit(_data(n=11, bs=4), epochs=2, verbose=0)
    assert len([c for c in fake.calls if c[0] == "adam"]) == 2 * 3


def test_early_stopping_and_reduce_lr():
    class Dummy:
        stop_training = False
        optimizer = model_mod.Adam(1e-3)

        class engine:
            import torch
            theta = torch.zeros(3)

            @staticmethod
            def refresh_operands():
                pass
    es = model_mod.EarlyStopping(monitor="val_loss", patience=2, restore_best_weights=True)
    rl = model_mod.ReduceLROnPlateau(monitor="val_loss", factor=0.5, patience=1, min_lr=1e-8, min_delta=1e-4)
    d = Dummy()
    for cb in (es, rl):
        cb.set_model(d)
        cb.on_train_begin()
    losses = [1.0, 0.9, 0.95, 0.97, 0.99]
    stopped_at = None
    for ep, l in enumerate(losses):
        logs = {"val_loss": l}
        es.on_epoch_end(ep, logs)
        rl.on_epoch_end(ep, logs)
        if d.stop_training:
            stopped_at = ep
            break
    assert stopped_at == 3                      # best at epoch 1, patience 2
    assert d.optimizer.learning_rate == pytest.approx(1e-3 * 0.25)   # halved at epochs 2 and 3


def test_metric_definitions_match_keras_semantics():
    y = np.eye(2, dtype=np.float32)[[0, 1, 1, 0, 1]]
    p = np.array([[.9, .1], [.2, .8], [.6, .4], [.3, .7], [.45, .55]], np.float32)
    assert M.CategoricalAccuracy().result(y, p) == pytest.approx(3 / 5)
    assert M.Precision(thresholds=0.5, class_id=1).result(y, p) == pytest.approx(2 / 3)
    assert M.Recall(thresholds=0.5, class_id=1).result(y, p) == pytest.approx(2 / 3)
    from sklearn.metrics import roc_auc_score
    rng = np.random.default_rng(0)
    yy = np.eye(2, dtype=np.float32)[rng.integers(0, 2, 400)]
    z = rng.normal(size=400) + yy[:, 1]
    pp = np.stack([1 / (1 + np.exp(z)), 1 / (1 + np.exp(-z))], -1).astype(np.float32)
    assert M.AUC().result(yy, pp) == pytest.approx(roc_auc_score(yy.reshape(-1), pp.reshape(-1)), abs=5e-3)


def test_class_weight_and_output_bias_formulas():
    from src.train import compute_class_weight, compute_output_bias
    labels = np.array([0] * 30 + [1] * 10)
    cw = compute_class_weight(labels)
    assert cw[0] == pytest.approx(0.5 * 40 / 30) and cw[1] == pytest.approx(0.5 * 40 / 10)
    np.testing.assert_allclose(compute_output_bias(labels), np.log([30 / 10, 10 / 30]))


def test_clip_aggregation_rules():
    """Host helpers kept from reference predict.py:283-308, 365-376 against the loop-for-loop restatement."""
    from oracle import clip_rules
    from src.predict import (highest_avg_contiguous_pred_prob, max_contiguous_b_line_preds,
                             predict_with_contiguity_threshold)
    assert max_contiguous_b_line_preds([0, 1, 1, 0, 1, 1, 1, 0]) == 3
    assert max_contiguous_b_line_preds([]) == 0
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 4, 5, 17, 100):
        pb = rng.random(n)
        probs = np.stack([1 - pb, pb], 1)
        np.testing.assert_array_equal(predict_with_contiguity_threshold(probs, 3, 0.5),
                                      clip_rules.contiguity_rule(probs, 3, 0.5))
        assert highest_avg_contiguous_pred_prob(pb, 4) == pytest.approx(clip_rules.sliding_window_rule(probs, 4)[1])
    assert highest_avg_contiguous_pred_prob([.1, .9, .8, .7, .6, .1], 4) == pytest.approx(0.75)
    assert highest_avg_contiguous_pred_prob([.2, .4], 4) == 0.0        # no full window: the reference's loop never runs


def test_engine_plan_and_unsupported_configs():
    from ab_line_classifier_b200.engine import build_plan
    with pytest.raises(NotImplementedError, match="IMG_DIM"):
        build_plan((96, 96, 3), 10, check_tiling=True)           # block3 maps would be 24x24: not tileable into 8x16 rectangles
    specs, feat = build_plan((128, 128, 3), 10)
    assert [s.name for s in specs] == ["block1_conv1", "block1_conv2", "block2_conv1", "block2_conv2",
                                       "block3_conv1", "block3_conv2", "block3_conv3"]
    assert [s.pool for s in specs] == [False, True, False, True, False, False, False]
    assert feat == (32, 32, 256)
    assert [(s.cin, s.cout, s.h) for s in specs][-1] == (256, 256, 32)
    with pytest.raises(NotImplementedError):
        build_plan((128, 128, 1), 10)


def test_save_and_pretrained_roundtrip(fake, tmp_path):
    m = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2)
    w = m.model.get_weights()
    path = model_mod.save_model(m.model, str(tmp_path / "m.npz"))
    m2 = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2, weights_path=path, seed=5)
    for k in w:
        np.testing.assert_array_equal(m2.model.get_weights()[k], w[k])


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ab_line_classifier_b200.engine import CutoffVGG16Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CutoffVGG16Engine()
    from ab_line_classifier_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.gap_fwd(torch.zeros(1, 2, 2, 64, dtype=torch.bfloat16), torch.zeros(1, 64))


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for d in ("ab_line_classifier_b200", "src"):
        for dp, _, fs in os.walk(os.path.join(root, d)):
            for f in fs:
                if f.endswith(".py"):
                    txt = open(os.path.join(dp, f)).read()
                    if "import oracle" in txt or "from oracle" in txt:
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_array_dataset_batches_shuffle_and_short_last_batch():
    from ab_line_classifier_b200.data import ArrayDataset
    x = np.arange(10 * 2 * 2 * 3, dtype=np.float32).reshape(10, 2, 2, 3)
    y = np.eye(2, dtype=np.float32)[np.arange(10) % 2]
    ds = ArrayDataset(x, y, 4, shuffle=True, seed=3, pin=False)
    for epoch in range(2):
        seen = []
        sizes = []
        for xb, yb in ds:
            sizes.append(len(xb))
            seen += [int(v) for v in (xb[:, 0, 0, 0] / 12).tolist()]
            np.testing.assert_array_equal(yb.numpy(), y[[int(v) for v in (xb[:, 0, 0, 0] / 12).tolist()]])
        assert sizes == [4, 4, 2] and sorted(seen) == list(range(10))          # every frame once, last batch short
    assert len(ds) == 3 and ds.n_frames == 10
    assert [len(b[0]) for b in ArrayDataset(x, None, 4)] == [4, 4, 2]


def test_array_dataset_data_parallel_shards_partition_every_global_batch():
    """the global batch is split contiguously, 1/world_size per rank; all ranks walk the same
    seeded order; a short last batch is cut to a multiple of world_size so local batches stay equal."""
    from ab_line_classifier_b200.data import ArrayDataset
    n, world, gb = 23, 4, 8
    x = np.arange(n * 3, dtype=np.float32).reshape(n, 1, 1, 3)
    y = np.eye(2, dtype=np.float32)[np.arange(n) % 2]
    for shuffle in (False, True):
        whole = [xb[:, 0, 0, 0].tolist() for xb, _ in ArrayDataset(x, y, gb, shuffle=shuffle, seed=9, pin=False)]
        parts = [[xb[:, 0, 0, 0].tolist() for xb, _ in ArrayDataset(x, y, gb, shuffle=shuffle, seed=9, pin=False,
                                                                  rank=r, world_size=world)] for r in range(world)]
        assert [len(p) for p in parts] == [3] * world                       # 8, 8 and 7 -> 4 frames: one dropped
        for b in range(3):
            local = [parts[r][b] for r in range(world)]
            assert len({len(l) for l in local}) == 1                        # equal local batches
            joined = sum(local, [])
            assert joined == whole[b][:len(joined)]                         # contiguous shares, in rank order
        assert len(sum(parts[0], [])) * world == 20
        assert len(ArrayDataset(x, y, gb, rank=1, world_size=world)) == 3
    # fewer leftover frames than ranks: the short batch is skipped on every rank
    assert [len(b[0]) for b in ArrayDataset(x[:18], None, gb, pin=False, rank=3, world_size=world)] == [2, 2]
    assert len(ArrayDataset(x[:18], None, gb, rank=3, world_size=world)) == 2
    with pytest.raises(ValueError):
        ArrayDataset(x, y, 6, world_size=4)
    with pytest.raises(ValueError):
        ArrayDataset(x, y, 8, rank=4, world_size=4)

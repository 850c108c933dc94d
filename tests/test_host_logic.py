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
        # the reference contract (predict.py:365-376): [N, C] frame probabilities in, np.array([1 - m, m]) out
        got = highest_avg_contiguous_pred_prob(probs, 4)
        assert isinstance(got, np.ndarray) and got.shape == (2,)
        np.testing.assert_allclose(got, clip_rules.sliding_window_rule(probs, 4), atol=1e-12)
        np.testing.assert_allclose(highest_avg_contiguous_pred_prob(pb, 4), got, atol=1e-12)   # 1-D extension
    two_d = np.array([[.9, .1], [.1, .9], [.2, .8], [.3, .7], [.4, .6], [.9, .1]])
    np.testing.assert_allclose(highest_avg_contiguous_pred_prob(two_d, 4), [0.25, 0.75])
    # no full window: the reference's loop never runs
    np.testing.assert_array_equal(highest_avg_contiguous_pred_prob(np.array([[.8, .2], [.6, .4]]), 4), [1.0, 0.0])


def test_engine_plan_and_unsupported_configs():
    from ab_line_classifier_b200.engine import build_plan
    with pytest.raises(NotImplementedError, match="IMG_DIM.*block2_pool"):
        build_plan((3, 3, 3), 10)                                # 1x1 block-2 maps: nothing left to pool
    assert build_plan((90, 90, 3), 10)[1] == (22, 22, 256)       # 45x45 block-2 maps: the pool drops the last row / column
    specs, feat = build_plan((96, 96, 3), 10, check_tiling=True)     # 24x24 block-3 maps: partial 8x16-pixel tiles
    assert feat == (24, 24, 256)
    from ab_line_classifier_b200.ops import igemm_gap_parts
    assert igemm_gap_parts(32, 32, 256) == 8 and igemm_gap_parts(24, 24, 256) == 6 and igemm_gap_parts(25, 25, 256) == 8
    assert igemm_gap_parts(32, 64, 128) == 16 and igemm_gap_parts(20, 24, 128) == 8      # 16-row tiles = 2 rectangles
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


def _keras_style_h5(path, weights, legacy_names=False, as_weights_file=False, model_config=None):
    """A file laid out like the ones Keras writes (save_model / save_weights / the keras-applications VGG16 notop file)."""
    from ab_line_classifier_b200.h5lite import H5Writer
    w = H5Writer()
    g = w.root if as_weights_file else w.group(w.root, "model_weights")
    layers = []
    for k in weights:
        if k.split("/")[0] not in layers:
            layers.append(k.split("/")[0])
    w.attr(g, "layer_names", np.array([n.encode() for n in layers + ["global_avgpool"]]))
    w.attr(g, "backend", b"tensorflow")
    if model_config is not None:
        w.attr(w.root, "model_config", model_config.encode())
    for ln in layers:
        lg = w.group(g, ln)
        names = []
        for part in ("kernel", "bias"):
            var = f"{ln}_{'W' if part == 'kernel' else 'b'}_1:0" if legacy_names else f"{ln}/{part}:0"
            names.append(var)
            parts = var.split("/")
            node = lg
            for p in parts[:-1]:
                node = w.group(node, p)
            w.dataset(node, parts[-1], weights[f"{ln}/{part}"])
        w.attr(lg, "weight_names", np.array([n.encode() for n in names]))
    w.group(g, "global_avgpool")        # a layer without weights, as Keras writes it
    w.save(path)


def test_h5lite_reads_what_it_writes_incl_many_links_and_attributes(tmp_path):
    from ab_line_classifier_b200.h5lite import H5File, H5Writer
    rng = np.random.default_rng(0)
    w = H5Writer()
    w.attr(w.root, "keras_version", b"2.9.0")
    w.attr(w.root, "an_int", np.int64(7))
    w.attr(w.root, "floats", np.arange(3, dtype=np.float64))
    ref = {}
    for i in range(37):                               # more links than one symbol-table node holds (8): several SNODs
        a = rng.normal(size=(i % 4 + 1, 3)).astype(np.float32 if i % 2 else np.float64)
        ref[f"g/d{i:02d}"] = a
        w.create_path(f"g/d{i:02d}", a)
    w.create_path("ints", np.arange(10, dtype=np.int32))
    w.create_path("empty", np.zeros((0, 4), np.float32))
    p = str(tmp_path / "t.h5")
    w.save(p)
    f = H5File(p)
    assert f.keys() == ["empty", "g", "ints"] and f.attrs["keras_version"] == b"2.9.0" and int(f.attrs["an_int"]) == 7
    np.testing.assert_array_equal(f.attrs["floats"], [0.0, 1.0, 2.0])
    assert len(f["g"].keys()) == 37 and not f["g/d00"].is_group and "g/nope" not in f
    for k, a in ref.items():
        got = f[k].read()
        assert got.dtype == a.dtype and np.array_equal(got, a)
    np.testing.assert_array_equal(f["ints"].read(), np.arange(10))
    assert f["empty"].read().shape == (0, 4)
    assert [k for k, _ in f.visit_datasets()][:2] == ["empty", "g/d00"]
    with pytest.raises(ValueError):
        H5File(data=b"not an hdf5 file at all" * 40)


def test_h5lite_reads_structures_the_writer_never_emits():
    """Hand-assembled HDF5 structures the HDF5 C library produces but H5Writer does not: an object-header continuation
    block, a version-3 attribute, a variable-length string through the global heap, a compact dataset and a chunked
    dataset behind a version-1 chunk B-tree (all per the HDF5 file format specification)."""
    import struct
    from ab_line_classifier_b200 import h5lite as H
    buf = bytearray(b"\0" * 96)

    def alloc(b):
        b = H._pad8(bytes(b))
        a = len(buf)
        buf.extend(b)
        return a

    # global heap collection holding the string b"hello world" as object 1
    s = b"hello world"
    gcol_body = struct.pack("<HHIQ", 1, 1, 0, len(s)) + H._pad8(s) + struct.pack("<HHIQ", 0, 0, 0, 0)
    gcol = alloc(b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(gcol_body)) + gcol_body)
    # vlen-string attribute, message version 3 (no padding, charset byte)
    vl_dt = struct.pack("<BBBBI", 0x19, 0x01, 0, 0, 16) + struct.pack("<BBBBI", 0x13, 0, 0, 0, 1)
    ds_scalar = struct.pack("<BBBB", 2, 0, 0, 0)                       # dataspace v2, scalar
    attr3 = struct.pack("<BBHHHB", 3, 0, 5, len(vl_dt), len(ds_scalar), 0) + b"note\0" + vl_dt + ds_scalar + \
        struct.pack("<IQI", len(s), gcol, 1)
    # chunked float32 dataset [5, 4] in chunks of [2, 4]: three chunks, the last one partial
    data = np.arange(20, dtype=np.float32).reshape(5, 4)
    chunks = []
    for r0 in (0, 2, 4):
        c = np.zeros((2, 4), np.float32)
        c[: min(2, 5 - r0)] = data[r0:r0 + 2]
        chunks.append((r0, alloc(c.tobytes())))
    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, 3, H.UNDEF, H.UNDEF)
    for r0, addr in chunks:
        tree += struct.pack("<II", 32, 0) + struct.pack("<QQQ", r0, 0, 0) + struct.pack("<Q", addr)
    tree += struct.pack("<II", 0, 0) + struct.pack("<QQQ", 6, 0, 0)
    btree = alloc(tree)
    layout = struct.pack("<BBB", 3, 2, 3) + struct.pack("<Q", btree) + struct.pack("<III", 2, 4, 4)
    cont = alloc(H._msg(0x08, layout) + H._msg(0x0C, attr3))             # continuation block: layout + attribute
    msgs = [H._msg(0x01, H._ds_bytes((5, 4))), H._msg(0x03, H._dt_bytes(np.float32)),
            H._msg(0x10, struct.pack("<QQ", cont, len(buf) - cont))]
    body = b"".join(msgs)
    chunked = alloc(struct.pack("<BBHII4x", 1, 0, 5, 1, len(body)) + body)
    # compact int16 dataset
    small = np.array([3, -4, 5], np.int16)
    lay_c = struct.pack("<BBH", 3, 0, small.nbytes) + small.tobytes()
    body = b"".join([H._msg(0x01, H._ds_bytes((3,))), H._msg(0x03, H._dt_bytes(np.int16)), H._msg(0x08, lay_c)])
    compact = alloc(struct.pack("<BBHII4x", 1, 0, 3, 1, len(body)) + body)
    # root group: symbol table with the two datasets
    heap = bytearray(b"\0" * 8)
    offs = {}
    for n in ("chunked", "compact"):
        offs[n] = len(heap)
        heap += H._pad8(n.encode() + b"\0")
    heap_data = alloc(heap)
    heap_addr = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), H.UNDEF, heap_data))
    ent = struct.pack("<QQII16x", offs["chunked"], chunked, 0, 0) + struct.pack("<QQII16x", offs["compact"], compact, 0, 0)
    snod = alloc(b"SNOD" + struct.pack("<BBH", 1, 0, 2) + ent + b"\0" * (40 * 6))
    gtree = alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, H.UNDEF, H.UNDEF) + struct.pack("<QQQ", 0, snod, offs["compact"]))
    body = H._msg(0x11, struct.pack("<QQ", gtree, heap_addr))
    root = alloc(struct.pack("<BBHII4x", 1, 0, 1, 1, len(body)) + body)
    sb = H.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + \
        struct.pack("<QQQQ", 0, H.UNDEF, len(buf), H.UNDEF) + struct.pack("<QQII", 0, root, 0, 0) + b"\0" * 16
    buf[:len(sb)] = sb
    f = H.H5File(data=bytes(buf))
    assert f.keys() == ["chunked", "compact"]
    np.testing.assert_array_equal(f["chunked"].read(), data)
    assert f["chunked"].attrs["note"] == "hello world"
    np.testing.assert_array_equal(f["compact"].read(), small)


def test_keras_h5_checkpoints_load_by_layer_name(fake, tmp_path):
    """f3: a model saved the way the reference saves it (Keras HDF5: model_weights/<layer>/<layer>/kernel:0, names with
    ':0', layer_names / weight_names attributes) loads into the B200 path; so does a weights-only file with the legacy
    variable names of the keras-applications VGG16 file; and save_model('...h5') writes that layout back."""
    import json
    src = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2, seed=3)
    w = src.model.get_weights()
    cfg = json.dumps(model_mod._keras_model_config(src.model))
    full = str(tmp_path / "model.h5")
    _keras_style_h5(full, w, model_config=cfg)
    got, parsed, opt = model_mod.read_keras_h5(full)
    assert sorted(got) == sorted(w) and opt == {} and parsed["class_name"] == "Functional"
    for k in w:
        np.testing.assert_array_equal(got[k], w[k])
    m = model_mod.load_model(full)                       # architecture from model_config: 8x8 input, cutoff 4
    assert m.engine.input_shape == (8, 8, 3) and m.engine.cutoff_layer == 4 and m.n_classes == 2
    for k in w:
        np.testing.assert_array_equal(m.get_weights()[k], w[k])
    # weights-only file, legacy names (block1_conv1_W_1:0 / _b_1:0), layer groups at the root: by-name transplant
    notop = str(tmp_path / "vgg16_notop.h5")
    conv_only = {k: v + 1.0 for k, v in w.items() if not k.startswith("output")}
    _keras_style_h5(notop, conv_only, legacy_names=True, as_weights_file=True)
    m2 = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2, seed=9, vgg16_weights=notop)          # VGG16(weights='imagenet')
    for k in conv_only:
        np.testing.assert_array_equal(m2.model.get_weights()[k], conv_only[k])
    m3 = model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2, seed=9, weights_path=full)            # model_utils.py:6-23
    for k in w:
        np.testing.assert_array_equal(m3.model.get_weights()[k], w[k])
    with pytest.warns(UserWarning, match="imagenet"):
        model_mod.CutoffVGG16(HP, [8, 8, 3], [], 2)
    # and back: save_model(..., '.h5') writes the Keras layout
    out = model_mod.save_model(src.model, str(tmp_path / "saved.h5"))
    from ab_line_classifier_b200.h5lite import H5File
    f = H5File(out)
    assert [n.decode() for n in f["model_weights"].attrs["layer_names"]][:2] == ["input", "block1_conv1"]
    assert [n.decode() for n in f["model_weights/block1_conv1"].attrs["weight_names"]] == \
        ["block1_conv1/kernel:0", "block1_conv1/bias:0"]
    np.testing.assert_array_equal(f["model_weights/block1_conv1/block1_conv1/kernel:0"].read(), w["block1_conv1/kernel"])
    m4 = model_mod.load_model(out)
    for k in w:
        np.testing.assert_array_equal(m4.get_weights()[k], w[k])


def test_keras_fixture_file_converts_to_npz(tmp_path):
    """tests/golden/keras_layout_cutoffvgg16_8x8.h5 (tests/golden/make_keras_layout_fixture.py) through the offline
    converter tools/keras_h5_to_npz.py."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fix = os.path.join(root, "tests", "golden", "keras_layout_cutoffvgg16_8x8.h5")
    dst = str(tmp_path / "conv.npz")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "keras_h5_to_npz.py"), fix, dst],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    z = np.load(dst)
    assert int(z["config/cutoff_layer"]) == 4 and tuple(z["config/input_shape"]) == (8, 8, 3)
    assert z["weights/block1_conv2/kernel"].shape == (3, 3, 64, 64) and z["weights/output/kernel"].shape == (64, 2)
    exp = np.load(os.path.join(root, "tests", "golden", "keras_layout_cutoffvgg16_8x8_expected.npz"))
    for k in exp.files:
        np.testing.assert_array_equal(z["weights/" + k], exp[k])


def test_b200_config_section_is_optional():
    """src/train.py must run on the reference's own config.yml, which has no B200 section."""
    from src.train import b200_options
    assert b200_options({"TRAIN": {}, "B200": {"SYNTH_TRAIN_FRAMES": 64, "SYNTH_VAL_FRAMES": 8, "AUGMENT": False}}) == (64, 8, False)
    assert b200_options({"TRAIN": {}}) == (2048, 256, True)
    assert b200_options({"TRAIN": {}, "B200": None}) == (2048, 256, True)


def test_launch_accounting_of_the_wgrad_entry_point(monkeypatch):
    """bench.py's gpu_launches comes from _native.launch_count: abl_conv3x3_wgrad is one kernel when the GEMM reduces its
    own pixel slices (Cin >= 128, the default) and two when wgrad_reduce_kernel follows it (Cin = 64, or the switch)."""
    from ab_line_classifier_b200 import _native
    args128 = [0] * 11 + [128, 256, 0]        # (x, dz, dw, db, db_parts, ws, ws_bytes, accumulate, B, H, W, Cin, Cout, stream)
    args64 = [0] * 11 + [64, 128, 0]
    monkeypatch.delenv("ABL_WGRAD_FUSED_REDUCE", raising=False)
    assert _native._wgrad_launches(args128) == 1 and _native._wgrad_launches(args64) == 2
    monkeypatch.setenv("ABL_WGRAD_FUSED_REDUCE", "0")
    assert _native._wgrad_launches(args128) == 2 and _native._wgrad_launches(args64) == 2
    monkeypatch.setenv("ABL_WGRAD_FUSED_REDUCE", "2")
    assert _native._wgrad_launches(args128) == 1 and _native._wgrad_launches(args64) == 1

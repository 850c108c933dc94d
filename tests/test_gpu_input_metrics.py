"""-m gpu: the augmentation kernel and the device-side metric counters against their oracles, through the C ABI
(SURVEY section 8 row f4).  Both are exact: the augmentation rounds every operation like the NumPy restatement and the
counters are integers."""
import numpy as np
import pytest
import torch

from oracle import augment as A
from oracle import metric_counts, synth

pytestmark = pytest.mark.gpu


def _metric_list():
    from ab_line_classifier_b200.metrics import AUC, Precision, Recall
    names = ["a_lines", "b_lines"]
    return (["accuracy", AUC(name="auc")] + [Precision(name="precision_" + n, thresholds=0.5, class_id=c) for c, n in enumerate(names)]
            + [Recall(name="recall_" + n, thresholds=0.5, class_id=c) for c, n in enumerate(names)])


def _raw(B, seed, cfg=None):
    from ab_line_classifier_b200 import augment as P
    return P.draw_raw(np.random.default_rng(seed), B, P.AugmentConfig(cfg))


@pytest.mark.parametrize("shape,u8", [((128, 128), True), ((128, 128), False), ((24, 40), True), ((176, 200), True),
                                      ((176, 200), False)])
@pytest.mark.parametrize("preprocess", [True, False])
def test_augment_kernel_is_bit_exact_against_the_oracle(shape, u8, preprocess):
    """128x128 and 24x40 ping-pong through shared memory, 176x200 (two planes = 275 KB) through the workspace."""
    from ab_line_classifier_b200 import augment as P
    from ab_line_classifier_b200 import ops
    H, W = shape
    B = 5
    rng = np.random.default_rng(H + W)
    frames = rng.integers(0, 256, (B, H, W, 3)).astype(np.uint8)
    if not u8:
        frames = (frames.astype(np.float32) + rng.random((B, H, W, 3)).astype(np.float32)).clip(0, 255)
    raw = _raw(B, seed=7)
    raw["brightness"] = rng.uniform(-40, 40, B).astype(np.float32)    # any delta: added as is, nothing is clipped
    raw["zoom"][0], raw["dx"][0], raw["dy"][0], raw["angle"][0], raw["flip"][0], raw["brightness"][0] = 1, 0, 0, 0, 0, 0
    assert (ops.augment_workspace_bytes(B, H, W) > 0) == (H * W > 29056)
    aug = P.DeviceAugmenter(preprocess=preprocess)
    out = aug(frames, params=P.params_from_raw(raw, H, W)).cpu().numpy()
    want = A.augment_batch(frames, raw, preprocess)
    assert out.shape == want.shape and np.array_equal(out, want), float(np.abs(out - want).max())
    ident = frames[0].astype(np.float32)
    assert np.array_equal(out[0], ident[..., ::-1] - A.BGR_MEAN if preprocess else ident)   # identity draw: the scaler only


def test_augmenter_draws_rotates_buffers_and_handles_empty_batches():
    from ab_line_classifier_b200 import augment as P
    frames, _ = synth.synth_frames(8, seed=3)
    aug = P.DeviceAugmenter({"ZOOM_RANGE": 0.1, "ROTATION_RANGE": 45}, seed=4)
    outs = [aug(frames) for _ in range(3)]
    torch.cuda.synchronize()
    assert len({o.data_ptr() for o in outs}) == 3                     # earlier batches stay valid
    assert not torch.equal(outs[0], outs[1])                          # fresh draws per batch
    assert aug.last_params.shape == (8, P.N_PARAMS)
    assert aug(frames[:0]).shape == (0, 128, 128, 3)
    again = P.DeviceAugmenter({"ZOOM_RANGE": 0.1, "ROTATION_RANGE": 45}, seed=4)(torch.from_numpy(frames).cuda())
    assert torch.equal(again, outs[0])                                # seeded; device-resident frames give the same result
    with pytest.raises(ValueError):
        aug(np.zeros((2, 8, 8), np.uint8))


@pytest.mark.parametrize("B,C,T", [(1, 2, 200), (257, 2, 200), (70001, 2, 200), (1000, 3, 50), (0, 2, 200)])
def test_metric_counters_equal_the_oracle(B, C, T):
    from ab_line_classifier_b200 import metrics as M
    from ab_line_classifier_b200 import ops
    rng = np.random.default_rng(B + C)
    thr = M.auc_thresholds(T)
    cthr = np.full(C, 1.0 / C, np.float32)
    z = rng.normal(size=(B, C)).astype(np.float32) * 3
    p = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    if B:
        p[::5, 0] = thr[rng.integers(1, T - 1, len(p[::5]))]         # exactly on thresholds: '>' is strict in fp32
        p[::7] = 1.0 / C                                             # ties: the first maximum wins
    y = np.eye(C, dtype=np.float32)[rng.integers(0, C, B)]
    n_cnt = ops.metrics_counter_count(C, T)
    assert n_cnt == 2 + 3 * C + 2 * (T + 1)
    counters = torch.zeros(n_cnt, device="cuda", dtype=torch.int64)
    pd, yd = torch.from_numpy(p).cuda().reshape(B, C), torch.from_numpy(y).cuda().reshape(B, C)
    td, cd = torch.from_numpy(thr).cuda(), torch.from_numpy(cthr).cuda()
    for rep in range(2):                                             # accumulates across calls
        for i in range(0, max(B, 1), 40000):
            ops.metrics_accumulate(pd[i:i + 40000].contiguous(), yd[i:i + 40000].contiguous(), td, cd, counters)
    want = metric_counts.counts_from_arrays(y, p, thr, cthr)
    assert np.array_equal(counters.cpu().numpy(), 2 * want)


def test_evaluate_and_fit_metrics_come_from_the_device_counters():
    from ab_line_classifier_b200 import metrics as M
    from ab_line_classifier_b200.data import ArrayDataset
    from src.models.models import get_model
    model_def, prep = get_model("cutoffvgg16")
    fr, lab = synth.synth_frames(200, seed=11)
    y = synth.one_hot(lab)
    hp = {"LR_EXTRACT": 3e-4, "LR_FINETUNE": 1e-5, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7,
          "EXTRACT_EPOCHS": 1}
    model = model_def(hp, [128, 128, 3], _metric_list(), 2)
    val = ArrayDataset(prep(fr), y, 64)
    model.model.compile(optimizer=model.optimizer_extract, metrics=model.metrics)
    res = model.evaluate(val, verbose=0)
    objs = model.model.metric_objs
    assert M.device_plan(objs, 2) is not None
    assert int(model.model.engine.metrics_read()[0]) == 200

    class Collect(M.Metric):          # not expressible by the counters: evaluate() then gathers arrays and uses result()
        name = "collect"

        def result(self, y_true, y_pred):
            self.seen = (np.array(y_true), np.array(y_pred))
            return 0.0

    col = Collect()
    objs.append(col)
    assert M.device_plan(objs, 2) is None
    res_arrays = model.model.evaluate(val, verbose=0)
    objs.pop()
    assert res[0] == pytest.approx(res_arrays[0], abs=1e-5)           # loss: float atomics
    assert res[1:] == res_arrays[1:-1]                                # same forward; counters vs whole arrays: identical
    assert np.array_equal(col.seen[0], y) and col.seen[1].shape == (200, 2)
    # training: accuracy / auc / precision / recall of every epoch are reported from the counters
    train = ArrayDataset(fr, y, 64, shuffle=True, seed=2)              # raw uint8 frames, pinned rotating batch buffers
    model.fit(train, epochs=2, validation_data=val, verbose=0)
    h = model.history_finetune.history
    for k in ("accuracy", "auc", "precision_a_lines", "recall_b_lines", "val_accuracy", "val_auc"):
        assert k in h and all(0.0 <= v <= 1.0 for v in h[k]), (k, h.get(k))
    assert int(model.model.engine.metrics_read()[0]) == 200           # last use: the validation pass


def test_fit_on_augmented_batches_from_the_dataset():
    """``prepare(train_set, shuffle=True, augment=True)`` (train_utils.py:85): RGB frames in, augmented + preprocessed
    device batches out; host-resident (pinned, fenced buffers) and device-resident datasets; frames at another size are
    resized first."""
    from ab_line_classifier_b200.data import ArrayDataset
    from src.models.models import get_model
    cfg = {"ZOOM_RANGE": 0.1, "HORIZONTAL_FLIP": True, "WIDTH_SHIFT_RANGE": 0.2, "HEIGHT_SHIFT_RANGE": 0.2,
           "SHEAR_RANGE": 10, "ROTATION_RANGE": 45, "BRIGHTNESS_RANGE": 0.3}
    fr, lab = synth.synth_frames(160, seed=12)
    y = synth.one_hot(lab)
    for kw in ({}, {"device": "cuda"}):
        ds = ArrayDataset(fr, y, 32, shuffle=True, seed=5, augment=cfg, **kw)
        batches = list(ds)
        assert len(batches) == 5
        x0, y0 = batches[0]
        assert x0.is_cuda and x0.dtype == torch.float32 and tuple(x0.shape) == (32, 128, 128, 3)
        lo, hi = float(x0.min()), float(x0.max())
        assert lo >= -123.68 - 0.5 and hi <= 255 - 103.939 + 0.5       # interpolated 0..255 + the small delta, mean-subtracted
        fill_b = float(np.float32(ds._aug.cfg.brightness_delta) - np.float32(103.939))
        assert float((x0 == fill_b).float().mean()) > 0.001            # constant fill (0 + delta) where the warp left the frame
    big = np.repeat(np.repeat(fr[:40], 2, axis=1), 2, axis=2)           # 256x256 raw frames
    ds = ArrayDataset(big, y[:40], 20, shuffle=True, seed=6, augment=cfg, img_dim=(128, 128))
    xb, _ = next(iter(ds))
    assert tuple(xb.shape) == (20, 128, 128, 3)
    # the reference-facing wrapper: Preprocessor(scale_fn).prepare(ds, shuffle=True, augment=True) (train_utils.py:84-85)
    from src.data.preprocessor import Preprocessor
    prepared = Preprocessor(get_model("cutoffvgg16")[1]).prepare((big[:8], lab[:8]), shuffle=True, augment=True)
    xb, yb = next(iter(prepared))
    assert xb.is_cuda and xb.dtype == torch.float32 and tuple(xb.shape) == (8, 128, 128, 3) and tuple(yb.shape) == (8, 2)
    model_def, _ = get_model("cutoffvgg16")
    hp = {"LR_EXTRACT": 3e-4, "LR_FINETUNE": 1e-5, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7,
          "EXTRACT_EPOCHS": 2}
    model = model_def(hp, [128, 128, 3], _metric_list(), 2)
    train = ArrayDataset(fr, y, 32, shuffle=True, seed=5, augment=cfg)
    model.fit(train, epochs=3, verbose=0)
    losses = model.history_extract.history["loss"] + model.history_finetune.history["loss"]
    assert np.isfinite(losses).all() and len(losses) == 4


def test_kernels_reproduce_the_committed_fixture():
    """The same inputs as tests/golden/make_golden_augment.py through the CUDA kernels == the committed vectors."""
    import importlib.util
    import os
    from ab_line_classifier_b200 import augment as P
    from ab_line_classifier_b200 import ops
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_augment", os.path.join(here, "make_golden_augment.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(here, "augment_metrics_seed77.npz"))
    frames, probs, y, thr = gen.inputs()
    out = P.DeviceAugmenter()(frames, params=P.params_from_raw(gen.RAW, 128, 128)).cpu().numpy()
    assert np.array_equal(out[:, ::4, ::4], gold["augmented_sub"])
    assert np.allclose(out.astype(np.float64).sum(axis=(1, 2)), gold["augmented_sum"], rtol=0, atol=1e-6)
    counters = torch.zeros(ops.metrics_counter_count(2, 200), device="cuda", dtype=torch.int64)
    ops.metrics_accumulate(torch.from_numpy(probs).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(thr).cuda(),
                           torch.full((2,), 0.5, device="cuda"), counters)
    assert np.array_equal(counters.cpu().numpy(), gold["counts"])

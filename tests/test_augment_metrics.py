"""CPU: the augmentation oracle against closed-form cases, the host-side parameter builder against the oracle's
restatement of Keras' matrices, and the count-based metric evaluation against the array-based definitions
(src/data/preprocessor.py:25-32, 56-63 and src/train.py:109-115)."""
import numpy as np
import pytest

from ab_line_classifier_b200 import augment as P
from ab_line_classifier_b200 import metrics as M
from oracle import augment as A
from oracle import metric_counts

IDENT = [1, 0, 0, 0, 1, 0, 0, 0]


def _img(h=16, w=20, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3)).astype(np.float32)


def test_projective_bilinear_identity_and_integer_shift():
    img = _img()
    assert np.array_equal(A.projective_bilinear(img, IDENT), img)
    # RandomTranslation by +25 % of the width: content moves 5 px to the right, the vacated columns take the fill value
    out = A.projective_bilinear(img, A.translation_transform(0.25, 0.0, 16, 20))
    assert np.array_equal(out[:, 5:], img[:, :-5]) and (out[:, :5] == 0).all()
    out = A.projective_bilinear(img, A.translation_transform(0.0, -0.125, 16, 20))        # 2 px up
    assert np.array_equal(out[:-2], img[2:]) and (out[-2:] == 0).all()


def test_projective_bilinear_half_pixel_blend_and_border_fill():
    img = _img()
    out = A.projective_bilinear(img, [1, 0, 0.5, 0, 1, 0, 0, 0])         # sample half a pixel to the right
    assert np.allclose(out[:, :-1], 0.5 * (img[:, :-1] + img[:, 1:]), atol=1e-4)
    assert np.allclose(out[:, -1], 0.5 * img[:, -1], atol=1e-4)          # the outside neighbour reads as fill value 0


def test_rotation_and_zoom_matrices():
    img = _img(16, 16)
    out = A.projective_bilinear(img, A.rotation_transform(np.pi / 2, 16, 16))
    assert np.abs(out - np.rot90(img, 1)).max() < 1e-2                   # quarter turn about the centre
    out = A.projective_bilinear(img, A.rotation_transform(2 * np.pi * 3, 16, 16))
    assert np.abs(out - img).max() < 1e-1                                # whole turns (RandomRotation(45) draws many)
    # zoom keeps the centre fixed: pixel ((W-1)/2, (H-1)/2) maps to itself
    t = A.zoom_transform(0.8, 17, 21)
    assert np.isclose(t[0] * 10 + t[2], 10) and np.isclose(t[4] * 8 + t[5], 8)
    out = A.projective_bilinear(_img(17, 21), t)
    assert np.allclose(out[8, 10], _img(17, 21)[8, 10], atol=1e-3)


def test_augment_frame_order_flip_brightness_preprocess():
    img = _img()
    # no geometry, flip on, brightness delta 0.2: mirrored columns, + 0.2 as adjust_brightness adds it (no 0..255
    # scaling, no clipping: the reference's own RandomBrightness layer, preprocessor.py:83-104), then BGR minus means
    out = A.augment_frame(img, 1.0, 0.0, 0.0, 0.0, 1, 0.2, preprocess=False)
    assert np.array_equal(out, img[:, ::-1] + np.float32(0.2))
    img[0, 0, 0] = 0                                                     # a black red value: goes below -mean, unclipped
    out = A.augment_frame(img, 1.0, 0.0, 0.0, 0.0, 0, -1.0, preprocess=True)
    assert np.array_equal(out, (img - np.float32(1.0))[..., ::-1] - A.BGR_MEAN) and out.min() < -A.BGR_MEAN.max()
    out = A.augment_frame(img, 1.0, 0.0, 0.0, 0.0, 0, 0.0, preprocess=True)
    assert np.allclose(out, img[..., ::-1] - A.BGR_MEAN, atol=1e-5)
    # the stages crop: a frame pushed out by the translation does not come back with the rotation
    out = A.augment_frame(img, 1.0, 1.0, 0.0, 0.3, 0, 0.0, preprocess=False)
    assert (out == 0).all()
    assert A.augment_batch(np.zeros((0, 16, 20, 3), np.float32), {k: [] for k in ("zoom", "dx", "dy", "angle", "flip", "brightness")}).shape == (0, 16, 20, 3)


def test_philox_known_answers_and_the_reference_brightness_constant():
    """Philox4x32-10 against the published Random123 known-answer vectors (both restatements), and the constant delta
    of ``tf.image.stateless_random_brightness(x, 0.3, (123, 0))`` (preprocessor.py:103) from product and oracle."""
    kats = [([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
            ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
            ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
             [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1])]
    for ctr, key, want in kats:
        assert P._philox4x32_10(ctr, key) == want
        assert A.philox_block(ctr, key) == want
    for factor in (0.3, 0.05, 1.0):
        d = P.reference_brightness_delta(factor)
        assert d.dtype == np.float32 and d == A.stateless_brightness_delta(factor) and -factor <= d < factor
    assert P.reference_brightness_delta(0.0) == 0.0
    cfg = P.AugmentConfig({"BRIGHTNESS_RANGE": 0.3})
    raw = P.draw_raw(np.random.default_rng(0), 7, cfg)
    assert (raw["brightness"] == cfg.brightness_delta).all()          # one constant for every frame of every batch
    assert np.array_equal(P.draw_raw(np.random.default_rng(1), 7, cfg)["brightness"], raw["brightness"])


def test_host_parameter_builder_matches_keras_matrices_and_ranges():
    cfg = P.AugmentConfig({"ZOOM_RANGE": 0.1, "HEIGHT_SHIFT_RANGE": 0.2, "WIDTH_SHIFT_RANGE": 0.2, "ROTATION_RANGE": 45,
                           "BRIGHTNESS_RANGE": 0.3, "HORIZONTAL_FLIP": True, "SHEAR_RANGE": 10})
    raw = P.draw_raw(np.random.default_rng(3), 4000, cfg)
    assert 0.9 <= raw["zoom"].min() < 0.91 and 1.09 < raw["zoom"].max() <= 1.1
    assert np.abs(raw["dx"]).max() <= 0.2 and np.abs(raw["dy"]).max() <= 0.2
    assert 44 * 2 * np.pi < np.abs(raw["angle"]).max() <= 45 * 2 * np.pi   # the factor is a fraction of 2 pi in Keras
    assert np.abs(raw["brightness"]).max() <= 0.3 and 0.45 < raw["flip"].mean() < 0.55
    assert set(np.unique(raw["flip"])) == {0.0, 1.0}
    H, W = 128, 96
    p = P.params_from_raw({k: v[:64] for k, v in raw.items()}, H, W)
    assert p.shape == (64, P.N_PARAMS) and p.dtype == np.float32
    for i in range(64):
        want = np.concatenate([A.zoom_transform(raw["zoom"][i], H, W)[:6],
                               A.translation_transform(raw["dx"][i], raw["dy"][i], H, W)[:6],
                               A.rotation_transform(raw["angle"][i], H, W)[:6],
                               [raw["flip"][i], raw["brightness"][i]]]).astype(np.float32)
        assert np.array_equal(p[i], want)
    with pytest.raises(ValueError):
        P.AugmentConfig({"BRIGHTNESS_RANGE": -0.5})
    with pytest.raises(ValueError):
        P.AugmentConfig({"ZOOM_RANGE": -0.1})


def _reference_metric_list():
    thr = 1.0 / 2
    return ([M.CategoricalAccuracy(), M.AUC(name="auc")]
            + [M.Precision(name=f"precision_{c}", thresholds=thr, class_id=c) for c in range(2)]
            + [M.Recall(name=f"recall_{c}", thresholds=thr, class_id=c) for c in range(2)])


def test_count_based_metrics_equal_array_based_definitions():
    rng = np.random.default_rng(5)
    objs = _reference_metric_list()
    plan = M.device_plan(objs, 2)
    assert plan is not None and plan[0].dtype == np.float32 and len(plan[0]) == 200 and list(plan[1]) == [0.5, 0.5]
    for n in (1, 7, 1000):
        z = rng.normal(size=(n, 2)).astype(np.float32) * 3
        p = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
        # predictions sitting exactly on thresholds (strict '>' in fp32 must be honoured) and exact ties
        p[::5, 0] = plan[0][rng.integers(1, 199, len(p[::5]))]
        p[::5, 1] = 1 - p[::5, 0]
        p[::11] = 0.5
        y = np.eye(2, dtype=np.float32)[rng.integers(0, 2, n)]
        total = None
        for i in range(0, n, 300):                       # accumulated batch by batch, as the device does
            c = metric_counts.counts_from_arrays(y[i:i + 300], p[i:i + 300], *plan)
            total = c if total is None else total + c
        counts = M.MetricCounts(total, 2, *plan)
        for m in objs:
            assert m.from_counts(counts) == pytest.approx(m.result(y, p), abs=1e-12), m.name


def test_metric_counts_hand_case_and_plan_rejections():
    thr = np.array([-1e-7, 0.5, 1 + 1e-7], np.float32)
    y = np.array([[1, 0], [0, 1], [0, 1]], np.float32)
    p = np.array([[0.8, 0.2], [0.5, 0.5], [0.1, 0.9]], np.float32)
    c = metric_counts.counts_from_arrays(y, p, thr, [0.5, 0.5])
    # frames 3; argmax hits: frame 0 (0==0), frame 1 (tie -> class 0 vs label 1: miss), frame 2 (hit)
    assert list(c[:2]) == [3, 2]
    assert list(c[2:5]) == [1, 0, 0]      # class 0: tp (0.8), fp none (0.5 is not > 0.5), fn none
    assert list(c[5:8]) == [1, 0, 1]      # class 1: tp 0.9; 0.5 is a miss -> fn
    # positives 0.8, 0.5, 0.9 -> buckets 2, 1, 2; negatives 0.2, 0.5, 0.1 -> all bucket 1
    assert list(c[8:12]) == [0, 1, 2, 0] and list(c[12:16]) == [0, 3, 0, 0]
    mc = M.MetricCounts(c, 2, thr, [0.5, 0.5])
    assert list(mc.auc_tp) == [3, 2, 0] and list(mc.auc_fp) == [3, 0, 0] and (mc.n_pos, mc.n_neg) == (3, 3)
    assert M.device_plan([M.Precision(thresholds=0.5)], 2) is None                     # no class_id
    assert M.device_plan([M.AUC(num_thresholds=50), M.AUC(name="b", num_thresholds=200)], 2) is None
    assert M.device_plan([M.Precision(thresholds=0.3, class_id=1), M.Recall(thresholds=0.6, class_id=1)], 2) is None

    class Custom(M.Metric):
        name = "custom"
    assert M.device_plan([Custom()], 2) is None
    assert len(M.device_plan([M.AUC(num_thresholds=50)], 3)[0]) == 50


def test_evaluate_falls_back_to_arrays_for_custom_metrics():
    from tests.fake_engine import FakeEngine
    from ab_line_classifier_b200.model import B200Model
    from oracle import synth

    class MeanB(M.Metric):
        name = "mean_b"

        def result(self, y_true, y_pred):
            return float(np.mean(y_pred[:, 1]))

    eng = FakeEngine()
    model = B200Model(eng, (8, 8, 3), 2, 0.45)
    x = np.random.default_rng(0).normal(size=(6, 8, 8, 3)).astype(np.float32)
    y = synth.one_hot(np.arange(6) % 2)
    from ab_line_classifier_b200.model import Adam
    model.compile(optimizer=Adam(1e-3), metrics=["accuracy", MeanB()])
    res = model.evaluate([(x[:4], y[:4]), (x[4:], y[4:])], verbose=0)
    p = model.predict([(x, None)], verbose=0)
    assert res[2] == pytest.approx(float(p[:, 1].mean()), abs=1e-6)
    model.compile(optimizer=Adam(1e-3), metrics=_reference_metric_list())
    res2 = model.evaluate([(x[:4], y[:4]), (x[4:], y[4:])], verbose=0)
    assert res2[0] == pytest.approx(res[0]) and res2[1] == pytest.approx(res[1])
    assert res2[2] == pytest.approx(M.AUC().result(y, p), abs=1e-9)


def test_oracles_reproduce_the_committed_fixture():
    """tests/golden/augment_metrics_seed77.npz (made by tests/golden/make_golden_augment.py): drift guard."""
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_augment", os.path.join(here, "make_golden_augment.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(here, "augment_metrics_seed77.npz"))
    frames, probs, y, thr = gen.inputs()
    out = A.augment_batch(frames, gen.RAW, preprocess=True)
    assert np.array_equal(out[:, ::4, ::4], gold["augmented_sub"])
    assert np.allclose(out.astype(np.float64).sum(axis=(1, 2)), gold["augmented_sum"], rtol=0, atol=1e-6)
    assert np.array_equal(probs, gold["probs"]) and np.array_equal(y, gold["y"])
    assert np.array_equal(metric_counts.counts_from_arrays(y, probs, thr, [0.5, 0.5]), gold["counts"])
    assert np.array_equal(out[0], frames[0].astype(np.float32)[..., ::-1] - A.BGR_MEAN)     # frame 0: identity draws


def test_preprocessor_prepare_mirrors_the_reference_pipeline():
    """src/data/preprocessor.py: prepare(ds, shuffle, augment) -> batches of (frames, one-hot), short last batch;
    resize = TF2 nearest neighbour; scaler paths (reference preprocessor.py:36-66, 69-80)."""
    from oracle import preprocess as OP
    from src.data import preprocessor as SP
    from src.models.models import get_model
    _, prep = get_model("cutoffvgg16")
    rng = np.random.default_rng(2)
    big = rng.integers(0, 256, (5, 150, 200, 3)).astype(np.uint8)
    assert np.array_equal(SP.resize_nearest(big, (128, 128)), OP.nearest_resize_tf(big, (128, 128)))
    assert SP.resize_nearest(big[:, :128, :128], (128, 128)) is not None
    n = SP.cfg['TRAIN']['BATCH_SIZE'] + 3
    frames = rng.integers(0, 256, (n, 128, 128, 3)).astype(np.uint8)
    labels = rng.integers(0, 2, n)
    p = SP.Preprocessor(prep)
    ds = p.prepare((frames, labels), shuffle=False, augment=False)      # uint8 out: scaler + resize run on the device
    batches = list(ds)
    assert [len(b[0]) for b in batches] == [SP.cfg['TRAIN']['BATCH_SIZE'], 3]
    assert batches[0][0].dtype.is_floating_point is False
    assert np.array_equal(np.concatenate([b[1].numpy() for b in batches]), np.eye(2, dtype=np.float32)[labels])
    aug = p.prepare((frames, labels), shuffle=True, augment=True)
    assert aug._aug is not None and aug.shuffle and aug.img_dim == (128, 128)
    assert aug._aug.cfg.rotation == SP.cfg['TRAIN']['DATA_AUG']['ROTATION_RANGE']
    # no scaler: x / 255 on the host (preprocessor.py:61); a foreign scaler is applied on the host as well
    x0 = next(iter(SP.Preprocessor(None).prepare((big, None))))[0].numpy()
    assert x0.shape == (5, 128, 128, 3) and np.allclose(x0, SP.resize_nearest(big, (128, 128)) / 255.)
    x1 = next(iter(SP.Preprocessor(lambda x: x * 2).prepare((big, None))))[0].numpy()
    assert np.allclose(x1, SP.resize_nearest(big, (128, 128)).astype(np.float32) * 2)
    with pytest.raises(NotImplementedError):
        SP.Preprocessor(None).prepare((frames, labels), augment=True)
    with pytest.raises(NotImplementedError):
        SP.Preprocessor(prep, resize_interpolation='bilinear')
    assert SP.RandomBrightness(0.3)(np.zeros((1, 2, 2, 3)))[0, 0, 0, 0] == P.reference_brightness_delta(0.3)

"""-m gpu: the reference-facing API (get_model / CutoffVGG16 / train.py / predict.py / gradcam.py mirrors)."""
import numpy as np
import pytest
import torch

from oracle import oracle, synth

pytestmark = pytest.mark.gpu
HP = {"LR_EXTRACT": 3e-4, "LR_FINETUNE": 9.3e-6, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7,
      "EXTRACT_EPOCHS": 2}


def _sets(n_train=96, n_val=32, bs=32):
    from ab_line_classifier_b200.data import ArrayDataset
    from src.models.models import get_model
    _, prep = get_model("cutoffvgg16")
    ftr, ltr = synth.synth_frames(n_train, seed=1)
    fva, lva = synth.synth_frames(n_val, seed=2)
    return (ArrayDataset(prep(ftr), synth.one_hot(ltr), bs, shuffle=True, seed=1),
            ArrayDataset(prep(fva), synth.one_hot(lva), bs), ltr, fva, lva)


def test_train_classifier_two_phase_predict_evaluate_save_load(tmp_path):
    from ab_line_classifier_b200.model import load_model, save_model
    from src.models.models import get_model
    from src import train as T
    model_def, prep = get_model("cutoffvgg16")
    train, val, ltr, fva, lva = _sets()
    model = T.train_classifier(model_def, train, val, HP, T.compute_output_bias(ltr), T.compute_class_weight(ltr),
                               verbose=False, epochs=3)
    assert model.metrics_names == ["loss", "accuracy", "auc", "precision_a_lines", "precision_b_lines",
                                   "recall_a_lines", "recall_b_lines"]
    assert model.history_extract.epoch == [0, 1] and model.history_finetune.epoch == [1, 2]
    assert {l.name for l in model.model.layers if l.trainable and l.kind in ("conv", "dense")} == {
        "block3_conv2", "block3_conv3", "output"}
    assert model.model.engine.opt_steps == {"adam": 2 * 3, "rmsprop": 2 * 3}
    res = model.evaluate(val, verbose=0)
    assert len(res) == 7 and np.isfinite(res).all()
    p = model.predict(val, verbose=0)
    assert p.shape == (32, 2) and np.allclose(p.sum(-1), 1, atol=1e-5)
    # the trained weights give the same probabilities in the oracle (bf16-faithful)
    w = model.model.get_weights()
    ref = oracle.forward(w, prep(fva), bf16=True)["probs"]
    assert np.abs(p - ref).max() < 5e-3
    # evaluate's loss is the unweighted mean categorical cross-entropy
    z = oracle.forward(w, prep(fva), bf16=True, as_torch=True)["logits"]
    assert abs(res[0] - float(oracle.loss_from_logits(z, synth.one_hot(lva)))) < 5e-3
    path = save_model(model.model, str(tmp_path / "m.npz"))
    again = load_model(path)
    assert np.array_equal(again.predict(val, verbose=0), p)
    assert np.allclose(again(prep(fva[:1])), p[:1], atol=0)        # model(x) callable, batch 1 (predict.py:440)


def test_head_only_phase_leaves_conv_weights_bit_identical():
    from ab_line_classifier_b200.model import CutoffVGG16
    train, val, *_ = _sets(64, 32)
    m = CutoffVGG16(dict(HP, EXTRACT_EPOCHS=1), [128, 128, 3], ["accuracy"], 2)
    before = m.model.get_weights()
    for layer in m.vgg16_layers:
        layer.trainable = False
    m.model.compile(optimizer=m.optimizer_extract, loss="categorical_crossentropy", metrics=["accuracy"])
    m.model.fit(train, epochs=1, verbose=0)
    after = m.model.get_weights()
    for k in before:
        if k.startswith("output"):
            assert not np.array_equal(before[k], after[k])
        else:
            assert np.array_equal(before[k], after[k]), k


def test_training_reduces_loss_on_separable_synthetic_frames():
    from ab_line_classifier_b200.model import Adam, CutoffVGG16
    train, val, *_ = _sets(256, 64, 64)
    m = CutoffVGG16(HP, [128, 128, 3], ["accuracy"], 2)
    m.model.compile(optimizer=Adam(1e-4), loss="categorical_crossentropy", metrics=["accuracy"])
    h = m.model.fit(train, epochs=4, validation_data=val, verbose=0)
    assert h.history["loss"][-1] < h.history["loss"][0]
    assert np.isfinite(h.history["val_loss"]).all()


def test_predict_set_clip_predictions_and_gradcam_explainer():
    from ab_line_classifier_b200.model import CutoffVGG16
    from src import predict as P
    from src.explainability.gradcam import GradCAMExplainer
    from src.models.models import get_model
    _, prep = get_model("cutoffvgg16")
    m = CutoffVGG16(HP, [128, 128, 3], [], 2)
    frames, labels = synth.synth_frames(70, seed=4)          # 70 = 64 + short batch of 6
    preds, probs = P.predict_set(m.model, prep, frames, threshold=0.5)
    assert len(preds) == 70 and probs.shape == (70, 2)
    assert preds == list((probs[:, 1] >= 0.5).astype(int))
    rows = P.compute_clip_predictions(m.model, prep, [("a", frames[:30], 0), ("b", frames[30:], 1)], 0.5, "average")
    assert rows[0]["score"] == pytest.approx(float(probs[:30, 1].mean()), abs=1e-6)
    rows = P.compute_clip_predictions(m.model, prep, [("a", frames[:30], 0)], 0.5, "sliding_window")
    assert rows[0]["score"] == pytest.approx(P.highest_avg_contiguous_pred_prob(probs[:30], 4)[1], abs=1e-6)
    ex = GradCAMExplainer(model=m.model)
    assert ex.last_conv_layer == "block3_conv3"
    heat, pr, overlays = ex.apply_gradcam(frames[:6])
    assert heat.shape == (6, 128, 128) and overlays.shape == (6, 128, 128, 3) and overlays.dtype == np.uint8
    ref, _ = oracle.gradcam_autograd(m.model.get_weights(), prep(frames[:2]), out_hw=(128, 128), bf16=True)
    assert np.abs(heat[:2] - ref).max() < 3e-2
    mean_ms, _ = P.clock_avg_runtime(m.model, 3, 5)
    assert mean_ms > 0


def test_array_dataset_host_and_device_residency_yield_the_same_batches():
    """Shuffled batches gathered by index into rotating pinned buffers (host) or by a device-side gather (dataset resident
    in HBM) are the same frames in the same order; buffers are recycled, so consumers must copy what they keep."""
    from ab_line_classifier_b200.data import ArrayDataset
    frames, labels = synth.synth_frames(70, seed=11)
    y = synth.one_hot(labels)
    host = ArrayDataset(frames, y, 16, shuffle=True, seed=5)
    dev = ArrayDataset(frames, y, 16, shuffle=True, seed=5, device="cuda")
    ref = np.random.default_rng(5).permutation(70)
    got = 0
    for (xh, yh), (xd, yd) in zip(host, dev):
        n = len(xh)
        assert xh.dtype == torch.uint8 and xh.is_pinned() and xd.is_cuda
        np.testing.assert_array_equal(xh.numpy(), frames[ref[got:got + n]])
        assert torch.equal(xd.cpu(), xh) and torch.equal(yd.cpu(), yh)
        got += n
    assert got == 70


@pytest.mark.parametrize("ext", [".npz", ".h5"])
def test_checkpoint_with_optimizer_state_resumes_training_bit_exactly(tmp_path, ext):
    """f3: save_model keeps the optimizer slots, their step counts and the dropout counter (Keras include_optimizer=True;
    the reference's save drops them, train.py:139-144), in .npz or in Keras' HDF5 layout.  Training 2 + 2 steps with a
    save / load in the middle must equal 4 uninterrupted steps bit for bit (deterministic kernels), and the reloaded
    model predicts exactly like the saved one."""
    from ab_line_classifier_b200.model import Adam, CutoffVGG16, load_model, save_model
    fr, lab = synth.synth_frames(16, seed=3)
    x, y = synth.vgg16_preprocess(fr), synth.one_hot(lab)

    def fresh():
        m = CutoffVGG16(HP, [128, 128, 3], ["accuracy"], 2, seed=77).model
        m.compile(optimizer=Adam(3e-4), loss="categorical_crossentropy", metrics=[])
        return m

    a = fresh()
    for _ in range(4):
        a.train_on_batch(x, y)
    b = fresh()
    for _ in range(2):
        b.train_on_batch(x, y)
    path = save_model(b, str(tmp_path / ("ckpt" + ext)))
    p_saved = b.predict(x, verbose=0)
    c = load_model(path)
    assert np.array_equal(c.predict(x, verbose=0), p_saved)
    assert c.engine.opt_steps == {"adam": 2} and c.engine.step_count == 2
    c.compile(optimizer=Adam(3e-4), loss="categorical_crossentropy", metrics=[])
    for _ in range(2):
        c.train_on_batch(x, y)
    assert torch.equal(c.engine.theta, a.engine.theta)
    for k in ("m", "v"):
        assert torch.equal(c.engine.opt_state[k], a.engine.opt_state[k])
    # weights only (Keras include_optimizer=False): loads, but Adam restarts
    d = load_model(save_model(b, str(tmp_path / ("weights_only" + ext)), include_optimizer=False))
    assert d.engine.opt_steps == {} and np.array_equal(d.predict(x, verbose=0), p_saved)

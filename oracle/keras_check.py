"""TEST INFRASTRUCTURE (oracle): pin the restatement against real Keras WHEN TensorFlow is importable.

TensorFlow 2.9 (reference requirements.txt:13) cannot be installed in this image, so today this script
reports "parity unpinned" and exits 0.  On a machine with TensorFlow it builds the reference graph exactly as
src/models/models.py:407-417 does (VGG16(include_top=False) layers[1:10] -> GlobalAveragePooling2D -> Dropout
-> Dense softmax), loads OUR seeded Glorot weights (oracle/synth.py, Keras names and layouts), and compares
logits / probabilities / one training-step loss with oracle/oracle.py on the committed golden frames; it also pins
the brightness constant, the projective-transform resampling and the metric definitions (section 8 f4).

    python oracle/keras_check.py            # prints PINNED ok / UNPINNED (TensorFlow missing) / mismatch
    python oracle/keras_check.py --emit     # additionally writes tests/golden/keras_pinned_seed1234.npz (+ a real Keras
                                            # .h5 of the same model): REAL Keras outputs on the committed golden frames --
                                            # probabilities, logits, class-weighted loss, two weight gradients, one Adam and
                                            # one RMSprop step.  Commit them: tests/test_oracle.py then checks the oracle
                                            # against Keras itself in every environment, and the parity claim is pinned
                                            # without TensorFlow in the image.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    try:
        import tensorflow as tf
    except Exception as e:  # noqa: BLE001
        print(f"UNPINNED: TensorFlow is not importable here ({type(e).__name__}); the oracle stays a restatement")
        return 0
    from oracle import oracle, synth
    from tensorflow.keras.applications.vgg16 import VGG16
    from tensorflow.keras.layers import Dense, Dropout, GlobalAveragePooling2D, Input
    from tensorflow.keras.models import Model

    params = synth.randomize_biases(synth.glorot_weights(10001))
    X_input = Input(shape=(128, 128, 3), name="input")
    vgg16 = VGG16(input_shape=(128, 128, 3), include_top=False, weights=None)      # models.py:409 with weights=None
    X = X_input
    layers = vgg16.layers[1:10]                                                      # models.py:410
    for layer in layers:                                                            # models.py:412-413
        X = layer(X)
    X = GlobalAveragePooling2D(name="global_avgpool")(X)                            # models.py:414
    X = Dropout(0.45)(X)                                                            # models.py:415
    Y = Dense(2, activation="softmax", name="output")(X)                            # models.py:416
    model = Model(inputs=X_input, outputs=Y)
    for layer in layers:
        if layer.name + "/kernel" in params:
            layer.set_weights([params[layer.name + "/kernel"], params[layer.name + "/bias"]])
    model.get_layer("output").set_weights([params["output/kernel"], params["output/bias"]])

    g = np.load(os.path.join(ROOT, "tests", "golden", "cutoffvgg16_seed1234.npz"))
    x = synth.vgg16_preprocess(g["frames_u8"])
    x_keras = tf.keras.applications.vgg16.preprocess_input(g["frames_u8"].astype(np.float32))
    ok = np.allclose(x, x_keras, atol=1e-4)
    print("preprocess_input (caffe mode) matches:", ok)
    p_keras = model(x, training=False).numpy()
    ref = oracle.forward(params, x)
    dp = float(np.abs(p_keras - ref["probs"]).max())
    print(f"max |softmax probs: Keras - oracle| = {dp:.3e}")
    y = synth.one_hot(np.arange(len(x)) % 2)
    loss_keras = float(tf.reduce_mean(tf.keras.losses.categorical_crossentropy(y, p_keras)).numpy())
    loss_ref, _, _ = oracle.loss_and_grads(params, x, y, None, np.ones((len(x), 256), np.float32), dropout=0.0)
    print(f"categorical_crossentropy: Keras {loss_keras:.6f}  oracle {float(loss_ref):.6f}")
    ok = ok and dp < 1e-4 and abs(loss_keras - float(loss_ref)) < 1e-4
    if "--emit" in sys.argv:
        emit_golden(tf, model, layers, params, x, y)
    ok = pin_input_pipeline_and_metrics(tf) and ok
    print("PINNED ok" if ok else "MISMATCH between Keras and the oracle")
    return 0 if ok else 1


def emit_golden(tf, model, layers, params, x, y):
    """Real-Keras golden vectors for tests/test_oracle.py::test_oracle_matches_keras_golden_vectors (consumed when the
    file exists).  Everything is computed by Keras/TF ops only; the oracle is not involved."""
    cw = np.array([0.75, 1.5], np.float32)
    xs = tf.constant(x)
    with tf.GradientTape() as tape:
        p = model(xs, training=False)
        # Keras: loss = sum_i w_i * CCE_i / batch (class_weight -> sample weights, SUM_OVER_BATCH_SIZE; train.py:135-136)
        li = tf.keras.losses.categorical_crossentropy(y, p)
        loss = tf.reduce_sum(li * tf.reduce_sum(y * cw, axis=-1)) / float(len(x))
    out_layer, c33 = model.get_layer("output"), [l for l in layers if l.name == "block3_conv3"][0]
    g_out, g_c33k, g_c33b = tape.gradient(loss, [out_layer.kernel, c33.kernel, c33.bias])
    logits = tf.keras.Model(model.inputs, out_layer.input)(xs, training=False).numpy() @ params["output/kernel"] + params["output/bias"]
    # one optimizer step each on a small tensor (Keras 2.9 Adam / RMSprop defaults as models.py:397-398 builds them)
    rng = np.random.default_rng(5)
    th0, gr = rng.normal(size=64).astype(np.float32), (rng.normal(size=64) * 0.01).astype(np.float32)
    steps = {}
    for name, opt in (("adam", tf.keras.optimizers.Adam(learning_rate=3e-4)), ("rmsprop", tf.keras.optimizers.RMSprop(learning_rate=9.3e-6))):
        v = tf.Variable(th0)
        for _ in range(3):
            opt.apply_gradients([(tf.constant(gr), v)])
        steps[name] = v.numpy()
    path = os.path.join(ROOT, "tests", "golden", "keras_pinned_seed1234.npz")
    np.savez(path, probs=p.numpy(), logits=logits, loss_class_weighted=np.float32(loss.numpy()),
             grad_output_kernel=g_out.numpy(), grad_block3_conv3_kernel=g_c33k.numpy(), grad_block3_conv3_bias=g_c33b.numpy(),
             opt_theta0=th0, opt_grad=gr, adam_3_steps=steps["adam"], rmsprop_3_steps=steps["rmsprop"],
             tf_version=np.array(tf.__version__), keras_version=np.array(tf.keras.__version__))
    model.save(os.path.join(ROOT, "tests", "golden", "keras_cutoffvgg16_seed10001.h5"))     # a REAL Keras .h5 for the f3 reader
    print("wrote", path, "and tests/golden/keras_cutoffvgg16_seed10001.h5")


def pin_input_pipeline_and_metrics(tf):
    """The pieces either side of the network: the reference's brightness constant
    (preprocessor.py:103), the projective-transform resampling behind RandomZoom/Translation/Rotation
    (preprocessor.py:26-29) and the Keras metric definitions (train.py:109-115) against oracle/augment.py and
    ab_line_classifier_b200/metrics.py."""
    from ab_line_classifier_b200 import metrics as M
    from oracle import augment as A
    ok = True
    d_tf = float(tf.image.stateless_random_brightness(tf.zeros((1, 2, 2, 3)), 0.3, (123, 0)).numpy().ravel()[0])
    d_or = float(A.stateless_brightness_delta(0.3))
    print(f"stateless_random_brightness delta: TF {d_tf:.9f}  oracle {d_or:.9f}")
    ok = ok and abs(d_tf - d_or) < 1e-7
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (1, 64, 96, 3)).astype(np.float32)
    for t in (A.zoom_transform(0.93, 64, 96), A.translation_transform(0.15, -0.1, 64, 96), A.rotation_transform(0.7, 64, 96)):
        got = tf.raw_ops.ImageProjectiveTransformV3(images=img, transforms=t[None], output_shape=[64, 96],
                                                    fill_value=0.0, interpolation="BILINEAR", fill_mode="CONSTANT").numpy()[0]
        err = float(np.abs(got - A.projective_bilinear(img[0], t)).max())
        print(f"ImageProjectiveTransformV3 vs oracle: max |diff| = {err:.3e}")
        ok = ok and err < 1e-3
    z = rng.normal(size=(500, 2)).astype(np.float32) * 2
    p = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(np.float32)
    y = np.eye(2, dtype=np.float32)[rng.integers(0, 2, 500)]
    pairs = [(tf.keras.metrics.AUC(name="auc"), M.AUC()), (tf.keras.metrics.CategoricalAccuracy(), M.CategoricalAccuracy()),
             (tf.keras.metrics.Precision(thresholds=0.5, class_id=1), M.Precision(thresholds=0.5, class_id=1)),
             (tf.keras.metrics.Recall(thresholds=0.5, class_id=0), M.Recall(thresholds=0.5, class_id=0))]
    for km, mm in pairs:
        km.update_state(y, p)
        a, b = float(km.result().numpy()), mm.result(y, p)
        print(f"{type(mm).__name__}: Keras {a:.6f}  here {b:.6f}")
        ok = ok and abs(a - b) < 1e-5
    return ok


if __name__ == "__main__":
    sys.exit(main())

"""TEST INFRASTRUCTURE (oracle): pin the restatement against real Keras WHEN TensorFlow is importable.

TensorFlow 2.9 (reference requirements.txt:13) cannot be installed in this image, so today this script
reports "parity unpinned" and exits 0.  On a machine with TensorFlow it builds the reference graph exactly as
src/models/models.py:407-417 does (VGG16(include_top=False) layers[1:10] -> GlobalAveragePooling2D -> Dropout
-> Dense softmax), loads OUR seeded Glorot weights (oracle/synth.py, Keras names and layouts), and compares
logits / probabilities / one training-step loss with oracle/oracle.py on the committed golden frames.

    python oracle/keras_check.py            # prints PINNED ok / UNPINNED (TensorFlow missing) / mismatch
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
    print("PINNED ok" if ok else "MISMATCH between Keras and the oracle")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

"""Writes tests/golden/keras_layout_cutoffvgg16_8x8.h5: a cutoffvgg16 model (CUTOFF_LAYER 4: block1_conv1, block1_conv2,
block1_pool; 8x8 input, 2 classes) in the HDF5 layout Keras' save_model produces -- /model_weights/<layer>/<layer>/
kernel:0 datasets, layer_names / weight_names attributes, model_config JSON -- plus the expected arrays as .npz.
TensorFlow / h5py are not installable in this image, so the file is produced by the repo's own HDF5 writer
(ab_line_classifier_b200/h5lite.py) following the HDF5 file-format specification; on a machine with TensorFlow,
oracle/keras_check.py --emit writes the same kind of file with real Keras (tests/golden/keras_*.h5).
    python tests/golden/make_keras_layout_fixture.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ab_line_classifier_b200.h5lite import H5Writer  # noqa: E402
from oracle import synth  # noqa: E402

params = synth.randomize_biases(synth.glorot_weights(4321, cutoff_layer=4))
layers = [{"class_name": "InputLayer", "name": "input", "config": {"name": "input", "batch_input_shape": [None, 8, 8, 3]}},
          {"class_name": "Conv2D", "name": "block1_conv1", "config": {"name": "block1_conv1", "filters": 64}},
          {"class_name": "Conv2D", "name": "block1_conv2", "config": {"name": "block1_conv2", "filters": 64}},
          {"class_name": "MaxPooling2D", "name": "block1_pool", "config": {"name": "block1_pool"}},
          {"class_name": "GlobalAveragePooling2D", "name": "global_avgpool", "config": {"name": "global_avgpool"}},
          {"class_name": "Dropout", "name": "dropout", "config": {"name": "dropout", "rate": 0.45}},
          {"class_name": "Dense", "name": "output", "config": {"name": "output", "units": 2, "activation": "softmax"}}]
w = H5Writer()
w.attr(w.root, "keras_version", b"2.9.0")
w.attr(w.root, "backend", b"tensorflow")
w.attr(w.root, "model_config", json.dumps({"class_name": "Functional", "config": {"name": "model", "layers": layers}}).encode())
mw = w.group(w.root, "model_weights")
w.attr(mw, "layer_names", np.array([l["name"].encode() for l in layers]))
w.attr(mw, "backend", b"tensorflow")
w.attr(mw, "keras_version", b"2.9.0")
for l in layers:
    g = w.group(mw, l["name"])
    mine = [k for k in params if k.split("/")[0] == l["name"]]
    if mine:
        w.attr(g, "weight_names", np.array([(k + ":0").encode() for k in mine]))
        for k in mine:
            w.dataset(w.group(g, l["name"]), k.split("/")[1] + ":0", params[k])
w.save(os.path.join(HERE, "keras_layout_cutoffvgg16_8x8.h5"))
np.savez(os.path.join(HERE, "keras_layout_cutoffvgg16_8x8_expected.npz"), **params)
print("wrote", os.path.join(HERE, "keras_layout_cutoffvgg16_8x8.h5"))

"""Offline converter: a Keras HDF5 model / weights file -> the .npz checkpoint format of ab_line_classifier_b200.

  python tools/keras_h5_to_npz.py results/models/cutoffvgg16_final_cropped.h5 out.npz      # a model train.py saved
  python tools/keras_h5_to_npz.py ~/.keras/models/vgg16_weights_tf_dim_ordering_tf_kernels_notop.h5 vgg16_notop.npz \
         --weights-only                                                                      # VGG16(weights='imagenet')

Reads the file with the built-in pure-Python HDF5 reader (ab_line_classifier_b200/h5lite.py; h5py is not required).
The .h5 can also be passed to load_model / weights_path / ABL_VGG16_WEIGHTS directly -- the .npz is just faster to open.
Reference: save_model(model.model, '...h5') src/train.py:139-144, load_model src/predict.py:50, by-name transplant
src/models/model_utils.py:6-23."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ab_line_classifier_b200.model import _arch_from_keras, read_keras_h5  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--weights-only", action="store_true", help="no architecture block (e.g. the VGG16 notop file)")
    ap.add_argument("--img-dim", type=int, nargs=2, default=None, help="H W when the file has no model_config")
    a = ap.parse_args()
    weights, cfg, opt = read_keras_h5(a.src)
    out = {"weights/" + k: v for k, v in weights.items()}
    if not a.weights_only:
        ishape = None if a.img_dim is None else (a.img_dim[0], a.img_dim[1], 3)
        ishape, n_classes, cutoff, dropout = _arch_from_keras(weights, cfg, ishape)
        out.update({"config/input_shape": np.asarray(ishape), "config/n_classes": n_classes, "config/dropout": dropout,
                    "config/cutoff_layer": cutoff})
    out.update({"optimizer/" + k: v for k, v in opt.items()})
    np.savez(a.dst, **out)
    for k, v in weights.items():
        print(f"{k:28s} {tuple(v.shape)}")
    print(f"wrote {a.dst}: {len(weights)} arrays" + ("" if a.weights_only else f", cutoff_layer {cutoff}, input {ishape}"))


if __name__ == "__main__":
    main()

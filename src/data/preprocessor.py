"""Counterpart of reference src/data/preprocessor.py for the cutoffvgg16 path.

``Preprocessor(scale_fn, resize_interpolation).prepare(ds, shuffle=False, augment=False)`` keeps the reference's
contract (preprocessor.py:16-66): shuffle over the whole set, resize every frame to ``DATA.IMG_DIM`` with
nearest-neighbour interpolation, one-hot the labels, batch by ``TRAIN.BATCH_SIZE`` (short last batch), optionally run
the augmentation Sequential on each batch, then the model's input scaler.  Differences, all forced by the scope:
  * ``ds`` is ``(frames, labels)`` -- decoded RGB frames ``[N,h,w,3]`` (uint8, any size) and integer class ids -- instead
    of a ``tf.data.Dataset`` of JPEG paths: JPEG decode stays in the caller's dataset plumbing;
  * resize, augmentation and the ``vgg16.preprocess_input`` scaler run on the B200 (``abl_resize_nearest_u8``,
    ``abl_augment_frames``, the uint8 path of block1_conv1), so the returned ``ArrayDataset`` yields raw uint8 frames
    (plain path) or augmented, scaled fp32 device batches (``augment=True``); any other ``scale_fn`` is applied on the
    host after a host-side nearest-neighbour resize (no augmentation then).
"""
import os

import numpy as np
import yaml

from ab_line_classifier_b200.augment import reference_brightness_delta
from ab_line_classifier_b200.data import ArrayDataset

cfg = yaml.full_load(open(os.getcwd() + "/config.yml", 'r'))

NEAREST_NEIGHBOR = 'nearest'        # tensorflow.image.ResizeMethod.NEAREST_NEIGHBOR


def resize_nearest(frames, img_dim):
    """``tf.image.resize(image, IMG_DIM, method=NEAREST_NEIGHBOR)`` of preprocessor.py:76-80 on the host (TF2
    half-pixel-centre rule: src = min(floor((dst + 0.5) * in / out), in - 1), evaluated in fp32)."""
    frames = np.asarray(frames)
    h, w = frames.shape[1:3]
    H, W = int(img_dim[0]), int(img_dim[1])
    if (h, w) == (H, W):
        return frames
    f = np.float32
    iy = np.minimum(np.floor((np.arange(H, dtype=f) + f(0.5)) * (f(h) / f(H))).astype(np.int64), h - 1)
    ix = np.minimum(np.floor((np.arange(W, dtype=f) + f(0.5)) * (f(w) / f(W))).astype(np.int64), w - 1)
    return frames[:, iy][:, :, ix]


class RandomBrightness:
    """The reference's own layer (preprocessor.py:83-104): ``tf.image.stateless_random_brightness(x, factor, (123, 0))``
    -- a stateless draw with a constant seed, i.e. the same delta on every call, added without clipping."""

    def __init__(self, factor=0.5, **kwargs):
        self.factor = factor
        self.delta = reference_brightness_delta(factor)

    def call(self, inputs, training=True):
        return np.asarray(inputs, np.float32) + self.delta

    __call__ = call


class Preprocessor:
    def __init__(self, scale_fn=None, resize_interpolation=NEAREST_NEIGHBOR):
        self.batch_size = cfg['TRAIN']['BATCH_SIZE']
        self.n_classes = len(cfg['DATA']['CLASSES'])
        self.img_dim = tuple(cfg['DATA']['IMG_DIM'])
        self.data_aug_cfg = dict(cfg['TRAIN']['DATA_AUG'])
        self.input_scaler = scale_fn
        if resize_interpolation != NEAREST_NEIGHBOR:
            raise NotImplementedError("only NEAREST_NEIGHBOR resizing (the reference's default) is implemented")
        self.interpolation = resize_interpolation
        self.seed = cfg['TRAIN'].get('SEED', 0)

    def _device_scaler(self):
        """True if the scaler is vgg16.preprocess_input, which the device applies itself to uint8 frames."""
        return getattr(self.input_scaler, "__name__", "") == "vgg16_preprocess"

    def prepare(self, ds, shuffle=False, augment=False):
        frames, labels = ds
        frames = np.asarray(frames)
        y = None if labels is None else np.eye(self.n_classes, dtype=np.float32)[np.asarray(labels).astype(int)]
        if self._device_scaler() and frames.dtype == np.uint8:
            # resize (+ augmentation) + scaler on the device; frames cross PCIe as uint8
            return ArrayDataset(frames, y, self.batch_size, shuffle=shuffle, seed=self.seed,
                                augment=self.data_aug_cfg if augment else None, img_dim=self.img_dim)
        if augment:
            raise NotImplementedError("augment=True needs uint8 frames and the vgg16 input scaler (the cutoffvgg16 path)")
        x = resize_nearest(frames, self.img_dim).astype(np.float32)
        x = x / 255. if self.input_scaler is None else self.input_scaler(x)
        return ArrayDataset(x, y, self.batch_size, shuffle=shuffle, seed=self.seed)

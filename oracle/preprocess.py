"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's frame preparation.

`Preprocessor._parse_fn` (reference src/data/preprocessor.py:69-80): decode_jpeg -> float32 ->
`tf.image.resize(image, IMG_DIM, method=NEAREST_NEIGHBOR)`; then `vgg16.preprocess_input` (preprocessor.py:60-63,
models.py:13).  TensorFlow 2's resize uses half-pixel centres: source index = min(floor((dst + 0.5) * in / out), in - 1)
with the scale and the product in float32 (third-party semantics: tensorflow/core/kernels/image/resize_nearest_neighbor_op).
Only tests/ may import this module.
"""
import numpy as np


def nearest_resize_tf(frames, out_hw):
    """frames [B,h,w,C] -> [B,H,W,C] with TF2 nearest-neighbour semantics."""
    frames = np.asarray(frames)
    _, h, w, _ = frames.shape
    H, W = out_hw
    sy, sx = np.float32(h) / np.float32(H), np.float32(w) / np.float32(W)
    iy = np.minimum(np.floor((np.arange(H, dtype=np.float32) + np.float32(0.5)) * sy).astype(np.int64), h - 1)
    ix = np.minimum(np.floor((np.arange(W, dtype=np.float32) + np.float32(0.5)) * sx).astype(np.int64), w - 1)
    return frames[:, iy][:, :, ix]


def vgg16_preprocess(frames_rgb):
    """keras.applications.vgg16.preprocess_input (caffe mode): RGB -> BGR, subtract the ImageNet BGR means."""
    x = np.asarray(frames_rgb, np.float32)[..., ::-1]
    return np.ascontiguousarray(x - np.array([103.939, 116.779, 123.68], np.float32))

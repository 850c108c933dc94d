"""Seeded synthetic lung-ultrasound-like frames (pure NumPy; shared by bench.py, src/train.py and the
test oracle).  Frames mimic masked sector scans: class 0 = A-lines (horizontal arcs), class 1 = B-lines
(vertical rays).  ``vgg16_preprocess`` is Keras' caffe-mode preprocess_input (reference models.py:13)."""
import numpy as np

VGG_MEAN_BGR = np.array([103.939, 116.779, 123.68], dtype=np.float32)


def synth_frames(n, seed=1234, h=128, w=128):
    """uint8 [n,h,w,3] grey (R=G=B) fan-sector frames; labels y = idx % 2 (0 A-lines, 1 B-lines)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ang = np.arctan2(xx - (w - 1) / 2.0, yy + 1.0)
    rad = np.hypot(xx - (w - 1) / 2.0, yy + 1.0)
    sector = (np.abs(ang) < np.deg2rad(35.0)) & (rad < 0.98 * h)
    frames = np.zeros((n, h, w), dtype=np.float32)
    labels = (np.arange(n) % 2).astype(np.int64)
    for i in range(n):
        img = rng.rayleigh(40.0, size=(h, w)).astype(np.float32)
        pleura = 0.15 * h + rng.uniform(-3, 3)
        img += 120.0 * np.exp(-0.5 * ((rad - pleura) / 1.5) ** 2)
        if labels[i] == 0:  # A-lines: horizontal arcs at equal depth spacing
            k = rng.integers(3, 6)
            for j in range(1, k + 1):
                img += (90.0 / j) * np.exp(-0.5 * ((rad - pleura * (j + 1)) / 1.5) ** 2)
        else:  # B-lines: vertical rays from the pleural line
            k = rng.integers(1, 5)
            for _ in range(k):
                a0 = rng.uniform(-0.5, 0.5)
                img += 110.0 * np.exp(-0.5 * ((ang - a0) / 0.03) ** 2) * (rad > pleura)
        frames[i] = np.where(sector, img, 0.0)
    frames = np.clip(frames, 0, 255).astype(np.uint8)
    return np.repeat(frames[..., None], 3, axis=-1), labels


def vgg16_preprocess(frames_rgb):
    """Keras ``vgg16.preprocess_input`` (caffe mode): RGB->BGR, subtract the ImageNet BGR means
    (reference binds it at models.py:13,44; applied in preprocessor.py:60-63)."""
    x = np.asarray(frames_rgb, dtype=np.float32)[..., ::-1]
    return np.ascontiguousarray(x - VGG_MEAN_BGR)


def one_hot(labels, n_classes=2):
    y = np.zeros((len(labels), n_classes), dtype=np.float32)
    y[np.arange(len(labels)), labels] = 1.0
    return y

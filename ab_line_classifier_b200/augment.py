"""Training-time augmentation of a batch on the device: host-side mirror of the ``tf.keras.Sequential`` the reference
builds in ``Preprocessor.__init__`` (src/data/preprocessor.py:25-32) and applies to every training batch
(preprocessor.py:56-57), followed by the input scaler (preprocessor.py:60-63).

The host only DRAWS the random parameters (Keras 2.9 distributions, config keys ``TRAIN.DATA_AUG.*``) and turns them
into the three output->input affine maps Keras hands to ImageProjectiveTransformV3; the resampling, flip, brightness
and vgg16 preprocessing run in one CUDA kernel (``abl_augment_frames``).  TensorFlow's random stream cannot be
reproduced, so the draws come from a seeded NumPy generator; given the same draws the result is what Keras computes.

Brightness is the exception: the reference does NOT use Keras' RandomBrightness but its own layer
(preprocessor.py:83-104), ``tf.image.stateless_random_brightness(inputs, factor, (123, 0))``: ONE delta in
[-factor, factor) from a stateless generator with a constant seed -- the same value for every frame of every batch --
added to the float image without scaling to the 0..255 range and without clipping.  ``reference_brightness_delta``
restates TF's stateless Philox sampling of that constant.
"""
import numpy as np
import torch

from . import ops

N_PARAMS = 20      # [zoom a0..b2 | translation a0..b2 | rotation a0..b2 | flip | brightness delta]


_M32 = 0xFFFFFFFF


def _philox4x32_10(counter, key):
    """One block of Philox4x32-10 as tensorflow/core/lib/random/philox_random.h computes it (uint32 lists in and out)."""
    c, k = [int(v) & _M32 for v in counter], [int(v) & _M32 for v in key]
    for rnd in range(10):
        if rnd:
            k = [(k[0] + 0x9E3779B9) & _M32, (k[1] + 0xBB67AE85) & _M32]
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & _M32, p1 & _M32, ((p0 >> 32) ^ c[3] ^ k[1]) & _M32, p0 & _M32]
    return c


def reference_brightness_delta(factor, seed=(123, 0)):
    """``tf.image.stateless_random_brightness(x, factor, seed)``'s delta: stateless_random_uniform(shape=[],
    minval=-factor, maxval=factor, seed) with TF 2.9's default Philox path -- the seed pair is scrambled into a key and
    a counter by one Philox block under a fixed key (stateless_random_ops.cc GenerateKey), the first 32-bit output of
    the next block becomes a float in [0, 1) by mantissa stuffing, and the result is u * (max - min) + min in fp32."""
    f = np.float32
    s0, s1 = int(seed[0]) & 0xFFFFFFFFFFFFFFFF, int(seed[1]) & 0xFFFFFFFFFFFFFFFF
    mix = _philox4x32_10([s0 & _M32, s0 >> 32, s1 & _M32, s1 >> 32], [0x3EC8F720, 0x02461E29])
    out = _philox4x32_10([0, 0, mix[2], mix[3]], [mix[0], mix[1]])
    u = np.array([(127 << 23) | (out[0] & 0x7FFFFF)], np.uint32).view(f)[0] - f(1.0)
    lo, hi = f(-factor), f(factor)
    return f(u * (hi - lo) + lo)


class AugmentConfig:
    """``cfg['TRAIN']['DATA_AUG']`` as the reference reads it: ZOOM_RANGE, HEIGHT_SHIFT_RANGE, WIDTH_SHIFT_RANGE,
    ROTATION_RANGE, BRIGHTNESS_RANGE (preprocessor.py:26-31).  HORIZONTAL_FLIP and SHEAR_RANGE exist in config.yml but
    the reference never reads them: RandomFlip("horizontal") is always part of the pipeline."""

    def __init__(self, cfg=None):
        cfg = dict(cfg or {})
        self.zoom = float(cfg.get("ZOOM_RANGE", 0.1))
        self.height_shift = float(cfg.get("HEIGHT_SHIFT_RANGE", 0.2))
        self.width_shift = float(cfg.get("WIDTH_SHIFT_RANGE", 0.2))
        self.rotation = float(cfg.get("ROTATION_RANGE", 45))
        self.brightness = float(cfg.get("BRIGHTNESS_RANGE", 0.3))
        for k in ("zoom", "height_shift", "width_shift", "rotation"):
            if getattr(self, k) < 0:
                raise ValueError(f"DATA_AUG {k} range cannot be negative")      # as the Keras layers raise
        if self.brightness < 0:
            raise ValueError("max_delta must be non-negative.")                 # tf.image.stateless_random_brightness
        self.brightness_delta = reference_brightness_delta(self.brightness)


def draw_raw(rng, B, cfg):
    """The random draws of one batch, one per frame, in Keras' units: zoom factor, (dx, dy) as fractions of (W, H),
    rotation angle in radians, flip (0/1), brightness delta.  Keras: RandomZoom draws U(1-z, 1+z) once and uses it
    for both axes (width_factor=None); RandomTranslation U(-f, f); RandomRotation U(-f*2pi, f*2pi); RandomFlip flips
    each frame with probability 1/2.  The brightness delta is the reference layer's constant (module docstring)."""
    f = np.float32
    return {
        "zoom": rng.uniform(1.0 - cfg.zoom, 1.0 + cfg.zoom, B).astype(f),
        "dx": rng.uniform(-cfg.width_shift, cfg.width_shift, B).astype(f),
        "dy": rng.uniform(-cfg.height_shift, cfg.height_shift, B).astype(f),
        "angle": rng.uniform(-cfg.rotation * 2.0 * np.pi, cfg.rotation * 2.0 * np.pi, B).astype(f),
        "flip": (rng.random(B) >= 0.5).astype(f),
        "brightness": np.full(B, cfg.brightness_delta, f),
    }


def params_from_raw(raw, H, W):
    """fp32 [B,20] kernel parameters from the raw draws: Keras' get_zoom_matrix / get_translation_matrix /
    get_rotation_matrix (first six entries of each 8-vector; the projective terms are 0), evaluated in fp32."""
    f = np.float32
    B = len(raw["zoom"])
    p = np.zeros((B, N_PARAMS), f)
    h1, w1 = f(H - 1), f(W - 1)
    z = raw["zoom"].astype(f)
    p[:, 0] = z
    p[:, 2] = (w1 / f(2.0)) * (f(1.0) - z)
    p[:, 4] = z
    p[:, 5] = (h1 / f(2.0)) * (f(1.0) - z)
    p[:, 6] = 1.0
    p[:, 8] = -(raw["dx"].astype(f) * f(W))
    p[:, 10] = 1.0
    p[:, 11] = -(raw["dy"].astype(f) * f(H))
    a = raw["angle"].astype(f)
    cos, sin = np.cos(a).astype(f), np.sin(a).astype(f)
    p[:, 12] = cos
    p[:, 13] = -sin
    p[:, 14] = (w1 - (cos * w1 - sin * h1)) / f(2.0)
    p[:, 15] = sin
    p[:, 16] = cos
    p[:, 17] = (h1 - (sin * w1 + cos * h1)) / f(2.0)
    p[:, 18] = raw["flip"].astype(f)
    p[:, 19] = raw["brightness"].astype(f)                  # added as is: tf.image.adjust_brightness on a float image
    return p


class DeviceAugmenter:
    """Draws per-frame parameters and launches the augmentation kernel.  ``__call__(frames)`` takes RGB frames
    [B,H,W,3] (uint8 or fp32 in [0,255], host or device) at IMG_DIM and returns a device fp32 tensor: preprocessed
    (BGR, mean-subtracted) when ``preprocess`` -- what the model's fit() consumes -- else the augmented RGB frames."""

    N_BUFFERS = 4        # rotating outputs: a batch stays valid while the next three are produced

    def __init__(self, cfg=None, seed=0, device=None, preprocess=True):
        self.cfg = cfg if isinstance(cfg, AugmentConfig) else AugmentConfig(cfg)
        self.rng = np.random.default_rng(seed)
        self.device = torch.device(device if device is not None else "cuda")
        self.preprocess = preprocess
        self._out, self._k, self._ws = [None] * self.N_BUFFERS, 0, None
        self.last_params = None

    def __call__(self, frames, params=None):
        x = frames if torch.is_tensor(frames) else torch.from_numpy(np.ascontiguousarray(frames))
        if x.dtype != torch.uint8:
            x = x.to(torch.float32)
        if x.dim() != 4 or x.shape[3] != 3:
            raise ValueError(f"expected frames [B,H,W,3], got {tuple(x.shape)}")
        B, H, W, _ = x.shape
        if params is None:
            params = params_from_raw(draw_raw(self.rng, B, self.cfg), H, W)
        self.last_params = params
        if B == 0:
            return torch.empty(0, H, W, 3, device=self.device, dtype=torch.float32)
        x = x.to(self.device, non_blocking=True).contiguous()
        p = torch.from_numpy(np.ascontiguousarray(params, np.float32)).to(self.device, non_blocking=True)
        out = self._out[self._k]
        if out is None or out.shape[0] < B or tuple(out.shape[1:]) != (H, W, 3):
            out = self._out[self._k] = torch.empty(B, H, W, 3, device=self.device, dtype=torch.float32)
        self._k = (self._k + 1) % self.N_BUFFERS
        need = ops.augment_workspace_bytes(B, H, W)
        if need and (self._ws is None or self._ws.numel() < need):
            self._ws = torch.empty(need, device=self.device, dtype=torch.uint8)
        ops.augment_frames(x, p, out[:B], self.preprocess, self._ws if need else None)
        return out[:B]

"""Training-time augmentation of a batch on the device: host-side mirror of the ``tf.keras.Sequential`` the reference
builds in ``Preprocessor.__init__`` (src/data/preprocessor.py:25-32) and applies to every training batch
(preprocessor.py:56-57), followed by the input scaler (preprocessor.py:60-63).

The host only DRAWS the random parameters (Keras 2.9 distributions, config keys ``TRAIN.DATA_AUG.*``) and turns them
into the three output->input affine maps Keras hands to ImageProjectiveTransformV3; the resampling, flip, brightness
and vgg16 preprocessing run in one CUDA kernel (``abl_augment_frames``).  TensorFlow's random stream cannot be
reproduced, so the draws come from a seeded NumPy generator; given the same draws the result is what Keras computes.
"""
import numpy as np
import torch

from . import ops

N_PARAMS = 20      # [zoom a0..b2 | translation a0..b2 | rotation a0..b2 | flip | brightness delta]


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
        if not 0.0 <= self.brightness <= 1.0:
            raise ValueError("BRIGHTNESS_RANGE must be within [0, 1]")          # RandomBrightness' own check


def draw_raw(rng, B, cfg):
    """The random draws of one batch, one per frame, in Keras' units: zoom factor, (dx, dy) as fractions of (W, H),
    rotation angle in radians, flip (0/1), brightness factor.  Keras: RandomZoom draws U(1-z, 1+z) once and uses it
    for both axes (width_factor=None); RandomTranslation U(-f, f); RandomRotation U(-f*2pi, f*2pi); RandomFlip flips
    each frame with probability 1/2; RandomBrightness U(-f, f) per frame."""
    f = np.float32
    return {
        "zoom": rng.uniform(1.0 - cfg.zoom, 1.0 + cfg.zoom, B).astype(f),
        "dx": rng.uniform(-cfg.width_shift, cfg.width_shift, B).astype(f),
        "dy": rng.uniform(-cfg.height_shift, cfg.height_shift, B).astype(f),
        "angle": rng.uniform(-cfg.rotation * 2.0 * np.pi, cfg.rotation * 2.0 * np.pi, B).astype(f),
        "flip": (rng.random(B) >= 0.5).astype(f),
        "brightness": rng.uniform(-cfg.brightness, cfg.brightness, B).astype(f),
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
    p[:, 19] = raw["brightness"].astype(f) * f(255.0)       # value_range (0, 255)
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

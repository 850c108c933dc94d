"""TEST INFRASTRUCTURE ONLY (never imported by the product).  NumPy restatement of the augmentation pipeline the
reference applies to training batches -- ``tf.keras.Sequential([RandomZoom(z, fill_mode='constant'),
RandomTranslation(h, w, fill_mode='constant'), RandomRotation(r, fill_mode='constant'), RandomFlip("horizontal"),
RandomBrightness(factor=b)])`` (src/data/preprocessor.py:25-32, 56-57; RandomBrightness is the reference's OWN layer,
preprocessor.py:83-104) -- followed by ``vgg16.preprocess_input`` (preprocessor.py:60-63).

The arithmetic lives in third-party code absent from /root/reference: tensorflow==2.9.0 (requirements.txt:13) with its
bundled Keras 2.9 (keras/layers/preprocessing/image_preprocessing.py: get_zoom_matrix, get_translation_matrix,
get_rotation_matrix, RandomFlip), the ImageProjectiveTransformV3 CPU kernel (tensorflow/core/kernels/image/
image_ops.h: ProjectiveGenerator, bilinear_interpolation, read_with_fill_value), tf.image.stateless_random_brightness /
adjust_brightness (python/ops/image_ops_impl.py) and the stateless Philox sampler behind it (core/kernels/
stateless_random_ops.cc GenerateKey, core/lib/random/philox_random.h, random_distributions.h Uint32ToFloat).  Their
published algorithm is restated below.  PARITY UNPINNED: TensorFlow cannot be installed here and the reference holds no
fixtures for this path; the restatement is anchored on closed-form cases in tests/test_augment_metrics.py.
The random draws are inputs (TF's random stream is not reproducible outside TF), except the brightness delta, which
the reference draws from a stateless generator with a constant seed."""
import numpy as np

F = np.float32
BGR_MEAN = np.array([103.939, 116.779, 123.68], F)


def zoom_transform(zoom, H, W):
    """get_zoom_matrix with width_zoom == height_zoom (RandomZoom(width_factor=None)): 8-vector [a0 a1 a2 b0 b1 b2 c0 c1]."""
    z = F(zoom)
    x_off = (F(W - 1) / F(2.0)) * (F(1.0) - z)
    y_off = (F(H - 1) / F(2.0)) * (F(1.0) - z)
    return np.array([z, 0, x_off, 0, z, y_off, 0, 0], F)


def translation_transform(dx_frac, dy_frac, H, W):
    """get_translation_matrix of (width_translate, height_translate) = (dx_frac*W, dy_frac*H)."""
    return np.array([1, 0, -(F(dx_frac) * F(W)), 0, 1, -(F(dy_frac) * F(H)), 0, 0], F)


def rotation_transform(angle, H, W):
    """get_rotation_matrix: rotation by ``angle`` radians about the frame centre."""
    a = F(angle)
    c, s = F(np.cos(a)), F(np.sin(a))
    x_off = (F(W - 1) - (c * F(W - 1) - s * F(H - 1))) / F(2.0)
    y_off = (F(H - 1) - (s * F(W - 1) + c * F(H - 1))) / F(2.0)
    return np.array([c, -s, x_off, s, c, y_off, 0, 0], F)


def projective_bilinear(img, t):
    """ImageProjectiveTransformV3(interpolation=BILINEAR, fill_mode=CONSTANT, fill_value=0) of one frame [H,W,C] fp32:
    out[y,x] = bilinear(img, (a0 x + a1 y + a2)/k, (b0 x + b1 y + b2)/k), k = c0 x + c1 y + 1; neighbours outside
    the frame read as 0.  All arithmetic in fp32, each operation rounded (as NumPy does)."""
    img = np.asarray(img, F)
    H, W = img.shape[:2]
    t = np.asarray(t, F)
    ox, oy = np.meshgrid(np.arange(W, dtype=F), np.arange(H, dtype=F))
    k = (t[6] * ox + t[7] * oy) + F(1.0)
    x = ((t[0] * ox + t[1] * oy) + t[2]) / k
    y = ((t[3] * ox + t[4] * oy) + t[5]) / k
    xf, yf = np.floor(x), np.floor(y)
    xc, yc = xf + F(1.0), yf + F(1.0)

    def read(yy, xx):
        inside = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1).astype(np.int64), np.clip(xx, 0, W - 1).astype(np.int64)]
        return np.where(inside[..., None], v, F(0.0)).astype(F)

    wx0, wx1 = (xc - x)[..., None], (x - xf)[..., None]
    v_lo = wx0 * read(yf, xf) + wx1 * read(yf, xc)
    v_hi = wx0 * read(yc, xf) + wx1 * read(yc, xc)
    return ((yc - y)[..., None] * v_lo + (y - yf)[..., None] * v_hi).astype(F)


def philox_block(counter, key):
    """Philox4x32-10 (philox_random.h): 10 rounds of (hi/lo products with 0xD2511F53 and 0xCD9E8D57, xor with the
    neighbouring words and the key), the key raised by the Weyl constants between rounds."""
    c = np.array(counter, np.uint64)
    k = np.array(key, np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for r in range(10):
        if r > 0:
            k = (k + np.array([0x9E3779B9, 0xBB67AE85], np.uint64)) & mask
        prod0 = np.uint64(0xD2511F53) * c[0]
        prod1 = np.uint64(0xCD9E8D57) * c[2]
        c = np.array([(prod1 >> np.uint64(32)) ^ c[1] ^ k[0], prod1 & mask, (prod0 >> np.uint64(32)) ^ c[3] ^ k[1],
                      prod0 & mask], np.uint64) & mask
    return [int(v) for v in c]


def stateless_brightness_delta(max_delta, seed=(123, 0)):
    """The delta of ``tf.image.stateless_random_brightness(image, max_delta, seed)`` (preprocessor.py:103): a scalar
    stateless_random_uniform in [-max_delta, max_delta).  GenerateKey scrambles the seed: block(counter = seed words,
    key = {0x3ec8f720, 0x02461e29}) -> key = words 0-1, counter = {0, 0, word 2, word 3}; the sample is word 0 of the
    next block, turned into [0,1) by Uint32ToFloat and scaled in fp32."""
    s0, s1 = int(seed[0]), int(seed[1])
    mix = philox_block([s0 & 0xFFFFFFFF, (s0 >> 32) & 0xFFFFFFFF, s1 & 0xFFFFFFFF, (s1 >> 32) & 0xFFFFFFFF],
                       [0x3EC8F720, 0x02461E29])
    word = philox_block([0, 0, mix[2], mix[3]], [mix[0], mix[1]])[0]
    u = np.array([(127 << 23) | (word & 0x7FFFFF)], np.uint32).view(F)[0] - F(1.0)
    return F(u * (F(max_delta) - F(-max_delta)) + F(-max_delta))


def augment_frame(rgb, zoom, dx_frac, dy_frac, angle, flip, brightness, preprocess=True):
    """One frame [H,W,3] (RGB, 0..255) through the five layers in the reference's order, then (optionally) the input
    scaler.  Draw units as Keras draws them: zoom factor, shifts as fractions of (W, H), angle in radians, flip 0/1;
    ``brightness`` is the delta adjust_brightness adds to the float image (no 0..255 scaling, no clipping)."""
    x = np.asarray(rgb, F)
    H, W = x.shape[:2]
    x = projective_bilinear(x, zoom_transform(zoom, H, W))
    x = projective_bilinear(x, translation_transform(dx_frac, dy_frac, H, W))
    x = projective_bilinear(x, rotation_transform(angle, H, W))
    if flip:
        x = x[:, ::-1]
    x = x + F(brightness)
    if preprocess:                                  # caffe mode: RGB -> BGR, subtract the ImageNet means, no scaling
        x = x[..., ::-1] - BGR_MEAN
    return np.ascontiguousarray(x, F)


def augment_batch(frames, raw, preprocess=True):
    """frames [B,H,W,3]; raw: dict of per-frame draws with keys zoom, dx, dy, angle, flip, brightness."""
    return np.stack([augment_frame(frames[i], raw["zoom"][i], raw["dx"][i], raw["dy"][i], raw["angle"][i],
                                   raw["flip"][i], raw["brightness"][i], preprocess) for i in range(len(frames))]) \
        if len(frames) else np.zeros((0,) + tuple(np.shape(frames)[1:]), F)

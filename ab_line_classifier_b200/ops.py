"""Launch shims: torch tensors (device memory + streams only) -> the C ABI of libabl_b200.so.

No torch math happens here; every function forwards raw device pointers and sizes to a
hand-written sm_100a kernel.  Missing library or non-CUDA tensors raise -- there is no fallback.
"""
import torch

from . import _native
from ._native import EPI_BIAS_RELU, EPI_RELU_MASK  # noqa: F401


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("ab_line_classifier_b200 ops need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("ab_line_classifier_b200 ops need contiguous tensors")
    return t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device.  The raw-handle query costs ~0.3 us; the public
    torch.cuda.current_stream() ~14 us, a third of the host time of a launch."""
    if _raw_stream is not None and _get_device is not None:
        return _raw_stream(_get_device())
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if t is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")


PARTS_HEADER = 64          # ABL_PARTS_HEADER: floats before the first row of a parts buffer
_scratch_cache = {}


def _scratch(device, key, nbytes):
    """Cached zero-initialised fp32 device scratch (grows on demand): parts buffers / workspaces for the convenience
    forms below.  The engine passes its own buffers instead."""
    k = (str(device), key)
    t = _scratch_cache.get(k)
    n = (int(nbytes) + 3) // 4
    if t is None or t.numel() < n:
        t = _scratch_cache[k] = torch.zeros(n, device=device, dtype=torch.float32)
    return t


def parts_bytes(C):
    """Capacity in bytes of a partial-sum buffer for C channels (abl_parts_bytes)."""
    return int(_native.load().abl_parts_bytes(int(C)))


def parts_buffer(C, device):
    return torch.zeros(parts_bytes(C) // 4, device=device, dtype=torch.float32)


def parts_reduce(parts, out, accumulate=False):
    """out[c] (+)= sum of the rows a producer left in `parts`, fixed order (abl_parts_reduce)."""
    _chk(parts, torch.float32, "parts"); _chk(out, torch.float32, "out")
    _native.call("abl_parts_reduce", _ptr(parts), _ptr(out), out.numel(), int(bool(accumulate)), _stream())


def _parts_args(final, parts, C, device, key):
    """(parts tensor or None, bytes) for a producer: the caller's parts buffer, or scratch when only the finished
    [C] vector was asked for (then the caller of this helper reduces into it)."""
    if parts is not None:
        _chk(parts, torch.float32, "parts")
        return parts, parts.numel() * 4
    if final is None:
        return None, 0
    _chk(final, torch.float32, key)
    t = _scratch(device, ("parts", key), parts_bytes(C))
    return t, t.numel() * 4


def conv3x3_igemm(x, w_kmajor, bias=None, mask_src=None, out_full=None, out_pool=None, flags=0, colsum=None,
                  pool_code=None, colsum_parts=None):
    """x bf16 [B,H,W,Cin]; w_kmajor bf16 [9,Cout,Cin]; see include/abl_b200.h abl_conv3x3_igemm.
    colsum_parts: parts buffer receiving the per-CTA column sums; colsum (fp32 [Cout]): the finished sums, OVERWRITTEN
    (this wrapper adds an abl_parts_reduce launch)."""
    _chk(pool_code, torch.int16, "pool_code")
    for t, n in ((x, "x"), (w_kmajor, "w"), (mask_src, "mask_src"), (out_full, "out_full"), (out_pool, "out_pool")):
        _chk(t, torch.bfloat16, n)
    _chk(bias, torch.float32, "bias")
    B, H, W, Cin = x.shape
    Cout = w_kmajor.shape[1]
    assert w_kmajor.shape == (9, Cout, Cin), w_kmajor.shape
    if out_full is not None:
        assert out_full.shape == (B, H, W, Cout)
    if out_pool is not None:
        assert out_pool.shape == (B, H // 2, W // 2, Cout)
    if mask_src is not None:
        assert mask_src.shape == (B, H, W, Cout)
    parts, pbytes = _parts_args(colsum, colsum_parts, Cout, x.device, "colsum")
    _native.call("abl_conv3x3_igemm", _ptr(x), _ptr(w_kmajor), _ptr(bias), _ptr(mask_src), _ptr(out_full),
                 _ptr(out_pool), _ptr(pool_code), _ptr(parts), pbytes, B, H, W, Cin, Cout, flags, _stream())
    if colsum is not None and colsum_parts is None:
        parts_reduce(parts, colsum)


def igemm_gap_supported(H, W, Cout):
    """Shapes abl_conv3x3_igemm_gap handles: any frame size (partial tiles are zero-padded), Cout % 128 == 0."""
    return H > 0 and W > 0 and Cout % 128 == 0


def igemm_gap_parts(H, W, Cout):
    """8 x 16-pixel rectangles per frame in abl_conv3x3_igemm_gap's partial sums (tiles counted rounding up: the
    kernel's tile is 8 rows for Cout % 256 == 0, else 16 rows = two rectangles)."""
    msub = 1 if Cout % 256 == 0 else 2
    return -(-H // (8 * msub)) * msub * -(-W // 16)


def conv3x3_igemm_gap(x, w_kmajor, bias, out_full, gap_part, relu_bits=None):
    """Last conv forward (+bias+ReLU) with the GAP partial sums fused in; out_full may be None.  relu_bits: int32
    [B,H,W,Cout/32] receiving the 1-bit ReLU mask of the activation (what a training step keeps instead of out_full)."""
    _chk(x, torch.bfloat16, "x"); _chk(w_kmajor, torch.bfloat16, "w"); _chk(bias, torch.float32, "bias")
    _chk(out_full, torch.bfloat16, "out_full"); _chk(gap_part, torch.float32, "gap_part")
    _chk(relu_bits, torch.int32, "relu_bits")
    B, H, W, Cin = x.shape
    Cout = w_kmajor.shape[1]
    assert gap_part.numel() >= B * igemm_gap_parts(H, W, Cout) * Cout
    assert relu_bits is None or relu_bits.numel() >= B * H * W * (Cout // 32)
    _native.call("abl_conv3x3_igemm_gap", _ptr(x), _ptr(w_kmajor), _ptr(bias), _ptr(out_full), _ptr(gap_part),
                 _ptr(relu_bits), B, H, W, Cin, Cout, _stream())


def gap_reduce(gap_part, gap, parts, hw):
    _chk(gap_part, torch.float32, "gap_part"); _chk(gap, torch.float32, "gap")
    B, Cc = gap.shape
    _native.call("abl_gap_reduce", _ptr(gap_part), _ptr(gap), B, parts, Cc, hw, _stream())


def dgrad_unpool_supported(H, W, Cout):
    """Shapes abl_conv3x3_dgrad_unpool handles (H, W = pooled resolution, Cout = channels of the pooled tensor)."""
    return H > 0 and W > 0 and Cout % 64 == 0 and Cout <= 128


def conv3x3_dgrad_unpool(dz, w_dgrad, pooled, pool_code, dz_out, db=None, db_parts=None):
    """dgrad + relu'(pooled) + MaxPoolGrad (+ BiasAddGrad) in one kernel; see abl_conv3x3_dgrad_unpool.
    db_parts: parts buffer for the bias partial sums; db: the finished fp32 [Cout] vector (overwritten, extra launch)."""
    for t, n in ((dz, "dz"), (w_dgrad, "w_dgrad"), (pooled, "pooled"), (dz_out, "dz_out")):
        _chk(t, torch.bfloat16, n)
    _chk(pool_code, torch.int16, "pool_code")
    B, H, W, Cin = dz.shape
    Cout = w_dgrad.shape[1]
    assert w_dgrad.shape == (9, Cout, Cin) and pooled.shape == (B, H, W, Cout)
    assert dz_out.shape == (B, 2 * H, 2 * W, Cout) and pool_code.shape == (B, H, W, Cout // 8)
    parts, pbytes = _parts_args(db, db_parts, Cout, dz.device, "db")
    _native.call("abl_conv3x3_dgrad_unpool", _ptr(dz), _ptr(w_dgrad), _ptr(pooled), _ptr(pool_code), _ptr(dz_out),
                 _ptr(parts), pbytes, B, H, W, Cin, Cout, _stream())
    if db is not None and db_parts is None:
        parts_reduce(parts, db)


def wgrad_workspace_bytes(B, H, W, Cin, Cout):
    n = _native.load().abl_conv3x3_wgrad_workspace_bytes(B, H, W, Cin, Cout)
    if n < 0:
        raise RuntimeError(f"abl_conv3x3_wgrad: unsupported shape B={B} H={H} W={W} Cin={Cin} Cout={Cout}")
    return n


def conv3x3_wgrad(x, dz, dw, db, workspace, accumulate=False, db_parts=None):
    """dw fp32 [3,3,Cin,Cout] (+)= x (*) dz; db fp32 [Cout] = sum dz (if given; overwritten).  db_parts: the parts buffer
    the producer of dz filled -- then dz is not read again for db."""
    _chk(x, torch.bfloat16, "x"); _chk(dz, torch.bfloat16, "dz")
    _chk(dw, torch.float32, "dw"); _chk(db, torch.float32, "db"); _chk(db_parts, torch.float32, "db_parts")
    B, H, W, Cin = x.shape
    Cout = dz.shape[3]
    assert dz.shape == (B, H, W, Cout) and dw.numel() == 9 * Cin * Cout
    _native.call("abl_conv3x3_wgrad", _ptr(x), _ptr(dz), _ptr(dw), _ptr(db), _ptr(db_parts), _ptr(workspace),
                 workspace.numel() * workspace.element_size(), int(bool(accumulate)), B, H, W, Cin, Cout, _stream())
    if db is not None and db_parts is None and not wgrad_fuses_bias(Cin):
        _native.launch_count += 1      # stand-alone column-sum pass over dz


def wgrad_fuses_bias(cin):
    """True when abl_conv3x3_wgrad produces db inside its GEMM (the Cin == 64 kernel, unless the generic one is forced)."""
    import os
    return cin == 64 and os.environ.get("ABL_WGRAD_GENERIC", "0") != "1"


def bias_grad(dz, db):
    """db[c] = sum over all leading dims of dz[..., c] (partial rows + fixed-order reduction; overwrites db)."""
    _chk(dz, torch.bfloat16, "dz"); _chk(db, torch.float32, "db")
    C = dz.shape[-1]
    parts, pbytes = _parts_args(db, None, C, dz.device, "bias_grad")
    _native.call("abl_bias_grad", _ptr(dz), _ptr(parts), pbytes, dz.numel() // C, C, _stream())
    parts_reduce(parts, db)


def _first_suffix(x):
    if x.dtype == torch.uint8:
        return "_u8"          # raw RGB frame: vgg16 preprocessing fused into the kernel
    _chk(x, torch.float32, "x")
    return ""


def conv_first_fwd(x, w_hwio, bias, out):
    """x fp32 [B,H,W,3] (preprocessed) or uint8 [B,H,W,3] (raw RGB)."""
    _chk(w_hwio, torch.float32, "w"); _chk(bias, torch.float32, "bias"); _chk(out, torch.bfloat16, "out")
    B, H, W, c = x.shape
    assert c == 3 and w_hwio.numel() == 27 * out.shape[3]
    _native.call("abl_conv_first_fwd" + _first_suffix(x), _ptr(x), _ptr(w_hwio), _ptr(bias), _ptr(out), B, H, W,
                 out.shape[3], _stream())


def conv_first_wgrad_workspace_bytes():
    return int(_native.load().abl_conv_first_wgrad_workspace_bytes())


def conv_first_wgrad(x, dz, dw, db, workspace=None):
    """dw fp32 [3,3,3,64], db fp32 [64] (both overwritten); workspace: >= conv_first_wgrad_workspace_bytes() bytes."""
    _chk(dz, torch.bfloat16, "dz"); _chk(dw, torch.float32, "dw"); _chk(db, torch.float32, "db")
    B, H, W, c = x.shape
    assert c == 3
    if workspace is None:
        workspace = _scratch(x.device, "first_wgrad", conv_first_wgrad_workspace_bytes())
    _native.call("abl_conv_first_wgrad" + _first_suffix(x), _ptr(x), _ptr(dz), _ptr(dw), _ptr(db), _ptr(workspace),
                 workspace.numel() * workspace.element_size(), B, H, W, dz.shape[3], _stream())


def prep_conv_weights(w_hwio, w_fwd=None, w_dgrad=None):
    _chk(w_hwio, torch.float32, "w"); _chk(w_fwd, torch.bfloat16, "w_fwd"); _chk(w_dgrad, torch.bfloat16, "w_dgrad")
    Cin, Cout = w_hwio.shape[-2], w_hwio.shape[-1]
    _native.call("abl_prep_conv_weights", _ptr(w_hwio), _ptr(w_fwd), _ptr(w_dgrad), Cin, Cout, _stream())


def prep_conv_weights_multi(layers):
    """layers: list of (w_hwio fp32 [3,3,Cin,Cout], w_fwd bf16 [9,Cout,Cin], w_dgrad bf16 [9,Cin,Cout]); one launch."""
    import ctypes as C
    n = len(layers)
    w = (C.c_void_p * n)(*[_ptr(l[0]) for l in layers])
    wf = (C.c_void_p * n)(*[_ptr(l[1]) for l in layers])
    wd = (C.c_void_p * n)(*[_ptr(l[2]) for l in layers])
    cin = (C.c_int * n)(*[int(l[0].shape[2]) for l in layers])
    cout = (C.c_int * n)(*[int(l[0].shape[3]) for l in layers])
    for l in layers:
        _chk(l[0], torch.float32, "w_hwio"); _chk(l[1], torch.bfloat16, "w_fwd"); _chk(l[2], torch.bfloat16, "w_dgrad")
    _native.call("abl_prep_conv_weights_multi", C.cast(w, C.c_void_p), C.cast(wf, C.c_void_p), C.cast(wd, C.c_void_p),
                 C.cast(cin, C.c_void_p), C.cast(cout, C.c_void_p), n, _stream())


def maxpool_bwd_relu(dp, a, pooled, dz, db=None, db_parts=None):
    """db_parts: parts buffer for the fused bias partial sums; db: finished fp32 [C] (overwritten, extra launch)."""
    for t in (dp, a, pooled, dz):
        _chk(t, torch.bfloat16, "tensor")
    B, H, W, Cc = a.shape
    parts, pbytes = _parts_args(db, db_parts, Cc, a.device, "db")
    _native.call("abl_maxpool_bwd_relu", _ptr(dp), _ptr(a), _ptr(pooled), _ptr(dz), _ptr(parts), pbytes, B, H, W, Cc,
                 _stream())
    if db is not None and db_parts is None:
        parts_reduce(parts, db)


def maxpool_bwd_code(dp, pool_code, dz, db=None, db_parts=None):
    """dz from the forward's 2-bit routing codes; dp must already carry relu'(pooled) (abl_maxpool_bwd_code)."""
    _chk(dp, torch.bfloat16, "dp"); _chk(dz, torch.bfloat16, "dz"); _chk(pool_code, torch.int16, "pool_code")
    B, H, W, Cc = dz.shape
    parts, pbytes = _parts_args(db, db_parts, Cc, dz.device, "db")
    _native.call("abl_maxpool_bwd_code", _ptr(dp), _ptr(pool_code), _ptr(dz), _ptr(parts), pbytes, B, H, W, Cc, _stream())
    if db is not None and db_parts is None:
        parts_reduce(parts, db)


def gap_fwd(a, gap):
    _chk(a, torch.bfloat16, "a"); _chk(gap, torch.float32, "gap")
    B, H, W, Cc = a.shape
    _native.call("abl_gap_fwd", _ptr(a), _ptr(gap), B, H * W, Cc, _stream())


def gap_bwd_relu(dgap, a, dz, db=None, db_parts=None):
    _chk(dgap, torch.float32, "dgap"); _chk(a, torch.bfloat16, "a"); _chk(dz, torch.bfloat16, "dz")
    B, H, W, Cc = a.shape
    parts, pbytes = _parts_args(db, db_parts, Cc, a.device, "db")
    _native.call("abl_gap_bwd_relu", _ptr(dgap), _ptr(a), _ptr(dz), _ptr(parts), pbytes, B, H * W, Cc, _stream())
    if db is not None and db_parts is None:
        parts_reduce(parts, db)


def gap_bwd_bits(dgap, relu_bits, dz, db=None, db_parts=None):
    """gap_bwd_relu from the 1-bit ReLU mask conv3x3_igemm_gap wrote (relu_bits int32 [B,H,W,C/32])."""
    _chk(dgap, torch.float32, "dgap"); _chk(relu_bits, torch.int32, "relu_bits"); _chk(dz, torch.bfloat16, "dz")
    B, H, W, Cc = dz.shape
    assert relu_bits.numel() >= B * H * W * (Cc // 32)
    parts, pbytes = _parts_args(db, db_parts, Cc, dz.device, "db")
    _native.call("abl_gap_bwd_bits", _ptr(dgap), _ptr(relu_bits), _ptr(dz), _ptr(parts), pbytes, B, H * W, Cc, _stream())
    if db is not None and db_parts is None:
        parts_reduce(parts, db)


def head_loss_ws(B, device):
    """Zeroed workspace for the deterministic batch-loss sum of head_fwd (abl_head_loss_ws_bytes)."""
    return torch.zeros(int(_native.load().abl_head_loss_ws_bytes(int(B))) // 4, device=device, dtype=torch.float32)


def head_fwd(gap, w, b, logits, probs, y=None, class_weight=None, inv_batch=1.0, loss_sum=None, dlogits=None,
             keep_mask=None, keep_mask_in=False, training=False, dropout=0.0, seed=0, step=0, rank=0, step_dev=None,
             loss_ws=None, loss_accumulate=True, gap_part=None, gap_parts=0, gap_hw=0):
    """loss_sum (+)= batch loss (loss_accumulate False: overwritten).  step_dev: int32[2] device counters {dropout step,
    optimizer step}, both advanced by the kernel (CUDA-graph replays).  gap_part (+ gap_parts rectangles per frame of
    gap_hw pixels in all): conv3x3_igemm_gap's partial sums -- the kernel finishes the pooling itself and WRITES gap."""
    for t in (gap, w, b, logits, probs, y, class_weight, loss_sum, dlogits, loss_ws, gap_part):
        _chk(t, torch.float32, "tensor")
    _chk(keep_mask, torch.uint8, "keep_mask")
    B, Cc = gap.shape
    n = w.shape[1]
    if loss_ws is None and loss_sum is not None and y is not None:
        loss_ws = _scratch(gap.device, "loss_ws", int(_native.load().abl_head_loss_ws_bytes(B)))
    if step_dev is not None:       # dropout counter in device memory (CUDA-graph replays)
        _chk(step_dev, torch.int32, "step_dev")
        _native.call("abl_head_fwd_stepdev", _ptr(gap), _ptr(w), _ptr(b), _ptr(logits), _ptr(probs), _ptr(y),
                     _ptr(class_weight), float(inv_batch), _ptr(loss_sum), _ptr(dlogits), _ptr(keep_mask),
                     int(bool(training)), float(dropout), int(seed), _ptr(step_dev), int(rank), B, Cc, n,
                     int(bool(loss_accumulate)), _ptr(loss_ws), _ptr(gap_part), int(gap_parts), int(gap_hw), _stream())
        return
    _native.call("abl_head_fwd", _ptr(gap), _ptr(w), _ptr(b), _ptr(logits), _ptr(probs), _ptr(y),
                 _ptr(class_weight), float(inv_batch), _ptr(loss_sum), _ptr(dlogits), _ptr(keep_mask),
                 int(bool(keep_mask_in)), int(bool(training)), float(dropout), int(seed), int(step), int(rank),
                 B, Cc, n, int(bool(loss_accumulate)), _ptr(loss_ws), _ptr(gap_part), int(gap_parts), int(gap_hw),
                 _stream())


def head_bwd(gap, w, dlogits, keep_mask, dropout, dw, db, dgap):
    for t in (gap, w, dlogits, dw, db, dgap):
        _chk(t, torch.float32, "tensor")
    B, Cc = gap.shape
    _native.call("abl_head_bwd", _ptr(gap), _ptr(w), _ptr(dlogits), _ptr(keep_mask), float(dropout), _ptr(dw),
                 _ptr(db), _ptr(dgap), B, Cc, w.shape[1], _stream())


def adam_flat(theta, g, m, v, alpha_t, b1=0.9, b2=0.999, eps=1e-7):
    for t in (theta, g, m, v):
        _chk(t, torch.float32, "tensor")
    _native.call("abl_adam_flat", _ptr(theta), _ptr(g), _ptr(m), _ptr(v), theta.numel(), float(alpha_t), float(b1),
                 float(b2), float(eps), _stream())


def adam_flat_tab(theta, g, m, v, alpha_tab, t_dev, b1=0.9, b2=0.999, eps=1e-7):
    """Adam with alpha_t looked up on the device: alpha_tab fp32 [T] (t = 1..T), t_dev int32 [1] (abl_adam_flat_tab)."""
    for t in (theta, g, m, v, alpha_tab):
        _chk(t, torch.float32, "tensor")
    _chk(t_dev, torch.int32, "t_dev")
    _native.call("abl_adam_flat_tab", _ptr(theta), _ptr(g), _ptr(m), _ptr(v), theta.numel(), _ptr(alpha_tab), _ptr(t_dev),
                 alpha_tab.numel(), float(b1), float(b2), float(eps), _stream())


def rmsprop_flat(theta, g, rms, lr, rho=0.9, eps=1e-7):
    for t in (theta, g, rms):
        _chk(t, torch.float32, "tensor")
    _native.call("abl_rmsprop_flat", _ptr(theta), _ptr(g), _ptr(rms), theta.numel(), float(lr), float(rho),
                 float(eps), _stream())


def resize_nearest_u8(src, dst):
    """src uint8 [B,h,w,3] -> dst uint8 [B,H,W,3], tf.image.resize NEAREST_NEIGHBOR semantics (abl_resize_nearest_u8)."""
    _chk(src, torch.uint8, "src"); _chk(dst, torch.uint8, "dst")
    B, h, w, _ = src.shape
    assert dst.shape[0] == B and dst.shape[3] == 3 and src.shape[3] == 3
    _native.call("abl_resize_nearest_u8", _ptr(src), _ptr(dst), B, h, w, dst.shape[1], dst.shape[2], _stream())


def augment_workspace_bytes(B, H, W):
    return int(_native.load().abl_augment_workspace_bytes(int(B), int(H), int(W)))


def augment_frames(src, params, dst, preprocess=True, workspace=None):
    """src uint8|fp32 [B,H,W,3] RGB in [0,255], params fp32 [B,20] -> dst fp32 [B,H,W,3]: the reference's augmentation
    Sequential + vgg16.preprocess_input in one launch (abl_augment_frames)."""
    if src.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"src: expected uint8 or float32, got {src.dtype}")
    _chk(params, torch.float32, "params"); _chk(dst, torch.float32, "dst")
    B, H, W, c = src.shape
    assert c == 3 and tuple(dst.shape) == (B, H, W, 3) and tuple(params.shape) == (B, 20), (src.shape, dst.shape, params.shape)
    if workspace is not None:
        _chk(workspace, torch.uint8, "workspace")
    _native.call("abl_augment_frames", _ptr(src), int(src.dtype == torch.uint8), _ptr(params), _ptr(dst),
                 _ptr(workspace), 0 if workspace is None else workspace.numel(), B, H, W, int(bool(preprocess)), _stream())


def metrics_counter_count(n_classes, num_thresholds):
    return int(_native.load().abl_metrics_counter_count(int(n_classes), int(num_thresholds)))


def metrics_accumulate(probs, y_true, auc_thresholds, class_thresholds, counters):
    """Add one batch (probs, one-hot y: fp32 [B,n]) to the device-side Keras metric counters (abl_metrics_accumulate)."""
    for t, n in ((probs, "probs"), (y_true, "y_true"), (auc_thresholds, "auc_thresholds"), (class_thresholds, "class_thresholds")):
        _chk(t, torch.float32, n)
    _chk(counters, torch.int64, "counters")
    B, n = probs.shape
    assert tuple(y_true.shape) == (B, n) and class_thresholds.numel() == n
    assert counters.numel() >= metrics_counter_count(n, auc_thresholds.numel())
    _native.call("abl_metrics_accumulate", _ptr(probs), _ptr(y_true), B, n, _ptr(auc_thresholds),
                 auc_thresholds.numel(), _ptr(class_thresholds), _ptr(counters), _stream())


def clip_aggregate(probs, clip_offsets, b_idx, threshold, window):
    """probs fp32 [N,n] (device), clip_offsets int64 [n_clips+1] (device) -> (mean [n_clips,n], run int32 [n_clips],
    win fp32 [n_clips]); see abl_clip_aggregate."""
    _chk(probs, torch.float32, "probs"); _chk(clip_offsets, torch.int64, "clip_offsets")
    n_clips, n = clip_offsets.numel() - 1, probs.shape[1]
    mean = torch.empty(n_clips, n, device=probs.device, dtype=torch.float32)
    run = torch.empty(n_clips, device=probs.device, dtype=torch.int32)
    win = torch.empty(n_clips, device=probs.device, dtype=torch.float32)
    _native.call("abl_clip_aggregate", _ptr(probs), _ptr(clip_offsets), n_clips, n, int(b_idx), float(threshold),
                 int(window), _ptr(mean), _ptr(run), _ptr(win), _stream())
    return mean, run, win


def opt_flat_dp(kind, theta_ptrs, grad_ptrs, m_or_rms, v, offset, n, rank, world, lr_or_alpha, b1_or_rho, b2=0.0,
                eps=1e-7, alpha_tab=None, t_dev=None):
    """Fused reduce-scatter(grad) -> Adam/RMSprop(shard) -> all-gather(theta) over peer pointers (abl_opt_flat_dp).
    theta_ptrs / grad_ptrs: lists of `world` device addresses (int) of every rank's flat buffers in this process."""
    import ctypes as C
    tp = (C.c_void_p * world)(*theta_ptrs)
    gp = (C.c_void_p * world)(*grad_ptrs)
    _chk(m_or_rms, torch.float32, "m_or_rms"); _chk(v, torch.float32, "v")
    _chk(alpha_tab, torch.float32, "alpha_tab"); _chk(t_dev, torch.int32, "t_dev")
    _native.call("abl_opt_flat_dp", {"adam": 0, "rmsprop": 1}[kind], C.cast(tp, C.c_void_p), C.cast(gp, C.c_void_p),
                 _ptr(m_or_rms), _ptr(v), int(offset), int(n), int(rank), int(world), float(lr_or_alpha),
                 float(b1_or_rho), float(b2), float(eps), _ptr(alpha_tab), _ptr(t_dev),
                 0 if alpha_tab is None else alpha_tab.numel(), _stream())


def gradcam(a, probs, w, heat):
    _chk(a, torch.bfloat16, "a"); _chk(probs, torch.float32, "probs"); _chk(w, torch.float32, "w")
    _chk(heat, torch.float32, "heat")
    B, H, W, Cc = a.shape
    _native.call("abl_gradcam", _ptr(a), _ptr(probs), _ptr(w), _ptr(heat), B, H, W, Cc, w.shape[1],
                 heat.shape[1], heat.shape[2], _stream())

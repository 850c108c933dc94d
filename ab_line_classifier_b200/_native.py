"""ctypes binding of libabl_b200.so (the C ABI declared in include/abl_b200.h).

This is the reference-side FFI shim: plain pointers and sizes, no torch types.  There is NO
fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libabl_b200.so")

EPI_BIAS_RELU = 1
EPI_RELU_MASK = 2

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_u, _ull = C.c_uint, C.c_ulonglong

# name -> (restype, argtypes); must list EVERY symbol of include/abl_b200.h (tests check this).
SIGNATURES = {
    "abl_abi_version": (_i, []),
    "abl_last_error": (C.c_char_p, []),
    "abl_parts_bytes": (_ll, [_i]),
    "abl_parts_reduce": (_i, [_vp, _vp, _i, _i, _vp]),
    "abl_conv3x3_igemm": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp]),
    "abl_conv3x3_igemm_gap": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "abl_gap_reduce": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "abl_conv3x3_dgrad_unpool": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _vp]),
    "abl_conv3x3_wgrad_workspace_bytes": (_ll, [_i, _i, _i, _i, _i]),
    "abl_conv3x3_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp]),
    "abl_bias_grad": (_i, [_vp, _vp, _ll, _ll, _i, _vp]),
    "abl_conv_first_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "abl_conv_first_wgrad_workspace_bytes": (_ll, []),
    "abl_conv_first_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "abl_conv_first_fwd_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "abl_conv_first_wgrad_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "abl_prep_conv_weights": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "abl_prep_conv_weights_multi": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "abl_maxpool_bwd_relu": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "abl_maxpool_bwd_code": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "abl_gap_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "abl_gap_bwd_relu": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "abl_gap_bwd_bits": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "abl_head_loss_ws_bytes": (_ll, [_i]),
    "abl_head_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _f, _ull, _u, _u,
                          _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "abl_head_fwd_stepdev": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _f, _ull, _vp, _u, _i, _i, _i,
                                  _i, _vp, _vp, _i, _i, _vp]),
    "abl_head_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "abl_adam_flat": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _vp]),
    "abl_adam_flat_tab": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _f, _f, _f, _vp]),
    "abl_rmsprop_flat": (_i, [_vp, _vp, _vp, _ll, _f, _f, _f, _vp]),
    "abl_resize_nearest_u8": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "abl_clip_aggregate": (_i, [_vp, _vp, _i, _i, _i, _f, _i, _vp, _vp, _vp, _vp]),
    "abl_opt_flat_dp": (_i, [_i, _vp, _vp, _vp, _vp, _ll, _ll, _i, _i, _f, _f, _f, _f, _vp, _vp, _i, _vp]),
    "abl_augment_workspace_bytes": (_ll, [_i, _i, _i]),
    "abl_augment_frames": (_i, [_vp, _i, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "abl_metrics_counter_count": (_i, [_i, _i]),
    "abl_metrics_accumulate": (_i, [_vp, _vp, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "abl_dev_mma_rate_probe": (_i, [_i, _i, _i, _i, _vp, _vp]),
    "abl_dev_write_probe": (_i, [_vp, _ll, _i, _i, _vp]),
    "abl_gradcam": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
}

_lib = None

# kernels launched per entry point.  block1_conv1's wgrad = GEMM + fixed-order reduction of the per-CTA sums; the 3x3
# wgrad GEMM reduces its pixel slices itself unless ABL_WGRAD_FUSED_REDUCE=0 brings the separate reduction kernel back.
LAUNCHES_PER_CALL = {"abl_conv_first_wgrad": 2, "abl_conv_first_wgrad_u8": 2}


def _wgrad_launches(args):
    """abl_conv3x3_wgrad(x, dz, dw, db, db_parts, ws, ws_bytes, accumulate, B, H, W, Cin, Cout, stream): the Cin >= 128
    GEMM reduces its pixel slices itself, the Cin = 64 one is followed by wgrad_reduce_kernel (see wgrad_tc.cu)."""
    mode = os.environ.get("ABL_WGRAD_FUSED_REDUCE", "1")
    fused = mode == "2" or (mode != "0" and int(args[11]) != 64)
    return 1 if fused else 2
launch_count = 0          # kernels launched through this module (bench.py reports it as gpu_launches)
timing = None             # set to a list to record (name, args, start_event, end_event) per call


def load():
    """dlopen the in-tree library and attach signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m ab_line_classifier_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.abl_abi_version() != 8:
        raise RuntimeError("libabl_b200.so ABI version mismatch")
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point; non-zero -> RuntimeError with abl_last_error()."""
    global launch_count
    lib = load()
    if timing is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        timing.append((name, args, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    launch_count += _wgrad_launches(args) if name == "abl_conv3x3_wgrad" else LAUNCHES_PER_CALL.get(name, 1)
    if rc != 0:
        msg = lib.abl_last_error().decode(errors="replace")
        raise RuntimeError(f"{name} failed (code {rc}): {msg}")
    return rc

"""-m gpu: every sm_100a kernel, called through the C ABI (ctypes), against the CPU oracle.

Tolerances: conv operands are bf16 with fp32 accumulation and outputs are stored in bf16, so a conv
output may differ from the fp32-accumulated oracle (same bf16-rounded operands) by one bf16 ulp of the
largest value: 2^-8 relative.  fp32 kernels (head, optimizers, GAP, Grad-CAM) are held to 1e-5..1e-6.
Integer / index work (pooling routes, ReLU masks, dropout masks) must be bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle

pytestmark = pytest.mark.gpu
BF16_ULP = 2.0 ** -8


def _ops():
    from ab_line_classifier_b200 import ops
    return ops


def _bf(t):
    return t.to(torch.bfloat16).float()


def _conv_cpu(x_bf, w, b=None, relu=True):
    y = F.conv2d(x_bf.permute(0, 3, 1, 2), _bf(w).permute(3, 2, 0, 1), b, padding=1)
    return (F.relu(y) if relu else y).permute(0, 2, 3, 1).contiguous()


def _close(got, ref, rel, what):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, f"{what}: max|err| {err:.4g} > {rel:.3g} * {scale:.4g}"


@pytest.fixture(params=["halo", "v1"])
def igemm_variant(request, monkeypatch):
    """Both main loops of abl_conv3x3_igemm: the haloed-tile kernel (default) and the v1 kernel that
    remains the fallback for frame sizes the 16-pixel-wide tiles do not divide."""
    monkeypatch.setenv("ABL_IGEMM_V1", "1" if request.param == "v1" else "0")
    return request.param


LAYERS = [(64, 64, 128, 128, True), (64, 128, 64, 64, False), (128, 128, 64, 64, True),
          (128, 256, 32, 32, False), (256, 256, 32, 32, False),
          # frames that are not a whole number of 8 x 16-pixel tiles (IMG_DIM 96, 100, 160, 224 ... give such block-2 / 3
          # maps): partial tiles, zero-filled by TMA on load, clipped on store, guarded in the epilogue
          (64, 64, 24, 40, True), (64, 128, 20, 20, False), (128, 128, 12, 24, True), (128, 256, 25, 25, False),
          (256, 256, 24, 24, False), (256, 512, 10, 6, False), (128, 128, 56, 56, True),
          # odd sides under a pool: Keras' 'valid' MaxPooling2D drops the last row / column
          (64, 64, 25, 41, True), (128, 128, 45, 45, True)]


@pytest.mark.parametrize("cin,cout,H,W,pool", LAYERS)
@pytest.mark.parametrize("B", [1, 3])
def test_conv3x3_forward_bias_relu_pool(cin, cout, H, W, pool, B, igemm_variant):
    ops = _ops()
    g = torch.Generator().manual_seed(cin * 7 + cout + B)
    x = _bf(torch.randn(B, H, W, cin, generator=g) * 0.5)
    w = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    wf = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), wf, None)
    assert torch.equal(wf.cpu(), w.permute(1, 0, 3, 2).reshape(9, cout, cin).to(torch.bfloat16))   # bit-exact re-layout
    full = torch.full((B, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    pl = torch.full((B, H // 2, W // 2, cout), float("nan"), device="cuda", dtype=torch.bfloat16) if pool else None
    ops.conv3x3_igemm(x.cuda().to(torch.bfloat16), wf, b.cuda(), None, full, pl, ops.EPI_BIAS_RELU)
    _close(full, _conv_cpu(x, w, b), 1.5 * BF16_ULP, "conv fwd")
    if pool:   # pooling of the kernel's own bf16 output is exact
        ref = F.max_pool2d(full.float().cpu().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        assert torch.equal(pl.float().cpu(), ref)


@pytest.mark.parametrize("B,H,W", [(3, 128, 128), (2, 64, 64), (1, 8, 128), (12, 128, 128)])
def test_blocked_cin64_cout64_pooled_forward_matches_reference(B, H, W, monkeypatch):
    """The pooled-only forward of a Cin = Cout = 64 layer (block1_conv2) runs the blocked main loop by default
    (conv3x3_blk64_pool_kernel: four adjacent pixels = one GEMM row, horizontal half of the max-pool in registers).
    Its pooled output and routing codes must equal those of the haloed kernel (ABL_BLOCKED64=0) -- the same bf16
    values up to the accumulation order -- and reproduce MaxPoolGrad of a full-resolution run exactly."""
    ops = _ops()
    g = torch.Generator().manual_seed(B + H)
    x = _bf(torch.randn(B, H, W, 64, generator=g) * 0.5)
    w = torch.randn(3, 3, 64, 64, generator=g) * (2.0 / 576) ** 0.5
    b = torch.randn(64, generator=g) * 0.1
    wf = torch.empty(9, 64, 64, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), wf, None)
    xg = x.cuda().to(torch.bfloat16)
    pl = torch.full((B, H // 2, W // 2, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    code = torch.full((B, H // 2, W // 2, 8), -1, device="cuda", dtype=torch.int16)
    ops.conv3x3_igemm(xg, wf, b.cuda(), None, None, pl, ops.EPI_BIAS_RELU, pool_code=code)       # blocked (default)
    pl_nc = torch.empty_like(pl)
    ops.conv3x3_igemm(xg, wf, b.cuda(), None, None, pl_nc, ops.EPI_BIAS_RELU)                     # without codes
    assert torch.equal(pl, pl_nc)
    if B <= 3:
        ref = F.max_pool2d(_conv_cpu(x, w, b).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
        _close(pl, ref, 1.5 * BF16_ULP, "blocked conv fwd + pool")
    monkeypatch.setenv("ABL_BLOCKED64", "0")
    full = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    pl_h = torch.empty_like(pl)
    code_h = torch.empty_like(code)
    ops.conv3x3_igemm(xg, wf, b.cuda(), None, full, pl_h, ops.EPI_BIAS_RELU, pool_code=code_h)    # haloed kernel
    _close(pl, pl_h, 1.5 * BF16_ULP, "blocked vs haloed pooled output")
    # where both kernels produced the same bf16 values in a window, they must route to the same element
    same = (pl == pl_h).all(dim=-1)
    assert float(same.float().mean()) > 0.5
    # the blocked kernel's codes name an element that holds the window maximum of ITS accumulation: routing a
    # gradient by them puts every value on an element whose haloed-kernel activation is within one ulp of the maximum
    dp = (torch.randn(B, H // 2, W // 2, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()
    dpm = torch.where(pl > 0, dp, torch.zeros_like(dp))
    dz = torch.empty_like(full)
    ops.maxpool_bwd_code(dpm, code, dz)
    routed = dz != 0
    up = pl_h.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    assert bool((full[routed].float() >= up[routed].float() * (1 - 2 * BF16_ULP) - 1e-6).all())
    pooled_back = F.max_pool2d(dz.float().abs().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(pooled_back, dpm.float().abs())          # exactly one element per window carries the gradient
    # dgrad + ReLU mask + fused BiasAddGrad of the same layer shape stays on the haloed kernel
    wd = torch.empty(9, 64, 64, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), None, wd)
    if B <= 3:
        dzz = _bf(torch.randn(B, H, W, 64, generator=g) * 0.1)
        act = _bf(torch.relu(torch.randn(B, H, W, 64, generator=g)))
        out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        colsum = torch.full((64,), -1.0, device="cuda")
        ops.conv3x3_igemm(dzz.cuda().to(torch.bfloat16), wd, None, act.cuda().to(torch.bfloat16), out, None,
                          ops.EPI_RELU_MASK, colsum=colsum)
        dx = F.conv_transpose2d(dzz.permute(0, 3, 1, 2), _bf(w).permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
        _close(out, dx * (act > 0), 1.5 * BF16_ULP, "dgrad (Cin = Cout = 64)")
        _close(colsum, out.float().sum(dim=(0, 1, 2)), 2e-5, "bias grad (separate pass, overwritten)")
        assert (out.float().cpu()[act == 0] == 0).all()


def test_prep_conv_weights_multi_equals_per_layer_launches():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    layers, refs = [], []
    for cin, cout in [(64, 64), (64, 128), (256, 256)]:
        w = torch.randn(3, 3, cin, cout, generator=g).cuda()
        wf = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
        wd = torch.empty(9, cin, cout, device="cuda", dtype=torch.bfloat16)
        rf, rd = torch.empty_like(wf), torch.empty_like(wd)
        ops.prep_conv_weights(w, rf, rd)
        layers.append((w, wf, wd))
        refs.append((rf, rd))
    ops.prep_conv_weights_multi(layers)
    for (w, wf, wd), (rf, rd) in zip(layers, refs):
        assert torch.equal(wf, rf) and torch.equal(wd, rd)


def test_conv3x3_pool_only_output_matches_full_path(igemm_variant):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 64, 64, 128, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(3, 3, 128, 128, generator=g) * 0.04).cuda()
    b = torch.zeros(128, device="cuda")
    wf = torch.empty(9, 128, 128, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w, wf, None)
    full = torch.empty(2, 64, 64, 128, device="cuda", dtype=torch.bfloat16)
    p1 = torch.empty(2, 32, 32, 128, device="cuda", dtype=torch.bfloat16)
    p2 = torch.empty_like(p1)
    ops.conv3x3_igemm(x, wf, b, None, full, p1, ops.EPI_BIAS_RELU)
    ops.conv3x3_igemm(x, wf, b, None, None, p2, ops.EPI_BIAS_RELU)     # inference mode: pooled only
    assert torch.equal(p1, p2)


@pytest.mark.parametrize("cin,cout,H,W,pool", LAYERS)
def test_conv3x3_dgrad_with_relu_mask(cin, cout, H, W, pool, igemm_variant):
    ops = _ops()
    g = torch.Generator().manual_seed(cin + cout)
    B = 2
    dz = _bf(torch.randn(B, H, W, cout, generator=g) * 0.1)
    w = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    act = _bf(torch.relu(torch.randn(B, H, W, cin, generator=g)))
    wd = torch.empty(9, cin, cout, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), None, wd)
    out = torch.full((B, H, W, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    colsum = torch.full((cin,), -1.0, device="cuda")     # fused BiasAddGrad of the layer below; OVERWRITTEN
    ops.conv3x3_igemm(dz.cuda().to(torch.bfloat16), wd, None, act.cuda().to(torch.bfloat16), out, None,
                      ops.EPI_RELU_MASK, colsum=colsum)
    dx = F.conv_transpose2d(dz.permute(0, 3, 1, 2), _bf(w).permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    _close(out, dx * (act > 0), 1.5 * BF16_ULP, "dgrad")
    # the column sums are those of the bf16 tensor the kernel wrote (fp32 accumulation, any order)
    _close(colsum, out.float().sum(dim=(0, 1, 2)), 2e-5, "fused bias grad (dgrad epilogue)")
    assert (out.float().cpu()[act == 0] == 0).all()          # ReLU mask is exact
    # per-CTA partial sums + fixed-order reduction: a second launch reproduces the column sums bit for bit
    colsum2 = torch.empty_like(colsum)
    ops.conv3x3_igemm(dz.cuda().to(torch.bfloat16), wd, None, act.cuda().to(torch.bfloat16), out, None,
                      ops.EPI_RELU_MASK, colsum=colsum2)
    assert torch.equal(colsum, colsum2)


@pytest.mark.parametrize("cin,cout,H,W,pool", LAYERS)
@pytest.mark.parametrize("B", [1, 5])
@pytest.mark.parametrize("reduce", ["in-kernel", "separate-kernel"])
def test_conv3x3_wgrad_and_bias_grad(cin, cout, H, W, pool, B, reduce, monkeypatch):
    """reduce: the fixed-order sum of the pixel slices inside the GEMM kernel (default: arrive / wait tickets, every CTA
    reduces its share of the unit) or by wgrad_reduce_kernel (ABL_WGRAD_FUSED_REDUCE=0)."""
    monkeypatch.setenv("ABL_WGRAD_FUSED_REDUCE", "2" if reduce == "in-kernel" else "0")      # 2: also for Cin = 64
    ops = _ops()
    g = torch.Generator().manual_seed(cin * 3 + cout + B)
    x = _bf(torch.relu(torch.randn(B, H, W, cin, generator=g)))
    dz = _bf(torch.randn(B, H, W, cout, generator=g) * 0.1)
    ws = torch.empty(ops.wgrad_workspace_bytes(B, H, W, cin, cout), device="cuda", dtype=torch.uint8)
    dw = torch.full((3, 3, cin, cout), float("nan"), device="cuda")
    db = torch.full((cout,), float("nan"), device="cuda")          # overwritten
    xg, dzg = x.cuda().to(torch.bfloat16), dz.cuda().to(torch.bfloat16)
    ops.conv3x3_wgrad(xg, dzg, dw, db, ws, accumulate=False)
    # the producer-side path: partial rows left by a stand-alone column-sum pass, finished by the wgrad reduction
    parts = ops.parts_buffer(cout, "cuda")
    ops._native.call("abl_bias_grad", dzg.data_ptr(), parts.data_ptr(), parts.numel() * 4, dzg.numel() // cout, cout,
                     ops._stream())
    db_p = torch.full((cout,), float("nan"), device="cuda")
    ops.conv3x3_wgrad(xg, dzg, torch.empty_like(dw), db_p, ws, accumulate=False, db_parts=parts)
    _close(db_p, dz.sum(dim=(0, 1, 2)), 1e-5, "bias grad from producer parts")
    ref = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (cout, cin, 3, 3), dz.permute(0, 3, 1, 2),
                                      padding=1).permute(2, 3, 1, 0)
    _close(dw, ref, 2e-4, "wgrad")                      # fp32 accumulation, different summation order
    _close(db, dz.sum(dim=(0, 1, 2)), 1e-5, "bias grad")
    dw2 = dw.clone()
    ops.conv3x3_wgrad(xg, dzg, dw2, None, ws, accumulate=True)
    _close(dw2, 2 * ref, 2e-4, "wgrad accumulate")
    dw3 = torch.empty_like(dw)
    ops.conv3x3_wgrad(xg, dzg, dw3, None, ws, accumulate=False)
    assert torch.equal(dw, dw3), "wgrad must be deterministic (fixed-order slice reduction)"
    for _ in range(3):          # back-to-back launches: the tickets re-arm themselves
        ops.conv3x3_wgrad(xg, dzg, dw3, db_p, ws, accumulate=False, db_parts=parts)
    assert torch.equal(dw, dw3)
    _close(db_p, dz.sum(dim=(0, 1, 2)), 1e-5, "bias grad from producer parts (repeat)")


@pytest.mark.parametrize("B,H,W", [(3, 128, 128), (2, 64, 64), (1, 32, 256), (4, 16, 32), (2, 48, 96),
                                   (3, 100, 100), (1, 25, 25), (2, 224, 224)])     # the last tile may be short
@pytest.mark.parametrize("gather", ["patch", "global", "patch-smem-a", "global-smem-a"])
def test_conv_first_forward_and_wgrad_fp32(B, H, W, gather, monkeypatch):
    """Both im2col sources: the shared-memory patch prefetched one tile ahead (tiles that are rectangles of one frame:
    W a power of two <= 128 or a multiple of 128) and the per-pixel global gather (any other shape, here 48x96); and both
    homes of the forward's A operand: tensor memory (default) and the swizzled shared-memory tile (ABL_FIRST_SMEM_A=1)."""
    if gather.startswith("global"):
        monkeypatch.setenv("ABL_FIRST_GATHER", "1")
    else:
        monkeypatch.delenv("ABL_FIRST_GATHER", raising=False)
    monkeypatch.setenv("ABL_FIRST_SMEM_A", "1" if gather.endswith("smem-a") else "0")
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(B, H, W, 3, generator=g) * 255 - 115).contiguous()
    w = torch.randn(3, 3, 3, 64, generator=g) * 0.1
    b = torch.randn(64, generator=g) * 0.1
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    ops.conv_first_fwd(x.cuda(), w.cuda(), b.cuda(), out)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1)).permute(0, 2, 3, 1)
    _close(out, ref, BF16_ULP, "conv_first fwd")        # fp32 math, bf16 store
    dz = _bf(torch.randn(B, H, W, 64, generator=g) * 0.1)
    dw, db = torch.full((3, 3, 3, 64), float("nan"), device="cuda"), torch.full((64,), float("nan"), device="cuda")
    ops.conv_first_wgrad(x.cuda(), dz.cuda().to(torch.bfloat16), dw, db)          # both overwritten
    dw_b, db_b = torch.empty_like(dw), torch.empty_like(db)
    ops.conv_first_wgrad(x.cuda(), dz.cuda().to(torch.bfloat16), dw_b, db_b)
    assert torch.equal(dw, dw_b) and torch.equal(db, db_b)                        # fixed-order sums: bit-reproducible
    refw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (64, 3, 3, 3), dz.permute(0, 3, 1, 2),
                                       padding=1).permute(2, 3, 1, 0)
    _close(dw, refw, 1e-2, "conv_first wgrad")
    _close(db, dz.sum(dim=(0, 1, 2)), 1e-4, "conv_first bias grad")


def test_conv_first_uint8_path_fuses_vgg16_preprocessing_bit_exactly():
    """f1: raw uint8 RGB frames in, caffe-mode preprocess_input (preprocessor.py:60-63) inside the kernel."""
    ops = _ops()
    from oracle import synth
    frames, _ = synth.synth_frames(3, seed=9)
    frames[0, :4, :4] = 255                  # exercise the borders / saturation
    x = synth.vgg16_preprocess(frames)
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(3, 3, 3, 64, generator=g) * 0.1).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    o32 = torch.empty(3, 128, 128, 64, device="cuda", dtype=torch.bfloat16)
    o8 = torch.empty_like(o32)
    ops.conv_first_fwd(torch.from_numpy(x).cuda(), w, b, o32)
    ops.conv_first_fwd(torch.from_numpy(frames).cuda(), w, b, o8)
    assert torch.equal(o32, o8)
    dz = (torch.randn(3, 128, 128, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()
    dw32, db32 = torch.zeros(3, 3, 3, 64, device="cuda"), torch.zeros(64, device="cuda")
    dw8, db8 = torch.zeros_like(dw32), torch.zeros_like(db32)
    ops.conv_first_wgrad(torch.from_numpy(x).cuda(), dz, dw32, db32)
    ops.conv_first_wgrad(torch.from_numpy(frames).cuda(), dz, dw8, db8)
    assert torch.equal(dw8, dw32) and torch.equal(db8, db32)      # same values in, fixed summation order: same bits


def test_maxpool_backward_routes_to_first_max_and_relu_masks():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    B, H, W, C = 2, 64, 64, 128
    a = _bf(torch.relu(torch.randn(B, H, W, C, generator=g)))
    a[0, :8, :8, :] = 0          # all-zero windows: gradient must vanish
    a[1, 0:2, 0:2, :] = 1.5      # positive ties: the first element of the window wins
    an = a.permute(0, 3, 1, 2).clone().requires_grad_(True)
    p = F.max_pool2d(an, 2, 2)
    dp = _bf(torch.randn(B, H // 2, W // 2, C, generator=g))
    p.backward(dp.permute(0, 3, 1, 2))
    ref = an.grad.permute(0, 2, 3, 1) * (a > 0)
    dz = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    db = torch.full((C,), 0.25, device="cuda")          # fused BiasAddGrad OVERWRITES db
    ops.maxpool_bwd_relu(dp.cuda().to(torch.bfloat16), a.cuda().to(torch.bfloat16),
                         p.detach().permute(0, 2, 3, 1).contiguous().cuda().to(torch.bfloat16), dz, db)
    assert torch.equal(dz.float().cpu(), ref)
    _close(db, ref.sum(dim=(0, 1, 2)), 1e-5, "fused bias grad (maxpool bwd)")


@pytest.mark.parametrize("cin,cout,H,W", [(64, 64, 128, 128), (128, 128, 64, 64), (64, 64, 45, 45), (128, 128, 37, 22)])
def test_pool_routing_codes_reproduce_maxpool_backward_bit_exactly(cin, cout, H, W, igemm_variant):
    """Training keeps 2-bit routing codes from the fused max-pool epilogue instead of the pre-pool activation:
    MaxPoolGrad + ReluGrad through the codes must equal the activation-based kernel bit for bit (first maximum wins,
    zero windows give no gradient)."""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + H)
    B = 2
    x = _bf(torch.randn(B, H, W, cin, generator=g) * 0.5)
    x[0, :6, :6, :] = 0                                   # all-zero windows after ReLU (bias <= 0 there)
    w = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = -torch.rand(cout, generator=g) * 0.05
    wf = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), wf, None)
    full = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    pooled = torch.empty(B, H // 2, W // 2, cout, device="cuda", dtype=torch.bfloat16)
    code = torch.zeros(B, H // 2, W // 2, cout // 8, device="cuda", dtype=torch.int16)
    ops.conv3x3_igemm(x.cuda().to(torch.bfloat16), wf, b.cuda(), None, full, pooled, ops.EPI_BIAS_RELU, pool_code=code)
    dp = (torch.randn(B, H // 2, W // 2, cout, generator=g) * 0.1).to(torch.bfloat16).cuda()
    dz_ref = torch.empty_like(full)
    db_ref = torch.zeros(cout, device="cuda")
    ops.maxpool_bwd_relu(dp, full, pooled, dz_ref, db_ref)
    dp_masked = torch.where(pooled > 0, dp, torch.zeros_like(dp))      # what the dgrad epilogue does (relu'(pooled))
    dz = torch.full_like(full, float("nan"))
    db = torch.zeros(cout, device="cuda")
    ops.maxpool_bwd_code(dp_masked, code, dz, db)
    assert torch.equal(dz, dz_ref)
    _close(db, db_ref, 1e-5, "bias grad through routing codes")
    # against autograd of the plain PyTorch ops (odd sides: the last row / column belongs to no window -> zero gradient)
    a = full.float().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    F.max_pool2d(F.relu(a), 2, 2).backward(dp.float().cpu().permute(0, 3, 1, 2))
    got, want = dz.float().cpu(), a.grad.permute(0, 2, 3, 1)
    assert torch.equal(got, want)                         # same first-maximum rule (row-major scan of the window)
    if H % 2:
        assert not got[:, H - 1].any()
    if W % 2:
        assert not got[:, :, W - 1].any()
    # pooled-only (inference-style) launch writes the same codes
    code2 = torch.zeros_like(code)
    ops.conv3x3_igemm(x.cuda().to(torch.bfloat16), wf, b.cuda(), None, None, pooled, ops.EPI_BIAS_RELU, pool_code=code2)
    assert torch.equal(code, code2)


@pytest.mark.parametrize("cin,cout,H,W", [(64, 128, 64, 64), (128, 256, 32, 32), (64, 128, 12, 20), (128, 256, 24, 24),
                                          (128, 256, 25, 25), (64, 64, 56, 56)])
def test_dgrad_unpool_equals_dgrad_then_maxpool_backward(cin, cout, H, W):
    """abl_conv3x3_dgrad_unpool (dgrad + relu'(pooled) + MaxPoolGrad + BiasAddGrad in one kernel) vs the two-kernel
    path through the routing codes: bit-identical dZ, bias gradient equal up to fp32 order.  (cin = channels of the
    pooled tensor = GEMM N; cout = channels of the incoming gradient = GEMM K.)"""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + cout)
    B = 3
    pooled = _bf(torch.relu(torch.randn(B, H, W, cin, generator=g))).to(torch.bfloat16).cuda()
    code = torch.randint(0, 1 << 16, (B, H, W, cin // 8), generator=g, dtype=torch.int32).to(torch.int16).cuda()
    dz = (torch.randn(B, H, W, cout, generator=g) * 0.1).to(torch.bfloat16).cuda()
    w = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    wd = torch.empty(9, cin, cout, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), None, wd)
    dp = torch.empty(B, H, W, cin, device="cuda", dtype=torch.bfloat16)
    ops.conv3x3_igemm(dz, wd, None, pooled, dp, None, ops.EPI_RELU_MASK)
    ref = torch.empty(B, 2 * H, 2 * W, cin, device="cuda", dtype=torch.bfloat16)
    db_ref = torch.zeros(cin, device="cuda")
    ops.maxpool_bwd_code(dp, code, ref, db_ref)
    out = torch.full_like(ref, float("nan"))
    db = torch.zeros(cin, device="cuda")
    ops.conv3x3_dgrad_unpool(dz, wd, pooled, code, out, db)
    assert torch.equal(out, ref)
    _close(db, db_ref, 1e-4, "fused bias grad (dgrad + unpool)")


@pytest.mark.parametrize("cin,cout,H,W", [(256, 256, 32, 32), (128, 128, 32, 64), (256, 256, 24, 24), (128, 128, 20, 24),
                                          (256, 256, 25, 25), (256, 256, 56, 56)])
def test_conv_with_fused_gap_partial_sums(cin, cout, H, W):
    """abl_conv3x3_igemm_gap: same activation as the plain forward conv, and the epilogue's per-rectangle sums reduce
    to the GlobalAveragePooling of that (bf16) activation; with out_full = NULL the sums are unchanged (inference)."""
    ops = _ops()
    g = torch.Generator().manual_seed(cin + W)
    B = 3
    x = (torch.randn(B, H, W, cin, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    wf = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w.cuda(), wf, None)
    ref = torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16)
    ops.conv3x3_igemm(x, wf, b, None, ref, None, ops.EPI_BIAS_RELU)
    parts = ops.igemm_gap_parts(H, W, cout)
    out = torch.full_like(ref, float("nan"))
    part = torch.full((B * parts, cout), float("nan"), device="cuda")
    ops.conv3x3_igemm_gap(x, wf, b, out, part)
    assert torch.equal(out, ref)
    gap = torch.empty(B, cout, device="cuda")
    ops.gap_reduce(part, gap, parts, H * W)
    _close(gap, ref.float().mean(dim=(1, 2)), 1e-5, "fused GAP")
    part2 = torch.full_like(part, float("nan"))
    ops.conv3x3_igemm_gap(x, wf, b, None, part2)
    assert torch.equal(part, part2)                       # deterministic, independent of the activation store
    # the head kernel finishes the pooling itself (no gap_reduce launch): same bits in gap, same logits
    wh, bh = torch.randn(cout, 2, device="cuda") * 0.1, torch.randn(2, device="cuda") * 0.1
    lg_a, pr_a = torch.empty(B, 2, device="cuda"), torch.empty(B, 2, device="cuda")
    lg_b, pr_b = torch.empty(B, 2, device="cuda"), torch.empty(B, 2, device="cuda")
    ops.head_fwd(gap, wh, bh, lg_a, pr_a)
    gap_b = torch.full_like(gap, float("nan"))
    ops.head_fwd(gap_b, wh, bh, lg_b, pr_b, gap_part=part, gap_parts=parts, gap_hw=H * W)
    assert torch.equal(gap_b, gap) and torch.equal(lg_a, lg_b) and torch.equal(pr_a, pr_b)
    # 1-bit ReLU mask instead of the activation (what a training step keeps), and the GAP gradient rebuilt from it
    if cout % 256:
        return
    bits = torch.full((B, H, W, cout // 32), -1, device="cuda", dtype=torch.int32)
    part3 = torch.full_like(part, float("nan"))
    ops.conv3x3_igemm_gap(x, wf, b, None, part3, relu_bits=bits)
    assert torch.equal(part, part3)
    on = (ref.float() > 0).view(B, H, W, cout // 32, 32).to(torch.int64)
    want = (on << torch.arange(32, device="cuda")).sum(-1)
    want = torch.where(want >= 2 ** 31, want - 2 ** 32, want).to(torch.int32)
    assert torch.equal(bits, want)                        # bit k of word g <=> channel 32 g + k is > 0, exactly
    dgap = torch.randn(B, cout, device="cuda")
    dz_a, dz_b = torch.empty_like(ref), torch.full_like(ref, float("nan"))
    db_a, db_b = torch.empty(cout, device="cuda"), torch.empty(cout, device="cuda")
    ops.gap_bwd_relu(dgap, ref, dz_a, db=db_a)
    ops.gap_bwd_bits(dgap, bits, dz_b, db=db_b)
    assert torch.equal(dz_a, dz_b)
    _close(db_b, db_a, 1e-5, "bias grad from the bit mask (value x count vs running sum)")
    db_c = torch.empty_like(db_b)
    ops.gap_bwd_bits(dgap, bits, dz_b, db=db_c)
    assert torch.equal(db_b, db_c)


def test_gap_forward_backward():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    a = _bf(torch.relu(torch.randn(3, 32, 32, 256, generator=g)))
    gap = torch.empty(3, 256, device="cuda")
    ops.gap_fwd(a.cuda().to(torch.bfloat16), gap)
    _close(gap, a.mean(dim=(1, 2)), 1e-5, "gap")
    dgap = torch.randn(3, 256, generator=g)
    dz = torch.empty(3, 32, 32, 256, device="cuda", dtype=torch.bfloat16)
    db = torch.zeros(256, device="cuda")
    ops.gap_bwd_relu(dgap.cuda(), a.cuda().to(torch.bfloat16), dz, db)
    ref = ((dgap[:, None, None, :] / 1024.0) * (a > 0)).to(torch.bfloat16)
    assert torch.equal(dz.cpu(), ref)
    _close(db, ref.float().sum(dim=(0, 1, 2)), 1e-5, "fused bias grad (gap bwd)")


@pytest.mark.parametrize("B", [1, 37])
def test_head_forward_loss_backward(B):
    ops = _ops()
    g = torch.Generator().manual_seed(B)
    C, n = 256, 2
    gap = torch.rand(B, C, generator=g)
    w = torch.randn(C, n, generator=g) * 0.15
    b = torch.tensor([0.1, -0.2])
    y = torch.zeros(B, n)
    y[torch.arange(B), torch.arange(B) % 2] = 1
    cw = {0: 0.75, 1: 1.5}
    logits, probs = torch.empty(B, n, device="cuda"), torch.empty(B, n, device="cuda")
    loss, dlog = torch.zeros(1, device="cuda"), torch.empty(B, n, device="cuda")
    mask = torch.empty(B, C, device="cuda", dtype=torch.uint8)
    ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, y.cuda(), torch.tensor([0.75, 1.5]).cuda(), 1.0 / B,
                 loss, dlog, mask, False, True, 0.45, 10001, 3, 0)
    keep = mask.float().cpu()
    assert set(np.unique(keep.numpy())) <= {0.0, 1.0}
    if B > 8:
        assert abs(keep.mean().item() - 0.55) < 0.03
    h = (gap * keep / 0.55).requires_grad_(True)
    wt, bt = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    z = h @ wt + bt
    L = oracle.loss_from_logits(z, y.numpy(), cw)          # Keras: sum_i w_i l_i / batch
    L.backward()
    _close(logits, z.detach(), 1e-5, "logits")
    _close(probs, torch.softmax(z.detach(), -1), 1e-5, "probs")
    _close(loss, L.detach().reshape(1), 1e-5, "loss")
    # the batch loss is one fixed-order sum: a second launch ACCUMULATES exactly the same value (2x), and an
    # overwriting launch reproduces it bit for bit (no float atomics)
    first = loss.clone()
    ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, y.cuda(), torch.tensor([0.75, 1.5]).cuda(), 1.0 / B,
                 loss, dlog, mask, True, True, 0.45, 10001, 3, 0)
    assert torch.equal(loss, first + first)
    ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, y.cuda(), torch.tensor([0.75, 1.5]).cuda(), 1.0 / B,
                 loss, dlog, mask, True, True, 0.45, 10001, 3, 0, loss_accumulate=False)
    assert torch.equal(loss, first)
    dw = torch.full((C, n), float("nan"), device="cuda")           # overwritten
    db, dgap = torch.full((n,), float("nan"), device="cuda"), torch.empty(B, C, device="cuda")
    ops.head_bwd(gap.cuda(), w.cuda(), dlog, mask, 0.45, dw, db, dgap)
    dw_b, db_b = torch.empty_like(dw), torch.empty_like(db)
    ops.head_bwd(gap.cuda(), w.cuda(), dlog, mask, 0.45, dw_b, db_b, torch.empty_like(dgap))
    assert torch.equal(dw, dw_b) and torch.equal(db, db_b)
    _close(dw, wt.grad, 1e-5, "dW")
    _close(db, bt.grad, 1e-5, "db")
    _close(dgap, h.grad * keep / 0.55, 1e-5, "dgap")
    # the realised mask is reproducible from (seed, step, rank) and differs across steps and ranks
    m2 = torch.empty_like(mask)
    ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, None, None, 1.0, None, None, m2, False, True, 0.45,
                 10001, 3, 0)
    assert torch.equal(mask, m2)
    if B > 8:
        ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, None, None, 1.0, None, None, m2, False, True,
                     0.45, 10001, 4, 0)
        assert not torch.equal(mask, m2)
        ops.head_fwd(gap.cuda(), w.cuda(), b.cuda(), logits, probs, None, None, 1.0, None, None, m2, False, True,
                     0.45, 10001, 3, 1)
        assert not torch.equal(mask, m2)


@pytest.mark.parametrize("n", [1, 3, 100003, 1736002])
def test_adam_and_rmsprop_flat(n):
    ops = _ops()
    rng = np.random.default_rng(n)
    th0 = rng.normal(size=n).astype(np.float32)
    th, m, v = torch.from_numpy(th0).cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    th_np, m_np, v_np = th0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in (1, 2, 3):
        g = (rng.normal(size=n) * 0.01).astype(np.float32)
        alpha = 3e-4 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ops.adam_flat(th, torch.from_numpy(g).cuda(), m, v, alpha)
        th_np, m_np, v_np = oracle.adam_step(th_np, g, m_np, v_np, t)
    np.testing.assert_allclose(th.cpu().numpy(), th_np, rtol=0, atol=2e-6)     # 3 steps of ~3e-4 each
    np.testing.assert_allclose(v.cpu().numpy(), v_np, rtol=1e-4, atol=1e-12)    # FMA contraction vs NumPy
    th, r = torch.from_numpy(th0).cuda(), torch.zeros(n, device="cuda")
    th_np, r_np = th0.copy(), np.zeros(n, np.float32)
    for t in (1, 2, 3):
        g = (rng.normal(size=n) * 0.01).astype(np.float32)
        ops.rmsprop_flat(th, torch.from_numpy(g).cuda(), r, 9.3e-6)
        th_np, r_np = oracle.rmsprop_step(th_np, g, r_np)
    np.testing.assert_allclose(th.cpu().numpy(), th_np, rtol=0, atol=1e-6)


def test_gradcam_matches_closed_form_oracle():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    B = 5
    a = _bf(torch.relu(torch.randn(B, 32, 32, 256, generator=g)))
    a[4] = 0                    # all-zero activation map -> zeros (the reference would emit NaN)
    w = torch.randn(256, 2, generator=g) * 0.15
    probs = torch.softmax(a.mean(dim=(1, 2)) @ w, -1).contiguous()
    for hw in ((128, 128), (32, 32)):
        heat = torch.empty(B, *hw, device="cuda")
        ops.gradcam(a.cuda().to(torch.bfloat16), probs.cuda(), w.cuda(), heat)
        ref = oracle.gradcam_closed_form(a.numpy(), probs.numpy(), w.numpy(), None if hw == (32, 32) else hw)
        np.testing.assert_allclose(heat.cpu().numpy(), ref, atol=2e-5)
        assert heat[4].abs().max().item() == 0.0


@pytest.mark.parametrize("h,w", [(128, 128), (300, 400), (64, 64), (97, 131), (480, 640), (1, 1)])
def test_resize_nearest_u8_matches_tf_semantics_bit_exactly(h, w):
    """f1: tf.image.resize(NEAREST_NEIGHBOR) of preprocessor.py:80 on raw uint8 frames (down-, up-sampling, odd sizes)."""
    ops = _ops()
    from oracle import preprocess
    rng = np.random.default_rng(h * 1000 + w)
    src = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    dst = torch.zeros(3, 128, 128, 3, device="cuda", dtype=torch.uint8)
    ops.resize_nearest_u8(torch.from_numpy(src).cuda(), dst)
    np.testing.assert_array_equal(dst.cpu().numpy(), preprocess.nearest_resize_tf(src, (128, 128)))


def test_clip_aggregate_matches_reference_rules_bit_exactly_on_runs():
    """f2: every clip in one launch vs the loop-for-loop restatement of predict.py:283-308, 365-376, 164.  Ragged clips:
    empty-ish (1 frame), shorter than the window, exactly the window, long; runs touching both ends."""
    ops = _ops()
    from oracle import clip_rules
    rng = np.random.default_rng(5)
    counts = [1, 2, 3, 4, 5, 31, 32, 33, 64, 257, 1000, 7]
    probs = []
    for i, n in enumerate(counts):
        pb = rng.random(n).astype(np.float32)
        if i % 3 == 0:
            pb[: n // 2 + 1] = 0.9          # run touching the start
        if i % 4 == 1:
            pb[n // 2:] = 0.8               # run touching the end
        if n == 64:
            pb[:] = 0.99                    # all positive
        probs.append(np.stack([1 - pb, pb], 1))
    allp = np.concatenate(probs, 0).astype(np.float32)
    off = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).cuda()
    thr, window = 0.5, 4
    mean, run, win = ops.clip_aggregate(torch.from_numpy(allp).cuda(), off, 1, thr, window)
    mean, run, win = mean.cpu().numpy(), run.cpu().numpy(), win.cpu().numpy()
    for i, p in enumerate(probs):
        assert run[i] == clip_rules.max_contiguous((p[:, 1] > thr).astype(int)), f"clip {i} ({counts[i]} frames)"
        for ct in (1, 3, 40):
            assert int(run[i] >= ct) == clip_rules.contiguity_rule(p, ct, thr)[1]
        assert win[i] == pytest.approx(clip_rules.sliding_window_rule(p, window)[1], abs=1e-6)
        np.testing.assert_allclose(mean[i], clip_rules.average_rule(p.astype(np.float64)), atol=5e-6)   # fp32 sum of <= 1000 terms


def test_error_behaviour_bad_shapes_and_alignment():
    ops = _ops()
    x = torch.zeros(1, 1, 20, 64, device="cuda", dtype=torch.bfloat16)        # a 2x2 pool needs two rows
    w = torch.zeros(9, 64, 64, device="cuda", dtype=torch.bfloat16)
    pl = torch.zeros(1, 1, 10, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="H, W >= 2"):                       # (through the C ABI: the wrapper checks shapes)
        ops._native.call("abl_conv3x3_igemm", x.data_ptr(), w.data_ptr(), torch.zeros(64, device="cuda").data_ptr(), None,
                         None, pl.data_ptr(), None, None, 0, 1, 1, 20, 64, 64, ops.EPI_BIAS_RELU, ops._stream())
    x = torch.zeros(1, 32, 32, 48, device="cuda", dtype=torch.bfloat16)       # Cin not a multiple of 64
    w = torch.zeros(9, 64, 48, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(1, 32, 32, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiples of 64"):
        ops.conv3x3_igemm(x, w, None, None, out, None, 0)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.gap_fwd(torch.zeros(1, 2, 2, 64, dtype=torch.bfloat16), torch.zeros(1, 64, device="cuda"))
    th = torch.zeros(9, device="cuda")
    with pytest.raises(RuntimeError, match="aligned"):
        ops.adam_flat(th[1:], th[1:], th[1:], th[1:], 1e-3)


def _guarded(shape, dtype, fill=0):
    """Output tensor carved out of the middle of a larger allocation whose margins hold a sentinel: an out-of-bounds
    write by a kernel (compute-sanitizer is not available on this pool) shows up as a changed margin."""
    n = int(np.prod(shape))
    pad = 4096
    raw = torch.full((n + 2 * pad,), 77 if dtype in (torch.uint8, torch.int16) else 7.0, device="cuda").to(dtype)
    view = raw[pad:pad + n].view(shape)
    view.fill_(fill)

    def check(what):
        sent = raw[0].item()
        assert bool((raw[:pad] == sent).all()) and bool((raw[pad + n:] == sent).all()), f"{what}: wrote outside its buffer"
    return view, check


def test_kernels_do_not_write_outside_their_output_buffers():
    ops = _ops()
    g = torch.Generator().manual_seed(123)
    bf = torch.bfloat16
    for B in (1, 3):
        # block1_conv2-shaped forward with pool + routing codes, then the un-pooling dgrad of the layer above it
        x = (torch.randn(B, 128, 128, 64, generator=g) * 0.5).to(bf).cuda()
        w = torch.randn(3, 3, 64, 64, generator=g) * 0.05
        wf = torch.empty(9, 64, 64, device="cuda", dtype=bf); wd = torch.empty(9, 64, 64, device="cuda", dtype=bf)
        ops.prep_conv_weights(w.cuda(), wf, wd)
        full, c1 = _guarded((B, 128, 128, 64), bf); pool, c2 = _guarded((B, 64, 64, 64), bf)
        code, c3 = _guarded((B, 64, 64, 8), torch.int16)
        ops.conv3x3_igemm(x, wf, torch.zeros(64, device="cuda"), None, full, pool, ops.EPI_BIAS_RELU, pool_code=code)
        c1("conv fwd full"); c2("conv fwd pooled"); c3("routing codes")
        w2 = torch.randn(3, 3, 64, 128, generator=g) * 0.05
        wd2 = torch.empty(9, 64, 128, device="cuda", dtype=bf)
        ops.prep_conv_weights(w2.cuda(), None, wd2)
        dz = (torch.randn(B, 64, 64, 128, generator=g) * 0.1).to(bf).cuda()
        up, c4 = _guarded((B, 128, 128, 64), bf); db, c5 = _guarded((64,), torch.float32)
        ops.conv3x3_dgrad_unpool(dz, wd2, pool, code, up, db)
        c4("dgrad unpool"); c5("fused bias grad")
        dzf, c6 = _guarded((B, 128, 128, 64), bf)
        ops.maxpool_bwd_code(pool, code, dzf)
        c6("maxpool_bwd_code")
        # last conv with fused GAP partial sums; masked dgrad with column sums
        x3 = (torch.randn(B, 32, 32, 256, generator=g) * 0.5).to(bf).cuda()
        w3 = torch.randn(3, 3, 256, 256, generator=g) * 0.03
        wf3 = torch.empty(9, 256, 256, device="cuda", dtype=bf); wd3 = torch.empty(9, 256, 256, device="cuda", dtype=bf)
        ops.prep_conv_weights(w3.cuda(), wf3, wd3)
        a3, c7 = _guarded((B, 32, 32, 256), bf); part, c8 = _guarded((B * 8, 256), torch.float32)
        ops.conv3x3_igemm_gap(x3, wf3, torch.zeros(256, device="cuda"), a3, part)
        c7("igemm_gap activation"); c8("gap partial sums")
        dx3, c9 = _guarded((B, 32, 32, 256), bf); cs, c10 = _guarded((256,), torch.float32)
        ops.conv3x3_igemm(x3, wd3, None, a3, dx3, None, ops.EPI_RELU_MASK, colsum=cs)
        c9("masked dgrad"); c10("column sums")
        # wgrad: gradient + workspace
        ws_bytes = ops.wgrad_workspace_bytes(B, 128, 128, 64, 64)
        ws, c11 = _guarded((ws_bytes,), torch.uint8); dw, c12 = _guarded((3, 3, 64, 64), torch.float32)
        dbw, c13 = _guarded((64,), torch.float32)
        ops.conv3x3_wgrad(x, full, dw, dbw, ws)
        c11("wgrad workspace"); c12("wgrad dw"); c13("wgrad db")
        # block1_conv1 forward (fp32 frames) and resize
        frames = (torch.rand(B, 128, 128, 3, generator=g) * 255 - 110).cuda()
        a1, c14 = _guarded((B, 128, 128, 64), bf)
        ops.conv_first_fwd(frames, (torch.randn(3, 3, 3, 64, generator=g) * 0.1).cuda(), torch.zeros(64, device="cuda"), a1)
        c14("conv_first fwd")
        rs, c15 = _guarded((B, 128, 128, 3), torch.uint8)
        ops.resize_nearest_u8(torch.randint(0, 256, (B, 77, 91, 3), generator=g, dtype=torch.uint8).cuda(), rs)
        c15("resize")
        # partial-sum buffers sized EXACTLY abl_parts_bytes(C): the producers' per-CTA rows and the header must fit
        for C, make in ((256, lambda pb: ops.conv3x3_igemm(x3, wd3, None, a3, dx3, None, ops.EPI_RELU_MASK, colsum_parts=pb)),
                        (64, lambda pb: ops.conv3x3_dgrad_unpool(dz, wd2, pool, code, up, db_parts=pb)),
                        (256, lambda pb: ops.gap_bwd_relu(torch.ones(B, 256, device="cuda"), a3, dx3, db_parts=pb)),
                        (64, lambda pb: ops.maxpool_bwd_code(pool, code, dzf, db_parts=pb))):
            pb, cp = _guarded((ops.parts_bytes(C) // 4,), torch.float32)
            make(pb)
            cp(f"parts buffer ({C} channels)")
            rows = int(pb[:1].view(torch.int32).item())
            assert 1 <= rows <= (ops.parts_bytes(C) // 4 - ops.PARTS_HEADER) // C
            out = torch.empty(C, device="cuda")
            ops.parts_reduce(pb, out)
            _close(out, pb[ops.PARTS_HEADER:ops.PARTS_HEADER + rows * C].view(rows, C).double().sum(0).float(), 1e-5,
                   "parts_reduce")
        # block1_conv1 filter gradient: per-CTA workspace + fixed-order reduction
        wsf, c16 = _guarded((ops.conv_first_wgrad_workspace_bytes(),), torch.uint8)
        dw1, c17 = _guarded((3, 3, 3, 64), torch.float32); db1, c18 = _guarded((64,), torch.float32)
        ops.conv_first_wgrad(frames, a1, dw1, db1, wsf)
        c16("conv_first wgrad workspace"); c17("conv_first dw"); c18("conv_first db")
        # head: loss workspace sized exactly abl_head_loss_ws_bytes(B)
        lws, c19 = _guarded((int(ops._native.load().abl_head_loss_ws_bytes(B)) // 4,), torch.float32)
        logits, c20 = _guarded((B, 2), torch.float32); probs, c21 = _guarded((B, 2), torch.float32)
        loss, c22 = _guarded((1,), torch.float32)
        yy = torch.zeros(B, 2, device="cuda"); yy[:, 0] = 1
        ops.head_fwd(torch.rand(B, 256, device="cuda"), torch.randn(256, 2, device="cuda"), torch.zeros(2, device="cuda"),
                     logits, probs, yy, None, 1.0 / B, loss, loss_ws=lws)
        c19("head loss workspace"); c20("logits"); c21("probs"); c22("loss")
        assert int(lws[:1].view(torch.int32).item()) == 0        # the ticket reset itself for the next launch
    # partial tiles: every output of a frame that is not a whole number of tiles stays inside its tensor
    B = 2
    x = (torch.randn(B, 24, 40, 64, generator=g) * 0.5).to(bf).cuda()
    w = torch.randn(3, 3, 64, 64, generator=g) * 0.05
    wf = torch.empty(9, 64, 64, device="cuda", dtype=bf); wd = torch.empty(9, 64, 64, device="cuda", dtype=bf)
    ops.prep_conv_weights(w.cuda(), wf, wd)
    full, c1 = _guarded((B, 24, 40, 64), bf); pool, c2 = _guarded((B, 12, 20, 64), bf)
    code, c3 = _guarded((B, 12, 20, 8), torch.int16)
    ops.conv3x3_igemm(x, wf, torch.zeros(64, device="cuda"), None, full, pool, ops.EPI_BIAS_RELU, pool_code=code)
    c1("partial conv fwd full"); c2("partial conv fwd pooled"); c3("partial routing codes")
    w2 = torch.randn(3, 3, 64, 128, generator=g) * 0.05
    wd2 = torch.empty(9, 64, 128, device="cuda", dtype=bf)
    ops.prep_conv_weights(w2.cuda(), None, wd2)
    dz = (torch.randn(B, 12, 20, 128, generator=g) * 0.1).to(bf).cuda()
    up, c4 = _guarded((B, 24, 40, 64), bf, fill=1); db, c5 = _guarded((64,), torch.float32)
    ops.conv3x3_dgrad_unpool(dz, wd2, pool, code, up, db)
    c4("partial dgrad unpool"); c5("partial fused bias grad")
    x3 = (torch.randn(B, 25, 25, 256, generator=g) * 0.5).to(bf).cuda()
    w3 = torch.randn(3, 3, 256, 256, generator=g) * 0.03
    wf3 = torch.empty(9, 256, 256, device="cuda", dtype=bf); wd3 = torch.empty(9, 256, 256, device="cuda", dtype=bf)
    ops.prep_conv_weights(w3.cuda(), wf3, wd3)
    a3, c7 = _guarded((B, 25, 25, 256), bf)
    part, c8 = _guarded((B * ops.igemm_gap_parts(25, 25, 256), 256), torch.float32)
    ops.conv3x3_igemm_gap(x3, wf3, torch.zeros(256, device="cuda"), a3, part)
    c7("partial igemm_gap activation"); c8("partial gap sums")
    dx3, c9 = _guarded((B, 25, 25, 256), bf); cs, c10 = _guarded((256,), torch.float32)
    ops.conv3x3_igemm(x3, wd3, None, a3, dx3, None, ops.EPI_RELU_MASK, colsum=cs)
    c9("partial masked dgrad"); c10("partial column sums")
    ws, c11 = _guarded((ops.wgrad_workspace_bytes(B, 25, 25, 256, 256),), torch.uint8)
    dw, c12 = _guarded((3, 3, 256, 256), torch.float32)
    ops.conv3x3_wgrad(x3, dx3, dw, None, ws)
    c11("partial wgrad workspace"); c12("partial wgrad dw")
    frames = (torch.rand(3, 25, 25, 3, generator=g) * 255 - 110).cuda()          # 1875 pixels: a short last tile
    a1, c14 = _guarded((3, 25, 25, 64), bf)
    ops.conv_first_fwd(frames, (torch.randn(3, 3, 3, 64, generator=g) * 0.1).cuda(), torch.zeros(64, device="cuda"), a1)
    c14("short-tile conv_first fwd")
    torch.cuda.synchronize()

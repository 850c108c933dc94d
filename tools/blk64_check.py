"""Developer tool: the blocked main loop for Cin = Cout = 64 (ABL_BLOCKED64=1, conv3x3_blk64_kernel) against the default
haloed-tile kernel on the same inputs -- forward (+pool, +routing codes), dgrad (+ReLU mask, +fused bias gradient) --
and their device times at batch 256.  python tools/blk64_check.py [B_time] [variant: 1 (default) | 2 | 3]
Variants 2 (ABL_BLOCKED64=2: five ring slots, no zero filter rows) and 3 (= 2 + the pooled-only forward epilogue that
pools horizontally in registers) have not been run on hardware yet."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import ops  # noqa: E402


VARIANT = sys.argv[2] if len(sys.argv) > 2 else "1"


def run(blocked, fn):
    os.environ["ABL_BLOCKED64"] = VARIANT if blocked else "0"
    try:
        return fn()
    finally:
        os.environ["ABL_BLOCKED64"] = "0"


def ulp_diff(a, b):
    a, b = a.float(), b.float()
    return float(((a - b).abs() / (torch.maximum(a.abs(), b.abs()) * 2.0 ** -8 + 1e-6)).max())


g = torch.Generator().manual_seed(3)
ok = True
for (B, H, W) in ((3, 128, 128), (2, 64, 64), (1, 8, 128)):
    x = (torch.randn(B, H, W, 64, generator=g) * 0.5).to(torch.bfloat16).cuda()
    w = (torch.randn(3, 3, 64, 64, generator=g) * (2.0 / 576) ** 0.5).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    wf = torch.empty(9, 64, 64, device="cuda", dtype=torch.bfloat16)
    wd = torch.empty(9, 64, 64, device="cuda", dtype=torch.bfloat16)
    ops.prep_conv_weights(w, wf, wd)

    def fwd():
        full = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        pl = torch.full((B, H // 2, W // 2, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        code = torch.full((B, H // 2, W // 2, 8), -1, device="cuda", dtype=torch.int16)
        ops.conv3x3_igemm(x, wf, b, None, full, pl, ops.EPI_BIAS_RELU, pool_code=code)
        pl2 = torch.full_like(pl, float("nan"))
        ops.conv3x3_igemm(x, wf, b, None, None, pl2, ops.EPI_BIAS_RELU)
        torch.cuda.synchronize()
        return full, pl, code, pl2

    f0, p0, c0, _ = run(False, fwd)
    f1, p1, c1, p1b = run(True, fwd)
    pool_own = torch.nn.functional.max_pool2d(f1.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    # routing codes of the blocked kernel must reproduce max-pool backward of its own output
    dp = (torch.randn(B, H // 2, W // 2, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()
    dz_ref, dz_code = torch.empty_like(f1), torch.empty_like(f1)
    ops.maxpool_bwd_relu(dp, f1, p1, dz_ref, None)
    ops.maxpool_bwd_code(torch.where(p1 > 0, dp, torch.zeros_like(dp)), c1, dz_code, None)
    res = {"fwd_ulp_vs_default": ulp_diff(f0, f1), "nan": bool(torch.isnan(f1.float()).any()),
           "pool_exact_of_own_output": bool(torch.equal(p1.float(), pool_own)),
           "pool_only_equal": bool(torch.equal(p1, p1b)),
           "codes_reproduce_maxpool_bwd": bool(torch.equal(dz_ref, dz_code))}
    act = torch.relu(torch.randn(B, H, W, 64, generator=g)).to(torch.bfloat16).cuda()
    dz = (torch.randn(B, H, W, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()

    def dgrad():
        out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        cs = torch.zeros(64, device="cuda")
        ops.conv3x3_igemm(dz, wd, None, act, out, None, ops.EPI_RELU_MASK, colsum=cs)
        torch.cuda.synchronize()
        return out, cs

    o0, s0 = run(False, dgrad)
    o1, s1 = run(True, dgrad)
    res.update({"dgrad_ulp_vs_default": ulp_diff(o0, o1), "mask_exact": bool((o1[act == 0] == 0).all()),
                "colsum_err": float((s1 - o1.float().sum(dim=(0, 1, 2))).abs().max() / (s0.abs().max() + 1e-9))})
    print((B, H, W), res, flush=True)
    ok = ok and res["fwd_ulp_vs_default"] <= 2.0 and res["dgrad_ulp_vs_default"] <= 2.0 and not res["nan"] and \
        res["pool_exact_of_own_output"] and res["pool_only_equal"] and res["codes_reproduce_maxpool_bwd"] and \
        res["mask_exact"] and res["colsum_err"] < 1e-4
print("blk64 parity:", "OK" if ok else "FAILED", flush=True)

Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 256
xs = [(torch.randn(Bt, 128, 128, 64, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(3)]
act = torch.relu(torch.randn(Bt, 128, 128, 64, device="cuda")).to(torch.bfloat16)
pl = torch.empty(Bt, 64, 64, 64, device="cuda", dtype=torch.bfloat16)
code = torch.empty(Bt, 64, 64, 8, device="cuda", dtype=torch.int16)
out = torch.empty(Bt, 128, 128, 64, device="cuda", dtype=torch.bfloat16)
cs = torch.zeros(64, device="cuda")
for blocked in (False, True):
    def timed(fn):
        for i in range(3):
            fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(12):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 12 * 1e3
    t_f = run(blocked, lambda: timed(lambda i: ops.conv3x3_igemm(xs[i % 3], wf, b, None, None, pl, ops.EPI_BIAS_RELU, pool_code=code)))
    t_d = run(blocked, lambda: timed(lambda i: ops.conv3x3_igemm(xs[i % 3], wd, None, act, out, None, ops.EPI_RELU_MASK, colsum=cs)))
    fl = 2 * Bt * 128 * 128 * 64 * 576 / 1e6
    print(f"blocked={blocked}: forward+pool+codes {t_f:.1f} us ({fl / t_f:.0f} TFLOP/s), dgrad+mask+colsum {t_d:.1f} us "
          f"({fl / t_d:.0f} TFLOP/s) at batch {Bt}", flush=True)

"""Developer tool: the blocked pooled-forward kernel for Cin = Cout = 64 (conv3x3_blk64_pool_kernel, default) against the
haloed-tile kernel (ABL_BLOCKED64=0) on the same inputs, and their device times.  python tools/blk64_check.py [B_time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import ops  # noqa: E402


def run(blocked, fn):
    os.environ["ABL_BLOCKED64"] = "1" if blocked else "0"
    try:
        return fn()
    finally:
        os.environ.pop("ABL_BLOCKED64", None)


g = torch.Generator().manual_seed(3)
w = (torch.randn(3, 3, 64, 64, generator=g) * (2.0 / 576) ** 0.5).cuda()
b = (torch.randn(64, generator=g) * 0.1).cuda()
wf = torch.empty(9, 64, 64, device="cuda", dtype=torch.bfloat16)
ops.prep_conv_weights(w, wf, None)
Bt = int(sys.argv[1]) if len(sys.argv) > 1 else 256
xs = [(torch.randn(Bt, 128, 128, 64, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(3)]
pl = torch.empty(Bt, 64, 64, 64, device="cuda", dtype=torch.bfloat16)
code = torch.empty(Bt, 64, 64, 8, device="cuda", dtype=torch.int16)
outs = {}
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
    run(blocked, lambda: ops.conv3x3_igemm(xs[0], wf, b, None, None, pl, ops.EPI_BIAS_RELU, pool_code=code))
    torch.cuda.synchronize()
    outs[blocked] = (pl.clone(), code.clone())
    fl = 2 * Bt * 128 * 128 * 64 * 576 / 1e6
    print(f"blocked={blocked}: forward+pool+codes {t_f:.1f} us ({fl / t_f:.0f} TFLOP/s) at batch {Bt}", flush=True)
d = (outs[True][0].float() - outs[False][0].float()).abs()
print(f"pooled outputs: max |diff| {float(d.max()):.3e} (bf16 ulp of the largest value "
      f"{float(outs[False][0].float().abs().max()) * 2 ** -8:.3e}); identical values {float((d == 0).float().mean()):.4f}; "
      f"identical codes {float((outs[True][1] == outs[False][1]).float().mean()):.4f}")

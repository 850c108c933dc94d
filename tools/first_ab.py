"""A/B of block1_conv1's forward: A operand in tensor memory (default) vs shared memory (ABL_FIRST_SMEM_A=1).
Usage: python tools/first_ab.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ab_line_classifier_b200 import ops  # noqa: E402


def run(B, H=128, W=128, reps=20):
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(B, H, W, 3, generator=g) * 255 - 115).cuda()
    x8 = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    w = (torch.randn(3, 3, 3, 64, generator=g) * 0.1).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    res = {}
    for mode in ("1", "0"):
        os.environ["ABL_FIRST_SMEM_A"] = mode
        for name, fn in (("fp32", lambda: ops.conv_first_fwd(x, w, b, out)),
                         ("u8", lambda: ops.conv_first_fwd(x8, w, b, out))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            res[(mode, name)] = (us, out.clone())
            gb = (x.numel() * (4 if name == "fp32" else 1) + out.numel() * 2) / 1e9
            print(f"A in {'smem' if mode == '1' else 'tmem'}  {name:5s} {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s")
    dz = (torch.randn(B, H, W, 64, generator=g) * 0.1).to(torch.bfloat16).cuda()
    wres = {}
    for mode in ("1", "0"):
        os.environ["ABL_FIRST_WGRAD_M128"] = mode
        dw, db = torch.empty(3, 3, 3, 64, device="cuda"), torch.empty(64, device="cuda")
        for name, xin in (("fp32", x), ("u8", x8)):
            fn = lambda: ops.conv_first_wgrad(xin, dz, dw, db)  # noqa: E731
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            wres[(mode, name)] = (dw.clone(), db.clone())
            gb = (xin.numel() * xin.element_size() + dz.numel() * 2) / 1e9
            print(f"wgrad M = {'128' if mode == '1' else ' 64'}  {name:5s} {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s (incl. reduction kernel)")
    for name in ("fp32", "u8"):
        same = all(torch.equal(a, b) for a, b in zip(wres[("1", name)], wres[("0", name)]))
        print(f"wgrad {name}: M64 == M128 bit for bit: {same}")
    # what a pure WRITE stream of the same size reaches on this board (the kernel's traffic is 92 % writes)
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    nbytes = out.numel() * 2
    src = res[("0", "fp32")][1]
    for label, fn in (("torch fill_ (write only)", lambda: out.fill_(1.0)),
                      ("cudaMemsetAsync (write only)", lambda: rt.cudaMemsetAsync(out.data_ptr(), 0, nbytes,
                                                                                  torch.cuda.current_stream().cuda_stream)),
                      ("torch sum (read only)", lambda: src.view(torch.int16).view(-1, 8192).sum(1, dtype=torch.int32)),
                      ("torch copy_ (read + write)", lambda: out.copy_(src))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        gb = out.numel() * 2 * (2 if "copy" in label else 1) / 1e9
        print(f"{label:30s} {us:7.1f} us  {gb / us * 1e6:7.0f} GB/s")
    for name in ("fp32", "u8"):
        same = torch.equal(res[("1", name)][1], res[("0", name)][1])
        print(f"{name}: tmem == smem bit for bit: {same}")


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 256)

"""HBM write bandwidth of SM-issued stores on this board (abl_dev_write_probe), beside torch's fill / copy and
cudaMemsetAsync: the write-side roofline of the HBM-bound kernels.  Usage: python tools/write_bw.py [MiB]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ab_line_classifier_b200 import _native  # noqa: E402

MODES = ["st.global.v4", "st.global.cs.v4", "st.global.wt.v4", "st.v4 + L2 evict_first", "st.global.v8",
         "TMA bulk store 4 KiB/warp", "TMA bulk store + evict_first"]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    nbytes = mib << 20
    buf = torch.empty(nbytes, device="cuda", dtype=torch.uint8)
    src = torch.ones(nbytes, device="cuda", dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    print(f"{mib} MiB buffer")
    for mode, name in enumerate(MODES):
        for cps in (1, 2, 4):
            try:
                us = timed(lambda: _native.call("abl_dev_write_probe", buf.data_ptr(), nbytes, mode, cps, st))
            except RuntimeError as e:
                print(f"{name:32s} {cps} CTA/SM  failed: {e}")
                continue
            print(f"{name:32s} {cps} CTA/SM {us:8.1f} us {nbytes / us / 1e3:8.0f} GB/s")
        if mode == 0:
            assert bool((buf.view(torch.int16) == 0x3F80).all()), "probe did not fill the buffer"
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemsetAsync.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
    for label, fn, mult in (("torch fill_", lambda: buf.fill_(7), 1),
                            ("cudaMemsetAsync(0)", lambda: rt.cudaMemsetAsync(buf.data_ptr(), 0, nbytes, st), 1),
                            ("cudaMemsetAsync(0x5a)", lambda: rt.cudaMemsetAsync(buf.data_ptr(), 0x5A, nbytes, st), 1),
                            ("torch copy_ (read + write)", lambda: buf.copy_(src), 2)):
        us = timed(fn)
        print(f"{label:32s}          {us:8.1f} us {mult * nbytes / us / 1e3:8.0f} GB/s")


if __name__ == "__main__":
    main()

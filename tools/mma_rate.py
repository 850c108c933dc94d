"""Developer tool: SM clocks per tcgen05.mma (M = 128, K = 16) versus N and issue pattern (abl_dev_mma_rate_probe).
Run on a B200; prints one line per (N, mode) with clocks per MMA on 1 CTA and on all SMs at once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import _native  # noqa: E402

sms = torch.cuda.get_device_properties(0).multi_processor_count
iters = 2000
st = torch.cuda.current_stream().cuda_stream
print(f"{'N':>4} {'mode':>22} {'clk/MMA 1 CTA':>14} {'clk/MMA all SMs':>16} {'math clk':>9}")
MODES = ((0, "plain SS", 4), (1, "weight-stationary pair", 8), (2, "rotating filter window", 4),
         (4, "2 accumulators plain", 8), (3, "4 accumulators plain", 16), (5, "ws quad fill-use-use-last", 16))
for mode, name, per_it in MODES:
    for n in (64, 128, 192, 256):
        if mode in (1, 5) and n == 192 or mode in (3, 5) and n > 128:
            continue
        res = []
        for ctas in (1, sms):
            out = torch.zeros(ctas, 2, device="cuda", dtype=torch.int64)
            _native.call("abl_dev_mma_rate_probe", n, mode, iters, ctas, out.data_ptr(), st)
            torch.cuda.synchronize()
            per = per_it * iters
            res.append(float(out[:, 1].double().mean()) / per)
        print(f"{n:>4} {name:>22} {res[0]:>14.1f} {res[1]:>16.1f} {n / 2:>9.0f}")

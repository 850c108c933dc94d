"""Developer tool: device time of the augmentation kernel and the metric-counter kernel at the bench batch (256 frames
of 128x128), with the HBM roofline of each (CUDA events, L2 flushed between launches by rotating 12 input batches)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ab_line_classifier_b200 import augment as P  # noqa: E402
from ab_line_classifier_b200 import metrics as M  # noqa: E402
from ab_line_classifier_b200 import ops  # noqa: E402

B, H, W, R = 256, 128, 128, 12
rng = np.random.default_rng(0)
out = {}
for name, dt in (("u8", torch.uint8), ("f32", torch.float32)):
    src = [torch.randint(0, 256, (B, H, W, 3), device="cuda", dtype=torch.uint8).to(dt) for _ in range(R)]
    dst = [torch.empty(B, H, W, 3, device="cuda") for _ in range(R)]
    par = torch.from_numpy(P.params_from_raw(P.draw_raw(rng, B, P.AugmentConfig()), H, W)).cuda()
    for i in range(R):
        ops.augment_frames(src[i], par, dst[i])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for rep in range(5):
        for i in range(R):
            ops.augment_frames(src[i], par, dst[i])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * R) * 1e3
    byts = B * H * W * 3 * (src[0].element_size() + 4)
    out["augment_" + name] = {"us_per_batch256": round(us, 1), "algorithmic_MB": round(byts / 1e6, 1),
                              "GBps": round(byts / us / 1e3, 1), "frames_per_s": round(B / us * 1e6)}
thr = torch.from_numpy(M.auc_thresholds(200)).cuda()
cthr = torch.full((2,), 0.5, device="cuda")
cnt = torch.zeros(ops.metrics_counter_count(2, 200), device="cuda", dtype=torch.int64)
p = torch.softmax(torch.randn(B, 2, device="cuda"), -1)
y = torch.eye(2, device="cuda")[torch.randint(0, 2, (B,), device="cuda")].contiguous()
for _ in range(5):
    ops.metrics_accumulate(p, y, thr, cthr, cnt)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(100):
    ops.metrics_accumulate(p, y, thr, cthr, cnt)
e1.record()
torch.cuda.synchronize()
out["metrics_accumulate"] = {"us_per_batch256": round(e0.elapsed_time(e1) / 100 * 1e3, 2)}
print(json.dumps(out))

"""Developer tool: fit()'s per-batch step (all layers, Adam, device-resident frames) at several batch sizes, CUDA events
over 30 steps.  ABL_TRAIN_GRAPHS=0 disables the whole-step CUDA graph used for batches <= 64."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

eng = CutoffVGG16Engine()
tr = set(eng.layer_names())
for B in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128, 256]:
    fr, lab = synth.synth_frames(B, seed=1)
    x = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
    y = torch.from_numpy(synth.one_hot(lab)).cuda()
    eng.reserve(B, True)
    kw = {"beta_1": 0.9, "beta_2": 0.999, "eps": 1e-7}

    def step(i):      # what fit() does per batch (small batches replay the whole-step CUDA graph)
        slot = eng.stage(x, y, slot=i & 1)
        eng.fit_step(tr, "adam", 3e-4, slot, B, B, **kw)

    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(30):
        step(i)
    e1.record()
    cpu_ms = (time.perf_counter() - w0) / 30 * 1e3           # time to ENQUEUE a step (host side)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"B={B:4d}  {ms:7.3f} ms/step  {B / ms * 1e3:9.0f} frames/s   host enqueue {cpu_ms:6.3f} ms/step")

"""Developer tool: full train step (all layers, Adam) at several batch sizes, CUDA events over 30 steps."""
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
    for _ in range(5):
        eng.train_step(x, y, tr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(30):
        eng.train_step(x, y, tr)
    e1.record()
    cpu_ms = (time.perf_counter() - w0) / 30 * 1e3           # time to ENQUEUE a step (host side)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"B={B:4d}  {ms:7.3f} ms/step  {B / ms * 1e3:9.0f} frames/s   host enqueue {cpu_ms:6.3f} ms/step")

"""Developer tool for ncu: N warm-up training steps + 1 profiled step at batch B (default 256)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = CutoffVGG16Engine()
fr, lab = synth.synth_frames(B, seed=1)
x = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
y = torch.from_numpy(synth.one_hot(lab)).cuda()
tr = set(eng.layer_names())
for _ in range(steps):
    eng.train_step(x, y, tr)
torch.cuda.synchronize()
print("done")

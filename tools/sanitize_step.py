"""Developer tool for `compute-sanitizer --tool memcheck python tools/sanitize_step.py`: one tiny pass over every kernel
(forward, Grad-CAM, clip aggregation, train step with every fusion, both optimizers) at batch 2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ab_line_classifier_b200 import ops, synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

eng = CutoffVGG16Engine()
fr, lab = synth.synth_frames(2, seed=1)
x, y = synth.vgg16_preprocess(fr), synth.one_hot(lab)
tr = set(eng.layer_names())
p = eng.predict_batch(x, use_graph=False)
p8 = eng.predict_batch(fr, use_graph=False)
raw = np.random.default_rng(0).integers(0, 256, size=(2, 200, 300, 3), dtype=np.uint8)
eng.predict_batch(raw, use_graph=False)
heat, _ = eng.gradcam_batch(x)
off = torch.tensor([0, 1, 2], dtype=torch.int64, device="cuda")
ops.clip_aggregate(p.clone(), off, 1, 0.5, 4)
for opt, lr in (("adam", 3e-4), ("rmsprop", 1e-5)):
    eng.train_step(x, y, tr, opt, lr)
    eng.train_step(fr, y, tr, opt, lr)          # uint8 input path
eng.train_step(x, y, set(list(eng.layer_names())[-3:]), "rmsprop", 1e-5)      # reference phase 2
torch.cuda.synchronize()
print("sanitize_step done", float(p.sum()), float(heat.max()))

"""Developer tool: per-launch device times of one training step (CUDA events around every C-ABI call)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import _native, synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = 5
eng = CutoffVGG16Engine()
fr, lab = synth.synth_frames(B, seed=1)
x = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
y = torch.from_numpy(synth.one_hot(lab)).cuda()
tr = set(eng.layer_names())
for _ in range(3):
    eng.train_step(x, y, tr)
torch.cuda.synchronize()
acc = {}
order = []
for r in range(reps):
    _native.timing = []
    eng.train_step(x, y, tr)
    torch.cuda.synchronize()
    recs, _native.timing = _native.timing, None
    for i, (name, a, e0, e1) in enumerate(recs):
        if name == "abl_conv3x3_igemm":
            shape = (a[9], a[10], a[11], a[12], a[13], a[14]); fl = 2.0 * a[9] * a[10] * a[11] * 9 * a[12] * a[13]
        elif name == "abl_conv3x3_igemm_gap":
            shape = (a[6], a[7], a[8], a[9], a[10], "gap"); fl = 2.0 * a[6] * a[7] * a[8] * 9 * a[9] * a[10]
        elif name == "abl_conv3x3_dgrad_unpool":
            shape = (a[7], a[8], a[9], a[10], a[11], "unpool"); fl = 2.0 * a[7] * a[8] * a[9] * 9 * a[10] * a[11]
        elif name == "abl_conv3x3_wgrad":
            shape = (a[8], a[9], a[10], a[11], a[12]); fl = 2.0 * a[8] * a[9] * a[10] * 9 * a[11] * a[12]
        else:
            shape, fl = (), 0.0
        key = (i, name, shape)
        if key not in acc:
            acc[key] = [0.0, fl]
            order.append(key)
        acc[key][0] += e0.elapsed_time(e1) / reps
tot = 0.0
for key in order:
    ms, fl = acc[key]
    tot += ms
    tf = f"{fl / ms / 1e9:8.1f} TF/s" if fl else ""
    print(f"{key[0]:3d} {key[1]:26s} {str(key[2]):34s} {ms * 1e3:9.1f} us {tf}")
print(f"total {tot:.3f} ms for B={B}")

"""Developer tool: host-side profile (cProfile) of the public CutoffVGG16.fit() -- where the Python time of a training
run goes once the device is fast.  python tools/fit_profile.py [N] [B] [device_u8|host_u8]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import metrics as M, synth  # noqa: E402
from ab_line_classifier_b200.data import ArrayDataset  # noqa: E402
from ab_line_classifier_b200.model import CutoffVGG16  # noqa: E402

N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "device_u8"
hp = {"LR_EXTRACT": 0.0003, "LR_FINETUNE": 0.0000093, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7,
      "EXTRACT_EPOCHS": 2}
fr, lab = synth.synth_frames(N, seed=3)
dev = "cuda" if mode.startswith("device") else None
train = ArrayDataset(fr, synth.one_hot(lab), B, shuffle=True, device=dev)
val = ArrayDataset(fr[:1024], synth.one_hot(lab[:1024]), B, device=dev)
mets = ["accuracy", M.AUC(name="auc"), M.Precision(name="precision_b_lines", class_id=1), M.Recall(name="recall_b_lines", class_id=1)]
model = CutoffVGG16(hp, [128, 128, 3], mets, 2)
model.fit(train, epochs=3, validation_data=val, verbose=0)          # warm: allocations, kernel attributes
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
model.fit(train, epochs=4, validation_data=val, verbose=0)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
print(f"[{mode}] fit(): 5 epochs of {N} frames in {dt:.3f} s ({5 * N / dt:.0f} frames/s); host returned after {t_host:.3f} s")

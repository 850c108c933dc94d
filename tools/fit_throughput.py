"""Developer tool: frames/s of the public two-phase CutoffVGG16.fit() on host-resident synthetic frames."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ab_line_classifier_b200 import metrics as M, synth  # noqa: E402
from ab_line_classifier_b200.data import ArrayDataset  # noqa: E402
from ab_line_classifier_b200.model import CutoffVGG16  # noqa: E402

N, B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, int(sys.argv[2]) if len(sys.argv) > 2 else 256
hp = {"LR_EXTRACT": 0.0003, "LR_FINETUNE": 0.0000093, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7,
      "EXTRACT_EPOCHS": 2}
fr, lab = synth.synth_frames(N, seed=3)
mode = sys.argv[3] if len(sys.argv) > 3 else "host_fp32"       # host_fp32 | host_u8 | device_u8 | host_u8_aug | device_u8_aug
aug = {"ZOOM_RANGE": 0.1, "WIDTH_SHIFT_RANGE": 0.2, "HEIGHT_SHIFT_RANGE": 0.2, "ROTATION_RANGE": 45,
       "BRIGHTNESS_RANGE": 0.3} if mode.endswith("_aug") else None     # reference config.yml TRAIN.DATA_AUG
frames = fr if "u8" in mode else synth.vgg16_preprocess(fr)
dev = "cuda" if mode.startswith("device") else None
train = ArrayDataset(frames, synth.one_hot(lab), B, shuffle=True, device=dev, augment=aug)
val = ArrayDataset(frames[:1024], synth.one_hot(lab[:1024]), B, device=dev)
mets = ["accuracy", M.AUC(name="auc"), M.Precision(name="precision_b_lines", class_id=1), M.Recall(name="recall_b_lines", class_id=1)]
model = CutoffVGG16(hp, [128, 128, 3], mets, 2)
t0 = time.perf_counter()
model.fit(train, epochs=4, validation_data=val, verbose=0)      # epochs 0-1 head only (Adam), 1-3 finetune (RMSprop)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
hist = model.model.history
print(f"[{mode}] fit(): 2 extract + 3 finetune epochs of {N} frames (+ validation on 1024) in {dt:.2f} s "
      f"= {5 * N / dt:.0f} train frames/s overall; last logs: { {k: round(float(v[-1]), 4) for k, v in hist.history.items()} }")

"""Developer tool: per-call fixed cost of B200Model.fit (epoch setup / teardown) vs its per-step cost at batch 256."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import synth  # noqa: E402
from ab_line_classifier_b200.model import Adam, CutoffVGG16  # noqa: E402

HP = {"LR_EXTRACT": 3e-4, "LR_FINETUNE": 9.3e-6, "DROPOUT": 0.45, "CUTOFF_LAYER": 10, "FINETUNE_LAYER": 7, "EXTRACT_EPOCHS": 6}
m = CutoffVGG16(HP, [128, 128, 3], ["accuracy"], 2).model
m.compile(optimizer=Adam(3e-4), loss="categorical_crossentropy", metrics=["accuracy"])
B = 256
xs, ys = [], []
for i in range(4):
    fr, lab = synth.synth_frames(B, seed=i)
    xs.append(torch.from_numpy(synth.vgg16_preprocess(fr)).pin_memory())
    ys.append(torch.from_numpy(synth.one_hot(lab)).pin_memory())


def run(steps):
    data = [(xs[i % 4], ys[i % 4]) for i in range(steps)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.fit(data, epochs=1, verbose=0)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


run(5)
for steps in (1, 5, 20, 20, 100):
    ms = run(steps)
    print(f"fit({steps:3d} steps): {ms:8.2f} ms total, {ms / steps:7.3f} ms/step", flush=True)

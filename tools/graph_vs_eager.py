"""Developer tool: is the batch-256 training step bound by launch gaps?  Times engine.fit_step eagerly and through the
whole-step CUDA graph (TRAIN_GRAPH_MAX_BATCH raised for the experiment) and prints the sum of the kernels' own times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ab_line_classifier_b200 import _native, synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = CutoffVGG16Engine()
fr, lab = synth.synth_frames(B, seed=1)
x = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
y = torch.from_numpy(synth.one_hot(lab)).cuda()
tr = frozenset(eng.layer_names())
eng.reserve(B, True)
loss = torch.zeros(1, device="cuda")


def run(n):
    slot = eng.stage(x, y, slot=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(n):
        eng.fit_step(tr, "adam", 3e-4, 0, B, B, loss_out=loss)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for graphs in (False, True):
    eng.train_graphs = graphs
    eng.TRAIN_GRAPH_MAX_BATCH = 4096
    eng._tg_seen.clear(); eng._tgraphs.clear(); eng._dev_counters = None
    run(5)
    print(f"graphs={graphs}: {run(20):.4f} ms/step (graphs captured: {len(eng._tgraphs)})", flush=True)
eng.train_graphs = False
_native.timing = []
eng.fit_step(tr, "adam", 3e-4, 0, B, B, loss_out=loss)
torch.cuda.synchronize()
recs, _native.timing = _native.timing, None
print(f"sum of per-call event times: {sum(e0.elapsed_time(e1) for _, _, e0, e1 in recs):.4f} ms over {len(recs)} calls")

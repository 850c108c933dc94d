"""Developer tool: data-parallel bring-up check on N GPUs (torchrun).  Prints a marker after every stage and dumps
all Python stacks if a stage takes more than 45 s (so a hang is located instead of eating the GPU budget).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/dp_check.py

Checks: NCCL init, blocking + async all-reduce, then that the DP train step (bucketed all-reduce overlapped with
backward) produces the gradient of the single-device step on the concatenated batch."""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
T0 = time.time()


def mark(msg):
    faulthandler.cancel_dump_traceback_later()
    print(f"[rank {os.environ.get('RANK', '0')} +{time.time() - T0:6.1f}s] {msg}", flush=True)
    faulthandler.dump_traceback_later(45, exit=True)


mark("start")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
mark("torch imported")
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
mark("process group up")
t = torch.ones(4, device="cuda") * (rank + 1)
dist.all_reduce(t)
torch.cuda.synchronize()
mark(f"blocking all_reduce ok: {t[0].item()}")
big = torch.ones(1736008, device="cuda")
hs = [dist.all_reduce(big[i * 500000:(i + 1) * 500000], async_op=True) for i in range(3)]
for h in hs:
    h.wait()
torch.cuda.synchronize()
mark(f"async all_reduce ok: {big[0].item()}")

from ab_line_classifier_b200 import synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

Bl = int(sys.argv[1]) if len(sys.argv) > 1 else 8           # frames per rank
G = Bl * world
fr, lab = synth.synth_frames(G, seed=77)
x, y = synth.vgg16_preprocess(fr), synth.one_hot(lab)
keep = (np.random.default_rng(5).random((G, 256)) >= 0.45).astype(np.uint8)
eng = CutoffVGG16Engine(rank=rank, world_size=world)
tr = set(eng.layer_names())
mark("engine built")
sl = slice(rank * Bl, (rank + 1) * Bl)
loss = eng.train_step(x[sl], y[sl], tr, keep_mask=keep[sl], apply_update=False, global_batch=G).clone()
torch.cuda.synchronize()
g_dp = eng.grad.clone()
mark(f"DP step done, local loss {loss.item():.5f}")
# single-device reference on the concatenated batch (every rank computes it; world_size forced to 1)
eng.world_size = 1
loss1 = eng.train_step(x, y, tr, keep_mask=keep, apply_update=False, global_batch=G).clone()
torch.cuda.synchronize()
g_1 = eng.grad.clone()
eng.world_size = world
ls = loss.clone()
dist.all_reduce(ls)
err = float((g_dp - g_1).abs().max() / g_1.abs().max())
mark(f"single-device step done: loss {loss1.item():.5f} vs DP sum {ls.item():.5f}; max|dg|/max|g| = {err:.3e}")
# every frame's dZ is independent of how the batch is split (per-frame bf16 roundings are the same); only the fp32
# summation order of the weight / bias gradient sums differs between the sharded and the concatenated batch
assert err < 1e-4, err
assert abs(ls.item() - loss1.item()) < 1e-5
# the DP gradient itself is bit-reproducible: no float atomics in any kernel, NCCL's reduction order is fixed
loss_b = eng.train_step(x[sl], y[sl], tr, keep_mask=keep[sl], apply_update=False, global_batch=G).clone()
torch.cuda.synchronize()
assert torch.equal(eng.grad, g_dp) and torch.equal(loss_b, loss), "two DP steps on the same inputs differ"
mark("DP step repeated: gradient and loss bit-identical")
# ---- training: the fused peer-memory optimizer (reduce-scatter + Adam/RMSprop + all-gather in one kernel) must keep
# every rank's weights bit-identical, repeat exactly, and equal NCCL all-reduce + the local optimizer kernel.
w0 = eng.get_weights()
steps = 4


def run_steps(opt, lr, xs, ys, ks, gb):
    eng.set_weights(w0)
    eng.opt_state.clear(); eng.opt_steps.clear(); eng.step_count = 0
    out = []
    for i in range(steps):
        eng.train_step(xs, ys, tr, opt, lr, keep_mask=ks, global_batch=gb)
        torch.cuda.synchronize()
        out.append(eng.theta.clone())
    return out


for opt, lr in (("adam", 3e-4), ("rmsprop", 1e-4)):
    th_steps = run_steps(opt, lr, x[sl], y[sl], keep[sl], G)
    th_dp = th_steps[-1]
    ths = [torch.empty_like(th_dp) for _ in range(world)]
    dist.all_gather(ths, th_dp)
    assert all(torch.equal(ths[0], t) for t in ths), f"{opt}: weights diverged across ranks"
    again = run_steps(opt, lr, x[sl], y[sl], keep[sl], G)
    assert all(torch.equal(a, b) for a, b in zip(th_steps, again)), f"{opt}: two fused-DP runs differ"
    # (a) same DP training through NCCL all-reduce + the local optimizer kernel (same update arithmetic, shared device
    #     function).  Two ranks: a + b is commutative, so the two paths must agree BIT FOR BIT after every step.  More
    #     ranks: NCCL's ring/tree order differs from the fused kernel's rank order by fp32 rounding of the gradient sum;
    #     one step must agree to that noise, later steps are reported (a 1e-9 weight difference can flip a bf16 weight
    #     rounding, hence a few ReLU masks, and the normalised optimizer steps amplify that).
    symm_saved, eng._symm = eng._symm, None
    th_nccl_steps = run_steps(opt, lr, x[sl], y[sl], keep[sl], G)
    again = run_steps(opt, lr, x[sl], y[sl], keep[sl], G)
    assert all(torch.equal(a, b) for a, b in zip(th_nccl_steps, again)), f"{opt}: two NCCL-path runs differ"
    d1 = (th_steps[0] - th_nccl_steps[0]).abs()
    d = (th_dp - th_nccl_steps[-1]).abs()
    exact = all(torch.equal(a, b) for a, b in zip(th_steps, th_nccl_steps))
    mark(f"{opt}: fused={symm_saved is not None}; vs NCCL all-reduce + local optimizer: bit-identical over {steps} steps: "
         f"{exact}; after 1 step max diff {float(d1.max()):.3e}; after {steps} steps max {float(d.max()):.3e}, "
         f"{float((d > 2e-5).float().mean()):.2e} of the weights differ by > 2e-5")
    if world == 2 and symm_saved is not None:
        assert exact, "2 ranks: fused peer-memory optimizer != NCCL all-reduce + optimizer"
    # more than two ranks: NCCL adds the ranks in ring/tree order, the fused kernel in rank order -- fp32 rounding of the
    # gradient sum (~1e-7 relative).  The first normalised step turns a gradient difference dg into lr * dg / 3.2e-6
    # (Adam) or lr * dg / 1e-7 (RMSprop) for weights whose gradient is itself at noise level, so the bound is on the
    # FRACTION of weights that moved differently at all, and on one full first step as the worst case.
    if world == 2:
        assert float(d1.max()) < 1e-7, "fused peer-memory optimizer != NCCL all-reduce + optimizer after one step"
    else:
        assert float((d1 > 1e-7).float().mean()) < 1e-3 and float(d1.max()) <= 6.4 * lr, (float(d1.max()), lr)
    assert float((d > 2e-5).float().mean()) < 5e-2 and float(d.median()) < 1e-6
    # (b) single-device training on the concatenated batch (world_size forced to 1): differs from DP only by the fp32
    #     order of the gradient sums -> one step within that noise; several steps as in (a).
    eng.world_size = 1
    th_1_steps = run_steps(opt, lr, x, y, keep, G)
    eng.world_size, eng._symm = world, symm_saved
    th_1 = th_1_steps[-1]
    moved = float((th_1 - torch.from_numpy(np.concatenate([np.pad(w0[k].ravel(), (0, (-w0[k].size) % 4))
                                                              for k in eng.segments])).cuda()).abs().max())
    d1 = (th_steps[0] - th_1_steps[0]).abs()
    d = (th_dp - th_1).abs()
    diff, frac_off = float(d.max()), float((d > 2e-5).float().mean())
    mark(f"{opt}: DP vs 1 GPU on the concatenated batch: after 1 step max diff {float(d1.max()):.3e} "
         f"({float((d1 > 1e-7).float().mean()):.2e} of the weights > 1e-7); after {steps} steps max {diff:.3e}, median "
         f"{float(d.median()):.1e}, {frac_off:.2e} of the weights differ by > 2e-5 (weights moved {moved:.3e})")
    # one step: Adam's first step is lr * g / (|g| + 3.2e-6): an element whose gradient is at summation-noise level
    # (|g| ~ 1e-9) may move differently by a fraction of lr; everything else agrees to fp32 rounding
    assert float((d1 > 1e-7).float().mean()) < 1e-3 and float(d1.max()) <= 6.4 * lr
    assert frac_off < 5e-2 and float(d.median()) < 1e-6 and moved > 1e-4, (opt, diff, frac_off, moved)
# ---- fit()'s per-batch step under data parallel replays a whole-step CUDA graph (forward, loss, backward, the two
# device barriers, the fused peer-memory optimizer, operand re-cast): it must reproduce the eager steps bit for bit.
if eng._symm is not None:
    for opt, lr in (("adam", 3e-4), ("rmsprop", 1e-4)):
        kw = {"beta_1": 0.9, "beta_2": 0.999, "eps": 1e-7} if opt == "adam" else {"rho": 0.9, "eps": 1e-7}
        res = []
        for use_graph in (False, True):
            eng.set_weights(w0)
            eng.opt_state.clear(); eng.opt_steps.clear(); eng.step_count = 0
            eng._tg_seen.clear(); eng._tgraphs.clear(); eng._dev_counters = None
            eng.train_graphs = use_graph
            eng.reserve(Bl, True)
            losses = torch.zeros(6, device="cuda")
            for i in range(6):
                slot = eng.stage(x[sl], y[sl], slot=i & 1)
                eng.fit_step(frozenset(tr), opt, lr, slot, Bl, G, loss_out=losses[i:i + 1], **kw)
            torch.cuda.synchronize()
            res.append((eng.theta.clone(), losses.clone(), len(eng._tgraphs)))
        eng.train_graphs = True
        same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        mark(f"{opt}: whole-step CUDA graph under DP ({res[1][2]} graphs captured) == eager steps bit for bit: {same}")
        assert res[1][2] == 2 and res[0][2] == 0 and same
        ths = [torch.empty_like(res[1][0]) for _ in range(world)]
        dist.all_gather(ths, res[1][0])
        assert all(torch.equal(ths[0], t) for t in ths), f"{opt}: graph replays diverged across ranks"
eng.set_weights(w0)
eng.opt_state.clear(); eng.opt_steps.clear(); eng.step_count = 0
eng._tg_seen.clear(); eng._tgraphs.clear(); eng._dev_counters = None
# ---- metric counters and epoch bookkeeping are summed over the ranks (every Keras metric is a ratio of counts)
from ab_line_classifier_b200 import metrics as M  # noqa: E402
plan = (M.auc_thresholds(200), np.full(2, 0.5, np.float32))
eng.metrics_begin(*plan)
eng.eval_batch(x[sl], y[sl], metrics=True)
local = torch.from_numpy(eng.metrics_read(reduce=False)).cuda()
glob = eng.metrics_read()
parts = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(parts, local)
assert np.array_equal(glob, sum(parts).cpu().numpy()) and int(glob[0]) == G, (glob[:8], G)
eng.world_size = 1
eng.metrics_begin(*plan)
eng.eval_batch(x, y, metrics=True)
full = eng.metrics_read()
eng.world_size = world
assert int(np.abs(glob - full).sum()) <= 2, (glob[:8], full[:8])      # same frames: shards vs one batch on one GPU
tot = eng.sum_over_ranks([1.0, float(rank)])
assert tot == [float(world), world * (world - 1) / 2.0], tot
mark(f"metric counters: {int(glob[0])} frames, {int(glob[1])} correct over {world} ranks; equal to the single-GPU batch: "
     f"{np.array_equal(glob, full)}")
dist.barrier()
faulthandler.cancel_dump_traceback_later()
dist.destroy_process_group()
print(f"[rank {rank}] dp_check ok", flush=True)

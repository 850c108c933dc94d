#!/usr/bin/env python
"""bench.py -- cutoffvgg16 training frames/s on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # B200 arm (hand-written sm_100a kernels)
  python bench.py --gpus N --global-batch 2048 ...         # BASELINE configs[2] as written: 2048 / N frames per GPU
  python bench.py --impl reference --gpus N ...            # CPU arm: torch-CPU restatement of the Keras path
                                                           # (TF 2.9 cannot be installed here)

Workload (config.workload): BASELINE.json configs[1] -- cutoffvgg16 single-GPU training, synthetic
2-class frames at 128x128, batch 256 per GPU, random-init (Glorot) weights, ALL layers trainable
(fwd + 7 wgrad + 6 dgrad + Adam = 18.23 GFLOP/frame).  N>1 without --global-batch: weak scaling, 256
frames per GPU (N=8 -> global batch 2048 = configs[2]); every default run also times configs[2] at its stated split
(global 2048 -> 2048/N per GPU) and reports it under extra.strong_scaling_global_2048.
A step = forward + loss + backward + gradient exchange + Adam update:
  `value`  frames resident in HBM when the timed region starts (engine.train_step on pre-staged batches)
  `e2e`    the same steps through the reference-facing call -- the inner Keras-model `fit()` of the CutoffVGG16 object
           (models.py:427 `self.model.fit(...)`) over batches in pinned HOST memory: every step's H2D copy and the D2H
           copy of its loss are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP = 6.096420864          # per frame
TRAIN_GFLOP = 18.232639488       # fwd + 7 wgrad + 6 dgrad
HPARAMS = {"LR_EXTRACT": 0.0003, "LR_FINETUNE": 0.0000093, "DROPOUT": 0.45, "CUTOFF_LAYER": 10,
           "FINETUNE_LAYER": 7, "EXTRACT_EPOCHS": 6}
PER_GPU_BATCH = 256


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons streamed every 100 ms DURING the timed regions (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                f = [v.strip() for v in line.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
                if self.stop_flag:
                    break
        except Exception:
            pass
        finally:
            if self.proc is not None:
                self.proc.terminate()

    def summary(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        try:
            load = [s for s in self.samples if float(s[2]) > 300.0] or self.samples     # samples under load
            sm = sorted(float(s[0]) for s in load)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
            return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                    "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples),
                    "samples_under_load": len(load), "reasons": reasons}
        except Exception as e:  # malformed nvidia-smi output must not kill the bench line
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"parse error: {e}"]}


def synth_batches(n_batches, batch, seed):
    from ab_line_classifier_b200 import synth
    xs, ys, raw = [], [], []
    for i in range(n_batches):
        fr, lab = synth.synth_frames(batch, seed=seed + i)
        raw.append(fr)
        xs.append(synth.vgg16_preprocess(fr))
        ys.append(synth.one_hot(lab))
    return xs, ys, raw


def cpu_train_frames_per_s(batch, iters, threads, warmup=1):
    """torch-CPU fp32 restatement of the Keras train step (oracle port): fwd + loss + full backward + Adam."""
    import torch
    from oracle import oracle, synth
    torch.set_num_threads(threads)
    fr, lab = synth.synth_frames(batch, seed=1234)
    x, y = synth.vgg16_preprocess(fr), synth.one_hot(lab)
    params = synth.glorot_weights(10001)
    mask = synth.dropout_mask(batch, 256, 0.45)
    m = {k: np.zeros_like(v) for k, v in params.items()}
    v = {k: np.zeros_like(p) for k, p in params.items()}
    times = []
    for it in range(iters + warmup):
        t0 = time.perf_counter()
        _, grads, _ = oracle.loss_and_grads(params, x, y, None, mask)
        for k in params:
            params[k], m[k], v[k] = oracle.adam_step(params[k], grads[k], m[k], v[k], it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch / (sum(times) / len(times)), sum(times) / len(times)


def run_reference(args):
    """CPU arm: rank 0 only; other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    batch = 32
    t_all0 = time.perf_counter()
    # warm-up + timed "steps": each step is one B=32 train step (bounded sample of the B=256 workload)
    fps_list = []
    n_iter = max(1, min(args.steps, 30))         # bounded sample (~0.35 s per step): the whole arm ends within a minute
    n_warm = max(1, min(args.warmup, 5))
    fps, sec = cpu_train_frames_per_s(batch, n_iter, threads, n_warm)
    line = {
        "impl": "reference", "metric": "cutoffvgg16 train frames/sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": n_iter, "warmup": n_warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, PER_GPU_BATCH, False, batch_note=f"CPU sample: batch {batch} per step"),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{n_iter} full train steps (fwd+loss+bwd+Adam) at batch {batch}, torch-CPU fp32 "
                                   "restatement of the Keras path (TensorFlow 2.9 not installable in this image)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, per_gpu_batch, strong, batch_note=None):
    cfg = {"workload": ("cutoffvgg16 training step, all 7 conv layers + Dense trainable, Adam (BASELINE configs[1]; "
                        "N=8 -> global batch 2048 = configs[2])" if not strong else
                        "cutoffvgg16 data-parallel training step at a fixed GLOBAL batch (BASELINE configs[2]), all layers "
                        "trainable, Adam"),
           "img_dim": [128, 128], "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * n_gpus,
           "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
           "gflop_per_frame": TRAIN_GFLOP,
           "l2": "per-step working set ~2 GB at batch 256 (1.05 GB kept activations + gradients) >> 126 MB L2; 4 distinct "
                 "resident input batches rotated"}
    if batch_note:
        cfg["note"] = batch_note
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (default = the BASELINE config)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fixed GLOBAL batch split over the GPUs (BASELINE configs[2]: 2048) -> strong scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 / C5 / strong-scaling measurements in `extra`")
    ap.add_argument("--trainable", default="all", choices=["all", "phase2", "phase1"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3
    # a hung collective must not eat the GPU budget: dump every Python stack and exit after 10 minutes
    import faulthandler
    faulthandler.dump_traceback_later(600, exit=True)

    import torch
    import torch.distributed as dist
    from ab_line_classifier_b200 import _native
    from ab_line_classifier_b200.model import CutoffVGG16, Adam

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    strong = args.global_batch > 0
    if strong and args.global_batch % world:
        raise SystemExit("--global-batch must be a multiple of the number of GPUs")
    B = args.global_batch // world if strong else args.batch
    model = CutoffVGG16(HPARAMS, [128, 128, 3], ["accuracy"], 2, rank=rank, world_size=world)
    inner, eng = model.model, model.model.engine
    names = eng.layer_names()
    trainable = {"all": set(names), "phase2": set(names[-3:]), "phase1": {"output"}}[args.trainable]
    for l in inner.layers:
        l.trainable = l.name in trainable
    inner.compile(optimizer=Adam(HPARAMS["LR_EXTRACT"]), loss="categorical_crossentropy", metrics=["accuracy"])

    n_rot = 4

    def make_batches(batch):
        xs, ys, raw = synth_batches(n_rot, batch, seed=1234 + 100 * rank)
        return ([torch.from_numpy(x).pin_memory() for x in xs], [torch.from_numpy(y).pin_memory() for y in ys],
                [torch.from_numpy(f).pin_memory() for f in raw])

    xs_pin, ys_pin, frames_u8 = make_batches(B)
    eng._ensure(B, True)
    xs_dev = [x.cuda() for x in xs_pin]
    ys_dev = [y.cuda() for y in ys_pin]

    state = {"slot": None}

    def make_device_step(xd, yd, batch):
        def device_step(i):
            # inputs already resident in HBM: the batch is placed into the engine's input slot by a device-to-device
            # copy on the copy stream (batch i+1 while step i computes), then the full step runs on it
            if state["slot"] is None:
                state["slot"] = eng.stage(xd[i % n_rot], yd[i % n_rot])
            slot = state["slot"]
            state["slot"] = eng.stage(xd[(i + 1) % n_rot], yd[(i + 1) % n_rot], slot=slot ^ 1)
            eng.train_step(None, None, trainable, "adam", HPARAMS["LR_EXTRACT"], global_batch=batch * world, slot=slot,
                           batch=batch)
        return device_step

    device_step = make_device_step(xs_dev, ys_dev, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _native.launch_count
    ms = timed(device_step, args.steps)
    launches = _native.launch_count - l0
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)

    # ---- per-kernel device times: instrumented steps (CUDA events around every launch), run RIGHT AFTER the timed region
    # so that they see the same clocks and power state as `value` (an earlier form ran them at the end of the process,
    # after the end-to-end runs, where the board had settled at a lower power-capped clock: its fractions were up to 10 %
    # below what the headline step time implies).  Every rank runs them (they contain the gradient exchange); rank 0
    # keeps the records.  Both input slots are staged and synchronised beforehand, so no input copy overlaps the timed
    # kernels; n_instr steps are enqueued back to back behind ONE spin kernel (~6 ms: the host enqueues 5 x (28 launches
    # + 56 events) while it runs), so the events are consumed back to back, no interval contains host launch latency,
    # and the per-kernel times are averages over n_instr consecutive steps.
    pk, pk_src = peaks()
    roof = None
    kernels = {}
    hbm_kernels = {}
    n_instr = 5
    for sl in (0, 1):
        eng.stage(xs_dev[sl], ys_dev[sl], slot=sl)
    torch.cuda.synchronize()
    eng.train_step(None, None, trainable, "adam", HPARAMS["LR_EXTRACT"], global_batch=B * world, slot=0, batch=B)   # warm-up of this code path
    torch.cuda.synchronize()
    _native.timing = []
    torch.cuda._sleep(12_000_000)
    for rep in range(n_instr):
        eng.train_step(None, None, trainable, "adam", HPARAMS["LR_EXTRACT"], global_batch=B * world, slot=rep & 1, batch=B)
    torch.cuda.synchronize()
    per_step = len(_native.timing) // n_instr          # (0 if the step replayed a CUDA graph: small batches)
    recs_all = [_native.timing[k * per_step:(k + 1) * per_step] for k in range(n_instr)]
    _native.timing = None
    state["slot"] = None
    recs = [(name, a, sum(r[i][2].elapsed_time(r[i][3]) for r in recs_all) / n_instr)
            for i, (name, a, _, _) in enumerate(recs_all[0])]


    # ---- e2e: the reference-facing call.  train.py:135 calls model.fit(...), whose body (models.py:427/432) is the inner
    # Keras model's fit(); here that is B200Model.fit on `steps` batches that live in pinned HOST memory.  Inside the
    # timed region, every step: H2D copy of that step's frames + labels (copy stream, overlapping the previous step),
    # the full train step, the Keras 'accuracy' state update, and an asynchronous D2H copy of the step's loss into
    # pinned memory; at the end of the epoch fit() synchronises and reads the metric counters back (Keras semantics).
    def run_fit(frames_host, steps):
        data = [(frames_host[i % n_rot], ys_pin[i % n_rot]) for i in range(steps)]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hist = inner.fit(data, epochs=1, verbose=0)
        e1.record()
        barrier()
        assert np.isfinite(hist.history["loss"][-1])
        return max_over_ranks(e0.elapsed_time(e1))

    e2e_steps = args.steps
    run_fit(xs_pin, 3)
    ms_e2e = run_fit(xs_pin, e2e_steps)
    e2e_value = B * world * e2e_steps / (ms_e2e / 1e3)
    # same, feeding RAW uint8 RGB frames (vgg16 preprocessing fused into block1_conv1: SURVEY 8 f1)
    run_fit(frames_u8, 3)
    ms_e2e_u8 = run_fit(frames_u8, e2e_steps)

    # Keras Model.train_on_batch: the host waits for every step's loss before it copies the next batch (no overlap)
    def tob_step(i):
        inner.train_on_batch(xs_pin[i % n_rot], ys_pin[i % n_rot])
    for i in range(2):
        tob_step(i)
    ms_tob = timed(tob_step, e2e_steps)
    sampler.stop_flag = True

    # ---- inference throughput (extra): forward only, resident inputs
    state["slot"] = None

    def infer_step(i):
        eng.load_input(xs_dev[i % n_rot])
        eng.forward_device(B, need_last_act=False)
        eng.release_slot()
    for i in range(3):
        infer_step(i)
    ms_inf = timed(infer_step, args.steps)
    infer_fps = B * world * args.steps / (ms_inf / 1e3)

    if rank == 0:
        tc_flops, tc_ms = 0.0, 0.0
        npix = B * 128 * 128
        hbm_bytes = {   # algorithmic bytes per launch of the HBM-bound kernels
            "abl_conv_first_fwd": npix * (3 * 4 + 64 * 2), "abl_conv_first_wgrad": npix * (3 * 4 + 64 * 2),
            "abl_gap_bwd_relu": B * 1024 * 256 * 4, "abl_gap_bwd_bits": B * 1024 * 256 * 2 + B * 1024 * 256 // 8,
            "abl_adam_flat": 28 * eng.n_flat,
            "abl_prep_conv_weights_multi": 8 * (eng.n_params - 514 - 1472)}
        for name, a, t in recs:
            k = kernels.setdefault(name, {"calls": 0, "ms": 0.0, "gflop": 0.0})
            k["calls"] += 1
            k["ms"] += t
            if name == "abl_conv3x3_igemm":        # (x, w, bias, mask, full, pool, code, parts, bytes, B, H, W, Cin, Cout, ...)
                fl = 2.0 * a[9] * a[10] * a[11] * 9 * a[12] * a[13]
            elif name == "abl_conv3x3_wgrad":      # (x, dz, dw, db, db_parts, ws, ws_bytes, acc, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[8] * a[9] * a[10] * 9 * a[11] * a[12]
            elif name == "abl_conv3x3_dgrad_unpool":   # (dz, w, pooled, code, out, parts, bytes, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[7] * a[8] * a[9] * 9 * a[10] * a[11]
            elif name == "abl_conv3x3_igemm_gap":      # (x, w, bias, out, gap_part, relu_bits, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[6] * a[7] * a[8] * 9 * a[9] * a[10]
            else:
                fl = 0.0
            k["gflop"] += fl / 1e9
            if fl:
                tc_flops += fl
                tc_ms += t
        for name, k in kernels.items():
            k["ms"] = round(k["ms"], 4)
            k["tflops"] = round(k["gflop"] / k["ms"], 1) if k["gflop"] and k["ms"] else None
            k["gflop"] = round(k["gflop"], 2)
            if name in hbm_bytes and k["ms"]:
                gbs = hbm_bytes[name] / (k["ms"] * 1e-3) / 1e9
                hbm_kernels[name] = {"ms": k["ms"], "algorithmic_mb": round(hbm_bytes[name] / 1e6, 2), "gb_per_s": round(gbs, 1),
                                     "frac_of_hbm_peak": round(gbs / pk.get("hbm_gbs", 6650.0), 4)}
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else 0.0
        # the timed region is a burst of ~0.1 s (20 steps of ~4 ms): the denominator is the BURST cuBLAS figure; the
        # sustained one (measured at the clocks a 4 s GEMM loop settles to) is reported beside it
        peak = pk.get("bf16_tflops", 1590.0)
        peak_sus = pk.get("bf16_tflops_sustained", 1400.0)
        step_ms_instr = sum(k["ms"] for k in kernels.values())
        # DRAM traffic of the same launches from the committed `ncu --set full` capture of one train step
        # (tools/ncu_summary.py): sum over the conv-family launches of one step
        traffic, traffic_note = None, "no ncu capture found under profiles/"
        for cap_name in ("r2_ncu_step.json", "r1_ncu_step.json"):
            try:
                with open(os.path.join(ROOT, "profiles", cap_name)) as f:
                    cap = json.load(f)["launches"]
                fam = [r for r in cap if ("conv3x3_" in r["kernel"] or "wgrad_tc_kernel" in r["kernel"]
                                         or "wgrad64_tc_kernel" in r["kernel"]) and "conv_first" not in r["kernel"]]
                if fam and B == PER_GPU_BATCH:
                    traffic = round(sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in fam) / 1e9, 3)
                    traffic_note = (f"GB per step: dram__bytes_read.sum + dram__bytes_write.sum over the {len(fam)} conv-family "
                                    f"launches of one step (profiles/{cap_name}); algorithmic activation+gradient bytes "
                                    "of the same launches: 9.45 GB (DESIGN.md section 4)")
                    break
            except Exception:
                continue
        roof = {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "tcgen05 implicit-GEMM conv family (conv3x3_halo_kernel / conv3x3_blk64_pool_kernel fwd, dgrad, "
                          "dgrad+unpool, wgrad_tc_kernel, wgrad64_tc_kernel): all launches of one step, algorithmic FLOPs / "
                          "summed CUDA-event durations (mean of 5 consecutive instrumented steps right after the timed region, inputs staged beforehand)",
                "peak_source": pk_src + ", bf16_tflops (burst: the timed region lasts ~0.1 s)",
                "frac_of_sustained_peak": round(achieved / peak_sus, 4), "sustained_peak": peak_sus,
                "share_of_step": round(tc_ms / step_ms_instr, 4) if step_ms_instr else None,
                "hbm_bound_kernels": hbm_kernels}

    # ---- extra: BASELINE configs[2] at its stated split (global 2048 -> 2048/N per GPU), configs[3] inference sweep,
    # configs[4] Grad-CAM over a 512-frame clip
    extras = {}
    if not args.no_extras and not strong:
        G = 2048
        if G % world == 0:
            B2 = G // world
            eng._ensure(B2, True)
            if B2 != B:         # one synthetic batch on the host, four distinct resident batches (rolled along the frames)
                xs1, ys1, _ = synth_batches(1, B2, seed=4321 + 100 * rank)
                xd0, yd0 = torch.from_numpy(xs1[0]).cuda(), torch.from_numpy(ys1[0]).cuda()
                x2d = [xd0.roll(k * (B2 // n_rot), 0) for k in range(n_rot)]
                y2d = [yd0.roll(k * (B2 // n_rot), 0) for k in range(n_rot)]
            else:
                x2d, y2d = xs_dev, ys_dev
            state["slot"] = None
            step2 = make_device_step(x2d, y2d, B2)
            for i in range(3):
                step2(i)
            n2 = max(3, min(args.steps, 10))
            ms2 = timed(step2, n2)
            extras["strong_scaling_global_2048"] = {
                "global_batch": G, "per_gpu_batch": B2, "n_gpus": world, "steps": n2, "ms_per_step": round(ms2 / n2, 4),
                "frames_per_s": round(G * n2 / (ms2 / 1e3), 1),
                "hbm_gb_allocated_per_gpu": round(torch.cuda.max_memory_allocated() / 1e9, 2),
                "note": "BASELINE configs[2] as written: the same 2048-frame global batch on 1/2/4/8 GPUs (2048/1024/512/256 "
                        "frames per GPU); strong-scaling efficiency = frames_per_s(N) / (N * frames_per_s(1))"}
            del x2d, y2d
        if world == 1:
            state["slot"] = None
            sweep = []
            from ab_line_classifier_b200 import synth
            fr_all, _ = synth.synth_frames(512, seed=3)
            x512 = torch.from_numpy(synth.vgg16_preprocess(fr_all)).cuda()
            flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            for Bi in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]:
                xi = x512[:Bi] if Bi <= 512 else x512.repeat(Bi // 512, 1, 1, 1)
                eng.reserve(Bi, False)
                flush = Bi * 10.06e6 < 126e6          # working set below the L2 size: flush between iterations
                for _ in range(3):
                    eng.predict_batch(xi)
                iters = 20 if Bi <= 256 else 5
                tot = 0.0
                for _ in range(iters):
                    if flush:
                        flush_buf.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    eng.predict_batch(xi)
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
                msi = tot / iters
                sweep.append({"batch": Bi, "ms": round(msi, 4), "frames_per_s": round(Bi / msi * 1e3, 1),
                              "tflops": round(Bi * FWD_GFLOP / msi, 1), "l2_flushed": flush,
                              "cuda_graph": Bi <= eng.GRAPH_MAX_BATCH})
                del xi
            extras["inference_sweep_batch_1_to_4096"] = {
                "rows": sweep, "note": "BASELINE configs[3] (predict.py path): engine.predict_batch on frames resident in HBM, "
                                       "CUDA events per call; batches <= 64 replay the captured forward graph"}
            for _ in range(2):
                eng.gradcam_batch(x512)
            tot = 0.0
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.gradcam_batch(x512)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            extras["gradcam_512_frame_clip"] = {
                "ms": round(tot / 5, 4), "heatmaps_per_s": round(512 / (tot / 5) * 1e3, 1),
                "note": "BASELINE configs[4]: forward to block3_conv3 + closed-form Grad-CAM + max-normalise + bilinear 32->128 "
                        "for 512 frames (gradcam.py:68-83), frames resident in HBM"}
            del x512, flush_buf

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        fps, sec = cpu_train_frames_per_s(32, 2, threads)
        cpu = {"value": round(fps, 2), "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "2 timed full train steps (fwd+loss+bwd+Adam) at batch 32 after 1 warm-up, torch-CPU fp32 "
                         "restatement of the Keras path (TensorFlow 2.9 not installable in this image)"}

    clocks = sampler.summary()          # also stops this rank's nvidia-smi child
    if rank == 0:
        line = {
            "metric": "cutoffvgg16 train frames/sec", "value": round(value, 1), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(world, B, strong), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "frames/s",
                    "h2d_bytes_per_step": int(xs_pin[0].numel() * 4 + ys_pin[0].numel() * 4) * world,
                    "d2h_bytes_per_step": 4 * world, "steps": e2e_steps, "ms_per_step": round(ms_e2e / e2e_steps, 4),
                    "api": "CutoffVGG16(...).model.fit(batches in pinned host memory, epochs=1): per step H2D frames+labels, "
                           "train step, Keras 'accuracy' update, async D2H of the loss; metric counters (3.3 KB) read back "
                           "once at the end of the epoch"},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu,
            "extra": {"train_tflops_per_gpu": round(value / world * TRAIN_GFLOP / 1e3, 1),
                      "whole_step_frac_of_burst_peak": round(value / world * TRAIN_GFLOP / 1e3 / pk.get("bf16_tflops", 1590.0), 4),
                      "whole_step_frac_of_sustained_peak": round(value / world * TRAIN_GFLOP / 1e3 / pk.get("bf16_tflops_sustained", 1400.0), 4),
                      "infer_frames_per_s": round(infer_fps, 1), "infer_ms_per_step": round(ms_inf / args.steps, 4),
                      "infer_tflops_per_gpu": round(infer_fps / world * FWD_GFLOP / 1e3, 1),
                      "e2e_uint8_frames": {"value": round(B * world * e2e_steps / (ms_e2e_u8 / 1e3), 1), "unit": "frames/s",
                                           "h2d_bytes_per_step": int(frames_u8[0].numel() + ys_pin[0].numel() * 4) * world,
                                           "note": "same fit() call on raw uint8 RGB frames, vgg16 preprocessing fused into block1_conv1"},
                      "e2e_train_on_batch": {"value": round(B * world * e2e_steps / (ms_tob / 1e3), 1), "unit": "frames/s",
                                             "ms_per_step": round(ms_tob / e2e_steps, 4),
                                             "note": "Keras train_on_batch: the host blocks on every step's loss, so the next "
                                                     "batch's H2D copy cannot overlap the step"},
                      "trainable": args.trainable, "kernels_one_step": kernels, **extras},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

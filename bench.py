#!/usr/bin/env python
"""bench.py -- cutoffvgg16 training frames/s on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # B200 arm (hand-written sm_100a kernels)
  python bench.py --impl reference --gpus N ...            # CPU arm: torch-CPU restatement of the Keras path
                                                           # (TF 2.9 cannot be installed here)

Workload (config.workload): BASELINE.json configs[1] -- cutoffvgg16 single-GPU training, synthetic
2-class frames at 128x128, batch 256 per GPU, random-init (Glorot) weights, ALL layers trainable
(fwd + 7 wgrad + 6 dgrad + Adam = 18.23 GFLOP/frame).  N>1: weak scaling, 256 frames per
GPU (N=8 -> global batch 2048 = configs[2]) with the bucketed NCCL gradient all-reduce.
A step = H2D-free forward + loss + backward + all-reduce + Adam update on frames resident in HBM
(`value`); `e2e` is the same step through the public model API from pinned host buffers with the
loss read back every step.
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
    n_iter = max(1, min(args.steps, 8))          # bounded sample: the whole arm ends within a few minutes
    n_warm = max(1, min(args.warmup, 2))
    fps, sec = cpu_train_frames_per_s(batch, n_iter, threads, n_warm)
    line = {
        "impl": "reference", "metric": "cutoffvgg16 train frames/sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": n_iter, "warmup": n_warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, batch_note=f"CPU sample: batch {batch} per step"),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{n_iter} full train steps (fwd+loss+bwd+Adam) at batch {batch}, torch-CPU fp32 "
                                   "restatement of the Keras path (TensorFlow 2.9 not installable in this image)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, batch_note=None):
    cfg = {"workload": "cutoffvgg16 training step, all 7 conv layers + Dense trainable, Adam (BASELINE configs[1]; "
                       "N=8 -> global batch 2048 = configs[2])",
           "img_dim": [128, 128], "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * n_gpus,
           "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
           "gflop_per_frame": TRAIN_GFLOP,
           "l2": "per-step working set ~2 GB (1.05 GB kept activations + gradients) >> 126 MB L2; 4 distinct resident input batches rotated"}
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    B = args.batch
    model = CutoffVGG16(HPARAMS, [128, 128, 3], ["accuracy"], 2, rank=rank, world_size=world)
    inner, eng = model.model, model.model.engine
    names = eng.layer_names()
    trainable = {"all": set(names), "phase2": set(names[-3:]), "phase1": {"output"}}[args.trainable]
    for l in inner.layers:
        l.trainable = l.name in trainable
    inner.compile(optimizer=Adam(HPARAMS["LR_EXTRACT"]), loss="categorical_crossentropy", metrics=["accuracy"])

    n_rot = 4
    xs, ys, raw_frames = synth_batches(n_rot, B, seed=1234 + 100 * rank)
    xs_pin = [torch.from_numpy(x).pin_memory() for x in xs]
    ys_pin = [torch.from_numpy(y).pin_memory() for y in ys]
    eng._ensure(B, True)
    xs_dev = [x.cuda() for x in xs_pin]
    ys_dev = [y.cuda() for y in ys_pin]

    state = {"slot": None}

    def device_step(i):
        # inputs already resident in HBM: the batch is placed into the engine's input slot by a device-to-device copy
        # on the copy stream (batch i+1 while step i computes), then the full step runs on it
        if state["slot"] is None:
            state["slot"] = eng.stage(xs_dev[i % n_rot], ys_dev[i % n_rot])
        slot = state["slot"]
        state["slot"] = eng.stage(xs_dev[(i + 1) % n_rot], ys_dev[(i + 1) % n_rot], slot=slot ^ 1)
        eng.train_step(None, None, trainable, "adam", HPARAMS["LR_EXTRACT"], global_batch=B * world, slot=slot, batch=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(args.warmup):
        device_step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _native.launch_count
    ms = timed(device_step, args.steps)
    launches = _native.launch_count - l0
    ms_per_step = ms / args.steps
    value = B * world * args.steps / (ms / 1e3)

    # ---- e2e: public engine API from pinned HOST buffers.  Every step: H2D copy of that step's frames + labels
    # (staged on the copy stream while the previous step computes), the full train step, and a D2H read of the
    # step's loss into pinned memory.
    ring = torch.zeros(64, device="cuda")
    loss_host = torch.zeros(64).pin_memory()

    def run_e2e(frames_host, steps):
        state["slot"] = None
        def body(_):
            pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        slot = eng.stage(frames_host[0], ys_pin[0])
        for i in range(steps):
            nslot = eng.stage(frames_host[(i + 1) % n_rot], ys_pin[(i + 1) % n_rot], slot=slot ^ 1) if i + 1 < steps else slot
            r = i % 64
            ring[r:r + 1].zero_()
            eng.train_step(None, None, trainable, "adam", HPARAMS["LR_EXTRACT"], global_batch=B * world, slot=slot,
                           batch=B, loss_out=ring[r:r + 1])
            loss_host[r:r + 1].copy_(ring[r:r + 1], non_blocking=True)
            slot = nslot
        e1.record()
        barrier()
        ms_ = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        assert torch.isfinite(loss_host[: min(steps, 64)]).all()
        return ms_

    e2e_steps = args.steps
    run_e2e(xs_pin, 3)
    ms_e2e = run_e2e(xs_pin, e2e_steps)
    e2e_value = B * world * e2e_steps / (ms_e2e / 1e3)
    # same, feeding RAW uint8 RGB frames (vgg16 preprocessing fused into block1_conv1: SURVEY 8 f1)
    frames_u8 = [torch.from_numpy(f).pin_memory() for f in raw_frames]
    run_e2e(frames_u8, 3)
    ms_e2e_u8 = run_e2e(frames_u8, e2e_steps)
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

    # ---- per-kernel device times of ONE instrumented step (CUDA events around every launch)
    pk, pk_src = peaks()
    roof = None
    kernels = {}
    # every rank runs the instrumented step (it contains the gradient all-reduce); rank 0 keeps the records
    state["slot"] = None
    _native.timing = []
    device_step(0)
    torch.cuda.synchronize()
    recs, _native.timing = _native.timing, None
    if rank == 0:
        tc_flops, tc_ms = 0.0, 0.0
        for name, a, e0, e1 in recs:
            t = e0.elapsed_time(e1)
            k = kernels.setdefault(name, {"calls": 0, "ms": 0.0, "gflop": 0.0})
            k["calls"] += 1
            k["ms"] += t
            if name == "abl_conv3x3_igemm":        # (x, w, bias, mask, full, pool, code, parts, bytes, B, H, W, Cin, Cout, ...)
                fl = 2.0 * a[9] * a[10] * a[11] * 9 * a[12] * a[13]
            elif name == "abl_conv3x3_wgrad":      # (x, dz, dw, db, db_parts, ws, ws_bytes, acc, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[8] * a[9] * a[10] * 9 * a[11] * a[12]
            elif name == "abl_conv3x3_dgrad_unpool":   # (dz, w, pooled, code, out, parts, bytes, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[7] * a[8] * a[9] * 9 * a[10] * a[11]
            elif name == "abl_conv3x3_igemm_gap":      # (x, w, bias, out, gap_part, B, H, W, Cin, Cout, stream)
                fl = 2.0 * a[5] * a[6] * a[7] * 9 * a[8] * a[9]
            else:
                fl = 0.0
            k["gflop"] += fl / 1e9
            if fl:
                tc_flops += fl
                tc_ms += t
        for k in kernels.values():
            k["ms"] = round(k["ms"], 4)
            k["tflops"] = round(k["gflop"] / k["ms"], 1) if k["gflop"] and k["ms"] else None
            k["gflop"] = round(k["gflop"], 2)
        achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else 0.0
        peak = pk.get("bf16_tflops_sustained", 1400.0)
        step_ms_instr = sum(k["ms"] for k in kernels.values())
        # DRAM traffic of the same launches from the committed `ncu --set full` capture of one train step
        # (profiles/r1_ncu_step.json, tools/ncu_summary.py): sum over the conv-family launches of one step
        traffic, traffic_note = None, "no ncu capture found under profiles/"
        try:
            with open(os.path.join(ROOT, "profiles", "r1_ncu_step.json")) as f:
                cap = json.load(f)["launches"]
            fam = [r for r in cap if ("conv3x3_" in r["kernel"] or "wgrad_tc_kernel" in r["kernel"]
                                     or "wgrad64_tc_kernel" in r["kernel"]) and "conv_first" not in r["kernel"]]
            if fam and B == PER_GPU_BATCH:
                traffic = round(sum(r.get("dram_read_bytes", 0) + r.get("dram_write_bytes", 0) for r in fam) / 1e9, 3)
                traffic_note = (f"GB per step: dram__bytes_read.sum + dram__bytes_write.sum over the {len(fam)} conv-family "
                                "launches of one step (profiles/r1_ncu_step.json); algorithmic activation+gradient bytes "
                                "of the same launches: see DESIGN.md section 4")
        except Exception:
            pass
        roof = {"bound": "tensor", "achieved": round(achieved, 1), "peak": peak, "unit": "TFLOP/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "tcgen05 implicit-GEMM conv family (conv3x3_halo_kernel fwd / dgrad / dgrad+unpool, wgrad_tc_kernel, wgrad64_tc_kernel): all "
                          "launches of one step, algorithmic FLOPs / summed CUDA-event durations",
                "peak_source": pk_src + ", bf16_tflops_sustained (kernels timed inside a long step)",
                "share_of_step": round(tc_ms / step_ms_instr, 4) if step_ms_instr else None}

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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "frames/s",
                    "h2d_bytes_per_step": int(xs_pin[0].numel() * 4 + ys_pin[0].numel() * 4) * world,
                    "d2h_bytes_per_step": 4 * world, "steps": e2e_steps, "ms_per_step": round(ms_e2e / e2e_steps, 4),
                    "api": "engine.stage(pinned fp32 frames, labels) + engine.train_step(slot) + D2H loss, every step"},
            "gpu_launches": launches,
            "roofline": roof, "cpu_baseline": cpu,
            "extra": {"train_tflops_per_gpu": round(value / world * TRAIN_GFLOP / 1e3, 1),
                      "frac_of_sustained_peak": round(value / world * TRAIN_GFLOP / 1e3 / pk.get("bf16_tflops_sustained", 1400.0), 4),
                      "infer_frames_per_s": round(infer_fps, 1), "infer_ms_per_step": round(ms_inf / args.steps, 4),
                      "infer_tflops_per_gpu": round(infer_fps / world * FWD_GFLOP / 1e3, 1),
                      "e2e_uint8_frames": {"value": round(B * world * e2e_steps / (ms_e2e_u8 / 1e3), 1), "unit": "frames/s",
                                           "h2d_bytes_per_step": int(frames_u8[0].numel() + ys_pin[0].numel() * 4) * world,
                                           "note": "raw uint8 RGB frames in, vgg16 preprocessing fused into block1_conv1"},
                      "trainable": args.trainable, "kernels_one_step": kernels},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Measurements for the BASELINE.json configs that bench.py's single line does not carry:
  C2  train step at batch 256: all layers / reference phase 2 (block3_conv2+3 + head, RMSprop) / phase 1 (head, Adam)
  C4  inference sweep, batch 1 .. 4096 on one GPU (predict.py path), small batches through the captured CUDA graph
  C5  Grad-CAM heat-maps for a clip of 512 frames (explainability/gradcam.py path)
plus achieved HBM GB/s of the memory-bound kernels.  One JSON object on stdout (commit it under profiles/).

Timing: CUDA events on the launching stream after warm-up; inputs resident in HBM; between timed iterations a
256 MiB buffer is overwritten whenever the working set of the case is smaller than the 126 MB L2.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ab_line_classifier_b200 import _native, synth  # noqa: E402
from ab_line_classifier_b200.engine import CutoffVGG16Engine  # noqa: E402

FWD_GFLOP, TRAIN_GFLOP, PHASE2_GFLOP = 6.096420864, 18.232639488, 9.72029952
PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters, flush):
    """Mean ms per call; with flush the L2 is overwritten (untimed) before every call."""
    if not flush:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    tot = 0.0
    for _ in range(iters):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "peaks": PEAKS}
    eng = CutoffVGG16Engine()
    names = eng.layer_names()

    # ---------------------------------------------------------------- C4 inference sweep
    fr, _ = synth.synth_frames(4096, seed=3)
    x_all = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
    sweep = []
    for B in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]:
        x = x_all[:B]
        eng.reserve(B, False)
        ws_bytes = B * 10.06e6                                 # min HBM traffic per frame
        flush = ws_bytes < 126e6
        row = {"batch": B, "l2_flushed": flush}
        for mode in (["graph", "launches"] if B <= eng.GRAPH_MAX_BATCH else ["launches"]):
            fn = (lambda: eng.predict_batch(x, use_graph=True)) if mode == "graph" else (
                lambda: eng.predict_batch(x, use_graph=False))
            for _ in range(3):
                fn()
            iters = 30 if B <= 256 else 8
            ms = timed(fn, iters, flush)
            row[mode] = {"ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                         "tflops": round(B * FWD_GFLOP / ms, 1)}
        sweep.append(row)
    out["inference_sweep"] = sweep
    del x_all

    # ---------------------------------------------------------------- predict() through the public model API, HOST frames
    # (reference predict.py:54-87 path: dataset of frames -> model.predict -> probabilities on the host); H2D copies of
    # batch i+1 overlap the forward pass of batch i, one D2H of the [N,2] probabilities at the end.
    from ab_line_classifier_b200.data import ArrayDataset
    from ab_line_classifier_b200.model import B200Model
    model = B200Model(eng, eng.input_shape, eng.n_classes, eng.dropout)
    api = {}
    for label, frames_host in (("fp32_preprocessed", synth.vgg16_preprocess(fr)), ("uint8_raw", fr)):
        ds = ArrayDataset(frames_host, None, 512)
        model.predict(ds)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        import time
        w0 = time.perf_counter()
        for _ in range(3):
            p = model.predict(ds)
        wall = (time.perf_counter() - w0) / 3
        api[label] = {"frames": len(fr), "batch": 512, "ms": round(wall * 1e3, 3),
                      "frames_per_s": round(len(fr) / wall, 1), "h2d_bytes": int(ds.x.numel() * ds.x.element_size())}
    out["predict_api_host_frames_4096"] = api
    del model

    # ---------------------------------------------------------------- C2 train-step variants at batch 256
    B = 256
    fr, lab = synth.synth_frames(B, seed=1)
    x = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
    y = torch.from_numpy(synth.one_hot(lab)).cuda()
    variants = {"all_layers_adam": (set(names), "adam", 3e-4, TRAIN_GFLOP),
                "reference_phase2_rmsprop": (set(names[-3:]), "rmsprop", 9.3e-6, PHASE2_GFLOP),
                "reference_phase1_adam_head_only": ({"output"}, "adam", 3e-4, FWD_GFLOP)}
    out["train_step_b256"] = {}
    for k, (tr, opt, lr, gf) in variants.items():
        fn = lambda: eng.train_step(x, y, tr, opt, lr)   # noqa: E731
        for _ in range(3):
            fn()
        ms = timed(fn, 20, False)
        out["train_step_b256"][k] = {"ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3, 1),
                                     "tflops": round(B * gf / ms, 1), "gflop_per_frame": gf}

    # ---------------------------------------------------------------- C5 Grad-CAM, 512-frame clip
    n = 512
    fr, _ = synth.synth_frames(n, seed=5)
    xg = torch.from_numpy(synth.vgg16_preprocess(fr)).cuda()
    eng.reserve(n, False)
    fn = lambda: eng.gradcam_batch(xg)   # noqa: E731
    for _ in range(3):
        fn()
    ms_total = timed(fn, 10, False)
    _native.timing = []
    fn()
    torch.cuda.synchronize()
    recs, _native.timing = _native.timing, None
    ms_kernel = sum(e0.elapsed_time(e1) for nme, a, e0, e1 in recs if nme == "abl_gradcam")
    bytes_kernel = n * (32 * 32 * 256 * 2 + 1024 + 128 * 128 * 4)
    out["gradcam_512_frames"] = {"ms_forward_plus_heatmaps": round(ms_total, 4),
                                 "frames_per_s": round(n / ms_total * 1e3, 1),
                                 "heatmap_kernel_ms": round(ms_kernel, 4),
                                 "heatmap_kernel_gbs": round(bytes_kernel / ms_kernel / 1e6, 1),
                                 "heatmap_kernel_frac_of_hbm_peak": round(bytes_kernel / ms_kernel / 1e6 / PEAKS["hbm_gbs"], 3),
                                 "out": "128x128 fp32 per frame (bilinear x4 of the 32x32 map)"}

    # ---------------------------------------------------------------- HBM-bound kernels of the train step (B = 256)
    tr = set(names)
    eng.train_step(x, y, tr)
    _native.timing = []
    eng.train_step(x, y, tr)
    torch.cuda.synchronize()
    recs, _native.timing = _native.timing, None
    P = eng.n_flat
    alg = {   # algorithmic bytes per launch: reads + writes, bf16 activations
        "abl_conv_first_fwd": B * (128 * 128 * 3 * 4 + 128 * 128 * 64 * 2),
        "abl_conv_first_wgrad": B * (128 * 128 * 3 * 4 + 128 * 128 * 64 * 2),
        "abl_gap_fwd": B * 32 * 32 * 256 * 2,
        "abl_gap_bwd_relu": B * 32 * 32 * 256 * 2 * 2,
        "abl_adam_flat": P * 28,
        "abl_prep_conv_weights_multi": (P - 1792 - 514) * (4 + 2 + 2),
    }
    hb = {}
    pools = iter([B * 64 * 64 * 128 * 2 * (2 + 0.25 + 0.25), B * 128 * 128 * 64 * 2 * (2 + 0.25 + 0.25)])
    for nme, a, e0, e1 in recs:
        ms = e0.elapsed_time(e1)
        by = alg.get(nme)
        if nme == "abl_maxpool_bwd_relu":
            by = next(pools)
        if by is None:
            continue
        key = nme if nme != "abl_maxpool_bwd_relu" else f"{nme}[{a[7]}x{a[8]}x{a[9]}]"
        hb[key] = {"ms": round(ms, 4), "alg_mbytes": round(by / 1e6, 1), "gbs": round(by / ms / 1e6, 1),
                   "frac_of_hbm_peak": round(by / ms / 1e6 / PEAKS["hbm_gbs"], 3)}
    out["hbm_bound_kernels_b256"] = hb
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Device engine for the cutoffvgg16 hot path: owns the HBM layout and sequences the sm_100a kernels.

Replaces what Keras does under ``CutoffVGG16.fit / predict`` (reference src/models/models.py:407-444):
forward (7 conv + 2 pool + GAP/Dropout/Dense/softmax), loss, backward for the trainable suffix of
layers, Adam / RMSprop, Grad-CAM.  torch is used for device memory, streams and (for data parallel)
``torch.distributed``; every numeric op is a kernel of libabl_b200.so.

HBM layout (per process / GPU)
  theta, grad, m, v, rms : flat fp32 [P] in Keras order and layout (HWIO kernels, (in,out) dense),
                           each tensor padded to 4 floats; gradients of all layers land in ONE buffer
                           so data-parallel all-reduce buckets are plain slices (no pack copy).
  w_fwd[l], w_dgrad[l]   : bf16 GEMM operands [9][Cout][Cin] / [9][Cin][Cout] re-cast after each update.
  act[l] / pooled[l]     : bf16 NHWC activations (full resolution kept only when training).
"""
import math
import os

import numpy as np
import torch

from . import ops
from .data import note_async_copy

# Keras VGG16(include_top=False).layers[1:]: (name, kind, out_channels)  (third-party layer list,
# sliced by the reference at models.py:410)
VGG16_LAYERS = [
    ("block1_conv1", "conv", 64), ("block1_conv2", "conv", 64), ("block1_pool", "pool", 0),
    ("block2_conv1", "conv", 128), ("block2_conv2", "conv", 128), ("block2_pool", "pool", 0),
    ("block3_conv1", "conv", 256), ("block3_conv2", "conv", 256), ("block3_conv3", "conv", 256),
    ("block3_pool", "pool", 0),
    ("block4_conv1", "conv", 512), ("block4_conv2", "conv", 512), ("block4_conv3", "conv", 512),
    ("block4_pool", "pool", 0),
    ("block5_conv1", "conv", 512), ("block5_conv2", "conv", 512), ("block5_conv3", "conv", 512),
    ("block5_pool", "pool", 0),
]


class ConvSpec:
    __slots__ = ("name", "cin", "cout", "h", "w", "pool", "vgg_index")

    def __init__(self, name, cin, cout, h, w, vgg_index):
        self.name, self.cin, self.cout, self.h, self.w = name, cin, cout, h, w
        self.pool = False
        self.vgg_index = vgg_index   # index in vgg16.layers[1:cutoff] (what FINETUNE_LAYER slices)


def build_plan(input_shape, cutoff_layer, check_tiling=False):
    """Conv specs for ``vgg16.layers[1:cutoff_layer]`` (models.py:410); a pool folds into the conv above it.
    check_tiling: kept for callers of round 1; the device kernels now take any feature-map size (frames that are not a
    whole number of 8 x 16-pixel tiles run with zero-padded partial tiles), so nothing extra is checked."""
    h, w, c = input_shape
    if c != 3:
        raise NotImplementedError("DATA.IMG_DIM must give 3-channel frames (VGG16 input)")
    kept = VGG16_LAYERS[: cutoff_layer - 1]
    if not kept or kept[0][1] != "conv":
        raise NotImplementedError("HPARAMS.CUTOFFVGG16.CUTOFF_LAYER must keep at least block1_conv1")
    specs = []
    for i, (name, kind, cout) in enumerate(kept):
        if kind == "conv":
            specs.append(ConvSpec(name, c, cout, h, w, i))
            c = cout
        else:
            if h < 2 or w < 2:
                raise NotImplementedError(f"DATA.IMG_DIM is too small for {name} (its input would be {h}x{w})")
            specs[-1].pool = True
            h, w = h // 2, w // 2          # Keras MaxPooling2D, padding 'valid': an odd side drops its last row / column
    return specs, (h, w, c)


class CutoffVGG16Engine:
    def __init__(self, input_shape=(128, 128, 3), n_classes=2, cutoff_layer=10, dropout=0.45, output_bias=None,
                 seed=10001, device=None, rank=0, world_size=1, process_group=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ab_line_classifier_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        ops._native.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("ab_line_classifier_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if self.device.index != torch.cuda.current_device():
            # the launch shims use the CURRENT device's stream and per-device kernel attributes: an engine on another
            # device would launch on the wrong stream (one process per GPU is the intended layout)
            raise RuntimeError(f"make {self.device} the current device first (torch.cuda.set_device({self.device.index})): "
                               "the engine launches on the current device's stream")
        self.input_shape = tuple(int(v) for v in input_shape)
        self.n_classes = int(n_classes)
        if not 1 <= self.n_classes <= 8:
            raise NotImplementedError("TRAIN.N_CLASSES must be in 1..8")
        self.cutoff_layer = int(cutoff_layer)
        self.dropout = float(dropout)
        self.seed, self.rank, self.world_size, self.pg = int(seed), int(rank), int(world_size), process_group
        self.specs, self.feat_shape = build_plan(self.input_shape, self.cutoff_layer, check_tiling=True)
        if self.specs[0].cout != 64:
            raise NotImplementedError("first layer must be block1_conv1 (64 filters)")
        # data parallel: fused reduce-scatter + optimizer + all-gather over NVLink peer memory (default) or NCCL
        # all-reduce + local optimizer (ABL_DP_FUSED=0, and always for apply_update=False gradient inspection)
        import os
        self.fused_dp = os.environ.get("ABL_DP_FUSED", "1") != "0"
        # training keeps 2-bit max-pool routing codes instead of the pre-pool activations (ABL_POOL_CODES=0: keep them)
        self.pool_codes = os.environ.get("ABL_POOL_CODES", "1") != "0"
        self.fuse_unpool = os.environ.get("ABL_FUSE_UNPOOL", "1") != "0"    # MaxPoolGrad inside the dgrad epilogue
        self.fuse_gap = os.environ.get("ABL_FUSE_GAP", "1") != "0"          # GAP partial sums inside the last conv
        # training keeps a 1-bit ReLU mask of the last conv's activation instead of the activation (needs fuse_gap)
        self.relu_bits = os.environ.get("ABL_RELU_BITS", "1") != "0"
        self._gap_parts = 0
        self.step_count = 0           # dropout counter
        self.opt_steps = {}           # optimizer name -> steps taken (Adam bias correction)
        self._alloc_params(output_bias)
        self._cap = 0
        self._cap_train = False
        self._keep_last_act = False     # Grad-CAM with a cutoff that ends in a pool: keep the last conv's pre-pool output
        self._buf = {}
        self._graphs = {}             # (batch, uint8 input, buffer generation) -> captured forward CUDA graph
        self._tgraphs, self._tg_seen, self._alpha_tabs = {}, {}, {}      # whole-train-step graphs (fit_step)
        self._step_dev = self._t_dev = None
        self._dev_counters = None
        self.train_graphs = os.environ.get("ABL_TRAIN_GRAPHS", "1") != "0"
        self.dp_graphs = os.environ.get("ABL_DP_GRAPHS", "1") != "0"      # whole-step graphs under data parallel
        # data parallel: all-reduce gradient buckets while backward still runs (True) or once after it (False).
        # Measured at 4 x B200, batch 256/GPU: 4.91 ms/step overlapped vs 4.68 ms after backward -- the persistent
        # conv kernels own every SM (1 CTA of ~220 KB shared memory each, static tile assignment), so a CTA whose SM
        # is held by a concurrent NCCL kernel starts late and stretches the whole launch; the 6.9 MB all-reduce
        # itself takes < 0.1 ms.  Default: after backward.
        import os
        self.dp_overlap = os.environ.get("ABL_DP_OVERLAP", "0") == "1"
        # input ring: two (frames, labels) slots filled on a copy stream so the H2D copy of batch i+1
        # overlaps the kernels of batch i
        self._slot = 0
        self._slot_u8 = [False, False]
        self._raw = [None, None]             # device staging of raw frames that still need the resize
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._slot_ready = [None, None]      # event: H2D into the slot finished (recorded on the copy stream)
        self._slot_free = [None, None]       # event: last kernel reading the slot finished (compute stream)

    # ------------------------------------------------------------------ parameters
    def _alloc_params(self, output_bias):
        self.segments = {}     # Keras weight name -> (offset, numel, shape)
        off = 0

        def add(name, shape):
            nonlocal off
            n = int(np.prod(shape))
            self.segments[name] = (off, n, tuple(shape))
            off += (n + 3) // 4 * 4

        self.layer_ranges = {}     # layer name -> (start, end) in the flat buffer
        for s in self.specs:
            start = off
            add(s.name + "/kernel", (3, 3, s.cin, s.cout))
            add(s.name + "/bias", (s.cout,))
            self.layer_ranges[s.name] = (start, off)
        start = off
        add("output/kernel", (self.feat_shape[2], self.n_classes))
        add("output/bias", (self.n_classes,))
        self.layer_ranges["output"] = (start, off)
        self.n_flat = off
        self.n_params = sum(n for _, n, _ in self.segments.values())
        dev = self.device
        self._metrics = None       # device-side Keras metric counters (metrics_begin)
        self._symm = None          # symmetric-memory handle (peer pointers + device barrier) when DP is fused
        self._grad_unreduced = False
        if self.world_size > 1 and self.fused_dp:
            try:
                self._alloc_symmetric(off)
            except Exception as e:  # noqa: BLE001 -- e.g. a driver / container without peer-memory support
                import sys
                print(f"[ab_line_classifier_b200] symmetric memory unavailable ({type(e).__name__}: {e}); "
                      "data parallel falls back to NCCL all-reduce + local optimizer", file=sys.stderr)
                self._symm = None
        if self._symm is None:
            self.theta = torch.zeros(off, device=dev, dtype=torch.float32)
            self.grad = torch.zeros(off, device=dev, dtype=torch.float32)
        self.opt_state = {}
        self.w_fwd, self.w_dgrad = {}, {}
        for s in self.specs[1:]:
            self.w_fwd[s.name] = torch.empty(9, s.cout, s.cin, device=dev, dtype=torch.bfloat16)
            self.w_dgrad[s.name] = torch.empty(9, s.cin, s.cout, device=dev, dtype=torch.bfloat16)
        self.class_weight = torch.ones(self.n_classes, device=dev, dtype=torch.float32)
        self.init_glorot(self.seed, output_bias)

    def _alloc_symmetric(self, n_flat):
        """theta and grad in ONE symmetric allocation: after the rendezvous every rank holds a peer pointer to every
        rank's buffers (NVLink P2P) and a device-side barrier; abl_opt_flat_dp reads / writes through them."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        if self.world_size > 8 or not dist.is_initialized():
            return
        group = self.pg if self.pg is not None else dist.group.WORLD
        buf = symm.empty(2 * n_flat, dtype=torch.float32, device=self.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        self._symm_buf, self._symm = buf, hdl
        self.theta, self.grad = buf[:n_flat], buf[n_flat:]
        self._theta_peers = [int(p) for p in hdl.buffer_ptrs]
        self._grad_peers = [int(p) + 4 * n_flat for p in hdl.buffer_ptrs]
        self._grad_unreduced = False
        torch.cuda.synchronize()
        hdl.barrier()

    def layer_names(self):
        return [s.name for s in self.specs] + ["output"]

    def view(self, name, buf=None):
        off, n, shape = self.segments[name]
        return (self.theta if buf is None else buf)[off:off + n].view(shape)

    def init_glorot(self, seed, output_bias=None):
        """Keras default init (Glorot uniform kernels, zero biases), as ``VGG16(weights=None)`` + Dense
        would give; the reference's ImageNet download (models.py:409) needs network access."""
        rng = np.random.default_rng(seed)
        w = {}
        for s in self.specs:
            lim = math.sqrt(6.0 / (9 * s.cin + 9 * s.cout))
            w[s.name + "/kernel"] = rng.uniform(-lim, lim, size=(3, 3, s.cin, s.cout)).astype(np.float32)
            w[s.name + "/bias"] = np.zeros(s.cout, np.float32)
        c = self.feat_shape[2]
        lim = math.sqrt(6.0 / (c + self.n_classes))
        w["output/kernel"] = rng.uniform(-lim, lim, size=(c, self.n_classes)).astype(np.float32)
        w["output/bias"] = (np.zeros(self.n_classes, np.float32) if output_bias is None
                            else np.asarray(output_bias, np.float32).reshape(self.n_classes))
        self.set_weights(w)

    def set_weights(self, weights, strict=True):
        """weights: dict Keras name ('block1_conv1/kernel', ..., 'output/bias') -> array in Keras layout."""
        host = self.theta.cpu().numpy()
        for name, (off, n, shape) in self.segments.items():
            key = name if name in weights else name + ":0"
            if key not in weights:
                if strict:
                    raise KeyError(f"missing weight {name}")
                continue
            a = np.asarray(weights[key], np.float32)
            if a.shape != shape:
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            host[off:off + n] = a.reshape(-1)
        self.theta.copy_(torch.from_numpy(host))
        self.refresh_operands()

    def get_weights(self):
        host = self.theta.cpu().numpy()
        return {name: host[off:off + n].reshape(shape).copy() for name, (off, n, shape) in self.segments.items()}

    def refresh_operands(self, layers=None):
        """Re-cast fp32 master kernels into the bf16 GEMM operand layouts."""
        todo = [(self.view(s.name + "/kernel"), self.w_fwd[s.name], self.w_dgrad[s.name])
                for s in self.specs[1:] if layers is None or s.name in layers]
        if todo:
            ops.prep_conv_weights_multi(todo)       # one launch for all layers

    # ------------------------------------------------------------------ buffers
    def _ensure(self, batch, training):
        if batch <= self._cap and (self._cap_train or not training):
            return
        cap = max(batch, self._cap)
        training = training or self._cap_train
        dev, bf = self.device, torch.bfloat16
        b = {}
        H, W, _ = self.input_shape
        b["x"] = [torch.empty(cap, H, W, 3, device=dev, dtype=torch.float32) for _ in range(2)]
        b["x8"] = [torch.empty(cap, H, W, 3, device=dev, dtype=torch.uint8) for _ in range(2)]
        max_full = 0
        for s in self.specs:
            if s.pool:
                b["pool_" + s.name] = torch.empty(cap, s.h // 2, s.w // 2, s.cout, device=dev, dtype=bf)
                if training:
                    if self.pool_codes:
                        # 2-bit routing codes of the fused max-pool (1/8 of the pooled bytes) replace the
                        # full-resolution activation that MaxPoolGrad would otherwise have to keep and re-read
                        b["code_" + s.name] = torch.empty(cap, s.h // 2, s.w // 2, s.cout // 8, device=dev,
                                                          dtype=torch.int16)
                    else:
                        b["act_" + s.name] = torch.empty(cap, s.h, s.w, s.cout, device=dev, dtype=bf)
            else:
                b["act_" + s.name] = torch.empty(cap, s.h, s.w, s.cout, device=dev, dtype=bf)
            max_full = max(max_full, s.h * s.w * s.cout)
        C = self.feat_shape[2]
        n = self.n_classes
        b["gap"] = torch.empty(cap, C, device=dev, dtype=torch.float32)
        last = self.specs[-1]
        if last.pool and self._keep_last_act:
            b["gcam_act"] = torch.empty(cap, last.h, last.w, last.cout, device=dev, dtype=bf)
        if self.fuse_gap and not last.pool and ops.igemm_gap_supported(last.h, last.w, last.cout):
            self._gap_parts = ops.igemm_gap_parts(last.h, last.w, last.cout)
            b["gap_part"] = torch.empty(cap * self._gap_parts, C, device=dev, dtype=torch.float32)
            if training and self.relu_bits and last.cout % 256 == 0:       # (the N = 256 kernel writes the bits)
                # after the GAP sums only the SIGN of the last activation is used (ReluGrad): the forward writes one bit
                # per element here and never the activation; abl_gap_bwd_bits reads it back
                b["bits_last"] = torch.empty(cap, last.h, last.w, last.cout // 32, device=dev, dtype=torch.int32)
        b["logits"] = torch.empty(cap, n, device=dev, dtype=torch.float32)
        b["probs"] = torch.empty(cap, n, device=dev, dtype=torch.float32)
        b["y"] = [torch.empty(cap, n, device=dev, dtype=torch.float32) for _ in range(2)]
        b["loss"] = torch.zeros(1, device=dev, dtype=torch.float32)
        b["loss_ws"] = ops.head_loss_ws(cap, dev)       # per-block loss sums + ticket (deterministic batch loss)
        if training:
            # per-CTA partial sums of the fused bias gradients: written by the kernel that produces a layer's dZ, added
            # in a fixed order by that layer's wgrad reduction (no float atomics anywhere: bit-reproducible training)
            b["bpart"] = ops.parts_buffer(max(s.cout for s in self.specs), dev)
            b["first_ws"] = torch.empty(ops.conv_first_wgrad_workspace_bytes(), device=dev, dtype=torch.uint8)
            b["head_scratch"] = torch.zeros(C * n + 4 + n, device=dev, dtype=torch.float32)
            b["dlogits"] = torch.empty(cap, n, device=dev, dtype=torch.float32)
            b["keep"] = torch.empty(cap, C, device=dev, dtype=torch.uint8)
            b["dgap"] = torch.empty(cap, C, device=dev, dtype=torch.float32)
            b["dz_a"] = torch.empty(cap * max_full, device=dev, dtype=bf)
            b["dz_b"] = torch.empty(cap * max_full, device=dev, dtype=bf)
            b["dp"] = torch.empty(cap * max_full // 4, device=dev, dtype=bf)
            ws = 0
            for s in self.specs[1:]:
                ws = max(ws, ops.wgrad_workspace_bytes(cap, s.h, s.w, s.cin, s.cout))
            b["ws"] = torch.empty(ws, device=dev, dtype=torch.uint8)
        self._buf, self._cap, self._cap_train = b, cap, training
        self._slot_ready, self._slot_free = [None, None], [None, None]

    def _x(self, B):
        """Current slot's frames: fp32 preprocessed, or raw uint8 RGB (preprocessing fused into block1_conv1)."""
        return (self._buf["x8"] if self._slot_u8[self._slot] else self._buf["x"])[self._slot][:B]

    def _in_act(self, li, B):
        """Input activation of conv li (bf16), li >= 1."""
        prev = self.specs[li - 1]
        key = ("pool_" if prev.pool else "act_") + prev.name
        return self._buf[key][:B]

    # ------------------------------------------------------------------ forward
    def forward_device(self, B, training=False, keep_mask_in=False, y=None, inv_batch=None, loss_out=None,
                       need_last_act=True, step_dev=None):
        """Runs the network on self._buf['x'][:B] (already resident).  Returns probs view [B,n].
        training: dropout on, class-weighted loss + dlogits.  Otherwise, if y is given, the UNweighted
        loss sum_i l_i * inv_batch is accumulated into loss_out (Keras evaluate / validation)."""
        b = self._buf
        for li, s in enumerate(self.specs):
            full = b.get("act_" + s.name)
            full = full[:B] if full is not None else None
            pooled = b["pool_" + s.name][:B] if s.pool else None
            if li == 0:
                ops.conv_first_fwd(self._x(B), self.view(s.name + "/kernel"), self.view(s.name + "/bias"), full)
                if s.pool:
                    raise NotImplementedError("a pool directly after block1_conv1 is not a VGG16 prefix")
            elif li == len(self.specs) - 1 and "gap_part" in b:
                # last conv: GlobalAveragePooling partial sums come out of the epilogue; without Grad-CAM / training
                # (need_last_act False) the activation itself is never written
                bits = b["bits_last"][:B] if (training and "bits_last" in b) else None
                ops.conv3x3_igemm_gap(self._in_act(li, B), self.w_fwd[s.name], self.view(s.name + "/bias"),
                                      full if ((bits is None) if training else need_last_act) else None,
                                      b["gap_part"], relu_bits=bits)
            else:
                code = b.get("code_" + s.name)
                if li == len(self.specs) - 1 and s.pool and need_last_act and not training and "gcam_act" in b:
                    full = b["gcam_act"][:B]             # Grad-CAM wants the map BEFORE the trailing pool
                ops.conv3x3_igemm(self._in_act(li, B), self.w_fwd[s.name], self.view(s.name + "/bias"), None,
                                  full, pooled, ops.EPI_BIAS_RELU, pool_code=code[:B] if code is not None else None)
        last = self.specs[-1]
        gp = {}
        if "gap_part" in b:
            # the per-rectangle sums are finished inside the head kernel (it writes b["gap"] for the backward pass)
            gp = dict(gap_part=b["gap_part"], gap_parts=self._gap_parts, gap_hw=last.h * last.w)
        else:
            feat = b["pool_" + last.name][:B] if last.pool else b["act_" + last.name][:B]
            ops.gap_fwd(feat, b["gap"][:B])
        # the step's loss OVERWRITES its destination in training (one deterministic sum per batch); evaluation
        # accumulates batch after batch into the caller's scalar (loss_out given) or overwrites the engine's own
        accumulate = loss_out is not None and not training
        if loss_out is None:
            loss_out = b["loss"]
        if training:
            ops.head_fwd(b["gap"][:B], self.view("output/kernel"), self.view("output/bias"), b["logits"][:B],
                         b["probs"][:B], y, self.class_weight, inv_batch, loss_out, b["dlogits"][:B],
                         b["keep"][:B], keep_mask_in, True, self.dropout, self.seed, self.step_count, self.rank,
                         step_dev=step_dev, loss_ws=b["loss_ws"], loss_accumulate=False, **gp)
        elif y is not None:
            ops.head_fwd(b["gap"][:B], self.view("output/kernel"), self.view("output/bias"), b["logits"][:B],
                         b["probs"][:B], y, None, inv_batch, loss_out, loss_ws=b["loss_ws"], loss_accumulate=accumulate,
                         **gp)
        else:
            ops.head_fwd(b["gap"][:B], self.view("output/kernel"), self.view("output/bias"), b["logits"][:B],
                         b["probs"][:B], **gp)
        return b["probs"][:B]

    def reserve(self, batch, training=False):
        """Allocate (or grow) activation / input buffers; call before stage() when the batch size or the
        training mode is about to grow, since growing re-allocates the input slots."""
        self._ensure(int(batch), bool(training))

    @staticmethod
    def _as_tensor(a, dtype=None):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        if dtype is not None and a.dtype != dtype:
            a = a.to(dtype)
        return a

    def stage(self, x, y=None, slot=None):
        """Asynchronously copy a batch into an input slot on the copy stream and return the slot index.
        x: fp32 [B,H,W,3] frames as ``vgg16_preprocess`` leaves them, or RAW uint8 RGB frames (then the
        preprocessing runs inside block1_conv1).  Host tensors should be pinned for a truly async copy."""
        x = self._as_tensor(x)
        if x.dtype != torch.uint8:
            x = self._as_tensor(x, torch.float32)
        B = x.shape[0]
        u8 = x.dtype == torch.uint8
        resize = u8 and x.dim() == 4 and x.shape[3] == 3 and tuple(x.shape[1:]) != self.input_shape
        if tuple(x.shape[1:]) != self.input_shape and not resize:
            raise ValueError(f"expected frames of shape {self.input_shape}, got {tuple(x.shape[1:])} "
                             "(only raw uint8 RGB frames are resized on the device)")
        self._ensure(B, self._cap_train)
        slot = (self._slot ^ 1) if slot is None else slot
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_stream(cur) if self._slot_free[slot] is None else self._copy_stream.wait_event(
                self._slot_free[slot])
            if x.is_cuda:            # device-resident dataset: the batch was gathered on the caller's stream just now
                self._copy_stream.wait_stream(cur)
            if resize:
                # decoded frames at their native size: nearest-neighbour resize to IMG_DIM on the device, as
                # Preprocessor._parse_fn does on the host (preprocessor.py:76-80); the copy stream runs it
                raw = self._raw[slot]
                if raw is None or raw.numel() < x.numel():
                    raw = self._raw[slot] = torch.empty(x.numel(), device=self.device, dtype=torch.uint8)
                rv = raw[: x.numel()].view(x.shape)
                rv.copy_(x, non_blocking=True)
                ops.resize_nearest_u8(rv, self._buf["x8"][slot][:B])
            else:
                (self._buf["x8"] if u8 else self._buf["x"])[slot][:B].copy_(x, non_blocking=True)
            if y is not None:
                self._buf["y"][slot][:B].copy_(self._as_tensor(y, torch.float32), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        for t in (x, y):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(self._copy_stream)   # a fresh gather / augmentation output is read on the copy stream
            else:
                note_async_copy(t, ev)               # a pinned batch buffer must not be refilled before this copy ran
        self._slot_u8[slot] = u8
        self._slot_ready[slot] = ev
        return slot

    def use_slot(self, slot):
        """Make ``slot`` the current input: the compute stream waits for its H2D copy."""
        self._slot = slot
        if self._slot_ready[slot] is not None:
            torch.cuda.current_stream().wait_event(self._slot_ready[slot])

    def release_slot(self):
        """Record that every kernel reading the current slot has been enqueued (the slot may be refilled after it)."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._slot_free[self._slot] = ev

    def load_input(self, x, y=None):
        """Synchronous-order variant: stage into the other slot and switch to it.  Returns the batch size."""
        self.use_slot(self.stage(x, y))
        return x.shape[0]

    GRAPH_MAX_BATCH = 64     # below this the forward pass is launch-bound (9 kernels of a few microseconds each)

    def predict_batch(self, x, use_graph=None, need_last_act=False):
        """x: [B,H,W,3] fp32 (BGR, mean-subtracted, as vgg16_preprocess leaves it) or raw uint8 RGB.
        Returns device probs.  Small batches (the reference's batch-1 loops: ``clock_avg_runtime`` predict.py:425-448,
        ``WandbGradcamEvalCallback._inference`` train_utils.py:280-293) replay a captured CUDA graph of the 10
        forward launches instead of launching them one by one."""
        B = x.shape[0]
        self._ensure(B, False)
        if use_graph is None:
            use_graph = B <= self.GRAPH_MAX_BATCH
        if not use_graph:
            self.load_input(x)
            probs = self.forward_device(B, need_last_act=need_last_act)
            self.release_slot()
            return probs
        u8 = (x.dtype == torch.uint8) if torch.is_tensor(x) else (np.asarray(x).dtype == np.uint8)
        g = self._forward_graph(B, u8, need_last_act)
        self.use_slot(self.stage(x, slot=0))
        g.replay()
        self.release_slot()
        return self._buf["probs"][:B]

    def _forward_graph(self, B, u8, need_last_act=False):
        """CUDA graph of forward_device(B) reading input slot 0 (captured once per (batch, input dtype, buffers))."""
        key = (B, bool(u8), id(self._buf), bool(need_last_act))
        g = self._graphs.get(key)
        if g is not None:
            return g
        self._slot, self._slot_u8[0] = 0, bool(u8)
        for _ in range(2):                       # warm-up outside capture: one-time kernel attributes, lazy module load
            self.forward_device(B, need_last_act=need_last_act)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward_device(B, need_last_act=need_last_act)
        self._graphs = {k: v for k, v in self._graphs.items() if k[2] == id(self._buf)}   # drop graphs of old buffers
        self._graphs[key] = g
        return g

    # ------------------------------------------------------------------ backward + update
    def trainable_suffix(self, trainable):
        names = self.layer_names()
        tr = [n in trainable for n in names]
        if not any(tr):
            raise ValueError("no trainable layer")
        return names, tr, tr.index(True)

    def backward_device(self, B, trainable, reduce_hook=None):
        """Gradients of the loss for the trainable layers into self.grad (zeroed first).
        reduce_hook(start, end) is called as soon as grad[start:end] is final (DP buckets)."""
        b = self._buf
        names, tr, lowest = self.trainable_suffix(trainable)
        # no zero-fill: every kernel below OVERWRITES its slice of self.grad (fixed-order sums, no atomics)
        keep = b["keep"][:B]
        if tr[-1]:
            dwk, dbk = self.view("output/kernel", self.grad), self.view("output/bias", self.grad)
        else:   # still need dgap; dump head grads in scratch
            nk = self.feat_shape[2] * self.n_classes
            dwk, dbk = b["head_scratch"][:nk], b["head_scratch"][(nk + 3) // 4 * 4:][: self.n_classes]
        ops.head_bwd(b["gap"][:B], self.view("output/kernel"), b["dlogits"][:B], keep, self.dropout, dwk, dbk,
                     b["dgap"][:B])
        if reduce_hook and tr[-1]:
            reduce_hook(*self.layer_ranges["output"])
        nconv = len(self.specs)
        if lowest >= nconv:
            return
        last = self.specs[-1]
        cur, other = b["dz_a"], b["dz_b"]
        dz = cur[: B * last.h * last.w * last.cout].view(B, last.h, last.w, last.cout)
        # BiasAddGrad of a layer is fused into whichever kernel PRODUCES that layer's dZ (GAP backward, max-pool
        # backward, or the dgrad epilogue of the layer above): it leaves per-CTA partial sums in b["bpart"], and the
        # layer's own wgrad reduction adds them in a fixed order -- dZ is never re-read for it.  Exceptions:
        # block1_conv1's bias gradient comes out of its own wgrad GEMM (ones column), and so does that of the
        # Cin == 64 layers, whose wgrad GEMM has idle accumulator rows (wgrad64_tc_kernel).
        def wgrad_has_bias(i):
            return i > 0 and self.specs[i].cin == 64 and ops.wgrad_fuses_bias(self.specs[i].cin)

        def bias_parts_of(i):
            return b["bpart"] if (tr[i] and i > 0 and not wgrad_has_bias(i)) else None

        if last.pool:
            # CUTOFF_LAYER ends in a pool (4, 7, 11 ...): the average runs over the POOLED map.  Its gradient, masked by
            # relu'(pooled) (pooled = max of post-ReLU values: zero iff the whole window is), is routed back through the
            # forward's codes; the bias gradient of the conv rides along as usual.
            ph, pw = last.h // 2, last.w // 2
            dp = b["dp"][: B * ph * pw * last.cout].view(B, ph, pw, last.cout)
            ops.gap_bwd_relu(b["dgap"][:B], b["pool_" + last.name][:B], dp)
            code = b.get("code_" + last.name)
            if code is not None:
                ops.maxpool_bwd_code(dp, code[:B], dz, db_parts=bias_parts_of(nconv - 1))
            else:
                ops.maxpool_bwd_relu(dp, b["act_" + last.name][:B], b["pool_" + last.name][:B], dz,
                                     db_parts=bias_parts_of(nconv - 1))
        elif "bits_last" in b:
            ops.gap_bwd_bits(b["dgap"][:B], b["bits_last"][:B], dz, db_parts=bias_parts_of(nconv - 1))
        else:
            ops.gap_bwd_relu(b["dgap"][:B], b["act_" + last.name][:B], dz, db_parts=bias_parts_of(nconv - 1))
        for li in range(nconv - 1, lowest - 1, -1):
            s = self.specs[li]
            if tr[li]:
                dw = self.view(s.name + "/kernel", self.grad)
                if li == 0:
                    ops.conv_first_wgrad(self._x(B), dz, dw, self.view(s.name + "/bias", self.grad), b["first_ws"])
                else:
                    ops.conv3x3_wgrad(self._in_act(li, B), dz, dw, self.view(s.name + "/bias", self.grad), b["ws"],
                                      accumulate=False, db_parts=bias_parts_of(li))
                if reduce_hook:
                    reduce_hook(*self.layer_ranges[s.name])
            if li > lowest:
                prev = self.specs[li - 1]
                if prev.pool:
                    dp = b["dp"][: B * s.h * s.w * s.cin].view(B, s.h, s.w, s.cin)
                    ndz = other[: B * prev.h * prev.w * prev.cout].view(B, prev.h, prev.w, prev.cout)
                    code = b.get("code_" + prev.name)
                    # (an odd side before the pool: the un-pooled tensor has a row / column no window covers -- the
                    # two-kernel path below zero-fills it)
                    even = prev.h == 2 * s.h and prev.w == 2 * s.w
                    if code is not None and self.fuse_unpool and even and ops.dgrad_unpool_supported(s.h, s.w, s.cin):
                        # dgrad + relu'(pooled) + MaxPoolGrad (+ BiasAddGrad) in one kernel
                        ops.conv3x3_dgrad_unpool(dz, self.w_dgrad[s.name], b["pool_" + prev.name][:B], code[:B], ndz,
                                                 db_parts=bias_parts_of(li - 1))
                    elif code is not None:
                        # relu'(pooled) in the dgrad epilogue, then routing by the forward's 2-bit codes
                        ops.conv3x3_igemm(dz, self.w_dgrad[s.name], None, b["pool_" + prev.name][:B], dp, None,
                                          ops.EPI_RELU_MASK)
                        ops.maxpool_bwd_code(dp, code[:B], ndz, db_parts=bias_parts_of(li - 1))
                    else:
                        ops.conv3x3_igemm(dz, self.w_dgrad[s.name], None, None, dp, None, 0)
                        ops.maxpool_bwd_relu(dp, b["act_" + prev.name][:B], b["pool_" + prev.name][:B], ndz,
                                             db_parts=bias_parts_of(li - 1))
                else:
                    ndz = other[: B * prev.h * prev.w * prev.cout].view(B, prev.h, prev.w, prev.cout)
                    ops.conv3x3_igemm(dz, self.w_dgrad[s.name], None, b["act_" + prev.name][:B], ndz, None,
                                      ops.EPI_RELU_MASK, colsum_parts=bias_parts_of(li - 1))
                dz = ndz
                cur, other = other, cur

    def _runs(self, trainable):
        """Contiguous [start,end) runs of the flat buffer covering the trainable layers."""
        runs = []
        for n in self.layer_names():
            if n in trainable:
                s, e = self.layer_ranges[n]
                if runs and runs[-1][1] == s:
                    runs[-1][1] = e
                else:
                    runs.append([s, e])
        return runs

    def apply_optimizer(self, kind, lr, trainable, beta_1=0.9, beta_2=0.999, rho=0.9, eps=1e-7, alpha_tab=None):
        """Keras 2.9 Adam / RMSprop on the trainable slices (Appendix A.5), then re-cast bf16 operands.
        alpha_tab (CUDA-graph capture): Adam's alpha_t comes from this device table indexed by the device step count."""
        key = kind
        t = self.opt_steps.get(key, 0) + 1
        if alpha_tab is None:
            self.opt_steps[key] = t          # (graph replays mirror the counter on the host themselves)
        fused = self._symm is not None and self._grad_unreduced
        if fused:
            self._symm.barrier()             # every rank's gradients are final
        if kind == "adam":
            if "m" not in self.opt_state:
                self.opt_state["m"] = torch.zeros(self.n_flat, device=self.device)
                self.opt_state["v"] = torch.zeros(self.n_flat, device=self.device)
            alpha = lr * math.sqrt(1.0 - beta_2 ** t) / (1.0 - beta_1 ** t)
            for s, e in self._runs(trainable):
                if fused:
                    ops.opt_flat_dp("adam", self._theta_peers, self._grad_peers, self.opt_state["m"],
                                    self.opt_state["v"], s, e - s, self.rank, self.world_size, alpha, beta_1, beta_2, eps,
                                    alpha_tab=alpha_tab, t_dev=self._t_dev if alpha_tab is not None else None)
                elif alpha_tab is not None:
                    ops.adam_flat_tab(self.theta[s:e], self.grad[s:e], self.opt_state["m"][s:e], self.opt_state["v"][s:e],
                                      alpha_tab, self._t_dev, beta_1, beta_2, eps)
                else:
                    ops.adam_flat(self.theta[s:e], self.grad[s:e], self.opt_state["m"][s:e], self.opt_state["v"][s:e],
                                  alpha, beta_1, beta_2, eps)
        elif kind == "rmsprop":
            if "rms" not in self.opt_state:
                self.opt_state["rms"] = torch.zeros(self.n_flat, device=self.device)
            for s, e in self._runs(trainable):
                if fused:
                    ops.opt_flat_dp("rmsprop", self._theta_peers, self._grad_peers, self.opt_state["rms"], None, s,
                                    e - s, self.rank, self.world_size, lr, rho, 0.0, eps)
                else:
                    ops.rmsprop_flat(self.theta[s:e], self.grad[s:e], self.opt_state["rms"][s:e], lr, rho, eps)
        else:
            raise ValueError(f"unknown optimizer {kind}")
        if fused:
            self._symm.barrier()             # every shard of the new weights has landed everywhere; grads may be reused
            self._grad_unreduced = False
        self.refresh_operands(trainable)

    def set_class_weight(self, class_weight):
        cw = np.ones(self.n_classes, np.float32)
        if class_weight is not None:
            for k, v in dict(class_weight).items():
                cw[int(k)] = float(v)
        self.class_weight.copy_(torch.from_numpy(cw))

    def eval_batch(self, x, y, loss_out=None, inv_batch=1.0, metrics=False):
        """Inference forward + unweighted categorical cross-entropy sum (Keras evaluate).  Device probs.
        ``metrics``: also add the batch to the device-side metric counters (metrics_begin)."""
        B = x.shape[0]
        self._ensure(B, False)
        self.load_input(x, y)
        probs = self.forward_device(B, y=self._buf["y"][self._slot][:B], inv_batch=inv_batch, loss_out=loss_out,
                                    need_last_act=False)
        if metrics:
            self.metrics_update(B)
        else:
            self.release_slot()
        return probs

    # ---- Keras metric state on the device (the reference's list: src/train.py:109-115)
    def metrics_begin(self, auc_thresholds, class_thresholds):
        """Zero the counters for a new epoch / evaluation.  Thresholds: fp32 arrays from metrics.device_plan()."""
        m = self._metrics
        key = (np.asarray(auc_thresholds, np.float32).tobytes(), np.asarray(class_thresholds, np.float32).tobytes())
        if m is None or m["key"] != key:
            thr = torch.from_numpy(np.ascontiguousarray(auc_thresholds, np.float32)).to(self.device)
            cthr = torch.from_numpy(np.ascontiguousarray(class_thresholds, np.float32)).to(self.device)
            if cthr.numel() != self.n_classes:
                raise ValueError("one class threshold per class expected")
            n = ops.metrics_counter_count(self.n_classes, thr.numel())
            m = self._metrics = {"key": key, "thr": thr, "cthr": cthr,
                                 "counters": torch.zeros(n, device=self.device, dtype=torch.int64)}
        m["counters"].zero_()

    def metrics_update(self, B):
        """Add the batch just run (softmax output and the labels of the current input slot) to the counters, then
        release the slot (its label buffer is read here, after the step's own release)."""
        m = self._metrics
        if m is None:
            raise RuntimeError("metrics_begin() must be called first")
        ops.metrics_accumulate(self._buf["probs"][:B], self._buf["y"][self._slot][:B], m["thr"], m["cthr"], m["counters"])
        self.release_slot()

    def metrics_read(self, reduce=True):
        """The counters as a host int64 array (one small D2H copy; synchronises).  Data parallel: every metric is a
        ratio of counts, so the SUM over ranks gives the metric of all frames seen by all ranks (``reduce``)."""
        c = self._metrics["counters"]
        if reduce and self.world_size > 1:
            import torch.distributed as dist
            c = c.clone()
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.pg)
        return c.cpu().numpy()

    def sum_over_ranks(self, values):
        """Host floats summed over the data-parallel ranks (epoch-end loss bookkeeping); identity on one GPU."""
        if self.world_size == 1:
            return [float(v) for v in values]
        import torch.distributed as dist
        t = torch.tensor([float(v) for v in values], device=self.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg)
        return t.cpu().tolist()

    def train_step(self, x, y, trainable, optimizer="adam", lr=3e-4, global_batch=None, keep_mask=None,
                   apply_update=True, bucket_bytes=2 << 20, loss_out=None, slot=None, batch=None, defer_reduce=False):
        """One step: H2D copy of (x, y) [or a slot pre-filled by stage()], forward (dropout on), loss,
        backward, [DP all-reduce], update.  Returns the device scalar sum_i w_i*l_i / global_batch for THIS
        rank's frames."""
        B = batch if slot is not None else x.shape[0]
        if slot is not None and (B > self._cap or not self._cap_train):
            raise RuntimeError("engine.reserve(batch, training=True) must be called before staging a training batch")
        self._ensure(B, True)
        b = self._buf
        if slot is None:
            self.load_input(x, y)
        else:
            self.use_slot(slot)
        if keep_mask is not None:
            if isinstance(keep_mask, np.ndarray):
                keep_mask = torch.from_numpy(np.ascontiguousarray(keep_mask).astype(np.uint8))
            b["keep"][:B].copy_(keep_mask, non_blocking=True)
        gb = float(global_batch if global_batch is not None else B * self.world_size)
        self.forward_device(B, training=True, keep_mask_in=keep_mask is not None, y=b["y"][self._slot][:B],
                            inv_batch=1.0 / gb, loss_out=loss_out)
        self.step_count += 1
        reducer = None
        # fused DP: leave self.grad un-reduced; apply_optimizer() sums it across ranks inside the optimizer kernel.
        # (apply_update=False without defer_reduce keeps the NCCL all-reduce so callers can inspect the full gradient.)
        fused = self._symm is not None and (apply_update or defer_reduce)
        if fused:
            self._grad_unreduced = True
        if self.world_size > 1 and not fused:
            from .dp import GradBucketReducer
            if not self.dp_overlap:
                bucket_bytes = 1 << 40            # one all-reduce of every trainable slice after backward
            reducer = GradBucketReducer(self.grad, bucket_bytes, self.pg)
        self.backward_device(B, trainable, reduce_hook=reducer.hook if reducer else None)
        if reducer:
            reducer.wait()
        self.release_slot()
        if apply_update:
            self.apply_optimizer(optimizer, lr, trainable)
        return b["loss"] if loss_out is None else loss_out

    # ------------------------------------------------------------------ whole-step CUDA graph (small batches)
    TRAIN_GRAPH_MAX_BATCH = 64    # above this the step is GPU-bound; below, ~0.7 ms of host enqueue time per step dominates
    ALPHA_TABLE_LEN = 32768       # Adam alpha_t for t = 1..T (beyond T it equals lr to fp32 precision)

    def fit_step(self, trainable, optimizer, lr, slot, batch, global_batch, loss_out=None, **kw):
        """What fit() does per batch: train_step on a staged slot + apply_optimizer.  For small batches the ~33
        launches (forward, loss, backward, optimizer, operand re-cast) are captured ONCE per configuration into a CUDA
        graph and replayed; the dropout step and the Adam step count advance in device memory.  Data parallel: the
        captured step contains the two device-side barriers and the fused peer-memory optimizer kernel, so every
        rank replays the same graph (all ranks see the same batch sizes, hence capture on the same step)."""
        trainable = frozenset(trainable)
        eligible = (self.train_graphs and (self.world_size == 1 or (self._symm is not None and self.dp_graphs))
                    and batch <= self.TRAIN_GRAPH_MAX_BATCH and batch <= self._cap and self._cap_train)
        key = (batch, trainable, optimizer, float(lr), tuple(sorted(kw.items())), slot, self._slot_u8[slot],
               id(self._buf), float(global_batch)) if eligible else None
        if eligible:
            self._tg_seen[key] = self._tg_seen.get(key, 0) + 1
        if not eligible or self._tg_seen[key] <= 2:       # the first two steps of a configuration run eagerly
            self.train_step(None, None, trainable, apply_update=False, loss_out=loss_out, global_batch=global_batch,
                            slot=slot, batch=batch, defer_reduce=True)
            self.apply_optimizer(optimizer, lr, trainable, **kw)
            return
        b = self._buf
        if self._step_dev is None:
            # {dropout step, optimizer step}: both advanced on the device by the head kernel's last block
            self._step_dev = torch.zeros(2, device=self.device, dtype=torch.int32)
            self._t_dev = self._step_dev[1:2]
        t_host = self.opt_steps.get(optimizer, 0)
        if self._dev_counters != (self.step_count, optimizer, t_host):      # re-sync after eager steps
            self._step_dev[0:1].fill_(self.step_count)
            self._t_dev.fill_(t_host)
        g = self._tgraphs.get(key)
        self.use_slot(slot)
        if g is None:
            alpha_tab = None
            if optimizer == "adam":
                b1, b2 = kw.get("beta_1", 0.9), kw.get("beta_2", 0.999)
                tk = (float(lr), b1, b2)
                if tk not in self._alpha_tabs:
                    tt = np.arange(1, self.ALPHA_TABLE_LEN + 1, dtype=np.float64)
                    tab = np.array([lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t) for t in tt], np.float32)
                    self._alpha_tabs[tk] = torch.from_numpy(tab).to(self.device)
                alpha_tab = self._alpha_tabs[tk]
            if optimizer == "adam" and "m" not in self.opt_state or optimizer == "rmsprop" and "rms" not in self.opt_state:
                raise RuntimeError("optimizer state must exist before capture")      # two eager steps guarantee it
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                # the head kernel overwrites the loss and advances both device counters: no fill / add kernels
                self.forward_device(batch, training=True, y=b["y"][slot][:batch], inv_batch=1.0 / float(global_batch),
                                    loss_out=b["loss"], step_dev=self._step_dev)
                self.backward_device(batch, trainable)
                if self._symm is not None:
                    self._grad_unreduced = True          # the fused optimizer sums the ranks' gradients itself
                if optimizer == "adam":
                    self.apply_optimizer(optimizer, lr, trainable, alpha_tab=alpha_tab, **kw)
                elif self._symm is not None:
                    t_keep = self.opt_steps.get("rmsprop", 0)
                    self.apply_optimizer(optimizer, lr, trainable, **kw)      # no step-dependent scalar in RMSprop
                    self.opt_steps["rmsprop"] = t_keep                        # (the replay bookkeeping counts the step)
                else:
                    self._apply_rmsprop_captured(lr, trainable, **kw)
            self._tgraphs = {k: v for k, v in self._tgraphs.items() if k[7] == id(self._buf)}
            self._tgraphs[key] = g
            # capture does not execute: fall through to the first replay
        g.replay()
        self.step_count += 1
        self.opt_steps[optimizer] = t_host + 1
        self._dev_counters = (self.step_count, optimizer, t_host + 1)
        self.release_slot()
        if loss_out is not None:
            loss_out.copy_(b["loss"])

    def _apply_rmsprop_captured(self, lr, trainable, rho=0.9, eps=1e-7, **_):
        """RMSprop has no step-dependent scalar: the eager kernels are captured as they are (host counter untouched)."""
        if "rms" not in self.opt_state:
            self.opt_state["rms"] = torch.zeros(self.n_flat, device=self.device)
        for s, e in self._runs(trainable):
            ops.rmsprop_flat(self.theta[s:e], self.grad[s:e], self.opt_state["rms"][s:e], lr, rho, eps)
        self.refresh_operands(trainable)

    # ------------------------------------------------------------------ Grad-CAM
    def gradcam_batch(self, x, out_hw=None):
        """Per-frame Grad-CAM heat-maps (gradcam.py:68-81) -> (device heat [B,oh,ow] fp32, device probs)."""
        last = self.specs[-1]
        if last.pool and not self._keep_last_act:
            # The reference takes the LAST CONV layer's output (gradcam.py:33-37), i.e. the map before a trailing pool.
            # The gradient of the class probability reaches it through MaxPool + GAP: every window hands
            # dp/dgap_c / (H'W') to one of its elements, so the spatial mean is dp/dgap_c / (HW) -- the same alpha_c as
            # without the pool, and the closed form applies to the pre-pool map unchanged.
            self._keep_last_act = True
            self._cap = 0                                   # re-allocate with the extra activation buffer
        probs = self.predict_batch(x, need_last_act=True)
        B = x.shape[0]
        oh, ow = out_hw if out_hw is not None else self.input_shape[:2]
        heat = torch.empty(B, oh, ow, device=self.device, dtype=torch.float32)
        act = self._buf["gcam_act"] if last.pool else self._buf["act_" + last.name]
        ops.gradcam(act[:B], probs, self.view("output/kernel"), heat)
        return heat, probs

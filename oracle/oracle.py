"""torch-CPU restatement of the cutoffvgg16 hot path (TEST INFRASTRUCTURE, parity unpinned --
see oracle/__init__.py).  Follows SURVEY.md Appendix A:

* graph        -- src/models/models.py:407-417 (VGG16.layers[1:cutoff] -> GAP -> Dropout -> Dense softmax)
* loss         -- models.py:426,431 ``categorical_crossentropy`` + ``class_weight`` (train.py:135-136):
                  L = (1/b) * sum_i w_{y_i} * (logsumexp(z_i) - z_i[y_i])
* optimizers   -- models.py:397-398 Keras 2.9 Adam / RMSprop (epsilon 1e-7, eps added to the
                  UNCORRECTED sqrt(v))
* Grad-CAM     -- src/explainability/gradcam.py:68-83, defined per frame

``bf16=True`` emulates the device numerics: conv operands (weights and inter-layer activations)
rounded to bf16, fp32 accumulation; block1_conv1 and the head stay fp32 like the CUDA path.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .synth import layer_plan


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


class _RoundBF16(torch.autograd.Function):
    """Round to bf16 in forward; in backward round the gradient to bf16 too (the device stores
    inter-layer gradients in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return _bf16_round(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16_round(g)


def _conv_names(cutoff_layer):
    return [n for n, k, _ in layer_plan(cutoff_layer) if k == "conv"]


def forward(params, x_nhwc, cutoff_layer=10, training=False, drop_mask=None, dropout=0.45,
            bf16=False, keep=None, as_torch=False):
    """x_nhwc fp32 [B,H,W,3] (already BGR / mean-subtracted).  Returns dict with every activation
    (NHWC numpy, Keras layer names), 'gap', 'logits', 'probs'."""
    p = {k: (v if torch.is_tensor(v) else _t(v)) for k, v in params.items()}
    x = x_nhwc if torch.is_tensor(x_nhwc) else _t(x_nhwc)
    a = x.permute(0, 3, 1, 2)
    acts = {}
    first = True
    for name, kind, _ in layer_plan(cutoff_layer):
        if kind == "conv":
            w = p[name + "/kernel"].permute(3, 2, 0, 1)  # HWIO -> OIHW (cross-correlation, no flip)
            if bf16 and not first:
                w = _RoundBF16.apply(w)
            a = F.relu(F.conv2d(a, w, p[name + "/bias"], stride=1, padding=1))
            if bf16:
                a = _RoundBF16.apply(a)
            first = False
        else:
            a = F.max_pool2d(a, 2, 2)
        acts[name] = a
    gap = a.mean(dim=(2, 3))
    h = gap
    if training:
        assert drop_mask is not None, "training forward needs the realised keep-mask"
        m = drop_mask if torch.is_tensor(drop_mask) else _t(drop_mask)
        h = gap * m / (1.0 - dropout)
    logits = h @ p["output/kernel"] + p["output/bias"]
    probs = torch.softmax(logits, dim=-1)
    out = {"gap": gap, "logits": logits, "probs": probs}
    out.update(acts)
    if as_torch:
        return out
    res = {}
    for k, v in out.items():
        if keep is not None and k not in keep:
            continue
        v = v.detach()
        res[k] = (v.permute(0, 2, 3, 1) if v.dim() == 4 else v).contiguous().numpy()
    return res


def loss_from_logits(logits, y_onehot, class_weight=None):
    """Keras CCE on a softmax output with class_weight: sum_i w_i*l_i / b (Appendix A.4)."""
    y = y_onehot if torch.is_tensor(y_onehot) else _t(y_onehot)
    li = torch.logsumexp(logits, dim=-1) - (logits * y).sum(-1)
    if class_weight is not None:
        cw = torch.tensor([class_weight[i] for i in range(y.shape[1])], dtype=torch.float32)
        li = li * (y * cw).sum(-1)
    return li.sum() / y.shape[0]


def loss_and_grads(params, x_nhwc, y_onehot, class_weight=None, drop_mask=None, dropout=0.45,
                   cutoff_layer=10, trainable=None, bf16=False, loss_scale_batch=None):
    """One training forward+backward.  Returns (loss, grads dict in Keras layout, fwd dict).
    ``trainable``: iterable of layer names (default: all).  ``loss_scale_batch``: divide the
    summed loss by this instead of the local batch (data-parallel shards divide by the GLOBAL batch)."""
    names = _conv_names(cutoff_layer) + ["output"]
    trainable = set(names if trainable is None else trainable)
    p = {}
    for k, v in params.items():
        t = _t(np.array(v, dtype=np.float32, copy=True))
        t.requires_grad_(k.split("/")[0] in trainable)
        p[k] = t
    out = forward(p, x_nhwc, cutoff_layer, training=True, drop_mask=drop_mask, dropout=dropout,
                  bf16=bf16, as_torch=True)
    loss = loss_from_logits(out["logits"], y_onehot, class_weight)
    if loss_scale_batch is not None:
        loss = loss * (np.asarray(y_onehot).shape[0] / float(loss_scale_batch))
    loss.backward()
    grads = {k: t.grad.numpy() for k, t in p.items() if t.grad is not None}
    fwd = {k: v.detach().numpy() for k, v in out.items() if k in ("logits", "probs", "gap")}
    return float(loss.detach()), grads, fwd


def adam_step(theta, g, m, v, t, lr=3e-4, b1=0.9, b2=0.999, eps=1e-7):
    """Keras 2.9 ``Adam._resource_apply_dense`` (Appendix A.5); t is the 1-based step."""
    theta = np.asarray(theta, np.float32); g = np.asarray(g, np.float32)
    m = (b1 * m + (1 - b1) * g).astype(np.float32)
    v = (b2 * v + (1 - b2) * g * g).astype(np.float32)
    alpha = np.float32(lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t))
    theta = (theta - alpha * m / (np.sqrt(v) + np.float32(eps))).astype(np.float32)
    return theta, m, v


def rmsprop_step(theta, g, rms, lr=9.3e-6, rho=0.9, eps=1e-7):
    """Keras 2.9 RMSprop, momentum 0, not centered, rms0 = 0 (Appendix A.5)."""
    theta = np.asarray(theta, np.float32); g = np.asarray(g, np.float32)
    rms = (rho * rms + (1 - rho) * g * g).astype(np.float32)
    theta = (theta - np.float32(lr) * g / (np.sqrt(rms) + np.float32(eps))).astype(np.float32)
    return theta, rms


def _resize_bilinear_np(img, oh, ow):
    """cv2.resize(INTER_LINEAR) semantics: half-pixel centres, edge replicate (gradcam.py:81)."""
    ih, iw = img.shape
    ys = (np.arange(oh) + 0.5) * ih / oh - 0.5
    xs = (np.arange(ow) + 0.5) * iw / ow - 0.5
    y0 = np.floor(ys).astype(int); x0 = np.floor(xs).astype(int)
    fy = (ys - y0).astype(np.float32); fx = (xs - x0).astype(np.float32)
    y0c = np.clip(y0, 0, ih - 1); y1c = np.clip(y0 + 1, 0, ih - 1)
    x0c = np.clip(x0, 0, iw - 1); x1c = np.clip(x0 + 1, 0, iw - 1)
    top = img[y0c][:, x0c] * (1 - fx) + img[y0c][:, x1c] * fx
    bot = img[y1c][:, x0c] * (1 - fx) + img[y1c][:, x1c] * fx
    return (top * (1 - fy)[:, None] + bot * fy[:, None]).astype(np.float32)


def gradcam_autograd(params, x_nhwc, cutoff_layer=10, out_hw=None, bf16=False):
    """Per-frame Grad-CAM exactly as gradcam.py:68-80 (batch of one): gradient of the predicted
    class PROBABILITY w.r.t. the post-ReLU last conv output, spatial mean -> alpha, channel MEAN of
    alpha*A, ReLU, divide by max (0/0 := 0).  Returns (heat [B,h,w], probs [B,n])."""
    x = _t(x_nhwc)
    heats, probs = [], []
    last = _conv_names(cutoff_layer)[-1]
    # Grad-CAM needs the activation of the last conv BEFORE any trailing pool; with cutoff 10 the
    # last kept layer is block3_conv3 itself.
    for i in range(x.shape[0]):
        p = {k: _t(v) for k, v in params.items()}
        out = forward(p, x[i:i + 1], cutoff_layer, bf16=bf16, as_torch=True)
        A = out[last].detach().clone().requires_grad_(True)
        tail = A
        seen = False
        for name, kind, _ in layer_plan(cutoff_layer):
            if seen and kind == "pool":
                tail = F.max_pool2d(tail, 2, 2)
            if name == last:
                seen = True
        logits = tail.mean(dim=(2, 3)) @ p["output/kernel"] + p["output/bias"]
        pr = torch.softmax(logits, dim=-1)
        c = int(torch.argmax(pr[0]))
        pr[0, c].backward()
        alpha = A.grad.mean(dim=(0, 2, 3))                       # [C]
        heat = (alpha[None, :, None, None] * A.detach()).mean(dim=1)[0]
        heat = torch.clamp(heat, min=0.0).numpy()
        mx = heat.max()
        heat = heat / mx if mx > 0 else np.zeros_like(heat)
        if out_hw is not None:
            heat = _resize_bilinear_np(heat, *out_hw)
        heats.append(heat.astype(np.float32)); probs.append(pr.detach().numpy()[0])
    return np.stack(heats), np.stack(probs)


def gradcam_closed_form(A_nhwc, probs, w_dense, out_hw=None):
    """Closed form for a GAP->Dense->softmax head:
    alpha_k = (1/HW) * sum_j p_c (delta_cj - p_j) W[k,j]."""
    A = np.asarray(A_nhwc, np.float32)
    B, H, W, C = A.shape
    heats = []
    for i in range(B):
        p = probs[i].astype(np.float32)
        c = int(np.argmax(p))
        dpdz = p[c] * ((np.arange(len(p)) == c).astype(np.float32) - p)
        alpha = (w_dense @ dpdz) / np.float32(H * W)
        heat = np.maximum((A[i] * alpha).mean(axis=-1), 0.0)
        mx = heat.max()
        heat = heat / mx if mx > 0 else np.zeros_like(heat)
        if out_hw is not None:
            heat = _resize_bilinear_np(heat, *out_hw)
        heats.append(heat.astype(np.float32))
    return np.stack(heats)

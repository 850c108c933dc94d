"""Developer bring-up script (NOT a test, NOT the product): per-kernel checks on a B200 against
torch-on-GPU fp32 references with detailed diagnostics.  Each group runs in its own process so a
device trap in one kernel does not hide the others.  Usage: python tools/gpu_check.py [group ...]
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["fwd", "dgrad", "wgrad", "first", "pointwise", "head", "optim", "gradcam", "e2e"]


def _setup():
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    return torch


def _report(name, got, ref, tol):
    import torch
    got = got.float(); ref = ref.float()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    mx = err.max().item()
    idx = err.argmax().item()
    bad = (err > tol * scale).sum().item()
    ok = mx <= tol * scale and torch.isfinite(got).all().item()
    print(f"[{'OK ' if ok else 'BAD'}] {name}: max|err|={mx:.4g} (scale {scale:.4g}, rel {mx / scale:.3g}) "
          f"bad={bad}/{got.numel()} argmax_idx={idx}", flush=True)
    if not ok:
        flat_g, flat_r = got.flatten(), ref.flatten()
        print("      got[idx], ref[idx] =", flat_g[idx].item(), flat_r[idx].item(),
              " unravel:", [int(v) for v in torch.unravel_index(torch.tensor(idx), got.shape)], flush=True)
        print("      got[:8] =", [round(v, 4) for v in flat_g[:8].tolist()], flush=True)
        print("      ref[:8] =", [round(v, 4) for v in flat_r[:8].tolist()], flush=True)
    return ok


def _conv_ref(torch, x_bf, w_hwio, bias=None, relu=True):
    import torch.nn.functional as F
    y = F.conv2d(x_bf.float().permute(0, 3, 1, 2), w_hwio.to(torch.bfloat16).float().permute(3, 2, 0, 1),
                 bias, padding=1)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


LAYERS = [(64, 64, 128, 128, True), (64, 128, 64, 64, False), (128, 128, 64, 64, True),
          (128, 256, 32, 32, False), (256, 256, 32, 32, False)]


def group_fwd():
    torch = _setup()
    import torch.nn.functional as F
    from ab_line_classifier_b200 import ops
    ok = True
    for (cin, cout, H, W, pool) in LAYERS:
        for B in (1, 3):
            x = (torch.randn(B, H, W, cin, device="cuda") * 0.5).to(torch.bfloat16)
            w = torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5
            b = torch.randn(cout, device="cuda") * 0.1
            wf = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
            ops.prep_conv_weights(w, wf, None)
            ref_wf = w.permute(1, 0, 3, 2).reshape(9, cout, cin).to(torch.bfloat16)
            ok &= _report(f"prep w_fwd {cin}->{cout}", wf, ref_wf, 0.0)
            full = torch.full((B, H, W, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
            pl = torch.full((B, H // 2, W // 2, cout), float("nan"), device="cuda", dtype=torch.bfloat16) if pool else None
            ops.conv3x3_igemm(x, wf, b, None, full, pl, ops.EPI_BIAS_RELU)
            torch.cuda.synchronize()
            ref = _conv_ref(torch, x, w, b)
            ok &= _report(f"conv fwd {cin}->{cout} {H}x{W} B={B}", full, ref, 1e-2)
            if pool:
                refp = F.max_pool2d(full.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
                ok &= _report(f"  fused pool {cin}->{cout}", pl, refp, 0.0)
    return ok


def group_dgrad():
    torch = _setup()
    import torch.nn.functional as F
    from ab_line_classifier_b200 import ops
    ok = True
    for (cin, cout, H, W, _) in LAYERS:
        B = 2
        dz = (torch.randn(B, H, W, cout, device="cuda") * 0.1).to(torch.bfloat16)
        w = torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5
        act = torch.relu(torch.randn(B, H, W, cin, device="cuda")).to(torch.bfloat16)   # layer-below activation
        wd = torch.empty(9, cin, cout, device="cuda", dtype=torch.bfloat16)
        ops.prep_conv_weights(w, None, wd)
        out = torch.full((B, H, W, cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv3x3_igemm(dz, wd, None, act, out, None, ops.EPI_RELU_MASK)
        torch.cuda.synchronize()
        wq = w.to(torch.bfloat16).float().permute(3, 2, 0, 1)
        dx = F.conv_transpose2d(dz.float().permute(0, 3, 1, 2), wq, padding=1).permute(0, 2, 3, 1)
        ref = dx * (act.float() > 0)
        ok &= _report(f"dgrad {cout}->{cin} {H}x{W}", out, ref, 1e-2)
    return ok


def group_wgrad():
    torch = _setup()
    from ab_line_classifier_b200 import ops
    ok = True
    for (cin, cout, H, W, _) in LAYERS:
        for B in (1, 4):
            x = torch.relu(torch.randn(B, H, W, cin, device="cuda")).to(torch.bfloat16)
            dz = (torch.randn(B, H, W, cout, device="cuda") * 0.1).to(torch.bfloat16)
            ws = torch.empty(ops.wgrad_workspace_bytes(B, H, W, cin, cout), device="cuda", dtype=torch.uint8)
            dw = torch.full((3, 3, cin, cout), float("nan"), device="cuda")
            db = torch.zeros(cout, device="cuda")
            ops.conv3x3_wgrad(x, dz, dw, db, ws, accumulate=False)
            torch.cuda.synchronize()
            ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (cout, cin, 3, 3),
                                              dz.float().permute(0, 3, 1, 2), padding=1).permute(2, 3, 1, 0)
            ok &= _report(f"wgrad {cin}->{cout} {H}x{W} B={B}", dw, ref, 5e-3)
            ok &= _report(f"  bias grad", db, dz.float().sum(dim=(0, 1, 2)), 5e-3)
    return ok


def group_first():
    torch = _setup()
    import torch.nn.functional as F
    from ab_line_classifier_b200 import ops
    ok = True
    B, H, W = 3, 128, 128
    x = (torch.rand(B, H, W, 3, device="cuda") * 255 - 115).contiguous()
    w = torch.randn(3, 3, 3, 64, device="cuda") * 0.1
    b = torch.randn(64, device="cuda") * 0.1
    out = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    ops.conv_first_fwd(x, w, b, out)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=1)).permute(0, 2, 3, 1)
    ok &= _report("conv_first fwd", out, ref, 5e-3)
    dz = (torch.randn(B, H, W, 64, device="cuda") * 0.1).to(torch.bfloat16)
    dw = torch.zeros(3, 3, 3, 64, device="cuda"); db = torch.zeros(64, device="cuda")
    ops.conv_first_wgrad(x, dz, dw, db)
    refw = torch.nn.grad.conv2d_weight(x.permute(0, 3, 1, 2), (64, 3, 3, 3), dz.float().permute(0, 3, 1, 2),
                                       padding=1).permute(2, 3, 1, 0)
    ok &= _report("conv_first wgrad", dw, refw, 1e-4)
    ok &= _report("conv_first bias grad", db, dz.float().sum(dim=(0, 1, 2)), 1e-4)
    return ok


def group_pointwise():
    torch = _setup()
    import torch.nn.functional as F
    from ab_line_classifier_b200 import ops
    ok = True
    B, H, W, Cc = 2, 64, 64, 128
    a = torch.relu(torch.randn(B, H, W, Cc, device="cuda")).to(torch.bfloat16)
    a[0, :8, :8, :] = 0       # all-zero windows
    a[1, 0:2, 0:2, :] = 1.5   # positive ties -> first element wins
    a_nchw = a.float().permute(0, 3, 1, 2).requires_grad_(True)
    p = F.max_pool2d(a_nchw, 2, 2)
    dp = (torch.randn(B, H // 2, W // 2, Cc, device="cuda")).to(torch.bfloat16)
    p.backward(dp.float().permute(0, 3, 1, 2))
    ref = a_nchw.grad.permute(0, 2, 3, 1) * (a.float() > 0)
    pooled = p.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    dz = torch.full_like(a, float("nan"))
    ops.maxpool_bwd_relu(dp, a, pooled, dz)
    ok &= _report("maxpool_bwd_relu", dz, ref, 0.0)
    a3 = torch.relu(torch.randn(3, 32, 32, 256, device="cuda")).to(torch.bfloat16)
    gap = torch.empty(3, 256, device="cuda")
    ops.gap_fwd(a3, gap)
    ok &= _report("gap_fwd", gap, a3.float().mean(dim=(1, 2)), 1e-5)
    dgap = torch.randn(3, 256, device="cuda")
    dz3 = torch.empty_like(a3)
    ops.gap_bwd_relu(dgap, a3, dz3)
    ref3 = (dgap[:, None, None, :] / 1024.0) * (a3.float() > 0)
    ok &= _report("gap_bwd_relu", dz3, ref3.to(torch.bfloat16), 0.0)
    return ok


def group_head():
    torch = _setup()
    from ab_line_classifier_b200 import ops
    ok = True
    B, Cc, n = 37, 256, 2
    gap = torch.rand(B, Cc, device="cuda")
    w = torch.randn(Cc, n, device="cuda") * 0.15
    b = torch.tensor([0.1, -0.2], device="cuda")
    y = torch.zeros(B, n, device="cuda"); y[torch.arange(B), torch.arange(B) % 2] = 1
    cw = torch.tensor([0.75, 1.5], device="cuda")
    logits = torch.empty(B, n, device="cuda"); probs = torch.empty(B, n, device="cuda")
    loss = torch.zeros(1, device="cuda"); dlog = torch.empty(B, n, device="cuda")
    mask = torch.empty(B, Cc, device="cuda", dtype=torch.uint8)
    ops.head_fwd(gap, w, b, logits, probs, y, cw, 1.0 / B, loss, dlog, mask, False, True, 0.45, 10001, 3, 0)
    keep = mask.float()
    print("      dropout keep fraction:", keep.mean().item(), "(expect ~0.55)")
    ok &= abs(keep.mean().item() - 0.55) < 0.03
    h = (gap * keep / 0.55).requires_grad_(True)
    wt = w.clone().requires_grad_(True); bt = b.clone().requires_grad_(True)
    z = h @ wt + bt
    li = torch.logsumexp(z, -1) - (z * y).sum(-1)
    L = (li * (y * cw).sum(-1)).sum() / B
    L.backward()
    ok &= _report("head logits", logits, z.detach(), 1e-5)
    ok &= _report("head probs", probs, torch.softmax(z.detach(), -1), 1e-5)
    ok &= _report("head loss", loss, L.detach().reshape(1), 1e-5)
    dw = torch.zeros(Cc, n, device="cuda"); db = torch.zeros(n, device="cuda"); dgap = torch.empty(B, Cc, device="cuda")
    ops.head_bwd(gap, w, dlog, mask, 0.45, dw, db, dgap)
    ok &= _report("head dW", dw, wt.grad, 1e-5)
    ok &= _report("head db", db, bt.grad, 1e-5)
    ok &= _report("head dgap", dgap, h.grad * keep / 0.55, 1e-5)
    # inference mode
    ops.head_fwd(gap, w, b, logits, probs)
    ok &= _report("head probs (inference)", probs, torch.softmax(gap @ w + b, -1), 1e-5)
    return ok


def group_optim():
    torch = _setup()
    import numpy as np
    from ab_line_classifier_b200 import ops
    from oracle import oracle
    ok = True
    n = 100003
    th = torch.randn(n, device="cuda"); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    th_np, m_np, v_np = th.cpu().numpy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in (1, 2, 3):
        g = torch.randn(n, device="cuda") * 0.01
        alpha = 3e-4 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ops.adam_flat(th, g, m, v, alpha)
        th_np, m_np, v_np = oracle.adam_step(th_np, g.cpu().numpy(), m_np, v_np, t)
    ok &= _report("adam theta (3 steps)", th, torch.from_numpy(th_np).cuda(), 1e-6)
    ok &= _report("adam v", v, torch.from_numpy(v_np).cuda(), 1e-5)
    th = torch.randn(n, device="cuda"); r = torch.zeros(n, device="cuda")
    th_np, r_np = th.cpu().numpy(), np.zeros(n, np.float32)
    for t in (1, 2, 3):
        g = torch.randn(n, device="cuda") * 0.01
        ops.rmsprop_flat(th, g, r, 9.3e-6)
        th_np, r_np = oracle.rmsprop_step(th_np, g.cpu().numpy(), r_np)
    ok &= _report("rmsprop theta (3 steps)", th, torch.from_numpy(th_np).cuda(), 1e-6)
    return ok


def group_gradcam():
    torch = _setup()
    import numpy as np
    from ab_line_classifier_b200 import ops
    from oracle import oracle
    ok = True
    B = 5
    a = torch.relu(torch.randn(B, 32, 32, 256, device="cuda")).to(torch.bfloat16)
    a[4] = 0   # all-zero map -> zeros, not NaN
    w = torch.randn(256, 2, device="cuda") * 0.15
    probs = torch.softmax(a.float().mean(dim=(1, 2)) @ w, -1).contiguous()
    heat = torch.empty(B, 128, 128, device="cuda")
    ops.gradcam(a, probs, w, heat)
    ref = oracle.gradcam_closed_form(a.float().cpu().numpy(), probs.cpu().numpy(), w.cpu().numpy(), (128, 128))
    ok &= _report("gradcam 128x128", heat, torch.from_numpy(ref).cuda(), 1e-4)
    heat32 = torch.empty(B, 32, 32, device="cuda")
    ops.gradcam(a, probs, w, heat32)
    ref32 = oracle.gradcam_closed_form(a.float().cpu().numpy(), probs.cpu().numpy(), w.cpu().numpy(), None)
    ok &= _report("gradcam 32x32", heat32, torch.from_numpy(ref32).cuda(), 1e-4)
    return ok


def group_e2e():
    torch = _setup()
    import numpy as np
    from ab_line_classifier_b200.engine import CutoffVGG16Engine
    from oracle import oracle, synth
    ok = True
    B = 8
    fr, lab = synth.synth_frames(B, seed=1234)
    x, y = synth.vgg16_preprocess(fr), synth.one_hot(lab)
    params = synth.randomize_biases(synth.glorot_weights(10001))
    eng = CutoffVGG16Engine()
    eng.set_weights(params)
    probs = eng.predict_batch(x).cpu().numpy()
    logits = eng._buf["logits"][:B].cpu().numpy()
    ref32 = oracle.forward(params, x)
    refbf = oracle.forward(params, x, bf16=True)
    print("      logits gpu :", logits[:3].round(4).tolist())
    print("      logits fp32:", ref32["logits"][:3].round(4).tolist())
    print("      logits bf16:", refbf["logits"][:3].round(4).tolist())
    ok &= _report("e2e logits vs fp32 oracle", torch.from_numpy(logits), torch.from_numpy(ref32["logits"]), 2e-2)
    ok &= _report("e2e logits vs bf16 oracle", torch.from_numpy(logits), torch.from_numpy(refbf["logits"]), 5e-3)
    ok &= _report("e2e block3_conv3 vs bf16 oracle", eng._buf["act_block3_conv3"][:B].float().cpu(),
                  torch.from_numpy(refbf["block3_conv3"]), 2e-2)
    print("      argmax equal:", (probs.argmax(-1) == ref32["probs"].argmax(-1)).all())
    # one train step, all layers trainable, explicit dropout mask, class weights
    mask = synth.dropout_mask(B, 256, 0.45)
    cw = {0: 0.75, 1: 1.5}
    eng.set_class_weight(cw)
    loss = eng.train_step(x, y, set(eng.layer_names()), keep_mask=mask, apply_update=False)
    l_ref, g_ref, _ = oracle.loss_and_grads(params, x, y, cw, mask, bf16=True)
    l32, g32, _ = oracle.loss_and_grads(params, x, y, cw, mask)
    print("      loss gpu / bf16 oracle / fp32 oracle:", float(loss.item()), l_ref, l32)
    ok &= abs(float(loss.item()) - l_ref) < 5e-3 * max(1.0, abs(l_ref))
    for name in eng.segments:
        g = eng.view(name, eng.grad).cpu()
        ok &= _report(f"grad {name} vs bf16 oracle", g, torch.from_numpy(g_ref[name]), 3e-2)
        _report(f"   (info) vs fp32 oracle", g, torch.from_numpy(g32[name]), 5e-2)
    return ok


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        t0 = time.time()
        good = globals()["group_" + sys.argv[2]]()
        import torch
        torch.cuda.synchronize()
        print(f"group {sys.argv[2]}: {'PASS' if good else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        sys.exit(0 if good else 1)
    groups = sys.argv[1:] or GROUPS
    rc = 0
    for g in groups:
        print(f"===== {g} =====", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], timeout=300)
            if r.returncode != 0:
                rc = 1
                print(f"group {g}: exit code {r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            rc = 1
            print(f"group {g}: TIMEOUT", flush=True)
    sys.exit(rc)

"""Data-parallel path on CPU: world_size 2, gloo backend.  Checks the bucketed all-reduce plumbing
(ab_line_classifier_b200/dp.py) and that SUM of shard gradients (loss scaled by 1/global batch)
equals the single-device gradient of the concatenated batch."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ab_line_classifier_b200.dp import GradBucketReducer
    from oracle import oracle, synth
    rng = np.random.default_rng(2)
    params = {"block1_conv1/kernel": rng.normal(size=(3, 3, 3, 4)).astype(np.float32) * 0.3,
              "block1_conv1/bias": rng.normal(size=4).astype(np.float32) * 0.1,
              "block1_conv2/kernel": rng.normal(size=(3, 3, 4, 5)).astype(np.float32) * 0.3,
              "block1_conv2/bias": rng.normal(size=5).astype(np.float32) * 0.1,
              "output/kernel": rng.normal(size=(5, 2)).astype(np.float32),
              "output/bias": rng.normal(size=2).astype(np.float32)}
    G = 8
    x = rng.normal(size=(G, 8, 8, 3)).astype(np.float32)
    y = synth.one_hot(np.arange(G) % 2)
    mask = np.ones((G, 5), np.float32)
    cw = {0: 0.75, 1: 1.5}
    sl = slice(rank * G // world, (rank + 1) * G // world)          # contiguous shard per rank
    _, g, _ = oracle.loss_and_grads(params, x[sl], y[sl], cw, mask[sl], 0.0, 3, loss_scale_batch=G)
    # flat layout in layer order, reduced in REVERSE layer order (as backward produces them)
    names = list(params)
    offs, off = {}, 0
    for n in names:
        offs[n] = (off, off + params[n].size)
        off += (params[n].size + 3) // 4 * 4
    flat = torch.zeros(off)
    for n in names:
        flat[offs[n][0]:offs[n][1]] = torch.from_numpy(g[n].reshape(-1))
    red = GradBucketReducer(flat, bucket_bytes=400)
    layers = [("output/kernel", "output/bias"), ("block1_conv2/kernel", "block1_conv2/bias"),
              ("block1_conv1/kernel", "block1_conv1/bias")]
    for k, b in layers:
        red.hook(offs[k][0], (offs[b][1] + 3) // 4 * 4)
    red.wait()
    if rank == 0:
        _, g_all, _ = oracle.loss_and_grads(params, x, y, cw, mask, 0.0, 3)
        err = max(float(np.abs(flat[offs[n][0]:offs[n][1]].numpy() - g_all[n].reshape(-1)).max()) for n in names)
        q.put((err, red.buckets))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_device_gradient():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, buckets = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5, err
    # 3 layers of 12 / 185 / 112 floats with a 400-byte bucket: head alone is below the threshold and is
    # coalesced with block1_conv2; block1_conv1 flushes on its own.
    assert len(buckets) == 2 and buckets[0][1] > buckets[1][1] and buckets[0][0] == buckets[1][1]

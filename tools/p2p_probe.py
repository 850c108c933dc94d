"""Developer tool: which peer-memory mechanisms work on this box?  (torchrun, >= 2 GPUs)
  A. torch.distributed._symmetric_memory (empty + rendezvous: peer pointers, multicast pointer, device barrier)
  B. CUDA IPC of a caching-allocator tensor (UntypedStorage._share_cuda_ / _new_shared_cuda)
"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.enable()
faulthandler.dump_traceback_later(80, exit=True)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
T0 = time.time()


def say(m):
    print(f"[rank {rank} +{time.time() - T0:5.1f}s] {m}", flush=True)


say(f"can_device_access_peer(0,1)={torch.cuda.can_device_access_peer(0, 1)}")
# ---- A
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local_rank))
    t.fill_(rank + 1)
    hdl = symm.rendezvous(t, dist.group.WORLD)
    say(f"A: rendezvous ok; world={hdl.world_size} multicast={hdl.has_multicast_support} mc_ptr={hdl.multicast_ptr:#x} "
        f"ptrs={[hex(p) for p in hdl.buffer_ptrs]}")
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
    torch.cuda.synchronize()
    say(f"A: peer buffer first element = {peer[0].item()} (expect {(rank + 1) % world + 1})")
    hdl.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        hdl.barrier()
    e1.record()
    torch.cuda.synchronize()
    say(f"A: device barrier = {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
    dst = torch.empty(1 << 20, device="cuda")
    e0.record()
    for _ in range(20):
        dst.copy_(peer)
    e1.record()
    torch.cuda.synchronize()
    say(f"A: 4 MiB peer read = {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
except Exception as e:  # noqa: BLE001
    say(f"A: FAILED {type(e).__name__}: {e}")

# ---- B
try:
    g = torch.full((1 << 20,), float(rank + 10), device="cuda")
    h = g.untyped_storage()._share_cuda_()
    hs = [None] * world
    dist.all_gather_object(hs, h)
    other = (rank + 1) % world
    st = torch.UntypedStorage._new_shared_cuda(*hs[other])
    pt = torch.empty(0, dtype=torch.float32, device=st.device).set_(st)
    torch.cuda.synchronize()
    say(f"B: IPC storage device={st.device} first element = {pt[0].item()} (expect {other + 10})")
    dist.barrier()
except Exception as e:  # noqa: BLE001
    say(f"B: FAILED {type(e).__name__}: {e}")
dist.barrier()
faulthandler.cancel_dump_traceback_later()
dist.destroy_process_group()

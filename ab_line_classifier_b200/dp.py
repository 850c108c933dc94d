"""Data-parallel gradient exchange (new functionality: the reference is single-device).

One process per GPU; every rank holds the full flat fp32 gradient buffer.  As backward finishes a
layer it reports the slice [start, end); slices are coalesced into buckets of >= ``bucket_bytes`` and
each bucket is all-reduced (SUM) asynchronously with torch.distributed (NCCL over NVLink/NVSwitch on
GPUs, gloo in the CPU tests) while the remaining dgrad/wgrad kernels run.  Each rank scales its loss by
1/GLOBAL batch, so the SUM equals the single-device gradient on the concatenated batch.
"""
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, grad_flat, bucket_bytes=2 << 20, group=None):
        self.grad, self.bucket_bytes, self.group = grad_flat, int(bucket_bytes), group
        self.pending, self.handles, self.buckets = [], [], []

    def hook(self, start, end):
        """grad[start:end] is final on this rank."""
        self.pending.append((start, end))
        if sum(e - s for s, e in self.pending) * self.grad.element_size() >= self.bucket_bytes:
            self.flush()

    def flush(self):
        if not self.pending:
            return
        s, e = min(p[0] for p in self.pending), max(p[1] for p in self.pending)
        self.buckets.append((s, e))
        self.handles.append(dist.all_reduce(self.grad[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.pending = []

    def wait(self):
        self.flush()
        for h in self.handles:
            h.wait()
        self.handles = []

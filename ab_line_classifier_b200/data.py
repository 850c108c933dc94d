"""Host input pipeline stand-in for the reference's tf.data ``Preprocessor.prepare`` output
(src/data/preprocessor.py:36-66): batches of (x [b,H,W,3], y one-hot fp32 [b,n]), reshuffled over the full dataset
every epoch, last batch short (no drop_remainder).

x is either fp32 frames as ``vgg16_preprocess`` leaves them or RAW uint8 RGB frames (4x smaller; preprocessing and
resize then run on the device).  Two residencies:
  * host (default): frames in pinned memory; every batch is gathered by index into one of a few rotating pinned
    buffers (no per-epoch copy of the whole dataset), so the engine's H2D copies stay asynchronous;
  * device (``device="cuda"``): the whole dataset lives in HBM (180 GB per B200: 1 M uint8 frames = 49 GB) and a batch
    is one device-side gather -- the shuffle never touches the host.

``augment=cfg['TRAIN']['DATA_AUG']`` reproduces ``prepare(..., augment=True)`` (preprocessor.py:56-63): x must then be
RGB frames in [0,255] (uint8 or fp32, NOT yet preprocessed); every batch is augmented and preprocessed on the device
(augment.py) and arrives as a device fp32 tensor.
"""
import numpy as np
import torch

# Pinned batch buffers are refilled by the host while earlier asynchronous copies out of them may still be queued on a
# CUDA stream (the host can run many steps ahead of the device).  Whoever enqueues such a copy reports it here; the
# dataset waits for that event before it overwrites the buffer.
_tracked, _copy_fences = set(), {}


def note_async_copy(t, event=None):
    """Called after an asynchronous H2D copy out of host tensor ``t`` was enqueued (``event``: recorded behind it)."""
    if not torch.is_tensor(t) or t.is_cuda:
        return
    key = t.untyped_storage().data_ptr()
    if key in _tracked:
        if event is None:
            event = torch.cuda.Event()
            event.record()
        _copy_fences[key] = event


def _wait_async_copies(t):
    ev = _copy_fences.pop(t.untyped_storage().data_ptr(), None)
    if ev is not None:
        ev.synchronize()


class ArrayDataset:
    N_PINNED = 4      # rotating pinned batch buffers: a buffer is reused only after 3 later batches were staged

    def __init__(self, x, y=None, batch_size=64, shuffle=False, seed=0, pin=True, device=None, augment=None,
                 img_dim=None, rank=0, world_size=1):
        x = np.asarray(x) if not torch.is_tensor(x) else x
        if not torch.is_tensor(x):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint8 if x.dtype == np.uint8 else np.float32))
        y = None if y is None else (y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)))
        self.device = torch.device(device) if device is not None else None
        if self.device is not None:
            x = x.to(self.device)
            y = None if y is None else y.to(self.device)
        elif pin and torch.cuda.is_available():
            x = x.pin_memory()
            y = None if y is None else y.pin_memory()
        self.x, self.y = x, y
        self.batch_size, self.shuffle = int(batch_size), shuffle
        self.rng = np.random.default_rng(seed)
        # data parallel: batch_size is the GLOBAL batch; every rank walks the same (seeded) order
        # and takes its contiguous 1/world_size of each global batch, so the ranks' gradient SUM is the single-device
        # gradient on the whole batch.  A short last batch is cut to a multiple of world_size (at most world_size - 1
        # frames of an epoch are skipped) so that all ranks see equal local batches.
        self.rank, self.world_size = int(rank), int(world_size)
        if not 0 <= self.rank < self.world_size:
            raise ValueError("need 0 <= rank < world_size")
        if self.batch_size % self.world_size:
            raise ValueError("the global batch size must be a multiple of world_size")
        self._pinned = pin and torch.cuda.is_available() and self.device is None
        self._bufs, self._k = None, 0
        self._aug, self.img_dim = None, (None if img_dim is None else tuple(img_dim))
        if augment is not None and augment is not False:
            from .augment import DeviceAugmenter
            self._aug = DeviceAugmenter(None if augment is True else augment, seed=seed + 1,
                                        device=self.device if self.device is not None else "cuda")

    def __len__(self):
        n = len(self.x)
        full, rest = divmod(n, self.batch_size)
        return full + (1 if rest >= self.world_size else 0)

    def _local(self, lo, hi):
        """This rank's contiguous share [lo', hi') of the global batch [lo, hi)."""
        per = (hi - lo) // self.world_size
        return lo + self.rank * per, lo + (self.rank + 1) * per

    @property
    def n_frames(self):
        return len(self.x)

    def _gather(self, idx):
        """Rows ``idx`` of (x, y): device gather, or host gather into the next rotating pinned buffer."""
        if self.device is not None:        # idx already lives on the device (one permutation upload per epoch)
            return self.x.index_select(0, idx), (None if self.y is None else self.y.index_select(0, idx))
        if not self._pinned:
            return self.x[idx], (None if self.y is None else self.y[idx])
        if self._bufs is None:
            b = self.batch_size // self.world_size
            self._bufs = [(torch.empty((b,) + tuple(self.x.shape[1:]), dtype=self.x.dtype).pin_memory(),
                           None if self.y is None else torch.empty((b,) + tuple(self.y.shape[1:]), dtype=self.y.dtype).pin_memory())
                          for _ in range(self.N_PINNED)]
            for bx, by in self._bufs:
                _tracked.update(t.untyped_storage().data_ptr() for t in (bx, by) if t is not None)
        bx, by = self._bufs[self._k]
        self._k = (self._k + 1) % self.N_PINNED
        _wait_async_copies(bx)
        if by is not None:
            _wait_async_copies(by)
        n = len(idx)
        torch.index_select(self.x, 0, idx, out=bx[:n])
        if self.y is not None:
            torch.index_select(self.y, 0, idx, out=by[:n])
        return bx[:n], (None if self.y is None else by[:n])

    def _batches(self):
        n = len(self.x)
        if not self.shuffle:
            for i in range(0, n, self.batch_size):
                lo, hi = self._local(i, min(i + self.batch_size, n))
                if hi > lo:
                    yield self.x[lo:hi], (None if self.y is None else self.y[lo:hi])
            return
        perm = torch.from_numpy(self.rng.permutation(n))
        if self.device is not None:
            # a per-batch upload of the indices from pageable memory would block the host behind all queued kernels
            perm = perm.to(self.device)
        for i in range(0, n, self.batch_size):
            lo, hi = self._local(i, min(i + self.batch_size, n))
            if hi > lo:
                yield self._gather(perm[lo:hi])

    def _augmented(self, x):
        """One batch of RGB frames -> device fp32, augmented and preprocessed (preprocessor.py:56-63)."""
        if self.img_dim is not None and tuple(x.shape[1:3]) != self.img_dim:
            if x.dtype != torch.uint8:
                raise ValueError("only raw uint8 frames are resized on the device")
            from . import ops
            xd = x.to(self._aug.device, non_blocking=True)
            note_async_copy(x)
            x = torch.empty((x.shape[0],) + self.img_dim + (3,), device=xd.device, dtype=torch.uint8)
            ops.resize_nearest_u8(xd, x)
        out = self._aug(x)
        note_async_copy(x)
        return out

    def __iter__(self):
        if self._aug is None:
            yield from self._batches()
            return
        for x, y in self._batches():
            yield self._augmented(x), y

    def __del__(self):
        try:
            for bx, by in (getattr(self, "_bufs", None) or []):
                for t in (bx, by):
                    if t is not None:
                        _tracked.discard(t.untyped_storage().data_ptr())
                        _copy_fences.pop(t.untyped_storage().data_ptr(), None)
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass

    def unbatch(self):
        """(frame, label) pairs, like ``tf.data.Dataset.unbatch`` (train_utils.py:285)."""
        for i in range(len(self.x)):
            yield self.x[i], (None if self.y is None else self.y[i])

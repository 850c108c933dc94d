"""Host input pipeline stand-in for the reference's tf.data ``Preprocessor.prepare`` output
(src/data/preprocessor.py:36-66): batches of (x fp32 [b,H,W,3] preprocessed, y one-hot fp32 [b,n]),
reshuffled over the full dataset every epoch, last batch short (no drop_remainder).  Frames are held in
pinned host memory so the engine's H2D copies are asynchronous."""
import numpy as np
import torch


class ArrayDataset:
    def __init__(self, x, y=None, batch_size=64, shuffle=False, seed=0, pin=True):
        # raw uint8 RGB frames stay uint8 (4x smaller H2D copies; preprocessing / resize run on the device)
        x = np.asarray(x)
        self.x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint8 if x.dtype == np.uint8 else np.float32))
        self.y = None if y is None else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        if pin and torch.cuda.is_available():
            self.x = self.x.pin_memory()
            self.y = None if self.y is None else self.y.pin_memory()
        self.batch_size, self.shuffle = int(batch_size), shuffle
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        return (len(self.x) + self.batch_size - 1) // self.batch_size

    @property
    def n_frames(self):
        return len(self.x)

    def __iter__(self):
        n = len(self.x)
        if self.shuffle:
            perm = torch.from_numpy(self.rng.permutation(n))
            x, y = self.x[perm], (None if self.y is None else self.y[perm])
        else:
            x, y = self.x, self.y
        for i in range(0, n, self.batch_size):
            yield x[i:i + self.batch_size], (None if y is None else y[i:i + self.batch_size])

    def unbatch(self):
        """(frame, label) pairs, like ``tf.data.Dataset.unbatch`` (train_utils.py:285)."""
        for i in range(len(self.x)):
            yield self.x[i], (None if self.y is None else self.y[i])

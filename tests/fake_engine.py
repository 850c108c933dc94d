"""CPU test double for CutoffVGG16Engine: same host-facing surface, numerics from the oracle.
Lives in tests/ only -- the product has no CPU path."""
import numpy as np
import torch

from ab_line_classifier_b200.engine import build_plan
from oracle import metric_counts, oracle, synth


class FakeEngine:
    calls = []          # (kind, lr, frozenset(trainable)) per optimizer application, across instances

    def __init__(self, input_shape=(8, 8, 3), n_classes=2, cutoff_layer=3, dropout=0.45, output_bias=None,
                 seed=10001, device=None, rank=0, world_size=1, process_group=None):
        self.input_shape, self.n_classes, self.cutoff_layer, self.dropout = tuple(input_shape), n_classes, cutoff_layer, dropout
        self.specs, self.feat_shape = build_plan(self.input_shape, cutoff_layer)
        self.device, self.rank, self.world_size = torch.device("cpu"), rank, world_size
        self.params = synth.glorot_weights(seed, cutoff_layer, n_classes, output_bias)
        self.n_params = sum(v.size for v in self.params.values())
        self.theta = torch.zeros(1)
        self.state, self.steps, self.step_count = {}, {}, 0
        self.class_weight = None
        self._buf = {"probs": torch.zeros(0, n_classes)}
        self.grads = None

    def layer_names(self):
        return [s.name for s in self.specs] + ["output"]

    def _ensure(self, b, t): pass
    def reserve(self, b, training=False): pass

    def stage(self, x, y=None, slot=None):
        self._slots = getattr(self, "_slots", {})
        slot = 0 if slot is None else slot
        self._slots[slot] = (x, y)
        return slot
    _cap = 1 << 30

    def use_slot(self, slot):
        self._cur = slot

    def release_slot(self): pass

    def forward_device(self, B, need_last_act=True):
        return self.predict_batch(self._slots[self._cur][0])

    def refresh_operands(self, layers=None): pass
    def get_weights(self): return {k: v.copy() for k, v in self.params.items()}

    def set_weights(self, w):
        for k in self.params:
            self.params[k] = np.asarray(w[k], np.float32).copy()

    def set_class_weight(self, cw):
        self.class_weight = None if cw is None else {int(k): float(v) for k, v in dict(cw).items()}

    def _np(self, a):
        return a.numpy() if torch.is_tensor(a) else np.asarray(a, np.float32)

    def predict_batch(self, x):
        p = oracle.forward(self.params, self._np(x), self.cutoff_layer)["probs"]
        self._buf["probs"] = torch.from_numpy(p)
        return self._buf["probs"]

    def eval_batch(self, x, y, loss_out=None, inv_batch=1.0, metrics=False):
        out = oracle.forward(self.params, self._np(x), self.cutoff_layer, as_torch=True)
        li = oracle.loss_from_logits(out["logits"], self._np(y)) * len(x) * inv_batch
        if loss_out is not None:
            loss_out += float(li)
        self._buf["probs"], self._last_y = out["probs"].detach(), self._np(y)
        if metrics:
            self.metrics_update(len(x))
        return out["probs"].detach()

    def metrics_begin(self, auc_thresholds, class_thresholds):
        self._met = [np.asarray(auc_thresholds, np.float32), np.asarray(class_thresholds, np.float32), None]

    def metrics_update(self, B):
        c = metric_counts.counts_from_arrays(self._last_y[:B], self._np(self._buf["probs"])[:B], self._met[0], self._met[1])
        self._met[2] = c if self._met[2] is None else self._met[2] + c

    def metrics_read(self, reduce=True):
        return self._met[2]

    def sum_over_ranks(self, values):
        return [float(v) for v in values]

    def train_step(self, x, y, trainable, optimizer="adam", lr=3e-4, global_batch=None, keep_mask=None,
                   apply_update=True, bucket_bytes=0, loss_out=None, slot=None, batch=None, defer_reduce=False):
        if slot is not None:
            x, y = self._slots[slot]
        x, y = self._np(x), self._np(y)
        mask = synth.dropout_mask(len(x), self.feat_shape[2], self.dropout, step=self.step_count)
        self.step_count += 1
        loss, self.grads, fwd = oracle.loss_and_grads(self.params, x, y, self.class_weight, mask, self.dropout,
                                                      self.cutoff_layer, trainable)
        self._buf["probs"], self._last_y = torch.from_numpy(fwd["probs"]), y
        t = torch.tensor([loss], dtype=torch.float32)
        if loss_out is not None:
            loss_out.copy_(t)          # the step's loss overwrites its destination (engine.forward_device)
            return loss_out
        return t

    def fit_step(self, trainable, optimizer, lr, slot, batch, global_batch, loss_out=None, **kw):
        self.train_step(None, None, trainable, apply_update=False, loss_out=loss_out, global_batch=global_batch,
                        slot=slot, batch=batch, defer_reduce=True)
        self.apply_optimizer(optimizer, lr, trainable, **kw)

    def apply_optimizer(self, kind, lr, trainable, beta_1=0.9, beta_2=0.999, rho=0.9, eps=1e-7):
        FakeEngine.calls.append((kind, lr, frozenset(trainable)))
        t = self.steps.get(kind, 0) + 1
        self.steps[kind] = t
        for k, g in self.grads.items():
            if k.split("/")[0] not in trainable:
                continue
            if kind == "adam":
                m, v = self.state.get(("m", k), 0.0), self.state.get(("v", k), 0.0)
                self.params[k], self.state[("m", k)], self.state[("v", k)] = oracle.adam_step(
                    self.params[k], g, m, v, t, lr, beta_1, beta_2, eps)
            else:
                r = self.state.get(("r", k), 0.0)
                self.params[k], self.state[("r", k)] = oracle.rmsprop_step(self.params[k], g, r, lr, rho, eps)

"""Host-side mirror of the reference's model-definition API for cutoffvgg16.

``CutoffVGG16`` has the constructor, ``fit`` (two-phase extract/finetune schedule), ``evaluate``,
``predict``, ``metrics_names`` and ``.model`` of src/models/models.py:387-444.  ``.model`` is a
``B200Model``: the stand-in for the inner ``tf.keras.Model`` (compile / fit / evaluate / predict /
__call__ / layers / get_layer / save) whose numerics run on the CutoffVGG16Engine.
"""
import os

import numpy as np
import torch

from . import metrics as M
from .engine import CutoffVGG16Engine


class Layer:
    """Stand-in for a Keras layer handle: the reference only toggles ``.trainable`` and reads ``.name``."""

    def __init__(self, name, kind, trainable=True):
        self.name, self.kind, self.trainable = name, kind, trainable
        # gradcam.py:35-36 finds the last conv by scanning ``layer._keras_api_names``
        self._keras_api_names = {"conv": ("keras.layers.Conv2D",), "pool": ("keras.layers.MaxPool2D",),
                                 "gap": ("keras.layers.GlobalAveragePooling2D",), "dropout": ("keras.layers.Dropout",),
                                 "dense": ("keras.layers.Dense",), "input": ("keras.layers.InputLayer",)}[kind]

    def __repr__(self):
        return f"<Layer {self.name} ({self.kind}) trainable={self.trainable}>"


class _Optimizer:
    kind = "?"

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)

    @property
    def lr(self):
        return self.learning_rate

    @lr.setter
    def lr(self, v):
        self.learning_rate = float(v)


class Adam(_Optimizer):
    """keras.optimizers.Adam defaults (beta_1 .9, beta_2 .999, epsilon 1e-7); models.py:397."""
    kind = "adam"

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon


class RMSprop(_Optimizer):
    """keras.optimizers.RMSprop defaults (rho .9, momentum 0, epsilon 1e-7, not centered); models.py:398."""
    kind = "rmsprop"

    def __init__(self, learning_rate=0.001, rho=0.9, epsilon=1e-7):
        super().__init__(learning_rate)
        self.rho, self.epsilon = rho, epsilon


class History:
    def __init__(self):
        self.epoch, self.history = [], {}


class Callback:
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def on_train_batch_end(self, batch, logs=None): pass


class EarlyStopping(Callback):
    """keras.callbacks.EarlyStopping as configured at train.py:155-156."""

    def __init__(self, monitor="val_loss", verbose=0, patience=0, mode="min", restore_best_weights=False, min_delta=0.0):
        self.monitor, self.verbose, self.patience, self.mode = monitor, verbose, patience, mode
        self.restore_best_weights, self.min_delta = restore_best_weights, abs(min_delta)

    def _better(self, cur, best):
        return cur < best - self.min_delta if self.mode == "min" else cur > best + self.min_delta

    def on_train_begin(self, logs=None):
        self.wait, self.stopped_epoch, self.best_weights = 0, 0, None
        self.best = np.inf if self.mode == "min" else -np.inf

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.engine.theta.clone()
        self.wait += 1
        if self._better(cur, self.best):
            self.best, self.wait = cur, 0
            if self.restore_best_weights:
                self.best_weights = self.model.engine.theta.clone()
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                self.model.engine.theta.copy_(self.best_weights)
                self.model.engine.refresh_operands()


class ReduceLROnPlateau(Callback):
    """keras.callbacks.ReduceLROnPlateau as configured at train.py:158-159."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, min_lr=0.0, min_delta=1e-4, cooldown=0):
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_lr, self.min_delta, self.cooldown = min_lr, min_delta, cooldown

    def on_train_begin(self, logs=None):
        self.best, self.wait, self.cooldown_counter = np.inf, 0, 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        logs["lr"] = self.model.optimizer.learning_rate
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if cur < self.best - self.min_delta:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = self.model.optimizer.learning_rate
                if old > np.float32(self.min_lr):
                    self.model.optimizer.learning_rate = max(old * self.factor, self.min_lr)
                    self.cooldown_counter, self.wait = self.cooldown, 0


def _iter_batches(data):
    """A dataset is anything that yields (x, y) or x batches each time it is iterated (the tf.data
    contract the reference relies on, preprocessor.py:36-66), or a callable returning such an iterable."""
    it = data() if callable(data) else data
    for item in it:
        if isinstance(item, (tuple, list)) and len(item) >= 2:
            yield item[0], item[1]
        else:
            yield item, None


def _cat_labels(ys):
    """Per-batch one-hot labels (numpy, pinned or device tensors) -> one float32 numpy array."""
    if ys and torch.is_tensor(ys[0]):
        return torch.cat([y.float() for y in ys]).cpu().numpy()
    return np.concatenate([np.asarray(y, np.float32) for y in ys], 0)


class B200Model:
    """Inner model (stand-in for the ``tf.keras.Model`` built at models.py:407-421)."""

    def __init__(self, engine, input_shape, n_classes, dropout):
        self.engine = engine
        self.input_shape_, self.n_classes, self.dropout = tuple(input_shape), n_classes, dropout
        self.layers = [Layer("input", "input", False)]
        for s in engine.specs:
            self.layers.append(Layer(s.name, "conv"))
            if s.pool:
                self.layers.append(Layer(s.name.replace("conv" + s.name[-1], "pool"), "pool", False))
        self.layers += [Layer("global_avgpool", "gap", False), Layer("dropout", "dropout", False),
                        Layer("output", "dense")]
        self.optimizer = None
        self.metric_objs = []
        self.stop_training = False
        self.history = None

    # ---- Keras-like surface
    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}")

    def summary(self, print_fn=print):
        print_fn(f"B200Model cutoffvgg16: {len(self.engine.specs)} conv layers, {self.engine.n_params:,} parameters")
        for l in self.layers:
            print_fn(f"  {l.name:16s} {l.kind:8s} trainable={l.trainable}")

    def compile(self, optimizer=None, loss="categorical_crossentropy", metrics=None, run_eagerly=None):
        if loss != "categorical_crossentropy":
            raise NotImplementedError("only loss='categorical_crossentropy' (models.py:426,431) is implemented")
        if not isinstance(optimizer, _Optimizer):
            raise TypeError("optimizer must be ab_line_classifier_b200.model.Adam or RMSprop")
        self.optimizer = optimizer
        self.metric_objs = M.resolve(metrics)

    @property
    def metrics_names(self):
        return ["loss"] + [m.name for m in self.metric_objs]

    def trainable_names(self):
        return {l.name for l in self.layers if l.kind in ("conv", "dense") and l.trainable}

    def get_weights(self):
        return self.engine.get_weights()

    def set_weights(self, weights):
        self.engine.set_weights(weights)

    def __call__(self, x, training=False):
        """model(x) -> softmax probabilities as numpy [B,n] (predict.py:440, train_utils.py:288)."""
        x = np.asarray(x, np.float32) if not torch.is_tensor(x) else x
        return self.engine.predict_batch(x).cpu().numpy()

    def predict(self, test_data, verbose=1, batch_size=512, return_device=False):
        """Keras Model.predict: concatenated softmax probabilities [N, n_classes] (numpy; return_device=True keeps the
        result in HBM for a following device-side step such as the clip aggregation)."""
        outs = []
        if isinstance(test_data, np.ndarray) or torch.is_tensor(test_data):
            test_data = [(test_data[i:i + batch_size], None) for i in range(0, len(test_data), batch_size)]
        eng = self.engine
        it = _iter_batches(test_data)
        nxt = next(it, None)
        slot = None
        while nxt is not None:
            x = nxt[0]
            B = len(x)
            if slot is None:                                   # nothing staged ahead (first batch, or the buffers grew)
                eng.reserve(B, False)
                slot = eng.stage(x)
            cur = slot
            nxt, slot = next(it, None), None
            if nxt is not None and len(nxt[0]) <= eng._cap:
                slot = eng.stage(nxt[0], slot=cur ^ 1)         # H2D copy of the next batch overlaps this forward pass
            eng.use_slot(cur)
            outs.append(eng.forward_device(B, need_last_act=False).clone())   # stays on device: one D2H at the end
            eng.release_slot()
        if not outs:
            return (torch.zeros(0, self.n_classes, device=self.engine.device) if return_device
                    else np.zeros((0, self.n_classes), np.float32))
        return torch.cat(outs) if return_device else torch.cat(outs).cpu().numpy()

    def _run_eval(self, data):
        eng = self.engine
        eng._ensure(1, False)
        loss = torch.zeros(1, device=eng.device, dtype=torch.float32)
        plan = M.device_plan(self.metric_objs, self.n_classes)
        if plan is not None:
            eng.metrics_begin(*plan)          # Keras metric state accumulates on the device (abl_metrics_accumulate)
        probs, ys, n = [], [], 0
        for x, y in _iter_batches(data):
            if y is None:
                raise ValueError("evaluate needs (x, y) batches")
            n += len(x)
            p = eng.eval_batch(x, y, loss_out=loss, metrics=plan is not None)
            if plan is None:                  # a metric the counters cannot express: collect arrays for result()
                probs.append(p.clone())
                ys.append(y.clone() if torch.is_tensor(y) else np.array(y, np.float32))   # batch buffers are recycled
        # data parallel: every rank evaluated its own shard (or a replica) -- loss and metrics are those of all frames
        loss_sum, n = eng.sum_over_ranks([float(loss.item()), n])
        if n == 0:
            return [0.0] + [0.0] * len(self.metric_objs)
        res = [loss_sum / n]
        if plan is not None:
            counts = M.MetricCounts(eng.metrics_read(), self.n_classes, *plan)
            res += [m.from_counts(counts) for m in self.metric_objs]
        else:
            p, yt = torch.cat(probs).cpu().numpy(), _cat_labels(ys)
            res += [m.result(yt, p) for m in self.metric_objs]
        return res

    def evaluate(self, test_data, verbose=1):
        res = self._run_eval(test_data)
        if verbose:
            print(" - ".join(f"{n}: {v:.4f}" for n, v in zip(self.metrics_names, res)))
        return res

    def train_on_batch(self, x, y, class_weight=None):
        """Keras ``Model.train_on_batch``: H2D copy of (x, y), one optimizer step, returns the batch loss (D2H)."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model before train_on_batch()")
        eng, opt = self.engine, self.optimizer
        if class_weight is not None:
            eng.set_class_weight(class_weight)
        trainable = self.trainable_names()
        kw = ({"beta_1": opt.beta_1, "beta_2": opt.beta_2, "eps": opt.epsilon} if opt.kind == "adam"
              else {"rho": opt.rho, "eps": opt.epsilon})
        loss = eng.train_step(x, y, trainable, apply_update=False, global_batch=len(x) * eng.world_size,
                              defer_reduce=True)
        eng.apply_optimizer(opt.kind, opt.learning_rate, trainable, **kw)
        return float(loss.item())

    def fit(self, train_data, epochs=1, initial_epoch=0, validation_data=None, callbacks=None, verbose=1,
            class_weight=None):
        if self.optimizer is None:
            raise RuntimeError("compile() the model before fit()")
        eng = self.engine
        eng.set_class_weight(class_weight)
        trainable = self.trainable_names()
        hist = History()
        self.history = hist
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        self.stop_training = False
        for cb in callbacks:
            cb.on_train_begin() if hasattr(cb, "on_train_begin") else None
        # per-step losses: written by the step into a device ring and copied to pinned host memory right behind it
        # (4 bytes, asynchronous), so the host never waits for a step inside the epoch
        ring = torch.zeros(1024, device=eng.device, dtype=torch.float32)
        on_gpu = torch.device(eng.device).type == "cuda"
        ring_host = torch.zeros(1024, dtype=torch.float32)
        if on_gpu:
            ring_host = ring_host.pin_memory()
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch) if hasattr(cb, "on_epoch_begin") else None
            opt = self.optimizer
            tot_loss, n_seen, slot_i = 0.0, 0, 0
            sizes, probs, ys = [], [], []
            plan = M.device_plan(self.metric_objs, self.n_classes) if self.metric_objs else None
            if plan is not None:
                eng.metrics_begin(*plan)

            def drain():
                nonlocal tot_loss, slot_i
                if slot_i:
                    if on_gpu:
                        torch.cuda.current_stream().synchronize()
                    vals = ring_host[:slot_i].numpy()
                    tot_loss += float(np.dot(vals, np.asarray(sizes[-slot_i:], np.float64)))
                    slot_i = 0

            kw = ({"beta_1": opt.beta_1, "beta_2": opt.beta_2, "eps": opt.epsilon} if opt.kind == "adam"
                  else {"rho": opt.rho, "eps": opt.epsilon})
            # software pipeline: the H2D copy of batch i+1 (copy stream) overlaps the kernels of batch i
            batches = _iter_batches(train_data)
            nxt = next(batches, None)
            nslot = None
            while nxt is not None:
                (x, y) = nxt
                B = len(x)
                if nslot is None:                                  # nothing staged ahead (first batch, or the buffers grew)
                    eng.reserve(B, True)
                    nslot = eng.stage(x, y)
                slot, nslot = nslot, None
                nxt = next(batches, None)
                if nxt is not None and len(nxt[0]) <= eng._cap:
                    # a LARGER next batch would re-allocate the input slots under the staged one: it is staged after
                    # this step instead (reserve + stage at the top of the loop)
                    nslot = eng.stage(nxt[0], nxt[1], slot=slot ^ 1)
                eng.fit_step(trainable, opt.kind, opt.learning_rate, slot, B, B * eng.world_size,
                             loss_out=ring[slot_i:slot_i + 1], **kw)
                ring_host[slot_i:slot_i + 1].copy_(ring[slot_i:slot_i + 1], non_blocking=True)
                if plan is not None:
                    eng.metrics_update(B)      # training-batch outputs (dropout on), as Keras reports them
                elif self.metric_objs:
                    probs.append(eng._buf["probs"][:B].clone())
                    ys.append(y.clone() if torch.is_tensor(y) else np.array(y, np.float32))   # batch buffers are recycled
                sizes.append(B)
                n_seen += B
                slot_i += 1
                if slot_i == ring.numel():
                    drain()
            drain()
            # each step's loss is sum_local(w_i l_i) / GLOBAL batch, so under data parallel the ranks' epoch sums add
            # up to the sum of the global-batch means (what single-device Keras logs)
            tot_loss, n_glob = eng.sum_over_ranks([tot_loss, n_seen])
            logs = {"loss": tot_loss * eng.world_size / max(n_glob, 1)}
            if plan is not None and n_glob:
                counts = M.MetricCounts(eng.metrics_read(), self.n_classes, *plan)
                for m in self.metric_objs:
                    logs[m.name] = m.from_counts(counts)
            elif self.metric_objs and probs:
                p, yt = torch.cat(probs).cpu().numpy(), _cat_labels(ys)
                for m in self.metric_objs:
                    logs[m.name] = m.result(yt, p)
            if validation_data is not None:
                for n, v in zip(self.metrics_names, self._run_eval(validation_data)):
                    logs["val_" + n] = v
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            for cb in callbacks:
                cb.on_epoch_end(epoch, logs) if hasattr(cb, "on_epoch_end") else None
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - " + " - ".join(f"{k}: {v:.4g}" for k, v in logs.items()))
            if self.stop_training:
                break
        for cb in callbacks:
            cb.on_train_end() if hasattr(cb, "on_train_end") else None
        return hist

    # ---- checkpoint interchange (Keras names and layouts: .npz, or Keras' legacy HDF5 layout through h5lite)
    def save(self, path, include_optimizer=True):
        save_model(self, path, include_optimizer=include_optimizer)


def _is_hdf5(path):
    with open(path, "rb") as fh:
        return fh.read(8) == b"\x89HDF\r\n\x1a\n"


def _inner(model):
    return model.model if hasattr(model, "model") and not isinstance(model, B200Model) else model


def _keras_model_config(inner):
    """model_config JSON in the shape Keras' functional API writes (class names, layer names, the arguments that define
    the architecture).  Enough for load_model() here to rebuild the engine; not a byte-for-byte Keras config."""
    H, W, C = (int(v) for v in inner.input_shape_)
    layers = [{"class_name": "InputLayer", "name": "input",
               "config": {"name": "input", "batch_input_shape": [None, H, W, C], "dtype": "float32"}}]
    for l in inner.layers[1:]:
        if l.kind == "conv":
            spec = next(s for s in inner.engine.specs if s.name == l.name)
            cfg = {"name": l.name, "filters": spec.cout, "kernel_size": [3, 3], "strides": [1, 1], "padding": "same",
                   "activation": "relu", "use_bias": True, "trainable": bool(l.trainable)}
            layers.append({"class_name": "Conv2D", "name": l.name, "config": cfg})
        elif l.kind == "pool":
            layers.append({"class_name": "MaxPooling2D", "name": l.name,
                           "config": {"name": l.name, "pool_size": [2, 2], "strides": [2, 2], "padding": "valid"}})
        elif l.kind == "gap":
            layers.append({"class_name": "GlobalAveragePooling2D", "name": l.name, "config": {"name": l.name}})
        elif l.kind == "dropout":
            layers.append({"class_name": "Dropout", "name": l.name, "config": {"name": l.name, "rate": float(inner.dropout)}})
        elif l.kind == "dense":
            layers.append({"class_name": "Dense", "name": l.name,
                           "config": {"name": l.name, "units": int(inner.n_classes), "activation": "softmax",
                                      "use_bias": True, "trainable": bool(l.trainable)}})
    return {"class_name": "Functional", "config": {"name": "model", "layers": layers,
                                                    "input_layers": [["input", 0, 0]], "output_layers": [["output", 0, 0]]}}


def _optimizer_state(eng):
    """Everything a resumed run needs besides the weights: optimizer slots (flat, engine layout), their step counts and
    the dropout step counter.  (The reference saves no optimizer state at all -- train.py:139-144 -- so a resumed Keras
    run restarts Adam's moments; here resume is exact.)"""
    # dropout masks are a counter hash of (seed, step, rank, frame, channel): the seed belongs to the resumable state
    st = {"step_count": np.asarray(eng.step_count, np.int64), "seed": np.asarray(getattr(eng, "seed", 0), np.int64)}
    for k, v in getattr(eng, "opt_state", {}).items():
        st["slot_" + k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in getattr(eng, "opt_steps", {}).items():
        st["steps_" + k] = np.asarray(v, np.int64)
    return st


def _restore_optimizer_state(eng, st):
    eng.step_count = int(st.get("step_count", 0))
    if "seed" in st:
        eng.seed = int(st["seed"])
    for k, v in st.items():
        if k.startswith("slot_"):
            t = torch.from_numpy(np.ascontiguousarray(v, np.float32))
            eng.opt_state[k[5:]] = t.to(eng.device) if torch.device(eng.device).type == "cuda" else t
        elif k.startswith("steps_"):
            eng.opt_steps[k[6:]] = int(v)


def save_model(model, path, include_optimizer=True):
    """Counterpart of ``tf.keras.models.save_model(model.model, path)`` (train.py:139-144): weights under Keras
    names/layouts + the architecture config [+ optimizer state, as Keras' include_optimizer=True].
    ``*.h5`` / ``*.hdf5``: Keras' legacy HDF5 layout (``model_weights/<layer>/<layer>/kernel:0`` datasets, ``layer_names``
    / ``weight_names`` attributes, ``model_config``), written by h5lite; anything else: ``.npz``."""
    inner = _inner(model)
    w = inner.get_weights()
    eng = inner.engine
    opt = _optimizer_state(eng) if include_optimizer else {}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    if str(path).endswith((".h5", ".hdf5")):
        import json
        from .h5lite import H5Writer
        wr = H5Writer()
        wr.attr(wr.root, "backend", b"ab_line_classifier_b200")
        wr.attr(wr.root, "keras_version", b"2.9.0")
        wr.attr(wr.root, "model_config", json.dumps(_keras_model_config(inner)).encode("utf8"))
        mw = wr.group(wr.root, "model_weights")
        names = [l.name for l in inner.layers]
        wr.attr(mw, "layer_names", np.array([n.encode("utf8") for n in names]))
        wr.attr(mw, "backend", b"ab_line_classifier_b200")
        wr.attr(mw, "keras_version", b"2.9.0")
        for l in inner.layers:
            g = wr.group(mw, l.name)
            mine = [k for k in w if k.split("/")[0] == l.name]
            if mine:
                wr.attr(g, "weight_names", np.array([(k + ":0").encode("utf8") for k in mine]))
                for k in mine:       # Keras nests by the '/' in the variable name: <layer>/<layer>/kernel:0
                    sub = wr.group(g, l.name)
                    wr.dataset(sub, k.split("/", 1)[1] + ":0", w[k])
        if opt:
            og = wr.group(wr.root, "b200_optimizer_state")
            for k, v in opt.items():
                wr.dataset(og, k, np.atleast_1d(v))
        wr.save(path)
        return path
    cfg = dict(input_shape=np.asarray(inner.input_shape_), n_classes=inner.n_classes, dropout=inner.dropout,
               cutoff_layer=eng.cutoff_layer)
    if not str(path).endswith(".npz"):
        path = str(path) + ".npz"
    np.savez(path, **{"weights/" + k: v for k, v in w.items()}, **{"config/" + k: v for k, v in cfg.items()},
             **{"optimizer/" + k: v for k, v in opt.items()})
    return path


def read_keras_h5(path):
    """(weights dict, model_config dict or None, optimizer-state dict) from a Keras HDF5 file: a full model
    (``save_model``: weights below ``model_weights``) or a weights file (``save_weights`` / the keras-applications
    ``vgg16_weights_tf_dim_ordering_tf_kernels_notop.h5``: layer groups at the root).  Variables are keyed
    ``<layer>/kernel`` / ``<layer>/bias`` whatever the file calls them (``kernel:0``, legacy ``block1_conv1_W_1:0``):
    within a layer the array with more than one axis is the kernel, the vector is the bias."""
    import json
    from .h5lite import H5File
    f = H5File(path)
    g = f["model_weights"] if "model_weights" in f else f
    layer_names = g.attrs.get("layer_names")
    layers = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.asarray(layer_names).ravel()] \
        if layer_names is not None else g.keys()
    weights = {}
    for ln in layers:
        if ln not in g:
            continue
        lg = g[ln]
        if not lg.is_group:
            continue
        wn = lg.attrs.get("weight_names")
        if wn is not None:
            paths = [n.decode("utf8") if isinstance(n, bytes) else str(n) for n in np.asarray(wn).ravel()]
            arrays = [lg[p].read() for p in paths]
        else:
            arrays = [d.read() for _, d in lg.visit_datasets()]
        for a in arrays:
            key = ln + ("/kernel" if a.ndim > 1 else "/bias")
            if key in weights:
                raise ValueError(f"{path}: layer {ln} holds more than one {key.split('/')[1]} (not a Conv2D / Dense layer)")
            weights[key] = np.asarray(a, np.float32)
    cfg = f.attrs.get("model_config")
    if cfg is not None:
        cfg = json.loads(cfg.decode("utf8") if isinstance(cfg, bytes) else cfg)
    opt = {}
    if "b200_optimizer_state" in f:
        for k, d in f["b200_optimizer_state"].visit_datasets():
            v = d.read()
            opt[k] = v if k.startswith("slot_") else v.reshape(-1)[0]      # scalars were stored as 1-element arrays
    return weights, cfg, opt


def _arch_from_keras(weights, cfg, input_shape=None):
    """(input_shape, n_classes, cutoff_layer, dropout) of a cutoffvgg16 model from its Keras config / weight shapes."""
    from .engine import VGG16_LAYERS
    names = [n for n, _, _ in VGG16_LAYERS]
    dropout, layer_list = 0.45, None
    if cfg is not None:
        layer_list = [l.get("name", l.get("config", {}).get("name")) for l in cfg.get("config", {}).get("layers", [])]
        for l in cfg.get("config", {}).get("layers", []):
            c = l.get("config", {})
            if l.get("class_name") == "Dropout":
                dropout = float(c.get("rate", dropout))
            if l.get("class_name") == "InputLayer" and c.get("batch_input_shape") and input_shape is None:
                input_shape = tuple(int(v) for v in c["batch_input_shape"][1:])
    present = layer_list if layer_list is not None else [k.split("/")[0] for k in weights]
    kept = [i for i, n in enumerate(names) if n in present]
    if not kept or "output/kernel" not in weights:
        raise ValueError("not a cutoffvgg16 model: no VGG16 block layers / no 'output' Dense layer in the file")
    cutoff = max(kept) + 2            # vgg16.layers[1:cutoff] keeps VGG16_LAYERS[:cutoff - 1]
    if input_shape is None:
        input_shape = (128, 128, 3)   # config.yml DATA.IMG_DIM default of the reference
    return tuple(input_shape), int(weights["output/kernel"].shape[1]), cutoff, dropout


def load_model(path, compile=False, device=None, input_shape=None):
    """Counterpart of ``load_model(model_path, compile=False)`` (predict.py:50): ``.npz`` files written by save_model, or
    a Keras HDF5 model / weights file (the format the reference's train.py writes)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Could not find model weights at file path: {path}")
    if _is_hdf5(path):
        weights, cfg, opt = read_keras_h5(path)
        ishape, n_classes, cutoff, dropout = _arch_from_keras(weights, cfg, input_shape)
    else:
        z = np.load(path)
        cfg = {k[len("config/"):]: z[k] for k in z.files if k.startswith("config/")}
        ishape, n_classes = tuple(int(v) for v in cfg["input_shape"]), int(cfg["n_classes"])
        cutoff, dropout = int(cfg["cutoff_layer"]), float(cfg["dropout"])
        weights = {k[len("weights/"):]: z[k] for k in z.files if k.startswith("weights/")}
        opt = {k[len("optimizer/"):]: z[k] for k in z.files if k.startswith("optimizer/")}
    eng = CutoffVGG16Engine(ishape, n_classes, cutoff, dropout, device=device)
    model = B200Model(eng, eng.input_shape, eng.n_classes, eng.dropout)
    model.set_weights(weights)
    if opt:
        _restore_optimizer_state(eng, opt)     # resume: Adam / RMSprop slots, step counts, dropout counter
    return model


def _stored_weights(weights_path):
    if _is_hdf5(weights_path):
        return read_keras_h5(weights_path)[0]
    z = np.load(weights_path)
    return {k[len("weights/"):]: z[k] for k in z.files if k.startswith("weights/")}


def initialize_with_pretrained_weights(model, weights_path):
    """model_utils.py:6-23: copy weights by layer name from a saved model (.npz or Keras .h5) into ``model``."""
    assert os.path.exists(weights_path), f"Could not find model weights at file path: {weights_path}"
    have = {(k[:-2] if k.endswith(":0") else k): v for k, v in _stored_weights(weights_path).items()}
    cur = model.get_weights()
    for k, v in have.items():
        if k in cur and cur[k].shape == v.shape:
            cur[k] = v
        else:
            print(f"{k.split('/')[0]} is not in the new model.")
    model.set_weights(cur)
    return model


class CutoffVGG16:
    """Drop-in for the reference class of the same name (src/models/models.py:387-444)."""

    def __init__(self, model_config, input_shape, metrics, n_classes, mixed_precision=False, output_bias=None,
                 weights_path=None, device=None, seed=10001, rank=0, world_size=1, process_group=None,
                 vgg16_weights=None):
        self.lr_extract = model_config['LR_EXTRACT']
        self.lr_finetune = model_config['LR_FINETUNE']
        self.dropout = model_config['DROPOUT']
        self.cutoff_layer = model_config['CUTOFF_LAYER']
        self.finetune_layer = model_config['FINETUNE_LAYER']
        self.extract_epochs = model_config['EXTRACT_EPOCHS']
        self.optimizer_extract = Adam(learning_rate=self.lr_extract)
        self.optimizer_finetune = RMSprop(learning_rate=self.lr_finetune)
        if output_bias is not None and hasattr(output_bias, "value"):   # keras Constant initializer
            output_bias = output_bias.value
        self.output_bias = output_bias
        self.n_classes = n_classes
        self.input_shape = list(input_shape)
        self.metrics = metrics
        self.mixed_precision = mixed_precision     # stored and unused, as in the reference (models.py:403)
        self.weights_path = weights_path
        self._engine_kw = dict(device=device, seed=seed, rank=rank, world_size=world_size, process_group=process_group)
        # the reference builds on VGG16(weights='imagenet') (models.py:409, a download); offline, point this at the
        # keras-applications file vgg16_weights_tf_dim_ordering_tf_kernels_notop.h5 (or its .npz conversion)
        self.vgg16_weights = vgg16_weights if vgg16_weights is not None else os.environ.get("ABL_VGG16_WEIGHTS")
        self.model = self.define_model()

    def define_model(self):
        eng = CutoffVGG16Engine(tuple(self.input_shape), self.n_classes, self.cutoff_layer, self.dropout,
                                output_bias=self.output_bias, **self._engine_kw)
        model = B200Model(eng, self.input_shape, self.n_classes, self.dropout)
        # vgg16.layers[1:cutoff_layer] (models.py:410): conv AND pool layers, in Keras order
        self.vgg16_layers = [l for l in model.layers if l.kind in ("conv", "pool")]
        if self.vgg16_weights:
            # VGG16(weights='imagenet'): the convolutional base starts from the published filters (by layer name)
            have = _stored_weights(self.vgg16_weights)
            cur = model.get_weights()
            missing = [k for k in cur if k.split("/")[0] != "output" and (k not in have or have[k].shape != cur[k].shape)]
            if missing:
                raise ValueError(f"{self.vgg16_weights}: no weights of the right shape for {missing}")
            cur.update({k: have[k] for k in cur if k.split("/")[0] != "output"})
            model.set_weights(cur)
        elif self.weights_path is None:
            import warnings
            warnings.warn("CutoffVGG16: no pretrained VGG16 weights supplied (ABL_VGG16_WEIGHTS / vgg16_weights=...): the "
                          "convolutional base is Glorot-initialised, whereas the reference starts from "
                          "VGG16(weights='imagenet') (models.py:409).  Phase 1 then trains the head on random features.",
                          stacklevel=3)
        if self.weights_path is not None:
            model = initialize_with_pretrained_weights(model, self.weights_path)
        return model

    def fit(self, train_data, epochs=1, validation_data=None, callbacks=None, verbose=1, class_weight=None):
        for layer in self.vgg16_layers:
            layer.trainable = False
        self.model.compile(optimizer=self.optimizer_extract, loss='categorical_crossentropy', metrics=self.metrics,
                           run_eagerly=True)
        history_extract = self.model.fit(train_data, epochs=self.extract_epochs, validation_data=validation_data,
                                         callbacks=callbacks, verbose=verbose, class_weight=class_weight)
        for layer in self.vgg16_layers[self.finetune_layer:]:
            layer.trainable = True
        self.model.compile(optimizer=self.optimizer_finetune, loss='categorical_crossentropy', metrics=self.metrics,
                           run_eagerly=True)
        history_finetune = self.model.fit(train_data, epochs=epochs, initial_epoch=history_extract.epoch[-1],
                                          validation_data=validation_data, callbacks=callbacks, verbose=verbose,
                                          class_weight=class_weight)
        self.history_extract, self.history_finetune = history_extract, history_finetune

    def evaluate(self, test_data, verbose=1):
        return self.model.evaluate(test_data, verbose=verbose)

    def predict(self, test_data, verbose=1):
        return self.model.predict(test_data, verbose=verbose)

    @property
    def metrics_names(self):
        return self.model.metrics_names

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

    # ---- checkpoint interchange (Keras names and layouts; .npz because h5py is absent here)
    def save(self, path):
        save_model(self, path)


def save_model(model, path):
    """Counterpart of ``tf.keras.models.save_model(model.model, path)`` (train.py:139-144): weights under
    Keras names/layouts + the architecture config, as .npz."""
    inner = model.model if hasattr(model, "model") and not isinstance(model, B200Model) else model
    w = inner.get_weights()
    cfg = dict(input_shape=np.asarray(inner.input_shape_), n_classes=inner.n_classes, dropout=inner.dropout,
               cutoff_layer=inner.engine.cutoff_layer)
    if not str(path).endswith(".npz"):
        path = str(path) + ".npz"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    np.savez(path, **{"weights/" + k: v for k, v in w.items()}, **{"config/" + k: v for k, v in cfg.items()})
    return path


def load_model(path, compile=False, device=None):
    """Counterpart of ``load_model(model_path, compile=False)`` (predict.py:50) for files written by save_model."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Could not find model weights at file path: {path}")
    z = np.load(path)
    cfg = {k[len("config/"):]: z[k] for k in z.files if k.startswith("config/")}
    eng = CutoffVGG16Engine(tuple(int(v) for v in cfg["input_shape"]), int(cfg["n_classes"]), int(cfg["cutoff_layer"]),
                            float(cfg["dropout"]), device=device)
    model = B200Model(eng, eng.input_shape, eng.n_classes, eng.dropout)
    model.set_weights({k[len("weights/"):]: z[k] for k in z.files if k.startswith("weights/")})
    return model


def initialize_with_pretrained_weights(model, weights_path):
    """model_utils.py:6-23: copy weights by layer name from a saved model into ``model``."""
    assert os.path.exists(weights_path), f"Could not find model weights at file path: {weights_path}"
    z = np.load(weights_path)
    have = {k[len("weights/"):]: z[k] for k in z.files if k.startswith("weights/")}
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
                 weights_path=None, device=None, seed=10001, rank=0, world_size=1, process_group=None):
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
        self.model = self.define_model()

    def define_model(self):
        eng = CutoffVGG16Engine(tuple(self.input_shape), self.n_classes, self.cutoff_layer, self.dropout,
                                output_bias=self.output_bias, **self._engine_kw)
        model = B200Model(eng, self.input_shape, self.n_classes, self.dropout)
        # vgg16.layers[1:cutoff_layer] (models.py:410): conv AND pool layers, in Keras order
        self.vgg16_layers = [l for l in model.layers if l.kind in ("conv", "pool")]
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

"""``Session`` shim: the three ``tf.Session.run`` idioms the meta-learner uses, dispatched to the C ABI.

  sess.run(model.minimize_op, feed_dict={input_ph, label_ph[, lr_ph, drop_rate_ph]})   reptile.py:115-121, :269-279
  sess.run(model.predictions, feed_dict={input_ph[, is_training_ph: False]})           reptile.py:503-506, :520
  sess.run(pre_step_op)                                                                 reptile.py:112-113
Every call copies its feed to the device, runs on slot 0 of the model's engine and (for predictions) copies
the result back - the same crossings the reference makes.  The throughput path is runner.TaskRunner.
"""
from __future__ import annotations

import numpy as np
import torch

from .efficientlab import EfficientLab, Handle


class WeightDecayOp(Handle):
    """variables.weight_decay(rate): var <- var * rate over all trainables (meta_learners/variables.py:48-55)."""

    def __init__(self, rate: float):
        super().__init__("weight_decay", "op")
        self.rate = float(rate)


class Session:
    def __init__(self, model: EfficientLab = None):
        self.model = model
        self.graph = None
        self._step = 0
        self._pending_decay = 1.0

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def bind(self, model: EfficientLab) -> "Session":
        self.model = model
        return self

    @staticmethod
    def _batch(x, channels):
        a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
        if a.ndim != 4 or a.shape[-1] != channels:
            raise ValueError("expected a batch of [H,W,%d] arrays, got shape %s" % (channels, a.shape))
        return a

    def run(self, fetches, feed_dict=None):
        m = self.model
        if m is None:
            raise RuntimeError("Session is not bound to a model")
        feed = feed_dict or {}
        if isinstance(fetches, WeightDecayOp):
            # applied by the scale kernel at the head of the next mliis_train_step (nothing runs in between:
            # reptile.py:112-121), so no torch arithmetic touches the parameters
            self._pending_decay *= fetches.rate
            return None
        if fetches is m.minimize_op:
            eng = m.engine()
            x = torch.from_numpy(self._batch(feed[m.input_ph], 3)).to(eng.device, non_blocking=True)
            y = torch.from_numpy(self._batch(feed[m.label_ph], 2)).to(eng.device, non_blocking=True)
            lr = float(feed.get(m.lr_ph, m.lr_ph.default))
            if isinstance(m.final_layer_dropout_rate_ph, Handle) and m.final_layer_dropout_rate_ph in feed:
                if abs(float(feed[m.final_layer_dropout_rate_ph]) - m.final_layer_dropout_rate) > 1e-12:
                    raise NotImplementedError("feeding a dropout rate different from --final_layer_dropout_rate")
            if x.shape[0] > eng.max_batch:
                raise ValueError("batch %d exceeds max_batch %d" % (x.shape[0], eng.max_batch))
            self._step += 1
            eng.train_step(0, x, y, lr, seed=self._step, pre_decay_rate=self._pending_decay)
            self._pending_decay = 1.0
            return None
        if fetches is m.predictions:
            eng = m.engine()
            training = feed.get(m.is_training_ph, m.is_training_ph.default)
            if training:
                raise NotImplementedError("predictions with is_training=True (batch statistics) is not built; the "
                                          "meta-learner always feeds is_training_ph: False (eval.py:61-69)")
            x = torch.from_numpy(self._batch(feed[m.input_ph], 3)).to(eng.device, non_blocking=True)
            out = []
            for i in range(0, x.shape[0], eng.max_batch):      # the reference has no batch limit
                pred, _, _, _ = eng.predict(0, x[i:i + eng.max_batch].contiguous())
                out.append(pred)
            return torch.cat(out, 0).cpu().numpy()
        raise NotImplementedError("Session.run(%r) is not part of the hot path" % (fetches,))

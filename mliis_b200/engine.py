"""Device-side owner of the EfficientLab state for the B200 engine.

PyTorch is used for buffer ownership, streams and host<->device copies only; every numeric step is a call
through the C ABI (mliis_b200/native.py -> libmliis_b200.so).  Replaces the role of ``tf.Session`` + the TF
variable store in the reference (run_metasegnet.py:109; meta_learners/variables.py:58-80).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import native as N

ADAM_BETA1, ADAM_BETA2 = 0.0, 0.999


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Engine:
    """One engine per (process, GPU).  ``n_slots`` independent task slots share the kernels."""

    def __init__(self, image_size: int = 224, max_batch: int = 8, n_slots: int = 1, sgd: bool = False,
                 dice: bool = True, l2: bool = True, label_smoothing: float = 0.0, final_dropout_rate: float = 0.0,
                 rsd: Sequence[int] = (2, 4), gemm_mode: int = N.GEMM_FP32, device: int = 0, n_classes: int = 1,
                 staging_bytes: Optional[int] = None):
        if not torch.cuda.is_available():
            raise N.MliisError(N.MLIIS_ERR_DEVICE, "no CUDA device: mliis_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        flags = (N.LOSS_DICE if dice else 0) | (N.LOSS_L2 if l2 else 0)
        self.cfg = N.make_config(image_size, max_batch, n_slots, N.OPT_SGD if sgd else N.OPT_ADAM, flags, gemm_mode,
                                 label_smoothing, final_dropout_rate or 0.0, rsd, n_classes)
        self.gemm_mode = gemm_mode
        self.n_classes = n_classes
        self.n_out = n_classes + 1 if n_classes > 1 else 2
        self.ctx = N.Context(self.cfg, device)
        self.lib = N.lib()
        self.image_size = image_size
        self.max_batch = max_batch
        self.n_slots = n_slots
        self.n_theta = self.ctx.n_theta
        self.n_bn = self.ctx.n_bn
        self.n_dc = self.ctx.n_dc
        self.state_floats = self.ctx.state_floats
        # ONE arena, one layout per slot at a uniform stride: [state | workspace | staging].  Task-batched launches
        # (mliis_task_args.n_group) address slot k of a group as pointer + k * slot_stride for every per-slot pointer.
        al = lambda n: (int(n) + 255) // 256 * 256
        if staging_bytes is None:      # room for a 16-example task pool (images + 2-channel labels) + index / lr blocks
            staging_bytes = 16 * image_size * image_size * 5 * 4 + (1 << 16)
        self.state_bytes, self.ws_bytes, self.staging_bytes = al(self.state_floats * 4), al(self.ctx.workspace_bytes), al(staging_bytes)
        self.slot_stride = self.state_bytes + self.ws_bytes + self.staging_bytes
        self.arena = torch.zeros(n_slots * self.slot_stride + 256, dtype=torch.uint8, device=self.device)
        self._arena_off = (-self.arena.data_ptr()) % 256
        body = self.arena[self._arena_off:self._arena_off + n_slots * self.slot_stride]
        self._slot_bytes = body.view(n_slots, self.slot_stride)
        self.states = body.view(torch.float32).view(n_slots, self.slot_stride // 4)[:, :self.state_floats]
        self.workspaces = [self._slot_bytes[s, self.state_bytes:self.state_bytes + self.ws_bytes] for s in range(n_slots)]
        for s in range(n_slots):
            N.check(self.lib.mliis_slot_bind(self.ctx.handle, s, _ptr(self.states[s]), _ptr(self.workspaces[s])))
        self.o_bn = self.n_theta
        self.o_v = self.n_theta + 2 * self.n_bn
        self.o_pow = self.o_v + self.n_theta
        # gather/scatter index between the TF-ordered concatenation of variables and the flat layout
        idx = np.concatenate([np.arange(p.offset, p.offset + p.size, dtype=np.int64) for p in self.ctx.params])
        self._flat_index = torch.from_numpy(idx).to(self.device)
        self._sizes = [p.size for p in self.ctx.params]
        self._shapes = [p.shape for p in self.ctx.params]

    def staging(self, slot: int) -> torch.Tensor:
        """The slot's staging region (uint8 view of the arena): task inputs / outputs that a task-batched launch
        must find at the uniform slot stride live here."""
        o = self.state_bytes + self.ws_bytes
        return self._slot_bytes[slot, o:o + self.staging_bytes]

    # ---- streams ----
    @staticmethod
    def _stream(stream: Optional[torch.cuda.Stream] = None):
        s = stream if stream is not None else torch.cuda.current_stream()
        return C.c_void_p(s.cuda_stream)

    # ---- state views ----
    def theta(self, slot: int = 0) -> torch.Tensor:
        return self.states[slot, :self.n_theta]

    def bn_state(self, slot: int = 0) -> torch.Tensor:
        return self.states[slot, self.o_bn:self.o_bn + 2 * self.n_bn].view(2, self.n_bn)

    def adam_v(self, slot: int = 0) -> torch.Tensor:
        return self.states[slot, self.o_v:self.o_v + self.n_theta]

    def powers(self, slot: int = 0) -> torch.Tensor:
        return self.states[slot, self.o_pow:self.o_pow + 2]

    def pack_theta(self, variables: Sequence[np.ndarray]) -> torch.Tensor:
        """list of arrays in tf.trainable_variables() order -> flat engine layout (device)."""
        if len(variables) != len(self._sizes):
            raise ValueError("expected %d variables, got %d" % (len(self._sizes), len(variables)))
        parts = []
        for v, shape in zip(variables, self._shapes):
            a = np.asarray(v, dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError("variable shape %s != expected %s" % (a.shape, shape))
            parts.append(a.reshape(-1))
        cat = torch.from_numpy(np.concatenate(parts)).to(self.device)
        flat = torch.zeros(self.n_theta, dtype=torch.float32, device=self.device)
        flat[self._flat_index] = cat
        return flat

    def unpack_theta(self, flat: torch.Tensor) -> List[np.ndarray]:
        cat = flat[self._flat_index].cpu().numpy()
        out, o = [], 0
        for size, shape in zip(self._sizes, self._shapes):
            out.append(cat[o:o + size].reshape(shape).copy())
            o += size
        return out

    def tf_order_vector(self, flat: torch.Tensor) -> torch.Tensor:
        """flat engine layout -> 1-D tensor in TF creation order (no padding)."""
        return flat[self._flat_index]

    def init_state(self, slot: int, variables: Sequence[np.ndarray], moving_mean: np.ndarray,
                   moving_var: np.ndarray, adam_v: Optional[Sequence[np.ndarray]] = None,
                   beta1_power: float = ADAM_BETA1, beta2_power: float = ADAM_BETA2) -> None:
        st = self.states[slot]
        st.zero_()
        st[:self.n_theta] = self.pack_theta(variables)
        bn = self.bn_state(slot)
        bn[0] = torch.from_numpy(np.asarray(moving_mean, np.float32)).to(self.device)
        bn[1] = torch.from_numpy(np.asarray(moving_var, np.float32)).to(self.device)
        if adam_v is not None:
            st[self.o_v:self.o_v + self.n_theta] = self.pack_theta(adam_v)
        p = self.powers(slot)
        p[0] = beta1_power
        p[1] = beta2_power

    def copy_state(self, dst: torch.Tensor, src: torch.Tensor, what: int = N.STATE_ALL, stream=None) -> None:
        N.check(self.lib.mliis_state_copy(self.ctx.handle, _ptr(dst), _ptr(src), what, self._stream(stream)))

    # ---- compute ----
    def train_step(self, slot: int, images: torch.Tensor, labels: torch.Tensor, lr: float,
                   index: Optional[torch.Tensor] = None, batch: Optional[int] = None,
                   dc_mask: Optional[torch.Tensor] = None, drop_mask: Optional[torch.Tensor] = None, seed: int = 0,
                   pre_decay_rate: float = 1.0, loss_out: Optional[torch.Tensor] = None, stream=None,
                   seed_dev: Optional[torch.Tensor] = None) -> None:
        """seed_dev: optional int64 device scalar added to `seed` when the final-layer dropout mask is drawn; it is
        read at run time, so a CUDA graph that captured this call draws fresh masks on every replay."""
        B = int(batch if batch is not None else (index.numel() if index is not None else images.shape[0]))
        a = N.StepArgs(_ptr(images), _ptr(labels), _ptr(index), B, float(lr), float(pre_decay_rate), _ptr(dc_mask),
                       _ptr(drop_mask), int(seed), _ptr(loss_out), _ptr(seed_dev))
        N.check(self.lib.mliis_train_step(self.ctx.handle, slot, C.byref(a), self._stream(stream)))

    def forward(self, slot: int, images: torch.Tensor, training: bool, index: Optional[torch.Tensor] = None,
                batch: Optional[int] = None, dc_mask=None, drop_mask=None, seed: int = 0,
                want_logits: bool = True, stream=None) -> Optional[torch.Tensor]:
        B = int(batch if batch is not None else (index.numel() if index is not None else images.shape[0]))
        H = self.image_size
        if self.n_out == 2:
            logits = torch.empty(B, H, H, 2, dtype=torch.float32, device=self.device) if want_logits else None
        else:   # multi-class head: the LOW-resolution head output [B, H/4, W/4, n_out]
            logits = torch.empty(B, H // 4, H // 4, self.n_out, dtype=torch.float32,
                                 device=self.device) if want_logits else None
        N.check(self.lib.mliis_forward(self.ctx.handle, slot, _ptr(images), _ptr(index), B, int(training),
                                       _ptr(dc_mask), _ptr(drop_mask), int(seed), _ptr(logits),
                                       self._stream(stream)))
        return logits

    def loss_backward(self, slot: int, labels: torch.Tensor, batch: int, index: Optional[torch.Tensor] = None,
                      want_grads: bool = True, stream=None):
        grads = torch.empty(self.n_theta, dtype=torch.float32, device=self.device) if want_grads else None
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        N.check(self.lib.mliis_loss_backward(self.ctx.handle, slot, _ptr(labels), _ptr(index), batch, _ptr(grads),
                                             _ptr(loss), self._stream(stream)))
        return loss, grads

    def set_grads(self, slot: int, grads: torch.Tensor, stream=None) -> None:
        """Overwrites the slot's gradient buffer (data-parallel all-reduce result) before optimizer_step."""
        N.check(self.lib.mliis_set_grads(self.ctx.handle, slot, _ptr(grads), self._stream(stream)))

    def set_class_ids(self, slot: int, class_ids: torch.Tensor) -> None:
        """Multi-class head: int32 [n_pool] foreground class of every pool example (kept referenced)."""
        assert class_ids.dtype == torch.int32 and class_ids.is_cuda
        self._class_ids = getattr(self, "_class_ids", {})
        self._class_ids[slot] = class_ids
        N.check(self.lib.mliis_set_class_ids(self.ctx.handle, slot, _ptr(class_ids)))

    def predict_classes(self, slot: int, images: torch.Tensor, masks: Optional[torch.Tensor] = None,
                        index: Optional[torch.Tensor] = None, batch: Optional[int] = None,
                        want_class_map: bool = True, stream=None):
        """(class map int32 [B,H,W] with -1 = no class above 0.5, inter, union) of the multi-class head."""
        B = int(batch if batch is not None else (index.numel() if index is not None else images.shape[0]))
        H = self.image_size
        cmap = torch.empty(B, H, H, dtype=torch.int32, device=self.device) if want_class_map else None
        inter = uni = None
        if masks is not None:
            inter = torch.zeros(B, dtype=torch.int32, device=self.device)
            uni = torch.zeros(B, dtype=torch.int32, device=self.device)
        N.check(self.lib.mliis_predict_classes(self.ctx.handle, slot, _ptr(images), _ptr(masks), _ptr(index), B,
                                               _ptr(cmap), _ptr(inter), _ptr(uni), self._stream(stream)))
        return cmap, inter, uni

    def optimizer_step(self, slot: int, lr: float, stream=None) -> None:
        N.check(self.lib.mliis_optimizer_step(self.ctx.handle, slot, float(lr), 1.0, self._stream(stream)))

    def predict(self, slot: int, images: torch.Tensor, labels: Optional[torch.Tensor] = None,
                index: Optional[torch.Tensor] = None, batch: Optional[int] = None, want_pred: bool = True,
                want_logits: bool = False, stream=None):
        B = int(batch if batch is not None else (index.numel() if index is not None else images.shape[0]))
        H = self.image_size
        pred = torch.empty(B, H, H, 2, dtype=torch.float32, device=self.device) if want_pred else None
        logits = torch.empty(B, H, H, 2, dtype=torch.float32, device=self.device) if want_logits else None
        inter = uni = None
        if labels is not None:
            inter = torch.zeros(B, dtype=torch.int32, device=self.device)
            uni = torch.zeros(B, dtype=torch.int32, device=self.device)
        N.check(self.lib.mliis_predict(self.ctx.handle, slot, _ptr(images), _ptr(labels), _ptr(index), B, _ptr(pred),
                                       _ptr(logits), _ptr(inter), _ptr(uni), self._stream(stream)))
        return pred, logits, inter, uni

    def adapt_eval_task(self, slot: int, init_state: torch.Tensor, images: torch.Tensor, labels: torch.Tensor,
                        batch_index: torch.Tensor, lrs: torch.Tensor, n_steps: int, batch: int,
                        query_index: torch.Tensor, inter_out: torch.Tensor, union_out: torch.Tensor,
                        dc_mask: Optional[torch.Tensor] = None, seed: int = 0, pre_decay_rate: float = 1.0,
                        loss_out: Optional[torch.Tensor] = None, stream=None,
                        seed_dev: Optional[torch.Tensor] = None) -> None:
        a = N.TaskArgs(_ptr(init_state), _ptr(images), _ptr(labels), _ptr(batch_index), _ptr(lrs), int(n_steps),
                       int(batch), _ptr(query_index), int(query_index.numel()), _ptr(dc_mask), int(seed),
                       float(pre_decay_rate), _ptr(inter_out), _ptr(union_out), _ptr(loss_out), _ptr(seed_dev))
        N.check(self.lib.mliis_adapt_eval_task(self.ctx.handle, slot, C.byref(a), self._stream(stream)))

    def delta_accumulate(self, dsum: torch.Tensor, a: torch.Tensor, b: torch.Tensor, first: bool, stream=None):
        N.check(self.lib.mliis_delta_accumulate(self.ctx.handle, _ptr(dsum), _ptr(a), _ptr(b), int(first),
                                                self._stream(stream)))

    def meta_apply(self, theta: torch.Tensor, dsum: torch.Tensor, scale: float, stream=None):
        N.check(self.lib.mliis_meta_apply(self.ctx.handle, _ptr(theta), _ptr(dsum), float(scale),
                                          self._stream(stream)))

    # ---- the exchange step of a meta-update: ONE NCCL all-reduce through the C ABI ----
    def init_comm(self) -> None:
        """Creates the ctx's NCCL communicator when torch.distributed runs with world > 1 (torch.distributed is only
        the control plane that carries the 128-byte unique id)."""
        import torch.distributed as dist
        if getattr(self, "_comm_ready", False):
            return
        self._comm_ready = True
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        rank, world = dist.get_rank(), dist.get_world_size()
        buf = (C.c_uint8 * 128)()
        if rank == 0:
            N.check(self.lib.mliis_comm_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8, device=self.device)
        dist.broadcast(t, 0)
        raw = (C.c_uint8 * 128)(*t.cpu().tolist())
        N.check(self.lib.mliis_comm_init(self.ctx.handle, raw, rank, world))

    def meta_buffer(self) -> torch.Tensor:
        """[sum of task deltas | sum of BN moving statistics | #contributing slots | pad] (mliis_meta_buffer_floats)."""
        n = int(self.lib.mliis_meta_buffer_floats(self.ctx.handle))
        return torch.zeros(n, dtype=torch.float32, device=self.device)

    def meta_reduce(self, buf: torch.Tensor, rows: Optional[torch.Tensor], row_stride: int, first_slot: int,
                    n_rows: int, stream=None) -> None:
        N.check(self.lib.mliis_meta_reduce(self.ctx.handle, _ptr(buf), _ptr(rows), int(row_stride), int(first_slot),
                                           int(n_rows), self._stream(stream)))

    def allreduce_delta(self, buf: torch.Tensor, stream=None) -> None:
        N.check(self.lib.mliis_allreduce_delta(self.ctx.handle, _ptr(buf), buf.numel(), self._stream(stream)))

    def meta_finish(self, theta: torch.Tensor, buf: torch.Tensor, scale: float, first_slot: int, n_slots: int,
                    stream=None) -> None:
        N.check(self.lib.mliis_meta_finish(self.ctx.handle, _ptr(theta), _ptr(buf), float(scale), int(first_slot),
                                           int(n_slots), self._stream(stream)))

    # ---- debugging ----
    def debug_buffer(self, slot: int, name: str, batch: int) -> torch.Tensor:
        """Copy of a named activation buffer as [batch, rows_per_image, C] (tests only)."""
        ptr, rows, ch, ld = C.c_void_p(), C.c_int64(), C.c_int32(), C.c_int32()
        N.check(self.lib.mliis_debug_buffer(self.ctx.handle, slot, name.encode(), C.byref(ptr), C.byref(rows),
                                            C.byref(ch), C.byref(ld)))
        torch.cuda.synchronize()
        ws = self.workspaces[slot]
        base_off = ptr.value - ws.data_ptr()
        n = batch * rows.value * ld.value
        raw = ws[base_off:base_off + n * 4].view(torch.float32)
        return raw.view(batch, rows.value, ld.value)[:, :, :ch.value].clone()

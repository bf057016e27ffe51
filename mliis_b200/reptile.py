"""Gecko (Reptile) and FOMLIS (FOMAML) meta-learners for image segmentation.

Mirror of /root/reference/meta_learners/supervised_reptile/supervised_reptile/reptile.py: the same classes,
constructor and method signatures, the same order of calls into Python's ``random`` (so task / split /
mini-batch indices are bit-identical to the reference under the same seed), the same return values.

Two execution paths produce the same results:
  * the Session path follows the reference line by line (sess.run per inner step, export/import of variables
    through host numpy) - used when a feature outside the device fast path is requested (augmentation,
    per-task checkpoints, non-transductive evaluation);
  * the device fast path keeps theta / BN statistics / optimizer slots on the GPU: `evaluate` turns every task
    into a TaskPlan executed by runner.TaskRunner (one CUDA graph per task, tasks sharded over slots and
    ranks), `train_step` accumulates the per-task deltas with mliis_delta_accumulate, all-reduces them over
    NCCL when world_size > 1 and applies theta += eps/M * sum(delta) with mliis_meta_apply.
"""
from __future__ import annotations

import os
import random
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import native as N
from .metaseg import (DEFAULT_NUM_TEST_EXAMPLES, _mini_batches, _sample_mini_image_segmentation_dataset,
                      _sample_task_indices, _sample_train_test_segmentation_with_replacement,
                      _split_train_test_segmentation)
from .variables import (VariableState, add_vars, average_vars, interpolate_vars, scale_vars, subtract_vars)

DEFAULT_ITER_RANGE = [1, 5, 10, 25, 50, 100, 200]


def _dist():
    """(rank, world) of the task-parallel group; (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


from .native import GEMM_FP32 as N_GEMM_FP32


class _ExamplePool:
    """The distinct (image, mask) arrays one task feeds to the device, in first-use order.  An example the augmenter
    returned untouched is the support example itself (same array objects) and shares its row; every augmented copy is
    a new row.  Inner batches / the query set are then rows of this pool (TaskPlan.batch_index / query_index)."""

    def __init__(self):
        self.images, self.labels, self._row = [], [], {}

    def row(self, image, label) -> int:
        key = (id(image), id(label))
        r = self._row.get(key)
        if r is None:
            r = len(self.images)
            self._row[key] = r
            self.images.append(image)      # keeps the arrays alive: ids stay unique
            self.labels.append(label)
        return r

    def arrays(self):
        return (np.stack([np.asarray(a, np.float32) for a in self.images]),
                np.stack([np.asarray(a, np.float32) for a in self.labels]))


class Gecko:
    """A meta-learning session for image segmentation that extends Reptile (reptile.py:23-62)."""

    def __init__(self, session, variables=None, transductive=False, pre_step_op=None, lr_scheduler=None,
                 augment: bool = False, aug_rate: Optional[float] = None, fast_path: bool = True,
                 meta_task_slots: Optional[int] = None):
        self.session = session
        model = session.model
        self._model = model
        self._model_state = VariableState(session, variables or model.trainable_variables())
        self._full_state = VariableState(session, model.global_variables())
        self._transductive = transductive
        self._pre_step_op = pre_step_op
        self.eval_sample_number = 0
        self.lr_scheduler = lr_scheduler
        if augment:
            # host numpy augmentations (reptile.py:48-52), drawn with the reference's RNG call order.  The device fast
            # path stays on: every augmented example becomes a row of the task's example pool (_ExamplePool below)
            from .np_augmenters import Augmenter
            self.augmenter = Augmenter()
        else:
            self.augmenter = None
        self.aug_rate = aug_rate
        self.fast_path = fast_path
        self._runner = None
        # meta-training: 1 = the reference's sequential order (optimizer slots / BN statistics flow task -> task);
        # S > 1 = S task slots adapt the tasks of a meta-batch concurrently, each slot carrying its own optimizer
        # slots, BN statistics averaged over slots after the meta-step - the same documented semantics as S ranks
        # (SURVEY.md 8e); exact for the trainables under --sgd up to summation order
        if meta_task_slots is None:
            meta_task_slots = int(os.environ.get("MLIIS_META_TASK_SLOTS", "1"))
        self.meta_task_slots = max(1, int(meta_task_slots))
        self._train_slots = None
        print("Augmentation rate {}".format(self.aug_rate))
        print("Using transduction in meta-learning." if transductive else "Not using transduction in meta-learning.")
        self.meta_fn = "Reptile"
        print("Reptile meta-learning session instantiated.")

    # ------------------------------------------------------------------------------------------
    # helpers shared by both paths
    # ------------------------------------------------------------------------------------------
    def _next_seed(self) -> int:
        # dropout seeds come from a private counter: Python's `random` stream must stay reference-identical
        self._seed_counter = getattr(self, "_seed_counter", 0) + 1
        return self._seed_counter

    def _pre_decay(self) -> float:
        return float(self._pre_step_op.rate) if self._pre_step_op is not None else 1.0

    def _inner_lr(self, lr_ph, lr, step: int, eval_mode: bool) -> float:
        """Learning rate one inner step runs with (reptile.py:269-279 for evaluation, :114-121 for training)."""
        default = self._model.lr_ph.default
        if lr_ph is not None and lr is not None:
            return float(lr)
        if lr_ph is not None and self.lr_scheduler is not None:
            return float(self.lr_scheduler.cur_lr(cur_step=step))
        return float(default)

    def _get_runner(self, n_pool, n_steps, batch, n_query):
        from .runner import TaskRunner
        key = (n_pool, n_steps, batch, n_query, self._pre_decay())
        if self._runner is None or self._runner[0] != key:
            eng = self._model.engine()
            # task-batched launches (several slots per kernel launch) when at least two groups stay in flight; measured
            # on B200: 12 slots x 1 -> 111, 16 x 4 -> 115, 32 x 8 -> 120 tasks/s.  Same results to fp32 rounding.  Augmented
            # pools live outside the uniform-stride arena, which single-slot launches only can address.
            group = 1
            if self.augmenter is None and eng.gemm_mode != N_GEMM_FP32:
                # (with the partials sized per launch: 32 x 8 -> 134.7, 32 x 16 -> 140.1, 48 x 24 -> 143.2 tasks/s)
                for g in range(min(32, eng.n_slots // 2), 1, -1):
                    if eng.n_slots % g == 0:
                        group = g
                        break
            self._runner = (key, TaskRunner(eng, n_pool, n_steps, batch, n_query, use_graph=True,
                                            pre_decay_rate=self._pre_decay(), group=group))
        return self._runner[1]

    # ------------------------------------------------------------------------------------------
    # meta-training step (reptile.py:64-125)
    # ------------------------------------------------------------------------------------------
    def train_step(self, dataset, input_ph, label_ph, minimize_op, num_classes, num_shots, inner_batch_size,
                   inner_iters, replacement, meta_step_size, meta_batch_size, lr_ph=None, lr=None, verbose=False):
        num_classes = 1      # hardcoded binary Gecko (reptile.py:99-100)
        if self.fast_path:
            return self._train_step_device(dataset, num_shots, inner_batch_size, inner_iters, replacement,
                                           meta_step_size, meta_batch_size, lr_ph, lr, fomaml=False)
        old_vars = self._model_state.export_variables()
        new_vars = []
        for _ in range(meta_batch_size):
            mini_dataset = _sample_mini_image_segmentation_dataset(self.session, dataset, num_classes, num_shots)
            for i, batch in enumerate(_mini_batches(mini_dataset, inner_batch_size, inner_iters, replacement,
                                                    augmenter=self.augmenter)):
                inputs, labels = zip(*batch)
                if self._pre_step_op:
                    self.session.run(self._pre_step_op)
                # NB the reference's `if ... if ... else` (reptile.py:114-121): with lr given and no scheduler the
                # minimize op runs TWICE per batch (once with lr, once with the default lr)
                if (lr_ph is not None) and (lr is not None):
                    self.session.run(minimize_op, feed_dict={input_ph: inputs, label_ph: labels, lr_ph: lr})
                if (lr_ph is not None) and (self.lr_scheduler is not None):
                    self.session.run(minimize_op, feed_dict={input_ph: inputs, label_ph: labels,
                                                             lr_ph: self.lr_scheduler.cur_lr(cur_step=i)})
                else:
                    self.session.run(minimize_op, feed_dict={input_ph: inputs, label_ph: labels})
            new_vars.append(self._model_state.export_variables())
            self._model_state.import_variables(old_vars)
        new_vars = average_vars(new_vars)
        self._model_state.import_variables(interpolate_vars(old_vars, new_vars, meta_step_size))

    def _task_batches_for_training(self, rows, inner_batch_size, inner_iters, replacement):
        """Index batches of one meta-training task; overridden by FOMLIS (tail batch)."""
        return list(_mini_batches(rows, inner_batch_size, inner_iters, replacement, augmenter=None))

    def _array_batches_for_training(self, samples, inner_batch_size, inner_iters, replacement):
        """Array-space twin of `_task_batches_for_training` (reptile.py:108): Gecko passes no aug_rate here, so the
        Augmenter falls back to its own prob_to_return_original; overridden by FOMLIS (tail batch, aug_rate)."""
        return _mini_batches(samples, inner_batch_size, inner_iters, replacement, augmenter=self.augmenter)

    def _train_plan(self, task, rows, inner_batch_size, inner_iters, replacement, build: bool = True):
        """(images, labels, batches) of one meta-training task for the device path; batches are rows of `images`.
        Without an augmenter the pool is the task's first len(rows) records.  With one (run.sh: --augment) every inner
        batch holds freshly augmented copies drawn on the host with the reference's RNG order; they become extra rows
        of the pool (_ExamplePool).  build=False only consumes the RNG streams (tasks owned by another rank)."""
        if self.augmenter is None:
            batches = self._task_batches_for_training(rows, inner_batch_size, inner_iters, replacement)
            if not build:
                return None, None, batches
            images, labels = task.arrays()
            return images[:len(rows)], labels[:len(rows)], batches
        images, labels = task.arrays()
        samples = [[images[r], labels[r]] for r in rows]
        pool = _ExamplePool()
        for ex in samples:                       # the support examples first: rows 0..n-1 like the plain path
            pool.row(ex[0], ex[1])
        batches = [[pool.row(ex[0], ex[1]) for ex in batch]
                   for batch in self._array_batches_for_training(samples, inner_batch_size, inner_iters, replacement)]
        if not build:
            return None, None, batches
        pi, pl = pool.arrays()
        return pi, pl, batches

    def _train_lrs(self, lr_ph, lr, n_batches) -> List[List[float]]:
        """Per batch, the learning rates of the minimize runs the reference performs (reptile.py:114-121)."""
        out = []
        for i in range(n_batches):
            runs = []
            if (lr_ph is not None) and (lr is not None):
                runs.append(float(lr))
            if (lr_ph is not None) and (self.lr_scheduler is not None):
                runs.append(float(self.lr_scheduler.cur_lr(cur_step=i)))
            else:
                runs.append(float(self._model.lr_ph.default))
            out.append(runs)
        return out

    def _train_step_device(self, dataset, num_shots, inner_batch_size, inner_iters, replacement, meta_step_size,
                           meta_batch_size, lr_ph, lr, fomaml: bool):
        """Device-resident meta-step.  world_size == 1 reproduces the reference's sequential carry-over of the
        optimizer slots / BN moving statistics between the tasks of a meta-batch exactly; with more ranks the
        tasks are dealt round-robin, each rank carries its own slots, and the summed deltas (and the BN moving
        statistics) are all-reduced once per meta-step."""
        import torch
        eng = self._model.engine()
        rank, world = _dist()
        if self.meta_task_slots > 1 and eng.n_slots > 1:
            return self._train_step_slots(dataset, num_shots, inner_batch_size, inner_iters, replacement,
                                          meta_step_size, meta_batch_size, lr_ph, lr, fomaml)
        theta = eng.theta(0)
        old = theta.clone()
        if world > 1:
            eng.init_comm()
        if getattr(self, "_meta_buf", None) is None:
            self._meta_buf = eng.meta_buffer()
        buf = self._meta_buf
        dsum = buf[:eng.n_theta]                             # the delta sum is the head of the exchange buffer
        backup = torch.empty_like(theta) if fomaml else None
        n_local = 0
        for t in range(meta_batch_size):
            # every rank draws every task so that the `random` stream stays identical across ranks
            task, rows = _sample_task_indices(dataset, num_shots)
            mine = t % world == rank
            images, labels, batches = self._train_plan(task, rows, inner_batch_size, inner_iters, replacement, build=mine)
            if not mine:
                continue
            x = torch.from_numpy(np.ascontiguousarray(images)).to(eng.device, non_blocking=True)
            y = torch.from_numpy(np.ascontiguousarray(labels)).to(eng.device, non_blocking=True)
            lrs = self._fomaml_lrs(lr_ph, lr, len(batches)) if fomaml else self._train_lrs(lr_ph, lr, len(batches))
            for j, batch in enumerate(batches):
                idx = torch.tensor(batch, dtype=torch.int32, device=eng.device)
                if fomaml and j == inner_iters - 1:
                    backup.copy_(theta)                      # last_backup (reptile.py:635-636)
                for k, step_lr in enumerate(lrs[j]):
                    eng.train_step(0, x, y, step_lr, index=idx, pre_decay_rate=self._pre_decay() if k == 0 else 1.0,
                                   seed=self._next_seed())
            eng.delta_accumulate(dsum, theta, backup if fomaml else old, first=(n_local == 0))
            n_local += 1
            theta.copy_(old)                                 # import_variables(old_vars): trainables only
        # the ONE exchange step: [sum of deltas | BN statistics | count] -> ncclAllReduce -> theta += eps/M * sum
        eng.meta_reduce(buf, dsum, eng.n_theta, 0, 1 if n_local else 0)
        eng.allreduce_delta(buf)
        eng.meta_finish(theta, buf, float(meta_step_size) / float(meta_batch_size), 0, 1)

    def _train_step_slots(self, dataset, num_shots, inner_batch_size, inner_iters, replacement, meta_step_size,
                          meta_batch_size, lr_ph, lr, fomaml: bool):
        """Slot-parallel meta-step: this rank's tasks are dealt round-robin to S task slots; every slot replays one
        CUDA graph per task (theta <- theta_old, the inner steps, delta accumulation) on its own stream."""
        import torch
        from .runner import TrainSlots
        eng = self._model.engine()
        rank, world = _dist()
        plans = []
        for t in range(meta_batch_size):
            task, rows = _sample_task_indices(dataset, num_shots)        # every rank draws every task
            mine = t % world == rank
            images, labels, batches = self._train_plan(task, rows, inner_batch_size, inner_iters, replacement, build=mine)
            if mine:
                plans.append((images, labels, batches))
        if plans:
            n_batches = len(plans[0][2])
            lrs = self._fomaml_lrs(lr_ph, lr, n_batches) if fomaml else self._train_lrs(lr_ph, lr, n_batches)
            n_rows = len(plans[0][0])
            if self.augmenter is not None:           # room for one fresh copy per batch element
                n_rows = num_shots + sum(len(b) for b in plans[0][2])
            shape = (n_rows, tuple(len(b) for b in plans[0][2]), tuple(tuple(l) for l in lrs), fomaml,
                     self._pre_decay())
            n_max = min(self.meta_task_slots, eng.n_slots)
            # task-batched groups of slots when the rank's tasks fill whole groups (Reptile meta-batch 40 on 16 slots:
            # 5 chunks of 8); augmented pools live outside the uniform-stride arena, fp32 mode has no batched kernels
            group = 1
            if self.augmenter is None and eng.gemm_mode != N_GEMM_FP32:
                # at least two groups in flight: ONE lockstep group is a single serial chain of kernels (measured: FOMAML
                # meta-batch 5 as one group of five 19.0 meta-steps/s, as five single-slot graphs 20.5)
                # (Reptile meta-batch 40: 16 slots as 2 groups of 8 3.29 meta-steps/s, 20 slots as 2 groups of 10 3.47)
                for g in range(min(12, n_max // 2, len(plans) // 2), 1, -1):
                    if len(plans) % g == 0:
                        group = g
                        break
                forced = int(os.environ.get("MLIIS_TRAIN_GROUP", "0"))       # experiments: force the group size
                if forced >= 1 and forced <= n_max and len(plans) % forced == 0:
                    group = forced
            # no uniform group divides the rank's task count (FOMAML's meta-batch of 5; Reptile's 40 on 8 GPUs): two
            # lockstep groups of ceil / floor(n / 2) slots - kernels that serve 2-3 tasks, two chains that fill each
            # other's tails - instead of n single-slot graphs (MLIIS_TRAIN_SPLIT=0 keeps the single-slot graphs)
            sizes = None
            if (group == 1 and self.augmenter is None and eng.gemm_mode != N_GEMM_FP32 and 4 <= len(plans) <= n_max
                    and os.environ.get("MLIIS_TRAIN_SPLIT", "1") != "0"):
                sizes = ((len(plans) + 1) // 2, len(plans) // 2)
            shape = shape + (group, sizes)
            if self._train_slots is None or self._train_slots.shape != shape:
                try:
                    if sizes is not None:
                        self._train_slots = TrainSlots(eng, len(plans), shape[:-2], group_sizes=sizes)
                    else:
                        self._train_slots = TrainSlots(eng, (n_max // group) * group, shape[:-2], group=group)
                except ValueError:               # the pool does not fit the arena's staging region: single-slot launches
                    self._train_slots = TrainSlots(eng, n_max, shape[:-2], group=1)
                self._train_slots.shape = shape
            ts = self._train_slots
            ts.begin(eng.theta(0))
            ts.run_tasks(plans)
            buf = ts.finish()
            n_write = ts.n
        else:
            if getattr(self, "_meta_buf", None) is None:
                self._meta_buf = eng.meta_buffer()
            buf = self._meta_buf
            eng.meta_reduce(buf, None, eng.n_theta, 0, 0)     # this rank adapted no task of the meta-batch: zeros
            ts, n_write = None, 1
        if world > 1:
            eng.init_comm()
        eng.allreduce_delta(buf)                               # ONE ncclAllReduce over [deltas | BN statistics | count]
        eng.meta_finish(eng.theta(0), buf, float(meta_step_size) / float(meta_batch_size), 0, n_write)
        if ts is not None:
            ts.save_states()

    def _fomaml_lrs(self, lr_ph, lr, n_batches):
        d = float(self._model.lr_ph.default)
        return [[float(lr)] if (lr_ph is not None and lr is not None) else [d] for _ in range(n_batches)]

    # ------------------------------------------------------------------------------------------
    # evaluation (reptile.py:127-294)
    # ------------------------------------------------------------------------------------------
    def evaluate(self, dataset, input_ph, label_ph, minimize_op, predictions, num_classes, num_shots,
                 inner_batch_size, inner_iters, replacement, eval_all_tasks=False, num_tasks_to_sample=1,
                 test_shots=DEFAULT_NUM_TEST_EXAMPLES, verbose=False, save_fine_tuned_checkpoints=False,
                 save_fine_tuned_checkpoints_dir: Optional[str] = None, eval_sample_num: Optional[int] = None,
                 is_training_ph=None, lr_ph=None, lr: Optional[float] = None, drop_rate_ph=None,
                 drop_rate: Optional[float] = None, aug_rate: Optional[float] = None) -> Tuple[float, Dict[str, float]]:
        print("Evaluating {} meta-learning.".format(self.meta_fn))
        if aug_rate is None:
            aug_rate = self.aug_rate
        if eval_all_tasks:
            sampled_tasks = dataset
        else:
            random.shuffle(dataset)        # in place, like the reference (reptile.py:186-188)
            sampled_tasks = dataset[:num_tasks_to_sample]
        print("Evaluating {} {}-shot tasks.".format(len(sampled_tasks), num_shots))
        # per-task fine-tuned checkpoints need the adapted state of every task on the host: Session path
        device_ok = (self.fast_path and not save_fine_tuned_checkpoints and self._transductive
                     and is_training_ph is not None and drop_rate is None
                     and inner_iters > 0 and all(hasattr(t, "arrays") for t in sampled_tasks))
        if device_ok:
            ious, task_iou_map = self._evaluate_device(sampled_tasks, num_shots, test_shots, inner_batch_size,
                                                       inner_iters, replacement, lr_ph, lr, aug_rate)
        else:
            ious, task_iou_map = [], {}
            for sampled_task in sampled_tasks:
                sampled, task_name = _sample_mini_image_segmentation_dataset(
                    self.session, [sampled_task], num_classes, num_shots + test_shots, return_task_name=True)
                print("Evaluating {}".format(task_name))
                train_set, test_set = _split_train_test_segmentation(sampled, test_shots)
                task_iou = self._evaluate(train_set, test_set, input_ph, label_ph, minimize_op, predictions,
                                          inner_batch_size, inner_iters, replacement, verbose=verbose,
                                          save_fine_tuned_checkpoints=save_fine_tuned_checkpoints,
                                          save_fine_tuned_checkpoints_dir=save_fine_tuned_checkpoints_dir,
                                          eval_sample_num=eval_sample_num, is_training_ph=is_training_ph, lr_ph=lr_ph, lr=lr, task_name=task_name,
                                          drop_rate_ph=drop_rate_ph, drop_rate=drop_rate, aug_rate=aug_rate)
                ious.append(task_iou)
                task_iou_map[task_name] = task_iou
        mean_iou_score = np.nanmean(ious)
        print("Evaluated {} task/s".format(len(sampled_tasks)))
        print("Mean IoU from train on {} images and evaluate on {} test images: {}".format(num_shots, test_shots,
                                                                                             mean_iou_score))
        return mean_iou_score, task_iou_map

    def _evaluate_device(self, sampled_tasks, num_shots, test_shots, inner_batch_size, inner_iters, replacement,
                         lr_ph, lr, aug_rate=None):
        """All tasks of one evaluation pass on the device, sharded over slots and ranks.  The host draws the
        plans sequentially (same `random` / `np.random` consumption as the reference), the device adapts them
        concurrently.  With an augmenter (run.sh: --augment --aug_rate 0.5) every inner batch consists of freshly
        augmented copies (metaseg.py:258-302, np_augmenters.py:135-160): they are produced on the host by the
        bit-exact numpy mirror, in the reference's draw order, and staged as extra rows of the task's example pool;
        tasks are then drawn and run in chunks of n_slots so that only a few pools are alive at a time."""
        from .runner import TaskPlan, gather_owned, iou_from_counts
        rank, world = _dist()
        eng = self._model.engine()
        names, owned, task_plans = [], [], []
        n_pool = num_shots + test_shots
        if self.augmenter is not None:
            n_pool += inner_iters * inner_batch_size
        runner = self._get_runner(n_pool, inner_iters, inner_batch_size, test_shots)
        runner.set_init_state(eng.states[0])          # old_vars = _full_state.export_variables() (reptile.py:258)
        ious_local = np.full(len(sampled_tasks), np.nan, np.float64)

        def flush():
            for i, (inter, uni) in zip(owned, runner.run(task_plans)):
                ious_local[i] = iou_from_counts(inter, uni)
            owned.clear()
            task_plans.clear()

        for i, task in enumerate(sampled_tasks):
            obj, rows = _sample_task_indices([task], num_shots + test_shots)
            names.append(obj.name)
            print("Evaluating {}".format(obj.name))
            lrs = np.asarray([self._inner_lr(lr_ph, lr, s, True) for s in range(inner_iters)], np.float32)
            mine = i % world == rank
            if self.augmenter is None:
                train, test = _split_train_test_segmentation(rows, test_shots)
                batches = list(_mini_batches(train, inner_batch_size, inner_iters, replacement, augmenter=None))
                if mine:
                    images, labels = obj.arrays()
                    task_plans.append(TaskPlan(images[:n_pool], labels[:n_pool], np.asarray(batches, np.int32), lrs,
                                               np.asarray(test, np.int32), None, obj.name))
            else:
                # every rank runs the augmenter for every task: the two global RNG streams stay rank-identical
                images, labels = obj.arrays()
                samples = [[images[r], labels[r]] for r in rows]           # == task.sample(sess, n)
                train, test = _split_train_test_segmentation(samples, test_shots)
                pool = _ExamplePool()
                batches = [[pool.row(ex[0], ex[1]) for ex in batch]
                           for batch in _mini_batches(train, inner_batch_size, inner_iters, replacement,
                                                      augmenter=self.augmenter, aug_rate=aug_rate)]
                query = [pool.row(ex[0], ex[1]) for ex in test]
                if mine:
                    pi, pl = pool.arrays()
                    task_plans.append(TaskPlan(pi, pl, np.asarray(batches, np.int32), lrs, np.asarray(query, np.int32),
                                               None, obj.name))
            if mine:
                owned.append(i)
            if self.augmenter is not None and len(task_plans) >= eng.n_slots:
                flush()
        flush()
        eng.states[0].copy_(runner.init_state)        # slot 0 doubles as a task slot: put the model state back
        ious = [float(v) for v in gather_owned(ious_local, eng.device)]
        for n, v in zip(names, ious):
            print("Mean task IoU: {}".format(v))
        # the engine state was never touched (tasks ran on slot copies): _full_state.import_variables is a no-op
        return ious, dict(zip(names, ious))

    def _evaluate(self, train_set, test_set, input_ph, label_ph, minimize_op, predictions, inner_batch_size,
                  inner_iters, replacement, verbose=False, save_fine_tuned_checkpoints=False,
                  save_fine_tuned_checkpoints_dir: Optional[str] = None, eval_sample_num: Optional[int] = None,
                  is_training_ph=None, lr_ph=None, lr: Optional[float] = None, drop_rate_ph=None,
                  drop_rate: Optional[float] = None, aug_rate: Optional[float] = None,
                  task_name: Optional[str] = None):
        """Evaluates a single task's train-test split through the Session (reptile.py:235-294)."""
        snap = self._full_state.snapshot()         # device-side stand-in for export_variables (reptile.py:258)
        for inner_iter, batch in enumerate(_mini_batches(train_set, inner_batch_size, num_batches=inner_iters,
                                                         replacement=replacement, augmenter=self.augmenter,
                                                         aug_rate=aug_rate)):
            inputs, labels = zip(*batch)
            if self._pre_step_op:
                self.session.run(self._pre_step_op)
            feed = {input_ph: inputs, label_ph: labels}
            if (lr_ph is not None) and (lr is not None) and (drop_rate_ph is not None) and (drop_rate is not None):
                feed[drop_rate_ph] = drop_rate
                feed[lr_ph] = lr
            elif (lr_ph is not None) and (lr is not None):
                feed[lr_ph] = lr
            elif (lr_ph is not None) and (self.lr_scheduler is not None):
                feed[lr_ph] = self.lr_scheduler.cur_lr(cur_step=inner_iter)
            self.session.run(minimize_op, feed_dict=feed)
        if save_fine_tuned_checkpoints:          # reptile.py:281-285
            from .util import save_fine_tuned_checkpoint
            save_fine_tuned_checkpoint(os.path.join(save_fine_tuned_checkpoints_dir, task_name), self.session,
                                       step=inner_iters - 1, eval_sample_num=eval_sample_num)
        test_preds = self._test_predictions(train_set, test_set, input_ph, predictions, is_training_ph,
                                            task_name=task_name)
        class_iou = [self._iou(test_preds[j], test_set[j][1]) for j in range(len(test_preds))]
        class_iou = np.nanmean(class_iou)
        print("Mean task IoU: {}".format(class_iou))
        self._full_state.restore(snap)             # import_variables(old_vars) (reptile.py:293)
        return class_iou

    # ------------------------------------------------------------------------------------------
    # early stopping / k-shot learning curves (reptile.py:296-480)
    # ------------------------------------------------------------------------------------------
    def _run_minimize(self, minimize_op, input_ph, label_ph, inputs, labels, inner_iter, lr_ph, lr, lr_scheduler,
                      drop_rate_ph, drop_rate):
        feed = {input_ph: inputs, label_ph: labels}
        if (lr_ph is not None) and (lr is not None) and (drop_rate_ph is not None) and (drop_rate is not None):
            feed[drop_rate_ph] = drop_rate
            feed[lr_ph] = lr
        elif (lr_ph is not None) and (lr is not None):
            feed[lr_ph] = lr
        elif (lr_ph is not None) and (lr_scheduler is not None):
            feed[lr_ph] = lr_scheduler.cur_lr(cur_step=inner_iter)
        self.session.run(minimize_op, feed_dict=feed)

    def _early_stopping_learn(self, train_set, val_set, input_ph, label_ph, minimize_op, predictions,
                              inner_batch_size, min_steps, max_steps, replacement, is_training_ph=None, lr_ph=None,
                              lr_scheduler=None, lr=None, drop_rate_ph=None, drop_rate=None, patience=50,
                              inner_iters=None, aug_rate: Optional[float] = None):
        """Estimates the number of steps to take when learning a new task: adapt up to max_steps, predict the
        validation images and score them after EVERY step (one eval-mode forward + fused threshold on the device),
        stop after `patience` non-improving evaluations (reptile.py:443-480)."""
        from .hyperparam_search import EarlyStopper
        del inner_iters
        if lr_scheduler is not None and lr is not None:
            raise ValueError("Only lr_scheduler or lr should be speced. Not both.")
        snap = self._full_state.snapshot()
        early_stopper = EarlyStopper(patience, min_steps=min_steps)
        for inner_iter, batch in enumerate(_mini_batches(train_set, inner_batch_size, num_batches=max_steps,
                                                         replacement=replacement, augmenter=self.augmenter,
                                                         aug_rate=aug_rate)):
            inputs, labels = zip(*batch)
            if self._pre_step_op:
                self.session.run(self._pre_step_op)
            self._run_minimize(minimize_op, input_ph, label_ph, inputs, labels, inner_iter, lr_ph, lr, lr_scheduler,
                               drop_rate_ph, drop_rate)
            test_preds = self._test_predictions(train_set, val_set, input_ph, predictions, is_training_ph)
            ious = [self._iou(test_preds[j], val_set[j][1]) for j in range(len(test_preds))]
            miou = np.nanmean(ious)
            if not early_stopper.continue_training(miou, inner_iter + 1):
                break
        best_num_steps = early_stopper.best_num_steps()
        best_iou = early_stopper.best_metric()
        print("Best iteration found: {}, with mean-IoU {}".format(best_num_steps, best_iou))
        self._full_state.restore(snap)
        return best_num_steps, best_iou

    def evaluate_with_early_stopping(self, dataset, input_ph, label_ph, minimize_op, predictions, num_classes,
                                     num_shots, inner_batch_size, min_steps, max_steps, replacement,
                                     eval_all_tasks=False, num_tasks_to_sample=20,
                                     test_shots=DEFAULT_NUM_TEST_EXAMPLES, is_training_ph=None, lr_ph=None,
                                     lr: Optional[float] = None, drop_rate_ph=None, drop_rate: Optional[float] = None,
                                     aug_rate: Optional[float] = None,
                                     eval_tasks_with_median_early_stopping_iterations: bool = False
                                     ) -> Tuple[List[str], List[int], List[float]]:
        """Samples few-shot tasks, finds each one's best step count by early stopping on its held-out shots and
        returns (task names, best step counts, IoUs) (reptile.py:296-391)."""
        print("Evaluating {} meta-learning.".format(self.meta_fn))
        if eval_all_tasks:
            sampled_tasks = dataset
        else:
            random.shuffle(dataset)
            sampled_tasks = dataset[:num_tasks_to_sample]
        print("Evaluating {} {}-shot tasks.".format(len(sampled_tasks), num_shots))
        task_names, ious = [], []
        if min_steps != max_steps:
            num_steps = []
            for sampled_task in sampled_tasks:
                sampled, task_name = _sample_mini_image_segmentation_dataset(
                    self.session, [sampled_task], num_classes, num_shots + test_shots, return_task_name=True)
                task_names.append(task_name)
                train_set, test_set = _split_train_test_segmentation(sampled, test_shots)
                best_n_steps, best_miou = self._early_stopping_learn(
                    train_set, test_set, input_ph, label_ph, minimize_op, predictions, inner_batch_size,
                    min_steps=min_steps, max_steps=max_steps, replacement=replacement, is_training_ph=is_training_ph,
                    lr_ph=lr_ph, lr_scheduler=self.lr_scheduler, lr=lr, drop_rate_ph=drop_rate_ph,
                    drop_rate=drop_rate, aug_rate=aug_rate)
                ious.append(best_miou)
                num_steps.append(best_n_steps)
            estimated_best_num_steps = int(np.median(num_steps))
        else:
            estimated_best_num_steps = min_steps
            num_steps = [estimated_best_num_steps] * len(sampled_tasks)
        if eval_tasks_with_median_early_stopping_iterations or min_steps == max_steps:
            print("Estimated best number of steps {}".format(estimated_best_num_steps))
            mean_iou_score, task_iou_map = self.evaluate(
                dataset=sampled_tasks, input_ph=input_ph, label_ph=label_ph, minimize_op=minimize_op,
                predictions=predictions, num_classes=num_classes, num_shots=num_shots,
                inner_batch_size=inner_batch_size, inner_iters=estimated_best_num_steps, replacement=replacement,
                eval_all_tasks=eval_all_tasks, num_tasks_to_sample=num_tasks_to_sample, test_shots=test_shots,
                is_training_ph=is_training_ph, lr_ph=lr_ph, lr=lr, drop_rate_ph=drop_rate_ph, drop_rate=drop_rate,
                aug_rate=aug_rate)
            task_names = list(task_iou_map.keys())
            ious = list(task_iou_map.values())
        else:
            mean_iou_score = np.nanmean(ious)
        print("Evaluated {} task/s".format(len(sampled_tasks)))
        print("Mean IoU from train on {} images and evaluate on {} test images: {}".format(num_shots, test_shots,
                                                                                             mean_iou_score))
        return task_names, num_steps, ious

    def evaluate_m_k_shot_ranges_all_tasks(self, tasks, k_range, m, input_ph, label_ph, minimize_op, predictions,
                                           inner_batch_size, inner_iters, replacement, is_training_ph=None,
                                           lr_ph=None, lr=None, test_samples=20, iter_range=DEFAULT_ITER_RANGE,
                                           aug_rate: float = 0.5):
        """m repetitions of the k-shot sweep for every task (reptile.py:393-407)."""
        assert len(iter_range) == len(k_range)
        params = {"input_ph": input_ph, "label_ph": label_ph, "minimize_op": minimize_op, "predictions": predictions,
                  "inner_batch_size": inner_batch_size, "inner_iters": inner_iters, "replacement": replacement,
                  "is_training_ph": is_training_ph, "lr_ph": lr_ph, "lr": lr, "aug_rate": aug_rate}
        ks, results = [], []
        for task in tasks:
            for _ in range(m):
                res = self.evaluate_k_shot_range(task, k_range=k_range, iter_range=iter_range,
                                                 test_samples=test_samples, **params)
                print("k-shot results {}".format({k: r for k, r in zip(k_range, res)}))
                results.extend(res)
                ks.extend(k_range)
        return ks, results

    def evaluate_k_shot_range(self, task, k_range, iter_range=DEFAULT_ITER_RANGE, test_samples=20,
                              early_stopping_min_val_samples=5, esimate_inner_iters_with_early_stoppping: bool = True,
                              **params):
        """k-shot results of ONE task over a range of ks; from k = 2 * early_stopping_min_val_samples on, 20 % of the
        training shots are held out to estimate the step count by early stopping (reptile.py:409-441).  As in the
        reference, the estimated step count then sticks for the larger ks until it is re-estimated."""
        mious = []
        sampled_task, task_name = _sample_mini_image_segmentation_dataset(
            self.session, [task], num_classes=1, num_shots=max(k_range) + test_samples, return_task_name=True)
        training_examples, test_set = _split_train_test_segmentation(sampled_task, test_shots=test_samples)
        for i, k in enumerate(k_range):
            print("Evaluating {}-shot learning".format(k))
            train_set = training_examples[:k]
            if esimate_inner_iters_with_early_stoppping:
                if k >= early_stopping_min_val_samples * 2:
                    val_shots = int(0.2 * k)
                    print("Split training dataset into {} train shots and {} val shots for early stopping to "
                          "estimate number of steps.".format(k - val_shots, val_shots))
                    d_tr, d_val = _split_train_test_segmentation(train_set, test_shots=val_shots)
                    inner_iters, _ = self._early_stopping_learn(d_tr, d_val, min_steps=1, max_steps=500, **params)
                    params["inner_iters"] = inner_iters
            else:
                params["inner_iters"] = iter_range[i]
            mious.append(self._evaluate(train_set, test_set, **params))
        print("Evaluated task {} over k-range {}".format(task_name, k_range))
        return mious

    def _test_predictions(self, train_set, test_set, input_ph, predictions, is_training_ph=None,
                          task_name: Optional[str] = None):
        """reptile.py:482-524."""
        if os.environ.get("SAVE_PREDICTIONS"):
            raise NotImplementedError("SAVE_PREDICTIONS visualisation (utils/viz.py) is out of scope")
        if self._transductive:
            inputs, _ = zip(*test_set)
            feed = {input_ph: inputs}
            if is_training_ph is not None:
                feed[is_training_ph] = False
            return self.session.run(predictions, feed_dict=feed)
        res = []
        for test_sample in test_set:
            inputs, _ = zip(*train_set)
            inputs += (test_sample[0],)
            feed = {input_ph: inputs}
            if is_training_ph is not None:
                feed[is_training_ph] = False
            res.append(self.session.run(predictions, feed_dict=feed)[-1])
        return res

    @staticmethod
    def _iou(prediction: np.ndarray, label: np.ndarray, epsilon: float = 1e-7,
             class_of_interest_channel: Optional[Union[int, slice]] = 1, round_labels: bool = True):
        """IoU of two binary masks of ONE image (reptile.py:526-549)."""
        if len(prediction.shape) > 3:
            raise ValueError("Function is intended for single image masks, not batches.")
        if prediction.shape != label.shape:
            raise ValueError("prediction shape and label shape must be equal but are: {} and {} respectively.".format(
                prediction.shape, label.shape))
        if class_of_interest_channel is not None:
            prediction = prediction[:, :, class_of_interest_channel]
            label = label[:, :, class_of_interest_channel]
        prediction = np.round(prediction)
        if round_labels:
            label = np.round(label)
        intersection = np.logical_and(prediction, label)
        union = np.logical_or(label, prediction)
        return (np.sum(intersection) + epsilon) / (np.sum(union) + epsilon)


def measure(y_in, pred_in, thresh: float = .5):
    """Shaban et al. confusion counts (reptile.py:555-562), kept as the cross-check metric."""
    y, pred = y_in > thresh, pred_in > thresh
    tp = np.logical_and(y, pred).sum()
    tn = np.logical_and(np.logical_not(y), np.logical_not(pred)).sum()
    fp = np.logical_and(np.logical_not(y), pred).sum()
    fn = np.logical_and(y, np.logical_not(pred)).sum()
    return tp, tn, fp, fn


def iou_img(tp, fp, fn):
    return tp / float(max(tp + fp + fn, 1))


class FOMLIS(Gecko):
    """First-order MAML for image segmentation (reptile.py:569-663): the update direction of a task is the
    last inner step, theta_T - theta_{T-1}; with tail_shots the last mini-batch is a held-out tail set."""

    def __init__(self, *args, train_shots: Optional[int] = None, tail_shots: Optional[int] = None,
                 sample_train_val_with_replacement: bool = False, **kwargs):
        super(FOMLIS, self).__init__(*args, **kwargs)
        self.train_shots = train_shots - tail_shots if tail_shots is not None else train_shots
        self.tail_shots = tail_shots
        self.sample_train_val_with_replacement = sample_train_val_with_replacement
        if sample_train_val_with_replacement:
            print("Sampling train val with replacement.")
        self.meta_fn = "FOMAML"
        print("Specializing meta-learner to FOMAML.")

    def train_step(self, dataset, input_ph, label_ph, minimize_op, num_classes, num_shots, inner_batch_size,
                   inner_iters, replacement, meta_step_size, meta_batch_size, verbose=False, lr_ph=None, lr=None):
        if self.fast_path:
            return self._train_step_device(dataset, num_shots, inner_batch_size, inner_iters, replacement,
                                           meta_step_size, meta_batch_size, lr_ph, lr, fomaml=True)
        old_vars = self._model_state.export_variables()
        updates = []
        for _ in range(meta_batch_size):
            mini_dataset = _sample_mini_image_segmentation_dataset(self.session, dataset, num_classes, num_shots)
            for j, batch in enumerate(self._mini_batches(mini_dataset, inner_batch_size, inner_iters, replacement)):
                inputs, labels = zip(*batch)
                if j == inner_iters - 1:
                    last_backup = self._model_state.export_variables()
                if self._pre_step_op:
                    self.session.run(self._pre_step_op)
                feed = {input_ph: inputs, label_ph: labels}
                if (lr_ph is not None) and (lr is not None):
                    feed[lr_ph] = lr
                self.session.run(minimize_op, feed_dict=feed)
            updates.append(subtract_vars(self._model_state.export_variables(), last_backup))
            self._model_state.import_variables(old_vars)
        update = average_vars(updates)
        self._model_state.import_variables(add_vars(old_vars, scale_vars(update, meta_step_size)))

    def _mini_batches(self, mini_dataset, inner_batch_size, inner_iters, replacement):
        """reptile.py:649-663: T-1 batches from the train part, then the un-augmented tail set as the last batch."""
        if self.tail_shots is None:
            for value in _mini_batches(mini_dataset, inner_batch_size, inner_iters, replacement,
                                       augmenter=self.augmenter, aug_rate=self.aug_rate):
                yield value
            return
        if self.sample_train_val_with_replacement:
            train, tail = _sample_train_test_segmentation_with_replacement(mini_dataset, train_shots=self.train_shots,
                                                                           test_shots=self.tail_shots)
        else:
            train, tail = _split_train_test_segmentation(mini_dataset, test_shots=self.tail_shots)
        for batch in _mini_batches(train, inner_batch_size, inner_iters - 1, replacement, augmenter=self.augmenter,
                                   aug_rate=self.aug_rate):
            yield batch
        yield tail

    def _task_batches_for_training(self, rows, inner_batch_size, inner_iters, replacement):
        return list(self._mini_batches(rows, inner_batch_size, inner_iters, replacement))

    def _array_batches_for_training(self, samples, inner_batch_size, inner_iters, replacement):
        return self._mini_batches(samples, inner_batch_size, inner_iters, replacement)

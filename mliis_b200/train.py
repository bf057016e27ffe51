"""Meta-training driver.  Mirror of /root/reference/meta_learners/supervised_reptile/supervised_reptile/train.py
(train_gecko, :18-135): same signature, same loop (linear meta-step-size anneal :90-92, periodic evaluation on
100 sampled train/test tasks :100-121 - whose in-place shuffles advance the shared `random` stream exactly as in
the reference -, checkpoint every 100 meta-iterations :129-131, time_deadline :132-133)."""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np

from .checkpoint import Saver
from .reptile import Gecko
from .util import log_estimated_time_remaining
from .variables import weight_decay


def train_gecko(sess, model, train_set, test_set, save_dir, num_classes=5, num_shots=5, inner_batch_size=5,
                inner_iters=20, replacement=False, meta_step_size=0.1, meta_step_size_final=0.1, meta_batch_size=1,
                meta_iters=10000, eval_inner_batch_size=5, eval_inner_iters=50, eval_interval=10, weight_decay_rate=1,
                time_deadline=None, train_shots=None, transductive=False, meta_fn=Gecko, log_fn=print,
                save_checkpoint_every_n_meta_iters=100, max_checkpoints_to_keep=2, augment=False, lr_scheduler=None,
                lr=None, save_best_seen=False, num_tasks_to_eval=100, aug_rate: Optional[float] = None):
    """Train a model on a dataset."""
    os.makedirs(save_dir, exist_ok=True)
    saver = Saver(model, max_to_keep=max_checkpoints_to_keep)
    best_saver, best_save_dir = None, None
    if save_best_seen:
        best_save_dir = os.path.join(save_dir, "best_eval")
        os.makedirs(best_save_dir, exist_ok=True)
        best_saver = Saver(model, max_to_keep=1)
    best_eval_iou = -np.inf
    pre_step_op = weight_decay(weight_decay_rate) if weight_decay_rate != 1 else None
    reptile = meta_fn(sess, transductive=transductive, pre_step_op=pre_step_op, lr_scheduler=lr_scheduler,
                      augment=augment, aug_rate=aug_rate)
    writers = _summary_writers(save_dir)
    if not getattr(model, "variables_initialized", False):
        print("Initializing variables.")
        model.initialize()

    for i in range(meta_iters):
        begin_time = time.time()
        print("Reptile training step {} of {}".format(i + 1, meta_iters))
        frac_done = i / meta_iters
        print("{} done".format(frac_done))
        cur_meta_step_size = frac_done * meta_step_size_final + (1 - frac_done) * meta_step_size
        print("Current meta-step size: {}".format(cur_meta_step_size))
        reptile.train_step(train_set, model.input_ph, model.label_ph, model.minimize_op, num_classes=num_classes,
                           num_shots=(train_shots or num_shots), inner_batch_size=inner_batch_size,
                           inner_iters=inner_iters, replacement=replacement, meta_step_size=cur_meta_step_size,
                           meta_batch_size=meta_batch_size, lr_ph=model.lr_ph, lr=lr)
        if i % eval_interval == 0:
            print("Evaluating training performance.")
            mean_ious = []
            for dataset, writer in [(train_set, writers[0]), (test_set, writers[1])]:
                mean_iou, _ = reptile.evaluate(dataset, model.input_ph, model.label_ph, model.minimize_op,
                                               model.predictions, num_classes=num_classes, num_shots=num_shots,
                                               inner_batch_size=eval_inner_batch_size, inner_iters=eval_inner_iters,
                                               replacement=replacement, eval_all_tasks=False,
                                               num_tasks_to_sample=num_tasks_to_eval,
                                               save_fine_tuned_checkpoints=False,
                                               is_training_ph=model.is_training_ph, lr_ph=model.lr_ph)
                if writer is not None:
                    writer.add_scalar("IoU", mean_iou, i)
                    writer.add_scalar("meta_step_size", cur_meta_step_size, i)
                    writer.flush()
                mean_ious.append(mean_iou)
            log_fn("Train step %d: train=%f test=%f" % (i, mean_ious[0], mean_ious[1]))
            if save_best_seen and mean_ious[1] > best_eval_iou:
                best_eval_iou = mean_ious[1]
                print("Highest test-set evaluation IoU seen at step {}: {}".format(i, best_eval_iou))
                print("Saving checkpoint to {}.".format(best_save_dir))
                best_saver.save(sess, os.path.join(best_save_dir, "model.ckpt"), global_step=i)
        if i % save_checkpoint_every_n_meta_iters == 0 or i == meta_iters - 1:
            print("Saving checkpoint to {}.".format(save_dir))
            saver.save(sess, os.path.join(save_dir, "model.ckpt"), global_step=i)
        if time_deadline is not None and time.time() > time_deadline:
            break
        log_estimated_time_remaining(begin_time, i, meta_iters)
    return reptile


def _summary_writers(save_dir):
    """tf.summary.FileWriter(save_dir/{train,test}) (train.py:73-74) via torch's tensorboard writer if present."""
    try:
        from torch.utils.tensorboard import SummaryWriter
        return (SummaryWriter(os.path.join(save_dir, "train")), SummaryWriter(os.path.join(save_dir, "test")))
    except Exception:
        return (None, None)

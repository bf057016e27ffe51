"""Small host helpers mirrored from utils/util.py (:42-50 latest_checkpoint, :94-98, :124-136)."""
import os
import re
import time

import numpy as np


def latest_checkpoint(checkpoint_dir: str, ckpt_prefix: str = "model.ckpt", return_relative: bool = True) -> str:
    """First line of <dir>/checkpoint -> '<dir>/model.ckpt-N' (utils/util.py:42-48)."""
    with open(os.path.join(checkpoint_dir, "checkpoint")) as f:
        text = f.readline()
    found = re.findall(re.escape(ckpt_prefix + "-") + r"[0-9]+", text)
    if not found:
        raise FileNotFoundError("no '%s-N' entry in %s/checkpoint" % (ckpt_prefix, checkpoint_dir))
    return os.path.join(checkpoint_dir, found[0])


def log_estimated_time_remaining(start_time, cur_step, total_steps, unit_name="meta-step"):
    elapsed = (time.time() - start_time) / 60.0
    print("This {} took:".format(unit_name), elapsed, "minutes.")
    print("Estimated training hours remaining:%.4f" % ((total_steps - cur_step) * elapsed / 60.0))
    return elapsed


def validate_datasets(args, train_set, val_set, test_set):
    if not args.pretrained and not args.run_k_shot_learning_curves_experiment:
        assert len(train_set) > 0, "Training set must have examples."
    assert len(test_set) > 0, "Test set must have examples."
    if args.eval_val_tasks and val_set is not None and len(val_set) == 0:
        raise ValueError("Val set has no tasks to evaluate")


def ci95(a):
    """95% confidence half-width: 1.96 * std / sqrt(n)  (utils/util.py:133-136)."""
    return 1.96 * np.std(a) / np.sqrt(len(a))


def save_fine_tuned_checkpoint(save_fine_tuned_checkpoint_dir, sess, step=None, eval_sample_num=None):
    """utils/util.py:72-81: dump the adapted variables of ONE task as a TF bundle
    <dir>[/<eval_sample_num>]/model.ckpt-<step>."""
    from .checkpoint import Saver
    if save_fine_tuned_checkpoint_dir is None:
        raise ValueError("Must specify directory in which to save fine-tuned checkpoints if saving them.")
    if eval_sample_num is not None:
        save_fine_tuned_checkpoint_dir = os.path.join(save_fine_tuned_checkpoint_dir, str(eval_sample_num))
    os.makedirs(save_fine_tuned_checkpoint_dir, exist_ok=True)
    Saver(sess.model).save(sess, os.path.join(save_fine_tuned_checkpoint_dir, "model.ckpt"), global_step=step)
    print("Saved fine-tuned checkpoint to {}.".format(save_fine_tuned_checkpoint_dir))

"""Meta-test driver.  Mirror of evaluate_gecko (/root/reference/meta_learners/supervised_reptile/
supervised_reptile/eval.py:18-90): same signature and return value (mean IoU, {task name: [IoU per sample]})."""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Tuple

import numpy as np

from .reptile import Gecko
from .util import ci95
from .variables import weight_decay


def evaluate_gecko(sess, model, dataset, num_classes=1, num_shots=5, eval_inner_batch_size=5, eval_inner_iters=50,
                   replacement=False, num_samples=100, transductive=False, weight_decay_rate=1, meta_fn=Gecko,
                   visualize_predicted_segmentations=True, save_fine_tuned_checkpoints=False,
                   save_fine_tuned_checkpoints_dir: Optional[str] = None, lr_scheduler=None, lr=None, augment=False,
                   serially_eval_all_tasks: bool = False, aug_rate: Optional[float] = None
                   ) -> Tuple[float, Dict[str, List[float]]]:
    """Evaluates an image segmentation model on a dataset."""
    print("Evaluating with eval_inner_iters: {}".format(eval_inner_iters))
    print("Evaluating with lr: {}".format(lr))
    if save_fine_tuned_checkpoints:
        print("Saving fine-tuned checkpoints to {}".format(save_fine_tuned_checkpoints_dir))
    pre_step_op = weight_decay(weight_decay_rate) if weight_decay_rate != 1 else None
    gecko = meta_fn(sess, transductive=transductive, pre_step_op=pre_step_op, lr_scheduler=lr_scheduler,
                    augment=augment, aug_rate=aug_rate)
    mean_ious = []
    task_iou_map: Dict[str, List[float]] = {}
    for i in range(num_samples):
        mean_iou, task_iou_map_i = gecko.evaluate(
            dataset, model.input_ph, model.label_ph, model.minimize_op, model.predictions, num_classes=num_classes,
            num_shots=num_shots, inner_batch_size=eval_inner_batch_size, inner_iters=eval_inner_iters,
            replacement=replacement, eval_all_tasks=serially_eval_all_tasks,
            save_fine_tuned_checkpoints=save_fine_tuned_checkpoints,
            save_fine_tuned_checkpoints_dir=save_fine_tuned_checkpoints_dir, eval_sample_num=i,
            is_training_ph=model.is_training_ph, lr_ph=model.lr_ph, lr=lr)
        for key, val in task_iou_map_i.items():
            task_iou_map.setdefault(key, []).append(val)
        mean_ious.append(mean_iou)
    all_ious = list(itertools.chain(*task_iou_map.values()))
    ninety_five_perc_ci = ci95(all_ious)
    print("Mean of all {} task-splits: {} +/- 95% CI: {}".format(len(all_ious), np.nanmean(all_ious),
                                                                ninety_five_perc_ci))
    print("{} NaN values out of total number of samples: {}".format(np.count_nonzero(np.isnan(mean_ious)), num_samples))
    mean_iou = np.nanmean(mean_ious)
    print("Mean of samples:")
    print("{} mean IoU, +/- 95% CI: {}".format(mean_iou, ninety_five_perc_ci))
    print("Evaluated with eval_inner_iters: {}".format(eval_inner_iters))
    print("Evaluated with lr: {}".format(lr))
    return mean_iou, task_iou_map


def optimize_update_hyperparams(*args, **kwargs):
    raise NotImplementedError("update-hyperparameter search (eval.py:93; needs skopt) is out of scope - SURVEY 8f-4")


def run_k_shot_learning_curves_experiment(*args, **kwargs):
    raise NotImplementedError("k-shot learning curves (eval.py:190) are out of scope - SURVEY 8f-4")

"""Command-line surface of run_metasegnet.py, kept flag-for-flag compatible with the reference
(/root/reference/meta_learners/args.py:16-118) so existing launch scripts (run.sh:8-17) keep working.

The flags are declared as a table; the kwargs builders produce the same dictionaries as the reference's
model_kwargs (:121-160), train_kwargs (:187-212) and evaluate_kwargs (:215-236), except that the optimizer
is named ('adam' | 'sgd') instead of being a TensorFlow class (args.py:151-154).
"""
from __future__ import annotations

import argparse
from functools import partial

from .lr_schedulers import supported_learning_rate_schedulers

SUPPORTED_MODELS = {"efficientlab"}          # models/constants.py:1
SUPPORTED_SEARCH_ALGS = {"GP", "GBRT", "RF", "ET", "dummy"}

_S, _I, _F = str, int, float
# (flag, kind, default, extra)   kind: a type, 'flag' (store_true), or ('list', type)
_FLAGS = [
    ("--fine-tune-task", _S, None, {}), ("--fine-tuned-checkpoint", _S, None, {}),
    ("--pretrained", "flag", False, {}), ("--seed", _I, 0, {}), ("--checkpoint", _S, "model_checkpoint", {}),
    ("--classes", _I, 1, {}), ("--shots", _I, 5, {}), ("--train-shots", _I, 5, {}), ("--inner-batch", _I, 8, {}),
    ("--inner-iters", _I, 8, {}), ("--replacement", "flag", False, {}), ("--learning-rate", _F, 1e-3, {}),
    ("--meta-step", _F, 0.1, {}), ("--meta-step-final", _F, 0.1, {}), ("--meta-batch", _I, 5, {}),
    ("--meta-iters", _I, 400000, {}), ("--eval-batch", _I, 8, {}), ("--eval-iters", _I, 4, {}),
    ("--eval-samples", _I, 200, {}), ("--eval-interval", _I, 10, {}), ("--weight-decay", _F, 1, {}),
    ("--transductive", "flag", False, {}), ("--foml", "flag", False, {}), ("--foml-tail", _I, None, {}),
    ("--sgd", "flag", False, {}), ("--n_unet_encoding_stacks", _I, 4, {}), ("--data-dir", _S, None, {}),
    ("--loss_name", _S, "cross_entropy", {}), ("--save_fine_tuned_checkpoints", "flag", False, {}),
    ("--save_fine_tuned_checkpoints_train", "flag", False, {}),
    ("--save_fine_tuned_checkpoints_dir", _S, "/tmp/checkpoints/fine-tuned", {}),
    ("--model_name", _S, "efficientlab", {}), ("--start_num_feature_maps_power", _I, 5, {}),
    ("--restore_efficient_net_weights_from", _S, None, {}), ("--spatial_pyramid_pooling", "flag", False, {}),
    ("--skip_decoding", "flag", False, {}), ("--rsd", ("list", _I), None, {}),
    ("--feature_extractor_name", _S, "efficientnet-b0", {}), ("--learning_rate_scheduler", _S, "fixed", {}),
    ("--step_decay_rate", _F, 0.5, {}), ("--decay_after_n_steps", _I, 5, {}), ("--l2", "flag", False, {}),
    ("--l1", "flag", False, {}), ("--darc1", "flag", False, {}), ("--augment", "flag", False, {}),
    ("--final_layer_dropout_rate", _F, 0.0, {}), ("--image_size", _I, 320, {}), ("--label_smoothing", _F, 0.0, {}),
    ("--continue_training_from_checkpoint", _S, None, {}), ("--fss_1000", "flag", False, {}),
    ("--num_val_tasks", _I, 0, {}), ("--eval_val_tasks", "flag", False, {}),
    ("--serially_eval_all_test_tasks", "flag", False, {}), ("--optimize_update_hyperparms_on_val_set", "flag", False, {}),
    ("--num_configs_to_sample", _I, 100, {}), ("--meta_fine_tune_steps_on_train_val", _I, 0, {}),
    ("--uho_outer_iters", _I, 2, {}), ("--lr_search_range_low", _F, 0.0005, {}), ("--lr_search_range_high", _F, 0.05, {}),
    ("--drop_rate_search_range_low", _F, 0.2, {}), ("--drop_rate_search_range_high", _F, 0.2, {}),
    ("--aug_rate_search_range_low", _F, 0.5, {}), ("--aug_rate_search_range_high", _F, 0.5, {}),
    ("--batch_size_search_range_low", _I, 8, {}), ("--batch_size_search_range_high", _I, 8, {}),
    ("--run_k_shot_learning_curves_experiment", "flag", False, {}), ("--fp_k_test_set", "flag", False, {}),
    ("--disable_rsd_residual_connections", "flag", False, {}), ("--do_not_restore_final_layer_weights", "flag", False, {}),
    ("--eval_tasks_with_median_early_stopping_iterations", "flag", False, {}), ("--min_steps", _I, 0, {}),
    ("--max_steps", _I, 80, {}), ("--k_shot_iter_range", ("list", _I), None, {}),
    ("--sample_foml_train_val_with_replacement", "flag", False, {}), ("--aug_rate", _F, 0.5, {}),
    ("--uho_results_csv_name", _S, "val-set_hyper_param_search_results.csv", {}), ("--uho_estimator", _S, "GP", {}),
    # ---- additions of this implementation (absent from the reference) ----
    ("--gemm_mode", _S, "tf32x3", {"help": "numeric mode of the dense contractions: tf32x3 (tcgen05, fp32-class accuracy; "
                                          "default) | tf32 | fp32 (FFMA reference mode)"}),
    ("--task_slots", _I, 32, {"help": "concurrent task slots per GPU on the device fast path (launched as two or more "
                                       "task-batched groups: 32 -> 2 x 16; 0.56 GB of workspace per slot at 224x224)"}),
    ("--meta_task_slots", _I, 1, {"help": "meta-training: task slots adapting the tasks of a meta-batch concurrently "
                                          "(1 = the reference's sequential order; S > 1 = per-slot optimizer state, "
                                          "like S ranks)"}),
    ("--synthetic_tasks", _I, 0, {"help": "use N synthetic FSS-1000-shaped test tasks instead of tfrecord shards"}),
]


def argument_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for flag, kind, default, extra in _FLAGS:
        if kind == "flag":
            parser.add_argument(flag, action="store_true", default=default, **extra)
        elif isinstance(kind, tuple):
            parser.add_argument(flag, type=kind[1], nargs="+", default=default, **extra)
        else:
            parser.add_argument(flag, type=kind, default=default, **extra)
    return parser


def _args_meta_fn(pa):
    from .reptile import FOMLIS, Gecko
    if pa.foml:
        return partial(FOMLIS, train_shots=pa.train_shots, tail_shots=pa.foml_tail,
                       sample_train_val_with_replacement=pa.sample_foml_train_val_with_replacement)
    return Gecko


def model_kwargs(pa) -> dict:
    pa.model_name = pa.model_name.lower()
    if pa.model_name not in SUPPORTED_MODELS:
        raise ValueError("Model name must be in the set: {} but is {}".format(SUPPORTED_MODELS, pa.model_name))
    kw = {"learning_rate": pa.learning_rate}
    if pa.model_name == "efficientlab":
        kw["restore_ckpt_dir"] = pa.restore_efficient_net_weights_from
        if pa.spatial_pyramid_pooling:
            kw["spatial_pyramid_pooling"] = True
        if pa.skip_decoding:
            kw["skip_decoding"] = True
        kw["rsd"] = pa.rsd if pa.rsd else None
        for name in ("feature_extractor_name", "l2", "l1", "darc1", "final_layer_dropout_rate", "label_smoothing"):
            kw[name] = getattr(pa, name)
        if "dice" not in pa.loss_name:       # only `dice` is toggled by --loss_name (args.py:147-148)
            kw["dice"] = False
        if pa.disable_rsd_residual_connections:
            # the reference stores this under a key its constructor ignores (args.py:149-150 vs
            # efficientlab.py:27): kept as the same silent no-op
            kw["disable_rsd_residual_connections"] = True
    kw["optimizer"] = "sgd" if pa.sgd else "adam"
    kw["loss_name"] = pa.loss_name
    kw["n_unet_encoding_stacks"] = pa.n_unet_encoding_stacks
    kw["start_num_feature_maps_power"] = pa.start_num_feature_maps_power
    kw["n_rows"] = kw["n_cols"] = pa.image_size
    kw["gemm_mode"] = getattr(pa, "gemm_mode", "tf32x3")
    kw["task_slots"] = getattr(pa, "task_slots", 32)
    return kw


def hyper_search_kwargs(pa) -> dict:
    assert pa.uho_estimator in SUPPORTED_SEARCH_ALGS
    keys = ["lr_search_range_low", "lr_search_range_high", "drop_rate_search_range_low", "drop_rate_search_range_high",
            "aug_rate_search_range_low", "aug_rate_search_range_high", "batch_size_search_range_low",
            "batch_size_search_range_high"]
    kw = {k: getattr(pa, k) for k in keys}
    kw["estimator"] = pa.uho_estimator
    return kw


def optim_kwargs(pa) -> dict:
    return {"learning_rate": pa.learning_rate, "label_smoothing": pa.label_smoothing}


def train_kwargs(pa) -> dict:
    if pa.learning_rate_scheduler not in supported_learning_rate_schedulers:
        raise ValueError("Learning rate scheduler, {}, not in supported set: {}".format(
            pa.learning_rate_scheduler, supported_learning_rate_schedulers.keys()))
    return {
        "num_classes": pa.classes, "num_shots": pa.shots, "train_shots": (pa.train_shots or None),
        "inner_batch_size": pa.inner_batch, "inner_iters": pa.inner_iters, "replacement": pa.replacement,
        "meta_step_size": pa.meta_step, "meta_step_size_final": pa.meta_step_final, "meta_batch_size": pa.meta_batch,
        "meta_iters": pa.meta_iters, "eval_inner_batch_size": pa.eval_batch, "eval_inner_iters": pa.eval_iters,
        "eval_interval": pa.eval_interval, "weight_decay_rate": pa.weight_decay, "transductive": pa.transductive,
        "meta_fn": _args_meta_fn(pa), "aug_rate": pa.aug_rate,
    }


def evaluate_kwargs(pa) -> dict:
    return {
        "num_classes": pa.classes, "num_shots": pa.shots, "eval_inner_batch_size": pa.eval_batch,
        "eval_inner_iters": pa.eval_iters, "replacement": pa.replacement, "weight_decay_rate": pa.weight_decay,
        "num_samples": pa.eval_samples, "transductive": pa.transductive,
        "save_fine_tuned_checkpoints": pa.save_fine_tuned_checkpoints,
        "save_fine_tuned_checkpoints_dir": pa.save_fine_tuned_checkpoints_dir, "meta_fn": _args_meta_fn(pa),
        "augment": pa.augment, "lr": None,
        "eval_tasks_with_median_early_stopping_iterations": pa.eval_tasks_with_median_early_stopping_iterations,
        "aug_rate": pa.aug_rate,
    }

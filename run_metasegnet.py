#!/usr/bin/env python
"""Meta-trains and evaluates image segmentation models on the B200 engine.

Drop-in for /root/reference/run_metasegnet.py: same flags (mliis_b200/args.py), same flow (:28-211) - build the
model, load the tasks, restore the checkpoint or meta-train, evaluate on the training and meta-test tasks, print
the greppable summary line and write <checkpoint>/meta-test_results.json.

Differences: tasks come from the gzip-TFRecord shards under --data-dir when there are any (read without
TensorFlow, mliis_b200/fss1000.py), else they are synthetic FSS-1000-shaped tasks (`--synthetic_tasks N`); the
update-hyperparameter search uses a scikit-learn GP instead of scikit-optimize (mliis_b200/hyperparam_search.py).
Multi-GPU: launch with torchrun - tasks are sharded across ranks (mliis_b200/reptile.py).
"""
import datetime
import json
import os
import random

import numpy as np


def main():
    from mliis_b200.args import argument_parser, evaluate_kwargs, hyper_search_kwargs, model_kwargs, train_kwargs
    from mliis_b200.efficientlab import EfficientLab
    from mliis_b200.eval import (evaluate_gecko, optimize_update_hyperparams,
                                 run_k_shot_learning_curves_experiment)
    from mliis_b200.lr_schedulers import supported_learning_rate_schedulers
    from mliis_b200.fss1000 import get_fss_tasks, read_fp_k_shot_dataset, read_fss_1000_dataset
    from mliis_b200.metaseg import read_synthetic_dataset
    from mliis_b200.session import Session
    from mliis_b200.train import train_gecko
    from mliis_b200.checkpoint import Saver
    from mliis_b200.util import latest_checkpoint, validate_datasets

    verbose = True
    eval_train_tasks = True
    start_time = datetime.datetime.now()
    print("Experiment started at: {}".format(start_time))
    args = argument_parser().parse_args()
    if args.optimize_update_hyperparms_on_val_set:
        assert args.num_val_tasks > 0, \
            "Must specify number of validation tasks greater than 0 to optimize update hyperparams."

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")

    os.environ["MLIIS_META_TASK_SLOTS"] = str(getattr(args, "meta_task_slots", 1))   # read by Gecko / FOMLIS
    random.seed(args.seed)          # the ONLY seed the reference sets (run_metasegnet.py:43)
    print("Defining model architecture:")
    mk = model_kwargs(args)
    print("Using loss {}".format(mk["loss_name"]))
    restore_ckpt_dir = mk["restore_ckpt_dir"]
    model = EfficientLab(**mk)
    lr_scheduler = None
    sched_cls = supported_learning_rate_schedulers[args.learning_rate_scheduler]
    if sched_cls is not None:
        sk = {"decay_rate": args.step_decay_rate, "decay_after_n_steps": args.decay_after_n_steps} \
            if "step" in args.learning_rate_scheduler else {}
        lr_scheduler = sched_cls(args.learning_rate, train_kwargs(args)["eval_inner_iters"], **sk)
    print("{} instantiated.".format(args.model_name))
    print("Model contains {} trainable parameters.".format(model.n_params))
    print("Meta-learning with algorithm:")
    print("FOMAML" if args.foml else "Reptile")

    print("Setting up meta-learning dataset")
    data_dir = getattr(args, "data_dir", None)
    have_shards = bool(data_dir) and os.path.isdir(data_dir) and len(get_fss_tasks(data_dir)) > 0
    if have_shards and not args.synthetic_tasks and args.run_k_shot_learning_curves_experiment:
        test_set, test_task_names = read_fp_k_shot_dataset(data_dir, image_size=args.image_size)   # :80-83
        train_set, val_set = None, None
    elif have_shards and not args.synthetic_tasks:
        # run_metasegnet.py:84-96: gzip-TFRecord shards, one per class (mliis_b200/fss1000.py, no TensorFlow)
        if args.fp_k_test_set:
            ids_file = os.path.join(data_dir, "fp-k_test_set.txt")
            if not os.path.exists(ids_file):
                raise FileNotFoundError("--fp_k_test_set needs %s (the reference's data/fp-k_test_set.txt)" % ids_file)
            print("Holding out FP-k classes listed in {}".format(ids_file))
            dataset = read_fss_1000_dataset(data_dir, num_val_tasks=args.num_val_tasks, test_task_ids=ids_file,
                                            image_size=args.image_size)
        else:
            dataset = read_fss_1000_dataset(data_dir, num_val_tasks=args.num_val_tasks, image_size=args.image_size)
        train_set, val_set, test_set, _, _, test_task_names = dataset
    else:
        n_test = args.synthetic_tasks or 240
        train_set, val_set, test_set, _, _, test_task_names = read_synthetic_dataset(
            num_train_tasks=max(8, 760 if not args.synthetic_tasks else 4 * n_test), num_test_tasks=n_test,
            n_examples=max(10, args.shots + 5, (args.train_shots or 0)), image_size=args.image_size)
    if not val_set:
        val_set = None
    validate_datasets(args, train_set, val_set, test_set)
    if verbose:
        print("Found {} testing tasks:".format(len(test_set)))
        if train_set is not None:
            print("Found {} training tasks:".format(len(train_set)))

    with Session(model) as sess:
        if restore_ckpt_dir is not None and not args.pretrained:
            print("Restoring from checkpoint {}".format(restore_ckpt_dir))
            model.restore_model(sess, restore_ckpt_dir, filter_to_scopes=[model.feature_extractor_name])
        if not args.pretrained:
            print("Meta-training...")
            if args.continue_training_from_checkpoint is not None:
                ckpt = latest_checkpoint(args.continue_training_from_checkpoint)
                print("Continuing meta-training from checkpoint: {}".format(ckpt))
                Saver(model).restore(sess, ckpt)
            train_gecko(sess, model, train_set, val_set or test_set, args.checkpoint, lr_scheduler=lr_scheduler,
                        augment=args.augment, **train_kwargs(args))
        else:
            if args.do_not_restore_final_layer_weights:
                print("Restoring from checkpoint: {}".format(args.checkpoint))
                model.restore_model(sess, args.checkpoint, filter_out_scope=model.final_layer_scope,
                                    convert_ckpt_to_rel_path=True)
            else:
                ckpt = latest_checkpoint(args.checkpoint)
                print("Restoring from checkpoint: {}".format(ckpt))
                Saver(model).restore(sess, ckpt)

        eval_kwargs = evaluate_kwargs(args)
        if args.optimize_update_hyperparms_on_val_set:                       # run_metasegnet.py:136-165
            print("Optimizing the update routine hyperparams on the val set")
            assert val_set and len(val_set) > 0, "Dev set has no tasks"
            keep = eval_kwargs["save_fine_tuned_checkpoints"]
            eval_kwargs["save_fine_tuned_checkpoints"] = False
            # (the reference also passes b=args.uho_outer_iters, which its own function does not accept)
            estimated_lr, estimated_steps = optimize_update_hyperparams(
                sess, model, val_set, lr_scheduler=lr_scheduler,
                serially_eval_all_tasks=args.serially_eval_all_test_tasks,
                num_configs_to_sample=args.num_configs_to_sample, save_dir=args.checkpoint,
                results_csv_name=args.uho_results_csv_name,
                num_train_val_data_splits_to_sample_per_config=1 if args.fss_1000 else 4,
                max_steps=args.max_steps, min_steps=args.min_steps, **eval_kwargs, **hyper_search_kwargs(args))
            eval_kwargs["save_fine_tuned_checkpoints"] = keep
            eval_kwargs["eval_inner_iters"] = estimated_steps
            eval_kwargs["lr"] = estimated_lr
            if args.meta_fine_tune_steps_on_train_val > 0:
                print("Fine-tuning meta-learned init for {} meta-steps with optimized hyperparameters.".format(
                    args.meta_fine_tune_steps_on_train_val))
                tp = train_kwargs(args)
                tp["inner_iters"], tp["lr"] = estimated_steps, estimated_lr
                tp["meta_step_size"] = tp["meta_step_size_final"]
                train_gecko(sess, model, train_set + val_set, test_set,
                            os.path.join(args.checkpoint, "fine-tuned_on_train_val_with_optimized_update_hyperparams"),
                            lr_scheduler=lr_scheduler, augment=args.augment, **tp)
        del eval_kwargs["eval_tasks_with_median_early_stopping_iterations"]
        if args.run_k_shot_learning_curves_experiment:                       # run_metasegnet.py:167-171
            kk = dict(eval_kwargs)
            del kk["save_fine_tuned_checkpoints"]
            del kk["save_fine_tuned_checkpoints_dir"]
            run_k_shot_learning_curves_experiment(sess, model, test_set, lr_scheduler=lr_scheduler,
                                                  iter_range=args.k_shot_iter_range, **kk)
            print("Experiment finished at: {}".format(datetime.datetime.now()))
            return
        if args.eval_val_tasks and val_set:
            test_set = val_set
        print("Evaluating {}-shot learning on training tasks.".format(args.shots))
        mean_train_iou = None
        if eval_train_tasks:
            keep = eval_kwargs["save_fine_tuned_checkpoints"]
            eval_kwargs["save_fine_tuned_checkpoints"] = args.save_fine_tuned_checkpoints_train
            mean_train_iou, _ = evaluate_gecko(sess, model, train_set, visualize_predicted_segmentations=False,
                                               lr_scheduler=lr_scheduler, serially_eval_all_tasks=False, **eval_kwargs)
            eval_kwargs["save_fine_tuned_checkpoints"] = keep
        print("Evaluating {}-shot learning on meta-test tasks.".format(args.shots))
        mean_test_iou, task_name_iou_map = evaluate_gecko(
            sess, model, test_set, visualize_predicted_segmentations=False, lr_scheduler=lr_scheduler,
            serially_eval_all_tasks=args.serially_eval_all_test_tasks, **eval_kwargs)
        print("Evaluated meta-test tasks:")
        print(task_name_iou_map)
        if eval_train_tasks:
            print("Mean meta-train IoU: {}".format(mean_train_iou))
        # Do NOT change this print (it's used to grep logs)  -- run_metasegnet.py:199-200
        print("Mean IoU over all meta-test tasks: {}".format(mean_test_iou))
        if int(os.environ.get("RANK", "0")) == 0:
            os.makedirs(args.checkpoint, exist_ok=True)
            results_path = os.path.join(args.checkpoint, "meta-test_results.json")
            with open(results_path, "w") as f:
                json.dump(task_name_iou_map, f)
            print("Wrote results to {}".format(results_path))

    end_time = datetime.datetime.now()
    print("Experiment finished at: {}, taking {}".format(end_time, end_time - start_time))


if __name__ == "__main__":
    main()

"""Early stopping and update-hyperparameter optimisation (UHO) around the adaptation hot path.

Mirror of /root/reference/meta_learners/hyperparam_search.py:
  EarlyStopper :24-68 · run_m :71-92 · save_results :95-131 · compute_best_configuration :134-157 ·
  gp_update_hyperparameter_optimization :185-251 · lr_droprate_aug_rate_batch_size_gp_search :254-288

The reference drives the search with ``skopt.Optimizer("GP", acq_func="EI")``; scikit-optimize is not in this image,
so ``GPSearch`` below is a small ask/tell optimiser with the same contract built on scikit-learn's
``GaussianProcessRegressor`` (Matern-5/2 + white noise, expected improvement maximised over random candidates;
log-uniform reals, inclusive integers, ``n_initial_points`` random draws first).  It is a functional equivalent, not a
bit-identical replay of skopt's internal sampling.
"""
from __future__ import annotations

import csv
import operator
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

DROPOUT_RATE_NAME = "drop_rate"
AUG_RATE_NAME = "aug_rate"
BATCH_SIZE_NAME = "inner_batch_size"
LEARNING_RATE_NAME = "lr"
SUPPORTED_SEARCH_ALGS = {"GP"}


class EarlyStopper:
    """Stopping criterion from a metric and a patience (number of non-improving evaluations tolerated)."""

    def __init__(self, patience: int = 10, metric_should_increase: bool = True, min_steps: int = 0):
        self.patience = patience
        self.metric_should_increase = metric_should_increase
        self.eval_operator = operator.gt if metric_should_increase else operator.lt
        self._best_metric = None
        self._best_num_steps = min_steps if min_steps > 0 else None
        self.num_evals_without_improving = 0
        self.min_steps = min_steps
        print("Built EarlyStopper with patience {}".format(self.patience))

    def continue_training(self, metric, total_steps_taken) -> bool:
        if total_steps_taken <= self.min_steps:
            self._best_metric = metric          # warm-up: the latest metric is the one to beat
            return True
        if self._best_metric is None or self.eval_operator(metric, self._best_metric):
            self.num_evals_without_improving = 0
            self._best_metric = metric
            self._best_num_steps = total_steps_taken
            return True
        self.num_evals_without_improving += 1
        return self.num_evals_without_improving <= self.patience

    def best_metric(self):
        return self._best_metric

    def best_num_steps(self):
        return self._best_num_steps


def run_m(eval_fn: Callable, params: Dict, m: int = 1):
    """Calls eval_fn(**params) m times; concatenates its (task ids, best step counts, metrics) lists."""
    ids, steps, metrics = [], [], []
    for _ in range(m):
        a, b, c = eval_fn(**params)
        ids.extend(a)
        steps.extend(b)
        metrics.extend(c)
    return ids, steps, metrics


def save_results(results: List[Tuple[Dict, Tuple[List, List, List]]], path: str, metric_name: str = "mIoU",
                 append_if_exists: bool = False) -> str:
    """One csv row per (configuration, task); an existing file is appended to or side-stepped with a numeric suffix."""
    print("Saving results to {}".format(path))
    columns: Dict[str, list] = {"task_ID": [], "best_num_steps": [], metric_name: []}
    for config, (task_ids, num_steps, metrics) in results:
        for key, val in config.items():
            columns.setdefault(key, []).extend([val] * len(task_ids))
        columns["task_ID"].extend(task_ids)
        columns["best_num_steps"].extend(num_steps)
        columns[metric_name].extend(metrics)
    mode, header = "w", True
    if os.path.exists(path):
        if append_if_exists:
            mode, header = "a", False
        else:
            i = 0
            while os.path.exists(path + "_{}".format(i)):
                i += 1
            path = path + "_{}".format(i)
    names = list(columns)
    with open(path, mode, newline="") as f:
        w = csv.writer(f)
        if header:
            w.writerow(names)
        for row in zip(*(columns[n] for n in names)):
            w.writerow(row)
    print("Saved optimization raw results to {}".format(path))
    return path


def compute_best_configuration(results_list, metric_should_increase=True):
    better = operator.gt if metric_should_increase else operator.lt
    best_metric = -np.inf if metric_should_increase else np.inf
    best_config, best_step_num = None, None
    for sampled_config, (task_ids, num_steps, metrics) in results_list:
        mean_metric = np.mean(metrics)
        if better(mean_metric, best_metric):
            best_config, best_metric, best_step_num = sampled_config, mean_metric, np.median(num_steps)
    print("Best mIoU found: {}".format(best_metric))
    print("with median iteration: {}".format(best_step_num))
    print("and config: {}".format(best_config))
    return best_config, int(best_step_num), best_metric


def log_opt_progress(hyperparams, results_i, task_ids, num_steps, metrics, save_results_to):
    print("Results for hyperparams {}: task IDs: {}, best num steps: {}, mIoUs: {}".format(hyperparams, task_ids,
                                                                                         num_steps, metrics))
    print("mean mIoU: {}".format(np.nanmean(metrics)))
    if save_results_to is not None:
        save_results([results_i], save_results_to, append_if_exists=True)


def insert_sampled_into_full_set_of_hyperparams(sampled, hyperparams) -> Dict:
    hyperparams.update(sampled)
    return hyperparams


class GPSearch:
    """ask/tell minimiser over box-bounded dimensions (stand-in for skopt.Optimizer("GP", acq_func="EI")).

    dims: list of (name, low, high, kind) with kind "real" (log-uniform when low > 0, like the reference's
    prior="log-uniform") or "int" (inclusive)."""

    def __init__(self, dims: Sequence[Tuple[str, Any, Any, str]], n_initial_points: int = 5, seed: Optional[int] = None,
                 n_candidates: int = 2000):
        self.dims = list(dims)
        self.n_initial_points = n_initial_points
        self.rng = np.random.default_rng(seed)
        self.n_candidates = n_candidates
        self.X: List[List[float]] = []
        self.y: List[float] = []

    def _to_unit(self, x) -> np.ndarray:
        u = []
        for (name, lo, hi, kind), v in zip(self.dims, x):
            if kind == "real" and lo > 0:
                u.append((np.log(v) - np.log(lo)) / (np.log(hi) - np.log(lo)))
            else:
                u.append((float(v) - lo) / float(hi - lo))
        return np.asarray(u, np.float64)

    def _from_unit(self, u) -> list:
        x = []
        for (name, lo, hi, kind), t in zip(self.dims, u):
            t = float(min(max(t, 0.0), 1.0))
            if kind == "int":
                x.append(int(min(hi, max(lo, int(np.floor(lo + t * (hi - lo + 1)))))))
            elif lo > 0:
                x.append(float(np.exp(np.log(lo) + t * (np.log(hi) - np.log(lo)))))
            else:
                x.append(float(lo + t * (hi - lo)))
        return x

    def ask(self) -> list:
        d = len(self.dims)
        if len(self.X) < self.n_initial_points or d == 0:
            return self._from_unit(self.rng.random(d))
        from scipy.stats import norm
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
        U = np.stack([self._to_unit(x) for x in self.X])
        y = np.asarray(self.y, np.float64)
        scale = y.std() if y.std() > 0 else 1.0
        yn = (y - y.mean()) / scale
        kernel = ConstantKernel(1.0) * Matern(length_scale=np.full(d, 0.3), nu=2.5) + WhiteKernel(1e-3)
        gp = GaussianProcessRegressor(kernel=kernel, normalize_y=False, n_restarts_optimizer=2,
                                      random_state=int(self.rng.integers(1 << 31)))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")     # sklearn's kernel-bound ConvergenceWarnings on tiny samples
            gp.fit(U, yn)
        cand = self.rng.random((self.n_candidates, d))
        mu, sd = gp.predict(cand, return_std=True)
        best = yn.min()
        sd = np.maximum(sd, 1e-12)
        z = (best - mu) / sd
        ei = (best - mu) * norm.cdf(z) + sd * norm.pdf(z)
        return self._from_unit(cand[int(np.argmax(ei))])

    def tell(self, x, objective: float) -> None:
        self.X.append(list(x))
        self.y.append(float(objective))


def gp_update_hyperparameter_optimization(eval_fn: Callable, hyperparams: Dict, search_key_ranges: Dict, n: int,
                                          save_results_to: Optional[str] = "gp_hyper_param_search_results.csv",
                                          m: int = 1, metric_should_increase: bool = True, metric_name: str = "mIoU",
                                          base: int = 2, n_initial_points: Optional[int] = None,
                                          prior: str = "log-uniform", seed: Optional[int] = None):
    """GP regression of the metric over the ranges in ``search_key_ranges`` (keys must exist in ``hyperparams``);
    degenerate ranges (low == high) are not searched.  Returns (best config, median best step count, best metric,
    all results)."""
    for key in search_key_ranges:
        assert key in hyperparams, "key: {} not in hyperparams: {}".format(key, hyperparams)
    if n_initial_points is None:
        n_initial_points = int(n / 2)
    print("Sampling {} points initially at random.".format(n_initial_points))
    dims = []
    for key, (lo, hi) in search_key_ranges.items():
        if lo == hi:
            continue
        if isinstance(lo, float):
            dims.append((key, lo, hi, "real"))
        elif isinstance(lo, int):
            dims.append((key, lo, hi, "int"))
        else:
            raise ValueError("Value must be float, int, or str, but {} is {}".format(lo, type(lo)))
    opt = GPSearch(dims, n_initial_points=n_initial_points, seed=seed)
    results = []
    for i in range(n):
        print("Running configuration sample {} of {}.".format(i + 1, n))
        sampled_list = opt.ask()
        sampled = {d[0]: x for d, x in zip(dims, sampled_list)}
        print("With sampled hyperparams:")
        print(sampled)
        hyperparams = insert_sampled_into_full_set_of_hyperparams(sampled, hyperparams)
        task_ids, num_steps, metrics = run_m(eval_fn, hyperparams, m)
        objective = np.nanmean(metrics)
        if metric_should_increase:
            objective *= -1
        print("Objective value at sample {} of {}: {}".format(i + 1, n, objective))
        opt.tell(sampled_list, objective)
        results_i = (sampled, (task_ids, num_steps, metrics))
        results.append(results_i)
        log_opt_progress(hyperparams, results_i, task_ids, num_steps, metrics, save_results_to)
    best_config, expected_best_step_num, best_metric = compute_best_configuration(results, metric_should_increase)
    return best_config, expected_best_step_num, best_metric, results


def _ordered(lo, hi, cast):
    lo, hi = cast(lo), cast(hi)
    return [hi, lo] if lo > hi else [lo, hi]


def lr_droprate_aug_rate_batch_size_gp_search(eval_fn: Callable, params: Dict, lr_name: str = LEARNING_RATE_NAME,
                                              lr_search_range_low: float = 0.0005, lr_search_range_high: float = 0.05,
                                              droprate_name: str = DROPOUT_RATE_NAME,
                                              drop_rate_search_range_low: float = 0.2,
                                              drop_rate_search_range_high: float = 0.2,
                                              aug_rate_name: str = AUG_RATE_NAME, aug_rate_search_range_low: float = 0.5,
                                              aug_rate_search_range_high: float = 0.5,
                                              batch_size_name: str = BATCH_SIZE_NAME,
                                              batch_size_search_range_low: int = 8,
                                              batch_size_search_range_high: int = 8, n: int = 100,
                                              save_results_to: str = "hyper_param_search_results.csv", m: int = 1,
                                              metric_should_increase: bool = True, metric_name: str = "mIoU",
                                              seed: Optional[int] = None) -> Tuple[float, int]:
    """Joint search over learning rate, final-layer dropout rate, augmentation rate and inner batch size; returns the
    best learning rate and the expected (median) number of adaptation steps."""
    ranges = {lr_name: _ordered(lr_search_range_low, lr_search_range_high, float),
              droprate_name: _ordered(drop_rate_search_range_low, drop_rate_search_range_high, float),
              aug_rate_name: _ordered(aug_rate_search_range_low, aug_rate_search_range_high, float),
              batch_size_name: _ordered(batch_size_search_range_low, batch_size_search_range_high, int)}
    best_config, expected_best_step_num, _, _ = gp_update_hyperparameter_optimization(
        eval_fn=eval_fn, hyperparams=params, search_key_ranges=ranges, n=n, save_results_to=save_results_to, m=m,
        metric_should_increase=metric_should_increase, metric_name=metric_name, seed=seed)
    lr = best_config[lr_name] if lr_name in best_config else params[lr_name]
    return float(lr), int(expected_best_step_num)

"""Task / mini-batch samplers of the meta-learner (host side, integer work).

Mirror of /root/reference/meta_learners/metaseg.py:233-343.  Every function consumes Python's global
``random`` stream in exactly the reference's order (``random.sample(l, 1)`` for the task draw, one
``random.shuffle`` per epoch of ``_mini_batches``, one per train/test split), so with the same
``random.seed`` the sampled indices are bit-identical to the reference's.  The functions are generic over
what a "sample" is: the Session-compatible path passes ``[image, mask]`` pairs exactly like the reference,
the device fast path passes integer row indices into the task's example pool.
"""
from __future__ import annotations

import random
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np

from .synthetic import SyntheticSegmentationTask, make_synthetic_dataset

DEFAULT_NUM_TEST_EXAMPLES = 5    # metaseg.py:20


def _sample_mini_image_segmentation_dataset(sess, dataset, num_classes, num_shots, return_task_name: bool = False):
    """metaseg.py:233-255.  Samples one task and returns its first `num_shots` records (file order)."""
    l = list(dataset)
    class_obj = random.sample(l, 1)[0]
    if num_shots > class_obj.batch_size:
        warnings.warn("Requested {} examples but dataset can return max of {} examples.".format(
            num_shots, class_obj.batch_size))
        num_shots = class_obj.batch_size
    if not return_task_name:
        return class_obj.sample(sess, num_shots)
    return class_obj.sample(sess, num_shots), class_obj.name


def _sample_task_indices(dataset, num_shots) -> Tuple[object, List[int]]:
    """Index-space twin of `_sample_mini_image_segmentation_dataset`: same `random` consumption, returns the
    task object and the row indices [0..n) of the records `task.sample` would have returned."""
    l = list(dataset)
    class_obj = random.sample(l, 1)[0]
    if num_shots > class_obj.batch_size:
        warnings.warn("Requested {} examples but dataset can return max of {} examples.".format(
            num_shots, class_obj.batch_size))
        num_shots = class_obj.batch_size
    return class_obj, list(range(num_shots))


def _mini_batches(samples, batch_size, num_batches, replacement: bool = False, augmenter=None,
                  aug_rate: Optional[float] = None):
    """metaseg.py:258-302.  Without replacement every example is visited once per (reshuffled) epoch and a
    batch may straddle epochs, so with 5 shots and batch 8 a batch contains repeats."""
    if aug_rate is not None:
        prob_to_return_original = 1.0 - aug_rate
    else:
        prob_to_return_original = None
    samples = list(samples)
    if len(samples) == 0:
        raise ValueError('No samples to sample. `samples` has no length: {}'.format(samples))
    if replacement:
        for _ in range(num_batches):
            cur_batch = random.sample(samples, batch_size)
            if augmenter is not None:
                cur_batch = [augmenter.apply_augmentations(s[0], s[1], prob_to_return_original) for s in cur_batch]
            yield cur_batch
        return
    cur_batch = []
    batch_count = 0
    while True:
        random.shuffle(samples)
        for sample in samples:
            if augmenter is not None:
                sample = augmenter.apply_augmentations(sample[0], sample[1], prob_to_return_original)
            cur_batch.append(sample)
            if len(cur_batch) < batch_size:
                continue
            yield cur_batch
            cur_batch = []
            batch_count += 1
            if batch_count == num_batches:
                return


def _sample_train_test_segmentation_with_replacement(samples: List, train_shots: int = 5, test_shots: int = 5):
    """metaseg.py:313-318 (numpy's global, unseeded stream - as in the reference)."""
    indices = np.random.randint(len(samples), size=train_shots)
    train_set = [samples[x] for x in indices]
    indices = np.random.randint(len(samples), size=test_shots)
    test_set = [samples[x] for x in indices]
    return train_set, test_set


def _split_train_test_segmentation(samples, test_shots=1, test_train_test_split: bool = False,
                                   shuffle_before_split: bool = True):
    """metaseg.py:321-343."""
    samples = list(samples)[:]
    if shuffle_before_split:
        random.shuffle(samples)
    train_set = samples[:-test_shots]
    test_set = samples[-test_shots:]
    return train_set, test_set


def read_synthetic_dataset(num_train_tasks: int = 760, num_test_tasks: int = 240, n_examples: int = 10,
                           image_size: int = 224):
    """Stand-in for read_fss_1000_dataset (metaseg.py:24-121) when no tfrecord shards are available: the same
    return signature, synthetic FSS-1000-shaped tasks.  Test tasks are ids 0..239."""
    test = make_synthetic_dataset(num_test_tasks, n_examples, image_size, first_id=0)
    train = make_synthetic_dataset(num_train_tasks, n_examples, image_size, first_id=10000)
    return train, [], test, [t.name for t in train], [], [t.name for t in test]

"""FSS-1000 / FP-k-shot task readers over gzip-TFRecord shards, without TensorFlow.

Host-side mirror of
  * ``read_fss_1000_dataset``      meta_learners/metaseg.py:24-121
  * ``read_fp_k_shot_dataset``     meta_learners/metaseg.py:124-178
  * ``BinarySegmentationTask``     meta_learners/metaseg.py:181-230
  * ``get_fss_tasks`` / ``split_train_test_tasks``   data/fss_1000_utils.py:7-25

One shard per semantic class, named ``<class>.tfrecord.gzip`` (data/fss_1000_image_to_tfrecord.py:80).  ``sample(n)``
returns the FIRST n records in file order (the reference pipeline has no example-level shuffle,
data/input_fn.py:112-115).  Task objects also expose ``arrays()`` so ``Gecko`` takes the device fast path.

The official FSS-1000 test split (the reference's data/fss_test_set.txt, 240 class names, a DATA file of the
FSS-1000 authors) is not vendored here: pass ``test_task_ids`` as a list or as the path of that text file, or put
``fss_test_set.txt`` next to the shards; with ``test_task_ids=None`` and no such file the split is random, exactly
like the reference's ``test_task_ids=None`` branch.
"""
from __future__ import annotations

import glob
import os
import random
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import tfrecord

DEFAULT_NUM_TEST_EXAMPLES = 5   # metaseg.py:20
DEFAULT_K_SHOT_SET = [{"airliner", "aeroplane"}, {"bus"}, {"motorbike"}, {"potted_plant", "potted plant"},
                      {"television", "tvmonitor"}]   # metaseg.py:21
_SUFFIX = ".tfrecord.gzip"


def get_fss_tasks(data_dir: str) -> List[str]:
    return glob.glob(os.path.join(data_dir, "*.tfrecord*"))


def assert_train_test_split(train, test) -> None:
    for i in test:
        assert i not in train, "train-test leakage"


def split_train_test_tasks(all_tasks: List[str], n_test: int, reproducbile_splits: bool = False):
    """data/fss_1000_utils.py:7-19 (same ``random`` consumption: one shuffle of the task list)."""
    if not isinstance(all_tasks, list):
        all_tasks = list(all_tasks)
    if reproducbile_splits:
        all_tasks = sorted(all_tasks)
    else:
        random.shuffle(all_tasks)
    test_set = []
    for _ in range(n_test):
        test_set.append(all_tasks.pop())
    assert_train_test_split(all_tasks, test_set)
    return all_tasks, test_set


def load_task_id_list(path: str) -> List[str]:
    with open(path, "r") as f:
        return [line.rstrip("\n") for line in f if line.strip()]


class BinarySegmentationTask:
    """Segmentation maps for binary segmentations; label dimensions are [n_row, n_col, 2] (one-hot)."""

    def __init__(self, tfrecord_paths: Union[str, Sequence[str]], iterator=None, batch_size: int = 32, seed=None,
                 name: str = None, image_size: Optional[int] = None, verbose: bool = False):
        self.tfrecord_paths = tfrecord_paths
        self.batch_size = batch_size
        self.name = name
        self.image_size = 512 if image_size is None else image_size   # input_fn._IMAGE_WIDTH default
        self.iterator = iterator      # kept for signature compatibility; there is no TF iterator to share
        self._images = None
        self._masks = None
        if verbose:
            print("BinarySegmentationTask for data {} will return batches of size {}".format(self.name,
                                                                                             self.batch_size))

    def _materialise(self) -> None:
        if self._images is None:
            self._images, self._masks = tfrecord.load_examples(self.tfrecord_paths, self.image_size,
                                                               limit=self.batch_size)

    def arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        """(images f32 [n,S,S,3] in 0..255, masks f32 [n,S,S,2]) — the first ``batch_size`` records in file order."""
        self._materialise()
        return self._images, self._masks

    def sample(self, sess, num_images, verbose=False) -> List[List[np.ndarray]]:
        if num_images > self.batch_size:
            raise ValueError("Tried to sample {} examples.Cannot sample more than {} examples that generator was "
                             "initialized with.".format(num_images, self.batch_size))
        self._materialise()
        return [[image, mask] for image, mask in zip(self._images[:num_images], self._masks[:num_images])]

    def release(self) -> None:
        """Drops the decoded arrays (a 10-example 224x224 task holds 8 MB)."""
        self._images = self._masks = None


def _task_id(path: str) -> str:
    return os.path.basename(path).replace(_SUFFIX, "")


def _build(shards: List[str], image_size: Optional[int]):
    tasks, names = [], []
    for shard in shards:
        name = os.path.basename(shard)
        names.append(name)
        n = tfrecord.count_examples_in_tfrecords([shard])
        tasks.append(BinarySegmentationTask(tfrecord_paths=shard, batch_size=n, name=name, image_size=image_size))
    return tasks, names


def read_fss_1000_dataset(data_dir: str, num_val_tasks: int = 0, num_test_tasks: int = 240,
                          test_task_ids: Union[None, str, Sequence[str]] = "auto", image_size: Optional[int] = 224):
    """Returns (train_tasks, val_tasks, test_tasks, train_task_names, val_task_names, test_task_names)."""
    all_tasks = get_fss_tasks(data_dir)
    if isinstance(test_task_ids, str):
        if test_task_ids == "auto":
            # the reference holds out the OFFICIAL 240-class FSS-1000 test split (data/fss_1000_utils.py:31-37, :58;
            # metaseg.py:24-58).  A silent random split would score a checkpoint on classes it was meta-trained on,
            # so a missing list is an error; pass test_task_ids=None to ask for a random split explicitly.
            cand = os.path.join(data_dir, "fss_test_set.txt")
            if not os.path.exists(cand):
                raise FileNotFoundError(
                    "%s not found: copy the reference's data/fss_test_set.txt (the official FSS-1000 test split) next "
                    "to the shards, or pass test_task_ids=None for a random split of %d tasks" % (cand, num_test_tasks))
            test_task_ids = load_task_id_list(cand)
        else:
            test_task_ids = load_task_id_list(test_task_ids)
    if test_task_ids is None:
        train_shards, test_shards = split_train_test_tasks(all_tasks, num_test_tasks)
    else:
        ids = set(test_task_ids)
        train_shards, test_shards = [], []
        for task in all_tasks:
            (test_shards if _task_id(task) in ids else train_shards).append(task)
        assert all(_task_id(x) in ids for x in test_shards), "Test shard not in test_task_ids"
        assert all(_task_id(x) not in ids for x in train_shards), "Test set task found in train shards"
    train_shards, val_shards = split_train_test_tasks(train_shards, num_val_tasks, reproducbile_splits=True)
    print("{} training tasks, {} val tasks, {} test tasks.".format(len(train_shards), len(val_shards),
                                                                   len(test_shards)))
    train_tasks, train_names = _build(train_shards, image_size)
    val_tasks, val_names = _build(val_shards, image_size)
    test_tasks, test_names = _build(test_shards, image_size)
    return train_tasks, val_tasks, test_tasks, train_names, val_names, test_names


def read_fp_k_shot_dataset(data_dir: str, all_task_names=DEFAULT_K_SHOT_SET, image_size: Optional[int] = 224):
    """Each task = the union of the shards whose file name contains one of the synonyms (metaseg.py:124-178)."""
    all_tasks = get_fss_tasks(data_dir)
    print("{} tasks found.".format(len(all_tasks)))
    test_tasks, test_task_names = [], []
    for synonyms in all_task_names:
        task_shards, task_globs, task_name = [], [], None
        for i, synonym in enumerate(synonyms):
            synonym = synonym.replace(" ", "")
            if i == 0:
                task_name = synonym
            task_shards.extend(x for x in all_tasks if synonym in os.path.basename(x))
            task_globs.append(os.path.join(data_dir, "{}*.tfrecord*".format(synonym)))
        test_task_names.append(task_name)
        n = tfrecord.count_examples_in_tfrecords(task_shards)
        test_tasks.append(BinarySegmentationTask(tfrecord_paths=task_globs, batch_size=n, name=task_name,
                                                 image_size=image_size))
    return test_tasks, test_task_names


def write_task_shard(path: str, images_u8: np.ndarray, masks_u8: np.ndarray) -> int:
    """Writes one ``<class>.tfrecord.gzip`` shard in the reference writer's record schema."""
    return tfrecord.write_tfrecords(path, (tfrecord.make_example(i, m) for i, m in zip(images_u8, masks_u8)))

"""Synthetic FSS-1000-shaped few-shot segmentation tasks.

Each task object honours the contract of ``BinarySegmentationTask`` (meta_learners/metaseg.py:181-230):
``.name``, ``.batch_size`` and ``.sample(sess, n) -> [[image f32 [S,S,3] in 0..255, mask f32 [S,S,2]], ...]``
returning the FIRST n records in file order.  Records follow the tfrecord schema written by
data/fss_1000_image_to_tfrecord.py:99-134 and parsed by data/input_fn.py:45-65:
image uint8 [S,S,3]; mask uint8 [S,S] in {0,255} -> stack([255-m, m]) / 255.
"""
from __future__ import annotations

from typing import List

import numpy as np


def make_task_arrays(task_id: int, n_examples: int = 10, size: int = 224):
    """Returns (images uint8 [n,S,S,3], masks uint8 [n,S,S] in {0,255}); deterministic in task_id."""
    rng = np.random.default_rng(1000 + task_id)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    fg = rng.uniform(30, 225, size=3).astype(np.float32)
    bg = rng.uniform(30, 225, size=3).astype(np.float32)
    imgs = np.empty((n_examples, size, size, 3), np.uint8)
    masks = np.empty((n_examples, size, size), np.uint8)
    for e in range(n_examples):
        while True:
            m = np.zeros((size, size), bool)
            for _ in range(int(rng.integers(1, 4))):
                cy, cx = rng.uniform(0.3 * size, 0.7 * size, size=2)
                ry, rx = rng.uniform(0.08 * size, 0.35 * size, size=2)
                if rng.random() < 0.5:
                    m |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
                else:
                    m |= (np.abs(yy - cy) <= ry) & (np.abs(xx - cx) <= rx)
            frac = m.mean()
            if 0.05 <= frac <= 0.60:
                break
        # low-frequency illumination field + per-pixel noise
        fy, fx = rng.uniform(0.5, 2.0, size=2)
        ph = rng.uniform(0, 2 * np.pi, size=2)
        low = 18.0 * np.sin(2 * np.pi * fy * yy / size + ph[0]) * np.cos(2 * np.pi * fx * xx / size + ph[1])
        img = np.where(m[..., None], fg, bg).astype(np.float32) + low[..., None]
        img += rng.normal(0.0, 20.0, size=img.shape).astype(np.float32)
        imgs[e] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        masks[e] = m.astype(np.uint8) * 255
    return imgs, masks


def parse_records(images_u8: np.ndarray, masks_u8: np.ndarray):
    """data/input_fn.py:45-65 parse_example on already-decoded bytes."""
    image = images_u8.astype(np.float32)
    m = masks_u8.astype(np.float32)
    mask = np.stack([255.0 - m, m], axis=-1) / np.float32(255.0)
    return image, mask.astype(np.float32)


class SyntheticSegmentationTask:
    """Drop-in for BinarySegmentationTask backed by in-memory synthetic records."""

    def __init__(self, task_id: int, n_examples: int = 10, image_size: int = 224, name: str = None):
        self.task_id = task_id
        self.batch_size = n_examples
        self.image_size = image_size
        self.name = name or "synthetic_%04d" % task_id
        self._images = None
        self._masks = None

    def _materialise(self):
        if self._images is None:
            iu8, mu8 = make_task_arrays(self.task_id, self.batch_size, self.image_size)
            self._images, self._masks = parse_records(iu8, mu8)

    def arrays(self):
        """(images f32 [n,S,S,3], masks f32 [n,S,S,2]) in file order."""
        self._materialise()
        return self._images, self._masks

    def sample(self, sess, num_images, verbose=False) -> List[List[np.ndarray]]:
        if num_images > self.batch_size:
            raise ValueError("Tried to sample {} examples.Cannot sample more than {} examples that generator was "
                             "initialized with.".format(num_images, self.batch_size))
        self._materialise()
        return [[image, mask] for image, mask in zip(self._images[:num_images], self._masks[:num_images])]


def make_synthetic_dataset(n_tasks: int = 240, n_examples: int = 10, image_size: int = 224, first_id: int = 0):
    return [SyntheticSegmentationTask(first_id + t, n_examples, image_size) for t in range(n_tasks)]

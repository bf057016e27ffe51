#!/usr/bin/env python
"""Writes the FSS-1000 image folders as one gzip-TFRecord shard per class, without TensorFlow.

Drop-in for /root/reference/data/fss_1000_image_to_tfrecord.py (:23-44 flags, :46-59 folder layout <root>/<class>/
{N.jpg, N.png}, :137-160 writer): 224x224 pairs only (others are skipped), the mask's first channel is stored,
examples are shuffled within a task with Python's `random` (seed it for reproducible shards), record schema =
bytes features `image` / `mask` (mliis_b200/tfrecord.py).  Images are decoded with Pillow instead of imageio.

    python tools/fss1000_to_tfrecords.py --input_dir fewshot_data --tfrecord_dir fss1000_tfrecords [--overwrite 1]
"""
import argparse
import glob
import os
import random
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mliis_b200 import tfrecord  # noqa: E402

IMAGE_DIMS = 224


def parse_arguments(argv):
    parser = argparse.ArgumentParser(description="Writes FSS-1000 images to TFRecords.")
    parser.add_argument("--input_dir", type=str, default=None)
    parser.add_argument("--tfrecord_dir", required=True, type=str)
    parser.add_argument("--overwrite", required=False, default=False, type=bool)
    parser.add_argument("--image_dims", type=int, default=IMAGE_DIMS)
    args, _ = parser.parse_known_args(args=argv[1:])
    return args


def get_fss_dir_paths(data_dir):
    return glob.glob(os.path.join(data_dir, "*/"))


def get_image_mask_pairs(task, image_ext=".jpg", mask_ext=".png"):
    pairs = []
    for mask in glob.glob(os.path.join(task, "*" + mask_ext)):
        image = mask.replace(mask_ext, image_ext)
        if os.path.exists(image):
            pairs.append((image, mask))
        else:
            warnings.warn("No corresponding image found for mask: {}".format(mask))
    return pairs


def _read(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im)


def write_tfrecord(tfrecord_filename, filename_pairs, image_dims=IMAGE_DIMS):
    """Returns the number of examples written."""
    filename_pairs = list(filename_pairs)
    random.shuffle(filename_pairs)               # shuffle examples within a task (reference :151)

    def payloads():
        for image_filename, mask_filename in filename_pairs:
            im, mk = _read(image_filename), _read(mask_filename)
            if im.shape[:2] != (image_dims, image_dims) or mk.shape[:2] != (image_dims, image_dims):
                print("{} is not of expected image dimensions. Skipping this sample".format(image_filename))
                continue
            if im.ndim == 2:
                im = np.repeat(im[:, :, None], 3, axis=2)
            yield tfrecord.make_example(im[:, :, :3], mk)

    return tfrecord.write_tfrecords(tfrecord_filename, payloads())


def main(argv=None):
    start = time.time()
    args = parse_arguments(sys.argv if argv is None else argv)
    task_dirs = get_fss_dir_paths(args.input_dir)
    print("{} tasks found".format(len(task_dirs)))
    os.makedirs(args.tfrecord_dir, exist_ok=True)
    for task in task_dirs:
        task_name = os.path.basename(task.rstrip("/"))
        out = os.path.join(args.tfrecord_dir, task_name + ".tfrecord.gzip")
        if os.path.exists(out) and not args.overwrite:
            continue
        n = write_tfrecord(out, get_image_mask_pairs(task), args.image_dims)
        print("Wrote {} examples of task {} to {}".format(n, task_name, out))
    print("Finished.")
    print("Took {} minutes.".format((time.time() - start) / 60.0))


if __name__ == "__main__":
    main()

"""Supervised joint training of EfficientLab on all FSS-1000 classes with SGD/Adam (SURVEY.md 8f row 1,
BASELINE config 5) on the B200 engine.

Mirror of /root/reference/joint_train.py (:32-82 flags, :86-109 model kwargs, :120-150 shards / dataset,
:153-245 train loop, :248-269 IoU callback, :295-343 main) and joint_train/data/input_fn.py.  Differences, all forced
by what the reference does with memory:

* The reference label tensor is dense one-hot [224,224,1001] per image (50 MB uint8 in the shard, 200 MB fp32 on the
  device, again for logits and probabilities).  Here an example is (image uint8 [S,S,3], binary mask uint8 [S,S],
  class id) and the loss kernels (csrc/k_mc.cu) never materialise anything of size H*W*C.  Both shard layouts are
  read: the per-class few-shot shards ``<class>.tfrecord.gzip`` (class id = 1 + rank of the class name in the sorted
  class list, the writer's convention, data/fss_1000_image_to_joint_tfrecord_shards.py:140-158) and the reference's
  dense joint shards (converted on read: the foreground channel is the non-background channel with the largest sum).
* ``get_model_kwargs`` reads ``args.lsd``, which the parser never defines (joint_train.py:92-93): the script cannot
  start as shipped.  ``--rsd`` is used here.
* Data parallelism (the reference is single-GPU): under torchrun every rank takes batch_size / world
  examples of each global batch, BatchNorm statistics stay local (the reference's own <= 8-shard rule,
  models/efficientnet/utils.py:116-118) and the flat gradient (2.18 M floats) is averaged with ONE NCCL all-reduce per
  step before the optimizer.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import tfrecord
from .np_augmenters import Augmenter, additive_gaussian_noise, exposure, fliplr, translate

TRAIN_ID, VAL_ID, TEST_ID = "train", "val", "test"
SUPPORTED_MODELS = {"efficientlab"}


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Train segmentation model via SGD.")
    parser.add_argument("--data_dir", help="Path to folder containing tfrecords", required=True)
    parser.add_argument("--fp_k_test_set", action="store_true")
    parser.add_argument("--test_on_val_set", action="store_true")
    parser.add_argument("--model_name", default="EfficientLab")
    parser.add_argument("--rsd", type=int, nargs="+")
    parser.add_argument("--feature_extractor_name", type=str, default="efficientnet-b0")
    parser.add_argument("--image_size", type=int, default=224)
    parser.add_argument("--seperate_background_channel", action="store_true", default=False)
    parser.add_argument("--restore_efficient_net_weights_from", type=str, default=None)
    parser.add_argument("--sgd", action="store_true")
    parser.add_argument("--loss_name", default="ce_dice")
    parser.add_argument("--l2", action="store_true")
    parser.add_argument("--augment", action="store_true")
    parser.add_argument("--final_layer_dropout_rate", type=float, default=0.0)
    parser.add_argument("--batch_size", default=64, type=int)
    parser.add_argument("--epochs", default=200, type=int)
    parser.add_argument("--steps_per_epoch", type=int, default=None)
    parser.add_argument("--learning_rate", default=0.005, type=float)
    parser.add_argument("--final_learning_rate", default=5e-7, type=float)
    parser.add_argument("--label_smoothing", default=0.0, type=float)
    parser.add_argument("--val_batches", default=20, type=int)
    parser.add_argument("--pretrained", action="store_true", default=False)
    parser.add_argument("--eval_interval", default=2, type=int)
    parser.add_argument("--seed", default=0, type=int)
    parser.add_argument("--checkpoint", default="/tmp/model_checkpoint", type=str)
    # engine knobs (not in the reference)
    parser.add_argument("--gemm_mode", default="tf32x3", choices=["fp32", "tf32", "tf32x3"])
    parser.add_argument("--class_list", default=None,
                        help="text file with one class name per line (default: the shard names, sorted)")
    return parser.parse_args(argv)


def get_model_kwargs(parsed_args) -> dict:
    parsed_args.model_name = parsed_args.model_name.lower()
    if parsed_args.model_name not in SUPPORTED_MODELS:
        raise ValueError("Model name must be in the set: {}".format(SUPPORTED_MODELS))
    res = {"learning_rate": parsed_args.learning_rate,
           "restore_ckpt_dir": parsed_args.restore_efficient_net_weights_from,
           "feature_extractor_name": parsed_args.feature_extractor_name, "l2": parsed_args.l2,
           "final_layer_dropout_rate": parsed_args.final_layer_dropout_rate,
           "label_smoothing": parsed_args.label_smoothing,
           "optimizer": "sgd" if parsed_args.sgd else "adam", "loss_name": parsed_args.loss_name,
           "n_rows": parsed_args.image_size, "n_cols": parsed_args.image_size}
    if parsed_args.rsd:
        res["rsd"] = parsed_args.rsd
    if "dice" not in parsed_args.loss_name:
        res["dice"] = False
    return res


def get_train_test_shards_from_dir(data_dir, ext: str = ".tfrecord.gzip", test_on_val_set: bool = False,
                                   test_ids: Optional[Sequence[str]] = None):
    """joint_train.py:120-135: dense joint shards are told apart by 'train' / 'val' / 'test' in their file names.
    Per-class few-shot shards (`<class>.tfrecord.gzip`, what the meta-learning path reads) carry no such marker - class
    names like 'train' or 'contest' would even land in both lists - so that layout is detected and split by the
    test-id list instead (every shard trains when no list is given; IoU is then measured on training batches, like the
    reference's iou_callback does)."""
    all_shards = sorted(x for x in os.listdir(data_dir) if ext in x)
    dense = [x for x in all_shards if x.startswith((TRAIN_ID, TEST_ID, VAL_ID))]
    if len(dense) != len(all_shards):
        ids = set(test_ids or ())
        base = lambda x: x[:x.index(ext)]
        train_shards = [x for x in all_shards if base(x) not in ids]
        test_shards = [x for x in all_shards if base(x) in ids]
        return [os.path.join(data_dir, x) for x in train_shards], [os.path.join(data_dir, x) for x in test_shards]
    train_shards = [x for x in all_shards if TEST_ID not in x]
    test_shards = [x for x in all_shards if TRAIN_ID not in x]
    if test_on_val_set:
        train_shards = [x for x in train_shards if VAL_ID not in x]
        test_shards = [x for x in all_shards if VAL_ID in x]
        assert len(set(train_shards + test_shards)) == len(all_shards) - len([x for x in all_shards if TEST_ID in x])
    else:
        assert len(set(train_shards + test_shards)) == len(all_shards)
    assert len(set(test_shards).intersection(set(train_shards))) == 0
    return [os.path.join(data_dir, x) for x in train_shards], [os.path.join(data_dir, x) for x in test_shards]


# ---------------------------------------------------------------------------------------------------
# sparse examples
# ---------------------------------------------------------------------------------------------------
class SparseSegmentationData:
    """images uint8 [n,S,S,3], masks uint8 [n,S,S] in {0,255}, class_ids int32 [n] (1..n_classes)."""

    def __init__(self, images: np.ndarray, masks: np.ndarray, class_ids: np.ndarray, n_classes: int):
        assert images.shape[0] == masks.shape[0] == class_ids.shape[0]
        assert class_ids.min(initial=1) >= 1 and class_ids.max(initial=1) <= n_classes
        self.images, self.masks, self.class_ids, self.n_classes = images, masks, class_ids.astype(np.int32), n_classes

    def __len__(self):
        return self.images.shape[0]

    def dense_labels(self, rows: Sequence[int]) -> np.ndarray:
        """The reference's label layout [n,S,S,n_classes+1] (tests / small problems only)."""
        rows = list(rows)
        s = self.images.shape[1]
        out = np.zeros((len(rows), s, s, self.n_classes + 1), np.float32)
        for k, r in enumerate(rows):
            m = self.masks[r].astype(np.float32) / np.float32(255.0)
            out[k, :, :, 0] = 1.0 - m
            out[k, :, :, self.class_ids[r]] = m
        return out


def _decode_record(payload: bytes, size: int, class_id: Optional[int], n_out: Optional[int]):
    f = tfrecord.decode_example(payload)
    img = np.frombuffer(f["image"][0], np.uint8).reshape(size, size, 3)
    raw = np.frombuffer(f["mask"][0], np.uint8)
    if raw.size == size * size:                       # per-class few-shot shard: binary mask, class from the file
        if class_id is None:
            raise tfrecord.TFRecordError("binary-mask record needs a class id (shard name not in the class list)")
        return img, raw.reshape(size, size), class_id
    if raw.size % (size * size):
        raise tfrecord.TFRecordError("mask feature of %d bytes is not a multiple of %dx%d" % (raw.size, size, size))
    dense = raw.reshape(size, size, raw.size // (size * size))      # the reference's joint shard
    if n_out is not None and dense.shape[2] != n_out:
        raise tfrecord.TFRecordError("dense mask has %d channels, model has %d" % (dense.shape[2], n_out))
    sums = dense[:, :, 1:].reshape(-1, dense.shape[2] - 1).sum(0, dtype=np.int64)
    c = int(np.argmax(sums)) + 1
    return img, dense[:, :, c], c


def load_sparse_shards(paths: Sequence[str], image_size: int, class_names: Optional[Sequence[str]] = None,
                       n_classes: Optional[int] = None, limit: Optional[int] = None) -> SparseSegmentationData:
    """Reads few-shot per-class shards and / or dense joint shards into the sparse representation."""
    index = {name: i + 1 for i, name in enumerate(class_names)} if class_names is not None else {}
    if n_classes is None:
        n_classes = len(class_names) if class_names is not None else None
    images, masks, ids = [], [], []
    for path in paths:
        base = os.path.basename(path)
        for suffix in (".tfrecord.gzip", ".tfrecord"):
            if base.endswith(suffix):
                base = base[:-len(suffix)]
        cid = index.get(base)
        for payload in tfrecord.read_tfrecords(path):
            if limit is not None and len(images) >= limit:
                break
            im, mk, c = _decode_record(payload, image_size, cid, None if n_classes is None else n_classes + 1)
            images.append(im)
            masks.append(mk)
            ids.append(c)
    if not images:
        raise ValueError("no examples found in {}".format(list(paths)))
    ids = np.asarray(ids, np.int32)
    return SparseSegmentationData(np.stack(images), np.stack(masks), ids,
                                  int(n_classes if n_classes is not None else ids.max()))


class SparseBatcher:
    """Endless shuffled batches (the reference: shuffle files, interleave, shuffle buffer 400, batch - unseeded; here
    a seeded permutation per epoch).  Augmentation runs on a 2-channel [background, foreground] mask, which is what
    the reference's 1001-channel augmentation does to the only two channels that are not identically zero."""

    def __init__(self, data: SparseSegmentationData, batch_size: int, seed: int = 0,
                 augmenter: Optional[Augmenter] = None, rank: int = 0, world: int = 1):
        if batch_size % world:
            raise ValueError("batch_size %d is not divisible by the number of ranks %d" % (batch_size, world))
        self.data, self.batch_size, self.augmenter = data, batch_size, augmenter
        self.rank, self.world = rank, world
        self.rng = np.random.default_rng(seed)      # same stream on every rank: ranks slice the same global batch
        self._order = np.zeros(0, np.int64)

    def next_rows(self) -> np.ndarray:
        while self._order.size < self.batch_size:
            self._order = np.concatenate([self._order, self.rng.permutation(len(self.data))])
        rows, self._order = self._order[:self.batch_size], self._order[self.batch_size:]
        per = self.batch_size // self.world
        return rows[self.rank * per:(self.rank + 1) * per]

    def next_batch(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """(images f32 [b,S,S,3] in 0..255, masks f32 [b,S,S] in {0,1}, class ids int32 [b]) of this rank."""
        rows = self.next_rows()
        d = self.data
        images = d.images[rows].astype(np.float32)
        masks = d.masks[rows].astype(np.float32) / np.float32(255.0)
        if self.augmenter is not None:
            for k in range(len(rows)):
                two = np.stack([1.0 - masks[k], masks[k]], axis=2)
                im, mk = self.augmenter.apply_augmentations(images[k], two, return_image_mask_in_list=False)
                images[k], masks[k] = np.asarray(im, np.float32), np.asarray(mk, np.float32)[:, :, 1]
        return images, masks, d.class_ids[rows]


def make_augmenter() -> Augmenter:
    """joint_train.py:141-147: translate (mask filled with background), fliplr, gaussian noise, exposure."""
    return Augmenter(aug_funcs=[translate, fliplr, additive_gaussian_noise, exposure])


# ---------------------------------------------------------------------------------------------------
# training
# ---------------------------------------------------------------------------------------------------
def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return None, 0, 1


def average_gradients(grads, dist=None, world: int = 1):
    """ONE all-reduce of the flat gradient per step (mean over ranks: every rank holds the mean over its own
    batch slice, and slices have equal size)."""
    if dist is not None and world > 1:
        dist.all_reduce(grads, op=dist.ReduceOp.SUM)
        grads.div_(world)
    return grads


def linear_lr(i: int, epochs: int, initial_lr: float, final_lr: float) -> float:
    """joint_train.py:332-335."""
    frac_done = i / epochs
    return frac_done * final_lr + (1 - frac_done) * initial_lr


class JointTrainer:
    """One training / evaluation step of the joint-training model on slot 0 of the engine."""

    def __init__(self, model):
        import torch
        self.torch = torch
        self.model = model
        self.eng = model.engine()
        self.dist, self.rank, self.world = _dist()

    def _upload(self, images, masks, class_ids):
        t = self.torch
        dev = self.eng.device
        x = t.from_numpy(np.ascontiguousarray(images, np.float32)).to(dev, non_blocking=True)
        m = t.from_numpy(np.ascontiguousarray(masks, np.float32)).to(dev, non_blocking=True)
        c = t.from_numpy(np.ascontiguousarray(class_ids, np.int32)).to(dev, non_blocking=True)
        return x, m, c

    def train_step(self, images, masks, class_ids, lr: float, seed: int = 0) -> float:
        x, m, c = self._upload(images, masks, class_ids)
        eng, B = self.eng, int(x.shape[0])
        eng.set_class_ids(0, c)
        if self.world == 1:
            loss = self.torch.zeros(1, device=eng.device)
            eng.train_step(0, x, m, lr, batch=B, seed=seed, loss_out=loss)
        else:
            eng.forward(0, x, True, batch=B, seed=seed, want_logits=False)
            loss, grads = eng.loss_backward(0, m, B)
            eng.set_grads(0, average_gradients(grads, self.dist, self.world))
            eng.optimizer_step(0, lr)
        self._keep = (x, m, c)          # inputs stay referenced until the next step has been enqueued
        return float(loss.item())

    def evaluate_batch(self, images, masks, class_ids) -> float:
        """compute_iou_metric (joint_train.py:262-269): mean over the batch of |pred AND label| / |pred OR label|
        over all channels, predictions thresholded at 0.5."""
        x, m, c = self._upload(images, masks, class_ids)
        self.eng.set_class_ids(0, c)
        _, inter, uni = self.eng.predict_classes(0, x, m, want_class_map=False)
        i, u = inter.cpu().numpy().astype(np.float64), uni.cpu().numpy().astype(np.float64)
        self._keep = (x, m, c)
        return float(np.nanmean((i + 1e-7) / (u + 1e-7)))


def iou_callback(trainer: JointTrainer, batcher: SparseBatcher, val_batches: int) -> float:
    return float(np.nanmean([trainer.evaluate_batch(*batcher.next_batch()) for _ in range(val_batches)]))


def train(sess, model, batcher: SparseBatcher, epochs: int, steps_per_epoch: int, save_dir: str, lr_fn: Callable,
          restore_ckpt_dir: Optional[str] = None, val_batches: int = 20, save_checkpoint_every_n_epochs: int = 2,
          time_deadline=None, max_checkpoints_to_keep: int = 2, eval_interval: int = 2,
          val_batcher: Optional[SparseBatcher] = None) -> List[float]:
    """joint_train.py:153-245.  Returns the IoU history."""
    from .checkpoint import Saver
    assert isinstance(epochs, int) and isinstance(steps_per_epoch, int)
    trainer = JointTrainer(model)
    is_root = trainer.rank == 0
    if is_root:
        os.makedirs(save_dir, exist_ok=True)
        print("Logging to {}".format(save_dir))
    saver = Saver(model, max_to_keep=max_checkpoints_to_keep)
    if restore_ckpt_dir is not None:
        print("Restoring from checkpoint {}".format(restore_ckpt_dir))
        model.restore_model(sess, restore_ckpt_dir, filter_to_scopes=[model.feature_extractor_name])
    if not model.variables_initialized:
        print("Initializing variables.")
        model.initialize()
    print("Training...")
    if is_root:
        saver.save(sess, os.path.join(save_dir, "model.ckpt"), global_step=0)
    ious, step = [], 0
    for i in range(epochs):
        start_time = time.time()
        lr = lr_fn(i)
        if is_root:
            print("Epoch: ", i)
            print("lr: ", lr)
        loss = float("nan")
        for _ in range(steps_per_epoch):
            step += 1
            loss = trainer.train_step(*batcher.next_batch(), lr=lr, seed=step)
        elapsed = max(time.time() - start_time, 1e-9)
        if is_root:
            print("Finished epoch {} with {} steps.".format(i, steps_per_epoch))
            print("Iterations per second: {}".format(steps_per_epoch / elapsed))
        if i % eval_interval == 0:
            iou = iou_callback(trainer, val_batcher or batcher, val_batches)
            if is_root:
                print("Validating")
                print("Loss: {}".format(loss))
                print("IoU on epoch {} estimated on {} batches:".format(i, val_batches))
                print(iou)
            ious.append(iou)
        if is_root and (i % save_checkpoint_every_n_epochs == 0 or i == epochs - 1):
            print("Saving checkpoint to {}.".format(save_dir))
            saver.save(sess, os.path.join(save_dir, "model.ckpt"), global_step=i)
        if time_deadline is not None and time.time() > time_deadline:
            break
    if is_root:
        print("Training complete. History:")
        print("Train set Intersection over Union (IoU):")
        print(ious)
    return ious


def main(argv=None):
    from .efficientlab import EfficientLab
    from .session import Session
    start = time.time()
    args = parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rank = int(os.environ.get("RANK", "0"))
    test_ids_file = os.path.join(args.data_dir, "fss_test_set.txt")
    test_ids = [l.strip() for l in open(test_ids_file) if l.strip()] if os.path.exists(test_ids_file) else None
    train_shards, test_shards = get_train_test_shards_from_dir(args.data_dir, test_on_val_set=args.test_on_val_set,
                                                               test_ids=test_ids)
    per_class_layout = not any(os.path.basename(p).startswith((TRAIN_ID, TEST_ID, VAL_ID))
                               for p in train_shards + test_shards)
    if args.class_list:
        class_names = [l.strip() for l in open(args.class_list) if l.strip()]
    elif per_class_layout:
        # one shard per class: the class list is the set of shard names (the reference: TRAIN_TASK_IDS + TEST_TASK_IDS)
        class_names = sorted({os.path.basename(p).replace(".tfrecord.gzip", "") for p in train_shards + test_shards})
    else:
        raise ValueError("dense joint shards (train_*/test_*) carry a [H,W,n_classes+1] mask: pass --class_list with "
                         "the class names in channel order (the reference uses TRAIN_TASK_IDS + TEST_TASK_IDS)")
    num_classes = len(class_names)
    data = load_sparse_shards(train_shards, args.image_size, class_names, num_classes)
    augmenter = make_augmenter() if args.augment else None
    batcher = SparseBatcher(data, args.batch_size, seed=args.seed, augmenter=augmenter, rank=rank, world=world)
    mk = get_model_kwargs(args)
    restore_ckpt_dir = mk.pop("restore_ckpt_dir")
    mk.pop("loss_name")
    model = EfficientLab(n_classes=num_classes, seperate_background_channel=True, binary_iou_loss=False,
                         gemm_mode=args.gemm_mode, task_slots=1, max_batch=args.batch_size // world, **mk)
    steps_per_epoch = int(760 * 10 // args.batch_size) if args.steps_per_epoch is None else args.steps_per_epoch

    def lr_fn(i):
        return linear_lr(i, args.epochs, args.learning_rate, args.final_learning_rate)

    with Session(model) as sess:
        train(sess, model, batcher, args.epochs, steps_per_epoch=steps_per_epoch, save_dir=args.checkpoint,
              lr_fn=lr_fn, val_batches=args.val_batches, eval_interval=args.eval_interval,
              restore_ckpt_dir=restore_ckpt_dir)
    if rank == 0:
        print("Finished training")
        print("Experiment took {} hours".format((time.time() - start) / 3600.))


if __name__ == "__main__":
    main()

"""Host-side numpy augmentations for (image, one-hot mask) pairs.

Behavioural mirror of /root/reference/augmenters/np_augmenters.py (images float in [0, 255], masks [H, W, 2] with
channel 0 = background): the same six transforms, the same defaults and — so that a seeded run draws the same
augmentations — the same order of draws from the two global streams the reference uses (``np.random`` and Python's
``random``).  Reference quirks that change results are kept and marked QUIRK.

  additive_gaussian_noise  :9-12      exposure      :15-18     random_eraser   :21-36
  fliplr                   :39-42     translate     :85-97     rotate_img_mask :100-129
  Augmenter.apply_augmentations :135-160
"""
from __future__ import annotations

import random
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
from scipy.ndimage import rotate as _nd_rotate

_F32 = np.float32
_BACKGROUND = [1, 0]   # one-hot "background" pixel


def _as_f32(image, mask):
    return image.astype(_F32), mask.astype(_F32)


def additive_gaussian_noise(image, mask, mean_sd=5.1):
    """Per-pixel N(0, sd) noise, sd ~ |N(mean_sd, 1)|, clipped to [0, 255]."""
    sd = np.abs(np.random.normal(mean_sd, 1, 1))
    return _as_f32(np.clip(image + np.random.normal(0, sd, image.shape), 0., 255.), mask)


def exposure(image, mask, mean_sd=12.75):
    """One global brightness offset ~ N(0, sd), sd ~ |N(mean_sd, 1)|."""
    sd = np.abs(np.random.normal(mean_sd, 1, 1))
    return _as_f32(np.clip(image + np.random.normal(0, sd, 1), 0., 255.), mask)


def random_eraser(input_img, mask, s_l=0.02, s_h=0.10, r_1=0.3, r_2=1 / 0.3, v_l=0, v_h=255):
    """Random erasing (arXiv:1708.04896) of one rectangle; the erased area becomes background in the mask.
    Operates in place on its arguments, like the reference (the Augmenter hands it copies)."""
    img_h, img_w, _ = input_img.shape
    area = np.random.uniform(s_l, s_h) * img_h * img_w
    aspect = np.random.uniform(r_1, r_2)
    w, h = int(np.sqrt(area / aspect)), int(np.sqrt(area * aspect))
    top, left = np.random.randint(0, img_h), np.random.randint(0, img_w)
    value = np.random.uniform(v_l, v_h)
    input_img[top:top + h, left:left + w, :] = value
    mask[top:top + h, left:left + w, :] = _BACKGROUND
    return _as_f32(input_img, mask)


def fliplr(image, mask):
    return _as_f32(np.fliplr(image), np.fliplr(mask))


def _roll_and_fill(image, shift, roll, positive, roll_axis, fill):
    """np.roll by +-shift along ``roll_axis``; unless ``roll``, overwrite a band of width ``shift`` with ``fill``
    (random colour per channel when fill is None).
    QUIRK (kept): the band is taken along the OTHER axis than the one that was rolled (shift_img_lr rolls axis 0
    and fills columns; shift_img_ud rolls axis 1 and fills rows), reference :45-82."""
    image = np.roll(image, shift if positive else -shift, roll_axis)
    if roll:
        return image
    colour = fill if fill is not None else np.random.uniform(0, 255, image.shape[2])
    if roll_axis == 0:       # "lr"
        if positive:
            image[:, :shift] = colour
        else:
            image[:, -shift:] = colour
    else:                    # "ud"
        if positive:
            image[-shift:, :] = colour
        else:
            image[:shift, :] = colour
    return image


def shift_img_lr(image, shift, roll, right, fill: Optional[Union[int, List[int]]] = None):
    return _roll_and_fill(image, shift, roll, right, 0, fill)


def shift_img_ud(image, shift, roll, up, fill: Optional[Union[int, List[int]]] = None):
    return _roll_and_fill(image, shift, roll, up, 1, fill)


def translate(image, mask, max_shift=23, mask_fill=_BACKGROUND):
    """Random jitter of up to max_shift pixels, wrapped (roll) or filled."""
    vert = random.getrandbits(1)
    direction = random.getrandbits(1)
    shift = np.random.randint(1, max_shift + 1, 1)[0]
    roll = random.getrandbits(1)
    mover = shift_img_ud if vert else shift_img_lr
    image = mover(image, shift, roll, direction)
    mask = mover(mask, shift, roll, direction, fill=mask_fill)
    return _as_f32(image, mask)


def rotate_img_mask(image, mask, max_angle: int = 45, mask_fill=_BACKGROUND):
    """Rotation by a random integer angle with a random border mode; nearest-neighbour for the mask."""
    angle = np.random.randint(-max_angle, max_angle)
    mode = random.sample(['reflect', 'constant', 'mirror', 'wrap'], 1)[0]
    noise_border = False
    cval = 0
    if mode == "constant":
        if random.getrandbits(1):
            cval, noise_border = -256, True
        else:
            cval = np.random.randint(0, 256)
    image = _nd_rotate(image, angle=angle, reshape=False, mode=mode, cval=cval)
    if noise_border:
        outside = image == -256
        noise = np.random.randint(0, 256, size=image.shape)
        image[outside] = noise[outside]
    mask = _nd_rotate(mask, angle=angle, reshape=False, mode=mode, cval=-256, order=0)
    if mode == "constant":
        mask[mask[:, :, 0] == -256] = mask_fill
    return image, mask       # QUIRK (kept): no float32 cast here, dtype follows the input


cur_aug_funcs: List[Callable] = [random_eraser, translate, fliplr, additive_gaussian_noise, exposure, rotate_img_mask]


class Augmenter:
    """Image segmentation augmenter: with probability ``prob_to_return_original`` the pair is returned untouched,
    otherwise 1..len(aug_funcs) transforms are applied in a freshly shuffled order."""

    def __init__(self, aug_funcs: Optional[Sequence[Callable]] = None):
        if aug_funcs is None:
            aug_funcs = cur_aug_funcs     # QUIRK (kept): the module-level list itself is shuffled in place
        self.aug_funcs = aug_funcs
        self.prob_to_return_original = 1. / (len(aug_funcs) + 1)
        print("Initialized image segmentation augmenter.")

    def apply_augmentations(self, image, mask, prob_to_return_original=0.0, return_image_mask_in_list: bool = True):
        prob = prob_to_return_original if prob_to_return_original is not None else self.prob_to_return_original
        if np.random.rand() <= prob:
            return image, mask
        image, mask = image.copy(), mask.copy()
        random.shuffle(self.aug_funcs)
        num_to_apply = np.random.randint(1, len(self.aug_funcs) + 1)
        for fn in self.aug_funcs[:num_to_apply]:
            image, mask = fn(image, mask)
        return [image, mask] if return_image_mask_in_list else (image, mask)

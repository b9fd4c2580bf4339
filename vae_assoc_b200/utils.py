"""The data loaders the training script calls before `train()` (/root/reference/utils.py:142-195, called at
vae_assoc_ujichar_img_jnt.py:21-36): host-side unpickling of the UJI character set into the [N, 784 + 147] matrix the
hot path consumes.  Pure host code (no arithmetic of the train step); restated for Python 3 -- the reference's pickles
were written by Python 2 (`cPickle`), hence `encoding="latin1"`.

Differences that do not change results: one pre-sized output matrix is filled in place instead of a Python list of
per-sample arrays being re-packed by `np.array`; the iteration order (characters sorted by their last letter /
by key, samples in file order) is the reference's, because the image and trajectory rows are paired by position.
"""
import pickle

import numpy as np

from . import dataset


def _load(data, fname):
    if fname is not None:
        with open(fname, "rb") as f:
            return pickle.load(f, encoding="latin1")
    return data


def _keep(char, only_digits):
    # utils.py:155,172,189: `ord(char[-1]) > 57` drops everything after '9' when only_digits
    return not (only_digits and ord(char[-1]) > 57)


def extract_images(data=None, fname=None, only_digits=True, dtype=np.float32):
    """utils.py:142-158 -- every character image flattened and scaled by 1/255 (values in [0, 1]); characters ordered by
    the LAST letter of their key."""
    data_dict = _load(data, fname)
    if data_dict is None:
        return np.array([])
    chars = [c for c in sorted(data_dict.keys(), key=lambda k: k[-1]) if _keep(c, only_digits)]
    n = sum(len(data_dict[c]) for c in chars)
    if n == 0:
        return np.array([])
    width = int(np.asarray(data_dict[chars[0]][0]).size)
    out = np.empty((n, width), dtype=dtype)
    i = 0
    for c in chars:
        for d in data_dict[c]:
            out[i] = np.asarray(d).reshape(-1).astype(dtype) * (1. / 255.)
            i += 1
    return out


def extract_jnt_trajs(data=None, fname=None, only_digits=True, dtype=np.float32):
    """utils.py:160-176 -- flattened joint trajectories, characters in key order."""
    data_dict = _load(data, fname)
    rows = []
    if data_dict is not None:
        for c in sorted(data_dict.keys()):
            if _keep(c, only_digits):
                rows += [np.asarray(d).reshape(-1).astype(dtype) for d in data_dict[c]]
    return np.array(rows)


def extract_jnt_fa_parms(data=None, fname=None, only_digits=True, dtype=np.float32):
    """utils.py:178-195 -- function-approximator (RBF) parameters of the joint trajectories plus their per-column mean and
    standard deviation; the caller z-scores with them (vae_assoc_ujichar_img_jnt.py:29)."""
    data_dict = _load(data, fname)
    rows = []
    if data_dict is not None:
        for c in sorted(data_dict.keys()):
            if _keep(c, only_digits):
                rows += [d for d in data_dict[c]]
    fa_parms = np.array(rows)
    fa_mean = np.mean(fa_parms, axis=0)
    fa_std = np.std(fa_parms, axis=0)
    return fa_parms, fa_mean, fa_std


def load_paired_datasets(img_fname, fa_fname, only_digits=False, validation_ratio=.1, test_ratio=.1):
    """vae_assoc_ujichar_img_jnt.py:21-36 in one call: images | z-scored joint parameters side by side, split into
    train / validation / test.  Returns (data_sets, fa_mean, fa_std)."""
    img_data = extract_images(fname=img_fname, only_digits=only_digits)
    fa_data, fa_mean, fa_std = extract_jnt_fa_parms(fname=fa_fname, only_digits=only_digits)
    fa_data_normed = (fa_data - fa_mean) / fa_std
    aug_data = np.concatenate((img_data, fa_data_normed), axis=1)
    return dataset.construct_datasets(aug_data, validation_ratio=validation_ratio, test_ratio=test_ratio), fa_mean, fa_std

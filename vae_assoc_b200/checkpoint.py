"""Checkpoint I/O behind save_model / restore_model (reference: tf.train.Saver over ALL variables including the Adam
slots, vae_assoc.py:70,427-463).

* save: the library's own self-describing file (`vaeassoc_save`, include/vaeassoc.h: tensor table with the TF-style names
  `image/Variable_2`, `image_1/Variable` ..., parameters + both Adam slots + step), written to exactly the file name the
  caller gave.  `save_tf_v1` exports the same variables as a TensorFlow V1 checkpoint instead.
* load: the format is recognised by its first bytes --
    "VAEASSOC"   the library's format                      -> vaeassoc_load (matched by name and shape)
    "PK"         round-1 files (numpy .npz inside .ckpt)   -> read with numpy
    otherwise    a TensorFlow V1 checkpoint table (what the reference's Saver wrote, e.g. the shipped
                 model_batchsize64_nz4_lambda8_weight50.ckpt) -> tf_checkpoint.read_v1, variables matched by TF name:
                 `<scope>/Variable_k` parameters, `<name>/Adam`, `<name>/Adam_1` slots, `beta1_power` -> step count.
"""
import io
import math

import numpy as np

from . import tf_checkpoint


def save(model, path):
    model._check(model._lib.vaeassoc_save(model._h, str(path).encode()))


def _apply_named(model, get, has, step):
    names = model.variable_names()
    missing = [n for n in names if not has(n)]
    if missing:
        raise KeyError("checkpoint lacks variables %s" % missing[:3])
    shapes = [tuple(t.shape[:t.ndim]) for t in model._tensors]
    params = []
    for n, shp in zip(names, shapes):
        a = np.asarray(get(n), dtype=np.float32)
        if tuple(a.shape) != shp:
            raise ValueError("checkpoint variable %s has shape %s, the model expects %s" % (n, a.shape, shp))
        params.append(a)
    model.set_params(params)
    if all(has(n + "/Adam") and has(n + "/Adam_1") for n in names) and step is not None:
        model.set_adam_state([np.asarray(get(n + "/Adam"), np.float32) for n in names],
                             [np.asarray(get(n + "/Adam_1"), np.float32) for n in names], int(step))


def load(model, path):
    with open(path, "rb") as f:
        head = f.read(8)
    if head == b"VAEASSOC":
        model._check(model._lib.vaeassoc_load(model._h, str(path).encode()))
        return "vaeassoc"
    if head[:2] == b"PK":
        with open(path, "rb") as f:
            data = np.load(io.BytesIO(f.read()), allow_pickle=False)
        _apply_named(model, lambda n: data[n], lambda n: n in data.files,
                     int(data["__step__"]) if "__step__" in data.files else None)
        return "npz"
    tensors = tf_checkpoint.read_v1(path)
    step = None
    if "beta1_power" in tensors:
        # TF's Adam keeps beta1^t, not t (vae_assoc.py:373: tf.train.AdamOptimizer defaults, beta1 = 0.9)
        b1p = float(np.asarray(tensors["beta1_power"]).reshape(-1)[0])
        step = int(round(math.log(b1p) / math.log(0.9))) if 0.0 < b1p < 1.0 else 0
    _apply_named(model, lambda n: tensors[n], lambda n: n in tensors, step)
    return "tf_v1"


def save_tf_v1(model, path):
    """Exports every variable a tf.train.Saver would write (parameters, Adam slots, the two beta powers) as a TensorFlow
    V1 checkpoint, so that a TensorFlow-0.x build of the reference can `restore_model` it."""
    names = model.variable_names()
    params = model.get_params()
    m, v, step = model.get_adam_state()
    out = {}
    for n, p, mi, vi in zip(names, params, m, v):
        out[n] = p
        out[n + "/Adam"] = mi
        out[n + "/Adam_1"] = vi
    out["beta1_power"] = np.float32(0.9 ** step)
    out["beta2_power"] = np.float32(0.999 ** step)
    tf_checkpoint.write_v1(path, out)

"""Checkpoint I/O behind save_model / restore_model (reference: tf.train.Saver over ALL variables including the
Adam slots, vae_assoc.py:70,427-463).  Own flat format -- a numpy .npz written to exactly the file name the caller
gave (the reference writes `<fname>` V1 checkpoints; TensorFlow's binary format is not reproduced):

    <tf name>                 parameter            e.g. "image/Variable_2", "image_1/Variable"
    <tf name>/Adam            first-moment slot    (TF slot naming)
    <tf name>/Adam_1          second-moment slot
    beta1_power, beta2_power  TF's two power accumulators (derived from the step count)
    __step__, __manifest__    Adam step t, and a JSON manifest (roles, shapes, hyper-parameters)
"""
import io
import json

import numpy as np


def save(model, path):
    names = model.variable_names()
    roles = model.variable_roles()
    params = model.get_params()
    m, v, step = model.get_adam_state()
    out = {}
    for n, p, mi, vi in zip(names, params, m, v):
        out[n] = p
        out[n + "/Adam"] = mi
        out[n + "/Adam_1"] = vi
    out["beta1_power"] = np.float32(0.9 ** step)
    out["beta2_power"] = np.float32(0.999 ** step)
    out["__step__"] = np.int64(step)
    manifest = dict(format="vae_assoc_b200/1", names=names, roles=[[int(a), b] for a, b in roles],
                    shapes=[list(p.shape) for p in params], batch_size=int(model.batch_size), n_z=int(model.n_z),
                    learning_rate=float(model.learning_rate), assoc_lambda=float(model.assoc_lambda),
                    weights=[float(w) for w in model.weights], binary=[bool(b) for b in model.binary])
    out["__manifest__"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    buf = io.BytesIO()
    np.savez(buf, **out)
    with open(path, "wb") as f:          # exact file name (np.savez would append ".npz")
        f.write(buf.getvalue())


def load(model, path):
    with open(path, "rb") as f:
        data = np.load(io.BytesIO(f.read()), allow_pickle=False)
    names = model.variable_names()
    missing = [n for n in names if n not in data.files]
    if missing:
        raise KeyError("checkpoint %s lacks variables %s" % (path, missing[:3]))
    model.set_params([data[n] for n in names])
    if all((n + "/Adam") in data.files for n in names):
        model.set_adam_state([data[n + "/Adam"] for n in names], [data[n + "/Adam_1"] for n in names],
                             int(data["__step__"]))

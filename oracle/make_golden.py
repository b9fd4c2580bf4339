"""Mints the pinned fixtures under tests/golden/ from the fp64 oracle.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden

The reference ships no golden vectors and cannot run here, so these are OUR fixtures:
they freeze the oracle (whose backward is cross-checked against torch autograd and finite differences) so that a
later edit of the oracle cannot silently move the goal posts.  Inputs are regenerated from seeds; only small
outputs are stored.
"""
import os

import numpy as np

from . import philox, synth
from . import vae_assoc_oracle as vo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def tiny_archs(n_z=3):
    img = dict(scope="image", hidden_conv=False, n_hidden_recog_1=16, n_hidden_recog_2=12, n_hidden_gener_1=16,
               n_hidden_gener_2=12, n_input=20, n_z=n_z)
    jnt = dict(scope="joint", hidden_conv=False, n_hidden_recog_1=10, n_hidden_recog_2=8, n_hidden_gener_1=10,
               n_hidden_gener_2=8, n_input=7, n_z=n_z)
    return [img, jnt]


def case_inputs(archs, batch, seed):
    X = synth.synth_batch(archs, [True, False], data_seed=seed, proj_seed=1, row0=0, n_rows=batch)
    eps = [philox.eps_rows(seed, t, 0, batch, archs[0]["n_z"]) for t in range(3)]
    return X, eps


def run_case(archs, batch, transfer_fct, seed, lam=8.0, weights=(50.0, 1.0)):
    X, eps = case_inputs(archs, batch, seed)
    o = vo.OracleAssocVAE(archs, [True, False], transfer_fct, list(weights), lam, 1e-3, batch, seed=seed)
    rec = dict(costs=[], grad_sum=[], grad_abs=[], z_mean0=[], z_mean1=[], vae_costs=[], assoc=[])
    for t in range(3):
        c, g, pr = o.loss_and_grads(X, eps[t])
        rec["costs"].append(c)
        rec["grad_sum"].append([x.sum() for gs in g for x in gs])
        rec["grad_abs"].append([np.abs(x).sum() for gs in g for x in gs])
        rec["z_mean0"].append(pr["z_means"][0]); rec["z_mean1"].append(pr["z_means"][1])
        rec["vae_costs"].append(pr["vae_costs"]); rec["assoc"].append(pr["assoc_costs"][0])
        o.adam_step(g)
    rec["param_sum"] = [p.sum() for ps in o.params for p in ps]
    rec["param_abs"] = [np.abs(p).sum() for ps in o.params for p in ps]
    return {k: np.asarray(v, dtype=np.float64) for k, v in rec.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "tiny_relu": (tiny_archs(), 5, "relu", 11),
        "tiny_softplus": (tiny_archs(), 5, "softplus", 12),
        "ref_relu_b100": (vo.reference_archs(4), 100, "relu", 0),
        "ref_softplus_b16": (vo.reference_archs(4), 16, "softplus", 1),
    }
    out = {}
    for name, (archs, batch, f, seed) in cases.items():
        r = run_case(archs, batch, f, seed)
        for k, v in r.items():
            out["%s/%s" % (name, k)] = v
    # conv variant: forward/backward of one step (tensor sums only)
    archs = vo.reference_archs(4, conv=True)
    X, eps = case_inputs(archs, 4, 5)
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, 4, seed=5)
    c, g, pr = o.loss_and_grads(X, eps[0])
    out["conv_relu_b4/cost"] = np.float64(c)
    out["conv_relu_b4/grad_sum"] = np.asarray([x.sum() for gs in g for x in gs])
    out["conv_relu_b4/grad_abs"] = np.asarray([np.abs(x).sum() for gs in g for x in gs])
    # generator + Philox
    out["philox/eps_seed7_step3_row5"] = philox.eps_rows(7, 3, 5, 4, 6)
    Xs = synth.synth_batch(vo.reference_archs(4), [True, False], 0, 1, 1000, 3)
    out["synth/img_rows1000"] = Xs[0].astype(np.float32)
    out["synth/jnt_rows1000"] = Xs[1].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "oracle_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "oracle_golden.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()

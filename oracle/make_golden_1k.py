"""Mints tests/golden/latents_1k.npz: latent codes of the fp64 oracle after 1000 train steps.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_1k          (about 2-3 minutes of CPU)

BASELINE.json north_star: "latent codes must agree within the same tolerance after 1k steps".  The reference cannot run
here, so the frozen run is the CPU restatement: reference config (vae_assoc_ujichar_img_jnt.py:53-71
shapes, batch 100 = vae_assoc.py:27 default, relu as train() selects at vae_assoc.py:502, weights [50,1], lambda 8,
lr 1e-3), a 1000-pair synthetic data set cycled as train() does (vae_assoc.py:540-541: 10 batches per epoch, 100
epochs), eps of step t injected from the Philox stream.  Only small outputs are stored; inputs come from seeds
(`case()` below is what the GPU test calls to rebuild them).

Two runs are frozen: the fp64 numpy oracle (`z_img_k`, `z_jnt_k`, `costs`) and the SAME graph in fp32 -- the precision
the reference's TensorFlow graph computes in (tf.float32 placeholders / variables, vae_assoc.py:54,186) -- by the
torch-CPU twin (`z32_img_k`, `z32_jnt_k`, `costs32`).  Over 1000 Adam steps fp32 arithmetic drifts away from fp64
SYSTEMATICALLY (5.8e-4 of max|z| after 10 steps, 4.7e-2 after 100, ~1e-1 after 1000: identical to three digits for the
torch-CPU fp32 twin and for the CUDA fp32 path, see scripts/fp32_twin_drift.py), so "the reference's own
implementation" after 1k steps is represented by the fp32 run, and the fp64 run documents the precision effect.
"""
import os

import numpy as np

from . import philox, synth
from . import vae_assoc_oracle as vo

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "latents_1k.npz")
SEED, BATCH, N_DATA, STEPS = 7, 100, 1000, 1000
CHECKPOINTS = (1, 10, 100, 1000)


def case():
    """(archs, initial params as fp32-representable fp64, data [1000, .] per modality as fp32, eps(t))"""
    archs = vo.reference_archs(4)
    params = [[p.astype(np.float32).astype(np.float64) for p in ps] for ps in vo.init_params(archs, SEED)]
    data = [x.astype(np.float32) for x in synth.synth_batch(archs, [True, False], data_seed=SEED, proj_seed=1, row0=0, n_rows=N_DATA)]
    eps = lambda t: philox.eps_rows(SEED, t, 0, BATCH, 4).astype(np.float32)
    return archs, params, data, eps


def batch_of(data, t):
    i = (t % (N_DATA // BATCH)) * BATCH
    return [d[i:i + BATCH] for d in data]


def main():
    archs, params, data, eps = case()
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, BATCH, params=params)
    probe = batch_of(data, 0)
    rec = {"costs": []}
    for t in range(STEPS):
        rec["costs"].append(o.partial_fit([x.astype(np.float64) for x in batch_of(data, t)], eps(t).astype(np.float64)))
        if t + 1 in CHECKPOINTS:
            z = o.transform([x.astype(np.float64) for x in probe])
            rec["z_img_%d" % (t + 1)] = z[0]
            rec["z_jnt_%d" % (t + 1)] = z[1]
    # the same run in fp32 (torch-CPU twin, the reference graph at its own precision)
    import torch
    from . import torch_twin
    torch.set_num_threads(1)              # fixed summation order inside the CPU matmuls
    tw = torch_twin.TorchAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, BATCH, params, dtype=torch.float32)
    probe32 = [torch.tensor(x, dtype=torch.float32) for x in probe]
    rec["costs32"] = []
    for t in range(STEPS):
        X = [torch.tensor(x, dtype=torch.float32) for x in batch_of(data, t)]
        rec["costs32"].append(tw.partial_fit(X, torch.tensor(eps(t), dtype=torch.float32)))
        if t + 1 in CHECKPOINTS:
            with torch.no_grad():
                rec["z32_img_%d" % (t + 1)] = tw.encode(0, probe32[0])[0].numpy()
                rec["z32_jnt_%d" % (t + 1)] = tw.encode(1, probe32[1])[0].numpy()
    np.savez_compressed(OUT, **{k: np.asarray(v, dtype=np.float64) for k, v in rec.items()})
    print("wrote", OUT, "final cost", rec["costs"][-1])


if __name__ == "__main__":
    main()

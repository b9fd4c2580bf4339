"""Oracle of the device-resident synthetic paired-batch generator.  TEST INFRASTRUCTURE ONLY.

The generator REPLACES the reference's input path (dataset.py:22-72 + utils.py:142-195, which need
pickles we cannot download); it only mirrors that path's value ranges and layout:
  * image modality ("binary"): stroke-like, background exactly 0, lit pixels in [0.5, 1]
    (cf. zero canvas utils.py:245 and the /255 scaling utils.py:150);
  * joint modality: z-scored per column (vae_assoc_ujichar_img_jnt.py:29);
  * one fp32 row per pair, modalities side by side (vae_assoc_ujichar_img_jnt.py:34, vae_assoc.py:543).

Definition (global sample index i, shared code c_i ~ N(0, I_4)):
    image:  x[d] = 1{ sigmoid(2 * c_i . P[:, d] - 3.5) > u1 } * (0.5 + 0.5 * u2)
    joint:  x[d] = (c_i . P[:, d] + 0.3 * n) / sqrt(|P[:, d]|^2 + 0.09)
with P [4, n_input] ~ N(0, 1) fixed by `proj_seed`.  All randomness is Philox4x32-10 (oracle/philox.py);
the CUDA implementation is vae_assoc_b200/csrc/synth.cu.
"""
import numpy as np

from . import philox as px

CODE_DIM = 4


def projection(proj_seed, modality, n_input):
    """P [4, n_input]: element idx = k*n_input + d -> component idx%4 of philox((idx//4, 0, 0, 0), (seed, TAG_PROJ+m))."""
    n = CODE_DIM * n_input
    blk = np.arange((n + 3) // 4, dtype=np.uint64)
    v = px.normal4(blk, 0, 0, 0, proj_seed, px.TAG_PROJ + modality).reshape(-1)[:n]
    return v.reshape(CODE_DIM, n_input).astype(np.float32).astype(np.float64)


def codes(data_seed, row0, n_rows):
    return px.normal_rows(data_seed, px.TAG_CODE, row0, n_rows, CODE_DIM).astype(np.float32).astype(np.float64)


def synth_modality(data_seed, proj_seed, modality, n_input, binary, row0, n_rows):
    P = projection(proj_seed, modality, n_input)
    c = codes(data_seed, row0, n_rows)
    proj = c @ P
    rows = np.arange(n_rows, dtype=np.uint64) + np.uint64(row0)
    rlo, rhi = (rows & px.MASK)[:, None], (rows >> np.uint64(32))[:, None]
    if binary:
        nblk = (n_input + 1) // 2
        w = px.philox4x32_10(rlo, rhi, np.arange(nblk, dtype=np.uint64)[None, :], modality, data_seed, px.TAG_IMG)
        u = np.stack([px.u01(x) for x in w], axis=-1).reshape(n_rows, nblk * 2, 2)[:, :n_input, :]
        p = 1.0 / (1.0 + np.exp(-(2.0 * proj - 3.5)))
        return np.where(p > u[..., 0], 0.5 + 0.5 * u[..., 1], 0.0)
    nblk = (n_input + 3) // 4
    n = px.normal4(rlo, rhi, np.arange(nblk, dtype=np.uint64)[None, :], modality, data_seed, px.TAG_JNT)
    n = n.reshape(n_rows, nblk * 4)[:, :n_input]
    inv = 1.0 / np.sqrt((P * P).sum(0) + 0.09)
    return (proj + 0.3 * n) * inv[None, :]


def synth_batch(archs, binary, data_seed, proj_seed, row0, n_rows):
    return [synth_modality(data_seed, proj_seed, m, na["n_input"], binary[m], row0, n_rows)
            for m, na in enumerate(archs)]

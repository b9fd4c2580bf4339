"""Philox4x32-10 counter-based generator (numpy, vectorised) -- oracle side.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The CUDA side is
``vae_assoc_b200/csrc/philox.cuh``; both follow the published Random123
algorithm (Salmon et al., SC'11) and are pinned by its known-answer vectors
(tests/test_oracle_philox.py).

The reference draws its noise with ``tf.random_normal`` (vae_assoc.py:90) and
``np.random.normal`` (vae_assoc.py:414); TensorFlow's stream cannot be
reproduced without TensorFlow, so the new build defines its own counter layout:

    key     = (seed, tag)
    counter = (row_lo, row_hi, block, step)

``row`` is the GLOBAL sample index, so a G-way batch shard reproduces the 1-GPU
stream.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# stream tags (second key word)
TAG_EPS = 1       # reparameterisation noise           (vae_assoc.py:90)
TAG_CODE = 2      # synthetic generator: shared code c
TAG_IMG = 3       # synthetic generator: image uniforms
TAG_JNT = 4       # synthetic generator: joint noise
TAG_PROJ = 16     # synthetic generator: projection matrices (+ modality index)
TAG_PRIOR = 5     # generate(z_mu=None) prior draw      (vae_assoc.py:414)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable integer arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01(x):
    """uint32 -> uniform in (0,1), exactly representable in fp32: ((x>>8)+0.5)*2^-24."""
    return ((np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 0.5) * (1.0 / 16777216.0)


def box_muller(xa, xb):
    """Two uint32 words -> two standard normals (r cos t, r sin t)."""
    u1 = u01(xa)
    u2 = u01(xb)
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2.0 * np.pi * u2), r * np.sin(2.0 * np.pi * u2)


def normal4(c0, c1, c2, c3, k0, k1):
    """One Philox call -> 4 standard normals, stacked on a new last axis."""
    w = philox4x32_10(c0, c1, c2, c3, k0, k1)
    n0, n1 = box_muller(w[0], w[1])
    n2, n3 = box_muller(w[2], w[3])
    return np.stack([n0, n1, n2, n3], axis=-1)


def normal_rows(seed, tag, row0, n_rows, n_cols, step=0):
    """[n_rows, n_cols] standard normals; element (r, c) = component c%4 of
    philox(counter=(row_lo, row_hi, c//4, step), key=(seed, tag)), row = row0 + r."""
    rows = np.arange(n_rows, dtype=np.uint64) + np.uint64(row0)
    nblk = (n_cols + 3) // 4
    blk = np.arange(nblk, dtype=np.uint64)
    out = normal4((rows & MASK)[:, None], (rows >> np.uint64(32))[:, None], blk[None, :], step, seed, tag)
    return out.reshape(n_rows, nblk * 4)[:, :n_cols]


def eps_rows(seed, step, row0, n_rows, n_z):
    """Reparameterisation noise for global rows [row0, row0+n_rows) at optimiser step `step`."""
    return normal_rows(seed, TAG_EPS, row0, n_rows, n_z, step=step)

"""CPU tests of the oracle itself: Philox known-answer vectors, golden fixtures, two independent gradient
derivations, finite differences.  (The reference has no tests; parity is pinned by these.)"""
import os

import numpy as np
import pytest

from oracle import make_golden, philox, synth
from oracle import vae_assoc_oracle as vo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.npz")


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = tuple(int(x) for x in philox.philox4x32_10(*ctr, *key))
        assert got == want


def test_philox_normal_moments():
    n = philox.normal_rows(3, philox.TAG_EPS, 0, 20000, 4)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1.0) < 0.02
    # row/step addressing: a shard reproduces the global stream
    a = philox.eps_rows(9, 5, 0, 64, 4)
    b = philox.eps_rows(9, 5, 32, 32, 4)
    assert np.array_equal(a[32:], b)


@pytest.mark.parametrize("name", ["tiny_relu", "tiny_softplus", "ref_relu_b100", "ref_softplus_b16"])
def test_oracle_matches_golden(name):
    g = np.load(GOLD)
    cases = {"tiny_relu": (make_golden.tiny_archs(), 5, "relu", 11),
             "tiny_softplus": (make_golden.tiny_archs(), 5, "softplus", 12),
             "ref_relu_b100": (vo.reference_archs(4), 100, "relu", 0),
             "ref_softplus_b16": (vo.reference_archs(4), 16, "softplus", 1)}
    r = make_golden.run_case(*cases[name])
    for k, v in r.items():
        np.testing.assert_allclose(v, g["%s/%s" % (name, k)], rtol=1e-9, atol=1e-12, err_msg=k)


def test_golden_generator_and_eps():
    g = np.load(GOLD)
    np.testing.assert_allclose(philox.eps_rows(7, 3, 5, 4, 6), g["philox/eps_seed7_step3_row5"], rtol=0, atol=1e-12)
    Xs = synth.synth_batch(vo.reference_archs(4), [True, False], 0, 1, 1000, 3)
    np.testing.assert_array_equal(Xs[0].astype(np.float32), g["synth/img_rows1000"])
    np.testing.assert_allclose(Xs[1].astype(np.float32), g["synth/jnt_rows1000"], rtol=1e-6)


def _rand_case(conv, f, B=6, seed=3, nz=4):
    archs = vo.reference_archs(nz, conv=conv)
    rng = np.random.RandomState(seed)
    X = synth.synth_batch(archs, [True, False], 0, 1, 0, B)
    eps = rng.normal(size=(B, nz))
    P = vo.init_params(archs, seed)
    P = [[p if p.ndim > 1 else rng.normal(size=p.shape) * 0.05 for p in ps] for ps in P]   # exercise the biases
    return archs, X, eps, P


@pytest.mark.parametrize("conv,f", [(False, "relu"), (False, "softplus"), (True, "relu"), (True, "softplus")])
def test_backward_vs_autograd(conv, f):
    from oracle import torch_twin as tt
    archs, X, eps, P = _rand_case(conv, f)
    o = vo.OracleAssocVAE(archs, [True, False], f, [50.0, 1.0], 8.0, 1e-3, 6, params=P)
    c, g, _ = o.loss_and_grads(X, eps)
    t = tt.TorchAssocVAE(archs, [True, False], f, [50.0, 1.0], 8.0, 1e-3, 6, P)
    c2, g2 = t.cost_and_grads(X, eps)
    assert abs(c - c2) <= 1e-12 * abs(c2)
    for ga, gb in zip(g, g2):
        for a, b in zip(ga, gb):
            assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-30)


def test_backward_vs_finite_differences():
    archs = make_golden.tiny_archs()
    X, eps = make_golden.case_inputs(archs, 5, 11)
    o = vo.OracleAssocVAE(archs, [True, False], "softplus", [50.0, 1.0], 8.0, 1e-3, 5, seed=11)
    c, g, _ = o.loss_and_grads(X, eps[0])
    rng = np.random.RandomState(0)
    for m in range(2):
        for i, p in enumerate(o.params[m]):
            for _ in range(3):
                idx = tuple(rng.randint(0, s) for s in p.shape)
                old = p[idx]
                h = 1e-6
                p[idx] = old + h
                cp = o.loss(X, o.forward(X, eps[0]))["cost"]
                p[idx] = old - h
                cm = o.loss(X, o.forward(X, eps[0]))["cost"]
                p[idx] = old
                fd = (cp - cm) / (2 * h)
                assert abs(fd - g[m][i][idx]) <= 1e-5 * max(1.0, abs(fd)), (m, i, idx, fd, g[m][i][idx])


def test_adam_matches_twin_three_steps():
    from oracle import torch_twin as tt
    import torch
    archs = make_golden.tiny_archs()
    X, eps = make_golden.case_inputs(archs, 5, 11)
    P = vo.init_params(archs, 11)
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, 5, params=P)
    t = tt.TorchAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, 5, P)
    Xt = [torch.tensor(x) for x in X]
    for k in range(3):
        c1 = o.partial_fit(X, eps[k])
        c2 = t.partial_fit(Xt, torch.tensor(eps[k]))
        assert abs(c1 - c2) <= 1e-10 * abs(c2)
    flat = [p for ps in o.params for p in ps]
    for a, b in zip(flat, t.flat):
        np.testing.assert_allclose(a, b.detach().numpy(), rtol=1e-9, atol=1e-12)


def test_sharded_loss_sums_to_global():
    """mean terms scaled by 1/B_global, sum terms unscaled => shard costs/grads add up."""
    archs = make_golden.tiny_archs()
    X, eps = make_golden.case_inputs(archs, 8, 4)
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, 8, seed=4)
    c, g, _ = o.loss_and_grads(X, eps[0])
    cs, gs = 0.0, None
    for r in range(2):
        sl = slice(4 * r, 4 * r + 4)
        ci, gi, _ = o.loss_and_grads([x[sl] for x in X], eps[0][sl], global_batch=8)
        cs += ci
        gs = gi if gs is None else [[a + b for a, b in zip(x, y)] for x, y in zip(gs, gi)]
    assert abs(cs - c) <= 1e-12 * abs(c)
    for ga, gb in zip(g, gs):
        for a, b in zip(ga, gb):
            np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)


def test_deconv_same_is_shifted_relative_to_torch_default():
    """TF SAME transposed conv crops the full output at pad_before = 1 (k=5, s=2)."""
    rng = np.random.RandomState(0)
    y = rng.normal(size=(1, 3, 3, 2)); w = rng.normal(size=(5, 5, 4, 2))
    out = vo.conv2d_transpose(y, w, 2, "SAME")
    assert out.shape == (1, 6, 6, 4)
    full = np.zeros((1, 9, 9, 4))
    for iy in range(3):
        for ix in range(3):
            for ky in range(5):
                for kx in range(5):
                    full[0, iy * 2 + ky, ix * 2 + kx] += w[ky, kx] @ y[0, iy, ix]
    np.testing.assert_allclose(out, full[:, 1:7, 1:7], rtol=1e-12)


def test_latents_1k_golden_prefix():
    """tests/golden/latents_1k.npz (1000-step oracle run, `python -m oracle.make_golden_1k`): the first 10 steps are
    re-run here and must reproduce the frozen costs and latent codes exactly (the full run takes minutes)."""
    from oracle import make_golden_1k as g1k
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "latents_1k.npz"))
    assert gold["costs"].shape == (g1k.STEPS,) and np.isfinite(gold["costs"]).all()
    archs, params, data, eps = g1k.case()
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, g1k.BATCH, params=params)
    probe = [x.astype(np.float64) for x in g1k.batch_of(data, 0)]
    for t in range(10):
        c = o.partial_fit([x.astype(np.float64) for x in g1k.batch_of(data, t)], eps(t).astype(np.float64))
        np.testing.assert_allclose(c, gold["costs"][t], rtol=1e-12)
        if t + 1 in (1, 10):
            z = o.transform(probe)
            np.testing.assert_allclose(z[0], gold["z_img_%d" % (t + 1)], rtol=1e-10, atol=1e-12)
            np.testing.assert_allclose(z[1], gold["z_jnt_%d" % (t + 1)], rtol=1e-10, atol=1e-12)
    # training made progress: the cost after 1000 steps is far below the initial one
    assert gold["costs"][-1] < 0.6 * gold["costs"][0]

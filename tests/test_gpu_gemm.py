"""Unit parity of the dense-layer contractions (SIMT fp32 and tcgen05 tf32 kernels) through the C-ABI test hook
`vaeassoc_debug_gemm`, against numpy fp64.  Needs a B200: `pytest -m gpu`.

For the tcgen05 kernel the inputs are pre-rounded to tf32 (10-bit mantissa), so the products are exact in fp32 and
the only error left is fp32 accumulation: any mistake in the TMA boxes, swizzle, UMMA descriptors or the TMEM
epilogue shows up as O(1) error, not as rounding noise."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NN, NT, TN = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SOFTPLUS, ACT_SIGMOID = 0, 1, 2, 3


@pytest.fixture(scope="module")
def model():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vae_assoc_b200 import build, vae_assoc
    build.build(verbose=False)
    archs = [dict(scope="image", hidden_conv=False, n_hidden_recog_1=8, n_hidden_recog_2=8, n_hidden_gener_1=8,
                  n_hidden_gener_2=8, n_input=16, n_z=2)]
    m = vae_assoc.AssocVariationalAutoEncoder(archs, batch_size=4, precision="fp32")
    yield m
    m.close()


def tf32_round(x):
    """round-to-nearest (ties away) to a 10-bit mantissa, like cvt.rna.tf32.f32"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def act(kind, v):
    if kind == ACT_RELU:
        return np.maximum(v, 0)
    if kind == ACT_SOFTPLUS:
        return np.logaddexp(0, v)
    if kind == ACT_SIGMOID:
        return 1 / (1 + np.exp(-v))
    return v


def act_grad(kind, h):
    if kind == ACT_RELU:
        return (h > 0).astype(np.float64)
    if kind == ACT_SOFTPLUS:
        return 1 - np.exp(-h)
    return np.ones_like(h)


def run(model, kind, use_tc, M, N, K, seed, pad=0, a=ACT_NONE, with_bias=True):
    import torch
    rng = np.random.RandomState(seed)
    dev = model._dev

    def mat(r, c):
        ld = c + pad
        ld += (-ld) % 4
        full = np.zeros((r, ld), np.float32)
        full[:, :c] = tf32_round(rng.normal(size=(r, c)))
        return full, ld

    if kind == NN:
        A, lda = mat(M, K); B, ldb = mat(K, N)
    elif kind == NT:
        A, lda = mat(M, K); B, ldb = mat(N, K)
    else:
        A, lda = mat(K, M); B, ldb = mat(K, N)
    ldc = N + pad + ((-(N + pad)) % 4)
    Cinit = rng.normal(size=(M, ldc)).astype(np.float32) if kind == TN else np.full((M, ldc), 7.0, np.float32)
    bias = tf32_round(rng.normal(size=N)) if (kind == NN and with_bias) else None
    aux = np.abs(rng.normal(size=(M, ldc))).astype(np.float32) * (rng.uniform(size=(M, ldc)) > 0.4) if kind == NT else None
    bgrad0 = rng.normal(size=N).astype(np.float32) if kind == TN else None
    tA, tB, tC = (torch.as_tensor(x).to(dev) for x in (A, B, Cinit))
    tb = torch.as_tensor(bias).to(dev) if bias is not None else None
    taux = torch.as_tensor(aux.astype(np.float32)).to(dev) if aux is not None else None
    tbg = torch.as_tensor(bgrad0).to(dev) if bgrad0 is not None else None
    torch.cuda.synchronize()
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    rc = model._lib.vaeassoc_debug_gemm(model._h, kind, int(use_tc), M, N, K, p(tA), lda, p(tB), ldb, p(tC), ldc, p(tb),
                                        p(tbg), p(taux), ldc, a, 0)
    assert rc == 0, model._lib.vaeassoc_last_error(model._h).decode()
    got = tC.cpu().numpy()
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    if kind == NN:
        ref = A64[:, :K] @ B64[:, :N]
        if bias is not None:
            ref = ref + bias.astype(np.float64)
        ref = act(a, ref)
    elif kind == NT:
        ref = (A64[:, :K] @ B64[:, :K].T) * act_grad(a, aux[:, :N].astype(np.float64))
    else:
        ref = Cinit[:, :N].astype(np.float64) + A64[:, :M].T @ B64[:, :N]
    err = np.abs(got[:, :N] - ref).max() / max(np.abs(ref).max(), 1e-30)
    # padding columns: the TMA epilogue of the tcgen05 kernels clips in 16-byte units, so columns N .. roundup4(N)
    # may receive exact zeros (NN / NT) or +0 (TN); everything beyond stays untouched
    N4 = N + (-N) % 4
    if ldc > N4:
        assert np.array_equal(got[:, N4:], Cinit[:, N4:]), "kernel wrote outside the logical columns"
    if N4 > N:
        padc = got[:, N:N4]
        assert np.array_equal(padc, Cinit[:, N:N4]) or (kind != TN and not padc.any()), "bad values in the pad columns"
    if kind == TN:
        bg = tbg.cpu().numpy()
        bref = bgrad0.astype(np.float64) + B64[:, :N].sum(0)
        assert np.abs(bg - bref).max() / np.abs(bref).max() < 1e-5
    return err


SHAPES = [(256, 160, 96), (128, 128, 32), (100, 147, 200), (8192, 500, 784), (333, 500, 500), (64, 784, 500),
          (1000, 200, 147), (129, 33, 40)]


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_forward_nn(model, use_tc, M, N, K):
    for a in (ACT_NONE, ACT_RELU, ACT_SIGMOID):
        # sigmoid output is O(1) whatever the pre-activation: its error is the ABSOLUTE fp32 accumulation error of a
        # K-term sum (~|sum| sqrt(K) 2^-24), not a relative one
        assert run(model, NN, use_tc, M, N, K, seed=M + N + K + a, a=a) < (5e-5 if a == ACT_SIGMOID else 5e-6)


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_dgrad_nt(model, use_tc, M, N, K):
    for a in (ACT_NONE, ACT_RELU, ACT_SOFTPLUS):
        assert run(model, NT, use_tc, M, N, K, seed=M + 2 * N + K + a, a=a) < 5e-6


@pytest.mark.parametrize("use_tc", [0, 1])
@pytest.mark.parametrize("M,N,K", [(160, 256, 96), (500, 784, 8192), (784, 500, 8192), (147, 200, 1000), (200, 147, 100),
                                   (33, 129, 64), (500, 500, 333)])
def test_wgrad_tn(model, use_tc, M, N, K):
    assert run(model, TN, use_tc, M, N, K, seed=M + N + 3 * K) < 5e-6


def test_tc_unpadded_leading_dimension_with_slack(model):
    """row pitches that are multiples of 4 floats but wider than the logical width (e.g. 147 -> 148)"""
    assert run(model, NN, 1, 300, 147, 200, seed=1, pad=1) < 5e-6
    assert run(model, NT, 1, 300, 200, 147, seed=2, pad=1) < 5e-6
    assert run(model, TN, 1, 147, 200, 300, seed=3, pad=1) < 5e-6


SKINNY = [(8192, 8, 500), (100, 8, 200), (777, 6, 333), (8192, 500, 4), (100, 200, 4), (65, 147, 3), (300, 16, 64)]


@pytest.mark.parametrize("M,N,K", SKINNY)
def test_skinny_forward_and_dgrad(model, M, N, K):
    """the n_z-wide layers (heads N = 2 n_z, decoder input K = n_z) on the HBM-bound skinny kernels"""
    for a in (ACT_NONE, ACT_RELU):
        assert run(model, NN, 0, M, N, K, seed=M + N + K + a, a=a) < 5e-6
        assert run(model, NT, 0, M, N, K, seed=M + 2 * N + K + a, a=a) < 5e-6


@pytest.mark.parametrize("M,N,K", [(4, 500, 8192), (500, 8, 8192), (3, 147, 65), (200, 6, 100), (16, 64, 300), (64, 16, 300)])
def test_skinny_wgrad(model, M, N, K):
    assert run(model, TN, 0, M, N, K, seed=M + N + 3 * K) < 5e-6

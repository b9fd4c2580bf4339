"""hidden_conv=True image modality (BASELINE configs[3]): conv encoder (vae_assoc.py:169-199),
deconv.py transposed-conv decoder with its sigmoid after every layer (vae_assoc.py:249-291, :491), dense joint
modality -- CUDA path (im2col -> GEMM -> col2im) against the fp64 oracle, whose conv arithmetic is itself checked
against torch's conv2d / conv_transpose2d in tests/test_oracle.py.  Needs a B200: `pytest -m gpu`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import philox, synth                     # noqa: E402
from oracle import vae_assoc_oracle as vo            # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def va():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from vae_assoc_b200 import build, vae_assoc
    build.build(verbose=False)
    return vae_assoc


def pair(va, batch, f, precision, seed):
    archs = vo.reference_archs(4, conv=True)
    model = va.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=f, weights=[50, 1], assoc_lambda=8,
                                           learning_rate=1e-3, batch_size=batch, precision=precision, seed=seed)
    params = model.get_params()
    roles = model.variable_roles()
    assert [r for m, r in roles if m == 0] == vo.param_names(archs[0])
    assert [tuple(p.shape) for (m, _), p in zip(roles, params) if m == 0] == [tuple(s) for s in vo.param_shapes(archs[0])]
    rng = np.random.RandomState(seed + 7)
    per_mod = [[], []]
    for (m, _), p in zip(roles, params):
        per_mod[m].append(p.astype(np.float64) if p.ndim > 1 else rng.normal(size=p.shape) * 0.05)
    model.set_params(per_mod)
    per_mod = [[p.astype(np.float32).astype(np.float64) for p in ps] for ps in per_mod]
    oracle = vo.OracleAssocVAE(archs, [True, False], f, [50.0, 1.0], 8.0, 1e-3, batch, params=per_mod)
    return archs, model, oracle


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("tf32", 2e-3)])
@pytest.mark.parametrize("batch", [3, 64])
def test_conv_gradient_step(va, precision, tol, batch):
    """The image modality has no relu (linear convs, sigmoid deconvs); with softplus for the dense joint modality
    the whole graph is smooth, so the tf32 path is held to the 2e-3 bound against the EXACT oracle."""
    archs, model, oracle = pair(va, batch, "softplus", precision, seed=batch)
    X = [x.astype(np.float32) for x in synth.synth_batch(archs, [True, False], 3, 1, 0, batch)]
    eps = philox.eps_rows(3, 0, 0, batch, 4).astype(np.float32)
    cost = model.compute_gradients(X, eps)
    c_ref, g_ref, pr = oracle.loss_and_grads(X, eps)
    assert abs(cost - c_ref) <= tol * abs(c_ref), (cost, c_ref)
    for m in range(2):
        assert rel(model.z_means[m], pr["z_means"][m]) < tol
        assert rel(model.z_log_sigma_sqs[m], pr["z_log_sigma_sqs"][m]) < tol
        assert rel(model.x_reconstr_means[m], pr["x_reconstr_means"][m]) < tol
        assert rel(model.d_z_means[m], pr["d_z_means"][m]) < tol
    for g, r, n in zip(model.get_grads(), [g for gs in g_ref for g in gs], model.variable_roles()):
        assert g.shape == r.shape, (n, g.shape, r.shape)
        assert rel(g, r) < tol, (n, rel(g, r))
    model.close()


def test_conv_training_and_inference(va):
    archs, model, oracle = pair(va, 16, "relu", "fp32", seed=2)
    for t in range(3):
        X = [x.astype(np.float32) for x in synth.synth_batch(archs, [True, False], 0, 1, 16 * t, 16)]
        eps = philox.eps_rows(1, t, 0, 16, 4).astype(np.float32)
        c, c_ref = model.partial_fit(X, eps), oracle.partial_fit(X, eps)
        assert abs(c - c_ref) <= 1e-4 * abs(c_ref), (t, c, c_ref)
    z = np.zeros((16, 4), np.float32); z[0] = [1.0, -0.5, 0.3, 2.0]
    for a, b in zip(model.generate(z), oracle.generate(z)):
        assert rel(a, b) < 1e-4
    for a, b in zip(model.transform(X), oracle.transform(X)):
        assert rel(a, b) < 1e-4
    for a, b in zip(model.reconstruct(X, eps=eps), oracle.reconstruct(X, eps=eps)):
        assert rel(a, b) < 1e-4
    model.close()

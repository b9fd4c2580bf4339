"""Per-tensor parity report against the oracle (debug helper; run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import philox, synth, vae_assoc_oracle as vo
from vae_assoc_b200 import vae_assoc as va

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
fct = os.environ.get("FCT", "relu")
emul = os.environ.get("EMUL", "0") == "1"
for batch in [int(x) for x in (sys.argv[2:] or ["100", "8192"])]:
    archs = vo.reference_archs(4)
    model = va.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=fct, weights=[50, 1], assoc_lambda=8,
                                           learning_rate=1e-3, batch_size=batch, precision=precision, seed=5)
    params = model.get_params()
    per_mod = [[p.astype(np.float64) for p in params[:14]], [p.astype(np.float64) for p in params[14:]]]
    oracle = vo.OracleAssocVAE(archs, [True, False], fct, [50., 1.], 8.0, 1e-3, batch, params=per_mod, emulate_tf32=emul)
    xs = model.synth_batch(0, batch)
    X = [x.cpu().numpy() for x in xs]
    eps = philox.eps_rows(5, 0, 0, batch, 4).astype(np.float32)
    cost = model.compute_gradients(xs, eps)
    c_ref, g_ref, pr = oracle.loss_and_grads(X, eps)
    print("== batch", batch, precision, "cost", cost, c_ref, abs(cost - c_ref) / abs(c_ref))
    for m in range(2):
        print("  z_mean", m, rel(model.z_means[m], pr["z_means"][m]), "lv", rel(model.z_log_sigma_sqs[m], pr["z_log_sigma_sqs"][m]),
              "z", rel(model.z_array[m], pr["z_array"][m]), "xh", rel(model.x_reconstr_means[m], pr["x_reconstr_means"][m]),
              "dmu", rel(model.d_z_means[m], pr["d_z_means"][m]), "dlv", rel(model.d_z_log_sigma_sqs[m], pr["d_z_log_sigma_sqs"][m]))
    for g, r, n in zip(model.get_grads(), [g for gs in g_ref for g in gs], model.variable_roles()):
        e = np.abs(g.astype(np.float64) - r)
        idx = np.unravel_index(e.argmax(), e.shape)
        print("  %-10s rel %.3e  max|ref| %.4g  worst at %s got %.6g ref %.6g" % (n, rel(g, r), np.abs(r).max(), idx, g[idx], r[idx]))
    model.close()

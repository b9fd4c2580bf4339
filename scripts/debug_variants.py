"""One gradient step of the reference config under the small-batch switches of the tile kernel: every variant must agree
with the plain one (32-deep k-blocks, 128 rows per CTA, no split-K) to summation order."""
import os, subprocess, sys, json
import numpy as np
VARIANTS = {"plain": "VAEASSOC_NO_DEEP_K=1 VAEASSOC_NO_SMALL_ROWS=1 VAEASSOC_NO_SPLIT_HEADS=1", "default": "",
            "deep_k": "VAEASSOC_NO_SMALL_ROWS=1 VAEASSOC_NO_SPLIT_HEADS=1", "small_rows": "VAEASSOC_NO_DEEP_K=1 VAEASSOC_NO_SPLIT_HEADS=1",
            "split_heads": "VAEASSOC_NO_DEEP_K=1 VAEASSOC_NO_SMALL_ROWS=1"}
if len(sys.argv) > 2:
    sys.path.insert(0, ".")
    from oracle import synth, philox, vae_assoc_oracle as vo
    from vae_assoc_b200 import vae_assoc
    B = int(sys.argv[1])
    archs = vo.reference_archs(4)
    model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct="relu", weights=[50, 1], assoc_lambda=8,
                                                  learning_rate=1e-3, batch_size=B, precision="tf32", seed=0, eps_seed=3)
    X = [x.astype(np.float32) for x in synth.synth_batch(archs, [True, False], 0, 1, 0, B)]
    eps = philox.eps_rows(3, 0, 0, B, 4).astype(np.float32)
    c = float(model.compute_gradients(X, eps))
    g = model.get_grads()
    costs = [float(model.partial_fit(X, eps)) for _ in range(4)]
    np.savez(sys.argv[2], c=c, costs=np.array(costs), **{"g%d" % i: x for i, x in enumerate(g)}, **{"p%d" % i: x for i, x in enumerate(model.get_params())})
    sys.exit(0)
for B in (100, 128, 256):
    res = {}
    for name, env in VARIANTS.items():
        out = "/tmp/var_%s_%d.npz" % (name, B)
        subprocess.check_call("env %s python scripts/debug_variants.py %d %s" % (env, B, out), shell=True, stderr=subprocess.DEVNULL)
        res[name] = np.load(out)
    ref = res["plain"]
    for name in VARIANTS:
        if name == "plain": continue
        r = res[name]
        gd = [float(np.linalg.norm(r["g%d" % i].astype(np.float64) - ref["g%d" % i]) / max(np.linalg.norm(ref["g%d" % i]), 1e-30)) for i in range(28)]
        pd = [float(np.linalg.norm(r["p%d" % i].astype(np.float64) - ref["p%d" % i]) / max(np.linalg.norm(ref["p%d" % i]), 1e-30)) for i in range(28)]
        print("B=%d %-11s cost %.3e | worst grad L2 %.2e (tensor %d) | worst param after 4 steps %.2e (tensor %d) | costs %s" % (
            B, name, abs(float(r["c"]) - float(ref["c"])) / abs(float(ref["c"])), max(gd), int(np.argmax(gd)), max(pd), int(np.argmax(pd)),
            np.array2string(np.abs(r["costs"] - ref["costs"]) / np.abs(ref["costs"]), precision=1)))

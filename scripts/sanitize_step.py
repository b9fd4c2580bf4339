"""One small train step through every kernel family of the tf32 schedule (one-launch form) -- the target of
scripts/sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import synth, vae_assoc_oracle as vo
from vae_assoc_b200 import vae_assoc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
archs = vo.reference_archs(4)
model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct="relu", weights=[50, 1], assoc_lambda=8,
                                              learning_rate=1e-3, batch_size=B, precision="tf32", seed=0, eps_seed=3,
                                              use_graph=False)
X = [x.astype(np.float32) for x in synth.synth_batch(archs, [True, False], 0, 1, 0, B)]
costs = [float(model.partial_fit(X)) for _ in range(2)]
print("costs", costs)
assert np.isfinite(costs).all()
model.close()

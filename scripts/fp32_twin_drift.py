"""How reproducible is the 1000-step run of tests/golden/latents_1k.npz for ANY fp32 implementation?

Runs the torch-CPU restatement (oracle/torch_twin.py) in fp32 with 8 threads (a different summation order inside the
matmuls than the frozen 1-thread fp32 run) on the same weights / data / eps and prints the error measure of
tests/test_gpu_latents_1k.py (max |dz| / max |z_ref| per modality, relative cost error) against
  (a) the frozen fp32 run (same graph, same precision, other summation order), and
  (b) the frozen fp64 oracle run.
    PYTHONPATH=. python scripts/fp32_twin_drift.py
"""
import numpy as np, torch
from oracle import make_golden_1k as g1k, torch_twin
torch.set_num_threads(8)
gold = np.load(g1k.OUT)
archs, params, data, eps = g1k.case()
m = torch_twin.TorchAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, g1k.BATCH, params, dtype=torch.float32)
probe = [torch.tensor(x, dtype=torch.float32) for x in g1k.batch_of(data, 0)]
rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
print(" step | vs frozen fp32 run: img jnt cost | vs fp64 oracle: img jnt cost")
for t in range(g1k.STEPS):
    X = [torch.tensor(x, dtype=torch.float32) for x in g1k.batch_of(data, t)]
    c = m.partial_fit(X, torch.tensor(eps(t), dtype=torch.float32))
    k = t + 1
    if k in g1k.CHECKPOINTS:
        with torch.no_grad():
            z = [m.encode(i, probe[i])[0].numpy().astype(np.float64) for i in range(2)]
        print("%5d | %.2e %.2e %.2e | %.2e %.2e %.2e" % (
            k, rel(z[0], gold["z32_img_%d" % k]), rel(z[1], gold["z32_jnt_%d" % k]), abs(c - gold["costs32"][t]) / gold["costs32"][t],
            rel(z[0], gold["z_img_%d" % k]), rel(z[1], gold["z_jnt_%d" % k]), abs(c - gold["costs"][t]) / gold["costs"][t]))

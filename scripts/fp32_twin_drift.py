"""How far does ANY fp32 implementation drift from the fp64 oracle over the 1000-step run of tests/golden/latents_1k.npz?
Runs the torch-CPU restatement (oracle/torch_twin.py) in fp32 on the same weights / data / eps and prints the same error
measure as tests/test_gpu_latents_1k.py (max |dz| / max |z_ref| per modality, relative cost error)."""
import numpy as np, torch
from oracle import make_golden_1k as g1k, torch_twin
torch.set_num_threads(8)
gold = np.load(g1k.OUT)
archs, params, data, eps = g1k.case()
m = torch_twin.TorchAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, g1k.BATCH, params, dtype=torch.float32)
probe = [torch.tensor(x, dtype=torch.float32) for x in g1k.batch_of(data, 0)]
for t in range(g1k.STEPS):
    X = [torch.tensor(x, dtype=torch.float32) for x in g1k.batch_of(data, t)]
    c = m.partial_fit(X, torch.tensor(eps(t), dtype=torch.float32))
    if t + 1 in g1k.CHECKPOINTS:
        with torch.no_grad():
            z = [m.encode(k, probe[k])[0].numpy().astype(np.float64) for k in range(2)]
        r = [float(np.abs(z[k] - gold["z_%s_%d" % (n, t + 1)]).max() / np.abs(gold["z_%s_%d" % (n, t + 1)]).max()) for k, n in enumerate(("img", "jnt"))]
        print("%5d  img %.2e  jnt %.2e  cost %.2e" % (t + 1, r[0], r[1], abs(float(c) - gold["costs"][t]) / gold["costs"][t]))

"""Launch the three tcgen05 GEMM forms at the reference shapes a few times (target of ncu captures)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_assoc_b200 import vae_assoc
archs = [dict(scope="image", hidden_conv=False, n_hidden_recog_1=8, n_hidden_recog_2=8, n_hidden_gener_1=8,
              n_hidden_gener_2=8, n_input=16, n_z=2)]
model = vae_assoc.AssocVariationalAutoEncoder(archs, batch_size=4, precision="fp32")
dev = model._dev
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
def t(r, c): return torch.randn(r, c, device=dev)
p = lambda x: C.c_void_p(x.data_ptr())
cases = [("NN", 0, B, 500, 500, t(B, 500), 500, t(500, 500), 500, t(B, 500), 500),
         ("NN784", 0, B, 784, 500, t(B, 500), 500, t(500, 784), 784, t(B, 784), 784),
         ("NNj", 0, B, 200, 147, t(B, 148), 148, t(147, 200), 200, t(B, 200), 200),
         ("NT", 1, B, 500, 500, t(B, 500), 500, t(500, 500), 500, t(B, 500), 500),
         ("TN", 2, 500, 500, B, t(B, 500), 500, t(B, 500), 500, t(500, 500), 500)]
for name, kind, M, N, K, A, lda, Bm, ldb, Cm, ldc in cases:
    for r in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bias = torch.randn(1024, device=dev) if kind == 0 else None       # forward form: bias + relu + tf32 rounding epilogue
        rc = model._lib.vaeassoc_debug_gemm(model._h, kind, 1, M, N, K, p(A), lda, p(Bm), ldb, p(Cm), ldc,
                                            p(bias) if kind == 0 else None, None, None, 0, 1 if kind == 0 else 0, 1 if kind == 0 else 0)
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
    print(name, M, N, K, "%.1f us incl. plan+launch+sync" % (1e3 * e0.elapsed_time(e1)), "%.1f TFLOP/s" % (2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12))
model.close()

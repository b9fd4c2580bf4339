"""cuBLAS tf32 / bf16 GEMM throughput on this GPU (library reference point for the roofline notes; not on the product path)."""
import json, sys, torch
torch.backends.cuda.matmul.allow_tf32 = True
out = {}
for name, dt, n in (("tf32", torch.float32, 8192), ("bf16", torch.bfloat16, 8192)):
    a = torch.randn(n, n, device="cuda", dtype=dt); b = torch.randn(n, n, device="cuda", dtype=dt)
    for _ in range(3): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops_burst"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
out["how"] = "torch.matmul 8192^3, best of 10, CUDA events; tf32 = fp32 tensors with allow_tf32"
print(json.dumps(out))

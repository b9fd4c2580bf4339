"""Structured probes of the tcgen05 GEMM kernels (debug helper; run under gpurun)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vae_assoc_b200 import vae_assoc

archs = [dict(scope="image", hidden_conv=False, n_hidden_recog_1=8, n_hidden_recog_2=8, n_hidden_gener_1=8,
              n_hidden_gener_2=8, n_input=16, n_z=2)]
model = vae_assoc.AssocVariationalAutoEncoder(archs, batch_size=4, precision="fp32")
dev = model._dev
NN, NT, TN = 0, 1, 2


def gemm(kind, A, B, M, N, K, Cinit=None):
    tA = torch.as_tensor(np.ascontiguousarray(A, np.float32)).to(dev)
    tB = torch.as_tensor(np.ascontiguousarray(B, np.float32)).to(dev)
    tC = torch.zeros((M, N), dtype=torch.float32, device=dev) if Cinit is None else torch.as_tensor(Cinit).to(dev)
    torch.cuda.synchronize()
    rc = model._lib.vaeassoc_debug_gemm(model._h, kind, 1, M, N, K, C.c_void_p(tA.data_ptr()), A.shape[1],
                                        C.c_void_p(tB.data_ptr()), B.shape[1], C.c_void_p(tC.data_ptr()), N, None, None,
                                        None, 0, 0, 0)
    assert rc == 0, model._lib.vaeassoc_last_error(model._h).decode()
    return tC.cpu().numpy()


def logical(kind, A, B):
    if kind == NN:
        return A.astype(np.float64) @ B.astype(np.float64)
    if kind == NT:
        return A.astype(np.float64) @ B.astype(np.float64).T
    return A.astype(np.float64).T @ B.astype(np.float64)


def shapes(kind, M, N, K):
    return {NN: ((M, K), (K, N)), NT: ((M, K), (N, K)), TN: ((K, M), (K, N))}[kind]


np.set_printoptions(linewidth=200, precision=3, suppress=True)
for kind, name in ((NT, "NT"), (NN, "NN"), (TN, "TN")):
    for (M, N, K) in ((128, 128, 32), (128, 128, 64), (256, 256, 96), (256, 160, 96)):
        sa, sb = shapes(kind, M, N, K)
        rng = np.random.RandomState(0)
        A = rng.randint(-3, 4, size=sa).astype(np.float32)
        B = rng.randint(-3, 4, size=sb).astype(np.float32)
        got = gemm(kind, A, B, M, N, K)
        ref = logical(kind, A, B)
        bad = np.abs(got - ref) > 1e-3
        print("%s M=%d N=%d K=%d: mismatches %d / %d, |got|max %.3g |ref|max %.3g" % (name, M, N, K, bad.sum(), bad.size,
                                                                                   np.abs(got).max(), np.abs(ref).max()))
        if bad.any() and (M, N, K) == (128, 128, 32):
            print("  rows with errors:", np.where(bad.any(1))[0][:40])
            print("  cols with errors:", np.where(bad.any(0))[0][:40])
            print("  got[0,:16]", got[0, :16]); print("  ref[0,:16]", ref[0, :16])
            print("  got[1,:16]", got[1, :16]); print("  ref[1,:16]", ref[1, :16])
            # which single k contributes where: A selects k*, B[k, n] = 100*k + n (NN/TN), B[n, k] for NT
            for ks in (0, 1, 4, 8, 31):
                A2 = np.zeros(sa, np.float32); B2 = np.zeros(sb, np.float32)
                if kind == TN:
                    A2[ks, :] = 1
                else:
                    A2[:, ks] = 1
                kk, nn = np.meshgrid(np.arange(K), np.arange(N), indexing="ij")
                B2[:] = (100 * kk + nn) if kind != NT else (100 * kk + nn).T
                g2 = gemm(kind, A2, B2, M, N, K)
                print("   k*=%d got[0,:10]" % ks, g2[0, :10], " got[0,32:36]", g2[0, 32:36], "want", 100 * ks, "+n")
            # A structure: B selects n*=0 only via ones column; A[m,k] = 100*m + k
            mm, kk = np.meshgrid(np.arange(M), np.arange(K), indexing="ij")
            A3 = (mm + 0.0 * kk).astype(np.float32)
            A3 = A3 if kind != TN else A3.T.copy()
            B3 = np.zeros(sb, np.float32)
            if kind == NT:
                B3[:, 0] = 1
            else:
                B3[0, :] = 1
            g3 = gemm(kind, A3, B3, M, N, K)
            print("   A[m,k]=m, only k=0 of B: got[:12,0]", g3[:12, 0], " got[32:36,0]", g3[32:36, 0], " got[64:68,0]", g3[64:68, 0])
model.close()

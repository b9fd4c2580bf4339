"""Data parallelism (new in this build; the reference is single-process).

CPU (gloo, world_size 2): the host-side DP logic -- row sharding, the 1/B_global scaling rule of SURVEY.md 8e, the
global-row Philox addressing of eps and of the synthetic stream, all-reduce(SUM) of flat gradients + cost slot --
with the ORACLE as the compute (test infrastructure), against the single-process oracle.
GPU (NCCL, needs >= 2 B200s: `gpurun --gpus 2 -- pytest -m gpu tests/test_dp.py`): the CUDA path with its in-library
NCCL all-reduce against the single-GPU CUDA path at the same global batch."""
import os
import socket

import numpy as np
import pytest

from oracle import make_golden, philox, synth
from oracle import vae_assoc_oracle as vo


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _gloo_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    archs = make_golden.tiny_archs()
    Bg, B = 8, 8 // world
    row0 = rank * B
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, B, seed=4)
    costs = []
    for t in range(3):
        X = synth.synth_batch(archs, [True, False], 7, 1, t * Bg + row0, B)      # this rank's rows of the global stream
        eps = philox.eps_rows(9, t, row0, B, archs[0]["n_z"])                    # addressed by GLOBAL row
        c, g, _ = o.loss_and_grads(X, eps, global_batch=Bg)
        flat = torch.tensor(np.concatenate([x.ravel() for gs in g for x in gs] + [np.array([c])]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)                               # grads + cost slot in one buffer
        flat = flat.numpy()
        costs.append(flat[-1])
        k, gsum = 0, []
        for gs in g:
            row = []
            for x in gs:
                row.append(flat[k:k + x.size].reshape(x.shape)); k += x.size
            gsum.append(row)
        o.adam_step(gsum)
    if rank == 0:
        np.savez(out, costs=np.array(costs), **{"p%d" % i: p for i, p in enumerate(p for ps in o.params for p in ps)})
    dist.barrier()
    dist.destroy_process_group()


def test_dp_logic_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "dp.npz")
    mp.spawn(_gloo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    archs = make_golden.tiny_archs()
    o = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, 8, seed=4)
    costs = []
    for t in range(3):
        X = synth.synth_batch(archs, [True, False], 7, 1, t * 8, 8)
        costs.append(o.partial_fit(X, philox.eps_rows(9, t, 0, 8, archs[0]["n_z"])))
    np.testing.assert_allclose(got["costs"], costs, rtol=1e-12)
    for i, p in enumerate(p for ps in o.params for p in ps):
        np.testing.assert_allclose(got["p%d" % i], p, rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, out, precision, mode="nccl"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vae_assoc_b200 import vae_assoc
    archs = vo.reference_archs(4)
    Bg = 256
    B = Bg // world
    model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct="relu", weights=[50, 1],
                                                  assoc_lambda=8, learning_rate=1e-3, batch_size=B, precision=precision,
                                                  seed=0, eps_seed=5, global_batch=Bg, global_row0=rank * B)
    # peer: cudaIpc mapping, plain peer loads / stores; peer-nvls: symmetric memory + multimem through the NVSwitch
    os.environ["VAEASSOC_DP_SYMMETRIC"] = "1" if mode == "peer-nvls" else "0"
    model.init_data_parallel(peer=(mode != "nccl"))
    if mode == "peer-nvls" and model.dp_mode != "peer-nvls":
        # no symmetric memory / multicast on this box: the library must have fallen back consistently on every rank
        assert model.dp_mode in ("peer", "nccl")
        print("[dp] NVLS form unavailable here: %s" % getattr(model, "_peer_error", None))
    else:
        assert model.dp_mode == mode, (model.dp_mode, getattr(model, "_peer_error", None))
    costs = []
    for t in range(4):
        xs = model.synth_batch(t * Bg + rank * B, B)
        costs.append(float(model.partial_fit(xs)))          # Philox eps addressed by global row and step
    params = model.get_params()
    m, v, step = model.get_adam_state()                     # peer mode: pulls the shards owned by the other ranks
    if rank == world - 1:                                   # the LAST rank: its copy of every other shard came from a peer
        np.savez(out, costs=np.array(costs), step=step, **{"p%d" % i: p for i, p in enumerate(params)},
                 **{"m%d" % i: x for i, x in enumerate(m)}, **{"v%d" % i: x for i, x in enumerate(v)})
    dist.barrier()
    model.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["nccl", "peer", "peer-nvls"])
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_dp_matches_single_gpu(tmp_path, precision, mode):
    """nccl: ncclAllReduce of the flat gradients + replicated Adam; peer: the one-kernel reduce-scatter + sharded Adam +
    all-gather over NVLink peer memory (csrc/peer_adam.cu) through cudaIpc mappings; peer-nvls: the same kernel over
    symmetric memory with the reduction done inside the NVSwitch (multimem.ld_reduce / multimem.st).  All must reproduce
    the single-GPU run."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from vae_assoc_b200 import build, vae_assoc
    build.build(verbose=False)
    out = str(tmp_path / "dp.npz")
    mp.spawn(_nccl_worker, args=(2, _free_port(), out, precision, mode), nprocs=2, join=True)
    got = np.load(out)
    archs = vo.reference_archs(4)
    model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct="relu", weights=[50, 1],
                                                  assoc_lambda=8, learning_rate=1e-3, batch_size=256, precision=precision,
                                                  seed=0, eps_seed=5)
    costs = [float(model.partial_fit(model.synth_batch(t * 256, 256))) for t in range(4)]
    # identical up to the summation order of the batch reduction (2 shards + all-reduce vs one split-K kernel)
    np.testing.assert_allclose(got["costs"], costs, rtol=2e-5)
    for i, p in enumerate(model.get_params()):
        d = np.linalg.norm(got["p%d" % i].astype(np.float64) - p) / max(np.linalg.norm(p), 1e-30)
        # tf32 + relu: the shards' bias-gradient atomics / split-K order differ from the one-GPU run and Adam amplifies the
        # last-bit differences of single entries whose gradient passes near zero in one of the steps (see rel_l2 in
        # tests/test_gpu_parity.py): the whole L2 difference of a 200-entry bias vector sits in one to five entries that
        # moved by 1e-4 of lr-sized values (scripts/debug_variants.py: two schedules on ONE GPU differ by 2.9e-3 on such
        # a tensor after 4 steps with first-step gradients equal to 2e-5).  Measured here: 1.9e-4 .. 1.3e-3.  The 2-rank
        # NCCL schedule is in addition the four-launch form with the stand-alone loss kernels (another form of d a).
        tol = 1e-4 if precision == "fp32" else 3e-3
        assert d < tol, (i, d)
    m1, v1, step1 = model.get_adam_state()
    assert int(got["step"]) == int(step1) == 4
    for i in range(len(m1)):
        for name, ref in (("m", m1[i]), ("v", v1[i])):
            # L2-relative like the parameters: with tf32 + relu a flipped mask bit moves single entries (see rel_l2 in
            # test_gpu_parity.py); measured 3.7e-3 in the max norm on a bias slot with EITHER schedule
            d = np.linalg.norm(got["%s%d" % (name, i)].astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-30)
            assert d < (1e-4 if precision == "fp32" else 5e-3), (name, i, d)
    model.close()

#!/usr/bin/env python
"""Benchmark of the associated-VAE train step (BASELINE.json metric: paired samples/sec per train step).

    python bench.py --gpus N --steps K --warmup W                      # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU restatement of the reference graph

A "step" is one `partial_fit` of the reference model (vae_assoc.py:378-386) on one synthetic paired batch:
forward of both encoders/decoders, losses, backward, Adam.  Workload = BASELINE.json configs[1]/[2]:
reference architecture (image 784-500-500, joint 147-200-200, n_z 4, relu, weights [50,1], lambda 8, lr 1e-3),
8192 pairs per GPU (global batch 8192*N: 65536 at N=8), weak scaling.  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 7744400.0      # fwd 2 862 400 + bwd 4 882 000 (2*K*N per GEMM row)
N_PARAMS = 1434947


def ref_archs(n_z=4):
    img = dict(scope="image", hidden_conv=False, n_hidden_recog_1=500, n_hidden_recog_2=500, n_hidden_gener_1=500,
               n_hidden_gener_2=500, n_input=784, n_z=n_z)
    jnt = dict(scope="joint", hidden_conv=False, n_hidden_recog_1=200, n_hidden_recog_2=200, n_hidden_gener_1=200,
               n_hidden_gener_2=200, n_input=147, n_z=n_z)
    return [img, jnt]


def conv_archs(n_z=4):
    """BASELINE.json configs[3]: the commented-out conv image arch of vae_assoc_ujichar_img_jnt.py:72-80 (recog 16/64,
    gener 64/16, hidden_conv=True: conv encoder + deconv.py decoder) next to the dense joint arch."""
    img = dict(scope="image", hidden_conv=True, n_hidden_recog_1=16, n_hidden_recog_2=64, n_hidden_gener_1=64,
               n_hidden_gener_2=16, n_input=784, n_z=n_z)
    return [img, ref_archs(n_z)[1]]


def scaled_archs(n_z=64, width=2048):
    """BASELINE.json configs[4]: 4 x 2048 hidden layers per modality (2 encoder + 2 decoder layers through the
    reference's own arch dict), latent 64."""
    a = ref_archs(n_z)
    for na in a:
        for k in ("n_hidden_recog_1", "n_hidden_recog_2", "n_hidden_gener_1", "n_hidden_gener_2"):
            na[k] = width
    return a


CONFIGS = {
    # name: (archs builder, default pairs per GPU, workload description)
    "ref": (ref_archs, 8192, "assoc-VAE reference arch (img 784-500-500, jnt 147-200-200, n_z 4, relu, w [50,1], lambda 8, lr 1e-3)"),
    "conv": (conv_archs, 4096, "assoc-VAE with the hidden_conv=True image modality (conv encoder 16/32/64, deconv.py decoder "
                               "64/32/16/1 + dense 784x784) and the dense joint modality, n_z 4, relu, w [50,1], lambda 8"),
    "scaled": (scaled_archs, 16384, "scaled assoc-VAE (both modalities 2048-2048 hidden, n_z 64, relu, w [50,1], lambda 8)"),
}


def dense_flops_per_sample(archs):
    """2*K*N per GEMM row; forward + wgrad of every layer + dgrad of every layer but the two input layers."""
    tot = 0
    for na in archs:
        ni, r1, r2, nz = na["n_input"], na["n_hidden_recog_1"], na["n_hidden_recog_2"], na["n_z"]
        fwd = 2 * (ni * r1 + r1 * r2 + r2 * 2 * nz + nz * r1 + r1 * r2 + r2 * ni)
        tot += 3 * fwd - 2 * ni * r1
    return tot


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, src="fallback")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def cpu_reference_run(batch, steps, warmup, threads=None, archs=None):
    """CPU restatement (torch fp32, all host threads) of the reference graph at its op granularity; NOT TensorFlow
    (not installable here).  Returns (samples_per_s, ms_per_step, cores)."""
    import torch
    from oracle import synth, torch_twin
    from oracle import vae_assoc_oracle as vo
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    archs = archs or vo.reference_archs(4)
    params = vo.init_params(archs, 0)
    model = torch_twin.TorchAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, batch, params, dtype=torch.float32)
    X = [torch.tensor(x, dtype=torch.float32) for x in synth.synth_batch(archs, [True, False], 0, 1, 0, batch)]
    for _ in range(warmup):
        model.partial_fit(X)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.partial_fit(X)          # float(cost): one host read per step, as partial_fit does (vae_assoc.py:383-386)
    dt = time.perf_counter() - t0
    return batch * steps / dt, 1e3 * dt / steps, cores


def pool_size(bytes_per_batch):
    """distinct device-resident batches the timed loop rotates through: > 2 x the 126 MB L2"""
    return min(max(4, int(np.ceil(2.2 * 126e6 / bytes_per_batch))), 512)


def make_config(args, archs, desc, world):
    """`config` of the JSON line -- the SAME dict for this repo's arm and for `--impl reference` (the driver compares them)."""
    B = args.batch
    bytes_per_batch = B * sum(na["n_input"] for na in archs) * 4
    n = pool_size(bytes_per_batch)
    return dict(workload="%s, %d pairs per GPU per step" % (desc, B), name=args.config, global_batch=B * world,
                per_gpu_batch=B, parallelism="dp%d" % world,
                l2="GPU arm: inputs rotate through a device pool of %d distinct batches (%.0f MB > 126 MB L2)"
                   % (n, n * bytes_per_batch / 1e6))


def committed_profile(name):
    """numbers that come from committed profiler captures, never literals in this file"""
    p = os.path.join(ROOT, "profiles", name)
    return json.load(open(p)) if os.path.exists(p) else {}


def parity_check(archs, B, precision, dynamic_first):
    """Outside every timed region: ONE compute_gradients of the benchmarked path (same architecture, same per-GPU batch,
    same precision, fused tile segments) against the CPU oracle -- the operand-rounding oracle for tf32, the exact one
    for fp32 -- on a synthetic batch with injected Philox eps.  Returns {cost_rel, worst_grad_l2, worst_grad_max, ...}."""
    from oracle import philox
    from oracle import vae_assoc_oracle as vo
    from vae_assoc_b200 import vae_assoc
    if dynamic_first:
        os.environ["VAEASSOC_DYNAMIC_FIRST"] = "1"      # the task-queue mode data-parallel runs use
    try:
        model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=vae_assoc.relu, weights=[50, 1],
                                                      assoc_lambda=8, learning_rate=1e-3, batch_size=B, precision=precision, seed=0)
    finally:
        os.environ.pop("VAEASSOC_DYNAMIC_FIRST", None)
    params = model.get_params()
    rng = np.random.RandomState(5)
    params = [p if p.ndim > 1 else (0.05 * rng.normal(size=p.shape)).astype(np.float32) for p in params]
    model.set_params(params)
    per_mod, k = [], 0
    for na in archs:
        n = len(vo.param_names(na))
        per_mod.append([p.astype(np.float64) for p in params[k:k + n]]); k += n
    oracle = vo.OracleAssocVAE(archs, [True, False], "relu", [50.0, 1.0], 8.0, 1e-3, B, params=per_mod,
                               emulate_tf32=(precision == "tf32"))
    xs = model.synth_batch(0, B)
    X = [x.cpu().numpy() for x in xs]
    eps = philox.eps_rows(0, 0, 0, B, archs[0]["n_z"]).astype(np.float32)
    t0 = time.perf_counter()
    cost = float(model.compute_gradients(xs, eps))
    c_ref, g_ref, _ = oracle.loss_and_grads(X, eps)
    l2 = mx = 0.0
    worst = None
    for g, r, n in zip(model.get_grads(), [g for gs in g_ref for g in gs], model.variable_roles()):
        r = np.asarray(r, np.float64); g = np.asarray(g, np.float64)
        e2 = float(np.linalg.norm(g - r) / max(np.linalg.norm(r), 1e-30))
        em = float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-30))
        if e2 > l2:
            l2, worst = e2, "%d/%s" % n
        mx = max(mx, em)
    fused = model.launch_count()
    model.close()
    return dict(cost_rel=abs(cost - c_ref) / abs(c_ref), worst_grad_l2=l2, worst_grad_max=mx, worst_tensor=worst,
                oracle="fp64 numpy restatement, operands rounded to tf32 where the CUDA path rounds" if precision == "tf32"
                else "fp64 numpy restatement (exact)",
                bounds=dict(cost_rel=5e-4, worst_grad_l2=5e-4), batch=B, precision=precision,
                seconds=round(time.perf_counter() - t0, 2))


def timed_steps(model, pool, steps, warmup, barrier=None):
    """K steps from device-resident batches, CUDA events on the library's stream; returns ms per step."""
    import torch
    n = len(pool)
    for k in range(warmup):
        model.partial_fit_async(pool[k % n])
    (barrier or torch.cuda.synchronize)()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(steps):
        model.partial_fit_async(pool[(warmup + k) % n])
    ev1.record()
    (barrier or torch.cuda.synchronize)()
    return ev0.elapsed_time(ev1) / steps


def secondary_runs(local_rank):
    """The other BASELINE.json configurations, short (N = 1 only): reference arch at the reference's own batch 100, the fp32
    (1e-4) path at 8192, the hidden_conv modality at 4096 (configs[3]) and the scaled model at 16 384 (configs[4], one
    GPU's share).  Same timing rules as the headline (device events, inputs rotating through a pool larger than L2)."""
    import torch
    from vae_assoc_b200 import vae_assoc
    out = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    for name, cfg, B, precision, steps in (("ref_b100", "ref", 100, "tf32", 400), ("ref_fp32_b8192", "ref", 8192, "fp32", 12),
                                           ("conv_b4096", "conv", 4096, "tf32", 20), ("scaled_b16384", "scaled", 16384, "tf32", 12)):
        mk, _, desc = CONFIGS[cfg]
        archs = mk()
        try:
            model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=vae_assoc.relu, weights=[50, 1],
                                                          assoc_lambda=8, learning_rate=1e-3, batch_size=B, precision=precision, seed=0)
            n = pool_size(B * sum(na["n_input"] for na in archs) * 4)
            if B <= 256:
                n = min(n, 64)         # 100-pair batches are 0.37 MB: a 64-batch pool is cycled; the 5.7 MB of weights dominate anyway
            pool = [model.synth_batch(k * B, B) for k in range(n)]
            torch.cuda.synchronize()
            l0 = model.launch_count()
            ms = timed_steps(model, pool, steps, 5)
            launches = (model.launch_count() - l0) / float(steps + 5)
            flops = dense_flops_per_sample(archs) * B if cfg != "conv" else None
            out.append(dict(name=name, config=cfg, per_gpu_batch=B, dtype=precision, steps=steps, warmup=5, ms_per_step=ms,
                            value=B / (ms * 1e-3), unit="samples/s", launches_per_step=launches,
                            step_tflops=(flops / (ms * 1e-3) / 1e12) if flops else None, last_cost=float(model.last_cost())))
            model.close()
            del pool
            torch.cuda.empty_cache()
        except Exception as e:      # a secondary line must never take the headline down
            out.append(dict(name=name, error=str(e)[:300]))
    out.append(inference_run())
    return out, sampler.stop()


def inference_run(B=100, calls=200):
    """The inference surface at the callers' batch (baxter_vae_assoc_writer.py:142-145 pads to a full batch of 100):
    wall-clock per call of transform / generate / reconstruct with HOST arrays in and out -- vaeassoc_infer_host: one
    graph launch + one D2H per call -- next to the per-modality device entry points + .cpu() copies (round 1)."""
    import torch
    from vae_assoc_b200 import vae_assoc
    try:
        archs = ref_archs()
        model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=vae_assoc.relu, weights=[50, 1],
                                                      assoc_lambda=8, learning_rate=1e-3, batch_size=B, precision="tf32", seed=0)
        X = [x.cpu().numpy() for x in model.synth_batch(0, B)]
        Xd = [torch.as_tensor(x).cuda() for x in X]
        z = np.random.RandomState(0).normal(size=(B, archs[0]["n_z"])).astype(np.float32)
        zd = torch.as_tensor(z).cuda()
        res = dict(name="inference_b%d" % B, per_gpu_batch=B, dtype="tf32", calls=calls, unit="ms per call (wall clock, host arrays out)")
        for label, fn_host, fn_dev in (("transform", lambda: model.transform(X), lambda: model.transform(Xd)),
                                       ("generate", lambda: model.generate(z), lambda: model.generate(zd)),
                                       ("reconstruct", lambda: model.reconstruct(X), lambda: model.reconstruct(Xd))):
            for tag, fn in (("", fn_host), ("_device_entry_points", fn_dev)):
                for _ in range(10):
                    fn()
                model.synchronize()
                t0 = time.perf_counter()
                l0 = model.launch_count()
                for _ in range(calls):
                    fn()
                model.synchronize()
                res[label + tag + "_ms"] = (time.perf_counter() - t0) * 1e3 / calls
                res[label + tag + "_launches"] = (model.launch_count() - l0) / float(calls)
        model.close()
        return res
    except Exception as e:
        return dict(name="inference_b%d" % B, error=str(e)[:300])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="ref", choices=sorted(CONFIGS), help="ref = BASELINE configs[1]/[2] (the metric's "
                    "workload), conv = configs[3], scaled = configs[4]")
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU per step (default: the config's)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the benchmarked path")
    ap.add_argument("--quick", action="store_true", help="diagnostics: only the HBM-resident timing loop")
    args = ap.parse_args()

    # stdout carries exactly ONE JSON line: everything else a library prints there (NCCL's "NCCL version ..." banner comes
    # from C code) is sent to stderr by pointing file descriptor 1 at stderr; the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    mk_archs, default_batch, desc = CONFIGS[args.config]
    if args.batch <= 0:
        args.batch = default_batch
    archs = mk_archs()

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU implementation of the path (torch-CPU restatement: TensorFlow is not installable), all host
        # threads, EXACTLY --steps timed and --warmup untimed steps; each step = a bounded sample of the arm's per-GPU batch
        n_gpus = max(args.gpus, world)
        cpu_batch = min(args.batch, 8192 if args.config == "ref" else 2048)
        sps, ms, cores = cpu_reference_run(cpu_batch, args.steps, args.warmup, archs=archs)
        line = dict(impl="reference", metric="paired samples/sec/train step", value=sps, unit="samples/s",
                    n_gpus=n_gpus, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="fp32", data="synthetic",
                    config=make_config(args, archs, desc, n_gpus),
                    reference_note="CPU restatement (torch fp32) of the reference graph, not TensorFlow (not installable here); "
                                   "one process on rank 0, host cores only",
                    cpu_baseline=dict(value=sps, unit="samples/s", cores=cores, kind="port",
                                      sample="each step = %d pairs (bounded sample of the %d-pair per-GPU batch), %d timed "
                                             "steps after %d warm-up" % (cpu_batch, args.batch, args.steps, args.warmup)),
                    e2e=dict(value=sps, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        json_out.write(json.dumps(line) + "\n"); json_out.flush()
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from vae_assoc_b200 import vae_assoc

    B = args.batch
    Bg = B * world
    model = vae_assoc.AssocVariationalAutoEncoder(archs, [True, False], transfer_fct=vae_assoc.relu,
                                                  weights=[50, 1], assoc_lambda=8, learning_rate=1e-3, batch_size=B,
                                                  precision=args.precision, seed=0, use_graph=not args.no_graph,
                                                  global_batch=Bg, global_row0=rank * B)
    if world > 1:
        model.init_data_parallel()

    # device-resident pool of distinct synthetic batches, larger than L2 (126 MB) so that no step re-reads a cached input
    bytes_per_batch = B * sum(na["n_input"] for na in archs) * 4
    pool_n = pool_size(bytes_per_batch)
    pool = [model.synth_batch((k * world + rank) * B, B) for k in range(pool_n)]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-resident number: inputs already in HBM ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # nvidia-smi needs ~1 s to produce its first line: start before the warm-up
    for k in range(args.warmup):
        model.partial_fit_async(pool[k % pool_n])
    if rank == 0:
        # BEFORE the barrier: a rank that waited here after it would start its timed loop late and every other rank's
        # first step (a collective) would wait for it inside THEIR timed region
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 3.0:
            time.sleep(0.05)
    barrier()
    if rank == 0:
        sampler.lines.clear()           # keep only samples taken during the timed region
    l0 = model.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        model.partial_fit_async(pool[(args.warmup + k) % pool_n])
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = model.launch_count() - l0
    if ms_total < 400.0:
        # keep the same work running until the 100 ms clock sampler has seen it (these steps are not timed).  EVERY rank
        # runs the same number of extra steps (the count derives from the rank-reduced time): under data parallelism a
        # step contains all-reduces, and a rank that stepped alone would wait for its peers forever
        n_extra = int(min(5000, max(20, 600.0 / max(ms_total / args.steps, 1e-3))))
        for k in range(n_extra):
            model.partial_fit_async(pool[k % pool_n])
        model.synchronize()
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    last_cost = float(model.last_cost())
    ms_step = ms_total / args.steps
    value = Bg * args.steps / (ms_total * 1e-3)

    if args.quick:
        if rank == 0:
            json_out.write(json.dumps(dict(value=value, ms_per_step=ms_step, n_gpus=world, quick=True)) + "\n"); json_out.flush()
        model.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end: host batches -> H2D -> step -> D2H of the cost, every step ----------------------------------
    def e2e_run(host_pool, refill_guard):
        n = len(host_pool)
        tickets = [None] * n
        for k in range(min(args.warmup, 5)):
            tickets[k % n] = model.partial_fit_async(host_pool[k % n])
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for k in range(args.steps):
            if refill_guard and tickets[k % n] is not None:
                model.wait_uploaded(tickets[k % n])      # a pinned slot is handed out again only after its DMA has completed
            tickets[k % n] = model.partial_fit_async(host_pool[k % n])
        ev1.record()
        model.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        ms = max_over_ranks(max(ev0.elapsed_time(ev1), wall * 1e3))
        return Bg * args.steps / (ms * 1e-3), ms / args.steps

    host_n = 4
    pinned_pool = [[torch.empty(x.shape, dtype=torch.float32, pin_memory=True).copy_(x).numpy() for x in pool[k]] for k in range(host_n)]
    torch.cuda.synchronize()
    e2e_value, e2e_ms = e2e_run(pinned_pool, True)
    # the reference's callers hand partial_fit fresh PAGEABLE numpy arrays (vae_assoc.py:541-543): same loop, pageable memory
    pageable_pool = [[np.array(a, copy=True) for a in hp] for hp in pinned_pool]
    pg_value, pg_ms = e2e_run(pageable_pool, False)
    # device-resident data set (train(device_data=True)): the pool's rows uploaded once, batches gathered by index
    width = sum(na["n_input"] for na in archs)
    rows = np.concatenate([np.concatenate(hp, axis=1) for hp in pinned_pool], axis=0)
    data_dev = model.upload_dataset(rows)
    idx_rng = np.random.RandomState(rank)
    idx_pool = [np.ascontiguousarray(idx_rng.permutation(rows.shape[0])[:B].astype(np.int64)) for _ in range(8)]
    for k in range(5):
        model.partial_fit_indexed(data_dev, idx_pool[k % 8])
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        model.partial_fit_indexed(data_dev, idx_pool[k % 8])
    ev1.record()
    model.synchronize()
    barrier()
    dd_ms = max_over_ranks(max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3))
    dd_value = Bg * args.steps / (dd_ms * 1e-3)
    del data_dev

    # ---- per-kernel roofline: one eager step with CUDA events between ops, averaged ------------------------------
    import ctypes as C
    cap = 128
    names = C.create_string_buffer(32 * cap)
    ms = (C.c_float * cap)(); fl = (C.c_double * cap)(); by = (C.c_double * cap)()
    acc = {}
    reps = 5
    for r in range(reps + 1):
        ptrs, lds, keep = model._dev_args(pool[r % pool_n])
        n = model._lib.vaeassoc_profile_step(model._h, ptrs, lds, None, names, ms, fl, by, cap)
        assert n > 0, model._lib.vaeassoc_last_error(model._h)
        if r == 0:
            continue
        for i in range(n):
            nm = names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()
            a = acc.setdefault(nm, [0.0, fl[i], by[i]])
            a[0] += ms[i] / reps
    pk = peaks()
    kernels = []
    for nm, (t, f, b) in acc.items():
        k = dict(name=nm, ms=round(t, 5))
        if f > 0:
            k.update(bound="tensor", achieved=f / (t * 1e-3) / 1e12, unit="TFLOP/s")
            k["frac"] = k["achieved"] / pk["tensor"]
        elif b > 0:
            k.update(bound="hbm", achieved=b / (t * 1e-3) / 1e9, unit="GB/s")
            k["frac"] = k["achieved"] / pk["hbm"]
        kernels.append(k)
    kernels.sort(key=lambda k: -k["ms"])
    top = kernels[0]
    eager_ms = sum(k["ms"] for k in kernels)
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of THIS workload
    # (profiles/ncu_traffic.json names the capture it was read from); other workloads / kernels: not captured -> null
    prof = committed_profile("ncu_traffic.json")
    key = "%s_b%d_%s" % (args.config, B, args.precision)
    traffic = prof.get(key, {}).get("dram_bytes_per_launch", {}).get(top["name"])
    tf32_ref = committed_profile("r1_tf32_peak.json").get("tf32_tflops_burst")
    roofline = dict(bound=top.get("bound", "hbm"), kernel=top["name"], achieved=top.get("achieved"),
                    peak=pk["tensor"] if top.get("bound") == "tensor" else pk["hbm"], unit=top.get("unit"),
                    frac=top.get("frac"), traffic=traffic, traffic_unit="bytes per launch (DRAM read + write, ncu)",
                    traffic_source=prof.get(key, {}).get("source"),
                    peak_source=pk["src"],
                    share_of_step=top["ms"] / eager_ms,
                    note="peak = measured cuBLAS bf16 (MEASURED_PEAKS.json); this path computes in kind::tf32, whose nominal peak "
                         "is half of bf16; measured cuBLAS tf32 8192^3 on this pool: profiles/r1_tf32_peak.json"
                    if top.get("bound") == "tensor" else "")
    if top.get("bound") == "tensor" and top.get("achieved") and tf32_ref:
        roofline["frac_of_tf32_cublas"] = top["achieved"] / tf32_ref
    if top["name"] == "seg_step":
        roofline["note"] += ("; seg_step = ONE launch of the persistent tile kernel carrying the whole gradient step (all 34 "
                             "contractions, latent stage, reconstruction losses, cost): its fraction is the step-level figure, "
                             "not the best segment's")
    # algorithmic FLOPs of one step = sum over the contractions of the schedule (2*M*N*K each); for the dense configs this
    # is SURVEY 8d's per-sample figure x B (7 744 400 x B at the reference arch)
    step_flops = sum(f for (_, f, _) in acc.values())
    if args.config != "conv":
        assert abs(step_flops - dense_flops_per_sample(archs) * B) <= 1e-6 * step_flops, (step_flops, dense_flops_per_sample(archs) * B)
    step_tflops = step_flops / (ms_step * 1e-3) / 1e12

    cfg = make_config(args, archs, desc, world)
    line = dict(metric="paired samples/sec/train step", value=value, unit="samples/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.precision, data="synthetic", config=cfg, graph=not args.no_graph,
                clocks=clocks, gpu_launches=int(launches), dp_mode=model.dp_mode,
                e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=bytes_per_batch, d2h_bytes_per_step=4,
                         ms_per_step=e2e_ms,
                         path="AssocVariationalAutoEncoder.partial_fit_async(numpy pinned) -> vaeassoc_submit_host"),
                e2e_pageable=dict(value=pg_value, unit="samples/s", h2d_bytes_per_step=bytes_per_batch, d2h_bytes_per_step=4,
                                  ms_per_step=pg_ms, path="the same call with pageable numpy arrays (what the reference's callers pass)"),
                e2e_device_dataset=dict(value=dd_value, unit="samples/s", h2d_bytes_per_step=8 * B, d2h_bytes_per_step=4,
                                        ms_per_step=dd_ms / args.steps,
                                        path="train(device_data=True): partial_fit_indexed -> vaeassoc_submit_indexed (data set "
                                             "uploaded once, batches gathered on the device from host row indices)"),
                roofline=roofline, step_tflops_per_gpu=step_tflops, last_cost=last_cost,
                kernels=kernels[:40])
    model.close()
    del pool
    torch.cuda.empty_cache()

    if rank == 0 and not args.no_parity:
        try:
            line["parity"] = parity_check(archs, B, args.precision, dynamic_first=world > 1)
        except Exception as e:
            line["parity"] = dict(error=str(e)[:300])
    if rank == 0 and world == 1 and not args.no_secondary and args.config == "ref":
        line["secondary"], line["secondary_clocks"] = secondary_runs(local_rank)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_batch = min(B, 8192 if args.config == "ref" else 2048)
        sps, cms, cores = cpu_reference_run(cpu_batch, 5, 1, archs=archs)
        line["cpu_baseline"] = dict(value=sps, unit="samples/s", cores=cores, kind="port", ms_per_step=cms,
                                    sample="5 steps of %d pairs after 1 warm-up, torch-CPU fp32 restatement "
                                           "(not TensorFlow)" % cpu_batch)
    if world > 1:
        dist.barrier()
    if rank == 0:
        json_out.write(json.dumps(line) + "\n"); json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

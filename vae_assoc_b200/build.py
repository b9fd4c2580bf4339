"""Builds libvaeassoc.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m vae_assoc_b200.build [--force]

One translation unit per .cu file (compiled in parallel), then one link.  The result,
vae_assoc_b200/libvaeassoc.so, is git-ignored but travels with the repo snapshot to the GPU box.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libvaeassoc.so")
SOURCES = ["api.cu", "gemm_simt.cu", "gemm_skinny.cu", "gemm_group.cu", "conv.cu", "loss.cu", "adam.cu", "peer_adam.cu", "synth.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libvaeassoc cannot be built")
    return nvcc


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "vaeassoc.h"))
    return hdrs


def _compile(src, log):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    dig = _digest([os.path.join(CSRC, src)] + _deps())
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, src + ".log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
    with open(stamp, "w") as f:
        f.write(dig)
    if log:
        print("[build] compiled", src)
    return obj, True


def build_variant(name, defines, verbose=True):
    """A build-time variant of the library next to the default one (e.g. `w16`: 16 epilogue warps per CTA in the tile
    kernel): vae_assoc_b200/libvaeassoc_<name>.so, selected at run time with VAEASSOC_LIB=<path>."""
    obj_dir = OBJ + "_" + name
    lib = os.path.join(HERE, "libvaeassoc_%s.so" % name)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(obj_dir, src + ".log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s (%s):\n%s" % (src, name, r.stderr[-4000:]))
        objs.append(obj)
    r = subprocess.run([_nvcc(), "-shared", "-o", lib] + objs + ["-cudart", "static", "-ldl", "-lpthread"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    if verbose:
        print("[build] linked", lib)
    return lib


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [o for o, _ in res]
    if any(ch for _, ch in res) or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-cudart", "static", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
        if verbose:
            print("[build] linked", LIB)
    return LIB


if __name__ == "__main__":
    if "--w16" in sys.argv:
        build_variant("w16", ["VAEASSOC_EPI_WARPS=16"])
    elif "--timeline" in sys.argv:
        build_variant("tl", ["VAEASSOC_TIMELINE=1"])
    elif "--epi-debug" in sys.argv:
        build_variant("dbg", ["VAEASSOC_EPI_DEBUG=1"])
    else:
        build(force="--force" in sys.argv)

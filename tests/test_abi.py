"""CPU-side checks of the C-ABI: the library builds, loads without a GPU, exports every symbol the header
declares, the ctypes binding covers the header exactly, and creation fails LOUDLY without a device."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vaeassoc.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vaeassoc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    fns = header_functions()
    assert "vaeassoc_train_step" in fns and "vaeassoc_create" in fns and len(fns) >= 30


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vaeassoc_[a-z0-9_]+)", out))
    missing = [f for f in header_functions() if f not in exported]
    assert not missing, missing
    extra = [f for f in exported if f not in header_functions()]
    assert not extra, "exported but undeclared: %s" % extra


def test_binding_matches_header(built_lib):
    from vae_assoc_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load()
    assert lib.vaeassoc_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header(built_lib, tmp_path):
    """sizeof(vaeassoc_config / vaeassoc_tensor_info) as gcc sees the header == the ctypes mirror."""
    from vae_assoc_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "vaeassoc.h"\nint main(){printf("%zu %zu %zu\\n",'
                   'sizeof(vaeassoc_config),sizeof(vaeassoc_tensor_info),sizeof(vaeassoc_modality));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b, c = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert (a, b, c) == (C.sizeof(_lib.Config), C.sizeof(_lib.TensorInfo), C.sizeof(_lib.Modality))


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must refuse to run (no oracle / torch fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vae_assoc_b200 import _lib, vae_assoc
    archs = [dict(scope="image", hidden_conv=False, n_hidden_recog_1=8, n_hidden_recog_2=8, n_hidden_gener_1=8,
                  n_hidden_gener_2=8, n_input=16, n_z=2)]
    with pytest.raises(vae_assoc.VaeAssocError):
        vae_assoc.AssocVariationalAutoEncoder(archs, batch_size=4)
    lib = _lib.load()
    cfg = _lib.Config()
    cfg.abi_version = _lib.ABI_VERSION
    cfg.n_modalities = 1; cfg.batch_size = 4; cfg.n_z = 2
    cfg.mod[0].n_input = 16; cfg.mod[0].n_hidden_recog_1 = 8; cfg.mod[0].n_hidden_recog_2 = 8
    h = _lib.Handle()
    assert lib.vaeassoc_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CPU fallback" in lib.vaeassoc_last_error(None) or b"CUDA" in lib.vaeassoc_last_error(None)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under vae_assoc_b200/ may import it."""
    pkg = os.path.join(ROOT, "vae_assoc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "from oracle" not in text and "import oracle" not in text, f

"""The product library loads on a CPU-only box and exports every symbol include/asph.h declares; it never computes
without a GPU (asph_create -> ASPH_ERR_NO_DEVICE) and nothing in the product imports the oracle."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "asph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asph_[a-z0-9_]+)\s*\(", text)))


def _has_gpu():
    if not shutil.which("nvidia-smi"):
        return False
    return subprocess.call(["nvidia-smi", "-L"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 0


def test_header_declares_the_step_api():
    names = _declared_functions()
    for must in ("asph_create", "asph_destroy", "asph_step", "asph_step_physics", "asph_step_adaptivity", "asph_get_field",
                 "asph_get_neighbors_csr", "asph_set_state", "asph_create_distributed"):
        assert must in names


def test_product_library_exports_every_declared_symbol(asph):
    if not os.path.exists(asph.PRODUCT_LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = C.CDLL(asph.PRODUCT_LIB)
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    lib.asph_backend_name.restype = C.c_char_p
    assert lib.asph_backend_name() == b"cuda-sm100a"


def test_oracle_libraries_export_the_same_abi(asph, oracle32, oracle64):
    for lib in (oracle32, oracle64):
        missing = [n for n in _declared_functions() if not hasattr(lib, n)]
        assert not missing, missing


def test_product_is_built_for_sm_100a(asph):
    out = subprocess.run(["cuobjdump", "--list-elf", asph.PRODUCT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_no_cpu_fallback(asph, cuda_lib, default_params):
    """Without a GPU the product refuses to compute."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    pos = np.zeros((4, 2), np.float32); vel = np.zeros((4, 2), np.float32); mass = np.ones(4, np.float32)
    with pytest.raises(asph.AsphError) as e:
        asph.FluidSimulation(default_params, pos, vel, mass, lib=cuda_lib)
    assert e.value.code == 11  # ASPH_ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "adaptive-sph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "sph_oracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f
    out = subprocess.run(["ldd", os.path.join(pkg, "csrc", "libasph_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out

"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "poltergeist_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(P):
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(P.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for nm in names:
        assert hasattr(lib, nm), nm
    assert sorted(P._abi.EXPORTS) == names
    assert lib.pgb_abi_version() == 2


def test_no_cpu_fallback_without_device(P):
    """without a CUDA device the product fails loudly (PGB_ERR_NO_DEVICE); it never computes on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.PoltergeistError) as ei:
        P.engine.Context(0)
    assert ei.value.status == 5
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "poltergeist.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no oracle", "").replace("test oracle", "").lower() or f in ("_abi.py", "engine.py", "__init__.py"), f
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_plain_c_client_compiles_and_fails_loudly_without_a_gpu(P, tmp_path):
    """the header is valid C11 and a C program links against the library; without a device it reports
    PGB_ERR_NO_DEVICE instead of computing anything"""
    import subprocess
    import torch
    libdir = os.path.dirname(P.LIB_PATH)
    exe = str(tmp_path / "abi_client")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_client", "abi_client.c"), "-o", exe, "-L", libdir, "-lpoltergeist_b200", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    gexe = str(tmp_path / "abi_group_client")                   # the multi-GPU (pgb_group_*) client: same rules
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_client", "abi_group_client.c"), "-o", gexe, "-L", libdir, "-lpoltergeist_b200", "-lm",
                           f"-Wl,-rpath,{libdir}"])
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU test runs the client")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr
    out = subprocess.run([gexe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr

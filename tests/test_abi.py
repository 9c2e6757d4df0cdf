"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU, exports every symbol
include/vcg.h declares, and refuses to compute without a Blackwell GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vcg.h")).read()
    return sorted(set(re.findall(r"VCG_API[^;(]*?\b(vcg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from vcg_b200 import binding
    binding.build_library()
    lib = binding.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in vcg.h but not exported"
        assert n in binding.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(binding.PROTOTYPES) == set(names)
    assert b"sm_100a" in lib.vcg_version()


def test_library_has_no_libcuda_link_dependency():
    """The driver entry point for tensor-map encoding is resolved at run time, so dlopen works on GPU-less hosts."""
    import subprocess
    from vcg_b200 import binding
    out = subprocess.run(["ldd", binding.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    from vcg_b200 import binding
    lib = binding.load_library()
    cfg = binding.VcgConfig(16, 100, 128, 0, 0, 0, 8, 8)
    h = ctypes.c_void_p()
    assert lib.vcg_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert lib.vcg_last_error()
    with pytest.raises(RuntimeError):
        binding.check(1)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_python_api_has_no_cpu_fallback():
    from vcg_b200.engine import Engine
    from vcg_b200 import ops
    with pytest.raises(RuntimeError):
        Engine(16)
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))

"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols include/glare_b200.h
declares (no compute calls here -- there is no GPU in the authoring container)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "glare_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(glare_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(glare_lib):
    from glare_b200 import _lib
    names = header_functions()
    assert names, "no declarations parsed from the header"
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "glare_" in l)
    assert exported == names, "exports and include/glare_b200.h differ"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and header differ"
    for n in names:
        assert isinstance(getattr(glare_lib, n), ctypes._CFuncPtr)


def test_abi_version_and_error_strings(glare_lib):
    assert glare_lib.glare_abi_version() == 1
    assert glare_lib.glare_error_string(0) == b"ok"
    assert b"bad argument" in glare_lib.glare_error_string(-1)
    assert glare_lib.glare_flow_net_floats() == 9552


def test_sass_is_sm100a(glare_lib):
    from glare_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_operators_refuse_cpu_tensors(glare_lib):
    """No CPU fallback: like the reference op (deform_conv.py:143-144) CPU tensors raise."""
    from glare_b200 import ops
    with pytest.raises(NotImplementedError):
        ops.vq_lookup(torch.zeros(1, 3, 2, 2), torch.zeros(8, 4))
    with pytest.raises(NotImplementedError):
        ops.modulated_deform_conv(torch.zeros(1, 4, 3, 3), torch.zeros(1, 72, 3, 3), torch.zeros(1, 36, 3, 3),
                                  torch.zeros(4, 4, 3, 3), None, 1, 1, 1, 1, 4)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_engine_fails_loudly_without_gpu(sd_g, sd_v):
    from glare_b200.engine import GlareEngine
    with pytest.raises(RuntimeError):
        GlareEngine(sd_g, sd_v)

"""CPU: the C-ABI library loads and exports every symbol include/dfgpu.h declares
(no compute calls: there is no GPU here), and it fails loudly — never falls back —
when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dfgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from deepflow_b200 import build
    path = build.build_dfgpu()
    lib = C.CDLL(path if isinstance(path, str) and path.endswith(".so") else os.path.join(ROOT, "deepflow_b200", "libdfgpu.so"))
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"missing exports: {missing}"


def test_struct_layouts_match_header():
    """ctypes mirrors vs. sizes the compiler sees (checked through a tiny C probe)."""
    import subprocess
    import tempfile
    from deepflow_b200 import abi
    probe = r'''
#include <stdio.h>
#include "dfgpu.h"
int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(dfgpu_col), sizeof(dfgpu_frame), sizeof(dfgpu_ip),
 sizeof(dfgpu_dict_snapshot), sizeof(dfgpu_geo_entry), sizeof(dfgpu_cfg), sizeof(dfgpu_counters), sizeof(dfgpu_pos));return 0;}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        out = subprocess.check_output([os.path.join(d, "p")]).decode().split()
    exp = [C.sizeof(x) for x in (abi.Col, abi.Frame, abi.Ip, abi.DictSnapshot, abi.GeoEntry, abi.Cfg, abi.Counters, abi.Pos)]
    assert [int(x) for x in out] == exp


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from deepflow_b200.native import DfgpuError, Engine
    with pytest.raises(DfgpuError):
        Engine(0)

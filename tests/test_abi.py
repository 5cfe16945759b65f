"""CPU tests of the drop-in boundary: the library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no fallback).  No compute calls here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "goicp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(goicp_[a-z0-9_]+)\s*\(", text)) - {"goicp_allgather_fn"})


def test_library_exports_every_declared_symbol(pkg):
    pkg.build()
    lib = pkg.lib()
    decl = declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(lib, s), s
    assert sorted(pkg.ABI_SYMBOLS) == decl
    nm = subprocess.run(["nm", "-D", "--defined-only", pkg.LIB_PATH], capture_output=True, text=True).stdout
    for s in decl:
        assert re.search(rf"\bT {s}\b", nm), s


def test_struct_layouts_match_header(pkg):
    # sizes computed by the C compiler for the header's structs vs the ctypes mirrors
    src = r'''
    #include <stdio.h>
    #include "goicp_b200.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(goicp_params), sizeof(goicp_result), sizeof(goicp_icp_result), sizeof(goicp_snapshot), sizeof(goicp_inner_result)); return 0; }
    '''
    exe = os.path.join(ROOT, "cuda-go-icp_b200", "build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=src, text=True, check=True)
    sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(pkg.Params), C.sizeof(pkg.Result), C.sizeof(pkg.IcpResult), C.sizeof(pkg.Snapshot), C.sizeof(pkg.InnerResult)]


def test_defaults_mirror_reference_constructor(pkg):
    p = pkg.Params()
    pkg.lib().goicp_default_params(C.byref(p))
    assert p.dt_size == 300 and p.dt_expand == 2.0 and p.trim_fraction == 0.0 and p.do_trim == 1       # jly_goicp.cpp:55-63
    assert list(p.trans_cube) == [-0.5, -0.5, -0.5, 1.0]                                                # :50-53
    assert np.allclose(list(p.rot_cube), [-np.pi, -np.pi, -np.pi, 2 * np.pi], atol=1e-6)                 # :44-48
    assert p.icp_max_iter == 10000                                                                       # jly_icp3d.hpp:113


def test_no_gpu_means_loud_failure_not_a_cpu_fallback(pkg, bunny):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = pkg.GoICP(1e-3)
    g.pModel, g.pData = bunny["model_s"], bunny["data_s"]
    with pytest.raises(pkg.GoicpError) as e:
        g.BuildDT()
    assert e.value.code == 2 and "no CUDA device" in str(e.value)
    with pytest.raises(pkg.GoicpError):
        g.NN(bunny["data_s"])
    with pytest.raises(pkg.GoicpError):
        g.Register()


def test_product_never_touches_the_oracle():
    """the package sources must not import, link or open anything under oracle/"""
    pkgdir = os.path.join(ROOT, "cuda-go-icp_b200")
    for dirpath, _, files in os.walk(pkgdir):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "goicp_oracle" not in text and "libref_goicp" not in text and "from oracle" not in text and "import oracle" not in text, f
    ldd = subprocess.run(["ldd", os.path.join(pkgdir, "libgoicp_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd

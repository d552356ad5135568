"""The C-ABI library loads and exports every symbol include/rpt.h declares (no compute calls: no GPU here)."""
import ctypes as ct
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "rpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rpt_[a-z_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points(pkg):
    assert declared_functions() == sorted(pkg.ffi.RPT_SYMBOLS)


def test_library_exports_every_declared_symbol(pkg):
    path = pkg.ffi.LIB_PATH
    if not os.path.exists(path):
        import __graft_entry__ as graft

        graft.build_cuda()
    lib = ct.CDLL(path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/rpt.h but not exported"
    lib.rpt_abi_version.restype = ct.c_uint32
    assert lib.rpt_abi_version() == pkg.ffi.ABI_VERSION


def test_oracle_exports_the_same_surface(pkg, oracle):
    for name in ("scene_create", "scene_destroy", "render_pt", "trace_primary", "trace_rays", "last_error", "abi_version", "output_film",
                 "scene_bake_importance_map", "render_samples", "set_num_threads", "num_threads"):
        assert hasattr(oracle, "rpto_" + name)
    assert oracle.rpto_abi_version() == pkg.ffi.ABI_VERSION


def test_struct_sizes_match_the_c_layout(pkg):
    """ctypes mirrors vs the sizes the C compiler computes (catches field drift between rpt.h and ffi.py)."""
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "rpt.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(RptInstance), sizeof(RptMesh), sizeof(RptMaterial), sizeof(RptTexture), sizeof(RptEnvironment), sizeof(RptCamera), sizeof(RptSceneDesc), sizeof(RptRenderParams), sizeof(RptCounters), sizeof(RptSceneStats), sizeof(RptOutputSettings), sizeof(RptImapBake), sizeof(RptKernelTime), sizeof(RptMultiTimes));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = list(map(int, subprocess.check_output([exe]).split()))
    f = pkg.ffi
    mirrors = [f.RptInstance, f.RptMesh, f.RptMaterial, f.RptTexture, f.RptEnvironment, f.RptCamera, f.RptSceneDesc, f.RptRenderParams, f.RptCounters, f.RptSceneStats,
               f.RptOutputSettings, f.RptImapBake, f.RptKernelTime, f.RptMultiTimes]
    assert sizes == [ct.sizeof(m) for m in mirrors]


def test_product_path_fails_loudly_without_the_extension(pkg, tmp_path):
    with pytest.raises(pkg.ffi.RptError):
        pkg.ffi.load_library(str(tmp_path / "missing.so"))

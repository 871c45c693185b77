"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/syncgpu.h declares.
No compute call is made here (there is no GPU in the build container)."""
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "syncgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_everything_the_header_declares():
    from oatk_b200 import lib
    L = lib.library()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert missing == []
    assert sorted(lib.SYMBOLS) == names


def test_strerror_and_no_device_is_loud():
    from oatk_b200 import lib
    L = lib.library()
    assert L.sg_strerror(0) == b"ok"
    assert b"smers" in L.sg_strerror(-6)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(lib.SgError):
            lib.Context(0)          # no CPU fallback: creating a context without a GPU must fail


def test_sass_is_sm100a():
    import subprocess
    from oatk_b200 import lib
    out = subprocess.run(["cuobjdump", "-lelf", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_product_does_not_touch_the_oracle():
    """the product package must not import, link or execute anything under oracle/"""
    pkg = os.path.join(ROOT, "oatk_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in src and "liboracle" not in src and "sync_oracle" not in src and "libref" not in src, f


def test_host_layer_exports_everything_its_headers_declare():
    """liboatk_gpu.so (the reference-named functions over libsyncgpu) builds and carries every function declared in
    oatk_b200/host/*.h -- the names a maintainer links the reference's callers against"""
    import ctypes
    from oatk_b200.host import build_host
    host = os.path.join(ROOT, "oatk_b200", "host")
    names = set()
    for h in ("syncmer_gpu.h", "graph_gpu.h", "fastx_gpu.h"):
        txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(host, h)).read(), flags=re.S)
        txt = re.sub(r"#ifndef __(SYNCMER|GRAPH|SYNCASM)_H__.*?#endif", "", txt, flags=re.S)   # struct blocks, no prototypes
        names |= set(re.findall(r"^\s*(?:[a-z_0-9]+\s+\**)+([a-z_0-9]+)\s*\([^;{]*\)\s*;", txt, flags=re.M))
    for must in ("sr_read_mem", "sr_db_stat", "collect_syncmer_from_reads", "make_syncmer_graph", "process_mergeable_unitigs", "scg_consensus",
                 "read_error_correction", "scg_read_alignment", "scg_ra_utg_coverage", "scg_ra_arc_coverage", "scg_multiplex", "scg_demultiplex",
                 "asmg_pop_bubble", "asmg_drop_tip", "asmg_remove_weak_crosslink", "syncasm", "sr_read_files"):
        assert must in names, must
    try:
        L = ctypes.CDLL(build_host.build())
    except OSError as e:
        pytest.skip("host layer not loadable here: %s" % e)
    missing = sorted(n for n in names if not hasattr(L, n))
    assert missing == []

"""The C-ABI library loads on a CPU-only box and exports every symbol include/hnsw_b200.h
declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hnsw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pkg):
    lib = ctypes.CDLL(pkg.lib_path())
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "libhnsw_b200.so does not export %s" % name


def test_binding_covers_header(pkg):
    from pgvector_hnsw_partitioning_b200 import hnsw
    assert set(declared_symbols()) <= set(hnsw._SIGS), set(declared_symbols()) - set(hnsw._SIGS)


def test_opclass_surface(pkg):
    assert set(pkg.OPCLASSES) == {"vector_l2_ops", "vector_ip_ops", "vector_cosine_ops", "vector_l1_ops",
                                  "halfvec_l2_ops", "halfvec_ip_ops", "halfvec_cosine_ops", "halfvec_l1_ops"}
    with pytest.raises(pkg.HnswError):
        pkg.HnswIndex(8, "bit_hamming_ops")          # bit / sparsevec operator classes are out of scope


@pytest.mark.skipif(has_gpu(), reason="checks the no-device failure mode")
def test_no_cpu_fallback(pkg):
    with pytest.raises(pkg.HnswError) as ei:
        pkg.HnswIndex(8, "vector_l2_ops")
    assert "CUDA" in str(ei.value) or "device" in str(ei.value)


def test_partition_routing_matches_oracle(pkg, oracle):
    ids = np.arange(-5, 5000, dtype=np.int64)
    for P in (1, 2, 8, 13):
        got = pkg.partition_route(ids, P)
        want = np.array([oracle.splitmix64(int(i)) % P for i in ids], np.int32)
        assert (got == want).all()
        assert got.min() >= 0 and got.max() < P
    # roughly uniform
    cnt = np.bincount(pkg.partition_route(np.arange(80000), 8), minlength=8)
    assert cnt.min() > 9000 and cnt.max() < 11000


def test_level_draw_matches_oracle(pkg, oracle):
    L = pkg.load_library()
    L.hb_level_for.restype = ctypes.c_int
    L.hb_level_for.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int]
    for m in (4, 16, 48):
        for i in range(3000):
            assert L.hb_level_for(99, i, m) == oracle.level_for(99, i, m)


def test_pgvector_pages_info_on_cpu(pkg, oracle):
    """host logic of the on-disk page reader that needs no device: metapage fields and tuple counts of pages
    written in the recalled pgvector layout (tests/pgpages_writer.py)."""
    import numpy as np
    from conftest import clustered
    from pgpages_writer import write_pages
    x = clustered(500, 12, 8, seed=5)
    orc = oracle.Index(12, 8, 32, 0, 0, oracle.CANON, seed=1)
    orc.build(x, np.arange(500, dtype=np.int64) + 65537)
    g = orc.export()
    for scatter in (False, True):
        blob = write_pages(g, 8, 32, 12, scatter=scatter)
        assert pkg.pgvector_pages_info(blob) == (12, 8, 32, g.n, g.upper_rows)
    with pytest.raises(pkg.HnswError):
        pkg.pgvector_pages_info(b"\0" * 8192)


def test_header_is_plain_c_and_links(pkg, tmp_path):
    """include/hnsw_b200.h compiles as C99 -pedantic and every entry point it declares links against the .so
    (tests/c/abi_smoke.c); its no-GPU error paths behave."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "tests", "c", "abi_smoke.c")).read()
    missing = [n for n in declared_symbols() if "REF(%s)" % n not in src]
    assert not missing, "tests/c/abi_smoke.c does not reference %s" % missing
    so = pkg.lib_path()
    exe = str(tmp_path / "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "c", "abi_smoke.c"), "-o", exe, so, "-Wl,-rpath," + os.path.dirname(so)])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "entry points" in out.stdout


def test_c_example_builds(pkg, tmp_path):
    """examples/knn.c (the ABI end to end from C) compiles and links; without a device it stops at the first call."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = pkg.lib_path()
    exe = str(tmp_path / "knn")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "knn.c"),
                           so, "-Wl,-rpath," + os.path.dirname(so), "-lm", "-o", exe])
    if not has_gpu():
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 1 and "no CUDA device" in out.stderr


def test_postgres_glue_compiles_against_the_am_api(tmp_path):
    """pg/hnsw_b200_am.c (the index-AM glue a maintainer adds) compiles as C against tests/c/pgstub -- a stand-in for
    the PostgreSQL 16 headers that declares the IndexAmRoutine callbacks with their real signatures -- and against
    include/hnsw_b200.h, with -Werror: a callback or C-ABI signature that drifts breaks this test.  Every hb_* call
    the glue makes must resolve against the built library."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    obj = str(tmp_path / "glue.o")
    cmd = ["gcc", "-std=gnu99", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-DHNSW_B200_WITH_POSTGRES",
           "-I", os.path.join(root, "tests", "c", "pgstub"), "-I", os.path.join(root, "include"), "-c",
           os.path.join(root, "pg", "hnsw_b200_am.c"), "-o", obj]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    syms = subprocess.run(["nm", "-u", obj], capture_output=True, text=True).stdout
    used = sorted(set(re.findall(r"\bU (hb_[a-z0-9_]+)", syms)))
    assert {"hb_index_create", "hb_insert", "hb_beginscan", "hb_rescan", "hb_gettuple", "hb_endscan", "hb_bulk_delete",
            "hb_vacuum_repair", "hb_index_trim", "hb_scan_set_iterative"} <= set(used)
    declared = set(declared_symbols())
    assert set(used) <= declared
    defined = subprocess.run(["nm", "-g", "--defined-only", obj], capture_output=True, text=True).stdout
    assert "hnsw_b200_handler" in defined
    # without the macro the file is empty (the library build never depends on PostgreSQL)
    out = subprocess.run(["gcc", "-std=gnu99", "-c", os.path.join(root, "pg", "hnsw_b200_am.c"), "-o", str(tmp_path / "empty.o")],
                         capture_output=True, text=True)
    assert out.returncode == 0

"""CPU tests of the host layer: config parity with the oracle, the C-ABI struct marshalling, and
that the shared library loads and exports every symbol include/hmse.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import hmse_b200
from hmse_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_matches_oracle():
    assert np.array_equal(hmse_b200.gear_table(), oracle.gear_table())
    for avg, nc in [(4096, 2), (8192, 2), (8192, 1), (16384, 3), (65536, 2)]:
        a, b = hmse_b200.CDCConfig.for_avg(avg, nc), oracle.CDCConfig.for_avg(avg, nc)
        assert (a.min_size, a.avg_size, a.max_size, a.mask_s, a.mask_l, a.gear_seed) == \
               (b.min_size, b.avg_size, b.max_size, b.mask_s, b.mask_l, b.gear_seed)
    assert hmse_b200.SimConfig().seeds == oracle.SimConfig().seeds
    assert hmse_b200.SimConfig().bands == 32 and hmse_b200.SimConfig().rows == 4


def test_cdc_struct_marshalling():
    for _ in range(50):  # the gear array must survive until it is copied (regression: dangling temporary)
        s = api._cdc_struct(hmse_b200.CDCConfig())
        assert np.array_equal(np.frombuffer(bytes(s.gear), dtype=np.uint64), oracle.gear_table())
    assert (s.min_size, s.avg_size, s.max_size) == (2048, 8192, 32768)
    assert s.mask_s == oracle.PAPER_MASK_S and s.mask_l == oracle.PAPER_MASK_L
    assert C.sizeof(_lib.CdcCfg) == 16 + 16 + 2048


def test_config_validation():
    with pytest.raises(ValueError):
        hmse_b200.CDCConfig(32, 64, 128)
    with pytest.raises(ValueError):
        hmse_b200.CDCConfig(2048, 1024, 4096)
    with pytest.raises(ValueError):
        hmse_b200.SimConfig(128, 16, 4)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hmse.h")).read()
    declared = set(re.findall(r"\b(hmse_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"hmse_ctx", "hmse_cdc_cfg", "hmse_corpus_cfg"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hmse_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        hmse_b200.chunk(b"x" * 100)
    with pytest.raises(RuntimeError):
        hmse_b200.Context()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hmse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "tests/model" not in src or f == "deflate_core.h", f


def test_stream_piece_schedule_covers_the_stream():
    """IngestStream.schedule (host logic, no GPU): piece ends are increasing, 16-byte aligned except the last, end at n,
    with and without the ramp at the start (a prefetched stream needs none)."""
    from hmse_b200 import ingest, CDCConfig
    st = ingest.IngestStream.__new__(ingest.IngestStream)
    st.cdc = CDCConfig()
    for piece in (1 << 20, 64 << 20, 1 << 30, 2 << 30):
        st.piece = piece
        for n in (1, 4095, 5_000_000, (3 << 30) + 12345, 10_000_000_000):
            for ramp in (True, False):
                ends = st.schedule(n, ramp)
                assert ends[-1] == n and all(b > a for a, b in zip(ends, ends[1:])) and ends[0] > 0
                assert all(e % 16 == 0 for e in ends[:-1]), (piece, n, ramp)
                sizes = [ends[0]] + [b - a for a, b in zip(ends, ends[1:])]
                assert max(sizes) <= max(piece, 16) + 16 or len(ends) == 1 or n <= 2 * piece, (piece, n, ramp, sizes[:4])
            if n > 4 * piece:
                assert st.schedule(n, False)[0] == piece and st.schedule(n, True)[0] < piece


def test_header_is_plain_c_and_the_c_programs_build_without_a_gpu(tmp_path):
    """include/hmse.h must stay a plain C header (the boundary a firmware or host maintainer binds): it compiles alone as
    C11 with -Wall -Werror -pedantic, and the two C callers of tests/c_abi/ compile and LINK against the in-tree library
    here, with no GPU (they only run in the -m gpu suite)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = os.path.join(ROOT, "include")
    tu = tmp_path / "only_header.c"
    tu.write_text('#include "hmse.h"\nint main(void) { return HMSE_OK + (HMSE_ABI_VERSION > 0 ? 0 : 1); }\n')
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(tu), "-o", str(tmp_path / "only_header.o")])
    lib_dir = os.path.join(ROOT, "hmse_b200")
    if not os.path.exists(os.path.join(lib_dir, "libhmse_b200.so")):
        pytest.skip("library not built")
    for src, std in (("hmse_c_ingest.c", "c11"), ("hmse_c_sharded.c", "gnu11")):
        exe = str(tmp_path / src[:-2])
        subprocess.check_call(["gcc", "-O1", "-std=" + std, "-Wall", "-I", inc, "-I", os.path.join(cuda, "include"),
                               os.path.join(ROOT, "tests", "c_abi", src), "-o", exe, "-L", lib_dir, "-lhmse_b200",
                               "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib_dir,
                               "-Wl,-rpath," + os.path.join(cuda, "lib64")])
        assert os.path.exists(exe)

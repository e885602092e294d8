"""The C ABI driven from C: tests/c_abi/hmse_c_ingest.c (gcc + libcudart + libhmse_b200.so, no Python on its path) runs
the call sequence of INTEGRATION.md; its outputs must equal the Python binding's and the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import corpus

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "hmse_c_ingest")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    lib_dir = os.path.join(ROOT, "hmse_b200")
    cmd = ["gcc", "-O1", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(ROOT, "tests", "c_abi", "hmse_c_ingest.c"), "-o", exe, "-L", lib_dir, "-lhmse_b200",
           "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    subprocess.check_call(cmd)
    return exe


@pytest.mark.parametrize("with_dict", [True, False])
def test_c_program_matches_python_binding_and_oracle(ctx, corpus8, tmp_path, with_dict):
    import hmse_b200
    exe = _build(tmp_path)
    data = np.concatenate([corpus8[:3 << 20], corpus8[1 << 20:2 << 20]])
    zd = corpus.zdict() if with_dict else b""
    inp, dic, outp = tmp_path / "in.bin", tmp_path / "dict.bin", tmp_path / "out.bin"
    inp.write_bytes(data.tobytes())
    dic.write_bytes(zd)
    r = subprocess.run([exe, str(inp), str(dic) if with_dict else "-", str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    buf = outp.read_bytes()
    n, m, total = np.frombuffer(buf, dtype=np.uint64, count=3).tolist()
    o = 24
    cuts = np.frombuffer(buf, dtype=np.uint64, count=n, offset=o); o += 8 * n
    dg = np.frombuffer(buf, dtype=np.uint8, count=32 * n, offset=o).reshape(n, 32); o += 32 * n
    canon = np.frombuffer(buf, dtype=np.int64, count=n, offset=o); o += 8 * n
    sel = np.frombuffer(buf, dtype=np.uint64, count=m, offset=o); o += 8 * m
    offs = np.frombuffer(buf, dtype=np.uint64, count=m + 1, offset=o); o += 8 * (m + 1)
    blob = np.frombuffer(buf, dtype=np.uint8, count=total, offset=o)
    assert o + total == len(buf)
    # the oracle
    wcuts = oracle.chunk_c(data)
    assert np.array_equal(cuts, wcuts)
    wdg = oracle.digest(data, wcuts)
    assert np.array_equal(dg, wdg)
    wcanon, wfirst = oracle.dedup(wdg)
    assert np.array_equal(canon, wcanon) and np.array_equal(sel, np.flatnonzero(wfirst))
    outs = oracle.inflate_all(blob, offs, zd)
    starts = np.concatenate([[0], wcuts[:-1]]).astype(np.int64)
    raw = data.tobytes()
    for k, j in enumerate(sel.astype(np.int64)):
        assert outs[k] == raw[starts[j]:int(wcuts[j])]
    # the Python binding produces the same bytes
    res = hmse_b200.Ingest(ctx, hmse_b200.CDCConfig(), zd).run(ctx.stage(data))
    assert np.array_equal(res.blob.cpu().numpy(), blob) and np.array_equal(res.offsets.cpu().numpy().view(np.uint64), offs)
    assert "kernels launched" in r.stdout

"""Builds oracle/hmse_ref.c -> oracle/_build/libhmse_ref.so (gcc).  TEST INFRASTRUCTURE."""
import os
import subprocess


def build(force: bool = False) -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "hmse_ref.c")
    out_dir = os.path.join(here, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libhmse_ref.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-shared", "-fPIC", "-o", out, src])
    return out


if __name__ == "__main__":
    print(build(force=True))

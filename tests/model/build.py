"""Builds tests/model/deflate_model.cpp -> tests/model/libdeflate_model.so (g++).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    _fields_ = [("hash_bytes", C.c_int), ("chain_own", C.c_int), ("chain_dict", C.c_int), ("lazy", C.c_int),
                ("too_far", C.c_int), ("dict_hash_bits", C.c_int), ("mode", C.c_int), ("min_len", C.c_int), ("hist", C.c_int)]


def build(force=False):
    src = os.path.join(HERE, "deflate_model.cpp")
    hdr = os.path.join(HERE, "..", "..", "hmse_b200", "csrc", "deflate_core.h")
    out = os.path.join(HERE, "libdeflate_model.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.model_compress.restype = C.c_int64
    lib.model_compress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(Params), C.c_void_p,
                                   C.c_uint64, C.c_void_p]
    return lib

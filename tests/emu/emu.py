"""ctypes front end of the SIMT-emulated kernels (tests/emu): test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("err_a", C.c_uint32), ("err_b", C.c_uint32), ("checksum", C.c_uint32),
                ("blocks", C.c_uint32), ("declared", C.c_uint32), ("produced", C.c_uint64),
                ("consumed_bits", C.c_uint64), ("resume_bit", C.c_uint64), ("resume_out", C.c_uint64),
                ("trailer_seen", C.c_uint32), ("phase", C.c_uint32), ("stat", C.c_uint32 * 4),
                ("tokens", C.c_uint64), ("matches", C.c_uint64), ("deferred", C.c_uint64), ("ck_done", C.c_uint32),
                ("pad", C.c_uint32), ("cycles", C.c_uint64 * 12)]


def build(name: str, force: bool = False) -> str:
    src = os.path.join(HERE, name + ".cpp")
    lib = os.path.join(HERE, "lib" + name + ".so")
    csrc = os.path.join(HERE, "..", "..", "swift-png_b200", "csrc")
    deps = [src, os.path.join(HERE, "simt.h")] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if force or not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-I" + HERE, "-Wno-attributes",
                        "-o", lib, src], check=True)
    return lib


def load(name: str):
    return C.CDLL(build(name))

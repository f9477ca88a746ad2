"""ctypes wrapper around tests/host_emu/libemu.so (device headers compiled for the host)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emu", "emu.cpp")
LIB = os.path.join(HERE, "host_emu", "libemu.so")


def build():
    deps = [SRC] + [os.path.join(HERE, "..", "ckb_zkp_b200", "csrc", f)
                    for f in ("ptx.cuh", "field.cuh", "curve.cuh", "field_params.cuh", "serialize.cuh", "pairing.cuh", "pairing_params.cuh")]
    if (not os.path.exists(LIB)) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def to_u32(x, n):
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)], dtype=np.uint32)


def from_u32(a):
    return sum(int(w) << (32 * i) for i, w in enumerate(a))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)

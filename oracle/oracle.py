"""ctypes wrapper of the CPU oracle (oracle/libltrans_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Same call surface as
ltrans_b200.host.binding.LtransLib (prefix ``ora_``) so the parity tests feed
both with identical arrays.  PARITY UNPINNED BY THE REFERENCE (see ltrans_oracle.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import ltrans_b200  # noqa: E402,F401
from ltrans_b200.host.binding import LtransLib  # noqa: E402

LIB = os.path.join(_HERE, "libltrans_oracle.so")
ORA_RNG_PHILOX, ORA_RNG_MT = 1, 2


def build(force=False):
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB


class Oracle(LtransLib):
    def __init__(self):
        build()
        super().__init__(LIB, prefix="ora_")

    def set_threads(self, n):
        self.lib.ora_set_threads(self.ctx, C.c_int32(n))

    def set_rng(self, mode):
        self.lib.ora_set_rng(self.ctx, C.c_int32(mode))


def leaf():
    """Raw library handle with restypes set for the exported leaf functions."""
    build()
    lib = C.CDLL(LIB)
    d, i32, pd, pi = C.c_double, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.ora_polintd.restype = d; lib.ora_polintd.argtypes = [pd, pd, d]
    lib.ora_linint.argtypes = [pd, pd, i32, d, pd, pd]
    lib.ora_gridcell.restype = i32; lib.ora_gridcell.argtypes = [pd, pd, d, d]
    lib.ora_inpoly.restype = i32; lib.ora_inpoly.argtypes = [d, d, i32, pd, pd, i32]
    lib.ora_tspsi.argtypes = [i32, pd, pd, pd, pd, pi, pi]
    lib.ora_hval.restype = d; lib.ora_hval.argtypes = [d, i32, pd, pd, pd, pd, pi]
    lib.ora_hpval.restype = d; lib.ora_hpval.argtypes = [d, i32, pd, pd, pd, pd, pi]
    lib.ora_snhcsh.argtypes = [d, pd, pd, pd]
    lib.ora_slevel.restype = d; lib.ora_slevel.argtypes = [d, d, d, d, C.c_float, i32]
    lib.ora_mt_init_genrand.argtypes = [C.c_uint32]
    lib.ora_mt_init_by_array.argtypes = [C.POINTER(C.c_uint32), i32]
    lib.ora_mt_int32.restype = C.c_uint32
    lib.ora_mt_real1.restype = d; lib.ora_mt_real3.restype = d
    lib.ora_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    lib.ora_intersect_reflect.restype = i32
    lib.ora_intersect_reflect.argtypes = [C.c_void_p, d, d, d, d, pd, pd, pd, pd, pi, pi]
    lib.ora_mbounds.restype = i32; lib.ora_mbounds.argtypes = [C.c_void_p, d, d]
    lib.ora_ibounds.restype = i32; lib.ora_ibounds.argtypes = [C.c_void_p, d, d, pd]
    return lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))

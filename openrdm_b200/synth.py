"""openrdm_b200.synth -- ctypes binding of the HOST-ONLY synthetic input generator (libordm_synth.so,
host/ordm_synth_host.cc over include/ordm_synth.h): the same bits as ordm_synth_fill_device /
ordm_synth_fill_host of the GPU library, without mapping any CUDA code.  Used by the CPU reference
arm of bench.py and by the golden-fixture generators under tests/golden/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "host", "libordm_synth.so")
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.run(["make", "-s", "-C", os.path.join(PKG, "host"), "libordm_synth.so"], check=True)
        lib = C.CDLL(LIB_PATH)
        D, S, I = C.POINTER(C.c_double), C.c_size_t, C.c_int
        lib.ordm_synth_fill_host.argtypes = [C.c_uint64, S, S, S, I, D, D, D, D, D]
        lib.ordm_synth_rdms.argtypes = [C.c_uint64, I, I, I, I, D, D, D]
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else C.POINTER(C.c_double)()


def synth_rdms(seed: int, nact: int, nelec_a: int, nelec_b: int, ndet: int = 8, want_d2: bool = True):
    A1a = np.zeros((nact, nact), order="F"); A1b = np.zeros((nact, nact), order="F")
    D2 = np.zeros((nact * nact, nact * nact), order="F") if want_d2 else None
    if load().ordm_synth_rdms(seed, nact, nelec_a, nelec_b, ndet, _d(A1a), _d(A1b), _d(D2)):
        raise ValueError("bad synthetic RDM request")
    return A1a, A1b, D2


def synth_fill_host(seed: int, npts_total: int, p0: int, m: int, n: int, gga: bool):
    """w, phi, (phix, phiy, phiz) or None of the synthetic grid slice [p0, p0+m)."""
    w = np.zeros(m); phi = np.zeros((m, n), order="F")
    g = [np.zeros((m, n), order="F") for _ in range(3)] if gga else [None] * 3
    if load().ordm_synth_fill_host(seed, npts_total, p0, m, n, _d(w), _d(phi), _d(g[0]), _d(g[1]), _d(g[2])):
        raise ValueError("bad synthetic shard")
    return w, phi, (g if gga else None)

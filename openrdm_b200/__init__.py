"""openrdm_b200 -- B200-native MC-PDFT energy engine (OpenRDM's libmcpdft/libfunc hot path).

The product is the C-ABI shared library ``libopenrdm_b200.so`` (include/openrdm_b200.h) built
from ``csrc/`` (CUDA, sm_100a) and the C++ host mirror of OpenRDM's classes in ``host/``.
This module is only the ctypes binding the Python-side harnesses (tests/, bench.py) use to
call that C ABI; it adds no compute and has no fallback: if the library or a CUDA device is
missing, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORDM_LIB") or os.path.join(PKG, "libopenrdm_b200.so")  # ORDM_LIB: A/B builds in the lab

FUNC_SVWN, FUNC_PBE, FUNC_REVPBE, FUNC_BLYP, FUNC_BOP, FUNC_WPBE = 0, 1, 2, 3, 4, 5
FUNC_ID = {"SVWN": 0, "PBE": 1, "REVPBE": 2, "BLYP": 3, "BOP": 4, "WPBE": 5}
EX_LSDA, EC_VWN3, EX_PBE, EC_PBE, EX_REVPBE, EX_B88, EC_LYP, EC_B88_OP, EX_WPBE, EC_PW92 = range(10)
OPT_DIAGNOSTICS, OPT_PI_GRADIENT, OPT_FULLY_TRANSLATED, OPT_SERIAL_STAGES = 1, 2, 4, 8

VEC_NAMES = [
    "rhoa", "rhob", "rho",
    "rhoa_x", "rhoa_y", "rhoa_z", "rhob_x", "rhob_y", "rhob_z",
    "sigma_aa", "sigma_ab", "sigma_bb",
    "pi", "R", "tr_rhoa", "tr_rhob",
    "tr_sigma_aa", "tr_sigma_ab", "tr_sigma_bb",
    "pi_x", "pi_y", "pi_z", "w",
]
VEC_ID = {n: i for i, n in enumerate(VEC_NAMES)}


class OrdmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"openrdm_b200 error {code}: {msg}")
        self.code = code


class Result(C.Structure):
    _fields_ = [(k, C.c_double) for k in (
        "e_tot", "ex", "ec", "eclass", "n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b",
        "ms_density", "ms_pi", "ms_xc", "ms_reduce", "ms_total")] + [
        ("npts", C.c_uint64), ("gpu_launches", C.c_uint32), ("overlapped", C.c_uint32),
        ("pi_guard_points", C.c_uint32), ("density_tiled", C.c_uint32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


LOAD_PREFER_AO, LOAD_DENSE_RDM, LOAD_NO_GRADIENTS = 1, 2, 4


class DataInfo(C.Structure):
    _fields_ = [("npts", C.c_uint64), ("nao", C.c_int), ("nmo", C.c_int), ("ncore", C.c_int), ("nact", C.c_int),
                ("gga", C.c_int), ("used_ao", C.c_int), ("used_sparse_rdm", C.c_int), ("e_core", C.c_double), ("e_hartree", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class InputFile:
    """ordm_file_*: the data.h5 schema from the raw container (or HDF5 when the library was built with it)."""
    _NP = {0: np.float64, 1: np.int64, 2: np.int32}

    def __init__(self, path: str):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.ordm_file_open(path.encode(), C.byref(self.h))
        if rc:
            raise OrdmError(rc, self.lib.ordm_last_error(None).decode())

    def dataset(self, name: str) -> np.ndarray:
        dt, nd, dims, ptr = C.c_int(), C.c_int(), (C.c_uint64 * 4)(), C.c_void_p()
        rc = self.lib.ordm_file_dataset(self.h, name.encode(), C.byref(dt), C.byref(nd), dims, C.byref(ptr))
        if rc:
            raise KeyError(self.lib.ordm_last_error(None).decode())
        shape = tuple(int(dims[i]) for i in range(nd.value))
        n = int(np.prod(shape)) if shape else 1
        item = np.dtype(self._NP[dt.value]).itemsize
        buf = (C.c_char * (n * item)).from_address(ptr.value) if n else b""
        return np.frombuffer(buf, dtype=self._NP[dt.value], count=n).reshape(shape).copy()

    def close(self):
        if self.h:
            self.lib.ordm_file_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HybridInputs(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("e_reference", "e_nuclear_repulsion", "e_kinetic", "e_nuclear_attraction", "e_coulomb",
                                          "e_hf_exchange", "e_mp2_correlation")]


class HybridResult(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("e_total", "wf_contribution", "dft_contribution", "one_electron", "two_electron", "lambda_")]


METHODS = {"MCPDFT": 0, "1H_MCPDFT": 1, "1DH_MCPDFT": 2, "RS_MCPDFT": 3, "RS1H_MCPDFT": 4, "RS1DH_MCPDFT": 5}


def hybrid_energy(method: str, lam: float, ex: float, ec: float, **scalars) -> dict:
    """MCPDFTSolver::compute_energy's assembly (mcpdft_solver.cc:849-925); scalars = the fields of ordm_hybrid_inputs."""
    inp = HybridInputs(**{k: float(v) for k, v in scalars.items()})
    out = HybridResult()
    rc = load().ordm_hybrid_energy(METHODS[method], float(lam), C.byref(inp), float(ex), float(ec), C.byref(out))
    if rc:
        raise OrdmError(rc, load().ordm_last_error(None).decode())
    return {k: getattr(out, k) for k, _ in out._fields_}


_lib = None


def load(rebuild: bool = False) -> C.CDLL:
    """Load (building first if the sources are newer) the C-ABI library."""
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    if not os.environ.get("ORDM_LIB") and (rebuild or _build.stale()):   # missing, or older than any source under csrc/ / include/
        _build.build(force=rebuild)
    lib = C.CDLL(LIB_PATH)
    P, D, S, I = C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_int
    lib.ordm_version.restype = C.c_char_p
    lib.ordm_last_error.restype = C.c_char_p
    lib.ordm_last_error.argtypes = [P]
    lib.ordm_device_vector.restype = P
    lib.ordm_device_vector.argtypes = [P, I]
    lib.ordm_create.argtypes = [I, C.POINTER(P)]
    lib.ordm_destroy.argtypes = [P]
    lib.ordm_destroy.restype = None
    lib.ordm_set_stream.argtypes = [P, P]
    lib.ordm_set_options.argtypes = [P, C.c_uint]
    lib.ordm_synchronize.argtypes = [P]
    lib.ordm_set_omega.argtypes = [P, C.c_double]
    lib.ordm_host_alloc.argtypes = [S, C.POINTER(P)]
    lib.ordm_host_free.argtypes = [P]
    lib.ordm_set_grid.argtypes = [P, S, D]
    lib.ordm_set_orbitals.argtypes = [P, S, I, S, D, D, D, D]
    lib.ordm_set_orbitals_ao.argtypes = [P, S, I, S, D, D, D, D, I, D]
    lib.ordm_set_rdm1.argtypes = [P, I, D, D]
    lib.ordm_set_rdm2ab_dense.argtypes = [P, I, D]
    lib.ordm_set_active_space.argtypes = [P, I, I, D, D, D]
    IP = C.POINTER(C.c_int)
    lib.ordm_set_active_space_sparse.argtypes = [P, I, I, S, IP, IP, D, S, IP, IP, D, S, IP, IP, IP, IP, D, I]
    lib.ordm_rdm2ab_sparse_to_dense.argtypes = [I, S, IP, IP, IP, IP, D, I, D]
    for f in ("ordm_build_rho", "ordm_build_pi", "ordm_build_R", "ordm_translate", "ordm_fully_translate", "ordm_comm_detach"):
        getattr(lib, f).argtypes = [P]
    lib.ordm_integrated_densities.argtypes = [P, D, D]
    lib.ordm_functional.argtypes = [P, I, S, D, D, D, D, D, D]
    lib.ordm_energy.argtypes = [P, I, C.c_double, C.POINTER(Result)]
    lib.ordm_energy_async.argtypes = [P, I, C.c_double]
    lib.ordm_energy_wait.argtypes = [P, C.POINTER(Result)]
    lib.ordm_energy_host.argtypes = [P, I, C.c_double, S, I, S, D, D, D, D, D, S, C.POINTER(Result)]
    lib.ordm_file_open.argtypes = [C.c_char_p, C.POINTER(P)]
    lib.ordm_file_close.argtypes = [P]
    lib.ordm_file_close.restype = None
    lib.ordm_file_dataset.argtypes = [P, C.c_char_p, C.POINTER(I), C.POINTER(I), C.POINTER(C.c_uint64), C.POINTER(P)]
    lib.ordm_load_data.argtypes = [P, C.c_char_p, C.c_uint, C.POINTER(DataInfo)]
    lib.ordm_polyradical.argtypes = [P, D, D, D, D, S]
    lib.ordm_hybrid_energy.argtypes = [I, C.c_double, C.POINTER(HybridInputs), C.c_double, C.c_double, C.POINTER(HybridResult)]
    lib.ordm_get_vector.argtypes = [P, I, D, S]
    lib.ordm_set_vector.argtypes = [P, I, D, S]
    lib.ordm_get_orbitals.argtypes = [P, I, D, S]
    lib.ordm_synth_fill_device.argtypes = [P, C.c_uint64, S, S, S, I, I]
    lib.ordm_synth_fill_host.argtypes = [C.c_uint64, S, S, S, I, D, D, D, D, D]
    lib.ordm_synth_rdms.argtypes = [C.c_uint64, I, I, I, I, D, D, D]
    lib.ordm_comm_unique_id.argtypes = [C.c_char_p]
    lib.ordm_comm_attach.argtypes = [P, I, I, C.c_char_p]
    lib.ordm_comm_info.argtypes = [P, C.POINTER(I), C.POINTER(I), C.POINTER(I)]
    lib.ordm_pi_kernel_info.argtypes = [P, C.POINTER(I), C.POINTER(C.c_uint64)]
    lib.ordm_shard_range.argtypes = [S, I, I, C.POINTER(S), C.POINTER(S)]
    lib.ordm_shard_range.restype = None
    _lib = lib
    return lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else C.POINTER(C.c_double)()


def _f(a):
    return None if a is None else np.require(np.asarray(a, dtype=np.float64), requirements=["A", "F"])


def shard_range(npts_total: int, nranks: int, rank: int):
    p0, cnt = C.c_size_t(), C.c_size_t()
    load().ordm_shard_range(npts_total, nranks, rank, C.byref(p0), C.byref(cnt))
    return p0.value, cnt.value


def synth_rdms(seed: int, nact: int, nelec_a: int, nelec_b: int, ndet: int = 8, want_d2: bool = True):
    """Host-side synthetic active-space RDMs (no GPU needed)."""
    A1a = np.zeros((nact, nact), order="F"); A1b = np.zeros((nact, nact), order="F")
    D2 = np.zeros((nact * nact, nact * nact), order="F") if want_d2 else None
    rc = load().ordm_synth_rdms(seed, nact, nelec_a, nelec_b, ndet, _d(A1a), _d(A1b), _d(D2))
    if rc:
        raise OrdmError(rc, load().ordm_last_error(None).decode())
    return A1a, A1b, D2


def rdm2ab_sparse_to_dense(n: int, d2, upper_packed=False) -> np.ndarray:
    """(i, j, k, l, val) list -> dense (n^2, n^2) matrix in the reference convention (host only)."""
    t = [np.ascontiguousarray(x, dtype=np.int32) for x in d2[:4]] + [np.ascontiguousarray(d2[4], dtype=np.float64)]
    out = np.zeros((n * n, n * n), order="F")
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = load().ordm_rdm2ab_sparse_to_dense(n, len(t[4]), ip(t[0]), ip(t[1]), ip(t[2]), ip(t[3]), _d(t[4]), int(upper_packed), _d(out))
    if rc:
        raise OrdmError(rc, load().ordm_last_error(None).decode())
    return out


def synth_fill_host(seed: int, npts_total: int, p0: int, m: int, n: int, gga: bool):
    """Host copy of the synthetic grid slice [p0, p0+m): returns w, phi, (phix, phiy, phiz) or None."""
    w = np.zeros(m); phi = np.zeros((m, n), order="F")
    g = [np.zeros((m, n), order="F") for _ in range(3)] if gga else [None] * 3
    rc = load().ordm_synth_fill_host(seed, npts_total, p0, m, n, _d(w), _d(phi), _d(g[0]), _d(g[1]), _d(g[2]))
    if rc:
        raise OrdmError(rc, load().ordm_last_error(None).decode())
    return w, phi, (g if gga else None)


class Engine:
    """One context = one GPU = one grid shard.  Thin wrapper over the C ABI."""

    def __init__(self, device: int = 0):
        self.lib = load()
        self.ctx = C.c_void_p()
        rc = self.lib.ordm_create(device, C.byref(self.ctx))
        if rc:
            raise OrdmError(rc, self.lib.ordm_last_error(None).decode())
        self.npts = 0

    def _ck(self, rc):
        if rc:
            raise OrdmError(rc, self.lib.ordm_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.lib.ordm_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs
    def set_stream(self, cuda_stream_handle: int):
        self._ck(self.lib.ordm_set_stream(self.ctx, C.c_void_p(cuda_stream_handle)))

    def set_options(self, diagnostics=False, pi_gradient=False, fully_translated=False, serial_stages=False):
        self._ck(self.lib.ordm_set_options(self.ctx, (OPT_DIAGNOSTICS if diagnostics else 0) | (OPT_PI_GRADIENT if pi_gradient else 0)
                                           | (OPT_FULLY_TRANSLATED if fully_translated else 0) | (OPT_SERIAL_STAGES if serial_stages else 0)))

    def set_omega(self, omega: float):
        self._ck(self.lib.ordm_set_omega(self.ctx, float(omega)))

    def set_grid(self, w):
        w = _f(w)
        self.npts = int(w.shape[0])
        self._ck(self.lib.ordm_set_grid(self.ctx, self.npts, _d(w)))

    def set_orbitals(self, phi, grads=None):
        phi = _f(phi)
        npts, n = phi.shape
        g = [_f(x) for x in grads] if grads is not None else [None] * 3
        self._ck(self.lib.ordm_set_orbitals(self.ctx, npts, n, max(npts, 1), _d(phi), _d(g[0]), _d(g[1]), _d(g[2])))

    def set_orbitals_ao(self, ao, C, ao_grads=None):
        """AO values (npts, nao) [+ gradients] and MO coefficients (nao, nmo): phi = ao @ C on the device."""
        ao, Cm = _f(ao), _f(C)
        npts, nao = ao.shape
        g = [_f(x) for x in ao_grads] if ao_grads is not None else [None] * 3
        self._ck(self.lib.ordm_set_orbitals_ao(self.ctx, npts, nao, max(npts, 1), _d(ao), _d(g[0]), _d(g[1]), _d(g[2]), Cm.shape[1], _d(Cm)))

    def set_rdm1(self, D1a, D1b):
        D1a, D1b = _f(D1a), _f(D1b)
        self._ck(self.lib.ordm_set_rdm1(self.ctx, D1a.shape[0], _d(D1a), _d(D1b)))

    def set_rdm2ab_dense(self, D2ab, n):
        D2ab = _f(D2ab)
        self._ck(self.lib.ordm_set_rdm2ab_dense(self.ctx, n, _d(D2ab)))

    def set_active_space(self, ncore, A1a, A1b, D2act):
        A1a, A1b, D2act = _f(A1a), _f(A1b), _f(D2act)
        self._ck(self.lib.ordm_set_active_space(self.ctx, ncore, A1a.shape[0], _d(A1a), _d(A1b), _d(D2act)))

    def set_active_space_sparse(self, ncore, nact, d1a, d1b, d2, upper_packed=False):
        """d1a/d1b = (i, j, val) arrays, d2 = (i, j, k, l, val) arrays; see ordm_set_active_space_sparse."""
        def ints(a): return np.ascontiguousarray(a, dtype=np.int32)
        def ip(a): return a.ctypes.data_as(C.POINTER(C.c_int))
        a = [ints(x) for x in d1a[:2]] + [np.ascontiguousarray(d1a[2], dtype=np.float64)]
        b = [ints(x) for x in d1b[:2]] + [np.ascontiguousarray(d1b[2], dtype=np.float64)]
        t = [ints(x) for x in d2[:4]] + [np.ascontiguousarray(d2[4], dtype=np.float64)]
        self._ck(self.lib.ordm_set_active_space_sparse(self.ctx, ncore, nact, len(a[2]), ip(a[0]), ip(a[1]), _d(a[2]),
                                                       len(b[2]), ip(b[0]), ip(b[1]), _d(b[2]),
                                                       len(t[4]), ip(t[0]), ip(t[1]), ip(t[2]), ip(t[3]), _d(t[4]), int(upper_packed)))

    def synth_fill(self, seed, npts_total, p0, m, n, gga):
        self._ck(self.lib.ordm_synth_fill_device(self.ctx, seed, npts_total, p0, m, n, int(gga)))
        self.npts = m

    # ---- stages
    def build_rho(self): self._ck(self.lib.ordm_build_rho(self.ctx))
    def build_pi(self): self._ck(self.lib.ordm_build_pi(self.ctx))
    def build_R(self): self._ck(self.lib.ordm_build_R(self.ctx))
    def translate(self): self._ck(self.lib.ordm_translate(self.ctx))
    def fully_translate(self): self._ck(self.lib.ordm_fully_translate(self.ctx))

    def integrated_densities(self):
        n = (C.c_double * 3)(); t = (C.c_double * 3)()
        self._ck(self.lib.ordm_integrated_densities(self.ctx, n, t))
        return list(n), list(t)

    def functional(self, term, ra, rb, saa=None, sab=None, sbb=None) -> float:
        ra, rb, saa, sab, sbb = (_f(x) for x in (ra, rb, saa, sab, sbb))
        e = C.c_double()
        self._ck(self.lib.ordm_functional(self.ctx, term, ra.shape[0], _d(ra), _d(rb), _d(saa), _d(sab), _d(sbb), C.byref(e)))
        return e.value

    def energy(self, functional=FUNC_SVWN, eclass=0.0) -> dict:
        r = Result()
        self._ck(self.lib.ordm_energy(self.ctx, functional, eclass, C.byref(r)))
        return r.asdict()

    def energy_async(self, functional=FUNC_SVWN, eclass=0.0):
        """enqueue one step on the context's stream (ordm_energy_async); collect with energy_wait()"""
        self._ck(self.lib.ordm_energy_async(self.ctx, functional, eclass))

    def energy_wait(self) -> dict:
        r = Result()
        self._ck(self.lib.ordm_energy_wait(self.ctx, C.byref(r)))
        return r.asdict()

    def energy_host(self, functional, eclass, w, phi, grads=None, batch_points=0, ld=None) -> dict:
        """w, phi, grads: numpy arrays (ideally views of pinned memory, see host_empty)."""
        npts, n = phi.shape
        g = grads if grads is not None else [None] * 3
        r = Result()
        self._ck(self.lib.ordm_energy_host(self.ctx, functional, eclass, npts, n, ld or max(npts, 1), _d(w), _d(phi),
                                           _d(g[0]), _d(g[1]), _d(g[2]), batch_points, C.byref(r)))
        return r.asdict()

    def load_data(self, path: str, flags: int = 0) -> dict:
        """ordm_load_data: grid, orbitals and active-space RDMs from a data.h5-schema file into this context."""
        info = DataInfo()
        self._ck(self.lib.ordm_load_data(self.ctx, path.encode(), flags, C.byref(info)))
        self.npts = int(info.npts)
        return info.asdict()

    def polyradical(self, want_vectors=True):
        """MCPDFTSolver::polyradical_analysis: ([Tr U, Tr D, Tr N], U(r), D(r), N(r))."""
        tr = (C.c_double * 3)()
        v = [np.zeros(self.npts) for _ in range(3)] if want_vectors else [None] * 3
        self._ck(self.lib.ordm_polyradical(self.ctx, tr, _d(v[0]), _d(v[1]), _d(v[2]), self.npts))
        return list(tr), v[0], v[1], v[2]

    def get_vector(self, name) -> np.ndarray:
        out = np.zeros(self.npts)
        self._ck(self.lib.ordm_get_vector(self.ctx, VEC_ID[name], _d(out), self.npts))
        return out

    def set_vector(self, name, values):
        """MCPDFT::set_<name> (mcpdft.h:125-153): install a caller-provided per-point vector."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._ck(self.lib.ordm_set_vector(self.ctx, VEC_ID[name], _d(v), v.shape[0]))

    def get_orbitals(self, phi_out, grads_out=None):
        """Copy the resident phi (and grad phi) into caller-provided F-ordered (npts, n) arrays."""
        for comp, arr in enumerate([phi_out] + (list(grads_out) if grads_out is not None else [])):
            self._ck(self.lib.ordm_get_orbitals(self.ctx, comp, _d(arr), max(arr.shape[0], 1)))

    def synchronize(self): self._ck(self.lib.ordm_synchronize(self.ctx))

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load().ordm_comm_unique_id(buf)
        if rc:
            raise OrdmError(rc, load().ordm_last_error(None).decode())
        return buf.raw

    def comm_attach(self, nranks, rank, uid: bytes):
        self._ck(self.lib.ordm_comm_attach(self.ctx, nranks, rank, C.create_string_buffer(uid, 128)))

    def pi_kernel_info(self) -> dict:
        """Which kernel builds the active-space on-top pair density (see ordm_pi_kernel_info)."""
        n, m = C.c_int(), C.c_uint64()
        self._ck(self.lib.ordm_pi_kernel_info(self.ctx, C.byref(n), C.byref(m)))
        return {"int8_slice_pairs": n.value, "mma_per_tile": m.value}

    def comm_info(self) -> dict:
        n, r, p = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.lib.ordm_comm_info(self.ctx, C.byref(n), C.byref(r), C.byref(p)))
        return {"nranks": n.value, "rank": r.value, "allreduce": "nvlink peer stores (own kernel)" if p.value else "ncclAllReduce"}


def host_empty(shape, order="F") -> np.ndarray:
    """numpy array backed by page-locked host memory (ordm_host_alloc); freed with the process."""
    n = int(np.prod(shape))
    p = C.c_void_p()
    rc = load().ordm_host_alloc(n * 8, C.byref(p))
    if rc:
        raise OrdmError(rc, load().ordm_last_error(None).decode())
    buf = (C.c_double * n).from_address(p.value)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape, order=order)

"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loaders for the two CPU checkers of the MC-PDFT hot path:

* ``Port``  -> oracle/_build/libmcpdft_oracle.so, the plain-C restatement (mcpdft_oracle.c)
* ``Ref``   -> oracle/_ref/libopenrdm_ref.so, OpenRDM's own unmodified sources
               (/root/reference/src/libmcpdft, src/libfunc) behind ref_harness.cc

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (openrdm_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libmcpdft_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libopenrdm_ref.so")
REF_PSI4_SO = os.path.join(HERE, "_ref", "libopenrdm_ref_psi4.so")
REF_OMP_SO = os.path.join(HERE, "_ref", "libopenrdm_ref_omp.so")  # same sources, WITH_OPENMP ON
# functional ids of the C port's drivers: 1 SVWN, 0 PBE (energy.cc:114-120: anything but "SVWN" is PBE),
# plus the Psi4 plugin's REVPBE / BLYP / BOP (mcpdft_solver.cc:742-787)
FUNC_ID = {"SVWN": 1, "PBE": 0, "REVPBE": 2, "BLYP": 3, "BOP": 4, "WPBE": 5}

VEC_NAMES = [
    "rhoa", "rhob", "rho",
    "rhoa_x", "rhoa_y", "rhoa_z", "rhob_x", "rhob_y", "rhob_z",
    "sigma_aa", "sigma_ab", "sigma_bb",
    "pi", "R", "tr_rhoa", "tr_rhob",
    "tr_sigma_aa", "tr_sigma_ab", "tr_sigma_bb",
    "pi_x", "pi_y", "pi_z",
]
NVEC = len(VEC_NAMES)
GGA_ONLY = {n for n in VEC_NAMES if ("_x" in n or "_y" in n or "_z" in n or "sigma" in n)}


class OrcResult(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("e_tot", "ex", "ec", "eclass", "n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RefTiming(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("t_rho", "t_grad_rho", "t_pi", "t_grad_pi", "t_R_translate", "t_functional", "t_total")]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


@dataclass
class Outcome:
    scalars: dict
    vecs: dict
    timing: dict | None = None
    seconds: float | None = None


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  `make ref` is a no-op without /root/reference."""
    subprocess.run(["make", "-s", "-C", HERE, "port"] + (["ref"] if ref else []), check=True)


def _f64(a, order="F"):
    if a is None:
        return None
    return np.require(np.asarray(a, dtype=np.float64), requirements=["A", "O"] + (["F"] if order == "F" else ["C"]))


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else C.POINTER(C.c_double)()


def _vec_table(npts, want=True):
    if not want:
        return None, None
    arrs = [np.zeros(npts, dtype=np.float64) for _ in range(NVEC)]
    tbl = (C.POINTER(C.c_double) * NVEC)(*[_ptr(a) for a in arrs])
    return arrs, tbl


def _vecs_dict(arrs, is_gga):
    if arrs is None:
        return {}
    return {n: a for n, a in zip(VEC_NAMES, arrs) if is_gga or n not in GGA_ONLY}


class Port:
    """The plain-C restatement."""

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build(ref=False)
        self.lib = C.CDLL(PORT_SO)
        self.lib.orc_ex_lsda.restype = C.c_double
        self.lib.orc_ec_vwn3.restype = C.c_double
        self.lib.orc_ex_pbe.restype = C.c_double
        self.lib.orc_ec_pbe.restype = C.c_double
        for f in ("orc_ex_pbe_kappa", "orc_ex_b88", "orc_ec_lyp", "orc_ec_b88_op", "orc_ex_wpbe", "orc_ec_pw92"):
            getattr(self.lib, f).restype = C.c_double

    def energy(self, w, phi, D1a, D1b, D2ab, functional="SVWN", grads=None, eclass=0.0,
               with_pi_grad=False, want_vecs=True) -> Outcome:
        """Dense reference operator surface: phi (npts,n), D1 (n,n), D2ab (n^2,n^2)."""
        phi = _f64(phi); npts, n = phi.shape
        is_gga = grads is not None
        gx, gy, gz = (_f64(g) for g in grads) if is_gga else (None, None, None)
        w, D1a, D1b, D2ab = _f64(w), _f64(D1a), _f64(D1b), _f64(D2ab)
        arrs, tbl = _vec_table(npts, want_vecs)
        res = OrcResult()
        self.lib.orc_energy(C.c_size_t(npts), n, int(is_gga), FUNC_ID.get(functional, 0), _ptr(w), _ptr(phi),
                            _ptr(gx), _ptr(gy), _ptr(gz), _ptr(D1a), _ptr(D1b), _ptr(D2ab),
                            C.c_double(eclass), int(with_pi_grad), C.byref(res), tbl)
        return Outcome(res.asdict(), _vecs_dict(arrs, is_gga))

    def energy_coreactive(self, w, phi, ncore, A1a, A1b, D2act, functional="SVWN", grads=None,
                          eclass=0.0, with_pi_grad=False, want_vecs=True) -> Outcome:
        phi = _f64(phi); npts, n = phi.shape
        nact = n - ncore
        is_gga = grads is not None
        gx, gy, gz = (_f64(g) for g in grads) if is_gga else (None, None, None)
        w, A1a, A1b, D2act = _f64(w), _f64(A1a), _f64(A1b), _f64(D2act)
        arrs, tbl = _vec_table(npts, want_vecs)
        res = OrcResult()
        self.lib.orc_energy_coreactive(C.c_size_t(npts), ncore, nact, int(is_gga), FUNC_ID.get(functional, 0),
                                       _ptr(w), _ptr(phi), _ptr(gx), _ptr(gy), _ptr(gz), _ptr(A1a),
                                       _ptr(A1b), _ptr(D2act), C.c_double(eclass), int(with_pi_grad),
                                       C.byref(res), tbl)
        return Outcome(res.asdict(), _vecs_dict(arrs, is_gga))

    def set_translation(self, fully: bool):
        """False: MCPDFT::translate (mcpdft_energy's choice); True: MCPDFT::fully_translate."""
        self.lib.orc_set_translation(int(bool(fully)))

    def functional(self, which, w, ra, rb, saa=None, sab=None, sbb=None) -> float:
        """which in {'EX_LSDA','EC_VWN3','EX_PBE','EC_PBE'} on caller-provided vectors."""
        w, ra, rb = _f64(w), _f64(ra), _f64(rb)
        n = C.c_size_t(w.shape[0])
        saa, sab, sbb = (_f64(s) for s in (saa, sab, sbb))
        if which == "EX_LSDA":
            return self.lib.orc_ex_lsda(n, _ptr(ra), _ptr(rb), _ptr(w))
        if which == "EC_VWN3":
            return self.lib.orc_ec_vwn3(n, _ptr(ra), _ptr(rb), _ptr(w))
        if which == "EX_PBE":
            return self.lib.orc_ex_pbe(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sbb), _ptr(w))
        if which == "EC_PBE":
            return self.lib.orc_ec_pbe(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sab), _ptr(sbb), _ptr(w))
        if which in ("EX_PBE_I", "EX_REVPBE"):
            return self.lib.orc_ex_pbe_kappa(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sbb), _ptr(w), C.c_double(1.245 if which == "EX_REVPBE" else 0.804))
        if which == "EX_B88":
            return self.lib.orc_ex_b88(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sbb), _ptr(w))
        if which == "EC_LYP":
            return self.lib.orc_ec_lyp(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sab), _ptr(sbb), _ptr(w))
        if which == "EC_B88_OP":
            return self.lib.orc_ec_b88_op(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sbb), _ptr(w))
        if which == "EC_PW92":
            return self.lib.orc_ec_pw92(n, _ptr(ra), _ptr(rb), _ptr(w))
        if which == "EX_WPBE":
            return self.lib.orc_ex_wpbe(n, _ptr(ra), _ptr(rb), _ptr(saa), _ptr(sbb), _ptr(w), C.c_double(self.omega))
        raise ValueError(which)

    omega = 0.0

    def set_omega(self, omega: float):
        """MCPDFT_OMEGA (plugin.cc:50): range-separation parameter of EX_wPBE_I / functional "WPBE"."""
        self.omega = float(omega)
        self.lib.orc_set_omega(C.c_double(omega))

    def polyradical(self, w, phi, D1a, D1b):
        """MCPDFTSolver::polyradical_analysis restated (mcpdft_solver.cc:3001-3240): (traces[3], U(r), D(r), N(r))."""
        phi = _f64(phi); npts, n = phi.shape
        w, D1a, D1b = _f64(w), _f64(D1a), _f64(D1b)
        U, D, N = (np.zeros(npts) for _ in range(3))
        tr = (C.c_double * 3)()
        self.lib.orc_polyradical(C.c_size_t(npts), n, _ptr(w), _ptr(phi), _ptr(D1a), _ptr(D1b), _ptr(U), _ptr(D), _ptr(N), tr)
        return list(tr), U, D, N

    def hybrid_energy(self, method, lam, e_ref, e_nuc, e_kin, e_ne, e_coulomb, e_hf_x, e_mp2, ex, ec) -> float:
        """Energy assembly of MCPDFTSolver::compute_energy restated (mcpdft_solver.cc:849-925)."""
        self.lib.orc_hybrid_energy.restype = C.c_double
        return self.lib.orc_hybrid_energy(int(method), *(C.c_double(x) for x in (lam, e_ref, e_nuc, e_kin, e_ne, e_coulomb, e_hf_x, e_mp2, ex, ec)))

    def kron_tpdm(self, D1a, D1b):
        D1a, D1b = _f64(D1a), _f64(D1b)
        n = D1a.shape[0]
        out = np.zeros((n * n, n * n), dtype=np.float64, order="F")
        self.lib.orc_kron_tpdm(n, _ptr(D1a), _ptr(D1b), _ptr(out))
        return out


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_ref_psi4() -> bool:
    return os.path.exists(REF_PSI4_SO)


class RefPsi4:
    """The Psi4 plugin's own functional routines (src/libpsi4/lsda_gga_functionals.cc, compiled
    unmodified against stand-in Psi4 headers: oracle/shim_psi4, oracle/ref_psi4_harness.cc)."""

    ID = {"EX_LSDA": 0, "EC_VWN3": 1, "EX_PBE_I": 2, "EC_PBE_I": 3, "EX_REVPBE": 4, "EX_B88": 5, "EC_LYP": 6, "EC_B88_OP": 7,
          "EX_WPBE": 8, "EC_PW92": 9}

    def set_omega(self, omega: float):
        self.lib.ref_psi4_set_omega(C.c_double(omega))

    def __init__(self):
        if not have_ref_psi4():
            build(ref=True)
        if not have_ref_psi4():
            raise FileNotFoundError(REF_PSI4_SO)
        self.lib = C.CDLL(REF_PSI4_SO)
        self.lib.ref_psi4_functional.restype = C.c_double

    def functional(self, which, w, ra, rb, saa=None, sab=None, sbb=None) -> float:
        w, ra, rb = _f64(w), _f64(ra), _f64(rb)
        z = np.zeros_like(w)
        saa, sab, sbb = (_f64(s) if s is not None else z for s in (saa, sab, sbb))
        return self.lib.ref_psi4_functional(self.ID[which], C.c_size_t(w.shape[0]), _ptr(w), _ptr(ra), _ptr(rb),
                                            _ptr(saa), _ptr(sab), _ptr(sbb))


class Ref:
    """OpenRDM's own compiled hot path (oracle/_ref)."""

    FUNC_ID = {"EX_LSDA": 0, "EC_VWN3": 1, "EX_PBE": 2, "EC_PBE": 3}

    def __init__(self, openmp: bool = False):
        """openmp=True loads the WITH_OPENMP build of the same sources (timing comparison only)."""
        if not have_ref():
            build(ref=True)
        so = REF_OMP_SO if openmp else REF_SO
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        self.lib = C.CDLL(so)
        self.lib.ref_functional.restype = C.c_double
        self.lib.ref_energy_threads.restype = C.c_double

    def set_translation(self, fully: bool):
        self.lib.ref_set_translation(int(bool(fully)))

    def energy_as_written(self, w, phi, D1a, D1b, D2ab, functional="SVWN", grads=None, eclass=0.0,
                          want_vecs=True, quiet=True) -> Outcome:
        """mcpdft::mcpdft_energy() exactly as tests/main.cc drives it."""
        phi = _f64(phi); npts, n = phi.shape
        is_gga = grads is not None
        gx, gy, gz = (_f64(g) for g in grads) if is_gga else (None, None, None)
        w, D1a, D1b, D2ab = _f64(w), _f64(D1a), _f64(D1b), _f64(D2ab)
        arrs, tbl = _vec_table(npts, want_vecs)
        res = OrcResult(); secs = C.c_double(0)
        self.lib.ref_energy_as_written(C.c_size_t(npts), n, int(is_gga), functional.encode(), _ptr(w), _ptr(phi),
                                       _ptr(gx), _ptr(gy), _ptr(gz), _ptr(D1a), _ptr(D1b), _ptr(D2ab),
                                       C.c_double(eclass), int(quiet), C.byref(res), tbl, C.byref(secs))
        return Outcome(res.asdict(), _vecs_dict(arrs, is_gga), seconds=secs.value)

    def energy_staged(self, w, phi, ncore, A1a, A1b, D2act, functional="SVWN", grads=None, eclass=0.0,
                      with_pi_grad=False, want_vecs=True, quiet=True) -> Outcome:
        phi = _f64(phi); npts, n = phi.shape
        nact = n - ncore
        is_gga = grads is not None
        gx, gy, gz = (_f64(g) for g in grads) if is_gga else (None, None, None)
        w, A1a, A1b, D2act = _f64(w), _f64(A1a), _f64(A1b), _f64(D2act)
        arrs, tbl = _vec_table(npts, want_vecs)
        res = OrcResult(); tm = RefTiming()
        self.lib.ref_energy_staged(C.c_size_t(npts), ncore, nact, int(is_gga), functional.encode(), _ptr(w),
                                   _ptr(phi), _ptr(gx), _ptr(gy), _ptr(gz), _ptr(A1a), _ptr(A1b), _ptr(D2act),
                                   C.c_double(eclass), int(with_pi_grad), int(quiet), C.byref(res), tbl,
                                   C.byref(tm))
        return Outcome(res.asdict(), _vecs_dict(arrs, is_gga), timing=tm.asdict())

    def energy_threads(self, nthreads, w, phi, ncore, A1a, A1b, D2act, functional="SVWN", grads=None,
                       eclass=0.0, with_pi_grad=False) -> Outcome:
        phi = _f64(phi); npts, n = phi.shape
        nact = n - ncore
        is_gga = grads is not None
        gx, gy, gz = (_f64(g) for g in grads) if is_gga else (None, None, None)
        w, A1a, A1b, D2act = _f64(w), _f64(A1a), _f64(A1b), _f64(D2act)
        res = OrcResult()
        secs = self.lib.ref_energy_threads(int(nthreads), C.c_size_t(npts), ncore, nact, int(is_gga),
                                           functional.encode(), _ptr(w), _ptr(phi), _ptr(gx), _ptr(gy), _ptr(gz),
                                           _ptr(A1a), _ptr(A1b), _ptr(D2act), C.c_double(eclass),
                                           int(with_pi_grad), C.byref(res))
        return Outcome(res.asdict(), {}, seconds=secs)

    def functional(self, which, w, ra, rb, saa=None, sab=None, sbb=None) -> float:
        w, ra, rb = _f64(w), _f64(ra), _f64(rb)
        z = np.zeros_like(w)
        saa, sab, sbb = (_f64(s) if s is not None else z for s in (saa, sab, sbb))
        return self.lib.ref_functional(self.FUNC_ID[which], C.c_size_t(w.shape[0]), _ptr(w), _ptr(ra), _ptr(rb),
                                       _ptr(saa), _ptr(sab), _ptr(sbb))

    def build_tpdm(self, D1a, D1b):
        D1a, D1b = _f64(D1a), _f64(D1b)
        n = D1a.shape[0]
        out = np.zeros((n * n, n * n), dtype=np.float64, order="F")
        self.lib.ref_build_tpdm(n, _ptr(D1a), _ptr(D1b), _ptr(out))
        return out

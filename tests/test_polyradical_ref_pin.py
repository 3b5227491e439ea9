"""Pins the contraction of the polyradical oracle to OpenRDM's own compiled code.

MCPDFTSolver::polyradical_analysis (src/libpsi4/mcpdft_solver.cc:3001-3240) needs Psi4 and cannot be built here, so
`orc_polyradical` restates it.  Its per-point part, X(r) = sum_ij phi_i(r) phi_j(r) X_ij (:3146-3157), is however the very
contraction MCPDFT::build_density_functions performs for rho_a with D1a = X (mcpdft.cc:83), and its traces (:3185-3187) are
that function's integrated alpha density (mcpdft.cc:92).  So the unmodified reference library (oracle/_ref), fed the
matrices U, D, N as "alpha 1-RDMs" with a zero beta block, must reproduce the restatement's U(r), D(r), N(r) and traces;
what stays a restatement are the three element-wise formulas of :3104-3108, re-derived below in numpy."""
import numpy as np
import pytest


@pytest.mark.parametrize("n,npts,seed", [(4, 200, 1), (6, 333, 2), (7, 64, 3)])
def test_polyradical_contraction_matches_the_reference_build_rho(port, ref, n, npts, seed):
    rng = np.random.default_rng(seed)
    phi = np.asfortranarray(rng.uniform(-0.7, 0.7, (npts, n)))
    w = rng.random(npts) / npts
    occ_a, occ_b = rng.random(n), rng.random(n)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    D1a = np.asfortranarray(Q @ np.diag(occ_a) @ Q.T)
    D1b = np.asfortranarray(Q @ np.diag(occ_b) @ Q.T)
    tr, U, D, N = port.polyradical(w, phi, D1a, D1b)
    pop = D1a + D1b                                                   # n_ij (:3082-3084)
    mats = {"U": 1.0 - np.abs(1.0 - pop), "D": pop * (2.0 - pop), "N": pop ** 2 * (2.0 - pop) ** 2}   # :3104-3108, element-wise
    zero1 = np.zeros((n, n), order="F")
    zero2 = np.zeros((n * n, n * n), order="F")
    for k, (name, got) in enumerate((("U", U), ("D", D), ("N", N))):
        out = ref.energy_as_written(w, phi, np.asfortranarray(mats[name]), zero1, zero2, "SVWN")
        want = out.vecs["rhoa"]
        assert np.abs(got - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), name
        assert abs(tr[k] - out.scalars["n_a"]) <= 1e-12 * max(1.0, abs(tr[k])), name

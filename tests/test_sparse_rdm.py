"""Sparse RDM ingestion: Psi4-style element lists and PySCF-style
upper-packed COO groups.  CPU part: the host-side expansion to the reference's dense convention
against numpy and the oracle.  GPU part: the engine fed with lists against the oracle fed with the
dense matrices."""
import numpy as np
import pytest

from helpers import check_scalars, check_vectors, random_case


def psi4_lists(A1a, A1b, D2act, rng, drop=0.0):
    """Dense active RDMs -> Psi4-style lists: every non-zero element once (mcpdft_solver.h:71-83)."""
    n = A1a.shape[0]
    d1 = []
    for A in (A1a, A1b):
        i, j = np.nonzero(A)
        d1.append((i, j, A[i, j]))
    r, c = np.nonzero(D2act)
    # dense element (nu*n+mu, sigma*n+lambda) multiplies phi_mu phi_nu phi_lambda phi_sigma (mcpdft.cc:204)
    i, j, k, l = r % n, r // n, c % n, c // n
    perm = rng.permutation(len(r))
    return d1[0], d1[1], (i[perm], j[perm], k[perm], l[perm], D2act[r, c][perm])


def pyscf_coo(dm1a, dm1b, dm2ab):
    """Symmetric arrays -> upper-packed COO as libpyscf.py:246-463 writes them (i<=j, k<=l, values unscaled)."""
    n = dm1a.shape[0]
    d1 = []
    for A in (dm1a, dm1b):
        i, j = np.triu_indices(n)
        keep = np.abs(A[i, j]) > 1e-20
        d1.append((i[keep], j[keep], A[i, j][keep]))
    idx = [(i, j, k, l) for i in range(n) for j in range(i, n) for k in range(n) for l in range(k, n) if abs(dm2ab[i, j, k, l]) > 1e-20]
    i, j, k, l = (np.array(x) for x in zip(*idx))
    return d1[0], d1[1], (i, j, k, l, dm2ab[i, j, k, l])


def symmetric_case(rng, n, ndet=3):
    """Ensemble of determinants: symmetric dm1, dm2ab[i,j,k,l] = sum_c w_c a_c[i,j] b_c[k,l] (chemist's notation)."""
    w = rng.random(ndet) + 0.2; w /= w.sum()
    dm1a = np.zeros((n, n)); dm1b = np.zeros((n, n)); dm2 = np.zeros((n, n, n, n))
    for c in range(ndet):
        U, _ = np.linalg.qr(rng.normal(size=(n, n)))
        oa = np.zeros(n); ob = np.zeros(n)
        oa[rng.permutation(n)[:max(1, n // 2)]] = 1; ob[rng.permutation(n)[:max(1, n // 2)]] = 1
        a = U @ np.diag(oa) @ U.T; b = U @ np.diag(ob) @ U.T
        a = 0.5 * (a + a.T); b = 0.5 * (b + b.T)
        dm1a += w[c] * a; dm1b += w[c] * b; dm2 += w[c] * np.einsum("ij,kl->ijkl", a, b)
    return dm1a, dm1b, dm2


def dense_from_chemist(dm2):
    n = dm2.shape[0]
    D2 = np.zeros((n * n, n * n), order="F")
    for i in range(n):
        for j in range(n):
            D2[j * n + i, :] = dm2[i, j].T.reshape(-1)  # column l*n+k <- dm2[i,j,k,l]
    return D2


def test_psi4_list_expands_to_reference_convention(ordm):
    rng = np.random.default_rng(1)
    n = 4
    _, _, _, A1a, A1b, D2 = random_case(rng, n, 8, "nonsym")
    D2 = np.where(rng.random(D2.shape) < 0.6, D2, 0.0)
    _, _, t = psi4_lists(A1a, A1b, D2, rng)
    back = ordm.rdm2ab_sparse_to_dense(n, t, upper_packed=False)
    assert np.array_equal(back, D2)
    with pytest.raises(ordm.OrdmError):
        ordm.rdm2ab_sparse_to_dense(n, ([0], [0], [0], [n], [1.0]))  # index out of range


def test_pyscf_coo_expands_to_reference_convention(ordm, port):
    rng = np.random.default_rng(2)
    n = 3
    dm1a, dm1b, dm2 = symmetric_case(rng, n)
    _, _, t = pyscf_coo(dm1a, dm1b, dm2)
    assert (t[0] <= t[1]).all() and (t[2] <= t[3]).all()
    back = ordm.rdm2ab_sparse_to_dense(n, t, upper_packed=True)
    want = dense_from_chemist(dm2)
    assert np.allclose(back, want, rtol=0, atol=1e-16)
    with pytest.raises(ordm.OrdmError):
        ordm.rdm2ab_sparse_to_dense(n, ([1], [0], [0], [0], [1.0]), upper_packed=True)  # below the diagonal
    # and the oracle sees the same energy from either dense matrix
    phi, grads, w, *_ = random_case(rng, n, 200, "ensemble")
    a = port.energy(w, phi, dm1a, dm1b, back, "PBE", grads)
    b = port.energy(w, phi, dm1a, dm1b, want, "PBE", grads)
    assert abs(a.scalars["e_tot"] - b.scalars["e_tot"]) < 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("ncore,nact,npts,gga", [(0, 4, 700, True), (3, 5, 1000, True), (2, 6, 333, False)])
def test_engine_psi4_lists_vs_oracle(engine, port, ordm, ncore, nact, npts, gga):
    rng = np.random.default_rng(40 + nact)
    phi, grads, w, A1a, A1b, D2 = random_case(rng, nact, npts, "nonsym")
    D2 = np.where(rng.random(D2.shape) < 0.5, D2, 0.0)  # genuinely sparse
    if ncore:
        phic = np.asfortranarray(rng.uniform(-0.4, 0.4, (npts, ncore)))
        gc = [np.asfortranarray(rng.uniform(-0.4, 0.4, (npts, ncore))) for _ in range(3)]
        phi = np.asfortranarray(np.hstack([phic, phi]))
        grads = [np.asfortranarray(np.hstack([a, b])) for a, b in zip(gc, grads)]
    fn = "PBE" if gga else "SVWN"
    want = port.energy_coreactive(w, phi, ncore, A1a, A1b, D2, fn, grads if gga else None, eclass=0.5)
    d1a, d1b, t = psi4_lists(A1a, A1b, D2, rng)
    engine.set_options(diagnostics=True)
    engine.set_grid(w); engine.set_orbitals(phi, grads if gga else None)
    engine.set_active_space_sparse(ncore, nact, d1a, d1b, t, upper_packed=False)
    got = engine.energy(1 if gga else 0, 0.5)
    check_scalars(got, want.scalars)
    check_vectors(engine.get_vector, want.vecs, gga)


@pytest.mark.gpu
def test_engine_pyscf_coo_vs_oracle(engine, port, ordm):
    rng = np.random.default_rng(77)
    ncore, nact, npts = 2, 5, 900
    dm1a, dm1b, dm2 = symmetric_case(rng, nact)
    phi, grads, w, *_ = random_case(rng, ncore + nact, npts, "ensemble")
    want = port.energy_coreactive(w, phi, ncore, dm1a, dm1b, dense_from_chemist(dm2), "PBE", grads, eclass=-1.0)
    d1a, d1b, t = pyscf_coo(dm1a, dm1b, dm2)
    engine.set_options(diagnostics=True)
    engine.set_grid(w); engine.set_orbitals(phi, grads)
    engine.set_active_space_sparse(ncore, nact, d1a, d1b, t, upper_packed=True)
    got = engine.energy(1, -1.0)
    check_scalars(got, want.scalars)
    check_vectors(engine.get_vector, want.vecs, True)
    with pytest.raises(ordm.OrdmError):
        engine.set_active_space_sparse(ncore, nact, d1a, d1b, ([0], [0], [0], [nact], [1.0]))

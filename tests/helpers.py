"""Shared comparison helpers.  Tolerances are SURVEY.md section 8d / BASELINE.json north_star:
1e-10 Eh on energies and integrated densities, 1e-12*max(1,|x|) per point on rho, Pi (and the
spin densities, gradients, sigma), 1e-10 relative on R where rho > 1e-8, and a zeta-noise-aware
1e-7*rho on the translated quantities (zeta = sqrt(1-R) amplifies rounding noise of R near
closed-shell points: SURVEY.md section 7 "hard parts")."""
import numpy as np

E_TOL = 1e-10
PT_TOL = 1e-12

STRICT = ["rho", "rhoa", "rhob", "pi", "rhoa_x", "rhoa_y", "rhoa_z", "rhob_x", "rhob_y", "rhob_z",
          "sigma_aa", "sigma_ab", "sigma_bb"]
SCALARS = ["e_tot", "ex", "ec", "n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b"]


def random_case(rng, n, npts, kind="ensemble", ndet=3):
    """Same construction as tests/golden/make_golden.py::random_case."""
    env = np.where(rng.random(npts) < 0.85, 1.0, 2.0 ** -rng.integers(0, 47, npts).astype(float))
    phi = np.asfortranarray(rng.uniform(-0.5, 0.5, (npts, n)) * env[:, None])
    grads = [np.asfortranarray(rng.uniform(-0.5, 0.5, (npts, n)) * env[:, None]) for _ in range(3)]
    w = rng.random(npts) / max(npts, 1)
    D1a = np.zeros((n, n)); D1b = np.zeros((n, n)); D2 = np.zeros((n * n, n * n))
    c = rng.random(ndet) + 0.2; c /= c.sum()
    ne = max(1, n // 2)
    for k in range(ndet):
        U, _ = np.linalg.qr(rng.normal(size=(n, n)))
        oa = np.zeros(n); ob = np.zeros(n)
        oa[rng.permutation(n)[:ne]] = 1
        if kind == "polarised":
            pass
        else:
            ob[rng.permutation(n)[:ne]] = 1
        a = U @ np.diag(oa) @ U.T; b = U @ np.diag(ob) @ U.T
        D1a += c[k] * a; D1b += c[k] * b; D2 += c[k] * np.kron(a, b)
    if kind == "nonsym":
        D1a = D1a + 0.05 * rng.normal(size=(n, n)); D1b = D1b + 0.05 * rng.normal(size=(n, n))
        D2 = D2 + 0.02 * rng.normal(size=D2.shape)
    return phi, grads, w, np.asfortranarray(D1a), np.asfortranarray(D1b), np.asfortranarray(D2)


def check_scalars(got: dict, want: dict, tol=E_TOL, keys=SCALARS):
    for k in keys:
        # trn_a / trn_b individually carry the zeta = sqrt(1-R) noise (see module docstring)
        t = 1e-7 if k in ("trn_a", "trn_b") else tol
        assert abs(got[k] - want[k]) <= t * max(1.0, abs(want[k])), f"{k}: got {got[k]!r} want {want[k]!r} diff {got[k]-want[k]:.3e}"


def check_vectors(get, want: dict, gga: bool):
    """get(name) -> ndarray from the engine; want = oracle vectors."""
    rho = want["rho"]
    for k in STRICT:
        if k not in want or (not gga and ("_x" in k or "_y" in k or "_z" in k or "sigma" in k)):
            continue
        a, b = get(k), want[k]
        err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
        assert err.max(initial=0.0) <= PT_TOL, f"{k}: max err {err.max():.3e} at {err.argmax()}"
    live = rho > 1e-8
    if live.any():
        a, b = get("R")[live], want["R"][live]
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        assert rel.max() <= 1e-10, f"R: max rel err {rel.max():.3e}"
    for k in ("tr_rhoa", "tr_rhob"):
        a, b = get(k), want[k]
        assert (np.abs(a - b) <= 1e-7 * np.abs(rho) + 1e-300).all(), f"{k}: {np.abs(a-b).max():.3e}"
    if gga:
        g2 = want["sigma_aa"] + want["sigma_bb"] + 2 * want["sigma_ab"]
        for k in ("tr_sigma_aa", "tr_sigma_ab", "tr_sigma_bb"):
            a, b = get(k), want[k]
            assert (np.abs(a - b) <= 1e-7 * np.abs(g2) + 1e-300).all(), f"{k}: {np.abs(a-b).max():.3e}"

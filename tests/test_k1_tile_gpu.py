"""K1t, the TMA-tiled active-space density kernel (openrdm_b200/csrc/k1_tile.cuh), through the C ABI.

The kernel serves the energy path (no diagnostic vectors) on grids with orbital gradients and up to 28 active orbitals;
`ordm_result.density_tiled` says whether it ran.  Every case is checked against the oracle (scalars 1e-10, rho and Pi
1e-12 max(1, |x|) per point) AND against the thread-per-point kernel (`ORDM_K1_TILE=0`) on the same inputs -- over ragged
grid sizes (tail tiles, fewer tiles than SMs, more tiles than ring slots), every coefficient-block size NP = 4 .. 28,
staged core columns in front of the active block (`ORDM_CORE_KEEP`), closed-shell 1-RDMs (the spin-difference role has
nothing to do) and the requests it must decline.  compute-sanitizer is closed on this GPU pool, so memory safety rests on
these ragged sizes plus the zero-padded leading dimensions."""
import numpy as np
import pytest

from helpers import check_scalars

pytestmark = pytest.mark.gpu
FN = {"SVWN": 0, "PBE": 1}


def run(ordm, w, phi, grads, ncore, A1a, A1b, D2, fn="PBE", diagnostics=False, serial=False):
    eng = ordm.Engine(0)
    try:
        eng.set_options(diagnostics=diagnostics, serial_stages=serial)
        eng.set_grid(w); eng.set_orbitals(phi, grads); eng.set_active_space(ncore, A1a, A1b, D2)
        r = eng.energy(FN[fn], 0.0)
        vecs = {k: eng.get_vector(k).copy() for k in ("rho", "pi")}
        return r, vecs
    finally:
        eng.close()


def both_kernels(ordm, monkeypatch, *args, **kw):
    monkeypatch.setenv("ORDM_K1_TILE", "1")
    tiled = run(ordm, *args, **kw)
    monkeypatch.setenv("ORDM_K1_TILE", "0")
    plain = run(ordm, *args, **kw)
    return tiled, plain


def check_pair(tiled, plain, want):
    (rt, vt), (rp, vp) = tiled, plain
    assert rp["density_tiled"] == 0
    check_scalars(rt, want.scalars)
    check_scalars(rt, rp, tol=1e-12)
    for k in ("rho", "pi"):
        a, b, o = vt[k], vp[k], want.vecs[k]
        assert (np.abs(a - o) <= 1e-12 * np.maximum(1.0, np.abs(o))).all(), k
        assert (np.abs(a - b) <= 1e-13 * np.maximum(1.0, np.abs(b))).all(), k


@pytest.mark.parametrize("ncore,nact,npts", [
    (0, 4, 1), (1, 3, 127), (2, 4, 128), (3, 5, 129),          # NP = 4 / 8, one tile and its edges
    (0, 8, 4133), (5, 9, 1000), (7, 12, 513), (2, 13, 2049),    # NP = 8 .. 16, staged core columns up to 7
    (4, 16, 300), (6, 17, 777), (3, 20, 1301),                  # NP = 16 .. 20
    (1, 21, 260), (0, 24, 131), (2, 26, 200), (1, 28, 129),     # NP = 24 / 28: the widest ring slots
])
def test_tiled_density_matches_oracle_and_plain_kernel(ordm, port, monkeypatch, ncore, nact, npts):
    seed = 1000 + 37 * nact + ncore
    n = ncore + nact
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, n, True)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, max(1, nact // 2), max(1, (nact + 1) // 3), 3)   # alpha != beta: both roles work
    want = port.energy_coreactive(w, phi, ncore, A1a, A1b, D2, "PBE", grads, eclass=0.0)
    tiled, plain = both_kernels(ordm, monkeypatch, w, phi, grads, ncore, A1a, A1b, D2)
    # (overlapped == 1: few core orbitals beside a warp-specialised K2 -- the whole density kernel streams underneath K2 on
    #  the side stream, where the tiled kernel's 220 KB of shared memory would not fit beside K2's: k1_density keeps that mode)
    if tiled[0]["overlapped"] != 1:
        assert tiled[0]["density_tiled"] >= 1, "the tiled kernel declined an energy-path request it should serve"
    else:
        assert tiled[0]["density_tiled"] == 0
    check_pair(tiled, plain, want)


@pytest.mark.parametrize("keep", [0, 1, 5, 8, 11])
def test_staged_core_columns_in_split_mode(ordm, port, monkeypatch, keep):
    """ncore >= nact: the core sums run on the side stream beside K2 and leave `keep` columns (ORDM_CORE_KEEP) to the
    density kernel, which stages them in front of the active block."""
    ncore, nact, npts, seed = 19, 14, 5000, 4242   # 14 active orbitals: the warp-specialised K2 that overlaps; 40 tiles
    monkeypatch.setenv("ORDM_CORE_KEEP", str(keep))
    n = ncore + nact
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, n, True)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, 7, 6, 4)
    want = port.energy_coreactive(w, phi, ncore, A1a, A1b, D2, "PBE", grads, eclass=0.0)
    tiled, plain = both_kernels(ordm, monkeypatch, w, phi, grads, ncore, A1a, A1b, D2)
    assert tiled[0]["overlapped"] == 2 and tiled[0]["density_tiled"] == 1
    check_pair(tiled, plain, want)


def test_many_tiles_per_cta_and_closed_shell(ordm, port, monkeypatch):
    """More tiles than SMs x ring slots (every CTA wraps its ring several times), and Dsa == Dsb: the difference role only
    passes the chunks on."""
    ncore, nact, npts, seed = 2, 6, 148 * 128 * 9 + 77, 99
    n = ncore + nact
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, n, True)
    A1a, _, _ = ordm.synth_rdms(seed, nact, 3, 3, 1)
    D2 = np.asfortranarray(np.kron(A1a, A1a))
    want = port.energy_coreactive(w, phi, ncore, A1a, A1a, D2, "PBE", grads, eclass=0.0, want_vecs=False)
    monkeypatch.setenv("ORDM_K1_TILE", "1")
    rt, vt = run(ordm, w, phi, grads, ncore, A1a, A1a, D2)
    monkeypatch.setenv("ORDM_K1_TILE", "0")
    rp, vp = run(ordm, w, phi, grads, ncore, A1a, A1a, D2)
    assert rt["density_tiled"] >= 1 and rp["density_tiled"] == 0
    check_scalars(rt, want.scalars)
    check_scalars(rt, rp, tol=1e-12)
    assert abs(rt["n_a"] - rt["n_b"]) <= 1e-14 * abs(rt["n_a"])
    for k in ("rho", "pi"):
        assert (np.abs(vt[k] - vp[k]) <= 1e-13 * np.maximum(1.0, np.abs(vp[k]))).all(), k


def test_requests_the_tiled_kernel_declines(ordm, port, monkeypatch):
    """Diagnostic vectors, grids without orbital gradients and more than 28 active orbitals stay with k1_density."""
    monkeypatch.setenv("ORDM_K1_TILE", "1")
    seed, ncore, nact, npts = 5, 2, 6, 700
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, ncore + nact, True)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, 3, 2, 3)
    assert run(ordm, w, phi, grads, ncore, A1a, A1b, D2, diagnostics=True)[0]["density_tiled"] == 0
    assert run(ordm, w, phi, None, ncore, A1a, A1b, D2, fn="SVWN")[0]["density_tiled"] == 0
    assert run(ordm, w, phi, grads, ncore, A1a, A1b, D2)[0]["density_tiled"] == 1
    nact = 30
    w, phi, grads = ordm.synth_fill_host(seed, 64, 0, 64, nact, True)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, 15, 15, 2)
    r, _ = run(ordm, w, phi, grads, 0, A1a, A1b, D2)
    assert r["density_tiled"] == 0
    want = port.energy_coreactive(w, phi, 0, A1a, A1b, D2, "PBE", grads, eclass=0.0, want_vecs=False)
    check_scalars(r, want.scalars)


def test_streamed_batches_through_the_tiled_kernel(ordm, port, monkeypatch):
    """ordm_energy_host: every point batch goes through the tiled kernel on its staging buffers (a fresh tensor map per
    buffer), with the same scalars as the resident path."""
    monkeypatch.setenv("ORDM_K1_TILE", "1")
    seed, npts, ncore, nact = 17, 9001, 3, 7
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, ncore + nact, True)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, 4, 3, 4)
    want = port.energy_coreactive(w, phi, ncore, A1a, A1b, D2, "PBE", grads, eclass=0.25, want_vecs=False)
    eng = ordm.Engine(0)
    try:
        eng.set_active_space(ncore, A1a, A1b, D2)
        for batch in (2048, 0):
            st = eng.energy_host(FN["PBE"], 0.25, w, phi, grads, batch_points=batch)
            assert st["density_tiled"] >= 1
            check_scalars(st, want.scalars)
    finally:
        eng.close()


def test_rdm_updates_on_one_engine_reuse_the_device_operands(ordm, port):
    """An SCF loop calls ordm_set_active_space every iteration: the device operands (DMMA fragments, int8 image, pair table,
    K1t's coefficient block and tensor maps) are re-used or re-sized in place.  Growing, shrinking and repeating the active
    space on ONE engine must give what a fresh engine gives."""
    seed, npts, ncore = 31, 900, 3
    cases = [20, 6, 20, 13, 6]      # int8 path -> small wide-tile K2 -> int8 again (same sizes: buffers re-used) -> WS K2 -> small
    nmax = ncore + max(cases)
    w, phi, grads = ordm.synth_fill_host(seed, npts, 0, npts, nmax, True)
    eng = ordm.Engine(0)
    try:
        eng.set_grid(w)
        for i, nact in enumerate(cases):
            n = ncore + nact
            ph = np.asfortranarray(phi[:, :n]); gr = [np.asfortranarray(g[:, :n]) for g in grads]
            A1a, A1b, D2 = ordm.synth_rdms(seed + i, nact, max(1, nact // 2), max(1, nact // 3), 3)
            eng.set_orbitals(ph, gr); eng.set_active_space(ncore, A1a, A1b, D2)
            got = eng.energy(FN["PBE"], 0.0)
            fresh, _ = run(ordm, w, ph, gr, ncore, A1a, A1b, D2)
            check_scalars(got, fresh, tol=1e-13)
            if nact <= 13:
                want = port.energy_coreactive(w, ph, ncore, A1a, A1b, D2, "PBE", gr, eclass=0.0, want_vecs=False)
                check_scalars(got, want.scalars)
    finally:
        eng.close()

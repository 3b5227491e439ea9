"""Accuracy study (CPU, numpy) for an Ozaki-style int8 slicing of K2's contraction Y = X . D~:
how many 7-bit slices of X rows and D~ columns does Pi = rowsum(X o Y) need to stay within the parity bar
(|dPi| <= 1e-12 max(1, |Pi|)) on the bench's synthetic config-4-like data?  Lab tool, not product code."""
import sys, numpy as np
sys.path.insert(0, ".")
import openrdm_b200 as ob

def folded(D2, n):
    idx = [(t, u) for t in range(n) for u in range(t, n)]
    M = len(idx)
    D4 = D2.reshape(n, n, n, n)
    F = np.zeros((M, M))
    for I, (t, u) in enumerate(idx):
        rows = {(t, u), (u, t)}
        for J, (v, w) in enumerate(idx):
            cols = {(v, w), (w, v)}
            F[I, J] = sum(D4[a, b, c, d] for (a, b) in rows for (c, d) in cols)
    return 0.5 * (F + F.T), idx

def slices(A, axis, S):
    """7-bit round-to-nearest slices along rows (axis=1: one exponent per row) or columns (axis=0)."""
    amax = np.abs(A).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)  # |A| / 2^e < 1
    r = A / np.exp2(e)
    out = []
    for s in range(S):
        q = np.rint(r * np.exp2(6 + 7 * s))
        assert np.abs(q).max() <= 64
        out.append(q.astype(np.int8))
        r = r - q * np.exp2(-(6 + 7 * s))
    return out, e

def main():
    nact, ncore, npts = int(sys.argv[1]) if len(sys.argv) > 1 else 20, 83, 2048
    A1a, A1b, D2 = ob.synth_rdms(7, nact, nact // 2, nact // 2)
    Dt, idx = folded(np.asarray(D2), nact)
    w, phi, _ = ob.synth_fill_host(11, 5_000_000, 123_456, npts, ncore + nact, False)
    pa = phi[:, ncore:]
    X = np.stack([pa[:, t] * pa[:, u] for (t, u) in idx], axis=1)
    Xl, Dl = X.astype(np.longdouble), Dt.astype(np.longdouble)
    pi_ref = ((Xl @ Dl) * Xl).sum(axis=1)
    pi_f64 = ((X @ Dt) * X).sum(axis=1)
    scale = np.maximum(1.0, np.abs(pi_ref).astype(np.float64))
    print(f"nact={nact} M={len(idx)} |X|max={np.abs(X).max():.3g} |D~|max={np.abs(Dt).max():.3g} |Pi|max={np.abs(pi_ref).max():.3g}")
    print(f"fp64 gemm vs long double: max |dPi|/max(1,|Pi|) = {np.abs((pi_f64 - pi_ref).astype(np.float64) / scale).max():.3e}")
    for S in (6, 7, 8, 9):
        xs, ex = slices(X, 1, S)
        ds, ed = slices(Dt, 0, S)
        for G in (S - 1, S, S + 1):          # keep slice pairs with s + t <= G
            Y = np.zeros(X.shape, dtype=np.longdouble); npair = 0
            for s in range(S):
                for t in range(S):
                    if s + t > G: continue
                    npair += 1
                    acc = xs[s].astype(np.int64) @ ds[t].astype(np.int64)       # exact (int32 on the tensor core)
                    assert np.abs(acc).max() < 2**31
                    Y += acc.astype(np.longdouble) * np.exp2(np.longdouble(-(12 + 7 * (s + t))))
            Y = Y * np.exp2(ex).astype(np.longdouble) * np.exp2(ed).astype(np.longdouble)
            pi = (Y.astype(np.float64) * X).sum(axis=1)
            err = np.abs((pi - pi_ref).astype(np.float64) / scale).max()
            print(f"S={S} pairs(s+t<={G})={npair:3d}: max |dPi|/max(1,|Pi|) = {err:.3e}")

main()

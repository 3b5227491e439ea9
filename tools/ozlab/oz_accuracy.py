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

if not (len(sys.argv) > 2 and sys.argv[2] == 'bytes'):
    main()


BALANCED = len(sys.argv) > 3 and sys.argv[3] == 'balanced'
HEAD = 1 if BALANCED else 0   # one bit of headroom so that the carry cannot overflow the top digit


def byte_slices(A, axis, S):
    """Two's-complement byte slicing of a (8 S)-bit fixed-point image: top byte signed, the rest unsigned;
    pure bit extraction on the device (no per-digit arithmetic)."""
    amax = np.abs(A).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.floor(np.log2(np.where(amax > 0, amax, 1.0))) + 1, 0.0)
    V = np.rint(A / np.exp2(e) * np.exp2(8 * S - 1 - HEAD)).astype(np.int64)
    V = np.clip(V, -(1 << (8 * S - 1)), (1 << (8 * S - 1)) - 1)
    if BALANCED:   # signed digits in [-128, 127]: add 0x80 to every lower byte (carries ripple up), then re-centre
        V = V + sum(128 << (8 * k) for k in range(S - 1))
        out = [(((V >> (8 * k)) & 255) - 128) for k in range(S - 1)] + [V >> (8 * (S - 1))]
        assert max(int(np.abs(o).max()) for o in out) <= 128 and int(out[-1].max()) <= 127
    else:
        out = [((V >> (8 * k)) & 255) for k in range(S - 1)] + [V >> (8 * (S - 1))]
    return out, e


def main_bytes(nact=20, npts=2048, ncore=83):
    A1a, A1b, D2 = ob.synth_rdms(7, nact, nact // 2, nact // 2)
    Dt, idx = folded(np.asarray(D2), nact)
    w, phi, _ = ob.synth_fill_host(11, 5_000_000, 123_456, npts, ncore + nact, False)
    pa = phi[:, ncore:]
    X = np.stack([pa[:, t] * pa[:, u] for (t, u) in idx], axis=1)
    Xl, Dl = X.astype(np.longdouble), Dt.astype(np.longdouble)
    pi_ref = ((Xl @ Dl) * Xl).sum(axis=1)
    scale = np.maximum(1.0, np.abs(pi_ref).astype(np.float64))
    for S in (5, 6, 7):
        xs, ex = byte_slices(X, 1, S)
        ds, ed = byte_slices(Dt, 0, S)
        for drop in (S - 2, S - 1, S):          # keep byte pairs with k + l >= drop
            acc = np.zeros(X.shape, dtype=object); npair = 0; amax = 0
            groups = {}
            for k in range(S):
                for l in range(S):
                    if k + l < drop: continue
                    npair += 1
                    groups.setdefault(k + l, np.zeros(X.shape, dtype=np.int64))
                    groups[k + l] += xs[k] @ ds[l]
            Y = np.zeros(X.shape, dtype=np.longdouble)
            for gsum, a in groups.items():
                amax = max(amax, int(np.abs(a).max()))
                Y += a.astype(np.longdouble) * np.exp2(np.longdouble(8 * gsum - 2 * (8 * S - 1 - HEAD)))
            Y = Y * np.exp2(ex).astype(np.longdouble) * np.exp2(ed).astype(np.longdouble)
            pi = (Y.astype(np.float64) * X).sum(axis=1)
            err = np.abs((pi - pi_ref).astype(np.float64) / scale).max()
            print(f"bytes S={S} pairs(k+l>={drop})={npair:3d} groups={len(groups)} max|acc|=2^{np.log2(max(amax,1)):.1f}: max |dPi|/max(1,|Pi|) = {err:.3e}")


if len(sys.argv) > 2 and sys.argv[2] == "bytes":
    main_bytes(int(sys.argv[1]))

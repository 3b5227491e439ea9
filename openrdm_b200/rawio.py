"""openrdm_b200.rawio -- writer of the raw MC-PDFT input container (format: csrc/rawio.h) and converter from the
PySCF exporter's data.h5 (/root/reference/src/libpyscf/libpyscf.py:613-690).  The C reader behind
ordm_file_open / ordm_load_data is the product; this module stands where h5py stands on the producing side:
`write_container(path, {"/GRIDS/W": w, "/PHI/PHI_MO": phi, ...})` mirrors `f[name] = array`.

`make_data_h5_dict` builds the exporter's dataset dictionary (same names, shapes and packing rules) from plain arrays
-- what libpyscf.MCPDFT.kernel writes, minus the quantities that need PySCF itself."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"ORDMRAW1"
_DT = {np.dtype("float64"): 0, np.dtype("int64"): 1, np.dtype("int32"): 2}


def write_container(path: str, datasets: dict) -> None:
    """datasets: name -> array-like (float64 / int64 / int32; anything else is converted to float64 or int64)."""
    items = []
    for name, a in datasets.items():
        a = np.asarray(a)
        if a.dtype not in _DT:
            a = a.astype(np.int64 if np.issubdtype(a.dtype, np.integer) else np.float64)
        if a.ndim > 4:
            raise ValueError(f"{name}: at most 4 dimensions")
        items.append((name.encode(), a if a.flags.c_contiguous else np.ascontiguousarray(a)))   # row-major, as h5py stores numpy arrays (0-d stays 0-d)
    toc_len = sum(4 + len(n) + 4 + 4 + 8 * a.ndim + 8 + 8 for n, a in items)
    off = (24 + toc_len + 63) // 64 * 64
    toc, offsets = b"", []
    for n, a in items:
        toc += struct.pack("<I", len(n)) + n + struct.pack("<II", _DT[a.dtype], a.ndim)
        toc += b"".join(struct.pack("<Q", d) for d in a.shape) + struct.pack("<QQ", off, a.nbytes)
        offsets.append(off)
        off = (off + a.nbytes + 63) // 64 * 64
    assert len(toc) == toc_len
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<QQ", len(items), toc_len) + toc)
        for (n, a), o in zip(items, offsets):
            f.seek(o)
            f.write(a.tobytes())
        f.truncate(max(off, 24 + toc_len))


def convert_h5(h5_path: str, out_path: str) -> None:
    """data.h5 -> raw container (needs h5py; every dataset is copied under its HDF5 path)."""
    import h5py
    sets = {}
    with h5py.File(h5_path, "r") as f:
        f.visititems(lambda name, obj: sets.__setitem__("/" + name, obj[()]) if isinstance(obj, h5py.Dataset) else None)
    write_container(out_path, sets)


def _coo_upper_1(D, tol=1.0e-20):
    """write_rdm1s_d2sp_coo (libpyscf.py:246-328): nonzero elements with i <= j, row by row."""
    n = D.shape[0]
    r, c, v = [], [], []
    for i in range(n):
        for j in range(i, n):
            if abs(D[i, j]) > tol:
                r.append(i); c.append(j); v.append(D[i, j])
    return np.array(v, dtype=np.float64), np.array(r, dtype=np.int64), np.array(c, dtype=np.int64)


def _coo_upper_2(D4, tol=1.0e-20):
    """write_rdm2s_d2sp_coo (libpyscf.py:330-463): nonzero dm2[i, j, k, l] with i <= j and k <= l."""
    n = D4.shape[0]
    iu = np.triu_indices(n)
    sub = D4[iu[0][:, None], iu[1][:, None], iu[0][None, :], iu[1][None, :]]    # [pair(i<=j)][pair(k<=l)], same loop order
    keep = np.abs(sub) > tol
    a, b = np.nonzero(keep)
    return sub[keep].astype(np.float64), iu[0][a].astype(np.int64), iu[1][a].astype(np.int64), iu[0][b].astype(np.int64), iu[1][b].astype(np.int64)


def make_data_h5_dict(w, phi_mo=None, grads_mo=None, ncore=0, A1a=None, A1b=None, dm2ab=None, phi_ao=None, grads_ao=None, C=None,
                      e_core=0.0, e_hartree=0.0, coords=None, dense_rdms=True, sparse_rdms=True):
    """The exporter's datasets from plain arrays.  phi_*: (npts, n) ; dm2ab: (nact, nact, nact, nact) with
    dm2ab[i, j, k, l] multiplying phi_i phi_j (alpha) phi_k phi_l (beta), as PySCF's casdm2ab."""
    nact = A1a.shape[0]
    d = {"/N/N_COR": np.int64(ncore), "/N/N_CAS_ORB": np.int64(nact), "/E/E_CORE": np.float64(e_core), "/E/E_HARTREE": np.float64(e_hartree),
         "/GRIDS/W": np.asarray(w, dtype=np.float64)}
    if coords is not None:
        d["/GRIDS/X"], d["/GRIDS/Y"], d["/GRIDS/Z"] = (np.ascontiguousarray(coords[:, k]) for k in range(3))
    for tag, phi, grads in (("MO", phi_mo, grads_mo), ("AO", phi_ao, grads_ao)):
        if phi is None:
            continue
        d[f"/PHI/PHI_{tag}"] = np.ascontiguousarray(phi, dtype=np.float64)
        d[f"/N/N_{tag}"] = np.int64(phi.shape[1])
        if grads is not None:
            for s, g in zip("XYZ", grads):
                d[f"/PHI/PHI_{tag}_{s}"] = np.ascontiguousarray(g, dtype=np.float64)
    if C is not None:
        d["/C"] = np.ascontiguousarray(C, dtype=np.float64)
    if dense_rdms:
        d["/D/D1/ACT_D1A_MO"], d["/D/D1/ACT_D1B_MO"] = np.ascontiguousarray(A1a), np.ascontiguousarray(A1b)
        d["/D/D2/ACT_D2AB_MO"] = np.ascontiguousarray(dm2ab)
    if sparse_rdms:
        for tag, A in (("a", A1a), ("b", A1b)):
            v, r, c = _coo_upper_1(np.asarray(A))
            base = f"/SP_SYMM_D1/ACT_D1{tag}_MO/"
            d[base + "NNZ"], d[base + "VAL"], d[base + "ROW_IDX"], d[base + "COL_IDX"] = np.int64(len(v)), v, r, c
        v, i, j, k, l = _coo_upper_2(np.asarray(dm2ab))
        base = "/SP_SYMM_D2/ACT_D2ab_MO/"
        d[base + "NNZ"], d[base + "VAL"] = np.int64(len(v)), v
        d[base + "DIM1_IDX"], d[base + "DIM2_IDX"], d[base + "DIM3_IDX"], d[base + "DIM4_IDX"] = i, j, k, l
    return d


if __name__ == "__main__":   # python -m openrdm_b200.rawio data.h5 data.ordmraw   (needs h5py)
    import sys
    if len(sys.argv) != 3:
        sys.exit("usage: python -m openrdm_b200.rawio <data.h5> <out.ordmraw>")
    convert_h5(sys.argv[1], sys.argv[2])

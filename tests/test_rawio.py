"""Input files: the PySCF exporter's data.h5 schema through ordm_file_* / ordm_load_data.
CPU part: the raw container written by openrdm_b200/rawio.py is read back bit for bit by the C reader, the exporter's
packing rules are reproduced, bad files fail loudly.  GPU part: a context filled from a file gives the same energy as
one filled through the ordm_set_* calls -- MO and AO form, sparse and dense RDMs -- and the oracle's."""
import os

import numpy as np
import pytest

from helpers import check_scalars, check_vectors


def case(ordm, rng, npts=700, ncore=2, nact=5, nao=11):
    n = ncore + nact
    C = np.linalg.qr(rng.normal(size=(nao, nao)))[0][:, :n + 2]              # nao x nmo (two virtuals behind the occupied block)
    ao = rng.uniform(-0.6, 0.6, (npts, nao)) * np.where(rng.random(npts) < 0.9, 1.0, 2.0 ** -rng.integers(0, 40, npts))[:, None]
    gao = [rng.uniform(-0.6, 0.6, (npts, nao)) for _ in range(3)]
    phi, gphi = ao @ C, [g @ C for g in gao]
    w = rng.random(npts) / npts
    A1a, A1b, D2 = ordm.synth_rdms(11, nact, (nact + 1) // 2, nact // 2, 4)
    # reference convention D2ab[(nu n + mu), (sigma n + lambda)] (alpha pair (nu, sigma), beta pair (mu, lambda), mcpdft.cc:204,640-652)
    # -> PySCF's casdm2ab[i, j, k, l] = alpha pair (i, j), beta pair (k, l)
    dm2ab = np.ascontiguousarray(np.asarray(D2).reshape(nact, nact, nact, nact, order="F").transpose(1, 3, 0, 2))
    return dict(w=w, ao=ao, gao=gao, C=C, phi=phi, gphi=gphi, A1a=np.asarray(A1a), A1b=np.asarray(A1b), D2=np.asarray(D2), dm2ab=dm2ab,
                ncore=ncore, nact=nact, n=n)


def test_container_round_trip_and_packing_rules(tmp_path, ordm):
    from openrdm_b200 import rawio
    rng = np.random.default_rng(5)
    c = case(ordm, rng)
    d = rawio.make_data_h5_dict(c["w"], c["phi"], c["gphi"], c["ncore"], c["A1a"], c["A1b"], c["dm2ab"], c["ao"], c["gao"], c["C"], e_core=-3.5, e_hartree=1.25)
    path = str(tmp_path / "data.ordm")
    rawio.write_container(path, d)
    f = ordm.InputFile(path)
    try:
        for name, a in d.items():
            got = f.dataset(name)
            a = np.asarray(a)
            assert got.shape == a.shape and np.array_equal(got, a), name
        with pytest.raises(KeyError):
            f.dataset("/NOPE")
    finally:
        f.close()
    # the exporter's COO rules (libpyscf.py:246-463): i <= j (and k <= l), |value| > 1e-20, values unscaled
    v, i, j, k, l = (d["/SP_SYMM_D2/ACT_D2ab_MO/" + s] for s in ("VAL", "DIM1_IDX", "DIM2_IDX", "DIM3_IDX", "DIM4_IDX"))
    assert (i <= j).all() and (k <= l).all() and np.array_equal(v, c["dm2ab"][i, j, k, l]) and int(d["/SP_SYMM_D2/ACT_D2ab_MO/NNZ"]) == len(v)
    assert np.allclose(c["dm2ab"], c["dm2ab"].transpose(1, 0, 2, 3), atol=1e-15) and np.allclose(c["dm2ab"], c["dm2ab"].transpose(0, 1, 3, 2), atol=1e-15)
    back = ordm.rdm2ab_sparse_to_dense(c["nact"], (i, j, k, l, v), upper_packed=True)        # the packed list, expanded by symmetry ...
    na = c["nact"]
    full = np.array(np.meshgrid(*[np.arange(na)] * 4, indexing="ij")).reshape(4, -1)
    want = ordm.rdm2ab_sparse_to_dense(na, (full[0], full[1], full[2], full[3], c["dm2ab"].reshape(-1)), upper_packed=False)   # ... equals every element listed once
    assert np.allclose(back, want, atol=1e-15)


def test_bad_files_fail_loudly(tmp_path, ordm):
    p = str(tmp_path / "junk.bin")
    open(p, "wb").write(b"not a container at all, just bytes" * 4)
    with pytest.raises(ordm.OrdmError):
        ordm.InputFile(p)
    open(p, "wb").write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)       # an HDF5 signature: needs the HDF5 build of the library
    with pytest.raises(ordm.OrdmError) as ei:
        ordm.InputFile(p)
    assert "HDF5" in str(ei.value)
    with pytest.raises(ordm.OrdmError):
        ordm.InputFile(str(tmp_path / "missing.ordm"))
    from openrdm_b200 import rawio
    rawio.write_container(p, {"/GRIDS/W": np.ones(4)})
    raw = bytearray(open(p, "rb").read())
    raw[24 + 4 + len("/GRIDS/W") + 8 + 8:24 + 4 + len("/GRIDS/W") + 8 + 16] = (1 << 40).to_bytes(8, "little")   # offset beyond the file
    open(p, "wb").write(raw)
    with pytest.raises(ordm.OrdmError):
        ordm.InputFile(p)


@pytest.mark.gpu
@pytest.mark.parametrize("form,flags", [("mo-sparse", 0), ("mo-dense", 2), ("ao-sparse", 1), ("lda", 4)])
def test_load_data_matches_direct_calls_and_oracle(tmp_path, ordm, port, form, flags):
    from openrdm_b200 import rawio
    rng = np.random.default_rng(17)
    c = case(ordm, rng)
    d = rawio.make_data_h5_dict(c["w"], c["phi"], c["gphi"], c["ncore"], c["A1a"], c["A1b"], c["dm2ab"], c["ao"], c["gao"], c["C"], e_core=-3.5, e_hartree=1.25)
    if form == "ao-sparse":
        for k in [k for k in d if k.startswith("/PHI/PHI_MO")]:
            del d[k]
    path = str(tmp_path / "data.ordm")
    rawio.write_container(path, d)
    gga = form != "lda"
    fn = "PBE" if gga else "SVWN"
    n = c["n"]
    phi = np.asfortranarray(c["phi"][:, :n]); grads = [np.asfortranarray(g[:, :n]) for g in c["gphi"]] if gga else None
    want = port.energy_coreactive(c["w"], phi, c["ncore"], c["A1a"], c["A1b"], c["D2"], fn, grads, eclass=-2.25)
    eng = ordm.Engine(0)
    try:
        eng.set_options(diagnostics=True)
        info = eng.load_data(path, flags)
        assert info["npts"] == len(c["w"]) and info["ncore"] == c["ncore"] and info["nact"] == c["nact"] and info["gga"] == int(gga)
        assert info["used_ao"] == int(form == "ao-sparse") and info["used_sparse_rdm"] == int(form != "mo-dense")
        assert info["e_core"] == -3.5 and info["e_hartree"] == 1.25
        got = eng.energy(ordm.FUNC_ID[fn], info["e_core"] + info["e_hartree"])
        check_scalars(got, want.scalars)
        check_vectors(eng.get_vector, want.vecs, gga)
    finally:
        eng.close()

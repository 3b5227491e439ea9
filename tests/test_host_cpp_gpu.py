"""GPU test of the C++ host mirror: OpenRDM's own test driver flow (tests/main.cc) run through
openrdm_b200/host/mcpdft_test on the real H2/STO-3G fixture (BASELINE.json configs[0] and [1]),
text files written in the reference's format (tests/h2_tsvwn_sto3g/README.txt)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "openrdm_b200", "host", "mcpdft_test")
# /root/reference/tests/main.cc compiled byte-for-byte unmodified against host/compat/{armadillo,mcpdft.h,energy.h}
# (openrdm_b200/host/Makefile, target mcpdft_test_ref; built where /root/reference exists, travels to the GPU box)
BIN_REF = os.path.join(ROOT, "openrdm_b200", "host", "mcpdft_test_ref")


def write_case(d, g, grads, eref):
    os.makedirs(d)
    phi, w = g["phi"], g["w"]
    npts, n = phi.shape
    with open(os.path.join(d, "grids.txt"), "w") as f:  # "w x y z" rows, utility.cc:5-34
        f.write(f"{npts}\n")
        for p in range(npts):
            f.write(f"{w[p]:.17e} 0.0 0.0 0.0\n")
    with open(os.path.join(d, "orbitals.txt"), "w") as f:
        f.write(f"{npts} {n}\n")
        np.savetxt(f, phi, fmt="%.17e")
    if grads is not None:
        with open(os.path.join(d, "gradients.txt"), "w") as f:  # phi_x phi_y phi_z interleaved per orbital
            f.write(f"{npts} {n}\n")
            inter = np.stack(grads, axis=2).reshape(npts, 3 * n)
            np.savetxt(f, inter, fmt="%.17e")
    with open(os.path.join(d, "opdm.txt"), "w") as f:
        f.write(f"{n} {n}\n")
        np.savetxt(f, g["D1a"], fmt="%.17e")
    open(os.path.join(d, "eclass.txt"), "w").write(f"{float(g['eclass']):.17e}\n")
    open(os.path.join(d, "eref.txt"), "w").write(f"{eref:.17e}\n")


@pytest.mark.parametrize("case,fn,pre", [("h2_tsvwn_sto3g", "SVWN", "svwn_"), ("h2_tpbe_sto3g", "PBE", "pbe_")])
def test_reference_driver_flow(tmp_path, golden_h2, ordm, case, fn, pre):
    if not os.path.exists(BIN):
        subprocess.run(["make", "-s", "-C", os.path.dirname(BIN)], check=True)
    g = golden_h2
    npts, n = g["phi"].shape
    grads = ordm.synth_fill_host(int(g["grad_seed"]), npts, 0, npts, n, True)[2] if fn == "PBE" else None
    want = float(g[pre + "e_tot"])  # what the unmodified reference returns on these inputs
    write_case(str(tmp_path / case), g, grads, want)
    r = subprocess.run([BIN, case, fn, "1e-10"], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    e = float(re.search(r"MCPDFT energy\s+=\s+(\S+)", r.stdout).group(1))
    assert abs(e - want) < 1e-10
    ex = float(re.search(r"Ex\s+=\s+(\S+)", r.stdout).group(1))
    assert abs(ex - float(g[pre + "ex"])) < 1e-10


@pytest.mark.parametrize("case,fn,pre", [("h2_tsvwn_sto3g", "SVWN", "svwn_"), ("h2_tpbe_sto3g", "PBE", "pbe_")])
def test_unmodified_reference_main_cc_against_the_mirror(tmp_path, golden_h2, ordm, case, fn, pre):
    """The reference's own tests/main.cc (no edits) linked to the B200 host mirror reproduces, on the H2 fixtures,
    the energy that the same main.cc linked to the reference's own library returns (golden, oracle/_ref)."""
    if not os.path.exists(BIN_REF):
        subprocess.run(["make", "-s", "-C", os.path.dirname(BIN_REF), "mcpdft_test_ref"], check=True)
    if not os.path.exists(BIN_REF):
        pytest.skip("mcpdft_test_ref not built (needs /root/reference/tests/main.cc at build time)")
    g = golden_h2
    npts, n = g["phi"].shape
    grads = ordm.synth_fill_host(int(g["grad_seed"]), npts, 0, npts, n, True)[2] if fn == "PBE" else None
    want = float(g[pre + "e_tot"])
    write_case(str(tmp_path / case), g, grads, want)
    r = subprocess.run([BIN_REF, case, fn], cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    e = float(re.search(r"MCPDFT energy\s+=\s+(\S+)", r.stdout).group(1))   # printed with 12 decimals (tests/main.cc:43)
    assert abs(e - want) < 1e-11 + 5e-13
    d = float(re.search(r"E\(MCPDFT\) - E\(Ref\)\s+=\s+(\S+)", r.stdout).group(1))
    assert abs(d) < 1e-10

"""Runs the kernel variants the default bench does not reach (for ncu): the grad-Pi / fully-translated
path on config-4 shapes and the wide-tile K2 of config 3.  Prints their event times."""
import sys
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import openrdm_b200 as ordm

def run(tag, npts, ncore, nact, nel, seed, **opts):
    n = ncore + nact
    eng = ordm.Engine(0)
    A1a, A1b, D2 = ordm.synth_rdms(seed, nact, nel, nel, 8)
    eng.set_options(**opts)
    eng.set_active_space(ncore, A1a, A1b, D2)
    eng.synth_fill(seed, npts, 0, npts, n, True)
    for _ in range(3):
        r = eng.energy(ordm.FUNC_PBE, 0.0)
    print(f"{tag}: total {r['ms_total']:.3f} ms  density {r['ms_density']:.3f}  pi {r['ms_pi']:.3f}  xc {r['ms_xc']:.3f}  E = {r['e_tot']:.12f}")
    eng.close()

run("config 4, 1M points, ft-tPBE (Pi + grad Pi, full Dt)", 1_000_000, 83, 20, 10, 4, fully_translated=True)
run("config 4, 1M points, tPBE", 1_000_000, 83, 20, 10, 4)
run("config 3, 500k points, tPBE (wide-tile K2)", 500_000, 2, 8, 5, 3)

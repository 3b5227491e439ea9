"""Full-grid golden energies of BASELINE.json's configs 4 and 5 from OpenRDM's own code.

    python tests/golden/make_golden_fullgrid.py [--configs 4,5] [--threads N]

Runs oracle/_ref (the reference's unmodified libmcpdft/libfunc sources, oracle/Makefile; energy.cc:27-188 staged
as mcpdft_energy runs it, minus the grad-Pi stage whose output never reaches the energy) over

* config 4: the WHOLE 5 000 000-point grid (83 core + 20 active orbitals, tPBE, seed 4);
* config 5: the first of eight shards of the 20 000 000-point grid = ordm_shard_range(20e6, 8, 0) = points
  [0, 2 500 000) (73 core + 26 active orbitals, seed 5), tSVWN and tPBE -- what rank 0 of the 8-GPU run holds;

in chunks of CHUNK = 65 536 points (SURVEY.md section 7 "Summation order": the reference's own serial sum over 2e7
terms carries ~1e-11 Eh of rounding noise; within a chunk each thread of ref_energy_threads sums its <= 8 192-point
slice serially as the reference does, chunks are combined in long double here).  The inputs are regenerated chunk by
chunk with the counter-based generator (include/ordm_synth.h, host-only build openrdm_b200/host/libordm_synth.so), so
the GPU test regenerates the same bits on the device.  Writes tests/golden/fullgrid.json: the totals and the
per-chunk partial sums (so that a test can also pin any chunk-aligned sub-range).

Needs /root/reference at build time only (oracle/_ref); cannot run on the GPU box's budget -- the JSON is committed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from openrdm_b200 import synth  # noqa: E402  (host-only generator; no CUDA library is mapped)

CHUNK = 65536
KEYS = ("ex", "ec", "n_tot", "n_a", "n_b", "trn_tot", "trn_a", "trn_b")
CASES = {
    "config4_tPBE_full": dict(seed=4, npts_total=5_000_000, p0=0, count=5_000_000, ncore=83, nact=20, nelec=10, functional="PBE"),
    "config5_tSVWN_shard0of8": dict(seed=5, npts_total=20_000_000, p0=0, count=2_500_000, ncore=73, nact=26, nelec=13, functional="SVWN"),
    "config5_tPBE_shard0of8": dict(seed=5, npts_total=20_000_000, p0=0, count=2_500_000, ncore=73, nact=26, nelec=13, functional="PBE"),
}


def run_case(ref, name, c, threads):
    n = c["ncore"] + c["nact"]
    gga = c["functional"] != "SVWN"
    A1a, A1b, D2 = synth.synth_rdms(c["seed"], c["nact"], c["nelec"], c["nelec"], 8)
    parts = {k: [] for k in KEYS}
    t0 = time.time()
    for b in range(c["p0"], c["p0"] + c["count"], CHUNK):
        m = min(CHUNK, c["p0"] + c["count"] - b)
        w, phi, grads = synth.synth_fill_host(c["seed"], c["npts_total"], b, m, n, gga)
        r = ref.energy_threads(threads, w, phi, c["ncore"], A1a, A1b, D2, c["functional"], grads, 0.0, with_pi_grad=False)
        for k in KEYS:
            parts[k].append(r.scalars[k])
        if (len(parts["ex"]) % 8) == 0:
            print(f"  {name}: {b + m - c['p0']}/{c['count']} points, {time.time() - t0:.0f} s", flush=True)
    tot = {k: float(np.sum(np.array(v, dtype=np.longdouble))) for k, v in parts.items()}
    tot["e_xc"] = float(np.longdouble(tot["ex"]) + np.longdouble(tot["ec"]))
    return dict(case=c, chunk_points=CHUNK, threads=threads, seconds=time.time() - t0, totals=tot, chunks=parts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="4,5")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(HERE, "fullgrid.json"))
    args = ap.parse_args()
    want = {int(x) for x in args.configs.split(",")}
    ref = O.Ref()
    out = {}
    if os.path.exists(args.out):
        out = json.load(open(args.out))
    out["_generator"] = "tests/golden/make_golden_fullgrid.py (oracle/_ref = OpenRDM's unmodified sources; see the script's docstring)"
    for name, c in CASES.items():
        if int(name[6]) not in want:
            continue
        print(name, flush=True)
        out[name] = run_case(ref, name, c, args.threads)
        print("  totals:", out[name]["totals"], flush=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

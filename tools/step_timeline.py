"""Lab helper: config 4 on one GPU, synchronous steps (ordm_energy) against stream-ordered ones (ordm_energy_async x K, one
wait): where the step's time goes and what the enqueue order costs.
    python tools/step_timeline.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import openrdm_b200 as O  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
npts, ncore, nact, seed = 5_000_000, 83, 20, 2024
A1a, A1b, D2 = O.synth_rdms(seed, nact, 10, 10, 8)
eng = O.Engine(0)
eng.set_options(diagnostics=False)
eng.set_active_space(ncore, A1a, A1b, D2)
eng.synth_fill(seed, npts, 0, npts, ncore + nact, True)
for _ in range(3):
    r = eng.energy(O.FUNC_PBE, 0.0)
keys = ("ms_density", "ms_pi", "ms_xc", "ms_reduce", "ms_total")
acc = {k: 0.0 for k in keys}
t0 = time.time()
for _ in range(steps):
    r = eng.energy(O.FUNC_PBE, 0.0)
    for k in keys:
        acc[k] += r[k] / steps
t_sync = (time.time() - t0) / steps * 1e3
print("synchronous steps: host wall %.3f ms/step; events:" % t_sync, {k: round(v, 3) for k, v in acc.items()}, "overlapped", r["overlapped"], "tiled", r["density_tiled"])
torch.cuda.synchronize()
t0 = time.time()
for _ in range(steps):
    eng.energy_async(O.FUNC_PBE, 0.0)
r = eng.energy_wait()
t_async = (time.time() - t0) / steps * 1e3
print("stream-ordered steps: host wall %.3f ms/step; last step's events:" % t_async, {k: round(r[k], 3) for k in keys})
eng.close()

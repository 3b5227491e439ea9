import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import openrdm_b200 as ordm
cfg = dict(npts=5_000_000, ncore=83, nact=20, nelec=10, seed=4)
n = cfg["ncore"] + cfg["nact"]
eng = ordm.Engine(0)
A1a, A1b, D2 = ordm.synth_rdms(cfg["seed"], cfg["nact"], cfg["nelec"], cfg["nelec"], 8)
eng.set_active_space(cfg["ncore"], A1a, A1b, D2)
eng.synth_fill(cfg["seed"], cfg["npts"], 0, cfg["npts"], n, True)
cnt = cfg["npts"]
w_h = ordm.host_empty((cnt,)); phi_h = ordm.host_empty((cnt, n)); g_h = [ordm.host_empty((cnt, n)) for _ in range(3)]
w_h[:] = eng.get_vector("w"); eng.get_orbitals(phi_h, g_h)
for batch in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
    eng.energy_host(1, 0.0, w_h, phi_h, g_h, batch_points=batch)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(2): r = eng.energy_host(1, 0.0, w_h, phi_h, g_h, batch_points=batch)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 2
    print(batch, round(dt * 1e3, 2), "ms", round(16.52 / dt, 2), "GB/s", r["e_tot"])

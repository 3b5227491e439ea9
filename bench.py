#!/usr/bin/env python
"""bench.py -- MC-PDFT grid points/s (rho + Pi + translated XC + quadrature) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 3|4|5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one evaluation of the whole hot path (mcpdft::mcpdft_energy, energy.cc:27-188) over
the synthetic grid of the chosen BASELINE.json config, inputs resident in HBM.  With N ranks the
fixed grid is sharded by contiguous point ranges (strong scaling), the RDMs are replicated and
one NCCL all-reduce combines the eight scalars.  Prints ONE JSON line on rank 0.

    value     grid points/s, device-resident inputs, CUDA events, max over ranks
    e2e       same metric through ordm_energy_host(): inputs in pinned HOST memory, H2D streamed
              inside the timed region, result scalars read back every step
    roofline  the dominant kernel (K2, on-top pair density, FP64 tensor pipe) against the cuBLAS
              DGEMM rate measured in this process; K1/K3 against MEASURED_PEAKS.json's HBM GB/s
    cpu_baseline   OpenRDM's own compiled code (oracle/_ref) on a bounded sample, all host cores

--impl reference times only the reference CPU arm (rank 0), same metric/config (its inputs come from the host-only
generator openrdm_b200/host/libordm_synth.so: the GPU library is not mapped in that arm).

Also in the line: `parity` (config 4: E_xc of the timed run against the full-grid golden of OpenRDM's own code,
tests/golden/fullgrid.json, asserted to 1e-10 Eh at every N), `value_fp64_dmma` (the same step with Pi from the FP64 DMMA
kernels, ORDM_K2_INT8=0), `e2e.grid_resident` (grid resident, RDMs re-uploaded and result read back every step: the
SCF-loop use) and `config5` (BASELINE.json's largest config, 20 M points / 26 active orbitals, tSVWN and tPBE, same N).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MCPDFT grid points/sec (rho+Pi+tXC+quadrature)"
UNIT = "grid points/s"
# the arithmetic type of the path (not a precision claim): FP64 throughout; for 18..20 active orbitals the products of the
# on-top contraction are formed exactly on the int8 tensor cores (roofline.pi_arithmetic says which kernel ran)
DTYPE = "f64"

# BASELINE.json configs 3-5 ("Config sizes")
CONFIGS = {
    3: dict(name="synthetic N2 cc-pVTZ (10e,8o), 500k-point grid, tPBE", npts=500_000, ncore=2, nact=8, nelec=5, functional="PBE", seed=3),
    4: dict(name="synthetic Fe-porphyrin-like (20e,20o), def2-TZVP-size basis, 5M-point grid, tPBE", npts=5_000_000, ncore=83, nact=20, nelec=10, functional="PBE", seed=4),
    5: dict(name="synthetic polyacene (26e,26o), 20M-point grid, tPBE", npts=20_000_000, ncore=73, nact=26, nelec=13, functional="PBE", seed=5),
}


def algorithmic(cfg):
    """Per-point algorithmic work."""
    n = cfg["ncore"] + cfg["nact"]
    M = cfg["nact"] * (cfg["nact"] + 1) // 2
    gga = cfg["functional"] == "PBE"
    return dict(
        M=M,
        k2_flops=M * (M + 1) + 2 * M,            # symmetric packed contraction: each unordered (I,J) once + row dot
        k2_flops_survey=2 * M * M + 2 * M,       # SURVEY.md 8(d) count for the unsymmetrised GEMM row
        k1_bytes=8 * n * (4 if gga else 1) + 8 + 8 * (5 if gga else 2),   # phi(+grad phi) + w in; rho, grad rho, pi_core out
        k3_bytes=48 if gga else 24,
    )


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self, t_begin=None, t_end=None):
        if self.proc:
            self.proc.terminate()
        rows = [r[1:] for r in self.rows if t_begin is None or (t_begin - 0.05 <= r[0] <= t_end + 0.05)]
        if len(rows) < 3:  # timed region shorter than the sampling period: use every sample under load
            rows = [r[1:] for r in self.rows]
        sm, mx, reasons, pw = [], 0.0, set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort(); pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w": pw[len(pw) // 2] if pw else None}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def dgemm_peak(torch, seconds=0.6):
    """cuBLAS DGEMM 8192^3 in this process: the FP64-tensor roofline denominator."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(3, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / reps
    del a, b
    torch.cuda.empty_cache()
    f = 2.0 * n ** 3
    return f / best * 1e-9, f / sus * 1e-9


def cpu_reference_leg(cfg, sample_pts, threads, steps=1, warmup=0):
    """OpenRDM's own compiled code (oracle/_ref) on a bounded sample of the same synthetic grid.
    Returns (points/s as written, points/s necessary-work-only, description)."""
    from oracle import oracle as O
    from openrdm_b200 import synth as ordm   # host-only generator (libordm_synth.so): no CUDA library in this arm
    if not O.have_ref():
        return None
    ref = O.Ref()
    n = cfg["ncore"] + cfg["nact"]
    gga = cfg["functional"] == "PBE"
    A1a, A1b, D2 = ordm.synth_rdms(cfg["seed"], cfg["nact"], cfg["nelec"], cfg["nelec"], 8)
    w, phi, grads = ordm.synth_fill_host(cfg["seed"], cfg["npts"], 0, sample_pts, n, gga)
    times = []
    for i in range(warmup + steps):
        r = ref.energy_threads(threads, w, phi, cfg["ncore"], A1a, A1b, D2, cfg["functional"], grads, 0.0, with_pi_grad=True)
        if i >= warmup:
            times.append(r.seconds)
    t_written = sum(times) / len(times)
    r2 = ref.energy_threads(threads, w, phi, cfg["ncore"], A1a, A1b, D2, cfg["functional"], grads, 0.0, with_pi_grad=False)
    # the reference's default build is serial (CMakeLists.txt:89): one core on 1/8 of the sample
    ns = max(64, sample_pts // 8)
    gs = [g[:ns] for g in grads] if grads is not None else None
    r1 = ref.energy_threads(1, w[:ns], phi[:ns], cfg["ncore"], A1a, A1b, D2, cfg["functional"], gs, 0.0, with_pi_grad=True)
    omp = None
    try:  # the reference's own OpenMP mode (configure:20), all host cores: slower than its serial build
        ro = O.Ref(openmp=True).energy_threads(1, w[:ns], phi[:ns], cfg["ncore"], A1a, A1b, D2, cfg["functional"], gs, 0.0, with_pi_grad=True)
        omp = ns / ro.seconds
    except Exception:
        pass
    return dict(as_written=sample_pts / t_written, necessary=sample_pts / r2.seconds, serial=ns / r1.seconds, omp=omp, seconds=t_written, e_tot=r.scalars["e_tot"],
                sample=f"first {sample_pts} points of the {cfg['npts']}-point grid, {threads} threads x contiguous slices, "
                       f"reference kernels unmodified (build_rho on all {n} orbitals; build_ontop_pair_density(+_gradients as mcpdft_energy "
                       f"does for GGA) on the {cfg['nact']} active orbitals; closed-form core terms added by the harness)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npts", type=int, default=0, help="override the grid size (debug)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): BASELINE.json's grid sharded over the ranks; weak: that grid PER rank")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config5", action="store_true", help="skip the config-5 sub-record (20 M points; needs ~64 GB per rank at N=1)")
    ap.add_argument("--no-fp64", action="store_true", help="skip the extra FP64-pure (DMMA) loop")
    ap.add_argument("--lean", action="store_true", help="profiling runs: only the resident loop (+ isolated stages): no e2e, CPU baseline, config 5, FP64-pure loop")
    ap.add_argument("--cpu-sample", type=int, default=0)
    args = ap.parse_args()

    if args.lean:
        args.no_e2e = args.no_cpu = args.no_config5 = args.no_fp64 = True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(CONFIGS[args.config])
    if args.npts:
        cfg["npts"] = args.npts
    if args.scaling == "weak":
        cfg["npts"] *= world
        cfg["name"] += f" (weak scaling: x{world} points)"
    alg = algorithmic(cfg)
    n = cfg["ncore"] + cfg["nact"]
    gga = cfg["functional"] == "PBE"
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cores = os.cpu_count() or 1
    workload = {"workload": cfg["name"], "npts": cfg["npts"], "n_occ": n, "ncore": cfg["ncore"], "nact": cfg["nact"], "M_pairs": alg["M"],
                "functional": "t" + cfg["functional"], "sharding": f"grid points, {world} contiguous shard(s), RDMs replicated",
                "l2_policy": ("inputs larger than L2 (no flush needed)" if cfg["npts"] * n * 8 * (4 if gga else 1) > 1.4e8
                              else "inputs smaller than the 126 MB L2 (parity-size case, not a bench target)")}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample or (16384 if cfg["nact"] >= 20 else 131072)
        sample = min(sample, cfg["npts"])
        leg = cpu_reference_leg(cfg, sample, cores, steps=steps, warmup=warmup)
        if leg is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
            return 0
        v = leg["as_written"]
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": leg["seconds"] * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE,
            "data": "synthetic", "config": workload,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": leg["sample"],
                             "value_necessary_work_only": leg["necessary"], "value_serial_1_core": leg["serial"],
                             "value_reference_openmp_build_all_cores": leg["omp"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import torch.distributed as dist
    import openrdm_b200 as ordm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = ordm.Engine(local_rank)
    stream = torch.cuda.Stream(device=local_rank)
    eng.set_stream(stream.cuda_stream)
    A1a, A1b, D2 = ordm.synth_rdms(cfg["seed"], cfg["nact"], cfg["nelec"], cfg["nelec"], 8)
    eng.set_options(diagnostics=False)
    eng.set_active_space(cfg["ncore"], A1a, A1b, D2)
    pik = eng.pi_kernel_info()   # int8 tensor-core K2 (18..20 active orbitals) or the FP64 DMMA kernels
    i8_pairs, i8_mma = pik["int8_slice_pairs"], pik["mma_per_tile"]
    pi_arithmetic = (f"exact int8 tensor-core products (tcgen05.mma kind::i8, CTA pairs) of {i8_pairs} byte-slice pairs of 46-bit fixed-point images; "
                     "epilogue on the integer and FP32 pipes (exact int64 sums against 27-bit orbital images + FP32 remainder terms for the four high-order groups, FP32 for the three that weigh <= 2^-32; FP64 only outside the MMA phase); guarded by an a-posteriori error estimate "
                     "with FP64 (DMMA) retry" if i8_pairs else "FP64 (DMMA.8x8x4)")
    p0, cnt = ordm.shard_range(cfg["npts"], world, rank)
    eng.synth_fill(cfg["seed"], cfg["npts"], p0, cnt, n, gga)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(ordm.Engine.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        eng.comm_attach(world, rank, bytes(uid.cpu().numpy().tobytes()))
    fid = ordm.FUNC_PBE if gga else ordm.FUNC_SVWN
    allreduce = eng.comm_info()["allreduce"] if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # warm-up: W synchronous steps; their per-kernel event times (averaged) fill `kernels` -- the timed loop below enqueues
    # its steps stream-ordered (ordm_energy_async) and reads kernel times of the last step only
    ms_k = {"density": 0.0, "pi": 0.0, "xc": 0.0, "reduce": 0.0}
    nwarm = max(warmup, 1)
    for i in range(nwarm):   # (the first step carries the kernels' one-time module load: left out of the average when W > 1)
        res = eng.energy(fid, 0.0)
        if i > 0 or nwarm == 1:
            for k in ms_k:
                ms_k[k] += res["ms_" + k] / max(nwarm - 1, 1)
    barrier()
    t_begin = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):        # K steps back to back on the engine's stream: kernels, all-reduce and the D2H read of
            eng.energy_async(fid, 0.0)   # the result scalars of every step; the host runs ahead and collects the last result
        e1.record(stream)
    res = eng.energy_wait()
    launches = res["gpu_launches"] * steps
    barrier()
    t_end = time.time()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_step = ms_total / steps
    value = cfg["npts"] / (ms_step * 1e-3)
    for k in ms_k:
        ms_k[k] = max_over_ranks(ms_k[k])
    overlapped = int(res.get("overlapped", 0))

    # ---- the same step with the stages back to back on one stream: isolated per-kernel durations
    # (in the default mode K1 streams underneath K2, so their event times overlap)
    ms_iso, ms_step_serial = None, None
    if overlapped:
        eng.set_options(diagnostics=False, serial_stages=True)
        iso_steps = max(3, min(steps, 10))
        eng.energy(fid, 0.0)
        barrier()
        ms_iso = {"density": 0.0, "pi": 0.0, "xc": 0.0, "reduce": 0.0}
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(iso_steps):
                r_iso = eng.energy(fid, 0.0)
                for k in ms_iso:
                    ms_iso[k] += r_iso["ms_" + k]
            e1.record(stream)
        barrier()
        ms_step_serial = max_over_ranks(e0.elapsed_time(e1)) / iso_steps
        for k in ms_iso:
            ms_iso[k] = max_over_ranks(ms_iso[k] / iso_steps)
        assert abs(r_iso["e_tot"] - res["e_tot"]) <= 1e-12 * max(1.0, abs(res["e_tot"])), (r_iso["e_tot"], res["e_tot"])
        eng.set_options(diagnostics=False)

    # ---- end-to-end: pinned host buffers -> ordm_energy_host, H2D inside the timed region
    e2e = None
    if not args.no_e2e:
        w_h = ordm.host_empty((cnt,)); phi_h = ordm.host_empty((cnt, n))
        g_h = [ordm.host_empty((cnt, n)) for _ in range(3)] if gga else None
        # fill the pinned buffers from the device-resident copy (same bits as the timed inputs;
        # the host generator would need minutes for 2e9 values)
        w_h[:] = eng.get_vector("w")
        eng.get_orbitals(phi_h, g_h)
        e2e_steps = max(1, min(steps, 3))
        eng.energy_host(fid, 0.0, w_h, phi_h, g_h)  # warm-up (allocates staging)
        barrier()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(e2e_steps):
                r2 = eng.energy_host(fid, 0.0, w_h, phi_h, g_h)
            e1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
        wall_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
        ms_e2e = max(ms_e2e, wall_e2e)  # host-side call time (the copies are issued from the host) bounds it from below
        h2d = 8 * cfg["npts"] * (n * (4 if gga else 1) + 1)
        e2e = {"value": cfg["npts"] / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * 8 * world,
               "ms_per_step": ms_e2e, "steps": e2e_steps, "e_tot_matches_resident": abs(r2["e_tot"] - res["e_tot"]) <= 1e-10 * max(1.0, abs(res["e_tot"])),
               "what": "ordm_energy_host: grid, orbitals and weights in pinned host memory, streamed H2D inside the timed region every step"}
        del w_h, phi_h, g_h
        # the SCF-loop use of the same API: the grid stays resident, every step uploads new RDMs (host folding / packing of the
        # 2-RDM included), runs the path and reads the result scalars back
        scf_steps = max(3, min(steps, 10))
        eng.set_active_space(cfg["ncore"], A1a, A1b, D2); eng.energy(fid, 0.0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(scf_steps):
            eng.set_active_space(cfg["ncore"], A1a, A1b, D2)
            r3 = eng.energy(fid, 0.0)
        barrier()
        ms_scf = max_over_ranks((time.perf_counter() - t0) * 1e3) / scf_steps
        e2e["grid_resident"] = {"value": cfg["npts"] / (ms_scf * 1e-3), "unit": UNIT, "ms_per_step": ms_scf, "steps": scf_steps,
                                "h2d_bytes_per_step": 8 * (2 * cfg["nact"] ** 2 + cfg["nact"] ** 4) * world, "d2h_bytes_per_step": 8 * 8 * world,
                                "timed_with": "host wall clock around ordm_set_active_space + ordm_energy (both synchronous), max over ranks",
                                "e_tot_matches_resident": abs(r3["e_tot"] - res["e_tot"]) <= 1e-12 * max(1.0, abs(res["e_tot"]))}

    # ---- parity of the timed run: config 4's E_xc against OpenRDM's own code over the same 5 M points
    parity = None
    if args.config == 4 and not args.npts and args.scaling == "strong":
        try:
            with open(os.path.join(ROOT, "tests", "golden", "fullgrid.json")) as f:
                gold = json.load(f)["config4_tPBE_full"]["totals"]
            parity = {"e_xc": res["e_tot"], "reference_e_xc": gold["e_xc"], "abs_diff": abs(res["e_tot"] - gold["e_xc"]),
                      "n_tot_abs_diff": abs(res["n_tot"] - gold["n_tot"]), "tolerance": 1e-10,
                      "reference": "oracle/_ref (OpenRDM's unmodified sources) over the whole grid, tests/golden/fullgrid.json"}
            assert parity["abs_diff"] <= 1e-10 and parity["n_tot_abs_diff"] <= 1e-10, parity
        except FileNotFoundError:
            pass
    eng.close()
    del eng

    def quick_rate(config_id, functional, k2_int8, nsteps):
        """resident rate of another configuration / kernel choice at the same N: (points/s, ms per step, e_tot)"""
        c5 = dict(CONFIGS[config_id]); c5["functional"] = functional
        g5 = functional == "PBE"
        n5 = c5["ncore"] + c5["nact"]
        old_env = os.environ.get("ORDM_K2_INT8")
        if k2_int8 is not None:
            os.environ["ORDM_K2_INT8"] = k2_int8
        e5 = ordm.Engine(local_rank)
        if old_env is None:
            os.environ.pop("ORDM_K2_INT8", None)
        else:
            os.environ["ORDM_K2_INT8"] = old_env
        try:
            e5.set_stream(stream.cuda_stream)
            a, b, d2 = ordm.synth_rdms(c5["seed"], c5["nact"], c5["nelec"], c5["nelec"], 8)
            e5.set_options(diagnostics=False)
            e5.set_active_space(c5["ncore"], a, b, d2)
            q0, qn = ordm.shard_range(c5["npts"], world, rank)
            e5.synth_fill(c5["seed"], c5["npts"], q0, qn, n5, g5)
            if world > 1:
                uid5 = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid5 = torch.frombuffer(bytearray(ordm.Engine.comm_unique_id()), dtype=torch.uint8).cuda()
                dist.broadcast(uid5, 0)
                e5.comm_attach(world, rank, bytes(uid5.cpu().numpy().tobytes()))
            f5 = ordm.FUNC_PBE if g5 else ordm.FUNC_SVWN
            for _ in range(3):
                r5 = e5.energy(f5, 0.0)
            barrier()
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(nsteps):
                    e5.energy_async(f5, 0.0)
                e1.record(stream)
            r5 = e5.energy_wait()
            barrier()
            ms5 = max_over_ranks(e0.elapsed_time(e1)) / nsteps
            return c5["npts"] / (ms5 * 1e-3), ms5, r5["e_tot"], e5.pi_kernel_info()["int8_slice_pairs"]
        finally:
            e5.close()

    # ---- the same step with Pi from the FP64 tensor-pipe kernels (no int8 path): the FP64-pure rate
    fp64_pure = None
    if i8_pairs and not args.no_fp64:
        v64, ms64, e64, _ = quick_rate(args.config, cfg["functional"], "0", max(3, min(steps, 10)))
        fp64_pure = {"value": v64, "ms_per_step": ms64, "pi_arithmetic": "FP64 (DMMA.8x8x4), ORDM_K2_INT8=0",
                     "abs_diff_e_tot_vs_int8_path": abs(e64 - res["e_tot"])}
    # ---- BASELINE.json's largest config at this N (20 M points, 73 core + 26 active orbitals), tSVWN and tPBE
    config5 = None
    if args.config == 4 and not args.npts and not args.no_config5:
        config5 = {"workload": CONFIGS[5]["name"].replace(", tPBE", ""), "npts": CONFIGS[5]["npts"], "n_gpus": world, "scaling": "strong"}
        for fn5 in ("SVWN", "PBE"):
            v5, ms5, e5tot, p5 = quick_rate(5, fn5, None, 5)
            config5["t" + fn5] = {"value": v5, "unit": UNIT, "ms_per_step": ms5, "steps": 5, "warmup": 3, "e_tot": e5tot,
                                  "pi_arithmetic": "int8 tensor cores" if p5 else "FP64 (DMMA.8x8x4)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines
    peaks, peak_src = measured_peaks()
    dg_burst, dg_sus = dgemm_peak(torch)
    pts_rank = cfg["npts"] / world  # per-kernel times are per rank
    k2_tf = alg["k2_flops"] * pts_rank / (ms_k["pi"] * 1e-3) * 1e-12
    k1_gbs = alg["k1_bytes"] * pts_rank / (ms_k["density"] * 1e-3) * 1e-9
    k3_gbs = alg["k3_bytes"] * pts_rank / (ms_k["xc"] * 1e-3) * 1e-9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01c_traffic.json")) as f:
            tr = json.load(f)
        if args.config == 4 and world == 1 and not args.npts:
            traffic = tr["config4_n1"]["k2_ontop"]
    except Exception:
        pass
    if i8_pairs:
        # int8 tensor-core K2 (csrc/k2_int8.cuh): algorithmic work = the exact slice-pair products of the block-upper-
        # triangular contraction, 2 ops per int8 multiply-add; the FP64 flop count of the same contraction is kept beside it
        kp = (alg["M"] + 31) // 32 * 32
        i8_ops = 2 * i8_pairs * 32 * sum(kp - 32 * kc for kc in range(kp // 32))
        sm_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        i8_peak = 2 * 8192 * 148 * sm_mhz * 1e6 * 1e-12
        def i8_top(ms):
            return i8_ops * pts_rank / (ms * 1e-3) * 1e-12
        traffic, traffic_source = None, None
        try:   # dram__bytes of one launch from the committed ncu --set full capture of this kernel on this config (not re-measured here)
            with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
                tr = json.load(f)
            if args.config == 4 and world == 1 and not args.npts:
                traffic, traffic_source = tr["config4_n1"]["k2i_ontop_pair"], "profiles/r02_traffic.json (ncu --set full, same command; not measured in this run)"
        except Exception:
            pass
        roofline = {"kernel": f"k2i_ontop_pair (Pi, tcgen05.mma cta_group::2 kind::i8, {i8_pairs} slice pairs)", "bound": "tensor", "achieved": i8_top(ms_k["pi"]), "peak": i8_peak,
                    "unit": "TOP/s", "frac": i8_top(ms_k["pi"]) / i8_peak, "traffic": traffic, "traffic_source": traffic_source, "pi_arithmetic": pi_arithmetic,
                    "peak_source": f"8192 int8 MAC/clk/SM x 148 SMs x {sm_mhz:.0f} MHz: the issue rate measured by tools/ozlab/i8mma_bench "
                                   "(112.2 clk per 128x224x32 MMA, profiles/r01f_i8mma_microbench.log); MEASURED_PEAKS.json has no int8 figure "
                                   f"(2 x its dense bf16 burst = {2 * peaks.get('bf16_tflops', 0):.0f} TOP/s)",
                    "algorithmic_int8_ops_per_point": i8_ops, "mma_per_256_point_pair_tile": i8_mma,
                    "fp64_flops_per_point_of_the_same_contraction": alg["k2_flops"], "fp64_equivalent_tflops": k2_tf,
                    "fp64_equivalent_vs_cublas_dgemm": k2_tf / dg_sus,
                    "dgemm_peak_source": f"cuBLAS DGEMM 8192^3 measured in this process (sustained {dg_sus:.2f}, burst {dg_burst:.2f} TFLOP/s)",
                    "share_of_step": ms_k["pi"] / ms_step}
    else:
        roofline = {"kernel": "k2_ontop (Pi, DMMA.8x8x4)", "bound": "tensor", "achieved": k2_tf, "peak": dg_sus, "unit": "TFLOP/s", "frac": k2_tf / dg_sus,
                    "traffic": traffic, "traffic_source": "profiles/r01c_traffic.json (ncu --set full; not measured in this run)" if traffic else None, "pi_arithmetic": pi_arithmetic, "peak_source": f"cuBLAS DGEMM 8192^3 measured in this process (sustained {dg_sus:.2f}, burst {dg_burst:.2f} TFLOP/s)",
                    "algorithmic_flops_per_point": alg["k2_flops"], "survey_flops_per_point": alg["k2_flops_survey"],
                    "achieved_survey_count": alg["k2_flops_survey"] * pts_rank / (ms_k["pi"] * 1e-3) * 1e-12,
                    "share_of_step": ms_k["pi"] / ms_step}
    def kernel_table(ms):
        g1 = alg["k1_bytes"] * pts_rank / (ms["density"] * 1e-3) * 1e-9
        g3 = alg["k3_bytes"] * pts_rank / (ms["xc"] * 1e-3) * 1e-9
        t2 = alg["k2_flops"] * pts_rank / (ms["pi"] * 1e-3) * 1e-12
        return {
            "k1_density": {"ms": ms["density"], "bound": "hbm", "achieved": g1, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g1 / peaks["hbm_gbs"], "bytes_per_point": alg["k1_bytes"]},
            "k2_ontop": ({"ms": ms["pi"], "bound": "tensor", "achieved": i8_top(ms["pi"]), "peak": i8_peak, "unit": "TOP/s", "frac": i8_top(ms["pi"]) / i8_peak,
                          "fp64_equivalent_tflops": t2} if i8_pairs else
                         {"ms": ms["pi"], "bound": "tensor", "achieved": t2, "peak": dg_sus, "unit": "TFLOP/s", "frac": t2 / dg_sus}),
            "k3_translate_xc": {"ms": ms["xc"], "bound": "hbm (FP64-ALU limited, DESIGN.md)", "achieved": g3, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g3 / peaks["hbm_gbs"], "bytes_per_point": alg["k3_bytes"],
                                "fp64_alu": {"instr_per_point_estimate": 500 if gga else 330, "achieved_ginstr_s": (500 if gga else 330) * pts_rank / (ms["xc"] * 1e-3) * 1e-9,
                                             "peak_ginstr_s": 16700.0, "note": "DFMA issue peak 33.4 TFLOP/s / 2 (profiles/r01_fp64_peaks_microbench.log); ncu: FP64 pipe 46-48 % active (profiles/r01c_ncu_full_k1_k2_k3.csv)"}},
            "k4_reduce_allreduce": {"ms": ms["reduce"]},
        }
    kernels = kernel_table(ms_k)
    kernels["peak_source"] = f"MEASURED_PEAKS.json ({peak_src})"
    kernels["mode"] = {2: "overlapped/split: K1's core-orbital sums (k1_core_light) stream underneath k2_ontop on a second stream, its active-space part "
                          "follows; k1_density.ms = both parts, overlapping k2_ontop.ms in time",
                       1: "overlapped: k1_density streams underneath k2_ontop on a second stream (their ms overlap in time); k3 joins both",
                       0: "stages back to back on one stream"}[overlapped]
    if i8_pairs and overlapped == 2:
        kernels["mode"] = ("overlapped/split: K1's core-orbital sums (k1_core_light<.,8>, one deep-unrolled CTA per SM) stream beside the int8 k2i_ontop "
                           "on a second stream, its active-space part follows; k1_density.ms = both parts, overlapping k2_ontop.ms in time")
    if ms_iso:
        kernels["isolated"] = kernel_table(ms_iso)
        kernels["isolated"]["ms_per_step"] = ms_step_serial
        kernels["isolated"]["mode"] = "ORDM_OPT_SERIAL_STAGES: K1 -> K2 -> K3 back to back on one stream, same inputs"

    cpu = None
    if not args.no_cpu and world == 1:
        sample = args.cpu_sample or (16384 if cfg["nact"] >= 20 else 131072)
        leg = cpu_reference_leg(cfg, min(sample, cfg["npts"]), cores)
        if leg:
            cpu = {"value": leg["as_written"], "unit": UNIT, "cores": cores, "kind": "reference", "sample": leg["sample"],
                   "value_necessary_work_only": leg["necessary"], "value_serial_1_core": leg["serial"],
                   "value_reference_openmp_build_all_cores": leg["omp"]}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": DTYPE, "data": "synthetic", "config": workload,
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
           "result": {k: res[k] for k in ("e_tot", "ex", "ec", "n_tot", "trn_tot")}, "parity": parity,
           "value_fp64_dmma": fp64_pure, "config5": config5, "allreduce": allreduce}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- particle-steps/s of one Strang step of GEMPIC's 1d2v Hamiltonian splitting.

    python bench.py --gpus N --steps K --warmup W          (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Weibel instability 1d2v, HamiltonianSplitting{1,2},
32 cells, degree 3/2 :galerkin, dt = 0.05, 1e8 particles per B200 (weak scaling: N GPUs hold
N x 1e8 particles, rho/j moments all-reduced over NCCL every depositing sub-step).
A "step" is one strang_splitting!(h, dt, 1): five streaming particle passes
(HE, Hp2, Hp1, Hp2, HE) + the replicated field solves.

value   device-resident path (fields stay on the GPU), CUDA events on the library stream.
e2e     the reference-facing call with HOST field buffers: every step copies e1,e2,b host->device
        and e1,e2,b device->host inside the timed region (gempic_hs_strang_splitting_host),
        exactly what the Julia shim does for the aliased e_dofs/b_dofs arrays.
roofline  dominant kernel = operatorHp1 pass: 48 algorithmic B/particle
        over its mean device time, measured with CUDA events inside the timed region.
cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/, "port";
        Julia is not installed, so the reference itself cannot run) with OpenMP chunking like
        the reference's Threads.@spawn chunks, on a bounded particle sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload (test/test_vm_1d2v.jl:11-31) ----------------------------
K_WEIBEL = 1.25
L_WEIBEL = 5.02654824574          # xmax of test_vm_1d2v.jl:21 (2 pi / k)
SIGMA = (0.2, 0.005773502691896)
BETA = 1e-4
NX = 32
DT = 0.05
DEG = 3
# workloads: BASELINE.json configs[1] (default, the metric's config), configs[2] and configs[3] at one GPU's share
WORKLOADS = {
    "weibel": dict(name="Weibel instability 1d2v HamiltonianSplitting, 32 cells, deg 3/2 galerkin, dt 0.05 (BASELINE configs[1])",
                   L=L_WEIBEL, sigma=SIGMA, kind="uniform", alpha=0.0, k=K_WEIBEL, integrator="hs"),
    "landau": dict(name="strong Landau damping 1d2v HamiltonianSplitting, 32 cells, deg 3/2 galerkin, dt 0.05 (BASELINE configs[2])",
                   L=4 * math.pi, sigma=(1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="hs"),
    "boris": dict(name="Boris-splitting 1d2v push (HamiltonianSplittingBoris), strong Landau load, 32 cells, deg 3/2, dt 0.05 "
                       "(BASELINE configs[3])",
                  L=4 * math.pi, sigma=(1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="boris"),
}
WORKLOADS["2d3v"] = dict(name="2d3v HamiltonianSplitting with TwoDMaxwell, 64x64 cells, deg 3 splines, Landau load along x1, dt 0.05 "
                              "(BASELINE configs[4] at one GPU's share: 1.25e8 particles per GPU)",
                         L=4 * math.pi, sigma=(1.0, 1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="hs2d")
NX2 = 64
# 2d3v rows x1,x2,v1,v2,v3,w: HE 40 R + 24 W, Hp3 48 R + 16 W, Hp1/Hp2 48 R + 24 W; sort 2 x 48 + keys
BYTES2 = {"fused[HE,Hp3]{2,3}": 72, "fused[HE,HE,Hp3]{2,3}": 72, "operatorHE{2,3}": 64, "operatorHp3{2,3}": 64, "operatorHp1{2,3}": 72, "operatorHp2{2,3}": 72, "cell sort 2d": 112,
          "operatorHp2{2,3}+sort": 96, "cell histogram after Hp2": 24,
          "strang_step": 2 * 64 + 2 * 64 + 3 * 72}
# algorithmic DRAM bytes per particle of each pass (fp64 SoA rows x, v1, v2, w; SURVEY section 8d / DESIGN.md)
BYTES = {"operatorHE": 40, "operatorHp2": 40, "operatorHp1": 48, "strang_step": 208,
         "fused[HE,Hp2,Hp1,Hp2]": 56, "fused[HE,HE,Hp2,Hp1,Hp2]": 56, "boris_step": 56, "boris_strang_step": 56,
         "operatorHE+j2": 48, "j2 deposit": 24}


def ncu_traffic(tag):
    """dram bytes per particle of a pass from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(tag)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def weibel_b0(mod, mx, wl=None):
    """B3(0) = beta cos(kx) through l2projection! (test_vm_1d2v.jl:33,80-84); B = 0 for the Landau loads
    (examples/strong_landau_damping_1d2v.jl:36-39)"""
    b = np.zeros(NX)
    if wl is None or wl["kind"] == "uniform":
        mx.l2projection(b, lambda x: BETA * math.cos(2 * math.pi * x / L_WEIBEL), DEG - 1)
    return b


def host_state(wl, n, rng):
    """seeded host particle set of a workload (the CPU arm; the GPU arm samples on the device)"""
    L = wl["L"]
    st = np.empty((4, n))
    if wl["kind"] == "uniform":
        st[0] = rng.uniform(0, L, n)
    else:   # inverse CDF of 1 + alpha cos(kx) by Newton (particle_sampling.jl:294-311)
        u = rng.uniform(0, 1, n)
        x = u * L
        for _ in range(30):
            x -= (x + wl["alpha"] / wl["k"] * np.sin(wl["k"] * x) - u * L) / (1 + wl["alpha"] * np.cos(wl["k"] * x))
        st[0] = np.mod(x, L)
    st[1] = wl["sigma"][0] * rng.normal(size=n)
    st[2] = wl["sigma"][1] * rng.normal(size=n)
    st[3] = L
    return st


# =============================== CPU reference arm ==============================================
def cpu_reference(n_cpu: int, steps: int, warmup: int, threads: int | None = None, workload: str = "weibel"):
    """Times the oracle's strang_splitting (C restatement of src/hamiltonian_splitting_1d2v.jl / _boris.jl,
    OpenMP over particle chunks with private deposit buffers like the reference's @spawn chunks)."""
    from oracle import oracle as orc

    orc.build()
    wl = WORKLOADS[workload]
    cores = threads or orc.max_threads()
    n_cpu -= n_cpu % cores
    rng = np.random.default_rng(1234)
    mesh = orc.OneDGrid(0.0, wl["L"], NX)
    pg = orc.ParticleGroup(1, 2, n_cpu)
    pg.array[:, :] = host_state(wl, n_cpu, rng)
    ks0 = orc.ParticleMeshCoupling1D(mesh, n_cpu, DEG, "galerkin")
    ks1 = orc.ParticleMeshCoupling1D(mesh, n_cpu, DEG - 1, "galerkin")
    mx = orc.Maxwell1DFEM(mesh, DEG)
    e1, e2, rho = np.zeros(NX), np.zeros(NX), np.zeros(NX)
    b = weibel_b0(orc, mx, wl)
    orc.solve_poisson(e1, pg, ks0, mx, rho)
    if wl["integrator"] == "boris":
        h = orc.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1, e2], b)
        h.staggering(DT)
        cores = 1   # the oracle's Boris loops are serial, like the reference's (hamiltonian_splitting_boris.jl:189-288)
    else:
        h = orc.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, n_chunks=cores)
    for _ in range(warmup):
        h.strang_splitting(DT, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        h.strang_splitting(DT, 1)
    dt = time.perf_counter() - t0
    return {"value": n_cpu * steps / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_cpu} particles x {steps} Strang steps (same {workload} 1d2v config), C restatement of the "
                      f"reference Julia path with OpenMP chunks -- Julia unavailable", "seconds": dt}, dt / steps


def cpu_reference_2d(n_cpu: int, steps: int, warmup: int):
    """the serial C restatement of the 2d3v operators (oracle/splitting2d3v.py) on a bounded sample"""
    from oracle import maxwell2d as m2
    from oracle import oracle as orc
    from oracle import splitting2d3v as s2

    wl = WORKLOADS["2d3v"]
    L = wl["L"]
    rng = np.random.default_rng(1234)
    st1 = host_state(dict(wl, sigma=(1.0, 1.0)), n_cpu, rng)
    pg = s2.ParticleGroup23(n_cpu)
    pg.array[0], pg.array[1] = st1[0], rng.uniform(0, L, n_cpu)
    pg.array[2], pg.array[3], pg.array[4] = st1[1], st1[2], rng.normal(size=n_cpu)
    pg.array[5] = L * L
    mx = m2.TwoDMaxwell(orc.TwoDGrid(0.0, L, NX2, 0.0, L, NX2), DEG)
    nd = NX2 * NX2
    e, b = [np.zeros(nd) for _ in range(3)], [np.zeros(nd) for _ in range(3)]
    h = s2.HamiltonianSplitting2D3V(mx, pg, e, b)
    rho = h.charge_density()
    mx.compute_e_from_rho(e, rho - rho.mean())
    for _ in range(warmup):
        h.strang_splitting(DT, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        h.strang_splitting(DT, 1)
    dt = time.perf_counter() - t0
    return {"value": n_cpu * steps / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
            "sample": f"{n_cpu} particles x {steps} Strang steps (same 2d3v config), serial C restatement of the 2D "
                      f"extension of the reference operators -- the reference has no 2d3v integrator and Julia is unavailable",
            "seconds": dt}, dt / steps


def run_ours_2d(args):
    import torch

    import __graft_entry__ as ge

    gp = ge.load_package()
    dc = gp.DistributedContext()
    if dc.world_size != args.gpus and dc.world_size > 1:
        args.gpus = dc.world_size
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: libgempic_b200 has no CPU path")
    dc.init_library_comm()
    Lib = gp.load()
    wl = WORKLOADS["2d3v"]
    L = wl["L"]
    n_local = args.particles if args.particles != 100_000_000 else 125_000_000
    n_global = n_local * dc.world_size
    pg = gp.ParticleGroup(2, 3, n_local, common_weight=1.0 / n_global)
    pg.sample("landau", 0.0, L, alpha=wl["alpha"], k=wl["k"], sigma=wl["sigma"], seed=1234, first_index=dc.rank * n_local)
    mx = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L, NX2, 0.0, L, NX2), DEG)
    nd = NX2 * NX2
    e, b = [np.zeros(nd) for _ in range(3)], [np.zeros(nd) for _ in range(3)]
    h = gp.HamiltonianSplitting2D3V(mx, pg, e, b, resident=True)
    h.set_sort_interval(args.sort_interval)
    h.set_fusion(bool(args.fuse))
    rho = h.charge_density()
    total_charge = float(rho.sum())
    mx.compute_e_from_rho(e, rho - rho.mean())
    b[2][:] = BETA * np.cos(2 * np.pi * (np.arange(nd) % NX2) / NX2)
    h.upload_fields()
    r0 = h.gauss_residual()
    stream = torch.cuda.ExternalStream(gp.stream_ptr(), device=torch.device("cuda", dc.local_rank))

    def timed(fn):
        dc.barrier(); gp.synchronize(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        fn()
        ev1.record(stream)
        gp.synchronize(); torch.cuda.synchronize(); dc.barrier()
        return dc.max_over_ranks(ev0.elapsed_time(ev1))

    h.strang_splitting(DT, args.warmup)
    gp.synchronize()
    sampler = ClockSampler(dc.local_rank)
    if dc.rank == 0:
        sampler.start()
    _lib = sys.modules["gempic_jl_b200._lib"]
    _lib.check(Lib.gempic_profile_enable(C.c_int(1)))
    gp.launch_count(reset=True)
    ms = timed(lambda: h.strang_splitting(DT, args.steps))
    launches = gp.launch_count()
    prof, slot = {}, 0
    while True:
        tag = C.create_string_buffer(64)
        t, cnt = C.c_double(), C.c_int64()
        if Lib.gempic_profile_read(C.c_int(slot), tag, C.c_int(64), C.byref(t), C.byref(cnt)) != 0:
            break
        prof[tag.value.decode()] = (t.value, cnt.value)
        slot += 1
    _lib.check(Lib.gempic_profile_enable(C.c_int(0)))
    clocks = sampler.stop() if dc.rank == 0 else None
    value = n_global * args.steps / (ms * 1e-3)

    # e2e: host field buffers every step (6 x nx*ny doubles in, 6 out), synchronous
    h.sync_fields()
    h_host = gp.HamiltonianSplitting2D3V(mx, pg, e, b, resident=False)
    h_host.set_sort_interval(args.sort_interval)
    h_host.set_fusion(bool(args.fuse))
    for _ in range(2):
        h_host.strang_splitting(DT, 1)
    dc.barrier(); gp.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h_host.strang_splitting(DT, 1)
    gp.synchronize()
    e2e_s = dc.max_over_ranks(time.perf_counter() - t0)

    # sanity: the timed state is a real simulation (charge exact, Gauss law conserved, finite fields)
    h.upload_fields()
    r1 = h.gauss_residual()
    assert abs(total_charge - L * L) < 1e-9 * L * L, "total charge is wrong"
    assert np.max(np.abs(r1 - r0)) < 1e-10 * np.max(np.abs(rho)), "discrete Gauss law violated -- the step did not do its work"
    assert all(np.all(np.isfinite(v)) for v in e + b)

    if dc.rank == 0:
        peak, peak_src = peaks()
        passes = {k: v for k, v in prof.items() if k in BYTES2 and k not in ("cell sort 2d", "cell histogram after Hp2")}
        dom = max(passes, key=lambda k: passes[k][0]) if passes else None
        roof = None
        if dom:
            t_ms = prof[dom][0] / prof[dom][1]
            achieved = BYTES2[dom] * n_local / (t_ms * 1e-3) / 1e9
            tr = ncu_traffic(dom)
            roof = {"bound": "hbm", "kernel": f"k2_pass<{dom}>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": (tr["dram_bytes_per_particle"] * n_local if tr else None),
                    "traffic_note": (f"{tr['dram_bytes_per_particle']} B/particle, {tr['source']}" if tr else None),
                    "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES2[dom], "avg_launch_ms": t_ms,
                    "all_passes": {k: {"avg_ms": v[0] / max(v[1], 1), "launches": v[1],
                                       "GBps": BYTES2.get(k, 0) * n_local / (v[0] / max(v[1], 1) * 1e-3) / 1e9 if v[1] else None}
                                   for k, v in prof.items()},
                    "step_GBps_vs_unfused_bytes": BYTES2["strang_step"] * n_local / (ms / args.steps * 1e-3) / 1e9}
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            cpu, _ = cpu_reference_2d(min(args.cpu_particles, 400_000), 2, 1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cfg = {"workload": wl["name"], "particles_per_gpu": int(n_local), "n_cells": [NX2, NX2], "spline_degree": [DEG, DEG - 1],
               "dt": DT, "parallelism": f"particles sharded over {dc.world_size} GPU(s), NCCL allreduce of j1/j2/j3",
               "l2_policy": "inputs (48 B/particle x N >> 126 MB L2) stream from HBM every pass; no flush needed",
               "kernels": ("fused [HE,(HE,)Hp3] + Hp2, Hp1, Hp2, Hp3 passes, one strang_splitting!(h, dt, K) call" if args.fuse else
                           "one pass per operator (HE, Hp3, Hp2, Hp1, Hp2, Hp3, HE)") + f", cell sort every {args.sort_interval} step(s)"}
        emit({"metric": "particle-steps/s per Strang step", "value": value, "unit": "particle-steps/s", "n_gpus": dc.world_size,
              "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
              "e2e": {"value": n_global * args.steps / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": 6 * nd * 8,
                      "d2h_bytes_per_step": 6 * nd * 8, "steps": args.steps,
                      "what": "gempic_hs2d_strang_splitting_host: host e[3], b[3] in and out, synchronous"},
              "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu})
    dc.finalize()
    return 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_cpu = args.cpu_particles
    if args.workload == "2d3v":
        n_cpu = min(n_cpu, 400_000)
        base, sec_per_step = cpu_reference_2d(n_cpu, args.steps, args.warmup)
    else:
        base, sec_per_step = cpu_reference(n_cpu, args.steps, args.warmup, workload=args.workload)
    line = {
        "impl": "reference", "metric": "particle-steps/s per Strang step", "value": base["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_cpu, note="CPU arm: bounded sample of the same workload on the host cores"),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, n_per_gpu, note=None):
    cfg = {"workload": WORKLOADS[args.workload]["name"],
           "particles_per_gpu": int(n_per_gpu), "n_cells": NX, "spline_degree": [DEG, DEG - 1], "dt": DT,
           "parallelism": f"particles sharded over {args.gpus} GPU(s), NCCL allreduce of rho/j",
           "l2_policy": "inputs (32 B/particle x N >> 126 MB L2) stream from HBM every pass; no flush needed",
           "kernels": ("one fused pass per step [push_v_epart, push_v_bpart, push_v_epart, push_x_accumulate_j]"
                       if WORKLOADS[args.workload]["integrator"] == "boris" else
                       "fused passes [HE,(HE,)Hp2,Hp1,Hp2] + HE, one strang_splitting!(h, dt, K) call" if args.fuse
                       else "one pass per reference operator, K calls of strang_splitting!(h, dt, 1)")}
    if note:
        cfg["note"] = note
    return cfg


# =============================== GPU arm ==========================================================
def run_ours(args):
    import torch

    import __graft_entry__ as ge

    gp = ge.load_package()
    dc = gp.DistributedContext()
    if dc.world_size != args.gpus and dc.world_size > 1:
        args.gpus = dc.world_size
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device: libgempic_b200 has no CPU path")
    dc.init_library_comm()
    L = gp.load()
    n_local = args.particles
    n_global = n_local * dc.world_size
    first = dc.rank * n_local

    wl = WORKLOADS[args.workload]
    boris = wl["integrator"] == "boris"
    Lw = wl["L"]
    mesh = gp.OneDGrid(0.0, Lw, NX)
    pg = gp.ParticleGroup(1, 2, n_local, common_weight=1.0 / n_global)
    pg.sample(wl["kind"], 0.0, Lw, alpha=wl["alpha"], k=wl["k"], sigma=wl["sigma"], seed=1234, first_index=first)
    ks0 = gp.ParticleMeshCoupling1D(mesh, n_local, DEG, "galerkin")
    ks1 = gp.ParticleMeshCoupling1D(mesh, n_local, DEG - 1, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, DEG)
    e1, e2, rho = np.zeros(NX), np.zeros(NX), np.zeros(NX)
    b = weibel_b0(gp, mx, wl)
    gp.solve_poisson(e1, pg, ks0, mx, rho)
    total_charge = float(rho.sum())

    if boris:
        h = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1, e2], b, resident=True)
        h.staggering(DT)
        args.fuse = 1   # strang_splitting!(h::HamiltonianSplittingBoris, dt, K) is one call of K fused passes
    else:
        h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, resident=True)
        if args.fuse:
            h.set_fusion(True)

    stream = torch.cuda.ExternalStream(gp.stream_ptr(), device=torch.device("cuda", dc.local_rank))

    def timed(fn, steps):
        dc.barrier()
        gp.synchronize()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        gp.synchronize()
        torch.cuda.synchronize()
        dc.barrier()
        return dc.max_over_ranks(ev0.elapsed_time(ev1))   # ms, max over ranks

    # ---- value: device-resident -------------------------------------------------------------
    # fused: one strang_splitting!(h, dt, K) call (the trailing HE of a step rides in the next step's pass);
    # unfused: K calls of one step, one kernel per reference operator
    step = lambda: h.strang_splitting(DT, 1)
    for _ in range(args.warmup):
        step()
    gp.synchronize()
    sampler = ClockSampler(dc.local_rank)
    if dc.rank == 0:
        sampler.start()
    _lib = sys.modules["gempic_jl_b200._lib"]
    _lib.check(L.gempic_profile_enable(C.c_int(1)))
    gp.launch_count(reset=True)
    if args.fuse:
        ms = timed(lambda: h.strang_splitting(DT, args.steps), 1)
    else:
        ms = timed(step, args.steps)
    launches = gp.launch_count()
    prof = {}
    slot = 0
    while True:
        tag = C.create_string_buffer(64)
        t, cnt = C.c_double(), C.c_int64()
        rc = L.gempic_profile_read(C.c_int(slot), tag, C.c_int(64), C.byref(t), C.byref(cnt))
        if rc != 0:
            break
        prof[tag.value.decode()] = (t.value, cnt.value)
        slot += 1
    _lib.check(L.gempic_profile_enable(C.c_int(0)))
    clocks = sampler.stop() if dc.rank == 0 else None
    ms_per_step = ms / args.steps
    value = n_global * args.steps / (ms * 1e-3)

    # ---- e2e: host field buffers every step ---------------------------------------------------
    h.sync_fields()
    if boris:
        # same object family, host-buffer entry point (gempic_boris_strang_splitting_host); the staggered mid fields
        # live in the splitting object, so the host variant is re-staggered from the current state
        h_host = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1, e2], b, resident=False)
        h_host.staggering(DT)
    else:
        h_host = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, resident=False)
        if args.fuse:
            h_host.set_fusion(True)
    for _ in range(min(args.warmup, 3)):
        h_host.strang_splitting(DT, 1)
    e2e_steps = args.steps
    dc.barrier()
    gp.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_host.strang_splitting(DT, 1)      # synchronous: H2D fields, 5 passes + solves, D2H fields
    gp.synchronize()
    e2e_s = dc.max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_global * e2e_steps / e2e_s

    # ---- sanity: the timed state is a real simulation -------------------------------------------
    rho2, ep = np.zeros(NX), np.zeros(NX)
    gp.solve_poisson(ep, pg, ks0, mx, rho2)
    assert abs(rho2.sum() - Lw) < 1e-9 * Lw and abs(total_charge - Lw) < 1e-9 * Lw, \
        "charge is not conserved -- the step did not do its work"
    assert np.all(np.isfinite(e1)) and np.all(np.isfinite(b))

    if dc.rank == 0:
        peak, peak_src = peaks()
        dom = max(prof, key=lambda k: prof[k][0]) if prof else None   # the pass with the largest share of the step
        roof = None
        if dom in prof and prof[dom][1] > 0:
            t_ms = prof[dom][0] / prof[dom][1]
            achieved = BYTES[dom] * n_local / (t_ms * 1e-3) / 1e9
            tr = ncu_traffic(dom)
            roof = {"bound": "hbm", "kernel": f"k_pass<{dom}>", "achieved": achieved,
                    "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": (tr["dram_bytes_per_particle"] * n_local if tr else None),
                    "traffic_note": (f"bytes per launch = {tr['dram_bytes_per_particle']} B/particle (ncu dram__bytes_read+write, "
                                     f"{tr['source']}) x {n_local} particles" if tr else None),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_particle": BYTES[dom], "avg_launch_ms": t_ms,
                    "all_passes": {k: {"avg_ms": v[0] / max(v[1], 1), "launches": v[1],
                                       "GBps": BYTES.get(k, 0) * n_local / (v[0] / max(v[1], 1) * 1e-3) / 1e9 if v[1] else None}
                                   for k, v in prof.items()},
                    "step_GBps_vs_unfused_bytes": BYTES["boris_strang_step" if boris else "strang_step"] * n_local / (ms_per_step * 1e-3) / 1e9}
        n_cpu = args.cpu_particles
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            cpu, _ = cpu_reference(n_cpu, args.cpu_steps, 1, workload=args.workload)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": "particle-steps/s per Strang step", "value": value, "unit": "particle-steps/s", "n_gpus": dc.world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_local),
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": 3 * NX * 8,
                    "d2h_bytes_per_step": 3 * NX * 8, "steps": e2e_steps,
                    "what": ("gempic_boris_strang_splitting_host" if boris else "gempic_hs_strang_splitting_host") +
                            ": host e1,e2,b in, e1,e2,b out, synchronous"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
        }
        emit(line)
    dc.finalize()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    # Libraries below us (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on stdout,
    # so everything else is routed to stderr and the JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=100_000_000, help="particles per GPU")
    ap.add_argument("--cpu-particles", type=int, default=4_000_000, help="bounded sample for the CPU arm")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="weibel", choices=sorted(WORKLOADS), help="weibel = BASELINE configs[1] (the metric's config)")
    ap.add_argument("--sort-interval", type=int, default=1, help="2d3v: cell sort every k Strang steps")
    ap.add_argument("--fuse", type=int, default=1, help="1: fused particle passes (default), 0: one kernel per reference operator")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours_2d(args) if args.workload == "2d3v" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

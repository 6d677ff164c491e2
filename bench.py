#!/usr/bin/env python
"""bench.py -- particle-steps/s of one Strang step of GEMPIC's Hamiltonian splitting on B200.

    python bench.py --gpus N --steps K --warmup W          (ours; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload of the headline line (BASELINE.json configs[1], the metric's config): Weibel instability 1d2v,
HamiltonianSplitting{1,2}, 32 cells, degree 3/2 :galerkin, dt = 0.05, 1e8 particles per B200 (weak scaling: N GPUs
hold N x 1e8 particles, the rho/j moments are all-reduced over NCCL every depositing sub-step).
A "step" is one strang_splitting!(h, dt, 1) = HB HE Hp2 Hp1 Hp2 HE HB (hamiltonian_splitting.jl:98-108); with the
default fused passes that is ONE streaming particle pass [HE,HE,Hp2,Hp1,Hp2] of 56 B/particle + one field kernel.

value     device-resident path (fields stay on the GPU): a region of K steps timed with CUDA events on the library
          stream, barrier + synchronize on both sides, max over ranks.  The region is repeated until >= 1 s of device
          time has been measured: `value` / `ms_per_step` are the MEDIAN region, `sustained` is all regions together
          (what a long run gets once the board sits at its power cap), `clocks` are sampled over all of them.
e2e       the reference-facing call with HOST field buffers: every step copies e1,e2,b host->device and back inside
          the timed region (gempic_hs_strang_splitting_host), exactly what the Julia shim does for the aliased
          e_dofs / b_dofs arrays; regions of K calls repeated for >= 1 s like `value`, median reported
          (`first_region_value`: the burst figure before the power cap bites).  `e2e_loop` is the example's whole loop body
          (examples/strong_landau_damping_1d2v.jl:46-59): strang_splitting!; solve_poisson!; write_step! per step.
roofline  the pass with the largest share of the step (the fused pass: 56 algorithmic B/particle) over
          its mean device time, CUDA events around every launch inside the timed regions.
configs   short regions of the other BASELINE configs at this GPU count -- strong Landau 1d2v (configs[2]), Boris
          (configs[3]), 2d3v (configs[4]) at 1.25e8 particles per GPU -- and one strong-scaling point (1e8 particles
          in total over the N GPUs).
sharded_parity  (N > 1) sharded vs un-sharded runs of the three integrators on a 4e5 / 1.2e5-particle sub-problem
          before the timed region (gempic.jl_b200/selfcheck.py); the run fails above 1e-11.
cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/, "port"; Julia is not installed,
          so the reference itself cannot run) with OpenMP chunking like the reference's Threads.@spawn chunks, on a
          bounded particle sample, the thread count pinned to the host's cores (and a 1-thread figure beside it).
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
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

# ---- workloads (test/test_vm_1d2v.jl:11-31, examples/strong_landau_damping_1d2v.jl:7-39) ------
K_WEIBEL = 1.25
L_WEIBEL = 5.02654824574          # xmax of test_vm_1d2v.jl:21 (2 pi / k)
SIGMA = (0.2, 0.005773502691896)
BETA = 1e-4
NX = 32
NX2 = 64
DT = 0.05
DEG = 3
WORKLOADS = {
    "weibel": dict(name="Weibel instability 1d2v HamiltonianSplitting, 32 cells, deg 3/2 galerkin, dt 0.05 (BASELINE configs[1])",
                   L=L_WEIBEL, sigma=SIGMA, kind="uniform", alpha=0.0, k=K_WEIBEL, integrator="hs", particles=100_000_000),
    "landau": dict(name="strong Landau damping 1d2v HamiltonianSplitting, 32 cells, deg 3/2 galerkin, dt 0.05 (BASELINE configs[2])",
                   L=4 * math.pi, sigma=(1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="hs", particles=125_000_000),
    "boris": dict(name="Boris-splitting 1d2v push (HamiltonianSplittingBoris), strong Landau load, 32 cells, deg 3/2, dt 0.05 "
                       "(BASELINE configs[3])",
                  L=4 * math.pi, sigma=(1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="boris", particles=125_000_000),
    "2d3v": dict(name="2d3v HamiltonianSplitting with TwoDMaxwell, 64x64 cells, deg 3 splines, Landau load along x1, dt 0.05 "
                      "(BASELINE configs[4] at one GPU's share: 1.25e8 particles per GPU)",
                 L=4 * math.pi, sigma=(1.0, 1.0, 1.0), kind="landau", alpha=0.5, k=0.5, integrator="hs2d", particles=125_000_000),
}
# algorithmic DRAM bytes per particle of each pass (fp64 SoA rows; SURVEY section 8d / DESIGN.md section 4)
BYTES = {"operatorHE": 40, "operatorHp2": 40, "operatorHp1": 48, "operatorHp1+rho": 48,
         "fused[HE,Hp2,Hp1,Hp2]": 56, "fused[HE,HE,Hp2,Hp1,Hp2]": 56, "boris_step": 56,
         "operatorHE+j2": 48, "j2 deposit": 24, "add_charge": 16, "write_step sums": 32, "loop tail [HE,rho,diag]": 48,
         # 2d3v rows x1,x2,v1,v2,v3,w: HE 40 R + 24 W, Hp3 48 R + 16 W, Hp1/Hp2 48 R + 24 W; sort 2 x 48 + keys
         "fused[HE,Hp3]{2,3}": 72, "fused[HE,HE,Hp3]{2,3}": 72, "operatorHE{2,3}": 64, "operatorHp3{2,3}": 64,
         "operatorHp1{2,3}": 72, "operatorHp1{2,3}+hist": 72, "operatorHp2{2,3}": 72, "cell sort 2d": 112, "operatorHp2{2,3}+sort": 96,
         "cell histogram after Hp2": 24, "fused[HE,Hp3,Hp2]{2,3}": 80, "fused[HE,HE,Hp3,Hp2]{2,3}": 80, "operatorHp3{2,3} sorted": 64,
         "fused[Hp1,Hp2,Hp3]{2,3}+sort": 96, "fused[Hp1,Hp2,Hp3]{2,3}": 72, "add_charge2d": 24}
STEP_BYTES = {"hs": 208, "boris": 160, "hs2d": 2 * 64 + 2 * 64 + 3 * 72}   # one pass per reference operator


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
            time.sleep(0.12)   # first sample before the region starts
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
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
        # "under load": the board draws well above idle
        loaded = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(loaded), "power_w_max": max(power)}


def weibel_b0(mx, wl):
    """B3(0) = beta cos(kx) through l2projection! (test_vm_1d2v.jl:33,80-84); B = 0 for the Landau loads
    (examples/strong_landau_damping_1d2v.jl:36-39)"""
    b = np.zeros(NX)
    if wl["kind"] == "uniform":
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
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference(n_cpu: int, steps: int, warmup: int, threads: int | None = None, workload: str = "weibel"):
    """Times the oracle's strang_splitting (C restatement of src/hamiltonian_splitting_1d2v.jl / _boris.jl,
    OpenMP over particle chunks with private deposit buffers like the reference's @spawn chunks).  The team size is
    set explicitly (an OMP_NUM_THREADS=1 inherited from torchrun is ignored)."""
    from oracle import oracle as orc

    orc.build()
    wl = WORKLOADS[workload]
    cores = orc.set_threads(threads or host_cores())
    n_cpu -= n_cpu % cores
    rng = np.random.default_rng(1234)
    mesh = orc.OneDGrid(0.0, wl["L"], NX)
    pg = orc.ParticleGroup(1, 2, n_cpu)
    pg.array[:, :] = host_state(wl, n_cpu, rng)
    ks0 = orc.ParticleMeshCoupling1D(mesh, n_cpu, DEG, "galerkin")
    ks1 = orc.ParticleMeshCoupling1D(mesh, n_cpu, DEG - 1, "galerkin")
    mx = orc.Maxwell1DFEM(mesh, DEG)
    e1, e2, rho = np.zeros(NX), np.zeros(NX), np.zeros(NX)
    b = weibel_b0(mx, wl)
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
                      f"reference Julia path with OpenMP chunks on {cores} pinned thread(s) -- Julia unavailable", "seconds": dt}, dt / steps


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


def cpu_baseline_block(args, workload):
    """cpu_baseline of the line: all host cores (pinned) and, beside it, one thread"""
    if workload == "2d3v":
        cpu, _ = cpu_reference_2d(min(args.cpu_particles, 400_000), 2, 1)
        return {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    cpu, _ = cpu_reference(args.cpu_particles, args.cpu_steps, 1, workload=workload)
    out = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if cpu["cores"] > 1:
        one, _ = cpu_reference(max(args.cpu_particles // 8, 100_000), 2, 1, threads=1, workload=workload)
        out["value_1_thread"] = one["value"]
    return out


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
        if base["cores"] > 1:
            one, _ = cpu_reference(max(n_cpu // 8, 100_000), 2, 1, threads=1, workload=args.workload)
            base["value_1_thread"] = one["value"]
    line = {
        "impl": "reference", "metric": "particle-steps/s per Strang step", "value": base["value"], "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, n_cpu, bool(args.fuse),
                                  note="CPU arm: bounded sample of the same workload on the host cores"),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "value_1_thread") if k in base},
        "e2e": {"value": base["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(workload, gpus, n_per_gpu, fuse, note=None, sort_interval=1):
    wl = WORKLOADS[workload]
    two_d = wl["integrator"] == "hs2d"
    cfg = {"workload": wl["name"], "particles_per_gpu": int(n_per_gpu),
           "n_cells": [NX2, NX2] if two_d else NX, "spline_degree": [DEG, DEG - 1], "dt": DT,
           "parallelism": f"particles sharded over {gpus} GPU(s), NCCL allreduce of " + ("j1/j2/j3" if two_d else "rho/j"),
           "l2_policy": f"inputs ({48 if two_d else 32} B/particle x N >> 126 MB L2) stream from HBM every pass; no flush needed"}
    if wl["integrator"] == "boris":
        cfg["kernels"] = "one fused pass per step [push_v_epart, push_v_bpart, push_v_epart, push_x_accumulate_j]"
    elif two_d:
        cfg["kernels"] = (("fused passes [HE,(HE,)Hp3,Hp2] + [Hp1,Hp2,Hp3] with the cell sort riding in the second one" if fuse else
                           "one pass per operator (HE, Hp3, Hp2, Hp1, Hp2, Hp3, HE)") + f", cell sort every {sort_interval} step(s)")
    else:
        cfg["kernels"] = ("fused pass [HE,(HE,)Hp2,Hp1,Hp2] + one field kernel per step, one strang_splitting!(h, dt, K) call"
                          if fuse else "one pass per reference operator, K calls of strang_splitting!(h, dt, 1)")
    if note:
        cfg["note"] = note
    return cfg


# =============================== GPU arm ==========================================================
class Bench:
    """one process per GPU: the library, its communicator and the timing helpers"""

    def __init__(self, args):
        import torch

        import __graft_entry__ as ge

        self.torch = torch
        self.gp = gp = ge.load_package()
        self.dc = dc = gp.DistributedContext()
        if dc.world_size != args.gpus and dc.world_size > 1:
            args.gpus = dc.world_size
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl ours) needs a CUDA device: libgempic_b200 has no CPU path")
        dc.init_library_comm()
        self.L = gp.load()
        self._lib = sys.modules["gempic_jl_b200._lib"]
        self.stream = torch.cuda.ExternalStream(gp.stream_ptr(), device=torch.device("cuda", dc.local_rank))
        self.args = args

    def timed(self, fn):
        """device ms of fn() on the library stream, barrier + synchronize on both sides, max over ranks"""
        dc, gp, torch = self.dc, self.gp, self.torch
        dc.barrier()
        gp.synchronize()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        fn()
        ev1.record(self.stream)
        gp.synchronize()
        torch.cuda.synchronize()
        dc.barrier()
        return dc.max_over_ranks(ev0.elapsed_time(ev1))

    def profile(self, on):
        self._lib.check(self.L.gempic_profile_enable(C.c_int(1 if on else 0)))

    def profile_read(self):
        prof, slot = {}, 0
        while True:
            tag = C.create_string_buffer(64)
            t, cnt = C.c_double(), C.c_int64()
            if self.L.gempic_profile_read(C.c_int(slot), tag, C.c_int(64), C.byref(t), C.byref(cnt)) != 0:
                break
            prof[tag.value.decode()] = (t.value, cnt.value)
            slot += 1
        return prof

    def regions(self, step_k, steps, min_seconds, max_regions=64):
        """repeat the K-step region until >= min_seconds of device time; returns the list of region ms"""
        ms = [self.timed(step_k)]
        while sum(ms) < min_seconds * 1e3 and len(ms) < max_regions:
            ms.append(self.timed(step_k))
        return ms


def build_1d(B, workload, n_local, resident=True, fuse=True):
    gp, dc = B.gp, B.dc
    wl = WORKLOADS[workload]
    Lw = wl["L"]
    n_global = n_local * dc.world_size
    mesh = gp.OneDGrid(0.0, Lw, NX)
    pg = gp.ParticleGroup(1, 2, n_local, common_weight=1.0 / n_global)
    pg.sample(wl["kind"], 0.0, Lw, alpha=wl["alpha"], k=wl["k"], sigma=wl["sigma"], seed=1234, first_index=dc.rank * n_local)
    ks0 = gp.ParticleMeshCoupling1D(mesh, n_local, DEG, "galerkin")
    ks1 = gp.ParticleMeshCoupling1D(mesh, n_local, DEG - 1, "galerkin")
    mx = gp.Maxwell1DFEM(mesh, DEG)
    e1, e2, rho = np.zeros(NX), np.zeros(NX), np.zeros(NX)
    b = weibel_b0(mx, wl)
    gp.solve_poisson(e1, pg, ks0, mx, rho)
    S = dict(wl=wl, pg=pg, ks0=ks0, ks1=ks1, mx=mx, e1=e1, e2=e2, b=b, rho=rho, total_charge=float(rho.sum()), L=Lw,
             n_local=n_local, n_global=n_global)
    if wl["integrator"] == "boris":
        h = gp.HamiltonianSplittingBoris(mx, ks0, ks1, pg, [e1, e2], b, resident=resident)
        h.staggering(DT)
    else:
        h = gp.HamiltonianSplitting(1, 2, mx, ks0, ks1, pg, [e1, e2], b, resident=resident, fuse=fuse)
    S["h"] = h
    return S


def build_2d(B, n_local, resident=True, fuse=True, sort_interval=1):
    gp, dc = B.gp, B.dc
    wl = WORKLOADS["2d3v"]
    L = wl["L"]
    n_global = n_local * dc.world_size
    pg = gp.ParticleGroup(2, 3, n_local, common_weight=1.0 / n_global)
    pg.sample("landau", 0.0, L, alpha=wl["alpha"], k=wl["k"], sigma=wl["sigma"], seed=1234, first_index=dc.rank * n_local)
    mx = gp.TwoDMaxwell(gp.TwoDGrid(0.0, L, NX2, 0.0, L, NX2), DEG)
    nd = NX2 * NX2
    e, b = [np.zeros(nd) for _ in range(3)], [np.zeros(nd) for _ in range(3)]
    h = gp.HamiltonianSplitting2D3V(mx, pg, e, b, resident=resident)
    h.set_sort_interval(sort_interval)
    h.set_fusion(bool(fuse))
    rho = h.charge_density()
    mx.compute_e_from_rho(e, rho - rho.mean())
    b[2][:] = BETA * np.cos(2 * np.pi * (np.arange(nd) % NX2) / NX2)
    if resident:
        h.upload_fields()
    return dict(wl=wl, pg=pg, mx=mx, e=e, b=b, h=h, rho=rho, total_charge=float(rho.sum()), L=L, n_local=n_local,
                n_global=n_global, nd=nd)


def roofline_block(prof, n_local, ms_per_step, integrator, exclude=()):
    peak, peak_src = peaks()
    passes = {k: v for k, v in prof.items() if k in BYTES and v[1] > 0 and k not in exclude}
    if not passes:
        return None
    dom = max(passes, key=lambda k: passes[k][0])   # the pass with the largest share of the step
    t_ms = prof[dom][0] / prof[dom][1]
    achieved = BYTES[dom] * n_local / (t_ms * 1e-3) / 1e9
    tr = ncu_traffic(dom)
    return {"bound": "hbm", "kernel": f"pass<{dom}>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (tr["dram_bytes_per_particle"] * n_local if tr else None),
            "traffic_note": (f"NOT measured in this run: {tr['dram_bytes_per_particle']} B/particle (ncu dram__bytes_read+write) "
                             f"from the committed capture {tr['source']} x {n_local} particles" if tr else None),
            "peak_source": peak_src, "algorithmic_bytes_per_particle": BYTES[dom], "avg_launch_ms": t_ms,
            "all_passes": {k: {"avg_ms": v[0] / max(v[1], 1), "launches": v[1], "bytes_per_particle": BYTES.get(k),
                               "GBps": BYTES[k] * n_local / (v[0] / v[1] * 1e-3) / 1e9 if v[1] and k in BYTES else None,
                               "frac": BYTES[k] * n_local / (v[0] / v[1] * 1e-3) / 1e9 / peak if v[1] and k in BYTES else None}
                           for k, v in prof.items()},
            "step_GBps_vs_unfused_bytes": STEP_BYTES[integrator] * n_local / (ms_per_step * 1e-3) / 1e9}


def measure(B, S, steps, warmup, min_seconds, with_clocks=True):
    """device-resident regions of `steps` Strang steps of the splitting S['h']"""
    gp, dc = B.gp, B.dc
    h = S["h"]
    h.strang_splitting(DT, warmup)
    gp.synchronize()
    sampler = ClockSampler(dc.local_rank).start() if (dc.rank == 0 and with_clocks) else None
    B.profile(True)
    gp.launch_count(reset=True)
    ms = B.regions(lambda: h.strang_splitting(DT, steps), steps, min_seconds)
    launches = gp.launch_count()
    prof = B.profile_read()
    B.profile(False)
    clocks = sampler.stop() if sampler else None
    med = statistics.median(ms)
    n_global = S["n_global"]
    return {"ms_regions": ms, "ms_per_step": med / steps, "value": n_global * steps / (med * 1e-3),
            "sustained": {"value": n_global * steps * len(ms) / (sum(ms) * 1e-3), "seconds": sum(ms) * 1e-3, "regions": len(ms),
                          "ms_per_step_min": min(ms) / steps, "ms_per_step_max": max(ms) / steps},
            "launches_per_region": launches / len(ms), "prof": prof, "clocks": clocks}


def e2e_1d(B, S, steps, fuse):
    """the reference-facing call with HOST field buffers, one step per call, wall clock, max over ranks"""
    gp, dc = B.gp, B.dc
    S["h"].sync_fields()
    if S["wl"]["integrator"] == "boris":
        # the staggered mid fields live in the splitting object, so the host variant is re-staggered from the current state
        hh = gp.HamiltonianSplittingBoris(S["mx"], S["ks0"], S["ks1"], S["pg"], [S["e1"], S["e2"]], S["b"], resident=False)
        hh.staggering(DT)
    else:
        hh = gp.HamiltonianSplitting(1, 2, S["mx"], S["ks0"], S["ks1"], S["pg"], [S["e1"], S["e2"]], S["b"], resident=False, fuse=fuse)
    for _ in range(3):
        hh.strang_splitting(DT, 1)
    # like the device-resident measurement: regions of `steps` calls repeated until >= min_seconds, median region reported
    secs = []
    while sum(secs) < B.args.min_seconds and len(secs) < 64:
        dc.barrier()
        gp.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            hh.strang_splitting(DT, 1)      # synchronous: H2D fields, pass(es) + solves, D2H fields
        gp.synchronize()
        secs.append(dc.max_over_ranks(time.perf_counter() - t0))
    sec = statistics.median(secs)
    out = {"value": S["n_global"] * steps / sec, "unit": "particle-steps/s", "h2d_bytes_per_step": 3 * NX * 8,
           "d2h_bytes_per_step": 3 * NX * 8, "steps": steps, "regions": len(secs),
           "first_region_value": S["n_global"] * steps / secs[0],
           "what": ("gempic_boris_strang_splitting_host" if S["wl"]["integrator"] == "boris" else "gempic_hs_strang_splitting_host") +
                   ": host e1,e2,b in, e1,e2,b out, synchronous; median of the regions (wall clock, max over ranks)"}
    S["h_host"] = hh
    return out


def e2e_loop_1d(B, S, steps):
    """the example's loop body per step (examples/strong_landau_damping_1d2v.jl:46-59): strang_splitting!(h, dt, 1);
    solve_poisson!(efield_poisson, ...); write_step!(thdiag, ...) -- all through host buffers"""
    gp, dc = B.gp, B.dc
    hh = S["h_host"]
    th = gp.TimeHistoryDiagnostics(S["pg"], S["mx"], S["ks0"], S["ks1"])
    ep, rho = np.zeros(NX), np.zeros(NX)
    e_n = [S["e1"].copy(), S["e2"].copy()]

    def body(j):
        e_n[0][:], e_n[1][:] = S["e1"], S["e2"]
        hh.strang_splitting(DT, 1)
        gp.solve_poisson(ep, S["pg"], S["ks0"], S["mx"], rho)
        gp.write_step(th, j * DT, DEG, [S["e1"], S["e2"]], S["b"], e_n, ep)

    for j in range(3):
        body(j)
    dc.barrier()
    gp.synchronize()
    B.profile(True)
    t0 = time.perf_counter()
    for j in range(steps):
        body(j)
    gp.synchronize()
    sec = dc.max_over_ranks(time.perf_counter() - t0)
    prof = B.profile_read()
    B.profile(False)
    bytes_pp = sum(BYTES.get(k, 0) * v[1] for k, v in prof.items()) / max(steps, 1)
    row = th.data[-1]
    assert np.all(np.isfinite(row)) and row[1] > 0.0, "write_step! returned garbage"
    return {"value": S["n_global"] * steps / sec, "unit": "particle-steps/s", "steps": steps,
            "particle_bytes_per_step": bytes_pp,
            "passes_per_step": {k: v[1] / steps for k, v in prof.items() if v[1]},
            "pass_ms": {k: v[0] / v[1] for k, v in prof.items() if v[1]}, "ms_per_step": sec / steps * 1e3,
            "h2d_bytes_per_step": (3 + 6) * NX * 8, "d2h_bytes_per_step": (3 + 2) * NX * 8 + 11 * 8,
            "what": "strang_splitting!(h, dt, 1); solve_poisson!; write_step! per step, host buffers, synchronous"}


def config_entry(B, workload, n_local, steps, min_seconds, fuse=True, sort_interval=1):
    """a short device-resident measurement of another BASELINE config at this GPU count"""
    integ = WORKLOADS[workload]["integrator"]
    S = build_2d(B, n_local, fuse=fuse, sort_interval=sort_interval) if integ == "hs2d" else build_1d(B, workload, n_local, fuse=fuse)
    M = measure(B, S, steps, 3, min_seconds)
    peak, _ = peaks()
    roof = roofline_block(M["prof"], n_local, M["ms_per_step"], integ, exclude=("cell histogram after Hp2", "cell sort 2d"))
    # sanity: still a simulation
    if integ == "hs2d":
        S["h"].sync_fields()
        assert all(np.all(np.isfinite(v)) for v in S["e"] + S["b"])
        assert abs(S["total_charge"] - S["L"] ** 2) < 1e-9 * S["L"] ** 2
    else:
        rho2, ep = np.zeros(NX), np.zeros(NX)
        B.gp.solve_poisson(ep, S["pg"], S["ks0"], S["mx"], rho2)
        assert abs(rho2.sum() - S["L"]) < 1e-9 * S["L"], "charge is not conserved"
    out = {"workload": WORKLOADS[workload]["name"], "particles_per_gpu": int(n_local), "n_gpus": B.dc.world_size,
           "value": M["value"], "unit": "particle-steps/s", "ms_per_step": M["ms_per_step"], "steps": steps,
           "regions": M["sustained"]["regions"], "sustained_value": M["sustained"]["value"],
           "passes": ({k: {"avg_ms": v["avg_ms"], "frac": v["frac"], "bytes_per_particle": v["bytes_per_particle"]}
                       for k, v in roof["all_passes"].items() if v["frac"]} if roof else None),
           "dominant_pass": roof["kernel"] if roof else None, "frac": roof["frac"] if roof else None,
           "step_frac_of_peak": (sum(BYTES.get(k, 0) * v[1] for k, v in M["prof"].items()) / max(len(M["ms_regions"]) * steps, 1)
                                 * n_local / (M["ms_per_step"] * 1e-3) / 1e9 / peak),
           "clocks": M["clocks"]}
    del S
    gc.collect()
    return out


def run_ours(args):
    B = Bench(args)
    gp, dc = B.gp, B.dc
    workload = args.workload
    wl = WORKLOADS[workload]
    integ = wl["integrator"]
    two_d = integ == "hs2d"
    n_local = args.particles if args.particles else wl["particles"]
    fuse = bool(args.fuse)

    # ---- sharded parity on a small sub-problem (N > 1) ------------------------------------------------
    parity = None
    if dc.world_size > 1 and not args.no_parity:
        errs = gp.sharded_parity(dc)
        worst = dc.max_over_ranks(max(errs.values()) if errs else 0.0)
        parity = {"max_rel_err": worst, "nranks": dc.world_size, "tolerance": 1e-11, "cases": errs if dc.rank == 0 else None,
                  "what": "sharded vs un-sharded HamiltonianSplitting{1,2} (fused / per operator), HamiltonianSplittingBoris "
                          "(fused step / separate pushes), HamiltonianSplitting{2,3} (fused / per operator): fields, currents, "
                          "rank 0's particles; 400003 / 120001 particles, 3 steps"}
        assert worst < 1e-11, f"sharded parity failed: {errs}"

    # ---- value: device-resident ---------------------------------------------------------------------
    S = build_2d(B, n_local, fuse=fuse, sort_interval=args.sort_interval) if two_d else build_1d(B, workload, n_local, fuse=fuse)
    if two_d:
        r0 = S["h"].gauss_residual()
    M = measure(B, S, args.steps, args.warmup, args.min_seconds)

    # ---- e2e: host field buffers every step ---------------------------------------------------------
    if two_d:
        S["h"].sync_fields()
        hh = gp.HamiltonianSplitting2D3V(S["mx"], S["pg"], S["e"], S["b"], resident=False)
        hh.set_sort_interval(args.sort_interval)
        hh.set_fusion(fuse)
        for _ in range(2):
            hh.strang_splitting(DT, 1)
        dc.barrier(); gp.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            hh.strang_splitting(DT, 1)
        gp.synchronize()
        sec = dc.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": S["n_global"] * args.steps / sec, "unit": "particle-steps/s", "h2d_bytes_per_step": 6 * S["nd"] * 8,
               "d2h_bytes_per_step": 6 * S["nd"] * 8, "steps": args.steps,
               "what": "gempic_hs2d_strang_splitting_host: host e[3], b[3] in and out, synchronous"}
        e2e_loop = None
        # sanity: the timed state is a real simulation (charge exact, Gauss law conserved, finite fields)
        S["h"].upload_fields()
        r1 = S["h"].gauss_residual()
        assert abs(S["total_charge"] - S["L"] ** 2) < 1e-9 * S["L"] ** 2, "total charge is wrong"
        assert np.max(np.abs(r1 - r0)) < 1e-10 * np.max(np.abs(S["rho"])), "discrete Gauss law violated -- the step did not do its work"
        assert all(np.all(np.isfinite(v)) for v in S["e"] + S["b"])
    else:
        e2e = e2e_1d(B, S, args.steps, fuse)
        e2e_loop = e2e_loop_1d(B, S, min(args.steps, 30)) if integ == "hs" else None
        rho2, ep = np.zeros(NX), np.zeros(NX)
        gp.solve_poisson(ep, S["pg"], S["ks0"], S["mx"], rho2)
        assert abs(rho2.sum() - S["L"]) < 1e-9 * S["L"] and abs(S["total_charge"] - S["L"]) < 1e-9 * S["L"], \
            "charge is not conserved -- the step did not do its work"
        assert np.all(np.isfinite(S["e1"])) and np.all(np.isfinite(S["b"]))

    roof = roofline_block(M["prof"], n_local, M["ms_per_step"], integ, exclude=("cell histogram after Hp2", "cell sort 2d"))
    del S
    gc.collect()

    # ---- the other BASELINE configs at this GPU count + one strong-scaling point ------------------------
    configs = None
    if not args.no_configs:
        configs = {}
        for w in ("landau", "boris", "2d3v", "weibel"):
            if w == workload:
                continue
            configs[w] = config_entry(B, w, WORKLOADS[w]["particles"], 20 if w != "2d3v" else 10, 0.4,
                                      sort_interval=args.sort_interval)
        n_strong = 100_000_000 // dc.world_size
        configs["strong_weibel_1e8_total"] = config_entry(B, "weibel", n_strong, 50, 0.3)
        configs["strong_weibel_1e8_total"]["scaling"] = "strong: 1e8 particles in total over the N GPUs"

    if dc.rank == 0:
        cpu = None
        if args.gpus == 1 and not args.no_cpu:
            cpu = cpu_baseline_block(args, workload)
        line = {
            "metric": "particle-steps/s per Strang step", "value": M["value"], "unit": "particle-steps/s", "n_gpus": dc.world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": M["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(workload, args.gpus, n_local, fuse, sort_interval=args.sort_interval),
            "timing": {"regions": M["sustained"]["regions"], "region_ms": M["ms_regions"],
                       "rule": "K-step region repeated until >= %.1f s of device time; value = median region" % args.min_seconds},
            "sustained": M["sustained"],
            "e2e": e2e, "e2e_loop": e2e_loop,
            "gpu_launches": int(round(M["launches_per_region"])), "clocks": M["clocks"], "roofline": roof, "cpu_baseline": cpu,
            "sharded_parity": parity, "configs": configs,
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
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--cpu-particles", type=int, default=4_000_000, help="bounded sample for the CPU arm")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-vs-unsharded self-check at N > 1")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="repeat the timed region until this much device time")
    ap.add_argument("--workload", default="weibel", choices=sorted(WORKLOADS), help="weibel = BASELINE configs[1] (the metric's config)")
    ap.add_argument("--sort-interval", type=int, default=1, help="2d3v: cell sort every k Strang steps")
    ap.add_argument("--fuse", type=int, default=1, help="1: fused particle passes (default), 0: one kernel per reference operator")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3   # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

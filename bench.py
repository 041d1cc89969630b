#!/usr/bin/env python3
"""Benchmark of the per-SCF-iteration grid hot path (XC + Hartree build) — BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path for a fixed density matrix P: rho = 2 phi^T P phi + rescale, LDA
pointwise, Becke/Poisson Hartree potential, and the [J | XC] contractions (reference src/dft.cpp:100-102 minus
the host eigen-solve).  Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).

 value  : ms per step with P already resident in HBM (dftgrid_iteration_device), timed with CUDA events on the
          library's stream between barriers, max over ranks.
 e2e    : the same through the public C ABI with HOST buffers (dftgrid_iteration: P host->device, [J|XC|E_xc|N]
          device->host inside the timed region), wall clock between device synchronisations, max over ranks.
 roofline: the dominant kernel (the [J|XC] DMMA contraction) against the FP64 tensor peak measured on this pool.
 cpu_baseline / --impl reference: the unmodified reference classes (oracle/_ref) timed on this box's host cores.

Multi-GPU (torchrun, one process per GPU): grid points are sharded by (atom, radial shell); the work of one
molecule is split, so scaling is STRONG.  torch.distributed carries only the NCCL id and the timing reductions.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from dftcxx_b200 import systems  # noqa: E402
from dftcxx_b200.molecule import DATA, Molecule  # noqa: E402

FP64_TENSOR_PEAK_TFLOPS = 37.05  # measured: tools/microbench/fp64_peak.cu on this pool (profiles/r01_fp64_peak_microbench.txt)
FILE_WORKLOADS = {
    "h2o_sto3g": "h2o_sto3g.in", "benzene": "benzene_p631_fine.in", "ethane": "ethane_p631_fine.in",
    "ch4": "ch4_p631_fine.in", "ch4_dense422": "ch4_p631_dense422.in",
}
METRIC = "xc_hartree_build_ms_per_scf_iter"


def load_workload(name):
    if name in FILE_WORKLOADS:
        mol = Molecule.from_file(os.path.join(DATA, "molecules", FILE_WORKLOADS[name]))
        st = mol.settings
        return mol, (st.radial_points, st.lebedev_order, st.lmax)
    fac, prm = systems.WORKLOADS[name]
    return fac(), prm


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        if not sm:
            return None
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        smax = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm
def reference_fit(workload, mol, prm, threads):
    """CPU time of ONE iteration of the reference for `workload`.

    Small workloads (<= 40k points) run whole.  The big synthetic clusters cannot (the reference needs of the order of
    10 min per iteration plus ~20 min of serial grid construction for (H2O)64), so a BOUNDED SAMPLE is timed instead:
    the first m molecules of the SAME cluster (same geometry generator, basis and grid), m = 4 and 8, through the
    reference's own classes.  Its per-phase times are scaled with the reference's loop bounds:
        Hartree phase  ~ a * Npts*(Natoms-1)*nlm   (src/moleculargrid.cpp:342-380, the 78 % hot loop)
        rho + XC phase ~ b * Npts*nb^2             (src/gridpoint.cpp:82-84, src/dft.cpp:424-432)
    with a, b taken from the larger sample (the smaller one is reported as a consistency check).  The nb^2 J assembly
    hidden inside the Hartree phase is NOT scaled up, so the extrapolation is a lower bound of the reference's time.
    Returns (ms, description, details)."""
    from oracle import refpy

    os.environ["OMP_NUM_THREADS"] = str(threads)
    nr, lo, lm = prm
    nlm = (lm + 1) ** 2
    nang = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194][lo]

    def run(m):
        extra = ("radial_points = %d" % nr, "lebedev_order = %d" % lo, "lmax = %d" % lm)
        with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
            f.write(m.to_input(grid=None, extra=extra))
            path = f.name
        try:
            r = refpy.Ref(path, full=False, fast=refpy.available(fast=True))
            P = systems.synthetic_density(m)
            best = None
            for _ in range(2):
                t, ph, _, _, _ = r.time_iteration(P)
                if best is None or t < best[0]:
                    best = (t, ph.copy())
            r.close()
        finally:
            os.remove(path)
        return best

    npts = mol.natoms * nr * nang
    if npts <= 40000:
        ms, _ = run(mol)
        return ms, "whole workload, 1 iteration (best of 2)", {}
    sizes = (4, 8)
    sub = [systems.water_cluster(m) if workload.startswith("h2o") else systems.alkane(m) for m in sizes]
    coef, ts = [], []
    for m in sub:
        n = m.natoms * nr * nang
        t, ph = run(m)
        ts.append(t)
        coef.append((ph[1] / (n * (m.natoms - 1) * nlm), (ph[0] + ph[2] + ph[3]) / (n * m.nbf ** 2)))
    a, b = coef[-1]
    full = a * npts * (mol.natoms - 1) * nlm + b * npts * mol.nbf ** 2
    full_small = coef[0][0] * npts * (mol.natoms - 1) * nlm + coef[0][1] * npts * mol.nbf ** 2
    desc = ("bounded sample: sub-clusters of the same geometry with %d and %d molecules (%.1f s and %.1f s of CPU per iteration); per-phase "
            "times of the larger one scaled by the reference's loop bounds (Hartree ~ Npts*(Natoms-1)*nlm, rho+XC ~ Npts*nb^2) to the full "
            "workload: EXTRAPOLATED lower bound (the smaller sample extrapolates to %.0f s)" % (sizes[0], sizes[1], ts[0] / 1e3, ts[1] / 1e3, full_small / 1e3))
    return float(full), desc, {"a_ms": a, "b_ms": b, "sample_ms": ts}


def run_reference(args, mol, prm, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refpy

    threads = os.cpu_count() or 1
    kind = "reference" if refpy.available() else "unavailable"
    if kind == "unavailable":
        emit({"impl": "reference", "unavailable": "oracle/_ref was not built (reference sources absent at build time)"})
        return
    vals = []
    desc = ""
    for _ in range(max(1, min(args.steps, 2))):
        ms, desc, det = reference_fit(args.workload, mol, prm, threads)
        vals.append(ms)
    v = float(np.median(vals))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "ms", "n_gpus": args.gpus, "steps": len(vals), "warmup": 0,
           "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args.workload, mol, prm),
           "cpu_baseline": {"value": v, "unit": "ms", "cores": threads, "kind": "reference", "sample": desc},
           "e2e": {"value": v, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def workload_config(name, mol, prm):
    nr, lo, lm = prm
    nang = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194][lo]
    return {"workload": "%s / %s / %dx%d grid, lmax %d" % (mol.name or name, mol.basis, nr, nang, lm), "name": name,
            "natoms": int(mol.natoms), "nbf": int(mol.nbf), "npoints": int(mol.natoms * nr * nang), "nlm": (lm + 1) ** 2,
            "l2_policy": "inputs exceed L2 (Phi is %.2f GB vs 126 MB)" % (mol.natoms * nr * nang * mol.nbf * 8 / 1e9)
            if mol.natoms * nr * nang * mol.nbf * 8 > 2 * 126e6 else "L2 flushed between steps by a 256 MB device memset"}


# ------------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="h2o64")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-peer", action="store_true", help="multi-GPU: sum [J | XC] with ncclAllReduce instead of the peer-memory kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # Libraries (NCCL, torch) print banners on stdout; the contract is ONE JSON line there.  Everything written to fd 1
    # while the benchmark runs is diverted to stderr and the JSON line goes to the saved descriptor at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    mol, prm = load_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args, mol, prm, emit)

    import torch

    from dftcxx_b200.grid import MolecularGrid, comm_unique_id

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE under torchrun")
    dist = None
    comm_id = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm_id = box[0]

    def barrier():
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    g = MolecularGrid(mol, device=local, rank=rank, nranks=world)
    g.set_grid_parameters(*prm)
    t0 = time.time()
    g.create_grid(comm_id)
    build_wall = time.time() - t0
    peer_path = False
    if world > 1 and not args.no_peer:
        # the [J | XC] sum runs in the library's own peer-memory kernels when the ranks' GPUs have a P2P path
        peer_path = g.connect_peers(dist)  # all ranks agree: peer path only if every rank mapped every buffer
    tb = g.timings()
    P = systems.synthetic_density(mol)
    flush = None
    small = g.npoints * mol.nbf * 8 <= 2 * 126e6
    if small:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % local)

    # ---- value: P resident, device-timed
    g.upload_density(P)
    for _ in range(args.warmup):
        g.iteration_device()
    # the clock sampler forks nvidia-smi (tens of ms from a process this size): start it BEFORE the barrier, otherwise the
    # other ranks' timers include rank 0's fork while they wait for it in the first collective
    sampler = ClockSampler(local) if rank == 0 else None
    g.synchronize()
    barrier()
    n0 = g.launch_count()
    if small:
        # small inputs: flush L2 between steps, time each step separately
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            g.timer_start()
            g.iteration_device()
            tot += g.timer_stop()
        dev_ms = tot
    else:
        g.timer_start()
        for _ in range(args.steps):
            g.iteration_device()
        dev_ms = g.timer_stop()
    g.synchronize()
    barrier()
    launches = g.launch_count() - n0
    phases = g.timings()
    ms_step = max_over_ranks(dev_ms / args.steps)

    # ---- e2e: host buffers through the C ABI.  P and the result matrices live in pinned host memory (the SCF
    # driver's own buffers); every step copies P host->device and [J | XC | E_xc | N_el] device->host.
    P_host = torch.from_numpy(np.ascontiguousarray(P)).pin_memory().numpy()
    out_host = (torch.empty((mol.nbf, mol.nbf), dtype=torch.float64).pin_memory().numpy(),
                torch.empty((mol.nbf, mol.nbf), dtype=torch.float64).pin_memory().numpy())
    for _ in range(2):
        g.iteration(P_host, out=out_host)
    g.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if small:
            flush.zero_()
            torch.cuda.synchronize()
        J, XC, exc, nel = g.iteration(P_host, out=out_host)
    g.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel ([J|XC] contraction), live event timings of the last step
    npts_loc = g.nloc
    flops_contract = 2.0 * npts_loc * mol.nbf * (mol.nbf + 1)  # 2 symmetric matrices, Npts*nb*(nb+1) flop each
    con_ms = max_over_ranks(phases["contract"])
    peaks, peak_kind = measured_peaks()
    ach = flops_contract / (con_ms * 1e-3) / 1e12
    roof = {"kernel": "k_contract_tma (+k_contract_reduce)", "bound": "tensor", "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS,
            "unit": "TFLOP/s", "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": None,
            "peak_source": "FP64 DMMA peak measured on this pool with tools/microbench/fp64_peak.cu (MEASURED_PEAKS.json has no FP64 entry; "
                           "its HBM figure %s GB/s is the denominator for the streaming kernels, %s)" % (peaks.get("hbm_gbs"), peak_kind),
            "ms": con_ms}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get(args.workload, {}).get("k_contract_dram_bytes")
        except Exception:
            pass

    if rank == 0:
        out = {"metric": METRIC, "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": workload_config(args.workload, mol, prm),
               "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(mol.nbf ** 2 * 8),
                       "d2h_bytes_per_step": int((2 * mol.nbf ** 2 + 2) * 8)},
               "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
               "collectives": ("none (single GPU)" if world == 1 else
                               "shell sums + rho_lm: ncclAllReduce; [J | XC]: " + ("peer-memory kernels (k_contract_reduce_publish + k_peer_sum)"
                                                                                  if peer_path else "ncclAllReduce")),
               "phases_ms": {k: round(v, 4) for k, v in phases.items()},
               "grid_build": {"wall_s": round(build_wall, 3), "becke_ms": tb["becke"], "phi_ms": tb["phi"],
                              "becke_cell_functions_per_s": g.nloc * mol.natoms * (mol.natoms - 1) / (tb["becke"] * 1e-3) * world
                              if tb["becke"] > 0 else None,
                              "phi_gridpt_basis_evals_per_s": g.nloc * mol.nbf / (tb["phi"] * 1e-3) * world if tb["phi"] > 0 else None,
                              "phi_hbm_write_frac": (g.nloc * mol.nbf * 8 / (tb["phi"] * 1e-3) / 1e9) / peaks.get("hbm_gbs", 6650.0)
                              if tb["phi"] > 0 else None},
               "results": {"exc": exc, "nelec": nel}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                ms, desc, det = reference_fit(args.workload, mol, prm, os.cpu_count() or 1)
                out["cpu_baseline"] = {"value": ms, "unit": "ms", "cores": os.cpu_count() or 1, "kind": "reference", "sample": desc}
            except Exception as e:  # the oracle is a checker, never a dependency of the measured path
                out["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count() or 1, "kind": "unavailable", "sample": str(e)[:200]}
        emit(out)
    g.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Benchmark of the per-SCF-iteration grid hot path (XC + Hartree build) — BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one pass of the hot path for a fixed density matrix P: rho = 2 phi^T P phi + rescale, LDA
pointwise, Becke/Poisson Hartree potential, and the contraction of the grid's Fock contribution F_grid = 2J + XC
with E_J, E_xc and the electron count (reference src/dft.cpp:100-102 minus the host eigen-solve; src/dft.cpp:334 is
the only consumer of J and XC and needs their sum).  Prints ONE JSON line on rank 0 (DESIGN.md "Measurement").

 value  : ms per step with P already resident in HBM (dftgrid_fock_device), timed with CUDA events on the
          library's stream between barriers, max over ranks.  pair_ms_per_step: the same with J and XC as two matrices.
 e2e    : the same through the public C ABI with HOST buffers (dftgrid_fock: P host->device, [F|E_J|E_xc|N]
          device->host inside the timed region), wall clock between device synchronisations, max over ranks.
 roofline: the dominant kernel (the F_grid DMMA contraction) against the FP64 tensor peak measured on this pool.
 results.parity: J / XC / F of this very run against the fixture the unmodified reference produced for this workload
          (tests/golden); the run exits non-zero when a BASELINE tolerance is exceeded.
 scf    : whole SCF iterations of the drop-in C++ host on this workload (device-resident algebra + grid path).
 cpu_baseline / --impl reference: the unmodified reference classes (oracle/_ref) timed on this box's host cores.

Multi-GPU: grid points are sharded by (atom, radial shell); the work of one molecule is split, so scaling is STRONG.
Under torchrun (one process per GPU) torch.distributed carries only the NCCL id and the timing reductions; without
torchrun `--gpus N` drives N devices from this one process (dftgrid_create_multi).
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
def reference_fit(workload, mol, prm, threads, whole=False):
    """CPU time of ONE iteration of the reference for `workload`.

    Small workloads (<= 40k points), or any workload with whole=True, run whole.  The big synthetic clusters cannot within
    a bench run (the unmodified reference needs of the order of an hour per (H2O)64 iteration plus its serial grid
    construction), so a BOUNDED SAMPLE is timed instead: the first m molecules of the SAME cluster (same geometry
    generator, basis and grid) for three sizes m, through the reference's own classes.  The per-phase times are fitted by
    least squares (through the origin) to the reference's loop bounds
        Hartree phase  ~ a * Npts*(Natoms-1)*nlm   (src/moleculargrid.cpp:342-380, the 78 % hot loop)
        rho + XC phase ~ b * Npts*nb^2             (src/gridpoint.cpp:82-84, src/dft.cpp:424-432)
    and evaluated at the full size; the relative residuals of the fit at the three sizes are reported.  The nb^2 J assembly
    hidden inside the Hartree phase is NOT scaled up, so the extrapolation is a lower bound of the reference's time.  Note
    that the nb^2 terms of oracle/_ref run on the header shim (the image has no Eigen).  The shim serves rows of large column-major
    matrices from a row-major mirror, so the reference's nb^2 strided dot products (src/dft.cpp:424-432,
    src/atomicgrid.cpp:479-487) read contiguous memory — a real Eigen build walks them with stride nb and is SLOWER; the CPU
    figure is therefore on the fast side, the GPU/CPU ratio on the conservative side.
    Returns (ms, description, details)."""
    from oracle import refpy

    os.environ["OMP_NUM_THREADS"] = str(threads)
    nr, lo, lm = prm
    nlm = (lm + 1) ** 2
    nang = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194][lo]

    def run(m, reps=2):
        extra = ("radial_points = %d" % nr, "lebedev_order = %d" % lo, "lmax = %d" % lm)
        with tempfile.NamedTemporaryFile("w", suffix=".in", delete=False) as f:
            f.write(m.to_input(grid=None, extra=extra))
            path = f.name
        try:
            r = refpy.Ref(path, full=False, fast=refpy.available(fast=True))
            P = systems.synthetic_density(m)
            best = None
            for _ in range(reps):
                t, ph, _, _, _ = r.time_iteration(P)
                if best is None or t < best[0]:
                    best = (t, ph.copy())
            r.close()
        finally:
            os.remove(path)
        return best

    npts = mol.natoms * nr * nang
    if npts <= 40000 or whole:
        t0 = time.time()
        ms, ph = run(mol, reps=1 if whole and npts > 40000 else 2)
        return ms, "whole workload, 1 iteration (measured; %.0f s including the reference's grid construction)" % (time.time() - t0), \
            {"extrapolated": False, "measured_s": ms / 1e3, "phases_ms": [float(x) for x in ph]}
    sizes = (4, 8, 16) if workload.startswith("h2o") else (4, 8, 12)
    sub = [systems.water_cluster(m) if workload.startswith("h2o") else systems.alkane(m) for m in sizes]
    xa, ya, xb, yb, ts = [], [], [], [], []
    for m in sub:
        n = m.natoms * nr * nang
        t, ph = run(m, reps=1 if m is sub[-1] else 2)
        ts.append(t)
        xa.append(n * (m.natoms - 1) * nlm)
        ya.append(ph[1])
        xb.append(n * m.nbf ** 2)
        yb.append(ph[0] + ph[2] + ph[3])
    xa, ya, xb, yb = (np.array(v, dtype=float) for v in (xa, ya, xb, yb))
    a = float(xa @ ya / (xa @ xa))
    b = float(xb @ yb / (xb @ xb))
    resid = [float((a * xa[i] + b * xb[i]) / ts[i] - 1.0) for i in range(len(sizes))]
    full = a * npts * (mol.natoms - 1) * nlm + b * npts * mol.nbf ** 2
    lo_hi = [float(ya[i] / xa[i] * npts * (mol.natoms - 1) * nlm + yb[i] / xb[i] * npts * mol.nbf ** 2) for i in range(len(sizes))]
    desc = ("bounded sample: sub-clusters of the same geometry with %s molecules (%s s of CPU per iteration, measured); per-phase times "
            "least-squares fitted to the reference's loop bounds (Hartree ~ Npts*(Natoms-1)*nlm, rho+XC ~ Npts*nb^2), fit residuals %s; "
            "evaluated at the full workload: EXTRAPOLATED lower bound (single-sample extrapolations: %s s); the nb^2 terms run on the "
            "header shim, whose row-major mirror makes the reference's strided nb^2 dot products contiguous (a real Eigen build is slower there: "
            "the CPU figure errs on the fast side); the same fit on the build container's 8 cores gave 225.7 s for (H2O)32 and 1281.8 s for (H2O)64 against "
            "214.2 s and 1331.5 s for the WHOLE workloads really run there (profiles/r02e_bench_reference_whole_*_build_container.json)" % ("/".join(str(x) for x in sizes), "/".join("%.1f" % (t / 1e3) for t in ts),
                                                         "/".join("%+.1f%%" % (100 * r) for r in resid), "/".join("%.0f" % (v / 1e3) for v in lo_hi)))
    return float(full), desc, {"extrapolated": True, "measured_s": [t / 1e3 for t in ts], "sample_molecules": list(sizes), "a_ms": a, "b_ms": b,
                               "fit_residuals": resid, "single_sample_extrapolations_ms": lo_hi}


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
    ms, desc, det = reference_fit(args.workload, mol, prm, threads, whole=args.whole)
    v = float(ms)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "ms", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
           "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args.workload, mol, prm),
           "cpu_baseline": dict({"value": v, "unit": "ms", "cores": threads, "kind": "reference", "sample": desc}, **det),
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
GOLDEN_OF = {"h2o64": "h2o64_p631_fine", "h2o32": "h2o32_p631_fine", "c40h82_fine": "c40h82_p631_fine", "h2o8": "h2o8_p631_fine",
             "benzene": "benzene_p631_fine", "ethane": "ethane_p631_fine", "ch4": "ch4_p631_fine", "h2o_sto3g": "h2o_sto3g",
             "ch4_dense422": "ch4_p631_dense422"}


def load_fixture(workload, nbf):
    """Reference outputs for this workload's synthetic P (tests/golden, generated from the unmodified reference), or None."""
    name = GOLDEN_OF.get(workload)
    path = os.path.join(ROOT, "tests", "golden", (name or "") + ".npz")
    if not name or not os.path.exists(path):
        return None
    z = np.load(path)
    out = {"fixture": "tests/golden/%s.npz" % name}
    iu = np.triu_indices(nbf)
    for k in ("J", "XC"):
        if k in z:
            out[k] = z[k]
        else:
            M = np.zeros((nbf, nbf))
            M[iu] = z[k + "_triu"]
            out[k] = M + np.triu(M, 1).T
    out["exc"], out["nel"] = float(z["exc"]), float(z["nel"])
    out["P_checksum"] = z["P_checksum"] if "P_checksum" in z else None
    out["P"] = z["P"] if "P" in z else None
    return out


def parity_block(g, P, fx):
    """J / XC / F of one more iteration (all ranks take part) against the reference fixture.  BASELINE tolerances:
    J, XC 1e-10 absolute; F = 2J + XC accordingly 3e-10; E_xc 1e-9; electron count 1e-9."""
    J, XC, exc, nel = g.iteration(P)
    F, ej, exc2, nel2 = g.fock(P)
    Fr = 2.0 * fx["J"] + fx["XC"]
    ejr = 2.0 * float(np.einsum("ij,ij->", P, fx["J"]))
    r = {"fixture": fx["fixture"], "max_abs_dJ": float(np.max(np.abs(J - fx["J"]))), "max_abs_dXC": float(np.max(np.abs(XC - fx["XC"]))),
         "max_abs_dF": float(np.max(np.abs(F - Fr))), "dExc": float(abs(exc - fx["exc"])), "dNel": float(abs(nel - fx["nel"])),
         "dEJ_rel": float(abs(ej - ejr) / max(1.0, abs(ejr))), "tol": {"J": 1e-10, "XC": 1e-10, "F": 3e-10, "Exc": 1e-9, "Nel": 1e-9, "EJ_rel": 1e-12}}
    r["ok"] = bool(r["max_abs_dJ"] <= 1e-10 and r["max_abs_dXC"] <= 1e-10 and r["max_abs_dF"] <= 3e-10 and r["dExc"] <= 1e-9 and r["dNel"] <= 1e-9
                   and r["dEJ_rel"] <= 1e-12)
    return r


def scf_block(workload, ngpus, iters=3):
    """Whole SCF iterations of the drop-in C++ host (`dftcxx -i`, libdfthost.so) on this workload: device-resident algebra
    + fused grid build per iteration.  Wall-clock per iteration as the host's own timer prints it."""
    import ctypes

    name = GOLDEN_OF.get(workload)
    infile = os.path.join(DATA, "molecules", (name or "") + ".in")
    lib = os.path.join(ROOT, "dftcxx_b200", "libdfthost.so")
    if not name or not os.path.exists(infile) or not os.path.exists(lib):
        return None
    L = ctypes.CDLL(lib)
    dp = ctypes.POINTER(ctypes.c_double)
    L.dfthost_scf2.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp]
    L.dfthost_last_timings.argtypes = [dp, ctypes.c_int]
    L.dfthost_last_error.restype = ctypes.c_char_p
    e = np.zeros((iters, 6))
    t0 = time.time()
    n = L.dfthost_scf2(infile.encode(), 0, ngpus, 0, iters, iters, e.ctypes.data_as(dp), None, None)
    wall = time.time() - t0
    if n != iters:
        return {"error": L.dfthost_last_error().decode()[:200]}
    t = np.zeros((iters, 4))
    L.dfthost_last_timings(t.ctypes.data_as(dp), iters)
    return {"input": "dftcxx_b200/data/molecules/%s.in" % name, "iterations": iters, "ms_per_iteration_wall": [round(x, 3) for x in t[:, 0]],
            "device_algebra_ms": [round(x, 3) for x in t[:, 1]], "device_grid_ms": [round(x, 3) for x in t[:, 2]],
            "purification_steps": [int(x) for x in t[:, 3]], "energies": [float(x) for x in e[:, 0]], "setup_wall_s": round(wall - t[:, 0].sum() / 1e3, 2),
            "what": "DFT::scf loop body (src/dft.cpp:100-103): F' = X^T F X, projector, P mixing, rho, LDA, Hartree, F_grid, energies; H/X/P/F resident in HBM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="h2o64")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--whole", action="store_true", help="--impl reference: really run one whole iteration instead of the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scf", action="store_true", help="skip the whole-SCF-iteration block (C++ host)")
    ap.add_argument("--no-peer", action="store_true", help="multi-GPU: sum the matrices with ncclAllReduce instead of the peer-memory kernels")
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
    # two ways to N GPUs: torchrun (one process per GPU, what the driver launches) or plain `python bench.py --gpus N`
    # (ONE process, dftgrid_create_multi — the mode `dftcxx -i ... --gpus N` uses)
    single_process = world == 1 and args.gpus > 1
    ngpus_total = args.gpus if single_process else world
    if args.no_peer:
        os.environ["DFTGRID_DEVELOPER"] = "1"
        os.environ["DFTGRID_NO_PEER"] = "1"
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

    g = MolecularGrid(mol, device=local, rank=rank, nranks=world, ngpus=args.gpus if single_process else 1)
    g.set_grid_parameters(*prm)
    t0 = time.time()
    g.create_grid(comm_id)
    build_wall = time.time() - t0
    peer_path = g.peer_active() if single_process else False
    if world > 1 and not args.no_peer:
        # the matrix sum runs in the library's own peer-memory kernels when the ranks' GPUs have a P2P path
        peer_path = g.connect_peers(dist)  # all ranks agree: peer path only if every rank mapped every buffer
    tb = g.timings()
    P = systems.synthetic_density(mol)
    devices = list(range(args.gpus)) if single_process else [local]
    small = g.npoints * mol.nbf * 8 <= 2 * 126e6 * ngpus_total
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device="cuda:%d" % d) for d in devices] if small else None

    def flush_l2():
        for f in flush:
            f.zero_()
        for d in devices:
            torch.cuda.synchronize(d)

    def timed(step_fn, steps):
        """K steps back to back between device synchronisations, CUDA events on the library's stream(s); small inputs are
        timed step by step with an L2 flush (a 256 MB memset) in between.  Returns total ms on this rank."""
        g.synchronize()
        barrier()
        if small:
            tot = 0.0
            for _ in range(steps):
                flush_l2()
                g.timer_start()
                step_fn()
                tot += g.timer_stop()
        else:
            g.timer_start()
            for _ in range(steps):
                step_fn()
            tot = g.timer_stop()
        g.synchronize()
        barrier()
        return tot

    # ---- value: the step the drop-in host issues every SCF iteration — the fused Fock build (rho, LDA, Hartree potential,
    # F_grid = 2J + XC in one contraction, E_J, E_xc, N_el) — with P resident in HBM, device-timed
    g.upload_density(P)
    for _ in range(args.warmup):
        g.fock_device()
    # the clock sampler forks nvidia-smi (tens of ms from a process this size): start it BEFORE the barrier, otherwise the
    # other ranks' timers include rank 0's fork while they wait for it in the first collective
    sampler = ClockSampler(local) if rank == 0 else None
    g.synchronize()
    n0 = g.launch_count()
    dev_ms = timed(g.fock_device, args.steps)
    launches = g.launch_count() - n0
    phases = g.timings()
    ms_step = max_over_ranks(dev_ms / args.steps)
    # the same with J and XC as two separate matrices (what the reference's own four calls return; round-1 headline)
    for _ in range(2):
        g.iteration_device()
    pair_ms = max_over_ranks(timed(g.iteration_device, args.steps) / args.steps)
    phases_pair = g.timings()

    # ---- e2e: host buffers through the C ABI.  P and F live in pinned host memory (the SCF driver's own buffers);
    # every step copies P host->device and [F_grid | E_J | E_xc | N_el] device->host.
    P_host = torch.from_numpy(np.ascontiguousarray(P)).pin_memory().numpy()
    F_host = torch.empty((mol.nbf, mol.nbf), dtype=torch.float64).pin_memory().numpy()
    for _ in range(2):
        g.fock(P_host, out=F_host)
    g.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if small:
            flush_l2()
        F, ej, exc, nel = g.fock(P_host, out=F_host)
    g.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    clocks = sampler.stop() if sampler else None

    screen_frac = g.screen_fraction()
    # ---- roofline of the dominant kernel (the DMMA contraction of F_grid), live event timings of the last fused step
    npts_loc = g.npoints / ngpus_total  # the mean shard (shards are cut at equal estimated work, not equal point counts)
    flops_contract = 1.0 * npts_loc * mol.nbf * (mol.nbf + 1)  # one symmetric matrix: Npts*nb*(nb+1) flop
    con_ms = max_over_ranks(phases["contract"])
    rho_ms = max_over_ranks(phases["rho"])
    peaks, peak_kind = measured_peaks()
    ach = flops_contract / (con_ms * 1e-3) / 1e12
    roof = {"kernel": "k_contract_tma (+k_contract_reduce), fused F_grid = 2J + XC", "bound": "tensor", "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS,
            "unit": "TFLOP/s", "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": None,
            "peak_source": "FP64 DMMA peak measured on this pool with tools/microbench/fp64_peak.cu (MEASURED_PEAKS.json has no FP64 entry; "
                           "its HBM figure %s GB/s is the denominator for the streaming kernels, %s)" % (peaks.get("hbm_gbs"), peak_kind),
            "ms": con_ms,
            "screen_work_fraction": screen_frac,
            "note": "achieved = ALGORITHMIC flops Npts*nb*(nb+1) / time; the block map of Phi lets the kernel skip insignificant 32-column "
                    "blocks (|phi| <= 1e-20), so it executes about screen_work_fraction of them and frac can exceed what a dense evaluation "
                    "reaches (dense: 10.5 ms = 0.999 at (H2O)64, DMMA pipe 92.7 % busy, profiles/r02_*)",
            "second_kernel": {"kernel": "k_rho_tma", "ms": rho_ms, "achieved_full_2Nnb2": 2.0 * npts_loc * mol.nbf ** 2 / (rho_ms * 1e-3) / 1e12 if rho_ms > 0 else None,
                              "note": "executes half of 2*Npts*nb^2 (P symmetric): frac of peak on executed flops = achieved/2/peak"}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and ngpus_total == 1:  # the capture is of the whole grid on one GPU; a shard's launch reads its own rows only
        try:
            tj = json.load(open(prof)).get(args.workload, {})
            roof["traffic"] = tj.get("k_contract_fock_dram_bytes")
            roof["traffic_algorithmic"] = tj.get("k_contract_fock_algorithmic_bytes")
            roof["traffic_source"] = tj.get("source_r02b") or tj.get("source_r02") or tj.get("source")
        except Exception:
            pass

    # ---- parity against the reference fixture of this workload (every rank takes part in the iterations)
    fx = load_fixture(args.workload, mol.nbf)
    parity = None
    if fx is not None:
        Pfx = fx["P"] if fx["P"] is not None else P
        if fx["P_checksum"] is not None:
            chk = np.array([Pfx.sum(), np.abs(Pfx).sum(), np.trace(Pfx)])
            assert np.allclose(chk, fx["P_checksum"], rtol=1e-13, atol=0), "synthetic P does not match the fixture's"
        parity = parity_block(g, Pfx, fx)
    g.close()

    if rank == 0:
        out = {"metric": METRIC, "value": ms_step, "unit": "ms", "n_gpus": ngpus_total, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": workload_config(args.workload, mol, prm),
               "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": int(mol.nbf ** 2 * 8),
                       "d2h_bytes_per_step": int((mol.nbf ** 2 + 3) * 8), "call": "dftgrid_fock (P in, F_grid + E_J + E_xc + N_el out)"},
               "gpu_launches": int(launches), "roofline": roof, "clocks": clocks,
               "step": "fused Fock build: rho, LDA, Becke/Poisson Hartree potential, F_grid = 2J + XC in one contraction, E_J, E_xc, N_el",
               "pair_ms_per_step": pair_ms, "pair_phases_ms": {k: round(v, 4) for k, v in phases_pair.items()},
               "process_model": "single process, %d devices (dftgrid_create_multi)" % args.gpus if single_process else
                                ("one process per GPU (torchrun)" if world > 1 else "single GPU"),
               "collectives": ("none (single GPU)" if ngpus_total == 1 else
                               "shell sums + rho_lm: ncclAllReduce; matrix sum: " + ("peer-memory kernels (k_contract_reduce_publish + k_peer_sum)"
                                                                                    if peer_path else "ncclAllReduce")),
               "phases_ms": {k: round(v, 4) for k, v in phases.items()},
               "grid_build": {"wall_s": round(build_wall, 3), "becke_ms": tb["becke"], "phi_ms": tb["phi"],
                              "becke_cell_functions_per_s": npts_loc * mol.natoms * (mol.natoms - 1) / (tb["becke"] * 1e-3) * ngpus_total
                              if tb["becke"] > 0 else None,
                              "phi_gridpt_basis_evals_per_s": npts_loc * mol.nbf / (tb["phi"] * 1e-3) * ngpus_total if tb["phi"] > 0 else None,
                              "phi_hbm_write_frac": (npts_loc * mol.nbf * 8 / (tb["phi"] * 1e-3) / 1e9) / peaks.get("hbm_gbs", 6650.0)
                              if tb["phi"] > 0 else None},
               "results": {"exc": exc, "nelec": nel, "e_j": ej, "parity": parity}}
        if (world == 1) and not args.no_scf:
            try:
                out["scf"] = scf_block(args.workload, args.gpus)
            except Exception as e:
                out["scf"] = {"error": str(e)[:200]}
        if world == 1 and not single_process and not args.no_cpu_baseline:
            try:
                ms, desc, det = reference_fit(args.workload, mol, prm, os.cpu_count() or 1)
                out["cpu_baseline"] = dict({"value": ms, "unit": "ms", "cores": os.cpu_count() or 1, "kind": "reference", "sample": desc}, **det)
            except Exception as e:  # the oracle is a checker, never a dependency of the measured path
                out["cpu_baseline"] = {"value": None, "unit": "ms", "cores": os.cpu_count() or 1, "kind": "unavailable", "sample": str(e)[:200]}
        emit(out)
    if dist:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("PARITY FAILURE: %s\n" % json.dumps(parity))
        sys.exit(3)


if __name__ == "__main__":
    main()

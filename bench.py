#!/usr/bin/env python
"""bench.py -- time-to-solution of the multiple-scattering hot path on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one synthetic cluster at one wavelength: solver->update()
(Mie factors, FF source, FF matrix assembly) + solver->solve() (FF GMRES, SH source, SH matrix assembly,
SH GMRES) + the Result cross sections -- what Simulation::scan_wavelengths does per wavelength
(srcAna/Simulation.cpp:643-667).

Workload (BASELINE.json configs[3], the largest single-GPU configuration; C5 needs 8 GPUs for its dense
921.6 GB matrix): 200 Si spheres r = 50 nm on the first 200 sites of the 190 nm cubic lattice of
examples/ManyParticles.xml, nMax = 8, lambda = 800 nm, theta = 45, phi = 90, E_theta = 1, FH + SH, dense
operator, Belos-style GMRES tol 1e-5 / restart 30 / <= 20 restarts.  The same total work is used at every
N (strong scaling); rows of particles are sharded across ranks.  `--workload c5` runs the 1000-sphere
nMax = 10 cluster (8 GPUs).

Operator form: `--operator pairs` (default) streams the compact pair form (unscaled A^T, B^T of the pairs
i < j: 4.08 GB per harmonic on C4), `--operator dense` the reference's full slab (16.4 GB).  Either is far
larger than the 126 MB L2, so consecutive matvecs / steps cannot hit in L2 (config.l2: "inputs larger than
L2").  The roofline numerator is the bytes of the form actually streamed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from optimet_b200 import xmlgen  # noqa: E402


def workload(name):
    belos = [("Solver", "string", "GMRES"), ("Convergence Tolerance", "double", "1.0e-5"),
             ("Maximum Iterations", "int", "600"), ("Num Blocks", "int", "30"), ("Block Size", "int", "1"),
             ("Maximum Restarts", "int", "20")]
    if name == "c4":
        xyz = xmlgen.cube_sites(7, 200, 190.0)
        return dict(name="C4: 200 Si spheres r=50nm, 190nm cubic lattice (first 200 sites), nMax=8, 800nm, FH+SH, dense",
                    xml=xmlgen.cluster_xml(xyz, 50.0, 8, 800.0, belos=belos), nobj=200, nMax=8)
    if name == "c5":
        xyz = xmlgen.random_sites(1000, 2200.0, 150.0, 20261017)
        return dict(name="C5: 1000 Si spheres r=50nm, random in (2200nm)^3 (std::mt19937_64 seed 20261017, min distance 150nm), nMax=10, 800nm, FH+SH, dense",
                    xml=xmlgen.cluster_xml(xyz, 50.0, 10, 800.0, belos=belos), nobj=1000, nMax=10)
    if name == "small":
        xyz = xmlgen.cube_sites(3, 27, 190.0)
        return dict(name="small: 27 Si spheres, nMax=6 (smoke-sized)", xml=xmlgen.cluster_xml(xyz, 50.0, 6, 800.0, belos=belos),
                    nobj=27, nMax=6)
    raise SystemExit("unknown workload " + name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_time(wl, threads, iters_ff, iters_sh, sample_rows=None, repeat=1):
    """The reference's CPU path (oracle port, all host threads) on a bounded sample of the workload,
    scaled to the full workload: assembly per block-row, matvec per byte, SH source per particle."""
    from oracle import oracle as O
    from optimet_b200 import host as H
    O.set_threads(threads)
    case = H.Case(xml=wl["xml"])
    a = case.arrays()
    info = case.info()
    nobj, nMax = info["nobj"], info["nMax"]
    n2 = 2 * nMax * (nMax + 2)
    N = n2 * nobj
    orc = O.Case()
    for j in range(nobj):
        orc.add_sphere(list(a["xyz"][j]), float(a["radius"][j]), nMax, O.MODEL_SILICON, [1.0, 0.0])
    orc.set_source(info["wavelength"], np.deg2rad(45.0), np.deg2rad(90.0), 1.0, 0.0, True)
    rows = sample_rows if sample_rows is not None else max(1, min(threads, nobj))
    t_asm = t_mv = t_src = t_sh = 0.0
    nmv = 3
    for _ in range(repeat):
        t0 = time.perf_counter()
        S = orc.matrix(1, 0, rows)  # rows block-rows x all columns, threads over block-rows
        t_asm += time.perf_counter() - t0
        x = np.ones(N, dtype=np.complex128)
        t0 = time.perf_counter()
        for _ in range(nmv):
            O.matvec(S, x)
        t_mv += (time.perf_counter() - t0) / nmv
        t0 = time.perf_counter()
        orc.source()
        t_src += time.perf_counter() - t0
    t_asm, t_mv, t_src = t_asm / repeat, t_mv / repeat, t_src / repeat
    # a sample of the same blocks on the reference's OWN compiled Coupling (oracle/_ref/libpath_ref.so, built from the
    # reference sources; single thread: its f2c'd AMOS keeps static state and is not thread-safe)
    ref_note = ""
    try:
        from oracle import reference_build as RB
        if RB.have():
            k = complex(orc.info()["waveK"])
            xyz = np.asarray(a["xyz"], dtype=float)
            jobs = [(0, j) for j in range(1, min(nobj, 65))]
            t0 = time.perf_counter()
            for i, j in jobs:
                d = xyz[i] - xyz[j]
                r = float(np.linalg.norm(d))
                RB.coupling([r, float(np.arccos(d[2] / r)), float(np.arctan2(d[1], d[0]))], k, nMax, True)
            per_block = (time.perf_counter() - t0) / max(1, len(jobs))
            ref_note = ("; the reference's own compiled Coupling takes %.1f ms per block on one thread (%d blocks timed; "
                        "the port: %.1f ms per block and thread)"
                        % (per_block * 1e3, len(jobs), t_asm * min(threads, rows) / (rows * (nobj - 1)) * 1e3))
    except Exception:
        pass
    # SH source: per particle cost from the oracle on `rows` particles is not separable through the case API;
    # time the whole SH source once on a reduced cluster of `rows` particles
    orc2 = O.Case()
    for j in range(rows):
        orc2.add_sphere(list(a["xyz"][j]), float(a["radius"][j]), nMax, O.MODEL_SILICON, [1.0, 0.0])
    orc2.set_source(info["wavelength"], np.deg2rad(45.0), np.deg2rad(90.0), 1.0, 0.0, True)
    xi = np.ones(n2 * rows, dtype=np.complex128) * 1e-3
    orc2.sh_source(xi)  # builds the CG tables (once per run in the reference, not counted per step)
    t0 = time.perf_counter()
    orc2.sh_source(xi)
    t_sh = time.perf_counter() - t0
    scale = nobj / float(rows)
    total = (2 * t_asm * scale                      # FF + SH assembly
             + (iters_ff + iters_sh + 2) * t_mv * scale  # matvecs (+1 initial residual each)
             + t_src                                 # FF source (all particles)
             + t_sh * scale)                         # SH source
    sample = ("oracle port, %d threads: FF assembly of %d of %d block-rows (%.2fs), %d matvecs on that %dx%d slab "
              "(%.3fs each), FF source (%.2fs), SH source on %d particles (%.2fs); scaled to the full workload with "
              "the GPU run's GMRES iteration counts (%d FF + %d SH)%s"
              % (threads, rows, nobj, t_asm, nmv, n2 * rows, N, t_mv, t_src, rows, t_sh, iters_ff, iters_sh, ref_note))
    return total, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--matvec-variant", type=int, default=None)
    ap.add_argument("--operator", default="pairs", choices=["pairs", "dense", "aca", "rot"],
                    help="pairs: compact A^T/B^T-of-i<j form (default); dense: the reference's full slab; "
                         "aca: the reference's ACA-compressed operator (eps 1e-3; results differ at that level); "
                         "rot: rotated-axial form (exact; phases + axial A/B + Wigner small-d per pair, csrc/ob_rot.cu)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the extra end-to-end steps in the rotated-axial form")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (ob_set_option)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = workload(args.workload)
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference's own CPU implementation of the path (oracle port; the reference itself cannot be built
        # here: Eigen/Boost/GSL/HDF5 absent), all host threads, bounded sample per step
        iters = (30, 30)
        p = os.path.join(ROOT, "profiles", "last_iters_%s.json" % args.workload)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
                iters = (d["iters_ff"], d["iters_sh"])
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_reference_time(wl, threads, *iters, sample_rows=1)
        vals = []
        sample = ""
        for _ in range(max(1, min(args.steps, 3))):
            v, sample = cpu_reference_time(wl, threads, *iters)
            vals.append(v)
        v = float(np.mean(vals))
        print(json.dumps({
            "impl": "reference", "metric": "time_to_solution_s", "value": v, "unit": "s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "c128 (complex FP64)", "data": "synthetic",
            "config": {"workload": wl["name"], "l2": "inputs larger than L2"},
            "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    from optimet_b200 import capi, host as H

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    case = H.Case(xml=wl["xml"])
    solver = H.Solver(case, device=local_rank)
    if world > 1:
        uid = [capi.Library.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        solver.comm_init(uid[0], rank, world)
    if args.matvec_variant is not None:
        solver.set_option("matvec_variant", args.matvec_variant)
    if args.operator == "aca":
        solver.set_aca_mode(1)
    else:
        solver.set_option("operator", {"pairs": 1, "dense": 0, "rot": 3}[args.operator])
    for opt in args.opt:
        k, v = opt.split("=")
        solver.set_option(k, float(v))
    lib = capi.load()
    ctx = C.c_void_p(H.load().obh_solver_ctx(solver.s))
    opts = case.gmres_defaults()
    info = case.info()
    nobj, nMax = info["nobj"], info["nMax"]
    n2 = 2 * nMax * (nMax + 2)
    N = n2 * nobj
    first, count = capi.Library.partition(nobj, world, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_resident():
        cs = (C.c_double * 5)()
        st = (C.c_int * 2)()
        rc = lib.ob_run(ctx, C.byref(opts), 1, None, None, None, None, cs, st)
        if rc:
            raise RuntimeError(lib.ob_last_error(ctx).decode())
        return list(cs), list(st)

    # ---------------- warm-up ----------------
    res = None
    for _ in range(args.warmup):
        res = solver.step(fetch=True)
    barrier()

    # ---------------- timed: device-resident arm (`value`) ----------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    acc = {}
    ms = C.c_double()
    barrier()
    lib.ob_timer(ctx, 0, None)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cs, st = run_resident()
        tm = solver.ctx_timings()
        for k, v in tm.items():
            acc[k] = v if k == "operator_bytes" else acc.get(k, 0.0) + v
    lib.ob_timer(ctx, 1, C.byref(ms))
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(dev_ms.item()) / args.steps

    # ---------------- timed: end-to-end arm through the host adaptor (`e2e`) ----------------
    barrier()
    lib.ob_timer(ctx, 0, None)
    for _ in range(args.steps):
        res = solver.step(fetch=True)
    lib.ob_timer(ctx, 1, C.byref(ms))
    barrier()
    e2e_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_per_step = float(e2e_ms.item()) / args.steps
    clocks = sampler.stop() if sampler else None

    # ---------------- extra (not the headline): the same end-to-end steps in the rotated-axial operator form ----------------
    # (csrc/ob_rot.cu: exact, 12-15x fewer operator bytes than the pair form; bound by shared-memory wavefronts rather
    # than HBM, so it is reported beside the TMA-streamed pair form the north star specifies, not instead of it)
    alt = None
    if args.operator == "pairs" and not args.no_alt:
        solver.set_option("operator", 3)
        for _ in range(max(1, args.warmup)):
            r2 = solver.step(fetch=True)
        barrier()
        lib.ob_timer(ctx, 0, None)
        for _ in range(args.steps):
            r2 = solver.step(fetch=True)
        lib.ob_timer(ctx, 1, C.byref(ms))
        barrier()
        a_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
        cs2 = torch.tensor([r2[k] for k in ("ext", "sca", "abs", "sca_SH", "abs_SH")], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(a_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cs2, op=dist.ReduceOp.SUM)
        tm2 = solver.ctx_timings()
        alt = {"operator": "rot (phases + axial A/B + Wigner small-d per pair, k_matvec_rot)",
               "e2e_ms_per_step": float(a_ms.item()) / args.steps,
               "matvec_ms_per_apply": tm2["matvec_ms"] / max(1.0, tm2["matvec_count"]),
               "operator_bytes_per_apply": tm2["operator_bytes"], "iters_ff": r2["iters_ff"], "iters_sh": r2["iters_sh"],
               "cross_sections": [float(v) for v in cs2.tolist()]}
        solver.set_option("operator", 1)

    # cross sections are per-rank partial sums over the rank's own particles (linear): add them up
    cs_t = torch.tensor([res[k] for k in ("ext", "sca", "abs", "sca_SH", "abs_SH")], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cs_t, op=dist.ReduceOp.SUM)
    mv_ms = acc["matvec_ms"] / max(1.0, acc["matvec_count"])
    mv_t = torch.tensor([mv_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(mv_t, op=dist.ReduceOp.MAX)

    if rank == 0:
        if alt is not None:
            ref_cs = [float(v) for v in cs_t.tolist()]
            alt["max_rel_diff_of_cross_sections_vs_headline_form"] = max(
                abs(a_ / b_ - 1.0) for a_, b_ in zip(alt["cross_sections"], ref_cs) if b_ != 0.0)
        peak, peak_src = measured_peaks()
        m_loc = n2 * count
        # dense 16 M_loc N + 32 N per apply; pair form 32 n^2 per local pair + 32 N
        # (= 4 N^2 (1 - 1/N_obj) + 32 N on one GPU), reported by the library for the form actually streamed
        mv_bytes = acc["operator_bytes"]
        achieved = mv_bytes / (float(mv_t.item()) * 1e-3) / 1e9
        os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
        try:
            with open(os.path.join(ROOT, "profiles", "last_iters_%s.json" % args.workload), "w") as f:
                json.dump({"iters_ff": st[0], "iters_sh": st[1]}, f)
        except Exception:
            pass
        h2d = (3 + 1) * 8 * nobj + 7 * 16 * nobj + 2 * 16 * (n2 // 2)
        d2h = 4 * 16 * N + 5 * 8
        out = {
            "metric": "time_to_solution_s", "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "c128 (complex FP64)", "data": "synthetic",
            "config": {"workload": wl["name"], "N": N, "rows_per_gpu": m_loc, "gmres": "belos tol=1e-5 restart=30",
                       "iters_ff": st[0], "iters_sh": st[1], "l2": "inputs larger than L2",
                       "operator": args.operator,
                       "phases_ms_per_step": {k: acc[k] / args.steps for k in acc
                                              if k not in ("matvec_count", "launches", "operator_bytes")
                                              and not (k.startswith("trace_") and acc[k] == 0)},
                       "matvecs_per_step": acc["matvec_count"] / args.steps,
                       "cross_sections": dict(zip(["ext", "sca", "abs", "sca_SH", "abs_SH"], [float(x) for x in cs_t.tolist()])),
                       "wall_s_resident_arm": wall},
            "clocks": clocks,
            "e2e": {"value": e2e_per_step / 1e3, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(acc["launches"]),
            "alt_operator": alt,
            "roofline": {"bound": "hbm", "kernel": {"pairs": "k_matvec_pairs (TMA-streamed complex-FP64 pair-form block matvec)",
                                                    "dense": "k_matvec (TMA-streamed complex-FP64 dense block matvec)",
                                                    "aca": "k_matvec_aca (complex-FP64 U(Vx) low-rank + dense near blocks)",
                                                    "rot": "k_matvec_rot (rotation - axial translation - rotation per pair; FP64 latency bound, not HBM)"}[args.operator],
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_launch": mv_bytes, "avg_launch_ms": float(mv_t.item())},
        }
        # assembly (north_star: "achieved FP64 FLOP/s against B200 FP64 peak" + store bandwidth).  Algorithmic flops per
        # VTAC block from SURVEY.md section 8(d) (replay count of the reference recursion); one block per unit
        # processed: a pair in the pair form, an off-diagonal block in the dense form.
        F = {3: 15.8e3, 6: 154.5e3, 8: 427.3e3, 10: 960.5e3, 12: 1883.6e3}.get(nMax)
        asm_ms = 0.5 * (acc["assemble_ff"] + acc["assemble_sh"]) / args.steps
        if args.operator == "pairs":
            units = nobj * (nobj - 1) // 2 // world
            asm_bytes = 32.0 * (n2 // 2) ** 2 * units
        elif args.operator == "rot":
            units = nobj * (nobj - 1) // 2 // world
            asm_bytes = acc["operator_bytes"] - 32.0 * N
        elif args.operator == "aca":
            units = count * (nobj - 1)
            asm_bytes = solver.ctx().aca_stats(1)["stored_bytes"]
            out["aca"] = solver.ctx().aca_stats(1)
        else:
            units = count * (nobj - 1)
            asm_bytes = 16.0 * n2 * n2 * count * nobj
        fp64 = C.c_double()
        lib.ob_measure_fp64_peak(ctx, C.byref(fp64))
        out["assembly"] = {"kernel": {"pairs": "k_assemble_pairs", "dense": "k_assemble", "rot": "k_assemble_axial + k_rot_tables",
                                      "aca": "k_assemble + k_aca_compress + pack"}[args.operator],
                           "ms_per_harmonic": asm_ms, "units_per_launch": units, "stored_bytes_per_launch": asm_bytes,
                           "store_GBps": asm_bytes / (asm_ms * 1e-3) / 1e9,
                           "algorithmic_flops_per_unit": F,
                           "achieved_tflops": (units * F / (asm_ms * 1e-3) / 1e12) if F else None,
                           "fp64_peak_tflops_measured": fp64.value,
                           "frac_of_fp64_peak": (units * F / (asm_ms * 1e-3) / 1e12 / fp64.value) if F else None}
        if world == 1 and not args.no_cpu_baseline:
            v, sample = cpu_reference_time(wl, threads, st[0], st[1])
            out["cpu_baseline"] = {"value": v, "unit": "s", "cores": threads, "kind": "port", "sample": sample}
        tr = os.path.join(ROOT, "profiles", "traffic_matvec.json")
        if os.path.exists(tr) and world == 1:
            try:
                with open(tr) as f:
                    ent = json.load(f).get("%s_%s" % (args.operator, args.workload))
                if ent:
                    out["roofline"]["traffic"] = ent["dram_bytes_per_launch"]
                    out["roofline"]["traffic_source"] = ent["source"]
            except Exception:
                pass
        print(json.dumps(out))
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the ARC radiation hot path: columns/s for one radiation step = RRTMG SW + LW, each producing the
full, clear-sky and clean (aerosol-free) streams, on the configuration BASELINE.json quotes the metric on.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference, all host threads

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
from __future__ import annotations

import argparse
import ctypes as C
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

from wrfchem_arc_interactions_b200 import abi, ktables, radiation as R, synth  # noqa: E402

METRIC = "columns/sec (SW+LW, full+clear+clean)"
WORKLOADS = {
    # name: (ni, nj, nk, synth kwargs)
    "C1": (32, 32, 40, {}),
    "C2": (425, 300, 50, {}),
    "C4": (425, 300, 50, dict(cloudy_frac=1.0, with_re=True)),
    "C5s": (500, 250, 100, {}),      # one-eighth of C5 (1M x 100)
}
SW_STATS = ("swupt", "swuptc", "swuptcln", "swdnt", "swdntc", "swdntcln", "swupb", "swupbc", "swupbcln", "swdnb", "swdnbc", "swdnbcln")
LW_STATS = ("lwupt", "lwuptc", "lwuptcln", "lwdnt", "lwdntc", "lwdntcln", "lwupb", "lwupbc", "lwupbcln", "lwdnb", "lwdnbc", "lwdnbcln")


def flops_per_column(nk, nlay_lw):
    """Algorithmic FLOPs of one column (SURVEY.md 8d / DESIGN.md): SW 112*Ls*(30 + 3*240), LW 140*Ll*(60 + 2*55)."""
    ls = nk + 1
    return 112.0 * ls * (30 + 3 * 240), 140.0 * nlay_lw * (60 + 2 * 55)


def sw_solve_flops_per_column(nk):
    """Share of the SW figure that k_sw_solve executes: taumol 30 + per variant layer prep 40 + reftra 120 +
    combine/direct 25 + the bottom-up half of the adding method 22 = 207; the top-down half (23) and the accumulation over
    g (10) run in k_sw_sweep."""
    return 112.0 * (nk + 1) * (30 + 3 * 207)


KERNEL_CLASSES = ("sw_mcica", "sw_prep", "sw_solve", "sw_sweep", "sw_reduce", "lw_mcica", "lw_prep", "lw_solve", "lw_sweep", "lw_reduce")


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def tile_dims(dom, j0, j1):
    d = dict(dom["dims"])
    d["jts"], d["jte"] = j0, j1
    return d


def run_reference(args, dom, psw, plw):
    """CPU arm: the C++ restatement of the reference's Fortran (no Fortran compiler in this image), all host threads,
    one column per call internally, on a bounded sample (the first rows of the same workload)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    orc = O.oracle_mt(0)
    orc.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
    nthreads = orc.nthreads
    ni, nj = dom["ni"], dom["nj"]
    flags = R.common_flags(dom)
    outs_sw, outs_lw = R.alloc_outputs(dom, "sw"), R.alloc_outputs(dom, "lw")

    def step(rows):
        d = tile_dims(dom, 1, rows)
        t = time.perf_counter()
        orc.RRTMG_LWRAD(d, **R.lw_kwargs(dom, outs_lw, **flags))
        orc.RRTMG_SWRAD(d, **R.sw_kwargs(dom, outs_sw, **flags))
        return time.perf_counter() - t
    # calibrate: rows so that one step is about `target` seconds
    rows = min(nj, max(1, nthreads // 4))
    dt = step(rows)
    target = args.ref_seconds
    rows = int(min(nj, max(1, round(rows * target / max(dt, 1e-3)))))
    for _ in range(args.warmup):
        step(rows)
    times = [step(rows) for _ in range(args.steps)]
    ncol = rows * ni
    total = float(np.sum(times))
    val = ncol * args.steps / total
    return val, total / args.steps * 1e3, {"value": val, "unit": "columns/s", "cores": nthreads, "kind": "port",
                                           "sample": "first %d of %d rows (%d columns) of the workload, SW+LW with clean call, per step" % (rows, nj, ncol)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="CPU arm: target seconds per step")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0)
    ap.add_argument("--clean", type=int, default=1, help="clean_atm_diag (1: full+clear+clean, 0: full+clear only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aer", action="store_true", help="skip the separately reported aerosol-optics stage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    ni, nj, nk, skw = WORKLOADS[args.workload]
    tmp = tempfile.mkdtemp(prefix="arc_bench_")
    psw, plw = ktables.write_files(tmp)

    if args.impl == "reference":
        if rank != 0:
            return
        dom = synth.make_domain(ni, min(nj, 64), nk, seed=synth.SEED, **skw)
        val, ms, cb = run_reference(args, dom, psw, plw)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %dx%d columns x %d levels, MOSAIC-style 4-wavelength aerosol optics in, 40%% cloudy, 25%% night, clean_atm_diag=1; CPU arm runs a row sample of it" % (args.workload, ni, nj, nk)},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C++ restatement of the v3.9.1 Fortran (gfortran absent in this image), -O2 no-FMA, one column per call, j-rows over all host threads"}))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]                      # NCCL prints its version banner to stdout at these levels: one JSON line only
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # every rank owns one tile of the workload size (weak scaling: columns are independent, j-slab partition)
    dom = synth.make_domain(ni, nj, nk, seed=synth.SEED + 1000 * rank, **skw)
    ncol = ni * nj
    nsun = int((dom["xcoszen"] > 0).sum())
    lib = R.lib()
    lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw, device=local_rank)
    L = lib.lib
    L.arc_rad_domain_stats.restype = C.c_int
    L.arc_rad_domain_stats.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_void_p]
    L.arc_rad_measure_fp32_tflops.restype = C.c_float
    L.arc_rad_driver_post.restype = C.c_int
    L.arc_rad_driver_post.argtypes = [C.POINTER(abi.ArcDims), C.c_int] + [abi.c_fp] * 6
    nlay_lw = lib.lw_nlayers()
    flags = R.common_flags(dom, clean_atm_diag=args.clean)
    dims = abi.make_dims(dom["dims"])

    # ---- device-resident arm --------------------------------------------------------------------------------
    ddom = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in dom.items()}
    like = ddom["xcoszen"]
    o_sw, o_lw = R.alloc_outputs(dom, "sw", like=like, ext=False), R.alloc_outputs(dom, "lw", like=like, ext=False)
    rthraten = torch.zeros_like(ddom["t3d"]); swdown = torch.zeros_like(like)
    nst = len(SW_STATS) + len(LW_STATS)
    stats = torch.zeros(nst, 5, dtype=torch.float64, device=dev)
    fptrs = (abi.c_fp * nst)(*[abi.fptr(int(o_sw[n].data_ptr())) for n in SW_STATS], *[abi.fptr(int(o_lw[n].data_ptr())) for n in LW_STATS])
    kw_sw, kw_lw = R.sw_kwargs(ddom, o_sw, **flags), R.lw_kwargs(ddom, o_lw, **flags)
    P = lambda t: abi.fptr(int(t.data_ptr()))

    def step_device():
        lib.RRTMG_LWSW(dims, kw_lw, kw_sw)        # arc_rad_lwsw: LW then SW as one continuous multi-stream pipeline
        lib.check(L.arc_rad_driver_post(C.byref(dims), abi.ARC_MEM_DEVICE, P(o_lw["rthratenlw"]), P(o_sw["rthratensw"]), P(rthraten),
                                        P(o_sw["gsw"]), P(ddom["albedo"]), P(swdown)))
        lib.check(L.arc_rad_domain_stats(C.byref(dims), abi.ARC_MEM_DEVICE, nst, fptrs, C.c_void_p(int(stats.data_ptr()))))
        if dist is not None:
            # domain-mean forcing terms: sums add, extrema combine through max of (-min, max)
            sums = stats[:, :3].contiguous(); ext = torch.stack([-stats[:, 3], stats[:, 4]], 1)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM); dist.all_reduce(ext, op=dist.ReduceOp.MAX)
            return sums, ext
        return stats, None

    stream = torch.cuda.ExternalStream(int(L.arc_rad_stream()), device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    fp32_peak = float(L.arc_rad_measure_fp32_tflops())
    sampler = ClockSampler(local_rank)
    kms = {}
    barrier()
    sampler.start()
    launches0 = int(L.arc_rad_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
        for n in KERNEL_CLASSES:
            kms[n] = kms.get(n, 0.0) + float(L.arc_rad_last_kernel_ms(n.encode()))
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    launches = int(L.arc_rad_launch_count()) - launches0
    clocks = sampler.stop()
    # per-kernel durations with every kernel alone on the GPU (sweep overlap off), outside the timed region: the timed
    # region's per-class times include the slow-down from the other stream's kernels running beside them
    kms_alone = {}
    L.arc_rad_set_overlap.restype = C.c_int
    prev = L.arc_rad_set_overlap(0)
    n_alone = 2
    for _ in range(n_alone):
        step_device()
        for n in KERNEL_CLASSES:
            kms_alone[n] = kms_alone.get(n, 0.0) + float(L.arc_rad_last_kernel_ms(n.encode())) / n_alone
    L.arc_rad_set_overlap(prev)
    t = torch.tensor([max(dev_ms, 0.0), wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    ms_step = dev_ms / args.steps
    value = ncol * world * args.steps / (dev_ms * 1e-3)

    # ---- end-to-end arm: host (pinned) buffers through the same public call, copies inside the timed region ---
    e2e = None
    if not args.no_e2e:
        def pin(a):
            tt = torch.from_numpy(a).pin_memory()
            return tt.numpy(), tt
        keep = []
        hdom = {}
        for k, v in dom.items():
            if isinstance(v, np.ndarray) and v.ndim >= 2:
                a, tt = pin(v); keep.append(tt); hdom[k] = a
            else:
                hdom[k] = v
        h_sw, h_lw = R.alloc_outputs(dom, "sw", ext=False), R.alloc_outputs(dom, "lw", ext=False)
        for o in (h_sw, h_lw):
            for k in list(o):
                a, tt = pin(o[k]); keep.append(tt); o[k] = a
        hk_sw, hk_lw = R.sw_kwargs(hdom, h_sw, **flags), R.lw_kwargs(hdom, h_lw, **flags)
        in_sw = sum(v.nbytes for k, v in hk_sw.items() if isinstance(v, np.ndarray) and k not in h_sw and k in (R.SW_FIELDS_3D + R.SW_FIELDS_2D)
                    and k not in ("rho3d", "dz8w", "qg3d", "gaer300", "gaer999", "waer300", "waer999"))
        # arrays both adapters take (t3d, p3d, qv3d, ...) are uploaded once per slab by the combined call
        in_lw = sum(v.nbytes for k, v in hk_lw.items() if isinstance(v, np.ndarray) and k not in h_lw and k in (R.LW_FIELDS_3D + R.LW_FIELDS_2D)
                    and k not in ("rho3d", "dz8w", "qg3d") and k not in hk_sw)
        out_b = sum(v.nbytes for v in h_sw.values()) + sum(v.nbytes for v in h_lw.values())
        # outputs the call may leave partly unwritten (SW night columns) are uploaded first to keep the caller's values
        inout_b = sum(h_sw[k].nbytes for k in ("rthratensw", "gsw", "swupflx", "swupflxc", "swupflxcln", "swdnflx", "swdnflxc", "swdnflxcln"))
        hfp = (abi.c_fp * nst)(*[abi.fptr(h_sw[n]) for n in SW_STATS], *[abi.fptr(h_lw[n]) for n in LW_STATS])
        hstats = np.zeros((nst, 5), np.float64)

        def step_host():
            lib.RRTMG_LWSW(dims, hk_lw, hk_sw)
            lib.check(L.arc_rad_domain_stats(C.byref(dims), abi.ARC_MEM_HOST, nst, hfp, C.c_void_p(hstats.ctypes.data)))
        for _ in range(2):
            step_host()
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            step_host()
        torch.cuda.synchronize(dev)
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": ncol * world * n_e2e / float(te[0]), "unit": "columns/s",
               "h2d_bytes_per_step": int(in_sw + in_lw + inout_b), "d2h_bytes_per_step": int(out_b),
               "steps": n_e2e, "note": "host pinned WRF-layout arrays through one RRTMG_LWSW step (arc_rad_lwsw: LW then SW) + statistics; the library pipelines j-slabs: upload / compute / download overlap on copy streams, shared inputs uploaded once, LW and SW of a slab chained and slabs not joined (the next slab's LW kernels start under the last SW sweep)"}

    # ---- aerosol optical-property stage (MOSAIC 8-bin sectional), reported separately ------------------
    aer = None
    if not args.no_aer and rank == 0:
        bins, alt, _ = synth.make_aerosol(dom, nbin=8)
        dbins = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in bins]
        dalt = torch.from_numpy(alt).to(dev)
        ao = R.alloc_aer_outputs(dom, like=like)
        for _ in range(2):
            lib.optical_averaging(dims, "sectional", dbins, dalt, ddom["dz8w"], ao)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter(); n_a = 3; kms_a = 0.0
        for _ in range(n_a):
            lib.optical_averaging(dims, "sectional", dbins, dalt, ddom["dz8w"], ao)
            kms_a += float(L.arc_rad_last_kernel_ms(b"aer_optics"))
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / n_a
        aer = {"columns_per_s": ncol / dt, "ms": dt * 1e3, "kernel_ms": kms_a / n_a, "config": "MOSAIC 8-bin sectional, 9 species classes, 4 SW + 16 LW wavelengths, %d levels" % nk,
               "column_aod400_median": float(ao["tauaer400"].sum(dim=1).median()), "parity": "self-consistent only (module_optical_averaging.F is not in the reference repository)"}
        del dbins, ao

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    fsw, flw = flops_per_column(nk, nlay_lw)
    f_solve = sw_solve_flops_per_column(nk)
    sw_solve_ms = kms["sw_solve"] / args.steps
    ach = f_solve * nsun / (sw_solve_ms * 1e-3) / 1e12 if sw_solve_ms > 0 else 0.0
    alone_ms = kms_alone.get("sw_solve", 0.0)
    ls = nk + 1
    nstream = 2 + (1 if args.clean else 0)
    # HBM-bound sweep kernels: algorithmic bytes = the level records read once + the group partials written once
    # (DESIGN.md 3.2): SW 28 B per (sunlit column, g, level, stream), LW 16 B per (column, g, level, stream)
    sw_sweep_bytes = nsun * 112.0 * (ls + 1) * 28.0 * nstream
    lw_sweep_bytes = ncol * 140.0 * (nlay_lw + 1) * 16.0 * (2 if args.clean else 1)
    def gbps(bytes_, ms):
        return bytes_ / 1e9 / (ms * 1e-3) if ms > 0 else None
    roofline = {"kernel": "k_sw_solve", "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak if fp32_peak > 0 else None,
                # DRAM bytes of k_sw_solve: ncu --set full (profiles/r1_summary.md) measured 15.00 GB read+write for a
                # 24,566-sunlit-column launch = 0.611 MB per sunlit column (the level records written for k_sw_sweep)
                "traffic": 0.611e6 * nsun, "traffic_unit": "bytes per step (all k_sw_solve launches)",
                "issue_slot_utilisation_ncu": 0.81,
                "ms_per_step": sw_solve_ms,
                "ms_per_step_alone": alone_ms,
                "frac_alone": (f_solve * nsun / (alone_ms * 1e-3) / 1e12 / fp32_peak) if alone_ms > 0 and fp32_peak > 0 else None,
                "note": "ms_per_step is measured in the timed region, where the sweep kernels of the previous chunk run beside the solver on a second stream; *_alone is the same kernel with the overlap off (outside the timed region)",
                "peak_source": "FP32 FMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 figure; nominal 74.5)",
                "algorithmic_flops_per_column": f_solve,
                "other_kernels": {
                    "k_sw_sweep": {"bound": "hbm", "unit": "GB/s", "peak": 6545.6, "algorithmic_GB_per_step": sw_sweep_bytes / 1e9,
                                   "achieved": gbps(sw_sweep_bytes, kms["sw_sweep"] / args.steps),
                                   "achieved_alone": gbps(sw_sweep_bytes, kms_alone.get("sw_sweep", 0.0))},
                    "k_lw_sweep": {"bound": "hbm", "unit": "GB/s", "peak": 6545.6, "algorithmic_GB_per_step": lw_sweep_bytes / 1e9,
                                   "achieved": gbps(lw_sweep_bytes, kms["lw_sweep"] / args.steps),
                                   "achieved_alone": gbps(lw_sweep_bytes, kms_alone.get("lw_sweep", 0.0))},
                    "k_lw_solve": {"bound": "fp32", "unit": "TFLOP/s", "peak": fp32_peak,
                                   "algorithmic_flops_per_column": 140.0 * nlay_lw * (60 + 2 * 42),
                                   "achieved": 140.0 * nlay_lw * (60 + 2 * 42) * ncol / (kms["lw_solve"] / args.steps * 1e-3) / 1e12 if kms["lw_solve"] > 0 else None,
                                   "achieved_alone": 140.0 * nlay_lw * (60 + 2 * 42) * ncol / (kms_alone["lw_solve"] * 1e-3) / 1e12 if kms_alone.get("lw_solve", 0) > 0 else None}}}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        a2 = argparse.Namespace(**vars(args)); a2.steps, a2.warmup, a2.ref_seconds = 1, 0, args.cpu_baseline_seconds
        sdom = synth.make_domain(ni, min(nj, 64), nk, seed=synth.SEED, **skw)
        _, _, cpu_baseline = run_reference(a2, sdom, psw, plw)
    out = {
        "metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %dx%d columns x %d levels per GPU (SW %d / LW %d layers), 4-wavelength aerosol optics in, 40%% cloudy, 25%% night columns, clean_atm_diag=%d, all six flux profiles out" % (args.workload, ni, nj, nk, nk + 1, nlay_lw, args.clean),
                   "l2": "inputs (%.0f MB) and workspaces exceed the 126 MB L2" % (sum(v.nbytes for v in dom.values() if isinstance(v, np.ndarray)) / 1e6),
                   "columns_per_gpu": ncol, "sunlit_columns": nsun, "partition": "j-slabs, one tile per rank; NCCL all-reduce of 24x5 domain statistics per step" if world > 1 else "single tile"},
        "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
        "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
        "kernel_ms_per_step_alone": kms_alone,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "aer_optics": aer,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the ARC radiation hot path: columns/s for one radiation step = RRTMG LW + SW, each producing the full,
clear-sky and clean (aerosol-free) streams, plus the domain statistics the ARC decomposition consumes, on the configuration
BASELINE.json quotes the metric on.

    python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo); N > 1 under torchrun
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference, all host threads

One JSON line on stdout (rank 0).  DESIGN.md section 6 says what every key means.

Multi-GPU: ONE domain is cut into j-slabs by partition.jslabs (balanced on LW + sunlit SW cost), one slab per rank, no
collective on the data path; per step the 24x5 domain sums are all-reduced and the 24 TOA / surface fields are all-gathered
(NCCL), after which every rank holds the global fields for the order statistics and Moran's I.
  --scaling weak   (default) the domain grows with N: ni x (nj * N) rows of the workload's grid, about one workload per rank
  --scaling strong the domain is fixed (e.g. --workload C3: 2000 x 2000 x 60, BASELINE config 3) and divided by N
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

from wrfchem_arc_interactions_b200 import abi, ktables, partition, radiation as R, synth  # noqa: E402

METRIC = "columns/sec (SW+LW, full+clear+clean)"
WORKLOADS = {
    # name: (ni, nj, nk, synth kwargs)
    "C1": (32, 32, 40, {}),
    "C2": (425, 300, 50, {}),
    "C3": (2000, 2000, 60, {}),       # 4 M columns: strong scaling over 2/4/8 GPUs
    "C4": (425, 300, 50, dict(cloudy_frac=1.0, with_re=True)),
    "C5s": (500, 250, 100, {}),      # one-eighth of C5 (1M x 100)
}
SW_STATS = ("swupt", "swuptc", "swuptcln", "swdnt", "swdntc", "swdntcln", "swupb", "swupbc", "swupbcln", "swdnb", "swdnbc", "swdnbcln")
LW_STATS = ("lwupt", "lwuptc", "lwuptcln", "lwdnt", "lwdntc", "lwdntcln", "lwupb", "lwupbc", "lwupbcln", "lwdnb", "lwdnbc", "lwdnbcln")
KERNEL_CLASSES = ("sw_mcica", "sw_prep", "sw_solve", "sw_sweep", "sw_reduce", "lw_mcica", "lw_prep", "lw_solve", "lw_reduce")
PERC = np.array([50.0, 25.0, 75.0, 5.0, 95.0], np.float32)       # calc_standard_stats: median, quartiles, 5th / 95th


def flop_model(nk, nlay_lw):
    """Algorithmic FLOPs per column: SW 112*Ls*(30 + 3*240) of which k_sw_solve executes
    taumol 30 + per variant (layer prep 40 + reftra 120 + combine/direct 25 + bottom-up adding 22 = 207); LW
    140*Ll*(60 + 2*55), all of it in k_lw_band."""
    ls = nk + 1
    return {"sw_total": 112.0 * ls * (30 + 3 * 240), "sw_solve": 112.0 * ls * (30 + 3 * 207), "lw_band": 140.0 * nlay_lw * (60 + 2 * 55)}


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


def global_coszen(ni, nj, night_frac=0.25, seed=synth.SEED + 5):
    """cos(zenith) of the whole domain, same distribution as synth.make_domain: the partition is decided on it."""
    rng = np.random.default_rng(seed)
    cz = rng.uniform(0.05, 1.0, (nj, ni))
    night = rng.uniform(0, 1, (nj, ni)) < night_frac
    return np.where(night, -rng.uniform(0.0, 0.5, (nj, ni)), cz).astype(np.float32)


def run_reference(args, dom, psw, plw, reps):
    """CPU arm: the C++ restatement of the reference's Fortran (no Fortran compiler in this image), all host threads,
    one column per call internally, on a bounded sample (the first rows of the same workload); median over `reps` steps."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    orc = O.oracle_mt(0)
    orc.init(dom["p_top"], dom["dims"]["kme"], psw, plw)
    nthreads = orc.nthreads
    ni, nj = dom["ni"], dom["nj"]
    flags = R.common_flags(dom, clean_atm_diag=args.clean)
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
    rows = int(min(nj, max(1, round(rows * args.ref_seconds / max(dt, 1e-3)))))
    for _ in range(args.warmup):
        step(rows)
    times = [step(rows) for _ in range(max(reps, 1))]
    ncol = rows * ni
    med = float(np.median(times))
    val = ncol / med
    return val, med * 1e3, {"value": val, "unit": "columns/s", "cores": nthreads, "kind": "port", "repetitions": len(times),
                            "spread": [ncol / max(times), ncol / min(times)],
                            "sample": "first %d of %d rows (%d columns) of the workload, SW+LW with clean_atm_diag=%d, median of %d steps" % (
                                rows, nj, ncol, args.clean, len(times)),
                            "build": "oracle/Makefile: g++ -O3 -ffp-contract=off -fno-fast-math (stand-in for gfortran -O3 on baseline x86-64: no FMA contraction)"}


def ncu_metrics():
    """Per-kernel DRAM traffic / issue utilisation from the tracked ncu summary of this round (profiles/r2_ncu_metrics.json,
    written by tools/profile_summary.py from one `ncu --set full` capture; keyed by the commit it was taken at)."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_metrics.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="CPU arm: target seconds per step")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=5.0)
    ap.add_argument("--clean", type=int, default=1, help="clean_atm_diag (1: full+clear+clean, 0: full+clear only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aer", action="store_true", help="skip the separately reported aerosol-optics stage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C5 clean-vs-full cost ratio and the C4 (modal, all-cloudy) leg")
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
        val, ms, cb = run_reference(args, dom, psw, plw, reps=max(args.steps, 3))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": "columns/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %dx%d columns x %d levels, MOSAIC-style 4-wavelength aerosol optics in, 40%% cloudy, 25%% night, clean_atm_diag=%d; CPU arm runs a row sample of it" % (args.workload, ni, nj, nk, args.clean)},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "columns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "C++ restatement of the v3.9.1 Fortran (gfortran absent in this image), one column per call, j-rows over all host threads; value = median step"}))
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

    # ---- one domain, cut into j-slabs ------------------------------------------------------------------------------
    nj_glob = nj * world if args.scaling == "weak" else nj
    cz_glob = global_coszen(ni, nj_glob, night_frac=0.25)
    slabs = partition.jslabs(cz_glob, world)                   # [(jts, jte)] 1-based inclusive, balanced on LW + sunlit SW cost
    j0, j1 = slabs[rank]
    rows = j1 - j0 + 1
    dom = synth.make_domain(ni, rows, nk, seed=synth.SEED + 1000 * rank, **skw)
    dom["xcoszen"] = np.ascontiguousarray(cz_glob[j0 - 1:j1])
    ncol = ni * rows
    ncol_glob = ni * nj_glob
    nsun = int((dom["xcoszen"] > 0).sum())
    lib = R.lib()
    lib.init(dom["p_top"], dom["dims"]["kme"], psw, plw, device=local_rank)
    L = lib.lib
    L.arc_rad_domain_stats.restype = C.c_int
    L.arc_rad_domain_stats.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_void_p]
    L.arc_rad_percentiles.restype = C.c_int
    L.arc_rad_percentiles.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_int, abi.c_fp, C.c_void_p]
    L.arc_rad_morans_i.restype = C.c_int
    L.arc_rad_morans_i.argtypes = [C.POINTER(abi.ArcDims), C.c_int, C.c_int, C.POINTER(abi.c_fp), C.c_void_p]
    L.arc_rad_measure_fp32_tflops.restype = C.c_float
    L.arc_rad_driver_post.restype = C.c_int
    L.arc_rad_driver_post.argtypes = [C.POINTER(abi.ArcDims), C.c_int] + [abi.c_fp] * 6
    nlay_lw = lib.lw_nlayers()
    flags = R.common_flags(dom, clean_atm_diag=args.clean)
    dims = abi.make_dims(dom["dims"])
    gdims = abi.make_dims(dict(dom["dims"], jms=1, jme=nj_glob, jts=1, jte=nj_glob))        # the gathered 2-D fields

    # ---- device-resident arm --------------------------------------------------------------------------------
    ddom = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in dom.items()}
    like = ddom["xcoszen"]
    o_sw, o_lw = R.alloc_outputs(dom, "sw", like=like, ext=False), R.alloc_outputs(dom, "lw", like=like, ext=False)
    rthraten = torch.zeros_like(ddom["t3d"]); swdown = torch.zeros_like(like)
    nst = len(SW_STATS) + len(LW_STATS)
    stats = torch.zeros(nst, 5, dtype=torch.float64, device=dev)
    fields2d = [o_sw[n] for n in SW_STATS] + [o_lw[n] for n in LW_STATS]
    fptrs = (abi.c_fp * nst)(*[abi.fptr(int(t.data_ptr())) for t in fields2d])
    kw_sw, kw_lw = R.sw_kwargs(ddom, o_sw, **flags), R.lw_kwargs(ddom, o_lw, **flags)
    P = lambda t: abi.fptr(int(t.data_ptr()))
    stream = torch.cuda.ExternalStream(int(L.arc_rad_stream()), device=dev)
    # gather of the 24 fields of every slab into global fields on every rank (partition.SlabGather: one all-gather per step)
    maxrows = max(b - a + 1 for a, b in slabs)
    if dist is not None:
        gather = partition.SlabGather(dist, slabs, rank, nst, ni, dev)
        glob = gather.glob
        gptrs = (abi.c_fp * nst)(*[abi.fptr(int(glob[f].data_ptr())) for f in range(nst)])
    else:
        gather, glob, gptrs = None, None, fptrs
    pct = torch.zeros(nst, 5, dtype=torch.float32, device=dev)
    mor = torch.zeros(nst, dtype=torch.float32, device=dev)

    def step_device():
        lib.RRTMG_LWSW(dims, kw_lw, kw_sw)        # arc_rad_lwsw: LW then SW as one continuous multi-stream pipeline
        lib.check(L.arc_rad_driver_post(C.byref(dims), abi.ARC_MEM_DEVICE, P(o_lw["rthratenlw"]), P(o_sw["rthratensw"]), P(rthraten),
                                        P(o_sw["gsw"]), P(ddom["albedo"]), P(swdown)))
        lib.check(L.arc_rad_domain_stats(C.byref(dims), abi.ARC_MEM_DEVICE, nst, fptrs, C.c_void_p(int(stats.data_ptr()))))
        sums, ext = stats, None
        if dist is not None:
            with torch.cuda.stream(stream):       # the collectives run on the library's stream: the CUDA events below see them
                # domain-mean forcing terms: sums add, extrema combine through max of (-min, max)
                sums = stats[:, :3].contiguous(); ext = torch.stack([-stats[:, 3], stats[:, 4]], 1)
                dist.all_reduce(sums, op=dist.ReduceOp.SUM); dist.all_reduce(ext, op=dist.ReduceOp.MAX)
                gather(fields2d)                   # the 24 TOA / surface fields of every slab -> global fields on every rank
            stream.synchronize()
        # order statistics and Moran's I of calc_standard_stats on the global fields (they do not combine from partial results)
        lib.check(L.arc_rad_percentiles(C.byref(gdims), abi.ARC_MEM_DEVICE, nst, gptrs, 5, abi.fptr(PERC), C.c_void_p(int(pct.data_ptr()))))
        lib.check(L.arc_rad_morans_i(C.byref(gdims), abi.ARC_MEM_DEVICE, nst, gptrs, C.c_void_p(int(mor.data_ptr()))))
        return sums, ext

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
            kms[n] = kms.get(n, 0.0) + max(float(L.arc_rad_last_kernel_ms(n.encode())), 0.0)
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    launches = int(L.arc_rad_launch_count()) - launches0
    clocks = sampler.stop()
    # per-kernel durations with every kernel alone on the GPU (stream overlap off), outside the timed region: the timed
    # region's per-class times include the slow-down from the other streams' kernels running beside them
    kms_alone = {}
    L.arc_rad_set_overlap.restype = C.c_int
    prev = L.arc_rad_set_overlap(0)
    n_alone = 2
    for _ in range(n_alone):
        step_device()
        for n in KERNEL_CLASSES:
            kms_alone[n] = kms_alone.get(n, 0.0) + max(float(L.arc_rad_last_kernel_ms(n.encode())), 0.0) / n_alone
    L.arc_rad_set_overlap(prev)
    t = torch.tensor([max(dev_ms, 0.0), wall_ms], dtype=torch.float64, device=dev)
    bal = torch.tensor([float(ncol), float(nsun)], dtype=torch.float64, device=dev)
    balance = [bal.clone() for _ in range(world)]
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(balance, bal)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    ms_step = dev_ms / args.steps
    value = ncol_glob * args.steps / (dev_ms * 1e-3)

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
        hfields = [h_sw[n] for n in SW_STATS] + [h_lw[n] for n in LW_STATS]
        hfp = (abi.c_fp * nst)(*[abi.fptr(a) for a in hfields])
        hstats = np.zeros((nst, 5), np.float64)
        hpct = np.zeros((nst, 5), np.float32)
        if dist is not None:
            hsend = torch.zeros(nst, maxrows, ni, dtype=torch.float32).pin_memory()

        def step_host():
            lib.RRTMG_LWSW(dims, hk_lw, hk_sw)
            lib.check(L.arc_rad_domain_stats(C.byref(dims), abi.ARC_MEM_HOST, nst, hfp, C.c_void_p(hstats.ctypes.data)))
            if dist is not None:
                # the 24 2-D fields (host) -> device -> all-gather -> global statistics on the device
                for f, a in enumerate(hfields):
                    hsend[f, :rows] = torch.from_numpy(a)
                with torch.cuda.stream(stream):
                    gather.send.copy_(hsend, non_blocking=True)
                    hs = torch.from_numpy(hstats).to(dev)
                    sums = hs[:, :3].contiguous(); ext = torch.stack([-hs[:, 3], hs[:, 4]], 1)
                    dist.all_reduce(sums, op=dist.ReduceOp.SUM); dist.all_reduce(ext, op=dist.ReduceOp.MAX)
                    gather.exchange()
                stream.synchronize()
                lib.check(L.arc_rad_percentiles(C.byref(gdims), abi.ARC_MEM_DEVICE, nst, gptrs, 5, abi.fptr(PERC), C.c_void_p(int(pct.data_ptr()))))
                lib.check(L.arc_rad_morans_i(C.byref(gdims), abi.ARC_MEM_DEVICE, nst, gptrs, C.c_void_p(int(mor.data_ptr()))))
                return float(pct.sum().cpu())          # device -> host read of the step's result
            lib.check(L.arc_rad_percentiles(C.byref(dims), abi.ARC_MEM_HOST, nst, hfp, 5, abi.fptr(PERC), C.c_void_p(hpct.ctypes.data)))
            return float(hpct.sum())
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
        e2e_s = float(te[0]) / n_e2e
        e2e = {"value": ncol_glob / e2e_s, "unit": "columns/s",
               "h2d_bytes_per_step": int(in_sw + in_lw + inout_b), "d2h_bytes_per_step": int(out_b),
               "ms_per_step": e2e_s * 1e3, "steps": n_e2e, "bytes_are": "per rank",
               "aggregate_h2d_GBps": (in_sw + in_lw + inout_b) * world / e2e_s / 1e9,
               "note": "host pinned WRF-layout arrays through one RRTMG_LWSW step (arc_rad_lwsw: LW then SW) + statistics; the library pipelines j-slabs: upload / compute / download overlap on copy streams, shared inputs uploaded once, LW and SW of a slab chained and slabs not joined; aggregate_h2d_GBps = all ranks' uploads over the step time (the copies are hidden under compute, so this is demand, not the PCIe ceiling)"}

    # ---- aerosol optical-property stage, reported separately; sectional 8-bin, or modal for C4 ------------
    def aer_leg(adom, dd, modal, chain_step=None):
        bins, alt, sg = synth.make_aerosol(adom, nbin=8, modal=modal) if modal else synth.make_aerosol(adom, nbin=8)
        dbins = [{k: torch.from_numpy(v).to(dev) for k, v in b.items()} for b in bins]
        dalt = torch.from_numpy(alt).to(dev)
        ao = R.alloc_aer_outputs(adom, like=dd["xcoszen"])
        mode = "modal" if modal else "sectional"
        for _ in range(2):
            lib.optical_averaging(abi.make_dims(adom["dims"]), mode, dbins, dalt, dd["dz8w"], ao, sigmag=sg if modal else None)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter(); n_a = 3; kms_a = 0.0
        for _ in range(n_a):
            lib.optical_averaging(abi.make_dims(adom["dims"]), mode, dbins, dalt, dd["dz8w"], ao, sigmag=sg if modal else None)
            kms_a += float(L.arc_rad_last_kernel_ms(b"aer_optics"))
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / n_a
        nc = adom["ni"] * adom["nj"]
        nbin = len(bins)
        # flop model of the Chebyshev-Mie evaluation: per (level, section, wavelength) the refractive-index mixing,
        # bilinear weights and polynomial recurrence (~100) + 3 quantities x 4 corner tables x 50 coefficients x 2 (1200)
        flops = float(nc) * adom["nk"] * 8 * 20 * 1300.0
        chained = None
        if chain_step is not None:
            # second headline (north_star's target sentence includes the optics): the optics stage writes tau / omega / g of the
            # 4 SW wavelengths and the 16 LW band optical depths straight into the device arrays RRTMG_SWRAD / RRTMG_LWRAD read;
            # one step = optics + the whole radiation step above, timed on the device
            co = {k: dd[k] for k in ao}
            def both():
                lib.optical_averaging(abi.make_dims(adom["dims"]), mode, dbins, dalt, dd["dz8w"], co, sigmag=sg if modal else None)
                chain_step()
            both()
            torch.cuda.synchronize(dev)
            ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_c = 3
            ca.record(stream)
            for _ in range(n_c):
                both()
            cb.record(stream)
            torch.cuda.synchronize(dev)
            cms = ca.elapsed_time(cb) / n_c
            chained = {"ms_per_step": cms, "columns_per_s": nc / (cms * 1e-3),
                       "note": "aerosol optics (8 sections x 20 wavelengths) -> RRTMG SW+LW with full / clear / clean streams -> statistics, device-resident; the optics outputs are the radiation inputs (no copy between the stages)"}
        return {"columns_per_s": nc / dt, "ms": dt * 1e3, "kernel_ms": kms_a / n_a, "rrtmg_plus_optics": chained,
                "config": ("MADE/SORGAM 3 modes -> 8 sections" if modal else "MOSAIC 8-bin sectional") + ", 9 species classes, 4 SW + 16 LW wavelengths, %d levels, %d inputs bins" % (adom["nk"], nbin),
                "column_aod400_median": float(ao["tauaer400"].sum(dim=1).median()),
                "roofline": {"kernel": "k_aer_mie (+ k_aer_prep)", "bound": "fp32", "unit": "TFLOP/s", "peak": fp32_peak,
                             "algorithmic_flops_per_column": flops / nc, "achieved": flops / (kms_a / n_a * 1e-3) / 1e12,
                             "frac": flops / (kms_a / n_a * 1e-3) / 1e12 / fp32_peak if fp32_peak > 0 else None},
                "parity": "self-consistent only (module_optical_averaging.F is not in the reference repository)"}

    aer = None
    if not args.no_aer and rank == 0:
        aer = aer_leg(dom, ddom, modal=(args.workload == "C4"), chain_step=step_device if world == 1 else None)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- extras at N = 1: C5 clean-vs-full cost ratio and config 4 (modal optics + all-cloudy + re_*) ---------
    extras = None
    if not args.no_extras and world == 1:
        del ddom, o_sw, o_lw, rthraten, fields2d
        torch.cuda.empty_cache()
        extras = {}

        def small_step(name, clean, nsteps=3):
            n_i, n_j, n_k, kw = WORKLOADS[name]
            xd = synth.make_domain(n_i, n_j, n_k, seed=synth.SEED + 3, **kw)
            lib.init(xd["p_top"], xd["dims"]["kme"], psw, plw, device=local_rank)
            xstream = torch.cuda.ExternalStream(int(L.arc_rad_stream()), device=dev)       # re-initialisation recreates the library's streams
            xdd = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) and v.ndim >= 2 else v) for k, v in xd.items()}
            xs, xl = R.alloc_outputs(xd, "sw", like=xdd["xcoszen"], ext=False), R.alloc_outputs(xd, "lw", like=xdd["xcoszen"], ext=False)
            res = {}
            for cl in clean:
                fl = R.common_flags(xd, clean_atm_diag=cl)
                a_, b_ = R.lw_kwargs(xdd, xl, **fl), R.sw_kwargs(xdd, xs, **fl)
                xdm = abi.make_dims(xd["dims"])
                for _ in range(2):
                    lib.RRTMG_LWSW(xdm, a_, b_)
                torch.cuda.synchronize(dev)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record(xstream)
                for _ in range(nsteps):
                    lib.RRTMG_LWSW(xdm, a_, b_)
                eb.record(xstream)
                torch.cuda.synchronize(dev)
                res[cl] = ea.elapsed_time(eb) / nsteps
            return xd, xdd, res
        xd, xdd, r5 = small_step("C5s", (1, 0))
        extras["c5_clean_vs_full"] = {"workload": "C5s: 500x250 = 125,000 columns x 100 levels (one-eighth of config 5), SW 101 / LW 113 layers",
                                      "ms_full_clear_clean": r5[1], "ms_full_clear": r5[0], "ratio": r5[1] / r5[0],
                                      "reference_structure": "<= 2.0 SW (spcvmc_sw incl. taumol_sw repeats) and ~1.5 LW (rtrnmc repeats), SURVEY 8d"}
        del xd, xdd
        torch.cuda.empty_cache()
        if args.workload != "C4":
            xd, xdd, r4 = small_step("C4", (1,))
            c4 = {"workload": "C4: 425x300 columns x 50 levels, every column cloudy, re_cloud / re_ice / re_snow given (inflg 5, iceflg 5)",
                  "ms_per_step": r4[1], "columns_per_s": 127500 / (r4[1] * 1e-3)}
            if not args.no_aer:
                c4["aer_optics_modal"] = aer_leg(xd, xdd, modal=True)
                c4["columns_per_s_with_optics"] = 127500 / ((r4[1] + c4["aer_optics_modal"]["kernel_ms"]) * 1e-3)
            extras["c4_modal_all_cloudy"] = c4
            del xd, xdd
            torch.cuda.empty_cache()

    fm = flop_model(nk, nlay_lw)
    met = ncu_metrics() or {}

    def fp32_entry(kernel, cls, flops_col, ncols):
        ms_in, ms_al = kms.get(cls, 0.0) / args.steps, kms_alone.get(cls, 0.0)
        ach = flops_col * ncols / (ms_in * 1e-3) / 1e12 if ms_in > 0 else None
        acha = flops_col * ncols / (ms_al * 1e-3) / 1e12 if ms_al > 0 else None
        m = met.get(kernel, {})
        return {"kernel": kernel, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": ach / fp32_peak if ach and fp32_peak > 0 else None,
                "traffic": m.get("dram_bytes_per_column") * ncols if m.get("dram_bytes_per_column") is not None else None,
                "traffic_unit": "DRAM bytes per step (ncu dram__bytes_read + write per column of the tracked capture x columns)",
                "issue_slot_utilisation_ncu": m.get("issue_active_pct"), "ncu_source": met.get("_source"),
                "ms_per_step": ms_in, "ms_per_step_alone": ms_al, "achieved_alone": acha,
                "frac_alone": acha / fp32_peak if acha and fp32_peak > 0 else None, "algorithmic_flops_per_column": flops_col}
    nstream = 2 + (1 if args.clean else 0)
    # HBM-bound SW sweep: algorithmic bytes = the level records read once: 28 B per (sunlit column, g, level, stream)
    sw_sweep_bytes = nsun * 112.0 * (nk + 2) * 28.0 * nstream
    hbm_peak = 6545.6
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def gbps(bytes_, ms):
        return bytes_ / 1e9 / (ms * 1e-3) if ms > 0 else None
    r_sw = fp32_entry("k_sw_solve", "sw_solve", fm["sw_solve"], nsun)
    r_lw = fp32_entry("k_lw_band", "lw_solve", fm["lw_band"], ncol)
    r_sweep = {"kernel": "k_sw_sweep", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "algorithmic_GB_per_step": sw_sweep_bytes / 1e9,
               "achieved": gbps(sw_sweep_bytes, kms.get("sw_sweep", 0.0) / args.steps), "achieved_alone": gbps(sw_sweep_bytes, kms_alone.get("sw_sweep", 0.0)),
               "ms_per_step": kms.get("sw_sweep", 0.0) / args.steps, "ms_per_step_alone": kms_alone.get("sw_sweep", 0.0)}
    if r_sweep["achieved"]:
        r_sweep["frac"] = r_sweep["achieved"] / hbm_peak
    dominant = r_sw if (r_sw["ms_per_step"] >= r_lw["ms_per_step"]) else r_lw
    others = [x for x in (r_sw, r_lw) if x is not dominant] + [r_sweep]
    if aer is not None:
        others.append(aer["roofline"])
    roofline = dict(dominant)
    roofline["peak_source"] = "FP32 FMA microbenchmark run in this process (MEASURED_PEAKS.json has no FP32 figure; nominal 74.5)"
    roofline["note"] = ("ms_per_step is measured in the timed region, where kernels of the other streams run beside this one; *_alone is the "
                        "same kernel with the stream overlap off (outside the timed region)")
    roofline["other_kernels"] = others
    roofline["whole_step"] = {"algorithmic_flops": fm["sw_total"] * nsun + fm["lw_band"] * ncol,
                              "achieved": (fm["sw_total"] * nsun + fm["lw_band"] * ncol) / (ms_step * 1e-3) / 1e12, "unit": "TFLOP/s"}
    roofline["whole_step"]["frac"] = roofline["whole_step"]["achieved"] / fp32_peak if fp32_peak > 0 else None
    cpu_baseline = None
    if not args.no_cpu_baseline:
        a2 = argparse.Namespace(**vars(args)); a2.warmup, a2.ref_seconds = 0, args.cpu_baseline_seconds
        sdom = synth.make_domain(ni, min(nj, 64), nk, seed=synth.SEED, **skw)
        _, _, cpu_baseline = run_reference(a2, sdom, psw, plw, reps=3)
    out = {
        "metric": METRIC, "value": value, "unit": "columns/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %dx%d columns x %d levels%s (SW %d / LW %d layers), 4-wavelength aerosol optics in, %s, 25%% night columns, clean_atm_diag=%d, all six flux profiles out, 13 domain statistics of 24 TOA/surface fields" % (
                       args.workload, ni, nj_glob, nk, (" = %d x the workload's rows" % world) if (world > 1 and args.scaling == "weak") else "", nk + 1, nlay_lw,
                       "every column cloudy" if skw.get("cloudy_frac") == 1.0 else "40% cloudy", args.clean),
                   "l2": "inputs (%.0f MB per rank) and workspaces exceed the 126 MB L2" % (sum(v.nbytes for v in dom.values() if isinstance(v, np.ndarray)) / 1e6),
                   "columns_total": ncol_glob, "columns_per_rank": [int(b[0]) for b in balance], "sunlit_per_rank": [int(b[1]) for b in balance],
                   "partition": ("partition.jslabs: %d j-slabs of one %dx%d domain, rows per rank %s (balanced on ni + 1.6 x sunlit per row); no data-path collective; per step NCCL all-reduce of 24x5 sums and all-gather of the 24 2-D fields (%.1f MB), then order statistics and Moran's I on the gathered fields" % (
                                     world, ni, nj_glob, [b - a + 1 for a, b in slabs], 4.0 * nst * maxrows * ni * world / 1e6)) if world > 1 else "single tile"},
        "clocks": clocks, "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
        "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
        "kernel_ms_per_step_alone": kms_alone,
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "aer_optics": aer, "extras": extras,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

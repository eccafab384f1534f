#!/usr/bin/env python
"""bench.py -- headline benchmark of the MM-PIHM hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--size 1M|100k|...] [--fbr]

A "step" is one 60 s model step of the synthetic watershed (SolveCVode +
Summary, src/pihm.c:53-57) during the rain pulse of the synthetic storm.
Metric (BASELINE.json): simulated days per wall-second at 1M triangles, with
RHS evals/s and the achieved HBM GB/s of the RHS kernels beside it.

Prints ONE JSON line (rank 0).  --impl reference times the reference's own CPU
implementation (oracle/_ref, OpenMP on all host cores) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import mm_pihm_b200  # noqa: E402,F401
from mm_pihm_b200 import watershed as W  # noqa: E402

T0 = 2 * 3600.0          # model time of step 0: one hour into the 6 h rain pulse
STEP = 60.0
B_RHS = {False: 376.0, True: 476.0}     # algorithmic bytes / element / RHS
# dram__bytes_read.sum + dram__bytes_write.sum of one RHS (k_pre + k_main) from the committed
# ncu --set full capture (profiles/); None where no capture exists
RHS_TRAFFIC_BYTES = {("1M", False): 384.7e6, ("1M", True): 437.4e6}
# FP64 side figure: 2 x DFMA + DADD + DMUL thread instructions of k_pre + k_main from the same kind
# of capture (profiles/r02_rhs_1M_ncu_summary.md): 179 + 581 MFLOP per RHS at 1M triangles (pihm), 179 + 1069 (fbr)
RHS_FP64_FLOP = {("1M", False): 0.760e9, ("1M", True): 1.248e9}
FP64_LANES_PER_SM = 64      # sm__sass_thread_inst_executed_op_dfma_pred_on peak_sustained per SM and cycle


def fp64_side(size, fbr, rhs_ms, clk):
    """achieved FP64 rate of one RHS evaluation next to the HBM figure (the kernels are bound by
    dependent FP64 chains, not by memory): counted flops / CUDA-event time; peak = SMs x 64 lanes x 2 x SM clock"""
    flop = RHS_FP64_FLOP.get((size, fbr))
    if flop is None or not rhs_ms:
        return None
    import torch
    nsm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or 1965.0
    peak = nsm * FP64_LANES_PER_SM * 2 * mhz * 1e6 / 1e12
    ach = flop / (rhs_ms * 1e-3) / 1e12
    return {"flop_per_rhs": flop, "achieved_tflops": ach, "peak_tflops": peak, "frac": ach / peak,
            "source": "profiles/r02_rhs_1M_ncu_summary.md (2 x DFMA + DADD + DMUL thread instructions)"}


def forcing_at(tb, k):
    return W.storm_forcing(tb, T0 + k * STEP)


def span_settings(args):
    """--span day: one simulated day from the RelaxIc state at t = 0 -- the dry hour before
    the storm, the 6 h rain pulse, 17 h of recession -- with the first simulated hour as warm-up: 60 + 1380 model
    steps.  The rain column is the only forcing that changes; it is re-uploaded every LSM step (15 model steps)
    inside the timed region.  Side legs that would repeat the day (strong scaling, CPU baseline) are off."""
    global T0
    if args.span == "day":
        T0 = 0.0
        args.warmup, args.steps = 60, 1380
        args.no_strong = args.no_cpu = True


def workload_config(size, fbr, ne_glob, nr_glob, world, span="storm"):
    """the `config` object -- identical in both arms (the driver compares them)"""
    where = ("in the rain pulse (t0 = 2 h)" if span == "storm" else
             "of one simulated day (t0 = 0: dry hour, 6 h rain pulse, recession), first simulated hour = warm-up")
    txt = (f"pihm{'-fbr' if fbr else ''} synthetic {size}-triangle watershed ({ne_glob} elements, {nr_glob} river "
           f"segments), 60 s model steps (SolveCVode + Summary/MassBalance) {where}, "
           f"reltol 1e-3 abstol 1e-4")
    if world > 1:
        txt += (f"; weak scaling: 1M triangles per GPU, mesh partitioned over {world} GPUs; "
                f"value = sim-days/s x (triangles / 1M)")
    return {"workload": txt, "nelem": ne_glob, "nriver": nr_glob,
            "nsv": (5 if fbr else 3) * ne_glob + 2 * nr_glob, "parallelism": f"mesh-partition x{world}",
            "l2": "RHS working set 376 B x nelem and 21 state-sized vectors exceed the 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, device=0):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        if not self.rows:       # timed region shorter than the sampling period: one direct query
            try:
                q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap")
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20)
                self.rows = [l.strip() for l in out.stdout.splitlines() if l.strip()]
            except Exception:
                pass
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def e2e_leg(args, torch, model, y, reset, step, forc_host, barrier, max_over_ranks, stream, ev0, ev1, tb):
    """the same K steps through the C ABI with HOST buffers: every step the three forcing columns come from pinned
    host memory and the state goes back to it (the drop-in driver's per-step traffic).  --e2e sync: the copies sit on
    the compute stream between two steps (pihm_b200_set_forcing_col, pihm_b200_vec_download); --e2e pipelined: the
    next step's columns travel while this step computes and the state of this step while the next one does
    (pihm_b200_forcing_prefetch / _commit, pihm_b200_vec_download_async; the last pull is awaited inside the timed
    region)."""
    K, Wu = args.steps, args.warmup
    pipelined = args.e2e == "pipelined"
    reset()
    host_y = [torch.empty(model.nsv, dtype=torch.float64).pin_memory().numpy() for _ in range(2)]
    forc_tabs = {}
    keys = sorted(forc_host) if forc_host else range(0, Wu + K + 15, 15)
    base = None
    for k in keys:
        if forc_host:
            tab = forc_host[k]
        else:               # day span: one table, the rain column rewritten per LSM step
            base = forcing_at(tb, 0)[:W.F_ETT + 1] if base is None else base
            tab = base.copy(); tab[W.F_PCPDRP] = W.storm_rain(T0 + k * STEP) * W.storm_modulation(tb)
        t = torch.from_numpy(np.ascontiguousarray(tab)).pin_memory()
        forc_tabs[k] = t.numpy()
        forc_tabs[("keep", k)] = t
    cols = (W.F_PCPDRP, W.F_EDIR, W.F_ETT)

    def columns(k):
        tab = forc_tabs[(k // 15) * 15]
        return [tab[c] for c in cols]

    def run(k):
        if pipelined:
            step(k, e2e="pipelined", host_forc=columns(k + 1), host_y=host_y[k & 1])
        else:
            step(k, e2e="sync", host_forc=forc_tabs[(k // 15) * 15], host_y=host_y[0])
    if pipelined:
        model.forcing_prefetch(cols, columns(0))
    for k in range(Wu):
        run(k)
    if pipelined:
        model.transfer_wait()
    barrier()
    ev0.record(stream)
    for k in range(Wu, Wu + K):
        run(k)
    if pipelined:
        model.transfer_wait()           # the last state has arrived in host memory
    ev1.record(stream)
    barrier()
    return max_over_ranks(ev0.elapsed_time(ev1))


# --------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mm_pihm_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    fbr = bool(args.fbr)
    size = args.size
    if world > 1:
        # weak scaling: 1M triangles per GPU (BASELINE config[4]: 8M triangles on 8 GPUs)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        size = {2: "2M", 4: "4M", 8: "8M"}.get(world)
        if size is None:
            raise SystemExit("--gpus must be 1, 2, 4 or 8")
    tb = W.make_named(size, fbr=fbr)
    ne_glob, nr_glob = tb["nelem"], tb["nriver"]
    if world > 1:
        from mm_pihm_b200 import partition as PT
        part = PT.partition(tb, world, parts=[rank])[0]
        y0 = tb["y0"][part["state_idx"]]
        tb = part                                    # local tables (owned + ghosts) from here on
        model = lib.Model(tb, device=local)
        uid = [lib.Model.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model.comm_init(rank, world, uid[0])
    else:
        y0 = tb["y0"]
        model = lib.Model(tb, device=local, reorder=args.reorder)
    ne, nr = tb["nelem"], tb["nriver"]
    stream = torch.cuda.current_stream()
    model.set_stream(stream.cuda_stream)
    cv = lib.Cvode(model)
    y = model.N_VNew(y0)
    tb["y0"] = y0
    K, Wu = args.steps, args.warmup
    model.set_diagnostics(True)     # Summary() + MassBalance() on the device after every step (update.c:3-160)
    paths = model.comm_paths() if world > 1 else {"halo": None}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reset():
        y.upload(tb["y0"])
        model.set_stale_ovlflow(np.zeros((3, ne)))
        model.set_forcing(forcing_at(tb, 0), np.zeros(nr))
        model.set_ws0(y)                # after the forcing table: its ws0.surf column is a placeholder
        cv.SetCVodeParam(y)

    # forcing tables of every LSM step (15 model steps) the run crosses, generated before any
    # timed region: the synthetic generator is numpy on the host and not part of the hot path
    day = args.span == "day"
    forc_host = {} if day else {k: forcing_at(tb, k) for k in range(0, Wu + K + 15, 15)}
    rain_mod = W.storm_modulation(tb) if day else None      # day span: only the rain column changes
    rivbc0 = np.zeros(nr)

    def step(k, e2e=False, host_forc=None, host_y=None):
        if e2e == "pipelined":
            model.forcing_commit()      # this step's columns, prefetched under the previous step's kernels
            model.forcing_prefetch((W.F_PCPDRP, W.F_EDIR, W.F_ETT), host_forc)      # the NEXT step's
        elif e2e:
            # the drop-in driver's per-step traffic: forcing columns in, state out
            for c in (W.F_PCPDRP, W.F_EDIR, W.F_ETT):
                model.set_forcing_col(c, host_forc[c])
        elif day:
            if k % 15 == 0:
                model.set_forcing_col(W.F_PCPDRP, W.storm_rain(T0 + k * STEP) * rain_mod)
        elif k % 15 == 0:
            model.set_forcing(forc_host[k], rivbc0)
            model.Summary(y)            # the table's ws0.surf column is a placeholder: ws0.surf = y[SURF] again
        cv.SolveCVode((k + 1) * STEP, y)
        model.SummaryMB(y, STEP)        # like the reference's step (pihm.c:53-57): mass balance, ws0 <- y
        if e2e == "pipelined":
            model.download_async(y, host_y)
        elif e2e:
            model.L.pihm_b200_vec_download(y.h, host_y.ctypes.data)

    # ---- device-resident timing ------------------------------------------------
    reset()
    for k in range(Wu):
        step(k)
    st0 = cv.stats(); l0 = model.launches
    clocks = ClockSampler(local); clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(Wu, Wu + K):
        step(k)
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clk = clocks.stop()
    st1 = cv.stats(); l1 = model.launches
    rhs_evals = (st1["nfe"] + st1["nfeLS"]) - (st0["nfe"] + st0["nfeLS"])
    nst = st1["nst"] - st0["nst"]
    # whole-job throughput: simulated days per second, scaled by the mesh size in units of
    # 1M triangles (== sim-days/s at 1M triangles for N = 1; N GPUs carry N x 1M triangles)
    mtri = ne_glob / 1.0e6 if world > 1 else 1.0
    value = (K * STEP / 86400.0) / (ms * 1e-3) * mtri

    # ---- end to end through the C ABI with host buffers --------------------------
    ms_e2e = None
    if args.e2e == "auto":
        args.e2e = "pipelined" if world == 1 else "sync"
    if not args.no_e2e:
        ms_e2e = e2e_leg(args, torch, model, y, reset, step, forc_host, barrier, max_over_ranks, stream, ev0, ev1, tb)
    e2e_value = (K * STEP / 86400.0) / (ms_e2e * 1e-3) * mtri if ms_e2e else None

    # ---- RHS kernels: live CUDA-event duration over back-to-back launches ----------
    yv = y
    yd = model.N_VNew()
    nrep = 30
    for _ in range(3):
        model.ode_dev(0.0, yv, yd)
    barrier()
    ev0.record(stream)
    for _ in range(nrep):
        model.ode_dev(0.0, yv, yd)
    ev1.record(stream)
    barrier()
    rhs_ms = max_over_ranks(ev0.elapsed_time(ev1)) / nrep

    # ---- the same kernels in situ: events around every RHS evaluation the integrator issues, and the
    # host time spent at its synchronisation points (a few steps; not part of the timed regions above)
    reset()
    for k in range(2):
        step(k)
    cv.profile(True)
    nprof = min(K, 10)
    for k in range(2, 2 + nprof):
        step(k)
    pf = cv.get_profile()
    cv.profile(False)
    rhs_ms_in_situ = max_over_ranks(pf["rhs_ms"] / max(pf["rhs_evals"], 1))
    host_wait_ms = max_over_ranks(pf["host_wait_ms"] / nprof)
    # ---- N_Vector / integrator kernels in situ: a pair of CUDA events around every vector kernel of a few
    # more model steps (programmatic serialization off meanwhile), true bytes read + written per launch
    cv.profile(2)
    for k in range(2 + nprof, 2 + nprof + min(K, 5)):
        step(k)
    kp = cv.get_kernel_profile()
    cv.profile(False)
    peak, peak_src = measured_peak()
    nown_elem, nsv_local = model.nown_elem, model.nsv
    achieved = B_RHS[fbr] * nown_elem / (rhs_ms * 1e-3) / 1e9      # per GPU
    if world > 1:
        tot = torch.tensor([float(rhs_evals), float(l1 - l0)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        rhs_evals = int(tot[0].item())
    cv.close(); model.close()
    del cv, model, y, yd
    # ---- strong scaling side figure: the 8M-triangle mesh of BASELINE config[4] on these N GPUs
    strong = None
    if not args.no_strong and not fbr:
        if world == 8 and size == "8M":
            strong = {"mesh": "8M", "ms_per_step": ms / K, "ms_per_rhs_eval": ms / max(rhs_evals, 1),
                      "rhs_evals_per_step": rhs_evals / K, "rhs_ms": rhs_ms, "steps": K, "same_as": "main run"}
        else:
            strong = strong_scaling_leg(args, rank, world, local, stream, barrier, max_over_ranks, steps=K, warmup=Wu)
    if rank != 0:
        dist.destroy_process_group()
        return

    out = {
        "metric": "simulated days/wall-sec at 1M triangles; RHS evals/s and achieved HBM GB/s",
        "value": value, "unit": "sim-days/s", "n_gpus": world, "steps": K, "warmup": Wu,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(size, fbr, ne_glob, nr_glob, world, args.span),
        "rhs_evals_per_s": rhs_evals / (ms * 1e-3), "rhs_evals": rhs_evals, "cvode_steps": nst,
        "ms_per_rhs_eval": ms / max(rhs_evals, 1),      # whole step / evaluations: RHS + vector kernels + host
        "rhs_ms": rhs_ms,                               # k_pre + k_main, back to back (30 calls, same input)
        "rhs_ms_in_situ": rhs_ms_in_situ,               # the same pair inside the integrator (event to event)
        "host_wait_ms_per_step": host_wait_ms,          # host time at the integrator's synchronisation points
        "halo_path": paths["halo"], "allreduce_path": paths["halo"],     # p2p: inside our kernels over NVLink peer memory
        "strong_scaling_8M": strong,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": RHS_TRAFFIC_BYTES.get((size, fbr)),
                     "kernel": "k_pre + k_main (one RHS evaluation, per GPU)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": B_RHS[fbr] * nown_elem,
                     "frac_in_situ": B_RHS[fbr] * nown_elem / (rhs_ms_in_situ * 1e-3) / 1e9 / peak,
                     "traffic_source": "profiles/r02_rhs_1M_ncu_summary.md: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of k_pre + k_main"},
        "fp64": fp64_side(size, fbr, rhs_ms, clk),
        # per fused vector kernel of the integrator, in situ on rank 0: algorithmic bytes (vector passes x 8 N)
        # / event-to-event time / the same measured peak; short kernels (2-3 passes of 24 MB) carry their launch
        # latency and reduction tail in the denominator
        "vector_roofline": {
            "peak": peak, "unit": "GB/s",
            "kernels": {nm: {"launches": v["launches"], "us": 1e3 * v["ms"] / max(v["launches"], 1),
                             "gbs": v["bytes"] / max(v["ms"], 1e-9) / 1e6,
                             "frac": v["bytes"] / max(v["ms"], 1e-9) / 1e6 / peak} for nm, v in kp.items()},
            "all": {"gbs": sum(v["bytes"] for v in kp.values()) / max(sum(v["ms"] for v in kp.values()), 1e-9) / 1e6,
                    "frac": sum(v["bytes"] for v in kp.values()) / max(sum(v["ms"] for v in kp.values()), 1e-9) / 1e6 / peak},
        },
        "e2e": ({"value": e2e_value, "unit": "sim-days/s", "ms_per_step": ms_e2e / K,
                 "h2d_bytes_per_step": 3 * 8 * ne, "d2h_bytes_per_step": 8 * nsv_local,
                 "transfers": args.e2e,
                 "note": ("pipelined: the per-step copies overlap the step's kernels on the library's copy stream; the "
                          "device-resident leg (`value`) re-uploads the whole 11-column forcing table synchronously "
                          "at every LSM step (k % 15 == 0: 88 MB), which this leg replaces by its per-step columns"
                          if args.e2e == "pipelined" else
                          "sync: the per-step copies sit on the compute stream between two steps")} if ms_e2e else None),
        "gpu_launches": int(l1 - l0),
        "clocks": clk,
    }
    if not args.no_cpu and world == 1:
        out["cpu_baseline"] = cpu_reference(tb, fbr, steps=args.cpu_steps, warmup=Wu)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def strong_scaling_leg(args, rank, world, local, stream, barrier, max_over_ranks, steps=6, warmup=3):
    """BASELINE config[4] / north_star '>= 6x on 8 B200 for an 8M-triangle mesh': the SAME 8M mesh on the N GPUs of
    this run (N = 1: unpartitioned), the same model steps of the same storm as the main run (so that the line of
    N = 1 and the main line of N = 8 divide into the strong-scaling factor); ms per step and per RHS evaluation."""
    import torch
    import torch.distributed as dist
    from mm_pihm_b200 import lib, partition as PT
    tb = W.make_named("8M")
    nr = tb["nriver"]
    if world > 1:
        part = PT.partition(tb, world, parts=[rank])[0]
        y0 = tb["y0"][part["state_idx"]]
        tb = part
        model = lib.Model(tb, device=local)
        uid = [lib.Model.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        model.comm_init(rank, world, uid[0])
    else:
        y0 = tb["y0"]
        model = lib.Model(tb, device=local, reorder=args.reorder)
    ne, nr = tb["nelem"], tb["nriver"]
    model.set_stream(stream.cuda_stream)
    model.set_diagnostics(True)
    cv = lib.Cvode(model)
    y = model.N_VNew(y0)
    model.set_stale_ovlflow(np.zeros((3, ne)))
    model.set_forcing(forcing_at(tb, 0), np.zeros(nr))
    model.set_ws0(y)
    cv.SetCVodeParam(y)

    forc_host = {k: forcing_at(tb, k) for k in range(0, warmup + steps + 15, 15)}

    def step(k):
        if k % 15 == 0 and k > 0:
            model.set_forcing(forc_host[k], np.zeros(nr))
            model.Summary(y)
        cv.SolveCVode((k + 1) * STEP, y)
        model.SummaryMB(y, STEP)
    for k in range(warmup):
        step(k)
    st0 = cv.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(warmup, warmup + steps):
        step(k)
    ev1.record(stream)
    barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    st1 = cv.stats()
    evals = (st1["nfe"] + st1["nfeLS"]) - (st0["nfe"] + st0["nfeLS"])
    yd = model.N_VNew()
    for _ in range(3):
        model.ode_dev(0.0, y, yd)
    barrier()
    ev0.record(stream)
    for _ in range(20):
        model.ode_dev(0.0, y, yd)
    ev1.record(stream)
    barrier()
    rhs_ms = max_over_ranks(ev0.elapsed_time(ev1)) / 20
    cv.close(); model.close()
    return {"mesh": "8M", "ms_per_step": ms / steps, "ms_per_rhs_eval": ms / max(evals, 1),
            "rhs_evals_per_step": evals / steps, "rhs_ms": rhs_ms, "steps": steps, "warmup": warmup}


# --------------------------------------------------------------------------- reference / CPU baseline
def cpu_reference(tb, fbr, steps, warmup):
    """The reference's own CPU implementation (oracle/_ref: unmodified MM-PIHM +
    CVODE compiled from /root/reference) on the same workload, all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reflib
    ne, nr = tb["nelem"], tb["nriver"]
    cores = os.cpu_count() or 1
    cvomp = (not fbr) and ne >= 30000            # README.md:65-67: OpenMP N_Vector above ~30k elements
    ref = reflib.RefModel(fbr=fbr, cvode_omp=cvomp, threads=cores).create_from_tables(tb)
    ref.init_state(tb["y0"])
    ref.set_ovlflow(np.zeros((3, ne)))
    ref.set_cvode_param()
    f = None
    t_start = None
    for k in range(warmup + steps):
        if k == warmup:
            s0 = ref.stats()
            t_start = time.perf_counter()
        if k % 15 == 0 or f is None:
            f = forcing_at(tb, k)
        f[W.F_WS0SURF] = ref.get_ws()[:ne]
        ref.set_forcing(f, np.zeros(nr))
        ref.model_step(k)
    dt = time.perf_counter() - t_start
    s1 = ref.stats()
    rhs = (s1["nfe"] + s1["nfeLS"]) - (s0["nfe"] + s0["nfeLS"])
    return {"value": (steps * STEP / 86400.0) / dt, "unit": "sim-days/s", "cores": ref.threads,
            "kind": "reference",
            "sample": f"{steps} model steps (after {warmup} warm-up) of the same watershed and forcing; "
                      f"reference ODE() with OpenMP, {'OpenMP' if cvomp else 'serial'} N_Vector",
            "seconds": dt, "rhs_evals_per_s": rhs / dt, "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fbr = bool(args.fbr)
    size = {1: args.size, 2: "2M", 4: "4M", 8: "8M"}.get(args.gpus, args.size)
    args.size = size
    tb = W.make_named(size, fbr=fbr)
    steps = min(args.steps, args.ref_max_steps)
    cb = cpu_reference(tb, fbr, steps=steps, warmup=args.warmup)
    # same scaling of the metric as our arm: sim-days/s x (triangles / 1M) when the job carries N x 1M triangles
    scale = tb["nelem"] / 1.0e6 if args.gpus > 1 else 1.0
    out = {
        "impl": "reference",
        "metric": "simulated days/wall-sec at 1M triangles; RHS evals/s and achieved HBM GB/s",
        "value": cb["value"] * scale, "unit": "sim-days/s",
        "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(size, fbr, tb["nelem"], tb["nriver"], args.gpus, args.span),
        "rhs_evals_per_s": cb["rhs_evals_per_s"],
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"] * scale, "unit": "sim-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="1M", choices=sorted(W.SIZES))
    ap.add_argument("--fbr", action="store_true")
    ap.add_argument("--reorder", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the 8M strong-scaling side figure")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--ref-max-steps", type=int, default=30)
    ap.add_argument("--span", default="storm", choices=["storm", "day"],
                    help="storm: --steps model steps inside the rain pulse (default); day: one simulated day, see span_settings")
    ap.add_argument("--e2e", default="auto", choices=["auto", "sync", "pipelined"],
                    help="host-buffer leg: copies on the compute stream between two steps (sync), or overlapped with the "
                         "steps on the library's copy stream (pipelined, csrc/transfer.cu); auto = pipelined on one GPU, "
                         "sync on a partitioned run (the pipelined calls have only been run on unpartitioned contexts)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (side runs only: a line without e2e is not a bench line)")
    args = ap.parse_args()
    span_settings(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

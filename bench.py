#!/usr/bin/env python
"""bench.py — grid points/s of the fused 2-moment (SB2006) tendency kernel, Float64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cumicro|reference]
    torchrun ... bench.py --gpus N ...        (N > 1: one rank per GPU, weak scaling)

A "step" is one pass of the hot path (BMT.bulk_microphysics_tendencies, 2-moment warm
rain) over one batch of 2^24 synthetic grid points per GPU (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every field.
"""
from __future__ import annotations

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

KEYS = ("rho", "T", "q_tot", "q_lcl", "n_lcl", "q_rai", "n_rai")
OUTS = ("dq_lcl_dt", "dn_lcl_dt", "dq_rai_dt", "dn_rai_dt")
METRIC = "grid points/sec, 2M bulk tendencies FP64"
UNIT = "grid points/s"
BYTES_PER_POINT = 88            # 7 input + 4 live output Float64 columns
WORKLOAD = "2-moment Seifert-Beheng 2006 full tendency set Float64 over 2^24 grid points per GPU"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region (NVML, in-process
    thread, ~2 ms period; falls back to nvidia-smi -lms when NVML is unavailable)."""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.t, self.max = index, [], False, None, None

    def _loop_nvml(self):
        import pynvml as nv
        h = self.h
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates all GPUs of the box; map through CUDA_VISIBLE_DEVICES when set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    idx = self.index
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.nv = nv
            self.t = threading.Thread(target=self._loop_nvml, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": [], "samples": 0}
        nv = self.nv
        names = {}
        for nm, attr in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                         ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                         ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                         ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
            if hasattr(nv, attr):
                names[nm] = getattr(nv, attr)
        reasons = set()
        for _, rs in self.rows:
            for nm, bit in names.items():
                if rs & bit:
                    reasons.add(nm)
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.max),
                "reasons": sorted(reasons), "samples": len(self.rows)}


def reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  Julia is
    not installable here, so this is the CPU restatement (oracle/, OpenMP over points,
    all host threads) — kind 'port'.  Each step = one pass over a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cumicro
    from cumicro import CMP
    from cumicro.testing import synthetic_states_2m
    from oracle import oracle as orc
    n = args.ref_points
    st = synthetic_states_2m(n, seed=1234)
    block = CMP.pack_2m_warm(CMP.Microphysics2MParams(np.float64), CMP.ThermodynamicsParameters(np.float64))
    cols = [st[k] for k in KEYS]
    orc.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    threads = orc.num_threads()
    for _ in range(args.warmup):
        orc.bmt2m_warm(block, *cols)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.bmt2m_warm(block, *cols)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_points_per_step": n},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n} of the workload's points per step, OpenMP over points, {threads} threads "
                                       "(C++ restatement of the Julia scalar methods; Julia itself is not in the image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cumicro", choices=["cumicro", "reference"])
    ap.add_argument("--points", type=int, default=1 << 24, help="grid points per GPU")
    ap.add_argument("--ref-points", type=int, default=1 << 22, help="points per step of the CPU arm")
    ap.add_argument("--cpu-sample", type=int, default=1 << 22, help="points of the cpu_baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--diagnostics", action="store_true",
                    help="also all-reduce (NCCL) a 4-double diagnostic vector every step, as the multi-GPU host model would")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cumicro" else args.warmup

    if args.impl == "reference":
        return reference_arm(args)

    # stdout carries exactly ONE JSON line: libraries that print there on their own (NCCL's "NCCL version ..." banner at
    # communicator creation) are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import cumicro
    from cumicro import BMT, CMP
    from cumicro.testing import synthetic_states_2m

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the cumicro arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lib = cumicro._abi.load()
    n = args.points
    mp = CMP.Microphysics2MParams(np.float64)
    tps = CMP.ThermodynamicsParameters(np.float64)
    scheme = BMT.Microphysics2Moment()
    # every rank owns a different slab of the global grid (weak scaling: fixed points per GPU)
    st = synthetic_states_2m(n, seed=1234 + rank)
    pinned = {k: torch.from_numpy(st[k]).pin_memory() for k in KEYS}
    cols = {k: pinned[k].to(dev, non_blocking=True) for k in KEYS}
    outs = [torch.empty_like(cols["rho"]) for _ in range(4)]
    torch.cuda.synchronize()

    diag = torch.zeros(4, dtype=torch.float64, device=dev)

    def step():
        r = BMT.bulk_microphysics_tendencies(scheme, mp, tps, *[cols[k] for k in KEYS], out=outs)
        if args.diagnostics and dist is not None:
            dist.all_reduce(diag)     # the only collective of the path: optional global diagnostic sums
        return r

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- timed region: K steps, CUDA events on the launching (current) stream -------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = lib.cumicro_launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step()
            ev[i + 1].record()
        barrier()
    launches = lib.cumicro_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = n * world * args.steps / (total_ms_max * 1e-3)

    # ---- FP64 pipe peak measured in place --------------------------------------------------
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    flops = C.c_double(0)
    fp64_peak = None
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.cumicro_probe_fp64_fma(C.c_int64(1 << 15), 8, C.c_void_p(scratch.data_ptr()), C.byref(flops),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        if rc == 0:
            fp64_peak = max(fp64_peak or 0.0, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    # ---- e2e: host buffers in, host buffers out, through the C-ABI host entry point --------
    host_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
    BMT.bulk_microphysics_tendencies_host(scheme, mp, tps, *[pinned[k] for k in KEYS], out=host_out)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        BMT.bulk_microphysics_tendencies_host(scheme, mp, tps, *[pinned[k] for k in KEYS], out=host_out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n * world * args.e2e_steps / float(t.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = _peaks()
    kernel_ms = float(np.mean(per_launch_ms))
    achieved = BYTES_PER_POINT * n / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": n, "global_points": n * world,
                   "parallelism": f"column slabs x{world}, no data-path collective"
                                  + (" + NCCL all-reduce of 4 diagnostic doubles per step" if args.diagnostics and world > 1 else ""),
                   "l2": "inputs (7 x 134 MB columns) larger than the 126 MB L2; no flush needed",
                   "psd": "SB2006 limited rain PSD, log-uniform number concentrations"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": None, "peak_source": peak_src, "kernel": "pointwise_kernel_pipelined<double,7,4,Warm2MFused<7,SPEC=1>,128x7> (cp.async double-buffered inputs, 16 waves; specialised for the default SB2006 block structure)",
                     "kernel_ms": kernel_ms, "bytes_per_point": BYTES_PER_POINT},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 7 * 8 * n, "d2h_bytes_per_step": 4 * 8 * n,
                "steps": args.e2e_steps, "api": "cumicro_bmt2m_warm_host_f64 (pinned host buffers, chunked H2D/kernel/D2H)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if fp64_peak:
        line["fp64_probe_tflops"] = fp64_peak
    prof = os.path.join(ROOT, "profiles", "roofline_inputs.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            fl = pj.get("fp64_flops_per_point_2m_warm")
            if fl and fp64_peak:
                tf = fl * value / world / 1e12
                line["roofline_fp64"] = {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                                         "frac": tf / fp64_peak, "flops_per_point": fl, "source": pj.get("source")}
                ai = pj.get("fp64_arith_inst_per_point_2m_warm")
                if ai:   # DFMA, DMUL and DADD all take one FP64 issue slot: the pipe's own roofline is instructions, not flops
                    line["roofline_fp64"]["issue_frac"] = ai * value / world / (fp64_peak * 1e12 / 2.0)
                    line["roofline_fp64"]["fp64_inst_per_point"] = ai
            if pj.get("dram_bytes_per_launch_2m_warm"):
                line["roofline"]["traffic"] = pj["dram_bytes_per_launch_2m_warm"]
        except Exception:
            pass

    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle as orc
        orc.set_num_threads(os.cpu_count() or 1)
        m = min(args.cpu_sample, n)
        block = CMP.pack_2m_warm(mp, tps)
        sample = [st[k][:m] for k in KEYS]
        orc.bmt2m_warm(block, *[c[: m // 8] for c in sample])
        t0 = time.perf_counter()
        reps = 0
        while True:
            orc.bmt2m_warm(block, *sample)
            reps += 1
            if time.perf_counter() - t0 > 10.0 or reps >= 50:
                break
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m * reps / dt, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                                "sample": f"first {m} points of the workload x {reps} passes, OpenMP over points "
                                          "(C++ restatement of the reference's scalar Julia methods)"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

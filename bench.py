#!/usr/bin/env python
"""bench.py — grid points/s of the fused 2-moment (SB2006) tendency kernel, Float64.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cumicro|reference]
    torchrun ... bench.py --gpus N ...        (N > 1: one rank per GPU, weak scaling)

A "step" is one pass of the hot path (BMT.bulk_microphysics_tendencies, 2-moment warm
rain) over one batch of 2^24 synthetic grid points per GPU (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).  See DESIGN.md §5 for every field.

Legs of the cumicro arm (all on the device's current stream, CUDA-event timed):
  value       EXACTLY K steps after W warm-up steps, inputs resident in HBM (burst figure)
  sustained   the same step for >= 1 s, SM clock sampled throughout (what a long model run sees)
  e2e         the same metric through the C-ABI host entry point: pinned host buffers in, pinned host buffers out,
              H2D + kernel + D2H inside the timed region; beside it the measured pinned H2D / D2H link rates
  config5     (N > 1, or --config5) BASELINE config 5: the fused 1M + 2M + ice-nucleation kernel on this rank's column slab with
              the in-kernel domain diagnostics and their NCCL all-reduce (the only collective of the path, through the C-ABI
              cumicro_nccl_allreduce_f64) inside the timed region; sub-record `p2p`: the same step with the exchange done by
              peer-memory stores over NVLink inside the fused call's own finish kernel (cumicro_fused_1m2m_icenuc_p2p_f64)
  cpu_baseline (N = 1) the CPU restatement of the reference on the box's host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
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
KERNEL = ("warm2m_tile_kernel<double,7,LIM=1,128x6,ALL_OUT,TAB> (cm_tile2m.cuh: block-uniform tiles, bulk-copy inputs, "
          "fast SB2006 body cm_sb2006_fast.cuh with the per-parameter-block ventilation table)")


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _roofline_inputs():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "roofline_inputs.json")))
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle-reason samples DURING a timed region (NVML, in-process thread, ~2 ms period)."""

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
    """--impl reference: the reference's own CPU implementation of the path.  Julia is not installable here (no network, no
    depot), so this is the CPU restatement (oracle/, OpenMP over points, all host threads) — kind 'port'.  Each step = one
    pass over the SAME workload as the cumicro arm (2^24 points, same generator and seed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cumicro  # noqa: F401
    from cumicro import CMP
    from cumicro.testing import synthetic_states_2m
    from oracle import oracle as orc
    n = args.points
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
            "config": {"workload": WORKLOAD, "points_per_gpu": n, "global_points": n,
                       "psd": "SB2006 limited rain PSD, log-uniform number concentrations"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"the whole workload ({n} points) per step, OpenMP over points, {threads} threads "
                                       "(C++ restatement of the Julia scalar methods; Julia itself is not in the image)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _bind_near_gpu(torch, local):
    """Pin this process (and with it the first-touch placement of its pinned host buffers) to the CPUs of the GPU's NUMA node."""
    try:
        prop = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(f"{base}/numa_node").read().strip()
        cpus = open(f"{base}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = ids & allowed
        if use:
            os.sched_setaffinity(0, use)
        return {"numa_node": int(node), "cpus": len(use) if use else len(allowed), "bound": bool(use)}
    except Exception as e:  # noqa
        return {"numa_node": None, "bound": False, "why": str(e)[:80]}


def _link_rates(torch, dev, barrier, mb=256, reps=4):
    """Measured pinned-host <-> device copy rates of this rank [GB/s] (all ranks copy at the same time: the shared-host ceiling)."""
    nbytes = mb << 20
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = {}
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        dst.copy_(src, non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[name] = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    # both directions at once (what the chunked pipeline does): each direction timed on its own stream
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(s1):
        ev[0].record()
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        ev[1].record()
    with torch.cuda.stream(s2):
        ev[2].record()
        for _ in range(reps):
            h2.copy_(d2, non_blocking=True)
        ev[3].record()
    torch.cuda.synchronize()
    out["h2d_while_d2h"] = nbytes * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9
    out["d2h_while_h2d"] = nbytes * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9
    return out


def _config5(torch, dist, dev, lib, args, rank, world):
    """BASELINE config 5 on this rank's slab: fused 1M + 2M + ice nucleation (+ ARG2000) with in-kernel diagnostics, then the
    all-reduce of the 4 diagnostic doubles over NCCL through the C-ABI."""
    from cumicro import CMP, fused
    from cumicro.testing import arg_test_distribution, synthetic_states_fused
    n = args.points
    st = synthetic_states_fused(n, seed=4321 + rank)
    cols = [torch.from_numpy(st[k]).to(dev) for k in fused.IN_NAMES]
    tps = CMP.ThermodynamicsParameters(np.float64)
    mp1, mp2 = CMP.Microphysics1MParams(np.float64), CMP.Microphysics2MParams(np.float64)
    blk3 = CMP.pack_icenuc(tps, ad=arg_test_distribution("kappa"), dust=CMP.DustType("Kaolinite"), hom_linear=True)
    outs = [torch.empty_like(cols[0]) for _ in fused.OUT_NAMES]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    comm = C.c_void_p(0)
    if world > 1:
        # the host model's communicator: unique id from rank 0, exchanged over the existing process group
        idbuf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            st_ = lib.cumicro_nccl_unique_id(raw)
            if st_ != 0:
                return {"error": lib.cumicro_last_error().decode()}
            idbuf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        idd = idbuf.to(dev)
        dist.broadcast(idd, 0)
        raw = (C.c_char * 128).from_buffer_copy(bytes(idd.cpu().numpy().tobytes()))
        st_ = lib.cumicro_nccl_comm_init_rank(C.byref(comm), world, raw, rank)
        if st_ != 0:
            return {"error": lib.cumicro_last_error().decode()}

    def step(reduce):
        r = fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *cols, out=outs, diagnostics=True, reduce=False)
        if reduce and world > 1:
            st_ = lib.cumicro_nccl_allreduce_f64(comm, C.c_void_p(r["diag"].data_ptr()), C.c_int64(4), stream)
            assert st_ == 0, lib.cumicro_last_error()
        return r

    def timed(reduce, k, step=step):
        for _ in range(3):
            step(reduce)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            r = step(reduce)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / k], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    k = max(10, min(args.steps, 50))
    ms_red, r = timed(True, k)
    ms_loc, _ = timed(False, k)
    diag = r["diag"].cpu().numpy().tolist()
    if world > 1:
        lib.cumicro_nccl_comm_destroy(comm)

    # ---- the same step with the exchange inside the fused call's finish kernel (peer-memory stores over NVLink, csrc/cm_p2p.cuh)
    p2p = None
    if world > 1:
        from cumicro import collective
        win, err = None, ""
        try:
            win = collective.P2PWindow(rank, world).set_timeout(2.0).connect_with_torch_distributed()
        except Exception as e:  # noqa  (no peer access / IPC in this container: reported, the NCCL leg stands)
            err = repr(e)[:200]
        ok = torch.tensor([1.0 if (win is not None and not err) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)           # every rank connected, or nobody runs the leg
        if float(ok.item()) == 1.0:
            def step_p2p(_reduce):
                return fused.fused_1m2m_icenuc(mp1, mp2, tps, blk3, *cols, out=outs, diagnostics=True, p2p_window=win)
            ms_p2p, r2 = timed(True, k, step_p2p)
            calls, bad = win.status()
            d2 = r2["diag"].cpu().numpy()
            ref = np.asarray(diag)
            p2p = {"ms_per_step": ms_p2p, "value": n * world / (ms_p2p * 1e-3), "unit": UNIT,
                   "exchange_us": max(0.0, (ms_p2p - ms_loc) * 1e3), "calls": calls, "timed_out_call": bad,
                   "max_rel_diff_vs_nccl": float(np.max(np.abs(d2 - ref) / np.maximum(np.abs(ref), 1e-300))),
                   "collective": "peer-memory stores + rank-ordered sum inside the fused call's finish kernel "
                                 "(cumicro_fused_1m2m_icenuc_p2p_f64): no second launch, no library call"}
        else:
            p2p = {"unavailable": err or "a peer rank could not map the windows"}
        dist.barrier()
        torch.cuda.synchronize()
        if win is not None:
            win.destroy()
    return {"workload": "fused 1M + 2M + ice nucleation (+ARG2000, 3 modes) Float64, column slabs of 2^24 points per GPU, "
                        "in-kernel domain diagnostics + NCCL all-reduce of 4 doubles per step (cumicro_nccl_allreduce_f64)",
            "points_per_gpu": n, "global_points": n * world, "steps": k, "ms_per_step": ms_red,
            "value": n * world / (ms_red * 1e-3), "unit": UNIT,
            "ms_per_step_without_allreduce": ms_loc, "allreduce_us": max(0.0, (ms_red - ms_loc) * 1e3),
            "collective": "ncclAllReduce(sum, 4 x f64) per step" if world > 1 else "none (single GPU)",
            "diag_global": dict(zip(fused.DIAG_NAMES, diag)), "p2p": p2p}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cumicro", choices=["cumicro", "reference"])
    ap.add_argument("--points", type=int, default=1 << 24, help="grid points per GPU (and per step of the reference arm)")
    ap.add_argument("--cpu-sample", type=int, default=1 << 24, help="points of the cpu_baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustained-s", type=float, default=1.2, help="length of the sustained leg [s]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config5", action="store_true", help="run the config-5 leg at N = 1 too")
    ap.add_argument("--no-config5", action="store_true")
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
    numa = _bind_near_gpu(torch, local)      # before any pinned allocation: first touch lands on the GPU's NUMA node
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

    def step():
        return BMT.bulk_microphysics_tendencies(scheme, mp, tps, *[cols[k] for k in KEYS], out=outs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    # ---- timed region: EXACTLY K steps, CUDA events on the launching (current) stream -------------
    # Two events bracket the K launches; the timed region holds nothing but the K kernels, so the kernel's average launch duration is
    # total / K.  (An event between every two launches — the earlier form — costs ~7 us per step: tools/launch_probe2.py.)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    launches0 = lib.cumicro_launch_count()
    with ClockSampler(local) as clocks:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            step()
        ev[1].record()
        barrier()
    launches = lib.cumicro_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[1])
    total_ms_max = max_over_ranks(total_ms)
    value = n * world * args.steps / (total_ms_max * 1e-3)
    kernel_ms = total_ms / args.steps

    # ---- FP64 pipe peak measured in place (same power / clock state as the leg it is compared with) -------------
    def fp64_probe():
        scratch = torch.zeros(8, dtype=torch.float64, device=dev)
        flops = C.c_double(0)
        best = None
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.cumicro_probe_fp64_fma(C.c_int64(1 << 15), 8, C.c_void_p(scratch.data_ptr()), C.byref(flops),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
            e1.record()
            torch.cuda.synchronize()
            if rc == 0:
                best = max(best or 0.0, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best

    fp64_peak = fp64_probe()

    # ---- sustained leg: the same step for >= 1 s (clock under load, not a burst) -----------------------
    n_sus = max(args.steps, int(args.sustained_s * 1e3 / max(kernel_ms, 1e-3)) + 1)
    with ClockSampler(local) as clocks_sus:
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_sus):
            step()
        e1.record()
        barrier()
    sus_ms = max_over_ranks(e0.elapsed_time(e1))
    sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / n_sus,
                 "value": n * world * n_sus / (sus_ms * 1e-3), "unit": UNIT, "clocks": clocks_sus.summary(),
                 "fp64_probe_tflops": fp64_probe()}

    # ---- e2e: host buffers in, host buffers out, through the C-ABI host entry point --------
    link = _link_rates(torch, dev, barrier)
    host_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(4)]
    BMT.bulk_microphysics_tendencies_host(scheme, mp, tps, *[pinned[k] for k in KEYS], out=host_out)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        BMT.bulk_microphysics_tendencies_host(scheme, mp, tps, *[pinned[k] for k in KEYS], out=host_out)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n * world * args.e2e_steps / e2e_s
    h2d_bytes, d2h_bytes = 7 * 8 * n, 4 * 8 * n
    # the longer of the two copy legs at this rank's measured link rate (both directions busy) bounds the step
    link_bound_s = max(h2d_bytes / (link["h2d_while_d2h"] * 1e9), d2h_bytes / (link["d2h_while_h2d"] * 1e9))
    link_all = None
    if dist is not None:
        g = [None] * world
        dist.all_gather_object(g, {"rank": rank, **{k: round(v, 2) for k, v in link.items()}, "numa": numa})
        link_all = g

    # ---- config 5 with the only collective of the path -------------------------------------
    cfg5 = None
    if (world > 1 or args.config5) and not args.no_config5:
        try:
            cfg5 = _config5(torch, dist, dev, lib, args, rank, world)
        except Exception as e:  # noqa
            cfg5 = {"error": repr(e)[:300]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    frac_of_link = (link_bound_s * args.e2e_steps) / e2e_s
    link_note = "frac_of_link = time the longer copy leg needs at this rank's measured pinned-copy rate / measured e2e time"
    if link_all:
        # N > 1: every rank measured its link while ALL ranks were copying (the barrier aligns them: the host memory system's
        # worst case), and the ranks' rates differ (two PCIe / NUMA groups).  The job-level ceiling is the aggregate: all ranks'
        # bytes of the longer leg over the SUM of the measured concurrent rates of that direction.
        agg_h2d = sum(r["h2d_while_d2h"] for r in link_all) * 1e9
        agg_d2h = sum(r["d2h_while_h2d"] for r in link_all) * 1e9
        agg_bound_s = max(h2d_bytes * world / agg_h2d, d2h_bytes * world / agg_d2h)
        frac_of_link = (agg_bound_s * args.e2e_steps) / e2e_s
        link_note = ("frac_of_link = time the longer copy leg of ALL ranks needs at the SUM of the ranks' pinned-copy rates, measured while every rank "
                     "copies in both directions at once / measured e2e time (max over ranks); per-rank rates in `ranks`; above 1 when the chunked copies of "
                     "desynchronised ranks contend less for the host memory system than the barrier-aligned microbenchmark does")
    hbm_peak, peak_src = _peaks()
    ri = _roofline_inputs()
    achieved_gbs = BYTES_PER_POINT * n / (kernel_ms * 1e-3) / 1e9
    pts_per_gpu_s = n / (kernel_ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": n, "global_points": n * world,
                   "parallelism": f"column slabs x{world}, no data-path collective (config5 leg: + NCCL all-reduce of 4 diagnostic doubles per step)",
                   "l2": "inputs (7 x 134 MB columns) larger than the 126 MB L2; no flush needed",
                   "psd": "SB2006 limited rain PSD, log-uniform number concentrations"},
        "sustained": sustained,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": args.e2e_steps, "api": "cumicro_bmt2m_warm_host_f64 (pinned host buffers, chunked H2D/kernel/D2H)",
                "link_gbs": {k: round(v, 2) for k, v in link.items()},
                "frac_of_link": frac_of_link,
                "link_note": link_note,
                "numa": numa, "ranks": link_all},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    # ---- rooflines.  The binding resource is the FP64 pipe: algorithmic flops per point (the direct
    # formulation's FP64 work, counted by ncu on the round-1 kernel and used by the round-1 review) over the in-place DFMA probe.
    fl_alg = ri.get("fp64_flops_per_point_algorithmic", 556.5881729125977)
    roof = {"bound": "fp64", "unit": "TFLOP/s", "traffic": None, "kernel": KERNEL, "kernel_ms": kernel_ms,
            "flops_per_point_algorithmic": fl_alg,
            "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                    "bytes_per_point": BYTES_PER_POINT, "peak_source": peak_src,
                    "traffic_ncu": ri.get("dram_bytes_per_launch_2m_warm"),
                    "traffic_note": "DRAM bytes per launch from the committed ncu capture (profiles/), not measured in this run"}}
    if fp64_peak:
        ach = fl_alg * pts_per_gpu_s / 1e12
        roof.update({"achieved": ach, "peak": fp64_peak, "frac": ach / fp64_peak,
                     "peak_source": "cumicro_probe_fp64_fma timed in this run (in-place DFMA chain, 2 flops each)"})
        fe, ie, oe = ri.get("fp64_flops_per_point_executed"), ri.get("fp64_pipe_inst_per_point_executed"), ri.get("other_inst_per_point_executed")
        if fe and ie:
            # executed view: an FP64 instruction holds the sub-partition's issue port for two cycles, every other instruction for one
            inst_rate = fp64_peak * 1e12 / 2.0            # FP64 thread-instructions / s the pipe can start
            roof["executed"] = {"flops_per_point": fe, "fp64_pipe_inst_per_point": ie, "other_inst_per_point": oe,
                                "tflops": fe * pts_per_gpu_s / 1e12, "fp64_pipe_frac": ie * pts_per_gpu_s / inst_rate,
                                "issue_cycles_frac": ((2 * ie + oe) * pts_per_gpu_s / (2 * inst_rate)) if oe else None,
                                "source": ri.get("source")}
    else:
        roof.update({"achieved": None, "peak": None, "frac": None})
    line["roofline"] = roof
    line["fp64_probe_tflops"] = fp64_peak
    if cfg5 is not None:
        line["config5"] = cfg5

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
                                "sample": f"the workload's {m} points x {reps} passes, OpenMP over points "
                                          "(C++ restatement of the reference's scalar Julia methods)"}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the H-bond hot path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config C3): 100,000-atom membrane-protein-like synthetic system x 10,000 frames, triclinic
box, `set_hbonds_in_selection_and_water_around(3.5)`, 3.5 A + 60 deg angle criterion, selection `protein`.
One "step" = one full pass of the hot path over all 10,000 frames.

* `value`  : frames/s with the frames already resident in HBM (hb_submit_frames_device), device-timed.
* `e2e`    : frames/s through the C ABI with HOST (pinned) frames: H2D streaming of all frames and the D2H
             read of bond keys + occupancy matrix inside the timed region.
* `roofline`: algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak.
* `cpu_baseline`: the CPU oracle port (1 core) on a bounded sample of the same frames.
* `--impl reference`: the CPU path with all host cores on a bounded sample per step.

Multi-GPU (torchrun, one rank per GPU): weak scaling - every rank processes its own 10,000-frame block of
one long trajectory; the only exchange is the all-gather of distinct bond keys at the end of each pass.
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

WORKLOAD = "C3"
SELECTION = "protein"
AROUND = 3.5


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """DRAM bytes per frame of `kernel` from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        d = json.load(open(p))
        return d[kernel]
    except Exception:
        return None


def build_case(scale, n_frames):
    from cgraphs_b200 import synth
    system, cfg = synth.config(WORKLOAD, scale=scale)
    cfg["n_frames"] = n_frames
    return system, cfg


def algorithmic_bytes(top, n_bonds):
    """12 B x distinct atoms whose coordinates the method reads + one output bit per bond."""
    n_part = len(top.sel_atom) + len(top.wat_atom) + len(np.unique(top.h_atom))
    return 12 * n_part + (n_bonds + 7) // 8, n_part


# ------------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------------
_CPU_SHARED = {}     # inherited by forked workers: no pickling of topology or frames


def _cpu_worker(span):
    from oracle import hbond_port
    pt, xyz, method, around = _CPU_SHARED["pt"], _CPU_SHARED["xyz"], _CPU_SHARED["method"], _CPU_SHARED["around"]
    lo, hi = span
    t = time.perf_counter()
    res = hbond_port.run_block(pt, xyz[lo:hi], method, around, nb_frames=hi - lo)
    return time.perf_counter() - t, len(res)


def cpu_port_topology(system, cfg, frames_xyz):
    from oracle import hbond_port
    from cgraphs_b200.universe import Universe
    u = Universe(system.topology, frames_xyz[:1])
    return hbond_port.PortTopology(u, SELECTION, check_angle=cfg["check_angle"],
                                   add_donors_without_hydrogen=cfg["add_donors_without_hydrogen"])


def cpu_port_rate(system, cfg, frames_xyz, n_procs, pool=None):
    """frames/s of the oracle port over frames_xyz using n_procs processes (frame sharded)."""
    from oracle import hbond_port
    method = "water_around"
    if n_procs <= 1:
        pt = cpu_port_topology(system, cfg, frames_xyz)
        t = time.perf_counter()
        hbond_port.run_block(pt, frames_xyz, method, AROUND, nb_frames=len(frames_xyz))
        dt = time.perf_counter() - t
        return len(frames_xyz) / dt, dt
    bounds = np.linspace(0, len(frames_xyz), n_procs + 1).astype(int)
    spans = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    t = time.perf_counter()
    pool.map(_cpu_worker, spans, chunksize=1)
    dt = time.perf_counter() - t
    return len(frames_xyz) / dt, dt


def run_reference_arm(args):
    """The CPU path on all host cores: the validated port (the reference itself needs /root/reference and
    MDAnalysis stubs, which do not travel to the GPU box), frame-sharded over one process per core."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import multiprocessing as mp
    scale = args.scale
    system, cfg = build_case(scale, args.frames)
    cores = os.cpu_count() or 1
    per_step = int(min(512, max(cores * 16, 64)))
    xyz = system.frames(0, per_step)
    _CPU_SHARED.update(pt=cpu_port_topology(system, cfg, xyz), xyz=xyz, method="water_around", around=AROUND)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(0, 1)] * cores, chunksize=1)          # spin-up, untimed
        for i in range(args.warmup + args.steps):
            rate, dt = cpu_port_rate(system, cfg, xyz, cores, pool)
            if i >= args.warmup:
                times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = per_step / (ms / 1000.0)
    sample = "%d frames of the %d-atom workload per step, frame-sharded over %d processes" % (per_step, system.n_atoms, cores)
    line = {"impl": "reference", "metric": "hbond_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 distance / f32 angle", "data": "synthetic",
            "config": workload_config(system, cfg, args),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def wires_leg(args, system, d_xyz, local_rank, peak, bytes_frame):
    """Secondary measurement: WireAnalysis.set_water_wires(max_water=5) on the first frames of the
    same trajectory, resident in HBM; device time from hb_begin to the end of hb_finalize (CUDA events inside the
    library). Reported next to the headline, never part of it."""
    import torch
    from cgraphs_b200.mdhbond import WireAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    nf = min(args.wire_frames, d_xyz.shape[0])
    na = system.n_atoms
    max_water = 5
    wa = WireAnalysis(SELECTION, system.universe(1), device=local_rank)
    top = wa._topology
    tables, _ = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = wa._context()
    ctx.set_topology(tables)
    ctx.set_wires(wa._wire_ent_node, wa._n_wire_nodes, max_water, True)
    ctx.set_params(3.5, (max_water + 1) * 3.5 / 2., cos_threshold(60.0), True, False, _native.METHOD_WATER_AROUND)
    ctx.set_option("profile", 1)
    ms, kernels, nb, tm = [], {}, 0, None
    for it in range(3 + 5):
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        nb, _ = ctx.finalize()
        torch.cuda.synchronize()
        tm = ctx.timers()
        if it >= 3:
            ms.append(tm["ms_run"])
            for k, v in tm["ms_kernel"].items():
                if v:
                    kernels[k] = kernels.get(k, 0.0) + v / 5
    ms_pass = float(np.mean(ms))
    value = nf / (ms_pass * 1e-3)
    out = {"metric": "water_wire_frames_per_s", "value": value, "unit": "frames/s", "ms_per_pass": ms_pass,
           "workload": "first %d frames of the same trajectory resident in HBM, WireAnalysis.set_water_wires(max_water=%d): "
                       "10.5 A water shell, 3.5 A + 60 deg, selection '%s'" % (nf, max_water, SELECTION),
           "n_residue_pairs": int(nb), "edges_per_frame": tm["pairs_emitted"] / nf, "launches_per_pass": int(tm["n_launches"]),
           "kernel_ms_per_pass": {k: round(v, 3) for k, v in kernels.items()},
           "hbm_frac": bytes_frame * value / 1e9 / peak}
    if not args.no_e2e:
        # the same pass from pinned host memory (host -> device copies and the result read back inside the timed region)
        pin = _native.PinnedBuffer((nf, na, 3))
        try:
            host = pin.array
            chunk = max(1, (1 << 30) // (na * 12))
            for s0 in range(0, nf, chunk):
                host[s0:s0 + chunk] = d_xyz[s0:min(s0 + chunk, nf)].cpu().numpy()

            def one_pass_host():
                ctx.begin(nf)
                ctx.submit_raw(host.ctypes.data, na, nf)
                return ctx.results()

            one_pass_host()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                keys, words = one_pass_host()
            torch.cuda.synchronize()
            e2e_ms = 1000.0 * (time.perf_counter() - t0) / 3
            out["e2e"] = {"value": nf / (e2e_ms / 1000.0), "unit": "frames/s", "ms_per_pass": e2e_ms,
                          "h2d_bytes_per_step": int(ctx.timers()["h2d_bytes"]),
                          "d2h_bytes_per_step": int(keys.nbytes + words.nbytes), "host_memory": "pinned"}
        finally:
            pin.free()
    if not args.no_cpu:
        from oracle import hbond_port                   # CPU baseline leg: the validated port on a bounded sample
        n_cpu = 4
        sample = d_xyz[:n_cpu].cpu().numpy()
        from cgraphs_b200.universe import Universe
        t0 = time.perf_counter()
        hbond_port.run_wires(Universe(system.topology, sample, system.dimensions), SELECTION, max_water=max_water)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "first %d frames, oracle/hbond_port.py run_wires (%.1f s)" % (n_cpu, dt)}
    return out


def workload_config(system, cfg, args):
    return {"workload": "C3: %d-atom membrane-protein-like system x %d frames per GPU, triclinic box, "
                        "set_hbonds_in_selection_and_water_around(3.5), 3.5 A + 60 deg, selection 'protein'"
                        % (system.n_atoms, args.frames),
            "n_atoms": system.n_atoms, "n_frames_per_gpu": args.frames, "selection": SELECTION, "method": "water_around",
            "check_angle": True, "pbc": "none (reference semantics; skin system)",
            "cache": "inputs (%.1f GB of frames per GPU) are far larger than the 126 MB L2" % (system.n_atoms * 12 * args.frames / 1e9)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from cgraphs_b200 import build as hb_build
    hb_build.build()
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    from cgraphs_b200.mdhbond.sharding import exchange_keys_queued, merge_keys, merge_keys_device

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device is visible (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    system, cfg = build_case(args.scale, args.frames)
    nf, na = args.frames, system.n_atoms
    f0 = rank * nf                                      # weak scaling: each rank owns its own frame block
    # frames generated on the device (synthetic data), then one copy to pinned host memory for the e2e arm
    d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(f0, f0 + nf, dev, out=d_xyz)
    torch.cuda.synchronize()

    u = system.universe(1)
    hba = HbondAnalysis(SELECTION, u, check_angle=cfg["check_angle"], device=local_rank,
                        add_donors_without_hydrogen=cfg["add_donors_without_hydrogen"])
    top = hba._topology
    tables, group_names = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = hba._context()
    ctx.set_topology(tables)
    ctx.set_params(3.5, AROUND, cos_threshold(60.0), True, True, _native.METHOD_WATER_AROUND)
    ctx.set_option("profile", 1)
    if args.batch_bytes:
        ctx.set_option("batch_bytes", args.batch_bytes)
    # every kernel of the library and the NCCL exchange run on one torch stream, so a pair of CUDA events on that
    # stream brackets whole passes (search + finalize + key merge)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_option("compute_stream", stream.cuda_stream)
    torch.cuda.set_stream(stream)

    merged = []

    def one_pass_device(sync_merge=True):
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        if world > 1 and not sync_merge:
            # steady state: finalisation, the one exchange (distinct bond keys, GPU to GPU over NCCL) and the merge
            # are all queued behind the search; the host waits once, in finalize()
            ctx.finalize_async()
            merged.append(exchange_keys_queued(ctx)[1])  # [n_keys, overflow] on the device, checked after the loop
            return ctx.finalize()
        nb, wpr = ctx.finalize()
        if world > 1:                                   # warm-up: synchronous exchange, learns the gather capacity
            k = torch.empty(max(nb, 1), dtype=torch.int64, device=dev)
            ctx.keys_to_device(k.data_ptr(), sync=False)
            merge_keys_device(k[:nb], ctx=ctx, sync=True)
        return nb, wpr

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()                  # nvidia-smi every 200 ms from the warm-up on: the timed region itself is ~12 ms
    t_sampler = time.perf_counter()
    for _ in range(args.warmup):
        one_pass_device()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    launches = 0
    retries = 0
    kernel_ms = {}
    kernel_launches = {}
    ev0.record(stream)
    for _ in range(args.steps):
        nb, wpr = one_pass_device(sync_merge=False)     # gather capacity learnt during warm-up; no host wait here
        tm = ctx.timers()
        launches += tm["n_launches"]
        retries += tm["finalize_retries"]
        for k, v in tm["ms_kernel"].items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v
            kernel_launches[k] = kernel_launches.get(k, 0) + tm["launches"][k]
    ev1.record(stream)
    sync_all()
    wall_ms = 1000.0 * (time.perf_counter() - t0)
    dev_ms = ev0.elapsed_time(ev1)      # CUDA events on the launching stream around exactly `steps` passes
    while time.perf_counter() - t_sampler < 1.2:     # keep the same load up (untimed, rank-local: no collective)
        ctx.begin(nf)                                # until the clock sampler has a few readings
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        ctx.finalize()
    torch.cuda.synchronize()
    n_global_keys = None
    if retries:
        raise SystemExit("bench: a finalisation was redone inside the timed loop (bond table overflow)")
    for info in merged:
        n_out, over = info.tolist()
        if over:
            raise SystemExit("bench: the key gather overflowed its learnt capacity inside the timed loop")
        n_global_keys = n_out
    clocks = sampler.stop()
    # max over ranks of the device-timed pass (CUDA events begin -> finalize), cross-checked by wall clock
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    per_rank = None
    if dist is not None:      # per-rank view: device ms per step and the fused kernel's ms per step
        mine = torch.tensor([dev_ms / args.steps, kernel_ms.get("frame_fused", 0.0) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(x[0]), 4), round(float(x[1]), 4)] for x in allr]
    ms_per_step = dev_ms_max / args.steps
    value = world * nf / (ms_per_step / 1000.0)

    # ---- e2e: host frames through the C ABI (H2D inside the timed region, results read back) ----
    e2e = None
    if not args.no_e2e:
        # host memory: all frames of the pass at N=1; with several ranks on one host the pinned budget is shared,
        # so each rank streams a bounded number of distinct frames per pass (same metric, shorter pass)
        ef = nf if world == 1 else int(min(nf, max(2048, nf // world)))
        try:
            pin = _native.PinnedBuffer((ef, na, 3))
            host = pin.array
            pinned = True
        except Exception:
            pin, host, pinned = None, np.empty((ef, na, 3), dtype=np.float32), False
        chunk = max(1, (1 << 30) // (na * 12))
        for s in range(0, ef, chunk):
            host[s:s + chunk] = d_xyz[s:min(s + chunk, ef)].cpu().numpy()

        def one_pass_host():
            ctx.begin(ef)
            ctx.submit_raw(host.ctypes.data, na, ef)
            keys, words = ctx.results()
            if world > 1:
                merge_keys(keys, device=local_rank)
            return keys, words

        one_pass_host()
        sync_all()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            keys, words = one_pass_host()
        sync_all()
        e2e_ms = 1000.0 * (time.perf_counter() - t0)
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0]) / e2e_steps
        e2e = {"value": world * ef / (e2e_ms / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": int(ctx.timers()["h2d_bytes"]),
               "frames_per_step_per_gpu": ef, "frame_bytes_per_step": int(ef) * na * 12,
               "d2h_bytes_per_step": int(keys.nbytes + words.nbytes), "ms_per_step": e2e_ms,
               "host_memory": "pinned" if pinned else "pageable", "timer": "wall clock between device synchronisations"}
        if pin is not None:
            pin.free()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peak, peak_kind = measured_peaks()
    bytes_frame, n_part = algorithmic_bytes(top, nb)
    dom = max(kernel_ms, key=lambda k: kernel_ms[k]) if kernel_ms else None
    roofline = None
    if dom and kernel_ms[dom] > 0:
        n_l = kernel_launches[dom]
        frames_per_launch = args.steps * nf / n_l
        avg_ms = kernel_ms[dom] / n_l
        achieved = bytes_frame * frames_per_launch / (avg_ms * 1e-3) / 1e9
        tr = ncu_traffic(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    "traffic": tr["dram_bytes_per_frame"] * frames_per_launch if tr else None,
                    "traffic_source": tr["source"] if tr else None, "peak_source": peak_kind,
                    "algorithmic_bytes_per_launch": bytes_frame * frames_per_launch,
                    "algorithmic_bytes_per_frame": bytes_frame, "frames_per_launch": frames_per_launch,
                    "avg_launch_ms": avg_ms, "kernel_share_of_step": kernel_ms[dom] / sum(kernel_ms.values()),
                    "whole_path_frac": bytes_frame * (nf / (ms_per_step / 1e3)) / 1e9 / peak,
                    "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()}}

    # ---- CPU baseline: oracle port, 1 core, bounded sample of the same frames -----------------------
    cpu = None
    if not args.no_cpu:
        sample_n = args.cpu_frames
        sample = d_xyz[:sample_n].cpu().numpy()
        rate, dt = cpu_port_rate(system, cfg, sample, 1)
        cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": "first %d frames of the same %d-atom trajectory (%.1f s of CPU work), oracle/hbond_port.py"
                         % (sample_n, na, dt)}

    line = {"metric": "hbond_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 distance / f32 angle", "data": "synthetic",
            "config": workload_config(system, cfg, args),
            "atom_frames_per_s": value * na, "wall_ms_per_step": wall_ms_max / args.steps,
            "n_bonds": int(nb), "n_global_bonds": n_global_keys, "n_part_atoms": int(n_part),
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu}
    if world == 1 and args.wire_frames > 0:
        try:
            line["wires"] = wires_leg(args, system, d_xyz, local_rank, peak, bytes_frame)
        except Exception as e:                          # the secondary leg never takes the headline down
            line["wires"] = {"error": repr(e)[:300]}
    if per_rank is not None:
        line["per_rank_ms_per_step"] = {"columns": ["pass", "frame_fused kernel"], "ranks": per_rank}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=10000, help="frames per GPU (BASELINE config: 10000)")
    ap.add_argument("--scale", type=float, default=1.0, help="atom-count scale of the workload (1.0 = BASELINE config)")
    ap.add_argument("--cpu-frames", type=int, default=1024)
    ap.add_argument("--batch-bytes", type=int, default=0)
    ap.add_argument("--wire-frames", type=int, default=2048, help="frames of the secondary water-wire leg (0: skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("HB_BENCH_ALLOW_SHORT_WARMUP"):
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the H-bond hot path on BASELINE.json's headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scaling weak|strong]

Workload (config C3): 100,000-atom membrane-protein-like synthetic system x 10,000 frames, triclinic
box, `set_hbonds_in_selection_and_water_around(3.5)`, 3.5 A + 60 deg angle criterion, selection `protein`.
One "step" = one full pass of the hot path over all frames of the rank.

* `value`   : frames/s with the frames already resident in HBM (hb_submit_frames_device), device-timed.
* `parity`  : the CPU oracle port runs on the first frames of the very trajectory the GPU searched; the GPU's
              occupancy columns of those frames must equal the port's result dict, else the run exits non-zero.
* `e2e`     : frames/s through the C ABI with HOST (pinned) buffers inside the timed region: the XTC records of
              the trajectory (compressed, ~4 bytes per atom; decoded on the device) when that is the faster
              source, else the float32 frames; both legs are reported (`e2e_f32`, `e2e_xtc`), and `e2e_class`
              is the call a cgraphs user makes (HbondAnalysis.set_hbonds_in_selection_and_water_around on a
              numpy-backed Universe, result dict built).
* `roofline`: algorithmic bytes of the dominant kernel / its CUDA-event time, against the measured HBM peak.
* `cpu_baseline`: the CPU oracle port (1 core) on a bounded sample of the same frames.
* `--impl reference`: the reference's own CPU code (oracle/_ref, byte-compiled from /root/reference by
              oracle/make_ref.py) when present, else the validated port; all host cores, bounded sample per step.

Multi-GPU (torchrun, one rank per GPU): frames are sharded, the only exchange is the all-gather of distinct
bond keys at the end of each pass. `--scaling weak` (default): every rank processes its own `--frames` block
of one long trajectory; `--scaling strong`: `--frames` is the total, split over the ranks.
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
_WORKLOAD = [WORKLOAD]       # --workload C4 switches the device-resident legs to BASELINE config C4 (1M atoms; general kernels)
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
    system, cfg = synth.config(_WORKLOAD[0], scale=scale)
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


def _ref_available():
    return os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "cgraphs", "mdhbond", "hbond.pyc.bin"))


def _ref_process(conn, topology, dims, xyz):
    """A worker that owns ONE reference HbondAnalysis (built once, like ours) over its frame block and runs the
    reference's own frame loop (hbond.py:232-289) every time it is told to."""
    os.environ["CGRAPHS_REF"] = os.path.join(ROOT, "oracle", "_ref", "cgraphs")
    from oracle import ref_harness
    from cgraphs_b200.universe import Universe
    ref = ref_harness.load_reference()
    hba = ref.HbondAnalysis(SELECTION, Universe(topology, xyz, dims), check_angle=True)
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        t = time.perf_counter()
        hba.set_hbonds_in_selection_and_water_around(AROUND)
        conn.send((time.perf_counter() - t, len(hba.initial_results)))


def run_reference_arm(args):
    """The reference's CPU implementation on all host cores, frame-sharded over one process per core: the reference
    itself (oracle/_ref: the reference's mdhbond package byte-compiled by oracle/make_ref.py, run through MDAnalysis
    stubs) when it travelled to this box, else the validated port."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import multiprocessing as mp
    system, cfg = build_case(args.scale, args.frames)
    cores = os.cpu_count() or 1
    use_ref = _ref_available() and not args.port_baseline
    per_step = int(min(512, max(cores * (4 if use_ref else 16), 64)))
    xyz = system.frames(0, per_step)
    ctx = mp.get_context("fork")
    times = []
    if use_ref:
        bounds = np.linspace(0, per_step, cores + 1).astype(int)
        blocks = [xyz[a:b] for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        procs = []
        for blk in blocks:
            parent, child = ctx.Pipe()
            pr = ctx.Process(target=_ref_process, args=(child, system.topology, system.dimensions, blk), daemon=True)
            pr.start()
            procs.append((pr, parent))
        for _, c in procs:
            c.recv()                                                     # constructors done (untimed, like ours)

        def step():
            t = time.perf_counter()
            for _, c in procs:
                c.send("run")
            for _, c in procs:
                c.recv()
            return time.perf_counter() - t

        step()                                                           # first run, untimed
        for i in range(args.warmup + args.steps):
            dt = step()
            if i >= args.warmup:
                times.append(dt)
        for pr, c in procs:
            c.send("stop")
        for pr, c in procs:
            pr.join(timeout=10)
        kind, what = "reference", "cgraphs.mdhbond.HbondAnalysis (oracle/_ref, unmodified, MDAnalysis stubbed)"
    else:
        _CPU_SHARED.update(pt=cpu_port_topology(system, cfg, xyz), xyz=xyz, method="water_around", around=AROUND)
        with ctx.Pool(cores) as pool:
            pool.map(_cpu_worker, [(0, 1)] * cores, chunksize=1)          # spin-up, untimed
            for i in range(args.warmup + args.steps):
                rate, dt = cpu_port_rate(system, cfg, xyz, cores, pool)
                if i >= args.warmup:
                    times.append(dt)
        kind, what = "port", "oracle/hbond_port.py"
    ms = 1000.0 * float(np.mean(times))
    value = per_step / (ms / 1000.0)
    sample = "%d frames of the %d-atom workload per step, frame-sharded over %d processes, %s" % (
        per_step, system.n_atoms, cores, what)
    line = {"impl": "reference", "metric": "hbond_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64 distance / f32 angle", "data": "synthetic",
            "config": workload_config(system, cfg, args, env_int("WORLD_SIZE", 1)),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def results_of(keys, words, group_names, n_cols):
    """{key string: bool[n_cols]} of the rows that have a bond in the first n_cols frames."""
    from cgraphs_b200.mdhbond import HbondAnalysis
    full = HbondAnalysis._results_dict(keys, words, group_names, words.shape[1] * 32)
    out = {}
    for k, v in full.items():
        if v[:n_cols].any():
            out[k] = v[:n_cols].copy()
    return out


def parity_object(gpu, cpu, n_frames, what):
    a = {k: v.tobytes() for k, v in gpu.items()}
    b = {k: np.asarray(v, dtype=bool)[:n_frames].tobytes() for k, v in cpu.items() if np.asarray(v)[:n_frames].any()}
    equal = a == b
    out = {"checked_against": what, "frames": int(n_frames), "keys": len(b), "equal": bool(equal)}
    if not equal:
        out["only_gpu"] = sorted(set(a) - set(b))[:5]
        out["only_cpu"] = sorted(set(b) - set(a))[:5]
        out["different_rows"] = sorted(k for k in set(a) & set(b) if a[k] != b[k])[:5]
    return out


def bind_near_gpu(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host buffers it
    allocates afterwards land on that NUMA node (first touch) and its copy threads run there."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        info = {"bound": True, "cpus": "%d-%d (%d of %d)" % (cpus[0], cpus[-1], len(cpus), before)}
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = open("/sys/bus/pci/devices/%s/numa_node" % bus.lower()[-12:]).read().strip()
            info["numa_node"] = int(node)
        except Exception:
            pass
    except Exception as e:
        info["error"] = repr(e)[:120]
    return info


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
        # CPU baseline leg + parity gate: the validated port on the first frames; the GPU's wire lengths of those frames
        # (a fresh run over exactly these frames through the drop-in class) must be the same dict
        from oracle import hbond_port
        from cgraphs_b200.universe import Universe
        n_cpu = 4
        sample = d_xyz[:n_cpu].cpu().numpy()
        u = Universe(system.topology, sample, system.dimensions)
        t0 = time.perf_counter()
        ref = hbond_port.run_wires(u, SELECTION, max_water=max_water)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "first %d frames, oracle/hbond_port.py run_wires (%.1f s)" % (n_cpu, dt)}
        wg = WireAnalysis(SELECTION, u, device=local_rank)
        wg.set_water_wires(max_water=max_water)
        lists = lambda d: {k: [(-1 if not np.isfinite(x) else int(x)) for x in v] for k, v in d.items()}
        got, want = lists(wg.wire_lengths), lists(ref.wire_lengths)
        out["parity"] = {"checked_against": "oracle/hbond_port.py run_wires", "frames": n_cpu, "keys": len(want),
                         "equal": bool(got == want)}
    return out


def c4_leg(args, local_rank, peak):
    """BASELINE config C4 (1M atoms; the selection is too large for the one-CTA-per-frame kernel, so the general kernels
    run): a block of frames resident in HBM, device-timed like the headline; parity gate on its first frames."""
    import torch
    from cgraphs_b200 import synth
    from cgraphs_b200.mdhbond import HbondAnalysis, _native
    from cgraphs_b200.mdhbond.calib import cos_threshold
    from cgraphs_b200.mdhbond.topology import pack_systems
    system, cfg = synth.config("C4", scale=args.scale)
    nf, na = args.c4_frames, system.n_atoms
    dev = torch.device("cuda", local_rank)
    d_xyz = torch.empty((nf, na, 3), dtype=torch.float32, device=dev)
    system.frames_torch(0, nf, dev, out=d_xyz)
    torch.cuda.synchronize()
    hba = HbondAnalysis(SELECTION, system.universe(1), check_angle=True, device=local_rank)
    top = hba._topology
    tables, group_names = pack_systems([top], [top.key_ids("water_around")], [na])
    ctx = hba._context()
    ctx.set_topology(tables)
    ctx.set_params(3.5, AROUND, cos_threshold(60.0), True, True, _native.METHOD_WATER_AROUND)
    ctx.set_option("profile", 1)
    ctx.set_option("max_batch_frames", nf)
    ms, search, kernels, launches, nb = [], [], {}, 0, 0
    for it in range(3 + 5):
        ctx.begin(nf)
        ctx.submit_device(d_xyz.data_ptr(), na, nf)
        nb, _ = ctx.finalize()
        torch.cuda.synchronize()
        tm = ctx.timers()
        if it >= 3:
            ms.append(tm["ms_run"])
            search.append(tm["ms_total"])
            launches = int(tm["n_launches"])
            for k, v in tm["ms_kernel"].items():
                if v:
                    kernels[k] = kernels.get(k, 0.0) + v / 5
    ctx.begin(nf)
    ctx.submit_device(d_xyz.data_ptr(), na, nf)
    keys, words = ctx.results()
    bytes_frame, n_part = algorithmic_bytes(top, nb)
    ms_pass, ms_search = float(np.mean(ms)), float(np.mean(search))
    out = {"workload": "C4: %d-atom system, %d frames resident in HBM, selection 'protein' (%d selection points), "
                       "set_hbonds_in_selection_and_water_around(3.5), 3.5 A + 60 deg" % (na, nf, len(top.sel_atom)),
           "metric": "hbond_frames_per_s", "value": nf / (ms_pass * 1e-3), "unit": "frames/s", "ms_per_pass": ms_pass,
           "atom_frames_per_s": nf * na / (ms_pass * 1e-3), "n_bonds": int(nb), "launches_per_pass": launches,
           "kernel_ms_per_pass": {k: round(v, 3) for k, v in kernels.items()},
           "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": bytes_frame, "peak": peak, "unit": "GB/s",
                        "achieved_search_kernels": bytes_frame * nf / (ms_search * 1e-3) / 1e9,
                        "frac_search_kernels": bytes_frame * nf / (ms_search * 1e-3) / 1e9 / peak,
                        "achieved_whole_pass": bytes_frame * nf / (ms_pass * 1e-3) / 1e9,
                        "frac_whole_pass": bytes_frame * nf / (ms_pass * 1e-3) / 1e9 / peak}}
    if not args.no_cpu:
        n_cpu = 2
        sample = d_xyz[:n_cpu].cpu().numpy()
        from oracle import hbond_port
        pt = cpu_port_topology(system, cfg, sample)
        t0 = time.perf_counter()
        cpu_res = hbond_port.run_block(pt, sample, "water_around", AROUND, nb_frames=n_cpu)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": "first %d frames (%.1f s), oracle/hbond_port.py" % (n_cpu, dt)}
        out["parity"] = parity_object(results_of(keys, words, group_names, n_cpu), cpu_res, n_cpu,
                                      "oracle/hbond_port.py on the first frames of the C4 block")
    hba.close()
    del d_xyz
    torch.cuda.empty_cache()
    return out


def workload_config(system, cfg, args, world):
    nf_total = args.frames if args.scaling == "strong" else args.frames * world
    return {"workload": "%s: %d-atom membrane-protein-like system x %d frames %s, triclinic box, "
                        "set_hbonds_in_selection_and_water_around(3.5), 3.5 A + 60 deg, selection 'protein'"
                        % (_WORKLOAD[0], system.n_atoms, args.frames, "in total, sharded over the GPUs" if args.scaling == "strong" else "per GPU"),
            "n_atoms": system.n_atoms, "n_frames_per_gpu": args.frames if args.scaling == "weak" else None,
            "n_frames_total": nf_total, "scaling": args.scaling,
            "selection": SELECTION, "method": "water_around",
            "check_angle": True, "pbc": "none (reference semantics; skin system)",
            "cache": "inputs (%.1f GB of frames per GPU) are far larger than the 126 MB L2"
                     % (system.n_atoms * 12 * (nf_total / world) / 1e9)}


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
    from cgraphs_b200.mdhbond.sharding import exchange_keys_queued, merge_keys, merge_keys_device, shard_range

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device is visible (there is no CPU fallback)")
    numa = bind_near_gpu(local_rank) if not args.no_numa else {"bound": False}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    system, cfg = build_case(args.scale, args.frames)
    na = system.n_atoms
    if args.scaling == "strong":                        # args.frames frames in total, contiguous blocks of multiples of 32
        f0, f1 = shard_range(args.frames, rank, world)
        nf = f1 - f0
        nf_total = args.frames
    else:                                               # weak: each rank owns its own block of args.frames frames
        nf = args.frames
        f0 = rank * nf
        nf_total = nf * world
    if nf < 1:
        raise SystemExit("bench: rank %d has no frames (%d frames over %d ranks)" % (rank, args.frames, world))
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
    for kv in args.native_opt:
        ctx.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if args.workload == "C4":
        ctx.set_option("max_batch_frames", nf)
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
        # gather capacity learnt during warm-up; no host wait here (large bond tables - C4 - finalise synchronously:
        # their radix sort is sized from the bond count, so the exchange follows the finalisation)
        nb, wpr = one_pass_device(sync_merge=args.workload == "C4")
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
    # the device-resident result of this rank (same state every pass), kept for the parity gate below
    ctx.begin(nf)
    ctx.submit_device(d_xyz.data_ptr(), na, nf)
    dev_keys, dev_words = ctx.results()
    # max over ranks of the device-timed pass (CUDA events begin -> finalize), cross-checked by wall clock
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = float(t[0]), float(t[1])
    ms_per_step = dev_ms_max / args.steps
    value = nf_total / (ms_per_step / 1000.0)

    # ---- e2e legs: host buffers through the C ABI (H2D inside the timed region, results read back) ----
    def timed_passes(one_pass, steps):
        one_pass()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = one_pass()
        sync_all()
        ms = 1000.0 * (time.perf_counter() - t0)
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0]) / steps, res

    e2e_f32 = e2e_xtc = e2e_class = None
    parity_xtc = None
    merged_e2e = []
    e2e_steps = max(1, min(args.steps, 3))
    if not args.no_e2e:
        # host memory: all frames of the pass at N = 1; with several ranks on one host the pinned budget is shared, so
        # each rank keeps a ring of distinct frames and streams it as often as it takes to submit ALL its frames
        # (same frame count per pass as the device-resident leg, same bytes over PCIe)
        ring = nf if world == 1 else int(min(nf, args.ring_frames))
        pin = _native.PinnedBuffer((ring, na, 3))
        host = pin.array
        chunk = max(1, (1 << 30) // (na * 12))
        for s in range(0, ring, chunk):
            host[s:s + chunk] = d_xyz[s:min(s + chunk, ring)].cpu().numpy()

        def one_pass_f32():
            ctx.begin(nf)
            for s in range(0, nf, ring):
                ctx.submit_raw(host.ctypes.data, na, min(ring, nf - s))
            ctx.finalize_async()
            if world > 1:       # the queued device-side exchange of the device-resident loop (no host merge)
                merged_e2e.append(exchange_keys_queued(ctx)[1])
            return ctx.results()

        ms, (keys, words) = timed_passes(one_pass_f32, e2e_steps)
        tm = ctx.timers()
        e2e_f32 = {"value": nf_total / (ms / 1000.0), "unit": "frames/s", "source": "float32 frames [F][N][3] in pinned host memory",
                   "h2d_bytes_per_step": int(tm["h2d_bytes"]), "frames_per_step_per_gpu": nf, "distinct_frames_in_host_ring": ring,
                   "frame_bytes_per_step": int(nf) * na * 12, "d2h_bytes_per_step": int(keys.nbytes + words.nbytes),
                   "ms_per_step": ms, "h2d_gbs_per_gpu": tm["h2d_bytes"] / 1e9 / (ms / 1e3), "ms_h2d": tm["ms_h2d"],
                   "host_memory": "pinned", "timer": "wall clock between device synchronisations, max over ranks"}
        if world == 1 and ring == nf and not (np.array_equal(keys, dev_keys) and np.array_equal(words, dev_words)):
            raise SystemExit("bench: the host-frame pass and the device-resident pass disagree")

        # ---- XTC: the compressed records of the same frames (written once, outside the timed region) ----
        if not args.no_xtc:
            xring = ring if world == 1 else int(min(ring, args.ring_frames))
            t_enc = time.perf_counter()
            xbuf, xoff = _native.xtc_encode(host[:xring], in_scale=0.1, precision=1000.0)
            t_enc = time.perf_counter() - t_enc
            xpin = _native.PinnedBuffer((len(xbuf),), dtype=np.uint8)
            xpin.array[:] = xbuf
            if args.xtc_batch_bytes:
                ctx.set_option("batch_bytes", args.xtc_batch_bytes)

            def one_pass_xtc():
                ctx.begin(nf)
                for s in range(0, nf, xring):
                    k = min(xring, nf - s)
                    ctx.submit_xtc(xpin.array, xoff[:k + 1], na, k, scale=10.0)
                ctx.finalize_async()
                if world > 1:
                    merged_e2e.append(exchange_keys_queued(ctx)[1])
                return ctx.results()

            ms, (xkeys, xwords) = timed_passes(one_pass_xtc, e2e_steps)
            tm = ctx.timers()
            e2e_xtc = {"value": nf_total / (ms / 1000.0), "unit": "frames/s",
                       "source": "GROMACS XTC records (precision 1000) in pinned host memory, decoded on the device",
                       "h2d_bytes_per_step": int(tm["h2d_bytes"]), "frames_per_step_per_gpu": nf, "distinct_frames_in_host_ring": xring,
                       "xtc_bytes_per_atom": len(xbuf) / xring / na, "d2h_bytes_per_step": int(xkeys.nbytes + xwords.nbytes),
                       "ms_per_step": ms, "h2d_gbs_per_gpu": tm["h2d_bytes"] / 1e9 / (ms / 1e3), "ms_h2d": tm["ms_h2d"],
                       "ms_decode_kernel": tm["ms_kernel"].get("xtc_decode"), "encode_s_untimed": t_enc, "host_memory": "pinned",
                       "timer": "wall clock between device synchronisations, max over ranks"}
            if rank == 0 and not args.no_cpu:
                # parity gate of the XTC leg: oracle decode + oracle H-bond port on the first frames
                from oracle import xtc_port                  # (CPU-baseline leg: the only place bench.py runs the oracle)
                n_x = int(min(args.cpu_frames_xtc, xring))
                dec, _ = xtc_port.decode(xbuf, xoff[:n_x + 1], na, scale=10.0)
                pt = cpu_port_topology(system, cfg, dec)
                from oracle import hbond_port
                cpu_res = hbond_port.run_block(pt, dec, "water_around", AROUND, nb_frames=n_x)
                parity_xtc = parity_object(results_of(xkeys, xwords, group_names, n_x), cpu_res, n_x,
                                           "oracle/xtc_codec.c decode + oracle/hbond_port.py on the first frames of the XTC stream")
            if args.xtc_batch_bytes:
                ctx.set_option("batch_bytes", args.batch_bytes or (4 << 30))
            xpin.free()

        # ---- the call a cgraphs user makes: drop-in class on a numpy-backed Universe, result dict built ----
        if world == 1 and not args.no_class:
            from cgraphs_b200.universe import Universe
            try:
                import psutil
                room = psutil.virtual_memory().available
            except Exception:
                room = 0
            ncls = nf if room > 3 * host.nbytes else int(min(nf, 2048))
            pageable = np.array(host[:ncls])                           # an ordinary (pageable) numpy array
            hb2 = HbondAnalysis(SELECTION, Universe(system.topology, pageable, system.dimensions), check_angle=True,
                                device=local_rank)
            hb2.set_hbonds_in_selection_and_water_around(AROUND)
            torch.cuda.synchronize()
            tc = time.perf_counter()
            for _ in range(2):
                hb2.set_hbonds_in_selection_and_water_around(AROUND)
            torch.cuda.synchronize()
            ms = 1000.0 * (time.perf_counter() - tc) / 2
            e2e_class = {"value": ncls / (ms / 1000.0), "unit": "frames/s", "frames": ncls, "ms_per_call": ms,
                         "call": "HbondAnalysis.set_hbonds_in_selection_and_water_around(3.5) on a Universe over a pageable "
                                 "numpy array [F][N][3]; returns with initial_results (dict[str, bool[F]]) and the graph built",
                         "n_bonds": len(hb2.initial_results), "host_memory": "pageable (staged through pinned buffers by the library)"}
            hb2.close()
            del pageable
        pin.free()

    for info in merged_e2e:
        if int(info.tolist()[1]):
            raise SystemExit("bench: the key gather overflowed its learnt capacity inside an e2e pass")
    per_rank = None
    if dist is not None:      # per-rank view: device ms per step, the fused kernel's ms per step, e2e H2D GB/s
        mine = torch.tensor([dev_ms / args.steps, kernel_ms.get("frame_fused", 0.0) / args.steps,
                             e2e_f32["h2d_gbs_per_gpu"] if e2e_f32 else 0.0, e2e_xtc["h2d_gbs_per_gpu"] if e2e_xtc else 0.0],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 4) for v in x] for x in allr]

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

    # ---- CPU baseline + parity gate: oracle port, 1 core, the first frames of the very trajectory searched above ----
    cpu = None
    parity = None
    if not args.no_cpu:
        sample_n = int(min(args.cpu_frames, nf))
        sample = d_xyz[:sample_n].cpu().numpy()
        from oracle import hbond_port
        pt = cpu_port_topology(system, cfg, sample)
        tcpu = time.perf_counter()
        cpu_res = hbond_port.run_block(pt, sample, "water_around", AROUND, nb_frames=sample_n)
        dt = time.perf_counter() - tcpu
        cpu = {"value": sample_n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": "first %d frames of the same %d-atom trajectory (%.1f s of CPU work), oracle/hbond_port.py"
                         % (sample_n, na, dt)}
        parity = parity_object(results_of(dev_keys, dev_words, group_names, sample_n), cpu_res, sample_n,
                               "oracle/hbond_port.py on the first frames of the device-resident trajectory (rank 0)")

    # the headline end-to-end figure: the faster host source (both legs are in the line)
    e2e = None
    if e2e_f32 or e2e_xtc:
        best = e2e_xtc if (e2e_xtc and (not e2e_f32 or e2e_xtc["value"] > e2e_f32["value"])) else e2e_f32
        e2e = dict(best)
    line = {"metric": "hbond_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64 distance / f32 angle", "data": "synthetic",
            "config": workload_config(system, cfg, args, world),
            "atom_frames_per_s": value * na, "wall_ms_per_step": wall_ms_max / args.steps,
            "n_bonds": int(nb), "n_global_bonds": n_global_keys, "n_part_atoms": int(n_part),
            "clocks": clocks, "gpu_launches": int(launches), "parity": parity, "e2e": e2e, "e2e_f32": e2e_f32,
            "e2e_xtc": e2e_xtc, "parity_xtc": parity_xtc, "e2e_class": e2e_class, "roofline": roofline,
            "cpu_baseline": cpu, "host_numa": numa}
    if world == 1 and args.wire_frames > 0:
        try:
            line["wires"] = wires_leg(args, system, d_xyz, local_rank, peak, bytes_frame)
        except Exception as e:                          # the secondary leg never takes the headline down
            line["wires"] = {"error": repr(e)[:300]}
    if world == 1 and args.c4_frames > 0:
        try:
            line["c4"] = c4_leg(args, local_rank, peak)
        except Exception as e:
            line["c4"] = {"error": repr(e)[:300]}
    if per_rank is not None:
        line["per_rank"] = {"columns": ["pass ms", "frame_fused kernel ms", "e2e_f32 H2D GB/s", "e2e_xtc H2D GB/s"], "ranks": per_rank}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    bad = [n for n, p_ in (("parity", parity), ("parity_xtc", parity_xtc), ("wires.parity", (line.get("wires") or {}).get("parity")),
                           ("c4.parity", (line.get("c4") or {}).get("parity")))
           if p_ is not None and not p_["equal"]]
    if bad:
        sys.stderr.write("bench: PARITY GATE FAILED: %s\n" % ", ".join(bad))
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU; strong: --frames in total, sharded over the GPUs")
    ap.add_argument("--frames", type=int, default=10000, help="frames per GPU (weak) or in total (strong); BASELINE config: 10000")
    ap.add_argument("--scale", type=float, default=1.0, help="atom-count scale of the workload (1.0 = BASELINE config)")
    ap.add_argument("--workload", default="C3", choices=["C3", "C4"],
                    help="C3 = the headline (default); C4 = the 1M-atom config (use with --frames 592 --no-e2e --cpu-frames 2 --c4-frames 0)")
    ap.add_argument("--cpu-frames", type=int, default=1024, help="frames of the CPU baseline / parity gate")
    ap.add_argument("--cpu-frames-xtc", type=int, default=128, help="frames of the XTC leg's parity gate")
    ap.add_argument("--batch-bytes", type=int, default=0)
    ap.add_argument("--xtc-batch-bytes", type=int, default=0, help="batch_bytes of the XTC leg (0: the library default)")
    ap.add_argument("--ring-frames", type=int, default=2048, help="distinct frames per rank in pinned host memory when N > 1")
    ap.add_argument("--wire-frames", type=int, default=2048, help="frames of the secondary water-wire leg (0: skip)")
    ap.add_argument("--c4-frames", type=int, default=592, help="frames of the secondary C4 (1M atoms) leg, N = 1 only (0: skip)")
    ap.add_argument("--native-opt", action="append", default=[], help="name=value passed to hb_set_option (experiments)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-xtc", action="store_true")
    ap.add_argument("--no-class", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the CPUs next to its GPU")
    ap.add_argument("--port-baseline", action="store_true", help="--impl reference: time the port even if oracle/_ref exists")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not os.environ.get("HB_BENCH_ALLOW_SHORT_WARMUP"):
        args.warmup = max(args.warmup, 1)
    _WORKLOAD[0] = args.workload
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

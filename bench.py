#!/usr/bin/env python3
"""bench.py - MPC solves/sec of the batched bipedal MPC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload identical|randomized]

A "step" is one MPC tick (one multiple-shooting SQP iteration: LQ approximation, projected Riccati QP, filter line
search, feedback policy) for every instance of the batch, warm-started from the previous tick.  Default workload =
BASELINE.json configs[1]: Unitree H1, 'trot', horizon 1.0 s at dt 0.01 (N = 100 intervals + 3 event nodes), batch 4096
identical instances.  Prints ONE JSON line (rank 0).
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MPC solves/sec (H1 centroidal, N=100) at batch 4096"
UNIT = "solves/s"
BATCH = 4096
DT, HORIZON = 0.01, 1.0
MPC_DT = 0.02  # mpcDesiredFrequency 50 Hz (task.info:177): the closed loop advances by one MPC period per tick
MODEL = os.path.join(ROOT, "configs", "h1.model")
MODELS = {"h1": MODEL, "g1": os.path.join(ROOT, "configs", "g1.model")}
# algorithmic bytes per node: LQ record 2024 doubles written once + read once, policy record 550 written,
# K + uff (506) read by the forward sweep
LQ_REC, POLICY_REC, FWD_READ = 2024, 550, 506


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload(kind, batch, rank=0, model=None):
    import helpers
    from tools.ingest import read_model
    mdl = read_model(model or MODEL)
    x_init = np.asarray(mdl["initial_state"])
    nx = x_init.shape[0]
    dj = np.asarray(mdl["default_joint_state"])
    ME = 40
    if kind == "identical":
        # configs[1] schedule (events -0.95 + 0.35 j) continued so that the closed loop can advance for many ticks
        et = -0.95 + 0.35 * np.arange(ME - 1)
        ms = np.array([3] + [1, 2] * ((ME - 2) // 2) + [3], dtype=np.int32)[:len(et) + 1]
        ms[-1] = 3
        CMD = np.tile(np.array([0.3, 0.0, 0.0, 0.0]), (batch, 1))
        X0 = np.tile(x_init, (batch, 1))
        tt, ts = helpers.cmd_vel_target(x_init, 0.0, (0.3, 0.0, 0.0, 0.0), 1.0, mdl["com_height"], dj)
        TT = np.tile(tt, (batch, 1))
        TS = np.tile(ts, (batch, 1, 1))
        ET = np.zeros((batch, ME))
        MS = np.zeros((batch, ME + 1), dtype=np.int32)
        ET[:, :len(et)] = et
        MS[:, :len(ms)] = ms
        NE = np.full(batch, len(et), dtype=np.int32)
    else:
        lo = np.array([mdl[f"joint{j}_limits"][0] for j in range(nx - 12)])
        hi = np.array([mdl[f"joint{j}_limits"][1] for j in range(nx - 12)])
        X0, cmd, gait, phase = helpers.randomized_instances(batch, x_init, dj, lo, hi, seed=rank)
        X0[:, 8] = x_init[8] + (X0[:, 8] - 0.93)
        CMD = cmd
        TT = np.zeros((batch, 2))
        TS = np.zeros((batch, 2, nx))
        ET = np.zeros((batch, ME))
        MS = np.zeros((batch, ME + 1), dtype=np.int32)
        NE = np.zeros(batch, dtype=np.int32)
        for b in range(batch):
            TT[b], TS[b] = helpers.cmd_vel_target(X0[b], 0.0, cmd[b], 1.0, mdl["com_height"], dj)
            et, ms = helpers.tiled_schedule(gait[b], phase[b], t_hi=3.6)
            NE[b] = len(et)
            ET[b, :len(et)] = et
            MS[b, :len(ms)] = ms
    return dict(X0=X0, T0=np.zeros(batch), TT=TT, TS=TS, ET=ET, MS=MS, NE=NE, CMD=CMD, mdl=mdl)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def cpu_baseline_run(sample_instances, ticks, threads, kind="identical", march_native=True, model=None):
    """Times the CPU oracle (restatement of the reference path; the reference's OCS2 stack cannot be built here) on host cores."""
    from oracle import pyoracle
    L = None
    if march_native:
        try:
            out = os.path.join(ROOT, "oracle", "liboracle_native.so")
            pyoracle.build(march="native", out=out)
            L = pyoracle.lib(out)
        except Exception:
            L = None
    if L is None:
        pyoracle.build()
        L = pyoracle.lib()
    model = model or MODEL
    w = workload(kind, sample_instances, 0, model)
    ob = pyoracle.OracleBatch(model, sample_instances, L=L)
    for b, o in enumerate(ob.inst):
        o.set_dt_horizon(DT, HORIZON)
        o.set_mode_schedule(w["ET"][b, :w["NE"][b]], w["MS"][b, :w["NE"][b] + 1])
        o.set_target(w["TT"][b], w["TS"][b])
        ob.set_observation(b, 0.0, w["X0"][b])
        ob.set_cmd_vel(b, w["CMD"][b], 1.0)
    ob.run(threads=threads)  # cold tick (untimed)
    ob.run(threads=threads, shift_dt=MPC_DT)  # first warm tick (untimed)
    sec = 0.0
    for _ in range(ticks):
        sec += ob.run(threads=threads, shift_dt=MPC_DT)
    nodes = ob.inst[0].info()["n_nodes"]
    return sample_instances * ticks / sec, sec, nodes


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; OCS2 itself is unbuildable offline) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(threads * 8, 64)
    t0 = time.time()
    val, sec, nodes = cpu_baseline_run(sample, max(args.steps, 1), threads)
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": "BASELINE configs[1]: H1 trot, horizon 1.0 s, dt 0.01 (N=100 + 3 event nodes), identical instances", "nodes": nodes,
                   "note": "CPU oracle port of the reference path (OCS2/Pinocchio/HPIPM are un-vendored and unbuildable offline)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{sample} instances x {args.steps} warm ticks per step-set, one std::thread per core"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="identical", choices=["identical", "randomized"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--robot", default="h1", choices=["h1", "g1"], help="g1 = BASELINE configs[3] (use --batch 8192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default="full", choices=["full", "window", "none"])
    ap.add_argument("--gather-impl", default="nccl", choices=["ce", "nccl"], help="full-policy all-gather: ncclAllGather (default: fastest measured at 8 GPUs) or copy-engine pulls over peer-mapped (symmetric) memory, which keep the SMs free but do not reach NVLink speed")
    ap.add_argument("--opt", action="append", default=[], help="debug option name=value passed to bmpc_debug_set_option (kernel variants; not for reported numbers)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything else that native libraries print there (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line_dict):
        os.write(json_fd, (json.dumps(line_dict) + "\n").encode())

    import torch
    import torch.distributed as dist
    from bipedal_control_b200 import BatchedMpcMrtInterface

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    model = MODELS[args.robot]
    w = workload(args.workload, B, rank, model)
    mpc = BatchedMpcMrtInterface(B, model_file=model, device=local_rank, dt=DT, time_horizon=HORIZON)
    for kv in args.opt:
        name, val = kv.split("=")
        mpc.setOption(name, int(val))
    stream = torch.cuda.ExternalStream(mpc.stream(), device=local_rank)

    # device-resident inputs for the kernel-only metric
    dev = torch.device("cuda", local_rank)
    d_t0 = torch.tensor(w["T0"], device=dev)
    d_x0 = torch.tensor(w["X0"], device=dev)
    d_tt = torch.tensor(w["TT"], device=dev)
    d_ts = torch.tensor(w["TS"], device=dev)
    d_ne = torch.tensor(w["NE"], device=dev, dtype=torch.int32)
    d_et = torch.tensor(w["ET"], device=dev)
    d_ms = torch.tensor(w["MS"], device=dev, dtype=torch.int32)
    torch.cuda.synchronize()

    gather_bufs = {}
    gstream = torch.cuda.Stream(device=dev) if world > 1 else None
    gather_events = []   # events after which the library's policy buffer of that tick may be overwritten (policy buffers are double buffered in the library)
    symm = {}            # copy-engine all-gather: name -> (staging tensor in symmetric memory, rendezvous handle)
    gather_impl = {"kind": "nccl"}

    def policy_fields(v):
        NS, nx, nu = mpc.max_nodes, mpc.nx, mpc.nu
        ptr = (lambda n: getattr(v, n)) if v is not None else (lambda n: 0)
        return (("K", ptr("K"), NS * nu * nx), ("uff", ptr("uff"), NS * nu), ("x", ptr("x"), NS * nx), ("u", ptr("u"), NS * nu), ("t", ptr("times"), NS))

    def setup_copy_engine_gather():
        """Peer-mapped staging buffers (torch symmetric memory: CUDA VMM handles exchanged inside the node).  The all-gather then is 7 pulls per
        rank with copy engines over NVLink (no SMs, so the next tick's kernels keep the whole GPU); falls back to ncclAllGather if unavailable."""
        if world == 1 or args.gather != "full" or args.gather_impl != "ce":
            return
        try:
            import torch.distributed._symmetric_memory as symm_mem
            for name, _, per in policy_fields(None):
                t = symm_mem.empty(B * per, dtype=torch.float64, device=dev)
                symm[name] = (t, symm_mem.rendezvous(t, dist.group.WORLD.group_name))
            gather_impl["kind"] = "copy-engine pulls over peer-mapped (symmetric) memory"
        except Exception as e:   # noqa: BLE001
            symm.clear()
            gather_impl["kind"] = "nccl (symmetric memory unavailable: %s)" % str(e).splitlines()[0][:80]

    def gather_policies():
        """One all-gather of the solved feedback policies per tick (north star).  The gather runs on a side stream and overlaps with the next
        tick's compute: the library double-buffers its policies, so the buffer being gathered is only overwritten two ticks later, and that tick
        first waits until the buffer has been read.  'window' = only the nodes consumers read before the next tick."""
        if world == 1 or args.gather == "none":
            return 0
        v = mpc.getDeviceView()
        nbytes = 0
        done = torch.cuda.Event()
        ready = torch.cuda.Event()
        ready.record(stream)
        gstream.wait_event(ready)
        parity = len(gather_events) & 1
        with torch.cuda.stream(gstream):
            if symm:
                fields = policy_fields(v)
                for name, ptr, per in fields:
                    symm[name][0].copy_(_alias(ptr, B * per, dev))      # stage the shard where the peers can read it (local copy)
                h0 = symm["K"][1]
                h0.barrier(channel=0)                                    # every rank has staged this tick
                for name, ptr, per in fields:
                    key = (name, B * per, parity)
                    if key not in gather_bufs:
                        gather_bufs[key] = torch.empty(world * B * per, device=dev, dtype=torch.float64)
                # pulls on one stream, one peer after the other (measured at 8 GPUs: 25.5 ms / tick; one stream per peer: 40 ms; ncclAllGather: 23.3 ms)
                for step in range(world):
                    r = (rank - step) % world
                    for name, ptr, per in fields:
                        gather_bufs[(name, B * per, parity)].chunk(world)[r].copy_(symm[name][1].get_buffer(r, (B * per,), torch.float64))
                    nbytes += sum(per for _, _, per in fields) * B * 8
                h0.barrier(channel=0)                                    # nobody restages before every pull has finished
                done.record(gstream)
            else:
                for name, ptr, per in policy_fields(v):
                    src = _alias(ptr, B * per, dev)
                    if args.gather == "window":
                        k = 4  # t0 .. t0 + 1/50 s is covered by the first 3 nodes at dt 0.01; 4 for interpolation
                        src = src.view(B, mpc.max_nodes, -1)[:, :k].contiguous()
                    key = (name, src.numel(), parity)
                    if key not in gather_bufs:
                        gather_bufs[key] = torch.empty(world * src.numel(), device=dev, dtype=torch.float64)
                    dist.all_gather_into_tensor(gather_bufs[key], src.reshape(-1))
                    nbytes += src.numel() * 8 * world
                done.record(gstream)
        gather_events.append(done)
        return nbytes

    def wait_for_old_gather():
        # the tick about to start overwrites the policy buffer that was gathered two ticks ago
        if len(gather_events) >= 2:
            stream.wait_event(gather_events[-2])

    d_cmd = torch.tensor(w["CMD"], device=dev)

    def device_step():
        # closed loop, everything resident in HBM: next observation from the current policy, cmd_vel target, one MPC tick
        wait_for_old_gather()
        mpc.shiftObservations(MPC_DT)
        mpc.setTargetsFromCmdVelDevice(d_cmd.data_ptr(), 1.0)
        mpc.advanceMpcAsync()
        return gather_policies()

    perf_host = None

    e2e_state = {}

    def e2e_step():
        # the call sequence a host-side user makes every tick, all buffers in host memory:
        # policy evaluation -> new observation -> targets -> mode schedules -> solve -> performance indices back on the host
        nonlocal perf_host
        if "t" not in e2e_state:
            e2e_state["t"], e2e_state["x"] = mpc.getObservations()
        t_next = e2e_state["t"] + MPC_DT
        x_next, _, _ = mpc.evaluatePolicy(t_next, e2e_state["x"])
        e2e_state["t"], e2e_state["x"] = t_next, x_next
        wait_for_old_gather()
        mpc.setCurrentObservation(t_next, x_next)
        mpc.setTargetsFromCmdVel(w["CMD"], 1.0)
        mpc.setModeSchedule(w["ET"], w["MS"], w["NE"])
        mpc.advanceMpcAsync()
        gather_policies()
        perf_host = mpc.getPerformanceIndices()  # device -> host read of the step's result (PerformanceIndex per instance)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect_phases=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = []
        e0.record(stream)
        t_wall = time.time()
        for _ in range(steps):
            fn()
            if collect_phases:
                mpc.synchronize()
                phases.append(mpc.phaseTimes())
        if gstream is not None:
            stream.wait_stream(gstream)   # the timed region ends when the last policy all-gather has finished, not only the last tick
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        wall = (time.time() - t_wall) * 1e3
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, phases

    # cold start (t0 = 0) + warm-up ticks of the closed loop
    setup_copy_engine_gather()
    mpc.reset()
    mpc.setCurrentObservationDevice(d_t0.data_ptr(), d_x0.data_ptr())
    mpc.setTargetsFromCmdVelDevice(d_cmd.data_ptr(), 1.0)
    mpc.setModeScheduleDevice(d_et.shape[1], d_ne.data_ptr(), d_et.data_ptr(), d_ms.data_ptr())
    mpc.advanceMpcAsync()
    gather_policies()
    for _ in range(max(args.warmup, 3)):
        device_step()
    mpc.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    mpc.enablePhaseTiming(True)
    ms_dev, wall_dev, phases = timed(device_step, args.steps, collect_phases=True)
    launches = (mpc.launchCount() + 2) * args.steps
    mpc.enablePhaseTiming(False)
    # un-instrumented kernel-only timing (no per-step phase readback)
    ms_dev2, _, _ = timed(device_step, args.steps)
    ms_dev = min(ms_dev, ms_dev2)
    for _ in range(2):
        e2e_step()
    ms_e2e, wall_e2e, _ = timed(e2e_step, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    status = mpc.getStatus()
    pol_n = mpc.getPolicy(0, 1, with_gains=False)["n_nodes"][0]
    nodes_stage = int(pol_n) - 1

    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    peaks, peak_kind = measured_peaks()
    ph = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]} if phases else {}
    n_ = mpc.nx
    lq_rec, pol_rec, fwd_rd = 3 * n_ * n_ + n_ * (n_ + 1) + 3 * n_, n_ * n_ + 3 * n_, n_ * n_ + n_   # SURVEY.md 8d general formula (H1: 2024, 550, 506)
    stages_total = B * nodes_stage
    # algorithmic bytes per launch of the three heavy kernels: the LQ kernel writes the LQ record, the Riccati kernel
    # reads it and produces the gains of the policy record, the forward sweep reads K + uff
    kernels = {"k_lq_pack": ("lq", lq_rec), "k_riccati_warp": ("riccati", lq_rec + pol_rec), "k_project": ("projection", lq_rec), "k_forward": ("forward", fwd_rd)}
    rl_all = {}
    for kn, (phase, doubles) in kernels.items():
        ms_k = ph.get(phase, 0.0)
        if ms_k > 0:
            a = stages_total * doubles * 8 / (ms_k * 1e-3) / 1e9
            rl_all[kn] = {"kernel_ms": ms_k, "achieved": a, "frac": a / peaks["hbm_gbs"]}
    dominant = max((k for k in rl_all if k in ("k_lq_pack", "k_riccati_warp", "k_project")), key=lambda k: rl_all[k]["kernel_ms"]) if rl_all else "k_riccati_warp"
    ric_ms = rl_all.get(dominant, {}).get("kernel_ms", 0.0)
    achieved = rl_all.get(dominant, {}).get("achieved", 0.0)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(args.robot, {}).get(dominant)
    tick_bytes = stages_total * (2 * lq_rec + pol_rec + fwd_rd) * 8
    nx = mpc.nx
    h2d = 2 * B * 8 * (1 + nx) + B * 2 * 8 * (1 + nx) + B * (4 + 40 * 8 + 41 * 4)   # evaluatePolicy query + observation, targets, mode schedules
    d2h = B * 8 * 8 + B * 8 * (nx + mpc.nu) + B * 4 + 4 * int(ph.get("linesearch_trials", 1))  # performance indices + evaluatePolicy result
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("BASELINE configs[1]: H1 trot, horizon 1.0 s, dt 0.01 (N=100 intervals + 3 event nodes = 103 stages), batch 4096 identical instances per GPU; closed loop: every step advances t0 by 1/50 s, takes x0 from the previous policy and solves one warm-started tick"
                                if args.workload == "identical" else "BASELINE configs[2]: H1 randomized states / velocity references / gaits (seed = rank), batch 4096 per GPU, warm-started tick"),
                   "robot": args.robot, "batch_per_gpu": B, "stages": nodes_stage, "l2": "per-tick working set (>5 GB of stage records) exceeds the 126 MB L2; no explicit flush",
                   "policy_gather": (args.gather + ", one all-gather per tick on a side stream, overlapped with the next tick (double-buffered policies); " + gather_impl["kind"]) if world > 1 else "n/a (1 GPU)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps,
                "note": "host buffers every step: bmpc_evaluate_policy -> bmpc_set_observations -> bmpc_set_targets_from_cmd_vel -> bmpc_set_mode_schedules -> bmpc_advance -> bmpc_get_performance"},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"] if peaks["hbm_gbs"] else None,
                     "traffic": traffic, "peak_kind": peak_kind, "kernel_ms": ric_ms, "all_kernels": rl_all,
                     "note": "FP64 small-matrix work: the dominant kernel is FP64-pipe/latency bound, not HBM bound (DESIGN.md section 4)",
                     "whole_tick_frac": (tick_bytes / (ms_dev / args.steps * 1e-3) / 1e9) / peaks["hbm_gbs"]},
        "phase_ms": ph,
        "status_nonzero": int(np.count_nonzero(status & ~16)),
    }
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        sample = max(threads * 16, 64)
        val, sec, _ = cpu_baseline_run(sample, 24, threads, kind=args.workload, model=model)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{sample} instances x 24 warm closed-loop ticks of the same workload, one std::thread per core ({sec:.1f} s)"}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_ALIAS_KEEP = []


def _alias(ptr, numel, dev):
    """torch tensor aliasing library-owned device memory (float64)."""
    import torch

    class _Arr:
        pass

    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}
    t = torch.as_tensor(a, device=dev)
    _ALIAS_KEEP.append(a)
    if len(_ALIAS_KEEP) > 64:
        del _ALIAS_KEEP[:32]
    return t


if __name__ == "__main__":
    main()

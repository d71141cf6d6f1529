#!/usr/bin/env python3
"""bench.py - MPC solves/sec of the batched bipedal MPC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload identical|randomized] [--robot h1|g1]

A "step" is one MPC tick (one multiple-shooting SQP iteration: LQ approximation, projected Riccati QP, filter line
search, feedback policy) for every instance of the batch, warm-started from the previous tick.  Default workload =
BASELINE.json configs[1]: Unitree H1, 'trot', horizon 1.0 s at dt 0.01 (N = 100 intervals + 3 event nodes), batch 4096
identical instances.  Prints ONE JSON line (rank 0).

--impl reference times the CPU arm (the oracle port of the reference path: OCS2 itself is unbuildable offline) on
all host cores, with the same `config`, metric and unit.
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

UNIT = "solves/s"
DT, HORIZON = 0.01, 1.0
MPC_DT = 0.02  # mpcDesiredFrequency 50 Hz (task.info:177): the closed loop advances by one MPC period per tick
MODELS = {"h1": os.path.join(ROOT, "configs", "h1.model"), "g1": os.path.join(ROOT, "configs", "g1.model")}
DEFAULT_BATCH = {"h1": 4096, "g1": 8192}
ME = 40                      # capacity of each instance's mode schedule
EVENT_NODES = {"identical": 3, "randomized": 16}   # event nodes inside the horizon: configs[1] has exactly 3 switches on the grid


def metric_name(robot, batch):
    return f"MPC solves/sec ({robot.upper()} centroidal, N=100) at batch {batch}"


def bench_config(robot, workload, batch, world, gather):
    """The `config` object of the JSON line: identical for the GPU arm and the CPU (reference) arm of the same command line."""
    names = {("h1", "identical"): "BASELINE configs[1]: H1 trot, horizon 1.0 s, dt 0.01 (N=100 intervals + 3 event nodes = 103 stages), identical instances",
             ("h1", "randomized"): "BASELINE configs[2]: H1 randomized states / velocity references / gaits incl. FLY (seed = rank), horizon 1.0 s, dt 0.01",
             ("g1", "identical"): "BASELINE configs[3] morphology: G1 trot, horizon 1.0 s, dt 0.01 (N=100 intervals + 3 event nodes = 103 stages), identical instances",
             ("g1", "randomized"): "BASELINE configs[3]: G1 randomized states / velocity references / gaits (seed = rank), horizon 1.0 s, dt 0.01"}
    return {"workload": names[(robot, workload)] + "; closed loop: every step advances t0 by 1/50 s, takes x0 from the previous policy and solves one warm-started tick",
            "robot": robot, "batch_per_gpu": batch, "dt": DT, "horizon": HORIZON, "sqp_iterations": 1,
            "projection": "upstream luConstraintProjection (Eigen::FullPivLU)",
            "l2": "per-tick working set (> 5 GB of LQ / stage records) exceeds the 126 MB L2; no explicit flush",
            "policy_gather": (gather + ": one all-gather of the solved policies per tick, overlapped with the next tick") if world > 1 else "n/a (1 GPU)"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def workload(kind, batch, rank=0, model=None):
    import helpers
    from tools.ingest import read_model
    mdl = read_model(model or MODELS["h1"])
    x_init = np.asarray(mdl["initial_state"])
    nx = x_init.shape[0]
    dj = np.asarray(mdl["default_joint_state"])
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
        if mdl["name"] != "h1":
            X0[:, 12:] = x_init[12:] + 0.5 * (X0[:, 12:] - dj)
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


# ------------------------------------------------------------------------------------------------ CPU arm (the only place that executes oracle/)
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
    model = model or MODELS["h1"]
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
    sample = args.cpu_sample or max(threads * 8, 64)
    steps = max(args.steps, 1)
    t0 = time.time()
    val, sec, nodes = cpu_baseline_run(sample, steps, threads, kind=args.workload, model=MODELS[args.robot])
    B = args.batch or DEFAULT_BATCH[args.robot]
    line = {
        "metric": metric_name(args.robot, B), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "impl": "reference",
        "config": bench_config(args.robot, args.workload, B, args.gpus, args.gather),
        "stages": nodes - 1,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} instances of the same workload x {steps} warm closed-loop ticks, one std::thread per core ({sec:.1f} s); CPU oracle port of the reference path (OCS2 / Pinocchio / HPIPM are un-vendored and unbuildable offline)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))


def cpu_baseline_subprocess(args, steps=24):
    """cpu_baseline of the GPU arm: the CPU arm in its own process, so that the GPU arm's process never maps oracle/ libraries."""
    threads = os.cpu_count() or 1
    sample = max(threads * 16, 64)
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", "1", "--workload", args.workload, "--robot", args.robot,
           "--cpu-sample", str(sample)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env["CUDA_VISIBLE_DEVICES"] = ""
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)["cpu_baseline"]
    raise RuntimeError("CPU arm printed no JSON line: " + out.stderr[-400:])


# ------------------------------------------------------------------------------------------------ DRAM traffic probe (ncu on a tiny separate run)
def traffic_probe(args, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, measured NOW on this box with ncu on a separate 3-tick run of the
    same workload (the timed run itself is never profiled).  Returns bytes or None."""
    out_csv = os.path.join(ROOT, "gpurun_out", "traffic_probe.csv")
    os.makedirs(os.path.dirname(out_csv), exist_ok=True)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", f"regex:{kernel}", "-s", "2", "-c", "1", "--csv", "--log-file", out_csv,
           sys.executable, os.path.abspath(__file__), "--probe", "--workload", args.workload, "--robot", args.robot, "--batch", str(args.batch or DEFAULT_BATCH[args.robot])]
    try:
        subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        total = 0.0
        found = False
        import csv
        with open(out_csv) as fh:
            rows = [r for r in csv.reader(l for l in fh if not l.startswith("==")) if r]
        hdr = rows[0]
        ni, vi, ui = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows[1:]:
            if r[ni].startswith("dram__bytes_"):
                v = float(r[vi].replace(",", ""))
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
                total += v * scale
                found = True
        return total if found else None
    except Exception:   # noqa: BLE001
        return None


def run_probe(args):
    from bipedal_control_b200 import BatchedMpcMrtInterface
    B = args.batch or DEFAULT_BATCH[args.robot]
    model = MODELS[args.robot]
    w = workload(args.workload, B, 0, model)
    mpc = BatchedMpcMrtInterface(B, model_file=model, dt=DT, time_horizon=HORIZON, max_event_nodes=EVENT_NODES[args.workload])
    mpc.setCurrentObservation(w["T0"], w["X0"])
    mpc.setTargetsFromCmdVel(w["CMD"], 1.0)
    mpc.setModeSchedule(w["ET"], w["MS"], w["NE"])
    for _ in range(4):
        mpc.advanceMpc()
        mpc.shiftObservations(MPC_DT)
    mpc.close()


# ------------------------------------------------------------------------------------------------ GPU arm
class Runner:
    """One batched MPC handle + its device-resident closed loop, shared by the headline measurement and the extra configs."""

    def __init__(self, torch, robot, kind, B, rank, local_rank):
        from bipedal_control_b200 import BatchedMpcMrtInterface
        self.torch = torch
        self.B = B
        self.model = MODELS[robot]
        self.w = w = workload(kind, B, rank, self.model)
        self.mpc = BatchedMpcMrtInterface(B, model_file=self.model, device=local_rank, dt=DT, time_horizon=HORIZON, max_event_nodes=EVENT_NODES[kind])
        self.dev = dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.ExternalStream(self.mpc.stream(), device=local_rank)
        self.d_t0 = torch.tensor(w["T0"], device=dev)
        self.d_x0 = torch.tensor(w["X0"], device=dev)
        self.d_ne = torch.tensor(w["NE"], device=dev, dtype=torch.int32)
        self.d_et = torch.tensor(w["ET"], device=dev)
        self.d_ms = torch.tensor(w["MS"], device=dev, dtype=torch.int32)
        self.d_cmd = torch.tensor(w["CMD"], device=dev)
        torch.cuda.synchronize()
        self.e2e_state = {}
        self.perf_host = None
        self.after_tick = lambda: 0     # hook: multi-GPU policy exchange
        self.before_tick = lambda: None

    def cold_start(self):
        m = self.mpc
        m.reset()
        m.setCurrentObservationDevice(self.d_t0.data_ptr(), self.d_x0.data_ptr())
        m.setTargetsFromCmdVelDevice(self.d_cmd.data_ptr(), 1.0)
        m.setModeScheduleDevice(self.d_et.shape[1], self.d_ne.data_ptr(), self.d_et.data_ptr(), self.d_ms.data_ptr())
        m.advanceMpcAsync()
        self.after_tick()

    def device_step(self):
        # closed loop, everything resident in HBM: next observation from the newest policy, cmd_vel target, one MPC tick
        m = self.mpc
        self.before_tick()
        m.shiftObservations(MPC_DT)
        m.setTargetsFromCmdVelDevice(self.d_cmd.data_ptr(), 1.0)
        m.advanceMpcAsync()
        return self.after_tick()

    def e2e_step(self):
        # the call sequence a host-side user makes every tick, all buffers in host memory:
        # policy evaluation -> new observation -> targets -> mode schedules -> solve -> performance indices back on the host
        m, w, s = self.mpc, self.w, self.e2e_state
        if "t" not in s:
            m.synchronize()
            s["t"], s["x"] = m.getObservations()
        t_next = s["t"] + MPC_DT
        m.synchronize()
        x_next, _, _ = m.evaluatePolicy(t_next, s["x"])
        s["t"], s["x"] = t_next, x_next
        self.before_tick()
        m.setCurrentObservation(t_next, x_next)
        m.setTargetsFromCmdVel(w["CMD"], 1.0)
        m.setModeSchedule(w["ET"], w["MS"], w["NE"])
        m.advanceMpc()
        self.after_tick()
        self.perf_host = m.getPerformanceIndices()  # device -> host read of the step's result (PerformanceIndex per instance)

    def e2e_bytes(self):
        B, nx, nu = self.B, self.mpc.nx, self.mpc.nu
        h2d = 2 * B * 8 * (1 + nx) + B * 4 * 8 + B * (4 + ME * 8 + (ME + 1) * 4)   # evaluatePolicy query + observation, velocity commands (targets are built on the device), mode schedules
        d2h = B * 8 * 8 + B * 8 * (nx + nu) + B * 4 + 32                                     # performance indices + evaluatePolicy result + tick counters
        return int(h2d), int(d2h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="identical", choices=["identical", "randomized"])
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default: 4096 for h1, 8192 for g1)")
    ap.add_argument("--robot", default="h1", choices=["h1", "g1"], help="g1 = BASELINE configs[3]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] / configs[3] side measurements and the ncu traffic probe")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--gather", default="full", choices=["full", "window", "none"])
    ap.add_argument("--gather-impl", default="native", choices=["native", "torch"], help="policy all-gather: libbmpc's own NCCL exchange (bmpc_exchange_*) or torch.distributed")
    ap.add_argument("--gather-ctas", type=int, default=16, help="cap on the SMs NCCL may use for the policy all-gather (0: NCCL default)")
    ap.add_argument("--gather-ce", type=int, default=2, help="1: NCCL copy-engine all-gather (NCCL >= 2.28, symmetric windows, no SM at all); 2: symmetric windows with NCCL's SM kernels")
    ap.add_argument("--opt", action="append", default=[], help="debug option name=value passed to bmpc_debug_set_option (not for reported numbers)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.probe:
        return run_probe(args)

    # stdout carries exactly ONE JSON line: everything else that native libraries print there (e.g. "NCCL version ...") goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line_dict):
        os.write(json_fd, (json.dumps(line_dict) + "\n").encode())

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch or DEFAULT_BATCH[args.robot]
    run = Runner(torch, args.robot, args.workload, B, rank, local_rank)
    mpc, dev, stream = run.mpc, run.dev, run.stream
    for kv in args.opt:
        name, val = kv.split("=")
        mpc.setOption(name, int(val))

    # ---- multi-GPU: one all-gather of the solved policies per tick (north star), behind the library's exchange API
    def attach_exchange(r, window):
        from bipedal_control_b200.sharding import PolicyExchange
        r.exchange = PolicyExchange(r.mpc, dist, rank, world, window=window, impl=args.gather_impl, max_ctas=args.gather_ctas, copy_engines=args.gather_ce)
        r.after_tick = r.exchange.after_tick
        r.before_tick = r.exchange.before_tick

    run.exchange = None
    if world > 1 and args.gather != "none":
        attach_exchange(run, args.gather == "window")
    exchange = run.exchange

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(r, fn, steps, collect_phases=False):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = []
        e0.record(r.stream)
        t_wall = time.time()
        for _ in range(steps):
            fn()
            if collect_phases:
                r.mpc.synchronize()
                phases.append(r.mpc.phaseTimes())
        if getattr(r, "exchange", None) is not None:
            r.exchange.join(r.stream)   # the timed region ends when the last policy all-gather has finished, not only the last tick
        e1.record(r.stream)
        barrier()
        ms = e0.elapsed_time(e1)
        wall = (time.time() - t_wall) * 1e3
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, phases

    # cold start (t0 = 0) + warm-up ticks of the closed loop
    W = max(args.warmup, 3)
    run.cold_start()
    for _ in range(W):
        run.device_step()
    mpc.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    mpc.enablePhaseTiming(True)
    ms_ph, _, phases = timed(run, run.device_step, args.steps, collect_phases=True)
    launches_per_tick = mpc.launchCount() + 2      # + k_shift_observations, k_cmd_vel_targets of the closed loop
    mpc.enablePhaseTiming(False)
    # un-instrumented kernel-only timing (no per-step phase readback): this is `value`
    ms_dev, _, _ = timed(run, run.device_step, args.steps)
    for _ in range(2):
        run.e2e_step()
    ms_e2e, _, _ = timed(run, run.e2e_step, args.steps)
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    status = mpc.getStatus()
    stats = mpc.tickStats()
    pol_n = mpc.getPolicy(0, 1, with_gains=False)["n_nodes"][0]
    nodes_stage = int(pol_n) - 1
    n_all = mpc.getPolicy(0, B, with_gains=False)["n_nodes"] if args.workload != "identical" else None
    stages_total = int((n_all - 1).sum()) if n_all is not None else B * nodes_stage

    value = world * B * args.steps / (ms_dev * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    peaks, peak_kind = measured_peaks()
    ph = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]} if phases else {}
    n_ = mpc.nx
    lq_rec, pol_rec, fwd_rd = 3 * n_ * n_ + n_ * (n_ + 1) + 3 * n_, n_ * n_ + 3 * n_, n_ * n_ + n_   # SURVEY.md 8d general formula (H1: 2024, 550, 506)
    # algorithmic bytes per launch of the heavy kernels: the LQ kernel writes the LQ record, the projection reads it, the
    # Riccati kernel reads it and produces the gains of the policy record, the forward sweep reads K + uff
    kernels = {"k_lq_pack": ("lq", lq_rec), "k_riccati_warp": ("riccati", lq_rec + pol_rec), "k_project": ("projection", lq_rec), "k_forward": ("forward", fwd_rd)}
    rl_all = {}
    for kn, (phase, doubles) in kernels.items():
        ms_k = ph.get(phase, 0.0)
        if ms_k > 0:
            a = stages_total * doubles * 8 / (ms_k * 1e-3) / 1e9
            rl_all[kn] = {"kernel_ms": ms_k, "achieved": a, "frac": a / peaks["hbm_gbs"]}
    dominant = max((k for k in rl_all if k in ("k_lq_pack", "k_riccati_warp", "k_project")), key=lambda k: rl_all[k]["kernel_ms"]) if rl_all else "k_lq_pack"
    dom_ms = rl_all.get(dominant, {}).get("kernel_ms", 0.0)
    achieved = rl_all.get(dominant, {}).get("achieved", 0.0)
    tick_bytes = stages_total * (2 * lq_rec + pol_rec + fwd_rd) * 8
    # FP64 roofline of the LQ kernel: exact flop count of the oracle's counting scalar (profiles/r02_flop_count.json), measured DFMA peak
    fp64 = None
    fpath = os.path.join(ROOT, "profiles", "r02_flop_count.json")
    if os.path.exists(fpath) and "k_lq_pack" in rl_all:
        with open(fpath) as fh:
            fc = json.load(fh)
        per_stage = fc.get(args.robot, {}).get("lq_flops_per_stage")
        if per_stage:
            tf = stages_total * per_stage / (rl_all["k_lq_pack"]["kernel_ms"] * 1e-3) / 1e12
            fp64 = {"kernel": "k_lq_pack", "flops_per_stage": per_stage, "achieved_tflops": tf, "peak_tflops": fc.get("dfma_peak_tflops", 35.9), "frac": tf / fc.get("dfma_peak_tflops", 35.9),
                    "note": "flops = the model / LQ arithmetic counted by the oracle's counting scalar for ONE stage with analytic derivatives (not the oracle's dual numbers); peak = DFMA microbenchmark on this pool (profiles/r01_fp64_peak_microbench.txt)"}
    h2d, d2h = run.e2e_bytes()
    traffic = None
    if world == 1 and not args.no_extra:
        traffic = traffic_probe(args, dominant)
    cfg = bench_config(args.robot, args.workload, B, world, args.gather)
    line = {
        "metric": metric_name(args.robot, B), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "stages": nodes_stage,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "note": "host buffers every step: bmpc_evaluate_policy -> bmpc_set_observations -> bmpc_set_targets_from_cmd_vel -> bmpc_set_mode_schedules -> bmpc_advance -> bmpc_get_performance; "
                        "the full PrimalSolution (2.1 GB per 4096-instance tick) stays on the device: host consumers read it per instance with bmpc_get_policy / bmpc_evaluate_policy"},
        "gpu_launches": int(launches_per_tick * args.steps),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"] if peaks["hbm_gbs"] else None,
                     "traffic": traffic, "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch, probed on this box by a separate 4-tick run of the same workload" if traffic else None,
                     "peak_kind": peak_kind, "kernel_ms": dom_ms, "all_kernels": rl_all, "fp64": fp64,
                     "note": "FP64 small-matrix work: the dominant kernel is FP64-issue / latency bound, not HBM bound (DESIGN.md section 4)",
                     "whole_tick_frac": (tick_bytes / (ms_dev / args.steps * 1e-3) / 1e9) / peaks["hbm_gbs"]},
        "phase_ms": ph,
        "linesearch": {"max_trials": stats["max_trials"], "mean_trials": stats["total_trials"] / B},
        "status_nonzero": int(np.count_nonzero(status & ~16)),
    }
    if exchange is not None:
        line["policy_exchange"] = exchange.describe()
        # extra keys at N > 1: the consumed-window exchange (only the nodes consumers read before the next tick) and the
        # randomised distribution of BASELINE configs[4] (seed = rank), both with the same closed loop; shorter runs, every rank takes part
        extra = {}
        try:
            if args.no_extra:
                raise StopIteration
            if args.gather == "full":
                attach_exchange(run, True)
                for _ in range(2):
                    run.device_step()
                msw, _, _ = timed(run, run.device_step, 8)
                extra["consumed_window"] = {"value": world * B * 8 / (msw * 1e-3), "unit": UNIT, "ms_per_step": msw / 8, "steps": 8, "policy_exchange": run.exchange.describe()}
            run.mpc.close()
            r2 = Runner(torch, args.robot, "randomized", B, rank, local_rank)
            r2.exchange = None
            attach_exchange(r2, False)
            r2.cold_start()
            for _ in range(3):
                r2.device_step()
            ms2, _, _ = timed(r2, r2.device_step, 8)
            nz = torch.tensor([int(np.count_nonzero(r2.mpc.getStatus() & ~16))], device=dev)
            dist.all_reduce(nz)
            extra["configs[4] randomized"] = {"value": world * B * 8 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / 8, "steps": 8, "status_nonzero": int(nz.item()),
                                              "workload": bench_config(args.robot, "randomized", B, world, args.gather)["workload"], "policy_exchange": r2.exchange.describe()}
            r2.mpc.close()
        except StopIteration:
            pass
        except Exception as e:   # noqa: BLE001
            extra["error"] = str(e)[:200]
        line["other_configs"] = extra
    if rank == 0 and world == 1 and not args.no_extra:
        # the other single-GPU configs of BASELINE.json, shorter runs (extra keys; the headline stays configs[1])
        extra = {}
        run.mpc.close()
        for key, robot, kind in (("configs[2]", "h1", "randomized"), ("configs[3]", "g1", "randomized")):
            if (robot, kind) == (args.robot, args.workload):
                continue
            try:
                r2 = Runner(torch, robot, kind, DEFAULT_BATCH[robot], 0, local_rank)
                r2.cold_start()
                for _ in range(3):
                    r2.device_step()
                r2.mpc.synchronize()
                ms2, _, _ = timed(r2, r2.device_step, 8)
                st2 = r2.mpc.tickStats()
                extra[key] = {"metric": metric_name(robot, r2.B), "value": r2.B * 8 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / 8, "steps": 8,
                              "mean_linesearch_trials": st2["total_trials"] / r2.B, "status_nonzero": int(np.count_nonzero(r2.mpc.getStatus() & ~16)),
                              "workload": bench_config(robot, kind, r2.B, 1, "full")["workload"]}
                r2.mpc.close()
                del r2
            except Exception as e:   # noqa: BLE001
                extra[key] = {"error": str(e)[:200]}
        line["other_configs"] = extra
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        except Exception as e:   # noqa: BLE001
            line["cpu_baseline"] = {"error": str(e)[:200]}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

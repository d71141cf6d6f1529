#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle (oracle/, FP64).

The reference holds no golden vectors for this path and its OCS2 stack cannot be built or imported here (DESIGN.md
section 2), so these fixtures are *regression pins of the oracle*, not outputs of the reference: they freeze the oracle's
results so that (a) a later edit of the oracle that changes any number fails `-m "not gpu"` tests, and (b) the CUDA path
is compared on the GPU box with numbers that were produced in this container, independently of the oracle build there.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Cases (inputs are stored next to the outputs, so a test needs nothing but the .npz and the model file):
  h1_stance_n20      BASELINE configs[0]: H1 'stance', dt 0.015, horizon 0.3, cold tick + warm tick
  h1_trot_n100       BASELINE configs[1]: H1 'trot', dt 0.01, horizon 1.0 (103 stages), cold tick + warm tick
  h1_random4         BASELINE configs[2] distributions, instances 0, 1, 4, 6 of seed 0 (one per gait), cold tick + warm tick
  g1_trot_n100       BASELINE configs[3]: G1 'trot', dt 0.01, horizon 1.0
  h1_random4_pinv    the h1_random4 instances solved with the selectable Moore-Penrose projection ("projection_mode" 0), first tick
All other fixtures use the default projection = upstream's luConstraintProjection (Eigen::FullPivLU, emulated by the oracle).
Gains are stored for a subset of nodes (first 4, every 10th, last) to keep the files small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle.pyoracle import Oracle, build  # noqa: E402
from tools.ingest import read_model  # noqa: E402


def gain_nodes(n_stages):
    ks = sorted(set([0, 1, 2, 3] + list(range(0, n_stages, 10)) + [n_stages - 1]))
    return np.array([k for k in ks if k < n_stages], dtype=np.int32)


def run_case(model, dt, horizon, et, ms, tt, ts, x0, ticks=2):
    o = Oracle(model)
    o.reset()
    o.set_dt_horizon(dt, horizon)
    o.set_mode_schedule(et, ms)
    o.set_target(tt, ts)
    out = {}
    for tick in range(ticks):
        o.run(0.0, x0)
        so, io = o.solution(), o.info()
        n = len(so["t"])
        gk = gain_nodes(n - 1)
        out[f"t{tick}_times"] = np.asarray(so["t"])
        out[f"t{tick}_events"] = np.asarray(so["events"], dtype=np.int32)
        out[f"t{tick}_x"] = np.asarray(so["x"])
        out[f"t{tick}_u"] = np.asarray(so["u"])
        out[f"t{tick}_uff"] = np.asarray(so["uff"])
        out[f"t{tick}_gain_nodes"] = gk
        out[f"t{tick}_K"] = np.asarray(so["K"])[gk]
        out[f"t{tick}_perf"] = np.concatenate([io["before"], io["after"], [io["step"], io["armijo"]]])
    return out


def main():
    build()
    cases = {}
    h1 = os.path.join(ROOT, "configs", "h1.model")
    g1 = os.path.join(ROOT, "configs", "g1.model")
    m = read_model(h1)
    x0 = np.asarray(m["initial_state"])
    dj = np.asarray(m["default_joint_state"])

    # configs[0]
    c = dict(dt=0.015, horizon=0.3, et=np.array([-1.0, 5.0]), ms=np.array([3, 3, 3], dtype=np.int32), tt=np.array([0.0, 1.0]), ts=np.stack([x0, x0]), x0=x0)
    cases["h1_stance_n20"] = (h1, c)
    # configs[1]
    et, ms = helpers.config2(len(x0), x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0.0, 0.0, 0.0), 1.0, m["com_height"], dj)
    cases["h1_trot_n100"] = (h1, dict(dt=0.01, horizon=1.0, et=et, ms=ms, tt=tt, ts=ts, x0=x0))
    # configs[3]
    mg = read_model(g1)
    xg = np.asarray(mg["initial_state"])
    ttg, tsg = helpers.cmd_vel_target(xg, 0.0, (0.3, 0.0, 0.0, 0.0), 1.0, mg["com_height"], np.asarray(mg["default_joint_state"]))
    cases["g1_trot_n100"] = (g1, dict(dt=0.01, horizon=1.0, et=et, ms=ms, tt=ttg, ts=tsg, x0=xg))

    for name, (model, c) in cases.items():
        out = run_case(model, c["dt"], c["horizon"], c["et"], c["ms"], c["tt"], c["ts"], c["x0"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), dt=c["dt"], horizon=c["horizon"], event_times=c["et"], mode_sequence=c["ms"],
                            target_times=c["tt"], target_states=c["ts"], x0=c["x0"], **out)
        print(name, "nodes", len(out["t0_times"]), "perf after warm tick", out["t1_perf"][3:6])

    # configs[2]: four instances of the seed-0 randomised batch, one per gait (divergent contact modes incl. FLY)
    nx = len(x0)
    lo = np.array([m[f"joint{j}_limits"][0] for j in range(nx - 12)])
    hi = np.array([m[f"joint{j}_limits"][1] for j in range(nx - 12)])
    X0, cmd, gait, phase = helpers.randomized_instances(4096, x0, dj, lo, hi, seed=0)
    picks = [0, 1, 4, 6]   # trot, standing_trot, stance, flying_trot (FLY phases)
    blob = dict(instances=np.array(picks, dtype=np.int32), dt=0.01, horizon=1.0)
    for b in picks:
        et, ms = helpers.tiled_schedule(gait[b], phase[b])
        tt, ts = helpers.cmd_vel_target(X0[b], 0.0, cmd[b], 1.0, m["com_height"], dj)
        out = run_case(h1, 0.01, 1.0, et, ms, tt, ts, X0[b])
        blob.update({f"i{b}_event_times": et, f"i{b}_mode_sequence": ms, f"i{b}_target_times": tt, f"i{b}_target_states": ts, f"i{b}_x0": X0[b],
                     f"i{b}_gait": np.array(str(gait[b]))})
        blob.update({f"i{b}_{k}": v for k, v in out.items()})
        print("h1_random4", b, gait[b], "nodes", len(out["t0_times"]), "step", out["t1_perf"][6])
    np.savez_compressed(os.path.join(HERE, "h1_random4.npz"), **blob)

    # the same four instances with the selectable Moore-Penrose projection (oracle switch / product option "projection_mode" 0): first tick only, no gains
    from oracle import pyoracle
    pyoracle.set_projection_mode(0)
    try:
        blob = dict(instances=np.array(picks, dtype=np.int32), dt=0.01, horizon=1.0)
        for b in picks:
            et, ms = helpers.tiled_schedule(gait[b], phase[b])
            tt, ts = helpers.cmd_vel_target(X0[b], 0.0, cmd[b], 1.0, m["com_height"], dj)
            out = run_case(h1, 0.01, 1.0, et, ms, tt, ts, X0[b], ticks=1)
            blob.update({f"i{b}_event_times": et, f"i{b}_mode_sequence": ms, f"i{b}_target_times": tt, f"i{b}_target_states": ts, f"i{b}_x0": X0[b]})
            blob.update({f"i{b}_{k}": v for k, v in out.items() if not k.endswith("_K") and not k.endswith("gain_nodes")})
        np.savez_compressed(os.path.join(HERE, "h1_random4_pinv.npz"), **blob)
    finally:
        pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)


if __name__ == "__main__":
    main()

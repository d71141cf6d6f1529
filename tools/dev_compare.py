#!/usr/bin/env python3
"""Developer diagnostic (GPU box): stage-by-stage comparison of the CUDA path with the CPU oracle.

Usage: python tools/dev_compare.py [config]   (config in {1, 2, 3}); prints max-abs differences of every intermediate.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bipedal_control_b200 import BatchedMpcMrtInterface  # noqa: E402
from oracle.pyoracle import Oracle, build  # noqa: E402
import helpers  # noqa: E402

MODEL = os.path.join(ROOT, "configs", "h1.model")


def compare(name, a, b, tol=1e-8):
    a, b = np.asarray(a), np.asarray(b)
    err = np.abs(a - b).max() if a.size else 0.0
    scale = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    flag = "" if err <= tol * max(1.0, scale) else "   <<<<<<<<"
    print(f"  {name:28s} max|diff| = {err:10.3e}   (scale {scale:9.3e}){flag}")
    return err


def run(config, ticks=2):
    build()
    o = Oracle(MODEL)
    x_init = o.initial_state()
    nx, nu, nj = o.nx, o.nu, o.nj
    from tools.ingest import read_model
    mdl = read_model(MODEL)
    if config == 1:
        dt, hor = 0.015, 0.3
        et, ms = np.array([-1.0, 5.0]), np.array([3, 3, 3], dtype=np.int32)
        tt, ts = np.array([0.0, 1.0]), np.stack([x_init, x_init])
        x0 = x_init.copy()
    else:
        dt, hor = 0.01, 1.0
        et, ms = helpers.config2(nx, x_init, None, None)
        x0 = x_init.copy()
        if config == 3:
            X, cmd, gait, phase = helpers.randomized_instances(8, x_init, mdl["default_joint_state"], -1.0, 1.0, seed=0)
            x0 = X[3]
            et, ms = helpers.tiled_schedule(gait[3], phase[3])
            print("gait", gait[3], "phase", phase[3])
            tt, ts = helpers.cmd_vel_target(x0, 0.0, cmd[3], 1.0, mdl["com_height"], mdl["default_joint_state"])
        else:
            tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0), 1.0, mdl["com_height"], mdl["default_joint_state"])
    o.set_dt_horizon(dt, hor)
    o.set_mode_schedule(et, ms)
    o.set_target(tt, ts)
    B = 3
    g = BatchedMpcMrtInterface(B, model_file=MODEL, dt=dt, time_horizon=hor)
    g.setCurrentObservation(0.0, x0)
    g.setTargetTrajectories(tt, ts)
    g.setModeSchedule(et, ms)
    for tick in range(ticks):
        print(f"=== config {config} tick {tick}")
        o.run(0.0, x0)
        g.advanceMpc()
        so, io, stp = o.solution(), o.info(), o.step()
        pol = g.getPolicy()
        perf = g.getPerformanceIndices()
        print("  status", g.getStatus(), "launches", g.launchCount(), "oracle info", {k: (v if not isinstance(v, np.ndarray) else np.round(v, 8)) for k, v in io.items()})
        n = pol["n_nodes"][1]
        print("  n_nodes gpu", pol["n_nodes"], "oracle", len(so["t"]))
        if n != len(so["t"]):
            print("  NODE COUNT MISMATCH"); print(pol["t"][1][:n]); print(so["t"]); return
        compare("node times", pol["t"][1][:n], so["t"])
        compare("node events", pol["events"][1][:n], so["events"])
        # LQ records
        rec = g.debugCopy("lq_record", 1)
        prj = g.debugCopy("proj_record", 1)
        worst = dict(A=0, B=0, b=0, q=0, r=0, Px=0, Pe=0, PN=0, K=0)
        for k in range(n - 1):
            L = o.node_lq(k)
            e = helpers.expand_lq_record(rec[k], nj, mdl["total_mass"])
            if L["type"] != 0:
                worst["b"] = max(worst["b"], np.abs(e["b"] - L["b"]).max())
                continue
            worst["A"] = max(worst["A"], np.abs(e["A"] - L["A"]).max())
            worst["B"] = max(worst["B"], np.abs(e["B"] - L["B"]).max())
            worst["b"] = max(worst["b"], np.abs(e["b"] - L["b"]).max())
            worst["q"] = max(worst["q"], np.abs(e["q"] - L["q"]).max())
            worst["r"] = max(worst["r"], np.abs(e["r"] - L["r"]).max())
            P = o.node_projection(k, L["m"])
            nxa = nx - 3
            X = list(range(6)) + list(range(9, nx))
            Pxj = prj[k][:nj * nxa].reshape(nj, nxa)
            Pej = prj[k][nj * nxa:nj * nxa + nj]
            Nn = prj[k][nj * nxa + nj:nj * nxa + nj + nj * 8].reshape(nj, 8)
            mj = int(prj[k][nj * nxa + nj + nj * 8 + 12])
            worst["Px"] = max(worst["Px"], np.abs(Pxj - P["Px"][12:][:, X]).max(), np.abs(P["Px"][12:, 6:9]).max())
            worst["Pe"] = max(worst["Pe"], np.abs(Pej - P["Pe"][12:]).max())
            Pu_j = P["Pu"][12:, :]
            Pu_j = Pu_j[:, np.abs(Pu_j).max(axis=0) > 0] if Pu_j.size else Pu_j
            if mj > 0:   # the bases need not be orthonormal (FullPivLU kernel): compare the orthogonal projectors onto their spans
                qo, _ = np.linalg.qr(Pu_j); qg, _ = np.linalg.qr(Nn[:, :mj])
                worst["PN"] = max(worst["PN"], np.abs(qo @ qo.T - qg @ qg.T).max() if qo.shape[1] == mj else 9.9)
            # raw constraint rows of the contact velocities (upstream stacking order, zero-force rows skipped)
            vel = [i for i in range(L["D"].shape[0]) if np.abs(L["D"][i, :12]).max() == 0.0]
            if e["nrows"] == len(vel):
                worst["CD"] = max(worst.get("CD", 0), np.abs(e["Cv"] - L["C"][vel][:, X]).max(), np.abs(e["Dv"] - L["D"][vel][:, 12:]).max(), np.abs(e["ev"] - L["e"][vel]).max())
            worst["K"] = max(worst["K"], np.abs(pol["K"][1][k] - P["K"]).max())
        for kk, v in worst.items():
            print(f"  LQ worst {kk:3s} {v:10.3e}")
        compare("perf before", perf[1][0:3], io["before"])
        compare("perf after", perf[1][3:6], io["after"])
        compare("step size", perf[1][6], io["step"])
        compare("armijo", perf[1][7], io["armijo"])
        dx = g.debugCopy("dx", 1)[:n]
        du = g.debugCopy("du", 1)[:n - 1]
        compare("dx", dx, stp["dx"])
        compare("du", du, stp["du"])
        compare("x", pol["x"][1][:n], so["x"])
        compare("u", pol["u"][1][:n], so["u"])
        compare("uff", pol["uff"][1][:n], so["uff"], tol=1e-7)
        compare("K", pol["K"][1][:n], so["K"], tol=1e-7)
        compare("instances identical (x)", pol["x"][0], pol["x"][2], tol=0)
        compare("instances identical (K)", pol["K"][0], pol["K"][2], tol=0)
        tq = 0.013
        xq = x0 + 0.01
        xo_, uo_, mo_ = o.evaluate_policy(tq, xq)
        xg, ug, mg = g.evaluatePolicy(tq, xq)
        compare("evaluatePolicy x", xg[1], xo_)
        compare("evaluatePolicy u", ug[1], uo_, tol=1e-7)
        print("  mode", mg[1], mo_)


if __name__ == "__main__":
    cfgs = [int(a) for a in sys.argv[1:]] or [1, 2, 3]
    for c in cfgs:
        run(c)

"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances (FP64 end to end): 1e-8 relative on trajectories / gains / performance indices for identical linearisation
points (far inside the north star's 1e-4 on cost and constraint residuals); closed-loop sequences of several ticks
accumulate rounding through re-linearisation and are held to 1e-6.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = os.path.join(ROOT, "configs", "h1.model")
REL = 1e-8


def _gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from bipedal_control_b200 import BatchedMpcMrtInterface
    return BatchedMpcMrtInterface


def _mdl():
    from tools.ingest import read_model
    return read_model(MODEL)


def _close(a, b, rel=REL, what=""):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, np.abs(b).max())
    err = np.abs(a - b).max()
    assert err <= rel * scale, f"{what}: max|diff| {err:.3e} > {rel:.0e} * {scale:.3e}"


def _compare_tick(g, o, inst, rel=REL):
    pol = g.getPolicy(inst, 1)
    so, io = o.solution(), o.info()
    n = int(pol["n_nodes"][0])
    assert n == len(so["t"])
    _close(pol["t"][0][:n], so["t"], 1e-14, "node times")
    assert list(pol["events"][0][:n]) == list(so["events"])
    perf = g.getPerformanceIndices()[inst]
    _close(perf[0:3], io["before"], rel, "performance before")
    _close(perf[3:6], io["after"], rel, "performance after")
    assert perf[6] == io["step"]
    _close(perf[7], io["armijo"], rel, "armijo descent metric")
    _close(pol["x"][0][:n], so["x"], rel, "x")
    _close(pol["u"][0][:n], so["u"], rel, "u")
    _close(pol["uff"][0][:n], so["uff"], rel, "uff")
    _close(pol["K"][0][:n], so["K"], rel, "K")
    return pol, so


def test_config1_stance_plumbing(oracle_h1):
    """BASELINE configs[0]: H1 'stance', N = 20, dt 0.015, batch 1, cold start then a warm tick."""
    G = _gpu()
    o = oracle_h1
    x0 = o.initial_state()
    o.reset(); o.set_dt_horizon(0.015, 0.3); o.set_mode_schedule([-1.0, 5.0], [3, 3, 3]); o.set_target([0.0, 1.0], [x0, x0])
    g = G(1, model_file=MODEL, dt=0.015, time_horizon=0.3)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories([0.0, 1.0], [x0, x0]); g.setModeSchedule([-1.0, 5.0], [3, 3, 3])
    for _ in range(2):
        o.run(0.0, x0); g.advanceMpc()
        assert g.getStatus()[0] == 0
        pol, so = _compare_tick(g, o, 0)
        assert pol["n_nodes"][0] == 21
    xo, uo, mo = o.evaluate_policy(0.007, x0 + 0.01)
    xg, ug, mg = g.evaluatePolicy(0.007, x0 + 0.01)
    _close(xg[0], xo, REL, "evaluatePolicy x"); _close(ug[0], uo, REL, "evaluatePolicy u"); assert mg[0] == mo == 3
    g.close()


def test_config2_trot_identical_instances(oracle_h1):
    """BASELINE configs[1] at full size: all 4096 outputs bit-identical to each other, instance 0 within tolerance of the oracle."""
    import helpers
    G = _gpu()
    m = _mdl()
    o = oracle_h1
    x0 = o.initial_state()
    et, ms = helpers.config2(o.nx, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_dt_horizon(0.01, 1.0); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    B = 4096
    g = G(B, model_file=MODEL, dt=0.01, time_horizon=1.0)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    for tick in range(2):
        o.run(0.0, x0); g.advanceMpc()
        assert not g.getStatus().any()
        pol, so = _compare_tick(g, o, 0)
        assert pol["n_nodes"][0] == 104   # 100 intervals + 3 event nodes + terminal
    # bit-identical across the batch (checks every wave of CTAs, not only the first 148 x occupancy instances)
    perf = g.getPerformanceIndices()
    assert (perf == perf[0]).all()
    for first in (1, 700, 2047, 4095):
        p = g.getPolicy(first, 1)
        for k in ("x", "u", "uff", "K"):
            assert np.array_equal(p[k][0], pol[k][0]), k
    g.close()


def _randomized_batch(model, B, seed, all_gaits):
    """BASELINE configs[2] / configs[3] distributions for `model` (seeded), as arrays ready for the C ABI."""
    import helpers
    from tools.ingest import read_model
    m = read_model(model)
    nj = m["nj"]
    lo = np.array([m[f"joint{j}_limits"][0] for j in range(nj)]); hi = np.array([m[f"joint{j}_limits"][1] for j in range(nj)])
    X0, cmd, gait, phase = helpers.randomized_instances(B, np.asarray(m["initial_state"]), np.asarray(m["default_joint_state"]), lo, hi, seed=seed)
    if m["name"] != "h1" and nj == 12:   # the helper draws around the H1 stance: re-centre heights / joints on this robot's initial state
        X0[:, 8] = m["initial_state"][8] + (X0[:, 8] - 0.93)
        X0[:, 12:] = np.asarray(m["initial_state"])[12:] + 0.5 * (X0[:, 12:] - np.asarray(m["default_joint_state"]))
    ME = 40
    ET, MS, NE = np.zeros((B, ME)), np.zeros((B, ME + 1), dtype=np.int32), np.zeros(B, dtype=np.int32)
    TT, TS = np.zeros((B, 2)), np.zeros((B, 2, 12 + nj))
    for b in range(B):
        g_ = gait[b] if (all_gaits or b % 4 == 0) else "trot"
        gait[b] = g_
        et, ms = helpers.tiled_schedule(g_, phase[b], t_hi=3.0)
        NE[b] = len(et); ET[b, :len(et)] = et; MS[b, :len(ms)] = ms
        TT[b], TS[b] = helpers.cmd_vel_target(X0[b], 0.0, cmd[b], 1.0, m["com_height"], m["default_joint_state"])
    return m, X0, gait, ET, MS, NE, TT, TS


def _full_size_randomized(model, B, seed, all_gaits, n_check=64):
    """Full-size randomised batch on the GPU; n_check instances (spread over the batch = over every wave of CTAs, every gait represented) are
    solved by the oracle as well (one thread per host core): trajectories, gains, performance indices of two ticks within tolerance, identical
    accepted step sizes, and the same number of line-search failures (alpha_min reached, status bit 16)."""
    from oracle.pyoracle import OracleBatch
    G = _gpu()
    m, X0, gait, ET, MS, NE, TT, TS = _randomized_batch(model, B, seed, all_gaits)
    check = set(np.linspace(0, B - 1, n_check - 12).astype(int).tolist() + [1, 2, 3, 5, 8, 13, B // 2 + 1, B - 2])
    for kind in ("trot", "standing_trot", "flying_trot", "stance"):   # every gait is represented
        check.add(int(np.flatnonzero(gait == kind)[len(check) % 7]))
    check = sorted(check)
    kinds = set(gait[check].tolist())
    assert {"trot", "standing_trot", "flying_trot", "stance"} <= kinds
    assert any(0 in MS[b, :NE[b] + 1] for b in check)          # FLY phases are among the checked instances
    g = G(B, model_file=model, dt=0.01, time_horizon=1.0)
    g.setCurrentObservation(np.zeros(B), X0); g.setTargetTrajectories(TT, TS); g.setModeSchedule(ET, MS, NE)
    ob = OracleBatch(model, len(check))
    for i, b in enumerate(check):
        o = ob.inst[i]
        o.set_dt_horizon(0.01, 1.0); o.set_mode_schedule(ET[b, :NE[b]], MS[b, :NE[b] + 1]); o.set_target(TT[b], TS[b])
        ob.set_observation(i, 0.0, X0[b])
    for tick in range(2):
        g.advanceMpc()
        st = g.getStatus()
        assert not (st & ~16).any()
        ob.run(threads=os.cpu_count() or 1)
        rejected_oracle = 0
        for i, b in enumerate(check):
            _compare_tick(g, ob.inst[i], b, rel=1e-7 if tick else REL)
            rejected_oracle += int(ob.inst[i].info()["step"] == 0.0)
        assert int(np.count_nonzero(st[check] & 16)) == rejected_oracle
        stats = g.tickStats()
        assert stats["failed_instances"] == 0 and stats["max_trials"] >= 1 and stats["total_trials"] >= B
    g.close()


def test_config3_randomized_divergent_modes():
    """BASELINE configs[2] at full size: H1, batch 4096, seed-0 distributions: divergent contact modes incl. FLY, non-zero momentum (rotating
    stance feet: the dependent sixth row of each stance foot is inconsistent, which is where upstream's FullPivLU projection and a pseudo-inverse differ)."""
    _full_size_randomized(MODEL, 4096, 0, all_gaits=True)


def test_closed_loop_shift_and_gait_schedule(h1_model_path):
    """Several closed-loop ticks (t0 += 1/50 s, x0 from the policy) with the library's GaitSchedule bookkeeping, vs the oracle."""
    import helpers
    from oracle.pyoracle import OracleBatch
    G = _gpu()
    m = _mdl()
    B = 3
    g = G(B, model_file=MODEL, dt=0.015, time_horizon=1.0)
    ob = OracleBatch(h1_model_path, B)
    cmds = np.array([[0.3, 0.0, 0.0, 0.0], [0.0, 0.1, 0.0, 0.2], [-0.2, 0.0, 0.0, -0.1]])
    x0 = g.initialState()
    gaits = ["trot", "flying_trot", "standing_trot"]
    for b in range(B):
        modes, times = helpers.GAITS[gaits[b]]
        g.insertGait((modes, times), 1.0, 2.0, instance=b)          # GaitReceiver inserts the new gait at the end of the horizon
        ob.inst[b].gait_insert(modes, times, 1.0, 2.0)
        ob.inst[b].use_gait_schedule()
        ob.set_observation(b, 0.0, x0); ob.set_cmd_vel(b, cmds[b], 1.0)
    g.useGaitSchedule(True)
    t = np.zeros(B); X = np.tile(x0, (B, 1))
    for tick in range(8):
        g.setCurrentObservation(t, X); g.setTargetsFromCmdVel(cmds, 1.0)
        g.advanceMpc()
        ob.run(threads=1, shift_dt=0.0 if tick == 0 else 0.05)
        assert not (g.getStatus() & ~16).any()
        for b in range(B):
            _compare_tick(g, ob.inst[b], b, rel=1e-6)
            eg, mg = g.gaitPeek(b); eo, mo = ob.inst[b].gait_peek()
            np.testing.assert_allclose(eg, eo, atol=1e-12); assert list(mg) == list(mo)
        # next observation: the optimized state 0.05 s ahead (device-side shift == host-side evaluatePolicy == oracle)
        xo, _, _ = g.evaluatePolicy(t + 0.05, X)
        g.shiftObservations(0.05)
        t2, X2 = g.getObservations()
        np.testing.assert_allclose(t2, t + 0.05); np.testing.assert_allclose(X2, xo, atol=1e-14)
        for b in range(B):
            xb, _, _ = ob.inst[b].evaluate_policy(t[b] + 0.05, X[b])
            _close(X2[b], xb, 1e-6, "shifted observation")
        t, X = t2, X2
    # device-resident continuation: observations stay in HBM (bmpc_shift_observations), the gait schedules are tiled on the device from them
    import torch
    d_cmd = torch.tensor(cmds, device="cuda")
    g.setCurrentObservation(t, X)
    for tick in range(3):
        g.setTargetsFromCmdVelDevice(d_cmd.data_ptr(), 1.0)
        g.advanceMpc()
        ob.run(threads=1, shift_dt=0.05)
        for b in range(B):
            _compare_tick(g, ob.inst[b], b, rel=1e-6)
            eg, mg = g.gaitPeek(b); eo, mo = ob.inst[b].gait_peek()
            np.testing.assert_allclose(eg, eo, atol=1e-12); assert list(mg) == list(mo)
        g.shiftObservations(0.05)
    g.close()


def test_device_resident_inputs_match_host_inputs():
    """The *_device entry points (inputs already in HBM) give bit-identical results to the host entry points."""
    import torch
    import helpers
    G = _gpu()
    m = _mdl()
    B = 64
    x0 = np.asarray(m["initial_state"])
    et, ms = helpers.config2(22, x0, None, None)
    cmd = np.tile([0.3, 0.0, 0.0, 0.1], (B, 1))
    ga = G(B, model_file=MODEL, dt=0.01, time_horizon=0.5); gb = G(B, model_file=MODEL, dt=0.01, time_horizon=0.5)
    ga.setCurrentObservation(0.0, x0); ga.setTargetsFromCmdVel(cmd, 1.0); ga.setModeSchedule(et, ms); ga.advanceMpc()
    dev = torch.device("cuda", 0)
    d_t = torch.zeros(B, dtype=torch.float64, device=dev); d_x = torch.tensor(np.tile(x0, (B, 1)), device=dev); d_c = torch.tensor(cmd, device=dev)
    d_ne = torch.full((B,), len(et), dtype=torch.int32, device=dev)
    d_et = torch.tensor(np.tile(et, (B, 1)), device=dev); d_ms = torch.tensor(np.tile(ms, (B, 1)), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    gb.setCurrentObservationDevice(d_t.data_ptr(), d_x.data_ptr()); gb.setTargetsFromCmdVelDevice(d_c.data_ptr(), 1.0)
    gb.setModeScheduleDevice(len(et), d_ne.data_ptr(), d_et.data_ptr(), d_ms.data_ptr()); gb.advanceMpc()
    pa, pb = ga.getPolicy(), gb.getPolicy()
    for k in ("x", "u", "uff", "K", "t"):
        assert np.array_equal(pa[k], pb[k]), k
    ga.close(); gb.close()


def test_error_paths():
    G = _gpu()
    from bipedal_control_b200 import BmpcError
    with pytest.raises(BmpcError):
        G(4, model_file="/nonexistent.model")
    g = G(2, model_file=MODEL, dt=0.01, time_horizon=0.2)
    with pytest.raises(BmpcError):
        g.advanceMpc()                       # targets / schedules not set
    with pytest.raises(BmpcError):
        g.getPolicy()                        # no solution yet
    x0 = g.initialState()
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories([0.0], [x0])
    # a swing phase without a lift-off time: the reference throws (SwingTrajectoryPlanner.cpp:191-212); here a status bit is raised
    g.setModeSchedule([0.05], [1, 3])
    with pytest.raises(BmpcError) as ei:
        g.advanceMpc()
    assert ei.value.code == -1           # BMPC_ERR_INVALID; the tick itself completed and every getter keeps working
    assert (g.getStatus() & 4).all()
    assert g.tickStats()["status_or"] & 4
    g.close()


@pytest.mark.parametrize("n_intervals", [1, 2, 3, 4, 7])
def test_short_and_ragged_horizons(oracle_h1, n_intervals):
    """Horizons of 1..7 intervals (stage counts around the 3-stages-per-warp packing of the LQ kernel, the single-stage Riccati sweep) with
    an event inside the horizon where it fits: every packing remainder and the event-node paths, against the oracle."""
    import helpers
    G = _gpu()
    m = _mdl()
    o = oracle_h1
    x0 = o.initial_state()
    dt = 0.01
    hor = n_intervals * dt
    # a switch (LF -> RF) on the third grid point: inside the horizon (one event node) for n_intervals >= 3, at / beyond tf otherwise
    et = np.array([-0.5, 0.02, 0.37])
    ms = np.array([3, 1, 2, 3], dtype=np.int32)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.2, 0.1, 0.0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_dt_horizon(dt, hor); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    g = G(5, model_file=MODEL, dt=dt, time_horizon=hor)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    for tick in range(2):
        # the warm tick starts from a perturbed observation: on a converged 1-interval problem the accept / reject decision of the line
        # search would otherwise be taken on rounding noise
        xo = x0 if tick == 0 else x0 + 0.01 * np.cos(np.arange(len(x0)))
        if tick == 1: g.setCurrentObservation(0.0, xo)
        o.run(0.0, xo); g.advanceMpc()
        assert not g.getStatus().any()
        pol, so = _compare_tick(g, o, 3)
        assert pol["n_nodes"][0] == n_intervals + 1 + (1 if n_intervals >= 3 else 0)
    g.close()


def test_line_search_rejects_and_halves(oracle_h1):
    """A far-off initial state forces alpha < 1 for the first tick; the accepted step size must match the oracle's."""
    import helpers
    G = _gpu()
    m = _mdl()
    o = oracle_h1
    x0 = o.initial_state().copy()
    x0[8] -= 0.25; x0[10] += 0.5; x0[0:3] = [1.5, -1.0, 0.8]; x0[14] -= 0.6
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(o.initial_state(), 0.0, (0.5, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_dt_horizon(0.01, 0.6); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    g = G(2, model_file=MODEL, dt=0.01, time_horizon=0.6)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    steps = []
    for _ in range(3):
        o.run(0.0, x0); g.advanceMpc()
        perf = g.getPerformanceIndices()[0]
        assert perf[6] == o.info()["step"]
        steps.append(perf[6])
        _close(perf[3:6], o.info()["after"], 1e-7, "performance after")
    g.close()
    o.reset()


def test_config4_g1_second_morphology():
    """BASELINE configs[3] at full size: Unitree G1 (12 leg joints, nx = nu = 24), N = 100, batch 8192: authored config (configs/g1); three
    quarters of the instances trot, the rest draw from all gaits."""
    G = _gpu()
    g = G(2, model_file=os.path.join(ROOT, "configs", "g1.model"))
    assert g.nx == 24 and g.nu == 24
    g.close()
    _full_size_randomized(os.path.join(ROOT, "configs", "g1.model"), 8192, 4, all_gaits=False)


def test_two_sqp_iterations_and_reset(oracle_h1):
    """sqpIteration = 2 (task.info:70 allows any count): both iterations re-linearise on the GPU exactly as the oracle does; reset drops the warm start."""
    import helpers
    G = _gpu()
    m = _mdl()
    o = oracle_h1
    x0 = o.initial_state().copy()
    x0[0] = 0.2; x0[7] = 0.05
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0.1, 0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_sqp_iterations(2); o.set_dt_horizon(0.01, 0.5); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    g = G(5, model_file=MODEL, dt=0.01, time_horizon=0.5, sqp_iterations=2)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    try:
        for _ in range(2):
            o.run(0.0, x0); g.advanceMpc()
            pol = g.getPolicy(2, 1); so = o.solution(); n = int(pol["n_nodes"][0])
            _close(pol["x"][0][:n], so["x"], 1e-7, "x after two iterations")
            _close(pol["u"][0][:n], so["u"], 1e-7, "u after two iterations")
            _close(pol["K"][0][:n], so["K"], 1e-7, "K after two iterations")
            _close(g.getPerformanceIndices()[2][3:6], o.info()["after"], 1e-7, "performance after")
        first = g.getPolicy(0, 1)["x"][0].copy()
        g.reset(); o.reset()
        o.run(0.0, x0); g.advanceMpc()   # cold start again: must reproduce a cold-start solve, not the warm one
        pol = g.getPolicy(0, 1); so = o.solution(); n = int(pol["n_nodes"][0])
        _close(pol["x"][0][:n], so["x"], 1e-7, "x after reset")
        assert np.abs(pol["x"][0] - first).max() > 1e-6
    finally:
        o.set_sqp_iterations(1); o.reset()
        g.close()


def test_two_handles_with_different_robots_coexist():
    """H1 and G1 handles in one process: the model constants are re-uploaded per tick, results must not interfere."""
    import helpers
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"])
    et, ms = helpers.config2(22, x0, None, None)
    gh = G(3, model_file=MODEL, dt=0.01, time_horizon=0.3)
    gg = G(3, model_file=os.path.join(ROOT, "configs", "g1.model"), dt=0.01, time_horizon=0.3)
    xg = gg.initialState()
    gh.setCurrentObservation(0.0, x0); gh.setTargetTrajectories([0.0], [x0]); gh.setModeSchedule(et, ms)
    gg.setCurrentObservation(0.0, xg); gg.setTargetTrajectories([0.0], [xg]); gg.setModeSchedule(et, ms)
    gh.advanceMpc(); ref = gh.getPolicy(0, 1)["K"].copy()
    gg.advanceMpc()
    gh.reset(); gh.advanceMpc()
    assert np.array_equal(gh.getPolicy(0, 1)["K"], ref)
    assert not gg.getStatus().any() and not gh.getStatus().any()
    gh.close(); gg.close()


def test_moore_penrose_option_matches_oracle_variant(h1_model_path):
    """"projection_mode" 0 (Householder QR on per-foot compressed rows) against the oracle switched to its Moore-Penrose projection, on
    randomised instances whose stance feet rotate (where the two projections are different QPs)."""
    from oracle import pyoracle
    from oracle.pyoracle import Oracle
    G = _gpu()
    B = 16
    m, X0, gait, ET, MS, NE, TT, TS = _randomized_batch(MODEL, B, 7, True)
    g = G(B, model_file=MODEL, dt=0.01, time_horizon=1.0)
    g.setOption("projection_mode", 0)
    g.setCurrentObservation(np.zeros(B), X0); g.setTargetTrajectories(TT, TS); g.setModeSchedule(ET, MS, NE)
    pyoracle.set_projection_mode(0)
    try:
        oracles = {}
        for b in (0, 3, 9, 15):
            o = Oracle(h1_model_path)
            o.set_dt_horizon(0.01, 1.0); o.set_mode_schedule(ET[b, :NE[b]], MS[b, :NE[b] + 1]); o.set_target(TT[b], TS[b])
            oracles[b] = o
        for tick in range(2):
            g.advanceMpc()
            assert not (g.getStatus() & ~16).any()
            for b, o in oracles.items():
                o.run(0.0, X0[b])
                _compare_tick(g, o, b, rel=1e-7 if tick else REL)
    finally:
        pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)
    g.close()


def test_numerical_failure_is_contained_and_reported():
    """A non-finite observation in one instance: the tick completes, bmpc_advance returns BMPC_ERR_NUMERIC, that instance is flagged, stores
    nothing non-finite (it keeps its previous policy; none on a cold start) and recovers on the next tick; the other instances are bit-identical to a clean run."""
    import helpers
    from bipedal_control_b200 import BmpcError
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"])
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
    B, bad = 6, 4
    clean = G(B, model_file=MODEL, dt=0.01, time_horizon=0.5); dirty = G(B, model_file=MODEL, dt=0.01, time_horizon=0.5)
    for g in (clean, dirty):
        g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    clean.advanceMpc(); dirty.advanceMpc()                 # tick 1: both fine
    X = np.tile(x0, (B, 1)); X[:, 6] += 0.01
    Xbad = X.copy(); Xbad[bad, 3] = np.nan
    clean.setCurrentObservation(0.0, X); dirty.setCurrentObservation(0.0, Xbad)
    before = dirty.getPolicy(bad, 1)
    clean.advanceMpc()
    with pytest.raises(BmpcError) as ei:
        dirty.advanceMpc()                                  # tick 2: instance `bad` fails
    assert ei.value.code == -3
    st = dirty.getStatus()
    assert st[bad] & 8 and not (np.delete(st, bad) & ~16).any()
    assert dirty.tickStats()["failed_instances"] == 1
    pc, pd = clean.getPolicy(), dirty.getPolicy()
    for k in ("x", "u", "uff", "K", "t"):
        assert np.array_equal(np.delete(pc[k], bad, axis=0), np.delete(pd[k], bad, axis=0)), k
        assert np.isfinite(pd[k][bad]).all()
        assert np.array_equal(pd[k][bad], before[k][0]), k      # the failed instance kept its previous policy
    clean.setCurrentObservation(0.0, X); dirty.setCurrentObservation(0.0, X)
    clean.advanceMpc(); dirty.advanceMpc()                 # tick 3: the instance solves again from valid data
    assert not (dirty.getStatus() & ~16).any()
    assert np.isfinite(dirty.getPolicy(bad, 1)["K"]).all()
    # a failure on the very first tick leaves the instance without a policy; it cold-starts on the next one
    g = G(3, model_file=MODEL, dt=0.01, time_horizon=0.3)
    Xn = np.tile(x0, (3, 1)); Xn[1, 12] = np.inf
    g.setCurrentObservation(0.0, Xn); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    with pytest.raises(BmpcError):
        g.advanceMpc()
    assert list(g.getPolicy(0, 3, with_gains=False)["n_nodes"] > 0) == [True, False, True]
    g.setCurrentObservation(0.0, np.tile(x0, (3, 1)))
    g.advanceMpc()
    fresh = G(1, model_file=MODEL, dt=0.01, time_horizon=0.3)
    fresh.setCurrentObservation(0.0, x0); fresh.setTargetTrajectories(tt, ts); fresh.setModeSchedule(et, ms); fresh.advanceMpc()
    p, pf = g.getPolicy(1, 1), fresh.getPolicy(0, 1)
    assert not g.getStatus().any()
    for k in ("x", "u", "uff", "K"):
        assert np.array_equal(p[k][0], pf[k][0]), k          # instance 1 solved a cold-start tick
    g.close(); fresh.close(); clean.close(); dirty.close()


def test_async_tick_publishes_on_completion_and_getters_do_not_tear():
    """MRT semantics (BipedalController.cpp:191-200 vs :332-351): while a tick is in flight the getters keep serving the previous policy; a second
    thread hammering evaluatePolicy / getPolicy during a closed loop only ever sees complete policies of one tick."""
    import threading
    import helpers
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"])
    et, ms = helpers.config2(22, x0, None, None)
    B = 512
    cmd = np.tile([0.3, 0.0, 0.0, 0.0], (B, 1))

    def make():
        g = G(B, model_file=MODEL, dt=0.01, time_horizon=1.0)
        g.setCurrentObservation(0.0, x0); g.setTargetsFromCmdVel(cmd, 1.0); g.setModeSchedule(et, ms)
        return g
    # serial reference run: policy of instance 3 after every tick
    g = make()
    ref = []
    for tick in range(6):
        g.advanceMpc()
        p = g.getPolicy(3, 1, with_gains=True)
        ref.append({k: p[k][0].copy() for k in ("t", "x", "u", "uff", "K")})
        g.shiftObservations(0.02)
    g.close()
    g = make()
    g.advanceMpc()
    first = g.getPolicy(3, 1)
    g.shiftObservations(0.02)
    g.advanceMpcAsync()
    during = g.getPolicy(3, 1)            # the tick is (almost certainly) still running: previous policy, complete
    assert any(np.array_equal(during["x"][0], r["x"]) and np.array_equal(during["K"][0], r["K"]) for r in ref[:2])
    g.synchronize()
    assert not g.poll()
    after = g.getPolicy(3, 1)
    assert np.array_equal(after["x"][0], ref[1]["x"]) and not np.array_equal(after["x"][0], first["x"][0])
    # reader thread against the MPC thread
    stop = threading.Event(); seen = []; errors = []

    def reader():
        try:
            while not stop.is_set():
                p = g.getPolicy(3, 1)
                hits = [i for i, r in enumerate(ref) if np.array_equal(p["t"][0], r["t"])]
                if len(hits) != 1:
                    errors.append("time grid of no tick"); return
                r = ref[hits[0]]
                if not all(np.array_equal(p[k][0], r[k]) for k in ("x", "u", "uff", "K")):
                    errors.append(f"torn policy at tick {hits[0]}"); return
                seen.append(hits[0])
                xo, uo, mo = g.evaluatePolicy(r["t"][0] + 0.004, x0)
                if not (np.isfinite(xo).all() and np.isfinite(uo).all()):
                    errors.append("non-finite evaluatePolicy"); return
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = threading.Thread(target=reader); th.start()
    for tick in range(2, 6):
        g.shiftObservations(0.02)
        g.advanceMpcAsync()
        while g.poll():
            pass
    stop.set(); th.join(timeout=30)
    assert not errors, errors
    assert len(seen) > 0 and seen == sorted(seen)      # complete policies only, never an older one after a newer one
    final = g.getPolicy(3, 1)
    assert np.array_equal(final["K"][0], ref[5]["K"])
    g.close()


def test_reset_of_one_instance():
    """bmpc_reset(h, instance): that instance cold-starts on the next tick (and has no policy in between), the others continue warm."""
    import helpers
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"]).copy(); x0[0] = 0.15
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    g = G(4, model_file=MODEL, dt=0.01, time_horizon=0.4)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    g.advanceMpc(); cold = g.getPolicy(0, 1)["x"][0].copy()
    g.advanceMpc(); warm2 = g.getPolicy(0, 1)["x"][0].copy()
    assert np.abs(cold - warm2).max() > 1e-7
    g.reset(2)
    assert list(g.getPolicy(0, 4, with_gains=False)["n_nodes"] > 0) == [True, True, False, True]
    xo, uo, _ = g.evaluatePolicy(0.0, x0)
    assert np.array_equal(xo[2], x0) and not uo[2].any() and uo[1].any()
    g.advanceMpc()
    p = g.getPolicy(0, 4)
    assert np.array_equal(p["x"][2], cold)                          # cold start again
    assert np.array_equal(p["x"][0], p["x"][1]) and np.array_equal(p["x"][0], p["x"][3]) and not np.array_equal(p["x"][0], cold)
    g.close()


def test_time_grid_capacity_is_reported():
    """More event nodes inside the horizon than max_event_nodes: status bit 32 and BMPC_ERR_CAPACITY instead of a silently distorted grid; a
    larger max_event_nodes solves the same problem."""
    import helpers
    from bipedal_control_b200 import BmpcError
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"])
    et = np.array([-0.5, 0.033, 0.071, 0.112, 0.155, 0.9])     # four switches off the grid inside a 0.2 s horizon: 8 extra nodes
    ms = np.array([3, 1, 3, 2, 3, 1, 3], dtype=np.int32)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.1, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
    from oracle import pyoracle
    nodes_expected = len(pyoracle.time_discretization(0.0, 0.2, 0.01, et)[0])     # the grid restarts at every event: 20 intervals become 20 + 8 nodes
    assert nodes_expected > 21 + 4
    g = G(2, model_file=MODEL, dt=0.01, time_horizon=0.2, max_event_nodes=4)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    with pytest.raises(BmpcError) as ei:
        g.advanceMpc()
    assert ei.value.code == -4 and (g.getStatus() & 32).all()
    g.close()
    g = G(2, model_file=MODEL, dt=0.01, time_horizon=0.2, max_event_nodes=8)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    g.advanceMpc()
    assert not g.getStatus().any() and g.getPolicy(0, 1, with_gains=False)["n_nodes"][0] == nodes_expected
    g.close()


def test_feedback_policy_rollout_matches_oracle(oracle_h1):
    """the device-side closed-loop rollout between MPC ticks (MRT_BASE::rolloutPolicy, Dormand-Prince 5(4) with odeint's step control,
    sub-intervals split at mode switches) against the oracle's restatement: from a perturbed observation, over 8 MRT periods, and across an event."""
    import helpers
    G = _gpu()
    m = _mdl()
    o = oracle_h1
    x0 = o.initial_state()
    et, ms = helpers.config2(o.nx, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_dt_horizon(0.01, 1.0); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    B = 5
    g = G(B, model_file=MODEL, dt=0.01, time_horizon=1.0)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    for _ in range(2):
        o.run(0.0, x0); g.advanceMpc()
    # (1) perturbed observations at t = 0, 8 MRT periods of 2.5 ms
    X = np.tile(x0, (B, 1)) + 0.01 * np.cos(np.arange(B * len(x0))).reshape(B, -1)
    g.setCurrentObservation(np.zeros(B), X)
    g.rolloutObservations(0.02, 8)
    t2, X2 = g.getObservations()
    np.testing.assert_allclose(t2, 0.02)
    for b in range(B):
        xr, steps = o.rollout_policy(0.0, X[b], 0.02, 8)
        assert steps >= 8
        _close(X2[b], xr, 1e-7, "rollout state")
    assert np.abs(X2 - X).max() > 1e-3 and not (g.getStatus() & 8).any()
    # (2) one period across the mode switch at t = 0.10 (two sub-intervals, the second one started weakEpsilon late)
    xs, _, _ = g.evaluatePolicy(np.full(B, 0.09), X)
    g.setCurrentObservation(np.full(B, 0.09), xs)
    g.rolloutObservations(0.02, 1)
    t3, X3 = g.getObservations()
    np.testing.assert_allclose(t3, 0.11)
    for b in (0, B - 1):
        xr, steps = o.rollout_policy(0.09, xs[b], 0.02, 1)
        assert steps >= 2
        _close(X3[b], xr, 1e-7, "rollout state across an event")
    # (3) the rollout differs from the perfect-model shift by the discretisation error of the plan, not by more
    g.setCurrentObservation(np.zeros(B), np.tile(x0, (B, 1)))
    g.rolloutObservations(0.02, 8)
    _, Xr = g.getObservations()
    xe, _, _ = g.evaluatePolicy(np.full(B, 0.02), np.tile(x0, (B, 1)))
    assert 1e-6 < np.abs(Xr - xe).max() < 2e-2
    g.close()
    o.reset()


def test_native_policy_exchange_single_rank():
    """bmpc_exchange_* (NCCL bound at run time, one all-gather of the policy slab per tick) with a one-rank communicator: the gathered slab is the
    policy the getters return, for two consecutive ticks (double-buffered slabs); with and without an SM cap."""
    import torch
    import torch.distributed as dist
    import helpers
    G = _gpu()
    m = _mdl()
    x0 = np.asarray(m["initial_state"])
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
    own_pg = not dist.is_initialized()
    if own_pg:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{33500 + os.getpid() % 2000}", rank=0, world_size=1)
    try:
        for max_ctas, mode in ((0, 0), (8, 0), (0, 1), (16, 2)):   # plain / SM cap / copy engines / symmetric windows (the latter two fall back to plain on old NCCL)
            B = 6
            g = G(B, model_file=MODEL, dt=0.01, time_horizon=0.3, max_event_nodes=3)
            g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
            g.exchangeInit(dist, 0, 1, max_ctas=max_ctas, copy_engines=mode)
            for tick in range(3):
                g.advanceMpcAsync()
                g.exchangeStart()
                g.exchangeWait()
                g.synchronize()
                ptr, nbytes, nranks, ce = g.exchangeView()
                assert nranks == 1 and ptr and nbytes == g.getDeviceView().slab_bytes

                class _Arr:
                    pass
                a = _Arr()
                a.__cuda_array_interface__ = {"shape": (nbytes // 8,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}
                slab = torch.as_tensor(a, device="cuda").cpu().numpy()
                pol = g.getPolicy()
                nK = B * g.max_nodes * g.nu * g.nx
                assert np.array_equal(slab[:nK].reshape(pol["K"].shape), pol["K"])
                nU = B * g.max_nodes * g.nu
                assert np.array_equal(slab[nK:nK + nU].reshape(pol["uff"].shape), pol["uff"])
            g.exchangeDestroy()
            g.close()
    finally:
        if own_pg:
            dist.destroy_process_group()


def test_position_error_gain_variant(tmp_path):
    """positionErrorGain != 0 (legged_hunter_config/config/task/task.info:12 uses 20.0; BipedalRobotInterface.cpp:350-359,
    BipedalRobotPreComputation.cpp:71-80): the z rows of the zero-velocity / normal-velocity constraints also see gain * (p_z - z_ref), which puts
    the base height into the constraint Jacobian.  H1 with the gain switched on, trot with swing phases, against the oracle."""
    import helpers
    from oracle.pyoracle import Oracle
    G = _gpu()
    model = str(tmp_path / "h1_gain20.model")
    txt = open(MODEL).read()
    assert "position_error_gain d 1 0.0" in txt
    open(model, "w").write(txt.replace("position_error_gain d 1 0.0", "position_error_gain d 1 20.0", 1))
    m = _mdl()
    o = Oracle(model)
    x0 = o.initial_state().copy()
    x0[8] += 0.01; x0[0] = 0.1; x0[13] += 0.05
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0.05, 0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    o.set_dt_horizon(0.01, 0.6); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    g = G(3, model_file=model, dt=0.01, time_horizon=0.6)
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    for tick in range(3):
        o.run(0.0, x0); g.advanceMpc()
        assert not (g.getStatus() & ~16).any()
        _compare_tick(g, o, 1, rel=1e-7 if tick else REL)
    # the gain really changes the problem: base-height column of the gains differs from the gain-free solve
    g0 = G(1, model_file=MODEL, dt=0.01, time_horizon=0.6)
    g0.setCurrentObservation(0.0, x0); g0.setTargetTrajectories(tt, ts); g0.setModeSchedule(et, ms); g0.advanceMpc()
    g.reset(); g.advanceMpc()
    assert np.abs(g.getPolicy(0, 1)["K"][0][:, :, 8] - g0.getPolicy(0, 1)["K"][0][:, :, 8]).max() > 1e-3
    from bipedal_control_b200 import BmpcError
    with pytest.raises(BmpcError):
        g.setOption("projection_mode", 0)     # the Moore-Penrose option does not carry the position rows
    g.close(); g0.close()

"""Known-answer tests that pin the CPU oracle.

The reference ships no tests and its solver stack cannot be built offline (parity unpinned), so every check here is an
independent numpy computation: finite differences, brute-force momentum sums, dense KKT solves, closed-form values.
"""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import Oracle


@pytest.fixture(scope="module")
def o(h1_model_path):
    pyoracle.build()
    return Oracle(h1_model_path)


def fd_jac(fun, z, eps=1e-6):
    f0 = np.asarray(fun(z)).reshape(-1)
    J = np.zeros((f0.size, z.size))
    for i in range(z.size):
        zp, zm = z.copy(), z.copy()
        zp[i] += eps
        zm[i] -= eps
        J[:, i] = (np.asarray(fun(zp)).reshape(-1) - np.asarray(fun(zm)).reshape(-1)) / (2 * eps)
    return J


# ---------------------------------------------------------------- E-1 model ingestion
def test_model_masses(o, h1_model_path):
    from tools.ingest import read_model
    m = read_model(h1_model_path)
    assert abs(o.total_mass - 51.641) < 1e-9          # sum of the URDF inertials incl. 4 x 0.01 kg soles
    assert abs(m["base_mass"] - 29.955) < 1e-9        # pelvis + torso + arms lumped (5.39 + 17.789 + 2 * 3.388)
    assert m["nj"] == 10 and m["nc"] == 4
    np.testing.assert_allclose(m["contact0_offset"], [0.19, 0.0, -0.06])
    np.testing.assert_allclose(m["contact1_offset"], [-0.1, 0.0, -0.06])


# ---------------------------------------------------------------- E-2 kinematics / dynamics identities
def test_contact_height_at_initial_state(o):
    x0 = o.initial_state()
    u = np.zeros(o.nu)
    _, pos, _ = o.flow_map(x0, u)
    expect = 0.93 - 0.1742 - 0.4 * np.cos(0.5) - 0.4 * np.cos(0.5) - 0.06
    np.testing.assert_allclose(pos[:, 2], expect, atol=1e-12)


def test_weight_compensation_gives_zero_linear_momentum_rate(o):
    x0 = o.initial_state()
    u = np.zeros(o.nu)
    u[2:12:3] = o.total_mass * 9.81 / 4
    f, _, vel = o.flow_map(x0, u)
    np.testing.assert_allclose(f[0:3], 0.0, atol=1e-12)
    np.testing.assert_allclose(f[6:], 0.0, atol=1e-12)
    np.testing.assert_allclose(vel, 0.0, atol=1e-12)


def test_cmm_matches_brute_force_momentum(o):
    rng = np.random.default_rng(3)
    q = np.concatenate([rng.normal(0, 0.3, 3), rng.uniform(-0.5, 0.5, 3), o.initial_state()[12:] + rng.normal(0, 0.3, o.nj)])
    v = rng.normal(0, 1.0, o.nq)
    A, com = o.cmm(q)
    eps = 1e-6
    c1, m, I1 = o.bodies(q + eps * v)
    c0, _, I0 = o.bodies(q - eps * v)
    c, _, I = o.bodies(q)
    np.testing.assert_allclose(com, (m[:, None] * c).sum(0) / m.sum(), atol=1e-13)
    cdot = (c1 - c0) / (2 * eps)
    lin = (m[:, None] * cdot).sum(0)
    # body angular velocity from the derivative of the world inertia: Idot = [w]x I - I [w]x  -> solve for w
    ang = np.zeros(3)
    for b in range(len(m)):
        Id = (I1[b] - I0[b]) / (2 * eps)
        # linear system for w: vec([w]x I - I [w]x) = Id
        Mx = np.zeros((9, 3))
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1
            S = np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
            Mx[:, k] = (S @ I[b] - I[b] @ S).reshape(-1)
        w = np.linalg.lstsq(Mx, Id.reshape(-1), rcond=None)[0]
        ang += I[b] @ w + m[b] * np.cross(c[b] - com, cdot[b])
    hv = A @ v
    np.testing.assert_allclose(hv[:3], lin, atol=1e-6)
    np.testing.assert_allclose(hv[3:], ang, atol=1e-5)
    # A_b has the block structure [m I, *; 0, *]
    np.testing.assert_allclose(A[:3, :3], o.total_mass * np.eye(3), atol=1e-12)
    np.testing.assert_allclose(A[3:, :3], 0.0, atol=1e-12)


# ---------------------------------------------------------------- E-3 Jacobians vs central differences
def test_dual_number_jacobians_match_finite_differences(o):
    rng = np.random.default_rng(1)
    x = o.initial_state() + rng.normal(0, 0.2, o.nx)
    u = rng.normal(0, 1.0, o.nu)
    u[:12] *= 50
    L = o.linearize(x, u)
    A = fd_jac(lambda z: o.flow_map(z, u)[0], x)
    B = fd_jac(lambda z: o.flow_map(x, z)[0], u)
    assert np.abs(A - L["A"]).max() < 1e-7 * max(1, np.abs(L["A"]).max())
    assert np.abs(B - L["B"]).max() < 1e-7
    assert np.abs(fd_jac(lambda z: o.flow_map(z, u)[2], x) - L["dvdx"].reshape(12, -1)).max() < 1e-7
    assert np.abs(fd_jac(lambda z: o.flow_map(x, z)[2], u) - L["dvdu"].reshape(12, -1)).max() < 1e-7
    assert np.abs(fd_jac(lambda z: o.flow_map(z, u)[1], x) - L["dpdx"].reshape(12, -1)).max() < 1e-7
    # structure the CUDA kernels rely on: rows 0..2 and 12.. of df/dx vanish, so do the base-position columns
    assert np.abs(L["A"][0:3]).max() == 0 and np.abs(L["A"][12:]).max() == 0 and np.abs(L["A"][:, 6:9]).max() == 0


# ---------------------------------------------------------------- E-4 Riccati vs dense KKT
def _random_lq(rng, N, nx, ms):
    A, B, b, Q, R, P, q, r = [], [], [], [], [], [], [], []
    for k in range(N):
        m = ms[k]
        A.append(np.eye(nx) + 0.1 * rng.normal(size=(nx, nx)))
        B.append(rng.normal(size=(nx, m)))
        b.append(0.1 * rng.normal(size=nx))
        W = rng.normal(size=(nx + m, nx + m))
        H = W @ W.T + 0.5 * np.eye(nx + m)
        Q.append(H[:nx, :nx])
        R.append(H[nx:, nx:])
        P.append(H[nx:, :nx])
        q.append(rng.normal(size=nx))
        r.append(rng.normal(size=m))
    return A, B, b, Q, R, P, q, r


def _kkt_solve(A, B, b, Q, R, P, q, r, dx0):
    N, nx = len(A), A[0].shape[0]
    ms = [Bk.shape[1] for Bk in B]
    # variables: dx_1..dx_N, du_0..du_{N-1}; dx_0 fixed
    nxv = N * nx
    nuv = sum(ms)
    uoff = np.concatenate([[0], np.cumsum(ms)])
    nv = nxv + nuv
    H = np.zeros((nv, nv))
    g = np.zeros(nv)
    E = np.zeros((N * nx, nv))
    d = np.zeros(N * nx)
    for k in range(N):
        us = slice(nxv + uoff[k], nxv + uoff[k + 1])
        H[us, us] += R[k]
        g[us] += r[k]
        if k == 0:
            g[us] += P[k] @ dx0
        else:
            xs = slice((k - 1) * nx, k * nx)
            H[xs, xs] += Q[k]
            g[xs] += q[k]
            H[us, xs] += P[k]
            H[xs, us] += P[k].T
        # dynamics: dx_{k+1} - A dx_k - B du_k = b
        rows = slice(k * nx, (k + 1) * nx)
        E[rows, k * nx:(k + 1) * nx] = np.eye(nx)
        E[rows, us] = -B[k]
        if k == 0:
            d[rows] = b[k] + A[k] @ dx0
        else:
            E[rows, (k - 1) * nx:k * nx] = -A[k]
            d[rows] = b[k]
    KKT = np.block([[H, E.T], [E, np.zeros((N * nx, N * nx))]])
    sol = np.linalg.solve(KKT, np.concatenate([-g, d]))
    dx = np.vstack([dx0, sol[:nxv].reshape(N, nx)])
    du = [sol[nxv + uoff[k]:nxv + uoff[k + 1]] for k in range(N)]
    return dx, du


def test_riccati_matches_dense_kkt():
    rng = np.random.default_rng(7)
    N, nx = 7, 5
    ms = [3, 2, 0, 3, 1, 0, 2]   # stage-varying input dimension including event stages (m = 0)
    lq = _random_lq(rng, N, nx, ms)
    for k in range(N):
        if ms[k] == 0:
            lq[0][k] = np.eye(nx)   # event stage: identity jump map
    dx0 = rng.normal(size=nx)
    dx, du, K = pyoracle.riccati(*lq, dx0)
    dx_ref, du_ref = _kkt_solve(*lq, dx0)
    np.testing.assert_allclose(dx, dx_ref, atol=1e-9)
    for k in range(N):
        np.testing.assert_allclose(du[k], du_ref[k], atol=1e-9)
        assert K[k].shape == (ms[k], nx)


# ---------------------------------------------------------------- E-5 projection
def _mode_constraints(o, mode, rng):
    x = o.initial_state() + rng.normal(0, 0.1, o.nx)
    x[0:6] = rng.normal(0, 0.2, 6)   # non-zero momentum: the stance-foot rows become inconsistent in C
    u = np.zeros(o.nu)
    u[12:] = rng.normal(0, 0.3, o.nj)
    o.set_dt_horizon(0.01, 0.03)
    o.set_mode_schedule([-5.0, 5.0], [3, mode, 3])
    o.set_target([0.0], [o.initial_state()])
    o.reset()
    o.run(0.0, x)
    return o.node_lq(0)


@pytest.mark.parametrize("mode,rows,rank", [(3, 12, 10), (1, 14, 13), (2, 14, 13)])
def test_projection_rank_and_pseudo_inverse(o, mode, rows, rank):
    rng = np.random.default_rng(5)
    L = _mode_constraints(o, mode, rng)
    assert L["nc_rows"] == rows and L["rank"] == rank and L["m"] == o.nu - rank
    assert np.linalg.matrix_rank(L["D"], tol=1e-9) == rank
    pyoracle.set_projection_mode(0)   # the Moore-Penrose variant ("projection_mode" 0 of the product)
    try:
        Px, Pu, Pe, rk = pyoracle.project(L["C"], L["D"], L["e"])
    finally:
        pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)
    assert rk == rank
    assert np.abs(L["D"] @ Pu).max() < 1e-12
    np.testing.assert_allclose(Pu.T @ Pu, np.eye(o.nu - rank), atol=1e-12)
    Dp = np.linalg.pinv(L["D"], rcond=1e-9)
    np.testing.assert_allclose(Px, -Dp @ L["C"], atol=1e-9)
    np.testing.assert_allclose(Pe, -Dp @ L["e"], atol=1e-9)


def test_projection_full_rank_fly_mode():
    rng = np.random.default_rng(9)
    D = rng.normal(size=(6, 10))
    C = rng.normal(size=(6, 7))
    e = rng.normal(size=6)
    Px, Pu, Pe, rk = pyoracle.project(C, D, e)
    assert rk == 6
    np.testing.assert_allclose(D @ Px + C, 0, atol=1e-12)
    np.testing.assert_allclose(D @ Pe + e, 0, atol=1e-12)
    np.testing.assert_allclose(D @ Pu, 0, atol=1e-12)


def test_projected_qp_matches_constrained_kkt(o):
    """Projected + Riccati solution of one real trot tick equals a dense equality-constrained KKT solve (un-projected)."""
    x0 = o.initial_state()
    o.set_dt_horizon(0.01, 0.06)
    o.set_mode_schedule([-5.0, 0.03, 5.0], [3, 1, 2, 3])
    o.set_target([0.0], [x0])
    o.reset()
    xs0 = x0.copy()
    xs0[6:] += 0.01   # zero momentum and zero joint velocities keep the (rank-deficient) stance-foot rows consistent
    o.run(0.0, xs0)
    st = o.step()
    n = o.info()["n_nodes"]
    N = n - 1
    nx, nu = o.nx, o.nu
    lqs = [o.node_lq(k) for k in range(N)]
    # dense KKT in (dx_1..dx_N, du_k for intermediate stages) with dynamics and state-input equality constraints
    var_u = [k for k in range(N) if lqs[k]["type"] == 0]
    nv = N * nx + len(var_u) * nu
    H = np.zeros((nv, nv))
    g = np.zeros(nv)
    rows_E, rhs = [], []
    dx0 = st["dx"][0]
    uidx = {k: N * nx + i * nu for i, k in enumerate(var_u)}
    for k in range(N):
        L = lqs[k]
        xs = None if k == 0 else slice((k - 1) * nx, k * nx)
        if L["type"] == 0:
            us = slice(uidx[k], uidx[k] + nu)
            H[us, us] += L["R"]
            g[us] += L["r"]
            if xs is not None:
                H[xs, xs] += L["Q"]
                g[xs] += L["q"]
            A, B = L["A"], L["B"]
        else:
            A, B = np.eye(nx), None
        for i in range(nx):
            row = np.zeros(nv)
            row[k * nx + i] = 1.0
            rr = L["b"][i]
            if B is not None:
                row[uidx[k]:uidx[k] + nu] = -B[i]
            if xs is None:
                rr += A[i] @ dx0
            else:
                row[xs] = -A[i]
            rows_E.append(row)
            rhs.append(rr)
        if L["type"] == 0:
            for i in range(L["nc_rows"]):
                row = np.zeros(nv)
                row[uidx[k]:uidx[k] + nu] = L["D"][i]
                rr = -L["e"][i]
                if xs is None:
                    rr -= L["C"][i] @ dx0
                else:
                    row[xs] = L["C"][i]
                rows_E.append(row)
                rhs.append(rr)
    E = np.array(rows_E)
    d = np.array(rhs)
    # rank-deficient, (slightly) inconsistent constraint rows: least-squares KKT via pseudo-inverse of the constraint block
    KKT = np.block([[H, E.T], [E, np.zeros((E.shape[0], E.shape[0]))]])
    sol = np.linalg.lstsq(KKT, np.concatenate([-g, d]), rcond=1e-12)[0]
    dx_ref = np.vstack([dx0, sol[:N * nx].reshape(N, nx)])
    assert np.abs(st["dx"] - dx_ref).max() < 1e-6
    for k in var_u:
        assert np.abs(st["du"][k] - sol[uidx[k]:uidx[k] + nu]).max() < 1e-5 * max(1.0, np.abs(st["du"][k]).max())


# ---------------------------------------------------------------- E-6 friction cone and barrier
def test_friction_cone_closed_form(o):
    h, g, H, pen = o.friction([0.0, 0.0, 100.0])
    assert abs(h - (0.5 * 100.0 - 5.0)) < 1e-12
    np.testing.assert_allclose(g, [0, 0, 0.5])
    np.testing.assert_allclose(H, np.diag([-1 / 5.0, -1 / 5.0, 0.0]), atol=1e-15)
    Fz = 51.641 * 9.81 / 4
    h2, *_ = o.friction([0, 0, Fz])
    assert h2 > 5.0   # nominal stance force is on the log branch of the barrier


def test_relaxed_barrier_is_c1_at_delta(o):
    d = 5.0
    lo, hi = o.barrier(d - 1e-9), o.barrier(d + 1e-9)
    assert abs(lo[0] - hi[0]) < 1e-8 and abs(lo[1] - hi[1]) < 1e-8
    p = o.barrier(10.0)
    np.testing.assert_allclose(p, [-0.1 * np.log(10.0), -0.1 / 10.0, 0.1 / 100.0])
    p = o.barrier(1.0)
    np.testing.assert_allclose(p[1:], [0.1 * (1.0 - 10.0) / 25.0, 0.1 / 25.0])


# ---------------------------------------------------------------- E-7 splines
def test_spline_cpg_hits_knots():
    tq = np.array([0.0, 0.175, 0.35, 0.0875])
    pos, vel = pyoracle.spline(0.0, 0.0, 0.05, 0.05, 0.35, 0.0, 0.0, tq)
    np.testing.assert_allclose(pos[:3], [0.0, 0.05, 0.0], atol=1e-15)
    np.testing.assert_allclose(vel[:3], [0.05, 0.0, 0.0], atol=1e-15)
    assert 0 < pos[3] < 0.05


def test_swing_scaling(o):
    # 0.35 s swing: scaling 1 -> peak height 0.05; 0.03 s swing: scaling 0.2 -> peak height 0.01
    et = np.array([0.0, 0.35, 0.70, 1.05])
    ms = np.array([3, 2, 1, 2, 3])   # left leg (contacts 0,1) swings during mode RF
    zv, zp = o.swing(et, ms, np.array([0.175, 0.0 + 1e-9]))
    np.testing.assert_allclose(zp[0, 0], 0.05, atol=1e-12)
    np.testing.assert_allclose(zv[1, 0], 0.05, atol=1e-6)
    np.testing.assert_allclose(zp[0, 2:], 0.0)
    et = np.array([0.0, 0.27, 0.30, 0.57])
    ms = np.array([3, 1, 0, 2, 3])   # FLY for 0.03 s: the left leg leaves at 0.27 and lands at 0.57 -> long swing, right leg: 0.0 -> 0.30
    zv, zp = o.swing(et, ms, np.array([0.15]))
    np.testing.assert_allclose(zp[0, 2], 0.05, atol=1e-12)   # right leg swing 0..0.30 s (scaling 1)
    et = np.array([0.0, 0.03, 1.0])
    ms = np.array([3, 0, 3, 3])
    zv, zp = o.swing(et, ms, np.array([0.015]))
    np.testing.assert_allclose(zp[0], 0.2 * 0.05, atol=1e-12)


def test_swing_undefined_liftoff_raises(o):
    with pytest.raises(RuntimeError):
        o.swing(np.array([0.5]), np.array([1, 3]), np.array([0.1]))


# ---------------------------------------------------------------- E-8 gait tiling
def test_gait_schedule_insert_trot(h1_model_path):
    o = Oracle(h1_model_path)
    et, ms = o.gait_peek()
    np.testing.assert_allclose(et, [0.5])
    assert list(ms) == [3, 3]
    # GaitReceiver inserts at the end of the horizon; phaseTransitionStanceTime is skipped when already in STANCE
    o.gait_insert([1, 2], [0.0, 0.35, 0.70], 1.0, 2.0)
    et, ms = o.gait_peek()
    np.testing.assert_allclose(et[:4], [0.5, 1.0, 1.35, 1.70])
    assert list(ms[:4]) == [3, 3, 1, 2] and ms[-1] == 3
    et2, ms2 = o.gait_get(0.2 - 1.0, 1.2 + 1.0)
    assert et2[-1] >= 2.2 and ms2[-1] == 3 and ms2[0] == 3
    assert np.all(np.diff(et2) > 0)


def test_time_discretization_with_events():
    t, e = pyoracle.time_discretization(0.0, 1.0, 0.01, [-0.25, 0.10, 0.45, 0.80, 1.15])
    assert len(t) == 104 and list(e).count(1) == 3 and list(e).count(2) == 3
    i = list(e).index(1)
    assert abs(t[i] - 0.10) < 1e-12 and e[i + 1] == 2 and t[i + 1] == t[i]
    assert t[0] == 0.0 and t[-1] == 1.0
    # an event between grid points shortens the interval before it and restarts the grid after it
    t, e = pyoracle.time_discretization(0.0, 0.1, 0.03, [0.05])
    np.testing.assert_allclose(t, [0.0, 0.03, 0.05, 0.05, 0.08, 0.1])
    assert list(e) == [0, 0, 1, 2, 0, 0]


# ---------------------------------------------------------------- E-9 end to end
def test_end_to_end_config1_and_warm_start(h1_model_path):
    o = Oracle(h1_model_path)
    x0 = o.initial_state()
    o.set_dt_horizon(0.015, 0.3)
    o.set_mode_schedule([-1.0, 5.0], [3, 3, 3])
    o.set_target([0.0], [x0])
    o.run(0.0, x0)
    i1 = o.info()
    assert i1["n_nodes"] == 21 and i1["step"] == 1.0
    th0 = np.sqrt(i1["before"][1] + i1["before"][2])
    th1 = np.sqrt(i1["after"][1] + i1["after"][2])
    assert th1 < th0
    s = o.solution()
    assert np.isfinite(s["K"]).all() and np.abs(s["K"]).max() > 1.0
    np.testing.assert_allclose(s["x"][0], x0)
    o.run(0.0, x0)
    i2 = o.info()
    np.testing.assert_allclose(i2["before"], i1["after"], atol=1e-9)   # warm start re-linearises at the accepted iterate
    assert i2["after"][0] <= i2["before"][0] + 1e-9
    xo, uo, mode = o.evaluate_policy(0.0, x0)
    np.testing.assert_allclose(xo, s["x"][0] if False else o.solution()["x"][0], atol=1e-12)
    assert mode == 3


def test_cmd_vel_target_matches_numpy(o, h1_model_path):
    import helpers
    from tools.ingest import read_model
    m = read_model(h1_model_path)
    x = o.initial_state()
    x[9] = 0.4
    x[6:8] = [0.2, -0.1]
    o.set_target_cmd_vel(0.3, x, [0.3, 0.1, 0.0, 0.2], 1.0)
    t, s = o.get_target()
    tt, ts = helpers.cmd_vel_target(x, 0.3, (0.3, 0.1, 0.0, 0.2), 1.0, m["com_height"], m["default_joint_state"])
    np.testing.assert_allclose(t, tt)
    np.testing.assert_allclose(s, ts, atol=1e-14)


# ------------------------------------------------------------------------------------------------ upstream's FullPivLU projection, emulated
def _random_stage_constraints(o, mode, seed):
    """(C, D, e) of one stage at a random state / input, taken from the oracle's own transcription (node_lq)."""
    rng = np.random.default_rng(seed)
    x = o.initial_state().copy()
    x[0:6] = rng.normal(0.0, 0.1, 6)
    x[9:12] += rng.uniform(-0.1, 0.1, 3)
    x[12:] += rng.uniform(-0.15, 0.15, o.nx - 12)
    o.reset(); o.set_dt_horizon(0.01, 0.02)
    o.set_mode_schedule([-1.0, 5.0], [3, mode, 3]); o.set_target([0.0, 1.0], [x, x])
    o.run(0.0, x)
    n = o.node_lq(0)
    return n["C"], n["D"], n["e"]


@pytest.mark.parametrize("mode,rank", [(3, 10), (1, 13), (2, 13)])
def test_fullpivlu_emulation_properties(o, mode, rank):
    """The emulation of upstream's luConstraintProjection (Eigen::FullPivLU: kernel + solve) finds the same rank as the default
    Moore-Penrose projection, its kernel spans the same null space, and its particular solution satisfies a full-rank sub-system
    exactly while ignoring the dependent stance-foot row."""
    from oracle import pyoracle
    Cm, D, e = _random_stage_constraints(o, mode, seed=11 + mode)
    pyoracle.set_projection_mode(0)
    try:
        Px0, Pu0, Pe0, r0 = pyoracle.project(Cm, D, e)
        pyoracle.set_projection_mode(1)
        Px1, Pu1, Pe1, r1 = pyoracle.project(Cm, D, e)
    finally:
        pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)
    assert r0 == r1 == rank == np.linalg.matrix_rank(D, tol=1e-9 * np.abs(D).max())
    assert np.abs(D @ Pu1).max() < 1e-10 * max(1.0, np.abs(D).max())
    # same null space: the orthogonal projector onto it is the same
    Q1, _ = np.linalg.qr(Pu1)
    assert np.abs(Pu0 @ Pu0.T - Q1 @ Q1.T).max() < 1e-9
    # residual of the particular solutions: Moore-Penrose = least squares (smallest residual), FullPivLU = exact on `rank` rows
    res0 = D @ Pe0 + e
    res1 = D @ Pe1 + e
    assert np.linalg.norm(res0) <= np.linalg.norm(res1) + 1e-12
    assert np.sum(np.abs(res1) > 1e-9 * max(1.0, np.abs(e).max())) <= D.shape[0] - rank
    # the two particular solutions differ only inside null(D) plus the part driven by the inconsistency
    assert np.abs(D @ (Pe1 - Pe0) - (res1 - res0)).max() < 1e-10


def test_projection_choice_identical_at_consistent_points_and_small_near_feasibility(h1_model_path):
    """Quantifies the difference between upstream's projection (default) and the selectable Moore-Penrose variant: with the stance feet at rest (config 2, cold start from the initial state) the
    equality constraints are consistent and both projections give the same solution to rounding; one warm tick later the stance-foot rows are
    inconsistent at second order only (foot angular velocity x dx) and the solutions agree to 1e-4 relative on cost and to 1e-3 on the inputs."""
    import helpers
    from oracle import pyoracle
    from tools.ingest import read_model
    m = read_model(h1_model_path)
    sols = {}
    for pm in (0, 1):
        pyoracle.set_projection_mode(pm)
        try:
            oo = Oracle(h1_model_path)
            x0 = oo.initial_state()
            et, ms = helpers.config2(oo.nx, x0, None, None)
            tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0, 0, 0), 1.0, m["com_height"], m["default_joint_state"])
            oo.reset(); oo.set_dt_horizon(0.01, 1.0); oo.set_mode_schedule(et, ms); oo.set_target(tt, ts)
            ticks = []
            for _ in range(2):
                oo.run(0.0, x0)
                so, io = oo.solution(), oo.info()
                ticks.append((np.array(so["x"]), np.array(so["u"]), np.array(so["K"]), np.array(io["after"])))
            sols[pm] = ticks
        finally:
            pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)
    (xa, ua, Ka, pa), (xb, ub, Kb, pb) = sols[0][0], sols[1][0]
    assert np.abs(xa - xb).max() < 1e-10 and np.abs(ua - ub).max() < 1e-9 * np.abs(ua).max() and np.abs(Ka - Kb).max() < 1e-9 * np.abs(Ka).max()
    (xa, ua, Ka, pa), (xb, ub, Kb, pb) = sols[0][1], sols[1][1]
    assert abs(pa[0] - pb[0]) < 1e-4 * abs(pa[0]) and abs(pa[2] - pb[2]) < 1e-4      # cost, equality-constraint SSE (the north star's 1e-4)
    assert np.abs(xa - xb).max() < 1e-3 and np.abs(ua - ub).max() < 1e-3 * np.abs(ua).max()


@pytest.mark.parametrize("mode", [3, 1, 2, 0])
def test_complete_pivoting_on_joint_block_equals_fullpivlu_on_stacked_rows(o, mode):
    """Design check for the CUDA FullPivLU projection (k_project<NJ, true>): the zero-force rows of open contacts are identity
    rows on columns the velocity rows never touch, so Gaussian elimination with complete pivoting of the joint block Dv alone (what one warp does:
    lane = joint column) gives the same Px / Pe joint rows as the oracle's FullPivLU emulation on the full stacked D of the stage."""
    from oracle import pyoracle
    Cm, D, e = _random_stage_constraints(o, mode, seed=5 + mode)
    nu = D.shape[1]; nj = nu - 12
    vel = [i for i in range(D.shape[0]) if np.abs(D[i, :12]).max() == 0.0]
    n_open = (D.shape[0] - len(vel)) // 3
    A = D[vel][:, 12:].copy(); Cv = Cm[vel]; ev = e[vel]
    nr = A.shape[0]
    rowid = list(range(nr)); colp = []; used = [False] * nj; piv = []
    for k in range(min(nr, nj)):
        best, bl, bi = -1.0, None, None
        for l in range(nj):            # first maximum in column-major order over the columns that are not pivots yet
            if used[l]:
                continue
            for i in range(k, nr):
                if abs(A[i, l]) > best:
                    best, bl, bi = abs(A[i, l]), l, i
        if not best > 0.0:
            break
        A[[k, bi], :] = A[[bi, k], :]; rowid[k], rowid[bi] = rowid[bi], rowid[k]
        p = A[k, bl]; piv.append(abs(p))
        for i in range(k + 1, nr):
            f = A[i, bl] / p
            for l in range(nj):
                if l == bl:
                    A[i, l] = f
                elif not used[l]:
                    A[i, l] -= f * A[k, l]
        used[bl] = True; colp.append(bl)
    thr = np.finfo(float).eps * min(D.shape[0], nu) * max([1.0 if n_open else 0.0] + piv)
    rank = sum(1 for p in piv if p > thr)

    def solve(rhs):
        g = np.array([rhs[rowid[i]] for i in range(nr)], dtype=float)
        for i in range(1, rank):
            for l in range(i):
                g[i] -= A[i, colp[l]] * g[l]
        for i in range(rank - 1, -1, -1):
            s = g[i]
            for j in range(i + 1, rank):
                s -= A[i, colp[j]] * g[j]
            g[i] = s / A[i, colp[i]]
        x = np.zeros(nj)
        for i in range(rank):
            x[colp[i]] = g[i]
        return x

    Pxj = -np.stack([solve(Cv[:, c]) for c in range(Cm.shape[1])], axis=1)
    Pej = -solve(ev)
    pyoracle.set_projection_mode(1)
    try:
        Px1, Pu1, Pe1, r1 = pyoracle.project(Cm, D, e)
    finally:
        pyoracle.set_projection_mode(pyoracle.DEFAULT_PROJECTION_MODE)
    assert r1 == rank + 3 * n_open
    scale = max(1.0, np.abs(Px1).max())
    assert np.abs(Px1[12:] - Pxj).max() < 1e-11 * scale and np.abs(Pe1[12:] - Pej).max() < 1e-11 * max(1.0, np.abs(Pe1).max())


def _gauss_jordan_lanes(Cv, Dv, ev, n_open, nu_full):
    """numpy restatement of lu_project (csrc/bmpc_kernels_project.cuh), lane by lane: Gauss-Jordan with complete pivoting on the augmented
    columns [Dv | Cv | ev], rows never swapped, scatter -col[row(k)] / pivot(k) into W at the pivot column's joint index."""
    nr, nj = Dv.shape
    nxa = Cv.shape[1]
    cols = np.concatenate([Dv, Cv, ev[:, None]], axis=1).astype(float)    # column j = lane j
    ncols = cols.shape[1]
    rowdone = [False] * nr
    used = [False] * nj
    spc, sip = [-1] * nr, [0.0] * nr
    maxpiv = 1.0 if n_open > 0 else 0.0
    epsd = np.finfo(float).eps * min(nr + 3 * n_open, nu_full)
    rank = 0
    for kk in range(min(nr, nj)):
        best, bl, prow = 0.0, None, None
        for lane in range(nj):           # warp maximum over the unused rows of the unused columns
            if not used[lane]:
                for i in range(nr):
                    if not rowdone[i]:
                        best = max(best, abs(cols[i, lane]))
        if best > 0.0:                   # first coefficient within 1e-10 of the maximum in (column, row) order
            tie = best * (1.0 - 1e-10)
            for lane in range(nj):
                if used[lane] or bl is not None:
                    continue
                for i in range(nr):
                    if not rowdone[i] and abs(cols[i, lane]) >= tie:
                        bl, prow = lane, i
                        break
        if bl is None:
            break
        p = cols[prow, bl]
        mp = max(maxpiv, abs(p))
        if not abs(p) > epsd * mp:
            break
        maxpiv = mp
        ip = 1.0 / p
        f = cols[:, bl] * ip
        f[prow] = 0.0
        cp = cols[prow, :].copy()
        cols -= np.outer(f, cp)
        used[bl] = True; rowdone[prow] = True; spc[prow] = bl; sip[prow] = ip; rank += 1
    W = np.zeros((nj, 32))
    free = [l for l in range(nj) if not used[l]]
    for lane in range(ncols):
        if lane < nj:
            if used[lane]:
                continue
            wl = 24 + free.index(lane)
        elif lane < nj + nxa:
            c = lane - nj
            wl = c if c < 6 else c + 3
        else:
            wl = 6
        for i in range(nr):
            if spc[i] >= 0:
                W[spc[i], wl] = -cols[i, lane] * sip[i]
        if lane < nj:
            W[lane, wl] = 1.0
    return W, rank


@pytest.mark.parametrize("mode", [3, 1, 2, 0])
@pytest.mark.parametrize("robot", ["h1", "g1"])
def test_in_warp_gauss_jordan_equals_fullpivlu_emulation(mode, robot):
    """The algorithm the CUDA projection runs (Gauss-Jordan on augmented columns, no row swaps, scatter by pivot column) gives the Px / Pe joint
    rows and a kernel basis that agree with the oracle's Eigen::FullPivLU emulation on the full stacked D, for every contact mode and both robots."""
    import os
    oo = Oracle(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", f"{robot}.model"))
    Cm, D, e = _random_stage_constraints(oo, mode, seed=21 + mode)
    nu = D.shape[1]; nj = nu - 12
    vel = [i for i in range(D.shape[0]) if np.abs(D[i, :12]).max() == 0.0]
    n_open = (D.shape[0] - len(vel)) // 3
    X = list(range(6)) + list(range(9, nu))
    assert np.abs(Cm[vel][:, 6:9]).max() == 0.0           # translation invariance: the base-position columns vanish
    W, rank = _gauss_jordan_lanes(Cm[vel][:, X], D[vel][:, 12:], e[vel], n_open, nu)
    Px1, Pu1, Pe1, r1 = pyoracle.project(Cm, D, e)        # default = FullPivLU emulation
    assert r1 == rank + 3 * n_open
    cols = [c if c < 6 else c + 3 for c in range(len(X))]
    scale = max(1.0, np.abs(Px1).max())
    assert np.abs(W[:, cols] - Px1[12:][:, X]).max() < 1e-11 * scale
    assert np.abs(W[:, 6] - Pe1[12:]).max() < 1e-11 * max(1.0, np.abs(Pe1).max())
    mj = nj - rank
    N = W[:, 24:24 + mj]
    assert mj == 0 or np.abs(D[vel][:, 12:] @ N).max() < 1e-10
    # same null space of the joint block as upstream's kernel() restricted to the joint rows (force columns of closed contacts are free there)
    if mj > 0:
        Nj = Pu1[12:, :]
        Nj = Nj[:, np.abs(Nj).max(axis=0) > 0]
        assert Nj.shape[1] == mj and np.linalg.matrix_rank(np.concatenate([N, Nj], axis=1), tol=1e-9) == mj


# ---------------------------------------------------------------- design check of the CUDA forward sweep (k_forward)
def test_forward_sweep_with_the_original_dynamics_and_their_structure(h1_model_path):
    """k_forward never forms the closed-loop map Phi = At + Bt Kt of the projected problem: it applies the ORIGINAL RK2 sensitivities of the compact LQ
    record, dx+ = A_d dx + B_d du + b with du = K dx + kappa, and relies on their structure (rows 0..2 of A_d - I vanish and B_d rows 0..2 are
    dt/m [I I I I]; rows 12.. of A_d - I vanish and B_d rows 12.. are dt I; the base-position columns 6..8 of A_d - I vanish).  Checked here on the oracle's
    dense matrices and its own step (computed in projected coordinates) over cold + warm tick of a trot with an event inside the horizon."""
    import helpers
    from tools.ingest import read_model
    mdl = read_model(h1_model_path)
    o = Oracle(h1_model_path)
    x0 = o.initial_state()
    et, ms = helpers.config2(o.nx, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0.1, 0.0, 0.2), 1.0, mdl["com_height"], mdl["default_joint_state"])
    o.set_dt_horizon(0.01, 0.25); o.set_mode_schedule(et, ms); o.set_target(tt, ts)
    nx, nu, mass = o.nx, o.nu, o.total_mass
    for tick in range(2):
        o.run(0.0, x0)
        st = o.step()
        n = o.info()["n_nodes"]
        seen_event = False
        for k in range(n - 1):
            lq = o.node_lq(k)
            dx, dxn = st["dx"][k], st["dx"][k + 1]
            if lq["type"] != 0:                                    # event node: A = I, no input
                seen_event = True
                np.testing.assert_allclose(dxn, dx + lq["b"], atol=1e-12)
                continue
            A, B, dt = lq["A"], lq["B"], lq["dt"]
            np.testing.assert_allclose(dxn, A @ dx + B @ st["du"][k] + lq["b"], rtol=0, atol=1e-9 * max(1.0, np.abs(dxn).max()))
            AmI = A - np.eye(nx)
            assert np.abs(AmI[0:3]).max() == 0.0 and np.abs(AmI[12:]).max() == 0.0 and np.abs(AmI[:, 6:9]).max() == 0.0
            np.testing.assert_allclose(B[0:3, 0:12], dt / mass * np.tile(np.eye(3), (1, 4)), rtol=1e-14, atol=0)
            assert np.abs(B[0:3, 12:]).max() == 0.0 and np.abs(B[12:, 0:12]).max() == 0.0
            np.testing.assert_allclose(B[12:, 12:], dt * np.eye(nu - 12), rtol=1e-14, atol=0)
        assert seen_event

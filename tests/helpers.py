"""Shared helpers for the parity tests: synthetic workloads and record expansion."""
import numpy as np

MODES = dict(FLY=0, LF=1, RF=2, STANCE=3)
GAITS = {
    "stance": ([3], [0.0, 0.5]),
    "trot": ([1, 2], [0.0, 0.35, 0.70]),
    "standing_trot": ([1, 3, 2, 3], [0.0, 0.30, 0.35, 0.65, 0.70]),
    "flying_trot": ([1, 0, 2, 0], [0.0, 0.27, 0.30, 0.57, 0.60]),
}


def tiled_schedule(gait, phase_offset, t_lo=-1.2, t_hi=2.4):
    """STANCE until the first event, then the gait template tiled, closed by STANCE (what GaitSchedule produces)."""
    modes, times = GAITS[gait]
    period = times[-1]
    et = [t_lo + (-phase_offset) % period]
    seq = [3]
    while et[-1] < t_hi:
        for i, m in enumerate(modes):
            seq.append(m)
            et.append(et[-1] + times[i + 1] - times[i])
    seq.append(3)
    # the last tiled phase must be followed by the final STANCE: drop nothing, schedule is [STANCE, tiles..., STANCE]
    return np.array(et), np.array(seq, dtype=np.int32)


def config2(nx, x_init, default_joints, com_height, cmd=(0.3, 0.0, 0.0, 0.0)):
    """H1 trot, t0 = 0, horizon 1.0, dt 0.01."""
    et = -0.95 + 0.35 * np.arange(9)
    ms = np.array([3, 1, 2, 1, 2, 1, 2, 1, 2, 3], dtype=np.int32)
    return et, ms


def cmd_vel_target(x_obs, t_obs, cmd, time_to_target, com_height, default_joints):
    """numpy restatement of TargetTrajectoriesPublisher.cpp:76-99 used to feed both oracle and GPU identically."""
    nx = x_obs.shape[0]
    z, y, x = x_obs[9:12]
    cz, sz, cy, sy, cx, sx = np.cos(z), np.sin(z), np.cos(y), np.sin(y), np.cos(x), np.sin(x)
    R = np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx], [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx], [-sy, cy * sx, cy * cx]])
    vr = R @ np.asarray(cmd[:3])
    s0, s1 = np.zeros(nx), np.zeros(nx)
    s0[0:3] = vr
    s1[0:3] = vr
    s0[6:12] = [x_obs[6], x_obs[7], com_height, x_obs[9], 0, 0]
    s1[6:12] = [x_obs[6] + vr[0] * time_to_target, x_obs[7] + vr[1] * time_to_target, com_height, x_obs[9] + cmd[3] * time_to_target, 0, 0]
    s0[12:] = default_joints
    s1[12:] = default_joints
    return np.array([t_obs, t_obs + time_to_target]), np.stack([s0, s1])


def randomized_instances(B, x_init, default_joints, joint_lo, joint_hi, seed=0):
    """SURVEY.md 8d config 3 distributions."""
    rng = np.random.default_rng(seed)
    nx = x_init.shape[0]
    X = np.tile(x_init, (B, 1))
    X[:, 0:6] = rng.normal(0.0, 0.1, (B, 6))
    X[:, 6:8] = rng.uniform(-0.5, 0.5, (B, 2))
    X[:, 8] = 0.93 + rng.uniform(-0.03, 0.03, B)
    X[:, 9] = rng.uniform(-np.pi, np.pi, B)
    X[:, 10:12] = rng.uniform(-0.1, 0.1, (B, 2))
    X[:, 12:] = np.clip(default_joints + rng.uniform(-0.15, 0.15, (B, nx - 12)), joint_lo, joint_hi)
    cmd = np.stack([rng.uniform(-0.5, 0.5, B), rng.uniform(-0.2, 0.2, B), np.zeros(B), rng.uniform(-0.3, 0.3, B)], axis=1)
    gait = rng.choice(["trot", "standing_trot", "flying_trot", "stance"], size=B, p=[0.55, 0.2, 0.15, 0.1])
    phase = rng.uniform(0.0, 0.70, B)
    return X, cmd, gait, phase


def expand_lq_record(rec, nj, total_mass):
    """Dense (A, B, b, q, r) of one stage from the GPU's compact LQ record (layout: csrc/bmpc_device.cuh Dims)."""
    nx = nu = 12 + nj
    nxa = nx - 3
    o_b, o_ad = 0, nx
    o_bd = o_ad + 9 * nxa
    o_q = o_bd + 9 * nu
    o_r = o_q + nx
    o_hb = o_r + nu
    o_cv = o_hb + 24
    o_dv = o_cv + 12 * nxa   # 12 row slots: raw rows (default, FullPivLU projection) or <= 10 compressed rows ("projection_mode" 0)
    o_ev = o_dv + 12 * nj
    o_misc = o_ev + 12
    o_fo = o_misc + 12
    misc = rec[o_misc:o_misc + 12]
    dt = misc[0]
    X = list(range(6)) + list(range(9, nx))
    A = np.eye(nx)
    A[np.ix_(range(3, 12), X)] += rec[o_ad:o_ad + 9 * nxa].reshape(9, nxa)
    B = np.zeros((nx, nu))
    for c in range(4):
        B[0:3, 3 * c:3 * c + 3] = np.eye(3) * dt / total_mass
    B[3:12, :] = rec[o_bd:o_bd + 9 * nu].reshape(9, nu)
    B[12:, 12:] = np.eye(nj) * dt
    nrows = int(misc[4])
    return dict(A=A, B=B, b=rec[o_b:o_b + nx].copy(), q=rec[o_q:o_q + nx].copy(), r=rec[o_r:o_r + nu].copy(), hb=rec[o_hb:o_hb + 24].reshape(4, 6).copy(),
                Cv=rec[o_cv:o_cv + nrows * nxa].reshape(nrows, nxa).copy(), Dv=rec[o_dv:o_dv + nrows * nj].reshape(nrows, nj).copy(),
                ev=rec[o_ev:o_ev + nrows].copy(), dt=dt, dq=misc[1], dr=misc[2], mode=int(misc[3]), nrows=nrows, type=int(misc[5]),
                perf=misc[6:9].copy(), fo=rec[o_fo:o_fo + 12].copy())

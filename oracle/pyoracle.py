"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE; parity unpinned - see oracle_model.hpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(march: str | None = None, out: str | None = None) -> str:
    """Compile the oracle.  Default flags are portable (x86-64-v3); bench may ask for 'native'."""
    out = out or os.path.join(_HERE, "liboracle.so")
    flags = ["-O3", f"-march={march or 'x86-64-v3'}", "-std=c++17", "-fPIC", "-pthread", "-shared"]
    src = os.path.join(_HERE, "oracle_api.cpp")
    deps = [src, os.path.join(_HERE, "oracle_solver.hpp"), os.path.join(_HERE, "oracle_model.hpp")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps) and march is None:
        return out
    subprocess.check_call(["g++", *flags, "-o", out, src])
    return out


def lib(path: str | None = None):
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(p):
        build()
    L = C.CDLL(p)
    L.orc_last_error.restype = C.c_char_p
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_char_p]
    L.orc_batch_create.restype = C.c_void_p
    L.orc_batch_create.argtypes = [C.c_char_p, C.c_int]
    L.orc_batch_instance.restype = C.c_void_p
    L.orc_batch_instance.argtypes = [C.c_void_p, C.c_int]
    L.orc_batch_run.restype = C.c_double
    L.orc_batch_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
    L.orc_batch_set_cmd_vel.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_double]
    L.orc_batch_get_observation.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
    L.orc_batch_destroy.argtypes = [C.c_void_p]
    L.orc_batch_set_observation.argtypes = [C.c_void_p, C.c_int, C.c_double, c_dp]
    L.orc_batch_get_observation.restype = None
    L.orc_total_mass.restype = C.c_double
    L.orc_total_mass.argtypes = [C.c_void_p]
    if path is None:
        _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(c_dp)


def _pi(a):
    return a.ctypes.data_as(c_ip)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Oracle:
    """One OCP instance: mirrors ocs2::MPC_BASE::run + MPC_MRT_Interface::evaluatePolicy for a single robot."""

    def __init__(self, model_path: str, handle=None, owner=True, L=None):
        self.L = L or lib()
        self._owner = owner
        self.h = C.c_void_p(handle) if handle is not None else C.c_void_p(self.L.orc_create(model_path.encode()))
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())
        nx, nu, nj = C.c_int(), C.c_int(), C.c_int()
        self.L.orc_dims(self.h, C.byref(nx), C.byref(nu), C.byref(nj))
        self.nx, self.nu, self.nj = nx.value, nu.value, nj.value
        self.nq = 6 + self.nj

    def __del__(self):
        try:
            if self._owner and self.h:
                self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- configuration
    @property
    def total_mass(self):
        return self.L.orc_total_mass(self.h)

    def initial_state(self):
        x = np.zeros(self.nx)
        self.L.orc_get_initial_state(self.h, _p(x))
        return x

    def set_dt_horizon(self, dt, horizon):
        self.L.orc_set_dt_horizon(self.h, C.c_double(dt), C.c_double(horizon))

    def set_sqp_iterations(self, n):
        self.L.orc_set_sqp_iterations(self.h, C.c_int(n))

    def reset(self):
        self.L.orc_reset(self.h)

    def set_mode_schedule(self, event_times, modes):
        et, ms = _d(event_times), _i(modes)
        assert len(ms) == len(et) + 1
        self.L.orc_set_mode_schedule(self.h, C.c_int(len(et)), _p(et), _pi(ms))

    def use_gait_schedule(self):
        self.L.orc_use_gait_schedule(self.h)

    def set_target(self, times, states):
        t, s = _d(times), _d(states)
        self.L.orc_set_target(self.h, C.c_int(len(t)), _p(t), _p(s))

    def set_target_cmd_vel(self, t_obs, x_obs, cmd, time_to_target):
        x, c = _d(x_obs), _d(cmd)
        self.L.orc_set_target_cmd_vel(self.h, C.c_double(t_obs), _p(x), _p(c), C.c_double(time_to_target))

    def get_target(self):
        t = np.zeros(16)
        s = np.zeros((16, self.nx))
        n = self.L.orc_get_target(self.h, _p(t), _p(s))
        return t[:n].copy(), s[:n].copy()

    def gait_insert(self, modes, times, start, final):
        m, t = _i(modes), _d(times)
        rc = self.L.orc_gait_insert(self.h, C.c_int(len(m)), _pi(m), _p(t), C.c_double(start), C.c_double(final))
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())

    def gait_get(self, lower, upper, cap=256):
        et = np.zeros(cap)
        ms = np.zeros(cap + 1, dtype=np.int32)
        n = self.L.orc_gait_get(self.h, C.c_double(lower), C.c_double(upper), C.c_int(cap), _p(et), _pi(ms))
        if n < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return et[:n].copy(), ms[:n + 1].copy()

    def gait_peek(self, cap=256):
        et = np.zeros(cap)
        ms = np.zeros(cap + 1, dtype=np.int32)
        n = self.L.orc_gait_peek(self.h, C.c_int(cap), _p(et), _pi(ms))
        return et[:n].copy(), ms[:n + 1].copy()

    # ---- solve
    def run(self, t0, x0):
        x = _d(x0)
        rc = self.L.orc_run(self.h, C.c_double(t0), _p(x))
        if rc < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return rc

    def solution(self):
        n = self.L.orc_num_nodes(self.h)
        t = np.zeros(n)
        ev = np.zeros(n, dtype=np.int32)
        self.L.orc_get_times(self.h, _p(t), _pi(ev))
        x = np.zeros((n, self.nx))
        u = np.zeros((n, self.nu))
        uff = np.zeros((n, self.nu))
        K = np.zeros((n, self.nu, self.nx))
        self.L.orc_get_solution(self.h, _p(x), _p(u), _p(uff), _p(K))
        return dict(t=t, events=ev, x=x, u=u, uff=uff, K=K)

    def info(self):
        o = np.zeros(12)
        self.L.orc_get_info(self.h, _p(o))
        return dict(before=o[0:3].copy(), after=o[3:6].copy(), step=o[6], trials=int(o[7]), armijo=o[8], dx_norm=o[9],
                    du_norm=o[10], n_nodes=int(o[11]))

    def step(self):
        n = self.L.orc_num_nodes(self.h)
        dx = np.zeros((n, self.nx))
        xl = np.zeros((n, self.nx))
        du = np.zeros((n - 1, self.nu))
        ul = np.zeros((n - 1, self.nu))
        self.L.orc_get_step(self.h, _p(dx), _p(du), _p(xl), _p(ul))
        return dict(dx=dx, du=du, x_lin=xl, u_lin=ul)

    def node_lq(self, k):
        nx, nu = self.nx, self.nu
        A, B, b = np.zeros((nx, nx)), np.zeros((nx, nu)), np.zeros(nx)
        Q, R, q, r = np.zeros((nx, nx)), np.zeros((nu, nu)), np.zeros(nx), np.zeros(nu)
        Cm, D, e = np.zeros((16, nx)), np.zeros((16, nu)), np.zeros(16)
        meta = np.zeros(5, dtype=np.int32)
        tdt = np.zeros(2)
        rc = self.L.orc_get_node_lq(self.h, C.c_int(k), _p(A), _p(B), _p(b), _p(Q), _p(R), _p(q), _p(r), _p(Cm), _p(D), _p(e), _pi(meta), _p(tdt))
        assert rc == 0
        nr = int(meta[2])
        Cm = Cm.reshape(-1)[:nr * nx].reshape(nr, nx).copy()
        D = D.reshape(-1)[:nr * nu].reshape(nr, nu).copy()
        return dict(type=int(meta[0]), mode=int(meta[1]), nc_rows=nr, m=int(meta[3]), rank=int(meta[4]), t=tdt[0], dt=tdt[1],
                    A=A, B=B, b=b, Q=Q, R=R, q=q, r=r, C=Cm, D=D, e=e[:nr].copy())

    def node_projection(self, k, m):
        nx, nu = self.nx, self.nu
        Px, Pu, Pe, K = np.zeros((nu, nx)), np.zeros(nu * max(m, 1)), np.zeros(nu), np.zeros((nu, nx))
        rc = self.L.orc_get_node_projection(self.h, C.c_int(k), _p(Px), _p(Pu), _p(Pe), _p(K))
        assert rc == 0
        return dict(Px=Px, Pu=Pu[:nu * m].reshape(nu, m).copy(), Pe=Pe, K=K)

    def evaluate_policy(self, t, x):
        xm = _d(x)
        xo, uo, mode = np.zeros(self.nx), np.zeros(self.nu), C.c_int()
        self.L.orc_evaluate_policy(self.h, C.c_double(t), _p(xm), _p(xo), _p(uo), C.byref(mode))
        return xo, uo, mode.value

    # ---- model maths
    def rollout_policy(self, t, x, time_step, substeps=1):
        """MRT_BASE::rolloutPolicy [UPSTREAM]: state after integrating the closed loop (u = uff(t) + K(t) x) from t to t + time_step."""
        x = _d(x).copy()
        steps = self.L.orc_rollout_policy(self.h, C.c_double(t), _p(x), C.c_double(time_step), C.c_int(substeps))
        if steps < 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return x, steps

    def flow_map(self, x, u):
        x, u = _d(x), _d(u)
        f, pos, vel = np.zeros(self.nx), np.zeros((4, 3)), np.zeros((4, 3))
        self.L.orc_flow_map(self.h, _p(x), _p(u), _p(f), _p(pos), _p(vel))
        return f, pos, vel

    def count_flow_map(self, x, u):
        """Exact operation counts {add, mul, div, trig} of one flow-map + contact-kinematics evaluation (counting scalar)."""
        x, u = _d(x), _d(u)
        out = (C.c_ulonglong * 4)()
        self.L.orc_count_flow_map(self.h, _p(x), _p(u), out)
        return dict(add=int(out[0]), mul=int(out[1]), div=int(out[2]), trig=int(out[3]))

    def linearize(self, x, u):
        x, u = _d(x), _d(u)
        nx, nu = self.nx, self.nu
        f, A, B = np.zeros(nx), np.zeros((nx, nx)), np.zeros((nx, nu))
        dpdx, dvdx, dvdu = np.zeros((4, 3, nx)), np.zeros((4, 3, nx)), np.zeros((4, 3, nu))
        self.L.orc_linearize(self.h, _p(x), _p(u), _p(f), _p(A), _p(B), _p(dpdx), _p(dvdx), _p(dvdu))
        return dict(f=f, A=A, B=B, dpdx=dpdx, dvdx=dvdx, dvdu=dvdu)

    def cmm(self, q):
        q = _d(q)
        A, com = np.zeros((6, self.nq)), np.zeros(3)
        self.L.orc_cmm(self.h, _p(q), _p(A), _p(com))
        return A, com

    def bodies(self, q):
        q = _d(q)
        nb = self.nj + 1
        c, m, I = np.zeros((nb, 3)), np.zeros(nb), np.zeros((nb, 3, 3))
        self.L.orc_bodies(self.h, _p(q), _p(c), _p(m), _p(I))
        return c, m, I

    def friction(self, F):
        F = _d(F)
        h = C.c_double()
        g, H, pen = np.zeros(3), np.zeros((3, 3)), np.zeros(3)
        self.L.orc_friction(self.h, _p(F), C.byref(h), _p(g), _p(H), _p(pen))
        return h.value, g, H, pen

    def barrier(self, h):
        pen = np.zeros(3)
        self.L.orc_barrier(self.h, C.c_double(h), _p(pen))
        return pen

    def swing(self, event_times, modes, tq):
        et, ms, tq = _d(event_times), _i(modes), _d(tq)
        zv, zp = np.zeros((len(tq), 4)), np.zeros((len(tq), 4))
        rc = self.L.orc_swing(self.h, C.c_int(len(et)), _p(et), _pi(ms), C.c_int(len(tq)), _p(tq), _p(zv), _p(zp))
        if rc != 0:
            raise RuntimeError(self.L.orc_last_error().decode())
        return zv, zp


def time_discretization(t0, tf, dt, events, cap=4096):
    L = lib()
    ev = _d(events)
    t = np.zeros(cap)
    e = np.zeros(cap, dtype=np.int32)
    n = L.orc_time_discretization(C.c_double(t0), C.c_double(tf), C.c_double(dt), C.c_int(len(ev)), _p(ev), C.c_int(cap), _p(t), _pi(e))
    assert n > 0
    return t[:n].copy(), e[:n].copy()


def spline(ts, ps, vs, mid, tf, pf, vf, tq):
    L = lib()
    tq = _d(tq)
    pos, vel = np.zeros(len(tq)), np.zeros(len(tq))
    L.orc_spline(*[C.c_double(v) for v in (ts, ps, vs, mid, tf, pf, vf)], C.c_int(len(tq)), _p(tq), _p(pos), _p(vel))
    return pos, vel


DEFAULT_PROJECTION_MODE = 1


def set_projection_mode(mode: int):
    """1 (default): emulation of upstream's FullPivLU projection (luConstraintProjection); 0: Moore-Penrose (process-wide switch)."""
    lib().orc_set_projection_mode(C.c_int(mode))


def project(Cm, D, e):
    L = lib()
    Cm, D, e = _d(Cm), _d(D), _d(e)
    nr, nu = D.shape
    nx = Cm.shape[1]
    Px, Pu, Pe = np.zeros((nu, nx)), np.zeros(nu * nu), np.zeros(nu)
    rank = L.orc_project(C.c_int(nr), C.c_int(nu), C.c_int(nx), _p(Cm), _p(D), _p(e), _p(Px), _p(Pu), _p(Pe))
    if rank < 0:
        raise RuntimeError(L.orc_last_error().decode())
    m = nu - rank
    return Px, Pu[:nu * m].reshape(nu, m).copy(), Pe, rank


def riccati(A, B, b, Q, R, P, q, r, dx0):
    """Lists of per-stage matrices (B[k] is nx x m_k)."""
    L = lib()
    N = len(A)
    nx = A[0].shape[0]
    m = _i([Bk.shape[1] for Bk in B])
    cat = lambda lst: _d(np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in lst] + [np.zeros(1)]))
    Ac, Bc, bc, Qc, Rc, Pc, qc, rc = map(cat, (A, B, b, Q, R, P, q, r))
    dx = np.zeros((N + 1, nx))
    du = np.zeros(int(m.sum()) + 1)
    Kt = np.zeros(int((m * nx).sum()) + 1)
    d0 = _d(dx0)
    st = L.orc_riccati(C.c_int(N), C.c_int(nx), _pi(m), _p(Ac), _p(Bc), _p(bc), _p(Qc), _p(Rc), _p(Pc), _p(qc), _p(rc), _p(d0), _p(dx), _p(du), _p(Kt))
    if st != 0:
        raise RuntimeError("riccati failed")
    dus, Ks, o, ok = [], [], 0, 0
    for k in range(N):
        dus.append(du[o:o + m[k]].copy())
        o += m[k]
        Ks.append(Kt[ok:ok + m[k] * nx].reshape(m[k], nx).copy())
        ok += m[k] * nx
    return dx, dus, Ks


class OracleBatch:
    def __init__(self, model_path, B, L=None):
        self.L = L or lib()
        self.h = C.c_void_p(self.L.orc_batch_create(model_path.encode(), C.c_int(B)))
        if not self.h:
            raise RuntimeError(self.L.orc_last_error().decode())
        self.B = B
        self.model_path = model_path
        self.inst = [Oracle(model_path, handle=self.L.orc_batch_instance(self.h, i), owner=False, L=self.L) for i in range(B)]

    def set_observation(self, i, t0, x0):
        x = _d(x0)
        self.L.orc_batch_set_observation(self.h, C.c_int(i), C.c_double(t0), _p(x))

    def set_cmd_vel(self, i, cmd, time_to_target):
        c = _d(cmd)
        self.L.orc_batch_set_cmd_vel(self.h, C.c_int(i), _p(c), C.c_double(time_to_target))

    def get_observation(self, i):
        t = C.c_double()
        x = np.zeros(self.inst[i].nx)
        self.L.orc_batch_get_observation(self.h, C.c_int(i), C.byref(t), _p(x))
        return t.value, x

    def run(self, first=0, count=None, threads=1, shift_dt=0.0):
        sec = self.L.orc_batch_run(self.h, C.c_int(first), C.c_int(self.B - first if count is None else count), C.c_int(threads), C.c_double(shift_dt))
        if sec < 0:
            raise RuntimeError("oracle batch run failed")
        return sec

    def __del__(self):
        try:
            self.L.orc_batch_destroy(self.h)
        except Exception:
            pass

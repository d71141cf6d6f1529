"""Host-side mirror of the reference's MPC interface for a *batch* of robots, on top of the C ABI (include/bmpc.h).

Names follow the OCS2 classes the reference drives from `BipedalController`
(bipedal_controllers/src/BipedalController.cpp:147-206, 282-351):

    BatchedMpcMrtInterface.reset()                  <- MPC_MRT_Interface::reset / resetMpcNode      (:147-148)
    .setCurrentObservation(t, x)                    <- MPC_MRT_Interface::setCurrentObservation     (:191)
    .setTargetTrajectories(times, states)           <- ReferenceManager::setTargetTrajectories       (:145, :153)
    .setModeSchedule(event_times, mode_sequence)    <- ReferenceManager::setModeSchedule
    .insertGait(name, start, final)                 <- GaitSchedule::insertModeSequenceTemplate      (GaitReceiver.cpp:49-59)
    .advanceMpc()                                   <- MPC_MRT_Interface::advanceMpc                 (:339)
    .evaluatePolicy(t, x) -> (xOpt, uOpt, mode)     <- MPC_MRT_Interface::evaluatePolicy             (:200)
    .getPolicy()                                    <- MPC_MRT_Interface::getPolicy                  (:203)

The compute path is the CUDA library only: if libbmpc.so is missing or no GPU is present this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbmpc.so")
REPO_ROOT = os.path.dirname(_HERE)
DEFAULT_MODELS = {"h1": os.path.join(REPO_ROOT, "configs", "h1.model"), "g1": os.path.join(REPO_ROOT, "configs", "g1.model")}

MODE_FLY, MODE_LF, MODE_RF, MODE_STANCE = 0, 1, 2, 3


class BmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bmpc error {code}: {msg}")
        self.code = code


class _Config(C.Structure):
    _fields_ = [("model_file", C.c_char_p), ("task_file", C.c_char_p), ("reference_file", C.c_char_p), ("gait_file", C.c_char_p),
                ("urdf_file", C.c_char_p), ("batch", C.c_int), ("device", C.c_int), ("dt", C.c_double), ("time_horizon", C.c_double),
                ("max_events", C.c_int), ("max_target_points", C.c_int), ("sqp_iterations", C.c_int), ("max_event_nodes", C.c_int)]


class DeviceView(C.Structure):
    _fields_ = [("n_nodes", C.c_void_p), ("times", C.c_void_p), ("events", C.c_void_p), ("x", C.c_void_p), ("u", C.c_void_p),
                ("uff", C.c_void_p), ("K", C.c_void_p), ("max_nodes", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("batch", C.c_int),
                ("slab", C.c_void_p), ("slab_bytes", C.c_ulonglong)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def load_library(path: str | None = None):
    """Loads libbmpc.so (built in-tree by `make -C bipedal_control_b200/csrc`).  No fallback: raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: build the CUDA extension (python -c 'import __graft_entry__ as g; g.build()')")
    L = C.CDLL(p)
    L.bmpc_last_error.restype = C.c_char_p
    L.bmpc_last_error.argtypes = [C.c_void_p]
    L.bmpc_create.argtypes = [C.POINTER(_Config), C.POINTER(C.c_void_p)]
    L.bmpc_destroy.argtypes = [C.c_void_p]
    L.bmpc_get_stream.restype = C.c_void_p
    L.bmpc_get_stream.argtypes = [C.c_void_p]
    for name in ("bmpc_get_dims", "bmpc_get_initial_state", "bmpc_export_model", "bmpc_reset", "bmpc_set_observations", "bmpc_set_observations_device",
                 "bmpc_set_target_trajectories", "bmpc_set_target_trajectories_device", "bmpc_set_targets_from_cmd_vel", "bmpc_set_targets_from_cmd_vel_device", "bmpc_shift_observations", "bmpc_get_observations", "bmpc_set_mode_schedules",
                 "bmpc_set_mode_schedules_device", "bmpc_gait_insert", "bmpc_gait_insert_named", "bmpc_use_gait_schedule", "bmpc_gait_peek",
                 "bmpc_advance", "bmpc_advance_async", "bmpc_synchronize", "bmpc_get_policy", "bmpc_get_device_view", "bmpc_get_performance",
                 "bmpc_get_status", "bmpc_evaluate_policy", "bmpc_get_launch_count", "bmpc_get_phase_times", "bmpc_enable_phase_timing",
                 "bmpc_poll", "bmpc_get_device_view_inflight", "bmpc_get_tick_stats", "bmpc_rollout_observations", "bmpc_set_rollout_settings",
                 "bmpc_exchange_create_id", "bmpc_exchange_init", "bmpc_exchange_start", "bmpc_exchange_wait", "bmpc_exchange_view", "bmpc_exchange_destroy",
                 "bmpc_debug_copy", "bmpc_debug_record_sizes", "bmpc_debug_set_option"):
        getattr(L, name).restype = C.c_int
    if path is None:
        _lib = L
    return L


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


class BatchedMpcMrtInterface:
    """B independent centroidal MPC problems solved in lockstep on one GPU."""

    def __init__(self, batch: int, model_file: str | None = None, robot: str = "h1", device: int = 0, dt: float = 0.0, time_horizon: float = 0.0,
                 max_events: int = 0, max_target_points: int = 0, sqp_iterations: int = 0, task_file: str | None = None,
                 reference_file: str | None = None, gait_file: str | None = None, urdf_file: str | None = None, max_event_nodes: int = 0):
        self.L = load_library()
        enc = lambda s: s.encode() if s else None
        if model_file is None and task_file is None:
            model_file = DEFAULT_MODELS[robot]
        cfg = _Config(enc(model_file), enc(task_file), enc(reference_file), enc(gait_file), enc(urdf_file), int(batch), int(device), float(dt),
                      float(time_horizon), int(max_events), int(max_target_points), int(sqp_iterations), int(max_event_nodes))
        h = C.c_void_p()
        rc = self.L.bmpc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise BmpcError(rc, self.L.bmpc_last_error(None).decode())
        self.h = h
        nx, nu, b, ns = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.bmpc_get_dims(self.h, C.byref(nx), C.byref(nu), C.byref(b), C.byref(ns)))
        self.nx, self.nu, self.batch, self.max_nodes = nx.value, nu.value, b.value, ns.value
        self.nj = self.nx - 12

    def close(self):
        if getattr(self, "h", None):
            self.L.bmpc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise BmpcError(rc, self.L.bmpc_last_error(self.h).decode())
        return rc

    # ------------------------------------------------------------------ configuration / inputs
    def initialState(self):
        x = np.zeros(self.nx)
        self._ck(self.L.bmpc_get_initial_state(self.h, _p(x)))
        return x

    def exportModel(self, path):
        self._ck(self.L.bmpc_export_model(self.h, path.encode()))

    def reset(self, instance: int = -1):
        """MPC_MRT_Interface::reset: drops warm start, policy and gait bookkeeping of one instance (all if instance < 0)."""
        self._ck(self.L.bmpc_reset(self.h, C.c_int(instance)))

    def setCurrentObservation(self, t, x):
        t = _d(np.broadcast_to(t, (self.batch,)))
        x = _d(np.broadcast_to(x, (self.batch, self.nx)))
        self._ck(self.L.bmpc_set_observations(self.h, _p(t), _p(x)))

    def setCurrentObservationDevice(self, t_ptr: int, x_ptr: int):
        self._ck(self.L.bmpc_set_observations_device(self.h, C.c_void_p(t_ptr), C.c_void_p(x_ptr)))

    def setTargetTrajectories(self, times, states):
        times = _d(times)
        if times.ndim == 1:
            times = _d(np.broadcast_to(times, (self.batch, times.shape[0])))
        npts = times.shape[1]
        states = _d(np.broadcast_to(states, (self.batch, npts, self.nx)))
        self._ck(self.L.bmpc_set_target_trajectories(self.h, C.c_int(npts), _p(times), _p(states)))

    def setTargetTrajectoriesDevice(self, npts: int, times_ptr: int, states_ptr: int):
        self._ck(self.L.bmpc_set_target_trajectories_device(self.h, C.c_int(npts), C.c_void_p(times_ptr), C.c_void_p(states_ptr)))

    def setTargetsFromCmdVel(self, cmd, time_to_target):
        cmd = _d(np.broadcast_to(cmd, (self.batch, 4)))
        self._ck(self.L.bmpc_set_targets_from_cmd_vel(self.h, _p(cmd), C.c_double(time_to_target)))

    def setTargetsFromCmdVelDevice(self, cmd_ptr: int, time_to_target: float):
        self._ck(self.L.bmpc_set_targets_from_cmd_vel_device(self.h, C.c_void_p(cmd_ptr), C.c_double(time_to_target)))

    def shiftObservations(self, dt: float):
        """t0 += dt, x0 = optimized state at the new time (device side, perfect-model closed loop)."""
        self._ck(self.L.bmpc_shift_observations(self.h, C.c_double(dt)))

    def rolloutObservations(self, time_step: float, substeps: int = 1):
        """MRT_BASE::rolloutPolicy for the batch: the observations advance by `time_step` under the feedback policy u = uff(t) + K(t) x (device side)."""
        self._ck(self.L.bmpc_rollout_observations(self.h, C.c_double(time_step), C.c_int(substeps)))

    def setRolloutSettings(self, abs_tol=1e-5, rel_tol=1e-3, initial_time_step=0.015):
        self._ck(self.L.bmpc_set_rollout_settings(self.h, C.c_double(abs_tol), C.c_double(rel_tol), C.c_double(initial_time_step)))

    def getObservations(self):
        t, x = np.zeros(self.batch), np.zeros((self.batch, self.nx))
        self._ck(self.L.bmpc_get_observations(self.h, _p(t), _p(x)))
        return t, x

    def setModeSchedule(self, event_times, mode_sequence, n_events=None):
        """event_times [B, stride] (or [stride] for all), mode_sequence [B, stride+1]; n_events[B] optional."""
        et = _d(event_times)
        ms = _i(mode_sequence)
        if et.ndim == 1:
            et = _d(np.broadcast_to(et, (self.batch, et.shape[0])))
            ms = _i(np.broadcast_to(ms, (self.batch, ms.shape[0])))
        stride = et.shape[1]
        assert ms.shape[1] == stride + 1
        ne = _i(np.full(self.batch, stride) if n_events is None else n_events)
        self._ck(self.L.bmpc_set_mode_schedules(self.h, C.c_int(stride), _pi(ne), _p(et), _pi(ms)))

    def setModeScheduleDevice(self, stride: int, n_events_ptr: int, event_times_ptr: int, mode_sequence_ptr: int):
        self._ck(self.L.bmpc_set_mode_schedules_device(self.h, C.c_int(stride), C.c_void_p(n_events_ptr), C.c_void_p(event_times_ptr), C.c_void_p(mode_sequence_ptr)))

    def insertGait(self, gait, start_time, final_time, instance=-1):
        if isinstance(gait, str):
            self._ck(self.L.bmpc_gait_insert_named(self.h, C.c_int(instance), gait.encode(), C.c_double(start_time), C.c_double(final_time)))
        else:
            modes, times = _i(gait[0]), _d(gait[1])
            self._ck(self.L.bmpc_gait_insert(self.h, C.c_int(instance), C.c_int(len(modes)), _pi(modes), _p(times), C.c_double(start_time), C.c_double(final_time)))

    def useGaitSchedule(self, enable=True):
        self._ck(self.L.bmpc_use_gait_schedule(self.h, C.c_int(1 if enable else 0)))

    def gaitPeek(self, instance, cap=256):
        et = np.zeros(cap)
        ms = np.zeros(cap + 1, dtype=np.int32)
        n = self._ck(self.L.bmpc_gait_peek(self.h, C.c_int(instance), C.c_int(cap), _p(et), _pi(ms)))
        return et[:n].copy(), ms[:n + 1].copy()

    # ------------------------------------------------------------------ solve
    def advanceMpc(self):
        self._ck(self.L.bmpc_advance(self.h))

    def advanceMpcAsync(self):
        self._ck(self.L.bmpc_advance_async(self.h))

    def synchronize(self):
        self._ck(self.L.bmpc_synchronize(self.h))

    def poll(self) -> bool:
        """True while a tick is still running; publishes it (makes its policy the current one) once it has finished."""
        return self._ck(self.L.bmpc_poll(self.h)) == 1

    def tickStats(self):
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.bmpc_get_tick_stats(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(total_trials=a.value, max_trials=b.value, failed_instances=c.value, status_or=d.value)

    # ------------------------------------------------------------------ outputs
    def getPolicy(self, first=0, count=None, with_gains=True):
        count = self.batch - first if count is None else count
        NS, nx, nu = self.max_nodes, self.nx, self.nu
        n = np.zeros(count, dtype=np.int32)
        t = np.zeros((count, NS))
        ev = np.zeros((count, NS), dtype=np.int32)
        x = np.zeros((count, NS, nx))
        u = np.zeros((count, NS, nu))
        uff = np.zeros((count, NS, nu))
        K = np.zeros((count, NS, nu, nx)) if with_gains else None
        self._ck(self.L.bmpc_get_policy(self.h, C.c_int(first), C.c_int(count), _pi(n), _p(t), _pi(ev), _p(x), _p(u), _p(uff), _p(K) if with_gains else None))
        return dict(n_nodes=n, t=t, events=ev, x=x, u=u, uff=uff, K=K)

    def getDeviceView(self, inflight: bool = False) -> DeviceView:
        """Device pointers of the published policy; inflight=True: of the policy the tick in flight is writing (valid in stream order)."""
        v = DeviceView()
        self._ck((self.L.bmpc_get_device_view_inflight if inflight else self.L.bmpc_get_device_view)(self.h, C.byref(v)))
        return v

    def getPerformanceIndices(self):
        """[B, 8]: cost/dynamicsSSE/eqSSE before, the same after the step, step size, armijo descent metric."""
        p = np.zeros((self.batch, 8))
        self._ck(self.L.bmpc_get_performance(self.h, _p(p)))
        return p

    def getStatus(self):
        s = np.zeros(self.batch, dtype=np.int32)
        self._ck(self.L.bmpc_get_status(self.h, _pi(s)))
        return s

    def evaluatePolicy(self, t, x):
        t = _d(np.broadcast_to(t, (self.batch,)))
        x = _d(np.broadcast_to(x, (self.batch, self.nx)))
        xo, uo, mo = np.zeros((self.batch, self.nx)), np.zeros((self.batch, self.nu)), np.zeros(self.batch, dtype=np.int32)
        self._ck(self.L.bmpc_evaluate_policy(self.h, _p(t), _p(x), _p(xo), _p(uo), _pi(mo)))
        return xo, uo, mo

    # ------------------------------------------------------------------ multi-GPU policy exchange (one process per GPU)
    def exchangeInit(self, dist, rank: int, world: int, max_ctas: int = 0, copy_engines: int = 0):
        """bmpc_exchange_init: rank 0 creates the NCCL id, `dist` (torch.distributed, any backend) carries its 128 bytes to the other ranks.
        copy_engines: 0 plain buffers, 1 NCCL copy-engine all-gather (no SM), 2 symmetric windows with NCCL's SM kernels.  Device views obtained
        before this call are stale afterwards (modes 1 and 2 move the policy slabs into NCCL-registered memory)."""
        import torch
        buf = (C.c_char * 128)()
        if rank == 0:
            rc = self.L.bmpc_exchange_create_id(buf)
            if rc != 0:
                raise BmpcError(rc, self.L.bmpc_last_error(None).decode())
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().tolist())
        idb = (C.c_char * 128).from_buffer_copy(raw)
        self._ck(self.L.bmpc_exchange_init(self.h, C.c_int(rank), C.c_int(world), idb, C.c_int(int(max_ctas)), C.c_int(int(copy_engines))))

    def exchangeStart(self):
        self._ck(self.L.bmpc_exchange_start(self.h))

    def exchangeWait(self):
        self._ck(self.L.bmpc_exchange_wait(self.h))

    def exchangeView(self):
        """(device pointer of the gathered slabs, bytes per slab, number of ranks, active mode: 0 plain, 1 copy engines, 2 symmetric windows)."""
        p, n, r = C.c_void_p(), C.c_ulonglong(), C.c_int()
        mode = self._ck(self.L.bmpc_exchange_view(self.h, C.byref(p), C.byref(n), C.byref(r)))
        return int(p.value or 0), int(n.value), int(r.value), int(mode)

    def exchangeDestroy(self):
        self._ck(self.L.bmpc_exchange_destroy(self.h))

    # ------------------------------------------------------------------ instrumentation
    def launchCount(self):
        return self.L.bmpc_get_launch_count(self.h)

    def enablePhaseTiming(self, enable=True):
        self._ck(self.L.bmpc_enable_phase_timing(self.h, C.c_int(1 if enable else 0)))

    def phaseTimes(self):
        ms = (C.c_float * 9)()
        self._ck(self.L.bmpc_get_phase_times(self.h, ms))
        names = ["setup", "lq", "projection", "riccati", "policy_expand", "forward", "linesearch", "finalize"]
        out = {k: float(ms[i]) for i, k in enumerate(names)}
        out["linesearch_trials"] = int(ms[8])   # largest number of trials any instance needed
        return out

    def stream(self) -> int:
        return int(self.L.bmpc_get_stream(self.h) or 0)

    def setOption(self, name, value):
        self._ck(self.L.bmpc_debug_set_option(self.h, name.encode(), C.c_int(int(value))))

    def recordSizes(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.bmpc_debug_record_sizes(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def debugCopy(self, name, instance):
        a, b, c = self.recordSizes()
        cap = self.max_nodes * max(a, b, c, self.nx * self.nu)
        buf = np.zeros(cap)
        n = self._ck(self.L.bmpc_debug_copy(self.h, name.encode(), C.c_int(instance), _p(buf), C.c_int(cap)))
        return buf[:n].reshape(self.max_nodes, -1).copy()


# ---------------------------------------------------------------------- synthetic workloads (BASELINE.json configs)
def trot_schedule():
    """Config 2 mode schedule: events -0.95 + 0.35 j (j = 0..8), modes STANCE, LF, RF, ..., STANCE."""
    et = -0.95 + 0.35 * np.arange(9)
    ms = np.array([3, 1, 2, 1, 2, 1, 2, 1, 2, 3], dtype=np.int32)
    return et, ms

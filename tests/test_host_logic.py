"""CPU tests of the product's host side: C ABI surface, file ingestion, gait bookkeeping, multi-rank sharding (gloo)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bipedal_control_b200", "libbmpc.so")
REF = "/root/reference/bipedal_robot_example/unitree_h1"


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bipedal_control_b200", "csrc")])
    L = C.CDLL(LIB)
    L.bmpc_last_error.restype = C.c_char_p
    L.bmpc_last_error.argtypes = [C.c_void_p]
    return L


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bmpc.h")).read()
    names = set(re.findall(r"\b(bmpc_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (bmpc_[a-z_0-9]+)", out))
    missing = names - exported
    assert not missing, f"declared in include/bmpc.h but not exported: {sorted(missing)}"


def test_exchange_binds_nccl_at_run_time():
    """bmpc_exchange_create_id dlopens libnccl.so.2 and returns a 128-byte id: no link-time NCCL dependency, no GPU needed for this step.
    (In a subprocess: a process that loads the system NCCL first can no longer import a PyTorch that ships a newer one.)"""
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "nccl" not in out
    code = ("import ctypes as C; L = C.CDLL(%r); L.bmpc_last_error.restype = C.c_char_p; L.bmpc_last_error.argtypes = [C.c_void_p]; "
            "b = (C.c_char * 128)(); rc = L.bmpc_exchange_create_id(b); print(rc, any(bytes(b)), L.bmpc_last_error(None))" % LIB)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.stdout.split()[:2] == ["0", "True"], r.stdout + r.stderr


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_blackwell_instructions_present():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    assert "DMMA" in sass        # FP64 tensor-core tiles of the Riccati recursion
    assert "UBLKCP" in sass      # TMA bulk copies staging the stage records


def test_create_fails_loudly_without_gpu(lib):
    """No CPU fallback: on a machine without a CUDA device bmpc_create must return BMPC_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")

    class Cfg(C.Structure):
        _fields_ = [("model_file", C.c_char_p), ("task_file", C.c_char_p), ("reference_file", C.c_char_p), ("gait_file", C.c_char_p), ("urdf_file", C.c_char_p),
                    ("batch", C.c_int), ("device", C.c_int), ("dt", C.c_double), ("time_horizon", C.c_double), ("max_events", C.c_int), ("max_target_points", C.c_int), ("sqp_iterations", C.c_int), ("max_event_nodes", C.c_int)]
    cfg = Cfg(os.path.join(ROOT, "configs", "h1.model").encode(), None, None, None, None, 4, 0, 0.0, 0.0, 0, 0, 0, 0)
    h = C.c_void_p()
    rc = lib.bmpc_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value
    assert b"CUDA" in lib.bmpc_last_error(None)
    from bipedal_control_b200 import BatchedMpcMrtInterface, BmpcError
    with pytest.raises(BmpcError):
        BatchedMpcMrtInterface(2)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_cpp_ingestion_of_reference_files_matches_committed_model(lib, tmp_path):
    from tools.ingest import read_model, build_model
    out = str(tmp_path / "h1_cpp.model")
    rc = lib.bmpc_convert_model(f"{REF}/h1_ocs2_config/config/task/task.info".encode(), f"{REF}/h1_ocs2_config/config/command/reference.info".encode(),
                                f"{REF}/h1_ocs2_config/config/command/gait.info".encode(), f"{REF}/h1_description/urdf/h1_with_sole.urdf".encode(), out.encode())
    assert rc == 0, lib.bmpc_last_error(None)
    a, b = read_model(out), read_model(os.path.join(ROOT, "configs", "h1.model"))
    assert set(b) <= set(a)
    for k, v in b.items():
        if isinstance(v, str):
            continue
        np.testing.assert_allclose(np.asarray(a[k], dtype=float), np.asarray(v, dtype=float), atol=1e-12, err_msg=k)
    # and the python ingestion (tools/ingest.py) regenerates the committed file from the reference tree
    py = build_model(f"{REF}/h1_ocs2_config/config/task/task.info", f"{REF}/h1_ocs2_config/config/command/reference.info",
                     f"{REF}/h1_ocs2_config/config/command/gait.info", f"{REF}/h1_description/urdf/h1_with_sole.urdf", "h1")
    np.testing.assert_allclose(py["R_joint"], b["R_joint"], atol=1e-14)
    assert abs(py["total_mass"] - 51.641) < 1e-9


def test_missing_files_are_reported(lib):
    rc = lib.bmpc_convert_model(b"/nonexistent/task.info", b"/nonexistent/reference.info", None, b"/nonexistent/robot.urdf", b"/tmp/x.model")
    assert rc == -1 and b"not found" in lib.bmpc_last_error(None)   # BipedalRobotInterface.cpp:71-90 throws std::invalid_argument


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "host_shim.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "host_shim.cpp")])
    L = C.CDLL(so)
    L.shim_gait_create.restype = C.c_void_p
    return L


def _arr(a, t):
    a = np.ascontiguousarray(a, dtype=t)
    return a, a.ctypes.data_as(C.POINTER(C.c_int if t == np.int32 else C.c_double))


def test_gait_schedule_matches_oracle(shim, h1_model_path):
    """bmpc_gait.h and the oracle restate GaitSchedule.cpp independently; drive both through the same call sequence."""
    from oracle.pyoracle import Oracle
    import helpers
    o = Oracle(h1_model_path)
    im, imp = _arr([3, 3], np.int32)
    ie, iep = _arr([0.5], np.float64)
    tm, tmp = _arr([3], np.int32)
    tt, ttp = _arr([0.0, 1.0], np.float64)
    g = C.c_void_p(shim.shim_gait_create(2, imp, 1, iep, 1, tmp, ttp, C.c_double(0.4)))
    rng = np.random.default_rng(0)
    t = 0.0
    for step in range(60):
        if step in (5, 20, 33, 47):
            name = ["trot", "flying_trot", "stance", "standing_trot"][[5, 20, 33, 47].index(step)]
            modes, times = helpers.GAITS[name]
            m, mp = _arr(modes, np.int32)
            tm_, tp_ = _arr(times, np.float64)
            assert shim.shim_gait_insert(g, len(modes), mp, tp_, C.c_double(t + 1.0), C.c_double(t + 2.0)) == 0
            o.gait_insert(modes, times, t + 1.0, t + 2.0)
        et = np.zeros(256)
        ms = np.zeros(257, dtype=np.int32)
        n = shim.shim_gait_get(g, C.c_double(t - 1.0), C.c_double(t + 2.0), 256, et.ctypes.data_as(C.POINTER(C.c_double)), ms.ctypes.data_as(C.POINTER(C.c_int)))
        assert n >= 0
        eo, mo = o.gait_get(t - 1.0, t + 2.0)
        np.testing.assert_allclose(et[:n], eo, atol=1e-12)
        assert list(ms[:n + 1]) == list(mo)
        assert ms[0] == 3 and ms[n] == 3
        t += 0.02 + 0.05 * rng.random()
    shim.shim_gait_destroy(g)


def test_tiled_schedule_helper_is_swing_complete(oracle_h1):
    """Every synthetic schedule used by the benchmarks gives each swing phase a lift-off and a touch-down."""
    import helpers
    for gait in helpers.GAITS:
        for phase in (0.0, 0.13, 0.41, 0.69):
            et, ms = helpers.tiled_schedule(gait, phase, t_hi=3.6)
            assert len(et) <= 40 and len(ms) == len(et) + 1
            oracle_h1.swing(et, ms, np.linspace(0.0, 3.0, 31))   # raises if a lift-off / touch-down time is undefined


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bipedal_control_b200.sharding import shard_range, all_gather_policies
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    B_total, NS, nx, nu = 10, 5, 22, 22
    lo, hi = shard_range(B_total, rank, world)
    local = {"K": torch.full((hi - lo, NS, nu, nx), float(rank + 1), dtype=torch.float64), "uff": torch.arange(lo, hi, dtype=torch.float64).repeat_interleave(NS * nu).view(hi - lo, NS, nu)}
    full = all_gather_policies(local, B_total, world)
    ok = full["K"].shape[0] == B_total and torch.equal(full["uff"][:, 0, 0], torch.arange(B_total, dtype=torch.float64))
    for r in range(world):
        a, b = shard_range(B_total, r, world)
        ok = ok and bool((full["K"][a:b] == r + 1).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_policy_all_gather_world_size_2_gloo():
    """Multi-GPU path on CPU: instances are sharded in contiguous blocks and the policies are all-gathered once per tick."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def _exchange_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from bipedal_control_b200.sharding import PolicyExchange
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    class FakeMpc:   # geometry of a shard; the slab layout is the library's: [K | uff | x | u | times | events | n_nodes]
        batch, max_nodes, nx, nu = 3, 6, 22, 22
    m = FakeMpc()
    nK, nU, nX, nT = m.batch * m.max_nodes * m.nu * m.nx, m.batch * m.max_nodes * m.nu, m.batch * m.max_nodes * m.nx, m.batch * m.max_nodes
    n = nK + 2 * nU + nX + nT + (nT + m.batch + 1) // 2
    tick = {"i": 0}
    slab = lambda: torch.arange(n, dtype=torch.float64) + 1e6 * rank + 1e3 * tick["i"]
    ok = True
    for window in (False, True):
        ex = PolicyExchange(m, dist, rank, world, window=window, window_nodes=2, slab_provider=slab)
        for t in range(3):
            tick["i"] = t
            ex.before_tick()
            out = ex.after_tick()
            for r in range(world):
                part = ex.shard(out, r)
                expect = torch.arange(n, dtype=torch.float64) + 1e6 * r + 1e3 * t
                if window:
                    K = expect[:nK].view(m.batch, m.max_nodes, -1)[:, :2].reshape(-1)
                    ok = ok and torch.equal(part[:K.numel()], K) and part.numel() == 2 * m.batch * (m.nu * m.nx + 2 * m.nu + m.nx + 1)
                else:
                    ok = ok and torch.equal(part, expect)
        ok = ok and ex.describe()["collectives_per_tick"] == 1
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_policy_exchange_world_size_2_gloo():
    """The exchange object bench.py uses at N > 1 (one collective per tick over the policy slab; consumed-window variant), on CPU with gloo."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res)


def test_shard_range_covers_batch():
    from bipedal_control_b200.sharding import shard_range
    for B in (1, 7, 4096, 32768):
        for w in (1, 2, 4, 8):
            r = [shard_range(B, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == B and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


G1_URDF = "/root/reference/bipedal_robot_example/unitree_g1/g1_description/g1.urdf"


@pytest.mark.skipif(not os.path.exists(G1_URDF), reason="reference tree not mounted")
def test_g1_model_ingestion_cpp_vs_python(lib, tmp_path):
    """Second morphology: authored configs/g1/*.info + the reference's g1.urdf (sole points via the contact_frames extension)."""
    from tools.ingest import read_model
    out = str(tmp_path / "g1_cpp.model")
    cfg = os.path.join(ROOT, "configs", "g1")
    rc = lib.bmpc_convert_model(f"{cfg}/task.info".encode(), f"{cfg}/reference.info".encode(), f"{cfg}/gait.info".encode(), G1_URDF.encode(), out.encode())
    assert rc == 0, lib.bmpc_last_error(None)
    a, b = read_model(out), read_model(os.path.join(ROOT, "configs", "g1.model"))
    for k, v in b.items():
        if not isinstance(v, str):
            np.testing.assert_allclose(np.asarray(a[k], dtype=float), np.asarray(v, dtype=float), atol=1e-12, err_msg=k)
    assert b["nj"] == 12 and abs(b["total_mass"] - 32.239) < 1e-3     # 32.239 kg over 44 links


def test_g1_oracle_soles_on_ground_and_jacobians():
    from oracle.pyoracle import Oracle
    o = Oracle(os.path.join(ROOT, "configs", "g1.model"))
    assert o.nx == 24 and o.nu == 24
    x0 = o.initial_state()
    _, pos, _ = o.flow_map(x0, np.zeros(o.nu))
    assert np.abs(pos[:, 2]).max() < 1e-3
    rng = np.random.default_rng(2)
    x = x0 + rng.normal(0, 0.1, o.nx)
    u = rng.normal(0, 1.0, o.nu)
    L = o.linearize(x, u)
    eps = 1e-6
    for i in (3, 10, 13, 17, 23):
        xp, xm = x.copy(), x.copy()
        xp[i] += eps; xm[i] -= eps
        fd = (o.flow_map(xp, u)[0] - o.flow_map(xm, u)[0]) / (2 * eps)
        assert np.abs(fd - L["A"][:, i]).max() < 1e-6

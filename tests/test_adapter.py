"""include/bmpc_ocs2_adapter.hpp (BmpcSolver : ocs2::SolverBase, BmpcMpc : ocs2::MPC_BASE) against minimal OCS2 stand-ins (tests/stubs/).

CPU: the header compiles (-Wall -Wextra -Werror) against the stand-ins and compiles to nothing when OCS2 headers are absent.
GPU: a small driver built from it runs two MPC ticks the way BipedalController does (MPC_BASE::run -> SolverBase::run -> runImpl) and its
PrimalSolution / LinearController output is compared with the C ABI's Python mirror on the same inputs.
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "tests", "stubs")]
SRC = os.path.join(ROOT, "tests", "adapter", "adapter_main.cpp")
MODEL = os.path.join(ROOT, "configs", "h1.model")


def test_adapter_header_compiles_against_ocs2_interfaces():
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", *INC, SRC], check=True)


def test_adapter_header_is_inert_without_ocs2(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include <bmpc_ocs2_adapter.hpp>\n#ifndef BMPC_OCS2_ADAPTER_DISABLED\n#error "adapter should be disabled without OCS2 headers"\n#endif\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


def test_adapter_overrides_every_pure_virtual_of_the_stub_interfaces(tmp_path):
    # instantiating the classes fails to compile if a pure virtual of SolverBase / MPC_BASE is left unimplemented
    src = tmp_path / "t.cpp"
    src.write_text('#include <bmpc_ocs2_adapter.hpp>\n#include <type_traits>\nstatic_assert(!std::is_abstract<bmpc::BmpcSolver>::value && !std::is_abstract<bmpc::BmpcMpc>::value, "abstract");\n'
                   'static_assert(std::is_base_of<ocs2::SolverBase, bmpc::BmpcSolver>::value && std::is_base_of<ocs2::MPC_BASE, bmpc::BmpcMpc>::value, "bases");\nint main() { return 0; }\n')
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", *INC, str(src)], check=True)


@pytest.mark.gpu
def test_adapter_runs_like_the_python_mirror(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from bipedal_control_b200 import BatchedMpcMrtInterface
    from tools.ingest import read_model
    exe = tmp_path / "adapter_main"
    libdir = os.path.join(ROOT, "bipedal_control_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", *INC, SRC, "-o", str(exe), "-L", libdir, "-lbmpc", f"-Wl,-rpath,{libdir}"], check=True)
    m = read_model(MODEL)
    x0 = np.asarray(m["initial_state"]).copy(); x0[0] = 0.1
    et, ms = helpers.config2(22, x0, None, None)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.3, 0.0, 0.0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    inp = tmp_path / "in.txt"
    with open(inp, "w") as fh:
        fh.write(f"{len(x0)}\n" + " ".join(repr(float(v)) for v in x0) + f"\n{len(et)}\n" + " ".join(repr(float(v)) for v in et) + "\n" + " ".join(str(int(v)) for v in ms) + f"\n{len(tt)}\n")
        for k in range(len(tt)):
            fh.write(repr(float(tt[k])) + " " + " ".join(repr(float(v)) for v in ts[k]) + "\n")
    out = subprocess.run([str(exe), MODEL, str(inp)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    head = lines[0].split()
    n = int(head[1])
    assert int(head[3]) == 2 and int(head[5]) == 2            # preSolverRun called once per tick, two iterations logged
    post_events = [int(v) for v in head[head.index("events") + 1:]]
    perf = np.array([float(v) for v in lines[1].split()[1:]])
    X = np.array([[float(v) for v in ln.split()[3:3 + 22]] for ln in lines[2:2 + n]])
    U = np.array([[float(v) for v in ln.split()[3 + 22 + 1:]] for ln in lines[2:2 + n]])
    T = np.array([float(ln.split()[1]) for ln in lines[2:2 + n]])
    uq = np.array([float(v) for v in lines[2 + n].split()[1:]])
    assert lines[3 + n].startswith("reset ok 0")
    g = BatchedMpcMrtInterface(2, model_file=MODEL)            # same defaults as the adapter: dt and horizon of the task file
    g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
    g.advanceMpc(); g.advanceMpc()
    pol = g.getPolicy(0, 1)
    assert n == pol["n_nodes"][0]
    assert np.array_equal(T, pol["t"][0][:n]) and np.array_equal(X, pol["x"][0][:n]) and np.array_equal(U, pol["u"][0][:n])
    assert post_events == [k for k in range(n) if pol["events"][0][k] == 2]
    assert np.array_equal(perf, g.getPerformanceIndices()[0][3:6])
    _, ug, _ = g.evaluatePolicy(0.013, x0 + 0.01)
    assert np.abs(uq - ug[0]).max() <= 1e-9 * max(1.0, np.abs(ug[0]).max())    # LinearController::computeInput == bmpc_evaluate_policy
    g.close()

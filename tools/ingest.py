#!/usr/bin/env python3
"""Model ingestion (test infrastructure + data generation, numpy only).

Reads the reference's own robot description and OCS2 configuration files
  - URDF                       (e.g. unitree_h1/h1_description/urdf/h1_with_sole.urdf)
  - task.info / reference.info / gait.info  (Boost property-tree INFO files)
and emits the compact numeric model file (``configs/<robot>.model``) that the
GPU-side library and the CPU oracle load.  `/root/reference` does not exist on
the GPU box, so the derived numbers travel instead of the raw files.

What is restated here (reference file:line, relative to /root/reference):
  - model reduction == centroidal_model::createPinocchioInterface(urdf, jointNames)
    [UPSTREAM OCS2], called at ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:117:
    every joint that is not listed in model_settings.jointNames is treated as fixed
    and the child body inertia is lumped into the parent (at joint angle zero).
  - input cost weight R == BipedalRobotInterface::initializeInputCostWeight,
    ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:239-269.
  - INFO matrix loading with the optional `scaling` key == loadData::loadEigenMatrix
    [UPSTREAM OCS2]; used at BipedalRobotInterface.cpp:107-108, 275-276.

This file is an independent (python) second implementation of what the C++ host
library does in `bipedal_control_b200/csrc/ingest.cpp`; tests compare the two.
"""
from __future__ import annotations

import argparse
import math
import os
import re
import sys
import xml.etree.ElementTree as ET
from collections import OrderedDict

import numpy as np

# ----------------------------------------------------------------------------- INFO files


def _strip_comment(line: str) -> str:
    # Boost INFO comments start with ';' ; the reference files also use '//' trailers.
    for tok in (";", "//"):
        k = line.find(tok)
        if k >= 0:
            line = line[:k]
    return line.strip()


def parse_info(path: str) -> OrderedDict:
    """Parse the subset of Boost property-tree INFO used by the reference configs."""
    root: OrderedDict = OrderedDict()
    stack = [root]
    pending_key = None
    with open(path, "r") as fh:
        for raw in fh:
            line = _strip_comment(raw)
            if not line:
                continue
            # a line can be: "key value", "key", "{", "}", "key {"
            while line:
                if line.startswith("{"):
                    child = OrderedDict()
                    if pending_key is None:
                        raise ValueError(f"{path}: '{{' without key")
                    stack[-1][pending_key] = child
                    stack.append(child)
                    pending_key = None
                    line = line[1:].strip()
                    continue
                if line.startswith("}"):
                    stack.pop()
                    pending_key = None
                    line = line[1:].strip()
                    continue
                parts = line.split(None, 1)
                key = parts[0]
                rest = parts[1].strip() if len(parts) > 1 else ""
                if rest.startswith("{"):
                    pending_key = key
                    line = rest
                    continue
                if rest == "":
                    pending_key = key
                    stack[-1].setdefault(key, "")
                    line = ""
                else:
                    val = rest.split()[0]
                    stack[-1][key] = val
                    pending_key = key
                    line = ""
    return root


def info_get(tree, dotted: str, default=None):
    node = tree
    for k in dotted.split("."):
        if not isinstance(node, dict) or k not in node:
            return default
        node = node[k]
    return node


def info_list(tree, dotted: str):
    node = info_get(tree, dotted)
    if not isinstance(node, dict):
        return []
    out = []
    i = 0
    while f"[{i}]" in node:
        out.append(node[f"[{i}]"])
        i += 1
    return out


def info_matrix(tree, dotted: str, rows: int, cols: int) -> np.ndarray:
    node = info_get(tree, dotted)
    m = np.zeros((rows, cols))
    if not isinstance(node, dict):
        return m
    scaling = float(node.get("scaling", 1.0))
    for k, v in node.items():
        mm = re.match(r"\((\d+),(\d+)\)", k)
        if mm:
            r, c = int(mm.group(1)), int(mm.group(2))
            if r < rows and c < cols:
                m[r, c] = float(v)
    return m * scaling


# ----------------------------------------------------------------------------- URDF


def rpy_to_R(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _vec(s, n=3):
    v = [float(t) for t in s.split()]
    assert len(v) == n
    return np.array(v)


def parse_urdf(path: str):
    root = ET.parse(path).getroot()
    links = OrderedDict()
    for l in root.findall("link"):
        name = l.get("name")
        inertial = l.find("inertial")
        mass, com, I = 0.0, np.zeros(3), np.zeros((3, 3))
        if inertial is not None:
            o = inertial.find("origin")
            xyz = _vec(o.get("xyz", "0 0 0")) if o is not None else np.zeros(3)
            rpy = _vec(o.get("rpy", "0 0 0")) if o is not None else np.zeros(3)
            mass = float(inertial.find("mass").get("value"))
            it = inertial.find("inertia")
            ixx, ixy, ixz = float(it.get("ixx")), float(it.get("ixy")), float(it.get("ixz"))
            iyy, iyz, izz = float(it.get("iyy")), float(it.get("iyz")), float(it.get("izz"))
            I0 = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
            Ri = rpy_to_R(rpy)
            I = Ri @ I0 @ Ri.T
            com = xyz
        links[name] = dict(mass=mass, com=com, I=I)
    joints = OrderedDict()
    for j in root.findall("joint"):
        name = j.get("name")
        jtype = j.get("type")
        if jtype == "floating":
            continue
        o = j.find("origin")
        xyz = _vec(o.get("xyz", "0 0 0")) if o is not None else np.zeros(3)
        rpy = _vec(o.get("rpy", "0 0 0")) if o is not None else np.zeros(3)
        ax = j.find("axis")
        axis = _vec(ax.get("xyz")) if ax is not None else np.array([1.0, 0, 0])
        lim = j.find("limit")
        lo = float(lim.get("lower", "-1e9")) if lim is not None else -1e9
        hi = float(lim.get("upper", "1e9")) if lim is not None else 1e9
        joints[name] = dict(type=jtype, parent=j.find("parent").get("link"), child=j.find("child").get("link"),
                            xyz=xyz, R=rpy_to_R(rpy), axis=axis, lo=lo, hi=hi)
    return links, joints


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


class Body:
    """mass / com / inertia-about-com of a (lumped) body, expressed in its joint frame."""

    def __init__(self):
        self.m = 0.0
        self.mc = np.zeros(3)     # first moment
        self.Io = np.zeros((3, 3))  # inertia about the frame origin

    def add(self, m, c, Ic):
        # c, Ic expressed in this body's frame
        self.m += m
        self.mc += m * c
        self.Io += Ic + m * (skew(c).T @ skew(c))

    def finish(self):
        c = self.mc / self.m if self.m > 0 else np.zeros(3)
        Ic = self.Io - self.m * (skew(c).T @ skew(c))
        return self.m, c, Ic


def reduce_model(links, joints, joint_names, contact_names, root_link=None, contact_frames=None):
    """Tree with only `joint_names` movable; everything else lumped at angle zero."""
    children = {}
    child_links = set()
    for jn, j in joints.items():
        children.setdefault(j["parent"], []).append(jn)
        child_links.add(j["child"])
    if root_link is None:
        roots = [l for l in links if l not in child_links]
        # a dummy 'world' link with a floating joint is skipped by parse_urdf
        roots = [r for r in roots if r in children or links[r]["mass"] > 0]
        root_link = roots[0]

    mov = []            # list of dict(parent, R, p, axis, body)
    base_body = Body()
    contacts = {}

    def visit(link, mov_idx, R_acc, p_acc):
        """link frame expressed in the frame of movable joint mov_idx (-1 = base): x_mov = R_acc x_link + p_acc."""
        body = base_body if mov_idx < 0 else mov[mov_idx]["body"]
        L = links[link]
        if L["mass"] > 0:
            body.add(L["mass"], R_acc @ L["com"] + p_acc, R_acc @ L["I"] @ R_acc.T)
        if link in contact_names:
            contacts[link] = (mov_idx, p_acc.copy())
        for cname, (cparent, cxyz) in (contact_frames or {}).items():   # bmpc extension: contact points that are not URDF links
            if cparent == link:
                contacts[cname] = (mov_idx, R_acc @ np.asarray(cxyz) + p_acc)
        for jn in children.get(link, []):
            j = joints[jn]
            Rj = R_acc @ j["R"]
            pj = R_acc @ j["xyz"] + p_acc
            if jn in joint_names:
                assert j["type"] in ("revolute", "continuous"), jn
                mov.append(dict(name=jn, parent=mov_idx, R=Rj, p=pj, axis=j["axis"] / np.linalg.norm(j["axis"]),
                                body=Body(), lo=j["lo"], hi=j["hi"]))
                visit(j["child"], len(mov) - 1, np.eye(3), np.zeros(3))
            else:
                visit(j["child"], mov_idx, Rj, pj)

    visit(root_link, -1, np.eye(3), np.zeros(3))
    # order joints as in joint_names (OCS2 keeps URDF/Pinocchio order; for the reference robots the
    # two orders coincide - assert it so a silent permutation cannot happen)
    order = [m["name"] for m in mov]
    assert order == list(joint_names), f"joint order mismatch: {order} vs {joint_names}"
    return base_body, mov, [contacts[c] for c in contact_names]


# ----------------------------------------------------------------------------- kinematics (for R only)


def euler_zyx_to_R(e):
    z, y, x = e
    cz, sz, cy, sy, cx, sx = math.cos(z), math.sin(z), math.cos(y), math.sin(y), math.cos(x), math.sin(x)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return Rz @ Ry @ Rx


def rot_axis(axis, ang):
    K = skew(axis)
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


def fk(mov, contacts, q):
    """World placement of every joint frame and contact point; q = [xyz, zyx euler, joints]."""
    nj = len(mov)
    Rw = [None] * nj
    pw = [None] * nj
    Rb = euler_zyx_to_R(q[3:6])
    pb = np.array(q[0:3])
    for i, m in enumerate(mov):
        Rp, pp = (Rb, pb) if m["parent"] < 0 else (Rw[m["parent"]], pw[m["parent"]])
        Rw[i] = Rp @ m["R"] @ rot_axis(m["axis"], q[6 + i])
        pw[i] = Rp @ m["p"] + pp
    cpos = []
    for (par, off) in contacts:
        Rp, pp = (Rb, pb) if par < 0 else (Rw[par], pw[par])
        cpos.append(Rp @ off + pp)
    return Rb, pb, Rw, pw, cpos


def contact_joint_jacobians(mov, contacts, q):
    """3nc x nj block of the LOCAL_WORLD_ALIGNED frame Jacobians w.r.t. the leg joints."""
    nj = len(mov)
    Rb, pb, Rw, pw, cpos = fk(mov, contacts, q)
    J = np.zeros((3 * len(contacts), nj))
    for ci, (par, off) in enumerate(contacts):
        k = par
        while k >= 0:
            a = Rw[k] @ mov[k]["axis"]
            J[3 * ci:3 * ci + 3, k] = np.cross(a, cpos[ci] - pw[k])
            k = mov[k]["parent"]
    return J


# ----------------------------------------------------------------------------- compact model writer

MODE_IDS = {"FLY": 0, "LF": 1, "RF": 2, "STANCE": 3}


def build_model(task_file, reference_file, gait_file, urdf_file, name):
    task = parse_info(task_file)
    ref = parse_info(reference_file)
    gait = parse_info(gait_file) if gait_file else OrderedDict()
    joint_names = info_list(task, "model_settings.jointNames")
    contact_names = info_list(task, "model_settings.contactNames3DoF")
    links, joints = parse_urdf(urdf_file)
    cf = info_get(task, "contact_frames")
    contact_frames = {k: (v["parent"], [float(v["x"]), float(v["y"]), float(v["z"])]) for k, v in cf.items()} if isinstance(cf, dict) else {}
    base_body, mov, contacts = reduce_model(links, joints, joint_names, contact_names, contact_frames=contact_frames)
    nj = len(mov)
    nc = len(contacts)
    nx = 12 + nj
    nu = 3 * nc + nj

    out = OrderedDict()
    out["name"] = name
    out["nj"] = nj
    out["nc"] = nc
    m, c, I = base_body.finish()
    out["base_mass"] = m
    out["base_com"] = c
    out["base_inertia"] = I.reshape(-1)
    total_mass = m
    for i, mj in enumerate(mov):
        bm, bc, bI = mj["body"].finish()
        total_mass += bm
        out[f"joint{i}_name"] = mj["name"]
        out[f"joint{i}_parent"] = mj["parent"]
        out[f"joint{i}_R"] = mj["R"].reshape(-1)
        out[f"joint{i}_p"] = mj["p"]
        out[f"joint{i}_axis"] = mj["axis"]
        out[f"joint{i}_mass"] = bm
        out[f"joint{i}_com"] = bc
        out[f"joint{i}_inertia"] = bI.reshape(-1)
        out[f"joint{i}_limits"] = np.array([mj["lo"], mj["hi"]])
    for i, (par, off) in enumerate(contacts):
        out[f"contact{i}_name"] = contact_names[i]
        out[f"contact{i}_parent"] = par
        out[f"contact{i}_offset"] = off
    out["total_mass"] = total_mass

    init_state = info_matrix(task, "initialState", nx, 1).reshape(-1)
    out["initial_state"] = init_state
    Q = info_matrix(task, "Q", nx, nx)
    assert np.allclose(Q, np.diag(np.diag(Q))), "only diagonal Q is supported"
    out["Q_diag"] = np.diag(Q).copy()
    Rt = info_matrix(task, "R", 6 * nc, 6 * nc)
    assert np.allclose(Rt, np.diag(np.diag(Rt))), "only diagonal task-space R is supported"
    out["R_taskspace_diag"] = np.diag(Rt).copy()
    # BipedalRobotInterface.cpp:239-269: R_joint = J^T R_v J with J the stacked contact Jacobians at initialState
    J = contact_joint_jacobians(mov, contacts, init_state[6:])
    Rj = J.T @ Rt[3 * nc:, 3 * nc:] @ J
    out["R_force_diag"] = np.diag(Rt)[:3 * nc].copy()
    out["R_joint"] = Rj.reshape(-1)

    out["default_joint_state"] = info_matrix(ref, "defaultJointState", nj, 1).reshape(-1)
    out["com_height"] = float(info_get(ref, "comHeight"))
    out["target_displacement_velocity"] = float(info_get(ref, "targetDisplacementVelocity"))
    out["target_rotation_velocity"] = float(info_get(ref, "targetRotationVelocity"))

    out["friction_coefficient"] = float(info_get(task, "frictionConeSoftConstraint.frictionCoefficient"))
    out["barrier_mu"] = float(info_get(task, "frictionConeSoftConstraint.mu"))
    out["barrier_delta"] = float(info_get(task, "frictionConeSoftConstraint.delta"))
    # FrictionConeConstraint.h:66-67 constructor defaults
    out["friction_regularization"] = 25.0
    out["friction_gripper_force"] = 0.0
    out["friction_hessian_shift"] = 1e-6

    out["position_error_gain"] = float(info_get(task, "model_settings.positionErrorGain"))
    out["phase_transition_stance_time"] = float(info_get(task, "model_settings.phaseTransitionStanceTime"))
    out["swing_liftoff_velocity"] = float(info_get(task, "swing_trajectory_config.liftOffVelocity"))
    out["swing_touchdown_velocity"] = float(info_get(task, "swing_trajectory_config.touchDownVelocity"))
    out["swing_height"] = float(info_get(task, "swing_trajectory_config.swingHeight"))
    out["swing_time_scale"] = float(info_get(task, "swing_trajectory_config.swingTimeScale"))

    out["sqp_dt"] = float(info_get(task, "sqp.dt"))
    out["sqp_iterations"] = int(info_get(task, "sqp.sqpIteration"))
    out["sqp_delta_tol"] = float(info_get(task, "sqp.deltaTol"))
    out["sqp_g_max"] = float(info_get(task, "sqp.g_max"))
    out["sqp_g_min"] = float(info_get(task, "sqp.g_min"))
    out["mpc_time_horizon"] = float(info_get(task, "mpc.timeHorizon"))
    out["mpc_desired_frequency"] = float(info_get(task, "mpc.mpcDesiredFrequency"))
    out["centroidal_model_type"] = int(info_get(task, "centroidalModelType", 0))

    # schedules
    ims = info_list(ref, "initialModeSchedule.modeSequence")
    out["initial_mode_sequence"] = np.array([MODE_IDS[s] for s in ims], dtype=int)
    out["initial_event_times"] = np.array([float(v) for v in info_list(ref, "initialModeSchedule.eventTimes")])
    out["default_template_modes"] = np.array([MODE_IDS[s] for s in info_list(ref, "defaultModeSequenceTemplate.modeSequence")], dtype=int)
    out["default_template_times"] = np.array([float(v) for v in info_list(ref, "defaultModeSequenceTemplate.switchingTimes")])
    gaits = info_list(gait, "list")
    out["n_gaits"] = len(gaits)
    for gi, g in enumerate(gaits):
        out[f"gait{gi}_name"] = g
        out[f"gait{gi}_modes"] = np.array([MODE_IDS[s] for s in info_list(gait, f"{g}.modeSequence")], dtype=int)
        out[f"gait{gi}_times"] = np.array([float(v) for v in info_list(gait, f"{g}.switchingTimes")])
    return out


def write_model(out, path):
    with open(path, "w") as fh:
        fh.write("# bmpc compact model file v1 (derived numbers; generated by tools/ingest.py)\n")
        for k, v in out.items():
            if isinstance(v, str):
                fh.write(f"{k} s 1 {v}\n")
            elif isinstance(v, (int, np.integer)):
                fh.write(f"{k} i 1 {int(v)}\n")
            elif isinstance(v, float):
                fh.write(f"{k} d 1 {v!r}\n")
            else:
                a = np.asarray(v)
                if a.dtype.kind in "iu":
                    fh.write(f"{k} i {a.size} " + " ".join(str(int(t)) for t in a.reshape(-1)) + "\n")
                else:
                    fh.write(f"{k} d {a.size} " + " ".join(repr(float(t)) for t in a.reshape(-1)) + "\n")


def read_model(path):
    out = OrderedDict()
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            parts = line.split()
            k, t, n = parts[0], parts[1], int(parts[2])
            vals = parts[3:3 + n]
            if t == "s":
                out[k] = vals[0] if vals else ""
            elif t == "i":
                out[k] = int(vals[0]) if n == 1 and not k.endswith(("_modes", "_sequence")) else np.array([int(v) for v in vals], dtype=int)
            else:
                out[k] = float(vals[0]) if n == 1 and not k.endswith(("_times",)) else np.array([float(v) for v in vals])
    return out


REF = "/root/reference/bipedal_robot_example"
ROBOTS = {
    "g1": dict(task=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "g1", "task.info"),
               reference=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "g1", "reference.info"),
               gait=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "g1", "gait.info"),
               urdf=f"{REF}/unitree_g1/g1_description/g1.urdf"),
    "h1": dict(task=f"{REF}/unitree_h1/h1_ocs2_config/config/task/task.info",
               reference=f"{REF}/unitree_h1/h1_ocs2_config/config/command/reference.info",
               gait=f"{REF}/unitree_h1/h1_ocs2_config/config/command/gait.info",
               urdf=f"{REF}/unitree_h1/h1_description/urdf/h1_with_sole.urdf"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--robot", default="h1")
    ap.add_argument("--task")
    ap.add_argument("--reference")
    ap.add_argument("--gait")
    ap.add_argument("--urdf")
    ap.add_argument("--out")
    a = ap.parse_args()
    cfg = dict(ROBOTS.get(a.robot, {}))
    for k in ("task", "reference", "gait", "urdf"):
        if getattr(a, k):
            cfg[k] = getattr(a, k)
    out = build_model(cfg["task"], cfg["reference"], cfg.get("gait"), cfg["urdf"], a.robot)
    path = a.out or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", f"{a.robot}.model")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    write_model(out, path)
    print(f"wrote {path}: nj={out['nj']} nc={out['nc']} total_mass={out['total_mass']:.6f} base_mass={out['base_mass']:.6f}")


if __name__ == "__main__":
    sys.exit(main())

"""debug: N = 1..3 horizons with every kernel-variant combination vs the oracle (run on the GPU box)"""
import itertools, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers
from bipedal_control_b200 import BatchedMpcMrtInterface as G
from oracle.pyoracle import Oracle, build
from tools.ingest import read_model
build()
MODEL = os.path.join(ROOT, "configs", "h1.model")
m = read_model(MODEL); o = Oracle(MODEL); x0 = o.initial_state()
for n in (1, 2, 3):
    dt = 0.01; hor = n * dt
    et = np.array([-0.5, 0.02, 0.37]); ms = np.array([3, 1, 2, 3], dtype=np.int32)
    tt, ts = helpers.cmd_vel_target(x0, 0.0, (0.2, 0.1, 0.0, 0.1), 1.0, m["com_height"], m["default_joint_state"])
    o.reset(); o.set_dt_horizon(dt, hor); o.set_mode_schedule(et, ms); o.set_target(tt, ts); o.run(0.0, x0)
    io = o.info(); so = o.solution()
    print("N", n, "oracle before", io["before"], "after", io["after"], "step", io["step"])
    for lq, ric, ls in itertools.product((3, 2), (1, 0), (1, 0)):
        g = G(5, model_file=MODEL, dt=dt, time_horizon=hor)
        g.setOption("lq_mode", lq); g.setOption("riccati_mode", ric); g.setOption("ls_mode", ls)
        g.setCurrentObservation(0.0, x0); g.setTargetTrajectories(tt, ts); g.setModeSchedule(et, ms)
        g.advanceMpc()
        p = g.getPerformanceIndices()[3]; pol = g.getPolicy(3, 1)
        nn = int(pol["n_nodes"][0])
        print("  lq", lq, "ric", ric, "ls", ls, "before", p[0:3], "after", p[3:6], "step", p[6], "status", g.getStatus()[3],
              "|dx|", np.abs(pol["x"][0][:nn] - so["x"]).max(), "|du|", np.abs(pol["u"][0][:nn] - so["u"]).max(), "|dK|", np.abs(pol["K"][0][:nn] - so["K"]).max())
        g.close()

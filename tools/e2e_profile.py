import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch, numpy as np
class A: pass
run = bench.Runner(torch, "h1", "identical", 4096, 0, 0)
run.cold_start()
for _ in range(3): run.e2e_step()
m, w, s = run.mpc, run.w, run.e2e_state
T = {}
def tm(name, f):
    t0 = time.perf_counter(); r = f(); T[name] = T.get(name, 0) + time.perf_counter() - t0; return r
N = 10
for _ in range(N):
    t_next = s["t"] + bench.MPC_DT
    tm("sync0", m.synchronize)
    x_next, _, _ = tm("evaluatePolicy", lambda: m.evaluatePolicy(t_next, s["x"]))
    s["t"], s["x"] = t_next, x_next
    tm("before", run.before_tick)
    tm("setObs", lambda: m.setCurrentObservation(t_next, x_next))
    tm("setTargets", lambda: m.setTargetsFromCmdVel(w["CMD"], 1.0))
    tm("setModes", lambda: m.setModeSchedule(w["ET"], w["MS"], w["NE"]))
    tm("advance", m.advanceMpc)
    tm("after", run.after_tick)
    tm("getPerf", m.getPerformanceIndices)
for k, v in T.items(): print(f"{k:16s} {1e3 * v / N:8.3f} ms")
print("total", 1e3 * sum(T.values()) / N)

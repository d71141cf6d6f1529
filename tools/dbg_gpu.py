import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import bench
from bipedal_control_b200 import BatchedMpcMrtInterface
for B in (3, 600, 4096):
    w = bench.workload('identical', B)
    g = BatchedMpcMrtInterface(B, model_file=bench.MODEL, dt=0.01, time_horizon=1.0)
    g.setCurrentObservation(w['T0'], w['X0']); g.setTargetsFromCmdVel(w['CMD'], 1.0); g.setModeSchedule(w['ET'], w['MS'], w['NE'])
    for tick in range(4):
        g.advanceMpc()
        st = g.getStatus(); pf = g.getPerformanceIndices()
        print(B, tick, 'status uniq', np.unique(st), 'perf0', np.round(pf[0], 6), 'perf spread', np.abs(pf - pf[0]).max(), g.phaseTimes()['linesearch_trials'])
        g.shiftObservations(0.02); g.setTargetsFromCmdVel(w['CMD'], 1.0) if False else None
        t, x = g.getObservations()
        g.setCurrentObservation(t, x); g.setTargetsFromCmdVel(w['CMD'], 1.0)
    g.close()

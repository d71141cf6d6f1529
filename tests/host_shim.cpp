// Test-only shim: exposes the product's host-side GaitSchedule (bipedal_control_b200/csrc/bmpc_gait.h) to ctypes so that the
// CPU test-suite can compare it with the oracle's restatement without a GPU.  Built on demand by tests/test_host_logic.py.
#include "../bipedal_control_b200/csrc/bmpc_gait.h"
using namespace bmpc;
extern "C" {
void* shim_gait_create(int n_modes, const int* modes, int n_events, const double* events, int nt, const int* tmodes, const double* ttimes, double pts) {
  auto* g = new GaitSchedule();
  g->ms.modeSequence.assign(modes, modes + n_modes); g->ms.eventTimes.assign(events, events + n_events);
  g->tmpl.modes.assign(tmodes, tmodes + nt); g->tmpl.times.assign(ttimes, ttimes + nt + 1); g->phaseTransitionStanceTime = pts;
  return g;
}
void shim_gait_destroy(void* h) { delete static_cast<GaitSchedule*>(h); }
int shim_gait_insert(void* h, int n, const int* modes, const double* times, double start, double fin) {
  try { GaitTemplate t; t.modes.assign(modes, modes + n); t.times.assign(times, times + n + 1); static_cast<GaitSchedule*>(h)->insertModeSequenceTemplate(t, start, fin); return 0; } catch (...) { return -1; }
}
int shim_gait_get(void* h, double lo, double hi, int cap, double* et, int* ms) {
  try { const ModeSchedule& s = static_cast<GaitSchedule*>(h)->getModeSchedule(lo, hi); const int n = (int)s.eventTimes.size(); if (n > cap) return -2;
    for (int i = 0; i < n; ++i) et[i] = s.eventTimes[i]; for (int i = 0; i <= n; ++i) ms[i] = s.modeSequence[i]; return n; } catch (...) { return -1; }
}
}

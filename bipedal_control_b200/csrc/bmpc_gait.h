// Gait bookkeeping of one OCP instance on fixed-capacity arrays, callable from host and device.
//
// Semantics of ocs2_bipedal_robot/src/gait/GaitSchedule.cpp:46-137 (insertModeSequenceTemplate, getModeSchedule, tileModeSequenceTemplate) on
// [UPSTREAM] ocs2::ModeSchedule {eventTimes[n], modeSequence[n + 1]}.  The reference keeps one GaitSchedule per robot in std::vectors on the host;
// here every instance of the batch owns a slice of device arrays and the per-tick update (SwitchedModelReferenceManager::modifyReferences,
// SwitchedModelReferenceManager.cpp:62-69) runs as one thread per instance (k_gait_schedule), so a tick has no O(B) host loop.
// The same functions are compiled for the host by tests/host_shim.cpp and compared with the oracle's restatement on the CPU.
#pragma once
#include "bmpc_model.h"

#if defined(__CUDACC__)
#define BMPC_HD __host__ __device__
#else
#define BMPC_HD
#endif

namespace bmpc {

enum ModeNumber { FLY = 0, LF = 1, RF = 2, STANCE = 3 };   // gait/MotionPhaseDefinition.h:47-52
constexpr int GAIT_TMAX = 8;                               // phases of a mode-sequence template (gait.info: at most 4)
enum GaitStatus { GAIT_OK = 0, GAIT_CAPACITY = 1, GAIT_TILING_ORDER = 2 };

struct GaitTemplateArrays { int n; int modes[GAIT_TMAX]; double times[GAIT_TMAX + 1]; };

// view of one instance's schedule: n events, nm modes (nm == n + 1 between calls), capacity `cap` events / cap + 1 modes
struct GaitView {
  int cap; int* n; double* ev; int* modes;
  GaitTemplateArrays* tmpl;
};

BMPC_HD inline int gait_lower_bound(const double* a, int n, double t) {
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < t) lo = mid + 1; else hi = mid; }
  return lo;
}

// GaitSchedule::tileModeSequenceTemplate (GaitSchedule.cpp:107-137): appends startTime, then whole templates until the last event reaches finalTime,
// then the closing STANCE.  nm = number of modes currently stored.
BMPC_HD inline int gait_tile(GaitView g, int& n, int& nm, double startTime, double finalTime) {
  const GaitTemplateArrays& t = *g.tmpl;
  if (t.n == 0) return GAIT_OK;                                   // no template: the last mode continues for ever
  if (n > 0 && startTime <= g.ev[n - 1]) return GAIT_TILING_ORDER;   // the reference throws here
  if (n >= g.cap) return GAIT_CAPACITY;
  g.ev[n++] = startTime;
  while (g.ev[n - 1] < finalTime) {
    if (n + t.n > g.cap || nm + t.n + 1 > g.cap + 1) return GAIT_CAPACITY;
    for (int i = 0; i < t.n; ++i) { g.modes[nm++] = t.modes[i]; g.ev[n] = g.ev[n - 1] + (t.times[i + 1] - t.times[i]); ++n; }
  }
  if (nm > g.cap) return GAIT_CAPACITY;
  g.modes[nm++] = STANCE;
  return GAIT_OK;
}

// GaitSchedule::insertModeSequenceTemplate (GaitSchedule.cpp:46-73; called from GaitReceiver::preSolverRun, GaitReceiver.cpp:49-59)
BMPC_HD inline int gait_insert(GaitView g, const GaitTemplateArrays& tmpl, double startTime, double finalTime, double phaseTransitionStanceTime) {
  *g.tmpl = tmpl;
  int n = *g.n, nm = n + 1;
  const int index = gait_lower_bound(g.ev, n, startTime);
  if (index < n) { n = index; nm = index + 1; }                  // delete the old logic from the index
  double transition = phaseTransitionStanceTime;
  if (nm > 0 && g.modes[nm - 1] == STANCE) transition = 0.0;     // already standing: no intermediate stance phase
  if (transition > 0.0) {
    if (n >= g.cap) return GAIT_CAPACITY;
    g.ev[n++] = startTime; g.modes[nm++] = STANCE;
  }
  const int rc = gait_tile(g, n, nm, startTime + transition, finalTime);
  *g.n = n;
  return rc;
}

// GaitSchedule::getModeSchedule (GaitSchedule.cpp:78-102): drops the events before lowerBoundTime except the last one, makes the first remaining
// mode STANCE, removes the closing STANCE and re-tiles up to upperBoundTime.  (The reference erases end() - 1 of an empty vector when there is no
// event at all; here that case simply starts tiling at upperBoundTime.)
BMPC_HD inline int gait_get_mode_schedule(GaitView g, double lowerBoundTime, double upperBoundTime) {
  int n = *g.n, nm = n + 1;
  const int index = gait_lower_bound(g.ev, n, lowerBoundTime);
  if (index > 0) {
    const int drop = index - 1;
    if (drop > 0) {
      for (int i = 0; i + drop < n; ++i) g.ev[i] = g.ev[i + drop];
      for (int i = 0; i + drop < nm; ++i) g.modes[i] = g.modes[i + drop];
      n -= drop; nm -= drop;
    }
    g.modes[0] = STANCE;
  }
  const double tilingStart = n == 0 ? upperBoundTime : g.ev[n - 1];
  if (n > 0) --n;
  if (nm > 0) --nm;
  const int rc = gait_tile(g, n, nm, tilingStart, upperBoundTime);
  *g.n = n;
  return rc;
}

}  // namespace bmpc

// Host-side gait bookkeeping of one OCP instance.
// Mirrors ocs2_bipedal_robot/src/gait/GaitSchedule.cpp:38-137 (insertModeSequenceTemplate, getModeSchedule,
// tileModeSequenceTemplate) and [UPSTREAM] ocs2::ModeSchedule.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <vector>
#include "bmpc_model.h"

namespace bmpc {

enum ModeNumber { FLY = 0, LF = 1, RF = 2, STANCE = 3 };   // gait/MotionPhaseDefinition.h:47-52

struct ModeSchedule {
  std::vector<double> eventTimes;
  std::vector<int> modeSequence;
};

class GaitSchedule {
 public:
  ModeSchedule ms;
  GaitTemplate tmpl;
  double phaseTransitionStanceTime = 0.4;

  // GaitSchedule.cpp:46-73
  void insertModeSequenceTemplate(const GaitTemplate& t, double startTime, double finalTime) {
    tmpl = t;
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t index = std::lower_bound(et.begin(), et.end(), startTime) - et.begin();
    if (index < et.size()) { et.erase(et.begin() + index, et.end()); seq.erase(seq.begin() + index + 1, seq.end()); }
    double transition = phaseTransitionStanceTime;
    if (!seq.empty() && seq.back() == STANCE) transition = 0.0;
    if (transition > 0.0) { et.push_back(startTime); seq.push_back(STANCE); }
    tile(startTime + transition, finalTime);
  }
  // GaitSchedule.cpp:78-102: trims the past, forces the first remaining mode to STANCE, re-tiles up to upperBoundTime
  const ModeSchedule& getModeSchedule(double lowerBoundTime, double upperBoundTime) {
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t index = std::lower_bound(et.begin(), et.end(), lowerBoundTime) - et.begin();
    if (index > 0) {
      et.erase(et.begin(), et.begin() + index - 1);
      seq.erase(seq.begin(), seq.begin() + index - 1);
      seq.front() = STANCE;
    }
    const double tilingStart = et.empty() ? upperBoundTime : et.back();
    if (!et.empty()) et.pop_back();
    if (!seq.empty()) seq.pop_back();
    tile(tilingStart, upperBoundTime);
    return ms;
  }

 private:
  // GaitSchedule.cpp:107-137
  void tile(double startTime, double finalTime) {
    auto& et = ms.eventTimes; auto& seq = ms.modeSequence;
    const size_t n = tmpl.modes.size();
    if (n == 0) return;
    if (!et.empty() && startTime <= et.back()) throw std::runtime_error("The initial time for template-tiling is not greater than the last event time.");
    et.push_back(startTime);
    while (et.back() < finalTime)
      for (size_t i = 0; i < n; ++i) { seq.push_back(tmpl.modes[i]); et.push_back(et.back() + (tmpl.times[i + 1] - tmpl.times[i])); }
    seq.push_back(STANCE);
  }
};

}  // namespace bmpc

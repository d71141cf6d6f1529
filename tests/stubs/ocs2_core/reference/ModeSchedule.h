// test stub of ocs2_core/reference/ModeSchedule.h
#pragma once
#include <utility>
#include <vector>
#include <ocs2_core/Types.h>
namespace ocs2 {
struct ModeSchedule {
  ModeSchedule() : ModeSchedule(std::vector<scalar_t>{}, std::vector<size_t>{0}) {}
  ModeSchedule(std::vector<scalar_t> eventTimesInput, std::vector<size_t> modeSequenceInput) : eventTimes(std::move(eventTimesInput)), modeSequence(std::move(modeSequenceInput)) {}
  std::vector<scalar_t> eventTimes;
  std::vector<size_t> modeSequence;
};
}  // namespace ocs2

// Test-only shim: exposes the product's gait bookkeeping (bipedal_control_b200/csrc/bmpc_gait.h: the same functions the device kernels
// k_gait_insert / k_gait_schedule call, compiled here for the host) to ctypes so that the CPU test-suite can compare it with the oracle's
// restatement of GaitSchedule.cpp without a GPU.  Built on demand by tests/test_host_logic.py.
#include <vector>
#include "../bipedal_control_b200/csrc/bmpc_gait.h"
using namespace bmpc;
namespace {
struct Shim { int cap, n; std::vector<double> ev; std::vector<int> modes; GaitTemplateArrays tmpl; double pts;
  GaitView view() { return GaitView{cap, &n, ev.data(), modes.data(), &tmpl}; } };
GaitTemplateArrays make_template(int n, const int* modes, const double* times) {
  GaitTemplateArrays t{}; t.n = n; for (int i = 0; i < n; ++i) t.modes[i] = modes[i]; for (int i = 0; i <= n; ++i) t.times[i] = times[i]; return t;
}
}  // namespace
extern "C" {
void* shim_gait_create(int n_modes, const int* modes, int n_events, const double* events, int nt, const int* tmodes, const double* ttimes, double pts) {
  auto* g = new Shim(); g->cap = 64; g->ev.assign(g->cap, 0.0); g->modes.assign(g->cap + 1, 0);
  g->n = n_events; for (int i = 0; i < n_events; ++i) g->ev[i] = events[i]; for (int i = 0; i < n_modes; ++i) g->modes[i] = modes[i];
  g->tmpl = make_template(nt, tmodes, ttimes); g->pts = pts;
  return g;
}
void shim_gait_destroy(void* h) { delete static_cast<Shim*>(h); }
int shim_gait_insert(void* h, int n, const int* modes, const double* times, double start, double fin) {
  auto* g = static_cast<Shim*>(h);
  if (n > GAIT_TMAX) return -3;
  return -gait_insert(g->view(), make_template(n, modes, times), start, fin, g->pts);
}
int shim_gait_get(void* h, double lo, double hi, int cap, double* et, int* ms) {
  auto* g = static_cast<Shim*>(h);
  const int rc = gait_get_mode_schedule(g->view(), lo, hi);
  if (rc != GAIT_OK) return -rc;
  if (g->n > cap) return -9;
  for (int i = 0; i < g->n; ++i) et[i] = g->ev[i];
  for (int i = 0; i <= g->n; ++i) ms[i] = g->modes[i];
  return g->n;
}
}

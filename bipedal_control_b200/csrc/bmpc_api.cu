// libbmpc.so: C ABI (include/bmpc.h) + host orchestration of one batched MPC tick on one B200.
//
// Host-side mirror of ocs2::MPC_BASE::run / SolverBase::run / SqpSolver::runImpl [UPSTREAM] as driven by
// MPC_MRT_Interface::advanceMpc (bipedal_controllers/src/BipedalController.cpp:332-351), and of the MRT side
// (updatePolicy / evaluatePolicy, BipedalController.cpp:191-206) that reads the policy while the MPC thread solves.
//
// Threading model (mirrors MPC_MRT_Interface [UPSTREAM]: one MPC thread, one real-time reader, callback threads that set references):
//   * a tick reads policy buffer `cur` (warm start) and writes buffer 1 - cur on the compute stream; nothing a getter can see changes
//     until the tick is PUBLISHED (cur flips) -- by bmpc_synchronize, or lazily by the first getter that finds the tick finished
//     (MRT_BASE::updatePolicy semantics);
//   * getters (bmpc_get_policy, bmpc_evaluate_policy, bmpc_get_performance, bmpc_get_status) serve buffer `cur` on a second stream and never
//     wait for a tick in flight; a reader count per buffer keeps the next tick from overwriting a buffer that is being copied;
//   * set_* calls copy into pinned staging under the handle mutex; the tick uploads staging at its start and set_* waits for that upload
//     (microseconds) before touching staging again.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bmpc.h"
#include "bmpc_exchange.h"
#include "bmpc_gait.h"
#include "bmpc_kernels.cuh"

namespace bmpc {
void finalize_model(HostModel& m);
}

using namespace bmpc;

namespace {
std::string g_create_error;
std::mutex g_create_mtx;

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) throw CudaError(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #call); } while (0)

// All handles of a process share ONE __constant__ image of the robot model per device.  The image is tagged with a hash of its content; a tick
// whose model differs from the resident image first waits for every tick in flight on that device, then uploads its own (handles of the same
// robot never wait; handles of different robots are serialised, not corrupted).
std::mutex g_image_mtx;
unsigned long long g_image_id[64] = {};

unsigned long long fnv1a(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h ? h : 1ull;
}

struct DevPool {   // every device / pinned allocation of a handle, so that a failing bmpc_create cannot leak
  std::vector<void*> dev, host;
  template <class T> T* d(size_t n) { T* p = nullptr; n = std::max<size_t>(n, 1); CK(cudaMalloc(&p, n * sizeof(T))); dev.push_back(p); CK(cudaMemset(p, 0, n * sizeof(T))); return p; }
  template <class T> T* h(size_t n) { T* p = nullptr; n = std::max<size_t>(n, 1); CK(cudaMallocHost(&p, n * sizeof(T))); host.push_back(p); std::memset(p, 0, n * sizeof(T)); return p; }
  void release() { for (void* p : dev) cudaFree(p); for (void* p : host) cudaFreeHost(p); dev.clear(); host.clear(); }
};
}  // namespace

struct bmpc_handle {
  HostModel model; unsigned long long model_id = 0;
  int B = 0, NS = 0, ME = 0, TP = 0, nj = 0, nx = 0, nu = 0, device = 0, sqp_iterations = 1;
  double dt = 0, horizon = 0;
  double rollout_abs = 1e-5, rollout_rel = 1e-3, rollout_dt = 0.015;   // task.info:159-167 (AbsTolODE, RelTolODE, timeStep)
  cudaStream_t stream = nullptr, io_stream = nullptr;
  cudaEvent_t ev_done = nullptr, ev_inputs = nullptr;
  std::string err;
  DevPool pool;
  // synchronisation (see the header comment)
  std::mutex mtx, eval_mtx; std::condition_variable cv; int readers[2] = {0, 0};
  bool pending = false, upload_inflight = false;
  int last_rc = BMPC_OK;
  // inputs (device) and their pinned staging copies
  double *d_t0 = nullptr, *d_x0 = nullptr, *d_tgt_t = nullptr, *d_tgt_x = nullptr, *d_ev_t = nullptr;
  int *d_n_ev = nullptr, *d_ev_mode = nullptr;
  double *h_t0 = nullptr, *h_x0 = nullptr, *h_tgt_t = nullptr, *h_tgt_x = nullptr, *h_ev_t = nullptr, *h_cmd = nullptr, *d_cmd = nullptr;
  int *h_n_ev = nullptr, *h_ev_mode = nullptr;
  bool obs_dirty = false, tgt_dirty = false, sched_dirty = false, have_obs = false, have_tgt = false, have_sched = false;
  int npts = 0;
  // node grid work arrays
  double *d_st_t = nullptr, *d_st_dt = nullptr, *d_xref = nullptr, *d_zref = nullptr;
  int* d_st_mode = nullptr;
  // primal solutions (double buffered), with the mode schedule and the performance / status of the tick that produced them
  int* s_n[2] = {nullptr, nullptr}; int* s_ev[2] = {nullptr, nullptr};
  double *s_t[2] = {nullptr, nullptr}, *s_x[2] = {nullptr, nullptr}, *s_u[2] = {nullptr, nullptr}, *s_uff[2] = {nullptr, nullptr}, *s_K[2] = {nullptr, nullptr};
  int* s_nev[2] = {nullptr, nullptr}; double* s_evt[2] = {nullptr, nullptr}; int* s_evm[2] = {nullptr, nullptr};
  double* s_perf[2] = {nullptr, nullptr}; int* s_status[2] = {nullptr, nullptr};
  double* slab[2] = {nullptr, nullptr}; size_t slab_doubles = 0;
  int cur = 0; bool have_solution = false;
  // work
  double *d_lq = nullptr, *d_proj = nullptr, *d_stage = nullptr, *d_ric = nullptr, *d_dx = nullptr, *d_du = nullptr, *d_norms = nullptr, *d_alpha = nullptr;
  int* d_counters = nullptr;
  int* h_counters = nullptr;
  size_t rec = 0, prec = 0, krec = 0, srec = 0;
  int projection_mode = 1;   // 1: upstream's FullPivLU projection (default), 0: Moore-Penrose (Householder QR on per-foot compressed rows)
  // gait bookkeeping
  GaitDev gait{}; int* d_gait_rc = nullptr; bool use_gait = false;
  // stats
  int launches = 0; bool timing = false; cudaEvent_t tev[10] = {}; float phase_ms[9] = {};
  int linesearch_trials = 0, max_trials = 0, failed_instances = 0, status_or = 0;
  // multi-GPU policy exchange (bmpc_exchange_*)
  struct Exchange {
    NcclApi::Comm comm = nullptr; int rank = 0, nranks = 1; bool copy_engines = false, symmetric = false; int max_ctas = 0;
    cudaStream_t stream = nullptr; cudaEvent_t ready = nullptr, done[2] = {nullptr, nullptr}; bool started[2] = {false, false};
    void* recv[2] = {nullptr, nullptr}; void* send_slab[2] = {nullptr, nullptr}; bool nccl_mem = false;
    NcclApi::Window win_recv[2] = {nullptr, nullptr}, win_send[2] = {nullptr, nullptr};
    int last = -1; long count = 0;
  } ex;
  // scratch for policy evaluation / rollout
  double *d_default_joints = nullptr, *d_jc = nullptr;
  double *d_eval_t = nullptr, *d_eval_x = nullptr, *d_eval_xo = nullptr, *d_eval_uo = nullptr; int* d_eval_m = nullptr;
};

namespace {

Dev make_dev(bmpc_handle* h) {
  const int w = 1 - h->cur;
  Dev d;
  d.B = h->B; d.NS = h->NS; d.ME = h->ME; d.TP = h->TP; d.npts = h->npts;
  d.dt_nom = h->dt; d.horizon = h->horizon;
  d.t0 = h->d_t0; d.x0 = h->d_x0; d.tgt_t = h->d_tgt_t; d.tgt_x = h->d_tgt_x;
  d.n_ev = h->d_n_ev; d.ev_t = h->d_ev_t; d.ev_mode = h->d_ev_mode;
  d.n_nodes = h->s_n[w]; d.node_t = h->s_t[w]; d.node_ev = h->s_ev[w];
  d.st_t = h->d_st_t; d.st_dt = h->d_st_dt; d.st_mode = h->d_st_mode; d.xref = h->d_xref; d.zref = h->d_zref;
  d.p_n = h->have_solution ? h->s_n[h->cur] : nullptr; d.p_t = h->s_t[h->cur]; d.p_x = h->s_x[h->cur]; d.p_u = h->s_u[h->cur];
  d.p_ev = h->s_ev[h->cur]; d.p_uff = h->s_uff[h->cur]; d.p_K = h->s_K[h->cur];
  d.s_x = h->s_x[w]; d.s_u = h->s_u[w]; d.s_uff = h->s_uff[w]; d.s_K = h->s_K[w];
  d.lq = h->d_lq; d.proj = h->d_proj; d.stage = h->d_stage; d.ric = h->d_ric; d.jc = h->d_jc; d.dx = h->d_dx; d.du = h->d_du;
  d.perf = h->s_perf[w]; d.alpha = h->d_alpha; d.norms = h->d_norms; d.status = h->s_status[w]; d.counters = h->d_counters;
  return d;
}

void upload_inputs(bmpc_handle* h) {   // pinned staging -> device, on the compute stream; caller holds h->mtx
  cudaStream_t st = h->stream; const int B = h->B;
  bool any = false;
  if (h->obs_dirty) { CK(cudaMemcpyAsync(h->d_t0, h->h_t0, sizeof(double) * B, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(h->d_x0, h->h_x0, sizeof(double) * B * h->nx, cudaMemcpyHostToDevice, st)); h->obs_dirty = false; any = true; }
  if (h->tgt_dirty) { CK(cudaMemcpyAsync(h->d_tgt_t, h->h_tgt_t, sizeof(double) * B * h->TP, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(h->d_tgt_x, h->h_tgt_x, sizeof(double) * B * h->TP * h->nx, cudaMemcpyHostToDevice, st)); h->tgt_dirty = false; any = true; }
  if (h->sched_dirty) {
    CK(cudaMemcpyAsync(h->d_n_ev, h->h_n_ev, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_ev_t, h->h_ev_t, sizeof(double) * B * h->ME, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_ev_mode, h->h_ev_mode, sizeof(int) * B * (h->ME + 1), cudaMemcpyHostToDevice, st));
    h->sched_dirty = false; any = true;
  }
  if (any) { CK(cudaEventRecord(h->ev_inputs, st)); h->upload_inflight = true; }
}
// before the host overwrites pinned staging: the previous upload must have left it (caller holds h->mtx)
void staging_ready(bmpc_handle* h) { if (h->upload_inflight) { CK(cudaEventSynchronize(h->ev_inputs)); h->upload_inflight = false; } }

template <int NJ>
void set_kernel_attributes() {
  CK(cudaFuncSetAttribute(k_riccati_warp<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RIC_WPC * sizeof(RicWarpSmem<NJ>))));
  CK(cudaFuncSetAttribute(k_riccati_warp<NJ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k_lq_pack<NJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LqPackSmem<NJ>)));
  CK(cudaFuncSetAttribute(k_lq_pack<NJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LqPackSmem<NJ>)));
  CK(cudaFuncSetAttribute(k_policy_expand<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(PolSmem<NJ>))));
  CK(cudaFuncSetAttribute(k_forward<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FWD_WPC * sizeof(FwdSmem<NJ>))));
}

// the model image this handle's kernels read (see g_image_mtx)
void ensure_model_image(bmpc_handle* h) {
  std::lock_guard<std::mutex> lk(g_image_mtx);
  const int dev = h->device & 63;
  if (g_image_id[dev] == h->model_id) return;
  CK(cudaDeviceSynchronize());   // ticks of handles that hold another robot finish on the old image first
  CK(cudaMemcpyToSymbol(c_model, &h->model.dev, sizeof(DevModel), 0, cudaMemcpyHostToDevice));
  g_image_id[dev] = h->model_id;
}

// enqueue one tick on the compute stream; caller holds h->mtx, no tick pending, no reader on the buffer about to be written
template <int NJ>
void tick(bmpc_handle* h) {
  cudaStream_t st = h->stream;
  const int B = h->B, NS = h->NS, w = 1 - h->cur;
  auto mark = [&](int i) { if (h->timing) CK(cudaEventRecord(h->tev[i], st)); };
  h->launches = 0;
  ensure_model_image(h);
  if (h->ex.comm && h->ex.started[w]) CK(cudaStreamWaitEvent(st, h->ex.done[w], 0));   // the all-gather that reads the slab this tick overwrites
  upload_inputs(h);
  CK(cudaMemsetAsync(h->s_status[w], 0, sizeof(int) * B, st));
  CK(cudaMemsetAsync(h->d_counters, 0, sizeof(int) * CNT_N, st));
  Dev d = make_dev(h);
  const int nodes = B * NS;
  mark(0);
  if (h->use_gait) { k_gait_schedule<<<(B + 127) / 128, 128, 0, st>>>(h->gait, d); ++h->launches; }   // modifyReferences: per-instance gait tiling on the device
  // the mode schedule this policy is solved with travels with the policy buffer (evaluatePolicy reports the mode from it)
  CK(cudaMemcpyAsync(h->s_nev[w], h->d_n_ev, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(h->s_evt[w], h->d_ev_t, sizeof(double) * B * h->ME, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(h->s_evm[w], h->d_ev_mode, sizeof(int) * B * (h->ME + 1), cudaMemcpyDeviceToDevice, st));
  k_time_grid<<<(B + 127) / 128, 128, 0, st>>>(d); ++h->launches;
  k_node_setup<NJ><<<(nodes + 127) / 128, 128, 0, st>>>(d); ++h->launches;
  mark(1);
  for (int iter = 0; iter < h->sqp_iterations; ++iter) {
    constexpr int G = LqPackSmem<NJ>::G;
    const int NP = (NS + G - 1) / G;
    if (h->projection_mode == 1) k_lq_pack<NJ, true><<<(B * NP + 3) / 4, 128, sizeof(LqPackSmem<NJ>), st>>>(d);
    else k_lq_pack<NJ, false><<<(B * NP + 3) / 4, 128, sizeof(LqPackSmem<NJ>), st>>>(d);
    ++h->launches;
    if (iter == 0) mark(2);
    if (h->projection_mode == 1) k_project<NJ, true><<<(nodes + 3) / 4, 128, 0, st>>>(d);     // upstream's FullPivLU projection
    else k_project<NJ, false><<<(nodes + 3) / 4, 128, 0, st>>>(d);                            // Moore-Penrose
    ++h->launches;
    if (iter == 0) mark(3);
    k_riccati_warp<NJ><<<(B + RIC_WPC - 1) / RIC_WPC, 32 * RIC_WPC, RIC_WPC * sizeof(RicWarpSmem<NJ>), st>>>(d); ++h->launches;
    if (iter == 0) mark(4);
    k_policy_expand<NJ><<<(nodes + 3) / 4, 128, 4 * sizeof(PolSmem<NJ>), st>>>(d); ++h->launches;
    if (iter == 0) mark(5);
    k_forward<NJ><<<(B + FWD_WPC - 1) / FWD_WPC, 32 * FWD_WPC, FWD_WPC * sizeof(FwdSmem<NJ>), st>>>(d); ++h->launches;
    if (iter == 0) mark(6);
    // filter line search: device-side backtracking loop, one CTA per instance; then the accepted step
    k_linesearch<NJ><<<B, LS_THREADS, 0, st>>>(d); ++h->launches;
    k_update<NJ><<<(unsigned)(((size_t)nodes * Dims<NJ>::NX + 255) / 256), 256, 0, st>>>(d); ++h->launches;
    if (iter == 0) mark(7);
  }
  k_policy_fill<NJ><<<B, 128, 0, st>>>(d); ++h->launches;
  mark(8);
  CK(cudaMemcpyAsync(h->h_counters, h->d_counters, sizeof(int) * CNT_N, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(h->ev_done, st));
  CK(cudaGetLastError());
  h->pending = true;
}

// the tick in flight has finished: make its policy the current one (caller holds h->mtx)
void publish(bmpc_handle* h) {
  h->cur = 1 - h->cur; h->have_solution = true; h->pending = false;
  h->linesearch_trials = h->h_counters[CNT_TRIALS]; h->max_trials = h->h_counters[CNT_MAXTRIALS];
  h->failed_instances = h->h_counters[CNT_FAIL]; h->status_or = h->h_counters[CNT_STATUS];
  if (h->timing) for (int i = 0; i < 8; ++i) cudaEventElapsedTime(&h->phase_ms[i], h->tev[i], h->tev[i + 1]);
  h->last_rc = BMPC_OK;
  if (h->status_or & 32) { h->last_rc = BMPC_ERR_CAPACITY; h->err = "[bmpc] time grid overflow: more event nodes inside the horizon than max_event_nodes (status bit 32)"; }
  else if (h->status_or & 4) { h->last_rc = BMPC_ERR_INVALID; h->err = "[bmpc] swing phase without lift-off / touch-down time in the mode schedule (status bit 4)"; }
  else if (h->failed_instances > 0) { h->last_rc = BMPC_ERR_NUMERIC; h->err = "[bmpc] " + std::to_string(h->failed_instances) + " instance(s) failed numerically (see bmpc_get_status); their policies are open loop"; }
}
void try_publish(bmpc_handle* h) { if (h->pending && cudaEventQuery(h->ev_done) == cudaSuccess) publish(h); }
// wait for the tick in flight (if any) and publish it; lk is released while waiting so that getters keep being served
void wait_and_publish(bmpc_handle* h, std::unique_lock<std::mutex>& lk) {
  while (h->pending) {
    lk.unlock();
    const cudaError_t e = cudaEventSynchronize(h->ev_done);
    lk.lock();
    if (e != cudaSuccess) { h->pending = false; throw CudaError(std::string("CUDA: ") + cudaGetErrorString(e) + " while waiting for the tick"); }
    if (h->pending && cudaEventQuery(h->ev_done) == cudaSuccess) publish(h);
  }
}

struct ReadGuard {   // pins the current policy buffer while a getter copies from it
  bmpc_handle* h; int c;
  explicit ReadGuard(bmpc_handle* h_) : h(h_) {
    std::lock_guard<std::mutex> lk(h->mtx);
    try_publish(h);
    if (!h->have_solution) throw std::invalid_argument("[bmpc] no solution yet");
    c = h->cur; ++h->readers[c];
  }
  ~ReadGuard() { { std::lock_guard<std::mutex> lk(h->mtx); --h->readers[c]; } h->cv.notify_all(); }
};

GaitTemplateArrays to_template_arrays(const std::vector<int>& modes, const std::vector<double>& times) {
  if (modes.size() > (size_t)GAIT_TMAX || times.size() != modes.size() + 1) throw std::invalid_argument("[bmpc] mode-sequence template: at most 8 phases, n + 1 switching times");
  GaitTemplateArrays t{}; t.n = (int)modes.size();
  for (size_t i = 0; i < modes.size(); ++i) t.modes[i] = modes[i];
  for (size_t i = 0; i < times.size(); ++i) t.times[i] = times[i];
  return t;
}
// (re)initialises the gait bookkeeping of instances [first, first + count) from reference.info's initialModeSchedule /
// defaultModeSequenceTemplate (BipedalRobotInterface.cpp:209-234); caller holds h->mtx
void init_gaits(bmpc_handle* h, int first, int count) {
  const auto& ev = h->model.init_events; const auto& ms = h->model.init_modes;
  if ((int)ev.size() > h->ME || ms.size() != ev.size() + 1) throw std::invalid_argument("[bmpc] initialModeSchedule does not fit max_events");
  double* d_ev = nullptr; int* d_ms = nullptr;
  CK(cudaMalloc(&d_ev, sizeof(double) * std::max<size_t>(ev.size(), 1))); CK(cudaMalloc(&d_ms, sizeof(int) * ms.size()));
  CK(cudaMemcpyAsync(d_ev, ev.data(), sizeof(double) * ev.size(), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_ms, ms.data(), sizeof(int) * ms.size(), cudaMemcpyHostToDevice, h->stream));
  k_gait_init<<<(count + 127) / 128, 128, 0, h->stream>>>(h->gait, first, count, (int)ev.size(), d_ev, d_ms, to_template_arrays(h->model.default_template.modes, h->model.default_template.times));
  CK(cudaGetLastError()); CK(cudaStreamSynchronize(h->stream));
  cudaFree(d_ev); cudaFree(d_ms);
}

int fail(bmpc_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else { std::lock_guard<std::mutex> lk(g_create_mtx); g_create_error = msg; }
  return code;
}

#define API_BEGIN try {
#define API_END(h) } catch (const CudaError& e) { return fail(h, BMPC_ERR_CUDA, e.what()); } \
  catch (const std::invalid_argument& e) { return fail(h, BMPC_ERR_INVALID, e.what()); } \
  catch (const std::length_error& e) { return fail(h, BMPC_ERR_CAPACITY, e.what()); } \
  catch (const std::exception& e) { return fail(h, BMPC_ERR_INVALID, e.what()); }

void set_slab_pointers(bmpc_handle* h, int i, double* slab) {
  const size_t B = h->B, NS = h->NS, nx = h->nx, nu = h->nu;
  const size_t nK = B * NS * nu * nx, nU = B * NS * nu, nX = B * NS * nx, nT = B * NS;
  h->slab[i] = slab;
  h->s_K[i] = slab; h->s_uff[i] = slab + nK; h->s_x[i] = h->s_uff[i] + nU; h->s_u[i] = h->s_x[i] + nX; h->s_t[i] = h->s_u[i] + nU;
  h->s_ev[i] = reinterpret_cast<int*>(h->s_t[i] + nT); h->s_n[i] = h->s_ev[i] + B * NS;
}
// Tear-down of the policy exchange.  With symmetric windows every rank has the peers' slabs mapped, and NCCL's window / communicator tear-down touches
// that peer memory: a rank that released its buffers while a slower peer was still deregistering made the peer fault (illegal memory access after
// bmpc_destroy on the slower rank).  So the release is fenced by two tiny collectives on the same communicator: (1) every rank has entered the release
// (no exchange in flight anywhere), (2) every rank has deregistered its windows -- only then is memory freed and the communicator destroyed.  Like
// ncclCommDestroy itself, bmpc_exchange_destroy / bmpc_destroy of a handle with an exchange is therefore a collective call.
void exchange_fence(bmpc_handle* h) {
  auto& ex = h->ex; NcclApi& N = nccl_api();
  if (!ex.comm || !ex.stream || ex.nranks <= 1) return;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * (size_t)(ex.nranks + 1)) != cudaSuccess) return;
  cudaMemsetAsync(buf, 0, sizeof(double) * (size_t)(ex.nranks + 1), ex.stream);
  N.AllGather(buf + ex.nranks, buf, 1, NcclApi::kFloat64, ex.comm, ex.stream);
  cudaStreamSynchronize(ex.stream);
  cudaFree(buf);
}
void exchange_release(bmpc_handle* h) {
  auto& ex = h->ex; NcclApi& N = nccl_api();
  if (ex.stream) cudaStreamSynchronize(ex.stream);
  if (ex.comm && ex.symmetric) exchange_fence(h);
  if (ex.comm) {
    for (int i = 0; i < 2; ++i) {
      if (ex.win_recv[i] && N.CommWindowDeregister) N.CommWindowDeregister(ex.comm, ex.win_recv[i]);
      if (ex.win_send[i] && N.CommWindowDeregister) N.CommWindowDeregister(ex.comm, ex.win_send[i]);
    }
  }
  if (ex.comm && ex.symmetric) exchange_fence(h);
  for (int i = 0; i < 2; ++i) if (ex.send_slab[i]) {
    // the policy slabs live in NCCL memory: move them back into plain device memory before it is released
    const size_t bytes = h->slab_doubles * sizeof(double);
    double* p = nullptr;
    if (cudaMalloc(&p, bytes) == cudaSuccess) { cudaMemcpy(p, ex.send_slab[i], bytes, cudaMemcpyDeviceToDevice); h->pool.dev.push_back(p); set_slab_pointers(h, i, p); }
  }
  if (ex.comm) { N.CommDestroy(ex.comm); ex.comm = nullptr; }   // before the NCCL memory goes: the communicator still has it mapped / registered
  for (int i = 0; i < 2; ++i) if (ex.recv[i]) { if (ex.nccl_mem) N.MemFree(ex.recv[i]); else cudaFree(ex.recv[i]); ex.recv[i] = nullptr; }
  for (int i = 0; i < 2; ++i) if (ex.send_slab[i]) { N.MemFree(ex.send_slab[i]); ex.send_slab[i] = nullptr; }
  if (ex.ready) cudaEventDestroy(ex.ready);
  for (auto& e : ex.done) if (e) cudaEventDestroy(e);
  if (ex.stream) cudaStreamDestroy(ex.stream);
  ex = bmpc_handle::Exchange();
}

void destroy_impl(bmpc_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->io_stream) cudaStreamSynchronize(h->io_stream);
  exchange_release(h);
  h->pool.release();
  for (auto& e : h->tev) if (e) cudaEventDestroy(e);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  if (h->ev_inputs) cudaEventDestroy(h->ev_inputs);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->io_stream) cudaStreamDestroy(h->io_stream);
  delete h;
}

}  // namespace

extern "C" {

const char* bmpc_last_error(const bmpc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int bmpc_create(const bmpc_config* cfg, bmpc_handle** out) {
  bmpc_handle* h = nullptr;
  try {
    if (!cfg || !out) throw std::invalid_argument("[bmpc] null config");
    if (cfg->batch <= 0) throw std::invalid_argument("[bmpc] batch must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw CudaError("CUDA: no device available (libbmpc has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) throw std::invalid_argument("[bmpc] invalid device ordinal");
    CK(cudaSetDevice(cfg->device));
    h = new bmpc_handle();
    if (cfg->model_file && cfg->model_file[0]) h->model = load_compact_model(cfg->model_file);
    else if (cfg->task_file && cfg->urdf_file && cfg->reference_file)
      h->model = load_reference_files(cfg->task_file, cfg->reference_file, cfg->gait_file ? cfg->gait_file : "", cfg->urdf_file);
    else throw std::invalid_argument("[bmpc] either model_file or task_file + reference_file + urdf_file must be given");
    h->model_id = fnv1a(&h->model.dev, sizeof(DevModel));
    h->device = cfg->device;
    h->B = cfg->batch; h->nj = h->model.nj; h->nx = h->model.nx; h->nu = h->model.nu;
    h->dt = cfg->dt > 0 ? cfg->dt : h->model.sqp_dt;
    h->horizon = cfg->time_horizon > 0 ? cfg->time_horizon : h->model.time_horizon;
    h->ME = cfg->max_events > 0 ? cfg->max_events : 40;
    h->TP = cfg->max_target_points > 0 ? cfg->max_target_points : 4;
    h->sqp_iterations = cfg->sqp_iterations > 0 ? cfg->sqp_iterations : h->model.sqp_iterations;
    const int max_event_nodes = cfg->max_event_nodes > 0 ? cfg->max_event_nodes : 16;
    h->NS = (int)std::ceil(h->horizon / h->dt - 1e-9) + 1 + max_event_nodes;   // nominal grid + event nodes inside the horizon (two per event off the grid, one per event on it)
    const size_t B = h->B, NS = h->NS, nx = h->nx, nu = h->nu;
    if (h->nj == 10) { h->rec = Dims<10>::REC; h->prec = Dims<10>::PREC; h->krec = RDims<10>::KREC; h->srec = SDims<10>::SREC; set_kernel_attributes<10>(); }
    else { h->rec = Dims<12>::REC; h->prec = Dims<12>::PREC; h->krec = RDims<12>::KREC; h->srec = SDims<12>::SREC; set_kernel_attributes<12>(); }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->io_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_inputs, cudaEventDisableTiming));
    for (auto& e : h->tev) CK(cudaEventCreate(&e));
    DevPool& P = h->pool;
    h->d_t0 = P.d<double>(B); h->d_x0 = P.d<double>(B * nx); h->d_tgt_t = P.d<double>(B * h->TP); h->d_tgt_x = P.d<double>(B * h->TP * nx);
    h->d_n_ev = P.d<int>(B); h->d_ev_t = P.d<double>(B * h->ME); h->d_ev_mode = P.d<int>(B * (h->ME + 1));
    h->h_cmd = P.h<double>(B * 4); h->d_cmd = P.d<double>(B * 4); h->h_t0 = P.h<double>(B); h->h_x0 = P.h<double>(B * nx); h->h_tgt_t = P.h<double>(B * h->TP); h->h_tgt_x = P.h<double>(B * h->TP * nx);
    h->h_n_ev = P.h<int>(B); h->h_ev_t = P.h<double>(B * h->ME); h->h_ev_mode = P.h<int>(B * (h->ME + 1));
    h->d_st_t = P.d<double>(B * NS); h->d_st_dt = P.d<double>(B * NS); h->d_st_mode = P.d<int>(B * NS);
    h->d_xref = P.d<double>(B * NS * nx); h->d_zref = P.d<double>(B * NS * 4);
    for (int i = 0; i < 2; ++i) {
      // one contiguous slab per policy buffer: [K | uff | x | u | t | events | n_nodes], so that the multi-GPU exchange is ONE collective
      const size_t nK = B * NS * nu * nx, nU = B * NS * nu, nX = B * NS * nx, nT = B * NS;
      const size_t ints = (B * NS + B + 1) / 2;   // events + n_nodes, in units of doubles
      h->slab_doubles = nK + 2 * nU + nX + nT + ints;
      set_slab_pointers(h, i, P.d<double>(h->slab_doubles));
      h->s_nev[i] = P.d<int>(B); h->s_evt[i] = P.d<double>(B * h->ME); h->s_evm[i] = P.d<int>(B * (h->ME + 1));
      h->s_perf[i] = P.d<double>(B * 8); h->s_status[i] = P.d<int>(B);
    }
    h->d_lq = P.d<double>(B * NS * h->rec); h->d_proj = P.d<double>(B * NS * h->prec); h->d_stage = P.d<double>(B * NS * h->srec); h->d_ric = P.d<double>(B * NS * h->krec);
    h->d_dx = P.d<double>(B * NS * nx); h->d_du = P.d<double>(B * NS * nu);
    h->d_norms = P.d<double>(B * 2); h->d_alpha = P.d<double>(B);
    h->d_counters = P.d<int>(CNT_N); h->h_counters = P.h<int>(CNT_N);
    {  // static entries of the projected stage records (identity rows / columns of At); everything else was zeroed by the pool
      const size_t nrec = B * NS;
      if (h->nj == 10) k_stage_static<10><<<(unsigned)((nrec + 255) / 256), 256>>>(h->d_stage, nrec);
      else k_stage_static<12><<<(unsigned)((nrec + 255) / 256), 256>>>(h->d_stage, nrec);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    }
    h->d_default_joints = P.d<double>(MAXJ);
    CK(cudaMemcpy(h->d_default_joints, h->model.default_joint_state.data(), sizeof(double) * h->nj, cudaMemcpyHostToDevice));
    {  // packed per-joint constants for the lane = joint kernels
      std::vector<double> jc((size_t)MAXJ * 28, 0.0); const DevModel& dm = h->model.dev;
      for (int j = 0; j < h->nj; ++j) {
        double* p = jc.data() + (size_t)j * 28;
        for (int i = 0; i < 9; ++i) { p[i] = dm.Rj[j][i]; p[19 + i] = dm.inertia[j][i]; }
        for (int i = 0; i < 3; ++i) { p[9 + i] = dm.pj[j][i]; p[12 + i] = dm.axis[j][i]; p[16 + i] = dm.com[j][i]; }
        p[15] = dm.mass[j];
      }
      h->d_jc = P.d<double>((size_t)MAXJ * 28);
      CK(cudaMemcpy(h->d_jc, jc.data(), sizeof(double) * jc.size(), cudaMemcpyHostToDevice));
    }
    h->d_eval_t = P.d<double>(B); h->d_eval_x = P.d<double>(B * nx); h->d_eval_xo = P.d<double>(B * nx); h->d_eval_uo = P.d<double>(B * nu); h->d_eval_m = P.d<int>(B);
    h->gait.B = h->B; h->gait.cap = h->ME;
    h->gait.n = P.d<int>(B); h->gait.ev = P.d<double>(B * h->ME); h->gait.modes = P.d<int>(B * (h->ME + 1)); h->gait.tmpl = P.d<GaitTemplateArrays>(B);
    h->d_gait_rc = P.d<int>(1);
    init_gaits(h, 0, h->B);
    *out = h;
    return BMPC_OK;
  } catch (const CudaError& e) { destroy_impl(h); return fail(nullptr, BMPC_ERR_CUDA, e.what()); }
  catch (const std::invalid_argument& e) { destroy_impl(h); return fail(nullptr, BMPC_ERR_INVALID, e.what()); }
  catch (const std::exception& e) { destroy_impl(h); return fail(nullptr, BMPC_ERR_INVALID, e.what()); }
}

void bmpc_destroy(bmpc_handle* h) { destroy_impl(h); }

int bmpc_get_dims(const bmpc_handle* h, int* nx, int* nu, int* batch, int* max_nodes) {
  if (!h) return BMPC_ERR_INVALID;
  if (nx) *nx = h->nx; if (nu) *nu = h->nu; if (batch) *batch = h->B; if (max_nodes) *max_nodes = h->NS;
  return BMPC_OK;
}
int bmpc_get_initial_state(const bmpc_handle* h, double* x) { if (!h || !x) return BMPC_ERR_INVALID; for (int i = 0; i < h->nx; ++i) x[i] = h->model.initial_state[i]; return BMPC_OK; }
int bmpc_export_model(const bmpc_handle* h, const char* path) {
  bmpc_handle* hh = const_cast<bmpc_handle*>(h);
  API_BEGIN if (!h || !path) throw std::invalid_argument("[bmpc] null argument"); save_compact_model(h->model, path); return BMPC_OK; API_END(hh)
}

// host-only: ingest the reference's own files and write the compact model file (no GPU needed)
int bmpc_convert_model(const char* task_file, const char* reference_file, const char* gait_file, const char* urdf_file, const char* out_path) {
  try {
    if (!task_file || !reference_file || !urdf_file || !out_path) throw std::invalid_argument("[bmpc] null argument");
    HostModel m = load_reference_files(task_file, reference_file, gait_file ? gait_file : "", urdf_file);
    save_compact_model(m, out_path);
    return BMPC_OK;
  } catch (const std::exception& e) { return fail(nullptr, BMPC_ERR_INVALID, e.what()); }
}

int bmpc_reset(bmpc_handle* h, int instance) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  if (instance >= h->B) throw std::invalid_argument("[bmpc] instance out of range");
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);
  if (instance < 0) { h->have_solution = false; init_gaits(h, 0, h->B); }
  else {
    // one instance: its warm start is dropped by emptying its previous solution (the initializer is used for every node of the next tick)
    if (h->have_solution) { CK(cudaMemsetAsync(h->s_n[h->cur] + instance, 0, sizeof(int), h->stream)); CK(cudaStreamSynchronize(h->stream)); }
    init_gaits(h, instance, 1);
  }
  return BMPC_OK; API_END(h)
}

int bmpc_set_observations(bmpc_handle* h, const double* t, const double* x) {
  API_BEGIN if (!h || !t || !x) throw std::invalid_argument("[bmpc] null argument");
  std::lock_guard<std::mutex> lk(h->mtx);
  staging_ready(h);
  std::memcpy(h->h_t0, t, sizeof(double) * h->B); std::memcpy(h->h_x0, x, sizeof(double) * h->B * h->nx);
  h->obs_dirty = true; h->have_obs = true; return BMPC_OK; API_END(h)
}
int bmpc_set_observations_device(bmpc_handle* h, const double* t, const double* x) {
  API_BEGIN if (!h || !t || !x) throw std::invalid_argument("[bmpc] null argument");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  CK(cudaMemcpyAsync(h->d_t0, t, sizeof(double) * h->B, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->d_x0, x, sizeof(double) * h->B * h->nx, cudaMemcpyDeviceToDevice, h->stream));
  h->obs_dirty = false; h->have_obs = false; return BMPC_OK; API_END(h)
}
int bmpc_set_target_trajectories(bmpc_handle* h, int npts, const double* times, const double* states) {
  API_BEGIN if (!h || !times || !states) throw std::invalid_argument("[bmpc] null argument");
  if (npts < 1 || npts > h->TP) throw std::length_error("[bmpc] number of target points exceeds max_target_points");
  std::lock_guard<std::mutex> lk(h->mtx);
  staging_ready(h);
  for (int b = 0; b < h->B; ++b) {
    std::memcpy(h->h_tgt_t + (size_t)b * h->TP, times + (size_t)b * npts, sizeof(double) * npts);
    std::memcpy(h->h_tgt_x + (size_t)b * h->TP * h->nx, states + (size_t)b * npts * h->nx, sizeof(double) * npts * h->nx);
  }
  h->npts = npts; h->tgt_dirty = true; h->have_tgt = true; return BMPC_OK; API_END(h)
}
int bmpc_set_target_trajectories_device(bmpc_handle* h, int npts, const double* times, const double* states) {
  API_BEGIN if (!h || !times || !states) throw std::invalid_argument("[bmpc] null argument");
  if (npts < 1 || npts > h->TP) throw std::length_error("[bmpc] number of target points exceeds max_target_points");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  CK(cudaMemcpy2DAsync(h->d_tgt_t, sizeof(double) * h->TP, times, sizeof(double) * npts, sizeof(double) * npts, h->B, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->d_tgt_x, sizeof(double) * h->TP * h->nx, states, sizeof(double) * npts * h->nx, sizeof(double) * npts * h->nx, h->B, cudaMemcpyDeviceToDevice, h->stream));
  h->npts = npts; h->tgt_dirty = false; h->have_tgt = true; return BMPC_OK; API_END(h)
}
// TargetTrajectoriesPublisher.cpp:41-99
int bmpc_set_targets_from_cmd_vel(bmpc_handle* h, const double* cmd, double time_to_target) {
  // cmdVelToTargetTrajectories (TargetTrajectoriesPublisher.cpp:76-99) for every instance, from the observation set before this call.
  API_BEGIN if (!h || !cmd) throw std::invalid_argument("[bmpc] null argument");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  if (!h->have_obs) throw std::invalid_argument("[bmpc] set observations (host) before bmpc_set_targets_from_cmd_vel");
  if (h->TP < 2) throw std::length_error("[bmpc] max_target_points < 2");
  staging_ready(h);
  const int B = h->B;
  try_publish(h);
  if (!h->pending) {
    // no tick in flight: the same device kernel as bmpc_set_targets_from_cmd_vel_device; only the 4 command values per instance cross PCIe
    cudaStream_t st = h->stream;
    std::memcpy(h->h_cmd, cmd, sizeof(double) * 4 * (size_t)B);
    upload_inputs(h);   // the observation the command refers to
    CK(cudaMemcpyAsync(h->d_cmd, h->h_cmd, sizeof(double) * 4 * (size_t)B, cudaMemcpyHostToDevice, st));
    if (h->nj == 10) k_cmd_vel_targets<10><<<(B + 127) / 128, 128, 0, st>>>(B, h->TP, h->d_t0, h->d_x0, h->d_cmd, time_to_target, h->model.com_height, h->d_default_joints, h->d_tgt_t, h->d_tgt_x);
    else k_cmd_vel_targets<12><<<(B + 127) / 128, 128, 0, st>>>(B, h->TP, h->d_t0, h->d_x0, h->d_cmd, time_to_target, h->model.com_height, h->d_default_joints, h->d_tgt_t, h->d_tgt_x);
    CK(cudaEventRecord(h->ev_inputs, st)); h->upload_inflight = true;   // h_cmd is pinned staging like the others
    h->npts = 2; h->tgt_dirty = false; h->have_tgt = true; CK(cudaGetLastError());
    return BMPC_OK;
  }
  // a tick is in flight (MRT use: the RT thread feeds the next tick while the MPC thread solves): nothing may queue behind it here, or the next
  // set_* would wait for the whole tick in staging_ready -- the targets are built on the host into the staging buffers and travel with the next tick
  const int nx = h->nx, nj = h->nj;
  for (int b = 0; b < B; ++b) {
    const double* x = h->h_x0 + (size_t)b * nx; const double* c = cmd + (size_t)b * 4;
    const double z = x[9], y = x[10], r = x[11];
    const double cz = std::cos(z), sz = std::sin(z), cy = std::cos(y), sy = std::sin(y), cx = std::cos(r), sx = std::sin(r);
    const double R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, -sy, cy * sx, cy * cx};
    const double vr[3] = {R[0] * c[0] + R[1] * c[1] + R[2] * c[2], R[3] * c[0] + R[4] * c[1] + R[5] * c[2], R[6] * c[0] + R[7] * c[1] + R[8] * c[2]};
    double* s0 = h->h_tgt_x + (size_t)b * h->TP * nx; double* s1 = s0 + nx;
    std::fill(s0, s0 + 2 * nx, 0.0);
    s0[0] = s1[0] = vr[0]; s0[1] = s1[1] = vr[1]; s0[2] = s1[2] = vr[2];
    s0[6] = x[6]; s0[7] = x[7]; s0[8] = h->model.com_height; s0[9] = x[9];
    s1[6] = x[6] + vr[0] * time_to_target; s1[7] = x[7] + vr[1] * time_to_target; s1[8] = h->model.com_height; s1[9] = x[9] + c[3] * time_to_target;
    for (int j = 0; j < nj; ++j) { s0[12 + j] = h->model.default_joint_state[j]; s1[12 + j] = h->model.default_joint_state[j]; }
    h->h_tgt_t[(size_t)b * h->TP] = h->h_t0[b]; h->h_tgt_t[(size_t)b * h->TP + 1] = h->h_t0[b] + time_to_target;
  }
  h->npts = 2; h->tgt_dirty = true; h->have_tgt = true; return BMPC_OK; API_END(h)
}
int bmpc_set_mode_schedules(bmpc_handle* h, int stride, const int* n_events, const double* event_times, const int* mode_sequence) {
  API_BEGIN if (!h || !n_events || !event_times || !mode_sequence) throw std::invalid_argument("[bmpc] null argument");
  std::lock_guard<std::mutex> lk(h->mtx);
  for (int b = 0; b < h->B; ++b) if (n_events[b] < 0 || n_events[b] > h->ME || n_events[b] > stride) throw std::length_error("[bmpc] mode schedule exceeds max_events");
  staging_ready(h);
  if (stride == h->ME) {   // same row length as the staging arrays: three block copies (entries beyond n_events are never read)
    std::memcpy(h->h_n_ev, n_events, sizeof(int) * (size_t)h->B);
    std::memcpy(h->h_ev_t, event_times, sizeof(double) * (size_t)h->B * h->ME);
    std::memcpy(h->h_ev_mode, mode_sequence, sizeof(int) * (size_t)h->B * (h->ME + 1));
  } else
  for (int b = 0; b < h->B; ++b) {
    const int ne = n_events[b];
    h->h_n_ev[b] = ne;
    std::memcpy(h->h_ev_t + (size_t)b * h->ME, event_times + (size_t)b * stride, sizeof(double) * ne);
    std::memcpy(h->h_ev_mode + (size_t)b * (h->ME + 1), mode_sequence + (size_t)b * (stride + 1), sizeof(int) * (ne + 1));
  }
  h->sched_dirty = true; h->have_sched = true; h->use_gait = false; return BMPC_OK; API_END(h)
}
int bmpc_set_mode_schedules_device(bmpc_handle* h, int stride, const int* n_events, const double* event_times, const int* mode_sequence) {
  API_BEGIN if (!h || !n_events || !event_times || !mode_sequence) throw std::invalid_argument("[bmpc] null argument");
  if (stride > h->ME) throw std::length_error("[bmpc] mode schedule stride exceeds max_events");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  CK(cudaMemcpyAsync(h->d_n_ev, n_events, sizeof(int) * h->B, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->d_ev_t, sizeof(double) * h->ME, event_times, sizeof(double) * stride, sizeof(double) * stride, h->B, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpy2DAsync(h->d_ev_mode, sizeof(int) * (h->ME + 1), mode_sequence, sizeof(int) * (stride + 1), sizeof(int) * (stride + 1), h->B, cudaMemcpyDeviceToDevice, h->stream));
  h->sched_dirty = false; h->have_sched = true; h->use_gait = false; return BMPC_OK; API_END(h)
}

int bmpc_gait_insert(bmpc_handle* h, int instance, int n_modes, const int* modes, const double* switching_times, double start_time, double final_time) {
  API_BEGIN if (!h || !modes || !switching_times || n_modes <= 0) throw std::invalid_argument("[bmpc] null argument");
  if (instance >= h->B) throw std::invalid_argument("[bmpc] instance out of range");
  const GaitTemplateArrays t = to_template_arrays(std::vector<int>(modes, modes + n_modes), std::vector<double>(switching_times, switching_times + n_modes + 1));
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);   // GaitReceiver's receivedGaitMutex_ (GaitReceiver.cpp:52,65); stream order puts it after the tick in flight
  const int first = instance < 0 ? 0 : instance, count = instance < 0 ? h->B : 1;
  int rc = 0;
  CK(cudaMemsetAsync(h->d_gait_rc, 0, sizeof(int), h->stream));
  k_gait_insert<<<(count + 127) / 128, 128, 0, h->stream>>>(h->gait, first, count, t, start_time, final_time, h->model.phase_transition_stance_time, h->d_gait_rc);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&rc, h->d_gait_rc, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  lk.unlock();                               // getters are not held up while the stream drains
  CK(cudaStreamSynchronize(h->stream));
  if (rc == GAIT_CAPACITY) throw std::length_error("[bmpc] gait schedule exceeds max_events");
  if (rc == GAIT_TILING_ORDER) throw std::invalid_argument("[bmpc] The initial time for template-tiling is not greater than the last event time.");   // GaitSchedule.cpp:118-120 throws the same
  return BMPC_OK; API_END(h)
}
int bmpc_gait_insert_named(bmpc_handle* h, int instance, const char* gait_name, double start_time, double final_time) {
  if (!h || !gait_name) return BMPC_ERR_INVALID;
  for (const auto& g : h->model.gaits)
    if (g.name == gait_name) return bmpc_gait_insert(h, instance, (int)g.modes.size(), g.modes.data(), g.times.data(), start_time, final_time);
  return fail(h, BMPC_ERR_INVALID, std::string("[bmpc] unknown gait '") + gait_name + "'");
}
int bmpc_use_gait_schedule(bmpc_handle* h, int enable) { if (!h) return BMPC_ERR_INVALID; std::lock_guard<std::mutex> lk(h->mtx); h->use_gait = enable != 0; if (h->use_gait) h->sched_dirty = false; return BMPC_OK; }
int bmpc_gait_peek(const bmpc_handle* hc, int instance, int cap, double* event_times, int* mode_sequence) {
  bmpc_handle* h = const_cast<bmpc_handle*>(hc);
  API_BEGIN if (!h || instance < 0 || instance >= h->B || !event_times || !mode_sequence) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  int n = 0;
  std::vector<double> ev(h->ME); std::vector<int> ms(h->ME + 1);
  CK(cudaMemcpyAsync(&n, h->gait.n + instance, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ev.data(), h->gait.ev + (size_t)instance * h->ME, sizeof(double) * h->ME, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(ms.data(), h->gait.modes + (size_t)instance * (h->ME + 1), sizeof(int) * (h->ME + 1), cudaMemcpyDeviceToHost, h->stream));
  lk.unlock();
  CK(cudaStreamSynchronize(h->stream));
  if (n > cap) return BMPC_ERR_CAPACITY;
  for (int i = 0; i < n; ++i) event_times[i] = ev[i];
  for (int i = 0; i <= n; ++i) mode_sequence[i] = ms[i];
  return n; API_END(h)
}

int bmpc_advance_async(bmpc_handle* h) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);                                         // one tick at a time (MPC_BASE::run is not re-entrant either)
  if (!h->have_tgt) throw std::invalid_argument("[bmpc] target trajectories not set");
  if (!h->have_sched && !h->use_gait) throw std::invalid_argument("[bmpc] mode schedules not set (bmpc_set_mode_schedules or bmpc_use_gait_schedule)");
  const int w = 1 - h->cur;
  h->cv.wait(lk, [&] { return h->readers[w] == 0; });              // nobody is still copying the buffer this tick overwrites
  if (h->nj == 10) tick<10>(h); else tick<12>(h);
  return BMPC_OK; API_END(h)
}
int bmpc_synchronize(bmpc_handle* h) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);
  return h->last_rc; API_END(h)
}
int bmpc_advance(bmpc_handle* h) { const int rc = bmpc_advance_async(h); return rc != BMPC_OK ? rc : bmpc_synchronize(h); }
int bmpc_poll(bmpc_handle* h) {
  if (!h) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  try_publish(h);
  return h->pending ? 1 : 0;
}

int bmpc_get_policy(bmpc_handle* h, int first, int count, int* n_nodes, double* times, int* events, double* x, double* u, double* uff, double* K) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  if (first < 0 || count < 0 || first + count > h->B) throw std::invalid_argument("[bmpc] instance range out of bounds");
  CK(cudaSetDevice(h->device));
  ReadGuard g(h);
  const int c = g.c; const size_t NS = h->NS, nx = h->nx, nu = h->nu, f = first, n = count;
  std::lock_guard<std::mutex> io(h->eval_mtx);   // one user of the io stream at a time
  cudaStream_t st = h->io_stream;
  if (n_nodes) CK(cudaMemcpyAsync(n_nodes, h->s_n[c] + f, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
  if (times) CK(cudaMemcpyAsync(times, h->s_t[c] + f * NS, sizeof(double) * n * NS, cudaMemcpyDeviceToHost, st));
  if (events) CK(cudaMemcpyAsync(events, h->s_ev[c] + f * NS, sizeof(int) * n * NS, cudaMemcpyDeviceToHost, st));
  if (x) CK(cudaMemcpyAsync(x, h->s_x[c] + f * NS * nx, sizeof(double) * n * NS * nx, cudaMemcpyDeviceToHost, st));
  if (u) CK(cudaMemcpyAsync(u, h->s_u[c] + f * NS * nu, sizeof(double) * n * NS * nu, cudaMemcpyDeviceToHost, st));
  if (uff) CK(cudaMemcpyAsync(uff, h->s_uff[c] + f * NS * nu, sizeof(double) * n * NS * nu, cudaMemcpyDeviceToHost, st));
  if (K) CK(cudaMemcpyAsync(K, h->s_K[c] + f * NS * nu * nx, sizeof(double) * n * NS * nu * nx, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return BMPC_OK; API_END(h)
}
int bmpc_get_device_view(bmpc_handle* h, bmpc_device_view* v) {
  if (!h || !v) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  try_publish(h);
  if (!h->have_solution) return BMPC_ERR_INVALID;
  const int c = h->cur;
  v->n_nodes = h->s_n[c]; v->times = h->s_t[c]; v->events = h->s_ev[c]; v->x = h->s_x[c]; v->u = h->s_u[c]; v->uff = h->s_uff[c]; v->K = h->s_K[c];
  v->max_nodes = h->NS; v->nx = h->nx; v->nu = h->nu; v->batch = h->B; v->slab = h->slab[c]; v->slab_bytes = h->slab_doubles * sizeof(double);
  return BMPC_OK;
}
int bmpc_get_device_view_inflight(bmpc_handle* h, bmpc_device_view* v) {
  if (!h || !v) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  try_publish(h);
  if (!h->have_solution && !h->pending) return BMPC_ERR_INVALID;
  const int c = h->pending ? 1 - h->cur : h->cur;
  v->n_nodes = h->s_n[c]; v->times = h->s_t[c]; v->events = h->s_ev[c]; v->x = h->s_x[c]; v->u = h->s_u[c]; v->uff = h->s_uff[c]; v->K = h->s_K[c];
  v->max_nodes = h->NS; v->nx = h->nx; v->nu = h->nu; v->batch = h->B; v->slab = h->slab[c]; v->slab_bytes = h->slab_doubles * sizeof(double);
  return BMPC_OK;
}
int bmpc_get_performance(bmpc_handle* h, double* perf) {
  API_BEGIN if (!h || !perf) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  ReadGuard g(h);
  std::lock_guard<std::mutex> io(h->eval_mtx);
  CK(cudaMemcpyAsync(perf, h->s_perf[g.c], sizeof(double) * h->B * 8, cudaMemcpyDeviceToHost, h->io_stream)); CK(cudaStreamSynchronize(h->io_stream));
  return BMPC_OK; API_END(h)
}
int bmpc_get_status(bmpc_handle* h, int* status) {
  API_BEGIN if (!h || !status) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  ReadGuard g(h);
  std::lock_guard<std::mutex> io(h->eval_mtx);
  CK(cudaMemcpyAsync(status, h->s_status[g.c], sizeof(int) * h->B, cudaMemcpyDeviceToHost, h->io_stream)); CK(cudaStreamSynchronize(h->io_stream));
  return BMPC_OK; API_END(h)
}
int bmpc_evaluate_policy(bmpc_handle* h, const double* t, const double* x, double* x_opt, double* u_opt, int* mode) {
  API_BEGIN if (!h || !t || !x) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  ReadGuard g(h);
  std::lock_guard<std::mutex> io(h->eval_mtx);   // the query / result scratch buffers are shared
  const int c = g.c, B = h->B; cudaStream_t st = h->io_stream;
  CK(cudaMemcpyAsync(h->d_eval_t, t, sizeof(double) * B, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(h->d_eval_x, x, sizeof(double) * B * h->nx, cudaMemcpyHostToDevice, st));
  if (h->nj == 10) k_evaluate_policy<10><<<B, 32, 0, st>>>(B, h->NS, h->ME, h->s_n[c], h->s_t[c], h->s_x[c], h->s_uff[c], h->s_K[c], h->s_nev[c], h->s_evt[c], h->s_evm[c], h->d_eval_t, h->d_eval_x, h->d_eval_xo, h->d_eval_uo, h->d_eval_m);
  else k_evaluate_policy<12><<<B, 32, 0, st>>>(B, h->NS, h->ME, h->s_n[c], h->s_t[c], h->s_x[c], h->s_uff[c], h->s_K[c], h->s_nev[c], h->s_evt[c], h->s_evm[c], h->d_eval_t, h->d_eval_x, h->d_eval_xo, h->d_eval_uo, h->d_eval_m);
  if (x_opt) CK(cudaMemcpyAsync(x_opt, h->d_eval_xo, sizeof(double) * B * h->nx, cudaMemcpyDeviceToHost, st));
  if (u_opt) CK(cudaMemcpyAsync(u_opt, h->d_eval_uo, sizeof(double) * B * h->nu, cudaMemcpyDeviceToHost, st));
  if (mode) CK(cudaMemcpyAsync(mode, h->d_eval_m, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st)); CK(cudaGetLastError());
  return BMPC_OK; API_END(h)
}

// Device-resident drivers for closed-loop batches (observations and targets never leave HBM).  They run on the compute stream, i.e. after
// the tick in flight, and use the policy that tick produces.
int bmpc_shift_observations(bmpc_handle* h, double dt) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  if (!h->have_solution && !h->pending) throw std::invalid_argument("[bmpc] no solution yet");
  upload_inputs(h);
  const int c = h->pending ? 1 - h->cur : h->cur, B = h->B;   // stream order: the pending tick has written buffer 1 - cur by the time this kernel runs
  if (h->nj == 10) k_shift_observations<10><<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->NS, dt, h->s_n[c], h->s_t[c], h->s_x[c], h->d_t0, h->d_x0);
  else k_shift_observations<12><<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->NS, dt, h->s_n[c], h->s_t[c], h->s_x[c], h->d_t0, h->d_x0);
  h->have_obs = false; CK(cudaGetLastError());
  return BMPC_OK; API_END(h)
}
// MRT_BASE::rolloutPolicy [UPSTREAM] (BipedalController.cpp:322, task.info:159-167), batched and device resident: see k_rollout
int bmpc_rollout_observations(bmpc_handle* h, double time_step, int substeps) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  if (!(time_step > 0.0) || substeps < 1) throw std::invalid_argument("[bmpc] rollout needs time_step > 0 and substeps >= 1");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  if (!h->have_solution && !h->pending) throw std::invalid_argument("[bmpc] no solution yet");
  ensure_model_image(h);
  upload_inputs(h);
  const int c = h->pending ? 1 - h->cur : h->cur, B = h->B;
  if (h->nj == 10) k_rollout<10><<<(B + 3) / 4, 128, 0, h->stream>>>(B, h->NS, h->ME, h->s_n[c], h->s_t[c], h->s_uff[c], h->s_K[c], h->s_nev[c], h->s_evt[c], h->d_t0, h->d_x0, time_step, substeps, h->rollout_abs, h->rollout_rel, h->rollout_dt, h->s_status[c]);
  else k_rollout<12><<<(B + 3) / 4, 128, 0, h->stream>>>(B, h->NS, h->ME, h->s_n[c], h->s_t[c], h->s_uff[c], h->s_K[c], h->s_nev[c], h->s_evt[c], h->d_t0, h->d_x0, time_step, substeps, h->rollout_abs, h->rollout_rel, h->rollout_dt, h->s_status[c]);
  h->have_obs = false; CK(cudaGetLastError());
  return BMPC_OK; API_END(h)
}
int bmpc_set_rollout_settings(bmpc_handle* h, double abs_tol, double rel_tol, double initial_time_step) {
  if (!h || !(abs_tol > 0.0) || !(rel_tol >= 0.0) || !(initial_time_step > 0.0)) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  h->rollout_abs = abs_tol; h->rollout_rel = rel_tol; h->rollout_dt = initial_time_step;
  return BMPC_OK;
}
int bmpc_set_targets_from_cmd_vel_device(bmpc_handle* h, const double* cmd_dev, double time_to_target) {
  API_BEGIN if (!h || !cmd_dev) return BMPC_ERR_INVALID;
  if (h->TP < 2) throw std::length_error("[bmpc] max_target_points < 2");
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  upload_inputs(h);
  const int B = h->B;
  if (h->nj == 10) k_cmd_vel_targets<10><<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->TP, h->d_t0, h->d_x0, cmd_dev, time_to_target, h->model.com_height, h->d_default_joints, h->d_tgt_t, h->d_tgt_x);
  else k_cmd_vel_targets<12><<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->TP, h->d_t0, h->d_x0, cmd_dev, time_to_target, h->model.com_height, h->d_default_joints, h->d_tgt_t, h->d_tgt_x);
  h->npts = 2; h->tgt_dirty = false; h->have_tgt = true; CK(cudaGetLastError());
  return BMPC_OK; API_END(h)
}
int bmpc_get_observations(bmpc_handle* h, double* t, double* x) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  upload_inputs(h);
  if (t) CK(cudaMemcpyAsync(t, h->d_t0, sizeof(double) * h->B, cudaMemcpyDeviceToHost, h->stream));
  if (x) CK(cudaMemcpyAsync(x, h->d_x0, sizeof(double) * h->B * h->nx, cudaMemcpyDeviceToHost, h->stream));
  lk.unlock();
  CK(cudaStreamSynchronize(h->stream));
  return BMPC_OK; API_END(h)
}

// ------------------------------------------------------------------------------------------------ multi-GPU policy exchange
int bmpc_exchange_create_id(bmpc_exchange_id* id) {
  try {
    if (!id) throw std::invalid_argument("[bmpc] null argument");
    NcclApi& N = nccl_api(); N.load();
    static_assert(sizeof(bmpc_exchange_id) == sizeof(NcclApi::UniqueId), "id size");
    N.check(N.GetUniqueId(reinterpret_cast<NcclApi::UniqueId*>(id)), "ncclGetUniqueId");
    return BMPC_OK;
  } catch (const std::exception& e) { return fail(nullptr, BMPC_ERR_INVALID, e.what()); }
}
int bmpc_exchange_init(bmpc_handle* h, int rank, int nranks, const bmpc_exchange_id* id, int max_ctas, int use_copy_engines) {
  API_BEGIN if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("[bmpc] bad exchange arguments");
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);
  exchange_release(h);
  NcclApi& N = nccl_api(); N.load();
  auto& ex = h->ex;
  ex.rank = rank; ex.nranks = nranks; ex.max_ctas = max_ctas;
  NcclApi::UniqueId uid; std::memcpy(&uid, id, sizeof(uid));
  const bool can_config = N.CommInitRankConfig && N.version >= 22800;
  // mode 1: copy-engine collectives (NCCL >= 2.28: CTA policy "zero" + symmetric windows): the all-gather then uses no SM at all;
  // mode 2: symmetric windows with NCCL's SM kernels (fewer CTAs reach the same bandwidth than with unregistered buffers)
  const bool windows = use_copy_engines != 0 && can_config && N.MemAlloc && N.CommWindowRegister;
  ex.copy_engines = windows && use_copy_engines == 1;
  ex.symmetric = windows;
  if (can_config) {
    NcclApi::ConfigV22800 cfg = N.default_config();
    if (max_ctas > 0) { cfg.maxCTAs = max_ctas; cfg.minCTAs = 1; }
    if (ex.copy_engines) cfg.CTAPolicy = NcclApi::kCtaPolicyZero;
    N.check(N.CommInitRankConfig(&ex.comm, nranks, uid, rank, &cfg), "ncclCommInitRankConfig");
  } else N.check(N.CommInitRank(&ex.comm, nranks, uid, rank), "ncclCommInitRank");
  CK(cudaStreamCreateWithFlags(&ex.stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&ex.ready, cudaEventDisableTiming));
  for (auto& e : ex.done) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const size_t bytes = h->slab_doubles * sizeof(double);
  ex.nccl_mem = windows;
  for (int i = 0; i < 2; ++i) {
    if (ex.nccl_mem) N.check(N.MemAlloc(&ex.recv[i], bytes * nranks), "ncclMemAlloc"); else CK(cudaMalloc(&ex.recv[i], bytes * nranks));
  }
  if (windows) {
    // the policy slabs themselves move into NCCL-allocated, window-registered memory: the all-gather reads them in place (no staging copy)
    for (int i = 0; i < 2; ++i) {
      void* p = nullptr;
      N.check(N.MemAlloc(&p, bytes), "ncclMemAlloc");
      CK(cudaMemcpy(p, h->slab[i], bytes, cudaMemcpyDeviceToDevice));
      auto& dv = h->pool.dev;
      dv.erase(std::remove(dv.begin(), dv.end(), (void*)h->slab[i]), dv.end());
      CK(cudaFree(h->slab[i]));
      set_slab_pointers(h, i, static_cast<double*>(p));
      ex.send_slab[i] = p;
      N.check(N.CommWindowRegister(ex.comm, p, bytes, &ex.win_send[i], NcclApi::kWinCollSymmetric), "ncclCommWindowRegister");
      N.check(N.CommWindowRegister(ex.comm, ex.recv[i], bytes * nranks, &ex.win_recv[i], NcclApi::kWinCollSymmetric), "ncclCommWindowRegister");
    }
  }
  return BMPC_OK; API_END(h)
}
// all-gather of the newest policy slab (the one the tick in flight writes), ordered after that tick, on the exchange stream
int bmpc_exchange_start(bmpc_handle* h) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::lock_guard<std::mutex> lk(h->mtx);
  auto& ex = h->ex;
  if (!ex.comm) throw std::invalid_argument("[bmpc] bmpc_exchange_init has not been called");
  if (!h->have_solution && !h->pending) throw std::invalid_argument("[bmpc] no solution yet");
  NcclApi& N = nccl_api();
  const int c = h->pending ? 1 - h->cur : h->cur;
  CK(cudaEventRecord(ex.ready, h->stream));
  CK(cudaStreamWaitEvent(ex.stream, ex.ready, 0));
  N.check(N.AllGather(h->slab[c], ex.recv[c], h->slab_doubles, NcclApi::kFloat64, ex.comm, ex.stream), "ncclAllGather");
  CK(cudaEventRecord(ex.done[c], ex.stream));
  ex.started[c] = true; ex.last = c; ++ex.count;
  return BMPC_OK; API_END(h)
}
int bmpc_exchange_wait(bmpc_handle* h) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  if (h->ex.stream) CK(cudaStreamSynchronize(h->ex.stream));
  return BMPC_OK; API_END(h)
}
// device pointer of the gathered slabs of the last started exchange: nranks consecutive slabs of slab_bytes (rank r at offset r * slab_bytes);
// valid after bmpc_exchange_wait (or for work ordered after it) until the exchange after the next one starts
int bmpc_exchange_view(bmpc_handle* h, const void** gathered, unsigned long long* slab_bytes, int* nranks) {
  if (!h || !h->ex.comm || h->ex.last < 0) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  if (gathered) *gathered = h->ex.recv[h->ex.last];
  if (slab_bytes) *slab_bytes = h->slab_doubles * sizeof(double);
  if (nranks) *nranks = h->ex.nranks;
  return h->ex.copy_engines ? 1 : (h->ex.symmetric ? 2 : 0);
}
int bmpc_exchange_destroy(bmpc_handle* h) {
  API_BEGIN if (!h) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);
  exchange_release(h);
  return BMPC_OK; API_END(h)
}

int bmpc_get_launch_count(const bmpc_handle* h) { return h ? h->launches : 0; }
int bmpc_debug_set_option(bmpc_handle* h, const char* name, int value) {
  if (!h || !name) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  if (std::string(name) == "projection_mode") {
    if (value < 0 || value > 1) return BMPC_ERR_INVALID;
    if (value == 0 && h->model.dev.gain != 0.0) return fail(h, BMPC_ERR_INVALID, "[bmpc] the Moore-Penrose projection option does not support positionErrorGain != 0");
    h->projection_mode = value; return BMPC_OK;
  }
  return BMPC_ERR_INVALID;
}
int bmpc_enable_phase_timing(bmpc_handle* h, int enable) { if (!h) return BMPC_ERR_INVALID; std::lock_guard<std::mutex> lk(h->mtx); h->timing = enable != 0; return BMPC_OK; }
int bmpc_get_phase_times(bmpc_handle* h, float* ms) {
  if (!h || !ms) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  for (int i = 0; i < 8; ++i) ms[i] = h->phase_ms[i];
  ms[8] = (float)h->max_trials;
  return BMPC_OK;
}
int bmpc_get_tick_stats(bmpc_handle* h, int* total_trials, int* max_trials, int* failed_instances, int* status_or) {
  if (!h) return BMPC_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mtx);
  try_publish(h);
  if (total_trials) *total_trials = h->linesearch_trials; if (max_trials) *max_trials = h->max_trials;
  if (failed_instances) *failed_instances = h->failed_instances; if (status_or) *status_or = h->status_or;
  return BMPC_OK;
}
void* bmpc_get_stream(bmpc_handle* h) { return h ? (void*)h->stream : nullptr; }

int bmpc_debug_record_sizes(const bmpc_handle* h, int* lq_rec, int* proj_rec, int* ric_rec) {
  if (!h) return BMPC_ERR_INVALID;
  if (lq_rec) *lq_rec = (int)h->rec; if (proj_rec) *proj_rec = (int)h->prec; if (ric_rec) *ric_rec = (int)h->krec;
  return BMPC_OK;
}
int bmpc_debug_copy(bmpc_handle* h, const char* name, int instance, double* dst, int capacity) {
  API_BEGIN if (!h || !name || !dst || instance < 0 || instance >= h->B) return BMPC_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  std::unique_lock<std::mutex> lk(h->mtx);
  wait_and_publish(h, lk);
  CK(cudaStreamSynchronize(h->stream));
  const std::string n(name); const size_t NS = h->NS, b = instance;
  const double* src = nullptr; size_t cnt = 0;
  const int c = h->cur;
  if (n == "lq_record") { src = h->d_lq + b * NS * h->rec; cnt = NS * h->rec; }
  else if (n == "proj_record") { src = h->d_proj + b * NS * h->prec; cnt = NS * h->prec; }
  else if (n == "stage_record") { src = h->d_stage + b * NS * h->srec; cnt = NS * h->srec; }
  else if (n == "riccati_record") { src = h->d_ric + b * NS * h->krec; cnt = NS * h->krec; }
  else if (n == "dx") { src = h->d_dx + b * NS * h->nx; cnt = NS * h->nx; }
  else if (n == "du") { src = h->d_du + b * NS * h->nu; cnt = NS * h->nu; }
  else if (n == "xref") { src = h->d_xref + b * NS * h->nx; cnt = NS * h->nx; }
  else if (n == "zref") { src = h->d_zref + b * NS * 4; cnt = NS * 4; }
  else if (n == "st_t") { src = h->d_st_t + b * NS; cnt = NS; }
  else if (n == "st_dt") { src = h->d_st_dt + b * NS; cnt = NS; }
  else if (n == "x") { src = h->s_x[c] + b * NS * h->nx; cnt = NS * h->nx; }
  else if (n == "u") { src = h->s_u[c] + b * NS * h->nu; cnt = NS * h->nu; }
  else throw std::invalid_argument("[bmpc] unknown debug buffer " + n);
  if ((size_t)capacity < cnt) throw std::length_error("[bmpc] debug buffer capacity too small");
  CK(cudaMemcpy(dst, src, sizeof(double) * cnt, cudaMemcpyDeviceToHost));
  return (int)cnt; API_END(h)
}

}  // extern "C"

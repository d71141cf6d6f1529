/* bmpc.h - C ABI of the B200-native batched bipedal MPC solver (libbmpc.so).
 *
 * Drop-in boundary for the hot path of zitongbai/bipedal_control: one multiple-shooting SQP iteration of the
 * centroidal switched-system OCP that ocs2_bipedal_robot defines (BipedalRobotInterface), for a batch of B
 * independent instances on one GPU.  Each entry point cites the reference interface it replaces
 * (paths relative to the reference repository root; [UPSTREAM] = OCS2 class that the reference calls but does
 * not vendor).  No exceptions cross this ABI: every call returns 0 on success or a negative bmpc_status, and
 * bmpc_last_error() returns the message.
 *
 * Layout conventions: all matrices row-major, double precision.
 *   state  x[nx]  = [h_lin/m (3), h_ang/m (3), base xyz (3), base ZYX Euler (3), leg joints (nj)]
 *   input  u[nu]  = [contact forces 4x3 (world), leg joint velocities (nj)]      (task.info:181-277)
 */
#ifndef BMPC_H
#define BMPC_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bmpc_handle bmpc_handle;

enum bmpc_status {
  BMPC_OK = 0,
  BMPC_ERR_INVALID = -1,   /* bad argument (reference: std::invalid_argument in BipedalRobotInterface.cpp:71-90) */
  BMPC_ERR_CUDA = -2,      /* CUDA runtime failure / no device: the library never falls back to the CPU */
  BMPC_ERR_NUMERIC = -3,   /* at least one instance reported a numerical failure (see bmpc_get_status) */
  BMPC_ERR_CAPACITY = -4   /* horizon / events exceed the capacities given at creation */
};

/* per-instance status bits returned by bmpc_get_status */
enum bmpc_instance_status {
  BMPC_INST_OK = 0,
  BMPC_INST_RICCATI_NOT_PD = 1,   /* reduced Hessian lost positive definiteness (HPIPM would fail the QP) */
  BMPC_INST_RANK_ANOMALY = 2,     /* constraint Jacobian lost rank beyond the structural stance-foot deficiency */
  BMPC_INST_SWING_UNDEFINED = 4,  /* lift-off / touch-down time undefined (SwingTrajectoryPlanner.cpp:191-212 throws) */
  BMPC_INST_NAN = 8,
  BMPC_INST_STEP_REJECTED = 16,   /* filter line search reached alpha_min: no step taken (informational) */
  BMPC_INST_GRID_OVERFLOW = 32    /* more event nodes inside the horizon than max_event_nodes: the time grid was truncated */
};
/* Failure semantics.  Upstream throws (HPIPM failure, SwingTrajectoryPlanner.cpp:191-212, ...) and BipedalController stops the controller
 * (BipedalController.cpp:344-348).  Here the tick always completes for the whole batch; bmpc_advance / bmpc_synchronize then return
 *   BMPC_ERR_NUMERIC  if any instance has bit 1, 2 or 8: that instance took no step and its policy is open loop (K = 0, uff = warm-start input),
 *                     so nothing non-finite is stored and the next tick warm-starts from valid data;
 *   BMPC_ERR_INVALID  if any instance has bit 4;   BMPC_ERR_CAPACITY if any instance has bit 32.
 * The other instances are unaffected and every getter keeps working after such a return. */

/* Construction: replaces `BipedalRobotInterface(taskFile, urdfFile, referenceFile)` +
 * `std::make_shared<SqpMpc>(mpcSettings, sqpSettings, ocp, initializer)`
 * (bipedal_controllers/src/BipedalController.cpp:282-306).  Either `model_file` (compact derived numbers written by
 * tools/ingest.py or bmpc_export_model) or the reference's own task/reference/gait .info + URDF files are given. */
typedef struct bmpc_config {
  const char* model_file;
  const char* task_file;
  const char* reference_file;
  const char* gait_file;
  const char* urdf_file;
  int batch;               /* number of independent OCP instances B */
  int device;              /* CUDA device ordinal */
  double dt;               /* <= 0: sqp.dt of the task file (task.info:69) */
  double time_horizon;     /* <= 0: mpc.timeHorizon (task.info:171) */
  int max_events;          /* capacity of each instance's mode schedule (0: default 40) */
  int max_target_points;   /* capacity of each instance's TargetTrajectories (0: default 4) */
  int sqp_iterations;      /* <= 0: sqp.sqpIteration (task.info:70) */
  int max_event_nodes;     /* extra grid nodes for mode switches inside the horizon: 2 per switch off the nominal grid, 1 per switch on it (0: default 16) */
} bmpc_config;

int bmpc_create(const bmpc_config* cfg, bmpc_handle** out);
void bmpc_destroy(bmpc_handle* h);
const char* bmpc_last_error(const bmpc_handle* h); /* h may be NULL: error of the last failed bmpc_create */

/* sizes: nx, nu, batch, maximum number of nodes per instance (N+1 including event nodes) */
int bmpc_get_dims(const bmpc_handle* h, int* nx, int* nu, int* batch, int* max_nodes);
int bmpc_get_initial_state(const bmpc_handle* h, double* x /* nx */);      /* task.info `initialState` */
int bmpc_export_model(const bmpc_handle* h, const char* path);             /* writes the compact model file */
/* host-only tool: reads task.info / reference.info / gait.info (may be NULL) / URDF exactly as BipedalRobotInterface does
 * (BipedalRobotInterface.cpp:67-204) and writes the compact model file; needs no GPU */
int bmpc_convert_model(const char* task_file, const char* reference_file, const char* gait_file, const char* urdf_file, const char* out_path);

/* MPC_MRT_Interface::reset / resetMpcNode (BipedalController.cpp:147-148): drops the warm start, the policy and the gait bookkeeping of one
 * instance, or of all instances if instance < 0.  Waits for a tick in flight. */
int bmpc_reset(bmpc_handle* h, int instance);

/* MPC_MRT_Interface::setCurrentObservation (BipedalController.cpp:191): SystemObservation{time, state} per instance.
 * t[B], x[B*nx], host memory (copied).  The *_device variants take device pointers (inputs already in HBM). */
int bmpc_set_observations(bmpc_handle* h, const double* t, const double* x);
int bmpc_set_observations_device(bmpc_handle* h, const double* t_dev, const double* x_dev);

/* ReferenceManager::setTargetTrajectories [UPSTREAM] (BipedalController.cpp:145,153): npts knots per instance,
 * times[B*npts], states[B*npts*nx]; the cost interpolates linearly (BipedalRobotQuadraticTrackingCost.h:60). */
int bmpc_set_target_trajectories(bmpc_handle* h, int npts, const double* times, const double* states);
int bmpc_set_target_trajectories_device(bmpc_handle* h, int npts, const double* times_dev, const double* states_dev);

/* Batched TargetTrajectoriesPublisher::cmdVelToTargetTrajectories (bipedal_controllers/src/TargetTrajectoriesPublisher.cpp:76-99):
 * builds the 2-knot targets from the observations set before the call and cmd[B*4] = (vx, vy, vz, yaw rate) in host memory.  Only the commands cross
 * PCIe (the targets are built by the device kernel of the _device variant) unless a tick is in flight, in which case they are built on the host and
 * travel with the next tick, so that the call never queues behind the solve. */
int bmpc_set_targets_from_cmd_vel(bmpc_handle* h, const double* cmd, double time_to_target);

int bmpc_set_targets_from_cmd_vel_device(bmpc_handle* h, const double* cmd_dev, double time_to_target); /* same, observations and cmd in HBM */

/* Closed-loop driver for device-resident batches: t0 += dt and x0 = optimized state trajectory of the newest policy (the one of the tick in
 * flight, in stream order) interpolated at the new time (perfect-model stand-in for the MRT_ROS_Dummy_Loop rollout [UPSTREAM] that
 * ocs2_bipedal_robot_ros/src/BipedalRobotDummyNode.cpp:72-86 runs between MPC ticks).  bmpc_get_observations reads them back. */
int bmpc_shift_observations(bmpc_handle* h, double dt);
/* MRT_BASE::rolloutPolicy [UPSTREAM] for the whole batch, device resident (mpcMrtInterface_->initRollout(&interface.getRollout()),
 * BipedalController.cpp:322; the dummy loop of ocs2_bipedal_robot_ros/src/BipedalRobotDummyNode.cpp:72-86 advances the observation with it):
 * integrates the closed loop xdot = f(x, uff(t) + K(t) x) of the newest policy from each instance's observation (t0, x0) over `substeps`
 * consecutive periods of time_step / substeps and stores the result as the new observation.  Integrator = upstream's rollout settings
 * (task.info:159-167: ODE45 = Dormand-Prince 5(4) with odeint's step-size control, AbsTolODE 1e-5, RelTolODE 1e-3, initial step `timeStep` 0.015;
 * sub-intervals split at the policy's mode switches); bmpc_set_rollout_settings overrides the three numbers.  Instances without a policy only
 * advance in time; an instance whose integration fails keeps its observation and gets status bit 8. */
int bmpc_rollout_observations(bmpc_handle* h, double time_step, int substeps);
int bmpc_set_rollout_settings(bmpc_handle* h, double abs_tol, double rel_tol, double initial_time_step);
int bmpc_get_observations(bmpc_handle* h, double* t, double* x);

/* ReferenceManager::setModeSchedule [UPSTREAM]: explicit ModeSchedule per instance; n_events[B],
 * event_times[B*stride], mode_sequence[B*(stride+1)] (mode ids: 0 FLY, 1 LF, 2 RF, 3 STANCE;
 * gait/MotionPhaseDefinition.h:47-52).  Disables the internal GaitSchedule for the next solves. */
int bmpc_set_mode_schedules(bmpc_handle* h, int stride, const int* n_events, const double* event_times, const int* mode_sequence);
int bmpc_set_mode_schedules_device(bmpc_handle* h, int stride, const int* n_events_dev, const double* event_times_dev, const int* mode_sequence_dev);

/* GaitSchedule (ocs2_bipedal_robot/src/gait/GaitSchedule.cpp): per-instance gait bookkeeping, device resident (one thread per instance; a tick
 * has no per-instance host work, and the observations may be host or device resident).
 * bmpc_gait_insert == GaitSchedule::insertModeSequenceTemplate (GaitSchedule.cpp:46-73; called from
 * GaitReceiver::preSolverRun, ocs2_bipedal_robot_ros/src/gait/GaitReceiver.cpp:49-59); instance < 0 applies to all.
 * bmpc_use_gait_schedule(1) makes bmpc_advance derive each instance's ModeSchedule with
 * GaitSchedule::getModeSchedule(t0 - T, tf + T) as SwitchedModelReferenceManager::modifyReferences does
 * (SwitchedModelReferenceManager.cpp:62-69). */
int bmpc_gait_insert(bmpc_handle* h, int instance, int n_modes, const int* modes, const double* switching_times, double start_time, double final_time);
int bmpc_gait_insert_named(bmpc_handle* h, int instance, const char* gait_name, double start_time, double final_time);
int bmpc_use_gait_schedule(bmpc_handle* h, int enable);
int bmpc_gait_peek(const bmpc_handle* h, int instance, int cap, double* event_times, int* mode_sequence); /* returns n_events */

/* MPC_MRT_Interface::advanceMpc -> MPC_BASE::run(t, x) -> SqpSolver::runImpl [UPSTREAM] (BipedalController.cpp:339):
 * one MPC tick for all B instances: reference update, LQ approximation, projected Riccati QP, filter line search,
 * feedback policy.  Synchronous: the new policy is current when it returns. */
int bmpc_advance(bmpc_handle* h);
/* The same tick, only ENQUEUED on the library's stream: the call returns as soon as the kernels are queued (no host synchronisation inside the
 * tick -- the line-search loop runs on the device).  The policy the getters serve stays the previous one until the tick is published:
 * by bmpc_synchronize (waits), or by the first getter / bmpc_poll that finds the tick finished (MRT_BASE::updatePolicy semantics).
 * Threading (MPC_MRT_Interface [UPSTREAM], BipedalController.cpp:191-200 vs :332-351): one thread at a time calls bmpc_advance* on a handle; any
 * other thread may call the set_* functions and the getters (bmpc_get_policy, bmpc_evaluate_policy, bmpc_get_performance, bmpc_get_status,
 * bmpc_get_device_view) concurrently: they never wait for a tick in flight.  A second bmpc_advance* first waits for the previous tick.
 * Handles of different robots may be used concurrently; their ticks are serialised on the device's shared model image, handles of the same
 * robot overlap freely. */
int bmpc_advance_async(bmpc_handle* h);
int bmpc_synchronize(bmpc_handle* h);   /* waits for the tick in flight, publishes it, returns its status (see "Failure semantics") */
int bmpc_poll(bmpc_handle* h);          /* publishes the tick in flight if it has finished; returns 1 while a tick is still running, else 0 */

/* PrimalSolution [UPSTREAM] of instances [first, first+count): any output pointer may be NULL.
 * n_nodes[count]; times[count*max_nodes]; events[count*max_nodes] (0 none, 1 pre-event, 2 post-event);
 * x[count*max_nodes*nx]; u[count*max_nodes*nu]; uff[count*max_nodes*nu]; K[count*max_nodes*nu*nx]
 * (LinearController: u = uff + K x).  Host memory. */
int bmpc_get_policy(bmpc_handle* h, int first, int count, int* n_nodes, double* times, int* events, double* x, double* u, double* uff, double* K);

/* device pointers of the current (published) policy buffers; they stay untouched until the tick AFTER the next one starts (double buffering).
 * bmpc_get_device_view_inflight: the buffers the tick in flight is writing (or the current ones if none is in flight) -- valid for work that
 * is ordered after the tick on bmpc_get_stream(), e.g. the multi-GPU policy exchange. */
typedef struct bmpc_device_view {
  const int* n_nodes; const double* times; const int* events;
  const double* x; const double* u; const double* uff; const double* K;
  int max_nodes, nx, nu, batch;
  /* the arrays above live in ONE contiguous allocation [K | uff | x | u | times | events | n_nodes] of slab_bytes bytes starting at slab (= K):
   * the multi-GPU exchange of a whole shard's policies is a single all-gather of this range */
  const void* slab; unsigned long long slab_bytes;
} bmpc_device_view;
int bmpc_get_device_view(bmpc_handle* h, bmpc_device_view* v);
int bmpc_get_device_view_inflight(bmpc_handle* h, bmpc_device_view* v);

/* Multi-GPU: one process per GPU, each with its own handle over a contiguous shard of the batch (instances are independent: the data path needs
 * no collective).  The only exchange is the all-gather of the solved policies once per tick (every rank then holds every robot's policy):
 *   rank 0: bmpc_exchange_create_id(&id); broadcast the 128 bytes to the other ranks (MPI, torch.distributed, a file ...);
 *   every rank: bmpc_exchange_init(h, rank, nranks, &id, max_ctas, use_copy_engines);
 *   every tick: bmpc_advance_async(h); bmpc_exchange_start(h);   -> ONE ncclAllGather of the policy slab on its own stream, overlapped with the
 *               next tick (the slab is double buffered; the tick that overwrites it waits for the gather by itself);
 *   consumers: bmpc_exchange_wait(h); bmpc_exchange_view(h, &ptr, &slab_bytes, &nranks)  (rank r's slab at ptr + r * slab_bytes, laid out as
 *              bmpc_device_view describes).
 * max_ctas > 0 caps the SMs NCCL may use (the solver keeps the rest); use_copy_engines = 1 asks for NCCL's copy-engine all-gather (NCCL >= 2.28,
 * symmetric windows, CTA policy "zero": no SM at all), 2 for symmetric windows with NCCL's SM kernels; in both the policy slabs move into
 * NCCL-registered memory (device views obtained earlier are stale) and are gathered in place.  bmpc_exchange_view returns the active mode
 * (0 plain buffers, 1, 2).  NCCL is loaded at run time (libnccl.so.2) only when these functions are used.
 * Tear-down is collective (as ncclCommDestroy is): bmpc_exchange_destroy, or bmpc_destroy of a handle with an exchange, must be called by every rank;
 * with symmetric windows the ranks are fenced inside the call so that no rank frees memory a slower peer still has mapped. */
typedef struct bmpc_exchange_id { char bytes[128]; } bmpc_exchange_id;
int bmpc_exchange_create_id(bmpc_exchange_id* id);
int bmpc_exchange_init(bmpc_handle* h, int rank, int nranks, const bmpc_exchange_id* id, int max_ctas, int use_copy_engines);
int bmpc_exchange_start(bmpc_handle* h);
int bmpc_exchange_wait(bmpc_handle* h);
int bmpc_exchange_view(bmpc_handle* h, const void** gathered, unsigned long long* slab_bytes, int* nranks);
int bmpc_exchange_destroy(bmpc_handle* h);

/* PerformanceIndex [UPSTREAM] per instance: perf[B*8] = {cost, dynamicsViolationSSE, equalityConstraintsSSE} before the
 * step, the same three after the accepted step, step size alpha, armijo descent metric. */
int bmpc_get_performance(bmpc_handle* h, double* perf);
int bmpc_get_status(bmpc_handle* h, int* status /* B */);

/* MPC_MRT_Interface::evaluatePolicy(t, x, &xOpt, &uOpt, &mode) (BipedalController.cpp:200), batched:
 * t[B], x[B*nx] -> x_opt[B*nx], u_opt[B*nu], mode[B].  Host memory. */
int bmpc_evaluate_policy(bmpc_handle* h, const double* t, const double* x, double* x_opt, double* u_opt, int* mode);

/* number of kernels launched by the last bmpc_advance (for bench.py's gpu_launches) and per-phase device times in
 * milliseconds of the last tick: ms[0..7] = {setup, lq, projection, riccati, policy_expand, forward, linesearch + step, policy completion}, ms[8] = largest number of
 * line-search trials any instance needed */
int bmpc_get_launch_count(const bmpc_handle* h);
int bmpc_get_phase_times(bmpc_handle* h, float* ms /* 9 */);
int bmpc_enable_phase_timing(bmpc_handle* h, int enable);
/* statistics of the last published tick: line-search trials summed over the batch, their maximum, number of failed instances, OR of all status bits */
int bmpc_get_tick_stats(bmpc_handle* h, int* total_trials, int* max_trials, int* failed_instances, int* status_or);
void* bmpc_get_stream(bmpc_handle* h); /* cudaStream_t the library launches on */

/* Test hooks (used by tests/ to compare intermediate device data with the oracle): copies the named device buffer of one
 * instance to host.  names: "lq_record", "proj_record", "stage_record", "riccati_record", "dx", "du", "xref", "zref", "st_t", "st_dt", "x", "u" */
int bmpc_debug_copy(bmpc_handle* h, const char* name, int instance, double* dst, int capacity_doubles);
int bmpc_debug_record_sizes(const bmpc_handle* h, int* lq_rec, int* proj_rec, int* ric_rec);
/* options:
 *   "projection_mode"  1 (default): upstream's constraint projection, Eigen::FullPivLU kernel() / solve() (LinearAlgebra::luConstraintProjection
 *                      [UPSTREAM], task.info:76);  0: Moore-Penrose projection (Householder QR, minimum-norm particular solution, orthonormal
 *                      null-space basis).  The two give the same QP whenever the stance feet are at rest at the linearisation point. */
int bmpc_debug_set_option(bmpc_handle* h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* BMPC_H */

// CUDA kernels of one SQP tick (sm_100a, FP64), split by phase.  Kernel <-> reference mapping:
//   bmpc_kernels_setup.cuh      k_gait_schedule / k_gait_insert : GaitSchedule::getModeSchedule / insertModeSequenceTemplate (gait/GaitSchedule.cpp:46-137) per instance;
//                               k_time_grid / k_node_setup : timeDiscretizationWithEvents, ModeSchedule::modeAtTime, TargetTrajectories::getDesiredState,
//                               SwingTrajectoryPlanner::getZvelocityConstraint (foot_planner/SwingTrajectoryPlanner.cpp:50-118),
//                               multiple_shooting::initializeStateInputTrajectories + BipedalRobotInitializer::compute [UPSTREAM / initializer]
//   bmpc_kernels_lq.cuh         k_lq_pack :
//                               multiple_shooting::setupIntermediateNode / setupEventNode (dynamics RK2 sensitivity, cost, soft friction cone,
//                               zero-force / zero-velocity / normal-velocity constraints)  -> compact LQ record
//   bmpc_kernels_project.cuh    k_project : LinearAlgebra::luConstraintProjection replacement (Householder QR, min-norm particular solution)
//                               + changeOfInputVariables on the FP64 tensor cores -> projected stage record (SDims)
//   bmpc_kernels_riccati.cuh    k_riccati_warp : HPIPM backward Riccati recursion (DMMA m8n8k4, TMA-staged records)
//   bmpc_kernels_policy.cuh     k_policy_expand : Riccati feedback -> K, uff, closed-loop stage maps ; k_forward : HPIPM forward substitution,
//                               armijoDescentMetric, PerformanceIndex reduction
//   bmpc_kernels_linesearch.cuh k_linesearch : SqpSolver::computePerformance + FilterLinesearch::acceptStep (device-side backtracking loop) ;
//                               k_update / k_policy_fill : incrementTrajectory + multiple_shooting::toPrimalSolution (LinearController uff, K)
//   bmpc_kernels_io.cuh         k_evaluate_policy (MPC_MRT_Interface::evaluatePolicy), k_rollout (MRT_BASE::rolloutPolicy), k_shift_observations, k_cmd_vel_targets (TargetTrajectoriesPublisher.cpp:76-99)
#pragma once
#include "bmpc_device.cuh"
#include "bmpc_gait.h"

#include "bmpc_kernels_common.cuh"
#include "bmpc_kernels_setup.cuh"
#include "bmpc_kernels_lq.cuh"
#include "bmpc_kernels_project.cuh"
#include "bmpc_kernels_riccati.cuh"
#include "bmpc_kernels_policy.cuh"
#include "bmpc_kernels_linesearch.cuh"
#include "bmpc_kernels_io.cuh"

// Thread-level model mathematics of the centroidal OCP (sm_100a, FP64).
//
// Hand-derived replacements for the CppADCodeGen libraries the reference dlopens:
//   - PinocchioCentroidalDynamicsAD flow map + exact Jacobians (call site
//     ocs2_bipedal_robot/src/dynamics/BipedalRobotDynamicsAD.cpp:46-56),
//   - PinocchioEndEffectorKinematicsCppAd contact velocity + Jacobians (built at
//     ocs2_bipedal_robot/src/BipedalRobotInterface.cpp:169-178, used by
//     src/constraint/EndEffectorLinearConstraint.cpp:89-111).
// The derivatives are analytic (spatial-algebra recursion over the kinematic tree), not dual numbers:
//   d(A(q) v)/dq_k |v = S_k x* h_sub(k) - I_sub(k) (S_k x v_parent(k))      (momentum, about the world origin)
//   d(pdot_i)/dq_k |v = a_k x u_w + w_k x (a_k x (p_i - o_k))                (contact point velocity)
// and are checked against the oracle's forward-mode duals in tests/.
#pragma once
#include <cuda_runtime.h>
#include "bmpc_model.h"

namespace bmpc {

__constant__ DevModel c_model;

struct v3 { double x, y, z; };
__device__ __forceinline__ v3 mk(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ v3 operator+(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ v3 operator-(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ v3 operator*(double s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
struct m3 { double m[9]; };
__device__ __forceinline__ v3 mul(const m3& R, v3 v) { return mk(R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z, R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z); }
__device__ __forceinline__ v3 mulc(const double* R, const double* v) { return mk(R[0] * v[0] + R[1] * v[1] + R[2] * v[2], R[3] * v[0] + R[4] * v[1] + R[5] * v[2], R[6] * v[0] + R[7] * v[1] + R[8] * v[2]); }
__device__ __forceinline__ m3 mul(const m3& A, const m3& B) {
  m3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
__device__ __forceinline__ m3 mulc(const m3& A, const double* B) {
  m3 C;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C.m[3 * i + j] = A.m[3 * i] * B[j] + A.m[3 * i + 1] * B[3 + j] + A.m[3 * i + 2] * B[6 + j];
  return C;
}
// symmetric 3x3 stored as xx, xy, xz, yy, yz, zz
struct s3 { double xx, xy, xz, yy, yz, zz; };
__device__ __forceinline__ v3 mul(const s3& I, v3 v) { return mk(I.xx * v.x + I.xy * v.y + I.xz * v.z, I.xy * v.x + I.yy * v.y + I.yz * v.z, I.xz * v.x + I.yz * v.y + I.zz * v.z); }
__device__ __forceinline__ s3 operator+(const s3& a, const s3& b) { s3 r; r.xx = a.xx + b.xx; r.xy = a.xy + b.xy; r.xz = a.xz + b.xz; r.yy = a.yy + b.yy; r.yz = a.yz + b.yz; r.zz = a.zz + b.zz; return r; }
// R I R^T for a (symmetric) body inertia given row-major 3x3 in constant memory
__device__ __forceinline__ s3 rotate_inertia(const m3& R, const double* I) {
  m3 T = mulc(R, I);
  s3 r;
  r.xx = T.m[0] * R.m[0] + T.m[1] * R.m[1] + T.m[2] * R.m[2];
  r.xy = T.m[0] * R.m[3] + T.m[1] * R.m[4] + T.m[2] * R.m[5];
  r.xz = T.m[0] * R.m[6] + T.m[1] * R.m[7] + T.m[2] * R.m[8];
  r.yy = T.m[3] * R.m[3] + T.m[4] * R.m[4] + T.m[5] * R.m[5];
  r.yz = T.m[3] * R.m[6] + T.m[4] * R.m[7] + T.m[5] * R.m[8];
  r.zz = T.m[6] * R.m[6] + T.m[7] * R.m[7] + T.m[8] * R.m[8];
  return r;
}
// spatial inertia about the world origin: mass, first moment h = m c, inertia about the origin
struct SI { double M; v3 h; s3 I; };
__device__ __forceinline__ SI body_si(double m, v3 c, const s3& Ic) {
  SI s; s.M = m; s.h = m * c;
  const double cc = dot(c, c);
  s.I.xx = Ic.xx + m * (cc - c.x * c.x); s.I.xy = Ic.xy - m * c.x * c.y; s.I.xz = Ic.xz - m * c.x * c.z;
  s.I.yy = Ic.yy + m * (cc - c.y * c.y); s.I.yz = Ic.yz - m * c.y * c.z; s.I.zz = Ic.zz + m * (cc - c.z * c.z);
  return s;
}
__device__ __forceinline__ SI operator+(const SI& a, const SI& b) { SI s; s.M = a.M + b.M; s.h = a.h + b.h; s.I = a.I + b.I; return s; }
struct Mom { v3 n, p; };   // moment about the world origin, linear momentum
__device__ __forceinline__ Mom operator+(const Mom& a, const Mom& b) { Mom r; r.n = a.n + b.n; r.p = a.p + b.p; return r; }
// momentum of a composite body moving with twist (w, vO)
__device__ __forceinline__ Mom si_apply(const SI& s, v3 w, v3 vO) { Mom r; r.p = s.M * vO + cross(w, s.h); r.n = mul(s.I, w) + cross(s.h, vO); return r; }

__device__ __forceinline__ m3 rodrigues(const double* a, double ang) {
  double s, c; sincos(ang, &s, &c);
  const double t = 1.0 - c;
  m3 r;
  r.m[0] = c + t * a[0] * a[0]; r.m[1] = t * a[0] * a[1] - s * a[2]; r.m[2] = t * a[0] * a[2] + s * a[1];
  r.m[3] = t * a[0] * a[1] + s * a[2]; r.m[4] = c + t * a[1] * a[1]; r.m[5] = t * a[1] * a[2] - s * a[0];
  r.m[6] = t * a[0] * a[2] - s * a[1]; r.m[7] = t * a[1] * a[2] + s * a[0]; r.m[8] = c + t * a[2] * a[2];
  return r;
}
__device__ __forceinline__ void inv3(const double* a, double* r) {
  const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  const double id = 1.0 / (a[0] * c00 + a[1] * c01 + a[2] * c02);
  r[0] = c00 * id; r[1] = (a[2] * a[7] - a[1] * a[8]) * id; r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id; r[4] = (a[0] * a[8] - a[2] * a[6]) * id; r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id; r[7] = (a[1] * a[6] - a[0] * a[7]) * id; r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// gait/MotionPhaseDefinition.h:57-76 : mode -> leg in stance (both sole points of a foot share the flag)
__device__ __forceinline__ bool leg_in_stance(int mode, int leg) { return leg == 0 ? (mode == 1 || mode == 3) : (mode == 2 || mode == 3); }

template <int NJ>
struct Dims {
  static constexpr int NX = 12 + NJ, NU = 12 + NJ, NXA = NX - 3, NL = NJ / 2;
  // compact LQ record produced by the LQ kernel (doubles)
  // constraint rows of the contact velocities, MAXROWS = 12: either the raw rows in upstream's stacking order (two sole points per foot:
  // 3 zero-velocity rows per closed contact, 1 normal-velocity row per open contact; consumed by the FullPivLU projection) or the per-foot
  // orthogonally compressed rows (<= 10, consumed by the Moore-Penrose projection)
  static constexpr int MAXROWS = 12;
  static constexpr int R_B = 0, R_AD = R_B + NX, R_BD = R_AD + 9 * NXA, R_Q = R_BD + 9 * NU, R_R = R_Q + NX, R_HB = R_R + NU,
                       R_CV = R_HB + 24, R_DV = R_CV + MAXROWS * NXA, R_EV = R_DV + MAXROWS * NJ, R_MISC = R_EV + MAXROWS, R_FO = R_MISC + 12,
                       REC = ((R_FO + 12 + 3) / 4) * 4;
  // misc slots
  static constexpr int M_DT = 0, M_DQ = 1, M_DR = 2, M_MODE = 3, M_NROWS = 4, M_TYPE = 5, M_PCOST = 6, M_PDYN = 7, M_PEQ = 8;
  // projection record: Pxj[NJ][NXA], Pej[NJ], N[NJ][8], meta (mj, rank_flag), Pxj[:, base height]
  // P_PX8: the base-height column (state 8) of Pxj; non-zero only with positionErrorGain != 0 (the other base-position columns vanish always)
  static constexpr int P_PX = 0, P_PE = P_PX + NJ * NXA, P_N = P_PE + NJ, P_FO = P_N + NJ * 8, P_META = P_FO + 12, P_PX8 = P_META + 4, PREC = ((P_PX8 + NJ + 3) / 4) * 4;
};

// map a state index (0..NX-1, not 6..8) to its column in the "active x" set X = {0..5, 9..NX-1}
__device__ __forceinline__ int xcol(int s) { return s < 6 ? s : s - 3; }

// ------------------------------------------------------------------------------------------------ streaming flow map (values only, registers only)
// Flow map / contact velocities for the line search and the rollout, structured so that nothing per-joint is stored:
//   sum_j A_j(q) qd_j = sum_i I_i (w_i^J, v_i^J),  (w_i^J, v_i^J) = sum_{j on the path to body i} qd_j S_j   (joint-induced link twist)
// so one root-to-leaf pass per leg accumulates the total spatial inertia, the joint-induced momentum and the tip twists; no composite
// inertias, no per-joint arrays, no local memory.  xb = x[0:12], qj = x[12:], uf = u[0:12], qd = u[12:]; f = rows 0..11 of the flow map
// (rows 12.. are qd).  Fully unrolled: the model constants become immediate constant-bank operands.
// MODEL_VALUES_UNROLL = 1: the joint loop of a leg stays rolled.  Fully unrolled the function is 7 k instructions of straight-line code that every warp
// runs through once: k_linesearch then waits for instruction fetch (ncu: 29 % of its stall samples are no_instruction); rolled, the model constants
// are indexed constant-bank loads instead of immediate operands but the body stays in the instruction cache.
#ifndef MODEL_VALUES_UNROLL
#define MODEL_VALUES_UNROLL 1
#endif
template <int NJ>
__device__ __noinline__ void model_values(const double (&xb)[12], const double (&qj)[NJ], const double (&uf)[12], const double (&qd)[NJ], double (&f)[12], v3 (&vc)[NCON], v3* pc_out = nullptr) {
  constexpr int NL = Dims<NJ>::NL, MVU = MODEL_VALUES_UNROLL;
  const DevModel& M = c_model;
  const double mass = M.total_mass, imass = 1.0 / mass;
  double sz, cz, sy, cy, sx, cx;
  sincos(xb[9], &sz, &cz); sincos(xb[10], &sy, &cy); sincos(xb[11], &sx, &cx);
  m3 Rb;
  Rb.m[0] = cz * cy; Rb.m[1] = cz * sy * sx - sz * cx; Rb.m[2] = cz * sy * cx + sz * sx;
  Rb.m[3] = sz * cy; Rb.m[4] = sz * sy * sx + cz * cx; Rb.m[5] = sz * sy * cx - cz * sx;
  Rb.m[6] = -sy;     Rb.m[7] = cy * sx;                Rb.m[8] = cy * cx;
  const v3 pb = mk(xb[6], xb[7], xb[8]);
  v3 bax[3];
  bax[0] = mk(0.0, 0.0, 1.0); bax[1] = mk(-sz, cz, 0.0); bax[2] = mk(cz * cy, sz * cy, -sy);
  SI tot = body_si(M.base_mass, mulc(Rb.m, M.base_com) + pb, rotate_inertia(Rb, M.base_inertia));
  Mom hJ; hJ.n = mk(0.0, 0.0, 0.0); hJ.p = mk(0.0, 0.0, 0.0);
  v3 pc[NCON], wtip[2], vtip[2];
#pragma unroll
  for (int leg = 0; leg < 2; ++leg) {
    m3 Rp = Rb; v3 pp = pb;
    v3 wJ = mk(0.0, 0.0, 0.0), vJ = mk(0.0, 0.0, 0.0);
#pragma unroll MVU
    for (int i = 0; i < NL; ++i) {
      const int j = leg * NL + i;
      const v3 o = mulc(Rp.m, M.pj[j]) + pp;
      const m3 Rfix = mulc(Rp, M.Rj[j]);
      const v3 a = mulc(Rfix.m, M.axis[j]);
      const m3 Rw = mul(Rfix, rodrigues(M.axis[j], qj[j]));
      const SI bj = body_si(M.mass[j], mulc(Rw.m, M.com[j]) + o, rotate_inertia(Rw, M.inertia[j]));
      tot = tot + bj;
      wJ = wJ + qd[j] * a; vJ = vJ + qd[j] * cross(o, a);
      hJ = hJ + si_apply(bj, wJ, vJ);
      Rp = Rw; pp = o;
    }
    pc[2 * leg] = mulc(Rp.m, M.coff[2 * leg]) + pp; pc[2 * leg + 1] = mulc(Rp.m, M.coff[2 * leg + 1]) + pp;
    wtip[leg] = wJ; vtip[leg] = vJ;
  }
  const v3 com = imass * tot.h;
  double A22[9], A22i[9], A12[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const Mom m = si_apply(tot, bax[k], cross(pb, bax[k]));
    const v3 an = m.n - cross(com, m.p);
    A22[k] = an.x; A22[3 + k] = an.y; A22[6 + k] = an.z; A12[k] = m.p.x; A12[3 + k] = m.p.y; A12[6 + k] = m.p.z;
  }
  inv3(A22, A22i);
  const v3 ml = mk(mass * xb[0], mass * xb[1], mass * xb[2]) - hJ.p;
  const v3 ma = mk(mass * xb[3], mass * xb[4], mass * xb[5]) - (hJ.n - cross(com, hJ.p));
  const v3 w = mk(A22i[0] * ma.x + A22i[1] * ma.y + A22i[2] * ma.z, A22i[3] * ma.x + A22i[4] * ma.y + A22i[5] * ma.z, A22i[6] * ma.x + A22i[7] * ma.y + A22i[8] * ma.z);
  const v3 vlin = imass * (ml - mk(A12[0] * w.x + A12[1] * w.y + A12[2] * w.z, A12[3] * w.x + A12[4] * w.y + A12[5] * w.z, A12[6] * w.x + A12[7] * w.y + A12[8] * w.z));
  v3 Ftot = mk(0.0, 0.0, 0.0), tau = mk(0.0, 0.0, 0.0);
#pragma unroll
  for (int c = 0; c < NCON; ++c) { const v3 F = mk(uf[3 * c], uf[3 * c + 1], uf[3 * c + 2]); Ftot = Ftot + F; tau = tau + cross(pc[c] - com, F); }
  f[0] = Ftot.x * imass; f[1] = Ftot.y * imass; f[2] = Ftot.z * imass - 9.81;
  f[3] = tau.x * imass; f[4] = tau.y * imass; f[5] = tau.z * imass;
  f[6] = vlin.x; f[7] = vlin.y; f[8] = vlin.z; f[9] = w.x; f[10] = w.y; f[11] = w.z;
  const v3 we3 = w.x * bax[0] + w.y * bax[1] + w.z * bax[2];
  const v3 ve3 = vlin + w.x * cross(pb, bax[0]) + w.y * cross(pb, bax[1]) + w.z * cross(pb, bax[2]);
#pragma unroll
  for (int c = 0; c < NCON; ++c) vc[c] = cross(we3 + wtip[c / 2], pc[c]) + ve3 + vtip[c / 2];
  if (pc_out) {
#pragma unroll
    for (int c = 0; c < NCON; ++c) pc_out[c] = pc[c];
  }
}

// ------------------------------------------------------------------------------------------------ base record (values only)
// Everything the Jacobian columns need, computed once per (stage, RK2 evaluation) by the base pass (lane = joint) and consumed by the
// column pass (lane = column) of k_lq_pack.  Layout in doubles:
template <int NJ>
struct BaseDims {
  static constexpr int NX = Dims<NJ>::NX;
  static constexpr int B_PB = 0, B_BAX = 3, B_WE = 12, B_VE = 24, B_TOT = 36, B_HTOT = 46, B_COM = 52, B_A22I = 55, B_A12 = 64, B_ALE = 73, B_AAE = 82,
                       B_PC = 91, B_VC = 103, B_F = 115, B_FTOT = B_F + NX, B_J = B_FTOT + 3, JS = 34, BASE = ((B_J + JS * NJ + 3) / 4) * 4;
  static constexpr int J_O = 0, J_A = 3, J_SI = 6, J_AL = 16, J_AA = 19, J_W = 22, J_V = 25, J_HN = 28, J_HP = 31;
};
__device__ __forceinline__ void st3(double* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ v3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st_si(double* p, const SI& s) { p[0] = s.M; st3(p + 1, s.h); p[4] = s.I.xx; p[5] = s.I.xy; p[6] = s.I.xz; p[7] = s.I.yy; p[8] = s.I.yz; p[9] = s.I.zz; }
__device__ __forceinline__ SI ld_si(const double* p) { SI s; s.M = p[0]; s.h = ld3(p + 1); s.I.xx = p[4]; s.I.xy = p[5]; s.I.xz = p[6]; s.I.yy = p[7]; s.I.yz = p[8]; s.I.zz = p[9]; return s; }

// ------------------------------------------------------------------------------------------------ warp-cooperative base record
// Computed by one warp with lane j = leg joint j (two serial chains of NL joints): per-joint data stays in
// the lane's registers, parents / children are reached with shuffles, level by level along each leg.  The record is written to shared memory.
__device__ __forceinline__ v3 shfl_v3(v3 a, int src) { return mk(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src), __shfl_sync(0xffffffffu, a.z, src)); }
__device__ __forceinline__ m3 shfl_m3(const m3& a, int src) { m3 r; for (int i = 0; i < 9; ++i) r.m[i] = __shfl_sync(0xffffffffu, a.m[i], src); return r; }
__device__ __forceinline__ SI shfl_si(const SI& a, int src) {
  SI r; r.M = __shfl_sync(0xffffffffu, a.M, src); r.h = shfl_v3(a.h, src);
  r.I.xx = __shfl_sync(0xffffffffu, a.I.xx, src); r.I.xy = __shfl_sync(0xffffffffu, a.I.xy, src); r.I.xz = __shfl_sync(0xffffffffu, a.I.xz, src);
  r.I.yy = __shfl_sync(0xffffffffu, a.I.yy, src); r.I.yz = __shfl_sync(0xffffffffu, a.I.yz, src); r.I.zz = __shfl_sync(0xffffffffu, a.I.zz, src);
  return r;
}
__device__ __forceinline__ Mom shfl_mom(const Mom& a, int src) { Mom r; r.n = shfl_v3(a.n, src); r.p = shfl_v3(a.p, src); return r; }
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

// jc: per-joint constants of this lane's joint in shared memory, layout [Rj 9 | pj 3 | axis 3 | mass 1 | com 3 | inertia 9] (28 doubles)
// SEG = 32: one stage per warp (lane = joint).  SEG = 16: two stages per warp, one per half-warp (lane & 15 = joint); x, u, base and jc are
// then per-half pointers and every shuffle stays inside the lane's 16-lane segment (segment base hb).
// sum over the NJ joint lanes of a segment, result in every lane of the segment.  Power-of-two segments: xor butterfly (idle lanes hold 0).
// Other segment sizes (SEG == NJ): suffix doubling along each leg (NL lanes), then the two leg totals are added.
template <int SEG, int NL>
__device__ __forceinline__ double seg_sum(double v, int lane_w) {
  if constexpr ((SEG & (SEG - 1)) == 0) {
    for (int o = SEG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  } else {
    static_assert(SEG == 2 * NL, "non power-of-two segments hold exactly the two legs");
    const int jl = lane_w % SEG, hb = lane_w - jl, li = jl % NL;
#pragma unroll
    for (int o = 1; o < NL; o <<= 1) { const double t = __shfl_sync(0xffffffffu, v, lane_w + o); if (li + o < NL) v += t; }
    return __shfl_sync(0xffffffffu, v, hb) + __shfl_sync(0xffffffffu, v, hb + NL);
  }
}
template <int NJ, int SEG = 32>
__device__ __forceinline__ void warp_model_base(const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ base, int lane_w, const double* __restrict__ jc, bool do_write = true) {
  using BD = BaseDims<NJ>;
  constexpr int NL = Dims<NJ>::NL;
  static_assert(NJ <= SEG, "one lane per joint");
  const DevModel& M = c_model;
  const double mass = M.total_mass, imass = 1.0 / mass;
  const int lane = lane_w % SEG, hb = lane_w - lane;   // lane: index inside the segment; lane_w +- 1 shuffles stay inside a leg
  const bool act = lane < NJ && hb + SEG <= 32;        // lanes beyond the last complete segment (SEG = 10: lanes 30, 31) only tag along
  constexpr int WR = (SEG > NJ) ? SEG - 1 : 0;         // lane of the segment that writes the shared part of the record
  const int j = act ? lane : 0, lvl_own = j % NL;
  double sz, cz, sy, cy, sx, cx;
  {   // one sincos per lane (lane % 3 picks the Euler angle), broadcast with shuffles instead of three evaluations on every lane
    double s_, c_; sincos(x[9 + lane % 3], &s_, &c_);
    sz = __shfl_sync(0xffffffffu, s_, hb + 0); cz = __shfl_sync(0xffffffffu, c_, hb + 0);
    sy = __shfl_sync(0xffffffffu, s_, hb + 1); cy = __shfl_sync(0xffffffffu, c_, hb + 1);
    sx = __shfl_sync(0xffffffffu, s_, hb + 2); cx = __shfl_sync(0xffffffffu, c_, hb + 2);
  }
  m3 Rb;
  Rb.m[0] = cz * cy; Rb.m[1] = cz * sy * sx - sz * cx; Rb.m[2] = cz * sy * cx + sz * sx;
  Rb.m[3] = sz * cy; Rb.m[4] = sz * sy * sx + cz * cx; Rb.m[5] = sz * sy * cx - cz * sx;
  Rb.m[6] = -sy;     Rb.m[7] = cy * sx;                Rb.m[8] = cy * cx;
  const v3 pb = mk(x[6], x[7], x[8]);
  v3 bax[3];
  bax[0] = mk(0.0, 0.0, 1.0); bax[1] = mk(-sz, cz, 0.0); bax[2] = mk(cz * cy, sz * cy, -sy);
  // ---- forward kinematics, one level of both legs at a time
  // local transform of the own joint (parallel over lanes), then the chain R = Rp Rloc, o = Rp pj + pp level by level (shuffles)
  m3 Rloc = mul(*reinterpret_cast<const m3*>(jc), rodrigues(jc + 12, x[12 + j]));
  m3 R = Rb; v3 o = pb;
#pragma unroll 1
  for (int lvl = 0; lvl < NL; ++lvl) {
    m3 Rp = shfl_m3(R, lane_w - 1); v3 pp = shfl_v3(o, lane_w - 1);
    if (lvl == 0) { Rp = Rb; pp = pb; }
    if (act && lvl_own == lvl) { o = mulc(Rp.m, jc + 9) + pp; R = mul(Rp, Rloc); }
  }
  // joint axis in world = R * axis (the rotation about the axis leaves it invariant), body COM and inertia
  const v3 a = mulc(R.m, jc + 12);
  const v3 cb = mulc(R.m, jc + 16) + o;
  const s3 Icb = rotate_inertia(R, jc + 19);
  // contact points live on the tip joints of the legs
  v3 pc[NCON];
#pragma unroll
  for (int c = 0; c < NCON; ++c) { const v3 local = mulc(R.m, M.coff[c]) + o; pc[c] = shfl_v3(local, hb + (c / 2) * NL + NL - 1); }
  const v3 cbase = mulc(Rb.m, M.base_com) + pb;
  const s3 Ibase = rotate_inertia(Rb, M.base_inertia);
  // ---- composite inertias, leaf to root
  SI comp = body_si(act ? jc[15] : 0.0, cb, Icb);
#pragma unroll 1
  for (int lvl = NL - 2; lvl >= 0; --lvl) {
    const SI child = shfl_si(comp, lane_w + 1);
    if (act && lvl_own == lvl) comp = comp + child;
  }
  const SI tot = body_si(M.base_mass, cbase, Ibase) + shfl_si(comp, hb) + shfl_si(comp, hb + NL);
  const v3 com = imass * tot.h;
  v3 Alin_e[3], Aang_e[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { const Mom m = si_apply(tot, bax[k], cross(pb, bax[k])); Alin_e[k] = m.p; Aang_e[k] = m.n - cross(com, m.p); }
  double A22[9], A22i[9], A12[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) { A22[k] = Aang_e[k].x; A22[3 + k] = Aang_e[k].y; A22[6 + k] = Aang_e[k].z; A12[k] = Alin_e[k].x; A12[3 + k] = Alin_e[k].y; A12[6 + k] = Alin_e[k].z; }
  inv3(A22, A22i);
  v3 Alin, Aang;
  { const Mom m = si_apply(comp, a, cross(o, a)); Alin = m.p; Aang = m.n - cross(com, m.p); }
  // ---- generalized velocity: v_b = A_b^-1 (m h - sum_j A_j qd_j)
  const double qd = act ? u[12 + j] : 0.0;
  v3 ml = mk(mass * x[0] - seg_sum<SEG, NL>(qd * Alin.x, lane_w), mass * x[1] - seg_sum<SEG, NL>(qd * Alin.y, lane_w), mass * x[2] - seg_sum<SEG, NL>(qd * Alin.z, lane_w));
  v3 ma = mk(mass * x[3] - seg_sum<SEG, NL>(qd * Aang.x, lane_w), mass * x[4] - seg_sum<SEG, NL>(qd * Aang.y, lane_w), mass * x[5] - seg_sum<SEG, NL>(qd * Aang.z, lane_w));
  const v3 w = mk(A22i[0] * ma.x + A22i[1] * ma.y + A22i[2] * ma.z, A22i[3] * ma.x + A22i[4] * ma.y + A22i[5] * ma.z, A22i[6] * ma.x + A22i[7] * ma.y + A22i[8] * ma.z);
  const v3 vlin = imass * (ml - mk(A12[0] * w.x + A12[1] * w.y + A12[2] * w.z, A12[3] * w.x + A12[4] * w.y + A12[5] * w.z, A12[6] * w.x + A12[7] * w.y + A12[8] * w.z));
  v3 Ftot = mk(0.0, 0.0, 0.0), tau = mk(0.0, 0.0, 0.0);
#pragma unroll
  for (int c = 0; c < NCON; ++c) { const v3 F = mk(u[3 * c], u[3 * c + 1], u[3 * c + 2]); Ftot = Ftot + F; tau = tau + cross(pc[c] - com, F); }
  // ---- link twists, root to leaf
  const double wr[3] = {w.x, w.y, w.z};
  v3 we[4], ve[4];
  we[0] = mk(0.0, 0.0, 0.0); ve[0] = vlin;
#pragma unroll
  for (int k = 0; k < 3; ++k) { we[k + 1] = we[k] + wr[k] * bax[k]; ve[k + 1] = ve[k] + wr[k] * cross(pb, bax[k]); }
  v3 wj = we[3], vj = ve[3];
#pragma unroll 1
  for (int lvl = 0; lvl < NL; ++lvl) {
    v3 wp = shfl_v3(wj, lane_w - 1), vp = shfl_v3(vj, lane_w - 1);
    if (lvl == 0) { wp = we[3]; vp = ve[3]; }
    if (act && lvl_own == lvl) { wj = wp + qd * a; vj = vp + qd * cross(o, a); }
  }
  v3 vc[NCON];
#pragma unroll
  for (int c = 0; c < NCON; ++c) { const v3 local = cross(wj, pc[c]) + vj; vc[c] = shfl_v3(local, hb + (c / 2) * NL + NL - 1); }
  // ---- subtree momenta, leaf to root
  Mom hs;
  { const double mj = act ? jc[15] : 0.0; hs.p = mj * (vj + cross(wj, cb)); hs.n = mul(Icb, wj) + cross(cb, hs.p); }
#pragma unroll 1
  for (int lvl = NL - 2; lvl >= 0; --lvl) {
    const Mom child = shfl_mom(hs, lane_w + 1);
    if (act && lvl_own == lvl) hs = hs + child;
  }
  Mom htot;
  { Mom bm; bm.p = M.base_mass * (ve[3] + cross(we[3], cbase)); bm.n = mul(Ibase, we[3]) + cross(cbase, bm.p); htot = bm + shfl_mom(hs, hb) + shfl_mom(hs, hb + NL); }
  // ---- write the record: per-joint part by the joint's lane, shared part by lane WR of the segment (an idle lane when there is one)
  if (act && do_write) {
    double* J = base + BD::B_J + BD::JS * j;
    st3(J + BD::J_O, o); st3(J + BD::J_A, a); st_si(J + BD::J_SI, comp); st3(J + BD::J_AL, Alin); st3(J + BD::J_AA, Aang);
    st3(J + BD::J_W, wj); st3(J + BD::J_V, vj); st3(J + BD::J_HN, hs.n); st3(J + BD::J_HP, hs.p);
    base[BD::B_F + 12 + j] = qd;
  }
  if (lane == WR && hb + SEG <= 32 && do_write) {
    st3(base + BD::B_PB, pb);
#pragma unroll
    for (int k = 0; k < 3; ++k) { st3(base + BD::B_BAX + 3 * k, bax[k]); st3(base + BD::B_ALE + 3 * k, Alin_e[k]); st3(base + BD::B_AAE + 3 * k, Aang_e[k]); }
#pragma unroll
    for (int k = 0; k < 4; ++k) { st3(base + BD::B_WE + 3 * k, we[k]); st3(base + BD::B_VE + 3 * k, ve[k]); st3(base + BD::B_PC + 3 * k, pc[k]); st3(base + BD::B_VC + 3 * k, vc[k]); }
    st_si(base + BD::B_TOT, tot); st3(base + BD::B_HTOT, htot.n); st3(base + BD::B_HTOT + 3, htot.p); st3(base + BD::B_COM, com);
#pragma unroll
    for (int i = 0; i < 9; ++i) { base[BD::B_A22I + i] = A22i[i]; base[BD::B_A12 + i] = A12[i]; }
    double* f = base + BD::B_F;
    f[0] = Ftot.x * imass; f[1] = Ftot.y * imass; f[2] = Ftot.z * imass - 9.81;
    f[3] = tau.x * imass; f[4] = tau.y * imass; f[5] = tau.z * imass;
    f[6] = vlin.x; f[7] = vlin.y; f[8] = vlin.z; f[9] = w.x; f[10] = w.y; f[11] = w.z;
    st3(base + BD::B_FTOT, Ftot);
  }
}

// relaxed log barrier [UPSTREAM RelaxedBarrierPenalty], settings task.info:280-287
__device__ __forceinline__ void barrier_penalty(double h, double& p, double& dp, double& ddp) {
  const double mu = c_model.bar_mu, de = c_model.bar_delta;
  if (h > de) { const double ih = __drcp_rn(h); p = -mu * log(h); dp = -mu * ih; ddp = mu * ih * ih; }
  else { const double ide = __drcp_rn(de), dh = (h - 2.0 * de) * ide; p = mu * (-log(de) + 0.5 * dh * dh - 0.5); dp = mu * dh * ide; ddp = mu * ide * ide; }
}
// friction cone value (constraint/FrictionConeConstraint.cpp:160-166)
__device__ __forceinline__ double friction_cone(double fx, double fy, double fz) {
  return c_model.mu_f * (fz + c_model.fr_grip) - sqrt(fx * fx + fy * fy + c_model.fr_reg);
}

// stage cost value (tracking + soft friction cones), cost/BipedalRobotQuadraticTrackingCost.h:57-63, common/utils.h:63-77
template <int NJ>
__device__ double stage_cost_value(int mode, const double* x, const double* u, const double* xref) {
  constexpr int NX = Dims<NJ>::NX;
  const DevModel& M = c_model;
  const bool st0 = leg_in_stance(mode, 0), st1 = leg_in_stance(mode, 1);
  const int nst = 2 * (int(st0) + int(st1));
  const double fz = nst > 0 ? M.total_mass * 9.81 / nst : 0.0;
  double c = 0.0;
#pragma unroll 1
  for (int i = 0; i < NX; ++i) { const double d = x[i] - xref[i]; c += 0.5 * M.Qdiag[i] * d * d; }
#pragma unroll 1
  for (int k = 0; k < NCON; ++k) {
    const bool st = (k / 2 == 0) ? st0 : st1;
    const double dx = u[3 * k], dy = u[3 * k + 1], dz = u[3 * k + 2] - (st ? fz : 0.0);
    c += 0.5 * (M.Rforce[3 * k] * dx * dx + M.Rforce[3 * k + 1] * dy * dy + M.Rforce[3 * k + 2] * dz * dz);
    if (st) { double p, dp, ddp; barrier_penalty(friction_cone(u[3 * k], u[3 * k + 1], u[3 * k + 2]), p, dp, ddp); c += p; }
  }
#pragma unroll 1
  for (int i = 0; i < NJ; ++i) {
    double s = 0.0;
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) s += M.Rjoint[i * NJ + j] * u[12 + j];
    c += 0.5 * u[12 + i] * s;
  }
  return c;
}

}  // namespace bmpc

// K1: fused per-quadrature-point crystal-plasticity update (sm_100a, fp64).
//
// One launch replaces ExaCMechModel::ModelSetup (src/mechanics_ecmech.cpp:192-258):
//   StateVarsSetup / StressSetup copies      src/mechanics_model.cpp:158-180
//   matGrad = 0, vel_grad = 0 memsets         src/mechanics_ecmech.cpp:212,216
//   exaconstit::kernel::grad_calc             src/mechanics_kernels.cpp:7-78
//   kernel_setup                              src/mechanics_ecmech.cpp:22-100
//   ecmech::matModelBase::getResponseECM      src/mechanics_ecmech.cpp:183 (external ExaCMech evptn model)
//   kernel_postprocessing (+ transpose)       src/mechanics_ecmech.cpp:106-172
// Thread = quadrature point, 8 lanes = element: the velocity gradient comes from the nodal
// velocities through the register butterfly of exab200_common.cuh; the 35 doubles/qpt of scratch
// arrays the reference allocates (src/mechanics_ecmech.hpp:56-64) never exist.  The per-point
// arithmetic lives in material_point.hpp (shared with the host-compiled unit check of the test-suite).
#pragma once
#include "exab200_common.cuh"
#include "material_point.hpp"

namespace exab {

// ------------------------------------------------------------------------------------------
// K1 kernel.  MODE as in k_operator.cuh (LVEC: vel is an L-vector gathered through e2n; EVEC:
// vel is the E-vector the reference's ModelSetup receives).  Layouts: QFs (vdim,q,e),
// jac (3,3,q,e), matgrad[(pt)*36 + j*6 + i] = d sigma_i / d eps_j (after the reference's transpose,
// src/mechanics_ecmech.cpp:159-169; layout 0 reproduces the EA-on-device quirk, :155; layout 2 = compact record).
// The material description is a kernel parameter: its tables are constant-bank operands.
// The 8x8 Jacobian of each thread lives in shared memory, entry k of thread t at smem[k * kJS + t]
// (conflict-free, and off the local-memory / L2 path).
// fail_count is incremented for points whose local solve did not converge.
// ------------------------------------------------------------------------------------------
#ifndef EXAB_K1_THREADS
#define EXAB_K1_THREADS 128
#endif
constexpr int kJS = EXAB_K1_THREADS;
constexpr int kK1SmemBytes = 64 * kJS * 8;  // 64 KB: one 8x8 Jacobian per thread
constexpr int kIdleRecord = 96;            // doubles per idle thread: history (<= 40) | stress (6) | tangent (36)
template <int NSLIP, int KIN, int MODE, int MINB>
__global__ void __launch_bounds__(kJS, MINB) k_model_setup(const __grid_constant__ MatDev m, double dt,
                                                           const double* __restrict__ jac,
                                                           const double* __restrict__ vel, const int* __restrict__ e2n,
                                                           long nnodes, const double* __restrict__ stress0,
                                                           const double* __restrict__ hist0, double* __restrict__ stress1,
                                                           double* __restrict__ hist1, double* __restrict__ matgrad,
                                                           long nelems, int layout, int* __restrict__ fail_count,
                                                           double* __restrict__ idle_scratch) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  constexpr int nsv = NSLIP + iH_Gdot + 2;
  // ---- grad_calc: velocity gradient at this point ----
  double v0 = 0, v1 = 0, v2 = 0;
  if (active) {
    if (MODE == 0) {
      const long nid = e2n[e * 8 + lex_to_native(lane)];
      v0 = vel[nid]; v1 = vel[nnodes + nid]; v2 = vel[2 * nnodes + nid];
    } else {
      const long o = e * 24 + lex_to_native(lane);
      v0 = vel[o]; v1 = vel[o + 8]; v2 = vel[o + 16];
    }
  }
  double d[3][3];
  nodal_to_qp_grad(v0, lane, d[0][0], d[0][1], d[0][2]);
  nodal_to_qp_grad(v1, lane, d[1][0], d[1][1], d[1][2]);
  nodal_to_qp_grad(v2, lane, d[2][0], d[2][1], d[2][2]);
#ifdef EXAB_K1_LOCKSTEP
  // CTA-lockstep solve loop: every thread of the CTA has to run it (a barrier per pass).  The inactive threads of the
  // last CTA shadow point 0 with a zero velocity gradient (d = 0 from the butterfly of zeros: a one-pass solve) and
  // write into a scratch record instead of the output arrays.
  const long p = active ? e * 8 + lane : 0;
#else
  if (!active) return;
  const long p = e * 8 + lane;
#endif
  double L[3][3];
  {
    double J[9], adj[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = jac[p * 9 + i];
    const double idet = 1.0 / adjugate(J, adj);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int t = 0; t < 3; ++t) L[i][t] = (d[i][0] * adj[t] + d[i][1] * adj[3 + t] + d[i][2] * adj[6 + t]) * idet;
  }
  extern __shared__ double smJ[];
  double* h1p = hist1 + p * nsv;
  double* s1p = stress1 + p * 6;
  double* kp = matgrad + p * (layout == mat::kTangentCompact ? 32 : 36);
#ifdef EXAB_K1_LOCKSTEP
  if (!active) {
    double* rec = idle_scratch + (long)threadIdx.x * kIdleRecord;
    h1p = rec; s1p = rec + 48; kp = rec + 56;
  }
#endif
  const int nf = mat::update_point<NSLIP, KIN, kJS>(m, dt, L, hist0 + p * nsv, stress0 + p * 6, h1p, s1p, kp, layout,
                                                    smJ + threadIdx.x);
  if (nf < 0 && active) atomicAdd(fail_count, 1);
}

// init_state_vars (src/mechanics_ecmech.hpp:249-300): overwrite the ExaCMech-owned history slots
// with the model's initial values; quaternion slots (set from the grain file) are left alone.
__global__ void k_hist_init(const __grid_constant__ MatDev m, double* __restrict__ hist, long npts) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  double* h = hist + p * m.nhist;
  h[iH_shrateEff] = 0.0;
  h[iH_shrEff] = 0.0;
  h[iH_flowStr] = 0.0;
  h[iH_nFEval] = 0.0;
  h[iH_H] = (m.kin == KIN_KMBALD) ? log(m.rho_dd_init) : m.kappa0;
  h[m.nhist - 2] = 1.0;
  h[m.nhist - 1] = 0.0;
  for (int j = 0; j < 5; ++j) h[iH_E + j] = 0.0;
  for (int j = 0; j < m.nslip; ++j) h[iH_Gdot + j] = 0.0;
}

// calcDpMat (src/mechanics_ecmech.hpp:303-357): D^p = sum gdot P rotated to the sample frame,
// written as a full 3x3 (vdim 9) quadrature function like the reference.
__global__ void k_calc_dp(const __grid_constant__ MatDev m, const double* __restrict__ hist, double* __restrict__ dp9, long npts) {
  using namespace mat;
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npts) return;
  const double* h = hist + p * m.nhist;
  double dphat[5] = {0, 0, 0, 0, 0}, C[9], dsm[5], s6[6];
  for (int a = 0; a < m.nslip; ++a)
    for (int i = 0; i < 5; ++i) dphat[i] += m.P[a][i] * h[iH_Gdot + a];
  quat_to_tensor(&h[iH_Q], C);
  rot_vecd<false>(C, dphat, dsm);
  vecd_to_svec(dsm, s6);
  double* o = dp9 + p * 9;  // Reshape(3,3,npts) column-major; symmetric so order is immaterial
  o[0] = s6[0]; o[4] = s6[1]; o[8] = s6[2];
  o[5] = o[7] = s6[3];
  o[2] = o[6] = s6[4];
  o[1] = o[3] = s6[5];
}

}  // namespace exab

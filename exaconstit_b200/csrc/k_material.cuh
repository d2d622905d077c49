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
// arrays the reference allocates (src/mechanics_ecmech.hpp:56-64) never exist.
//
// The constitutive algorithm (evptn: 5 deviatoric lattice elastic strains + 3 exponential-map
// rotation increments solved by a trust-region dogleg Newton, backward Euler, hardness advanced
// with beginning-of-step slip rates, Kirchhoff-stress resolved shear stresses, constant-modulus
// EOS p = K(1/V - 1)) is the published ExaCMech formulation; it is implemented here
// independently for the GPU and checked point by point against the CPU oracle.
#pragma once
#include "exab200_common.cuh"

namespace exab {

constexpr int kMaxSlip = 24;
enum { KIN_VOCE = 0, KIN_VOCE_NL = 1, KIN_KMBALD = 2 };
enum { XTAL_FCC = 0, XTAL_BCC = 1, XTAL_HCP = 2 };

// history layout (src/mechanics_ecmech.hpp:136-141,165-185)
constexpr int iH_shrateEff = 0, iH_shrEff = 1, iH_flowStr = 2, iH_nFEval = 3, iH_E = 4, iH_Q = 9, iH_H = 13,
              iH_Gdot = 14;

struct MatDev {
  int xtal, kin, nslip, nhist, withGAthermal, pad_;
  double P[kMaxSlip][5];
  double Q[kMaxSlip][3];
  double Kdiag[5], bulk, gmod, Kvd;  // Kvd: hexagonal volumetric <-> c-axis deviator coupling (0 for cubic)
  double tol, gruneisen, dtde, tK0;
  // Voce power law
  double xm, gam_w0, h0, tausi, taus0, xmprime, xms, gamss0, kappa0;
  double pl_t_min, pl_t_max, pl_max, ln_ovf;  // power-law guards, precomputed on the host
  // KMBalD
  double mu_ref, tau_a, p_exp, q_exp, gam_wo, gam_ro, wrD, k1, k2o, ninv, gamma_o, rho_dd_init;
  double c_1[kMaxSlip], go[kMaxSlip], s_[kMaxSlip];
};

namespace mat {

constexpr double sqr2 = 1.4142135623730951, sqr3 = 1.7320508075688772;
constexpr double sqr2i = 0.7071067811865475, sqr6i = 0.4082482904638631;
constexpr double sqr2b3 = 0.816496580927726, sqr3b2 = 1.224744871391589;
constexpr double idp_tiny_sqrt = 1.0e-90, idp_eps_sqrt = 1.0e-8;
constexpr double gam_ratio_min = 1.0e-60, gam_ratio_ovf = 1.0e45;
constexpr double e_scale = 5.0e-4, r_scale = 1.0e-2;
// The 8x8 Jacobian of each thread lives in shared memory, entry k of thread t at smem[k * kJS + t]
// (conflict-free, and off the local-memory / L2 path).
constexpr int kJS = 128;

__device__ __forceinline__ void svec_to_vecd(const double* s, double* v) {
  v[0] = sqr2i * (s[0] - s[1]);
  v[1] = sqr6i * (2.0 * s[2] - s[0] - s[1]);
  v[2] = sqr2 * s[5];
  v[3] = sqr2 * s[4];
  v[4] = sqr2 * s[3];
}
__device__ __forceinline__ void vecd_to_svec(const double* v, double* s) {
  const double t1 = sqr2i * v[0], t2 = sqr6i * v[1];
  s[0] = t1 - t2;
  s[1] = -t1 - t2;
  s[2] = sqr2b3 * v[1];
  s[3] = sqr2i * v[4];
  s[4] = sqr2i * v[3];
  s[5] = sqr2i * v[2];
}
// 5-vector <-> symmetric tensor as 6 entries (xx,yy,zz,yz,xz,xy)
__device__ __forceinline__ void quat_to_tensor(const double* q, double* c) {
  const double x0 = q[0], x1 = q[1], x2 = q[2], x3 = q[3];
  c[0] = x0 * x0 + x1 * x1 - x2 * x2 - x3 * x3;
  c[1] = 2.0 * (x1 * x2 - x0 * x3);
  c[2] = 2.0 * (x1 * x3 + x0 * x2);
  c[3] = 2.0 * (x1 * x2 + x0 * x3);
  c[4] = x0 * x0 - x1 * x1 + x2 * x2 - x3 * x3;
  c[5] = 2.0 * (x2 * x3 - x0 * x1);
  c[6] = 2.0 * (x1 * x3 - x0 * x2);
  c[7] = 2.0 * (x2 * x3 + x0 * x1);
  c[8] = x0 * x0 - x1 * x1 - x2 * x2 + x3 * x3;
}
// rotate a deviatoric 5-vector: TRANSPOSE=false  v_out = 5vec(C T C^T) ; true  5vec(C^T T C)
template <bool TRANSPOSE>
__device__ __forceinline__ void rot_vecd(const double* C, const double* v, double* out) {
  double s[6];
  vecd_to_svec(v, s);
  const double T[9] = {s[0], s[5], s[4], s[5], s[1], s[3], s[4], s[3], s[2]};
  double A[9];  // A = R T with R = C or C^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) a += (TRANSPOSE ? C[3 * k + i] : C[3 * i + k]) * T[3 * k + j];
      A[3 * i + j] = a;
    }
  // B = A R^T, only the 6 needed entries
  auto ent = [&](int i, int j) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) a += A[3 * i + k] * (TRANSPOSE ? C[3 * k + j] : C[3 * j + k]);
    return a;
  };
  const double b00 = ent(0, 0), b11 = ent(1, 1), b22 = ent(2, 2), b12 = ent(1, 2), b02 = ent(0, 2), b01 = ent(0, 1);
  out[0] = sqr2i * (b00 - b11);
  out[1] = sqr6i * (2.0 * b22 - b00 - b11);
  out[2] = sqr2 * b01;
  out[3] = sqr2 * b02;
  out[4] = sqr2 * b12;
}
// Me(e): 5vec(E W - W E) = Me(e) w   (structure constants of the dev-sym / skew commutator)
__device__ __forceinline__ void comm_Me(const double* e, double M[5][3]) {
  M[0][0] = -e[4];            M[0][1] = -e[3];                 M[0][2] = 2.0 * e[2];
  M[1][0] = -sqr3 * e[4];     M[1][1] = sqr3 * e[3];           M[1][2] = 0.0;
  M[2][0] = e[3];             M[2][1] = -e[4];                 M[2][2] = -2.0 * e[0];
  M[3][0] = -e[2];            M[3][1] = e[0] - sqr3 * e[1];    M[3][2] = e[4];
  M[4][0] = e[0] + sqr3 * e[1]; M[4][1] = e[2];                M[4][2] = -e[3];
}
// Mw(w): Me(e) w = Mw(w) e
__device__ __forceinline__ void comm_Mw(const double* w, double M[5][5]) {
  M[0][0] = 0.0;          M[0][1] = 0.0;            M[0][2] = 2.0 * w[2]; M[0][3] = -w[1];      M[0][4] = -w[0];
  M[1][0] = 0.0;          M[1][1] = 0.0;            M[1][2] = 0.0;        M[1][3] = sqr3 * w[1]; M[1][4] = -sqr3 * w[0];
  M[2][0] = -2.0 * w[2];  M[2][1] = 0.0;            M[2][2] = 0.0;        M[2][3] = w[0];       M[2][4] = -w[1];
  M[3][0] = w[1];         M[3][1] = -sqr3 * w[1];   M[3][2] = -w[0];      M[3][3] = 0.0;        M[3][4] = w[2];
  M[4][0] = w[0];         M[4][1] = sqr3 * w[0];    M[4][2] = w[1];       M[4][3] = -w[2];      M[4][4] = 0.0;
}
// right Jacobian of the exponential map
__device__ __forceinline__ void exp_Jr(const double* xi, double Jm[3][3]) {
  const double th2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2], th = sqrt(th2);
  double a, b;
  if (th < 1e-4) { a = 0.5 - th2 / 24.0; b = 1.0 / 6.0 - th2 / 120.0; }
  else { a = (1.0 - cos(th)) / th2; b = (th - sin(th)) / (th2 * th); }
  const double X[3][3] = {{0.0, -xi[2], xi[1]}, {xi[2], 0.0, -xi[0]}, {-xi[1], xi[0], 0.0}};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double x2 = X[i][0] * X[0][j] + X[i][1] * X[1][j] + X[i][2] * X[2][j];
      Jm[i][j] = (i == j ? 1.0 : 0.0) - a * X[i][j] + b * x2;
    }
}

// ---- kinetics ---------------------------------------------------------------------------
__device__ __forceinline__ void kin_power_law(const MatDev& m, double gam_w, double g, double gi, double xmi, double tau,
                                              double& gdot, double& dgdot_dtau) {
  gdot = 0.0;
  dgdot_dtau = 0.0;
  const double t = tau * gi, at = fabs(t);
  if (at <= m.pl_t_min) return;
  if (at > m.pl_t_max) {  // linear extrapolation beyond the overflow guard
    const double d = gam_w * m.pl_max * xmi * gi;
    const double g0 = gam_w * m.pl_t_max * m.pl_max;
    gdot = (g0 + d * g * (at - m.pl_t_max)) * (t > 0 ? 1.0 : -1.0);
    dgdot_dtau = d;
    return;
  }
  const double pl = exp((xmi - 1.0) * log(at));
  gdot = gam_w * t * pl;
  dgdot_dtau = gam_w * pl * xmi * gi;
}

__device__ __forceinline__ void kin_kmbald(const MatDev& m, double g, double gam_w, double gam_r, double c_e,
                                           double tau, double& gdot, double& dgdot_dtau) {
  gdot = 0.0;
  dgdot_dtau = 0.0;
  const double at = fabs(tau), sgn = tau >= 0 ? 1.0 : -1.0;
  double gAth, g_i;
  if (m.withGAthermal) { gAth = g; g_i = 1.0 / m.tau_a; }
  else { gAth = m.tau_a; g_i = 1.0 / g; }
  if (at <= gAth) return;
  const double at_0 = (at - gAth) * g_i;
  double gdot_r, dgdot_r;
  {
    const double x = (at - gAth) / m.wrD;
    if (x < gam_ratio_min) return;
    if (x < idp_eps_sqrt) { gdot_r = gam_r * x; dgdot_r = gam_r / m.wrD; }
    else { const double ex = exp(-x); gdot_r = gam_r * (1.0 - ex); dgdot_r = gam_r * ex / m.wrD; }
  }
  double gdot_w, dgdot_w;
  if (at_0 >= 1.0) {
    const double xn = c_e * m.p_exp;
    const double lg = xn * log(at_0);
    if (lg > m.ln_ovf) { gdot = sgn * gdot_r; dgdot_dtau = dgdot_r; return; }
    gdot_w = gam_w * exp(lg);
    dgdot_w = gdot_w * xn / at_0 * g_i;
  } else {
    const bool p1 = m.p_exp == 1.0, q1 = m.q_exp == 1.0;
    const double pf = p1 ? at_0 : pow(at_0, m.p_exp);
    const double dpf = p1 ? g_i : m.p_exp * pow(at_0, m.p_exp - 1.0) * g_i;
    const double qa = 1.0 - pf;
    const double ef = exp(-c_e * (q1 ? qa : pow(qa, m.q_exp)));
    const double dqf = q1 ? dpf : m.q_exp * pow(qa, m.q_exp - 1.0) * dpf;
    const double qb = 1.0 + pf;
    const double eb = exp(-c_e * (q1 ? qb : pow(qb, m.q_exp)));
    const double dqb = q1 ? dpf : m.q_exp * pow(qb, m.q_exp - 1.0) * dpf;
    gdot_w = gam_w * (ef - eb);
    dgdot_w = gam_w * c_e * (ef * dqf + eb * dqb);
    if (gdot_w <= gam_ratio_min * gam_w) return;
  }
  const double inv = 1.0 / (gdot_w + gdot_r);
  gdot = sgn * gdot_w * gdot_r * inv;
  dgdot_dtau = (dgdot_w * gdot_r * gdot_r + dgdot_r * gdot_w * gdot_w) * inv * inv;
}

// backward-Euler hardness update with the supplied effective shear rate
__device__ __forceinline__ double kin_update_h(const MatDev& m, double h_n, double dt, double shr) {
  double k2 = m.k2o, sat = m.taus0;
  if (m.kin == KIN_KMBALD) {
    if (shr > idp_tiny_sqrt) k2 = m.k2o * pow(m.gamma_o / shr, m.ninv);
  } else {
    if (shr > idp_tiny_sqrt && m.xms != 0.0) sat = m.taus0 * pow(shr / m.gamss0, m.xms);
  }
  double h = h_n;
  for (int it = 0; it < 50; ++it) {
    double sd, ds;
    if (m.kin == KIN_KMBALD) {
      const double t = exp(-0.5 * h);
      ds = (-0.5 * m.k1 * t) * shr;
      sd = (m.k1 * t - k2) * shr;
    } else if (m.kin == KIN_VOCE_NL && m.xmprime != 1.0) {
      const double r = (sat - h) / (sat - m.tausi);
      const double rp = (r > 0) ? pow(r, m.xmprime) : 0.0;
      ds = (r > 0) ? -m.h0 * m.xmprime * pow(r, m.xmprime - 1.0) / (sat - m.tausi) * shr : 0.0;
      sd = m.h0 * rp * shr;
    } else {
      const double t1 = m.h0 / (sat - m.tausi);
      ds = -t1 * shr;
      sd = t1 * (sat - h) * shr;
    }
    const double r = h - h_n - dt * sd;
    const double dh = -r / (1.0 - dt * ds);
    h += dh;
    if (fabs(dh) <= 1e-14 * fabs(h) + 1e-300) break;
  }
  return h;
}

// ---- dense LU (n = 8) with partial pivoting on a local array ------------------------------
#define JIDX(i, j) (((i) * 8 + (j)) * kJS)
__device__ __noinline__ bool lu_factor8(double* A, int* piv) {
  for (int k = 0; k < 8; ++k) {
    int p = k;
    double mx = fabs(A[JIDX(k, k)]);
    for (int i = k + 1; i < 8; ++i) {
      const double v = fabs(A[JIDX(i, k)]);
      if (v > mx) { mx = v; p = i; }
    }
    if (mx == 0.0) return false;
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < 8; ++j) { const double t = A[JIDX(k, j)]; A[JIDX(k, j)] = A[JIDX(p, j)]; A[JIDX(p, j)] = t; }
    const double inv = 1.0 / A[JIDX(k, k)];
    for (int i = k + 1; i < 8; ++i) {
      const double f = A[JIDX(i, k)] * inv;
      A[JIDX(i, k)] = f;
      for (int j = k + 1; j < 8; ++j) A[JIDX(i, j)] -= f * A[JIDX(k, j)];
    }
  }
  return true;
}
__device__ __noinline__ void lu_solve8(const double* A, const int* piv, double* b) {
  // rows of L were swapped in full during factorisation: apply every interchange first
  for (int k = 0; k < 8; ++k) {
    const int p = piv[k];
    if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
  }
  for (int k = 0; k < 8; ++k)
    for (int i = k + 1; i < 8; ++i) b[i] -= A[JIDX(i, k)] * b[k];
  for (int i = 7; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < 8; ++j) s -= A[JIDX(i, j)] * b[j];
    b[i] = s / A[JIDX(i, i)];
  }
}

// ---- the 8-unknown update problem -----------------------------------------------------------
template <int NSLIP>
struct Problem {
  double dt, dt_ri, detVi, tK;
  double e_n[5], q_n[4], d_sm[5], w_sm[3];
  double eps_si, rot_si, T1_shift;
  static constexpr int NG = (NSLIP == 24) ? 24 : 1;  // per-system resistances only differ for HCP families
  double g[NG], c_e[NG], gam_w, gam_r;
  // state of the last evaluation
  double e_f[5], q_f[4], C[9];
  double shrate, disRate;
  double gdot[NSLIP];
  __device__ __forceinline__ double gv(int a) const { return g[NG == 1 ? 0 : a]; }
  __device__ __forceinline__ double cev(int a) const { return c_e[NG == 1 ? 0 : a]; }

  __device__ void kin_vals(const MatDev& m, double h) {
    if (m.kin == KIN_KMBALD) {
      const double sq = exp(0.5 * h);
      for (int a = 0; a < NG; ++a) { g[a] = m.go[a] + m.s_[a] * sq; c_e[a] = m.c_1[a] / tK * m.mu_ref; }
      gam_w = m.gam_wo / sq;
      gam_r = m.gam_ro * sq * sq;
    } else {
      for (int a = 0; a < NG; ++a) { g[a] = h; c_e[a] = 0.0; }
      gam_w = m.gam_w0;
      gam_r = 0.0;
    }
  }

  // residual R[8]; if Jac != nullptr also the 8x8 Jacobian (row-major).  Deliberately not inlined: one
  // copy of this body keeps the kernel inside the instruction cache.
  __device__ __noinline__ void eval(const MatDev& m, const double* x, double* R, double* Jac) {
    double edot[5], xi[3];
#pragma unroll
    for (int i = 0; i < 5; ++i) { const double de = e_scale * x[i]; e_f[i] = e_n[i] + de; edot[i] = de * dt_ri; }
#pragma unroll
    for (int k = 0; k < 3; ++k) xi[k] = r_scale * x[5 + k];
    // q_f = q_n * exp-map(xi)
    {
      double A[4];
      const double th = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
      if (th > idp_eps_sqrt) {
        const double s = sin(0.5 * th) / th;
        A[0] = cos(0.5 * th); A[1] = s * xi[0]; A[2] = s * xi[1]; A[3] = s * xi[2];
      } else {
        A[0] = 1.0; A[1] = 0.5 * xi[0]; A[2] = 0.5 * xi[1]; A[3] = 0.5 * xi[2];
        const double n = 1.0 / sqrt(A[0] * A[0] + A[1] * A[1] + A[2] * A[2] + A[3] * A[3]);
        A[0] *= n; A[1] *= n; A[2] *= n; A[3] *= n;
      }
      q_f[0] = q_n[0] * A[0] - q_n[1] * A[1] - q_n[2] * A[2] - q_n[3] * A[3];
      q_f[1] = q_n[0] * A[1] + q_n[1] * A[0] + q_n[2] * A[3] - q_n[3] * A[2];
      q_f[2] = q_n[0] * A[2] - q_n[1] * A[3] + q_n[2] * A[0] + q_n[3] * A[1];
      q_f[3] = q_n[0] * A[3] + q_n[1] * A[2] - q_n[2] * A[1] + q_n[3] * A[0];
    }
    quat_to_tensor(q_f, C);
    double d_lat[5], w_lat[3];
    rot_vecd<true>(C, d_sm, d_lat);
#pragma unroll
    for (int k = 0; k < 3; ++k) w_lat[k] = C[0 + k] * w_sm[0] + C[3 + k] * w_sm[1] + C[6 + k] * w_sm[2];
    double T[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) T[i] = m.Kdiag[i] * e_f[i];
    T[1] += T1_shift;
    double dp[5] = {0, 0, 0, 0, 0}, wp[3] = {0, 0, 0};
    double dDp[5][5], dWp[3][5];
    if (Jac) {
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) dDp[i][j] = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 5; ++j) dWp[k][j] = 0.0;
    }
    shrate = 0.0;
    disRate = 0.0;
    const double gi0 = 1.0 / g[0], xmi0 = (m.kin == KIN_KMBALD) ? 0.0 : 1.0 / m.xm;  // Voce: one resistance for all systems
#pragma unroll 1
    for (int a = 0; a < NSLIP; ++a) {
      double tau = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) tau += m.P[a][i] * T[i];
      double gd, dg;
      if (m.kin == KIN_KMBALD) kin_kmbald(m, gv(a), gam_w, gam_r, cev(a), tau, gd, dg);
      else kin_power_law(m, gam_w, gv(a), gi0, xmi0, tau, gd, dg);
      gdot[a] = gd;
      shrate += fabs(gd);
      disRate += tau * gd;
#pragma unroll
      for (int i = 0; i < 5; ++i) dp[i] += gd * m.P[a][i];
#pragma unroll
      for (int k = 0; k < 3; ++k) wp[k] += gd * m.Q[a][k];
      if (Jac) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const double dga = dg * m.P[a][j] * m.Kdiag[j];
#pragma unroll
          for (int i = 0; i < 5; ++i) dDp[i][j] += m.P[a][i] * dga;
#pragma unroll
          for (int k = 0; k < 3; ++k) dWp[k][j] += m.Q[a][k] * dga;
        }
      }
    }
    double Me[5][3], Medot[5][3];
    comm_Me(e_f, Me);
    comm_Me(edot, Medot);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double ewp = Me[i][0] * wp[0] + Me[i][1] * wp[1] + Me[i][2] * wp[2];
      R[i] = eps_si * (edot[i] + ewp + dp[i] - d_lat[i]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double xe_dp = 0.0, xedot_e = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) { xe_dp += 0.5 * Me[i][k] * dp[i]; xedot_e += 0.5 * Medot[i][k] * e_f[i]; }
      R[5 + k] = rot_si * dt * (xi[k] * dt_ri + wp[k] - w_lat[k] + (xe_dp - 0.5 * xedot_e));
    }
    if (!Jac) return;
    double Mwp[5][5], JrM[3][3], Mdl[5][3], Mdp[5][3];
    comm_Mw(wp, Mwp);
    exp_Jr(xi, JrM);
    comm_Me(d_lat, Mdl);
    comm_Me(dp, Mdp);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double v = (i == j ? dt_ri : 0.0) + Mwp[i][j] + dDp[i][j];
#pragma unroll
        for (int k = 0; k < 3; ++k) v += Me[i][k] * dWp[k][j];
        Jac[JIDX(i, j)] = eps_si * v * e_scale;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) v += Mdl[i][l] * JrM[l][k];
        Jac[JIDX(i, 5 + k)] = -eps_si * v * r_scale;
      }
    }
    const double Wl[3][3] = {{0.0, -w_lat[2], w_lat[1]}, {w_lat[2], 0.0, -w_lat[0]}, {-w_lat[1], w_lat[0], 0.0}};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        double t = -0.5 * Mdp[j][k];
#pragma unroll
        for (int i = 0; i < 5; ++i) t += 0.5 * Me[i][k] * dDp[i][j];
        t += -0.5 * (-0.5 * Me[j][k] * dt_ri + 0.5 * Medot[j][k]);
        Jac[JIDX(5 + k, j)] = rot_si * dt * (dWp[k][j] + t) * e_scale;
      }
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        double v = (k == l ? dt_ri : 0.0);
#pragma unroll
        for (int n = 0; n < 3; ++n) v -= Wl[k][n] * JrM[n][l];
        Jac[JIDX(5 + k, 5 + l)] = rot_si * dt * v * r_scale;
      }
    }
  }
};

__device__ __forceinline__ double norm8(const double* v) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// trust-region dogleg Newton; returns number of residual evaluations, negative on failure.
// On return R/J hold the residual and (unfactored) Jacobian at the returned x.  The Jacobian is
// factored in place: grad = J^T R and J grad are formed first, and the linear model of any dogleg
// step s = -a grad + b nr follows from J nr = -R as  R + J s = (1 - b) R - a (J grad).
template <int NSLIP>
__device__ __noinline__ int solve_trdl(const MatDev& m, Problem<NSLIP>& prob, double* x, double* R, double* J, double tol) {
  double Rt[8], xt[8];
  prob.eval(m, x, R, J);
  int nfev = 1;
  double res = norm8(R);
  double delta = 1.0e2;
  const double xiLG = 0.75, xiLO = 0.35, xiIncDelta = 1.5, xiDecDelta = 0.25;
  const double deltaMin = 1e-12, deltaMax = 1e4;
  for (int it = 0; it < 200; ++it) {
    if (res <= tol) return nfev;
    double grad[8], Jg[8], nr[8];
    int piv[8];
    for (int j = 0; j < 8; ++j) { double s = 0; for (int i = 0; i < 8; ++i) s += J[JIDX(i, j)] * R[i]; grad[j] = s; }
    for (int i = 0; i < 8; ++i) { double s = 0; for (int j = 0; j < 8; ++j) s += J[JIDX(i, j)] * grad[j]; Jg[i] = s; }
    for (int i = 0; i < 8; ++i) nr[i] = -R[i];
    const bool have_newton = lu_factor8(J, piv);
    if (have_newton) lu_solve8(J, piv, nr);
    double g2 = 0, Jg2 = 0;
    for (int i = 0; i < 8; ++i) { g2 += grad[i] * grad[i]; Jg2 += Jg[i] * Jg[i]; }
    const double nrn = have_newton ? norm8(nr) : 1e300;
    bool accepted = false;
    while (!accepted) {
      double ca, cb, pred;  // step = -ca * grad + cb * nr
      if (have_newton && nrn <= delta) {
        ca = 0.0; cb = 1.0;
        pred = res;
      } else {
        const double alpha = (Jg2 > 0) ? g2 / Jg2 : 0.0;
        const double cpn = alpha * sqrt(g2);
        if (cpn >= delta || !have_newton) {
          ca = delta / sqrt(g2 > 0 ? g2 : 1.0); cb = 0.0;
        } else {
          double a = 0, b = 0, c = -delta * delta;
          for (int i = 0; i < 8; ++i) {
            const double cp = -alpha * grad[i], d = nr[i] - cp;
            a += d * d; b += 2.0 * cp * d; c += cp * cp;
          }
          const double beta = (-b + sqrt(fmax(0.0, b * b - 4 * a * c))) / (2 * a);
          ca = alpha * (1.0 - beta); cb = beta;
        }
        double l2 = 0.0;
        for (int i = 0; i < 8; ++i) { const double l = (1.0 - cb) * R[i] - ca * Jg[i]; l2 += l * l; }
        pred = res - sqrt(l2);
      }
      double sn = 0.0;
      for (int i = 0; i < 8; ++i) { const double st = cb * nr[i] - ca * grad[i]; xt[i] = x[i] + st; sn += st * st; }
      sn = sqrt(sn);
      // Jacobian evaluated at the trial point straight into J (the factored old one is dead: a rejected
      // step only needs grad, J grad and nr), so an accepted step needs no second evaluation
      prob.eval(m, xt, Rt, J);
      ++nfev;
      const double rest = norm8(Rt);
      const bool finite = isfinite(rest);
      const double rho = (finite && pred > 0) ? (res - rest) / pred : -1.0;
      if (finite && rest < res) {
        accepted = true;
        for (int i = 0; i < 8; ++i) { x[i] = xt[i]; R[i] = Rt[i]; }
        if (rho > xiLG && sn >= 0.99 * delta) delta = fmin(deltaMax, delta * xiIncDelta);
        else if (rho < xiLO) delta = fmax(deltaMin, fmax(delta, sn) * xiDecDelta * 2.0);
        res = rest;
      } else {
        delta = fmin(delta, sn) * xiDecDelta;
        if (delta < deltaMin) { prob.eval(m, x, R, J); return -(nfev + 1); }  // restore state at x
      }
    }
  }
  if (res > tol) return -nfev;
  return nfev;
}

}  // namespace mat

// ------------------------------------------------------------------------------------------
// K1 kernel.  MODE as in k_operator.cuh (LVEC: vel is an L-vector gathered through e2n; EVEC:
// vel is the E-vector the reference's ModelSetup receives).  Layouts: QFs (vdim,q,e),
// jac (3,3,q,e), matgrad[(pt)*36 + j*6 + i] = d sigma_i / d eps_j (after the reference's transpose,
// src/mechanics_ecmech.cpp:159-169; TRANSPOSE=false reproduces the EA-on-device quirk, :155).
// fail_count is incremented for points whose local solve did not converge.
// ------------------------------------------------------------------------------------------
constexpr int kJS = mat::kJS;
constexpr int kK1SmemBytes = 64 * kJS * 8;  // 64 KB: one 8x8 Jacobian per thread
template <int NSLIP, int MODE, int MINB>
__global__ void __launch_bounds__(kJS, MINB) k_model_setup(const MatDev* __restrict__ mp, double dt, double temp_k, const double* __restrict__ jac,
                                                     const double* __restrict__ vel, const int* __restrict__ e2n,
                                                     long nnodes, const double* __restrict__ stress0,
                                                     const double* __restrict__ hist0, double* __restrict__ stress1,
                                                     double* __restrict__ hist1, double* __restrict__ matgrad,
                                                     long nelems, int transpose, int* __restrict__ fail_count) {
  using namespace mat;
  const MatDev& m = *mp;  // read-only, uniform across the grid (broadcast loads)
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  const int nsv = m.nhist;
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
  if (!active) return;
  const long p = e * 8 + lane;
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
  // ---- kernel_setup ----
  const double* h0 = hist0 + p * nsv;
  double* h1 = hist1 + p * nsv;
  const int ind_int_eng = nsv - 1, ind_vols = nsv - 2;
  double w_vec[3], d_svec_p[7], s_svec_p[7];
  w_vec[0] = 0.5 * (L[2][1] - L[1][2]);
  w_vec[1] = 0.5 * (L[0][2] - L[2][0]);
  w_vec[2] = 0.5 * (L[1][0] - L[0][1]);
  const double d_mean = -(1.0 / 3.0) * (L[0][0] + L[1][1] + L[2][2]);
  d_svec_p[0] = L[0][0] + d_mean;
  d_svec_p[1] = L[1][1] + d_mean;
  d_svec_p[2] = L[2][2] + d_mean;
  d_svec_p[3] = 0.5 * (L[2][1] + L[1][2]);
  d_svec_p[4] = 0.5 * (L[2][0] + L[0][2]);
  d_svec_p[5] = 0.5 * (L[1][0] + L[0][1]);
  d_svec_p[6] = -3.0 * d_mean;
  Problem<NSLIP> prob;
  svec_to_vecd(d_svec_p, prob.d_sm);
  double dEff;
  {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) n2 += prob.d_sm[i] * prob.d_sm[i];
    dEff = sqr2b3 * sqrt(n2);
  }
  const double vOld = h0[ind_vols];
  const double vNew = vOld * exp(d_svec_p[6] * dt);
  const double volInc = vNew - vOld;
  {
    double sig[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sig[i] = stress0[p * 6 + i];
    const double sm = -(1.0 / 3.0) * (sig[0] + sig[1] + sig[2]);
#pragma unroll
    for (int i = 0; i < 6; ++i) s_svec_p[i] = sig[i];
    s_svec_p[0] += sm; s_svec_p[1] += sm; s_svec_p[2] += sm;
    s_svec_p[6] = sm;
  }
  // ---- getResponseECM (evptn) ----
#pragma unroll
  for (int k = 0; k < 3; ++k) prob.w_sm[k] = w_vec[k];
#pragma unroll
  for (int i = 0; i < 5; ++i) prob.e_n[i] = h0[iH_E + i];
  {
    double q[4], n = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i] = h0[iH_Q + i]; n += q[i] * q[i]; }
    n = 1.0 / sqrt(n);
#pragma unroll
    for (int i = 0; i < 4; ++i) prob.q_n[i] = q[i] * n;
  }
  const double eOld = h0[ind_int_eng], pOld = s_svec_p[6];
  double tkelv = m.tK0 + eOld * m.dtde;
  (void)temp_k;
  const double eta = 1.0 / vNew - 1.0;
  double eNew = eOld - volInc * pOld;
  double pEOS = m.bulk * eta + m.gruneisen * eNew;
  eNew = eOld - 0.5 * volInc * (pOld + pEOS);
  pEOS = m.bulk * eta + m.gruneisen * eNew;
  const double dp_dlnV = -m.bulk / vNew;
  double shr_n = 0.0;
#pragma unroll 4
  for (int a = 0; a < NSLIP; ++a) shr_n += fabs(h0[iH_Gdot + a]);
  const double h_u = kin_update_h(m, h0[iH_H], dt, shr_n);
  prob.dt = dt;
  prob.dt_ri = 1.0 / dt;
  prob.detVi = 1.0 / vNew;
  prob.tK = tkelv;
  prob.T1_shift = m.Kvd * log(vNew) / sqr3;
  prob.kin_vals(m, h_u);
  const double halfVMidDt = 0.25 * (vOld + vNew) * dt;
  double dEDev = halfVMidDt * (s_svec_p[0] * d_svec_p[0] + s_svec_p[1] * d_svec_p[1] + s_svec_p[2] * d_svec_p[2] +
                               2.0 * (s_svec_p[3] * d_svec_p[3] + s_svec_p[4] * d_svec_p[4] + s_svec_p[5] * d_svec_p[5]));
  {
    const double eps_dot = fmax(dEff * sqr3b2, 1.0e-12 / dt);
    prob.eps_si = fmin(1.0 / eps_dot, 1.0e6 * dt);
    prob.rot_si = prob.dt_ri * prob.eps_si;
  }
  extern __shared__ double smJ[];
  double* J = smJ + threadIdx.x;  // J(i,j) at J[JIDX(i,j)]
  double x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, R[8];
  int nfev = solve_trdl<NSLIP>(m, prob, x, R, J, m.tol);
  if (nfev < 0) { atomicAdd(fail_count, 1); nfev = -nfev; }
  // ---- history out (StateVarsSetup copy + updates + kernel_postprocessing) ----
  h1[iH_shrateEff] = prob.shrate;
  h1[iH_shrEff] = h0[iH_shrEff] + prob.shrate * dt;
  {
    double flow = prob.gv(0);
    if (dEff > idp_tiny_sqrt) flow = prob.disRate / dEff;
    double plw = (dEff > idp_tiny_sqrt) ? flow * dEff * dt : 0.0;  // kernel_postprocessing :135-140
    h1[iH_flowStr] = plw + h0[iH_flowStr];
  }
  h1[iH_nFEval] = (double)nfev;
#pragma unroll
  for (int i = 0; i < 5; ++i) h1[iH_E + i] = prob.e_f[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) h1[iH_Q + i] = prob.q_f[i];
  h1[iH_H] = h_u;
#pragma unroll 4
  for (int a = 0; a < NSLIP; ++a) h1[iH_Gdot + a] = prob.gdot[a];
  h1[ind_vols] = vNew;
  // ---- stress out ----
  double sig_lat[5], sig_sm[5], s6[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) sig_lat[i] = prob.detVi * m.Kdiag[i] * prob.e_f[i];
  sig_lat[1] += prob.detVi * prob.T1_shift;
  const double p_tot = pEOS - m.Kvd * prob.e_f[1] * prob.detVi / sqr3;  // hexagonal: c-axis strain carries pressure
  rot_vecd<false>(prob.C, sig_lat, sig_sm);
  vecd_to_svec(sig_sm, s6);
  dEDev += halfVMidDt * (s6[0] * d_svec_p[0] + s6[1] * d_svec_p[1] + s6[2] * d_svec_p[2] +
                         2.0 * (s6[3] * d_svec_p[3] + s6[4] * d_svec_p[4] + s6[5] * d_svec_p[5]));
  h1[ind_int_eng] = eNew + dEDev;
  {
    double* so = stress1 + p * 6;
    so[0] = s6[0] - p_tot; so[1] = s6[1] - p_tot; so[2] = s6[2] - p_tot;
    so[3] = s6[3]; so[4] = s6[4]; so[5] = s6[5];
  }
  // ---- algorithmic tangent ----
  {
    int piv[8];
    const bool ok = lu_factor8(J, piv);  // J is consumed here
    double Msl[5][3], JrM[3][3];
    const double xi[3] = {r_scale * x[5], r_scale * x[6], r_scale * x[7]};
    comm_Me(sig_lat, Msl);
    exp_Jr(xi, JrM);
    double dsd[5][5], s1c[5];
    for (int c = 0; c < 5; ++c) {
      // rhs = eps_si * R5[c][:] = eps_si * (row c of R5) = eps_si * 5vec(C^T B_c C)
      double ec[5] = {0, 0, 0, 0, 0}, rhs[8];
      ec[c] = 1.0;
      rot_vecd<true>(prob.C, ec, rhs);
#pragma unroll
      for (int i = 0; i < 5; ++i) rhs[i] *= prob.eps_si;
      rhs[5] = rhs[6] = rhs[7] = 0.0;
      if (ok) lu_solve8(J, piv, rhs);
      else for (int i = 0; i < 8; ++i) rhs[i] = 0.0;
      s1c[c] = rhs[1];
      double dl[5], col[5];
      double jr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) jr[k] = r_scale * (JrM[k][0] * rhs[5] + JrM[k][1] * rhs[6] + JrM[k][2] * rhs[7]);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        dl[j] = prob.detVi * m.Kdiag[j] * e_scale * rhs[j] - (Msl[j][0] * jr[0] + Msl[j][1] * jr[1] + Msl[j][2] * jr[2]);
      rot_vecd<false>(prob.C, dl, col);
#pragma unroll
      for (int i = 0; i < 5; ++i) dsd[i][c] = col[i] / dt;
    }
    // 5x5 deviatoric operator -> 6x6 Voigt, engineering shear columns
    const double Tm[5][6] = {{sqr2i, -sqr2i, 0, 0, 0, 0},
                             {-sqr6i, -sqr6i, 2.0 * sqr6i, 0, 0, 0},
                             {0, 0, 0, 0, 0, sqr2},
                             {0, 0, 0, 0, sqr2, 0},
                             {0, 0, 0, sqr2, 0, 0}};
    const double Bm[6][5] = {{sqr2i, -sqr6i, 0, 0, 0}, {-sqr2i, -sqr6i, 0, 0, 0}, {0, sqr2b3, 0, 0, 0},
                             {0, 0, 0, 0, sqr2i},      {0, 0, 0, sqr2i, 0},       {0, 0, sqr2i, 0, 0}};
    double hexa[6] = {0, 0, 0, 0, 0, 0}, hexb[6] = {0, 0, 0, 0, 0, 0};
    if (m.Kvd != 0.0) {
      const double kc = m.Kvd * prob.detVi / sqr3;
      for (int j = 0; j < 6; ++j) {
        double v = 0.0;
        for (int c = 0; c < 5; ++c) v += kc * e_scale * s1c[c] / dt * Tm[c][j];
        hexa[j] = (j >= 3) ? 0.5 * v : v;
      }
      const double e1[5] = {0.0, kc, 0.0, 0.0, 0.0};
      double e1sm[5];
      rot_vecd<false>(prob.C, e1, e1sm);
      vecd_to_svec(e1sm, hexb);
    }
    double* K = matgrad + p * 36;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double v = 0.0;
        for (int a = 0; a < 5; ++a) {
          double t = 0.0;
          for (int b = 0; b < 5; ++b) t += dsd[a][b] * Tm[b][j];
          v += Bm[i][a] * t;
        }
        if (j >= 3) v *= 0.5;
        if (i < 3) v += hexa[j];
        if (j < 3) {
          v += -s6[i] + hexb[i];
          if (i < 3) v += -dp_dlnV;
        }
        K[transpose ? (j * 6 + i) : (i * 6 + j)] = v;
      }
  }
}

// init_state_vars (src/mechanics_ecmech.hpp:249-300): overwrite the ExaCMech-owned history slots
// with the model's initial values; quaternion slots (set from the grain file) are left alone.
__global__ void k_hist_init(MatDev m, double* __restrict__ hist, long npts) {
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
__global__ void k_calc_dp(MatDev m, const double* __restrict__ hist, double* __restrict__ dp9, long npts) {
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

// Per-quadrature-point crystal-plasticity update: the arithmetic of K1 (k_material.cuh), written as
// host/device inline functions so that the very same source is (a) inlined into the sm_100a kernel and
// (b) compiled for the host by the CPU test-suite (tests/hostcheck), which checks it point by point
// against the oracle without a GPU.  Nothing here is a CPU fallback: the library only ever launches the
// CUDA kernel.
//
// What one call of update_point replaces (src/mechanics_ecmech.cpp:192-258 per point):
//   kernel_setup                              src/mechanics_ecmech.cpp:22-100
//   ecmech::matModelBase::getResponseECM      src/mechanics_ecmech.cpp:183 (external ExaCMech evptn model)
//   kernel_postprocessing (+ transpose)       src/mechanics_ecmech.cpp:106-172
//
// The constitutive algorithm (evptn: 5 deviatoric lattice elastic strains + 3 exponential-map rotation
// increments solved by a trust-region dogleg Newton, backward Euler, hardness advanced with
// beginning-of-step slip rates, Kirchhoff-stress resolved shear stresses, constant-modulus EOS
// p = K(1/V - 1)) is the published ExaCMech formulation, implemented here independently for the GPU.
//
// Design for the SM (one thread = one point):
//   * the material description is a kernel parameter (constant bank): with the 12-system loops fully
//     unrolled every Schmid-tensor entry has a compile-time address and reaches the DFMAs through one uniform
//     load (LDCU into a uniform register; sm_100a DFMAs take no constant-bank operand) -- no per-thread loads;
//   * d D^p / d tau and d W^p / d tau are accumulated as (dg P)(x)P (15 unique entries) and Q(x)(dg P) (15)
//     instead of 40 products per system, from the same 8 table entries the D^p / W^p sums use;
//   * integer power-law exponents (1/m - 1 = 49 for the reference's Voce parameters) are evaluated by
//     repeated squaring for all systems at once instead of exp(n log x) per system;
//   * the 8x8 Newton system is factored in registers (row-wise Doolittle, 36 doubles of U live) straight
//     from the Jacobian in shared memory, which stays intact for the dogleg quantities (J^T R is accumulated
//     while the rows pass through registers); a growth check falls back to the partially pivoted in-place
//     factorisation;
//   * there is a single inlined call site of the residual/Jacobian evaluation (a small state machine
//     drives trial / re-evaluation / final passes), so the code stays inside the instruction cache.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EXAB_HD __host__ __device__ __forceinline__
#define EXAB_HDN __host__ __device__ __noinline__
#else
#define EXAB_HD inline
#define EXAB_HDN inline
#endif

// host-only path statistics for the test harness (tests/hostcheck): which solver branches ran
#if defined(EXAB_POINT_STATS) && !defined(__CUDA_ARCH__)
namespace exab { extern long g_point_stats[8]; }
#define EXAB_STAT(k) (__atomic_fetch_add(&::exab::g_point_stats[k], 1L, __ATOMIC_RELAXED))
#else
#define EXAB_STAT(k) ((void)0)
#endif

// unroll factor of the 12-system slip loops of the power-law path (tuning knob: full unrolling maximises ILP and
// uses immediate constant operands, partial unrolling shrinks the hot loop's instruction footprint)
#ifndef EXAB_SLIP_UNROLL
#define EXAB_SLIP_UNROLL 12
#endif
// CTA-lockstep local solve in the K1 kernel (default; -DEXAB_K1_NO_LOCKSTEP for the A/B): see solve_point
#if !defined(EXAB_K1_NO_LOCKSTEP) && !defined(EXAB_K1_LOCKSTEP)
#define EXAB_K1_LOCKSTEP 1
#endif
#define EXAB_PRAGMA_(x) _Pragma(#x)
#define EXAB_PRAGMA(x) EXAB_PRAGMA_(x)
#define EXAB_UNROLL_SLIP EXAB_PRAGMA(unroll EXAB_SLIP_UNROLL)

namespace exab {

constexpr int kMaxSlip = 24;
enum { KIN_VOCE = 0, KIN_VOCE_NL = 1, KIN_KMBALD = 2 };
enum { XTAL_FCC = 0, XTAL_BCC = 1, XTAL_HCP = 2 };

// history layout (src/mechanics_ecmech.hpp:136-141,165-185)
constexpr int iH_shrateEff = 0, iH_shrEff = 1, iH_flowStr = 2, iH_nFEval = 3, iH_E = 4, iH_Q = 9, iH_H = 13,
              iH_Gdot = 14;

struct MatDev {
  int xtal, kin, nslip, nhist, withGAthermal;
  int pl_n;         // > 0: the power-law exponent 1/m - 1 is this integer (repeated squaring)
  int force_pivot;  // testing: always take the pivoted-LU fallback
  int pad_;
  double P[kMaxSlip][5];
  double Q[kMaxSlip][3];
  double Kdiag[5], bulk, gmod, Kvd;  // Kvd: hexagonal volumetric <-> c-axis deviator coupling (0 for cubic)
  double tol, gruneisen, dtde, tK0;
  // Voce power law
  double xm, xmi, gam_w0, h0, tausi, taus0, xmprime, xms, gamss0, kappa0;
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
constexpr double kGrowthMax = 64.0;  // largest multiplier tolerated by the unpivoted factorisation

// Compact tangent record (32 doubles in the point's 36-double slot; cubic crystals): [0..24] the 5x5 deviatoric
// operator dsd (row-major, per unit strain increment), [25] kvol = -dp/dlnV, [26..31] the deviatoric Cauchy stress s'.
// It expands to the Voigt 6x6 (engineering-shear columns) as
//   K = Bm dsd Tm' + (-s'_i + [i<3] kvol) for the three normal-strain columns,
// Tm' (5x6): deviatoric 5-vector of a Voigt strain with halved shear entries, Bm (6x5): Voigt stress of a 5-vector.
constexpr int kTangentCompact = 2;
// S = K eps for a Voigt (engineering shear) strain-like vector eps, straight from the compact record
EXAB_HD void compact_apply(const double* __restrict__ rec, const double* eps, double* S) {
  const double v[5] = {sqr2i * (eps[0] - eps[1]), sqr6i * (2.0 * eps[2] - eps[0] - eps[1]), sqr2i * eps[5], sqr2i * eps[4],
                       sqr2i * eps[3]};
  double w[5];
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    double t = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) t += rec[a * 5 + b] * v[b];
    w[a] = t;
  }
  const double tr = eps[0] + eps[1] + eps[2];
  S[0] = sqr2i * w[0] - sqr6i * w[1] + (rec[25] - rec[26]) * tr;
  S[1] = -sqr2i * w[0] - sqr6i * w[1] + (rec[25] - rec[27]) * tr;
  S[2] = sqr2b3 * w[1] + (rec[25] - rec[28]) * tr;
  S[3] = sqr2i * w[4] - rec[29] * tr;
  S[4] = sqr2i * w[3] - rec[30] * tr;
  S[5] = sqr2i * w[2] - rec[31] * tr;
}
// the Voigt 6x6 at K36[j*6+i] = d sigma_i / d eps_j from the compact record
EXAB_HD void compact_expand(const double* __restrict__ rec, double* K36) {
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double e[6] = {0, 0, 0, 0, 0, 0};
    e[j] = 1.0;
    compact_apply(rec, e, &K36[j * 6]);
  }
}

// packed index of the symmetric 5x5 accumulator, i <= j
EXAB_HD constexpr int sidx(int i, int j) { return i * 5 - (i * (i - 1)) / 2 + (j - i); }
EXAB_HD constexpr int sidx_sym(int i, int j) { return i <= j ? sidx(i, j) : sidx(j, i); }

EXAB_HD void svec_to_vecd(const double* s, double* v) {
  v[0] = sqr2i * (s[0] - s[1]);
  v[1] = sqr6i * (2.0 * s[2] - s[0] - s[1]);
  v[2] = sqr2 * s[5];
  v[3] = sqr2 * s[4];
  v[4] = sqr2 * s[3];
}
EXAB_HD void vecd_to_svec(const double* v, double* s) {
  const double t1 = sqr2i * v[0], t2 = sqr6i * v[1];
  s[0] = t1 - t2;
  s[1] = -t1 - t2;
  s[2] = sqr2b3 * v[1];
  s[3] = sqr2i * v[4];
  s[4] = sqr2i * v[3];
  s[5] = sqr2i * v[2];
}
EXAB_HD void quat_to_tensor(const double* q, double* c) {
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
EXAB_HD void rot_vecd(const double* C, const double* v, double* out) {
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
  double b[6];
  const int ei[6] = {0, 1, 2, 1, 0, 0}, ej[6] = {0, 1, 2, 2, 2, 1};
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) a += A[3 * ei[n] + k] * (TRANSPOSE ? C[3 * k + ej[n]] : C[3 * ej[n] + k]);
    b[n] = a;
  }
  out[0] = sqr2i * (b[0] - b[1]);
  out[1] = sqr6i * (2.0 * b[2] - b[0] - b[1]);
  out[2] = sqr2 * b[5];
  out[3] = sqr2 * b[4];
  out[4] = sqr2 * b[3];
}
// Me(e): 5vec(E W - W E) = Me(e) w   (structure constants of the dev-sym / skew commutator)
EXAB_HD void comm_Me(const double* e, double M[5][3]) {
  M[0][0] = -e[4];              M[0][1] = -e[3];               M[0][2] = 2.0 * e[2];
  M[1][0] = -sqr3 * e[4];       M[1][1] = sqr3 * e[3];         M[1][2] = 0.0;
  M[2][0] = e[3];               M[2][1] = -e[4];               M[2][2] = -2.0 * e[0];
  M[3][0] = -e[2];              M[3][1] = e[0] - sqr3 * e[1];  M[3][2] = e[4];
  M[4][0] = e[0] + sqr3 * e[1]; M[4][1] = e[2];                M[4][2] = -e[3];
}
// Mw(w): Me(e) w = Mw(w) e
EXAB_HD void comm_Mw(const double* w, double M[5][5]) {
  M[0][0] = 0.0;         M[0][1] = 0.0;          M[0][2] = 2.0 * w[2]; M[0][3] = -w[1];       M[0][4] = -w[0];
  M[1][0] = 0.0;         M[1][1] = 0.0;          M[1][2] = 0.0;        M[1][3] = sqr3 * w[1]; M[1][4] = -sqr3 * w[0];
  M[2][0] = -2.0 * w[2]; M[2][1] = 0.0;          M[2][2] = 0.0;        M[2][3] = w[0];        M[2][4] = -w[1];
  M[3][0] = w[1];        M[3][1] = -sqr3 * w[1]; M[3][2] = -w[0];      M[3][3] = 0.0;         M[3][4] = w[2];
  M[4][0] = w[0];        M[4][1] = sqr3 * w[0];  M[4][2] = w[1];       M[4][3] = -w[2];       M[4][4] = 0.0;
}
// sine and cosine of half the rotation angle (one sincos serves the exponential map and its Jacobian)
EXAB_HD void half_angle(double th, double& sh, double& ch) {
#if defined(__CUDA_ARCH__)
  sincos(0.5 * th, &sh, &ch);
#else
  sh = sin(0.5 * th);
  ch = cos(0.5 * th);
#endif
}
// right Jacobian of the exponential map from the half-angle sine/cosine:
//   (1 - cos th)/th^2 = 2 sh^2/th^2,  (th - sin th)/th^3 = (th - 2 sh ch)/th^3
EXAB_HD void exp_Jr(const double* xi, double th2, double th, double sh, double ch, double Jm[3][3]) {
  double a, b;
  if (th < 1e-4) { a = 0.5 - th2 / 24.0; b = 1.0 / 6.0 - th2 / 120.0; }
  else { a = 2.0 * sh * sh / th2; b = (th - 2.0 * sh * ch) / (th2 * th); }
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
// Kocks-Mecking balanced thermally-activated + drag-limited kinetics (see oracle/ecmech_port.hpp)
EXAB_HD void kin_kmbald(const MatDev& m, double g, double gam_w, double gam_r, double c_e, double tau, double& gdot,
                        double& dgdot_dtau) {
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
EXAB_HD double kin_update_h(const MatDev& m, double h_n, double dt, double shr) {
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

// ---- 8x8 linear algebra; J(i,j) lives at J[(i*8+j)*JS] (JS = threads per CTA in shared memory) ----
#define EXAB_JIDX(i, j) (((i) * 8 + (j)) * JS)

// pivoted in-place factorisation / solve: the fallback of the register factorisation below
template <int JS>
EXAB_HDN bool lu_factor8(double* A, int* piv) {
  for (int k = 0; k < 8; ++k) {
    int p = k;
    double mx = fabs(A[EXAB_JIDX(k, k)]);
    for (int i = k + 1; i < 8; ++i) {
      const double v = fabs(A[EXAB_JIDX(i, k)]);
      if (v > mx) { mx = v; p = i; }
    }
    if (!(mx > 0.0)) return false;
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < 8; ++j) {
        const double t = A[EXAB_JIDX(k, j)];
        A[EXAB_JIDX(k, j)] = A[EXAB_JIDX(p, j)];
        A[EXAB_JIDX(p, j)] = t;
      }
    const double inv = 1.0 / A[EXAB_JIDX(k, k)];
    for (int i = k + 1; i < 8; ++i) {
      const double f = A[EXAB_JIDX(i, k)] * inv;
      A[EXAB_JIDX(i, k)] = f;
      for (int j = k + 1; j < 8; ++j) A[EXAB_JIDX(i, j)] -= f * A[EXAB_JIDX(k, j)];
    }
  }
  return true;
}
template <int JS>
EXAB_HDN void lu_solve8(const double* A, const int* piv, double* b) {
  // rows of L were swapped in full during factorisation: apply every interchange first
  for (int k = 0; k < 8; ++k) {
    const int p = piv[k];
    if (p != k) { const double t = b[k]; b[k] = b[p]; b[p] = t; }
  }
  for (int k = 0; k < 8; ++k)
    for (int i = k + 1; i < 8; ++i) b[i] -= A[EXAB_JIDX(i, k)] * b[k];
  for (int i = 7; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < 8; ++j) s -= A[EXAB_JIDX(i, j)] * b[j];
    b[i] = s / A[EXAB_JIDX(i, i)];
  }
}

// Row-wise Doolittle factorisation in registers of the (intact) matrix J, one right-hand side carried
// along; while the rows pass through registers the steepest-descent direction grad = J^T R is accumulated too.
// Returns false when a multiplier exceeds kGrowthMax or anything is not finite (-> pivoted path).
template <int JS>
EXAB_HD bool lu_solve_reg(const double* J, double* b, const double* R, double* grad) {
  double U[8][8], inv[8];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 8; ++j) grad[j] = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = J[EXAB_JIDX(i, j)];
#pragma unroll
    for (int j = 0; j < 8; ++j) grad[j] += a[j] * R[i];
    double bi = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double f = a[k] * inv[k];
      bad |= !(fabs(f) <= kGrowthMax);
#pragma unroll
      for (int j = k + 1; j < 8; ++j) a[j] -= f * U[k][j];
      bi -= f * b[k];
    }
#pragma unroll
    for (int j = i; j < 8; ++j) U[i][j] = a[j];
    inv[i] = 1.0 / a[i];
    b[i] = bi;
  }
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int j = i + 1; j < 8; ++j) s -= U[i][j] * b[j];
    b[i] = s * inv[i];
    bad |= !(fabs(b[i]) <= 1.0e300);
  }
  return !bad;
}
// Same factorisation written back in place (L multipliers below the diagonal, U above, 1/u_ii on it) -- but only
// after it has succeeded: on a growth failure J is left untouched for the pivoted fallback.
template <int JS>
EXAB_HD bool lu_factor_reg_store(double* J) {
  double U[8][8], L[8][8], inv[8];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = J[EXAB_JIDX(i, j)];
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double f = a[k] * inv[k];
      bad |= !(fabs(f) <= kGrowthMax);
      L[i][k] = f;
#pragma unroll
      for (int j = k + 1; j < 8; ++j) a[j] -= f * U[k][j];
    }
#pragma unroll
    for (int j = i; j < 8; ++j) U[i][j] = a[j];
    inv[i] = 1.0 / a[i];
    bad |= !(fabs(inv[i]) <= 1.0e300);
  }
  if (bad) return false;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int k = 0; k < i; ++k) J[EXAB_JIDX(i, k)] = L[i][k];
    J[EXAB_JIDX(i, i)] = inv[i];
#pragma unroll
    for (int j = i + 1; j < 8; ++j) J[EXAB_JIDX(i, j)] = U[i][j];
  }
  return true;
}
// NR right-hand sides r[c][0..7] solved at once from the stored factors (each factor entry is read once)
template <int JS, int NR>
EXAB_HD void lu_solve_stored(const double* J, double r[NR][8]) {
#pragma unroll
  for (int i = 1; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double l = J[EXAB_JIDX(i, k)];
#pragma unroll
      for (int c = 0; c < NR; ++c) r[c][i] -= l * r[c][k];
    }
#pragma unroll
  for (int i = 7; i >= 0; --i) {
#pragma unroll
    for (int j = i + 1; j < 8; ++j) {
      const double u = J[EXAB_JIDX(i, j)];
#pragma unroll
      for (int c = 0; c < NR; ++c) r[c][i] -= u * r[c][j];
    }
    const double d = J[EXAB_JIDX(i, i)];
#pragma unroll
    for (int c = 0; c < NR; ++c) r[c][i] *= d;
  }
}

EXAB_HD double norm8(const double* v) {
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// ---- the 8-unknown update problem -----------------------------------------------------------
// KIN: 0 = power law (Voce / Voce-NL, one slip resistance), 1 = KMBalD
template <int NSLIP, int KIN, int JS>
struct Point {
  double dt, dt_ri, detVi, tK;
  double e_n[5], q_n[4], d_sm[5], w_sm[3];
  double eps_si, rot_si, T1_shift;
  static constexpr int NG = (NSLIP == 24) ? 24 : 1;  // per-system resistances only differ for HCP families
  double g[NG], c_e[NG], gam_w, gam_r;
  // state of the last evaluation
  double e_f[5], q_f[4], C[9];
  double shrate, disRate;
  EXAB_HD double gv(int a) const { return g[NG == 1 ? 0 : a]; }
  EXAB_HD double cev(int a) const { return c_e[NG == 1 ? 0 : a]; }

  EXAB_HD void kin_vals(const MatDev& m, double h) {
    if (KIN == 1) {
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

  // One slip system's contribution: D^p += gd P, W^p += gd Q and the tangent sums S += dg P (x) P (15 unique entries),
  // Wq += dg Q (x) P, all from the same 8 table entries (on sm_100a every constant operand of a DFMA costs an LDCU
  // into a uniform register first, so the entries are loaded once and v = dg P is formed explicitly)
  EXAB_HD static void accumulate_slip(const MatDev& m, int a, double gd, double dg, double* dp, double* wp, double* S,
                                      double* Wq) {
    double Pa[5], Qa[3], v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { Pa[i] = m.P[a][i]; v[i] = dg * Pa[i]; dp[i] += gd * Pa[i]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { Qa[k] = m.Q[a][k]; wp[k] += gd * Qa[k]; }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = i; j < 5; ++j) S[sidx(i, j)] += v[i] * Pa[j];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 5; ++j) Wq[k * 5 + j] += Qa[k] * v[j];
  }

  // Residual R[8] at the scaled unknowns x; with want_jac also the 8x8 Jacobian into J; with gout != nullptr
  // the slip rates are written there.  e_f, q_f, C, shrate, disRate describe the evaluated state.
  EXAB_HD void eval(const MatDev& m, const double* x, double* R, double* J, bool want_jac, double* gout) {
    double edot[5], xi[3];
#pragma unroll
    for (int i = 0; i < 5; ++i) { const double de = e_scale * x[i]; e_f[i] = e_n[i] + de; edot[i] = de * dt_ri; }
#pragma unroll
    for (int k = 0; k < 3; ++k) xi[k] = r_scale * x[5 + k];
    const double th2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2], th = sqrt(th2);
    double sh = 0.0, ch = 1.0;
    // q_f = q_n * exp-map(xi)
    {
      double A[4];
      if (th > idp_eps_sqrt) {
        half_angle(th, sh, ch);
        const double s = sh / th;
        A[0] = ch; A[1] = s * xi[0]; A[2] = s * xi[1]; A[3] = s * xi[2];
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

    // ---- slip systems: D^p, W^p and (want_jac) S = sum dg P(x)P, Wq = sum dg Q(x)P ----
    double dp[5] = {0, 0, 0, 0, 0}, wp[3] = {0, 0, 0};
    double S[15], Wq[15];
#pragma unroll
    for (int n = 0; n < 15; ++n) { S[n] = 0.0; Wq[n] = 0.0; }
    shrate = 0.0;
    if (KIN == 0) {
      // power law, one resistance: tau/g for all systems, then |tau/g|^(1/m - 1) for all systems at once
      const double gi = 1.0 / g[0], xmi = m.xmi;
      double tt[NSLIP], pl[NSLIP];
EXAB_UNROLL_SLIP
      for (int a = 0; a < NSLIP; ++a) {
        double tau = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) tau += m.P[a][i] * T[i];
        tt[a] = tau * gi;
      }
      if (m.pl_n > 0) {
        double b[NSLIP];
#pragma unroll
        for (int a = 0; a < NSLIP; ++a) { b[a] = fabs(tt[a]); pl[a] = 1.0; }
        for (int n = m.pl_n; n; n >>= 1) {
          if (n & 1) {
#pragma unroll
            for (int a = 0; a < NSLIP; ++a) pl[a] *= b[a];
          }
#pragma unroll
          for (int a = 0; a < NSLIP; ++a) b[a] *= b[a];
        }
      } else {
#pragma unroll 2
        for (int a = 0; a < NSLIP; ++a) {
          const double at = fabs(tt[a]);
          pl[a] = (at > m.pl_t_min) ? exp((xmi - 1.0) * log(at)) : 0.0;
        }
      }
      const double gw_xmi_gi = gam_w * xmi * gi;
EXAB_UNROLL_SLIP
      for (int a = 0; a < NSLIP; ++a) {
        const double t = tt[a], at = fabs(t);
        double gd, dg;
        if (at <= m.pl_t_min) { gd = 0.0; dg = 0.0; }
        else if (at > m.pl_t_max) {  // linear extrapolation beyond the overflow guard
          const double d = gam_w * m.pl_max * xmi * gi;
          const double g0 = gam_w * m.pl_t_max * m.pl_max;
          gd = (g0 + d * g[0] * (at - m.pl_t_max)) * (t > 0 ? 1.0 : -1.0);
          dg = d;
        } else {
          gd = gam_w * t * pl[a];
          dg = gw_xmi_gi * pl[a];
        }
        if (gout) gout[a] = gd;
        shrate += fabs(gd);
        accumulate_slip(m, a, gd, dg, dp, wp, S, Wq);
      }
    } else {
#pragma unroll 1
      for (int a = 0; a < NSLIP; ++a) {
        double tau = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) tau += m.P[a][i] * T[i];
        double gd, dg;
        kin_kmbald(m, gv(a), gam_w, gam_r, cev(a), tau, gd, dg);
        if (gout) gout[a] = gd;
        shrate += fabs(gd);
        accumulate_slip(m, a, gd, dg, dp, wp, S, Wq);
      }
    }
    // plastic dissipation rate sum_a tau_a gdot_a = T . D^p
    disRate = T[0] * dp[0] + T[1] * dp[1] + T[2] * dp[2] + T[3] * dp[3] + T[4] * dp[4];

    // Strain-rate and spin equations.  The plastic velocity gradient acts in the unstretched lattice (no [e, W^p] /
    // [e, D^p] terms: what the reference's goldens pin, see oracle/ecmech_port.hpp Options::slip_stretch_terms); the
    // spin equation keeps the skew part of edot e.
    double Me[5][3], Medot[5][3];
    comm_Me(e_f, Me);
    comm_Me(edot, Medot);
#pragma unroll
    for (int i = 0; i < 5; ++i) R[i] = eps_si * (edot[i] + dp[i] - d_lat[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double xedot_e = 0.0;
#pragma unroll
      for (int i = 0; i < 5; ++i) xedot_e += 0.5 * Medot[i][k] * e_f[i];
      R[5 + k] = rot_si * dt * (xi[k] * dt_ri + wp[k] - w_lat[k] - 0.5 * xedot_e);
    }
    if (!want_jac) return;
    // dDp(i,j) = S(i,j) K_j ; dWp(k,j) = Wq(k,j) K_j
    double JrM[3][3], Mdl[5][3];
    exp_Jr(xi, th2, th, sh, ch, JrM);
    comm_Me(d_lat, Mdl);
    const double ce = eps_si * e_scale, cr = eps_si * r_scale;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const double v = S[sidx_sym(i, j)] * m.Kdiag[j] + (i == j ? dt_ri : 0.0);
        J[EXAB_JIDX(i, j)] = ce * v;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int l = 0; l < 3; ++l) v += Mdl[i][l] * JrM[l][k];
        J[EXAB_JIDX(i, 5 + k)] = -cr * v;
      }
    }
    const double Wl[3][3] = {{0.0, -w_lat[2], w_lat[1]}, {w_lat[2], 0.0, -w_lat[0]}, {-w_lat[1], w_lat[0], 0.0}};
    const double rde = rot_si * dt * e_scale, rdr = rot_si * dt * r_scale;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const double t = Wq[k * 5 + j] * m.Kdiag[j] - 0.5 * (-0.5 * Me[j][k] * dt_ri + 0.5 * Medot[j][k]);
        J[EXAB_JIDX(5 + k, j)] = rde * t;
      }
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        double v = (k == l ? dt_ri : 0.0);
#pragma unroll
        for (int n = 0; n < 3; ++n) v -= Wl[k][n] * JrM[n][l];
        J[EXAB_JIDX(5 + k, 5 + l)] = rdr * v;
      }
    }
  }
};

// Trust-region dogleg Newton with the reference solver's acceptance rules (oracle/ecmech_port.hpp solve_trdl),
// organised around ONE call site of Point::eval.  Invariants: (x, R, res) is the last accepted point.  Every trial
// point is evaluated with its Jacobian, so an accepted step needs nothing else; what a rejected step needs of the
// old point (grad = J^T R, J grad, the Newton step) is formed at the start of each Newton iteration while the
// Jacobian is still intact (the factorisation works in registers).  The last pass re-evaluates the accepted point
// to emit the slip rates; only a failed solve needs one extra Jacobian pass before it.  Returns the number of
// trial evaluations, negative on failure; J holds the (unfactored) Jacobian at the returned x.
template <int NSLIP, int KIN, int JS>
EXAB_HD int solve_point(const MatDev& m, Point<NSLIP, KIN, JS>& P, double* x, double* J, double tol, double* gout) {
  const double xiLG = 0.75, xiLO = 0.35, xiIncDelta = 1.5, xiDecDelta = 0.25;
  const double deltaMin = 1e-12, deltaMax = 1e4;
  enum { TRIAL = 0, REJAC = 2, FINAL = 3 };
  double R[8], xt[8], Rt[8], nr[8], grad[8], Jg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = 0.0; xt[i] = 0.0; R[i] = 0.0; nr[i] = 0.0; grad[i] = 0.0; Jg[i] = 0.0; }
  double res = 0.0, delta = 1.0e2, pred = 0.0, sn = 0.0, g2 = 0.0, Jg2 = 0.0, nrn = 1e300;
  int nfev = 0, iters = 0, what = TRIAL;
  bool first = true, failed = false, jac_valid = false, inner = false, have_newton = false;
#if defined(__CUDA_ARCH__) && defined(EXAB_K1_LOCKSTEP)
  // CTA lockstep: every pass of the loop body starts at a CTA-wide barrier, so the four warps of a CTA stream the body
  // (most of the kernel's 147 KB, far beyond the 32 KB L1.5 instruction cache) together instead of each warp missing
  // on its own; threads whose solve is finished keep arriving at the barrier until the whole CTA is done (inactive
  // threads of the last CTA do the same in the kernel).  The CTA's residency was already set by its slowest warp, so
  // the alignment costs nothing: measured on B200 at 128^3 in the elastic-plastic transition 62.1 -> 53.0 ms per call.
  bool done = false;
#endif
  for (;;) {
#if defined(__CUDA_ARCH__) && defined(EXAB_K1_LOCKSTEP)
    if (!__syncthreads_or(!done)) break;
    if (done) continue;
#endif
    P.eval(m, xt, Rt, J, what != FINAL, what == FINAL ? gout : nullptr);
#if defined(__CUDA_ARCH__) && defined(EXAB_K1_LOCKSTEP)
    if (what == FINAL) { done = true; continue; }
#else
    if (what == FINAL) break;
#endif
    if (what == REJAC) {
      jac_valid = true;
      EXAB_STAT(2);
    } else {
      EXAB_STAT(0);
      ++nfev;
      const double rest = norm8(Rt);
      if (first) {
        first = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) R[i] = Rt[i];
        res = rest;
        jac_valid = true;
        if (!isfinite(rest)) failed = true;
      } else {
        const bool finite = isfinite(rest);
        const double rho = (finite && pred > 0) ? (res - rest) / pred : -1.0;
        if (finite && rest < res) {
#pragma unroll
          for (int i = 0; i < 8; ++i) { x[i] = xt[i]; R[i] = Rt[i]; }
          if (rho > xiLG && sn >= 0.99 * delta) delta = fmin(deltaMax, delta * xiIncDelta);
          else if (rho < xiLO) delta = fmax(deltaMin, fmax(delta, sn) * xiDecDelta * 2.0);
          res = rest;
          inner = false;
          jac_valid = true;
        } else {
          delta = fmin(delta, sn) * xiDecDelta;
          EXAB_STAT(3);
          if (delta < deltaMin) failed = true;
          jac_valid = false;  // J now holds the rejected trial point's Jacobian
        }
      }
    }
    // ---- next action ----
    if (!inner && !failed) {  // start of a Newton iteration (J is the Jacobian at x)
      if (res <= tol || iters >= 200) {
        if (res > tol) failed = true;
      } else {
        ++iters;
#pragma unroll
        for (int i = 0; i < 8; ++i) nr[i] = -R[i];
        have_newton = !m.force_pivot && lu_solve_reg<JS>(J, nr, R, grad);
        bool pivoted = false;
        if (!have_newton) {
          EXAB_STAT(4);
          pivoted = true;
          for (int j = 0; j < 8; ++j) { double s = 0; for (int i = 0; i < 8; ++i) s += J[EXAB_JIDX(i, j)] * R[i]; grad[j] = s; }
        }
        // J grad for the Cauchy point and the linear model of a dogleg step
        g2 = 0.0; Jg2 = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) s += J[EXAB_JIDX(i, j)] * grad[j];
          Jg[i] = s;
          Jg2 += s * s;
          g2 += grad[i] * grad[i];
        }
        if (pivoted) {  // partially pivoted in-place factorisation (destroys J; the next trial rebuilds it)
          for (int i = 0; i < 8; ++i) nr[i] = -R[i];
          int piv[8];
          have_newton = lu_factor8<JS>(J, piv);
          if (have_newton) lu_solve8<JS>(J, piv, nr);
          jac_valid = false;
        }
        nrn = have_newton ? norm8(nr) : 1e300;
        if (!(nrn <= 1e300)) { have_newton = false; nrn = 1e300; }
        inner = true;
        if (have_newton && nrn <= delta) {
#pragma unroll
          for (int i = 0; i < 8; ++i) xt[i] = x[i] + nr[i];
          sn = nrn;
          pred = res;
          what = TRIAL;
          continue;
        }
      }
    }
    if (failed || !inner) {  // converged, or gave up: emit the state of the accepted point
      what = jac_valid ? FINAL : REJAC;
#pragma unroll
      for (int i = 0; i < 8; ++i) xt[i] = x[i];
      continue;
    }
    // ---- dogleg / Cauchy step inside the trust region (Newton step too long, or a rejected trial) ----
    {
      EXAB_STAT(5);
      double ca, cb;  // step = -ca * grad + cb * nr
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
        // linear model: R + J s = (1 - cb) R - ca (J grad), because J nr = -R
        double l2 = 0.0;
        for (int i = 0; i < 8; ++i) { const double l = (1.0 - cb) * R[i] - ca * Jg[i]; l2 += l * l; }
        pred = res - sqrt(l2);
      }
      double s2 = 0.0;
      for (int i = 0; i < 8; ++i) { const double st = cb * nr[i] - ca * grad[i]; xt[i] = x[i] + st; s2 += st * st; }
      sn = sqrt(s2);
      what = TRIAL;
    }
  }
  return failed ? -nfev : nfev;
}

// ------------------------------------------------------------------------------------------
// One material point.  L(i,t) = d v_i / d x_t; h0/s0: beginning-of-step history (m.nhist) and Cauchy stress
// (Voigt 11,22,33,23,13,12); h1/s1: end-of-step; K: the tangent d sigma_i / d eps_j, `layout` 1: 36 entries at
// [j*6+i] (the layout after the reference's transpose, src/mechanics_ecmech.cpp:159-169), 0: at [i*6+j], 2: the
// compact record of kTangentCompact (cubic crystals only); J: this point's 8x8 scratch, J(i,j) at J[(i*8+j)*JS].
// Returns the number of trial evaluations of the local solve, negative when it failed.
// ------------------------------------------------------------------------------------------
template <int NSLIP, int KIN, int JS>
EXAB_HD int update_point(const MatDev& m, double dt, const double L[3][3], const double* __restrict__ h0,
                         const double* __restrict__ s0, double* __restrict__ h1, double* __restrict__ s1,
                         double* __restrict__ K, int layout, double* J) {
  const int nsv = NSLIP + iH_Gdot + 2;
  const int ind_int_eng = nsv - 1, ind_vols = nsv - 2;
  // ---- kernel_setup ----
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
  Point<NSLIP, KIN, JS> prob;
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
    for (int i = 0; i < 6; ++i) sig[i] = s0[i];
    const double sm = -(1.0 / 3.0) * (sig[0] + sig[1] + sig[2]);
#pragma unroll
    for (int i = 0; i < 6; ++i) s_svec_p[i] = sig[i];
    s_svec_p[0] += sm; s_svec_p[1] += sm; s_svec_p[2] += sm;
    s_svec_p[6] = sm;
  }
  // ---- getResponseECM (evptn) ----
#pragma unroll
  for (int k = 0; k < 3; ++k) prob.w_sm[k] = w_vec[k];
  {
    // the beginning-of-step lattice strain is carried along with the step's volume change, e_u = e_n (V_old/V_new)^(1/3)
    // (oracle/ecmech_port.hpp Options::vol_convect: what the reference's cyclic goldens pin)
    const double f = cbrt(vOld / vNew);
#pragma unroll
    for (int i = 0; i < 5; ++i) prob.e_n[i] = f * h0[iH_E + i];
  }
  {
    double q[4], n = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { q[i] = h0[iH_Q + i]; n += q[i] * q[i]; }
    n = 1.0 / sqrt(n);
#pragma unroll
    for (int i = 0; i < 4; ++i) prob.q_n[i] = q[i] * n;
  }
  const double eOld = h0[ind_int_eng], pOld = s_svec_p[6];
  const double tkelv = m.tK0 + eOld * m.dtde;
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
  prob.T1_shift = (m.Kvd != 0.0) ? m.Kvd * log(vNew) / sqr3 : 0.0;
  prob.kin_vals(m, h_u);
  const double halfVMidDt = 0.25 * (vOld + vNew) * dt;
  double dEDev = halfVMidDt * (s_svec_p[0] * d_svec_p[0] + s_svec_p[1] * d_svec_p[1] + s_svec_p[2] * d_svec_p[2] +
                               2.0 * (s_svec_p[3] * d_svec_p[3] + s_svec_p[4] * d_svec_p[4] + s_svec_p[5] * d_svec_p[5]));
  {
    const double eps_dot = fmax(dEff * sqr3b2, 1.0e-12 / dt);
    prob.eps_si = fmin(1.0 / eps_dot, 1.0e6 * dt);
    prob.rot_si = prob.dt_ri * prob.eps_si;
  }
  double x[8];
  const int nfev_s = solve_point<NSLIP, KIN, JS>(m, prob, x, J, m.tol, h1 + iH_Gdot);
  const int nfev = nfev_s < 0 ? -nfev_s : nfev_s;
  // ---- history out (StateVarsSetup copy + updates + kernel_postprocessing) ----
  h1[iH_shrateEff] = prob.shrate;
  h1[iH_shrEff] = h0[iH_shrEff] + prob.shrate * dt;
  {
    double flow = prob.gv(0);
    if (dEff > idp_tiny_sqrt) flow = prob.disRate * prob.detVi / dEff;  // Cauchy stress : D^p (per current volume)
    const double plw = (dEff > idp_tiny_sqrt) ? flow * dEff * dt : 0.0;  // kernel_postprocessing :135-140
    h1[iH_flowStr] = plw + h0[iH_flowStr];
  }
  h1[iH_nFEval] = (double)nfev;
#pragma unroll
  for (int i = 0; i < 5; ++i) h1[iH_E + i] = prob.e_f[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) h1[iH_Q + i] = prob.q_f[i];
  h1[iH_H] = h_u;
  h1[ind_vols] = vNew;
  // ---- stress out ----
  // 5x5 rotation of deviatoric 5-vectors for the end-of-step orientation: lattice -> sample is R5 v, sample -> lattice
  // is R5^T v.  Built once (one copy of the tensor rotation in the code) and reused for the stress and the tangent.
  double R5[5][5];
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    double ej[5] = {0, 0, 0, 0, 0}, colj[5];
    ej[j] = 1.0;
    rot_vecd<false>(prob.C, ej, colj);
#pragma unroll
    for (int i = 0; i < 5; ++i) R5[i][j] = colj[i];
  }
  double sig_lat[5], sig_sm[5], s6[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) sig_lat[i] = prob.detVi * m.Kdiag[i] * prob.e_f[i];
  sig_lat[1] += prob.detVi * prob.T1_shift;
  const double p_tot = pEOS - m.Kvd * prob.e_f[1] * prob.detVi / sqr3;  // hexagonal: c-axis strain carries pressure
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) t += R5[i][j] * sig_lat[j];
    sig_sm[i] = t;
  }
  vecd_to_svec(sig_sm, s6);
  dEDev += halfVMidDt * (s6[0] * d_svec_p[0] + s6[1] * d_svec_p[1] + s6[2] * d_svec_p[2] +
                         2.0 * (s6[3] * d_svec_p[3] + s6[4] * d_svec_p[4] + s6[5] * d_svec_p[5]));
  h1[ind_int_eng] = eNew + dEDev;
  s1[0] = s6[0] - p_tot; s1[1] = s6[1] - p_tot; s1[2] = s6[2] - p_tot;
  s1[3] = s6[3]; s1[4] = s6[4]; s1[5] = s6[5];
  // ---- algorithmic tangent by implicit differentiation through the converged Jacobian ----
  {
    int piv[8];
    bool ok = !m.force_pivot && lu_factor_reg_store<JS>(J);
    const bool pivoted = !ok;
    if (pivoted) {
      EXAB_STAT(6);
      ok = lu_factor8<JS>(J, piv);  // J is intact: the register factorisation only writes on success
    }
    // right-hand sides: column c = eps_si * (row c of R5) on the strain rows, 0 on the rotation rows
    double r[5][8];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
#pragma unroll
      for (int i = 0; i < 5; ++i) r[c][i] = prob.eps_si * R5[c][i];
      r[c][5] = r[c][6] = r[c][7] = 0.0;
    }
    if (!pivoted) {
      lu_solve_stored<JS, 5>(J, r);
    } else {
      for (int c = 0; c < 5; ++c) {
        if (ok) lu_solve8<JS>(J, piv, r[c]);
        else for (int i = 0; i < 8; ++i) r[c][i] = 0.0;
      }
    }
    double Msl[5][3], JrM[3][3];
    const double xi[3] = {r_scale * x[5], r_scale * x[6], r_scale * x[7]};
    {
      const double th2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2], th = sqrt(th2);
      double sh = 0.0, ch = 1.0;
      if (th >= 1e-4) half_angle(th, sh, ch);
      exp_Jr(xi, th2, th, sh, ch, JrM);
    }
    comm_Me(sig_lat, Msl);
    double dsd[5][5], s1c[5];
    const double idt = 1.0 / dt;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      s1c[c] = r[c][1];
      double dl[5], col[5], jr[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) jr[k] = r_scale * (JrM[k][0] * r[c][5] + JrM[k][1] * r[c][6] + JrM[k][2] * r[c][7]);
#pragma unroll
      for (int j = 0; j < 5; ++j)
        dl[j] = prob.detVi * m.Kdiag[j] * e_scale * r[c][j] - (Msl[j][0] * jr[0] + Msl[j][1] * jr[1] + Msl[j][2] * jr[2]);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j) t += R5[i][j] * dl[j];
        col[i] = t;
      }
#pragma unroll
      for (int i = 0; i < 5; ++i) dsd[i][c] = col[i] * idt;
    }
    if (layout == kTangentCompact) {
      // private format of the fused PA path: the 6x6 is K = Bm dsd Tm' + (-s' + [i<3] kvol) (x) (1,1,1,0,0,0)
#pragma unroll
      for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) K[a * 5 + b] = dsd[a][b];
      K[25] = -dp_dlnV;
#pragma unroll
      for (int i = 0; i < 6; ++i) K[26 + i] = s6[i];
      return nfev_s;
    }
    // 5x5 deviatoric operator -> 6x6 Voigt (engineering-shear columns): K6 = Bm dsd Tm with the sparse maps
    //   Tm (5x6): vecd = Tm eps6 ; Bm (6x5): svec = Bm vecd ; shear columns halved
    double Mt[5][6];
#pragma unroll
    for (int a = 0; a < 5; ++a) {
      Mt[a][0] = sqr2i * dsd[a][0] - sqr6i * dsd[a][1];
      Mt[a][1] = -sqr2i * dsd[a][0] - sqr6i * dsd[a][1];
      Mt[a][2] = 2.0 * sqr6i * dsd[a][1];
      Mt[a][3] = 0.5 * sqr2 * dsd[a][4];
      Mt[a][4] = 0.5 * sqr2 * dsd[a][3];
      Mt[a][5] = 0.5 * sqr2 * dsd[a][2];
    }
    double hexa[6] = {0, 0, 0, 0, 0, 0}, hexb[6] = {0, 0, 0, 0, 0, 0};
    if (m.Kvd != 0.0) {
      const double kc = m.Kvd * prob.detVi / sqr3;
      const double f = kc * e_scale * idt;
      hexa[0] = f * (sqr2i * s1c[0] - sqr6i * s1c[1]);
      hexa[1] = f * (-sqr2i * s1c[0] - sqr6i * s1c[1]);
      hexa[2] = f * (2.0 * sqr6i * s1c[1]);
      hexa[3] = 0.5 * f * sqr2 * s1c[4];
      hexa[4] = 0.5 * f * sqr2 * s1c[3];
      hexa[5] = 0.5 * f * sqr2 * s1c[2];
      double e1sm[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) e1sm[i] = kc * R5[i][1];
      vecd_to_svec(e1sm, hexb);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double col[6];
      col[0] = sqr2i * Mt[0][j] - sqr6i * Mt[1][j];
      col[1] = -sqr2i * Mt[0][j] - sqr6i * Mt[1][j];
      col[2] = sqr2b3 * Mt[1][j];
      col[3] = sqr2i * Mt[4][j];
      col[4] = sqr2i * Mt[3][j];
      col[5] = sqr2i * Mt[2][j];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = col[i];
        if (i < 3) v += hexa[j];
        if (j < 3) {
          v += -s6[i] + hexb[i];
          if (i < 3) v += -dp_dlnV;
        }
        K[layout ? (j * 6 + i) : (i * 6 + j)] = v;
      }
    }
  }
  return nfev_s;
}

}  // namespace mat
}  // namespace exab

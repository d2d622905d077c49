// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of the crystal-plasticity update ExaConstit obtains from the
// third-party library ExaCMech through
//   mat_model_base->getResponseECM(...)            (src/mechanics_ecmech.cpp:183)
// ExaCMech (LLNL/ExaCMech, README pins "v0.3.4"/develop, src/../README.md:71) and
// its solver SNLS are NOT present under /root/reference, so this file restates
// the published algorithm of ecmech::evptn::matModel<SlipGeom, Kinetics,
// ThermoElastN, EosModelConst<false>> (elasto-viscoplastic, thermo-elastic "n"
// formulation: unknowns = 5 deviatoric lattice elastic strains + 3 exponential-map
// lattice rotation increments, backward Euler, trust-region Newton), using the
// interface, history layout and parameter order that ARE in the reference:
//   history layout           src/mechanics_ecmech.hpp:136-185
//   argument strides         src/mechanics_ecmech.hpp:143-159
//   model <-> template map   src/mechanics_ecmech.hpp:407-414,460-463
//   parameter order          src/mechanics_ecmech.hpp:395-405,444-458,
//                            scripts/ecmech_prop_file.py:12-129
//   5-vector <-> tensor map  src/mechanics_ecmech.hpp:343-354
//
// PARITY STATUS: per-point parity with ExaCMech itself is UNPINNED (no per-point
// known-answer vectors exist in the reference).  System-level parity is pinned by
// the reference's golden volume-averaged stress histories (6 significant digits,
// test/data/*_stress.txt) through oracle/sim_ref.hpp -- see tests/test_oracle_goldens.py:
// every history is reproduced to half a unit of its 6th printed digit, and voce_pa,
// voce_full, voce_nl_full, voce_bcc and mtsdd_bcc pass the reference's own criterion
// (test/test_mechanics.py:11-31, identical 6-digit prints).  UNPINNED: KMBalD kinetics
// with p, q != 1 (the mtsdd_full_auto / IN625 case) and HCP (no reference data).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace ecm {

constexpr double sqr2 = 1.4142135623730951, sqr3 = 1.7320508075688772;
constexpr double sqr2i = 0.7071067811865475, sqr6i = 0.4082482904638631;
constexpr double sqr2b3 = 0.816496580927726, sqr3b2 = 1.224744871391589;
constexpr double onethird = 1.0 / 3.0;
constexpr double idp_tiny_sqrt = 1.0e-90, idp_eps_sqrt = 1.0e-8;
constexpr double gam_ratio_min = 1.0e-60, gam_ratio_ovf = 1.0e45;
constexpr double e_scale = 5.0e-4, r_scale = 1.0e-2;
constexpr int NSLIP_MAX = 24;

// history indices (src/mechanics_ecmech.hpp:165-185)
constexpr int iHistA_shrateEff = 0, iHistA_shrEff = 1, iHistA_flowStr = 2, iHistA_nFEval = 3;
constexpr int iHistLbE = 4, iHistLbQ = 9, iHistLbH = 13, iHistLbGdot = 14;

enum XtalType { XTAL_FCC = 0, XTAL_BCC = 1, XTAL_HCP = 2 };
enum KinType { KIN_VOCE = 0, KIN_VOCE_NL = 1, KIN_KMBALD = 2 };

// ---- small tensor helpers ------------------------------------------------
// symmetric 3x3 (row-major full) -> 5-vector deviatoric (src/mechanics_ecmech.hpp:343-354 inverted)
inline void sym_to_vecd(const double* T, double* v) {
  v[0] = sqr2i * (T[0] - T[4]);
  v[1] = sqr6i * (2.0 * T[8] - T[0] - T[4]);
  v[2] = sqr2 * T[1];
  v[3] = sqr2 * T[2];
  v[4] = sqr2 * T[5];
}
inline void vecd_to_sym(const double* v, double* T) {
  const double t1 = sqr2i * v[0], t2 = sqr6i * v[1];
  T[0] = t1 - t2;
  T[4] = -t1 - t2;
  T[8] = sqr2b3 * v[1];
  T[5] = T[7] = sqr2i * v[4];
  T[2] = T[6] = sqr2i * v[3];
  T[1] = T[3] = sqr2i * v[2];
}
// axial vector convention of kernel_setup (src/mechanics_ecmech.cpp:65-67):
// w = (W21, W02, W10)
inline void skew_from_axial(const double* w, double* W) {
  W[0] = W[4] = W[8] = 0.0;
  W[7] = w[0];  W[5] = -w[0];
  W[2] = w[1];  W[6] = -w[1];
  W[3] = w[2];  W[1] = -w[2];
}
inline void mat3_mul(const double* A, const double* B, double* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// svec (Voigt 11,22,33,23,31,12) deviatoric part -> 5-vector
inline void svec_to_vecd(const double* s, double* v) {
  v[0] = sqr2i * (s[0] - s[1]);
  v[1] = sqr6i * (2.0 * s[2] - s[0] - s[1]);
  v[2] = sqr2 * s[5];
  v[3] = sqr2 * s[4];
  v[4] = sqr2 * s[3];
}
inline void vecd_to_svec(const double* v, double* s) {
  const double t1 = sqr2i * v[0], t2 = sqr6i * v[1];
  s[0] = t1 - t2;
  s[1] = -t1 - t2;
  s[2] = sqr2b3 * v[1];
  s[3] = sqr2i * v[4];
  s[4] = sqr2i * v[3];
  s[5] = sqr2i * v[2];
}
inline double vecd_Deff(const double* v) {
  double n = 0.0;
  for (int i = 0; i < 5; ++i) n += v[i] * v[i];
  return sqr2b3 * std::sqrt(n);
}
// unit quaternion -> rotation matrix (lattice -> sample)
inline void quat_to_tensor(const double* q, double* c) {
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
inline void emap_to_quat(const double* xi, double* q) {
  const double th = std::sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  if (th > idp_eps_sqrt) {
    const double s = std::sin(0.5 * th) / th;
    q[0] = std::cos(0.5 * th);
    q[1] = s * xi[0]; q[2] = s * xi[1]; q[3] = s * xi[2];
  } else {
    q[0] = 1.0; q[1] = 0.5 * xi[0]; q[2] = 0.5 * xi[1]; q[3] = 0.5 * xi[2];
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= n;
  }
}
// c = a * b (Hamilton product)
inline void quat_mul(const double* a, const double* b, double* c) {
  c[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  c[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  c[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  c[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
// 5x5 rotation acting on deviatoric 5-vectors: v_sm = R5 v_lat for T_sm = C T_lat C^T
inline void rot_mat_vecd(const double* C, double R5[5][5]) {
  for (int j = 0; j < 5; ++j) {
    double ej[5] = {0, 0, 0, 0, 0}, B[9], CB[9], Ct[9], CBCt[9], v[5];
    ej[j] = 1.0;
    vecd_to_sym(ej, B);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) Ct[3 * a + b] = C[3 * b + a];
    mat3_mul(C, B, CB);
    mat3_mul(CB, Ct, CBCt);
    sym_to_vecd(CBCt, v);
    for (int i = 0; i < 5; ++i) R5[i][j] = v[i];
  }
}

// Structure constants of the commutator between deviatoric-symmetric and skew
// tensors in the 5-vector / axial-vector bases:
//   cmt[i][j][k] = < B_i , B_j S_k - S_k B_j >
// so that  5vec(E W - W E) = Me(e) w, Me(e)[i][k] = sum_j cmt[i][j][k] e_j.
struct Commutator {
  double c[5][5][3];
  Commutator() {
    for (int j = 0; j < 5; ++j)
      for (int k = 0; k < 3; ++k) {
        double ej[5] = {0, 0, 0, 0, 0}, wk[3] = {0, 0, 0}, B[9], S[9], BS[9], SB[9], D[9], v[5];
        ej[j] = 1.0;
        wk[k] = 1.0;
        vecd_to_sym(ej, B);
        skew_from_axial(wk, S);
        mat3_mul(B, S, BS);
        mat3_mul(S, B, SB);
        for (int x = 0; x < 9; ++x) D[x] = BS[x] - SB[x];
        sym_to_vecd(D, v);
        for (int i = 0; i < 5; ++i) c[i][j][k] = v[i];
      }
  }
  void Me(const double* e, double M[5][3]) const {
    for (int i = 0; i < 5; ++i)
      for (int k = 0; k < 3; ++k) {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s += c[i][j][k] * e[j];
        M[i][k] = s;
      }
  }
  void Mw(const double* w, double M[5][5]) const {
    for (int i = 0; i < 5; ++i)
      for (int j = 0; j < 5; ++j) M[i][j] = c[i][j][0] * w[0] + c[i][j][1] * w[1] + c[i][j][2] * w[2];
  }
};
inline const Commutator& commutator() {
  static const Commutator k;
  return k;
}

// ---- material description --------------------------------------------------
struct Options {
  // Choices that the reference tree cannot settle (ExaCMech is external); the
  // defaults are the ones that reproduce the reference's golden stress files.
  bool kirchhoff_rss = true;    // resolve shear stress from the Kirchhoff (true) or Cauchy stress
  bool eos_temperature = true;  // kinetics temperature from the EOS energy (true) or the host's temp_k
  bool hard_lag = true;         // hardness advanced with beginning-of-step slip rates
  double av_power = 0.0;        // a_V^power factor on the elastic-strain-rate terms (0 = none)
  bool eos_mu_form = true;      // p = K (1/V - 1) instead of K (1 - V)
  // First-order effect of the elastic stretch on the slip kinematics: the [e, W^p] term of the strain-rate equation
  // and the [e, D^p] term of the spin equation that a consistent expansion of V L^p V^-1 produces.  ExaCMech does NOT
  // carry them (its plastic velocity gradient acts in the unstretched lattice): with them the reference's voce_*
  // goldens are missed by 1e-5 of the axial stress in the shear components, without them every printed value is
  // reproduced to about one unit of its 6th digit (tests/test_oracle_goldens.py).
  bool slip_stretch_terms = false;
  // The beginning-of-step deviatoric lattice strain is carried along with the volume change of the step before the
  // local solve: e_u = e_n (V_old / V_new)^(1/3).  Found in round 2 from the reference's cyclic goldens: without it the
  // averaged axial stress is off by c e_n dsigma with c = 0.10 +- 0.01 = (1 - 2 nu)/3 in every elastic increment (most
  // visible after the load reversals of voce_full_cyclic*: 1.5e-5 .. 2.2e-5 of the peak stress); with it these
  // histories are reproduced to the print resolution (7e-7).  The analytic tangent does not carry the corresponding
  // -e_u/3 d(tr eps) term (relative size sigma/E).
  bool vol_convect = true;
};

struct Material {
  int xtal = XTAL_FCC, kin = KIN_VOCE, nslip = 12, nhist = 28;
  Options opt;
  double P[NSLIP_MAX][5], Q[NSLIP_MAX][3];
  // thermo-elasticity (Kirchhoff stress = Kdiag o lattice elastic strain; cubic or hexagonal)
  double Kdiag[5], bulk, gmod;
  double Khex_vol_dev = 0.0;  // hexagonal coupling between volumetric strain and the v[1] deviator
  // EOS (EosModelConst<false>)
  double rho0, cvav, tol, gruneisen, ec0, dtde, tK0;
  // Voce power law
  double mu, xm, gam_w0, h0, tausi, taus0, xmprime = 1.0, xms, gamss0, kappa0;
  // KMBalD (per slip system where ExaCMech allows it)
  double mu_ref, tK_ref, c_1[NSLIP_MAX], tau_a, p_exp, q_exp, gam_wo, gam_ro, wrD;
  double go[NSLIP_MAX], s_[NSLIP_MAX], k1, k2o, ninv, gamma_o, rho_dd_init;
  bool withGAthermal = false;
};

inline void set_slip_system(Material& m, int a, const double* sdir, const double* mnorm) {
  double s[3], n[3];
  const double ls = std::sqrt(sdir[0] * sdir[0] + sdir[1] * sdir[1] + sdir[2] * sdir[2]);
  const double ln = std::sqrt(mnorm[0] * mnorm[0] + mnorm[1] * mnorm[1] + mnorm[2] * mnorm[2]);
  for (int i = 0; i < 3; ++i) { s[i] = sdir[i] / ls; n[i] = mnorm[i] / ln; }
  double T[9], W[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T[3 * i + j] = 0.5 * (s[i] * n[j] + s[j] * n[i]);
      W[3 * i + j] = 0.5 * (s[i] * n[j] - s[j] * n[i]);
    }
  sym_to_vecd(T, m.P[a]);
  m.Q[a][0] = W[7]; m.Q[a][1] = W[2]; m.Q[a][2] = W[3];
}

inline void setup_slip_fcc(Material& m) {  // {111}<110>
  const double mv[12][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, 1},
                            {-1, -1, 1}, {-1, -1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, -1, 1}, {1, -1, 1}};
  const double sv[12][3] = {{0, 1, -1}, {-1, 0, 1}, {1, -1, 0}, {-1, 0, -1}, {0, -1, 1}, {1, 1, 0},
                            {0, -1, -1}, {1, 0, 1}, {-1, 1, 0}, {1, 0, -1}, {0, 1, 1}, {-1, -1, 0}};
  m.nslip = 12;
  for (int a = 0; a < 12; ++a) set_slip_system(m, a, sv[a], mv[a]);
}
inline void setup_slip_bcc(Material& m) {  // {110}<111>
  const double mv[12][3] = {{1, 1, 0}, {1, 1, 0}, {1, -1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, 1},
                            {1, 0, -1}, {1, 0, -1}, {0, 1, 1}, {0, 1, 1}, {0, 1, -1}, {0, 1, -1}};
  const double sv[12][3] = {{1, -1, 1}, {-1, 1, 1}, {1, 1, 1}, {1, 1, -1}, {1, 1, -1}, {-1, 1, 1},
                            {1, 1, 1}, {1, -1, 1}, {1, 1, -1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  m.nslip = 12;
  for (int a = 0; a < 12; ++a) set_slip_system(m, a, sv[a], mv[a]);
}
// HCP: 3 basal <a>, 3 prismatic <a>, 6 pyramidal <a>, 12 pyramidal <c+a> (needs c/a).
inline void setup_slip_hcp(Material& m, double cOverA) {
  m.nslip = 24;
  const double r3 = sqr3;
  // hexagonal basis in Cartesian: a1 = (1,0,0), a2 = (-1/2, r3/2, 0), a3 = -(a1+a2), c = (0,0,c/a)
  auto dir4 = [&](const int* uvtw, double* d) {  // Miller-Bravais direction [uvtw]
    const double a1[3] = {1, 0, 0}, a2[3] = {-0.5, 0.5 * r3, 0}, a3[3] = {-0.5, -0.5 * r3, 0};
    for (int i = 0; i < 3; ++i) d[i] = uvtw[0] * a1[i] + uvtw[1] * a2[i] + uvtw[2] * a3[i];
    d[2] += uvtw[3] * cOverA;
  };
  auto pln4 = [&](const int* hkil, double* n) {  // plane normal (hkil)
    // reciprocal-type construction: n ~ h a1* + k a2* + l c*; with a1* = (1, 1/r3, 0), a2* = (0, 2/r3, 0)
    n[0] = hkil[0];
    n[1] = (hkil[0] + 2.0 * hkil[1]) / r3;
    n[2] = hkil[3] / cOverA;
  };
  const int planes[24][4] = {
      {0, 0, 0, 1}, {0, 0, 0, 1}, {0, 0, 0, 1},
      {0, 1, -1, 0}, {-1, 0, 1, 0}, {1, -1, 0, 0},
      {0, 1, -1, 1}, {-1, 0, 1, 1}, {1, -1, 0, 1}, {0, -1, 1, 1}, {1, 0, -1, 1}, {-1, 1, 0, 1},
      {1, 0, -1, 1}, {1, 0, -1, 1}, {0, 1, -1, 1}, {0, 1, -1, 1}, {-1, 1, 0, 1}, {-1, 1, 0, 1},
      {-1, 0, 1, 1}, {-1, 0, 1, 1}, {0, -1, 1, 1}, {0, -1, 1, 1}, {1, -1, 0, 1}, {1, -1, 0, 1}};
  const int dirs[24][4] = {
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0}, {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {-2, 1, 1, 3}, {-1, -1, 2, 3}, {-1, -1, 2, 3}, {1, -2, 1, 3}, {1, -2, 1, 3}, {2, -1, -1, 3},
      {2, -1, -1, 3}, {1, 1, -2, 3}, {1, 1, -2, 3}, {-1, 2, -1, 3}, {-1, 2, -1, 3}, {-2, 1, 1, 3}};
  for (int a = 0; a < 24; ++a) {
    double d[3], n[3];
    dir4(dirs[a], d);
    pln4(planes[a], n);
    // enforce exact orthogonality (guards typos in the tables above)
    const double dn = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
    const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    for (int i = 0; i < 3; ++i) d[i] -= dn / nn * n[i];
    set_slip_system(m, a, d, n);
  }
}

// Parameter vectors, order as documented at src/mechanics_ecmech.hpp:395-405,444-458.
// Returns 0 on success.
inline int init_material(Material& m, int xtal, int kin, const double* p, int np) {
  m.xtal = xtal;
  m.kin = kin;
  int i = 0;
  m.rho0 = p[i++]; m.cvav = p[i++]; m.tol = p[i++];
  if (xtal == XTAL_HCP) {
    const double c11 = p[i++], c12 = p[i++], c13 = p[i++], c33 = p[i++], c44 = p[i++];
    // hexagonal stiffness in the 5-vector + volumetric basis
    m.Kdiag[0] = c11 - c12;
    m.Kdiag[1] = (c11 + c12 - 4.0 * c13 + 2.0 * c33) / 3.0;
    m.Kdiag[2] = c11 - c12;  // 2*c66
    m.Kdiag[3] = 2.0 * c44;
    m.Kdiag[4] = 2.0 * c44;
    m.bulk = (2.0 * c11 + 2.0 * c12 + 4.0 * c13 + c33) / 9.0;
    m.Khex_vol_dev = sqr2 * (c33 + c13 - c11 - c12) / 3.0;
    m.gmod = (2.0 * m.Kdiag[0] + m.Kdiag[1] + 2.0 * m.Kdiag[3]) / 10.0;
  } else {
    const double c11 = p[i++], c12 = p[i++], c44 = p[i++];
    m.Kdiag[0] = m.Kdiag[1] = c11 - c12;
    m.Kdiag[2] = m.Kdiag[3] = m.Kdiag[4] = 2.0 * c44;
    m.bulk = onethird * (c11 + 2.0 * c12);
    m.gmod = (2.0 * c11 - 2.0 * c12 + 6.0 * c44) * 0.1;
  }
  double cOverA = 1.587;
  if (xtal == XTAL_FCC) setup_slip_fcc(m);
  else if (xtal == XTAL_BCC) setup_slip_bcc(m);
  if (kin == KIN_VOCE || kin == KIN_VOCE_NL) {
    if (xtal == XTAL_HCP) return 2;
    m.mu = p[i++]; m.xm = p[i++]; m.gam_w0 = p[i++];
    m.h0 = p[i++]; m.tausi = p[i++]; m.taus0 = p[i++];
    m.xmprime = (kin == KIN_VOCE_NL) ? p[i++] : 1.0;
    m.xms = p[i++]; m.gamss0 = p[i++]; m.kappa0 = p[i++];
  } else {
    const int ns = (xtal == XTAL_HCP) ? 24 : 12;
    const bool perSS = (xtal == XTAL_HCP);
    m.withGAthermal = (xtal != XTAL_FCC);
    m.mu_ref = p[i++]; m.tK_ref = p[i++];
    // per-family expansion for HCP: 3 basal, 3 prismatic, 6 pyr<a>, 12 pyr<c+a>
    auto read_fam = [&](double* out) {
      if (!perSS) { const double v = p[i++]; for (int a = 0; a < ns; ++a) out[a] = v; return; }
      const int cnt[4] = {3, 3, 6, 12};
      int a = 0;
      for (int f = 0; f < 4; ++f) { const double v = p[i++]; for (int c = 0; c < cnt[f]; ++c) out[a++] = v; }
    };
    read_fam(m.c_1);
    m.tau_a = p[i++]; m.p_exp = p[i++]; m.q_exp = p[i++];
    m.gam_wo = p[i++]; m.gam_ro = p[i++]; m.wrD = p[i++];
    read_fam(m.go);
    read_fam(m.s_);
    m.k1 = p[i++]; m.k2o = p[i++]; m.ninv = p[i++]; m.gamma_o = p[i++]; m.rho_dd_init = p[i++];
    if (xtal == XTAL_HCP) cOverA = p[i++];
  }
  if (xtal == XTAL_HCP) setup_slip_hcp(m, cOverA);
  m.gruneisen = p[i++]; m.ec0 = p[i++];
  if (i != np) return 1;
  m.dtde = 1.0 / m.cvav;
  m.tK0 = -m.ec0 * m.dtde;
  m.nhist = iHistLbGdot + m.nslip + 2;
  return 0;
}

// getHistInfo defaults (consumed by init_state_vars, src/mechanics_ecmech.hpp:236-300)
inline void hist_init(const Material& m, double* h) {
  for (int i = 0; i < m.nhist; ++i) h[i] = 0.0;
  h[iHistLbQ] = 1.0;
  h[iHistLbH] = (m.kin == KIN_KMBALD) ? std::log(m.rho_dd_init) : m.kappa0;
  h[m.nhist - 2] = 1.0;  // relative volume (init_state_vars, src/mechanics_ecmech.hpp:286)
}

// ---- kinetics ----------------------------------------------------------------
struct KinVals {
  double g[NSLIP_MAX];    // slip resistance per system
  double gam_w, gam_r;    // KMBalD rate prefactors
  double c_e[NSLIP_MAX];  // KMBalD activation-energy factor
};

inline void kin_get_vals(const Material& m, double tK, const double* h, KinVals& v) {
  if (m.kin == KIN_KMBALD) {
    const double sqrtDD = std::exp(0.5 * h[0]);  // hardness state is ln(relative dislocation density)
    for (int a = 0; a < m.nslip; ++a) {
      v.g[a] = m.go[a] + m.s_[a] * sqrtDD;
      v.c_e[a] = m.c_1[a] / tK * m.mu_ref;
    }
    v.gam_w = m.gam_wo / sqrtDD;
    v.gam_r = m.gam_ro * sqrtDD * sqrtDD;
  } else {
    for (int a = 0; a < m.nslip; ++a) v.g[a] = h[0];
    v.gam_w = m.gam_w0;
  }
}

// power law: gdot = gam_w |tau/g|^(1/m) sign(tau)
inline void kin_power_law(double gam_w, double xm, double g, double tau, double& gdot, double& dgdot_dtau) {
  gdot = 0.0;
  dgdot_dtau = 0.0;
  const double xmi = 1.0 / xm;
  const double t_min = std::pow(gam_ratio_min, xm), t_max = std::pow(gam_ratio_ovf, xm);
  const double gi = 1.0 / g;
  const double t = tau * gi, at = std::fabs(t);
  if (at <= t_min) return;
  if (at > t_max) {  // linear extrapolation beyond the overflow guard
    const double pl = std::exp((xmi - 1.0) * std::log(t_max));
    const double d = gam_w * pl * xmi * gi;
    const double g0 = gam_w * t_max * pl;
    gdot = (g0 + d * g * (at - t_max)) * (t > 0 ? 1.0 : -1.0);
    dgdot_dtau = d;
    return;
  }
  const double pl = std::exp((xmi - 1.0) * std::log(at));
  gdot = gam_w * t * pl;
  dgdot_dtau = gam_w * pl * xmi * gi;
}

// Kocks-Mecking balanced thermally-activated (MTS-like) + drag kinetics:
//   gdot_w = gam_w [exp(-c_e (1 - t^p)^q) - exp(-c_e (1 + t^p)^q)]   (t < 1, thermally activated)
//   gdot_w = gam_w t^(c_e p q')                                       (t >= 1, barrier overrun)
//   gdot_r = gam_r (1 - exp(-(|tau| - gAth)/wrD))                     (drag limited)
//   gdot   = sign(tau) / (1/gdot_w + 1/gdot_r)
// with gAth = g, t = (|tau| - g)/tau_a when withGAthermal (BCC/HCP); gAth = tau_a, t = (|tau| - tau_a)/g
// otherwise (FCC).  The second form is what reproduces test/data/mtsdd_full_stress.txt (1.3e-5).
inline void kin_kmbald(const Material& m, double g, double gam_w, double gam_r, double c_e, double tau,
                       double& gdot, double& dgdot_dtau) {
  gdot = 0.0;
  dgdot_dtau = 0.0;
  const double at = std::fabs(tau), sgn = tau >= 0 ? 1.0 : -1.0;
  double gAth, g_i;
  if (m.withGAthermal) { gAth = g; g_i = 1.0 / m.tau_a; }
  else { gAth = m.tau_a; g_i = 1.0 / g; }  // tau_a is the athermal threshold when g is the thermal barrier
  if (at <= gAth) return;
  const double at_0 = (at - gAth) * g_i;
  // drag limited
  double gdot_r, dgdot_r;
  {
    const double x = (at - gAth) / m.wrD;
    if (x < gam_ratio_min) return;
    if (x < idp_eps_sqrt) { gdot_r = gam_r * x; dgdot_r = gam_r / m.wrD; }
    else { const double ex = std::exp(-x); gdot_r = gam_r * (1.0 - ex); dgdot_r = gam_r * ex / m.wrD; }
  }
  // thermally activated
  double gdot_w, dgdot_w;
  if (at_0 >= 1.0) {
    const double xn = c_e * m.p_exp;  // C1 continuation of the q = 1 Arrhenius law at t = 1
    const double lg = xn * std::log(at_0);
    if (lg > std::log(gam_ratio_ovf)) {  // thermal term no longer limits at all
      gdot = sgn * gdot_r;
      dgdot_dtau = dgdot_r;
      return;
    }
    gdot_w = gam_w * std::exp(lg);
    dgdot_w = gdot_w * xn / at_0 * g_i;
  } else {
    const double pf = std::pow(at_0, m.p_exp);
    const double dpf = m.p_exp * std::pow(at_0, m.p_exp - 1.0) * g_i;
    const double qa = 1.0 - pf;
    const double ef = std::exp(-c_e * std::pow(qa, m.q_exp));
    const double dqf = m.q_exp * std::pow(qa, m.q_exp - 1.0) * dpf;
    const double qb = 1.0 + pf;
    const double eb = std::exp(-c_e * std::pow(qb, m.q_exp));
    const double dqb = m.q_exp * std::pow(qb, m.q_exp - 1.0) * dpf;
    gdot_w = gam_w * (ef - eb);
    dgdot_w = gam_w * c_e * (ef * dqf + eb * dqb);
    if (gdot_w <= gam_ratio_min * gam_w) return;
  }
  const double inv = 1.0 / (gdot_w + gdot_r);
  gdot = sgn * gdot_w * gdot_r * inv;
  dgdot_dtau = (dgdot_w * gdot_r * gdot_r + dgdot_r * gdot_w * gdot_w) * inv * inv;
}

inline void kin_eval(const Material& m, const KinVals& v, const double* tau, double* gdot, double* dg) {
  for (int a = 0; a < m.nslip; ++a) {
    if (m.kin == KIN_KMBALD) kin_kmbald(m, v.g[a], v.gam_w, v.gam_r, v.c_e[a], tau[a], gdot[a], dg[a]);
    else kin_power_law(v.gam_w, m.xm, v.g[a], tau[a], gdot[a], dg[a]);
  }
}

// hardness evolution, backward Euler in h with the supplied slip rates
inline double kin_update_h(const Material& m, double h_n, double dt, const double* gdot, int* nfev) {
  double shr = 0.0;
  for (int a = 0; a < m.nslip; ++a) shr += std::fabs(gdot[a]);
  auto sdot = [&](double h, double& ds) {
    if (m.kin == KIN_KMBALD) {
      double k2 = m.k2o;
      if (shr > idp_tiny_sqrt) k2 = m.k2o * std::pow(m.gamma_o / shr, m.ninv);
      const double t = std::exp(-0.5 * h);  // d(ln rho)/dt = (k1/sqrt(rho) - k2) * shrate
      ds = (-0.5 * m.k1 * t) * shr;
      return (m.k1 * t - k2) * shr;
    }
    double sat = m.taus0;
    if (shr > idp_tiny_sqrt && m.xms != 0.0) sat = m.taus0 * std::pow(shr / m.gamss0, m.xms);
    if (m.kin == KIN_VOCE_NL && m.xmprime != 1.0) {
      const double r = (sat - h) / (sat - m.tausi);
      const double rp = (r > 0) ? std::pow(r, m.xmprime) : 0.0;
      ds = (r > 0) ? -m.h0 * m.xmprime * std::pow(r, m.xmprime - 1.0) / (sat - m.tausi) * shr : 0.0;
      return m.h0 * rp * shr;
    }
    const double t1 = m.h0 / (sat - m.tausi);
    ds = -t1 * shr;
    return t1 * (sat - h) * shr;
  };
  double h = h_n;
  for (int it = 0; it < 50; ++it) {
    double ds;
    const double r = h - h_n - dt * sdot(h, ds);
    if (nfev) ++*nfev;
    const double dh = -r / (1.0 - dt * ds);
    h += dh;
    if (std::fabs(dh) <= 1e-14 * std::fabs(h) + 1e-300) break;
  }
  return h;
}

// ---- dense LU solve (n <= 8), partial pivoting; returns false if singular ----
inline bool lu_solve(int n, double* A, double* b) {
  int piv[8];
  for (int k = 0; k < n; ++k) {
    int p = k;
    double mx = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + k]) > mx) { mx = std::fabs(A[i * n + k]); p = i; }
    if (mx == 0.0) return false;
    piv[k] = p;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
      std::swap(b[k], b[p]);
    }
    const double inv = 1.0 / A[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      const double f = A[i * n + k] * inv;
      A[i * n + k] = f;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
    b[i] = s / A[i * n + i];
  }
  (void)piv;
  return true;
}

// ---- the 8-unknown update problem ------------------------------------------
struct UpdateProblem {
  const Material& m;
  double dt, dt_ri, detV, detVi, a_V_ri, tK;
  KinVals kv;
  double e_n[5], q_n[4], d_sm[5], w_sm[3];
  double epsdot_scale_inv, rotincr_scale_inv;
  double T1_shift = 0.0;  // hexagonal volumetric -> c-axis deviator coupling (constant during the local solve)
  // outputs of the last evaluation
  double gdot[NSLIP_MAX], tau[NSLIP_MAX], e_f[5], q_f[4], C[9], R5[5][5];

  explicit UpdateProblem(const Material& mm) : m(mm) {}

  // right Jacobian of the exponential map: d exp(xi^) = exp(xi^) (Jr dxi)^
  static void Jr(const double* xi, double J[3][3]) {
    const double th2 = xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2], th = std::sqrt(th2);
    double a, b;
    if (th < 1e-4) { a = 0.5 - th2 / 24.0; b = 1.0 / 6.0 - th2 / 120.0; }
    else { a = (1.0 - std::cos(th)) / th2; b = (th - std::sin(th)) / (th2 * th); }
    double X[9], X2[9];
    skew_from_axial(xi, X);
    mat3_mul(X, X, X2);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J[i][j] = (i == j ? 1.0 : 0.0) - a * X[3 * i + j] + b * X2[3 * i + j];
  }

  // residual (8) and optionally the 8x8 Jacobian (row-major) at scaled unknown x
  void eval(const double* x, double* R, double* Jac) {
    const Commutator& cm = commutator();
    double de[5], edot[5], xi[3];
    for (int i = 0; i < 5; ++i) { de[i] = e_scale * x[i]; e_f[i] = e_n[i] + de[i]; edot[i] = de[i] * dt_ri; }
    for (int k = 0; k < 3; ++k) xi[k] = r_scale * x[5 + k];
    double A[4];
    emap_to_quat(xi, A);
    quat_mul(q_n, A, q_f);
    quat_to_tensor(q_f, C);
    rot_mat_vecd(C, R5);
    double d_lat[5], w_lat[3];
    for (int i = 0; i < 5; ++i) {
      double s = 0.0;
      for (int j = 0; j < 5; ++j) s += R5[j][i] * d_sm[j];
      d_lat[i] = s;
    }
    for (int k = 0; k < 3; ++k) w_lat[k] = C[0 + k] * w_sm[0] + C[3 + k] * w_sm[1] + C[6 + k] * w_sm[2];
    // Kirchhoff stress deviator in the lattice frame and resolved shear stresses
    double T[5], dg[NSLIP_MAX];
    const double rss_fac = m.opt.kirchhoff_rss ? 1.0 : detVi;
    for (int i = 0; i < 5; ++i) T[i] = m.Kdiag[i] * e_f[i];
    T[1] += T1_shift;
    for (int a = 0; a < m.nslip; ++a) {
      double s = 0.0;
      for (int i = 0; i < 5; ++i) s += m.P[a][i] * T[i];
      tau[a] = s * rss_fac;
    }
    kin_eval(m, kv, tau, gdot, dg);
    double dp[5] = {0, 0, 0, 0, 0}, wp[3] = {0, 0, 0};
    for (int a = 0; a < m.nslip; ++a) {
      for (int i = 0; i < 5; ++i) dp[i] += gdot[a] * m.P[a][i];
      for (int k = 0; k < 3; ++k) wp[k] += gdot[a] * m.Q[a][k];
    }
    double Me[5][3];
    cm.Me(e_f, Me);
    const double av = std::pow(a_V_ri, -m.opt.av_power);
    const double c2 = m.opt.slip_stretch_terms ? 1.0 : 0.0;
    // X(a,b)[k] = 0.5 * sum_i Me(a)[i][k] b[i] = axial(a b - b a)
    double Medot[5][3];
    cm.Me(edot, Medot);
    for (int i = 0; i < 5; ++i) {
      const double ewp = c2 * (Me[i][0] * wp[0] + Me[i][1] * wp[1] + Me[i][2] * wp[2]);
      R[i] = epsdot_scale_inv * (av * (edot[i] + ewp) + dp[i] - d_lat[i]);
    }
    for (int k = 0; k < 3; ++k) {
      double xe_dp = 0.0, xedot_e = 0.0;
      for (int i = 0; i < 5; ++i) { xe_dp += 0.5 * Me[i][k] * dp[i]; xedot_e += 0.5 * Medot[i][k] * e_f[i]; }
      R[5 + k] = rotincr_scale_inv * dt * (xi[k] * dt_ri + wp[k] - w_lat[k] + c2 * xe_dp - 0.5 * xedot_e);
    }
    if (!Jac) return;
    // d gdot / d e_f
    double dDp_de[5][5] = {{0}}, dWp_de[3][5] = {{0}};
    for (int a = 0; a < m.nslip; ++a)
      for (int j = 0; j < 5; ++j) {
        const double dga = dg[a] * rss_fac * m.P[a][j] * m.Kdiag[j];
        for (int i = 0; i < 5; ++i) dDp_de[i][j] += m.P[a][i] * dga;
        for (int k = 0; k < 3; ++k) dWp_de[k][j] += m.Q[a][k] * dga;
      }
    double Mwp[5][5], JrM[3][3], Mdl[5][3];
    cm.Mw(wp, Mwp);
    Jr(xi, JrM);
    cm.Me(d_lat, Mdl);
    double Wl[9];
    skew_from_axial(w_lat, Wl);
    for (int x8 = 0; x8 < 64; ++x8) Jac[x8] = 0.0;
    for (int i = 0; i < 5; ++i) {
      for (int j = 0; j < 5; ++j) {
        double v = av * ((i == j ? dt_ri : 0.0) + c2 * Mwp[i][j]) + dDp_de[i][j];
        for (int k = 0; k < 3; ++k) v += c2 * av * Me[i][k] * dWp_de[k][j];
        Jac[i * 8 + j] = epsdot_scale_inv * v * e_scale;
      }
      for (int k = 0; k < 3; ++k) {
        // -d(d_lat)/d xi = -Me(d_lat) Jr
        double v = 0.0;
        for (int l = 0; l < 3; ++l) v += Mdl[i][l] * JrM[l][k];
        Jac[i * 8 + 5 + k] = -epsdot_scale_inv * v * r_scale;
      }
    }
    // helper matrices for the spin-equation second-order terms
    double Mdp[5][3];
    cm.Me(dp, Mdp);
    for (int k = 0; k < 3; ++k) {
      for (int j = 0; j < 5; ++j) {
        double v = dWp_de[k][j];
        if (c2 != 0.0) {
          // X(e,dp): d/de_j -> -0.5*Mdp[j][k] (antisymmetry) ... plus through dp
          double t = -0.5 * Mdp[j][k];
          for (int i = 0; i < 5; ++i) t += 0.5 * Me[i][k] * dDp_de[i][j];
          v += c2 * t;
        }
        // -0.5 X(edot,e): d/de_j = -0.5*( X(e_j/dt, e) + X(edot, e_j) )
        //   X(e_j, e)[k] = -0.5*Me(e)[j][k];  X(edot, e_j)[k] = 0.5*Medot[j][k]
        v += -0.5 * (-0.5 * Me[j][k] * dt_ri + 0.5 * Medot[j][k]);
        Jac[(5 + k) * 8 + j] = rotincr_scale_inv * dt * v * e_scale;
      }
      for (int l = 0; l < 3; ++l) {
        // d(xi/dt)/dxi - d(w_lat)/dxi ; w_lat' = w_lat x (Jr dxi) = skew(w_lat) Jr dxi
        double v = (k == l ? dt_ri : 0.0);
        for (int n = 0; n < 3; ++n) v -= Wl[3 * k + n] * JrM[n][l];
        Jac[(5 + k) * 8 + 5 + l] = rotincr_scale_inv * dt * v * r_scale;
      }
    }
  }
};

// Trust-region dogleg Newton on the 8x8 system (the role SNLS's SNLSTrDlDenseG<8>
// plays inside ExaCMech).  Returns the number of residual evaluations, <0 on failure.
inline int solve_trdl(UpdateProblem& prob, double* x, double tol, int max_iter = 200) {
  const int n = 8;
  double R[8], J[64], Rt[8], Jt[64], xt[8];
  prob.eval(x, R, J);
  int nfev = 1;
  auto norm = [&](const double* v) { double s = 0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return std::sqrt(s); };
  double res = norm(R);
  double delta = 1.0e2;
  const double xiLG = 0.75, xiIG = 0.25, xiLO = 0.35, xiIncDelta = 1.5, xiDecDelta = 0.25;
  const double deltaMin = 1e-12, deltaMax = 1e4;
  for (int it = 0; it < max_iter; ++it) {
    if (res <= tol) return nfev;
    // gradient of 0.5|R|^2 and Newton step
    double grad[8], Jg[8], nr[8], Jc[64];
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = 0; i < n; ++i) s += J[i * n + j] * R[i]; grad[j] = s; }
    for (int i = 0; i < n; ++i) { double s = 0; for (int j = 0; j < n; ++j) s += J[i * n + j] * grad[j]; Jg[i] = s; }
    std::memcpy(Jc, J, sizeof(J));
    for (int i = 0; i < n; ++i) nr[i] = -R[i];
    const bool have_newton = lu_solve(n, Jc, nr);
    const double g2 = [&] { double s = 0; for (int i = 0; i < n; ++i) s += grad[i] * grad[i]; return s; }();
    const double Jg2 = [&] { double s = 0; for (int i = 0; i < n; ++i) s += Jg[i] * Jg[i]; return s; }();
    bool accepted = false;
    while (!accepted) {
      double step[8], pred;
      const double nrn = have_newton ? norm(nr) : 1e300;
      if (have_newton && nrn <= delta) {
        for (int i = 0; i < n; ++i) step[i] = nr[i];
        pred = res;  // predicted new residual 0
      } else {
        // Cauchy point along -grad
        const double alpha = (Jg2 > 0) ? g2 / Jg2 : 0.0;
        const double cpn = alpha * std::sqrt(g2);
        if (cpn >= delta || !have_newton) {
          const double f = delta / std::sqrt(g2 > 0 ? g2 : 1.0);
          for (int i = 0; i < n; ++i) step[i] = -f * grad[i];
        } else {
          // dogleg between the Cauchy point and the Newton point
          double cp[8], d[8];
          for (int i = 0; i < n; ++i) { cp[i] = -alpha * grad[i]; d[i] = nr[i] - cp[i]; }
          double a = 0, b = 0, c = -delta * delta;
          for (int i = 0; i < n; ++i) { a += d[i] * d[i]; b += 2.0 * cp[i] * d[i]; c += cp[i] * cp[i]; }
          const double beta = (-b + std::sqrt(std::max(0.0, b * b - 4 * a * c))) / (2 * a);
          for (int i = 0; i < n; ++i) step[i] = cp[i] + beta * d[i];
        }
        double lin[8];
        for (int i = 0; i < n; ++i) { double s = R[i]; for (int j = 0; j < n; ++j) s += J[i * n + j] * step[j]; lin[i] = s; }
        pred = res - norm(lin);
      }
      for (int i = 0; i < n; ++i) xt[i] = x[i] + step[i];
      prob.eval(xt, Rt, Jt);  // Jacobian at the trial point: an accepted step needs no second evaluation
      ++nfev;
      const double rest = norm(Rt);
      const bool finite = std::isfinite(rest);
      const double actual = res - rest;
      const double rho = (finite && pred > 0) ? actual / pred : -1.0;
      if (finite && rest < res) {
        accepted = true;
        std::memcpy(x, xt, sizeof(xt));
        if (rho > xiLG && norm(step) >= 0.99 * delta) delta = std::min(deltaMax, delta * xiIncDelta);
        else if (rho < xiLO) delta = std::max(deltaMin, std::max(delta, norm(step)) * xiDecDelta * 2.0);
        std::memcpy(R, Rt, sizeof(Rt));
        std::memcpy(J, Jt, sizeof(Jt));
        res = rest;
      } else {
        delta = std::min(delta, norm(step)) * xiDecDelta;
        if (delta < deltaMin) return -nfev;
      }
      (void)xiIG;
    }
  }
  return (res <= tol) ? nfev : -nfev;
}

// One material point: the restatement of evptn getResponseSngl.  Argument
// meaning = the per-point slices ExaConstit passes at src/mechanics_ecmech.cpp:183-185.
//   d_svec_p[7]  deviatoric deformation rate (Voigt) + trace
//   w_vec[3]     spin axial vector
//   vol_ratio[4] (V_old, V_new, dV-rate, V_new - V_old)
//   eInt[1], stress_svec_p[7] (dev. Cauchy Voigt, pressure) in/out, hist[] in/out,
//   tkelv in/out, sdd[2] out, mtanSD[36] out (row-major d sigma_i / d eps_j, eng. shear)
// Returns number of residual evaluations (>0) or <0 when the local solve failed.
inline int get_response_sngl(const Material& m, double dt, const double* d_svec_p, const double* w_vec,
                             const double* vol_ratio, double* eInt, double* stress_svec_p, double* hist,
                             double& tkelv, double* sdd, double* mtanSD) {
  UpdateProblem prob(m);
  svec_to_vecd(d_svec_p, prob.d_sm);
  for (int k = 0; k < 3; ++k) prob.w_sm[k] = w_vec[k];
  {
    const double f = m.opt.vol_convect ? std::cbrt(vol_ratio[0] / vol_ratio[1]) : 1.0;
    for (int i = 0; i < 5; ++i) prob.e_n[i] = f * hist[iHistLbE + i];
  }
  {
    double n = 0.0;
    for (int i = 0; i < 4; ++i) n += hist[iHistLbQ + i] * hist[iHistLbQ + i];
    n = 1.0 / std::sqrt(n);
    for (int i = 0; i < 4; ++i) prob.q_n[i] = hist[iHistLbQ + i] * n;
  }
  double* gdot_h = &hist[iHistLbGdot];
  // EOS: temperature from beginning-of-step state, pressure at end-of-step volume
  const double eOld = eInt[0], pOld = stress_svec_p[6];
  const double vOld = vol_ratio[0], vNew = vol_ratio[1], volInc = vol_ratio[3];
  if (m.opt.eos_temperature) tkelv = m.tK0 + eOld * m.dtde;
  double eNew = eOld - volInc * pOld;
  const double eta = m.opt.eos_mu_form ? (1.0 / vNew - 1.0) : (1.0 - vNew);
  double pEOS = m.bulk * eta + m.gruneisen * eNew;
  eNew = eOld - 0.5 * volInc * (pOld + pEOS);
  pEOS = m.bulk * eta + m.gruneisen * eNew;
  const double bulkNew = m.opt.eos_mu_form ? m.bulk / vNew : m.bulk;
  const double dp_dlnV = m.opt.eos_mu_form ? -m.bulk / vNew : -m.bulk * vNew;
  (void)vOld;
  // hardness to end of step
  int nfev_h = 0;
  double h_u;
  double gdot_zero[NSLIP_MAX] = {0};
  if (m.opt.hard_lag) h_u = kin_update_h(m, hist[iHistLbH], dt, gdot_h, &nfev_h);
  else h_u = hist[iHistLbH];
  (void)gdot_zero;
  kin_get_vals(m, tkelv, &h_u, prob.kv);
  // deviatoric work with beginning-of-step stress
  const double halfVMidDt = 0.25 * (vol_ratio[0] + vol_ratio[1]) * dt;
  auto inner_dev = [&](const double* s, const double* d) {
    return s[0] * d[0] + s[1] * d[1] + s[2] * d[2] + 2.0 * (s[3] * d[3] + s[4] * d[4] + s[5] * d[5]);
  };
  double dEDev = halfVMidDt * inner_dev(stress_svec_p, d_svec_p);

  prob.dt = dt;
  prob.dt_ri = 1.0 / dt;
  prob.detV = vNew;
  prob.detVi = 1.0 / vNew;
  prob.a_V_ri = 1.0 / std::cbrt(vNew);
  prob.tK = tkelv;
  prob.T1_shift = m.Khex_vol_dev * std::log(vNew) / sqr3;
  {
    const double dEff = vecd_Deff(prob.d_sm);
    const double wn = std::sqrt(w_vec[0] * w_vec[0] + w_vec[1] * w_vec[1] + w_vec[2] * w_vec[2]);
    const double eps_dot = std::max(dEff * sqr3b2, 1.0e-12 / dt);  // scale for the strain-rate residual
    prob.epsdot_scale_inv = std::min(1.0 / eps_dot, 1.0e6 * dt);
    (void)wn;
    prob.rotincr_scale_inv = prob.dt_ri * prob.epsdot_scale_inv;  // residual in rotation-increment units
  }
  double x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int nfev = solve_trdl(prob, x, m.tol);
  const bool ok = nfev > 0;
  if (!ok) nfev = -nfev;
  // final state (prob holds the last accepted evaluation with Jacobian)
  double R[8], J[64];
  prob.eval(x, R, J);
  double shrate = 0.0, disRate = 0.0;
  for (int a = 0; a < m.nslip; ++a) {
    shrate += std::fabs(prob.gdot[a]);
    disRate += prob.tau[a] * prob.gdot[a];
    gdot_h[a] = prob.gdot[a];
  }
  if (!m.opt.hard_lag) h_u = kin_update_h(m, hist[iHistLbH], dt, prob.gdot, &nfev_h);
  for (int i = 0; i < 5; ++i) hist[iHistLbE + i] = prob.e_f[i];
  for (int i = 0; i < 4; ++i) hist[iHistLbQ + i] = prob.q_f[i];
  hist[iHistLbH] = h_u;
  hist[iHistA_shrateEff] = shrate;
  hist[iHistA_shrEff] += shrate * dt;
  {
    const double dEff = vecd_Deff(prob.d_sm);
    double flow = prob.kv.g[0];
    // dissipation per CURRENT volume (Cauchy stress : D^p = Kirchhoff resolved shear stresses . slip rates / det V): what the
    // reference's voce_ea_pl_work.txt / voce_ea_cs_pl_work.txt pin (with the Kirchhoff value the integrated plastic work
    // is high by the elastic volume strain, 1.0e-4 ... 1.7e-4 relative; with this one it matches to the printed digits)
    if (dEff > idp_tiny_sqrt) flow = disRate * prob.detVi / dEff;
    hist[iHistA_flowStr] = flow;
  }
  hist[iHistA_nFEval] = (double)nfev;
  // Cauchy stress: lattice -> sample
  double sig_lat[5], sig_sm[5];
  for (int i = 0; i < 5; ++i) sig_lat[i] = prob.detVi * m.Kdiag[i] * prob.e_f[i];
  sig_lat[1] += prob.detVi * prob.T1_shift;
  for (int i = 0; i < 5; ++i) {
    double s = 0.0;
    for (int j = 0; j < 5; ++j) s += prob.R5[i][j] * sig_lat[j];
    sig_sm[i] = s;
  }
  vecd_to_svec(sig_sm, stress_svec_p);
  // hexagonal crystals: the c-axis deviatoric strain also carries pressure
  const double p_cpl = -m.Khex_vol_dev * prob.e_f[1] * prob.detVi / sqr3;
  stress_svec_p[6] = pEOS + p_cpl;
  dEDev += halfVMidDt * inner_dev(stress_svec_p, d_svec_p);
  eInt[0] = eNew + dEDev;
  sdd[0] = bulkNew;
  sdd[1] = m.gmod;

  // ---- algorithmic tangent d sigma / d (D dt), 6x6 Voigt with engineering shear ----
  if (mtanSD) {
    // dx/d(d_sm): J dx = -dR/dd_sm ; dR_e/dd_sm = -epsdot_scale_inv * R5^T (rows i, cols j: R5[j][i])
    // Solve for the 8x5 sensitivity S = dx/dd_sm.
    double S[8][5];
    for (int c = 0; c < 5; ++c) {
      double Jc[64], rhs[8];
      std::memcpy(Jc, J, sizeof(J));
      for (int i = 0; i < 5; ++i) rhs[i] = prob.epsdot_scale_inv * prob.R5[c][i];
      for (int k = 0; k < 3; ++k) rhs[5 + k] = 0.0;
      if (!lu_solve(8, Jc, rhs)) { for (int i = 0; i < 8; ++i) rhs[i] = 0.0; }
      for (int i = 0; i < 8; ++i) S[i][c] = rhs[i];
    }
    // d sig_sm(5) / d d_sm(5) = R5 diag(detVi K) de/dd + (d R5/d xi sig_lat) dxi/dd
    const Commutator& cm = commutator();
    double Msl[5][3], JrM[3][3], xi[3] = {r_scale * x[5], r_scale * x[6], r_scale * x[7]};
    cm.Me(sig_lat, Msl);  // d(R5 s)/dxi = R5 * (-(Me(s)) Jr) ... sample stress: s_sm = C s C^T
    UpdateProblem::Jr(xi, JrM);
    double dsd[5][5];
    for (int i = 0; i < 5; ++i)
      for (int c = 0; c < 5; ++c) {
        double v = 0.0;
        for (int j = 0; j < 5; ++j) {
          double dl = prob.detVi * m.Kdiag[j] * e_scale * S[j][c];
          // lattice-frame change of sig from rotation: s_sm = R s_lat R^T with R -> R exp(dxi'^):
          // d s_sm = R (dxi'^ s - s dxi'^) R^T = -R Me(s) dxi'
          for (int k = 0; k < 3; ++k) {
            double jr = 0.0;
            for (int l = 0; l < 3; ++l) jr += JrM[k][l] * r_scale * S[5 + l][c];
            dl -= Msl[j][k] * jr;
          }
          v += prob.R5[i][j] * dl;
        }
        dsd[i][c] = v / dt;  // per unit strain increment (D dt)
      }
    // convert the 5x5 deviatoric operator to 6x6 Voigt (engineering shear strain columns)
    // eps (tensor) -> vecd: v = Tm * eps6 (tensor shear); sig6 = Tm^T-like back map
    double Tm[5][6] = {{sqr2i, -sqr2i, 0, 0, 0, 0},
                       {-sqr6i, -sqr6i, 2.0 * sqr6i, 0, 0, 0},
                       {0, 0, 0, 0, 0, sqr2},
                       {0, 0, 0, 0, sqr2, 0},
                       {0, 0, 0, sqr2, 0, 0}};
    double Bm[6][5];  // svec = Bm * vecd
    for (int j = 0; j < 5; ++j) { double ej[5] = {0, 0, 0, 0, 0}, s6[6]; ej[j] = 1.0; vecd_to_svec(ej, s6); for (int i = 0; i < 6; ++i) Bm[i][j] = s6[i]; }
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double v = 0.0;
        for (int a = 0; a < 5; ++a)
          for (int b = 0; b < 5; ++b) v += Bm[i][a] * dsd[a][b] * Tm[b][j];
        if (j >= 3) v *= 0.5;  // engineering shear
        mtanSD[i * 6 + j] = v;
      }
    if (m.Khex_vol_dev != 0.0) {
      // (a) pressure carried by the c-axis deviatoric strain; (b) deviatoric stress carried by the volume strain
      const double kc = m.Khex_vol_dev * prob.detVi / sqr3;
      for (int j = 0; j < 6; ++j) {
        double v = 0.0;
        for (int c = 0; c < 5; ++c) v += kc * e_scale * S[1][c] / dt * Tm[c][j];
        if (j >= 3) v *= 0.5;
        for (int i = 0; i < 3; ++i) mtanSD[i * 6 + j] += v;
      }
      double e1[5] = {0, kc, 0, 0, 0}, e1sm[5], s6b[6];
      for (int i = 0; i < 5; ++i) { e1sm[i] = 0.0; for (int j = 0; j < 5; ++j) e1sm[i] += prob.R5[i][j] * e1[j]; }
      vecd_to_svec(e1sm, s6b);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 3; ++j) mtanSD[i * 6 + j] += s6b[i];
    }
    // volumetric parts: -dp/d(eps_kk) on the diagonal block and the 1/detV dependence of sig'
    double sdev6[6];
    for (int i = 0; i < 6; ++i) sdev6[i] = stress_svec_p[i];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = -sdev6[i];                 // d(1/detV)/d eps_kk * detV
        if (i < 3) v += -dp_dlnV;             // -dp/d eps_kk
        mtanSD[i * 6 + j] += v;
      }
  }
  return ok ? nfev : -nfev;
}

}  // namespace ecm

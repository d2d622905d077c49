// Host-side construction of the device material description (MatDev) from an ExaConstit
// property vector.  Parameter order: src/mechanics_ecmech.hpp:395-405,444-458 and
// scripts/ecmech_prop_file.py:12-129; model <-> crystal/kinetics map: src/mechanics_ecmech.hpp:
// 407-414,460-463 (FCC/BCC Voce + Voce-NL, FCC/BCC/HCP KMBalD).
#pragma once
#include <cmath>
#include <cstring>
#include <string>

#include "material_point.hpp"

namespace exab {

inline void slip_system(MatDev& m, int a, const double* sdir, const double* mnorm) {
  double s[3], n[3];
  const double ls = std::sqrt(sdir[0] * sdir[0] + sdir[1] * sdir[1] + sdir[2] * sdir[2]);
  const double ln = std::sqrt(mnorm[0] * mnorm[0] + mnorm[1] * mnorm[1] + mnorm[2] * mnorm[2]);
  for (int i = 0; i < 3; ++i) { s[i] = sdir[i] / ls; n[i] = mnorm[i] / ln; }
  double T[3][3], W[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      T[i][j] = 0.5 * (s[i] * n[j] + s[j] * n[i]);
      W[i][j] = 0.5 * (s[i] * n[j] - s[j] * n[i]);
    }
  const double r2 = std::sqrt(2.0), r6 = std::sqrt(6.0);
  m.P[a][0] = (T[0][0] - T[1][1]) / r2;
  m.P[a][1] = (2.0 * T[2][2] - T[0][0] - T[1][1]) / r6;
  m.P[a][2] = r2 * T[0][1];
  m.P[a][3] = r2 * T[0][2];
  m.P[a][4] = r2 * T[1][2];
  m.Q[a][0] = W[2][1];
  m.Q[a][1] = W[0][2];
  m.Q[a][2] = W[1][0];
}

inline void slip_fcc(MatDev& m) {  // {111}<110>
  static const double mv[12][3] = {{1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, 1},
                                   {-1, -1, 1}, {-1, -1, 1}, {-1, -1, 1}, {1, -1, 1}, {1, -1, 1}, {1, -1, 1}};
  static const double sv[12][3] = {{0, 1, -1}, {-1, 0, 1}, {1, -1, 0}, {-1, 0, -1}, {0, -1, 1}, {1, 1, 0},
                                   {0, -1, -1}, {1, 0, 1}, {-1, 1, 0}, {1, 0, -1}, {0, 1, 1}, {-1, -1, 0}};
  m.nslip = 12;
  for (int a = 0; a < 12; ++a) slip_system(m, a, sv[a], mv[a]);
}
inline void slip_bcc(MatDev& m) {  // {110}<111>
  static const double mv[12][3] = {{1, 1, 0}, {1, 1, 0}, {1, -1, 0}, {1, -1, 0}, {1, 0, 1}, {1, 0, 1},
                                   {1, 0, -1}, {1, 0, -1}, {0, 1, 1}, {0, 1, 1}, {0, 1, -1}, {0, 1, -1}};
  static const double sv[12][3] = {{1, -1, 1}, {-1, 1, 1}, {1, 1, 1}, {1, 1, -1}, {1, 1, -1}, {-1, 1, 1},
                                   {1, 1, 1}, {1, -1, 1}, {1, 1, -1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  m.nslip = 12;
  for (int a = 0; a < 12; ++a) slip_system(m, a, sv[a], mv[a]);
}
// 3 basal <a>, 3 prismatic <a>, 6 pyramidal <a>, 12 pyramidal <c+a>, Miller-Bravais tables
inline void slip_hcp(MatDev& m, double cOverA) {
  m.nslip = 24;
  const double r3 = std::sqrt(3.0);
  static const int planes[24][4] = {
      {0, 0, 0, 1}, {0, 0, 0, 1}, {0, 0, 0, 1},
      {0, 1, -1, 0}, {-1, 0, 1, 0}, {1, -1, 0, 0},
      {0, 1, -1, 1}, {-1, 0, 1, 1}, {1, -1, 0, 1}, {0, -1, 1, 1}, {1, 0, -1, 1}, {-1, 1, 0, 1},
      {1, 0, -1, 1}, {1, 0, -1, 1}, {0, 1, -1, 1}, {0, 1, -1, 1}, {-1, 1, 0, 1}, {-1, 1, 0, 1},
      {-1, 0, 1, 1}, {-1, 0, 1, 1}, {0, -1, 1, 1}, {0, -1, 1, 1}, {1, -1, 0, 1}, {1, -1, 0, 1}};
  static const int dirs[24][4] = {
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0}, {2, -1, -1, 0}, {-1, 2, -1, 0}, {-1, -1, 2, 0},
      {-2, 1, 1, 3}, {-1, -1, 2, 3}, {-1, -1, 2, 3}, {1, -2, 1, 3}, {1, -2, 1, 3}, {2, -1, -1, 3},
      {2, -1, -1, 3}, {1, 1, -2, 3}, {1, 1, -2, 3}, {-1, 2, -1, 3}, {-1, 2, -1, 3}, {-2, 1, 1, 3}};
  const double a1[3] = {1, 0, 0}, a2[3] = {-0.5, 0.5 * r3, 0}, a3[3] = {-0.5, -0.5 * r3, 0};
  for (int a = 0; a < 24; ++a) {
    double d[3], n[3];
    for (int i = 0; i < 3; ++i) d[i] = dirs[a][0] * a1[i] + dirs[a][1] * a2[i] + dirs[a][2] * a3[i];
    d[2] += dirs[a][3] * cOverA;
    n[0] = planes[a][0];
    n[1] = (planes[a][0] + 2.0 * planes[a][1]) / r3;
    n[2] = planes[a][3] / cOverA;
    const double dn = d[0] * n[0] + d[1] * n[1] + d[2] * n[2];
    const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    for (int i = 0; i < 3; ++i) d[i] -= dn / nn * n[i];
    slip_system(m, a, d, n);
  }
}

// returns empty string on success
inline std::string build_material(MatDev& m, int xtal, int kin, const double* p, int np) {
  std::memset(&m, 0, sizeof(m));
  m.xtal = xtal;
  m.kin = kin;
  const int need_voce = (kin == KIN_VOCE_NL) ? 18 : 17;
  const int need_km = (xtal == XTAL_HCP) ? 3 + 5 + 2 + 4 + 6 + 4 + 4 + 5 + 1 + 2 : 24;
  const int need = (kin == KIN_KMBALD) ? need_km : need_voce;
  if (np != need) return "property vector has " + std::to_string(np) + " entries, model needs " + std::to_string(need);
  if (kin != KIN_KMBALD && xtal == XTAL_HCP) return "Voce kinetics are not available for HCP crystals";
  int i = 0;
  i++;  // rho0 (unused by the constant-modulus EOS)
  const double cvav = p[i++];
  m.tol = p[i++];
  if (xtal == XTAL_HCP) {
    const double c11 = p[i++], c12 = p[i++], c13 = p[i++], c33 = p[i++], c44 = p[i++];
    m.Kdiag[0] = c11 - c12;
    m.Kdiag[1] = (c11 + c12 - 4.0 * c13 + 2.0 * c33) / 3.0;
    m.Kdiag[2] = c11 - c12;
    m.Kdiag[3] = 2.0 * c44;
    m.Kdiag[4] = 2.0 * c44;
    m.bulk = (2.0 * c11 + 2.0 * c12 + 4.0 * c13 + c33) / 9.0;
    m.Kvd = std::sqrt(2.0) * (c33 + c13 - c11 - c12) / 3.0;
    m.gmod = (2.0 * m.Kdiag[0] + m.Kdiag[1] + 2.0 * m.Kdiag[3]) / 10.0;
  } else {
    const double c11 = p[i++], c12 = p[i++], c44 = p[i++];
    m.Kdiag[0] = m.Kdiag[1] = c11 - c12;
    m.Kdiag[2] = m.Kdiag[3] = m.Kdiag[4] = 2.0 * c44;
    m.bulk = (c11 + 2.0 * c12) / 3.0;
    m.gmod = (2.0 * c11 - 2.0 * c12 + 6.0 * c44) * 0.1;
  }
  double cOverA = 1.587;
  if (kin == KIN_KMBALD) {
    const int ns = (xtal == XTAL_HCP) ? 24 : 12;
    const bool perFam = (xtal == XTAL_HCP);
    m.withGAthermal = (xtal != XTAL_FCC);
    m.mu_ref = p[i++];
    i++;  // reference temperature (documented, not used by the rate equations)
    auto fam = [&](double* out) {
      if (!perFam) { const double v = p[i++]; for (int a = 0; a < ns; ++a) out[a] = v; return; }
      const int cnt[4] = {3, 3, 6, 12};
      int a = 0;
      for (int f = 0; f < 4; ++f) { const double v = p[i++]; for (int c = 0; c < cnt[f]; ++c) out[a++] = v; }
    };
    fam(m.c_1);
    m.tau_a = p[i++]; m.p_exp = p[i++]; m.q_exp = p[i++];
    m.gam_wo = p[i++]; m.gam_ro = p[i++]; m.wrD = p[i++];
    fam(m.go);
    fam(m.s_);
    m.k1 = p[i++]; m.k2o = p[i++]; m.ninv = p[i++]; m.gamma_o = p[i++]; m.rho_dd_init = p[i++];
    if (xtal == XTAL_HCP) cOverA = p[i++];
  } else {
    i++;  // shear modulus (reported only)
    m.xm = p[i++]; m.gam_w0 = p[i++];
    m.h0 = p[i++]; m.tausi = p[i++]; m.taus0 = p[i++];
    m.xmprime = (kin == KIN_VOCE_NL) ? p[i++] : 1.0;
    m.xms = p[i++]; m.gamss0 = p[i++]; m.kappa0 = p[i++];
  }
  if (xtal == XTAL_FCC) slip_fcc(m);
  else if (xtal == XTAL_BCC) slip_bcc(m);
  else slip_hcp(m, cOverA);
  if (kin != KIN_KMBALD) {
    m.xmi = 1.0 / m.xm;
    m.pl_t_min = std::pow(1.0e-60, m.xm);
    m.pl_t_max = std::pow(1.0e45, m.xm);
    m.pl_max = std::exp((m.xmi - 1.0) * std::log(m.pl_t_max));
    // integer exponent 1/m - 1 (49 for the reference's Voce sets): evaluated by repeated squaring
    const double n = m.xmi - 1.0;
    m.pl_n = (n >= 1.0 && n <= 512.0 && n == std::floor(n)) ? (int)n : 0;
  }
  m.ln_ovf = std::log(1.0e45);
  m.gruneisen = p[i++];
  const double ec0 = p[i++];
  m.dtde = 1.0 / cvav;
  m.tK0 = -ec0 * m.dtde;
  m.nhist = iH_Gdot + m.nslip + 2;
  return std::string();
}

}  // namespace exab

// TEST INFRASTRUCTURE: compiles the per-point arithmetic of the CUDA material kernel
// (exaconstit_b200/csrc/material_point.hpp -- the very source that is inlined into k_model_setup) for the
// host, so the CPU test-suite can check it point by point against the oracle without a GPU.  The product never
// loads this library; the shipped path is the CUDA kernel only.
#include <cstring>
#include <string>

#define EXAB_POINT_STATS 1
#include "../../exaconstit_b200/csrc/material_host.hpp"

namespace exab { long g_point_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0}; }

using namespace exab;

namespace {
double adjugate(const double* J, double* adj) {
  const double J11 = J[0], J21 = J[1], J31 = J[2], J12 = J[3], J22 = J[4], J32 = J[5], J13 = J[6], J23 = J[7], J33 = J[8];
  adj[0] = (J22 * J33) - (J23 * J32); adj[1] = (J32 * J13) - (J12 * J33); adj[2] = (J12 * J23) - (J22 * J13);
  adj[3] = (J31 * J23) - (J21 * J33); adj[4] = (J11 * J33) - (J13 * J31); adj[5] = (J21 * J13) - (J11 * J23);
  adj[6] = (J21 * J32) - (J31 * J22); adj[7] = (J31 * J12) - (J11 * J32); adj[8] = (J11 * J22) - (J12 * J21);
  return J11 * adj[0] + J21 * adj[1] + J31 * adj[2];
}

template <int NSLIP, int KIN>
long run(const MatDev& m, long ne, double dt, const double* jac, const double* G, const double* velE, const double* s0,
         const double* h0, double* s1, double* h1, double* mg, long* nfev_sum, int layout) {
  const int nsv = m.nhist;
  long nfail = 0, nf = 0;
#pragma omp parallel for reduction(+ : nfail, nf)
  for (long p = 0; p < ne * 8; ++p) {
    const long e = p / 8;
    const int q = (int)(p % 8);
    double adj[9], L[3][3], d[3][3];
    const double idet = 1.0 / adjugate(jac + p * 9, adj);
    for (int i = 0; i < 3; ++i)
      for (int s = 0; s < 3; ++s) {
        double v = 0.0;
        for (int a = 0; a < 8; ++a) v += velE[e * 24 + i * 8 + a] * G[q * 24 + s * 8 + a];
        d[i][s] = v;
      }
    for (int i = 0; i < 3; ++i)
      for (int t = 0; t < 3; ++t) L[i][t] = (d[i][0] * adj[t] + d[i][1] * adj[3 + t] + d[i][2] * adj[6 + t]) * idet;
    double J[64];
    const int r = mat::update_point<NSLIP, KIN, 1>(m, dt, L, h0 + p * nsv, s0 + p * 6, h1 + p * nsv, s1 + p * 6, mg + p * 36, layout, J);
    if (r < 0) ++nfail;
    nf += r < 0 ? -r : r;
  }
  if (nfev_sum) *nfev_sum = nf;
  return nfail;
}
}  // namespace

extern "C" {
// returns the number of failed points, or -1 on a bad material description
long hostcheck_model_setup(int xtal, int kin, const double* props, int nprops, int force_pivot, int disable_powi, long ne,
                           double dt, const double* jac, const double* G, const double* velE, const double* s0,
                           const double* h0, double* s1, double* h1, double* mg, long* nfev_sum, int layout) {
  MatDev m;
  if (!build_material(m, xtal, kin, props, nprops).empty()) return -1;
  m.force_pivot = force_pivot;
  if (disable_powi) m.pl_n = 0;
  const bool km = kin == KIN_KMBALD;
  if (m.nslip == 12) {
    if (km) return run<12, 1>(m, ne, dt, jac, G, velE, s0, h0, s1, h1, mg, nfev_sum, layout);
    return run<12, 0>(m, ne, dt, jac, G, velE, s0, h0, s1, h1, mg, nfev_sum, layout);
  }
  if (!km) return -1;
  return run<24, 1>(m, ne, dt, jac, G, velE, s0, h0, s1, h1, mg, nfev_sum, layout);
}
// solver path counters since the last call: 0 trial evaluations, 1 unused, 2 Jacobian re-evaluations (failed
// solves only), 3 rejected trials, 4 pivoted-LU fallbacks in the Newton loop, 5 dogleg / Cauchy steps,
// 6 pivoted-LU fallbacks in the tangent
void hostcheck_stats(long* out8) {
  for (int i = 0; i < 8; ++i) { out8[i] = g_point_stats[i]; g_point_stats[i] = 0; }
}
// expands compact tangent records (layout 2) to the Voigt 6x6 layout, npts points, 36-double slots in and out
void hostcheck_compact_expand(long npts, const double* rec, double* K36) {
  for (long p = 0; p < npts; ++p) mat::compact_expand(rec + p * 36, K36 + p * 36);
}
int hostcheck_nhist(int xtal, int kin, const double* props, int nprops) {
  MatDev m;
  if (!build_material(m, xtal, kin, props, nprops).empty()) return -1;
  return m.nhist;
}
int hostcheck_pl_n(int xtal, int kin, const double* props, int nprops) {
  MatDev m;
  if (!build_material(m, xtal, kin, props, nprops).empty()) return -1;
  return m.pl_n;
}
}

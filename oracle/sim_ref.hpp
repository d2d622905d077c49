// ORACLE (test infrastructure only -- never linked into the product path).
//
// CPU restatement of the ExaConstit solution loop around the hot path, on an
// auto-generated Cartesian voxel mesh of p=1 hexes, so that the reference's golden
// volume-averaged stress histories (test/data/*_stress.txt) can be reproduced:
//   time loop                 src/mechanics_driver.cpp:837-907
//   SystemDriver::Solve/SolveInit/UpdateVelocity/UpdateModel
//                             src/system_driver.cpp:221-288,293-319,327-427,429-468
//   ExaNewtonSolver / ExaNewtonLSSolver   src/mechanics_solver.cpp:39-143,155-280
//   NonlinearMechOperator::Mult/Setup/GetGradient/GetUpdateBCsAction
//                             src/mechanics_operator.cpp:288-348,436-483
//   PA / EA gradient operator  src/mechanics_operator_ext.cpp:95-174,228-328
//   ExaCMechModel::ModelSetup  src/mechanics_ecmech.cpp:192-258
//   mfem::CGSolver (external; restated from MFEM's published algorithm)
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <vector>

#include "ecmech_port.hpp"
#include "fem_ref.hpp"

namespace orc {

using dvec = std::vector<double>;

struct BCSet {
  int step;                // 1-based step at which this set becomes active
  std::vector<int> ids;    // boundary attributes: 1 z_min 2 x_min 3 y_min 4 z_max 5 x_max 6 y_max
  std::vector<int> comps;  // component codes (src/BCData.cpp:27-117); negative = velocity-gradient BC
  dvec vals;               // 3 per id
  dvec vgrad;              // 9, row-major L (BCs.essential_vel_grad, src/option_parser.cpp:216-226); empty if unused
};

// Time.Auto (src/option_parser.cpp, src/system_driver.cpp:225-274, src/mechanics_driver.cpp:212,845-848)
struct AutoTime {
  bool on = false;
  double dt_start = 1.0, dt_min = 1.0, dt_scale = 0.25, t_final = 1.0;
};

struct SimConfig {
  int nx = 1, ny = 1, nz = 1;
  double len[3] = {1, 1, 1};
  int xtal = 0, kin = 0;
  dvec props;
  double temp_k = 298.0;
  std::vector<int> grain_ids;  // per element, 1-based
  dvec quats;                  // 4 per grain
  dvec dts;
  std::vector<BCSet> bcs;
  int assembly = 0;   // 0 PA, 1 EA
  int integ = 0;      // 0 full integration, 1 B-bar
  int nl_solver = 0;  // 0 NR, 1 NRLS
  double nr_rel = 5e-5, nr_abs = 5e-10;
  int nr_iter = 25;
  double kr_rel = 1e-7, kr_abs = 1e-27;
  int kr_iter = 1000;
  bool true_jacobi = false;  // false = reference behaviour (dinv never refreshed, quirk C.1)
  ecm::Options opt;
  int verbose = 0;
  AutoTime auto_time;
};

struct SimStats {
  long newton_iters = 0, pcg_iters = 0, model_setups = 0, grad_mults = 0;
  int failed_points = 0;
  double pcg_seconds = 0.0;  // wall time inside pcg() (bench.py's CPU legs split a step into its CG and non-CG parts)
};

// ExaCMechModel::ModelSetup for one batch of elements: begin->end copies, grad_calc,
// kernel_setup, getResponseECM, kernel_postprocessing (src/mechanics_ecmech.cpp:22-258).
inline int model_setup(const ecm::Material& mat, long ne, double dt, double temp_k, const double* jac,
                       const double* G, const double* velE, const double* stress0, const double* hist0,
                       double* stress1, double* hist1, double* ddsdde, bool transpose_tangent = true) {
  const int nsv = mat.nhist;
  const int ind_int_eng = nsv - 1, ind_vols = ind_int_eng - 1, ind_pl_work = ecm::iHistA_flowStr;
  int nfail = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : nfail)
  for (long e = 0; e < ne; ++e) {
    double vg[72];
    for (int x = 0; x < 72; ++x) vg[x] = 0.0;
    // grad_calc on this element (layout (i,t,q): vg[q*9 + t*3 + i])
    for (int q = 0; q < 8; ++q) {
      double adj[9];
      const double detJ = adjugate(&jac[(e * 8 + q) * 9], adj);
      const double c = 1.0 / detJ;
      for (int t = 0; t < 3; ++t)
        for (int s = 0; s < 3; ++s)
          for (int r = 0; r < 8; ++r)
            for (int i = 0; i < 3; ++i)
              vg[q * 9 + t * 3 + i] += velE[e * 24 + i * 8 + r] * G[q * 24 + s * 8 + r] * (c * adj[3 * s + t]);
    }
    for (int q = 0; q < 8; ++q) {
      const long p = e * 8 + q;
      const double* L = &vg[q * 9];
      auto vgrad = [&](int i, int t) { return L[t * 3 + i]; };
      double* sv = &hist1[p * nsv];
      for (int i = 0; i < nsv; ++i) sv[i] = hist0[p * nsv + i];  // StateVarsSetup
      double* sig = &stress1[p * 6];
      for (int i = 0; i < 6; ++i) sig[i] = stress0[p * 6 + i];  // StressSetup
      // kernel_setup (src/mechanics_ecmech.cpp:57-98)
      double w_vec[3], d_svec_p[7], vol_ratio[4], stress_svec_p[7], eng_int[1], sdd[2], mtan[36];
      double tempk = temp_k;
      eng_int[0] = sv[ind_int_eng];
      w_vec[0] = 0.5 * (vgrad(2, 1) - vgrad(1, 2));
      w_vec[1] = 0.5 * (vgrad(0, 2) - vgrad(2, 0));
      w_vec[2] = 0.5 * (vgrad(1, 0) - vgrad(0, 1));
      const double d_mean = -ecm::onethird * (vgrad(0, 0) + vgrad(1, 1) + vgrad(2, 2));
      d_svec_p[0] = vgrad(0, 0) + d_mean;
      d_svec_p[1] = vgrad(1, 1) + d_mean;
      d_svec_p[2] = vgrad(2, 2) + d_mean;
      d_svec_p[3] = 0.5 * (vgrad(2, 1) + vgrad(1, 2));
      d_svec_p[4] = 0.5 * (vgrad(2, 0) + vgrad(0, 2));
      d_svec_p[5] = 0.5 * (vgrad(1, 0) + vgrad(0, 1));
      d_svec_p[6] = -3.0 * d_mean;
      double d_vecd[5];
      ecm::svec_to_vecd(d_svec_p, d_vecd);
      const double dEff = ecm::vecd_Deff(d_vecd);
      vol_ratio[0] = sv[ind_vols];
      vol_ratio[1] = vol_ratio[0] * std::exp(d_svec_p[6] * dt);
      vol_ratio[3] = vol_ratio[1] - vol_ratio[0];
      vol_ratio[2] = vol_ratio[3] / (dt * 0.5 * (vol_ratio[0] + vol_ratio[1]));
      for (int i = 0; i < 6; ++i) stress_svec_p[i] = sig[i];
      const double stress_mean = -ecm::onethird * (sig[0] + sig[1] + sig[2]);
      stress_svec_p[0] += stress_mean;
      stress_svec_p[1] += stress_mean;
      stress_svec_p[2] += stress_mean;
      stress_svec_p[6] = stress_mean;
      // getResponseECM
      const int rc = ecm::get_response_sngl(mat, dt, d_svec_p, w_vec, vol_ratio, eng_int, stress_svec_p, sv,
                                            tempk, sdd, mtan);
      if (rc < 0) ++nfail;
      // kernel_postprocessing (src/mechanics_ecmech.cpp:128-169)
      sv[ind_vols] = vol_ratio[1];
      sv[ind_int_eng] = eng_int[0];
      if (dEff > ecm::idp_tiny_sqrt) sv[ind_pl_work] *= dEff * dt;
      else sv[ind_pl_work] = 0.0;
      sv[ind_pl_work] += hist0[p * nsv + ind_pl_work];
      const double sm = -stress_svec_p[6];
      for (int i = 0; i < 6; ++i) sig[i] = stress_svec_p[i];
      sig[0] += sm; sig[1] += sm; sig[2] += sm;
      // tangent: ExaCMech row-major -> transposed in place so that k[j*6+i] = K(i,j)
      double* K = &ddsdde[p * 36];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) K[(transpose_tangent ? j * 6 + i : i * 6 + j)] = mtan[i * 6 + j];
    }
  }
  return nfail;
}

class VoxelSim {
 public:
  SimConfig cfg;
  ecm::Material mat;
  long ne, nn, ndof;
  std::vector<int> e2n;
  dvec x_beg, x_end, x_ref, G, W;
  dvec stress0, stress1, hist0, hist1, matgrad;
  dvec jac, velE, c81, D81, ea, eds, dres;
  std::vector<char> ess;    // per true dof
  std::vector<char> ess_vg; // essential dofs driven by the velocity gradient
  double vgradL[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool has_vgrad = false;
  dvec ess_val;             // velocity value on essential dofs
  std::vector<std::vector<int>> n2e;  // node -> (elem*8+local) list, for race-free scatter
  double dt = 0.0;
  SimStats stats;
  dvec dinv;

  explicit VoxelSim(const SimConfig& c) : cfg(c) {
    ne = (long)c.nx * c.ny * c.nz;
    nn = (long)(c.nx + 1) * (c.ny + 1) * (c.nz + 1);
    ndof = 3 * nn;
    e2n.resize(ne * 8);
    x_beg.resize(ndof);
    voxel_mesh(c.nx, c.ny, c.nz, c.len[0], c.len[1], c.len[2], e2n.data(), x_beg.data());
    x_end = x_beg;
    x_ref = x_beg;
    G.resize(192);
    W.resize(8);
    hex8_dshape(G.data(), W.data());
    int rc = ecm::init_material(mat, c.xtal, c.kin, c.props.data(), (int)c.props.size());
    if (rc) std::fprintf(stderr, "orc: bad property vector (rc=%d)\n", rc);
    mat.opt = c.opt;
    const long npts = ne * 8;
    const int nsv = mat.nhist;
    stress0.assign(npts * 6, 0.0);
    stress1.assign(npts * 6, 0.0);
    hist0.assign(npts * nsv, 0.0);
    hist1.assign(npts * nsv, 0.0);
    matgrad.assign(npts * 36, 0.0);
    jac.resize(npts * 9);
    velE.resize(ne * 24);
    dres.resize(npts * 9);
    // history init: setStateVarData (src/mechanics_driver.cpp:1058-1154) then
    // init_state_vars (src/mechanics_ecmech.hpp:249-300)
    dvec hinit(nsv);
    ecm::hist_init(mat, hinit.data());
    for (long e = 0; e < ne; ++e) {
      const int g = c.grain_ids[e] - 1;
      for (int q = 0; q < 8; ++q) {
        double* h = &hist0[(e * 8 + q) * nsv];
        for (int i = 0; i < nsv; ++i) h[i] = hinit[i];
        for (int i = 0; i < 4; ++i) h[ecm::iHistLbQ + i] = c.quats[4 * g + i];
        h[nsv - 2] = 1.0;
        h[nsv - 1] = 0.0;
      }
    }
    ess.assign(ndof, 0);
    ess_vg.assign(ndof, 0);
    ess_val.assign(ndof, 0.0);
    n2e.resize(nn);
    for (long e = 0; e < ne; ++e)
      for (int a = 0; a < 8; ++a) n2e[e2n[e * 8 + a]].push_back((int)(e * 8 + a));
    dinv.assign(ndof, 1.0);
  }

  bool on_face(long n, int attr) const {
    const int px = cfg.nx + 1, py = cfg.ny + 1;
    const int i = (int)(n % px), j = (int)((n / px) % py), k = (int)(n / ((long)px * py));
    switch (attr) {
      case 1: return k == 0;
      case 2: return i == 0;
      case 3: return j == 0;
      case 4: return k == cfg.nz;
      case 5: return i == cfg.nx;
      case 6: return j == cfg.ny;
    }
    return false;
  }

  // UpdateEssBdr + the essential-dof list (src/system_driver.cpp:321-324)
  void set_bcs(const BCSet& b) {
    std::fill(ess.begin(), ess.end(), 0);
    std::fill(ess_vg.begin(), ess_vg.end(), 0);
    std::fill(ess_val.begin(), ess_val.end(), 0.0);
    has_vgrad = false;
    for (int i = 0; i < 9; ++i) vgradL[i] = (b.vgrad.size() == 9) ? b.vgrad[i] : 0.0;
    // ess_vel attributes first, then ess_vgrad ones (BCManager::updateBCData, src/BCManager.cpp:10-140): a dof
    // claimed by both kinds ends up velocity-gradient driven (UpdateVelocity applies that block last)
    for (int pass = 0; pass < 2; ++pass)
      for (size_t s = 0; s < b.ids.size(); ++s) {
        const bool vg = b.comps[s] < 0;
        if (vg != (pass == 1)) continue;
        const int code = std::abs(b.comps[s]);
        const bool cmp[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 1}};
        for (long n = 0; n < nn; ++n)
          if (on_face(n, b.ids[s]))
            for (int d = 0; d < 3; ++d)
              if (cmp[code][d]) {
                ess[d * nn + n] = 1;
                if (vg) { ess_vg[d * nn + n] = 1; has_vgrad = true; }
                else ess_val[d * nn + n] = b.vals[3 * s + d];
              }
      }
  }
  // UpdateVelocity (src/system_driver.cpp:327-427): velocity BCs overwrite their components; velocity-gradient
  // BCs get v = L (x - x_min) on the CURRENT node coordinates (origin = component-wise minimum of the mesh,
  // recomputed every call), i.e. a constant true strain rate
  void update_velocity(dvec& v) const {
    for (long i = 0; i < ndof; ++i)
      if (ess[i] && !ess_vg[i]) v[i] = ess_val[i];
    if (!has_vgrad) return;
    double org[3];
    for (int j = 0; j < 3; ++j) {
      double m = x_end[j * nn];
      for (long n = 1; n < nn; ++n) m = std::min(m, x_end[j * nn + n]);
      org[j] = m;
    }
    for (long n = 0; n < nn; ++n)
      for (int d = 0; d < 3; ++d)
        if (ess_vg[d * nn + n]) {
          double s = 0.0;
          for (int j = 0; j < 3; ++j) s += vgradL[3 * d + j] * (x_end[j * nn + n] - org[j]);
          v[d * nn + n] = s;
        }
  }

  void scatter(const dvec& yE, dvec& yL) const {
#pragma omp parallel for schedule(static)
    for (long n = 0; n < nn; ++n)
      for (int i = 0; i < 3; ++i) {
        double s = 0.0;
        for (int id : n2e[n]) s += yE[(long)(id / 8) * 24 + i * 8 + (id % 8)];
        yL[i * nn + n] = s;
      }
  }

  // NonlinearMechOperator::Setup<upd_crds> (src/mechanics_operator.cpp:311-348)
  void setup(const dvec& k, bool upd_crds) {
    if (upd_crds)
      {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < ndof; ++i) x_end[i] = k[i] * dt + x_beg[i];
      }
    dvec xE(ne * 24);
    gather(ne, nn, e2n.data(), x_end.data(), xE.data());
    jacobians(ne, G.data(), xE.data(), jac.data());
    gather(ne, nn, e2n.data(), k.data(), velE.data());
    const bool transpose = true;  // CPU path (src/mechanics_ecmech.cpp:155)
    stats.failed_points += model_setup(mat, ne, dt, cfg.temp_k, jac.data(), G.data(), velE.data(),
                                       stress0.data(), hist0.data(), stress1.data(), hist1.data(),
                                       matgrad.data(), transpose);
    ++stats.model_setups;
    if (cfg.integ == 1) { eds.resize(ne * 24); ic_assemble_eds(ne, jac.data(), W.data(), G.data(), eds.data()); }
  }

  // residual action: Hform->Mult (spec: MultVec, src/mechanics_operator_ext.cpp:176-202)
  void residual_action(dvec& y) {
    dvec yE(ne * 24, 0.0);
    if (cfg.integ == 1) {
      ic_addmult_pa(ne, jac.data(), W.data(), G.data(), eds.data(), stress1.data(), yE.data());
    } else {
      assemble_pa(ne, jac.data(), W.data(), stress1.data(), dres.data());
      addmult_pa(ne, G.data(), dres.data(), yE.data());
    }
    scatter(yE, y);
    for (long i = 0; i < ndof; ++i)
      if (ess[i]) y[i] = 0.0;
  }

  // NonlinearMechOperator::Mult (src/mechanics_operator.cpp:288-308)
  void mult(const dvec& k, dvec& y) {
    setup(k, true);
    residual_action(y);
  }

  // Hform->GetGradient: AssembleGradPA / AssembleEA (+ diagonal)
  void get_gradient() {
    // PA: the gradient is the plain operator also with B-bar integration (ICExaNLFIntegrator inherits
    // ExaNLFIntegrator::AssembleGradPA / AddMultGradPA, src/mechanics_integrators.hpp:107-110)
    if (cfg.assembly == 0) {
      c81.resize(ne * 8 * 81);
      D81.resize(ne * 8 * 81);
      transform_matgrad_4d(ne * 8, matgrad.data(), c81.data());
      assemble_grad_pa(ne, dt, jac.data(), W.data(), c81.data(), D81.data());
    } else {
      ea.assign(ne * 576, 0.0);
      if (cfg.integ == 1) ic_assemble_ea(ne, dt, jac.data(), W.data(), G.data(), eds.data(), matgrad.data(), ea.data());
      else assemble_ea(ne, dt, jac.data(), W.data(), G.data(), matgrad.data(), ea.data());
    }
    if (cfg.true_jacobi) {
      dvec dE(ne * 24, 0.0), d(ndof);
      if (cfg.assembly == 0 && cfg.integ == 0) assemble_grad_diag_pa(ne, dt, jac.data(), W.data(), G.data(), matgrad.data(), dE.data());
      else if (cfg.assembly == 0) ic_assemble_grad_diag_pa(ne, dt, jac.data(), W.data(), G.data(), eds.data(), matgrad.data(), dE.data());
      else ea_diag(ne, ea.data(), dE.data());
      scatter(dE, d);
      for (long i = 0; i < ndof; ++i) dinv[i] = ess[i] ? 1.0 : 1.0 / d[i];
    }
  }

  // gradient operator TMult<local_action> (src/mechanics_operator_ext.cpp:136-174,278-328)
  void grad_mult(const dvec& x, dvec& y, bool local_action = false) {
    dvec xm(x);
    if (!local_action)
      for (long i = 0; i < ndof; ++i)
        if (ess[i]) xm[i] = 0.0;
    dvec xE(ne * 24), yE(ne * 24, 0.0);
    gather(ne, nn, e2n.data(), xm.data(), xE.data());
    if (cfg.assembly == 0) addmult_grad_pa(ne, G.data(), D81.data(), xE.data(), yE.data());
    else ea_mult(ne, ea.data(), xE.data(), yE.data());
    scatter(yE, y);
    if (!local_action)
      for (long i = 0; i < ndof; ++i)
        if (ess[i]) y[i] = 0.0;
    ++stats.grad_mults;
  }

  static double dot(const dvec& a, const dvec& b) {
    double s = 0.0;
    const long n = (long)a.size();
#pragma omp parallel for reduction(+ : s)
    for (long i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
  }

  // mfem::CGSolver::Mult with the MechOperatorJacobiSmoother as preconditioner,
  // iterative_mode = false (src/mechanics_solver.cpp:72, src/mechanics_operator_ext.cpp:37-55)
  int pcg(const dvec& b, dvec& x) {
    const auto t0 = std::chrono::steady_clock::now();
    const int it = pcg_body(b, x);
    stats.pcg_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return it;
  }
  int pcg_body(const dvec& b, dvec& x) {
    const long n = ndof;
    dvec r(b), z(n), d(n);
    std::fill(x.begin(), x.end(), 0.0);
    auto prec = [&](const dvec& in, dvec& out) {
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n; ++i) out[i] = dinv[i] * in[i];
    };
    prec(r, z);
    d = z;
    double nom = dot(d, r);
    const double r0 = std::max(nom * cfg.kr_rel * cfg.kr_rel, cfg.kr_abs * cfg.kr_abs);
    if (nom <= r0) return 0;
    grad_mult(d, z);
    double den = dot(z, d);
    if (den <= 0.0) {
      if (den == 0.0) return 0;
    }
    int i = 1;
    for (;;) {
      const double alpha = nom / den;
#pragma omp parallel for schedule(static)
      for (long j = 0; j < n; ++j) { x[j] += alpha * d[j]; r[j] -= alpha * z[j]; }
      prec(r, z);
      const double betanom = dot(r, z);
      if (betanom <= r0) break;
      if (++i > cfg.kr_iter) break;
      const double beta = betanom / nom;
#pragma omp parallel for schedule(static)
      for (long j = 0; j < n; ++j) d[j] = z[j] + beta * d[j];
      grad_mult(d, z);
      den = dot(d, z);
      if (den <= 0.0 && den == 0.0) break;
      nom = betanom;
    }
    stats.pcg_iters += std::min(i, cfg.kr_iter);
    return i;
  }

  // ExaNewtonSolver::Mult / ExaNewtonLSSolver::Mult; returns converged flag
  bool newton(dvec& x, int* iters) {
    dvec r(ndof), c(ndof), xp(ndof);
    mult(x, r);
    double norm = std::sqrt(dot(r, r));
    const double norm_max = std::max(cfg.nr_rel * norm, cfg.nr_abs);
    double scale = 1.0;
    int it = 0;
    bool converged = false;
    for (it = 0; true; ++it) {
      if (cfg.verbose) std::printf("  Newton iteration %2d : ||r|| = %.6e\n", it, norm);
      if (norm <= norm_max) { converged = true; break; }
      if (it >= cfg.nr_iter) { converged = false; break; }
      get_gradient();
      pcg(r, c);
      if (cfg.nl_solver == 1) {
        xp = x;
        for (long i = 0; i < ndof; ++i) x[i] = xp[i] - c[i];
        mult(x, r);
        const double q1 = norm, q3 = std::sqrt(dot(r, r));
        for (long i = 0; i < ndof; ++i) x[i] = xp[i] - 0.5 * c[i];
        mult(x, r);
        const double q2 = std::sqrt(dot(r, r));
        const double eps = (3.0 * q1 - 4.0 * q2 + q3) / (4.0 * (q1 - 2.0 * q2 + q3));
        if ((q1 - 2.0 * q2 + q3) > 0 && eps > 0 && eps < 1) scale = eps;
        else if (q3 < q1) scale = 1.0;
        else scale = 0.05;
        x = xp;
      }
      for (long i = 0; i < ndof; ++i) x[i] -= scale * c[i];
      mult(x, r);
      const double norm_prev = norm;
      norm = std::sqrt(dot(r, r));
      if (cfg.nl_solver == 0) scale = (norm / norm_prev > 0.5) ? 0.5 : 1.0;
    }
    stats.newton_iters += it;
    *iters = it;
    return converged;
  }

  // SystemDriver::SolveInit + GetUpdateBCsAction (src/system_driver.cpp:293-319,
  // src/mechanics_operator.cpp:446-483)
  void solve_init(const dvec& xprev, dvec& x) {
    dvec deltaF(ndof, 0.0), b(ndof, 0.0), resid(ndof);
    for (long i = 0; i < ndof; ++i)
      if (ess[i]) deltaF[i] = x[i] - xprev[i];
    setup(xprev, false);
    get_gradient();
    grad_mult(deltaF, b, true);
    residual_action(resid);
    for (long i = 0; i < ndof; ++i) {
      if (ess[i]) b[i] = 0.0;
      b[i] += resid[i];
    }
    dvec sol(ndof, 0.0);
    pcg(b, sol);
    for (long i = 0; i < ndof; ++i) x[i] = -sol[i] + xprev[i];
  }

  void vol_avg(const dvec& qf, int vdim, double* out, bool avg) {
    double vol;
    vol_sum(ne, vdim, jac.data(), W.data(), qf.data(), out, &vol);
    if (avg) for (int c = 0; c < vdim; ++c) out[c] /= vol;
  }

  // main time loop; avg_stress: nsteps x 6.  With Time.Auto the step count is ceil(t_final / dt_min) at most
  // (src/mechanics_driver.cpp:212) and the loop stops at the last step; out_dts (may be null) receives the step
  // sizes actually taken, *out_nsteps their number.
  int run(double* avg_stress, double* extra /* nsteps x 16 or null */, int* iters /* nsteps x 2 or null */,
          double* out_dts = nullptr, int* out_nsteps = nullptr) {
    const AutoTime& at = cfg.auto_time;
    const int nsteps = at.on ? (int)std::ceil(at.t_final / at.dt_min) : (int)cfg.dts.size();
    begin();
    if (out_nsteps) *out_nsteps = 0;
    for (int ti = 1; ti <= nsteps; ++ti) {
      bool last_step = false;
      const int rc = step(at.on ? 0.0 : cfg.dts[ti - 1], &avg_stress[(ti - 1) * 6], extra ? &extra[(ti - 1) * 16] : nullptr,
                          iters ? &iters[(ti - 1) * 2] : nullptr, out_dts ? &out_dts[ti - 1] : nullptr, &last_step);
      if (rc) return ti;
      if (out_nsteps) *out_nsteps = ti;
      if (last_step) break;
    }
    return 0;
  }

  // time-loop state for step-at-a-time driving (run() above, and the bench's CPU-baseline leg, which times single steps)
  dvec v_cur, v_prev;
  double t_now = 0.0, dt_class = 0.0;
  int ti_now = 0;
  void begin() {
    v_cur.assign(ndof, 0.0);
    v_prev.assign(ndof, 0.0);
    t_now = 0.0;
    dt_class = cfg.auto_time.dt_start;
    ti_now = 0;
  }

  // one pass of the time loop (src/mechanics_driver.cpp:837-907); dt_fixed is ignored with Time.Auto.  Returns 0, or 1 if
  // the Newton solve failed.
  int step(double dt_fixed, double* avg_stress /*6*/, double* ex /*16 or null*/, int* iters /*2 or null*/,
           double* out_dt /*or null*/, bool* last /*or null*/) {
    dvec& v = v_cur;
    dvec& vprev = v_prev;
    const AutoTime& at = cfg.auto_time;
    double& t = t_now;
    const int ti = ++ti_now;
    {
      dt = at.on ? std::min(dt_class, at.t_final - t) : dt_fixed;
      t += dt;
      bool last_step = at.on && std::fabs(t - at.t_final) <= std::fabs(1e-3 * dt);
      const long pcg0 = stats.pcg_iters;
      for (const BCSet& b : cfg.bcs)
        if (b.step == ti) {
          vprev = v;
          set_bcs(b);
          update_velocity(v);
          solve_init(vprev, v);
        }
      update_velocity(v);
      int nit = 0;
      bool ok;
      if (at.on) {
        // SystemDriver::Solve, auto_time branch (src/system_driver.cpp:225-274)
        if (last_step) dt_class = dt;
        const double dt_old = dt_class;
        const dvec vsave(v);
        ok = newton(v, &nit);
        if (!ok) {
          for (int retry = 0; !ok && retry < 2; ++retry) {
            v = vsave;
            dt_class *= at.dt_scale;
            if (dt_class < at.dt_min) dt_class = at.dt_min;
            dt = dt_class;
            ok = newton(v, &nit);
          }
          t = t - dt_old + dt_class;
          last_step = std::fabs(t - at.t_final) <= std::fabs(1e-3 * dt);
        }
        const double factor = ((double)cfg.nr_iter * at.dt_scale) / (double)nit;
        if (out_dt) *out_dt = dt;
        dt_class *= factor;
        if (dt_class < at.dt_min) dt_class = at.dt_min;
      } else {
        ok = newton(v, &nit);
        if (out_dt) *out_dt = dt;
      }
      if (cfg.verbose) std::printf("step %d: t %.6f dt %.6f newton its %d converged %d\n", ti, t, dt, nit, (int)ok);
      if (!ok) return 1;
      // UpdateModel: swap begin/end, then averages over the end-of-step (current) mesh
      stress0.swap(stress1);
      hist0.swap(hist1);
      x_beg = x_end;
      vol_avg(stress0, 6, avg_stress, true);
      if (ex) {
        dvec tmp(mat.nhist);
        vol_avg(hist0, mat.nhist, tmp.data(), false);
        ex[0] = tmp[ecm::iHistA_flowStr];
        // D^p (calcDpMat, src/mechanics_ecmech.hpp:303-357), volume averaged, Voigt order.  The reference
        // evaluates it from matVars1 AFTER the begin/end swap (src/mechanics_ecmech.hpp:308 called at
        // src/system_driver.cpp:526), i.e. from the previous step's slip rates and orientations: reproduced.
        const long npts = ne * 8;
        dvec dp(npts * 6);
        for (long p = 0; p < npts; ++p) {
          const double* h = &hist1[p * mat.nhist];
          double dphat[5] = {0, 0, 0, 0, 0}, C[9], R5[5][5], dsm[5], s6[6];
          for (int a = 0; a < mat.nslip; ++a)
            for (int i = 0; i < 5; ++i) dphat[i] += mat.P[a][i] * h[ecm::iHistLbGdot + a];
          ecm::quat_to_tensor(&h[ecm::iHistLbQ], C);
          ecm::rot_mat_vecd(C, R5);
          for (int i = 0; i < 5; ++i) { dsm[i] = 0; for (int j = 0; j < 5; ++j) dsm[i] += R5[i][j] * dphat[j]; }
          ecm::vecd_to_svec(dsm, s6);
          for (int i = 0; i < 6; ++i) dp[p * 6 + i] = s6[i];
        }
        vol_avg(dp, 6, &ex[1], true);
        // <F>: CalculateDeformationGradient (src/mechanics_operator.cpp:393-427): gradient of the current coordinates
        // with respect to the REFERENCE mesh at its quadrature points, F(i,t) at [t*3+i]; averaged over the
        // current mesh like the other quantities (src/system_driver.cpp:496-517) -> ex[7..15]
        {
          dvec xrE(ne * 24), xcE(ne * 24), jref(npts * 9), F(npts * 9, 0.0);
          gather(ne, nn, e2n.data(), x_ref.data(), xrE.data());
          jacobians(ne, G.data(), xrE.data(), jref.data());
          gather(ne, nn, e2n.data(), x_beg.data(), xcE.data());
          grad_calc(ne, jref.data(), G.data(), xcE.data(), F.data());
          vol_avg(F, 9, &ex[7], true);
        }
      }
      if (iters) { iters[0] = nit; iters[1] = (int)(stats.pcg_iters - pcg0); }
      if (last) *last = last_step;
    }
    return 0;
  }
};

}  // namespace orc

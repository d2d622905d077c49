// ORACLE (test infrastructure only -- never linked into the product path).
// extern "C" surface of the CPU restatement so tests/ and bench.py's cpu_baseline leg
// can drive it through ctypes.  See fem_ref.hpp / ecmech_port.hpp / sim_ref.hpp for the
// reference file:line each routine follows.
#include <chrono>
#include <cstring>

#include "sim_ref.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// bench.py's CPU legs: launchers such as torch.distributed.run export OMP_NUM_THREADS=1; the baseline must use every
// host thread it is allowed to, so the count is set explicitly.
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

void orc_hex8_dshape(double* G, double* W) { hex8_dshape(G, W); }
void orc_voxel_mesh(int nx, int ny, int nz, const double* len, int* e2n, double* coords) {
  voxel_mesh(nx, ny, nz, len[0], len[1], len[2], e2n, coords);
}
void orc_gather(long ne, long nn, const int* e2n, const double* xL, double* xE) { gather(ne, nn, e2n, xL, xE); }
void orc_scatter_add(long ne, long nn, const int* e2n, const double* yE, double* yL) {
  scatter_add(ne, nn, e2n, yE, yL);
}
void orc_jacobians(long ne, const double* G, const double* xE, double* jac) { jacobians(ne, G, xE, jac); }
void orc_grad_calc(long ne, const double* jac, const double* G, const double* xE, double* out) {
  grad_calc(ne, jac, G, xE, out);
}
void orc_transform_matgrad_4d(long npts, const double* k36, double* c81) { transform_matgrad_4d(npts, k36, c81); }
void orc_assemble_pa(long ne, const double* jac, const double* W, const double* stress, double* d) {
  assemble_pa(ne, jac, W, stress, d);
}
void orc_addmult_pa(long ne, const double* G, const double* d, double* yE) { addmult_pa(ne, G, d, yE); }
void orc_assemble_grad_pa(long ne, double dt, const double* jac, const double* W, const double* c81, double* D) {
  assemble_grad_pa(ne, dt, jac, W, c81, D);
}
void orc_addmult_grad_pa(long ne, const double* G, const double* D, const double* xE, double* yE) {
  addmult_grad_pa(ne, G, D, xE, yE);
}
void orc_assemble_grad_diag_pa(long ne, double dt, const double* jac, const double* W, const double* G,
                               const double* k36, double* dE) {
  assemble_grad_diag_pa(ne, dt, jac, W, G, k36, dE);
}
void orc_assemble_ea(long ne, double dt, const double* jac, const double* W, const double* G, const double* k36,
                     double* ea) {
  assemble_ea(ne, dt, jac, W, G, k36, ea);
}
void orc_ea_mult(long ne, const double* ea, const double* xE, double* yE) { ea_mult(ne, ea, xE, yE); }
void orc_ea_diag(long ne, const double* ea, double* dE) { ea_diag(ne, ea, dE); }
void orc_ic_assemble_eds(long ne, const double* jac, const double* W, const double* G, double* eds) {
  ic_assemble_eds(ne, jac, W, G, eds);
}
void orc_ic_addmult_pa(long ne, const double* jac, const double* W, const double* G, const double* eds,
                       const double* stress, double* yE) {
  ic_addmult_pa(ne, jac, W, G, eds, stress, yE);
}
void orc_ic_assemble_ea(long ne, double dt, const double* jac, const double* W, const double* G, const double* eds,
                        const double* k36, double* ea) {
  ic_assemble_ea(ne, dt, jac, W, G, eds, k36, ea);
}
void orc_ic_assemble_grad_diag_pa(long ne, double dt, const double* jac, const double* W, const double* G,
                                  const double* eds, const double* k36, double* dE) {
  ic_assemble_grad_diag_pa(ne, dt, jac, W, G, eds, k36, dE);
}
void orc_vol_sum(long ne, int vdim, const double* jac, const double* W, const double* qf, double* sums, double* vol) {
  vol_sum(ne, vdim, jac, W, qf, sums, vol);
}

static void set_opts(ecm::Options& o, const double* opts) {
  if (!opts) return;
  o.kirchhoff_rss = opts[0] != 0.0;
  o.eos_temperature = opts[1] != 0.0;
  o.hard_lag = opts[2] != 0.0;
  o.av_power = opts[3];
  o.slip_stretch_terms = opts[4] != 0.0;
  o.eos_mu_form = opts[5] != 0.0;
  o.vol_convect = opts[6] != 0.0;
}

int orc_nhist(int xtal, int kin) { return ecm::iHistLbGdot + (xtal == ecm::XTAL_HCP ? 24 : 12) + 2; }

int orc_hist_init(int xtal, int kin, const double* props, int nprops, double* h) {
  ecm::Material m;
  int rc = ecm::init_material(m, xtal, kin, props, nprops);
  if (rc) return rc;
  ecm::hist_init(m, h);
  return 0;
}

// getResponseECM over npts points with the strides of src/mechanics_ecmech.hpp:143-159
int orc_get_response(int xtal, int kin, const double* props, int nprops, const double* opts, double dt, long npts,
                     const double* d_svec_p, const double* w_vec, const double* vol_ratio, double* eng_int,
                     double* stress_svec_p, double* hist, double* tempk, double* sdd, double* mtan) {
  ecm::Material m;
  int rc = ecm::init_material(m, xtal, kin, props, nprops);
  if (rc) return -1000 - rc;
  set_opts(m.opt, opts);
  int nfail = 0;
#pragma omp parallel for reduction(+ : nfail)
  for (long p = 0; p < npts; ++p) {
    int r = ecm::get_response_sngl(m, dt, &d_svec_p[7 * p], &w_vec[3 * p], &vol_ratio[4 * p], &eng_int[p],
                                   &stress_svec_p[7 * p], &hist[m.nhist * p], tempk[p], &sdd[2 * p],
                                   mtan ? &mtan[36 * p] : nullptr);
    if (r < 0) ++nfail;
  }
  return nfail;
}

// Residual + Jacobian of the local 8x8 problem at a given scaled x, for derivative checks.
int orc_local_problem(int xtal, int kin, const double* props, int nprops, const double* opts, double dt,
                      const double* d_svec_p, const double* w_vec, double vnew, const double* hist, double tK,
                      const double* x, double* R, double* J) {
  ecm::Material m;
  int rc = ecm::init_material(m, xtal, kin, props, nprops);
  if (rc) return rc;
  set_opts(m.opt, opts);
  ecm::UpdateProblem prob(m);
  ecm::svec_to_vecd(d_svec_p, prob.d_sm);
  for (int k = 0; k < 3; ++k) prob.w_sm[k] = w_vec[k];
  for (int i = 0; i < 5; ++i) prob.e_n[i] = hist[ecm::iHistLbE + i];
  for (int i = 0; i < 4; ++i) prob.q_n[i] = hist[ecm::iHistLbQ + i];
  double h = hist[ecm::iHistLbH];
  ecm::kin_get_vals(m, tK, &h, prob.kv);
  prob.dt = dt; prob.dt_ri = 1.0 / dt; prob.detV = vnew; prob.detVi = 1.0 / vnew;
  prob.a_V_ri = 1.0 / std::cbrt(vnew); prob.tK = tK;
  prob.epsdot_scale_inv = 1.0; prob.rotincr_scale_inv = prob.dt_ri;
  prob.eval(x, R, J);
  return 0;
}

// ExaCMechModel::ModelSetup on a batch of elements (src/mechanics_ecmech.cpp:192-258)
int orc_model_setup(int xtal, int kin, const double* props, int nprops, const double* opts, long ne, double dt,
                    double temp_k, const double* jac, const double* G, const double* velE, const double* stress0,
                    const double* hist0, double* stress1, double* hist1, double* ddsdde) {
  ecm::Material m;
  int rc = ecm::init_material(m, xtal, kin, props, nprops);
  if (rc) return -1000 - rc;
  set_opts(m.opt, opts);
  return model_setup(m, ne, dt, temp_k, jac, G, velE, stress0, hist0, stress1, hist1, ddsdde, true);
}

static int make_config(SimConfig& c, int nx, int ny, int nz, const double* len, int xtal, int kin, const double* props, int nprops,
                       double temp_k, const int* grain_ids, const double* quats, int ngrains, const double* dts,
                       int nsteps, int nbc, const int* bc_steps, const int* bc_counts, const int* bc_ids,
                       const int* bc_comps, const double* bc_vals, const double* bc_vgrads, int assembly, int integ,
                       int nl_solver, const double* nr, const double* kr, int true_jacobi, const double* opts, int verbose,
                       const double* auto_time) {
  c.nx = nx; c.ny = ny; c.nz = nz;
  for (int i = 0; i < 3; ++i) c.len[i] = len[i];
  c.xtal = xtal; c.kin = kin;
  c.props.assign(props, props + nprops);
  c.temp_k = temp_k;
  c.grain_ids.assign(grain_ids, grain_ids + (long)nx * ny * nz);
  c.quats.assign(quats, quats + 4 * (long)ngrains);
  c.dts.assign(dts, dts + nsteps);
  int off = 0;
  for (int s = 0; s < nbc; ++s) {
    BCSet b;
    b.step = bc_steps[s];
    for (int i = 0; i < bc_counts[s]; ++i) {
      b.ids.push_back(bc_ids[off + i]);
      b.comps.push_back(bc_comps[off + i]);
      for (int d = 0; d < 3; ++d) b.vals.push_back(bc_vals[3 * (off + i) + d]);
    }
    if (bc_vgrads) b.vgrad.assign(bc_vgrads + 9 * s, bc_vgrads + 9 * s + 9);
    off += bc_counts[s];
    c.bcs.push_back(b);
  }
  c.assembly = assembly; c.integ = integ; c.nl_solver = nl_solver;
  c.nr_rel = nr[0]; c.nr_abs = nr[1]; c.nr_iter = (int)nr[2];
  c.kr_rel = kr[0]; c.kr_abs = kr[1]; c.kr_iter = (int)kr[2];
  c.true_jacobi = true_jacobi != 0;
  set_opts(c.opt, opts);
  c.verbose = verbose;
  if (auto_time && auto_time[0] != 0.0) {
    c.auto_time.on = true;
    c.auto_time.dt_start = auto_time[1]; c.auto_time.dt_min = auto_time[2];
    c.auto_time.dt_scale = auto_time[3]; c.auto_time.t_final = auto_time[4];
    if (nsteps < (int)std::ceil(c.auto_time.t_final / c.auto_time.dt_min)) return -1;
  }
  return 0;
}

// Full quasi-static simulation on a voxel mesh.  Returns 0 or the failing step.
//   bc_* : nbc sets; set s has bc_counts[s] (id, comp, 3 vals) entries, concatenated; comp < 0 = velocity-gradient
//          BC; bc_vgrads (may be null): 9 per set, row-major L
//   nr = {rel, abs, iters}, kr = {rel, abs, iters}
//   auto_time (may be null): {on, dt_start, dt_min, dt_scale, t_final}; then nsteps must be >= ceil(t_final/dt_min)
//   out_stress nsteps x 6; out_extra nsteps x 16 (may be null); out_iters nsteps x 2 (may be null)
//   out_stats = {newton_iters, pcg_iters, model_setups, grad_mults, failed_points, seconds, steps taken}
//   out_dts (may be null): step sizes taken
int orc_sim_run2(int nx, int ny, int nz, const double* len, int xtal, int kin, const double* props, int nprops,
                 double temp_k, const int* grain_ids, const double* quats, int ngrains, const double* dts,
                 int nsteps, int nbc, const int* bc_steps, const int* bc_counts, const int* bc_ids,
                 const int* bc_comps, const double* bc_vals, const double* bc_vgrads, int assembly, int integ,
                 int nl_solver, const double* nr, const double* kr, int true_jacobi, const double* opts, int verbose,
                 const double* auto_time, double* out_stress, double* out_extra, int* out_iters, double* out_stats,
                 double* out_hist /* final hist0, may be null */, double* out_stress_qp /* final stress0 */,
                 double* out_dts) {
  SimConfig c;
  if (make_config(c, nx, ny, nz, len, xtal, kin, props, nprops, temp_k, grain_ids, quats, ngrains, dts, nsteps, nbc, bc_steps,
                  bc_counts, bc_ids, bc_comps, bc_vals, bc_vgrads, assembly, integ, nl_solver, nr, kr, true_jacobi, opts, verbose,
                  auto_time))
    return -1;
  VoxelSim sim(c);
  auto t0 = std::chrono::steady_clock::now();
  int taken = 0;
  int rc = sim.run(out_stress, out_extra, out_iters, out_dts, &taken);
  auto t1 = std::chrono::steady_clock::now();
  if (out_stats) {
    out_stats[0] = (double)sim.stats.newton_iters;
    out_stats[1] = (double)sim.stats.pcg_iters;
    out_stats[2] = (double)sim.stats.model_setups;
    out_stats[3] = (double)sim.stats.grad_mults;
    out_stats[4] = (double)sim.stats.failed_points;
    out_stats[5] = std::chrono::duration<double>(t1 - t0).count();
    out_stats[6] = (double)taken;
  }
  if (out_hist) std::memcpy(out_hist, sim.hist0.data(), sim.hist0.size() * sizeof(double));
  if (out_stress_qp) std::memcpy(out_stress_qp, sim.stress0.data(), sim.stress0.size() * sizeof(double));
  return rc;
}

// Step-at-a-time driving of the same simulation (bench.py's CPU legs time single steps of it).
void* orc_sim_create(int nx, int ny, int nz, const double* len, int xtal, int kin, const double* props, int nprops,
                     double temp_k, const int* grain_ids, const double* quats, int ngrains, int nbc, const int* bc_steps,
                     const int* bc_counts, const int* bc_ids, const int* bc_comps, const double* bc_vals,
                     const double* bc_vgrads, int assembly, int integ, int nl_solver, const double* nr, const double* kr,
                     int true_jacobi, const double* opts, int verbose) {
  SimConfig c;
  if (make_config(c, nx, ny, nz, len, xtal, kin, props, nprops, temp_k, grain_ids, quats, ngrains, nullptr, 0, nbc, bc_steps,
                  bc_counts, bc_ids, bc_comps, bc_vals, bc_vgrads, assembly, integ, nl_solver, nr, kr, true_jacobi, opts, verbose,
                  nullptr))
    return nullptr;
  VoxelSim* sim = new VoxelSim(c);
  sim->begin();
  return sim;
}
// out[12] = {newton iters, pcg iters, model setups, grad mults (all of this step), seconds, avg stress[6], seconds inside
// the CG solves}; returns 0 / 1 (Newton failed)
int orc_sim_step(void* h, double dt, double* out) {
  VoxelSim* sim = static_cast<VoxelSim*>(h);
  const SimStats s0 = sim->stats;
  int iters[2] = {0, 0};
  auto t0 = std::chrono::steady_clock::now();
  const int rc = sim->step(dt, &out[5], nullptr, iters, nullptr, nullptr);
  auto t1 = std::chrono::steady_clock::now();
  out[0] = (double)(sim->stats.newton_iters - s0.newton_iters);
  out[1] = (double)(sim->stats.pcg_iters - s0.pcg_iters);
  out[2] = (double)(sim->stats.model_setups - s0.model_setups);
  out[3] = (double)(sim->stats.grad_mults - s0.grad_mults);
  out[4] = std::chrono::duration<double>(t1 - t0).count();
  out[11] = sim->stats.pcg_seconds - s0.pcg_seconds;
  return rc;
}
void orc_sim_destroy(void* h) { delete static_cast<VoxelSim*>(h); }

}  // extern "C"

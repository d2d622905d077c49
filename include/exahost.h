/* exahost -- C entry points of the C++ host layer (exaconstit_b200/csrc/host_sim.cu) that mirrors the
 * reference's SystemDriver / NonlinearMechOperator / ExaNewtonSolver / CGSolver classes on top of the
 * exab200 kernels, for a z-slab of an auto-generated voxel mesh.  These are what bench.py and the
 * system-level parity tests drive; a reference-side build would instead bind exab200.h directly from
 * its own classes.
 */
#ifndef EXAHOST_H
#define EXAHOST_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct exahost_sim exahost_sim;

typedef struct {
  /* mesh: Mesh.Auto (src/mechanics_driver.cpp:247-253): global nx*ny*nz_total hexes, this rank owns
   * element layers [z0, z0 + nz_local) */
  int nx, ny, nz_local, z0, nz_total;
  double length[3];
  /* Model.ExaCMech + Properties */
  int xtal, slip, nprops;
  const double* props;
  double temp_k;
  const int* grain_ids; /* nx*ny*nz_local, 1-based, x fastest (grains.txt order) */
  const double* quats;  /* 4 per grain */
  int ngrains;
  /* Solvers */
  int assembly, integ, nl_solver; /* PA|EA, FULL|BBAR, NR|NRLS */
  double newton_rel_tol, newton_abs_tol;
  int newton_iter;
  double krylov_rel_tol, krylov_abs_tol;
  int krylov_iter;
  int true_jacobi; /* 0 = reference behaviour (smoother never refreshed => identity), 1 = Jacobi */
  /* parallel layout */
  int rank, nranks, device;
  const void* nccl_id; /* 128-byte ncclUniqueId shared by all ranks (NULL when nranks == 1) */
  int verbose;
} exahost_config;

/* z-slab element partition of an nx*ny*nz_total voxel mesh over `nranks` ranks and the ownership / interface-plane
 * index sets the exchanges work with (host arithmetic only; exahost_create validates its config against it and the
 * exchange kernels index with the same numbers):
 *   out[0] z0 (first element layer)   out[1] nz_local           out[2] local nodes (both interface planes included)
 *   out[3] nodes per plane            out[4] uniquely-owned nodes (all but the top plane; the last rank owns it too)
 *   out[5] node offset of the bottom interface plane in the local L-vector components, out[6] of the top one
 *   out[7] local elements             out[8] has lower neighbour  out[9] has upper neighbour
 *   out[10] leading warp tiles (4 elements) of the bottom layer   out[11] first warp tile of the top layer
 * returns 0, or 1 for an impossible layout (fewer layers than ranks, nranks outside 1..8). */
int exahost_slab_layout(int nx, int ny, int nz_total, int rank, int nranks, long* out12);
const char* exahost_last_error(void);
int exahost_nccl_unique_id(void* out128);
int exahost_create(const exahost_config* cfg, exahost_sim** out);
void exahost_destroy(exahost_sim* sim);
int exahost_set_bcs(exahost_sim* sim, const unsigned char* mask_per_node, const double* h_ess_val_L);
/* velocity-gradient ("constant strain rate") BCs: BCs.essential_comps < 0 + BCs.essential_vel_grad
 * (src/system_driver.cpp:346-426); mask_vgrad_per_node bit i = component i driven by L (row-major 3x3) */
int exahost_set_vgrad(exahost_sim* sim, const unsigned char* mask_vgrad_per_node, const double* L9);
int exahost_step(exahost_sim* sim, double dt, int bc_changed, const double* h_ess_val_in, double* h_vel_out,
                 double* out16);
/* Time.Auto (src/system_driver.cpp:225-274): ctl6 = {dt_class, t, dt_min, dt_scale, t_final, last_step} in/out */
int exahost_step_auto(exahost_sim* sim, double* ctl6, int bc_changed, const double* h_ess_val_in, double* h_vel_out,
                      double* out16);
int exahost_kernel_timing(exahost_sim* sim, int enable);
int exahost_kernel_time(exahost_sim* sim, int which, double* total_ms, long* count, int reset);
int exahost_set_tuning(exahost_sim* sim, int ctas_per_sm, int variant);
int exahost_extra_avgs(exahost_sim* sim, double* out16);
int exahost_get(exahost_sim* sim, int which, double* h_out);
long exahost_counter(exahost_sim* sim, int which);
int exahost_comm_handle(exahost_sim* sim, void* out64);
int exahost_set_peers(exahost_sim* sim, const void* handles_nranks_x64);
void* exahost_stream(exahost_sim* sim);
void* exahost_ctx(exahost_sim* sim);

#ifdef __cplusplus
}
#endif
#endif

/* exab200 -- C ABI of the B200-native ExaConstit hot path.
 *
 * The reference has no C ABI or plugin loader: its material models and integrators are C++
 * virtuals compiled in (src/mechanics_operator.cpp:66-210).  Each entry point below is what a
 * reference-side shim class binds in place of the virtual it replaces; the reference file:line
 * is cited per function and the shim code is shown in INTEGRATION.md.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer to fp64 (or int32 where stated) owned by the
 *    caller; nothing is allocated or freed across the ABI except the opaque context;
 *  - layouts are the reference's (src/mechanics_integrators.cpp:217-218,410-414,580-585):
 *      E-vector  X(a,i,e) -> x[e*24 + i*8 + a]        (a: NATIVE hex vertex order)
 *      L-vector  byNODES  -> x[i*nnodes + node]
 *      jacobian  J(i,s,q,e) -> jac[((e*8+q)*3 + s)*3 + i]
 *      quadrature functions (vdim,q,e) -> qf[(e*8+q)*vdim + c]; stress Voigt 11,22,33,23,13,12
 *      matGrad   K(i,j,q,e) -> k[(e*8+q)*36 + j*6 + i] = d sigma_i / d eps_j (engineering shear)
 *      ea_data   E(r,c,e) -> ea[e*576 + c*24 + r]
 *  - p = 1 hexahedra, 8 nodes, 8 Gauss points (IntRules.Get(CUBE, 3), x fastest);
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *  - return value 0 = success; != 0 = error (exab200_last_error() gives the text).  The
 *    reference aborts the job on errors (MFEM_ABORT); the shim maps non-zero to that.
 */
#ifndef EXAB200_H
#define EXAB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct exab200_ctx exab200_ctx;

enum { EXAB200_FCC = 0, EXAB200_BCC = 1, EXAB200_HCP = 2 };              /* Model.ExaCMech.xtal_type */
enum { EXAB200_POWERVOCE = 0, EXAB200_POWERVOCENL = 1, EXAB200_MTSDD = 2 }; /* Model.ExaCMech.slip_type */
enum { EXAB200_PA = 0, EXAB200_EA = 1 };                                   /* Solvers.assembly */
enum { EXAB200_INTEG_FULL = 0, EXAB200_INTEG_BBAR = 1 };                   /* Solvers.integ_model */

typedef struct {
  int xtal;            /* EXAB200_FCC ... */
  int slip;            /* EXAB200_POWERVOCE ... */
  int nprops;          /* length of props (17 voce, 18 voce-nl, 24 mtsdd cubic) */
  const double* props; /* HOST pointer; order of src/mechanics_ecmech.hpp:395-405,444-458 */
  double temp_k;       /* Properties.temperature */
  long nelems;         /* local elements */
  long nnodes;         /* local nodes (L-vectors hold 3*nnodes doubles) */
  const int* e2n;      /* HOST pointer, 8*nelems node ids in NATIVE hex vertex order; may be NULL if
                          only the E-vector entry points are used */
  int assembly;        /* EXAB200_PA / EXAB200_EA */
  int integ;           /* EXAB200_INTEG_FULL / EXAB200_INTEG_BBAR */
  int device;          /* CUDA device ordinal */
} exab200_config;

const char* exab200_last_error(void);
int exab200_version(void);

/* ECMechXtalModel ctor (src/mechanics_ecmech.hpp:126-245) + operator scratch set-up
 * (src/mechanics_operator.cpp:227-262). */
int exab200_create(const exab200_config* cfg, exab200_ctx** out);
void exab200_destroy(exab200_ctx* ctx);

/* numStateVars = numHist + ne + 1 (src/mechanics_ecmech.hpp:141): 28 (fcc/bcc), 40 (hcp). */
int exab200_num_state_vars(const exab200_ctx* ctx);

/* Essential true-dof set (Hform->SetEssentialBC, src/mechanics_operator.cpp:279-285), given as a
 * HOST byte mask per node: bit i set = component i essential.  NULL clears it. */
int exab200_set_essential_mask(exab200_ctx* ctx, const unsigned char* h_mask_per_node);

/* ECMechXtalModel::init_state_vars (src/mechanics_ecmech.hpp:249-300). */
int exab200_hist_init(exab200_ctx* ctx, double* d_hist, void* stream);

/* NonlinearMechOperator::SetupJacobianTerms (src/mechanics_operator.cpp:350-391) with
 * ExaModel::UpdateEndCoords (src/mechanics_model.cpp:445-481) fused: J at the quadrature points of
 * x_beg + dt*vel (d_vel_L may be NULL -> J of x_beg). */
int exab200_setup_jacobians(exab200_ctx* ctx, const double* d_xbeg_L, const double* d_vel_L, double dt,
                            double* d_jac, void* stream);

/* ExaCMechModel::ModelSetup (src/mechanics_ecmech.cpp:192-258), velocity as L-vector (fused
 * restriction, src/mechanics_operator.cpp:344-345) or as E-vector exactly like the reference. */
int exab200_model_setup(exab200_ctx* ctx, double dt, const double* d_jac, const double* d_vel_L,
                        const double* d_stress0, const double* d_hist0, double* d_stress1, double* d_hist1,
                        double* d_matgrad, void* stream);
int exab200_model_setup_evec(exab200_ctx* ctx, double dt, const double* d_jac, const double* d_vel_E,
                             const double* d_stress0, const double* d_hist0, double* d_stress1,
                             double* d_hist1, double* d_matgrad, void* stream);
/* Where exab200_model_setup puts the material tangent and what the L-vector gradient entry points read.  VOIGT36
 * (default) is the reference's matGrad layout above, in the caller's d_matgrad.  COMPACT is a private record of the
 * fused PA path for cubic crystals (5x5 deviatoric operator, -dp/dlnV, deviatoric stress: 32 doubles, expanding to
 * exactly the same 6x6) kept densely packed in context-owned memory (256 B per point; d_matgrad is then neither
 * written nor read) that the gradient apply streams with 11 % fewer bytes; with it the E-vector entry points and
 * exab200_ea_assemble, which are defined on the reference layout, return an error, and exab200_grad_mult needs the
 * Jacobian array of the last exab200_setup_jacobians call. */
enum { EXAB200_TANGENT_VOIGT36 = 0, EXAB200_TANGENT_COMPACT = 1 };
int exab200_set_tangent_format(exab200_ctx* ctx, int format);

/* number of points whose local Newton solve failed in model_setup calls since the last query
 * (ExaCMech raises on failure; the shim turns >0 into MFEM_ABORT).  Synchronises the stream. */
int exab200_failed_points(exab200_ctx* ctx, void* stream, int* out);

/* Residual: ExaNLFIntegrator / ICExaNLFIntegrator AssemblePA + AddMultPA
 * (src/mechanics_integrators.cpp:160-314,518-557,1809-2088).  _evec: y_E += ... like AddMultPA.
 * L-vector form = MultVec (src/mechanics_operator_ext.cpp:176-202): y_L is overwritten and
 * essential dofs are zero. */
int exab200_residual_evec(exab200_ctx* ctx, const double* d_jac, const double* d_stress, double* d_y_E, void* stream);
int exab200_residual(exab200_ctx* ctx, const double* d_jac, const double* d_stress, double* d_y_L, void* stream);

/* Hform->GetGradient(x): AssembleGradPA / AssembleEA (src/mechanics_integrators.cpp:331-513,
 * 756-1017,1195-1604).  PA keeps only the two pointers (the 81-entry operands are never formed);
 * EA assembles the element matrices into ctx-owned storage. */
int exab200_grad_setup(exab200_ctx* ctx, double dt, const double* d_matgrad, const double* d_jac, void* stream);

/* Gradient operator action.
 * _evec: ExaNLFIntegrator::AddMultGradPA (src/mechanics_integrators.cpp:562-622), y_E += K_E x_E.
 * L-vector: PA/EANonlinearMechOperatorGradExt::TMult<local_action>
 * (src/mechanics_operator_ext.cpp:136-174,278-328): y_L overwritten; unless local_action the
 * essential dofs of x are treated as zero and y[ess] = 0. */
int exab200_grad_mult_evec(exab200_ctx* ctx, const double* d_x_E, double* d_y_E, void* stream);
int exab200_grad_mult(exab200_ctx* ctx, const double* d_x_L, double* d_y_L, int local_action, void* stream);
/* Same operator for a device-resident CG loop (the role mfem::CGSolver's oper->Mult + Dot(d, z) pair plays,
 * src/system_driver.cpp:166-178): flags = EXAB200_LOCAL_ACTION | EXAB200_NO_ZERO (y_L is accumulated into; the
 * caller zeroed it); if d_dot_accum != NULL, x^T K x (essential dofs of x as zero) is ADDED to *d_dot_accum
 * from the element contributions, so the CG denominator costs no extra pass over the vectors. */
enum { EXAB200_LOCAL_ACTION = 1, EXAB200_NO_ZERO = 2 };
int exab200_grad_mult_ex(exab200_ctx* ctx, const double* d_x_L, double* d_y_L, int flags, double* d_dot_accum,
                         void* stream);

/* The same operator for one z-slab of a multi-GPU run, WITH what ParNonlinearForm wraps around it in the reference
 * (P^T ... P: the shared-dof sum, src/mechanics_operator_ext.cpp:149,157) and the all-reduce of the fused CG
 * denominator -- one kernel over NVLink peer memory: the two boundary element layers are processed first, a few CTAs
 * push their interface-plane sums into the neighbours' mailboxes and add the neighbours' while the interior streams,
 * and *d_dot_accum ends up summed over all ranks.  Mailboxes are CUDA-IPC mapped buffers laid out as documented in
 * exaconstit_b200/csrc/exab200_p2p.cuh; seq_halo / seq_scal are the caller's monotonically increasing sequence
 * numbers of that protocol.  exab200_grad_mult_halo_supported() tells whether the context can use it (PA with compact
 * tangent records, >= 3 element layers); otherwise call exab200_grad_mult_ex and exchange separately. */
typedef struct exab200_halo {
  double* mailbox;           /* this rank's mailbox (device) */
  double* lo;                /* lower / upper z-neighbour's mailbox, NULL at the ends of the rank line */
  double* hi;
  double* peers[8];          /* every rank's mailbox (scalar all-reduce) */
  long long spin_limit;      /* clock64 ticks before a lost peer is reported instead of waited for */
  int rank, nranks;
  long plane;                /* nodes per interface plane */
  long layer_elems;          /* elements per z-layer; elements are ordered layer by layer */
  unsigned long long seq_halo, seq_scal;
} exab200_halo;
int exab200_grad_mult_halo_supported(exab200_ctx* ctx, const exab200_halo* h);
int exab200_grad_mult_halo(exab200_ctx* ctx, const double* d_x_L, double* d_y_L, int flags, double* d_dot_accum,
                           const exab200_halo* h, void* stream);

/* Deterministic operator option.  By default the L-vector kernels scatter element contributions with
 * red.global.add.f64, whose order varies run to run (results differ in the last bits).  With on != 0 the gradient
 * apply, the residual and the diagonal write E-vectors and a second kernel sums the contributions of the (at most 8)
 * elements around every node in ascending element order -- the explicit form of ElementRestriction::MultTranspose
 * (src/mechanics_operator_ext.cpp:149,198) -- and x^T K x is reduced in a fixed order: bitwise reproducible results at
 * the cost of one extra E-vector round trip per apply.  PA with compact tangent records only. */
int exab200_set_deterministic(exab200_ctx* ctx, int on);

/* AssembleGradDiagonalPA / EA AssembleDiagonal (src/mechanics_integrators.cpp:625-748,1607-1805;
 * src/mechanics_operator_ext.cpp:95-123,228-265).  L form: diag[ess] = 1. */
int exab200_grad_diag_evec(exab200_ctx* ctx, double* d_diag_E, void* stream);
int exab200_grad_diag(exab200_ctx* ctx, double* d_diag_L, void* stream);

/* AssembleEA into caller storage (src/mechanics_integrators.cpp:756-1017,1195-1604): emat += ... */
int exab200_ea_assemble(exab200_ctx* ctx, double dt, const double* d_matgrad, const double* d_jac,
                        double* d_emat, void* stream);
/* EA apply on E-vectors (src/mechanics_operator_ext.cpp:303-314). */
int exab200_ea_mult_evec(exab200_ctx* ctx, const double* d_emat, const double* d_x_E, double* d_y_E, void* stream);

/* ComputeVolAvgTensor (src/mechanics_kernels.hpp:19-134): d_out[0..vdim-1] = sum detJ W f,
 * d_out[vdim] = sum detJ W (local sums; the caller reduces over ranks and divides). vdim <= 40. */
int exab200_vol_sum(exab200_ctx* ctx, const double* d_jac, const double* d_qf, int vdim, double* d_out, void* stream);

/* ECMechXtalModel::calcDpMat (src/mechanics_ecmech.hpp:303-357): d_dp holds 9 doubles per point. */
int exab200_calc_dp(exab200_ctx* ctx, const double* d_hist, double* d_dp, void* stream);

/* exaconstit::kernel::grad_calc (src/mechanics_kernels.cpp:7-78) for an L-vector field
 * (used by CalculateDeformationGradient, src/mechanics_operator.cpp:393-427): d_grad 9 per point,
 * grad(i,t) at [t*3+i], overwritten. */
int exab200_grad_calc(exab200_ctx* ctx, const double* d_jac, const double* d_field_L, double* d_grad, void* stream);

/* Kernel launch counter (all launches issued through this context). */
long exab200_launch_count(const exab200_ctx* ctx);

/* Tuning knobs for the PA gradient apply: persistent CTAs per SM and tile variant -- benchmarking only.
 *   J streamed from HBM: CTA-tile kernel 0: 16 elems x 4 stages, 1: 16x2, 2: 32x2, 3: 8x4, 4: 16x3;
 *     warp-private pipelines 10: 4 warps x 2 stages (default, 2 CTAs/SM), 11: 4x3, 12: 8x2, 13: 2x3, 14: 4x4, 15: 3x3;
 *   J rebuilt in registers from the end coordinates stored by exab200_setup_jacobians (L-vector entry points, used
 *     whenever the bound J array is the one that call wrote): 20: 4x2, 21: 4x3, 22: 8x2, 23: 4x4, 24: 2x3, 25: 3x3,
 *     26: 2x2 (default, 6 CTAs/SM), 27: 1x2, 28: 1x3, 29: 3x2;  99: disable the rebuilt-J path;  98 / 97: L2
 *     evict_first hint on the operand stream on / off;
 *   compact tangent records (tiled, swizzled TMA): 30: 2x2 (default, 6 CTAs/SM), 31: 2x3, 32: 4x2, 33: 1x2, 34: 2x4, 35: 1x3;
 *   element-matrix apply of the L-vector EA path (bulk TMA pipeline): 40: 2x2 (default, 3 CTAs/SM), 41: 1x2, 42: 1x3,
 *     43: 2x3;  49: the plain (non-pipelined) kernel.
 *   variant + 100*k additionally sets the material-update occupancy target to k CTAs/SM. */
int exab200_set_tuning(exab200_ctx* ctx, int ctas_per_sm, int variant);

#ifdef __cplusplus
}
#endif
#endif /* EXAB200_H */

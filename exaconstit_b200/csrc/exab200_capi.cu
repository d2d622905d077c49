// C ABI of the exab200 library (see include/exab200.h for the contract and the reference
// interfaces each entry point replaces).
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/exab200.h"
#include "k_material.cuh"
#include "k_operator.cuh"
#include "material_host.hpp"

using namespace exab;

namespace {
thread_local std::string g_err;
int fail(const std::string& s) { g_err = s; return 1; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return 2;
}
#define CK(call)                                        \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

}  // namespace

struct exab200_ctx {
  exab200_config cfg;
  MatDev mat;
  int device = 0, sm_count = 148;
  int* d_e2n = nullptr;
  unsigned char* d_ess = nullptr;
  int* d_e2n_ess = nullptr;     // connectivity with the essential-dof bits in bits 28..30 (compact PA apply)
  bool have_ess = false;
  int* d_fail = nullptr;
  double* d_k1_idle = nullptr;  // scratch records for the idle threads of K1's last CTA
  // gradient operator state
  double grad_dt = 0.0;
  const double* d_matgrad = nullptr;
  const double* d_jac = nullptr;
  double* d_ea = nullptr;  // EA element matrices (assembly == EA)
  long launches = 0;
  int ctas_per_sm = 2;
  int variant = 10;
  // coordinate-rebuilt Jacobians (JX kernels): end coordinates written by setup_jacobians, valid for
  // the Jacobian array `xend_jac` they were written with
  double* d_xend = nullptr;
  const double* xend_jac = nullptr;
  int variant_jx = 26, ctas_jx = 6;  // variant_jx < 0 disables the JX path
  int l2_hint = 1;                   // operand stream of the gradient apply marked L2 evict_first
  // compact tangent records (EXAB200_TANGENT_COMPACT): written by model_setup, streamed by k_grad_mult_pa_c
  int tangent_fmt = 0;
  double* d_tan = nullptr;  // ctx-owned packed records, 32 doubles per point
  CUtensorMap tmap;
  bool tmap_valid = false;
  int variant_c = 30, ctas_c = 6;
  // operator apply with the interface-plane exchange folded in (exab200_grad_mult_halo)
  unsigned long long* d_halo_cnt = nullptr;
  unsigned long long halo_tiles_cum = 0, halo_ctas_cum = 0;
  int halo_ctas = 16;  // exchange CTAs of the fused kernel (B200, 8 ranks: 8 -> 137 us, 16 -> 115 us per apply with 4-deep loads)
  // deterministic (owner-computes) scatter: exab200_set_deterministic
  int deterministic = 0;
  int* d_n2e = nullptr;            // 8 slots per node: element*8 + native local node, ascending, -1 padded
  double* d_yE = nullptr;          // E-vector scratch, 24 doubles per element
  double* d_red_partial = nullptr; // block partials of the fixed-order reductions (gather dot, volume sums)
  unsigned int* d_red_counter = nullptr;
  int variant_ea = 40, ctas_ea = 3;
#ifndef EXAB_K1_MINB_DEFAULT
#define EXAB_K1_MINB_DEFAULT 2
#endif
  int k1_min_blocks = EXAB_K1_MINB_DEFAULT;  // K1 occupancy target (blocks of 128 threads per SM)  // PA gradient-apply tile configuration, see kVariants
};

static constexpr int kRedBlocksMax = 1184;  // 8 x 148 SMs
static inline unsigned eblocks(long nelems, int threads) { return (unsigned)((nelems * 8 + threads - 1) / threads); }
#define NEED_L(c) \
  if (!(c) || !(c)->d_e2n) return fail("L-vector entry point needs e2n in the config")
#define POST_LAUNCH(c)          \
  do {                          \
    ++(c)->launches;            \
    CK(cudaPeekAtLastError());  \
  } while (0)


// PA gradient-apply tile configurations {elements per tile, pipeline stages}; smem/CTA =
// stages * ept * 2896 B, so variant 0 runs 1 CTA (128 thr) per SM, 1 -> 2 CTAs, 2 -> 1 CTA of 256 thr,
// 3 -> up to 4 CTAs of 64 thr.
template <int EPT, int STAGES, int MODE, bool ESS>
static int launch_gm(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st) {
  using SM = GradMultSmem<EPT, STAGES>;
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device (context), not per process
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_grad_mult_pa<EPT, STAGES, MODE, ESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
    attr_set = true;
  }
  const long ntiles = (c->cfg.nelems + EPT - 1) / EPT;
  long grid = (long)c->sm_count * c->ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  k_grad_mult_pa<EPT, STAGES, MODE, ESS><<<(unsigned)grid, EPT * 8, sizeof(SM), st>>>(c->d_matgrad, c->d_jac, x, y, io,
                                                                                       c->cfg.nelems, c->grad_dt);
  POST_LAUNCH(c);
  return 0;
}
template <int NW, int STAGES, int MODE, bool ESS, bool JX = false>
static int launch_gmw(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  constexpr int smem = NW * STAGES * (JX ? kWarpStageBytesJX : kWarpStageBytes) + NW * STAGES * 8;
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device (context), not per process
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_grad_mult_pa_w<NW, STAGES, MODE, ESS, JX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long nwt = (c->cfg.nelems + 3) / 4;
  long grid = (long)c->sm_count * (JX ? c->ctas_jx : c->ctas_per_sm);
  if (grid * NW > nwt) grid = (nwt + NW - 1) / NW;
  k_grad_mult_pa_w<NW, STAGES, MODE, ESS, JX><<<(unsigned)grid, NW * 32, smem, st>>>(c->d_matgrad, c->d_jac, x, y, io,
                                                                                      c->cfg.nelems, c->grad_dt, dot, c->d_xend,
                                                                                      c->l2_hint);
  POST_LAUNCH(c);
  return 0;
}
// L-vector apply with the Jacobians rebuilt from the end coordinates {warps per CTA, stages}
template <bool ESS>
static int launch_grad_mult_jx(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  switch (c->variant_jx) {
    case 21: return launch_gmw<4, 3, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 22: return launch_gmw<8, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 23: return launch_gmw<4, 4, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 24: return launch_gmw<2, 3, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 25: return launch_gmw<3, 3, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 26: return launch_gmw<2, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 20: return launch_gmw<4, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 27: return launch_gmw<1, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 28: return launch_gmw<1, 3, LVEC, ESS, true>(c, x, y, io, st, dot);
    case 29: return launch_gmw<3, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
    default: return launch_gmw<2, 2, LVEC, ESS, true>(c, x, y, io, st, dot);
  }
}
// CUtensorMap over the packed tangent records [npts][32 doubles]; boxes of 32 rows x 16
// doubles, 128-byte swizzle.  cuTensorMapEncodeTiled is a driver entry point: resolved at run time, no -lcuda.
static int encode_tangent_map(exab200_ctx* c) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail("cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {32, (cuuint64_t)c->cfg.nelems * 8};
  const cuuint64_t strides[1] = {32 * sizeof(double)};
  const cuuint32_t box[2] = {16, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, c->d_tan, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  c->tmap_valid = true;
  return 0;
}
template <int NW, int STAGES, bool ESS>
static int launch_gmc(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  constexpr int smem = NW * STAGES * kWarpStageBytesC + NW * STAGES * 8 + 1024;
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device (context), not per process
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_grad_mult_pa_c<NW, STAGES, ESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long nwt = (c->cfg.nelems + 3) / 4;
  long grid = (long)c->sm_count * c->ctas_c;
  if (grid * NW > nwt) grid = (nwt + NW - 1) / NW;
  if (ESS) io.e2n = c->d_e2n_ess;
  k_grad_mult_pa_c<NW, STAGES, ESS><<<(unsigned)grid, NW * 32, smem, st>>>(c->tmap, x, y, io, c->cfg.nelems, c->grad_dt, dot,
                                                                           c->d_xend, HaloArgs{});
  POST_LAUNCH(c);
  return 0;
}
// owner-computes mode: the compact kernel writes the E-vector, k_evec_to_lvec sums per node in a fixed order
template <bool ESS>
static int launch_gmc_evout(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  constexpr int NW = 2, STAGES = 2;
  constexpr int smem = NW * STAGES * kWarpStageBytesC + NW * STAGES * 8 + 1024;
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_grad_mult_pa_c<NW, STAGES, ESS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long nwt = (c->cfg.nelems + 3) / 4;
  long grid = (long)c->sm_count * c->ctas_c;
  if (grid * NW > nwt) grid = (nwt + NW - 1) / NW;
  if (ESS) io.e2n = c->d_e2n_ess;
  k_grad_mult_pa_c<NW, STAGES, ESS, false, true><<<(unsigned)grid, NW * 32, smem, st>>>(c->tmap, x, c->d_yE, io, c->cfg.nelems,
                                                                                        c->grad_dt, nullptr, c->d_xend, HaloArgs{});
  POST_LAUNCH(c);
  const unsigned gb = (unsigned)std::min<long>(std::min<long>((c->cfg.nnodes + 255) / 256, (long)c->sm_count * 8), kRedBlocksMax);
  k_evec_to_lvec<ESS><<<gb, 256, 0, st>>>(c->d_yE, c->d_n2e, y, io.essmask, c->cfg.nnodes, 0, dot ? x : nullptr, c->d_red_partial,
                                          c->d_red_counter, dot);
  POST_LAUNCH(c);
  return 0;
}
// E-vector scratch -> L-vector for the residual (essential dofs 0) and the diagonal (essential dofs 1)
static int gather_evec(exab200_ctx* c, double* y, bool ess, int ess_one, cudaStream_t st) {
  const unsigned gb = (unsigned)std::min<long>((c->cfg.nnodes + 255) / 256, (long)c->sm_count * 8);
  if (ess) k_evec_to_lvec<true><<<gb, 256, 0, st>>>(c->d_yE, c->d_n2e, y, c->d_ess, c->cfg.nnodes, ess_one, nullptr, nullptr, nullptr, nullptr);
  else k_evec_to_lvec<false><<<gb, 256, 0, st>>>(c->d_yE, c->d_n2e, y, nullptr, c->cfg.nnodes, ess_one, nullptr, nullptr, nullptr, nullptr);
  POST_LAUNCH(c);
  return 0;
}
// the same kernel with the interface-plane exchange and the all-reduce of the denominator folded in
template <bool ESS>
static int launch_gmc_halo(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot, HaloArgs h) {
  constexpr int NW = 2, STAGES = 2;
  constexpr int smem = NW * STAGES * kWarpStageBytesC + NW * STAGES * 8 + 1024;
  static bool attr_set_dev[64] = {};
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_grad_mult_pa_c<NW, STAGES, ESS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (ESS) io.e2n = c->d_e2n_ess;
  const long nwt = (c->cfg.nelems + 3) / 4;
  long compute = (long)c->sm_count * c->ctas_c - h.ncomm;   // the exchange CTAs take the place of compute CTAs: one wave
  if (compute * NW > nwt) compute = (nwt + NW - 1) / NW;
  c->halo_tiles_cum += (unsigned long long)(h.nb_lo + h.nb_hi);
  c->halo_ctas_cum += (unsigned long long)compute;
  h.counters = c->d_halo_cnt;
  h.tiles_target = c->halo_tiles_cum;
  h.ctas_target = c->halo_ctas_cum;
  k_grad_mult_pa_c<NW, STAGES, ESS, true><<<(unsigned)(compute + h.ncomm), NW * 32, smem, st>>>(c->tmap, x, y, io, c->cfg.nelems,
                                                                                               c->grad_dt, dot, c->d_xend, h);
  POST_LAUNCH(c);
  return 0;
}
template <bool ESS>
static int launch_grad_mult_compact(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  switch (c->variant_c) {
    case 31: return launch_gmc<2, 3, ESS>(c, x, y, io, st, dot);
    case 32: return launch_gmc<4, 2, ESS>(c, x, y, io, st, dot);
    case 33: return launch_gmc<1, 2, ESS>(c, x, y, io, st, dot);
    case 34: return launch_gmc<2, 4, ESS>(c, x, y, io, st, dot);
    case 35: return launch_gmc<1, 3, ESS>(c, x, y, io, st, dot);
    default: return launch_gmc<2, 2, ESS>(c, x, y, io, st, dot);
  }
}
template <int NW, int STAGES, bool ESS>
static int launch_eap(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot, int ctas) {
  constexpr int smem = NW * STAGES * kEaStageBytes + NW * STAGES * 8;
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device (context), not per process
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_ea_mult_p<NW, STAGES, ESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const long nwt = (c->cfg.nelems + 3) / 4;
  long grid = (long)c->sm_count * ctas;
  if (grid * NW > nwt) grid = (nwt + NW - 1) / NW;
  k_ea_mult_p<NW, STAGES, ESS><<<(unsigned)grid, NW * 32, smem, st>>>(c->d_ea, x, y, io, c->cfg.nelems, dot);
  POST_LAUNCH(c);
  return 0;
}
// pipelined EA apply {warps per CTA, stages}: 40: 2x2 (3 CTAs/SM, default), 41: 1x2 (6), 42: 1x3 (4), 43: 2x3 (2); 49: plain kernel
template <bool ESS>
static int launch_ea_mult(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot) {
  switch (c->variant_ea) {
    case 41: return launch_eap<1, 2, ESS>(c, x, y, io, st, dot, c->ctas_ea);
    case 42: return launch_eap<1, 3, ESS>(c, x, y, io, st, dot, c->ctas_ea);
    case 43: return launch_eap<2, 3, ESS>(c, x, y, io, st, dot, c->ctas_ea);
    case 49:
      k_ea_mult<LVEC, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_ea, x, y, io, c->cfg.nelems, dot);
      POST_LAUNCH(c);
      return 0;
    default: return launch_eap<2, 2, ESS>(c, x, y, io, st, dot, c->ctas_ea);
  }
}
template <int MODE, bool ESS>
static int launch_grad_mult_pa(exab200_ctx* c, const double* x, double* y, ElemIO io, cudaStream_t st, double* dot = nullptr) {
  if (dot && c->variant < 10) return fail("fused dot product needs a warp-pipelined variant (>= 10)");
  switch (c->variant) {
    case 0: return launch_gm<16, 4, MODE, ESS>(c, x, y, io, st);
    case 1: return launch_gm<16, 2, MODE, ESS>(c, x, y, io, st);
    case 2: return launch_gm<32, 2, MODE, ESS>(c, x, y, io, st);
    case 3: return launch_gm<8, 4, MODE, ESS>(c, x, y, io, st);
    case 4: return launch_gm<16, 3, MODE, ESS>(c, x, y, io, st);
    // warp-private pipelines {warps per CTA, stages}
    case 10: return launch_gmw<4, 2, MODE, ESS>(c, x, y, io, st, dot);
    case 11: return launch_gmw<4, 3, MODE, ESS>(c, x, y, io, st, dot);
    case 12: return launch_gmw<8, 2, MODE, ESS>(c, x, y, io, st, dot);
    case 13: return launch_gmw<2, 3, MODE, ESS>(c, x, y, io, st, dot);
    case 14: return launch_gmw<4, 4, MODE, ESS>(c, x, y, io, st, dot);
    case 15: return launch_gmw<3, 3, MODE, ESS>(c, x, y, io, st, dot);
    default: return launch_gmw<4, 2, MODE, ESS>(c, x, y, io, st, dot);
  }
}

__global__ void k_set_ess_one(double* __restrict__ v, const unsigned char* __restrict__ ess, long nnodes) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnodes) return;
  const unsigned m = ess[n];
  if (m & 1) v[n] = 1.0;
  if (m & 2) v[nnodes + n] = 1.0;
  if (m & 4) v[2 * nnodes + n] = 1.0;
}


__global__ void __launch_bounds__(256) k_grad_calc(const double* __restrict__ jac, const double* __restrict__ f,
                                                   double* __restrict__ out, ElemIO io, long nelems) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  double c0 = 0, c1 = 0, c2 = 0;
  if (active) {
    const long nid = io.e2n[e * 8 + lex_to_native(lane)];
    c0 = f[nid]; c1 = f[io.nnodes + nid]; c2 = f[2 * io.nnodes + nid];
  }
  double d[3][3];
  nodal_to_qp_grad(c0, lane, d[0][0], d[0][1], d[0][2]);
  nodal_to_qp_grad(c1, lane, d[1][0], d[1][1], d[1][2]);
  nodal_to_qp_grad(c2, lane, d[2][0], d[2][1], d[2][2]);
  if (!active) return;
  double J[9], adj[9];
  const long p = e * 8 + lane;
  for (int i = 0; i < 9; ++i) J[i] = jac[p * 9 + i];
  const double idet = 1.0 / adjugate(J, adj);
  for (int t = 0; t < 3; ++t)
    for (int i = 0; i < 3; ++i)
      out[p * 9 + t * 3 + i] = (d[i][0] * adj[t] + d[i][1] * adj[3 + t] + d[i][2] * adj[6 + t]) * idet;
}


template <int NSLIP, int KIN, int MODE, int MINB>
static int launch_k1(exab200_ctx* c, double dt, const double* d_jac, const double* d_vel, const double* s0, const double* h0,
                     double* s1, double* h1, double* mg, cudaStream_t st) {
  static bool attr_set_dev[64] = {};  // cudaFuncSetAttribute is per device (context), not per process
  bool& attr_set = attr_set_dev[c->cfg.device & 63];
  if (!attr_set) {
    CK(cudaFuncSetAttribute(k_model_setup<NSLIP, KIN, MODE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kK1SmemBytes));
    attr_set = true;
  }
  const unsigned nb = eblocks(c->cfg.nelems, kJS);
  k_model_setup<NSLIP, KIN, MODE, MINB><<<nb, kJS, kK1SmemBytes, st>>>(c->mat, dt, d_jac, d_vel,
                                                                         MODE == LVEC ? c->d_e2n : nullptr, c->cfg.nnodes, s0, h0, s1,
                                                                         h1, c->tangent_fmt ? c->d_tan : mg, c->cfg.nelems,
                                                                         c->tangent_fmt ? mat::kTangentCompact : 1, c->d_fail, c->d_k1_idle);
  POST_LAUNCH(c);
  return 0;
}
template <int NSLIP, int KIN, int MODE>
static int launch_k1_occ(exab200_ctx* c, double dt, const double* d_jac, const double* d_vel, const double* s0,
                         const double* h0, double* s1, double* h1, double* mg, cudaStream_t st) {
  switch (c->k1_min_blocks) {
    case 1: return launch_k1<NSLIP, KIN, MODE, 1>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
    case 2: return launch_k1<NSLIP, KIN, MODE, 2>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
    default: return launch_k1<NSLIP, KIN, MODE, 3>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
  }
}
template <int MODE>
static int launch_k1_model(exab200_ctx* c, double dt, const double* d_jac, const double* d_vel, const double* s0,
                           const double* h0, double* s1, double* h1, double* mg, cudaStream_t st) {
  const bool km = c->mat.kin == KIN_KMBALD;
  if (c->mat.nslip == 12) {
    if (km) return launch_k1_occ<12, 1, MODE>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
    return launch_k1_occ<12, 0, MODE>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
  }
  if (!km) return fail("24 slip systems are only available with KMBalD kinetics");
  return launch_k1_occ<24, 1, MODE>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
}

// The reference skips the tangent transpose for EA on a device backend (src/mechanics_ecmech.cpp:155) and then
// reads the row-major matrix as column-major; we always store d sigma_i/d eps_j at [j*6+i].
static int model_setup_impl(exab200_ctx* c, int mode, double dt, const double* d_jac, const double* d_vel,
                            const double* s0, const double* h0, double* s1, double* h1, double* mg, void* stream) {
  if (!c) return fail("null ctx");
  if (!(dt > 0.0)) return fail("dt must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == LVEC) return launch_k1_model<LVEC>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
  return launch_k1_model<EVEC>(c, dt, d_jac, d_vel, s0, h0, s1, h1, mg, st);
}


extern "C" {

const char* exab200_last_error(void) { return g_err.c_str(); }
int exab200_version(void) { return 100; }

int exab200_create(const exab200_config* cfg, exab200_ctx** out) {
  if (!cfg || !out) return fail("null argument");
  if (cfg->nelems <= 0) return fail("nelems must be positive");
  exab200_ctx* c = new exab200_ctx();
  c->cfg = *cfg;
  c->cfg.props = nullptr;
  c->cfg.e2n = nullptr;
  std::string msg = build_material(c->mat, cfg->xtal, cfg->slip, cfg->props, cfg->nprops);
  if (!msg.empty()) { delete c; return fail(msg); }
  // integ_model = BBAR with assembly = PA is accepted like the reference accepts it: the residual and the diagonal are
  // B-bar (ICExaNLFIntegrator::AssemblePA / AddMultPA / AssembleGradDiagonalPA, src/mechanics_integrators.cpp:1809-2088,
  // 1607-1805) while the gradient apply is the plain operator, because ICExaNLFIntegrator inherits
  // ExaNLFIntegrator::AssembleGradPA / AddMultGradPA (src/mechanics_integrators.hpp:107-110) -- a quirk of the reference
  // (SURVEY.md App. C.4); production runs B-bar with EA.
  c->device = cfg->device;
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaSetDevice"); }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, c->device);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaGetDeviceProperties"); }
  if (prop.major < 10) { delete c; return fail("exab200 kernels are built for sm_100a (Blackwell) only"); }
  c->sm_count = prop.multiProcessorCount;
  if (cfg->e2n) {
    for (long i = 0; i < 8 * cfg->nelems; ++i)
      if (cfg->e2n[i] < 0 || cfg->e2n[i] >= cfg->nnodes) { delete c; return fail("e2n entry out of range"); }
    CK(cudaMalloc(&c->d_e2n, sizeof(int) * 8 * cfg->nelems));
    CK(cudaMemcpy(c->d_e2n, cfg->e2n, sizeof(int) * 8 * cfg->nelems, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_ess, cfg->nnodes));
    CK(cudaMemset(c->d_ess, 0, cfg->nnodes));
    if (cfg->assembly == EXAB200_PA) CK(cudaMalloc(&c->d_xend, sizeof(double) * 3 * cfg->nnodes));
  }
  CK(cudaMalloc(&c->d_fail, sizeof(int)));
  CK(cudaMemset(c->d_fail, 0, sizeof(int)));
  CK(cudaMalloc(&c->d_k1_idle, sizeof(double) * kJS * kIdleRecord));
  // fixed-order reductions (volume sums, the deterministic gather's dot): block partials + arrival counters
  CK(cudaMalloc(&c->d_red_partial, sizeof(double) * kRedBlocksMax * 41));
  CK(cudaMalloc(&c->d_red_counter, 2 * sizeof(unsigned int)));
  CK(cudaMemset(c->d_red_counter, 0, 2 * sizeof(unsigned int)));
  if (cfg->assembly == EXAB200_EA) CK(cudaMalloc(&c->d_ea, sizeof(double) * 576 * cfg->nelems));
  *out = c;
  return 0;
}

void exab200_destroy(exab200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaFree(c->d_e2n);
  cudaFree(c->d_ess);
  cudaFree(c->d_e2n_ess);
  cudaFree(c->d_fail);
  cudaFree(c->d_k1_idle);
  cudaFree(c->d_halo_cnt);
  cudaFree(c->d_n2e); cudaFree(c->d_yE); cudaFree(c->d_red_partial); cudaFree(c->d_red_counter);
  cudaFree(c->d_ea);
  cudaFree(c->d_xend);
  cudaFree(c->d_tan);
  delete c;
}

int exab200_num_state_vars(const exab200_ctx* c) { return c ? c->mat.nhist : -1; }
long exab200_launch_count(const exab200_ctx* c) { return c ? c->launches : -1; }
int exab200_set_tuning(exab200_ctx* c, int ctas_per_sm, int variant) {
  if (!c || ctas_per_sm < 1 || ctas_per_sm > 16 || variant < 0 || variant > 499) return fail("bad tuning");
  const int v = variant % 100;
  if (variant >= 100) c->k1_min_blocks = variant / 100;  // e.g. 210 -> K1 with 2 blocks/SM, K2 variant 10
  if (v >= 20 && v <= 29) { c->variant_jx = v; c->ctas_jx = ctas_per_sm; return 0; }
  if (v >= 30 && v <= 35) { c->variant_c = v; c->ctas_c = ctas_per_sm; return 0; }
  if (v >= 40 && v <= 49) { c->variant_ea = v; c->ctas_ea = ctas_per_sm; return 0; }
  if (v == 99) { c->variant_jx = -1; return 0; }
  if (v == 98 || v == 97) { c->l2_hint = (v == 98); return 0; }  // 98 / 97: L2 evict_first hint on / off  // stream J from HBM (the E-vector entry points always do)
  if (v > 15) return fail("bad tuning");
  c->ctas_per_sm = ctas_per_sm;
  c->variant = v;
  return 0;
}

int exab200_set_tangent_format(exab200_ctx* c, int fmt) {
  if (!c) return fail("null ctx");
  if (fmt != EXAB200_TANGENT_VOIGT36 && fmt != EXAB200_TANGENT_COMPACT) return fail("unknown tangent format");
  if (fmt == EXAB200_TANGENT_COMPACT) {
    if (c->cfg.assembly != EXAB200_PA) return fail("compact tangent records are for partial assembly");
    if (c->mat.Kvd != 0.0) return fail("compact tangent records are defined for cubic crystals only");
    if (!c->d_e2n) return fail("compact tangent records need the L-vector connectivity");
  }
  if (fmt == EXAB200_TANGENT_COMPACT && !c->d_tan) CK(cudaMalloc(&c->d_tan, sizeof(double) * 32 * 8 * c->cfg.nelems));
  c->tangent_fmt = fmt;
  c->tmap_valid = false;
  return 0;
}

int exab200_set_essential_mask(exab200_ctx* c, const unsigned char* h_mask) {
  if (!c || !c->d_ess) return fail("context has no L-vector connectivity");
  if (h_mask) {
    CK(cudaMemcpy(c->d_ess, h_mask, c->cfg.nnodes, cudaMemcpyHostToDevice));
    c->have_ess = false;
    for (long i = 0; i < c->cfg.nnodes; ++i)
      if (h_mask[i] & 7) { c->have_ess = true; break; }
    if (c->have_ess) {
      if (c->cfg.nnodes > (long)kEssNodeMask) return fail("more than 2^28 nodes in one context");
      const long n = 8 * c->cfg.nelems;
      if (!c->d_e2n_ess) CK(cudaMalloc(&c->d_e2n_ess, sizeof(int) * n));
      k_fold_ess_mask<<<(unsigned)((n + 255) / 256), 256>>>(c->d_e2n, c->d_ess, c->d_e2n_ess, n);
      POST_LAUNCH(c);
      CK(cudaDeviceSynchronize());
    }
  } else {
    CK(cudaMemset(c->d_ess, 0, c->cfg.nnodes));
    c->have_ess = false;
  }
  return 0;
}

int exab200_hist_init(exab200_ctx* c, double* d_hist, void* stream) {
  if (!c) return fail("null ctx");
  const long npts = c->cfg.nelems * 8;
  k_hist_init<<<(unsigned)((npts + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c->mat, d_hist, npts);
  POST_LAUNCH(c);
  return 0;
}

int exab200_setup_jacobians(exab200_ctx* c, const double* d_xbeg, const double* d_vel, double dt, double* d_jac,
                            void* stream) {
  NEED_L(c);
  ElemIO io{c->d_e2n, nullptr, c->cfg.nnodes};
  k_jacobians<<<eblocks(c->cfg.nelems, 256), 256, 0, (cudaStream_t)stream>>>(d_xbeg, d_vel, dt, d_jac, io, c->cfg.nelems,
                                                                             c->d_xend);
  c->xend_jac = d_jac;
  POST_LAUNCH(c);
  return 0;
}

int exab200_model_setup(exab200_ctx* c, double dt, const double* d_jac, const double* d_vel_L, const double* s0,
                        const double* h0, double* s1, double* h1, double* mg, void* stream) {
  NEED_L(c);
  return model_setup_impl(c, LVEC, dt, d_jac, d_vel_L, s0, h0, s1, h1, mg, stream);
}
int exab200_model_setup_evec(exab200_ctx* c, double dt, const double* d_jac, const double* d_vel_E, const double* s0,
                             const double* h0, double* s1, double* h1, double* mg, void* stream) {
  return model_setup_impl(c, EVEC, dt, d_jac, d_vel_E, s0, h0, s1, h1, mg, stream);
}

int exab200_failed_points(exab200_ctx* c, void* stream, int* out) {
  if (!c || !out) return fail("null argument");
  CK(cudaMemcpyAsync(out, c->d_fail, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaMemsetAsync(c->d_fail, 0, sizeof(int), (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int exab200_residual_evec(exab200_ctx* c, const double* d_jac, const double* d_stress, double* d_y_E, void* stream) {
  if (!c) return fail("null ctx");
  ElemIO io{nullptr, nullptr, 0};
  const unsigned nb = eblocks(c->cfg.nelems, 256);
  if (c->cfg.integ == EXAB200_INTEG_BBAR)
    k_residual<EVEC, true><<<nb, 256, 0, (cudaStream_t)stream>>>(d_stress, d_jac, d_y_E, io, c->cfg.nelems);
  else
    k_residual<EVEC, false><<<nb, 256, 0, (cudaStream_t)stream>>>(d_stress, d_jac, d_y_E, io, c->cfg.nelems);
  POST_LAUNCH(c);
  return 0;
}

int exab200_residual(exab200_ctx* c, const double* d_jac, const double* d_stress, double* d_y_L, void* stream) {
  NEED_L(c);
  if (c->deterministic) {
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(c->d_yE, 0, sizeof(double) * 24 * c->cfg.nelems, st));
    ElemIO ioe{nullptr, nullptr, 0};
    const unsigned nbe = eblocks(c->cfg.nelems, 256);
    if (c->cfg.integ == EXAB200_INTEG_BBAR) k_residual<EVEC, true><<<nbe, 256, 0, st>>>(d_stress, d_jac, c->d_yE, ioe, c->cfg.nelems);
    else k_residual<EVEC, false><<<nbe, 256, 0, st>>>(d_stress, d_jac, c->d_yE, ioe, c->cfg.nelems);
    POST_LAUNCH(c);
    return gather_evec(c, d_y_L, c->have_ess, 0, st);
  }
  CK(cudaMemsetAsync(d_y_L, 0, sizeof(double) * 3 * c->cfg.nnodes, (cudaStream_t)stream));
  ElemIO io{c->d_e2n, c->have_ess ? c->d_ess : nullptr, c->cfg.nnodes};
  const unsigned nb = eblocks(c->cfg.nelems, 256);
  if (c->cfg.integ == EXAB200_INTEG_BBAR)
    k_residual<LVEC, true><<<nb, 256, 0, (cudaStream_t)stream>>>(d_stress, d_jac, d_y_L, io, c->cfg.nelems);
  else
    k_residual<LVEC, false><<<nb, 256, 0, (cudaStream_t)stream>>>(d_stress, d_jac, d_y_L, io, c->cfg.nelems);
  POST_LAUNCH(c);
  return 0;
}

int exab200_ea_assemble(exab200_ctx* c, double dt, const double* d_matgrad, const double* d_jac, double* d_emat,
                        void* stream) {
  if (!c) return fail("null ctx");
  if (c->tangent_fmt) return fail("element assembly reads the reference's 36-entry tangent layout (tangent format 0)");
  const unsigned nb = eblocks(c->cfg.nelems, 128);
  if (c->cfg.integ == EXAB200_INTEG_BBAR)
    k_assemble_ea<true><<<nb, 128, 0, (cudaStream_t)stream>>>(d_matgrad, d_jac, d_emat, c->cfg.nelems, dt);
  else
    k_assemble_ea<false><<<nb, 128, 0, (cudaStream_t)stream>>>(d_matgrad, d_jac, d_emat, c->cfg.nelems, dt);
  POST_LAUNCH(c);
  return 0;
}

int exab200_grad_setup(exab200_ctx* c, double dt, const double* d_matgrad, const double* d_jac, void* stream) {
  if (!c) return fail("null ctx");
  c->grad_dt = dt;
  if (c->tangent_fmt && !c->tmap_valid) {
    const int rc = encode_tangent_map(c);
    if (rc) return rc;
  }
  c->d_matgrad = d_matgrad;
  c->d_jac = d_jac;
  if (c->cfg.assembly == EXAB200_EA) {  // context-owned element matrices, lane-interleaved private layout
    CK(cudaMemsetAsync(c->d_ea, 0, sizeof(double) * 576 * c->cfg.nelems, (cudaStream_t)stream));
    const unsigned nb = eblocks(c->cfg.nelems, 128);
    if (c->cfg.integ == EXAB200_INTEG_BBAR)
      k_assemble_ea<true, true><<<nb, 128, 0, (cudaStream_t)stream>>>(d_matgrad, d_jac, c->d_ea, c->cfg.nelems, dt);
    else
      k_assemble_ea<false, true><<<nb, 128, 0, (cudaStream_t)stream>>>(d_matgrad, d_jac, c->d_ea, c->cfg.nelems, dt);
    POST_LAUNCH(c);
    return 0;
  }
  return 0;
}

int exab200_grad_mult_evec(exab200_ctx* c, const double* d_x_E, double* d_y_E, void* stream) {
  if (!c || !c->d_matgrad) return fail("grad_setup has not been called");
  if (c->tangent_fmt) return fail("E-vector entry points read the reference's 36-entry tangent layout (tangent format 0)");
  ElemIO io{nullptr, nullptr, 0};
  if (c->cfg.assembly == EXAB200_EA) {
    k_ea_mult<EVEC, true><<<eblocks(c->cfg.nelems, 256), 256, 0, (cudaStream_t)stream>>>(c->d_ea, d_x_E, d_y_E, io, c->cfg.nelems, nullptr);
    POST_LAUNCH(c);
    return 0;
  }
  return launch_grad_mult_pa<EVEC, false>(c, d_x_E, d_y_E, io, (cudaStream_t)stream);
}

int exab200_grad_mult_ex(exab200_ctx* c, const double* d_x_L, double* d_y_L, int flags, double* d_dot_accum, void* stream) {
  NEED_L(c);
  if (!c->d_matgrad) return fail("grad_setup has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  const bool local_action = flags & EXAB200_LOCAL_ACTION;
  if (!(flags & EXAB200_NO_ZERO)) CK(cudaMemsetAsync(d_y_L, 0, sizeof(double) * 3 * c->cfg.nnodes, st));
  const bool ess = c->have_ess && !local_action;
  ElemIO io{c->d_e2n, ess ? c->d_ess : nullptr, c->cfg.nnodes};
  if (c->cfg.assembly == EXAB200_EA) {
    if (ess) return launch_ea_mult<true>(c, d_x_L, d_y_L, io, st, d_dot_accum);
    return launch_ea_mult<false>(c, d_x_L, d_y_L, io, st, d_dot_accum);
  }
  if (c->tangent_fmt) {
    if (!(c->d_xend && c->xend_jac == c->d_jac))
      return fail("compact tangent records need the Jacobian array written by the last exab200_setup_jacobians call");
    if (c->deterministic) {
      if (ess) return launch_gmc_evout<true>(c, d_x_L, d_y_L, io, st, d_dot_accum);
      return launch_gmc_evout<false>(c, d_x_L, d_y_L, io, st, d_dot_accum);
    }
    if (ess) return launch_grad_mult_compact<true>(c, d_x_L, d_y_L, io, st, d_dot_accum);
    return launch_grad_mult_compact<false>(c, d_x_L, d_y_L, io, st, d_dot_accum);
  }
  if (c->variant_jx >= 0 && c->d_xend && c->xend_jac == c->d_jac) {
    if (ess) return launch_grad_mult_jx<true>(c, d_x_L, d_y_L, io, st, d_dot_accum);
    return launch_grad_mult_jx<false>(c, d_x_L, d_y_L, io, st, d_dot_accum);
  }
  if (ess) return launch_grad_mult_pa<LVEC, true>(c, d_x_L, d_y_L, io, st, d_dot_accum);
  return launch_grad_mult_pa<LVEC, false>(c, d_x_L, d_y_L, io, st, d_dot_accum);
}

int exab200_set_deterministic(exab200_ctx* c, int on) {
  NEED_L(c);
  if (!on) { c->deterministic = 0; return 0; }
  if (c->cfg.assembly != EXAB200_PA || !c->tangent_fmt)
    return fail("the deterministic (owner-computes) scatter is implemented for the PA path with compact tangent records");
  if (!c->d_n2e) {
    const long ne = c->cfg.nelems, nn = c->cfg.nnodes;
    std::vector<int> e2n(8 * ne), n2e(8 * nn, -1), fill(nn, 0);
    CK(cudaMemcpy(e2n.data(), c->d_e2n, sizeof(int) * 8 * ne, cudaMemcpyDeviceToHost));
    for (long e = 0; e < ne; ++e)
      for (int a = 0; a < 8; ++a) {
        const int nd = e2n[e * 8 + a];
        if (fill[nd] >= 8) return fail("deterministic scatter: a node with more than 8 elements around it");
        n2e[(long)nd * 8 + fill[nd]++] = (int)(e * 8 + a);
      }
    CK(cudaMalloc(&c->d_n2e, sizeof(int) * 8 * nn));
    CK(cudaMemcpy(c->d_n2e, n2e.data(), sizeof(int) * 8 * nn, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_yE, sizeof(double) * 24 * ne));
  }
  c->deterministic = 1;
  return 0;
}

int exab200_grad_mult_halo_supported(exab200_ctx* c, const exab200_halo* h) {
  if (!c || !h || !c->cfg.nnodes) return 0;
  if (c->deterministic) return 0;                                           // owner-computes scatter: exchange separately
  if (c->cfg.assembly != EXAB200_PA || !c->tangent_fmt) return 0;          // compact-record PA kernel only
  if (h->nranks < 2 || h->nranks > 8 || h->layer_elems <= 0 || c->cfg.nelems % h->layer_elems) return 0;
  const long layers = c->cfg.nelems / h->layer_elems;
  const long nwt = (c->cfg.nelems + 3) / 4, nb_lo = (h->layer_elems + 3) / 4, hi_start = (c->cfg.nelems - h->layer_elems) / 4;
  if (layers < 3 || hi_start < nb_lo || nwt < 64) return 0;                // the two boundary layers must be distinct tiles
  return 1;
}

int exab200_grad_mult_halo(exab200_ctx* c, const double* d_x_L, double* d_y_L, int flags, double* d_dot_accum,
                           const exab200_halo* hd, void* stream) {
  NEED_L(c);
  if (!c->d_matgrad) return fail("grad_setup has not been called");
  if (!exab200_grad_mult_halo_supported(c, hd)) return fail("exab200_grad_mult_halo: unsupported configuration (see exab200_grad_mult_halo_supported)");
  if (!(c->d_xend && c->xend_jac == c->d_jac))
    return fail("compact tangent records need the Jacobian array written by the last exab200_setup_jacobians call");
  cudaStream_t st = (cudaStream_t)stream;
  if (!c->d_halo_cnt) {
    CK(cudaMalloc(&c->d_halo_cnt, 2 * sizeof(unsigned long long)));
    CK(cudaMemset(c->d_halo_cnt, 0, 2 * sizeof(unsigned long long)));
    if (const char* e = std::getenv("EXAB200_HALO_CTAS")) c->halo_ctas = std::max(1, std::min(64, std::atoi(e)));
  }
  const bool local_action = flags & EXAB200_LOCAL_ACTION;
  if (!(flags & EXAB200_NO_ZERO)) CK(cudaMemsetAsync(d_y_L, 0, sizeof(double) * 3 * c->cfg.nnodes, st));
  const bool ess = c->have_ess && !local_action;
  ElemIO io{c->d_e2n, ess ? c->d_ess : nullptr, c->cfg.nnodes};
  HaloArgs h{};
  h.mb = hd->mailbox; h.lo = hd->lo; h.hi = hd->hi;
  for (int r = 0; r < 8; ++r) h.peers.p[r] = hd->peers[r];
  h.peers.spin_limit = hd->spin_limit;
  h.rank = hd->rank; h.nranks = hd->nranks; h.ncomm = c->halo_ctas;
  h.nn = c->cfg.nnodes; h.plane = hd->plane;
  const long nwt = (c->cfg.nelems + 3) / 4;
  h.nb_lo = hd->lo ? (hd->layer_elems + 3) / 4 : 0;
  h.hi_start = (c->cfg.nelems - hd->layer_elems) / 4;
  h.nb_hi = hd->hi ? nwt - h.hi_start : 0;
  h.seq_halo = hd->seq_halo; h.seq_scal = hd->seq_scal;
  if (ess) return launch_gmc_halo<true>(c, d_x_L, d_y_L, io, st, d_dot_accum, h);
  return launch_gmc_halo<false>(c, d_x_L, d_y_L, io, st, d_dot_accum, h);
}

int exab200_grad_mult(exab200_ctx* c, const double* d_x_L, double* d_y_L, int local_action, void* stream) {
  return exab200_grad_mult_ex(c, d_x_L, d_y_L, local_action ? EXAB200_LOCAL_ACTION : 0, nullptr, stream);
}

int exab200_ea_mult_evec(exab200_ctx* c, const double* d_emat, const double* d_x_E, double* d_y_E, void* stream) {
  if (!c) return fail("null ctx");
  ElemIO io{nullptr, nullptr, 0};
  k_ea_mult<EVEC><<<eblocks(c->cfg.nelems, 256), 256, 0, (cudaStream_t)stream>>>(d_emat, d_x_E, d_y_E, io, c->cfg.nelems, nullptr);
  POST_LAUNCH(c);
  return 0;
}

int exab200_grad_diag_evec(exab200_ctx* c, double* d_diag_E, void* stream) {
  if (!c || !c->d_matgrad) return fail("grad_setup has not been called");
  if (c->tangent_fmt) return fail("E-vector entry points read the reference's 36-entry tangent layout (tangent format 0)");
  ElemIO io{nullptr, nullptr, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (c->cfg.assembly == EXAB200_EA)
    k_ea_diag<EVEC, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_ea, d_diag_E, io, c->cfg.nelems);
  else if (c->cfg.integ == EXAB200_INTEG_BBAR)
    k_grad_diag<EVEC, false, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_matgrad, c->d_jac, d_diag_E, io, c->cfg.nelems, c->grad_dt);
  else
    k_grad_diag<EVEC><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_matgrad, c->d_jac, d_diag_E, io, c->cfg.nelems, c->grad_dt);
  POST_LAUNCH(c);
  return 0;
}

int exab200_grad_diag(exab200_ctx* c, double* d_diag_L, void* stream) {
  NEED_L(c);
  if (!c->d_matgrad) return fail("grad_setup has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->deterministic) {
    CK(cudaMemsetAsync(c->d_yE, 0, sizeof(double) * 24 * c->cfg.nelems, st));
    ElemIO ioe{nullptr, nullptr, 0};
    const unsigned nbe = eblocks(c->cfg.nelems, 256);
    if (c->cfg.integ == EXAB200_INTEG_BBAR) k_grad_diag<EVEC, true, true><<<nbe, 256, 0, st>>>(c->d_tan, c->d_jac, c->d_yE, ioe, c->cfg.nelems, c->grad_dt);
    else k_grad_diag<EVEC, true, false><<<nbe, 256, 0, st>>>(c->d_tan, c->d_jac, c->d_yE, ioe, c->cfg.nelems, c->grad_dt);
    POST_LAUNCH(c);
    return gather_evec(c, d_diag_L, c->have_ess, 1, st);
  }
  CK(cudaMemsetAsync(d_diag_L, 0, sizeof(double) * 3 * c->cfg.nnodes, st));
  ElemIO io{c->d_e2n, nullptr, c->cfg.nnodes};
  if (c->cfg.assembly == EXAB200_EA)
    k_ea_diag<LVEC, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_ea, d_diag_L, io, c->cfg.nelems);
  else if (c->cfg.integ == EXAB200_INTEG_BBAR && c->tangent_fmt)
    k_grad_diag<LVEC, true, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_tan, c->d_jac, d_diag_L, io, c->cfg.nelems, c->grad_dt);
  else if (c->cfg.integ == EXAB200_INTEG_BBAR)
    k_grad_diag<LVEC, false, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_matgrad, c->d_jac, d_diag_L, io, c->cfg.nelems, c->grad_dt);
  else if (c->tangent_fmt)
    k_grad_diag<LVEC, true><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_tan, c->d_jac, d_diag_L, io, c->cfg.nelems, c->grad_dt);
  else
    k_grad_diag<LVEC><<<eblocks(c->cfg.nelems, 256), 256, 0, st>>>(c->d_matgrad, c->d_jac, d_diag_L, io, c->cfg.nelems, c->grad_dt);
  POST_LAUNCH(c);
  if (c->have_ess) {
    k_set_ess_one<<<(unsigned)((c->cfg.nnodes + 255) / 256), 256, 0, st>>>(d_diag_L, c->d_ess, c->cfg.nnodes);
    POST_LAUNCH(c);
  }
  return 0;
}

int exab200_vol_sum(exab200_ctx* c, const double* d_jac, const double* d_qf, int vdim, double* d_out, void* stream) {
  if (!c) return fail("null ctx");
  if (vdim < 1 || vdim > 40) return fail("vdim out of range (1..40)");
  cudaStream_t st = (cudaStream_t)stream;
  const long npts = c->cfg.nelems * 8;
  long nb = (npts + 255) / 256;
  if (nb > (long)c->sm_count * 8) nb = (long)c->sm_count * 8;
  if (nb > kRedBlocksMax) nb = kRedBlocksMax;
  if (vdim <= 9)
    k_vol_sum<9><<<(unsigned)nb, 256, 0, st>>>(d_qf, d_jac, vdim, npts, d_out, c->d_red_partial, c->d_red_counter + 1);
  else
    k_vol_sum<40><<<(unsigned)nb, 256, 0, st>>>(d_qf, d_jac, vdim, npts, d_out, c->d_red_partial, c->d_red_counter + 1);
  POST_LAUNCH(c);
  return 0;
}

int exab200_calc_dp(exab200_ctx* c, const double* d_hist, double* d_dp, void* stream) {
  if (!c) return fail("null ctx");
  const long npts = c->cfg.nelems * 8;
  k_calc_dp<<<(unsigned)((npts + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c->mat, d_hist, d_dp, npts);
  POST_LAUNCH(c);
  return 0;
}

int exab200_grad_calc(exab200_ctx* c, const double* d_jac, const double* d_field_L, double* d_grad, void* stream) {
  NEED_L(c);
  ElemIO io{c->d_e2n, nullptr, c->cfg.nnodes};
  k_grad_calc<<<eblocks(c->cfg.nelems, 256), 256, 0, (cudaStream_t)stream>>>(d_jac, d_field_L, d_grad, io, c->cfg.nelems);
  POST_LAUNCH(c);
  return 0;
}

}  // extern "C"

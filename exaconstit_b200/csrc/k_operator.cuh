// Element kernels of the matrix-free operator path (sm_100a, fp64, no tensor cores: 3x3 / 6x6
// tensor algebra).  Reference behaviour each kernel replaces is cited at its definition
// (paths relative to the ExaConstit source tree).
#pragma once
#include <cuda.h>

#include "exab200_common.cuh"
#include "exab200_p2p.cuh"
#include "material_point.hpp"

namespace exab {

// How an element kernel reaches the nodal vectors.
//   LVEC: x/y are L-vectors (byNODES: comp*nnodes + node); gather through e2n, scatter with
//         red.global.add.f64; essential dofs masked on the fly (bit i of essmask[node]).
//   EVEC: x/y are E-vectors X(a,i,e) = x[e*24 + i*8 + a_native] exactly as the reference's
//         integrator virtuals receive them (src/mechanics_integrators.cpp:580-581).
enum VecMode { LVEC = 0, EVEC = 1 };

struct ElemIO {
  const int* __restrict__ e2n;               // 8*nelems, NATIVE order (LVEC)
  const unsigned char* __restrict__ essmask;  // nnodes or nullptr (LVEC)
  long nnodes;
};
// connectivity with the essential-dof bits folded in (k_grad_mult_pa_c<..., ESS = true>): node | mask << 28
constexpr int kEssShift = 28;
constexpr int kEssNodeMask = (1 << kEssShift) - 1;
__global__ void __launch_bounds__(256) k_fold_ess_mask(const int* __restrict__ e2n, const unsigned char* __restrict__ ess,
                                                       int* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int nd = e2n[i];
    out[i] = nd | ((int)(ess[nd] & 7) << kEssShift);
  }
}

// ------------------------------------------------------------------------------------------
// K2: gradient (tangent-stiffness) operator apply, y += K x, matrix-free from the 6x6 material
// tangent and the Jacobians.  Replaces, in one launch:
//   ExaModel::TransformMatGradTo4D            src/mechanics_model.cpp:949-1061
//   ExaNLFIntegrator::AssembleGradPA          src/mechanics_integrators.cpp:331-513
//   ExaNLFIntegrator::AddMultGradPA           src/mechanics_integrators.cpp:562-622
//   and in LVEC mode also ElementRestriction Mult/MultTranspose and the essential-dof masking of
//   PANonlinearMechOperatorGradExt::TMult     src/mechanics_operator_ext.cpp:136-174
// y_{a,k} = sum_q dt W_q detJ (dN_a/dx_p) C(p,k,l,m) (dv_l/dx_m), C(p,k,l,m) = K(v(p,k), v(l,m)).
//
// Persistent CTAs; each stage holds one tile of EPT elements: matGrad (36 dbl/qpt) and J
// (9 dbl/qpt) are streamed HBM -> shared memory with cp.async.bulk (TMA 1-D) completing on an
// mbarrier, STAGES deep, so >= (STAGES-1) * EPT * 2880 B are in flight per SM while the current
// tile is contracted out of shared memory.  The two 1152-B halves of an element's matGrad block
// are skewed by 16 B in shared memory so that the per-lane LDS.128 stream is bank-conflict free.
// ------------------------------------------------------------------------------------------
constexpr int kCHalfBytes = 4 * 36 * 8;           // 1152: four quadrature points of matGrad
constexpr int kCElemSmem = 2 * kCHalfBytes + 16;  // 2320: skewed element block
constexpr int kJElemBytes = 8 * 9 * 8;            // 576

template <int EPT, int STAGES>
struct GradMultSmem {
  alignas(16) unsigned char c[STAGES][EPT * kCElemSmem];
  alignas(16) unsigned char j[STAGES][EPT * kJElemBytes];
  alignas(8) uint64_t full[STAGES];
};

template <int EPT, int STAGES, int MODE, bool ESS>
__global__ void __launch_bounds__(EPT * 8) k_grad_mult_pa(const double* __restrict__ matgrad,
                                                          const double* __restrict__ jac,
                                                          const double* __restrict__ x, double* __restrict__ y,
                                                          ElemIO io, long nelems, double dt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GradMultSmem<EPT, STAGES>& sm = *reinterpret_cast<GradMultSmem<EPT, STAGES>*>(smem_raw);
  const int tid = threadIdx.x;
  const int lane = tid & 7;          // node / quadrature point (lexicographic) inside the element
  const int el = tid >> 3;           // element slot inside the tile
  const long ntiles = (nelems + EPT - 1) / EPT;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&sm.full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  // producer: warp 0, one lane per element of the tile
  auto issue = [&](long tile, int s) {
    if (tid < 32) {
      const long e0 = tile * EPT;
      const int ne = (int)min((long)EPT, nelems - e0);
      if (tid == 0) mbar_arrive_expect_tx(&sm.full[s], (uint32_t)ne * (2 * kCHalfBytes + kJElemBytes));
      __syncwarp();
      for (int i = tid; i < ne; i += 32) {
        const unsigned char* gc = reinterpret_cast<const unsigned char*>(matgrad + (e0 + i) * 288);
        unsigned char* sc = &sm.c[s][i * kCElemSmem];
        bulk_g2s(sc, gc, kCHalfBytes, &sm.full[s]);
        bulk_g2s(sc + kCHalfBytes + 16, gc + kCHalfBytes, kCHalfBytes, &sm.full[s]);
        bulk_g2s(&sm.j[s][i * kJElemBytes], jac + (e0 + i) * 72, kJElemBytes, &sm.full[s]);
      }
    }
  };

  // prologue: fill the ring
  {
    long t = blockIdx.x;
    for (int s = 0; s < STAGES; ++s, t += gridDim.x)
      if (t < ntiles) issue(t, s);
  }

  // register-prefetched nodal data of the next tile (hides the e2n -> x dependent loads)
  auto load_nodal = [&](long tile, long& nid, unsigned& msk, double& x0, double& x1, double& x2) {
    const long e = tile * EPT + el;
    nid = -1; msk = 0; x0 = x1 = x2 = 0.0;
    if (tile < ntiles && e < nelems) {
      if (MODE == LVEC) {
        nid = io.e2n[e * 8 + lex_to_native(lane)];
        if (ESS) msk = io.essmask[nid];
        x0 = (msk & 1) ? 0.0 : x[nid];
        x1 = (msk & 2) ? 0.0 : x[io.nnodes + nid];
        x2 = (msk & 4) ? 0.0 : x[2 * io.nnodes + nid];
      } else {
        nid = e * 24 + lex_to_native(lane);
        x0 = x[nid]; x1 = x[nid + 8]; x2 = x[nid + 16];
      }
    }
  };

  long nid_n; unsigned msk_n; double xn0, xn1, xn2;
  load_nodal(blockIdx.x, nid_n, msk_n, xn0, xn1, xn2);

  int s = 0;
  uint32_t phase = 0;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long nid = nid_n; const unsigned msk = msk_n;
    const double u0 = xn0, u1 = xn1, u2 = xn2;
    load_nodal(tile + gridDim.x, nid_n, msk_n, xn0, xn1, xn2);

    // reference-space velocity gradient at this lane's quadrature point: d(i,s) = dv_i/dxi_s
    double d00, d01, d02, d10, d11, d12, d20, d21, d22;
    nodal_to_qp_grad(u0, lane, d00, d01, d02);
    nodal_to_qp_grad(u1, lane, d10, d11, d12);
    nodal_to_qp_grad(u2, lane, d20, d21, d22);

    mbar_wait(&sm.full[s], phase);

    const long e = tile * EPT + el;
    const bool active = e < nelems;
    double t00 = 0, t01 = 0, t02 = 0, t10 = 0, t11 = 0, t12 = 0, t20 = 0, t21 = 0, t22 = 0;  // T(j,k)
    if (active) {
      const double* Jq = reinterpret_cast<const double*>(&sm.j[s][el * kJElemBytes]) + lane * 9;
      double J[9], adj[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) J[i] = Jq[i];
      const double det = adjugate(J, adj);
      const double c = dt * kWq / det;
      // det * dv_i/dx_t = sum_s d(i,s) adj(s,t)
      const double g00 = d00 * adj[0] + d01 * adj[3] + d02 * adj[6];
      const double g01 = d00 * adj[1] + d01 * adj[4] + d02 * adj[7];
      const double g02 = d00 * adj[2] + d01 * adj[5] + d02 * adj[8];
      const double g10 = d10 * adj[0] + d11 * adj[3] + d12 * adj[6];
      const double g11 = d10 * adj[1] + d11 * adj[4] + d12 * adj[7];
      const double g12 = d10 * adj[2] + d11 * adj[5] + d12 * adj[8];
      const double g20 = d20 * adj[0] + d21 * adj[3] + d22 * adj[6];
      const double g21 = d20 * adj[1] + d21 * adj[4] + d22 * adj[7];
      const double g22 = d20 * adj[2] + d21 * adj[5] + d22 * adj[8];
      // Voigt strain-rate-like vector (engineering shear), scaled by dt W / det
      const double eps[6] = {c * g00, c * g11, c * g22, c * (g12 + g21), c * (g02 + g20), c * (g01 + g10)};
      // S_I = sum_J K(I,J) eps_J, K(I,J) at [J*6 + I]; stream the 36 doubles in memory order
      const unsigned char* cb = &sm.c[s][el * kCElemSmem + (lane >> 2) * (kCHalfBytes + 16) + (lane & 3) * 288];
      const double2* C2 = reinterpret_cast<const double2*>(cb);
      double S[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int Jc = 0; Jc < 6; ++Jc) {
        const double2 a = C2[Jc * 3 + 0], b = C2[Jc * 3 + 1], cc = C2[Jc * 3 + 2];
        S[0] += a.x * eps[Jc]; S[1] += a.y * eps[Jc];
        S[2] += b.x * eps[Jc]; S[3] += b.y * eps[Jc];
        S[4] += cc.x * eps[Jc]; S[5] += cc.y * eps[Jc];
      }
      // T(j,k) = sum_p adj(j,p) S(p,k); S symmetric 3x3 from Voigt (0,1,2,3=yz,4=xz,5=xy)
      t00 = adj[0] * S[0] + adj[1] * S[5] + adj[2] * S[4];
      t01 = adj[0] * S[5] + adj[1] * S[1] + adj[2] * S[3];
      t02 = adj[0] * S[4] + adj[1] * S[3] + adj[2] * S[2];
      t10 = adj[3] * S[0] + adj[4] * S[5] + adj[5] * S[4];
      t11 = adj[3] * S[5] + adj[4] * S[1] + adj[5] * S[3];
      t12 = adj[3] * S[4] + adj[4] * S[3] + adj[5] * S[2];
      t20 = adj[6] * S[0] + adj[7] * S[5] + adj[8] * S[4];
      t21 = adj[6] * S[5] + adj[7] * S[1] + adj[8] * S[3];
      t22 = adj[6] * S[4] + adj[7] * S[3] + adj[8] * S[2];
    }
    // all lanes of the warp take part in the butterflies
    const double y0 = qp_grad_to_nodal(t00, t10, t20, lane);
    const double y1 = qp_grad_to_nodal(t01, t11, t21, lane);
    const double y2 = qp_grad_to_nodal(t02, t12, t22, lane);
    if (active) {
      if (MODE == LVEC) {
        if (!(msk & 1)) red_add_f64(&y[nid], y0);
        if (!(msk & 2)) red_add_f64(&y[io.nnodes + nid], y1);
        if (!(msk & 4)) red_add_f64(&y[2 * io.nnodes + nid], y2);
      } else {
        y[nid] += y0; y[nid + 8] += y1; y[nid + 16] += y2;
      }
    }
    // release the stage and refill it with the tile STAGES grid-strides ahead
    __syncthreads();
    const long tnext = tile + (long)STAGES * gridDim.x;
    if (tnext < ntiles) issue(tnext, s);
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }
}

// ------------------------------------------------------------------------------------------
// K2, warp-private pipelines (the default).  Same arithmetic as k_grad_mult_pa, but every warp owns
// its own STAGES-deep ring of 4-element sub-tiles (4 x 2320 B skewed matGrad + 2304 B J per stage)
// with its own mbarriers and refills a stage itself right after its last shared-memory read
// (__syncwarp + proxy fence), so there is no CTA-wide barrier anywhere in the loop; the element
// connectivity is prefetched two tiles ahead and the nodal values one tile ahead, so the
// e2n -> x dependent-load chain never stalls the contraction.
// ------------------------------------------------------------------------------------------
constexpr int kWarpStageBytes = 4 * kCElemSmem + 4 * kJElemBytes;  // 11584
constexpr int kWarpStageBytesJX = 4 * kCElemSmem;                  // 9280: matGrad only

// JX = true (LVEC only): the Jacobians are not streamed from HBM (576 B/element, 20 % of the operand bytes)
// but rebuilt in registers from the end-of-step nodal coordinates `xend` (an L-vector, mostly L1/L2 hits)
// with the same butterfly k_jacobians uses, i.e. bit-identical to the stored J.
template <int NW, int STAGES, int MODE, bool ESS, bool JX>
__global__ void __launch_bounds__(NW * 32) k_grad_mult_pa_w(const double* __restrict__ matgrad,
                                                            const double* __restrict__ jac,
                                                            const double* __restrict__ x, double* __restrict__ y,
                                                            ElemIO io, long nelems, double dt,
                                                            double* __restrict__ dot_accum,
                                                            const double* __restrict__ xend, int l2_hint) {
  static_assert(!JX || MODE == LVEC, "coordinate-rebuilt Jacobians need the L-vector connectivity");
  const uint64_t pol = l2_hint ? l2_policy_evict_first() : 0ull;
  constexpr int SB = JX ? kWarpStageBytesJX : kWarpStageBytes;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), l32 = threadIdx.x & 31;  // warp-uniform for the compiler
  const int lane = l32 & 7;  // node / quadrature point
  const int el = l32 >> 3;   // element slot in the warp's sub-tile
  unsigned char* ring = smem_raw + (size_t)w * STAGES * SB;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * STAGES * SB) + w * STAGES;
  const long nwt = (nelems + 3) >> 2;                 // warp tiles
  const long stride = (long)gridDim.x * NW;
  const long wt0 = (long)blockIdx.x * NW + w;

  if (l32 == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  auto issue = [&](long wt, int s) {
    const long e0 = wt << 2;
    const int ne = (int)min(4L, nelems - e0);
    unsigned char* sc = ring + s * SB;
    if (l32 == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)ne * (2 * kCHalfBytes + (JX ? 0 : kJElemBytes)));
    __syncwarp();
    if (l32 < 2 * ne) {
      const int i = l32 >> 1, h = l32 & 1;
      unsigned char* dst = sc + i * kCElemSmem + h * (kCHalfBytes + 16);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(matgrad + (e0 + i) * 288) + h * kCHalfBytes;
      if (l2_hint) bulk_g2s_hint(dst, src, kCHalfBytes, &full[s], pol);
      else bulk_g2s(dst, src, kCHalfBytes, &full[s]);
    } else if (!JX && l32 == 8) {
      if (l2_hint) bulk_g2s_hint(sc + 4 * kCElemSmem, jac + e0 * 72, (uint32_t)ne * kJElemBytes, &full[s], pol);
      else bulk_g2s(sc + 4 * kCElemSmem, jac + e0 * 72, (uint32_t)ne * kJElemBytes, &full[s]);
    }
  };
  {
    long t = wt0;
    for (int s = 0; s < STAGES; ++s, t += stride)
      if (t < nwt) issue(t, s);
  }

  // Connectivity is kept as the raw 32-bit value until it is used one iteration later, so that nothing
  // (not even a sign extension) depends on the load in the iteration that issues it.
  auto load_nid = [&](long wt) -> int {
    const long e = (wt << 2) + el;
    if (wt >= nwt || e >= nelems) return -1;
    if (MODE == LVEC) return io.e2n[e * 8 + lex_to_native(lane)];
    return 0;  // EVEC: the offset is recomputed from the tile index
  };
  auto evec_off = [&](long wt) -> long { return ((wt << 2) + el) * 24 + lex_to_native(lane); };
  auto load_x = [&](int nid, long wt, unsigned& msk, double& x0, double& x1, double& x2, double& c0, double& c1,
                    double& c2) {
    msk = 0; x0 = x1 = x2 = 0.0; c0 = c1 = c2 = 0.0;
    if (nid < 0) return;
    if (MODE == LVEC) {
      if (ESS) msk = io.essmask[nid];
      x0 = x[nid]; x1 = x[io.nnodes + nid]; x2 = x[2 * io.nnodes + nid];
      if (JX) { c0 = xend[nid]; c1 = xend[io.nnodes + nid]; c2 = xend[2 * io.nnodes + nid]; }
    } else {
      const long o = evec_off(wt);
      x0 = x[o]; x1 = x[o + 8]; x2 = x[o + 16];
    }
  };

  int nid_c = load_nid(wt0), nid_n = load_nid(wt0 + stride);
  unsigned msk_c; double xc0, xc1, xc2, cc0, cc1, cc2;
  load_x(nid_c, wt0, msk_c, xc0, xc1, xc2, cc0, cc1, cc2);

  int s = 0;
  uint32_t phase = 0;
  double xdoty = 0.0;  // sum over this lane's (node, element) pairs of x . (K_e x_e): assembles to x^T K x
  for (long wt = wt0; wt < nwt; wt += stride) {
    // prefetch: connectivity two tiles ahead, nodal values one tile ahead
    const int nid_n2 = load_nid(wt + 2 * stride);
    unsigned msk_n; double xn0, xn1, xn2, cn0, cn1, cn2;
    load_x(nid_n, wt + stride, msk_n, xn0, xn1, xn2, cn0, cn1, cn2);

    const double u0 = (msk_c & 1) ? 0.0 : xc0, u1 = (msk_c & 2) ? 0.0 : xc1, u2 = (msk_c & 4) ? 0.0 : xc2;
    double d00, d01, d02, d10, d11, d12, d20, d21, d22;
    nodal_to_qp_grad(u0, lane, d00, d01, d02);
    nodal_to_qp_grad(u1, lane, d10, d11, d12);
    nodal_to_qp_grad(u2, lane, d20, d21, d22);
    double J[9];
    if (JX) {
      nodal_to_qp_grad(cc0, lane, J[0], J[3], J[6]);
      nodal_to_qp_grad(cc1, lane, J[1], J[4], J[7]);
      nodal_to_qp_grad(cc2, lane, J[2], J[5], J[8]);
    }

    mbar_wait(&full[s], phase);

    const bool active = nid_c >= 0;
    double t00 = 0, t01 = 0, t02 = 0, t10 = 0, t11 = 0, t12 = 0, t20 = 0, t21 = 0, t22 = 0;
    if (active) {
      const unsigned char* sc = ring + s * SB;
      double adj[9];
      if (!JX) {
        const double* Jq = reinterpret_cast<const double*>(sc + 4 * kCElemSmem + el * kJElemBytes) + lane * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) J[i] = Jq[i];
      }
      const double det = adjugate(J, adj);
      const double c = dt * kWq / det;
      const double g00 = d00 * adj[0] + d01 * adj[3] + d02 * adj[6];
      const double g01 = d00 * adj[1] + d01 * adj[4] + d02 * adj[7];
      const double g02 = d00 * adj[2] + d01 * adj[5] + d02 * adj[8];
      const double g10 = d10 * adj[0] + d11 * adj[3] + d12 * adj[6];
      const double g11 = d10 * adj[1] + d11 * adj[4] + d12 * adj[7];
      const double g12 = d10 * adj[2] + d11 * adj[5] + d12 * adj[8];
      const double g20 = d20 * adj[0] + d21 * adj[3] + d22 * adj[6];
      const double g21 = d20 * adj[1] + d21 * adj[4] + d22 * adj[7];
      const double g22 = d20 * adj[2] + d21 * adj[5] + d22 * adj[8];
      const double eps[6] = {c * g00, c * g11, c * g22, c * (g12 + g21), c * (g02 + g20), c * (g01 + g10)};
      const double2* C2 = reinterpret_cast<const double2*>(sc + el * kCElemSmem + (lane >> 2) * (kCHalfBytes + 16) + (lane & 3) * 288);
      double S[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int Jc = 0; Jc < 6; ++Jc) {
        const double2 a = C2[Jc * 3 + 0], b = C2[Jc * 3 + 1], cc = C2[Jc * 3 + 2];
        S[0] += a.x * eps[Jc]; S[1] += a.y * eps[Jc];
        S[2] += b.x * eps[Jc]; S[3] += b.y * eps[Jc];
        S[4] += cc.x * eps[Jc]; S[5] += cc.y * eps[Jc];
      }
      t00 = adj[0] * S[0] + adj[1] * S[5] + adj[2] * S[4];
      t01 = adj[0] * S[5] + adj[1] * S[1] + adj[2] * S[3];
      t02 = adj[0] * S[4] + adj[1] * S[3] + adj[2] * S[2];
      t10 = adj[3] * S[0] + adj[4] * S[5] + adj[5] * S[4];
      t11 = adj[3] * S[5] + adj[4] * S[1] + adj[5] * S[3];
      t12 = adj[3] * S[4] + adj[4] * S[3] + adj[5] * S[2];
      t20 = adj[6] * S[0] + adj[7] * S[5] + adj[8] * S[4];
      t21 = adj[6] * S[5] + adj[7] * S[1] + adj[8] * S[3];
      t22 = adj[6] * S[4] + adj[7] * S[3] + adj[8] * S[2];
    }
    // this stage's shared memory is dead: refill it before the output butterflies
    __syncwarp();
    {
      const long tnext = wt + (long)STAGES * stride;
      if (tnext < nwt) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tnext, s);
      }
    }
    const double y0 = qp_grad_to_nodal(t00, t10, t20, lane);
    const double y1 = qp_grad_to_nodal(t01, t11, t21, lane);
    const double y2 = qp_grad_to_nodal(t02, t12, t22, lane);
    if (active) {
      xdoty += u0 * y0 + u1 * y1 + u2 * y2;
      if (MODE == LVEC) {
        if (!(msk_c & 1)) red_add_f64(&y[nid_c], y0);
        if (!(msk_c & 2)) red_add_f64(&y[io.nnodes + nid_c], y1);
        if (!(msk_c & 4)) red_add_f64(&y[2 * io.nnodes + nid_c], y2);
      } else {
        const long o = evec_off(wt);
        y[o] += y0; y[o + 8] += y1; y[o + 16] += y2;
      }
    }
    nid_c = nid_n; nid_n = nid_n2;
    msk_c = msk_n; xc0 = xn0; xc1 = xn1; xc2 = xn2;
    cc0 = cn0; cc1 = cn1; cc2 = cn2;
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }
  if (dot_accum) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) xdoty += __shfl_xor_sync(kFull, xdoty, m);
    if (l32 == 0) red_add_f64(dot_accum, xdoty);
  }
}

// ------------------------------------------------------------------------------------------
// K2, compact-tangent variant (the default of the fused L-vector PA path for cubic crystals).  The operand is the
// 32-double record K1 writes in layout kTangentCompact (mat::compact_apply: 25 + 6 + 1 values instead of 36), densely
// packed at 256 B per point in a context-owned array, so 11 % fewer bytes leave HBM; J is rebuilt from the end coordinates as
// in the JX kernels.  Each warp stage is two 32-row x 128-B boxes fetched by TWO tiled TMA copies
// (cp.async.bulk.tensor.2d, CUtensorMap over [npts][32 doubles]) with the 128-byte hardware
// swizzle: row r's 16-byte chunk c lands at chunk (c ^ (r & 7)), which makes the per-lane LDS.128 stream (lane =
// row) bank-conflict free without padding or per-point copies.
// ------------------------------------------------------------------------------------------
constexpr int kBoxBytes = 32 * 128;               // 32 points x 16 doubles
constexpr int kWarpStageBytesC = 2 * kBoxBytes;   // 8192

// HALO variant: the operator apply of one z-slab of a multi-GPU run WITH its interface-plane exchange and the all-reduce
// of the fused CG denominator, in one kernel over NVLink peer memory (the role ParNonlinearForm's P^T ... P plays around
// PANonlinearMechOperatorGradExt::TMult, src/mechanics_operator_ext.cpp:149,157).  The warp tiles of the two boundary
// element layers are processed FIRST and counted off; the first `ncomm` CTAs do no element work: they wait for that
// count, push this rank's partial sums on the interface planes into the neighbours' mailboxes, wait for theirs and add
// them -- while the other CTAs stream the interior layers -- and CTA 0 finally all-reduces x^T K x over the ranks once
// every compute CTA has contributed.  Nothing of the exchange is left on the critical path but the last flag wait.
struct HaloArgs {
  double* mb;                       // this rank's mailbox, the lower / upper neighbour's (nullptr at the ends)
  double* lo;
  double* hi;
  exab_p2p::PeerTable peers;
  int rank, nranks, ncomm;
  long nn, plane;
  long nb_lo, nb_hi, hi_start;      // boundary warp tiles: [0, nb_lo) and [hi_start, hi_start + nb_hi)
  unsigned long long seq_halo, seq_scal;
  unsigned long long* counters;     // [0] boundary tiles done, [1] compute CTAs done (both cumulative over launches)
  unsigned long long tiles_target, ctas_target;
};

// register cap: 12 resident warps per SM (6 CTAs of 2 warps, 3 of 4, 12 of 1)
template <int NW, int STAGES, bool ESS, bool HALO = false, bool EVOUT = false>
__global__ void __launch_bounds__(NW * 32, 12 / NW) k_grad_mult_pa_c(const __grid_constant__ CUtensorMap tmap,
                                                            const double* __restrict__ x, double* __restrict__ y,
                                                            ElemIO io, long nelems, double dt,
                                                            double* __restrict__ dot_accum,
                                                            const double* __restrict__ xend,
                                                            const __grid_constant__ HaloArgs h) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ncomm = HALO ? h.ncomm : 0;
  if (HALO && (int)blockIdx.x < ncomm) {
    __shared__ bool s_ok;
    if (threadIdx.x == 0)
      while (exab_p2p::ld_acquire_gpu(&h.counters[0]) < h.tiles_target) __nanosleep(200);
    __syncthreads();
    exab_p2p::halo_exchange_blocks(y, h.mb, h.lo, h.hi, h.nn, h.plane, h.seq_halo, h.peers.spin_limit, blockIdx.x,
                                   (unsigned)ncomm, &s_ok);
    if (blockIdx.x == 0 && dot_accum) {
      if (threadIdx.x == 0)
        while (exab_p2p::ld_acquire_gpu(&h.counters[1]) < h.ctas_target) __nanosleep(200);
      __syncthreads();
      if (threadIdx.x < 32) exab_p2p::warp_allreduce_p2p(dot_accum, h.peers, h.rank, h.nranks, 1, h.seq_scal, threadIdx.x);
    }
    return;
  }
  // warp index through a broadcast: tells the compiler it is warp-uniform (uniform loop bounds, no re-convergence
  // barriers around the shuffles)
  const int w = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), l32 = threadIdx.x & 31;
  const int lane = l32 & 7;  // node / quadrature point
  const int el = l32 >> 3;   // element slot in the warp's sub-tile
  // the swizzle pattern is a function of the shared-memory address: 1024-byte aligned stages
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ring = base + (size_t)w * STAGES * kWarpStageBytesC;
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)NW * STAGES * kWarpStageBytesC) + w * STAGES;
  const long nwt = (nelems + 3) >> 2;  // warp tiles of 4 elements = 32 points
  const long stride = (long)(gridDim.x - ncomm) * NW;
  const long wt0 = (long)(blockIdx.x - ncomm) * NW + w;   // first ITERATION index of this warp
  const uint64_t pol = l2_policy_evict_first();
  // iteration index -> warp tile: identity, or boundary layers first (HALO)
  const long nb_lo = HALO ? h.nb_lo : 0, nb_hi = HALO ? h.nb_hi : 0, nb = nb_lo + nb_hi;
  auto tile_of = [&](long it) -> long {
    if (!HALO) return it;
    if (it < nb_lo) return it;
    if (it < nb) return h.hi_start + (it - nb_lo);
    return it - nb_hi;
  };

  if (l32 == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  auto issue = [&](long wt, int s) {
    if (l32 == 0) {
      unsigned char* sc = ring + s * kWarpStageBytesC;
      mbar_arrive_expect_tx(&full[s], (uint32_t)kWarpStageBytesC);  // out-of-range rows are zero-filled and counted
      tma_load_2d(sc, &tmap, 0, (int)(wt << 5), &full[s], pol);
      tma_load_2d(sc + kBoxBytes, &tmap, 16, (int)(wt << 5), &full[s], pol);
    }
  };
  {
    long t = wt0;
    for (int s = 0; s < STAGES; ++s, t += stride)
      if (t < nwt) issue(tile_of(t), s);
  }

  // With ESS the connectivity entries carry the node's essential-dof bits in bits 28..30 (kEssShift; written by
  // k_fold_ess_mask when the mask is set): one scattered byte gather less per lane and tile -- the kernel is bound by
  // the L1TEX data-pipe wavefronts (shuffles + LDS + gathers + reds: 87 % of peak in ncu), not by instruction issue.
  auto load_nid = [&](long it) -> int {
    if (it >= nwt) return -1;
    const long e = (tile_of(it) << 2) + el;
    if (e >= nelems) return -1;
    return io.e2n[e * 8 + lex_to_native(lane)];
  };
  auto load_x = [&](int& nid, unsigned& msk, double& x0, double& x1, double& x2, double& c0, double& c1, double& c2) {
    msk = 0; x0 = x1 = x2 = 0.0; c0 = c1 = c2 = 0.0;
    if (nid < 0) return;
    if (ESS) { msk = (unsigned)nid >> kEssShift; nid &= kEssNodeMask; }
    x0 = x[nid]; x1 = x[io.nnodes + nid]; x2 = x[2 * io.nnodes + nid];
    c0 = xend[nid]; c1 = xend[io.nnodes + nid]; c2 = xend[2 * io.nnodes + nid];
  };

  int nid_c = load_nid(wt0), nid_n = load_nid(wt0 + stride);
  unsigned msk_c; double xc0, xc1, xc2, cc0, cc1, cc2;
  load_x(nid_c, msk_c, xc0, xc1, xc2, cc0, cc1, cc2);

  // this lane's row of the two boxes and its swizzle key
  const int row_off = l32 * 128, key = (l32 & 7) << 4;
  const LaneSigns sg = lane_signs(lane);

  int s = 0;
  uint32_t phase = 0;
  double xdoty = 0.0;
  for (long wt = wt0; wt < nwt; wt += stride) {
    const int nid_n2 = load_nid(wt + 2 * stride);
    unsigned msk_n; double xn0, xn1, xn2, cn0, cn1, cn2;
    load_x(nid_n, msk_n, xn0, xn1, xn2, cn0, cn1, cn2);

    const double u0 = (msk_c & 1) ? 0.0 : xc0, u1 = (msk_c & 2) ? 0.0 : xc1, u2 = (msk_c & 4) ? 0.0 : xc2;
    double d00, d01, d02, d10, d11, d12, d20, d21, d22;
    nodal_to_qp_grad(u0, sg, d00, d01, d02);
    nodal_to_qp_grad(u1, sg, d10, d11, d12);
    nodal_to_qp_grad(u2, sg, d20, d21, d22);
    double J[9];
    nodal_to_qp_grad(cc0, sg, J[0], J[3], J[6]);
    nodal_to_qp_grad(cc1, sg, J[1], J[4], J[7]);
    nodal_to_qp_grad(cc2, sg, J[2], J[5], J[8]);

    mbar_wait(&full[s], phase);

    // No branch around the arithmetic: an inactive 8-lane group (only in the last, partial tile) works on
    // zero-filled rows and zero coordinates; whatever it produces (inf/NaN from det = 0) stays inside the group
    // and is never stored.  Uniform control flow keeps the shuffles free of re-convergence overhead.
    const bool active = nid_c >= 0;
    double t00, t01, t02, t10, t11, t12, t20, t21, t22;
    {
      const unsigned char* sc = ring + s * kWarpStageBytesC + row_off;
      double rec[32];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const double2 v = *reinterpret_cast<const double2*>(sc + (c >> 3) * kBoxBytes + (((c & 7) << 4) ^ key));
        rec[2 * c] = v.x;
        rec[2 * c + 1] = v.y;
      }
      double adj[9];
      const double det = adjugate(J, adj);
      const double c = dt * kWq / det;
      const double g00 = d00 * adj[0] + d01 * adj[3] + d02 * adj[6];
      const double g01 = d00 * adj[1] + d01 * adj[4] + d02 * adj[7];
      const double g02 = d00 * adj[2] + d01 * adj[5] + d02 * adj[8];
      const double g10 = d10 * adj[0] + d11 * adj[3] + d12 * adj[6];
      const double g11 = d10 * adj[1] + d11 * adj[4] + d12 * adj[7];
      const double g12 = d10 * adj[2] + d11 * adj[5] + d12 * adj[8];
      const double g20 = d20 * adj[0] + d21 * adj[3] + d22 * adj[6];
      const double g21 = d20 * adj[1] + d21 * adj[4] + d22 * adj[7];
      const double g22 = d20 * adj[2] + d21 * adj[5] + d22 * adj[8];
      const double eps[6] = {c * g00, c * g11, c * g22, c * (g12 + g21), c * (g02 + g20), c * (g01 + g10)};
      double S[6];
      mat::compact_apply(rec, eps, S);
      t00 = adj[0] * S[0] + adj[1] * S[5] + adj[2] * S[4];
      t01 = adj[0] * S[5] + adj[1] * S[1] + adj[2] * S[3];
      t02 = adj[0] * S[4] + adj[1] * S[3] + adj[2] * S[2];
      t10 = adj[3] * S[0] + adj[4] * S[5] + adj[5] * S[4];
      t11 = adj[3] * S[5] + adj[4] * S[1] + adj[5] * S[3];
      t12 = adj[3] * S[4] + adj[4] * S[3] + adj[5] * S[2];
      t20 = adj[6] * S[0] + adj[7] * S[5] + adj[8] * S[4];
      t21 = adj[6] * S[5] + adj[7] * S[1] + adj[8] * S[3];
      t22 = adj[6] * S[4] + adj[7] * S[3] + adj[8] * S[2];
    }
    // this stage's shared memory is dead: refill it before the output butterflies
    __syncwarp();
    {
      const long tnext = wt + (long)STAGES * stride;
      if (tnext < nwt) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tile_of(tnext), s);
      }
    }
    const double y0 = qp_grad_to_nodal(t00, t10, t20, sg);
    const double y1 = qp_grad_to_nodal(t01, t11, t21, sg);
    const double y2 = qp_grad_to_nodal(t02, t12, t22, sg);
    if (active) {
      if (EVOUT) {
        // owner-computes (deterministic) mode: the element's contributions go to an E-vector Y(a, i, e) with plain
        // stores; k_evec_to_lvec sums them per node in a fixed order afterwards
        const long o = ((tile_of(wt) << 2) + el) * 24 + lex_to_native(lane);
        y[o] = y0; y[o + 8] = y1; y[o + 16] = y2;
      } else {
        xdoty += u0 * y0 + u1 * y1 + u2 * y2;
        if (!(msk_c & 1)) red_add_f64(&y[nid_c], y0);
        if (!(msk_c & 2)) red_add_f64(&y[io.nnodes + nid_c], y1);
        if (!(msk_c & 4)) red_add_f64(&y[2 * io.nnodes + nid_c], y2);
      }
    }
    if (HALO && wt < nb) {  // a boundary-layer tile is complete: count it off for the exchange CTAs
      __threadfence();
      __syncwarp();
      if (l32 == 0) atomicAdd(&h.counters[0], 1ull);
    }
    nid_c = nid_n; nid_n = nid_n2;
    msk_c = msk_n; xc0 = xn0; xc1 = xn1; xc2 = xn2;
    cc0 = cn0; cc1 = cn1; cc2 = cn2;
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }
  if (dot_accum) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) xdoty += __shfl_xor_sync(kFull, xdoty, m);
    if (l32 == 0) red_add_f64(dot_accum, xdoty);
  }
  if (HALO) {  // this CTA's share of x^T K x is in: CTA 0 all-reduces once every compute CTA has said so
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(&h.counters[1], 1ull);
  }
}

// ------------------------------------------------------------------------------------------
// Residual action: y_{a,k} += sum_q W_q [adj(J) sigma](j,k) G(a,j,q)
// Replaces ExaNLFIntegrator::AssemblePA (three passes, src/mechanics_integrators.cpp:160-314)
// + AddMultPA (:518-557) [+ restriction^T and essential-dof zeroing of MultVec,
// src/mechanics_operator_ext.cpp:176-202].  BBAR selects ICExaNLFIntegrator::AddMultPA
// (:1961-2088) with the element-average shape gradients computed in-kernel (:1895-1952).
// One thread per quadrature point, 8 lanes per element; plain coalescable global loads (this
// kernel runs once per Newton residual, not in the PCG loop).
// ------------------------------------------------------------------------------------------
template <int MODE, bool BBAR>
__global__ void __launch_bounds__(256) k_residual(const double* __restrict__ stress, const double* __restrict__ jac,
                                                  double* __restrict__ y, ElemIO io, long nelems) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  double J[9], adj[9], S[6], det = 1.0;
  if (active) {
    const double* Jq = jac + (e * 8 + lane) * 9;
    const double* Sq = stress + (e * 8 + lane) * 6;
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = Jq[i];
#pragma unroll
    for (int i = 0; i < 6; ++i) S[i] = Sq[i];
    det = adjugate(J, adj);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) adj[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) S[i] = 0.0;
  }
  double y0, y1, y2;
  if (!BBAR) {
    const double w = active ? kWq : 0.0;
    const double t00 = w * (adj[0] * S[0] + adj[1] * S[5] + adj[2] * S[4]);
    const double t01 = w * (adj[0] * S[5] + adj[1] * S[1] + adj[2] * S[3]);
    const double t02 = w * (adj[0] * S[4] + adj[1] * S[3] + adj[2] * S[2]);
    const double t10 = w * (adj[3] * S[0] + adj[4] * S[5] + adj[5] * S[4]);
    const double t11 = w * (adj[3] * S[5] + adj[4] * S[1] + adj[5] * S[3]);
    const double t12 = w * (adj[3] * S[4] + adj[4] * S[3] + adj[5] * S[2]);
    const double t20 = w * (adj[6] * S[0] + adj[7] * S[5] + adj[8] * S[4]);
    const double t21 = w * (adj[6] * S[5] + adj[7] * S[1] + adj[8] * S[3]);
    const double t22 = w * (adj[6] * S[4] + adj[7] * S[3] + adj[8] * S[2]);
    y0 = qp_grad_to_nodal(t00, t10, t20, lane);
    y1 = qp_grad_to_nodal(t01, t11, t21, lane);
    y2 = qp_grad_to_nodal(t02, t12, t22, lane);
  } else {
    // B-bar: y_{a,I} = sum_q W detJ [ b_p(a,q) dev-part + eDS_I(a) * mean-part ]
    //   Bbar^T sigma = b_p s_{pI} + (1/3)(eDS_I - b_I) tr(sigma)
    // => sum_q W detJ b_p (s_{pI} - tr/3 delta_{pI})   (plain operator on the deviatoric-shifted stress)
    //    + eDS_I(a) * sum_q W detJ tr(sigma)/3
    const double w = active ? kWq : 0.0;
    const double tr3 = (S[0] + S[1] + S[2]) * (1.0 / 3.0);
    const double D0 = S[0] - tr3, D1 = S[1] - tr3, D2 = S[2] - tr3;
    const double t00 = w * (adj[0] * D0 + adj[1] * S[5] + adj[2] * S[4]);
    const double t01 = w * (adj[0] * S[5] + adj[1] * D1 + adj[2] * S[3]);
    const double t02 = w * (adj[0] * S[4] + adj[1] * S[3] + adj[2] * D2);
    const double t10 = w * (adj[3] * D0 + adj[4] * S[5] + adj[5] * S[4]);
    const double t11 = w * (adj[3] * S[5] + adj[4] * D1 + adj[5] * S[3]);
    const double t12 = w * (adj[3] * S[4] + adj[4] * S[3] + adj[5] * D2);
    const double t20 = w * (adj[6] * D0 + adj[7] * S[5] + adj[8] * S[4]);
    const double t21 = w * (adj[6] * S[5] + adj[7] * D1 + adj[8] * S[3]);
    const double t22 = w * (adj[6] * S[4] + adj[7] * S[3] + adj[8] * D2);
    y0 = qp_grad_to_nodal(t00, t10, t20, lane);
    y1 = qp_grad_to_nodal(t01, t11, t21, lane);
    y2 = qp_grad_to_nodal(t02, t12, t22, lane);
    // element sums: volume, mean-stress moment, and eDS numerators sum_q W b_c(a,q)
    double vol = w * det, pm = w * det * tr3;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) { vol += shfl_xor_d(vol, m); pm += shfl_xor_d(pm, m); }
    const double e0 = qp_grad_to_nodal(w * adj[0], w * adj[3], w * adj[6], lane);
    const double e1 = qp_grad_to_nodal(w * adj[1], w * adj[4], w * adj[7], lane);
    const double e2 = qp_grad_to_nodal(w * adj[2], w * adj[5], w * adj[8], lane);
    const double f = active ? pm / vol : 0.0;
    y0 += e0 * f; y1 += e1 * f; y2 += e2 * f;
  }
  if (active) {
    if (MODE == LVEC) {
      const long nid = io.e2n[e * 8 + lex_to_native(lane)];
      const unsigned msk = io.essmask ? io.essmask[nid] : 0u;
      if (!(msk & 1)) red_add_f64(&y[nid], y0);
      if (!(msk & 2)) red_add_f64(&y[io.nnodes + nid], y1);
      if (!(msk & 4)) red_add_f64(&y[2 * io.nnodes + nid], y2);
    } else {
      const long o = e * 24 + lex_to_native(lane);
      y[o] += y0; y[o + 8] += y1; y[o + 16] += y2;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Diagonal of the gradient operator.  Replaces ExaNLFIntegrator::AssembleGradDiagonalPA
// (src/mechanics_integrators.cpp:625-748) [+ restriction^T, src/mechanics_operator_ext.cpp:95-123].
// lane = quadrature point; the 24 per-point partials are summed over the element with a
// reduce-scatter butterfly (12 + 6 + 3 doubles), leaving node `lane`'s 3 entries in each lane.
// ------------------------------------------------------------------------------------------
template <int MODE, bool COMPACT = false, bool BBAR = false>
__global__ void __launch_bounds__(256) k_grad_diag(const double* __restrict__ matgrad, const double* __restrict__ jac,
                                                   double* __restrict__ diag, ElemIO io, long nelems, double dt) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  double v[8][3];
#pragma unroll
  for (int a = 0; a < 8; ++a) v[a][0] = v[a][1] = v[a][2] = 0.0;
  double J[9], adj[9], K[36], det = 1.0;
  if (active) {
    const double* Jq = jac + (e * 8 + lane) * 9;
    const double* Kq = matgrad + (e * 8 + lane) * (COMPACT ? 32 : 36);
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = Jq[i];
    if (COMPACT) {
      double rec[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rec[i] = Kq[i];
      mat::compact_expand(rec, K);
    } else {
#pragma unroll
      for (int i = 0; i < 36; ++i) K[i] = Kq[i];
    }
    det = adjugate(J, adj);
  } else {
#pragma unroll
    for (int i = 0; i < 9; ++i) adj[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 36; ++i) K[i] = 0.0;
  }
  // B-bar (ICExaNLFIntegrator::AssembleGradDiagonalPA, src/mechanics_integrators.cpp:1607-1805): element-average shape
  // gradients eDS(a,c) = sum_q W b_c(a,q) / sum_q W detJ (src/mechanics_integrators.cpp:1895-1952); lane a of the
  // element's 8-lane group ends up holding node a's, every lane then fetches all eight by shuffle
  double eds[8][3];
  if (BBAR) {
    const double w = active ? kWq : 0.0;
    double vol = w * det;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) vol += shfl_xor_d(vol, m);
    const double ivol = active ? 1.0 / vol : 0.0;
    const double own0 = qp_grad_to_nodal(w * adj[0], w * adj[3], w * adj[6], lane) * ivol;
    const double own1 = qp_grad_to_nodal(w * adj[1], w * adj[4], w * adj[7], lane) * ivol;
    const double own2 = qp_grad_to_nodal(w * adj[2], w * adj[5], w * adj[8], lane) * ivol;
    const int grp = (threadIdx.x & 31) & ~7;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      eds[a][0] = __shfl_sync(kFull, own0, grp | a);
      eds[a][1] = __shfl_sync(kFull, own1, grp | a);
      eds[a][2] = __shfl_sync(kFull, own2, grp | a);
    }
  }
  if (active) {
    const double c = dt * kWq / det;
    const int vg[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      double g[3], b[3];
      shape_grad(a, lane, g);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) b[cc] = g[0] * adj[cc] + g[1] * adj[3 + cc] + g[2] * adj[6 + cc];
      if (!BBAR) {
#pragma unroll
        for (int I = 0; I < 3; ++I) {
          double s = 0.0;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            double w = 0.0;
#pragma unroll
            for (int r = 0; r < 3; ++r) w += b[r] * K[vg[I][r] * 6 + vg[I][p]];
            s += b[p] * w;
          }
          v[a][I] = c * s;
        }
      } else {
        // columns of the 6x3 B-bar block of node a (rows in Voigt order), b normalised by detJ
        const double idet = 1.0 / det, i3 = 1.0 / 3.0;
        const double bx = b[0] * idet, by = b[1] * idet, bz = b[2] * idet;
        const double b4 = i3 * (eds[a][0] - bx), b5 = b4 + bx;
        const double b6 = i3 * (eds[a][1] - by), b7 = b6 + by;
        const double b8 = i3 * (eds[a][2] - bz), b9 = b8 + bz;
        const double Bb[3][6] = {{b5, b4, b4, 0.0, bz, by}, {b6, b7, b6, bz, 0.0, bx}, {b8, b8, b9, by, bx, 0.0}};
        const double cw = dt * kWq * det;
#pragma unroll
        for (int I = 0; I < 3; ++I) {
          double s = 0.0;
#pragma unroll
          for (int R = 0; R < 6; ++R) {
            double w = 0.0;
#pragma unroll
            for (int S = 0; S < 6; ++S) w += K[S * 6 + R] * Bb[I][S];
            s += Bb[I][R] * w;
          }
          v[a][I] = cw * s;
        }
      }
    }
  }
  // reduce-scatter over the 8 lanes: after the three stages lane l holds sum_q v_q[l][:]
  double w4[4][3];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int I = 0; I < 3; ++I) {
      const bool hi = lane & 4;
      const double send = hi ? v[a][I] : v[a + 4][I];
      const double keep = hi ? v[a + 4][I] : v[a][I];
      w4[a][I] = keep + shfl_xor_d(send, 4);
    }
  double w2[2][3];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int I = 0; I < 3; ++I) {
      const bool hi = lane & 2;
      const double send = hi ? w4[a][I] : w4[a + 2][I];
      const double keep = hi ? w4[a + 2][I] : w4[a][I];
      w2[a][I] = keep + shfl_xor_d(send, 2);
    }
  double w1[3];
#pragma unroll
  for (int I = 0; I < 3; ++I) {
    const bool hi = lane & 1;
    const double send = hi ? w2[0][I] : w2[1][I];
    const double keep = hi ? w2[1][I] : w2[0][I];
    w1[I] = keep + shfl_xor_d(send, 1);
  }
  if (active) {
    if (MODE == LVEC) {
      const long nid = io.e2n[e * 8 + lex_to_native(lane)];
      red_add_f64(&diag[nid], w1[0]);
      red_add_f64(&diag[io.nnodes + nid], w1[1]);
      red_add_f64(&diag[2 * io.nnodes + nid], w1[2]);
    } else {
      const long o = e * 24 + lex_to_native(lane);
      diag[o] += w1[0]; diag[o + 8] += w1[1]; diag[o + 16] += w1[2];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Jacobians at the quadrature points from nodal coordinates: J(i,s,q,e) = dx_i/dxi_s.
// Replaces mesh->DeleteGeometricFactors()/GetGeometricFactors(JACOBIANS) + the transposing copy
// of NonlinearMechOperator::SetupJacobianTerms (src/mechanics_operator.cpp:350-391); with
// vel != nullptr the coordinates are x_beg + dt*vel, i.e. ExaModel::UpdateEndCoords
// (src/mechanics_model.cpp:445-481) fused in.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_jacobians(const double* __restrict__ xbeg, const double* __restrict__ vel,
                                                   double dt, double* __restrict__ jac, ElemIO io, long nelems,
                                                   double* __restrict__ xend) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  double c0 = 0, c1 = 0, c2 = 0;
  if (active) {
    const long nid = io.e2n[e * 8 + lex_to_native(lane)];
    c0 = xbeg[nid]; c1 = xbeg[io.nnodes + nid]; c2 = xbeg[2 * io.nnodes + nid];
    if (vel) { c0 += dt * vel[nid]; c1 += dt * vel[io.nnodes + nid]; c2 += dt * vel[2 * io.nnodes + nid]; }
    // the coordinates J is built from, kept for the PA gradient apply (every element sharing the node
    // stores the same bits)
    if (xend) { xend[nid] = c0; xend[io.nnodes + nid] = c1; xend[2 * io.nnodes + nid] = c2; }
  }
  double J[9];
  nodal_to_qp_grad(c0, lane, J[0], J[3], J[6]);
  nodal_to_qp_grad(c1, lane, J[1], J[4], J[7]);
  nodal_to_qp_grad(c2, lane, J[2], J[5], J[8]);
  if (active) {
    double* o = jac + (e * 8 + lane) * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) o[i] = J[i];
  }
}

// ------------------------------------------------------------------------------------------
// Element-matrix assembly (24x24 per element).  Replaces ExaNLFIntegrator::AssembleEA
// (src/mechanics_integrators.cpp:756-1017) and, with BBAR, ICExaNLFIntegrator::AssembleEA
// (:1195-1604; eDS from :1895-1952 computed in-kernel).
//   E(l+8I, k+8Kc, e) = sum_q c_q sum_{R,S} B_l(R,I) K(R,S) B_k(S,Kc),  memory [e*576 + col*24 + row]
// lane = column node k; it loops the element's 8 points and keeps its 24x3 column block in
// registers; rows/cols are in NATIVE node order like the reference.
// ------------------------------------------------------------------------------------------
// PRIV selects the context-private storage of the L-vector EA path: the 72 entries lane k owns (its 3 columns) are
// interleaved with the other lanes' in 16-byte chunks, value v = comp*24 + row_native at
// ea[e*576 + ((v >> 1) * 8 + lane) * 2 + (v & 1)], so that every LDG/STG.128 of the 8 lanes of an element covers 128
// contiguous bytes (the reference layout gives each lane a private 192-byte column: 32 sectors per warp request).
template <bool BBAR, bool PRIV = false>
__global__ void __launch_bounds__(128) k_assemble_ea(const double* __restrict__ matgrad, const double* __restrict__ jac,
                                                     double* __restrict__ ea, long nelems, double dt) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;  // lexicographic column node
  const long e = gt >> 3;
  if (e >= nelems) return;  // no shuffles in this kernel
  double acc[8][3][3];
#pragma unroll
  for (int l = 0; l < 8; ++l)
#pragma unroll
    for (int I = 0; I < 3; ++I) acc[l][I][0] = acc[l][I][1] = acc[l][I][2] = 0.0;
  // B-bar: element-average physical gradients eDS(a,c) and volume
  double eds[8][3];
  if (BBAR) {
    double vol = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) eds[a][0] = eds[a][1] = eds[a][2] = 0.0;
    for (int q = 0; q < 8; ++q) {
      double J[9], adj[9];
      const double* Jq = jac + (e * 8 + q) * 9;
#pragma unroll
      for (int i = 0; i < 9; ++i) J[i] = Jq[i];
      vol += kWq * adjugate(J, adj);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        double g[3];
        shape_grad(a, q, g);
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) eds[a][cc] += kWq * (g[0] * adj[cc] + g[1] * adj[3 + cc] + g[2] * adj[6 + cc]);
      }
    }
    const double iv = 1.0 / vol;
#pragma unroll
    for (int a = 0; a < 8; ++a) { eds[a][0] *= iv; eds[a][1] *= iv; eds[a][2] *= iv; }
  }
  for (int q = 0; q < 8; ++q) {
    double J[9], adj[9], K[36];
    const double* Jq = jac + (e * 8 + q) * 9;
    const double* Kq = matgrad + (e * 8 + q) * 36;
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = Jq[i];
#pragma unroll
    for (int i = 0; i < 36; ++i) K[i] = Kq[i];
    const double det = adjugate(J, adj);
    const double idet = 1.0 / det;
    const double c = dt * kWq * det;  // on true physical gradients
    auto bmat = [&](int a, double B[6][3]) {
      double g[3], b[3];
      shape_grad(a, q, g);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) b[cc] = idet * (g[0] * adj[cc] + g[1] * adj[3 + cc] + g[2] * adj[6 + cc]);
      double m0 = 0, m1 = 0, m2 = 0;
      if (BBAR) {
        m0 = (eds[a][0] - b[0]) * (1.0 / 3.0);
        m1 = (eds[a][1] - b[1]) * (1.0 / 3.0);
        m2 = (eds[a][2] - b[2]) * (1.0 / 3.0);
      }
      B[0][0] = m0 + b[0]; B[0][1] = m1;        B[0][2] = m2;
      B[1][0] = m0;        B[1][1] = m1 + b[1]; B[1][2] = m2;
      B[2][0] = m0;        B[2][1] = m1;        B[2][2] = m2 + b[2];
      B[3][0] = 0.0;       B[3][1] = b[2];      B[3][2] = b[1];
      B[4][0] = b[2];      B[4][1] = 0.0;       B[4][2] = b[0];
      B[5][0] = b[1];      B[5][1] = b[0];      B[5][2] = 0.0;
    };
    double Bk[6][3], KB[6][3];
    bmat(lane, Bk);
#pragma unroll
    for (int R = 0; R < 6; ++R)
#pragma unroll
      for (int Kc = 0; Kc < 3; ++Kc) {
        double s = 0.0;
#pragma unroll
        for (int S = 0; S < 6; ++S) s += K[S * 6 + R] * Bk[S][Kc];
        KB[R][Kc] = c * s;
      }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      double Bl[6][3];
      bmat(l, Bl);
#pragma unroll
      for (int I = 0; I < 3; ++I)
#pragma unroll
        for (int Kc = 0; Kc < 3; ++Kc) {
          double s = 0.0;
#pragma unroll
          for (int R = 0; R < 6; ++R) s += Bl[R][I] * KB[R][Kc];
          acc[l][I][Kc] += s;
        }
    }
  }
  const int kn = lex_to_native(lane);
#pragma unroll
  for (int Kc = 0; Kc < 3; ++Kc) {
    double* col = ea + e * 576 + (long)(kn + 8 * Kc) * 24;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const int ln = lex_to_native(l);
#pragma unroll
      for (int I = 0; I < 3; ++I) {
        if (PRIV) {
          const int v = Kc * 24 + ln + 8 * I;
          ea[e * 576 + ((v >> 1) * 8 + lane) * 2 + (v & 1)] += acc[l][I][Kc];
        } else {
          col[ln + 8 * I] += acc[l][I][Kc];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Element-matrix apply: Y(j,e) += sum_i E(i,j,e) X(i,e) (the reference applies the stored matrix
// transposed, src/mechanics_operator_ext.cpp:303-314).  lane = node; it owns columns
// j = lane_native + 8*comp.  LVEC fuses restriction, restriction^T and essential-dof masking
// (EANonlinearMechOperatorGradExt::TMult, :278-328).
// ------------------------------------------------------------------------------------------
template <int MODE, bool PRIV = false>
__global__ void __launch_bounds__(256) k_ea_mult(const double* __restrict__ ea, const double* __restrict__ x,
                                                 double* __restrict__ y, ElemIO io, long nelems,
                                                 double* __restrict__ dot_accum) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  const bool active = e < nelems;
  const int an = lex_to_native(lane);
  long nid = 0; unsigned msk = 0;
  double u0 = 0, u1 = 0, u2 = 0;
  if (active) {
    if (MODE == LVEC) {
      nid = io.e2n[e * 8 + an];
      msk = io.essmask ? io.essmask[nid] : 0u;
      u0 = (msk & 1) ? 0.0 : x[nid];
      u1 = (msk & 2) ? 0.0 : x[io.nnodes + nid];
      u2 = (msk & 4) ? 0.0 : x[2 * io.nnodes + nid];
    } else {
      nid = e * 24 + an;
      u0 = x[nid]; u1 = x[nid + 8]; u2 = x[nid + 16];
    }
  }
  // all 24 element dofs to every lane: xe[i] with i = a_native + 8*comp
  double xe[24];
  const int base = (threadIdx.x & 31) & ~7;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int src = base + lex_to_native(a);  // lane holding native node a (lex_to_native is an involution)
    xe[a] = __shfl_sync(kFull, u0, src);
    xe[a + 8] = __shfl_sync(kFull, u1, src);
    xe[a + 16] = __shfl_sync(kFull, u2, src);
  }
  double r[3] = {0.0, 0.0, 0.0};
  if (active) {
#pragma unroll
  for (int cmp = 0; cmp < 3; ++cmp) {
    // reference layout: this lane's column is 192 contiguous bytes; private layout: chunk-interleaved over the lanes
    const double2* col = reinterpret_cast<const double2*>(ea + e * 576 + (PRIV ? (long)(cmp * 12 * 8 + lane) * 2 : (long)(an + 8 * cmp) * 24));
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const double2 m = col[PRIV ? i * 8 : i];
      s0 += m.x * xe[2 * i];
      s1 += m.y * xe[2 * i + 1];
    }
    r[cmp] = s0 + s1;
  }
  if (MODE == LVEC) {
    if (!(msk & 1)) red_add_f64(&y[nid], r[0]);
    if (!(msk & 2)) red_add_f64(&y[io.nnodes + nid], r[1]);
    if (!(msk & 4)) red_add_f64(&y[2 * io.nnodes + nid], r[2]);
  } else {
    y[nid] += r[0]; y[nid + 8] += r[1]; y[nid + 16] += r[2];
  }
  }
  if (dot_accum) {
    double v = u0 * r[0] + u1 * r[1] + u2 * r[2];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
    if ((threadIdx.x & 31) == 0 && v != 0.0) red_add_f64(dot_accum, v);
  }
}

// ------------------------------------------------------------------------------------------
// Element-matrix apply of the L-vector EA path, pipelined like K2: persistent CTAs, every warp owns a STAGES-deep ring
// of 4-element sub-tiles (4 x 4608 B of the context-private, lane-interleaved element matrices: ONE bulk TMA copy per
// stage since consecutive elements are contiguous), completion on the warp's own mbarriers, refill right after the
// last shared-memory read.  In shared memory lane k's i-th 16-byte chunk sits at ((c*12 + i)*8 + k)*16, so the 8 lanes
// of an element read 128 contiguous bytes per LDS.128: conflict-free.
// ------------------------------------------------------------------------------------------
constexpr int kEaElemBytes = 576 * 8;
constexpr int kEaStageBytes = 4 * kEaElemBytes;  // 18432

template <int NW, int STAGES, bool ESS>
__global__ void __launch_bounds__(NW * 32) k_ea_mult_p(const double* __restrict__ ea, const double* __restrict__ x,
                                                       double* __restrict__ y, ElemIO io, long nelems,
                                                       double* __restrict__ dot_accum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int w = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), l32 = threadIdx.x & 31;
  const int lane = l32 & 7, el = l32 >> 3;
  unsigned char* ring = smem_raw + (size_t)w * STAGES * kEaStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NW * STAGES * kEaStageBytes) + w * STAGES;
  const long nwt = (nelems + 3) >> 2;
  const long stride = (long)gridDim.x * NW;
  const long wt0 = (long)blockIdx.x * NW + w;
  const uint64_t pol = l2_policy_evict_first();
  if (l32 == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  auto issue = [&](long wt, int s) {
    if (l32 == 0) {
      const long e0 = wt << 2;
      const uint32_t bytes = (uint32_t)min(4L, nelems - e0) * kEaElemBytes;
      mbar_arrive_expect_tx(&full[s], bytes);
      bulk_g2s_hint(ring + s * kEaStageBytes, ea + e0 * 576, bytes, &full[s], pol);
    }
  };
  {
    long t = wt0;
    for (int s = 0; s < STAGES; ++s, t += stride)
      if (t < nwt) issue(t, s);
  }
  auto load_nid = [&](long wt) -> int {
    const long e = (wt << 2) + el;
    if (wt >= nwt || e >= nelems) return -1;
    return io.e2n[e * 8 + lex_to_native(lane)];
  };
  auto load_x = [&](int nid, unsigned& msk, double& x0, double& x1, double& x2) {
    msk = 0; x0 = x1 = x2 = 0.0;
    if (nid < 0) return;
    if (ESS) msk = io.essmask[nid];
    x0 = x[nid]; x1 = x[io.nnodes + nid]; x2 = x[2 * io.nnodes + nid];
  };
  int nid_c = load_nid(wt0), nid_n = load_nid(wt0 + stride);
  unsigned msk_c; double xc0, xc1, xc2;
  load_x(nid_c, msk_c, xc0, xc1, xc2);
  const int base = l32 & ~7;
  int s = 0;
  uint32_t phase = 0;
  double xdoty = 0.0;
  for (long wt = wt0; wt < nwt; wt += stride) {
    const int nid_n2 = load_nid(wt + 2 * stride);
    unsigned msk_n; double xn0, xn1, xn2;
    load_x(nid_n, msk_n, xn0, xn1, xn2);
    const double u0 = (msk_c & 1) ? 0.0 : xc0, u1 = (msk_c & 2) ? 0.0 : xc1, u2 = (msk_c & 4) ? 0.0 : xc2;
    // all 24 element dofs to every lane: xe[a_native + 8*comp]
    double xe[24];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int src = base + lex_to_native(a);  // lane holding native node a (lex_to_native is an involution)
      xe[a] = __shfl_sync(kFull, u0, src);
      xe[a + 8] = __shfl_sync(kFull, u1, src);
      xe[a + 16] = __shfl_sync(kFull, u2, src);
    }
    mbar_wait(&full[s], phase);
    const bool active = nid_c >= 0;
    double r[3] = {0.0, 0.0, 0.0};
    if (active) {
      const double2* blk = reinterpret_cast<const double2*>(ring + s * kEaStageBytes + el * kEaElemBytes) + lane;
#pragma unroll
      for (int cmp = 0; cmp < 3; ++cmp) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          const double2 m = blk[(cmp * 12 + i) * 8];
          s0 += m.x * xe[2 * i];
          s1 += m.y * xe[2 * i + 1];
        }
        r[cmp] = s0 + s1;
      }
    }
    __syncwarp();
    {
      const long tnext = wt + (long)STAGES * stride;
      if (tnext < nwt) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(tnext, s);
      }
    }
    if (active) {
      xdoty += u0 * r[0] + u1 * r[1] + u2 * r[2];
      if (!(msk_c & 1)) red_add_f64(&y[nid_c], r[0]);
      if (!(msk_c & 2)) red_add_f64(&y[io.nnodes + nid_c], r[1]);
      if (!(msk_c & 4)) red_add_f64(&y[2 * io.nnodes + nid_c], r[2]);
    }
    nid_c = nid_n; nid_n = nid_n2;
    msk_c = msk_n; xc0 = xn0; xc1 = xn1; xc2 = xn2;
    if (++s == STAGES) { s = 0; phase ^= 1; }
  }
  if (dot_accum) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) xdoty += __shfl_xor_sync(kFull, xdoty, m);
    if (l32 == 0) red_add_f64(dot_accum, xdoty);
  }
}

// EA AssembleDiagonal (src/mechanics_operator_ext.cpp:228-265): diag of the element matrices.
template <int MODE, bool PRIV = false>
__global__ void __launch_bounds__(256) k_ea_diag(const double* __restrict__ ea, double* __restrict__ diag, ElemIO io,
                                                 long nelems) {
  const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 7;
  const long e = gt >> 3;
  if (e >= nelems) return;
  const int an = lex_to_native(lane);
#pragma unroll
  for (int cmp = 0; cmp < 3; ++cmp) {
    const int j = an + 8 * cmp;
    const int pv = cmp * 24 + j;  // private layout: entry (row j) of this lane's column j
    const double v = PRIV ? ea[e * 576 + ((pv >> 1) * 8 + lane) * 2 + (pv & 1)] : ea[e * 576 + (long)j * 24 + j];
    if (MODE == LVEC) red_add_f64(&diag[cmp * io.nnodes + io.e2n[e * 8 + an]], v);
    else diag[e * 24 + j] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Volume-weighted sums of a quadrature function: out[c] = sum w f_c (c < vdim), out[vdim] = sum w,
// w = detJ W.  One pass over the data (the reference makes `vdim` passes,
// src/mechanics_kernels.hpp:51-133).  out must be zeroed; block partials are combined with
// one red.add per component per block.
// ------------------------------------------------------------------------------------------
template <int VDIM_MAX>
__global__ void __launch_bounds__(256) k_vol_sum(const double* __restrict__ qf, const double* __restrict__ jac, int vdim,
                                                 long npts, double* __restrict__ out, double* __restrict__ partial,
                                                 unsigned int* __restrict__ counter) {
  __shared__ double red[8][VDIM_MAX + 1];
  double acc[VDIM_MAX + 1];
#pragma unroll
  for (int c = 0; c <= VDIM_MAX; ++c) acc[c] = 0.0;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (long)gridDim.x * blockDim.x) {
    double J[9], adj[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = jac[p * 9 + i];
    const double w = adjugate(J, adj) * kWq;
    acc[VDIM_MAX] += w;
#pragma unroll
    for (int c = 0; c < VDIM_MAX; ++c)
      if (c < vdim) acc[c] += w * qf[p * vdim + c];
  }
#pragma unroll
  for (int c = 0; c <= VDIM_MAX; ++c) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc[c] += __shfl_xor_sync(kFull, acc[c], m);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int c = 0; c <= VDIM_MAX; ++c) red[warp][c] = acc[c];
  __syncthreads();
  // block partials, then the last block to finish sums them in block order: bitwise reproducible averages
  __shared__ bool last;
  if (threadIdx.x <= VDIM_MAX) {
    const int c = threadIdx.x;
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][c];
    partial[(long)blockIdx.x * (VDIM_MAX + 1) + c] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x <= VDIM_MAX) {
    const int c = threadIdx.x;
    double s = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partial[(long)b * (VDIM_MAX + 1) + c]);
    if (c < vdim) out[c] = s;
    else if (c == VDIM_MAX) out[vdim] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Owner-computes scatter (the deterministic operator option): y_L(node) = sum of the E-vector entries of the elements
// around the node, in ascending element order (n2e: 8 slots per node, element*8 + native local node, -1 padded) -- the
// role of ElementRestriction::MultTranspose (src/mechanics_operator_ext.cpp:149) without atomics, so the operator, the
// residual and the diagonal are bitwise reproducible run to run.  With x != nullptr the block partial sums of
// x_masked . y are written too and the last block adds them in block order into *dot_out (the CG denominator).
// ------------------------------------------------------------------------------------------
template <bool ESS>
__global__ void __launch_bounds__(256) k_evec_to_lvec(const double* __restrict__ yE, const int* __restrict__ n2e,
                                                      double* __restrict__ yL, const unsigned char* __restrict__ essmask,
                                                      long nnodes, int ess_value_is_one, const double* __restrict__ x,
                                                      double* __restrict__ partial, unsigned int* __restrict__ counter,
                                                      double* __restrict__ dot_out) {
  double dot = 0.0;
  for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < nnodes; n += (long)gridDim.x * blockDim.x) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int4 s0 = reinterpret_cast<const int4*>(n2e)[2 * n], s1 = reinterpret_cast<const int4*>(n2e)[2 * n + 1];
    const int ids[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (ids[k] >= 0) {
        const long o = (long)(ids[k] >> 3) * 24 + (ids[k] & 7);
        a0 += yE[o]; a1 += yE[o + 8]; a2 += yE[o + 16];
      }
    }
    const unsigned msk = ESS ? essmask[n] : 0u;
    const double fill = ess_value_is_one ? 1.0 : 0.0;
    if (x) {
      if (!(msk & 1)) dot += x[n] * a0;
      if (!(msk & 2)) dot += x[nnodes + n] * a1;
      if (!(msk & 4)) dot += x[2 * nnodes + n] * a2;
    }
    yL[n] = (msk & 1) ? fill : a0;
    yL[nnodes + n] = (msk & 2) ? fill : a1;
    yL[2 * nnodes + n] = (msk & 4) ? fill : a2;
  }
  if (!x) return;
  __shared__ double red[8];
  __shared__ bool last;
  for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(kFull, dot, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  double t = 0.0;
  for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(&partial[b]);
  *dot_out += t;
}

}  // namespace exab

// Shared device helpers for the exab200 kernels (sm_100a).
//
// Work decomposition used by every element kernel: 8 consecutive lanes own one hexahedron.
// Lane l = bx + 2*by + 4*bz is at the same time node (bx,by,bz) and quadrature point
// (bx,by,bz) of the element in LEXICOGRAPHIC order; the reference's NATIVE (MFEM hex vertex)
// node order differs only by the swaps 2<->3 and 6<->7 (kNat below).  Nodal <-> quadrature
// transforms of the trilinear basis are done as three 2-point butterflies over
// __shfl_xor(1|2|4): sum-factorisation in registers, no shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exab {

constexpr double kAlpha = 0.78867513459481288225;  // (1 + 1/sqrt(3))/2 : N_a(xi_q), a == q
constexpr double kBeta = 0.21132486540518711775;   // (1 - 1/sqrt(3))/2 : N_a(xi_q), a != q
constexpr double kWq = 0.125;                      // Gauss weight of the 2x2x2 rule on [0,1]^3
constexpr unsigned kFull = 0xffffffffu;

// lexicographic lane -> NATIVE hex vertex index (src/mechanics_integrators.cpp E-vector order)
__device__ __forceinline__ int lex_to_native(int l) { return l ^ ((l >> 1) & 1); }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }

// Nodal values (one per lane) -> reference-space gradient (d/dxi, d/deta, d/dzeta) at this
// lane's quadrature point.  6 double shuffles.
__device__ __forceinline__ void nodal_to_qp_grad(double u, int lane, double& gx, double& gy, double& gz) {
  const double sx = (lane & 1) ? 1.0 : -1.0, sy = (lane & 2) ? 1.0 : -1.0, sz = (lane & 4) ? 1.0 : -1.0;
  double o = shfl_xor_d(u, 1);
  const double I = kAlpha * u + kBeta * o;
  const double D = sx * (u - o);
  double oI = shfl_xor_d(I, 2), oD = shfl_xor_d(D, 2);
  const double II = kAlpha * I + kBeta * oI;
  const double DI = kAlpha * D + kBeta * oD;
  const double ID = sy * (I - oI);
  const double oII = shfl_xor_d(II, 4), oDI = shfl_xor_d(DI, 4), oID = shfl_xor_d(ID, 4);
  gx = kAlpha * DI + kBeta * oDI;
  gy = kAlpha * ID + kBeta * oID;
  gz = sz * (II - oII);
}

// Same transforms with the lane's three signs hoisted by the caller (persistent kernels: computed once per thread).
// A sign is kept as the bit to XOR into the high word (0 for +1, 0x80000000 for -1): the flip is one integer LOP3
// instead of a DMUL on the fp64 pipe, bit-identical to the multiplication by +-1.0 (27 per element and lane in K2).
struct LaneSigns { unsigned sx, sy, sz; };
__device__ __forceinline__ LaneSigns lane_signs(int lane) {
  return LaneSigns{(lane & 1) ? 0u : 0x80000000u, (lane & 2) ? 0u : 0x80000000u, (lane & 4) ? 0u : 0x80000000u};
}
__device__ __forceinline__ double flip(double v, unsigned m) {
  return __hiloint2double(__double2hiint(v) ^ (int)m, __double2loint(v));
}
__device__ __forceinline__ void nodal_to_qp_grad(double u, const LaneSigns& sg, double& gx, double& gy, double& gz) {
  double o = shfl_xor_d(u, 1);
  const double I = kAlpha * u + kBeta * o;
  const double D = flip(u - o, sg.sx);
  double oI = shfl_xor_d(I, 2), oD = shfl_xor_d(D, 2);
  const double II = kAlpha * I + kBeta * oI;
  const double DI = kAlpha * D + kBeta * oD;
  const double ID = flip(I - oI, sg.sy);
  const double oII = shfl_xor_d(II, 4), oDI = shfl_xor_d(DI, 4), oID = shfl_xor_d(ID, 4);
  gx = kAlpha * DI + kBeta * oDI;
  gy = kAlpha * ID + kBeta * oID;
  gz = flip(II - oII, sg.sz);
}
// Transpose with 6 double shuffles instead of 8: in the y and x stages the partner combines what this lane needs of
// it into ONE value before sending (the partner's sign is the opposite of this lane's, so it can apply it itself).
__device__ __forceinline__ double qp_grad_to_nodal(double tx, double ty, double tz, const LaneSigns& sg) {
  const double otx = shfl_xor_d(tx, 4), oty = shfl_xor_d(ty, 4), otz = shfl_xor_d(tz, 4);
  const double DI = kAlpha * tx + kBeta * otx;
  const double sID = flip(kAlpha * ty + kBeta * oty, sg.sy);  // sy * ID
  const double II = flip(tz + otz, sg.sz);
  // I = alpha II + sy ID + [beta II_p + sy ID_p], and sy ID_p = -(sy_p ID_p)
  const double rI = shfl_xor_d(kBeta * II - sID, 2), oDI = shfl_xor_d(DI, 2);
  const double I = kAlpha * II + sID + rI;
  const double sD = flip(kAlpha * DI + kBeta * oDI, sg.sx);  // sx * D
  const double r = shfl_xor_d(kBeta * I - sD, 1);
  return kAlpha * I + sD + r;
}

// Transpose of the above: per-quadrature-point (t_xi, t_eta, t_zeta) -> nodal value
//   Y(a) = sum_q [ G(a,0,q) t_xi(q) + G(a,1,q) t_eta(q) + G(a,2,q) t_zeta(q) ].  8 double shuffles.
__device__ __forceinline__ double qp_grad_to_nodal(double tx, double ty, double tz, int lane) {
  const double sx = (lane & 1) ? 1.0 : -1.0, sy = (lane & 2) ? 1.0 : -1.0, sz = (lane & 4) ? 1.0 : -1.0;
  const double otx = shfl_xor_d(tx, 4), oty = shfl_xor_d(ty, 4), otz = shfl_xor_d(tz, 4);
  const double DI = kAlpha * tx + kBeta * otx;
  const double ID = kAlpha * ty + kBeta * oty;
  const double II = sz * (tz + otz);
  const double oDI = shfl_xor_d(DI, 2), oID = shfl_xor_d(ID, 2), oII = shfl_xor_d(II, 2);
  const double I = kAlpha * II + kBeta * oII + sy * (ID + oID);
  const double D = kAlpha * DI + kBeta * oDI;
  const double oI = shfl_xor_d(I, 1), oD = shfl_xor_d(D, 1);
  return kAlpha * I + kBeta * oI + sx * (D + oD);
}

// Shape-function gradient G(a, s, q) of lexicographic node a at lexicographic point q.
__device__ __forceinline__ void shape_grad(int a, int q, double g[3]) {
  const double nx = ((a ^ q) & 1) ? kBeta : kAlpha, ny = ((a ^ q) & 2) ? kBeta : kAlpha,
               nz = ((a ^ q) & 4) ? kBeta : kAlpha;
  const double dx = (a & 1) ? 1.0 : -1.0, dy = (a & 2) ? 1.0 : -1.0, dz = (a & 4) ? 1.0 : -1.0;
  g[0] = dx * ny * nz;
  g[1] = nx * dy * nz;
  g[2] = nx * ny * dz;
}

// adj[3*r+c] = adj(J)(r,c), J given as J[3*s+i] = dx_i/dxi_s (reference layout); returns det J.
__device__ __forceinline__ double adjugate(const double* J, double* adj) {
  const double J11 = J[0], J21 = J[1], J31 = J[2];
  const double J12 = J[3], J22 = J[4], J32 = J[5];
  const double J13 = J[6], J23 = J[7], J33 = J[8];
  adj[0] = (J22 * J33) - (J23 * J32);
  adj[1] = (J32 * J13) - (J12 * J33);
  adj[2] = (J12 * J23) - (J22 * J13);
  adj[3] = (J31 * J23) - (J21 * J33);
  adj[4] = (J11 * J33) - (J13 * J31);
  adj[5] = (J21 * J13) - (J11 * J23);
  adj[6] = (J21 * J32) - (J31 * J22);
  adj[7] = (J31 * J12) - (J11 * J32);
  adj[8] = (J11 * J22) - (J12 * J21);
  return J11 * adj[0] + J21 * adj[1] + J31 * adj[2];
}

// ---- mbarrier / bulk-copy (TMA 1-D) wrappers --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk async copy, completion counted on an mbarrier (UBLKCP in SASS).
// bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Same copy with an L2 eviction-priority hint: the operand stream of the gradient apply is read exactly once per
// launch, so it is marked evict_first and leaves the L2 to the CG vectors (d, z, r) that the neighbouring vector
// kernels and the gather / red.add traffic of the apply itself reuse.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// 2-D tiled tensor copy global -> shared (TMA with a CUtensorMap descriptor, UTMALDG in SASS), completion counted
// on an mbarrier; c0 = innermost coordinate (elements), c1 = row.
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tensor_map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst_smem)),
      "l"(tensor_map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

}  // namespace exab

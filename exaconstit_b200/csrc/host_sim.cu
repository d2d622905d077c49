// Host side of the hot path: C++ classes that mirror the reference's plugin/operator interface
// (same names, argument meaning and error behaviour) and drive the exab200 C ABI.  MFEM is absent
// in this image, so the few MFEM facilities the path needs (device vectors, CG, the L<->E
// restriction, shared-dof sums) are provided here for the structured voxel mesh; with MFEM present
// the same classes sit behind mfem::Vector Read()/Write() pointers.
//
//   ExaModel / ExaCMechModel            src/mechanics_model.hpp:17-241, src/mechanics_ecmech.hpp:12-107
//   NonlinearMechOperator               src/mechanics_operator.hpp, src/mechanics_operator.cpp:288-483
//   gradient operator (PA/EA ext)       src/mechanics_operator_ext.cpp:95-174,228-328
//   MechOperatorJacobiSmoother          src/mechanics_operator_ext.cpp:11-55
//   CGSolver                            mfem::CGSolver as configured at src/system_driver.cpp:166-178
//   ExaNewtonSolver / ExaNewtonLSSolver src/mechanics_solver.cpp:39-143,155-280
//   SystemDriver                        src/system_driver.cpp:221-319,327-333,429-468
//
// Multi-GPU: one process per GPU, z-slab element partition; each rank keeps a local L-vector with
// both interface node planes.  The only data-path exchanges are (i) the sum of interface-plane
// partial results with the two z-neighbours after every operator/residual/diagonal action
// (ncclSend/ncclRecv pair inside one group) and (ii) ncclAllReduce of the CG / Newton scalars.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/exab200.h"
#include "../../include/exahost.h"
#include "exab200_p2p.cuh"

namespace exahost {

static thread_local std::string g_err;
struct Abort { std::string msg; };  // the reference calls MFEM_ABORT; we unwind to the C boundary
#define HCK(call)                                                                           \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) throw Abort{std::string(#call) + ": " + cudaGetErrorString(e_)}; \
  } while (0)
#define XCK(call)                                                    \
  do {                                                               \
    if ((call) != 0) throw Abort{std::string(exab200_last_error())}; \
  } while (0)

// ---------------------------------------------------------------- NCCL (resolved lazily) ----
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define SYM(f, n) f = reinterpret_cast<decltype(f)>(dlsym(h, n))
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return GetUniqueId && CommInitRank && AllReduce && Send && Recv && GroupStart && GroupEnd;
  }
};
static NcclApi g_nccl;
#define NCK(call)                                                                                   \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess) throw Abort{std::string(#call) + ": " + g_nccl.GetErrorString(r_)};      \
  } while (0)

// ---------------------------------------------------------------- device vector -------------
struct Vector {
  double* d = nullptr;
  long n = 0;
  bool own = true;
  Vector() = default;
  explicit Vector(long n_) { SetSize(n_); }
  Vector(const Vector&) = delete;
  Vector& operator=(const Vector&) = delete;
  ~Vector() { if (own) cudaFree(d); }
  // non-owning view of part of another allocation
  void MakeRef(double* p, long n_) {
    if (own) cudaFree(d);
    own = false; d = p; n = n_;
  }
  void SetSize(long n_) {
    if (n_ == n) return;
    if (own) cudaFree(d);
    own = true;
    d = nullptr;
    n = n_;
    if (n > 0) HCK(cudaMalloc(&d, sizeof(double) * n));
  }
  long Size() const { return n; }
  const double* Read() const { return d; }
  double* Write() { return d; }
  double* ReadWrite() { return d; }
};

// ---------------------------------------------------------------- vector kernels ------------
constexpr int kRedBlocks = 592;  // 4 x 148 SMs
__global__ void __launch_bounds__(256) k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, long nn,
                                                     long n_owned, double* __restrict__ partial) {
  // dot over the owned nodes of a byNODES L-vector: comps c in 0..2, nodes [0, n_owned)
  double s = 0.0;
  for (int c = 0; c < 3; ++c) {
    const long off = c * nn;
    for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < n_owned; n += (long)gridDim.x * blockDim.x)
      s += a[off + n] * b[off + n];
  }
  __shared__ double red[8];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_dot_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  __shared__ double red[8];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    *out = t;
  }
}
// x += alpha d ; r -= alpha z     (CG update, fused)
__global__ void k_cg_update(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d,
                            const double* __restrict__ z, double alpha, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { x[i] += alpha * d[i]; r[i] -= alpha * z[i]; }
}
// d = z + beta d
__global__ void k_xpby(double* __restrict__ d, const double* __restrict__ z, double beta, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = z[i] + beta * d[i];
}
// CG step 1 (device-resident scalars): alpha = nom/den; x += alpha d; r -= alpha z;
// partial[blk] = sum over uniquely-owned dofs of r . (M r), M = diag(dinv) or identity
__global__ void __launch_bounds__(256) k_cg_step1(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d,
                                                  const double* __restrict__ z, const double* __restrict__ dinv,
                                                  const double* __restrict__ nom, const double* __restrict__ den, long nn,
                                                  long n_owned, double* __restrict__ partial, double* __restrict__ den_next) {
  const double alpha = *nom / *den;
  if (blockIdx.x == 0 && threadIdx.x == 0) *den_next = 0.0;  // accumulator of the next fused d^T A d
  double s = 0.0;
  // component by component: no 64-bit division in the streaming loop
  for (int c = 0; c < 3; ++c) {
    const long off = c * nn;
    for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (long)gridDim.x * blockDim.x) {
      const long i = off + n;
      x[i] += alpha * d[i];
      const double rn = r[i] - alpha * z[i];
      r[i] = rn;
      if (n < n_owned) s += rn * (dinv ? dinv[i] * rn : rn);
    }
  }
  __shared__ double red[8];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
// CG step 2: beta = betanom/nom; d = M r + beta d; y = 0 (output of the next operator apply)
__global__ void __launch_bounds__(256) k_cg_step2(double* __restrict__ d, const double* __restrict__ r,
                                                  const double* __restrict__ dinv, double* __restrict__ y,
                                                  const double* __restrict__ betanom, const double* __restrict__ nom, long n) {
  const double beta = *betanom / *nom;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double ri = r[i];
    d[i] = (dinv ? dinv[i] * ri : ri) + beta * d[i];
    y[i] = 0.0;
  }
}

// y = a x + b y
__global__ void k_axpby(double* __restrict__ y, const double* __restrict__ x, double a, double b, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i] + b * y[i];
}
// y = dinv .* r
__global__ void k_jacobi(double* __restrict__ y, const double* __restrict__ dinv, const double* __restrict__ r, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = dinv[i] * r[i];
}
__global__ void k_jacobi_setup(double* __restrict__ dinv, const double* __restrict__ diag, const unsigned char* __restrict__ ess,
                               long nn, double damping) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * nn) return;
  const long c = i / nn, n = i - c * nn;
  dinv[i] = ((ess[n] >> c) & 1) ? damping : damping / diag[i];
}
// v[ess] = val[ess]    (UpdateVelocity);  with keep_other=false also zero the non-essential dofs
__global__ void k_set_ess(double* __restrict__ v, const double* __restrict__ val, const unsigned char* __restrict__ ess,
                          long nn, int mode /*0: v[ess]=val; 1: v = ess ? val - v : 0; 2: v[ess] = 0*/,
                          const double* __restrict__ other) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * nn) return;
  const long c = i / nn, n = i - c * nn;
  const bool e = (ess[n] >> c) & 1;
  if (mode == 0) { if (e) v[i] = val[i]; }
  else if (mode == 1) v[i] = e ? (val[i] - other[i]) : 0.0;
  else if (mode == 2) { if (e) v[i] = 0.0; }
  else if (mode == 3) { if (e) v[i] = 1.0; }
}
// interface-plane add: dst[c*nn + off + i] += src[c*plane + i]
__global__ void k_plane_add(double* __restrict__ v, const double* __restrict__ buf, long nn, long off, long plane) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * plane) return;
  const long c = i / plane, n = i - c * plane;
  v[c * nn + off + n] += buf[i];
}
__global__ void k_plane_pack(const double* __restrict__ v, double* __restrict__ buf, long nn, long off, long plane) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * plane) return;
  const long c = i / plane, n = i - c * plane;
  buf[i] = v[c * nn + off + n];
}
// component-wise minimum of a byNODES L-vector: partial[c*gridDim.x + blk], then out[c]
__global__ void __launch_bounds__(256) k_min3_partial(const double* __restrict__ x, long nn, double* __restrict__ partial) {
  __shared__ double red[8];
  for (int c = 0; c < 3; ++c) {
    double m = 1.0e300;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long)gridDim.x * blockDim.x) m = fmin(m, x[c * nn + i]);
    for (int k = 16; k > 0; k >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = red[0];
      for (int w = 1; w < 8; ++w) t = fmin(t, red[w]);
      partial[c * gridDim.x + blockIdx.x] = t;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_min3_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  __shared__ double red[8];
  for (int c = 0; c < 3; ++c) {
    double m = 1.0e300;
    for (int i = threadIdx.x; i < nb; i += 256) m = fmin(m, partial[c * nb + i]);
    for (int k = 16; k > 0; k >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, k));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = red[0];
      for (int w = 1; w < 8; ++w) t = fmin(t, red[w]);
      out[c] = t;
    }
    __syncthreads();
  }
}
// velocity-gradient BC (src/system_driver.cpp:404-426): v_d = sum_j L(d,j) (x_j - origin_j) on the marked dofs
struct VGrad { double L[9]; };
__global__ void k_vgrad_vel(double* __restrict__ v, const double* __restrict__ x, const unsigned char* __restrict__ vgmask,
                            const double* __restrict__ origin, VGrad g, long nn) {
  const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nn) return;
  const unsigned m = vgmask[n];
  if (!m) return;
  const double r0 = x[n] - origin[0], r1 = x[nn + n] - origin[1], r2 = x[2 * nn + n] - origin[2];
  for (int d = 0; d < 3; ++d)
    if ((m >> d) & 1) {
      double s = 0.0;
      s += g.L[3 * d + 0] * r0;
      s += g.L[3 * d + 1] * r1;
      s += g.L[3 * d + 2] * r2;
      v[d * nn + n] = s;
    }
}
static long g_host_launches = 0;
static inline long& g_host_launches_ref() { return g_host_launches; }
static inline unsigned nb(long n) { ++g_host_launches; return (unsigned)((n + 255) / 256); }

// ---------------------------------------------------------------- per-kernel timing ---------
// CUDA-event pairs recorded on the launching stream around one kernel family; elapsed times are
// harvested after the step (no host sync inside the hot loop).
struct KernelTimer {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  double total_ms = 0.0;
  long count = 0;
  bool enabled = false;
  ~KernelTimer() { for (auto e : ev) cudaEventDestroy(e); }
  void Begin(cudaStream_t s) {
    if (!enabled || used + 2 > 8192) return;
    if (used + 2 > ev.size()) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev.push_back(a); ev.push_back(b); }
    cudaEventRecord(ev[used], s);
  }
  void End(cudaStream_t s) {
    if (!enabled || used + 2 > 8192 || used + 2 > ev.size()) return;
    cudaEventRecord(ev[used + 1], s);
    used += 2;
  }
  void Harvest() {  // stream must be synchronised
    for (size_t i = 0; i + 1 < used; i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) { total_ms += ms; ++count; }
    }
    used = 0;
  }
};

// ---------------------------------------------------------------- NVLink peer-memory collectives
// protocol and device helpers: exab200_p2p.cuh
using exab_p2p::PeerTable;
using exab_p2p::warp_allreduce_p2p;
using exab_p2p::ld_cg;
using exab_p2p::kMbScal;
using exab_p2p::kMbHalo;

// The last block (when ar_val != nullptr) is not part of the plane exchange: it all-reduces one scalar
// (the CG denominator accumulated by the operator kernel) in the same launch.
__global__ void __launch_bounds__(256) k_halo_p2p(double* __restrict__ v, double* mb, double* lo, double* hi, long nn,
                                                  long plane, unsigned long long seq, double* ar_val, PeerTable peers,
                                                  int rank, int nranks, unsigned long long seq_scal) {
  const unsigned nblk = ar_val ? gridDim.x - 1 : gridDim.x;
  if (ar_val && blockIdx.x == nblk) {
    if (threadIdx.x < 32) warp_allreduce_p2p(ar_val, peers, rank, nranks, 1, seq_scal, threadIdx.x);
    return;
  }
  __shared__ bool s_ok;
  exab_p2p::halo_exchange_blocks(v, mb, lo, hi, nn, plane, seq, peers.spin_limit, blockIdx.x, nblk, &s_ok);
}

// in-place sum of val[0..n) over all ranks, n <= 8; one warp
__global__ void k_allreduce_p2p(double* __restrict__ val, PeerTable peers, int rank, int nranks, int n,
                                unsigned long long seq) {
  warp_allreduce_p2p(val, peers, rank, nranks, n, seq, threadIdx.x);
}
// block partial sums -> one scalar -> all-reduce, one launch
__global__ void __launch_bounds__(256) k_reduce_allreduce_p2p(const double* __restrict__ partial, int nb, double* __restrict__ out,
                                                              PeerTable peers, int rank, int nranks, unsigned long long seq) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  __shared__ double red[8];
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      *out = t;
    }
    __syncwarp();
    warp_allreduce_p2p(out, peers, rank, nranks, 1, seq, threadIdx.x);
  }
}

// CG step 1 with the reduction folded in: the last block to finish sums the block partials in a fixed order
// (bitwise reproducible), all-reduces the scalar over the ranks through the peer mailboxes when nranks > 1 and
// publishes betanom -- one launch instead of two per CG iteration.
__global__ void __launch_bounds__(256) k_cg_step1_fused(double* __restrict__ x, double* __restrict__ r,
                                                        const double* __restrict__ d, const double* __restrict__ z,
                                                        const double* __restrict__ dinv, const double* __restrict__ nom,
                                                        const double* __restrict__ den, long nn, long n_owned,
                                                        double* __restrict__ partial, double* __restrict__ den_next,
                                                        unsigned int* __restrict__ counter, double* __restrict__ d_bet,
                                                        PeerTable peers, int rank, int nranks, unsigned long long seq) {
  const double alpha = *nom / *den;
  if (blockIdx.x == 0 && threadIdx.x == 0) *den_next = 0.0;  // accumulator of the next fused d^T A d
  double s = 0.0;
  // component by component: no 64-bit division in the streaming loop
  for (int c = 0; c < 3; ++c) {
    const long off = c * nn;
    for (long n = (long)blockIdx.x * blockDim.x + threadIdx.x; n < nn; n += (long)gridDim.x * blockDim.x) {
      const long i = off + n;
      x[i] += alpha * d[i];
      const double rn = r[i] - alpha * z[i];
      r[i] = rn;
      if (n < n_owned) s += rn * (dinv ? dinv[i] * rn : rn);
    }
  }
  __shared__ double red[8];
  __shared__ bool last;
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;  // wraps back to 0 for the next launch
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += ld_cg(&partial[i]);
  for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      double u = 0.0;
      for (int w = 0; w < 8; ++w) u += red[w];
      *d_bet = u;
    }
    __syncwarp();
    if (nranks > 1) warp_allreduce_p2p(d_bet, peers, rank, nranks, 1, seq, threadIdx.x);
  }
}

// The whole vector part of a CG iteration in ONE launch: step 1 (r -= alpha z, betanom = r.Mr with the
// block reduction, the fixed-order final sum and -- for nranks > 1 -- the peer-memory all-reduce done by the last
// block to arrive; z = 0 for the next apply), a grid-wide barrier on a generation flag, then x += alpha d and step 2
// (d = M r + beta d) on the same slices while r is still in L2: nine vector passes instead of ten.  Replaces k_cg_step1_fused + k_cg_step2 (one launch and one full re-read of r and d
// less per iteration).  All blocks must be co-resident (the grid is sized from the occupancy query by the caller).
__global__ void __launch_bounds__(256, 4) k_cg_fused(double* __restrict__ x, double* __restrict__ r, double* __restrict__ d,
                                                  double* __restrict__ z, const double* __restrict__ dinv,
                                                  const double* __restrict__ nom, const double* __restrict__ den, long nn,
                                                  long n_owned, double* __restrict__ partial, double* __restrict__ den_next,
                                                  unsigned int* __restrict__ counter, double* __restrict__ d_bet,
                                                  PeerTable peers, int rank, int nranks, unsigned long long seq,
                                                  unsigned long long* __restrict__ gen_flag, unsigned long long gen,
                                                  volatile double* __restrict__ host_bet /* mapped pinned: {betanom, gen} */) {
  const double nom_v = *nom;
  const double alpha = nom_v / *den;
  if (blockIdx.x == 0 && threadIdx.x == 0) *den_next = 0.0;  // accumulator of the next fused d^T A d
  double s = 0.0;
  // four independent elements per thread and pass: with one, the loads in flight (151 k threads x 2 x 8 B) cover only
  // half of what HBM needs (ncu: 50 % DRAM throughput, long_scoreboard 69 per issue)
  const long first = (long)blockIdx.x * blockDim.x + threadIdx.x, step = (long)gridDim.x * blockDim.x;
  for (int c = 0; c < 3; ++c) {
    const long off = c * nn;
    for (long n = first; n < nn; n += 4 * step) {
      double rv[4], zv[4], mv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long m = n + u * step;
        if (m < nn) { rv[u] = r[off + m]; zv[u] = z[off + m]; mv[u] = dinv ? dinv[off + m] : 1.0; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long m = n + u * step;
        if (m < nn) {
          const double rn = rv[u] - alpha * zv[u];
          r[off + m] = rn;
          z[off + m] = 0.0;  // output of the next operator apply
          if (m < n_owned) s += rn * (dinv ? mv[u] * rn : rn);
        }
      }
    }
  }
  __shared__ double red[8];
  __shared__ bool last;
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
    __threadfence();
    last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;  // wraps back to 0 for the next launch
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += ld_cg(&partial[i]);
    for (int m = 16; m > 0; m >>= 1) t += __shfl_xor_sync(0xffffffffu, t, m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x < 32) {
      if (threadIdx.x == 0) {
        double u = 0.0;
        for (int w = 0; w < 8; ++w) u += red[w];
        *d_bet = u;
      }
      __syncwarp();
      if (nranks > 1) warp_allreduce_p2p(d_bet, peers, rank, nranks, 1, seq, threadIdx.x);
      __syncwarp();
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(gen_flag), "l"(gen) : "memory");
        // the host's stopping test reads betanom straight from mapped pinned memory (no copy-engine op and no event
        // between this kernel and the operator apply that follows it in the stream)
        host_bet[0] = *d_bet;
        __threadfence_system();
        host_bet[1] = __longlong_as_double((long long)gen);
      }
    }
  }
  // grid barrier: everyone waits for this launch's generation
  if (threadIdx.x == 0) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(gen_flag) : "memory");
      if (v < gen) __nanosleep(40);
    } while (v < gen);
  }
  __syncthreads();
  const double beta = ld_cg(d_bet) / nom_v;
  for (int c = 0; c < 3; ++c) {
    const long off = c * nn;
    for (long n = first; n < nn; n += 4 * step) {
      double rv[4], dv[4], xv[4], mv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long m = n + u * step;
        if (m < nn) { rv[u] = r[off + m]; dv[u] = d[off + m]; xv[u] = x[off + m]; mv[u] = dinv ? dinv[off + m] : 1.0; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long m = n + u * step;
        if (m < nn) {
          x[off + m] = xv[u] + alpha * dv[u];   // step 1's solution update, done here where d is read anyway
          d[off + m] = (dinv ? mv[u] * rv[u] : rv[u]) + beta * dv[u];
        }
      }
    }
  }
}

// ---------------------------------------------------------------- communicator --------------
class SlabComm {
 public:
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  long nn = 0, plane = 0, n_owned = 0;
  Vector send_lo, send_hi, recv_lo, recv_hi, partial, scal, counter;
  double* h_scal = nullptr;  // pinned
  long n_allreduce = 0, n_halo = 0;
  // NVLink peer-memory path
  Vector mailbox;
  PeerTable peers{};
  bool use_p2p = false;
  unsigned long long seq_halo = 0, seq_scal = 0;
  std::vector<void*> opened;

  void Init(int rank_, int nranks_, const void* nccl_id, cudaStream_t s, long nn_, long plane_) {
    if (nranks_ < 1 || nranks_ > 8 || rank_ < 0 || rank_ >= nranks_)
      throw Abort{"SlabComm: 1 <= nranks <= 8 (one box of NVLink peers: PeerTable and the mailbox layout hold 8 ranks) and 0 <= rank < nranks"};
    rank = rank_; nranks = nranks_; stream = s; nn = nn_; plane = plane_;
    peers.spin_limit = 20000000000LL;  // ~10 s of clock64 ticks; EXAHOST_SPIN_LIMIT_S overrides
    if (const char* e = std::getenv("EXAHOST_SPIN_LIMIT_S")) {
      const long long ticks = (long long)(std::atof(e) * 2.0e9);
      if (ticks > 0) peers.spin_limit = ticks;
    }
    n_owned = (rank == nranks - 1) ? nn : nn - plane;  // = exahost_slab_layout out[4] (checked by exahost_create)
    partial.SetSize(kRedBlocks);
    scal.SetSize(8);
    counter.SetSize(2);
    HCK(cudaMemset(counter.d, 0, 2 * sizeof(double)));
    HCK(cudaMallocHost(&h_scal, 64 * sizeof(double)));
    if (nranks > 1) {
      if (!g_nccl.load()) throw Abort{"NCCL library could not be loaded"};
      ncclUniqueId id;
      std::memcpy(&id, nccl_id, sizeof(id));
      NCK(g_nccl.CommInitRank(&comm, nranks, id, rank));
      send_lo.SetSize(3 * plane); send_hi.SetSize(3 * plane); recv_lo.SetSize(3 * plane); recv_hi.SetSize(3 * plane);
      mailbox.SetSize(kMbHalo + 12 * plane);
      HCK(cudaMemset(mailbox.d, 0, sizeof(double) * mailbox.n));
    }
  }
  void GetHandle(void* out64) {
    cudaIpcMemHandle_t h;
    HCK(cudaIpcGetMemHandle(&h, mailbox.d));
    static_assert(sizeof(h) == 64, "ipc handle size");
    std::memcpy(out64, &h, 64);
  }
  void SetPeers(const void* handles) {
    if (!handles) {  // back to the NCCL exchanges (a rank could not map a peer's mailbox)
      use_p2p = false;
      halo_fused = -1;
      return;
    }
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) { peers.p[r] = mailbox.d; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, static_cast<const char*>(handles) + 64 * r, 64);
      void* ptr = nullptr;
      HCK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      opened.push_back(ptr);
      peers.p[r] = static_cast<double*>(ptr);
    }
    use_p2p = true;
  }
  bool PeerError() {
    if (!use_p2p) return false;
    unsigned long long e = 0;
    cudaMemcpy(&e, reinterpret_cast<unsigned long long*>(mailbox.d) + 17, 8, cudaMemcpyDeviceToHost);
    return e != 0;
  }
  ~SlabComm() {
    for (void* p : opened) cudaIpcCloseMemHandle(p);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    if (h_scal) cudaFreeHost(h_scal);
    if (h_poll) cudaFreeHost(const_cast<double*>(h_poll));
  }
  // operator apply with the exchange folded in (exab200_grad_mult_halo): decided once per context
  int halo_fused = -1;
  bool HaloFusedOk(exab200_ctx* ctx, long layer_elems) {
    if (halo_fused < 0) {
      halo_fused = 0;
      if (use_p2p && nranks > 1 && !std::getenv("EXAHOST_NO_HALO_FUSION")) {
        exab200_halo h = MakeHaloDescriptor(layer_elems);
        halo_fused = exab200_grad_mult_halo_supported(ctx, &h);
      }
    }
    return halo_fused == 1;
  }
  exab200_halo MakeHaloDescriptor(long layer_elems) const {
    exab200_halo h;
    std::memset(&h, 0, sizeof(h));
    h.mailbox = mailbox.d;
    h.lo = rank > 0 ? peers.p[rank - 1] : nullptr;
    h.hi = rank < nranks - 1 ? peers.p[rank + 1] : nullptr;
    for (int r = 0; r < nranks; ++r) h.peers[r] = peers.p[r];
    h.spin_limit = peers.spin_limit;
    h.rank = rank; h.nranks = nranks;
    h.plane = plane; h.layer_elems = layer_elems;
    h.seq_halo = seq_halo; h.seq_scal = seq_scal;
    return h;
  }
  exab200_halo NextHaloDescriptor(long layer_elems) {
    ++seq_halo; ++seq_scal; ++n_halo; ++n_allreduce;
    return MakeHaloDescriptor(layer_elems);
  }
  // Sum the partial results living on the interface planes with the z-neighbours (the role of
  // P->MultTranspose followed by P->Mult in the reference, src/mechanics_operator_ext.cpp:149,157).
  void HaloSum(double* v, double* ar_scalar = nullptr) {
    if (nranks == 1) return;
    const bool lo = rank > 0, hi = rank < nranks - 1;
    const long top = nn - plane;
    if (use_p2p) {
      ++seq_halo;
      unsigned blocks = (unsigned)std::min<long>((3 * plane + 255) / 256, 140);
      if (ar_scalar) { ++seq_scal; ++n_allreduce; }
      k_halo_p2p<<<blocks + (ar_scalar ? 1 : 0), 256, 0, stream>>>(v, mailbox.d, lo ? peers.p[rank - 1] : nullptr,
                                                                   hi ? peers.p[rank + 1] : nullptr, nn, plane, seq_halo,
                                                                   ar_scalar, peers, rank, nranks, seq_scal);
      ++g_host_launches;
      ++n_halo;
      return;
    }
    if (lo) k_plane_pack<<<nb(3 * plane), 256, 0, stream>>>(v, send_lo.d, nn, 0, plane);
    if (hi) k_plane_pack<<<nb(3 * plane), 256, 0, stream>>>(v, send_hi.d, nn, top, plane);
    NCK(g_nccl.GroupStart());
    if (lo) { NCK(g_nccl.Send(send_lo.d, 3 * plane, ncclDouble, rank - 1, comm, stream)); NCK(g_nccl.Recv(recv_lo.d, 3 * plane, ncclDouble, rank - 1, comm, stream)); }
    if (hi) { NCK(g_nccl.Send(send_hi.d, 3 * plane, ncclDouble, rank + 1, comm, stream)); NCK(g_nccl.Recv(recv_hi.d, 3 * plane, ncclDouble, rank + 1, comm, stream)); }
    NCK(g_nccl.GroupEnd());
    if (lo) k_plane_add<<<nb(3 * plane), 256, 0, stream>>>(v, recv_lo.d, nn, 0, plane);
    if (hi) k_plane_add<<<nb(3 * plane), 256, 0, stream>>>(v, recv_hi.d, nn, top, plane);
    ++n_halo;
    if (ar_scalar) AllReduceDevice(ar_scalar, 1);
  }
  // global dot product over uniquely-owned dofs; blocking (returns the value on the host)
  double Dot(const double* a, const double* b) {
    k_dot_partial<<<kRedBlocks, 256, 0, stream>>>(a, b, nn, n_owned, partial.d);
    k_dot_final<<<1, 256, 0, stream>>>(partial.d, kRedBlocks, scal.d);
    g_host_launches += 2;
    AllReduceDevice(scal.d, 1);
    HCK(cudaMemcpyAsync(h_scal, scal.d, sizeof(double), cudaMemcpyDeviceToHost, stream));
    HCK(cudaStreamSynchronize(stream));
    return h_scal[0];
  }
  // CG step 1 + reduction (+ all-reduce) in one launch; falls back to two launches on the NCCL-only path
  void CgStep1(double* x, double* r, const double* d, const double* z, const double* dinv, const double* d_nom,
               const double* d_den, double* d_den_next, double* d_bet) {
    ++g_host_launches;
    if (nranks > 1 && !use_p2p) {
      k_cg_step1<<<kRedBlocks, 256, 0, stream>>>(x, r, d, z, dinv, d_nom, d_den, nn, n_owned, partial.d, d_den_next);
      ReduceToDevice(d_bet);
      return;
    }
    if (nranks > 1) { ++seq_scal; ++n_allreduce; }
    k_cg_step1_fused<<<kRedBlocks, 256, 0, stream>>>(x, r, d, z, dinv, d_nom, d_den, nn, n_owned, partial.d, d_den_next,
                                                     reinterpret_cast<unsigned int*>(counter.d), d_bet, peers, rank,
                                                     nranks, seq_scal);
  }
  // whole vector part of a CG iteration in one launch (k_cg_fused); false if this configuration has to use the
  // separate kernels (NCCL-only exchanges, or a grid that could not be made co-resident)
  int fused_grid = -1;
  unsigned long long fused_gen = 0;
  volatile double* h_poll = nullptr;  // mapped pinned {betanom, generation}
  double* d_poll = nullptr;
  bool CgFused(double* x, double* r, double* d, double* z, const double* dinv, const double* d_nom, const double* d_den,
               double* d_den_next, double* d_bet) {
    if (nranks > 1 && !use_p2p) return false;
    if (fused_grid < 0) {
      int dev = 0, sms = 0, occ = 0;
      HCK(cudaGetDevice(&dev));
      HCK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      HCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cg_fused, 256, 0));
      fused_grid = std::min<long>(kRedBlocks, (long)sms * occ);
      if (std::getenv("EXAHOST_NO_CG_FUSION")) fused_grid = 0;
      if (fused_grid > 0) {
        void* hp = nullptr;
        HCK(cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
        std::memset(hp, 0, 64);
        h_poll = static_cast<volatile double*>(hp);
        HCK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_poll), hp, 0));
      }
    }
    if (fused_grid <= 0) return false;
    ++g_host_launches;
    if (nranks > 1) { ++seq_scal; ++n_allreduce; }
    ++fused_gen;
    k_cg_fused<<<fused_grid, 256, 0, stream>>>(x, r, d, z, dinv, d_nom, d_den, nn, n_owned, partial.d, d_den_next,
                                               reinterpret_cast<unsigned int*>(counter.d), d_bet, peers, rank, nranks, seq_scal,
                                               reinterpret_cast<unsigned long long*>(counter.d) + 1, fused_gen, d_poll);
    return true;
  }
  // betanom of the last CgFused launch; spins on the mapped word (checks the stream now and then so that a failed
  // launch cannot hang the host)
  double WaitBetanom() {
    unsigned long spins = 0;
    for (;;) {
      const double tag = h_poll[1];
      long long t;
      std::memcpy(&t, &tag, 8);
      if ((unsigned long long)t == fused_gen) return h_poll[0];
      if ((++spins & 0xfffff) == 0) {
        const cudaError_t q = cudaStreamQuery(stream);
        if (q != cudaErrorNotReady) {
          const double tag2 = h_poll[1];
          std::memcpy(&t, &tag2, 8);
          if ((unsigned long long)t == fused_gen) return h_poll[0];
          throw Abort{std::string("CG: the vector kernel did not deliver its result: ") + cudaGetErrorString(q)};
        }
      }
    }
  }
  // partial sums -> one device scalar (+ allreduce), no host involvement
  void ReduceToDevice(double* d_out) {
    ++g_host_launches_ref();
    if (nranks > 1 && use_p2p) {
      ++seq_scal; ++n_allreduce;
      k_reduce_allreduce_p2p<<<1, 256, 0, stream>>>(partial.d, kRedBlocks, d_out, peers, rank, nranks, seq_scal);
      return;
    }
    k_dot_final<<<1, 256, 0, stream>>>(partial.d, kRedBlocks, d_out);
    AllReduceDevice(d_out, 1);
  }
  // component-wise minimum over ranks (vgrad origin, src/system_driver.cpp:399); once per step: NCCL
  void AllReduceMinDevice(double* d_buf, int n) {
    if (nranks == 1) return;
    ++n_allreduce;
    NCK(g_nccl.AllReduce(d_buf, d_buf, n, ncclDouble, ncclMin, comm, stream));
  }
  void AllReduceDevice(double* d_buf, int n) {
    if (nranks == 1) return;
    ++n_allreduce;
    if (use_p2p && n <= 8) {
      ++seq_scal;
      k_allreduce_p2p<<<1, 32, 0, stream>>>(d_buf, peers, rank, nranks, n, seq_scal);
      ++g_host_launches;
      return;
    }
    NCK(g_nccl.AllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, comm, stream));
  }
  // sum-reduce a small device buffer in place and fetch it
  void AllReduceFetch(double* d_buf, int n, double* h_out) {
    if (nranks > 1) {
      // chunks of 8 go through the peer-memory kernel
      if (use_p2p) { for (int o = 0; o < n; o += 8) AllReduceDevice(d_buf + o, std::min(8, n - o)); }
      else { NCK(g_nccl.AllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, comm, stream)); ++n_allreduce; }
    }
    HCK(cudaMemcpyAsync(h_scal, d_buf, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    HCK(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; ++i) h_out[i] = h_scal[i];
  }
};

// ---------------------------------------------------------------- ExaModel ------------------
enum class Assembly { PA = EXAB200_PA, EA = EXAB200_EA };

// Plugin surface of src/mechanics_model.hpp:17-241 (quadrature data as raw device arrays).
class ExaModel {
 public:
  int numProps, numStateVars;
  double dt = 0.0, t = 0.0;
  Vector *stress0, *stress1, *matGrad, *matVars0, *matVars1;
  Assembly assembly;
  ExaModel(Vector* s0, Vector* s1, Vector* mg, Vector* v0, Vector* v1, int nProps, int nStateVars, Assembly a)
      : numProps(nProps), numStateVars(nStateVars), stress0(s0), stress1(s1), matGrad(mg), matVars0(v0), matVars1(v1), assembly(a) {}
  virtual ~ExaModel() {}
  // src/mechanics_model.hpp:109-111; `vel` is the velocity L-vector here (restriction fused in the kernel)
  virtual void ModelSetup(const int nqpts, const int nelems, const int space_dim, const int nnodes, const Vector& jacobian,
                          const Vector& loc_grad, const Vector& vel) = 0;
  virtual void UpdateModelVars() = 0;
  void SetModelDt(double dt_) { dt = dt_; }
  double GetModelDt() const { return dt; }
  void UpdateStress() { std::swap(stress0->d, stress1->d); }        // src/mechanics_model.cpp:435-438
  void UpdateStateVars() { std::swap(matVars0->d, matVars1->d); }   // src/mechanics_model.cpp:440-443
};

class ExaCMechModel : public ExaModel {
 public:
  exab200_ctx* ctx;
  cudaStream_t stream;
  long model_setups = 0;
  ExaCMechModel(exab200_ctx* c, cudaStream_t s, Vector* s0, Vector* s1, Vector* mg, Vector* v0, Vector* v1, int nProps,
                Assembly a)
      : ExaModel(s0, s1, mg, v0, v1, nProps, exab200_num_state_vars(c), a), ctx(c), stream(s) {}
  void ModelSetup(const int, const int, const int, const int, const Vector& jacobian, const Vector&, const Vector& vel) override {
    XCK(exab200_model_setup(ctx, dt, jacobian.Read(), vel.Read(), stress0->Read(), matVars0->Read(), stress1->Write(),
                            matVars1->Write(), matGrad->Write(), stream));
    ++model_setups;
  }
  void UpdateModelVars() override {}
};

// ---------------------------------------------------------------- operators -----------------
class Operator {
 public:
  virtual ~Operator() {}
  virtual void Mult(const Vector& x, Vector& y) const = 0;
};

class NonlinearMechOperator;

// The matrix-free gradient (PA or EA flavour, src/mechanics_operator_ext.cpp:63-328).
class GradientOperator : public Operator {
 public:
  const NonlinearMechOperator* op;
  bool local_action = false;
  explicit GradientOperator(const NonlinearMechOperator* o) : op(o) {}
  void Mult(const Vector& x, Vector& y) const override;       // TMult<false>
  void LocalMult(const Vector& x, Vector& y) const;           // TMult<true>
  // y += K x into a pre-zeroed y and *d_den += x^T K x (fused CG denominator); halo-summed
  void MultAccDot(const Vector& x, Vector& y, double* d_den) const;
  void AssembleDiagonal(Vector& diag) const;
};

class NonlinearMechOperator : public Operator {
 public:
  exab200_ctx* ctx;
  SlabComm* comm;
  cudaStream_t stream;
  ExaModel* model;
  long nelems, nnodes;
  Vector* x_beg;                 // beginning-of-step coordinates (L-vector)
  mutable Vector el_jac, diag;
  mutable GradientOperator jacobian;
  Vector ess_mask_dev;           // device copy of the per-node essential mask (bytes)
  mutable long grad_mults = 0, residuals = 0;
  long layer_elems = 0;          // elements per z-layer of the slab (exchange folded into the operator apply)
  bool need_diag = false;        // set when the smoother refreshes its inverse diagonal (true_jacobi)
  mutable KernelTimer tm_grad_mult, tm_model_setup;

  NonlinearMechOperator(exab200_ctx* c, SlabComm* cm, cudaStream_t s, ExaModel* m, long ne, long nn, Vector* xb)
      : ctx(c), comm(cm), stream(s), model(m), nelems(ne), nnodes(nn), x_beg(xb), el_jac(ne * 72), diag(3 * nn), jacobian(this) {}

  // src/mechanics_operator.cpp:279-285
  void UpdateEssTDofs(const unsigned char* h_mask) { XCK(exab200_set_essential_mask(ctx, h_mask)); }

  // src/mechanics_operator.cpp:311-348: end coordinates, Jacobians, material update
  template <bool upd_crds>
  void Setup(const Vector& k) const {
    XCK(exab200_setup_jacobians(ctx, x_beg->Read(), upd_crds ? k.Read() : nullptr, model->dt, el_jac.Write(), stream));
    tm_model_setup.Begin(stream);
    model->ModelSetup(8, (int)nelems, 3, 8, el_jac, el_jac /*shape gradients are analytic in-kernel*/, k);
    tm_model_setup.End(stream);
  }
  // src/mechanics_operator.cpp:288-308: y = H(k)
  void Mult(const Vector& k, Vector& y) const override {
    Setup<true>(k);
    XCK(exab200_residual(ctx, el_jac.Read(), model->stress1->Read(), y.Write(), stream));
    comm->HaloSum(y.Write());
    ++residuals;
  }
  // src/mechanics_operator.cpp:436-443
  Operator& GetGradient(const Vector&) const {
    XCK(exab200_grad_setup(ctx, model->dt, model->matGrad->Read(), el_jac.Read(), stream));
    // The reference assembles the diagonal here on every call (src/mechanics_operator.cpp:441) although its smoother
    // never picks it up; nothing reads `diag` unless the real Jacobi smoother is on.
    if (need_diag) jacobian.AssembleDiagonal(diag);
    return jacobian;
  }
  // src/mechanics_operator.cpp:446-483
  Operator& GetUpdateBCsAction(const Vector& k, const Vector& x, Vector& y) const {
    Setup<false>(k);
    Vector resid(y.Size());
    XCK(exab200_grad_setup(ctx, model->dt, model->matGrad->Read(), el_jac.Read(), stream));
    jacobian.LocalMult(x, y);
    XCK(exab200_residual(ctx, el_jac.Read(), model->stress1->Read(), resid.Write(), stream));
    comm->HaloSum(resid.Write());
    ++residuals;
    // y[ess] = 0; y += resid
    k_set_ess<<<nb(3 * nnodes), 256, 0, stream>>>(y.Write(), nullptr, ess_dev(), nnodes, 2, nullptr);
    k_axpby<<<nb(3 * nnodes), 256, 0, stream>>>(y.Write(), resid.Read(), 1.0, 1.0, 3 * nnodes);
    HCK(cudaStreamSynchronize(stream));  // resid goes out of scope
    return jacobian;
  }
  const unsigned char* ess_dev() const { return reinterpret_cast<const unsigned char*>(ess_mask_dev.d); }
  // src/mechanics_operator.hpp:92-96
  const unsigned char* GetEssTDofList() const { return ess_dev(); }   // per-node component mask instead of an index list
  ExaModel* GetModel() const { return model; }
  // src/mechanics_operator.cpp:393-427: F = I + grad_X u at the quadrature points of the REFERENCE mesh
  void CalculateDeformationGradient(const Vector& x_ref, const Vector& x_cur, Vector& jac_ref, Vector& def_grad) const {
    XCK(exab200_setup_jacobians(ctx, x_ref.Read(), nullptr, 0.0, jac_ref.Write(), stream));
    XCK(exab200_grad_calc(ctx, jac_ref.Read(), x_cur.Read(), def_grad.Write(), stream));
  }
};

void GradientOperator::Mult(const Vector& x, Vector& y) const {
  op->tm_grad_mult.Begin(op->stream);
  XCK(exab200_grad_mult(op->ctx, x.Read(), y.Write(), 0, op->stream));
  op->tm_grad_mult.End(op->stream);
  op->comm->HaloSum(y.Write());
  ++op->grad_mults;
}
void GradientOperator::MultAccDot(const Vector& x, Vector& y, double* d_den) const {
  SlabComm* cm = op->comm;
  if (cm->HaloFusedOk(op->ctx, op->layer_elems)) {
    // one kernel: operator apply + interface-plane exchange + all-reduce of the denominator (boundary layers first,
    // the exchange overlaps the interior)
    exab200_halo h = cm->NextHaloDescriptor(op->layer_elems);
    op->tm_grad_mult.Begin(op->stream);
    XCK(exab200_grad_mult_halo(op->ctx, x.Read(), y.Write(), EXAB200_NO_ZERO, d_den, &h, op->stream));
    op->tm_grad_mult.End(op->stream);
    ++op->grad_mults;
    return;
  }
  op->tm_grad_mult.Begin(op->stream);
  XCK(exab200_grad_mult_ex(op->ctx, x.Read(), y.Write(), EXAB200_NO_ZERO, d_den, op->stream));
  op->tm_grad_mult.End(op->stream);
  cm->HaloSum(y.Write(), d_den);  // interface-plane sum + all-reduce of the denominator in one launch
  ++op->grad_mults;
}
void GradientOperator::LocalMult(const Vector& x, Vector& y) const {
  XCK(exab200_grad_mult(op->ctx, x.Read(), y.Write(), 1, op->stream));
  op->comm->HaloSum(y.Write());
  ++op->grad_mults;
}
void GradientOperator::AssembleDiagonal(Vector& diag) const {
  XCK(exab200_grad_diag(op->ctx, diag.Write(), op->stream));
  // shared nodes: sum partial diagonals, then essential entries back to 1 (they were summed twice)
  if (op->comm->nranks > 1) {
    op->comm->HaloSum(diag.Write());
    k_set_ess<<<nb(3 * op->nnodes), 256, 0, op->stream>>>(diag.Write(), nullptr, op->ess_dev(), op->nnodes, 3, nullptr);
  }
}

// src/mechanics_operator_ext.cpp:11-55.  Reference quirk: Setup() runs only in
// the constructor with diag == 1, so dinv stays 1; `refresh` = true gives a real Jacobi preconditioner.
class MechOperatorJacobiSmoother {
 public:
  long N;
  Vector dinv;
  double damping;
  cudaStream_t stream;
  bool refresh;
  MechOperatorJacobiSmoother(long nn, cudaStream_t s, bool refresh_, double dmp = 1.0)
      : N(3 * nn), dinv(3 * nn), damping(dmp), stream(s), refresh(refresh_) {
    // dinv = damping / 1
    std::vector<double> ones(N, damping);
    HCK(cudaMemcpy(dinv.d, ones.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  }
  void Setup(const Vector& diag, const unsigned char* ess, long nn) {
    k_jacobi_setup<<<nb(N), 256, 0, stream>>>(dinv.Write(), diag.Read(), ess, nn, damping);
  }
  void Mult(const Vector& x, Vector& y) const { k_jacobi<<<nb(N), 256, 0, stream>>>(y.Write(), dinv.Read(), x.Read(), N); }
};

// mfem::CGSolver::Mult, iterative_mode = false, with the smoother as preconditioner.
class CGSolver {
 public:
  SlabComm* comm;
  cudaStream_t stream;
  double rel_tol = 1e-7, abs_tol = 1e-27;
  int max_iter = 1000;
  const Operator* oper = nullptr;
  const MechOperatorJacobiSmoother* prec = nullptr;
  mutable Vector r, d, z, scal, dz;  // d and z are the two halves of dz (one L2 persistence window covers both)
  mutable int final_iter = 0, converged = 0;
  mutable long total_iters = 0;
  double* h_bet = nullptr;  // pinned ring for the stopping test
  cudaEvent_t ev[8];
  CGSolver(SlabComm* c, cudaStream_t s, long n) : comm(c), stream(s), r(n), scal(8), dz(2 * n + 32) {
    const long n_al = (n + 15) & ~15L;  // keep z 128-byte aligned
    d.MakeRef(dz.d, n);
    z.MakeRef(dz.d + n_al, n);
    HCK(cudaMallocHost(&h_bet, 8 * sizeof(double)));
    for (int i = 0; i < 8; ++i) HCK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  ~CGSolver() {
    if (h_bet) cudaFreeHost(h_bet);
    for (int i = 0; i < 8; ++i) cudaEventDestroy(ev[i]);
  }
  void SetOperator(const Operator& op) { oper = &op; }
  // Device-resident formulation of mfem::CGSolver::Mult (iterative_mode = false): the scalars nom, den,
  // betanom live on the device; alpha/beta are formed inside the vector kernels; the denominator d^T A d is
  // accumulated by the operator kernel itself.  The only host wait per iteration is on betanom (for the
  // reference's stopping test), and it is placed AFTER the next direction update + operator apply have been
  // enqueued -- those touch neither x nor r, so a positive test simply discards them -- which hides the
  // host round trip behind the operator kernel.  Iteration-for-iteration identical to the reference loop.
  void Mult(const Vector& b, Vector& x) const {
    const long n = b.Size();
    const long nn = n / 3;
    const GradientOperator* A = static_cast<const GradientOperator*>(oper);
    const double* dinv = prec->refresh ? prec->dinv.Read() : nullptr;  // identity smoother otherwise
    // device scalars ping-pong between iterations: {nom, betanom} in scal[0..1], denominators in scal[2..3]
    HCK(cudaMemcpyAsync(r.d, b.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    HCK(cudaMemsetAsync(x.d, 0, sizeof(double) * n, stream));
    if (dinv) prec->Mult(r, d);
    else HCK(cudaMemcpyAsync(d.d, r.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    double nom = comm->Dot(d.d, r.d);
    const double r0 = std::max(nom * rel_tol * rel_tol, abs_tol * abs_tol);
    converged = 0;
    final_iter = 0;
    if (nom <= r0) { converged = 1; return; }
    double init[4] = {nom, 0.0, 0.0, 0.0};
    HCK(cudaMemcpyAsync(scal.d, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    HCK(cudaMemsetAsync(z.d, 0, sizeof(double) * n, stream));
    A->MultAccDot(d, z, scal.d + 2);
    {
      double den;
      HCK(cudaMemcpyAsync(&den, scal.d + 2, sizeof(double), cudaMemcpyDeviceToHost, stream));
      HCK(cudaStreamSynchronize(stream));
      if (den <= 0.0 && den == 0.0) return;
    }
    int i = 1;
    final_iter = max_iter;
    for (;;) {
      const int a = (i - 1) & 1;
      double* d_nom = scal.d + a; double* d_bet = scal.d + (a ^ 1);
      double* d_den = scal.d + 2 + a; double* d_den_next = scal.d + 2 + (a ^ 1);
      // step 1 and (speculatively: it touches neither x nor r) the next direction d = M r + beta d with z zeroed
      const bool fused = comm->CgFused(x.d, r.d, d.d, z.d, dinv, d_nom, d_den, d_den_next, d_bet);
      if (!fused) comm->CgStep1(x.d, r.d, d.d, z.d, dinv, d_nom, d_den, d_den_next, d_bet);
      const int slot = i & 7;
      if (!fused) {
        HCK(cudaMemcpyAsync(h_bet + slot, d_bet, sizeof(double), cudaMemcpyDeviceToHost, stream));
        HCK(cudaEventRecord(ev[slot], stream));
      }
      const bool last = (i + 1 > max_iter);
      if (!last) {
        // speculative operator apply
        if (!fused) k_cg_step2<<<nb(n), 256, 0, stream>>>(d.d, r.d, dinv, z.d, d_bet, d_nom, n);
        A->MultAccDot(d, z, d_den_next);
      }
      double betanom;
      if (fused) betanom = comm->WaitBetanom();
      else { HCK(cudaEventSynchronize(ev[slot])); betanom = h_bet[slot]; }
      // mfem::CGSolver leaves the loop on betanom < 0 and on a non-positive denominator; a NaN (den == 0, or a
      // peer-memory exchange that timed out and poisoned its result) must not run max_iter applies on garbage
      if (!std::isfinite(betanom) || betanom < 0.0) { converged = 0; final_iter = i; break; }
      if (betanom <= r0) { converged = 1; final_iter = i; break; }
      if (++i > max_iter) break;
    }
    total_iters += std::min(i, max_iter);
  }
};

// src/mechanics_solver.cpp:39-143 (NR) and :155-280 (NRLS)
class ExaNewtonSolver {
 public:
  SlabComm* comm;
  cudaStream_t stream;
  const NonlinearMechOperator* oper_mech = nullptr;
  CGSolver* prec = nullptr;
  MechOperatorJacobiSmoother* smoother = nullptr;
  double rel_tol = 5e-5, abs_tol = 5e-10;
  int max_iter = 25, print_level = -1;
  bool line_search = false;
  mutable Vector r, c, x_prev;
  mutable int final_iter = 0, converged = 0;
  mutable double final_norm = 0.0;
  ExaNewtonSolver(SlabComm* cm, cudaStream_t s, long n) : comm(cm), stream(s), r(n), c(n), x_prev(n) {}
  double Norm(const Vector& v) const { return std::sqrt(comm->Dot(v.d, v.d)); }
  void CGSolverSolve(Operator& op, const Vector& b, Vector& x) const { prec->SetOperator(op); prec->Mult(b, x); }
  void Mult(Vector& x) const {
    const long n = x.Size();
    oper_mech->Mult(x, r);
    double norm = Norm(r);
    const double norm0 = norm;
    const double norm_max = std::max(rel_tol * norm, abs_tol);
    double scale = 1.0;
    int it;
    for (it = 0; true; ++it) {
      if (!std::isfinite(norm)) {
        if (comm->PeerError()) throw Abort{"peer-memory collective timed out waiting for a neighbour"};
        throw Abort{"Newton: residual norm is not finite"};
      }
      if (print_level >= 0 && comm->rank == 0) {
        std::printf("Newton iteration %2d : ||r|| = %g", it, norm);
        if (it > 0) std::printf(", ||r||/||r_0|| = %g", norm / norm0);
        std::printf("\n");
      }
      if (norm <= norm_max) { converged = 1; break; }
      if (it >= max_iter) { converged = 0; break; }
      Operator& grad = oper_mech->GetGradient(x);
      if (smoother->refresh) smoother->Setup(oper_mech->diag, oper_mech->ess_dev(), oper_mech->nnodes);
      prec->SetOperator(grad);
      prec->Mult(r, c);
      if (line_search) {
        HCK(cudaMemcpyAsync(x_prev.d, x.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
        k_axpby<<<nb(n), 256, 0, stream>>>(x.d, c.d, -1.0, 1.0, n);
        oper_mech->Mult(x, r);
        const double q1 = norm, q3 = Norm(r);
        HCK(cudaMemcpyAsync(x.d, x_prev.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
        k_axpby<<<nb(n), 256, 0, stream>>>(x.d, c.d, -0.5, 1.0, n);
        oper_mech->Mult(x, r);
        const double q2 = Norm(r);
        const double eps = (3.0 * q1 - 4.0 * q2 + q3) / (4.0 * (q1 - 2.0 * q2 + q3));
        if ((q1 - 2.0 * q2 + q3) > 0 && eps > 0 && eps < 1) scale = eps;
        else if (q3 < q1) scale = 1.0;
        else scale = 0.05;
        HCK(cudaMemcpyAsync(x.d, x_prev.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
      }
      k_axpby<<<nb(n), 256, 0, stream>>>(x.d, c.d, -scale, 1.0, n);
      oper_mech->Mult(x, r);
      const double norm_prev = norm;
      norm = Norm(r);
      if (!line_search) scale = (norm / norm_prev > 0.5) ? 0.5 : 1.0;
    }
    final_iter = it;
    final_norm = norm;
  }
};

}  // namespace exahost

using namespace exahost;

// ---------------------------------------------------------------- SystemDriver + C API ------
struct exahost_sim {
  exahost_config cfg;
  cudaStream_t stream = nullptr;
  exab200_ctx* ctx = nullptr;
  SlabComm comm;
  long nelems = 0, nnodes = 0, plane = 0;
  Vector stress0, stress1, matVars0, matVars1, matGrad;
  Vector x_beg, x_ref, v_sol, v_prev, ess_val, tmp, sums;
  std::vector<unsigned char> h_mask;
  std::unique_ptr<ExaCMechModel> model;
  std::unique_ptr<NonlinearMechOperator> oper;
  std::unique_ptr<MechOperatorJacobiSmoother> smoother;
  std::unique_ptr<CGSolver> cg;
  std::unique_ptr<ExaNewtonSolver> newton;
  double* h_pinned = nullptr;  // staging for host-buffer steps
  cudaEvent_t ev[4];
  long newton_total = 0;
  // velocity-gradient ("constant strain rate") boundary conditions
  Vector vg_mask_dev, vg_scratch;
  bool has_vgrad = false;
  VGrad vgrad{};

  // SystemDriver::UpdateVelocity (src/system_driver.cpp:327-427): velocity BCs overwrite their components; the
  // velocity-gradient ones get L (x - x_min) on the current coordinates, x_min = component-wise minimum of the
  // whole mesh, recomputed every call (MPI_Allreduce MIN in the reference)
  void UpdateVelocity() {
    k_set_ess<<<nb(3 * nnodes), 256, 0, stream>>>(v_sol.d, ess_val.d, oper->ess_dev(), nnodes, 0, nullptr);
    if (!has_vgrad) return;
    k_min3_partial<<<kRedBlocks, 256, 0, stream>>>(x_beg.d, nnodes, vg_scratch.d + 8);
    k_min3_final<<<1, 256, 0, stream>>>(vg_scratch.d + 8, kRedBlocks, vg_scratch.d);
    g_host_launches += 2;
    comm.AllReduceMinDevice(vg_scratch.d, 3);
    k_vgrad_vel<<<nb(nnodes), 256, 0, stream>>>(v_sol.d, x_beg.d, reinterpret_cast<const unsigned char*>(vg_mask_dev.d),
                                                vg_scratch.d, vgrad, nnodes);
  }
  // SystemDriver::SolveInit (src/system_driver.cpp:293-319)
  void SolveInit() {
    const long n = 3 * nnodes;
    Vector deltaF(n), b(n), x(n);
    // deltaF = ess ? v_sol - v_prev : 0
    k_set_ess<<<nb(n), 256, 0, stream>>>(deltaF.d, v_sol.d, oper->ess_dev(), nnodes, 1, v_prev.d);
    Operator& op = oper->GetUpdateBCsAction(v_prev, deltaF, b);
    newton->CGSolverSolve(op, b, x);
    // v_sol = -x + v_prev
    HCK(cudaMemcpyAsync(v_sol.d, v_prev.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    k_axpby<<<nb(n), 256, 0, stream>>>(v_sol.d, x.d, -1.0, 1.0, n);
    HCK(cudaStreamSynchronize(stream));
  }
  // SystemDriver::UpdateModel (src/system_driver.cpp:429-468): swap + volume-averaged stress
  void UpdateModel(double* avg_stress) {
    model->UpdateModelVars();
    model->UpdateStress();
    model->UpdateStateVars();
    XCK(exab200_vol_sum(ctx, oper->el_jac.Read(), stress0.Read(), 6, sums.Write(), stream));
    double h[7];
    comm.AllReduceFetch(sums.d, 7, h);
    for (int i = 0; i < 6; ++i) avg_stress[i] = h[i] / h[6];
  }
};

extern "C" {

const char* exahost_last_error(void) { return g_err.c_str(); }

int exahost_slab_layout(int nx, int ny, int nz_total, int rank, int nranks, long* out) {
  if (nx < 1 || ny < 1 || nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks || nz_total < nranks) {
    g_err = "exahost_slab_layout: need nx, ny >= 1, 1 <= nranks <= 8, 0 <= rank < nranks and at least one element layer per rank";
    return 1;
  }
  const long base = nz_total / nranks, rem = nz_total % nranks;      // the first `rem` ranks take one layer more
  const long z0 = rank * base + std::min<long>(rank, rem), nzl = base + (rank < rem ? 1 : 0);
  const long plane = (long)(nx + 1) * (ny + 1), nn = plane * (nzl + 1), layer = (long)nx * ny, ne = layer * nzl;
  out[0] = z0; out[1] = nzl; out[2] = nn; out[3] = plane;
  out[4] = (rank == nranks - 1) ? nn : nn - plane;
  out[5] = 0; out[6] = nn - plane;
  out[7] = ne; out[8] = rank > 0; out[9] = rank < nranks - 1;
  out[10] = (layer + 3) / 4; out[11] = (ne - layer) / 4;
  return 0;
}

int exahost_nccl_unique_id(void* out128) {
  if (!g_nccl.load()) { g_err = "NCCL library could not be loaded"; return 1; }
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_err = "ncclGetUniqueId failed"; return 1; }
  std::memcpy(out128, &id, sizeof(id));
  return 0;
}

int exahost_create(const exahost_config* cfg, exahost_sim** out) {
  exahost_sim* s = nullptr;
  try {
    s = new exahost_sim();
    s->cfg = *cfg;
    HCK(cudaSetDevice(cfg->device));
    HCK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    const long nx = cfg->nx, ny = cfg->ny, nzl = cfg->nz_local;
    long lay[12];
    if (exahost_slab_layout(cfg->nx, cfg->ny, cfg->nz_total, cfg->rank, cfg->nranks, lay)) throw Abort{g_err};
    if (lay[0] != cfg->z0 || lay[1] != nzl) throw Abort{"exahost_create: (z0, nz_local) is not this rank's slab of exahost_slab_layout"};
    s->nelems = nx * ny * nzl;
    s->plane = (nx + 1) * (ny + 1);
    s->nnodes = s->plane * (nzl + 1);
    // element -> node map, NATIVE hex vertex order (Mesh::MakeCartesian3D, x fastest)
    std::vector<int> e2n(8 * s->nelems);
    static const int hv[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
    for (long k = 0; k < nzl; ++k)
      for (long j = 0; j < ny; ++j)
        for (long i = 0; i < nx; ++i) {
          const long e = (k * ny + j) * nx + i;
          for (int a = 0; a < 8; ++a)
            e2n[e * 8 + a] = (int)(((k + hv[a][2]) * (ny + 1) + (j + hv[a][1])) * (nx + 1) + (i + hv[a][0]));
        }
    exab200_config ec;
    ec.xtal = cfg->xtal; ec.slip = cfg->slip; ec.nprops = cfg->nprops; ec.props = cfg->props; ec.temp_k = cfg->temp_k;
    ec.nelems = s->nelems; ec.nnodes = s->nnodes; ec.e2n = e2n.data(); ec.assembly = cfg->assembly; ec.integ = cfg->integ;
    ec.device = cfg->device;
    XCK(exab200_create(&ec, &s->ctx));
    // fused PA path on cubic crystals: K1 hands K2 the compact tangent record (11 % fewer operand bytes)
    if (cfg->assembly == EXAB200_PA && cfg->xtal != EXAB200_HCP) XCK(exab200_set_tangent_format(s->ctx, EXAB200_TANGENT_COMPACT));
    const int nsv = exab200_num_state_vars(s->ctx);
    const long npts = s->nelems * 8, n = 3 * s->nnodes;
    s->stress0.SetSize(npts * 6); s->stress1.SetSize(npts * 6);
    s->matVars0.SetSize(npts * nsv); s->matVars1.SetSize(npts * nsv);
    s->matGrad.SetSize(npts * 36);
    s->x_beg.SetSize(n); s->v_sol.SetSize(n); s->v_prev.SetSize(n); s->ess_val.SetSize(n); s->tmp.SetSize(n);
    s->sums.SetSize(48);
    HCK(cudaMemset(s->stress0.d, 0, sizeof(double) * npts * 6));
    HCK(cudaMemset(s->stress1.d, 0, sizeof(double) * npts * 6));
    HCK(cudaMemset(s->v_sol.d, 0, sizeof(double) * n));
    HCK(cudaMemset(s->v_prev.d, 0, sizeof(double) * n));
    HCK(cudaMemset(s->ess_val.d, 0, sizeof(double) * n));
    // coordinates (byNODES) of this slab
    {
      std::vector<double> xc(n);
      for (long k = 0; k <= nzl; ++k)
        for (long j = 0; j <= ny; ++j)
          for (long i = 0; i <= nx; ++i) {
            const long nd = (k * (ny + 1) + j) * (nx + 1) + i;
            xc[nd] = cfg->length[0] * i / nx;
            xc[s->nnodes + nd] = cfg->length[1] * j / ny;
            xc[2 * s->nnodes + nd] = cfg->length[2] * (k + cfg->z0) / cfg->nz_total;
          }
      HCK(cudaMemcpy(s->x_beg.d, xc.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
      s->x_ref.SetSize(n);
      HCK(cudaMemcpy(s->x_ref.d, xc.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    // history: setStateVarData (src/mechanics_driver.cpp:1058-1154: state-file values + per-grain quaternion
    // at offset 9) followed by init_state_vars
    {
      std::vector<double> h((size_t)npts * nsv, 0.0);
      for (long e = 0; e < s->nelems; ++e) {
        const int g = cfg->grain_ids[e] - 1;
        if (g < 0 || g >= cfg->ngrains) throw Abort{"grain id out of range"};
        for (int q = 0; q < 8; ++q)
          for (int i = 0; i < 4; ++i) h[(size_t)(e * 8 + q) * nsv + 9 + i] = cfg->quats[4 * g + i];
      }
      HCK(cudaMemcpy(s->matVars0.d, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
      XCK(exab200_hist_init(s->ctx, s->matVars0.d, s->stream));
      HCK(cudaStreamSynchronize(s->stream));
    }
    s->comm.Init(cfg->rank, cfg->nranks, cfg->nccl_id, s->stream, s->nnodes, s->plane);
    if (s->comm.n_owned != lay[4] || s->nnodes != lay[2] || s->nelems != lay[7]) throw Abort{"slab layout mismatch"};
    const Assembly as = cfg->assembly == EXAB200_EA ? Assembly::EA : Assembly::PA;
    s->model.reset(new ExaCMechModel(s->ctx, s->stream, &s->stress0, &s->stress1, &s->matGrad, &s->matVars0, &s->matVars1,
                                     cfg->nprops, as));
    s->oper.reset(new NonlinearMechOperator(s->ctx, &s->comm, s->stream, s->model.get(), s->nelems, s->nnodes, &s->x_beg));
    s->oper->ess_mask_dev.SetSize((s->nnodes + 7) / 8 + 1);
    s->h_mask.assign(s->nnodes, 0);
    HCK(cudaMemset(s->oper->ess_mask_dev.d, 0, s->nnodes));
    s->smoother.reset(new MechOperatorJacobiSmoother(s->nnodes, s->stream, cfg->true_jacobi != 0));
    s->cg.reset(new CGSolver(&s->comm, s->stream, n));
    // Optional L2 persistence window over the search direction d and the operator result z (written, gathered /
    // accumulated and read again within one CG iteration while 4+ GB of operands stream past them).  OFF by default:
    // measured on B200 at 128^3 it costs 2.7 % of the step (4.43 vs 4.31 s; the window is larger than the carve-out, so
    // a random subset persists and the rest is demoted to streaming) -- the evict_first hint on the operand stream
    // already protects the vectors.  EXAHOST_L2_PERSIST=1 enables it for experiments.
    {
      const char* env = std::getenv("EXAHOST_L2_PERSIST");
      cudaDeviceProp prop;
      if (env && env[0] == '1' && cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess &&
          prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        const size_t want = sizeof(double) * (size_t)s->cg->dz.n;
        const size_t carve = std::min((size_t)prop.persistingL2CacheMaxSize, want);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
          cudaStreamAttrValue attr;
          std::memset(&attr, 0, sizeof(attr));
          attr.accessPolicyWindow.base_ptr = s->cg->dz.d;
          attr.accessPolicyWindow.num_bytes = std::min(want, (size_t)prop.accessPolicyMaxWindowSize);
          attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)attr.accessPolicyWindow.num_bytes);
          attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
          attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
          if (cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
          else if (cfg->verbose) std::printf("L2 persistence: %.1f MB carve-out over a %.1f MB window\n", carve / 1e6, attr.accessPolicyWindow.num_bytes / 1e6);
        } else cudaGetLastError();
      }
    }
    s->cg->rel_tol = cfg->krylov_rel_tol; s->cg->abs_tol = cfg->krylov_abs_tol; s->cg->max_iter = cfg->krylov_iter;
    s->cg->prec = s->smoother.get();
    s->newton.reset(new ExaNewtonSolver(&s->comm, s->stream, n));
    s->newton->rel_tol = cfg->newton_rel_tol; s->newton->abs_tol = cfg->newton_abs_tol; s->newton->max_iter = cfg->newton_iter;
    s->newton->line_search = cfg->nl_solver == 1;
    s->newton->oper_mech = s->oper.get();
    s->newton->prec = s->cg.get();
    s->newton->smoother = s->smoother.get();
    s->oper->need_diag = s->smoother->refresh;
    s->oper->layer_elems = nx * ny;
    s->newton->print_level = cfg->verbose ? 0 : -1;
    HCK(cudaMallocHost(&s->h_pinned, sizeof(double) * n));
    for (int i = 0; i < 4; ++i) HCK(cudaEventCreate(&s->ev[i]));
    *out = s;
    return 0;
  } catch (const Abort& a) {
    g_err = a.msg;
    delete s;
    return 1;
  }
}

void exahost_destroy(exahost_sim* s) {
  if (!s) return;
  cudaSetDevice(s->cfg.device);
  cudaStreamSynchronize(s->stream);
  s->newton.reset(); s->cg.reset(); s->smoother.reset(); s->oper.reset(); s->model.reset();
  if (s->ctx) exab200_destroy(s->ctx);
  if (s->h_pinned) cudaFreeHost(s->h_pinned);
  cudaStream_t st = s->stream;
  delete s;
  cudaStreamDestroy(st);
}

// SystemDriver::UpdateEssBdr (src/system_driver.cpp:321-324): per-node component mask + prescribed values
int exahost_set_bcs(exahost_sim* s, const unsigned char* mask, const double* h_ess_val) {
  try {
    s->h_mask.assign(mask, mask + s->nnodes);
    s->oper->UpdateEssTDofs(mask);
    HCK(cudaMemcpy(s->oper->ess_mask_dev.d, mask, s->nnodes, cudaMemcpyHostToDevice));
    HCK(cudaMemcpy(s->ess_val.d, h_ess_val, sizeof(double) * 3 * s->nnodes, cudaMemcpyHostToDevice));
    return 0;
  } catch (const Abort& a) { g_err = a.msg; return 1; }
}

// SystemDriver::UpdateEssBdr for velocity-gradient attributes: per-node component mask of the dofs driven by
// BCs.essential_vel_grad (row-major L); mask == NULL switches them off.  The total essential mask of
// exahost_set_bcs must include these dofs.
int exahost_set_vgrad(exahost_sim* s, const unsigned char* mask_vgrad, const double* L9) {
  try {
    HCK(cudaSetDevice(s->cfg.device));
    s->has_vgrad = false;
    if (!mask_vgrad) return 0;
    s->vg_mask_dev.SetSize((s->nnodes + 7) / 8 + 1);
    s->vg_scratch.SetSize(8 + 3 * kRedBlocks);
    HCK(cudaMemcpy(s->vg_mask_dev.d, mask_vgrad, s->nnodes, cudaMemcpyHostToDevice));
    for (int i = 0; i < 9; ++i) s->vgrad.L[i] = L9[i];
    for (long i = 0; i < s->nnodes; ++i)
      if (mask_vgrad[i] & 7) { s->has_vgrad = true; break; }
    if (s->comm.nranks > 1) s->has_vgrad = true;  // another rank may hold the marked faces; the MIN is collective
    return 0;
  } catch (const Abort& a) { g_err = a.msg; return 1; }
}

// Time.Auto state of SystemDriver::Solve (src/system_driver.cpp:225-274)
struct AutoCtl { double dt_class, t, dt_min, dt_scale, t_final; int last_step; };

// One time step of the reference's loop (src/mechanics_driver.cpp:837-907).
//   bc_changed: run the SolveInit corrector first (src/mechanics_driver.cpp:866-878)
//   h_ess_val_in  (may be NULL): prescribed velocities for this step, HOST buffer, copied H2D inside the call
//   h_vel_out     (may be NULL): converged velocity, HOST buffer, copied D2H inside the call
//   out[16]: {newton_iters, pcg_iters, converged, model_setups, grad_mults, wall seconds, avg_stress[6],
//             device ms of the solve (CUDA events, copies excluded), device ms end to end (copies included),
//             dt actually taken, 0}
static int step_impl(exahost_sim* s, double dt, int bc_changed, const double* h_ess_val_in, double* h_vel_out, double* out,
                     AutoCtl* at) {
  try {
    HCK(cudaSetDevice(s->cfg.device));
    const long n = 3 * s->nnodes;
    auto t0 = std::chrono::steady_clock::now();
    const long pcg0 = s->cg->total_iters, ms0 = s->model->model_setups, gm0 = s->oper->grad_mults;
    HCK(cudaEventRecord(s->ev[0], s->stream));
    if (h_ess_val_in) {
      std::memcpy(s->h_pinned, h_ess_val_in, sizeof(double) * n);
      HCK(cudaMemcpyAsync(s->ess_val.d, s->h_pinned, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
    }
    HCK(cudaEventRecord(s->ev[1], s->stream));
    if (at) {  // src/mechanics_driver.cpp:845-855
      dt = std::min(at->dt_class, at->t_final - at->t);
      at->t += dt;
      at->last_step = std::fabs(at->t - at->t_final) <= std::fabs(1e-3 * dt);
    }
    s->model->SetModelDt(dt);
    if (bc_changed) {
      HCK(cudaMemcpyAsync(s->v_prev.d, s->v_sol.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, s->stream));
      s->UpdateVelocity();
      s->SolveInit();
    }
    s->UpdateVelocity();
    int newton_iters = 0;
    if (at) {
      if (at->last_step) at->dt_class = dt;
      const double dt_old = at->dt_class;
      HCK(cudaMemcpyAsync(s->tmp.d, s->v_sol.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, s->stream));  // xprev
      s->newton->Mult(s->v_sol);
      newton_iters += s->newton->final_iter;
      if (!s->newton->converged) {
        for (int retry = 0; !s->newton->converged && retry < 2; ++retry) {
          HCK(cudaMemcpyAsync(s->v_sol.d, s->tmp.d, sizeof(double) * n, cudaMemcpyDeviceToDevice, s->stream));
          at->dt_class *= at->dt_scale;
          if (at->dt_class < at->dt_min) at->dt_class = at->dt_min;
          dt = at->dt_class;
          s->model->SetModelDt(dt);
          s->newton->Mult(s->v_sol);
          newton_iters += s->newton->final_iter;
        }
        at->t = at->t - dt_old + at->dt_class;
        at->last_step = std::fabs(at->t - at->t_final) <= std::fabs(1e-3 * dt);
      }
      const double factor = ((double)s->newton->max_iter * at->dt_scale) / (double)s->newton->final_iter;
      at->dt_class *= factor;
      if (at->dt_class < at->dt_min) at->dt_class = at->dt_min;
    } else {
      s->newton->Mult(s->v_sol);
      newton_iters = s->newton->final_iter;
    }
    // Failures are made collective before anyone throws: the Newton verdict rests on all-reduced norms (identical on
    // every rank), the failed-point count and the peer-timeout flag are summed over the ranks here, so either every
    // rank aborts or none does (a rank that throws alone would leave its neighbours spinning in the next exchange).
    int nfail = 0;
    XCK(exab200_failed_points(s->ctx, s->stream, &nfail));
    double fl[2] = {(double)nfail, s->comm.PeerError() ? 1.0 : 0.0};
    if (s->comm.nranks > 1 && fl[1] == 0.0 && std::isfinite(s->newton->final_norm)) {
      HCK(cudaMemcpyAsync(s->sums.d, fl, sizeof(fl), cudaMemcpyHostToDevice, s->stream));
      s->comm.AllReduceFetch(s->sums.d, 2, fl);
    }
    if (fl[1] != 0.0 || s->comm.PeerError()) throw Abort{"peer-memory collective timed out waiting for a neighbour"};
    if (!s->newton->converged) throw Abort{"Newton Solver did not converge."};  // MFEM_VERIFY, src/system_driver.cpp:287
    if (fl[0] != 0.0) throw Abort{"material update failed at " + std::to_string((long)fl[0]) + " quadrature points (all ranks)"};
    double avg[6];
    s->UpdateModel(avg);
    // x_beg = x_cur (src/mechanics_driver.cpp:907): x_beg += dt * v
    k_axpby<<<nb(n), 256, 0, s->stream>>>(s->x_beg.d, s->v_sol.d, dt, 1.0, n);
    HCK(cudaEventRecord(s->ev[2], s->stream));
    if (h_vel_out) {
      HCK(cudaMemcpyAsync(s->h_pinned, s->v_sol.d, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
      HCK(cudaStreamSynchronize(s->stream));
      std::memcpy(h_vel_out, s->h_pinned, sizeof(double) * n);
    }
    HCK(cudaEventRecord(s->ev[3], s->stream));
    HCK(cudaStreamSynchronize(s->stream));
    auto t1 = std::chrono::steady_clock::now();
    s->oper->tm_grad_mult.Harvest();
    s->oper->tm_model_setup.Harvest();
    float ms_dev = 0.f, ms_e2e = 0.f;
    HCK(cudaEventElapsedTime(&ms_dev, s->ev[1], s->ev[2]));
    HCK(cudaEventElapsedTime(&ms_e2e, s->ev[0], s->ev[3]));
    out[12] = ms_dev;
    out[13] = ms_e2e;
    out[14] = dt;
    out[15] = 0.0;
    s->newton_total += newton_iters;
    out[0] = s->newton->final_iter;
    out[1] = (double)(s->cg->total_iters - pcg0);
    out[2] = s->newton->converged;
    out[3] = (double)(s->model->model_setups - ms0);
    out[4] = (double)(s->oper->grad_mults - gm0);
    out[5] = std::chrono::duration<double>(t1 - t0).count();
    for (int i = 0; i < 6; ++i) out[6 + i] = avg[i];
    return 0;
  } catch (const Abort& a) { g_err = a.msg; return 1; }
}

int exahost_step(exahost_sim* s, double dt, int bc_changed, const double* h_ess_val_in, double* h_vel_out, double* out) {
  return step_impl(s, dt, bc_changed, h_ess_val_in, h_vel_out, out, nullptr);
}

// Time.Auto step: ctl = {dt_class, t, dt_min, dt_scale, t_final, last_step}: dt_class / t / last_step are updated in
// place exactly like SystemDriver::Solve + the driver loop do (src/system_driver.cpp:225-274,
// src/mechanics_driver.cpp:845-889); out[14] = the step size taken.
int exahost_step_auto(exahost_sim* s, double* ctl6, int bc_changed, const double* h_ess_val_in, double* h_vel_out,
                      double* out) {
  AutoCtl at{ctl6[0], ctl6[1], ctl6[2], ctl6[3], ctl6[4], 0};
  const int rc = step_impl(s, 0.0, bc_changed, h_ess_val_in, h_vel_out, out, &at);
  ctl6[0] = at.dt_class; ctl6[1] = at.t; ctl6[5] = (double)at.last_step;
  return rc;
}

// copy quadrature state to HOST buffers (parity checks): which = 0 stress0 (6/pt), 1 matVars0 (nsv/pt), 2 v_sol, 3 x_beg
int exahost_get(exahost_sim* s, int which, double* h_out) {
  try {
    HCK(cudaSetDevice(s->cfg.device));
    HCK(cudaStreamSynchronize(s->stream));
    const Vector* v = which == 0 ? &s->stress0 : which == 1 ? &s->matVars0 : which == 2 ? &s->v_sol : &s->x_beg;
    HCK(cudaMemcpy(h_out, v->d, sizeof(double) * v->n, cudaMemcpyDeviceToHost));
    return 0;
  } catch (const Abort& a) { g_err = a.msg; return 1; }
}

// Additional volume averages of SystemDriver::UpdateModel (src/system_driver.cpp:470-553), evaluated on the state
// just committed by exahost_step: out[0] = volume integral of the plastic work (not divided by the volume, as the
// reference prints it), out[1..9] = volume-averaged deformation gradient F(i,t) at [1 + t*3 + i]
// (CalculateDeformationGradient, src/mechanics_operator.cpp:393-427), out[10..15] = volume-averaged D^p in Voigt order.
int exahost_extra_avgs(exahost_sim* s, double* out16) {
  try {
    HCK(cudaSetDevice(s->cfg.device));
    const int nsv = exab200_num_state_vars(s->ctx);
    const long npts = s->nelems * 8;
    double h[48];
    XCK(exab200_vol_sum(s->ctx, s->oper->el_jac.Read(), s->matVars0.Read(), nsv, s->sums.Write(), s->stream));
    s->comm.AllReduceFetch(s->sums.d, nsv + 1, h);
    out16[0] = h[2];
    Vector jac_ref(s->nelems * 72), q9(npts * 9);
    s->oper->CalculateDeformationGradient(s->x_ref, s->x_beg, jac_ref, q9);
    XCK(exab200_vol_sum(s->ctx, s->oper->el_jac.Read(), q9.Read(), 9, s->sums.Write(), s->stream));
    s->comm.AllReduceFetch(s->sums.d, 10, h);
    for (int i = 0; i < 9; ++i) out16[1 + i] = h[i] / h[9];
    // the reference reads matVars1 after the swap (src/mechanics_ecmech.hpp:308): previous step's state
    XCK(exab200_calc_dp(s->ctx, s->matVars1.Read(), q9.Write(), s->stream));
    XCK(exab200_vol_sum(s->ctx, s->oper->el_jac.Read(), q9.Read(), 9, s->sums.Write(), s->stream));
    s->comm.AllReduceFetch(s->sums.d, 10, h);
    const int pick[6] = {0, 4, 8, 5, 2, 1};
    for (int i = 0; i < 6; ++i) out16[10 + i] = h[pick[i]] / h[9];
    HCK(cudaStreamSynchronize(s->stream));
    return 0;
  } catch (const Abort& a) { g_err = a.msg; return 1; }
}

long exahost_counter(exahost_sim* s, int which) {
  switch (which) {
    case 0: return exab200_launch_count(s->ctx) + g_host_launches;
    case 1: return s->comm.n_allreduce;
    case 2: return s->comm.n_halo;
    case 3: return s->model->model_setups;
    case 4: return s->oper->grad_mults;
    case 5: return s->cg->total_iters;
    case 6: return s->newton_total;
    case 7: return exab200_num_state_vars(s->ctx);
  }
  return -1;
}

// per-kernel CUDA-event timing: which = 0 gradient apply (memset + kernel), 1 material update
int exahost_kernel_timing(exahost_sim* s, int enable) {
  s->oper->tm_grad_mult.enabled = s->oper->tm_model_setup.enabled = enable != 0;
  return 0;
}
int exahost_kernel_time(exahost_sim* s, int which, double* total_ms, long* count, int reset) {
  KernelTimer& t = which == 0 ? s->oper->tm_grad_mult : s->oper->tm_model_setup;
  *total_ms = t.total_ms;
  *count = t.count;
  if (reset) { t.total_ms = 0.0; t.count = 0; }
  return 0;
}
int exahost_set_tuning(exahost_sim* s, int ctas_per_sm, int variant) {
  // 95 / 96: reference 36-entry tangent layout / compact records (only meaningful before the first step)
  // 97 / 98: deterministic (owner-computes) scatter on / off; the exchange then stays a separate kernel
  if (variant == 97 || variant == 98) { s->comm.halo_fused = -1; return exab200_set_deterministic(s->ctx, variant == 97); }
  if (variant == 95 || variant == 96) return exab200_set_tangent_format(s->ctx, variant == 96 ? EXAB200_TANGENT_COMPACT : EXAB200_TANGENT_VOIGT36);
  return exab200_set_tuning(s->ctx, ctas_per_sm, variant);
}

// NVLink peer-memory collectives: every rank publishes the CUDA-IPC handle of its mailbox, the caller gathers
// the handles of all ranks (torch.distributed) and hands them back.
int exahost_comm_handle(exahost_sim* s, void* out64) {
  try { HCK(cudaSetDevice(s->cfg.device)); s->comm.GetHandle(out64); return 0; }
  catch (const Abort& a) { g_err = a.msg; return 1; }
}
int exahost_set_peers(exahost_sim* s, const void* handles) {
  try { HCK(cudaSetDevice(s->cfg.device)); s->comm.SetPeers(handles); return 0; }
  catch (const Abort& a) { g_err = a.msg; return 1; }
}

void* exahost_stream(exahost_sim* s) { return s->stream; }
void* exahost_ctx(exahost_sim* s) { return s->ctx; }

}  // extern "C"

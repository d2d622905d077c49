// NVLink peer-memory collectives: device-side protocol shared by the host layer's exchange kernels (host_sim.cu) and the
// operator kernel that folds the interface-plane exchange into the apply (k_operator.cuh, k_grad_mult_pa_c<..., HALO>).
//
// Every rank owns a small "mailbox" in device memory that its peers map through CUDA IPC.  The latency-critical
// exchanges of the CG loop store straight into the peers' mailboxes over NVLink and synchronise through
// release/acquire flags there:
//   * interface-plane sum with the z-neighbours
//   * all-reduce of up to 8 scalars over all ranks, summed in rank order => bitwise identical on every rank and
//     run to run
// Sequence numbers are monotonic and slots are double-buffered by parity, which is sufficient because a rank can only
// be one collective ahead of a peer it exchanges with.
// Mailbox layout (doubles): [0,64) flags as u64: 0 halo-from-lo, 1 halo-from-hi, 8+r scalar-from-rank-r,
// 16 block counter, 17 error;  [64,192) scalar slots [par][rank][8];  [192, 192+12*plane) halo slots
// [from-lo | from-hi][par][3*plane].
#pragma once
#include <cuda_runtime.h>

namespace exab_p2p {

constexpr int kMbScal = 64, kMbHalo = 192;

// all ranks' mailboxes + the spin limit (clock64 ticks): a lost peer becomes an error instead of a hang -- the waiter
// raises the mailbox error flag AND poisons what it was about to produce with NaN, so that the CG scalars turn
// non-finite and the host loop stops at its next stopping test instead of iterating on stale data
struct PeerTable {
  double* p[8];
  long long spin_limit;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long seq, double* mb, long long limit) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq) {
    if (clock64() - t0 > limit) { reinterpret_cast<unsigned long long*>(mb)[17] = 1ull; return false; }
  }
  return true;
}
__device__ __forceinline__ double ld_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double quiet_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// one warp: in-place sum of val[0..n) over all ranks through the peers' scalar slots (rank order)
__device__ __forceinline__ void warp_allreduce_p2p(double* val, const PeerTable& peers, int rank, int nranks, int n,
                                                   unsigned long long seq, int lane) {
  const int par = (int)(seq & 1);
  double* mb = peers.p[rank];
  if (lane < nranks) {
    double* dst = peers.p[lane] + kMbScal + (par * 8 + rank) * 8;
    for (int k = 0; k < n; ++k) dst[k] = val[k];
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned long long*>(peers.p[lane]) + 8 + rank, seq);
  }
  const bool ok = (lane < nranks) ? spin_until(reinterpret_cast<unsigned long long*>(mb) + 8 + lane, seq, mb, peers.spin_limit) : true;
  const bool all_ok = __all_sync(0xffffffffu, ok);
  __syncwarp();
  if (lane < n) {
    double s = 0.0;
    for (int r = 0; r < nranks; ++r) s += ld_cg(&mb[kMbScal + (par * 8 + r) * 8 + lane]);
    val[lane] = all_ok ? s : quiet_nan();
  }
}

// Interface-plane exchange by `nblk` cooperating blocks (block index `blk`): push this rank's bottom / top plane of v
// into the lower / upper neighbour's mailbox, flag them once all blocks have pushed, wait for the neighbours' planes and
// add them.  counter: a u64 in this rank's mailbox (flags[16]).
__device__ __forceinline__ void halo_exchange_blocks(double* __restrict__ v, double* mb, double* lo, double* hi, long nn, long plane,
                                                     unsigned long long seq, long long spin_limit, unsigned blk, unsigned nblk,
                                                     bool* s_ok) {
  const int par = (int)(seq & 1);
  const long n3 = 3 * plane, top = nn - plane;
  // kU independent elements per thread and pass: the exchange blocks are few (latency-bound loads)
  constexpr int kU = 8;
  const long step = (long)nblk * blockDim.x, first = (long)blk * blockDim.x + threadIdx.x;
  for (int c = 0; c < 3; ++c) {
    const long vo = c * nn, mo = c * plane;
    double* dlo = lo ? lo + kMbHalo + (2 + par) * n3 + mo : nullptr;
    double* dhi = hi ? hi + kMbHalo + (0 + par) * n3 + mo : nullptr;
    for (long n = first; n < plane; n += kU * step) {
      double a[kU], b[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long m = n + u * step;
        if (m < plane) {
          if (lo) a[u] = ld_cg(&v[vo + m]);
          if (hi) b[u] = ld_cg(&v[vo + top + m]);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long m = n + u * step;
        if (m < plane) {
          if (lo) dlo[m] = a[u];
          if (hi) dhi[m] = b[u];
        }
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(mb);
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(&flags[16], 1ull);
    if (prev == nblk - 1) {
      flags[16] = 0ull;
      __threadfence_system();
      if (lo) st_release_sys(reinterpret_cast<unsigned long long*>(lo) + 1, seq);  // I am the lower one's "hi"
      if (hi) st_release_sys(reinterpret_cast<unsigned long long*>(hi) + 0, seq);  // I am the upper one's "lo"
    }
    bool ok = true;
    if (lo) ok = spin_until(&flags[0], seq, mb, spin_limit) && ok;
    if (hi) ok = spin_until(&flags[1], seq, mb, spin_limit) && ok;
    *s_ok = ok;
  }
  __syncthreads();
  const double poison = *s_ok ? 0.0 : quiet_nan();
  for (int c = 0; c < 3; ++c) {
    const long vo = c * nn, mo = c * plane;
    const double* slo = mb + kMbHalo + (0 + par) * n3 + mo;
    const double* shi = mb + kMbHalo + (2 + par) * n3 + mo;
    for (long n = first; n < plane; n += kU * step) {
      double a[kU], b[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long m = n + u * step;
        if (m < plane) {
          if (lo) a[u] = ld_cg(&v[vo + m]) + ld_cg(&slo[m]);
          if (hi) b[u] = ld_cg(&v[vo + top + m]) + ld_cg(&shi[m]);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long m = n + u * step;
        if (m < plane) {
          if (lo) v[vo + m] = a[u] + poison;
          if (hi) v[vo + top + m] = b[u] + poison;
        }
      }
    }
  }
}

}  // namespace exab_p2p

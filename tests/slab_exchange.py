"""Test harness for the z-slab exchange semantics (world-size-2 gloo runs on CPU, tests/test_parallel_gloo.py).

The partition, ownership and interface-plane index sets come from the C++ host layer itself (exahost_slab_layout,
csrc/host_sim.cu: the numbers exahost_create sizes its vectors with and the exchange kernels index with); this module
only states the two exchanges -- interface-plane sum with the z-neighbours, owned-dof dot product -- with
`torch.distributed` tensors so that they can be exercised without GPUs and compared with the single-domain oracle.

Layout: rank r owns element layers [z0, z0 + nz_local) and keeps a local byNODES L-vector over node planes
z0 .. z0 + nz_local (both interface planes included).  Uniquely-owned dofs = all local nodes except the top plane
(the last rank owns its top plane too)."""
import numpy as np

from exaconstit_b200 import host


class SlabLayout:
    def __init__(self, nx, ny, nz, rank, nranks):
        self.nx, self.ny, self.nz, self.rank, self.nranks = nx, ny, nz, rank, nranks
        lay = host.slab_layout(nx, ny, nz, rank, nranks)
        self.z0, self.nzl = lay["z0"], lay["nz_local"]
        self.z1 = self.z0 + self.nzl
        self.plane, self.nnodes, self.nelems, self.n_owned = lay["plane"], lay["nnodes"], lay["nelems"], lay["n_owned"]
        self.lo_offset, self.hi_offset = lay["lo_offset"], lay["hi_offset"]
        self.has_lo, self.has_hi = bool(lay["has_lo"]), bool(lay["has_hi"])
        self.nn_global = self.plane * (nz + 1)

    # ---- global <-> local field maps (byNODES L-vectors, x-fastest element arrays) ----
    def local_nodes_of_global(self, xg):
        xg = np.asarray(xg).reshape(3, self.nn_global)
        return xg[:, self.z0 * self.plane:(self.z1 + 1) * self.plane].reshape(-1).copy()

    def local_elems_of_global(self, eg, per_elem=1):
        eg = np.asarray(eg).reshape(self.nx * self.ny * self.nz, per_elem)
        return eg[self.nx * self.ny * self.z0: self.nx * self.ny * self.z1].reshape(-1).copy()

    def plane_slices(self, which):
        """Index arrays (into the local L-vector) of the bottom ('lo') or top ('hi') interface plane."""
        off = self.lo_offset if which == "lo" else self.hi_offset
        return np.concatenate([c * self.nnodes + off + np.arange(self.plane) for c in range(3)])

    # ---- exchanges, stated with torch.distributed (any backend) ----
    def halo_sum(self, v, dist):
        """Sum the partial results on the interface planes with the z-neighbours (both copies end equal)."""
        import torch
        if self.nranks == 1:
            return v
        lo, hi = self.has_lo, self.has_hi
        ilo, ihi = torch.as_tensor(self.plane_slices("lo")), torch.as_tensor(self.plane_slices("hi"))
        reqs, rlo, rhi = [], None, None
        if lo:
            slo = v[ilo].contiguous()
            rlo = torch.empty_like(slo)
            reqs += [dist.isend(slo, self.rank - 1), dist.irecv(rlo, self.rank - 1)]
        if hi:
            shi = v[ihi].contiguous()
            rhi = torch.empty_like(shi)
            reqs += [dist.isend(shi, self.rank + 1), dist.irecv(rhi, self.rank + 1)]
        for r in reqs:
            r.wait()
        if lo:
            v[ilo] += rlo
        if hi:
            v[ihi] += rhi
        return v

    def dot(self, a, b, dist):
        """Global dot product counting every shared node once."""
        import torch
        s = torch.zeros(1, dtype=a.dtype)
        for c in range(3):
            sl = slice(c * self.nnodes, c * self.nnodes + self.n_owned)
            s += torch.dot(a[sl], b[sl])
        if self.nranks > 1:
            dist.all_reduce(s)
        return float(s)

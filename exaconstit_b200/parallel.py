"""Host-side logic of the z-slab element partition used across GPUs (one process per GPU).

The C++ `SlabComm` (csrc/host_sim.cu) implements these semantics with NCCL on device buffers; this
module states them with `torch.distributed` tensors so they can be exercised with the gloo backend on
CPU (world_size 2) and so that Python-side drivers can scatter/gather global fields.

Layout: rank r owns element layers [z0[r], z0[r+1]) and keeps a local byNODES L-vector over node planes
z0[r] .. z0[r+1] (both interface planes included).  Uniquely-owned dofs = all local nodes except the top
plane (the last rank owns its top plane too)."""
import numpy as np

from . import voxel


class SlabLayout:
    def __init__(self, nx, ny, nz, rank, nranks):
        self.nx, self.ny, self.nz, self.rank, self.nranks = nx, ny, nz, rank, nranks
        z0s = voxel.slab_partition(nz, nranks)
        self.z0, self.z1 = int(z0s[rank]), int(z0s[rank + 1])
        self.nzl = self.z1 - self.z0
        self.plane = (nx + 1) * (ny + 1)
        self.nnodes = self.plane * (self.nzl + 1)
        self.nelems = nx * ny * self.nzl
        self.n_owned = self.nnodes if rank == nranks - 1 else self.nnodes - self.plane
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
        off = 0 if which == "lo" else self.nnodes - self.plane
        return np.concatenate([c * self.nnodes + off + np.arange(self.plane) for c in range(3)])

    # ---- exchanges, stated with torch.distributed (any backend) ----
    def halo_sum(self, v, dist):
        """Sum the partial results on the interface planes with the z-neighbours (both copies end equal)."""
        import torch
        if self.nranks == 1:
            return v
        lo, hi = self.rank > 0, self.rank < self.nranks - 1
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

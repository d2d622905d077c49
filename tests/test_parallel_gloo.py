"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: the C++ host layer's slab partition, ownership and
interface-plane index sets (exahost_slab_layout) driving interface-plane sums and owned-dof dot products over
torch.distributed, checked against the single-domain oracle operator."""
import os

import numpy as np
import pytest


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        sys.path.insert(0, os.path.join(root, "tests"))
        from exaconstit_b200 import voxel
        import slab_exchange as parallel
        from oracle import orc
        n = (4, 3, 5)
        nx, ny, nz = n
        rng = np.random.default_rng(0)  # same stream on both ranks => same global fields
        e2n_g, coords_g = voxel.voxel_mesh(nx, ny, nz)
        nn_g, ne_g = coords_g.size // 3, nx * ny * nz
        coords_g = coords_g + 0.05 * (rng.random(coords_g.size) - 0.5)
        S = rng.normal(size=(ne_g * 8, 6, 6))
        k36_g = (S + S.transpose(0, 2, 1) + 10 * np.eye(6)).reshape(ne_g, 8 * 36)
        x_g, z_g = rng.normal(size=3 * nn_g), rng.normal(size=3 * nn_g)
        G, W = orc.hex8_dshape()
        dt = 0.7
        # single-domain reference
        jac_g = orc.jacobians(G, orc.gather(e2n_g, coords_g))
        y_ref = orc.scatter_add(e2n_g, orc.grad_mult_pa(dt, jac_g, W, G, k36_g.ravel(), orc.gather(e2n_g, x_g)), nn_g)
        dot_ref = float(x_g @ z_g)
        # this rank's slab
        lay = parallel.SlabLayout(nx, ny, nz, rank, world)
        e2n_l, _ = voxel.voxel_mesh(nx, ny, lay.nzl, z0=lay.z0, nz_total=nz)
        coords_l = lay.local_nodes_of_global(coords_g)
        k36_l = lay.local_elems_of_global(k36_g, 8 * 36)
        x_l, z_l = lay.local_nodes_of_global(x_g), lay.local_nodes_of_global(z_g)
        jac_l = orc.jacobians(G, orc.gather(e2n_l, coords_l))
        assert np.allclose(jac_l, lay.local_elems_of_global(jac_g.reshape(ne_g, 72), 72), rtol=0, atol=1e-15)
        y_l = orc.scatter_add(e2n_l, orc.grad_mult_pa(dt, jac_l, W, G, k36_l, orc.gather(e2n_l, x_l)), lay.nnodes)
        y_l = lay.halo_sum(torch.tensor(y_l), dist).numpy()
        err = np.abs(y_l - lay.local_nodes_of_global(y_ref)).max() / np.abs(y_ref).max()
        d = lay.dot(torch.tensor(x_l), torch.tensor(z_l), dist)
        derr = abs(d - dot_ref) / abs(dot_ref)
        # essential BC masks of the slab agree with the global ones
        mg, vg = voxel.essential_bcs(nx, ny, nz, [1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0]] * 3 + [[0, 0, 1e-3]])
        ml, vl = voxel.essential_bcs(nx, ny, lay.nzl, [1, 2, 3, 4], [3, 1, 2, 3], [[0, 0, 0]] * 3 + [[0, 0, 1e-3]],
                                     z0=lay.z0, nz_total=nz)
        plane = lay.plane
        ok_mask = np.array_equal(ml, mg[lay.z0 * plane:(lay.z1 + 1) * plane])
        ok_val = np.array_equal(vl, lay.local_nodes_of_global(vg))
        owned = torch.tensor([float(lay.n_owned)])
        dist.all_reduce(owned)
        q.put((rank, err, derr, ok_mask, ok_val, float(owned) == nn_g))
    finally:
        dist.destroy_process_group()


def test_slab_exchange_world2(orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, derr, ok_mask, ok_val, ok_owned in res:
        assert err < 1e-13, (rank, err)
        assert derr < 1e-13, (rank, derr)
        assert ok_mask and ok_val and ok_owned


def test_slab_partition_covers_mesh():
    from exaconstit_b200 import host, voxel
    import slab_exchange as parallel
    for nz, nr in [(5, 2), (128, 8), (7, 3), (4, 4), (130, 8)]:
        z0 = voxel.slab_partition(nz, nr)
        assert z0[0] == 0 and z0[-1] == nz and np.all(np.diff(z0) >= 1)
        lays = [parallel.SlabLayout(3, 2, nz, r, nr) for r in range(nr)]
        assert [l.z0 for l in lays] == [int(z) for z in z0[:-1]] and lays[-1].z1 == nz     # C++ and Python partitions agree
        assert sum(l.nelems for l in lays) == 3 * 2 * nz
        assert sum(l.n_owned for l in lays) == 4 * 3 * (nz + 1)
        for r, l in enumerate(lays):
            assert l.has_lo == (r > 0) and l.has_hi == (r < nr - 1)
            assert l.lo_offset == 0 and l.hi_offset == l.nnodes - l.plane
            lay = host.slab_layout(3, 2, nz, r, nr)
            # the operator kernel's boundary-first order: leading tiles cover the bottom layer, the tail the top layer
            assert lay["lo_tiles"] * 4 >= 6 and lay["hi_tile_start"] * 4 <= l.nelems - 6
    with pytest.raises(host.HostError):
        host.slab_layout(3, 2, 4, 0, 9)          # more ranks than one NVLink box holds
    with pytest.raises(host.HostError):
        host.slab_layout(3, 2, 2, 0, 4)          # fewer layers than ranks


def test_slab_bc_masks_tile_the_global_masks():
    """Essential / velocity-gradient masks and prescribed values built per z-slab (what each rank hands to
    exahost_set_bcs / exahost_set_vgrad) agree with the global ones on every node, duplicated interface planes
    included."""
    from exaconstit_b200 import voxel
    nx, ny, nz, nranks = 4, 3, 7, 3
    ids, comps = [1, 2, 3, 4, 6], [3, -1, -2, -3, 7]
    vals = np.arange(15, dtype=float).reshape(5, 3)
    gm, gv = voxel.essential_bcs(nx, ny, nz, ids, comps, vals)
    gg = voxel.vgrad_mask(nx, ny, nz, ids, comps)
    plane = (nx + 1) * (ny + 1)
    z0 = voxel.slab_partition(nz, nranks)
    assert z0[0] == 0 and z0[-1] == nz and np.all(np.diff(z0) >= 2)
    nn_g = plane * (nz + 1)
    for r in range(nranks):
        nzl = int(z0[r + 1] - z0[r])
        m, v = voxel.essential_bcs(nx, ny, nzl, ids, comps, vals, int(z0[r]), nz)
        g = voxel.vgrad_mask(nx, ny, nzl, ids, comps, int(z0[r]), nz)
        lo, hi = int(z0[r]) * plane, (int(z0[r]) + nzl + 1) * plane
        assert np.array_equal(m, gm[lo:hi]) and np.array_equal(g, gg[lo:hi])
        nn_l = plane * (nzl + 1)
        for d in range(3):
            assert np.array_equal(v[d * nn_l:(d + 1) * nn_l], gv[d * nn_g + lo:d * nn_g + hi])
    assert np.all((gg & ~gm) == 0)      # velocity-gradient dofs are a subset of the essential ones

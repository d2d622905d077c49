"""Voxel-mesh helpers of the host layer: the auto-generated Cartesian hex mesh the reference builds
with Mesh::MakeCartesian3D(..., sfc_ordering=false) (src/mechanics_driver.cpp:247-253), its boundary
attribute convention (src/mechanics_driver.cpp:1196-1231: 1 z_min, 2 x_min, 3 y_min, 4 z_max, 5 x_max,
6 y_max), synthetic Voronoi grains and the z-slab partition used across GPUs."""
import numpy as np

_HEX = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])


def voxel_mesh(nx, ny, nz, length=(1.0, 1.0, 1.0), z0=0, nz_total=None):
    """Element->node map (NATIVE hex vertex order) and nodal coordinates (byNODES) of an nx*ny*nz slab whose
    first element layer is global layer z0 of an nz_total-layer mesh."""
    nz_total = nz if nz_total is None else nz_total
    px, py, pz = nx + 1, ny + 1, nz + 1
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    coords = np.concatenate([(length[0] * i / nx).ravel(), (length[1] * j / ny).ravel(),
                             (length[2] * (k + z0) / nz_total).ravel()]).astype(np.float64)
    ek, ej, ei = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    e2n = np.empty((nx * ny * nz, 8), dtype=np.int32)
    for a in range(8):
        e2n[:, a] = (((ek + _HEX[a, 2]) * py + (ej + _HEX[a, 1])) * px + (ei + _HEX[a, 0])).ravel()
    return e2n.ravel(), coords


def face_nodes(nx, ny, nz, attr, z0=0, nz_total=None):
    """Local node ids on boundary attribute `attr` of the GLOBAL box for a slab [z0, z0+nz)."""
    nz_total = nz if nz_total is None else nz_total
    px, py, pz = nx + 1, ny + 1, nz + 1
    k, j, i = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
    kg = k + z0
    sel = {1: kg == 0, 2: i == 0, 3: j == 0, 4: kg == nz_total, 5: i == nx, 6: j == ny}[attr]
    return np.nonzero(sel.ravel())[0]


_CMP = {0: (0, 0, 0), 1: (1, 0, 0), 2: (0, 1, 0), 3: (0, 0, 1), 4: (1, 1, 0), 5: (0, 1, 1), 6: (1, 0, 1), 7: (1, 1, 1)}


def essential_bcs(nx, ny, nz, ids, comps, vals, z0=0, nz_total=None):
    """Per-node essential-component mask (bit i = component i) and the prescribed velocity L-vector
    (src/BCData.cpp:27-117, src/system_driver.cpp:327-333)."""
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    mask = np.zeros(nn, dtype=np.uint8)
    val = np.zeros(3 * nn)
    vals = np.asarray(vals, dtype=float).reshape(-1, 3)
    for s, attr in enumerate(ids):
        nodes = face_nodes(nx, ny, nz, attr, z0, nz_total)
        for d in range(3):
            if _CMP[abs(comps[s])][d]:
                mask[nodes] |= (1 << d)
                val[d * nn + nodes] = vals[s, d]
    return mask, val


def vgrad_mask(nx, ny, nz, ids, comps, z0=0, nz_total=None):
    """Per-node mask of the components driven by a velocity-gradient BC: the attributes whose component code is
    negative (BCs.essential_comps, src/option_parser.cpp:179-194)."""
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    mask = np.zeros(nn, dtype=np.uint8)
    for s, attr in enumerate(ids):
        if comps[s] >= 0:
            continue
        nodes = face_nodes(nx, ny, nz, attr, z0, nz_total)
        for d in range(3):
            if _CMP[abs(comps[s])][d]:
                mask[nodes] |= (1 << d)
    return mask


def uniaxial_mask(nx, ny, nz):
    return essential_bcs(nx, ny, nz, [1, 2, 3, 4], [3, 1, 2, 3], np.zeros((4, 3)))[0]


def uniaxial_velocity(coords, nn, rate):
    """A smooth velocity field compatible with the uniaxial BCs (used to seed benchmarks)."""
    x, y, z = coords[:nn], coords[nn:2 * nn], coords[2 * nn:]
    return np.concatenate([-0.35 * rate * x, -0.35 * rate * y, rate * z])


def voronoi_grains(nx, ny, nz, ngrains, seed):
    """Voronoi tessellation of `ngrains` uniformly random seeds evaluated at voxel centroids, ids 1..G,
    x fastest like grains.txt."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    seeds = rng.random((ngrains, 3))
    k, j, i = np.meshgrid((np.arange(nz) + 0.5) / nz, (np.arange(ny) + 0.5) / ny, (np.arange(nx) + 0.5) / nx,
                          indexing="ij")
    pts = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    _, idx = cKDTree(seeds).query(pts)
    return (idx + 1).astype(np.int32)


def random_quats(ngrains, seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(ngrains, 4))
    return q / np.linalg.norm(q, axis=1)[:, None]


def slab_partition(nz, nranks):
    """Contiguous z-slabs: rank r owns element layers [z0[r], z0[r+1])."""
    base, rem = divmod(nz, nranks)
    sizes = [base + (1 if r < rem else 0) for r in range(nranks)]
    z0 = np.concatenate([[0], np.cumsum(sizes)])
    return z0

"""Synthetic ScanNet-/nuScenes-shaped point clouds.

CPU/numpy only; used by bench.py, the tests and the golden generator.  A scene is
a set of UNIQUE voxels (so serialization codes are unique and argsort stability
is irrelevant).

ScanNet-shaped: an 8 m x 6 m x 3 m room voxelised at 0.02 m (grid < 400x300x150
=> serialized depth 9).  Candidate voxels are the floor, the four walls and the
faces of 12 axis-aligned furniture boxes; a smooth pseudo-random "scan coverage"
field picks exactly N of them, so surfaces are locally dense (realistic
neighbour occupancy and grid-pool ratios) and the voxel count is exact.
"""
import numpy as np


def _faces_room(nx, ny, nz):
    xs, ys, zs = np.arange(nx), np.arange(ny), np.arange(nz)
    out = []
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    out.append((np.stack([X.ravel(), Y.ravel(), np.zeros(X.size, np.int64)], 1), (0, 0, 1)))
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    out.append((np.stack([X.ravel(), np.zeros(X.size, np.int64), Z.ravel()], 1), (0, 1, 0)))
    out.append((np.stack([X.ravel(), np.full(X.size, ny - 1), Z.ravel()], 1), (0, -1, 0)))
    Y, Z = np.meshgrid(ys, zs, indexing="ij")
    out.append((np.stack([np.zeros(Y.size, np.int64), Y.ravel(), Z.ravel()], 1), (1, 0, 0)))
    out.append((np.stack([np.full(Y.size, nx - 1), Y.ravel(), Z.ravel()], 1), (-1, 0, 0)))
    return out


def _faces_box(x0, y0, x1, y1, h):
    out = []
    X, Y = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), indexing="ij")
    out.append((np.stack([X.ravel(), Y.ravel(), np.full(X.size, h)], 1), (0, 0, 1)))
    for (xa, xb, ya, yb, nrm) in ((x0, x1, y0, y0 + 1, (0, -1, 0)), (x0, x1, y1 - 1, y1, (0, 1, 0)),
                                  (x0, x0 + 1, y0, y1, (-1, 0, 0)), (x1 - 1, x1, y0, y1, (1, 0, 0))):
        X, Y, Z = np.meshgrid(np.arange(xa, xb), np.arange(ya, yb), np.arange(1, h), indexing="ij")
        out.append((np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1), nrm))
    return out


def scannet_scene(n_points=120000, seed=0, room_m=(8.0, 6.0, 3.0), grid_size=0.02, n_boxes=12, roughness=0.5):
    """-> dict(coord f32 [N,3], grid_coord i32 [N,3], feat f32 [N,6]) with exactly n_points unique voxels."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = (int(round(r / grid_size)) for r in room_m)
    faces = _faces_room(nx, ny, nz)
    for _ in range(n_boxes):
        sx, sy = rng.uniform(0.4, 2.0, 2) / grid_size
        h = int(rng.uniform(0.4, min(2.0, room_m[2] * 0.9)) / grid_size)
        sx, sy = int(min(sx, nx - 4)), int(min(sy, ny - 4))
        x0 = int(rng.integers(1, max(2, nx - sx - 1))); y0 = int(rng.integers(1, max(2, ny - sy - 1)))
        faces += _faces_box(x0, y0, x0 + sx, y0 + sy, h)
    vox = np.concatenate([f[0] for f in faces]).astype(np.int64)
    nrm = np.concatenate([np.tile(np.asarray(f[1], np.float32), (len(f[0]), 1)) for f in faces])
    # sensor noise: a fraction of surface voxels sits one voxel off the ideal plane
    # (gives the ~2.3x first-level grid-pool ratio of real scans instead of 4x)
    off = (rng.uniform(0, 1, len(vox)) < roughness)[:, None] * np.rint(nrm).astype(np.int64)
    vox = vox + off
    ok = (vox >= 0).all(1) & (vox[:, 0] < nx) & (vox[:, 1] < ny) & (vox[:, 2] < nz)
    vox, nrm = vox[ok], nrm[ok]
    key = (vox[:, 0] * ny + vox[:, 1]) * nz + vox[:, 2]
    _, first = np.unique(key, return_index=True)
    vox, nrm = vox[first], nrm[first]
    if len(vox) < n_points:
        raise ValueError(f"room offers only {len(vox)} surface voxels < {n_points}")
    # smooth scan-coverage field: keep the n_points voxels with the largest value
    p = vox.astype(np.float64) * grid_size
    f = np.zeros(len(vox))
    for _ in range(6):
        k = rng.normal(0, 1.2, 3); ph = rng.uniform(0, 2 * np.pi)
        f += np.sin(p @ k + ph)
    f += 0.35 * rng.standard_normal(len(vox))
    keep = np.argsort(-f, kind="stable")[:n_points]
    keep = keep[rng.permutation(n_points)]          # points arrive in no particular order
    vox, nrm = vox[keep], nrm[keep]
    vox = vox - vox.min(0)
    coord = ((vox + rng.uniform(0.05, 0.95, vox.shape)) * grid_size).astype(np.float32)
    coord[:, :2] -= coord[:, :2].mean(0)             # CenterShift(apply_z=False)-like
    color = rng.uniform(-1, 1, (n_points, 3)).astype(np.float32)
    return dict(coord=coord, grid_coord=vox.astype(np.int32),
                feat=np.concatenate([color, nrm], 1).astype(np.float32))


def small_room(n_points=2000, seed=0):
    """BASELINE config 1: 2 m x 2 m x 1 m room, 2 000 unique voxels @ 0.02 m."""
    return scannet_scene(n_points, seed, room_m=(2.0, 2.0, 1.0), n_boxes=3)


def nuscenes_sweep(n_points=30000, seed=0, grid_size=0.05):
    """nuScenes-shaped sweep: 32-beam spinning LiDAR over a ground plane plus boxes,
    clipped to [-51.2,51.2]^2 x [-4,2.4], voxelised at 0.05 m (depth 11), 4-ch feat."""
    rng = np.random.default_rng(seed)
    pts = []
    elev = np.deg2rad(np.linspace(-30.0, 10.0, 32))
    n_az = 4096
    az = np.linspace(0, 2 * np.pi, n_az, endpoint=False)
    boxes = [(rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(1, 5), rng.uniform(1, 5)) for _ in range(40)]
    for e in elev:
        d = np.stack([np.cos(e) * np.cos(az), np.cos(e) * np.sin(az), np.full(n_az, np.sin(e))], 1)
        t = np.full(n_az, 70.0)
        if e < 0:
            t = np.minimum(t, 1.84 / -np.sin(e))     # sensor 1.84 m above the ground plane
        for (bx, by, sx, sy) in boxes:               # crude ray/box-footprint hit
            tx = (bx - np.sign(d[:, 0]) * sx / 2) / np.where(d[:, 0] == 0, 1e-9, d[:, 0])
            hit_y = np.abs(tx * d[:, 1] - by) < sy / 2
            hit_z = (tx * d[:, 2] + 1.84 > 0) & (tx * d[:, 2] + 1.84 < 2.0)
            t = np.where((tx > 1) & hit_y & hit_z, np.minimum(t, tx), t)
        p = d * t[:, None] + rng.normal(0, 0.01, (n_az, 3))
        pts.append(p[t < 69.0])
    p = np.concatenate(pts)
    ok = (np.abs(p[:, 0]) < 51.2) & (np.abs(p[:, 1]) < 51.2) & (p[:, 2] > -4) & (p[:, 2] < 2.4)
    p = p[ok]
    g = np.floor(p / grid_size).astype(np.int64)
    g -= g.min(0)
    key = (g[:, 0] * 4096 + g[:, 1]) * 4096 + g[:, 2]
    _, first = np.unique(key, return_index=True)
    first = first[rng.permutation(len(first))][:n_points]
    p, g = p[first], g[first]
    strength = rng.uniform(0, 1, (len(p), 1))
    return dict(coord=p.astype(np.float32), grid_coord=g.astype(np.int32),
                feat=np.concatenate([p, strength], 1).astype(np.float32))


def collate(scenes):
    """pointcept/datasets/utils.py:15-41 for our dict-of-arrays scenes: concat + cumulative offset."""
    out = {k: np.concatenate([s[k] for s in scenes]) for k in scenes[0]}
    out["offset"] = np.cumsum([len(s["coord"]) for s in scenes]).astype(np.int64)
    return out

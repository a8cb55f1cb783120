"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's test-time fragment pipeline.

  GridSample(mode="test")   pointcept/datasets/transform.py:796-933
  collate_fn                pointcept/datasets/utils.py:15-41
  vote accumulation         pointcept/engines/test.py:198-268

Pinned: tests/golden/fragments.npz holds the outputs of the reference's own transform.py / utils.py executed here
(tests/golden/make_golden_fragments.py, NumPy 2.3: the division by np.array(grid_size) runs in float64).  One
degree of freedom is NOT pinned by the reference: np.argsort's default kind is unstable, so the order of the points
inside a voxel -- hence which point of a voxel lands in fragment i -- is whatever numpy's introsort produces.  This
restatement (and the CUDA path) use a STABLE sort; the tests compare everything that is determined (hash keys, grid,
inverse, counts, the voxel sequence of every fragment, the per-voxel point sets over all fragments) bit-exactly.
"""
import numpy as np

FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)


def fnv_hash_vec(arr):
    """transform.py:918-933 (FNV64-1A as written there: multiply, then xor, per column)"""
    arr = arr.astype(np.uint64)
    h = FNV_OFFSET * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= FNV_PRIME
        h = np.bitwise_xor(h, arr[:, j])
    return h


def ravel_hash_vec(arr):
    """transform.py:907-916"""
    arr = arr - arr.min(0)
    arr = arr.astype(np.uint64)
    mx = arr.max(0).astype(np.uint64) + np.uint64(1)
    keys = np.zeros(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1] - 1):
        keys += arr[:, j]
        keys *= mx[j + 1]
    keys += arr[:, -1]
    return keys


def grid_sample_plan(coord, grid_size, hash_type="fnv", legacy_f32=False):
    """transform.py:825-836 (+ stable tie order).  -> dict(grid_coord, key, order, inverse, count, start)"""
    coord = np.asarray(coord)
    scaled = coord / np.float32(grid_size) if legacy_f32 else coord.astype(np.float64) / np.float64(grid_size)
    grid = np.floor(scaled).astype(np.int64)
    mn = grid.min(0)
    grid = grid - mn
    key = fnv_hash_vec(grid) if hash_type == "fnv" else ravel_hash_vec(grid)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    _, inv_sorted, count = np.unique(ks, return_inverse=True, return_counts=True)
    inverse = np.zeros_like(inv_sorted)
    inverse[order] = inv_sorted
    start = np.concatenate([[0], np.cumsum(count)])
    return dict(grid_coord=grid, key=key, order=order, inverse=inverse, count=count, start=start, min_grid=mn)


def fragment_index(plan):
    """transform.py:866-870: fragment i takes the (i % count)-th point of every voxel's run"""
    count, start, order = plan["count"], plan["start"], plan["order"]
    return np.stack([order[start[:-1] + i % count] for i in range(int(count.max()))])


def softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def vote(n_points, num_classes, fragments):
    """test.py:198-268: fragments = iterable of (index, logits) -> (pred float64 [n, C], labels)"""
    pred = np.zeros((n_points, num_classes), dtype=np.float64)
    for idx, logits in fragments:
        pred[idx] += softmax(logits.astype(np.float64))
    return pred, pred.argmax(1)

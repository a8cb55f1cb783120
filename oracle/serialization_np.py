"""Integer side of the oracle (numpy): space-filling-curve keys, argsort /
inverse, patch padding maps, grid-pool plan.  Bit-exact contracts.

TEST INFRASTRUCTURE -- see oracle/__init__.py.

Reference lines restated here:
  * Morton key            pointcept/models/utils/serialization/z_order.py:40-50, 66-101
  * Hilbert key           pointcept/models/utils/serialization/hilbert.py:91-198 (gray2binary 69-88)
  * encode(order, batch)  pointcept/models/utils/serialization/default.py:9-24
  * Point.serialization   pointcept/models/utils/structure.py:47-102
  * patch padding maps    pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py:188-244
  * grid-pool plan        point_transformer_v3m1_base.py:464-505
"""
import numpy as np

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")


def bit_length(v: int) -> int:
    return int(v).bit_length()


def morton3(x, y, z, depth):
    """z_order.py:40-50: x -> bit 3i+2, y -> 3i+1, z -> 3i+0."""
    x = x.astype(np.int64); y = y.astype(np.int64); z = z.astype(np.int64)
    key = np.zeros_like(x)
    for i in range(depth):
        key |= ((x >> i) & 1) << (3 * i + 2)
        key |= ((y >> i) & 1) << (3 * i + 1)
        key |= ((z >> i) & 1) << (3 * i + 0)
    return key


def hilbert3(x, y, z, depth):
    """hilbert.py:91-198 on integers (Skilling's transpose form).

    The reference unpacks each coordinate into `depth` bits (MSB first), walks
    bits MSB->LSB and dims 0..2; where the current bit of dim d is set it
    inverts the lower bits of dim 0, otherwise it exchanges the differing lower
    bits of dim 0 and dim d (hilbert.py:150-170).  The bits are then
    interleaved bit-major (x,y,z per level; hilbert.py:173) and Gray-decoded by
    a prefix XOR from the MSB (hilbert.py:69-88, 176).
    """
    X = [x.astype(np.int64).copy(), y.astype(np.int64).copy(), z.astype(np.int64).copy()]
    for b in range(depth):               # b = 0 is the MSB of the depth-bit window
        q = np.int64(1) << (depth - 1 - b)
        low = q - 1                      # mask of the bits below the current one
        for d in range(3):
            m = (X[d] & q) != 0
            # bit set: invert the lower bits of dim 0
            X[0] = np.where(m, X[0] ^ low, X[0])
            # bit clear: exchange differing lower bits of dim 0 and dim d
            t = np.where(m, 0, (X[0] ^ X[d]) & low)
            X[d] = X[d] ^ t
            X[0] = X[0] ^ t
    g = morton3(X[0], X[1], X[2], depth)  # bit-major interleave, x most significant per level
    s = 1
    while s < 64:                         # Gray -> binary == prefix XOR from the MSB
        g = g ^ (g >> s)
        s <<= 1
    return g


def encode(grid_coord, batch, depth, order):
    """default.py:9-24."""
    g = np.asarray(grid_coord)
    x, y, z = g[:, 0], g[:, 1], g[:, 2]
    if order == "z":
        code = morton3(x, y, z, depth)
    elif order == "z-trans":
        code = morton3(y, x, z, depth)
    elif order == "hilbert":
        code = hilbert3(x, y, z, depth)
    elif order == "hilbert-trans":
        code = hilbert3(y, x, z, depth)
    else:
        raise NotImplementedError(order)
    if batch is not None:
        code = (np.asarray(batch).astype(np.int64) << (depth * 3)) | code
    return code


def offset2batch(offset):
    """utils/misc.py:12-24."""
    offset = np.asarray(offset, dtype=np.int64)
    counts = np.diff(offset, prepend=0)
    return np.repeat(np.arange(len(offset), dtype=np.int64), counts)


def serialization(grid_coord, batch, orders=ORDERS, depth=None):
    """structure.py:47-102 without the shuffle (callers permute rows).

    Returns code[k,N], order[k,N], inverse[k,N] (int64) and depth."""
    g = np.asarray(grid_coord)
    if depth is None:
        depth = bit_length(int(g.max())) if g.size else 0
    assert depth <= 16
    code = np.stack([encode(g, batch, depth, o) for o in orders])
    order = np.argsort(code, axis=1, kind="stable").astype(np.int64)
    inverse = np.zeros_like(order)
    n = code.shape[1]
    for k in range(code.shape[0]):
        inverse[k, order[k]] = np.arange(n, dtype=np.int64)
    return code, order, inverse, depth


def patch_maps(offset, K):
    """point_transformer_v3m1_base.py:188-244 (flash branch semantics).

    pad[p]   : index into the scene-concatenated *sorted* sequence for padded slot p
    unpad[i] : padded slot of sorted position i
    cu_seqlens (int32): patch boundaries in padded slots."""
    offset = np.asarray(offset, dtype=np.int64)
    counts = np.diff(offset, prepend=0)
    padded = np.where(counts > K, (counts + K - 1) // K * K, counts)
    start = np.concatenate([[0], offset[:-1]]) if len(offset) else np.zeros(0, np.int64)
    pstart = np.concatenate([[0], np.cumsum(padded)[:-1]]) if len(offset) else np.zeros(0, np.int64)
    total = int(padded.sum())
    pad = np.zeros(total, np.int64)
    unpad = np.zeros(int(counts.sum()), np.int64)
    cu = []
    for b in range(len(offset)):
        n, npad, s, p = int(counts[b]), int(padded[b]), int(start[b]), int(pstart[b])
        unpad[s:s + n] = p + np.arange(n)
        j = np.arange(npad)
        # filler slots of the last patch replay the tail of the previous patch
        pad[p:p + npad] = s + np.where(j < n, j, j - K)
        cu.append(np.arange(p, p + npad, K, dtype=np.int32))
    cu.append(np.array([total], np.int32))
    return pad, unpad, np.concatenate(cu).astype(np.int32)


def pool_plan(code, stride, serialized_depth):
    """point_transformer_v3m1_base.py:464-505 (without the shuffle).

    `indices` / `head_indices` are only defined up to a within-cluster
    permutation in the reference (unstable torch.sort, :485-489); we return the
    stable choice.  Everything else is well defined."""
    pd = (int(np.ceil(stride)) - 1).bit_length()
    if pd > serialized_depth:
        pd = 0
    c = code >> (pd * 3)
    uniq, cluster, counts = np.unique(c[0], return_inverse=True, return_counts=True)
    indices = np.argsort(cluster, kind="stable").astype(np.int64)
    idx_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    head = indices[idx_ptr[:-1]]
    new_code = c[:, head]
    order = np.argsort(new_code, axis=1, kind="stable").astype(np.int64)
    inverse = np.zeros_like(order)
    for k in range(order.shape[0]):
        inverse[k, order[k]] = np.arange(order.shape[1], dtype=np.int64)
    return dict(pooling_depth=pd, cluster=cluster.astype(np.int64), counts=counts.astype(np.int64),
                indices=indices, idx_ptr=idx_ptr, head_indices=head, code=new_code,
                order=order, inverse=inverse, depth=serialized_depth - pd)


def fnv_hash_vec(arr):
    """pointcept/datasets/transform.py:918-933 (uint64 FNV, multiply-then-xor)."""
    arr = np.asarray(arr).copy().astype(np.uint64, copy=False)
    h = np.uint64(14695981039346656037) * np.ones(arr.shape[0], dtype=np.uint64)
    for j in range(arr.shape[1]):
        h *= np.uint64(1099511628211)
        h = np.bitwise_xor(h, arr[:, j])
    return h

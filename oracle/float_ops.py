"""numpy fp32 restatement of the float sparse-convolution semantics behind the reference's MinkowskiEngine layer
API (lib/minkowski_sparse_conv_layers.py) -- TEST INFRASTRUCTURE ONLY.

MinkowskiEngine (~0.5.4, README.md:49) is not vendored in the reference and cannot be installed offline, so this
restates its PUBLISHED behaviour as pinned by the reference's own call sites: absolute coordinates (multiples of
the tensor stride), HYPER_CUBE offsets with the first spatial axis fastest (`minkowski_expand_coord_2x`,
lib/minkowski_sparse_conv_layers.py:401-408; `unfold_kernel` of lossl_coord_me/model.py:335-337), odd kernels
centred, even kernels 0..k-1, stride-2 output coordinates floor(c / 2ts) * 2ts, kernel [K, C_in, C_out].
Parity of float kernels is by tolerance (1e-2 relative in fp16, BASELINE.json), not bit-exact.
"""
import numpy as np


def me_kernel_offsets(kernel_size):
    ks = (kernel_size,) * 3 if isinstance(kernel_size, int) else tuple(kernel_size)
    out = []
    for k in range(ks[0] * ks[1] * ks[2]):
        r, d = k, []
        for a in range(3):
            d.append(r % ks[a] - ((ks[a] - 1) // 2 if ks[a] % 2 == 1 else 0))
            r //= ks[a]
        out.append(d)
    return np.array(out, dtype=np.int64)


def _key(c):
    c = c.astype(np.int64)
    return (c[:, 0] << 54) | ((c[:, 1] + (1 << 16)) << 36) | ((c[:, 2] + (1 << 16)) << 18) | (c[:, 3] + (1 << 16))


def me_lookup(in_coords, out_coords, kernel_size, scale):
    """table[k, o] = input row + 1 of the voxel at out + offset(k) * scale, 0 if absent."""
    offs = me_kernel_offsets(kernel_size)
    ik = _key(in_coords)
    order = np.argsort(ik)
    sk = ik[order]
    table = np.zeros((offs.shape[0], out_coords.shape[0]), dtype=np.int32)
    for k, d in enumerate(offs):
        q = out_coords.astype(np.int64).copy()
        q[:, 1:] += d[None] * scale
        ok = (q[:, 1:] >= 0).all(1)
        qk = _key(np.where(ok[:, None], q, 0))
        pos = np.minimum(np.searchsorted(sk, qk), sk.shape[0] - 1)
        hit = ok & (sk[pos] == qk)
        table[k, hit] = order[pos[hit]] + 1
    return table


def sparse_conv_f32(feats, weight, table, bias=None):
    """out[o] = sum_k feats[table[k,o]-1] @ weight[k]   (weight [K, C_in, C_out]), fp32 accumulate."""
    out = np.zeros((table.shape[1], weight.shape[2]), dtype=np.float32)
    f = feats.astype(np.float32)
    for k in range(table.shape[0]):
        o = np.nonzero(table[k])[0]
        if o.size:
            out[o] += f[table[k, o] - 1] @ weight[k].astype(np.float32)
    if bias is not None:
        out += bias.astype(np.float32).reshape(1, -1)
    return out


def act(x, name, slope=0.0):
    if name in (None, 'None', 'none'):
        return x
    if name == 'relu':
        return np.maximum(x, 0)
    if name in ('leaky_relu', 'prelu'):
        return np.where(x < 0, x * slope, x)
    raise NotImplementedError(name)

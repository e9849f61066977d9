"""Training support for the float sparse convolution (first slice): a torch.autograd.Function around
`fpcc_spconv_f16`.

  forward   out[m]  = sum_k  A[nbr_k(m)] @ W[k] (+ bias)            fused tcgen05 kernel, fp32 accumulation
  dgrad     dA[j]   = sum_k  dOut[nbr_k^T(j)] @ W[k]^T               the SAME kernel on the transposed kernel map: the
                                                                    parameter layout [K, C_in, C_out] is exactly the
                                                                    [K, C_out', C_in'] operand the kernel wants
  wgrad     dW[k]   = A[in_k]^T @ dOut[out_k]                        `fpcc_spconv_wgrad_f16`: tcgen05 with BOTH operands
                                                                    MN-major (the contraction runs over the gathered
                                                                    rows), split-K tiles meeting in fp32 atomics; a
                                                                    per-offset library GEMM only above 256 channels
  dbias     = sum_m dOut[m]

MinkowskiEngine computes the same three products per kernel offset (ME backward = transposed gather-GEMM-scatter)."""
import torch

from . import ops

WGRAD_TC = True  # False: per-offset library GEMMs (A/B switch for the tests)


def transpose_table(table: torch.Tensor, n_in: int) -> torch.Tensor:
    """k-major map of the transposed convolution: tt[k, j] = m + 1 where table[k, m] == j + 1 (0 = none).  Each
    (offset, input row) pairs with at most one output row, so the scatter has no collisions."""
    kv, n_out = table.shape
    tt = torch.zeros((kv, n_in), dtype=torch.int32, device=table.device)
    k_idx, m_idx = torch.nonzero(table, as_tuple=True)
    tt[k_idx, (table[k_idx, m_idx] - 1).long()] = (m_idx + 1).to(torch.int32)
    return tt


# Tensors derived from a kernel map (transposed map for dgrad, compacted pair lists + tile offsets for wgrad) are the same
# for every layer and every step that uses the map: built once per table.  The entry keeps the table itself alive, so its
# address cannot be reused by another tensor while the entry exists; a handful of maps are live at a time.
_derived_cache = {}
_DERIVED_MAX = 64


def _derived(table: torch.Tensor, n_in: int) -> dict:
    key = (table.data_ptr(), tuple(table.shape), n_in, table.device.index)
    ent = _derived_cache.get(key)
    if ent is None:
        if len(_derived_cache) >= _DERIVED_MAX:
            _derived_cache.pop(next(iter(_derived_cache)))
        ent = {'table': table}
        _derived_cache[key] = ent
    return ent


def _transposed(table, n_in):
    ent = _derived(table, n_in)
    if 'tt' not in ent:
        ent['tt'] = transpose_table(table, n_in)
    return ent['tt']


def _compacted(table, n_in):
    ent = _derived(table, n_in)
    if 'pairs' not in ent:
        in_map, out_map, offsets = ops.kmap_compact(table)
        ent['pairs'] = (in_map, out_map, offsets.tolist())
    return ent['pairs']


class SparseConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias, table, compute_dtype):
        kv, c_in, c_out = weight.shape
        cin_p, cout_p = max(16, (c_in + 7) // 8 * 8), max(16, c_out)
        f = torch.nn.functional.pad(feats.to(compute_dtype), (0, cin_p - c_in)).contiguous()
        w_t = torch.zeros((kv, cout_p, cin_p), dtype=compute_dtype, device=weight.device)
        w_t[:, :c_out, :c_in] = weight.detach().permute(0, 2, 1).to(compute_dtype)
        b = None
        if bias is not None:
            b = torch.zeros(cout_p, dtype=torch.float32, device=weight.device)
            b[:c_out] = bias.detach().float().reshape(-1)
        out = ops.spconv_f16(f, w_t, table, bias=b, out_dtype=torch.float32)[:, :c_out]
        ctx.save_for_backward(feats, weight, table)
        ctx.has_bias, ctx.compute_dtype = bias is not None, compute_dtype
        return out.to(feats.dtype if feats.dtype.is_floating_point else torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        feats, weight, table = ctx.saved_tensors
        kv, c_in, c_out = weight.shape
        dt = ctx.compute_dtype
        g = grad_out.contiguous()
        d_feats = d_weight = d_bias = None
        if ctx.needs_input_grad[0]:
            cin_p, cout_p = max(16, c_in), max(16, (c_out + 7) // 8 * 8)
            gp = torch.nn.functional.pad(g.to(dt), (0, cout_p - c_out)).contiguous()
            w = torch.zeros((kv, cin_p, cout_p), dtype=dt, device=weight.device)   # [K, "C_out" = c_in, "C_in" = c_out]
            w[:, :c_in, :c_out] = weight.detach().to(dt)
            tt = _transposed(table, feats.shape[0])
            d_feats = ops.spconv_f16(gp, w, tt, out_dtype=torch.float32)[:, :c_in].to(feats.dtype)
        if ctx.needs_input_grad[1]:
            in_map, out_map, off = _compacted(table, feats.shape[0])
            if WGRAD_TC and ops.wgrad_supported(c_in, c_out):
                # tensor cores: contraction over the gathered rows, both operands MN-major (fpcc_spconv_wgrad_f16)
                d_weight = ops.spconv_wgrad_f16(feats.to(dt), g.to(dt), in_map, out_map, off, c_in, c_out)
            else:  # library GEMM per offset (wide layers)
                a32, g32 = feats.float(), g.float()
                d_weight = torch.zeros_like(weight, dtype=torch.float32)
                for k in range(kv):
                    s, e = off[k], off[k + 1]
                    if e > s:
                        d_weight[k] = a32[in_map[s:e].long()].t() @ g32[out_map[s:e].long()]
            d_weight = d_weight.to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_bias = g.float().sum(0)
        return d_feats, d_weight, d_bias, None, None


def sparse_conv(feats, weight, bias, table, compute_dtype=torch.float16):
    """differentiable sparse convolution: feats [n_in, C_in], weight [K, C_in, C_out], k-major table [K, n_out]"""
    return SparseConvFunction.apply(feats, weight, bias, table, compute_dtype)

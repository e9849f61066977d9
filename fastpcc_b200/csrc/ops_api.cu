// C-ABI entry points for the GEMM family and the element-wise requant epilogues.
#include <stdlib.h>

#include "common.cuh"

namespace fpcc {

int launch_pairs_simt(const PairArgs &a, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s);
int launch_conv_simt(const int8_t *feats, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                     int64_t ld, int n_out, const int32_t *zp_comp, const EpiParams &ep, void *out, cudaStream_t s);
// tensor-core (tcgen05) paths, igemm_tc.cu; return FPCC_ERR_UNSUPPORTED when the shape is not covered
int launch_conv_tc(const int8_t *feats, int n_in, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                   int64_t ld, int n_out, const int32_t *row_perm, const EpiParams &ep, void *out, cudaStream_t s);
int launch_pairs_tc(const PairArgs &a, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s);
bool tc_enabled();

// ---------------------------------------------------------------------------------------------
// element-wise requant / prelu (src/element_wise/*.cu).  One thread per 4 consecutive channels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) requant_kernel(const int32_t *__restrict__ in, int64_t total, int ch, EpiParams ep,
                                                      void *__restrict__ out) {
    const bool has_slope = ep.slope != nullptr;
    const int32_t slope = has_slope ? ep.slope[0] : 0;
    const int64_t zp = ep.zp[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % ch);
        int64_t o = epi_value(in[i], ep.bias ? ep.bias[c] : 0, has_slope, slope, ep.mul[ep.mul_is_scalar ? 0 : c], zp, ep.shift);
        epi_store(out, i, o, ep.out_type, nullptr, false, 0);
    }
}

// Four consecutive channels per thread (16-byte loads; 4/8/16-byte stores); needs ch % 4 == 0 and 16-byte aligned
// tensors.  This is the shape of every requant on the codec path (RequantFxpToScaledInt8 between layers).
template <int OUT>
__global__ void __launch_bounds__(256) requant_vec4_kernel(const int4 *__restrict__ in, int64_t total4, int ch4, EpiParams ep,
                                                           void *__restrict__ out) {
    const bool has_slope = ep.slope != nullptr;
    const int32_t slope = has_slope ? ep.slope[0] : 0;
    const int64_t zp = ep.zp[0];
    const uint32_t mul0 = ep.mul[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % ch4) * 4;
        const int4 v = __ldg(&in[i]);
        const int32_t a[4] = {v.x, v.y, v.z, v.w};
        int32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int64_t r = epi_value(a[q], ep.bias ? __ldg(&ep.bias[c + q]) : 0, has_slope, slope,
                                  ep.mul_is_scalar ? mul0 : __ldg(&ep.mul[c + q]), zp, ep.shift);
            if (OUT == FPCC_OUT_I8) o[q] = (int32_t)(r < -128 ? -128 : (r > 127 ? 127 : r));
            else if (OUT == FPCC_OUT_I16) o[q] = (int32_t)(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
            else o[q] = clamp_i32(r);
        }
        if (OUT == FPCC_OUT_I8)
            ((uint32_t *)out)[i] = (uint32_t)(o[0] & 0xff) | ((uint32_t)(o[1] & 0xff) << 8) | ((uint32_t)(o[2] & 0xff) << 16) | ((uint32_t)(o[3] & 0xff) << 24);
        else if (OUT == FPCC_OUT_I16)
            ((uint2 *)out)[i] = make_uint2((uint32_t)(o[0] & 0xffff) | ((uint32_t)o[1] << 16), (uint32_t)(o[2] & 0xffff) | ((uint32_t)o[3] << 16));
        else
            ((int4 *)out)[i] = make_int4(o[0], o[1], o[2], o[3]);
    }
}

// HI: shift in [32, 62] (every RequantFxpToScaledInt8 of a PTQ-converted model: 23 + requant_shift is ~48).  The
// quotient is then the high word of the 64-bit sum shifted by shift - 32: it always fits int32, so no clamp of the
// input is needed, and |x*mul| < 2^62, |zp| < 2^60, half <= 2^61 keep the sum inside int64.
template <bool SLOPE, bool HI>
__device__ __forceinline__ void requant_scalar_fast_loop(const int4 *__restrict__ in, int64_t total16, uint4 *__restrict__ out,
                                                         int32_t slope, int32_t B, int32_t thr, int32_t mul, int64_t c0, int shift,
                                                         int ch16, int64_t ld16) {
    const int64_t c_pos = c0, c_neg = c0 - 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (int64_t)gridDim.x * blockDim.x) {
        int4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldg(&in[4 * i + q]);
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int32_t a[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
            int32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int32_t x = a[e];
                if (SLOPE) {  // 0 <= slope <= 2^25: |result| <= |x|, the product is <= 0 for x < 0 so the -1 applies
                    int64_t p;
                    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(p) : "r"(x), "r"(slope), "l"((int64_t)((1 << 24) - 1)));
                    const int32_t pv = (int32_t)__funnelshift_r((uint32_t)p, (uint32_t)((uint64_t)p >> 32), 25);
                    x = x < 0 ? pv : x;
                }
                if (!HI) x = max(min(x, B), -B);
                int64_t t;
                asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(t) : "r"(x), "r"(mul), "l"(x < thr ? c_neg : c_pos));
                o[e] = HI ? ((int32_t)((uint64_t)t >> 32) >> (shift - 32))
                          : (int32_t)__funnelshift_r((uint32_t)t, (uint32_t)((uint64_t)t >> 32), shift);
            }
            uint32_t up;  // saturating pack: d = c[15:0] << 16 | sat8(a) << 8 | sat8(b)
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(up) : "r"(o[3]), "r"(o[2]), "r"(0));
            asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(w[q]) : "r"(o[1]), "r"(o[0]), "r"(up));
        }
        out[ld16 == ch16 ? i : (i / ch16) * ld16 + (i % ch16)] = make_uint4(w[0], w[1], w[2], w[3]);  // ld16: output row pitch / 16
    }
}

// RequantFxpToScaledInt8 between layers: ONE multiplier, no bias, no PReLU, int8 out.  16 elements per thread per
// step (four 16-byte loads in flight, one 16-byte store) and, when the ranges allow it, exact 32-bit arithmetic:
// v is first clamped to [-B, B] with B the smallest magnitude that already saturates int8, so v*mul + zp fits a
// 64-bit accumulator whose shifted value fits 32 bits; "t < 0" is decided as v < thr = ceil(-zp / mul).
__global__ void __launch_bounds__(256) requant_scalar_i8_kernel(const int4 *__restrict__ in, int64_t total16, EpiParams ep,
                                                                uint4 *__restrict__ out, int ch16, int64_t ld16) {
    const int64_t zp = ep.zp[0];
    const uint32_t mul = ep.mul[0];
    const int shift = ep.shift;
    const bool has_slope = ep.slope != nullptr;  // optional Q6.25 PReLU in front (PReLUIn32Out32 -> Requant chains)
    const int32_t slope = has_slope ? ep.slope[0] : 0;
    const int64_t half = shift > 0 ? (int64_t)1 << (shift - 1) : 0;
    const int64_t c0 = zp + half;
    const int64_t azp = zp < 0 ? -zp : zp;
    const bool hi = shift >= 32 && shift <= 62;
    // mul == 0: the sign of v*mul + zp does not depend on v, which "v < thr" cannot express for v == INT32_MAX
    bool fast = (hi || (shift <= 31 && (uint32_t)c0 != 0u)) && mul != 0u && mul < (1u << 31) && azp < ((int64_t)1 << 60) &&
                (!has_slope || (slope >= 0 && slope <= (1 << 25)));
    int32_t B = 0, thr = 0;
    if (fast) {
        int64_t b = 0;
        if (!hi) {
            const int64_t num = ((int64_t)129 << shift) + azp;
            b = mul ? (num + (int64_t)mul - 1) / (int64_t)mul : 0;
            fast = b <= 2147483646ll;  // inputs span all of int32 here: the clamp must not bind below saturation
            fast = fast && ((num + (int64_t)mul + azp + ((int64_t)1 << 31)) >> shift) < 2147483647ll;
        }
        int64_t t;
        if (mul == 0) t = zp < 0 ? 2147483647ll : -2147483648ll;
        else { const int64_t nz = -zp, m = (int64_t)mul; t = nz >= 0 ? (nz + m - 1) / m : -((-nz) / m); }
        if (hi && t > 2147483647ll) fast = false;  // no input clamp on this path: x == INT32_MAX must still compare right
        t = t > 2147483647ll ? 2147483647ll : (t < -2147483648ll ? -2147483648ll : t);
        B = (int32_t)b; thr = (int32_t)t;
    }
    // `fast` and `has_slope` are grid-uniform: one branch-free loop body per case, so that the 16 independent element
    // chains of a thread interleave (a per-element fast/slow branch serialises them)
    if (fast) {
        if (hi) {
            if (has_slope) requant_scalar_fast_loop<true, true>(in, total16, out, slope, B, thr, (int32_t)mul, c0, shift, ch16, ld16);
            else requant_scalar_fast_loop<false, true>(in, total16, out, slope, B, thr, (int32_t)mul, c0, shift, ch16, ld16);
        } else {
            if (has_slope) requant_scalar_fast_loop<true, false>(in, total16, out, slope, B, thr, (int32_t)mul, c0, shift, ch16, ld16);
            else requant_scalar_fast_loop<false, false>(in, total16, out, slope, B, thr, (int32_t)mul, c0, shift, ch16, ld16);
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total16; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w[4];
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            const int4 v = __ldg(&in[4 * i + q]);
            const int32_t a[4] = {v.x, v.y, v.z, v.w};
            uint32_t word = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int64_t r = epi_value(a[e], 0, has_slope, slope, mul, zp, shift);
                word |= (uint32_t)((int32_t)(r < -128 ? -128 : (r > 127 ? 127 : r)) & 0xff) << (8 * e);
            }
            w[q] = word;
        }
        out[ld16 == ch16 ? i : (i / ch16) * ld16 + (i % ch16)] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void __launch_bounds__(256) prelu_kernel(const int32_t *__restrict__ in, int64_t total, const int32_t *__restrict__ slope,
                                                    int32_t *__restrict__ out) {
    const int32_t sl = slope[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = clamp_i32(prelu_q25((int64_t)in[i], sl));
}

static int ew_grid(int64_t total) {
    int64_t b = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;  // 8 resident 256-thread blocks per SM, grid-stride beyond that
    return (int)(b < cap ? b : cap);
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_requant_ld(const int32_t *in, int64_t rows, int ch, const fpcc_epilogue *e, void *out, int64_t out_ld, void *stream) {
    FPCC_REQUIRE(in && out, "requant_ld: NULL pointer");
    FPCC_REQUIRE(rows > 0 && ch > 0, "requant_ld: rows and channels must be positive");
    int rc = check_epilogue(e, false);
    if (rc) return rc;
    FPCC_REQUIRE(e->out_type == FPCC_OUT_I8 && e->mul_is_scalar && !e->bias, "requant_ld: one multiplier, no bias, int8 output (RequantFxpToScaledInt8)");
    FPCC_REQUIRE(ch % 16 == 0 && out_ld % 16 == 0 && out_ld >= ch && (((uintptr_t)in | (uintptr_t)out) & 15) == 0,
                 "requant_ld: channels, output pitch and pointers must be multiples of 16");
    const int64_t total16 = rows * (ch / 16);
    requant_scalar_i8_kernel<<<ew_grid(total16), 256, 0, (cudaStream_t)stream>>>((const int4 *)in, total16, to_params(e), (uint4 *)out, ch / 16, out_ld / 16);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_requant(const int32_t *in, int64_t rows, int ch, const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(in && out, "requant: NULL pointer");
    FPCC_REQUIRE(rows > 0 && ch > 0, "requant: rows and channels must be positive");
    int rc = check_epilogue(e, false);
    if (rc) return rc;
    int64_t total = rows * ch;
    cudaStream_t s = (cudaStream_t)stream;
    EpiParams ep = to_params(e);
    if (e->out_type == FPCC_OUT_I8 && e->mul_is_scalar && !e->bias && total % 16 == 0 &&
        (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        requant_scalar_i8_kernel<<<ew_grid(total / 16), 256, 0, s>>>((const int4 *)in, total / 16, ep, (uint4 *)out, 1, 1);
    } else if (ch % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        int64_t t4 = total / 4;
        if (e->out_type == FPCC_OUT_I8) requant_vec4_kernel<FPCC_OUT_I8><<<ew_grid(t4), 256, 0, s>>>((const int4 *)in, t4, ch / 4, ep, out);
        else if (e->out_type == FPCC_OUT_I16) requant_vec4_kernel<FPCC_OUT_I16><<<ew_grid(t4), 256, 0, s>>>((const int4 *)in, t4, ch / 4, ep, out);
        else requant_vec4_kernel<FPCC_OUT_I32><<<ew_grid(t4), 256, 0, s>>>((const int4 *)in, t4, ch / 4, ep, out);
    } else {
        requant_kernel<<<ew_grid(total), 256, 0, s>>>(in, total, ch, ep, out);
    }
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_prelu_i32(const int32_t *in, int64_t numel, const int32_t *slope, int32_t *out, void *stream) {
    FPCC_REQUIRE(in && out && slope && numel > 0, "prelu: bad arguments");
    prelu_kernel<<<ew_grid(numel), 256, 0, (cudaStream_t)stream>>>(in, numel, slope, out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_gemm_i8(const int8_t *A, const int8_t *B, const int32_t *C, int c_mode, int32_t *D, int M, int N, int K,
                            void *stream) {
    FPCC_REQUIRE(A && B && D, "gemm_i8: NULL pointer");
    FPCC_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_i8: bad sizes");
    FPCC_REQUIRE(c_mode >= 0 && c_mode <= 2 && (c_mode == 0 || C), "gemm_i8: bad C operand");
    PairArgs a = {A, B, nullptr, nullptr, nullptr, 1, M, N, K, 1, C, c_mode, 0};
    EpiParams ep = {};
    return launch_pairs_simt(a, ep, D, ceil_div(M, 64), (cudaStream_t)stream);
}

extern "C" int fpcc_gather_gemm_scatter_i8(const int8_t *A, const int8_t *B, int32_t *D, const int32_t *gather_idx,
                                           const int32_t *scatter_idx, int L, int N, int K, void *stream) {
    FPCC_REQUIRE(A && B && D && gather_idx && scatter_idx, "gather_gemm_scatter_i8: NULL pointer");
    FPCC_REQUIRE(L > 0 && N > 0 && K > 0, "gather_gemm_scatter_i8: bad sizes");
    PairArgs a = {A, B, gather_idx, scatter_idx, nullptr, 1, L, N, K, 2, nullptr, 0, 0};
    EpiParams ep = {};
    return launch_pairs_simt(a, ep, D, ceil_div(L, 64), (cudaStream_t)stream);
}

extern "C" int fpcc_spconv_i8(const int8_t *in_feats, int n_in, int c_in, const int8_t *weight, int kvol, int c_out,
                              const int32_t *nbr_table, int64_t ld, int n_out, const int32_t *row_perm,
                              const int32_t *zp_comp, const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(in_feats && weight && nbr_table && out, "spconv_i8: NULL pointer");
    FPCC_REQUIRE(n_in > 0 && n_out > 0 && c_in > 0 && c_out > 0 && kvol > 0 && ld >= n_out, "spconv_i8: bad sizes");
    int rc = check_epilogue(e, true);
    if (rc) return rc;
    FPCC_REQUIRE(e->row_bias == nullptr, "spconv_i8: row_bias belongs to the linear kernels");
    EpiParams ep = to_params(e);
    if (tc_enabled() && !zp_comp) {
        rc = launch_conv_tc(in_feats, n_in, c_in, weight, kvol, c_out, nbr_table, ld, n_out, row_perm, ep, out, (cudaStream_t)stream);
        if (rc != FPCC_ERR_UNSUPPORTED) return rc;
    }
    FPCC_REQUIRE(!row_perm, "spconv_i8: a grouped (row_perm) table needs the tensor-core path (see fpcc_gemm_engine)");
    FPCC_REQUIRE(!e->post_requant_mul, "spconv_i8: the fused second stage needs the tensor-core path (see fpcc_gemm_engine)");
    FPCC_REQUIRE(e->out_ld == 0, "spconv_i8: an output pitch needs the tensor-core path (see fpcc_gemm_engine)");
    return launch_conv_simt(in_feats, c_in, weight, kvol, c_out, nbr_table, ld, n_out, zp_comp, ep, out, (cudaStream_t)stream);
}

extern "C" int fpcc_linear_i8(const int8_t *A, int m, int k, const int8_t *W, int n, const int32_t *sel_row,
                              const int32_t *sel_out, const int32_t *sel_offsets, int n_groups, int n_sel,
                              const fpcc_epilogue *e, void *out, void *stream) {
    FPCC_REQUIRE(A && W && out, "linear_i8: NULL pointer");
    FPCC_REQUIRE(m > 0 && n > 0 && k > 0, "linear_i8: bad sizes");
    int rc = check_epilogue(e, true);
    if (rc) return rc;
    EpiParams ep = to_params(e);
    PairArgs a;
    int max_tiles;
    if (sel_row) {
        FPCC_REQUIRE(sel_out && sel_offsets && n_groups > 0 && n_sel > 0, "linear_i8: incomplete selection");
        a = {A, W, sel_row, sel_out, sel_offsets, n_groups, n_sel, n, k, 0, nullptr, 0, 1};
        max_tiles = ceil_div(n_sel, 64) + n_groups;
    } else {
        a = {A, W, nullptr, nullptr, nullptr, 1, m, n, k, 0, nullptr, 0, 0};
        max_tiles = ceil_div(m, 64);
    }
    if (tc_enabled()) {
        rc = launch_pairs_tc(a, ep, out, max_tiles, (cudaStream_t)stream);
        if (rc != FPCC_ERR_UNSUPPORTED) return rc;
    }
    FPCC_REQUIRE(!e->post_requant_mul, "linear_i8: the fused second stage needs the tensor-core path (see fpcc_gemm_engine)");
    FPCC_REQUIRE(e->out_ld == 0, "linear_i8: an output pitch needs the tensor-core path (see fpcc_gemm_engine)");
    return launch_pairs_simt(a, ep, out, max_tiles, (cudaStream_t)stream);
}

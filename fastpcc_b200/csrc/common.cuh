// Shared helpers for the fastpcc_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fastpcc_b200.h"

namespace fpcc {

void set_error(const char *fmt, ...);

#define FPCC_REQUIRE(cond, ...)                     \
    do {                                            \
        if (!(cond)) {                              \
            ::fpcc::set_error(__VA_ARGS__);         \
            return FPCC_ERR_INVALID;                \
        }                                           \
    } while (0)

#define FPCC_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            ::fpcc::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FPCC_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define FPCC_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) {                                                          \
            ::fpcc::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return FPCC_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
int sm_count();

// ---- coordinate keys -------------------------------------------------------------------------
// packed = batch:10 | x:18 | y:18 | z:18 ; stored key = packed + 1 so that 0 stays "empty"
__device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
    return ((unsigned)x < (1u << 18)) & ((unsigned)y < (1u << 18)) & ((unsigned)z < (1u << 18)) & ((unsigned)b < 1023u);
}
__device__ __forceinline__ uint64_t pack_key(int b, int x, int y, int z) {
    return (((uint64_t)(unsigned)b << 54) | ((uint64_t)(unsigned)x << 36) | ((uint64_t)(unsigned)y << 18) | (uint64_t)(unsigned)z) + 1ull;
}
__device__ __forceinline__ uint32_t hash_slot(uint64_t key, uint32_t capacity) {
    key ^= key >> 33;
    key *= 0xff51afd7ed558ccdULL;
    key ^= key >> 33;
    key *= 0xc4ceb9fe1a85ec53ULL;
    key ^= key >> 33;
    return (uint32_t)(((key >> 32) * (uint64_t)capacity) >> 32);
}

// ---- integer epilogue arithmetic (bias_prelu_requant.cu:6-37, prelu.cu:6-21) ---------------------
// round-half-away-from-zero arithmetic shift: sign(v) * ((|v| + 2^(s-1)) >> s)  (requant.cu:16-20).
// For v < 0:  -((-v + h) >> s) = ceil((v - h) / 2^s) = (v + h - 1) >> s  with h = 2^(s-1), so one branch-free
// form serves both signs: (v + h - [v < 0]) >> s.
__device__ __forceinline__ int64_t rha_shift(int64_t v, int s) {
    if (s <= 0) return v;
    return (v + (((int64_t)1 << (s - 1)) - (int64_t)((uint64_t)v >> 63))) >> s;
}
__device__ __forceinline__ int64_t prelu_q25(int64_t v, int32_t slope) {
    return v < 0 ? rha_shift(v * (int64_t)slope, 25) : v;
}
__device__ __forceinline__ int32_t clamp_i32(int64_t v) {
    return (int32_t)(v < (int64_t)INT32_MIN ? (int64_t)INT32_MIN : (v > (int64_t)INT32_MAX ? (int64_t)INT32_MAX : v));
}

struct EpiParams {  // device-side copy of fpcc_epilogue with scalars resolved to pointers
    const int32_t *bias;
    const int32_t *slope;
    const uint32_t *mul;
    const int64_t *zp;
    int32_t shift;
    int32_t out_type;
    int32_t mul_is_scalar;
    const int32_t *residual;
    const int32_t *post_slope;
    const int32_t *row_bias;
    const uint8_t *row_idx;
    int32_t row_bias_bound;
    const uint32_t *post_mul;     // fused second stage (fpcc_epilogue::post_requant_mul ...), NULL = off
    const int64_t *post_zp;
    int32_t post_shift;
    const int32_t *post_slope2;
    int64_t out_ld;               // int8 outputs of the fused kernels: row pitch in elements (0 = dense rows of ch)
    int8_t *aux_out;              // dual output: int8 rows of the second stage beside the int32 rows (NULL = off)
};
static inline EpiParams to_params(const fpcc_epilogue *e) {
    EpiParams p;
    p.bias = e->bias; p.slope = e->slope; p.mul = e->requant_mul; p.zp = e->zero_point;
    p.shift = e->shift; p.out_type = e->out_type; p.mul_is_scalar = e->mul_is_scalar;
    p.residual = e->residual; p.post_slope = e->post_slope;
    p.row_bias = e->row_bias; p.row_idx = e->row_idx; p.row_bias_bound = e->row_bias_bound;
    p.post_mul = e->post_requant_mul; p.post_zp = e->post_zero_point; p.post_shift = e->post_shift; p.post_slope2 = e->post_requant_slope; p.aux_out = e->aux_out; p.out_ld = e->out_ld;
    return p;
}
int check_epilogue(const fpcc_epilogue *e, bool allow_residual);

// Applies the epilogue to one accumulator.  `slope_v`/`post_v`/`zp_v` are the pre-loaded scalars
// (has_slope/has_post tell whether they apply); returns the value before the final narrowing store.
__device__ __forceinline__ int64_t epi_value(int32_t acc, int32_t bias_v, bool has_slope, int32_t slope_v,
                                             uint32_t mul_v, int64_t zp_v, int shift) {
    int64_t v = (int64_t)acc + (int64_t)bias_v;
    if (has_slope) v = prelu_q25(v, slope_v);
    return rha_shift(v * (int64_t)mul_v + zp_v, shift);
}
__device__ __forceinline__ void epi_store(void *out, int64_t idx, int64_t o, int out_type, const int32_t *residual,
                                          bool has_post, int32_t post_v) {
    if (out_type == FPCC_OUT_I8) {
        ((int8_t *)out)[idx] = (int8_t)(o < -128 ? -128 : (o > 127 ? 127 : o));
    } else if (out_type == FPCC_OUT_I16) {
        ((int16_t *)out)[idx] = (int16_t)(o < -32768 ? -32768 : (o > 32767 ? 32767 : o));
    } else {
        int32_t r = clamp_i32(o);
        if (residual) {
            r = (int32_t)((uint32_t)r + (uint32_t)residual[idx]);  // torch int32 add wraps (cuda_ops.py:90)
            if (has_post) r = clamp_i32(prelu_q25((int64_t)r, post_v));
        }
        ((int32_t *)out)[idx] = r;
    }
}

// Pair-list GEMM arguments: D[out[i]] (+)= A[in[i]] * W[group]^T over pair ranges per weight group.
// raw: 0 -> fused epilogue store, 1 -> D = acc (+ C by c_mode), 2 -> D += acc
struct PairArgs {
    const int8_t *A;
    const int8_t *W;         // [n_groups, N, K]
    const int32_t *in_idx;   // per pair, or NULL (identity)
    const int32_t *out_idx;  // per pair, or NULL (identity)
    const int32_t *offsets;  // device [n_groups+1] pair range of every weight group, or NULL (one group of n_pairs)
    int n_groups, n_pairs, N, K;
    int raw;
    const int32_t *C;        // raw==1: bias (c_mode 1) or full matrix (c_mode 2)
    int c_mode;
    int bias_per_group;      // epilogue bias/mul indexed by group*N + col
};

}  // namespace fpcc

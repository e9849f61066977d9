// tcgen05 (5th-gen tensor core) int8 implicit GEMM for sm_100a: the hot kernel of the codec.
//
//   D[128 x N] (int32, TMEM)  +=  A[128 x 128B] (smem, gathered rows)  x  W[N x 128B]^T (smem, TMA)
//
// One CTA owns a tile of 128 output rows and up to 256 output channels; the accumulator lives in
// tensor memory for the whole tile, so a sparse convolution is OUTPUT-STATIONARY: all kernel offsets
// that have at least one neighbour in the tile are accumulated by `tcgen05.mma.kind::i8` before a
// single fused epilogue (bias + Q6.25 PReLU + requant [+ residual + PReLU]) writes int8 / Q8.23 once.
// No atomics, no int32 round trip through HBM, deterministic.
//
// Warp roles (192 threads):
//   warps 0-3  gather producers: one output row each; 16-byte cp.async with zero-fill for missing
//              neighbours, written straight into the 128B-swizzled K-major layout the MMA expects;
//              afterwards the same four warps run the epilogue (TMEM lane quarter = warp id).
//   warp 4     weight producer: one TMA box (N x 128 B, SWIZZLE_128B) per pipeline stage.
//   warp 5     TMEM allocation + the single MMA-issuing thread.
// smem stages hand over through mbarriers (full: 128 gather arrivals + TMA transaction bytes; empty:
// tcgen05.commit); the accumulator hands over to the epilogue through one more mbarrier.
//
// Two front-ends share the kernel: MODE_CONV reads the k-major neighbour table of coords.cu, MODE_PAIRS
// reads (in,out) pair lists grouped by weight block (dense linear, occupied-children linear).
//
// Replaces lib/int_sparse_conv/src/gather_gemm_scatter.cu:11-144 / gemm.cu:11-127 (CUTLASS 2.x
// mma.sync m16n8k32, one launch per kernel offset, read-modify-write of D per offset) plus the separate
// element-wise epilogue kernels of src/element_wise/*.cu.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace fpcc {

constexpr int TC_M = 128;         // rows per tile (UMMA M)
constexpr int TC_KB = 128;        // K bytes per stage (one 128B swizzle atom)
constexpr int TC_MAX_KVOL = 32;   // offsets tracked by the per-tile activity mask

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// SM100 shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO),
// the leading-dimension offset is unused for a single 128 B atom along K.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
    return d;
}
// kind::i8 instruction descriptor: D = s32, A = B = s8, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_i8(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
// kind::f16 instruction descriptor: D = f32, A = B = f16 (fmt 0) or bf16 (fmt 1), K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_f16(int n, int bf16) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
struct TcArgs {
    // common
    const int8_t *A;   // activations [*, K]
    int K;             // contraction bytes (C_in), multiple of 16
    int N;             // output channels of one weight block (C_out)
    int n_tile;        // channels per CTA: multiple of 16, <= 256
    int tmem_cols;     // power of two >= n_tile, >= 32
    // MODE_CONV
    const int32_t *nbr;
    int64_t ld;
    int n_out, kvol;
    // MODE_PAIRS
    const int32_t *in_idx, *out_idx, *offsets;
    int n_groups, n_pairs, bias_per_group;
    int dbg;  // FPCC_TC_DEBUG bitmask (experiments only): 1 skip A loads, 2 skip B loads, 4 skip epilogue math, 8 skip MMA
};

// ---------------------------------------------------------------------------------------------
// persistent kernel: one CTA per SM loops over tiles; the accumulator is double-buffered in TMEM so the
// epilogue of tile t (8 warps, int64 requant arithmetic) overlaps the gathers + MMAs of tile t+1; TMEM
// allocation, barrier setup and weight-descriptor prefetch are paid once per CTA instead of once per tile.
//
//   warps 0-7   epilogue: TMEM lane quarter = warp % 4, column half = warp / 4
//   warps 8-11  gather producers (one output row per thread) + per-tile metadata (source rows, offset mask)
//   warp 12     weight producer (TMA)          warp 13   TMEM owner + MMA issuer
// ---------------------------------------------------------------------------------------------
// Epilogue of one 32-column chunk of one output row.  Values stay in registers; the body is specialised at compile
// time on the output type, the PReLU and the occupancy row-bias so that the 32x unrolled arithmetic carries no
// per-element selects: ~22 (no PReLU) / ~32 (PReLU) integer instructions per element, all 64-bit exact.
struct EpiCtx {
    const int2 *chan;          // smem (bias, mul) pairs of this chunk
    int32_t slope, post;
    int64_t zp, half;          // half = 2^(shift-1) (0 when shift == 0)
    int shift, sgn;            // sgn = 1 when shift > 0 (round-half-away correction for negatives)
    const int32_t *row_bias;   // 32 entries of the occupancy-indexed bias row (ROWBIAS)
    const int32_t *residual;   // 32 entries, or NULL
    bool has_post;
    int nvalid;
};

__device__ __forceinline__ int32_t sat_s8(int64_t r) { int32_t o; asm("cvt.sat.s8.s64 %0, %1;" : "=r"(o) : "l"(r)); return o; }
__device__ __forceinline__ int32_t sat_s16(int64_t r) { int32_t o; asm("cvt.sat.s16.s64 %0, %1;" : "=r"(o) : "l"(r)); return o; }
__device__ __forceinline__ int32_t sat_s32(int64_t r) { int32_t o; asm("cvt.sat.s32.s64 %0, %1;" : "=r"(o) : "l"(r)); return o; }

template <int OUT, bool SLOPE, bool ROWBIAS>
__device__ __forceinline__ void epi_chunk(const uint32_t (&acc)[32], const EpiCtx &cx, void *optr, bool vec) {
    int32_t o[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        const int2 bm = cx.chan[q];
        int32_t a32 = (int32_t)acc[q];
        if (ROWBIAS) a32 = (int32_t)((uint32_t)a32 + (uint32_t)__ldg(&cx.row_bias[q < cx.nvalid ? q : 0]));
        int64_t v = (int64_t)a32 + (int64_t)bm.x;
        if (SLOPE) {  // Q6.25 PReLU on negatives, round half away (bias_prelu_requant.cu:17-22)
            int64_t t = v * (int64_t)cx.slope;
            t = (t + ((1ll << 24) - (int64_t)((uint64_t)t >> 63))) >> 25;
            v = v < 0 ? t : v;
        }
        int64_t r = v * (int64_t)(uint32_t)bm.y + cx.zp;
        r = (r + (cx.half - (int64_t)(((uint64_t)r >> 63) & (uint64_t)cx.sgn))) >> cx.shift;
        o[q] = OUT == FPCC_OUT_I8 ? sat_s8(r) : (OUT == FPCC_OUT_I16 ? sat_s16(r) : sat_s32(r));
    }
    if (OUT == FPCC_OUT_I32 && cx.residual) {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            int32_t rv = (int32_t)((uint32_t)o[q] + (uint32_t)__ldg(&cx.residual[q < cx.nvalid ? q : 0]));  // int32 add wraps
            o[q] = cx.has_post ? sat_s32(prelu_q25((int64_t)rv, cx.post)) : rv;
        }
    }
    if (vec) {
        if (OUT == FPCC_OUT_I8) {
            uint4 *dst = reinterpret_cast<uint4 *>(optr);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t w[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int q = h * 16 + t * 4;
                    w[t] = (uint32_t)(o[q] & 0xff) | ((uint32_t)(o[q + 1] & 0xff) << 8) | ((uint32_t)(o[q + 2] & 0xff) << 16) |
                           ((uint32_t)(o[q + 3] & 0xff) << 24);
                }
                dst[h] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            int4 *dst = reinterpret_cast<int4 *>(optr);
#pragma unroll
            for (int t = 0; t < 8; ++t) dst[t] = make_int4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            if (q < cx.nvalid) {
                if (OUT == FPCC_OUT_I8) ((int8_t *)optr)[q] = (int8_t)o[q];
                else if (OUT == FPCC_OUT_I16) ((int16_t *)optr)[q] = (int16_t)o[q];
                else ((int32_t *)optr)[q] = o[q];
            }
        }
    }
}

template <int OUT>
__device__ __forceinline__ void epi_dispatch(const uint32_t (&acc)[32], const EpiCtx &cx, void *optr, bool vec, bool slope, bool rb) {
    if (slope) { if (rb) epi_chunk<OUT, true, true>(acc, cx, optr, vec); else epi_chunk<OUT, true, false>(acc, cx, optr, vec); }
    else { if (rb) epi_chunk<OUT, false, true>(acc, cx, optr, vec); else epi_chunk<OUT, false, false>(acc, cx, optr, vec); }
}

// ---- floating-point epilogue (kind::f16 path): v = acc + bias; act; [+ residual; post act]; cast -------------
// act codes: 0 none, 1 relu, 2 leaky-relu / PReLU with one slope.  out_type: 0 fp16, 1 bf16, 2 fp32.
struct FEpi {
    const float *bias;      // [N] or NULL
    const void *residual;   // [rows, N] in the OUTPUT dtype, or NULL
    float slope, post_slope;
    int act, post_act, out_type;
};
__device__ __forceinline__ float f_act(float v, int act, float slope) {
    return act == 1 ? fmaxf(v, 0.f) : (act == 2 ? (v < 0.f ? v * slope : v) : v);
}
__device__ __forceinline__ float f_load(const void *p, int64_t i, int t) {
    return t == 0 ? __half2float(((const __half *)p)[i]) : (t == 1 ? __bfloat162float(((const __nv_bfloat16 *)p)[i]) : ((const float *)p)[i]);
}
__device__ __forceinline__ void epi_chunk_f(const uint32_t (&acc)[32], const int2 *chan, const FEpi &fe, const void *res,
                                            void *optr, int nvalid, bool vec) {
    float o[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        float v = __uint_as_float(acc[q]) + __int_as_float(chan[q].x);
        v = f_act(v, fe.act, fe.slope);
        if (res) v += f_load(res, q < nvalid ? q : 0, fe.out_type);
        o[q] = f_act(v, fe.post_act, fe.post_slope);
    }
    if (vec && fe.out_type != 2) {
        uint4 *dst = reinterpret_cast<uint4 *>(optr);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int q = h * 8 + t * 2;
                if (fe.out_type == 0) { __half2 hv = __floats2half2_rn(o[q], o[q + 1]); w[t] = *reinterpret_cast<uint32_t *>(&hv); }
                else { __nv_bfloat162 bv = __floats2bfloat162_rn(o[q], o[q + 1]); w[t] = *reinterpret_cast<uint32_t *>(&bv); }
            }
            dst[h] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    } else if (vec) {
        float4 *dst = reinterpret_cast<float4 *>(optr);
#pragma unroll
        for (int t = 0; t < 8; ++t) dst[t] = make_float4(o[4 * t], o[4 * t + 1], o[4 * t + 2], o[4 * t + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            if (q < nvalid) {
                if (fe.out_type == 0) ((__half *)optr)[q] = __float2half_rn(o[q]);
                else if (fe.out_type == 1) ((__nv_bfloat16 *)optr)[q] = __float2bfloat16_rn(o[q]);
                else ((float *)optr)[q] = o[q];
            }
        }
    }
}

constexpr int P_THREADS = 448;
constexpr int P_EPI_WARPS = 8;
constexpr int P_PROD_WARP0 = 8;

struct PMeta {  // per-tile metadata, double buffered
    uint32_t kmask;
    int32_t group, begin, end;
};

template <int STAGES>
struct PSmem {
    static size_t bytes(int n_tile, int rows_k) {
        return 1024 + (size_t)STAGES * (TC_M * TC_KB + (size_t)n_tile * TC_KB) + 2 * (size_t)rows_k * TC_M * 4 +
               (size_t)n_tile * 8 + 512;
    }
};

__device__ __forceinline__ void pairs_tile_lookup(const TcArgs &a, int tile_m, PMeta *m) {
    int g = 0, begin = tile_m * TC_M, end = a.n_pairs;
    if (a.offsets) {
        int t = tile_m;
        begin = end = -1;
        for (g = 0; g < a.n_groups; ++g) {
            int lo = a.offsets[g], hi = a.offsets[g + 1];
            int nt = (hi - lo + TC_M - 1) / TC_M;
            if (t < nt) { begin = lo + t * TC_M; end = hi; break; }
            t -= nt;
        }
    }
    m->group = g; m->begin = begin; m->end = end;
}

// KIND: 0 = int8 x int8 -> int32 (kind::i8, integer requant epilogue), 1 = fp16, 2 = bf16 (kind::f16, fp32
// accumulation, floating-point epilogue).  a.K is the contraction length in BYTES in every case.
template <int MODE, int STAGES, int KIND>
__global__ void __launch_bounds__(P_THREADS, 1) igemm_tc_persistent(TcArgs a, const __grid_constant__ CUtensorMap tmap_w,
                                                                    EpiParams ep, FEpi fe, void *__restrict__ out, int tiles_m,
                                                                    int tiles_n) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int b_bytes = a.n_tile * TC_KB;
    constexpr int a_bytes = TC_M * TC_KB;
    uint8_t *sA = smem;
    uint8_t *sB = smem + STAGES * a_bytes;
    const int rows_k = MODE == 0 ? a.kvol : 2;
    int32_t *rows_s = (int32_t *)(sB + (size_t)STAGES * b_bytes);  // [2 slots][rows_k][128]
    int2 *chan_s = (int2 *)(rows_s + 2 * rows_k * TC_M);          // [n_tile] (bias, mul) of the current channel block
    uint64_t *bars = (uint64_t *)(chan_s + a.n_tile);
    uint64_t *full = bars, *empty = bars + STAGES;
    uint64_t *meta_full = bars + 2 * STAGES, *tmem_full = meta_full + 2, *tmem_empty = tmem_full + 2;
    PMeta *meta = (PMeta *)(tmem_empty + 2);
    uint32_t *tmem_ptr = (uint32_t *)(meta + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = (a.K + TC_KB - 1) / TC_KB;
    const int total_tiles = tiles_m * tiles_n;

    if (warp == 13 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], TC_M + 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&meta_full[b], TC_M);
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], P_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    if (warp == 13) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)(2 * a.tmem_cols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // per-channel epilogue constants of a single channel block are staged once (grouped weights reload per tile)
    const bool chan_static = tiles_n == 1 && !(MODE == 1 && a.bias_per_group);
    if (chan_static)
        for (int c = tid; c < a.n_tile; c += P_THREADS) {
            if (KIND == 0) chan_s[c] = c < a.N ? make_int2(ep.bias ? ep.bias[c] : 0, (int)ep.mul[ep.mul_is_scalar ? 0 : c]) : make_int2(0, 0);
            else chan_s[c] = make_int2(c < a.N && fe.bias ? __float_as_int(fe.bias[c]) : 0, 0);
        }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp >= P_PROD_WARP0 && warp < P_PROD_WARP0 + 4) {
        // ================= metadata + gather producers =================
        const int r = tid - P_PROD_WARP0 * 32;
        const uint32_t row_smem = r * TC_KB, sw = r & 7;
        int it = 0;       // running pipeline step across tiles
        int arrived = 0;  // steps [0, arrived) have been published on their full barrier
        int j = 0;        // local tile counter
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
            const int slot = j & 1;
            const int tile_m = tile / tiles_n;
            int32_t *rows = rows_s + slot * rows_k * TC_M;
            mbar_wait(&tmem_empty[slot], ((j >> 1) & 1) ^ 1);  // slot's previous tile (j-2) fully consumed
            if (r == 0) meta[slot].kmask = 0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (MODE == 0) {
                const int m = tile_m * TC_M + r;
                uint32_t mine = 0;
                for (int k = 0; k < a.kvol; ++k) {
                    int32_t v = m < a.n_out ? __ldg(&a.nbr[(int64_t)k * a.ld + m]) : 0;
                    rows[k * TC_M + r] = v - 1;
                    mine |= (uint32_t)(v != 0) << k;
                }
                mine = __reduce_or_sync(0xffffffffu, mine);
                if (lane == 0 && mine) atomicOr(&meta[slot].kmask, mine);
            } else {
                if (r == 0) { pairs_tile_lookup(a, tile_m, &meta[slot]); }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int p = meta[slot].begin + r;
                const bool ok = meta[slot].begin >= 0 && p < meta[slot].end;
                rows[r] = ok ? (a.in_idx ? __ldg(&a.in_idx[p]) : p) : -1;
                rows[TC_M + r] = ok ? (a.out_idx ? __ldg(&a.out_idx[p]) : p) : -1;
                if (r == 0) meta[slot].kmask = meta[slot].begin >= 0 && meta[slot].begin < meta[slot].end ? 1u : 0u;
            }
            __threadfence_block();
            mbar_arrive(&meta_full[slot]);
            mbar_wait(&meta_full[slot], (j >> 1) & 1);
            uint32_t rem = meta[slot].kmask;
            const int total = __popc(rem) * n_chunks;
            int k = 0;
            for (int i = 0; i < total; ++i, ++it) {
                const int kc = i % n_chunks;
                if (kc == 0) { k = __ffs(rem) - 1; rem &= rem - 1; }
                const int stage = it % STAGES;
                mbar_wait(&empty[stage], ((it / STAGES) & 1) ^ 1);
                const int32_t src = rows[(MODE == 0 ? k : 0) * TC_M + r];
                const int8_t *gsrc = a.A + (src >= 0 ? (int64_t)src * a.K : 0) + kc * TC_KB;
                const uint32_t dst = smem_u32(sA + stage * a_bytes) + row_smem;
                if (!(a.dbg & 1)) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const bool ok = src >= 0 && kc * TC_KB + c * 16 < a.K;
                        cp_async16(dst + ((c ^ sw) << 4), ok ? (const void *)(gsrc + c * 16) : (const void *)a.A, ok ? 16u : 0u);
                    }
                }
                cp_async_commit();
                if (it - arrived >= STAGES - 1) {  // keep at most STAGES-1 steps in flight per thread
                    cp_async_wait<STAGES - 1>();
                    fence_proxy_async();
                    for (; arrived <= it - (STAGES - 1); ++arrived) mbar_arrive(&full[arrived % STAGES]);
                }
            }
            // Publish the tail of this tile before blocking on the next tile's accumulator: the MMAs of this
            // tile (and through them the epilogue that frees that accumulator) wait for these arrivals.
            cp_async_wait<0>();
            fence_proxy_async();
            for (; arrived < it; ++arrived) mbar_arrive(&full[arrived % STAGES]);
        }
    } else if (warp == 12) {
        // ================= weight producer (TMA) =================
        if (lane == 0) {
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
                const int slot = j & 1;
                const int n0 = (tile % tiles_n) * a.n_tile;
                mbar_wait(&meta_full[slot], (j >> 1) & 1);
                uint32_t rem = meta[slot].kmask;
                const int total = __popc(rem) * n_chunks;
                const int group = MODE == 1 ? meta[slot].group : 0;
                int k = 0;
                for (int i = 0; i < total; ++i, ++it) {
                    const int kc = i % n_chunks;
                    if (kc == 0) { k = __ffs(rem) - 1; rem &= rem - 1; }
                    const int stage = it % STAGES;
                    mbar_wait(&empty[stage], ((it / STAGES) & 1) ^ 1);
                    if (a.dbg & 2) { mbar_arrive(&full[stage]); continue; }
                    mbar_expect_tx(&full[stage], (uint32_t)b_bytes);
                    const int wrow = (MODE == 0 ? k : group) * a.N + n0;
                    tma_load_2d(smem_u32(sB + (size_t)stage * b_bytes), &tmap_w, &full[stage], kc * TC_KB, wrow);
                }
            }
        }
    } else if (warp == 13) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = KIND == 0 ? umma_idesc_i8(a.n_tile) : umma_idesc_f16(a.n_tile, KIND == 2);
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
                const int slot = j & 1;
                mbar_wait(&meta_full[slot], (j >> 1) & 1);
                const int total = __popc(meta[slot].kmask) * n_chunks;
                mbar_wait(&tmem_empty[slot], ((j >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(slot * a.tmem_cols);
                for (int i = 0; i < total; ++i, ++it) {
                    const int stage = it % STAGES;
                    mbar_wait(&full[stage], (it / STAGES) & 1);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * a_bytes));
                    const uint64_t bd = umma_desc_sw128(smem_u32(sB + (size_t)stage * b_bytes));
                    if (!(a.dbg & 8)) {
#pragma unroll
                        for (int q = 0; q < TC_KB / 32; ++q) {  // 32 bytes of K per instruction for both kinds
                            if (KIND == 0) umma_i8(tacc, ad + 2 * q, bd + 2 * q, idesc, (uint32_t)(i > 0 || q > 0));
                            else umma_f16(tacc, ad + 2 * q, bd + 2 * q, idesc, (uint32_t)(i > 0 || q > 0));
                        }
                    }
                    umma_commit(&empty[stage]);
                }
                if (total > 0) umma_commit(&tmem_full[slot]);
                else mbar_arrive(&tmem_full[slot]);
            }
        }
    } else if (warp < P_EPI_WARPS) {
        // ================= epilogue =================
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;  // tile row == TMEM lane
        const bool has_slope = KIND == 0 && ep.slope != nullptr, has_post = KIND == 0 && ep.post_slope != nullptr;
        const int32_t slope = has_slope ? ep.slope[0] : 0, post = has_post ? ep.post_slope[0] : 0;
        const int64_t zp = KIND == 0 ? ep.zp[0] : 0;
        const int shift = ep.shift;
        const int cols_half = ((a.n_tile / 2 + 31) / 32) * 32;  // columns per half, multiple of 32
        const int c_begin = half * cols_half, c_end = min(a.n_tile, c_begin + cols_half);
        int j = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
            const int slot = j & 1;
            const int tile_m = tile / tiles_n, n0 = (tile % tiles_n) * a.n_tile;
            mbar_wait(&meta_full[slot], (j >> 1) & 1);
            const bool have_acc = meta[slot].kmask != 0;
            const int pbase = (MODE == 1 && a.bias_per_group) ? meta[slot].group * a.N : 0;
            if (!chan_static) {  // grouped weights / several channel blocks: restage (bias, mul) of this tile's block
                asm volatile("bar.sync 2, 256;" ::: "memory");  // every epilogue warp is done with the previous tile's values
                const int t = warp * 32 + lane;
                if (t < a.n_tile) {
                    const int pc = pbase + min(n0 + t, a.N - 1);
                    if (KIND == 0) chan_s[t] = make_int2(ep.bias ? __ldg(&ep.bias[pc]) : 0, (int)__ldg(&ep.mul[ep.mul_is_scalar ? 0 : pc]));
                    else chan_s[t] = make_int2(fe.bias ? __float_as_int(__ldg(&fe.bias[pc])) : 0, 0);
                }
                asm volatile("bar.sync 2, 256;" ::: "memory");
            }
            const int32_t *rows = rows_s + slot * rows_k * TC_M;
            const int64_t m = MODE == 0 ? (int64_t)tile_m * TC_M + r : (int64_t)rows[TC_M + r];
            const bool row_ok = MODE == 0 ? (m < a.n_out) : (m >= 0);
            mbar_wait(&tmem_full[slot], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(slot * a.tmem_cols) + ((uint32_t)(quarter * 32) << 16);
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t acc[32];
                if (have_acc) {
                    tmem_ld32(tacc + (uint32_t)c0, acc);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; ++q) acc[q] = 0;
                }
                const int nb = n0 + c0;
                if (!row_ok || nb >= a.N) continue;
                if (a.dbg & 4) { if (c0 == 0) ((int32_t *)out)[m] = (int32_t)acc[0]; continue; }
                if (KIND != 0) {
                    const int esz = fe.out_type == 2 ? 4 : 2;
                    const int nvalid = min(32, a.N - nb);
                    const void *res = fe.residual ? (const char *)fe.residual + (m * a.N + nb) * esz : nullptr;
                    epi_chunk_f(acc, chan_s + c0, fe, res, (char *)out + (m * a.N + nb) * esz, nvalid,
                                nvalid == 32 && (a.N & 7) == 0);
                    continue;
                }
                EpiCtx cx;
                cx.chan = chan_s + c0;
                cx.slope = slope; cx.post = post; cx.zp = zp; cx.shift = shift;
                cx.half = shift > 0 ? (int64_t)1 << (shift - 1) : 0; cx.sgn = shift > 0;
                cx.row_bias = ep.row_bias ? ep.row_bias + (int64_t)__ldg(&ep.row_idx[m]) * a.N + nb : nullptr;
                cx.residual = ep.residual ? ep.residual + m * a.N + nb : nullptr;
                cx.has_post = has_post;
                cx.nvalid = min(32, a.N - nb);
                void *optr = (char *)out + (m * a.N + nb) * (ep.out_type == FPCC_OUT_I8 ? 1 : (ep.out_type == FPCC_OUT_I16 ? 2 : 4));
                const bool vec = cx.nvalid == 32 && (a.N & 15) == 0;
                const bool rb = ep.row_bias != nullptr;
                if (ep.out_type == FPCC_OUT_I8) epi_dispatch<FPCC_OUT_I8>(acc, cx, optr, vec, has_slope, rb);
                else if (ep.out_type == FPCC_OUT_I32) epi_dispatch<FPCC_OUT_I32>(acc, cx, optr, vec, has_slope, rb);
                else epi_dispatch<FPCC_OUT_I16>(acc, cx, optr, false, has_slope, rb);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[slot]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * a.tmem_cols)) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// tensor-pipe ceiling probe: back-to-back kind::i8 MMAs (M=128, N=256, K=32) on resident smem tiles, no
// loads in the loop.  Gives the measured int8 peak that the roofline of this kernel is quoted against.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) mma_i8_peak_kernel(int iters, int n) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (TC_M * TC_KB + 256 * TC_KB) / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) {
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = umma_idesc_i8(n);
        const uint64_t ad = umma_desc_sw128(smem_u32(smem));
        const uint64_t bd = umma_desc_sw128(smem_u32(smem + TC_M * TC_KB));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_i8(tmem_base, ad + 2 * j, bd + 2 * j, idesc, 1u);
        }
        umma_commit(&done);
    }
    mbar_wait(&done, 0);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_tc_mode = -1;  // -1: read FPCC_TC from the environment on first use
static int g_sm_budget = 0;  // persistent GEMM grids use at most this many SMs (0 = all)

bool tc_enabled() {
    if (g_tc_mode < 0) {
        const char *e = getenv("FPCC_TC");
        g_tc_mode = (e && e[0] == '0') ? 0 : 1;
    }
    return g_tc_mode == 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// weights are static per layer: one descriptor per (pointer, rows, K, box rows)
static int weight_tensor_map(const int8_t *w, int64_t rows, int K, int box_rows, CUtensorMap *out) {
    static std::mutex mu;
    static std::map<std::tuple<const void *, int64_t, int, int>, CUtensorMap> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple((const void *)w, rows, K, box_rows);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return FPCC_OK; }
    EncodeTiledFn fn = encode_fn();
    FPCC_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FPCC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows %lld, K %d, box %d)", (int)r, (long long)rows, K, box_rows);
    if (cache.size() > 4096) cache.clear();
    cache[key] = m;
    *out = m;
    return FPCC_OK;
}

static bool tc_shape_ok(const void *A, const void *W, int K, int N) {
    return K >= 32 && K % 16 == 0 && N >= 16 && (((uintptr_t)A | (uintptr_t)W) & 15) == 0;
}

static int pick_tile(int N, int *n_tile, int *tmem_cols) {
    int nt = N >= 256 ? 256 : ((N + 15) / 16) * 16;
    int cols = 32;
    while (cols < nt) cols <<= 1;
    *n_tile = nt;
    *tmem_cols = cols;
    return (N + nt - 1) / nt;
}

template <int MODE, int KIND>
static int launch_tc(TcArgs &a, const void *W, int64_t w_rows, int tiles_m, const EpiParams &ep, const FEpi &fe, void *out,
                     cudaStream_t s) {
    int n_blocks_n = pick_tile(a.N, &a.n_tile, &a.tmem_cols);
    { const char *e = getenv("FPCC_TC_DEBUG"); a.dbg = e ? atoi(e) : 0; }
    CUtensorMap tmap;
    int rc = weight_tensor_map((const int8_t *)W, w_rows, a.K, a.n_tile, &tmap);
    if (rc) return rc;
    constexpr int STAGES = 4;
    const int rows_k = MODE == 0 ? a.kvol : 2;
    size_t smem = PSmem<STAGES>::bytes(a.n_tile, rows_k);
    FPCC_REQUIRE(smem <= 227 * 1024, "igemm_tc: %zu bytes of shared memory exceed the 227 KB limit", smem);
    auto kern = igemm_tc_persistent<MODE, STAGES, KIND>;
    static bool configured = false;
    if (!configured) {
        FPCC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int total = tiles_m * n_blocks_n;
    int sms = g_sm_budget > 0 && g_sm_budget < sm_count() ? g_sm_budget : sm_count();
    int grid = total < sms ? total : sms;
    kern<<<grid, P_THREADS, smem, s>>>(a, tmap, ep, fe, out, tiles_m, n_blocks_n);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

int launch_conv_tc(const int8_t *feats, int n_in, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                   int64_t ld, int n_out, const EpiParams &ep, void *out, cudaStream_t s) {
    (void)n_in;
    if (!tc_shape_ok(feats, weight, c_in, c_out) || kvol > TC_MAX_KVOL) return FPCC_ERR_UNSUPPORTED;
    TcArgs a = {};
    a.A = feats; a.K = c_in; a.N = c_out;
    a.nbr = nbr; a.ld = ld; a.n_out = n_out; a.kvol = kvol;
    FEpi fe = {};
    return launch_tc<0, 0>(a, weight, (int64_t)kvol * c_out, ceil_div(n_out, TC_M), ep, fe, out, s);
}

int launch_pairs_tc(const PairArgs &p, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s) {
    (void)max_tiles;
    if (p.raw != 0 || !tc_shape_ok(p.A, p.W, p.K, p.N)) return FPCC_ERR_UNSUPPORTED;
    TcArgs a = {};
    a.A = p.A; a.K = p.K; a.N = p.N;
    a.in_idx = p.in_idx; a.out_idx = p.out_idx; a.offsets = p.offsets;
    a.n_groups = p.n_groups; a.n_pairs = p.n_pairs; a.bias_per_group = p.bias_per_group;
    int tiles = ceil_div(p.n_pairs, TC_M) + (p.offsets ? p.n_groups : 0);
    FEpi fe = {};
    return launch_tc<1, 0>(a, p.W, (int64_t)p.n_groups * p.N, tiles, ep, fe, out, s);
}

}  // namespace fpcc

// ---- fp16 / bf16 entry points (MinkowskiEngine / torchsparse float layers) ---------------------------------
static int f16_common(int dtype, int out_type, int act, int post_act, int c_in, int c_out, const void *A, const void *W, const void *out) {
    FPCC_REQUIRE(dtype == 0 || dtype == 1, "f16 GEMM: dtype must be 0 (fp16) or 1 (bf16)");
    FPCC_REQUIRE(out_type >= 0 && out_type <= 2, "f16 GEMM: out_type must be 0 (fp16), 1 (bf16) or 2 (fp32)");
    FPCC_REQUIRE(act >= 0 && act <= 2 && post_act >= 0 && post_act <= 2, "f16 GEMM: bad activation code");
    FPCC_REQUIRE(c_in >= 16 && c_in % 8 == 0 && c_out >= 16, "f16 GEMM: needs C_in %% 8 == 0, C_in >= 16, C_out >= 16 (got %d -> %d)", c_in, c_out);
    FPCC_REQUIRE((((uintptr_t)A | (uintptr_t)W | (uintptr_t)out) & 15) == 0, "f16 GEMM: pointers must be 16-byte aligned");
    return FPCC_OK;
}

extern "C" int fpcc_spconv_f16(const void *feats, int dtype, int n_in, int c_in, const void *weight, int kvol, int c_out,
                               const int32_t *nbr_table, int64_t ld, int n_out, const float *bias, int act, float slope,
                               const void *residual, int post_act, float post_slope, void *out, int out_type, void *stream) {
    using namespace fpcc;
    FPCC_REQUIRE(feats && weight && nbr_table && out, "spconv_f16: NULL pointer");
    FPCC_REQUIRE(n_in > 0 && n_out > 0 && kvol > 0 && kvol <= TC_MAX_KVOL && ld >= n_out, "spconv_f16: bad sizes");
    int rc = f16_common(dtype, out_type, act, post_act, c_in, c_out, feats, weight, out);
    if (rc) return rc;
    TcArgs a = {};
    a.A = (const int8_t *)feats; a.K = 2 * c_in; a.N = c_out;
    a.nbr = nbr_table; a.ld = ld; a.n_out = n_out; a.kvol = kvol;
    FEpi fe = {bias, residual, slope, post_slope, act, post_act, out_type};
    EpiParams ep = {};
    if (dtype == 0) return launch_tc<0, 1>(a, weight, (int64_t)kvol * c_out, ceil_div(n_out, TC_M), ep, fe, out, (cudaStream_t)stream);
    return launch_tc<0, 2>(a, weight, (int64_t)kvol * c_out, ceil_div(n_out, TC_M), ep, fe, out, (cudaStream_t)stream);
}

extern "C" int fpcc_linear_f16(const void *A, int dtype, int m, int k, const void *W, int n, const int32_t *sel_row,
                               const int32_t *sel_out, const int32_t *sel_offsets, int n_groups, int n_sel, const float *bias,
                               int act, float slope, const void *residual, int post_act, float post_slope, void *out,
                               int out_type, void *stream) {
    using namespace fpcc;
    FPCC_REQUIRE(A && W && out && m > 0, "linear_f16: bad arguments");
    int rc = f16_common(dtype, out_type, act, post_act, k, n, A, W, out);
    if (rc) return rc;
    TcArgs a = {};
    a.A = (const int8_t *)A; a.K = 2 * k; a.N = n;
    int tiles;
    if (sel_row) {
        FPCC_REQUIRE(sel_out && sel_offsets && n_groups > 0 && n_sel > 0, "linear_f16: incomplete selection");
        a.in_idx = sel_row; a.out_idx = sel_out; a.offsets = sel_offsets; a.n_groups = n_groups; a.n_pairs = n_sel; a.bias_per_group = 0;
        tiles = ceil_div(n_sel, TC_M) + n_groups;
    } else {
        a.n_groups = 1; a.n_pairs = m;
        tiles = ceil_div(m, TC_M);
        n_groups = 1;
    }
    FEpi fe = {bias, residual, slope, post_slope, act, post_act, out_type};
    EpiParams ep = {};
    if (dtype == 0) return launch_tc<1, 1>(a, W, (int64_t)n_groups * n, tiles, ep, fe, out, (cudaStream_t)stream);
    return launch_tc<1, 2>(a, W, (int64_t)n_groups * n, tiles, ep, fe, out, (cudaStream_t)stream);
}

extern "C" int fpcc_mma_i8_peak(int iters, int n, double *tops_out, void *stream) {
    using namespace fpcc;
    FPCC_REQUIRE(iters > 0 && tops_out && n >= 16 && n <= 256 && n % 16 == 0, "mma_i8_peak: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    size_t smem = 1024 + TC_M * TC_KB + 256 * TC_KB;
    FPCC_CUDA(cudaFuncSetAttribute(mma_i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    FPCC_CUDA(cudaEventCreate(&e0));
    FPCC_CUDA(cudaEventCreate(&e1));
    int sms = sm_count();
    mma_i8_peak_kernel<<<sms, 128, smem, s>>>(iters / 8 + 1, n);  // warm-up
    FPCC_CUDA(cudaEventRecord(e0, s));
    mma_i8_peak_kernel<<<sms, 128, smem, s>>>(iters, n);
    FPCC_CUDA(cudaEventRecord(e1, s));
    FPCC_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    FPCC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops_out = 2.0 * TC_M * n * 32.0 * 4.0 * iters * sms / (ms * 1e-3) / 1e12;
    return FPCC_OK;
}

extern "C" int fpcc_gemm_engine(int k, int n, int kvol, int has_zp_comp) {
    return fpcc::tc_enabled() && !has_zp_comp && k >= 32 && k % 16 == 0 && n >= 16 && kvol <= fpcc::TC_MAX_KVOL;
}

extern "C" int fpcc_set_sm_budget(int sms) {
    fpcc::g_sm_budget = sms > 0 ? sms : 0;
    return FPCC_OK;
}

extern "C" int fpcc_set_tc_mode(int mode) {
    fpcc::g_tc_mode = mode ? 1 : 0;
    return FPCC_OK;
}

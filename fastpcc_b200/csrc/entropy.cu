// Entropy head: integer softmax (LUT), 16-bit PMF/CDF quantisation and symbol-range extraction.
// One warp per row; lane-strided (coalesced) loads; warp shuffles for max / sum / prefix sums.
// HBM-bound: 4*S bytes read per row, 2*ld bytes (CDF) or 4 bytes (range) written.
//
// Replaces lib/int_sparse_conv/src/softmax.cu:41-144 and the torch passes of
// models/convolutional/lossl_coord_int/model.py:344-353 (which end in an n x 255 x 2 B device-to-host
// copy in the reference; here the table stays in HBM, or is never materialised on the encoder side).
#include "common.cuh"

namespace fpcc {

constexpr int EXP_LUT_SIZE = 12 * 512 + 1;
__device__ const int32_t g_exp_lut[EXP_LUT_SIZE] = {
#include "exp_lut.inc"
};

__device__ __forceinline__ int32_t warp_max(int32_t v) {
    return __reduce_max_sync(0xffffffffu, v);
}
__device__ __forceinline__ int32_t warp_sum(int32_t v) {
    return __reduce_add_sync(0xffffffffu, v);
}

// Row statistics shared by all three kernels.  `pre_shift` = 7 turns Q8.23 logits into the Q15.16
// softmax input (model.py:347), 0 takes Q15.16 directly (softmax_int32).
struct RowStat {
    int32_t row_max;  // max + 64 (softmax.cu:71)
    uint64_t inv;
};
__device__ __forceinline__ int32_t lut_of(int32_t row_max, int32_t x) {
    int32_t id = (row_max - x) >> 7;
    return __ldg(&g_exp_lut[id > EXP_LUT_SIZE - 1 ? EXP_LUT_SIZE - 1 : id]);
}
__device__ __forceinline__ RowStat row_stat(const int32_t *__restrict__ row, int S, int pre_shift, int lane) {
    int32_t m = INT32_MIN;
    for (int j = lane; j < S; j += 32) m = max(m, row[j] >> pre_shift);
    RowStat st;
    st.row_max = warp_max(m) + 64;
    int32_t sum = 0;
    for (int j = lane; j < S; j += 32) sum += lut_of(st.row_max, row[j] >> pre_shift);
    sum = warp_sum(sum);
    st.inv = sum > 0 ? ((1ull << 32) + (uint64_t)(sum >> 1)) / (uint64_t)sum : (1ull << 32) / (uint64_t)S;
    return st;
}
__device__ __forceinline__ uint32_t prob_q32(const RowStat &st, int32_t x) {
    uint64_t p = (uint64_t)lut_of(st.row_max, x) * st.inv;
    return p > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)p;
}

__global__ void __launch_bounds__(256) softmax_kernel(const int32_t *__restrict__ in, int64_t rows, int S, uint32_t *__restrict__ out) {
    int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int32_t *row = in + r * S;
        RowStat st = row_stat(row, S, 0, lane);
        for (int j = lane; j < S; j += 32) out[r * S + j] = prob_q32(st, row[j]);
    }
}

// MODE 0: write the inclusive uint16 CDF (pitch ld, pad entries = 0xFFFF);  MODE 1: write the packed
// range start | (freq-1)<<16 of symbols[r].
template <int MODE>
__global__ void __launch_bounds__(256) cdf_kernel(const int32_t *__restrict__ logits, int64_t logits_ld, int64_t rows, int S,
                                                  uint16_t *__restrict__ cdf, int ld, const int32_t *__restrict__ symbols,
                                                  uint32_t *__restrict__ ranges) {
    int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int32_t *row = logits + r * logits_ld;
        RowStat st = row_stat(row, S, 7, lane);
        uint32_t running = 0;
        int sym = MODE == 1 ? symbols[r] : 0;
        sym = sym < 0 ? 0 : (sym > S - 1 ? S - 1 : sym);
        uint32_t lo = 0, hi = 0;
        for (int j0 = 0; j0 < (MODE == 0 ? ld : S); j0 += 32) {
            int j = j0 + lane;
            uint32_t v = 0;
            if (j < S) v = (uint32_t)(((uint64_t)prob_q32(st, row[j] >> 7) * (uint64_t)(65536 - S)) >> 32) + 1u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += t;
            }
            uint32_t c = running + v;  // inclusive CDF at j
            running += __shfl_sync(0xffffffffu, v, 31);
            if (j == S - 1) c = 65535;  // model.py:351
            if (MODE == 0) {
                if (j < ld) cdf[r * ld + j] = j < S ? (uint16_t)c : (uint16_t)0xFFFF;
            } else {
                if (j == sym - 1) lo = c;
                if (j == sym) hi = (j == S - 1) ? 65536u : c;  // simple_rans_wrapper.cpp:89-90
            }
        }
        if (MODE == 1) {
            lo = __reduce_or_sync(0xffffffffu, lo);
            hi = __reduce_or_sync(0xffffffffu, hi);
            if (lane == 0) ranges[r] = lo | ((hi - lo - 1u) << 16);
        }
    }
}

// S <= 256: the whole row lives in registers.  Lane l owns the 8 CONTIGUOUS entries 8l .. 8l+7 (two 16-byte loads
// when the row pitch allows), looks every exponential up once, and the inclusive CDF is a 7-add local prefix plus
// ONE warp scan of the lane totals (the strided kernel above re-reads the row three times and scans 8 chunks).
// The normaliser's 64-bit division is a double division: num < 2^33 and sum < 2^31 are exact doubles and the
// quotient's rounding error (< 2^-20 / sum) is below the distance 1 / sum to the next integer, so the truncated
// result is the exact floor.
template <int MODE>
__global__ void __launch_bounds__(256) cdf_row_kernel(const int32_t *__restrict__ logits, int64_t logits_ld, int64_t rows, int S,
                                                      uint16_t *__restrict__ cdf, int ld, const int32_t *__restrict__ symbols,
                                                      uint32_t *__restrict__ ranges) {
    const int lane = threadIdx.x & 31;
    const bool vec_in = (logits_ld & 3) == 0 && ((uintptr_t)logits & 15) == 0 && 8 * lane + 8 <= logits_ld;
    const bool vec_out = MODE == 0 && ld == 256 && ((uintptr_t)cdf & 15) == 0;
    const uint32_t scale = (uint32_t)(65536 - S);
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const int32_t *row = logits + r * logits_ld;
        int32_t x[8];
        if (vec_in) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(row) + 2 * lane), b = __ldg(reinterpret_cast<const int4 *>(row) + 2 * lane + 1);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = 8 * lane + i < S ? __ldg(&row[8 * lane + i]) : 0;
        }
        int32_t m = INT32_MIN;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            x[i] >>= 7;  // Q8.23 -> Q15.16 (model.py:347)
            if (8 * lane + i < S) m = max(m, x[i]);
        }
        const int32_t row_max = warp_max(m) + 64;  // softmax.cu:71
        int32_t e[8];
        int32_t sum = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            e[i] = 8 * lane + i < S ? lut_of(row_max, x[i]) : 0;
            sum += e[i];
        }
        sum = warp_sum(sum);
        uint64_t inv;
        if (sum > 0) inv = (uint64_t)__double2ull_rz(__ddiv_rn((double)((1ull << 32) + (uint64_t)(sum >> 1)), (double)sum));
        else inv = (1ull << 32) / (uint64_t)S;
        uint32_t c[8];
        uint32_t run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint64_t p = (uint64_t)(uint32_t)e[i] * inv;
            const uint32_t p32 = p > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)p;
            run += 8 * lane + i < S ? __umulhi(p32, scale) + 1u : 0u;
            c[i] = run;
        }
        uint32_t incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t base = incl - run;  // CDF value just before this lane's first entry
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            c[i] += base;
            if (8 * lane + i == S - 1) c[i] = 65535u;  // model.py:351
        }
        if (MODE == 0) {
            if (vec_out) {
                uint32_t w[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const uint32_t lo = 8 * lane + 2 * t < S ? c[2 * t] : 0xFFFFu, hi = 8 * lane + 2 * t + 1 < S ? c[2 * t + 1] : 0xFFFFu;
                    w[t] = (lo & 0xFFFFu) | (hi << 16);
                }
                reinterpret_cast<uint4 *>(cdf + r * ld)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (8 * lane + i < ld) cdf[r * ld + 8 * lane + i] = 8 * lane + i < S ? (uint16_t)c[i] : (uint16_t)0xFFFF;
            }
        } else {
            int sym = symbols[r];
            sym = sym < 0 ? 0 : (sym > S - 1 ? S - 1 : sym);
            if ((sym >> 3) == lane) {
                const int si = sym & 7;
                uint32_t lo = base, hi = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i == si) hi = c[i];
                    if (i + 1 == si) lo = c[i];
                }
                if (sym == S - 1) hi = 65536u;  // simple_rans_wrapper.cpp:89-90
                ranges[r] = lo | ((hi - lo - 1u) << 16);
            }
        }
    }
}

__global__ void __launch_bounds__(256) table_ranges_kernel(const uint16_t *__restrict__ cdf, int64_t n_cdf, int S,
                                                           const int32_t *__restrict__ symbols, int64_t rows,
                                                           uint32_t *__restrict__ ranges) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint16_t *row = n_cdf == 1 ? cdf : cdf + r * S;
    int s = symbols[r];
    s = s < 0 ? 0 : (s > S - 1 ? S - 1 : s);
    uint32_t lo = s == 0 ? 0u : row[s - 1];
    uint32_t hi = s == S - 1 ? 65536u : row[s];
    ranges[r] = lo | ((hi - lo - 1u) << 16);
}

static int row_grid(int64_t rows) {
    int64_t b = (rows + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(b < cap ? b : cap);
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_softmax_i32(const int32_t *in, int64_t rows, int c, uint32_t *out, void *stream) {
    FPCC_REQUIRE(in && out, "softmax_i32: NULL pointer");
    FPCC_REQUIRE(rows > 0 && c > 1 && rows * c <= 0xFFFFFFFFll, "softmax_i32: need N > 0, C > 1, N*C < 2^32");
    softmax_kernel<<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(in, rows, c, out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_quantize_cdf(const int32_t *logits, int64_t logits_ld, int64_t rows, int s, uint16_t *cdf, int ld, void *stream) {
    FPCC_REQUIRE(logits && cdf, "quantize_cdf: NULL pointer");
    FPCC_REQUIRE(logits_ld >= s, "quantize_cdf: logits pitch %lld < s = %d", (long long)logits_ld, s);
    FPCC_REQUIRE(rows > 0 && s > 1 && s < 65536 && ld >= s, "quantize_cdf: bad sizes");
    if (s <= 256) cdf_row_kernel<0><<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(logits, logits_ld, rows, s, cdf, ld, nullptr, nullptr);
    else cdf_kernel<0><<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(logits, logits_ld, rows, s, cdf, ld, nullptr, nullptr);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_cdf_symbol_ranges(const int32_t *logits, int64_t logits_ld, int64_t rows, int s, const int32_t *symbols,
                                      uint32_t *ranges, void *stream) {
    FPCC_REQUIRE(logits && symbols && ranges, "cdf_symbol_ranges: NULL pointer");
    FPCC_REQUIRE(logits_ld >= s, "cdf_symbol_ranges: logits pitch %lld < s = %d", (long long)logits_ld, s);
    FPCC_REQUIRE(rows > 0 && s > 1 && s < 65536, "cdf_symbol_ranges: bad sizes");
    if (s <= 256) cdf_row_kernel<1><<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(logits, logits_ld, rows, s, nullptr, 0, symbols, ranges);
    else cdf_kernel<1><<<row_grid(rows), 256, 0, (cudaStream_t)stream>>>(logits, logits_ld, rows, s, nullptr, 0, symbols, ranges);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_table_symbol_ranges(const uint16_t *cdf, int64_t n_cdf, int s, const int32_t *symbols, int64_t rows,
                                        uint32_t *ranges, void *stream) {
    FPCC_REQUIRE(cdf && symbols && ranges, "table_symbol_ranges: NULL pointer");
    FPCC_REQUIRE(rows > 0 && s > 0 && (n_cdf == 1 || n_cdf == rows), "table_symbol_ranges: bad sizes");
    table_ranges_kernel<<<ceil_div(rows, 256), 256, 0, (cudaStream_t)stream>>>(cdf, n_cdf, s, symbols, rows, ranges);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

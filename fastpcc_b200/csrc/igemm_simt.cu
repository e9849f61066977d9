// CUDA-core (dp4a) int8 GEMM family.  This is the shape-generic path: it serves the layers whose
// contraction is too thin for the tensor pipe (C_in = 1 or 8: first conv, occupancy embeds -- pure
// bandwidth), arbitrary alignments of the drop-in API, and it is the on-device cross-check of the
// tcgen05 kernels in igemm_tc.cu.  Two kernels share one 64x64x64 tile engine:
//
//   conv_tile_kernel  : output-stationary sparse convolution from the k-major neighbour table
//                       (all kernel offsets accumulate in registers; fused epilogue; no atomics).
//   pairs_tile_kernel : pair-list GEMM  D[out[i]] (+)= A[in[i]] * W[k]^T  for dense GEMM, the mirror of
//                       cutlass_gather_gemm_scatter_int8, LinearIn8W8, and "occupied children only"
//                       Linear(C->8C) (pairs grouped by child slot, one weight block per group).
//
// Replaces lib/int_sparse_conv/src/gather_gemm_scatter.cu:11-144, gemm.cu:11-127 and the Python loop
// over kernel offsets at lib/int_sparse_conv/cuda_ops.py:153-166.
#include "common.cuh"

namespace fpcc {

constexpr int BM = 64, BN = 64, BK = 64;  // BK in bytes (= int8 elements)
constexpr int LDS_W = BK / 4 + 1;         // padded row pitch in 32-bit words

struct TileSmem {
    uint32_t a[BM][LDS_W];
    uint32_t b[BN][LDS_W];
};

// Copies bytes [k0, k0+64) of row `src` (row pitch K bytes; src < 0 -> zeros) into smem row `r`;
// thread handles 16-byte segment `seg`.  Falls back to byte loads at ragged / unaligned edges.
__device__ __forceinline__ void load_row_seg(uint32_t *dst_row, const int8_t *__restrict__ base, int64_t src, int K, int k0,
                                             int seg, bool vec_ok) {
    uint32_t w[4] = {0, 0, 0, 0};
    int kb = k0 + seg * 16;
    if (src >= 0 && kb < K) {
        const int8_t *p = base + src * (int64_t)K + kb;
        if (vec_ok && kb + 16 <= K) {
            int4 v = *reinterpret_cast<const int4 *>(p);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            int lim = K - kb < 16 ? K - kb : 16;
            for (int t = 0; t < lim; ++t) w[t >> 2] |= (uint32_t)(uint8_t)p[t] << ((t & 3) * 8);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) dst_row[seg * 4 + q] = w[q];
}

__device__ __forceinline__ void tile_mma(const TileSmem &s, int ty, int tx, int32_t (&acc)[4][4]) {
#pragma unroll
    for (int w = 0; w < BK / 4; ++w) {
        uint32_t a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = s.a[ty + 16 * i][w];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = s.b[tx + 16 * j][w];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a((int)a[i], (int)b[j], acc[i][j]);
    }
}

struct EpiScalars {
    bool has_slope, has_post;
    int32_t slope, post;
    int64_t zp;
};
__device__ __forceinline__ EpiScalars load_scalars(const EpiParams &e) {
    EpiScalars s;
    s.has_slope = e.slope != nullptr;
    s.has_post = e.post_slope != nullptr;
    s.slope = s.has_slope ? e.slope[0] : 0;
    s.post = s.has_post ? e.post_slope[0] : 0;
    s.zp = e.zp[0];
    return s;
}

// ---------------------------------------------------------------------------------------------
// output-stationary sparse convolution
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_tile_kernel(const int8_t *__restrict__ feats, int c_in,
                                                        const int8_t *__restrict__ weight, int kvol, int c_out,
                                                        const int32_t *__restrict__ nbr, int64_t ld, int n_out,
                                                        const int32_t *__restrict__ zp_comp, EpiParams ep,
                                                        void *__restrict__ out, bool vec_ok) {
    __shared__ TileSmem s;
    __shared__ int32_t src_row[BM];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    int32_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;

    for (int k = 0; k < kvol; ++k) {
        int any = 0;
        if (tid < BM) {
            int m = m0 + tid;
            int32_t v = m < n_out ? nbr[(int64_t)k * ld + m] : 0;
            src_row[tid] = v - 1;
            any = v != 0;
        }
        if (!__syncthreads_or(any)) continue;  // no output row of this tile has a neighbour at offset k
        if (zp_comp) {                          // cuda_ops.py:157-162: one compensation row per matched pair
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (src_row[ty + 16 * i] >= 0)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int n = n0 + tx + 16 * j;
                        if (n < c_out) acc[i][j] += zp_comp[k * c_out + n];
                    }
        }
        const int8_t *wk = weight + (int64_t)k * c_out * c_in;
        for (int k0 = 0; k0 < c_in; k0 += BK) {
            int r = tid >> 2, seg = tid & 3;
            load_row_seg(s.a[r], feats, src_row[r], c_in, k0, seg, vec_ok);
            int n = n0 + r;
            load_row_seg(s.b[r], wk, n < c_out ? n : -1, c_in, k0, seg, vec_ok);
            __syncthreads();
            tile_mma(s, ty, tx, acc);
            __syncthreads();
        }
    }
    EpiScalars sc = load_scalars(ep);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty + 16 * i;
        if (m >= n_out) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx + 16 * j;
            if (n >= c_out) continue;
            int64_t o = epi_value(acc[i][j], ep.bias ? ep.bias[n] : 0, sc.has_slope, sc.slope,
                                  ep.mul[ep.mul_is_scalar ? 0 : n], sc.zp, ep.shift);
            epi_store(out, (int64_t)m * c_out + n, o, ep.out_type, ep.residual, sc.has_post, sc.post);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pair-list GEMM
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) pairs_tile_kernel(PairArgs a, EpiParams ep, void *__restrict__ out, bool vec_ok) {
    __shared__ TileSmem s;
    __shared__ int32_t src_row[BM];
    __shared__ int32_t dst_row[BM];
    __shared__ int sh_group, sh_begin, sh_end;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int n0 = blockIdx.y * BN;
    if (tid == 0) {
        int g = 0, begin = blockIdx.x * BM, end = a.n_pairs;
        if (a.offsets) {  // find the (group, tile-in-group) this block owns
            int t = blockIdx.x;
            begin = end = -1;
            for (g = 0; g < a.n_groups; ++g) {
                int lo = a.offsets[g], hi = a.offsets[g + 1];
                int nt = (hi - lo + BM - 1) / BM;
                if (t < nt) { begin = lo + t * BM; end = hi; break; }
                t -= nt;
            }
        }
        sh_group = g; sh_begin = begin; sh_end = end;
    }
    __syncthreads();
    const int g = sh_group, begin = sh_begin, end = sh_end;
    if (begin < 0 || begin >= end) return;
    if (tid < BM) {
        int p = begin + tid;
        bool ok = p < end;
        src_row[tid] = ok ? (a.in_idx ? a.in_idx[p] : p) : -1;
        dst_row[tid] = ok ? (a.out_idx ? a.out_idx[p] : p) : -1;
    }
    __syncthreads();
    int32_t acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    const int8_t *wg = a.W + (int64_t)g * a.N * a.K;
    for (int k0 = 0; k0 < a.K; k0 += BK) {
        int r = tid >> 2, seg = tid & 3;
        load_row_seg(s.a[r], a.A, src_row[r], a.K, k0, seg, vec_ok);
        int n = n0 + r;
        load_row_seg(s.b[r], wg, n < a.N ? n : -1, a.K, k0, seg, vec_ok);
        __syncthreads();
        tile_mma(s, ty, tx, acc);
        __syncthreads();
    }
    EpiScalars sc;
    if (a.raw == 0) sc = load_scalars(ep);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = dst_row[ty + 16 * i];
        if (m < 0) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx + 16 * j;
            if (n >= a.N) continue;
            int64_t idx = (int64_t)m * a.N + n;
            if (a.raw == 1) {
                int32_t c = a.c_mode == 1 ? a.C[n] : (a.c_mode == 2 ? a.C[idx] : 0);
                ((int32_t *)out)[idx] = (int32_t)((uint32_t)acc[i][j] + (uint32_t)c);
            } else if (a.raw == 2) {
                ((int32_t *)out)[idx] = (int32_t)((uint32_t)((int32_t *)out)[idx] + (uint32_t)acc[i][j]);
            } else {
                int pc = a.bias_per_group ? g * a.N + n : n;
                if (ep.row_bias) acc[i][j] += ep.row_bias[(int64_t)ep.row_idx[m] * a.N + n];
                int64_t o = epi_value(acc[i][j], ep.bias ? ep.bias[pc] : 0, sc.has_slope, sc.slope,
                                      ep.mul[ep.mul_is_scalar ? 0 : pc], sc.zp, ep.shift);
                epi_store(out, idx, o, ep.out_type, ep.residual, sc.has_post, sc.post);
            }
        }
    }
}

static bool vec_ok_for(const void *a, const void *w, int K) {
    return (K % 16 == 0) && (((uintptr_t)a | (uintptr_t)w) & 15) == 0;
}

int launch_pairs_simt(const PairArgs &a, const EpiParams &ep, void *out, int max_tiles, cudaStream_t s) {
    dim3 grid(max_tiles, ceil_div(a.N, BN));
    pairs_tile_kernel<<<grid, 256, 0, s>>>(a, ep, out, vec_ok_for(a.A, a.W, a.K));
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

int launch_conv_simt(const int8_t *feats, int c_in, const int8_t *weight, int kvol, int c_out, const int32_t *nbr,
                     int64_t ld, int n_out, const int32_t *zp_comp, const EpiParams &ep, void *out, cudaStream_t s) {
    dim3 grid(ceil_div(n_out, BM), ceil_div(c_out, BN));
    conv_tile_kernel<<<grid, 256, 0, s>>>(feats, c_in, weight, kvol, c_out, nbr, ld, n_out, zp_comp, ep, out,
                                          vec_ok_for(feats, weight, c_in));
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

}  // namespace fpcc

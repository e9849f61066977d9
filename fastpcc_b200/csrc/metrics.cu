// Geometry distortion metric on the device: exact nearest neighbour of every query point in a reference
// cloud, the kernel under the D1 (point-to-point) / D2 (point-to-plane) PSNR that the reference obtains from the MPEG
// `pc_error` subprocess (lib/metrics/pc_error_wrapper.py:40-106, called from lib/evaluators.py:49-124).
// Brute force in integer arithmetic: exact squared distances, lowest reference index among equal distances.
// Bound: CUDA-core integer ALU (3 subtractions + 3 multiply-adds + compare/select per pair); reference points are
// staged once per CTA tile in shared memory and broadcast, so HBM/L2 traffic is nr*16 B per CTA.
#include "common.cuh"

namespace fpcc {

constexpr int NN_THREADS = 256;
constexpr int NN_QPT = 4;       // queries per thread (registers)
constexpr int NN_TILE = 1024;   // reference points per shared-memory tile (16 KB)

template <typename D>  // D = uint32_t when 3 * 4^coord_bits fits 32 bits, else uint64_t
__global__ void __launch_bounds__(NN_THREADS) nn_kernel(const int32_t *__restrict__ query, int nq, int q_ld, int q_col,
                                                        const int32_t *__restrict__ ref, int nr, int r_ld, int r_col,
                                                        int64_t *__restrict__ d2_out, int32_t *__restrict__ idx_out) {
    __shared__ int4 tile[NN_TILE];
    int qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], bi[NN_QPT];
    D best[NN_QPT];
    const int q0 = blockIdx.x * (NN_THREADS * NN_QPT) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j) {
        const int q = q0 + j * NN_THREADS;
        const int32_t *p = query + (int64_t)min(q, nq - 1) * q_ld + q_col;
        qx[j] = p[0]; qy[j] = p[1]; qz[j] = p[2];
        best[j] = ~(D)0; bi[j] = -1;
    }
    for (int base = 0; base < nr; base += NN_TILE) {
        const int cnt = min(NN_TILE, nr - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += NN_THREADS) {
            const int32_t *p = ref + (int64_t)(base + t) * r_ld + r_col;
            tile[t] = make_int4(p[0], p[1], p[2], 0);
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
            const int4 r = tile[t];
#pragma unroll
            for (int j = 0; j < NN_QPT; ++j) {
                const int dx = qx[j] - r.x, dy = qy[j] - r.y, dz = qz[j] - r.z;
                D d;
                if (sizeof(D) == 4) d = (D)((uint32_t)(dx * dx) + (uint32_t)(dy * dy) + (uint32_t)(dz * dz));
                else d = (D)((uint64_t)((int64_t)dx * dx) + (uint64_t)((int64_t)dy * dy) + (uint64_t)((int64_t)dz * dz));
                if (d < best[j]) { best[j] = d; bi[j] = base + t; }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NN_QPT; ++j) {
        const int q = q0 + j * NN_THREADS;
        if (q < nq) {
            d2_out[q] = (int64_t)best[j];
            if (idx_out) idx_out[q] = bi[j];
        }
    }
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_nn_search(const int32_t *query, int nq, int q_ld, int q_col, const int32_t *ref, int nr, int r_ld,
                              int r_col, int coord_bits, int64_t *d2_out, int32_t *idx_out, void *stream) {
    FPCC_REQUIRE(query && ref && d2_out, "nn_search: NULL pointer");
    FPCC_REQUIRE(nq > 0 && nr > 0, "nn_search: empty cloud (nq=%d nr=%d)", nq, nr);
    FPCC_REQUIRE(q_ld >= 3 + q_col && r_ld >= 3 + r_col && q_col >= 0 && r_col >= 0, "nn_search: bad row layout");
    FPCC_REQUIRE(coord_bits >= 1 && coord_bits <= 30, "nn_search: coord_bits must be in 1..30");
    const int grid = ceil_div(nq, NN_THREADS * NN_QPT);
    cudaStream_t s = (cudaStream_t)stream;
    if (coord_bits <= 14)  // |d| < 2^15 per axis with signed slack: 3 * 2^30 < 2^32
        nn_kernel<uint32_t><<<grid, NN_THREADS, 0, s>>>(query, nq, q_ld, q_col, ref, nr, r_ld, r_col, d2_out, idx_out);
    else
        nn_kernel<uint64_t><<<grid, NN_THREADS, 0, s>>>(query, nq, q_ld, q_col, ref, nr, r_ld, r_col, d2_out, idx_out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

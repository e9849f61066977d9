// Front end of the codec path: float points -> unique voxels in Morton order, and the kd-tree
// partition of large clouds.  Replaces the host NumPy of lib/datasets/KITTIOdometry/dataset.py:96-102,117-118
// (min, scale, round, np.unique, Morton argsort) and lib/data_utils.py:187-234 (_kd_tree_partition).
// All kernels are HBM-bound streaming passes; the sort is cub's radix sort restricted to the 3*coord_bits key bits.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fpcc {

// ---- float min (dataset.py:96 `xyz.min(0)`) ---------------------------------------------------------------
// order-preserving float -> uint32 map so that one atomicMin per block serves negative and positive values
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void fe_min_init_kernel(uint32_t *ord3) {
    if (threadIdx.x < 3) ord3[threadIdx.x] = 0xffffffffu;
}

__global__ void __launch_bounds__(256) fe_min_kernel(const float *__restrict__ pts, int64_t n, int ld, uint32_t *__restrict__ ord3) {
    uint32_t m[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = pts + i * ld;
#pragma unroll
        for (int a = 0; a < 3; ++a) m[a] = min(m[a], f2ord(__ldg(p + a)));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        m[a] = __reduce_min_sync(0xffffffffu, m[a]);
        if ((threadIdx.x & 31) == 0 && m[a] != 0xffffffffu) atomicMin(&ord3[a], m[a]);
    }
}

__global__ void fe_min_decode_kernel(const uint32_t *ord3, float *min3) {
    if (threadIdx.x < 3) min3[threadIdx.x] = ord2f(ord3[threadIdx.x]);
}

// ---- quantise + Morton key (dataset.py:97-100, space_filling_curves/__init__.py:46-88) ------------------------
__device__ __forceinline__ uint64_t fe_split3(uint32_t a) {
    uint64_t x = a & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ uint32_t fe_compact3(uint64_t x) {
    x &= 0x1249249249249249ull;
    x = (x ^ (x >> 2)) & 0x10c30c30c30c30c3ull;
    x = (x ^ (x >> 4)) & 0x100f00f00f00f00full;
    x = (x ^ (x >> 8)) & 0x1f0000ff0000ffull;
    x = (x ^ (x >> 16)) & 0x1f00000000ffffull;
    x = (x ^ (x >> 32)) & 0x1fffffull;
    return (uint32_t)x;
}

__global__ void __launch_bounds__(256) fe_key_kernel(const float *__restrict__ pts, int64_t n, int ld, const float *__restrict__ min3,
                                                     float scale, int coord_bits, int msb_axis, uint64_t *__restrict__ keys,
                                                     int32_t *__restrict__ overflow) {
    const float m0 = min3[0], m1 = min3[1], m2 = min3[2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float *p = pts + i * ld;
        // (p - min) * scale in float32, two roundings (numpy: `xyz -= org_point; xyz *= scale`), then rint (np.round)
        const float q0 = rintf(__fmul_rn(__fsub_rn(__ldg(p + 0), m0), scale));
        const float q1 = rintf(__fmul_rn(__fsub_rn(__ldg(p + 1), m1), scale));
        const float q2 = rintf(__fmul_rn(__fsub_rn(__ldg(p + 2), m2), scale));
        const float lim = (float)(1u << coord_bits);
        if (!(q0 >= 0.f && q0 < lim && q1 >= 0.f && q1 < lim && q2 >= 0.f && q2 < lim)) {  // also catches NaN
            *overflow = 1;
            keys[i] = 0;
            continue;
        }
        const uint64_t sx = fe_split3((uint32_t)q0), sy = fe_split3((uint32_t)q1), sz = fe_split3((uint32_t)q2);
        keys[i] = msb_axis == 0 ? ((sx << 2) | (sy << 1) | sz) : ((sz << 2) | (sy << 1) | sx);
    }
}

__global__ void __launch_bounds__(256) fe_flag_kernel(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void __launch_bounds__(256) fe_write_kernel(const uint64_t *__restrict__ keys, int64_t n, const int32_t *__restrict__ pos,
                                                       int msb_axis, int batch, int32_t *__restrict__ out_coords,
                                                       int32_t *__restrict__ n_out_dev, const int32_t *__restrict__ overflow) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pos[i];  // inclusive count of distinct keys up to i
    if (i == n - 1) *n_out_dev = *overflow ? -1 : p;
    if (i > 0 && pos[i - 1] == p) return;
    const uint64_t k = keys[i];
    const uint32_t hi = fe_compact3(k >> 2), mid = fe_compact3(k >> 1), lo = fe_compact3(k);
    reinterpret_cast<int4 *>(out_coords)[p - 1] =
        msb_axis == 0 ? make_int4(batch, (int)hi, (int)mid, (int)lo) : make_int4(batch, (int)lo, (int)mid, (int)hi);
}

// ---- kd-tree split (data_utils.py:196-205) -----------------------------------------------------------------
__global__ void __launch_bounds__(256) kd_moments_kernel(const int32_t *__restrict__ coords, int64_t n, unsigned long long *__restrict__ sums6) {
    unsigned long long s[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int4 c = reinterpret_cast<const int4 *>(coords)[i];
        const unsigned long long x = (unsigned)c.y, y = (unsigned)c.z, z = (unsigned)c.w;
        s[0] += x; s[1] += y; s[2] += z;
        s[3] += x * x; s[4] += y * y; s[5] += z * z;
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[a] += __shfl_xor_sync(0xffffffffu, s[a], o);
        if ((threadIdx.x & 31) == 0 && s[a]) atomicAdd(&sums6[a], s[a]);
    }
}

// axis = argmax of the variance, first maximum wins (np.argmax(np.var(coord, 0))).  n^2 * var = n*sum(x^2) - sum(x)^2
// is compared exactly in 128-bit integers (the reference compares float64 variances: equal unless two axes tie
// to ~1e-16 relative).
__global__ void kd_axis_kernel(const unsigned long long *sums6, int64_t n, int32_t *info) {
    if (threadIdx.x != 0) return;
    unsigned __int128 best = 0;
    int axis = 0;
    for (int a = 0; a < 3; ++a) {
        const unsigned __int128 v = (unsigned __int128)(unsigned long long)n * sums6[3 + a] - (unsigned __int128)sums6[a] * sums6[a];
        if (a == 0 || v > best) { best = v; axis = a; }
    }
    info[0] = axis;
}

__global__ void __launch_bounds__(256) kd_hist_kernel(const int32_t *__restrict__ coords, int64_t n, const int32_t *__restrict__ info,
                                                      int n_bins, int32_t *__restrict__ hist) {
    const int axis = info[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = coords[4 * i + 1 + axis];
        atomicAdd(&hist[min(max(v, 0), n_bins - 1)], 1);
    }
}

// split value = k-th smallest (1-based) of the column = the smallest v with cum[v] >= k  (torch.kthvalue)
__global__ void __launch_bounds__(256) kd_kth_kernel(const int32_t *__restrict__ cum, int n_bins, int64_t k, int32_t *__restrict__ info) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_bins) return;
    if ((int64_t)cum[v] >= k && (v == 0 || (int64_t)cum[v - 1] < k)) {
        info[1] = v;
        info[2] = cum[v];  // rows with value <= v : size of the left partition
    }
}

__global__ void __launch_bounds__(256) kd_flag_kernel(const int32_t *__restrict__ coords, int64_t n, const int32_t *__restrict__ info,
                                                      int32_t *__restrict__ flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = coords[4 * i + 1 + info[0]] <= info[1] ? 1 : 0;
}

// stable two-way partition: left rows keep their order in [0, n_left), right rows theirs in [n_left, n)
__global__ void __launch_bounds__(256) kd_scatter_kernel(const int32_t *__restrict__ coords, int64_t n, const int32_t *__restrict__ info,
                                                         const int32_t *__restrict__ pos, int32_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = reinterpret_cast<const int4 *>(coords)[i];
    const int incl = pos[i];
    const int a = info[0];
    const bool left = (a == 0 ? c.y : (a == 1 ? c.z : c.w)) <= info[1];
    const int64_t dst = left ? (int64_t)incl - 1 : (int64_t)info[2] + (i - incl);
    reinterpret_cast<int4 *>(out)[dst] = c;
}

static inline size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
static inline int stream_grid(int64_t n) {
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace fpcc

using namespace fpcc;

extern "C" size_t fpcc_voxelize_workspace(int64_t n) {
    if (n < 1) n = 1;
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DeviceRadixSort::SortKeys((void *)nullptr, sort_tmp, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)n, 0, 63);
    cub::DeviceScan::InclusiveSum((void *)nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)n);
    return 256 + 2 * align256((size_t)n * 8) + 2 * align256((size_t)n * 4) + align256(sort_tmp > scan_tmp ? sort_tmp : scan_tmp);
}

extern "C" int fpcc_voxelize_f32(const float *points, int64_t n, int ld, float scale, int coord_bits, int msb_axis, int batch,
                                 int32_t *out_coords, float *min_xyz_dev, int32_t *n_out_dev, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(points && out_coords && min_xyz_dev && n_out_dev && workspace, "voxelize_f32: NULL pointer");
    FPCC_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && ld >= 3, "voxelize_f32: bad sizes (n=%lld ld=%d)", (long long)n, ld);
    FPCC_REQUIRE(coord_bits >= 1 && coord_bits <= 21, "voxelize_f32: coord_bits must be in 1..21");
    FPCC_REQUIRE(msb_axis == 0 || msb_axis == 2, "voxelize_f32: msb_axis must be 0 or 2");
    FPCC_REQUIRE(workspace_bytes >= fpcc_voxelize_workspace(n), "voxelize_f32: workspace too small");
    FPCC_REQUIRE(((uintptr_t)out_coords & 15) == 0, "voxelize_f32: out_coords must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    char *w = (char *)workspace;
    uint32_t *ord3 = (uint32_t *)w;
    int32_t *overflow = (int32_t *)(w + 16);
    w += 256;
    uint64_t *keys = (uint64_t *)w; w += align256((size_t)n * 8);
    uint64_t *sorted = (uint64_t *)w; w += align256((size_t)n * 8);
    int32_t *flags = (int32_t *)w; w += align256((size_t)n * 4);
    int32_t *pos = (int32_t *)w; w += align256((size_t)n * 4);
    size_t tmp_bytes = workspace_bytes - (size_t)(w - (char *)workspace);
    const int g = stream_grid(n);
    fe_min_init_kernel<<<1, 32, 0, s>>>(ord3);
    FPCC_CUDA(cudaMemsetAsync(overflow, 0, 4, s));
    fe_min_kernel<<<g, 256, 0, s>>>(points, n, ld, ord3);
    fe_min_decode_kernel<<<1, 32, 0, s>>>(ord3, min_xyz_dev);
    fe_key_kernel<<<g, 256, 0, s>>>(points, n, ld, min_xyz_dev, scale, coord_bits, msb_axis, keys, overflow);
    FPCC_LAUNCH_CHECK();
    FPCC_CUDA(cub::DeviceRadixSort::SortKeys(w, tmp_bytes, keys, sorted, (int)n, 0, 3 * coord_bits, s));
    fe_flag_kernel<<<ceil_div(n, 256), 256, 0, s>>>(sorted, n, flags);
    FPCC_LAUNCH_CHECK();
    tmp_bytes = workspace_bytes - (size_t)(w - (char *)workspace);
    FPCC_CUDA(cub::DeviceScan::InclusiveSum(w, tmp_bytes, flags, pos, (int)n, s));
    fe_write_kernel<<<ceil_div(n, 256), 256, 0, s>>>(sorted, n, pos, msb_axis, batch, out_coords, n_out_dev, overflow);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" size_t fpcc_kd_split_workspace(int64_t n, int coord_bits) {
    if (n < 1) n = 1;
    const int64_t bins = (int64_t)1 << coord_bits;
    size_t scan_tmp = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, scan_tmp, (const int32_t *)nullptr, (int32_t *)nullptr, (int)(n > bins ? n : bins));
    return 256 + 2 * align256((size_t)bins * 4) + 2 * align256((size_t)n * 4) + align256(scan_tmp);
}

extern "C" int fpcc_kd_split(const int32_t *coords, int64_t n, int coord_bits, int32_t *out_coords, int32_t *info_dev,
                             void *workspace, size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(coords && out_coords && info_dev && workspace, "kd_split: NULL pointer");
    FPCC_REQUIRE(n >= 2 && n < ((int64_t)1 << 31), "kd_split: need 2 <= n < 2^31");
    FPCC_REQUIRE(coord_bits >= 1 && coord_bits <= 21, "kd_split: coord_bits must be in 1..21");
    FPCC_REQUIRE((((uintptr_t)coords | (uintptr_t)out_coords) & 15) == 0, "kd_split: coords must be 16-byte aligned");
    FPCC_REQUIRE(workspace_bytes >= fpcc_kd_split_workspace(n, coord_bits), "kd_split: workspace too small");
    // sum(x^2) must fit 64 bits: n * 4^coord_bits < 2^64
    FPCC_REQUIRE(2 * coord_bits + 31 <= 64 || n < ((int64_t)1 << (64 - 2 * coord_bits)), "kd_split: n * 4^coord_bits overflows");
    cudaStream_t s = (cudaStream_t)stream;
    const int bins = 1 << coord_bits;
    char *w = (char *)workspace;
    unsigned long long *sums6 = (unsigned long long *)w;
    w += 256;
    int32_t *hist = (int32_t *)w; w += align256((size_t)bins * 4);
    int32_t *cum = (int32_t *)w; w += align256((size_t)bins * 4);
    int32_t *flags = (int32_t *)w; w += align256((size_t)n * 4);
    int32_t *pos = (int32_t *)w; w += align256((size_t)n * 4);
    size_t tmp_bytes = workspace_bytes - (size_t)(w - (char *)workspace);
    const int g = stream_grid(n);
    FPCC_CUDA(cudaMemsetAsync(sums6, 0, 48, s));
    FPCC_CUDA(cudaMemsetAsync(hist, 0, (size_t)bins * 4, s));
    kd_moments_kernel<<<g, 256, 0, s>>>(coords, n, sums6);
    kd_axis_kernel<<<1, 32, 0, s>>>(sums6, n, info_dev);
    kd_hist_kernel<<<g, 256, 0, s>>>(coords, n, info_dev, bins, hist);
    FPCC_LAUNCH_CHECK();
    FPCC_CUDA(cub::DeviceScan::InclusiveSum(w, tmp_bytes, hist, cum, bins, s));
    kd_kth_kernel<<<ceil_div(bins, 256), 256, 0, s>>>(cum, bins, n / 2, info_dev);
    kd_flag_kernel<<<ceil_div(n, 256), 256, 0, s>>>(coords, n, info_dev, flags);
    FPCC_LAUNCH_CHECK();
    tmp_bytes = workspace_bytes - (size_t)(w - (char *)workspace);
    FPCC_CUDA(cub::DeviceScan::InclusiveSum(w, tmp_bytes, flags, pos, (int)n, s));
    kd_scatter_kernel<<<ceil_div(n, 256), 256, 0, s>>>(coords, n, info_dev, pos, out_coords);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

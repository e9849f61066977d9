// Coordinate hashing, kernel-map construction, stride-2 downsample / generative upsample coordinate
// generation and Morton codes.  All HBM/L2-latency-bound integer work: one thread per coordinate or
// per (coordinate, offset) probe, 16-byte coordinate loads, k-major tables so that every warp writes
// 128 contiguous bytes.
//
// Replaces lib/int_sparse_conv/src/hashmap/hashmap_cuda.cuh:171-191 (insert), :221-275 (lookup),
// lib/int_sparse_conv/cuda_ops.py:132-151 (compaction, host-synchronising in the reference),
// models/convolutional/lossl_coord_int/model.py:261-295 (get_bin) and :86-91 (child generation),
// lib/space_filling_curves/src/morton3d.cu:8-37.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fpcc {

// ---------------------------------------------------------------------------------------------
// hash table
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ int4 load_bxyz(const int32_t *coords, int64_t i, int layout) {
    int4 c = reinterpret_cast<const int4 *>(coords)[i];
    // layout 0: (b,x,y,z) ; layout 1: (x,y,z,b)  -> return as (b,x,y,z) in (x,y,z,w) slots
    return layout == 0 ? c : make_int4(c.w, c.x, c.y, c.z);
}

__global__ void __launch_bounds__(256) hash_insert_kernel(unsigned long long *keys, int32_t *vals, uint32_t capacity,
                                                          const int32_t *coords, int n, int layout) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = load_bxyz(coords, i, layout);
    if (!coord_in_range(c.x, c.y, c.z, c.w)) return;  // outside the key range: never matched by a lookup either
    unsigned long long key = pack_key(c.x, c.y, c.z, c.w);
    uint32_t slot = hash_slot(key, capacity);
    while (true) {
        unsigned long long prev = atomicCAS(&keys[slot], 0ull, key);
        if (prev == 0ull || prev == key) {
            vals[slot] = i + 1;
            return;
        }
        slot = slot + 1 == capacity ? 0 : slot + 1;
    }
}

__device__ __forceinline__ int32_t hash_find(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ vals,
                                             uint32_t capacity, unsigned long long key) {
    uint32_t slot = hash_slot(key, capacity);
    while (true) {
        unsigned long long cur = __ldg(&keys[slot]);
        if (cur == key) return __ldg(&vals[slot]);
        if (cur == 0ull) return 0;
        slot = slot + 1 == capacity ? 0 : slot + 1;
    }
}

struct KernelGeom {
    int ks[3];
    int st[3];
    int kvol;
    int me;  // 1: MinkowskiEngine convention (absolute coordinates, x fastest, even kernels 0..k-1, offsets scaled by st)
};

// offset of kernel index k along each axis, in the reference's enumeration order
// (hashmap_cuda.cuh:239-258): odd volume -> x fastest, even volume -> z fastest.
__device__ __forceinline__ void kernel_offset(const KernelGeom &g, int k, int &dx, int &dy, int &dz) {
    int d[3];
    if (g.me) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            d[a] = (k % g.ks[a] - ((g.ks[a] & 1) ? (g.ks[a] - 1) / 2 : 0)) * g.st[a];
            k /= g.ks[a];
        }
    } else if (g.kvol & 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            d[a] = k % g.ks[a] - (g.ks[a] - 1) / 2;
            k /= g.ks[a];
        }
    } else {
#pragma unroll
        for (int a = 2; a >= 0; --a) {
            d[a] = k % g.ks[a] - (g.ks[a] - 1) / 2;
            k /= g.ks[a];
        }
    }
    dx = d[0]; dy = d[1]; dz = d[2];
}

// grid = (ceil(n_out/256), kvol): each block probes 256 consecutive outputs for ONE offset, so the
// offset arithmetic is block-uniform and the k-major store is a contiguous 1 KB line per block.
__global__ void __launch_bounds__(256) kmap_lookup_kernel(const unsigned long long *__restrict__ keys,
                                                          const int32_t *__restrict__ vals, uint32_t capacity,
                                                          const int32_t *__restrict__ out_coords, int n_out, int layout,
                                                          KernelGeom g, int32_t *__restrict__ table, int k_major, int64_t ld) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (o >= n_out) return;
    int dx, dy, dz;
    kernel_offset(g, k, dx, dy, dz);
    int4 c = load_bxyz(out_coords, o, layout);
    int x, y, z;
    if (g.me) { x = c.y + dx; y = c.z + dy; z = c.w + dz; }
    else { x = c.y * g.st[0] + dx; y = c.z * g.st[1] + dy; z = c.w * g.st[2] + dz; }
    int32_t v = 0;
    if (coord_in_range(c.x, x, y, z)) v = hash_find(keys, vals, capacity, pack_key(c.x, x, y, z));
    if (k_major) table[(int64_t)k * ld + o] = v;
    else table[(int64_t)o * g.kvol + k] = v;
}

// 3x3x3 kernel map of a FINER pyramid level derived from the map of its parent level (octree neighbour finding), no
// hashing: the neighbour of child (parent j, slot s) at offset d lives in the parent-level neighbour
// D = floor((s + d) / 2) of j, at slot (s + d) mod 2, if that child exists; its row is the first-child row of that
// parent plus the number of its occupied slots below.  Reads: one entry of the (8x smaller, cache-resident) parent
// table, one occupancy byte, one child base -- instead of a random 12-byte probe of a hash table of all fine nodes.
// Slots are 4x + 2y + z with children stored in slot order (Morton, x most significant); occupancy bit 7 - slot.
// One thread per fine node for all 27 offsets: per axis the target s + d (d = -1,0,1) falls into one of only two
// parent-level offsets (D0 = s - 1, D0 + 1), so the node needs 8 parent-table entries (+ occupancy byte and child base
// of each), not 27; parent / slot are read once; the k-major stores of consecutive threads stay coalesced.
__global__ void __launch_bounds__(256) kmap_from_parent_kernel(const int32_t *__restrict__ ctable, int64_t ldc,
                                                               const uint8_t *__restrict__ cocc, const int32_t *__restrict__ cbase,
                                                               const int32_t *__restrict__ parent, const uint8_t *__restrict__ slot,
                                                               int n_fine, int32_t *__restrict__ table, int64_t ld) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_fine) return;
    const int j = parent[i];
    const int s = slot[i];
    const int sx = (s >> 2) & 1, sy = (s >> 1) & 1, sz = s & 1;
    uint32_t pocc[8];
    int32_t pbase[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // q = a + 2b + 4c: parent offset (sx - 1 + a, sy - 1 + b, sz - 1 + c)
        const int K = (sx + (q & 1)) + 3 * (sy + ((q >> 1) & 1)) + 9 * (sz + (q >> 2));
        const int32_t jn = __ldg(&ctable[(int64_t)K * ldc + j]);
        pocc[q] = jn ? (uint32_t)__ldg(&cocc[jn - 1]) : 0u;
        pbase[q] = jn ? __ldg(&cbase[jn - 1]) : 0;
    }
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
        const int tz = sz + dz;
        const bool cz = ((tz >> 1) - (sz - 1)) != 0;  // arithmetic shift = floor
        uint32_t o4[4];
        int32_t b4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { o4[q] = cz ? pocc[4 + q] : pocc[q]; b4[q] = cz ? pbase[4 + q] : pbase[q]; }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int ty = sy + dy;
            const bool cy = ((ty >> 1) - (sy - 1)) != 0;
            const uint32_t o2a = cy ? o4[2] : o4[0], o2b = cy ? o4[3] : o4[1];
            const int32_t b2a = cy ? b4[2] : b4[0], b2b = cy ? b4[3] : b4[1];
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int tx = sx + dx;
                const bool cx = ((tx >> 1) - (sx - 1)) != 0;
                const uint32_t o = cx ? o2b : o2a;
                const int32_t b = cx ? b2b : b2a;
                const int sn = ((tx & 1) << 2) | ((ty & 1) << 1) | (tz & 1);
                const int32_t v = ((o >> (7 - sn)) & 1u) ? b + __popc(o >> (8 - sn)) + 1 : 0;
                const int k = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);  // odd kernels enumerate x fastest (hashmap_cuda.cuh:239-258)
                table[(int64_t)k * ld + i] = v;
            }
        }
    }
}

// Row grouping for the tensor-core conv: the kernel skips an offset for a whole 128-row tile only when NO row of
// the tile has that neighbour, so rows with equal neighbour patterns should share tiles.  row_masks gives the
// sort key (bit k = neighbour k present); permute_table rewrites the table in the sorted row order.
__global__ void __launch_bounds__(256) kmap_row_masks_kernel(const int32_t *__restrict__ table, int kvol, int n_out, int64_t ld,
                                                             int32_t *__restrict__ masks) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    uint32_t m = 0;
    for (int k = 0; k < kvol; ++k) m |= (uint32_t)(table[(int64_t)k * ld + o] != 0) << (k < 31 ? k : 31);
    masks[o] = (int32_t)(m & 0x7FFFFFFFu);  // non-negative: any signed sort groups equal patterns
}
__global__ void __launch_bounds__(256) kmap_permute_kernel(const int32_t *__restrict__ table, int n_out, int64_t ld,
                                                           const int32_t *__restrict__ perm, int32_t *__restrict__ out, int64_t ld_out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (j >= n_out) return;
    out[(int64_t)k * ld_out + j] = table[(int64_t)k * ld + perm[j]];
}

// ---------------------------------------------------------------------------------------------
// compaction: k-major table -> offset-major (in,out) pair lists, output index ascending
// ---------------------------------------------------------------------------------------------
constexpr int CHUNK = 2048;  // outputs per block (256 threads x 8)

__global__ void __launch_bounds__(256) compact_count_kernel(const int32_t *__restrict__ table, int n_out, int64_t ld,
                                                            int omit_k, int nchunks, int32_t *__restrict__ counts) {
    int k = blockIdx.y, chunk = blockIdx.x;
    int base = chunk * CHUNK;
    int c = 0;
    if (k != omit_k) {
        for (int j = threadIdx.x; j < CHUNK; j += 256) {
            int o = base + j;
            if (o < n_out && table[(int64_t)k * ld + o] != 0) ++c;
        }
    }
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    int tot = BR(tmp).Sum(c);
    if (threadIdx.x == 0) counts[k * nchunks + chunk] = tot;
}

__global__ void __launch_bounds__(256) compact_write_kernel(const int32_t *__restrict__ table, int n_out, int64_t ld,
                                                            int omit_k, int nchunks, const int32_t *__restrict__ starts,
                                                            int32_t *__restrict__ in_map, int32_t *__restrict__ out_map) {
    int k = blockIdx.y, chunk = blockIdx.x;
    if (k == omit_k) return;
    // thread t owns 8 CONSECUTIVE outputs so that the block-wide exclusive scan preserves output order
    int o0 = chunk * CHUNK + threadIdx.x * 8;
    int32_t v[8];
    int c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int o = o0 + j;
        v[j] = o < n_out ? table[(int64_t)k * ld + o] : 0;
        c += v[j] != 0;
    }
    typedef cub::BlockScan<int, 256> BS;
    __shared__ typename BS::TempStorage tmp;
    int pos;
    BS(tmp).ExclusiveSum(c, pos);
    pos += starts[k * nchunks + chunk];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (v[j] != 0) {
            in_map[pos] = v[j] - 1;
            out_map[pos] = o0 + j;
            ++pos;
        }
    }
}

__global__ void compact_offsets_kernel(const int32_t *starts, const int32_t *counts, int kvol, int nchunks, int32_t *offsets) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < kvol) offsets[k] = starts[k * nchunks];
    if (k == kvol) offsets[kvol] = starts[kvol * nchunks - 1] + counts[kvol * nchunks - 1];
}

// ---------------------------------------------------------------------------------------------
// downsample / upsample
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) ds_flag_kernel(const int32_t *__restrict__ coords, int n, int32_t *__restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = reinterpret_cast<const int4 *>(coords)[i];
    int f = 1;
    if (i > 0) {
        int4 p = reinterpret_cast<const int4 *>(coords)[i - 1];
        f = (c.x != p.x) | ((c.y >> 1) != (p.y >> 1)) | ((c.z >> 1) != (p.z >> 1)) | ((c.w >> 1) != (p.w >> 1));
    }
    flags[i] = f;
}

__global__ void __launch_bounds__(256) ds_write_kernel(const int32_t *__restrict__ coords, int n,
                                                       const int32_t *__restrict__ pos, int32_t *__restrict__ out_coords,
                                                       uint8_t *__restrict__ out_occ, int32_t *__restrict__ parent_of_child,
                                                       uint8_t *__restrict__ slot_of_child, int32_t *__restrict__ n_out_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = pos[i] - 1;
    int4 c = reinterpret_cast<const int4 *>(coords)[i];
    if (parent_of_child) parent_of_child[i] = p;
    if (slot_of_child) slot_of_child[i] = (uint8_t)(((c.y & 1) << 2) | ((c.z & 1) << 1) | (c.w & 1));
    if (i == n - 1) *n_out_dev = p + 1;
    bool first = i == 0 || pos[i - 1] != pos[i];
    if (!first) return;
    reinterpret_cast<int4 *>(out_coords)[p] = make_int4(c.x, c.y >> 1, c.z >> 1, c.w >> 1);
    unsigned occ = 0;
    for (int j = i; j < n && j < i + 8; ++j) {
        if (j > i && pos[j] != pos[i]) break;
        int4 q = reinterpret_cast<const int4 *>(coords)[j];
        int kidx = ((q.y & 1) << 2) | ((q.z & 1) << 1) | (q.w & 1);
        occ |= 1u << (7 - kidx);
    }
    out_occ[p] = (uint8_t)occ;
}

__global__ void __launch_bounds__(256) us_count_kernel(const uint8_t *__restrict__ occ, int n, int32_t *__restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = __popc((unsigned)occ[i]);
}

__global__ void __launch_bounds__(256) us_write_kernel(const int32_t *__restrict__ coords, const uint8_t *__restrict__ occ, int n,
                                                       const int32_t *__restrict__ base, int32_t *__restrict__ child_coords,
                                                       int32_t *__restrict__ child_parent, uint8_t *__restrict__ child_slot,
                                                       int32_t *__restrict__ n_child_dev, const int32_t *__restrict__ shift_add) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned o = occ[i];
    int j = base[i];
    if (i == n - 1) *n_child_dev = j + __popc(o);
    int4 c = reinterpret_cast<const int4 *>(coords)[i];
    int ax = 0, ay = 0, az = 0;
    if (shift_add) { ax = shift_add[0]; ay = shift_add[1]; az = shift_add[2]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (o & (1u << (7 - k))) {
            if (child_coords)
                reinterpret_cast<int4 *>(child_coords)[j] =
                    make_int4(c.x, (c.y << 1) + ((k >> 2) & 1) + ax, (c.z << 1) + ((k >> 1) & 1) + ay, (c.w << 1) + (k & 1) + az);
            if (child_parent) child_parent[j] = i;
            if (child_slot) child_slot[j] = (uint8_t)k;
            ++j;
        }
    }
}

__global__ void __launch_bounds__(256) occ_bits_kernel(const uint8_t *__restrict__ occ, int n, int32_t *__restrict__ bits) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 8) return;
    bits[t] = (occ[t >> 3] >> (7 - (t & 7))) & 1;
}

// ---------------------------------------------------------------------------------------------
// Morton codes (21 bits per axis)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t split_by_3(uint32_t a) {
    uint64_t x = a & 0x1fffff;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void __launch_bounds__(256) morton_kernel(const int32_t *__restrict__ xyz, int64_t ld, int n, int msb_axis,
                                                     int64_t *__restrict__ codes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t *p = xyz + (int64_t)i * ld;
    uint64_t sx = split_by_3((uint32_t)p[0]), sy = split_by_3((uint32_t)p[1]), sz = split_by_3((uint32_t)p[2]);
    uint64_t code = msb_axis == 0 ? ((sx << 2) | (sy << 1) | sz) : ((sz << 2) | (sy << 1) | sx);
    codes[i] = (int64_t)code;
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_hash_insert_coords(int64_t *keys, int32_t *vals, int capacity, const int32_t *coords, int n,
                                       int layout, void *stream) {
    FPCC_REQUIRE(keys && vals && coords, "hash_insert_coords: NULL pointer");
    FPCC_REQUIRE(n >= 0 && capacity > n, "hash_insert_coords: capacity %d must exceed n %d", capacity, n);
    FPCC_REQUIRE(layout == 0 || layout == 1, "hash_insert_coords: bad layout");
    FPCC_REQUIRE(((uintptr_t)coords & 15) == 0, "hash_insert_coords: coords must be 16-byte aligned");
    if (n == 0) return FPCC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    hash_insert_kernel<<<ceil_div(n, 256), 256, 0, s>>>((unsigned long long *)keys, vals, (uint32_t)capacity, coords, n, layout);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_kmap_lookup(const int64_t *keys, const int32_t *vals, int capacity, const int32_t *out_coords,
                                int n_out, int layout, int ksx, int ksy, int ksz, int sx, int sy, int sz, int convention,
                                int32_t *table, int k_major, int64_t ld, void *stream) {
    FPCC_REQUIRE(keys && vals && out_coords && table, "kmap_lookup: NULL pointer");
    FPCC_REQUIRE(ksx > 0 && ksy > 0 && ksz > 0 && sx > 0 && sy > 0 && sz > 0, "kmap_lookup: bad kernel geometry");
    FPCC_REQUIRE(capacity > 0 && n_out >= 0, "kmap_lookup: bad sizes");
    FPCC_REQUIRE(!k_major || ld >= n_out, "kmap_lookup: ld %lld < n_out %d", (long long)ld, n_out);
    FPCC_REQUIRE(((uintptr_t)out_coords & 15) == 0, "kmap_lookup: coords must be 16-byte aligned");
    if (n_out == 0) return FPCC_OK;
    KernelGeom g;
    g.ks[0] = ksx; g.ks[1] = ksy; g.ks[2] = ksz;
    g.st[0] = sx; g.st[1] = sy; g.st[2] = sz;
    g.kvol = ksx * ksy * ksz;
    g.me = convention == 1;
    FPCC_REQUIRE(convention == 0 || convention == 1, "kmap_lookup: unknown convention");
    FPCC_REQUIRE(g.kvol <= 65535, "kmap_lookup: kernel volume too large");
    dim3 grid(ceil_div(n_out, 256), g.kvol);
    kmap_lookup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)keys, vals, (uint32_t)capacity,
                                                               out_coords, n_out, layout, g, table, k_major, ld);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_kmap_from_parent(const int32_t *coarse_table, int64_t ld_coarse, int n_coarse, const uint8_t *coarse_occ,
                                     const int32_t *child_base, const int32_t *parent, const uint8_t *slot, int n_fine,
                                     int32_t *table, int64_t ld, void *stream) {
    FPCC_REQUIRE(coarse_table && coarse_occ && child_base && parent && slot && table, "kmap_from_parent: NULL pointer");
    FPCC_REQUIRE(n_coarse > 0 && ld_coarse >= n_coarse && n_fine >= 0 && ld >= n_fine, "kmap_from_parent: bad sizes");
    if (n_fine == 0) return FPCC_OK;
    kmap_from_parent_kernel<<<ceil_div(n_fine, 256), 256, 0, (cudaStream_t)stream>>>(coarse_table, ld_coarse, coarse_occ, child_base,
                                                                                                  parent, slot, n_fine, table, ld);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_kmap_row_masks(const int32_t *table, int kvol, int n_out, int64_t ld, int32_t *masks, void *stream) {
    FPCC_REQUIRE(table && masks, "kmap_row_masks: NULL pointer");
    FPCC_REQUIRE(kvol > 0 && n_out >= 0 && ld >= n_out, "kmap_row_masks: bad sizes");
    if (n_out == 0) return FPCC_OK;
    kmap_row_masks_kernel<<<ceil_div(n_out, 256), 256, 0, (cudaStream_t)stream>>>(table, kvol, n_out, ld, masks);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_kmap_permute(const int32_t *table, int kvol, int n_out, int64_t ld, const int32_t *perm, int32_t *out,
                                 int64_t ld_out, void *stream) {
    FPCC_REQUIRE(table && perm && out, "kmap_permute: NULL pointer");
    FPCC_REQUIRE(kvol > 0 && kvol <= 65535 && n_out >= 0 && ld >= n_out && ld_out >= n_out, "kmap_permute: bad sizes");
    if (n_out == 0) return FPCC_OK;
    kmap_permute_kernel<<<dim3(ceil_div(n_out, 256), kvol), 256, 0, (cudaStream_t)stream>>>(table, n_out, ld, perm, out, ld_out);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

static size_t scan_temp_bytes(int n) {
    size_t t = 0;
    cub::DeviceScan::ExclusiveSum((void *)nullptr, t, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    return (t + 255) & ~(size_t)255;
}

extern "C" size_t fpcc_kmap_compact_workspace(int kvol, int n_out) {
    int nchunks = ceil_div(n_out > 0 ? n_out : 1, CHUNK);
    size_t cnt = ((size_t)kvol * nchunks * sizeof(int32_t) + 255) & ~(size_t)255;
    return 2 * cnt + scan_temp_bytes(kvol * nchunks) + 256;
}

extern "C" int fpcc_kmap_compact(const int32_t *table, int kvol, int n_out, int64_t ld, int omit_k, int32_t *in_map,
                                 int32_t *out_map, int32_t *offsets, void *workspace, size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(table && in_map && out_map && offsets && workspace, "kmap_compact: NULL pointer");
    FPCC_REQUIRE(kvol > 0 && n_out > 0 && ld >= n_out, "kmap_compact: bad sizes");
    FPCC_REQUIRE(workspace_bytes >= fpcc_kmap_compact_workspace(kvol, n_out), "kmap_compact: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    int nchunks = ceil_div(n_out, CHUNK);
    size_t cnt = ((size_t)kvol * nchunks * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t *counts = (int32_t *)workspace;
    int32_t *starts = (int32_t *)((char *)workspace + cnt);
    void *tmp = (char *)workspace + 2 * cnt;
    size_t tmp_bytes = scan_temp_bytes(kvol * nchunks);
    dim3 grid(nchunks, kvol);
    compact_count_kernel<<<grid, 256, 0, s>>>(table, n_out, ld, omit_k, nchunks, counts);
    FPCC_LAUNCH_CHECK();
    FPCC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, counts, starts, kvol * nchunks, s));
    compact_write_kernel<<<grid, 256, 0, s>>>(table, n_out, ld, omit_k, nchunks, starts, in_map, out_map);
    FPCC_LAUNCH_CHECK();
    compact_offsets_kernel<<<ceil_div(kvol + 1, 128), 128, 0, s>>>(starts, counts, kvol, nchunks, offsets);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" size_t fpcc_scan_workspace(int n) {
    if (n < 1) n = 1;
    size_t arr = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255;
    size_t t = 0;
    cub::DeviceScan::InclusiveSum((void *)nullptr, t, (const int32_t *)nullptr, (int32_t *)nullptr, n);
    size_t t2 = scan_temp_bytes(n);
    if (t2 > t) t = t2;
    return 2 * arr + ((t + 255) & ~(size_t)255) + 256;
}

extern "C" int fpcc_downsample(const int32_t *coords, int n, int32_t *out_coords, uint8_t *out_occ,
                               int32_t *parent_of_child, uint8_t *slot_of_child, int32_t *n_out_dev, void *workspace,
                               size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(coords && out_coords && out_occ && n_out_dev && workspace, "downsample: NULL pointer");
    FPCC_REQUIRE(n > 0, "downsample: empty input");
    FPCC_REQUIRE(workspace_bytes >= fpcc_scan_workspace(n), "downsample: workspace too small");
    FPCC_REQUIRE((((uintptr_t)coords | (uintptr_t)out_coords) & 15) == 0, "downsample: coords must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    size_t arr = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t *flags = (int32_t *)workspace;
    int32_t *pos = (int32_t *)((char *)workspace + arr);
    void *tmp = (char *)workspace + 2 * arr;
    size_t tmp_bytes = workspace_bytes - 2 * arr;
    ds_flag_kernel<<<ceil_div(n, 256), 256, 0, s>>>(coords, n, flags);
    FPCC_LAUNCH_CHECK();
    FPCC_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, flags, pos, n, s));
    ds_write_kernel<<<ceil_div(n, 256), 256, 0, s>>>(coords, n, pos, out_coords, out_occ, parent_of_child, slot_of_child, n_out_dev);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_upsample(const int32_t *coords, const uint8_t *occ, int n, int32_t *child_coords,
                             int32_t *child_parent, uint8_t *child_slot, int32_t *n_child_dev, const int32_t *shift_add,
                             void *workspace, size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(coords && occ && n_child_dev && workspace, "upsample: NULL pointer");
    FPCC_REQUIRE(n > 0, "upsample: empty input");
    FPCC_REQUIRE(workspace_bytes >= fpcc_scan_workspace(n), "upsample: workspace too small");
    FPCC_REQUIRE((((uintptr_t)coords | (uintptr_t)child_coords) & 15) == 0, "upsample: coords must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    size_t arr = ((size_t)n * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t *cnt = (int32_t *)workspace;
    int32_t *base = (int32_t *)((char *)workspace + arr);
    void *tmp = (char *)workspace + 2 * arr;
    size_t tmp_bytes = workspace_bytes - 2 * arr;
    us_count_kernel<<<ceil_div(n, 256), 256, 0, s>>>(occ, n, cnt);
    FPCC_LAUNCH_CHECK();
    FPCC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, base, n, s));
    us_write_kernel<<<ceil_div(n, 256), 256, 0, s>>>(coords, occ, n, base, child_coords, child_parent, child_slot, n_child_dev, shift_add);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_occ_to_bits(const uint8_t *occ, int n, int32_t *bits_i32, void *stream) {
    FPCC_REQUIRE(occ && bits_i32 && n >= 0, "occ_to_bits: bad arguments");
    if (n == 0) return FPCC_OK;
    occ_bits_kernel<<<ceil_div((int64_t)n * 8, 256), 256, 0, (cudaStream_t)stream>>>(occ, n, bits_i32);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_morton_encode(const int32_t *xyz, int64_t ld, int n, int msb_axis, int64_t *codes, void *stream) {
    FPCC_REQUIRE(xyz && codes && n >= 0 && ld >= 3, "morton_encode: bad arguments");
    FPCC_REQUIRE(msb_axis == 0 || msb_axis == 2, "morton_encode: msb_axis must be 0 or 2");
    if (n == 0) return FPCC_OK;
    morton_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(xyz, ld, n, msb_axis, codes);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

namespace fpcc {
// dst[i] = src[idx[i]] for 16-byte rows (coordinates): the permutation into Morton order after the sort
// (model.py:396-398 `xyz = xyz[order]`).  torch's index kernel moves these rows at ~60 GB/s; one int4 per thread
// runs at the sector rate of the random reads.
__global__ void __launch_bounds__(256) gather_rows16_kernel(const int4 *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
                                                            int4 *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __ldg(&src[__ldg(&idx[i])]);
}
}  // namespace fpcc

namespace fpcc {
__global__ void __launch_bounds__(256) occ_bits_q8_kernel(const uint8_t *__restrict__ occ, int64_t n, uint32_t q0, uint32_t q1,
                                                          uint8_t *__restrict__ out, int64_t out_ld) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t o = occ[i];
        uint32_t w0 = 0, w1 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // channel k = bit 7 - k (model.py:63: cur_bin, most significant child first)
            w0 |= (((o >> (7 - k)) & 1u) ? q1 : q0) << (8 * k);
            w1 |= (((o >> (3 - k)) & 1u) ? q1 : q0) << (8 * k);
        }
        *reinterpret_cast<uint4 *>(out + i * out_ld) = make_uint4(w0, w1, 0u, 0u);
    }
}
}  // namespace fpcc

extern "C" int fpcc_occ_bits_q8(const uint8_t *occ, int64_t n, int q0, int q1, int8_t *out, int64_t out_ld, void *stream) {
    FPCC_REQUIRE(occ && out && n >= 0, "occ_bits_q8: bad arguments");
    FPCC_REQUIRE(out_ld >= 16 && out_ld % 16 == 0 && ((uintptr_t)out & 15) == 0, "occ_bits_q8: 16-byte aligned rows expected");
    FPCC_REQUIRE(q0 >= -128 && q0 <= 127 && q1 >= -128 && q1 <= 127, "occ_bits_q8: int8 levels expected");
    if (n == 0) return FPCC_OK;
    const int64_t blocks = (n + 255) / 256, cap = (int64_t)fpcc::sm_count() * 16;
    fpcc::occ_bits_q8_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(occ, n, (uint32_t)(q0 & 0xff), (uint32_t)(q1 & 0xff),
                                                                                                   (uint8_t *)out, out_ld);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_gather_rows16(const void *src, const int64_t *idx, int64_t n, void *dst, void *stream) {
    FPCC_REQUIRE(src && idx && dst && n >= 0, "gather_rows16: bad arguments");
    FPCC_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "gather_rows16: rows must be 16-byte aligned");
    if (n == 0) return FPCC_OK;
    const int64_t blocks = (n + 255) / 256, cap = (int64_t)fpcc::sm_count() * 16;
    fpcc::gather_rows16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>((const int4 *)src, idx, n, (int4 *)dst);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

namespace fpcc {
__global__ void __launch_bounds__(256) slot_table_kernel(const int32_t *__restrict__ parent, const uint8_t *__restrict__ slot,
                                                         int n, int32_t *__restrict__ table, int64_t ld) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int g = blockIdx.y;
    table[(int64_t)g * ld + j] = slot[j] == g ? parent[j] + 1 : 0;
}
}  // namespace fpcc

extern "C" int fpcc_slot_table(const int32_t *child_parent, const uint8_t *child_slot, int n_child, int32_t *table,
                               int64_t ld, void *stream) {
    FPCC_REQUIRE(child_parent && child_slot && table && n_child > 0 && ld >= n_child, "slot_table: bad arguments");
    dim3 grid(fpcc::ceil_div(n_child, 256), 8);
    fpcc::slot_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(child_parent, child_slot, n_child, table, ld);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

// ---------------------------------------------------------------------------------------------
// im2col for thin inputs: patches[o, k*c + j] = feats[table[k][o]-1, j] (0 where there is no neighbour).
// Turns a sparse conv with C_in = 1 / 8 (first conv, occupancy embeds with kernel = stride) into ONE dense
// contraction of K = kvol*C_in per output row, which the tensor-core linear kernel then runs.
// ---------------------------------------------------------------------------------------------
namespace fpcc {
// Generic form: one thread per (output row, offset); byte-granular, any C.
__global__ void __launch_bounds__(256) gather_patches_kernel(const int8_t *__restrict__ feats, int c,
                                                             const int32_t *__restrict__ table, int64_t ld, int kvol, int n_out,
                                                             int8_t *__restrict__ patches, int kp) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (o >= n_out) return;
    int32_t v = table[(int64_t)k * ld + o];
    int8_t *dst = patches + (int64_t)o * kp + k * c;
    for (int j = 0; j < c; ++j) dst[j] = v ? feats[(int64_t)(v - 1) * c + j] : (int8_t)0;
}

// C = 8 (occupancy-bit embeds) with 8-byte aligned rows: a block owns 32 output rows.  The table is read with the
// rows of one offset contiguous (coalesced), the 8-byte feature rows are gathered into a shared-memory tile
// [32 rows][kp], and the tile leaves as whole 16-byte vectors of contiguous patch rows.
constexpr int GP_ROWS = 32;
__global__ void __launch_bounds__(256) gather_patches8_kernel(const int8_t *__restrict__ feats, const int32_t *__restrict__ table,
                                                              int64_t ld, int kvol, int n_out, int8_t *__restrict__ patches, int kp) {
    extern __shared__ __align__(16) uint8_t tile[];  // [GP_ROWS][kp]
    const int o0 = blockIdx.x * GP_ROWS;
    const int rows = min(GP_ROWS, n_out - o0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = warp; k < kvol; k += 8) {   // lanes = consecutive rows of one offset
        uint2 val = make_uint2(0, 0);
        if (lane < rows) {
            const int32_t v = __ldg(&table[(int64_t)k * ld + o0 + lane]);
            if (v) val = __ldg(reinterpret_cast<const uint2 *>(feats) + (v - 1));
        }
        *reinterpret_cast<uint2 *>(tile + lane * kp + k * 8) = val;
    }
    for (int i = threadIdx.x; i < GP_ROWS * ((kp - kvol * 8) / 8); i += 256) {  // zero the pad columns
        const int per = (kp - kvol * 8) / 8;
        *reinterpret_cast<uint2 *>(tile + (i / per) * kp + kvol * 8 + (i % per) * 8) = make_uint2(0, 0);
    }
    __syncthreads();
    const int n16 = rows * kp / 16;  // kp % 16 == 0: the rows of the block are one contiguous byte range
    uint4 *dst = reinterpret_cast<uint4 *>(patches + (int64_t)o0 * kp);
    for (int i = threadIdx.x; i < n16; i += 256) dst[i] = reinterpret_cast<const uint4 *>(tile)[i];
}
}  // namespace fpcc

extern "C" int fpcc_gather_patches(const int8_t *feats, int c, const int32_t *table, int64_t ld, int kvol, int n_out,
                                   int8_t *patches, int kp, void *stream) {
    FPCC_REQUIRE(feats && table && patches, "gather_patches: NULL pointer");
    FPCC_REQUIRE(c > 0 && kvol > 0 && n_out > 0 && ld >= n_out && kp >= kvol * c, "gather_patches: bad sizes");
    if (c == 8 && kp % 16 == 0 && (size_t)fpcc::GP_ROWS * kp <= 48 * 1024 && (((uintptr_t)feats & 7) == 0) && (((uintptr_t)patches & 15) == 0)) {
        fpcc::gather_patches8_kernel<<<fpcc::ceil_div(n_out, fpcc::GP_ROWS), 256, (size_t)fpcc::GP_ROWS * kp, (cudaStream_t)stream>>>(
            feats, table, ld, kvol, n_out, patches, kp);
    } else {
        dim3 grid(fpcc::ceil_div(n_out, 256), kvol);
        fpcc::gather_patches_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feats, c, table, ld, kvol, n_out, patches, kp);
    }
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

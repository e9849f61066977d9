// Batched multi-stream range coders on the GPU: 32-bit rANS, 16-bit probabilities, byte-wise
// renormalisation, L = 2^23 -- bitstreams byte-identical to the reference's CPU coders.
//
// rANS is a serial recurrence per stream (one 32-bit state) and byte-identical output forbids
// splitting a stream, so parallelism is ACROSS streams (one warp per frame / partition) and, inside a
// stream, everything that does not depend on the state is hoisted off the critical path:
//   encode: symbol ranges arrive precomputed (entropy.cu); a warp loads 32 of them coalesced, each lane
//           derives the exact-division constants of its entry (Alverson reciprocal, as the reference's
//           RansEncSymbolInit), and the state update itself is ~10 dependent integer instructions.
//   decode: CDF rows are state-independent, so rows are prefetched 4 symbols ahead as one 16-byte load
//           per lane; the symbol search is a warp-wide compare + redux instead of a binary search.
//
// Replaces lib/entropy_models/rans_coder/rans_byte.h:66-296,
// models/convolutional/lossy_coord_v3/rans_coder/simple_rans_wrapper.cpp:67-95,126-145,206-239 and
// lib/entropy_models/rans_coder/rans_wrapper.cpp:89-185,206-279,326-428.
#include "common.cuh"

namespace fpcc {

constexpr uint32_t RANS_L = 1u << 23;

// Exact x/freq for x < 2^31 by multiply-high (rans_byte.h:201-259): per-entry constants.
struct EncSym {
    uint32_t x_max, rcp, bias;
    uint16_t cmpl, rcp_shift;
};
__device__ __forceinline__ EncSym make_enc_sym(uint32_t start, uint32_t freq, uint32_t bits) {
    EncSym s;
    s.x_max = ((RANS_L >> bits) << 8) * freq;
    s.cmpl = (uint16_t)((1u << bits) - freq);
    if (freq < 2) {
        s.rcp = ~0u;
        s.rcp_shift = 0;
        s.bias = start + (1u << bits) - 1;
    } else {
        uint32_t shift = 32 - __clz(freq - 1);  // ceil(log2(freq))
        s.rcp = (uint32_t)(((1ull << (shift + 31)) + freq - 1) / freq);
        s.rcp_shift = (uint16_t)(shift - 1);
        s.bias = start;
    }
    return s;
}

// Encoding is two kernels.  (1) enc_prepare_kernel, fully parallel: every entry gets its exact-division
// constants {x_max, rcp, bias, cmpl | rcp_shift << 16}.  (2) rans_encode_kernel, one warp per stream: entries
// of stream b are [rng_off[b], rng_off[b+1]) in DECODE order and are consumed back to front (rANS is LIFO);
// bytes are written backwards from the end of the stream's slot, collected in a register and stored as
// aligned 32-bit words.
__global__ void __launch_bounds__(256) enc_prepare_kernel(const uint32_t *__restrict__ ranges, const uint8_t *__restrict__ bits,
                                                          int64_t total, uint4 *__restrict__ recs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t r = ranges[i];
    EncSym es = make_enc_sym(r & 0xFFFFu, (r >> 16) + 1u, bits ? (uint32_t)bits[i] : 16u);
    recs[i] = make_uint4(es.x_max, es.rcp, es.bias, (uint32_t)es.cmpl | ((uint32_t)es.rcp_shift << 16));
}

// One rANS step on record r, branch-free.  The state recurrence is all that stays serial: the bytes that leave
// (0..2, low byte first) are handed to the packer warp as one word, low 16 bits of the old state | count << 16.
__device__ __forceinline__ int enc_step(uint32_t &x, const uint4 r, uint32_t &emit) {
    // RansEncRenorm, rans_byte.h:77-89: x < 2^31 and x_max >= 2^15, so at most two bytes leave.  The three possible
    // renormalised states and their quotients are formed side by side and selected afterwards: the serial
    // dependency per symbol is shift -> multiply-high -> select -> shift -> multiply-add.
    const uint32_t x1 = x >> 8, x2 = x >> 16;
    const bool c1 = x >= r.x, c2 = x1 >= r.x;
    const uint32_t q0 = __umulhi(x, r.y), q1 = __umulhi(x1, r.y), q2 = __umulhi(x2, r.y);  // exact x / freq (rans_byte.h:201-259)
    const uint32_t xs = c2 ? x2 : (c1 ? x1 : x);
    const uint32_t qs = c2 ? q2 : (c1 ? q1 : q0);
    const int nb = (int)c1 + (int)c2;
    emit = (x & 0xFFFFu) | ((uint32_t)nb << 16);
    x = (xs + r.z) + (qs >> (r.w >> 16)) * (r.w & 0xFFFFu);
    return nb;
}

// Two warps per stream.  Warp 0: all lanes stream the records of the stream, back to front, through a
// shared-memory ring with cp.async (ENC_RING-1 chunks of ENC_CH records in flight); lane 0 runs the serial
// recurrence on them at shared-memory latency.  Warp 1 packs: per chunk it turns the per-symbol byte counts into
// positions with a warp scan and stores the bytes (backwards from the end of the stream's slot), off the critical
// path of the recurrence.  The two warps hand chunks over through named barriers (double-buffered).
constexpr int ENC_CH = 256, ENC_RING = 4;
__device__ __forceinline__ void nbar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id) {
    __threadfence_block();
    asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}

__global__ void __launch_bounds__(64) rans_encode_kernel(const uint4 *__restrict__ recs, const int64_t *__restrict__ rng_off,
                                                         uint8_t *__restrict__ out, int64_t out_stride,
                                                         int32_t *__restrict__ out_len, uint32_t *__restrict__ state_io,
                                                         int do_flush) {
    __shared__ uint4 ring[ENC_RING][ENC_CH];
    __shared__ __align__(16) uint32_t emit[2][ENC_CH];
    __shared__ int emit_n[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x;
    const int64_t lo = rng_off[b], hi = rng_off[b + 1];
    uint8_t *const base = out + (int64_t)b * out_stride;
    uint8_t *ptr = base + out_stride;
    uint32_t x = RANS_L;
    bool overflow = false;
    if (state_io) {  // resume a stream left open by an earlier call (RansEncoder keeps its state between encode() calls)
        x = state_io[2 * b];
        uint32_t written = state_io[2 * b + 1];
        if (written == 0xFFFFFFFFu) overflow = true; else ptr -= written;
    }
    // named barriers: 1 + buf = "emit[buf] filled", 3 + buf = "emit[buf] drained", 5 = join
    if (warp == 1) {  // ---- packer ----
        for (int64_t c = 0;; ++c) {
            const int buf = (int)(c & 1);
            nbar_sync(1 + buf);
            if (emit_n[buf] == 0) break;
            const uint4 w0 = reinterpret_cast<const uint4 *>(emit[buf])[2 * lane], w1 = reinterpret_cast<const uint4 *>(emit[buf])[2 * lane + 1];
            const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            int local = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) local += (int)(w[k] >> 16);
            int incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            uint8_t *q = ptr - (incl - local);  // first byte of this lane's first symbol goes to q[-1]
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int nb = (int)(w[k] >> 16);
                if (nb >= 1) q[-1] = (uint8_t)(w[k] & 0xffu);
                if (nb >= 2) q[-2] = (uint8_t)((w[k] >> 8) & 0xffu);
                q -= nb;
            }
            ptr -= __shfl_sync(0xffffffffu, incl, 31);
            nbar_arrive(3 + buf);
        }
        nbar_sync(5);
        return;
    }
    int64_t top = hi;
    {
        // Chunk c holds records [hi - (c+1)*ENC_CH, hi - c*ENC_CH).
        const int64_t n_chunks = overflow ? 0 : (hi - lo) / ENC_CH;
        auto issue = [&](int64_t c) {
            if (c < n_chunks) {
                const uint4 *src = recs + (hi - (c + 1) * ENC_CH);
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&ring[c % ENC_RING][0]);
#pragma unroll
                for (int t = 0; t < ENC_CH / 32; ++t)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)((t * 32 + lane) * 16)), "l"(src + t * 32 + lane) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int c = 0; c < ENC_RING - 1; ++c) issue(c);
        int64_t room = ptr - base;  // bytes still free in front of the write position (lane 0 keeps it exact)
        int64_t c = 0;
        for (; c < n_chunks; ++c) {
            const int buf = (int)(c & 1);
            issue(c + ENC_RING - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(ENC_RING - 1) : "memory");
            __syncwarp();
            if (c >= 2) nbar_sync(3 + buf);  // packer has drained this buffer (chunk c-2)
            int go = 1;
            if (lane == 0) {
                go = room >= 2 * ENC_CH + 16;  // worst case two bytes per record; the flush needs 4 more
                if (go) {
                    const uint4 *rb = ring[c % ENC_RING];
                    uint32_t *em = emit[buf];
                    int total = 0;
                    constexpr int U = 8;
#pragma unroll 1
                    for (int i = ENC_CH - U; i >= 0; i -= U) {
                        uint4 r[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) r[u] = rb[i + U - 1 - u];
#pragma unroll
                        for (int u = 0; u < U; ++u) total += enc_step(x, r[u], em[ENC_CH - U - i + u]);
                    }
                    room -= total;
                }
                emit_n[buf] = go ? ENC_CH : 0;
            }
            go = __shfl_sync(0xffffffffu, go, 0);
            nbar_arrive(1 + buf);
            if (!go) break;
            top -= ENC_CH;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (c == n_chunks) {  // normal end: post the terminating empty chunk
            const int buf = (int)(c & 1);
            if (c >= 2) nbar_sync(3 + buf);
            if (lane == 0) emit_n[buf] = 0;
            __syncwarp();
            nbar_arrive(1 + buf);
        }
        nbar_sync(5);  // every byte of the chunks above is in place
        room = __shfl_sync(0xffffffffu, (int)room, 0);
        ptr = base + room;
    }
    if (lane != 0) return;
    // tail (and unaligned resume points): plain loop with byte stores
    uint8_t *const guard = base + 8;
    for (; top > lo; --top) {
        const uint4 r = __ldg(&recs[top - 1]);
        while (x >= r.x) {
            if (ptr > guard) *--ptr = (uint8_t)(x & 0xff); else overflow = true;
            x >>= 8;
        }
        const uint32_t q = __umulhi(x, r.y) >> (r.w >> 16);
        x = x + r.z + q * (r.w & 0xFFFFu);
    }
    if (!do_flush) {
        state_io[2 * b] = x;
        state_io[2 * b + 1] = overflow ? 0xFFFFFFFFu : (uint32_t)(base + out_stride - ptr);
        out_len[b] = overflow ? -1 : (int32_t)(base + out_stride - ptr);
        return;
    }
    // RansEncFlush, rans_byte.h:109-121
    if (ptr - 4 >= base && !overflow) {
        ptr -= 4;
        for (int t = 0; t < 4; ++t) ptr[t] = (uint8_t)(x >> (8 * t));
        out_len[b] = (int32_t)(base + out_stride - ptr);
    } else {
        out_len[b] = -1;
    }
}

// ---- decoding ------------------------------------------------------------------------------------

struct ByteReader {
    const uint8_t *p;
    uint32_t pos, len, err;
    __device__ __forceinline__ uint32_t next() {
        uint32_t v = 0;
        if (pos < len) v = p[pos]; else err = 1;
        ++pos;
        return v;
    }
};

__global__ void rans_dec_init_kernel(fpcc_rans_dec_state *st, const uint8_t *__restrict__ bytes,
                                     const int64_t *__restrict__ byte_off, const int32_t *__restrict__ byte_len, int n) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const uint8_t *p = bytes + byte_off[b];
    fpcc_rans_dec_state s;
    s.len = (uint32_t)byte_len[b];
    s.err = s.len < 4;
    s.x = s.err ? RANS_L : ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    if (s.x < RANS_L) { s.x = RANS_L; s.err = 1; }  // not a valid final encoder state
    s.pos = 4;
    st[b] = s;
}

__device__ __forceinline__ uint32_t dec_advance(uint32_t x, ByteReader &br, uint32_t start, uint32_t freq, uint32_t bits) {
    x = freq * (x >> bits) + (x & ((1u << bits) - 1u)) - start;  // RansDecAdvance, rans_byte.h:149-165
    while (x < RANS_L) x = (x << 8) | br.next();
    return x;
}

// Big-endian window over the byte stream (next byte on top), refilled with aligned 32-bit loads that are
// issued one refill ahead of their use.  Reads at most 3 bytes before / 7 bytes past the stream.
__device__ __forceinline__ int32_t le_mask(uint32_t e, int32_t cf_plus_1) {  // -1 if e <= cf else 0 (e, cf < 2^16)
    int32_t m;
    asm("{\n\t.reg .s32 d;\n\tsub.s32 d, %1, %2;\n\tshr.s32 %0, d, 31;\n\t}" : "=r"(m) : "r"((int32_t)e), "r"(cf_plus_1));
    return m;
}

struct ByteWindow {
    const uint32_t *wp;
    uint64_t win;
    int avail;
    uint32_t nextw;
    __device__ __forceinline__ void init(const uint8_t *p) {
        uintptr_t a = (uintptr_t)p;
        uint32_t mis = (uint32_t)(a & 3);
        wp = reinterpret_cast<const uint32_t *>(a - mis);
        uint32_t w = __byte_perm(*wp++, 0, 0x0123);
        win = (uint64_t)(w << (8 * mis)) << 32;
        avail = 4 - (int)mis;
        nextw = *wp++;
        refill();
    }
    __device__ __forceinline__ void refill() {  // written with selects: no branch in the serial loop
        const bool need = avail <= 4;
        const uint64_t w = (uint64_t)__byte_perm(nextw, 0, 0x0123) << ((32 - 8 * avail) & 63);
        win |= need ? w : 0ull;
        avail += need ? 4 : 0;
        if (need) {
            nextw = *wp++;
            if (((uintptr_t)wp & 127) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(wp + 64));  // two lines ahead
        }
    }
    __device__ __forceinline__ void consume(int n) {  // drops n (0..2) bytes from the front of the window
        win <<= 8 * n;
        avail -= n;
        refill();
    }
    // Shifts the next n (0..2) bytes of the stream into x from the right: x = (x << 8n) | bytes.
    __device__ __forceinline__ uint32_t shift_in(uint32_t x, int n) {
        x = __funnelshift_l((uint32_t)(win >> 32), x, 8 * n);
        win <<= 8 * n;
        avail -= n;
        refill();
        return x;
    }
};

// ---- fast path: one CDF row per symbol, S <= 256 entries, 256-entry (512 B) row pitch -----------------------
// Two warps per stream.  Warp 1 is a one-thread producer: it streams the rows, which do not depend on the coder
// state, into a shared-memory ring with bulk async copies (8 rows = 4 KB per copy, 8 slots), so the decoder never
// waits on L2/HBM.  Warp 0 decodes; all its lanes carry the same state.  The symbol search is two ballots: a
// coarse one over every 8th entry and a fine one over the 9 entries around the hit.
constexpr int DEC_SLOT_ROWS = 8, DEC_SLOTS = 8;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "DEC_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni DEC_DONE;\n\t"
        "bra.uni DEC_WAIT;\n\t"
        "DEC_DONE:\n\t"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(64) rans_decode_rows_kernel(fpcc_rans_dec_state *st, const uint8_t *__restrict__ bytes,
                                                              const int64_t *__restrict__ byte_off, const uint16_t *__restrict__ cdf,
                                                              int S, const int64_t *__restrict__ row_off, int32_t *__restrict__ symbols) {
    __shared__ __align__(128) uint16_t ring[DEC_SLOTS][DEC_SLOT_ROWS][256];
    __shared__ uint64_t full[DEC_SLOTS], empty[DEC_SLOTS];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t lo = row_off[b], hi = row_off[b + 1];
    const int64_t n = hi - lo;
    const int64_t nslots = (n + DEC_SLOT_ROWS - 1) / DEC_SLOT_ROWS;
    if (threadIdx.x == 0) {
        for (int i = 0; i < DEC_SLOTS; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[i])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&empty[i])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        if (lane == 0) {
            for (int64_t sl = 0; sl < nslots; ++sl) {
                const int slot = (int)(sl % DEC_SLOTS);
                bar_wait(&empty[slot], (uint32_t)(((sl / DEC_SLOTS) & 1) ^ 1));
                const int64_t rows = min((int64_t)DEC_SLOT_ROWS, n - sl * DEC_SLOT_ROWS);
                const uint32_t nbytes = (uint32_t)(rows * 512);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&full[slot])), "r"(nbytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr(&ring[slot][0][0])), "l"(cdf + (lo + sl * DEC_SLOT_ROWS) * 256), "r"(nbytes),
                               "r"(smem_addr(&full[slot])) : "memory");
            }
        }
        return;
    }
    fpcc_rans_dec_state s = st[b];
    const uint8_t *p = bytes + byte_off[b];
    uint32_t x = s.x, pos = s.pos;
    ByteWindow bw;
    bw.init(p + pos);
    // Lane l owns entries 8l .. 8l+7 of the current row (one 16-byte shared load, issued one symbol ahead: rows do
    // not depend on the coder state).  With the state known, every lane counts its entries <= cf (the CDF is
    // non-decreasing, so they form a prefix), one warp reduction gives the symbol, and two broadcast loads fetch
    // its range: compare -> redux -> load -> multiply-add is the whole serial dependency.
    const int nv = min(8, max(0, S - 1 - 8 * lane));  // entries j >= S-1 never count (simple_rans_wrapper.cpp:225-228)
    for (int64_t sl = 0; sl < nslots; ++sl) {
        const int slot = (int)(sl % DEC_SLOTS);
        bar_wait(&full[slot], (uint32_t)((sl / DEC_SLOTS) & 1));
        const int rows = (int)min((int64_t)DEC_SLOT_ROWS, n - sl * DEC_SLOT_ROWS);
        const int64_t i0 = lo + sl * DEC_SLOT_ROWS;
        uint4 ev = reinterpret_cast<const uint4 *>(ring[slot][0])[lane];
        for (int r = 0; r < rows; ++r) {
            const uint16_t *row = ring[slot][r];
            const uint32_t e0 = ev.x & 0xFFFFu, e1 = ev.x >> 16, e2 = ev.y & 0xFFFFu, e3 = ev.y >> 16;
            const uint32_t e4 = ev.z & 0xFFFFu, e5 = ev.z >> 16, e6 = ev.w & 0xFFFFu, e7 = ev.w >> 16;
            if (r + 1 < rows) ev = reinterpret_cast<const uint4 *>(ring[slot][r + 1])[lane];
            const uint32_t cf = x & 0xFFFFu;
            // entries < 2^16: (e - cf - 1) >> 31 is -1 exactly when e <= cf; eight independent subtract/shift pairs and
            // a three-input add tree instead of a chain of compare-and-select
            const int32_t t1 = (int32_t)cf + 1;
            const int cnt8 = -((le_mask(e0, t1) + le_mask(e1, t1) + le_mask(e2, t1)) + (le_mask(e3, t1) + le_mask(e4, t1) + le_mask(e5, t1)) +
                               (le_mask(e6, t1) + le_mask(e7, t1)));
            const int cnt = min(cnt8, nv);
            const int sym = __reduce_add_sync(0xffffffffu, cnt);  // #{j < S-1 : cdf[j] <= cf}
            const uint32_t start = sym > 0 ? (uint32_t)row[sym - 1] : 0u;
            const uint32_t end = sym == S - 1 ? 0x10000u : (uint32_t)row[sym];
            x = (end - start) * (x >> 16) + cf - start;                   // RansDecAdvance, rans_byte.h:149-165
            // renormalisation, 0..2 bytes: both refilled candidates are formed from the window before the compares
            // resolve, then selected
            const uint32_t whi = (uint32_t)(bw.win >> 32);
            const uint32_t c1 = __funnelshift_l(whi, x, 8), c2 = __funnelshift_l(whi, x, 16);
            const bool p0 = x >= RANS_L, p1 = x >= (1u << 15);
            const int nb = (int)!p0 + (int)!p1;
            x = p0 ? x : (p1 ? c1 : c2);
            bw.consume(nb);
            pos += nb;
            if (lane == 0) symbols[i0 + r] = sym;
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&empty[slot])) : "memory");
    }
    if (lane == 0) {
        s.x = x; s.pos = pos; s.err = s.err | (uint32_t)(pos > s.len);
        st[b] = s;
    }
}

// ---- generic path: one warp per stream; binary search on a shared table, a per-stream table, or long rows ----
__global__ void __launch_bounds__(32) rans_decode_kernel(fpcc_rans_dec_state *st, const uint8_t *__restrict__ bytes,
                                                         const int64_t *__restrict__ byte_off, const uint16_t *__restrict__ cdf,
                                                         int64_t n_cdf, int S, int ld, const int64_t *__restrict__ row_off,
                                                         int32_t *__restrict__ symbols, int rows_per_stream,
                                                         const int32_t *__restrict__ s_per_stream) {
    const int b = blockIdx.x, lane = threadIdx.x;
    if (s_per_stream) S = s_per_stream[b];
    fpcc_rans_dec_state s = st[b];
    ByteReader br = {bytes + byte_off[b], s.pos, s.len, s.err};
    uint32_t x = s.x;
    const int64_t lo = row_off[b], hi = row_off[b + 1];
    {
        // generic path: binary search by every lane on the same row (shared table, or S > 256)
        for (int64_t i = lo; i < hi; ++i) {
            const uint16_t *row = n_cdf == 1 ? cdf : cdf + (rows_per_stream ? (int64_t)b : i) * (int64_t)ld;
            uint32_t cf = x & 0xFFFFu;
            int a = 0, c = S;
            while (a < c) {
                int mid = a + ((c - a) >> 1);
                if ((uint32_t)__ldg(&row[mid]) <= cf) a = mid + 1; else c = mid;
            }
            int sym = a > S - 1 ? S - 1 : a;
            uint32_t start = sym == 0 ? 0u : __ldg(&row[sym - 1]);
            uint32_t end = sym == S - 1 ? 65536u : __ldg(&row[sym]);
            x = dec_advance(x, br, start, end - start, 16);
            if (lane == 0) symbols[i] = sym;
        }
    }
    if (lane == 0) {
        s.x = x; s.pos = br.pos; s.err = br.err;
        st[b] = s;
    }
}

// ---- BinaryRansCoder ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) binary_ranges_kernel(const uint8_t *__restrict__ sym, const uint32_t *__restrict__ prob,
                                                            int64_t total, uint32_t *__restrict__ ranges) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t p = prob[i];  // P(1) * 65536
    uint32_t start = sym[i] ? 65536u - p : 0u;
    uint32_t freq = sym[i] ? p : 65536u - p;
    ranges[i] = start | ((freq - 1u) << 16);
}

__global__ void __launch_bounds__(32) binary_decode_kernel(const uint8_t *__restrict__ bytes, const int64_t *__restrict__ byte_off,
                                                           const int32_t *__restrict__ byte_len, const uint32_t *__restrict__ prob,
                                                           int64_t n, uint8_t *__restrict__ symbols, int32_t *__restrict__ err) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const uint8_t *p = bytes + byte_off[b];
    ByteReader br = {p, 4, (uint32_t)byte_len[b], (uint32_t)(byte_len[b] < 4)};
    uint32_t x = br.err ? RANS_L : ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    if (x < RANS_L) { x = RANS_L; br.err = 1; }
    const uint32_t *pr = prob + (int64_t)b * n;
    uint8_t *out = symbols + (int64_t)b * n;
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
        uint32_t mine = i0 + lane < n ? pr[i0 + lane] : 1u;
        int cnt = (int)min((int64_t)32, n - i0);
        uint32_t got = 0;
        for (int j = 0; j < cnt; ++j) {
            uint32_t pj = __shfl_sync(0xffffffffu, mine, j);
            uint32_t thr = 65536u - pj;
            uint32_t bit = (x & 0xFFFFu) >= thr;
            x = bit ? dec_advance(x, br, thr, pj, 16) : dec_advance(x, br, 0, thr, 16);
            got |= bit << j;
        }
        if (i0 + lane < n) out[i0 + lane] = (got >> lane) & 1u;
    }
    if (lane == 0 && br.err) atomicExch(err, 1);
}

// ---- IndexedRansCoder ----------------------------------------------------------------------------
struct Tables {
    const uint32_t *flat;
    const int64_t *off;
    const int32_t *len;
    const int32_t *offsets;
    int n_tables;
};

// number of coder entries symbol i expands to: 1, plus sign + Elias-gamma bits when it escapes
__device__ __forceinline__ int indexed_entries(const Tables &t, int overflow, int32_t sym, int tbl, int32_t *v_out, int32_t *gamma_out, int *sign_out) {
    int32_t nsym = t.len[tbl] - 1;
    int32_t v = sym - t.offsets[tbl];
    int32_t gamma = 0;
    int sign = 0, n = 1;
    if (overflow) {
        int32_t maxv = nsym - 1;
        sign = v < 0;
        if (sign) { gamma = -v; v = maxv; }
        else if (v >= maxv) { gamma = v - maxv + 1; v = maxv; }
        if (v == maxv) {
            int nb = 32 - __clz(gamma);  // gamma >= 1
            n += 1 + nb + (nb - 1);      // sign, nb value bits, nb-1 zeros (rans_wrapper.cpp:153-167)
        }
    }
    *v_out = v; *gamma_out = gamma; *sign_out = sign;
    return n;
}

__global__ void __launch_bounds__(256) indexed_count_kernel(Tables t, int overflow, const int32_t *__restrict__ symbols,
                                                            const int32_t *__restrict__ indexes, int64_t n, int64_t total,
                                                            int32_t *__restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int tbl = indexes ? indexes[i] : (int)((i % n) % t.n_tables);
    int32_t v, g; int sgn;
    counts[i] = indexed_entries(t, overflow, symbols[i], tbl, &v, &g, &sgn);
}

// Entries are written in DECODE order: symbol first, then (escape only) the unary zeros, the value bits MSB
// first ... i.e. the exact reverse of the encoder's push order at rans_wrapper.cpp:153-170.
__global__ void __launch_bounds__(256) indexed_ranges_kernel(Tables t, int overflow, const int32_t *__restrict__ symbols,
                                                             const int32_t *__restrict__ indexes, int64_t n, int64_t total,
                                                             const int64_t *__restrict__ entry_pos, uint32_t *__restrict__ ranges,
                                                             uint8_t *__restrict__ bits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int tbl = indexes ? indexes[i] : (int)((i % n) % t.n_tables);
    int32_t v, gamma; int sign;
    int cnt = indexed_entries(t, overflow, symbols[i], tbl, &v, &gamma, &sign);
    const uint32_t *cdf = t.flat + t.off[tbl];
    int64_t p = entry_pos[i];
    uint32_t lo = cdf[v], hi = cdf[v + 1];
    ranges[p] = lo | ((hi - lo - 1u) << 16);
    bits[p] = 16;
    ++p;
    if (cnt > 1) {
        int nb = 32 - __clz(gamma);
        // decoder reads: (nb-1) zeros, then bits nb-1 .. 0 of gamma (the top one is the terminating 1), then sign
        for (int z = 0; z < nb - 1; ++z) { ranges[p] = 0u; bits[p] = 1; ++p; }
        for (int k = nb - 1; k >= 0; --k) { ranges[p] = (uint32_t)((gamma >> k) & 1); bits[p] = 1; ++p; }
        ranges[p] = (uint32_t)sign; bits[p] = 1;
    }
}

__global__ void __launch_bounds__(32) indexed_decode_kernel(Tables t, int overflow, const uint8_t *__restrict__ bytes,
                                                            const int64_t *__restrict__ byte_off, const int32_t *__restrict__ byte_len,
                                                            const int32_t *__restrict__ indexes, int64_t n,
                                                            int32_t *__restrict__ symbols, int32_t *__restrict__ err) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const uint8_t *p = bytes + byte_off[b];
    ByteReader br = {p, 4, (uint32_t)byte_len[b], (uint32_t)(byte_len[b] < 4)};
    uint32_t x = br.err ? RANS_L : ((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    if (x < RANS_L) { x = RANS_L; br.err = 1; }
    for (int64_t j = 0; j < n; ++j) {
        int tbl = indexes ? indexes[(int64_t)b * n + j] : (int)(j % t.n_tables);
        const uint32_t *cdf = t.flat + t.off[tbl];
        int len = t.len[tbl];
        uint32_t cf = x & 0xFFFFu;
        int a = 1, c = len;  // upper_bound(cdf+1, cdf+len, cf) - cdf - 1, rans_wrapper.cpp:243
        while (a < c) {
            int mid = a + ((c - a) >> 1);
            if (__ldg(&cdf[mid]) <= cf) a = mid + 1; else c = mid;
        }
        int32_t v = a - 1;
        uint32_t lo = __ldg(&cdf[v]), hi = __ldg(&cdf[v + 1]);
        x = dec_advance(x, br, lo, hi - lo, 16);
        if (overflow && v == len - 2) {
            int32_t maxv = len - 2;
            int nb = 0;
            while ((x & 1u) == 0 && !br.err) { ++nb; x = dec_advance(x, br, 0, 1, 1); }
            x = dec_advance(x, br, 1, 1, 1);
            v = 1 << nb;
            while (--nb >= 0) {
                int32_t bit = (int32_t)(x & 1u);
                x = dec_advance(x, br, (uint32_t)bit, 1, 1);
                v |= bit << nb;
            }
            int32_t sign = (int32_t)(x & 1u);
            x = dec_advance(x, br, (uint32_t)sign, 1, 1);
            v = sign ? -v : v + maxv - 1;
        }
        if (lane == 0) symbols[(int64_t)b * n + j] = v + t.offsets[tbl];
    }
    if (lane == 0 && br.err) atomicExch(err, 1);
}

// ---- pmf -> quantised CDF (cdf_ops.cpp:4-109), one thread per table ---------------------------------
__global__ void pmf_to_cdf_kernel(double *__restrict__ pmf, int n_tables, int pmf_size, int32_t *__restrict__ offsets,
                                  int overflow, uint32_t *__restrict__ cdf_out, int32_t *__restrict__ cdf_len) {
    int tbl = blockIdx.x * blockDim.x + threadIdx.x;
    if (tbl >= n_tables) return;
    double *p = pmf + (int64_t)tbl * pmf_size;
    uint32_t *cdf = cdf_out + (int64_t)tbl * (pmf_size + 2);
    int len = overflow ? pmf_size + 2 : pmf_size + 1;
    for (int i = 0; i < len; ++i) cdf[i] = 0;
    double total = 0.;
    for (int i = 0; i < pmf_size; ++i) total += p[i];
    double over = 1. - total > 0. ? 1. - total : 0.;
    if (overflow) total += over;
    for (int i = 1; i < pmf_size; ++i) p[i] = p[i - 1] + p[i];
    for (int i = 0; i < pmf_size; ++i) cdf[i + 1] = (uint32_t)round(65536.0 * (p[i] / total));
    cdf[len - 1] = 65536u;
    if (overflow) {
        int s = 0, e = 0;
        for (int i = 0; i < len - 1; i++) if (cdf[i + 1] != cdf[i]) { s = i; break; }
        for (int i = len - 2; i > 0; i--) if (cdf[i - 1] != cdf[i]) { e = i; break; }
        offsets[tbl] += s;
        if (s > e) { s = len - 3; e = s + 1; }
        int nlen = e - s + 2;
        for (int i = 0; i < nlen - 1; i++) cdf[i] = cdf[i + s];
        len = nlen;
        cdf[len - 1] = 65536u;
    }
    for (int i = 0; i < len - 1; i++) {
        if (cdf[i + 1] == cdf[i]) {
            uint32_t best = ~0u;
            int steal = -1;
            for (int j = 0; j < len - 1; j++) {
                uint32_t f = cdf[j + 1] - cdf[j];
                if (f > 1 && f < best) { best = f; steal = j; }
            }
            if (steal < 0) { len = -1; break; }
            if (steal < i) { for (int j = steal + 1; j <= i; j++) cdf[j]--; }
            else { for (int j = i + 1; j <= steal; j++) cdf[j]++; }
        }
    }
    cdf_len[tbl] = len;
}

}  // namespace fpcc

using namespace fpcc;

extern "C" int fpcc_rans_encode(const uint32_t *ranges, const uint8_t *bits, const int64_t *rng_off, int n_streams,
                                int64_t total_entries, uint8_t *out, int64_t out_stride, int32_t *out_len,
                                uint32_t *state_io, int do_flush, void *workspace, size_t workspace_bytes, void *stream) {
    FPCC_REQUIRE(rng_off && out && out_len, "rans_encode: NULL pointer");
    FPCC_REQUIRE(n_streams > 0 && out_stride >= 16 && total_entries >= 0, "rans_encode: bad sizes");
    FPCC_REQUIRE(total_entries == 0 || ranges, "rans_encode: NULL ranges");
    FPCC_REQUIRE(do_flush || state_io, "rans_encode: an open (unflushed) stream needs state_io");
    FPCC_REQUIRE(total_entries == 0 || (workspace && workspace_bytes >= (size_t)total_entries * sizeof(uint4) &&
                                        ((uintptr_t)workspace & 15) == 0),
                 "rans_encode: workspace must hold 16 bytes per entry (16-byte aligned)");
    cudaStream_t s = (cudaStream_t)stream;
    uint4 *recs = (uint4 *)workspace;
    if (total_entries > 0) {
        enc_prepare_kernel<<<ceil_div(total_entries, 256), 256, 0, s>>>(ranges, bits, total_entries, recs);
        FPCC_LAUNCH_CHECK();
    }
    rans_encode_kernel<<<n_streams, 64, 0, s>>>(recs, rng_off, out, out_stride, out_len, state_io, do_flush);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_rans_dec_init(fpcc_rans_dec_state *st, const uint8_t *bytes, const int64_t *byte_off,
                                  const int32_t *byte_len, int n_streams, void *stream) {
    FPCC_REQUIRE(st && bytes && byte_off && byte_len && n_streams > 0, "rans_dec_init: bad arguments");
    rans_dec_init_kernel<<<ceil_div(n_streams, 64), 64, 0, (cudaStream_t)stream>>>(st, bytes, byte_off, byte_len, n_streams);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_rans_decode(fpcc_rans_dec_state *st, const uint8_t *bytes, const int64_t *byte_off,
                                const uint16_t *cdf, int64_t n_cdf, int s, int ld, const int64_t *row_off, int n_streams,
                                int32_t *symbols, int rows_per_stream, const int32_t *s_per_stream, void *stream) {
    FPCC_REQUIRE(st && bytes && byte_off && cdf && row_off && symbols, "rans_decode: NULL pointer");
    FPCC_REQUIRE(n_streams > 0 && s > 0 && ld >= s, "rans_decode: bad sizes");
    FPCC_REQUIRE(ld != 256 || n_cdf == 1 || ((uintptr_t)cdf & 15) == 0, "rans_decode: padded CDF rows must be 16-byte aligned");
    if (n_cdf != 1 && !rows_per_stream && !s_per_stream && ld == 256 && s <= 256 && ((uintptr_t)cdf & 15) == 0)
        rans_decode_rows_kernel<<<n_streams, 64, 0, (cudaStream_t)stream>>>(st, bytes, byte_off, cdf, s, row_off, symbols);
    else
        rans_decode_kernel<<<n_streams, 32, 0, (cudaStream_t)stream>>>(st, bytes, byte_off, cdf, n_cdf, s, ld, row_off, symbols,
                                                                        rows_per_stream, s_per_stream);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_rans_binary_ranges(const uint8_t *symbols, const uint32_t *prob, int64_t total, uint32_t *ranges, void *stream) {
    FPCC_REQUIRE(symbols && prob && ranges && total > 0, "rans_binary_ranges: bad arguments");
    binary_ranges_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(symbols, prob, total, ranges);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_rans_binary_decode(const uint8_t *bytes, const int64_t *byte_off, const int32_t *byte_len,
                                       const uint32_t *prob, int64_t n, int n_streams, uint8_t *symbols, int32_t *err_dev,
                                       void *stream) {
    FPCC_REQUIRE(bytes && byte_off && byte_len && prob && symbols && err_dev, "rans_binary_decode: NULL pointer");
    FPCC_REQUIRE(n > 0 && n_streams > 0, "rans_binary_decode: bad sizes");
    binary_decode_kernel<<<n_streams, 32, 0, (cudaStream_t)stream>>>(bytes, byte_off, byte_len, prob, n, symbols, err_dev);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_indexed_count(const uint32_t *cdf_flat, const int64_t *cdf_off, const int32_t *cdf_len, int n_tables,
                                  const int32_t *offsets, int overflow, const int32_t *symbols, const int32_t *indexes,
                                  int64_t n, int n_streams, int32_t *entries_per_symbol, void *stream) {
    FPCC_REQUIRE(cdf_flat && cdf_off && cdf_len && offsets && symbols && entries_per_symbol, "indexed_count: NULL pointer");
    FPCC_REQUIRE(n > 0 && n_streams > 0 && n_tables > 0, "indexed_count: bad sizes");
    Tables t = {cdf_flat, cdf_off, cdf_len, offsets, n_tables};
    int64_t total = n * n_streams;
    indexed_count_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(t, overflow, symbols, indexes, n, total, entries_per_symbol);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_indexed_ranges(const uint32_t *cdf_flat, const int64_t *cdf_off, const int32_t *cdf_len, int n_tables,
                                   const int32_t *offsets, int overflow, const int32_t *symbols, const int32_t *indexes,
                                   int64_t n, int n_streams, const int64_t *entry_pos, uint32_t *ranges, uint8_t *bits,
                                   void *stream) {
    FPCC_REQUIRE(cdf_flat && cdf_off && cdf_len && offsets && symbols && entry_pos && ranges && bits, "indexed_ranges: NULL pointer");
    FPCC_REQUIRE(n > 0 && n_streams > 0 && n_tables > 0, "indexed_ranges: bad sizes");
    Tables t = {cdf_flat, cdf_off, cdf_len, offsets, n_tables};
    int64_t total = n * n_streams;
    indexed_ranges_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(t, overflow, symbols, indexes, n, total, entry_pos, ranges, bits);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_indexed_decode(const uint32_t *cdf_flat, const int64_t *cdf_off, const int32_t *cdf_len, int n_tables,
                                   const int32_t *offsets, int overflow, const uint8_t *bytes, const int64_t *byte_off,
                                   const int32_t *byte_len, const int32_t *indexes, int64_t n, int n_streams,
                                   int32_t *symbols, int32_t *err_dev, void *stream) {
    FPCC_REQUIRE(cdf_flat && cdf_off && cdf_len && offsets && bytes && byte_off && byte_len && symbols && err_dev,
                 "indexed_decode: NULL pointer");
    FPCC_REQUIRE(n > 0 && n_streams > 0 && n_tables > 0, "indexed_decode: bad sizes");
    Tables t = {cdf_flat, cdf_off, cdf_len, offsets, n_tables};
    indexed_decode_kernel<<<n_streams, 32, 0, (cudaStream_t)stream>>>(t, overflow, bytes, byte_off, byte_len, indexes, n, symbols, err_dev);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

extern "C" int fpcc_pmf_to_quantized_cdf(double *pmf, int n_tables, int pmf_size, int32_t *offsets, int overflow,
                                         uint32_t *cdf_out, int32_t *cdf_len, void *stream) {
    FPCC_REQUIRE(pmf && offsets && cdf_out && cdf_len, "pmf_to_quantized_cdf: NULL pointer");
    FPCC_REQUIRE(n_tables > 0 && pmf_size >= 2, "pmf_to_quantized_cdf: need at least one table of >= 2 symbols");
    pmf_to_cdf_kernel<<<ceil_div(n_tables, 64), 64, 0, (cudaStream_t)stream>>>(pmf, n_tables, pmf_size, offsets, overflow, cdf_out, cdf_len);
    FPCC_LAUNCH_CHECK();
    return FPCC_OK;
}

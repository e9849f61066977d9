/*
 * oracle/rans_oracle.c -- CPU restatement (plain C) of the reference's range coders.
 *
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.  The product (fastpcc_b200) never
 * links, loads or calls anything in this file.
 *
 * Parity status: PINNED.  tests/test_oracle_rans.py checks every function here against
 *   (a) the reference's own C++ coders compiled from /root/reference into oracle/_ref
 *       (oracle/build_ref.py) on randomised inputs, and
 *   (b) the committed known-answer vectors in tests/golden/rans_kat.json (SURVEY.md 8c,
 *       KAT-A..I, minted from the compiled reference by tests/golden/make_rans_kat.py).
 *
 * What is restated (reference file:line):
 *   32-bit byte-wise rANS primitives ........ lib/entropy_models/rans_coder/rans_byte.h:66-165
 *   fast-division symbol tables .............. rans_byte.h:190-296 (results identical to the
 *                                              plain put/advance; restated with plain division)
 *   RansEncoder::encode / encode_bin / flush . models/convolutional/lossy_coord_v3/rans_coder/
 *                                              simple_rans_wrapper.cpp:67-95, 97-124, 126-134
 *   RansDecoder::flush / decode / decode_bin . simple_rans_wrapper.cpp:139-145, 206-239, 241-270
 *   IndexedRansCoder encode/decode (+overflow
 *   Elias-gamma escape, +index arrays) ........ lib/entropy_models/rans_coder/rans_wrapper.cpp:89-185, 206-279
 *   BinaryRansCoder encode/decode ............ rans_wrapper.cpp:326-382, 385-428
 *   pmf_to_quantized_cdf ...................... lib/entropy_models/rans_coder/cdf_ops.cpp:4-109
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_L (1u << 23)       /* lower bound of the normalisation interval (rans_byte.h:66) */
#define FO_PREC 16u
#define FO_SCALE (1u << 16)

/* ---- primitives: rans_byte.h:77-165 ------------------------------------------------ */

/* one encoder step; writes bytes backwards through *pp */
static inline uint32_t fo_put(uint32_t x, uint8_t **pp, uint32_t start, uint32_t freq, uint32_t bits)
{
    uint32_t x_max = ((FO_L >> bits) << 8) * freq;
    uint8_t *p = *pp;
    while (x >= x_max) {
        *--p = (uint8_t)(x & 0xff);
        x >>= 8;
    }
    *pp = p;
    return ((x / freq) << bits) + (x % freq) + start;
}

static inline void fo_flush_state(uint32_t x, uint8_t **pp)
{
    uint8_t *p = *pp - 4;
    p[0] = (uint8_t)x; p[1] = (uint8_t)(x >> 8); p[2] = (uint8_t)(x >> 16); p[3] = (uint8_t)(x >> 24);
    *pp = p;
}

static inline uint32_t fo_dec_init(const uint8_t **pp)
{
    const uint8_t *p = *pp;
    uint32_t x = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    *pp = p + 4;
    return x;
}

static inline uint32_t fo_advance(uint32_t x, const uint8_t **pp, uint32_t start, uint32_t freq, uint32_t bits)
{
    uint32_t mask = (1u << bits) - 1;
    x = freq * (x >> bits) + (x & mask) - start;
    const uint8_t *p = *pp;
    while (x < FO_L) x = (x << 8) | *p++;
    *pp = p;
    return x;
}

/* ---- "simple" single-stream coder: simple_rans_wrapper.cpp ------------------------- */

typedef struct {
    uint32_t x;
    uint8_t *buf;
    size_t cap;
    uint8_t *ptr;
} fo_simple_enc;

fo_simple_enc *fo_simple_enc_new(size_t cap)
{
    fo_simple_enc *e = (fo_simple_enc *)malloc(sizeof(*e));
    e->buf = (uint8_t *)malloc(cap);
    e->cap = cap;
    e->ptr = e->buf + cap;
    e->x = FO_L;
    return e;
}

void fo_simple_enc_free(fo_simple_enc *e)
{
    if (e) { free(e->buf); free(e); }
}

/* simple_rans_wrapper.cpp:67-95.  cdf is [n_cdf, S] uint16 (n_cdf == n or 1); symbols are
 * pushed last -> first; symbol s covers [cdf[s-1] (0 for s==0), cdf[s] (65536 for s==S-1)). */
uint64_t fo_simple_enc_encode(fo_simple_enc *e, const uint16_t *cdf, size_t n_cdf, size_t S,
                              const uint16_t *sym, size_t n)
{
    for (size_t i = n; i-- > 0;) {
        const uint16_t *row = (n_cdf == 1) ? cdf : cdf + i * S;
        uint32_t s = sym[i];
        uint32_t lo = s == 0 ? 0u : row[s - 1];
        uint32_t hi = s == S - 1 ? FO_SCALE : row[s];
        e->x = fo_put(e->x, &e->ptr, lo, hi - lo, FO_PREC);
    }
    return (uint64_t)(e->buf + e->cap - e->ptr);
}

/* simple_rans_wrapper.cpp:97-124.  cdf[i] = P(0)*65536 threshold: sym 0 -> [0,c), 1 -> [c,65536) */
uint64_t fo_simple_enc_encode_bin(fo_simple_enc *e, const uint16_t *cdf, size_t n_cdf,
                                  const uint8_t *sym, size_t n)
{
    for (size_t i = n; i-- > 0;) {
        uint32_t c = (n_cdf == 1) ? cdf[0] : cdf[i];
        uint32_t lo = sym[i] ? c : 0u;
        uint32_t hi = sym[i] ? FO_SCALE : c;
        e->x = fo_put(e->x, &e->ptr, lo, hi - lo, FO_PREC);
    }
    return (uint64_t)(e->buf + e->cap - e->ptr);
}

/* simple_rans_wrapper.cpp:126-134: 4-byte LE state header, then reset. Returns byte count. */
size_t fo_simple_enc_flush(fo_simple_enc *e, uint8_t *out, size_t out_cap)
{
    fo_flush_state(e->x, &e->ptr);
    size_t len = (size_t)(e->buf + e->cap - e->ptr);
    if (out && len <= out_cap) memcpy(out, e->ptr, len);
    e->ptr = e->buf + e->cap;
    e->x = FO_L;
    return len;
}

typedef struct {
    uint32_t x;
    const uint8_t *ptr;
} fo_simple_dec;

fo_simple_dec *fo_simple_dec_new(void) { return (fo_simple_dec *)calloc(1, sizeof(fo_simple_dec)); }
void fo_simple_dec_free(fo_simple_dec *d) { free(d); }

/* simple_rans_wrapper.cpp:139-145 (borrowed pointer: caller keeps `bytes` alive) */
void fo_simple_dec_flush(fo_simple_dec *d, const uint8_t *bytes)
{
    d->ptr = bytes;
    d->x = fo_dec_init(&d->ptr);
}

/* std::upper_bound over one uint16 row, clamped to S-1 (simple_rans_wrapper.cpp:225-228) */
static inline uint32_t fo_row_search(const uint16_t *row, size_t S, uint32_t cf)
{
    size_t lo = 0, hi = S;
    while (lo < hi) {
        size_t mid = lo + ((hi - lo) >> 1);
        if ((uint32_t)row[mid] <= cf) lo = mid + 1; else hi = mid;
    }
    if (lo > S - 1) lo = S - 1;
    return (uint32_t)lo;
}

/* simple_rans_wrapper.cpp:206-239 */
void fo_simple_dec_decode(fo_simple_dec *d, const uint16_t *cdf, size_t n_cdf, size_t S,
                          uint16_t *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        const uint16_t *row = (n_cdf == 1) ? cdf : cdf + i * S;
        uint32_t cf = d->x & (FO_SCALE - 1);
        uint32_t s = fo_row_search(row, S, cf);
        uint32_t lo = s == 0 ? 0u : row[s - 1];
        uint32_t hi = s == S - 1 ? FO_SCALE : row[s];
        d->x = fo_advance(d->x, &d->ptr, lo, hi - lo, FO_PREC);
        out[i] = (uint16_t)s;
    }
}

/* simple_rans_wrapper.cpp:241-270 */
void fo_simple_dec_decode_bin(fo_simple_dec *d, const uint16_t *cdf, size_t n_cdf, uint8_t *out, size_t n)
{
    for (size_t i = 0; i < n; ++i) {
        uint32_t c = (n_cdf == 1) ? cdf[0] : cdf[i];
        uint32_t cf = d->x & (FO_SCALE - 1);
        uint32_t s = cf >= c;
        uint32_t lo = s ? c : 0u;
        uint32_t hi = s ? FO_SCALE : c;
        d->x = fo_advance(d->x, &d->ptr, lo, hi - lo, FO_PREC);
        out[i] = (uint8_t)s;
    }
}

/* ---- IndexedRansCoder: rans_wrapper.cpp:89-279 ------------------------------------- */
/* Tables are passed flat: cdf_flat holds T concatenated CDFs (each starts with 0 and ends with
 * 65536), cdf_off[t] is the start of table t, cdf_len[t] its entry count (symbols = len-1).     */

/* Encodes one stream. Output is written backwards from buf+cap; returns the start pointer offset
 * (bytes are buf[ret .. cap)).  Returns (size_t)-1 when the buffer is too small.               */
size_t fo_indexed_encode(const uint32_t *cdf_flat, const int64_t *cdf_off, const int32_t *cdf_len,
                         size_t n_tables, const int32_t *offsets, int overflow,
                         const int32_t *sym, const int32_t *idx /* may be NULL */, size_t n,
                         uint8_t *buf, size_t cap)
{
    uint8_t *p = buf + cap;
    uint8_t *guard = buf + 64;
    uint32_t x = FO_L;
    for (size_t f = 0; f < n; ++f) {
        size_t i = n - 1 - f;
        size_t t = idx ? (size_t)idx[i] : i % n_tables;
        const uint32_t *cdf = cdf_flat + cdf_off[t];
        int32_t nsym = cdf_len[t] - 1;
        int32_t v = sym[i] - offsets[t];
        if (p < guard) return (size_t)-1;
        if (overflow) {
            int32_t sign = v < 0;
            int32_t maxv = nsym - 1;
            int32_t gamma = 0;
            if (sign) { gamma = -v; v = maxv; }
            else if (v >= maxv) { gamma = v - maxv + 1; v = maxv; }
            if (v == maxv) {
                /* sign bit, then gamma bits LSB first, then (bit_count-1) zeros; each bit is coded
                 * at precision 1 with freq 1 (rans_wrapper.cpp:153-167)                           */
                x = fo_put(x, &p, (uint32_t)sign, 1, 1);
                int32_t nb = 0;
                while (gamma != 0) { x = fo_put(x, &p, (uint32_t)(gamma & 1), 1, 1); gamma >>= 1; ++nb; }
                while (--nb > 0) x = fo_put(x, &p, 0, 1, 1);
            }
        }
        x = fo_put(x, &p, cdf[v], cdf[v + 1] - cdf[v], FO_PREC);
    }
    fo_flush_state(x, &p);
    return (size_t)(p - buf);
}

void fo_indexed_decode(const uint32_t *cdf_flat, const int64_t *cdf_off, const int32_t *cdf_len,
                       size_t n_tables, const int32_t *offsets, int overflow,
                       const uint8_t *bytes, const int32_t *idx /* may be NULL */, int32_t *out, size_t n)
{
    const uint8_t *p = bytes;
    uint32_t x = fo_dec_init(&p);
    for (size_t j = 0; j < n; ++j) {
        size_t t = idx ? (size_t)idx[j] : j % n_tables;
        const uint32_t *cdf = cdf_flat + cdf_off[t];
        int32_t len = cdf_len[t];
        uint32_t cf = x & (FO_SCALE - 1);
        /* upper_bound(cdf+1, cdf+len, cf) - cdf - 1 (rans_wrapper.cpp:243) */
        int32_t lo = 1, hi = len;
        while (lo < hi) {
            int32_t mid = lo + ((hi - lo) >> 1);
            if (cdf[mid] <= cf) lo = mid + 1; else hi = mid;
        }
        int32_t v = lo - 1;
        x = fo_advance(x, &p, cdf[v], cdf[v + 1] - cdf[v], FO_PREC);
        if (overflow) {
            int32_t maxv = len - 2;
            if (v == maxv) {
                int32_t nb = 0;
                while ((x & 1u) == 0) { ++nb; x = fo_advance(x, &p, 0, 1, 1); }
                x = fo_advance(x, &p, 1, 1, 1);
                v = 1 << nb;
                while (--nb >= 0) {
                    int32_t bit = (int32_t)(x & 1u);
                    x = fo_advance(x, &p, (uint32_t)bit, 1, 1);
                    v |= bit << nb;
                }
                int32_t sign = (int32_t)(x & 1u);
                x = fo_advance(x, &p, (uint32_t)sign, 1, 1);
                v = sign ? -v : v + maxv - 1;
            }
        }
        out[j] = v + offsets[t];
    }
}

/* ---- BinaryRansCoder: rans_wrapper.cpp:326-428.  prob = P(1)*65536 in [1,65535] ------ */

size_t fo_binary_encode(const uint8_t *sym, const uint32_t *prob, size_t n, uint8_t *buf, size_t cap)
{
    uint8_t *p = buf + cap;
    uint8_t *guard = buf + 16;
    uint32_t x = FO_L;
    for (size_t f = 0; f < n; ++f) {
        size_t i = n - 1 - f;
        if (p < guard) return (size_t)-1;
        if (sym[i] == 0) x = fo_put(x, &p, 0, FO_SCALE - prob[i], FO_PREC);
        else x = fo_put(x, &p, FO_SCALE - prob[i], prob[i], FO_PREC);
    }
    fo_flush_state(x, &p);
    return (size_t)(p - buf);
}

void fo_binary_decode(const uint8_t *bytes, const uint32_t *prob, uint8_t *out, size_t n)
{
    const uint8_t *p = bytes;
    uint32_t x = fo_dec_init(&p);
    for (size_t j = 0; j < n; ++j) {
        uint32_t thr = FO_SCALE - prob[j];
        if ((x & (FO_SCALE - 1)) < thr) { out[j] = 0; x = fo_advance(x, &p, 0, thr, FO_PREC); }
        else { out[j] = 1; x = fo_advance(x, &p, thr, prob[j], FO_PREC); }
    }
}

/* ---- pmf_to_quantized_cdf: cdf_ops.cpp:4-109 ----------------------------------------- */
/* pmf is modified in place (prefix sums) exactly as the reference does.  cdf_out must hold
 * pmf_size+2 entries.  Returns the CDF length; *offset is adjusted in overflow mode.           */
int32_t fo_pmf_to_quantized_cdf(double *pmf, size_t pmf_size, int32_t *offset, int overflow, uint32_t *cdf_out)
{
    size_t len = overflow ? pmf_size + 2 : pmf_size + 1;
    uint32_t *cdf = cdf_out;
    for (size_t i = 0; i < len; ++i) cdf[i] = 0;
    double total = 0.;
    for (size_t i = 0; i < pmf_size; ++i) total += pmf[i];            /* std::accumulate, left to right */
    double over = 1. - total > 0. ? 1. - total : 0.;
    if (overflow) total += over;
    for (size_t i = 1; i < pmf_size; ++i) pmf[i] = pmf[i - 1] + pmf[i]; /* std::partial_sum */
    for (size_t i = 0; i < pmf_size; ++i)
        cdf[i + 1] = (uint32_t)round((double)FO_SCALE * (pmf[i] / total));
    cdf[len - 1] = FO_SCALE;

    if (overflow) {
        size_t s = 0, e = 0;
        for (size_t i = 0; i < len - 1; i++) if (cdf[i + 1] != cdf[i]) { s = i; break; }
        for (size_t i = len - 2; i > 0; i--) if (cdf[i - 1] != cdf[i]) { e = i; break; }
        offset[0] += (int32_t)s;
        if (s > e) {
            /* reference line 53 is `assert(cdf_start = cdf.size() - 2)` -- an ASSIGNMENT inside a live
             * assert (NDEBUG is off); its value is overwritten on the next line, so it has no effect. */
            s = len - 3;
            e = s + 1;
        }
        size_t nlen = e - s + 1 + 1;
        for (size_t i = 0; i < nlen - 1; i++) cdf[i] = cdf[i + s];
        len = nlen;
        cdf[len - 1] = FO_SCALE;
    }

    for (size_t i = 0; i < len - 1; i++) {
        if (cdf[i + 1] == cdf[i]) {
            uint32_t best = ~0u;
            size_t steal = 0;
            int found = 0;
            for (size_t j = 0; j < len - 1; j++) {
                uint32_t f = cdf[j + 1] - cdf[j];
                if (f > 1 && f < best) { best = f; steal = j; found = 1; }
            }
            if (!found) return -1;
            if (steal < i) { for (size_t j = steal + 1; j <= i; j++) cdf[j]--; }
            else { for (size_t j = i + 1; j <= steal; j++) cdf[j]++; }
        }
    }
    return (int32_t)len;
}

/* OpenMP C kernels behind oracle/int_ops_fast.py: the element-wise epilogues and the CDF head of the reference path as
 * a competent CPU implementation would run them (one fused pass per operator over all host cores), for the TIMED CPU
 * arm of bench.py.  TEST INFRASTRUCTURE ONLY; never linked into the product.
 *
 * Restates (same integer results as oracle/int_ops.py, pinned by tests/test_oracle_fast_backend.py):
 *   fo_requant        lib/int_sparse_conv/src/element_wise/{requant,bias_requant,prelu_requant,bias_prelu_requant}.cu
 *   fo_prelu          .../element_wise/prelu.cu:6-21
 *   fo_quantize_pmf   src/softmax.cu:41-106 + models/convolutional/lossl_coord_int/model.py:344-353
 *   fo_f32_to_i32_add exact float accumulators of a BLAS GEMM back to int32 (+ optional int32 addend)
 */
#include <stdint.h>
#include <stddef.h>

static inline int64_t rha(int64_t x, int shift) { /* requant.cu:16-20 */
    if (shift <= 0) return x;
    const int64_t half = (int64_t)1 << (shift - 1);
    return x >= 0 ? (x + half) >> shift : -((-x + half) >> shift);
}

static inline int64_t prelu_q25(int64_t v, int32_t slope) { /* bias_prelu_requant.cu:17-22 */
    return v < 0 ? rha(v * (int64_t)slope, 25) : v;
}

void fo_requant(const int32_t *in, int64_t rows, int ch, const int32_t *bias, int has_slope, int32_t slope,
                const uint32_t *mul, int64_t zp, int shift, int out_type, void *out) {
    const int64_t lo = out_type == 0 ? -128 : (out_type == 1 ? -32768 : -2147483648LL);
    const int64_t hi = out_type == 0 ? 127 : (out_type == 1 ? 32767 : 2147483647LL);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        const int32_t *x = in + r * ch;
        for (int c = 0; c < ch; ++c) {
            int64_t v = (int64_t)x[c] + (bias ? (int64_t)bias[c] : 0);
            if (has_slope) v = prelu_q25(v, slope);
            int64_t q = rha(v * (int64_t)mul[c] + zp, shift);
            q = q < lo ? lo : (q > hi ? hi : q);
            if (out_type == 0) ((int8_t *)out)[r * ch + c] = (int8_t)q;
            else if (out_type == 1) ((int16_t *)out)[r * ch + c] = (int16_t)q;
            else ((int32_t *)out)[r * ch + c] = (int32_t)q;
        }
    }
}

void fo_prelu(const int32_t *in, int64_t total, int32_t slope, int32_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < total; ++i) {
        int64_t v = prelu_q25((int64_t)in[i], slope);
        out[i] = (int32_t)(v < -2147483648LL ? -2147483648LL : (v > 2147483647LL ? 2147483647LL : v));
    }
}

/* logits int32 Q8.23 [n, S] (row pitch ld) -> uint16 inclusive CDF [n, S], last entry 65535 */
void fo_quantize_pmf(const int32_t *logits, int64_t n, int S, int64_t ld, const int32_t *lut, int lut_size, uint16_t *cdf) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const int32_t *x = logits + r * ld;
        uint16_t *o = cdf + r * S;
        int64_t mx = (int64_t)(x[0] >> 7);
        for (int c = 1; c < S; ++c) { const int64_t v = (int64_t)(x[c] >> 7); mx = v > mx ? v : mx; }
        mx += 64; /* "+ (1 << 6) for rounding", softmax.cu:71 */
        int64_t sum = 0;
        for (int c = 0; c < S; ++c) {
            int64_t idx = (mx - (int64_t)(x[c] >> 7)) >> 7;
            if (idx > lut_size - 1) idx = lut_size - 1;
            sum += lut[idx];
        }
        const uint64_t inv = sum > 0 ? (uint64_t)((((int64_t)1 << 32) + (sum >> 1)) / sum) : (uint64_t)(((int64_t)1 << 32) / S);
        uint64_t acc = 0;
        for (int c = 0; c < S; ++c) {
            int64_t idx = (mx - (int64_t)(x[c] >> 7)) >> 7;
            if (idx > lut_size - 1) idx = lut_size - 1;
            uint64_t p = (uint64_t)lut[idx] * inv;
            if (p > 0xFFFFFFFFull) p = 0xFFFFFFFFull;
            acc += ((p * (uint64_t)(65536 - S)) >> 32) + 1;
            o[c] = (uint16_t)acc;
        }
        o[S - 1] = 65535;
    }
}

/* out[i] = (int32) acc[i] (+ add[i % ch_add] when add_mode == 1, + add[i] when add_mode == 2); returns max |value| */
double fo_f32_to_i32_add(const float *acc, int64_t rows, int ch, const int32_t *add, int add_mode, int32_t *out) {
    double worst = 0.0;
#pragma omp parallel for schedule(static) reduction(max : worst)
    for (int64_t r = 0; r < rows; ++r)
        for (int c = 0; c < ch; ++c) {
            double v = (double)acc[r * ch + c];
            if (add_mode == 1) v += (double)add[c];
            else if (add_mode == 2) v += (double)add[r * ch + c];
            const double a = v < 0 ? -v : v;
            worst = a > worst ? a : worst;
            out[r * ch + c] = (int32_t)v;
        }
    return worst;
}

/* D[scatter[i], :] += (int32) upd[i, :]   (scatter indices unique within one call) */
void fo_scatter_add_f32(const float *upd, int64_t L, int ch, const int32_t *scatter, int32_t *D) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L; ++i) {
        int32_t *d = D + (int64_t)scatter[i] * ch;
        const float *u = upd + i * ch;
        for (int c = 0; c < ch; ++c) d[c] += (int32_t)u[c];
    }
}

/* dst[i, :] = (float) src[gather[i], :]  (gather == NULL: identity) */
void fo_gather_i8_to_f32(const int8_t *src, int ch, const int32_t *gather, int64_t L, float *dst) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < L; ++i) {
        const int8_t *s = src + (int64_t)(gather ? gather[i] : i) * ch;
        float *d = dst + i * ch;
        for (int c = 0; c < ch; ++c) d[c] = (float)s[c];
    }
}

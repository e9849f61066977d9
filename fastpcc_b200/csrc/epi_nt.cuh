// Rounding-tie analysis of the integer requantisation (host + device; also compiled by the CPU test tests/epi_nt_host.cpp).
//
// The reference rounds half away from zero:  rha(r, s) = (r + 2^(s-1) - [r < 0]) >> s,  r = v * mul + zp
// (requant.cu:16-20, bias_prelu_requant.cu:24-33).  The "- [r < 0]" term changes the result only at an exact tie,
// r == 2^(s-1) (mod 2^s).  For the shifts of a PTQ-converted model (28 ... 48) and |v| bounded by the static accumulator
// bound, most channels cannot produce a tie at all: v * mul == 2^(s-1) - zp (mod 2^s) has its solutions spaced
// 2^(s - ctz(mul)) apart and usually none of them lies in [-A, A].  For such a channel
//     rha(v * mul + zp, s) == (v * mul + zp + 2^(s-1)) >> s        for every |v| <= A,
// i.e. ONE multiply-add with a per-channel 64-bit constant and one shift, no sign handling.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FPCC_HD __host__ __device__ __forceinline__
#else
#define FPCC_HD static inline
#endif

namespace fpcc {

FPCC_HD int ctz_u32(uint32_t m) {  // m != 0
#ifdef __CUDA_ARCH__
    return __ffs((int)m) - 1;
#else
    return __builtin_ctz(m);
#endif
}

// true when some |v| <= A may make v * mul + zp an exact rounding tie of the shift s (1 <= s <= 62); false PROVES that no
// such v exists.  (The sign of r at the tie is ignored: conservative.)
FPCC_HD bool tie_possible(uint32_t mul, int64_t zp, int s, int64_t A) {
    if (mul == 0u || s < 1 || s > 62 || A < 0) return true;
    const uint64_t mask = ((uint64_t)1 << s) - 1;
    const uint64_t d = (((uint64_t)1 << (s - 1)) - (uint64_t)zp) & mask;  // v * mul == d (mod 2^s)
    const int tz = ctz_u32(mul);
    if (tz >= s) return d == 0;  // v * mul == 0 (mod 2^s) for every v
    if (d & (((uint64_t)1 << tz) - 1)) return false;  // the left side is a multiple of 2^tz, the right side is not
    const uint64_t m1 = (uint64_t)(mul >> tz);  // odd
    const int s1 = s - tz;
    const uint64_t mask1 = ((uint64_t)1 << s1) - 1;
    uint64_t x = m1;  // Newton iteration for m1^-1 mod 2^64: correct to 3 bits, doubled five times
    x *= 2 - m1 * x; x *= 2 - m1 * x; x *= 2 - m1 * x; x *= 2 - m1 * x; x *= 2 - m1 * x;
    const uint64_t v0 = (x * (d >> tz)) & mask1;  // the solutions are v == v0 (mod 2^s1), 0 <= v0 < 2^s1
    const uint64_t a = (uint64_t)A;
    return v0 <= a || (((uint64_t)1 << s1) - v0) <= a;
}

}  // namespace fpcc

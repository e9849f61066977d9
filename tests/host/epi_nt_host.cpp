// CPU check of fastpcc_b200/csrc/epi_nt.cuh (the tie analysis behind the "no-tie" integer epilogue): compiled and run
// by tests/test_epi_nt_host.py.  Exhaustive on small shifts, randomised + constructed ties on the shifts of converted
// models.  Reference arithmetic: requant.cu:16-20 (round half away from zero).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../fastpcc_b200/csrc/epi_nt.cuh"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static int64_t rha(int64_t r, int s) { return (r + (((int64_t)1 << (s - 1)) - (int64_t)((uint64_t)r >> 63))) >> s; }
static int64_t nt(int64_t r, int s) { return (r + ((int64_t)1 << (s - 1))) >> s; }

int main() {
    long checked = 0, proven = 0;
    // 1. exhaustive: small shifts and bounds, every v
    for (int it = 0; it < 200000; ++it) {
        const int s = 1 + (int)(rnd() % 20);
        uint32_t mul = (uint32_t)(rnd() % (1u << (1 + rnd() % 16))) << (rnd() % 8);
        if (mul == 0) mul = 1;
        const int64_t zp = (rnd() & 1) ? 0 : (int64_t)(rnd() % (1ull << 22)) - (1 << 21);
        const int64_t A = (int64_t)(rnd() % 600);
        const bool poss = fpcc::tie_possible(mul, zp, s, A);
        bool found = false;
        for (int64_t v = -A; v <= A; ++v) {
            const int64_t r = v * (int64_t)mul + zp;
            const uint64_t M = (uint64_t)1 << s;
            if ((((uint64_t)r) & (M - 1)) == (M >> 1)) found = true;
            if (!poss && rha(r, s) != nt(r, s)) { printf("FAIL formula s=%d mul=%u zp=%lld v=%lld\n", s, mul, (long long)zp, (long long)v); return 1; }
        }
        if (found && !poss) { printf("FAIL missed tie s=%d mul=%u zp=%lld A=%lld\n", s, mul, (long long)zp, (long long)A); return 1; }
        if (!found && poss) {
            // allowed to be conservative only through the two-candidate test; with exhaustive v it must be exact here
            printf("FAIL conservative s=%d mul=%u zp=%lld A=%lld\n", s, mul, (long long)zp, (long long)A); return 1;
        }
        ++checked; proven += !poss;
    }
    // 2. converted-model regime: shifts 24..62, multipliers < 2^31, bounds up to 2^31; sampled v incl. the candidate ties
    for (int it = 0; it < 400000; ++it) {
        const int s = 24 + (int)(rnd() % 39);
        uint32_t mul = (uint32_t)(rnd() & 0x7fffffffu);
        if (rnd() % 4 == 0) mul &= ~((1u << (rnd() % 24)) - 1);  // trailing zeros: ties become reachable
        if (mul == 0) mul = 1u << 20;
        int64_t zp = 0;
        if (rnd() % 3 == 0) zp = (int64_t)(rnd() % (1ull << 40)) - ((int64_t)1 << 39);
        const int64_t A = (int64_t)(rnd() % (1ull << (10 + rnd() % 22)));
        const bool poss = fpcc::tie_possible(mul, zp, s, A);
        // candidate tie values, recomputed independently by 128-bit search over the residue class
        const int tz = __builtin_ctz(mul);
        const unsigned __int128 M = (unsigned __int128)1 << s;
        const uint64_t d = (uint64_t)((((unsigned __int128)1 << (s - 1)) - (unsigned __int128)(__int128)zp) & (M - 1));
        bool found = false;
        int64_t tie_v = 0;
        if (tz >= s) { found = d == 0; tie_v = -1 <= A ? -(A > 0) : 0; }
        else if ((d & (((uint64_t)1 << tz) - 1)) == 0) {
            const int s1 = s - tz;
            const uint64_t m1 = mul >> tz, mask1 = ((uint64_t)1 << s1) - 1, d1 = d >> tz;
            uint64_t inv = 1;  // bit-by-bit inverse (independent of the Newton form in the header)
            for (int b = 1; b < 64; ++b) if (((m1 * inv) >> b) & 1) inv |= (uint64_t)1 << b;
            const uint64_t v0 = (inv * d1) & mask1;
            if (v0 <= (uint64_t)A) { found = true; tie_v = (int64_t)v0; }
            if (((uint64_t)1 << s1) - v0 <= (uint64_t)A) { found = true; tie_v = (int64_t)v0 - ((int64_t)1 << s1); }
        }
        if (found != poss) { printf("FAIL regime s=%d mul=%u zp=%lld A=%lld found=%d poss=%d\n", s, mul, (long long)zp, (long long)A, found, poss); return 1; }
        if (found && tz < s) {  // the candidate really is a tie
            const __int128 r = (__int128)tie_v * mul + zp;
            if ((uint64_t)((unsigned __int128)r & (M - 1)) != (uint64_t)(M >> 1)) { printf("FAIL candidate not a tie\n"); return 1; }
        }
        if (!poss) {
            for (int k = 0; k < 64; ++k) {
                int64_t v = k == 0 ? -A : (k == 1 ? A : (k == 2 ? 0 : (int64_t)(rnd() % (2 * (uint64_t)A + 1)) - A));
                const int64_t r = v * (int64_t)mul + zp;  // |v * mul| < 2^62
                if (rha(r, s) != nt(r, s)) { printf("FAIL formula2 s=%d mul=%u zp=%lld v=%lld\n", s, mul, (long long)zp, (long long)v); return 1; }
            }
            ++proven;
        }
        ++checked;
    }
    printf("ok %ld cases, %ld proven tie-free\n", checked, proven);
    return 0;
}

// Fixed-point arithmetic of the int8 digit-plane path (i8path.cu, verify.cu), shared between device code and the host-side
// known-answer test (tests/test_i8digits_cpu.py compiles this header with g++).
//
// A row of G_s = L' diag(e_s) is scaled by a power of two 2^E >= max_k |G_s[j,k]| / I8_FILL, rounded to nearest to I8_FRAC
// fractional bits (w = rn(G 2^(I8_FRAC - E))) and written as I8_DIG BALANCED radix-256 digits g_i in [-128, 127]:
// w = sum_i g_i 256^(I8_DIG - 1 - i). The digit sums S_i = sum_k g_i[k] xi[k] over n non-zero ternary mask entries obey
// |S_i| <= 128 n and are exact in int32; W = sum_i S_i 256^(I8_DIG - 1 - i) = sum_k w[k] xi[k] is exact in 64 bits while
// n < 2^14 (|W| <= n 2^I8_FRAC).
//   GMS_I8_DIG = 7 (default): 49 fractional bits, the precision of round 1's seven sign-magnitude radix-128 planes (the top
//                plane only holds -2 .. 2).
//   GMS_I8_DIG = 6 (build variant lib/dig6, tools only): 47 fractional bits in six full-range planes (largest six-digit value
//                127 (256^6 - 1) / 255 = 0.99608 * 2^47, hence I8_FILL = 0.996). 14 % fewer MMAs, operand bytes and TMEM columns
//                (32.3 ms against 37.3 ms per 50k cassava crosses) for 4 x the rounding error: 4e-13 instead of 1e-13 relative on
//                the variances, both far inside the 1e-9 contract that verify.cu enforces unit by unit - but a covariance that
//                cancels to 1e-5 of its variance scale carries the error relative to that scale, and with six planes one of the
//                90,000 entries of the T = 24 case of tests/test_gpu_parity.py (int8 mode against the CPU restatement) lands at
//                1.8e-9 of its own (guarded) size. Parity is the gate, so seven planes stay the product.
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define GMS_HD __host__ __device__ __forceinline__
#else
#define GMS_HD inline
#endif

namespace gms {

#ifndef GMS_I8_DIG
#define GMS_I8_DIG 7
#endif
constexpr int I8_DIG = GMS_I8_DIG;                          // balanced radix-256 digit planes
static_assert(I8_DIG == 6 || I8_DIG == 7, "GMS_I8_DIG must be 6 or 7");
constexpr int I8_FRAC = I8_DIG == 6 ? 47 : 49;              // fractional bits below the row scale
constexpr long long I8_WMAX = I8_DIG == 6 ? 127LL * 0x010101010101LL : (1LL << 49);     // largest |w| (6: what six digits hold)
constexpr double I8_FILL = I8_DIG == 6 ? 0.996 : 1.0;       // rowmax <= I8_FILL * scale
// error of one stored entry relative to its row scale: 2^-(I8_FRAC + 1) from rounding to nearest, plus 2^-52 because the
// epilogue truncates a row sum (|sum| <= n_j * scale) to 53 bits when it converts it to FP64 with integer instructions
constexpr double I8_U = (I8_DIG == 6 ? 0x1p-48 : 0x1p-50) + 0x1p-52;

// exponent E of the power-of-two scale of a row whose largest magnitude is v > 0
GMS_HD int i8_scale_exp(double v) {
    int e = 0;
    const double f = frexp(v, &e);                          // v = f 2^e, f in [0.5, 1)
    return f > I8_FILL ? e + 1 : e;                         // I8_FILL = 1: v < 2^e always
}

// fixed-point image of x = G / scale, |x| <= I8_FILL (clamped: a caller that breaks the rule gets a saturated entry, never
// a wrapped one)
GMS_HD long long i8_fixed(double x) {
    const double sc = (double)(1LL << I8_FRAC);
#if defined(__CUDA_ARCH__)
    long long w = __double2ll_rn(x * sc);
#else
    long long w = (long long)nearbyint(x * sc);
#endif
    return w > I8_WMAX ? I8_WMAX : (w < -I8_WMAX ? -I8_WMAX : w);
}

// I8_DIG balanced digits, most significant first
GMS_HD void i8_split(long long w, signed char (&d)[I8_DIG]) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int i = I8_DIG - 1; i >= 0; --i) {
        const int g = (int)((w + 128) & 255) - 128;
        d[i] = (signed char)g;
        w = (w - g) >> 8;                                   // exact: w - g is a multiple of 256
    }
}

// W = sum_i S_i 256^(I8_DIG - 1 - i) as one 64-bit integer, built from 32-bit pairs (|256 S_i + S_j| <= 128 n * 257 < 2^31
// for n < 2^15); the additions wrap modulo 2^64, so the result is exact whenever the true W fits
GMS_HD long long i8_combine(const int (&s)[I8_DIG]) {
    constexpr int o = I8_DIG - 6;
    const int w01 = s[o] * 256 + s[o + 1], w23 = s[o + 2] * 256 + s[o + 3], w45 = s[o + 4] * 256 + s[o + 5];
    unsigned long long w = ((unsigned long long)(long long)w01 << 32) + ((unsigned long long)(long long)w23 << 16) + (unsigned long long)(long long)w45;
    if (I8_DIG == 7) w += (unsigned long long)(long long)s[0] << 48;
    return (long long)w;
}

#if defined(__CUDA_ARCH__)
#define GMS_CLZ(x) __clz((int)(x))
#define GMS_FSL(lo, hi, s) __funnelshift_l((lo), (hi), (s))
#define GMS_FSR(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#else
#define GMS_CLZ(x) ((x) ? __builtin_clz(x) : 32)
#define GMS_FSL(lo, hi, s) ((unsigned)(((((unsigned long long)(hi)) << 32 | (lo)) << ((s) & 31)) >> 32))
#define GMS_FSR(lo, hi, s) ((unsigned)(((((unsigned long long)(hi)) << 32 | (lo)) >> ((s) & 31))))
#endif

// The FP64 bit pattern of  (+-) W 2^(E - I8_FRAC)  built with INTEGER instructions only (on B200 every FP64 instruction of the
// epilogue stalls the tcgen05 MMAs of the same SM - tools/tc_contention_probe.cu - while INT32 work overlaps with them).
// eb = (exponent field of the row scale 2^E) + 63 - I8_FRAC - 1; keep = 0 zeroes the value, flip = 1 negates it (the ternary mask entry
// of the row). |W| >= 2^53 is truncated to 53 bits (relative error < 2^-52, part of I8_U). Row scales are >= 2^-900 and
// W < 2^63, so the exponent field never leaves the normal range.
GMS_HD void i8_cvt_bits(long long w, int eb, unsigned keep, unsigned flip, unsigned& out_hi, unsigned& out_lo) {
    const unsigned sgn = (((unsigned)((unsigned long long)w >> 63)) ^ flip) << 31;
    const unsigned long long a = (unsigned long long)(w < 0 ? -w : w);
    const unsigned hi = (unsigned)(a >> 32), lo = (unsigned)a;
    const unsigned x = hi ? hi : lo, y = hi ? lo : 0u;
    const int lz = GMS_CLZ(x);                              // 32 only if W == 0 (zeroed below)
    const unsigned nh = GMS_FSL(y, x, lz), nl = y << (lz & 31);      // (x:y) << lz: bit 31 of nh is the leading one
    const unsigned mh = nh >> 11, ml = GMS_FSR(nl, nh, 11);          // 53-bit significand, its leading one at bit 20 of mh
    const int e = eb - lz - (hi ? 0 : 32);
    const unsigned bh = (mh + ((unsigned)e << 20)) | sgn;            // the leading one carries into the exponent field
    const bool k = keep != 0u && x != 0u;
    out_hi = k ? bh : 0u;
    out_lo = k ? ml : 0u;
}

}  // namespace gms

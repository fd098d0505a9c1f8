// Fixed-point arithmetic of the int8 digit-plane path (i8path.cu, verify.cu), shared between device code and the host-side
// known-answer test (tests/test_i8digits_cpu.py compiles this header with g++).
//
// A row of G_s = L' diag(e_s) is scaled by a power of two 2^E with max_k |G_s[j,k]| <= 0.996 * 2^E, rounded to nearest
// to I8_FRAC = 47 fractional bits (w = rn(G 2^(47 - E)), |w| <= 0.996 * 2^47) and written as I8_DIG = 6 BALANCED radix-256
// digits g_i in [-128, 127]:  w = sum_i g_i 256^(5 - i).  Six planes of full-range int8 digits carry 48 bits, one more bit
// per plane than sign-magnitude radix-128 digits; the largest six-digit value is 127 (256^6 - 1) / 255 = 0.99608 * 2^47,
// hence the 0.996. The digit sums S_i = sum_k g_i[k] xi[k] over n non-zero ternary mask entries obey |S_i| <= 128 n and
// are exact in int32; W = sum_i S_i 256^(5 - i) is exact in 64 bits while n < 2^15 (|W| <= 1.004 * 2^47 n).
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define GMS_HD __host__ __device__ __forceinline__
#else
#define GMS_HD inline
#endif

namespace gms {

constexpr int I8_DIG = 6;                                   // balanced radix-256 digit planes
constexpr int I8_FRAC = 47;                                 // fractional bits below the row scale
constexpr long long I8_WMAX = 127LL * 0x010101010101LL;     // largest magnitude six balanced digits hold on both sides
constexpr double I8_FILL = 0.996;                           // rowmax <= I8_FILL * scale
// error of one stored entry relative to its row scale: 2^-48 from rounding to 47 fractional bits, plus 2^-52 because the
// epilogue truncates a row sum (|sum| <= n_j * scale) to 53 bits when it converts it to FP64 with integer instructions
constexpr double I8_U = 0x1p-48 * (1.0 + 0x1p-4);

// exponent E of the power-of-two scale of a row whose largest magnitude is v > 0
GMS_HD int i8_scale_exp(double v) {
    int e = 0;
    const double f = frexp(v, &e);                          // v = f 2^e, f in [0.5, 1)
    return f > I8_FILL ? e + 1 : e;
}

// fixed-point image of x = G / scale, |x| <= 0.996 (clamped: a caller that breaks the rule gets a saturated entry, never
// a wrapped one)
GMS_HD long long i8_fixed(double x) {
#if defined(__CUDA_ARCH__)
    long long w = __double2ll_rn(x * 0x1p47);
#else
    long long w = (long long)nearbyint(x * 0x1p47);
#endif
    return w > I8_WMAX ? I8_WMAX : (w < -I8_WMAX ? -I8_WMAX : w);
}

// six balanced digits, most significant first
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

// W = sum_i S_i 256^(5 - i) as one 64-bit integer: three 32-bit pairs (|S_i 256 + S_j| <= 128 n * 257 < 2^31 for n < 2^15)
GMS_HD long long i8_combine(int s0, int s1, int s2, int s3, int s4, int s5) {
    const int w01 = s0 * 256 + s1, w23 = s2 * 256 + s3, w45 = s4 * 256 + s5;
    return (long long)(((unsigned long long)(long long)w01 << 32) + ((unsigned long long)(long long)w23 << 16)) + (long long)w45;
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

// The FP64 bit pattern of  (+-) W 2^(E - 47)  built with INTEGER instructions only (on B200 every FP64 instruction of the
// epilogue stalls the tcgen05 MMAs of the same SM - tools/tc_contention_probe.cu - while INT32 work overlaps with them).
// eb = (exponent field of the row scale 2^E) + 15; keep = 0 zeroes the value, flip = 1 negates it (the ternary mask entry
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

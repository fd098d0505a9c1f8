// Known-answer test of csrc/i8digits.cuh on the host (g++): digit split, digit-sum recombination and the integer-only
// FP64 conversion against exact 128-bit / long double arithmetic. Prints "ok <cases>" or the first failure.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <random>
#include "../../genomicmateselectr_b200/csrc/i8digits.cuh"
using namespace gms;

static int fails = 0;
#define CHECK(c, ...) do { if (!(c)) { if (fails++ < 5) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } } while (0)

int main() {
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(-I8_FILL, I8_FILL);
    long long cases = 0;
    // 1. split: exact reconstruction, digits in range, rounding error <= 2^-48
    for (int it = 0; it < 2000000; ++it) {
        double x = U(rng);
        if (it < 8) { const double edge[8] = {0.0, I8_FILL, -I8_FILL, ldexp(1.0, -I8_FRAC), -ldexp(1.0, -I8_FRAC), ldexp(1.0, -I8_FRAC - 1), 0.5, -0.5}; x = edge[it]; }
        const long long w = i8_fixed(x);
        signed char d[I8_DIG];
        i8_split(w, d);
        long long r = 0;
        for (int i = 0; i < I8_DIG; ++i) r = r * 256 + d[i];
        CHECK(r == w, "split: x=%a w=%lld rebuilt %lld", x, w, r);
        CHECK(fabs(ldexp((double)w, -I8_FRAC) - x) <= ldexp(1.0, -I8_FRAC - 1), "rounding: x=%a", x);
        ++cases;
    }
    if (I8_DIG == 6) {   // saturation instead of wrap-around
        signed char d[I8_DIG];
        i8_split(i8_fixed(1.0), d);
        for (int i = 0; i < I8_DIG; ++i) CHECK(d[i] == 127, "clamp+ digit %d = %d", i, d[i]);
        i8_split(i8_fixed(-1.0), d);
        for (int i = 0; i < I8_DIG; ++i) CHECK(d[i] == -127, "clamp- digit %d = %d", i, d[i]);
    }
    // 2. scale rule: rowmax <= 0.996 scale, scale minimal
    for (int it = 0; it < 200000; ++it) {
        const double v = ldexp(0.5 + 0.5 * (double)(rng() >> 11) * 0x1p-53, (int)(rng() % 1800) - 890);
        const int e = i8_scale_exp(v);
        CHECK(v <= I8_FILL * ldexp(1.0, e) && v > I8_FILL * ldexp(1.0, e - 1) * (1 - 1e-15), "scale: v=%a e=%d", v, e);
        ++cases;
    }
    // 3. digit sums -> W: exact for n < 2^14 terms, worst-case signs included
    for (int it = 0; it < 3000; ++it) {
        const int n = it < 4 ? 16383 : 1 + (int)(rng() % 16383);
        int S[I8_DIG] = {};
        __int128 exact = 0;
        for (int k = 0; k < n; ++k) {
            long long w;
            int xi;
            if (it == 0) { w = I8_WMAX; xi = 1; }
            else if (it == 1) { w = -I8_WMAX; xi = 1; }
            else if (it == 2) { w = -I8_WMAX; xi = -1; }
            else if (it == 3) { w = (k & 1) ? I8_WMAX : -I8_WMAX; xi = (k & 1) ? 1 : -1; }
            else { w = i8_fixed(U(rng)); xi = (int)(rng() % 3) - 1; }
            signed char d[I8_DIG];
            i8_split(w, d);
            for (int i = 0; i < I8_DIG; ++i) S[i] += d[i] * xi;
            exact += (__int128)w * xi;
        }
        const long long W = i8_combine(S);
        CHECK((__int128)W == exact, "combine: n=%d W=%lld", n, W);
        // 4. conversion of this W at a few scales, all four (keep, flip) combinations
        for (int q = 0; q < 8; ++q) {
            const int E = q == 0 ? -900 : (q == 1 ? 900 : (int)(rng() % 1600) - 800);
            const int Ef = E + 1023;
            for (unsigned keep = 0; keep < 2; ++keep)
                for (unsigned flip = 0; flip < 2; ++flip) {
                    unsigned bh, bl;
                    i8_cvt_bits(W, Ef + 63 - I8_FRAC - 1, keep, flip, bh, bl);
                    const uint64_t bits = ((uint64_t)bh << 32) | bl;
                    double got;
                    memcpy(&got, &bits, 8);
                    // truncation towards zero of |W| to 53 bits
                    unsigned long long a = (unsigned long long)(W < 0 ? -W : W);
                    int sh = 0;
                    while ((a >> sh) >= (1ull << 53)) ++sh;
                    const double mag = ldexp((double)((a >> sh) << sh), E - I8_FRAC);
                    double want = (keep && W != 0) ? ((W < 0) != (flip != 0) ? -mag : mag) : 0.0;
                    CHECK(memcmp(&got, &want, 8) == 0 || (got == 0.0 && want == 0.0 && !(keep && W != 0)), "cvt: W=%lld E=%d keep=%u flip=%u got %a want %a", W, E, keep, flip, got, want);
                    if (!keep || W == 0) CHECK(bits == 0, "cvt zero pattern");
                    ++cases;
                }
        }
    }
    // 5. conversion on small and boundary magnitudes
    const long long special[] = {1, -1, 2, 3, (1LL << 31) - 1, 1LL << 31, (1LL << 32) - 1, 1LL << 32, (1LL << 52) + 1, (1LL << 53) - 1, 1LL << 53, (1LL << 53) + 1,
                                 (1LL << 62) - 1, (1LL << 62) + 12345, -(1LL << 62), 0};
    for (long long W : special) {
        unsigned bh, bl;
        i8_cvt_bits(W, 1023 + 63 - I8_FRAC - 1, 1, 0, bh, bl);
        const uint64_t bits = ((uint64_t)bh << 32) | bl;
        double got;
        memcpy(&got, &bits, 8);
        unsigned long long a = (unsigned long long)(W < 0 ? -W : W);
        int sh = 0;
        while ((a >> sh) >= (1ull << 53)) ++sh;
        const double want = (W < 0 ? -1.0 : 1.0) * ldexp((double)((a >> sh) << sh), -I8_FRAC);
        CHECK(got == want, "cvt special W=%lld got %a want %a", W, got, want);
        ++cases;
    }
    if (fails) { printf("FAILED %d checks\n", fails); return 1; }
    printf("ok %lld\n", cases);
    return 0;
}

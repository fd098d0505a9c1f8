// Host check of the int8 path's block ownership (csrc/common.cuh): for every chromosome size the 64-row tiles together own
// every unordered pair of 64-blocks exactly once, the diagonal block comes first, the linear block index and its inverse
// agree, and no tile owns more than one block more than another.
#include "../../genomicmateselectr_b200/csrc/common.cuh"
#include <algorithm>
#include <cstdio>
#include <set>
#include <utility>
using namespace gms;
int main() {
    for (int nb = 1; nb <= 300; ++nb) {
        std::set<std::pair<int, int>> seen;
        long long cnt = 0;
        int lo = 1 << 30, hi = 0;
        for (int J = 0; J < nb; ++J) {
            if (i8_blocks_before(nb, J) != cnt) { std::printf("prefix nb=%d J=%d\n", nb, J); return 1; }
            const int n = i8_nassign(nb, J);
            lo = std::min(lo, n); hi = std::max(hi, n);
            for (int i = 0; i < n; ++i, ++cnt) {
                int J2, i2;
                i8_block_decode(nb, cnt, J2, i2);
                if (J2 != J || i2 != i) { std::printf("decode nb=%d J=%d i=%d\n", nb, J, i); return 1; }
                const int K = (J + i) % nb;
                if (i == 0 && K != J) { std::printf("diagonal nb=%d J=%d\n", nb, J); return 1; }
                if (!seen.insert({std::min(J, K), std::max(J, K)}).second) { std::printf("twice nb=%d %d %d\n", nb, J, K); return 1; }
            }
        }
        if (cnt != i8_blocks_of_chrom(nb) || (long long)seen.size() != cnt || hi - lo > 1) { std::printf("count nb=%d\n", nb); return 1; }
    }
    std::printf("ok 300 chromosome sizes\n");
    return 0;
}

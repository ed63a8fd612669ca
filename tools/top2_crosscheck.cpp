// CPU cross-check of the vocabulary kernel's per-chunk top-2 (vocab_top2_pair_kernel's epilogue, csrc/linear_tc.cu) against a
// brute-force canonical top-2 (value descending, column ascending) on random 32-column chunks with many exact ties and ragged
// tails.  `new_top2` restates the two-level epilogue the kernel runs now; `old_top2` restates the four insertion chains it ran
// before — kept because this harness is how their tie-order defect was found (the runner-up's COLUMN was not canonical when two
// equal runner-up values sat in one chain and the chunk's best arrived later in that chain).
//   g++ -O2 -o top2_crosscheck tools/top2_crosscheck.cpp && ./top2_crosscheck [iterations]
// prints "new_vs_ref <mismatches> old_vs_ref <mismatches>"; tests/test_host_logic.py expects 0 and > 0.
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <algorithm>
static void old_top2(const float* x, float& C1, int& K1, float& C2, int& K2) {
    float c1[4], c2[4]; int k1[4], k2[4];
    for (int u = 0; u < 4; ++u) { c1[u] = -INFINITY; c2[u] = -INFINITY; k1[u] = 0xFFFF; k2[u] = 0xFFFF; }
    for (int j = 0; j < 32; ++j) {
        const int u = j & 3; const float xv = x[j];
        const bool g1 = xv > c1[u];
        const float loser = fminf(c1[u], xv);
        const int kl = g1 ? k1[u] : j;
        c1[u] = fmaxf(c1[u], xv);
        k1[u] = g1 ? j : k1[u];
        const bool g2 = loser > c2[u];
        c2[u] = fmaxf(c2[u], loser);
        k2[u] = g2 ? kl : k2[u];
    }
    auto better = [](float av, int ai, float bv, int bi) { return av > bv || (av == bv && ai < bi); };
    auto merge2 = [&](float& p1, int& i1, float& p2, int& i2, float q1, int j1_, float q2, int j2_) {
        const bool qf = better(q1, j1_, p1, i1);
        const float r1 = qf ? q1 : p1; const int ri1 = qf ? j1_ : i1;
        const float s1 = qf ? p1 : p2; const int si1 = qf ? i1 : i2;
        const float s2 = qf ? q2 : q1; const int si2 = qf ? j2_ : j1_;
        const bool sf = better(s2, si2, s1, si1);
        p1 = r1; i1 = ri1;
        p2 = sf ? s2 : s1; i2 = sf ? si2 : si1;
    };
    merge2(c1[0], k1[0], c2[0], k2[0], c1[1], k1[1], c2[1], k2[1]);
    merge2(c1[2], k1[2], c2[2], k2[2], c1[3], k1[3], c2[3], k2[3]);
    merge2(c1[0], k1[0], c2[0], k2[0], c1[2], k1[2], c2[2], k2[2]);
    C1 = c1[0]; K1 = k1[0]; C2 = c2[0]; K2 = k2[0];
}
static void new_top2(const float* x, float& C1, int& K1, float& C2, int& K2) {
    float gb[4]; int gi[4];
    for (int g = 0; g < 4; ++g) {
        float bv = x[8 * g]; int bi = 8 * g;
        for (int t = 1; t < 8; ++t) { const float xv = x[8 * g + t]; const bool gt = xv > bv; bv = fmaxf(bv, xv); bi = gt ? 8 * g + t : bi; }
        gb[g] = bv; gi[g] = bi;
    }
    float c1v = gb[0]; int k1v = gi[0];
    for (int g = 1; g < 4; ++g) { const bool gt = gb[g] > c1v; c1v = fmaxf(c1v, gb[g]); k1v = gt ? gi[g] : k1v; }
    const int wg = k1v >> 3, wt = k1v & 7;
    float sv = -INFINITY; int st_ = 0;
    for (int t = 0; t < 8; ++t) {
        const float xs = wg == 0 ? x[t] : (wg == 1 ? x[8 + t] : (wg == 2 ? x[16 + t] : x[24 + t]));
        const float xv = t == wt ? -INFINITY : xs;
        const bool gt = xv > sv; sv = fmaxf(sv, xv); st_ = gt ? t : st_;
    }
    float c2v = -INFINITY; int k2v = 0xFFFF;
    for (int g = 0; g < 4; ++g) {
        const float cv = g == wg ? sv : gb[g];
        const int ci = g == wg ? 8 * g + st_ : gi[g];
        const bool gt = cv > c2v; c2v = fmaxf(c2v, cv); k2v = gt ? ci : k2v;
    }
    C1 = c1v; K1 = c1v == -INFINITY ? 0xFFFF : k1v; C2 = c2v; K2 = k2v;
}
static void ref_top2(const float* x, float& C1, int& K1, float& C2, int& K2) {
    C1 = -INFINITY; K1 = 0xFFFF; C2 = -INFINITY; K2 = 0xFFFF;
    for (int j = 0; j < 32; ++j) if (x[j] > C1) { C1 = x[j]; K1 = j; }
    for (int j = 0; j < 32; ++j) if (j != K1 && x[j] > C2) { C2 = x[j]; K2 = j; }
}
int main(int argc, char** argv) {
    srand(1);
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    long bad_new = 0, bad_old = 0;
    for (long it = 0; it < n; ++it) {
        float x[32];
        const int range = 2 + rand() % 40, nv = (it % 7 == 0) ? 1 + rand() % 32 : 32;
        for (int j = 0; j < 32; ++j) x[j] = j < nv ? (float)(rand() % range) - range / 2 : -INFINITY;
        float r1, r2, a1, a2, b1, b2; int ri1, ri2, i1, i2, j1, j2;
        ref_top2(x, r1, ri1, r2, ri2); old_top2(x, a1, i1, a2, i2); new_top2(x, b1, j1, b2, j2);
        bad_old += (a1 != r1 || i1 != ri1 || a2 != r2 || i2 != ri2);
        bad_new += (b1 != r1 || j1 != ri1 || b2 != r2 || j2 != ri2);
    }
    printf("new_vs_ref %ld old_vs_ref %ld\n", bad_new, bad_old);
    return bad_new != 0;
}

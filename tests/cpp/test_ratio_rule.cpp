// Exactness note of DESIGN.md section 4 / pbsc_walk_thread.cuh (eval4): the reference accepts an extension when
//     (double)kmerFreq / (double)maxfreq >= cutoff,  cutoff in {0.125, 0.2, 0.25, 0.3, 0.6, 2}
// (PacBio/LongReadCorrectByOverlap.cpp:733-781, compiled without FMA contraction); the kernel tests  kmerFreq * q >= p * maxfreq
// with cutoff = p/q in integers.  This program checks that the two agree for every 0 <= a <= b <= 6000 and, for b up to 2^31 - 1,
// for the a around every cutoff's boundary p*b/q — the only place where rounding could matter.
#include <cstdint>
#include <cstdio>
#include <random>

struct Cut { int p, q; double c; };
static const Cut cuts[] = {{1, 8, 0.125}, {1, 5, 0.2}, {1, 4, 0.25}, {3, 10, 0.3}, {3, 5, 0.6}, {2, 1, 2.0}};

static long bad = 0;
static void check(int64_t a, int64_t b)
{
    if (a < 0 || a > b || b <= 0) return;
    volatile double ratio = (double)a / (double)b;   // one correctly rounded IEEE division, as in the reference
    for (const Cut& k : cuts)
    {
        const bool ref = ratio >= k.c;
        const bool ours = a * k.q >= (int64_t)k.p * b;
        if (ref != ours) { if (bad++ < 10) printf("a=%lld b=%lld cutoff %d/%d: reference %d, integer rule %d\n", (long long)a, (long long)b, k.p, k.q, (int)ref, (int)ours); }
    }
}

int main()
{
    long n = 0;
    for (int64_t b = 1; b <= 6000; b++) for (int64_t a = 0; a <= b; a++) { check(a, b); n++; }
    std::mt19937_64 rng(7);
    for (int it = 0; it < 4000000; it++)
    {
        const int64_t b = 1 + (int64_t)(rng() % 2147483647ull);
        for (const Cut& k : cuts)
        {
            const int64_t edge = (int64_t)k.p * b / k.q;
            for (int d = -2; d <= 2; d++) { check(edge + d, b); n++; }
        }
        check((int64_t)(rng() % (uint64_t)(b + 1)), b); n++;
    }
    printf("%s: %ld (a, b) pairs, %ld disagreements\n", bad ? "FAILED" : "ok", n, bad);
    return bad ? 1 : 0;
}

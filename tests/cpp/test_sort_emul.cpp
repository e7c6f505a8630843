// Host check of csrc/stl_sort_emul.cuh against the real libstdc++ std::sort / std::partial_sort.
// Elements are (key << 32 | payload); the comparator looks at the key only, so equal keys expose the
// permutation.  Exit code 0 on success.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../longreadselfcorrect_b200/csrc/stl_sort_emul.cuh"

struct KeyGreater { bool operator()(uint64_t a, uint64_t b) const { return (a >> 32) > (b >> 32); } };

int main()
{
    std::mt19937_64 rng(12345);
    long checked = 0;
    for (int iter = 0; iter < 20000; iter++)
    {
        long n = (iter < 200) ? iter : (long)(rng() % 3000);
        int nkeys = 1 + (int)(rng() % (iter % 7 == 0 ? 3 : (iter % 5 == 0 ? 40 : 100000)));
        std::vector<uint64_t> a(n), b;
        int pattern = iter % 11;
        for (long i = 0; i < n; i++)
        {
            uint64_t key;
            if (pattern == 3) key = i;                       // ascending
            else if (pattern == 4) key = n - i;              // descending
            else if (pattern == 5) key = (i % 2) ? i : n - i; // organ pipe-ish
            else key = rng() % nkeys;
            a[i] = (key << 32) | (uint64_t)i;
        }
        b = a;
        std::sort(a.begin(), a.end(), KeyGreater());
        pbsc::stlsort::sort(b.data(), n, KeyGreater());
        if (a != b) { fprintf(stderr, "MISMATCH sort n=%ld iter=%d\n", n, iter); return 1; }
        checked++;
    }
    // heap-sort fallback path: compare with std::partial_sort(first, last, last)
    for (int iter = 0; iter < 3000; iter++)
    {
        long n = (long)(rng() % 600);
        std::vector<uint64_t> a(n), b;
        for (long i = 0; i < n; i++) a[i] = ((rng() % 20) << 32) | (uint64_t)i;
        b = a;
        std::partial_sort(a.begin(), a.end(), a.end(), KeyGreater());
        pbsc::stlsort::heap_sort_(b.data(), n, KeyGreater());
        if (a != b) { fprintf(stderr, "MISMATCH heap n=%ld iter=%d\n", n, iter); return 1; }
    }
    // median-of-3 killer to drive std::sort into the depth limit
    for (long n : {64L, 500L, 2000L, 5000L})
    {
        std::vector<uint64_t> a(n), b;
        // classic anti-quicksort sequence (Musser)
        long k = n / 2;
        for (long i = 1; i <= k; i++) { if (i % 2 == 1) { a[i - 1] = i; a[i] = k + i; } a[k + i - 1] = 2 * i; }
        for (long i = 0; i < n; i++) a[i] = (a[i] << 32) | (uint64_t)i;
        b = a;
        std::sort(a.begin(), a.end(), KeyGreater());
        pbsc::stlsort::sort(b.data(), n, KeyGreater());
        if (a != b) { fprintf(stderr, "MISMATCH killer n=%ld\n", n); return 1; }
        struct KeyLess { bool operator()(uint64_t x, uint64_t y) const { return (x >> 32) < (y >> 32); } };
        std::vector<uint64_t> c = b, d;
        for (long i = 0; i < n; i++) c[i] = (c[i] & 0xffffffff00000000ull) | (uint64_t)i;
        // feed the killer again with ascending comparator
        long kk = n / 2; std::vector<uint64_t> e(n);
        for (long i = 1; i <= kk; i++) { if (i % 2 == 1) { e[i - 1] = i; e[i] = kk + i; } e[kk + i - 1] = 2 * i; }
        for (long i = 0; i < n; i++) e[i] = (e[i] << 32) | (uint64_t)i;
        d = e;
        std::sort(e.begin(), e.end(), KeyLess());
        pbsc::stlsort::sort(d.data(), n, KeyLess());
        if (e != d) { fprintf(stderr, "MISMATCH killer-less n=%ld\n", n); return 1; }
    }
    printf("ok %ld\n", checked);
    return 0;
}

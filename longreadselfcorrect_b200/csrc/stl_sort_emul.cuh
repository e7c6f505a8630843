// stl_sort_emul.cuh — the permutation GNU libstdc++'s std::sort produces, reproduced on the device.
//
// The reference sorts the query k-mer intervals with `std::sort(..., std::greater<>)` on the interval
// start only (PacBio/IntervalTree.cpp:18, IntervalTree.h:27-30).  std::sort is not stable, and the order in
// which IntervalTree::findOverlapping returns equal-start intervals (identical k-mers of the query)
// decides which hit isSupportedByNewSeed keeps (PacBio/LongReadCorrectByOverlap.cpp:580-626).  The
// reference binary links libstdc++ (GCC 13), whose std::sort is the classic introsort: median-of-3
// quicksort down to 16-element runs with a 2*floor(log2 n) depth limit and heap-sort fallback, finished
// by one insertion sort.  This header restates that published algorithm for 64-bit elements compared
// by a caller-supplied strict-weak `less`-style predicate comp(a, b) ("a goes before b").
// tests/test_sort_emulation.py checks it against the real std::sort.
#ifndef PBSC_STL_SORT_EMUL_CUH
#define PBSC_STL_SORT_EMUL_CUH

#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace pbsc {
namespace stlsort {

template <class T> __host__ __device__ inline void swp(T& a, T& b) { T t = a; a = b; b = t; }

template <class T, class Comp>
__host__ __device__ inline void push_heap_(T* first, long hole, long top, T value, Comp comp)
{
    long parent = (hole - 1) / 2;
    while (hole > top && comp(first[parent], value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <class T, class Comp>
__host__ __device__ inline void adjust_heap_(T* first, long hole, long len, T value, Comp comp)
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value, comp);
}

// std::partial_sort(first, last, last): make_heap + sort_heap
template <class T, class Comp>
__host__ __device__ inline void heap_sort_(T* first, long len, Comp comp)
{
    if (len >= 2)
    {
        long parent = (len - 2) / 2;
        while (true)
        {
            T value = first[parent];
            adjust_heap_(first, parent, len, value, comp);
            if (parent == 0) break;
            parent--;
        }
    }
    long last = len;
    while (last > 1)
    {
        --last;
        T value = first[last];
        first[last] = first[0];
        adjust_heap_(first, 0, last, value, comp);
    }
}

template <class T, class Comp>
__host__ __device__ inline void move_median_to_first_(T* a0, long result, long a, long b, long c, Comp comp)
{
    if (comp(a0[a], a0[b]))
    {
        if (comp(a0[b], a0[c])) swp(a0[result], a0[b]);
        else if (comp(a0[a], a0[c])) swp(a0[result], a0[c]);
        else swp(a0[result], a0[a]);
    }
    else if (comp(a0[a], a0[c])) swp(a0[result], a0[a]);
    else if (comp(a0[b], a0[c])) swp(a0[result], a0[c]);
    else swp(a0[result], a0[b]);
}

template <class T, class Comp>
__host__ __device__ inline long unguarded_partition_(T* a0, long first, long last, long pivot, Comp comp)
{
    while (true)
    {
        while (comp(a0[first], a0[pivot])) ++first;
        --last;
        while (comp(a0[pivot], a0[last])) --last;
        if (!(first < last)) return first;
        swp(a0[first], a0[last]);
        ++first;
    }
}

template <class T, class Comp>
__host__ __device__ inline void unguarded_linear_insert_(T* a0, long last, Comp comp)
{
    T val = a0[last];
    long next = last - 1;
    while (comp(val, a0[next]))
    {
        a0[last] = a0[next];
        last = next;
        --next;
    }
    a0[last] = val;
}

template <class T, class Comp>
__host__ __device__ inline void insertion_sort_(T* a0, long first, long last, Comp comp)
{
    if (first == last) return;
    for (long i = first + 1; i != last; ++i)
    {
        if (comp(a0[i], a0[first]))
        {
            T val = a0[i];
            for (long j = i; j > first; --j) a0[j] = a0[j - 1];
            a0[first] = val;
        }
        else unguarded_linear_insert_(a0, i, comp);
    }
}

// std::sort(a, a + n, comp)
template <class T, class Comp>
__host__ __device__ inline void sort(T* a, long n, Comp comp)
{
    if (n <= 0) return;
    // __introsort_loop with an explicit stack; sub-ranges are disjoint, so processing order is irrelevant
    long stack_first[64], stack_last[64];
    int stack_depth[64];
    int sp = 0;
    int lg = 0;
    for (long t = n; t > 1; t >>= 1) lg++;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg; sp = 1;
    while (sp > 0)
    {
        --sp;
        long first = stack_first[sp], last = stack_last[sp];
        int depth = stack_depth[sp];
        while (last - first > 16)
        {
            if (depth == 0) { heap_sort_(a + first, last - first, comp); break; }
            --depth;
            const long mid = first + (last - first) / 2;
            move_median_to_first_(a, first, first + 1, mid, last - 1, comp);
            const long cut = unguarded_partition_(a, first + 1, last, first, comp);
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; sp++;
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16)
    {
        insertion_sort_(a, 0, 16, comp);
        for (long i = 16; i != n; ++i) unguarded_linear_insert_(a, i, comp);
    }
    else insertion_sort_(a, 0, n, comp);
}

}  // namespace stlsort
}  // namespace pbsc
#endif

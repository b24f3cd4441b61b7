// introsort.cuh — move-for-move emulation of libstdc++'s std::sort (GCC's bits/stl_algo.h: __sort =
// __introsort_loop + __final_insertion_sort, threshold 16, depth limit 2*floor(log2 n), heapsort
// fallback) for the oct-tree's "careful" phase.
//
// Why: the reference sorts vPrevSizeAndPointerToNode with std::sort and a comparator on
// (nKeys, UL.x) only (orb_slam3/src/ORBextractor.cc:539-560, :707).  Nodes with equal keys are
// "equivalent", std::sort is not stable, and their final order decides which nodes get expanded —
// i.e. which keypoints are selected.  Bit-exact keypoint sets therefore
// need the very same permutation libstdc++ produces.  tests/test_introsort.py compiles this header
// with g++ and checks it against std::sort on tie-heavy inputs.
//
// Usable from host and device code.  On the device one thread runs the partition phase (introsort_loop) and the
// whole CTA the final stable pass (stable_rank); n is a few hundred at most.
#pragma once
#if defined(__CUDACC__)
#define VSG_HD __host__ __device__ __forceinline__
#else
#define VSG_HD inline
#endif

namespace vsg {

// One sort element packed into 64 bits so that a compare is one load and a move one load + one store (the sort runs
// on a single thread of the CTA over shared memory): count (pair.first = node key count, 24 bits) | ulx
// (pair.second->UL.x, 16 bits) | ref (which node; NOT part of the ordering, 24 bits).
struct SortItem {
    unsigned long long v;
};

VSG_HD SortItem make_sort_item(int count, int ulx, int ref) {
    SortItem it;
    it.v = ((unsigned long long)(unsigned)count << 40) | ((unsigned long long)((unsigned)ulx & 0xFFFFu) << 24) |
           (unsigned long long)((unsigned)ref & 0xFFFFFFu);
    return it;
}
VSG_HD int sort_item_ref(const SortItem &it) { return (int)(it.v & 0xFFFFFFull); }

VSG_HD bool item_less(const SortItem &a, const SortItem &b) {  // compareNodes, ORBextractor.cc:539-560: (count, UL.x)
    return (a.v >> 24) < (b.v >> 24);
}

VSG_HD void item_swap(SortItem &a, SortItem &b) { SortItem t = a; a = b; b = t; }

// __unguarded_linear_insert
VSG_HD void linear_insert_unguarded(SortItem *v, int last) {
    const SortItem val = v[last];
    int next = last - 1;
    while (item_less(val, v[next])) {
        v[last] = v[next];
        last = next;
        --next;
    }
    v[last] = val;
}

// __insertion_sort on [first, last)
VSG_HD void insertion_sort(SortItem *v, int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (item_less(v[i], v[first])) {
            const SortItem val = v[i];
            for (int k = i; k > first; --k) v[k] = v[k - 1];  // move_backward(first, i, i + 1)
            v[first] = val;
        } else {
            linear_insert_unguarded(v, i);
        }
    }
}

// __push_heap / __adjust_heap on the sub-array starting at `base`
VSG_HD void heap_push(SortItem *v, int base, int hole, int top, const SortItem &value) {
    int parent = (hole - 1) / 2;
    while (hole > top && item_less(v[base + parent], value)) {
        v[base + hole] = v[base + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    v[base + hole] = value;
}

VSG_HD void heap_adjust(SortItem *v, int base, int hole, int len, const SortItem &value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (item_less(v[base + child], v[base + child - 1])) --child;
        v[base + hole] = v[base + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        v[base + hole] = v[base + child - 1];
        hole = child - 1;
    }
    heap_push(v, base, hole, top, value);
}

// __partial_sort(first, last, last) == make_heap + sort_heap
VSG_HD void heap_sort(SortItem *v, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            const SortItem val = v[first + parent];
            heap_adjust(v, first, parent, len, val);
            if (parent == 0) break;
            --parent;
        }
    }
    // __heap_select's scan over [middle, last) is empty because middle == last
    int end = last;
    while (end - first > 1) {
        --end;
        const SortItem val = v[end];   // __pop_heap(first, end, end)
        v[end] = v[first];
        heap_adjust(v, first, 0, end - first, val);
    }
}

// __move_median_to_first(result, a, b, c)
VSG_HD void median_to_first(SortItem *v, int result, int a, int b, int c) {
    if (item_less(v[a], v[b])) {
        if (item_less(v[b], v[c])) item_swap(v[result], v[b]);
        else if (item_less(v[a], v[c])) item_swap(v[result], v[c]);
        else item_swap(v[result], v[a]);
    } else if (item_less(v[a], v[c])) item_swap(v[result], v[a]);
    else if (item_less(v[b], v[c])) item_swap(v[result], v[c]);
    else item_swap(v[result], v[b]);
}

// __unguarded_partition(first, last, pivot)
VSG_HD int partition_unguarded(SortItem *v, int first, int last, int pivot) {
    const SortItem pv = v[pivot];   // the pivot sits in front of [first, last) and is never swapped: read it once
    while (true) {
        while (item_less(v[first], pv)) ++first;
        --last;
        while (item_less(pv, v[last])) --last;
        if (!(first < last)) return first;
        item_swap(v[first], v[last]);
        ++first;
    }
}

// __introsort_loop(v, v + n, 2 * lg(n)): median-of-3 quicksort partitions (heapsort below the depth limit) until every
// unsorted run is at most 16 long.  The first half of std::sort; inherently sequential (data-dependent swaps).
VSG_HD void introsort_loop(SortItem *v, int n) {
    if (n <= 0) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    // explicit stack instead of the recursion on the right part (the two parts are disjoint, so the processing order
    // does not change the outcome)
    int stack_first[40], stack_last[40], stack_depth[40];   // at most one push per level of the 2*lg(n) depth budget
    int sp = 0;
    stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg; sp = 1;
    while (sp > 0) {
        --sp;
        const int first = stack_first[sp];
        int last = stack_last[sp];
        int depth = stack_depth[sp];
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort(v, first, last);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            median_to_first(v, first, first + 1, mid, last - 1);
            const int cut = partition_unguarded(v, first + 1, last, first);
            stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth; ++sp;
            last = cut;
        }
    }
}

// __final_insertion_sort, the second half of std::sort.  Both of its loops only ever move an element in front of
// strictly greater ones, i.e. it is a STABLE sort of whatever arrangement introsort_loop left — so any stable sort
// produces the same permutation, e.g. the rank of every element computed independently (stable_rank below).
VSG_HD void final_insertion_sort(SortItem *v, int n) {
    if (n > 16) {
        insertion_sort(v, 0, 16);
        for (int i = 16; i < n; ++i) linear_insert_unguarded(v, i);
    } else {
        insertion_sort(v, 0, n);
    }
}

// Position of v[j] after a stable sort of v[0..n): elements ordered before it plus equivalent ones in front of it.
VSG_HD int stable_rank(const SortItem *v, int n, int j) {
    const unsigned long long key = v[j].v >> 24;
    int rank = 0;
    for (int i = 0; i < n; ++i) {
        const unsigned long long ki = v[i].v >> 24;
        rank += (ki < key) || (ki == key && i < j);
    }
    return rank;
}

// std::sort(v, v + n, compareNodes), one thread
VSG_HD void libstdcxx_sort(SortItem *v, int n) {
    if (n <= 0) return;
    introsort_loop(v, n);
    final_insertion_sort(v, n);
}

}  // namespace vsg

// Checks visual_sgraphs_b200/csrc/introsort.cuh (the std::sort emulation used by the device oct-tree)
// against the real libstdc++ std::sort with the reference's comparator semantics
// (compareNodes, orb_slam3/src/ORBextractor.cc:539-560) on tie-heavy and adversarial inputs.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>

#include "../../visual_sgraphs_b200/csrc/introsort.cuh"

struct FakeNode { int ulx; int id; };
typedef std::pair<int, FakeNode *> Entry;

static bool entry_less(Entry &a, Entry &b) {
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ulx < b.second->ulx;
}

static long run_case(const std::vector<int> &counts, const std::vector<int> &ulx) {
    const int n = (int)counts.size();
    std::vector<FakeNode> nodes(n);
    std::vector<Entry> ref(n);
    std::vector<vsg::SortItem> mine(n);
    for (int i = 0; i < n; ++i) {
        nodes[i] = {ulx[i], i};
        ref[i] = std::make_pair(counts[i], &nodes[i]);
        mine[i] = vsg::make_sort_item(counts[i], ulx[i], i);
    }
    std::vector<vsg::SortItem> two(mine), ranked(n);
    std::sort(ref.begin(), ref.end(), entry_less);
    vsg::libstdcxx_sort(mine.data(), n);
    // the device's two-phase form: sequential partition phase, then every element placed by its stable rank
    vsg::introsort_loop(two.data(), n);
    for (int j = 0; j < n; ++j) ranked[vsg::stable_rank(two.data(), n, j)] = two[j];
    for (int i = 0; i < n; ++i)
        if (ref[i].second->id != vsg::sort_item_ref(mine[i]) || ref[i].second->id != vsg::sort_item_ref(ranked[i])) return i + 1;
    return 0;
}

int main() {
    std::mt19937 rng(12345);
    long cases = 0;
    for (int n = 0; n <= 700; ++n) {
        for (int rep = 0; rep < 6; ++rep) {
            std::vector<int> c(n), u(n);
            const int cmax = 2 + (int)(rng() % (rep < 3 ? 4 : 40));
            const int umax = 1 + (int)(rng() % (rep % 2 ? 6 : 300));
            for (int i = 0; i < n; ++i) { c[i] = 2 + (int)(rng() % cmax); u[i] = (int)(rng() % umax); }
            if (long bad = run_case(c, u)) { std::printf("MISMATCH n=%d rep=%d at %ld\n", n, rep, bad - 1); return 1; }
            ++cases;
        }
    }
    // sorted / reversed / organ-pipe / all-equal / median-of-3 killer (drives the heapsort fallback)
    for (int n : {17, 33, 64, 100, 257, 1000, 4096}) {
        std::vector<int> u(n, 0), c(n);
        for (int i = 0; i < n; ++i) c[i] = i;
        if (run_case(c, u)) { std::puts("MISMATCH sorted"); return 1; }
        for (int i = 0; i < n; ++i) c[i] = n - i;
        if (run_case(c, u)) { std::puts("MISMATCH reversed"); return 1; }
        for (int i = 0; i < n; ++i) c[i] = std::min(i, n - i);
        if (run_case(c, u)) { std::puts("MISMATCH organ"); return 1; }
        for (int i = 0; i < n; ++i) c[i] = 7;
        if (run_case(c, u)) { std::puts("MISMATCH equal"); return 1; }
        // Musser-style killer for median-of-3 quicksort
        const int k = n / 2;
        for (int i = 0; i < n; ++i) {
            if (i < k) c[i] = (i % 2 == 0) ? i + 1 : k + i + (k % 2 ? 0 : 1);
            else c[i] = (i - k + 1) * 2;
        }
        if (run_case(c, u)) { std::puts("MISMATCH killer"); return 1; }
        cases += 5;
    }
    std::printf("OK %ld cases\n", cases);
    return 0;
}

#include "bvh_build.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>

#include <algorithm>
#include <new>
#include <atomic>
#include <thread>

namespace take {
namespace {

// Fork-join helper: run `left` on a new thread when the budget allows, `right` inline.
struct Forker {
    std::atomic<int> budget;
    explicit Forker(int threads) : budget(threads > 1 ? threads - 1 : 0) {}
    template <typename A, typename B>
    void both(bool big, A left, B right) {
        if (big && budget.fetch_sub(1) > 0) {
            // the slot goes back the moment the forked side is done -- not when both are -- so that the side still
            // running can fork again at its next big node (SAH splits are far from balanced)
            std::thread t([&] { left(); budget.fetch_add(1); });
            right();
            t.join();
        } else {
            if (big) budget.fetch_add(1);
            left();
            right();
        }
    }
};

// ------------------------------------------------------------------------------------------
// Reference-order tree (src/bvh.cpp:8-45).  A subtree over m primitives has 2m-1 nodes numbered in
// post-order, so a subtree starting at node index `base` puts its left child's nodes at
// [base, base+2*ml-1), the right child's right after, and itself last: the numbering the reference's
// push_back order produces, computed without a shared vector so subtrees can build in parallel.
// ------------------------------------------------------------------------------------------
struct KeyId {
    double key;
    int32_t id;
};

// f(tid, a, b) over [lo, hi) split into `parts` contiguous chunks on their own threads (the huge ranges at the top of a
// tree, where the recursion itself offers no parallelism yet)
template <typename F>
void run_chunks(int64_t lo, int64_t hi, int parts, F f) {
    if (parts <= 1) { f(0, lo, hi); return; }
    std::vector<std::thread> pool;
    const int64_t m = hi - lo;
    for (int t = 1; t < parts; ++t) pool.emplace_back(f, t, lo + m * t / parts, lo + m * (t + 1) / parts);
    f(0, lo, lo + m / parts);
    for (auto &t : pool) t.join();
}

// std::sort, run on several threads, with the IDENTICAL result -- ties included.
// The reference's tree depends on the order libstdc++'s introsort leaves equal keys in, so no other sorting algorithm may
// stand in for it.  But introsort's recursion is a tree of independent sub-ranges: after a partition step the right part is
// sorted by a recursive call and the left part by the next loop iteration (bits/stl_algo.h: __introsort_loop), with the same
// remaining depth budget.  Running the two on different threads changes nothing about what either does.  The steps
// themselves are libstdc++'s own (__unguarded_partition_pivot, __partial_sort for the heapsort fallback,
// __final_insertion_sort), called in std::sort's order, so the sequence of comparisons and moves inside every sub-range
// is the one std::sort performs.  Without libstdc++ this is plain std::sort.
template <typename It, typename Cmp>
void introsort_loop_forked(It first, It last, long depth_limit, Cmp comp, Forker &fork) {
#if defined(__GLIBCXX__)
    while (last - first > 16) {  // _S_threshold
        if (depth_limit == 0) {
            std::__partial_sort(first, last, last, comp);
            return;
        }
        --depth_limit;
        It cut = std::__unguarded_partition_pivot(first, last, comp);
        if (last - first > (1 << 15)) {
            const long d = depth_limit;
            fork.both(true, [&, d] { introsort_loop_forked(cut, last, d, comp, fork); },
                      [&, d] { introsort_loop_forked(first, cut, d, comp, fork); });
            return;
        }
        introsort_loop_forked(cut, last, depth_limit, comp, fork);
        last = cut;
    }
#endif
}
// The forked twin leans on libstdc++ internals; a different libstdc++ could change them under our feet.  So before its
// first use in a process it is checked against plain std::sort on a few adversarial inputs (heavy ties, presorted, organ
// pipe; ~10 ms): if any position differs -- or TAKE_PLAIN_STD_SORT is set -- every call falls back to std::sort itself.
bool forked_sort_trusted();

template <typename It, typename Less>
void exact_std_sort(It first, It last, Less less, int threads, bool checked = true) {
#if defined(__GLIBCXX__)
    if (threads > 1 && last - first > (1 << 16) && (!checked || forked_sort_trusted())) {
        auto comp = __gnu_cxx::__ops::__iter_comp_iter(less);
        Forker fork(threads);
        introsort_loop_forked(first, last, (long)std::__lg(last - first) * 2, comp, fork);
        std::__final_insertion_sort(first, last, comp);
        return;
    }
#endif
    std::sort(first, last, less);
}

struct RefBuilder {
    const Aabb *boxes;
    RefNode *nodes;
    int32_t *ids;
    Forker fork;
    int top_threads = 1;

    int32_t build(int64_t lo, int64_t hi, int64_t base) {
        int64_t m = hi - lo;
        if (m == 1) {
            const Aabb &b = boxes[ids[lo]];
            RefNode &n = nodes[base];
            for (int a = 0; a < 3; ++a) { n.lo[a] = b.lo[a]; n.hi[a] = b.hi[a]; }
            n.left = n.right = -1;
            n.prim = ids[lo];
            n.pad = 0;
            return (int32_t)base;
        }
        // near the root the per-primitive passes run in chunks on all threads (the recursion has not fanned out yet)
        const int parts = m >= (1 << 18) ? std::min<int>(top_threads, (int)(m >> 16)) : 1;
        RefNode big;
        for (int a = 0; a < 3; ++a) { big.lo[a] = INFINITY; big.hi[a] = -INFINITY; }
        {   // merge(), src/bbox.h:45-55 (min / max are exact: the order of the union does not matter)
            std::vector<RefNode> part((size_t)parts, big);
            run_chunks(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                RefNode t = part[tid];
                for (int64_t i = a0; i < a1; ++i) {
                    const Aabb &b = boxes[ids[i]];
                    for (int a = 0; a < 3; ++a) {
                        t.lo[a] = std::min(t.lo[a], b.lo[a]);
                        t.hi[a] = std::max(t.hi[a], b.hi[a]);
                    }
                }
                part[tid] = t;
            });
            for (int t = 0; t < parts; ++t)
                for (int a = 0; a < 3; ++a) {
                    big.lo[a] = std::min(big.lo[a], part[t].lo[a]);
                    big.hi[a] = std::max(big.hi[a], part[t].hi[a]);
                }
        }
        double ex = big.hi[0] - big.lo[0], ey = big.hi[1] - big.lo[1], ez = big.hi[2] - big.lo[2];
        int axis = (ex > ey && ex > ez) ? 0 : (ey > ex && ey > ez) ? 1 : 2;  // largest_axis(), bbox.h:34-43
        {
            std::vector<KeyId> tmp((size_t)m);
            run_chunks(lo, hi, parts, [&](int, int64_t a0, int64_t a1) {
                for (int64_t i = a0; i < a1; ++i) {
                    const Aabb &b = boxes[ids[i]];
                    tmp[i - lo] = {(b.hi[axis] + b.lo[axis]) * 0.5, ids[i]};  // centre = (p_max + p_min) * (1/2)
                }
            });
            // The tie order among equal centroids is whatever libstdc++'s introsort yields for this sequence of
            // comparison outcomes; the reference sorts whole BBoxWithID values with the same comparator, so the
            // permutation is identical.
            // (near the root, where this recursion offers no parallelism yet, the sort itself forks: exact_std_sort)
            exact_std_sort(tmp.begin(), tmp.end(), [](const KeyId &a, const KeyId &b) { return a.key < b.key; },
                           m >= (1 << 18) ? top_threads : 1);
            run_chunks(lo, hi, parts, [&](int, int64_t a0, int64_t a1) {
                for (int64_t i = a0; i < a1; ++i) ids[i] = tmp[i - lo].id;
            });
        }
        int64_t ml = m / 2, mr = m - ml;
        int64_t mid = lo + ml;
        int32_t l = -1, r = -1;
        fork.both(m > 65536, [&] { l = build(lo, mid, base); }, [&] { r = build(mid, hi, base + 2 * ml - 1); });
        big.left = l;
        big.right = r;
        big.prim = -1;
        big.pad = 0;
        int64_t self = base + 2 * ml - 1 + 2 * mr - 1;
        nodes[self] = big;
        return (int32_t)self;
    }
};

// ------------------------------------------------------------------------------------------
// Fast tree: binned SAH over centroids.
// ------------------------------------------------------------------------------------------
struct TmpNode {
    Aabb box;
    int32_t left = -1, right = -1;  // inner
    int64_t first = 0;
    int32_t count = 0;              // leaf
};

// Node pool of the SAH builder: raw storage, a node is initialised when it is handed out (value-initialising 2n nodes up
// front was a serial gigabyte of page faults at 10 M primitives; now the building threads touch their own nodes).
struct TmpNodePool {
    TmpNode *p = nullptr;
    size_t n = 0;
    explicit TmpNodePool(size_t count) : p((TmpNode *)malloc(count * sizeof(TmpNode))), n(count) {}
    ~TmpNodePool() { free(p); }
    TmpNodePool(const TmpNodePool &) = delete;
    TmpNodePool &operator=(const TmpNodePool &) = delete;
    TmpNode &operator[](size_t i) { return p[i]; }
    const TmpNode &operator[](size_t i) const { return p[i]; }
    size_t size() const { return n; }
};

inline double half_area(const double *lo, const double *hi) {
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

// One primitive as the builder moves it around: its box and id side by side (64 bytes), so that every pass over a range
// streams memory instead of gathering 48-byte boxes through an index array.
struct alignas(64) PrimRef {
    Aabb box;
    int64_t id;
    int64_t pad;
};
static_assert(sizeof(PrimRef) == 64, "PrimRef");

struct Bounds {  // box of a primitive set and box of its centroids
    Aabb bb, cb;
    void clear() {
        for (int a = 0; a < 3; ++a) { bb.lo[a] = cb.lo[a] = INFINITY; bb.hi[a] = cb.hi[a] = -INFINITY; }
    }
    void add(const Aabb &b) {
        for (int a = 0; a < 3; ++a) {
            bb.lo[a] = std::min(bb.lo[a], b.lo[a]);
            bb.hi[a] = std::max(bb.hi[a], b.hi[a]);
            const double c = 0.5 * (b.lo[a] + b.hi[a]);
            cb.lo[a] = std::min(cb.lo[a], c);
            cb.hi[a] = std::max(cb.hi[a], c);
        }
    }
    void merge(const Bounds &o) {
        for (int a = 0; a < 3; ++a) {
            bb.lo[a] = std::min(bb.lo[a], o.bb.lo[a]); bb.hi[a] = std::max(bb.hi[a], o.bb.hi[a]);
            cb.lo[a] = std::min(cb.lo[a], o.cb.lo[a]); cb.hi[a] = std::max(cb.hi[a], o.cb.hi[a]);
        }
    }
};

// Two streaming passes per node: (1) bin the three axes, (2) partition the records in place while accumulating the
// bounds of the two sides -- so a child starts with its boxes known and never re-reads its range for them.
struct SahBuilder {
    static const int NBINS = 32;
    static const int SMALL = 16;  // ranges up to this size take the sparse sweep
    PrimRef *refs;
    PrimRef *scratch;  // same length as refs: source of the chunk-parallel partitions near the root
    TmpNodePool nodes;
    std::atomic<int32_t> next{0};
    int max_leaf;
    double c_trav, c_isect;
    Forker fork;

    SahBuilder(PrimRef *r, PrimRef *sc, int64_t n, int ml, int threads)
        : refs(r), scratch(sc), nodes((size_t)std::max<int64_t>(2 * n, 2)), max_leaf(ml), c_trav(1.0), c_isect(1.2), fork(threads) {
        if (const char *e = getenv("TAKE_SAH_CISECT")) c_isect = atof(e);  // tuning knob: cost of a leaf test relative to a node visit
    }

    int32_t alloc() {
        const int32_t i = next.fetch_add(1);
        new (&nodes[i]) TmpNode();
        return i;
    }

    void make_leaf(TmpNode &n, int64_t lo, int64_t hi) {
        n.first = lo;
        n.count = (int32_t)(hi - lo);
    }

    // Run f(tid, a, b) over [lo, hi) split into `parts` contiguous chunks on their own threads (used only for the few huge
    // ranges at the top of the tree, where the recursion itself offers no parallelism yet).
    template <typename F>
    static void chunked(int64_t lo, int64_t hi, int parts, F f) { run_chunks(lo, hi, parts, f); }
    int top_threads = 1;  // threads for the chunked top-level loops

    struct Bins {
        Aabb box[3][NBINS];
        int64_t cnt[3][NBINS];
        void clear() {
            for (int x = 0; x < 3; ++x)
                for (int b = 0; b < NBINS; ++b) {
                    cnt[x][b] = 0;
                    for (int a = 0; a < 3; ++a) { box[x][b].lo[a] = INFINITY; box[x][b].hi[a] = -INFINITY; }
                }
        }
    };

    // bounds of the whole input (the only range whose bounds no parent hands down)
    Bounds range_bounds(int64_t lo, int64_t hi) {
        const int parts = (hi - lo) >= (1 << 18) ? std::min<int>(top_threads, (int)((hi - lo) >> 16)) : 1;
        std::vector<Bounds> pb((size_t)parts);
        chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
            Bounds t;
            t.clear();
            for (int64_t i = a0; i < a1; ++i) t.add(refs[i].box);
            pb[tid] = t;
        });
        for (int t = 1; t < parts; ++t) pb[0].merge(pb[t]);
        return pb[0];
    }

    void build(int32_t node_id, int64_t lo, int64_t hi, const Bounds &bounds) {
        TmpNode &node = nodes[node_id];
        const int64_t m = hi - lo;
        const int parts = m >= (1 << 18) ? std::min<int>(top_threads, (int)(m >> 16)) : 1;
        const Aabb &bb = bounds.bb, &cb = bounds.cb;
        node.box = bb;
        if (m == 1) { make_leaf(node, lo, hi); return; }

        // pass 1: bin all three axes and pick the cheapest split
        double scale3[3];
        bool axis_ok[3];
        for (int x = 0; x < 3; ++x) {
            double ext = cb.hi[x] - cb.lo[x];
            axis_ok[x] = ext > 0;
            scale3[x] = axis_ok[x] ? NBINS / ext : 0.0;
        }
        double best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        const double parent_area = half_area(bb.lo, bb.hi);
        if (m <= SMALL) {
            // Most NODES are tiny, and for them clearing, filling and sweeping 3 x 32 dense bins is the whole cost of the
            // build.  The sweep only changes where a bin is occupied, so a handful of primitives are ordered by bin index
            // and swept directly: same candidate splits, same boxes, same costs, same choice as the dense sweep below.
            for (int axis = 0; axis < 3; ++axis) {
                if (!axis_ok[axis]) continue;
                int key[SMALL], ord[SMALL];
                for (int i = 0; i < (int)m; ++i) {
                    const Aabb &b = refs[lo + i].box;
                    int k = (int)((0.5 * (b.lo[axis] + b.hi[axis]) - cb.lo[axis]) * scale3[axis]);
                    key[i] = std::min(std::max(k, 0), NBINS - 1);
                    int j = i;
                    while (j > 0 && key[ord[j - 1]] > key[i]) { ord[j] = ord[j - 1]; --j; }
                    ord[j] = i;
                }
                // suffix: box area and count of everything in bins >= key[ord[i]] (valid at the first entry of a bin group)
                double right_area[SMALL + 1];
                Aabb acc;
                for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
                for (int i = (int)m - 1; i >= 0; --i) {
                    const Aabb &b = refs[lo + ord[i]].box;
                    for (int a = 0; a < 3; ++a) { acc.lo[a] = std::min(acc.lo[a], b.lo[a]); acc.hi[a] = std::max(acc.hi[a], b.hi[a]); }
                    right_area[i] = half_area(acc.lo, acc.hi);
                }
                for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
                for (int i = 0; i < (int)m - 1; ++i) {
                    const Aabb &b = refs[lo + ord[i]].box;
                    for (int a = 0; a < 3; ++a) { acc.lo[a] = std::min(acc.lo[a], b.lo[a]); acc.hi[a] = std::max(acc.hi[a], b.hi[a]); }
                    if (key[ord[i + 1]] == key[ord[i]]) continue;  // inside a bin group: not a split position
                    const int64_t cl = i + 1, cr = m - cl;
                    double cost = half_area(acc.lo, acc.hi) * cl + right_area[i + 1] * cr;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = key[ord[i]]; }
                }
            }
        } else {
        std::vector<Bins> pbins((size_t)parts);
        chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
            Bins &B = pbins[tid];
            B.clear();
            for (int64_t i = a0; i < a1; ++i) {
                const Aabb &b = refs[i].box;
                for (int x = 0; x < 3; ++x) {
                    if (!axis_ok[x]) continue;
                    int k = (int)((0.5 * (b.lo[x] + b.hi[x]) - cb.lo[x]) * scale3[x]);
                    k = std::min(std::max(k, 0), NBINS - 1);
                    B.cnt[x][k]++;
                    for (int a = 0; a < 3; ++a) {
                        B.box[x][k].lo[a] = std::min(B.box[x][k].lo[a], b.lo[a]);
                        B.box[x][k].hi[a] = std::max(B.box[x][k].hi[a], b.hi[a]);
                    }
                }
            }
        });
        Bins &bins = pbins[0];
        for (int t = 1; t < parts; ++t)
            for (int x = 0; x < 3; ++x)
                for (int b = 0; b < NBINS; ++b) {
                    bins.cnt[x][b] += pbins[t].cnt[x][b];
                    for (int a = 0; a < 3; ++a) {
                        bins.box[x][b].lo[a] = std::min(bins.box[x][b].lo[a], pbins[t].box[x][b].lo[a]);
                        bins.box[x][b].hi[a] = std::max(bins.box[x][b].hi[a], pbins[t].box[x][b].hi[a]);
                    }
                }

        // best split over the three axes
        for (int axis = 0; axis < 3; ++axis) {
            if (!axis_ok[axis]) continue;
            const Aabb *bin_box = bins.box[axis];
            const int64_t *bin_cnt = bins.cnt[axis];
            double right_area[NBINS];
            int64_t right_cnt[NBINS];
            Aabb acc;
            for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
            int64_t cnt = 0;
            for (int b = NBINS - 1; b > 0; --b) {
                cnt += bin_cnt[b];
                for (int a = 0; a < 3; ++a) {
                    acc.lo[a] = std::min(acc.lo[a], bin_box[b].lo[a]);
                    acc.hi[a] = std::max(acc.hi[a], bin_box[b].hi[a]);
                }
                right_area[b] = half_area(acc.lo, acc.hi);
                right_cnt[b] = cnt;
            }
            for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
            cnt = 0;
            for (int b = 0; b < NBINS - 1; ++b) {  // split between bin b and b+1
                cnt += bin_cnt[b];
                for (int a = 0; a < 3; ++a) {
                    acc.lo[a] = std::min(acc.lo[a], bin_box[b].lo[a]);
                    acc.hi[a] = std::max(acc.hi[a], bin_box[b].hi[a]);
                }
                if (cnt == 0 || right_cnt[b + 1] == 0) continue;
                double cost = half_area(acc.lo, acc.hi) * cnt + right_area[b + 1] * right_cnt[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        }
        // pass 2: partition in place, accumulating the bounds of both sides
        int64_t mid = -1;
        Bounds bl, br;
        bl.clear(); br.clear();
        if (best_axis >= 0) {
            double split_cost = c_trav + c_isect * best_cost / (parent_area > 0 ? parent_area : 1.0);
            double leaf_cost = c_isect * (double)m;
            if (m <= max_leaf && leaf_cost <= split_cost) { make_leaf(node, lo, hi); return; }
            double ext = cb.hi[best_axis] - cb.lo[best_axis];
            double scale = NBINS / ext;
            auto goes_left = [&](const Aabb &b) {
                int k = (int)((0.5 * (b.lo[best_axis] + b.hi[best_axis]) - cb.lo[best_axis]) * scale);
                k = std::min(std::max(k, 0), NBINS - 1);
                return k <= best_bin;
            };
            if (parts > 1) {
                // chunk-wise counting partition through the scratch copy (which side a primitive lands on is all that
                // matters; the order inside a side only permutes leaf slots)
                std::vector<int64_t> nleft((size_t)parts + 1, 0);
                std::vector<Bounds> pl((size_t)parts), pr((size_t)parts);
                chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                    int64_t c = 0;
                    for (int64_t i = a0; i < a1; ++i) {
                        c += goes_left(refs[i].box) ? 1 : 0;
                        scratch[i] = refs[i];
                    }
                    nleft[tid + 1] = c;
                });
                for (int t = 0; t < parts; ++t) nleft[t + 1] += nleft[t];
                const int64_t total_left = nleft[parts];
                chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                    int64_t l = lo + nleft[tid], r = lo + total_left + (a0 - lo) - nleft[tid];
                    Bounds tl, tr;
                    tl.clear(); tr.clear();
                    for (int64_t i = a0; i < a1; ++i) {
                        const PrimRef &p = scratch[i];
                        if (goes_left(p.box)) { refs[l++] = p; tl.add(p.box); }
                        else { refs[r++] = p; tr.add(p.box); }
                    }
                    pl[tid] = tl; pr[tid] = tr;
                });
                for (int t = 0; t < parts; ++t) { bl.merge(pl[t]); br.merge(pr[t]); }
                mid = lo + total_left;
            } else {
                // Hoare partition from both ends; every record is inspected exactly once
                int64_t i = lo, j = hi - 1;
                for (;;) {
                    while (i <= j && goes_left(refs[i].box)) { bl.add(refs[i].box); ++i; }
                    while (i <= j && !goes_left(refs[j].box)) { br.add(refs[j].box); --j; }
                    if (i >= j) break;
                    std::swap(refs[i], refs[j]);
                    bl.add(refs[i].box); br.add(refs[j].box);
                    ++i; --j;
                }
                mid = i;
            }
        }
        if (mid <= lo || mid >= hi) {
            // all centroids coincide (or binning failed): leaf if allowed, otherwise split the range in half
            if (m <= max_leaf) { make_leaf(node, lo, hi); return; }
            mid = lo + m / 2;
            bl.clear(); br.clear();
            for (int64_t i = lo; i < mid; ++i) bl.add(refs[i].box);
            for (int64_t i = mid; i < hi; ++i) br.add(refs[i].box);
        }
        int32_t l = alloc(), r = alloc();
        nodes[node_id].left = l;
        nodes[node_id].right = r;
        fork.both(m > 32768, [&, bl] { build(l, lo, mid, bl); }, [&, br] { build(r, mid, hi, br); });
    }
};

// Conservative FP32 image of an FP64 bound: round outward, move out by `pad`, and one more ulp to absorb the
// rounding of that subtraction / addition.
// nextafterf(f, +-INFINITY) on the bit pattern (the libm call was a third of the flatten time: ~36 calls per 4-wide node)
inline float step_up(float f) {
    if (!(f < INFINITY)) return f;  // +inf, NaN
    if (f == 0.0f) return 1.401298464324817e-45f;
    uint32_t b;
    memcpy(&b, &f, 4);
    b = f > 0.0f ? b + 1u : b - 1u;
    memcpy(&f, &b, 4);
    return f;
}
inline float step_down(float f) { return -step_up(-f); }
inline float round_down(double v, float pad) {
    float f = (float)v;
    if ((double)f > v) f = step_down(f);
    return step_down(f - pad);
}
inline float round_up(double v, float pad) {
    float f = (float)v;
    if ((double)f < v) f = step_up(f);
    return step_up(f + pad);
}

inline int32_t leaf_code(int64_t first, int32_t count) { return ~(int32_t)((first << 3) | (int64_t)(count - 1)); }

// Flattening = numbering the inner nodes in pre-order and writing their children's conservative FP32 boxes.  It walks
// the temporary tree in an order unrelated to its memory order, so it is bound by cache misses; for big trees the top
// `split_depth` levels are flattened serially, every subtree hanging below them is flattened into its own buffer on its
// own thread, and the buffers are appended with their links shifted.  The node order is then "top part, subtree,
// subtree, ..." (each subtree still contiguous and in pre-order); any order is a valid tree.
struct FlatTask {
    int32_t tmp_id;   // subtree root in the temporary tree
    int32_t out_id;   // parent node in the output ...
    int32_t slot;     // ... and the child slot that will point to the subtree
};
struct FlatItem { int32_t tmp_id, out_id, depth; };

struct BinaryPolicy {
    typedef FastNode Node;
    const TmpNodePool &tmp;
    float pad;
    static void set_link(FastNode &n, int k, int32_t idx) { (k == 0 ? n.child0 : n.child1) = idx; }
    static void shift_links(FastNode &n, int32_t base) {
        if (n.child0 >= 0) n.child0 += base;
        if (n.child1 >= 0) n.child1 += base;
    }
    void set_child(FastNode &n, int which, const TmpNode *c, int32_t idx, int32_t cnt) const {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            if (c) { lo[a] = round_down(c->box.lo[a], pad); hi[a] = round_up(c->box.hi[a], pad); }
            else { lo[a] = INFINITY; hi[a] = -INFINITY; }  // empty child: never hit
        }
        if (which == 0) {
            n.c0lox = lo[0]; n.c0hix = hi[0]; n.c0loy = lo[1]; n.c0hiy = hi[1]; n.c0loz = lo[2]; n.c0hiz = hi[2];
            n.child0 = idx; n.count0 = cnt;
        } else {
            n.c1lox = lo[0]; n.c1hix = hi[0]; n.c1loy = lo[1]; n.c1hiy = hi[1]; n.c1loz = lo[2]; n.c1hiz = hi[2];
            n.child1 = idx; n.count1 = cnt;
        }
    }
    // children of the output node that stands for temporary node `t`
    int kids(const TmpNode &t, int32_t *out_kids) const {
        out_kids[0] = t.left; out_kids[1] = t.right;
        return 2;
    }
    static const int WIDTH = 2;
    static const int32_t EMPTY_LINK = ~0;
    // the tree is a single leaf (or empty): wrap it in a root node
    void wrap_root(FastNode &n, const TmpNode &r) const {
        memset(&n, 0, sizeof(FastNode));
        if (r.count > 0) set_child(n, 0, &r, leaf_code(r.first, r.count), r.count);
        else set_child(n, 0, nullptr, ~0, 0);
        set_child(n, 1, nullptr, ~0, 0);
    }
};

// Collapse the binary SAH tree into 4-wide nodes: start from a binary node's two children and keep replacing the inner
// child with the largest surface area by its own two children until there are four (or only leaves are left).
struct WidePolicy {
    typedef WideNode Node;
    const TmpNodePool &tmp;
    float pad;
    static void set_link(WideNode &n, int k, int32_t idx) { n.child[k] = idx; }
    static void shift_links(WideNode &n, int32_t base) {
        for (int k = 0; k < 4; ++k)
            if (n.child[k] >= 0 && n.child[k] != TAKE_WIDE_EMPTY) n.child[k] += base;
    }
    void set_child(WideNode &n, int k, const TmpNode *c, int32_t idx, int32_t cnt) const {
        if (c) {
            n.lox[k] = round_down(c->box.lo[0], pad); n.hix[k] = round_up(c->box.hi[0], pad);
            n.loy[k] = round_down(c->box.lo[1], pad); n.hiy[k] = round_up(c->box.hi[1], pad);
            n.loz[k] = round_down(c->box.lo[2], pad); n.hiz[k] = round_up(c->box.hi[2], pad);
        } else {
            n.lox[k] = n.loy[k] = n.loz[k] = INFINITY;
            n.hix[k] = n.hiy[k] = n.hiz[k] = -INFINITY;
        }
        n.child[k] = idx;
        n.count[k] = cnt;
    }
    int kids(const TmpNode &t, int32_t *out_kids) const {
        int nk = 0;
        out_kids[nk++] = t.left;
        out_kids[nk++] = t.right;
        while (nk < 4) {
            int best = -1;
            double best_area = -1;
            for (int k = 0; k < nk; ++k) {
                const TmpNode &c = tmp[out_kids[k]];
                if (c.count > 0) continue;  // a leaf stays a leaf
                double a = half_area(c.box.lo, c.box.hi);
                if (a > best_area) { best_area = a; best = k; }
            }
            if (best < 0) break;
            const TmpNode &c = tmp[out_kids[best]];
            out_kids[best] = c.left;
            out_kids[nk++] = c.right;
        }
        return nk;
    }
    static const int WIDTH = 4;
    static const int32_t EMPTY_LINK = TAKE_WIDE_EMPTY;
    void wrap_root(WideNode &n, const TmpNode &r) const {
        memset(&n, 0, sizeof(WideNode));
        for (int k = 0; k < 4; ++k) set_child(n, k, nullptr, TAKE_WIDE_EMPTY, 0);
        if (r.count > 0) set_child(n, 0, &r, leaf_code(r.first, r.count), r.count);
    }
};

template <typename P>
struct FlattenerT {
    typedef typename P::Node Node;
    P pol;
    std::vector<Node> &out;
    int depth = 0;

    // Pre-order flatten of the subtree under temporary node `root` into `dst` (explicit stack: no recursion depth limits
    // on degenerate trees); dst[0] must already exist and stands for `root`.  Inner children found at depth
    // `defer_depth` are not expanded but listed in `tasks` (when given).  Returns the deepest level reached.
    int flatten(std::vector<Node> &dst, int32_t root, int depth0, int defer_depth, std::vector<FlatTask> *tasks) const {
        std::vector<FlatItem> stack;
        stack.push_back({root, 0, depth0});
        int deepest = depth0;
        while (!stack.empty()) {
            FlatItem it = stack.back();
            stack.pop_back();
            deepest = std::max(deepest, it.depth);
            int32_t kids[4];
            const int nk = pol.kids(pol.tmp[it.tmp_id], kids);
            for (int k = 0; k < P::WIDTH; ++k) {
                if (k >= nk) { pol.set_child(dst[it.out_id], k, nullptr, P::EMPTY_LINK, 0); continue; }
                const TmpNode *c = &pol.tmp[kids[k]];
                if (c->count > 0) {
                    pol.set_child(dst[it.out_id], k, c, leaf_code(c->first, c->count), c->count);
                } else if (tasks && it.depth >= defer_depth) {
                    pol.set_child(dst[it.out_id], k, c, 0, 0);  // link patched when the subtree has its place
                    tasks->push_back({kids[k], it.out_id, k});
                } else {
                    int32_t id = (int32_t)dst.size();
                    dst.emplace_back();
                    pol.set_child(dst[it.out_id], k, c, id, 0);
                    stack.push_back({kids[k], id, it.depth + 1});
                }
            }
        }
        return deepest;
    }

    void run(int32_t root, int threads) {
        out.clear();
        const TmpNode &r = pol.tmp[root];
        out.emplace_back();
        if (r.count > 0 || r.left < 0) {  // root is a leaf (or the tree is empty): wrap it
            pol.wrap_root(out[0], r);
            depth = 1;
            return;
        }
        const size_t n_tmp = pol.tmp.size();
        if (threads <= 1 || n_tmp < (size_t)(1 << 18) || getenv("TAKE_SERIAL_FLATTEN")) {  // (the knob is for A/B runs)
            out.reserve(n_tmp / P::WIDTH + 2);
            depth = flatten(out, root, 1, 0, nullptr);
            return;
        }
        // top of the tree serially, down to a level with enough subtrees to balance the threads
        const int split_depth = P::WIDTH == 2 ? 9 : 5;
        std::vector<FlatTask> tasks;
        depth = flatten(out, root, 1, split_depth, &tasks);
        const size_t nt = tasks.size();
        std::vector<std::vector<Node>> parts(nt);
        std::vector<int> deep(nt, 0);
        std::atomic<size_t> next{0};
        auto worker = [&] {
            for (;;) {
                const size_t t = next.fetch_add(1);
                if (t >= nt) return;
                parts[t].emplace_back();
                deep[t] = flatten(parts[t], tasks[t].tmp_id, split_depth + 1, 0, nullptr);
            }
        };
        {
            std::vector<std::thread> pool;
            for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
            worker();
            for (auto &t : pool) t.join();
        }
        std::vector<size_t> base(nt + 1, out.size());
        for (size_t t = 0; t < nt; ++t) base[t + 1] = base[t] + parts[t].size();
        out.resize(base[nt]);
        next = 0;
        auto copier = [&] {
            for (;;) {
                const size_t t = next.fetch_add(1);
                if (t >= nt) return;
                Node *dst = out.data() + base[t];
                for (size_t i = 0; i < parts[t].size(); ++i) {
                    Node n = parts[t][i];
                    P::shift_links(n, (int32_t)base[t]);
                    dst[i] = n;
                }
                P::set_link(out[tasks[t].out_id], tasks[t].slot, (int32_t)base[t]);
                std::vector<Node>().swap(parts[t]);
            }
        };
        {
            std::vector<std::thread> pool;
            for (int t = 1; t < threads; ++t) pool.emplace_back(copier);
            copier();
            for (auto &t : pool) t.join();
        }
        for (size_t t = 0; t < nt; ++t) depth = std::max(depth, deep[t]);
    }
};

}  // namespace

// Self-test of exact_std_sort (tests/test_capi_host.py): sorts `n` (key, id) pairs drawn from `distinct` different keys
// (0: all keys different) in the given arrangement with std::sort and with the forked version, returns the number of
// positions where the two results differ (must be 0).  pattern: 0 random, 1 ascending, 2 descending, 3 all equal,
// 4 organ pipe.
static int64_t sort_selftest_impl(int64_t n, int64_t distinct, int pattern, int threads, uint64_t seed, bool checked);
int64_t sort_selftest(int64_t n, int64_t distinct, int pattern, int threads, uint64_t seed) {
    return sort_selftest_impl(n, distinct, pattern, threads, seed, true);
}
namespace {
bool forked_sort_trusted() {
    static const bool ok = [] {
        const char *e = getenv("TAKE_PLAIN_STD_SORT");
        if (e && *e && *e != '0') return false;
        const int64_t n = (1 << 16) + 4321;  // just above the size from which exact_std_sort forks
        int64_t bad = 0;
        bad += sort_selftest_impl(n, 0, 0, 4, 1, false);      // random, all different
        bad += sort_selftest_impl(n, 7, 0, 4, 2, false);      // heavy ties
        bad += sort_selftest_impl(n, 0, 1, 4, 3, false);      // ascending
        bad += sort_selftest_impl(n, 0, 4, 4, 4, false);      // organ pipe (drives introsort towards its heapsort fallback)
        bad += sort_selftest_impl(n, 1000, 2, 4, 5, false);   // descending with ties
        if (bad) fprintf(stderr, "[take_gpu] the multi-threaded twin of std::sort disagrees with this libstdc++'s std::sort: "
                                 "using plain std::sort for the reference-order tree\n");
        return bad == 0;
    }();
    return ok;
}
}  // namespace
static int64_t sort_selftest_impl(int64_t n, int64_t distinct, int pattern, int threads, uint64_t seed, bool checked) {
    std::vector<KeyId> a((size_t)n);
    uint64_t x = seed * 2862933555777941757ULL + 3037000493ULL;
    for (int64_t i = 0; i < n; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        double k;
        if (pattern == 1) k = (double)i;
        else if (pattern == 2) k = (double)(n - i);
        else if (pattern == 3) k = 1.0;
        else if (pattern == 4) k = (double)(i < n / 2 ? i : n - i);
        else k = (double)(x >> 11);
        if (distinct > 0) k = fmod(k, (double)distinct);
        a[i] = {k, (int32_t)i};
    }
    std::vector<KeyId> b(a);
    auto less = [](const KeyId &p, const KeyId &q) { return p.key < q.key; };
    std::sort(a.begin(), a.end(), less);
    exact_std_sort(b.begin(), b.end(), less, threads, checked);
    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) bad += (a[i].id != b[i].id || a[i].key != b[i].key) ? 1 : 0;
    return bad;
}

void build_reference_tree(const Aabb *boxes, int64_t n, int threads, RefTree &out) {
    out.nodes.clear();
    out.dfs_rank.clear();
    out.root = -1;
    if (n <= 0) return;
    out.nodes.resize((size_t)(2 * n - 1));
    std::vector<int32_t> ids((size_t)n);
    for (int64_t i = 0; i < n; ++i) ids[i] = (int32_t)i;
    RefBuilder b{boxes, out.nodes.data(), ids.data(), Forker(threads)};
    b.top_threads = std::max(1, threads);
    out.root = b.build(0, n, 0);
    // After the recursion ids[] lists the primitives in left-to-right leaf order.
    out.dfs_rank.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) out.dfs_rank[ids[i]] = (int32_t)i;
}

void build_fast_tree(const Aabb *boxes, int64_t n, int max_leaf, float pad, int threads, FastTree &out) {
    out.nodes.clear();
    out.leaf_prims.assign((size_t)std::max<int64_t>(n, 0), 0);
    const int64_t nn = std::max<int64_t>(n, 0);
    timespec tsr; clock_gettime(CLOCK_MONOTONIC, &tsr);
    const double tr0 = tsr.tv_sec * 1e3 + tsr.tv_nsec * 1e-6;
    // records moved by the partitions + the scratch copy of the chunk-parallel ones (64-byte aligned)
    PrimRef *refs = nn ? (PrimRef *)aligned_alloc(64, (size_t)nn * sizeof(PrimRef)) : nullptr;
    PrimRef *scratch = nn ? (PrimRef *)aligned_alloc(64, (size_t)nn * sizeof(PrimRef)) : nullptr;
    const int fill_parts = nn >= (1 << 16) ? std::max(1, threads) : 1;
    SahBuilder::chunked(0, nn, fill_parts, [&](int, int64_t a0, int64_t a1) {
        for (int64_t i = a0; i < a1; ++i) { refs[i].box = boxes[i]; refs[i].id = i; refs[i].pad = 0; }
    });
    SahBuilder b(refs, scratch, n, std::max(1, max_leaf), threads);
    b.top_threads = std::max(1, threads);
    int32_t root = b.alloc();
    const bool times = getenv("TAKE_BUILD_TIMES") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t0 = now();
    if (n > 0) {
        const Bounds all = b.range_bounds(0, n);
        b.build(root, 0, n, all);
    }
    SahBuilder::chunked(0, nn, fill_parts, [&](int, int64_t a0, int64_t a1) {
        for (int64_t i = a0; i < a1; ++i) out.leaf_prims[i] = (int32_t)refs[i].id;
    });
    free(refs);
    free(scratch);
    double t1 = now();
    // the binary and the 4-wide image of the tree are independent: flatten them side by side
    FlattenerT<BinaryPolicy> f{BinaryPolicy{b.nodes, pad}, out.nodes};
    FlattenerT<WidePolicy> wf{WidePolicy{b.nodes, pad}, out.wide};
    if (threads > 1 && nn >= (1 << 16)) {
        const int tw_threads = std::max(1, threads / 2);
        std::thread tw([&] { wf.run(root, tw_threads); });
        f.run(root, std::max(1, threads - tw_threads));
        tw.join();
    } else {
        f.run(root, 1);
        wf.run(root, 1);
    }
    out.depth = f.depth;
    out.wide_depth = wf.depth;
    double t2 = now();
    if (times) fprintf(stderr, "[take_gpu] fast tree: records %.0f ms, sah build %.0f ms, flatten (binary | 4-wide) %.0f ms\n", t0 - tr0, t1 - t0, t2 - t1);
    // SAH cost of the final tree (diagnostic): partial sums over a fixed number of chunks, added in chunk order, so the
    // value does not depend on the thread count
    double cost = 0, root_area = n > 0 ? half_area(b.nodes[root].box.lo, b.nodes[root].box.hi) : 0;
    if (root_area > 0) {
        const int64_t used = b.next.load();
        const int CH = 64;
        double part[CH];
        std::atomic<int> next_chunk{0};
        auto work = [&] {
            for (;;) {
                const int c = next_chunk.fetch_add(1);
                if (c >= CH) return;
                double acc = 0;
                for (int64_t i = used * c / CH; i < used * (c + 1) / CH; ++i) {
                    const TmpNode &t = b.nodes[i];
                    double a = half_area(t.box.lo, t.box.hi) / root_area;
                    acc += t.count > 0 ? a * b.c_isect * t.count : a * b.c_trav;
                }
                part[c] = acc;
            }
        };
        std::vector<std::thread> pool;
        const int workers = used >= (1 << 18) ? std::max(1, threads) : 1;
        for (int t = 1; t < workers; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        for (int c = 0; c < CH; ++c) cost += part[c];
    }
    out.sah_cost = cost;
}

}  // namespace take

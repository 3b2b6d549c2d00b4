#include "bvh_build.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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
            std::thread t(left);
            right();
            t.join();
            budget.fetch_add(1);
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

struct RefBuilder {
    const Aabb *boxes;
    RefNode *nodes;
    int32_t *ids;
    Forker fork;

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
        RefNode big;
        for (int a = 0; a < 3; ++a) { big.lo[a] = INFINITY; big.hi[a] = -INFINITY; }
        for (int64_t i = lo; i < hi; ++i) {  // merge(), src/bbox.h:45-55
            const Aabb &b = boxes[ids[i]];
            for (int a = 0; a < 3; ++a) {
                big.lo[a] = std::min(big.lo[a], b.lo[a]);
                big.hi[a] = std::max(big.hi[a], b.hi[a]);
            }
        }
        double ex = big.hi[0] - big.lo[0], ey = big.hi[1] - big.lo[1], ez = big.hi[2] - big.lo[2];
        int axis = (ex > ey && ex > ez) ? 0 : (ey > ex && ey > ez) ? 1 : 2;  // largest_axis(), bbox.h:34-43
        {
            std::vector<KeyId> tmp((size_t)m);
            for (int64_t i = lo; i < hi; ++i) {
                const Aabb &b = boxes[ids[i]];
                tmp[i - lo] = {(b.hi[axis] + b.lo[axis]) * 0.5, ids[i]};  // centre = (p_max + p_min) * (1/2)
            }
            // The tie order among equal centroids is whatever libstdc++'s introsort yields for this sequence of
            // comparison outcomes; the reference sorts whole BBoxWithID values with the same comparator, so the
            // permutation is identical.
            std::sort(tmp.begin(), tmp.end(), [](const KeyId &a, const KeyId &b) { return a.key < b.key; });
            for (int64_t i = lo; i < hi; ++i) ids[i] = tmp[i - lo].id;
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

inline double half_area(const double *lo, const double *hi) {
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0;
    return dx * dy + dy * dz + dz * dx;
}

struct SahBuilder {
    static const int NBINS = 32;
    const Aabb *boxes;
    int32_t *ids;
    std::vector<TmpNode> nodes;
    std::atomic<int32_t> next{0};
    int max_leaf;
    double c_trav, c_isect;
    Forker fork;

    SahBuilder(const Aabb *b, int32_t *i, int64_t n, int ml, int threads)
        : boxes(b), ids(i), nodes((size_t)std::max<int64_t>(2 * n, 2)), max_leaf(ml), c_trav(1.0), c_isect(1.2), fork(threads) {
        if (const char *e = getenv("TAKE_SAH_CISECT")) c_isect = atof(e);  // tuning knob: cost of a leaf test relative to a node visit
    }

    int32_t alloc() { return next.fetch_add(1); }

    void make_leaf(TmpNode &n, int64_t lo, int64_t hi) {
        n.first = lo;
        n.count = (int32_t)(hi - lo);
    }

    // Run f(tid, a, b) over [lo, hi) split into `parts` contiguous chunks on their own threads (used only for the few huge
    // ranges at the top of the tree, where the recursion itself offers no parallelism yet).
    template <typename F>
    static void chunked(int64_t lo, int64_t hi, int parts, F f) {
        if (parts <= 1) { f(0, lo, hi); return; }
        std::vector<std::thread> pool;
        const int64_t m = hi - lo;
        for (int t = 1; t < parts; ++t) pool.emplace_back(f, t, lo + m * t / parts, lo + m * (t + 1) / parts);
        f(0, lo, lo + m / parts);
        for (auto &t : pool) t.join();
    }
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

    void build(int32_t node_id, int64_t lo, int64_t hi) {
        TmpNode &node = nodes[node_id];
        int64_t m = hi - lo;
        const int parts = m >= (1 << 18) ? std::min<int>(top_threads, (int)(m >> 16)) : 1;
        Aabb bb, cb;
        for (int a = 0; a < 3; ++a) { bb.lo[a] = cb.lo[a] = INFINITY; bb.hi[a] = cb.hi[a] = -INFINITY; }
        {
            std::vector<Aabb> pbb(parts, bb), pcb(parts, cb);
            chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                Aabb tb = pbb[tid], tc = pcb[tid];
                for (int64_t i = a0; i < a1; ++i) {
                    const Aabb &b = boxes[ids[i]];
                    for (int a = 0; a < 3; ++a) {
                        tb.lo[a] = std::min(tb.lo[a], b.lo[a]);
                        tb.hi[a] = std::max(tb.hi[a], b.hi[a]);
                        double c = 0.5 * (b.lo[a] + b.hi[a]);
                        tc.lo[a] = std::min(tc.lo[a], c);
                        tc.hi[a] = std::max(tc.hi[a], c);
                    }
                }
                pbb[tid] = tb; pcb[tid] = tc;
            });
            for (int t = 0; t < parts; ++t)
                for (int a = 0; a < 3; ++a) {
                    bb.lo[a] = std::min(bb.lo[a], pbb[t].lo[a]); bb.hi[a] = std::max(bb.hi[a], pbb[t].hi[a]);
                    cb.lo[a] = std::min(cb.lo[a], pcb[t].lo[a]); cb.hi[a] = std::max(cb.hi[a], pcb[t].hi[a]);
                }
        }
        node.box = bb;
        if (m == 1) { make_leaf(node, lo, hi); return; }

        // bin all three axes in one pass (per-thread bins for the huge ranges, merged afterwards)
        double scale3[3];
        bool axis_ok[3];
        for (int x = 0; x < 3; ++x) {
            double ext = cb.hi[x] - cb.lo[x];
            axis_ok[x] = ext > 0;
            scale3[x] = axis_ok[x] ? NBINS / ext : 0.0;
        }
        std::vector<Bins> pbins(parts);
        chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
            Bins &B = pbins[tid];
            B.clear();
            for (int64_t i = a0; i < a1; ++i) {
                const Aabb &b = boxes[ids[i]];
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
        double best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        double parent_area = half_area(bb.lo, bb.hi);
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
        int64_t mid = -1;
        if (best_axis >= 0) {
            double split_cost = c_trav + c_isect * best_cost / (parent_area > 0 ? parent_area : 1.0);
            double leaf_cost = c_isect * (double)m;
            if (m <= max_leaf && leaf_cost <= split_cost) { make_leaf(node, lo, hi); return; }
            double ext = cb.hi[best_axis] - cb.lo[best_axis];
            double scale = NBINS / ext;
            auto goes_left = [&](int32_t id) {
                const Aabb &b = boxes[id];
                int k = (int)((0.5 * (b.lo[best_axis] + b.hi[best_axis]) - cb.lo[best_axis]) * scale);
                k = std::min(std::max(k, 0), NBINS - 1);
                return k <= best_bin;
            };
            if (parts > 1) {
                // chunk-wise counting partition through a scratch copy (which side a primitive lands on is all that
                // matters; the order inside a side only permutes leaf slots)
                std::vector<int64_t> nleft(parts + 1, 0);
                chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                    int64_t c = 0;
                    for (int64_t i = a0; i < a1; ++i) c += goes_left(ids[i]) ? 1 : 0;
                    nleft[tid + 1] = c;
                });
                for (int t = 0; t < parts; ++t) nleft[t + 1] += nleft[t];
                const int64_t total_left = nleft[parts];
                std::vector<int32_t> scratch(ids + lo, ids + hi);
                chunked(lo, hi, parts, [&](int tid, int64_t a0, int64_t a1) {
                    int64_t l = lo + nleft[tid], r = lo + total_left + (a0 - lo) - nleft[tid];
                    for (int64_t i = a0; i < a1; ++i) {
                        const int32_t id = scratch[i - lo];
                        if (goes_left(id)) ids[l++] = id; else ids[r++] = id;
                    }
                });
                mid = lo + total_left;
            } else {
                int32_t *p = std::partition(ids + lo, ids + hi, goes_left);
                mid = p - ids;
            }
        }
        if (mid <= lo || mid >= hi) {
            // all centroids coincide (or binning failed): leaf if allowed, otherwise split the range in half
            if (m <= max_leaf) { make_leaf(node, lo, hi); return; }
            mid = lo + m / 2;
        }
        int32_t l = alloc(), r = alloc();
        nodes[node_id].left = l;
        nodes[node_id].right = r;
        fork.both(m > 32768, [&] { build(l, lo, mid); }, [&] { build(r, mid, hi); });
    }
};

// Conservative FP32 image of an FP64 bound: round outward, move out by `pad`, and one more ulp to absorb the
// rounding of that subtraction / addition.
inline float round_down(double v, float pad) {
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return nextafterf(f - pad, -INFINITY);
}
inline float round_up(double v, float pad) {
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return nextafterf(f + pad, INFINITY);
}

inline int32_t leaf_code(int64_t first, int32_t count) { return ~(int32_t)((first << 3) | (int64_t)(count - 1)); }

struct Flattener {
    const std::vector<TmpNode> &tmp;
    std::vector<FastNode> &out;
    float pad;
    int depth = 0;

    void set_child(FastNode &n, int which, const TmpNode *c, int32_t idx, int32_t cnt) {
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

    // iterative pre-order flatten (explicit stack: no recursion depth limits on degenerate trees)
    void run(int32_t root) {
        struct Item { int32_t tmp_id, out_id, depth; };
        std::vector<Item> stack;
        out.clear();
        out.reserve(tmp.size() / 2 + 2);
        const TmpNode &r = tmp[root];
        out.emplace_back();
        if (r.count > 0 || r.left < 0) {  // root is a leaf (or the tree is empty): wrap it
            memset(&out[0], 0, sizeof(FastNode));
            if (r.count > 0) set_child(out[0], 0, &r, leaf_code(r.first, r.count), r.count);
            else set_child(out[0], 0, nullptr, ~0, 0);
            set_child(out[0], 1, nullptr, ~0, 0);
            depth = 1;
            return;
        }
        stack.push_back({root, 0, 1});
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            depth = std::max(depth, it.depth);
            const TmpNode &t = tmp[it.tmp_id];
            const TmpNode *kids[2] = {&tmp[t.left], &tmp[t.right]};
            int32_t kid_ids[2] = {t.left, t.right};
            for (int k = 0; k < 2; ++k) {
                const TmpNode *c = kids[k];
                if (c->count > 0) {
                    set_child(out[it.out_id], k, c, leaf_code(c->first, c->count), c->count);
                } else {
                    int32_t id = (int32_t)out.size();
                    out.emplace_back();
                    set_child(out[it.out_id], k, c, id, 0);
                    stack.push_back({kid_ids[k], id, it.depth + 1});
                }
            }
        }
    }
};

// Collapse the binary SAH tree into 4-wide nodes: start from a binary node's two children and keep replacing the inner
// child with the largest surface area by its own two children until there are four (or only leaves are left).
struct WideFlattener {
    const std::vector<TmpNode> &tmp;
    std::vector<WideNode> &out;
    float pad;
    int depth = 0;

    void set_child(WideNode &n, int k, const TmpNode *c, int32_t idx, int32_t cnt) {
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

    void run(int32_t root) {
        struct Item { int32_t tmp_id, out_id, depth; };
        std::vector<Item> stack;
        out.clear();
        out.reserve(tmp.size() / 3 + 2);
        const TmpNode &r = tmp[root];
        out.emplace_back();
        if (r.count > 0 || r.left < 0) {  // root is a leaf or the tree is empty
            memset(&out[0], 0, sizeof(WideNode));
            for (int k = 0; k < 4; ++k) set_child(out[0], k, nullptr, TAKE_WIDE_EMPTY, 0);
            if (r.count > 0) set_child(out[0], 0, &r, leaf_code(r.first, r.count), r.count);
            depth = 1;
            return;
        }
        stack.push_back({root, 0, 1});
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            depth = std::max(depth, it.depth);
            int32_t kids[4];
            int nk = 0;
            kids[nk++] = tmp[it.tmp_id].left;
            kids[nk++] = tmp[it.tmp_id].right;
            while (nk < 4) {
                int best = -1;
                double best_area = -1;
                for (int k = 0; k < nk; ++k) {
                    const TmpNode &c = tmp[kids[k]];
                    if (c.count > 0) continue;  // a leaf stays a leaf
                    double a = half_area(c.box.lo, c.box.hi);
                    if (a > best_area) { best_area = a; best = k; }
                }
                if (best < 0) break;
                const TmpNode &c = tmp[kids[best]];
                kids[best] = c.left;
                kids[nk++] = c.right;
            }
            for (int k = 0; k < 4; ++k) {
                if (k >= nk) { set_child(out[it.out_id], k, nullptr, TAKE_WIDE_EMPTY, 0); continue; }
                const TmpNode *c = &tmp[kids[k]];
                if (c->count > 0) {
                    set_child(out[it.out_id], k, c, leaf_code(c->first, c->count), c->count);
                } else {
                    int32_t id = (int32_t)out.size();
                    out.emplace_back();
                    set_child(out[it.out_id], k, c, id, 0);
                    stack.push_back({kids[k], id, it.depth + 1});
                }
            }
        }
    }
};

}  // namespace

void build_reference_tree(const Aabb *boxes, int64_t n, int threads, RefTree &out) {
    out.nodes.clear();
    out.dfs_rank.clear();
    out.root = -1;
    if (n <= 0) return;
    out.nodes.resize((size_t)(2 * n - 1));
    std::vector<int32_t> ids((size_t)n);
    for (int64_t i = 0; i < n; ++i) ids[i] = (int32_t)i;
    RefBuilder b{boxes, out.nodes.data(), ids.data(), Forker(threads)};
    out.root = b.build(0, n, 0);
    // After the recursion ids[] lists the primitives in left-to-right leaf order.
    out.dfs_rank.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) out.dfs_rank[ids[i]] = (int32_t)i;
}

void build_fast_tree(const Aabb *boxes, int64_t n, int max_leaf, float pad, int threads, FastTree &out) {
    out.nodes.clear();
    out.leaf_prims.assign((size_t)std::max<int64_t>(n, 0), 0);
    for (int64_t i = 0; i < n; ++i) out.leaf_prims[i] = (int32_t)i;
    SahBuilder b(boxes, out.leaf_prims.data(), n, std::max(1, max_leaf), threads);
    b.top_threads = std::max(1, threads);
    int32_t root = b.alloc();
    if (n > 0) b.build(root, 0, n);
    Flattener f{b.nodes, out.nodes, pad};
    f.run(root);
    out.depth = f.depth;
    WideFlattener wf{b.nodes, out.wide, pad};
    wf.run(root);
    out.wide_depth = wf.depth;
    // SAH cost of the final tree (diagnostic)
    double cost = 0, root_area = n > 0 ? half_area(b.nodes[root].box.lo, b.nodes[root].box.hi) : 0;
    if (root_area > 0) {
        int32_t used = b.next.load();
        for (int32_t i = 0; i < used; ++i) {
            const TmpNode &t = b.nodes[i];
            double a = half_area(t.box.lo, t.box.hi) / root_area;
            cost += t.count > 0 ? a * b.c_isect * t.count : a * b.c_trav;
        }
    }
    out.sah_cost = cost;
}

}  // namespace take

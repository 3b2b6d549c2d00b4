// Host-side acceleration-structure builders (replace build_bvh / construct_bvh of the reference:
// src/scene.cpp:4-23, src/bvh.cpp:8-45).  Two trees are built from the same primitive boxes:
//
//  * the REFERENCE-ORDER tree: the reference's own topology (object-median split on the largest-extent axis,
//    libstdc++ std::sort by centroid, one primitive per leaf, post-order node numbering, root last) with FP64
//    boxes.  It serves the TAKE_ISECT_EXACT kernel and defines each primitive's DFS rank, which is how the
//    reference resolves equal-t hits (the later leaf in its left-to-right DFS wins, src/bvh.cpp:94-108).
//  * the FAST tree: binned-SAH BVH2, up to `max_leaf` primitives per leaf, flattened into 64-byte nodes that
//    hold BOTH children's boxes as conservative FP32 (rounded outward + padded), fetched as four float4.
#pragma once
#include <stdint.h>

#include <vector>

namespace take {

struct Aabb {
    double lo[3], hi[3];
};

struct RefNode {  // 64 bytes; mirrors BVHNode (src/bvh.h:5-10)
    double lo[3], hi[3];
    int32_t left, right, prim, pad;
};

struct alignas(16) FastNode {  // 64 bytes = 4 x float4
    float c0lox, c0hix, c0loy, c0hiy;  // child 0: x and y slabs
    float c1lox, c1hix, c1loy, c1hiy;  // child 1: x and y slabs
    float c0loz, c0hiz, c1loz, c1hiz;  // both children: z slabs
    int32_t child0, child1;            // >= 0: inner node index; < 0: leaf, ~child = (first slot << 3) | (count - 1)
    int32_t count0, count1;            // leaf primitive counts again (diagnostics; 0 for inner children)
};
static_assert(sizeof(FastNode) == 64, "FastNode must be 64 bytes");

// 4-wide node, 128 bytes = 8 x float4: per-axis lo / hi planes of the four children (SoA, one float4 each), the four
// child links, and padding to a full 128-byte line.  Links as in FastNode; an unused child is TAKE_WIDE_EMPTY.
struct alignas(16) WideNode {
    float lox[4], hix[4], loy[4], hiy[4], loz[4], hiz[4];
    int32_t child[4];
    int32_t count[4];  // leaf primitive counts (diagnostics)
};
static_assert(sizeof(WideNode) == 128, "WideNode must be 128 bytes");
#define TAKE_WIDE_EMPTY 0x7fffffff

struct RefTree {
    std::vector<RefNode> nodes;
    int32_t root = -1;
    std::vector<int32_t> dfs_rank;  // per primitive: position in the reference's DFS leaf order
};

struct FastTree {
    std::vector<FastNode> nodes;     // node 0 is the root (always an inner node, possibly with an empty child 1)
    std::vector<WideNode> wide;      // the same tree collapsed to 4-wide nodes (node 0 is the root)
    std::vector<int32_t> leaf_prims; // primitive ids in leaf order; a leaf is a contiguous slot range
    int depth = 0, wide_depth = 0;
    double sah_cost = 0;
};

void build_reference_tree(const Aabb *boxes, int64_t n, int threads, RefTree &out);
// std::sort vs. its forked twin on synthetic keys: number of differing positions (0 = identical)
int64_t sort_selftest(int64_t n, int64_t distinct, int pattern, int threads, uint64_t seed);
// pad: absolute outward padding applied to every FP32 box after outward rounding.
// max_leaf <= 8 (the leaf code keeps count-1 in 3 bits); n < 2^28.
void build_fast_tree(const Aabb *boxes, int64_t n, int max_leaf, float pad, int threads, FastTree &out);

}  // namespace take

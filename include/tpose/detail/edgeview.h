// tpose/detail/edgeview.h -- the unique-edge view of a triangle list that the step kernels work on
// (host only; built by libtpose_cuda at upload, unit-tested on its own in tests/test_edgeview_cpu.py).
//
// Side j of triangle t runs from corner j to corner j+1 (the reference's halfedge 3t+j).  The two
// sides of an interior edge share one entry of `edges`, stored (first < second end point), so that
// k_edges integrates the edge once; a side that runs second -> first carries the "reversed" bit.
// `corners` lists, grouped by vertex, the (triangle, corner) pairs that sit on it: k_corners
// evaluates the variants that move corner j of triangle t in the warp of that vertex.
// Any triangle list is accepted: open meshes, edges shared by more than two triangles, repeated
// vertices inside a triangle (zero-length sides).
#ifndef TPOSE_DETAIL_EDGEVIEW_H
#define TPOSE_DETAIL_EDGEVIEW_H

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace tpose {
namespace edgeview {

constexpr std::int32_t kReversed = (std::int32_t)0x80000000u;

struct view {
    std::vector<std::int32_t> edges;       // 2 per edge: first, second end point (first <= second)
    std::vector<std::int32_t> tri_edges;   // 4 per triangle: edge id | kReversed for sides 0, 1, 2; [3] = 0
    std::vector<std::int32_t> corner_off;  // np + 1 offsets into corners
    std::vector<std::int32_t> corners;     // 3 * nt entries, triangle << 2 | corner
    int ne = 0;
};

// tri: nt x 4 ints (w ignored).  Returns false if a triangle references a vertex outside [0, np)
// (np must stay below 2^21, nt below 2^22 / 3: the limits of tpose_cuda_create).
inline bool build(const std::int32_t* tri, int nt, int np, view& out, std::vector<std::uint64_t>& keys)
{
    keys.resize((std::size_t)3 * nt);
    for (int t = 0; t < nt; t++) {
        const std::uint32_t v[4] = {(std::uint32_t)tri[4 * t], (std::uint32_t)tri[4 * t + 1], (std::uint32_t)tri[4 * t + 2], (std::uint32_t)tri[4 * t]};
        if (v[0] >= (std::uint32_t)np || v[1] >= (std::uint32_t)np || v[2] >= (std::uint32_t)np) return false;     // (negative indices wrap high)
        for (int j = 0; j < 3; j++) {
            const std::uint32_t x = v[j], y = v[j + 1];
            const std::uint32_t lo = x < y ? x : y, hi = x < y ? y : x;
            // (lo, hi) in the top 42 bits, the side slot below
            keys[(std::size_t)3 * t + j] = ((std::uint64_t)lo << 43) | ((std::uint64_t)hi << 22) | (std::uint64_t)(3 * t + j);
        }
    }
    // ascending keys.  A vertex has a handful of sides, so: counting sort on the low end point, insertion sort inside
    // each bucket -- the order std::sort gives, in linear time (3 000 triangles: 255 -> ~60 us; this runs on the host
    // after every topology edit of the schedule and at every level switch of a pair)
    {
        static thread_local std::vector<std::uint32_t> start;
        static thread_local std::vector<std::uint64_t> sorted;
        start.assign((std::size_t)np + 1, 0u);
        for (std::uint64_t k : keys) start[(std::size_t)(k >> 43) + 1]++;
        for (int v = 0; v < np; v++) start[(std::size_t)v + 1] += start[(std::size_t)v];
        sorted.resize(keys.size());
        for (std::uint64_t k : keys) sorted[start[(std::size_t)(k >> 43)]++] = k;       // start[v] ends up at the END of bucket v
        std::size_t b0 = 0;
        for (int v = 0; v < np; v++) {
            const std::size_t b1 = start[(std::size_t)v];
            for (std::size_t i = b0 + 1; i < b1; i++) {
                const std::uint64_t k = sorted[i];
                std::size_t j = i;
                for (; j > b0 && sorted[j - 1] > k; j--) sorted[j] = sorted[j - 1];
                sorted[j] = k;
            }
            b0 = b1;
        }
        keys.swap(sorted);
    }
    out.edges.resize((std::size_t)6 * nt);
    out.tri_edges.assign((std::size_t)4 * nt, 0);
    out.corner_off.assign((std::size_t)np + 1, 0);
    out.corners.resize((std::size_t)3 * nt);
    int ne = 0;
    std::uint64_t prev = ~0ull;
    for (std::size_t i = 0; i < keys.size(); i++) {
        const std::uint64_t pair = keys[i] >> 22;
        const int slot = (int)(keys[i] & 0x3FFFFFu), t = slot / 3, j = slot - 3 * t;
        if (pair != prev) {
            out.edges[(std::size_t)2 * ne] = (std::int32_t)(pair >> 21);
            out.edges[(std::size_t)2 * ne + 1] = (std::int32_t)(pair & 0x1FFFFFu);
            ne++; prev = pair;
        }
        const bool rev = tri[4 * t + j] > tri[4 * t + (j + 1) % 3];
        out.tri_edges[(std::size_t)4 * t + j] = (ne - 1) | (rev ? kReversed : 0);
    }
    out.ne = ne;
    out.edges.resize((std::size_t)2 * ne);
    // corners grouped by vertex (counting sort): gradient.cs sums a vertex's contributions, the
    // order inside a group is irrelevant (wrapping integer adds)
    for (int t = 0; t < nt; t++)
        for (int j = 0; j < 3; j++) out.corner_off[(std::size_t)tri[4 * t + j] + 1]++;
    for (int v = 0; v < np; v++) out.corner_off[(std::size_t)v + 1] += out.corner_off[(std::size_t)v];
    static thread_local std::vector<std::int32_t> fill;
    fill.assign(out.corner_off.begin(), out.corner_off.begin() + np);
    for (int t = 0; t < nt; t++)
        for (int j = 0; j < 3; j++) out.corners[(std::size_t)fill[(std::size_t)tri[4 * t + j]]++] = (t << 2) | j;
    return true;
}

}  // namespace edgeview
}  // namespace tpose

#endif

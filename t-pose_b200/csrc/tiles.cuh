// tiles.cuh -- the tile-streaming form of the edge integrals: every texel of the image crosses HBM
// once per step, as raw RGBA8, staged into shared memory by TMA.
//
// eval.cuh's k_edges gathers 16-byte row-prefix entries from a (W+1) x H table at the nine crossings
// of every (edge, row): a 33 MB table per 1080p image, touched in 32-byte sectors scattered over
// three quarters of it -- 3.5x the algorithmic bytes.  Here the same integrals
//      record(edge, move) = sign * sum_rows P[r][cross(r)]           (P = row prefix of the six moments)
// are accumulated TILE BY TILE:
//
//   k_build_coarse (once per image)  C[tx][r] = P[r][tx*64]: the prefix at the left border of every
//            64-pixel tile column, column-major so that one tile's slice is contiguous
//            (16 bytes per row and tile column: 0.5 MB at 1080p -- L2-resident).
//   k_setup  (per step, one thread per unique mesh edge)  the ten snapped end points of the edge's nine
//            moves (triangle.vs:58-84 -- same float32 expression tree as common.cuh::to_fixed), the
//            range of tiles the moves can cross, that range reduced over groups of 32 edges, and
//            zeroed records.  Its last block is the service block of k_edges (previous frame's geterr +
//            topology watch).
//   k_tiles  (per step, one CTA per 64 x TH tile)
//            1. one thread issues a TMA tensor load (cp.async.bulk.tensor.2d, UTMALDG) of the tile's raw
//               pixels and a bulk copy (UBLKCP) of its coarse slice, both signalled through one mbarrier;
//            2. meanwhile all threads cull the edge groups, then the edges of the surviving groups,
//               against the tile and stage the survivors' snapped end points in shared memory;
//            3. the tile-local row prefix is built in shared memory on top of the coarse slice: every
//               entry is the same packed 16-byte absolute prefix the table path reads from HBM;
//            4. one thread per (edge, move) item walks the item's rows inside the tile (EdgeWalk, exact
//               integer scan conversion) and sums the entries at the crossings this tile owns;
//            5. the partial records of an edge (9 moves x 6 x int64 = 432 contiguous bytes) are added to
//               the global records by ONE bulk reduction (cp.reduce.async.bulk .add.u64) per edge --
//               the TMA unit does the read-modify-write in L2, the SM issues no atomics.
//   k_corners (eval.cuh) consumes the records unchanged.
//
// Every pixel is read exactly once per step: HBM traffic = 4 B per pixel + the coarse slice (0.25 B)
// + the small per-edge arrays, i.e. the algorithmic bytes of SURVEY.md 8(d).
#pragma once
#include "eval.cuh"
#include "table.cuh"

namespace tpose_dev {

constexpr int kTileW = 64;              // pixels per tile row: 256 bytes, the TMA box's inner extent
constexpr int kTileThreads = 256;
constexpr int kTilePitch = kTileW + 4;  // entries per shared-memory table row (slot 64 = the prefix at x = 64; +3 pad
                                        // so that consecutive rows start 4 bank groups apart)
constexpr int kSetupThreads = 256;
constexpr int kListCap = 512;           // edges in a tile's work list; more than that: the tile walks edge-id chunks
constexpr int kGroupCap = 256;          // surviving edge groups per tile
constexpr int kEdgeGroup = 32;          // edges per cull group (one warp of k_setup)
constexpr int kBatchEdges = 28;         // edges per walk round: 28 * 9 = 252 items <= 256 threads
constexpr int kSnapPerEdge = 10;        // a, a+dx, a-dx, a+dy, a-dy, b, b+dx, ...  (int2 each)
constexpr unsigned kEmptyBox = 0x00FF00FFu;   // tx0 = 255 > tx1 = 0, ty0 = 255 > ty1 = 0

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// TMA: box (kTileW x TH) of the [H][W] uint32 image at pixel (x, y) -> shared memory; out-of-range texels read 0
__device__ __forceinline__ void tma_load_tile(unsigned dst, const void* tmap, int x, int y, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// global[0 .. bytes) += shared[0 .. bytes) as 64-bit integers, by the TMA unit
__device__ __forceinline__ void bulk_reduce_add_u64(void* gdst, unsigned ssrc, unsigned bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                 ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- k_build_coarse ---------------------------------------------------------------------------
// One block per image row: the packed prefix (table.cuh::part_pack) at every multiple of kTileW and
// at W, written column-major: coarse[tx * H + r], tx = 0 .. ntx (ntx = ceil(W / kTileW)).
__global__ void __launch_bounds__(kTableThreads)
k_build_coarse(const uchar4* __restrict__ img, int ipitch, int W, int H, uint4* __restrict__ coarse, int ntx)
{
    __shared__ Part s_warp[kTableThreads / 32];
    __shared__ Part s_carry;
    const int r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uchar4* row = img + (size_t)r * ipitch;
    if (tid == 0) { s_carry = Part{0, 0, 0, 0, 0}; coarse[r] = make_uint4(0, 0, 0, 0); }
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += kTableThreads) {
        const int x = x0 + tid;
        Part v = Part{0, 0, 0, 0, 0};
        if (x < W) v = part_of(__ldg(row + x));
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const Part u = part_shfl_up(v, d);
            if (lane >= d) v = part_add(v, u);
        }
        if (lane == 31) s_warp[wid] = v;
        __syncthreads();
        Part base = s_carry;
        for (int w2 = 0; w2 < wid; w2++) base = part_add(base, s_warp[w2]);
        v = part_add(v, base);
        // v = inclusive prefix through pixel x = exclusive prefix at column x + 1
        if (x < W && (((x + 1) % kTileW) == 0 || x + 1 == W)) {
            const int tx = (x + 1 + kTileW - 1) / kTileW;          // W itself goes to column ntx
            coarse[(size_t)tx * H + r] = part_pack(v);
        }
        __syncthreads();
        if (tid == kTableThreads - 1) s_carry = v;
        __syncthreads();
    }
    (void)ntx;
}

// ---- k_setup --------------------------------------------------------------------------------------
__device__ __forceinline__ int tile_col(int c, int ntx) { return min(c / kTileW, ntx - 1); }

template <int TH>
__global__ void __launch_bounds__(kSetupThreads)
k_setup(const EvalJob* __restrict__ jobs, int set, unsigned flags, double rel_tol, int frame)
{
    __shared__ EvalJob J;
    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(EvalJob) / sizeof(int)))
        reinterpret_cast<int*>(&J)[tid] = reinterpret_cast<const int*>(jobs + blockIdx.y)[tid];
    __syncthreads();
    if ((flags & EV_HONOUR_STOP) && __ldcg(&J.run_state[RS_STOP])) return;
    set ^= J.flip;
    if ((flags & EV_FINAL) && blockIdx.x != gridDim.x - 1) return;      // check-only evaluation: nothing to integrate
    if (blockIdx.x == gridDim.x - 1) { service_block(J, set, flags, rel_tol, frame, tid, kSetupThreads, gridDim.y == 1); return; }

    const Xform xf{J.ratio, 0.5f * (float)J.W, 0.5f * (float)J.H, 256 * J.H};
    const int ne_up = (J.NE + kEdgeGroup - 1) / kEdgeGroup * kEdgeGroup;      // whole warps: the group boxes are warp reductions
    for (int base = blockIdx.x * kSetupThreads; base < ne_up; base += (gridDim.x - 1) * kSetupThreads) {
        const int e = base + tid;
        unsigned tx0 = 255u, tx1 = 0u, ty0 = 255u, ty1 = 0u;
        if (e < J.NE) {
            const int2 ev = __ldg(J.edges + e);
            const float2 pa = J.pts[ev.x], pb = J.pts[ev.y];
            int2 s[kSnapPerEdge];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                float dx = 0.0f, dy = 0.0f;
                if (k) fd_offset(k - 1, J.dp, dx, dy);
                s[k] = to_fixed(pa, dx, dy, xf);
                s[5 + k] = to_fixed(pb, dx, dy, xf);
            }
            int4* out = reinterpret_cast<int4*>(J.esnap + (size_t)e * kSnapPerEdge);
#pragma unroll
            for (int k = 0; k < 5; k++) out[k] = make_int4(s[2 * k].x, s[2 * k].y, s[2 * k + 1].x, s[2 * k + 1].y);
            int xmin = s[0].x, xmax = s[0].x, ymin = s[0].y, ymax = s[0].y;
#pragma unroll
            for (int k = 1; k < kSnapPerEdge; k++) {
                xmin = min(xmin, s[k].x); xmax = max(xmax, s[k].x);
                ymin = min(ymin, s[k].y); ymax = max(ymax, s[k].y);
            }
            // rows with a pixel centre in [ymin, ymax), columns of the crossings (see EdgeWalk): both closed here
            const int r_lo = max((ymin - 128 + 255) >> 8, 0), r_hi = min((ymax - 128 + 255) >> 8, J.H) - 1;
            if (r_lo <= r_hi) {
                const int c_lo = min(max((xmin + 127) >> 8, 0), J.W), c_hi = min(max((xmax + 127) >> 8, 0), J.W);
                tx0 = (unsigned)tile_col(c_lo, J.ntx); tx1 = (unsigned)tile_col(c_hi, J.ntx);
                ty0 = (unsigned)(r_lo / TH); ty1 = (unsigned)(r_hi / TH);
            }
            J.ebox[e] = tx0 | (tx1 << 8) | (ty0 << 16) | (ty1 << 24);
        }
        const unsigned g0 = __reduce_min_sync(0xffffffffu, tx0), g1 = __reduce_max_sync(0xffffffffu, tx1);
        const unsigned g2 = __reduce_min_sync(0xffffffffu, ty0), g3 = __reduce_max_sync(0xffffffffu, ty1);
        if ((tid & 31) == 0 && base + tid < ne_up) J.gbox[(base + tid) / kEdgeGroup] = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
        // zero this block's records (coalesced): k_tiles adds into them
        const int e_first = base, e_last = min(base + kSetupThreads, J.NE);
        if (e_first < e_last) {
            uint4* z = reinterpret_cast<uint4*>(J.rec + (size_t)e_first * kEdgeMoves);
            const int n16 = (e_last - e_first) * kEdgeMoves * (int)(sizeof(EdgeRec) / 16);
            for (int i = tid; i < n16; i += kSetupThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// ---- k_tiles --------------------------------------------------------------------------------------

__device__ __forceinline__ bool box_hits(unsigned box, unsigned tx, unsigned ty)
{
    return (box & 255u) <= tx && tx <= ((box >> 8) & 255u) && ((box >> 16) & 255u) <= ty && ty <= (box >> 24);
}

// slot of the prefix at tile column x (0 .. 64) inside a shared-memory table row: a permutation inside
// every 16-entry segment (by the segment number), so that the four segment threads of a row -- and the
// two rows a quarter warp covers, kTilePitch apart -- store to different bank groups
__device__ __forceinline__ int tile_slot(int x) { return x ^ ((x >> 4) & 3); }

template <typename REM>
__device__ __forceinline__ void walk_tile_rows(int nrows, int q, const int dq, REM rem, const REM den, const REM dr,
                                               const uint4* __restrict__ ent, const int x0, const int xlim, const int W,
                                               unsigned& n, unsigned& odd, unsigned& sr, unsigned& sg, unsigned& sb, unsigned long long& s2)
{
#pragma unroll 2
    for (int i = 0; i < nrows; i++) {
        const int c = min(max(q, 0), W);
        const int x = c - x0;
        if ((unsigned)x < (unsigned)xlim) {                 // this tile owns the crossing
            const uint4 e = ent[tile_slot(x)];
            n += (unsigned)c; odd += e.x >> 20; sr += e.x & 0xFFFFFu; sg += e.y; sb += e.z; s2 += e.w;
        }
        ent += kTilePitch;
        q += dq; rem += dr;
        if (rem >= den) { rem -= den; ++q; }
    }
}

struct TileSmem {
    // (the dynamic shared memory of k_tiles, carved by hand: see tile_smem_bytes)
};

template <int TH>
__host__ __device__ constexpr int tile_raw_bytes() { return kTileW * TH * 4; }
template <int TH>
__host__ __device__ constexpr int tile_stage_bytes() { return kBatchEdges * kEdgeMoves * (int)sizeof(EdgeRec); }
template <int TH>
__host__ __device__ constexpr int tile_region0_bytes()       // raw pixels, later the staged partial records
{
    return tile_raw_bytes<TH>() > tile_stage_bytes<TH>() ? tile_raw_bytes<TH>() : tile_stage_bytes<TH>();
}
template <int TH>
__host__ __device__ constexpr int tile_smem_bytes()
{
    return tile_region0_bytes<TH>()                       // raw tile | staging
         + TH * kTilePitch * 16                           // prefix entries
         + TH * 16                                        // coarse slice
         + kBatchEdges * kSnapPerEdge * 8                 // snapped end points of the batch
         + kListCap * 4 + kGroupCap * 4                   // work list, surviving groups
         + 128;                                           // alignment slack
}

template <int TH>
__global__ void __launch_bounds__(kTileThreads, 2)
k_tiles(const EvalJob* __restrict__ jobs, unsigned flags)
{
    static_assert(kTileW * TH % kTileThreads == 0, "whole segments per thread");
    constexpr int SEG = kTileW * TH / kTileThreads;          // pixels per scan thread
    static_assert(SEG == 16, "the slot permutation and the segment scan are written for 16-pixel segments (TH = 64)");
    constexpr int TPR = kTileW / SEG;                        // scan threads per row

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ EvalJob J;
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ int s_ngroups, s_nlist, s_nz[kBatchEdges];

    const int tid = threadIdx.x;
    if (tid < (int)(sizeof(EvalJob) / sizeof(int)))
        reinterpret_cast<int*>(&J)[tid] = reinterpret_cast<const int*>(jobs + blockIdx.y)[tid];
    if (tid == 0) { s_ngroups = 0; s_nlist = 0; }
    __syncthreads();
    if ((flags & EV_HONOUR_STOP) && __ldcg(&J.run_state[RS_STOP])) return;
    if (flags & EV_FINAL) return;
    const int ntiles = J.ntx * J.nty;
    if ((int)blockIdx.x >= ntiles) return;                   // (a batch launches the largest job's grid)
    const int ty = (int)blockIdx.x / J.ntx, tx = (int)blockIdx.x - ty * J.ntx;
    const int x0 = tx * kTileW, y0 = ty * TH;
    const int rows_here = min(TH, J.H - y0);

    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uchar4* s_px = reinterpret_cast<uchar4*>(base);                                   // region 0: raw tile ...
    EdgeRec* s_stage = reinterpret_cast<EdgeRec*>(base);                              // ... then the partial records
    uint4* s_ent = reinterpret_cast<uint4*>(base + tile_region0_bytes<TH>());
    uint4* s_coarse = s_ent + TH * kTilePitch;
    int2* s_snap = reinterpret_cast<int2*>(s_coarse + TH);
    int* s_list = reinterpret_cast<int*>(s_snap + kBatchEdges * kSnapPerEdge);
    int* s_groups = s_list + kListCap;

    // 1. the tile's pixels and coarse slice are on their way while the edges are culled
    const unsigned bar = smem_u32(&s_bar);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (unsigned)(tile_raw_bytes<TH>() + rows_here * 16));
        tma_load_tile(smem_u32(s_px), J.tmap, x0, y0, bar);
        bulk_load(smem_u32(s_coarse), J.coarse + ((size_t)tx * J.H + y0), (unsigned)rows_here * 16u, bar);
    }

    // 2. cull: groups of 32 edges first, then the edges of the groups that survive
    const int ngroups_all = (J.NE + kEdgeGroup - 1) / kEdgeGroup;
    bool chunked = J.NE <= kListCap;                         // small meshes: every edge is tested directly, no overflow possible
    if (!chunked) {
        for (int g = tid; g < ngroups_all; g += kTileThreads)
            if (box_hits(__ldg(J.gbox + g), (unsigned)tx, (unsigned)ty)) {
                const int at = atomicAdd(&s_ngroups, 1);
                if (at < kGroupCap) s_groups[at] = g;
            }
        __syncthreads();
        const int ng = s_ngroups;
        if (ng > kGroupCap) chunked = true;
        else {
            for (int p = tid; p < ng * kEdgeGroup; p += kTileThreads) {
                const int e = s_groups[p / kEdgeGroup] * kEdgeGroup + (p & (kEdgeGroup - 1));
                if (e < J.NE && box_hits(__ldg(J.ebox + e), (unsigned)tx, (unsigned)ty)) {
                    const int at = atomicAdd(&s_nlist, 1);
                    if (at < kListCap) s_list[at] = e;
                }
            }
            __syncthreads();
            if (s_nlist > kListCap) chunked = true;
        }
    }

    // 3. row prefix of the tile on top of the coarse slice
    mbar_wait(bar, 0);
    {
        const int row = tid / TPR, seg = tid % TPR;
        const uint4* px4 = reinterpret_cast<const uint4*>(s_px + row * kTileW + seg * SEG);
        unsigned pr = 0, pg = 0, pb = 0, ps = 0, po = 0;
        unsigned pxs[SEG];
#pragma unroll
        for (int i = 0; i < SEG / 4; i++) {
            const uint4 v = px4[i];
            pxs[4 * i] = v.x & 0x00FFFFFFu; pxs[4 * i + 1] = v.y & 0x00FFFFFFu; pxs[4 * i + 2] = v.z & 0x00FFFFFFu; pxs[4 * i + 3] = v.w & 0x00FFFFFFu;
        }
#pragma unroll
        for (int i = 0; i < SEG; i++) {
            const unsigned p = pxs[i];
            pr += p & 255u; pg += (p >> 8) & 255u; pb += p >> 16;
            ps = __dp4a(p, p, ps);
            po += __dp4a(p, 0x00010101u, 0u) & 1u;
        }
        // exclusive scan of the segment totals over the TPR threads of this row (adjacent lanes)
        unsigned er = pr, eg = pg, eb = pb, es = ps, eo = po;
#pragma unroll
        for (int d = 1; d < TPR; d <<= 1) {
            const unsigned ur = __shfl_up_sync(0xffffffffu, er, d), ug = __shfl_up_sync(0xffffffffu, eg, d);
            const unsigned ub = __shfl_up_sync(0xffffffffu, eb, d), us = __shfl_up_sync(0xffffffffu, es, d);
            const unsigned uo = __shfl_up_sync(0xffffffffu, eo, d);
            if (seg >= d) { er += ur; eg += ug; eb += ub; es += us; eo += uo; }
        }
        const uint4 c = (row < rows_here) ? s_coarse[row] : make_uint4(0u, 0u, 0u, 0u);
        // running absolute prefix in the packed entry format (common.cuh::Moments): x = sum R | odd << 20
        unsigned ax = c.x + (er - pr) + ((eo - po) << 20), ay = c.y + (eg - pg), az = c.z + (eb - pb), aw = c.w + (es - ps);
        uint4* out = s_ent + row * kTilePitch;
#pragma unroll
        for (int i = 0; i < SEG; i++) {
            out[tile_slot(seg * SEG + i)] = make_uint4(ax, ay, az, aw);
            const unsigned p = pxs[i];
            ax += (p & 255u) + ((__dp4a(p, 0x00010101u, 0u) & 1u) << 20);
            ay += (p >> 8) & 255u; az += p >> 16;
            aw = __dp4a(p, p, aw);
        }
        if (seg == TPR - 1) out[kTileW] = make_uint4(ax, ay, az, aw);
    }
    __syncthreads();        // entries complete; the raw pixels are dead (region 0 becomes the staging area)

    // 4. + 5. walk the work list in rounds of kBatchEdges edges
    // crossings this tile owns: x0 <= c < x0 + 64, and c == W as well where the image ends exactly on this tile's right border
    const int xlim = kTileW + ((tx == J.ntx - 1 && x0 + kTileW == J.W) ? 1 : 0);
    const int row_end = y0 + rows_here;
    const int nchunks = chunked ? (J.NE + kListCap - 1) / kListCap : 1;
    bool pending = false;                                    // this thread has a bulk reduction in flight that reads the staging area
    for (int chunk = 0; chunk < nchunks; chunk++) {
        if (chunked) {
            // fallback work list: the edges of one id range, tested directly
            if (tid == 0) s_nlist = 0;
            __syncthreads();
            for (int e = chunk * kListCap + tid; e < min((chunk + 1) * kListCap, J.NE); e += kTileThreads)
                if (box_hits(__ldg(J.ebox + e), (unsigned)tx, (unsigned)ty)) s_list[atomicAdd(&s_nlist, 1)] = e;
            __syncthreads();
        }
        const int nlist = s_nlist;
        for (int first = 0; first < nlist; first += kBatchEdges) {
            const int nb = min(kBatchEdges, nlist - first);
            // the staging area and the snap cache are reused: the previous round's reductions must have read them
            if (pending) { bulk_wait_read(); pending = false; }
            __syncthreads();
            for (int i = tid; i < nb * (kSnapPerEdge / 2); i += kTileThreads) {
                const int el = i / (kSnapPerEdge / 2), k = i - el * (kSnapPerEdge / 2);
                reinterpret_cast<int4*>(s_snap)[i] = __ldg(reinterpret_cast<const int4*>(J.esnap + (size_t)s_list[first + el] * kSnapPerEdge) + k);
            }
            if (tid < nb) s_nz[tid] = 0;
            __syncthreads();
            if (tid < nb * kEdgeMoves) {
                const int el = tid / kEdgeMoves, q = tid - el * kEdgeMoves;
                const int2 a = s_snap[el * kSnapPerEdge + ((q >= 1 && q <= 4) ? q : 0)];
                const int2 b = s_snap[el * kSnapPerEdge + 5 + ((q >= 5) ? q - 4 : 0)];
                EdgeWalk w;
                w.init_rows(a, b, y0, row_end);
                unsigned n = 0, odd = 0, sr = 0, sg = 0, sb = 0; unsigned long long s2 = 0;
                if (!w.done()) {
                    const uint4* ent = s_ent + (w.r - y0) * kTilePitch;
                    if (w.den < (1ll << 30))
                        walk_tile_rows<int>(w.r_end - w.r, w.q, w.dq, (int)w.rem, (int)w.den, (int)w.dr, ent, x0, xlim, J.W, n, odd, sr, sg, sb, s2);
                    else
                        walk_tile_rows<long long>(w.r_end - w.r, w.q, w.dq, w.rem, w.den, w.dr, ent, x0, xlim, J.W, n, odd, sr, sg, sb, s2);
                }
                const long long sgn = w.sign;
                EdgeRec rec;
                rec.r = sgn * (long long)sr; rec.g = sgn * (long long)sg; rec.b = sgn * (long long)sb;
                rec.s2 = sgn * (long long)s2; rec.n = sgn * (long long)n; rec.odd = sgn * (long long)odd;
                s_stage[tid] = rec;
                if (n | odd | sr | sg | sb | (unsigned)(s2 != 0ull)) s_nz[el] = 1;
            }
            fence_async_smem();          // the staged records must be visible to the TMA unit (async proxy)
            __syncthreads();
            // the thread of an edge's first item issues the edge's reduction (the instruction takes uniform
            // operands, so a warp issues its lanes one by one: spread them over the warps)
            if (tid < nb * kEdgeMoves && tid % kEdgeMoves == 0 && s_nz[tid / kEdgeMoves]) {
                const int el = tid / kEdgeMoves;
                bulk_reduce_add_u64(J.rec + (size_t)s_list[first + el] * kEdgeMoves, smem_u32(s_stage + el * kEdgeMoves),
                                    (unsigned)(kEdgeMoves * sizeof(EdgeRec)));
                bulk_commit();
                pending = true;
            }
        }
    }
    if (pending) bulk_wait_read();       // shared memory must outlive the reads of the reductions
}

}  // namespace tpose_dev

// capi.cu -- the C ABI of include/tpose_cuda.h over the sm_100a kernels.
// There is no CPU path in this library: without a CUDA device every entry point fails.
#include "../../include/tpose_cuda.h"
#include "../../include/tpose/detail/edgeview.h"
#include "carry.cuh"
#include "eval.cuh"
#include "mview.cuh"
#include "raster.cuh"
#include "rgbd.cuh"
#include "table.cuh"
#include "tiles.cuh"

#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <vector>

using namespace tpose_dev;

// Launch geometry of one evaluation.  Grid sizes are quantised (power of two, at most one full
// wave) and the kernels stride, so a captured sequence stays valid while the mesh grows.
// path: 0 = row-prefix table gathers (k_edges), 1 = TMA-staged tiles (k_setup + k_tiles)
struct EvalShape { int group, grid_e, grid_t; bool narrow; int path, grid_s, grid_tiles; int grid_b; bool dense = false; };
constexpr int kPathTable = 0, kPathTiles = 1;
constexpr int kTileH = 64;
// a sequence of evaluations (see issue_sequence) and its instantiated graph
struct SeqKey { const EvalJob* jobs; int njobs; EvalShape shape; int kind, len, first_set; unsigned flags; double tol; Early early = Early{}; };
struct SeqGraph { SeqKey key; cudaGraphExec_t exec; uint64_t stamp; };
// one instantiated frame (tpose_cuda_frame with device-visible caller buffers: see frame_mapped)
struct FrameGraph { const void* in; void* ten; void* cn; void* ca; void* out; int set, nt, np, slot, flavour; EvalShape shape; const EvalJob* job; cudaGraphExec_t exec; };

struct ImageSlot {
    uchar4* rgba = nullptr;    // raster-resolution image on the device, rows ipitch pixels apart
    uint4* table = nullptr;    // table path: row-prefix moments, (W+1) entries per row
    uint4* coarse = nullptr;   // tile path: prefix at every tile column's left border, [ntx+1][H]
    void* tmap = nullptr;      // tile path: CUtensorMap of rgba in device memory (128 bytes)
    int W = 0, H = 0, pitch = 0, ipitch = 0, ntx = 0, nty = 0;
    size_t rgba_cap = 0, table_cap = 0, coarse_cap = 0;
    bool ready = false;
};

struct tpose_cuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int maxT = 0, maxP = 0;
    int NT = 0, NP = 0;
    float ratio = 1.5f;
    float step_scale = 1.0f;          // multiplies shift.cs's step constant (1 = the reference)
    bool have_mesh = false;
    ImageSlot img[2];
    uchar4* staging = nullptr; size_t staging_cap = 0;   // source-resolution upload buffer
    // pinned ring for small uploads (points, triangles, colours): the caller's buffer is copied
    // here and the H2D copy is asynchronous, so an upload does not stall the stream
    static constexpr int kStageSlots = 4;
    unsigned char* h_stage[kStageSlots] = {}; size_t h_stage_cap = 0; cudaEvent_t h_stage_ev[kStageSlots] = {}; int h_stage_next = 0;
    float2* pts = nullptr;
    int4* tri = nullptr;
    int* cn[2] = {nullptr, nullptr}; int4* ca[2] = {nullptr, nullptr}; int* ten[2] = {nullptr, nullptr};
    int cur = 0;                      // buffer set holding the last committed evaluation
    int2* gr = nullptr;
    // unique-edge view of the uploaded triangles (rebuilt on the host whenever they change)
    std::vector<int32_t> tri_host;    // mirror of the index buffer, 4 ints per triangle
    int2* edges = nullptr; int4* tri_edges = nullptr; int* corner_off = nullptr; int* corners = nullptr;
    EdgeRec* rec = nullptr;
    int2* esnap = nullptr; unsigned* ebox = nullptr; unsigned* gbox = nullptr;   // tile path, per step
    int path = kPathTable;            // measured faster on B200; TPOSE_CUDA_PATH=tiles selects the TMA tiles
    int NE = 0, edges_np = 0;
    bool edges_dirty = true;
    std::vector<uint64_t> scratch_keys; tpose::edgeview::view edge_view;
    int4* wtri = nullptr; int* whe = nullptr; unsigned* wmask = nullptr; bool watching = false;
    // run_state | err_state | wmask | pts share one allocation ("report"), so that what a run's caller reads
    // next -- verdict, watch mask, points -- comes back in a single copy
    unsigned char* report = nullptr; size_t report_pts_off = 0; unsigned char* h_report = nullptr;
    int* d_h_report = nullptr; int* d_h_hist[4] = {nullptr, nullptr, nullptr, nullptr};   // device aliases of h_report / h_hist_* (k_run_out)
    bool run_alias_known = false, hist_alias_known = false;
    unsigned char* wbuf = nullptr;    // wtri | whe, packed at the current NT: one upload
    unsigned* h_run_mask = nullptr; bool run_mask_valid = false;   // watch mask of the frame the last run stopped on
    // per-frame history of tpose_cuda_run (tpose_cuda_set_history): device arrays + pinned host copies
    int hist_cap = 0, hist_frames = 0, hist_np = 0, hist_nt = 0, hist_np_frames = 0;   // hist_np_frames: frames the last run recorded
    float2* d_hist_pts = nullptr; int* d_hist_ten = nullptr; float* d_hist_err = nullptr; unsigned* d_hist_mask = nullptr;
    int2* d_hist_gr = nullptr;
    float2* h_hist_pts = nullptr; int* h_hist_ten = nullptr; float* h_hist_err = nullptr; unsigned* h_hist_mask = nullptr;
    unsigned* ticket = nullptr;       // [1] topology trigger accumulator ([0] spare)
    float* err_state = nullptr;       // [0] toterr [1] |relerr| [2] newerr [3] converged flag
    int* run_state = nullptr;         // RS_STOP, RS_STEPS, RS_TRIG
    int4* geterr_scratch = nullptr;   // kGeterrScratchBytes (eval.cuh, geterr_block)
    EvalJob* job_dev = nullptr;       // this context's job record (single launches)
    EvalJob job_host{};               // what job_dev holds
    bool job_valid = false;
    EvalJob* batch_jobs = nullptr; int batch_cap = 0;   // job array when this context leads a batch
    int* ids = nullptr; size_t ids_cap = 0;
    unsigned char* flush = nullptr; size_t flush_cap = 0;
    unsigned char* list_buf = nullptr; size_t list_cap = 0;     // tpose_cuda_eval_list scratch
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;   // ev_mid: between k_edges / k_tiles and k_corners of a single step
    cudaEvent_t ev_mid0 = nullptr;    // tile path: between k_setup and k_tiles
    cudaEvent_t ev_batch = nullptr;   // end of the last batch this context led (the other jobs' streams wait on it)
    float last_ms = 0.0f;
    int64_t launches = 0;
    int group = 0;                    // lanes per edge (0 = pick from NT and H)
    std::vector<SeqGraph> graphs; uint64_t graph_clock = 0; bool use_graphs = true;
    std::vector<FrameGraph> frame_graphs;      // tpose_cuda_frame with pinned caller buffers
    const void* frame_host[5] = {(const void*)1, nullptr, nullptr, nullptr, nullptr}; void* frame_alias[5] = {};   // last pointer set seen and its device aliases
    bool mid_valid = false, time_kernels = false;
    // read-back cache filled by tpose_cuda_run in the same synchronisation as its verdict: the points and
    // toterr as the run left them (what the caller's next Buffer::retrieve / get_toterr would fetch)
    float* h_run_pts = nullptr; bool run_cache_valid = false; float run_toterr = 0.0f;
    std::string err;
};

static std::string g_create_error;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            char b_[512];                                                                 \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            ctx->err = b_;                                                                \
            return TPOSE_ERR_CUDA;                                                        \
        }                                                                                 \
    } while (0)

static int fail(tpose_cuda_ctx* ctx, int code, const char* msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

// Host -> device copy of a caller-owned buffer that the caller may reuse as soon as we return:
// through the pinned ring when it fits (asynchronous), otherwise synchronous.
static int upload_bytes(tpose_cuda_ctx* ctx, void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) return TPOSE_OK;
    ctx->run_cache_valid = false;
    if (bytes <= ctx->h_stage_cap) {
        const int slot = ctx->h_stage_next;
        ctx->h_stage_next = (slot + 1) % tpose_cuda_ctx::kStageSlots;
        CK(cudaEventSynchronize(ctx->h_stage_ev[slot]));          // the copy that last used this slot is done
        memcpy(ctx->h_stage[slot], src, bytes);
        CK(cudaMemcpyAsync(dst, ctx->h_stage[slot], bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->h_stage_ev[slot], ctx->stream));
        return TPOSE_OK;
    }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" const char* tpose_cuda_version(void)
{
    return "tpose-cuda 0.3 (sm_100a; shared-edge integrals over row-prefix moment tables or TMA-staged image tiles)";
}

extern "C" const char* tpose_cuda_last_error(const tpose_cuda_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int tpose_cuda_create(int device, int max_triangles, int max_points, void* stream, tpose_cuda_ctx** out)
{
    if (!out || max_triangles <= 0 || max_points < 3 || max_points > (1 << 21) || 13ll * max_triangles > TPOSE_REF_MAXT) {
        g_create_error = "tpose_cuda_create: bad arguments (need 13*max_triangles <= 2<<18, max_points <= 2<<20)";
        return TPOSE_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        (void)cudaGetLastError();
        g_create_error = "tpose_cuda_create: no usable CUDA device (this library has no CPU fallback)";
        return TPOSE_ERR_NODEVICE;
    }
    tpose_cuda_ctx* ctx = new tpose_cuda_ctx();
    ctx->device = device;
    ctx->maxT = max_triangles;
    ctx->maxP = max_points;
    auto bail = [&](const char* what, cudaError_t e) {
        g_create_error = std::string("tpose_cuda_create: ") + what + ": " + cudaGetErrorString(e);
        delete ctx;
        return (int)TPOSE_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if (stream) ctx->stream = (cudaStream_t)stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
        ctx->own_stream = true;
    }
    const size_t nv = (size_t)kVariants * max_triangles;
#define ALLOC(ptr, bytes) if ((e = cudaMalloc((void**)&(ptr), (bytes))) != cudaSuccess) return bail("cudaMalloc " #ptr, e)
    const size_t mask_words = ((size_t)max_triangles + 31) / 32;
    ctx->report_pts_off = ((32 + sizeof(unsigned) * mask_words + 15) / 16) * 16;
    ALLOC(ctx->report, ctx->report_pts_off + sizeof(float2) * max_points);
    ctx->run_state = (int*)ctx->report;
    ctx->err_state = (float*)(ctx->report + 16);
    ctx->wmask = (unsigned*)(ctx->report + 32);
    ctx->pts = (float2*)(ctx->report + ctx->report_pts_off);
    ALLOC(ctx->geterr_scratch, kGeterrScratchBytes);
    ALLOC(ctx->wbuf, (sizeof(int4) + 3 * sizeof(int)) * (size_t)max_triangles);
    ctx->wtri = (int4*)ctx->wbuf; ctx->whe = (int*)(ctx->wbuf + sizeof(int4) * (size_t)max_triangles);
    ALLOC(ctx->tri, sizeof(int4) * max_triangles);
    for (int s = 0; s < 2; s++) {
        ALLOC(ctx->cn[s], sizeof(int) * nv);
        ALLOC(ctx->ca[s], sizeof(int4) * nv);
        ALLOC(ctx->ten[s], sizeof(int) * nv);
    }
    ALLOC(ctx->gr, sizeof(int2) * max_points);
    ALLOC(ctx->edges, sizeof(int2) * 3 * (size_t)max_triangles);
    ALLOC(ctx->tri_edges, sizeof(int4) * max_triangles);
    ALLOC(ctx->corner_off, sizeof(int) * ((size_t)max_points + 1));
    ALLOC(ctx->corners, sizeof(int) * 3 * (size_t)max_triangles);
    ALLOC(ctx->rec, sizeof(EdgeRec) * kEdgeMoves * 3 * (size_t)max_triangles);
    ALLOC(ctx->esnap, sizeof(int2) * kSnapPerEdge * 3 * (size_t)max_triangles);
    ALLOC(ctx->ebox, sizeof(unsigned) * 3 * (size_t)max_triangles);
    ALLOC(ctx->gbox, sizeof(unsigned) * ((3 * (size_t)max_triangles + kEdgeGroup - 1) / kEdgeGroup + 1));
    ALLOC(ctx->ticket, sizeof(unsigned) * 2);
    ALLOC(ctx->job_dev, sizeof(EvalJob));
#undef ALLOC
    for (int s = 0; s < 2; s++) {
        cudaMemsetAsync(ctx->cn[s], 0, sizeof(int) * nv, ctx->stream);
        cudaMemsetAsync(ctx->ca[s], 0, sizeof(int4) * nv, ctx->stream);
        cudaMemsetAsync(ctx->ten[s], 0, sizeof(int) * nv, ctx->stream);
    }
    cudaMemsetAsync(ctx->gr, 0, sizeof(int2) * max_points, ctx->stream);
    cudaMemsetAsync(ctx->ticket, 0, sizeof(unsigned) * 2, ctx->stream);
    cudaMemsetAsync(ctx->run_state, 0, sizeof(int) * 4, ctx->stream);
    const float st[4] = {1.0f, 0.0f, 0.0f, 0.0f};   // toterr starts at 1 (triangulation.hpp:648)
    cudaMemcpyAsync(ctx->err_state, st, sizeof st, cudaMemcpyHostToDevice, ctx->stream);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaEventCreate(&ctx->ev_mid);
    cudaEventCreate(&ctx->ev_mid0);
    cudaEventCreateWithFlags(&ctx->ev_batch, cudaEventDisableTiming);
    {
        const size_t cap = std::max((sizeof(int4) + 3 * sizeof(int)) * (size_t)max_triangles, sizeof(float2) * (size_t)max_points);
        bool ok = cap <= (4u << 20);
        for (int i = 0; ok && i < tpose_cuda_ctx::kStageSlots; i++)
            ok = cudaHostAlloc((void**)&ctx->h_stage[i], cap, cudaHostAllocDefault) == cudaSuccess &&
                 cudaEventCreateWithFlags(&ctx->h_stage_ev[i], cudaEventDisableTiming) == cudaSuccess;
        if (ok) ctx->h_stage_cap = cap; else (void)cudaGetLastError();
    }
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail("init", e);
    const char* g = getenv("TPOSE_CUDA_GROUP");
    if (g) ctx->group = atoi(g);
    // TPOSE_CUDA_PATH = table | tiles: which kernels integrate the edges (both produce the same records, bit for bit)
    const char* pth = getenv("TPOSE_CUDA_PATH");
    if (pth && std::string(pth) == "table") ctx->path = kPathTable;
    else if (pth && std::string(pth) == "tiles") ctx->path = kPathTiles;
    if (ctx->path == kPathTiles) {
        cudaError_t ea = cudaFuncSetAttribute(k_tiles<kTileH>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem_bytes<kTileH>());
        if (ea != cudaSuccess) return bail("k_tiles shared memory", ea);
    }
    const char* ng = getenv("TPOSE_CUDA_NO_GRAPH");
    if (ng && atoi(ng)) ctx->use_graphs = false;
    *out = ctx;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_destroy(tpose_cuda_ctx* ctx)
{
    if (!ctx) return TPOSE_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    for (auto& g : ctx->frame_graphs) cudaGraphExecDestroy(g.exec);
    for (auto& s : ctx->img) { cudaFree(s.rgba); cudaFree(s.table); cudaFree(s.coarse); cudaFree(s.tmap); }
    cudaFree(ctx->esnap); cudaFree(ctx->ebox); cudaFree(ctx->gbox);
    for (int s = 0; s < 2; s++) { cudaFree(ctx->cn[s]); cudaFree(ctx->ca[s]); cudaFree(ctx->ten[s]); }
    cudaFree(ctx->staging); cudaFree(ctx->tri);
    cudaFree(ctx->edges); cudaFree(ctx->tri_edges); cudaFree(ctx->corner_off); cudaFree(ctx->corners); cudaFree(ctx->rec);
    cudaFree(ctx->gr); cudaFree(ctx->wbuf); cudaFree(ctx->ticket); cudaFree(ctx->geterr_scratch);
    cudaFree(ctx->report); cudaFree(ctx->ids); cudaFree(ctx->flush); cudaFree(ctx->list_buf);
    cudaFree(ctx->job_dev); cudaFree(ctx->batch_jobs);
    if (ctx->h_report) cudaFreeHost(ctx->h_report);
    cudaFree(ctx->d_hist_pts); cudaFree(ctx->d_hist_ten); cudaFree(ctx->d_hist_err); cudaFree(ctx->d_hist_mask); cudaFree(ctx->d_hist_gr);
    if (ctx->h_hist_pts) cudaFreeHost(ctx->h_hist_pts);
    if (ctx->h_hist_ten) cudaFreeHost(ctx->h_hist_ten);
    if (ctx->h_hist_err) cudaFreeHost(ctx->h_hist_err);
    if (ctx->h_hist_mask) cudaFreeHost(ctx->h_hist_mask);
    for (int i = 0; i < tpose_cuda_ctx::kStageSlots; i++) {
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->h_stage_ev[i]) cudaEventDestroy(ctx->h_stage_ev[i]);
    }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
    if (ctx->ev_mid0) cudaEventDestroy(ctx->ev_mid0);
    if (ctx->ev_batch) cudaEventDestroy(ctx->ev_batch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TPOSE_OK;
}

// ---- images ---------------------------------------------------------------------------------

static int build_table(tpose_cuda_ctx* ctx, ImageSlot& s)
{
    const size_t need_table = (size_t)(s.W + 1) * s.H * sizeof(uint4);
    if (s.table_cap < need_table) { cudaFree(s.table); s.table = nullptr; CK(cudaMalloc((void**)&s.table, need_table)); s.table_cap = need_table; }
    s.pitch = s.W + 1;
    const size_t smem = (size_t)s.ipitch * 4;
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_build_table<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_build_table<true><<<s.H, kTableThreads, smem, ctx->stream>>>(s.rgba, s.ipitch, s.W, s.H, s.table, s.pitch);
    ctx->launches++;
    CK(cudaGetLastError());
    return TPOSE_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else (void)cudaGetLastError();
    }
    return fn;
}

// tile path: coarse column prefixes + the tensor map the tiles are loaded through
static int build_tiles(tpose_cuda_ctx* ctx, ImageSlot& s)
{
    s.ntx = (s.W + kTileW - 1) / kTileW;
    s.nty = (s.H + kTileH - 1) / kTileH;
    const size_t need = (size_t)(s.ntx + 1) * s.H * sizeof(uint4);
    if (s.coarse_cap < need) { cudaFree(s.coarse); s.coarse = nullptr; CK(cudaMalloc((void**)&s.coarse, need)); s.coarse_cap = need; }
    if (!s.tmap) CK(cudaMalloc(&s.tmap, sizeof(CUtensorMap)));
    k_build_coarse<<<s.H, kTableThreads, 0, ctx->stream>>>(s.rgba, s.ipitch, s.W, s.H, s.coarse, s.ntx);
    ctx->launches++;
    CK(cudaGetLastError());
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(ctx, TPOSE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver (TPOSE_CUDA_PATH=table selects the table kernels)");
    alignas(64) CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)s.W, (cuuint64_t)s.H};
    const cuuint64_t gstride[1] = {(cuuint64_t)s.ipitch * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTileW, (cuuint32_t)kTileH};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, s.rgba, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, TPOSE_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    return upload_bytes(ctx, s.tmap, &tm, sizeof tm);
}

static int set_image_common(tpose_cuda_ctx* ctx, int slot, const uint8_t* src, bool src_on_host,
                            int w, int h, int rw, int rh)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (slot < 0 || slot > 1 || !src || w <= 0 || h <= 0 || rw <= 0 || rh <= 0)
        return fail(ctx, TPOSE_ERR_ARG, "set_image: bad arguments");
    if (rw > TPOSE_MAX_WIDTH || rh > TPOSE_MAX_HEIGHT)
        return fail(ctx, TPOSE_ERR_CAPACITY, "set_image: raster exceeds 4095 x 8192");
    CK(cudaSetDevice(ctx->device));
    ImageSlot& s = ctx->img[slot];
    s.ready = false;
    const int ipitch = (rw + 3) & ~3;                 // rows start 16-byte aligned (bulk copies, the tensor map's stride)
    const size_t need_rgba = (size_t)ipitch * rh * sizeof(uchar4);
    if (s.rgba_cap < need_rgba) { cudaFree(s.rgba); s.rgba = nullptr; CK(cudaMalloc((void**)&s.rgba, need_rgba)); s.rgba_cap = need_rgba; }
    s.W = rw; s.H = rh; s.ipitch = ipitch;
    const cudaMemcpyKind kind = src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (ipitch != rw) CK(cudaMemsetAsync(s.rgba, 0, need_rgba, ctx->stream));
    if (rw == w && rh == h) {
        if (ipitch == rw) CK(cudaMemcpyAsync(s.rgba, src, need_rgba, kind, ctx->stream));
        else CK(cudaMemcpy2DAsync(s.rgba, (size_t)ipitch * 4, src, (size_t)w * 4, (size_t)w * 4, (size_t)h, kind, ctx->stream));
    } else {
        const size_t need_src = (size_t)w * h * sizeof(uchar4);
        if (ctx->staging_cap < need_src) { cudaFree(ctx->staging); ctx->staging = nullptr; CK(cudaMalloc((void**)&ctx->staging, need_src)); ctx->staging_cap = need_src; }
        CK(cudaMemcpyAsync(ctx->staging, src, need_src, kind, ctx->stream));
        dim3 grid((rw + 127) / 128, rh);
        k_resample<<<grid, 128, 0, ctx->stream>>>(ctx->staging, w, h, s.rgba, ipitch, rw, rh);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    const int rc = ctx->path == kPathTiles ? build_tiles(ctx, s) : build_table(ctx, s);
    if (rc) return rc;
    s.ready = true;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_image(tpose_cuda_ctx* ctx, int slot, const uint8_t* rgba_host, int w, int h, int rw, int rh)
{
    int rc = set_image_common(ctx, slot, rgba_host, true, w, h, rw, rh);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));   // the caller may reuse (or free) its pixels
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_image_device(tpose_cuda_ctx* ctx, int slot, const uint8_t* rgba_dev, int w, int h, int rw, int rh)
{
    return set_image_common(ctx, slot, rgba_dev, false, w, h, rw, rh);
}

// ---- mesh -----------------------------------------------------------------------------------

extern "C" int tpose_cuda_upload(tpose_cuda_ctx* ctx, const float* pts_xy, int np, const int32_t* tri_xyzw, int nt,
                                 const int32_t* colors_rgba, float ratio)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!pts_xy || !tri_xyzw || np < 3 || nt < 1) return fail(ctx, TPOSE_ERR_ARG, "upload: bad arguments");
    if (nt > ctx->maxT || np > ctx->maxP) return fail(ctx, TPOSE_ERR_CAPACITY, "upload: mesh larger than the context capacity");
    if (!(ratio > 0.0f)) return fail(ctx, TPOSE_ERR_ARG, "upload: RATIO must be positive");
    for (int t = 0; t < nt; t++)
        for (int j = 0; j < 3; j++) {
            const int v = tri_xyzw[4 * t + j];
            if (v < 0 || v >= np) return fail(ctx, TPOSE_ERR_ARG, "upload: triangle references a vertex outside [0,NP)");
        }
    CK(cudaSetDevice(ctx->device));
    int rc = upload_bytes(ctx, ctx->pts, pts_xy, sizeof(float2) * np);
    if (rc) return rc;
    if ((rc = upload_bytes(ctx, ctx->tri, tri_xyzw, sizeof(int4) * nt)) != 0) return rc;
    if (nt != ctx->NT || np != ctx->edges_np || ctx->tri_host.size() != (size_t)4 * nt ||
        memcmp(ctx->tri_host.data(), tri_xyzw, sizeof(int32_t) * 4 * nt) != 0) {
        ctx->tri_host.assign(tri_xyzw, tri_xyzw + (size_t)4 * nt);
        ctx->edges_dirty = true;
    }
    if (colors_rgba) {   // triangulation.hpp:633-641: every variant row gets the triangle's colour
        if ((rc = upload_bytes(ctx, ctx->ca[0], colors_rgba, sizeof(int4) * nt)) != 0) return rc;
        for (int s = 0; s < 2; s++)
            for (int k = (s == 0 ? 1 : 0); k < kVariants; k++)
                CK(cudaMemcpyAsync(ctx->ca[s] + (size_t)k * nt, ctx->ca[0], sizeof(int4) * nt, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (nt != ctx->NT) ctx->watching = false;   // a watch topology is only valid for the same NT
    ctx->NT = nt; ctx->NP = np; ctx->ratio = ratio; ctx->have_mesh = true;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_upload_points(tpose_cuda_ctx* ctx, const float* pts_xy, int np)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "upload_points: no mesh uploaded");
    if (!pts_xy || np != ctx->NP) return fail(ctx, TPOSE_ERR_ARG, "upload_points: NP differs from the uploaded mesh");
    CK(cudaSetDevice(ctx->device));
    return upload_bytes(ctx, ctx->pts, pts_xy, sizeof(float2) * np);
}

// ---- evaluation -----------------------------------------------------------------------------

// Unique-edge view of the uploaded triangles: side j of triangle t runs from corner j to corner
// j+1; the two sides of an interior edge share one entry of `edges` (stored first < second end
// point), so k_edges integrates it once.  Also the triangle corners that sit on every vertex
// (k_corners evaluates the moved variants grouped by the vertex they move).
static int rebuild_edges(tpose_cuda_ctx* ctx)
{
    if (!ctx->edges_dirty) return TPOSE_OK;
    const int nt = ctx->NT, np = ctx->NP;
    if (ctx->tri_host.size() < (size_t)4 * nt) return fail(ctx, TPOSE_ERR_STATE, "index buffer mirror is incomplete");
    tpose::edgeview::view& v = ctx->edge_view;
    if (!tpose::edgeview::build(ctx->tri_host.data(), nt, np, v, ctx->scratch_keys))
        return fail(ctx, TPOSE_ERR_ARG, "index buffer references a vertex outside [0,NP)");
    // through the page-locked staging slots (the view's vectors are reused): no synchronisation here
    int rc;
    if ((rc = upload_bytes(ctx, ctx->edges, v.edges.data(), sizeof(int2) * v.ne)) != 0) return rc;
    if ((rc = upload_bytes(ctx, ctx->tri_edges, v.tri_edges.data(), sizeof(int4) * nt)) != 0) return rc;
    if ((rc = upload_bytes(ctx, ctx->corner_off, v.corner_off.data(), sizeof(int) * (np + 1))) != 0) return rc;
    if ((rc = upload_bytes(ctx, ctx->corners, v.corners.data(), sizeof(int) * 3 * nt)) != 0) return rc;
    ctx->NE = v.ne; ctx->edges_np = np;
    ctx->edges_dirty = false;
    return TPOSE_OK;
}

static bool valid_group(int g) { return g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32; }

// Lanes per (edge, move) item.  More lanes shorten the serial row walk of an item (what a single
// image needs: it cannot fill the machine anyway) but repeat its set-up (two snaps, two integer
// divisions) per lane (what a batch pays for).  So: as many lanes as it takes to put ~250 k lanes
// on the device, but at least 4 rows per lane.  The mean vertical extent of an edge is about
// 0.6*H*sqrt(2/NT) rows.
static int pow2_floor(double x)
{
    int g = 1;
    while (g < 32 && 2.0 * g <= x) g <<= 1;
    return g;
}

static int group_for(double mean_rows, double total_items)
{
    const int fill = pow2_floor(250000.0 / (total_items > 1.0 ? total_items : 1.0));
    const int rows = pow2_floor(mean_rows / 4.0);
    return fill < rows ? fill : rows;
}

static double mean_edge_rows(int NT, int H) { return 0.6 * (double)H * sqrt(2.0 / (double)NT); }

static int pick_group(const tpose_cuda_ctx* ctx, int NT, int H)
{
    if (valid_group(ctx->group)) return ctx->group;
    return group_for(mean_edge_rows(NT, H), (double)ctx->NE * kEdgeMoves);
}

// 32-bit colour sums are exact while 255*W*H < 2^31
static bool narrow_sums(int W, int H) { return 255ll * W * H < (1ll << 31); }

constexpr int kSMs = 148;
constexpr int kMaxGridBlocks = kSMs * 4;

static int quantised_grid(long long blocks_needed, int cap = kMaxGridBlocks)
{
    int g = 32;
    while (g < blocks_needed && g < cap) g <<= 1;
    return g < cap ? g : cap;
}

static EvalShape make_shape(int group, int max_np, int max_ne, int max_nt, bool narrow, int path, int max_tiles, bool single_job)
{
    const int per_block = kEdgeThreads / group;
    // k_edges' shape (eval.cuh): the dense one where it spares a single image a second pass through the chain
    const long long edge_blocks = ((long long)max_ne * kEdgeMoves + per_block - 1) / per_block;
    const bool dense = single_job && path == kPathTable && edge_blocks > kSMs * kEdgeBlocksWide && edge_blocks <= kSMs * kEdgeBlocksDense;
    // k_corners: blocks for the base variants in proportion to NT (one variant per thread), and vertex blocks up to
    // what is resident together with them, so that neither part of the grid takes a second turn on a busy SM
    const int base_blocks = quantised_grid((max_nt + kCornerThreads - 1) / kCornerThreads, 128) < 16 ? 16
                            : quantised_grid((max_nt + kCornerThreads - 1) / kCornerThreads, 128);
    EvalShape sh{group, quantised_grid(edge_blocks, kSMs * (dense ? kEdgeBlocksDense : kEdgeBlocksWide)),
                 quantised_grid((max_np + kCornerWarps - 1) / kCornerWarps, kSMs * kCornerBlocksPerSM - base_blocks), narrow, path, 0, 0, base_blocks, dense};
    if (path == kPathTiles) {
        // (group, grid_e and narrow do not exist on the tile path: keep them out of the graph key)
        sh.group = 1; sh.grid_e = 0; sh.narrow = true;
        sh.grid_s = quantised_grid((max_ne + kSetupThreads - 1) / kSetupThreads);
        sh.grid_tiles = max_tiles;
    }
    return sh;
}

template <int G>
static void launch_edges(const EvalJob* jobs, int njobs, const EvalShape& sh, int set, unsigned flags, double tol, int frame, cudaStream_t st, const Early& stop)
{
    const dim3 grid(sh.grid_e + 1, njobs);   // + the service block
    if (sh.dense) {
        if (sh.narrow) k_edges<G, unsigned, true><<<grid, kEdgeThreads, 0, st>>>(jobs, set, flags, tol, frame, stop);
        else k_edges<G, unsigned long long, true><<<grid, kEdgeThreads, 0, st>>>(jobs, set, flags, tol, frame, stop);
    } else {
        if (sh.narrow) k_edges<G, unsigned, false><<<grid, kEdgeThreads, 0, st>>>(jobs, set, flags, tol, frame, stop);
        else k_edges<G, unsigned long long, false><<<grid, kEdgeThreads, 0, st>>>(jobs, set, flags, tol, frame, stop);
    }
}

// one evaluation of every job: k_edges (or k_setup + k_tiles) then k_corners
// `stop`: what a single-job launch is told as kernel arguments (eval.cuh, Early), all null for a batch
static void launch_eval(const EvalShape& sh, const EvalJob* jobs, int njobs, int set, unsigned flags, double tol, int frame, cudaStream_t st,
                        cudaEvent_t mid = nullptr, cudaEvent_t mid0 = nullptr, const Early& stop = Early{})
{
    if (sh.path == kPathTiles) {
        k_setup<kTileH><<<dim3(sh.grid_s + 1, njobs), kSetupThreads, 0, st>>>(jobs, set, flags, tol, frame);   // + the service block
        if (mid0) cudaEventRecord(mid0, st);
        k_tiles<kTileH><<<dim3(sh.grid_tiles, njobs), kTileThreads, tile_smem_bytes<kTileH>(), st>>>(jobs, flags);
        if (mid) cudaEventRecord(mid, st);
        k_corners<EdgeRec><<<dim3(sh.grid_t + sh.grid_b, njobs), kCornerThreads, 0, st>>>(jobs, set, flags, frame, sh.grid_b, stop);   // (k_tiles adds 64-bit words)
        return;
    }
    switch (sh.group) {
    case 1: launch_edges<1>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    case 2: launch_edges<2>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    case 4: launch_edges<4>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    case 8: launch_edges<8>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    case 16: launch_edges<16>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    default: launch_edges<32>(jobs, njobs, sh, set, flags, tol, frame, st, stop); break;
    }
    if (mid) cudaEventRecord(mid, st);
    if (sh.narrow) k_corners<EdgeRecNarrow><<<dim3(sh.grid_t + sh.grid_b, njobs), kCornerThreads, 0, st>>>(jobs, set, flags, frame, sh.grid_b, stop);
    else k_corners<EdgeRec><<<dim3(sh.grid_t + sh.grid_b, njobs), kCornerThreads, 0, st>>>(jobs, set, flags, frame, sh.grid_b, stop);
}
static int launches_per_eval(const EvalShape& sh) { return sh.path == kPathTiles ? 3 : 2; }

// ---- sequences of evaluations, replayed as CUDA graphs ------------------------------------------
// A sequence is `len` evaluations into alternating buffer sets starting at `first_set`.
//   kind 0 (step): every evaluation carries `flags`
//   kind 1 (run):  evaluation 0 carries `flags`, the others add EV_CHECK_PREV, the last EV_FINAL
// The host cannot issue two launches per step as fast as the device retires them, so sequences of
// kGraphMinLen or more evaluations are captured once and replayed; the job records are read from
// device memory at run time, so a captured sequence does not depend on the mesh or the image.
constexpr int kGraphMinLen = 1;
constexpr size_t kGraphCacheSize = 64;

static void issue_direct(const SeqKey& k, cudaStream_t st, cudaEvent_t mid = nullptr, cudaEvent_t mid0 = nullptr)
{
    for (int j = 0; j < k.len; j++) {
        unsigned f = k.flags;
        if (k.kind == 1) {
            if (j > 0) f |= EV_CHECK_PREV;
            if (j == k.len - 1) f |= EV_FINAL;
        }
        launch_eval(k.shape, k.jobs, k.njobs, (k.first_set + j) & 1, f, k.tol, k.kind == 1 ? j : -1, st, mid, mid0, k.early);
    }
}

static bool same_early(const Early& a, const Early& b)
{
    return a.stop == b.stop && a.edges == b.edges && a.edges_cap == b.edges_cap && a.corner_off == b.corner_off && a.pts == b.pts && a.np_cap == b.np_cap;
}

// the kernel arguments of a single-job launch of this context (eval.cuh, Early); `in_run`: with the run's stop flag
static Early early_of(const tpose_cuda_ctx* ctx, bool in_run)
{
    return Early{in_run ? ctx->run_state : nullptr, ctx->edges, 3 * ctx->maxT, ctx->corner_off, ctx->pts, ctx->maxP};
}

static bool same_key(const SeqKey& a, const SeqKey& b)
{
    return a.jobs == b.jobs && a.njobs == b.njobs && a.kind == b.kind && a.len == b.len && a.first_set == b.first_set && same_early(a.early, b.early) &&
           a.flags == b.flags && a.tol == b.tol && a.shape.group == b.shape.group && a.shape.grid_e == b.shape.grid_e &&
           a.shape.grid_t == b.shape.grid_t && a.shape.narrow == b.shape.narrow && a.shape.path == b.shape.path &&
           a.shape.grid_s == b.shape.grid_s && a.shape.grid_tiles == b.shape.grid_tiles && a.shape.grid_b == b.shape.grid_b && a.shape.dense == b.shape.dense;
}

static int issue_sequence(tpose_cuda_ctx* ctx, const SeqKey& key)
{
    ctx->run_cache_valid = false; ctx->run_mask_valid = false;
    ctx->launches += (int64_t)launches_per_eval(key.shape) * key.len;
    if (key.len < kGraphMinLen || !ctx->use_graphs || (key.len == 1 && ctx->time_kernels)) {
        const bool mid = key.len == 1 && ctx->time_kernels;
        issue_direct(key, ctx->stream, mid ? ctx->ev_mid : nullptr, mid ? ctx->ev_mid0 : nullptr);
        ctx->mid_valid = mid;
        CK(cudaGetLastError());
        return TPOSE_OK;
    }
    ctx->mid_valid = false;
    ctx->graph_clock++;
    for (auto& g : ctx->graphs)
        if (same_key(g.key, key)) {
            g.stamp = ctx->graph_clock;
            CK(cudaGraphLaunch(g.exec, ctx->stream));
            return TPOSE_OK;
        }
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        issue_direct(key, ctx->stream);
        e = cudaStreamEndCapture(ctx->stream, &graph);
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {          // no graphs on this stream (e.g. it is being captured by the caller): launch directly
        (void)cudaGetLastError();
        ctx->use_graphs = false;
        issue_direct(key, ctx->stream);
        CK(cudaGetLastError());
        return TPOSE_OK;
    }
    if (ctx->graphs.size() >= kGraphCacheSize) {
        size_t oldest = 0;
        for (size_t i = 1; i < ctx->graphs.size(); i++) if (ctx->graphs[i].stamp < ctx->graphs[oldest].stamp) oldest = i;
        cudaGraphExecDestroy(ctx->graphs[oldest].exec);
        ctx->graphs.erase(ctx->graphs.begin() + (long)oldest);
    }
    ctx->graphs.push_back(SeqGraph{key, exec, ctx->graph_clock});
    CK(cudaGraphLaunch(exec, ctx->stream));
    return TPOSE_OK;
}

// `nsteps` plain steps: even-length graph chunks (the set parity is the same after each), then singles
static int issue_steps(tpose_cuda_ctx* ctx, const EvalJob* jobs, int njobs, const EvalShape& sh, unsigned flags, int first_set, int nsteps)
{
    int left = nsteps, set = first_set;
    for (int chunk : {64, 16, 4})
        while (left >= chunk) {
            int rc = issue_sequence(ctx, SeqKey{jobs, njobs, sh, 0, chunk, set, flags, 0.0, njobs == 1 ? early_of(ctx, false) : Early{}});
            if (rc) return rc;
            left -= chunk;
        }
    if (left > 0) {
        int rc = issue_sequence(ctx, SeqKey{jobs, njobs, sh, 0, left, set, flags, 0.0, njobs == 1 ? early_of(ctx, false) : Early{}});
        if (rc) return rc;
    }
    return TPOSE_OK;
}

// the job record of this context for (slot, flavour)
static EvalJob make_job(const tpose_cuda_ctx* ctx, int slot, int flavour)
{
    const ImageSlot& s = ctx->img[slot];
    EvalJob j;
    memset(&j, 0, sizeof j);
    j.table = s.table; j.pitch = s.pitch; j.W = s.W; j.H = s.H;
    j.tmap = s.tmap; j.coarse = s.coarse; j.ntx = s.ntx; j.nty = s.nty;
    j.esnap = ctx->esnap; j.ebox = ctx->ebox; j.gbox = ctx->gbox;
    j.pts = ctx->pts; j.tri = ctx->tri; j.NT = ctx->NT; j.NP = ctx->NP; j.NE = ctx->NE;
    j.edges = ctx->edges; j.tri_edges = ctx->tri_edges; j.corner_off = ctx->corner_off; j.corners = ctx->corners;
    j.rec = ctx->rec;
    j.ratio = ctx->ratio;
    j.dp = dp_for(ctx->NT, flavour);
    j.step = (flavour == TPOSE_WARP ? 0.00003f : 0.00005f) * ctx->step_scale;
    for (int k = 0; k < 2; k++) { j.cn[k] = ctx->cn[k]; j.ca[k] = ctx->ca[k]; j.ten[k] = ctx->ten[k]; }
    j.gr = ctx->gr;
    j.wtri = ctx->watching ? ctx->wtri : nullptr;
    j.whe = ctx->watching ? ctx->whe : nullptr;
    j.wmask = ctx->wmask;
    j.hist_pts = ctx->d_hist_pts; j.hist_ten = ctx->d_hist_ten; j.hist_err = ctx->d_hist_err; j.hist_mask = ctx->d_hist_mask;
    j.hist_gr = ctx->d_hist_gr;
    j.err_state = ctx->err_state; j.run_state = ctx->run_state; j.geterr_scratch = ctx->geterr_scratch;
    j.trig = ctx->ticket + 1;
    return j;
}

// make job_dev current (a 200-byte copy, only when something changed)
static int sync_job(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    int rc = rebuild_edges(ctx);
    if (rc) return rc;
    const EvalJob j = make_job(ctx, slot, flavour);
    if (!ctx->job_valid || memcmp(&j, &ctx->job_host, sizeof j) != 0) {
        ctx->job_host = j;
        CK(cudaMemcpyAsync(ctx->job_dev, &ctx->job_host, sizeof j, cudaMemcpyHostToDevice, ctx->stream));
        ctx->job_valid = true;
    }
    return TPOSE_OK;
}

static EvalShape shape_of(const tpose_cuda_ctx* ctx, int slot)
{
    const ImageSlot& im = ctx->img[slot];
    return make_shape(pick_group(ctx, ctx->NT, im.H), ctx->NP, ctx->NE, ctx->NT, narrow_sums(im.W, im.H), ctx->path, im.ntx * im.nty, true);
}

static int eval(tpose_cuda_ctx* ctx, int slot, int flavour, int set, unsigned flags, double tol = 0.0)
{
    int rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    rc = issue_sequence(ctx, SeqKey{ctx->job_dev, 1, shape_of(ctx, slot), 0, 1, set, flags, tol, early_of(ctx, false)});
    ctx->mid_valid = false;      // only tpose_cuda_step brackets its kernels with events
    return rc;
}

static int check_ready(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (slot < 0 || slot > 1 || (flavour != TPOSE_TRIANGULATE && flavour != TPOSE_WARP)) return fail(ctx, TPOSE_ERR_ARG, "bad slot or flavour");
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "no mesh uploaded");
    if (!ctx->img[slot].ready) return fail(ctx, TPOSE_ERR_STATE, "no image in this slot");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, TPOSE_ERR_CUDA, cudaGetErrorString(e));
    return TPOSE_OK;
}

static unsigned fused_flags(int flavour)
{
    return flavour == TPOSE_WARP
        ? (EV_WRITE_CN | EV_WRITE_TEN | EV_MEAN_STORED | EV_GRAD_SHIFT)
        : (EV_WRITE_CN | EV_WRITE_CA | EV_WRITE_TEN | EV_MEAN_DIVIDE | EV_GRAD_SHIFT);
}

extern "C" int tpose_cuda_colors(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    return eval(ctx, slot, flavour, ctx->cur, flavour == TPOSE_WARP ? EV_WRITE_CN : (EV_WRITE_CN | EV_WRITE_CA));
}

extern "C" int tpose_cuda_energy(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if ((rc = rebuild_edges(ctx)) != 0) return rc;                       // corner_off: which vertices are referenced
    k_zero_gradient<<<(ctx->NP + 255) / 256, 256, 0, ctx->stream>>>(ctx->gr, ctx->NP, ctx->corner_off);   // triangle.vs:95-97
    ctx->launches++;
    const unsigned fl = EV_WRITE_TEN | EV_MEAN_STORED | (flavour == TPOSE_WARP ? 0u : EV_MEAN_DIVIDE);
    return eval(ctx, slot, flavour, ctx->cur, fl);
}

extern "C" int tpose_cuda_shift(tpose_cuda_ctx* ctx, int flavour)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    ctx->run_cache_valid = false;
    k_gradient<<<(ctx->NT + 127) / 128, 128, 0, ctx->stream>>>(ctx->tri, ctx->NT, ctx->ten[ctx->cur], ctx->gr);
    k_shift<<<(ctx->NP + 127) / 128, 128, 0, ctx->stream>>>(ctx->pts, ctx->NP, ctx->gr, ctx->ratio,
                                                            (flavour == TPOSE_WARP ? 0.00003f : 0.00005f) * ctx->step_scale);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return TPOSE_OK;
}

extern "C" int tpose_cuda_step(tpose_cuda_ctx* ctx, int slot, int flavour, int nsteps)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if (nsteps < 0) return fail(ctx, TPOSE_ERR_ARG, "step: nsteps < 0");
    rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = issue_steps(ctx, ctx->job_dev, 1, shape_of(ctx, slot), fused_flags(flavour), ctx->cur ^ 1, nsteps);
    if (rc) return rc;
    ctx->cur = (ctx->cur + nsteps) & 1;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TPOSE_OK;
}

// Launch j evaluates the points after j shifts into set (first+j)&1 and carries the checks of frame
// j-1; launch max_steps is check-only.  See eval.cuh.
// ---- the frame's host traffic without DMA operations --------------------------------------------------
// Page-locked host memory is visible to the device (unified addressing), so when the caller's frame buffers
// are pinned the frame becomes ONE graph: a kernel that pulls the points in over PCIe, the step, and a kernel
// that pushes ten | cn | ca | points out with coalesced 16-byte stores -- no copy-engine operation (each
// costs microseconds of fixed latency, and a frame has four), one stream synchronisation.
__global__ void k_frame_in(const float2* __restrict__ host_pts, float2* __restrict__ pts, int np)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) pts[i] = host_pts[i];
}

struct FrameOut { const int* ten; const int* cn; const int4* ca; const float2* pts; int* h_ten; int* h_cn; int4* h_ca; float2* h_pts; int nv, np; };

__device__ __forceinline__ void copy_words(const int* __restrict__ src, int* __restrict__ dst, int n, int tid, int nthreads)
{
    // 16-byte chunks where both sides allow it (cudaMalloc / cudaHostAlloc memory is at least 256-byte aligned)
    const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    const int n4 = vec ? n >> 2 : 0;
    for (int i = tid; i < n4; i += nthreads) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(src)[i];
    for (int i = 4 * n4 + tid; i < n; i += nthreads) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_frame_out(FrameOut f)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    if (f.h_ten) copy_words(f.ten, f.h_ten, f.nv, tid, nthreads);
    if (f.h_cn) copy_words(f.cn, f.h_cn, f.nv, tid, nthreads);
    if (f.h_ca) copy_words(reinterpret_cast<const int*>(f.ca), reinterpret_cast<int*>(f.h_ca), 4 * f.nv, tid, nthreads);
    if (f.h_pts) copy_words(reinterpret_cast<const int*>(f.pts), reinterpret_cast<int*>(f.h_pts), 2 * f.np, tid, nthreads);
}

// What a run's caller reads -- the verdict block (run_state | err_state | watch mask | points) and, with the history
// on, the frames the run actually took (RS_STEPS of them, not max_steps) -- pushed into the context's page-locked
// host buffers by one kernel behind the run instead of one to five copy-engine operations.
struct RunOut {
    const int* report; int* h_report; int report_words;
    const int* run_state;
    const int* pts; const int* ten; const int* err; const int* mask;      // d_hist_* (null: no history)
    int* h_pts; int* h_ten; int* h_err; int* h_mask;
    int np, nt, words;
};
__global__ void __launch_bounds__(256) k_run_out(RunOut r)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    copy_words(r.report, r.h_report, r.report_words, tid, nthreads);
    if (!r.pts) return;
    const int done = r.run_state[RS_STEPS];
    copy_words(r.pts, r.h_pts, 2 * r.np * done, tid, nthreads);
    copy_words(r.ten, r.h_ten, r.nt * done, tid, nthreads);
    copy_words(r.err, r.h_err, done, tid, nthreads);
    copy_words(r.mask, r.h_mask, r.words * done, tid, nthreads);
}

// the device-side alias of a host pointer if the memory is page-locked and mapped, else nullptr
static void* device_alias(const void* host)
{
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}


static int frame_mapped(tpose_cuda_ctx* ctx, int slot, int flavour, const float* pts_in, int32_t* ten13, int32_t* cn13, int32_t* ca13, float* pts_out,
                        bool* done)
{
    *done = false;
    if (!ctx->use_graphs) return TPOSE_OK;
    // every buffer the caller passed must be device-visible; remember the verdict per pointer set
    const void* host5[5] = {pts_in, ten13, cn13, ca13, pts_out};
    if (memcmp(host5, ctx->frame_host, sizeof host5) != 0) {          // (the attribute queries are not free: once per pointer set)
        for (int i = 0; i < 5; i++) ctx->frame_alias[i] = device_alias(host5[i]);
        memcpy(ctx->frame_host, host5, sizeof host5);
    }
    void* d_in = ctx->frame_alias[0]; void* d_ten = ctx->frame_alias[1]; void* d_cn = ctx->frame_alias[2];
    void* d_ca = ctx->frame_alias[3]; void* d_out = ctx->frame_alias[4];
    if ((pts_in && !d_in) || (ten13 && !d_ten) || (cn13 && !d_cn) || (ca13 && !d_ca) || (pts_out && !d_out)) return TPOSE_OK;
    int rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    const int set = ctx->cur ^ 1;
    const EvalShape sh = shape_of(ctx, slot);
    FrameGraph* g = nullptr;
    for (auto& c : ctx->frame_graphs)
        if (c.in == d_in && c.ten == d_ten && c.cn == d_cn && c.ca == d_ca && c.out == d_out && c.set == set && c.nt == ctx->NT && c.np == ctx->NP &&
            c.slot == slot && c.flavour == flavour && c.job == ctx->job_dev && c.shape.group == sh.group && c.shape.grid_e == sh.grid_e &&
            c.shape.grid_t == sh.grid_t && c.shape.narrow == sh.narrow && c.shape.path == sh.path && c.shape.grid_s == sh.grid_s &&
            c.shape.grid_tiles == sh.grid_tiles && c.shape.grid_b == sh.grid_b) g = &c;
    if (!g) {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            if (d_in) k_frame_in<<<(ctx->NP + 255) / 256, 256, 0, ctx->stream>>>((const float2*)d_in, ctx->pts, ctx->NP);
            launch_eval(sh, ctx->job_dev, 1, set, fused_flags(flavour), 0.0, -1, ctx->stream, nullptr, nullptr, early_of(ctx, false));
            FrameOut f{ctx->ten[set], ctx->cn[set], ctx->ca[set], ctx->pts, (int*)d_ten, (int*)d_cn, (int4*)d_ca, (float2*)d_out, kVariants * ctx->NT, ctx->NP};
            k_frame_out<<<64, 256, 0, ctx->stream>>>(f);
            e = cudaStreamEndCapture(ctx->stream, &graph);
        }
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return TPOSE_OK; }       // (stream under foreign capture etc.: take the copy path)
        if (ctx->frame_graphs.size() >= 8) { cudaGraphExecDestroy(ctx->frame_graphs.front().exec); ctx->frame_graphs.erase(ctx->frame_graphs.begin()); }
        ctx->frame_graphs.push_back(FrameGraph{d_in, d_ten, d_cn, d_ca, d_out, set, ctx->NT, ctx->NP, slot, flavour, sh, ctx->job_dev, exec});
        g = &ctx->frame_graphs.back();
    }
    ctx->run_cache_valid = false; ctx->run_mask_valid = false; ctx->mid_valid = false;
    CK(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += launches_per_eval(sh) + (d_in ? 2 : 1);
    ctx->cur = set;
    CK(cudaStreamSynchronize(ctx->stream));
    *done = true;
    return TPOSE_OK;
}

// One frame of the reference's loop as a single call (triangulate/main.cpp:196-204, warp/main.cpp:220-229):
// optional pointbuf->fill, doenergy + doshift (+ the colours of the pipeline), and the read-backs the
// frame ends with -- one stream synchronisation in all.
extern "C" int tpose_cuda_frame(tpose_cuda_ctx* ctx, int slot, int flavour, const float* pts_in_xy,
                                int32_t* ten13, int32_t* cn13, int32_t* ca13, float* pts_out_xy)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    bool done = false;
    if ((rc = frame_mapped(ctx, slot, flavour, pts_in_xy, ten13, cn13, ca13, pts_out_xy, &done)) != 0 || done) return rc;
    if (pts_in_xy && (rc = upload_bytes(ctx, ctx->pts, pts_in_xy, sizeof(float2) * ctx->NP)) != 0) return rc;
    if ((rc = sync_job(ctx, slot, flavour)) != 0) return rc;
    ctx->cur ^= 1;
    rc = issue_sequence(ctx, SeqKey{ctx->job_dev, 1, shape_of(ctx, slot), 0, 1, ctx->cur, fused_flags(flavour), 0.0, early_of(ctx, false)});
    if (rc) return rc;
    ctx->mid_valid = false;
    const size_t nv = (size_t)kVariants * ctx->NT;
    const int c = ctx->cur;
    if (ten13) CK(cudaMemcpyAsync(ten13, ctx->ten[c], sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn13) CK(cudaMemcpyAsync(cn13, ctx->cn[c], sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca13) CK(cudaMemcpyAsync(ca13, ctx->ca[c], sizeof(int4) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (pts_out_xy) CK(cudaMemcpyAsync(pts_out_xy, ctx->pts, sizeof(float2) * ctx->NP, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

// Evaluation j works on the points after j shifts, writes set (first+j)&1 and carries the checks of
// frame j-1; evaluation max_steps is check-only.  See eval.cuh.
static int issue_run(tpose_cuda_ctx* lead, const EvalJob* jobs, int njobs, const EvalShape& sh, int first_set,
                     int flavour, int max_steps, double rel_tol, unsigned extra_flags = 0u)
{
    return issue_sequence(lead, SeqKey{jobs, njobs, sh, 1, max_steps + 1, first_set, fused_flags(flavour) | EV_HONOUR_STOP | extra_flags, rel_tol,
                                       njobs == 1 ? early_of(lead, true) : Early{}});
}

extern "C" int tpose_cuda_run(tpose_cuda_ctx* ctx, int slot, int flavour, int max_steps, double rel_tol,
                              int* steps_done, uint32_t* trigger_flags, float* last_err)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if (max_steps < 1) return fail(ctx, TPOSE_ERR_ARG, "run: max_steps < 1");
    const bool hist = ctx->hist_frames > 0;
    if (hist && max_steps > ctx->hist_frames) return fail(ctx, TPOSE_ERR_CAPACITY, "run: max_steps exceeds the history capacity (tpose_cuda_set_history)");
    rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->run_state, 0, sizeof(int) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int first = ctx->cur ^ 1;
    rc = issue_run(ctx, ctx->job_dev, 1, shape_of(ctx, slot), first, flavour, max_steps, rel_tol, hist ? (unsigned)EV_HISTORY : 0u);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->hist_np_frames = 0;
    int rs[4]; float es[4];
    if (!ctx->h_report && cudaHostAlloc((void**)&ctx->h_report, ctx->report_pts_off + sizeof(float2) * (size_t)ctx->maxP, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError(); ctx->h_report = nullptr;
    }
    if (hist) { ctx->hist_np = ctx->NP; ctx->hist_nt = ctx->NT; }
    // the device-visible aliases of the context's own page-locked buffers (looked up once)
    if (!ctx->run_alias_known) {
        ctx->run_alias_known = true;
        ctx->d_h_report = (int*)device_alias(ctx->h_report);
    }
    if (hist && !ctx->hist_alias_known) {
        ctx->hist_alias_known = true;
        ctx->d_h_hist[0] = (int*)device_alias(ctx->h_hist_pts); ctx->d_h_hist[1] = (int*)device_alias(ctx->h_hist_ten);
        ctx->d_h_hist[2] = (int*)device_alias(ctx->h_hist_err); ctx->d_h_hist[3] = (int*)device_alias(ctx->h_hist_mask);
    }
    const bool pushed = ctx->d_h_report && (!hist || (ctx->d_h_hist[0] && ctx->d_h_hist[1] && ctx->d_h_hist[2] && ctx->d_h_hist[3]));
    const size_t report_bytes = ctx->report_pts_off + sizeof(float2) * (size_t)ctx->NP;      // (a multiple of 8)
    if (pushed) {
        RunOut r{(const int*)ctx->report, ctx->d_h_report, (int)(report_bytes / 4), ctx->run_state,
                 hist ? (const int*)ctx->d_hist_pts : nullptr, (const int*)ctx->d_hist_ten, (const int*)ctx->d_hist_err, (const int*)ctx->d_hist_mask,
                 ctx->d_h_hist[0], ctx->d_h_hist[1], ctx->d_h_hist[2], ctx->d_h_hist[3], ctx->NP, ctx->NT, (ctx->NT + 31) / 32};
        k_run_out<<<hist ? 32 : 4, 256, 0, ctx->stream>>>(r);
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        if (hist) {     // every frame the run may have taken, in the same synchronisation as the verdict
            const size_t f = (size_t)max_steps, words = ((size_t)ctx->NT + 31) / 32;
            CK(cudaMemcpyAsync(ctx->h_hist_pts, ctx->d_hist_pts, sizeof(float2) * ctx->NP * f, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_hist_ten, ctx->d_hist_ten, sizeof(int) * ctx->NT * f, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_hist_err, ctx->d_hist_err, sizeof(float) * f, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(ctx->h_hist_mask, ctx->d_hist_mask, sizeof(unsigned) * words * f, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (ctx->h_report) {
            // verdict + watch mask + points in one copy (the mask region is sized for maxT: a few KB)
            CK(cudaMemcpyAsync(ctx->h_report, ctx->report, report_bytes, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        } else {
            CK(cudaMemcpyAsync(rs, ctx->run_state, sizeof rs, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(es, ctx->err_state, sizeof es, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (ctx->h_report) {
        memcpy(rs, ctx->h_report, sizeof rs); memcpy(es, ctx->h_report + 16, sizeof es);
        ctx->h_run_mask = (unsigned*)(ctx->h_report + 32);
        ctx->h_run_pts = (float*)(ctx->h_report + ctx->report_pts_off);
    }
    ctx->run_cache_valid = ctx->h_report != nullptr; ctx->run_toterr = es[0];
    ctx->run_mask_valid = !hist && ctx->h_report && ctx->watching && (rs[RS_TRIG] & kTrigTopology);
    if (hist) ctx->hist_np_frames = rs[RS_STEPS];
    if (rs[RS_STEPS] > 0) ctx->cur = (first + rs[RS_STEPS] - 1) & 1;
    if (steps_done) *steps_done = rs[RS_STEPS];
    if (trigger_flags) *trigger_flags = (uint32_t)rs[RS_TRIG];
    if (last_err) *last_err = es[1];
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_history(tpose_cuda_ctx* ctx, int frames)
{
    if (!ctx || frames < 0) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (frames > ctx->hist_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_hist_pts); cudaFree(ctx->d_hist_ten); cudaFree(ctx->d_hist_err); cudaFree(ctx->d_hist_mask); cudaFree(ctx->d_hist_gr);
        ctx->d_hist_gr = nullptr;
        if (ctx->h_hist_pts) cudaFreeHost(ctx->h_hist_pts);
        if (ctx->h_hist_ten) cudaFreeHost(ctx->h_hist_ten);
        if (ctx->h_hist_err) cudaFreeHost(ctx->h_hist_err);
        if (ctx->h_hist_mask) cudaFreeHost(ctx->h_hist_mask);
        ctx->d_hist_pts = nullptr; ctx->d_hist_ten = nullptr; ctx->d_hist_err = nullptr; ctx->d_hist_mask = nullptr;
        ctx->h_hist_pts = nullptr; ctx->h_hist_ten = nullptr; ctx->h_hist_err = nullptr; ctx->h_hist_mask = nullptr;
        ctx->hist_alias_known = false;
        ctx->hist_cap = 0;
        const size_t f = (size_t)frames, words = ((size_t)ctx->maxT + 31) / 32;
        CK(cudaMalloc((void**)&ctx->d_hist_pts, sizeof(float2) * ctx->maxP * f));
        CK(cudaMalloc((void**)&ctx->d_hist_ten, sizeof(int) * ctx->maxT * f));
        CK(cudaMalloc((void**)&ctx->d_hist_err, sizeof(float) * f));
        CK(cudaMalloc((void**)&ctx->d_hist_mask, sizeof(unsigned) * words * f));
        CK(cudaMalloc((void**)&ctx->d_hist_gr, sizeof(int2) * ctx->maxP * f));
        CK(cudaHostAlloc((void**)&ctx->h_hist_pts, sizeof(float2) * ctx->maxP * f, cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&ctx->h_hist_ten, sizeof(int) * ctx->maxT * f, cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&ctx->h_hist_err, sizeof(float) * f, cudaHostAllocDefault));
        CK(cudaHostAlloc((void**)&ctx->h_hist_mask, sizeof(unsigned) * words * f, cudaHostAllocDefault));
        ctx->hist_cap = frames;
    }
    ctx->hist_frames = frames;          // 0 switches the recording off (the buffers stay)
    return TPOSE_OK;
}

extern "C" int tpose_cuda_history_frame(tpose_cuda_ctx* ctx, int frame, float* pts_xy, uint32_t* mask_words, int nwords,
                                        int32_t* ten_base, float* toterr)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (frame < 0 || frame >= ctx->hist_np_frames) return fail(ctx, TPOSE_ERR_STATE, "history_frame: the last run recorded no such frame");
    const size_t np = (size_t)ctx->hist_np, nt = (size_t)ctx->hist_nt, words = (nt + 31) / 32;
    if (pts_xy) memcpy(pts_xy, ctx->h_hist_pts + (size_t)frame * np, sizeof(float2) * np);
    if (ten_base) memcpy(ten_base, ctx->h_hist_ten + (size_t)frame * nt, sizeof(int) * nt);
    if (toterr) *toterr = ctx->h_hist_err[frame];
    if (mask_words) for (int i = 0; i < nwords; i++) mask_words[i] = (size_t)i < words ? ctx->h_hist_mask[(size_t)frame * words + i] : 0u;
    return TPOSE_OK;
}

// A caller that rolls the run back to recorded frame `frame` (because its own per-frame code edited
// the mesh there) must also see pgradbuf as that frame left it: a vertex that the edit leaves without
// triangles keeps that entry from then on (triangle.vs:93-98 resets referenced vertices only).
extern "C" int tpose_cuda_history_restore_gradient(tpose_cuda_ctx* ctx, int frame)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (frame < 0 || frame >= ctx->hist_np_frames) return fail(ctx, TPOSE_ERR_STATE, "history_restore_gradient: the last run recorded no such frame");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->gr, ctx->d_hist_gr + (size_t)frame * ctx->hist_np, sizeof(int2) * (size_t)ctx->hist_np,
                       cudaMemcpyDeviceToDevice, ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_toterr(tpose_cuda_ctx* ctx, float toterr)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (ctx->run_cache_valid && ctx->run_toterr == toterr) return TPOSE_OK;   // the device already holds this value
    CK(cudaSetDevice(ctx->device));
    const bool keep = ctx->run_cache_valid;
    int rc = upload_bytes(ctx, ctx->err_state, &toterr, sizeof(float));
    if (rc) return rc;
    ctx->run_cache_valid = keep; ctx->run_toterr = toterr;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_get_toterr(tpose_cuda_ctx* ctx, float* toterr)
{
    if (!ctx || !toterr) return TPOSE_ERR_ARG;
    if (ctx->run_cache_valid) { *toterr = ctx->run_toterr; return TPOSE_OK; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(toterr, ctx->err_state, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_watch(tpose_cuda_ctx* ctx, const int32_t* tri_xyzw, const int32_t* halfedges, int nt)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!tri_xyzw) { ctx->watching = false; return TPOSE_OK; }
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "watch: no mesh uploaded");
    if (!halfedges || nt != ctx->NT) return fail(ctx, TPOSE_ERR_ARG, "watch: needs halfedges and the uploaded mesh's NT");
    for (int t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++) {
            const int v = tri_xyzw[4 * t + k], h = halfedges[3 * t + k];
            if (v < 0 || v >= ctx->NP || h < -1 || h >= 3 * nt) return fail(ctx, TPOSE_ERR_ARG, "watch: index out of range");
        }
    CK(cudaSetDevice(ctx->device));
    const bool keep = ctx->run_cache_valid;
    // wtri | whe packed back to back at this NT, staged together: one H2D copy
    ctx->whe = (int*)(ctx->wbuf + sizeof(int4) * (size_t)nt);
    const size_t b_tri = sizeof(int4) * (size_t)nt, b_he = sizeof(int) * 3 * (size_t)nt;
    if (b_tri + b_he <= ctx->h_stage_cap) {
        const int slot = ctx->h_stage_next;
        ctx->h_stage_next = (slot + 1) % tpose_cuda_ctx::kStageSlots;
        CK(cudaEventSynchronize(ctx->h_stage_ev[slot]));
        memcpy(ctx->h_stage[slot], tri_xyzw, b_tri);
        memcpy(ctx->h_stage[slot] + b_tri, halfedges, b_he);
        CK(cudaMemcpyAsync(ctx->wbuf, ctx->h_stage[slot], b_tri + b_he, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->h_stage_ev[slot], ctx->stream));
    } else {
        int rc = upload_bytes(ctx, ctx->wtri, tri_xyzw, b_tri);
        if (!rc) rc = upload_bytes(ctx, ctx->whe, halfedges, b_he);
        if (rc) return rc;
    }
    ctx->run_cache_valid = keep;
    ctx->watching = true;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_watch_mask(tpose_cuda_ctx* ctx, uint32_t* mask_words, int nwords)
{
    if (!ctx || !mask_words || nwords < 0) return TPOSE_ERR_ARG;
    if (!ctx->run_mask_valid) return fail(ctx, TPOSE_ERR_STATE, "watch_mask: the last call was not a tpose_cuda_run that stopped on TPOSE_TRIG_TOPOLOGY");
    const int have = (ctx->NT + 31) / 32;
    for (int i = 0; i < nwords; i++) mask_words[i] = i < have ? ctx->h_run_mask[i] : 0u;
    return TPOSE_OK;
}

// ---- batches ---------------------------------------------------------------------------------

static int batch_prepare(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, EvalShape* shape)
{
    if (!ctxs || n < 1 || !ctxs[0]) return TPOSE_ERR_ARG;
    tpose_cuda_ctx* lead = ctxs[0];
    tpose_cuda_ctx* ctx = lead;   // for CK
    std::vector<EvalJob> jobs((size_t)n);
    int mxp = 0, mxe = 0, mxt = 0, mxn = 0; double rows = 0.0, items = 0.0; bool narrow = true;
    for (int i = 0; i < n; i++) {
        const int slot = slots ? slots[i] : 0;
        int rc = check_ready(ctxs[i], slot, flavour);
        if (rc) { if (ctxs[i] && ctxs[i] != lead) lead->err = "batch: job " + std::to_string(i) + ": " + ctxs[i]->err; return rc; }
        if (ctxs[i]->device != lead->device) return fail(lead, TPOSE_ERR_ARG, "batch: contexts live on different devices");
        if (ctxs[i]->path != lead->path) return fail(lead, TPOSE_ERR_ARG, "batch: contexts use different kernel paths");
        for (int k = 0; k < i; k++) if (ctxs[k] == ctxs[i]) return fail(lead, TPOSE_ERR_ARG, "batch: a context appears twice");
        if (ctxs[i] != lead) cudaStreamSynchronize(ctxs[i]->stream);   // its uploads are done before we touch its buffers
        ctxs[i]->run_cache_valid = false;
        rc = rebuild_edges(ctxs[i]);
        if (rc) { if (ctxs[i] != lead) lead->err = "batch: job " + std::to_string(i) + ": " + ctxs[i]->err; return rc; }
        jobs[(size_t)i] = make_job(ctxs[i], slot, flavour);
        jobs[(size_t)i].flip = ctxs[i]->cur ^ lead->cur;   // contexts need not have done the same number of steps
        mxp = ctxs[i]->NP > mxp ? ctxs[i]->NP : mxp;
        mxe = ctxs[i]->NE > mxe ? ctxs[i]->NE : mxe;
        mxn = std::max(mxn, ctxs[i]->NT);
        mxt = std::max(mxt, ctxs[i]->img[slot].ntx * ctxs[i]->img[slot].nty);
        items += (double)ctxs[i]->NE * kEdgeMoves;
        narrow = narrow && narrow_sums(ctxs[i]->img[slot].W, ctxs[i]->img[slot].H);
        rows += mean_edge_rows(ctxs[i]->NT, ctxs[i]->img[slot].H);
    }
    if (lead->batch_cap < n) {
        cudaFree(lead->batch_jobs); lead->batch_jobs = nullptr; lead->batch_cap = 0;
        CK(cudaMalloc((void**)&lead->batch_jobs, sizeof(EvalJob) * n));
        lead->batch_cap = n;
    }
    CK(cudaMemcpyAsync(lead->batch_jobs, jobs.data(), sizeof(EvalJob) * n, cudaMemcpyHostToDevice, lead->stream));
    CK(cudaStreamSynchronize(lead->stream));   // `jobs` is a local
    *shape = make_shape(valid_group(lead->group) ? lead->group : group_for(rows / n, items), mxp, mxe, mxn, narrow, lead->path, mxt, false);
    return TPOSE_OK;
}

extern "C" int tpose_cuda_batch_step(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, int nsteps)
{
    EvalShape shape{};
    int rc = batch_prepare(ctxs, n, slots, flavour, &shape);
    if (rc) return rc;
    tpose_cuda_ctx* ctx = ctxs[0];
    if (nsteps < 0) return fail(ctx, TPOSE_ERR_ARG, "batch_step: nsteps < 0");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    rc = issue_steps(ctx, ctx->batch_jobs, n, shape, fused_flags(flavour), ctx->cur ^ 1, nsteps);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    // the batch ran on the lead context's stream; whatever the caller does next with another job's
    // context (retrieve, upload, step) goes to that context's own stream and must come after it
    CK(cudaEventRecord(ctx->ev_batch, ctx->stream));
    for (int i = 1; i < n; i++)
        if (ctxs[i]->stream != ctx->stream) CK(cudaStreamWaitEvent(ctxs[i]->stream, ctx->ev_batch, 0));
    for (int i = 0; i < n; i++) ctxs[i]->cur = (ctxs[i]->cur + nsteps) & 1;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_batch_run(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, int max_steps,
                                    double rel_tol, int* steps_done, uint32_t* trigger_flags)
{
    EvalShape shape{};
    int rc = batch_prepare(ctxs, n, slots, flavour, &shape);
    if (rc) return rc;
    tpose_cuda_ctx* ctx = ctxs[0];
    if (max_steps < 1) return fail(ctx, TPOSE_ERR_ARG, "batch_run: max_steps < 1");
    for (int i = 0; i < n; i++) CK(cudaMemsetAsync(ctxs[i]->run_state, 0, sizeof(int) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int first = ctx->cur ^ 1;
    rc = issue_run(ctx, ctx->batch_jobs, n, shape, first, flavour, max_steps, rel_tol);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<int> rs((size_t)4 * n);
    for (int i = 0; i < n; i++)
        CK(cudaMemcpyAsync(&rs[(size_t)4 * i], ctxs[i]->run_state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        const int done = rs[(size_t)4 * i + RS_STEPS];
        if (done > 0) ctxs[i]->cur = ((ctxs[i]->cur ^ 1) + done - 1) & 1;
        if (steps_done) steps_done[i] = done;
        if (trigger_flags) trigger_flags[i] = (uint32_t)rs[(size_t)4 * i + RS_TRIG];
    }
    return TPOSE_OK;
}

// ---- read back ------------------------------------------------------------------------------

extern "C" int tpose_cuda_retrieve(tpose_cuda_ctx* ctx, int32_t* ten13, int32_t* cn13, int32_t* ca13, float* pts_xy)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "retrieve: no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    const size_t nv = (size_t)kVariants * ctx->NT;
    const int c = ctx->cur;
    if (ten13) CK(cudaMemcpyAsync(ten13, ctx->ten[c], sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn13) CK(cudaMemcpyAsync(cn13, ctx->cn[c], sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca13) CK(cudaMemcpyAsync(ca13, ctx->ca[c], sizeof(int4) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (pts_xy) CK(cudaMemcpyAsync(pts_xy, ctx->pts, sizeof(float2) * ctx->NP, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_retrieve_base(tpose_cuda_ctx* ctx, int32_t* ten, int32_t* cn, int32_t* ca)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "retrieve_base: no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->NT;
    const int c = ctx->cur;
    if (ten) CK(cudaMemcpyAsync(ten, ctx->ten[c], sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn) CK(cudaMemcpyAsync(cn, ctx->cn[c], sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca) CK(cudaMemcpyAsync(ca, ctx->ca[c], sizeof(int4) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_retrieve_gradient(tpose_cuda_ctx* ctx, int32_t* gr_xy)
{
    if (!ctx || !gr_xy) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "retrieve_gradient: no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(gr_xy, ctx->gr, sizeof(int2) * ctx->NP, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

// element size, capacity (elements) and device address of one binding point
static bool buffer_info(tpose_cuda_ctx* ctx, int which, size_t* elem, size_t* cap, void** dev)
{
    const size_t nv = (size_t)kVariants * ctx->maxT;
    const int c = ctx->cur;
    switch (which) {
    case TPOSE_BUF_POINTS:   *elem = sizeof(float2); *cap = ctx->maxP; *dev = ctx->pts; return true;
    case TPOSE_BUF_INDEX:    *elem = sizeof(int4);   *cap = ctx->maxT; *dev = ctx->tri; return true;
    case TPOSE_BUF_COLACC:   *elem = sizeof(int4);   *cap = nv; *dev = ctx->ca[c]; return true;
    case TPOSE_BUF_COLNUM:   *elem = sizeof(int);    *cap = nv; *dev = ctx->cn[c]; return true;
    case TPOSE_BUF_TENERGY:  *elem = sizeof(int);    *cap = nv; *dev = ctx->ten[c]; return true;
    case TPOSE_BUF_GRADIENT: *elem = sizeof(int2);   *cap = ctx->maxP; *dev = ctx->gr; return true;
    case TPOSE_BUF_PENERGY:
    case TPOSE_BUF_NRING:    *elem = sizeof(int);    *cap = (size_t)TPOSE_REF_MAXT; *dev = nullptr; return true;
    default: return false;
    }
}

extern "C" int tpose_cuda_read(tpose_cuda_ctx* ctx, int which, size_t first, size_t count, void* out)
{
    if (!ctx) return TPOSE_ERR_ARG;
    size_t elem, cap; void* dev;
    if (!out || !buffer_info(ctx, which, &elem, &cap, &dev)) return fail(ctx, TPOSE_ERR_ARG, "read: bad buffer or null output");
    if (first + count > cap) return fail(ctx, TPOSE_ERR_CAPACITY, "read: range exceeds the buffer");
    if (count == 0) return TPOSE_OK;
    if (!dev) { memset(out, 0, elem * count); return TPOSE_OK; }   // penergy / nring: dead in the reference
    if (which == TPOSE_BUF_POINTS && ctx->run_cache_valid && first + count <= (size_t)ctx->NP) {
        memcpy(out, ctx->h_run_pts + 2 * first, elem * count);     // fetched together with the run's verdict
        return TPOSE_OK;
    }
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, (char*)dev + elem * first, elem * count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_write(tpose_cuda_ctx* ctx, int which, size_t first, size_t count, const void* in)
{
    if (!ctx) return TPOSE_ERR_ARG;
    size_t elem, cap; void* dev;
    if (!in || !buffer_info(ctx, which, &elem, &cap, &dev)) return fail(ctx, TPOSE_ERR_ARG, "write: bad buffer or null input");
    if (first + count > cap) return fail(ctx, TPOSE_ERR_CAPACITY, "write: range exceeds the buffer");
    if (count == 0 || !dev) return TPOSE_OK;
    ctx->run_cache_valid = false;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync((char*)dev + elem * first, in, elem * count, cudaMemcpyHostToDevice, ctx->stream));
    if (which == TPOSE_BUF_INDEX) {   // keep the host mirror the edge view is derived from
        if (ctx->tri_host.size() < (size_t)4 * (first + count)) ctx->tri_host.resize((size_t)4 * (first + count), 0);
        memcpy(ctx->tri_host.data() + 4 * first, in, elem * count);
        ctx->edges_dirty = true;
    }
    if (which == TPOSE_BUF_COLACC)   // both buffer sets see uploaded colours (warp flavour reads them every frame)
        CK(cudaMemcpyAsync((char*)ctx->ca[ctx->cur ^ 1] + elem * first, in, elem * count, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_raster_ids(tpose_cuda_ctx* ctx, int slot, int32_t* ids_host)
{
    int rc = check_ready(ctx, slot, TPOSE_TRIANGULATE);
    if (rc) return rc;
    if (!ids_host) return fail(ctx, TPOSE_ERR_ARG, "raster_ids: null output");
    const ImageSlot& s = ctx->img[slot];
    const size_t bytes = (size_t)s.W * s.H * sizeof(int);
    if (ctx->ids_cap < bytes) { cudaFree(ctx->ids); ctx->ids = nullptr; CK(cudaMalloc((void**)&ctx->ids, bytes)); ctx->ids_cap = bytes; }
    CK(cudaMemsetAsync(ctx->ids, 0xFF, bytes, ctx->stream));
    const int warps_per_block = 8;
    k_raster_ids<<<(ctx->NT + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, ctx->stream>>>(
        ctx->pts, ctx->tri, ctx->NT, ctx->ratio, s.W, s.H, ctx->ids);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ids_host, ctx->ids, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

static int render_common(tpose_cuda_ctx* ctx, int raster_w, int raster_h, const float* other_pts_xy, float s, int lines, uint8_t* rgba_host);

extern "C" int tpose_cuda_render(tpose_cuda_ctx* ctx, int raster_w, int raster_h, const float* other_pts_xy, float s, uint8_t* rgba_host)
{
    return render_common(ctx, raster_w, raster_h, other_pts_xy, s, 0, rgba_host);
}

// the viewer with its `n` key down (view/main.cpp:58-69,112-121): the triangles' outlines in white on top
extern "C" int tpose_cuda_render_wire(tpose_cuda_ctx* ctx, int raster_w, int raster_h, const float* other_pts_xy, float s, uint8_t* rgba_host)
{
    return render_common(ctx, raster_w, raster_h, other_pts_xy, s, 1, rgba_host);
}

static int render_common(tpose_cuda_ctx* ctx, int raster_w, int raster_h, const float* other_pts_xy, float s, int lines, uint8_t* rgba_host)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "render: no mesh uploaded");
    if (!rgba_host || raster_w <= 0 || raster_h <= 0 || raster_w > TPOSE_MAX_WIDTH || raster_h > TPOSE_MAX_HEIGHT)
        return fail(ctx, TPOSE_ERR_ARG, "render: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const size_t npix = (size_t)raster_w * raster_h;
    // scratch: ids (int) | colours (uchar4) | mixed points (float2)
    const size_t need = npix * (sizeof(int) + sizeof(uchar4)) + sizeof(float2) * (size_t)ctx->NP;
    if (ctx->ids_cap < need) { cudaFree(ctx->ids); ctx->ids = nullptr; ctx->ids_cap = 0; CK(cudaMalloc((void**)&ctx->ids, need)); ctx->ids_cap = need; }
    uchar4* colours = (uchar4*)(ctx->ids + npix);
    float2* mixed = (float2*)(colours + npix);
    const float2* pos = ctx->pts;
    if (other_pts_xy) {
        // the other embedding goes through the (unused during display) gradient buffer's neighbour: stage it after `mixed`
        CK(cudaMemcpyAsync(mixed, other_pts_xy, sizeof(float2) * ctx->NP, cudaMemcpyHostToDevice, ctx->stream));
        k_mix_points<<<(ctx->NP + 127) / 128, 128, 0, ctx->stream>>>(ctx->pts, mixed, ctx->NP, s, mixed);
        ctx->launches++;
        pos = mixed;
    }
    CK(cudaMemsetAsync(ctx->ids, 0xFF, npix * sizeof(int), ctx->stream));
    const int warps_per_block = 8;
    k_raster_ids<<<(ctx->NT + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, ctx->stream>>>(
        pos, ctx->tri, ctx->NT, ctx->ratio, raster_w, raster_h, ctx->ids);
    k_flat_colors<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(ctx->ids, ctx->ca[ctx->cur], (int)npix, colours);
    ctx->launches += 2;
    if (lines) {
        k_wire<<<(3 * ctx->NT + 127) / 128, 128, 0, ctx->stream>>>(pos, ctx->tri, ctx->NT, ctx->ratio, raster_w, raster_h, colours);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgba_host, colours, npix * sizeof(uchar4), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

// A few instances on their own (see k_eval_list): entries = n x (v0, v1, v2, k).
extern "C" int tpose_cuda_eval_list(tpose_cuda_ctx* ctx, int slot, const int32_t* entries_v0v1v2k, int n, int reset_gradient,
                                    int32_t* ten_out, int32_t* cn_out, int32_t* ca_out_rgba)
{
    int rc = check_ready(ctx, slot, TPOSE_TRIANGULATE);
    if (rc) return rc;
    if (n < 0 || (n > 0 && !entries_v0v1v2k)) return fail(ctx, TPOSE_ERR_ARG, "eval_list: bad arguments");
    ImageSlot& im = ctx->img[slot];
    if (!im.table) return fail(ctx, TPOSE_ERR_STATE, "eval_list: needs the row-prefix table (contexts on the tile path do not keep one)");
    if (reset_gradient) {
        // the mode-1 draw this call stands in for resets pgradbuf of every referenced vertex (triangle.vs:93-98)
        if ((rc = rebuild_edges(ctx)) != 0) return rc;
        k_zero_gradient<<<(ctx->NP + 255) / 256, 256, 0, ctx->stream>>>(ctx->gr, ctx->NP, ctx->corner_off);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    if (n == 0) return TPOSE_OK;
    for (int i = 0; i < n; i++) {
        const int32_t* e = entries_v0v1v2k + 4 * (size_t)i;
        if (e[0] < 0 || e[0] >= ctx->NP || e[1] < 0 || e[1] >= ctx->NP || e[2] < 0 || e[2] >= ctx->NP || e[3] < 0 || e[3] >= kVariants)
            return fail(ctx, TPOSE_ERR_ARG, "eval_list: entry references a vertex outside [0,NP) or a variant outside [0,13)");
    }
    // scratch: entries (int4) | ten | cn | ca (int4), grow-only
    const size_t need = (size_t)n * (sizeof(int4) * 2 + sizeof(int) * 2);
    if (ctx->list_cap < need) {
        cudaFree(ctx->list_buf); ctx->list_buf = nullptr; ctx->list_cap = 0;
        CK(cudaMalloc((void**)&ctx->list_buf, need + need / 2));
        ctx->list_cap = need + need / 2;
    }
    int4* d_en = (int4*)ctx->list_buf; int4* d_ca = d_en + n; int* d_ten = (int*)(d_ca + n); int* d_cn = d_ten + n;
    const bool keep = ctx->run_cache_valid;
    rc = upload_bytes(ctx, d_en, entries_v0v1v2k, sizeof(int4) * (size_t)n);
    ctx->run_cache_valid = keep;                  // (nothing the cache mirrors changed)
    if (rc) return rc;
    const int blocks = (n + kListThreads / 32 - 1) / (kListThreads / 32);
    const float dp = dp_for(ctx->NT, TPOSE_TRIANGULATE);
    if (narrow_sums(im.W, im.H))
        k_eval_list<unsigned><<<blocks, kListThreads, 0, ctx->stream>>>(im.table, im.pitch, im.W, im.H, ctx->pts, ctx->ratio, dp, d_en, n, d_ten, d_cn, d_ca);
    else
        k_eval_list<unsigned long long><<<blocks, kListThreads, 0, ctx->stream>>>(im.table, im.pitch, im.W, im.H, ctx->pts, ctx->ratio, dp, d_en, n, d_ten, d_cn, d_ca);
    ctx->launches++;
    CK(cudaGetLastError());
    if (ten_out) CK(cudaMemcpyAsync(ten_out, d_ten, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn_out) CK(cudaMemcpyAsync(cn_out, d_cn, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca_out_rgba) CK(cudaMemcpyAsync(ca_out_rgba, d_ca, sizeof(int4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_last_elapsed_ms(tpose_cuda_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_step_scale(tpose_cuda_ctx* ctx, float scale)
{
    if (!ctx || !(scale > 0.0f)) return TPOSE_ERR_ARG;
    ctx->step_scale = scale;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_kernel_timing(tpose_cuda_ctx* ctx, int on)
{
    if (!ctx) return TPOSE_ERR_ARG;
    ctx->time_kernels = on != 0;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_last_kernel_ms(tpose_cuda_ctx* ctx, float* edges_ms, float* corners_ms)
{
    if (!ctx || !edges_ms || !corners_ms) return TPOSE_ERR_ARG;
    if (!ctx->mid_valid) return fail(ctx, TPOSE_ERR_STATE, "last_kernel_ms: needs tpose_cuda_set_kernel_timing(ctx, 1) and a single tpose_cuda_step(.., 1) as the last call");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(edges_ms, ctx->path == kPathTiles ? ctx->ev_mid0 : ctx->ev0, ctx->ev_mid));   // k_edges, or k_tiles alone
    CK(cudaEventElapsedTime(corners_ms, ctx->ev_mid, ctx->ev1));
    return TPOSE_OK;
}

// tile path: the step's first kernel (k_setup: snapped end points, tile ranges, zeroed records); 0 on the table path
extern "C" int tpose_cuda_last_setup_ms(tpose_cuda_ctx* ctx, float* setup_ms)
{
    if (!ctx || !setup_ms) return TPOSE_ERR_ARG;
    if (!ctx->mid_valid) return fail(ctx, TPOSE_ERR_STATE, "last_setup_ms: needs tpose_cuda_set_kernel_timing(ctx, 1) and a single tpose_cuda_step(.., 1) as the last call");
    *setup_ms = 0.0f;
    if (ctx->path != kPathTiles) return TPOSE_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(setup_ms, ctx->ev0, ctx->ev_mid0));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_kernel_path(const tpose_cuda_ctx* ctx) { return ctx ? ctx->path : -1; }

#ifdef TPOSE_TIMELINE
// development aid: out[2][6][8192] globaltimer stamps (ns) per warp; reset != 0 zeroes them
extern "C" int tpose_cuda_debug_timeline(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    if (out) cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 2 * 6 * kTlWarps);
    if (reset) {
        void* p = nullptr;
        cudaGetSymbolAddress(&p, g_timeline);
        cudaMemset(p, 0, sizeof(unsigned long long) * 2 * 6 * kTlWarps);
        cudaDeviceSynchronize();
    }
    return 0;
}
#endif

extern "C" int tpose_cuda_sync(tpose_cuda_ctx* ctx)
{
    if (!ctx) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int64_t tpose_cuda_launch_count(const tpose_cuda_ctx* ctx)
{
    return ctx ? ctx->launches : 0;
}

extern "C" int tpose_cuda_host_alloc(size_t bytes, void** out)
{
    if (!out || bytes == 0) return TPOSE_ERR_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? TPOSE_OK : TPOSE_ERR_CUDA;
}

extern "C" int tpose_cuda_host_free(void* p)
{
    return cudaFreeHost(p) == cudaSuccess ? TPOSE_OK : TPOSE_ERR_CUDA;
}

// reads `n16` 16-byte words and keeps the compiler from dropping the loads
__global__ void k_read_sink(const uint4* __restrict__ p, size_t n16, unsigned* sink)
{
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldcs(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

// Measurement aid: evicts everything the context's kernels could find in L2.  A buffer larger than L2 is
// written and then read back: after the write alone L2 would be full of DIRTY lines of the flush buffer, and
// the timed kernel would pay their write-back on every miss -- an artefact of the flush, not of the workload.
// After the read pass L2 holds clean lines of the buffer's tail.
extern "C" int tpose_cuda_flush_l2(tpose_cuda_ctx* ctx, size_t bytes)
{
    if (!ctx || bytes == 0) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    bytes = (bytes + 15) & ~(size_t)15;
    if (ctx->flush_cap < bytes + 16) { cudaFree(ctx->flush); ctx->flush = nullptr; CK(cudaMalloc((void**)&ctx->flush, bytes + 16)); ctx->flush_cap = bytes + 16; }
    CK(cudaMemsetAsync(ctx->flush, 0x5a, bytes, ctx->stream));
    k_read_sink<<<148 * 8, 256, 0, ctx->stream>>>((const uint4*)ctx->flush, bytes / 16, (unsigned*)(ctx->flush + bytes));
    CK(cudaGetLastError());
    return TPOSE_OK;
}

// ---- multiview (source/multiview.hpp) -----------------------------------------------------------

namespace {
// The multiview entry points take no context: their state -- last error, last kernel time and a
// grow-only device workspace with its own stream and events per device -- is per host thread, so
// they are thread-compatible like the context API, and a call costs no cudaMalloc / cudaFree.
thread_local std::string g_mview_error;
thread_local float g_mview_ms = 0.0f;

struct MviewWs {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    unsigned char* buf = nullptr; size_t cap = 0, used = 0;
    ~MviewWs()
    {
        if (device < 0) return;
        // (thread exit: release quietly; the primary context may already be gone at process exit)
        if (cudaSetDevice(device) == cudaSuccess) {
            if (buf) cudaFree(buf);
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            if (stream) cudaStreamDestroy(stream);
        }
        (void)cudaGetLastError();
    }
};
thread_local std::vector<std::unique_ptr<MviewWs>> g_mview_ws;      // released when the thread ends

int mview_fail(int code, const std::string& msg) { g_mview_error = msg; return code; }

#define MCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return mview_fail(TPOSE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// the calling thread's workspace on `device`, holding at least `bytes`
int mview_workspace(int device, size_t bytes, MviewWs** out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        (void)cudaGetLastError();
        return mview_fail(TPOSE_ERR_NODEVICE, "no usable CUDA device (this library has no CPU fallback)");
    }
    MCK(cudaSetDevice(device));
    MviewWs* ws = nullptr;
    for (auto& w : g_mview_ws) if (w->device == device) ws = w.get();
    if (!ws) {
        auto fresh = std::make_unique<MviewWs>();      // (dropped again if any of its resources cannot be created)
        fresh->device = device;
        MCK(cudaStreamCreateWithFlags(&fresh->stream, cudaStreamNonBlocking));
        MCK(cudaEventCreate(&fresh->e0)); MCK(cudaEventCreate(&fresh->e1));
        g_mview_ws.push_back(std::move(fresh));
        ws = g_mview_ws.back().get();
    }
    if (ws->cap < bytes) {
        MCK(cudaStreamSynchronize(ws->stream));
        if (ws->buf) cudaFree(ws->buf);
        ws->buf = nullptr; ws->cap = 0;
        const size_t want = bytes + bytes / 2 + 4096;
        MCK(cudaMalloc((void**)&ws->buf, want));
        ws->cap = want;
    }
    ws->used = 0;
    *out = ws;
    return TPOSE_OK;
}

template <typename T> T* carve(MviewWs* ws, size_t count)
{
    T* p = reinterpret_cast<T*>(ws->buf + ws->used);
    ws->used += (count * sizeof(T) + 255) & ~(size_t)255;
    return p;
}
size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

bool offsets_ok(const int* offsets, int nbatch)
{
    if (!offsets || nbatch < 1 || offsets[0] != 0) return false;
    for (int i = 0; i < nbatch; i++) if (offsets[i + 1] < offsets[i]) return false;
    return true;
}
}  // namespace

extern "C" const char* tpose_cuda_mview_last_error(void) { return g_mview_error.c_str(); }
extern "C" float tpose_cuda_mview_last_ms(void) { return g_mview_ms; }

static int fmatrix_common(int device, const float* A_xy, const float* B_xy, const int* offsets, int nbatch,
                          const float* weights, int iters, float* F9_out, const char* who)
{
    if (!A_xy || !B_xy || !F9_out || !offsets_ok(offsets, nbatch) || iters < 0)
        return mview_fail(TPOSE_ERR_ARG, std::string(who) + ": bad arguments");
    const size_t n = (size_t)offsets[nbatch];
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, 2 * padded(n * sizeof(float2)) + padded((nbatch + 1) * sizeof(int)) +
                                         padded((size_t)nbatch * 9 * sizeof(float)) + padded(n * sizeof(float)) + 1024, &ws);
    if (rc) return rc;
    float2* dA = carve<float2>(ws, n); float2* dB = carve<float2>(ws, n);
    int* dO = carve<int>(ws, nbatch + 1); float* dF = carve<float>(ws, (size_t)nbatch * 9);
    float* dW = weights ? carve<float>(ws, n) : nullptr;
    cudaStream_t st = ws->stream;
    MCK(cudaMemcpyAsync(dA, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dB, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dO, offsets, (nbatch + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    if (weights) MCK(cudaMemcpyAsync(dW, weights, n * sizeof(float), cudaMemcpyHostToDevice, st));
    MCK(cudaEventRecord(ws->e0, st));
    k_f8point<<<nbatch, kF8Threads, 0, st>>>(dA, dB, dO, dW, iters, dF);
    MCK(cudaEventRecord(ws->e1, st));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(F9_out, dF, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&g_mview_ms, ws->e0, ws->e1);
    return TPOSE_OK;
}

extern "C" int tpose_cuda_f8point(int device, const float* A_xy, const float* B_xy, const int* offsets, int nbatch, float* F9_out)
{
    return fmatrix_common(device, A_xy, B_xy, offsets, nbatch, nullptr, 0, F9_out, "f8point");
}

extern "C" int tpose_cuda_fsampson(int device, const float* A_xy, const float* B_xy, const float* weights, const int* offsets,
                                   int nbatch, int iterations, float* F9_out)
{
    return fmatrix_common(device, A_xy, B_xy, offsets, nbatch, weights, iterations, F9_out, "fsampson");
}

extern "C" int tpose_cuda_triangulate(int device, const float* F9, const float* K9, const float* A_xy, const float* B_xy,
                                      const int* offsets, int nbatch, int optimal_correction, float* X4_out,
                                      float* A_corr, float* B_corr)
{
    if (!F9 || !K9 || !A_xy || !B_xy || !X4_out || !offsets_ok(offsets, nbatch)) return mview_fail(TPOSE_ERR_ARG, "triangulate: bad arguments");
    const size_t n = (size_t)offsets[nbatch];
    int longest = 0;
    for (int i = 0; i < nbatch; i++) longest = std::max(longest, offsets[i + 1] - offsets[i]);
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, 4 * padded(n * sizeof(float2)) + padded((nbatch + 1) * sizeof(int)) +
                                         2 * padded((size_t)nbatch * 9 * sizeof(float)) + padded(n * sizeof(float4)) + 1024, &ws);
    if (rc) return rc;
    float2* dA = carve<float2>(ws, n); float2* dB = carve<float2>(ws, n);
    int* dO = carve<int>(ws, nbatch + 1);
    float* dF = carve<float>(ws, (size_t)nbatch * 9); float* dK = carve<float>(ws, (size_t)nbatch * 9);
    float4* dX = carve<float4>(ws, n);
    float2* dAc = A_corr ? carve<float2>(ws, n) : nullptr;
    float2* dBc = B_corr ? carve<float2>(ws, n) : nullptr;
    cudaStream_t st = ws->stream;
    MCK(cudaMemcpyAsync(dA, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dB, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dO, offsets, (nbatch + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dF, F9, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dK, K9, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    MCK(cudaEventRecord(ws->e0, st));
    if (longest > 0) {
        const dim3 grid((longest + 127) / 128, nbatch);
        k_triangulate<<<grid, 128, 0, st>>>(dF, dK, dA, dB, dO, optimal_correction, dX, dAc, dBc);
    }
    MCK(cudaEventRecord(ws->e1, st));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(X4_out, dX, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (A_corr) MCK(cudaMemcpyAsync(A_corr, dAc, n * sizeof(float2), cudaMemcpyDeviceToHost, st));
    if (B_corr) MCK(cudaMemcpyAsync(B_corr, dBc, n * sizeof(float2), cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&g_mview_ms, ws->e0, ws->e1);
    return TPOSE_OK;
}

extern "C" int tpose_cuda_select_matches(int device, const float* originA_xy, const float* pointsA_xy, const float* warp_originA_xy, int npa,
                                         const float* pointsB_xy, const float* originB_xy, const float* warp_originB_xy, int npb,
                                         float ratio, double max_dist, float* X_xy, float* Y_xy, float* weights, float* dist, int* n_out)
{
    if (npa < 0 || npb < 0 || !X_xy || !Y_xy || !weights || !n_out || (npa > 0 && (!originA_xy || !pointsA_xy || !warp_originA_xy)) ||
        (npb > 0 && (!pointsB_xy || !originB_xy || !warp_originB_xy)) || !(ratio > 0.0f))
        return mview_fail(TPOSE_ERR_ARG, "select_matches: bad arguments");
    *n_out = 0;
    const size_t n = (size_t)npa + (size_t)npb;
    if (n == 0) return TPOSE_OK;
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, 8 * padded(n * sizeof(float2)) + 2 * padded(n * sizeof(float)) + 2048, &ws);
    if (rc) return rc;
    float2* dOA = carve<float2>(ws, npa); float2* dPA = carve<float2>(ws, npa); float2* dWA = carve<float2>(ws, npa);
    float2* dPB = carve<float2>(ws, npb); float2* dOB = carve<float2>(ws, npb); float2* dWB = carve<float2>(ws, npb);
    float2* dX = carve<float2>(ws, n); float2* dY = carve<float2>(ws, n); float* dW = carve<float>(ws, n); float* dD = carve<float>(ws, n);
    int* dN = carve<int>(ws, 1);
    cudaStream_t st = ws->stream;
    if (npa) {
        MCK(cudaMemcpyAsync(dOA, originA_xy, sizeof(float2) * npa, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dPA, pointsA_xy, sizeof(float2) * npa, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dWA, warp_originA_xy, sizeof(float2) * npa, cudaMemcpyHostToDevice, st));
    }
    if (npb) {
        MCK(cudaMemcpyAsync(dPB, pointsB_xy, sizeof(float2) * npb, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dOB, originB_xy, sizeof(float2) * npb, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dWB, warp_originB_xy, sizeof(float2) * npb, cudaMemcpyHostToDevice, st));
    }
    k_select_matches<<<1, kSelectThreads, 0, st>>>(dOA, dPA, dWA, npa, dPB, dOB, dWB, npb, ratio, max_dist, dX, dY, dW, dist ? dD : nullptr, dN);
    MCK(cudaGetLastError());
    int kept = 0;
    MCK(cudaMemcpyAsync(&kept, dN, sizeof(int), cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    if (kept > 0) {
        MCK(cudaMemcpyAsync(X_xy, dX, sizeof(float2) * kept, cudaMemcpyDeviceToHost, st));
        MCK(cudaMemcpyAsync(Y_xy, dY, sizeof(float2) * kept, cudaMemcpyDeviceToHost, st));
        MCK(cudaMemcpyAsync(weights, dW, sizeof(float) * kept, cudaMemcpyDeviceToHost, st));
    }
    if (dist) MCK(cudaMemcpyAsync(dist, dD, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    *n_out = kept;
    return TPOSE_OK;
}

// geterr (source/triangulation.hpp:653-674) over an array of energies: the kernel-side routine of the runs on its own
extern "C" int tpose_cuda_geterr(int device, const int32_t* energies, int n, float* toterr_inout, float* relerr_out, int* below_tol_out, double rel_tol,
                                 int route)
{
    if (!energies || n < 0 || !toterr_inout) return mview_fail(TPOSE_ERR_ARG, "geterr: bad arguments");
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, padded(sizeof(int) * (size_t)(n > 0 ? n : 1)) + padded(4 * sizeof(float)) + padded(kGeterrScratchBytes) + 1024, &ws);
    if (rc) return rc;
    int* dE = carve<int>(ws, n > 0 ? n : 1); float* dS = carve<float>(ws, 4); int4* dScratch = carve<int4>(ws, kGeterrThreads);
    cudaStream_t st = ws->stream;
    const float st4[4] = {*toterr_inout, 0.0f, 0.0f, 0.0f};
    float out4[4];
    if (n > 0) MCK(cudaMemcpyAsync(dE, energies, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dS, st4, sizeof st4, cudaMemcpyHostToDevice, st));
    MCK(cudaEventRecord(ws->e0, st));
    k_geterr<<<1, kGeterrThreads, 0, st>>>(dE, n, dS, rel_tol, route < 0 ? -1 : (route ? 1 : 0), dScratch);
    MCK(cudaGetLastError());
    MCK(cudaEventRecord(ws->e1, st));
    MCK(cudaMemcpyAsync(out4, dS, sizeof out4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&g_mview_ms, ws->e0, ws->e1);         // tpose_cuda_mview_last_ms: the kernel alone
    *toterr_inout = out4[0];
    if (relerr_out) *relerr_out = out4[1];
    if (below_tol_out) *below_tol_out = out4[3] != 0.0f;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_carry(int device, const int32_t* tri_xyzw, int nt, const float* from_xy, const float* to_xy, int nv,
                                float* pts_xy, int npts, float ratio)
{
    if (!tri_xyzw || !from_xy || !to_xy || !pts_xy || nt < 0 || nv < 0 || npts < 0) return mview_fail(TPOSE_ERR_ARG, "carry: bad arguments");
    if (nt == 0 || nv == 0 || npts == 0) return TPOSE_OK;          // triangulation.hpp:467,495: nothing to do
    for (int t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++)
            if (tri_xyzw[4 * t + k] < 0 || tri_xyzw[4 * t + k] >= nv) return mview_fail(TPOSE_ERR_ARG, "carry: triangle references a vertex outside [0,nv)");
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, padded(sizeof(int4) * nt) + 2 * padded(sizeof(float2) * nv) + padded(sizeof(float2) * npts) + 1024, &ws);
    if (rc) return rc;
    int4* dT = carve<int4>(ws, nt); float2* dF = carve<float2>(ws, nv); float2* dTo = carve<float2>(ws, nv); float2* dP = carve<float2>(ws, npts);
    cudaStream_t st = ws->stream;
    MCK(cudaMemcpyAsync(dT, tri_xyzw, sizeof(int4) * nt, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dF, from_xy, sizeof(float2) * nv, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dTo, to_xy, sizeof(float2) * nv, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dP, pts_xy, sizeof(float2) * npts, cudaMemcpyHostToDevice, st));
    k_carry<<<(int)(((long long)npts * kCarryLanes + kCarryThreads - 1) / kCarryThreads), kCarryThreads, 0, st>>>(dT, nt, dF, dTo, dP, npts, ratio);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(pts_xy, dP, sizeof(float2) * npts, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_hdlt(int device, const float* PA12, const float* PB12, const float* A_xy, const float* B_xy,
                               int n, float* X4_out)
{
    if (!PA12 || !PB12 || !A_xy || !B_xy || !X4_out || n < 1) return mview_fail(TPOSE_ERR_ARG, "hdlt: bad arguments");
    MviewWs* ws = nullptr;
    int rc = mview_workspace(device, 2 * padded(n * sizeof(float2)) + 2 * padded(12 * sizeof(float)) + padded(n * sizeof(float4)) + 1024, &ws);
    if (rc) return rc;
    float2* dA = carve<float2>(ws, n); float2* dB = carve<float2>(ws, n);
    float* dPA = carve<float>(ws, 12); float* dPB = carve<float>(ws, 12); float4* dX = carve<float4>(ws, n);
    cudaStream_t st = ws->stream;
    MCK(cudaMemcpyAsync(dA, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dB, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dPA, PA12, 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dPB, PB12, 12 * sizeof(float), cudaMemcpyHostToDevice, st));
    k_hdlt<<<(n + 127) / 128, 128, 0, st>>>(dPA, dPB, dA, dB, n, dX);
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(X4_out, dX, n * sizeof(float4), cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return TPOSE_OK;
}

// ---- software/rgbd's plane-fit meshing core (rgbd/main.cpp:244-349) ------------------------------------------
extern "C" int tpose_cuda_rgbd_planefit(int device, const int32_t* tri_xyzw, int nt, const float* pts_xy, int np,
                                        const float* T9, const float* TB9, const float* pixels_xy, const float* points_xyzw,
                                        const float* normals_xyzw, int npoints, const float* intr_fx_fy_ppx_ppy,
                                        float* verts_xyzw, float* ring, int32_t* tri_counts)
{
    if (!tri_xyzw || !pts_xy || !T9 || !TB9 || !intr_fx_fy_ppx_ppy || !verts_xyzw || nt < 1 || np < 3 || npoints < 0 ||
        (npoints > 0 && (!pixels_xy || !points_xyzw || !normals_xyzw)))
        return mview_fail(TPOSE_ERR_ARG, "rgbd_planefit: bad arguments");
    for (int t = 0; t < nt; t++)
        for (int k = 0; k < 3; k++)
            if (tri_xyzw[4 * t + k] < 0 || tri_xyzw[4 * t + k] >= np) return mview_fail(TPOSE_ERR_ARG, "rgbd_planefit: triangle references a vertex outside [0,np)");
    MviewWs* ws = nullptr;
    const size_t npt = (size_t)std::max(npoints, 1);
    int rc = mview_workspace(device, padded(sizeof(int4) * nt) + padded(sizeof(float2) * np) + padded(sizeof(float2) * npt) + 2 * padded(sizeof(float4) * npt) +
                                         padded(sizeof(float4) * 3 * nt) + 2 * padded(sizeof(int) * nt) + padded(sizeof(float4) * np) + padded(sizeof(float) * np) + 4096, &ws);
    if (rc) return rc;
    int4* dT = carve<int4>(ws, nt); float2* dP = carve<float2>(ws, np); float2* dPx = carve<float2>(ws, npt);
    float4* dX = carve<float4>(ws, npt); float4* dN = carve<float4>(ws, npt); float4* dC = carve<float4>(ws, 3 * (size_t)nt);
    int* dHas = carve<int>(ws, nt); int* dCnt = carve<int>(ws, nt); float4* dV = carve<float4>(ws, np); float* dR = carve<float>(ws, np);
    cudaStream_t st = ws->stream;
    MCK(cudaMemcpyAsync(dT, tri_xyzw, sizeof(int4) * nt, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dP, pts_xy, sizeof(float2) * np, cudaMemcpyHostToDevice, st));
    if (npoints > 0) {
        MCK(cudaMemcpyAsync(dPx, pixels_xy, sizeof(float2) * npoints, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dX, points_xyzw, sizeof(float4) * npoints, cudaMemcpyHostToDevice, st));
        MCK(cudaMemcpyAsync(dN, normals_xyzw, sizeof(float4) * npoints, cudaMemcpyHostToDevice, st));
    }
    RgbdMap map; RgbdIntr intr{intr_fx_fy_ppx_ppy[0], intr_fx_fy_ppx_ppy[1], intr_fx_fy_ppx_ppy[2], intr_fx_fy_ppx_ppy[3]};
    memcpy(map.T, T9, sizeof map.T); memcpy(map.TB, TB9, sizeof map.TB);
    MCK(cudaEventRecord(ws->e0, st));
    k_rgbd_fit<<<nt, kRgbdThreads, 0, st>>>(dT, nt, dP, map, dPx, dX, dN, npoints, intr, dC, dHas, dCnt);
    k_rgbd_gather<<<(np + 127) / 128, 128, 0, st>>>(dT, nt, np, dC, dHas, dV, dR);
    MCK(cudaEventRecord(ws->e1, st));
    MCK(cudaGetLastError());
    MCK(cudaMemcpyAsync(verts_xyzw, dV, sizeof(float4) * np, cudaMemcpyDeviceToHost, st));
    if (ring) MCK(cudaMemcpyAsync(ring, dR, sizeof(float) * np, cudaMemcpyDeviceToHost, st));
    if (tri_counts) MCK(cudaMemcpyAsync(tri_counts, dCnt, sizeof(int) * nt, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&g_mview_ms, ws->e0, ws->e1);
    return TPOSE_OK;
}

// capi.cu -- the C ABI of include/tpose_cuda.h over the sm_100a kernels.
// There is no CPU path in this library: without a CUDA device every entry point fails.
#include "../../include/tpose_cuda.h"
#include "eval.cuh"
#include "raster.cuh"
#include "table.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace tpose_dev;

struct ImageSlot {
    uchar4* rgba = nullptr;    // raster-resolution image on the device
    uint4* table = nullptr;    // row-prefix moments, (W+1) entries per row
    int W = 0, H = 0, pitch = 0;
    size_t rgba_cap = 0, table_cap = 0;
};

struct tpose_cuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int maxT = 0, maxP = 0;
    int NT = 0, NP = 0;
    float ratio = 1.5f;
    bool have_mesh = false;
    ImageSlot img[2];
    uchar4* staging = nullptr; size_t staging_cap = 0;   // source-resolution upload buffer
    float2* pts = nullptr;
    int4* tri = nullptr;
    int* cn = nullptr; int4* ca = nullptr; int* ten = nullptr;
    int2* gr = nullptr; int2* gr_acc = nullptr;
    int* hist = nullptr;
    unsigned* ticket = nullptr;
    float* err_state = nullptr;       // [0] toterr [1] |relerr| [2] newerr
    int* run_state = nullptr;         // [0] stop flag [1] steps done
    int* ids = nullptr; size_t ids_cap = 0;
    unsigned char* flush = nullptr; size_t flush_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.0f;
    int64_t launches = 0;
    int group = 0;                    // lanes per edge (0 = pick from NT and H)
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

extern "C" const char* tpose_cuda_version(void)
{
    return "tpose-cuda 0.1 (sm_100a; edge-integral evaluator over row-prefix moment tables)";
}

extern "C" const char* tpose_cuda_last_error(const tpose_cuda_ctx* ctx)
{
    return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

extern "C" int tpose_cuda_create(int device, int max_triangles, int max_points, void* stream, tpose_cuda_ctx** out)
{
    if (!out || max_triangles <= 0 || max_points < 3 || 13ll * max_triangles > TPOSE_REF_MAXT) {
        g_create_error = "tpose_cuda_create: bad arguments (need 13*max_triangles <= 2<<18)";
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
    ALLOC(ctx->pts, sizeof(float2) * max_points);
    ALLOC(ctx->tri, sizeof(int4) * max_triangles);
    ALLOC(ctx->cn, sizeof(int) * nv);
    ALLOC(ctx->ca, sizeof(int4) * nv);
    ALLOC(ctx->ten, sizeof(int) * nv);
    ALLOC(ctx->gr, sizeof(int2) * max_points);
    ALLOC(ctx->gr_acc, sizeof(int2) * max_points);
    ALLOC(ctx->hist, sizeof(int) * max_triangles);
    ALLOC(ctx->ticket, sizeof(unsigned));
    ALLOC(ctx->err_state, sizeof(float) * 4);
    ALLOC(ctx->run_state, sizeof(int) * 4);
#undef ALLOC
    cudaMemsetAsync(ctx->cn, 0, sizeof(int) * nv, ctx->stream);
    cudaMemsetAsync(ctx->ca, 0, sizeof(int4) * nv, ctx->stream);
    cudaMemsetAsync(ctx->ten, 0, sizeof(int) * nv, ctx->stream);
    cudaMemsetAsync(ctx->gr, 0, sizeof(int2) * max_points, ctx->stream);
    cudaMemsetAsync(ctx->gr_acc, 0, sizeof(int2) * max_points, ctx->stream);
    cudaMemsetAsync(ctx->ticket, 0, sizeof(unsigned), ctx->stream);
    cudaMemsetAsync(ctx->run_state, 0, sizeof(int) * 4, ctx->stream);
    const float st[4] = {1.0f, 0.0f, 0.0f, 0.0f};   // toterr starts at 1 (triangulation.hpp:648)
    cudaMemcpyAsync(ctx->err_state, st, sizeof st, cudaMemcpyHostToDevice, ctx->stream);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail("init", e);
    const char* g = getenv("TPOSE_CUDA_GROUP");
    if (g) ctx->group = atoi(g);
    *out = ctx;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_destroy(tpose_cuda_ctx* ctx)
{
    if (!ctx) return TPOSE_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->img) { cudaFree(s.rgba); cudaFree(s.table); }
    cudaFree(ctx->staging); cudaFree(ctx->pts); cudaFree(ctx->tri); cudaFree(ctx->cn); cudaFree(ctx->ca);
    cudaFree(ctx->ten); cudaFree(ctx->gr); cudaFree(ctx->gr_acc); cudaFree(ctx->hist); cudaFree(ctx->ticket);
    cudaFree(ctx->err_state); cudaFree(ctx->run_state); cudaFree(ctx->ids); cudaFree(ctx->flush);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TPOSE_OK;
}

// ---- images ---------------------------------------------------------------------------------

static int build_table(tpose_cuda_ctx* ctx, ImageSlot& s)
{
    const bool bulk = (s.W % 4 == 0);
    const size_t smem = bulk ? (size_t)s.W * 4 : 0;
    if (bulk) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_build_table<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_build_table<true><<<s.H, kTableThreads, smem, ctx->stream>>>(s.rgba, s.W, s.H, s.table, s.pitch);
    } else {
        k_build_table<false><<<s.H, kTableThreads, 0, ctx->stream>>>(s.rgba, s.W, s.H, s.table, s.pitch);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TPOSE_OK;
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
    const size_t need_rgba = (size_t)rw * rh * sizeof(uchar4);
    const size_t need_table = (size_t)(rw + 1) * rh * sizeof(uint4);
    if (s.rgba_cap < need_rgba) { cudaFree(s.rgba); s.rgba = nullptr; CK(cudaMalloc((void**)&s.rgba, need_rgba)); s.rgba_cap = need_rgba; }
    if (s.table_cap < need_table) { cudaFree(s.table); s.table = nullptr; CK(cudaMalloc((void**)&s.table, need_table)); s.table_cap = need_table; }
    s.W = rw; s.H = rh; s.pitch = rw + 1;
    const cudaMemcpyKind kind = src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (rw == w && rh == h) {
        CK(cudaMemcpyAsync(s.rgba, src, need_rgba, kind, ctx->stream));
    } else {
        const size_t need_src = (size_t)w * h * sizeof(uchar4);
        if (ctx->staging_cap < need_src) { cudaFree(ctx->staging); ctx->staging = nullptr; CK(cudaMalloc((void**)&ctx->staging, need_src)); ctx->staging_cap = need_src; }
        CK(cudaMemcpyAsync(ctx->staging, src, need_src, kind, ctx->stream));
        dim3 grid((rw + 127) / 128, rh);
        k_resample<<<grid, 128, 0, ctx->stream>>>(ctx->staging, w, h, s.rgba, rw, rh);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return build_table(ctx, s);
}

extern "C" int tpose_cuda_set_image(tpose_cuda_ctx* ctx, int slot, const uint8_t* rgba_host, int w, int h, int rw, int rh)
{
    return set_image_common(ctx, slot, rgba_host, true, w, h, rw, rh);
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
    CK(cudaMemcpyAsync(ctx->pts, pts_xy, sizeof(float2) * np, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->tri, tri_xyzw, sizeof(int4) * nt, cudaMemcpyHostToDevice, ctx->stream));
    if (colors_rgba) {   // triangulation.hpp:633-641: every variant row gets the triangle's colour
        for (int k = 0; k < kVariants; k++)
            CK(cudaMemcpyAsync(ctx->ca + (size_t)k * nt, colors_rgba, sizeof(int4) * nt, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->NT = nt; ctx->NP = np; ctx->ratio = ratio; ctx->have_mesh = true;
    CK(cudaStreamSynchronize(ctx->stream));   // caller may reuse its host buffers
    return TPOSE_OK;
}

extern "C" int tpose_cuda_upload_points(tpose_cuda_ctx* ctx, const float* pts_xy, int np)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "upload_points: no mesh uploaded");
    if (!pts_xy || np != ctx->NP) return fail(ctx, TPOSE_ERR_ARG, "upload_points: NP differs from the uploaded mesh");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->pts, pts_xy, sizeof(float2) * np, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

// ---- evaluation -----------------------------------------------------------------------------

static int pick_group(const tpose_cuda_ctx* ctx, int H)
{
    if (ctx->group == 1 || ctx->group == 2 || ctx->group == 4 || ctx->group == 8 || ctx->group == 16 || ctx->group == 32)
        return ctx->group;
    // mean vertical extent of an edge is about H*sqrt(2/NT)*0.6; aim at <= ~8 rows per lane
    const float rows = 0.6f * (float)H * sqrtf(2.0f / (float)ctx->NT);
    int g = 1;
    while (g < 32 && rows / g > 8.0f) g <<= 1;
    return g;
}

template <int G, int TPB>
static void launch_eval(const EvalParams& p, cudaStream_t st)
{
    constexpr int threads = ((TPB * kEdgeVariants * G + 31) / 32) * 32;
    const int blocks = (p.NT + TPB - 1) / TPB;
    k_eval<G, TPB><<<blocks, threads, 0, st>>>(p);
}

static int eval(tpose_cuda_ctx* ctx, int slot, int flavour, unsigned flags, bool with_stop)
{
    const ImageSlot& s = ctx->img[slot];
    EvalParams p;
    p.table = s.table; p.pitch = s.pitch; p.W = s.W; p.H = s.H;
    p.pts = ctx->pts; p.tri = ctx->tri; p.NT = ctx->NT; p.NP = ctx->NP;
    p.ratio = ctx->ratio;
    p.dp = dp_for(ctx->NT, flavour);
    p.step = flavour == TPOSE_WARP ? 0.00003f : 0.00005f;
    p.cn = ctx->cn; p.ca = ctx->ca; p.ten = ctx->ten; p.gr = ctx->gr; p.gr_acc = ctx->gr_acc;
    p.stop = with_stop ? ctx->run_state : nullptr;
    p.hist = ctx->hist; p.ticket = ctx->ticket; p.flags = flags;
    switch (pick_group(ctx, s.H)) {
    case 1: launch_eval<1, 8>(p, ctx->stream); break;
    case 2: launch_eval<2, 4>(p, ctx->stream); break;
    case 4: launch_eval<4, 2>(p, ctx->stream); break;
    case 8: launch_eval<8, 1>(p, ctx->stream); break;
    case 16: launch_eval<16, 1>(p, ctx->stream); break;
    default: launch_eval<32, 1>(p, ctx->stream); break;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TPOSE_OK;
}

static int check_ready(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (slot < 0 || slot > 1 || (flavour != TPOSE_TRIANGULATE && flavour != TPOSE_WARP)) return fail(ctx, TPOSE_ERR_ARG, "bad slot or flavour");
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "no mesh uploaded");
    if (!ctx->img[slot].table) return fail(ctx, TPOSE_ERR_STATE, "no image in this slot");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(ctx, TPOSE_ERR_CUDA, cudaGetErrorString(e));
    return TPOSE_OK;
}

static unsigned fused_flags(int flavour)
{
    return flavour == TPOSE_WARP
        ? (EV_WRITE_CN | EV_WRITE_TEN | EV_MEAN_STORED | EV_GRAD_SHIFT | EV_HISTORY)
        : (EV_WRITE_CN | EV_WRITE_CA | EV_WRITE_TEN | EV_MEAN_DIVIDE | EV_GRAD_SHIFT | EV_HISTORY);
}

extern "C" int tpose_cuda_colors(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    return eval(ctx, slot, flavour, flavour == TPOSE_WARP ? EV_WRITE_CN : (EV_WRITE_CN | EV_WRITE_CA), false);
}

extern "C" int tpose_cuda_energy(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    k_zero_gradient<<<(ctx->NP + 255) / 256, 256, 0, ctx->stream>>>(ctx->gr, ctx->NP);   // triangle.vs:95-97
    ctx->launches++;
    const unsigned fl = EV_WRITE_TEN | EV_MEAN_STORED | EV_HISTORY | (flavour == TPOSE_WARP ? 0u : EV_MEAN_DIVIDE);
    return eval(ctx, slot, flavour, fl, false);
}

extern "C" int tpose_cuda_shift(tpose_cuda_ctx* ctx, int flavour)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    k_gradient<<<(ctx->NT + 127) / 128, 128, 0, ctx->stream>>>(ctx->tri, ctx->NT, ctx->ten, ctx->gr);
    k_shift<<<(ctx->NP + 127) / 128, 128, 0, ctx->stream>>>(ctx->pts, ctx->NP, ctx->gr, ctx->ratio,
                                                            flavour == TPOSE_WARP ? 0.00003f : 0.00005f);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return TPOSE_OK;
}

extern "C" int tpose_cuda_step(tpose_cuda_ctx* ctx, int slot, int flavour, int nsteps)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if (nsteps < 0) return fail(ctx, TPOSE_ERR_ARG, "step: nsteps < 0");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const unsigned fl = fused_flags(flavour);
    for (int i = 0; i < nsteps; i++) {
        rc = eval(ctx, slot, flavour, fl, false);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_run(tpose_cuda_ctx* ctx, int slot, int flavour, int max_steps, float rel_tol,
                              int* steps_done, uint32_t* trigger_flags, float* last_err)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if (max_steps < 0) return fail(ctx, TPOSE_ERR_ARG, "run: max_steps < 0");
    CK(cudaMemsetAsync(ctx->run_state, 0, sizeof(int) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const unsigned fl = fused_flags(flavour);
    for (int i = 0; i < max_steps; i++) {
        rc = eval(ctx, slot, flavour, fl, true);
        if (rc) return rc;
        k_geterr<<<1, 32, 0, ctx->stream>>>(ctx->hist, ctx->NT, ctx->err_state, rel_tol, ctx->run_state, ctx->run_state + 1);
        ctx->launches++;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    int rs[4]; float es[4];
    CK(cudaMemcpyAsync(rs, ctx->run_state, sizeof rs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(es, ctx->err_state, sizeof es, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (steps_done) *steps_done = rs[1];
    if (trigger_flags) *trigger_flags = rs[0] ? TPOSE_TRIG_CONVERGED : TPOSE_TRIG_MAXSTEPS;
    if (last_err) *last_err = es[1];
    return TPOSE_OK;
}

extern "C" int tpose_cuda_set_toterr(tpose_cuda_ctx* ctx, float toterr)
{
    if (!ctx) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->err_state, &toterr, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

extern "C" int tpose_cuda_get_toterr(tpose_cuda_ctx* ctx, float* toterr)
{
    if (!ctx || !toterr) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(toterr, ctx->err_state, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TPOSE_OK;
}

// ---- read back ------------------------------------------------------------------------------

extern "C" int tpose_cuda_retrieve(tpose_cuda_ctx* ctx, int32_t* ten13, int32_t* cn13, int32_t* ca13, float* pts_xy)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "retrieve: no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    const size_t nv = (size_t)kVariants * ctx->NT;
    if (ten13) CK(cudaMemcpyAsync(ten13, ctx->ten, sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn13) CK(cudaMemcpyAsync(cn13, ctx->cn, sizeof(int) * nv, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca13) CK(cudaMemcpyAsync(ca13, ctx->ca, sizeof(int4) * nv, cudaMemcpyDeviceToHost, ctx->stream));
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
    if (ten) CK(cudaMemcpyAsync(ten, ctx->ten, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (cn) CK(cudaMemcpyAsync(cn, ctx->cn, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ca) CK(cudaMemcpyAsync(ca, ctx->ca, sizeof(int4) * n, cudaMemcpyDeviceToHost, ctx->stream));
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

extern "C" int tpose_cuda_last_elapsed_ms(tpose_cuda_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return TPOSE_OK;
}

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

extern "C" int tpose_cuda_flush_l2(tpose_cuda_ctx* ctx, size_t bytes)
{
    if (!ctx || bytes == 0) return TPOSE_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->flush_cap < bytes) { cudaFree(ctx->flush); ctx->flush = nullptr; CK(cudaMalloc((void**)&ctx->flush, bytes)); ctx->flush_cap = bytes; }
    CK(cudaMemsetAsync(ctx->flush, 0x5a, bytes, ctx->stream));
    return TPOSE_OK;
}

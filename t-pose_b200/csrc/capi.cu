// capi.cu -- the C ABI of include/tpose_cuda.h over the sm_100a kernels.
// There is no CPU path in this library: without a CUDA device every entry point fails.
#include "../../include/tpose_cuda.h"
#include "eval.cuh"
#include "mview.cuh"
#include "raster.cuh"
#include "table.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

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
    int* cn[2] = {nullptr, nullptr}; int4* ca[2] = {nullptr, nullptr}; int* ten[2] = {nullptr, nullptr};
    int cur = 0;                      // buffer set holding the last committed evaluation
    int2* gr = nullptr; int2* gr_acc = nullptr;
    int4* wtri = nullptr; int* whe = nullptr; bool watching = false;
    unsigned* ticket = nullptr;       // [0] ticket [1] topology trigger accumulator
    float* err_state = nullptr;       // [0] toterr [1] |relerr| [2] newerr [3] converged flag
    int* run_state = nullptr;         // RS_STOP, RS_STEPS, RS_TRIG
    EvalJob* job_dev = nullptr;       // this context's job record (single launches)
    EvalJob job_host{};               // what job_dev holds
    bool job_valid = false;
    EvalJob* batch_jobs = nullptr; int batch_cap = 0;   // job array when this context leads a batch
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
    for (int s = 0; s < 2; s++) {
        ALLOC(ctx->cn[s], sizeof(int) * nv);
        ALLOC(ctx->ca[s], sizeof(int4) * nv);
        ALLOC(ctx->ten[s], sizeof(int) * nv);
    }
    ALLOC(ctx->gr, sizeof(int2) * max_points);
    ALLOC(ctx->gr_acc, sizeof(int2) * max_points);
    ALLOC(ctx->wtri, sizeof(int4) * max_triangles);
    ALLOC(ctx->whe, sizeof(int) * 3 * max_triangles);
    ALLOC(ctx->ticket, sizeof(unsigned) * 2);
    ALLOC(ctx->err_state, sizeof(float) * 4);
    ALLOC(ctx->run_state, sizeof(int) * 4);
    ALLOC(ctx->job_dev, sizeof(EvalJob));
#undef ALLOC
    for (int s = 0; s < 2; s++) {
        cudaMemsetAsync(ctx->cn[s], 0, sizeof(int) * nv, ctx->stream);
        cudaMemsetAsync(ctx->ca[s], 0, sizeof(int4) * nv, ctx->stream);
        cudaMemsetAsync(ctx->ten[s], 0, sizeof(int) * nv, ctx->stream);
    }
    cudaMemsetAsync(ctx->gr, 0, sizeof(int2) * max_points, ctx->stream);
    cudaMemsetAsync(ctx->gr_acc, 0, sizeof(int2) * max_points, ctx->stream);
    cudaMemsetAsync(ctx->ticket, 0, sizeof(unsigned) * 2, ctx->stream);
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
    for (int s = 0; s < 2; s++) { cudaFree(ctx->cn[s]); cudaFree(ctx->ca[s]); cudaFree(ctx->ten[s]); }
    cudaFree(ctx->staging); cudaFree(ctx->pts); cudaFree(ctx->tri);
    cudaFree(ctx->gr); cudaFree(ctx->gr_acc); cudaFree(ctx->wtri); cudaFree(ctx->whe); cudaFree(ctx->ticket);
    cudaFree(ctx->err_state); cudaFree(ctx->run_state); cudaFree(ctx->ids); cudaFree(ctx->flush);
    cudaFree(ctx->job_dev); cudaFree(ctx->batch_jobs);
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
        for (int s = 0; s < 2; s++)
            for (int k = 0; k < kVariants; k++)
                CK(cudaMemcpyAsync(ctx->ca[s] + (size_t)k * nt, colors_rgba, sizeof(int4) * nt, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (nt != ctx->NT) ctx->watching = false;   // a watch topology is only valid for the same NT
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

static int pick_group(const tpose_cuda_ctx* ctx, int NT, int H)
{
    if (ctx->group == 1 || ctx->group == 2 || ctx->group == 4 || ctx->group == 8 || ctx->group == 16 || ctx->group == 32)
        return ctx->group;
    // mean vertical extent of an edge is about H*sqrt(2/NT)*0.6; aim at <= ~8 rows per lane
    const float rows = 0.6f * (float)H * sqrtf(2.0f / (float)NT);
    int g = 1;
    while (g < 32 && rows / g > 8.0f) g <<= 1;
    return g;
}

template <int G, int TPB>
static void launch_eval(const EvalJob* jobs, int njobs, int max_nt, int set, unsigned flags, double tol, cudaStream_t st)
{
    const dim3 grid((max_nt + TPB - 1) / TPB, njobs);
    k_eval<G, TPB><<<grid, EvalCfg<G, TPB>::kThreads, 0, st>>>(jobs, set, flags, tol);
}

static void launch_eval_group(int group, const EvalJob* jobs, int njobs, int max_nt, int set, unsigned flags, double tol, cudaStream_t st)
{
    switch (group) {
    case 1: launch_eval<1, 8>(jobs, njobs, max_nt, set, flags, tol, st); break;
    case 2: launch_eval<2, 4>(jobs, njobs, max_nt, set, flags, tol, st); break;
    case 4: launch_eval<4, 2>(jobs, njobs, max_nt, set, flags, tol, st); break;
    case 8: launch_eval<8, 1>(jobs, njobs, max_nt, set, flags, tol, st); break;
    case 16: launch_eval<16, 1>(jobs, njobs, max_nt, set, flags, tol, st); break;
    default: launch_eval<32, 1>(jobs, njobs, max_nt, set, flags, tol, st); break;
    }
}

// the job record of this context for (slot, flavour)
static EvalJob make_job(const tpose_cuda_ctx* ctx, int slot, int flavour)
{
    const ImageSlot& s = ctx->img[slot];
    EvalJob j;
    memset(&j, 0, sizeof j);
    j.table = s.table; j.pitch = s.pitch; j.W = s.W; j.H = s.H;
    j.pts = ctx->pts; j.tri = ctx->tri; j.NT = ctx->NT; j.NP = ctx->NP;
    j.ratio = ctx->ratio;
    j.dp = dp_for(ctx->NT, flavour);
    j.step = flavour == TPOSE_WARP ? 0.00003f : 0.00005f;
    for (int k = 0; k < 2; k++) { j.cn[k] = ctx->cn[k]; j.ca[k] = ctx->ca[k]; j.ten[k] = ctx->ten[k]; }
    j.gr = ctx->gr; j.gr_acc = ctx->gr_acc;
    j.wtri = ctx->watching ? ctx->wtri : nullptr;
    j.whe = ctx->watching ? ctx->whe : nullptr;
    j.err_state = ctx->err_state; j.run_state = ctx->run_state;
    j.trig = ctx->ticket + 1; j.ticket = ctx->ticket;
    return j;
}

// make job_dev current (a 200-byte copy, only when something changed)
static int sync_job(tpose_cuda_ctx* ctx, int slot, int flavour)
{
    const EvalJob j = make_job(ctx, slot, flavour);
    if (!ctx->job_valid || memcmp(&j, &ctx->job_host, sizeof j) != 0) {
        ctx->job_host = j;
        CK(cudaMemcpyAsync(ctx->job_dev, &ctx->job_host, sizeof j, cudaMemcpyHostToDevice, ctx->stream));
        ctx->job_valid = true;
    }
    return TPOSE_OK;
}

static int eval(tpose_cuda_ctx* ctx, int slot, int flavour, int set, unsigned flags, double tol = 0.0)
{
    int rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    launch_eval_group(pick_group(ctx, ctx->NT, ctx->img[slot].H), ctx->job_dev, 1, ctx->NT, set, flags, tol, ctx->stream);
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
    k_zero_gradient<<<(ctx->NP + 255) / 256, 256, 0, ctx->stream>>>(ctx->gr, ctx->NP);   // triangle.vs:95-97
    ctx->launches++;
    const unsigned fl = EV_WRITE_TEN | EV_MEAN_STORED | (flavour == TPOSE_WARP ? 0u : EV_MEAN_DIVIDE);
    return eval(ctx, slot, flavour, ctx->cur, fl);
}

extern "C" int tpose_cuda_shift(tpose_cuda_ctx* ctx, int flavour)
{
    if (!ctx) return TPOSE_ERR_ARG;
    if (!ctx->have_mesh) return fail(ctx, TPOSE_ERR_STATE, "no mesh uploaded");
    CK(cudaSetDevice(ctx->device));
    k_gradient<<<(ctx->NT + 127) / 128, 128, 0, ctx->stream>>>(ctx->tri, ctx->NT, ctx->ten[ctx->cur], ctx->gr);
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
    rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const unsigned fl = fused_flags(flavour);
    for (int i = 0; i < nsteps; i++) {
        ctx->cur ^= 1;
        rc = eval(ctx, slot, flavour, ctx->cur, fl);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    return TPOSE_OK;
}

// Launch j evaluates the points after j shifts into set (first+j)&1 and carries the checks of frame
// j-1; launch max_steps is check-only.  See eval.cuh.
static void issue_run(tpose_cuda_ctx* lead, const EvalJob* jobs, int njobs, int max_nt, int group, int first_set,
                      int flavour, int max_steps, double rel_tol)
{
    const unsigned fl = fused_flags(flavour) | EV_HONOUR_STOP;
    for (int j = 0; j <= max_steps; j++) {
        unsigned f = fl;
        if (j > 0) f |= EV_CHECK_PREV;
        if (j == max_steps) f |= EV_FINAL;
        launch_eval_group(group, jobs, njobs, max_nt, (first_set + j) & 1, f, rel_tol, lead->stream);
        lead->launches++;
    }
}

extern "C" int tpose_cuda_run(tpose_cuda_ctx* ctx, int slot, int flavour, int max_steps, double rel_tol,
                              int* steps_done, uint32_t* trigger_flags, float* last_err)
{
    int rc = check_ready(ctx, slot, flavour);
    if (rc) return rc;
    if (max_steps < 1) return fail(ctx, TPOSE_ERR_ARG, "run: max_steps < 1");
    rc = sync_job(ctx, slot, flavour);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->run_state, 0, sizeof(int) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int first = ctx->cur ^ 1;
    issue_run(ctx, ctx->job_dev, 1, ctx->NT, pick_group(ctx, ctx->NT, ctx->img[slot].H), first, flavour, max_steps, rel_tol);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    int rs[4]; float es[4];
    CK(cudaMemcpyAsync(rs, ctx->run_state, sizeof rs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(es, ctx->err_state, sizeof es, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (rs[RS_STEPS] > 0) ctx->cur = (first + rs[RS_STEPS] - 1) & 1;
    if (steps_done) *steps_done = rs[RS_STEPS];
    if (trigger_flags) *trigger_flags = (uint32_t)rs[RS_TRIG];
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
    CK(cudaMemcpyAsync(ctx->wtri, tri_xyzw, sizeof(int4) * nt, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->whe, halfedges, sizeof(int) * 3 * nt, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->watching = true;
    return TPOSE_OK;
}

// ---- batches ---------------------------------------------------------------------------------

static int batch_prepare(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, int* max_nt, int* group)
{
    if (!ctxs || n < 1 || !ctxs[0]) return TPOSE_ERR_ARG;
    tpose_cuda_ctx* lead = ctxs[0];
    tpose_cuda_ctx* ctx = lead;   // for CK
    std::vector<EvalJob> jobs((size_t)n);
    int mx = 0; double rows = 0.0;
    for (int i = 0; i < n; i++) {
        const int slot = slots ? slots[i] : 0;
        int rc = check_ready(ctxs[i], slot, flavour);
        if (rc) { if (ctxs[i] && ctxs[i] != lead) lead->err = "batch: job " + std::to_string(i) + ": " + ctxs[i]->err; return rc; }
        if (ctxs[i]->device != lead->device) return fail(lead, TPOSE_ERR_ARG, "batch: contexts live on different devices");
        for (int k = 0; k < i; k++) if (ctxs[k] == ctxs[i]) return fail(lead, TPOSE_ERR_ARG, "batch: a context appears twice");
        if (ctxs[i] != lead) cudaStreamSynchronize(ctxs[i]->stream);   // its uploads are done before we touch its buffers
        jobs[(size_t)i] = make_job(ctxs[i], slot, flavour);
        mx = ctxs[i]->NT > mx ? ctxs[i]->NT : mx;
        rows += 0.6 * ctxs[i]->img[slot].H * sqrt(2.0 / ctxs[i]->NT);
    }
    if (lead->batch_cap < n) {
        cudaFree(lead->batch_jobs); lead->batch_jobs = nullptr; lead->batch_cap = 0;
        CK(cudaMalloc((void**)&lead->batch_jobs, sizeof(EvalJob) * n));
        lead->batch_cap = n;
    }
    CK(cudaMemcpyAsync(lead->batch_jobs, jobs.data(), sizeof(EvalJob) * n, cudaMemcpyHostToDevice, lead->stream));
    CK(cudaStreamSynchronize(lead->stream));   // `jobs` is a local
    *max_nt = mx;
    int g = lead->group;
    if (!(g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32)) {
        g = 1;
        while (g < 32 && rows / n / g > 8.0) g <<= 1;
    }
    *group = g;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_batch_step(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, int nsteps)
{
    int max_nt = 0, group = 1;
    int rc = batch_prepare(ctxs, n, slots, flavour, &max_nt, &group);
    if (rc) return rc;
    tpose_cuda_ctx* ctx = ctxs[0];
    if (nsteps < 0) return fail(ctx, TPOSE_ERR_ARG, "batch_step: nsteps < 0");
    // all jobs must write the same set parity: bring every context to the leader's
    for (int i = 1; i < n; i++)
        if (ctxs[i]->cur != ctx->cur) return fail(ctx, TPOSE_ERR_STATE, "batch: contexts are out of step (mix of stepped and fresh contexts)");
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const unsigned fl = fused_flags(flavour);
    int cur = ctx->cur;
    for (int s = 0; s < nsteps; s++) {
        cur ^= 1;
        launch_eval_group(group, ctx->batch_jobs, n, max_nt, cur, fl, 0.0, ctx->stream);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    for (int i = 0; i < n; i++) ctxs[i]->cur = cur;
    return TPOSE_OK;
}

extern "C" int tpose_cuda_batch_run(tpose_cuda_ctx* const* ctxs, int n, const int* slots, int flavour, int max_steps,
                                    double rel_tol, int* steps_done, uint32_t* trigger_flags)
{
    int max_nt = 0, group = 1;
    int rc = batch_prepare(ctxs, n, slots, flavour, &max_nt, &group);
    if (rc) return rc;
    tpose_cuda_ctx* ctx = ctxs[0];
    if (max_steps < 1) return fail(ctx, TPOSE_ERR_ARG, "batch_run: max_steps < 1");
    for (int i = 1; i < n; i++)
        if (ctxs[i]->cur != ctx->cur) return fail(ctx, TPOSE_ERR_STATE, "batch: contexts are out of step");
    for (int i = 0; i < n; i++) CK(cudaMemsetAsync(ctxs[i]->run_state, 0, sizeof(int) * 4, ctx->stream));
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const int first = ctx->cur ^ 1;
    issue_run(ctx, ctx->batch_jobs, n, max_nt, group, first, flavour, max_steps, rel_tol);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    std::vector<int> rs((size_t)4 * n);
    for (int i = 0; i < n; i++)
        CK(cudaMemcpyAsync(&rs[(size_t)4 * i], ctxs[i]->run_state, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) {
        const int done = rs[(size_t)4 * i + RS_STEPS];
        if (done > 0) ctxs[i]->cur = (first + done - 1) & 1;
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
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync((char*)dev + elem * first, in, elem * count, cudaMemcpyHostToDevice, ctx->stream));
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

// ---- multiview (source/multiview.hpp) -----------------------------------------------------------

namespace {
std::string g_mview_error;
float g_mview_ms = 0.0f;

struct DevBuf {                       // RAII device allocation for one call
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T* as() { return (T*)p; }
};

int mview_fail(int code, const std::string& msg) { g_mview_error = msg; return code; }

#define MCK(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return mview_fail(TPOSE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int mview_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        (void)cudaGetLastError();
        return mview_fail(TPOSE_ERR_NODEVICE, "no usable CUDA device (this library has no CPU fallback)");
    }
    if (cudaSetDevice(device) != cudaSuccess) return mview_fail(TPOSE_ERR_CUDA, "cudaSetDevice failed");
    return TPOSE_OK;
}

bool offsets_ok(const int* offsets, int nbatch)
{
    if (!offsets || nbatch < 1 || offsets[0] != 0) return false;
    for (int i = 0; i < nbatch; i++) if (offsets[i + 1] < offsets[i]) return false;
    return true;
}
}  // namespace

extern "C" const char* tpose_cuda_mview_last_error(void) { return g_mview_error.c_str(); }
extern "C" float tpose_cuda_mview_last_ms(void) { return g_mview_ms; }

extern "C" int tpose_cuda_f8point(int device, const float* A_xy, const float* B_xy, const int* offsets, int nbatch, float* F9_out)
{
    if (!A_xy || !B_xy || !F9_out || !offsets_ok(offsets, nbatch)) return mview_fail(TPOSE_ERR_ARG, "f8point: bad arguments");
    int rc = mview_device(device);
    if (rc) return rc;
    const size_t n = (size_t)offsets[nbatch];
    DevBuf dA, dB, dO, dF;
    MCK(dA.alloc(n * sizeof(float2))); MCK(dB.alloc(n * sizeof(float2)));
    MCK(dO.alloc((nbatch + 1) * sizeof(int))); MCK(dF.alloc((size_t)nbatch * 9 * sizeof(float)));
    MCK(cudaMemcpy(dA.p, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dB.p, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dO.p, offsets, (nbatch + 1) * sizeof(int), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    MCK(cudaEventCreate(&e0)); MCK(cudaEventCreate(&e1));
    MCK(cudaEventRecord(e0, 0));
    k_f8point<<<nbatch, kF8Threads>>>(dA.as<float2>(), dB.as<float2>(), dO.as<int>(), dF.as<float>());
    MCK(cudaEventRecord(e1, 0));
    MCK(cudaGetLastError());
    MCK(cudaMemcpy(F9_out, dF.p, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaEventElapsedTime(&g_mview_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return TPOSE_OK;
}

extern "C" int tpose_cuda_triangulate(int device, const float* F9, const float* K9, const float* A_xy, const float* B_xy,
                                      const int* offsets, int nbatch, int optimal_correction, float* X4_out,
                                      float* A_corr, float* B_corr)
{
    if (!F9 || !K9 || !A_xy || !B_xy || !X4_out || !offsets_ok(offsets, nbatch)) return mview_fail(TPOSE_ERR_ARG, "triangulate: bad arguments");
    int rc = mview_device(device);
    if (rc) return rc;
    const size_t n = (size_t)offsets[nbatch];
    int longest = 0;
    for (int i = 0; i < nbatch; i++) longest = std::max(longest, offsets[i + 1] - offsets[i]);
    DevBuf dA, dB, dO, dF, dK, dX, dAc, dBc;
    MCK(dA.alloc(n * sizeof(float2))); MCK(dB.alloc(n * sizeof(float2)));
    MCK(dO.alloc((nbatch + 1) * sizeof(int)));
    MCK(dF.alloc((size_t)nbatch * 9 * sizeof(float))); MCK(dK.alloc((size_t)nbatch * 9 * sizeof(float)));
    MCK(dX.alloc(n * sizeof(float4)));
    if (A_corr) MCK(dAc.alloc(n * sizeof(float2)));
    if (B_corr) MCK(dBc.alloc(n * sizeof(float2)));
    MCK(cudaMemcpy(dA.p, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dB.p, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dO.p, offsets, (nbatch + 1) * sizeof(int), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dF.p, F9, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dK.p, K9, (size_t)nbatch * 9 * sizeof(float), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    MCK(cudaEventCreate(&e0)); MCK(cudaEventCreate(&e1));
    MCK(cudaEventRecord(e0, 0));
    if (longest > 0) {
        const dim3 grid((longest + 127) / 128, nbatch);
        k_triangulate<<<grid, 128>>>(dF.as<float>(), dK.as<float>(), dA.as<float2>(), dB.as<float2>(), dO.as<int>(),
                                     optimal_correction, dX.as<float4>(), A_corr ? dAc.as<float2>() : nullptr,
                                     B_corr ? dBc.as<float2>() : nullptr);
    }
    MCK(cudaEventRecord(e1, 0));
    MCK(cudaGetLastError());
    MCK(cudaMemcpy(X4_out, dX.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    if (A_corr) MCK(cudaMemcpy(A_corr, dAc.p, n * sizeof(float2), cudaMemcpyDeviceToHost));
    if (B_corr) MCK(cudaMemcpy(B_corr, dBc.p, n * sizeof(float2), cudaMemcpyDeviceToHost));
    cudaEventElapsedTime(&g_mview_ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return TPOSE_OK;
}

extern "C" int tpose_cuda_hdlt(int device, const float* PA12, const float* PB12, const float* A_xy, const float* B_xy,
                               int n, float* X4_out)
{
    if (!PA12 || !PB12 || !A_xy || !B_xy || !X4_out || n < 1) return mview_fail(TPOSE_ERR_ARG, "hdlt: bad arguments");
    int rc = mview_device(device);
    if (rc) return rc;
    DevBuf dA, dB, dPA, dPB, dX;
    MCK(dA.alloc(n * sizeof(float2))); MCK(dB.alloc(n * sizeof(float2)));
    MCK(dPA.alloc(12 * sizeof(float))); MCK(dPB.alloc(12 * sizeof(float))); MCK(dX.alloc(n * sizeof(float4)));
    MCK(cudaMemcpy(dA.p, A_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dB.p, B_xy, n * sizeof(float2), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dPA.p, PA12, 12 * sizeof(float), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dPB.p, PB12, 12 * sizeof(float), cudaMemcpyHostToDevice));
    k_hdlt<<<(n + 127) / 128, 128>>>(dPA.as<float>(), dPB.as<float>(), dA.as<float2>(), dB.as<float2>(), n, dX.as<float4>());
    MCK(cudaGetLastError());
    MCK(cudaMemcpy(X4_out, dX.p, n * sizeof(float4), cudaMemcpyDeviceToHost));
    return TPOSE_OK;
}

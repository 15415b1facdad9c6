// table.cuh -- builds the packed row-prefix moment table of one image: one streaming pass that
// reads every RGBA8 texel once and writes one 16-byte entry per (row, column boundary).
// This is the only kernel that touches the image itself; it is HBM-bound (4 B in, 16 B out per
// pixel).  The row is staged into shared memory with one TMA bulk copy (cp.async.bulk, UBLKCP in
// SASS) signalled through an mbarrier, scanned 256 pixels at a time, and stored with coalesced
// 16-byte vector stores.
//
// The texel arithmetic it tabulates is triangle.fs:29-33 / :40-42 of software/triangulate/shader
// (int(255*c) == the stored byte on the native raster), summed ahead of time.
#pragma once
#include "common.cuh"

namespace tpose_dev {

constexpr int kTableThreads = 256;

struct Part { unsigned r, g, b, s2, odd; };

__device__ __forceinline__ Part part_add(Part a, Part b)
{
    return Part{a.r + b.r, a.g + b.g, a.b + b.b, a.s2 + b.s2, a.odd + b.odd};
}
__device__ __forceinline__ Part part_shfl_up(Part a, int d)
{
    return Part{__shfl_up_sync(0xffffffffu, a.r, d), __shfl_up_sync(0xffffffffu, a.g, d),
                __shfl_up_sync(0xffffffffu, a.b, d), __shfl_up_sync(0xffffffffu, a.s2, d),
                __shfl_up_sync(0xffffffffu, a.odd, d)};
}
__device__ __forceinline__ Part part_of(uchar4 px)
{
    const unsigned r = px.x, g = px.y, b = px.z;
    return Part{r, g, b, r * r + g * g + b * b, (r + g + b) & 1u};
}
__device__ __forceinline__ uint4 part_pack(Part a)
{
    return make_uint4(a.r | (a.odd << 20), a.g, a.b, a.s2);
}

__device__ __forceinline__ unsigned smem_u32(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}

// One block per image row.  BULK: rows are 16-byte aligned and ipitch is a multiple of 4 pixels.
template <bool BULK>
__global__ void __launch_bounds__(kTableThreads)
k_build_table(const uchar4* __restrict__ img, int ipitch, int W, int H, uint4* __restrict__ table, int pitch)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ Part s_warp[kTableThreads / 32];
    __shared__ Part s_carry;

    const int r = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uchar4* row = img + (size_t)r * ipitch;      // image rows are ipitch pixels apart (a multiple of 4: 16-byte aligned rows)
    uint4* out = table + (size_t)r * pitch;
    const uchar4* src = row;

    if (BULK) {
        uchar4* s_row = reinterpret_cast<uchar4*>(s_raw);
        const unsigned bar = smem_u32(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned bytes = (unsigned)ipitch * 4u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(s_row)), "l"(row), "r"(bytes), "r"(bar) : "memory");
        }
        // everybody waits for phase 0 of the barrier
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        }
        src = s_row;
    }

    if (tid == 0) { s_carry = Part{0, 0, 0, 0, 0}; out[0] = make_uint4(0, 0, 0, 0); }
    __syncthreads();

    for (int x0 = 0; x0 < W; x0 += kTableThreads) {
        const int x = x0 + tid;
        Part v = Part{0, 0, 0, 0, 0};
        if (x < W) v = part_of(BULK ? src[x] : __ldg(src + x));
        // inclusive scan inside the warp
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
        if (x < W) out[x + 1] = part_pack(v);
        __syncthreads();
        if (tid == kTableThreads - 1) s_carry = v;
        __syncthreads();
    }
}

// Optional resampler for the reference's 1/1.5 window (Tiny::window(w/1.5, h/1.5),
// software/triangulate/main.cpp:53): bilinear fetch at the raster pixel centre, float32 with one
// rounding per operation, truncated to a byte like triangle.fs:31-33 truncates int(255*c).
__global__ void k_resample(const uchar4* __restrict__ src, int w, int h, uchar4* __restrict__ dst, int rpitch, int rw, int rh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= rw || j >= rh) return;
    // texel-space position of the raster pixel centre: u*w - 0.5
    const float fx = __fsub_rn(__fdiv_rn(__fmul_rn(__fadd_rn((float)i, 0.5f), (float)w), (float)rw), 0.5f);
    const float fy = __fsub_rn(__fdiv_rn(__fmul_rn(__fadd_rn((float)j, 0.5f), (float)h), (float)rh), 0.5f);
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = __fsub_rn(fx, flx), ay = __fsub_rn(fy, fly);
    int x0 = (int)flx, y0 = (int)fly;
    int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    x0 = max(x0, 0); y0 = max(y0, 0);
    const uchar4 p00 = src[(size_t)y0 * w + x0], p10 = src[(size_t)y0 * w + x1];
    const uchar4 p01 = src[(size_t)y1 * w + x0], p11 = src[(size_t)y1 * w + x1];
    auto mix = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) {
        const float top = __fadd_rn(__fmul_rn((float)a, __fsub_rn(1.0f, ax)), __fmul_rn((float)b, ax));
        const float bot = __fadd_rn(__fmul_rn((float)c, __fsub_rn(1.0f, ax)), __fmul_rn((float)d, ax));
        const float v = __fadd_rn(__fmul_rn(top, __fsub_rn(1.0f, ay)), __fmul_rn(bot, ay));
        return (unsigned char)min(max((int)v, 0), 255);
    };
    dst[(size_t)j * rpitch + i] = make_uchar4(mix(p00.x, p10.x, p01.x, p11.x), mix(p00.y, p10.y, p01.y, p11.y),
                                          mix(p00.z, p10.z, p01.z, p11.z), 255);
}

}  // namespace tpose_dev

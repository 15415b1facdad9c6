// common.cuh -- device-side conventions shared by every kernel of the hot path.
//
// Vertex transform follows software/triangulate/shader/triangle.vs:58-84 (warp copy :61-87);
// the raster conventions the GL driver would decide (snap, fill rule, clipping) are the ones
// written down in DESIGN.md "Raster conventions" and restated by oracle/step_oracle.c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpose_dev {

constexpr int kVariants = 13;        // base + 3 vertices x 4 directions (triangle.vs:66-78)
constexpr int kEdgeVariants = 27;    // per triangle: 3 edges x (base + 4 moves of a + 4 moves of b)
constexpr float kSnapLimit = 32768.0f;

// dp cooling: triangle.vs:60-62 (triangulate), warp/shader/triangle.vs:63-65 (warp)
__host__ __device__ inline float dp_for(int nt, int flavour)
{
#ifdef __CUDA_ARCH__
    float den = flavour ? __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(9.0f, (float)nt), 1000.0f))
                        : __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(4.0f, (float)nt), 3000.0f));
    return __fdiv_rn(0.05f, den);
#else
    volatile float a = flavour ? 9.0f * (float)nt : 4.0f * (float)nt;
    volatile float b = flavour ? a / 1000.0f : a / 3000.0f;
    volatile float c = 1.0f + b;
    return 0.05f / c;
#endif
}

struct Xform {
    float ratio, hw, hh;
    int H256;
};

__device__ __forceinline__ int snap256(float w)
{
    w = fminf(fmaxf(w, -kSnapLimit), kSnapLimit);
    return __float2int_rn(__fmul_rn(w, 256.0f));
}

// p + (dx,dy) -> snapped window coordinates, y down.  One rounding per operation, no FMA.
__device__ __forceinline__ int2 to_fixed(float2 p, float dx, float dy, const Xform& xf)
{
    float tx = __fadd_rn(p.x, dx);
    float ty = __fadd_rn(p.y, dy);
    tx = __fdiv_rn(tx, xf.ratio);
    float xw = __fadd_rn(__fmul_rn(tx, xf.hw), xf.hw);
    float yw = __fadd_rn(__fmul_rn(ty, xf.hh), xf.hh);
    return make_int2(snap256(xw), xf.H256 - snap256(yw));
}

// direction s of the finite difference (triangle.vs:66-78): +x, -x, +y, -y
__device__ __forceinline__ void fd_offset(int s, float dp, float& dx, float& dy)
{
    dx = (s == 0) ? dp : (s == 1) ? -dp : 0.0f;
    dy = (s == 2) ? dp : (s == 3) ? -dp : 0.0f;
}

__device__ __forceinline__ int64_t floordiv64(int64_t a, int64_t b)  // b > 0
{
    int64_t q = a / b;
    if ((a % b) < 0) --q;
    return q;
}

// Scan conversion of one edge: for row r (centre 256r+128) the number of pixel centres of that
// row strictly left of the edge, c(r) = ceil((x_edge(yc) - 128) / 256), kept as an incremental
// quotient/remainder pair.  Rows are visited r, r+stride, ...
//
// With the edge sorted so that ya < yb, dy = yb-ya, dx = xb-xa, m = yc-ya >= 0:
//   c(r) = floor( (xa + 128 + floor((dx*m - 1)/dy)) / 256 )
// (nested-floor identity; checked against the plain ceil-division on 3e5 random edges), which
// needs one division by dy instead of one by 256*dy and stays in 32 bits for on-screen edges.
struct EdgeWalk {
    long long rem, den, dr;
    int q, dq;
    int r, r_end;       // first row of this lane, one past the last row of the edge (clipped)
    int sign;           // +1 if the directed edge a->b goes down (yb > ya), -1 if up, 0 if horizontal
    __device__ __forceinline__ void init(int2 a, int2 b, int H, int lane, int stride)
    {
        sign = (b.y > a.y) - (b.y < a.y);
        if (a.y > b.y) { int2 t = a; a = b; b = t; }
        const int lo = a.y, hi = b.y;
        // rows with lo <= 256r+128 < hi  <=>  ceil((lo-128)/256) <= r < ceil((hi-128)/256)
        int r0 = (lo - 128 + 255) >> 8;
        int r1 = (hi - 128 + 255) >> 8;
        r0 = max(r0, 0);
        r_end = min(r1, H);
        r = r0 + lane;
        q = 0; dq = 0; rem = 0; den = 1; dr = 0;
        if (sign == 0 || r >= r_end) { r = r_end; return; }
        const int dy = hi - lo, dx = b.x - a.x;       // |.| <= 2^24
        const int m = 256 * r + 128 - lo;             // 0 <= m < 2^24
        int w, wr;
        const unsigned adx = (unsigned)abs(dx);
        if (__umulhi(adx, (unsigned)m) == 0 && adx * (unsigned)m < 0x7fffffffu) {   // 32-bit path
            const int z = dx * m - 1;
            w = z / dy; wr = z - w * dy;
            if (wr < 0) { wr += dy; --w; }
        } else {
            const long long z = (long long)dx * m - 1;
            long long w64 = z / dy, wr64 = z - w64 * dy;
            if (wr64 < 0) { wr64 += dy; --w64; }
            w = (int)w64; wr = (int)wr64;
        }
        const int s = a.x + 128 + w;
        q = s >> 8;
        den = 256ll * dy;
        rem = (long long)(s & 255) * dy + wr;
        const int t = stride * dx;                    // |t| <= 2^29
        dq = t / dy;
        int tr = t - dq * dy;
        if (tr < 0) { tr += dy; --dq; }
        dr = 256ll * tr;
    }
    // The rows of the edge that fall inside [row_lo, row_hi) (a tile), stride 1: the same state as
    // init() followed by next() up to the first of those rows.
    __device__ __forceinline__ void init_rows(int2 a, int2 b, int row_lo, int row_hi)
    {
        sign = (b.y > a.y) - (b.y < a.y);
        if (a.y > b.y) { int2 t = a; a = b; b = t; }
        const int lo = a.y, hi = b.y;
        const int r0 = max((lo - 128 + 255) >> 8, row_lo);
        r_end = min((hi - 128 + 255) >> 8, row_hi);
        r = r0;
        q = 0; dq = 0; rem = 0; den = 1; dr = 0;
        if (sign == 0 || r >= r_end) { r = r_end; return; }
        const int dy = hi - lo, dx = b.x - a.x;
        const int m = 256 * r + 128 - lo;             // 0 <= m < dy
        int w, wr;
        const unsigned adx = (unsigned)abs(dx);
        if (__umulhi(adx, (unsigned)m) == 0 && adx * (unsigned)m < 0x7fffffffu) {
            const int z = dx * m - 1;
            w = z / dy; wr = z - w * dy;
            if (wr < 0) { wr += dy; --w; }
        } else {
            const long long z = (long long)dx * m - 1;
            long long w64 = z / dy, wr64 = z - w64 * dy;
            if (wr64 < 0) { wr64 += dy; --w64; }
            w = (int)w64; wr = (int)wr64;
        }
        const int s = a.x + 128 + w;
        q = s >> 8;
        den = 256ll * dy;
        rem = (long long)(s & 255) * dy + wr;
        dq = dx / dy;
        int tr = dx - dq * dy;
        if (tr < 0) { tr += dy; --dq; }
        dr = 256ll * tr;
    }
    __device__ __forceinline__ bool done() const { return r >= r_end; }
    __device__ __forceinline__ int cross(int W) const { return min(max(q, 0), W); }
    __device__ __forceinline__ void next(int stride)
    {
        r += stride;
        q += dq; rem += dr;
        if (rem >= den) { rem -= den; ++q; }
    }
};

// Packed row-prefix moment entry (16 bytes): exclusive prefix over pixels [0,c) of one row.
//   x: sum R (20 bits) | count of pixels with odd R+G+B (12 bits, << 20)
//   y: sum G   z: sum B   w: sum (R^2+G^2+B^2)
// Width <= 4095 keeps every field in range (255*4095 < 2^20, 195075*4095 < 2^32).
struct Moments {
    int n, odd;
    long long r, g, b, s2;
    __device__ __forceinline__ void zero() { n = odd = 0; r = g = b = s2 = 0; }
    __device__ __forceinline__ void add_entry(uint4 e, int c)
    {
        n += c;
        odd += (int)(e.x >> 20);
        r += (long long)(e.x & 0xFFFFFu);
        g += (long long)e.y;
        b += (long long)e.z;
        s2 += (long long)e.w;
    }
};

// energy of a covered set from its moments and an integer mean colour (DESIGN.md "closed form"):
//   sum_px trunc(0.5*|b-m|^2) = ( S2 - 2 m.S + n |m|^2 - #{px : parity(b) != parity(m)} ) / 2
__device__ __forceinline__ long long energy_closed_form(int n, int odd, long long sr, long long sg,
                                                        long long sb, long long s2, int m0, int m1, int m2)
{
    const long long mm = (long long)m0 * m0 + (long long)m1 * m1 + (long long)m2 * m2;
    const long long ms = (long long)m0 * sr + (long long)m1 * sg + (long long)m2 * sb;
    const int mism = ((m0 + m1 + m2) & 1) ? (n - odd) : odd;
    return (s2 - 2 * ms + (long long)n * mm - (long long)mism) >> 1;
}

}  // namespace tpose_dev

// DRAM fetch granularity probe (B200): how many bytes does DRAM deliver for one cold 16-byte load?
//   k_stride<S>: 1 M threads, one __ldg of 16 bytes each at stride S        -> profiles/r02y_fetch_granularity_probe_strides.csv
//   k_kind<K>:   the same at stride 512 through other instructions          -> profiles/r02y_fetch_granularity_probe_flavours.csv
//                (0 __ldg, 1 ld.global.L2::64B, 2 cp.async.cg, 3 cp.async.ca, 4 / 5 cp.async.bulk of 16 / 32 bytes)
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o fetch_granularity fetch_granularity.cu
// Run:   ncu --clock-control none --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum ./fetch_granularity [limit]
// Result: 32 / 64 / 128 / 128 / 128 bytes per load at stride 32 / 64 / 128 / 256 / 512 -- a plain load that misses L2
// pulls the whole 128-byte line; .L2::64B and bulk copies pull 64; cudaLimitMaxL2FetchGranularity (the optional
// argument) is accepted and changes nothing.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int S> __global__ void k_stride(const char* base, size_t n, unsigned* sink)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t j = (i * 2654435761ull) % n;          // neighbours in a warp are not neighbours in memory
    uint4 v = __ldg(reinterpret_cast<const uint4*>(base + j * S));
    if (v.x + v.y + v.z + v.w == 0x12345678u) *sink = 1;
}
template <int KIND> __global__ void k_kind(const char* base, size_t n, unsigned* sink)
{
    __shared__ __align__(128) uint4 s[256];
    __shared__ __align__(8) unsigned long long bar;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const uint4* p = reinterpret_cast<const uint4*>(base + i * 512);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (KIND == 0) v = __ldg(p);
    else if (KIND == 1) asm volatile("ld.global.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (KIND == 2) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s[threadIdx.x])), "l"(p));
        asm volatile("cp.async.commit_group;"); asm volatile("cp.async.wait_group 0;" ::: "memory");
        v = s[threadIdx.x];
    } else if (KIND == 3) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s[threadIdx.x])), "l"(p));
        asm volatile("cp.async.commit_group;"); asm volatile("cp.async.wait_group 0;" ::: "memory");
        v = s[threadIdx.x];
    } else if (KIND == 4 || KIND == 5) {
        // one bulk copy of 16 (KIND 4) or 32 (KIND 5) bytes per thread, all on one mbarrier
        const unsigned b = smem_u32(&bar);
        const unsigned bytes = KIND == 4 ? 16u : 32u;
        if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncthreads();
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes * 128u) : "memory");
        __syncthreads();
        if (threadIdx.x < 128)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(&s[threadIdx.x * 2])), "l"(p), "r"(bytes), "r"(b) : "memory");
        unsigned done = 0;
        while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(0u) : "memory");
        v = s[threadIdx.x];
    }
    if (v.x + v.y + v.z + v.w == 0x12345678u) *sink = 1;
}
int main(int argc, char** argv)
{
    if (argc > 1) {
        size_t g = 0; cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now %zu\n", g);
    }
    const size_t n = 1 << 20;
    char* buf; unsigned* sink;
    cudaMalloc(&buf, n * 512); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, n * 512);
    const int T = 256, B = (int)(n / T);
    k_stride<32><<<B, T>>>(buf, n, sink);
    k_stride<64><<<B, T>>>(buf, n, sink);
    k_stride<128><<<B, T>>>(buf, n, sink);
    k_stride<256><<<B, T>>>(buf, n, sink);
    k_stride<512><<<B, T>>>(buf, n, sink);
    k_kind<0><<<B, T>>>(buf, n, sink);
    k_kind<1><<<B, T>>>(buf, n, sink);
    k_kind<2><<<B, T>>>(buf, n, sink);
    k_kind<3><<<B, T>>>(buf, n, sink);
    k_kind<4><<<B, T>>>(buf, n, sink);
    k_kind<5><<<B, T>>>(buf, n, sink);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

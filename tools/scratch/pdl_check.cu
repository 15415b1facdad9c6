// Does programmatic dependent launch overlap a secondary kernel's prologue with its primary -- in a plain
// stream and inside a captured graph?  Prints globaltimer stamps (ns, relative).
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void A(unsigned long long* out, int spin_us, int early) {
    if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned long long t0 = gt();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t0;
    while (gt() - t0 < (unsigned long long)spin_us * 1000ull) {}
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = gt();
}
__global__ void B(unsigned long long* out) {
    unsigned long long t0 = gt();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long t1 = gt();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[2] = t0; out[3] = t1; }
}
static void launch(cudaStream_t st, unsigned long long* d, int pdl, int early) {
    A<<<64, 128, 0, st>>>(d, 30, early);
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(64); cfg.blockDim = dim3(128); cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, B, d);
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64); unsigned long long h[4];
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int mode = 0; mode < 2; mode++) for (int pdl = 0; pdl < 2; pdl++) for (int early = 0; early < 2; early++) {
        for (int rep = 0; rep < 3; rep++) {
            if (mode == 0) launch(st, d, pdl, early);
            else {
                cudaGraph_t g; cudaGraphExec_t e;
                cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
                launch(st, d, pdl, early);
                cudaError_t er = cudaStreamEndCapture(st, &g);
                if (er != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(er)); return 1; }
                er = cudaGraphInstantiate(&e, g, 0);
                if (er != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(er)); return 1; }
                cudaGraphLaunch(e, st); cudaStreamSynchronize(st); cudaGraphLaunch(e, st);
            }
            cudaStreamSynchronize(st);
        }
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("%s pdl=%d early_trigger=%d: A %5.1f us long; B started %+6.1f us after A's end, passed wait %+6.1f us after A's end  (%s)\n",
               mode ? "graph " : "stream", pdl, early, (h[1] - h[0]) / 1e3, ((double)h[2] - (double)h[1]) / 1e3, ((double)h[3] - (double)h[1]) / 1e3,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

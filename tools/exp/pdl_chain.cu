// Micro-experiment: latency per dependent kernel node in a CUDA graph, with and without programmatic dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_plain(float* p, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
__global__ void k_pdl(float* p, int n) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
int main() {
    float* d; int n = 128 * 256;
    cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    for (int mode = 0; mode < 2; mode++) {
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < 40; k++) {
            if (mode == 0) k_plain<<<128, 256, 0, st>>>(d, n);
            else {
                cudaLaunchConfig_t cfg = {}; cfg.gridDim = 128; cfg.blockDim = 256; cfg.stream = st;
                cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                a[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = a; cfg.numAttrs = 1;
                cudaLaunchKernelEx(&cfg, k_pdl, d, n);
            }
        }
        cudaStreamEndCapture(st, &g);
        cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
        if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int w = 0; w < 5; w++) cudaGraphLaunch(ge, st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int w = 0; w < 50; w++) cudaGraphLaunch(ge, st);
        cudaEventRecord(e1, st); cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s: %.2f us per kernel node (40-node chain, 50 replays) err=%s\n", mode ? "PDL  " : "plain", ms * 1e3 / (50 * 40), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

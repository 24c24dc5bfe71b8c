// Probe: does programmatic stream serialization overlap the serial tail of kernel N with the body of kernel N+1?
// Every CTA spins ~body cycles; the "finisher" (last CTA) then waits on the grid dependency, triggers, and spins ~tail.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *p, long long body, long long tail, int early) {
    long long t0 = clock64();
    while (clock64() - t0 < body) {}
    const bool fin = blockIdx.x == gridDim.x - 1;
    if (!fin) {
        if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    t0 = clock64();
    while (clock64() - t0 < tail) {}
    if (threadIdx.x == 0) atomicAdd(p, 1);
}
float run(int pdl, int early, int smem_kb, int threads, int grid) {
    int *d;
    cudaMalloc(&d, 4);
    cudaMemset(d, 0, 4);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int reps = 200;
    for (int w = 0; w < 2; w++) {
        if (w == 1) cudaEventRecord(a, s);
        for (int i = 0; i < reps; i++) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem_kb * 1024;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = pdl ? 1 : 0;
            cudaLaunchKernelEx(&cfg, k, d, 40000LL, 20000LL, early);
        }
    }
    cudaEventRecord(b, s);
    cudaStreamSynchronize(s);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("pdl %d early %d smem %3d KB threads %d grid %d: %.2f us per kernel (%s)\n", pdl, early, smem_kb, threads, grid, ms * 1e3 / reps,
           cudaGetErrorString(cudaGetLastError()));
    return ms;
}
int main() {
    run(0, 0, 0, 256, 296);
    run(1, 0, 0, 256, 296);
    run(1, 1, 0, 256, 296);
    run(0, 1, 112, 256, 296);
    run(1, 1, 112, 256, 296);
    return 0;
}

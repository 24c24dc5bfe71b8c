// Probe: PDL kernel whose launcher CTA triggers, then starts a fire-and-forget DRIVER grid that tail-launches two ordered
// children.  Does the next kernel's griddepcontrol.wait cover the driver and its tail-launched children?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void child(int *p, int expect_mod) {
    // children must run in order: c1 sees counter % 4 == 2, c2 sees 3
    int v = atomicAdd(p, 1);
    if ((v & 3) != expect_mod) atomicAdd(p + 1, 1); // order violation count
}
__global__ void driver(int *p) {
    int v = atomicAdd(p, 1);
    if ((v & 3) != 1) atomicAdd(p + 1, 1);
    child<<<1, 1, 0, cudaStreamTailLaunch>>>(p, 2);
    child<<<1, 1, 0, cudaStreamTailLaunch>>>(p, 3);
}
__global__ void k(int *p, int *log, int rep, int spin) {
    // streaming part: nothing
    for (volatile int i = 0; i < spin; i++) {}
    const bool launcher = threadIdx.x == 0 && blockIdx.x == gridDim.x - 1;
    if (!launcher) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        return;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    log[rep] = *(volatile int *)p; // must be 4 * rep: everything of the previous kernel, descendants included, is done
    atomicAdd(p, 1);
    driver<<<1, 1, 0, cudaStreamFireAndForget>>>(p);
}
int main() {
    int *d, *log;
    const int reps = 200;
    cudaMalloc(&d, 8);
    cudaMemset(d, 0, 8);
    cudaMalloc(&log, reps * 4);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < reps; i++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(296);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        a[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = a;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k, d, log, i, 20000);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    int h[2], hl[reps];
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hl, log, reps * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < reps; i++) bad += (hl[i] != 4 * i);
    printf("chain: %s, counter %d (want %d), order violations %d, reps that saw an incomplete predecessor %d (log[1]=%d log[2]=%d)\n",
           cudaGetErrorString(e), h[0], 4 * reps, h[1], bad, hl[1], hl[2]);
    return 0;
}

// Minimal probe: griddepcontrol in a kernel built with -rdc=true (device runtime linked), launched normally and with PDL,
// with and without a device-side tail launch.  Prints what completes.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void child(int *p) { atomicAdd(p, 1); }
// mode bits: 1 trigger, 2 wait, 4 = block 0 launches its child BEFORE it triggers, 8 = fire-and-forget instead of tail launch,
// 16 = block 0 never triggers explicitly
__global__ void k(int *p, int mode, int do_child) {
    const bool launcher = threadIdx.x == 0 && blockIdx.x == 0;
    if (mode & 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    if ((mode & 4) && launcher && do_child) {
        if (mode & 8) child<<<1, 1, 0, cudaStreamFireAndForget>>>(p);
        else child<<<1, 1, 0, cudaStreamTailLaunch>>>(p);
    }
    if ((mode & 1) && !((mode & 16) && blockIdx.x == 0)) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (launcher) {
        atomicAdd(p, 1);
        if (do_child && !(mode & 4)) {
            if (mode & 8) child<<<1, 1, 0, cudaStreamFireAndForget>>>(p);
            else child<<<1, 1, 0, cudaStreamTailLaunch>>>(p);
        }
    }
}
int run(const char *name, int mode, int pdl, int do_child, int reps) {
    int *d;
    cudaMalloc(&d, 4);
    cudaMemset(d, 0, 4);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int i = 0; i < reps; i++) {
        if (pdl) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(296);
            cfg.blockDim = dim3(256);
            cfg.stream = s;
            cudaLaunchAttribute a[1];
            a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            a[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = a;
            cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k, d, mode, do_child);
        } else {
            k<<<296, 256, 0, s>>>(d, mode, do_child);
        }
    }
    cudaError_t e = cudaStreamSynchronize(s);
    int h = -1;
    cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
    printf("%-40s mode %d pdl %d child %d reps %d -> %s, counter %d\n", name, mode, pdl, do_child, reps, cudaGetErrorString(e), h);
    fflush(stdout);
    return 0;
}
int main(int argc, char **argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    switch (which) {
    case 0: return run("plain launch, no griddepcontrol", 0, 0, 0, 10);
    case 1: return run("plain launch, wait only", 2, 0, 0, 10);
    case 2: return run("plain launch, trigger+wait", 3, 0, 0, 10);
    case 3: return run("pdl launch, trigger+wait", 3, 1, 0, 10);
    case 4: return run("plain launch, trigger+wait, child", 3, 0, 1, 10);
    case 5: return run("pdl launch, trigger+wait, child", 3, 1, 1, 10);
    case 6: return run("pdl launch, no griddepcontrol, child", 0, 1, 1, 10);
    case 7: return run("plain, wait only, tail child", 2, 0, 1, 10);
    case 8: return run("plain, trigger only, tail child after", 1, 0, 1, 10);
    case 9: return run("plain, child before trigger", 1 | 2 | 4, 0, 1, 10);
    case 10: return run("plain, trigger+wait, f&f child after", 1 | 2 | 8, 0, 1, 10);
    case 11: return run("pdl, child before trigger", 1 | 2 | 4, 1, 1, 10);
    case 12: return run("pdl, launcher block never triggers", 1 | 2 | 16, 1, 1, 10);
    case 13: return run("pdl, f&f child after trigger", 1 | 2 | 8, 1, 1, 10);
    case 14: return run("pdl, wait only, tail child", 2, 1, 1, 10);
    }
    return 0;
}

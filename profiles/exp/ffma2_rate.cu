// ffma2_rate.cu -- issue rate of packed FP32 (FFMA2 / FADD2) against scalar FFMA on sm_100a, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu ; ./ffma2_rate
// Each thread runs CH independent accumulator chains for ITERS trips; one CTA of W warps per SM (W = 4, 8, 16: 1, 2, 4
// warps per scheduler).  Prints cycles per warp instruction per scheduler and the FMA lanes per clock per SM it implies.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void k(float *out, int iters, long long *cycles) {
    float2 acc[CH];
    const float2 a = make_float2(1.0f + threadIdx.x * 1e-9f, 1.0f - threadIdx.x * 1e-9f), b = make_float2(1e-9f, -1e-9f);
    float2 xa[CH], xb[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) {
        acc[c] = make_float2(c * 0.5f, c * 0.25f);
        xa[c] = make_float2(1.0f + out[c] * 1e-9f, 1.0f - out[c + 32] * 1e-9f);
        xb[c] = make_float2(1e-9f * out[c + 64], -1e-9f * out[c + 96]);
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i0 = 0; i0 < iters; i0 += CH) {
#pragma unroll
      for (int i = 0; i < CH; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0) {        // scalar FFMA x2 (two instructions, two lanes-worth of work each ... i.e. 2 FFMA)
                acc[c].x = __fmaf_rn(acc[c].x, a.x, b.x);
                acc[c].y = __fmaf_rn(acc[c].y, a.y, b.y);
            } else if (MODE == 1) { // packed FFMA2, all operands register pairs
                acc[c] = __ffma2_rn(acc[c], a, b);
            } else if (MODE == 2) { // packed FADD2
                acc[c] = __fadd2_rn(acc[c], b);
            } else if (MODE == 3) { // FFMA2 with a scalar-broadcast multiplicand (R.F32 form)
                acc[c] = __ffma2_rn(make_float2(a.x, a.x), acc[c], b);
            } else if (MODE == 4) { // FFMA2 with three distinct vector-register pairs per instruction (the RMSD loop's form)
                acc[c] = __ffma2_rn(xa[c], xb[(c + i) & (CH - 1)], acc[c]);
            } else if (MODE == 5) { // scalar-broadcast multiplicand + two distinct pairs
                acc[c] = __ffma2_rn(make_float2(xa[c].x, xa[c].x), xb[(c + i) & (CH - 1)], acc[c]);
            } else if (MODE == 6) { // scalar FFMA, three distinct registers
                acc[c].x = __fmaf_rn(xa[c].x, xb[(c + i) & (CH - 1)].x, acc[c].x);
                acc[c].y = __fmaf_rn(xa[c].y, xb[(c + i) & (CH - 1)].y, acc[c].y);
            }
        }
      }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CH; c++) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE, int CH>
void run(const char *name, int warps) {
    float *out;
    long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 8);
    const int iters = 20480;
    k<MODE, CH><<<148, warps * 32>>>(out, iters, cyc);
    k<MODE, CH><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst_per_thread = (double)iters * CH * ((MODE == 0 || MODE == 6) ? 2 : 1);
    const double warp_inst_per_sched = inst_per_thread * warps / 4.0;
    const double cpi = (double)h / warp_inst_per_sched;
    const double lanes = ((MODE == 0 || MODE == 6) ? 1.0 : 2.0) * 32.0 / cpi * 4.0;  // FP32 lanes (FMA or ADD results) per clock per SM
    printf("%-34s warps/SM %2d chains %2d: %.3f cycles per warp instruction per scheduler -> %.1f FP32 results/clk/SM\n", name, warps, CH, cpi, lanes);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0, 8>("scalar FFMA", w);
        run<1, 8>("FFMA2 (pairs)", w);
        run<2, 8>("FADD2", w);
        run<3, 8>("FFMA2 (scalar-broadcast operand)", w);
    }
    for (int w : {4, 8, 16}) {
        run<4, 8>("FFMA2, 3 distinct register pairs", w);
        run<5, 8>("FFMA2, broadcast + 2 distinct pairs", w);
        run<6, 8>("scalar FFMA, 3 distinct registers", w);
    }
    // dependent-issue latency: one warp per scheduler, 1 / 2 / 4 chains
    run<0, 1>("scalar FFMA, 1 chain (x2 instr)", 4);
    run<1, 1>("FFMA2, 1 chain", 4);
    run<1, 2>("FFMA2, 2 chains", 4);
    run<1, 4>("FFMA2, 4 chains", 4);
    run<2, 1>("FADD2, 1 chain", 4);
    run<3, 1>("FFMA2 broadcast, 1 chain", 4);
    run<1, 16>("FFMA2 (pairs)", 16);
    run<0, 16>("scalar FFMA", 16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

// kernels_center.cuh -- periodic centre of geometry / mass over a batch of frames.
//
// Reference-order ("exact") passes.  Per-atom arithmetic is the reference's, operation for operation
// and in f32 (so every per-atom value, in particular every image decision, is the reference's);
// only the SUMMATION differs: per-thread f32 partials over <= 64 atoms, then f64 tree reduction
// (the reference sums sequentially in f32 and drifts; see DESIGN.md "Accumulation").
//
//   k_trig    Bai-Breen circular mean          iterators.rs:1152-1191 / 1314-1357, auxiliary.rs:59-99
//   k_unwrap  refined pass: unwrap around c0   iterators.rs:1237-1266 / 1404-1438, vector3d.rs:561-569
//   k_naive   plain mean                       iterators.rs:886-903
//
// Grid = (blocks per frame, frames); the last CTA of each frame finishes the frame (common.cuh).
#pragma once
#include "common.cuh"

namespace groan {

constexpr int kFlush = 64; // f32 partials are flushed to f64 every kFlush atoms

template <bool WEIGHTED>
__global__ void __launch_bounds__(kThreads) k_trig(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                    float *c0_out) {
    __shared__ double smem[6 * (kThreads / 32)];
    __shared__ int sh_flag;
    const int f = blockIdx.y, nb = gridDim.x;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float sx = pi_x2() / lx, sy = pi_x2() / ly, sz = pi_x2() / lz; // iterators.rs:1154
    const float *fr = fv.frame(f);
    float a[6] = {0, 0, 0, 0, 0, 0};
    double d[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += nb * blockDim.x) {
        const float *p = fr + (size_t)g.atom(i) * 3;
        const float m = WEIGHTED ? __ldg(g.mass + i) : 1.0f;
        float s, c;
        // auxiliary.rs:59-83: wrap, theta = pos * scaling, sum_xi += m cos, sum_zeta += m sin
        sincosf(wrap_coordinate(__ldg(p + 0), lx) * sx, &s, &c);
        a[0] += m * c; a[3] += m * s;
        sincosf(wrap_coordinate(__ldg(p + 1), ly) * sy, &s, &c);
        a[1] += m * c; a[4] += m * s;
        sincosf(wrap_coordinate(__ldg(p + 2), lz) * sz, &s, &c);
        a[2] += m * c; a[5] += m * s;
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 6; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) d[k] += (double)a[k];
    block_sum<6>(d, smem);
    double tot[6];
    if (frame_finish<6>(d, partials + (size_t)f * nb * 6, tickets + f, nb, tot, &sh_flag) && threadIdx.x == 0) {
        // auxiliary.rs:87-99: (atan2(-zeta, -xi) + PI) / scaling, in f32
        const float sc[3] = {sx, sy, sz};
        for (int k = 0; k < 3; k++) {
            const float xi = (float)tot[k], ze = (float)tot[3 + k];
            c0_out[f * 3 + k] = (atan2f(-ze, -xi) + 3.14159265358979323846f) / sc[k];
        }
    }
}

// centre = sum(m * (c0 + vector_to(c0, x))) / sum(m); geometry: m = 1, divisor = n
template <bool WEIGHTED>
__global__ void __launch_bounds__(kThreads) k_unwrap(FrameView fv, GroupView g, const float *c0_in, double *partials,
                                                      unsigned int *tickets, float *out) {
    __shared__ double smem[4 * (kThreads / 32)];
    __shared__ int sh_flag;
    const int f = blockIdx.y, nb = gridDim.x;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float cx = c0_in[f * 3 + 0], cy = c0_in[f * 3 + 1], cz = c0_in[f * 3 + 2];
    const float *fr = fv.frame(f);
    float a[4] = {0, 0, 0, 0};
    double d[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += nb * blockDim.x) {
        const float *p = fr + (size_t)g.atom(i) * 3;
        const float m = WEIGHTED ? __ldg(g.mass + i) : 1.0f;
        const float nx = cx + vector_to_1(cx, __ldg(p + 0), lx);
        const float ny = cy + vector_to_1(cy, __ldg(p + 1), ly);
        const float nz = cz + vector_to_1(cz, __ldg(p + 2), lz);
        a[0] += nx * m; a[1] += ny * m; a[2] += nz * m; a[3] += m;
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 4; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) d[k] += (double)a[k];
    block_sum<4>(d, smem);
    double tot[4];
    if (frame_finish<4>(d, partials + (size_t)f * nb * 4, tickets + f, nb, tot, &sh_flag) && threadIdx.x == 0) {
        const double div = WEIGHTED ? tot[3] : (double)g.n;
        for (int k = 0; k < 3; k++) out[f * 3 + k] = (float)(tot[k] / div);
    }
}

__global__ void __launch_bounds__(kThreads) k_naive(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                     float *out) {
    __shared__ double smem[3 * (kThreads / 32)];
    __shared__ int sh_flag;
    const int f = blockIdx.y, nb = gridDim.x;
    const float *fr = fv.frame(f);
    float a[3] = {0, 0, 0};
    double d[3] = {0, 0, 0};
    int cnt = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += nb * blockDim.x) {
        const float *p = fr + (size_t)g.atom(i) * 3;
        a[0] += __ldg(p + 0); a[1] += __ldg(p + 1); a[2] += __ldg(p + 2);
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 3; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] += (double)a[k];
    block_sum<3>(d, smem);
    double tot[3];
    if (frame_finish<3>(d, partials + (size_t)f * nb * 3, tickets + f, nb, tot, &sh_flag) && threadIdx.x == 0)
        for (int k = 0; k < 3; k++) out[f * 3 + k] = (float)(tot[k] / (double)g.n);
}

// Vector3D::distance between two per-frame centres (System::group_distance, analysis.rs:348-360)
__device__ __forceinline__ float distance_dim(float ax, float ay, float az, float bx, float by, float bz, int dim, float lx,
                                              float ly, float lz) {
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    switch (dim) {
    case 0: return 0.0f;
    case 1: return min_image(ax - bx, lx);
    case 2: return min_image(ay - by, ly);
    case 3: return min_image(az - bz, lz);
    case 4: dx = min_image(ax - bx, lx); dy = min_image(ay - by, ly); break;
    case 5: dx = min_image(ax - bx, lx); dz = min_image(az - bz, lz); break;
    case 6: dy = min_image(ay - by, ly); dz = min_image(az - bz, lz); break;
    default: dx = min_image(ax - bx, lx); dy = min_image(ay - by, ly); dz = min_image(az - bz, lz); break;
    }
    return sqrtf((dx * dx + dy * dy) + dz * dz); // nalgebra Vector3::magnitude (vector3d.rs:467-483)
}

__global__ void k_center_distance(const float *c1, const float *c2, FrameView fv, int dim, int n_frames, float *out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    out[f] = distance_dim(c1[f * 3], c1[f * 3 + 1], c1[f * 3 + 2], c2[f * 3], c2[f * 3 + 1], c2[f * 3 + 2], dim, lx, ly, lz);
}

} // namespace groan

// kernels_center.cuh -- periodic centre of geometry / mass over a batch of frames.
//
// (1) Reference-order ("exact") passes.  Per-atom arithmetic is the reference's, operation for operation
//     and in f32 (so every per-atom value, in particular every image decision, is the reference's);
//     only the SUMMATION differs: per-thread f32 partials over <= 64 atoms, then an f64 tree
//     (the reference sums sequentially in f32 and drifts; see DESIGN.md "Accumulation").
//
//       k_trig    Bai-Breen circular mean          iterators.rs:1152-1191 / 1314-1357, auxiliary.rs:59-99
//       k_unwrap  refined pass: unwrap around c0   iterators.rs:1237-1266 / 1404-1438, vector3d.rs:561-569
//       k_naive   plain mean                       iterators.rs:886-903
//
// (2) Single-pass kernel k_center_fast for group_get_center / group_get_com (DESIGN.md "Single-pass
//     centre"): every atom is unwrapped around a PILOT (the group's first atom) instead of around the
//     Bai-Breen estimate c0, and the sums, the extent of the unwrapped group and the circular-mean sums
//     are taken in the same pass.  If the unwrapped group is shorter than half the box on every axis
//     -- the reference's own validity condition (iterators.rs:1195) -- c0 provably lies inside the
//     group's arc, no atom changes image between the two choices of origin, and the reference's result
//     is mean(unwrapped) + k*L with k fixed by c0.  c0 is only needed to pick k, so its trig sums use
//     the SFU (MUFU.SIN/COS); frames where that decision is within the SFU error of the box edge, or
//     whose group is not compact, are flagged and re-done by the exact passes.
//
// Grid = (CTAs per frame, frames); the last CTA of each frame finishes the frame (common.cuh).
#pragma once
#include "common.cuh"

namespace groan {

constexpr int kFlush = 64; // f32 partials are flushed to f64 every kFlush atoms

// frames are skipped when a flag array is given and the frame's flag is zero (exact passes re-doing
// only the frames a single-pass kernel could not certify)
__device__ __forceinline__ bool frame_skipped(const int *flags, int f) { return flags && flags[f] == 0; }

template <bool WEIGHTED>
__global__ void __launch_bounds__(kThreads) k_trig(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                    float *c0_out, const int *flags, int to_cartesian = 0) {
    __shared__ FrameReduceSmem<6, 0> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    if (frame_skipped(flags, f)) return;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float sx = pi_x2() / lx, sy = pi_x2() / ly, sz = pi_x2() / lz; // iterators.rs:1154
    float a[6] = {0, 0, 0, 0, 0, 0};
    double d[6] = {0, 0, 0, 0, 0, 0};
    int cnt = 0;
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float m = WEIGHTED ? __ldg(g.mass + i) : 1.0f;
        float s, c;
        // auxiliary.rs:59-83: wrap, theta = pos * scaling, sum_xi += m cos, sum_zeta += m sin
        sincosf(wrap_coordinate(x, lx) * sx, &s, &c);
        a[0] += m * c; a[3] += m * s;
        sincosf(wrap_coordinate(y, ly) * sy, &s, &c);
        a[1] += m * c; a[4] += m * s;
        sincosf(wrap_coordinate(z, lz) * sz, &s, &c);
        a[2] += m * c; a[5] += m * s;
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 6; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    });
#pragma unroll
    for (int k = 0; k < 6; k++) d[k] += (double)a[k];
    double tot[6];
    if (frame_reduce<6, 0>(d, nullptr, nullptr, partials + (size_t)f * nb * 6, tickets + f, nb, sm, tot, nullptr, nullptr) &&
        threadIdx.x == 0) {
        // auxiliary.rs:87-99: (atan2(-zeta, -xi) + PI) / scaling, in f32
        const float sc[3] = {sx, sy, sz};
        for (int k = 0; k < 3; k++) {
            const float xi = (float)tot[k], ze = (float)tot[3 + k];
            c0_out[f * 3 + k] = (atan2f(-ze, -xi) + 3.14159265358979323846f) / sc[k];
        }
        if (fv.tric && to_cartesian) fv.shear(f).to_x(c0_out[f * 3], c0_out[f * 3 + 1], c0_out[f * 3 + 2]);
    }
}

// centre = sum(m * (c0 + vector_to(c0, x))) / sum(m); geometry: m = 1, divisor = n
template <bool WEIGHTED>
// to_cartesian (triclinic extension only): 1 = `out` is a result and goes back through Shear::to_x; 0 = it is the COM the
// covariance pass will use, which works in the sheared picture too
__global__ void __launch_bounds__(kThreads) k_unwrap(FrameView fv, GroupView g, const float *c0_in, double *partials,
                                                      unsigned int *tickets, float *out, const int *flags, int to_cartesian = 1) {
    __shared__ FrameReduceSmem<4, 0> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    if (frame_skipped(flags, f)) return;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float cx = c0_in[f * 3 + 0], cy = c0_in[f * 3 + 1], cz = c0_in[f * 3 + 2];
    float a[4] = {0, 0, 0, 0};
    double d[4] = {0, 0, 0, 0};
    int cnt = 0;
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float m = WEIGHTED ? __ldg(g.mass + i) : 1.0f;
        const float nx = cx + vector_to_1(cx, x, lx);
        const float ny = cy + vector_to_1(cy, y, ly);
        const float nz = cz + vector_to_1(cz, z, lz);
        a[0] += nx * m; a[1] += ny * m; a[2] += nz * m; a[3] += m;
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 4; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    });
#pragma unroll
    for (int k = 0; k < 4; k++) d[k] += (double)a[k];
    double tot[4];
    if (frame_reduce<4, 0>(d, nullptr, nullptr, partials + (size_t)f * nb * 4, tickets + f, nb, sm, tot, nullptr, nullptr) &&
        threadIdx.x == 0) {
        const double div = WEIGHTED ? tot[3] : (double)g.n;
        for (int k = 0; k < 3; k++) out[f * 3 + k] = (float)(tot[k] / div);
        if (fv.tric && to_cartesian) fv.shear(f).to_x(out[f * 3], out[f * 3 + 1], out[f * 3 + 2]);
    }
}

__global__ void __launch_bounds__(kThreads) k_naive(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                     float *out) {
    __shared__ FrameReduceSmem<3, 0> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float a[3] = {0, 0, 0};
    double d[3] = {0, 0, 0};
    int cnt = 0;
    for_each_group_atom(fv, g, f, [&](uint32_t, float x, float y, float z) {
        a[0] += x; a[1] += y; a[2] += z;
        if (++cnt == kFlush) {
#pragma unroll
            for (int k = 0; k < 3; k++) { d[k] += (double)a[k]; a[k] = 0.0f; }
            cnt = 0;
        }
    });
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] += (double)a[k];
    double tot[3];
    if (frame_reduce<3, 0>(d, nullptr, nullptr, partials + (size_t)f * nb * 3, tickets + f, nb, sm, tot, nullptr, nullptr) &&
        threadIdx.x == 0)
        for (int k = 0; k < 3; k++) out[f * 3 + k] = (float)(tot[k] / (double)g.n);
}

// ---------------------------------------------------------------- single pass
// min-image displacement from the pilot, branch-free: d - L * rint(d / L).  Differs from the reference's
// vector_to (vector3d.rs:561-569) only in the last ulp and at the |d| = L/2 tie, which the extent check
// below excludes before the result is trusted.
// rint on the FMA pipe instead of FRND (which shares the XU pipe with the SFU trig): adding 1.5 * 2^23 rounds
// |q| < 2^22 to an integer, and that add rides on the FFMA that forms q = (x - p) / L.
// |q| >= 2^22 would defeat it; such a frame fails the extent check and is re-done by the exact passes.
__device__ __forceinline__ float pilot_delta(float x, float p, float L, float invL) {
    const float d = x - p;
    const float k = __fmaf_rn(d, invL, 12582912.0f) - 12582912.0f; // rint(d / L)
    return __fmaf_rn(-L, k, d);
}

constexpr double kExtentSlack = 1.0 - 1e-5; // unwrapped extent must be below (L/2) * slack on every axis
constexpr double kEdgeBand = 2e-5;          // c0 within this fraction of L of the box edge: image count ambiguous

// finishing thread of the single-pass centre: certify the pass and place the mean in the reference's image
template <bool WEIGHTED>
__device__ inline void finish_center(const double (&tot)[10], const float *tmn, const float *tmx, float px, float py, float pz,
                                     const float *L, uint32_t n, float *out3, int *flag) {
    const double M = WEIGHTED ? tot[3] : (double)n;
    const float p[3] = {px, py, pz};
    int redo = 0;
    for (int k = 0; k < 3; k++) {
        const double Lk = (double)L[k];
        if (!((double)tmx[k] - (double)tmn[k] < 0.5 * Lk * kExtentSlack)) redo = 1; // not compact: images may differ
        // circular mean of the group, iterators.rs:1152-1191, from the pilot-relative sums:
        // theta_i = theta_p + phi_i  =>  (xi, zeta) = R(theta_p) (C, S)
        const double C = tot[4 + k], S = tot[7 + k];
        if (!(C * C + S * S >= 1e-6 * (double)n * (double)n)) redo = 1; // resultant too short to trust the SFU sums
        // f32 library trig is plenty here (c0 only decides an integer, guarded by the edge band) and keeps this serial
        // section short: f64 sin/cos/atan2 cost thousands of cycles each in a single thread
        const float scf = 6.2831853f / L[k];
        const float pwf = p[k] - L[k] * floorf(p[k] / L[k]);
        float stf, ctf;
        sincosf(pwf * scf, &stf, &ctf);
        const double ct = (double)ctf, st = (double)stf;
        const double xi = ct * C - st * S, ze = st * C + ct * S;
        const double c0 = ((double)atan2f((float)(-ze), (float)(-xi)) + 3.141592653589793) / (double)scf; // in [0, L]
        if (!(c0 >= kEdgeBand * Lk && c0 <= (1.0 - kEdgeBand) * Lk)) redo = 1;  // which side of the edge decides k
        const double um = (double)p[k] + tot[k] / M;                           // mean of the unwrapped group
        out3[k] = (float)(um + Lk * rint((c0 - um) / Lk));                      // the image within L/2 of c0
    }
    *flag = redo;
}

// sums: [0..2] sum m*d, [3] sum m, [4..6] sum cos(phi), [7..9] sum sin(phi); min/max of d per axis
template <bool WEIGHTED>
__global__ void __launch_bounds__(kThreads) k_center_fast(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                           float *out, int *flags) {
    __shared__ FrameReduceSmem<10, 3> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    float px, py, pz;
    fv.load_atom(f, g.atom(0), px, py, pz);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const float sx = pi_x2() * ix, sy = pi_x2() * iy, sz = pi_x2() * iz;
    float a[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float dx = pilot_delta(x, px, L[0], ix), dy = pilot_delta(y, py, L[1], iy), dz = pilot_delta(z, pz, L[2], iz);
        if (WEIGHTED) {
            const float m = __ldg(g.mass + i);
            a[0] = __fmaf_rn(m, dx, a[0]); a[1] = __fmaf_rn(m, dy, a[1]); a[2] = __fmaf_rn(m, dz, a[2]);
            a[3] += m;
        } else {
            a[0] += dx; a[1] += dy; a[2] += dz;
        }
        float s, c;
        __sincosf(dx * sx, &s, &c); a[4] += c; a[7] += s;
        __sincosf(dy * sy, &s, &c); a[5] += c; a[8] += s;
        __sincosf(dz * sz, &s, &c); a[6] += c; a[9] += s;
        mn[0] = fminf(mn[0], dx); mx[0] = fmaxf(mx[0], dx);
        mn[1] = fminf(mn[1], dy); mx[1] = fmaxf(mx[1], dy);
        mn[2] = fminf(mn[2], dz); mx[2] = fmaxf(mx[2], dz);
    });
    double tot[10];
    float tmn[3], tmx[3];
    if (frame_reduce<10, 3>(a, mn, mx, partials + (size_t)f * nb * 16, tickets + f, nb, sm, tot, tmn, tmx) && threadIdx.x == 0) {
        finish_center<WEIGHTED>(tot, tmn, tmx, px, py, pz, L, g.n, out + f * 3, flags + f);
        if (fv.tric) fv.shear(f).to_x(out[f * 3], out[f * 3 + 1], out[f * 3 + 2]);  // the centre back from the sheared picture
    }
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

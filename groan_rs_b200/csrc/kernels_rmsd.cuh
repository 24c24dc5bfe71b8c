// kernels_rmsd.cuh -- RMSD with Kabsch fit over a batch of frames (reference-order passes).
//
//   k_ref_prepare  extract_data_from_system on the reference  rmsd.rs:425-446,479-492
//   k_cov          shift+wrap the target group, covariance sums, last CTA: SVD + RMSD   rmsd.rs:547-603
//   k_fit          fit_structure over ALL atoms              rmsd.rs:508-528
//
// RMSD is evaluated from sums (no second pass over rotated coordinates):
//   sum_i w_i |r^T pc_i - qc_i|^2 = sum w|pc|^2 + sum w|qc|^2 - 2 sum_ab r_ab Hw_ab,  Hw = sum_i w_i pc_i qc_i^T
// with H = sum_i pc_i qc_i^T (UNWEIGHTED, rmsd.rs:566-570) feeding the SVD.  The sums are accumulated in
// f64 (products of f32 pairs are exact there), so the cancellation in the identity is harmless.
#pragma once
#include "common.cuh"

namespace groan {

struct RefView {
    const float4 *pc; // group order: (y_ref - box_centre_ref).xyz, w = mass
    double sum_mpp;   // sum m |pc|^2
    double sum_m;     // sum m
    float com[3];     // reference.group_get_com(group) (rmsd.rs:133,198)
};

// reference side, once: y = wrap(x + (bc - com)); pc = y - bc; (pc, m) -> float4; sums m|pc|^2 and m
__global__ void __launch_bounds__(kThreads) k_ref_prepare(FrameView fv, GroupView g, const float *com, float4 *pc_out,
                                                           double *partials, unsigned int *tickets, double *sums_out) {
    __shared__ double smem[2 * (kThreads / 32)];
    __shared__ int sh_flag;
    const int nb = gridDim.x;
    float lx, ly, lz;
    fv.lengths(0, lx, ly, lz);
    const float bx = lx / 2.0f, by = ly / 2.0f, bz = lz / 2.0f; // get_box_center, mod.rs:298-308
    const float shx = bx - com[0], shy = by - com[1], shz = bz - com[2];
    const float *fr = fv.frame(0);
    double d[2] = {0, 0};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += nb * blockDim.x) {
        const float *p = fr + (size_t)g.atom(i) * 3;
        const float m = __ldg(g.mass + i);
        const float px = wrap_coordinate(__ldg(p + 0) + shx, lx) - bx;
        const float py = wrap_coordinate(__ldg(p + 1) + shy, ly) - by;
        const float pz = wrap_coordinate(__ldg(p + 2) + shz, lz) - bz;
        pc_out[i] = make_float4(px, py, pz, m);
        d[0] += (double)m * ((double)px * px + (double)py * py + (double)pz * pz);
        d[1] += (double)m;
    }
    block_sum<2>(d, smem);
    double tot[2];
    if (frame_finish<2>(d, partials, tickets, nb, tot, &sh_flag) && threadIdx.x == 0) {
        sums_out[0] = tot[0];
        sums_out[1] = tot[1];
    }
}

constexpr int kCovSums = 19; // H[9], Hw[9], sum m|qc|^2

__global__ void __launch_bounds__(kThreads) k_cov(FrameView fv, GroupView g, RefView ref, const float *com_in,
                                                   double *partials, unsigned int *tickets, float *rmsd_out, float *rot_out) {
    __shared__ double smem[kCovSums * (kThreads / 32)];
    __shared__ int sh_flag;
    const int f = blockIdx.y, nb = gridDim.x;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float bx = lx / 2.0f, by = ly / 2.0f, bz = lz / 2.0f;
    const float shx = bx - com_in[f * 3 + 0], shy = by - com_in[f * 3 + 1], shz = bz - com_in[f * 3 + 2];
    const float *fr = fv.frame(f);
    // f64 accumulation: products of two f32 are exact in f64, so the identity above stays accurate
    // down to rmsd ~ 1e-6 nm (an f32 partial would leave ~1e-4 nm at rmsd = 0).
    double d[kCovSums];
#pragma unroll
    for (int k = 0; k < kCovSums; k++) d[k] = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += nb * blockDim.x) {
        const float *p = fr + (size_t)g.atom(i) * 3;
        const float4 r = __ldg(ref.pc + i);
        // shift_and_wrap_coordinates (rmsd.rs:479-492) then q - centroid_q (rmsd.rs:564), f32 like the reference
        const double q[3] = {(double)(wrap_coordinate(__ldg(p + 0) + shx, lx) - bx),
                             (double)(wrap_coordinate(__ldg(p + 1) + shy, ly) - by),
                             (double)(wrap_coordinate(__ldg(p + 2) + shz, lz) - bz)};
        const double pc[3] = {(double)r.x, (double)r.y, (double)r.z};
        const double m = (double)r.w;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const double mp = m * pc[u];
#pragma unroll
            for (int v = 0; v < 3; v++) {
                d[u * 3 + v] = fma(pc[u], q[v], d[u * 3 + v]);
                d[9 + u * 3 + v] = fma(mp, q[v], d[9 + u * 3 + v]);
            }
        }
        d[18] = fma(m, fma(q[0], q[0], fma(q[1], q[1], q[2] * q[2])), d[18]);
    }
    block_sum<kCovSums>(d, smem);
    double tot[kCovSums];
    if (frame_finish<kCovSums>(d, partials + (size_t)f * nb * kCovSums, tickets + f, nb, tot, &sh_flag) && threadIdx.x == 0) {
        double r[9];
        kabsch_rotation(tot, r);
        double cross = 0.0;
        for (int k = 0; k < 9; k++) cross += r[k] * tot[9 + k];
        double msd = (ref.sum_mpp + tot[18] - 2.0 * cross) / ref.sum_m;
        if (msd < 0.0) msd = 0.0;
        rmsd_out[f] = (float)sqrt(msd);
        for (int k = 0; k < 9; k++) rot_out[f * 9 + k] = (float)r[k];
    }
}

// fit_structure, rmsd.rs:508-528, every atom of every frame:
//   translate(bc - com_tgt) incl. wrap (atom.rs:498-511); translate_nopbc(-bc); rotate_nopbc(r) = r * x
//   (vector3d.rs:359; nalgebra gemv order ((r0 x0) + r1 x1) + r2 x2); translate_nopbc(com_ref)
__global__ void __launch_bounds__(kThreads) k_fit(float *xyz, const float *box, size_t n_atoms, const float *com_tgt,
                                                   const float *rot, float rcx, float rcy, float rcz) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    const float bx = lx / 2.0f, by = ly / 2.0f, bz = lz / 2.0f;
    const float shx = bx - com_tgt[f * 3 + 0], shy = by - com_tgt[f * 3 + 1], shz = bz - com_tgt[f * 3 + 2];
    const float nbx = -bx, nby = -by, nbz = -bz;
    float r[9];
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = rot[f * 9 + k];
    float *fr = xyz + (size_t)f * n_atoms * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_atoms; i += (size_t)gridDim.x * blockDim.x) {
        float *p = fr + i * 3;
        const float vx = wrap_coordinate(p[0] + shx, lx) + nbx;
        const float vy = wrap_coordinate(p[1] + shy, ly) + nby;
        const float vz = wrap_coordinate(p[2] + shz, lz) + nbz;
        const float ox = (r[0] * vx + r[1] * vy) + r[2] * vz;
        const float oy = (r[3] * vx + r[4] * vy) + r[5] * vz;
        const float oz = (r[6] * vx + r[7] * vy) + r[8] * vz;
        p[0] = ox + rcx;
        p[1] = oy + rcy;
        p[2] = oz + rcz;
    }
}

// ---------------------------------------------------------------- wrap / translate (in place)
// MutAtomIteratorWithBox::wrap / translate, iterators.rs:1520,1548 -> atom.rs:498-545 -> vector3d.rs:380-417
template <bool TRANSLATE, bool SHIFTS>
__global__ void __launch_bounds__(kThreads) k_wrap(float *xyz, const float *box, size_t n_atoms, GroupView g, float tx,
                                                    float ty, float tz, int8_t *shifts) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    float *fr = xyz + (size_t)f * n_atoms * 3;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        float *p = fr + (size_t)g.atom(i) * 3;
        float x = p[0], y = p[1], z = p[2];
        if (TRANSLATE) { x += tx; y += ty; z += tz; }
        int kx, ky, kz;
        x = wrap_coordinate_count(x, lx, kx);
        y = wrap_coordinate_count(y, ly, ky);
        z = wrap_coordinate_count(z, lz, kz);
        p[0] = x; p[1] = y; p[2] = z;
        if (SHIFTS) {
            int8_t *s = shifts + ((size_t)f * g.n + i) * 3;
            s[0] = (int8_t)kx; s[1] = (int8_t)ky; s[2] = (int8_t)kz;
        }
    }
}

// triclinic EXTENSION (no reference counterpart; definition in DESIGN.md, oracle orc_tric_wrap1):
// z, then y, then x, removing whole box vectors with the reference's strict loop comparisons.
template <bool TRANSLATE, bool SHIFTS>
__global__ void __launch_bounds__(kThreads) k_wrap_tric(float *xyz, const float *box, size_t n_atoms, GroupView g, float tx,
                                                         float ty, float tz, int8_t *shifts) {
    const int f = blockIdx.y;
    float B[9];
#pragma unroll
    for (int k = 0; k < 9; k++) B[k] = __ldg(box + f * 9 + k);
    float *fr = xyz + (size_t)f * n_atoms * 3;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        float *p = fr + (size_t)g.atom(i) * 3;
        float x = p[0], y = p[1], z = p[2];
        if (TRANSLATE) { x += tx; y += ty; z += tz; }
        int kx = 0, ky = 0, kz = 0;
        while (z > B[8]) { x -= B[6]; y -= B[7]; z -= B[8]; kz--; }
        while (z < 0.0f) { x += B[6]; y += B[7]; z += B[8]; kz++; }
        while (y > B[4]) { x -= B[3]; y -= B[4]; ky--; }
        while (y < 0.0f) { x += B[3]; y += B[4]; ky++; }
        while (x > B[0]) { x -= B[0]; kx--; }
        while (x < 0.0f) { x += B[0]; kx++; }
        p[0] = x; p[1] = y; p[2] = z;
        if (SHIFTS) {
            int8_t *s = shifts + ((size_t)f * g.n + i) * 3;
            s[0] = (int8_t)kx; s[1] = (int8_t)ky; s[2] = (int8_t)kz;
        }
    }
}

} // namespace groan

// kernels_rmsd.cuh -- RMSD with Kabsch fit over a batch of frames, plus wrap / translate.
//
//   k_ref_prepare  extract_data_from_system on the reference      rmsd.rs:425-446,479-492
//   k_cov          exact pass: shift+wrap the target group, covariance sums in f64, SVD + RMSD   rmsd.rs:547-603
//   k_rmsd_fast    single pass: COM, covariance and RMSD sums relative to a pilot atom (DESIGN.md "Single-pass RMSD")
//   k_fit          fit_structure over ALL atoms                   rmsd.rs:508-528
//   k_wrap         atoms_wrap / atoms_translate                   modifying.rs:73,201; vector3d.rs:380-417
//
// RMSD is evaluated from sums (no second pass over rotated coordinates):
//   sum_i w_i |r^T pc_i - qc_i|^2 = sum w|pc|^2 + sum w|qc|^2 - 2 sum_ab r_ab Hw_ab,  Hw = sum_i w_i pc_i qc_i^T
// with H = sum_i pc_i qc_i^T (UNWEIGHTED, rmsd.rs:566-570) feeding the SVD.
#pragma once
#include "common.cuh"
#include "kernels_center.cuh"

namespace groan {

// The prepared reference lives in blocks of kRefBlock atoms, structure-of-arrays inside a block:
// [pc.x * 256][pc.y * 256][pc.z * 256][w * 256] with pc = y_ref - box_centre_ref and w the reference mass.
// A bulk copy of whole blocks therefore lands in shared memory in a layout that consecutive lanes read
// without bank conflicts and straight into the register pairs of the packed f32x2 arithmetic.
constexpr int kRefBlock = 256;
__host__ __device__ inline size_t ref_floats(size_t n) { return ((n + kRefBlock - 1) / kRefBlock) * (size_t)(4 * kRefBlock); }
__device__ __forceinline__ size_t ref_word(uint32_t i) { return (size_t)(i >> 8) * (4 * kRefBlock) + (i & (kRefBlock - 1)); }
__device__ __forceinline__ float4 ref_at(const float *ref, uint32_t i) {
    const float *b = ref + ref_word(i);
    return make_float4(__ldg(b), __ldg(b + kRefBlock), __ldg(b + 2 * kRefBlock), __ldg(b + 3 * kRefBlock));
}

struct RefView {
    const float *pc;   // block-SoA (see above), group order
    double sum_wpp;    // sum w |pc|^2
    double sum_w;      // sum w
    double sum_pc[3];  // sum pc
    double sum_wpc[3]; // sum w pc
    double sum_m_target; // sum of the TARGET group's own masses (constant over the frames; used by the quad kernels)
    float com[3];      // reference.group_get_com(group) (rmsd.rs:133,198)
};

constexpr int kRefSums = 8; // w|pc|^2, w, pc[3], w pc[3]

// reference side, once: y = wrap(x + (bc - com)); pc = y - bc; (pc, w) -> float4; constant sums
__global__ void __launch_bounds__(kThreads) k_ref_prepare(FrameView fv, GroupView g, const float *com, float *pc_out,
                                                           double *partials, unsigned int *tickets, double *sums_out) {
    __shared__ FrameReduceSmem<kRefSums, 0> sm;
    const int nb = gridDim.x;
    float lx, ly, lz;
    fv.lengths(0, lx, ly, lz);
    const float bx = lx / 2.0f, by = ly / 2.0f, bz = lz / 2.0f; // get_box_center, mod.rs:298-308
    const float shx = bx - com[0], shy = by - com[1], shz = bz - com[2];
    double d[kRefSums] = {0, 0, 0, 0, 0, 0, 0, 0};
    const Shear shr = fv.shear(0);
    for_each_group_atom(fv, g, 0, [&](uint32_t i, float x, float y, float z) {
        const float m = __ldg(g.mass + i);
        float px = wrap_coordinate(x + shx, lx) - bx;
        float py = wrap_coordinate(y + shy, ly) - by;
        float pz = wrap_coordinate(z + shz, lz) - bz;
        if (fv.tric) shr.to_x(px, py, pz);  // wrapped in the sheared picture, stored (and rotated) in Cartesian coordinates
        float *o = pc_out + ref_word(i);
        o[0] = px; o[kRefBlock] = py; o[2 * kRefBlock] = pz; o[3 * kRefBlock] = m;
        d[0] += (double)m * ((double)px * px + (double)py * py + (double)pz * pz);
        d[1] += (double)m;
        d[2] += (double)px; d[3] += (double)py; d[4] += (double)pz;
        d[5] += (double)m * px; d[6] += (double)m * py; d[7] += (double)m * pz;
    });
    double tot[kRefSums];
    if (frame_reduce<kRefSums, 0>(d, nullptr, nullptr, partials, tickets, nb, sm, tot, nullptr, nullptr) && threadIdx.x == 0)
        for (int k = 0; k < kRefSums; k++) sums_out[k] = tot[k];
}

// Kabsch + RMSD from the covariance sums (thread 0 of the finishing CTA), rmsd.rs:573-599
__device__ inline double finish_kabsch(const double H[9], const double Hw[9], double sum_wqq, const RefView &ref, double r[9],
                                       double *ratio) {
    kabsch_rotation(H, r);
    double cross = 0.0;
    for (int k = 0; k < 9; k++) cross += r[k] * Hw[k];
    const double T = ref.sum_wpp + sum_wqq;
    double R = T - 2.0 * cross;
    if (ratio) *ratio = (T > 0.0) ? R * fast_rcp(T) : 1.0;
    if (R < 0.0) R = 0.0;
    return fast_sqrt(R * fast_rcp(ref.sum_w));
}

constexpr int kCovSums = 19; // H[9], Hw[9], sum w|qc|^2

__global__ void __launch_bounds__(kThreads) k_cov(FrameView fv, GroupView g, RefView ref, const float *com_in, double *partials,
                                                   unsigned int *tickets, float *rmsd_out, float *rot_out, const int *flags) {
    __shared__ FrameReduceSmem<kCovSums, 0> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    if (frame_skipped(flags, f)) return;
    float lx, ly, lz;
    fv.lengths(f, lx, ly, lz);
    const float bx = lx / 2.0f, by = ly / 2.0f, bz = lz / 2.0f;
    const float shx = bx - com_in[f * 3 + 0], shy = by - com_in[f * 3 + 1], shz = bz - com_in[f * 3 + 2];
    // f64 accumulation: products of two f32 are exact in f64, so the identity above stays accurate
    // down to rmsd ~ 1e-6 nm (an f32 partial would leave ~1e-4 nm at rmsd = 0 for a small group).
    double d[kCovSums];
#pragma unroll
    for (int k = 0; k < kCovSums; k++) d[k] = 0.0;
    const Shear shr = fv.shear(f);
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float4 r = ref_at(ref.pc, i);
        // shift_and_wrap_coordinates (rmsd.rs:479-492) then q - centroid_q (rmsd.rs:564), f32 like the reference
        float qx = wrap_coordinate(x + shx, lx) - bx, qy = wrap_coordinate(y + shy, ly) - by, qz = wrap_coordinate(z + shz, lz) - bz;
        if (fv.tric) shr.to_x(qx, qy, qz);
        const double q[3] = {(double)qx, (double)qy, (double)qz};
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
    });
    double tot[kCovSums];
    if (frame_reduce<kCovSums, 0>(d, nullptr, nullptr, partials + (size_t)f * nb * kCovSums, tickets + f, nb, sm, tot, nullptr,
                                  nullptr) &&
        threadIdx.x == 0) {
        double r[9];
        rmsd_out[f] = (float)finish_kabsch(tot, tot + 9, tot[18], ref, r, nullptr);
        for (int k = 0; k < 9; k++) rot_out[f * 9 + k] = (float)r[k];
    }
}

// ---------------------------------------------------------------- single pass
// With d_i the min-image displacement of atom i from the pilot p (the group's first atom), u_i = p + d_i is
// the group made whole, com = p + delta with delta = sum m d / sum m, and -- as long as the whole group is
// shorter than half the box (checked) -- the reference's wrap(x + (bc - com)) - bc equals d_i - delta.  Hence
//   H  = sum pc d^T   - (sum pc)   delta^T          Hw = sum w pc d^T - (sum w pc) delta^T
//   sum w|qc|^2 = sum w|d|^2 - 2 delta . sum w d + |delta|^2 sum w
// and one pass over the frame suffices.  Sums are f32 FFMA per thread, f64 across threads; the finishing
// thread flags the frame for the exact f64 passes when the cancellation in the RMSD identity has eaten more
// than the f32 products can give (ratio test) or the group is not compact.
//
// sums: [0..8] sum pc_a d_b, [9..17] sum w pc_a d_b, [18..20] sum w d, [21] sum w|d|^2, [22..24] sum m d, [25] sum m
constexpr int kFastSums = 26;
constexpr double kCancelGuard = 2e-5;

// d: displacement the sums are formed with (Cartesian); du: the same displacement in the picture the box is orthogonal in
// (identical unless the triclinic extension is on), whose extent certifies the pass
template <bool SAME_MASS>
__device__ __forceinline__ void rmsd_accumulate(float (&a)[kFastSums], float (&mn)[3], float (&mx)[3], const float (&d)[3],
                                                const float4 &r, float m, const float (&du)[3]) {
    const float pc[3] = {r.x, r.y, r.z};
    const float w = r.w;
#pragma unroll
    for (int u = 0; u < 3; u++) {
        const float wp = w * pc[u];
#pragma unroll
        for (int v = 0; v < 3; v++) {
            a[u * 3 + v] = __fmaf_rn(pc[u], d[v], a[u * 3 + v]);
            a[9 + u * 3 + v] = __fmaf_rn(wp, d[v], a[9 + u * 3 + v]);
        }
    }
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const float wd = w * d[v];
        a[18 + v] += wd;
        a[21] = __fmaf_rn(wd, d[v], a[21]);
        mn[v] = fminf(mn[v], du[v]);
        mx[v] = fmaxf(mx[v], du[v]);
    }
    if (!SAME_MASS) {
#pragma unroll
        for (int v = 0; v < 3; v++) a[22 + v] = __fmaf_rn(m, d[v], a[22 + v]);
        a[25] += m;
    }
}

template <bool SAME_MASS>
__device__ inline void finish_rmsd(const double (&tot)[kFastSums], const float *tmn, const float *tmx, float px, float py, float pz,
                                   const float *L, const RefView &ref, float *rmsd_out, float *rot9, float *com3, int *flag) {
    int redo = 0;
    double delta[3];
    const double inv_m = fast_rcp(SAME_MASS ? ref.sum_w : tot[25]);
    for (int k = 0; k < 3; k++) {
        if (!((double)tmx[k] - (double)tmn[k] < 0.5 * (double)L[k] * kExtentSlack)) redo = 1;
        delta[k] = (SAME_MASS ? tot[18 + k] : tot[22 + k]) * inv_m;
    }
    double H[9], Hw[9];
    for (int u = 0; u < 3; u++)
        for (int v = 0; v < 3; v++) {
            H[u * 3 + v] = tot[u * 3 + v] - ref.sum_pc[u] * delta[v];
            Hw[u * 3 + v] = tot[9 + u * 3 + v] - ref.sum_wpc[u] * delta[v];
        }
    const double dd = delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2];
    const double wqq = tot[21] - 2.0 * (delta[0] * tot[18] + delta[1] * tot[19] + delta[2] * tot[20]) + dd * ref.sum_w;
    double r[9], ratio;
    *rmsd_out = (float)finish_kabsch(H, Hw, wqq, ref, r, &ratio);
    if (!(ratio >= kCancelGuard)) redo = 1; // f32 products cannot resolve this RMSD: exact f64 passes
    for (int k = 0; k < 9; k++) rot9[k] = (float)r[k];
    com3[0] = (float)((double)px + delta[0]);
    com3[1] = (float)((double)py + delta[1]);
    com3[2] = (float)((double)pz + delta[2]);
    *flag = redo;
}

template <bool SAME_MASS>
__global__ void __launch_bounds__(kThreads, 2) k_rmsd_fast(FrameView fv, GroupView g, RefView ref, double *partials,
                                                            unsigned int *tickets, float *rmsd_out, float *rot_out, float *com_out,
                                                            int *flags) {
    __shared__ FrameReduceSmem<kFastSums, 3> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    float px, py, pz;
    fv.load_atom(f, g.atom(0), px, py, pz);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    float a[kFastSums];
#pragma unroll
    for (int k = 0; k < kFastSums; k++) a[k] = 0.0f;
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    const Shear shr = fv.shear(f);
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float4 r = ref_at(ref.pc, i);
        const float du[3] = {pilot_delta(x, px, L[0], ix), pilot_delta(y, py, L[1], iy), pilot_delta(z, pz, L[2], iz)};
        float d[3] = {du[0], du[1], du[2]};
        if (fv.tric) shr.to_x(d[0], d[1], d[2]);  // linear map: sums of the Cartesian displacements are what Kabsch needs
        rmsd_accumulate<SAME_MASS>(a, mn, mx, d, r, SAME_MASS ? 0.0f : __ldg(g.mass + i), du);
    });
    double tot[kFastSums];
    float tmn[3], tmx[3];
    if (frame_reduce<kFastSums, 3>(a, mn, mx, partials + (size_t)f * nb * (kFastSums + 6), tickets + f, nb, sm, tot, tmn, tmx) &&
        threadIdx.x == 0)
        finish_rmsd<SAME_MASS>(tot, tmn, tmx, px, py, pz, L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, flags + f);
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
    for_each_atom_inplace(xyz, n_atoms, f, 0u, (uint32_t)n_atoms, [&](uint32_t, float &x, float &y, float &z) {
        const float vx = wrap_coordinate(x + shx, lx) + nbx;
        const float vy = wrap_coordinate(y + shy, ly) + nby;
        const float vz = wrap_coordinate(z + shz, lz) + nbz;
        const float ox = (r[0] * vx + r[1] * vy) + r[2] * vz;
        const float oy = (r[3] * vx + r[4] * vy) + r[5] * vz;
        const float oz = (r[6] * vx + r[7] * vy) + r[8] * vz;
        x = ox + rcx;
        y = oy + rcy;
        z = oz + rcz;
    });
}

// ---------------------------------------------------------------- wrap / translate (in place)
// MutAtomIteratorWithBox::wrap / translate, iterators.rs:1520,1548 -> atom.rs:498-545 -> vector3d.rs:380-417
template <bool TRANSLATE, bool SHIFTS>
__global__ void __launch_bounds__(kThreads) k_wrap(float *xyz, const float *box, size_t n_atoms, GroupView g, float tx,
                                                    float ty, float tz, int8_t *shifts) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    float *fr = xyz + (size_t)f * n_atoms * 3;
    if (!SHIFTS && !g.idx) { // contiguous range, positions only: quads of atoms, 128-bit loads and stores
        for_each_atom_inplace(xyz, n_atoms, f, g.first, g.n, [&](uint32_t, float &x, float &y, float &z) {
            if (TRANSLATE) { x += tx; y += ty; z += tz; }
            x = wrap_coordinate(x, lx);
            y = wrap_coordinate(y, ly);
            z = wrap_coordinate(z, lz);
        });
        return;
    }
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

// ---------------------------------------------------------------- whole groups / molecules, centering (in place)
// System::make_group_whole, modifying.rs:437-465: pos = c + vector_to(c, pos) with c = group_estimate_center of the frame
__global__ void __launch_bounds__(kThreads) k_make_whole(float *xyz, const float *box, size_t n_atoms, GroupView g, const float *c0) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    const float cx = c0[f * 3 + 0], cy = c0[f * 3 + 1], cz = c0[f * 3 + 2];
    float *fr = xyz + (size_t)f * n_atoms * 3;
    if (!g.idx) {
        for_each_atom_inplace(xyz, n_atoms, f, g.first, g.n, [&](uint32_t, float &x, float &y, float &z) {
            x = cx + vector_to_1(cx, x, lx);
            y = cy + vector_to_1(cy, y, ly);
            z = cz + vector_to_1(cz, z, lz);
        });
        return;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        float *p = fr + (size_t)g.atom(i) * 3;
        const float x = p[0], y = p[1], z = p[2];
        p[0] = cx + vector_to_1(cx, x, lx);
        p[1] = cy + vector_to_1(cy, y, ly);
        p[2] = cz + vector_to_1(cz, z, lz);
    }
}

// System::make_molecules_whole, modifying.rs:338-391.  mol_ref[i] = reference atom (lowest index) of atom i's molecule,
// kNoMolecule for atoms of monoatomic molecules (untouched).  The reference atom is wrapped into the box; every other atom
// goes to ref + vector_to(ref, pos).  A thread reads its reference atom while that atom's own thread may be storing the
// wrapped value: wrap is idempotent per component (the result lies in [0, L]), so wrapping whatever is read gives the same.
constexpr uint32_t kNoMolecule = 0xFFFFFFFFu;
__global__ void __launch_bounds__(kThreads) k_mol_whole(float *xyz, const float *box, size_t n_atoms, const uint32_t *mol_ref) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    float *fr = xyz + (size_t)f * n_atoms * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_atoms; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = __ldg(mol_ref + i);
        if (r == kNoMolecule) continue;
        const volatile float *q = fr + (size_t)r * 3;
        const float rx = wrap_coordinate(q[0], lx), ry = wrap_coordinate(q[1], ly), rz = wrap_coordinate(q[2], lz);
        float *p = fr + i * 3;
        if (r == (uint32_t)i) {
            p[0] = rx; p[1] = ry; p[2] = rz;
        } else {
            const float x = p[0], y = p[1], z = p[2];
            p[0] = rx + vector_to_1(rx, x, lx);
            p[1] = ry + vector_to_1(ry, y, ly);
            p[2] = rz + vector_to_1(rz, z, lz);
        }
    }
}

// System::atoms_center / atoms_center_mass, utility.rs:109-130,168-189: shift = box_centre - estimate (f32), components
// outside `dim` zeroed (Vector3D::filter), then atoms_translate: pos += shift; wrap (atom.rs:498-511)
__global__ void __launch_bounds__(kThreads) k_center_atoms(float *xyz, const float *box, size_t n_atoms, const float *c0, int dim) {
    const int f = blockIdx.y;
    const float lx = __ldg(box + f * 9), ly = __ldg(box + f * 9 + 4), lz = __ldg(box + f * 9 + 8);
    const bool kx = dim == 1 || dim == 4 || dim == 5 || dim == 7, ky = dim == 2 || dim == 4 || dim == 6 || dim == 7,
               kz = dim == 3 || dim == 5 || dim == 6 || dim == 7;
    const float sx = kx ? lx / 2.0f - c0[f * 3 + 0] : 0.0f, sy = ky ? ly / 2.0f - c0[f * 3 + 1] : 0.0f,
                sz = kz ? lz / 2.0f - c0[f * 3 + 2] : 0.0f;
    for_each_atom_inplace(xyz, n_atoms, f, 0u, (uint32_t)n_atoms, [&](uint32_t, float &x, float &y, float &z) {
        x = wrap_coordinate(x + sx, lx);
        y = wrap_coordinate(y + sy, ly);
        z = wrap_coordinate(z + sz, lz);
    });
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

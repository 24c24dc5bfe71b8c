// kernels_bonds.cuh -- the two users of the cell grid in the reference, as consumers of the same neighbour walk
// (kernels_cells.cuh): bond guessing (System::guess_bonds, src/system/guess.rs:362-470) and the donor / acceptor search of
// HBondAnalysis (src/system/hbonds.rs:160-335).  One warp per query atom, the candidates of the (up to 27) cells around it
// in the lanes, Vector3D::distance with the reference's arithmetic on the original coordinates; what differs from the plain
// cutoff search is the test a candidate has to pass and what is written for it.
#pragma once
#include "kernels_cells.cuh"

namespace groan {

// walk the cells around (ax, ay, az): fn(candidate float4 (x, y, z, position in the binned group), d2, valid) per lane and step
template <typename F>
__device__ __forceinline__ void for_each_neighbour(float ax, float ay, float az, const BoxOrtho &B, const CellGeom &cg, const uint32_t *of,
                                                   const float4 *sb, bool fold, int lane, F &&fn) {
    // The binned atoms are sorted by cell index with x running fastest, so the (up to three) neighbour cells of one (y, z) row
    // are ONE contiguous range of the sorted array -- plus the cell at the other end of the row when the atom's cell is the first
    // or the last one (periodic wrap).  9 ranges (+ up to 9 wrap cells) instead of 27: with the small cells of bond guessing
    // (0.17 nm: half an atom per cell) most of a walk is range bookkeeping.
    int ys[3], zs[3], my, mz;
    const int cx = cell_coord(ax, B.lx, cg.nx);
    axis_cells(cell_coord(ay, B.ly, cg.ny), cg.ny, ys, my);
    axis_cells(cell_coord(az, B.lz, cg.nz), cg.nz, zs, mz);
    const int x_lo = cg.nx >= 3 ? max(cx - 1, 0) : 0, x_hi = cg.nx >= 3 ? min(cx + 1, cg.nx - 1) : cg.nx - 1;
    const int x_wrap = cg.nx >= 3 ? (cx == 0 ? cg.nx - 1 : (cx == cg.nx - 1 ? 0 : -1)) : -1;
    const int rows = my * mz, ncell = x_wrap >= 0 ? 2 * rows : rows;
    uint32_t my_lo = 0, my_hi = 0;
    if (lane < ncell) {
        const int r = lane < rows ? lane : lane - rows, ky = r % my, kz = r / my;
        const uint32_t row = ((uint32_t)zs[kz] * cg.ny + ys[ky]) * cg.nx;
        if (lane < rows) {
            my_lo = of[row + x_lo];
            my_hi = of[row + x_hi + 1];
        } else {
            my_lo = of[row + x_wrap];
            my_hi = of[row + x_wrap + 1];
        }
    }
    for (int k = 0; k < ncell; k++) {
        const uint32_t lo = __shfl_sync(0xffffffffu, my_lo, k), hi = __shfl_sync(0xffffffffu, my_hi, k);
        for (uint32_t s0 = lo; s0 < hi; s0 += 32) {
            const uint32_t s = s0 + lane;
            const float4 b = s < hi ? sb[s] : make_float4(0.f, 0.f, 0.f, 0.f);
            float dx, dy, dz;
            if (fold) { // both atoms within L/4 of the box: |min_image(d)| = min(|d|, ||d| - L|) exactly
                const float rx = fabsf(ax - b.x), ry = fabsf(ay - b.y), rz = fabsf(az - b.z);
                dx = fminf(rx, fabsf(rx - B.lx));
                dy = fminf(ry, fabsf(ry - B.ly));
                dz = fminf(rz, fabsf(rz - B.lz));
            } else {
                dx = min_image(ax - b.x, B.lx);
                dy = min_image(ay - b.y, B.ly);
                dz = min_image(az - b.z, B.lz);
            }
            fn(b, (dx * dx + dy * dy) + dz * dz, s < hi);
        }
    }
}

// warp-aggregated append: lanes with `want` get consecutive slots of the frame's list (one atomic per warp and call)
__device__ __forceinline__ unsigned long long warp_append(bool want, unsigned long long *cursor, int lane, unsigned long long &mine) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (m == 0u) return ~0ull;
    unsigned long long base = 0;
    if (lane == 0) {
        base = atomicAdd(cursor, (unsigned long long)__popc(m));
        mine += __popc(m);
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------- guess_bonds (guess.rs:427-470)
// atom1 over every atom with a van der Waals radius, atom2 over its neighbours: a bond if distance < (vdw1 + vdw2) * factor.
// The reference inserts (min, max) into a set; here a pair is reported once, from its lower index.  `sorted` bins ALL atoms.
__global__ void __launch_bounds__(kThreads) k_guess_bonds(FrameView fv, uint32_t n_atoms, CellGeom cg, const uint32_t *offsets, const float4 *sorted,
                                                           size_t cells, const float *vdw, float factor, unsigned long long *count,
                                                           uint32_t *pairs, unsigned long long capacity, unsigned long long *cursor,
                                                           const unsigned int *far) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    BoxOrtho B;
    load_box(fv.box, f, B);
    const float *fr = fv.frame(f);
    const uint32_t *of = offsets + (size_t)f * (cells + 1);
    const float4 *sb = sorted + (size_t)f * n_atoms;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const bool b_near = far[f] == 0u;
    unsigned long long mine = 0;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_atoms; i += warps) {
        const float v1 = __ldg(vdw + i);
        if (!(v1 >= 0.0f)) continue; // no radius: the reference lists the atom in its warning and skips it
        const float *p = fr + (size_t)i * 3;
        const float ax = __ldg(p), ay = __ldg(p + 1), az = __ldg(p + 2);
        const bool fold = b_near && atom_in_fold_range<7>(ax, ay, az, B);
        for_each_neighbour(ax, ay, az, B, cg, of, sb, fold, lane, [&](const float4 &b, float d2, bool valid) {
            const uint32_t j = __float_as_uint(b.w);
            bool hit = false;
            if (valid && j > i) {
                const float v2 = __ldg(vdw + j);
                hit = v2 >= 0.0f && sqrt1_rn(d2) < (v1 + v2) * factor; // Atom::distance < limit, both in f32
            }
            const unsigned long long at = warp_append(hit, cursor + f, lane, mine);
            if (hit && pairs && at < capacity) {
                uint32_t *o = pairs + ((size_t)f * capacity + at) * 2;
                o[0] = i;
                o[1] = j;
            }
        });
    }
    if (lane == 0 && mine) atomicAdd(count + f, mine);
}

// ---------------------------------------------------------------- HBondAnalysis::analyze_single (hbonds.rs:240-320)
// donors[d] with hydrogens hyd[hyd_off[d] .. hyd_off[d + 1]); `sorted` bins the acceptor group `gacc`.  For every acceptor other
// than the donor itself with distance(acceptor, donor) <= max_distance (d2 < le2, the smallest float whose root exceeds it) and
// every hydrogen of the donor with angle(donor - hydrogen - acceptor) >= min_angle: one record (donor, hydrogen, acceptor) +
// (distance, angle in degrees).  calc_angle (hbonds.rs:322-335): vector_to from the hydrogen to either atom, acos of the
// normalised dot product, NaN resolved by the two distances.
__device__ __forceinline__ float hbond_angle(const float *h, const float *d, const float *a, const BoxOrtho &B) {
    const float L[3] = {B.lx, B.ly, B.lz};
    float hd[3], ha[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        hd[k] = vector_to_1(h[k], d[k], L[k]);
        ha[k] = vector_to_1(h[k], a[k], L[k]);
    }
    const float dot = (hd[0] * ha[0] + hd[1] * ha[1]) + hd[2] * ha[2];
    const float l1 = sqrtf((hd[0] * hd[0] + hd[1] * hd[1]) + hd[2] * hd[2]), l2 = sqrtf((ha[0] * ha[0] + ha[1] * ha[1]) + ha[2] * ha[2]);
    const float ang = acosf(dot / (l1 * l2)) * 57.29577951308232f; // f32::to_degrees
    if (ang == ang) return ang;
    // handle_nan: hydrogen closer to the acceptor than the donor is -> 180, else 0
    float q1 = 0.f, q2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float u = min_image(h[k] - a[k], L[k]), v = min_image(d[k] - a[k], L[k]);
        q1 = k == 0 ? u * u : q1 + u * u;
        q2 = k == 0 ? v * v : q2 + v * v;
    }
    return sqrtf(q1) < sqrtf(q2) ? 180.0f : 0.0f;
}

__global__ void __launch_bounds__(kThreads) k_hbonds(FrameView fv, GroupView gacc, CellGeom cg, const uint32_t *offsets, const float4 *sorted,
                                                      size_t cells, const uint32_t *donors, const uint32_t *hyd_off, const uint32_t *hyd,
                                                      uint32_t n_donors, float le2, float min_angle, unsigned long long *count, uint32_t *dha,
                                                      float *dist_angle, unsigned long long capacity, unsigned long long *cursor,
                                                      const unsigned int *far) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    BoxOrtho B;
    load_box(fv.box, f, B);
    const float *fr = fv.frame(f);
    const uint32_t *of = offsets + (size_t)f * (cells + 1);
    const float4 *sb = sorted + (size_t)f * gacc.n;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const bool b_near = far[f] == 0u;
    unsigned long long mine = 0;
    for (uint32_t di = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); di < n_donors; di += warps) {
        const uint32_t don = __ldg(donors + di), h0 = __ldg(hyd_off + di), h1 = __ldg(hyd_off + di + 1);
        const float *p = fr + (size_t)don * 3;
        const float dpos[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        const bool fold = b_near && atom_in_fold_range<7>(dpos[0], dpos[1], dpos[2], B);
        for_each_neighbour(dpos[0], dpos[1], dpos[2], B, cg, of, sb, fold, lane, [&](const float4 &b, float d2, bool valid) {
            const uint32_t acc = valid ? gacc.atom(__float_as_uint(b.w)) : 0u;
            const bool near = valid && acc != don && d2 < le2;
            if (__ballot_sync(0xffffffffu, near) == 0u) return;
            const float apos[3] = {b.x, b.y, b.z};
            const float dist = sqrt1_rn(d2);
            for (uint32_t h = h0; h < h1; h++) { // warp-uniform trip count: the donor's hydrogens
                const uint32_t hy = __ldg(hyd + h);
                const float *q = fr + (size_t)hy * 3;
                const float hpos[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
                float ang = 0.0f;
                bool hit = false;
                if (near) {
                    ang = hbond_angle(hpos, dpos, apos, B);
                    hit = !(ang < min_angle);
                }
                const unsigned long long at = warp_append(hit, cursor + f, lane, mine);
                if (hit && dha && at < capacity) {
                    uint32_t *o = dha + ((size_t)f * capacity + at) * 3;
                    o[0] = don;
                    o[1] = hy;
                    o[2] = acc;
                    float *r = dist_angle + ((size_t)f * capacity + at) * 2;
                    r[0] = dist;
                    r[1] = ang;
                }
            }
        });
    }
    if (lane == 0 && mine) atomicAdd(count + f, mine);
}

} // namespace groan

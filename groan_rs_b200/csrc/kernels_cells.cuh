// kernels_cells.cuh -- cutoff pair search through a cell grid: the sparse counterpart of the all-pairs kernels.
//
// Reference: CellGrid::new (cellgrid.rs:301-381) bins the atoms of a group into cells, CellGrid::neighbors_iter (:383-420)
// walks the cells around a reference point, and the users keep the atoms within a cutoff (guess.rs:362-470 bond guessing,
// hbonds.rs:240-335 donor/acceptor search).  Here, per frame:
//
//   k_cell_count   cell of every atom of group B (from its wrapped position) + histogram
//   k_cell_scan    exclusive prefix sum of the histogram (one CTA per frame)
//   k_cell_fill    counting sort: (x, y, z, position in B) of every atom, cell by cell
//   k_cell_query   one warp per atom of group A: the 27 cells around it (fewer when the grid is narrower than 3 cells, so
//                  that no cell is visited twice: NeighborsRange::convert, cellgrid.rs:207-231), lanes over the atoms of a
//                  cell, Vector3D::distance with the reference's arithmetic on the ORIGINAL coordinates (bit-identical to
//                  the all-pairs matrix), pairs below the cutoff counted and, if asked for, appended to the frame's list
//
// Cells are at least cutoff * (1 + 1e-4) wide: two atoms closer than the cutoff then sit in the same or in adjacent cells
// even if the f32 cell assignment of either is off by a rounding error at a cell boundary.
// The order of the emitted pairs is undefined, as is the order of neighbors_iter (cellgrid.rs:141-144).
#pragma once
#include "common.cuh"
#include "kernels_pairs.cuh"

namespace groan {

struct CellGeom {
    int nx, ny, nz; // cells per axis
};

__device__ __forceinline__ int cell_coord(float x, float L, int n) {
    const float w = wrap_coordinate(x, L); // in [0, L]
    int c = (int)(w * ((float)n / L));
    return c < 0 ? 0 : (c >= n ? n - 1 : c); // w == L (wrap keeps x == L) and rounding at the upper face
}
__device__ __forceinline__ uint32_t cell_index(float x, float y, float z, const BoxOrtho &B, const CellGeom &cg) {
    const int cx = cell_coord(x, B.lx, cg.nx), cy = cell_coord(y, B.ly, cg.ny), cz = cell_coord(z, B.lz, cg.nz);
    return ((uint32_t)cz * cg.ny + cy) * cg.nx + cx;
}

// cells per axis for a frame: the grid differs from frame to frame only through the box, and all frames share one
// geometry chosen on the host from the SMALLEST box of the batch (cells can only get wider for the other frames)
// far[f] is set when an atom of group B lies more than L/4 outside the box: the query then uses the reference's loop
// form of the minimum image instead of the branch-free one-step fold (kernels_pairs.cuh)
__global__ void __launch_bounds__(kThreads) k_cell_count(FrameView fv, GroupView gb, CellGeom cg, uint32_t *cell_of, uint32_t *counts,
                                                          size_t cells, unsigned int *far) {
    const int f = blockIdx.y;
    BoxOrtho B;
    load_box(fv.box, f, B);
    const float *fr = fv.frame(f);
    bool ok = true;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < gb.n; j += gridDim.x * blockDim.x) {
        const float *p = fr + (size_t)gb.atom(j) * 3;
        const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
        const uint32_t c = cell_index(x, y, z, B, cg);
        ok = ok && atom_in_fold_range<7>(x, y, z, B);
        cell_of[(size_t)f * gb.n + j] = c;
        atomicAdd(counts + (size_t)f * cells + c, 1u);
    }
    if (!ok) atomicOr(far + f, 1u);
}

// offsets[c] = number of atoms in cells < c; cursor[c] = the same (k_cell_fill advances it).  One CTA per frame.
__global__ void __launch_bounds__(1024) k_cell_scan(const uint32_t *counts, uint32_t *offsets, uint32_t *cursor, size_t cells) {
    __shared__ uint32_t part[1024];
    const int f = blockIdx.x;
    const uint32_t *cn = counts + (size_t)f * cells;
    uint32_t *of = offsets + (size_t)f * (cells + 1), *cu = cursor + (size_t)f * cells;
    const size_t per = (cells + 1023) / 1024, lo = threadIdx.x * per, hi = lo + per < cells ? lo + per : cells;
    uint32_t sum = 0;
    for (size_t c = lo; c < hi; c++) sum += cn[c];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { // Hillis-Steele inclusive scan of the 1024 partial sums
        const uint32_t v = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    for (size_t c = lo; c < hi; c++) {
        of[c] = run;
        cu[c] = run;
        run += cn[c];
    }
    if (threadIdx.x == 1023) of[cells] = part[1023];
}

__global__ void __launch_bounds__(kThreads) k_cell_fill(FrameView fv, GroupView gb, const uint32_t *cell_of, uint32_t *cursor, float4 *sorted,
                                                         size_t cells) {
    const int f = blockIdx.y;
    const float *fr = fv.frame(f);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < gb.n; j += gridDim.x * blockDim.x) {
        const float *p = fr + (size_t)gb.atom(j) * 3;
        const uint32_t c = cell_of[(size_t)f * gb.n + j];
        const uint32_t pos = atomicAdd(cursor + (size_t)f * cells + c, 1u);
        sorted[(size_t)f * gb.n + pos] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __uint_as_float(j));
    }
}

// neighbour offsets along one axis of n cells around cell c without visiting a cell twice (cellgrid.rs:207-231)
__device__ __forceinline__ void axis_cells(int c, int n, int (&out)[3], int &m) {
    if (n >= 3) {
        out[0] = c == 0 ? n - 1 : c - 1;
        out[1] = c;
        out[2] = c == n - 1 ? 0 : c + 1;
        m = 3;
    } else {
        for (int k = 0; k < n; k++) out[k] = k;
        m = n;
    }
}

__global__ void __launch_bounds__(kThreads) k_cell_query(FrameView fv, GroupView ga, uint32_t nb_atoms, CellGeom cg, const uint32_t *offsets,
                                                          const float4 *sorted, size_t cells, float cutoff2, unsigned long long *count,
                                                          uint32_t *pairs, float *dist, unsigned long long capacity,
                                                          unsigned long long *cursor, const unsigned int *far) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    BoxOrtho B;
    load_box(fv.box, f, B);
    const float *fr = fv.frame(f);
    const uint32_t *of = offsets + (size_t)f * (cells + 1);
    const float4 *sb = sorted + (size_t)f * nb_atoms;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const bool b_near = far[f] == 0u;
    unsigned long long mine = 0;
    for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < ga.n; i += warps) {
        const float *p = fr + (size_t)ga.atom(i) * 3;
        const float ax = __ldg(p), ay = __ldg(p + 1), az = __ldg(p + 2);
        int xs[3], ys[3], zs[3], mx, my, mz;
        axis_cells(cell_coord(ax, B.lx, cg.nx), cg.nx, xs, mx);
        axis_cells(cell_coord(ay, B.ly, cg.ny), cg.ny, ys, my);
        axis_cells(cell_coord(az, B.lz, cg.nz), cg.nz, zs, mz);
        // the (up to 27) cell ranges are fetched by 27 lanes at once and handed round by shuffles: one L2 round trip per
        // atom instead of one per cell
        const int ncell = mx * my * mz;
        const bool fold = b_near && atom_in_fold_range<7>(ax, ay, az, B); // warp-uniform
        uint32_t my_lo = 0, my_hi = 0;
        if (lane < ncell) {
            const int kx = lane % mx, ky = (lane / mx) % my, kz = lane / (mx * my);
            const uint32_t c = ((uint32_t)zs[kz] * cg.ny + ys[ky]) * cg.nx + xs[kx];
            my_lo = of[c];
            my_hi = of[c + 1];
        }
        for (int k = 0; k < ncell; k++) {
            const uint32_t lo = __shfl_sync(0xffffffffu, my_lo, k), hi = __shfl_sync(0xffffffffu, my_hi, k);
            // four candidates per lane and trip, their loads issued before the first use
            for (uint32_t s0 = lo; s0 < hi; s0 += 128) {
                float4 b[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t s = s0 + u * 32 + lane;
                    b[u] = s < hi ? sb[s] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t s = s0 + u * 32 + lane;
                    // Vector3D::distance, XYZ (vector3d.rs:458-486), compared before the square root: cutoff2 is the
                    // smallest float whose sqrtf reaches the cutoff (host: cutoff_squared_threshold), sqrtf is monotone.
                    // With both atoms within L/4 of the box |min_image(d)| = min(|d|, ||d| - L|) exactly:
                    // no loops, no divergence bookkeeping; otherwise the reference's loops.
                    float dx, dy, dz;
                    if (fold) {
                        const float rx = fabsf(ax - b[u].x), ry = fabsf(ay - b[u].y), rz = fabsf(az - b[u].z);
                        dx = fminf(rx, fabsf(rx - B.lx));
                        dy = fminf(ry, fabsf(ry - B.ly));
                        dz = fminf(rz, fabsf(rz - B.lz));
                    } else {
                        dx = min_image(ax - b[u].x, B.lx);
                        dy = min_image(ay - b[u].y, B.ly);
                        dz = min_image(az - b[u].z, B.lz);
                    }
                    const float d2 = (dx * dx + dy * dy) + dz * dz;
                    const bool hit = s < hi && d2 < cutoff2;
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m == 0u) continue;
                    const int n_hit = __popc(m);
                    if (lane == 0) mine += n_hit;
                    if (pairs) { // one atomic per warp and step; pairs beyond the capacity are counted but not stored
                        unsigned long long base = 0;
                        if (lane == 0) base = atomicAdd(cursor + f, (unsigned long long)n_hit);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const unsigned long long at = base + __popc(m & ((1u << lane) - 1u));
                        if (hit && at < capacity) {
                            uint32_t *o = pairs + ((size_t)f * capacity + at) * 2;
                            o[0] = i;
                            o[1] = __float_as_uint(b[u].w);
                            if (dist) dist[(size_t)f * capacity + at] = sqrt1_rn(d2);
                        }
                    }
                }
            }
        }
    }
    if (lane == 0 && mine) atomicAdd(count + f, mine);
}

// ---------------------------------------------------------------- query, group A sorted by cell as well
// k_cell_query spends a 16-byte L2 load and the whole candidate bookkeeping on every single distance.  When group A is
// binned too (the same three kernels), one CTA takes one cell of A: its atoms go through in tiles of four held in
// registers (one tile per warp at a time), and every candidate a lane loads is compared with all four.
constexpr int kCellTileA = 4;
constexpr int kCellStage = 384; // staged hits per warp (flushed when fewer than 32 * kCellTileA slots are left)

template <bool STORE>
__global__ void __launch_bounds__(kThreads) k_cell_query_tiled(FrameView fv, uint32_t na_atoms, uint32_t nb_atoms, CellGeom cg,
                                                                const uint32_t *offsets_a, const float4 *sorted_a, const uint32_t *offsets_b,
                                                                const float4 *sorted_b, size_t cells, float cutoff2, unsigned long long *count,
                                                                uint32_t *pairs, float *dist, unsigned long long capacity,
                                                                unsigned long long *cursor, const unsigned int *far_a, const unsigned int *far_b) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    BoxOrtho B;
    load_box(fv.box, f, B);
    const uint32_t *ofa = offsets_a + (size_t)f * (cells + 1), *ofb = offsets_b + (size_t)f * (cells + 1);
    const float4 *sa = sorted_a + (size_t)f * na_atoms, *sb = sorted_b + (size_t)f * nb_atoms;
    const bool fold = far_a[f] == 0u && far_b[f] == 0u; // every atom of both groups within L/4 of the box: one-step fold
    const uint32_t wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long mine = 0;
    // Hits are staged per warp in shared memory and appended to the frame's list kCellStage / 2 or more at a time: one
    // atomic on the list cursor and one coalesced burst of 8-byte stores per flush.  (One atomic per 32 candidates -- every
    // step has hits at liquid density -- serialised the whole grid on a single L2 address: 9.9 ms against 2.0 ms counting only.)
    __shared__ uint2 stage_p[STORE ? kThreads / 32 : 1][STORE ? kCellStage : 1];
    __shared__ float stage_d[STORE ? kThreads / 32 : 1][STORE ? kCellStage : 1];
    uint32_t fill = 0; // warp-uniform
    uint32_t lane_hits = 0; // !STORE: far below 2^32 evaluations per lane
    auto flush = [&]() {
        if (fill == 0u) return;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor + f, (unsigned long long)fill);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (uint32_t k = lane; k < fill; k += 32) {
            const unsigned long long at = base + k;
            if (at < capacity) { // pairs beyond the capacity are counted but not stored
                reinterpret_cast<uint2 *>(pairs)[(size_t)f * capacity + at] = stage_p[wid][k];
                if (dist) dist[(size_t)f * capacity + at] = stage_d[wid][k];
            }
        }
        __syncwarp();
        fill = 0u;
    };
    // one CTA per cell of A (grid-stride), its warps take the cell's tiles in turn: units of a few hundred distance
    // evaluations keep the SMs evenly loaded (one warp per cell left 17 % of the warp slots busy: long, uneven tasks)
    for (uint32_t ca = blockIdx.x; ca < cells; ca += gridDim.x) {
        const uint32_t alo = ofa[ca], ahi = ofa[ca + 1];
        if (alo == ahi) continue;
        const int cx = (int)(ca % cg.nx), cy = (int)((ca / cg.nx) % cg.ny), cz = (int)(ca / ((uint32_t)cg.nx * cg.ny));
        int xs[3], ys[3], zs[3], mx, my, mz;
        axis_cells(cx, cg.nx, xs, mx);
        axis_cells(cy, cg.ny, ys, my);
        axis_cells(cz, cg.nz, zs, mz);
        const int ncell = mx * my * mz;
        uint32_t my_lo = 0, my_hi = 0;
        if (lane < ncell) {
            const int kx = lane % mx, ky = (lane / mx) % my, kz = lane / (mx * my);
            const int vx = kx == 0 ? xs[0] : (kx == 1 ? xs[1] : xs[2]), vy = ky == 0 ? ys[0] : (ky == 1 ? ys[1] : ys[2]),
                      vz = kz == 0 ? zs[0] : (kz == 1 ? zs[1] : zs[2]);
            const uint32_t c = ((uint32_t)vz * cg.ny + vy) * cg.nx + vx;
            my_lo = ofb[c];
            my_hi = ofb[c + 1];
        }
        for (uint32_t a0 = alo + wid * kCellTileA; a0 < ahi; a0 += nw * kCellTileA) {
            float ax[kCellTileA], ay[kCellTileA], az[kCellTileA];
            uint32_t ai[kCellTileA];
            const int na = (int)min((uint32_t)kCellTileA, ahi - a0);
#pragma unroll
            for (int t = 0; t < kCellTileA; t++) {
                const float4 a = sa[min(a0 + t, ahi - 1)]; // warp-uniform address: one transaction, broadcast
                ax[t] = a.x; ay[t] = a.y; az[t] = a.z; ai[t] = __float_as_uint(a.w);
            }
            for (int k = 0; k < ncell; k++) {
                const uint32_t lo = __shfl_sync(0xffffffffu, my_lo, k), hi = __shfl_sync(0xffffffffu, my_hi, k);
                for (uint32_t s0 = lo; s0 < hi; s0 += 64) {
                    float4 b[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const uint32_t s = s0 + u * 32 + lane;
                        b[u] = s < hi ? sb[s] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const bool in = s0 + u * 32 + lane < hi;
                        if (STORE && fill > (uint32_t)(kCellStage - 32 * kCellTileA)) flush(); // room for this batch's hits
#pragma unroll
                        for (int t = 0; t < kCellTileA; t++) {
                            float dx, dy, dz;
                            if (fold) {
                                const float rx = fabsf(ax[t] - b[u].x), ry = fabsf(ay[t] - b[u].y), rz = fabsf(az[t] - b[u].z);
                                dx = fminf(rx, fabsf(rx - B.lx));
                                dy = fminf(ry, fabsf(ry - B.ly));
                                dz = fminf(rz, fabsf(rz - B.lz));
                            } else {
                                dx = min_image(ax[t] - b[u].x, B.lx);
                                dy = min_image(ay[t] - b[u].y, B.ly);
                                dz = min_image(az[t] - b[u].z, B.lz);
                            }
                            const float d2 = (dx * dx + dy * dy) + dz * dz;
                            const bool hit = in && t < na && d2 < cutoff2;
                            if (!STORE) { // counting only: a private counter per lane, no vote, no branch
                                lane_hits += hit ? 1u : 0u;
                                continue;
                            }
                            const unsigned m = __ballot_sync(0xffffffffu, hit);
                            if (m == 0u) continue;
                            const int n_hit = __popc(m);
                            if (lane == 0) mine += n_hit;
                            if (STORE) {
                                if (hit) {
                                    const uint32_t k = fill + __popc(m & ((1u << lane) - 1u));
                                    stage_p[wid][k] = make_uint2(ai[t], __float_as_uint(b[u].w));
                                    if (dist) stage_d[wid][k] = sqrt1_rn(d2);
                                }
                                fill += (uint32_t)n_hit;
                            }
                        }
                    }
                }
            }
        }
    }
    if (STORE) flush();
    if (!STORE) {
        const uint32_t tot = __reduce_add_sync(0xffffffffu, lane_hits);
        if (lane == 0) mine += tot;
    }
    if (lane == 0 && mine) atomicAdd(count + f, mine);
}

} // namespace groan

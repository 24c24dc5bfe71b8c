// groan_pairs.cu -- host side of the distance ops of libgroan_gpu.so: group_all_distances, its fused reduction, and the
// cutoff pair search through a cell grid (include/groan_gpu.h).  Split from groan_gpu.cu so that the translation units
// compile in parallel.
#include "ctx.cuh"
#include "kernels_pairs.cuh"
#include "kernels_cells.cuh"

using namespace groan;
using namespace groan_host;
static_assert(sizeof(PairPartial) == kPairPartialBytes, "ctx.cuh: d_pair_partials");

namespace {

template <int DIM, typename BOX>
int launch_pairs(groan_gpu_ctx *ctx, const Group &a, const Group &b, float *d_out) {
    const bool vec = (b.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
    dim3 grid((unsigned)((b.n + (size_t)kThreads * kPairJ - 1) / ((size_t)kThreads * kPairJ)),
              (unsigned)((a.n + kPairRows - 1) / kPairRows), (unsigned)ctx->n_frames);
    if (grid.y > 65535u || grid.z > 65535u) return GROAN_ECAPACITY;  // > 2M rows: such a matrix does not fit any memory anyway
    if (vec)
        k_pairs<DIM, BOX, true><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), d_out);
    else
        k_pairs<DIM, BOX, false><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), d_out);
    LAUNCHED();
    return GROAN_OK;
}

// orthogonal box, 2-D / 3-D distance: packed one-step min-image (kernels_pairs.cuh "fast paths")
template <int DIM, typename BOX = BoxOrtho>
int launch_pairs_fast(groan_gpu_ctx *ctx, const Group &a, const Group &b, float *d_out) {
    const bool vec = (b.n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15) == 0);
    dim3 grid((unsigned)((b.n + (size_t)kThreads * kPairJ - 1) / ((size_t)kThreads * kPairJ)),
              (unsigned)((a.n + kFastRows - 1) / kFastRows), (unsigned)ctx->n_frames);
    if (grid.y > 65535u || grid.z > 65535u) return GROAN_ECAPACITY;
    if (vec)
        k_pairs_fast<DIM, true, BOX><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), d_out);
    else
        k_pairs_fast<DIM, false, BOX><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), d_out);
    LAUNCHED();
    return GROAN_OK;
}

template <typename BOX>
int dispatch_pairs(groan_gpu_ctx *ctx, int dim, const Group &a, const Group &b, float *d_out) {
    if (std::is_same<BOX, BoxOrtho>::value) {
        switch (dim) {
        case 4: return launch_pairs_fast<4>(ctx, a, b, d_out);
        case 5: return launch_pairs_fast<5>(ctx, a, b, d_out);
        case 6: return launch_pairs_fast<6>(ctx, a, b, d_out);
        case 7: return launch_pairs_fast<7>(ctx, a, b, d_out);
        default: break;
        }
    } else {
        // triclinic extension, 2-D / 3-D: the same kernel with the 27-image d^2 (kernels_pairs.cuh pair_d2_tric)
        switch (dim) {
        case 4: return launch_pairs_fast<4, BoxTric>(ctx, a, b, d_out);
        case 5: return launch_pairs_fast<5, BoxTric>(ctx, a, b, d_out);
        case 6: return launch_pairs_fast<6, BoxTric>(ctx, a, b, d_out);
        case 7: return launch_pairs_fast<7, BoxTric>(ctx, a, b, d_out);
        default: break;
        }
    }
    switch (dim) {
    case 0: return launch_pairs<0, BOX>(ctx, a, b, d_out);
    case 1: return launch_pairs<1, BOX>(ctx, a, b, d_out);
    case 2: return launch_pairs<2, BOX>(ctx, a, b, d_out);
    case 3: return launch_pairs<3, BOX>(ctx, a, b, d_out);
    case 4: return launch_pairs<4, BOX>(ctx, a, b, d_out);
    case 5: return launch_pairs<5, BOX>(ctx, a, b, d_out);
    case 6: return launch_pairs<6, BOX>(ctx, a, b, d_out);
    case 7: return launch_pairs<7, BOX>(ctx, a, b, d_out);
    default: return GROAN_EINVAL;
    }
}

struct ReduceOut {
    float *dmin;
    uint32_t *imin;
    float *dmax;
    uint32_t *imax;
    unsigned long long *count;
};

template <int DIM, typename BOX>
int launch_pairs_reduce(groan_gpu_ctx *ctx, const Group &a, const Group &b, float cutoff, const ReduceOut &o) {
    size_t nb = (b.n + (size_t)kThreads * kPairJ - 1) / ((size_t)kThreads * kPairJ);
    nb = std::max<size_t>(1, std::min<size_t>(nb, kMaxBlocksPerFrame));
    nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / ctx->n_frames));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    k_pairs_reduce<DIM, BOX><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), cutoff,
                                                                  (PairPartial *)ctx->d_pair_partials, ctx->d_tickets, o.dmin,
                                                                  o.imin, o.dmax, o.imax, o.count);
    LAUNCHED();
    return GROAN_OK;
}

template <int DIM, typename BOX = BoxOrtho>
int launch_pairs_reduce_fast(groan_gpu_ctx *ctx, const Group &a, const Group &b, float cutoff, const ReduceOut &o) {
    // persistent CTAs (4 per SM in total) striding over work units of (1024 B atoms) x (256 A atoms)
    const size_t units = ((b.n + (size_t)kThreads * kPairJ - 1) / ((size_t)kThreads * kPairJ)) * ((a.n + kSliceA - 1) / kSliceA);
    size_t nb = std::max<size_t>(1, ((size_t)kSMs * 4) / ctx->n_frames);
    nb = std::min<size_t>(nb, units);
    nb = std::max<size_t>(1, std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / ctx->n_frames)));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    const float c2 = cutoff_squared_threshold(cutoff);
    if (o.count)
        k_pairs_reduce_fast<DIM, true, BOX><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), cutoff, c2,
                                                                            (PairPartial *)ctx->d_pair_partials, ctx->d_tickets, o.dmin,
                                                                            o.imin, o.dmax, o.imax, o.count);
    else
        k_pairs_reduce_fast<DIM, false, BOX><<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(a), view_of(b), cutoff, c2,
                                                                             (PairPartial *)ctx->d_pair_partials, ctx->d_tickets, o.dmin,
                                                                             o.imin, o.dmax, o.imax, o.count);
    LAUNCHED();
    return GROAN_OK;
}

template <typename BOX>
int dispatch_pairs_reduce(groan_gpu_ctx *ctx, int dim, const Group &a, const Group &b, float cutoff, const ReduceOut &o) {
    if (std::is_same<BOX, BoxOrtho>::value) {
        switch (dim) {
        case 4: return launch_pairs_reduce_fast<4>(ctx, a, b, cutoff, o);
        case 5: return launch_pairs_reduce_fast<5>(ctx, a, b, cutoff, o);
        case 6: return launch_pairs_reduce_fast<6>(ctx, a, b, cutoff, o);
        case 7: return launch_pairs_reduce_fast<7>(ctx, a, b, cutoff, o);
        default: break;
        }
    } else {
        // triclinic extension, 2-D / 3-D: the same kernel with the 27-image d^2 (kernels_pairs.cuh pair_d2_tric)
        switch (dim) {
        case 4: return launch_pairs_reduce_fast<4, BoxTric>(ctx, a, b, cutoff, o);
        case 5: return launch_pairs_reduce_fast<5, BoxTric>(ctx, a, b, cutoff, o);
        case 6: return launch_pairs_reduce_fast<6, BoxTric>(ctx, a, b, cutoff, o);
        case 7: return launch_pairs_reduce_fast<7, BoxTric>(ctx, a, b, cutoff, o);
        default: break;
        }
    }
    switch (dim) {
    case 0: return launch_pairs_reduce<0, BOX>(ctx, a, b, cutoff, o);
    case 1: return launch_pairs_reduce<1, BOX>(ctx, a, b, cutoff, o);
    case 2: return launch_pairs_reduce<2, BOX>(ctx, a, b, cutoff, o);
    case 3: return launch_pairs_reduce<3, BOX>(ctx, a, b, cutoff, o);
    case 4: return launch_pairs_reduce<4, BOX>(ctx, a, b, cutoff, o);
    case 5: return launch_pairs_reduce<5, BOX>(ctx, a, b, cutoff, o);
    case 6: return launch_pairs_reduce<6, BOX>(ctx, a, b, cutoff, o);
    case 7: return launch_pairs_reduce<7, BOX>(ctx, a, b, cutoff, o);
    default: return GROAN_EINVAL;
    }
}
}  // namespace

extern "C" {

int groan_gpu_all_distances(groan_gpu_ctx *ctx, int g1, int g2, int dim, float *out) {
    if (!ctx || dim < 0 || dim > 7) return GROAN_EINVAL;
    const Group *a = get_group(ctx, g1), *b = get_group(ctx, g2);
    if (!a || !b) return GROAN_ENOGROUP;  // group_get_n_atoms (analysis.rs:407-408)
    bool tric = false;
    int rc = check_box(ctx, true, &tric);
    if (rc) return rc;
    if (a->n == 0 || b->n == 0) return GROAN_OK;  // empty matrix, not an error (analysis.rs:412)
    if (!out) return GROAN_EINVAL;
    rc = check_pair_positions(ctx, *a, *b);
    if (rc) return rc;
    const size_t bytes = ctx->n_frames * a->n * b->n * sizeof(float);
    float *d_out = out;
    if (classify(out) != PK_DEVICE) {
        rc = ensure_tmp(ctx, bytes);
        if (rc) return rc;
        d_out = (float *)ctx->d_tmp;
    }
    rc = tric ? dispatch_pairs<BoxTric>(ctx, dim, *a, *b, d_out) : dispatch_pairs<BoxOrtho>(ctx, dim, *a, *b, d_out);
    if (rc) return rc;
    return deliver(ctx, out, d_out, bytes);
}

int groan_gpu_all_distances_reduce(groan_gpu_ctx *ctx, int g1, int g2, int dim, float cutoff, float *dmin, uint32_t *imin,
                                   float *dmax, uint32_t *imax, uint64_t *count) {
    if (!ctx || dim < 0 || dim > 7) return GROAN_EINVAL;
    const Group *a = get_group(ctx, g1), *b = get_group(ctx, g2);
    if (!a || !b) return GROAN_ENOGROUP;
    bool tric = false;
    int rc = check_box(ctx, true, &tric);
    if (rc) return rc;
    if (a->n == 0 || b->n == 0) return GROAN_EEMPTY;  // min/max of an empty matrix: Option::unwrap panics in the documented consumer
    rc = check_pair_positions(ctx, *a, *b);
    if (rc) return rc;
    const size_t F = ctx->n_frames;
    // scratch layout inside d_res (8 floats per frame): dmin | dmax | imin(2) | imax(2) | count(u64)
    float *s = ctx->d_res;
    ReduceOut o;
    o.dmin = target_of<float>(dmin, s);
    o.dmax = target_of<float>(dmax, s + F);
    o.imin = target_of<uint32_t>(imin, (uint32_t *)(s + 2 * F));
    o.imax = target_of<uint32_t>(imax, (uint32_t *)(s + 4 * F));
    o.count = count ? target_of<unsigned long long>(count, (unsigned long long *)(s + 6 * F)) : nullptr;
    rc = tric ? dispatch_pairs_reduce<BoxTric>(ctx, dim, *a, *b, cutoff, o) : dispatch_pairs_reduce<BoxOrtho>(ctx, dim, *a, *b, cutoff, o);
    if (rc) return rc;
    if ((rc = deliver(ctx, dmin, o.dmin, F * sizeof(float)))) return rc;
    if ((rc = deliver(ctx, dmax, o.dmax, F * sizeof(float)))) return rc;
    if ((rc = deliver(ctx, imin, o.imin, F * 2 * sizeof(uint32_t)))) return rc;
    if ((rc = deliver(ctx, imax, o.imax, F * 2 * sizeof(uint32_t)))) return rc;
    return deliver(ctx, count, o.count, F * sizeof(uint64_t));
}

// ---- cutoff pair search through a cell grid ---------------------------------------
int groan_gpu_pairs_within(groan_gpu_ctx *ctx, int g1, int g2, float cutoff, uint64_t *count, uint32_t *pairs, float *dist,
                           size_t capacity) {
    if (!ctx || !count || !(cutoff > 0.0f) || (dist && !pairs)) return GROAN_EINVAL;
    const Group *a = get_group(ctx, g1), *b = get_group(ctx, g2);
    if (!a || !b) return GROAN_ENOGROUP;
    // CellGrid::new: the box must exist and be orthogonal (cellgrid.rs:308-312), then the positions of the group
    int rc = check_box(ctx, false, nullptr);
    if (rc) return rc;
    rc = check_pair_positions(ctx, *a, *b);
    if (rc) return rc;
    const size_t F = ctx->n_frames, nb_atoms = b->n;
    if (!pairs) capacity = 0;
    // one grid geometry for the batch, from the smallest box: cells at least cutoff * (1 + 1e-4) wide in every frame
    float lmin[3] = {3.0e38f, 3.0e38f, 3.0e38f};
    for (size_t f = 0; f < F; f++)
        for (int k = 0; k < 3; k++) lmin[k] = std::min(lmin[k], ctx->h_box[f * 9 + 4 * k]);
    long nc[3];
    for (int k = 0; k < 3; k++) nc[k] = std::max<long>(1, std::min<long>(1024, (long)std::floor((double)lmin[k] / ((double)cutoff * 1.0001))));
    const size_t cell_cap = std::max<size_t>(4096, std::min<size_t>((size_t)8 << 20, 4 * nb_atoms + 4096));
    while ((size_t)nc[0] * nc[1] * nc[2] > cell_cap) {  // wider cells are always correct, only slower
        const int k = nc[0] >= nc[1] && nc[0] >= nc[2] ? 0 : (nc[1] >= nc[2] ? 1 : 2);
        nc[k] = (nc[k] + 1) / 2;
    }
    CellGeom cg = {(int)nc[0], (int)nc[1], (int)nc[2]};
    const float cutoff2 = cutoff_squared_threshold(cutoff);
    const size_t cells = (size_t)nc[0] * nc[1] * nc[2];
    // scratch layout (one allocation): results first, then per-frame grid storage for as many frames as fit ~1.5 GB.
    // Group A is binned as well when it has at least one atom per cell on average: one warp then serves a whole cell of A
    // (k_cell_query_tiled); sparse query groups keep one warp per atom (k_cell_query).
    const size_t na_atoms = a->n;
    const bool tiled = na_atoms >= cells && !(ctx->flags & GROAN_FLAG_NO_TMA);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    auto grid_bytes = [&](size_t atoms) { return up(atoms * 4) + up(cells * 4) * 2 + up((cells + 1) * 4) + up(atoms * 16); };
    const size_t per_frame = grid_bytes(nb_atoms) + (tiled ? grid_bytes(na_atoms) : 0);
    const size_t fb = std::max<size_t>(1, std::min<size_t>(F, ((size_t)3 << 29) / std::max<size_t>(per_frame, 1)));
    const bool stage_pairs = pairs && classify(pairs) != PK_DEVICE, stage_dist = dist && classify(dist) != PK_DEVICE;
    const size_t o_count = 0, o_cursor = up(F * 8), o_far = o_cursor + up(F * 8), o_pairs = o_far + up(2 * F * 4),
                 o_dist = o_pairs + (stage_pairs ? up(F * capacity * 8) : 0), o_grid = o_dist + (stage_dist ? up(F * capacity * 4) : 0);
    rc = ensure_tmp(ctx, o_grid + fb * per_frame);
    if (rc) return rc;
    char *base = (char *)ctx->d_tmp;
    unsigned long long *d_count = (unsigned long long *)(base + o_count), *d_cursor = (unsigned long long *)(base + o_cursor);
    unsigned int *d_far_b = (unsigned int *)(base + o_far), *d_far_a = d_far_b + F;
    uint32_t *d_pairs = pairs ? (stage_pairs ? (uint32_t *)(base + o_pairs) : pairs) : nullptr;
    float *d_dist = dist ? (stage_dist ? (float *)(base + o_dist) : dist) : nullptr;
    CK(cudaMemsetAsync(base, 0, o_pairs, ctx->compute));
    struct CellLists {
        uint32_t *cell_of, *counts, *fill, *offsets;
        float4 *sorted;
    };
    auto carve = [&](char *p0, size_t atoms) {
        CellLists c;
        c.cell_of = (uint32_t *)p0;
        c.counts = (uint32_t *)(p0 + fb * up(atoms * 4));
        c.fill = (uint32_t *)((char *)c.counts + fb * up(cells * 4));
        c.offsets = (uint32_t *)((char *)c.fill + fb * up(cells * 4));
        c.sorted = (float4 *)((char *)c.offsets + fb * up((cells + 1) * 4));
        return c;
    };
    const CellLists lb = carve(base + o_grid, nb_atoms);
    const CellLists la = tiled ? carve(base + o_grid + fb * grid_bytes(nb_atoms), na_atoms) : CellLists();
    const GroupView ga = view_of(*a), gb = view_of(*b);
    for (size_t f0 = 0; f0 < F; f0 += fb) {
        const size_t nf = std::min(fb, F - f0);
        FrameView fv = frames_of(ctx);
        fv.xyz += f0 * ctx->n_atoms * 3;
        fv.box += f0 * 9;
        // counting sort of a group by cell: histogram, prefix sum, scatter
        auto build = [&](const GroupView &gv, size_t atoms, const CellLists &cl, unsigned int *far) -> int {
            CK(cudaMemsetAsync(cl.counts, 0, nf * cells * 4, ctx->compute));
            const unsigned nbk = (unsigned)std::max<size_t>(1, std::min<size_t>((atoms + kThreads - 1) / kThreads, (size_t)kSMs * 8));
            if (atoms) {
                k_cell_count<<<dim3(nbk, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, gv, cg, cl.cell_of, cl.counts, cells, far + f0);
                LAUNCHED();
            }
            k_cell_scan<<<(unsigned)nf, 1024, 0, ctx->compute>>>(cl.counts, cl.offsets, cl.fill, cells);
            LAUNCHED();
            if (atoms) {
                k_cell_fill<<<dim3(nbk, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, gv, cl.cell_of, cl.fill, cl.sorted, cells);
                LAUNCHED();
            }
            return GROAN_OK;
        };
        rc = build(gb, nb_atoms, lb, d_far_b);
        if (rc) return rc;
        uint32_t *pp = d_pairs ? d_pairs + f0 * capacity * 2 : nullptr;
        float *dp = d_dist ? d_dist + f0 * capacity : nullptr;
        if (tiled) {
            rc = build(ga, na_atoms, la, d_far_a);
            if (rc) return rc;
            const unsigned nq = (unsigned)std::max<size_t>(1, std::min<size_t>(cells, (size_t)kSMs * 64));
            if (pp)
                k_cell_query_tiled<true><<<dim3(nq, (unsigned)nf), kThreads, 0, ctx->compute>>>(
                    fv, (uint32_t)na_atoms, (uint32_t)nb_atoms, cg, la.offsets, la.sorted, lb.offsets, lb.sorted, cells, cutoff2, d_count + f0, pp,
                    dp, (unsigned long long)capacity, d_cursor + f0, d_far_a + f0, d_far_b + f0);
            else
                k_cell_query_tiled<false><<<dim3(nq, (unsigned)nf), kThreads, 0, ctx->compute>>>(
                    fv, (uint32_t)na_atoms, (uint32_t)nb_atoms, cg, la.offsets, la.sorted, lb.offsets, lb.sorted, cells, cutoff2, d_count + f0, pp,
                    dp, (unsigned long long)capacity, d_cursor + f0, d_far_a + f0, d_far_b + f0);
            LAUNCHED();
        } else if (na_atoms) {
            const unsigned nqa = (unsigned)std::max<size_t>(1, std::min<size_t>((na_atoms + 7) / 8, (size_t)kSMs * 16));
            k_cell_query<<<dim3(nqa, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, ga, (uint32_t)nb_atoms, cg, lb.offsets, lb.sorted, cells, cutoff2,
                                                                               d_count + f0, pp, dp, (unsigned long long)capacity, d_cursor + f0,
                                                                               d_far_b + f0);
            LAUNCHED();
        }
    }
    if ((rc = deliver(ctx, count, d_count, F * sizeof(uint64_t)))) return rc;
    if (stage_pairs && (rc = deliver(ctx, pairs, d_pairs, F * capacity * 8))) return rc;
    if (stage_dist && (rc = deliver(ctx, dist, d_dist, F * capacity * 4))) return rc;
    return GROAN_OK;
}
}  // extern "C"

#include "groan_bonds.inl"

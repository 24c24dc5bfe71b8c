// groan_bonds.inl -- host side of the two users of the cell grid (include/groan_gpu.h): groan_gpu_guess_bonds
// (System::guess_bonds, src/system/guess.rs:362-470) and groan_gpu_hbonds (HBondAnalysis::analyze_single,
// src/system/hbonds.rs:240-320).  Both bin one group into cells per frame (kernels_cells.cuh: histogram, prefix sum, scatter)
// and walk the neighbourhood of their query atoms with a warp each (kernels_bonds.cuh).
// Included at the end of groan_pairs.cu: the grid kernels of kernels_cells.cuh are plain __global__ functions and may be
// defined in one translation unit only.
#include "kernels_bonds.cuh"

namespace {

// smallest float t with sqrtf(t) > c: (d2 < t) <=> (sqrtf(d2) <= c) for every float d2 >= 0
float cutoff_squared_threshold_le(float c) {
    if (!(c >= 0.0f)) return 0.0f;
    float t = c * c;
    if (std::isinf(t)) return t;
    while (std::sqrt(t) > c && t > 0.0f) t = std::nextafter(t, 0.0f);
    while (std::sqrt(t) <= c) t = std::nextafter(t, INFINITY);
    return t;
}

struct Grid {
    CellGeom cg;
    size_t cells = 0, fb = 0;       // cells per frame, frames per pass
    uint32_t *cell_of = nullptr, *counts = nullptr, *fill = nullptr, *offsets = nullptr;
    float4 *sorted = nullptr;
    unsigned long long *d_count = nullptr, *d_cursor = nullptr;
    unsigned int *d_far = nullptr;
    char *out = nullptr;            // staging for results that go to host memory
};

// one grid geometry for the batch, from the smallest box: cells at least `width` * (1 + 1e-4) wide in every frame; scratch
// carved from ctx->d_tmp: [count F][cursor F][far F][out_bytes][per-frame grid storage x fb]
int plan_grid(groan_gpu_ctx *ctx, float width, size_t n_binned, size_t out_bytes, Grid *g) {
    const size_t F = ctx->n_frames;
    float lmin[3] = {3.0e38f, 3.0e38f, 3.0e38f};
    for (size_t f = 0; f < F; f++)
        for (int k = 0; k < 3; k++) lmin[k] = std::min(lmin[k], ctx->h_box[f * 9 + 4 * k]);
    long nc[3];
    for (int k = 0; k < 3; k++) nc[k] = std::max<long>(1, std::min<long>(1024, (long)std::floor((double)lmin[k] / ((double)width * 1.0001))));
    // at most one cell per four binned atoms: a warp walks a neighbourhood 32 candidates at a time, so cells much emptier than
    // that only add range bookkeeping (bond guessing asks for 0.17 nm cells: half an atom each) and make the one-CTA prefix sum
    // over the cells the longest kernel of the call
    const size_t cell_cap = std::max<size_t>(4096, n_binned / 4);
    while ((size_t)nc[0] * nc[1] * nc[2] > cell_cap) {  // wider cells are always correct, only slower
        const int k = nc[0] >= nc[1] && nc[0] >= nc[2] ? 0 : (nc[1] >= nc[2] ? 1 : 2);
        nc[k] = (nc[k] + 1) / 2;
    }
    g->cg = {(int)nc[0], (int)nc[1], (int)nc[2]};
    g->cells = (size_t)nc[0] * nc[1] * nc[2];
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t per_frame = up(n_binned * 4) + up(g->cells * 4) * 2 + up((g->cells + 1) * 4) + up(n_binned * 16);
    g->fb = std::max<size_t>(1, std::min<size_t>(F, ((size_t)3 << 29) / std::max<size_t>(per_frame, 1)));
    const size_t o_cursor = up(F * 8), o_far = o_cursor + up(F * 8), o_out = o_far + up(F * 4), o_grid = o_out + up(out_bytes);
    int rc = ensure_tmp(ctx, o_grid + g->fb * per_frame);
    if (rc) return rc;
    char *base = (char *)ctx->d_tmp;
    g->d_count = (unsigned long long *)base;
    g->d_cursor = (unsigned long long *)(base + o_cursor);
    g->d_far = (unsigned int *)(base + o_far);
    g->out = base + o_out;
    CK(cudaMemsetAsync(base, 0, o_out, ctx->compute));
    char *p0 = base + o_grid;
    g->cell_of = (uint32_t *)p0;
    g->counts = (uint32_t *)(p0 + g->fb * up(n_binned * 4));
    g->fill = (uint32_t *)((char *)g->counts + g->fb * up(g->cells * 4));
    g->offsets = (uint32_t *)((char *)g->fill + g->fb * up(g->cells * 4));
    g->sorted = (float4 *)((char *)g->offsets + g->fb * up((g->cells + 1) * 4));
    return GROAN_OK;
}

// counting sort of a group by cell for frames [f0, f0 + nf): histogram, prefix sum, scatter
int build_grid(groan_gpu_ctx *ctx, const FrameView &fv, const GroupView &gv, const Grid &g, size_t f0, size_t nf) {
    CK(cudaMemsetAsync(g.counts, 0, nf * g.cells * 4, ctx->compute));
    const unsigned nbk = (unsigned)std::max<size_t>(1, std::min<size_t>(((size_t)gv.n + kThreads - 1) / kThreads, (size_t)kSMs * 8));
    if (gv.n) {
        k_cell_count<<<dim3(nbk, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, gv, g.cg, g.cell_of, g.counts, g.cells, g.d_far + f0);
        LAUNCHED();
    }
    k_cell_scan<<<(unsigned)nf, 1024, 0, ctx->compute>>>(g.counts, g.offsets, g.fill, g.cells);
    LAUNCHED();
    if (gv.n) {
        k_cell_fill<<<dim3(nbk, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, gv, g.cell_of, g.fill, g.sorted, g.cells);
        LAUNCHED();
    }
    return GROAN_OK;
}

}  // namespace

extern "C" {

int groan_gpu_guess_bonds(groan_gpu_ctx *ctx, const float *vdw, float radius_factor, uint64_t *count, uint32_t *pairs, size_t capacity) {
    if (!ctx || !vdw || !count || !(radius_factor > 0.0f) || classify(vdw) == PK_DEVICE) return GROAN_EINVAL;  // vdw: host array
    // CellGrid::new(self, "all", cell size) first (guess.rs:371): box, then positions of every atom
    int rc = check_box(ctx, false, nullptr);
    if (rc) return rc;
    rc = check_positions(ctx, ctx->all);
    if (rc) return rc;
    const size_t F = ctx->n_frames, N = ctx->n_atoms;
    if (!pairs) capacity = 0;
    // get_cell_size (guess.rs:397-407): twice the largest radius times the factor = the longest possible bond
    float max_vdw = -INFINITY;
    for (size_t i = 0; i < N; i++)
        if (vdw[i] >= 0.0f) max_vdw = std::max(max_vdw, vdw[i]);
    if (!(max_vdw > 0.0f)) {  // no atom has a radius: every atom is skipped (guess.rs:433-440), no bonds
        if (classify(count) == PK_DEVICE) {
            CK(cudaMemsetAsync(count, 0, F * sizeof(uint64_t), ctx->compute));
        } else {
            std::memset(count, 0, F * sizeof(uint64_t));
        }
        return GROAN_OK;
    }
    const float width = 2.0f * radius_factor * max_vdw;
    const bool stage = pairs && classify(pairs) != PK_DEVICE;
    Grid g;
    rc = plan_grid(ctx, width, N, (stage ? F * capacity * 8 : 0) + ((N * sizeof(float) + 255) & ~(size_t)255), &g);
    if (rc) return rc;
    float *d_vdw = (float *)g.out;
    uint32_t *d_pairs = pairs ? (stage ? (uint32_t *)(g.out + ((N * sizeof(float) + 255) & ~(size_t)255)) : pairs) : nullptr;
    CK(cudaMemcpyAsync(d_vdw, vdw, N * sizeof(float), cudaMemcpyDefault, ctx->compute));
    if (classify(vdw) == PK_PAGEABLE) CK(cudaStreamSynchronize(ctx->compute));
    const GroupView gall = view_of(ctx->all);
    for (size_t f0 = 0; f0 < F; f0 += g.fb) {
        const size_t nf = std::min(g.fb, F - f0);
        FrameView fv = frames_of(ctx);
        fv.xyz += f0 * N * 3;
        fv.box += f0 * 9;
        rc = build_grid(ctx, fv, gall, g, f0, nf);
        if (rc) return rc;
        const unsigned nq = (unsigned)std::max<size_t>(1, std::min<size_t>((N + 7) / 8, (size_t)kSMs * 16));
        k_guess_bonds<<<dim3(nq, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, (uint32_t)N, g.cg, g.offsets, g.sorted, g.cells, d_vdw, radius_factor,
                                                                            g.d_count + f0, d_pairs ? d_pairs + f0 * capacity * 2 : nullptr,
                                                                            (unsigned long long)capacity, g.d_cursor + f0, g.d_far + f0);
        LAUNCHED();
    }
    if ((rc = deliver(ctx, count, g.d_count, F * sizeof(uint64_t)))) return rc;
    if (stage && (rc = deliver(ctx, pairs, d_pairs, F * capacity * 8))) return rc;
    return GROAN_OK;
}

int groan_gpu_hbonds(groan_gpu_ctx *ctx, int acc_gid, const uint32_t *donors, const uint32_t *hyd_offsets, const uint32_t *hydrogens,
                     size_t n_donors, float max_distance, float min_angle, uint64_t *count, uint32_t *dha, float *dist_angle, size_t capacity) {
    if (!ctx || !count || !(max_distance > 0.0f) || (n_donors && (!donors || !hyd_offsets || !hydrogens)) || ((dha == nullptr) != (dist_angle == nullptr)))
        return GROAN_EINVAL;
    const Group *acc = get_group(ctx, acc_gid);
    if (!acc) return GROAN_ENOGROUP;
    // CellGrid::new_from_group(system, acceptors, max_distance) (hbonds.rs:166-172): box first, then the acceptors' positions
    int rc = check_box(ctx, false, nullptr);
    if (rc) return rc;
    rc = check_positions(ctx, *acc);
    if (rc) return rc;
    const size_t F = ctx->n_frames, N = ctx->n_atoms;
    const size_t n_hyd = n_donors ? hyd_offsets[n_donors] : 0;
    for (size_t d = 0; d < n_donors; d++)
        if (donors[d] >= N || hyd_offsets[d + 1] < hyd_offsets[d]) return GROAN_EINVAL;
    for (size_t h = 0; h < n_hyd; h++)
        if (hydrogens[h] >= N) return GROAN_EINVAL;
    if (ctx->has_valid && n_donors) {  // donor and hydrogen positions (hbonds.rs:252-258,284-289): asked once per call as an ad-hoc group
        std::vector<uint32_t> need(donors, donors + n_donors);
        need.insert(need.end(), hydrogens, hydrogens + n_hyd);
        std::sort(need.begin(), need.end());
        need.erase(std::unique(need.begin(), need.end()), need.end());
        Group tmp;
        tmp.set = true;
        tmp.n = need.size();
        tmp.idx = need;
        tmp.contiguous = false;
        CK(cudaMalloc(&tmp.d_idx, need.size() * sizeof(uint32_t)));
        cudaError_t e = cudaMemcpy(tmp.d_idx, need.data(), need.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
        rc = e == cudaSuccess ? check_positions(ctx, tmp) : GROAN_ECUDA;
        ctx->valid_cache.pop_back();  // the answer belongs to a group that is about to disappear
        cudaFree(tmp.d_idx);
        if (rc) return rc;
    }
    if (!dha) capacity = 0;
    const bool stage = dha && classify(dha) != PK_DEVICE;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b_don = up(n_donors * 4), b_off = up((n_donors + 1) * 4), b_hyd = up(std::max<size_t>(n_hyd, 1) * 4);
    Grid g;
    rc = plan_grid(ctx, max_distance, acc->n, b_don + b_off + b_hyd + (stage ? up(F * capacity * 12) + up(F * capacity * 8) : 0), &g);
    if (rc) return rc;
    uint32_t *d_don = (uint32_t *)g.out, *d_off = (uint32_t *)(g.out + b_don), *d_hyd = (uint32_t *)(g.out + b_don + b_off);
    uint32_t *d_dha = dha ? (stage ? (uint32_t *)(g.out + b_don + b_off + b_hyd) : dha) : nullptr;
    float *d_da = dha ? (stage ? (float *)(g.out + b_don + b_off + b_hyd + up(F * capacity * 12)) : dist_angle) : nullptr;
    if (n_donors) {
        CK(cudaMemcpyAsync(d_don, donors, n_donors * 4, cudaMemcpyDefault, ctx->compute));
        CK(cudaMemcpyAsync(d_off, hyd_offsets, (n_donors + 1) * 4, cudaMemcpyDefault, ctx->compute));
        if (n_hyd) CK(cudaMemcpyAsync(d_hyd, hydrogens, n_hyd * 4, cudaMemcpyDefault, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));  // the three host arrays may be pageable and short-lived
    }
    const float le2 = cutoff_squared_threshold_le(max_distance);
    const GroupView gacc = view_of(*acc);
    for (size_t f0 = 0; f0 < F; f0 += g.fb) {
        const size_t nf = std::min(g.fb, F - f0);
        FrameView fv = frames_of(ctx);
        fv.xyz += f0 * N * 3;
        fv.box += f0 * 9;
        rc = build_grid(ctx, fv, gacc, g, f0, nf);
        if (rc) return rc;
        if (n_donors && acc->n) {
            const unsigned nq = (unsigned)std::max<size_t>(1, std::min<size_t>((n_donors + 7) / 8, (size_t)kSMs * 16));
            k_hbonds<<<dim3(nq, (unsigned)nf), kThreads, 0, ctx->compute>>>(fv, gacc, g.cg, g.offsets, g.sorted, g.cells, d_don, d_off, d_hyd,
                                                                           (uint32_t)n_donors, le2, min_angle, g.d_count + f0,
                                                                           d_dha ? d_dha + f0 * capacity * 3 : nullptr,
                                                                           d_da ? d_da + f0 * capacity * 2 : nullptr, (unsigned long long)capacity,
                                                                           g.d_cursor + f0, g.d_far + f0);
            LAUNCHED();
        }
    }
    if ((rc = deliver(ctx, count, g.d_count, F * sizeof(uint64_t)))) return rc;
    if (stage && (rc = deliver(ctx, dha, d_dha, F * capacity * 12))) return rc;
    if (stage && (rc = deliver(ctx, dist_angle, d_da, F * capacity * 8))) return rc;
    return GROAN_OK;
}

}  // extern "C"

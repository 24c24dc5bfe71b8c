// groan_gpu.cu -- host side of libgroan_gpu.so: the C ABI declared in include/groan_gpu.h.
//
// One ctx = one GPU = one host thread.  Frames live in two device slots so that the host->device copy
// of batch k+1 (copy stream, pinned staging) overlaps the kernels of batch k (compute stream).
// Every op evaluates one groan_rs function for every frame of the current batch; host code here only
// validates arguments exactly like the reference does (error order included), picks launch
// geometry and delivers results.  There is no CPU fallback anywhere in this file.
#include "../../include/groan_gpu.h"

#include "ctx.cuh"
#include "kernels_center.cuh"
#include "kernels_rmsd.cuh"
#include "kernels_synth.cuh"
#include "kernels_tma.cuh"
#include "kernels_quad.cuh"

using namespace groan;
using namespace groan_host;
static_assert(kRefSums == kRefSumsHost, "ctx.cuh: Ref::sums");

namespace {

// blocks per frame for a streaming pass over g atoms of each of F frames
int blocks_per_frame(size_t g, size_t F) {
    size_t nb = (g + (size_t)kThreads * 8 - 1) / ((size_t)kThreads * 8);
    nb = std::max<size_t>(nb, 1);
    nb = std::min<size_t>(nb, kMaxBlocksPerFrame);
    nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / std::max<size_t>(F, 1)));
    return (int)nb;
}

// single-pass kernels: one wave of long-lived CTAs (occ per SM), so that the per-CTA reduction and the
// last-CTA finish are amortised over hundreds of atoms per thread
int blocks_per_frame_fast(size_t g, size_t F, int occ) {
    const size_t octs = (g + 7) / 8;
    size_t nb = (octs + (size_t)kThreads * 2 - 1) / ((size_t)kThreads * 2);
    nb = std::max<size_t>(nb, 1);
    nb = std::min<size_t>(nb, std::max<size_t>(1, ((size_t)kSMs * (size_t)occ) / std::max<size_t>(F, 1)));
    nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / std::max<size_t>(F, 1)));
    return (int)nb;
}

// The ring-fed kernels need a contiguous group (one byte range per frame), enough atoms to fill the ring, and a
// 16-byte aligned coordinate buffer (cudaMalloc'ed slots always are; attached buffers are checked)
bool tma_ok(const groan_gpu_ctx *ctx, const Group &g, int occ) {
    return occ > 0 && g.contiguous && g.n >= 4096 && (reinterpret_cast<uintptr_t>(ctx->cur_xyz) & 15) == 0 &&
           !(ctx->flags & GROAN_FLAG_NO_TMA);
}

// ---- quad kernels (kernels_quad.cuh) -----------------------------------------------------------------
// The centre kernels take any frame size (the 16-byte phase of the group is worked out per frame).  The RMSD kernels read a
// reference laid out relative to the aligned body, one copy per phase (launch_rmsd_quad_t).
// Triclinic extension: the single-pass centre runs on the ring as well (k_center_quad<., TRIC>); RMSD stays with the gather kernels.
bool quad_center_ok(const groan_gpu_ctx *ctx, const Group &g) { return tma_ok(ctx, g, 2); }
bool quad_ok(const groan_gpu_ctx *ctx, const Group &g) { return quad_center_ok(ctx, g) && !ctx->batch_tric; }

uint32_t quad_head(const groan_gpu_ctx *ctx, const Group &g, size_t f) { return (uint32_t)((4 - ((f * ctx->n_atoms + g.first) & 3)) & 3); }

int blocks_per_frame_quad(size_t g, size_t F, int occ, size_t chunk) {
    size_t nb = (g + chunk - 1) / chunk;  // at least one chunk per CTA
    nb = std::max<size_t>(nb, 1);
    nb = std::min<size_t>(nb, std::max<size_t>(1, ((size_t)kSMs * (size_t)occ) / std::max<size_t>(F, 1)));
    nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / std::max<size_t>(F, 1)));
    return (int)nb;
}

// build the permuted reference for every phase the frames of the batch put the group at
int ensure_quad_ref(groan_gpu_ctx *ctx, groan_gpu_ctx::Ref &R, const Group &g, QuadRef *out) {
    for (size_t f = 0; f < std::min<size_t>(ctx->n_frames, 4); f++) {
        const uint32_t head = quad_head(ctx, g, f);
        if (R.d_pq[head]) continue;
        const uint32_t body = ((uint32_t)g.n - head) & ~3u;
        const size_t bytes = quad_ref_floats(g.n) * sizeof(float);
        CK(cudaMalloc(&R.d_pq[head], bytes));
        CK(cudaMemsetAsync(R.d_pq[head], 0, bytes, ctx->compute));
        const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((body / 4 + kThreads - 1) / kThreads, (size_t)kSMs * 8));
        k_ref_permute<<<nb, kThreads, 0, ctx->compute>>>(R.d_pc, R.d_pq[head], head, body);
        LAUNCHED();
    }
    for (int h = 0; h < 4; h++) out->v[h] = R.d_pq[h];
    return GROAN_OK;
}

// The permuted reference is laid out relative to the 16-byte aligned body of the group, and where that body starts depends on
// the frame when a frame is not a multiple of four atoms long: frame f puts the group at phase (f * n_atoms + first) mod 4.
// All frames of one phase ("head") share one permuted copy.  Launching the whole batch at once would stream up to four
// 64 MB copies through the 126 MB L2 at the same time; instead the frames are sorted by phase and each phase gets its own
// launch (one wave of CTAs each), so that one copy at a time is resident.  n_atoms % 4 == 0 (or a single frame) is the
// one-launch case.
int quad_head_lists(groan_gpu_ctx *ctx, const Group &g, int counts[4], int offsets[4]) {
    const size_t F = ctx->n_frames;
    std::vector<int> order;
    order.reserve(F);
    int off = 0;
    for (int h = 0; h < 4; h++) {
        offsets[h] = off;
        for (size_t f = 0; f < F; f++)
            if (quad_head(ctx, g, f) == (uint32_t)h) order.push_back((int)f);
        counts[h] = (int)order.size() - off;
        off = (int)order.size();
    }
    if (ctx->head_list_key_frames != F || ctx->head_list_key_first != g.first) {
        if (!ctx->d_head_list) CK(cudaMalloc(&ctx->d_head_list, ctx->max_frames * sizeof(int)));
        CK(cudaMemcpyAsync(ctx->d_head_list, order.data(), F * sizeof(int), cudaMemcpyHostToDevice, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));  // `order` is a local; once per (batch size, group start)
        ctx->head_list_key_frames = F;
        ctx->head_list_key_first = g.first;
    }
    return GROAN_OK;
}

// The second tier of the fused centre + RMSD kernels (kernels_quad.cuh, finish_center_moments): the sine-sum centre pass over the
// frames the fused launch just counted, launched behind EVERY fused launch with room for kSecondCap frames (the CTAs of unused
// slots exit at once: ~4 us).  Launching it from the device only when needed costs ~45 us each time, and predicting the frames
// inside the fused kernel makes a whole one-wave grid wait for its slowest frame (profiles/r2_ring.md).
constexpr int kSecondCap = 4;
template <int CENTER>
int launch_second_tier(groan_gpu_ctx *ctx, const Group &g, float *d_center, FallbackPlan fp) {
    if (CENTER == 0 || !fp.enabled) return GROAN_OK;  // GROAN_FLAG_HOST_FALLBACK: rmsd_common gates the pass by the per-frame flags
    typedef QuadCfg<false, kQuadCenterStages, kQuadCenterThreads> C;
    fp.want_rmsd = 0;
    fp.want_center = 1;
    fp.feedback = nullptr;
    dim3 grid((unsigned)fp.nb_second, (unsigned)kSecondCap);
    k_center_quad<CENTER == 2><<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets,
                                                                                     d_center, ctx->d_flags, fp, ctx->d_second_list, 3, nullptr);
    LAUNCHED();
    return GROAN_OK;
}

template <bool SAME_MASS, int CENTER, bool TRIC = false>
int launch_rmsd_quad_t(groan_gpu_ctx *ctx, const Group &g, const RefView &rv, const QuadRef &d_pq, float *d_center, float *d_rmsd,
                       float *d_rot, FallbackPlan fp) {
    typedef QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads> C;
    if (ctx->n_atoms % 4 == 0 || ctx->n_frames == 1) {
        dim3 grid((unsigned)blocks_per_frame_quad(g.n, ctx->n_frames, 2, C::kAtoms), (unsigned)ctx->n_frames);
        k_rmsd_quad<SAME_MASS, CENTER, TRIC><<<grid, kQuadRmsdThreads, C::kBytes, ctx->compute>>>(TRIC ? frames_of_geom(ctx) : frames_of(ctx), view_of(g), rv, d_pq, ctx->d_partials,
                                                                                       ctx->d_tickets, d_center, d_rmsd, d_rot, ctx->d_cen,
                                                                                       ctx->d_flags, fp, nullptr);
        LAUNCHED();
        return launch_second_tier<CENTER>(ctx, g, d_center, fp);
    }
    int counts[4], offsets[4];
    int rc = quad_head_lists(ctx, g, counts, offsets);
    if (rc) return rc;
    for (int h = 0; h < 4; h++) {
        if (!counts[h]) continue;
        size_t nb = (size_t)blocks_per_frame_quad(g.n, (size_t)counts[h], 2, C::kAtoms);
        nb = std::max<size_t>(1, std::min<size_t>(nb, kPartialSlots / ctx->n_frames));  // partial records are indexed by the frame number
        dim3 grid((unsigned)nb, (unsigned)counts[h]);
        fp.n_report = counts[h];
        k_rmsd_quad<SAME_MASS, CENTER, TRIC><<<grid, kQuadRmsdThreads, C::kBytes, ctx->compute>>>(TRIC ? frames_of_geom(ctx) : frames_of(ctx), view_of(g), rv, d_pq, ctx->d_partials,
                                                                                       ctx->d_tickets, d_center, d_rmsd, d_rot, ctx->d_cen,
                                                                                       ctx->d_flags, fp, ctx->d_head_list + offsets[h]);
        LAUNCHED();
        rc = launch_second_tier<CENTER>(ctx, g, d_center, fp);
        if (rc) return rc;
    }
    return GROAN_OK;
}

int launch_rmsd_quad(groan_gpu_ctx *ctx, const Group &g, const RefView &rv, const QuadRef &d_pq, bool same_mass, int center_mode,
                     float *d_center, float *d_rmsd, float *d_rot, const FallbackPlan &fp) {
#define GO(SM, CM) return launch_rmsd_quad_t<SM, CM>(ctx, g, rv, d_pq, d_center, d_rmsd, d_rot, fp)
    if (center_mode == 3) { // triclinic extension: RMSD only, in the sheared picture (k_rmsd_quad<., 0, TRIC>)
        if (same_mass) return launch_rmsd_quad_t<true, 0, true>(ctx, g, rv, d_pq, d_center, d_rmsd, d_rot, fp);
        return launch_rmsd_quad_t<false, 0, true>(ctx, g, rv, d_pq, d_center, d_rmsd, d_rot, fp);
    }
    if (same_mass) {
        if (center_mode == 0) GO(true, 0);
        if (center_mode == 1) GO(true, 1);
        GO(true, 2);
    }
    // target masses != reference masses: only the RMSD-only kernel exists.  The fused variants would need 70 accumulator
    // registers and spill inside the loop (which costs ~40 %, profiles/exp/README.md); rmsd_common runs the centre separately.
    if (center_mode == 0) GO(false, 0);
    return GROAN_EINVAL;
#undef GO
}

template <bool SAME_MASS, int CENTER, bool TRIC = false>
int set_quad_attr(groan_gpu_ctx *ctx) {
    CK(cudaFuncSetAttribute(k_rmsd_quad<SAME_MASS, CENTER, TRIC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads>::kBytes));
    return GROAN_OK;
}

int blocks_per_frame_quad(size_t g, size_t F, int occ, size_t chunk);
// CTAs per frame of the exact quad passes (k_trig_quad / k_center_quad with ext_pilot; k_cov_quad).  They depend on the batch
// only, never on how many frames are re-done, so that a frame gets the same bits whether it is re-done from the device, from
// the host or under GROAN_FLAG_EXACT_ONLY.
int nb_exact_quad(const groan_gpu_ctx *ctx, const Group &g, bool cov) {
    const size_t chunk = 1024;
    size_t nb = std::max<size_t>(1, (g.n + chunk - 1) / chunk);
    nb = std::min<size_t>(nb, (size_t)kSMs * (cov ? 2 : (size_t)std::max(1, ctx->occ_center_quad)));
    nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / std::max<size_t>(ctx->n_frames, 1)));
    return (int)nb;
}

FallbackPlan fallback_plan(groan_gpu_ctx *ctx, const Group &g, bool want_center, bool center_weighted, float *center_out, bool want_rmsd,
                           float *rmsd_out, float *rot_out, bool quad_exact = false, const QuadRef *qr = nullptr) {
    FallbackPlan fp;
    const int slot = &g == &ctx->all ? GROAN_MAX_GROUPS : (int)(&g - ctx->groups);
    fp.feedback = (quad_exact && slot >= 0 && slot <= GROAN_MAX_GROUPS) ? ctx->d_feedback + slot : nullptr;
    fp.enabled = (ctx->flags & GROAN_FLAG_HOST_FALLBACK) ? 0 : 1;
    fp.n_frames = (int)ctx->n_frames;
    fp.nb_exact = blocks_per_frame_fast(g.n, ctx->n_frames, 4);
    fp.nb_cov = blocks_per_frame_fast(g.n, ctx->n_frames, 2);
    fp.frames_done = ctx->d_frames_done;
    fp.c0 = ctx->d_c0;
    fp.want_center = want_center;
    fp.center_weighted = center_weighted;
    fp.want_rmsd = want_rmsd;
    fp.center_out = center_out;
    fp.com = ctx->d_cen;
    fp.rmsd_out = rmsd_out;
    fp.rot_out = rot_out;
    fp.second_flags = ctx->d_flags2;
    fp.second_count = ctx->d_second_any;
    fp.second_ticket = ctx->d_second_any + 1;
    fp.second_cap = kSecondCap;
    fp.second_list = ctx->d_second_list;
    fp.n_report = (int)ctx->n_frames;
    // the second tier runs for one or two frames of a batch (the CTAs of its other slots exit at once): two CTAs per SM and frame
    {
        const size_t chunk = QuadCfg<false, kQuadCenterStages, kQuadCenterThreads>::kAtoms;
        size_t nb = std::max<size_t>(1, (g.n + chunk - 1) / chunk);
        nb = std::min<size_t>(nb, (size_t)kSMs * 2);
        nb = std::min<size_t>(nb, std::max<size_t>(1, kPartialSlots / std::max<size_t>(ctx->n_frames, 1)));
        fp.nb_second = (int)nb;
    }
    fp.second_smem = (int)QuadCfg<false, kQuadCenterStages, kQuadCenterThreads>::kBytes;
    fp.quad_exact = quad_exact ? 1 : 0;
    fp.slow_count = ctx->d_slow_count;
    fp.slow_list = ctx->d_slow_list;
    fp.nb_xc = nb_exact_quad(ctx, g, false);
    fp.nb_xv = nb_exact_quad(ctx, g, true);
    fp.cov_smem = (int)QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads>::kBytes;
    for (int h = 0; h < 4; h++) fp.ref_pq.v[h] = qr ? qr->v[h] : nullptr;
    return fp;
}

int ensure_slots(groan_gpu_ctx *ctx) {
    if (ctx->d_slot[0]) return GROAN_OK;
    const size_t bytes = ctx->max_frames * ctx->n_atoms * 3 * sizeof(float) + 256;
    for (int s = 0; s < 2; s++) CK(cudaMalloc(&ctx->d_slot[s], bytes));
    return GROAN_OK;
}

}  // namespace
namespace groan {
// first (frame, position in the group) whose atom has no position: min over f * n + i of the zero entries of the bitmap.
// mol_ref != nullptr restricts the scan to atoms of polyatomic molecules (make_molecules_whole, modifying.rs:362-380).
__global__ void __launch_bounds__(kThreads) k_first_invalid(const uint8_t *valid, size_t n_atoms, GroupView g, size_t n_frames,
                                                             const uint32_t *mol_ref, unsigned long long *first) {
    const size_t total = n_frames * (size_t)g.n;
    unsigned long long best = ~0ull;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) {
        const size_t f = k / g.n;
        const uint32_t a = g.atom((uint32_t)(k - f * g.n));
        if (mol_ref && __ldg(mol_ref + a) == kNoMolecule) continue;
        if (!__ldg(valid + f * n_atoms + a)) {
            best = k;
            break;  // k only grows along a thread's stride
        }
    }
    if (best != ~0ull) atomicMin(first, best);
}
}  // namespace groan

namespace groan_host {
int first_invalid(groan_gpu_ctx *ctx, const Group *gq, bool *found, size_t *frame, size_t *pos) {
    *found = false;
    if (!ctx->has_valid) return GROAN_OK;
    for (const auto &c : ctx->valid_cache)
        if (c.group == (const void *)gq) {
            *found = c.found;
            *frame = c.frame;
            *pos = c.atom;
            return GROAN_OK;
        }
    const Group &g = gq ? *gq : ctx->all;
    groan_gpu_ctx::ValidAnswer ans = {(const void *)gq, false, 0, 0};
    if (g.n) {
        if (!ctx->d_valid_first) CK(cudaMalloc(&ctx->d_valid_first, sizeof(unsigned long long)));
        CK(cudaMemsetAsync(ctx->d_valid_first, 0xff, sizeof(unsigned long long), ctx->compute));
        const size_t total = ctx->n_frames * g.n;
        const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((total + kThreads - 1) / kThreads, (size_t)kSMs * 16));
        k_first_invalid<<<nb, kThreads, 0, ctx->compute>>>(ctx->d_valid, ctx->n_atoms, view_of(g), ctx->n_frames,
                                                          gq ? nullptr : ctx->d_mol_ref, ctx->d_valid_first);
        LAUNCHED();
        unsigned long long h = ~0ull;
        CK(cudaMemcpyAsync(&h, ctx->d_valid_first, sizeof(h), cudaMemcpyDeviceToHost, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));
        if (h != ~0ull) {
            ans.found = true;
            ans.frame = (size_t)(h / g.n);
            ans.atom = (size_t)(h % g.n);
        }
    }
    ctx->valid_cache.push_back(ans);
    *found = ans.found;
    *frame = ans.frame;
    *pos = ans.atom;
    return GROAN_OK;
}

static size_t atom_of(const Group &g, size_t i) { return g.contiguous ? (size_t)g.first + i : (size_t)g.idx[i]; }

int check_positions(groan_gpu_ctx *ctx, const Group &g) {
    bool found;
    size_t f = 0, i = 0;
    int rc = first_invalid(ctx, &g, &found, &f, &i);
    if (rc) return rc;
    if (!found) return GROAN_OK;
    ctx->err_a = f;
    ctx->err_b = atom_of(g, i);
    return GROAN_ENOPOS;
}

// per frame the reference looks at a[0], then every atom of b, then a[1..] (atom.rs:780-790, row-major scan): with the first
// missing position of each group (frame-major, group order) the first failure of that scan follows
int check_pair_positions(groan_gpu_ctx *ctx, const Group &a, const Group &b) {
    if (!ctx->has_valid || a.n == 0 || b.n == 0) return GROAN_OK;
    bool fa, fb;
    size_t af = 0, ai = 0, bf = 0, bi = 0;
    int rc = first_invalid(ctx, &a, &fa, &af, &ai);
    if (rc) return rc;
    rc = first_invalid(ctx, &b, &fb, &bf, &bi);
    if (rc) return rc;
    if (!fa && !fb) return GROAN_OK;
    const bool take_a = fa && (!fb || af < bf || (af == bf && ai == 0));
    ctx->err_a = take_a ? af : bf;
    ctx->err_b = take_a ? atom_of(a, ai) : atom_of(b, bi);
    return GROAN_ENOPOS;
}

// switch to the other slot; the copy stream may only overwrite it once every kernel that used it is done
int begin_batch(groan_gpu_ctx *ctx, size_t F, const float *box, bool use_slot) {
    if (F == 0) return GROAN_EINVAL;
    if (F > ctx->max_frames) return GROAN_ECAPACITY;
    const int prev = ctx->slot, next = ctx->slot ^ 1;
    if (ctx->have_frames) {
        CK(cudaEventRecord(ctx->ev_done[prev], ctx->compute));
        ctx->done_recorded[prev] = true;
    }
    if (use_slot) {
        int rc = ensure_slots(ctx);
        if (rc) return rc;
        if (ctx->done_recorded[next]) CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_done[next], 0));
    }
    ctx->slot = next;
    ctx->n_frames = F;
    ctx->have_frames = true;
    ctx->has_valid = false;
    ctx->valid_cache.clear();
    ctx->have_box = (box != nullptr);
    ctx->h_box.assign(F * 9, 0.0f);
    if (box) {
        std::memcpy(ctx->h_box.data(), box, F * 9 * sizeof(float));
        // the box slot is small; order it after the kernels that still read the same slot's previous content
        if (ctx->done_recorded[next]) CK(cudaStreamWaitEvent(ctx->copy, ctx->ev_done[next], 0));
        CK(cudaMemcpyAsync(ctx->d_box[next], ctx->h_box.data(), F * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->copy));
    } else {
        CK(cudaMemsetAsync(ctx->d_box[next], 0, F * 9 * sizeof(float), ctx->copy));
    }
    return GROAN_OK;
}

int end_batch(groan_gpu_ctx *ctx) {
    CK(cudaEventRecord(ctx->ev_h2d, ctx->copy));
    CK(cudaStreamWaitEvent(ctx->compute, ctx->ev_h2d, 0));
    return GROAN_OK;
}
}  // namespace groan_host
namespace {

// ---- centre passes -------------------------------------------------------------------------------
// to_cartesian matters for triclinic batches only: intermediate results stay in the sheared picture (0), results go back (1)
int run_trig(groan_gpu_ctx *ctx, const Group &g, bool weighted, float *c0_out, const int *flags, int to_cartesian = 0) {
    const int nb = blocks_per_frame_fast(g.n, ctx->n_frames, 4);
    dim3 grid(nb, (unsigned)ctx->n_frames);
    if (weighted)
        k_trig<true><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, c0_out, flags,
                                                          to_cartesian);
    else
        k_trig<false><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, c0_out, flags,
                                                           to_cartesian);
    LAUNCHED();
    return GROAN_OK;
}

int run_unwrap(groan_gpu_ctx *ctx, const Group &g, bool weighted, const float *c0, float *out, const int *flags, int to_cartesian = 1) {
    const int nb = blocks_per_frame_fast(g.n, ctx->n_frames, 4);
    dim3 grid(nb, (unsigned)ctx->n_frames);
    if (weighted)
        k_unwrap<true><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), c0, ctx->d_partials, ctx->d_tickets, out, flags,
                                                            to_cartesian);
    else
        k_unwrap<false><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), c0, ctx->d_partials, ctx->d_tickets, out, flags,
                                                             to_cartesian);
    LAUNCHED();
    return GROAN_OK;
}

// A group whose every frame was flagged by the last single pass (a membrane: it spans the box on every frame) goes straight
// to the exact passes: no wasted pass, no device-side launch (~45 us per call).  The answer of the previous call arrives
// through a host-mapped word the last finishing thread writes; nothing is waited for, a stale answer only costs time.  Every
// 32nd call tries the single pass again.
bool skip_single_pass(groan_gpu_ctx *ctx, const Group &g) {
    if (ctx->flags & GROAN_FLAG_HOST_FALLBACK) return false;
    const int slot = &g == &ctx->all ? GROAN_MAX_GROUPS : (int)(&g - ctx->groups);
    if (slot < 0 || slot > GROAN_MAX_GROUPS) return false;
    const unsigned long long w = *(volatile unsigned long long *)(ctx->h_feedback + slot);
    const unsigned flagged = (unsigned)(w >> 32), frames = (unsigned)(w & 0xffffffffu);
    if (frames == 0 || flagged != frames) {
        ctx->fast_skips[slot] = 0;
        return false;
    }
    if (++ctx->fast_skips[slot] % 32 == 0) return false;
    return true;
}
int mark_all_flagged(groan_gpu_ctx *ctx) {  // groan_gpu_fallback_frames counts non-zero flags
    CK(cudaMemsetAsync(ctx->d_flags, 1, ctx->n_frames * sizeof(int), ctx->compute));
    return GROAN_OK;
}

// The exact passes of a contiguous group on the ring (kernels_quad.cuh).  sel == nullptr: every frame of the batch; otherwise
// the frames whose flag is set.
int run_exact_center_quad(groan_gpu_ctx *ctx, const Group &g, bool weighted, float *out, const int *sel) {
    typedef QuadCfg<false, kQuadCenterStages, kQuadCenterThreads> C;
    FallbackPlan off = fallback_plan(ctx, g, false, false, nullptr, false, nullptr, nullptr);
    off.enabled = 0;
    dim3 grid((unsigned)off.nb_xc, (unsigned)ctx->n_frames);
    const int mode = sel ? 1 : 0;
    k_trig_quad<<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, ctx->d_c0, sel, mode);
    LAUNCHED();
    if (weighted)
        k_center_quad<true><<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, out,
                                                                                   ctx->d_flags, off, sel, mode, ctx->d_c0);
    else
        k_center_quad<false><<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, out,
                                                                                    ctx->d_flags, off, sel, mode, ctx->d_c0);
    LAUNCHED();
    return GROAN_OK;
}

// group_get_com of the target (-> ctx->d_cen), then the f64 covariance around it
int run_exact_rmsd_quad(groan_gpu_ctx *ctx, const Group &g, const RefView &rv, const QuadRef &qr, float *d_rmsd, float *d_rot, const int *sel) {
    int rc = run_exact_center_quad(ctx, g, true, ctx->d_cen, sel);
    if (rc) return rc;
    typedef QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads> C;
    dim3 grid((unsigned)nb_exact_quad(ctx, g, true), (unsigned)ctx->n_frames);
    k_cov_quad<<<grid, kQuadRmsdThreads, C::kBytes, ctx->compute>>>(frames_of(ctx), view_of(g), rv, qr, ctx->d_cen, ctx->d_partials, ctx->d_tickets,
                                                                     d_rmsd, d_rot, sel, sel ? 1 : 0);
    LAUNCHED();
    return GROAN_OK;
}

// group_get_center / group_get_com.  Single pass first (kernels_center.cuh); the reference-order passes --
// estimate (always geometric, iterators.rs:1407) then unwrap -- then re-do only the frames it flagged
// (their CTAs exit at once for every other frame).  GROAN_FLAG_EXACT_ONLY runs the reference-order passes alone.
int run_get_center(groan_gpu_ctx *ctx, const Group &g, bool weighted, float *out) {
    const int *flags = nullptr;
    if (ctx->occ_center_quad > 0 && quad_center_ok(ctx, g) && ctx->batch_tric) {
        // triclinic extension: the single pass on the ring in the sheared picture, then -- like behind the gather kernel -- the
        // reference-order passes over the frames it flagged (they exit at once for every other frame)
        if (!(ctx->flags & GROAN_FLAG_EXACT_ONLY)) {
            typedef QuadCfg<false, kQuadCenterStages, kQuadCenterThreads> C;
            dim3 grid(blocks_per_frame_quad(g.n, ctx->n_frames, ctx->occ_center_quad, C::kAtoms), (unsigned)ctx->n_frames);
            FallbackPlan fp = fallback_plan(ctx, g, true, weighted, out, false, nullptr, nullptr, false);
            fp.enabled = 0;
            if (weighted)
                k_center_quad<true, true><<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials,
                                                                                              ctx->d_tickets, out, ctx->d_flags, fp, nullptr, 0, nullptr);
            else
                k_center_quad<false, true><<<grid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials,
                                                                                               ctx->d_tickets, out, ctx->d_flags, fp, nullptr, 0, nullptr);
            LAUNCHED();
            flags = ctx->d_flags;
        }
    } else if (ctx->occ_center_quad > 0 && quad_center_ok(ctx, g)) {
        if (ctx->flags & GROAN_FLAG_EXACT_ONLY) return run_exact_center_quad(ctx, g, weighted, out, nullptr);
        if (skip_single_pass(ctx, g)) {
            int rc = mark_all_flagged(ctx);
            return rc ? rc : run_exact_center_quad(ctx, g, weighted, out, nullptr);
        }
        typedef QuadCfg<false, kQuadCenterStages, kQuadCenterThreads> C;
        dim3 grid(blocks_per_frame_quad(g.n, ctx->n_frames, ctx->occ_center_quad, C::kAtoms), (unsigned)ctx->n_frames);
        const size_t smem = C::kBytes;
        const FallbackPlan fp = fallback_plan(ctx, g, true, weighted, out, false, nullptr, nullptr, true);
        if (weighted)
            k_center_quad<true><<<grid, kQuadCenterThreads, smem, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets,
                                                                           out, ctx->d_flags, fp, nullptr, 0, nullptr);
        else
            k_center_quad<false><<<grid, kQuadCenterThreads, smem, ctx->compute>>>(frames_of(ctx), view_of(g), ctx->d_partials, ctx->d_tickets,
                                                                            out, ctx->d_flags, fp, nullptr, 0, nullptr);
        LAUNCHED();
        if (fp.enabled) return GROAN_OK;
        return run_exact_center_quad(ctx, g, weighted, out, ctx->d_flags);
    } else if (!(ctx->flags & GROAN_FLAG_EXACT_ONLY)) {
        const int nb = blocks_per_frame_fast(g.n, ctx->n_frames, ctx->occ_center);
        dim3 grid(nb, (unsigned)ctx->n_frames);
        if (weighted)
            k_center_fast<true><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, out,
                                                                     ctx->d_flags);
        else
            k_center_fast<false><<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(g), ctx->d_partials, ctx->d_tickets, out,
                                                                      ctx->d_flags);
        LAUNCHED();
        flags = ctx->d_flags;
    }
    int rc = run_trig(ctx, g, false, ctx->d_c0, flags);
    if (rc) return rc;
    return run_unwrap(ctx, g, weighted, ctx->d_c0, out, flags);
}

// validation shared by the centre ops, in the reference's order (analysis.rs:105-120, iterators.rs:1152-1191)
int check_center_args(groan_gpu_ctx *ctx, int gid, bool weighted, const Group **gp, bool allow_triclinic = false) {
    if (!ctx) return GROAN_EINVAL;
    const Group *g = get_group(ctx, gid);
    if (!g) return GROAN_ENOGROUP;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    if (g->n == 0) return GROAN_EEMPTY;
    int rc = check_box(ctx, allow_triclinic, nullptr);  // triclinic boxes: extension of the centre ops only (GROAN_FLAG_TRICLINIC)
    if (rc) return rc;
    rc = check_positions(ctx, *g);
    if (rc) return rc;
    if (weighted) {
        rc = check_masses(ctx, *g);
        if (rc) return rc;
    }
    *gp = g;
    return GROAN_OK;
}
int run_wrap(groan_gpu_ctx *ctx, const Group &g, bool translate, const float t[3], int8_t *shifts, bool tric) {
    int8_t *d_sh = nullptr;
    const size_t sh_bytes = ctx->n_frames * g.n * 3;
    if (shifts) {
        if (classify(shifts) == PK_DEVICE) {
            d_sh = shifts;
        } else {
            int rc = ensure_tmp(ctx, sh_bytes);
            if (rc) return rc;
            d_sh = (int8_t *)ctx->d_tmp;
        }
    }
    size_t nb = (g.n + kThreads - 1) / kThreads;
    nb = std::max<size_t>(1, std::min<size_t>(nb, (size_t)kMaxBlocksPerFrame * 4));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    const float tx = translate ? t[0] : 0.0f, ty = translate ? t[1] : 0.0f, tz = translate ? t[2] : 0.0f;
    float *xyz = ctx->cur_xyz;
    const float *box = ctx->d_box[ctx->slot];
    const GroupView gv = view_of(g);
#define WRAP_LAUNCH(K, TR, SH) K<TR, SH><<<grid, kThreads, 0, ctx->compute>>>(xyz, box, ctx->n_atoms, gv, tx, ty, tz, d_sh)
    if (tric) {
        if (translate) { if (d_sh) WRAP_LAUNCH(k_wrap_tric, true, true); else WRAP_LAUNCH(k_wrap_tric, true, false); }
        else { if (d_sh) WRAP_LAUNCH(k_wrap_tric, false, true); else WRAP_LAUNCH(k_wrap_tric, false, false); }
    } else {
        if (translate) { if (d_sh) WRAP_LAUNCH(k_wrap, true, true); else WRAP_LAUNCH(k_wrap, true, false); }
        else { if (d_sh) WRAP_LAUNCH(k_wrap, false, true); else WRAP_LAUNCH(k_wrap, false, false); }
    }
#undef WRAP_LAUNCH
    LAUNCHED();
    if (shifts && d_sh != shifts) return deliver(ctx, shifts, d_sh, sh_bytes);
    return GROAN_OK;
}

int check_wrap_args(groan_gpu_ctx *ctx, int gid, const Group **gp, bool *tric) {
    if (!ctx) return GROAN_EINVAL;
    const Group *g = get_group(ctx, gid);
    if (!g) return GROAN_ENOGROUP;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    int rc = check_box(ctx, true, tric);
    if (rc) return rc;
    rc = check_positions(ctx, *g);
    if (rc) return rc;
    *gp = g;
    return GROAN_OK;
}

// center != nullptr: also produce group_get_center (center_weighted = 0) or group_get_com (1) of the same group
int rmsd_common(groan_gpu_ctx *ctx, int gid, float *rmsd, float *rot, bool fit, float *center = nullptr, int center_weighted = 0) {
    if (!ctx) return GROAN_EINVAL;
    const Group *g = get_group(ctx, gid);
    if (!g) return GROAN_ENOGROUP;
    if (!ctx->refs[ref_slot(gid)].set) return GROAN_ENOREF;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    // extract_data_from_system(target): box first (rmsd.rs:430), then group_get_com's own checks
    bool tric = false;
    int rc = check_box(ctx, !fit, &tric);  // triclinic extension: RMSD and rotation, not the fit (fit_structure stays orthogonal)
    if (rc) return rc;
    if (g->n == 0) return GROAN_EEMPTY;
    rc = check_positions(ctx, *g);
    if (rc) return rc;
    rc = check_masses(ctx, *g);
    if (rc) return rc;
    groan_gpu_ctx::Ref &R = ctx->refs[ref_slot(gid)];
    if (R.n != g->n) {  // number_of_positions_consistent, rmsd.rs:405-422
        ctx->err_a = R.n;
        ctx->err_b = g->n;
        return GROAN_EGROUPSIZE;
    }
    float *d_rmsd = target_of<float>(rmsd, ctx->d_res);
    float *d_rot = target_of<float>(rot, ctx->d_rot);
    RefView rv;
    rv.pc = R.d_pc;
    rv.sum_wpp = R.sums[0];
    rv.sum_w = R.sums[1];
    for (int k = 0; k < 3; k++) {
        rv.sum_pc[k] = R.sums[2 + k];
        rv.sum_wpc[k] = R.sums[5 + k];
        rv.com[k] = R.com[k];
    }
    rv.sum_m_target = g->mass_sum;
    const int *flags = nullptr;
    float *d_center = center ? target_of<float>(center, ctx->d_cen2) : nullptr;
    bool center_done = false, device_fallback = false, second_tier = false;
    const bool qx = quad_ok(ctx, *g);  // contiguous group on the ring: fast and exact passes of kernels_quad.cuh
    // triclinic extension: the single pass on the ring in the sheared picture; the exact passes stay with the gather kernels
    const bool qt = !qx && ctx->batch_tric && quad_center_ok(ctx, *g) && !(ctx->flags & GROAN_FLAG_EXACT_ONLY);
    QuadRef qr;
    if (qx || qt) {
        rc = ensure_quad_ref(ctx, R, *g, &qr);
        if (rc) return rc;
    }
    const bool skip = qx && !(ctx->flags & GROAN_FLAG_EXACT_ONLY) && skip_single_pass(ctx, *g);
    if (skip) {
        rc = mark_all_flagged(ctx);  // flags stays nullptr: the exact passes below run for every frame
        if (rc) return rc;
    } else if (!(ctx->flags & GROAN_FLAG_EXACT_ONLY) && qx) {
        // quad kernels: RMSD, optionally with the centre, from one read of the frame (kernels_quad.cuh)
        const bool fused = center != nullptr && R.same_mass;  // see launch_rmsd_quad
        const FallbackPlan fp = fallback_plan(ctx, *g, fused, center_weighted != 0, d_center, true, d_rmsd, d_rot, true, &qr);
        device_fallback = fp.enabled != 0;
        rc = launch_rmsd_quad(ctx, *g, rv, qr, R.same_mass, fused ? (center_weighted ? 2 : 1) : 0, d_center, d_rmsd, d_rot, fp);
        if (rc) return rc;
        flags = ctx->d_flags;
        center_done = fused;
        second_tier = fused;
    } else if (qt) {
        FallbackPlan fp = fallback_plan(ctx, *g, false, false, nullptr, true, d_rmsd, d_rot, false, &qr);
        fp.enabled = 0;  // flagged frames: the host-launched reference-order passes below, gated by the flags
        rc = launch_rmsd_quad(ctx, *g, rv, qr, R.same_mass, 3, nullptr, d_rmsd, d_rot, fp);
        if (rc) return rc;
        flags = ctx->d_flags;
    } else if (!(ctx->flags & GROAN_FLAG_EXACT_ONLY)) {
        // single pass: COM, covariance and RMSD sums relative to a pilot atom (kernels_rmsd.cuh)
        const int nbf = blocks_per_frame_fast(g->n, ctx->n_frames, ctx->occ_rmsd);
        dim3 fgrid(nbf, (unsigned)ctx->n_frames);
        if (center && R.same_mass && !tric) {
            // centre and RMSD from one gather of the group (kernels_quad.cuh k_rmsd_fast_center)
            if (center_weighted)
                k_rmsd_fast_center<2><<<fgrid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(*g), rv, ctx->d_partials, ctx->d_tickets,
                                                                             d_center, d_rmsd, d_rot, ctx->d_cen, ctx->d_flags);
            else
                k_rmsd_fast_center<1><<<fgrid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(*g), rv, ctx->d_partials, ctx->d_tickets,
                                                                             d_center, d_rmsd, d_rot, ctx->d_cen, ctx->d_flags);
            center_done = true;
        } else if (R.same_mass)
            k_rmsd_fast<true><<<fgrid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(*g), rv, ctx->d_partials, ctx->d_tickets,
                                                                     d_rmsd, d_rot, ctx->d_cen, ctx->d_flags);
        else
            k_rmsd_fast<false><<<fgrid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(*g), rv, ctx->d_partials, ctx->d_tickets,
                                                                      d_rmsd, d_rot, ctx->d_cen, ctx->d_flags);
        LAUNCHED();
        flags = ctx->d_flags;
    }
    // reference-order passes (all frames, or only the flagged ones): group_get_com of the target
    // (geometric estimate, mass-weighted unwrap), then shift + wrap + covariance in f64.  Skipped here when the
    // single-pass kernel tail-launches them itself for the frames that need them.
    ctx->second_valid = second_tier;
    if (!device_fallback && qx) {
        rc = run_exact_rmsd_quad(ctx, *g, rv, qr, d_rmsd, d_rot, flags);
        if (rc) return rc;
    } else if (!device_fallback) {
        rc = run_trig(ctx, *g, false, ctx->d_c0, flags);
        if (rc) return rc;
        rc = run_unwrap(ctx, *g, true, ctx->d_c0, ctx->d_cen, flags, 0);  // the COM stays in the picture k_cov wraps in
        if (rc) return rc;
        const int nb = blocks_per_frame_fast(g->n, ctx->n_frames, 2);
        dim3 grid(nb, (unsigned)ctx->n_frames);
        k_cov<<<grid, kThreads, 0, ctx->compute>>>(frames_of_geom(ctx), view_of(*g), rv, ctx->d_cen, ctx->d_partials, ctx->d_tickets, d_rmsd,
                                                    d_rot, flags);
        LAUNCHED();
    }
    if (fit) {
        size_t fb = (ctx->n_atoms + kThreads - 1) / kThreads;
        fb = std::max<size_t>(1, std::min<size_t>(fb, (size_t)kMaxBlocksPerFrame * 4));
        dim3 fgrid((unsigned)fb, (unsigned)ctx->n_frames);
        k_fit<<<fgrid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->d_box[ctx->slot], ctx->n_atoms, ctx->d_cen, d_rot, R.com[0],
                                                     R.com[1], R.com[2]);
        LAUNCHED();
    }
    if (center) {
        if (center_done) {
            // frames the fused pass flagged: reference-order centre passes for those frames only (d_c0 already
            // holds their Bai-Breen estimate from the RMSD fallback above)
            if (!device_fallback) {
                if (ctx->second_valid) {
                    // host-launched second tier of the fused quad kernel: the sine-sum centre pass over the frames it marked;
                    // what that pass cannot certify either gets bit 1 of its flag set and is redone below
                    typedef QuadCfg<false, kQuadCenterStages, kQuadCenterThreads> C;
                    FallbackPlan f2 = fallback_plan(ctx, *g, true, center_weighted != 0, d_center, false, nullptr, nullptr);
                    f2.second_count = nullptr;
                    dim3 sgrid((unsigned)f2.nb_second, (unsigned)ctx->n_frames);
                    if (center_weighted)
                        k_center_quad<true><<<sgrid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(
                            frames_of(ctx), view_of(*g), ctx->d_partials, ctx->d_tickets, d_center, ctx->d_flags, f2, ctx->d_flags2, 1, nullptr);
                    else
                        k_center_quad<false><<<sgrid, kQuadCenterThreads, C::kBytes, ctx->compute>>>(
                            frames_of(ctx), view_of(*g), ctx->d_partials, ctx->d_tickets, d_center, ctx->d_flags, f2, ctx->d_flags2, 1, nullptr);
                    LAUNCHED();
                }
                if (qx) {
                    rc = run_exact_center_quad(ctx, *g, center_weighted != 0, d_center, flags);
                } else {
                    rc = run_unwrap(ctx, *g, center_weighted != 0, ctx->d_c0, d_center, flags);
                }
                if (rc) return rc;
            }
        } else {
            rc = run_get_center(ctx, *g, center_weighted != 0, d_center);
            if (rc) return rc;
        }
        rc = deliver(ctx, center, d_center, ctx->n_frames * 3 * sizeof(float));
        if (rc) return rc;
    }
    rc = deliver(ctx, rmsd, d_rmsd, ctx->n_frames * sizeof(float));
    if (rc) return rc;
    return deliver(ctx, rot, d_rot, ctx->n_frames * 9 * sizeof(float));
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int groan_gpu_create(int device, size_t n_atoms, size_t max_frames, groan_gpu_ctx **out) {
    if (!out || n_atoms == 0 || max_frames == 0 || n_atoms > 0xFFFFFFF0ull) return GROAN_EINVAL;
    *out = nullptr;
    // frames map to gridDim.y (<= 65535), and maybe_launch_fallback counts finished frames in 16 bits
    if (max_frames > kMaxFramesPerBatch) return GROAN_ECAPACITY;
    groan_gpu_ctx *ctx = new (std::nothrow) groan_gpu_ctx();
    if (!ctx) return GROAN_EINVAL;
    ctx->device = device;
    ctx->n_atoms = n_atoms;
    ctx->max_frames = max_frames;
    int rc = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&ctx->own_compute, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking));
        ctx->compute = ctx->own_compute;
        CK(cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming));
        for (int s = 0; s < 2; s++) {
            CK(cudaEventCreateWithFlags(&ctx->ev_done[s], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&ctx->ev_stage[s], cudaEventDisableTiming));
            CK(cudaMalloc(&ctx->d_box[s], max_frames * 9 * sizeof(float)));
        }
        const size_t slots = std::max(max_frames, kPartialSlots) + kMaxBlocksPerFrame;
        CK(cudaMalloc(&ctx->d_partials, slots * kMaxSums * sizeof(double)));
        CK(cudaMalloc(&ctx->d_pair_partials, slots * kPairPartialBytes));
        CK(cudaMalloc(&ctx->d_tickets, (max_frames + 1) * sizeof(unsigned int)));
        CK(cudaMemset(ctx->d_tickets, 0, (max_frames + 1) * sizeof(unsigned int)));
        CK(cudaMalloc(&ctx->d_c0, max_frames * 3 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_cen, max_frames * 3 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_cen2, max_frames * 3 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_res, max_frames * 8 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_rot, max_frames * 9 * sizeof(float)));
        CK(cudaMalloc(&ctx->d_frames_done, sizeof(unsigned int)));
        CK(cudaMemset(ctx->d_frames_done, 0, sizeof(unsigned int)));
        CK(cudaMalloc(&ctx->d_flags, max_frames * sizeof(int)));
        CK(cudaMemset(ctx->d_flags, 0, max_frames * sizeof(int)));
        CK(cudaMalloc(&ctx->d_flags2, max_frames * sizeof(int)));
        CK(cudaMemset(ctx->d_flags2, 0, max_frames * sizeof(int)));
        CK(cudaMalloc(&ctx->d_second_list, max_frames * sizeof(int)));
        CK(cudaMalloc(&ctx->d_slow_list, max_frames * sizeof(int)));
        CK(cudaHostAlloc(&ctx->h_feedback, (GROAN_MAX_GROUPS + 1) * sizeof(unsigned long long), cudaHostAllocMapped));
        std::memset(ctx->h_feedback, 0, (GROAN_MAX_GROUPS + 1) * sizeof(unsigned long long));
        CK(cudaHostGetDevicePointer(&ctx->d_feedback, ctx->h_feedback, 0));
        CK(cudaMalloc(&ctx->d_slow_count, sizeof(unsigned int)));
        CK(cudaMemset(ctx->d_slow_count, 0, sizeof(unsigned int)));
        CK(cudaMalloc(&ctx->d_second_any, 2 * sizeof(unsigned int)));  // [frames counted by the fused launch, tickets of the pass]
        CK(cudaMemset(ctx->d_second_any, 0, 2 * sizeof(unsigned int)));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_center, k_center_fast<false>, kThreads, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_rmsd, k_rmsd_fast<true>, kThreads, 0));
        ctx->occ_center = std::max(1, std::min(ctx->occ_center, 8));
        ctx->occ_rmsd = std::max(1, std::min(ctx->occ_rmsd, 8));
        const int sq = (int)QuadCfg<false, kQuadCenterStages, kQuadCenterThreads>::kBytes;
        CK(cudaFuncSetAttribute(k_center_quad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sq));
        CK(cudaFuncSetAttribute(k_center_quad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sq));
        CK(cudaFuncSetAttribute(k_center_quad<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sq));
        CK(cudaFuncSetAttribute(k_center_quad<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sq));
        CK(cudaFuncSetAttribute(k_trig_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, sq));
        CK(cudaFuncSetAttribute(k_cov_quad, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads>::kBytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->occ_center_quad, k_center_quad<false>, kQuadCenterThreads, sq));
        ctx->occ_center_quad = std::min(ctx->occ_center_quad, 4);
        int qrc = set_quad_attr<true, 0>(ctx);
        if (!qrc) qrc = set_quad_attr<true, 1>(ctx);
        if (!qrc) qrc = set_quad_attr<true, 2>(ctx);
        if (!qrc) qrc = set_quad_attr<false, 0>(ctx);
        if (!qrc) qrc = set_quad_attr<true, 0, true>(ctx);
        if (!qrc) qrc = set_quad_attr<false, 0, true>(ctx);
        if (qrc) return qrc;
        return GROAN_OK;
    }();
    if (rc) {
        groan_gpu_destroy(ctx);
        return rc;
    }
    ctx->all.set = true;
    ctx->all.contiguous = true;
    ctx->all.first = 0;
    ctx->all.n = n_atoms;
    *out = ctx;
    return GROAN_OK;
}

void groan_gpu_destroy(groan_gpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; s++) {
        if (ctx->d_slot[s]) cudaFree(ctx->d_slot[s]);
        if (ctx->d_box[s]) cudaFree(ctx->d_box[s]);
        if (ctx->h_stage[s]) cudaFreeHost(ctx->h_stage[s]);
        if (ctx->d_quant[s]) cudaFree(ctx->d_quant[s]);
        if (ctx->d_origin[s]) cudaFree(ctx->d_origin[s]);
        if (ctx->d_xtc[s]) cudaFree(ctx->d_xtc[s]);
        if (ctx->d_xtc_params[s]) cudaFree(ctx->d_xtc_params[s]);
        if (ctx->d_sel[s]) cudaFree(ctx->d_sel[s]);
        if (ctx->ev_done[s]) cudaEventDestroy(ctx->ev_done[s]);
        if (ctx->ev_stage[s]) cudaEventDestroy(ctx->ev_stage[s]);
    }
    for (auto &g : ctx->groups) {
        if (g.d_idx) cudaFree(g.d_idx);
        if (g.d_mass) cudaFree(g.d_mass);
    }
    if (ctx->all.d_mass) cudaFree(ctx->all.d_mass);
    for (auto &r : ctx->refs) {
        if (r.d_pc) cudaFree(r.d_pc);
        for (float *q : r.d_pq)
            if (q) cudaFree(q);
    }
    void *bufs[] = {ctx->d_partials, ctx->d_pair_partials, ctx->d_tickets, ctx->d_c0, ctx->d_cen, ctx->d_cen2, ctx->d_res, ctx->d_rot, ctx->d_tmp, ctx->d_flags, ctx->d_flags2, ctx->d_second_any, ctx->d_second_list, ctx->d_slow_list, ctx->d_slow_count, ctx->d_head_list, ctx->d_valid, ctx->d_valid_first, ctx->d_frames_done, ctx->d_mol_ref,
                    ctx->d_xtc_status, ctx->d_sel_atoms};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (ctx->h_feedback) cudaFreeHost(ctx->h_feedback);
    if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
    if (ctx->own_compute) cudaStreamDestroy(ctx->own_compute);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    cudaGetLastError();
    delete ctx;
}

int groan_gpu_set_flags(groan_gpu_ctx *ctx, unsigned flags) {
    if (!ctx) return GROAN_EINVAL;
    ctx->flags = flags;
    return GROAN_OK;
}

int groan_gpu_set_stream(groan_gpu_ctx *ctx, void *cuda_stream) {
    if (!ctx) return GROAN_EINVAL;
    CK(cudaStreamSynchronize(ctx->compute));
    ctx->compute = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_compute;
    return GROAN_OK;
}

int groan_gpu_sync(groan_gpu_ctx *ctx) {
    if (!ctx) return GROAN_EINVAL;
    CK(cudaStreamSynchronize(ctx->copy));
    CK(cudaStreamSynchronize(ctx->compute));
    return GROAN_OK;
}

const char *groan_gpu_strerror(int status) {
    switch (status) {
    case GROAN_OK: return "ok";
    case GROAN_ENOBOX: return "system has no simulation box";
    case GROAN_ENOTORTHO: return "simulation box is not orthogonal";
    case GROAN_EEMPTY: return "group is empty";
    case GROAN_ENOPOS: return "atom has no position";
    case GROAN_ENOMASS: return "atom has no mass";
    case GROAN_EGROUPSIZE: return "group has an inconsistent number of atoms in reference and target";
    case GROAN_EZEROBOX: return "box length is zero";
    case GROAN_ENOGROUP: return "group does not exist";
    case GROAN_EINVAL: return "invalid argument";
    case GROAN_ECUDA: return "CUDA runtime failure";
    case GROAN_ENOFRAMES: return "no frames pushed or attached";
    case GROAN_ENOREF: return "RMSD reference not set";
    case GROAN_ECAPACITY: return "more frames than the ctx was created for";
    default: return "unknown status";
    }
}

const char *groan_gpu_last_cuda_error(groan_gpu_ctx *ctx) { return ctx ? ctx->cuda_err.c_str() : ""; }

int groan_gpu_error_detail(groan_gpu_ctx *ctx, size_t *a, size_t *b) {
    if (!ctx) return GROAN_EINVAL;
    if (a) *a = ctx->err_a;
    if (b) *b = ctx->err_b;
    return GROAN_OK;
}

uint64_t groan_gpu_launch_count(groan_gpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int groan_gpu_fallback_frames(groan_gpu_ctx *ctx, size_t *n) {
    if (!ctx || !n) return GROAN_EINVAL;
    *n = 0;
    if (!ctx->have_frames) return GROAN_OK;
    std::vector<int> h(ctx->n_frames);
    CK(cudaMemcpyAsync(h.data(), ctx->d_flags, ctx->n_frames * sizeof(int), cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    for (int v : h) *n += (v != 0);
    return GROAN_OK;
}

int groan_gpu_second_pass_frames(groan_gpu_ctx *ctx, size_t *n) {
    if (!ctx || !n) return GROAN_EINVAL;
    *n = 0;
    if (!ctx->have_frames || !ctx->second_valid) return GROAN_OK;
    std::vector<int> h(ctx->n_frames);
    CK(cudaMemcpyAsync(h.data(), ctx->d_flags2, ctx->n_frames * sizeof(int), cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    for (int v : h) *n += (v != 0);
    return GROAN_OK;
}

// ---- groups ---------------------------------------------------------------------------------------
int groan_gpu_set_group(groan_gpu_ctx *ctx, int gid, const uint32_t *idx, size_t n, const float *mass) {
    if (ctx && gid == GROAN_GROUP_ALL) {
        // the built-in "all" group (System::new creates "all" / "All" over every atom, groups.rs): its index list is fixed,
        // only the masses can be attached (or removed: mass == NULL)
        if (idx || n != ctx->n_atoms) return GROAN_EINVAL;
        Group &g = ctx->all;
        CK(cudaStreamSynchronize(ctx->compute));
        if (g.d_mass) { cudaFree(g.d_mass); g.d_mass = nullptr; }
        groan_gpu_ctx::Ref &R = ctx->refs[GROAN_MAX_GROUPS];
        if (R.d_pc) { cudaFree(R.d_pc); R.d_pc = nullptr; }
        for (float *&q : R.d_pq)
            if (q) { cudaFree(q); q = nullptr; }
        R.set = false;
        g.has_mass = (mass != nullptr);
        g.no_mass_at = -1;
        g.mass.clear();
        g.mass_sum = 0.0;
        if (mass) {
            g.mass.assign(mass, mass + n);
            for (size_t i = 0; i < n; i++) {
                g.mass_sum += (double)mass[i];
                if (mass[i] < 0.0f && g.no_mass_at < 0) g.no_mass_at = (long)i;
            }
            CK(cudaMalloc(&g.d_mass, n * sizeof(float)));
            CK(cudaMemcpy(g.d_mass, mass, n * sizeof(float), cudaMemcpyHostToDevice));
        }
        return GROAN_OK;
    }
    if (!ctx || gid < 0 || gid >= GROAN_MAX_GROUPS || (n && !idx)) return GROAN_EINVAL;
    for (size_t i = 0; i < n; i++) {
        if (idx[i] >= ctx->n_atoms) return GROAN_EINVAL;
        if (i && idx[i] <= idx[i - 1]) return GROAN_EINVAL;  // container.rs:51-115: sorted, unique
    }
    Group &g = ctx->groups[gid];
    ctx->valid_cache.clear();
    CK(cudaStreamSynchronize(ctx->compute));
    if (g.d_idx) { cudaFree(g.d_idx); g.d_idx = nullptr; }
    if (g.d_mass) { cudaFree(g.d_mass); g.d_mass = nullptr; }
    if (ctx->refs[gid].d_pc) { cudaFree(ctx->refs[gid].d_pc); ctx->refs[gid].d_pc = nullptr; }
    for (float *&q : ctx->refs[gid].d_pq)
        if (q) { cudaFree(q); q = nullptr; }
    ctx->refs[gid].set = false;
    g.set = true;
    g.n = n;
    g.idx.assign(idx, idx + n);
    g.contiguous = true;
    g.first = n ? idx[0] : 0;
    for (size_t i = 0; i < n; i++)
        if (idx[i] != g.first + i) { g.contiguous = false; break; }
    g.has_mass = (mass != nullptr);
    g.no_mass_at = -1;
    g.mass.clear();
    if (mass) g.mass.assign(mass, mass + n);
    g.mass_sum = 0.0;
    for (float m : g.mass) g.mass_sum += (double)m;
    if (n) {
        CK(cudaMalloc(&g.d_idx, n * sizeof(uint32_t)));
        CK(cudaMemcpy(g.d_idx, idx, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (mass) {
            for (size_t i = 0; i < n; i++)
                if (mass[i] < 0.0f) { g.no_mass_at = (long)i; break; }
            CK(cudaMalloc(&g.d_mass, n * sizeof(float)));
            CK(cudaMemcpy(g.d_mass, mass, n * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    if (g.contiguous) g.idx.clear();
    return GROAN_OK;
}

// ---- frames ---------------------------------------------------------------------------------------
}  // extern "C"

namespace groan_host {
// host -> device on the copy stream; a pageable source bounces through two pinned chunks so that the host memcpy of
// chunk i + 1 overlaps the DMA of chunk i
int h2d_on_copy_stream(groan_gpu_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (classify(src) != PK_PAGEABLE) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->copy));
        return GROAN_OK;
    }
    for (int s = 0; s < 2; s++)
        if (!ctx->h_stage[s]) CK(cudaMallocHost(&ctx->h_stage[s], kStageBytes));
    size_t off = 0;
    int s = 0;
    while (off < bytes) {
        const size_t chunk = std::min(kStageBytes, bytes - off);
        CK(cudaEventSynchronize(ctx->ev_stage[s]));
        std::memcpy(ctx->h_stage[s], reinterpret_cast<const char *>(src) + off, chunk);
        CK(cudaMemcpyAsync(reinterpret_cast<char *>(dst) + off, ctx->h_stage[s], chunk, cudaMemcpyHostToDevice, ctx->copy));
        CK(cudaEventRecord(ctx->ev_stage[s], ctx->copy));
        off += chunk;
        s ^= 1;
    }
    return GROAN_OK;
}
}  // namespace groan_host

extern "C" {

int groan_gpu_push_frames(groan_gpu_ctx *ctx, const float *xyz, const float *box, size_t n_frames) {
    if (!ctx || !xyz) return GROAN_EINVAL;
    int rc = begin_batch(ctx, n_frames, box, true);
    if (rc) return rc;
    ctx->attached = false;
    float *dst = ctx->d_slot[ctx->slot];
    ctx->cur_xyz = dst;
    rc = h2d_on_copy_stream(ctx, dst, xyz, n_frames * ctx->n_atoms * 3 * sizeof(float));
    if (!rc) rc = end_batch(ctx);
    if (rc) ctx->have_frames = false;  // a failed upload leaves no current batch rather than a half-switched one
    return rc;
}

int groan_gpu_push_frames_quantized(groan_gpu_ctx *ctx, const void *q, int elem_bytes, const int32_t *origin, float precision,
                                    const float *box, size_t n_frames) {
    if (!ctx || !q || (elem_bytes != 2 && elem_bytes != 4) || !(precision > 0.0f)) return GROAN_EINVAL;
    int rc = begin_batch(ctx, n_frames, box, true);
    if (rc) return rc;
    ctx->attached = false;
    const int slot = ctx->slot;
    float *dst = ctx->d_slot[slot];
    ctx->cur_xyz = dst;
    const size_t cap = ctx->max_frames * ctx->n_atoms * 3 * sizeof(int32_t);
    if (!ctx->d_quant[0]) {
        for (int s = 0; s < 2; s++) {
            CK(cudaMalloc(&ctx->d_quant[s], cap));
            CK(cudaMalloc(&ctx->d_origin[s], ctx->max_frames * 3 * sizeof(int32_t)));
        }
        ctx->quant_bytes = cap;
    }
    const size_t count = n_frames * ctx->n_atoms * 3;
    rc = h2d_on_copy_stream(ctx, ctx->d_quant[slot], q, count * (size_t)elem_bytes);
    if (rc) return rc;
    if (origin) CK(cudaMemcpyAsync(ctx->d_origin[slot], origin, n_frames * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy));
    // xdrfile.c:844: inv_precision = 1.0 / *precision (double quotient stored in a float); :915 coordinate = int * inv_precision
    const float inv = (float)(1.0 / (double)precision);
    const unsigned nb = (unsigned)std::min<size_t>((ctx->n_atoms * 3 + kThreads * 4 - 1) / (kThreads * 4), (size_t)kMaxBlocksPerFrame * 4);
    dim3 grid(std::max(1u, nb), (unsigned)n_frames);
    // on the COPY stream: ordered behind its own upload, overlapping the kernels still running on the previous batch
    if (elem_bytes == 2)
        k_dequantize<int16_t><<<grid, kThreads, 0, ctx->copy>>>((const int16_t *)ctx->d_quant[slot], origin ? ctx->d_origin[slot] : nullptr,
                                                                 inv, dst, ctx->n_atoms);
    else
        k_dequantize<int32_t><<<grid, kThreads, 0, ctx->copy>>>((const int32_t *)ctx->d_quant[slot], origin ? ctx->d_origin[slot] : nullptr,
                                                                 inv, dst, ctx->n_atoms);
    LAUNCHED();
    return end_batch(ctx);
}

int groan_gpu_attach_frames(groan_gpu_ctx *ctx, float *d_xyz, const float *box, size_t n_frames) {
    if (!ctx || !d_xyz || classify(d_xyz) != PK_DEVICE) return GROAN_EINVAL;
    int rc = begin_batch(ctx, n_frames, box, false);
    if (rc) return rc;
    ctx->attached = true;
    ctx->cur_xyz = d_xyz;
    return end_batch(ctx);
}

int groan_gpu_set_valid(groan_gpu_ctx *ctx, const uint8_t *valid) {
    if (!ctx) return GROAN_EINVAL;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    ctx->valid_cache.clear();
    ctx->has_valid = false;
    if (!valid) return GROAN_OK;
    if (!ctx->d_valid) CK(cudaMalloc(&ctx->d_valid, ctx->max_frames * ctx->n_atoms));
    // host or device source; ordered on the compute stream like the ops that will ask about it
    CK(cudaMemcpyAsync(ctx->d_valid, valid, ctx->n_frames * ctx->n_atoms, cudaMemcpyDefault, ctx->compute));
    if (classify(valid) == PK_PAGEABLE) CK(cudaStreamSynchronize(ctx->compute));
    ctx->has_valid = true;
    return GROAN_OK;
}

int groan_gpu_get_frames(groan_gpu_ctx *ctx, float *xyz_out) {
    if (!ctx || !xyz_out) return GROAN_EINVAL;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    return deliver(ctx, xyz_out, ctx->cur_xyz, ctx->n_frames * ctx->n_atoms * 3 * sizeof(float));
}

int groan_gpu_get_frames_quantized(groan_gpu_ctx *ctx, int32_t *q_out, float precision) {
    if (!ctx || !q_out || !(precision > 0.0f)) return GROAN_EINVAL;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    const size_t count = ctx->n_frames * ctx->n_atoms * 3;
    int32_t *d_q = q_out;
    if (classify(q_out) != PK_DEVICE) {
        int rc = ensure_tmp(ctx, count * sizeof(int32_t));
        if (rc) return rc;
        d_q = (int32_t *)ctx->d_tmp;
    }
    const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((count + kThreads * 4 - 1) / (kThreads * 4), (size_t)kSMs * 16));
    k_quantize<<<nb, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, precision, d_q, count);
    LAUNCHED();
    return deliver(ctx, q_out, d_q, count * sizeof(int32_t));
}

// ---- centres --------------------------------------------------------------------------------------
int groan_gpu_estimate_center(groan_gpu_ctx *ctx, int gid, int weighted, float *out) {
    const Group *g = nullptr;
    int rc = check_center_args(ctx, gid, weighted != 0, &g, true);
    if (rc) return rc;
    if (!out) return GROAN_EINVAL;
    float *d_out = target_of<float>(out, ctx->d_cen);
    rc = run_trig(ctx, *g, weighted != 0, d_out, nullptr, 1);
    if (rc) return rc;
    return deliver(ctx, out, d_out, ctx->n_frames * 3 * sizeof(float));
}

int groan_gpu_get_center(groan_gpu_ctx *ctx, int gid, int weighted, float *out) {
    const Group *g = nullptr;
    int rc = check_center_args(ctx, gid, weighted != 0, &g, true);
    if (rc) return rc;
    if (!out) return GROAN_EINVAL;
    float *d_out = target_of<float>(out, ctx->d_cen);
    rc = run_get_center(ctx, *g, weighted != 0, d_out);
    if (rc) return rc;
    return deliver(ctx, out, d_out, ctx->n_frames * 3 * sizeof(float));
}

int groan_gpu_get_center_naive(groan_gpu_ctx *ctx, int gid, float *out) {
    if (!ctx || !out) return GROAN_EINVAL;
    const Group *g = get_group(ctx, gid);
    if (!g) return GROAN_ENOGROUP;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    if (g->n == 0) return GROAN_EEMPTY;
    int rc = check_positions(ctx, *g);
    if (rc) return rc;
    float *d_out = target_of<float>(out, ctx->d_cen);
    const int nb = blocks_per_frame(g->n, ctx->n_frames);
    dim3 grid(nb, (unsigned)ctx->n_frames);
    k_naive<<<grid, kThreads, 0, ctx->compute>>>(frames_of(ctx), view_of(*g), ctx->d_partials, ctx->d_tickets, d_out);
    LAUNCHED();
    return deliver(ctx, out, d_out, ctx->n_frames * 3 * sizeof(float));
}

// ---- distances ------------------------------------------------------------------------------------
int groan_gpu_group_distance(groan_gpu_ctx *ctx, int g1, int g2, int dim, float *out) {
    if (!ctx || !out || dim < 0 || dim > 7) return GROAN_EINVAL;
    const Group *a = nullptr, *b = nullptr;
    int rc = check_center_args(ctx, g1, false, &a);  // group_get_center(group1)? then group2 (analysis.rs:353-354)
    if (rc) return rc;
    rc = check_center_args(ctx, g2, false, &b);
    if (rc) return rc;
    rc = run_get_center(ctx, *a, false, ctx->d_cen);
    if (rc) return rc;
    rc = run_get_center(ctx, *b, false, ctx->d_cen2);
    if (rc) return rc;
    float *d_out = target_of<float>(out, ctx->d_res);
    const unsigned F = (unsigned)ctx->n_frames;
    k_center_distance<<<(F + 127) / 128, 128, 0, ctx->compute>>>(ctx->d_cen, ctx->d_cen2, frames_of(ctx), dim, (int)F, d_out);
    LAUNCHED();
    return deliver(ctx, out, d_out, F * sizeof(float));
}

// ---- wrap / translate -----------------------------------------------------------------------------
int groan_gpu_wrap(groan_gpu_ctx *ctx, int gid, int8_t *shifts) {
    const Group *g = nullptr;
    bool tric = false;
    int rc = check_wrap_args(ctx, gid, &g, &tric);
    if (rc) return rc;
    if (g->n == 0) return GROAN_OK;
    return run_wrap(ctx, *g, false, nullptr, shifts, tric);
}

int groan_gpu_translate(groan_gpu_ctx *ctx, int gid, const float t[3], int8_t *shifts) {
    if (!t) return GROAN_EINVAL;
    const Group *g = nullptr;
    bool tric = false;
    int rc = check_wrap_args(ctx, gid, &g, &tric);
    if (rc) return rc;
    if (g->n == 0) return GROAN_OK;
    return run_wrap(ctx, *g, true, t, shifts, tric);
}

// ---- whole groups / molecules, centering ---------------------------------------
int groan_gpu_make_group_whole(groan_gpu_ctx *ctx, int gid) {
    // group_estimate_center(group)? comes first (modifying.rs:439), with all of its checks
    const Group *g = nullptr;
    int rc = check_center_args(ctx, gid, false, &g);
    if (rc) return rc;
    rc = run_trig(ctx, *g, false, ctx->d_c0, nullptr);
    if (rc) return rc;
    size_t nb = (g->n + kThreads - 1) / kThreads;
    nb = std::max<size_t>(1, std::min<size_t>(nb, (size_t)kMaxBlocksPerFrame * 4));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    k_make_whole<<<grid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->d_box[ctx->slot], ctx->n_atoms, view_of(*g), ctx->d_c0);
    LAUNCHED();
    return GROAN_OK;
}

int groan_gpu_set_molecules(groan_gpu_ctx *ctx, const uint32_t *mol_ref) {
    if (!ctx || !mol_ref) return GROAN_EINVAL;
    for (size_t i = 0; i < ctx->n_atoms; i++) {
        const uint32_t r = mol_ref[i];
        if (r == kNoMolecule) continue;
        // the reference atom of a molecule is its lowest index (modifying.rs:258-283) and refers to itself
        if (r > i || mol_ref[r] != r) return GROAN_EINVAL;
    }
    CK(cudaStreamSynchronize(ctx->compute));
    if (!ctx->d_mol_ref) CK(cudaMalloc(&ctx->d_mol_ref, ctx->n_atoms * sizeof(uint32_t)));
    ctx->mol_ref.assign(mol_ref, mol_ref + ctx->n_atoms);
    ctx->valid_cache.clear();
    CK(cudaMemcpy(ctx->d_mol_ref, mol_ref, ctx->n_atoms * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return GROAN_OK;
}

int groan_gpu_make_molecules_whole(groan_gpu_ctx *ctx) {
    if (!ctx) return GROAN_EINVAL;
    if (!ctx->d_mol_ref) return GROAN_EINVAL;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    int rc = check_box(ctx, false, nullptr);  // simbox_check first (modifying.rs:350)
    if (rc) return rc;
    {  // first atom of a polyatomic molecule without a position (modifying.rs:362-380)
        bool found;
        size_t f = 0, i = 0;
        rc = first_invalid(ctx, nullptr, &found, &f, &i);
        if (rc) return rc;
        if (found) {
            ctx->err_a = f;
            ctx->err_b = i;
            return GROAN_ENOPOS;
        }
    }
    size_t nb = (ctx->n_atoms + kThreads - 1) / kThreads;
    nb = std::max<size_t>(1, std::min<size_t>(nb, (size_t)kMaxBlocksPerFrame * 4));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    k_mol_whole<<<grid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->d_box[ctx->slot], ctx->n_atoms, ctx->d_mol_ref);
    LAUNCHED();
    return GROAN_OK;
}

int groan_gpu_atoms_center(groan_gpu_ctx *ctx, int gid, int weighted, int dim) {
    if (dim < 0 || dim > 7) return GROAN_EINVAL;
    // group_estimate_center / group_estimate_com of the reference group first (utility.rs:114,173) ...
    const Group *g = nullptr;
    int rc = check_center_args(ctx, gid, weighted != 0, &g);
    if (rc) return rc;
    // ... then atoms_translate: every atom of the system needs a position (utility.rs:120-125)
    rc = check_positions(ctx, ctx->all);
    if (rc) return rc;
    rc = run_trig(ctx, *g, weighted != 0, ctx->d_c0, nullptr);
    if (rc) return rc;
    size_t nb = (ctx->n_atoms + kThreads - 1) / kThreads;
    nb = std::max<size_t>(1, std::min<size_t>(nb, (size_t)kMaxBlocksPerFrame * 4));
    dim3 grid((unsigned)nb, (unsigned)ctx->n_frames);
    k_center_atoms<<<grid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->d_box[ctx->slot], ctx->n_atoms, ctx->d_c0, dim);
    LAUNCHED();
    return GROAN_OK;
}

// ---- RMSD -----------------------------------------------------------------------------------------
int groan_gpu_rmsd_set_reference(groan_gpu_ctx *ctx, int gid, const float *ref_xyz, size_t n_ref_atoms, const uint32_t *ref_idx,
                                 size_t n_ref, const float ref_box[9], const float *ref_mass) {
    if (!ctx || !ref_xyz || gid < GROAN_GROUP_ALL || gid >= GROAN_MAX_GROUPS) return GROAN_EINVAL;
    const Group *g = get_group(ctx, gid);
    if (!g) return GROAN_ENOGROUP;
    if (!ref_box) return GROAN_ENOBOX;  // get_box_center (mod.rs:298-308) comes first in extract_data_from_system
    if (ref_box[1] != 0.0f || ref_box[2] != 0.0f || ref_box[5] != 0.0f) return GROAN_EINVAL;
    const bool ref_tric = ref_box[3] != 0.0f || ref_box[6] != 0.0f || ref_box[7] != 0.0f;
    if (ref_tric && !(ctx->flags & GROAN_FLAG_TRICLINIC)) return GROAN_ENOTORTHO;
    if (ref_box[0] == 0.0f || ref_box[4] == 0.0f || ref_box[8] == 0.0f) return GROAN_EZEROBOX;
    if (n_ref == 0) return GROAN_EEMPTY;
    if (!ref_idx && n_ref > n_ref_atoms) return GROAN_EINVAL;  // ref_idx == NULL: the first n_ref atoms of the reference
    for (size_t i = 0; ref_idx && i < n_ref; i++) {
        if (ref_idx[i] >= n_ref_atoms) return GROAN_EINVAL;
        if (i && ref_idx[i] <= ref_idx[i - 1]) return GROAN_EINVAL;
    }
    int rc = GROAN_OK;
    if (ref_mass) {
        // masses of the REFERENCE system's group: reference.group_get_com and the Kabsch weights (rmsd.rs:154,192)
        for (size_t i = 0; i < n_ref; i++)
            if (ref_mass[i] < 0.0f) {
                ctx->err_a = 0;
                ctx->err_b = ref_idx ? ref_idx[i] : i;
                return GROAN_ENOMASS;
            }
    } else {
        rc = check_masses(ctx, *g);
        if (rc) return rc;
        if (n_ref != g->n) {  // borrowing the target group's masses needs equal sizes
            ctx->err_a = n_ref;
            ctx->err_b = g->n;
            return GROAN_EGROUPSIZE;
        }
    }
    groan_gpu_ctx::Ref &R = ctx->refs[ref_slot(gid)];
    CK(cudaStreamSynchronize(ctx->compute));
    if (R.d_pc) { cudaFree(R.d_pc); R.d_pc = nullptr; }
    for (float *&q : R.d_pq)
        if (q) { cudaFree(q); q = nullptr; }
    R.set = false;
    float *d_ref = nullptr, *d_refbox = nullptr, *d_small = nullptr, *d_rmass = nullptr;
    uint32_t *d_ridx = nullptr;
    double *d_sums = nullptr;
    rc = [&]() -> int {
        CK(cudaMalloc(&d_ref, n_ref_atoms * 3 * sizeof(float)));
        CK(cudaMalloc(&d_refbox, 9 * sizeof(float)));
        CK(cudaMalloc(&d_small, 6 * sizeof(float)));
        if (ref_idx) CK(cudaMalloc(&d_ridx, n_ref * sizeof(uint32_t)));
        CK(cudaMalloc(&d_sums, kRefSums * sizeof(double)));
        CK(cudaMalloc(&R.d_pc, ref_floats(n_ref) * sizeof(float)));
        CK(cudaMemsetAsync(R.d_pc, 0, ref_floats(n_ref) * sizeof(float), ctx->compute));
        CK(cudaMemcpyAsync(d_ref, ref_xyz, n_ref_atoms * 3 * sizeof(float), cudaMemcpyDefault, ctx->compute));
        CK(cudaMemcpyAsync(d_refbox, ref_box, 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
        if (ref_idx) CK(cudaMemcpyAsync(d_ridx, ref_idx, n_ref * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->compute));
        if (ref_mass) {
            CK(cudaMalloc(&d_rmass, n_ref * sizeof(float)));
            CK(cudaMemcpyAsync(d_rmass, ref_mass, n_ref * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
        }
        FrameView fv;
        fv.xyz = d_ref;
        fv.box = d_refbox;
        fv.n_atoms = n_ref_atoms;
        fv.tric = ref_tric ? 1 : 0;
        GroupView gv;
        gv.idx = d_ridx;
        gv.first = 0;
        gv.n = (uint32_t)n_ref;
        gv.mass = ref_mass ? d_rmass : g->d_mass;
        const int nb = blocks_per_frame(n_ref, 1);
        // reference.group_get_com(group): geometric estimate, then mass-weighted unwrap (iterators.rs:1404-1438)
        k_trig<false><<<dim3(nb, 1), kThreads, 0, ctx->compute>>>(fv, gv, ctx->d_partials, ctx->d_tickets + ctx->max_frames, d_small,
                                                                   nullptr);
        LAUNCHED();
        k_unwrap<true><<<dim3(nb, 1), kThreads, 0, ctx->compute>>>(fv, gv, d_small, ctx->d_partials, ctx->d_tickets + ctx->max_frames,
                                                                    d_small + 3, nullptr, 0);
        LAUNCHED();
        k_ref_prepare<<<dim3(nb, 1), kThreads, 0, ctx->compute>>>(fv, gv, d_small + 3, R.d_pc, ctx->d_partials,
                                                                   ctx->d_tickets + ctx->max_frames, d_sums);
        LAUNCHED();
        CK(cudaMemcpyAsync(R.sums, d_sums, sizeof(R.sums), cudaMemcpyDeviceToHost, ctx->compute));
        CK(cudaMemcpyAsync(R.com, d_small + 3, 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->compute));
        CK(cudaStreamSynchronize(ctx->compute));
        R.same_mass = !ref_mass || (g->mass.size() == n_ref && std::memcmp(g->mass.data(), ref_mass, n_ref * sizeof(float)) == 0);
        return GROAN_OK;
    }();
    cudaFree(d_ref); cudaFree(d_refbox); cudaFree(d_small); cudaFree(d_ridx); cudaFree(d_sums); cudaFree(d_rmass);
    if (rc) return rc;
    R.n = n_ref;
    R.set = true;
    return GROAN_OK;
}

int groan_gpu_rmsd(groan_gpu_ctx *ctx, int gid, float *rmsd, float *rot) {
    if (!rmsd) return GROAN_EINVAL;
    return rmsd_common(ctx, gid, rmsd, rot, false);
}

int groan_gpu_center_rmsd(groan_gpu_ctx *ctx, int gid, int weighted, float *center, float *rmsd, float *rot) {
    if (!rmsd || !center) return GROAN_EINVAL;
    return rmsd_common(ctx, gid, rmsd, rot, false, center, weighted);
}

int groan_gpu_rmsd_fit(groan_gpu_ctx *ctx, int gid, float *rmsd) {
    if (!rmsd) return GROAN_EINVAL;
    return rmsd_common(ctx, gid, rmsd, nullptr, true);
}

// ---- synthetic workloads ----------------------------------------------------------------------------
int groan_gpu_synth_uniform(groan_gpu_ctx *ctx, uint64_t seed, uint64_t frame0, size_t n_frames, const float lo[3],
                            const float span[3], const float *box) {
    if (!ctx || !lo || !span) return GROAN_EINVAL;
    int rc = begin_batch(ctx, n_frames, box, true);
    if (rc) return rc;
    ctx->attached = false;
    ctx->cur_xyz = ctx->d_slot[ctx->slot];
    rc = end_batch(ctx);
    if (rc) return rc;
    dim3 grid(kMaxBlocksPerFrame, (unsigned)n_frames);
    k_synth_uniform<<<grid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->n_atoms, seed, frame0, lo[0], lo[1], lo[2], span[0],
                                                          span[1], span[2]);
    LAUNCHED();
    return GROAN_OK;
}

int groan_gpu_synth_blob(groan_gpu_ctx *ctx, uint64_t seed, uint64_t frame0, size_t n_frames, float scale, float nscale,
                         const float *rot, const float *centre, const float *box, int wrap) {
    if (!ctx || !rot || !centre || !box) return GROAN_EINVAL;
    int rc = begin_batch(ctx, n_frames, box, true);
    if (rc) return rc;
    ctx->attached = false;
    ctx->cur_xyz = ctx->d_slot[ctx->slot];
    rc = end_batch(ctx);
    if (rc) return rc;
    rc = ensure_tmp(ctx, n_frames * 12 * sizeof(float));
    if (rc) return rc;
    float *d_rot = (float *)ctx->d_tmp, *d_cen = d_rot + n_frames * 9;
    CK(cudaMemcpyAsync(d_rot, rot, n_frames * 9 * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
    CK(cudaMemcpyAsync(d_cen, centre, n_frames * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->compute));
    dim3 grid(kMaxBlocksPerFrame, (unsigned)n_frames);
    k_synth_blob<<<grid, kThreads, 0, ctx->compute>>>(ctx->cur_xyz, ctx->n_atoms, seed, frame0, scale, nscale, d_rot, d_cen,
                                                       ctx->d_box[ctx->slot], wrap);
    LAUNCHED();
    return GROAN_OK;
}

// reference structure of the blob workload (n_atoms x 3), written to a caller buffer (host or device)
int groan_gpu_synth_blob_ref(groan_gpu_ctx *ctx, uint64_t seed, float scale, const float centre[3], float *xyz_out) {
    if (!ctx || !centre || !xyz_out) return GROAN_EINVAL;
    const size_t bytes = ctx->n_atoms * 3 * sizeof(float);
    float *d_out = xyz_out;
    if (classify(xyz_out) != PK_DEVICE) {
        int rc = ensure_tmp(ctx, bytes);
        if (rc) return rc;
        d_out = (float *)ctx->d_tmp;
    }
    k_synth_blob_ref<<<kMaxBlocksPerFrame, kThreads, 0, ctx->compute>>>(d_out, ctx->n_atoms, seed, scale, centre[0], centre[1],
                                                                         centre[2]);
    LAUNCHED();
    return deliver(ctx, xyz_out, d_out, bytes);
}

}  // extern "C"

// ctx.cuh -- the context behind the C ABI (include/groan_gpu.h) and the host-side helpers every translation unit of
// libgroan_gpu.so shares: argument checks in the reference's order, result delivery, batch bookkeeping.
#pragma once
#include "../../include/groan_gpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"

namespace groan_host {
using namespace groan;

constexpr size_t kPairPartialBytes = 32;  // == sizeof(groan::PairPartial) (kernels_pairs.cuh; static_assert in groan_pairs.cu)
constexpr int kRefSumsHost = 8;  // == groan::kRefSums (kernels_rmsd.cuh; static_assert in groan_gpu.cu)


constexpr size_t kStageBytes = 32u << 20;  // pinned staging chunk for pageable sources
constexpr size_t kPartialSlots = 8192;     // (blocks per frame) x (frames) upper bound for reductions
constexpr int kMaxSums = 48;               // widest per-CTA partial record, in doubles
constexpr size_t kMaxFramesPerBatch = 32768;  // gridDim.y / the 16-bit frame counter of maybe_launch_fallback

struct Group {
    bool set = false;
    std::vector<uint32_t> idx;  // host copy (validity / error reporting)
    uint32_t *d_idx = nullptr;
    bool contiguous = false;
    uint32_t first = 0;
    size_t n = 0;
    bool has_mass = false;
    long no_mass_at = -1;  // position in the group of the first atom without mass
    float *d_mass = nullptr;
    std::vector<float> mass;  // host copy (compared with the RMSD reference's masses)
    double mass_sum = 0.0;    // sum of `mass` in f64, ascending order (set once in groan_gpu_set_group, never per call)
};

enum PtrKind { PK_DEVICE, PK_PINNED, PK_PAGEABLE };

inline PtrKind classify(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return PK_PAGEABLE;
    }
    switch (a.type) {
    case cudaMemoryTypeDevice:
    case cudaMemoryTypeManaged: return PK_DEVICE;
    case cudaMemoryTypeHost: return PK_PINNED;
    default: return PK_PAGEABLE;
    }
}


}  // namespace groan_host

struct groan_gpu_ctx {
    int device = 0;
    size_t n_atoms = 0, max_frames = 0;
    unsigned flags = 0;
    cudaStream_t own_compute = nullptr, compute = nullptr, copy = nullptr;

    // frames
    float *d_slot[2] = {nullptr, nullptr};
    float *d_box[2] = {nullptr, nullptr};
    int slot = 1;               // slot of the current batch (first push goes to 0)
    float *cur_xyz = nullptr;   // slot buffer or attached caller buffer
    bool attached = false;
    bool have_frames = false, have_box = false;
    bool batch_tric = false;    // some frame of the batch has a triclinic box (set by check_box)
    size_t n_frames = 0;
    std::vector<float> h_box;   // F x 9 of the current batch
    // Option<Vector3D> positions (groan_gpu_set_valid): the bitmap lives on the device; an op asks "first atom of this group
    // without a position" once per (bitmap, group) -- the answer is cached, so the per-op cost is a table lookup
    uint8_t *d_valid = nullptr;  // F x N, allocated on first use
    bool has_valid = false;
    struct ValidAnswer {
        const void *group;  // Group the question was asked for (nullptr: atoms of polyatomic molecules)
        bool found;
        size_t frame, atom;
    };
    std::vector<ValidAnswer> valid_cache;
    unsigned long long *d_valid_first = nullptr;
    cudaEvent_t ev_h2d = nullptr, ev_done[2] = {nullptr, nullptr};
    bool done_recorded[2] = {false, false};
    void *d_quant[2] = {nullptr, nullptr};       // quantised frames as uploaded (groan_gpu_push_frames_quantized), one per slot
    size_t quant_bytes = 0;
    int32_t *d_origin[2] = {nullptr, nullptr};   // their per-frame integer origins (F x 3)
    // xtc streams decoded on the device (groan_xtc.cu): the file's bytes as uploaded, per-frame parameters, damage flags
    void *d_xtc[2] = {nullptr, nullptr};
    size_t xtc_cap = 0;
    void *d_xtc_params[2] = {nullptr, nullptr};
    int *d_xtc_status = nullptr;
    size_t xtc_frames = 0;
    // partial frames (groan_gpu_push_group_frames): compact upload + the atom list
    void *d_sel[2] = {nullptr, nullptr};
    size_t sel_cap = 0;
    uint32_t *d_sel_atoms = nullptr;
    size_t sel_atoms_cap = 0;
    float *h_stage[2] = {nullptr, nullptr};
    cudaEvent_t ev_stage[2] = {nullptr, nullptr};

    groan_host::Group groups[GROAN_MAX_GROUPS];
    groan_host::Group all;  // GROAN_GROUP_ALL

    // scratch
    double *d_partials = nullptr;
    void *d_pair_partials = nullptr;
    unsigned int *d_tickets = nullptr;
    float *d_c0 = nullptr, *d_cen = nullptr, *d_cen2 = nullptr, *d_res = nullptr, *d_rot = nullptr;
    int *d_flags = nullptr;  // per frame: 1 = the single-pass kernel could not certify its result, redo exactly
    unsigned int *d_frames_done = nullptr;  // device-side fallback launch: frames finished by the running single-pass kernel
    int *d_flags2 = nullptr;                // per frame: 1 = the fused centre + RMSD kernel sent the frame through the sine-sum centre pass
    unsigned int *d_second_any = nullptr;   // device-side launch of that pass: how many frames of the running launch want it
    int *d_second_list = nullptr;           // ... and which
    unsigned int *d_slow_count = nullptr;   // the same for the frames a single pass of a contiguous group flags (exact quad passes)
    int *d_slow_list = nullptr;
    // feedback of the device-side fallback protocol, one word per group slot (pinned, mapped): see FallbackPlan::feedback
    unsigned long long *h_feedback = nullptr, *d_feedback = nullptr;
    unsigned fast_skips[GROAN_MAX_GROUPS + 1] = {};  // consecutive calls that went straight to the exact passes
    int *d_head_list = nullptr;             // frames of the batch sorted by the 16-byte phase of a group (launch_rmsd_quad), cached per key
    size_t head_list_key_frames = 0;
    uint32_t head_list_key_first = 0xffffffffu;
    bool second_valid = false;              // d_flags2 belongs to the last call
    int occ_center = 4, occ_rmsd = 2;  // resident CTAs per SM of the single-pass kernels
    int occ_center_quad = 0;                   // quad kernels (kernels_quad.cuh)
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    uint32_t *d_mol_ref = nullptr;      // make_molecules_whole: reference atom of every atom's molecule (groan_gpu_set_molecules)
    std::vector<uint32_t> mol_ref;      // host copy (position checks)

    // RMSD reference (per group id)
    struct Ref {
        bool set = false;
        size_t n = 0;
        float *d_pc = nullptr;  // block-SoA prepared reference (kernels_rmsd.cuh)
        float *d_pq[4] = {nullptr, nullptr, nullptr, nullptr};  // quad-permuted copies of the aligned body (kernels_quad.cuh), one per
                                                                 // head (atoms before the first 16-byte boundary), built on first use
        double sums[groan_host::kRefSumsHost] = {0, 0, 0, 0, 0, 0, 0, 0};
        bool same_mass = true;  // reference masses == the target group's masses
        float com[3] = {0, 0, 0};
    } refs[GROAN_MAX_GROUPS + 1];  // last slot: GROAN_GROUP_ALL

    uint64_t launches = 0;
    std::string cuda_err;
    size_t err_a = 0, err_b = 0;
};

namespace groan_host {


inline int cuda_fail(groan_gpu_ctx *c, cudaError_t e, const char *what) {
    if (c) c->cuda_err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return GROAN_ECUDA;
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);     \
    } while (0)
#define LAUNCHED()                                                   \
    do {                                                             \
        ctx->launches++;                                             \
        cudaError_t e_ = cudaGetLastError();                         \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, "kernel launch"); \
    } while (0)

inline const Group *get_group(groan_gpu_ctx *ctx, int gid) {
    if (gid == GROAN_GROUP_ALL) return &ctx->all;
    if (gid < 0 || gid >= GROAN_MAX_GROUPS || !ctx->groups[gid].set) return nullptr;
    return &ctx->groups[gid];
}

// slot of a group id in ctx->refs: the built-in "all" group owns the last one
inline int ref_slot(int gid) { return gid == GROAN_GROUP_ALL ? GROAN_MAX_GROUPS : gid; }

inline GroupView view_of(const Group &g) {
    GroupView v;
    v.idx = g.contiguous ? nullptr : g.d_idx;
    v.first = g.first;
    v.n = (uint32_t)g.n;
    v.mass = g.d_mass;
    return v;
}

inline FrameView frames_of(groan_gpu_ctx *ctx) {
    FrameView fv;
    fv.xyz = ctx->cur_xyz;
    fv.box = ctx->d_box[ctx->slot];
    fv.n_atoms = ctx->n_atoms;
    fv.tric = 0;
    return fv;
}

// the same, for the centre / RMSD kernels of the triclinic extension: they see sheared coordinates when the batch has a
// triclinic frame (the shear of an orthogonal frame is the identity bit for bit)
inline FrameView frames_of_geom(groan_gpu_ctx *ctx) {
    FrameView fv = frames_of(ctx);
    fv.tric = ctx->batch_tric ? 1 : 0;
    return fv;
}

// simbox_check (simbox.rs:230-236) over every frame of the batch; zero box = the reference's panic
inline int check_box(groan_gpu_ctx *ctx, bool allow_triclinic, bool *any_triclinic) {
    if (any_triclinic) *any_triclinic = false;
    ctx->batch_tric = false;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    if (!ctx->have_box) return GROAN_ENOBOX;
    for (size_t f = 0; f < ctx->n_frames; f++) {
        const float *b = &ctx->h_box[f * 9];
        if (b[1] != 0.0f || b[2] != 0.0f || b[5] != 0.0f) return GROAN_EINVAL;  // matrix2simbox rejects (xdrfile.rs:171)
        const bool tric = (b[3] != 0.0f || b[6] != 0.0f || b[7] != 0.0f);
        if (tric) {
            if (!allow_triclinic || !(ctx->flags & GROAN_FLAG_TRICLINIC)) return GROAN_ENOTORTHO;
            if (any_triclinic) *any_triclinic = true;
            ctx->batch_tric = true;
        }
        if (b[0] == 0.0f || b[4] == 0.0f || b[8] == 0.0f) return GROAN_EZEROBOX;
    }
    return GROAN_OK;
}

// first atom of the group (group order) without a position, in the first frame that has one (groan_gpu.cu; device scan,
// cached per bitmap and group)
int check_positions(groan_gpu_ctx *ctx, const Group &g);
// the same question for an arbitrary group or, with g == nullptr, for the atoms of polyatomic molecules (ctx->d_mol_ref)
int first_invalid(groan_gpu_ctx *ctx, const Group *g, bool *found, size_t *frame, size_t *pos);

inline int check_masses(groan_gpu_ctx *ctx, const Group &g) {
    if (!g.has_mass) {
        ctx->err_a = 0;
        ctx->err_b = g.n ? (g.contiguous ? g.first : g.idx[0]) : 0;
        return GROAN_ENOMASS;
    }
    if (g.no_mass_at >= 0) {
        ctx->err_a = 0;
        ctx->err_b = g.contiguous ? g.first + (size_t)g.no_mass_at : g.idx[(size_t)g.no_mass_at];
        return GROAN_ENOMASS;
    }
    return GROAN_OK;
}


inline int ensure_tmp(groan_gpu_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->tmp_bytes) return GROAN_OK;
    if (ctx->d_tmp) {
        CK(cudaStreamSynchronize(ctx->compute));
        CK(cudaFree(ctx->d_tmp));
        ctx->d_tmp = nullptr;
        ctx->tmp_bytes = 0;
    }
    CK(cudaMalloc(&ctx->d_tmp, bytes));
    ctx->tmp_bytes = bytes;
    return GROAN_OK;
}

// copy a device result to wherever the caller wants it (device / pinned: async; pageable: blocking)
inline int deliver(groan_gpu_ctx *ctx, void *out, const void *d_src, size_t bytes) {
    if (!out || out == d_src || bytes == 0) return GROAN_OK;
    const PtrKind k = classify(out);
    CK(cudaMemcpyAsync(out, d_src, bytes, cudaMemcpyDefault, ctx->compute));
    if (k == PK_PAGEABLE) CK(cudaStreamSynchronize(ctx->compute));
    return GROAN_OK;
}

template <typename T>
T *target_of(void *out, T *scratch) {
    return (out && classify(out) == PK_DEVICE) ? reinterpret_cast<T *>(out) : scratch;
}

// smallest float t with sqrtf(t) >= c: (d2 < t) <=> (sqrtf(d2) < c) for every float d2 >= 0
inline float cutoff_squared_threshold(float c) {
    if (!(c > 0.0f)) return 0.0f;
    float t = c * c;
    if (std::isinf(t)) return t;
    while (std::sqrt(t) >= c && t > 0.0f) t = std::nextafter(t, 0.0f);
    while (std::sqrt(t) < c) t = std::nextafter(t, INFINITY);
    return t;
}

// Atom::distance checks self first, then the other atom (atom.rs:780-790); scan order is row-major
int check_pair_positions(groan_gpu_ctx *ctx, const Group &a, const Group &b);

// batch bookkeeping (groan_gpu.cu): switch to the other device slot / make the uploaded batch visible to the compute stream
int begin_batch(groan_gpu_ctx *ctx, size_t F, const float *box, bool use_slot);
int end_batch(groan_gpu_ctx *ctx);
int h2d_on_copy_stream(groan_gpu_ctx *ctx, void *dst, const void *src, size_t bytes);

}  // namespace groan_host

// kernels_tma.cuh -- what the ring-fed kernels of kernels_quad.cuh share: the mbarrier / bulk-copy primitives, the geometry of a
// contiguous group inside a frame, and the device-side launch of the passes that re-do frames a single pass flags.
//
// A contiguous group of a frame is a contiguous byte range of the AoS coordinate buffer, so it can be moved global ->
// shared memory by 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier: no registers and no
// issue slots are spent on addresses or on keeping loads in flight, and the number of bytes in flight is set by the ring
// depth, not by occupancy.  (The second-generation kernels that lived here -- pairs of atoms, stride-3 LDS.32, sin + cos
// sums, optionally four frames per CTA -- were measured slower than the quad kernels in every configuration and are gone;
// profiles/r1_v2_to_v6_single_pass.md keeps their numbers.)
#pragma once
#include "common.cuh"
#include "kernels_center.cuh"
#include "kernels_rmsd.cuh"

namespace groan {


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the warp until the phase completes instead of spinning on issue slots
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
                 : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

struct BodyGeom {
    uint32_t head, body, tail;
};
__device__ __forceinline__ BodyGeom body_geom(const FrameView &fv, const GroupView &g, int f) {
    BodyGeom b;
    const size_t a0 = (size_t)f * fv.n_atoms + g.first;
    b.head = (uint32_t)((4 - (a0 & 3)) & 3);
    if (b.head > g.n) b.head = g.n;
    b.body = (g.n - b.head) & ~3u;
    b.tail = g.n - b.head - b.body;
    return b;
}

// packed helpers (sm_100 f32x2 pipe)
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

// ---------------------------------------------------------------- device-side launch of the fallback passes
// The reference-order passes are needed only for frames the single pass could not certify -- usually none.  Launching
// them from the host every time costs four grids of early-exiting CTAs per call (~17 us of a 145 us step,
// profiles/r1_launches.csv).  Instead the thread that finishes the LAST frame of the launch looks at the flags and,
// only if one is set, tail-launches the passes from the device (CUDA dynamic parallelism, cudaStreamTailLaunch: they
// run in order after this grid, before anything the host enqueues next on the stream).
// the permuted reference of the quad kernels (kernels_quad.cuh), one copy per head (0..3 atoms before the first 16-byte
// boundary of the group in a frame)
struct QuadRef {
    const float *v[4];
};
struct FallbackPlan {
    int enabled;              // 0: the host launches the fallback passes itself
    int n_frames;             // frames of the batch
    int nb_exact, nb_cov;     // CTAs per frame of k_trig / k_unwrap and of k_cov
    unsigned int *frames_done;
    float *c0;                // Bai-Breen estimates of the flagged frames
    int want_center, center_weighted, want_rmsd;
    float *center_out;        // want_center
    float *com, *rmsd_out, *rot_out; // want_rmsd
    // second tier of the fused centre + RMSD kernels (kernels_quad.cuh, finish_center_moments): frames whose image could
    // not be certified from the moments go through the sine-sum centre pass (k_center_quad, gated by second_flags) first
    int *second_flags;          // per frame, written by the fused kernel (diagnostics, and the gate under GROAN_FLAG_HOST_FALLBACK)
    unsigned int *second_count; // number of such frames of the running fused launch (re-armed by the second-tier pass, sel_mode 3)
    int *second_list;           // ... and which (any order): the pass runs over exactly these frames
    unsigned int *second_ticket; // CTAs of the second-tier pass that are done
    int second_cap;             // frames the host-launched pass has room for; more than that: the rest is launched from the device
    int nb_second;              // CTAs per frame of that k_center_quad launch
    int second_smem;            // its dynamic shared memory
    int n_report;               // finishing threads that will call maybe_launch_fallback (n_frames, or the length of the list)
    // contiguous groups on the ring (kernels_quad.cuh): the frames a single pass flags are re-done by the exact passes
    // k_trig_quad / k_center_quad(ext_pilot) / k_cov_quad over exactly those frames
    int quad_exact;             // 1: use them (0: k_trig / k_unwrap / k_cov over the whole batch, gated by the flags)
    unsigned int *slow_count;   // number of flagged frames of the running launch (re-armed by the thread that reads it)
    int *slow_list;             // ... and which
    int nb_xc, nb_xv;           // CTAs per frame of the centre-type passes and of k_cov_quad (the same in every mode: bit-identical results)
    int cov_smem;               // dynamic shared memory of k_cov_quad
    QuadRef ref_pq;             // permuted reference (want_rmsd)
    unsigned long long *feedback; // host-mapped word: (flagged frames << 32) | frames of the launch, written by the last finisher;
                                  // the host reads it before the NEXT call on the group and skips the single pass if it was all flagged
};

// sel_mode 0: every frame of the batch; 1: frames with sel[f] != 0; 2: frames sel[0 .. gridDim.y)
template <bool WEIGHTED, bool TRIC = false>
__global__ void k_center_quad(FrameView fv, GroupView g, double *partials, unsigned int *tickets, float *out, int *flags, FallbackPlan fp,
                              const int *sel, int sel_mode, const float *ext_pilot);
__global__ void k_trig_quad(FrameView fv, GroupView g, double *partials, unsigned int *tickets, float *c0_out, const int *sel, int sel_mode);
__global__ void k_cov_quad(FrameView fv, GroupView g, RefView ref, QuadRef ref_pq, const float *com_in, double *partials,
                           unsigned int *tickets, float *rmsd_out, float *rot_out, const int *sel, int sel_mode);

// frames_done counts finished frames in its low 16 bits and flagged ones above, so that the thread that finishes the last
// frame learns from its own atomic whether any frame needs the passes: no read-back of the flags (8 dependent L2 round
// trips at the very end of the kernel, ~3 us, in the first version).  Batches are far below 65535 frames (kPartialSlots).
__device__ __forceinline__ void maybe_launch_fallback(const FallbackPlan &fp, const FrameView &fv, const GroupView &g, const RefView &ref,
                                                      double *partials, unsigned int *tickets, int *flags, int my_flag, int my_second = 0,
                                                      int my_frame = 0) {
    if (!fp.enabled) return;
    if (my_second) fp.second_list[atomicAdd(fp.second_count, 1u)] = my_frame;
    if (my_flag && fp.quad_exact) fp.slow_list[atomicAdd(fp.slow_count, 1u)] = my_frame;
    __threadfence();
    const unsigned int done = atomicAdd(fp.frames_done, my_flag ? 0x10001u : 1u);
    if ((done & 0xffffu) != (unsigned)fp.n_report - 1) return;
    *fp.frames_done = 0u; // re-arm
    const bool slow = my_flag || (done >> 16) != 0u;
    // second tier: the host launches it behind every fused launch for up to second_cap frames; only what exceeds that is
    // launched from here (a device-side launch costs ~45 us, profiles/r2_ring.md)
    const unsigned int n_all = fp.second_count != nullptr ? *reinterpret_cast<volatile unsigned int *>(fp.second_count) : 0u;
    const unsigned int n_second = n_all > (unsigned)fp.second_cap ? n_all - (unsigned)fp.second_cap : 0u;
    const unsigned int n_slow = fp.quad_exact ? atomicExch(fp.slow_count, 0u) : 0u;
    if (fp.quad_exact && fp.feedback) *fp.feedback = ((unsigned long long)n_slow << 32) | (unsigned long long)(unsigned)fp.n_report;
    if (!slow && n_second == 0u) return;
    __threadfence();
    if (slow && fp.quad_exact) {
        FallbackPlan off = fp; // the exact passes flag nothing and launch nothing
        off.enabled = 0;
        const dim3 gx(fp.nb_xc, n_slow), gv(fp.nb_xv, n_slow);
        k_trig_quad<<<gx, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.c0, fp.slow_list, 2);
        if (fp.want_rmsd) {
            k_center_quad<true><<<gx, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.com, flags, off, fp.slow_list, 2, fp.c0);
            k_cov_quad<<<gv, 256, fp.cov_smem, cudaStreamTailLaunch>>>(fv, g, ref, fp.ref_pq, fp.com, partials, tickets, fp.rmsd_out, fp.rot_out,
                                                                        fp.slow_list, 2);
        }
        if (fp.want_center) {
            if (fp.center_weighted)
                k_center_quad<true><<<gx, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, off, fp.slow_list, 2, fp.c0);
            else
                k_center_quad<false><<<gx, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, off, fp.slow_list, 2, fp.c0);
        }
    } else if (slow) {
        const dim3 ge(fp.nb_exact, fp.n_frames), gc(fp.nb_cov, fp.n_frames);
        k_trig<false><<<ge, kThreads, 0, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.c0, flags);
        if (fp.want_rmsd) {
            k_unwrap<true><<<ge, kThreads, 0, cudaStreamTailLaunch>>>(fv, g, fp.c0, partials, tickets, fp.com, flags);
            k_cov<<<gc, kThreads, 0, cudaStreamTailLaunch>>>(fv, g, ref, fp.com, partials, tickets, fp.rmsd_out, fp.rot_out, flags);
        }
        if (fp.want_center) {
            if (fp.center_weighted) k_unwrap<true><<<ge, kThreads, 0, cudaStreamTailLaunch>>>(fv, g, fp.c0, partials, tickets, fp.center_out, flags);
            else k_unwrap<false><<<ge, kThreads, 0, cudaStreamTailLaunch>>>(fv, g, fp.c0, partials, tickets, fp.center_out, flags);
        }
    }
    if (n_second) {
        // the sine-sum centre pass over the frames the moments could not certify BEYOND the host-launched pass's capacity (tail
        // launches run in order: behind the exact passes above, which have consumed this launch's lists by then).  It is itself
        // a single-pass kernel with a device-side fallback: a centre-only plan, no third tier.
        FallbackPlan f2 = fp;
        f2.want_rmsd = 0;
        f2.want_center = 1;
        f2.second_count = nullptr;
        f2.feedback = nullptr;  // what happens to a handful of frames says nothing about the group
        f2.n_report = (int)n_second;
        const dim3 gs(fp.nb_second, n_second);
        const int *rest = fp.second_list + fp.second_cap;
        if (fp.center_weighted)
            k_center_quad<true><<<gs, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, f2, rest, 2, nullptr);
        else
            k_center_quad<false><<<gs, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, f2, rest, 2, nullptr);
    }
}

} // namespace groan

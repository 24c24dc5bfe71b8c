// kernels_tma.cuh -- the single-pass centre / RMSD kernels for contiguous groups, fed by the TMA.
//
// A contiguous group of a frame is a contiguous byte range of the AoS coordinate buffer, so it can be moved
// global -> shared memory by 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier:
// no registers and no issue slots are spent on addresses or on keeping loads in flight, and the number of
// bytes in flight is set by the ring depth, not by occupancy.  ncu on the register-staged version of these
// kernels (profiles/r1_v2_*.md) showed exactly that limit: 75 % of the warp stalls were long-scoreboard
// waits with 16 resident warps per SM.
//
//   CTA = 8 warps, all consumers.  Ring of 3-4 stages.  full[s] (count 1 + tx bytes) is armed by whoever issues the
//   copies and completed by the TMA; the last warp to finish reading a stage (an atomic counter in shared memory)
//   re-arms it and issues the copies for the chunk one ring ahead, so no warp is spent polling for free stages.
//
//   A CTA serves FPC frames at once (FPC = 4 when the batch allows it, else 1): its warps are split into FPC
//   groups, group q streams frame q's chunk, and ALL groups share one copy of the RMSD reference chunk of the
//   stage.  For a 4M-atom group the prepared reference (64 MB) lives in L2 and every frame needs all of it:
//   sharing it between 4 frames cuts the reference's L2 -> SM traffic from 16 to 4 B per atom per frame
//   (profiles/r1_summary.md; the cluster-multicast alternative was measured slower, r1_multicast_experiment.md).
//   Stage = reference blocks (12 KB for 512 atoms / 20 KB for 1024) + FPC x chunk x 12 B.
//   Coordinates are read from shared memory with stride-3 LDS.32 (3 is coprime to 32: conflict-free), the
//   reference with LDS.128.  Frame bytes carry an L2 evict-first policy, reference bytes evict-last, so the
//   64 MB reference of the 4M-atom workload stays in the 126 MB L2 while 48 MB frames stream through.
//
// The arithmetic and the certification logic are those of k_center_fast / k_rmsd_fast (kernels_center.cuh,
// kernels_rmsd.cuh); only the data movement differs.
#pragma once
#include "common.cuh"
#include "kernels_center.cuh"
#include "kernels_rmsd.cuh"

namespace groan {

constexpr int kTmaThreads = kThreads; // 8 warps, all consumers; the last warp to leave a stage refills it

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

// ring geometry: FPC frames per CTA, CH atoms per chunk and frame
template <bool WITH_REF, int STAGES, int FPC>
struct TmaCfg {
    static constexpr int CH = (FPC == 1) ? 1024 : 512;         // atoms per chunk (per frame)
    static constexpr int GT = kThreads / FPC;                  // threads serving one frame
    static constexpr size_t kRefBytes = WITH_REF ? (size_t)(CH / kRefBlock + 1) * kRefBlock * 16 : 0; // whole blocks, +1 if unaligned
    static constexpr size_t kFrameBytes = (size_t)CH * 12;
    static constexpr size_t kStageBytes = kRefBytes + FPC * kFrameBytes;
    static constexpr size_t kBytes = STAGES * kStageBytes + 128;
};
constexpr int kCenterStages = 4; // 48 KB of ring: 4 CTAs per SM
constexpr int kRmsdStages = 3;   // 96 KB (FPC = 1) / 108 KB (FPC = 4) of ring: 2 CTAs per SM

template <int STAGES>
struct TmaCtl {
    uint64_t full[STAGES];     // count 1 + tx bytes: armed by whoever issues the copies, completed by the TMA
    unsigned int done[STAGES]; // consumer warps that have finished reading the stage
};

// Geometry of a contiguous group inside frame f: `head` atoms before the first 16-byte boundary, a body whose
// length is a multiple of 4 atoms (so every chunk and its byte count are 16-byte multiples), then `tail` atoms.
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

struct RefPair {
    float2 x, y, z, w; // (atom i0, atom i1) per component
};

// Stream the body of the group through the ring, two atoms per call, for the frame this thread's warp group serves
// (frame f0 + threadIdx.x / GT):
//   fn(i0, i1, X, Y, Z, ref)   with X = (x of atom i0, x of atom i1) etc. and ref the reference (pc.xyz, w) of the
// two atoms, component-wise paired (zero when !WITH_REF).  Pairs feed the packed f32x2 arithmetic of sm_100 (FADD2 /
// FMUL2 / FFMA2: one issue slot for two lanes of work).  The two atoms of a pair sit half a chunk apart, so every
// LDS.32 of a warp walks consecutive atoms (stride 3 words for coordinates, stride 1 inside a reference block:
// conflict-free) and lands in a register pair the packed ops can use.
// The up-to-3 atoms before and after the 16-byte aligned body are NOT visited; the finishing thread adds them.
// All FPC frames of a CTA must share the body geometry (the host guarantees it: FPC > 1 only if n_atoms % 4 == 0).
template <bool WITH_REF, int STAGES, int FPC, typename F>
__device__ __forceinline__ void stream_pairs_tma(const FrameView &fv, const GroupView &g, int f0, const BodyGeom &bg, const float *ref_pc,
                                                 unsigned char *smem, TmaCtl<STAGES> &ctl, F &&fn) {
    typedef TmaCfg<WITH_REF, STAGES, FPC> C;
    constexpr int CH = C::CH, GT = C::GT;
    const int lane = threadIdx.x & 31;
    const int q = threadIdx.x / GT, tg = threadIdx.x % GT; // frame slot of this thread, index inside its group
    const uint32_t chunks = (bg.body + CH - 1) / CH;
    const uint32_t my_chunks = chunks > blockIdx.x ? (chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint64_t pol_frame = l2_policy_evict_first(), pol_ref = l2_policy_evict_last();
    auto issue = [&](uint32_t it) { // copies of this CTA's chunk `it` (reference once, FPC frames) into stage it % STAGES
        const uint32_t s = it % STAGES, c = blockIdx.x + it * gridDim.x;
        const uint32_t atoms = min((uint32_t)CH, bg.body - c * CH);
        const uint32_t i0 = bg.head + c * CH, b0 = i0 >> 8; // reference blocks covering group atoms [i0, i0 + atoms)
        const uint32_t ref_bytes = WITH_REF ? (((i0 + atoms - 1) >> 8) - b0 + 1) * (uint32_t)(kRefBlock * 16) : 0u;
        mbar_expect_tx(ctl.full + s, atoms * 12u * FPC + ref_bytes);
        unsigned char *dst = smem + s * C::kStageBytes;
        if (WITH_REF && ref_bytes) bulk_g2s(dst, ref_pc + (size_t)b0 * (4 * kRefBlock), ref_bytes, ctl.full + s, pol_ref);
#pragma unroll
        for (int k = 0; k < FPC; k++) {
            const char *src = reinterpret_cast<const char *>(fv.frame(f0 + k) + ((size_t)g.first + bg.head) * 3);
            bulk_g2s(dst + C::kRefBytes + k * C::kFrameBytes, src + (size_t)c * CH * 12, atoms * 12u, ctl.full + s, pol_frame);
        }
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(ctl.full + s, 1);
            ctl.done[s] = 0;
        }
        fence_mbar_init();
        for (uint32_t it = 0; it < (uint32_t)STAGES && it < my_chunks; it++) issue(it);
    }
    __syncthreads();
    for (uint32_t it = 0; it < my_chunks; it++) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t c = blockIdx.x + it * gridDim.x;
        const uint32_t atoms = min((uint32_t)CH, bg.body - c * CH); // a multiple of 4
        const float *sr = reinterpret_cast<const float *>(smem + s * C::kStageBytes);
        const float *sf = reinterpret_cast<const float *>(smem + s * C::kStageBytes + C::kRefBytes + q * C::kFrameBytes);
        const uint32_t i0 = bg.head + c * CH;
        const uint32_t lo = i0 & (kRefBlock - 1); // position of the chunk's first atom inside its reference block
        mbar_wait(ctl.full + s, ph);
        auto pair = [&](uint32_t j0, uint32_t j1) {
            const float2 X = make_float2(sf[3 * j0], sf[3 * j1]), Y = make_float2(sf[3 * j0 + 1], sf[3 * j1 + 1]),
                         Z = make_float2(sf[3 * j0 + 2], sf[3 * j1 + 2]);
            RefPair r;
            if (WITH_REF) {
                const uint32_t w0 = (uint32_t)ref_word(lo + j0), w1 = (uint32_t)ref_word(lo + j1);
                r.x = make_float2(sr[w0], sr[w1]);
                r.y = make_float2(sr[w0 + kRefBlock], sr[w1 + kRefBlock]);
                r.z = make_float2(sr[w0 + 2 * kRefBlock], sr[w1 + 2 * kRefBlock]);
                r.w = make_float2(sr[w0 + 3 * kRefBlock], sr[w1 + 3 * kRefBlock]);
            } else {
                r.x = r.y = r.z = r.w = make_float2(0.f, 0.f);
            }
            fn(i0 + j0, i0 + j1, X, Y, Z, r);
        };
        if (atoms == CH) {
#pragma unroll
            for (int u = 0; u < CH / 2 / GT; u++) pair(tg + u * GT, tg + u * GT + CH / 2);
        } else {
            const uint32_t half = atoms >> 1;
            for (uint32_t j = tg; j < half; j += GT) pair(j, j + half);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&ctl.done[s], 1u) == (unsigned)(kWarps - 1)) { // last reader of the stage: refill it
                ctl.done[s] = 0;
                __threadfence_block();
                if (it + STAGES < my_chunks) issue(it + STAGES);
            }
        }
    }
}

// packed helpers (sm_100 f32x2 pipe)
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
// min-image displacement from the pilot for two atoms at once; same arithmetic as pilot_delta (kernels_center.cuh)
__device__ __forceinline__ float2 pilot_delta2(float2 x, float negp, float L, float invL) {
    const float2 d = __fadd2_rn(x, splat(negp));
    const float2 k = __fadd2_rn(__ffma2_rn(d, splat(invL), splat(12582912.0f)), splat(-12582912.0f));
    return __ffma2_rn(splat(-L), k, d);
}

// Per-frame reduction for a CTA that serves FPC frames: one frame_reduce per frame slot, in which the threads of the other
// slots contribute neutral values.  Returns, for the calling thread's own slot, whether this CTA drew the frame's last
// ticket; the totals of the thread's own frame are left in tot / tmn / tmx.
template <int KS, int FPC>
__device__ __forceinline__ bool multi_frame_reduce(const float (&a)[KS], const float (&mn)[3], const float (&mx)[3], int f0, int nb,
                                                   double *partials, unsigned int *tickets, FrameReduceSmem<KS, 3> &sm,
                                                   double (&tot)[KS], float (&tmn)[3], float (&tmx)[3]) {
    constexpr int GT = kThreads / FPC;
    const int q = threadIdx.x / GT;
    bool mine_last = false;
#pragma unroll 1
    for (int k = 0; k < FPC; k++) {
        float b[KS], bmn[3], bmx[3];
#pragma unroll
        for (int i = 0; i < KS; i++) b[i] = (FPC == 1 || q == k) ? a[i] : 0.0f;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            bmn[i] = (FPC == 1 || q == k) ? mn[i] : 3.0e38f;
            bmx[i] = (FPC == 1 || q == k) ? mx[i] : -3.0e38f;
        }
        double t[KS];
        float t0[3], t1[3];
        const bool last = frame_reduce<KS, 3>(b, bmn, bmx, partials + (size_t)(f0 + k) * nb * (KS + 6), tickets + f0 + k, nb, sm, t, t0, t1);
        if (last && q == k) {
            mine_last = true;
#pragma unroll
            for (int i = 0; i < KS; i++) tot[i] = t[i];
#pragma unroll
            for (int i = 0; i < 3; i++) { tmn[i] = t0[i]; tmx[i] = t1[i]; }
        }
        __syncthreads(); // sm is reused by the next slot
    }
    return mine_last;
}

// ---------------------------------------------------------------- device-side launch of the fallback passes
// The reference-order passes are needed only for frames the single pass could not certify -- usually none.  Launching
// them from the host every time costs four grids of early-exiting CTAs per call (~17 us of a 145 us step,
// profiles/r1_launches.csv).  Instead the thread that finishes the LAST frame of the launch looks at the flags and,
// only if one is set, tail-launches the passes from the device (CUDA dynamic parallelism, cudaStreamTailLaunch: they
// run in order after this grid, before anything the host enqueues next on the stream).
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
    int *second_flags;          // per frame, written by the fused kernel (diagnostics, and the gate of the host-launched pass)
    unsigned int *second_count; // device-launched pass: number of such frames of the running launch (re-armed by the thread that reads it)
    int *second_list;           // ... and which (any order): the pass is launched over exactly these frames
    int nb_second;              // CTAs per frame of that k_center_quad launch
    int second_smem;            // its dynamic shared memory
    int n_report;               // finishing threads that will call maybe_launch_fallback (n_frames, or the length of the list)
};

// sel_mode 0: every frame of the batch; 1: frames with sel[f] != 0 (host-launched second tier); 2: frames sel[0 .. gridDim.y)
template <bool WEIGHTED>
__global__ void k_center_quad(FrameView fv, GroupView g, double *partials, unsigned int *tickets, float *out, int *flags, FallbackPlan fp,
                              const int *sel, int sel_mode);

// frames_done counts finished frames in its low 16 bits and flagged ones above, so that the thread that finishes the last
// frame learns from its own atomic whether any frame needs the passes: no read-back of the flags (8 dependent L2 round
// trips at the very end of the kernel, ~3 us, in the first version).  Batches are far below 65535 frames (kPartialSlots).
__device__ __forceinline__ void maybe_launch_fallback(const FallbackPlan &fp, const FrameView &fv, const GroupView &g, const RefView &ref,
                                                      double *partials, unsigned int *tickets, int *flags, int my_flag, int my_second = 0,
                                                      int my_frame = 0) {
    if (!fp.enabled) return;
    if (my_second) fp.second_list[atomicAdd(fp.second_count, 1u)] = my_frame;
    __threadfence();
    const unsigned int done = atomicAdd(fp.frames_done, my_flag ? 0x10001u : 1u);
    if ((done & 0xffffu) != (unsigned)fp.n_report - 1) return;
    *fp.frames_done = 0u; // re-arm
    const bool slow = my_flag || (done >> 16) != 0u;
    const unsigned int n_second = fp.second_count != nullptr ? atomicExch(fp.second_count, 0u) : 0u;
    if (!slow && n_second == 0u) return;
    __threadfence();
    const dim3 ge(fp.nb_exact, fp.n_frames), gc(fp.nb_cov, fp.n_frames);
    if (slow) {
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
        // the sine-sum centre pass over exactly the frames the moments could not certify (tail launches run in order: behind
        // the reference-order passes above, which have consumed this launch's flags by then).  It is itself a single-pass
        // kernel with a device-side fallback: a centre-only plan, no third tier.
        FallbackPlan f2 = fp;
        f2.want_rmsd = 0;
        f2.want_center = 1;
        f2.second_count = nullptr;
        f2.n_report = (int)n_second;
        const dim3 gs(fp.nb_second, n_second);
        if (fp.center_weighted)
            k_center_quad<true><<<gs, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, f2, fp.second_list, 2);
        else
            k_center_quad<false><<<gs, 256, fp.second_smem, cudaStreamTailLaunch>>>(fv, g, partials, tickets, fp.center_out, flags, f2, fp.second_list, 2);
    }
}

// ---------------------------------------------------------------- group_get_center / group_get_com, single pass
// per-thread sums (float2 = one partial per atom of the pair): [0..2] sum m d, [3] sum m, [4..6] sum cos, [7..9] sum sin
template <bool WEIGHTED>
__global__ void __launch_bounds__(kTmaThreads, 4) k_center_tma(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                                float *out, int *flags, FallbackPlan fp) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<10, 3> sm;
    __shared__ TmaCtl<kCenterStages> ctl;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = fr + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const float sx = pi_x2() * ix, sy = pi_x2() * iy, sz = pi_x2() * iz;
    const BodyGeom bg = body_geom(fv, g, f);
    float2 a2[10];
#pragma unroll
    for (int k = 0; k < 10; k++) a2[k] = make_float2(0.f, 0.f);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_pairs_tma<false, kCenterStages, 1>(fv, g, f, bg, nullptr, dyn_smem, ctl,
                                               [&](uint32_t i0, uint32_t i1, float2 X, float2 Y, float2 Z, const RefPair &) {
        const float2 dx = pilot_delta2(X, -px, L[0], ix), dy = pilot_delta2(Y, -py, L[1], iy), dz = pilot_delta2(Z, -pz, L[2], iz);
        if (WEIGHTED) {
            const float2 m = make_float2(__ldg(g.mass + i0), __ldg(g.mass + i1));
            a2[0] = __ffma2_rn(m, dx, a2[0]); a2[1] = __ffma2_rn(m, dy, a2[1]); a2[2] = __ffma2_rn(m, dz, a2[2]);
            a2[3] = __fadd2_rn(a2[3], m);
        } else {
            a2[0] = __fadd2_rn(a2[0], dx); a2[1] = __fadd2_rn(a2[1], dy); a2[2] = __fadd2_rn(a2[2], dz);
        }
        const float2 tx = __fmul2_rn(dx, splat(sx)), ty = __fmul2_rn(dy, splat(sy)), tz = __fmul2_rn(dz, splat(sz));
        float2 s, c;
        __sincosf(tx.x, &s.x, &c.x); __sincosf(tx.y, &s.y, &c.y);
        a2[4] = __fadd2_rn(a2[4], c); a2[7] = __fadd2_rn(a2[7], s);
        __sincosf(ty.x, &s.x, &c.x); __sincosf(ty.y, &s.y, &c.y);
        a2[5] = __fadd2_rn(a2[5], c); a2[8] = __fadd2_rn(a2[8], s);
        __sincosf(tz.x, &s.x, &c.x); __sincosf(tz.y, &s.y, &c.y);
        a2[6] = __fadd2_rn(a2[6], c); a2[9] = __fadd2_rn(a2[9], s);
        mn[0] = fminf(mn[0], fminf(dx.x, dx.y)); mx[0] = fmaxf(mx[0], fmaxf(dx.x, dx.y));
        mn[1] = fminf(mn[1], fminf(dy.x, dy.y)); mx[1] = fmaxf(mx[1], fmaxf(dy.x, dy.y));
        mn[2] = fminf(mn[2], fminf(dz.x, dz.y)); mx[2] = fmaxf(mx[2], fmaxf(dz.x, dz.y));
    });
    float a[10];
#pragma unroll
    for (int k = 0; k < 10; k++) a[k] = a2[k].x + a2[k].y;
    double tot[10];
    float tmn[3], tmx[3];
    if (frame_reduce<10, 3>(a, mn, mx, partials + (size_t)f * nb * 16, tickets + f, nb, sm, tot, tmn, tmx) && threadIdx.x == 0) {
        // the up-to-6 atoms outside the 16-byte aligned body, in f64 with the same definitions
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            const float pp[3] = {px, py, pz};
            const double m = WEIGHTED ? (double)__ldg(g.mass + i) : 1.0;
            if (WEIGHTED) tot[3] += m;
            for (int k = 0; k < 3; k++) {
                const float d = pilot_delta(__ldg(q + k), pp[k], L[k], 1.0f / L[k]);
                tot[k] += m * (double)d;
                float sn, cs;
                sincosf(d * (6.2831853f / L[k]), &sn, &cs);
                tot[4 + k] += (double)cs;
                tot[7 + k] += (double)sn;
                tmn[k] = fminf(tmn[k], d);
                tmx[k] = fmaxf(tmx[k], d);
            }
        }
        int flag = 0;
        finish_center<WEIGHTED>(tot, tmn, tmx, px, py, pz, L, g.n, out + f * 3, &flag);
        flags[f] = flag;
        maybe_launch_fallback(fp, fv, g, RefView(), partials, tickets, flags, flag);
    }
}

// ---------------------------------------------------------------- calc_rmsd (+ optionally the centre), single pass
// CENTER: 0 = RMSD only, 1 = also group_get_center (geometric), 2 = also group_get_com (mass-weighted = the COM the RMSD
// needs anyway).  A trajectory analysis usually wants several per-frame quantities of the same group; each extra pass
// costs another 12 B/atom of HBM, so the fused variants take them from one read of the frame.
// per-thread sums as float2 (one partial per atom of the pair): [0..25] as kFastSums of kernels_rmsd.cuh,
// then (CENTER != 0) [26..28] sum d (geometric centre), [29..31] sum cos, [32..34] sum sin
constexpr int kFusedSums = kFastSums + 9;

template <bool SAME_MASS, int CENTER, int FPC>
__global__ void __launch_bounds__(kTmaThreads, 2) k_rmsd_tma(FrameView fv, GroupView g, RefView ref, double *partials,
                                                              unsigned int *tickets, float *center_out, float *rmsd_out, float *rot_out,
                                                              float *com_out, int *flags, FallbackPlan fp) {
    constexpr int KS = CENTER ? kFusedSums : kFastSums;
    constexpr int GT = kThreads / FPC;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<KS, 3> sm;
    __shared__ TmaCtl<kRmsdStages> ctl;
    const int f0 = blockIdx.y * FPC, f = f0 + threadIdx.x / GT, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = fr + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const float sc[3] = {pi_x2() * ix, pi_x2() * iy, pi_x2() * iz};
    const BodyGeom bg = body_geom(fv, g, f0);
    float2 a2[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) a2[k] = make_float2(0.f, 0.f);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_pairs_tma<true, kRmsdStages, FPC>(fv, g, f0, bg, ref.pc, dyn_smem, ctl,
                                              [&](uint32_t i0, uint32_t i1, float2 X, float2 Y, float2 Z, const RefPair &r) {
        const float2 d[3] = {pilot_delta2(X, -px, L[0], ix), pilot_delta2(Y, -py, L[1], iy), pilot_delta2(Z, -pz, L[2], iz)};
        const float2 pc[3] = {r.x, r.y, r.z};
        const float2 w = r.w;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const float2 wp = __fmul2_rn(w, pc[u]);
#pragma unroll
            for (int v = 0; v < 3; v++) {
                a2[u * 3 + v] = __ffma2_rn(pc[u], d[v], a2[u * 3 + v]);
                a2[9 + u * 3 + v] = __ffma2_rn(wp, d[v], a2[9 + u * 3 + v]);
            }
        }
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const float2 wd = __fmul2_rn(w, d[v]);
            a2[18 + v] = __fadd2_rn(a2[18 + v], wd);
            a2[21] = __ffma2_rn(wd, d[v], a2[21]);
            mn[v] = fminf(mn[v], fminf(d[v].x, d[v].y));
            mx[v] = fmaxf(mx[v], fmaxf(d[v].x, d[v].y));
            if (CENTER) {
                if (CENTER == 1) a2[KS - 9 + v] = __fadd2_rn(a2[KS - 9 + v], d[v]);
                const float2 th = __fmul2_rn(d[v], splat(sc[v]));
                float2 s, c;
                __sincosf(th.x, &s.x, &c.x);
                __sincosf(th.y, &s.y, &c.y);
                a2[KS - 6 + v] = __fadd2_rn(a2[KS - 6 + v], c);
                a2[KS - 3 + v] = __fadd2_rn(a2[KS - 3 + v], s);
            }
        }
        if (!SAME_MASS) {
            const float2 m = make_float2(__ldg(g.mass + i0), __ldg(g.mass + i1));
#pragma unroll
            for (int v = 0; v < 3; v++) a2[22 + v] = __ffma2_rn(m, d[v], a2[22 + v]);
            a2[25] = __fadd2_rn(a2[25], m);
        }
    });
    float a[KS];
#pragma unroll
    for (int k = 0; k < KS; k++) a[k] = a2[k].x + a2[k].y;
    double tot[KS];
    float tmn[3], tmx[3];
    const bool last = multi_frame_reduce<KS, FPC>(a, mn, mx, f0, nb, partials, tickets, sm, tot, tmn, tmx);
    if (last && threadIdx.x % GT == 0) {
        // the up-to-6 atoms outside the 16-byte aligned body, in f64 with the same definitions
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            const float4 r = ref_at(ref.pc, i);
            const float pp[3] = {px, py, pz};
            const double pcd[3] = {(double)r.x, (double)r.y, (double)r.z}, w = (double)r.w;
            double d[3];
            for (int k = 0; k < 3; k++) {
                const float dk = pilot_delta(__ldg(q + k), pp[k], L[k], 1.0f / L[k]);
                d[k] = (double)dk;
                tmn[k] = fminf(tmn[k], dk);
                tmx[k] = fmaxf(tmx[k], dk);
                if (CENTER) {
                    float sn, cs;
                    sincosf(dk * (6.2831853f / L[k]), &sn, &cs);
                    if (CENTER == 1) tot[KS - 9 + k] += d[k];
                    tot[KS - 6 + k] += (double)cs;
                    tot[KS - 3 + k] += (double)sn;
                }
            }
            for (int u = 0; u < 3; u++)
                for (int v = 0; v < 3; v++) {
                    tot[u * 3 + v] += pcd[u] * d[v];
                    tot[9 + u * 3 + v] += w * pcd[u] * d[v];
                }
            for (int v = 0; v < 3; v++) {
                tot[18 + v] += w * d[v];
                tot[21] += w * d[v] * d[v];
            }
            if (!SAME_MASS) {
                const double m = (double)__ldg(g.mass + i);
                for (int v = 0; v < 3; v++) tot[22 + v] += m * d[v];
                tot[25] += m;
            }
        }
        double rt[kFastSums];
        for (int k = 0; k < kFastSums; k++) rt[k] = tot[k];
        int flag_r = 0, flag_c = 0;
        finish_rmsd<SAME_MASS>(rt, tmn, tmx, px, py, pz, L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, &flag_r);
        if (CENTER) {
            // centre: geometric (sum d / n) or mass-weighted with the target group's masses (= the COM of the RMSD)
            double ct[10];
            for (int k = 0; k < 3; k++) {
                ct[k] = CENTER == 2 ? (SAME_MASS ? tot[18 + k] : tot[22 + k]) : tot[KS - 9 + k];
                ct[4 + k] = tot[KS - 6 + k];
                ct[7 + k] = tot[KS - 3 + k];
            }
            ct[3] = SAME_MASS ? ref.sum_w : tot[25];
            if (CENTER == 2) finish_center<true>(ct, tmn, tmx, px, py, pz, L, g.n, center_out + f * 3, &flag_c);
            else finish_center<false>(ct, tmn, tmx, px, py, pz, L, g.n, center_out + f * 3, &flag_c);
        }
        flags[f] = flag_r | (flag_c << 1);
        maybe_launch_fallback(fp, fv, g, ref, partials, tickets, flags, flag_r | flag_c);
    }
}

} // namespace groan

// kernels_tma.cuh -- the single-pass centre / RMSD kernels for contiguous groups, fed by the TMA.
//
// A contiguous group of a frame is a contiguous byte range of the AoS coordinate buffer, so it can be moved
// global -> shared memory by 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier:
// no registers and no issue slots are spent on addresses or on keeping loads in flight, and the number of
// bytes in flight is set by the ring depth, not by occupancy.  ncu on the register-staged version of these
// kernels (profiles/r1_v2_*.md) showed exactly that limit: 75 % of the warp stalls were long-scoreboard
// waits with 16 resident warps per SM.
//
//   CTA = 8 warps, all consumers.  Ring of 3-4 stages; a stage holds one chunk of kChunk atoms: 12 KB of
//   coordinates (+ 16 KB of the RMSD reference, float4 per atom).  full[s] (count 1 + tx bytes) is armed by
//   whoever issues the copies and completed by the TMA; the last warp to finish reading a stage (an atomic
//   counter in shared memory) re-arms it and issues the copies for the chunk one ring ahead, so no warp is
//   spent polling for free stages.
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

constexpr int kChunk = 1024;          // atoms per stage
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

// ---- thread-block cluster helpers: the CTAs that process the same chunks of DIFFERENT frames form a cluster along y,
// and the reference chunk they all need is fetched from L2 once and multicast into every CTA's ring.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// atomic add on the counter at the same shared-memory offset in CTA `rank` of the cluster (distributed shared memory);
// acq_rel at cluster scope: every CTA's reads of a stage happen-before the copies issued by the CTA that arrives last
__device__ __forceinline__ uint32_t atomic_add_remote(unsigned int *ctr, uint32_t rank, uint32_t v) {
    uint32_t raddr, old;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(ctr)), "r"(rank));
    asm volatile("atom.acq_rel.cluster.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(raddr), "r"(v) : "memory");
    return old;
}
// bulk copy delivered to the same offset of every CTA in `mask`, completing on each CTA's own mbarrier at that offset
__device__ __forceinline__ void bulk_g2s_multicast(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask,
                                                   uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask), "l"(policy)
        : "memory");
}

template <bool WITH_REF, int STAGES>
struct TmaSmem {
    static constexpr size_t kFrameBytes = (size_t)kChunk * 12;
    // reference: whole blocks of kRefBlock atoms; a chunk that does not start on a block boundary touches one more
    static constexpr size_t kRefBytes = WITH_REF ? (size_t)(kChunk / kRefBlock + 1) * kRefBlock * 16 : 0;
    static constexpr size_t kStageBytes = kFrameBytes + kRefBytes;
    static constexpr size_t kBytes = STAGES * kStageBytes + 128;
};
constexpr int kCenterStages = 4; // 48 KB of ring: 4 CTAs per SM
constexpr int kRmsdStages = 3;   // 96 KB of ring: 2 CTAs per SM

template <int STAGES>
struct TmaCtl {
    uint64_t full[STAGES];     // count 1 + tx bytes: armed by whoever issues the copies, completed by the TMA
    unsigned int done[STAGES];  // consumer warps that have finished reading the stage
    unsigned int cdone[STAGES]; // cluster rank 0's copy only: CTAs of the cluster that have finished reading the stage
};

// Geometry of a contiguous group inside frame f: `head` atoms before the first 16-byte boundary, a body whose
// length is a multiple of 4 atoms (so every chunk and its byte count are 16-byte multiples), then `tail` atoms.
struct BodyGeom {
    uint32_t head, body, tail, chunks;
};
__device__ __forceinline__ BodyGeom body_geom(const FrameView &fv, const GroupView &g, int f) {
    BodyGeom b;
    const size_t a0 = (size_t)f * fv.n_atoms + g.first;
    b.head = (uint32_t)((4 - (a0 & 3)) & 3);
    if (b.head > g.n) b.head = g.n;
    b.body = (g.n - b.head) & ~3u;
    b.tail = g.n - b.head - b.body;
    b.chunks = (b.body + kChunk - 1) / kChunk;
    return b;
}

// Stream the body of the group in frame f through the ring, two atoms per call:
//   fn(i0, i1, X, Y, Z, ref)   with X = (x of atom i0, x of atom i1) etc. and ref the reference (pc.xyz, w) of the
// two atoms, component-wise paired (zero when !WITH_REF).  Pairs feed the packed f32x2 arithmetic of sm_100 (FADD2 / FMUL2 / FFMA2: one
// issue slot for two lanes of work), which is what lifts these kernels from issue-bound to HBM-bound
// (profiles/r1_v4_*).  The two atoms of a pair sit half a chunk apart, so every LDS.32 of a warp walks
// consecutive atoms (stride 3 words, conflict-free) and lands in a register pair the packed ops can use.
// The up-to-3 atoms before and after the 16-byte aligned body are NOT visited; the finishing thread adds them.
//
// There is no producer warp and nobody polls for a free stage: a warp that has read its share of a stage bumps
// done[s]; the warp that brings it to 8 is the last reader, so it re-arms full[s] and issues the copies of the
// chunk STAGES ahead into the stage it has just emptied.  Consumers only ever wait on full[s].
// profiling experiment only (GROAN_DEBUG_SKIP_REF=1): do not copy the reference into the ring, to measure how much
// of the RMSD kernel's time is its L2 -> SM traffic.  Results are garbage with it set; never set in tests or bench.
__device__ int g_debug_skip_ref = 0;

struct RefPair {
    float2 x, y, z, w; // (atom i0, atom i1) per component
};

template <bool WITH_REF, int STAGES, typename F>
__device__ __forceinline__ void stream_pairs_tma(const FrameView &fv, const GroupView &g, int f, const BodyGeom &bg,
                                                 const float *ref_pc, unsigned char *smem, TmaCtl<STAGES> &ctl, int cs, F &&fn) {
    // cs = cluster size along y (frames).  cs > 1: all CTAs of the cluster walk the same chunk sequence (same blockIdx.x,
    // same body geometry -- the host guarantees it) and share one L2 read of each reference chunk: the CTA that is the
    // LAST of the cluster to finish a stage (a counter in rank 0's shared memory) multicasts the next chunk into it.
    typedef TmaSmem<WITH_REF, STAGES> S;
    const int lane = threadIdx.x & 31;
    const bool mc = WITH_REF && cs > 1;
    const uint32_t rank = mc ? cluster_rank() : 0u;
    const float *fr = fv.frame(f);
    const char *src = reinterpret_cast<const char *>(fr + ((size_t)g.first + bg.head) * 3);
    const uint32_t my_chunks = bg.chunks > blockIdx.x ? (bg.chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint64_t pol_frame = l2_policy_evict_first(), pol_ref = l2_policy_evict_last();
    // byte counts of this CTA's chunk `it`
    auto geom = [&](uint32_t it, uint32_t &c, uint32_t &atoms, uint32_t &b0, uint32_t &ref_bytes) {
        c = blockIdx.x + it * gridDim.x;
        atoms = min((uint32_t)kChunk, bg.body - c * kChunk);
        const uint32_t i0 = bg.head + c * kChunk;
        b0 = i0 >> 8; // reference blocks covering group atoms [i0, i0 + atoms)
        ref_bytes = (WITH_REF && !g_debug_skip_ref) ? (((i0 + atoms - 1) >> 8) - b0 + 1) * (uint32_t)(kRefBlock * 16) : 0u;
    };
    // arm full[s] for everything that will land in the stage and copy this CTA's own frame chunk (+ reference if not multicast)
    auto issue_own = [&](uint32_t it) {
        uint32_t c, atoms, b0, ref_bytes;
        geom(it, c, atoms, b0, ref_bytes);
        const uint32_t s = it % STAGES;
        mbar_expect_tx(ctl.full + s, atoms * 12u + ref_bytes);
        unsigned char *dst = smem + s * S::kStageBytes;
        bulk_g2s(dst, src + (size_t)c * kChunk * 12, atoms * 12u, ctl.full + s, pol_frame);
        if (WITH_REF && !mc && ref_bytes)
            bulk_g2s(dst + S::kFrameBytes, ref_pc + (size_t)b0 * (4 * kRefBlock), ref_bytes, ctl.full + s, pol_ref);
    };
    // one L2 read of the reference chunk, delivered to every CTA of the cluster
    auto issue_multicast = [&](uint32_t it) {
        uint32_t c, atoms, b0, ref_bytes;
        geom(it, c, atoms, b0, ref_bytes);
        const uint32_t s = it % STAGES;
        if (ref_bytes)
            bulk_g2s_multicast(smem + s * S::kStageBytes + S::kFrameBytes, ref_pc + (size_t)b0 * (4 * kRefBlock), ref_bytes,
                               ctl.full + s, (uint16_t)((1u << cs) - 1u), pol_ref);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(ctl.full + s, 1);
            ctl.done[s] = 0;
            ctl.cdone[s] = 0;
        }
        fence_mbar_init();
    }
    if (mc) cluster_sync_all(); // every CTA's barriers exist before any multicast can signal them
    else __syncthreads();
    if (threadIdx.x == 0) {
        for (uint32_t it = 0; it < (uint32_t)STAGES && it < my_chunks; it++) {
            issue_own(it);
            if (mc && rank == 0) issue_multicast(it);
        }
    }
    for (uint32_t it = 0; it < my_chunks; it++) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t c = blockIdx.x + it * gridDim.x;
        const uint32_t atoms = min((uint32_t)kChunk, bg.body - c * kChunk); // a multiple of 4
        const float *sf = reinterpret_cast<const float *>(smem + s * S::kStageBytes);
        const float *sr = reinterpret_cast<const float *>(smem + s * S::kStageBytes + S::kFrameBytes);
        const uint32_t i0 = bg.head + c * kChunk;
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
        if (atoms == kChunk) {
#pragma unroll
            for (int u = 0; u < kChunk / 2 / kThreads; u++) pair(threadIdx.x + u * kThreads, threadIdx.x + u * kThreads + kChunk / 2);
        } else {
            const uint32_t half = atoms >> 1;
            for (uint32_t j = threadIdx.x; j < half; j += kThreads) pair(j, j + half);
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&ctl.done[s], 1u) == (unsigned)(kWarps - 1)) { // last reader of the stage: refill it
                ctl.done[s] = 0;
                __threadfence_block();
                if (it + STAGES < my_chunks) {
                    issue_own(it + STAGES);
                    // this CTA no longer reads the stage's reference and has armed full[s]; the last CTA to say so multicasts
                    if (mc && atomic_add_remote(&ctl.cdone[s], 0, 1u) % (uint32_t)cs == (uint32_t)(cs - 1)) issue_multicast(it + STAGES);
                }
            }
        }
    }
    if (mc) cluster_sync_all(); // no CTA may exit while a peer can still signal its barriers
}

// packed helpers (sm_100 f32x2 pipe)
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }
// min-image displacement from the pilot for two atoms at once; same arithmetic as pilot_delta (kernels_center.cuh)
__device__ __forceinline__ float2 pilot_delta2(float2 x, float negp, float L, float invL) {
    const float2 d = __fadd2_rn(x, splat(negp));
    const float2 k = __fadd2_rn(__ffma2_rn(d, splat(invL), splat(12582912.0f)), splat(-12582912.0f));
    return __ffma2_rn(splat(-L), k, d);
}

// ---------------------------------------------------------------- device-side launch of the fallback passes
// The reference-order passes are needed only for frames the single pass could not certify -- usually none.  Launching
// them from the host every time costs four grids of early-exiting CTAs per call (~17 us of a 145 us step,
// profiles/r1_launches.csv).  Instead the thread that finishes the LAST frame of the launch looks at the flags and,
// only if one is set, tail-launches the passes from the device (CUDA dynamic parallelism, cudaStreamTailLaunch: they
// run in order after this grid, before anything the host enqueues next on the stream).
struct FallbackPlan {
    int enabled;              // 0: the host launches the fallback passes itself
    int nb_exact, nb_cov;     // CTAs per frame of k_trig / k_unwrap and of k_cov
    unsigned int *frames_done;
    float *c0;                // Bai-Breen estimates of the flagged frames
    int want_center, center_weighted, want_rmsd;
    float *center_out;        // want_center
    float *com, *rmsd_out, *rot_out; // want_rmsd
};

__device__ __forceinline__ void maybe_launch_fallback(const FallbackPlan &fp, const FrameView &fv, const GroupView &g, const RefView &ref,
                                                      double *partials, unsigned int *tickets, const int *flags) {
    if (!fp.enabled) return;
    __threadfence();
    const unsigned int done = atomicAdd(fp.frames_done, 1u);
    if (done != gridDim.y - 1) return;
    *fp.frames_done = 0u; // re-arm
    __threadfence();
    int any = 0;
    for (unsigned int f = 0; f < gridDim.y; f++) any |= ((const volatile int *)flags)[f];
    if (!any) return;
    const dim3 ge(fp.nb_exact, gridDim.y), gc(fp.nb_cov, gridDim.y);
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
    stream_pairs_tma<false, kCenterStages>(fv, g, f, bg, nullptr, dyn_smem, ctl, 1,
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
                const double th = (double)d * 6.283185307179586 / (double)L[k];
                tot[4 + k] += cos(th);
                tot[7 + k] += sin(th);
                tmn[k] = fminf(tmn[k], d);
                tmx[k] = fmaxf(tmx[k], d);
            }
        }
        finish_center<WEIGHTED>(tot, tmn, tmx, px, py, pz, L, g.n, out + f * 3, flags + f);
        maybe_launch_fallback(fp, fv, g, RefView(), partials, tickets, flags);
    }
}

// ---------------------------------------------------------------- calc_rmsd, single pass
// per-thread sums as float2 (one partial per atom of the pair), same meaning as kFastSums of kernels_rmsd.cuh
template <bool SAME_MASS>
__global__ void __launch_bounds__(kTmaThreads, 2) k_rmsd_tma(FrameView fv, GroupView g, RefView ref, double *partials,
                                                              unsigned int *tickets, float *rmsd_out, float *rot_out, float *com_out,
                                                              int *flags, FallbackPlan fp, int cs) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<kFastSums, 3> sm;
    __shared__ TmaCtl<kRmsdStages> ctl;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = fr + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const BodyGeom bg = body_geom(fv, g, f);
    float2 a2[kFastSums];
#pragma unroll
    for (int k = 0; k < kFastSums; k++) a2[k] = make_float2(0.f, 0.f);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_pairs_tma<true, kRmsdStages>(fv, g, f, bg, ref.pc, dyn_smem, ctl, cs,
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
        }
        if (!SAME_MASS) {
            const float2 m = make_float2(__ldg(g.mass + i0), __ldg(g.mass + i1));
#pragma unroll
            for (int v = 0; v < 3; v++) a2[22 + v] = __ffma2_rn(m, d[v], a2[22 + v]);
            a2[25] = __fadd2_rn(a2[25], m);
        }
    });
    float a[kFastSums];
#pragma unroll
    for (int k = 0; k < kFastSums; k++) a[k] = a2[k].x + a2[k].y;
    double tot[kFastSums];
    float tmn[3], tmx[3];
    if (frame_reduce<kFastSums, 3>(a, mn, mx, partials + (size_t)f * nb * (kFastSums + 6), tickets + f, nb, sm, tot, tmn, tmx) &&
        threadIdx.x == 0) {
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
        finish_rmsd<SAME_MASS>(tot, tmn, tmx, px, py, pz, L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, flags + f);
        maybe_launch_fallback(fp, fv, g, ref, partials, tickets, flags);
    }
}

// ---------------------------------------------------------------- group_get_center (or group_get_com) + calc_rmsd in ONE pass
// A trajectory analysis usually wants several per-frame quantities of the same group; each extra pass costs
// another 12 B/atom of HBM.  This kernel produces the refined centre (geometric, or mass-weighted = the COM the
// RMSD needs anyway) and the Kabsch RMSD from a single read of the frame.
// sums: [0..25] as k_rmsd_tma, [26..28] sum d (geometric centre), [29..31] sum cos, [32..34] sum sin
constexpr int kFusedSums = kFastSums + 9;

template <bool SAME_MASS, bool WEIGHTED_CENTER>
__global__ void __launch_bounds__(kTmaThreads, 2) k_center_rmsd_tma(FrameView fv, GroupView g, RefView ref, double *partials,
                                                                     unsigned int *tickets, float *center_out, float *rmsd_out,
                                                                     float *rot_out, float *com_out, int *flags, FallbackPlan fp, int cs) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<kFusedSums, 3> sm;
    __shared__ TmaCtl<kRmsdStages> ctl;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = fr + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const float sc[3] = {pi_x2() * ix, pi_x2() * iy, pi_x2() * iz};
    const BodyGeom bg = body_geom(fv, g, f);
    float2 a2[kFusedSums];
#pragma unroll
    for (int k = 0; k < kFusedSums; k++) a2[k] = make_float2(0.f, 0.f);
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_pairs_tma<true, kRmsdStages>(fv, g, f, bg, ref.pc, dyn_smem, ctl, cs,
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
            if (!WEIGHTED_CENTER) a2[26 + v] = __fadd2_rn(a2[26 + v], d[v]);
            const float2 th = __fmul2_rn(d[v], splat(sc[v]));
            float2 s, c;
            __sincosf(th.x, &s.x, &c.x);
            __sincosf(th.y, &s.y, &c.y);
            a2[29 + v] = __fadd2_rn(a2[29 + v], c);
            a2[32 + v] = __fadd2_rn(a2[32 + v], s);
        }
        if (!SAME_MASS) {
            const float2 m = make_float2(__ldg(g.mass + i0), __ldg(g.mass + i1));
#pragma unroll
            for (int v = 0; v < 3; v++) a2[22 + v] = __ffma2_rn(m, d[v], a2[22 + v]);
            a2[25] = __fadd2_rn(a2[25], m);
        }
    });
    float a[kFusedSums];
#pragma unroll
    for (int k = 0; k < kFusedSums; k++) a[k] = a2[k].x + a2[k].y;
    double tot[kFusedSums];
    float tmn[3], tmx[3];
    if (frame_reduce<kFusedSums, 3>(a, mn, mx, partials + (size_t)f * nb * (kFusedSums + 6), tickets + f, nb, sm, tot, tmn, tmx) &&
        threadIdx.x == 0) {
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) { // atoms outside the 16-byte aligned body, in f64
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
                const double th = d[k] * 6.283185307179586 / (double)L[k];
                if (!WEIGHTED_CENTER) tot[26 + k] += d[k];
                tot[29 + k] += cos(th);
                tot[32 + k] += sin(th);
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
        // centre: geometric (sum d / n) or mass-weighted with the target group's masses (= the COM of the RMSD)
        double ct[10];
        for (int k = 0; k < 3; k++) {
            ct[k] = WEIGHTED_CENTER ? (SAME_MASS ? tot[18 + k] : tot[22 + k]) : tot[26 + k];
            ct[4 + k] = tot[29 + k];
            ct[7 + k] = tot[32 + k];
        }
        ct[3] = SAME_MASS ? ref.sum_w : tot[25];
        finish_center<WEIGHTED_CENTER>(ct, tmn, tmx, px, py, pz, L, g.n, center_out + f * 3, &flag_c);
        flags[f] = flag_r | (flag_c << 1);
        maybe_launch_fallback(fp, fv, g, ref, partials, tickets, flags);
    }
}

} // namespace groan

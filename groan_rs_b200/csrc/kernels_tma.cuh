// kernels_tma.cuh -- the single-pass centre / RMSD kernels for contiguous groups, fed by the TMA.
//
// A contiguous group of a frame is a contiguous byte range of the AoS coordinate buffer, so it can be moved
// global -> shared memory by 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier:
// no registers and no issue slots are spent on addresses or on keeping loads in flight, and the number of
// bytes in flight is set by the ring depth, not by occupancy.  ncu on the register-staged version of these
// kernels (profiles/r1_v2_*.md) showed exactly that limit: 75 % of the warp stalls were long-scoreboard
// waits with 16 resident warps per SM.
//
//   CTA = 8 consumer warps + 1 producer warp (one elected lane).  Ring of kStages stages; a stage holds one
//   chunk of kChunk atoms: 12 KB of coordinates (+ 16 KB of the RMSD reference, float4 per atom).
//   full[s]  (count 1 + tx bytes): producer arms it, the TMA completes it.
//   empty[s] (count 8): one arrival per consumer warp after its last read of the stage.
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

constexpr int kChunk = 1024;                 // atoms per stage
constexpr int kStages = 4;                   // ring depth
constexpr int kTmaThreads = kThreads + 32;   // 8 consumer warps + the producer warp
constexpr int kTmaWarps = kTmaThreads / 32;

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
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
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

template <bool WITH_REF>
struct TmaSmem {
    static constexpr size_t kFrameBytes = (size_t)kChunk * 12;
    static constexpr size_t kRefBytes = WITH_REF ? (size_t)kChunk * 16 : 0;
    static constexpr size_t kStageBytes = kFrameBytes + kRefBytes;
    static constexpr size_t kBytes = kStages * kStageBytes + 2 * kStages * sizeof(uint64_t) + 128;
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

// Stream the group's atoms of frame f through the ring.  Consumers call fn(i, x, y, z, ref) for each of their atoms
// (i = position in the group; ref = reference float4, undefined when !WITH_REF); the producer warp only issues copies.
template <bool WITH_REF, typename F>
__device__ __forceinline__ void stream_group_tma(const FrameView &fv, const GroupView &g, int f, const float4 *ref_pc,
                                                 unsigned char *smem_raw, F &&fn) {
    typedef TmaSmem<WITH_REF> S;
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * S::kStageBytes);
    uint64_t *empty = full + kStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BodyGeom bg = body_geom(fv, g, f);
    const float *fr = fv.frame(f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t my_chunks = bg.chunks > blockIdx.x ? (bg.chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (warp == kWarps) {
        // ---------------- producer
        if (lane == 0) {
            const uint64_t pol_frame = l2_policy_evict_first(), pol_ref = l2_policy_evict_last();
            const char *src = reinterpret_cast<const char *>(fr + ((size_t)g.first + bg.head) * 3);
            for (uint32_t it = 0; it < my_chunks; it++) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                const uint32_t c = blockIdx.x + it * gridDim.x;
                const uint32_t atoms = min((uint32_t)kChunk, bg.body - c * kChunk);
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, atoms * (WITH_REF ? 28u : 12u));
                unsigned char *dst = smem + s * S::kStageBytes;
                bulk_g2s(dst, src + (size_t)c * kChunk * 12, atoms * 12u, full + s, pol_frame);
                if (WITH_REF)
                    bulk_g2s(dst + S::kFrameBytes, reinterpret_cast<const char *>(ref_pc + bg.head + (size_t)c * kChunk), atoms * 16u,
                             full + s, pol_ref);
            }
        }
        __syncwarp();
    } else {
        // ---------------- consumers
        for (uint32_t it = 0; it < my_chunks; it++) {
            const uint32_t s = it % kStages, ph = (it / kStages) & 1;
            const uint32_t c = blockIdx.x + it * gridDim.x;
            const uint32_t atoms = min((uint32_t)kChunk, bg.body - c * kChunk);
            const float *sf = reinterpret_cast<const float *>(smem + s * S::kStageBytes);
            const float4 *sr = reinterpret_cast<const float4 *>(smem + s * S::kStageBytes + S::kFrameBytes);
            mbar_wait(full + s, ph);
#pragma unroll
            for (int u = 0; u < kChunk / kThreads; u++) {
                const uint32_t j = threadIdx.x + u * kThreads;
                if (j < atoms) {
                    const float4 r = WITH_REF ? sr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    fn(bg.head + c * kChunk + j, sf[3 * j], sf[3 * j + 1], sf[3 * j + 2], r);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        // the up-to-3 atoms before and after the 16-byte-aligned body
        if (blockIdx.x == 0 && threadIdx.x < bg.head + bg.tail) {
            const uint32_t i = threadIdx.x < bg.head ? threadIdx.x : bg.head + bg.body + (threadIdx.x - bg.head);
            const float *p = fr + ((size_t)g.first + i) * 3;
            const float4 r = WITH_REF ? __ldg(ref_pc + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            fn(i, __ldg(p), __ldg(p + 1), __ldg(p + 2), r);
        }
    }
}

// ---------------------------------------------------------------- group_get_center / group_get_com, single pass
template <bool WEIGHTED>
__global__ void __launch_bounds__(kTmaThreads) k_center_tma(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                             float *out, int *flags) {
    extern __shared__ unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<10, 3, kTmaWarps> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *p0 = fv.frame(f) + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    const float sx = pi_x2() * ix, sy = pi_x2() * iy, sz = pi_x2() * iz;
    float a[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_group_tma<false>(fv, g, f, nullptr, dyn_smem, [&](uint32_t i, float x, float y, float z, const float4 &) {
        const float dx = pilot_delta(x, px, L[0], ix), dy = pilot_delta(y, py, L[1], iy),
                    dz = pilot_delta(z, pz, L[2], iz);
        if (WEIGHTED) {
            const float m = __ldg(g.mass + i);
            a[0] = __fmaf_rn(m, dx, a[0]); a[1] = __fmaf_rn(m, dy, a[1]); a[2] = __fmaf_rn(m, dz, a[2]);
            a[3] += m;
        } else {
            a[0] += dx; a[1] += dy; a[2] += dz;
        }
        float s, c;
        __sincosf(dx * sx, &s, &c); a[4] += c; a[7] += s;
        __sincosf(dy * sy, &s, &c); a[5] += c; a[8] += s;
        __sincosf(dz * sz, &s, &c); a[6] += c; a[9] += s;
        mn[0] = fminf(mn[0], dx); mx[0] = fmaxf(mx[0], dx);
        mn[1] = fminf(mn[1], dy); mx[1] = fmaxf(mx[1], dy);
        mn[2] = fminf(mn[2], dz); mx[2] = fmaxf(mx[2], dz);
    });
    double tot[10];
    float tmn[3], tmx[3];
    if (frame_reduce<10, 3>(a, mn, mx, partials + (size_t)f * nb * 16, tickets + f, nb, sm, tot, tmn, tmx) && threadIdx.x == 0)
        finish_center<WEIGHTED>(tot, tmn, tmx, px, py, pz, L, g.n, out + f * 3, flags + f);
}

// ---------------------------------------------------------------- calc_rmsd, single pass
template <bool SAME_MASS>
__global__ void __launch_bounds__(kTmaThreads, 2) k_rmsd_tma(FrameView fv, GroupView g, RefView ref, double *partials,
                                                              unsigned int *tickets, float *rmsd_out, float *rot_out, float *com_out,
                                                              int *flags) {
    extern __shared__ unsigned char dyn_smem[];
    __shared__ FrameReduceSmem<kFastSums, 3, kTmaWarps> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *p0 = fv.frame(f) + (size_t)g.first * 3;
    const float px = __ldg(p0), py = __ldg(p0 + 1), pz = __ldg(p0 + 2);
    const float ix = 1.0f / L[0], iy = 1.0f / L[1], iz = 1.0f / L[2];
    float a[kFastSums];
#pragma unroll
    for (int k = 0; k < kFastSums; k++) a[k] = 0.0f;
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    stream_group_tma<true>(fv, g, f, ref.pc, dyn_smem, [&](uint32_t i, float x, float y, float z, const float4 &r) {
        const float d[3] = {pilot_delta(x, px, L[0], ix), pilot_delta(y, py, L[1], iy), pilot_delta(z, pz, L[2], iz)};
        rmsd_accumulate<SAME_MASS>(a, mn, mx, d, r, SAME_MASS ? 0.0f : __ldg(g.mass + i));
    });
    double tot[kFastSums];
    float tmn[3], tmx[3];
    if (frame_reduce<kFastSums, 3>(a, mn, mx, partials + (size_t)f * nb * (kFastSums + 6), tickets + f, nb, sm, tot, tmn, tmx) &&
        threadIdx.x == 0)
        finish_rmsd<SAME_MASS>(tot, tmn, tmx, px, py, pz, L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, flags + f);
}

} // namespace groan

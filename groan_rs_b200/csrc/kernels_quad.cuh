// kernels_quad.cuh -- the single-pass centre / RMSD kernels, third generation: one QUAD of four consecutive atoms per
// thread and iteration, 128-bit shared-memory reads, mixed-component f32x2 arithmetic, one trigonometric sum per axis.
//
// What limits kernels_tma.cuh (profiles/r1_summary.md): it is ISSUE bound, 72 warp instructions per atom at 54 % issue
// utilisation, while the frame bytes already arrive by TMA.  Per atom it spends 7 LDS.32 (stride-3 coordinates plus the
// block-SoA reference), 6 MUFU + 3 FMUL.RZ for sin AND cos, and ~18 instructions of loop control and epilogue.
// This version removes most of that:
//
//  * Four consecutive atoms are 48 contiguous bytes of the AoS frame: three LDS.128 per quad (thread stride 48 B:
//    conflict-free).  The twelve floats land in six register pairs of MIXED components,
//        A = (x0,y0)  B = (z0,x1)  C = (y1,z1)     A' = (x2,y2)  B' = (z2,x3)  C' = (y3,z3),
//    and all per-coordinate work (pilot displacement, minimum image, angle) is component-agnostic: it runs packed with
//    per-frame constants held in the same three patterns (x,y) (z,x) (y,z).
//  * The outer products pc (x) d run on those pairs as well.  For the atom pair (a, b) = (0,1) or (2,3):
//        acc1[u] += pc_u(a) * A      -> (H_ux, H_uy)      scalar-broadcast operand (SASS: R.F32)
//        acc2[u] += (pc_u(a), pc_u(b)) * B -> (H_uz of a, H_ux of b)
//        acc3[u] += pc_u(b) * C      -> (H_uy, H_uz)
//    three FFMA2 per row u and atom pair, nothing wasted; the accumulators are folded to H at the end.  The prepared
//    reference is stored so that (pc_u(a), pc_u(b)) IS a register pair after one LDS.128: per quad four LDS.128
//    [pcx_a pcx_b pcy_a pcy_b] [pcz_a pcz_b w_a w_b] x 2, laid out in planes so that consecutive lanes read consecutive
//    16-byte units (k_ref_permute).  7 LDS per quad instead of 28.
//  * Only the SINE sum per axis.  The Bai-Breen estimate c0 (iterators.rs:1152-1191) is needed for one thing: the
//    integer m = floor(c~ / L) that places the unwrapped mean in the reference's periodic image,
//    c~ being the circular mean unwrapped next to the group.  With the group certified compact (extent < L/2), c~ lies
//    inside [lo, hi] = [p + min d, p + max d].  If floor(lo / L) == floor(hi / L) no trigonometry is needed at all;
//    otherwise exactly one box boundary b = floor(hi / L) * L lies in (lo, hi], all angles 2 pi (u_i - b) / L lie in an
//    arc shorter than pi around 0, and c~ >= b  <=>  sum_i sin(2 pi u_i / L) >= 0.  |sum sin| / n bounds the angle of the
//    mean from below, so a fixed threshold on it (kSinGuard) replaces the edge-band test; frames below it are flagged
//    and re-done by the reference-order passes like before.
//  * Loop control: stage/phase counters instead of divisions, full chunks only in the hot loop.
//  * Two rings feed the quads: the kernels that also read the reference (k_rmsd_quad, k_cov_quad) run one
//    private cp.async ring per warp (stream_quads_warp: no barrier protocol at all), the centre kernels the CTA-wide TMA ring
//    (stream_quads: bulk copies on mbarriers, last reader refills).
#pragma once
#include "kernels_tma.cuh"

namespace groan {

constexpr int kQuadCenterThreads = 256;            // CTA size of k_center_quad; a chunk is one quad per thread
constexpr int kQuadRmsdThreads = 256;              // CTA size of k_rmsd_quad
constexpr int kQuadRefBlock = 128;                 // atoms per permuted reference block (32 quads = one warp)
constexpr double kSinGuard = 2.0e-4;               // |sum sin| / n below this: side of the boundary not certain
constexpr float kMagic = 12582912.0f;              // 1.5 * 2^23: (x + kMagic) - kMagic = rint(x) for |x| < 2^22

__host__ __device__ inline size_t quad_ref_floats(size_t body) {
    return ((body + kQuadRefBlock - 1) / kQuadRefBlock) * (size_t)(4 * kQuadRefBlock);
}

// block-SoA reference (group order, kernels_rmsd.cuh) -> quad-permuted reference (body order: group index - head).
// Block of 32 quads = 2 KB = 4 planes of 32 x 16 B; plane k of quad q holds
//   k = 0: pcx(0) pcx(1) pcy(0) pcy(1)   k = 1: pcz(0) pcz(1) w(0) w(1)   k = 2, 3: the same for atoms 2, 3 of the quad.
__global__ void __launch_bounds__(kThreads) k_ref_permute(const float *pc, float *pq, uint32_t head, uint32_t body) {
    const uint32_t quads = body / 4;
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        float4 r[4];
#pragma unroll
        for (int k = 0; k < 4; k++) r[k] = ref_at(pc, head + q * 4 + k);
        float4 *o = reinterpret_cast<float4 *>(pq) + (size_t)(q >> 5) * 128 + (q & 31);
        o[0] = make_float4(r[0].x, r[1].x, r[0].y, r[1].y);
        o[32] = make_float4(r[0].z, r[1].z, r[0].w, r[1].w);
        o[64] = make_float4(r[2].x, r[3].x, r[2].y, r[3].y);
        o[96] = make_float4(r[2].z, r[3].z, r[2].w, r[3].w);
    }
}


// a 3-vector sum held as three register pairs in the patterns of the quad: a = (x,y), b = (z,x), c = (y,z)
struct V3 {
    float2 a, b, c;
};
__device__ __forceinline__ V3 v3_zero() {
    V3 v;
    v.a = v.b = v.c = make_float2(0.f, 0.f);
    return v;
}
__device__ __forceinline__ void v3_add(V3 &s, const V3 &d) {
    s.a = __fadd2_rn(s.a, d.a);
    s.b = __fadd2_rn(s.b, d.b);
    s.c = __fadd2_rn(s.c, d.c);
}
// s += P.x * d(atom a) + P.y * d(atom b), d = the three pairs of an atom pair: a = d(a).xy, b = (d(a).z, d(b).x), c = d(b).yz
__device__ __forceinline__ void v3_fma(V3 &s, float2 P, const V3 &d) {
    s.a = __ffma2_rn(splat(P.x), d.a, s.a);
    s.b = __ffma2_rn(P, d.b, s.b);
    s.c = __ffma2_rn(splat(P.y), d.c, s.c);
}
__device__ __forceinline__ V3 v3_mul(float2 P, const V3 &d) {
    V3 r;
    r.a = __fmul2_rn(splat(P.x), d.a);
    r.b = __fmul2_rn(P, d.b);
    r.c = __fmul2_rn(splat(P.y), d.c);
    return r;
}
__device__ __forceinline__ float v3_x(const V3 &s) { return s.a.x + s.b.y; }
__device__ __forceinline__ float v3_y(const V3 &s) { return s.a.y + s.c.x; }
__device__ __forceinline__ float v3_z(const V3 &s) { return s.b.x + s.c.y; }

// per-frame constants in the three patterns
struct QuadConst {
    V3 negp; // -pilot
    V3 inv;  // 1 / L
    float negl[3]; // -L (scalars: the last step of the min-image runs as two scalar FFMAs, see quad_delta)
    float sc[3];   // 2 pi / L
};
__device__ __forceinline__ V3 v3_pattern(float x, float y, float z) {
    V3 v;
    v.a = make_float2(x, y);
    v.b = make_float2(z, x);
    v.c = make_float2(y, z);
    return v;
}

// min-image displacement from the pilot, pattern-wise; same arithmetic as pilot_delta (kernels_center.cuh).
// The last step has three different register operands: as an FFMA2 it would read six registers and hold the FMA pipe
// for three cycles (profiles/exp/ffma2_rate.cu: 3.0 cycles per warp instruction against 2.06 with an immediate operand
// and 1.03 for a scalar FFMA), so it is issued as two scalar FFMAs, and -L needs no register pairs.
__device__ __forceinline__ float2 quad_delta(float2 x, float2 negp, float2 inv, float nl0, float nl1) {
    const float2 d = __fadd2_rn(x, negp);
    const float2 k = __fadd2_rn(__ffma2_rn(d, inv, splat(kMagic)), splat(-kMagic));
    return make_float2(__fmaf_rn(nl0, k.x, d.x), __fmaf_rn(nl1, k.y, d.y));
}
// sin(2 pi x / L) on the SFU for both halves, from the coordinate AS LOADED: the unwrapped coordinate u = p + d differs
// from x by a whole number of box lengths, so sin(2 pi u / L) = sin(2 pi x / L); no phase constant, and the sines do not
// wait for the min-image chain.  (|x| < 65 L is certified by the finishing thread, so the f32 angle is good to 3e-5.)
__device__ __forceinline__ float2 quad_sin(float2 x, float sc0, float sc1) {
    return make_float2(__sinf(__fmul_rn(x.x, sc0)), __sinf(__fmul_rn(x.y, sc1)));
}

struct QuadMinMax {
    float mn[3], mx[3];
};
// the three pairs of an atom pair -> per-axis min / max (FMNMX3)
__device__ __forceinline__ void quad_minmax(QuadMinMax &m, const V3 &d) {
    m.mn[0] = fminf(m.mn[0], fminf(d.a.x, d.b.y)); m.mx[0] = fmaxf(m.mx[0], fmaxf(d.a.x, d.b.y));
    m.mn[1] = fminf(m.mn[1], fminf(d.a.y, d.c.x)); m.mx[1] = fmaxf(m.mx[1], fmaxf(d.a.y, d.c.x));
    m.mn[2] = fminf(m.mn[2], fminf(d.b.x, d.c.y)); m.mx[2] = fmaxf(m.mx[2], fmaxf(d.b.x, d.c.y));
}

template <bool WITH_REF, int STAGES, int NT>
struct QuadCfg {
    static constexpr int kAtoms = NT * 4; // atoms per chunk
    static constexpr size_t kFrameBytes = (size_t)kAtoms * 12;
    static constexpr size_t kRefBytes = WITH_REF ? (size_t)kAtoms * 16 : 0;
    static constexpr size_t kStageBytes = kFrameBytes + kRefBytes;
    static constexpr size_t kConstOff = STAGES * kStageBytes + 256; // ring, barriers, then 24 floats of per-frame constants
    static constexpr size_t kBytes = kConstOff + 128;
};
constexpr int kQuadCenterStages = 4; // 48 KB ring: 4 CTAs per SM
constexpr int kQuadRmsdStages = 4;   // 112 KB ring: 2 CTAs per SM (no static shared memory: the reduction scratch aliases the ring)

// control block of the quad ring: full[s] (count 1 + tx bytes) is armed by whoever issues the copies and completed by
// the TMA; empty[s] (count = warps) collects one arrival per warp that has finished reading stage s.  mbarrier.arrive
// returns the state BEFORE the arrival, so exactly one warp sees "pending == 1": it is the last reader, and it refills
// the stage (no polling, no separate counter to reset: the phase completes and re-arms itself).
template <int STAGES>
struct QuadCtl {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];
};
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(20000u)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_a(bar, parity)) {
        if (++spins > (1u << 24)) __trap(); // a protocol bug fails the launch instead of hanging the GPU
    }
}
// arrive (release) and return the pending count before this arrival
__device__ __forceinline__ uint32_t mbar_arrive_pending(uint32_t bar) {
    uint32_t cnt;
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%1];\n\tmbarrier.pending_count.b64 %0, st;\n\t}"
                 : "=r"(cnt)
                 : "r"(bar)
                 : "memory");
    return cnt;
}

// ---------------------------------------------------------------- per-warp ring (frame + reference), cp.async
// What the CTA-wide TMA ring costs the RMSD kernels (profiles/r2_async.md): the arithmetic alone runs the bench batch in
// 0.27 ms, the barrier protocol around it (two waits, two warp-elected arrivals with their answer looked at, the refill code)
// adds 0.065 ms even when the copies are 16 bytes long, and with real data every warp waits for the slowest of the CTA's
// eight before a stage is refilled.  Here every warp owns its slice of the ring outright: 32 quads = 1536 B of frame +
// one 2 KB block of the permuted reference per stage, filled by its own lanes with 16-byte cp.async (coalesced 512 B per
// warp instruction, the same L2 policies as the bulk copies) and waited for with cp.async.wait_group + __syncwarp: no
// mbarrier, no elected thread, no dependence on any other warp.  Chunk c of the frame still goes to CTA c % gridDim.x and
// warp w still takes quads [32 w, 32 w + 32) of it, so the sums are bit-identical to the TMA ring's.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int STAGES, int NT, typename F>
__device__ __forceinline__ void stream_quads_warp(const FrameView &fv, const GroupView &g, int f, const BodyGeom &bg, const float *ref_pq,
                                                  unsigned char *smem, F &&fn) {
    constexpr uint32_t CH = NT * 4, W = NT / 32;
    constexpr uint32_t kSlotF = 32 * 48, kSlotR = kQuadRefBlock * 16, kSlot = kSlotF + kSlotR;
    static_assert(kQuadRefBlock == 128, "one reference block per warp and stage");
    static_assert((size_t)W * STAGES * kSlot <= QuadCfg<true, STAGES, NT>::kStageBytes * STAGES, "per-warp rings fit the CTA ring");
    uint32_t t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    const uint32_t lane = t & 31, w = t >> 5;
    const uint32_t chunks = (bg.body + CH - 1) / CH;
    const uint32_t n = chunks > blockIdx.x ? (chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (n == 0) return;
    uint32_t ring;
    asm volatile("{\n\t.reg .u64 a;\n\tcvta.to.shared.u64 a, %1;\n\tcvt.u32.u64 %0, a;\n\t}" : "=r"(ring) : "l"(smem));
    ring += w * (uint32_t)(STAGES * kSlot);
    const uint64_t pol_frame = l2_policy_evict_first(), pol_ref = l2_policy_evict_last();
    // this lane's 16-byte units of the warp's slice of chunk blockIdx.x; a chunk further on: + gridDim.x chunks
    const char *gf = reinterpret_cast<const char *>(fv.frame(f) + ((size_t)g.first + bg.head) * 3) +
                     ((size_t)blockIdx.x * CH + w * 128u) * 12 + lane * 16u;
    const char *gr = reinterpret_cast<const char *>(ref_pq) + ((size_t)blockIdx.x * CH + w * 128u) * 16 + lane * 16u;
    const size_t fstep = (size_t)gridDim.x * CH * 12, rstep = (size_t)gridDim.x * CH * 16;
    const uint32_t dst0 = ring + lane * 16u;
    // The destination of a stage is carried as a per-thread address (d, below), not as dst0 + a warp-uniform stage offset:
    // with the offset in a uniform register ptxas 12.9 emits LDGSTS [R + UR + imm], desc[UR] with the offset copied over the
    // cache-policy descriptor, which the B200 rejects as an illegal instruction.
    auto issue_full = [&](uint32_t d) { // the warp's slice of a full chunk into the stage whose lane address is d
        cp_async16(d, gf, pol_frame);
        cp_async16(d + 512u, gf + 512, pol_frame);
        cp_async16(d + 1024u, gf + 1024, pol_frame);
        cp_async16(d + kSlotF, gr, pol_ref);
        cp_async16(d + kSlotF + 512u, gr + 512, pol_ref);
        cp_async16(d + kSlotF + 1024u, gr + 1024, pol_ref);
        cp_async16(d + kSlotF + 1536u, gr + 1536, pol_ref);
        gf += fstep;
        gr += rstep;
    };
    // the CTA's last chunk may be ragged (a multiple of 4 atoms): quads [0, lastq) exist, the reference is padded to blocks
    const uint32_t last_atoms = min(CH, bg.body - (blockIdx.x + (n - 1) * gridDim.x) * CH), lastq = last_atoms >> 2;
    auto issue_last = [&](uint32_t d) {
#pragma unroll
        for (uint32_t k = 0; k < 3; k++)
            if (w * 96u + k * 32u + lane < lastq * 3u) cp_async16(d + k * 512u, gf + k * 512u, pol_frame);
        if (w * 128u < last_atoms) {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) cp_async16(d + kSlotF + k * 512u, gr + k * 512u, pol_ref);
        }
    };
    const uint32_t off_f = ring + lane * 48u, off_r = ring + kSlotF + lane * 16u;
    struct QuadRegs {
        float4 c0, c1, c2, r[4];
    };
    auto load = [&](uint32_t so, QuadRegs &q) {
        q.c0 = lds128(off_f + so); q.c1 = lds128(off_f + so + 16u); q.c2 = lds128(off_f + so + 32u);
        q.r[0] = lds128(off_r + so);
        q.r[1] = lds128(off_r + so + 512u);
        q.r[2] = lds128(off_r + so + 1024u);
        q.r[3] = lds128(off_r + so + 1536u);
    };
    // prologue: the first STAGES chunks (always one group per stage, so that the wait below counts the same everywhere)
#pragma unroll
    for (uint32_t i = 0; i < (uint32_t)STAGES; i++) {
        if (i + 1 < n) issue_full(dst0 + i * kSlot);
        else if (i + 1 == n) issue_last(dst0 + i * kSlot);
        cp_async_commit();
    }
    uint32_t j = blockIdx.x * CH + t * 4, so = 0, i = 0, d = dst0;
    const uint32_t dend = dst0 + (STAGES - 1) * kSlot;
    const uint32_t jstep = gridDim.x * CH;
    // hot loop: this chunk and the one refilled behind it are full
    const uint32_t hot = n > (uint32_t)STAGES + 1 ? n - STAGES - 1 : 0;
    for (; i < hot; i++) {
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        QuadRegs q;
        load(so, q);
        fn(j, q.c0, q.c1, q.c2, q.r);
        __syncwarp(); // every lane has its quad in registers: the stage may be overwritten
        issue_full(d);
        cp_async_commit();
        j += jstep;
        so = so == (STAGES - 1) * kSlot ? 0u : so + kSlot;
        d = d == dend ? dst0 : d + kSlot;
    }
    for (; i < n; i++) {
        cp_async_wait<STAGES - 1>();
        __syncwarp();
        if (i + 1 < n || (t < lastq)) {
            QuadRegs q;
            load(so, q);
            fn(j, q.c0, q.c1, q.c2, q.r);
        }
        __syncwarp();
        if (i + STAGES + 1 < n) issue_full(d);
        else if (i + STAGES + 1 == n) issue_last(d);
        cp_async_commit();
        j += jstep;
        so = so == (STAGES - 1) * kSlot ? 0u : so + kSlot;
        d = d == dend ? dst0 : d + kSlot;
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------- CTA-wide TMA ring (frame bytes only)
// The centre kernels (k_center_quad, k_trig_quad: no reference, 4 CTAs per SM, at the HBM roofline) stream through this ring;
// the kernels that also read the reference use stream_quads_warp above.
// Stream the 16-byte aligned body of the group through the ring; fn(j, c0, c1, c2, r) per quad with j the body index of
// the quad's first atom, c0..c2 the twelve coordinates and r unused (the signature of stream_quads_warp).
// Chunk c of the frame goes to CTA c % gridDim.x.  Only the last chunk of a frame can be ragged (a multiple of 4 atoms),
// so every chunk of a CTA but its last runs the unchecked, stage-unrolled loop.
// Dynamic shared memory: [ring: STAGES x (frame chunk | reference chunk)] [QuadCtl].  After the call the ring is free
// (every copy issued has been consumed) and the caller may reuse it once the CTA has synchronised.
template <bool WITH_REF, int STAGES, int NT, typename F>
__device__ __forceinline__ void stream_quads(const FrameView &fv, const GroupView &g, int f, const BodyGeom &bg, const float *,
                                             unsigned char *smem, F &&fn) {
    static_assert(!WITH_REF, "frame + reference go through stream_quads_warp");
    typedef QuadCfg<WITH_REF, STAGES, NT> C;
    constexpr uint32_t CH = C::kAtoms, kSt = (uint32_t)C::kStageBytes;
    // threadIdx.x through a volatile asm: the value then lives in a register.  Left to itself the allocator re-reads
    // SR_TID.X (and re-derives the shared-memory window base from SR_CgaCtaId) in front of every quad, ~60 cycles of
    // dependent latency before the first LDS of a quad can issue (ncu source view of the round-1 kernel).
    uint32_t t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    const uint32_t lane = t & 31;
    const uint32_t chunks = (bg.body + CH - 1) / CH;
    const uint32_t my_chunks = chunks > blockIdx.x ? (chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const char *src0 = reinterpret_cast<const char *>(fv.frame(f) + ((size_t)g.first + bg.head) * 3);
    QuadCtl<STAGES> &ctl = *reinterpret_cast<QuadCtl<STAGES> *>(smem + STAGES * C::kStageBytes);
    // the ring's shared-memory address, also through a volatile asm (see above)
    uint32_t ring;
    asm volatile("{\n\t.reg .u64 a;\n\tcvta.to.shared.u64 a, %1;\n\tcvt.u32.u64 %0, a;\n\t}" : "=r"(ring) : "l"(smem));
    const uint32_t full0 = ring + STAGES * kSt, empty0 = full0 + STAGES * 8u;
    const uint64_t pol_frame = l2_policy_evict_first();
    auto issue = [&](uint32_t it) { // copy of this CTA's chunk `it` into stage it % STAGES (one thread)
        const uint32_t s = it % STAGES, c = blockIdx.x + it * gridDim.x;
        const uint32_t atoms = min(CH, bg.body - c * CH);
        mbar_expect_tx(ctl.full + s, atoms * 12u);
        bulk_g2s(smem + s * C::kStageBytes, src0 + (size_t)c * CH * 12, atoms * 12u, ctl.full + s, pol_frame);
    };
    if (t == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(ctl.full + s, 1);
            mbar_init(ctl.empty + s, NT / 32);
        }
        fence_mbar_init();
        for (uint32_t it = 0; it < (uint32_t)STAGES && it < my_chunks; it++) issue(it);
    }
    __syncthreads();
    if (my_chunks == 0) return;
    // this thread's quad inside a stage: coordinates at t * 48 B
    const uint32_t off_f = ring + t * 48u;
    uint32_t j = blockIdx.x * CH + t * 4;
    const uint32_t jstep = gridDim.x * CH;
    struct QuadRegs {
        float4 c0, c1, c2, r[4];
    };
    auto load = [&](uint32_t st, QuadRegs &q) { // st = byte offset of the stage
        q.c0 = lds128(off_f + st); q.c1 = lds128(off_f + st + 16u); q.c2 = lds128(off_f + st + 32u);
    };
    auto quad = [&](uint32_t st) {
        QuadRegs q;
        load(st, q);
        fn(j, q.c0, q.c1, q.c2, q.r);
    };
    static_assert(STAGES == 4, "the loop below walks the ring in pairs of stages");
    const uint32_t hot = (my_chunks - 1) & ~1u; // chunks of this CTA that are certainly full, in pairs
    uint32_t ph = 0, it = 0, st = 0;            // st = byte offset of the pair's first stage: 0 or 2 stages
    // Two chunks per trip (stages s, s + 1 with s = 0 or 2): loop control and address arithmetic are paid once per
    // eight atoms of a thread, while each stage keeps its own barriers and is refilled as soon as its last reader leaves.
    // The wait for a stage is issued one quad EARLY -- right after the loads of the quad before it, in front of that
    // quad's ~170 instructions of arithmetic -- so the round trip of the barrier check is covered by arithmetic instead
    // of standing between two quads (the ring is four stages deep: the stage is normally full long before).
    if (hot) mbar_wait_a(full0, 0u);
    for (; it < hot; it += 2) {
        const uint32_t fb = full0 + (st ? 16u : 0u), eb = empty0 + (st ? 16u : 0u);
        const uint32_t nst = st ^ (2u * kSt), nph = ph ^ (uint32_t)(st != 0);
        QuadRegs q;
        load(st, q);
        mbar_wait_a(fb + 8u, ph);
        fn(j, q.c0, q.c1, q.c2, q.r);
        __syncwarp();
        if (lane == 0 && mbar_arrive_pending(eb) == 1u && it + STAGES < my_chunks) issue(it + STAGES);
        j += jstep;
        load(st + kSt, q);
        if (it + 2 < my_chunks) mbar_wait_a(full0 + (nst ? 16u : 0u), nph);
        fn(j, q.c0, q.c1, q.c2, q.r);
        __syncwarp();
        if (lane == 0 && mbar_arrive_pending(eb + 8u) == 1u && it + 1 + STAGES < my_chunks) issue(it + 1 + STAGES);
        j += jstep;
        ph = nph;
        st = nst;
    }
    // the remaining one or two chunks; the very last one may be ragged.  Nothing is left to refill.
    for (; it < my_chunks; it++, j += jstep) {
        const uint32_t s = it % STAGES, c = blockIdx.x + it * gridDim.x;
        const uint32_t quads = min(CH, bg.body - c * CH) >> 2;
        mbar_wait_a(full0 + s * 8u, (it / STAGES) & 1u);
        if (t < quads) quad(s * kSt);
    }
}

// finishing thread of the single-pass centre, sine-only version (see the header): certify and place the mean.
// tot_md = sum m d (or sum d), M = sum m (or n), S = sum sin(2 pi u / L) (unweighted, geometric estimate: iterators.rs:1407)
__device__ inline void finish_center_sin(const double tot_md[3], double M, const double S[3], const float *tmn, const float *tmx,
                                         const float p[3], const float *L, uint32_t n, float *out3, int *flag) {
    int redo = 0;
    const double inv_m = fast_rcp(M);
    for (int k = 0; k < 3; k++) {
        const double Lk = (double)L[k], inv_l = fast_rcp(Lk);
        if (!((double)tmx[k] - (double)tmn[k] < 0.5 * Lk * kExtentSlack)) redo = 1; // not compact: images may differ
        if (!(fabs((double)p[k]) < 64.0 * Lk)) redo = 1;                            // f32 frac(p / L) no longer good enough
        const double lo = (double)p[k] + (double)tmn[k], hi = (double)p[k] + (double)tmx[k];
        const double mlo = floor_div(lo, Lk, inv_l), mhi = floor_div(hi, Lk, inv_l);
        double m = mlo;
        if (mlo != mhi) { // the group straddles the boundary mhi * L: which side is the circular mean on?
            if (!(fabs(S[k]) >= kSinGuard * (double)n)) redo = 1;
            m = S[k] > 0.0 ? mhi : mlo;
        }
        const double um = (double)p[k] + tot_md[k] * inv_m; // mean of the unwrapped group
        out3[k] = (float)(um - m * Lk);                 // c0 = c~ - m L lies in [0, L): the image within L/2 of c0
    }
    *flag = redo;
}

// finishing thread of the fused centre + RMSD kernels: the same placement of the mean, but WITHOUT any trigonometric sum.
// The loop accumulates Sd = sum d and Qd = sum d^2 per axis (unweighted; the image decision is geometric, iterators.rs:1407)
// instead of sum sin -- one FFMA per coordinate instead of FMUL + FMUL.RZ + MUFU.SIN + FADD, a quarter of the loop's issue slots.
// With theta_i = 2 pi (u_i - b) / L the angles of the unwrapped atoms about the straddled boundary b, tbar their mean and
// a_i = theta_i - tbar (sum a_i = 0):
//     S = sum sin theta_i = sin tbar * sum cos a_i + cos tbar * sum sin a_i,
//     sum cos a_i >= n - Q/2,   |sum sin a_i| = |sum (sin a_i - a_i)| <= sum |a_i|^3 / 6 <= A Q / 6,
// Q = sum a_i^2 (from Sd, Qd), A = max |a_i| (from the extent).  If |sin tbar| (n - Q/2) - |cos tbar| A Q / 6 >= kSinGuard n
// the sign of S -- the side of b the circular mean lies on -- is the sign of tbar, certified with the same margin the sine
// sum itself is held to.  Otherwise (*second = 1) the frame goes through the sine-sum pass (k_center_quad, gated), which
// is one more read of that frame; only what IT cannot certify reaches the reference-order passes.
__device__ inline void finish_center_moments(const double tot_md[3], double M, const double Sd[3], const double Qd[3], const float *tmn,
                                             const float *tmx, const float p[3], const float *L, uint32_t n, float *out3, int *flag,
                                             int *second) {
    int redo = 0, again = 0;
    const double inv_m = fast_rcp(M), inv_n = fast_rcp((double)n);
    for (int k = 0; k < 3; k++) {
        const double Lk = (double)L[k], inv_l = fast_rcp(Lk);
        if (!((double)tmx[k] - (double)tmn[k] < 0.5 * Lk * kExtentSlack)) redo = 1; // not compact: images may differ
        const double lo = (double)p[k] + (double)tmn[k], hi = (double)p[k] + (double)tmx[k];
        const double mlo = floor_div(lo, Lk, inv_l), mhi = floor_div(hi, Lk, inv_l);
        double m = mlo;
        if (mlo != mhi) { // the group straddles the boundary mhi * L: which side is the circular mean on?
            const double s = 6.283185307179586 * inv_l, dbar = Sd[k] * inv_n;
            const double tbar = s * (((double)p[k] - mhi * Lk) + dbar); // |tbar| < pi: b lies inside the extent, which is < L/2
            double var = Qd[k] * inv_n - dbar * dbar;                   // f32 partial sums: keep a relative slack
            var = (var > 0.0 ? var : 0.0) * 1.001 + 1e-9;
            const double q = s * s * var;                               // Q / n
            const double a = s * fmax((double)tmx[k] - dbar, dbar - (double)tmn[k]);
            const float st = __sinf((float)tbar), ct = __cosf((float)tbar);
            const double bound = fabs((double)st) * (1.0 - 0.5 * q) - fabs((double)ct) * a * q * (1.0 / 6.0);
            if (!(bound >= kSinGuard)) again = 1;
            m = tbar > 0.0 ? mhi : mlo;
        }
        const double um = (double)p[k] + tot_md[k] * inv_m; // mean of the unwrapped group
        out3[k] = (float)(um - m * Lk);                 // c0 = c~ - m L lies in [0, L): the image within L/2 of c0
    }
    *flag = redo;
    *second = redo ? 0 : again; // a frame that is not compact goes to the reference-order passes right away
}

// Per-frame constants, computed by one thread, parked in shared memory and read back by everybody with volatile loads.
// The detour is deliberate: if the compiler can see that the two x (y, z) of a pattern are the same value it keeps ONE
// copy and rebuilds the register pairs with MOVs in front of every packed instruction (~25 MOV per quad in SASS; a
// mov-through-asm is not enough, ptxas propagates copies through it).  Values loaded from different addresses stay
// in their own registers.  `scratch` = 24 floats of shared memory; contains a __syncthreads().
__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ QuadConst quad_constants(const float p[3], const float L[3], float *scratch) {
    if (threadIdx.x == 0) {
        const float inv[3] = {1.0f / L[0], 1.0f / L[1], 1.0f / L[2]};
        const int pat[6] = {0, 1, 2, 0, 1, 2};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            scratch[k] = -p[pat[k]];
            scratch[6 + k] = inv[pat[k]];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            scratch[12 + k] = -L[k];
            scratch[18 + k] = 6.283185307179586f * inv[k];
        }
    }
    __syncthreads();
    // plain loads from warp-uniform addresses: the compiler may keep the values in uniform registers (FFMA2 / FADD2 take
    // uniform register pairs as operands), which leaves the vector registers to the accumulators and the scheduler
    const float2 *c2 = reinterpret_cast<const float2 *>(scratch);
    QuadConst q;
    q.negp.a = c2[0]; q.negp.b = c2[1]; q.negp.c = c2[2];
    q.inv.a = c2[3]; q.inv.b = c2[4]; q.inv.c = c2[5];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        q.negl[k] = scratch[12 + k];
        q.sc[k] = scratch[18 + k];
    }
    return q;
}

// the six pairs of a quad -> min-image displacements of atom pair (0,1) in d01 and (2,3) in d23
__device__ __forceinline__ void quad_deltas(const QuadConst &q, const float4 &c0, const float4 &c1, const float4 &c2, V3 &d01, V3 &d23) {
    d01.a = quad_delta(make_float2(c0.x, c0.y), q.negp.a, q.inv.a, q.negl[0], q.negl[1]);
    d01.b = quad_delta(make_float2(c0.z, c0.w), q.negp.b, q.inv.b, q.negl[2], q.negl[0]);
    d01.c = quad_delta(make_float2(c1.x, c1.y), q.negp.c, q.inv.c, q.negl[1], q.negl[2]);
    d23.a = quad_delta(make_float2(c1.z, c1.w), q.negp.a, q.inv.a, q.negl[0], q.negl[1]);
    d23.b = quad_delta(make_float2(c2.x, c2.y), q.negp.b, q.inv.b, q.negl[2], q.negl[0]);
    d23.c = quad_delta(make_float2(c2.z, c2.w), q.negp.c, q.inv.c, q.negl[1], q.negl[2]);
}
// xa, xb, xc = the three coordinate pairs of an atom pair as loaded: (x,y) (z,x') (y',z')
__device__ __forceinline__ void quad_sines(const QuadConst &q, float2 xa, float2 xb, float2 xc, V3 &ssin) {
    ssin.a = __fadd2_rn(ssin.a, quad_sin(xa, q.sc[0], q.sc[1]));
    ssin.b = __fadd2_rn(ssin.b, quad_sin(xb, q.sc[2], q.sc[0]));
    ssin.c = __fadd2_rn(ssin.c, quad_sin(xc, q.sc[1], q.sc[2]));
}

// sine of one atom's axis for the finishing thread's head / tail atoms, same definition as quad_sin
__device__ __forceinline__ double edge_sin(float x, float L) {
    const float inv = 1.0f / L;
    return (double)__sinf(__fmul_rn(x, 6.283185307179586f * inv));
}

// ---------------------------------------------------------------- group_get_center / group_get_com
// sums: [0..2] sum m d, [3] sum m, [4..6] sum sin
// sel_mode != 0: only the selected frames are done (1: frames with sel[f] != 0, launched over the whole batch; 2: frames
// sel[blockIdx.y], launched from the device over exactly the frames that need it).
// ext_pilot == nullptr: the single pass -- pilot = the group's first atom, image of the mean decided by the sine sum; with
//   sel_mode != 0 this is the second tier of the fused centre + RMSD kernels, and a frame it cannot certify either gets bit 1
//   of its flag ORed in (the fused kernel's flag word: bit 0 RMSD, bit 1 centre).
// ext_pilot != nullptr: the EXACT pass -- the reference's second loop itself (iterators.rs:1237-1266): pilot = c0, the
//   Bai-Breen estimate of the frame (k_trig_quad), result = c0 + mean(min-image displacement from c0).  No compactness is
//   needed and nothing is flagged: this is what frames the single pass could not certify are re-done with.
// TRIC: the triclinic extension -- every atom goes through Shear::to_u as it leaves shared memory (the
//   sheared picture is an orthogonal periodic box, so everything after that is the orthogonal kernel) and the centre goes back
//   through Shear::to_x; with an orthogonal box both maps are the identity bit for bit.  Single pass only (no ext_pilot, no
//   sel): flagged frames are re-done by the host-launched reference-order passes, which know about the shear themselves.
template <bool WEIGHTED, bool TRIC>
__global__ void __launch_bounds__(kQuadCenterThreads, 3) k_center_quad(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                                 float *out, int *flags, FallbackPlan fp, const int *sel, int sel_mode,
                                                                 const float *ext_pilot) {
    static_assert(kQuadCenterThreads == 256, "maybe_launch_fallback launches this kernel with 256 threads");
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FrameReduceSmem<7, 3, kQuadCenterThreads / 32> &sm = *reinterpret_cast<FrameReduceSmem<7, 3, kQuadCenterThreads / 32> *>(dyn_smem); // reuses the ring once it has drained
    // sel_mode 3: the second tier of the fused kernels, launched from the host behind EVERY fused launch with room for
    // gridDim.y frames: the frames are sel[0 .. n), n = what the fused launch counted (capped; the rest was launched from the
    // device).  Every CTA takes a ticket when it is done; the last one re-arms the counter for the next fused launch.
    unsigned int n_sel = 0;
    if (sel_mode == 3) {
        n_sel = min(*reinterpret_cast<volatile unsigned int *>(fp.second_count), gridDim.y);
        fp.n_report = (int)n_sel;
    }
    auto done = [&]() {
        if (sel_mode != 3 || threadIdx.x != 0) return;
        if (atomicAdd(fp.second_ticket, 1u) == gridDim.x * gridDim.y - 1u) {
            *fp.second_ticket = 0u;
            *fp.second_count = 0u;
        }
    };
    if (sel_mode == 3 && blockIdx.y >= n_sel) {
        done();
        return;
    }
    const int f = sel_mode >= 2 ? sel[blockIdx.y] : (int)blockIdx.y, nb = gridDim.x;
    if (sel_mode == 1 && sel[f] == 0) return; // uniform for the CTA (nobody counts finished frames in this mode)
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = ext_pilot ? ext_pilot + (size_t)f * 3 : fr + (size_t)g.first * 3;
    float p[3] = {__ldg(p0), __ldg(p0 + 1), __ldg(p0 + 2)};
    Shear sh = {0.f, 0.f, 0.f};
    if (TRIC) {
        sh = fv.shear(f);
        sh.to_u(p[0], p[1], p[2]);
    }
    const QuadConst qc = quad_constants(p, L, reinterpret_cast<float *>(dyn_smem + QuadCfg<false, kQuadCenterStages, kQuadCenterThreads>::kConstOff));
    const BodyGeom bg = body_geom(fv, g, f);
    V3 smd = v3_zero(), ssin = v3_zero();
    float2 sm2 = make_float2(0.f, 0.f);
    QuadMinMax mm = {{3.0e38f, 3.0e38f, 3.0e38f}, {-3.0e38f, -3.0e38f, -3.0e38f}};
    stream_quads<false, kQuadCenterStages, kQuadCenterThreads>(fv, g, f, bg, nullptr, dyn_smem,
                                           [&](uint32_t j, const float4 &l0, const float4 &l1, const float4 &l2, const float4 (&)[4]) {
        float4 c0 = l0, c1 = l1, c2 = l2;
        if (TRIC) { // atoms (c0.x c0.y c0.z) (c0.w c1.x c1.y) (c1.z c1.w c2.x) (c2.y c2.z c2.w) into the sheared picture
            sh.to_u(c0.x, c0.y, c0.z);
            sh.to_u(c0.w, c1.x, c1.y);
            sh.to_u(c1.z, c1.w, c2.x);
            sh.to_u(c2.y, c2.z, c2.w);
        }
        V3 d01, d23;
        quad_deltas(qc, c0, c1, c2, d01, d23);
        if (WEIGHTED) {
            const float *mp = g.mass + bg.head + j;
            const float2 m01 = make_float2(__ldg(mp), __ldg(mp + 1)), m23 = make_float2(__ldg(mp + 2), __ldg(mp + 3));
            v3_fma(smd, m01, d01);
            v3_fma(smd, m23, d23);
            sm2 = __fadd2_rn(sm2, __fadd2_rn(m01, m23));
        } else {
            v3_add(smd, d01);
            v3_add(smd, d23);
        }
        quad_sines(qc, make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y), ssin);
        quad_sines(qc, make_float2(c1.z, c1.w), make_float2(c2.x, c2.y), make_float2(c2.z, c2.w), ssin);
        quad_minmax(mm, d01);
        quad_minmax(mm, d23);
    });
    __syncthreads(); // every warp has left the ring: its memory becomes the reduction scratch
    const float a[7] = {v3_x(smd), v3_y(smd), v3_z(smd), sm2.x + sm2.y, v3_x(ssin), v3_y(ssin), v3_z(ssin)};
    double tot[7];
    float tmn[3], tmx[3];
    if (frame_reduce<7, 3>(a, mm.mn, mm.mx, partials + (size_t)f * nb * 13, tickets + f, nb, sm, tot, tmn, tmx) && threadIdx.x == 0) {
        // the up-to-6 atoms outside the 16-byte aligned body, in f64 with the same definitions
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            const double m = WEIGHTED ? (double)__ldg(g.mass + i) : 1.0;
            if (WEIGHTED) tot[3] += m;
            float xq[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
            if (TRIC) sh.to_u(xq[0], xq[1], xq[2]);
            for (int k = 0; k < 3; k++) {
                const float xk = xq[k], d = pilot_delta(xk, p[k], L[k], 1.0f / L[k]);
                tot[k] += m * (double)d;
                tot[4 + k] += edge_sin(xk, L[k]);
                tmn[k] = fminf(tmn[k], d);
                tmx[k] = fmaxf(tmx[k], d);
            }
        }
        if (ext_pilot) {
            const double inv_m = 1.0 / (WEIGHTED ? tot[3] : (double)g.n);
            for (int k = 0; k < 3; k++) out[f * 3 + k] = (float)((double)p[k] + tot[k] * inv_m);
            return;
        }
        int flag = 0;
        finish_center_sin(tot, WEIGHTED ? tot[3] : (double)g.n, tot + 4, tmn, tmx, p, L, g.n, out + f * 3, &flag);
        if (TRIC) sh.to_x(out[f * 3], out[f * 3 + 1], out[f * 3 + 2]); // the centre back from the sheared picture
        if (sel_mode) {
            if (flag) flags[f] |= 2;
        } else {
            flags[f] = flag;
        }
        FallbackPlan mine = fp;  // this pass produces no second-tier frames of its own
        if (sel_mode == 3) mine.second_count = nullptr;
        maybe_launch_fallback(mine, fv, g, RefView(), partials, tickets, flags, flag, 0, f);
    }
    done();
}

// ---------------------------------------------------------------- calc_rmsd (+ optionally the centre)
// CENTER: 0 = RMSD only, 1 = also group_get_center, 2 = also group_get_com.
// canonical sums after the fold: [0..25] as kFastSums (kernels_rmsd.cuh), then (CENTER) [26..28] sum d, [29..31] sum d^2 (unweighted)
constexpr int kQuadSums = kFastSums + 6;

// TRIC (CENTER = 0 only): the triclinic extension, like k_rmsd_fast -- coordinates into the sheared
//   picture (Shear::to_u) as they leave shared memory, minimum image and compactness there, the displacement back to Cartesian
//   (Shear::to_x: a linear map, and sums of Cartesian displacements are what Kabsch needs) before the sums.  Identity bit for
//   bit on an orthogonal box.  Flagged frames are re-done by the host-launched reference-order passes.
template <bool SAME_MASS, int CENTER, bool TRIC = false>
__global__ void __launch_bounds__(kQuadRmsdThreads, 2) k_rmsd_quad(FrameView fv, GroupView g, RefView ref, QuadRef ref_pq, double *partials,
                                                               unsigned int *tickets, float *center_out, float *rmsd_out, float *rot_out,
                                                               float *com_out, int *flags, FallbackPlan fp, const int *sel) {
    constexpr int KS = CENTER ? kQuadSums : kFastSums;
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FrameReduceSmem<KS, 3, kQuadRmsdThreads / 32> &sm = *reinterpret_cast<FrameReduceSmem<KS, 3, kQuadRmsdThreads / 32> *>(dyn_smem); // reuses the ring once it has drained
    // sel: the frames of this launch (one launch per 16-byte phase of the group when the frame size is not a multiple of four
    // atoms, see launch_rmsd_quad); nullptr = the whole batch
    const int f = sel ? sel[blockIdx.y] : (int)blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float *p0 = fr + (size_t)g.first * 3;
    static_assert(!TRIC || CENTER == 0, "the triclinic variant is the RMSD-only kernel");
    float p[3] = {__ldg(p0), __ldg(p0 + 1), __ldg(p0 + 2)};
    Shear sh = {0.f, 0.f, 0.f};
    if (TRIC) {
        sh = fv.shear(f);
        sh.to_u(p[0], p[1], p[2]);
    }
    const QuadConst qc = quad_constants(p, L, reinterpret_cast<float *>(dyn_smem + QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads>::kConstOff));
    const BodyGeom bg = body_geom(fv, g, f);
    V3 h[3], hw[3], swd = v3_zero(), smd = v3_zero();
    // centre (CENTER != 0): sum d and sum d^2 per axis, unweighted, as pattern sums -- the moments finish_center_moments decides with
    V3 sdv = v3_zero(), cqv = v3_zero();
#pragma unroll
    for (int u = 0; u < 3; u++) h[u] = hw[u] = v3_zero();
    float2 sq = make_float2(0.f, 0.f);
    QuadMinMax mm = {{3.0e38f, 3.0e38f, 3.0e38f}, {-3.0e38f, -3.0e38f, -3.0e38f}};
    auto atom_pair = [&](const V3 &d, const float4 &r0, const float4 &r1, uint32_t i) {
        const float2 pc[3] = {make_float2(r0.x, r0.y), make_float2(r0.z, r0.w), make_float2(r1.x, r1.y)};
        const float2 w = make_float2(r1.z, r1.w);
        const V3 wd = v3_mul(w, d); // w d also serves Hw = sum pc (w d)^T: no separate w pc products
#pragma unroll
        for (int u = 0; u < 3; u++) {
            v3_fma(h[u], pc[u], d);
            v3_fma(hw[u], pc[u], wd);
        }
        v3_add(swd, wd);
        sq = __ffma2_rn(wd.a, d.a, sq);
        sq = __ffma2_rn(wd.b, d.b, sq);
        sq = __ffma2_rn(wd.c, d.c, sq);
        if (CENTER) v3_add(sdv, d);
        if (!TRIC) quad_minmax(mm, d);
        if (!SAME_MASS) {
            const float2 m = make_float2(__ldg(g.mass + i), __ldg(g.mass + i + 1));
            v3_fma(smd, m, d); // sum m is a constant of the group: ref.sum_m_target, no accumulator
        }
    };
    auto moments = [&](const V3 &d) {
        cqv.a = __ffma2_rn(d.a, d.a, cqv.a);
        cqv.b = __ffma2_rn(d.b, d.b, cqv.b);
        cqv.c = __ffma2_rn(d.c, d.c, cqv.c);
    };
    stream_quads_warp<kQuadRmsdStages, kQuadRmsdThreads>(fv, g, f, bg, ref_pq.v[bg.head], dyn_smem,
                                        [&](uint32_t j, const float4 &l0, const float4 &l1, const float4 &l2, const float4 (&r)[4]) {
        float4 c0 = l0, c1 = l1, c2 = l2;
        if (TRIC) { // atoms (c0.x c0.y c0.z) (c0.w c1.x c1.y) (c1.z c1.w c2.x) (c2.y c2.z c2.w) into the sheared picture
            sh.to_u(c0.x, c0.y, c0.z);
            sh.to_u(c0.w, c1.x, c1.y);
            sh.to_u(c1.z, c1.w, c2.x);
            sh.to_u(c2.y, c2.z, c2.w);
        }
        V3 d01, d23;
        quad_deltas(qc, c0, c1, c2, d01, d23);
        if (TRIC) { // the extent is certified in the sheared picture, the sums are formed with Cartesian displacements
            quad_minmax(mm, d01);
            quad_minmax(mm, d23);
            sh.to_x(d01.a.x, d01.a.y, d01.b.x);
            sh.to_x(d01.b.y, d01.c.x, d01.c.y);
            sh.to_x(d23.a.x, d23.a.y, d23.b.x);
            sh.to_x(d23.b.y, d23.c.x, d23.c.y);
        }
        atom_pair(d01, r[0], r[1], bg.head + j);
        if (CENTER) moments(d01);
        atom_pair(d23, r[2], r[3], bg.head + j + 2);
        if (CENTER) moments(d23);
    });
    __syncthreads(); // every warp has left the ring: its memory becomes the reduction scratch
    float a[KS];
#pragma unroll
    for (int u = 0; u < 3; u++) {
        a[u * 3 + 0] = v3_x(h[u]); a[u * 3 + 1] = v3_y(h[u]); a[u * 3 + 2] = v3_z(h[u]);
        a[9 + u * 3 + 0] = v3_x(hw[u]); a[9 + u * 3 + 1] = v3_y(hw[u]); a[9 + u * 3 + 2] = v3_z(hw[u]);
    }
    a[18] = v3_x(swd); a[19] = v3_y(swd); a[20] = v3_z(swd);
    a[21] = sq.x + sq.y;
    a[22] = v3_x(smd); a[23] = v3_y(smd); a[24] = v3_z(smd);
    a[25] = 0.0f;
    if (CENTER) {
        a[KS - 6] = v3_x(sdv); a[KS - 5] = v3_y(sdv); a[KS - 4] = v3_z(sdv);
        a[KS - 3] = v3_x(cqv); a[KS - 2] = v3_y(cqv); a[KS - 1] = v3_z(cqv);
    }
    double tot[KS];
    float tmn[3], tmx[3];
    if (frame_reduce<KS, 3>(a, mm.mn, mm.mx, partials + (size_t)f * nb * (KS + 6), tickets + f, nb, sm, tot, tmn, tmx) &&
        threadIdx.x == 0) {
        // the up-to-6 atoms outside the 16-byte aligned body, in f64 with the same definitions
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            const float4 r = ref_at(ref.pc, i);
            const double pcd[3] = {(double)r.x, (double)r.y, (double)r.z}, w = (double)r.w;
            double d[3];
            float xq[3] = {__ldg(q), __ldg(q + 1), __ldg(q + 2)}, dq[3];
            if (TRIC) sh.to_u(xq[0], xq[1], xq[2]);
            for (int k = 0; k < 3; k++) {
                dq[k] = pilot_delta(xq[k], p[k], L[k], 1.0f / L[k]);
                tmn[k] = fminf(tmn[k], dq[k]);
                tmx[k] = fmaxf(tmx[k], dq[k]);
            }
            if (TRIC) sh.to_x(dq[0], dq[1], dq[2]);
            for (int k = 0; k < 3; k++) {
                d[k] = (double)dq[k];
                if (CENTER) {
                    tot[KS - 6 + k] += d[k];
                    tot[KS - 3 + k] += d[k] * d[k];
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
            }
        }
        tot[25] = ref.sum_m_target;
        double rt[kFastSums];
        for (int k = 0; k < kFastSums; k++) rt[k] = tot[k];
        int flag_r = 0, flag_c = 0, second = 0;
        finish_rmsd<SAME_MASS>(rt, tmn, tmx, p[0], p[1], p[2], L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, &flag_r);
        if (CENTER) {
            // centre: geometric (sum d / n) or mass-weighted with the target group's masses (= the COM of the RMSD)
            double md[3];
            for (int k = 0; k < 3; k++) md[k] = CENTER == 2 ? (SAME_MASS ? tot[18 + k] : tot[22 + k]) : tot[KS - 6 + k];
            const double M = CENTER == 2 ? (SAME_MASS ? ref.sum_w : tot[25]) : (double)g.n;
            finish_center_moments(md, M, tot + (KS - 6), tot + (KS - 3), tmn, tmx, p, L, g.n, center_out + f * 3, &flag_c, &second);
            fp.second_flags[f] = second;
        }
        flags[f] = flag_r | (flag_c << 1);
        maybe_launch_fallback(fp, fv, g, ref, partials, tickets, flags, flag_r | flag_c, second, f);
    }
}

// ---------------------------------------------------------------- the exact passes of a contiguous group
// What a frame costs when the single pass cannot certify it (a group that spans the box: any membrane) used to be three
// reference-order passes at a tenth of the memory bandwidth (k_trig, k_unwrap, k_cov: precise sincosf, fmodf chains, f64
// sums behind 32-bit scalar loads).  These are the same passes on the ring + quad machinery: the Bai-Breen estimate with the
// SFU (k_trig_quad), the unwrap around it (k_center_quad with ext_pilot), the covariance in f64 (k_cov_quad).
//
// k_trig_quad -- estimate_center (iterators.rs:1152-1191, auxiliary.rs:59-99), geometric: c0 = (atan2(-sum sin th, -sum cos th)
// + pi) / s, th = 2 pi wrap(x) / L.  sin and cos are periodic, so th is taken from frac(x / L) in [-1/2, 1/2] -- where the
// SFU is at its best (|err| < 5e-7) whatever the coordinate -- instead of wrapping first.  The estimate only picks periodic
// images in the next pass; its error moves an atom to another image only if the atom lies within ~1e-7 nm of the point
// opposite to c0, where the reference's own f32 sums decide by rounding as well.
// sums: [0..2] sum cos, [3..5] sum sin
__global__ void __launch_bounds__(kQuadCenterThreads, 3) k_trig_quad(FrameView fv, GroupView g, double *partials, unsigned int *tickets,
                                                               float *c0_out, const int *sel, int sel_mode) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FrameReduceSmem<6, 0, kQuadCenterThreads / 32> &sm = *reinterpret_cast<FrameReduceSmem<6, 0, kQuadCenterThreads / 32> *>(dyn_smem);
    const int f = sel_mode == 2 ? sel[blockIdx.y] : (int)blockIdx.y, nb = gridDim.x;
    if (sel_mode == 1 && sel[f] == 0) return;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float zero[3] = {0.f, 0.f, 0.f};
    const QuadConst qc = quad_constants(zero, L, reinterpret_cast<float *>(dyn_smem + QuadCfg<false, kQuadCenterStages, kQuadCenterThreads>::kConstOff));
    const BodyGeom bg = body_geom(fv, g, f);
    V3 sc = v3_zero(), ss = v3_zero();
    auto turn = [](float2 x, float2 inv) { // frac(x / L) in [-1/2, 1/2], both halves
        const float2 t = __fmul2_rn(x, inv), k = __fadd2_rn(__fadd2_rn(t, splat(kMagic)), splat(-kMagic));
        return __fadd2_rn(t, make_float2(-k.x, -k.y));
    };
    auto pair = [&](float2 xa, float2 xb, float2 xc) {
        const float2 ta = turn(xa, qc.inv.a), tb = turn(xb, qc.inv.b), tc = turn(xc, qc.inv.c);
        const float tp = 6.283185307179586f;
        sc.a = __fadd2_rn(sc.a, make_float2(__cosf(ta.x * tp), __cosf(ta.y * tp)));
        ss.a = __fadd2_rn(ss.a, make_float2(__sinf(ta.x * tp), __sinf(ta.y * tp)));
        sc.b = __fadd2_rn(sc.b, make_float2(__cosf(tb.x * tp), __cosf(tb.y * tp)));
        ss.b = __fadd2_rn(ss.b, make_float2(__sinf(tb.x * tp), __sinf(tb.y * tp)));
        sc.c = __fadd2_rn(sc.c, make_float2(__cosf(tc.x * tp), __cosf(tc.y * tp)));
        ss.c = __fadd2_rn(ss.c, make_float2(__sinf(tc.x * tp), __sinf(tc.y * tp)));
    };
    stream_quads<false, kQuadCenterStages, kQuadCenterThreads>(fv, g, f, bg, nullptr, dyn_smem,
                                           [&](uint32_t, const float4 &c0, const float4 &c1, const float4 &c2, const float4 (&)[4]) {
        pair(make_float2(c0.x, c0.y), make_float2(c0.z, c0.w), make_float2(c1.x, c1.y));
        pair(make_float2(c1.z, c1.w), make_float2(c2.x, c2.y), make_float2(c2.z, c2.w));
    });
    __syncthreads();
    const float a[6] = {v3_x(sc), v3_y(sc), v3_z(sc), v3_x(ss), v3_y(ss), v3_z(ss)};
    double tot[6];
    if (frame_reduce<6, 0>(a, nullptr, nullptr, partials + (size_t)f * nb * 6, tickets + f, nb, sm, tot, nullptr, nullptr) && threadIdx.x == 0) {
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            for (int k = 0; k < 3; k++) {
                const float u = __ldg(q + k) * (1.0f / L[k]), v = (u - rintf(u)) * 6.283185307179586f;
                tot[k] += (double)__cosf(v);
                tot[3 + k] += (double)__sinf(v);
            }
        }
        for (int k = 0; k < 3; k++)
            c0_out[f * 3 + k] = (float)((atan2(-tot[3 + k], -tot[k]) + 3.14159265358979323846) * (double)L[k] * (1.0 / 6.283185307179586));
    }
}

// k_cov_quad -- kabsch_rmsd's sums for frames the f32 single pass cannot serve (not compact, or RMSD below what f32 products
// resolve).  q_i = wrap(x_i + (bc - com)) - bc (rmsd.rs:479-492) is the min-image displacement of x_i from com, whatever the
// shape of the group: com (the exact pass's group_get_com) is the pilot, and every product pc_u q_v, w pc_u q_v, w q_v q_v is
// formed and summed in f64 (exact products of two f32), like k_cov does -- on the FP64 pipe, which B200 has at half the FP32
// rate and which the rest of the library never touches.
__global__ void __launch_bounds__(kQuadRmsdThreads, 2) k_cov_quad(FrameView fv, GroupView g, RefView ref, QuadRef ref_pq, const float *com_in,
                                                              double *partials, unsigned int *tickets, float *rmsd_out, float *rot_out,
                                                              const int *sel, int sel_mode) {
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    FrameReduceSmem<kCovSums, 0, kQuadRmsdThreads / 32> &sm = *reinterpret_cast<FrameReduceSmem<kCovSums, 0, kQuadRmsdThreads / 32> *>(dyn_smem);
    const int f = sel_mode == 2 ? sel[blockIdx.y] : (int)blockIdx.y, nb = gridDim.x;
    if (sel_mode == 1 && sel[f] == 0) return;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *fr = fv.frame(f);
    const float p[3] = {__ldg(com_in + f * 3), __ldg(com_in + f * 3 + 1), __ldg(com_in + f * 3 + 2)};
    const QuadConst qc = quad_constants(p, L, reinterpret_cast<float *>(dyn_smem + QuadCfg<true, kQuadRmsdStages, kQuadRmsdThreads>::kConstOff));
    const BodyGeom bg = body_geom(fv, g, f);
    double acc[kCovSums];
#pragma unroll
    for (int k = 0; k < kCovSums; k++) acc[k] = 0.0;
    auto atom = [&](float dx, float dy, float dz, float px, float py, float pz, float w) {
        const double q[3] = {(double)dx, (double)dy, (double)dz}, pc[3] = {(double)px, (double)py, (double)pz}, m = (double)w;
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const double mp = m * pc[u];
#pragma unroll
            for (int v = 0; v < 3; v++) {
                acc[u * 3 + v] = fma(pc[u], q[v], acc[u * 3 + v]);
                acc[9 + u * 3 + v] = fma(mp, q[v], acc[9 + u * 3 + v]);
            }
        }
        acc[18] = fma(m, fma(q[0], q[0], fma(q[1], q[1], q[2] * q[2])), acc[18]);
    };
    stream_quads_warp<kQuadRmsdStages, kQuadRmsdThreads>(fv, g, f, bg, ref_pq.v[bg.head], dyn_smem,
                                        [&](uint32_t, const float4 &c0, const float4 &c1, const float4 &c2, const float4 (&r)[4]) {
        V3 d01, d23;
        quad_deltas(qc, c0, c1, c2, d01, d23);
        // unit k = 0: pcx(0) pcx(1) pcy(0) pcy(1); 1: pcz(0) pcz(1) w(0) w(1); 2, 3: the same for atoms 2, 3 (k_ref_permute)
        atom(d01.a.x, d01.a.y, d01.b.x, r[0].x, r[0].z, r[1].x, r[1].z);
        atom(d01.b.y, d01.c.x, d01.c.y, r[0].y, r[0].w, r[1].y, r[1].w);
        atom(d23.a.x, d23.a.y, d23.b.x, r[2].x, r[2].z, r[3].x, r[3].z);
        atom(d23.b.y, d23.c.x, d23.c.y, r[2].y, r[2].w, r[3].y, r[3].w);
    });
    __syncthreads();
    double tot[kCovSums];
    if (frame_reduce<kCovSums, 0>(acc, nullptr, nullptr, partials + (size_t)f * nb * kCovSums, tickets + f, nb, sm, tot, nullptr, nullptr) &&
        threadIdx.x == 0) {
        for (uint32_t t = 0; t < bg.head + bg.tail; t++) {
            const uint32_t i = t < bg.head ? t : bg.head + bg.body + (t - bg.head);
            const float *q = fr + ((size_t)g.first + i) * 3;
            const float4 r = ref_at(ref.pc, i);
            double d[3];
            for (int k = 0; k < 3; k++) d[k] = (double)pilot_delta(__ldg(q + k), p[k], L[k], 1.0f / L[k]);
            const double pcd[3] = {(double)r.x, (double)r.y, (double)r.z}, w = (double)r.w;
            for (int u = 0; u < 3; u++)
                for (int v = 0; v < 3; v++) {
                    tot[u * 3 + v] += pcd[u] * d[v];
                    tot[9 + u * 3 + v] += w * pcd[u] * d[v];
                }
            tot[18] += w * (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        }
        double r[9];
        rmsd_out[f] = (float)finish_kabsch(tot, tot + 9, tot[18], ref, r, nullptr);
        for (int k = 0; k < 9; k++) rot_out[f * 9 + k] = (float)r[k];
    }
}

// ---------------------------------------------------------------- centre + RMSD from ONE gather (index-list groups)
// The quad kernels need a contiguous, 16-byte aligned range.  For index lists (and whatever else falls back to
// k_rmsd_fast) the centre used to be a second gather of the same atoms; this is k_rmsd_fast with the three extra sums of
// the sine-only centre, finished like k_rmsd_quad.  SAME_MASS only (see launch_rmsd_quad).  CENTER: 1 = geometry, 2 = COM.
template <int CENTER>
__global__ void __launch_bounds__(kThreads, 2) k_rmsd_fast_center(FrameView fv, GroupView g, RefView ref, double *partials,
                                                                   unsigned int *tickets, float *center_out, float *rmsd_out,
                                                                   float *rot_out, float *com_out, int *flags) {
    constexpr int KS = kQuadSums;
    __shared__ FrameReduceSmem<KS, 3> sm;
    const int f = blockIdx.y, nb = gridDim.x;
    float L[3];
    fv.lengths(f, L[0], L[1], L[2]);
    const float *p0 = fv.frame(f) + (size_t)g.atom(0) * 3;
    const float p[3] = {__ldg(p0), __ldg(p0 + 1), __ldg(p0 + 2)};
    const float inv[3] = {1.0f / L[0], 1.0f / L[1], 1.0f / L[2]};
    const float sc[3] = {6.283185307179586f * inv[0], 6.283185307179586f * inv[1], 6.283185307179586f * inv[2]};
    float a[kFastSums], c[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kFastSums; k++) a[k] = 0.0f;
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for_each_group_atom(fv, g, f, [&](uint32_t i, float x, float y, float z) {
        const float4 r = ref_at(ref.pc, i);
        const float d[3] = {pilot_delta(x, p[0], L[0], inv[0]), pilot_delta(y, p[1], L[1], inv[1]), pilot_delta(z, p[2], L[2], inv[2])};
        rmsd_accumulate<true>(a, mn, mx, d, r, 0.0f, d);
        if (CENTER == 1) { c[0] += d[0]; c[1] += d[1]; c[2] += d[2]; }
        c[3] += __sinf(__fmul_rn(x, sc[0])); // same definition as quad_sin / edge_sin
        c[4] += __sinf(__fmul_rn(y, sc[1]));
        c[5] += __sinf(__fmul_rn(z, sc[2]));
    });
    float all[KS];
#pragma unroll
    for (int k = 0; k < kFastSums; k++) all[k] = a[k];
#pragma unroll
    for (int k = 0; k < 6; k++) all[kFastSums + k] = c[k];
    double tot[KS];
    float tmn[3], tmx[3];
    if (frame_reduce<KS, 3>(all, mn, mx, partials + (size_t)f * nb * (KS + 6), tickets + f, nb, sm, tot, tmn, tmx) && threadIdx.x == 0) {
        double rt[kFastSums];
        for (int k = 0; k < kFastSums; k++) rt[k] = tot[k];
        int flag_r = 0, flag_c = 0;
        finish_rmsd<true>(rt, tmn, tmx, p[0], p[1], p[2], L, ref, rmsd_out + f, rot_out + f * 9, com_out + f * 3, &flag_r);
        double md[3];
        for (int k = 0; k < 3; k++) md[k] = CENTER == 2 ? tot[18 + k] : tot[KS - 6 + k];
        finish_center_sin(md, CENTER == 2 ? ref.sum_w : (double)g.n, tot + (KS - 3), tmn, tmx, p, L, g.n, center_out + f * 3, &flag_c);
        flags[f] = flag_r | (flag_c << 1);
    }
}

} // namespace groan

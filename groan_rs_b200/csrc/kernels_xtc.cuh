// kernels_xtc.cuh -- xtc coordinate streams decoded on the GPU.
//
// The bit stream of a frame (format: xtc_codec.hpp) is sequential -- where a group starts depends on the run length and
// radix index left behind by the group before it -- but frames are independent, and inside a frame the state only changes
// at groups whose flag bit is set.  One WARP decodes one frame:
//   * while the flag stays clear, every group has the same size in bits, so the 32 lanes look at the flag bits of the next
//     32 groups at once (positions pos + lane * stride); a ballot finds the first set flag;
//   * the lanes in front of it decode their whole group (mixed-radix unpack of the large atom, the run of small atoms
//     relative to it, the water swap) and write the reader's floats, (float)int * (1 / precision) (xdrfile.c:844,915-917);
//   * the lane at the set flag reads the 5-bit run / radix change, decodes its group with the new run and hands the new
//     state to the warp (two shuffles).
// A stream without run-length changes (coordinates in no spatial order) advances 32 atoms per step; a water-rich one
// advances to the next change.  The lines the next steps will touch are prefetched into L1 (the stream is read once, front
// to back).  Batches of few, large frames use W warps per frame (one CTA): the same step over 32 W groups, the first set
// flag found through shared memory.  Divisions by the radices use reciprocals prepared on the host (xtc_recip).
#pragma once
#include "common.cuh"

namespace groan {

struct XtcFrameParams {
    unsigned long long base;   // byte offset of the frame's bit stream in the uploaded byte range (a multiple of 4: XDR)
    uint32_t nbytes;           // its length
    int32_t natoms;
    int32_t minint[3];
    uint32_t sizeint[3];
    int32_t bitsize;           // 0: three separate fields of bitsizeint[] bits
    int32_t bitsizeint[3];
    int32_t smallidx;
    float inv_precision;
    unsigned long long recip1, recip2;  // xtc_recip(sizeint[1]), xtc_recip(sizeint[2])
};

// floor(2^64 / d) + 1: v / d for any 64-bit v is umulhi(v, recip) or one less (xtc_divmod corrects)
__host__ __device__ inline unsigned long long xtc_recip(uint32_t d) { return d <= 1 ? 0ull : (~0ull) / d + 1ull; }
__device__ __forceinline__ unsigned long long xtc_divmod(unsigned long long v, uint32_t d, unsigned long long recip, uint32_t *rem) {
    if (d <= 1) { *rem = 0; return v; }
    unsigned long long q = __umul64hi(v, recip);
    unsigned long long r = v - q * d;
    if ((long long)r < 0) { q--; r += d; }  // the estimate is never more than one too large
    *rem = (uint32_t)r;
    return q;
}
__device__ unsigned long long g_xtc_magic_recip[73];  // xtc_recip(c_xtc_magic[k]), filled once per ctx (groan_xtc.cu)

__constant__ int c_xtc_magic[73] = {0,       0,       0,       0,       0,       0,       0,       0,        0,        8,       10,      12,      16,
                                    20,      25,      32,      40,      50,      64,      80,      101,      128,      161,     203,     256,     322,
                                    406,     512,     645,     812,     1024,    1290,    1625,    2048,     2580,     3250,    4096,    5060,    6501,
                                    8192,    10321,   13003,   16384,   20642,   26007,   32768,   41285,    52015,    65536,   82570,   104031,  131072,
                                    165140,  208063,  262144,  330280,  416127,  524287,  660561,  832255,   1048576,  1321122, 1664510, 2097152, 2642245,
                                    3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};

// n <= 32 bits at bit position `pos` of a big-endian (MSB-first) stream whose base is 4-byte aligned
__device__ __forceinline__ uint32_t xtc_bits(const uint32_t *__restrict__ s, unsigned long long pos, int n) {
    const unsigned long long w = pos >> 5;
    const uint32_t w0 = __byte_perm(__ldg(s + w), 0, 0x0123), w1 = __byte_perm(__ldg(s + w + 1), 0, 0x0123);
    const unsigned long long win = ((unsigned long long)w0 << 32) | w1;
    return (uint32_t)((win << (pos & 31)) >> (64 - n));
}

// three integers in mixed radix sizes[] packed into nbits bits as little-endian 8-bit chunks (xtc_codec.hpp unpack3)
__device__ __forceinline__ void xtc_unpack3(const uint32_t *__restrict__ s, unsigned long long pos, int nbits, uint32_t s1, uint32_t s2,
                                            unsigned long long r1, unsigned long long r2, int32_t out[3]) {
    if (nbits <= 64) {
        unsigned long long v = 0;
        int sh = 0;
        // whole 32-bit pieces first: four chunks each, byte-reversed (the chunks are little-endian, the stream is MSB-first)
        while (nbits - sh > 32) {
            v |= (unsigned long long)__byte_perm(xtc_bits(s, pos + sh, 32), 0, 0x0123) << sh;
            sh += 32;
        }
        while (nbits - sh > 8) {
            v |= (unsigned long long)xtc_bits(s, pos + sh, 8) << sh;
            sh += 8;
        }
        v |= (unsigned long long)xtc_bits(s, pos + sh, nbits - sh) << sh;
        uint32_t rem;
        const unsigned long long q2 = xtc_divmod(v, s2, r2, &rem);
        out[2] = (int32_t)rem;
        const unsigned long long q1 = xtc_divmod(q2, s1, r1, &rem);
        out[1] = (int32_t)rem;
        out[0] = (int32_t)(uint32_t)q1;
    } else {
        unsigned __int128 v = 0;
        int sh = 0;
        while (nbits - sh > 8) {
            v |= (unsigned __int128)xtc_bits(s, pos + sh, 8) << sh;
            sh += 8;
        }
        v |= (unsigned __int128)xtc_bits(s, pos + sh, nbits - sh) << sh;
        const unsigned __int128 q2 = v / s2;
        out[2] = (int32_t)(uint32_t)(v - q2 * s2);
        const unsigned __int128 q1 = q2 / s1;
        out[1] = (int32_t)(uint32_t)(q2 - q1 * s1);
        out[0] = (int32_t)(uint32_t)q1;
    }
}

__device__ __forceinline__ void xtc_emit(float *__restrict__ o, int32_t i, const int32_t v[3], float inv) {
    o[3 * (size_t)i + 0] = __int2float_rn(v[0]) * inv;
    o[3 * (size_t)i + 1] = __int2float_rn(v[1]) * inv;
    o[3 * (size_t)i + 2] = __int2float_rn(v[2]) * inv;
}

// one group at bit position gpos: the large atom, `hdr` bits of flag (+ run field), run / 3 small atoms; first atom index i0
__device__ __forceinline__ void xtc_group(const uint32_t *__restrict__ s, unsigned long long gpos, const XtcFrameParams &p, int big_bits, int hdr,
                                          int run, int smallidx, float *__restrict__ o, int32_t i0) {
    int32_t big[3];
    if (p.bitsize == 0) {
        big[0] = (int32_t)xtc_bits(s, gpos, p.bitsizeint[0]);
        big[1] = (int32_t)xtc_bits(s, gpos + p.bitsizeint[0], p.bitsizeint[1]);
        big[2] = (int32_t)xtc_bits(s, gpos + p.bitsizeint[0] + p.bitsizeint[1], p.bitsizeint[2]);
    } else {
        xtc_unpack3(s, gpos, p.bitsize, p.sizeint[1], p.sizeint[2], p.recip1, p.recip2, big);
    }
    big[0] += p.minint[0]; big[1] += p.minint[1]; big[2] += p.minint[2];
    if (run == 0) {
        xtc_emit(o, i0, big, p.inv_precision);
        return;
    }
    const uint32_t ms = (uint32_t)c_xtc_magic[smallidx];
    const unsigned long long mr = g_xtc_magic_recip[smallidx];
    const int32_t smallnum = (int32_t)(ms >> 1);
    unsigned long long q = gpos + big_bits + hdr;
    int32_t prev[3] = {big[0], big[1], big[2]};
    int32_t i = i0;
    for (int k = 0; k < run; k += 3) {
        int32_t v[3];
        xtc_unpack3(s, q, smallidx, ms, ms, mr, mr, v);
        q += smallidx;
        v[0] += prev[0] - smallnum; v[1] += prev[1] - smallnum; v[2] += prev[2] - smallnum;
        xtc_emit(o, i++, v, p.inv_precision);
        if (k == 0) xtc_emit(o, i++, big, p.inv_precision);  // the water swap: the large atom comes second
        prev[0] = v[0]; prev[1] = v[1]; prev[2] = v[2];
    }
}

constexpr int kXtcWarpsPerCta = 4;   // W = 1: four frames per CTA, one warp each
constexpr int kXtcWideWarps = 16;    // W = 16: one frame per CTA

// W = 1: one warp per frame (kXtcWarpsPerCta frames per CTA).  W > 1: one CTA of W warps per frame.
// xyz: frame f at xyz + f * frame_stride floats.  status[f] != 0: the stream of frame f is damaged (ran past its end).
template <int W>
__global__ void __launch_bounds__(W == 1 ? kXtcWarpsPerCta * 32 : W * 32) k_xtc_decode(const uint32_t *__restrict__ stream,
                                                                                       const XtcFrameParams *__restrict__ params, int n_frames,
                                                                                       float *__restrict__ xyz, size_t frame_stride, int *status) {
    __shared__ int s_first[W > 1 ? W : 1];
    __shared__ int s_state[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = W == 1 ? (int)(blockIdx.x * kXtcWarpsPerCta + warp) : (int)blockIdx.x;
    if (f >= n_frames) return;  // W == 1: whole warps leave; W > 1: the grid has exactly n_frames CTAs
    const int t = W == 1 ? lane : (int)threadIdx.x;  // this thread's group within a step
    constexpr int G = 32 * W;                         // groups examined per step
    const XtcFrameParams p = params[f];
    const uint32_t *s = stream + (p.base >> 2);
    float *o = xyz + (size_t)f * frame_stride;
    const int big_bits = p.bitsize ? p.bitsize : p.bitsizeint[0] + p.bitsizeint[1] + p.bitsizeint[2];
    const unsigned long long end_bits = (unsigned long long)p.nbytes * 8ull;
    int smallidx = p.smallidx, run = 0, bad = 0;
    unsigned long long pos = 0;
    int32_t i = 0;
    while (i < p.natoms) {
        const int per_group = 1 + run / 3;
        const uint32_t stride = (uint32_t)(big_bits + 1 + (run / 3) * smallidx);
        const bool valid = (long long)i + (long long)t * per_group < (long long)p.natoms;
        const unsigned long long gpos = pos + (unsigned long long)t * stride;
        // the stream is read once, front to back: pull the lines of the coming steps into L1 while this one is decoded
        {
            const unsigned long long ahead = (pos >> 3) + (unsigned long long)G * 128ull + (unsigned long long)t * 128ull;
            if (ahead < (unsigned long long)p.nbytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char *>(s) + ahead));
        }
        const bool in_range = gpos + big_bits < end_bits;
        const uint32_t flag = (valid && in_range) ? xtc_bits(s, gpos + big_bits, 1) : 1u;
        const unsigned ball = __ballot_sync(0xffffffffu, flag != 0u);
        int n0 = ball ? __ffs(ball) - 1 : 32;
        if (W > 1) {
            if (lane == 0) s_first[warp] = n0;
            __syncthreads();
            n0 = G;
#pragma unroll
            for (int w = W - 1; w >= 0; w--)
                if (s_first[w] < 32) n0 = w * 32 + s_first[w];
        }
        if (t < n0) xtc_group(s, gpos, p, big_bits, 1, run, smallidx, o, i + t * per_group);
        if (n0 == G) {
            pos += (unsigned long long)G * stride;
            i += G * per_group;
            if (W > 1) __syncthreads();  // s_first is rewritten by the next step
            continue;
        }
        const int32_t i_flag = i + n0 * per_group;
        if (i_flag >= p.natoms) break;  // the first "set flag" was a thread past the last atom: done
        const unsigned long long fpos = pos + (unsigned long long)n0 * stride;
        if (!(fpos + big_bits + 6 <= end_bits + 64)) { bad = 1; break; }
        int new_run = 0, is_smaller = 0;
        if (t == n0) {
            const int r = (int)xtc_bits(s, fpos + big_bits + 1, 5);
            is_smaller = r % 3;
            new_run = r - is_smaller;
            is_smaller--;
            if (W > 1) { s_state[0] = new_run; s_state[1] = is_smaller; }
            xtc_group(s, fpos, p, big_bits, 6, new_run, smallidx, o, i_flag);
        }
        if (W > 1) {
            __syncthreads();
            new_run = s_state[0];
            is_smaller = s_state[1];
            __syncthreads();  // s_state / s_first are rewritten by the next step
        } else {
            new_run = __shfl_sync(0xffffffffu, new_run, n0);
            is_smaller = __shfl_sync(0xffffffffu, is_smaller, n0);
        }
        pos = fpos + big_bits + 6 + (unsigned long long)(new_run / 3) * smallidx;
        i = i_flag + 1 + new_run / 3;
        run = new_run;
        smallidx += is_smaller;
        if (smallidx < 9 || smallidx > 72 || pos > end_bits + 64) { bad = 1; break; }
    }
    if ((W == 1 ? lane : (int)threadIdx.x) == 0) status[f] = bad;
}

// GroupXtcReader semantics on the device (molly_xtc.rs:441-462): a batch that holds only the atoms of `atoms` (ascending)
// is scattered into the full frames; every other atom keeps whatever the slot held before.
__global__ void __launch_bounds__(kThreads) k_scatter_group_frames(const float *__restrict__ compact, const uint32_t *__restrict__ atoms, uint32_t n_sel,
                                                                   float *__restrict__ xyz, size_t n_atoms) {
    const int f = blockIdx.y;
    const float *src = compact + (size_t)f * n_sel * 3;
    float *dst = xyz + (size_t)f * n_atoms * 3;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_sel; k += gridDim.x * blockDim.x) {
        const size_t a = __ldg(atoms + k);
        dst[3 * a + 0] = __ldg(src + 3 * (size_t)k + 0);
        dst[3 * a + 1] = __ldg(src + 3 * (size_t)k + 1);
        dst[3 * a + 2] = __ldg(src + 3 * (size_t)k + 2);
    }
}

}  // namespace groan

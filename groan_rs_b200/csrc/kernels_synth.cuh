// kernels_synth.cuh -- deterministic synthetic workloads, generated on the device.
//
// Bit-identical to oracle/groan_oracle.c (orc_synth_uniform / orc_synth_blob_ref / orc_synth_blob_frame):
// counter-based splitmix64 keyed by (seed, frame, atom, axis), integer -> f32 conversions that are
// exact, and f32 arithmetic without FMA contraction (the library is built with -fmad=false).
// BASELINE.json configs 4 and 5 cannot be stored (4.8 TB of frames), so they are generated per batch.
#pragma once
#include "common.cuh"

namespace groan {

constexpr uint64_t kRefFrame = 0xFFFFFFFFULL;

__device__ __forceinline__ uint64_t atom_hash(uint64_t key, uint64_t atom, uint64_t axis) {
    return splitmix64(key + 4 * atom + axis);
}
// Irwin-Hall(4) of the four 16-bit fields of one hash: integer in [-131070, 131070], exact in f32
__device__ __forceinline__ float ih4(uint64_t h) {
    const int k = (int)(h & 0xFFFF) + (int)((h >> 16) & 0xFFFF) + (int)((h >> 32) & 0xFFFF) + (int)((h >> 48) & 0xFFFF);
    return (float)(k - 131070);
}

// x = lo + u * span, u = (hash >> 40) * 2^-24
__global__ void __launch_bounds__(kThreads) k_synth_uniform(float *xyz, size_t n_atoms, uint64_t seed, uint64_t frame0,
                                                             float lox, float loy, float loz, float spx, float spy, float spz) {
    const int f = blockIdx.y;
    const uint64_t key = frame_key(seed, frame0 + (uint64_t)f);
    float *fr = xyz + (size_t)f * n_atoms * 3;
    const size_t total = n_atoms * 3;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / 3;
        const int k = (int)(e - i * 3);
        const float u = (float)(atom_hash(key, i, (uint64_t)k) >> 40) * 0x1p-24f;
        const float lo = k == 0 ? lox : (k == 1 ? loy : loz);
        const float sp = k == 0 ? spx : (k == 1 ? spy : spz);
        const float t = u * sp;
        fr[e] = lo + t;
    }
}

// reference structure of the blob: p_i * scale + centre
__global__ void __launch_bounds__(kThreads) k_synth_blob_ref(float *xyz, size_t n_atoms, uint64_t seed, float scale, float cx,
                                                              float cy, float cz) {
    const uint64_t key = frame_key(seed, kRefFrame);
    const size_t total = n_atoms * 3;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t i = e / 3;
        const int k = (int)(e - i * 3);
        const float p = ih4(atom_hash(key, i, (uint64_t)k)) * scale;
        xyz[e] = p + (k == 0 ? cx : (k == 1 ? cy : cz));
    }
}

// frame f: x = R_f * p_i + c_f + noise_{f,i}, optionally folded once into [0, L]
__global__ void __launch_bounds__(kThreads) k_synth_blob(float *xyz, size_t n_atoms, uint64_t seed, uint64_t frame0, float scale,
                                                          float nscale, const float *rot, const float *centre, const float *box,
                                                          int wrap) {
    const int f = blockIdx.y;
    const uint64_t kref = frame_key(seed, kRefFrame);
    const uint64_t key = frame_key(seed, frame0 + (uint64_t)f);
    float r[9], c[3], L[3];
#pragma unroll
    for (int k = 0; k < 9; k++) r[k] = rot[f * 9 + k];
#pragma unroll
    for (int k = 0; k < 3; k++) { c[k] = centre[f * 3 + k]; L[k] = box[f * 9 + 4 * k]; }
    float *fr = xyz + (size_t)f * n_atoms * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_atoms; i += (size_t)gridDim.x * blockDim.x) {
        float p[3];
#pragma unroll
        for (int k = 0; k < 3; k++) p[k] = ih4(atom_hash(kref, i, (uint64_t)k)) * scale;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float t0 = r[k * 3 + 0] * p[0], t1 = r[k * 3 + 1] * p[1], t2 = r[k * 3 + 2] * p[2];
            float s = (t0 + t1) + t2;
            s = s + c[k];
            const float nz = ih4(atom_hash(key, i, (uint64_t)k)) * nscale;
            s = s + nz;
            if (wrap) {
                if (s < 0.0f) s = s + L[k];
                else if (s > L[k]) s = s - L[k];
            }
            fr[i * 3 + k] = s;
        }
    }
}

// Frames as the xtc decoder's integer stage holds them: coordinate = (float)(q + origin) * inv_precision, the expression of
// external/xdrfile/xdrfile.c:915-917 (int -> float conversion, then one f32 multiply), so the floats are the reader's, bit
// for bit.  T = int16_t (relative to a per-frame, per-axis origin) or int32_t.
template <typename T>
__global__ void __launch_bounds__(kThreads) k_dequantize(const T *q, const int32_t *origin, float inv_precision, float *out, size_t n_atoms) {
    const int f = blockIdx.y;
    const size_t n3 = n_atoms * 3;
    const T *src = q + (size_t)f * n3;
    float *dst = out + (size_t)f * n3;
    int32_t o[3] = {0, 0, 0};
    if (origin) { o[0] = origin[f * 3]; o[1] = origin[f * 3 + 1]; o[2] = origin[f * 3 + 2]; }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % 3);
        dst[i] = (float)((int32_t)src[i] + (k == 0 ? o[0] : k == 1 ? o[1] : o[2])) * inv_precision;
    }
}

// The other direction, for writing a (fitted / wrapped / centred) batch as xtc: the integer lattice point xdrfile's encoder
// derives from every coordinate before it compresses (external/xdrfile/xdrfile.c:1018-1031): lf = x * precision +- 0.5 in
// f32 (sign of x), truncated to int.  No FMA contraction (-fmad=false), so the integers are the encoder's.
__global__ void __launch_bounds__(kThreads) k_quantize(const float *xyz, float precision, int32_t *out, size_t count) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const float x = xyz[i];
        const float lf = x >= 0.0f ? x * precision + 0.5f : x * precision - 0.5f;
        out[i] = (int32_t)lf;
    }
}

} // namespace groan

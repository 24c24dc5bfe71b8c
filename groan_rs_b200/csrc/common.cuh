// common.cuh -- device helpers shared by every kernel of libgroan_gpu (sm_100a).
//
// The library is compiled with -fmad=false: the reference (rustc) never contracts a*b+c into an
// FMA, and the parity-critical per-atom arithmetic below must round exactly like it.  Where an
// FMA is wanted for speed it is written explicitly (__fmaf_rn / fma()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace groan {

constexpr int kThreads = 256;           // CTA size of the streaming kernels
constexpr int kWarps = kThreads / 32;
constexpr int kSMs = 148;               // B200
constexpr int kMaxBlocksPerFrame = 592; // 4 resident CTAs x 148 SMs

// ---------------------------------------------------------------- reference scalar primitives
// Vector3D::wrap_coordinate, vector3d.rs:401-417 (strict comparisons; x == L stays L, 0 stays 0)
__device__ __forceinline__ float wrap_coordinate(float x, float L) {
    while (x > L) x -= L;
    while (x < 0.0f) x += L;
    return x;
}
__device__ __forceinline__ float wrap_coordinate_count(float x, float L, int &k) {
    k = 0;
    while (x > L) { x -= L; k--; }
    while (x < 0.0f) { x += L; k++; }
    return x;
}
// Vector3D::min_image, vector3d.rs:575-592
__device__ __forceinline__ float min_image(float d, float L) {
    const float h = L / 2.0f;
    while (d > h) d -= L;
    while (d < -h) d += L;
    return d;
}
// C fmodf(a, L) for L > 0: exact by construction for |a| < 2L (a - L is exact there, Sterbenz),
// CUDA's fmodf (0 ulp) otherwise.
__device__ __forceinline__ float fmod_exact(float a, float L) {
    const float aa = fabsf(a);
    if (aa < L) return a;
    if (aa < L + L) return copysignf(aa - L, a);
    return fmodf(a, L);
}
// floor_mod, vector3d.rs:28-30: (x % y + y) % y
__device__ __forceinline__ float floor_mod(float x, float y) { return fmod_exact(fmod_exact(x, y) + y, y); }
// one axis of Vector3D::vector_to, vector3d.rs:561-569: floor_mod(p - c + half, L) - half
__device__ __forceinline__ float vector_to_1(float c, float p, float L) {
    const float h = L / 2.0f;
    return floor_mod(p - c + h, L) - h;
}

__device__ __forceinline__ float pi_x2() { return 3.14159265358979323846f * 2.0f; } // auxiliary.rs:15

// ---------------------------------------------------------------- B200 streaming loads
// 256-bit global loads (LDG.E.256, sm_100+) with cache policy: frame coordinates are read once
// (no L1 allocation, evict-first in L2); the RMSD reference is re-read by every frame of the batch
// and should stay in the 126 MB L2 (evict-last).
struct f8 {
    float v[8];
};
__device__ __forceinline__ f8 ld256_stream(const void *p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ f8 ld256_keep(const void *p) {
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st256_stream(void *p, const f8 &r) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
                 "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double shfl_down_d(double v, int off) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_down_sync(0xffffffffu, lo, off);
    hi = __shfl_down_sync(0xffffffffu, hi, off);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_down_d(v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Per-frame reduction of KS sums and NM (min, max) pairs over all CTAs of a frame.
//
//   1. warp shuffles, then one shared-memory stage across the CTA's warps -> one record per CTA,
//      stored as f64 in partials[(frame, cta)][KS + 2*NM];
//   2. "last CTA of the frame finishes the frame": a ticket per frame; the CTA that draws the last
//      ticket re-reads all records with ALL its threads in a fixed order (deterministic result,
//      independent of scheduling) and leaves the totals in tot[] / tmn[] / tmx[] of EVERY thread.
//
// T is the per-thread accumulator type (float for the single-pass kernels, double for the exact ones).
template <int KS, int NM, int NW = kWarps>
struct FrameReduceSmem {
    double stage[(KS + 2 * NM) * NW];
    double tot[KS + 2 * NM];
    int last;
};

__device__ __forceinline__ double fold(double v, double u, int k, int ks, int nm) {
    return (k < ks) ? v + u : (k < ks + nm ? fmin(v, u) : fmax(v, u));
}

template <int KS, int NM, typename T, int NW>
__device__ __forceinline__ bool frame_reduce(const T (&sum)[KS], const float *mn, const float *mx, double *partials_frame,
                                             unsigned int *ticket, int blocks, FrameReduceSmem<KS, NM, NW> &sm, double (&tot)[KS],
                                             float *tmn, float *tmx) {
    constexpr int KT = KS + 2 * NM;
    static_assert(KT <= NW * 32, "record wider than a CTA");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < KS; k++) {
        const T s = warp_sum(sum[k]);
        if (lane == 0) sm.stage[k * NW + w] = (double)s;
    }
#pragma unroll
    for (int k = 0; k < NM; k++) {
        const float a = warp_min(mn[k]), b = warp_max(mx[k]);
        if (lane == 0) {
            sm.stage[(KS + k) * NW + w] = (double)a;
            sm.stage[(KS + NM + k) * NW + w] = (double)b;
        }
    }
    __syncthreads();
    if (threadIdx.x < KT) {
        const int k = threadIdx.x;
        double v = sm.stage[k * NW];
        for (int j = 1; j < NW; j++) v = fold(v, sm.stage[k * NW + j], k, KS, NM);
        partials_frame[(size_t)blockIdx.x * KT + k] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        sm.last = (t == (unsigned)(blocks - 1));
        if (sm.last) *ticket = 0u; // re-arm for the next launch on this frame
    }
    __syncthreads();
    if (!sm.last) return false;
    __threadfence();
    // parallel, fixed-order re-read: thread (k, js) folds records js, js + J, js + 2J, ...  The loads of a batch of four
    // records are issued before any is used (L2 loads, ~700 cycles each: one dependent load per record made this loop
    // 2-3 us of serial tail); the fold order, and with it the result, does not depend on the batching.
    constexpr int J = (NW * 32 / KT) < NW ? (NW * 32 / KT) : NW;
    const int k = threadIdx.x % KT, js = threadIdx.x / KT;
    if (js < J) {
        const double neutral = (k < KS) ? 0.0 : (k < KS + NM ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0xfff0000000000000LL));
        double v = neutral;
        for (int j0 = js; j0 < blocks; j0 += 4 * J) {
            double x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = j0 + u * J;
                x[u] = (j < blocks) ? __ldcg(partials_frame + (size_t)j * KT + k) : neutral;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) v = fold(v, x[u], k, KS, NM);
        }
        sm.stage[k * NW + js] = v;
    }
    __syncthreads();
    if (threadIdx.x < KT) {
        const int kk = threadIdx.x;
        double v = sm.stage[kk * NW];
        for (int j = 1; j < J; j++) v = fold(v, sm.stage[kk * NW + j], kk, KS, NM);
        sm.tot[kk] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < KS; q++) tot[q] = sm.tot[q];
#pragma unroll
    for (int q = 0; q < NM; q++) {
        tmn[q] = (float)sm.tot[KS + q];
        tmx[q] = (float)sm.tot[KS + NM + q];
    }
    return true;
}

// ---------------------------------------------------------------- 3x3 SVD / Kabsch rotation (f64)
// One-sided Jacobi; singular values sorted descending like nalgebra's Matrix3::svd (rmsd.rs:573).
__device__ inline void svd3(const double A[9], double U[9], double S[3], double V[9]) {
    double W[9];
    for (int i = 0; i < 9; i++) { W[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 64; sweep++) {
        int rotated = 0;
        for (int e = 0; e < 3; e++) {
            const int p = (e == 2) ? 1 : 0, q = (e == 0) ? 1 : 2;
            double al = 0, be = 0, ga = 0;
            for (int i = 0; i < 3; i++) {
                al += W[i * 3 + p] * W[i * 3 + p];
                be += W[i * 3 + q] * W[i * 3 + q];
                ga += W[i * 3 + p] * W[i * 3 + q];
            }
            if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
            rotated = 1;
            const double ze = (be - al) / (2.0 * ga);
            const double t = (ze >= 0 ? 1.0 : -1.0) / (fabs(ze) + sqrt(1.0 + ze * ze));
            const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
            for (int i = 0; i < 3; i++) {
                const double wp = W[i * 3 + p], wq = W[i * 3 + q];
                W[i * 3 + p] = c * wp - s * wq;
                W[i * 3 + q] = s * wp + c * wq;
                const double vp = V[i * 3 + p], vq = V[i * 3 + q];
                V[i * 3 + p] = c * vp - s * vq;
                V[i * 3 + q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < 3; j++) S[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++)
            if (S[b] > S[a]) {
                double ts = S[a]; S[a] = S[b]; S[b] = ts;
                for (int i = 0; i < 3; i++) {
                    double tw = W[i * 3 + a]; W[i * 3 + a] = W[i * 3 + b]; W[i * 3 + b] = tw;
                    double tv = V[i * 3 + a]; V[i * 3 + a] = V[i * 3 + b]; V[i * 3 + b] = tv;
                }
            }
    const double tiny = 1e-12 * (S[0] > 0 ? S[0] : 1.0);
    const int rank = (S[0] > tiny) + (S[1] > tiny) + (S[2] > tiny);
    for (int j = 0; j < rank; j++)
        for (int i = 0; i < 3; i++) U[i * 3 + j] = W[i * 3 + j] / S[j];
    if (rank == 0) {
        for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else if (rank == 1) {
        const double u0[3] = {U[0], U[3], U[6]};
        const int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double u1[3];
        for (int i = 0; i < 3; i++) u1[i] = ((i == m) ? 1.0 : 0.0) - u0[m] * u0[i];
        const double n = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
        for (int i = 0; i < 3; i++) { u1[i] /= n; U[i * 3 + 1] = u1[i]; }
        U[2] = u0[1] * u1[2] - u0[2] * u1[1];
        U[5] = u0[2] * u1[0] - u0[0] * u1[2];
        U[8] = u0[0] * u1[1] - u0[1] * u1[0];
    } else if (rank == 2) {
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
}

// f64 reciprocal / square root for the serial per-frame finish.  The library versions are ABI calls of ~100 dependent
// instructions each under -rdc, and the finish runs in ONE thread per frame while the rest of the GPU waits:
// an SFU seed in f32 plus two Newton steps in f64 FMAs gives the same result to ~1e-16 relative.
__device__ __forceinline__ double fast_rcp(double x) {
    const double ax = fabs(x);
    if (!(ax > 1e-30 && ax < 1e30)) return 1.0 / x;
    double y = (double)__frcp_rn((float)x);
    y = fma(fma(-x, y, 1.0), y, y);
    y = fma(fma(-x, y, 1.0), y, y);
    return y;
}
__device__ __forceinline__ double fast_sqrt(double x) {
    if (!(x > 1e-30 && x < 1e30)) return sqrt(x);
    double y = (double)rsqrtf((float)x);
    const double h = 0.5 * x;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    const double s = x * y;
    return fma(fma(-s, s, x), 0.5 * y, s);
}
// floor(x / L) for L > 0 with the quotient formed by a reciprocal: corrected with the exact remainder
__device__ __forceinline__ double floor_div(double x, double L, double invL) {
    double q = floor(x * invL);
    const double r = fma(-q, L, x);
    if (r < 0.0) q -= 1.0;
    else if (r >= L) q += 1.0;
    return q;
}

// r = U * diag(1, 1, sign det(U Vt)) * Vt, rmsd.rs:573-583 (row-major).
//
// For det(H) > 0 (every non-degenerate, non-reflected case) that matrix is the orthogonal polar factor of H, which a
// scaled Newton iteration X <- (g X + X^-T / g) / 2 reaches in ~8 steps -- about 10x fewer dependent operations than
// the one-sided Jacobi SVD, and this code runs in ONE thread at the very end of a frame while the rest of the GPU
// waits.  The iteration is self-correcting, so it runs in f32 (4-cycle FMAs, SFU reciprocal / square roots) until it
// stalls at f32 precision, and two unscaled f64 steps X <- (X + X^-T) / 2 (quadratic: 1e-6 -> 1e-12 -> 1e-24) finish it.
// Reflections (det < 0) and near-singular H (planar / collinear groups) keep the SVD path.
__device__ inline void kabsch_rotation(const double H[9], double r[9]) {
    double n2 = 0.0;
    for (int i = 0; i < 9; i++) n2 += H[i] * H[i];
    const double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
    const double n = fast_sqrt(n2);
    if (n2 > 0.0 && det > 1e-7 * n2 * n) {
        float X[9];
        const double inv_n = fast_rcp(n);
        for (int i = 0; i < 9; i++) X[i] = (float)(H[i] * inv_n);
        bool ok = false;
        for (int it = 0; it < 30; it++) {
            // Y = X^-T = cofactor(X) / det(X)
            float C[9];
            C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
            C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
            C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
            const float dx = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
            if (!(dx > 0.0f)) break;
            const float idx = __frcp_rn(dx);
            float nx = 0.0f, ny = 0.0f;
            for (int i = 0; i < 9; i++) { C[i] *= idx; nx += X[i] * X[i]; ny += C[i] * C[i]; }
            const float g = __fsqrt_rn(__fsqrt_rn(ny * __frcp_rn(nx))); // Frobenius-norm scaling
            const float a = 0.5f * g, b = 0.5f * __frcp_rn(g);
            float diff = 0.0f;
            for (int i = 0; i < 9; i++) {
                const float v = a * X[i] + b * C[i];
                diff += (v - X[i]) * (v - X[i]);
                X[i] = v;
            }
            if (diff < 1e-11f) { ok = true; break; }
        }
        if (ok) {
            double Y[9];
            for (int i = 0; i < 9; i++) Y[i] = (double)X[i];
            for (int it = 0; it < 2; it++) {
                double C[9];
                C[0] = Y[4] * Y[8] - Y[5] * Y[7]; C[1] = Y[5] * Y[6] - Y[3] * Y[8]; C[2] = Y[3] * Y[7] - Y[4] * Y[6];
                C[3] = Y[2] * Y[7] - Y[1] * Y[8]; C[4] = Y[0] * Y[8] - Y[2] * Y[6]; C[5] = Y[1] * Y[6] - Y[0] * Y[7];
                C[6] = Y[1] * Y[5] - Y[2] * Y[4]; C[7] = Y[2] * Y[3] - Y[0] * Y[5]; C[8] = Y[0] * Y[4] - Y[1] * Y[3];
                const double hd = 0.5 * fast_rcp(Y[0] * C[0] + Y[1] * C[1] + Y[2] * C[2]);
                for (int i = 0; i < 9; i++) Y[i] = fma(hd, C[i], 0.5 * Y[i]);
            }
            for (int i = 0; i < 9; i++) r[i] = Y[i];
            return;
        }
    }
    double U[9], S[3], V[9], M[9];
    svd3(H, U, S, V);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
    const double dm = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
    const double d = dm < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + d * U[i * 3 + 2] * V[j * 3 + 2];
}

// ---------------------------------------------------------------- atom accessors
// A group is either a contiguous index range (the common case: Group = list of ranges,
// container.rs:13-31) or a general ascending index list.
struct GroupView {
    const uint32_t *idx; // null for a contiguous range
    uint32_t first;      // first atom of a contiguous range
    uint32_t n;          // atoms in the group
    const float *mass;   // group order, nullable
    __device__ __forceinline__ uint32_t atom(uint32_t i) const { return idx ? __ldg(idx + i) : first + i; }
};

// Triclinic extension of the centre / RMSD path (the reference rejects such boxes).  A GROMACS box
// v1 = (a, 0, 0), v2 = (bx, by, 0), v3 = (cx, cy, cz) becomes an ORTHOGONAL periodic box of lengths (a, by, cz) under the shear
//     uz = z,   uy = y - z cy / cz,   ux = x - uy bx / by - z cx / cz
// (v2 -> (0, by, 0), v3 -> (0, 0, cz)); fractional coordinates are u_k / L_k, so "Bai-Breen on fractional coordinates" is
// the orthogonal algorithm on u, and because the map is linear, means and displacements computed in u map back with to_x.
// With bx = cx = cy = 0 both maps are the identity bit for bit (y - z * 0 == y).
struct Shear {
    float a21, a31, a32;
    __device__ __forceinline__ void to_u(float &x, float &y, float &z) const {
        y = y - z * a32;
        x = (x - y * a21) - z * a31;
    }
    __device__ __forceinline__ void to_x(float &x, float &y, float &z) const {
        x = (x + y * a21) + z * a31;
        y = y + z * a32;
    }
};

struct FrameView {
    const float *xyz; // F x N x 3
    const float *box; // F x 9
    size_t n_atoms;
    int tric;         // 1: the kernels see sheared coordinates (for_each_group_atom applies Shear::to_u)
    __device__ __forceinline__ Shear shear(int f) const {
        Shear s;
        s.a21 = __ldg(box + f * 9 + 3) / __ldg(box + f * 9 + 4);
        s.a31 = __ldg(box + f * 9 + 6) / __ldg(box + f * 9 + 8);
        s.a32 = __ldg(box + f * 9 + 7) / __ldg(box + f * 9 + 8);
        return s;
    }
    // the first atom of the group as the kernels see it (pilot of the single-pass kernels)
    __device__ __forceinline__ void load_atom(int f, uint32_t atom, float &x, float &y, float &z) const {
        const float *p = frame(f) + (size_t)atom * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
        if (tric) shear(f).to_u(x, y, z);
    }
    __device__ __forceinline__ const float *frame(int f) const { return xyz + (size_t)f * n_atoms * 3; }
    __device__ __forceinline__ void lengths(int f, float &lx, float &ly, float &lz) const {
        lx = __ldg(box + f * 9 + 0);
        ly = __ldg(box + f * 9 + 4);
        lz = __ldg(box + f * 9 + 8);
    }
};

// Visit every atom of the group in frame f that belongs to this CTA: fn(i, x, y, z) with i the
// position inside the group.  A contiguous group is streamed with 256-bit loads (three LDG.256 =
// eight atoms per thread, issued before any of them is used); the up-to-7 atoms before the first
// 32-byte boundary and after the last full octet, and index-list groups, use scalar loads.
template <typename F>
__device__ __forceinline__ void for_each_group_atom_raw(const FrameView &fv, const GroupView &g, int f, F &&fn);
// ... with the coordinates sheared into the orthogonal picture when the frame view says so (triclinic extension)
template <typename F>
__device__ __forceinline__ void for_each_group_atom(const FrameView &fv, const GroupView &g, int f, F &&fn) {
    if (fv.tric) {
        const Shear sh = fv.shear(f);
        for_each_group_atom_raw(fv, g, f, [&](uint32_t i, float x, float y, float z) {
            sh.to_u(x, y, z);
            fn(i, x, y, z);
        });
    } else {
        for_each_group_atom_raw(fv, g, f, fn);
    }
}
template <typename F>
__device__ __forceinline__ void for_each_group_atom_raw(const FrameView &fv, const GroupView &g, int f, F &&fn) {
    const float *fr = fv.frame(f);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (g.idx) {
        // index list: four atoms per thread and trip, all twelve coordinate loads in flight before the first use
        // (one atom per trip left the gather latency bound: 3 dependent-on-index loads at a time per thread)
        uint32_t i = tid;
        for (; i + 3 * nth < g.n; i += 4 * nth) {
            const uint32_t a0 = __ldg(g.idx + i), a1 = __ldg(g.idx + i + nth), a2 = __ldg(g.idx + i + 2 * nth), a3 = __ldg(g.idx + i + 3 * nth);
            const float *p0 = fr + (size_t)a0 * 3, *p1 = fr + (size_t)a1 * 3, *p2 = fr + (size_t)a2 * 3, *p3 = fr + (size_t)a3 * 3;
            const float x0 = __ldg(p0), y0 = __ldg(p0 + 1), z0 = __ldg(p0 + 2), x1 = __ldg(p1), y1 = __ldg(p1 + 1), z1 = __ldg(p1 + 2);
            const float x2 = __ldg(p2), y2 = __ldg(p2 + 1), z2 = __ldg(p2 + 2), x3 = __ldg(p3), y3 = __ldg(p3 + 1), z3 = __ldg(p3 + 2);
            fn(i, x0, y0, z0);
            fn(i + nth, x1, y1, z1);
            fn(i + 2 * nth, x2, y2, z2);
            fn(i + 3 * nth, x3, y3, z3);
        }
        for (; i < g.n; i += nth) {
            const float *p = fr + (size_t)__ldg(g.idx + i) * 3;
            fn(i, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        }
        return;
    }
    const size_t a0 = (size_t)f * fv.n_atoms + g.first; // global atom index of the group's first atom
    const bool base_ok = ((reinterpret_cast<uintptr_t>(fv.xyz) & 31) == 0);
    uint32_t head = base_ok ? (uint32_t)((8 - (a0 & 7)) & 7) : g.n;
    if (head > g.n) head = g.n;
    const uint32_t noct = (g.n - head) >> 3;
    const char *body = reinterpret_cast<const char *>(fr + ((size_t)g.first + head) * 3);
    for (uint32_t o = tid; o < noct; o += nth) {
        const f8 q0 = ld256_stream(body + (size_t)o * 96), q1 = ld256_stream(body + (size_t)o * 96 + 32),
                 q2 = ld256_stream(body + (size_t)o * 96 + 64);
        const uint32_t i = head + o * 8;
        fn(i + 0, q0.v[0], q0.v[1], q0.v[2]);
        fn(i + 1, q0.v[3], q0.v[4], q0.v[5]);
        fn(i + 2, q0.v[6], q0.v[7], q1.v[0]);
        fn(i + 3, q1.v[1], q1.v[2], q1.v[3]);
        fn(i + 4, q1.v[4], q1.v[5], q1.v[6]);
        fn(i + 5, q1.v[7], q2.v[0], q2.v[1]);
        fn(i + 6, q2.v[2], q2.v[3], q2.v[4]);
        fn(i + 7, q2.v[5], q2.v[6], q2.v[7]);
    }
    // head and tail atoms: at most 14, spread over the first threads of the frame's first CTA
    const uint32_t tail0 = head + noct * 8, ntail = g.n - tail0;
    if (tid < head + ntail) {
        const uint32_t i = tid < head ? tid : tail0 + (tid - head);
        const float *p = fr + ((size_t)g.first + i) * 3;
        fn(i, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    }
}

// In-place update of the contiguous atoms [first, first + n) of frame f: fn(i, x, y, z) with i the position inside the
// range and x, y, z by reference.  The 16-byte aligned body goes in quads -- three 128-bit loads and stores per four
// atoms, the same register picture as the quad kernels -- the up-to-3 atoms before and after it one by one.
template <typename F>
__device__ __forceinline__ void for_each_atom_inplace(float *xyz, size_t n_atoms, int f, uint32_t first, uint32_t n, F &&fn) {
    float *fr = xyz + (size_t)f * n_atoms * 3;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const size_t a0 = (size_t)f * n_atoms + first;
    const bool base_ok = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0);
    uint32_t head = base_ok ? (uint32_t)((4 - (a0 & 3)) & 3) : n;
    if (head > n) head = n;
    const uint32_t quads = (n - head) >> 2;
    float4 *body = reinterpret_cast<float4 *>(fr + ((size_t)first + head) * 3);
    for (uint32_t q = tid; q < quads; q += nth) {
        float4 v0 = body[q * 3], v1 = body[q * 3 + 1], v2 = body[q * 3 + 2];
        const uint32_t i = head + q * 4;
        fn(i, v0.x, v0.y, v0.z);
        fn(i + 1, v0.w, v1.x, v1.y);
        fn(i + 2, v1.z, v1.w, v2.x);
        fn(i + 3, v2.y, v2.z, v2.w);
        body[q * 3] = v0;
        body[q * 3 + 1] = v1;
        body[q * 3 + 2] = v2;
    }
    const uint32_t tail0 = head + quads * 4, ntail = n - tail0;
    if (tid < head + ntail) {
        const uint32_t i = tid < head ? tid : tail0 + (tid - head);
        float *p = fr + ((size_t)first + i) * 3;
        float x = p[0], y = p[1], z = p[2];
        fn(i, x, y, z);
        p[0] = x; p[1] = y; p[2] = z;
    }
}

// counter-based generator shared with oracle/groan_oracle.c (orc_splitmix64 / orc_hash)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t frame_key(uint64_t seed, uint64_t frame) {
    return splitmix64(seed ^ (0xD1B54A32D192ED03ULL * (frame + 1)));
}

} // namespace groan

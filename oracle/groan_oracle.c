/*
 * groan_oracle.c -- CPU restatement of groan_rs's PBC geometry hot path (see groan_oracle.h).
 * TEST INFRASTRUCTURE ONLY -- never linked into the product.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread (oracle/Makefile).
 * Citations are into /root/reference (Ladme/groan_rs v0.11.3).
 */
#include "groan_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define POS(xyz, stride, i) ((xyz) + (size_t)(i) * (stride))

/* auxiliary.rs:15  PI_X2 = consts::PI * 2.0f32 (f32 product) */
static const float PI_F = 3.14159265358979323846f;
static float pi_x2(void) { return PI_F * 2.0f; }

/* ------------------------------------------------------------------ scalar primitives */

/* vector3d.rs:401-417: strict comparisons, x == L stays L, x == 0 stays 0 */
float orc_wrap1(float x, float L) {
    float w = x;
    while (w > L) w -= L;
    while (w < 0.0f) w += L;
    return w;
}

/* same loop, also counting the net number of +L steps (image shift) */
static float wrap1_count(float x, float L, int *k) {
    float w = x;
    int c = 0;
    while (w > L) { w -= L; c--; }
    while (w < 0.0f) { w += L; c++; }
    *k = c;
    return w;
}

/* vector3d.rs:575-592 */
float orc_minimg1(float d, float L) {
    float h = L / 2.0f;
    float n = d;
    while (n > h) n -= L;
    while (n < -h) n += L;
    return n;
}

/* vector3d.rs:28-30: (x % y + y) % y, Rust % on f32 == C fmodf */
float orc_floor_mod(float x, float y) { return fmodf(fmodf(x, y) + y, y); }

/* vector3d.rs:561-569: floor_mod(p - c + half, L) - half per axis */
void orc_vector_to(const float c[3], const float p[3], const float L[3], float out[3]) {
    for (int k = 0; k < 3; k++) {
        float h = L[k] / 2.0f;
        out[k] = orc_floor_mod(p[k] - c[k] + h, L[k]) - h;
    }
}

/* vector3d.rs:458-486: 1-D oriented; 2/3-D = Vector3::new(..).magnitude() = sqrt((x*x + y*y) + z*z) */
float orc_distance(const float a[3], const float b[3], int dim, const float L[3]) {
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    switch (dim) {
    case ORC_DIM_NONE: return 0.0f;
    case ORC_DIM_X: return orc_minimg1(a[0] - b[0], L[0]);
    case ORC_DIM_Y: return orc_minimg1(a[1] - b[1], L[1]);
    case ORC_DIM_Z: return orc_minimg1(a[2] - b[2], L[2]);
    case ORC_DIM_XY: dx = orc_minimg1(a[0] - b[0], L[0]); dy = orc_minimg1(a[1] - b[1], L[1]); break;
    case ORC_DIM_XZ: dx = orc_minimg1(a[0] - b[0], L[0]); dz = orc_minimg1(a[2] - b[2], L[2]); break;
    case ORC_DIM_YZ: dy = orc_minimg1(a[1] - b[1], L[1]); dz = orc_minimg1(a[2] - b[2], L[2]); break;
    default:
        dx = orc_minimg1(a[0] - b[0], L[0]);
        dy = orc_minimg1(a[1] - b[1], L[1]);
        dz = orc_minimg1(a[2] - b[2], L[2]);
        break;
    }
    return sqrtf((dx * dx + dy * dy) + dz * dz);
}

/* io/xdrfile.rs:170-187 (matrix -> SimBox) + simbox.rs:185,230 (orthogonality gate) */
int orc_box_lengths(const float box9[9], float L[3]) {
    if (!box9) return ORC_ENOBOX;
    /* v2x = box[1][0], v3x = box[2][0], v3y = box[2][1]; v1y,v1z,v2z must be zero for Gromacs */
    if (box9[3] != 0.0f || box9[6] != 0.0f || box9[7] != 0.0f) return ORC_ENOTORTHO;
    if (box9[1] != 0.0f || box9[2] != 0.0f || box9[5] != 0.0f) return ORC_ENOTORTHO;
    L[0] = box9[0];
    L[1] = box9[4];
    L[2] = box9[8];
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    return ORC_OK;
}

/* ------------------------------------------------------------------ centres, ref32 */

/* iterators.rs:1152-1191 (geometry, mass = 1.0) and :1314-1357 (mass-weighted);
 * per atom auxiliary.rs:59-83; conversion auxiliary.rs:87-99 */
int orc_estimate_center(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass,
                        const float L[3], float out[3]) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    float s[3] = {pi_x2() / L[0], pi_x2() / L[1], pi_x2() / L[2]};
    float xi[3] = {0, 0, 0}, zeta[3] = {0, 0, 0};
    if (g == 0) {
        out[0] = out[1] = out[2] = NAN; /* iterators.rs:1183-1185 */
        return ORC_OK;
    }
    for (size_t i = 0; i < g; i++) {
        const float *p = POS(xyz, stride, idx[i]);
        float m = mass ? mass[i] : 1.0f;
        for (int k = 0; k < 3; k++) {
            float y = orc_wrap1(p[k], L[k]);
            float th = y * s[k];
            xi[k] += m * cosf(th);
            zeta[k] += m * sinf(th);
        }
    }
    for (int k = 0; k < 3; k++) out[k] = (atan2f(-zeta[k], -xi[k]) + PI_F) / s[k];
    return ORC_OK;
}

/* iterators.rs:1237-1266 */
int orc_get_center(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3], float out[3]) {
    float c0[3];
    int st = orc_estimate_center(xyz, stride, idx, g, NULL, L, c0);
    if (st) return st;
    float tot[3] = {0, 0, 0};
    for (size_t i = 0; i < g; i++) {
        float v[3];
        orc_vector_to(c0, POS(xyz, stride, idx[i]), L, v);
        for (int k = 0; k < 3; k++) {
            float np = c0[k] + v[k];
            tot[k] += np;
        }
    }
    float n = (float)g; /* n_atoms as f32 */
    for (int k = 0; k < 3; k++) out[k] = tot[k] / n;
    return ORC_OK;
}

/* iterators.rs:1404-1438: geometric estimate, then mass-weighted mean of unwrapped positions */
int orc_get_com(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass, const float L[3],
                float out[3]) {
    if (!mass) return ORC_ENOMASS;
    float c0[3];
    int st = orc_estimate_center(xyz, stride, idx, g, NULL, L, c0);
    if (st) return st;
    float tot[3] = {0, 0, 0}, sum = 0.0f;
    for (size_t i = 0; i < g; i++) {
        float v[3];
        orc_vector_to(c0, POS(xyz, stride, idx[i]), L, v);
        for (int k = 0; k < 3; k++) {
            float np = c0[k] + v[k];
            tot[k] += np * mass[i];
        }
        sum += mass[i];
    }
    for (int k = 0; k < 3; k++) out[k] = tot[k] / sum;
    return ORC_OK;
}

/* iterators.rs:886-903 */
int orc_get_center_naive(const float *xyz, size_t stride, const uint32_t *idx, size_t g, float out[3]) {
    float tot[3] = {0, 0, 0};
    for (size_t i = 0; i < g; i++)
        for (int k = 0; k < 3; k++) tot[k] += POS(xyz, stride, idx[i])[k];
    for (int k = 0; k < 3; k++) out[k] = tot[k] / (float)g;
    return ORC_OK;
}

/* analysis.rs:348-360 (group_get_center rejects empty groups, analysis.rs:106-108) */
int orc_group_distance(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2, size_t g2,
                       int dim, const float L[3], float *out) {
    if (g1 == 0 || g2 == 0) return ORC_EEMPTY;
    float c1[3], c2[3];
    int st = orc_get_center(xyz, stride, idx1, g1, L, c1);
    if (st) return st;
    st = orc_get_center(xyz, stride, idx2, g2, L, c2);
    if (st) return st;
    *out = orc_distance(c1, c2, dim, L);
    return ORC_OK;
}

/* analysis.rs:401-427 */
int orc_all_distances(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2, size_t g2,
                      int dim, const float L[3], float *out) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < g1; i++) {
        const float *a = POS(xyz, stride, idx1[i]);
        for (size_t j = 0; j < g2; j++) out[i * g2 + j] = orc_distance(a, POS(xyz, stride, idx2[j]), dim, L);
    }
    return ORC_OK;
}

/* the documented consumer (analysis.rs:390-399): Iterator::min_by keeps the first minimum,
 * Iterator::max_by keeps the last maximum, scanning the matrix in row-major order */
int orc_all_distances_minmax(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2,
                             size_t g2, int dim, const float L[3], float *dmin, uint32_t *imin, float *dmax,
                             uint32_t *imax, float cutoff, uint64_t *count_below) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    if (g1 == 0 || g2 == 0) return ORC_EEMPTY;
    float mn = 0, mx = 0;
    uint32_t mni = 0, mnj = 0, mxi = 0, mxj = 0;
    uint64_t cnt = 0;
    int first = 1;
    for (size_t i = 0; i < g1; i++) {
        const float *a = POS(xyz, stride, idx1[i]);
        for (size_t j = 0; j < g2; j++) {
            float d = orc_distance(a, POS(xyz, stride, idx2[j]), dim, L);
            if (first) { mn = mx = d; first = 0; }
            if (d < mn) { mn = d; mni = (uint32_t)i; mnj = (uint32_t)j; }
            if (d >= mx) { mx = d; mxi = (uint32_t)i; mxj = (uint32_t)j; }
            if (d < cutoff) cnt++;
        }
    }
    if (dmin) *dmin = mn;
    if (imin) { imin[0] = mni; imin[1] = mnj; }
    if (dmax) *dmax = mx;
    if (imax) { imax[0] = mxi; imax[1] = mxj; }
    if (count_below) *count_below = cnt;
    return ORC_OK;
}

/* iterators.rs:1548 -> atom.rs:535 -> vector3d.rs:380-384 */
int orc_wrap(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3], int8_t *shifts) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < g; i++) {
        float *p = POS(xyz, stride, idx[i]);
        for (int k = 0; k < 3; k++) {
            int c;
            p[k] = wrap1_count(p[k], L[k], &c);
            if (shifts) shifts[i * 3 + k] = (int8_t)c;
        }
    }
    return ORC_OK;
}

/* atom.rs:498-511: pos += t per component, then wrap */
int orc_translate(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float t[3], const float L[3],
                  int8_t *shifts) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < g; i++) {
        float *p = POS(xyz, stride, idx[i]);
        for (int k = 0; k < 3; k++) {
            int c;
            float v = p[k] + t[k];
            p[k] = wrap1_count(v, L[k], &c);
            if (shifts) shifts[i * 3 + k] = (int8_t)c;
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ cutoff pair search */
/* What CellGrid::neighbors_iter + a distance filter returns for every atom of group 1 over a grid of group 2
 * (cellgrid.rs:301-420; guess.rs:362-470): the pairs with Vector3D::distance(XYZ) < cutoff -- restated as the plain double
 * loop, which a cell grid with cells >= cutoff must reproduce as a SET (its order is undefined, cellgrid.rs:141-144).
 * pairs / dist may be NULL; at most `capacity` pairs are stored, in row-major order; the count is always complete. */
int orc_pairs_within(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2, size_t g2,
                     float cutoff, const float L[3], uint64_t *count, uint32_t *pairs, float *dist, size_t capacity) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    uint64_t n = 0;
    for (size_t i = 0; i < g1; i++)
        for (size_t j = 0; j < g2; j++) {
            const float d = orc_distance(POS(xyz, stride, idx1[i]), POS(xyz, stride, idx2[j]), 7, L);
            if (d < cutoff) {
                if (pairs && n < capacity) {
                    pairs[n * 2] = (uint32_t)i;
                    pairs[n * 2 + 1] = (uint32_t)j;
                    if (dist) dist[n] = d;
                }
                n++;
            }
        }
    *count = n;
    return ORC_OK;
}

/* ------------------------------------------------------------------ whole molecules / groups, centering */

/* Vector3D::filter (vector3d.rs): keep the components of `dim`, zero the others.
 * dim: 0 None, 1 X, 2 Y, 3 Z, 4 XY, 5 XZ, 6 YZ, 7 XYZ (Dimension, dimension.rs:15-25) */
static void filter_dim(float v[3], int dim) {
    const int keep_x = (dim == 1 || dim == 4 || dim == 5 || dim == 7);
    const int keep_y = (dim == 2 || dim == 4 || dim == 6 || dim == 7);
    const int keep_z = (dim == 3 || dim == 5 || dim == 6 || dim == 7);
    if (!keep_x) v[0] = 0.0f;
    if (!keep_y) v[1] = 0.0f;
    if (!keep_z) v[2] = 0.0f;
}

/* System::make_group_whole, modifying.rs:437-465: c = group_estimate_center; pos = c + vector_to(c, pos) */
int orc_make_group_whole(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3]) {
    float c[3];
    int st = orc_estimate_center(xyz, stride, idx, g, NULL, L, c);
    if (st) return st;
    for (size_t i = 0; i < g; i++) {
        float *p = POS(xyz, stride, idx[i]);
        float v[3];
        orc_vector_to(c, p, L, v);
        for (int k = 0; k < 3; k++) p[k] = c[k] + v[k];
    }
    return ORC_OK;
}

/* System::make_molecules_whole, modifying.rs:338-391.  mol_ref[i] = index of the reference atom of atom i's molecule (the
 * lowest index of the molecule, modifying.rs:258-283); ORC_NO_MOL for atoms of monoatomic molecules (left untouched).
 * The reference atom is wrapped into the box, every other atom goes to ref + vector_to(ref, pos). */
int orc_make_molecules_whole(float *xyz, size_t stride, size_t n, const uint32_t *mol_ref, const float L[3]) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < n; i++) /* references first: they are the lowest indices of their molecules in the reference, too */
        if (mol_ref[i] == (uint32_t)i) {
            float *p = POS(xyz, stride, i);
            for (int k = 0; k < 3; k++) p[k] = orc_wrap1(p[k], L[k]);
        }
    for (size_t i = 0; i < n; i++) {
        if (mol_ref[i] == ORC_NO_MOL || mol_ref[i] == (uint32_t)i) continue;
        const float *r = POS(xyz, stride, mol_ref[i]);
        float *p = POS(xyz, stride, i);
        float v[3];
        orc_vector_to(r, p, L, v);
        for (int k = 0; k < 3; k++) p[k] = r[k] + v[k];
    }
    return ORC_OK;
}

/* System::atoms_center / atoms_center_mass, utility.rs:109-130,168-189: shift = box_centre - estimate(reference group),
 * filtered by the dimension, then atoms_translate (all n atoms: pos += shift; wrap) */
int orc_atoms_center(float *xyz, size_t stride, size_t n, const uint32_t *idx, size_t g, const float *mass, int dim,
                     const float L[3]) {
    float c[3];
    int st = orc_estimate_center(xyz, stride, idx, g, mass, L, c);
    if (st) return st;
    float shift[3];
    for (int k = 0; k < 3; k++) shift[k] = L[k] / 2.0f - c[k]; /* get_box_center, mod.rs:298-308 */
    filter_dim(shift, dim);
    for (size_t i = 0; i < n; i++) {
        float *p = POS(xyz, stride, i);
        for (int k = 0; k < 3; k++) p[k] = orc_wrap1(p[k] + shift[k], L[k]);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ 3x3 SVD (f64 one-sided Jacobi) */
/* nalgebra's Matrix3::svd (rmsd.rs:573) is un-vendored; the optimal rotation is unique for
 * rank >= 2 H, pinned by rmsd.rs:618-780 to 1e-6.  Singular values sorted descending like nalgebra. */
static void svd3(const double A[9], double U[9], double S[3], double V[9]) {
    double W[9];
    memcpy(W, A, sizeof(W));
    for (int i = 0; i < 9; i++) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    static const int PQ[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int sweep = 0; sweep < 64; sweep++) {
        int rotated = 0;
        for (int e = 0; e < 3; e++) {
            int p = PQ[e][0], q = PQ[e][1];
            double al = 0, be = 0, ga = 0;
            for (int i = 0; i < 3; i++) {
                al += W[i * 3 + p] * W[i * 3 + p];
                be += W[i * 3 + q] * W[i * 3 + q];
                ga += W[i * 3 + p] * W[i * 3 + q];
            }
            if (ga == 0.0 || fabs(ga) <= 1e-17 * sqrt(al * be)) continue;
            rotated = 1;
            double ze = (be - al) / (2.0 * ga);
            double t = (ze >= 0 ? 1.0 : -1.0) / (fabs(ze) + sqrt(1.0 + ze * ze));
            double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
            for (int i = 0; i < 3; i++) {
                double wp = W[i * 3 + p], wq = W[i * 3 + q];
                W[i * 3 + p] = c * wp - s * wq;
                W[i * 3 + q] = s * wp + c * wq;
                double vp = V[i * 3 + p], vq = V[i * 3 + q];
                V[i * 3 + p] = c * vp - s * vq;
                V[i * 3 + q] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < 3; j++) S[j] = sqrt(W[j] * W[j] + W[3 + j] * W[3 + j] + W[6 + j] * W[6 + j]);
    /* sort descending (swap columns of W and V) */
    for (int a = 0; a < 2; a++)
        for (int b = a + 1; b < 3; b++)
            if (S[b] > S[a]) {
                double ts = S[a]; S[a] = S[b]; S[b] = ts;
                for (int i = 0; i < 3; i++) {
                    double tw = W[i * 3 + a]; W[i * 3 + a] = W[i * 3 + b]; W[i * 3 + b] = tw;
                    double tv = V[i * 3 + a]; V[i * 3 + a] = V[i * 3 + b]; V[i * 3 + b] = tv;
                }
            }
    double tiny = 1e-12 * (S[0] > 0 ? S[0] : 1.0);
    int rank = (S[0] > tiny) + (S[1] > tiny) + (S[2] > tiny);
    for (int j = 0; j < rank; j++)
        for (int i = 0; i < 3; i++) U[i * 3 + j] = W[i * 3 + j] / S[j];
    if (rank == 0) {
        for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else if (rank == 1) {
        /* any orthonormal completion */
        double u0[3] = {U[0], U[3], U[6]};
        int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[m] = 1.0;
        double d = u0[m];
        double u1[3] = {e[0] - d * u0[0], e[1] - d * u0[1], e[2] - d * u0[2]};
        double n = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
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

static double det3(const double M[9]) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

/* r = U * diag(1,1,sign det(U*Vt)) * Vt   (rmsd.rs:573-583); row-major out */
static void rotation_from_h(const double H[9], double r[9]) {
    double U[9], S[3], V[9], UVt[9];
    svd3(H, U, S, V);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) UVt[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
    double d = det3(UVt) < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1] + d * U[i * 3 + 2] * V[j * 3 + 2];
}

/* ------------------------------------------------------------------ RMSD, ref32 */

/* rmsd.rs:425-446 (+ get_box_center mod.rs:298-308, shift_and_wrap_coordinates rmsd.rs:479-492) */
int orc_rmsd_extract(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass, const float L[3],
                     float *y, float bc[3], float com[3]) {
    if (g == 0) return ORC_EEMPTY;
    for (int k = 0; k < 3; k++) bc[k] = L[k] / 2.0f;
    int st = orc_get_com(xyz, stride, idx, g, mass, L, com);
    if (st) return st;
    float sh[3] = {bc[0] - com[0], bc[1] - com[1], bc[2] - com[2]};
    for (size_t i = 0; i < g; i++) {
        const float *p = POS(xyz, stride, idx[i]);
        for (int k = 0; k < 3; k++) y[i * 3 + k] = orc_wrap1(p[k] + sh[k], L[k]);
    }
    return ORC_OK;
}

/* rmsd.rs:547-603.  p = reference, q = target (call order rmsd.rs:158-165,215-222).
 * H is NOT mass weighted (rmsd.rs:566-570); only the final sum is (rmsd.rs:592-599). */
void orc_kabsch(const float *p, const float *q, const float *w, size_t g, const float cp[3], const float cq[3], float sum_w,
                float r[9], float t[3], float *rmsd) {
    float h[9] = {0};
    for (size_t i = 0; i < g; i++) {
        float pc[3], qc[3];
        for (int k = 0; k < 3; k++) { pc[k] = p[i * 3 + k] - cp[k]; qc[k] = q[i * 3 + k] - cq[k]; }
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) h[a * 3 + b] += pc[a] * qc[b];
    }
    double H[9], R[9];
    for (int i = 0; i < 9; i++) H[i] = (double)h[i];
    rotation_from_h(H, R);
    for (int i = 0; i < 9; i++) r[i] = (float)R[i];
    float acc = 0.0f;
    for (size_t i = 0; i < g; i++) {
        float pc[3], qc[3], pr[3];
        for (int k = 0; k < 3; k++) { pc[k] = p[i * 3 + k] - cp[k]; qc[k] = q[i * 3 + k] - cq[k]; }
        /* r.transpose() * pc : column-by-column gemv -> ((rT_k0*pc0) + rT_k1*pc1) + rT_k2*pc2 */
        for (int k = 0; k < 3; k++) pr[k] = (r[0 * 3 + k] * pc[0] + r[1 * 3 + k] * pc[1]) + r[2 * 3 + k] * pc[2];
        float dx = pr[0] - qc[0], dy = pr[1] - qc[1], dz = pr[2] - qc[2];
        float n2 = (dx * dx + dy * dy) + dz * dz;
        acc += w[i] * n2;
    }
    *rmsd = sqrtf(acc / sum_w);
    for (int k = 0; k < 3; k++) t[k] = cq[k] - cp[k];
}

/* rmsd.rs:141-166 */
int orc_calc_rmsd(const float *ref_xyz, size_t ref_stride, const uint32_t *ref_idx, size_t g_ref, const float ref_L[3],
                  const float *mass, const float *tgt_xyz, size_t tgt_stride, const uint32_t *tgt_idx, size_t g_tgt,
                  const float tgt_L[3], float r[9], float *rmsd) {
    if (g_ref == 0 || g_tgt == 0) return ORC_EEMPTY;
    float *yr = (float *)malloc(sizeof(float) * 3 * g_ref);
    float *yt = (float *)malloc(sizeof(float) * 3 * g_tgt);
    float bcr[3], bct[3], comr[3], comt[3], t[3];
    int st = orc_rmsd_extract(ref_xyz, ref_stride, ref_idx, g_ref, mass, ref_L, yr, bcr, comr);
    if (!st) st = orc_rmsd_extract(tgt_xyz, tgt_stride, tgt_idx, g_tgt, mass, tgt_L, yt, bct, comt);
    if (!st && g_ref != g_tgt) st = ORC_EGROUPSIZE;
    if (!st) {
        float sw = 0.0f;
        for (size_t i = 0; i < g_ref; i++) sw += mass[i]; /* masses.iter().sum::<f32>() rmsd.rs:155 */
        orc_kabsch(yr, yt, mass, g_ref, bcr, bct, sw, r, t, rmsd);
    }
    free(yr);
    free(yt);
    return st;
}

/* rmsd.rs:508-528: for ALL atoms: translate(bc - com_tgt)+wrap; -= bc; rotate (vector3d.rs:359: r * x); += com_ref */
void orc_fit(float *xyz, size_t stride, size_t n, const float r[9], const float com_tgt[3], const float com_ref[3],
             const float L[3]) {
    float bc[3] = {L[0] / 2.0f, L[1] / 2.0f, L[2] / 2.0f};
    float sh[3] = {bc[0] - com_tgt[0], bc[1] - com_tgt[1], bc[2] - com_tgt[2]};
    float nbc[3] = {-bc[0], -bc[1], -bc[2]};
    for (size_t i = 0; i < n; i++) {
        float *p = POS(xyz, stride, i);
        float v[3], o[3];
        for (int k = 0; k < 3; k++) {
            v[k] = orc_wrap1(p[k] + sh[k], L[k]);
            v[k] = v[k] + nbc[k];
        }
        for (int k = 0; k < 3; k++) o[k] = (r[k * 3 + 0] * v[0] + r[k * 3 + 1] * v[1]) + r[k * 3 + 2] * v[2];
        for (int k = 0; k < 3; k++) p[k] = o[k] + com_ref[k];
    }
}

/* ------------------------------------------------------------------ exact64 flavour */

static double wrap1d(double x, double L) {
    while (x > L) x -= L;
    while (x < 0.0) x += L;
    return x;
}
static double floor_mod_d(double x, double y) { return fmod(fmod(x, y) + y, y); }

int orc_estimate_center_x64(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass,
                            const float L[3], double out[3]) {
    if (L[0] == 0.0f || L[1] == 0.0f || L[2] == 0.0f) return ORC_EZEROBOX;
    if (g == 0) { out[0] = out[1] = out[2] = NAN; return ORC_OK; }
    double xi[3] = {0, 0, 0}, ze[3] = {0, 0, 0};
    for (size_t i = 0; i < g; i++) {
        const float *p = POS(xyz, stride, idx[i]);
        double m = mass ? (double)mass[i] : 1.0;
        for (int k = 0; k < 3; k++) {
            double th = wrap1d((double)p[k], (double)L[k]) * (2.0 * M_PI / (double)L[k]);
            xi[k] += m * cos(th);
            ze[k] += m * sin(th);
        }
    }
    for (int k = 0; k < 3; k++) out[k] = (atan2(-ze[k], -xi[k]) + M_PI) / (2.0 * M_PI / (double)L[k]);
    return ORC_OK;
}

int orc_get_center_x64(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass, const float L[3],
                       double out[3]) {
    double c0[3];
    int st = orc_estimate_center_x64(xyz, stride, idx, g, NULL, L, c0);
    if (st) return st;
    double tot[3] = {0, 0, 0}, sum = 0.0;
    for (size_t i = 0; i < g; i++) {
        const float *p = POS(xyz, stride, idx[i]);
        double m = mass ? (double)mass[i] : 1.0;
        for (int k = 0; k < 3; k++) {
            double h = (double)L[k] / 2.0;
            double v = floor_mod_d((double)p[k] - c0[k] + h, (double)L[k]) - h;
            tot[k] += (c0[k] + v) * m;
        }
        sum += m;
    }
    for (int k = 0; k < 3; k++) out[k] = tot[k] / sum;
    return ORC_OK;
}

static int extract_x64(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass, const float L[3],
                       double *y, double bc[3]) {
    double com[3];
    for (int k = 0; k < 3; k++) bc[k] = (double)L[k] / 2.0;
    int st = orc_get_center_x64(xyz, stride, idx, g, mass, L, com);
    if (st) return st;
    for (size_t i = 0; i < g; i++) {
        const float *p = POS(xyz, stride, idx[i]);
        for (int k = 0; k < 3; k++) y[i * 3 + k] = wrap1d((double)p[k] + (bc[k] - com[k]), (double)L[k]);
    }
    return ORC_OK;
}

int orc_calc_rmsd_x64(const float *ref_xyz, size_t ref_stride, const uint32_t *ref_idx, size_t g_ref, const float ref_L[3],
                      const float *mass, const float *tgt_xyz, size_t tgt_stride, const uint32_t *tgt_idx, size_t g_tgt,
                      const float tgt_L[3], double r[9], double *rmsd) {
    if (g_ref == 0 || g_tgt == 0) return ORC_EEMPTY;
    if (!mass) return ORC_ENOMASS;
    if (g_ref != g_tgt) return ORC_EGROUPSIZE;
    size_t g = g_ref;
    double *p = (double *)malloc(sizeof(double) * 3 * g), *q = (double *)malloc(sizeof(double) * 3 * g);
    double cp[3], cq[3];
    int st = extract_x64(ref_xyz, ref_stride, ref_idx, g, mass, ref_L, p, cp);
    if (!st) st = extract_x64(tgt_xyz, tgt_stride, tgt_idx, g, mass, tgt_L, q, cq);
    if (!st) {
        double H[9] = {0}, sw = 0.0;
        for (size_t i = 0; i < g; i++) {
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) H[a * 3 + b] += (p[i * 3 + a] - cp[a]) * (q[i * 3 + b] - cq[b]);
            sw += (double)mass[i];
        }
        rotation_from_h(H, r);
        double acc = 0.0;
        for (size_t i = 0; i < g; i++) {
            double pc[3] = {p[i * 3] - cp[0], p[i * 3 + 1] - cp[1], p[i * 3 + 2] - cp[2]};
            double n2 = 0.0;
            for (int k = 0; k < 3; k++) {
                double pr = r[0 * 3 + k] * pc[0] + r[1 * 3 + k] * pc[1] + r[2 * 3 + k] * pc[2];
                double d = pr - (q[i * 3 + k] - cq[k]);
                n2 += d * d;
            }
            acc += (double)mass[i] * n2;
        }
        *rmsd = sqrt(acc / sw);
    }
    free(p);
    free(q);
    return st;
}

/* ------------------------------------------------------------------ triclinic extension (UNPINNED) */
/* No reference counterpart: groan_rs returns SimBoxError::NotOrthogonal (simbox.rs:230-236).
 * Definition (DESIGN.md "Triclinic extension"): GROMACS lower-triangular box, rows v1=(a,0,0),
 * v2=(b,c,0), v3=(d,e,f).  wrap = z, then y, then x, subtracting whole box VECTORS with the
 * reference's strict loop comparisons.  On an orthogonal box this is bit-identical to orc_wrap. */
void orc_tric_wrap1(float p[3], const float B[9], int sh[3]) {
    int kx = 0, ky = 0, kz = 0;
    while (p[2] > B[8]) { p[0] -= B[6]; p[1] -= B[7]; p[2] -= B[8]; kz--; }
    while (p[2] < 0.0f) { p[0] += B[6]; p[1] += B[7]; p[2] += B[8]; kz++; }
    while (p[1] > B[4]) { p[0] -= B[3]; p[1] -= B[4]; ky--; }
    while (p[1] < 0.0f) { p[0] += B[3]; p[1] += B[4]; ky++; }
    while (p[0] > B[0]) { p[0] -= B[0]; kx--; }
    while (p[0] < 0.0f) { p[0] += B[0]; kx++; }
    sh[0] = kx; sh[1] = ky; sh[2] = kz;
}

int orc_tric_wrap(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float B[9], int8_t *shifts) {
    if (B[0] == 0.0f || B[4] == 0.0f || B[8] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < g; i++) {
        int sh[3];
        orc_tric_wrap1(POS(xyz, stride, idx[i]), B, sh);
        if (shifts) for (int k = 0; k < 3; k++) shifts[i * 3 + k] = (int8_t)sh[k];
    }
    return ORC_OK;
}

static float dim_norm2(const float d[3], int dim) {
    float dx = (dim == ORC_DIM_X || dim == ORC_DIM_XY || dim == ORC_DIM_XZ || dim == ORC_DIM_XYZ) ? d[0] : 0.0f;
    float dy = (dim == ORC_DIM_Y || dim == ORC_DIM_XY || dim == ORC_DIM_YZ || dim == ORC_DIM_XYZ) ? d[1] : 0.0f;
    float dz = (dim == ORC_DIM_Z || dim == ORC_DIM_XZ || dim == ORC_DIM_YZ || dim == ORC_DIM_XYZ) ? d[2] : 0.0f;
    return (dx * dx + dy * dy) + dz * dz;
}

/* min-image: sequential z,y,x reduction by whole box vectors (strict comparisons against half the
 * diagonal element), then a search over the 27 neighbouring images for a STRICTLY smaller norm of
 * the selected components; image (0,0,0) is the incumbent, order kz outer, ky, kx inner. */
float orc_tric_distance(const float a[3], const float b[3], int dim, const float B[9]) {
    if (dim == ORC_DIM_NONE) return 0.0f;
    float d[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    float hz = B[8] / 2.0f, hy = B[4] / 2.0f, hx = B[0] / 2.0f;
    while (d[2] > hz) { d[0] -= B[6]; d[1] -= B[7]; d[2] -= B[8]; }
    while (d[2] < -hz) { d[0] += B[6]; d[1] += B[7]; d[2] += B[8]; }
    while (d[1] > hy) { d[0] -= B[3]; d[1] -= B[4]; }
    while (d[1] < -hy) { d[0] += B[3]; d[1] += B[4]; }
    while (d[0] > hx) { d[0] -= B[0]; }
    while (d[0] < -hx) { d[0] += B[0]; }
    float best[3] = {d[0], d[1], d[2]};
    float bn = dim_norm2(d, dim);
    for (int kz = -1; kz <= 1; kz++)
        for (int ky = -1; ky <= 1; ky++)
            for (int kx = -1; kx <= 1; kx++) {
                if (!kz && !ky && !kx) continue;
                float e[3] = {d[0], d[1], d[2]};
                float fz = (float)kz, fy = (float)ky, fx = (float)kx;
                e[0] += fz * B[6]; e[1] += fz * B[7]; e[2] += fz * B[8];
                e[0] += fy * B[3]; e[1] += fy * B[4];
                e[0] += fx * B[0];
                float n = dim_norm2(e, dim);
                if (n < bn) { bn = n; best[0] = e[0]; best[1] = e[1]; best[2] = e[2]; }
            }
    if (dim == ORC_DIM_X) return best[0];
    if (dim == ORC_DIM_Y) return best[1];
    if (dim == ORC_DIM_Z) return best[2];
    return sqrtf(bn);
}

int orc_tric_all_distances(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2, size_t g2,
                           int dim, const float B[9], float *out) {
    if (B[0] == 0.0f || B[4] == 0.0f || B[8] == 0.0f) return ORC_EZEROBOX;
    for (size_t i = 0; i < g1; i++)
        for (size_t j = 0; j < g2; j++)
            out[i * g2 + j] = orc_tric_distance(POS(xyz, stride, idx1[i]), POS(xyz, stride, idx2[j]), dim, B);
    return ORC_OK;
}

/* f64 brute force over (2*nimg+1)^3 images: the self-pin for the triclinic min-image */
double orc_tric_distance_brute64(const float a[3], const float b[3], int dim, const float B[9], int nimg) {
    double d[3] = {(double)a[0] - b[0], (double)a[1] - b[1], (double)a[2] - b[2]};
    int ux = (dim == ORC_DIM_X || dim == ORC_DIM_XY || dim == ORC_DIM_XZ || dim == ORC_DIM_XYZ);
    int uy = (dim == ORC_DIM_Y || dim == ORC_DIM_XY || dim == ORC_DIM_YZ || dim == ORC_DIM_XYZ);
    int uz = (dim == ORC_DIM_Z || dim == ORC_DIM_XZ || dim == ORC_DIM_YZ || dim == ORC_DIM_XYZ);
    double best = INFINITY;
    for (int kz = -nimg; kz <= nimg; kz++)
        for (int ky = -nimg; ky <= nimg; ky++)
            for (int kx = -nimg; kx <= nimg; kx++) {
                double ex = d[0] + kx * (double)B[0] + ky * (double)B[3] + kz * (double)B[6];
                double ey = d[1] + ky * (double)B[4] + kz * (double)B[7];
                double ez = d[2] + kz * (double)B[8];
                double n = (ux ? ex * ex : 0.0) + (uy ? ey * ey : 0.0) + (uz ? ez * ez : 0.0);
                if (n < best) best = n;
            }
    return sqrt(best);
}

/* ------------------------------------------------------------------ synthetic workloads */

uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint64_t orc_hash(uint64_t seed, uint64_t frame, uint64_t atom, uint64_t axis) {
    uint64_t key = orc_splitmix64(seed ^ (0xD1B54A32D192ED03ULL * (frame + 1)));
    return orc_splitmix64(key + 4 * atom + axis);
}

void orc_synth_uniform(float *xyz, size_t n, uint64_t seed, uint64_t frame, const float lo[3], const float span[3]) {
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            float u = (float)(orc_hash(seed, frame, i, (uint64_t)k) >> 40) * 0x1p-24f;
            float t = u * span[k];
            xyz[i * 3 + k] = lo[k] + t;
        }
}

/* Irwin-Hall(4) of 16-bit fields of one hash -> integer in [-2^17, 2^17), exact in f32 */
static float ih4(uint64_t h) {
    int32_t k = (int32_t)(h & 0xFFFF) + (int32_t)((h >> 16) & 0xFFFF) + (int32_t)((h >> 32) & 0xFFFF) + (int32_t)((h >> 48) & 0xFFFF);
    return (float)(k - 131070);
}

#define ORC_REF_FRAME 0xFFFFFFFFULL

void orc_synth_blob_ref(float *xyz, size_t n, uint64_t seed, float scale, const float centre[3]) {
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < 3; k++) {
            float p = ih4(orc_hash(seed, ORC_REF_FRAME, i, (uint64_t)k)) * scale;
            xyz[i * 3 + k] = p + centre[k];
        }
}

void orc_synth_blob_frame(float *xyz, size_t n, uint64_t seed, uint64_t frame, float scale, float nscale, const float rot[9],
                          const float centre[3], const float L[3], int wrap) {
    for (size_t i = 0; i < n; i++) {
        float p[3];
        for (int k = 0; k < 3; k++) p[k] = ih4(orc_hash(seed, ORC_REF_FRAME, i, (uint64_t)k)) * scale;
        for (int k = 0; k < 3; k++) {
            float t0 = rot[k * 3 + 0] * p[0], t1 = rot[k * 3 + 1] * p[1], t2 = rot[k * 3 + 2] * p[2];
            float s = (t0 + t1) + t2;
            s = s + centre[k];
            float nz = ih4(orc_hash(seed, frame, i, (uint64_t)k)) * nscale;
            s = s + nz;
            if (wrap) {
                if (s < 0.0f) s = s + L[k];
                else if (s > L[k]) s = s - L[k];
            }
            xyz[i * 3 + k] = s;
        }
    }
}

/* ------------------------------------------------------------------ restated CPU trajectory path */

/* 240-byte atom record mimicking groan_rs's Atom (atom.rs:23-71): position is an Option<Vector3D> */
typedef union {
    struct {
        size_t residue_number, atom_number, index;
        char residue_name[24], atom_name[24]; /* String = ptr,len,cap */
        char chain_tag, chain;
        float charge, mass, vdw;
        int has_charge, has_mass, has_vdw;
        char element_name[24], element_symbol[24];
        int has_pos;
        float pos[3];
        int has_vel;
        float vel[3];
        int has_force;
        float force[3];
        void *bonded_ptr;
        size_t bonded_len, bonded_cap;
    };
    char raw[240];
} atom_rec;
_Static_assert(sizeof(atom_rec) == 240, "atom record must be 240 bytes");

typedef struct {
    const float *frames, *boxes;
    size_t F, n;
    size_t F_store; /* frames actually held in `frames` (0 = F): frame f of the trajectory is stored frame f % F_store */
    const uint32_t *idx, *idx2;
    size_t g, g2;
    const float *mass_all;
    const float *ref_xyz;
    const float *ref_L;
    int ops, T, tid, dim;
    float *centers, *rmsd, *dmin, *dmax;
} traj_job;

static void *traj_thread(void *arg) {
    traj_job *J = (traj_job *)arg;
    size_t n = J->n;
    /* self.clone() (parallel.rs:236): each thread owns a deep copy of the System */
    atom_rec *atoms = (atom_rec *)calloc(n, sizeof(atom_rec));
    const size_t stride = sizeof(atom_rec) / sizeof(float);
    for (size_t i = 0; i < n; i++) {
        atoms[i].index = i;
        atoms[i].has_mass = 1;
        atoms[i].mass = J->mass_all ? J->mass_all[i] : 1.0f;
    }
    float *base = (float *)((char *)atoms + offsetof(atom_rec, pos));
    float *gmass = NULL;
    if (J->mass_all && J->idx) {
        gmass = (float *)malloc(sizeof(float) * (J->g ? J->g : 1));
        for (size_t i = 0; i < J->g; i++) gmass[i] = J->mass_all[J->idx[i]];
    }
    uint32_t *ref_idx = NULL;
    if (J->ops & 6) {
        ref_idx = (uint32_t *)malloc(sizeof(uint32_t) * (J->g ? J->g : 1));
        for (size_t i = 0; i < J->g; i++) ref_idx[i] = J->idx[i];
    }
    /* interleaved assignment parallel.rs:425-448 */
    const size_t F_store = J->F_store ? J->F_store : J->F;
    for (size_t ft = (size_t)J->tid; ft < J->F; ft += (size_t)J->T) {
        const size_t f = ft % F_store;
        const float *fr = J->frames + f * n * 3;
        const float *L = J->boxes + f * 3;
        /* FrameData::update_system xdrfile_xtc.rs:88-104: set position, reset velocity and force */
        for (size_t i = 0; i < n; i++) {
            atoms[i].has_pos = 1;
            atoms[i].pos[0] = fr[i * 3];
            atoms[i].pos[1] = fr[i * 3 + 1];
            atoms[i].pos[2] = fr[i * 3 + 2];
            atoms[i].has_vel = 0;
            atoms[i].has_force = 0;
        }
        if (J->ops & 1) orc_get_center(base, stride, J->idx, J->g, L, J->centers + f * 3);
        if (J->ops & 6) {
            float r[9], rm = 0.0f;
            orc_calc_rmsd(J->ref_xyz, 3, ref_idx, J->g, J->ref_L, gmass, base, stride, J->idx, J->g, L, r, &rm);
            J->rmsd[f] = rm;
            if (J->ops & 4) {
                float comt[3], comr[3];
                orc_get_com(base, stride, J->idx, J->g, gmass, L, comt);
                orc_get_com(J->ref_xyz, 3, ref_idx, J->g, gmass, J->ref_L, comr);
                /* fit walks every atom record */
                float bc[3] = {L[0] / 2.0f, L[1] / 2.0f, L[2] / 2.0f};
                (void)bc;
                for (size_t i = 0; i < n; i++) orc_fit(atoms[i].pos, 3, 1, r, comt, comr, L);
            }
        }
        if (J->ops & 8) {
            for (size_t i = 0; i < n; i++)
                for (int k = 0; k < 3; k++) atoms[i].pos[k] = orc_wrap1(atoms[i].pos[k], L[k]);
        }
        if (J->ops & 16) {
            orc_all_distances_minmax(base, stride, J->idx, J->g, J->idx2, J->g2, J->dim, L, J->dmin + f, NULL, J->dmax + f, NULL,
                                     0.0f, NULL);
        }
    }
    free(atoms);
    free(gmass);
    free(ref_idx);
    return NULL;
}

static double run_jobs(traj_job *proto, int T) {
    if (T < 1) T = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    traj_job *jobs = (traj_job *)malloc(sizeof(traj_job) * (size_t)T);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < T; t++) {
        jobs[t] = *proto;
        jobs[t].T = T;
        jobs[t].tid = t;
        pthread_create(&th[t], NULL, traj_thread, &jobs[t]);
    }
    for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double orc_baseline_traj(const float *frames, const float *boxes, size_t F, size_t n, const uint32_t *idx, size_t g,
                         const float *mass_all, const float *ref_xyz, const float ref_L[3], int ops, int T, float *centers,
                         float *rmsd) {
    traj_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.boxes = boxes; J.F = F; J.n = n; J.idx = idx; J.g = g; J.mass_all = mass_all;
    J.ref_xyz = ref_xyz; J.ref_L = ref_L; J.ops = ops & 15; J.centers = centers; J.rmsd = rmsd;
    return run_jobs(&J, T);
}

/* the same path over a trajectory of F_total frames of which only F_store distinct ones are held in memory (frame f is
 * stored frame f % F_store): lets the bench time many steps' worth of frames the way the reference splits a whole
 * trajectory among its threads (parallel.rs:425-448), without holding them all.  Outputs: F_store entries. */
double orc_baseline_traj_cyclic(const float *frames, const float *boxes, size_t F_store, size_t F_total, size_t n, const uint32_t *idx,
                                size_t g, const float *mass_all, const float *ref_xyz, const float ref_L[3], int ops, int T,
                                float *centers, float *rmsd) {
    traj_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.boxes = boxes; J.F = F_total; J.F_store = F_store; J.n = n; J.idx = idx; J.g = g; J.mass_all = mass_all;
    J.ref_xyz = ref_xyz; J.ref_L = ref_L; J.ops = ops & 15; J.centers = centers; J.rmsd = rmsd;
    return run_jobs(&J, T);
}

double orc_baseline_pairs(const float *frames, const float *boxes, size_t F, size_t n, const uint32_t *idx1, size_t g1,
                          const uint32_t *idx2, size_t g2, int dim, int T, float *dmin, float *dmax) {
    traj_job J;
    memset(&J, 0, sizeof(J));
    J.frames = frames; J.boxes = boxes; J.F = F; J.n = n; J.idx = idx1; J.g = g1; J.idx2 = idx2; J.g2 = g2;
    J.dim = dim; J.ops = 16; J.dmin = dmin; J.dmax = dmax;
    return run_jobs(&J, T);
}

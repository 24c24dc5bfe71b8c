/*
 * groan_oracle.h -- CPU restatement of groan_rs's per-frame PBC geometry hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (groan_rs_b200/, libgroan_gpu.so)
 * may include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker or
 * the CPU baseline.
 *
 * The reference (Ladme/groan_rs v0.11.3) is Rust and cannot be compiled in this image
 * (no cargo/rustc), so this is a restatement: every function cites the reference
 * file:line it follows.  Two flavours:
 *   ref32   -- same operation order and f32 rounding as the reference (sequential f32
 *              sums, while-loop wraps, fmodf floor_mod, (x*x+y*y)+z*z then sqrtf).
 *              Compile with -ffp-contract=off so no FMA is contracted (rustc never does).
 *   exact64 -- the same per-atom algorithm evaluated and accumulated in f64.
 * Parity pin: tests/test_oracle_kat.py checks ref32 against the reference's own golden
 * vectors.  The triclinic functions (orc_tric_*) have NO reference
 * counterpart (groan_rs returns SimBoxError::NotOrthogonal) -- their parity is UNPINNED
 * by the reference and self-pinned as described in DESIGN.md.
 *
 * Conventions: coordinates are AoS floats with a caller-given stride in floats
 * (3 for a plain [n][3] array as xdrfile's read_xtc emits; 60 for the 240-byte
 * atom-record layout the CPU baseline uses to mimic Vec<Atom>), groups are ascending
 * atom-index lists (container.rs:51-115,415-436), boxes are either L[3] (orthogonal
 * lengths) or the row-major 3x3 matrix read_xtc emits (io/xdrfile.rs:170-187).
 */
#ifndef GROAN_ORACLE_H
#define GROAN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes -- numerically identical to include/groan_gpu.h */
enum {
    ORC_OK = 0,
    ORC_ENOBOX = 1,     /* SimBoxError::DoesNotExist   errors.rs:556-582 */
    ORC_ENOTORTHO = 2,  /* SimBoxError::NotOrthogonal  simbox.rs:230-236 */
    ORC_EEMPTY = 3,     /* GroupError::EmptyGroup      analysis.rs:106-108 */
    ORC_ENOPOS = 4,     /* PositionError::NoPosition(i) */
    ORC_ENOMASS = 5,    /* MassError::NoMass(i) */
    ORC_EGROUPSIZE = 6, /* RMSDError::InconsistentGroup rmsd.rs:405-422 */
    ORC_EZEROBOX = 7    /* reference panics: vector3d.rs:402,576 */
};

/* Dimension enum, dimension.rs:15-25 */
enum { ORC_DIM_NONE = 0, ORC_DIM_X, ORC_DIM_Y, ORC_DIM_Z, ORC_DIM_XY, ORC_DIM_XZ, ORC_DIM_YZ, ORC_DIM_XYZ };

/* ---- scalar primitives (ref32) ---- */
float orc_wrap1(float x, float L);                 /* vector3d.rs:401-417 */
float orc_minimg1(float d, float L);               /* vector3d.rs:575-592 */
float orc_floor_mod(float x, float y);             /* vector3d.rs:28-30 */
void orc_vector_to(const float c[3], const float p[3], const float L[3], float out[3]); /* vector3d.rs:561-569 */
float orc_distance(const float a[3], const float b[3], int dim, const float L[3]);      /* vector3d.rs:458-486 */
int orc_box_lengths(const float box9[9], float L[3]); /* matrix2simbox io/xdrfile.rs:170-187 + simbox_check simbox.rs:230 */

/* ---- group ops, ref32 ---- */
int orc_estimate_center(const float *xyz, size_t stride, const uint32_t *idx, size_t g,
                        const float *mass, const float L[3], float out[3]); /* iterators.rs:1152-1191 / 1314-1357; mass==NULL -> 1.0 */
int orc_get_center(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3], float out[3]); /* iterators.rs:1237-1266 */
int orc_get_com(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass,
                const float L[3], float out[3]);                                                                   /* iterators.rs:1404-1438 */
int orc_get_center_naive(const float *xyz, size_t stride, const uint32_t *idx, size_t g, float out[3]);            /* iterators.rs:886-903 */
int orc_group_distance(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2,
                       size_t g2, int dim, const float L[3], float *out);                                          /* analysis.rs:348-360 */
int orc_all_distances(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2,
                      size_t g2, int dim, const float L[3], float *out /* g1*g2 row-major */);                     /* analysis.rs:401-427 */
/* documented consumer of the matrix: min_by keeps the FIRST minimum, max_by the LAST maximum (analysis.rs:390-399) */
int orc_all_distances_minmax(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2,
                             size_t g2, int dim, const float L[3], float *dmin, uint32_t *imin /*[2] = (i,j)*/,
                             float *dmax, uint32_t *imax, float cutoff, uint64_t *count_below);
int orc_wrap(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3], int8_t *shifts /* g*3 or NULL */); /* iterators.rs:1548, vector3d.rs:380 */
int orc_pairs_within(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2, size_t g2, float cutoff,
                     const float L[3], uint64_t *count, uint32_t *pairs, float *dist, size_t capacity); /* cellgrid.rs:301-420 */
#define ORC_NO_MOL 0xFFFFFFFFu
int orc_make_group_whole(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float L[3]);   /* modifying.rs:437-465 */
int orc_make_molecules_whole(float *xyz, size_t stride, size_t n, const uint32_t *mol_ref, const float L[3]); /* modifying.rs:338-391 */
int orc_atoms_center(float *xyz, size_t stride, size_t n, const uint32_t *idx, size_t g, const float *mass /* nullable */, int dim,
                     const float L[3]);                                                                /* utility.rs:109-189 */
int orc_translate(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float t[3], const float L[3],
                  int8_t *shifts);                                                                                /* atom.rs:498-511 */

/* ---- RMSD / Kabsch, ref32 ---- */
int orc_rmsd_extract(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass,
                     const float L[3], float *y /* g*3 */, float bc[3], float com[3]);  /* rmsd.rs:425-446,479-492 */
void orc_kabsch(const float *p, const float *q, const float *w, size_t g, const float cp[3], const float cq[3],
                float sum_w, float r[9] /* row-major */, float t[3], float *rmsd);     /* rmsd.rs:547-603 */
int orc_calc_rmsd(const float *ref_xyz, size_t ref_stride, const uint32_t *ref_idx, size_t g_ref, const float ref_L[3],
                  const float *mass /* reference group order */, const float *tgt_xyz, size_t tgt_stride,
                  const uint32_t *tgt_idx, size_t g_tgt, const float tgt_L[3], float r[9], float *rmsd); /* rmsd.rs:141-166 */
void orc_fit(float *xyz, size_t stride, size_t n, const float r[9], const float com_tgt[3], const float com_ref[3],
             const float L[3]);                                                        /* rmsd.rs:508-528 */

/* ---- exact64 flavour ---- */
int orc_estimate_center_x64(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass,
                            const float L[3], double out[3]);
int orc_get_center_x64(const float *xyz, size_t stride, const uint32_t *idx, size_t g, const float *mass /* NULL -> geometry */,
                       const float L[3], double out[3]);
int orc_calc_rmsd_x64(const float *ref_xyz, size_t ref_stride, const uint32_t *ref_idx, size_t g_ref, const float ref_L[3],
                      const float *mass, const float *tgt_xyz, size_t tgt_stride, const uint32_t *tgt_idx, size_t g_tgt,
                      const float tgt_L[3], double r[9], double *rmsd);

/* ---- triclinic extension (UNPINNED by the reference; see header comment) ---- */
/* box9 = row-major matrix, rows are box vectors v1=(a,0,0) v2=(b,c,0) v3=(d,e,f) */
void orc_tric_wrap1(float p[3], const float box9[9], int shifts[3] /* kx,ky,kz */);
int orc_tric_wrap(float *xyz, size_t stride, const uint32_t *idx, size_t g, const float box9[9], int8_t *shifts);
float orc_tric_distance(const float a[3], const float b[3], int dim, const float box9[9]);
int orc_tric_all_distances(const float *xyz, size_t stride, const uint32_t *idx1, size_t g1, const uint32_t *idx2,
                           size_t g2, int dim, const float box9[9], float *out);
double orc_tric_distance_brute64(const float a[3], const float b[3], int dim, const float box9[9], int nimg);

/* ---- deterministic synthetic workloads (bit-identical to the device generator) ---- */
uint64_t orc_splitmix64(uint64_t x);
uint64_t orc_hash(uint64_t seed, uint64_t frame, uint64_t atom, uint64_t axis);
/* uniform in [lo, lo+span) per axis: x = lo + u*span, u = (h>>40)*2^-24 */
void orc_synth_uniform(float *xyz, size_t n_atoms, uint64_t seed, uint64_t frame, const float lo[3], const float span[3]);
/* rigid blob: x = R*p_i + c + noise_i, p_i/noise_i Irwin-Hall(4) of 22-bit ints; optional one-step wrap into L */
void orc_synth_blob_ref(float *xyz, size_t n_atoms, uint64_t seed, float scale, const float centre[3]);
void orc_synth_blob_frame(float *xyz, size_t n_atoms, uint64_t seed, uint64_t frame, float scale, float nscale,
                          const float rot[9], const float centre[3], const float L[3], int wrap);

/* ---- restated groan_rs CPU trajectory path (baseline) ---- */
/* parallel.rs:208-269,425-448: T threads, thread t handles frames t, t+T, ...; each thread owns a
 * cloned "System" of 240-byte atom records, scatters the frame into it (xdrfile_xtc.rs:88-104),
 * then runs the requested ops.  ops bitmask: 1 = group_get_center, 2 = calc_rmsd, 4 = fit, 8 = atoms_wrap.
 * Returns wall seconds of the threaded region (frame generation/copy-in excluded). */
double orc_baseline_traj(const float *frames /* F*n*3 */, const float *boxes /* F*3 lengths */, size_t F, size_t n_atoms,
                         const uint32_t *idx, size_t g, const float *mass_all /* n_atoms */, const float *ref_xyz,
                         const float ref_L[3], int ops, int n_threads, float *centers /* F*3 */, float *rmsd /* F */);
double orc_baseline_traj_cyclic(const float *frames, const float *boxes, size_t F_store, size_t F_total, size_t n_atoms,
                                const uint32_t *idx, size_t g, const float *mass_all, const float *ref_xyz, const float ref_L[3],
                                int ops, int n_threads, float *centers, float *rmsd);
double orc_baseline_pairs(const float *frames, const float *boxes, size_t F, size_t n_atoms, const uint32_t *idx1, size_t g1,
                          const uint32_t *idx2, size_t g2, int dim, int n_threads, float *dmin /* F */, float *dmax /* F */);

#ifdef __cplusplus
}
#endif
#endif

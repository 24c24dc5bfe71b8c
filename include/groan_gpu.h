/*
 * groan_gpu.h -- C ABI of libgroan_gpu.so: groan_rs's per-frame PBC geometry hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary a groan_rs maintainer binds from Rust (see INTEGRATION.md for the
 * extern "C" block and the safe wrapper).  Conventions mirror the reference's only existing FFI,
 * the xdrfile binding (src/io/xdrfile.rs:27-100): opaque handle, caller-owned buffers, plain
 * pointers and sizes, `int` status (0 = ok), no exceptions or panics across the boundary.
 *
 * Data model (one ctx = one GPU = one host thread; multi-GPU = one ctx per rank):
 *   - a ctx is created for a System of `n_atoms` atoms and batches of up to `max_frames` frames;
 *   - groups are ascending, unique atom-index lists, exactly what Group/AtomContainer iterate
 *     (src/structures/container.rs:51-115,415-436), optionally with per-atom masses in group order;
 *   - a batch of frames is pushed as the xtc reader emits it: coordinates F x N x 3 f32 AoS
 *     (src/io/xtc_io/xdrfile_xtc.rs:25-31) and one 3x3 row-major box matrix per frame, box[i][j] =
 *     component j of box vector i (src/io/xdrfile.rs:170-187);
 *   - every op below evaluates one reference function for EVERY frame of the current batch.
 *
 * Output pointers may be device memory, pinned host memory (both: the call is asynchronous on the
 * ctx's compute stream; use groan_gpu_sync) or pageable host memory (the call blocks until the
 * result is there).  There is no CPU fallback: every op runs CUDA kernels or fails.
 */
#ifndef GROAN_GPU_H
#define GROAN_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct groan_gpu_ctx groan_gpu_ctx;

/* Status codes.  1..8 map 1:1 onto the reference's error enums so a Rust wrapper can rebuild them. */
enum groan_status {
    GROAN_OK = 0,
    GROAN_ENOBOX = 1,     /* SimBoxError::DoesNotExist          src/errors.rs:556-582, simbox.rs:230-236 */
    GROAN_ENOTORTHO = 2,  /* SimBoxError::NotOrthogonal         simbox.rs:185,230-236 */
    GROAN_EEMPTY = 3,     /* GroupError::EmptyGroup / RMSDError::EmptyGroup   analysis.rs:106-108 */
    GROAN_ENOPOS = 4,     /* PositionError::NoPosition(index)   errors.rs:227-258; index via groan_gpu_error_detail */
    GROAN_ENOMASS = 5,    /* MassError::NoMass(index)           errors.rs:290-305 */
    GROAN_EGROUPSIZE = 6, /* RMSDError::InconsistentGroup(n_ref, n_target)    rmsd.rs:405-422 */
    GROAN_EZEROBOX = 7,   /* reference panics "Box len should not be zero"    vector3d.rs:402,576 */
    GROAN_ENOGROUP = 8,   /* GroupError::NotFound */
    GROAN_EINVAL = 9,     /* bad argument (null pointer, unsorted indices, index >= n_atoms, ...) */
    GROAN_ECUDA = 10,     /* CUDA runtime failure; groan_gpu_last_cuda_error() has the text */
    GROAN_ENOFRAMES = 11, /* no batch pushed / attached */
    GROAN_ENOREF = 12,    /* groan_gpu_rmsd* before groan_gpu_rmsd_set_reference */
    GROAN_ECAPACITY = 13  /* more frames than max_frames */
};

/* Dimension, numerically the reference's enum order (src/structures/dimension.rs:15-25). */
enum groan_dim {
    GROAN_DIM_NONE = 0, GROAN_DIM_X = 1, GROAN_DIM_Y = 2, GROAN_DIM_Z = 3,
    GROAN_DIM_XY = 4, GROAN_DIM_XZ = 5, GROAN_DIM_YZ = 6, GROAN_DIM_XYZ = 7
};

#define GROAN_GROUP_ALL (-1) /* "all": every atom of the system (System::atoms_wrap, atoms_translate) */
#define GROAN_MAX_GROUPS 64

/* ctx flags */
#define GROAN_FLAG_TRICLINIC 1u  /* enable the triclinic EXTENSION (wrap, min-image distances); without it a
                                    non-orthogonal box returns GROAN_ENOTORTHO exactly like the reference */
#define GROAN_FLAG_EXACT_ONLY 2u /* disable the single-pass fast paths; always run the exact (reference-order) passes */
#define GROAN_FLAG_NO_TMA 4u     /* contiguous groups through the gather kernels (register-staged 256-bit loads) instead of the
                                    TMA-fed ring kernels; for tests (two independent implementations of the same sums) */
#define GROAN_FLAG_HOST_FALLBACK 16u /* launch the passes that re-do flagged frames from the host after every single-pass kernel
                                        instead of letting the kernel tail-launch them from the device when a frame needs them */

/* ---- lifetime -------------------------------------------------------------------------------- */
int groan_gpu_create(int device, size_t n_atoms, size_t max_frames, groan_gpu_ctx **out);
void groan_gpu_destroy(groan_gpu_ctx *ctx);
int groan_gpu_set_flags(groan_gpu_ctx *ctx, unsigned flags);
/* run all kernels on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = ctx's own */
int groan_gpu_set_stream(groan_gpu_ctx *ctx, void *cuda_stream);
int groan_gpu_sync(groan_gpu_ctx *ctx);
const char *groan_gpu_strerror(int status);
const char *groan_gpu_last_cuda_error(groan_gpu_ctx *ctx);
/* frame and atom index behind the last GROAN_ENOPOS / GROAN_ENOMASS / GROAN_EGROUPSIZE (a = n_ref, b = n_target) */
int groan_gpu_error_detail(groan_gpu_ctx *ctx, size_t *a, size_t *b);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t groan_gpu_launch_count(groan_gpu_ctx *ctx);
/* diagnostics (synchronises): how many frames of the last get_center / get_com / rmsd call the single-pass kernel could
 * not certify and handed to the reference-order passes (non-compact group, centre on the box edge, RMSD below f32 resolution) */
int groan_gpu_fallback_frames(groan_gpu_ctx *ctx, size_t *n);
/* diagnostics (synchronises): how many frames of the last centre + RMSD call (groan_gpu_center_rmsd) went through the second,
 * sine-sum centre pass because the moments the fused kernel accumulates could not decide the centre's periodic image (the mean
 * of the group within ~0.1 nm of a box face); still a single-pass kernel, one more read of those frames only */
int groan_gpu_second_pass_frames(groan_gpu_ctx *ctx, size_t *n);

/* ---- groups: Group::from_indices, container.rs:51-115 ---------------------------------------- */
/* idx ascending and unique, each < n_atoms; mass nullable (ops that need masses then fail GROAN_ENOMASS,
 * mass[i] < 0 marks "atom idx[i] has no mass").
 * gid == GROAN_GROUP_ALL attaches masses to the built-in group of every atom ("all" / "All" exist from System::new on,
 * src/system/mod.rs:150-170): idx must be NULL and n == n_atoms. */
int groan_gpu_set_group(groan_gpu_ctx *ctx, int gid, const uint32_t *idx, size_t n, const float *mass);

/* ---- frames: FrameData::update_system, xdrfile_xtc.rs:88-104 ---------------------------------- */
/* host -> device staging of a batch: chunks go through pinned buffers with cudaMemcpyAsync on the ctx's
 * copy stream into the slot NOT used by the previous batch, so the copy overlaps kernels still running
 * on the previous batch.  box: F x 9 row-major matrices, NULL = system has no box. */
int groan_gpu_push_frames(groan_gpu_ctx *ctx, const float *xyz, const float *box, size_t n_frames);
/* The same batch in the form the xtc decoder holds it one step before it emits floats: every xtc coordinate is an integer
 * lattice point, and the float the reader hands out is (float)int * (1 / precision) (external/xdrfile/xdrfile.c:844,915-917;
 * molly does the same).  Uploading the integers halves the bytes over PCIe, the bound of the host-fed path, when the span of
 * a frame fits 16 bits (65 nm at the usual precision of 1000); the floats are rebuilt on the device with the reader's own
 * expression, bit for bit (tests/test_gpu_parity.py::test_quantized_frames_are_the_readers_floats).
 * q: F x N x 3 integers of elem_bytes (2 = int16, 4 = int32); origin: F x 3 int32 added to every q of the frame (NULL = 0). */
int groan_gpu_push_frames_quantized(groan_gpu_ctx *ctx, const void *q, int elem_bytes, const int32_t *origin, float precision,
                                    const float *box, size_t n_frames);
/* ... and the current batch in the form the xtc ENCODER starts from (writing fitted trajectories): the
 * integer lattice point of every coordinate, (int)(x * precision +- 0.5) in f32 exactly as external/xdrfile/xdrfile.c:1018-1031
 * computes it.  q_out: F x N x 3 int32, host or device. */
int groan_gpu_get_frames_quantized(groan_gpu_ctx *ctx, int32_t *q_out, float precision);
/* zero-copy: operate in place on a caller-owned DEVICE buffer of F x N x 3 floats (16-byte aligned) */
int groan_gpu_attach_frames(groan_gpu_ctx *ctx, float *d_xyz, const float *box, size_t n_frames);
/* Option<Vector3D> positions: valid[f*N + i] == 0 marks "atom i has no position in frame f" (host array,
 * NULL = all valid, the default after every push/attach) */
int groan_gpu_set_valid(groan_gpu_ctx *ctx, const uint8_t *valid);
/* read the (possibly wrapped / translated / fitted) current batch back: F x N x 3 */
int groan_gpu_get_frames(groan_gpu_ctx *ctx, float *xyz_out);

/* ---- centres ---------------------------------------------------------------------------------- */
/* System::group_estimate_center / group_estimate_com  (analysis.rs:52; iterators.rs:1152-1191,1314-1357) */
int groan_gpu_estimate_center(groan_gpu_ctx *ctx, int gid, int weighted, float *out /* F x 3 */);
/* System::group_get_center / group_get_com            (analysis.rs:105,258; iterators.rs:1237-1266,1404-1438) */
int groan_gpu_get_center(groan_gpu_ctx *ctx, int gid, int weighted, float *out /* F x 3 */);
/* System::group_get_center_naive (iterators.rs:886-903) */
int groan_gpu_get_center_naive(groan_gpu_ctx *ctx, int gid, float *out /* F x 3 */);

/* ---- distances -------------------------------------------------------------------------------- */
/* System::group_distance (analysis.rs:348-360) */
int groan_gpu_group_distance(groan_gpu_ctx *ctx, int g1, int g2, int dim, float *out /* F */);
/* System::group_all_distances (analysis.rs:401-427): F x n1 x n2 row-major */
int groan_gpu_all_distances(groan_gpu_ctx *ctx, int g1, int g2, int dim, float *out);
/* the documented consumer of that matrix (analysis.rs:390-399), fused so the matrix is never written:
 * per frame min (FIRST minimum in row-major order, Iterator::min_by) and max (LAST maximum,
 * Iterator::max_by) with their (i, j) positions inside the groups, and the number of pairs with
 * distance < cutoff.  Any output pointer may be NULL. */
int groan_gpu_all_distances_reduce(groan_gpu_ctx *ctx, int g1, int g2, int dim, float cutoff, float *dmin /* F */,
                                   uint32_t *imin /* F x 2 */, float *dmax /* F */, uint32_t *imax /* F x 2 */,
                                   uint64_t *count /* F */);

/* ---- cutoff pair search through a cell grid -------------------------------- */
/* CellGrid::new over group g2 + CellGrid::neighbors_iter around every atom of g1 + the distance filter its users apply
 * (cellgrid.rs:301-420; guess.rs:362-470, hbonds.rs:240-335): every pair (i in g1, j in g2) with
 * Vector3D::distance(pos_i, pos_j, XYZ) < cutoff, per frame.  Distances are bit-identical to groan_gpu_all_distances.
 * count: F; pairs (nullable): F x capacity x 2 positions inside (g1, g2); dist (nullable, needs pairs): F x capacity.
 * A frame with more than `capacity` pairs keeps the first `capacity` it found and still reports the full count.
 * The order of the pairs is undefined (as the order of neighbors_iter is, cellgrid.rs:141-144).  Orthogonal boxes only. */
int groan_gpu_pairs_within(groan_gpu_ctx *ctx, int g1, int g2, float cutoff, uint64_t *count, uint32_t *pairs, float *dist,
                           size_t capacity);

/* ---- users of the cell grid ---------------------------------------------------- */
/* System::guess_bonds (src/system/guess.rs:362-470): per frame, every pair of atoms i < j with van der Waals radii whose
 * minimum-image distance is below (vdw[i] + vdw[j]) * radius_factor (reference default 0.55).  vdw: n_atoms floats, < 0 = the
 * atom has no radius (it gets no bonds; the reference lists it in its BondsGuessWarning).  count: F; pairs (nullable):
 * F x capacity x 2 atom indices, unordered; pairs beyond the capacity are counted but not stored.  Assigning the bonds to
 * atoms and the sanity checks on their number stay on the host (guess.rs:409-425,472-). */
int groan_gpu_guess_bonds(groan_gpu_ctx *ctx, const float *vdw, float radius_factor, uint64_t *count, uint32_t *pairs, size_t capacity);
/* HBondAnalysis::analyze_single (src/system/hbonds.rs:240-320) for one acceptor group and one list of donors, per frame:
 * donors[d] carries the hydrogens hydrogens[hyd_offsets[d] .. hyd_offsets[d + 1]) (HBondChainGroups, hbonds.rs:108-150).
 * A record for every (donor, hydrogen, acceptor) with acceptor != donor, distance(acceptor, donor) <= max_distance and
 * angle(donor - hydrogen - acceptor) >= min_angle (degrees; calc_angle hbonds.rs:322-335).  count: F; dha (nullable, together
 * with dist_angle): F x capacity x 3 atom indices; dist_angle: F x capacity x 2 (distance in nm, angle in degrees).
 * The order of the records is undefined (as is the order of CellGrid::neighbors_iter, cellgrid.rs:141-144). */
int groan_gpu_hbonds(groan_gpu_ctx *ctx, int acc_gid, const uint32_t *donors, const uint32_t *hyd_offsets, const uint32_t *hydrogens,
                     size_t n_donors, float max_distance, float min_angle, uint64_t *count, uint32_t *dha, float *dist_angle, size_t capacity);

/* ---- wrap / translate (in place on the current batch) ------------------------------------------ */
/* System::atoms_wrap / group_wrap (modifying.rs:201,215; vector3d.rs:380-417).  shifts (nullable):
 * F x G x 3 int8, net number of +L steps per axis (for the triclinic extension: multiples of box vectors) */
int groan_gpu_wrap(groan_gpu_ctx *ctx, int gid, int8_t *shifts);
/* System::atoms_translate / group_translate (modifying.rs:73; atom.rs:498-511) */
int groan_gpu_translate(groan_gpu_ctx *ctx, int gid, const float t[3], int8_t *shifts);

/* ---- whole groups / molecules, centering (in place on the current batch) ---- */
/* System::make_group_whole (modifying.rs:437-465): every atom of the group goes to c + vector_to(c, pos) with
 * c = group_estimate_center of the frame.  Errors as group_estimate_center. */
int groan_gpu_make_group_whole(groan_gpu_ctx *ctx, int gid);
/* The molecule topology lives on the host (bonds, System::mol_references: modifying.rs:258-283).  mol_ref[i], i < n_atoms:
 * index of the reference atom of atom i's molecule = the molecule's lowest atom index (mol_ref[r] == r for a reference atom),
 * GROAN_NO_MOLECULE for atoms of monoatomic molecules, which make_molecules_whole leaves untouched. */
#define GROAN_NO_MOLECULE 0xFFFFFFFFu
int groan_gpu_set_molecules(groan_gpu_ctx *ctx, const uint32_t *mol_ref);
/* System::make_molecules_whole (modifying.rs:338-391): reference atoms are wrapped into the box, every other atom of a
 * polyatomic molecule goes to ref + vector_to(ref, pos).  GROAN_EINVAL before groan_gpu_set_molecules. */
int groan_gpu_make_molecules_whole(groan_gpu_ctx *ctx);
/* System::atoms_center (weighted = 0) / atoms_center_mass (1) (utility.rs:109-130,168-189): all atoms are translated by
 * box_centre - group_estimate_center/com(gid), restricted to the axes of dim (GROAN_DIM_*), and wrapped. */
int groan_gpu_atoms_center(groan_gpu_ctx *ctx, int gid, int weighted, int dim);

/* ---- RMSD / Kabsch ----------------------------------------------------------------------------- */
/* RMSDConverterAnalyzer::new (rmsd.rs:186-203): the reference may be a different System (own atom count,
 * own index list for the same group name, rmsd.rs:823-841).  ref_mass: the n_ref masses of the REFERENCE
 * system's group (they weight the RMSD sum and reference.group_get_com, rmsd.rs:154,192); NULL = use the
 * masses given to groan_gpu_set_group (then n_ref must equal the group's size).  The target's own
 * group_get_com always uses the set_group masses.  ref_idx == NULL: the group is the first n_ref atoms of the reference
 * (n_ref <= n_ref_atoms).  gid may be GROAN_GROUP_ALL. */
int groan_gpu_rmsd_set_reference(groan_gpu_ctx *ctx, int gid, const float *ref_xyz, size_t n_ref_atoms,
                                 const uint32_t *ref_idx, size_t n_ref, const float ref_box[9], const float *ref_mass);
/* System::calc_rmsd / RMSDTrajRead::calc_rmsd (rmsd.rs:75,315): rmsd F; rot (nullable) F x 9 row-major r */
int groan_gpu_rmsd(groan_gpu_ctx *ctx, int gid, float *rmsd, float *rot);
/* group_get_center (weighted = 0) or group_get_com (weighted = 1) AND calc_rmsd of the same group from ONE read of every
 * frame: what a FrameAnalyze implementor that needs both (traj_convert.rs:76-83) would call per batch.  Results are those
 * of groan_gpu_get_center + groan_gpu_rmsd. */
int groan_gpu_center_rmsd(groan_gpu_ctx *ctx, int gid, int weighted, float *center /* F x 3 */, float *rmsd /* F */,
                          float *rot /* nullable */);
/* System::calc_rmsd_and_fit / RMSDTrajRead::calc_rmsd_and_fit (rmsd.rs:129,390; fit_structure :508-528):
 * also fits ALL atoms of every frame in place */
int groan_gpu_rmsd_fit(groan_gpu_ctx *ctx, int gid, float *rmsd);

/* ---- synthetic workloads (bench / test support; bit-identical to oracle/groan_oracle.c) -------- */
/* fills the ctx's current device slot with F generated frames and makes it the current batch */
int groan_gpu_synth_uniform(groan_gpu_ctx *ctx, uint64_t seed, uint64_t frame0, size_t n_frames, const float lo[3],
                            const float span[3], const float *box);
int groan_gpu_synth_blob(groan_gpu_ctx *ctx, uint64_t seed, uint64_t frame0, size_t n_frames, float scale, float nscale,
                         const float *rot /* F x 9 */, const float *centre /* F x 3 */, const float *box, int wrap);
/* the blob's reference structure (n_atoms x 3) into a caller buffer (host or device) */
int groan_gpu_synth_blob_ref(groan_gpu_ctx *ctx, uint64_t seed, float scale, const float centre[3], float *xyz_out);

#ifdef __cplusplus
}
#endif
#endif /* GROAN_GPU_H */

/*
 * groan_xtc.h -- xtc frames in and out of the GPU path (part of libgroan_gpu.so).
 *
 * What it replaces in the reference:
 *   - reading:  read_xtc (external/xdrfile/xdrfile_xtc.h:45-52, bound in src/io/xdrfile.rs:27-100 and called from
 *     src/io/xtc_io/xdrfile_xtc.rs:63-83) and the partial-frame GroupXtcReader built on molly
 *     (src/io/xtc_io/molly_xtc.rs:404-560): here a whole BATCH of frames is decoded at once -- on the host by a pool of
 *     threads into the integers groan_gpu_push_frames_quantized uploads, or on the GPU from the file's own bytes;
 *   - writing:  write_xtc (xdrfile_xtc.h:55-59; XtcWriter::write_frame, src/io/xtc_io/mod.rs:300-330): the fitted batch is
 *     quantised on the device and encoded on the host, byte for byte what the reference writes.
 * Conventions as in groan_gpu.h: plain pointers and sizes, caller-owned buffers, int status.
 */
#ifndef GROAN_XTC_H
#define GROAN_XTC_H

#include <stddef.h>
#include <stdint.h>

#include "groan_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* status codes of the groan_xtc_* functions (0 = ok) */
enum groan_xtc_status {
    GROAN_XTC_OK = 0,
    GROAN_XTC_EOF = 1,       /* no frame at this offset: end of the data (read_xtc: exdrENDOFFILE, xdrfile_xtc.rs:65-83) */
    GROAN_XTC_EMAGIC = 2,    /* not an xtc frame (magic number != 1995) */
    GROAN_XTC_ETRUNC = 3,    /* the data end inside a frame */
    GROAN_XTC_EFORMAT = 4,   /* inconsistent header or damaged bit stream */
    GROAN_XTC_ERAW = 5,      /* frame of <= 9 atoms: stored as plain floats, it has no integer lattice */
    GROAN_XTC_ECAPACITY = 6, /* output buffer too small / more frames than asked for */
    GROAN_XTC_ERANGE = 7,    /* int16 output requested but a frame spans more than 65535 lattice steps */
    GROAN_XTC_EINVAL = 8
};

/* Walk the frame headers of an xtc file held in memory: offsets[f] = where frame f starts, offsets[n] = where the last
 * one ends (capacity max_frames + 1 entries).  natoms: atoms per frame (all frames must agree, as in the reference:
 * xtc_io/mod.rs:110-125).  Stops after max_frames frames. */
int groan_xtc_scan(const uint8_t *data, size_t len, size_t max_frames, uint64_t *offsets, int32_t *natoms, size_t *n_frames);

/* Decode frames offsets[0..n_frames) with n_threads host threads (frames are independent).
 * atoms (nullable): ascending atom indices to keep -- GroupXtcReader semantics (molly_xtc.rs:441-462): decoding stops at the
 * last wanted atom and only those atoms are written, n_out = n_sel per frame; NULL = all atoms, n_out = natoms.
 * Any of the coordinate outputs may be NULL:
 *   xyz  F x n_out x 3 floats, exactly read_xtc's: (float)int * (float)(1.0 / precision) (xdrfile.c:844,915-917)
 *   q32  F x n_out x 3 lattice integers
 *   q16  F x n_out x 3 lattice integers minus origin[f] (F x 3, required with q16): the form groan_gpu_push_frames_quantized
 *        uploads; GROAN_XTC_ERANGE if a frame does not fit 16 bits
 * box F x 9, step F, time F, precision F: nullable. */
int groan_xtc_decode(const uint8_t *data, size_t len, const uint64_t *offsets, size_t n_frames, int n_threads, const uint32_t *atoms,
                     size_t n_sel, float *xyz, int32_t *q32, int16_t *q16, int32_t *origin, float *box, int32_t *step, float *time,
                     float *precision);

/* Encode n_frames frames of n_atoms atoms (xyz: floats, quantised like write_xtc does, xdrfile.c:1018-1031; or q: lattice
 * integers; exactly one of the two) into out[0..capacity); *len = bytes written.  Byte-identical to write_xtc. */
int groan_xtc_encode(const float *xyz, const int32_t *q, size_t n_frames, size_t n_atoms, const float *box, const int32_t *step,
                     const float *time, float precision, int n_threads, uint8_t *out, size_t capacity, size_t *len);

/* ---- GPU side ---------------------------------------------------------------------------------------------------- */
/* Stage a batch straight from the file's bytes: data[offsets[0] .. offsets[n_frames]) is copied host -> device as it is
 * (pinned source: one cudaMemcpyAsync on the copy stream) and decoded on the GPU into the ctx's frame slot; the boxes come
 * from the frame headers.  step / time / precision (nullable, n_frames each) receive the header fields.
 * Frames of <= 9 atoms are not compressed in the file: GROAN_EINVAL (use groan_gpu_push_frames). */
int groan_gpu_push_xtc(groan_gpu_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *offsets, size_t n_frames, int32_t *step,
                       float *time, float *precision);
/* number of frames of the last groan_gpu_push_xtc whose bit stream was damaged (synchronises) */
int groan_gpu_xtc_bad_frames(groan_gpu_ctx *ctx, size_t *n);

/* Partial frames (GroupXtcReader, molly_xtc.rs:404-470): xyz_sel holds only the atoms `atoms` (ascending, n_sel of them) of
 * every frame, F x n_sel x 3.  Only those bytes cross PCIe; on the device they are scattered into full frames in which every
 * other atom keeps its previous (stale) value, exactly like the reference's System after a partial read. */
int groan_gpu_push_group_frames(groan_gpu_ctx *ctx, const float *xyz_sel, const uint32_t *atoms, size_t n_sel, const float *box,
                                size_t n_frames);

/* The current batch (e.g. after groan_gpu_rmsd_fit) as xtc frames: quantised on the device with the writer's rounding,
 * encoded on the host with n_threads threads.  step / time: n_frames each. */
int groan_gpu_write_xtc(groan_gpu_ctx *ctx, float precision, const int32_t *step, const float *time, int n_threads, uint8_t *out,
                        size_t capacity, size_t *len);

#ifdef __cplusplus
}
#endif
#endif /* GROAN_XTC_H */

// groan_xtc.cu -- include/groan_xtc.h: xtc frames in and out of the GPU path.
//
// Host side: frame scan, a multi-threaded decoder (integer lattice or floats, all atoms or a sorted subset) and the
// encoder (xtc_codec.hpp).  Device side: the file's bytes are uploaded as they are and decoded by one warp per frame
// (kernels_xtc.cuh); partial frames are uploaded compact and scattered on the device.
#include "../../include/groan_xtc.h"

#include <atomic>
#include <thread>

#include "ctx.cuh"
#include "kernels_xtc.cuh"
#include "xtc_codec.hpp"

using namespace groan;
using namespace groan_host;
namespace gx = groan_xtc;

namespace {

int xtc_status(int st) {
    switch (st) {
    case gx::XTC_OK: return GROAN_XTC_OK;
    case gx::XTC_EOF: return GROAN_XTC_EOF;
    case gx::XTC_EMAGIC: return GROAN_XTC_EMAGIC;
    case gx::XTC_ETRUNC: return GROAN_XTC_ETRUNC;
    case gx::XTC_ERAW: return GROAN_XTC_ERAW;
    case gx::XTC_ECAPACITY: return GROAN_XTC_ECAPACITY;
    default: return GROAN_XTC_EFORMAT;
    }
}

// run fn(f) for f in [0, n) on up to n_threads threads; frames are handed out one at a time (they differ in cost);
// returns the first non-zero status
template <typename F>
int parallel_frames(size_t n, int n_threads, F &&fn) {
    const size_t T = std::max<size_t>(1, std::min<size_t>(n, (size_t)std::max(1, n_threads)));
    std::atomic<size_t> next(0);
    std::atomic<int> status(0);
    auto work = [&]() {
        for (;;) {
            const size_t f = next.fetch_add(1);
            if (f >= n || status.load() != 0) return;
            const int st = fn(f);
            if (st) {
                int expected = 0;
                status.compare_exchange_strong(expected, st);
            }
        }
    };
    if (T == 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (size_t t = 1; t < T; t++) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
    }
    return status.load();
}

}  // namespace

extern "C" {

int groan_xtc_scan(const uint8_t *data, size_t len, size_t max_frames, uint64_t *offsets, int32_t *natoms, size_t *n_frames) {
    if (!data || !offsets || !n_frames) return GROAN_XTC_EINVAL;
    *n_frames = 0;
    uint64_t off = 0;
    int32_t n0 = -1;
    while (*n_frames < max_frames) {
        gx::FrameInfo fi;
        const int st = gx::parse_frame(data, len, off, &fi);
        if (st == gx::XTC_EOF) break;
        if (st) return xtc_status(st);
        if (n0 < 0) n0 = fi.natoms;
        else if (fi.natoms != n0) return GROAN_XTC_EFORMAT;  // xtc_io/mod.rs:110-125: every frame has the system's atoms
        offsets[*n_frames] = off;
        (*n_frames)++;
        off = fi.next;
        offsets[*n_frames] = off;
    }
    if (*n_frames == 0) offsets[0] = 0;
    if (natoms) *natoms = n0 < 0 ? 0 : n0;
    return GROAN_XTC_OK;
}

int groan_xtc_decode(const uint8_t *data, size_t len, const uint64_t *offsets, size_t n_frames, int n_threads, const uint32_t *atoms,
                     size_t n_sel, float *xyz, int32_t *q32, int16_t *q16, int32_t *origin, float *box, int32_t *step, float *time,
                     float *precision) {
    if (!data || !offsets || (q16 && !origin) || (atoms && n_sel == 0)) return GROAN_XTC_EINVAL;
    for (size_t k = 1; atoms && k < n_sel; k++)
        if (atoms[k] <= atoms[k - 1]) return GROAN_XTC_EINVAL;
    return parallel_frames(n_frames, n_threads, [&](size_t f) -> int {
        gx::FrameInfo fi;
        int st = gx::parse_frame(data, len, offsets[f], &fi);
        if (st) return xtc_status(st);
        if (atoms && atoms[n_sel - 1] >= (uint32_t)fi.natoms) return GROAN_XTC_EINVAL;
        const size_t n_out = atoms ? n_sel : (size_t)fi.natoms;
        if (box) std::memcpy(box + f * 9, fi.box, sizeof(fi.box));
        if (step) step[f] = fi.step;
        if (time) time[f] = fi.time;
        if (precision) precision[f] = fi.raw ? 0.0f : fi.precision;
        float *fx = xyz ? xyz + f * n_out * 3 : nullptr;
        if (fi.raw) {
            if (q32 || q16) return GROAN_XTC_ERAW;
            if (fx) {
                size_t k = 0;
                for (int32_t i = 0; i < fi.natoms; i++) {
                    if (atoms && (k >= n_sel || atoms[k] != (uint32_t)i)) continue;
                    const size_t o = atoms ? k : (size_t)i;
                    for (int c = 0; c < 3; c++) fx[3 * o + c] = gx::be32f(data + fi.payload + 12 * (size_t)i + 4 * c);
                    k++;
                }
            }
            return GROAN_XTC_OK;
        }
        int32_t *fq = q32 ? q32 + f * n_out * 3 : nullptr;
        std::vector<int32_t> tmp;
        if (!fq && q16) {  // int16 needs the frame's extent first
            tmp.resize(n_out * 3);
            fq = tmp.data();
        }
        const float inv = gx::inv_precision(fi.precision);
        const int32_t n_decode = atoms ? (int32_t)atoms[n_sel - 1] + 1 : fi.natoms;
        size_t k = 0;  // next wanted atom (subset mode)
        st = gx::decode_lattice(data, fi, n_decode, [&](int32_t i, int32_t x, int32_t y, int32_t z) {
            size_t o = (size_t)i;
            if (atoms) {
                // the water swap emits atom i + 1 after atom i, so the output order is still ascending
                if (k >= n_sel || atoms[k] != (uint32_t)i) return;
                o = k++;
            }
            if (fq) { fq[3 * o] = x; fq[3 * o + 1] = y; fq[3 * o + 2] = z; }
            if (fx) { fx[3 * o] = (float)x * inv; fx[3 * o + 1] = (float)y * inv; fx[3 * o + 2] = (float)z * inv; }
        });
        if (st) return xtc_status(st);
        if (q16) {
            int32_t lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
            for (size_t a = 0; a < n_out; a++)
                for (int c = 0; c < 3; c++) {
                    lo[c] = std::min(lo[c], fq[3 * a + c]);
                    hi[c] = std::max(hi[c], fq[3 * a + c]);
                }
            int32_t org[3];
            for (int c = 0; c < 3; c++) {
                if ((int64_t)hi[c] - (int64_t)lo[c] > 65535) return GROAN_XTC_ERANGE;
                org[c] = (int32_t)(((int64_t)lo[c] + (int64_t)hi[c] + 1) / 2);
                if (hi[c] - org[c] > 32767) org[c] = hi[c] - 32767;
                origin[f * 3 + c] = org[c];
            }
            int16_t *o16 = q16 + f * n_out * 3;
            for (size_t a = 0; a < n_out; a++)
                for (int c = 0; c < 3; c++) o16[3 * a + c] = (int16_t)(fq[3 * a + c] - org[c]);
        }
        return GROAN_XTC_OK;
    });
}

int groan_xtc_encode(const float *xyz, const int32_t *q, size_t n_frames, size_t n_atoms, const float *box, const int32_t *step,
                     const float *time, float precision, int n_threads, uint8_t *out, size_t capacity, size_t *len) {
    if ((!xyz) == (!q) || !box || !out || !len || n_atoms == 0 || n_atoms > 0x7fffffffu / 3) return GROAN_XTC_EINVAL;
    if (n_atoms <= 9 && !xyz) return GROAN_XTC_ERAW;
    if (!(precision > 0.0f)) precision = 1000.0f;  // xdrfile.c:998-999
    std::vector<std::vector<uint8_t>> enc(n_frames);
    const int rc = parallel_frames(n_frames, n_threads, [&](size_t f) -> int {
        const int32_t st = step ? step[f] : (int32_t)f;
        const float tm = time ? time[f] : 0.0f;
        if (n_atoms <= 9) {
            gx::encode_frame_raw(xyz + f * n_atoms * 3, (int32_t)n_atoms, st, tm, box + f * 9, enc[f]);
            return 0;
        }
        std::vector<int32_t> lat(n_atoms * 3);
        if (q) {
            std::memcpy(lat.data(), q + f * n_atoms * 3, n_atoms * 3 * sizeof(int32_t));
        } else {
            const float *x = xyz + f * n_atoms * 3;
            for (size_t k = 0; k < n_atoms * 3; k++) lat[k] = gx::to_lattice(x[k], precision);
        }
        gx::encode_frame_lattice(lat.data(), (int32_t)n_atoms, st, tm, box + f * 9, precision, enc[f]);
        return 0;
    });
    if (rc) return rc;
    size_t total = 0;
    for (const auto &e : enc) total += e.size();
    *len = total;
    if (total > capacity) return GROAN_XTC_ECAPACITY;
    size_t o = 0;
    for (const auto &e : enc) {
        std::memcpy(out + o, e.data(), e.size());
        o += e.size();
    }
    return GROAN_XTC_OK;
}

// ================================================================================================ GPU side
int groan_gpu_push_xtc(groan_gpu_ctx *ctx, const uint8_t *data, size_t len, const uint64_t *offsets, size_t n_frames, int32_t *step,
                       float *time, float *precision) {
    if (!ctx || !data || !offsets || n_frames == 0) return GROAN_EINVAL;
    if (n_frames > ctx->max_frames) return GROAN_ECAPACITY;
    if (offsets[0] & 3) return GROAN_EINVAL;  // XDR: every item is a multiple of 4 bytes
    std::vector<XtcFrameParams> params(n_frames);
    std::vector<float> boxes(n_frames * 9);
    const uint64_t lo = offsets[0];
    uint64_t hi = lo;
    for (size_t f = 0; f < n_frames; f++) {
        gx::FrameInfo fi;
        const int st = gx::parse_frame(data, len, offsets[f], &fi);
        if (st || fi.raw || (size_t)fi.natoms != ctx->n_atoms) return GROAN_EINVAL;
        XtcFrameParams &p = params[f];
        p.base = fi.payload - lo;
        p.nbytes = fi.nbytes;
        p.natoms = fi.natoms;
        for (int k = 0; k < 3; k++) {
            p.minint[k] = fi.minint[k];
            p.sizeint[k] = fi.sizeint[k];
            p.bitsizeint[k] = fi.bitsizeint[k];
        }
        p.bitsize = fi.bitsize;
        p.smallidx = fi.smallidx;
        p.inv_precision = gx::inv_precision(fi.precision);
        p.recip1 = xtc_recip(fi.sizeint[1]);
        p.recip2 = xtc_recip(fi.sizeint[2]);
        std::memcpy(&boxes[f * 9], fi.box, sizeof(fi.box));
        if (step) step[f] = fi.step;
        if (time) time[f] = fi.time;
        if (precision) precision[f] = fi.precision;
        hi = std::max<uint64_t>(hi, fi.next);
    }
    const size_t bytes = (size_t)(hi - lo);
    // a compressed frame is never larger than its floats: <= 78 bits per atom + header
    const size_t cap = ctx->max_frames * (ctx->n_atoms * 12 + 256) + 64;
    if (bytes + 16 > cap) return GROAN_ECAPACITY;
    int rc = begin_batch(ctx, n_frames, boxes.data(), true);
    if (rc) return rc;
    rc = [&]() -> int {
        ctx->attached = false;
        const int slot = ctx->slot;
        ctx->cur_xyz = ctx->d_slot[slot];
        if (!ctx->d_xtc[0]) {
            for (int s = 0; s < 2; s++) {
                CK(cudaMalloc(&ctx->d_xtc[s], cap));
                CK(cudaMalloc(&ctx->d_xtc_params[s], ctx->max_frames * sizeof(XtcFrameParams)));
            }
            CK(cudaMalloc(&ctx->d_xtc_status, ctx->max_frames * sizeof(int)));
            ctx->xtc_cap = cap;
            unsigned long long recips[73];
            for (int k = 0; k < 73; k++) recips[k] = xtc_recip((uint32_t)gx::kMagicInts[k]);
            CK(cudaMemcpyToSymbol(g_xtc_magic_recip, recips, sizeof(recips)));
        }
        int r = h2d_on_copy_stream(ctx, ctx->d_xtc[slot], data + lo, bytes);
        if (r) return r;
        // the decoder's 64-bit window may look up to 8 bytes past the last frame
        CK(cudaMemsetAsync(reinterpret_cast<char *>(ctx->d_xtc[slot]) + bytes, 0, 16, ctx->copy));
        // pageable source: the runtime stages it before the call returns, `params` may go out of scope
        CK(cudaMemcpyAsync(ctx->d_xtc_params[slot], params.data(), n_frames * sizeof(XtcFrameParams), cudaMemcpyHostToDevice, ctx->copy));
        int r2 = end_batch(ctx);
        if (r2) return r2;
        // The decoder runs on the COMPUTE stream, behind the upload (end_batch made it wait for the copy stream): while it
        // decodes batch k and the analysis kernels run, the copy stream is already uploading batch k + 1 into the other
        // buffer (begin_batch orders that upload behind everything the compute stream did with the buffer before).
        // Many frames: one warp each; few large ones (fewer frames than the GPU has warp slots to fill): a CTA each.
        const uint32_t *d_stream = (const uint32_t *)ctx->d_xtc[slot];
        const XtcFrameParams *d_params = (const XtcFrameParams *)ctx->d_xtc_params[slot];
        if (n_frames >= (size_t)kSMs * 4 || ctx->n_atoms < 4096) {
            const unsigned nb = (unsigned)((n_frames + kXtcWarpsPerCta - 1) / kXtcWarpsPerCta);
            k_xtc_decode<1><<<nb, kXtcWarpsPerCta * 32, 0, ctx->compute>>>(d_stream, d_params, (int)n_frames, ctx->cur_xyz, ctx->n_atoms * 3,
                                                                          ctx->d_xtc_status);
        } else {
            k_xtc_decode<kXtcWideWarps><<<(unsigned)n_frames, kXtcWideWarps * 32, 0, ctx->compute>>>(d_stream, d_params, (int)n_frames,
                                                                                                    ctx->cur_xyz, ctx->n_atoms * 3, ctx->d_xtc_status);
        }
        LAUNCHED();
        ctx->xtc_frames = n_frames;
        return GROAN_OK;
    }();
    if (rc) ctx->have_frames = false;
    return rc;
}

int groan_gpu_xtc_bad_frames(groan_gpu_ctx *ctx, size_t *n) {
    if (!ctx || !n) return GROAN_EINVAL;
    *n = 0;
    if (!ctx->d_xtc_status || ctx->xtc_frames == 0) return GROAN_OK;
    std::vector<int> h(ctx->xtc_frames);
    CK(cudaStreamSynchronize(ctx->copy));
    CK(cudaStreamSynchronize(ctx->compute));
    CK(cudaMemcpy(h.data(), ctx->d_xtc_status, h.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int v : h) *n += (v != 0);
    return GROAN_OK;
}

int groan_gpu_push_group_frames(groan_gpu_ctx *ctx, const float *xyz_sel, const uint32_t *atoms, size_t n_sel, const float *box,
                                size_t n_frames) {
    if (!ctx || !xyz_sel || !atoms || n_sel == 0 || n_sel > ctx->n_atoms) return GROAN_EINVAL;
    for (size_t k = 0; k < n_sel; k++) {
        if (atoms[k] >= ctx->n_atoms) return GROAN_EINVAL;
        if (k && atoms[k] <= atoms[k - 1]) return GROAN_EINVAL;
    }
    // atoms outside the selection keep the values the slot held two batches ago (the reference's System keeps the last
    // values it read, molly_xtc.rs:441-462); make sure the slot exists and holds defined bytes
    const bool fresh = ctx->d_slot[0] == nullptr;
    int rc = begin_batch(ctx, n_frames, box, true);
    if (rc) return rc;
    rc = [&]() -> int {
        ctx->attached = false;
        const int slot = ctx->slot;
        ctx->cur_xyz = ctx->d_slot[slot];
        if (fresh)
            for (int s = 0; s < 2; s++) CK(cudaMemsetAsync(ctx->d_slot[s], 0, ctx->max_frames * ctx->n_atoms * 3 * sizeof(float), ctx->copy));
        const size_t cap = ctx->max_frames * n_sel * 3 * sizeof(float);
        if (cap > ctx->sel_cap) {
            CK(cudaStreamSynchronize(ctx->copy));
            for (int s = 0; s < 2; s++) {
                if (ctx->d_sel[s]) cudaFree(ctx->d_sel[s]);
                ctx->d_sel[s] = nullptr;
                CK(cudaMalloc(&ctx->d_sel[s], cap));
            }
            ctx->sel_cap = cap;
        }
        if (n_sel > ctx->sel_atoms_cap) {
            CK(cudaStreamSynchronize(ctx->copy));
            if (ctx->d_sel_atoms) cudaFree(ctx->d_sel_atoms);
            ctx->d_sel_atoms = nullptr;
            CK(cudaMalloc(&ctx->d_sel_atoms, n_sel * sizeof(uint32_t)));
            ctx->sel_atoms_cap = n_sel;
        }
        CK(cudaMemcpyAsync(ctx->d_sel_atoms, atoms, n_sel * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copy));
        int r = h2d_on_copy_stream(ctx, ctx->d_sel[slot], xyz_sel, n_frames * n_sel * 3 * sizeof(float));
        if (r) return r;
        const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((n_sel + kThreads - 1) / kThreads, (size_t)kSMs * 8));
        k_scatter_group_frames<<<dim3(nb, (unsigned)n_frames), kThreads, 0, ctx->copy>>>((const float *)ctx->d_sel[slot], ctx->d_sel_atoms,
                                                                                            (uint32_t)n_sel, ctx->cur_xyz, ctx->n_atoms);
        LAUNCHED();
        return end_batch(ctx);
    }();
    if (rc) ctx->have_frames = false;
    return rc;
}

int groan_gpu_write_xtc(groan_gpu_ctx *ctx, float precision, const int32_t *step, const float *time, int n_threads, uint8_t *out,
                        size_t capacity, size_t *len) {
    if (!ctx || !out || !len) return GROAN_EINVAL;
    if (!ctx->have_frames) return GROAN_ENOFRAMES;
    if (!ctx->have_box) return GROAN_ENOBOX;
    if (!(precision > 0.0f)) precision = 1000.0f;
    const size_t F = ctx->n_frames, N = ctx->n_atoms;
    if (N <= 9) {
        std::vector<float> x(F * N * 3);
        int rc = groan_gpu_get_frames(ctx, x.data());
        if (rc) return rc;
        rc = groan_xtc_encode(x.data(), nullptr, F, N, ctx->h_box.data(), step, time, precision, n_threads, out, capacity, len);
        return rc == GROAN_XTC_OK ? GROAN_OK : (rc == GROAN_XTC_ECAPACITY ? GROAN_ECAPACITY : GROAN_EINVAL);
    }
    std::vector<int32_t> q(F * N * 3);
    int rc = groan_gpu_get_frames_quantized(ctx, q.data(), precision);  // the writer's rounding, on the device
    if (rc) return rc;
    rc = groan_xtc_encode(nullptr, q.data(), F, N, ctx->h_box.data(), step, time, precision, n_threads, out, capacity, len);
    return rc == GROAN_XTC_OK ? GROAN_OK : (rc == GROAN_XTC_ECAPACITY ? GROAN_ECAPACITY : GROAN_EINVAL);
}

}  // extern "C"

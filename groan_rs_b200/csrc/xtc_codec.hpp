// xtc_codec.hpp -- host side of the xtc coordinate codec.
//
// The reference reads and writes xtc through its vendored C library (external/xdrfile/xdrfile.c:742-949 decoder,
// :950-1330 encoder; frame header xdrfile_xtc.c) or through `molly` (src/io/xtc_io/molly_xtc.rs).  This is an independent
// implementation of the same FILE FORMAT, written for the way the GPU path consumes it:
//   * a frame is parsed into a FrameInfo (header fields + where the compressed bit stream lies) without touching the
//     coordinates, so that the stream can be handed to the GPU decoder (kernels_xtc.cuh) as it lies in the file;
//   * the host decoder emits the INTEGER lattice points (what groan_gpu_push_frames_quantized uploads) or the reader's
//     floats, for all atoms or for a sorted subset (GroupXtcReader semantics: decoding stops at the last wanted atom,
//     molly_xtc.rs:404-470); frames are independent, so a batch is decoded by a pool of threads;
//   * the encoder reproduces the reference's output byte for byte (tests compare with oracle/_ref/libxdrfile.so and with
//     the golden short_trajectory_fit.xtc, rmsd.rs:952-994).
//
// Format (all integers / floats big-endian, XDR):
//   magic 1995 | natoms | step | time | box[9] | natoms | [natoms <= 9: 3 natoms raw floats]
//   precision | minint[3] | maxint[3] | smallidx | nbytes | bit stream padded to a multiple of 4 bytes
// Bit stream, MSB first, per "group": one LARGE atom (three lattice integers minus minint, packed in mixed radix
// sizeint[] into `bitsize` bits, or three fixed-width fields when a size exceeds 24 bits), one flag bit, and when it is set
// 5 bits holding run + (is_smaller + 1): run / 3 SMALL atoms follow, each three integers in mixed radix magic[smallidx]
// packed into `smallidx` bits, each relative to the atom before it (+ magic[smallidx] / 2); the first small atom and the
// large atom swap places in the output (the water-molecule trick).  After the group smallidx moves by is_smaller.
// Mixed-radix fields are stored as a little-endian sequence of 8-bit chunks (the last chunk holds the remaining bits).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <climits>
#include <vector>

namespace groan_xtc {

constexpr int kFirstIdx = 9;
// the format's table of small-integer radices, ~2^(i/3) (part of the file format: every xtc reader carries these numbers)
static const int kMagicInts[] = {0,       0,       0,       0,       0,       0,       0,       0,        0,        8,       10,      12,      16,
                                 20,      25,      32,      40,      50,      64,      80,      101,      128,      161,     203,     256,     322,
                                 406,     512,     645,     812,     1024,    1290,    1625,    2048,     2580,     3250,    4096,    5060,    6501,
                                 8192,    10321,   13003,   16384,   20642,   26007,   32768,   41285,    52015,    65536,   82570,   104031,  131072,
                                 165140,  208063,  262144,  330280,  416127,  524287,  660561,  832255,   1048576,  1321122, 1664510, 2097152, 2642245,
                                 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int kLastIdx = (int)(sizeof(kMagicInts) / sizeof(kMagicInts[0]));
constexpr int kXtcMagic = 1995;

enum Status { XTC_OK = 0, XTC_EOF = 1, XTC_EMAGIC = 2, XTC_ETRUNC = 3, XTC_EFORMAT = 4, XTC_ERAW = 5, XTC_ECAPACITY = 6 };

typedef unsigned __int128 u128;

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
inline float be32f(const uint8_t *p) {
    const uint32_t u = be32(p);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline void put32(uint8_t *p, uint32_t v) {
    p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}
inline void put32f(uint8_t *p, float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    put32(p, u);
}
inline int bit_length(u128 v) {
    int n = 0;
    while (v) { n++; v >>= 1; }
    return n;
}

// everything of a frame but its coordinates
struct FrameInfo {
    uint64_t offset = 0;      // of the frame in the file
    uint64_t next = 0;        // offset of the following frame
    int32_t natoms = 0, step = 0;
    float time = 0.f, box[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool raw = false;         // natoms <= 9: uncompressed floats at `payload`
    float precision = 0.f;
    int32_t minint[3] = {0, 0, 0}, maxint[3] = {0, 0, 0};
    uint32_t sizeint[3] = {0, 0, 0};
    int32_t bitsize = 0;      // bits of a packed large atom; 0 = three separate fields of bitsizeint[] bits
    int32_t bitsizeint[3] = {0, 0, 0};
    int32_t smallidx = 0;
    uint64_t payload = 0;     // offset of the bit stream (or of the raw floats) in the file
    uint32_t nbytes = 0;      // length of the bit stream
};

// parse the header of the frame at `off`; XTC_EOF when off == len
inline int parse_frame(const uint8_t *d, uint64_t len, uint64_t off, FrameInfo *fi) {
    if (off == len) return XTC_EOF;
    if (off + 56 > len) return XTC_ETRUNC;
    const uint8_t *p = d + off;
    if ((int32_t)be32(p) != kXtcMagic) return XTC_EMAGIC;
    fi->offset = off;
    fi->natoms = (int32_t)be32(p + 4);
    fi->step = (int32_t)be32(p + 8);
    fi->time = be32f(p + 12);
    for (int k = 0; k < 9; k++) fi->box[k] = be32f(p + 16 + 4 * k);
    const int32_t lsize = (int32_t)be32(p + 52);
    if (fi->natoms < 0 || lsize != fi->natoms) return XTC_EFORMAT;
    uint64_t q = off + 56;
    if (fi->natoms <= 9) {
        fi->raw = true;
        fi->payload = q;
        fi->nbytes = (uint32_t)fi->natoms * 12u;
        if (q + fi->nbytes > len) return XTC_ETRUNC;
        fi->next = q + fi->nbytes;
        return XTC_OK;
    }
    if (q + 36 > len) return XTC_ETRUNC;
    fi->raw = false;
    fi->precision = be32f(d + q);
    for (int k = 0; k < 3; k++) {
        fi->minint[k] = (int32_t)be32(d + q + 4 + 4 * k);
        fi->maxint[k] = (int32_t)be32(d + q + 16 + 4 * k);
        fi->sizeint[k] = (uint32_t)(fi->maxint[k] - fi->minint[k] + 1);
    }
    fi->smallidx = (int32_t)be32(d + q + 28);
    fi->nbytes = be32(d + q + 32);
    fi->payload = q + 36;
    if (fi->smallidx < kFirstIdx || fi->smallidx >= kLastIdx) return XTC_EFORMAT;
    if ((fi->sizeint[0] | fi->sizeint[1] | fi->sizeint[2]) > 0xffffffu) {
        fi->bitsize = 0;
        for (int k = 0; k < 3; k++) fi->bitsizeint[k] = bit_length((u128)fi->sizeint[k]);
    } else {
        fi->bitsize = bit_length((u128)fi->sizeint[0] * fi->sizeint[1] * fi->sizeint[2]);
        fi->bitsizeint[0] = fi->bitsizeint[1] = fi->bitsizeint[2] = 0;
    }
    const uint64_t padded = ((uint64_t)fi->nbytes + 3) & ~(uint64_t)3;
    if (fi->payload + padded > len) return XTC_ETRUNC;
    fi->next = fi->payload + padded;
    return XTC_OK;
}

// MSB-first bit reader over a byte range; reads past the end return zero bits
struct BitReader {
    const uint8_t *p;
    uint64_t nbits, pos;
    BitReader(const uint8_t *data, uint64_t nbytes) : p(data), nbits(nbytes * 8), pos(0) {}
    inline uint32_t get(int n) {  // n <= 32
        if (n == 0) return 0;
        const uint64_t byte = pos >> 3;
        const int sh = (int)(pos & 7);
        uint64_t w = 0;
        const uint64_t avail = (nbits >> 3) > byte ? (nbits >> 3) - byte : 0;
        if (avail >= 8) {
            uint64_t t;
            std::memcpy(&t, p + byte, 8);
            w = __builtin_bswap64(t);
        } else {
            for (uint64_t k = 0; k < avail; k++) w |= (uint64_t)p[byte + k] << (56 - 8 * k);
        }
        pos += (uint64_t)n;
        return (uint32_t)((w << sh) >> (64 - n));
    }
};

// three integers packed in mixed radix sizes[] into nbits bits (little-endian 8-bit chunks)
inline void unpack3(BitReader &br, int nbits, const uint32_t sizes[3], int32_t out[3]) {
    if (nbits <= 64) {
        uint64_t v = 0;
        int sh = 0, nb = nbits;
        while (nb > 8) { v |= (uint64_t)br.get(8) << sh; sh += 8; nb -= 8; }
        if (nb > 0) v |= (uint64_t)br.get(nb) << sh;
        const uint64_t q2 = v / sizes[2];
        out[2] = (int32_t)(v - q2 * sizes[2]);
        const uint64_t q1 = q2 / sizes[1];
        out[1] = (int32_t)(q2 - q1 * sizes[1]);
        out[0] = (int32_t)(uint32_t)q1;
    } else {
        u128 v = 0;
        int sh = 0, nb = nbits;
        while (nb > 8) { v |= (u128)br.get(8) << sh; sh += 8; nb -= 8; }
        if (nb > 0) v |= (u128)br.get(nb) << sh;
        const u128 q2 = v / sizes[2];
        out[2] = (int32_t)(uint32_t)(v - q2 * sizes[2]);
        const u128 q1 = q2 / sizes[1];
        out[1] = (int32_t)(uint32_t)(q2 - q1 * sizes[1]);
        out[0] = (int32_t)(uint32_t)q1;
    }
}

// Decode the first n_decode atoms of a compressed frame; sink(i, x, y, z) receives atom i's lattice integers in file
// order of the OUTPUT (the water swap applied).  n_decode <= natoms: decoding may stop early (partial-frame reads).
template <typename Sink>
inline int decode_lattice(const uint8_t *file, const FrameInfo &fi, int32_t n_decode, Sink &&sink) {
    if (fi.raw) return XTC_ERAW;
    BitReader br(file + fi.payload, fi.nbytes);
    int smallidx = fi.smallidx;
    int smallnum = kMagicInts[smallidx] / 2;
    int smaller = kMagicInts[smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx] / 2;
    uint32_t ssize[3] = {(uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx]};
    int run = 0;
    int32_t i = 0;
    const int32_t n = n_decode < fi.natoms ? n_decode : fi.natoms;
    while (i < n) {
        int32_t big[3];
        if (fi.bitsize == 0) {
            for (int k = 0; k < 3; k++) big[k] = (int32_t)br.get(fi.bitsizeint[k]);
        } else {
            unpack3(br, fi.bitsize, fi.sizeint, big);
        }
        for (int k = 0; k < 3; k++) big[k] += fi.minint[k];
        int is_smaller = 0;
        if (br.get(1)) {
            run = (int)br.get(5);
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller--;
        }
        if (br.pos > br.nbits + 64) return XTC_ETRUNC;
        if (run > 0) {
            int32_t prev[3] = {big[0], big[1], big[2]};
            for (int k = 0; k < run; k += 3) {
                int32_t s[3];
                unpack3(br, smallidx, ssize, s);
                for (int c = 0; c < 3; c++) s[c] += prev[c] - smallnum;
                if (i < n) sink(i, s[0], s[1], s[2]);
                i++;
                if (k == 0) {  // the first small atom comes out BEFORE the large one
                    if (i < n) sink(i, big[0], big[1], big[2]);
                    i++;
                }
                prev[0] = s[0]; prev[1] = s[1]; prev[2] = s[2];
            }
        } else {
            sink(i, big[0], big[1], big[2]);
            i++;
        }
        smallidx += is_smaller;
        if (smallidx < kFirstIdx || smallidx >= kLastIdx) return XTC_EFORMAT;
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > kFirstIdx ? kMagicInts[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = kMagicInts[smallidx] / 2;
        }
        ssize[0] = ssize[1] = ssize[2] = (uint32_t)kMagicInts[smallidx];
    }
    return XTC_OK;
}

// the float the reference's readers hand out for a lattice integer (xdrfile.c:844,915-917): int * (float)(1.0 / precision)
inline float inv_precision(float precision) { return (float)(1.0 / (double)precision); }

// ------------------------------------------------------------------------------------------------ encoder
struct BitWriter {
    std::vector<uint8_t> &out;
    uint64_t acc = 0;  // pending bits, right-aligned
    int nacc = 0;
    explicit BitWriter(std::vector<uint8_t> &o) : out(o) {}
    inline void put(int n, uint32_t v) {  // n <= 32, v < 2^n
        if (n == 0) return;
        acc = (acc << n) | (uint64_t)v;
        nacc += n;
        while (nacc >= 8) {
            out.push_back((uint8_t)(acc >> (nacc - 8)));
            nacc -= 8;
        }
        acc &= ((uint64_t)1 << nacc) - 1;
    }
    inline void flush() {
        if (nacc > 0) {
            out.push_back((uint8_t)(acc << (8 - nacc)));
            nacc = 0;
            acc = 0;
        }
    }
};

// three integers (each < sizes[k]) in mixed radix into nbits bits, little-endian 8-bit chunks, high bits zero-filled
inline void pack3(BitWriter &bw, int nbits, const uint32_t sizes[3], const uint32_t v3[3]) {
    u128 v = ((u128)v3[0] * sizes[1] + v3[1]) * sizes[2] + v3[2];
    int nb = nbits;
    // the reference emits the significant bytes of the value first and pads with a zero field; as a bit string that is the
    // same as emitting 8-bit chunks from the low end until fewer than 8 (or exactly the remaining) bits are left
    int nbytes = 0;
    for (u128 t = v; t; t >>= 8) nbytes++;
    if (nbytes == 0) nbytes = 1;
    if (nb >= nbytes * 8) {
        for (int k = 0; k < nbytes; k++) { bw.put(8, (uint32_t)(v & 0xff)); v >>= 8; }
        int rest = nb - nbytes * 8;
        while (rest > 0) { const int c = rest > 32 ? 32 : rest; bw.put(c, 0); rest -= c; }
    } else {
        for (int k = 0; k < nbytes - 1; k++) { bw.put(8, (uint32_t)(v & 0xff)); v >>= 8; }
        bw.put(nb - (nbytes - 1) * 8, (uint32_t)v);
    }
}

// lattice point of a coordinate as the reference's writer computes it (xdrfile.c:1018-1031): f32 product, +-0.5, truncation
inline int32_t to_lattice(float x, float precision) {
    float lf;
    if (x >= 0.0f) lf = (float)((double)(x * precision) + 0.5);
    else lf = (float)((double)(x * precision) - 0.5);
    return (int32_t)lf;
}

// Encode one frame from its lattice integers q (natoms x 3, clobbered: the water swap happens in place like in the
// reference) and append it to `out`.  natoms > 9.
inline void encode_frame_lattice(int32_t *q, int32_t natoms, int32_t step, float time, const float box[9], float precision,
                                 std::vector<uint8_t> &out) {
    const size_t base = out.size();
    out.resize(base + 56 + 36);
    uint8_t *h = out.data() + base;
    put32(h, (uint32_t)kXtcMagic);
    put32(h + 4, (uint32_t)natoms);
    put32(h + 8, (uint32_t)step);
    put32f(h + 12, time);
    for (int k = 0; k < 9; k++) put32f(h + 16 + 4 * k, box[k]);
    put32(h + 52, (uint32_t)natoms);
    put32f(h + 56, precision);
    int32_t mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    int mindiff = INT_MAX;
    for (int32_t i = 0; i < natoms; i++) {
        for (int k = 0; k < 3; k++) {
            const int32_t v = q[3 * i + k];
            if (v < mn[k]) mn[k] = v;
            if (v > mx[k]) mx[k] = v;
        }
        if (i > 0) {
            const int diff = std::abs(q[3 * i - 3] - q[3 * i]) + std::abs(q[3 * i - 2] - q[3 * i + 1]) + std::abs(q[3 * i - 1] - q[3 * i + 2]);
            if (diff < mindiff) mindiff = diff;
        }
    }
    uint32_t sizeint[3];
    for (int k = 0; k < 3; k++) {
        put32(h + 60 + 4 * k, (uint32_t)mn[k]);
        put32(h + 72 + 4 * k, (uint32_t)mx[k]);
        sizeint[k] = (uint32_t)(mx[k] - mn[k] + 1);
    }
    int bitsize, bitsizeint[3] = {0, 0, 0};
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
        bitsize = 0;
        for (int k = 0; k < 3; k++) bitsizeint[k] = bit_length((u128)sizeint[k]);
    } else {
        bitsize = bit_length((u128)sizeint[0] * sizeint[1] * sizeint[2]);
    }
    int smallidx = kFirstIdx;
    // (the reference lets smallidx reach kLastIdx when no two consecutive atoms are within 16 777 216 lattice steps of each
    // other and then reads past its table, xdrfile.c:1104-1118; such a frame has no small atoms, the last valid index serves)
    while (smallidx < kLastIdx - 1 && kMagicInts[smallidx] < mindiff) smallidx++;
    put32(h + 84, (uint32_t)smallidx);
    const int maxidx = kLastIdx < smallidx + 8 ? kLastIdx : smallidx + 8;
    const int minidx = maxidx - 8;
    int smaller = kMagicInts[kFirstIdx > smallidx - 1 ? kFirstIdx : smallidx - 1] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    uint32_t ssize[3] = {(uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx], (uint32_t)kMagicInts[smallidx]};
    const int larger = kMagicInts[maxidx] / 2;
    std::vector<uint8_t> bits;
    bits.reserve((size_t)natoms * 6 + 64);
    BitWriter bw(bits);
    int32_t prev[3] = {0, 0, 0};
    int prevrun = -1;
    int32_t i = 0;
    while (i < natoms) {
        int32_t *a = q + 3 * (size_t)i;
        int is_smaller;
        if (smallidx < maxidx && i >= 1 && std::abs(a[0] - prev[0]) < larger && std::abs(a[1] - prev[1]) < larger &&
            std::abs(a[2] - prev[2]) < larger)
            is_smaller = 1;
        else if (smallidx > minidx)
            is_smaller = -1;
        else
            is_smaller = 0;
        bool is_small = false;
        if (i + 1 < natoms && std::abs(a[0] - a[3]) < smallnum && std::abs(a[1] - a[4]) < smallnum && std::abs(a[2] - a[5]) < smallnum) {
            for (int k = 0; k < 3; k++) { const int32_t t = a[k]; a[k] = a[3 + k]; a[3 + k] = t; }  // water swap
            is_small = true;
        }
        const uint32_t big[3] = {(uint32_t)(a[0] - mn[0]), (uint32_t)(a[1] - mn[1]), (uint32_t)(a[2] - mn[2])};
        if (bitsize == 0) {
            for (int k = 0; k < 3; k++) bw.put(bitsizeint[k], big[k]);
        } else {
            pack3(bw, bitsize, sizeint, big);
        }
        prev[0] = a[0]; prev[1] = a[1]; prev[2] = a[2];
        a += 3;
        i++;
        int run = 0;
        uint32_t small[24];
        if (!is_small && is_smaller == -1) is_smaller = 0;
        while (is_small && run < 24) {
            long long d2 = 0;
            for (int k = 0; k < 3; k++) {
                const int t = a[k] - prev[k];
                d2 += (long long)t * t;
            }
            // the reference compares int products (tmpsum >= smaller * smaller); both sides are far below 2^31 here
            if (is_smaller == -1 && d2 >= (long long)smaller * smaller) is_smaller = 0;
            for (int k = 0; k < 3; k++) small[run++] = (uint32_t)(a[k] - prev[k] + smallnum);
            prev[0] = a[0]; prev[1] = a[1]; prev[2] = a[2];
            i++;
            a += 3;
            is_small = i < natoms && std::abs(a[0] - prev[0]) < smallnum && std::abs(a[1] - prev[1]) < smallnum &&
                       std::abs(a[2] - prev[2]) < smallnum;
        }
        if (run != prevrun || is_smaller != 0) {
            prevrun = run;
            bw.put(1, 1);
            bw.put(5, (uint32_t)(run + is_smaller + 1));
        } else {
            bw.put(1, 0);
        }
        for (int k = 0; k < run; k += 3) pack3(bw, smallidx, ssize, small + k);
        if (is_smaller != 0) {
            smallidx += is_smaller;
            if (is_smaller < 0) {
                smallnum = smaller;
                smaller = kMagicInts[smallidx - 1] / 2;
            } else {
                smaller = smallnum;
                smallnum = kMagicInts[smallidx] / 2;
            }
            ssize[0] = ssize[1] = ssize[2] = (uint32_t)kMagicInts[smallidx];
        }
    }
    bw.flush();
    const uint32_t nbytes = (uint32_t)bits.size();
    while (bits.size() & 3) bits.push_back(0);
    put32(out.data() + base + 88, nbytes);
    out.insert(out.end(), bits.begin(), bits.end());
}

// frames of <= 9 atoms are stored as raw floats
inline void encode_frame_raw(const float *xyz, int32_t natoms, int32_t step, float time, const float box[9], std::vector<uint8_t> &out) {
    const size_t base = out.size();
    out.resize(base + 56 + (size_t)natoms * 12);
    uint8_t *h = out.data() + base;
    put32(h, (uint32_t)kXtcMagic);
    put32(h + 4, (uint32_t)natoms);
    put32(h + 8, (uint32_t)step);
    put32f(h + 12, time);
    for (int k = 0; k < 9; k++) put32f(h + 16 + 4 * k, box[k]);
    put32(h + 52, (uint32_t)natoms);
    for (int32_t k = 0; k < natoms * 3; k++) put32f(h + 56 + 4 * k, xyz[k]);
}

}  // namespace groan_xtc

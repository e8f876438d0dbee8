// hb_xtc_host.h - host side of the XTC trajectory path: frame index, a frame writer (test / benchmark inputs) and a
// frame reader for the few frames the host itself needs (frame 0 for the constructor's bonded-hydrogen detection).
// The frames of a run are NOT decoded here: their compressed records cross PCIe as they lie in the file and are
// decoded on the device (hb_xtc.cuh).
//
// Format (GROMACS .xtc as read by MDAnalysis' bundled xdrfile for the reference, basic.py:46-47): XDR big-endian
// record {int 1995, int natoms, int step, float time, float box[9]} + coordinate block {int natoms, float precision,
// int minint[3], int maxint[3], int smallidx, int nbytes, bit stream}. A triple of integers (a, b, c) with ranges
// (sa, sb, sc) is stored as v = (a * sb + b) * sc + c in `nbits` bits: the bytes of v least significant first, eight
// bits each, the last chunk taking the remaining 1..8 bits; bits go into the stream most significant first.
// This implementation works on v as one 128-bit integer (the published library walks byte arrays).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

namespace xtc {

constexpr int kMagic = 1995;
constexpr int kFirstIdx = 9;
constexpr int kHeaderBytes = 52;            // magic .. box
constexpr int kCoordHeaderBytes = 40;       // natoms, precision, minint, maxint, smallidx, nbytes
static const int kMagicInts[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
    65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
    2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
constexpr int kLastIdx = 72;

typedef unsigned __int128 u128;

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline void put32(unsigned char* p, uint32_t v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; }
inline float bef(const unsigned char* p) { uint32_t v = be32(p); float f; memcpy(&f, &v, 4); return f; }
inline void putf(unsigned char* p, float f) { uint32_t v; memcpy(&v, &f, 4); put32(p, v); }

inline int bit_length(u128 v) { int n = 0; while (v) { ++n; v >>= 1; } return n; }
inline int bits_of_int(unsigned size) { int n = 0; unsigned long long num = 1; while (size >= num && n < 32) { ++n; num <<= 1; } return n; }

// length of the record starting at p (0 = not a frame / truncated)
inline size_t frame_bytes(const unsigned char* p, size_t avail, int* natoms) {
    if (avail < (size_t)kHeaderBytes + 4 || be32(p) != (uint32_t)kMagic) return 0;
    const int n = (int)be32(p + 4);
    if (n < 0) return 0;
    if (natoms) *natoms = n;
    size_t len;
    if (n <= 9) len = (size_t)kHeaderBytes + 4 + (size_t)12 * n;
    else {
        if (avail < (size_t)kHeaderBytes + kCoordHeaderBytes) return 0;
        const size_t nb = be32(p + kHeaderBytes + 36);
        len = (size_t)kHeaderBytes + kCoordHeaderBytes + ((nb + 3) & ~(size_t)3);
    }
    return len <= avail ? len : 0;
}

// ---- writer ---------------------------------------------------------------------------------------------------
struct BitWriter {
    std::vector<unsigned char>& out;
    unsigned long long acc = 0;     // pending bits in the low `n` bits
    int n = 0;
    explicit BitWriter(std::vector<unsigned char>& o) : out(o) {}
    void put(unsigned v, int bits) {                 // bits <= 32, most significant bit first
        if (bits <= 0) return;
        acc = (acc << bits) | (bits == 32 ? (unsigned long long)v : (unsigned long long)(v & ((1u << bits) - 1u)));
        n += bits;
        while (n >= 8) { out.push_back((unsigned char)(acc >> (n - 8))); n -= 8; }
    }
    void put_triple(u128 v, int nbits) {             // bytes of v least significant first, last chunk 1..8 bits
        int left = nbits;
        while (left > 8) { put((unsigned)(v & 0xff), 8); v >>= 8; left -= 8; }
        if (left > 0) put((unsigned)(v & 0xff), left);
    }
    void flush() { if (n > 0) { out.push_back((unsigned char)(acc << (8 - n))); n = 0; } }
};

// Encode one frame; xyz are multiplied by `in_scale` first (0.1: Angstrom -> nm; float32). box9 in nm. Appends the
// record to `out`; false if a coordinate is out of range.
inline bool encode_frame(const float* xyz, int natoms, float in_scale, float precision, int step, float time,
                         const float* box9, std::vector<unsigned char>& out) {
    const size_t start = out.size();
    out.resize(start + kHeaderBytes + 4);
    unsigned char* h = out.data() + start;
    put32(h, kMagic); put32(h + 4, (uint32_t)natoms); put32(h + 8, (uint32_t)step); putf(h + 12, time);
    for (int k = 0; k < 9; ++k) putf(h + 16 + 4 * k, box9 ? box9[k] : 0.f);
    put32(h + kHeaderBytes, (uint32_t)natoms);
    auto coord = [&](size_t idx) { return in_scale == 1.0f ? xyz[idx] : xyz[idx] * in_scale; };
    if (natoms <= 9) {
        for (int k = 0; k < 3 * natoms; ++k) { unsigned char b[4]; putf(b, coord((size_t)k)); out.insert(out.end(), b, b + 4); }
        return true;
    }
    if (!(precision > 0.f)) precision = 1000.f;
    std::vector<int> q((size_t)natoms * 3);
    int lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    long long mindiff = INT32_MAX;
    const float maxabs = (float)(INT32_MAX - 2);
    for (int i = 0; i < natoms; ++i) {
        long long d = 0;
        for (int k = 0; k < 3; ++k) {
            const float x = coord((size_t)3 * i + k);
            const float lf = x >= 0.f ? x * precision + 0.5f : x * precision - 0.5f;
            if (!(std::fabs(lf) <= maxabs)) { out.resize(start); return false; }
            const int v = (int)lf;
            q[(size_t)3 * i + k] = v;
            lo[k] = v < lo[k] ? v : lo[k];
            hi[k] = v > hi[k] ? v : hi[k];
            if (i > 0) d += std::llabs((long long)q[(size_t)3 * (i - 1) + k] - v);
        }
        if (i > 0 && d < mindiff) mindiff = d;
    }
    unsigned size[3];
    for (int k = 0; k < 3; ++k) {
        if ((float)hi[k] - (float)lo[k] >= maxabs) { out.resize(start); return false; }
        size[k] = (unsigned)(hi[k] - lo[k] + 1);
    }
    const bool wide = (size[0] | size[1] | size[2]) > 0xffffffu;
    int wbits[3] = {0, 0, 0}, bitsize = 0;
    if (wide) for (int k = 0; k < 3; ++k) wbits[k] = bits_of_int(size[k]);
    else bitsize = bit_length((u128)size[0] * size[1] * size[2]);
    int smallidx = kFirstIdx;
    while (smallidx < kLastIdx && kMagicInts[smallidx] < mindiff) ++smallidx;
    {
        unsigned char b[kCoordHeaderBytes - 4];
        putf(b, precision);
        for (int k = 0; k < 3; ++k) { put32(b + 4 + 4 * k, (uint32_t)lo[k]); put32(b + 16 + 4 * k, (uint32_t)hi[k]); }
        put32(b + 28, (uint32_t)smallidx);
        put32(b + 32, 0);                              // nbytes, patched below
        out.insert(out.end(), b, b + sizeof(b));
    }
    const size_t nbytes_at = out.size() - 4, data_at = out.size();
    const int maxidx = smallidx + 8 < kLastIdx ? smallidx + 8 : kLastIdx, minidx = maxidx - 8;
    int smaller = kMagicInts[smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    unsigned ssize = (unsigned)kMagicInts[smallidx];
    const int larger = kMagicInts[maxidx] / 2;
    BitWriter bw(out);
    auto near = [&](const int* a, const int* b, int lim) {
        return std::abs(a[0] - b[0]) < lim && std::abs(a[1] - b[1]) < lim && std::abs(a[2] - b[2]) < lim;
    };
    int prev[3] = {0, 0, 0}, prevrun = -1;
    int i = 0;
    while (i < natoms) {
        int* cur = &q[(size_t)3 * i];
        int change = 0;                                 // of smallidx after this group: +1, 0, -1
        if (smallidx < maxidx && i >= 1 && near(cur, prev, larger)) change = 1;
        else if (smallidx > minidx) change = -1;
        bool small = false;
        if (i + 1 < natoms && near(cur, cur + 3, smallnum)) {
            for (int k = 0; k < 3; ++k) { const int t = cur[k]; cur[k] = cur[k + 3]; cur[k + 3] = t; }   // swap the first two
            small = true;
        }
        if (wide) for (int k = 0; k < 3; ++k) bw.put((unsigned)(cur[k] - lo[k]), wbits[k]);
        else bw.put_triple(((u128)(unsigned)(cur[0] - lo[0]) * size[1] + (unsigned)(cur[1] - lo[1])) * size[2] + (unsigned)(cur[2] - lo[2]), bitsize);
        prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];
        ++i;
        if (!small && change == -1) change = 0;
        unsigned delta[24];
        int run = 0;
        while (small && run < 24) {
            const int* nx = &q[(size_t)3 * i];
            if (change == -1) {
                const long long a = nx[0] - prev[0], b = nx[1] - prev[1], c = nx[2] - prev[2];
                if (a * a + b * b + c * c >= (long long)smaller * smaller) change = 0;
            }
            for (int k = 0; k < 3; ++k) delta[run++] = (unsigned)(nx[k] - prev[k] + smallnum);
            prev[0] = nx[0]; prev[1] = nx[1]; prev[2] = nx[2];
            ++i;
            small = i < natoms && near(&q[(size_t)3 * i], prev, smallnum);
        }
        if (run != prevrun || change != 0) {
            prevrun = run;
            bw.put(1u, 1);
            bw.put((unsigned)(run + change + 1), 5);
        } else {
            bw.put(0u, 1);
        }
        for (int k = 0; k < run; k += 3) bw.put_triple(((u128)delta[k] * ssize + delta[k + 1]) * ssize + delta[k + 2], smallidx);
        if (change != 0) {
            smallidx += change;
            if (change < 0) { smallnum = smaller; smaller = kMagicInts[smallidx - 1] / 2; }
            else { smaller = smallnum; smallnum = kMagicInts[smallidx] / 2; }
            ssize = (unsigned)kMagicInts[smallidx];
        }
    }
    bw.flush();
    put32(out.data() + nbytes_at, (uint32_t)(out.size() - data_at));
    while ((out.size() - start) & 3) out.push_back(0);
    return true;
}

// ---- reader (host; single frames) -------------------------------------------------------------------------------
struct BitReader {
    const unsigned char* p;
    size_t nbits_total;
    size_t pos = 0;
    bool bad = false;
    unsigned long long get(int bits) {                 // bits <= 64, most significant first
        unsigned long long v = 0;
        if (pos + (size_t)bits > nbits_total) { bad = true; pos += (size_t)bits; return 0; }
        for (int k = 0; k < bits; ++k, ++pos) v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1u);
        return v;
    }
    u128 get_triple(int nbits) {
        u128 v = 0;
        int shift = 0, left = nbits;
        while (left > 8) { v |= (u128)get(8) << shift; shift += 8; left -= 8; }
        if (left > 0) v |= (u128)get(left) << shift;
        return v;
    }
};

// Decode the frame record at p: xyz_out[k] = fl32(fl32(int * fl32(1 / precision)) * scale) (scale = 10: Angstrom, the
// float32 in-place unit conversion MDAnalysis applies), box9 (nm, unscaled). Returns the record length, 0 on error.
inline size_t decode_frame(const unsigned char* p, size_t avail, int natoms_expected, float scale, float* xyz_out, float* box9,
                           int* step, float* time) {
    int natoms = 0;
    const size_t len = frame_bytes(p, avail, &natoms);
    if (!len || natoms != natoms_expected || (int)be32(p + kHeaderBytes) != natoms) return 0;
    if (step) *step = (int)be32(p + 8);
    if (time) *time = bef(p + 12);
    if (box9) for (int k = 0; k < 9; ++k) box9[k] = bef(p + 16 + 4 * k);
    const unsigned char* c = p + kHeaderBytes + 4;
    if (natoms <= 9) {
        for (int k = 0; k < 3 * natoms; ++k) xyz_out[k] = bef(c + 4 * k) * scale;
        return len;
    }
    const float precision = bef(c);
    int lo[3], hi[3];
    unsigned size[3];
    for (int k = 0; k < 3; ++k) { lo[k] = (int)be32(c + 4 + 4 * k); hi[k] = (int)be32(c + 16 + 4 * k); size[k] = (unsigned)(hi[k] - lo[k] + 1); }
    int smallidx = (int)be32(c + 28);
    const size_t nbytes = be32(c + 32);
    if (smallidx < kFirstIdx || smallidx > kLastIdx) return 0;
    const bool wide = (size[0] | size[1] | size[2]) > 0xffffffu;
    int wbits[3] = {0, 0, 0}, bitsize = 0;
    if (wide) for (int k = 0; k < 3; ++k) wbits[k] = bits_of_int(size[k]);
    else bitsize = bit_length((u128)size[0] * size[1] * size[2]);
    int smaller = kMagicInts[smallidx - 1 > kFirstIdx ? smallidx - 1 : kFirstIdx] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    unsigned ssize = (unsigned)kMagicInts[smallidx];
    const float inv = (float)(1.0 / (double)precision);
    BitReader br{c + 36, nbytes * 8};
    float* o = xyz_out;
    auto emit = [&](const int* v) { for (int k = 0; k < 3; ++k) *o++ = ((float)v[k] * inv) * scale; };
    int i = 0, run = 0;
    while (i < natoms && !br.bad) {
        int big[3];
        if (wide) for (int k = 0; k < 3; ++k) big[k] = (int)br.get(wbits[k]) + lo[k];
        else {
            u128 v = br.get_triple(bitsize);
            big[2] = (int)(v % size[2]) + lo[2]; v /= size[2];
            big[1] = (int)(v % size[1]) + lo[1]; v /= size[1];
            big[0] = (int)v + lo[0];
        }
        ++i;
        int change = 0;
        if (br.get(1)) {
            const int r = (int)br.get(5);
            change = r % 3;
            run = r - change;
            change -= 1;
        }
        if (run > 0) {
            if (i + run / 3 > natoms) return 0;
            int prev[3] = {big[0], big[1], big[2]};
            for (int k = 0; k < run; k += 3) {
                u128 v = br.get_triple(smallidx);
                int s[3];
                s[2] = (int)(v % ssize); v /= ssize;
                s[1] = (int)(v % ssize); v /= ssize;
                s[0] = (int)v;
                for (int d = 0; d < 3; ++d) s[d] += prev[d] - smallnum;
                emit(s);
                if (k == 0) emit(big);                 // stored swapped: the large atom follows its first small one
                prev[0] = s[0]; prev[1] = s[1]; prev[2] = s[2];
                ++i;
            }
        } else {
            emit(big);
        }
        smallidx += change;
        if (smallidx < kFirstIdx || smallidx > kLastIdx) return 0;
        if (change < 0) { smallnum = smaller; smaller = smallidx > kFirstIdx ? kMagicInts[smallidx - 1] / 2 : 0; }
        else if (change > 0) { smaller = smallnum; smallnum = kMagicInts[smallidx] / 2; }
        ssize = (unsigned)kMagicInts[smallidx];
    }
    return br.bad ? 0 : len;
}

}  // namespace xtc

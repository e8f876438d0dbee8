// hb_xtc.cuh - XTC frame records decoded on the device (frame sources; basic.py:46-47 reads any
// MDAnalysis format, for .xtc MDAnalysis decodes every frame on the CPU with xdrfile).
//
// The compressed records cross PCIe as they lie in the file (about 3.5 bytes per atom instead of 12) and one WARP
// decodes one frame into the float32 [atom][3] layout the search kernels read:
//
//   The bit stream is a sequence of groups {large atom: `bitsize` bits, flag bit, [5 bits: run + smallidx change],
//   run / 3 small atoms of `smallidx` bits each}. Where a group starts depends on all groups before it, but as long
//   as the flag bit is 0 nothing changes, so the warp speculates: lane k assumes the next k groups have no flag
//   and looks at "its" flag bit; the lanes up to and including the first one that finds a flag are right, all
//   decode their group in parallel, and the flagged lane's new run / smallidx become the warp's state. In water
//   (O, H, H: run = 6 for tens of thousands of groups) one step decodes 32 molecules.
//   A triple (a, b, c) with ranges (sa, sb, sc) is v = (a * sb + b) * sc + c, stored as the bytes of v least
//   significant first (the last chunk holds the remaining 1..8 bits), bits most significant first: read as a
//   64-bit big-endian window, byte-swapped, and divided as one integer (128-bit only when nbits > 64).
//   Decoded atoms are staged in shared memory and stored as full warp-wide float rows.
//
// Arithmetic of the published decoder (xdrfile_decompress_coord_float) and of MDAnalysis' unit conversion:
//   x = fl32(fl32(float(int) * fl32(1 / precision)) * scale), scale = 10 (nm -> Angstrom), every step rounded.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace {

#define XTC_WARPS 4
#define XTC_STAGE_ATOMS (32 * 9)

__constant__ int c_xtc_magic[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
    65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
    2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};

struct XtcRanges {            // atom ranges [b, e) whose coordinates are written (n = 0: all atoms)
    int n;
    int b[8], e[8];
};

struct XtcArgs {
    const unsigned char* raw;       // records of the batch, 4-byte aligned, 16 readable bytes past the end
    const long long* frame_off;     // [n_frames + 1] byte offsets of the records inside raw
    int n_frames;
    int n_atoms;
    float scale;
    float* out;                     // [n_frames][n_atoms][3]
    int* status;                    // set to 1 when a record is malformed
    XtcRanges ranges;
};

__device__ __forceinline__ unsigned xtc_be32(const unsigned char* p) {       // p is 4-byte aligned
    return __byte_perm(*reinterpret_cast<const unsigned*>(p), 0, 0x0123);
}
// `n` (1..64) bits of the big-endian bit stream starting at bit `pos` of `data` (4-byte aligned)
__device__ __forceinline__ unsigned long long xtc_bits(const unsigned char* data, unsigned long long pos, int n) {
    const unsigned* w = reinterpret_cast<const unsigned*>(data) + (pos >> 5);
    const unsigned a = __byte_perm(__ldg(w), 0, 0x0123), b = __byte_perm(__ldg(w + 1), 0, 0x0123), c = __byte_perm(__ldg(w + 2), 0, 0x0123);
    const int s = (int)(pos & 31);
    unsigned long long hi = ((unsigned long long)a << 32) | b;
    if (s) hi = (hi << s) | (c >> (32 - s));
    return hi >> (64 - n);
}
__device__ __forceinline__ unsigned long long xtc_bswap64(unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
// the integer v of a triple stored in `nbits` (<= 64) bits at `pos`
__device__ __forceinline__ unsigned long long xtc_triple64(const unsigned char* data, unsigned long long pos, int nbits) {
    const int k = (nbits - 1) >> 3, r = nbits - 8 * k;              // k full bytes, then r = 1..8 bits
    const unsigned long long w = xtc_bits(data, pos, nbits);
    const unsigned long long last = w & ((1ull << r) - 1ull);
    unsigned long long v = 0;
    if (k) v = xtc_bswap64((w >> r) << (64 - 8 * k));             // chunk 0 (first in the stream) is the lowest byte
    return v | (last << (8 * k));                                   // k <= 7 here
}

__device__ __forceinline__ void xtc_split(unsigned long long v, unsigned sb, unsigned sc, int nbits, int& a, int& b, int& c) {
    if (nbits <= 32) {
        const unsigned u = (unsigned)v, q = u / sc;
        c = (int)(u - q * sc);
        const unsigned q2 = q / sb;
        b = (int)(q - q2 * sb);
        a = (int)q2;
    } else {
        const unsigned long long q = v / sc;
        c = (int)(v - q * sc);
        const unsigned long long q2 = q / sb;
        b = (int)(q - q2 * sb);
        a = (int)q2;
    }
}
// general triple (nbits up to 72)
__device__ __noinline__ void xtc_triple_wide(const unsigned char* data, unsigned long long pos, int nbits, unsigned sb, unsigned sc,
                                             int& a, int& b, int& c) {
    unsigned __int128 v = 0;
    int shift = 0, left = nbits;
    while (left > 8) { v |= (unsigned __int128)xtc_bits(data, pos, 8) << shift; pos += 8; shift += 8; left -= 8; }
    v |= (unsigned __int128)xtc_bits(data, pos, left) << shift;
    c = (int)(unsigned)(v % sc); v /= sc;
    b = (int)(unsigned)(v % sb); v /= sb;
    a = (int)(unsigned)v;
}
__device__ __forceinline__ void xtc_triple(const unsigned char* data, unsigned long long pos, int nbits, unsigned sb, unsigned sc,
                                           int& a, int& b, int& c) {
    if (nbits <= 64) xtc_split(xtc_triple64(data, pos, nbits), sb, sc, nbits, a, b, c);
    else xtc_triple_wide(data, pos, nbits, sb, sc, a, b, c);
}
__device__ __forceinline__ int xtc_bit_length(unsigned __int128 v) {
    const unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
    return hi ? 128 - __clzll((long long)hi) : (lo ? 64 - __clzll((long long)lo) : 0);
}

__global__ void __launch_bounds__(XTC_WARPS * 32) k_xtc_decode(XtcArgs A) {
    __shared__ float stage[XTC_WARPS][XTC_STAGE_ATOMS * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * XTC_WARPS + warp;
    if (f >= A.n_frames) return;
    const unsigned char* rec = A.raw + A.frame_off[f];
    const long long rec_len = A.frame_off[f + 1] - A.frame_off[f];
    float* out = A.out + (size_t)f * A.n_atoms * 3;
    float* st = stage[warp];
    const int natoms = A.n_atoms;
    auto fail = [&]() { if (lane == 0) atomicExch(A.status, 1); };
    if (rec_len < 56 || xtc_be32(rec) != 1995u || (int)xtc_be32(rec + 4) != natoms || (int)xtc_be32(rec + 52) != natoms) { fail(); return; }
    auto wanted = [&](int atom) {
        if (A.ranges.n == 0) return true;
        bool w = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) w = w || (k < A.ranges.n && atom >= A.ranges.b[k] && atom < A.ranges.e[k]);
        return w;
    };
    if (natoms <= 9) {                                   // stored as plain floats
        if (rec_len < 56 + 12ll * natoms) { fail(); return; }
        for (int k = lane; k < 3 * natoms; k += 32) out[k] = __fmul_rn(__uint_as_float(xtc_be32(rec + 56 + 4 * k)), A.scale);
        return;
    }
    if (rec_len < 92) { fail(); return; }
    const float precision = __uint_as_float(xtc_be32(rec + 56));
    int lo[3];
    unsigned size[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = (int)xtc_be32(rec + 60 + 4 * k);
        size[k] = (unsigned)((int)xtc_be32(rec + 72 + 4 * k) - lo[k] + 1);
    }
    int smallidx = (int)xtc_be32(rec + 84);
    const unsigned long long nbits_total = 8ull * xtc_be32(rec + 88);
    if (smallidx < 9 || smallidx > 72 || 92ull + nbits_total / 8 > (unsigned long long)rec_len) { fail(); return; }
    const unsigned char* data = rec + 92;
    const bool wide = (size[0] | size[1] | size[2]) > 0xffffffu;
    int wb[3] = {0, 0, 0}, bitsize;
    if (wide) {
#pragma unroll
        for (int k = 0; k < 3; ++k) wb[k] = size[k] ? 32 - __clz(size[k]) : 0;      // bits of an int of this range
        bitsize = wb[0] + wb[1] + wb[2];
    } else {
        bitsize = xtc_bit_length((unsigned __int128)size[0] * size[1] * size[2]);
    }
    const float inv = (float)(1.0 / (double)precision);
    const float scale = A.scale;
    int smaller = c_xtc_magic[max(smallidx - 1, 9)] / 2;
    int smallnum = c_xtc_magic[smallidx] / 2;
    unsigned ssize = (unsigned)c_xtc_magic[smallidx];

    unsigned long long pos = 0;          // warp-uniform state
    int i = 0, run = 0;
    while (i < natoms) {
        const int per = 1 + run / 3;                                 // atoms of a group without a flag
        const unsigned long long glen = (unsigned long long)bitsize + 1ull + (unsigned long long)(run / 3) * smallidx;
        const unsigned long long p = pos + (unsigned long long)lane * glen;
        const int i0 = i + lane * per;
        const bool active = i0 < natoms && p + bitsize + 1 <= nbits_total;          // my flag bit lies inside the stream
        int flag = 0;
        if (active) flag = (int)xtc_bits(data, p + bitsize, 1);
        const unsigned am = __ballot_sync(0xffffffffu, active), fm = __ballot_sync(0xffffffffu, active && flag);
        if (!(am & 1u)) { fail(); return; }                          // lane 0 ran past the end of the stream
        const int first = fm ? __ffs(fm) - 1 : 32;
        const int ngroups = min(first + 1, __popc(am));             // active lanes are a prefix
        const bool mine = lane < ngroups;
        int myrun = run, change = 0;
        unsigned long long q = p + bitsize + 1;                      // first small atom of my group
        if (mine && flag) {
            const int r5 = p + bitsize + 6 <= nbits_total ? (int)xtc_bits(data, p + bitsize + 1, 5) : 31;   // 31: rejected below
            change = r5 % 3;
            myrun = r5 - change;
            change -= 1;
            q += 5;
        }
        const int nsmall = myrun / 3;
        if (mine) {
            if (nsmall > 8 || i0 + 1 + nsmall > natoms || q + (unsigned long long)nsmall * smallidx > nbits_total) {
                atomicExch(A.status, 1);
            } else {
                int big[3];
                if (wide) {
                    unsigned long long pp = p;
#pragma unroll
                    for (int k = 0; k < 3; ++k) { big[k] = wb[k] ? (int)xtc_bits(data, pp, wb[k]) : 0; pp += wb[k]; }
                } else {
                    xtc_triple(data, p, bitsize, size[1], size[2], big[0], big[1], big[2]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) big[k] += lo[k];
                float* o = st + (size_t)(i0 - i) * 3;
                auto emit = [&](float* dst, const int* v) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) dst[k] = __fmul_rn(__fmul_rn(__int2float_rn(v[k]), inv), scale);
                };
                if (nsmall == 0) {
                    emit(o, big);
                } else {
                    int prev[3] = {big[0], big[1], big[2]};
                    for (int t = 0; t < nsmall; ++t) {
                        int s[3];
                        xtc_triple(data, q + (unsigned long long)t * smallidx, smallidx, ssize, ssize, s[0], s[1], s[2]);
#pragma unroll
                        for (int k = 0; k < 3; ++k) { s[k] += prev[k] - smallnum; prev[k] = s[k]; }
                        // stream order L, S1, S2, ... is atom order S1, L, S2, ... (the first two were swapped)
                        emit(o + 3 * (t == 0 ? 0 : t + 1), s);
                    }
                    emit(o + 3, big);
                }
            }
        }
        // new warp state: from the flagged lane if there was one, else after the last group
        const int src = ngroups - 1;
        const int n_i = __shfl_sync(0xffffffffu, i0 + 1 + nsmall, src);
        const unsigned long long n_pos = __shfl_sync(0xffffffffu, q + (unsigned long long)nsmall * smallidx, src);
        const int n_run = __shfl_sync(0xffffffffu, myrun, src), n_change = __shfl_sync(0xffffffffu, change, src);
        __syncwarp();
        const int nout = min(n_i, natoms) - i;                      // atoms staged in this step
        for (int k = lane; k < nout * 3; k += 32)
            if (wanted(i + k / 3)) out[(size_t)i * 3 + k] = st[k];
        __syncwarp();
        if (n_i <= i || n_i > natoms || n_run > 24) { fail(); return; }
        i = n_i;
        pos = n_pos;
        run = n_run;
        if (n_change) {
            smallidx += n_change;
            if (smallidx < 9 || smallidx > 72) { fail(); return; }
            if (n_change < 0) { smallnum = smaller; smaller = smallidx > 9 ? c_xtc_magic[smallidx - 1] / 2 : 0; }
            else { smaller = smallnum; smallnum = c_xtc_magic[smallidx] / 2; }
            ssize = (unsigned)c_xtc_magic[smallidx];
        }
    }
}

}  // namespace

// hb_xtc.cuh - XTC frame records decoded on the device (frame sources; basic.py:46-47 reads any
// MDAnalysis format, for .xtc MDAnalysis decodes every frame on the CPU with xdrfile).
//
// The compressed records cross PCIe as they lie in the file (about 3.5 bytes per atom instead of 12) and one WARP
// decodes one frame into the float32 [atom][3] layout the search kernels read:
//
//   The bit stream is a sequence of groups {large atom: `bitsize` bits, flag bit, [5 bits: run + smallidx change],
//   run / 3 small atoms of `smallidx` bits each}. Where a group starts depends on all groups before it. Per window
//   of 4 kB of the stream (staged in shared memory with 16-byte loads) there are two phases, done by different warps of
//   the frame's CTA (the scanner warp runs one window ahead of the decoder warps; named barriers hand the buffers over):
//   SCAN   finds the groups: as long as the flag bit is 0 nothing changes, so lane k assumes the next k groups have
//          no flag and looks at "its" flag bit; the lanes up to and including the first one that finds a flag are
//          right and queue one descriptor each (bit offset, atom index, run, smallidx), the flagged lane's new run /
//          smallidx become the warp's state. In water (O, H, H: run = 6 for tens of thousands of groups) one step
//          finds 32 molecules; the dependent chain of a step is two shared-memory reads and a few shuffles.
//   DECODE takes 32 queued groups at a time, one per lane, whatever their flags were. A triple (a, b, c) with ranges
//          (sa, sb, sc) is v = (a * sb + b) * sc + c, stored as the bytes of v least significant first (the last
//          chunk holds the remaining 1..8 bits), bits most significant first: read as a 64-bit big-endian window,
//          byte-swapped and divided by multiplication with precomputed reciprocals (Granlund-Montgomery; the 73
//          possible small ranges come from constant memory, the frame's own ranges are derived once per frame;
//          128-bit arithmetic only when nbits > 64). Decoded atoms are staged in shared memory and stored as full
//          warp-wide float rows.
//
// Arithmetic of the published decoder (xdrfile_decompress_coord_float) and of MDAnalysis' unit conversion:
//   x = fl32(fl32(float(int) * fl32(1 / precision)) * scale), scale = 10 (nm -> Angstrom), every step rounded.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "hb_common.cuh"

namespace {

#define XTC_WARPS 4
#define XTC_STAGE_ATOMS (32 * 9)
#define XTC_WIN_BYTES 4096                 // bit-stream window per warp (16 more bytes are loaded behind it)
#define XTC_QCAP 512                       // group descriptors per window
#define XTC_GROUP_MAX_BITS 704             // 96 + 1 + 5 + 8 * 72, rounded up

__constant__ int c_xtc_magic[73] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
    1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
    65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
    2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
// reciprocals of c_xtc_magic[9..72] for xtc_fastdiv (filled by hb_create through xtc_upload_tables)
__constant__ unsigned long long c_xtc_mul[73];
__constant__ int c_xtc_shift[73];

struct XtcRanges {            // atom ranges [b, e) whose coordinates are written (n = 0: all atoms)
    int n;
    int b[8], e[8];
};

struct XtcArgs {
    const unsigned char* raw;       // records of the batch, 4-byte aligned, 64 readable bytes past the end
    const long long* frame_off;     // [n_frames + 1] byte offsets of the records inside raw
    int n_frames;
    int n_atoms;
    float scale;
    float* out;                     // [n_frames][n_atoms][3]
    int* status;                    // set to 1 when a record is malformed
    XtcRanges ranges;
};

// n / d for every 64-bit n by one multiplication: d >= 2, l = ceil(log2 d), mul = floor(2^64 (2^l - d) / d) + 1,
// shift = l - 1:  t = mulhi(mul, n);  q = (t + ((n - t) >> 1)) >> shift      (Granlund & Montgomery 1994, fig. 4.1)
__host__ __device__ inline void xtc_make_div(unsigned d, unsigned long long& mul, int& shift) {
    int fl = 0;
    while ((d >> fl) > 1u) ++fl;                                   // floor(log2 d)
    const int l = (d & (d - 1u)) ? fl + 1 : fl;
    const unsigned __int128 num = ((unsigned __int128)1 << 64) * ((((unsigned __int128)1) << l) - d);
    mul = (unsigned long long)(num / d) + 1ull;
    shift = l - 1;
}
__device__ __forceinline__ unsigned long long xtc_fastdiv(unsigned long long n, unsigned long long mul, int shift) {
    const unsigned long long t = __umul64hi(mul, n);
    return (t + ((n - t) >> 1)) >> shift;
}

__device__ __forceinline__ unsigned xtc_be32(const unsigned char* p) {       // p is 4-byte aligned
    return __byte_perm(*reinterpret_cast<const unsigned*>(p), 0, 0x0123);
}
// `n` (1..64) bits of the big-endian bit stream starting at bit `pos` of the 4-byte aligned shared-memory window
// (the window holds the stream as native 32-bit words: byte-swapped when it was staged)
__device__ __forceinline__ unsigned long long xtc_bits(const unsigned* win, unsigned pos, int n) {
    const unsigned* w = win + (pos >> 5);
    const unsigned a = w[0], b = w[1], c = w[2];
    const int s = (int)(pos & 31);
    unsigned long long hi = ((unsigned long long)a << 32) | b;
    if (s) hi = (hi << s) | (c >> (32 - s));
    return hi >> (64 - n);
}
// same for n <= 32: two words and one funnel shift
__device__ __forceinline__ unsigned xtc_bits32(const unsigned* win, unsigned pos, int n) {
    const unsigned* w = win + (pos >> 5);
    return __funnelshift_l(w[1], w[0], pos & 31u) >> (32 - n);
}
__device__ __forceinline__ unsigned long long xtc_bswap64(unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
// the integer v of a triple stored in `nbits` (<= 64) bits at `pos`
__device__ __forceinline__ unsigned long long xtc_triple64(const unsigned* win, unsigned pos, int nbits) {
    const int k = (nbits - 1) >> 3, r = nbits - 8 * k;              // k full bytes, then r = 1..8 bits
    const unsigned long long w = xtc_bits(win, pos, nbits);
    const unsigned long long last = w & ((1ull << r) - 1ull);
    unsigned long long v = 0;
    if (k) v = xtc_bswap64((w >> r) << (64 - 8 * k));             // chunk 0 (first in the stream) is the lowest byte
    return v | (last << (8 * k));                                   // k <= 7 here
}
struct XtcDiv { unsigned d; unsigned long long mul; int shift; };
// v = (a * sb + b) * sc + c  ->  a, b, c
__device__ __forceinline__ void xtc_split(unsigned long long v, const XtcDiv& sb, const XtcDiv& sc, int& a, int& b, int& c) {
    const unsigned long long q = sc.d > 1u ? xtc_fastdiv(v, sc.mul, sc.shift) : v;
    c = (int)(unsigned)(v - q * sc.d);
    const unsigned long long q2 = sb.d > 1u ? xtc_fastdiv(q, sb.mul, sb.shift) : q;
    b = (int)(unsigned)(q - q2 * sb.d);
    a = (int)(unsigned)q2;
}
// general triple (nbits up to 72)
__device__ __noinline__ void xtc_triple_wide(const unsigned* win, unsigned pos, int nbits, unsigned sb, unsigned sc, int* abc) {
    unsigned __int128 v = 0;
    int shift = 0, left = nbits;
    while (left > 8) { v |= (unsigned __int128)xtc_bits(win, pos, 8) << shift; pos += 8; shift += 8; left -= 8; }
    v |= (unsigned __int128)xtc_bits(win, pos, left) << shift;
    abc[2] = (int)(unsigned)(v % sc); v /= sc;
    abc[1] = (int)(unsigned)(v % sb); v /= sb;
    abc[0] = (int)(unsigned)v;
}
__device__ __forceinline__ void xtc_triple(const unsigned* win, unsigned pos, int nbits, const XtcDiv& sb, const XtcDiv& sc, int* abc) {
    if (nbits <= 64) xtc_split(xtc_triple64(win, pos, nbits), sb, sc, abc[0], abc[1], abc[2]);
    else xtc_triple_wide(win, pos, nbits, sb.d, sc.d, abc);
}
__device__ __forceinline__ int xtc_bit_length(unsigned __int128 v) {
    const unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
    return hi ? 128 - __clzll((long long)hi) : (lo ? 64 - __clzll((long long)lo) : 0);
}

struct __align__(16) XtcWindow {                  // one window of the bit stream and what the scan found in it
    unsigned win[XTC_WIN_BYTES / 4 + 8];          // the stream as native words (+ 16 bytes read-ahead, + padding)
    uint2 queue[XTC_QCAP];                        // x: bit offset inside the window | nsmall << 16 | flag << 20 | smallidx << 21, y: atom
    long long wbase_bits;                         // stream position of the window
    int nq, i_first, i_end;                       // groups queued (-1: no more windows), atoms before / behind them
};

struct XtcShared {
    XtcWindow w[2];
    float stage[XTC_WARPS - 1][XTC_STAGE_ATOMS * 3];
    int bad;
    int rb[8], re[8];                             // atom ranges to write
};

#define XTC_BAR_FULL 1                             // named barriers 1, 2: window b scanned (scanner arrives, decoders wait)
#define XTC_BAR_EMPTY 3                            // named barriers 3, 4: window b decoded (decoders arrive, scanner waits)
// (barrier numbers as immediates: with a register operand ptxas reserves all 16 barriers for the CTA, which costs occupancy)
template <int ID>
__device__ __forceinline__ void xtc_bar_sync_id() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(XTC_WARPS * 32) : "memory"); }
template <int ID>
__device__ __forceinline__ void xtc_bar_arrive_id() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(XTC_WARPS * 32) : "memory"); }
__device__ __forceinline__ void xtc_bar_sync(int base, int b) {
    __syncwarp();
    if (base == XTC_BAR_FULL) { if (b) xtc_bar_sync_id<XTC_BAR_FULL + 1>(); else xtc_bar_sync_id<XTC_BAR_FULL>(); }
    else { if (b) xtc_bar_sync_id<XTC_BAR_EMPTY + 1>(); else xtc_bar_sync_id<XTC_BAR_EMPTY>(); }
}
__device__ __forceinline__ void xtc_bar_arrive(int base, int b) {
    __syncwarp();
    __threadfence_block();                         // what this warp wrote to shared memory is visible to the waiting side
    if (base == XTC_BAR_FULL) { if (b) xtc_bar_arrive_id<XTC_BAR_FULL + 1>(); else xtc_bar_arrive_id<XTC_BAR_FULL>(); }
    else { if (b) xtc_bar_arrive_id<XTC_BAR_EMPTY + 1>(); else xtc_bar_arrive_id<XTC_BAR_EMPTY>(); }
}

// One CTA (XTC_WARPS warps) per frame, warp specialised: warp 0 stages windows and scans them, running one window ahead;
// the other warps decode the groups it queued.
__global__ void __launch_bounds__(XTC_WARPS * 32, 6) k_xtc_decode(XtcArgs A) {
    __shared__ __align__(16) XtcShared S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x;
    const unsigned char* rec = A.raw + A.frame_off[f];
    const long long rec_len = A.frame_off[f + 1] - A.frame_off[f];
    float* out = A.out + (size_t)f * A.n_atoms * 3;
    const int natoms = A.n_atoms;
    auto fail = [&]() { if (threadIdx.x == 0) atomicExch(A.status, 1); };
    // (every exit below is taken by the whole CTA: the conditions only depend on the record header)
    if (rec_len < 56 || xtc_be32(rec) != 1995u || (int)xtc_be32(rec + 4) != natoms || (int)xtc_be32(rec + 52) != natoms) { fail(); return; }
    const int n_ranges = A.ranges.n;
    if (threadIdx.x < 8) { S.rb[threadIdx.x] = A.ranges.b[threadIdx.x]; S.re[threadIdx.x] = A.ranges.e[threadIdx.x]; }
    if (natoms <= 9) {                                   // stored as plain floats
        if (rec_len < 56 + 12ll * natoms) { fail(); return; }
        for (int k = threadIdx.x; k < 3 * natoms; k += XTC_WARPS * 32) out[k] = __fmul_rn(__uint_as_float(xtc_be32(rec + 56 + 4 * k)), A.scale);
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
    const int smallidx0 = (int)xtc_be32(rec + 84);
    const unsigned long long nbits_total = 8ull * xtc_be32(rec + 88);
    if (smallidx0 < 9 || smallidx0 > 72 || size[0] == 0 || size[1] == 0 || size[2] == 0 ||
        92ull + nbits_total / 8 > (unsigned long long)rec_len) { fail(); return; }
    const unsigned char* data = rec + 92;
    const bool wide = (size[0] | size[1] | size[2]) > 0xffffffu;
    int wb[3] = {0, 0, 0}, bitsize;
    if (wide) {
#pragma unroll
        for (int k = 0; k < 3; ++k) wb[k] = 32 - __clz(size[k]);      // bits of an int of this range
        bitsize = wb[0] + wb[1] + wb[2];
    } else {
        bitsize = xtc_bit_length((unsigned __int128)size[0] * size[1] * size[2]);
    }
    if (threadIdx.x == 0) S.bad = 0;
    __syncthreads();

    if (warp == 0) {
        // ---- scanner: stage a window, find the groups that lie completely inside it, hand it over ----------------------
        // The scan only finds where groups start; it reads nothing but flag bits inside [0, lim] of the window and
        // always advances, so it needs no validation of its own: the decoders check every descriptor they use.
        unsigned long long pos0 = 0;         // bit position of the next group
        int i = 0, nsm = 0, smallidx = smallidx0;        // atoms found so far, small atoms per group until a flag changes it
        bool bad = false;
        for (int wi = 0;; ++wi) {
            XtcWindow& W = S.w[wi & 1];
            if (wi >= 2) xtc_bar_sync(XTC_BAR_EMPTY, wi & 1);     // the decoders are done with this buffer
            if (i >= natoms || bad || *reinterpret_cast<volatile int*>(&S.bad)) {
                if (lane == 0) { W.nq = -1; if (bad) S.bad = 1; }
                xtc_bar_arrive(XTC_BAR_FULL, wi & 1);
                break;
            }
            // window: 16-byte aligned global address at or before the byte of pos0
            const unsigned long long a0 = ((unsigned long long)(uintptr_t)data + (pos0 >> 3)) & ~15ull;
            const long long wbase_bits = 8ll * ((long long)a0 - (long long)(uintptr_t)data);      // may be negative (header bytes)
            {
                const uint4* g = reinterpret_cast<const uint4*>(a0);
                uint4* w = reinterpret_cast<uint4*>(W.win);
                for (int k = lane; k < XTC_WIN_BYTES / 16 + 1; k += 32) {
                    uint4 v = __ldg(g + k);
                    v.x = __byte_perm(v.x, 0, 0x0123); v.y = __byte_perm(v.y, 0, 0x0123);
                    v.z = __byte_perm(v.z, 0, 0x0123); v.w = __byte_perm(v.w, 0, 0x0123);
                    w[k] = v;
                }
                // the two windows after this one: into L2 while this one is scanned and decoded
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + XTC_WIN_BYTES + 256ull * lane));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a0 + XTC_WIN_BYTES + 256ull * lane + 128ull));
            }
            __syncwarp();
            const long long wend_bits = wbase_bits + 8ll * XTC_WIN_BYTES;
            const bool last_window = wend_bits >= (long long)nbits_total;
            const long long end_rel = (long long)nbits_total - wbase_bits;                      // stream end, window relative
            // a group may start up to `lim` (its flag bit lies inside the stream and, unless this is the last window,
            // the whole group inside the window)
            const int lim = (int)min(end_rel - bitsize - 1, last_window ? (1ll << 30) : (long long)(8 * XTC_WIN_BYTES - XTC_GROUP_MAX_BITS));
            int prel0 = (int)((long long)pos0 - wbase_bits);
            const int i_first = i;
            int nq = 0;
            while (i < natoms && nq <= XTC_QCAP - 32) {
                const int glen = bitsize + 1 + nsm * smallidx;
                const unsigned sidx_bits = (unsigned)smallidx << 21;
                const int prel = prel0 + lane * glen;
                const int i0 = i + lane * (1 + nsm);
                const bool active = i0 < natoms && prel <= lim;
                // flag bit and the five bits behind it in one read
                unsigned six = 0;
                if (active) six = xtc_bits32(W.win, (unsigned)(prel + bitsize), 6);
                const unsigned am = __ballot_sync(0xffffffffu, active), fm = __ballot_sync(0xffffffffu, six >> 5);
                if (!(am & 1u)) {
                    // lane 0 cannot go on: the window is used up (slide it), or the stream ended before the last atom
                    bad = last_window || nq == 0;
                    break;
                }
                const int f1 = __ffs(fm);                                    // flagged lanes are active lanes, active lanes a prefix
                const int ngroups = f1 ? f1 : __popc(am);                    // up to and including the first flag
                const int flag = (int)(six >> 5);
                const int r5 = (int)(six & 31u);
                const int fs = (r5 * 11) >> 5;                               // r5 / 3: small atoms of a flagged group
                const int nsmall = flag ? fs : nsm;
                const int change = flag ? r5 - 3 * fs : 1;                   // smallidx change + 1
                const int pnext = prel + bitsize + 1 + 5 * flag + nsmall * smallidx;
                HB_DASSERT(nq + ngroups <= XTC_QCAP && prel >= 0 && (lane >= ngroups || prel + bitsize + 6 <= 8 * XTC_WIN_BYTES + 128));
                if (lane < ngroups) W.queue[nq + lane] = make_uint2((unsigned)prel | ((unsigned)nsmall << 16) | (six >> 5 << 20) | sidx_bits,
                                                                    (unsigned)(i0 - i_first));
                const int src = ngroups - 1;
                i = __shfl_sync(0xffffffffu, i0 + 1 + nsmall, src);
                prel0 = __shfl_sync(0xffffffffu, pnext, src);
                const int st = __shfl_sync(0xffffffffu, nsmall | (change << 8), src);
                nsm = st & 0xff;
                smallidx = min(max(smallidx + (st >> 8) - 1, 0), 127);     // (kept sane for the arithmetic; decoders reject < 9 or > 72)
                nq += ngroups;
            }
            pos0 = (unsigned long long)(wbase_bits + (long long)prel0);
            if (lane == 0) { W.wbase_bits = wbase_bits; W.nq = nq; W.i_first = i_first; W.i_end = i; }
            xtc_bar_arrive(XTC_BAR_FULL, wi & 1);
        }
    } else {
        // ---- decoders: 32 queued groups per warp and pass -------------------------------------------------------------------
        XtcDiv dy{size[1], 0, 0}, dz{size[2], 0, 0};
        if (!wide) {
            if (size[1] > 1u) xtc_make_div(size[1], dy.mul, dy.shift);
            if (size[2] > 1u) xtc_make_div(size[2], dz.mul, dz.shift);
        }
        const float inv = (float)(1.0 / (double)precision);
        const float scale = A.scale;
        float* stage = S.stage[warp - 1];
        for (int wi = 0;; ++wi) {
            XtcWindow& W = S.w[wi & 1];
            xtc_bar_sync(XTC_BAR_FULL, wi & 1);
            const int nq = W.nq;
            if (nq < 0) break;
            const int i_first = W.i_first, i_end = W.i_end;
            const long long wbase_bits = W.wbase_bits;
            for (int g0 = 32 * (warp - 1); g0 < nq; g0 += 32 * (XTC_WARPS - 1)) {
                const int g = g0 + lane;
                const int a_first = (int)W.queue[g0].y;                                               // first atom of this pass
                const int a_end = g0 + 32 < nq ? (int)W.queue[g0 + 32].y : i_end - i_first;          // first atom of the next
                const int a_lim = min(a_end, a_first + XTC_STAGE_ATOMS);    // (a_end - a_first > 288 only for malformed streams)
                if (g < nq) {
                    const uint2 d = W.queue[g];
                    const unsigned prel = d.x & 0xffffu;
                    const int arel = (int)d.y, nsmall = (int)((d.x >> 16) & 15u), flag = (int)((d.x >> 20) & 1u);
                    const int sidx = (int)(d.x >> 21);
                    const long long gend = wbase_bits + (long long)prel + bitsize + 1 + 5 * flag + (long long)nsmall * sidx;
                    if (nsmall > 8 || sidx < 9 || sidx > 72 || arel < a_first || arel + 1 + nsmall > a_lim ||
                        i_first + arel + 1 + nsmall > natoms || gend > (long long)nbits_total) {
                        S.bad = 1;                                            // malformed stream (benign race: every writer writes 1)
                    } else {
                        int big[3];
                        if (wide) {
                            unsigned pp = prel;
#pragma unroll
                            for (int k = 0; k < 3; ++k) { big[k] = (int)xtc_bits(W.win, pp, wb[k]); pp += wb[k]; }
                        } else {
                            xtc_triple(W.win, prel, bitsize, dy, dz, big);
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) big[k] += lo[k];
                        HB_DASSERT(arel - a_first >= 0 && arel - a_first + 1 + nsmall <= XTC_STAGE_ATOMS && prel + bitsize <= 8u * XTC_WIN_BYTES + 128u);
                        float* o = stage + (size_t)(arel - a_first) * 3;
                        auto emit = [&](float* dst, const int* v) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) dst[k] = __fmul_rn(__fmul_rn(__int2float_rn(v[k]), inv), scale);
                        };
                        if (nsmall == 0) {
                            emit(o, big);
                        } else {
                            const XtcDiv ds{(unsigned)c_xtc_magic[sidx], c_xtc_mul[sidx], c_xtc_shift[sidx]};
                            const int smallnum = c_xtc_magic[sidx] / 2;
                            const unsigned q = prel + bitsize + 1 + (flag ? 5 : 0);
                            int prev[3] = {big[0], big[1], big[2]};
                            for (int t = 0; t < nsmall; ++t) {
                                int sm[3];
                                xtc_triple(W.win, q + (unsigned)(t * sidx), sidx, ds, ds, sm);
#pragma unroll
                                for (int k = 0; k < 3; ++k) { sm[k] += prev[k] - smallnum; prev[k] = sm[k]; }
                                // stream order L, S1, S2, ... is atom order S1, L, S2, ... (the first two were swapped)
                                emit(o + 3 * (t == 0 ? 0 : t + 1), sm);
                            }
                            emit(o + 3, big);
                        }
                    }
                }
                __syncwarp();
                // write the atoms of this pass: whole pass inside one wanted range (or no ranges) -> plain row copy
                const int abase = i_first + a_first, npass = max(a_lim - a_first, 0);
                int mode = n_ranges == 0 ? 1 : 0;                           // 1: all, 0: none, 2: mixed
                for (int k = 0; k < n_ranges; ++k) {
                    if (abase >= S.rb[k] && abase + npass <= S.re[k]) mode = 1;
                    else if (abase < S.re[k] && abase + npass > S.rb[k] && mode == 0) mode = 2;
                }
                if (mode == 1) {
                    for (int k = lane; k < npass * 3; k += 32) out[(size_t)abase * 3 + k] = stage[k];
                } else if (mode == 2) {
                    for (int k = lane; k < npass * 3; k += 32) {
                        const int atom = abase + k / 3;
                        bool w = false;
                        for (int r = 0; r < n_ranges; ++r) w = w || (atom >= S.rb[r] && atom < S.re[r]);
                        if (w) out[(size_t)abase * 3 + k] = stage[k];
                    }
                }
                __syncwarp();
            }
            xtc_bar_arrive(XTC_BAR_EMPTY, wi & 1);
        }
    }
    __syncthreads();
    if (S.bad) fail();
}

// reciprocal tables of the 64 small ranges (host; once per device)
inline cudaError_t xtc_upload_tables() {
    static const int magic[73] = {
        0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812,
        1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007, 32768, 41285, 52015,
        65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561, 832255, 1048576, 1321122, 1664510,
        2097152, 2642245, 3329021, 4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
    unsigned long long mul[73] = {0};
    int shift[73] = {0};
    for (int k = 9; k < 73; ++k) xtc_make_div((unsigned)magic[k], mul[k], shift[k]);
    cudaError_t e = cudaMemcpyToSymbol(c_xtc_mul, mul, sizeof(mul));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_xtc_shift, shift, sizeof(shift));
    return e;
}

}  // namespace

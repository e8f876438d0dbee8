/* TEST INFRASTRUCTURE - CPU restatement of the GROMACS XTC frame codec (not part of the product path).
 *
 * The reference reads trajectories through MDAnalysis.Universe(structure, trajectories)
 * (cgraphs/mdhbond/basic.py:46-47); for .xtc files MDAnalysis decodes every frame on the CPU with its bundled copy
 * of the xdrfile library (MDAnalysis/lib/formats/src/xdrfile.c, xdrfile 1.1.4: xdrfile_decompress_coord_float, the
 * same algorithm as GROMACS' libxdrf.c xdr3dfcoord) and converts nm to Angstrom in float32
 * (MDAnalysis/coordinates/XDR.py + base.py convert_pos_from_native: `x *= 10.0` on the float32 array).
 * Neither MDAnalysis nor xdrfile nor any .xtc file exists under /root/reference or in this image, so this file
 * restates the PUBLISHED algorithm of that third-party dependency from its specification:
 *
 *   frame record (XDR, big endian): int magic = 1995, int natoms, int step, float time, float box[3][3] (nm),
 *   then the coordinate block: int natoms; natoms <= 9: 3 * natoms floats, else float precision, int minint[3],
 *   int maxint[3], int smallidx, int nbytes, nbytes of bit stream padded to a multiple of 4.
 *   Bit stream: bits are written most significant first (sendbits); a triple of integers (x, y, z) with ranges
 *   (sx, sy, sz) is the number v = (x * sy + y) * sz + z written as its bytes, LEAST significant byte first, the
 *   last chunk holding the remaining 1..8 top bits (sendints / receiveints); every "large" atom is such a triple of
 *   the offsets from minint in `bitsize` bits, followed by a flag bit (1: five bits follow that hold run + change of
 *   smallidx), followed by run / 3 "small" atoms, each a triple of the differences to the previous atom plus
 *   smallnum in `smallidx` bits with ranges magicints[smallidx]; the first small atom of a run and the large atom
 *   before it are stored in swapped order (water: O after H).
 *
 * PARITY UNPINNED: there is no golden .xtc file and no real reader in this image to check this restatement against;
 * it is checked only for self-consistency (encode -> decode reproduces the quantised input, every adaptive path is
 * exercised). The CUDA decoder and the product's own host encoder (cgraphs_b200/csrc) are tested bit-exact against
 * this file. DESIGN.md says the same.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XTC_MAGIC 1995
#define FIRSTIDX 9
#define MAXABS (INT32_MAX - 2)

static const int magicints[] = {
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64,
    80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290,
    1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003,
    16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031,
    131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
    832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021,
    4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216};
#define LASTIDX ((int)(sizeof(magicints) / sizeof(*magicints)) - 1)

/* ---- XDR primitives (big endian) ------------------------------------------------------------------ */
static void put_u32(unsigned char* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
static uint32_t get_u32(const unsigned char* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
static void put_f32(unsigned char* p, float f) { uint32_t v; memcpy(&v, &f, 4); put_u32(p, v); }
static float get_f32(const unsigned char* p) { uint32_t v = get_u32(p); float f; memcpy(&f, &v, 4); return f; }

/* ---- bit stream ------------------------------------------------------------------------------------- */
typedef struct {
    unsigned char* c;      /* byte buffer */
    size_t cnt;            /* bytes complete */
    unsigned lastbits;     /* bits pending in lastbyte */
    unsigned lastbyte;
} bitbuf;

static int sizeofint(int size) {
    unsigned num = 1;
    int bits = 0;
    while ((unsigned)size >= num && bits < 32) { bits++; num <<= 1; }
    return bits;
}

static int sizeofints(int n, const unsigned sizes[]) {
    unsigned char bytes[32];
    unsigned nbytes = 1, bytecnt, tmp, num;
    int i, nbits = 0;
    bytes[0] = 1;
    for (i = 0; i < n; i++) {
        tmp = 0;
        for (bytecnt = 0; bytecnt < nbytes; bytecnt++) {
            tmp = bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = tmp & 0xff;
            tmp >>= 8;
        }
        while (tmp != 0) { bytes[bytecnt++] = tmp & 0xff; tmp >>= 8; }
        nbytes = bytecnt;
    }
    num = 1;
    nbytes--;
    while (bytes[nbytes] >= num) { nbits++; num *= 2; }
    return nbits + (int)nbytes * 8;
}

static void sendbits(bitbuf* b, int nbits, unsigned num) {
    while (nbits >= 8) {
        b->lastbyte = (b->lastbyte << 8) | ((num >> (nbits - 8)) & 0xff);
        b->c[b->cnt++] = (unsigned char)(b->lastbyte >> b->lastbits);
        nbits -= 8;
    }
    if (nbits > 0) {
        b->lastbyte = (b->lastbyte << nbits) | (num & ((1u << nbits) - 1u));
        b->lastbits += nbits;
        if (b->lastbits >= 8) {
            b->lastbits -= 8;
            b->c[b->cnt++] = (unsigned char)(b->lastbyte >> b->lastbits);
        }
    }
    if (b->lastbits > 0) b->c[b->cnt] = (unsigned char)(b->lastbyte << (8 - b->lastbits));
}

static void sendints(bitbuf* b, int n, int nbits, const unsigned sizes[], const unsigned nums[]) {
    unsigned char bytes[32];
    unsigned nbytes = 0, bytecnt, tmp;
    int i;
    tmp = nums[0];
    do { bytes[nbytes++] = tmp & 0xff; tmp >>= 8; } while (tmp != 0);
    for (i = 1; i < n; i++) {
        tmp = nums[i];
        for (bytecnt = 0; bytecnt < nbytes; bytecnt++) {
            tmp = bytes[bytecnt] * sizes[i] + tmp;
            bytes[bytecnt] = tmp & 0xff;
            tmp >>= 8;
        }
        while (tmp != 0) { bytes[bytecnt++] = tmp & 0xff; tmp >>= 8; }
        nbytes = bytecnt;
    }
    if ((unsigned)nbits >= nbytes * 8) {
        for (i = 0; i < (int)nbytes; i++) sendbits(b, 8, bytes[i]);
        sendbits(b, nbits - (int)nbytes * 8, 0);
    } else {
        for (i = 0; i < (int)nbytes - 1; i++) sendbits(b, 8, bytes[i]);
        sendbits(b, nbits - ((int)nbytes - 1) * 8, bytes[i]);
    }
}

static unsigned receivebits(bitbuf* b, int nbits) {
    const unsigned mask = nbits >= 32 ? 0xffffffffu : (1u << nbits) - 1u;
    unsigned num = 0;
    while (nbits >= 8) {
        b->lastbyte = (b->lastbyte << 8) | b->c[b->cnt++];
        num |= (b->lastbyte >> b->lastbits) << (nbits - 8);
        nbits -= 8;
    }
    if (nbits > 0) {
        if ((int)b->lastbits < nbits) {
            b->lastbits += 8;
            b->lastbyte = (b->lastbyte << 8) | b->c[b->cnt++];
        }
        b->lastbits -= nbits;
        num |= (b->lastbyte >> b->lastbits) & ((1u << nbits) - 1u);
    }
    return num & mask;
}

static void receiveints(bitbuf* b, int n, int nbits, const unsigned sizes[], int nums[]) {
    /* the published code keeps `num` and `p` in ints and divides by the unsigned sizes[i], i.e. the arithmetic is
       unsigned 32-bit (num < sizes[i] * 256 <= 2^32); stated with unsigned variables here */
    unsigned bytes[32];
    int i, j, nbytes = 0;
    unsigned p, num;
    bytes[1] = bytes[2] = bytes[3] = 0;
    while (nbits > 8) { bytes[nbytes++] = receivebits(b, 8); nbits -= 8; }
    if (nbits > 0) bytes[nbytes++] = receivebits(b, nbits);
    for (i = n - 1; i > 0; i--) {
        num = 0;
        for (j = nbytes - 1; j >= 0; j--) {
            num = (num << 8) | bytes[j];
            p = num / sizes[i];
            bytes[j] = p;
            num = num - p * sizes[i];
        }
        nums[i] = (int)num;
    }
    nums[0] = (int)(bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24));
}

/* ---- frame record -------------------------------------------------------------------------------------- */
#define HEADER_BYTES (4 * (4 + 9))      /* magic, natoms, step, time, box */

/* Upper bound of the bytes one encoded frame can take. */
size_t xtc_oracle_frame_bound(int natoms) { return HEADER_BYTES + 4 + 4 * 9 + (size_t)natoms * 3 * 5 + 64; }

/* Encode one frame (coordinates and box in nm). Returns the record length in bytes, 0 on error. */
size_t xtc_oracle_encode(const float* xyz, int natoms, float precision, int step, float time, const float* box9,
                         unsigned char* out) {
    unsigned char* p = out;
    int k;
    put_u32(p, XTC_MAGIC); put_u32(p + 4, (uint32_t)natoms); put_u32(p + 8, (uint32_t)step); put_f32(p + 12, time);
    for (k = 0; k < 9; k++) put_f32(p + 16 + 4 * k, box9[k]);
    p += HEADER_BYTES;
    put_u32(p, (uint32_t)natoms); p += 4;
    if (natoms <= 9) {
        for (k = 0; k < 3 * natoms; k++) { put_f32(p, xyz[k]); p += 4; }
        return (size_t)(p - out);
    }
    if (precision <= 0) precision = 1000;
    put_f32(p, precision); p += 4;

    int* ip = (int*)malloc((size_t)natoms * 3 * sizeof(int));
    unsigned char* cbuf = (unsigned char*)calloc((size_t)natoms * 3 * 5 + 64, 1);
    if (!ip || !cbuf) { free(ip); free(cbuf); return 0; }
    int minint[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, maxint[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    int mindiff = INT32_MAX, oldl[3] = {0, 0, 0};
    int i;
    for (i = 0; i < natoms; i++) {
        int l[3];
        for (k = 0; k < 3; k++) {
            float lf = xyz[3 * i + k] >= 0.0f ? xyz[3 * i + k] * precision + 0.5f : xyz[3 * i + k] * precision - 0.5f;
            if (fabsf(lf) > (float)MAXABS) { free(ip); free(cbuf); return 0; }
            l[k] = (int)lf;
            if (l[k] < minint[k]) minint[k] = l[k];
            if (l[k] > maxint[k]) maxint[k] = l[k];
            ip[3 * i + k] = l[k];
        }
        int diff = abs(oldl[0] - l[0]) + abs(oldl[1] - l[1]) + abs(oldl[2] - l[2]);
        if (diff < mindiff && i > 0) mindiff = diff;
        oldl[0] = l[0]; oldl[1] = l[1]; oldl[2] = l[2];
    }
    for (k = 0; k < 3; k++) { put_u32(p, (uint32_t)minint[k]); p += 4; }
    for (k = 0; k < 3; k++) { put_u32(p, (uint32_t)maxint[k]); p += 4; }
    for (k = 0; k < 3; k++)
        if ((float)maxint[k] - (float)minint[k] >= (float)MAXABS) { free(ip); free(cbuf); return 0; }
    unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
    for (k = 0; k < 3; k++) sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
        for (k = 0; k < 3; k++) bitsizeint[k] = (unsigned)sizeofint((int)sizeint[k]);
        bitsize = 0;
    } else {
        bitsize = sizeofints(3, sizeint);
    }
    int smallidx = FIRSTIDX;
    while (smallidx < LASTIDX && magicints[smallidx] < mindiff) smallidx++;
    put_u32(p, (uint32_t)smallidx); p += 4;
    int maxidx = smallidx + 8 < LASTIDX ? smallidx + 8 : LASTIDX;
    int minidx = maxidx - 8;
    int smaller = magicints[smallidx - 1 > FIRSTIDX ? smallidx - 1 : FIRSTIDX] / 2;
    int smallnum = magicints[smallidx] / 2;
    unsigned sizesmall[3];
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    int larger = magicints[maxidx] / 2;
    bitbuf b = {cbuf, 0, 0, 0};
    int prevrun = -1, prevcoord[3] = {0, 0, 0};
    i = 0;
    while (i < natoms) {
        int is_small = 0, is_smaller, run;
        int* thiscoord = ip + 3 * i;
        unsigned tmpcoord[30];
        if (smallidx < maxidx && i >= 1 && abs(thiscoord[0] - prevcoord[0]) < larger &&
            abs(thiscoord[1] - prevcoord[1]) < larger && abs(thiscoord[2] - prevcoord[2]) < larger)
            is_smaller = 1;
        else if (smallidx > minidx)
            is_smaller = -1;
        else
            is_smaller = 0;
        if (i + 1 < natoms) {
            if (abs(thiscoord[0] - thiscoord[3]) < smallnum && abs(thiscoord[1] - thiscoord[4]) < smallnum &&
                abs(thiscoord[2] - thiscoord[5]) < smallnum) {
                /* the first two atoms of a run change places (better compression of water: O after H) */
                for (k = 0; k < 3; k++) { int t = thiscoord[k]; thiscoord[k] = thiscoord[k + 3]; thiscoord[k + 3] = t; }
                is_small = 1;
            }
        }
        for (k = 0; k < 3; k++) tmpcoord[k] = (unsigned)(thiscoord[k] - minint[k]);
        if (bitsize == 0) {
            for (k = 0; k < 3; k++) sendbits(&b, (int)bitsizeint[k], tmpcoord[k]);
        } else {
            sendints(&b, 3, bitsize, sizeint, tmpcoord);
        }
        for (k = 0; k < 3; k++) prevcoord[k] = thiscoord[k];
        thiscoord += 3;
        i++;
        run = 0;
        if (is_small == 0 && is_smaller == -1) is_smaller = 0;
        while (is_small && run < 8 * 3) {
            if (is_smaller == -1) {
                const long long d0 = thiscoord[0] - prevcoord[0], d1 = thiscoord[1] - prevcoord[1], d2 = thiscoord[2] - prevcoord[2];
                if (d0 * d0 + d1 * d1 + d2 * d2 >= (long long)smaller * smaller) is_smaller = 0;
            }
            for (k = 0; k < 3; k++) tmpcoord[run++] = (unsigned)(thiscoord[k] - prevcoord[k] + smallnum);
            for (k = 0; k < 3; k++) prevcoord[k] = thiscoord[k];
            i++;
            thiscoord += 3;
            is_small = 0;
            if (i < natoms && abs(thiscoord[0] - prevcoord[0]) < smallnum && abs(thiscoord[1] - prevcoord[1]) < smallnum &&
                abs(thiscoord[2] - prevcoord[2]) < smallnum)
                is_small = 1;
        }
        if (run != prevrun || is_smaller != 0) {
            prevrun = run;
            sendbits(&b, 1, 1);                       /* the run length (or smallidx) changes */
            sendbits(&b, 5, (unsigned)(run + is_smaller + 1));
        } else {
            sendbits(&b, 1, 0);
        }
        for (k = 0; k < run; k += 3) sendints(&b, 3, smallidx, sizesmall, &tmpcoord[k]);
        if (is_smaller != 0) {
            smallidx += is_smaller;
            if (is_smaller < 0) {
                smallnum = smaller;
                smaller = magicints[smallidx - 1] / 2;
            } else {
                smaller = smallnum;
                smallnum = magicints[smallidx] / 2;
            }
            sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
        }
    }
    size_t nbytes = b.cnt + (b.lastbits != 0 ? 1 : 0);
    put_u32(p, (uint32_t)nbytes); p += 4;
    memcpy(p, cbuf, nbytes);
    p += nbytes;
    while ((p - out) & 3) *p++ = 0;
    free(ip);
    free(cbuf);
    return (size_t)(p - out);
}

/* Length of the frame record that starts at `in` (0: not a frame / truncated). */
size_t xtc_oracle_frame_bytes(const unsigned char* in, size_t avail, int* natoms) {
    if (avail < HEADER_BYTES + 4 || get_u32(in) != XTC_MAGIC) return 0;
    const int n = (int)get_u32(in + 4);
    if (natoms) *natoms = n;
    if (n <= 9) return HEADER_BYTES + 4 + (size_t)12 * n <= avail ? HEADER_BYTES + 4 + (size_t)12 * n : 0;
    if (avail < HEADER_BYTES + 4 + 4 * 9) return 0;
    const size_t nbytes = get_u32(in + HEADER_BYTES + 4 + 4 * 8);
    const size_t len = HEADER_BYTES + 4 + 4 * 9 + ((nbytes + 3) & ~(size_t)3);
    return len <= avail ? len : 0;
}

/* Decode one frame record: xyz_nm = int * fl32(1 / precision) (xdrfile), box in nm. If scale != 0, xyz_out = fl32(
 * xyz_nm * scale) as MDAnalysis' in-place float32 unit conversion does (scale = 10: Angstrom). Returns bytes consumed. */
size_t xtc_oracle_decode(const unsigned char* in, size_t avail, int natoms_expected, float scale, float* xyz_out,
                         float* box9, int* step, float* time, float* precision_out) {
    int natoms = 0, k, i;
    const size_t len = xtc_oracle_frame_bytes(in, avail, &natoms);
    if (!len || natoms != natoms_expected) return 0;
    if (step) *step = (int)get_u32(in + 8);
    if (time) *time = get_f32(in + 12);
    if (box9) for (k = 0; k < 9; k++) box9[k] = get_f32(in + 16 + 4 * k);
    const unsigned char* p = in + HEADER_BYTES;
    if ((int)get_u32(p) != natoms) return 0;
    p += 4;
    if (natoms <= 9) {
        for (k = 0; k < 3 * natoms; k++) {
            float v = get_f32(p + 4 * k);
            if (scale != 0.0f) v = v * scale;
            xyz_out[k] = v;
        }
        if (precision_out) *precision_out = 0.0f;
        return len;
    }
    const float precision = get_f32(p); p += 4;
    if (precision_out) *precision_out = precision;
    int minint[3], maxint[3];
    for (k = 0; k < 3; k++) { minint[k] = (int)get_u32(p); p += 4; }
    for (k = 0; k < 3; k++) { maxint[k] = (int)get_u32(p); p += 4; }
    unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
    for (k = 0; k < 3; k++) sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
        for (k = 0; k < 3; k++) bitsizeint[k] = (unsigned)sizeofint((int)sizeint[k]);
        bitsize = 0;
    } else {
        bitsize = sizeofints(3, sizeint);
    }
    int smallidx = (int)get_u32(p); p += 4;
    if (smallidx < FIRSTIDX || smallidx > LASTIDX) return 0;
    int smaller = magicints[smallidx - 1 > FIRSTIDX ? smallidx - 1 : FIRSTIDX] / 2;
    int smallnum = magicints[smallidx] / 2;
    unsigned sizesmall[3];
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    p += 4;                                           /* nbytes (already used for the record length) */
    bitbuf b = {(unsigned char*)p, 0, 0, 0};
    const float inv_precision = (float)(1.0 / precision);
    float* lfp = xyz_out;
    int run = 0, prevcoord[3], thiscoord[3];
    i = 0;
#define EMIT(c)                                                                \
    do {                                                                       \
        for (k = 0; k < 3; k++) {                                              \
            float v = (float)(c)[k] * inv_precision;                           \
            if (scale != 0.0f) v = v * scale;                                  \
            *lfp++ = v;                                                        \
        }                                                                      \
    } while (0)
    while (i < natoms) {
        if (bitsize == 0) {
            for (k = 0; k < 3; k++) thiscoord[k] = (int)receivebits(&b, (int)bitsizeint[k]);
        } else {
            receiveints(&b, 3, bitsize, sizeint, thiscoord);
        }
        i++;
        for (k = 0; k < 3; k++) { thiscoord[k] += minint[k]; prevcoord[k] = thiscoord[k]; }
        const int flag = (int)receivebits(&b, 1);
        int is_smaller = 0;
        if (flag == 1) {
            run = (int)receivebits(&b, 5);
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller--;
        }
        if (run > 0) {
            if (i + run / 3 > natoms) return 0;       /* corrupt stream */
            for (int r = 0; r < run; r += 3) {
                int small[3];
                receiveints(&b, 3, smallidx, sizesmall, small);
                i++;
                for (k = 0; k < 3; k++) small[k] += prevcoord[k] - smallnum;
                if (r == 0) {                         /* the first two atoms of a run were stored in swapped order */
                    EMIT(small);
                    EMIT(thiscoord);
                    for (k = 0; k < 3; k++) prevcoord[k] = small[k];
                } else {
                    EMIT(small);
                    for (k = 0; k < 3; k++) prevcoord[k] = small[k];
                }
            }
        } else {
            EMIT(thiscoord);
        }
        smallidx += is_smaller;
        if (smallidx < FIRSTIDX || smallidx > LASTIDX) return 0;
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > FIRSTIDX ? magicints[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = magicints[smallidx] / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    }
#undef EMIT
    return len;
}

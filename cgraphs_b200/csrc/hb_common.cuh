// hb_common.cuh - device-side data views, the exact DIST / ANG predicates, pair enumeration and the
// bond hash table shared by the general (multi-kernel) and the fused (one CTA per frame) paths.
#pragma once
#include "../../include/hbgpu.h"
#include <cuda_runtime.h>
#include <cstdint>

// Debug build (-DHB_DEBUG, cgraphs_b200/build.py --debug -> libhbgpu_debug.so): bounds checks on the indices that go into
// shared-memory and global scratch arrays. A violated check prints where it happened and traps the kernel (the host
// then fails with the CUDA error) - the stand-in for compute-sanitizer, which is closed on this GPU pool.
#ifdef HB_DEBUG
#include <cstdio>
#define HB_DASSERT(cond)                                                                                              \
    do {                                                                                                                 \
        if (!(cond)) {                                                                                                   \
            printf("HB_DASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
            __trap();                                                                                                    \
        }                                                                                                                \
    } while (0)
#else
#define HB_DASSERT(cond) do { } while (0)
#endif

namespace {


// ------------------------------------------------------------------------------------------------
// device-side views
// ------------------------------------------------------------------------------------------------
struct DevTopo {
    int n_systems;
    const long long* sys_atom_off;
    const int* sys_sel_off;
    const int* sys_wat_off;
    const int* sel_atom;
    const int* sel_don;
    const int* sel_acc;
    const int* sel_wat;
    const unsigned char* sel_bb;
    const int* wat_atom;
    const int* wat_ent;
    const int* ent_group;
    const int* ent_resid;
    const int* ent_residue;
    const int* ent_h_off;
    const int* ent_h_atom;
    // packed copies built by hb_set_topology (one 16-byte load instead of four 4-byte loads)
    const int4* sel_pack;    // {sel_don, sel_acc, sel_wat, sel_bb}
    const int4* ent_pack;    // {ent_resid, ent_residue, ent_h_off, hydrogen count}
    // inline hydrogen lists: four signed bytes, hydrogen atom = heavy atom + byte (-128 = none)
    const int2* sel_h;       // {donor entry's hydrogens, water entry's hydrogens} of a selection point
    const int* wat_h;        // hydrogens of a water point's entry
    int h_wide;              // some entry has more than 4 hydrogens or one further than 127 atoms away: use the CSR
};

struct Box6 {           // lower-triangular lattice: a = (ax, 0, 0), b = (bx, by, 0), c = (cx, cy, cz)
    float ax, bx, by, cx, cy, cz;
};

struct PbcGrid {        // minimum-image modes: grid in fractional coordinates. Per lattice axis either a window
    Box6 box;           // around the selection (non-periodic, `per` = 0) or the whole period (wraps, `per` = 1)
    float ia, ib, ic;   // 1 / ax, 1 / by, 1 / cz
    float c[3];         // window centre (absolute fractional coordinate)
    float half[3];      // window half width (0.5 for a periodic axis)
    float scale[3];     // cells per fractional unit
    int per[3];
};

struct FrameGrid {      // geometry of one cell grid of one frame
    float ox, oy, oz;   // origin                       (non-periodic mode)
    float inv;          // 1 / cell edge                (non-periodic mode)
    int nx, ny, nz;
    int ncell;
    PbcGrid p;          // minimum-image modes only
};

struct FrameState {     // per frame slot of a batch
    int bb_min[3];      // ordered-int encoded float min / max of the selection points
    int bb_max[3];
    int n_local;        // local waters appended so far
    int n_points;       // points in the pair-search set (selection + local waters)
    FrameGrid g1;       // selection grid used by the local-water test (cell >= around_radius)
    FrameGrid g2;       // pair-search grid (cell >= distance)
    float lo[3], hi[3]; // decoded bounding box of the selection points (minimum-image modes: fractional,
                        // relative to the first selection point and wrapped into [-0.5, 0.5])
};

struct BatchView {
    const float* xyz;        // coordinates of the batch
    const float* box;        // minimum-image modes: 9 floats per frame of the batch (lattice vectors as rows)
    long long frame0;        // global index (occupancy column) of the first frame of the batch
    int n_frames;            // frames in this batch
    long long atoms_per_frame;   // trajectory mode stride (atoms)
    long long atom_base;     // structure_batch: sys_atom_off of the first system of the batch
    int structure_batch;
    int max_sel, max_wat;    // per-frame capacities of the scratch arrays
    int cell_cap;            // cells per grid per frame (capacity)
    FrameState* fs;
    float4* selpos;          // [n_frames][max_sel]
    float4* sorted1;         // [n_frames][max_sel]          selection points in g1 cell order
    float4* lw;              // [n_frames][max_wat]          local waters (unordered)
    float4* sorted2;         // [n_frames][max_sel+max_wat]  pair-search points in g2 cell order
    int* cells1;             // [n_frames][cell_cap+1]
    int* cells2;             // [n_frames][cell_cap+1]
};

struct RunParams {
    double r2;               // fl(distance*distance)
    double around2;          // fl(around*around)
    float r2f_hi;            // float prefilter bounds (slightly above the double thresholds)
    float around2f_hi;
    float cell1;             // cell edge of g1
    float cell2;             // cell edge of g2
    float around_f;          // around_radius rounded up a little (bounding-box dilation)
    float cos_threshold;
    int check_angle;
    int excl_bb;
    int method;
    int self_pairs;          // 1: distance-0 entry pairs of one atom can pass (only with cut_angle >= 180 deg)
    int pbc;                 // HB_PBC_*
};

struct BondTable {
    unsigned long long* keys;   // open addressing, EMPTY = ~0
    unsigned int* bits;         // [capacity][wpr]
    unsigned int mask;          // capacity - 1
    int wpr;                    // 32-bit words per row
    int* n_keys;
    int* overflow;
    int* fused_overflow;        // a frame exceeded the shared-memory capacities of the fused kernel
    unsigned long long* n_hits;
    unsigned long long* phase_cycles;   // optional (HBGPU_TRACE): SM cycles thread 0 of each fused CTA spent per phase
    // water-wire mode (hb_wires.cuh): the search appends the H-bond entry pairs of every frame to a per-frame edge
    // list instead of setting table bits; k_wire_bfs consumes the lists and writes wire lengths into the table rows.
    // One pointer (null outside wire mode) so that the H-bond path's kernels carry nothing else for it.
    const struct EdgeSink* sink;
};

struct EdgeSink {               // lives in device memory, rewritten by the host before every search round
    int2* edges;                // [frames of the round][edge_cap]
    int* edge_cnt;              // [frames of the round]
    int edge_cap;
    int pad;
    long long edge_frame0;      // global frame of list 0
    int* wire_overflow;         // some list ran out of space: the host doubles edge_cap and re-runs the batch
};

#define EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
// tag bits in the id a point carries in the w component of its float4
#define WATER_TAG 0x40000000      // water point (id = index into the system's water points)
#define SELF_TAG 0x20000000       // selection point with more than one entry: has distance-0 entry pairs
#define POINT_ID_MASK 0x1fffffff

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__device__ __forceinline__ const float* frame_xyz(const BatchView& bv, const DevTopo& tp, int b, int& sys) {
    if (bv.structure_batch) {
        sys = (int)(bv.frame0 + b);
        return bv.xyz + 3ll * (tp.sys_atom_off[sys] - bv.atom_base);
    }
    sys = 0;
    return bv.xyz + 3ll * bv.atoms_per_frame * b;
}

__device__ __forceinline__ int cell_coord(float x, float o, float inv, int n) {
    int c = (int)floorf((x - o) * inv);
    return min(max(c, 0), n - 1);
}

__device__ __forceinline__ Box6 load_box(const float* b9) {
    Box6 B;
    B.ax = b9[0]; B.bx = b9[3]; B.by = b9[4]; B.cx = b9[6]; B.cy = b9[7]; B.cz = b9[8];
    return B;
}
// fractional coordinates (not wrapped) in the lower-triangular lattice
__device__ __forceinline__ void frac_coords(const PbcGrid& p, float x, float y, float z, float& sa, float& sb, float& sc) {
    sc = z * p.ic;
    sb = (y - sc * p.box.cy) * p.ib;
    sa = (x - sb * p.box.bx - sc * p.box.cx) * p.ia;
}
// cell coordinates in a minimum-image grid; false when the point lies outside a windowed axis
__device__ __forceinline__ bool cell3_pbc(const FrameGrid& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    float s[3];
    frac_coords(g.p, x, y, z, s[0], s[1], s[2]);
    const int n[3] = {g.nx, g.ny, g.nz};
    int c[3];
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float u = s[k] - g.p.c[k];
        u -= rintf(u);
        const float v = u + g.p.half[k];
        inside = inside && (g.p.per[k] || (v >= 0.f && v <= 2.f * g.p.half[k]));
        c[k] = min(max((int)floorf(v * g.p.scale[k]), 0), n[k] - 1);
    }
    cx = c[0]; cy = c[1]; cz = c[2];
    return inside;
}
template <bool PBC>
__device__ __forceinline__ void cell3(const FrameGrid& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    if constexpr (PBC) {
        cell3_pbc(g, x, y, z, cx, cy, cz);
    } else {
        cx = cell_coord(x, g.ox, g.inv, g.nx);
        cy = cell_coord(y, g.oy, g.inv, g.ny);
        cz = cell_coord(z, g.oz, g.inv, g.nz);
    }
}
template <bool PBC>
__device__ __forceinline__ int cell_of(const FrameGrid& g, float x, float y, float z) {
    int cx, cy, cz;
    cell3<PBC>(g, x, y, z, cx, cy, cz);
    return (cz * g.ny + cy) * g.nx + cx;
}
// neighbour cells along one axis of a minimum-image grid, each listed once
struct AxisList { int v[3]; int n; };
__device__ __forceinline__ AxisList axis_neighbors(int c, int n, int per) {
    AxisList a;
    if (!per) {
        const int lo = max(c - 1, 0), hi = min(c + 1, n - 1);
        a.n = hi - lo + 1;
        a.v[0] = lo; a.v[1] = lo + 1; a.v[2] = lo + 2;
    } else if (n >= 3) {
        a.n = 3;
        a.v[0] = c ? c - 1 : n - 1; a.v[1] = c; a.v[2] = c + 1 < n ? c + 1 : 0;
    } else {
        a.n = n;
        a.v[0] = 0; a.v[1] = 1; a.v[2] = 0;
    }
    return a;
}
// the x-neighbours of a row as one or two runs of contiguous cells [lo, hi]
struct XSegs { int lo[2], hi[2]; int n; };
__device__ __forceinline__ XSegs x_segments(int c, int n, int per) {
    XSegs x;
    x.n = 1;
    x.lo[1] = 0; x.hi[1] = -1;
    if (!per) { x.lo[0] = max(c - 1, 0); x.hi[0] = min(c + 1, n - 1); }
    else if (n < 3) { x.lo[0] = 0; x.hi[0] = n - 1; }
    else if (c == 0) { x.lo[0] = 0; x.hi[0] = 1; x.lo[1] = n - 1; x.hi[1] = n - 1; x.n = 2; }
    else if (c == n - 1) { x.lo[0] = n - 2; x.hi[0] = n - 1; x.lo[1] = 0; x.hi[1] = 0; x.n = 2; }
    else { x.lo[0] = c - 1; x.hi[0] = c + 1; }
    return x;
}

// DIST predicate, SURVEY A.2 / F2 (scipy cKDTree semantics on float32 inputs)
__device__ __forceinline__ bool dist_le(float ax, float ay, float az, float bx, float by, float bz, double r2) {
    double dx = (double)ax - (double)bx;
    double dy = (double)ay - (double)by;
    double dz = (double)az - (double)bz;
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return s <= r2;
}
__device__ __forceinline__ float dist2f(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return dx * dx + dy * dy + dz * dz;
}
// Minimum-image DIST (opt-in extension): the float64 difference vector is reduced along c, b, a in
// turn (n = rint(d_k / L_k), individually rounded multiply and subtract), then tested like DIST. When no
// reduction applies (n = 0 three times) the arithmetic is exactly that of dist_le.
__device__ __forceinline__ bool dist_le_pbc(float ax, float ay, float az, float bx, float by, float bz, double r2, const Box6& B) {
    double dx = (double)ax - (double)bx;
    double dy = (double)ay - (double)by;
    double dz = (double)az - (double)bz;
    double n = rint(__ddiv_rn(dz, (double)B.cz));
    dz = __dsub_rn(dz, __dmul_rn(n, (double)B.cz));
    dy = __dsub_rn(dy, __dmul_rn(n, (double)B.cy));
    dx = __dsub_rn(dx, __dmul_rn(n, (double)B.cx));
    n = rint(__ddiv_rn(dy, (double)B.by));
    dy = __dsub_rn(dy, __dmul_rn(n, (double)B.by));
    dx = __dsub_rn(dx, __dmul_rn(n, (double)B.bx));
    n = rint(__ddiv_rn(dx, (double)B.ax));
    dx = __dsub_rn(dx, __dmul_rn(n, (double)B.ax));
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return s <= r2;
}
// float32 minimum image of a difference vector, individually rounded (ANG under PBC; also the pre-filter)
__device__ __forceinline__ void min_image_f32(float& vx, float& vy, float& vz, const Box6& B) {
    float n = rintf(__fdiv_rn(vz, B.cz));
    vz = __fsub_rn(vz, __fmul_rn(n, B.cz));
    vy = __fsub_rn(vy, __fmul_rn(n, B.cy));
    vx = __fsub_rn(vx, __fmul_rn(n, B.cx));
    n = rintf(__fdiv_rn(vy, B.by));
    vy = __fsub_rn(vy, __fmul_rn(n, B.by));
    vx = __fsub_rn(vx, __fmul_rn(n, B.bx));
    n = rintf(__fdiv_rn(vx, B.ax));
    vx = __fsub_rn(vx, __fmul_rn(n, B.ax));
}
__device__ __forceinline__ float dist2f_pbc(float ax, float ay, float az, float bx, float by, float bz, const Box6& B,
                                            float ia, float ib, float ic) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float n = rintf(dz * ic);
    dz -= n * B.cz; dy -= n * B.cy; dx -= n * B.cx;
    n = rintf(dy * ib);
    dy -= n * B.by; dx -= n * B.bx;
    n = rintf(dx * ia);
    dx -= n * B.ax;
    return dx * dx + dy * dy + dz * dz;
}
template <bool PBC>
__device__ __forceinline__ bool within(float ax, float ay, float az, float bx, float by, float bz, float r2f_hi, double r2,
                                       const PbcGrid& p) {
    if constexpr (PBC) {
        if (dist2f_pbc(ax, ay, az, bx, by, bz, p.box, p.ia, p.ib, p.ic) > r2f_hi) return false;
        return dist_le_pbc(ax, ay, az, bx, by, bz, r2, p.box);
    } else {
        if (dist2f(ax, ay, az, bx, by, bz) > r2f_hi) return false;
        return dist_le(ax, ay, az, bx, by, bz, r2);
    }
}

// ANG predicate, SURVEY A.2 / F3 (helpfunctions.py:96-104,155): b = pair partner, h = hydrogen, a = other partner
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
template <bool PBC>
__device__ __forceinline__ bool ang_ok_exact(float bx, float by, float bz, float hx, float hy, float hz, float ax,
                                             float ay, float az, float cthr, const Box6& B) {
    float v1x = __fsub_rn(hx, bx), v1y = __fsub_rn(hy, by), v1z = __fsub_rn(hz, bz);
    float v2x = __fsub_rn(ax, hx), v2y = __fsub_rn(ay, hy), v2z = __fsub_rn(az, hz);
    if constexpr (PBC) {
        min_image_f32(v1x, v1y, v1z, B);
        min_image_f32(v2x, v2y, v2z, B);
    }
    float d12 = dot3_rn(v1x, v1y, v1z, v2x, v2y, v2z);
    float d11 = dot3_rn(v1x, v1y, v1z, v1x, v1y, v1z);
    float d22 = dot3_rn(v2x, v2y, v2z, v2x, v2y, v2z);
    float den = __fmul_rn(__fsqrt_rn(d11), __fsqrt_rn(d22));
    float c = __fdiv_rn(d12, den);
    if (c < -1.0f) c = -1.0f;
    return (c >= cthr) && (c <= 1.0f);   // NaN fails both
}
// Same decision, cheaper on average: an approximate cosine (rsqrt, FMA allowed; relative error < 1e-6)
// settles every case that is further than 1e-5 from both ends of the accepted interval [cthr, 1];
// only the rest (and NaN / degenerate vectors) runs the exact, individually rounded chain above.
template <bool PBC>
__device__ __forceinline__ bool ang_ok(float bx, float by, float bz, float hx, float hy, float hz, float ax,
                                       float ay, float az, float cthr, const Box6& B) {
    float v1x = hx - bx, v1y = hy - by, v1z = hz - bz;
    float v2x = ax - hx, v2y = ay - hy, v2z = az - hz;
    if constexpr (PBC) {
        min_image_f32(v1x, v1y, v1z, B);
        min_image_f32(v2x, v2y, v2z, B);
    }
    const float d12 = v1x * v2x + v1y * v2y + v1z * v2z;
    const float d11 = v1x * v1x + v1y * v1y + v1z * v1z;
    const float d22 = v2x * v2x + v2y * v2y + v2z * v2z;
    const float ca = d12 * rsqrtf(d11 * d22);
    if (ca < cthr - 1e-5f) return false;
    if (ca > cthr + 1e-5f && ca < 1.0f - 1e-5f) return true;
    return ang_ok_exact<PBC>(bx, by, bz, hx, hy, hz, ax, ay, az, cthr, B);
}

__device__ void make_grid(FrameGrid& g, const float* lo, const float* hi, float pad, float cell, int cap) {
    float ext[3];
    bool finite = true;
    for (int k = 0; k < 3; ++k) {
        ext[k] = (hi[k] - lo[k]) + 2.f * pad;
        finite = finite && isfinite(ext[k]) && ext[k] >= 0.f;
    }
    g.ox = lo[0] - pad; g.oy = lo[1] - pad; g.oz = lo[2] - pad;
    if (!finite) { g.inv = 0.f; g.nx = g.ny = g.nz = 1; g.ncell = 1; g.ox = g.oy = g.oz = 0.f; return; }
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        int n[3];
        double tot = 1.0;
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            float q = floorf(ext[k] / cell) + 1.f;
            if (q > 1024.f) ok = false;
            n[k] = (int)fminf(q, 1024.f);
            tot *= (double)n[k];
        }
        if (ok && tot <= (double)cap) {
            g.nx = n[0]; g.ny = n[1]; g.nz = n[2]; g.ncell = n[0] * n[1] * n[2];
            g.inv = 1.0f / cell;
            return;
        }
        cell *= 1.2f;
    }
    g.inv = 0.f; g.nx = g.ny = g.nz = 1; g.ncell = 1;
}

// Minimum-image grid. sR = fractional coordinates of the reference point (first selection point), [umin, umax] =
// fractional bounding box of the selection relative to it (each coordinate wrapped into [-0.5, 0.5]), pad = real
// distance every point of interest may lie outside that box. An axis whose padded box stays below 0.9 periods
// gets a window (no wrap, points outside are of no interest), otherwise the whole period with wrap-around.
// Cell thickness (distance between opposite cell faces) is >= `cell` on every axis.
__device__ void make_grid_pbc(FrameGrid& g, const Box6& B, const float* sR, const float* umin, const float* umax,
                              float pad, float cell, int cap) {
    PbcGrid& p = g.p;
    p.box = B;
    p.ia = 1.0f / B.ax; p.ib = 1.0f / B.by; p.ic = 1.0f / B.cz;
    const float bcx = B.by * B.cz, bcy = -B.bx * B.cz, bcz = B.bx * B.cy - B.by * B.cx;
    float w[3];                                    // perpendicular widths of the cell
    w[0] = B.ax * B.by * B.cz / sqrtf(bcx * bcx + bcy * bcy + bcz * bcz);
    w[1] = B.by * B.cz / sqrtf(B.cz * B.cz + B.cy * B.cy);
    w[2] = B.cz;
    for (int k = 0; k < 3; ++k) {
        const float padf = pad / w[k] * 1.001f + 1e-6f;
        const float width = (umax[k] - umin[k]) + 2.f * padf;
        if (width < 0.9f && isfinite(sR[k])) {
            p.per[k] = 0;
            p.c[k] = sR[k] + 0.5f * (umin[k] + umax[k]);
            p.half[k] = 0.5f * width;
        } else {
            p.per[k] = 1;
            p.c[k] = 0.5f;
            p.half[k] = 0.5f;
        }
    }
    int n[3] = {1, 1, 1};
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
        double tot = 1.0;
        bool ok = true;
        for (int k = 0; k < 3; ++k) {
            float q = floorf(2.f * p.half[k] * w[k] / cell);
            if (!(q >= 1.f)) q = 1.f;
            if (q > 1024.f) { ok = false; q = 1024.f; }
            n[k] = (int)q;
            tot *= (double)n[k];
        }
        if (ok && tot <= (double)cap) break;
        cell *= 1.2f;
        if (it == 63) n[0] = n[1] = n[2] = 1;
    }
    for (int k = 0; k < 3; ++k) p.scale[k] = (float)n[k] / (2.f * p.half[k]);
    g.nx = n[0]; g.ny = n[1]; g.nz = n[2]; g.ncell = n[0] * n[1] * n[2];
    g.ox = g.oy = g.oz = 0.f; g.inv = 0.f;
}

// ------------------------------------------------------------------------------------------------
// bond table: insert key, set bit
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

__device__ __forceinline__ void record_bond(const BondTable& bt, unsigned long long key, long long frame) {
    unsigned h = hash64(key) & bt.mask;
    for (unsigned probe = 0; probe <= bt.mask; ++probe) {
        unsigned long long cur = bt.keys[h];
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&bt.keys[h], EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { atomicAdd(bt.n_keys, 1); cur = key; }
        }
        if (cur == key) {
            HB_DASSERT(frame >= 0 && (frame >> 5) < bt.wpr && h <= bt.mask);
            atomicOr(&bt.bits[(size_t)h * bt.wpr + (frame >> 5)], 1u << (frame & 31));
            return;
        }
        h = (h + 1) & bt.mask;
    }
    atomicExch(bt.overflow, 1);
}

// record_bond with the first probe already loaded (`cur` = bt.keys[h]): the fused kernel requests the probes of several
// bonds together before it handles them
__device__ __forceinline__ void record_bond_probed(const BondTable& bt, unsigned long long key, long long frame, unsigned h,
                                                   unsigned long long cur) {
    for (unsigned probe = 0; probe <= bt.mask; ++probe) {
        if (probe) cur = bt.keys[h];
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&bt.keys[h], EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { atomicAdd(bt.n_keys, 1); cur = key; }
        }
        if (cur == key) {
            HB_DASSERT(frame >= 0 && (frame >> 5) < bt.wpr && h <= bt.mask);
            atomicOr(&bt.bits[(size_t)h * bt.wpr + (frame >> 5)], 1u << (frame & 31));
            return;
        }
        h = (h + 1) & bt.mask;
    }
    atomicExch(bt.overflow, 1);
}

// ------------------------------------------------------------------------------------------------
// K7: pair search
// ------------------------------------------------------------------------------------------------
struct PointInfo {
    int don, acc, wat;   // global entry indices or -1
    int bb;
    float x, y, z;
    int atom;            // atom index inside the frame
    int hd, hw;          // packed hydrogen offsets of the donor / water entry (complete unless tp.h_wide)
};
#define H_NONE4 ((int)0x80808080)

// everything the pair expansion needs to know about a point: independent loads, one round trip
__device__ __forceinline__ void load_info(const DevTopo& tp, const RunParams& rp, int sys, float4 p, PointInfo& o) {
    int id = __float_as_int(p.w);
    o.x = p.x; o.y = p.y; o.z = p.z;
    if (id & WATER_TAG) {
        const int w = tp.sys_wat_off[sys] + (id & POINT_ID_MASK);
        o.don = -1; o.acc = -1; o.bb = 0;
        o.wat = tp.wat_ent[w];
        o.atom = tp.wat_atom[w];
        o.hd = H_NONE4;
        o.hw = tp.wat_h[w];
    } else {
        const int s = tp.sys_sel_off[sys] + (id & POINT_ID_MASK);
        const int4 k = tp.sel_pack[s];
        const bool wa = rp.method == HB_METHOD_WATER_AROUND;
        o.don = k.x;
        o.acc = k.y;
        o.wat = wa ? k.z : -1;
        o.bb = k.w;
        o.atom = tp.sel_atom[s];
        const int2 h = tp.sel_h[s];
        o.hd = h.x;
        o.hw = h.y;
    }
}

// any hydrogen of entry `e` (owned by point `owner`) with ANG(other, h, owner) satisfied: CSR walk
template <bool PBC>
__device__ __forceinline__ bool any_h(const DevTopo& tp, const float* xyz, int e, const PointInfo& owner,
                                      const PointInfo& other, float cthr, const Box6& B) {
    int4 ep = tp.ent_pack[e];
    int s = ep.z, t = ep.z + ep.w;
    for (int k = s; k < t; ++k) {
        const float* h = xyz + 3ll * tp.ent_h_atom[k];
        if (ang_ok<PBC>(other.x, other.y, other.z, h[0], h[1], h[2], owner.x, owner.y, owner.z, cthr, B)) return true;
    }
    return false;
}
// same over an inline list of up to four hydrogens of the atom `atom` at (ox, oy, oz); (tx, ty, tz) is the
// partner. The coordinate loads are issued together. Not inlined: it is called from up to four places per
// pair and would otherwise multiply the register footprint of the calling kernels.
template <bool PBC>
__device__ __noinline__ bool any_h4(const float* xyz, int hl, int atom, float ox, float oy, float oz, float tx, float ty,
                                    float tz, float cthr, const Box6* B) {
    float h[4][3];
    int d[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        d[k] = (int)(signed char)((unsigned)hl >> (8 * k));
        if (d[k] != -128) {
            const float* p = xyz + 3ll * (atom + d[k]);
            h[k][0] = p[0]; h[k][1] = p[1]; h[k][2] = p[2];
        }
    }
    bool ok = false;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (d[k] != -128 && !ok) ok = ang_ok<PBC>(tx, ty, tz, h[k][0], h[k][1], h[k][2], ox, oy, oz, cthr, *B);
    return ok;
}

__device__ __noinline__ void emit_bond(const int4* ent_pack, const int* ent_group, int check_angle, BondTable bt,
                                       long long frame, int e0, int e1, int* nhit) {
    if (bt.sink) {                                     // wire mode: (acceptor, donor) / (selection, water) / (water, water)
        ++*nhit;
        const EdgeSink sk = *bt.sink;
        const int b = (int)(frame - sk.edge_frame0);
        HB_DASSERT(b >= 0);
        const int i = atomicAdd(&sk.edge_cnt[b], 1);
        if (i < sk.edge_cap) sk.edges[(size_t)b * sk.edge_cap + i] = make_int2(e0, e1);
        else atomicExch(sk.wire_overflow, 1);
        return;
    }
    int x = min(e0, e1), y = max(e0, e1);
    int4 px = ent_pack[x], py = ent_pack[y];
    int gx = ent_group[x], gy = ent_group[y];
    bool chk = px.x < py.x;                            // hbond.py:120-122
    if (!check_angle && px.y == py.y) return;          // hbond.py:125-127
    unsigned long long key = ((unsigned long long)(unsigned)(chk ? gx : gy) << 32) | (unsigned long long)(unsigned)(chk ? gy : gx);
    ++*nhit;
    record_bond(bt, key, frame);
}
__device__ __forceinline__ void emit(const DevTopo& tp, const RunParams& rp, const BondTable& bt, long long frame,
                                     int e0, int e1, int& nhit) {
    emit_bond(tp.ent_pack, tp.ent_group, rp.check_angle, bt, frame, e0, e1, &nhit);
}
// All entry pairs generated by the point pair (P, Q), P != Q as points.
// "Some hydrogen of this entry passes the angle test" is evaluated once per entry that takes part in a pair.
template <bool PBC>
__device__ __forceinline__ void process_pair(const DevTopo& tp, const RunParams& rp, const BondTable& bt, const float* xyz,
                                             long long frame, const PointInfo& P, const PointInfo& Q, int& nhit, const Box6& B) {
    const bool ang = rp.check_angle != 0;
    const float ct = rp.cos_threshold;
    const bool bb_skip = rp.excl_bb && P.bb && Q.bb;
    const bool need_pd = P.don >= 0 && ((!bb_skip && Q.acc >= 0) || Q.wat >= 0);
    const bool need_qd = Q.don >= 0 && ((!bb_skip && P.acc >= 0) || P.wat >= 0);
    const bool need_pw = P.wat >= 0;
    const bool need_qw = Q.wat >= 0;
    bool pd = need_pd, qd = need_qd, pw = need_pw, qw = need_qw;
    if (ang) {
        if (tp.h_wide) {
            if (need_pd) pd = any_h<PBC>(tp, xyz, P.don, P, Q, ct, B);
            if (need_qd) qd = any_h<PBC>(tp, xyz, Q.don, Q, P, ct, B);
            if (need_pw) pw = any_h<PBC>(tp, xyz, P.wat, P, Q, ct, B);
            if (need_qw) qw = any_h<PBC>(tp, xyz, Q.wat, Q, P, ct, B);
        } else {
            if (need_pd) pd = any_h4<PBC>(xyz, P.hd, P.atom, P.x, P.y, P.z, Q.x, Q.y, Q.z, ct, &B);
            if (need_qd) qd = any_h4<PBC>(xyz, Q.hd, Q.atom, Q.x, Q.y, Q.z, P.x, P.y, P.z, ct, &B);
            if (need_pw) pw = any_h4<PBC>(xyz, P.hw, P.atom, P.x, P.y, P.z, Q.x, Q.y, Q.z, ct, &B);
            if (need_qw) qw = any_h4<PBC>(xyz, Q.hw, Q.atom, Q.x, Q.y, Q.z, P.x, P.y, P.z, ct, &B);
        }
    }
    if (!bb_skip) {
        if (P.acc >= 0 && Q.don >= 0 && qd) emit(tp, rp, bt, frame, P.acc, Q.don, nhit);
        if (Q.acc >= 0 && P.don >= 0 && pd) emit(tp, rp, bt, frame, Q.acc, P.don, nhit);
    }
    if (P.wat >= 0 && Q.wat >= 0 && (pw || qw)) emit(tp, rp, bt, frame, P.wat, Q.wat, nhit);
    if (Q.wat >= 0) {
        if (P.don >= 0 && (pd || qw)) emit(tp, rp, bt, frame, P.don, Q.wat, nhit);
        if (P.acc >= 0 && qw) emit(tp, rp, bt, frame, P.acc, Q.wat, nhit);
    }
    if (P.wat >= 0) {
        if (Q.don >= 0 && (qd || pw)) emit(tp, rp, bt, frame, Q.don, P.wat, nhit);
        if (Q.acc >= 0 && pw) emit(tp, rp, bt, frame, Q.acc, P.wat, nhit);
    }
}

// The same expansion in three separable steps (used by the fused kernel, which runs the angle tests as
// one dense task list between steps one and three).
//   rules bit r: entry pair r exists.       r0 (P.acc,Q.don)  r1 (Q.acc,P.don)  r2 (P.wat,Q.wat)  r3 (P.don,Q.wat)
//                                            r4 (P.acc,Q.wat)  r5 (Q.don,P.wat)  r6 (Q.acc,P.wat)
//   flags bit f: some hydrogen of entry f passes the angle test.   f0 P.don  f1 Q.don  f2 P.wat  f3 Q.wat
__device__ __forceinline__ unsigned pair_rules(const PointInfo& P, const PointInfo& Q, bool excl_bb) {
    const bool bb_skip = excl_bb && P.bb && Q.bb;
    unsigned r = 0;
    if (!bb_skip && P.acc >= 0 && Q.don >= 0) r |= 1u;
    if (!bb_skip && Q.acc >= 0 && P.don >= 0) r |= 2u;
    if (P.wat >= 0 && Q.wat >= 0) r |= 4u;
    if (Q.wat >= 0 && P.don >= 0) r |= 8u;
    if (Q.wat >= 0 && P.acc >= 0) r |= 16u;
    if (P.wat >= 0 && Q.don >= 0) r |= 32u;
    if (P.wat >= 0 && Q.acc >= 0) r |= 64u;
    return r;
}
__device__ __forceinline__ unsigned rules_need(unsigned r) {      // flags that decide some existing rule
    unsigned n = 0;
    if (r & (2u | 8u)) n |= 1u;
    if (r & (1u | 32u)) n |= 2u;
    if (r & (4u | 32u | 64u)) n |= 4u;
    if (r & (4u | 8u | 16u)) n |= 8u;
    return n;
}
__device__ __forceinline__ unsigned rules_pass(unsigned f) {      // rules satisfied by the flags
    const bool pd = f & 1u, qd = f & 2u, pw = f & 4u, qw = f & 8u;
    return (qd ? 1u : 0u) | (pd ? 2u : 0u) | ((pw || qw) ? 4u : 0u) | ((pd || qw) ? 8u : 0u) | (qw ? 16u : 0u) |
           ((qd || pw) ? 32u : 0u) | (pw ? 64u : 0u);
}
__device__ __forceinline__ void rule_entries(const PointInfo& P, const PointInfo& Q, int r, int& e0, int& e1) {
    switch (r) {
        case 0: e0 = P.acc; e1 = Q.don; break;
        case 1: e0 = Q.acc; e1 = P.don; break;
        case 2: e0 = P.wat; e1 = Q.wat; break;
        case 3: e0 = P.don; e1 = Q.wat; break;
        case 4: e0 = P.acc; e1 = Q.wat; break;
        case 5: e0 = Q.don; e1 = P.wat; break;
        default: e0 = Q.acc; e1 = P.wat; break;
    }
}
__device__ __forceinline__ int h_count(int hl) {                  // bytes of the packed list that are not -128
    const unsigned x = (unsigned)hl ^ 0x80808080u;                // zero byte <=> none
    const unsigned nz = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // high bit set <=> byte non-zero
    return __popc(nz);
}

// Entry pairs of a point with itself (distance 0): acceptor-donor and selection-water entries of one atom
template <bool PBC>
__device__ void process_self(const DevTopo& tp, const RunParams& rp, const BondTable& bt, const float* xyz,
                             long long frame, const PointInfo& P, int& nhit, const Box6& B) {
    const bool ang = rp.check_angle != 0;
    const float ct = rp.cos_threshold;
    bool need = (P.acc >= 0 && P.don >= 0) || (P.wat >= 0 && (P.don >= 0 || P.acc >= 0));
    if (!need) return;
    bool pd = P.don >= 0 ? ((!ang) || any_h<PBC>(tp, xyz, P.don, P, P, ct, B)) : false;
    bool pw = P.wat >= 0 ? ((!ang) || any_h<PBC>(tp, xyz, P.wat, P, P, ct, B)) : false;
    if (P.acc >= 0 && P.don >= 0 && !(rp.excl_bb && P.bb) && pd) emit(tp, rp, bt, frame, P.acc, P.don, nhit);
    if (P.wat >= 0) {
        if (P.don >= 0 && (pd || pw)) emit(tp, rp, bt, frame, P.don, P.wat, nhit);
        if (P.acc >= 0 && pw) emit(tp, rp, bt, frame, P.acc, P.wat, nhit);
    }
}


// ------------------------------------------------------------------------------------------------
// neighbour walks shared by both paths. `cells[c]` holds the END offset of cell c in the cell-sorted
// point array (start = cells[c-1], 0 for c = 0); cells are ordered x-fastest, so the three x-neighbours
// of a (y, z) row are one contiguous run of points.
// ------------------------------------------------------------------------------------------------
// OWN_FIRST (general path): scan the water's own row of cells first, then the eight around it - a water next to the
// selection usually finds its hit at once. (The fused kernel keeps the plain order: its candidates are pre-filtered
// by the near bitmap and the shorter loop control is worth more there.)
template <bool PBC, typename CellT, bool OWN_FIRST = false>
__device__ __forceinline__ bool water_is_local(const FrameGrid& g, const CellT* cells, const float4* pts, float x,
                                               float y, float z, const RunParams& rp) {
    int cx, cy, cz;
    cell3<PBC>(g, x, y, z, cx, cy, cz);
    if constexpr (PBC) {
        const AxisList zl = axis_neighbors(cz, g.nz, g.p.per[2]), yl = axis_neighbors(cy, g.ny, g.p.per[1]);
        const XSegs xs = x_segments(cx, g.nx, g.p.per[0]);
        for (int a = 0; a < zl.n; ++a)
            for (int b = 0; b < yl.n; ++b) {
                const int row = (zl.v[a] * g.ny + yl.v[b]) * g.nx;
                for (int k = 0; k < xs.n; ++k) {
                    const int c_lo = row + xs.lo[k], c_hi = row + xs.hi[k];
                    const int s = c_lo ? (int)cells[c_lo - 1] : 0, e = (int)cells[c_hi];
                    for (int q = s; q < e; ++q) {
                        const float4 sp = pts[q];
                        if (within<true>(x, y, z, sp.x, sp.y, sp.z, rp.around2f_hi, rp.around2, g.p)) return true;
                    }
                }
            }
        return false;
    } else {
        if constexpr (OWN_FIRST) {
            for (int kz = 0; kz < 3; ++kz)
                for (int ky = 0; ky < 3; ++ky) {
                    const int zz = cz + ((kz + 1) % 3) - 1, yy = cy + ((ky + 1) % 3) - 1;      // offsets 0, +1, -1
                    if (zz < 0 || zz >= g.nz || yy < 0 || yy >= g.ny) continue;
                    const int c_lo = (zz * g.ny + yy) * g.nx + max(cx - 1, 0);
                    const int c_hi = (zz * g.ny + yy) * g.nx + min(cx + 1, g.nx - 1);
                    const int s = c_lo ? (int)cells[c_lo - 1] : 0;
                    const int e = (int)cells[c_hi];
                    for (int q = s; q < e; ++q) {
                        const float4 sp = pts[q];
                        if (dist2f(x, y, z, sp.x, sp.y, sp.z) <= rp.around2f_hi &&
                            dist_le(x, y, z, sp.x, sp.y, sp.z, rp.around2))
                            return true;
                    }
                }
            return false;
        }
        for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.nz - 1); ++zz)
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); ++yy) {
                int c_lo = (zz * g.ny + yy) * g.nx + max(cx - 1, 0);
                int c_hi = (zz * g.ny + yy) * g.nx + min(cx + 1, g.nx - 1);
                int s = c_lo ? (int)cells[c_lo - 1] : 0;
                int e = (int)cells[c_hi];
                for (int q = s; q < e; ++q) {
                    float4 sp = pts[q];
                    if (dist2f(x, y, z, sp.x, sp.y, sp.z) <= rp.around2f_hi &&
                        dist_le(x, y, z, sp.x, sp.y, sp.z, rp.around2))
                        return true;
                }
            }
        return false;
    }
}

// Calls f(q) for every cell-sorted point q > i that shares a neighbourhood with point i: half shell (own cell
// + 13 forward cells as 5 contiguous runs) in the non-periodic grid, the full de-duplicated shell filtered
// by q > i in a minimum-image grid (wrap-around breaks the forward ordering of cells).
template <bool PBC, typename CellT, typename F>
__device__ __forceinline__ void for_each_partner(const FrameGrid& g, const CellT* cells, int i, float x, float y, float z, F f) {
    int cx, cy, cz;
    cell3<PBC>(g, x, y, z, cx, cy, cz);
    bool wraps = false;
    if constexpr (PBC) wraps = g.p.per[0] | g.p.per[1] | g.p.per[2];
    if (wraps) {
        if constexpr (PBC) {
            const AxisList zl = axis_neighbors(cz, g.nz, g.p.per[2]), yl = axis_neighbors(cy, g.ny, g.p.per[1]);
            const XSegs xs = x_segments(cx, g.nx, g.p.per[0]);
            for (int a = 0; a < zl.n; ++a)
                for (int b = 0; b < yl.n; ++b) {
                    const int row = (zl.v[a] * g.ny + yl.v[b]) * g.nx;
                    for (int k = 0; k < xs.n; ++k) {
                        const int c_lo = row + xs.lo[k], c_hi = row + xs.hi[k];
                        const int s = max(c_lo ? (int)cells[c_lo - 1] : 0, i + 1), e = (int)cells[c_hi];
                        for (int q = s; q < e; ++q) f(q);
                    }
                }
        }
    } else {      // no wrap-around on any axis (always so without PBC; windowed selection with PBC): forward half shell
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const int xl = max(cx - 1, 0), xh = min(cx + 1, nx - 1);
#pragma unroll 1
        for (int run = 0; run < 5; ++run) {
            int s, e;
            if (run == 0) {
                s = i + 1;
                e = (int)cells[(cz * ny + cy) * nx + xh];
            } else {
                const int yy = run == 1 ? cy + 1 : cy + (run - 3);
                const int zz = run == 1 ? cz : cz + 1;
                if (yy < 0 || yy >= ny || zz >= nz) continue;
                const int row = (zz * ny + yy) * nx;
                s = (row + xl) ? (int)cells[row + xl - 1] : 0;
                e = (int)cells[row + xh];
            }
            for (int q = s; q < e; ++q) f(q);
        }
    }
}

// all H-bond entry pairs of cell-sorted point i with itself and with its partners (general path)
template <bool PBC, typename CellT>
__device__ __forceinline__ void search_point(const DevTopo& tp, const RunParams& rp, const BondTable& bt,
                                             const float* xyz, long long frame, int sys, const FrameGrid& g,
                                             const CellT* cells, const float4* pts, int i, int& nhit) {
    float4 p4 = pts[i];
    PointInfo P;
    load_info(tp, rp, sys, p4, P);
    if (rp.self_pairs) process_self<PBC>(tp, rp, bt, xyz, frame, P, nhit, g.p.box);
    for_each_partner<PBC>(g, cells, i, P.x, P.y, P.z, [&](int q) {
        const float4 q4 = pts[q];
        if (!within<PBC>(P.x, P.y, P.z, q4.x, q4.y, q4.z, rp.r2f_hi, rp.r2, g.p)) return;
        PointInfo Q;
        load_info(tp, rp, sys, q4, Q);
        process_pair<PBC>(tp, rp, bt, xyz, frame, P, Q, nhit, g.p.box);
    });
}

__device__ __forceinline__ void flush_hits(const BondTable& bt, int nhit) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nhit += __shfl_xor_sync(0xffffffffu, nhit, o);
    if ((threadIdx.x & 31) == 0 && nhit) atomicAdd(bt.n_hits, (unsigned long long)nhit);
}

}  // namespace

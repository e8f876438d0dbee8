// hb_general.cuh - general path: global-memory cell lists, any system size (one launch per stage per batch).
#pragma once
#include "hb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// Frame decode for planar sources (DCD trajectory records hold X[N], Y[N], Z[N] per frame): interleave the
// atoms [first, first + len) of every frame of the batch into the [atom][3] layout the search kernels read
// (replaces the reader's decode, reference basic.py:46-47 / MDAnalysis' DCD reader)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_interleave(const unsigned char* raw, long long frame_stride, long long x_off, long long y_off,
                                                    long long z_off, long long first, long long len, long long n_atoms, float* xyz) {
    const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= len) return;
    const unsigned char* fr = raw + (size_t)blockIdx.y * frame_stride;
    const float x = reinterpret_cast<const float*>(fr + x_off)[first + a];
    const float y = reinterpret_cast<const float*>(fr + y_off)[first + a];
    const float z = reinterpret_cast<const float*>(fr + z_off)[first + a];
    float* o = xyz + ((size_t)blockIdx.y * n_atoms + first + a) * 3;
    o[0] = x; o[1] = y; o[2] = z;
}

// ------------------------------------------------------------------------------------------------
// K0: reset per-frame state
// ------------------------------------------------------------------------------------------------
__global__ void k_init_frames(BatchView bv) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bv.n_frames) return;
    FrameState& f = bv.fs[b];
    for (int k = 0; k < 3; ++k) {
        f.bb_min[k] = 0x7fffffff;
        f.bb_max[k] = (int)0x80000000;
    }
    f.n_local = 0;
    f.n_points = 0;
}

// ------------------------------------------------------------------------------------------------
// K1: gather selection points, per-frame bounding box
// ------------------------------------------------------------------------------------------------
// fractional coordinates of the selection's reference point (its first point) and the lattice of frame b
__device__ __forceinline__ void frame_lattice(const BatchView& bv, int b, PbcGrid& p) {
    p.box = load_box(bv.box + 9ll * b);
    p.ia = 1.0f / p.box.ax; p.ib = 1.0f / p.box.by; p.ic = 1.0f / p.box.cz;
}

template <bool PBC>
__global__ void __launch_bounds__(256) k_gather_sel(BatchView bv, DevTopo tp) {
    int b = blockIdx.y;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    int s0 = tp.sys_sel_off[sys], s1 = tp.sys_sel_off[sys + 1];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    bool live = (s0 + i) < s1;
    if (live) {
        const float* p = xyz + 3ll * tp.sel_atom[s0 + i];
        x = p[0]; y = p[1]; z = p[2];
        bv.selpos[(size_t)b * bv.max_sel + i] = make_float4(x, y, z, __int_as_float(i));
    }
    if constexpr (PBC) {       // bounding box in fractional coordinates relative to the first selection point
        if (live) {
            PbcGrid pg;
            frame_lattice(bv, b, pg);
            const float* r = xyz + 3ll * tp.sel_atom[s0];
            float ra, rb, rc, sa, sb, sc;
            frac_coords(pg, r[0], r[1], r[2], ra, rb, rc);
            frac_coords(pg, x, y, z, sa, sb, sc);
            x = sa - ra; y = sb - rb; z = sc - rc;
            x -= rintf(x); y -= rintf(y); z -= rintf(z);
        }
    }
    float mn[3] = {live ? x : INFINITY, live ? y : INFINITY, live ? z : INFINITY};
    float mx[3] = {live ? x : -INFINITY, live ? y : -INFINITY, live ? z : -INFINITY};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0 && mn[0] <= mx[0]) {
        FrameState& f = bv.fs[b];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&f.bb_min[k], f2ord(mn[k]));
            atomicMax(&f.bb_max[k], f2ord(mx[k]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: grid geometry. g1 covers the selection bounding box with one padding cell on every side so that
// every water inside the dilated box maps to a valid cell; g2 covers the dilated box.
// Cell edges carry a 1e-3 relative margin over the cut-off so float32 binning can never separate a
// true neighbour pair by more than one cell; grids are capped at 1024 cells per axis and at
// cell_cap cells in total by coarsening (coarser cells stay correct with a 27-cell neighbourhood).
// ------------------------------------------------------------------------------------------------
template <bool PBC>
__global__ void k_grid_setup(BatchView bv, DevTopo tp, RunParams rp) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bv.n_frames) return;
    FrameState& f = bv.fs[b];
    for (int k = 0; k < 3; ++k) {
        f.lo[k] = ord2f(f.bb_min[k]);
        f.hi[k] = ord2f(f.bb_max[k]);
    }
    if constexpr (PBC) {
        int sys;
        const float* xyz = frame_xyz(bv, tp, b, sys);
        PbcGrid pg;
        frame_lattice(bv, b, pg);
        float sR[3] = {NAN, NAN, NAN};
        if (tp.sys_sel_off[sys + 1] > tp.sys_sel_off[sys]) {
            const float* r = xyz + 3ll * tp.sel_atom[tp.sys_sel_off[sys]];
            frac_coords(pg, r[0], r[1], r[2], sR[0], sR[1], sR[2]);
        }
        if (rp.method == HB_METHOD_WATER_AROUND) {
            make_grid_pbc(f.g1, pg.box, sR, f.lo, f.hi, rp.cell1, rp.cell1, bv.cell_cap);
            make_grid_pbc(f.g2, pg.box, sR, f.lo, f.hi, rp.around_f, rp.cell2, bv.cell_cap);
        } else {
            make_grid_pbc(f.g2, pg.box, sR, f.lo, f.hi, 0.f, rp.cell2, bv.cell_cap);
            f.g1 = f.g2;
        }
    } else {
        if (rp.method == HB_METHOD_WATER_AROUND) {
            make_grid(f.g1, f.lo, f.hi, rp.cell1, rp.cell1, bv.cell_cap);
            make_grid(f.g2, f.lo, f.hi, rp.around_f, rp.cell2, bv.cell_cap);
        } else {
            make_grid(f.g2, f.lo, f.hi, 0.f, rp.cell2, bv.cell_cap);
            f.g1 = f.g2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3-K5: counting sort of a point list into the cells of a grid. which = 1: selection points -> g1;
// which = 2: selection points ++ local waters -> g2.  Grids are (X, frames): X blocks stride over the
// points of their frame, so frames with few points cost few idle blocks.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int n_points_of(const BatchView& bv, const DevTopo& tp, int b, int which) {
    int sys = bv.structure_batch ? (int)(bv.frame0 + b) : 0;
    int nsel = tp.sys_sel_off[sys + 1] - tp.sys_sel_off[sys];
    return which == 1 ? nsel : nsel + bv.fs[b].n_local;
}
__device__ __forceinline__ float4 load_point(const BatchView& bv, const DevTopo& tp, int b, int i) {
    int sys = bv.structure_batch ? (int)(bv.frame0 + b) : 0;
    int nsel = tp.sys_sel_off[sys + 1] - tp.sys_sel_off[sys];
    if (i < nsel) return bv.selpos[(size_t)b * bv.max_sel + i];
    return bv.lw[(size_t)b * bv.max_wat + (i - nsel)];
}

__global__ void __launch_bounds__(256) k_zero_cells(BatchView bv, int which) {
    int b = blockIdx.y;
    const FrameGrid& g = which == 1 ? bv.fs[b].g1 : bv.fs[b].g2;
    int* cells = (which == 1 ? bv.cells1 : bv.cells2) + (size_t)b * (bv.cell_cap + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= g.ncell; i += gridDim.x * blockDim.x) cells[i] = 0;
}

template <bool PBC>
__global__ void __launch_bounds__(256) k_count(BatchView bv, DevTopo tp, int which) {
    int b = blockIdx.y;
    int n = n_points_of(bv, tp, b, which);
    const FrameGrid& g = which == 1 ? bv.fs[b].g1 : bv.fs[b].g2;
    int* cells = (which == 1 ? bv.cells1 : bv.cells2) + (size_t)b * (bv.cell_cap + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = load_point(bv, tp, b, i);
        atomicAdd(&cells[cell_of<PBC>(g, p.x, p.y, p.z)], 1);
    }
}

// exclusive scan of the per-cell counts of one frame (one block per frame), in place. Every warp owns one contiguous
// segment: it adds it up (pass 1), the 32 warp totals are combined once, then it scans its segment 32 cells at a time
// with a running carry (pass 2) - two block barriers in all instead of four per 1024 cells.
__global__ void __launch_bounds__(1024) k_scan(BatchView bv, int which) {
    const int b = blockIdx.x;
    const FrameGrid& g = which == 1 ? bv.fs[b].g1 : bv.fs[b].g2;
    int* cells = (which == 1 ? bv.cells1 : bv.cells2) + (size_t)b * (bv.cell_cap + 1);
    __shared__ int warp_tot[32];
    const int n = g.ncell;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int per = (((n + 31) / 32 + 31) / 32) * 32;           // cells per warp, a multiple of 32
    const int s = min(w * per, n), e = min(s + per, n);
    int sum = 0;
    for (int i = s + lane; i < e; i += 32) sum += cells[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) warp_tot[w] = sum;
    __syncthreads();
    int carry = 0;
    for (int k = 0; k < w; ++k) carry += warp_tot[k];
    int v_next = (s + lane < e) ? cells[s + lane] : 0;
    for (int i0 = s; i0 < e; i0 += 32) {
        const int i = i0 + lane;
        const int v = v_next;
        if (i0 + 32 < e) v_next = (i + 32 < e) ? cells[i + 32] : 0;      // next strip in flight before this one is stored
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) sc += t;
        }
        if (i < e) cells[i] = carry + sc - v;
        carry += __shfl_sync(0xffffffffu, sc, 31);
    }
    if (threadIdx.x == 0) {
        int total = 0;
        for (int k = 0; k < 32; ++k) total += warp_tot[k];
        cells[n] = total;
        if (which == 2) bv.fs[b].n_points = total;
    }
}

// after this kernel cells[c] holds the END of cell c; its start is cells[c-1] (0 for c = 0)
template <bool PBC>
__global__ void __launch_bounds__(256) k_scatter(BatchView bv, DevTopo tp, int which) {
    int b = blockIdx.y;
    int n = n_points_of(bv, tp, b, which);
    const FrameGrid& g = which == 1 ? bv.fs[b].g1 : bv.fs[b].g2;
    int* cells = (which == 1 ? bv.cells1 : bv.cells2) + (size_t)b * (bv.cell_cap + 1);
    float4* dst = which == 1 ? bv.sorted1 + (size_t)b * bv.max_sel : bv.sorted2 + (size_t)b * (bv.max_sel + bv.max_wat);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = load_point(bv, tp, b, i);
        dst[atomicAdd(&cells[cell_of<PBC>(g, p.x, p.y, p.z)], 1)] = p;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: local waters = waters with DIST <= around_radius to any selection entry (hbond.py:249-252)
// ------------------------------------------------------------------------------------------------
template <bool PBC>
__global__ void __launch_bounds__(256) k_local_water(BatchView bv, DevTopo tp, RunParams rp) {
    int b = blockIdx.y;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    int w0 = tp.sys_wat_off[sys], w1 = tp.sys_wat_off[sys + 1];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const FrameState& f = bv.fs[b];
    bool local = false;
    float x = 0.f, y = 0.f, z = 0.f;
    if (w0 + i < w1) {
        const float* p = xyz + 3ll * tp.wat_atom[w0 + i];
        x = p[0]; y = p[1]; z = p[2];
        bool inside;
        if constexpr (PBC) {
            int cx, cy, cz;
            inside = cell3_pbc(f.g1, x, y, z, cx, cy, cz);
        } else {
            float a = rp.around_f;
            inside = x >= f.lo[0] - a && x <= f.hi[0] + a && y >= f.lo[1] - a && y <= f.hi[1] + a &&
                     z >= f.lo[2] - a && z <= f.hi[2] + a;
        }
        if (inside)
            local = water_is_local<PBC, int, true>(f.g1, bv.cells1 + (size_t)b * (bv.cell_cap + 1),
                                        bv.sorted1 + (size_t)b * bv.max_sel, x, y, z, rp);
    }
    unsigned m = __ballot_sync(0xffffffffu, local);
    if (m) {
        int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == 0) base = atomicAdd(&bv.fs[b].n_local, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (local) {
            int pos = base + __popc(m & ((1u << lane) - 1u));
            bv.lw[(size_t)b * bv.max_wat + pos] = make_float4(x, y, z, __int_as_float(WATER_TAG | i));
        }
    }
}

// K6 with the selection grid in shared memory: when the cell-sorted selection points and the g1 cell offsets of a
// frame fit (the host checks the points, the kernel the cells), a block stages them once and tests `per_block` waters
// against them - the dependent cell-offset and point loads then cost shared-memory instead of L2 latency.
#define LW_SMEM_CELLS 4096
template <bool PBC>
__global__ void __launch_bounds__(256) k_local_water_smem(BatchView bv, DevTopo tp, RunParams rp, int per_block) {
    extern __shared__ float4 s_pts[];                    // [max_sel] points, then [LW_SMEM_CELLS + 1] cell offsets
    int* s_cells = (int*)(s_pts + bv.max_sel);
    const int b = blockIdx.y;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const int w0 = tp.sys_wat_off[sys], w1 = tp.sys_wat_off[sys + 1];
    const int n_sel = tp.sys_sel_off[sys + 1] - tp.sys_sel_off[sys];
    const FrameState& f = bv.fs[b];
    const int first = blockIdx.x * per_block;
    if (w0 + first >= w1) return;
    const int* g_cells = bv.cells1 + (size_t)b * (bv.cell_cap + 1);
    const float4* g_pts = bv.sorted1 + (size_t)b * bv.max_sel;
    const bool staged = f.g1.ncell <= LW_SMEM_CELLS;     // uniform per block
    if (staged) {
        for (int i = threadIdx.x; i < n_sel; i += blockDim.x) s_pts[i] = g_pts[i];
        for (int i = threadIdx.x; i <= f.g1.ncell; i += blockDim.x) s_cells[i] = g_cells[i];
    }
    __syncthreads();
    const int* cells = staged ? s_cells : g_cells;
    const float4* pts = staged ? s_pts : g_pts;
    for (int base = first; base < first + per_block; base += blockDim.x) {
        const int i = base + threadIdx.x;
        bool local = false;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < first + per_block && w0 + i < w1) {
            const float* p = xyz + 3ll * tp.wat_atom[w0 + i];
            x = p[0]; y = p[1]; z = p[2];
            bool inside;
            if constexpr (PBC) {
                int cx, cy, cz;
                inside = cell3_pbc(f.g1, x, y, z, cx, cy, cz);
            } else {
                const float a = rp.around_f;
                inside = x >= f.lo[0] - a && x <= f.hi[0] + a && y >= f.lo[1] - a && y <= f.hi[1] + a &&
                         z >= f.lo[2] - a && z <= f.hi[2] + a;
            }
            if (inside) local = water_is_local<PBC, int, true>(f.g1, cells, pts, x, y, z, rp);
        }
        const unsigned m = __ballot_sync(0xffffffffu, local);
        if (m) {
            const int lane = threadIdx.x & 31;
            int pos0 = 0;
            if (lane == 0) pos0 = atomicAdd(&bv.fs[b].n_local, __popc(m));
            pos0 = __shfl_sync(0xffffffffu, pos0, 0);
            if (local)
                bv.lw[(size_t)b * bv.max_wat + pos0 + __popc(m & ((1u << lane) - 1u))] = make_float4(x, y, z, __int_as_float(WATER_TAG | i));
        }
    }
}

// HB_METHOD_WATER_PRESENCE (HydrationAnalysis.set_presence_around, hydration.py:71-107): every local water of a
// frame sets its bit in the row of its water id.
__global__ void __launch_bounds__(256) k_presence_emit(BatchView bv, DevTopo tp, BondTable bt) {
    const int b = blockIdx.y;
    const int sys = bv.structure_batch ? (int)(bv.frame0 + b) : 0;
    const int n = bv.fs[b].n_local;
    const int w0 = tp.sys_wat_off[sys];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int w = __float_as_int(bv.lw[(size_t)b * bv.max_wat + i].w) & POINT_ID_MASK;
        record_bond(bt, (unsigned long long)(unsigned)tp.ent_group[tp.wat_ent[w0 + w]], bv.frame0 + b);
    }
}

template <bool PBC>
__global__ void __launch_bounds__(128) k_pairs(BatchView bv, DevTopo tp, RunParams rp, BondTable bt) {
    int b = blockIdx.y;
    const FrameState& f = bv.fs[b];
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const int* cells = bv.cells2 + (size_t)b * (bv.cell_cap + 1);
    const float4* pts = bv.sorted2 + (size_t)b * (bv.max_sel + bv.max_wat);
    long long frame = bv.frame0 + b;
    int nhit = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < f.n_points; i += gridDim.x * blockDim.x)
        search_point<PBC>(tp, rp, bt, xyz, frame, sys, f.g2, cells, pts, i, nhit);
    flush_hits(bt, nhit);
}

// ------------------------------------------------------------------------------------------------
// Ion contacts (HbondAnalysis.set_add_ion_interactions, reference hbond.py:57-94): every selection entry and,
// with include_water, every water of the water group within `distance` of an ion; key = id of the atom ':' id of
// the ion, no angle criterion, no orientation rule. Ions are few: their coordinates sit in shared memory and
// every point (one thread each) tests all of them with the exact DIST predicate.
// ------------------------------------------------------------------------------------------------
struct IonTable {
    const int* atom;      // atom index inside the frame
    const int* group;     // key group id of the ion's id string
    int n;
};

template <bool PBC>
__global__ void __launch_bounds__(256) k_ion_contacts(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, IonTable ions, int include_water) {
    __shared__ float4 ion_s[1024];
    const int b = blockIdx.y;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const long long frame = bv.frame0 + b;
    const int s0 = tp.sys_sel_off[sys], nsel = tp.sys_sel_off[sys + 1] - s0;
    const int w0 = tp.sys_wat_off[sys], nwp = include_water ? tp.sys_wat_off[sys + 1] - w0 : 0;
    PbcGrid pg;
    if constexpr (PBC) {
        pg.box = load_box(bv.box + 9ll * b);
        pg.ia = 1.0f / pg.box.ax; pg.ib = 1.0f / pg.box.by; pg.ic = 1.0f / pg.box.cz;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < nsel + nwp;
    float x = 0.f, y = 0.f, z = 0.f;
    int g0 = -1, g1 = -1;                       // key groups of the point (a selection atom of the water group has two)
    if (live) {
        int atom;
        if (i < nsel) {
            const int4 k = tp.sel_pack[s0 + i];
            atom = tp.sel_atom[s0 + i];
            g0 = tp.ent_group[k.x >= 0 ? k.x : k.y];
            if (include_water && k.z >= 0) g1 = tp.ent_group[k.z];
        } else {
            atom = tp.wat_atom[w0 + i - nsel];
            g0 = tp.ent_group[tp.wat_ent[w0 + i - nsel]];
        }
        const float* p = xyz + 3ll * atom;
        x = p[0]; y = p[1]; z = p[2];
    }
    int nhit = 0;
    for (int base = 0; base < ions.n; base += 1024) {
        const int m = min(1024, ions.n - base);
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += blockDim.x) {
            const float* q = xyz + 3ll * ions.atom[base + k];
            ion_s[k] = make_float4(q[0], q[1], q[2], __int_as_float(ions.group[base + k]));
        }
        __syncthreads();
        if (!live) continue;
        for (int k = 0; k < m; ++k) {
            const float4 q = ion_s[k];
            if (!within<PBC>(x, y, z, q.x, q.y, q.z, rp.r2f_hi, rp.r2, pg)) continue;
            const unsigned long long gi = (unsigned long long)(unsigned)__float_as_int(q.w);
            record_bond(bt, ((unsigned long long)(unsigned)g0 << 32) | gi, frame);
            ++nhit;
            if (g1 >= 0 && g1 != g0) { record_bond(bt, ((unsigned long long)(unsigned)g1 << 32) | gi, frame); ++nhit; }
        }
    }
    flush_hits(bt, nhit);
}

// ------------------------------------------------------------------------------------------------
// K7 in two dense steps (used when the batch-wide pair buffer is available):
//   K7a  every point walks its neighbourhood and appends the point pairs within `distance` to one pair buffer
//        for the whole sub-batch (records staged in shared memory, one global atomic per block and trip)
//   K7b  one thread per pair: topology of both points, angle tests, entry pairs, bond table
// A thread then carries one short chain of dependent global loads instead of one per partner of its point.
// Pair record: frame slot of the sub-batch << 48 | i << 24 | q  (cell-sorted point indices, < 2^24).
// ------------------------------------------------------------------------------------------------
#define PAIR_STAGE 2048

struct PairBuf {
    unsigned long long* rec;
    unsigned long long cap;
    unsigned long long* count;     // records appended (may exceed cap: the excess is dropped and flagged)
    int* overflow;
};

template <bool PBC>
__global__ void __launch_bounds__(128) k_collect_pairs(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, PairBuf pb) {
    __shared__ unsigned long long stage[PAIR_STAGE];
    __shared__ unsigned n_stage;
    __shared__ unsigned long long g_base;
    const int b = blockIdx.y;
    const FrameState& f = bv.fs[b];
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const int* cells = bv.cells2 + (size_t)b * (bv.cell_cap + 1);
    const float4* pts = bv.sorted2 + (size_t)b * (bv.max_sel + bv.max_wat);
    const long long frame = bv.frame0 + b;
    const int n = f.n_points;
    int nhit = 0;
    if (threadIdx.x == 0) n_stage = 0;
    __syncthreads();
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        if (i < n) {
            const float4 p = pts[i];
            if (rp.self_pairs) {
                PointInfo P;
                load_info(tp, rp, sys, p, P);
                process_self<PBC>(tp, rp, bt, xyz, frame, P, nhit, f.g2.p.box);
            }
            for_each_partner<PBC>(f.g2, cells, i, p.x, p.y, p.z, [&](int q) {
                const float4 q4 = pts[q];
                if (!within<PBC>(p.x, p.y, p.z, q4.x, q4.y, q4.z, rp.r2f_hi, rp.r2, f.g2.p)) return;
                const unsigned long long r = ((unsigned long long)b << 48) | ((unsigned long long)i << 24) | (unsigned long long)q;
                const unsigned k = atomicAdd(&n_stage, 1u);
                if (k < PAIR_STAGE) stage[k] = r;
                else {                                          // staging full: straight to the buffer
                    const unsigned long long g = atomicAdd(pb.count, 1ull);
                    if (g < pb.cap) pb.rec[g] = r; else *pb.overflow = 1;
                }
            });
        }
        __syncthreads();
        const unsigned m = min(n_stage, (unsigned)PAIR_STAGE);
        if (threadIdx.x == 0 && m) g_base = atomicAdd(pb.count, (unsigned long long)m);
        __syncthreads();
        for (unsigned k = threadIdx.x; k < m; k += blockDim.x) {
            const unsigned long long g = g_base + k;
            if (g < pb.cap) pb.rec[g] = stage[k]; else *pb.overflow = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) n_stage = 0;
        __syncthreads();
    }
    flush_hits(bt, nhit);
}

template <bool PBC>
__global__ void __launch_bounds__(128) k_expand_pairs(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, PairBuf pb) {
    const unsigned long long n = min(*pb.count, pb.cap);
    int nhit = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = pb.rec[k];
        const int b = (int)(r >> 48), i = (int)((r >> 24) & 0xffffffull), q = (int)(r & 0xffffffull);
        int sys;
        const float* xyz = frame_xyz(bv, tp, b, sys);
        const float4* pts = bv.sorted2 + (size_t)b * (bv.max_sel + bv.max_wat);
        PointInfo P, Q;
        load_info(tp, rp, sys, pts[i], P);
        load_info(tp, rp, sys, pts[q], Q);
        process_pair<PBC>(tp, rp, bt, xyz, bv.frame0 + b, P, Q, nhit, bv.fs[b].g2.p.box);
    }
    flush_hits(bt, nhit);
}

}  // namespace

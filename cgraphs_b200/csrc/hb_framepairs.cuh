// hb_framepairs.cuh - the pair search of the general path as one CTA per frame.
//
// When a frame's pair-search set (selection points + local waters) fits in shared memory but the
// frame did not fit the fully fused kernel (wide water shells: hb_fused.cuh also keeps the selection grid, the water
// tile ring and the local-water list resident), the general path still finds the local waters with its global-memory
// kernels (K1-K6, hb_general.cuh) and then runs THIS kernel instead of the g2 counting sort + k_collect_pairs +
// k_expand_pairs: the points are cell-sorted in shared memory and the pairs are expanded with the fused kernel's
// dense three-step scheme (hb_fused.cuh phase C: topology + rules, one task per hydrogen, emit), a tile of 1024
// points at a time so that the pair list stays small. One 1024-thread CTA per SM.
//
// Capacity misses (points, cells, pairs of a tile, hydrogen tasks of a chunk) set a status word; the host then
// switches the run back to the multi-kernel pair search and re-runs the batch (idempotent updates).
#pragma once
#include "hb_fused.cuh"

namespace {

#define FP_THREADS 1024      // one CTA per SM: the frame's points, cells and work lists take most of its shared memory

struct FpShared {            // small fixed part at the start of dynamic shared memory
    FrameGrid g2;
    int n_pairs;
    int n_tasks;
    int warp_sum[FP_THREADS / 32];
};

__device__ __forceinline__ void fp_sync() { __syncthreads(); }

// exclusive scan of n 16-bit counters in place (whole CTA)
__device__ void fp_scan_cells(unsigned short* cells, int n, FpShared* sh) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n + FP_THREADS - 1) / FP_THREADS;
    const int s = min(tid * per, n), e = min(s + per, n);
    int sum = 0;
    for (int i = s; i < e; ++i) sum += cells[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh->warp_sum[w] = incl;
    fp_sync();
    int base = 0;
    for (int k = 0; k < w; ++k) base += sh->warp_sum[k];
    int run = base + incl - sum;
    for (int i = s; i < e; ++i) {
        int v = cells[i];
        cells[i] = (unsigned short)run;
        run += v;
    }
    fp_sync();
}

// sort `n` points (src(i)) into the cells of g: zero, count, scan, scatter. After return cells[c] is the END of cell c.
template <bool PBC, typename Src>
__device__ void fp_cell_sort(const FrameGrid& g, unsigned short* cells, float4* dst, int n, Src src, FpShared* sh) {
    unsigned* words = reinterpret_cast<unsigned*>(cells);
    const int nwords = (g.ncell + 2) >> 1;
    for (int i = threadIdx.x; i < nwords; i += FP_THREADS) words[i] = 0u;
    fp_sync();
    for (int i = threadIdx.x; i < n; i += FP_THREADS) {
        float4 p = src(i);
        cell_add(words, cell_of<PBC>(g, p.x, p.y, p.z));
    }
    fp_sync();
    fp_scan_cells(cells, g.ncell, sh);
    for (int i = threadIdx.x; i < n; i += FP_THREADS) {
        float4 p = src(i);
        dst[cell_add(words, cell_of<PBC>(g, p.x, p.y, p.z))] = p;
    }
    fp_sync();
}

struct FramePairsCfg {
    int np_cap;               // points per frame
    int cell_cap;             // g2 cells per frame (16-bit counters)
    int pair_cap;             // point pairs of one tile of FP_THREADS points
    int task_cap;             // hydrogens to test per chunk of FP_THREADS pairs
    unsigned off_pts, off_pairs, off_tasks, off_flags, off_cells;   // byte offsets in dynamic shared memory
    unsigned smem_bytes;
    int* overflow;            // status word: some frame did not fit
};

template <bool PBC>
__global__ void __launch_bounds__(FP_THREADS) k_frame_pairs(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, FramePairsCfg fc) {
    extern __shared__ __align__(16) unsigned char smem[];
    FpShared* sh = reinterpret_cast<FpShared*>(smem);
    float4* psorted = reinterpret_cast<float4*>(smem + fc.off_pts);
    unsigned* pairs = reinterpret_cast<unsigned*>(smem + fc.off_pairs);
    uint2* tasks = reinterpret_cast<uint2*>(smem + fc.off_tasks);
    unsigned* pflags = reinterpret_cast<unsigned*>(smem + fc.off_flags);
    unsigned short* cells = reinterpret_cast<unsigned short*>(smem + fc.off_cells);

    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const long long frame = bv.frame0 + b;
    const FrameState& f = bv.fs[b];
    const int nsel = tp.sys_sel_off[sys + 1] - tp.sys_sel_off[sys];
    const int nlw = rp.method == HB_METHOD_WATER_AROUND ? f.n_local : 0;
    const int np = nsel + nlw;
    if (np > fc.np_cap || (PBC && f.g2.ncell > fc.cell_cap)) {  // uniform
        if (tid == 0) atomicExch(fc.overflow, 1);
        return;
    }
    if (tid == 0) {
        sh->g2 = f.g2;
        if constexpr (!PBC) {      // more cells than the counters here hold: the same box with larger cells
            if (f.g2.ncell > fc.cell_cap)
                make_grid(sh->g2, f.lo, f.hi, rp.method == HB_METHOD_WATER_AROUND ? rp.around_f : 0.f, rp.cell2, fc.cell_cap);
        }
    }
    fp_sync();
    const float4* selpos = bv.selpos + (size_t)b * bv.max_sel;
    const float4* lw = bv.lw + (size_t)b * bv.max_wat;
    fp_cell_sort<PBC>(sh->g2, cells, psorted, np, [&](int i) { return i < nsel ? selpos[i] : lw[i - nsel]; }, sh);
    const FrameGrid g2 = sh->g2;
    const bool ang = rp.check_angle != 0;
    const float ct = rp.cos_threshold;
    int nhit = 0;

    for (int tile = 0; tile < np; tile += FP_THREADS) {
        if (tid == 0) sh->n_pairs = 0;
        fp_sync();
        // C1: the points of this tile walk their half shells and append the point pairs within `distance`
        const int i = tile + tid;
        if (i < np) {
            collect_pairs<PBC>(rp, g2, cells, psorted, i, pairs, fc.pair_cap, &sh->n_pairs);
            if (rp.self_pairs) {
                PointInfo P;
                load_info(tp, rp, sys, psorted[i], P);
                process_self<PBC>(tp, rp, bt, xyz, frame, P, nhit, g2.p.box);
            }
        }
        fp_sync();
        const int npairs = sh->n_pairs;
        if (npairs > fc.pair_cap) {                             // uniform
            if (tid == 0) atomicExch(fc.overflow, 1);
            return;
        }
        // C2: pairs in chunks of 256, three dense steps per chunk (hb_fused.cuh, phase C)
        for (int base = 0; base < npairs; base += FP_THREADS) {
            const int k = base + tid;
            const bool have = k < npairs;
            PointInfo P, Q;
            unsigned rules = 0;
            if (tid == 0) sh->n_tasks = 0;
            pflags[tid] = 0u;
            fp_sync();
            // (a) topology of the two points, entry pairs that exist, one task per hydrogen to test
            if (have) {
                const unsigned pq = pairs[k];
                load_info(tp, rp, sys, psorted[pq & 0xffffu], P);
                load_info(tp, rp, sys, psorted[pq >> 16], Q);
                rules = pair_rules(P, Q, rp.excl_bb != 0);
                if (ang) {
                    const unsigned need = rules_need(rules);
                    if (tp.h_wide) {                            // hydrogen lists too long for the inline tables: CSR walk
                        unsigned fl = 0;
                        if ((need & 1u) && any_h<PBC>(tp, xyz, P.don, P, Q, ct, g2.p.box)) fl |= 1u;
                        if ((need & 2u) && any_h<PBC>(tp, xyz, Q.don, Q, P, ct, g2.p.box)) fl |= 2u;
                        if ((need & 4u) && any_h<PBC>(tp, xyz, P.wat, P, Q, ct, g2.p.box)) fl |= 4u;
                        if ((need & 8u) && any_h<PBC>(tp, xyz, Q.wat, Q, P, ct, g2.p.box)) fl |= 8u;
                        pflags[tid] = fl;
                    } else {
                        const int lists[4] = {P.hd, Q.hd, P.hw, Q.hw};
                        const int atoms[4] = {P.atom, Q.atom, P.atom, Q.atom};
                        int nt = 0;
#pragma unroll
                        for (int fl = 0; fl < 4; ++fl)
                            if (need & (1u << fl)) nt += h_count(lists[fl]);
                        int slot = nt ? atomicAdd(&sh->n_tasks, nt) : 0;
                        if (slot + nt <= fc.task_cap) {
#pragma unroll
                            for (int fl = 0; fl < 4; ++fl)
                                if (need & (1u << fl)) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const int d = (int)(signed char)((unsigned)lists[fl] >> (8 * j));
                                        if (d != -128) tasks[slot++] = make_uint2((unsigned)tid | ((unsigned)fl << 16), (unsigned)(atoms[fl] + d));
                                    }
                                }
                        }
                    }
                } else {
                    pflags[tid] = 0xfu;
                }
            }
            fp_sync();
            // (b) one thread per hydrogen: ANG(other, h, owner)
            const int ntasks = sh->n_tasks;
            if (ntasks > fc.task_cap) {                         // uniform
                if (tid == 0) atomicExch(fc.overflow, 1);
                return;
            }
            for (int t = tid; t < ntasks; t += FP_THREADS) {
                const uint2 tk = tasks[t];
                const unsigned kk = tk.x & 0xffffu, fl = tk.x >> 16;
                const unsigned pq = pairs[base + kk];
                const float4 a4 = psorted[pq & 0xffffu], b4 = psorted[pq >> 16];
                const float4 own = (fl & 1u) ? b4 : a4, oth = (fl & 1u) ? a4 : b4;
                const float* h = xyz + 3ll * (long long)tk.y;
                if (ang_ok<PBC>(oth.x, oth.y, oth.z, h[0], h[1], h[2], own.x, own.y, own.z, ct, g2.p.box)) atomicOr(&pflags[kk], 1u << fl);
            }
            fp_sync();
            // (c) entry pairs whose hydrogens passed
            if (have) {
                unsigned m = rules & rules_pass(pflags[tid]);
                while (m) {
                    const int r = __ffs(m) - 1;
                    m &= m - 1;
                    int e0, e1;
                    rule_entries(P, Q, r, e0, e1);
                    emit(tp, rp, bt, frame, e0, e1, nhit);
                }
            }
            fp_sync();                                    // pflags / tasks are rewritten by the next chunk
        }
    }
    flush_hits(bt, nhit);
}

}  // namespace

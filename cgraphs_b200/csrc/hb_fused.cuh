// hb_fused.cuh - fused fast path: ONE CTA processes ONE frame end to end; everything but the frame's
// coordinates and the bond table stays in shared memory.
//
// The CTA has 8 warps.
//
//   phase A  gather the selection points (<= sel_cap) into smem, bounding box, cell grid g1 (edge >= around_radius),
//            counting sort, "near" bitmap: bit c is set iff some selection point lies in the 27-cell neighbourhood of cell c.
//   phase B  the frame's water block streams through a ring of shared-memory stages with 1-D bulk async copies (TMA engine:
//            cp.async.bulk + mbarrier complete_tx); a tile is 32 x FUSED_WPL consecutive waters (O,H,H,O,H,H ... as they lie
//            in the coordinate array), so HBM is read in large aligned bursts exactly once. Warp w < stages owns stage w and
//            every stages-th tile: FUSED_WPL waters per lane, dilated-bounding-box test -> near-bitmap test -> the survivors'
//            ("candidates") indices go to a CTA-wide list; the owner refills its stage as soon as it has read it.
//   phase B2 candidates are tested densely against the 27 neighbouring g1 cells with the exact DIST predicate; local
//            waters (hbond.py:249-252) are appended to a shared list.
//   phase C  cell grid g2 (edge >= distance) over selection + local waters; half-shell walk collecting the point pairs
//            that pass the exact DIST test into a pair list (C1), then per chunk of 256 pairs (C2): one thread per pair
//            expands it into the reference's entry pairs and writes one task per hydrogen, one thread per task applies the
//            exact ANG test, one thread per pair records the bonds (hash-table insert + frame bit). Loads that depend on
//            each other through L2 / DRAM are requested several at a time (three tasks, two bonds, two points per trip):
//            the number of dependent round trips per frame, times the loaded memory latency, is what a frame's life is made of.
//
// Four CTAs are resident per SM, so the gathers of phases A / C of one frame overlap the streaming of others; the first
// wave of a launch starts staggered (FusedCfg::stagger), otherwise frames that all cost the same stream and compute in
// lock step.  Frames that exceed the shared capacities (sel_cap / lw_cap / pair_cap) raise `fused_overflow`; the host
// then re-runs the batch with doubled capacities or on the general path (hb_general.cuh).
#pragma once
#include "hb_common.cuh"

namespace {

#define FUSED_CWARPS 8
#define FUSED_CONSUMERS (FUSED_CWARPS * 32)
#define FUSED_THREADS FUSED_CONSUMERS
#ifndef FUSED_WPL
#define FUSED_WPL 3             // waters per lane and tile
#endif
#define FUSED_TILE (32 * FUSED_WPL)   // waters per tile: one warp takes a whole tile
#define FUSED_MAX_STAGES 8

struct FusedCfg {
    int sel_cap;              // selection points per frame that fit
    int lw_cap;               // local waters per frame that fit
    int cell_cap;             // cells per grid (packed 16-bit counters)
    int pair_cap;             // point pairs per frame
    int cand_cap;             // candidate waters per frame (indices, collected while the tiles stream)
    int task_cap;             // hydrogens to test per chunk of 256 pairs
    int regular;              // 1: water point i is atom w_first + i * w_stride in every system -> TMA tiles
    int w_stride;             // atoms between consecutive water oxygens
    int stages;               // stages of the ring = warps that stream (one stage each)
    int prefetch;             // tiles pulled into L2 ahead of the ring
    unsigned wait_ns;         // sleep between two polls of a barrier
    unsigned stage_bytes;     // bytes of one stage (multiple of 16)
    int cand16;               // candidate indices are 16-bit (every system has < 65536 waters)
    unsigned off_sorted, off_ring, off_lw, off_cells, off_near, off_cand, off_bar;   // byte offsets in dynamic smem
    unsigned smem_bytes;
    unsigned long long buf_begin, buf_end;   // address range of the coordinate batch (bounds for bulk copies)
    int debug;                // measurement only: bit 0 = no water becomes a candidate, bit 1 = skip phase C
    unsigned stagger;         // cycles over which the CTAs of the first wave spread their start (0 = all start together)
    unsigned wave_ctas;       // CTAs of the first wave (SMs x resident CTAs per SM)
    unsigned n_sm;
    unsigned pf_off, pf_bytes; // byte range of a frame that holds the selection points and their hydrogens (0 = none)
    int* sm_streams;          // [n_sm] frames streaming on each SM right now (admission: at most max_streams per SM)
    int max_streams;          // 0 = no limit
    unsigned long long* cta_trace;   // diagnostic (HBGPU_CTA_TRACE): [frame][8] = globaltimer at start / after the stagger / B begin / B end / B2 end / end, SM id
};

struct FusedShared {          // small fixed part at the start of dynamic shared memory
    FrameGrid g1, g2;
    float lo[3], hi[3];
    float red[FUSED_CWARPS][6];
    int n_lw;
    int n_cand;
    int n_pairs;
    int n_tasks;
    int overflow;
    int warp_sum[FUSED_CWARPS];
    unsigned stage_lo[FUSED_MAX_STAGES];     // [lo, hi): bytes of the stage the bulk copy filled
    unsigned stage_hi[FUSED_MAX_STAGES];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Waits take 32-bit shared-window addresses computed once outside the loop (the generic-pointer forms make the compiler
// rebuild the shared-window base - S2R SR_CgaCtaId, LEA - at every use). try_wait carries a suspend-time hint: the warp
// sleeps in hardware until the phase completes or ~1 us passed. The slow path is bounded (about a second): a lost copy
// must never hang the GPU, the caller flags the frame instead; between its polls the warp sleeps `ns` nanoseconds.
__device__ __forceinline__ bool mbar_try_wait_a(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(1000u)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ bool mbar_wait_slow_a(unsigned bar, unsigned parity, unsigned ns) {
    const long long t0 = clock64();
    for (;;) {
#pragma unroll 1
        for (int k = 0; k < 64; ++k) {
            if (ns) __nanosleep(ns);
            if (mbar_try_wait_a(bar, parity)) return true;
        }
        if (clock64() - t0 > 2000000000ll) return false;
    }
}
__device__ __forceinline__ void mbar_arrive_a(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(FUSED_CONSUMERS) : "memory"); }

// counting-sort helpers on packed 16-bit cell counters ---------------------------------------------
__device__ __forceinline__ unsigned cell_add(unsigned* words, int c) {   // returns the old 16-bit value
    unsigned old = atomicAdd(&words[c >> 1], (c & 1) ? 0x10000u : 1u);
    return (c & 1) ? (old >> 16) : (old & 0xffffu);
}

// exclusive scan of n 16-bit counters in place (all consumers)
__device__ void block_scan_cells(unsigned short* cells, int n, FusedShared* sh) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n + FUSED_CONSUMERS - 1) / FUSED_CONSUMERS;
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
    consumer_sync();
    int base = 0;
    for (int k = 0; k < w; ++k) base += sh->warp_sum[k];
    int run = base + incl - sum;
    for (int i = s; i < e; ++i) {
        int v = cells[i];
        cells[i] = (unsigned short)run;
        run += v;
    }
    consumer_sync();
}

// sort `n` points (src(i)) into the cells of g: zero, count, scan, scatter. After return cells[c] is the END of cell c.
template <bool PBC, typename Src>
__device__ void block_cell_sort(const FrameGrid& g, unsigned short* cells, float4* dst, int n, Src src, FusedShared* sh) {
    unsigned* words = reinterpret_cast<unsigned*>(cells);
    const int nwords = (g.ncell + 2) >> 1;
    for (int i = threadIdx.x; i < nwords; i += FUSED_CONSUMERS) words[i] = 0u;
    consumer_sync();
    for (int i = threadIdx.x; i < n; i += FUSED_CONSUMERS) {
        float4 p = src(i);
        cell_add(words, cell_of<PBC>(g, p.x, p.y, p.z));
    }
    consumer_sync();
    block_scan_cells(cells, g.ncell, sh);
    for (int i = threadIdx.x; i < n; i += FUSED_CONSUMERS) {
        float4 p = src(i);
        const int c = cell_of<PBC>(g, p.x, p.y, p.z);
        HB_DASSERT(c >= 0 && c < g.ncell);
        const unsigned at = cell_add(words, c);
        HB_DASSERT((int)at < n);
        dst[at] = p;
    }
    consumer_sync();
}

// point pairs (i, q), i < q in cell order, that pass the exact DIST test
template <bool PBC>
__device__ __forceinline__ void collect_pairs(const RunParams& rp, const FrameGrid& g, const unsigned short* cells,
                                              const float4* pts, int i, unsigned* pairs, int pair_cap, int* n_pairs) {
    const float4 p = pts[i];
    for_each_partner<PBC>(g, cells, i, p.x, p.y, p.z, [&](int q) {
        const float4 q4 = pts[q];
        if (!within<PBC>(p.x, p.y, p.z, q4.x, q4.y, q4.z, rp.r2f_hi, rp.r2, g.p)) return;
        const int k = atomicAdd(n_pairs, 1);
        HB_DASSERT(i >= 0 && i < 65536 && q > i && q < 65536);
        if (k < pair_cap) pairs[k] = (unsigned)i | ((unsigned)q << 16);
    });
}

// The two neighbour walks of the non-periodic fused path on 32-bit shared-window addresses (`cells_a`: 16-bit end offsets
// per cell, `pts_a`: float4 points in cell order): the same walks as water_is_local / for_each_partner in hb_common.cuh,
// but the generic-pointer forms make the compiler rebuild the shared-window base at every access of these short loops.
__device__ __forceinline__ bool water_is_local_a(const FrameGrid& g, unsigned cells_a, unsigned pts_a, float x, float y, float z,
                                                 const RunParams& rp) {
    int cx, cy, cz;
    cell3<false>(g, x, y, z, cx, cy, cz);
    const int xl = max(cx - 1, 0), xh = min(cx + 1, g.nx - 1);
    for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.nz - 1); ++zz)
        for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); ++yy) {
            const int row = (zz * g.ny + yy) * g.nx;
            const int c_lo = row + xl;
            const int s = c_lo ? (int)lds_u16(cells_a + 2u * (unsigned)(c_lo - 1)) : 0;
            const int e = (int)lds_u16(cells_a + 2u * (unsigned)(row + xh));
            for (int q = s; q < e; ++q) {
                const float4 sp = lds_f4(pts_a + 16u * (unsigned)q);
                if (dist2f(x, y, z, sp.x, sp.y, sp.z) <= rp.around2f_hi && dist_le(x, y, z, sp.x, sp.y, sp.z, rp.around2)) return true;
            }
        }
    return false;
}
__device__ __forceinline__ void collect_pairs_a(const RunParams& rp, const FrameGrid& g, unsigned cells_a, unsigned pts_a, int i,
                                                unsigned* pairs, int pair_cap, int* n_pairs) {
    const float4 p = lds_f4(pts_a + 16u * (unsigned)i);
    int cx, cy, cz;
    cell3<false>(g, p.x, p.y, p.z, cx, cy, cz);
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const int xl = max(cx - 1, 0), xh = min(cx + 1, nx - 1);
#pragma unroll 1
    for (int run = 0; run < 5; ++run) {         // own cell + 13 forward cells as five contiguous runs
        int s, e;
        if (run == 0) {
            s = i + 1;
            e = (int)lds_u16(cells_a + 2u * (unsigned)((cz * ny + cy) * nx + xh));
        } else {
            const int yy = run == 1 ? cy + 1 : cy + (run - 3);
            const int zz = run == 1 ? cz : cz + 1;
            if (yy < 0 || yy >= ny || zz >= nz) continue;
            const int row = (zz * ny + yy) * nx;
            s = (row + xl) ? (int)lds_u16(cells_a + 2u * (unsigned)(row + xl - 1)) : 0;
            e = (int)lds_u16(cells_a + 2u * (unsigned)(row + xh));
        }
        for (int q = s; q < e; ++q) {
            const float4 q4 = lds_f4(pts_a + 16u * (unsigned)q);
            if (dist2f(p.x, p.y, p.z, q4.x, q4.y, q4.z) > rp.r2f_hi || !dist_le(p.x, p.y, p.z, q4.x, q4.y, q4.z, rp.r2)) continue;
            const int k = atomicAdd(n_pairs, 1);
            HB_DASSERT(i >= 0 && i < 65536 && q > i && q < 65536);
            if (k < pair_cap) pairs[k] = (unsigned)i | ((unsigned)q << 16);
        }
    }
}

template <bool PBC>
__global__ void __launch_bounds__(FUSED_THREADS, PBC ? 3 : 4) k_frame_fused(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, FusedCfg fc) {
    extern __shared__ __align__(16) unsigned char smem[];
    FusedShared* sh = reinterpret_cast<FusedShared*>(smem);
    // regions by lifetime (plan_fused): [sorted1 | pairs] [ring | candidate coordinates | psorted, tasks, flags] [cand][lw]
    float4* sorted1 = reinterpret_cast<float4*>(smem + fc.off_sorted);     // selection points in g1 cell order (until the phase C sort)
    unsigned* pairs = reinterpret_cast<unsigned*>(smem + fc.off_sorted);   // phase C1 onwards: the pair list
    unsigned char* ring = smem + fc.off_ring;                              // phase B: tile stages
    float4* psorted = reinterpret_cast<float4*>(ring);                     // phase C: selection + local waters, g2 order
    uint2* tasks = reinterpret_cast<uint2*>(psorted + fc.sel_cap + fc.lw_cap);   // phase C: (pair | flag, hydrogen atom)
    unsigned* pflags = reinterpret_cast<unsigned*>(tasks + fc.task_cap);   // phase C: angle flags of the chunk's pairs
    float4* lw = reinterpret_cast<float4*>(smem + fc.off_lw);
    unsigned short* cells = reinterpret_cast<unsigned short*>(smem + fc.off_cells);
    unsigned* near = reinterpret_cast<unsigned*>(smem + fc.off_near);
    float4* cand = reinterpret_cast<float4*>(smem + fc.off_cand);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + fc.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const long long frame = bv.frame0 + b;
    const int s0 = tp.sys_sel_off[sys], nsel = tp.sys_sel_off[sys + 1] - s0;
    const int w0 = tp.sys_wat_off[sys], nwp = tp.sys_wat_off[sys + 1] - w0;
    const bool wa = rp.method == HB_METHOD_WATER_AROUND;
    const bool scan_waters = wa && nwp > 0 && nsel > 0;
    const bool stream = scan_waters && fc.regular;
    // selection points in gather order are only needed until the g1 sort, so they borrow the candidate buffers;
    // without a water scan there is no g1 sort and they go straight to sorted1
    float4* spos = scan_waters ? cand : sorted1;
    const int ntiles = (nwp + FUSED_TILE - 1) / FUSED_TILE;
    const int stride12 = fc.w_stride * 12;

    long long tick = bt.phase_cycles ? clock64() : 0;
    auto phase_done = [&](int k) {     // trace only: cycles thread 0 spent in phase k
        if (bt.phase_cycles && tid == 0) {
            const long long now = clock64();
            atomicAdd(&bt.phase_cycles[k], (unsigned long long)(now - tick));
            tick = now;
        }
    };
    auto cta_mark = [&](int k) {       // diagnostic only: wall-clock marks of consumer thread 0
        if (fc.cta_trace && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            fc.cta_trace[(size_t)b * 8 + k] = t;
            if (k == 0) {
                unsigned sm;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
                fc.cta_trace[(size_t)b * 8 + 6] = sm;
            }
        }
    };
    cta_mark(0);
    if (nsel > fc.sel_cap) {           // uniform over the CTA: this frame does not fit
        if (tid == 0) atomicExch(bt.fused_overflow, 1);
        return;
    }
    if (tid == 0) {
        sh->n_lw = 0;
        sh->n_cand = 0;
        sh->n_pairs = 0;
        sh->overflow = 0;
        for (int k = 0; k < fc.stages; ++k) mbar_init(&full[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (fc.pf_bytes) {                 // the selection's atom block -> L2 (phases A and C gather from it)
            unsigned long long a = ((unsigned long long)(uintptr_t)xyz + fc.pf_off) & ~15ull;
            unsigned long long e = ((unsigned long long)(uintptr_t)xyz + fc.pf_off + fc.pf_bytes + 15ull) & ~15ull;
            const unsigned long long l0 = (fc.buf_begin + 15ull) & ~15ull, l1 = fc.buf_end & ~15ull;
            if (a < l0) a = l0;
            if (e > l1) e = l1;
            if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
        }
        // All frames cost the same and the CTAs of a wave share HBM fairly, so a launch whose CTAs start together stays in
        // lock step: every resident frame streams at the same time (HBM saturated) and then every frame computes at the
        // same time (HBM idle). The first wave therefore starts spread over one frame time; the order persists.
        if (fc.stagger && (unsigned)b < fc.wave_ctas) {
            const unsigned per_sm = fc.wave_ctas / fc.n_sm;
            const unsigned long long rank = (unsigned long long)(((unsigned)b / fc.n_sm) % per_sm) * fc.n_sm + (unsigned)b % fc.n_sm;
            const long long delay = (long long)((unsigned long long)fc.stagger * rank / fc.wave_ctas);
            const long long t0 = clock64();
            while (clock64() - t0 < delay) __nanosleep(500);
        }
    }
    __syncthreads();
    cta_mark(1);

    // ---- the water block of this frame is streamed through the tile ring ------------------------------------------
    // A tile is 32 x FUSED_WPL consecutive waters. Warp w < stages owns stage w of the ring and the tiles w, w + stages, ...: it waits
    // for its tile, reads FUSED_WPL waters per lane, and - the stage being free again - one of its lanes issues the copy of its
    // next tile into it before the warp tests what it read. There is no producer warp and no "empty" barrier; the per-tile
    // costs (wait, refill, loop) are paid once per 256 waters, and the copies of the other stages are in flight meanwhile.
    // A tile advances by 256 * stride * 12 bytes, a multiple of 16, so every tile of the frame starts at the same offset
    // `head` inside its 16-byte aligned copy. Frames whose water block touches the edge of the batch buffer (the first and
    // last frame of a batch) copy the part of a tile that lies inside and read the few waters outside with plain loads.
    const unsigned long long gbase = stream ? (unsigned long long)(uintptr_t)(xyz + 3ll * tp.wat_atom[w0]) : 0ull;
    const unsigned tile_step = (unsigned)(FUSED_TILE * stride12);
    const unsigned long long start0 = gbase & ~15ull;
    const unsigned head = (unsigned)(gbase - start0);
    const unsigned long long lim_lo = (fc.buf_begin + 15ull) & ~15ull, lim_hi = fc.buf_end & ~15ull;
    const int nw_last = nwp - (ntiles - 1) * FUSED_TILE;
    const unsigned full_bytes = (head + (unsigned)((FUSED_TILE - 1) * stride12 + 12) + 15u) & ~15u;
    const unsigned last_bytes = (head + (unsigned)((nw_last - 1) * stride12 + 12) + 15u) & ~15u;
    const bool inside = stream && start0 >= lim_lo &&
                        start0 + (unsigned long long)(ntiles - 1) * tile_step + last_bytes <= lim_hi;
    auto issue_tile = [&](int t, int st) {      // one thread: copy tile t into stage st, pull a tile further ahead into L2 (any frame)
        unsigned char* dst = ring + (size_t)st * fc.stage_bytes;
        const unsigned long long src = start0 + (unsigned long long)t * tile_step;
        const unsigned bytes = t == ntiles - 1 ? last_bytes : full_bytes;
        if (inside) {
            mbar_expect_tx(&full[st], bytes);
            bulk_g2s(dst, reinterpret_cast<const void*>(src), bytes, &full[st]);
        } else {
            const unsigned long long end = src + bytes;
            const unsigned long long c0 = src > lim_lo ? src : lim_lo, c1 = end < lim_hi ? end : lim_hi;
            if (c1 > c0) {
                sh->stage_lo[st] = (unsigned)(c0 - src);
                sh->stage_hi[st] = (unsigned)(c1 - src);
                mbar_expect_tx(&full[st], (unsigned)(c1 - c0));
                bulk_g2s(dst + (c0 - src), reinterpret_cast<const void*>(c0), (unsigned)(c1 - c0), &full[st]);
            } else {
                sh->stage_lo[st] = 0;
                sh->stage_hi[st] = 0;
                mbar_arrive(&full[st]);
            }
        }
        const int tp2 = t + fc.prefetch;
        if (fc.prefetch > 0 && tp2 < ntiles) {
            unsigned long long a = start0 + (unsigned long long)tp2 * tile_step;
            unsigned long long e = a + (tp2 == ntiles - 1 ? last_bytes : tile_step);
            if (a < lim_lo) a = lim_lo;
            if (e > lim_hi) e = lim_hi;
            if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
        }
    };

    // the first tile of every stage is requested now, so that it lands while phase A runs (with the admission control on, it has
    // to wait until the frame is admitted)
    const bool early_issue = stream && fc.max_streams <= 0;
    if (early_issue && lane == 0 && warp < fc.stages && warp < ntiles) issue_tile(warp, warp);

    // ---- phase A: selection points, bounding box, grids ------------------------------------------
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    // two points per thread and trip: the atom indices are requested together, then the coordinates (two dependent round
    // trips per trip instead of two per point)
    for (int i0 = tid; i0 < nsel; i0 += 2 * FUSED_CONSUMERS) {
        int at[2];
        int4 kk4[2];
        float cx4[2], cy4[2], cz4[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = i0 + j * FUSED_CONSUMERS;
            if (i < nsel) { at[j] = tp.sel_atom[s0 + i]; kk4[j] = tp.sel_pack[s0 + i]; }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = i0 + j * FUSED_CONSUMERS;
            if (i < nsel) { const float* p = xyz + 3ll * at[j]; cx4[j] = p[0]; cy4[j] = p[1]; cz4[j] = p[2]; }
        }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = i0 + j * FUSED_CONSUMERS;
        if (i >= nsel) break;
        const int4 k = kk4[j];
        float x = cx4[j], y = cy4[j], z = cz4[j];
        const int multi = (k.x >= 0) + (k.y >= 0) + ((wa && k.z >= 0) ? 1 : 0);
        spos[i] = make_float4(x, y, z, __int_as_float(i | (multi > 1 ? SELF_TAG : 0)));
        if constexpr (PBC) {   // fractional coordinates relative to the first selection point, wrapped
            PbcGrid pg;
            pg.box = load_box(bv.box + 9ll * b);
            pg.ia = 1.0f / pg.box.ax; pg.ib = 1.0f / pg.box.by; pg.ic = 1.0f / pg.box.cz;
            const float* r = xyz + 3ll * tp.sel_atom[s0];
            float ra, rb, rc, sa, sb, sc;
            frac_coords(pg, r[0], r[1], r[2], ra, rb, rc);
            frac_coords(pg, x, y, z, sa, sb, sc);
            x = sa - ra; y = sb - rb; z = sc - rc;
            x -= rintf(x); y -= rintf(y); z -= rintf(z);
        }
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if (lane == 0)
        for (int k = 0; k < 3; ++k) { sh->red[warp][k] = mn[k]; sh->red[warp][3 + k] = mx[k]; }
    consumer_sync();
    // thread 0 builds the selection grid g1, thread 32 (another warp) the pair grid g2 at the same time: the grid sizing is a
    // serial loop of divisions
    const bool split_grids = !PBC && wa;
    if (tid == 0 || (tid == 32 && split_grids)) {
        // same values the general path obtains through its ordered-int atomics: min / max over the points
        float lo3[3], hi3[3];
        for (int k = 0; k < 3; ++k) {
            float a = sh->red[0][k], c = sh->red[0][3 + k];
            for (int w = 1; w < FUSED_CWARPS; ++w) { a = fminf(a, sh->red[w][k]); c = fmaxf(c, sh->red[w][3 + k]); }
            lo3[k] = a; hi3[k] = c;
        }
        if (tid == 0)
            for (int k = 0; k < 3; ++k) { sh->lo[k] = lo3[k]; sh->hi[k] = hi3[k]; }
        if constexpr (PBC) {
            PbcGrid pg;
            pg.box = load_box(bv.box + 9ll * b);
            pg.ia = 1.0f / pg.box.ax; pg.ib = 1.0f / pg.box.by; pg.ic = 1.0f / pg.box.cz;
            float sR[3] = {NAN, NAN, NAN};
            if (nsel > 0) {
                const float* r = xyz + 3ll * tp.sel_atom[s0];
                frac_coords(pg, r[0], r[1], r[2], sR[0], sR[1], sR[2]);
            }
            if (wa) {
                make_grid_pbc(sh->g1, pg.box, sR, lo3, hi3, rp.cell1, rp.cell1, fc.cell_cap);
                make_grid_pbc(sh->g2, pg.box, sR, lo3, hi3, rp.around_f, rp.cell2, fc.cell_cap);
            } else {
                make_grid_pbc(sh->g2, pg.box, sR, lo3, hi3, 0.f, rp.cell2, fc.cell_cap);
                sh->g1 = sh->g2;
            }
        } else if (wa) {
            if (tid == 0) make_grid(sh->g1, lo3, hi3, rp.cell1, rp.cell1, fc.cell_cap);
            else make_grid(sh->g2, lo3, hi3, rp.around_f, rp.cell2, fc.cell_cap);
        } else {
            make_grid(sh->g2, lo3, hi3, 0.f, rp.cell2, fc.cell_cap);
            sh->g1 = sh->g2;
        }
    }
    consumer_sync();

    phase_done(0);
    // ---- phase B: local waters ---------------------------------------------------------------------
    if (scan_waters) {
        // non-periodic grid geometry lives in registers; the larger minimum-image geometry stays in shared memory
        const FrameGrid g1 = sh->g1;
        for (int i = tid; i < (g1.ncell + 31) / 32; i += FUSED_CONSUMERS) near[i] = 0u;
        block_cell_sort<PBC>(g1, cells, sorted1, nsel, [&](int i) { return spos[i]; }, sh);   // (syncs cover the zeroing)
        for (int i = tid; i < nsel; i += FUSED_CONSUMERS) {
            const float4 p = sorted1[i];
            int cx, cy, cz;
            cell3<PBC>(g1, p.x, p.y, p.z, cx, cy, cz);
            if constexpr (PBC) {
                const AxisList zl = axis_neighbors(cz, g1.nz, g1.p.per[2]), yl = axis_neighbors(cy, g1.ny, g1.p.per[1]),
                               xl = axis_neighbors(cx, g1.nx, g1.p.per[0]);
                for (int a = 0; a < zl.n; ++a)
                    for (int bb = 0; bb < yl.n; ++bb)
                        for (int k = 0; k < xl.n; ++k) {
                            const int c = (zl.v[a] * g1.ny + yl.v[bb]) * g1.nx + xl.v[k];
                            atomicOr(&near[c >> 5], 1u << (c & 31));
                        }
            } else {
                // the (up to) three cells of a row are consecutive bits: one atomic per row, two where they straddle a word
                const int xl = max(cx - 1, 0), nb = min(cx + 1, g1.nx - 1) - xl + 1;
                for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g1.nz - 1); ++zz)
                    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g1.ny - 1); ++yy) {
                        const int r = (zz * g1.ny + yy) * g1.nx + xl;
                        const int sh5 = r & 31;
                        const unsigned m = (1u << nb) - 1u;
                        atomicOr(&near[r >> 5], m << sh5);
                        if (sh5 + nb > 32) atomicOr(&near[(r >> 5) + 1], m >> (32 - sh5));
                    }
            }
        }
        // Admission: at most max_streams frames of an SM stream at the same time. Frames that all cost the same drift into
        // lock step (every resident frame streams - HBM saturated - and then every frame computes - HBM idle); a frame that
        // finds its SM's share taken starts a little later and stays out of step from then on.
        unsigned my_sm = 0u;
        bool admitted = false;
        if (stream && fc.max_streams > 0 && tid == 0) {
            asm volatile("mov.u32 %0, %%smid;" : "=r"(my_sm));
            if (my_sm < fc.n_sm) {
                const long long t0 = clock64();
                for (;;) {
                    if (atomicAdd(&fc.sm_streams[my_sm], 1) < fc.max_streams) { admitted = true; break; }
                    atomicSub(&fc.sm_streams[my_sm], 1);
                    if (clock64() - t0 > 400000ll) break;             // (~0.2 ms: never wait for long)
                    __nanosleep(2000);
                }
            }
        }
        consumer_sync();
        phase_done(1);
        cta_mark(2);
        if (stream && !early_issue && lane == 0 && warp < fc.stages && warp < ntiles) issue_tile(warp, warp);

        const float a = rp.around_f;
        const float lox = sh->lo[0] - a, loy = sh->lo[1] - a, loz = sh->lo[2] - a;
        const float hix = sh->hi[0] + a, hiy = sh->hi[1] + a, hiz = sh->hi[2] + a;
        // Candidates (waters inside the dilated bounding box whose cell has a selection point in its neighbourhood) are
        // only COLLECTED while the tiles stream - their indices go to one CTA-wide list - and tested afterwards, densely,
        // by all consumers (phase B2). A warp that stopped to test 32 candidates in the middle of the stream held its
        // arrival on the next tiles back and, with a ring of three stages, everybody else's too.
        // candidate indices, 16-bit where every water index fits
        auto cand_put = [&](int pos, int i) {
            if (fc.cand16) reinterpret_cast<unsigned short*>(cand)[pos] = (unsigned short)i;
            else reinterpret_cast<unsigned*>(cand)[pos] = (unsigned)i;
        };
        auto cand_get = [&](int j) -> int {
            return fc.cand16 ? (int)reinterpret_cast<const unsigned short*>(cand)[j] : (int)reinterpret_cast<const unsigned*>(cand)[j];
        };
        const unsigned lt_mask = (1u << lane) - 1u;
        // cheap test first (non-periodic: dilated bounding box), then the near bitmap
        auto in_box = [&](float x, float y, float z) -> bool {
            if constexpr (PBC) return true;
            else return x >= lox && x <= hix && y >= loy && y <= hiy && z >= loz && z <= hiz;
        };
        auto near_hit = [&](float x, float y, float z) -> bool {
            int c;
            if constexpr (PBC) {
                int cx, cy, cz;
                if (!cell3_pbc(g1, x, y, z, cx, cy, cz)) return false;
                c = (cz * g1.ny + cy) * g1.nx + cx;
            } else {
                c = cell_of<false>(g1, x, y, z);
            }
            return (near[c >> 5] >> (c & 31)) & 1u;
        };
        if (stream) {
            const unsigned full_a = smem_u32(full);
            const unsigned lane_a = smem_u32(ring) + head + (unsigned)(lane * stride12);    // this lane's first water of stage 0
            const unsigned step32 = (unsigned)(32 * stride12);
            bool ok = true;
            // warp w < stages owns stage w: tiles w, w + stages, ... (the other warps wait at the barrier after the loop)
            const int st = warp;
            const unsigned bar = full_a + 8u * (unsigned)st;
            const unsigned a = lane_a + (unsigned)st * fc.stage_bytes;
            unsigned par = 0u;
            // refill of this warp's stage, frames inside the batch buffer: everything but the source address is loop invariant
            const unsigned stage_a = smem_u32(ring) + (unsigned)st * fc.stage_bytes;
            const unsigned long long src_step = (unsigned long long)fc.stages * tile_step;
            const unsigned long long pf_dist = (unsigned long long)fc.prefetch * tile_step;
            unsigned long long next_src = start0 + (unsigned long long)(warp + fc.stages) * tile_step;
            for (int t = warp; warp < fc.stages && t < ntiles; t += fc.stages, par ^= 1u) {
                if (!mbar_try_wait_a(bar, par) && !mbar_wait_slow_a(bar, par, fc.wait_ns)) { ok = false; break; }
                const int i0 = t * FUSED_TILE + lane;
                float x[FUSED_WPL], y[FUSED_WPL], z[FUSED_WPL];
#pragma unroll
                for (int k = 0; k < FUSED_WPL; ++k) {          // (always inside the stage; meaningful if the bytes were copied)
                    x[k] = lds_f32(a + k * step32); y[k] = lds_f32(a + k * step32 + 4u); z[k] = lds_f32(a + k * step32 + 8u);
                }
                if (!inside) {                          // bytes outside the batch buffer were not copied
                    const unsigned lo = sh->stage_lo[st], hi = sh->stage_hi[st];
#pragma unroll
                    for (int k = 0; k < FUSED_WPL; ++k) {
                        const unsigned off = head + (unsigned)((lane + 32 * k) * stride12);
                        if (i0 + 32 * k < nwp && !(off >= lo && off + 12u <= hi)) {
                            const float* g = xyz + 3ll * tp.wat_atom[w0 + i0 + 32 * k];
                            x[k] = g[0]; y[k] = g[1]; z[k] = g[2];
                        }
                    }
                }
                // Every lane's reads of the stage are ordered before the refill: the warp barrier orders the lanes' shared-memory
                // accesses, and the proxy fence orders them before the bulk copy's writes (async proxy).
                __syncwarp();
                if (lane == 0 && t + fc.stages < ntiles) {
                    const int u = t + fc.stages;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (inside) {
                        const unsigned bytes = u == ntiles - 1 ? last_bytes : full_bytes;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_a),
                                     "l"(next_src), "r"(bytes), "r"(bar) : "memory");
                        if (fc.prefetch > 0 && u + fc.prefetch < ntiles)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(next_src + pf_dist),
                                         "r"(u + fc.prefetch == ntiles - 1 ? last_bytes : tile_step) : "memory");
                        next_src += src_step;
                    } else {
                        issue_tile(u, st);
                    }
                }
                unsigned inb = 0u;
                if (t == ntiles - 1) {
#pragma unroll
                    for (int k = 0; k < FUSED_WPL; ++k)
                        if (i0 + 32 * k < nwp && in_box(x[k], y[k], z[k])) inb |= 1u << k;
                } else {
#pragma unroll
                    for (int k = 0; k < FUSED_WPL; ++k)
                        if (in_box(x[k], y[k], z[k])) inb |= 1u << k;
                }
                if (!__any_sync(0xffffffffu, inb != 0u)) continue;
                unsigned cm = 0u;
#pragma unroll
                for (int k = 0; k < FUSED_WPL; ++k)
                    if (((inb >> k) & 1u) && near_hit(x[k], y[k], z[k])) cm |= 1u << k;
                if (fc.debug & 1) cm = 0u;
                if (!__any_sync(0xffffffffu, cm != 0u)) continue;
                // append: warp scan of the per-lane counts, one shared-memory atomic per tile
                const int n = __popc(cm);
                int incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                int base = 0;
                if (lane == 31) base = atomicAdd(&sh->n_cand, incl);
                base = __shfl_sync(0xffffffffu, base, 31);
                int pos = base + incl - n;
                while (cm) {
                    const int k = __ffs(cm) - 1;
                    cm &= cm - 1u;
                    HB_DASSERT(pos >= 0 && i0 + 32 * k < nwp);
                    if (pos < fc.cand_cap) cand_put(pos, i0 + 32 * k);
                    else sh->overflow = 1;
                    ++pos;
                }
            }
            if (!ok) sh->overflow = 2;
        } else {
            // waters without a regular stride: index loads, one water per thread and trip
            for (int i0 = 0; i0 < nwp; i0 += FUSED_CONSUMERS) {
                const int i = i0 + tid;
                bool is_cand = false;
                if (i < nwp) {
                    const float* f = xyz + 3ll * tp.wat_atom[w0 + i];
                    const float x = f[0], y = f[1], z = f[2];
                    is_cand = in_box(x, y, z) && near_hit(x, y, z);
                }
                if (fc.debug & 1) is_cand = false;
                const unsigned m = __ballot_sync(0xffffffffu, is_cand);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sh->n_cand, __popc(m));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (is_cand) {
                        const int pos = base + __popc(m & lt_mask);
                        if (pos < fc.cand_cap) cand_put(pos, i);
                        else sh->overflow = 1;
                    }
                }
            }
        }
        phase_done(2);
        cta_mark(3);
        unsigned b2_t[5] = {0u, 0u, 0u, 0u, 0u}, b2_tk = 0u;                 // measurement (fused_debug & 4), warp 0
        const bool trb2 = (fc.debug & 4) && bt.phase_cycles && warp == 0;
        auto b2_lap = [&](int k) { if (trb2) { const unsigned now = (unsigned)clock(); b2_t[k] += now - b2_tk; b2_tk = now; } };
        if (trb2) b2_tk = (unsigned)clock();
        consumer_sync();
        if (admitted) atomicSub(&fc.sm_streams[my_sm], 1);      // (thread 0 only)
        b2_lap(0);
        // ---- phase B2: exact local-water test of the collected candidates, one per thread --------------------------------
        const int ncand = min(sh->n_cand, fc.cand_cap);
        // The candidates' coordinates were just streamed (L2 hits), but a round trip per round of 256 candidates was half
        // of this phase: all of them are gathered at once with asynchronous 4-byte copies into the drained tile ring.
        float* cxyz = reinterpret_cast<float*>(ring);
        const bool staged = stream && !sh->overflow && (size_t)ncand * 12 <= (size_t)fc.stages * fc.stage_bytes;
        if (staged) {
            for (int j = tid; j < ncand; j += FUSED_CONSUMERS) {
                const unsigned long long src = gbase + (unsigned long long)cand_get(j) * (unsigned long long)stride12;
                const unsigned dst = smem_u32(cxyz + 3 * j);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u), "l"(src + 4ull) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 8u), "l"(src + 8ull) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            consumer_sync();
        }
        b2_lap(1);
        const unsigned cells_a = smem_u32(cells), sorted1_a = smem_u32(sorted1);
        for (int j0 = 0; j0 < ncand; j0 += FUSED_CONSUMERS) {
            const int j = j0 + tid;
            bool local = false;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < ncand) {
                const int i = cand_get(j);
                HB_DASSERT(i >= 0 && i < nwp);
                if (staged) {
                    c = make_float4(cxyz[3 * j], cxyz[3 * j + 1], cxyz[3 * j + 2], __int_as_float(WATER_TAG | i));
                } else {
                    // (just streamed: an L2 hit; a regular water block needs no index load in front of it)
                    const float* f = stream ? reinterpret_cast<const float*>(gbase) + (long long)i * (3 * fc.w_stride)
                                            : xyz + 3ll * tp.wat_atom[w0 + i];
                    c = make_float4(f[0], f[1], f[2], __int_as_float(WATER_TAG | i));
                }
                b2_lap(2);
                if constexpr (PBC) local = water_is_local<true>(g1, cells, sorted1, c.x, c.y, c.z, rp);
                else local = water_is_local_a(g1, cells_a, sorted1_a, c.x, c.y, c.z, rp);
            }
            __syncwarp();
            b2_lap(3);
            const unsigned m = __ballot_sync(0xffffffffu, local);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sh->n_lw, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (local) {
                    const int pos = base + __popc(m & ((1u << lane) - 1u));
                    if (pos < fc.lw_cap) lw[pos] = c;
                    else sh->overflow = 1;
                }
            }
            b2_lap(4);
        }
        if (trb2 && lane == 0) {
            for (int k = 0; k < 5; ++k) atomicAdd(&bt.phase_cycles[30 + k], (unsigned long long)b2_t[k]);
            atomicAdd(&bt.phase_cycles[35], (unsigned long long)ncand);
        }
        phase_done(9);
    }
    consumer_sync();
    phase_done(3);
    cta_mark(4);
    if (sh->overflow || sh->n_lw > fc.lw_cap) {
        if (tid == 0) atomicExch(bt.fused_overflow, 1);
        return;
    }

    if (fc.debug & 2) return;
    // ---- phase C: pair search over selection + local waters --------------------------------------------
    const int nlw = sh->n_lw;
    const int np = nsel + nlw;
    block_cell_sort<PBC>(sh->g2, cells, psorted, np, [&](int i) { return i < nsel ? sorted1[i] : lw[i - nsel]; }, sh);
    int nhit = 0;
    const FrameGrid g2 = sh->g2;
    phase_done(4);
    const bool ang = rp.check_angle != 0;
    const float ct = rp.cos_threshold;
    // C1: every point walks its half shell and appends the point pairs within `distance` to one list
    for (int i = tid; i < np; i += FUSED_CONSUMERS) {
        if constexpr (PBC) collect_pairs<true>(rp, g2, cells, psorted, i, pairs, fc.pair_cap, &sh->n_pairs);
        else collect_pairs_a(rp, g2, smem_u32(cells), smem_u32(psorted), i, pairs, fc.pair_cap, &sh->n_pairs);
        if (rp.self_pairs) {
            const float4 p4 = psorted[i];
            if (__float_as_int(p4.w) & SELF_TAG) {
                PointInfo P;
                load_info(tp, rp, sys, p4, P);
                process_self<PBC>(tp, rp, bt, xyz, frame, P, nhit, g2.p.box);
            }
        }
    }
    phase_done(5);
    consumer_sync();
    const int npairs = sh->n_pairs;
    if (npairs > fc.pair_cap) {                         // uniform
        if (tid == 0) atomicExch(bt.fused_overflow, 1);
        return;
    }
    phase_done(6);
    // C2: pairs in chunks of 256, three dense steps per chunk
    for (int base = 0; base < npairs; base += FUSED_CONSUMERS) {
        const int k = base + tid;
        const bool have = k < npairs;
        PointInfo P, Q;
        unsigned rules = 0;
        if (tid == 0) sh->n_tasks = 0;
        pflags[tid] = 0u;
        consumer_sync();
        // (a) topology of the two points, entry pairs that exist, one task per hydrogen to test
        if (have) {
            const unsigned pq = pairs[k];
            load_info(tp, rp, sys, psorted[pq & 0xffffu], P);
            load_info(tp, rp, sys, psorted[pq >> 16], Q);
            rules = pair_rules(P, Q, rp.excl_bb != 0);
            if (ang) {
                const unsigned need = rules_need(rules);
                if (tp.h_wide) {                        // hydrogen lists too long for the inline tables: CSR walk here
                    unsigned f = 0;
                    if ((need & 1u) && any_h<PBC>(tp, xyz, P.don, P, Q, ct, g2.p.box)) f |= 1u;
                    if ((need & 2u) && any_h<PBC>(tp, xyz, Q.don, Q, P, ct, g2.p.box)) f |= 2u;
                    if ((need & 4u) && any_h<PBC>(tp, xyz, P.wat, P, Q, ct, g2.p.box)) f |= 4u;
                    if ((need & 8u) && any_h<PBC>(tp, xyz, Q.wat, Q, P, ct, g2.p.box)) f |= 8u;
                    pflags[tid] = f;
                } else {
                    const int lists[4] = {P.hd, Q.hd, P.hw, Q.hw};
                    const int atoms[4] = {P.atom, Q.atom, P.atom, Q.atom};
                    int nt = 0;
#pragma unroll
                    for (int f = 0; f < 4; ++f)
                        if (need & (1u << f)) nt += h_count(lists[f]);
                    int slot = nt ? atomicAdd(&sh->n_tasks, nt) : 0;
                    if (slot + nt <= fc.task_cap) {
#pragma unroll
                        for (int f = 0; f < 4; ++f)
                            if (need & (1u << f)) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int d = (int)(signed char)((unsigned)lists[f] >> (8 * j));
                                    HB_DASSERT(d == -128 || slot < fc.task_cap);
                                    if (d != -128) tasks[slot++] = make_uint2((unsigned)tid | ((unsigned)f << 16), (unsigned)(atoms[f] + d));
                                }
                            }
                    }
                }
            } else {
                pflags[tid] = 0xfu;
            }
        }
        phase_done(10);
        consumer_sync();
        phase_done(11);
        // (b) one thread per hydrogen: ANG(other, h, owner)
        const int ntasks = sh->n_tasks;
        if (ntasks > fc.task_cap) {                     // uniform
            if (tid == 0) atomicExch(bt.fused_overflow, 1);
            return;
        }
        // three tasks per thread and trip: their hydrogen coordinates (L2 / DRAM) are requested together, so a chunk pays one
        // round trip here instead of one per 256 tasks
        for (int t0 = tid; t0 < ntasks; t0 += 3 * FUSED_CONSUMERS) {
            uint2 tk[3];
            float hx[3], hy[3], hz[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int t = t0 + j * FUSED_CONSUMERS;
                if (t < ntasks) {
                    tk[j] = tasks[t];
                    const float* h = xyz + 3ll * (long long)tk[j].y;
                    hx[j] = h[0]; hy[j] = h[1]; hz[j] = h[2];
                }
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (t0 + j * FUSED_CONSUMERS >= ntasks) break;
                const unsigned kk = tk[j].x & 0xffffu, f = tk[j].x >> 16;
                HB_DASSERT(kk < FUSED_CONSUMERS && f < 4u && (int)(base + kk) < npairs);
                const unsigned pq = pairs[base + kk];
                const float4 a4 = psorted[pq & 0xffffu], b4 = psorted[pq >> 16];
                const float4 own = (f & 1u) ? b4 : a4, oth = (f & 1u) ? a4 : b4;
                if (ang_ok<PBC>(oth.x, oth.y, oth.z, hx[j], hy[j], hz[j], own.x, own.y, own.z, ct, g2.p.box)) atomicOr(&pflags[kk], 1u << f);
            }
        }
        phase_done(12);
        consumer_sync();
        phase_done(13);
        // (c) entry pairs whose hydrogens passed
        if (have) {
            unsigned m = rules & rules_pass(pflags[tid]);
            if (bt.sink) {                              // water-wire mode: edges go to the per-frame lists
                while (m) {
                    const int r = __ffs(m) - 1;
                    m &= m - 1;
                    int e0, e1;
                    rule_entries(P, Q, r, e0, e1);
                    emit(tp, rp, bt, frame, e0, e1, nhit);
                }
            }
            // Two entry pairs at a time (emit_bond of hb_common.cuh, hbond.py:119-127, in two steps): the entry tables of
            // both are requested together, then the first probes of the bond table - two dependent round trips per trip
            // instead of two per bond.
            while (m && !bt.sink) {
                unsigned long long key[2];
                unsigned hs[2];
                int ex[2], ey[2];
                int nk = 0;
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (m) {
                        const int r = __ffs(m) - 1;
                        m &= m - 1;
                        int e0, e1;
                        rule_entries(P, Q, r, e0, e1);
                        ex[j] = min(e0, e1); ey[j] = max(e0, e1);
                        nk = j + 1;
                    }
                int pxx[2], pxy[2], pyx[2], pyy[2], gx[2], gy[2];
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (j < nk) {
                        const int4 px = tp.ent_pack[ex[j]], py = tp.ent_pack[ey[j]];
                        pxx[j] = px.x; pxy[j] = px.y; pyx[j] = py.x; pyy[j] = py.y;
                        gx[j] = tp.ent_group[ex[j]]; gy[j] = tp.ent_group[ey[j]];
                    }
                unsigned long long cur[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    key[j] = EMPTY_KEY;
                    if (j < nk && (rp.check_angle || pxy[j] != pyy[j])) {          // hbond.py:125-127
                        const bool chk = pxx[j] < pyx[j];                          // hbond.py:120-122
                        key[j] = ((unsigned long long)(unsigned)(chk ? gx[j] : gy[j]) << 32) | (unsigned long long)(unsigned)(chk ? gy[j] : gx[j]);
                        hs[j] = hash64(key[j]) & bt.mask;
                        cur[j] = bt.keys[hs[j]];
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    if (key[j] != EMPTY_KEY) {
                        ++nhit;
                        record_bond_probed(bt, key[j], frame, hs[j], cur[j]);
                    }
            }
        }
        phase_done(14);
        consumer_sync();                                // pflags / tasks are rewritten by the next chunk
        phase_done(15);
    }
    phase_done(7);
    flush_hits(bt, nhit);
    phase_done(8);
    cta_mark(5);
}

}  // namespace

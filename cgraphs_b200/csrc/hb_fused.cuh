// hb_fused.cuh - fused fast path: ONE CTA processes ONE frame end to end, everything but the frame's
// coordinates and the bond table stays in shared memory.
//
//   phase A  gather the selection points (<= sel_cap), bounding box, cell grid g1, counting sort in smem
//   phase B  stream every water oxygen of the frame once: each warp pulls 32-water slabs of the
//            coordinate array into its own shared-memory ring with 1-D bulk async copies (TMA,
//            cp.async.bulk + mbarrier), tests them against the dilated bounding box and the g1 cells
//            (exact DIST predicate) and appends the local waters (hbond.py:249-252) to a shared list
//   phase C  cell grid g2 over selection + local waters, half-shell pair search with the exact DIST /
//            ANG predicates, key build, bond-table insert, occupancy bit
//
// HBM traffic per frame is the water slab read once plus a few gathers; several CTAs are resident per
// SM so the compute of one frame overlaps the streaming of others.  Frames that exceed the shared
// capacities (sel_cap / lw_cap) raise `fused_overflow`; the host then re-runs the batch on the general path.
#pragma once
#include "hb_common.cuh"

namespace {

#define FUSED_THREADS 256
#define FUSED_WARPS (FUSED_THREADS / 32)
#define FUSED_STAGES 2
#define FUSED_CHUNK 32   // waters per slab (one per lane)

struct FusedCfg {
    int sel_cap;              // selection points per frame that fit
    int lw_cap;               // local waters per frame that fit
    int cell_cap;             // cells per grid (packed 16-bit counters)
    int regular;              // 1: water point i is atom w_first + i * w_stride in every system -> TMA slabs
    int w_stride;             // atoms between consecutive water oxygens
    unsigned stage_bytes;     // bytes of one slab buffer (multiple of 16)
    unsigned off_spos, off_bc, off_lw, off_cells, off_bar;   // byte offsets in dynamic shared memory
    unsigned smem_bytes;
    unsigned long long buf_begin, buf_end;   // address range of the coordinate batch (bounds for bulk copies)
};

struct FusedShared {          // small fixed part at the start of dynamic shared memory
    FrameGrid g1, g2;
    float lo[3], hi[3];
    float red[FUSED_WARPS][6];
    int n_lw;
    int n_points;
    int overflow;
    int scan_carry;
    int warp_sum[FUSED_WARPS];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// counting-sort helpers on packed 16-bit cell counters ---------------------------------------------
__device__ __forceinline__ unsigned cell_add(unsigned* words, int c) {   // returns the old 16-bit value
    unsigned old = atomicAdd(&words[c >> 1], (c & 1) ? 0x10000u : 1u);
    return (c & 1) ? (old >> 16) : (old & 0xffffu);
}

// exclusive scan of n 16-bit counters in place (whole CTA); returns the total
__device__ int block_scan_cells(unsigned short* cells, int n, FusedShared* sh) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (n + FUSED_THREADS - 1) / FUSED_THREADS;
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
    __syncthreads();
    int base = 0;
    for (int k = 0; k < w; ++k) base += sh->warp_sum[k];
    int run = base + incl - sum;
    for (int i = s; i < e; ++i) {
        int v = cells[i];
        cells[i] = (unsigned short)run;
        run += v;
    }
    int total = 0;
    for (int k = 0; k < FUSED_WARPS; ++k) total += sh->warp_sum[k];
    __syncthreads();
    return total;
}

// sort `n` points (src(i)) into the cells of g: zero, count, scan, scatter. After return cells[c] is the END of cell c.
template <typename Src>
__device__ void block_cell_sort(const FrameGrid& g, unsigned short* cells, float4* dst, int n, Src src, FusedShared* sh) {
    unsigned* words = reinterpret_cast<unsigned*>(cells);
    const int nwords = (g.ncell + 2) >> 1;
    for (int i = threadIdx.x; i < nwords; i += FUSED_THREADS) words[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += FUSED_THREADS) {
        float4 p = src(i);
        cell_add(words, cell_of(g, p.x, p.y, p.z));
    }
    __syncthreads();
    block_scan_cells(cells, g.ncell, sh);
    for (int i = threadIdx.x; i < n; i += FUSED_THREADS) {
        float4 p = src(i);
        dst[cell_add(words, cell_of(g, p.x, p.y, p.z))] = p;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FUSED_THREADS) k_frame_fused(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, FusedCfg fc) {
    extern __shared__ __align__(16) unsigned char smem[];
    FusedShared* sh = reinterpret_cast<FusedShared*>(smem);
    float4* spos = reinterpret_cast<float4*>(smem + fc.off_spos);        // selection points, gather order
    float4* sorted1 = reinterpret_cast<float4*>(smem + fc.off_bc);       // selection points in g1 cell order
    unsigned char* stages = smem + fc.off_bc + (size_t)fc.sel_cap * 16;  // per-warp slab rings (phase B)
    float4* psorted = reinterpret_cast<float4*>(smem + fc.off_bc);       // phase C: aliases sorted1 + rings
    float4* lw = reinterpret_cast<float4*>(smem + fc.off_lw);
    unsigned short* cells = reinterpret_cast<unsigned short*>(smem + fc.off_cells);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + fc.off_bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const long long frame = bv.frame0 + b;
    const int s0 = tp.sys_sel_off[sys], nsel = tp.sys_sel_off[sys + 1] - s0;
    const int w0 = tp.sys_wat_off[sys], nwp = tp.sys_wat_off[sys + 1] - w0;
    const bool wa = rp.method == HB_METHOD_WATER_AROUND;

    if (tid == 0) {
        sh->n_lw = 0;
        sh->overflow = 0;
        for (int k = 0; k < FUSED_WARPS * FUSED_STAGES; ++k) mbar_init(&bars[k], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (nsel > fc.sel_cap) {           // uniform over the CTA: this frame does not fit
        if (tid == 0) atomicExch(bt.fused_overflow, 1);
        return;
    }

    // ---- phase A: selection points, bounding box, grids ------------------------------------------
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = tid; i < nsel; i += FUSED_THREADS) {
        const float* p = xyz + 3ll * tp.sel_atom[s0 + i];
        float x = p[0], y = p[1], z = p[2];
        spos[i] = make_float4(x, y, z, __int_as_float(i));
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
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
    __syncthreads();
    if (tid == 0) {
        // same values the general path obtains through its ordered-int atomics: min / max over the points
        for (int k = 0; k < 3; ++k) {
            float a = sh->red[0][k], c = sh->red[0][3 + k];
            for (int w = 1; w < FUSED_WARPS; ++w) { a = fminf(a, sh->red[w][k]); c = fmaxf(c, sh->red[w][3 + k]); }
            sh->lo[k] = a; sh->hi[k] = c;
        }
        if (wa) {
            make_grid(sh->g1, sh->lo, sh->hi, rp.cell1, rp.cell1, fc.cell_cap);
            make_grid(sh->g2, sh->lo, sh->hi, rp.around_f, rp.cell2, fc.cell_cap);
        } else {
            make_grid(sh->g2, sh->lo, sh->hi, 0.f, rp.cell2, fc.cell_cap);
            sh->g1 = sh->g2;
        }
    }
    __syncthreads();

    // ---- phase B: local waters ---------------------------------------------------------------------
    if (wa && nwp > 0 && nsel > 0) {
        block_cell_sort(sh->g1, cells, sorted1, nsel, [&](int i) { return spos[i]; }, sh);
        const FrameGrid g1 = sh->g1;
        const float a = rp.around_f;
        const float lox = sh->lo[0] - a, loy = sh->lo[1] - a, loz = sh->lo[2] - a;
        const float hix = sh->hi[0] + a, hiy = sh->hi[1] + a, hiz = sh->hi[2] + a;
        auto test_and_append = [&](bool live, float x, float y, float z, int i) {
            bool local = false;
            if (live && x >= lox && x <= hix && y >= loy && y <= hiy && z >= loz && z <= hiz)
                local = water_is_local(g1, cells, sorted1, x, y, z, rp);
            unsigned m = __ballot_sync(0xffffffffu, local);
            if (m) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&sh->n_lw, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (local) {
                    int pos = base + __popc(m & ((1u << lane) - 1u));
                    if (pos < fc.lw_cap) lw[pos] = make_float4(x, y, z, __int_as_float(WATER_TAG | i));
                    else sh->overflow = 1;
                }
            }
        };
        const int nchunk = (nwp + FUSED_CHUNK - 1) / FUSED_CHUNK;
        if (fc.regular) {
            // slab k of this warp = waters [32 c, 32 c + 32), c = warp + 8 k; bytes from the first to the last oxygen
            const unsigned char* gbase = reinterpret_cast<const unsigned char*>(xyz + 3ll * tp.wat_atom[w0]);
            const long long slab_stride = (long long)FUSED_CHUNK * fc.w_stride * 12;
            unsigned char* ring = stages + (size_t)warp * FUSED_STAGES * fc.stage_bytes;
            unsigned long long* wbar = bars + warp * FUSED_STAGES;
            const int nk = warp < nchunk ? (nchunk - warp + FUSED_WARPS - 1) / FUSED_WARPS : 0;
            auto slab = [&](int k, unsigned long long& start, unsigned& size, unsigned& head, int& nw) {
                int c = warp + k * FUSED_WARPS;
                nw = min(FUSED_CHUNK, nwp - c * FUSED_CHUNK);
                unsigned long long ga = (unsigned long long)(gbase + (long long)c * slab_stride);
                unsigned nbytes = (unsigned)(((nw - 1) * fc.w_stride * 3 + 3) * 4);
                start = ga & ~15ull;
                head = (unsigned)(ga - start);
                size = (head + nbytes + 15u) & ~15u;
                return start >= fc.buf_begin && start + size <= fc.buf_end && size <= fc.stage_bytes;
            };
            int issued = 0;      // TMA slabs issued so far (ring position), consumed likewise
            int k_issue = 0;
            auto issue_next = [&]() {    // lane 0: queue the next TMA-able slab, if any
                while (k_issue < nk) {
                    unsigned long long start; unsigned size, head; int nw;
                    bool ok = slab(k_issue, start, size, head, nw);
                    ++k_issue;
                    if (ok) {
                        int st = issued % FUSED_STAGES;
                        mbar_expect_tx(&wbar[st], size);
                        bulk_g2s(ring + (size_t)st * fc.stage_bytes, reinterpret_cast<const void*>(start), size, &wbar[st]);
                        ++issued;
                        return;
                    }
                }
            };
            __syncthreads();     // barrier init visible to every warp before the first wait
            if (lane == 0)
                for (int s = 0; s < FUSED_STAGES; ++s) issue_next();
            int consumed = 0;
            for (int k = 0; k < nk; ++k) {
                unsigned long long start; unsigned size, head; int nw;
                bool ok = slab(k, start, size, head, nw);
                const int c = warp + k * FUSED_WARPS;
                const int i = c * FUSED_CHUNK + lane;
                const bool live = lane < nw;
                float x = 0.f, y = 0.f, z = 0.f;
                if (ok) {
                    int st = consumed % FUSED_STAGES;
                    unsigned parity = (unsigned)((consumed / FUSED_STAGES) & 1);
                    unsigned spins = 0;
                    while (!mbar_try_wait(&wbar[st], parity)) {
                        if (++spins > (1u << 26)) { sh->overflow = 2; break; }     // never hang the GPU
                    }
                    if (live) {
                        const float* f = reinterpret_cast<const float*>(ring + (size_t)st * fc.stage_bytes + head) +
                                         lane * fc.w_stride * 3;
                        x = f[0]; y = f[1]; z = f[2];
                    }
                    ++consumed;
                    __syncwarp();                     // every lane has read its water before the slot is refilled
                    if (lane == 0) issue_next();
                } else if (live) {                    // slab touches the edge of the batch buffer: plain loads
                    const float* f = xyz + 3ll * tp.wat_atom[w0 + i];
                    x = f[0]; y = f[1]; z = f[2];
                }
                test_and_append(live, x, y, z, i);
            }
        } else {
            for (int c = warp; c < nchunk; c += FUSED_WARPS) {
                const int i = c * FUSED_CHUNK + lane;
                const bool live = i < nwp;
                float x = 0.f, y = 0.f, z = 0.f;
                if (live) {
                    const float* f = xyz + 3ll * tp.wat_atom[w0 + i];
                    x = f[0]; y = f[1]; z = f[2];
                }
                test_and_append(live, x, y, z, i);
            }
        }
    }
    __syncthreads();
    if (sh->overflow || sh->n_lw > fc.lw_cap) {
        if (tid == 0) atomicExch(bt.fused_overflow, 1);
        return;
    }

    // ---- phase C: pair search over selection + local waters --------------------------------------------
    const int nlw = sh->n_lw;
    const int np = nsel + nlw;
    block_cell_sort(sh->g2, cells, psorted, np, [&](int i) { return i < nsel ? spos[i] : lw[i - nsel]; }, sh);
    int nhit = 0;
    const FrameGrid g2 = sh->g2;
    for (int i = tid; i < np; i += FUSED_THREADS)
        search_point(tp, rp, bt, xyz, frame, sys, g2, cells, psorted, i, nhit);
    flush_hits(bt, nhit);
}

}  // namespace

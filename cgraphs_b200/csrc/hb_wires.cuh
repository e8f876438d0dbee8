// hb_wires.cuh - water wires (SURVEY 8f "next" #1; reference WireAnalysis.set_water_wires, waterwire.py:191-317).
//
// The neighbour stage (waterwire.py:200-248) is the search of hb_fused.cuh / hb_general.cuh run with the around
// radius (max_water + 1) * distance / 2 and without the backbone filter; in wire mode `emit_bond` appends every
// H-bond entry pair of a frame to that frame's edge list. k_wire_bfs then does what waterwire.py:250-309 does with
// the lists of one frame:
//
//   nodes   residues (`da_trans`: all selection entries with one residue-level id are one node, numbered in entry
//           order) and local waters
//   wires   for every residue s, breadth first over its own residue-water edges and all water-water edges down to
//           depth max_water; a residue t > s bonded to a water of depth k is joined to s by a wire of k waters;
//           the smallest k counts
//   direct  (acceptor, donor) pairs are wires of 0 waters
//
// All sources advance together: a water carries one bit per source residue (seen / frontier / next), a level is
// one pass over the water-water edges OR-ing frontier words into the neighbours, and the pairs found at level k
// are the new bits that reach the waters next to a target residue. Sources are taken in chunks of at most
// 32 * WIRE_CHUNK_WORDS so that the bit sets stay small. One CTA (1024 threads, one per SM) works on one frame; the
// bit sets live in its shared memory - the source chunk is narrowed until they fit - and only for frames with
// more than ~17,000 local waters in a per-CTA scratch area in global memory.
//
// Results go into the bond table with byte-wide columns: row word (frame >> 2), byte (frame & 3) = k (waters of the
// shortest wire, 1..127, 0 = none) | 0x80 if a direct bond exists; the last two words of a row hold
// ~min((frame << 32) | (acceptor entry << 1) | donor_is_second_node) over the direct bonds (0 = none), which is
// what the host needs to orient the key the way the reference's first discovery does (waterwire.py:277-279,299-301).
#pragma once
#include "hb_common.cuh"

namespace {

#define WIRE_CHUNK_WORDS 32
#define WIRE_THREADS 1024
#define WIRE_QUEUE 4096
#define WIRE_SMEM_WORDS (50 * 1024)      // 200 kB of bit sets per CTA (one CTA per SM)

struct WireCfg {
    const int* ent_node;     // [n_ent] residue node (< n_res) of a selection entry, n_res + water index of a water entry
    const int* direct_node;  // [n_ent] node a DIRECT bond of a selection entry is booked on (residuewise=False: the atom's
                             // own node, while wires are still booked on the node of the residue's first entry); may equal ent_node
    int n_res, n_wat;
    int max_water, allow_direct;
    int nb4;                 // row words that hold the per-frame bytes (the "first direct" tail follows)
    int cap_words;           // min(ceil(n_res / 32), WIRE_CHUNK_WORDS): words per node the scratch is sized for
    int smem_words;          // shared-memory budget of the bit sets (WIRE_SMEM_WORDS; smaller values are a test hook)
    // per-CTA scratch
    int* wl_of;              // [n_cta][n_wat]   water -> local index, -1 when idle
    int* lw_list;            // [n_cta][n_wat]
    int* ar_of;              // [n_cta][n_res]   residue node -> active index, -1 when idle
    int* ar_list;            // [n_cta][n_res]
    unsigned* wbits;         // [n_cta][3][n_wat * cap_words]   seen / frontier / next
    unsigned* tbits;         // [n_cta][2][n_res * cap_words]   found / level accumulator
    int2* rlist;             // [n_cta][edge_cap]               the frame's edges in local indices
    const int* status;       // status block: [2] fused_overflow, [3] pair_overflow, [6] wire_overflow, [7] frame-pairs overflow
};

__global__ void k_set_sink(EdgeSink* dst, EdgeSink v) { *dst = v; }

// ---- convex-hull option of the water wires (waterwire.py:218-228) -------------------------------------------------------
// The reference keeps only the local waters inside the Delaunay triangulation of the selection (scipy / Qhull, done on
// the host exactly as the reference does it) but goes on naming them by the UNFILTERED list: the j-th in-hull water
// (`true_w[j]`: its position enters the distance searches) is called `label_w[j]`, the j-th local water, and the angle
// criterion is evaluated with the labelled water's coordinates and hydrogens. The host hands over both index lists
// per frame; this kernel does the selection-water and water-water searches over them (a few hundred waters per frame:
// all pairs, no grid) and feeds the same pair expansion / edge lists as the ordinary search.
struct VirtualWaters {
    const long long* off;     // [n_frames + 1] offsets into true_w / label_w, indexed by the global frame number
    const int* true_w;        // water point whose position is searched
    const int* label_w;       // water point the pair is attributed to
    long long n_frames;
};

__global__ void __launch_bounds__(256) k_virtual_pairs(BatchView bv, DevTopo tp, RunParams rp, BondTable bt, VirtualWaters vw) {
    const int b = blockIdx.x;
    int sys;
    const float* xyz = frame_xyz(bv, tp, b, sys);
    const long long frame = bv.frame0 + b;
    if (frame >= vw.n_frames) return;
    const long long o0 = vw.off[frame];
    const int n = (int)(vw.off[frame + 1] - o0);
    const int* tw = vw.true_w + o0;
    const int* lw = vw.label_w + o0;
    const int s0 = tp.sys_sel_off[sys], nsel = tp.sys_sel_off[sys + 1] - s0, w0 = tp.sys_wat_off[sys];
    int nhit = 0;
    const Box6 nobox{};
    auto water_info = [&](int w, PointInfo& q) {
        const float* p = xyz + 3ll * tp.wat_atom[w0 + w];
        load_info(tp, rp, sys, make_float4(p[0], p[1], p[2], __int_as_float(WATER_TAG | w)), q);
    };
    for (long long k = threadIdx.x; k < (long long)nsel * n; k += blockDim.x) {          // selection - water
        const int s = (int)(k / n), j = (int)(k - (long long)s * n);
        const float* ps = xyz + 3ll * tp.sel_atom[s0 + s];
        const float* pt = xyz + 3ll * tp.wat_atom[w0 + tw[j]];
        if (dist2f(ps[0], ps[1], ps[2], pt[0], pt[1], pt[2]) > rp.r2f_hi || !dist_le(ps[0], ps[1], ps[2], pt[0], pt[1], pt[2], rp.r2)) continue;
        PointInfo P, Q;
        load_info(tp, rp, sys, make_float4(ps[0], ps[1], ps[2], __int_as_float(s)), P);
        water_info(lw[j], Q);
        process_pair<false>(tp, rp, bt, xyz, frame, P, Q, nhit, nobox);
    }
    for (long long k = threadIdx.x; k < (long long)n * n; k += blockDim.x) {             // water - water
        const int i = (int)(k / n), j = (int)(k - (long long)i * n);
        if (i >= j) continue;
        const float* pa = xyz + 3ll * tp.wat_atom[w0 + tw[i]];
        const float* pb = xyz + 3ll * tp.wat_atom[w0 + tw[j]];
        if (dist2f(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2]) > rp.r2f_hi || !dist_le(pa[0], pa[1], pa[2], pb[0], pb[1], pb[2], rp.r2)) continue;
        PointInfo P, Q;
        water_info(lw[i], P);
        water_info(lw[j], Q);
        process_pair<false>(tp, rp, bt, xyz, frame, P, Q, nhit, nobox);
    }
    flush_hits(bt, nhit);
}

__device__ __forceinline__ int table_slot(const BondTable& bt, unsigned long long key) {
    unsigned h = hash64(key) & bt.mask;
    for (unsigned probe = 0; probe <= bt.mask; ++probe) {
        unsigned long long cur = bt.keys[h];
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&bt.keys[h], EMPTY_KEY, key);
            if (cur == EMPTY_KEY) { atomicAdd(bt.n_keys, 1); cur = key; }
        }
        if (cur == key) return (int)h;
        h = (h + 1) & bt.mask;
    }
    atomicExch(bt.overflow, 1);
    return -1;
}

__device__ __forceinline__ void wire_record(const BondTable& bt, int s, int t, long long frame, unsigned byte) {
    const int h = table_slot(bt, ((unsigned long long)(unsigned)s << 32) | (unsigned)t);
    if (h >= 0) atomicOr(&bt.bits[(size_t)h * bt.wpr + (frame >> 2)], byte << (8 * (int)(frame & 3)));
}

__global__ void __launch_bounds__(WIRE_THREADS) k_wire_bfs(BondTable bt, WireCfg wc, const int2* edges, const int* edge_cnt,
                                                           int edge_cap, long long frame0, int n_frames) {
    // an incomplete edge list (a capacity overflow somewhere in the search) would give wires that are too long:
    // the host re-runs the batch after growing whatever overflowed, so this launch does nothing
    if (wc.status[2] | wc.status[3] | wc.status[6] | wc.status[7]) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n_res = wc.n_res;
    int* wl_of = wc.wl_of + (size_t)blockIdx.x * wc.n_wat;
    int* lw = wc.lw_list + (size_t)blockIdx.x * wc.n_wat;
    int* ar_of = wc.ar_of + (size_t)blockIdx.x * n_res;
    int* ar = wc.ar_list + (size_t)blockIdx.x * n_res;
    int2* R = wc.rlist + (size_t)blockIdx.x * edge_cap;
    const size_t wstride = (size_t)wc.n_wat * wc.cap_words, tstride = (size_t)n_res * wc.cap_words;
    extern __shared__ unsigned s_bits[];
    // HBGPU_TRACE: cycles thread 0 spends per phase (slots 16.. of the status block's cycle counters)
    long long t_last = bt.phase_cycles ? clock64() : 0;
    auto phase = [&](int slot) {
        if (bt.phase_cycles && threadIdx.x == 0) {
            const long long t = clock64();
            atomicAdd(&bt.phase_cycles[16 + slot], (unsigned long long)(t - t_last));
            t_last = t;
        }
    };
    __shared__ int s_nlw, s_nar, s_nsw, s_nww, s_any, s_nq;
    __shared__ unsigned s_queue[WIRE_QUEUE];      // pairs found at the current level: (source node index << 16) | target index

    for (int b = blockIdx.x; b < n_frames; b += gridDim.x) {
        const int ne = min(edge_cnt[b], edge_cap);
        const int2* E = edges + (size_t)b * edge_cap;
        const long long frame = frame0 + b;
        if (tid == 0) { s_nlw = 0; s_nar = 0; s_nsw = 0; s_nww = 0; }
        __syncthreads();
        // local waters and active residues (those with a residue-water edge) get dense indices; direct bonds
        // (acceptor entry, donor entry) are recorded on the way
        for (int i0 = tid; i0 < ne; i0 += 4 * nt) {           // four edges per thread in flight
            int2 e4[4];
            int n04[4], n14[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e4[u] = (i0 + u * nt < ne) ? E[i0 + u * nt] : make_int2(-1, -1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                n04[u] = e4[u].x >= 0 ? wc.ent_node[e4[u].x] : 0;
                n14[u] = e4[u].x >= 0 ? wc.ent_node[e4[u].y] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
            if (e4[u].x < 0) continue;
            const int2 e = e4[u];
            const int n0 = n04[u], n1 = n14[u];
            HB_DASSERT(n0 >= 0 && n1 >= 0 && n0 < n_res + wc.n_wat && n1 < n_res + wc.n_wat);
            if (n0 >= n_res && atomicCAS(&wl_of[n0 - n_res], -1, -2) == -1) lw[atomicAdd(&s_nlw, 1)] = n0 - n_res;
            if (n1 >= n_res && atomicCAS(&wl_of[n1 - n_res], -1, -2) == -1) lw[atomicAdd(&s_nlw, 1)] = n1 - n_res;
            if ((n0 >= n_res) != (n1 >= n_res)) {
                const int r = n0 < n_res ? n0 : n1;
                if (atomicCAS(&ar_of[r], -1, -2) == -1) ar[atomicAdd(&s_nar, 1)] = r;
            } else if (n0 < n_res && wc.allow_direct) {
                const int d0 = wc.direct_node[e.x], d1 = wc.direct_node[e.y];
                const int lo = min(d0, d1), hi = max(d0, d1);
                const int h = table_slot(bt, ((unsigned long long)(unsigned)lo << 32) | (unsigned)hi);
                if (h >= 0) {
                    unsigned* row = bt.bits + (size_t)h * bt.wpr;
                    atomicOr(&row[frame >> 2], 0x80u << (8 * (int)(frame & 3)));
                    const unsigned long long packed = ((unsigned long long)frame << 32) | ((unsigned long long)(unsigned)e.x << 1) |
                                                      (unsigned long long)(d1 == hi ? 1 : 0);
                    atomicMax((unsigned long long*)(row + wc.nb4), ~packed);
                }
            }
            }
        }
        __syncthreads();
        phase(0);
        const int nlw = s_nlw, nar = s_nar;
        for (int i = tid; i < nlw; i += nt) wl_of[lw[i]] = i;
        for (int i = tid; i < nar; i += nt) ar_of[ar[i]] = i;
        __syncthreads();
        // the edges in local indices: residue-water (active residue, local water) from the front of R,
        // water-water (local water, local water) from its back
        for (int i0 = tid; i0 < ne; i0 += 4 * nt) {
            int2 e4[4];
            int n04[4], n14[4], a4[4], b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e4[u] = (i0 + u * nt < ne) ? E[i0 + u * nt] : make_int2(-1, -1);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                n04[u] = e4[u].x >= 0 ? wc.ent_node[e4[u].x] : 0;
                n14[u] = e4[u].x >= 0 ? wc.ent_node[e4[u].y] : 0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {                     // local ids of both ends (water / active residue)
                a4[u] = b4[u] = -1;
                if (e4[u].x < 0) continue;
                const bool w0 = n04[u] >= n_res, w1 = n14[u] >= n_res;
                if (!w0 && !w1) continue;
                a4[u] = w0 ? wl_of[n04[u] - n_res] : ar_of[n04[u]];
                b4[u] = w1 ? wl_of[n14[u] - n_res] : ar_of[n14[u]];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (a4[u] < 0) continue;
                const bool w0 = n04[u] >= n_res, w1 = n14[u] >= n_res;
                if (w0 && w1) R[edge_cap - 1 - atomicAdd(&s_nww, 1)] = make_int2(a4[u], b4[u]);
                else R[atomicAdd(&s_nsw, 1)] = w0 ? make_int2(b4[u], a4[u]) : make_int2(a4[u], b4[u]);
            }
        }
        __syncthreads();
        phase(1);
        const int nsw = s_nsw, nww = s_nww;
        const int2* SW = R;
        const int2* WW = R + (edge_cap - nww);

        if (nar > 0 && wc.max_water >= 1) {
            // bit sets: shared memory when one word per node fits, with the widest source chunk that does
            const int per_word = 3 * nlw + 2 * nar;
            const bool in_smem = per_word <= wc.smem_words;
            const int words = min((nar + 31) >> 5, in_smem ? min(wc.smem_words / per_word, WIRE_CHUNK_WORDS) : wc.cap_words);
            const int chunk = words * 32;
            const int nw = nlw * words, ntw = nar * words;
            unsigned* seen = in_smem ? s_bits : wc.wbits + (size_t)blockIdx.x * 3 * wstride;
            unsigned* cur = in_smem ? seen + nw : seen + wstride;
            unsigned* nxt = in_smem ? cur + nw : cur + wstride;
            unsigned* found = in_smem ? nxt + nw : wc.tbits + (size_t)blockIdx.x * 2 * tstride;
            unsigned* acc = in_smem ? found + ntw : found + tstride;
            for (int base = 0; base < nar; base += chunk) {
                for (int i = tid; i < nw; i += nt) { cur[i] = 0; nxt[i] = 0; }
                for (int i = tid; i < ntw; i += nt) { found[i] = 0; acc[i] = 0; }
                __syncthreads();
                // depth 1: the waters bonded to each source of this chunk
                for (int i = tid; i < nsw; i += nt) {
                    const int2 e = SW[i];
                    const int s = e.x - base;
                    if (s >= 0 && s < chunk) atomicOr(&cur[(size_t)e.y * words + (s >> 5)], 1u << (s & 31));
                }
                __syncthreads();
                for (int i = tid; i < nw; i += nt) seen[i] = cur[i];
                phase(2);
                for (int k = 1; k <= wc.max_water; ++k) {        // (acc and nxt are zero here: their readers clear them)
                    if (tid == 0) s_nq = 0;
                    __syncthreads();
                    // sources whose frontier (depth k) holds a water next to residue t
                    for (int i = tid; i < nsw * words; i += nt) {
                        const int q = i / words, j = i - q * words;
                        const int2 e = SW[q];
                        const unsigned v = cur[(size_t)e.y * words + j];
                        if (v) atomicOr(&acc[(size_t)e.x * words + j], v);
                    }
                    __syncthreads();
                    phase(3);
                    for (int i = tid; i < ntw; i += nt) {
                        const unsigned av = acc[i];
                        if (!av) continue;
                        acc[i] = 0;
                        unsigned nb = av & ~found[i];
                        if (!nb) continue;
                        found[i] |= nb;
                        const int t = i / words, j = i - t * words;
                        const int tn = ar[t];
                        while (nb) {            // every pair is reported from its smaller node
                            const int bit = __ffs(nb) - 1;
                            nb &= nb - 1;
                            const int si = base + j * 32 + bit;
                            if (ar[si] >= tn) continue;
                            const int q = (nar <= 65536) ? atomicAdd(&s_nq, 1) : WIRE_QUEUE;
                            if (q < WIRE_QUEUE) s_queue[q] = ((unsigned)si << 16) | (unsigned)t;
                            else wire_record(bt, ar[si], tn, frame, (unsigned)k);
                        }
                    }
                    __syncthreads();
                    phase(4);
                    // the table updates (hash probe + atomic in global memory) spread over all threads
                    for (int i = tid; i < min(s_nq, WIRE_QUEUE); i += nt) {
                        const unsigned q = s_queue[i];
                        wire_record(bt, ar[q >> 16], ar[q & 0xffffu], frame, (unsigned)k);
                    }
                    phase(5);
                    if (k == wc.max_water) break;
                    // one step along the water-water edges
                    if (tid == 0) s_any = 0;        // (read two barriers ago, written after the next one)
                    // (one edge per thread, its words in an inner loop: the edge list is read from L2 once per level and
                    // the loads of a thread's edges are issued back to back)
                    for (int q0 = tid; q0 < nww; q0 += 4 * nt) {
                        int2 e[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) e[u] = (q0 + u * nt < nww) ? WW[q0 + u * nt] : make_int2(-1, -1);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (e[u].x < 0) continue;
                            const unsigned* ca = cur + (size_t)e[u].x * words;
                            const unsigned* cc = cur + (size_t)e[u].y * words;
                            unsigned* na = nxt + (size_t)e[u].x * words;
                            unsigned* nc = nxt + (size_t)e[u].y * words;
                            for (int j = 0; j < words; ++j) {
                                const unsigned va = ca[j], vc = cc[j];
                                if (va) atomicOr(&nc[j], va);
                                if (vc) atomicOr(&na[j], vc);
                            }
                        }
                    }
                    __syncthreads();
                    phase(6);
                    int any = 0;
                    for (int i = tid; i < nw; i += nt) {
                        const unsigned nv = nxt[i];
                        if (nv) nxt[i] = 0;
                        const unsigned n = nv & ~seen[i];
                        seen[i] |= n;
                        cur[i] = n;
                        any |= n != 0;
                    }
                    if (any) s_any = 1;
                    __syncthreads();
                    phase(7);
                    if (!s_any) break;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        for (int i = tid; i < nlw; i += nt) wl_of[lw[i]] = -1;
        for (int i = tid; i < nar; i += nt) ar_of[ar[i]] = -1;
        __syncthreads();
        phase(8);
    }
}

}  // namespace

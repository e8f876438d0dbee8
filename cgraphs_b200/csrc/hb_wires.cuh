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
// 32 * WIRE_CHUNK_WORDS so that the bit sets stay small. One CTA works on one frame; its bit sets live in a
// per-CTA scratch area in global memory (L2 resident).
//
// Results go into the bond table with byte-wide columns: row word (frame >> 2), byte (frame & 3) = k (waters of the
// shortest wire, 1..127, 0 = none) | 0x80 if a direct bond exists; the last two words of a row hold
// ~min((frame << 32) | (acceptor entry << 1) | donor_is_second_node) over the direct bonds (0 = none), which is
// what the host needs to orient the key the way the reference's first discovery does (waterwire.py:277-279,299-301).
#pragma once
#include "hb_common.cuh"

namespace {

#define WIRE_CHUNK_WORDS 32
#define WIRE_THREADS 256

struct WireCfg {
    const int* ent_node;     // [n_ent] residue node (< n_res) of a selection entry, n_res + water index of a water entry
    int n_res, n_wat;
    int max_water, allow_direct;
    int nb4;                 // row words that hold the per-frame bytes (the "first direct" tail follows)
    int cap_words;           // min(ceil(n_res / 32), WIRE_CHUNK_WORDS): words per node the scratch is sized for
    // per-CTA scratch
    int* wl_of;              // [n_cta][n_wat]   water -> local index, -1 when idle
    int* lw_list;            // [n_cta][n_wat]
    int* ar_of;              // [n_cta][n_res]   residue node -> active index, -1 when idle
    int* ar_list;            // [n_cta][n_res]
    unsigned* wbits;         // [n_cta][3][n_wat * cap_words]   seen / frontier / next
    unsigned* tbits;         // [n_cta][2][n_res * cap_words]   found / level accumulator
    const int* status;       // status block: [2] fused_overflow, [3] pair_overflow, [6] wire_overflow
};

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
    if (wc.status[2] | wc.status[3] | wc.status[6]) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n_res = wc.n_res;
    int* wl_of = wc.wl_of + (size_t)blockIdx.x * wc.n_wat;
    int* lw = wc.lw_list + (size_t)blockIdx.x * wc.n_wat;
    int* ar_of = wc.ar_of + (size_t)blockIdx.x * n_res;
    int* ar = wc.ar_list + (size_t)blockIdx.x * n_res;
    const size_t wstride = (size_t)wc.n_wat * wc.cap_words, tstride = (size_t)n_res * wc.cap_words;
    unsigned* seen = wc.wbits + (size_t)blockIdx.x * 3 * wstride;
    unsigned* cur = seen + wstride;
    unsigned* nxt = cur + wstride;
    unsigned* found = wc.tbits + (size_t)blockIdx.x * 2 * tstride;
    unsigned* acc = found + tstride;
    __shared__ int s_nlw, s_nar, s_any;

    for (int b = blockIdx.x; b < n_frames; b += gridDim.x) {
        const int ne = min(edge_cnt[b], edge_cap);
        const int2* E = edges + (size_t)b * edge_cap;
        const long long frame = frame0 + b;
        if (tid == 0) { s_nlw = 0; s_nar = 0; }
        __syncthreads();
        // local waters and active residues (those with a residue-water edge) get dense indices
        for (int i = tid; i < ne; i += nt) {
            const int2 e = E[i];
            const int n0 = wc.ent_node[e.x], n1 = wc.ent_node[e.y];
            if (n0 >= n_res && atomicCAS(&wl_of[n0 - n_res], -1, -2) == -1) lw[atomicAdd(&s_nlw, 1)] = n0 - n_res;
            if (n1 >= n_res && atomicCAS(&wl_of[n1 - n_res], -1, -2) == -1) lw[atomicAdd(&s_nlw, 1)] = n1 - n_res;
            if ((n0 >= n_res) != (n1 >= n_res)) {
                const int r = n0 < n_res ? n0 : n1;
                if (atomicCAS(&ar_of[r], -1, -2) == -1) ar[atomicAdd(&s_nar, 1)] = r;
            }
        }
        __syncthreads();
        const int nlw = s_nlw, nar = s_nar;
        for (int i = tid; i < nlw; i += nt) wl_of[lw[i]] = i;
        for (int i = tid; i < nar; i += nt) ar_of[ar[i]] = i;
        __syncthreads();

        if (nar > 0 && wc.max_water >= 1) {
            const int words = min((nar + 31) >> 5, wc.cap_words);
            const int chunk = words * 32;
            for (int base = 0; base < nar; base += chunk) {
                const int nw = nlw * words, ntw = nar * words;
                for (int i = tid; i < nw; i += nt) cur[i] = 0;
                for (int i = tid; i < ntw; i += nt) found[i] = 0;
                __syncthreads();
                // depth 1: the waters bonded to each source of this chunk
                for (int i = tid; i < ne; i += nt) {
                    const int2 e = E[i];
                    const int n0 = wc.ent_node[e.x], n1 = wc.ent_node[e.y];
                    if ((n0 >= n_res) == (n1 >= n_res)) continue;
                    const int s = ar_of[n0 < n_res ? n0 : n1] - base;
                    if (s < 0 || s >= chunk) continue;
                    const int w = wl_of[(n0 < n_res ? n1 : n0) - n_res];
                    atomicOr(&cur[(size_t)w * words + (s >> 5)], 1u << (s & 31));
                }
                __syncthreads();
                for (int i = tid; i < nw; i += nt) seen[i] = cur[i];
                for (int k = 1; k <= wc.max_water; ++k) {
                    for (int i = tid; i < ntw; i += nt) acc[i] = 0;
                    __syncthreads();
                    // sources whose frontier (depth k) holds a water next to residue t
                    for (int i = tid; i < ne; i += nt) {
                        const int2 e = E[i];
                        const int n0 = wc.ent_node[e.x], n1 = wc.ent_node[e.y];
                        if ((n0 >= n_res) == (n1 >= n_res)) continue;
                        const int t = ar_of[n0 < n_res ? n0 : n1];
                        const int w = wl_of[(n0 < n_res ? n1 : n0) - n_res];
                        for (int j = 0; j < words; ++j) {
                            const unsigned v = cur[(size_t)w * words + j];
                            if (v) atomicOr(&acc[(size_t)t * words + j], v);
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < ntw; i += nt) {
                        unsigned nb = acc[i] & ~found[i];
                        if (!nb) continue;
                        found[i] |= nb;
                        const int t = i / words, j = i - t * words;
                        const int tn = ar[t];
                        while (nb) {
                            const int bit = __ffs(nb) - 1;
                            nb &= nb - 1;
                            const int sn = ar[base + j * 32 + bit];
                            if (sn < tn) wire_record(bt, sn, tn, frame, (unsigned)k);   // every pair from its smaller node
                        }
                    }
                    if (k == wc.max_water) break;
                    // one step along the water-water edges
                    for (int i = tid; i < nw; i += nt) nxt[i] = 0;
                    if (tid == 0) s_any = 0;
                    __syncthreads();
                    for (int i = tid; i < ne; i += nt) {
                        const int2 e = E[i];
                        const int n0 = wc.ent_node[e.x], n1 = wc.ent_node[e.y];
                        if (n0 < n_res || n1 < n_res) continue;
                        const int a = wl_of[n0 - n_res], c = wl_of[n1 - n_res];
                        for (int j = 0; j < words; ++j) {
                            const unsigned va = cur[(size_t)a * words + j], vc = cur[(size_t)c * words + j];
                            if (va) atomicOr(&nxt[(size_t)c * words + j], va);
                            if (vc) atomicOr(&nxt[(size_t)a * words + j], vc);
                        }
                    }
                    __syncthreads();
                    int any = 0;
                    for (int i = tid; i < nw; i += nt) {
                        const unsigned n = nxt[i] & ~seen[i];
                        seen[i] |= n;
                        cur[i] = n;
                        any |= n != 0;
                    }
                    if (any) s_any = 1;
                    __syncthreads();
                    if (!s_any) break;
                }
                __syncthreads();
            }
        }
        // direct bonds: (acceptor entry, donor entry)
        if (wc.allow_direct) {
            for (int i = tid; i < ne; i += nt) {
                const int2 e = E[i];
                const int na = wc.ent_node[e.x], nd = wc.ent_node[e.y];
                if (na >= n_res || nd >= n_res) continue;
                const int lo = min(na, nd), hi = max(na, nd);
                const int h = table_slot(bt, ((unsigned long long)(unsigned)lo << 32) | (unsigned)hi);
                if (h < 0) continue;
                unsigned* row = bt.bits + (size_t)h * bt.wpr;
                atomicOr(&row[frame >> 2], 0x80u << (8 * (int)(frame & 3)));
                const unsigned long long packed = ((unsigned long long)frame << 32) | ((unsigned long long)(unsigned)e.x << 1) |
                                                  (unsigned long long)(nd == hi ? 1 : 0);
                atomicMax((unsigned long long*)(row + wc.nb4), ~packed);
            }
        }
        __syncthreads();
        for (int i = tid; i < nlw; i += nt) wl_of[lw[i]] = -1;
        for (int i = tid; i < nar; i += nt) ar_of[ar[i]] = -1;
        __syncthreads();
    }
}

}  // namespace

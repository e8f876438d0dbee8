// hb_finalize.cuh - table maintenance and finalisation: compaction, LSD radix sort of the bond keys,
// row gather, per-bond popcounts, rehash on growth.
#pragma once
#include "hb_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// finalize kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_u64(unsigned long long* p, size_t n, unsigned long long v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__global__ void k_compact(BondTable bt, unsigned long long* out_keys, unsigned* out_slots, int* counter) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = i <= bt.mask && bt.keys[i] != EMPTY_KEY;
    unsigned m = __ballot_sync(0xffffffffu, live);
    if (!m) return;
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (live) {
        int pos = base + __popc(m & ((1u << lane) - 1u));
        out_keys[pos] = bt.keys[i];
        out_slots[pos] = i;
    }
}

// LSD radix sort, 8 bits per pass, of (key64, slot32) pairs. One block handles a tile; stability inside
// a tile comes from processing the tile in warp-sized strips in order.
#define RS_TILE 2048
__global__ void __launch_bounds__(256) k_rs_hist(const unsigned long long* keys, int n, int shift, unsigned* hist /*[256][nblk]*/) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE && base + i < n; i += 256)
        atomicAdd(&h[(unsigned)(keys[base + i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(1024) k_rs_scan(unsigned* hist, size_t n) {   // single block exclusive scan
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (size_t base = 0; base < n; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned v = i < n ? hist[i] : 0, s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
        if (lane == 31) warp_sum[w] = s;
        __syncthreads();
        if (w == 0) {
            unsigned t = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { unsigned u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
            warp_sum[lane] = t;
        }
        __syncthreads();
        unsigned off = carry + (w ? warp_sum[w - 1] : 0);
        if (i < n) hist[i] = off + s - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(32) k_rs_scatter(const unsigned long long* keys, const unsigned* vals, int n, int shift,
                                                   const unsigned* hist, unsigned long long* okeys, unsigned* ovals) {
    // one warp per tile: strips of 32 in order keep the sort stable
    __shared__ unsigned off[256];
    for (int d = threadIdx.x; d < 256; d += 32) off[d] = hist[(size_t)d * gridDim.x + blockIdx.x];
    __syncwarp();
    int base = blockIdx.x * RS_TILE;
    int lane = threadIdx.x;
    for (int s = 0; s < RS_TILE && base + s < n; s += 32) {
        int i = base + s + lane;
        bool live = i < n;
        unsigned long long k = live ? keys[i] : 0;
        unsigned d = (unsigned)(k >> shift) & 255u;
        unsigned peers = __match_any_sync(0xffffffffu, live ? d : 0xffffffffu);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        unsigned pos = 0;
        if (live) pos = off[d] + rank;
        __syncwarp();
        if (live && rank == __popc(peers) - 1) off[d] = pos + 1;   // last peer advances the digit cursor
        __syncwarp();
        if (live) { okeys[pos] = k; ovals[pos] = vals[i]; }
    }
}

// Small results (the common case: a few thousand distinct bonds): one CTA sorts (key, slot) pairs in shared
// memory with a bitonic network; n <= SORT_SMALL_MAX, padded with the largest key to a power of two.
#define SORT_SMALL_MAX 16384                 // 12 bytes of dynamic shared memory per element: 192 kB
#define SORT_SPECULATE_MAX 4096              // bond tables up to twice this size are finalised without knowing the count
__global__ void __launch_bounds__(1024) k_sort_small(const unsigned long long* keys_in, const unsigned* vals_in,
                                                     const int* n_ptr, unsigned long long* keys_out, unsigned* vals_out, int cap) {
    extern __shared__ __align__(16) unsigned long long sort_smem[];
    unsigned long long* k = sort_smem;                                       // [cap]
    unsigned* v = reinterpret_cast<unsigned*>(sort_smem + cap);            // [cap]
    const int n = min(*n_ptr, cap);                // (a larger count is caught by the host and redone with radix passes)
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        k[i] = i < n ? keys_in[i] : 0xFFFFFFFFFFFFFFFFull;
        v[i] = i < n ? vals_in[i] : 0u;
    }
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                int lo = 2 * i - (i & (stride - 1));          // index of the lower element of pair i
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) {
                    k[lo] = b; k[hi] = a;
                    unsigned t = v[lo]; v[lo] = v[hi]; v[hi] = t;
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { keys_out[i] = k[i]; vals_out[i] = v[i]; }
}

// Multi-GPU key exchange, send side: this rank's block of the all-gather = [count, key_0 .. key_{cap-1}].
__global__ void k_pack_block(const unsigned long long* keys, const int* n_ptr, int n_max, unsigned long long* block, int cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(*n_ptr, n_max);
    if (i == 0) block[0] = (unsigned long long)n;
    else if (i <= cap) block[i] = (i - 1 < n) ? keys[i - 1] : EMPTY_KEY;
}

// Merge step of the multi-GPU key exchange: `gathered` holds, per rank, [count, key_0 .. key_{cap-1}] as the
// all-gather delivered it (keys sorted, padding arbitrary). One CTA loads the valid keys into shared memory,
// sorts them (bitonic), drops duplicates and writes the global sorted distinct list; info[0] = its length,
// info[1] = 1 if some rank had more than `cap` keys (the caller then repeats with a larger capacity).
__global__ void __launch_bounds__(1024) k_merge_keys(const unsigned long long* gathered, int world, int cap, int m,
                                                     int sorted_runs, unsigned long long* out, long long* info) {
    extern __shared__ __align__(16) unsigned long long mk[];
    __shared__ int part[1024];
    __shared__ int over;
    if (threadIdx.x == 0) over = 0;
    __syncthreads();
    // every rank's list is sorted: store odd lists reversed, so that the runs of length `cap` alternate
    // ascending / descending and only the merge phases (size > cap) of the bitonic network are left to do
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int w = i / cap, j = i - w * cap;
        unsigned long long key = EMPTY_KEY;
        if (w < world) {
            const long long cnt = (long long)gathered[(size_t)w * (cap + 1)];
            if (j == 0 && cnt > cap) over = 1;
            if (j < cnt) key = gathered[(size_t)w * (cap + 1) + 1 + j];
        }
        mk[(w & 1) ? (w * cap + (cap - 1 - j)) : i] = key;
    }
    __syncthreads();
    for (int size = sorted_runs ? 2 * cap : 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = mk[lo], b = mk[hi];
                if ((a > b) == up) { mk[lo] = b; mk[hi] = a; }
            }
            __syncthreads();
        }
    // distinct keys: each thread owns a contiguous slice
    const int per = (m + blockDim.x - 1) / blockDim.x;
    const int s0 = min((int)threadIdx.x * per, m), s1 = min(s0 + per, m);
    int cnt = 0;
    for (int i = s0; i < s1; ++i) cnt += (mk[i] != EMPTY_KEY && (i == 0 || mk[i] != mk[i - 1])) ? 1 : 0;
    part[threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < (int)blockDim.x; ++t) { int c = part[t]; part[t] = run; run += c; }
        info[0] = run;
        info[1] = over;
    }
    __syncthreads();
    int pos = part[threadIdx.x];
    for (int i = s0; i < s1; ++i)
        if (mk[i] != EMPTY_KEY && (i == 0 || mk[i] != mk[i - 1])) out[pos++] = mk[i];
}

// Merge step for many ranks: the distinct keys are far fewer than the gathered ones (most bonds exist on every
// rank), so they are first de-duplicated in a small hash set by all SMs (k_union_insert) and only the distinct
// keys are sorted by one CTA (k_union_finish). Cost follows the number of distinct keys, not world * cap.
#define UNION_SLOTS 32768
__global__ void k_union_insert(const unsigned long long* gathered, int world, int cap, unsigned long long* table, long long* info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= world * cap) return;
    const int w = i / cap, j = i - w * cap;
    const long long cnt = (long long)gathered[(size_t)w * (cap + 1)];
    if (j == 0 && cnt > cap) info[1] = 1;
    if (j >= cnt) return;
    const unsigned long long key = gathered[(size_t)w * (cap + 1) + 1 + j];
    unsigned h = hash64(key) & (UNION_SLOTS - 1);
    for (int probe = 0; probe < UNION_SLOTS; ++probe) {
        const unsigned long long cur = atomicCAS(&table[h], EMPTY_KEY, key);
        if (cur == EMPTY_KEY || cur == key) return;
        h = (h + 1) & (UNION_SLOTS - 1);
    }
}
__global__ void __launch_bounds__(1024) k_union_finish(const unsigned long long* table, unsigned long long* out, long long* info) {
    extern __shared__ __align__(16) unsigned long long mk[];      // up to 16384 keys
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < UNION_SLOTS; i += blockDim.x) {  // coalesced scan; order is irrelevant, a sort follows
        const unsigned long long k = table[i];
        if (k != EMPTY_KEY) mk[atomicAdd(&total, 1)] = k;
    }
    __syncthreads();
    const int n = total;
    int m = 2;
    while (m < n) m <<= 1;
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) mk[i] = EMPTY_KEY;
    __syncthreads();
    for (int size = 2; size <= m; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (m >> 1); i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = mk[lo], b = mk[hi];
                if ((a > b) == up) { mk[lo] = b; mk[hi] = a; }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = mk[i];
    if (threadIdx.x == 0) info[0] = n;
}

__global__ void k_gather_rows(const unsigned* bits, const unsigned* slots, const int* n_ptr, int n_max, int wpr, unsigned* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)min(*n_ptr, n_max) * wpr;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        size_t row = i / wpr, col = i % wpr;
        out[i] = bits[(size_t)slots[row] * wpr + col];
    }
}

__global__ void k_row_counts(const unsigned* rows, int n, int wpr, long long* counts) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    int lane = threadIdx.x & 31;
    long long c = 0;
    for (int w = lane; w < wpr; w += 32) c += __popc(rows[(size_t)row * wpr + w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) counts[row] = c;
}

// AND of the selected occupancy rows, word by word (NetworkAnalysis.compute_joint_occupancy, network.py:369-376);
// one thread per 32-frame word, rows walked in order. select == nullptr: all rows.
__global__ void k_joint_rows(const unsigned* rows, const unsigned char* select, int n, int wpr, unsigned tail_mask,
                             unsigned* out, unsigned long long* n_set) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0u;
    if (w < wpr) {
        acc = 0xffffffffu;
        for (int r = 0; r < n; ++r)
            if (!select || select[r]) acc &= rows[(size_t)r * wpr + w];
        if (w == wpr - 1) acc &= tail_mask;
        out[w] = acc;
    }
    unsigned c = __popc(acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_set, (unsigned long long)c);
}

// which bonds exist in one frame (NetworkAnalysis._compute_graph_in_frame, network.py:1186-1198)
__global__ void k_frame_column(const unsigned* rows, int n, int wpr, long long frame, unsigned char* present) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) present[r] = (rows[(size_t)r * wpr + (frame >> 5)] >> (frame & 31)) & 1u;
}

// re-insert the rows of an old table into a larger one
__global__ void k_rehash(BondTable oldt, BondTable newt) {
    unsigned slot = blockIdx.x;
    if (slot > oldt.mask) return;
    unsigned long long key = oldt.keys[slot];
    if (key == EMPTY_KEY) return;
    __shared__ unsigned dst;
    if (threadIdx.x == 0) {
        unsigned h = hash64(key) & newt.mask;
        for (;;) {
            unsigned long long cur = atomicCAS(&newt.keys[h], EMPTY_KEY, key);
            if (cur == EMPTY_KEY || cur == key) break;
            h = (h + 1) & newt.mask;
        }
        dst = h;
    }
    __syncthreads();
    for (int w = threadIdx.x; w < oldt.wpr; w += blockDim.x)
        newt.bits[(size_t)dst * newt.wpr + w] = oldt.bits[(size_t)slot * oldt.wpr + w];
}


}  // namespace

// libhbgpu - hand-written sm_100a kernels + C ABI for the cgraphs / Bridge H-bond hot path.
//
// Replaces the per-frame body of HbondAnalysis.set_hbonds_in_selection (reference
// cgraphs/mdhbond/hbond.py:96-135) and set_hbonds_in_selection_and_water_around (hbond.py:232-289),
// including check_angle (helpfunctions.py:137-159) and the result[key][frame] = True bookkeeping.
//
// Pipeline per batch of frames (all frames of a batch are processed by every kernel launch):
//   K0 init_frames        reset per-frame bounding boxes / counters
//   K1 gather_sel         selection points -> float4 + per-frame bounding box
//   K2 grid_setup         per-frame cell-grid geometry (non-periodic bounding-box grid)
//   K3 count / K4 scan / K5 scatter   counting sort of points into cells (cell list)
//   K6 local_water        waters within around_radius of any selection entry (hbond.py:249-252)
//   K7 pairs              half-shell neighbour search, exact DIST and ANG predicates, key build,
//                         insert into the bond hash table, set the bit (bond, frame)
// Finalize: compact the table, radix-sort the keys, gather occupancy rows in key order.
//
// Exactness:
//   DIST: float32 coordinates widened to double, s = fl(fl(dx*dx + dy*dy) + dz*dz) <= fl(r*r), inclusive.
//   ANG : float32 chain with individually rounded operations in the reference's order, compared as
//         cos_threshold <= c <= 1 (c < -1 clamps to -1), cos_threshold calibrated on the host.
// No fast-math, no FMA contraction on these paths (explicit _rn intrinsics).
#include "../../include/hbgpu.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define HB_VERSION "hbgpu 0.1 (sm_100a)"

#include "hb_general.cuh"
#include "hb_fused.cuh"
#include "hb_framepairs.cuh"
#include "hb_finalize.cuh"
#include "hb_wires.cuh"
#include "hb_xtc.cuh"
#include "hb_xtc_host.h"

namespace {

enum KernelClass { KC_INIT = 0, KC_GATHER, KC_GRID, KC_COUNT, KC_SCAN, KC_SCATTER, KC_LOCAL, KC_PAIRS, KC_FUSED, KC_WIRES, KC_XTC, KC_N };
const char* const kKernelNames[KC_N] = {"init_frames", "gather_sel", "grid_setup", "cell_count",
                                        "cell_scan", "cell_scatter", "local_water", "pairs", "frame_fused", "wire_bfs", "xtc_decode"};

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct hb_ctx {
    int device = 0;
    std::string err;
    cudaStream_t s_compute = nullptr, s_copy = nullptr;
    cudaStream_t s_own = nullptr;     // the library's own compute stream (s_compute may be a caller's stream)
    cudaDeviceProp prop{};

    // topology
    bool have_topo = false;
    DevTopo tp{};
    std::vector<void*> topo_allocs;
    std::vector<long long> h_sys_atom_off;
    std::vector<int> h_sys_sel_off, h_sys_wat_off;
    int n_ent = 0;
    int* d_ent_group = nullptr;
    int max_sel = 0, max_wat = 0;
    long long max_atoms = 0;

    // params
    bool have_params = false;
    hb_params prm{};
    RunParams rp{};

    // options
    long long opt_batch_bytes = 4ll << 30;
    long long opt_table_capacity = 0;   // 0: sized from the selection
    long long opt_max_batch_frames = 2048;
    int opt_profile = 0;
    int opt_force_radix = 0;          // 1: always use the multi-pass radix sort in hb_finalize (test hook)
    int opt_fused = 1;                // 1: use the one-CTA-per-frame kernel when a frame fits in shared memory
    long long opt_fused_lw_cap = 0;   // 0: automatic
    long long opt_fused_cell_cap = 0; // 0: automatic
    long long opt_fused_pair_cap = 0; // 0: automatic
    long long opt_fused_stages = 8;   // stages of the water-tile ring = warps that stream (one stage each)
    long long opt_fused_wait_ns = 100;
    long long opt_fused_prefetch = 2; // tiles prefetched into L2 ahead of the ring (more evicts lines before use)
    int opt_fused_sel_prefetch = 1;   // bulk L2 prefetch of the selection's atom block at the start of a frame
    long long opt_fused_max_streams = 0; // frames of one SM that may stream at the same time (0 = no limit)
    int* d_sm_streams = nullptr;
    long long occ_key = -1;           // resident fused CTAs per SM, cached per shared-memory plan
    int occ_per_sm = 0;
    long long opt_fused_stagger = 125000; // cycles over which the first wave of fused CTAs spreads its start
    int opt_fused_debug = 0;          // measurement hook (results are wrong when set): see FusedCfg::debug
    int fused_grow = 0;
    int sort_smem_cap = 0;            // dynamic shared memory k_sort_small may use was raised (per context / device)
    int fused_grow_hint = 0;          // capacities that an earlier run on this topology needed ...
    double fused_hint_around = 0.0;   // ... with this around radius (narrower shells start from the defaults)

    // fused-path plan (valid for the current run)
    bool water_regular = false;       // water point i of every system is atom first + i * stride
    int water_stride = 0;
    bool presence = false;            // HB_METHOD_WATER_PRESENCE
    // general path: pair search as one CTA per frame (hb_framepairs.cuh) while the frames fit
    int opt_frame_pairs = 1;
    int fp_level = 0;                 // 0: use it, 1: some frame did not fit - multi-kernel pair search
    FramePairsCfg fp{};
    bool fused_active = false;
    double fused_unfit_around = -1.0; // around radius of a run whose frames did not fit the fused kernel's shared memory
    FusedCfg fc{};                    // (same topology): later runs with a shell at least as wide start on the general path
    long long fused_fallbacks = 0;

    // run state
    bool running = false, finalized = false;
    long long n_frames_total = 0, next_frame = 0;
    BondTable bt{};
    long long table_cap = 0;
    long long alloc_sig[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int table_grows = 0;
    int* h_status = nullptr;          // pinned, 4 ints per slot: n_keys, overflow, fused_overflow, pad
    int last_status[6] = {0, 0, 0, 0, 0, 0};   // status block as of the last full_verify
    int* h_verify = nullptr;          // pinned landing zone of that read back
    unsigned long long* h_hits = nullptr;

    // batch scratch
    int batch_frames = 0;             // frames the general-path scratch holds
    int host_batch_frames = 0;        // frames one input buffer of hb_submit_frames holds
    long long opt_fused_batch_frames = 1 << 15;   // frames per launch of the fused kernel (device-resident input)
    int cell_cap = 0;
    BatchView bv{};
    std::vector<void*> scratch_allocs;
    PairBuf pairbuf{};                // batch-wide pair buffer of the general path (lazy)
    long long opt_pair_buf_bytes = 512ll << 20;
    bool use_pairbuf = true;
    float* d_in[2] = {nullptr, nullptr};
    float* h_stage[2] = {nullptr, nullptr};
    size_t in_bytes = 0;              // planned capacity of one input buffer
    size_t in_alloc = 0;              // bytes actually allocated per input buffer
    size_t stage_alloc = 0;           // bytes per pinned staging buffer
    unsigned char* d_raw[2] = {nullptr, nullptr};   // planar sources (DCD): raw frame records before the interleave
    size_t raw_alloc = 0;
    long long* d_xoff[2] = {nullptr, nullptr};      // XTC sources: record offsets of the frames in d_raw[k]
    long long* h_xoff[2] = {nullptr, nullptr};      // (pinned)
    size_t xoff_cap = 0;                            // frames
    int* d_xtc_status = nullptr;                    // set by k_xtc_decode when a record is malformed
    int* h_xtc_status = nullptr;                    // pinned copy, read back behind every decode
    cudaEvent_t ev_copied[2]{}, ev_done[2]{};
    cudaEvent_t ev_run0{}, ev_run1{};
    long long batch_counter = 0;

    struct Slot { bool busy; const float* d_xyz; long long f0; int nf; long long apf; size_t bytes; const float* d_box; bool fused; };
    Slot slots[2] = {{false, nullptr, 0, 0, 0, 0, nullptr, false}, {false, nullptr, 0, 0, 0, 0, nullptr, false}};
    bool last_batch_fused = false;
    float* d_box[2] = {nullptr, nullptr};      // lattice vectors of the frames in each input buffer
    float* h_box[2] = {nullptr, nullptr};      // pinned staging for them
    size_t box_cap = 0;                        // frames

    // results
    long long n_bonds = 0;
    unsigned long long* d_out_keys = nullptr;     // == fin_k[fin_cur] after hb_finalize
    unsigned* d_out_rows = nullptr;
    size_t out_rows_cap = 0;                      // bytes
    // finalize scratch, kept across runs
    unsigned long long* fin_k[2] = {nullptr, nullptr};
    unsigned* fin_v[2] = {nullptr, nullptr};
    unsigned* fin_hist = nullptr;
    int* fin_counter = nullptr;
    size_t fin_cap = 0, fin_hist_cap = 0;
    unsigned long long* fin_sorted_k = nullptr;   // sorted keys of the queued / last finalize
    long long fin_n_max = 0;
    bool fin_queued = false, fin_speculative = false;
    long long n_sel_total = 0, n_wat_total = 0;
    IonTable ions{nullptr, nullptr, 0};
    unsigned long long* merge_table = nullptr;    // hash set of hb_merge_keys_device (many ranks)
    void* nccl_comm = nullptr;                    // hb_comm_init: communicator of the key exchange (ncclComm_t)
    int comm_world = 0, comm_rank = 0;
    unsigned long long* comm_send = nullptr;      // [cap + 1] this rank's block
    unsigned long long* comm_recv = nullptr;      // [world][cap + 1]
    int comm_cap = 0;
    // water wires (hb_wires.cuh)
    struct Wire {
        bool on = false;
        int* d_ent_node = nullptr;                // owned here (hb_set_wires may be called before every run)
        int* d_direct_node = nullptr;             // hb_set_wire_direct_nodes (null: direct bonds are booked on ent_node)
        int n_res = 0, n_wat_nodes = 1, max_water = 0, allow_direct = 1;
        int nb4 = 0;                              // row words holding the per-frame bytes
        int2* d_edges = nullptr;
        int* d_cnt = nullptr;
        EdgeSink* d_sink = nullptr;
        long long edge_cap = 0;
        int frames = 0;                           // frames per search + BFS round
        int n_cta = 0;
        std::vector<void*> allocs;
        WireCfg cfg{};
        // convex-hull option: per-frame (position water, name water) lists from the host (hb_set_wire_virtual_waters)
        bool virt_on = false;
        VirtualWaters virt{nullptr, nullptr, nullptr, 0};
        std::vector<void*> virt_allocs;
    } wire;
    bool lw_attr_set[2] = {false, false};
    int opt_lw_smem = 1;                          // general path: stage the selection grid in shared memory when it fits
    long long opt_wire_edge_cap = 0;              // 0: automatic
    long long opt_wire_smem_words = WIRE_SMEM_WORDS;
    long long opt_wire_ctas_per_sm = 1;           // the BFS kernel takes a whole SM's shared memory
    BondTable bt_run{};                           // what the search kernels of the current launch get (wire mode: + edge lists)
    int ion_include_water = 1;
    // trajectory mode: atom ranges [begin, end) that hold every atom the kernels read; the rest of a frame
    // (inert atoms) is never copied to the device. Empty = copy whole frames.
    std::vector<std::pair<long long, long long>> copy_ranges;
    long long sel_block_lo = 0, sel_block_n = 0;   // atoms [lo, lo + n) hold the selection points and their hydrogens (0 = unknown)
    std::vector<unsigned char> atom_used;     // single-system topologies: 1 = some kernel reads this atom's coordinates
    int opt_copy_ranges = 1;
    int opt_stage_threads = 8;        // host threads that copy pageable frames into the pinned staging buffer
    int trace = 0;

    // timers
    hb_timers tm{};
    struct PendingEvent { cudaEvent_t a, b; int kc; };
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> event_pool;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return HB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

// HBGPU_TRACE=1: wall-clock trace of the host-side phases on stderr
struct Trace {
    hb_ctx* ctx; const char* what; std::chrono::steady_clock::time_point t0;
    Trace(hb_ctx* c, const char* w) : ctx(c), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* step) {
        if (!ctx->trace) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[hbgpu] %s/%s %.3f ms\n", what, step, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static int fail(hb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

template <typename T>
static int upload(hb_ctx* ctx, const T* src, size_t n, const T** dst, std::vector<void*>& owner) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    owner.push_back(p);
    if (n) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)p;
    return HB_OK;
}

static void free_list(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

static cudaEvent_t get_event(hb_ctx* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

static void drain_events(hb_ctx* ctx) {
    for (auto& pe : ctx->pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(pe.b) == cudaSuccess && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            if (pe.kc == -1) ctx->tm.ms_total += ms;
            else if (pe.kc == -2) ctx->tm.ms_h2d += ms;
            else ctx->tm.ms_kernel[pe.kc] += ms;
        }
        ctx->event_pool.push_back(pe.a);
        ctx->event_pool.push_back(pe.b);
    }
    ctx->pending.clear();
}

struct KTimer {   // times one kernel class when profiling is on
    hb_ctx* ctx; int kc; cudaEvent_t a{};
    KTimer(hb_ctx* c, int k) : ctx(c), kc(k) {
        ctx->tm.launches[kc]++;
        ctx->tm.n_launches++;
        if (ctx->opt_profile) { a = get_event(ctx); cudaEventRecord(a, ctx->s_compute); }
    }
    ~KTimer() {
        if (ctx->opt_profile) {
            cudaEvent_t b = get_event(ctx);
            cudaEventRecord(b, ctx->s_compute);
            ctx->pending.push_back({a, b, kc});
        }
    }
};

// Staging copy pageable -> pinned with several threads: one thread moves ~10 GB/s, PCIe takes 55 GB/s.
// copy(f0, f1) copies the frames [f0, f1) of the batch.
template <typename F>
static void parallel_frames(int nf, size_t bytes_per_frame, int max_threads, F copy) {
    int nt = (int)std::min<size_t>((size_t)std::max(max_threads, 1), (bytes_per_frame * (size_t)nf) >> 23);   // >= 8 MB per thread
    nt = std::min(nt, nf);
    if (nt <= 1) { copy(0, nf); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        const int a = (int)((long long)nf * t / nt), b = (int)((long long)nf * (t + 1) / nt);
        th.emplace_back([=]() { copy(a, b); });
    }
    for (auto& x : th) x.join();
}

// ---- NCCL, loaded at run time (the library links against nothing but the CUDA runtime) ---------------------------
struct HbNcclId { char internal[128]; };
struct HbNccl {
    void* handle = nullptr;
    int (*GetUniqueId)(HbNcclId*) = nullptr;
    int (*CommInitRank)(void**, int, HbNcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static HbNccl& hb_nccl() {
    static HbNccl n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("HBGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (n.handle) {
            n.GetUniqueId = (int (*)(HbNcclId*))dlsym(n.handle, "ncclGetUniqueId");
            n.CommInitRank = (int (*)(void**, int, HbNcclId, int))dlsym(n.handle, "ncclCommInitRank");
            n.CommDestroy = (int (*)(void*))dlsym(n.handle, "ncclCommDestroy");
            n.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(n.handle, "ncclAllGather");
            n.GetErrorString = (const char* (*)(int))dlsym(n.handle, "ncclGetErrorString");
            n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather;
        }
    }
    return n;
}

// Trajectory mode: atom ranges [begin, end) that cover every atom marked in ctx->atom_used; holes shorter than 4096
// atoms (48 kB) are copied along. Left empty (= copy whole frames) when the ranges would not save at least 10 %.
static void build_copy_ranges(hb_ctx* ctx) {
    ctx->copy_ranges.clear();
    const std::vector<unsigned char>& used = ctx->atom_used;
    const long long n = (long long)used.size(), gap = 4096;
    long long covered = 0, a = 0;
    while (a < n) {
        while (a < n && !used[a]) ++a;
        if (a >= n) break;
        long long b = a, last = a;
        while (b < n && b - last <= gap) { if (used[b]) last = b; ++b; }
        ctx->copy_ranges.push_back({a, last + 1});
        covered += last + 1 - a;
        a = last + 1;
    }
    if (ctx->copy_ranges.size() > 8 || covered * 10 > n * 9) ctx->copy_ranges.clear();   // not worth it
}

extern "C" {

const char* hb_version(void) { return HB_VERSION; }

const char* hb_kernel_name(int k) { return (k >= 0 && k < KC_N) ? kKernelNames[k] : nullptr; }

const char* hb_last_error(hb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hb_create(hb_ctx** out, int device) {
    if (!out) return HB_ERR_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return HB_ERR_NODEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "device ordinal out of range"; return HB_ERR_ARG; }
    hb_ctx* ctx = new (std::nothrow) hb_ctx();
    if (!ctx) return HB_ERR_NOMEM;
    ctx->device = device;
    ctx->trace = getenv("HBGPU_TRACE") != nullptr;
    ctx->opt_stage_threads = (int)std::min<unsigned>(16u, std::max<unsigned>(1u, std::thread::hardware_concurrency()));
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return HB_ERR_CUDA;
    }
    if (ctx->prop.major != 10) {
        g_create_error = "libhbgpu is built for sm_100a (B200) only; found sm_" + std::to_string(ctx->prop.major) +
                         std::to_string(ctx->prop.minor);
        delete ctx;
        return HB_ERR_NODEVICE;
    }
    if ((e = xtc_upload_tables()) != cudaSuccess) {
        g_create_error = std::string("constant tables: ") + cudaGetErrorString(e);
        delete ctx;
        return HB_ERR_CUDA;
    }
    cudaStreamCreateWithFlags(&ctx->s_own, cudaStreamNonBlocking);
    ctx->s_compute = ctx->s_own;
    cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking);
    for (int k = 0; k < 2; ++k) {
        cudaEventCreateWithFlags(&ctx->ev_copied[k], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming);
    }
    cudaEventCreate(&ctx->ev_run0);
    cudaEventCreate(&ctx->ev_run1);
    cudaHostAlloc((void**)&ctx->h_status, 16 * sizeof(int), cudaHostAllocDefault);
    cudaHostAlloc((void**)&ctx->h_verify, 8 * sizeof(int), cudaHostAllocDefault);
    cudaHostAlloc((void**)&ctx->h_hits, sizeof(unsigned long long), cudaHostAllocDefault);
    *out = ctx;
    return HB_OK;
}

static void free_run(hb_ctx* ctx) {
    free_list(ctx->scratch_allocs);
    ctx->pairbuf = PairBuf{};         // lived in scratch_allocs
    for (int k = 0; k < 2; ++k) {
        if (ctx->d_in[k]) cudaFree(ctx->d_in[k]);
        if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]);
        if (ctx->d_raw[k]) cudaFree(ctx->d_raw[k]);
        ctx->d_in[k] = nullptr;
        ctx->h_stage[k] = nullptr;
        ctx->d_raw[k] = nullptr;
    }
    ctx->in_alloc = ctx->stage_alloc = ctx->raw_alloc = 0;
    for (int k = 0; k < 2; ++k) {
        if (ctx->d_box[k]) cudaFree(ctx->d_box[k]);
        if (ctx->h_box[k]) cudaFreeHost(ctx->h_box[k]);
        ctx->d_box[k] = nullptr;
        ctx->h_box[k] = nullptr;
    }
    ctx->box_cap = 0;
    free_list(ctx->wire.allocs);
    ctx->wire.d_edges = nullptr;
    ctx->wire.d_cnt = nullptr;
    if (ctx->d_sm_streams) cudaFree(ctx->d_sm_streams);
    ctx->d_sm_streams = nullptr;
    if (ctx->bt.keys) cudaFree(ctx->bt.keys);
    if (ctx->bt.bits) cudaFree(ctx->bt.bits);
    if (ctx->bt.n_keys) cudaFree(ctx->bt.n_keys);
    ctx->bt = BondTable{};
    if (ctx->d_out_rows) cudaFree(ctx->d_out_rows);
    for (int k = 0; k < 2; ++k) {
        if (ctx->fin_k[k]) cudaFree(ctx->fin_k[k]);
        if (ctx->fin_v[k]) cudaFree(ctx->fin_v[k]);
        ctx->fin_k[k] = nullptr;
        ctx->fin_v[k] = nullptr;
    }
    if (ctx->fin_hist) cudaFree(ctx->fin_hist);
    if (ctx->fin_counter) cudaFree(ctx->fin_counter);
    ctx->fin_hist = nullptr;
    ctx->fin_counter = nullptr;
    ctx->fin_cap = ctx->fin_hist_cap = ctx->out_rows_cap = 0;
    ctx->d_out_keys = nullptr;
    ctx->d_out_rows = nullptr;
    ctx->running = ctx->finalized = false;
    for (auto& sl : ctx->slots) sl.busy = false;
}

int hb_destroy(hb_ctx* ctx) {
    if (!ctx) return HB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    drain_events(ctx);
    free_run(ctx);
    free_list(ctx->topo_allocs);
    if (ctx->wire.d_ent_node) cudaFree(ctx->wire.d_ent_node);
    if (ctx->wire.d_direct_node) cudaFree(ctx->wire.d_direct_node);
    free_list(ctx->wire.virt_allocs);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(ctx->ev_copied[k]); cudaEventDestroy(ctx->ev_done[k]); }
    cudaEventDestroy(ctx->ev_run0);
    cudaEventDestroy(ctx->ev_run1);
    if (ctx->merge_table) cudaFree(ctx->merge_table);
    if (ctx->nccl_comm) { hb_nccl().CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (ctx->comm_send) cudaFree(ctx->comm_send);
    if (ctx->comm_recv) cudaFree(ctx->comm_recv);
    for (int k = 0; k < 2; ++k) { if (ctx->d_xoff[k]) cudaFree(ctx->d_xoff[k]); if (ctx->h_xoff[k]) cudaFreeHost(ctx->h_xoff[k]); }
    if (ctx->d_xtc_status) cudaFree(ctx->d_xtc_status);
    if (ctx->h_xtc_status) cudaFreeHost(ctx->h_xtc_status);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_verify) cudaFreeHost(ctx->h_verify);
    if (ctx->h_hits) cudaFreeHost(ctx->h_hits);
    cudaStreamDestroy(ctx->s_own);
    cudaStreamDestroy(ctx->s_copy);
    delete ctx;
    return HB_OK;
}

int hb_set_option(hb_ctx* ctx, const char* name, int64_t value) {
    if (ctx && name && !strcmp(name, "frame_pairs")) { ctx->opt_frame_pairs = value != 0; return HB_OK; }
    if (ctx && name && !strcmp(name, "lw_smem")) { ctx->opt_lw_smem = value != 0; return HB_OK; }
    if (ctx && name && !strcmp(name, "wire_edge_cap")) { ctx->opt_wire_edge_cap = value; return HB_OK; }
    if (ctx && name && !strcmp(name, "wire_smem_words")) { ctx->opt_wire_smem_words = value; return HB_OK; }
    if (!ctx || !name) return HB_ERR_ARG;
    std::string n(name);
    if (n == "batch_bytes") ctx->opt_batch_bytes = std::max<long long>(value, 1 << 20);
    else if (n == "table_capacity") ctx->opt_table_capacity = std::max<long long>(value, 1024);
    else if (n == "profile") ctx->opt_profile = value != 0;
    else if (n == "max_batch_frames") ctx->opt_max_batch_frames = std::max<long long>(value, 1);
    else if (n == "fused") ctx->opt_fused = value != 0;
    else if (n == "force_radix_sort") ctx->opt_force_radix = value != 0;
    else if (n == "pair_buf_bytes") { ctx->opt_pair_buf_bytes = std::max<long long>(value, 0); ctx->use_pairbuf = value > 0; }
    else if (n == "compute_stream") {   // run all kernels on the caller's CUDA stream (0 = back to the library's own)
        if (ctx->running) return fail(ctx, HB_ERR_STATE, "compute_stream cannot change during a run");
        cudaStreamSynchronize(ctx->s_compute);
        ctx->s_compute = value ? (cudaStream_t)(uintptr_t)value : ctx->s_own;
    }
    else if (n == "copy_ranges") ctx->opt_copy_ranges = value != 0;
    else if (n == "stage_threads") ctx->opt_stage_threads = (int)std::min<long long>(std::max<long long>(value, 1), 64);
    else if (n == "fused_lw_cap") ctx->opt_fused_lw_cap = std::max<long long>(value, 0);
    else if (n == "fused_cell_cap") ctx->opt_fused_cell_cap = std::max<long long>(value, 0);
    else if (n == "fused_pair_cap") ctx->opt_fused_pair_cap = std::max<long long>(value, 0);
    else if (n == "fused_stages") ctx->opt_fused_stages = std::max<long long>(value, 2);
    else if (n == "fused_batch_frames") ctx->opt_fused_batch_frames = std::max<long long>(value, 1);
    else if (n == "fused_prefetch") ctx->opt_fused_prefetch = std::max<long long>(value, 0);
    else if (n == "fused_debug") ctx->opt_fused_debug = (int)value;
    else if (n == "fused_sel_prefetch") ctx->opt_fused_sel_prefetch = value != 0;
    else if (n == "fused_max_streams") ctx->opt_fused_max_streams = std::min<long long>(std::max<long long>(value, 0), 64);
    else if (n == "fused_stagger") ctx->opt_fused_stagger = std::min<long long>(std::max<long long>(value, 0), 4000000);
    else if (n == "fused_wait_ns") ctx->opt_fused_wait_ns = std::max<long long>(value, 0);
    else return fail(ctx, HB_ERR_ARG, "unknown option " + n);
    return HB_OK;
}

int hb_set_topology(hb_ctx* ctx, const hb_topology* t) {
    if (!ctx || !t) return HB_ERR_ARG;
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_topology during a run");
    if (t->n_systems < 1 || !t->sys_atom_off || !t->sys_sel_off || !t->sys_wat_off)
        return fail(ctx, HB_ERR_ARG, "topology: system offset tables missing");
    if (t->n_sel < 0 || t->n_wat < 0 || t->n_ent < 0 || t->n_ent_h < 0) return fail(ctx, HB_ERR_ARG, "topology: negative count");
    if (t->n_sel >= WATER_TAG || t->n_wat >= WATER_TAG) return fail(ctx, HB_ERR_ARG, "topology: too many points");
    CK(cudaSetDevice(ctx->device));
    free_list(ctx->topo_allocs);
    ctx->ions = IonTable{nullptr, nullptr, 0};
    ctx->wire.on = false;
    if (ctx->wire.d_ent_node) cudaFree(ctx->wire.d_ent_node);
    ctx->wire.d_ent_node = nullptr;
    if (ctx->wire.d_direct_node) cudaFree(ctx->wire.d_direct_node);
    ctx->wire.d_direct_node = nullptr;
    free_list(ctx->wire.virt_allocs);
    ctx->wire.virt_on = false;
    ctx->fused_unfit_around = -1.0;
    ctx->fused_grow_hint = 0;
    ctx->fp_level = 0;
    ctx->have_topo = false;       // (run allocations are kept: hb_begin compares everything that sizes them)
    int ns = t->n_systems;
    ctx->h_sys_atom_off.assign(t->sys_atom_off, t->sys_atom_off + ns + 1);
    ctx->h_sys_sel_off.assign(t->sys_sel_off, t->sys_sel_off + ns + 1);
    ctx->h_sys_wat_off.assign(t->sys_wat_off, t->sys_wat_off + ns + 1);
    ctx->max_sel = ctx->max_wat = 0;
    ctx->max_atoms = 0;
    ctx->n_sel_total = t->n_sel;
    ctx->n_wat_total = t->n_wat;
    for (int s = 0; s < ns; ++s) {
        int a = t->sys_sel_off[s + 1] - t->sys_sel_off[s], w = t->sys_wat_off[s + 1] - t->sys_wat_off[s];
        long long at = t->sys_atom_off[s + 1] - t->sys_atom_off[s];
        if (a < 0 || w < 0 || at < 0) return fail(ctx, HB_ERR_ARG, "topology: offsets not monotone");
        ctx->max_sel = std::max(ctx->max_sel, a);
        ctx->max_wat = std::max(ctx->max_wat, w);
        ctx->max_atoms = std::max(ctx->max_atoms, at);
    }
    if (t->sys_sel_off[ns] != t->n_sel || t->sys_wat_off[ns] != t->n_wat) return fail(ctx, HB_ERR_ARG, "topology: offsets do not cover the point tables");
    // validate indices on the host (cheap, once)
    for (int s = 0; s < ns; ++s) {
        long long at = t->sys_atom_off[s + 1] - t->sys_atom_off[s];
        for (int i = t->sys_sel_off[s]; i < t->sys_sel_off[s + 1]; ++i) {
            if (t->sel_atom[i] < 0 || t->sel_atom[i] >= at) return fail(ctx, HB_ERR_ARG, "topology: sel_atom out of range");
            int d = t->sel_don[i], a = t->sel_acc[i], w = t->sel_wat[i];
            if (d >= t->n_ent || a >= t->n_ent || w >= t->n_ent) return fail(ctx, HB_ERR_ARG, "topology: entry index out of range");
        }
        for (int i = t->sys_wat_off[s]; i < t->sys_wat_off[s + 1]; ++i) {
            if (t->wat_atom[i] < 0 || t->wat_atom[i] >= at) return fail(ctx, HB_ERR_ARG, "topology: wat_atom out of range");
            if (t->wat_ent[i] < 0 || t->wat_ent[i] >= t->n_ent) return fail(ctx, HB_ERR_ARG, "topology: wat_ent out of range");
        }
    }
    if (t->n_ent) {
        if (t->ent_h_off[0] != 0 || t->ent_h_off[t->n_ent] != t->n_ent_h) return fail(ctx, HB_ERR_ARG, "topology: hydrogen CSR inconsistent");
        for (int e = 0; e < t->n_ent; ++e)
            if (t->ent_h_off[e + 1] < t->ent_h_off[e]) return fail(ctx, HB_ERR_ARG, "topology: hydrogen CSR not monotone");
    }
    for (int k = 0; k < t->n_ent_h; ++k)
        if (t->ent_h_atom[k] < 0 || t->ent_h_atom[k] >= ctx->max_atoms) return fail(ctx, HB_ERR_ARG, "topology: hydrogen atom out of range");
    // atoms the kernels can touch (single system): selection points, water points, every listed hydrogen
    // (hb_set_ions adds the ions and rebuilds the ranges)
    ctx->atom_used.clear();
    if (ns == 1 && ctx->max_atoms > 0) {
        ctx->atom_used.assign((size_t)ctx->max_atoms, 0);
        for (int i = 0; i < t->n_sel; ++i) ctx->atom_used[t->sel_atom[i]] = 1;
        for (int i = 0; i < t->n_wat; ++i) ctx->atom_used[t->wat_atom[i]] = 1;
        for (int i = 0; i < t->n_ent_h; ++i) ctx->atom_used[t->ent_h_atom[i]] = 1;
    }
    build_copy_ranges(ctx);
    // the block of atoms that holds the selection points and their hydrogens (single system, waters excluded): the fused
    // kernel pulls it into L2 with one bulk prefetch when a frame starts, so its gathers do not wait for DRAM
    ctx->sel_block_lo = ctx->sel_block_n = 0;
    if (ns == 1 && t->n_sel > 0) {
        long long lo = t->sel_atom[0], hi = t->sel_atom[0], wlo = -1, whi = -2;
        for (int i = 0; i < t->n_sel; ++i) { lo = std::min<long long>(lo, t->sel_atom[i]); hi = std::max<long long>(hi, t->sel_atom[i]); }
        if (t->n_wat > 0) {
            wlo = whi = t->wat_atom[0];
            for (int i = 0; i < t->n_wat; ++i) { wlo = std::min<long long>(wlo, t->wat_atom[i]); whi = std::max<long long>(whi, t->wat_atom[i]); }
            whi += 8;      // (the hydrogens of the last water)
        }
        for (int i = 0; i < t->n_ent_h; ++i) {
            const long long h = t->ent_h_atom[i];
            if (h >= wlo && h <= whi) continue;
            lo = std::min(lo, h); hi = std::max(hi, h);
        }
        if ((hi - lo + 1) * 12 <= 512 * 1024) { ctx->sel_block_lo = lo; ctx->sel_block_n = hi - lo + 1; }
    }
    // water points at a constant atom stride (O, H, H, O, H, H ...) can be streamed as contiguous slabs
    ctx->water_regular = t->n_wat > 0;
    ctx->water_stride = 0;
    for (int s = 0; s < ns && ctx->water_regular; ++s) {
        int w0 = t->sys_wat_off[s], w1 = t->sys_wat_off[s + 1];
        for (int i = w0 + 1; i < w1; ++i) {
            int d = t->wat_atom[i] - t->wat_atom[i - 1];
            if (ctx->water_stride == 0) ctx->water_stride = d;
            if (d != ctx->water_stride || d < 1) { ctx->water_regular = false; break; }
        }
    }
    if (ctx->water_stride == 0) ctx->water_stride = 1;

    DevTopo& d = ctx->tp;
    d.n_systems = ns;
    int rc;
    static const int zero_off[2] = {0, 0};
    if ((rc = upload(ctx, (const long long*)t->sys_atom_off, ns + 1, &d.sys_atom_off, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sys_sel_off, ns + 1, &d.sys_sel_off, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sys_wat_off, ns + 1, &d.sys_wat_off, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sel_atom, t->n_sel, &d.sel_atom, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sel_don, t->n_sel, &d.sel_don, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sel_acc, t->n_sel, &d.sel_acc, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sel_wat, t->n_sel, &d.sel_wat, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->sel_bb, t->n_sel, &d.sel_bb, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->wat_atom, t->n_wat, &d.wat_atom, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->wat_ent, t->n_wat, &d.wat_ent, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->ent_group, t->n_ent, &d.ent_group, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->ent_resid, t->n_ent, &d.ent_resid, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->ent_residue, t->n_ent, &d.ent_residue, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->n_ent ? t->ent_h_off : zero_off, (size_t)t->n_ent + 1, &d.ent_h_off, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, t->ent_h_atom, t->n_ent_h, &d.ent_h_atom, ctx->topo_allocs))) return rc;
    {   // packed copies for the kernels
        std::vector<int4> sp((size_t)std::max(t->n_sel, 1)), ep((size_t)std::max(t->n_ent, 1));
        for (int i = 0; i < t->n_sel; ++i) sp[i] = make_int4(t->sel_don[i], t->sel_acc[i], t->sel_wat[i], t->sel_bb[i] ? 1 : 0);
        for (int e = 0; e < t->n_ent; ++e)
            ep[e] = make_int4(t->ent_resid[e], t->ent_residue[e], t->ent_h_off[e], t->ent_h_off[e + 1] - t->ent_h_off[e]);
        if ((rc = upload(ctx, sp.data(), sp.size(), &d.sel_pack, ctx->topo_allocs))) return rc;
        if ((rc = upload(ctx, ep.data(), ep.size(), &d.ent_pack, ctx->topo_allocs))) return rc;
        // inline hydrogen lists: offsets of the first four hydrogens from the heavy atom, one signed byte each
        d.h_wide = 0;
        auto hlist = [&](int e, int heavy_atom) {
            unsigned v = 0x80808080u;
            if (e >= 0) {
                int s0 = t->ent_h_off[e], n = t->ent_h_off[e + 1] - s0;
                if (n > 4) d.h_wide = 1;
                for (int k = 0; k < std::min(n, 4); ++k) {
                    int delta = t->ent_h_atom[s0 + k] - heavy_atom;
                    if (delta < -127 || delta > 127) { d.h_wide = 1; delta = 0; }
                    v = (v & ~(0xffu << (8 * k))) | ((unsigned)(delta & 0xff) << (8 * k));
                }
            }
            return (int)v;
        };
        std::vector<int2> sh((size_t)std::max(t->n_sel, 1));
        std::vector<int> wh((size_t)std::max(t->n_wat, 1));
        for (int i = 0; i < t->n_sel; ++i) sh[i] = make_int2(hlist(t->sel_don[i], t->sel_atom[i]), hlist(t->sel_wat[i], t->sel_atom[i]));
        for (int i = 0; i < t->n_wat; ++i) wh[i] = hlist(t->wat_ent[i], t->wat_atom[i]);
        if ((rc = upload(ctx, sh.data(), sh.size(), &d.sel_h, ctx->topo_allocs))) return rc;
        if ((rc = upload(ctx, wh.data(), wh.size(), &d.wat_h, ctx->topo_allocs))) return rc;
    }
    ctx->d_ent_group = (int*)d.ent_group;
    ctx->n_ent = t->n_ent;
    ctx->have_topo = true;
    return HB_OK;
}

int hb_set_ions(hb_ctx* ctx, const int32_t* ion_atom, const int32_t* ion_group, int32_t n_ions, int32_t include_water) {
    if (!ctx || n_ions < 0 || (n_ions > 0 && (!ion_atom || !ion_group))) return HB_ERR_ARG;
    if (!ctx->have_topo) return fail(ctx, HB_ERR_STATE, "hb_set_ions before hb_set_topology");
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_ions during a run");
    for (int i = 0; i < n_ions; ++i)
        if (ion_atom[i] < 0 || ion_atom[i] >= ctx->max_atoms || ion_group[i] < 0) return fail(ctx, HB_ERR_ARG, "hb_set_ions: index out of range");
    CK(cudaSetDevice(ctx->device));
    int rc;
    ctx->ions = IonTable{nullptr, nullptr, 0};
    if ((rc = upload(ctx, ion_atom, (size_t)n_ions, &ctx->ions.atom, ctx->topo_allocs))) return rc;
    if ((rc = upload(ctx, ion_group, (size_t)n_ions, &ctx->ions.group, ctx->topo_allocs))) return rc;
    ctx->ions.n = n_ions;
    ctx->ion_include_water = include_water != 0;
    // k_ion_contacts reads the ions' coordinates: they must be inside the atom ranges that reach the device
    if (!ctx->atom_used.empty()) {
        for (int i = 0; i < n_ions; ++i) ctx->atom_used[ion_atom[i]] = 1;
        build_copy_ranges(ctx);
    }
    return HB_OK;
}

int hb_set_wires(hb_ctx* ctx, const int32_t* ent_node, int32_t n_ent, int32_t n_res, int32_t max_water, int32_t allow_direct_bonds) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->have_topo) return fail(ctx, HB_ERR_STATE, "hb_set_wires before hb_set_topology");
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_wires during a run");
    if (!ent_node) { ctx->wire.on = false; return HB_OK; }
    if (n_ent != ctx->n_ent) return fail(ctx, HB_ERR_ARG, "hb_set_wires: size mismatch");
    if (ctx->tp.n_systems != 1) return fail(ctx, HB_ERR_ARG, "hb_set_wires: trajectories of one system only");
    if (n_res < 1 || max_water < 0 || max_water > 127) return fail(ctx, HB_ERR_ARG, "hb_set_wires: n_res >= 1 and 0 <= max_water <= 127");
    int n_wat_nodes = 1;
    for (int i = 0; i < n_ent; ++i) {
        if (ent_node[i] < 0 || ent_node[i] - n_res >= n_ent) return fail(ctx, HB_ERR_ARG, "hb_set_wires: node out of range");
        n_wat_nodes = std::max(n_wat_nodes, ent_node[i] - n_res + 1);
    }
    if (n_wat_nodes != ctx->wire.n_wat_nodes || n_res != ctx->wire.n_res) {   // the BFS scratch is sized by both
        free_list(ctx->wire.allocs);
        ctx->wire.d_edges = nullptr;
    }
    ctx->wire.n_wat_nodes = n_wat_nodes;
    CK(cudaSetDevice(ctx->device));
    if (ctx->wire.d_ent_node) cudaFree(ctx->wire.d_ent_node);
    ctx->wire.d_ent_node = nullptr;
    CK(cudaMalloc((void**)&ctx->wire.d_ent_node, std::max<size_t>((size_t)n_ent * sizeof(int), 16)));
    if (n_ent) CK(cudaMemcpy(ctx->wire.d_ent_node, ent_node, (size_t)n_ent * sizeof(int), cudaMemcpyHostToDevice));
    if (ctx->wire.d_direct_node) cudaFree(ctx->wire.d_direct_node);     // (set again after every hb_set_wires)
    ctx->wire.d_direct_node = nullptr;
    ctx->wire.n_res = n_res;
    ctx->wire.max_water = max_water;
    ctx->wire.allow_direct = allow_direct_bonds != 0;
    ctx->wire.on = true;
    return HB_OK;
}

int hb_set_wire_direct_nodes(hb_ctx* ctx, const int32_t* direct_node, int32_t n_ent) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->wire.on) return fail(ctx, HB_ERR_STATE, "hb_set_wire_direct_nodes before hb_set_wires");
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_wire_direct_nodes during a run");
    CK(cudaSetDevice(ctx->device));
    if (ctx->wire.d_direct_node) cudaFree(ctx->wire.d_direct_node);
    ctx->wire.d_direct_node = nullptr;
    if (!direct_node) return HB_OK;
    if (n_ent != ctx->n_ent) return fail(ctx, HB_ERR_ARG, "hb_set_wire_direct_nodes: size mismatch");
    for (int i = 0; i < n_ent; ++i)
        if (direct_node[i] < 0 || direct_node[i] >= ctx->wire.n_res + ctx->wire.n_wat_nodes) return fail(ctx, HB_ERR_ARG, "hb_set_wire_direct_nodes: node out of range");
    CK(cudaMalloc((void**)&ctx->wire.d_direct_node, std::max<size_t>((size_t)n_ent * sizeof(int), 16)));
    if (n_ent) CK(cudaMemcpy(ctx->wire.d_direct_node, direct_node, (size_t)n_ent * sizeof(int), cudaMemcpyHostToDevice));
    return HB_OK;
}

int hb_set_wire_virtual_waters(hb_ctx* ctx, const int64_t* frame_off, const int32_t* true_wat, const int32_t* label_wat, int64_t n_frames) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->have_topo) return fail(ctx, HB_ERR_STATE, "hb_set_wire_virtual_waters before hb_set_topology");
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_wire_virtual_waters during a run");
    CK(cudaSetDevice(ctx->device));
    free_list(ctx->wire.virt_allocs);
    ctx->wire.virt_on = false;
    ctx->wire.virt = VirtualWaters{nullptr, nullptr, nullptr, 0};
    if (!frame_off || n_frames <= 0) return HB_OK;
    if (ctx->tp.n_systems != 1) return fail(ctx, HB_ERR_ARG, "hb_set_wire_virtual_waters: trajectories of one system only");
    const int64_t total = frame_off[n_frames];
    if (frame_off[0] != 0 || total < 0 || (total > 0 && (!true_wat || !label_wat))) return fail(ctx, HB_ERR_ARG, "hb_set_wire_virtual_waters: bad offsets");
    for (int64_t f = 0; f < n_frames; ++f)
        if (frame_off[f + 1] < frame_off[f]) return fail(ctx, HB_ERR_ARG, "hb_set_wire_virtual_waters: offsets must not decrease");
    for (int64_t k = 0; k < total; ++k)
        if (true_wat[k] < 0 || true_wat[k] >= ctx->n_wat_total || label_wat[k] < 0 || label_wat[k] >= ctx->n_wat_total)
            return fail(ctx, HB_ERR_ARG, "hb_set_wire_virtual_waters: water point out of range");
    int rc;
    const long long* d_off = nullptr;
    const int *d_t = nullptr, *d_l = nullptr;
    if ((rc = upload(ctx, (const long long*)frame_off, (size_t)n_frames + 1, &d_off, ctx->wire.virt_allocs))) return rc;
    if ((rc = upload(ctx, true_wat, (size_t)total, &d_t, ctx->wire.virt_allocs))) return rc;
    if ((rc = upload(ctx, label_wat, (size_t)total, &d_l, ctx->wire.virt_allocs))) return rc;
    ctx->wire.virt = VirtualWaters{d_off, d_t, d_l, n_frames};
    ctx->wire.virt_on = true;
    return HB_OK;
}

int hb_set_entry_groups(hb_ctx* ctx, const int32_t* g, int32_t n) {
    if (!ctx || !g) return HB_ERR_ARG;
    if (!ctx->have_topo || n != ctx->n_ent) return fail(ctx, HB_ERR_ARG, "hb_set_entry_groups: size mismatch");
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_entry_groups during a run");
    CK(cudaSetDevice(ctx->device));
    if (n) CK(cudaMemcpy(ctx->d_ent_group, g, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
    return HB_OK;
}

int hb_set_params(hb_ctx* ctx, const hb_params* p) {
    if (!ctx || !p) return HB_ERR_ARG;
    if (ctx->running) return fail(ctx, HB_ERR_STATE, "hb_set_params during a run");
    if (!(p->distance > 0.0) || !std::isfinite(p->distance)) return fail(ctx, HB_ERR_ARG, "distance must be positive");
    if (p->method != HB_METHOD_IN_SELECTION && p->method != HB_METHOD_WATER_AROUND && p->method != HB_METHOD_ION_CONTACTS &&
        p->method != HB_METHOD_WATER_PRESENCE)
        return fail(ctx, HB_ERR_ARG, "unknown method");
    if (p->method == HB_METHOD_ION_CONTACTS && p->structure_batch) return fail(ctx, HB_ERR_ARG, "ion contacts: trajectories only");
    if ((p->method == HB_METHOD_WATER_AROUND || p->method == HB_METHOD_WATER_PRESENCE) &&
        (!(p->around_radius >= 0.0) || !std::isfinite(p->around_radius)))
        return fail(ctx, HB_ERR_ARG, "around_radius must be >= 0");
    if (p->pbc_mode != HB_PBC_NONE && p->pbc_mode != HB_PBC_ORTHO && p->pbc_mode != HB_PBC_TRICLINIC)
        return fail(ctx, HB_ERR_ARG, "unknown pbc_mode");
    if (ctx->have_params && (p->around_radius != ctx->prm.around_radius || p->method != ctx->prm.method)) ctx->fp_level = 0;
    ctx->prm = *p;
    RunParams& r = ctx->rp;
    r.r2 = p->distance * p->distance;
    r.around2 = p->around_radius * p->around_radius;
    // float32 pre-filter bounds: a little above the float64 thresholds (more under minimum image, where the
    // float32 reduction of the difference vector adds rounding of its own)
    const double slack = p->pbc_mode != HB_PBC_NONE ? 1e-4 : 1e-5;
    r.r2f_hi = nextafterf((float)(r.r2 * (1.0 + slack)), INFINITY);
    r.around2f_hi = nextafterf((float)(r.around2 * (1.0 + slack)), INFINITY);
    r.pbc = p->pbc_mode;
    r.cell2 = (float)(p->distance * 1.001);
    r.cell1 = (float)(std::max(p->around_radius, 1e-3) * 1.001);
    r.around_f = (float)(p->around_radius * 1.001 + 1e-6);
    r.cos_threshold = p->cos_threshold;
    r.check_angle = p->check_angle;
    r.excl_bb = p->exclude_backbone_backbone;
    // the presence method is the first half of the water-around pipeline (selection grid + local-water test)
    ctx->presence = p->method == HB_METHOD_WATER_PRESENCE;
    r.method = ctx->presence ? HB_METHOD_WATER_AROUND : p->method;
    r.self_pairs = 0;          // decided in hb_begin
    ctx->have_params = true;
    return HB_OK;
}

// Decide whether the one-CTA-per-frame kernel can serve this run and lay out its shared memory.
// `grow` > 0 doubles the local-water / pair capacities that many times (used after a fused overflow).
static bool plan_fused(hb_ctx* ctx, int grow = 0) {
    ctx->fused_active = false;
    if (!ctx->opt_fused || ctx->rp.method == HB_METHOD_ION_CONTACTS || ctx->presence) return false;
    if (ctx->fused_unfit_around >= 0.0 && ctx->rp.method == HB_METHOD_WATER_AROUND &&
        ctx->prm.around_radius >= ctx->fused_unfit_around) return false;
    const bool wa = ctx->rp.method == HB_METHOD_WATER_AROUND;
    FusedCfg fc{};
    fc.sel_cap = std::max(ctx->max_sel, 1);
    long long lw = 0;
    if (wa && ctx->max_wat > 0) {
        lw = ctx->opt_fused_lw_cap ? ctx->opt_fused_lw_cap : std::max<long long>(256, fc.sel_cap / 2);
        lw <<= grow;
        lw = std::min<long long>(lw, ctx->max_wat);
    }
    fc.lw_cap = (int)lw;
    if ((long long)fc.sel_cap + fc.lw_cap > 65535) return false;      // 16-bit cell cursors / pair indices
    fc.pair_cap = (int)((ctx->opt_fused_pair_cap ? ctx->opt_fused_pair_cap : 2048) << grow);
    fc.regular = (wa && ctx->water_regular && ctx->water_stride <= 8) ? 1 : 0;
    fc.w_stride = ctx->water_stride;
    fc.stages = fc.regular ? (int)std::min<long long>(std::max<long long>(ctx->opt_fused_stages, 2), FUSED_MAX_STAGES) : 0;
    fc.prefetch = (int)ctx->opt_fused_prefetch;
    fc.debug = ctx->opt_fused_debug;
    fc.wait_ns = (unsigned)ctx->opt_fused_wait_ns;
    fc.stage_bytes = fc.regular ? (unsigned)(((FUSED_TILE - 1) * fc.w_stride * 12 + 12 + 15 + 15) & ~15) : 0u;
    fc.task_cap = 1536 << grow;
    // Regions by lifetime (every region 16-byte aligned):
    //   sorted1 | pairs   selection points in g1 order (g1 sort .. phase C sort), then the pair list (phase C1 ..)
    //   ring              tile stages (phase B), gathered candidate coordinates (B2), then psorted + tasks + flags (phase C)
    //   cand, lw          candidate indices (B .. B2) and local waters (B2 .. C sort); before the g1 sort the two together
    //                     hold the selection points in gather order
    //   cells, near       cell counters of the current grid, near bitmap of g1
    auto al16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    fc.cand16 = ctx->max_wat <= 65535 ? 1 : 0;
    const size_t sorted_b = al16(std::max((size_t)fc.sel_cap * 16, (size_t)fc.pair_cap * 4));
    const size_t ring_b = al16(std::max((size_t)fc.stages * fc.stage_bytes,
                                        ((size_t)fc.sel_cap + fc.lw_cap) * 16 + (size_t)fc.task_cap * 8 + FUSED_CONSUMERS * 4));
    const size_t lw_b = (size_t)fc.lw_cap * 16;
    size_t cand_b = 0;
    if (wa) {
        fc.cand_cap = 2048 << std::min(grow, 3);
        cand_b = al16((size_t)fc.cand_cap * (fc.cand16 ? 2 : 4));
        if (cand_b + lw_b < (size_t)fc.sel_cap * 16) cand_b = al16((size_t)fc.sel_cap * 16 - lw_b);
        fc.cand_cap = (int)(cand_b / (fc.cand16 ? 2 : 4));
    }
    size_t off = al16(sizeof(FusedShared));
    fc.off_sorted = (unsigned)off;  off += sorted_b;
    fc.off_ring = (unsigned)off;    off += ring_b;
    fc.off_cand = (unsigned)off;    off += cand_b;
    fc.off_lw = (unsigned)off;      off += lw_b;
    fc.off_bar = (unsigned)off;     off += (size_t)2 * FUSED_MAX_STAGES * 8;
    // the cell counters take what is left of the shared memory that still lets four frames share an SM (finer cells =
    // fewer distance tests in phases B2 and C1): 2 bytes per cell + 1 bit for the near bitmap
    long long cc = ctx->opt_fused_cell_cap;
    if (!cc) {
        const long long budget = (long long)(ctx->prop.sharedMemPerMultiprocessor / 4) - (long long)ctx->prop.reservedSharedMemPerBlock - 64;
        cc = ((budget - (long long)off - 64) * 8 / 17) & ~31ll;
        cc = std::min<long long>(std::max<long long>(cc, 3712), 8192);
    }
    fc.cell_cap = (int)std::min<long long>(cc, 32768);
    fc.off_cells = (unsigned)off;   off += al16(((size_t)fc.cell_cap + 2) * 2);
    fc.off_near = (unsigned)off;    off += al16(((size_t)fc.cell_cap + 31) / 32 * 4);
    fc.smem_bytes = (unsigned)off;
    if (fc.smem_bytes > 112 * 1024) return false;                     // keep at least two frames resident per SM
    {   // the attribute belongs to the function (all contexts of the device share it): it is only ever raised
        static std::mutex mu;
        static long long attr_smem[64];
        std::lock_guard<std::mutex> lock(mu);
        const int d = ctx->device & 63;
        if ((long long)fc.smem_bytes > attr_smem[d]) {
            if (cudaFuncSetAttribute(k_frame_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fc.smem_bytes) != cudaSuccess ||
                cudaFuncSetAttribute(k_frame_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fc.smem_bytes) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            attr_smem[d] = (long long)fc.smem_bytes;
        }
    }
    ctx->fc = fc;
    ctx->fused_grow = grow;
    ctx->fused_active = true;
    return true;
}

static int alloc_table(hb_ctx* ctx, BondTable& bt, long long cap, int wpr) {
    bt = BondTable{};
    bt.mask = (unsigned)(cap - 1);
    bt.wpr = wpr;
    CK(cudaMalloc((void**)&bt.keys, (size_t)cap * sizeof(unsigned long long)));
    CK(cudaMalloc((void**)&bt.bits, (size_t)cap * wpr * sizeof(unsigned)));
    k_fill_u64<<<1184, 256, 0, ctx->s_compute>>>(bt.keys, (size_t)cap, EMPTY_KEY);
    CK(cudaMemsetAsync(bt.bits, 0, (size_t)cap * wpr * sizeof(unsigned), ctx->s_compute));
    CK(cudaGetLastError());
    return HB_OK;
}

int hb_begin(hb_ctx* ctx, int64_t n_frames_total) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->have_topo || !ctx->have_params) return fail(ctx, HB_ERR_STATE, "hb_begin before topology/params");
    if (n_frames_total < 1) return fail(ctx, HB_ERR_ARG, "n_frames_total must be >= 1");
    if (ctx->prm.structure_batch && n_frames_total != ctx->tp.n_systems)
        return fail(ctx, HB_ERR_ARG, "structure_batch: n_frames_total must equal n_systems");
    CK(cudaSetDevice(ctx->device));
    Trace tr(ctx, "begin");
    CK(cudaStreamSynchronize(ctx->s_copy));
    CK(cudaStreamSynchronize(ctx->s_compute));
    drain_events(ctx);
    tr.mark("sync");
    int wpr = (int)((n_frames_total + 31) / 32);
    if (ctx->wire.on) {        // one byte per frame, then the two "first direct bond" words (hb_wires.cuh)
        const bool method_ok = ctx->wire.virt_on ? ctx->rp.method == HB_METHOD_IN_SELECTION : ctx->rp.method == HB_METHOD_WATER_AROUND;
        if (ctx->prm.structure_batch || !method_ok || ctx->presence)
            return fail(ctx, HB_ERR_ARG, "water wires: HB_METHOD_WATER_AROUND on a trajectory (HB_METHOD_IN_SELECTION with hb_set_wire_virtual_waters)");
        if (ctx->wire.virt_on && (ctx->prm.pbc_mode != HB_PBC_NONE || n_frames_total > ctx->wire.virt.n_frames))
            return fail(ctx, HB_ERR_ARG, "water wires with virtual waters: no pbc_mode, and a water list for every frame");
        ctx->wire.nb4 = (int)((((n_frames_total + 3) / 4) + 1) & ~1ll);
        wpr = ctx->wire.nb4 + 2;
    }
    // a point paired with itself: with the angle criterion ANG(P, h, P) is 180 degrees (cos_arg = -1) and fails
    // unless the threshold admits -1; without it both entries share one residue and are skipped (hbond.py:125-127),
    // except for wires, which have no same-residue rule
    ctx->rp.self_pairs = ctx->wire.on ? ((!ctx->prm.check_angle || !(ctx->prm.cos_threshold > -1.0f)) ? 1 : 0)
                                      : ((ctx->prm.check_angle && !(ctx->prm.cos_threshold > -1.0f)) ? 1 : 0);
    long long want = ctx->opt_table_capacity;
    if (!want) {   // a few slots per selection point (+ as many local waters); the table doubles on demand
        want = 4 * (ctx->n_sel_total + (ctx->rp.method == HB_METHOD_WATER_AROUND ? std::min(ctx->n_sel_total, ctx->n_wat_total) : 0));
        want = std::min<long long>(std::max<long long>(want, 4096), 1ll << 22);
    }
    long long cap = 1024;
    while (cap < want) cap <<= 1;
    size_t npts = (size_t)ctx->max_sel + ctx->max_wat;
    long long cc = 4096;
    while (cc < (long long)(2 * npts)) cc <<= 1;
    cc = std::min<long long>(cc, 1ll << 22);
    // device budget: half for the two input buffers of hb_submit_frames, half for the general-path scratch
    auto round32 = [](long long f) { return f >= 32 ? f - f % 32 : f; };
    size_t per_frame_scratch = sizeof(FrameState) + (size_t)ctx->max_sel * 32 + (size_t)ctx->max_wat * 16 + npts * 16 +
                               2 * (size_t)(cc + 1) * 4;
    long long fb = (long long)(ctx->opt_batch_bytes / 2 / std::max<size_t>(per_frame_scratch, 1));
    fb = round32(std::min<long long>(std::max<long long>(1, std::min<long long>(fb, ctx->opt_max_batch_frames)), n_frames_total));
    long long hb = (long long)(ctx->opt_batch_bytes / 2 / std::max<size_t>((size_t)ctx->max_atoms * 24, 1));
    hb = round32(std::min<long long>(std::max<long long>(1, std::min<long long>(hb, ctx->opt_max_batch_frames)), n_frames_total));
    {
        const int hint = (ctx->rp.method == HB_METHOD_WATER_AROUND && ctx->prm.around_radius >= ctx->fused_hint_around) ? ctx->fused_grow_hint : 0;
        if (!plan_fused(ctx, hint) && hint) plan_fused(ctx, 0);
    }
    tr.mark("plan");
    // allocations of the previous run are reused when nothing that sizes them changed
    long long sig[8] = {wpr, ctx->max_sel, ctx->max_wat, ctx->max_atoms, fb, cc, ctx->prm.structure_batch + 2 * (ctx->wire.on ? ctx->wire.n_res + 1 : 0), hb};
    bool reuse = ctx->bt.keys && ctx->table_cap >= cap && memcmp(sig, ctx->alloc_sig, sizeof(sig)) == 0;
    if (!reuse) free_run(ctx);
    ctx->tm = hb_timers{};
    CK(cudaEventRecord(ctx->ev_run0, ctx->s_compute));
    ctx->n_frames_total = n_frames_total;
    ctx->next_frame = 0;
    ctx->batch_counter = 0;
    ctx->table_grows = 0;
    ctx->fused_fallbacks = 0;
    ctx->fin_queued = false;
    for (auto& sl : ctx->slots) sl.busy = false;
    int rc;
    if (reuse) {
        size_t tc = (size_t)ctx->table_cap;
        k_fill_u64<<<1184, 256, 0, ctx->s_compute>>>(ctx->bt.keys, tc, EMPTY_KEY);
        ctx->tm.n_launches += 1;
        CK(cudaMemsetAsync(ctx->bt.bits, 0, tc * wpr * sizeof(unsigned), ctx->s_compute));
        CK(cudaMemsetAsync(ctx->bt.n_keys, 0, 64 * sizeof(unsigned long long), ctx->s_compute));
        if (ctx->trace) { cudaStreamSynchronize(ctx->s_compute); tr.mark("clear(reuse)"); }
        ctx->running = true;
        ctx->finalized = false;
        return HB_OK;
    }
    // bond table
    if ((rc = alloc_table(ctx, ctx->bt, cap, wpr))) return rc;
    ctx->table_cap = cap;
    CK(cudaMalloc((void**)&ctx->bt.n_keys, 64 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->bt.n_keys, 0, 64 * sizeof(unsigned long long), ctx->s_compute));
    ctx->bt.overflow = ctx->bt.n_keys + 1;
    ctx->bt.fused_overflow = ctx->bt.n_keys + 2;
    ctx->bt.n_hits = (unsigned long long*)(ctx->bt.n_keys + 4);
    ctx->bt.phase_cycles = ctx->trace ? (unsigned long long*)(ctx->bt.n_keys + 8) : nullptr;

    ctx->cell_cap = (int)cc;
    ctx->batch_frames = (int)fb;
    ctx->host_batch_frames = (int)hb;
    BatchView& bv = ctx->bv;
    bv = BatchView{};
    bv.max_sel = ctx->max_sel;
    bv.max_wat = ctx->max_wat;
    bv.cell_cap = ctx->cell_cap;
    bv.structure_batch = ctx->prm.structure_batch;
    ctx->in_bytes = (size_t)hb * ctx->max_atoms * 12;
    memcpy(ctx->alloc_sig, sig, sizeof(sig));
    ctx->running = true;
    ctx->finalized = false;
    return HB_OK;
}

static int grow_table(hb_ctx* ctx) {
    long long ncap = ctx->table_cap * 2;
    BondTable nt;
    int rc = alloc_table(ctx, nt, ncap, ctx->bt.wpr);
    if (rc) return rc;
    nt.n_keys = ctx->bt.n_keys;
    nt.overflow = ctx->bt.overflow;
    nt.fused_overflow = ctx->bt.fused_overflow;
    nt.n_hits = ctx->bt.n_hits;
    nt.phase_cycles = ctx->bt.phase_cycles;
    k_rehash<<<(unsigned)ctx->table_cap, 128, 0, ctx->s_compute>>>(ctx->bt, nt);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->s_compute));
    cudaFree(ctx->bt.keys);
    cudaFree(ctx->bt.bits);
    ctx->bt = nt;
    ctx->table_cap = ncap;
    ctx->table_grows++;
    return HB_OK;
}

// shared-memory layout of k_frame_pairs for the current level; false when the selection alone does not fit
static bool plan_frame_pairs(hb_ctx* ctx) {
    if (!ctx->opt_frame_pairs || ctx->fp_level > 0) return false;
    FramePairsCfg c{};
    c.cell_cap = 16384;
    c.pair_cap = 8192;
    c.task_cap = 4096;
    size_t off = (sizeof(FpShared) + 15) & ~(size_t)15;
    c.off_pairs = (unsigned)off; off += (size_t)c.pair_cap * 4;
    c.off_tasks = (unsigned)off; off += (size_t)c.task_cap * 8;
    c.off_flags = (unsigned)off; off += (size_t)FP_THREADS * 4;
    c.off_cells = (unsigned)off; off += (((size_t)c.cell_cap + 2) * 2 + 15) & ~(size_t)15;
    c.off_pts = (unsigned)off;
    const size_t budget = 226 * 1024;
    if (off + 16 * 64 > budget) return false;
    long long np = (long long)((budget - off) / 16);
    np = std::min<long long>(np, std::min<long long>((long long)ctx->max_sel + ctx->max_wat, 65535));
    if (np < ctx->max_sel || np < 1) return false;
    c.np_cap = (int)np;
    c.smem_bytes = (unsigned)(off + (size_t)np * 16);
    c.overflow = ctx->bt.n_keys + 7;
    if (cudaFuncSetAttribute(k_frame_pairs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(k_frame_pairs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    ctx->fp = c;
    return true;
}

// the general path's per-batch scratch is allocated the first time that path runs
static int ensure_scratch(hb_ctx* ctx) {
    if (!ctx->scratch_allocs.empty()) return HB_OK;
    BatchView& bv = ctx->bv;
    size_t fb = (size_t)ctx->batch_frames, cc = (size_t)ctx->cell_cap, npts = (size_t)ctx->max_sel + ctx->max_wat;
    auto dalloc = [&](void** p, size_t bytes) -> int {
        CK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
        ctx->scratch_allocs.push_back(*p);
        return HB_OK;
    };
    int rc;
    if ((rc = dalloc((void**)&bv.fs, fb * sizeof(FrameState)))) return rc;
    if ((rc = dalloc((void**)&bv.selpos, fb * (size_t)ctx->max_sel * 16))) return rc;
    if ((rc = dalloc((void**)&bv.sorted1, fb * (size_t)ctx->max_sel * 16))) return rc;
    if ((rc = dalloc((void**)&bv.lw, fb * (size_t)ctx->max_wat * 16))) return rc;
    if ((rc = dalloc((void**)&bv.sorted2, fb * npts * 16))) return rc;
    if ((rc = dalloc((void**)&bv.cells1, fb * (cc + 1) * 4))) return rc;
    if ((rc = dalloc((void**)&bv.cells2, fb * (cc + 1) * 4))) return rc;
    ctx->pairbuf = PairBuf{};
    if (ctx->use_pairbuf && ctx->opt_pair_buf_bytes >= 4096 && npts < (1u << 24) && fb < 65536) {
        ctx->pairbuf.cap = (unsigned long long)ctx->opt_pair_buf_bytes / 8;
        if ((rc = dalloc((void**)&ctx->pairbuf.rec, ctx->pairbuf.cap * 8))) return rc;
        if ((rc = dalloc((void**)&ctx->pairbuf.count, 8))) return rc;
        ctx->pairbuf.overflow = ctx->bt.n_keys + 3;          // fourth word of the status block
    }
    return HB_OK;
}

// launch the kernel pipeline for one batch resident at d_xyz
}  // extern "C" (templates need C++ linkage)
template <bool PBC>
static int run_general_t(hb_ctx* ctx, BatchView bv) {
    cudaStream_t st = ctx->s_compute;
    const RunParams& rp = ctx->rp;
    const DevTopo& tp = ctx->tp;
    const int nf = bv.n_frames;
    bool wa = rp.method == HB_METHOD_WATER_AROUND;
    { KTimer t(ctx, KC_INIT); k_init_frames<<<(nf + 127) / 128, 128, 0, st>>>(bv); }
    dim3 gsel((bv.max_sel + 255) / 256, nf), gwat((bv.max_wat + 255) / 256, nf);
    if (bv.max_sel > 0) { KTimer t(ctx, KC_GATHER); k_gather_sel<PBC><<<gsel, 256, 0, st>>>(bv, tp); }
    { KTimer t(ctx, KC_GRID); k_grid_setup<PBC><<<(nf + 127) / 128, 128, 0, st>>>(bv, tp, rp); }
    // only the cells each frame's grid really has are cleared (the capacity is sized for the worst case)
    dim3 gzero(16, nf);
    const bool frame_pairs = !ctx->presence && plan_frame_pairs(ctx);
    if (wa) { KTimer t(ctx, KC_INIT); k_zero_cells<<<gzero, 256, 0, st>>>(bv, 1); }
    if (!frame_pairs) { KTimer t(ctx, KC_INIT); k_zero_cells<<<gzero, 256, 0, st>>>(bv, 2); }
    if (wa && bv.max_sel > 0 && bv.max_wat > 0) {
        { KTimer t(ctx, KC_COUNT); k_count<PBC><<<gsel, 256, 0, st>>>(bv, tp, 1); }
        { KTimer t(ctx, KC_SCAN); k_scan<<<nf, 1024, 0, st>>>(bv, 1); }
        { KTimer t(ctx, KC_SCATTER); k_scatter<PBC><<<gsel, 256, 0, st>>>(bv, tp, 1); }
        const size_t lw_smem = (size_t)bv.max_sel * 16 + (LW_SMEM_CELLS + 1) * 4;
        if (ctx->opt_lw_smem && lw_smem <= 96 * 1024) {
            // selection grid staged in shared memory, 2048 waters per block
            const int per_block = 2048;
            if (!ctx->lw_attr_set[PBC ? 1 : 0]) {      // per context: the attribute belongs to the context's device
                CK(cudaFuncSetAttribute(k_local_water_smem<PBC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                ctx->lw_attr_set[PBC ? 1 : 0] = true;
            }
            dim3 gw((bv.max_wat + per_block - 1) / per_block, nf);
            KTimer t(ctx, KC_LOCAL);
            k_local_water_smem<PBC><<<gw, 256, lw_smem, st>>>(bv, tp, rp, per_block);
        } else {
            KTimer t(ctx, KC_LOCAL);
            k_local_water<PBC><<<gwat, 256, 0, st>>>(bv, tp, rp);
        }
    }
    if (ctx->presence) {
        if (bv.max_sel > 0 && bv.max_wat > 0) {
            KTimer t(ctx, KC_PAIRS);
            k_presence_emit<<<dim3(8, nf), 256, 0, st>>>(bv, tp, ctx->bt_run);
        }
        CK(cudaGetLastError());
        return HB_OK;
    }
    int maxp = bv.max_sel + (wa ? bv.max_wat : 0);
    if (frame_pairs && maxp > 0) {
        KTimer t(ctx, KC_PAIRS);
        k_frame_pairs<PBC><<<nf, FP_THREADS, ctx->fp.smem_bytes, st>>>(bv, tp, rp, ctx->bt_run, ctx->fp);
    } else if (maxp > 0) {
        // selection + local waters: the kernels stride over the points a frame really has (far fewer than all waters)
        dim3 gall(std::min((maxp + 255) / 256, 48), nf);
        { KTimer t(ctx, KC_COUNT); k_count<PBC><<<gall, 256, 0, st>>>(bv, tp, 2); }
        { KTimer t(ctx, KC_SCAN); k_scan<<<nf, 1024, 0, st>>>(bv, 2); }
        { KTimer t(ctx, KC_SCATTER); k_scatter<PBC><<<gall, 256, 0, st>>>(bv, tp, 2); }
        dim3 gp(std::min((maxp + 127) / 128, 96), nf);
        if (ctx->pairbuf.rec && ctx->use_pairbuf) {
            PairBuf pb = ctx->pairbuf;
            pb.overflow = ctx->bt.n_keys + 3;
            CK(cudaMemsetAsync(pb.count, 0, 8, st));
            { KTimer t(ctx, KC_PAIRS); k_collect_pairs<PBC><<<gp, 128, 0, st>>>(bv, tp, rp, ctx->bt_run, pb); }
            { KTimer t(ctx, KC_PAIRS); k_expand_pairs<PBC><<<148 * 8, 128, 0, st>>>(bv, tp, rp, ctx->bt_run, pb); }
        } else {
            KTimer t(ctx, KC_PAIRS);
            k_pairs<PBC><<<gp, 128, 0, st>>>(bv, tp, rp, ctx->bt_run);
        }
    }
    CK(cudaGetLastError());
    return HB_OK;
}
extern "C" {
static int run_general(hb_ctx* ctx, const BatchView& bv) {
    return ctx->rp.pbc != HB_PBC_NONE ? run_general_t<true>(ctx, bv) : run_general_t<false>(ctx, bv);
}

// launch the kernels for one batch resident at d_xyz: one fused launch, or the general pipeline over
// sub-batches of at most `batch_frames` frames (the capacity of its scratch)
static int run_search(hb_ctx* ctx, const float* d_xyz, long long frame0, int nf, long long atoms_per_frame,
                      const float* buf0, size_t buf_bytes, const float* d_box, bool allow_fused) {
    BatchView bv = ctx->bv;
    bv.xyz = d_xyz;
    bv.box = d_box;
    bv.frame0 = frame0;
    bv.n_frames = nf;
    bv.atoms_per_frame = atoms_per_frame;
    bv.atom_base = bv.structure_batch ? ctx->h_sys_atom_off[frame0] : 0;
    cudaStream_t st = ctx->s_compute;
    if (ctx->fused_active && allow_fused) {
        FusedCfg fc = ctx->fc;
        fc.buf_begin = (unsigned long long)(uintptr_t)buf0;
        fc.buf_end = fc.buf_begin + buf_bytes;
        fc.n_sm = (unsigned)ctx->prop.multiProcessorCount;
        fc.pf_off = fc.pf_bytes = 0;
        if (ctx->opt_fused_sel_prefetch && !bv.structure_batch && ctx->sel_block_n > 0) {
            fc.pf_off = (unsigned)(ctx->sel_block_lo * 12);
            fc.pf_bytes = (unsigned)(ctx->sel_block_n * 12);
        }
        fc.max_streams = 0;
        fc.sm_streams = nullptr;
        if (ctx->opt_fused_max_streams > 0) {
            if (!ctx->d_sm_streams && cudaMalloc((void**)&ctx->d_sm_streams, 1024 * sizeof(int)) == cudaSuccess)
                cudaMemsetAsync(ctx->d_sm_streams, 0, 1024 * sizeof(int), st);
            if (ctx->d_sm_streams && fc.n_sm <= 1024 && (long long)nf >= 8ll * fc.n_sm) {   // (short launches: nothing to keep apart)
                cudaMemsetAsync(ctx->d_sm_streams, 0, fc.n_sm * sizeof(int), st);          // a failed launch must not leave counts behind
                fc.sm_streams = ctx->d_sm_streams;
                fc.max_streams = (int)ctx->opt_fused_max_streams;
            }
        }
        fc.stagger = 0;
        if (ctx->opt_fused_stagger > 0) {
            const long long occ_key = (long long)fc.smem_bytes * 2 + (ctx->rp.pbc != HB_PBC_NONE ? 1 : 0);
            if (ctx->occ_key != occ_key) {            // (asked once per shared-memory plan, not per launch)
                int q = 0;
                if (ctx->rp.pbc != HB_PBC_NONE) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_frame_fused<true>, FUSED_THREADS, fc.smem_bytes);
                else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, k_frame_fused<false>, FUSED_THREADS, fc.smem_bytes);
                ctx->occ_per_sm = q;
                ctx->occ_key = occ_key;
            }
            const int per_sm = ctx->occ_per_sm;
            fc.n_sm = (unsigned)ctx->prop.multiProcessorCount;
            fc.wave_ctas = (unsigned)std::max(per_sm, 1) * fc.n_sm;
            if ((long long)nf >= 2ll * fc.wave_ctas) fc.stagger = (unsigned)ctx->opt_fused_stagger;   // short launches: the delay would not pay
        }
        const char* trace_path = getenv("HBGPU_CTA_TRACE");      // diagnostic: per-CTA wall-clock marks appended to this file
        fc.cta_trace = nullptr;
        if (trace_path && cudaMalloc((void**)&fc.cta_trace, (size_t)nf * 64) == cudaSuccess) cudaMemsetAsync(fc.cta_trace, 0, (size_t)nf * 64, st);
        {
            KTimer t(ctx, KC_FUSED);
            if (ctx->rp.pbc != HB_PBC_NONE) k_frame_fused<true><<<nf, FUSED_THREADS, fc.smem_bytes, st>>>(bv, ctx->tp, ctx->rp, ctx->bt_run, fc);
            else k_frame_fused<false><<<nf, FUSED_THREADS, fc.smem_bytes, st>>>(bv, ctx->tp, ctx->rp, ctx->bt_run, fc);
        }
        CK(cudaGetLastError());
        if (fc.cta_trace) {
            std::vector<unsigned long long> h((size_t)nf * 8);
            cudaStreamSynchronize(st);
            if (cudaMemcpy(h.data(), fc.cta_trace, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
                if (FILE* f = fopen(trace_path, "ab")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
            }
            cudaFree(fc.cta_trace);
        }
        ctx->tm.fused_frames += nf;
        ctx->last_batch_fused = true;
    } else if (ctx->rp.method == HB_METHOD_ION_CONTACTS) {
        ctx->last_batch_fused = false;
        const long long pts = (long long)ctx->max_sel + (ctx->ion_include_water ? ctx->max_wat : 0);
        if (ctx->ions.n > 0 && pts > 0) {
            dim3 grid((unsigned)((pts + 255) / 256), (unsigned)nf);
            KTimer t(ctx, KC_PAIRS);
            if (ctx->rp.pbc != HB_PBC_NONE) k_ion_contacts<true><<<grid, 256, 0, st>>>(bv, ctx->tp, ctx->rp, ctx->bt_run, ctx->ions, ctx->ion_include_water);
            else k_ion_contacts<false><<<grid, 256, 0, st>>>(bv, ctx->tp, ctx->rp, ctx->bt_run, ctx->ions, ctx->ion_include_water);
            CK(cudaGetLastError());
        }
    } else {
        ctx->last_batch_fused = false;
        int rc = ensure_scratch(ctx);
        if (rc) return rc;
        bv = ctx->bv;                     // scratch pointers
        bv.box = d_box;
        bv.atoms_per_frame = atoms_per_frame;
        bv.atom_base = bv.structure_batch ? ctx->h_sys_atom_off[frame0] : 0;
        for (int done = 0; done < nf; done += ctx->batch_frames) {
            bv.n_frames = std::min(nf - done, ctx->batch_frames);
            bv.frame0 = frame0 + done;
            bv.xyz = bv.structure_batch ? d_xyz : d_xyz + (size_t)done * atoms_per_frame * 3;
            bv.box = d_box ? d_box + (size_t)done * 9 : nullptr;
            if ((rc = run_general(ctx, bv))) return rc;
        }
    }
    return HB_OK;
}

// edge lists and the per-CTA scratch of k_wire_bfs (allocated the first time a wire run launches; the edge lists are
// re-allocated with twice the capacity after a wire overflow)
static int ensure_wire_buffers(hb_ctx* ctx) {
    hb_ctx::Wire& w = ctx->wire;
    if (w.d_edges) return HB_OK;
    free_list(w.allocs);
    auto dalloc = [&](void** p, size_t bytes) -> int {
        CK(cudaMalloc(p, std::max<size_t>(bytes, 16)));
        w.allocs.push_back(*p);
        return HB_OK;
    };
    if (!w.edge_cap) {
        long long guess = ctx->opt_wire_edge_cap ? ctx->opt_wire_edge_cap
                                                 : 4ll * ctx->max_sel + std::min<long long>(4ll * ctx->max_wat, 16384);
        w.edge_cap = std::max<long long>(guess, 256);
    }
    long long fr = (long long)((512ll << 20) / (w.edge_cap * 8));
    w.frames = (int)std::min<long long>(std::max<long long>(fr, 1), 1024);
    const int n_wat = w.n_wat_nodes, n_res = w.n_res;
    const int cap_words = std::min((n_res + 31) / 32, WIRE_CHUNK_WORDS);
    const size_t per_cta = (size_t)n_wat * 8 + (size_t)n_res * 8 + 3 * (size_t)n_wat * cap_words * 4 + 2 * (size_t)n_res * cap_words * 4 +
                           (size_t)w.edge_cap * 8;
    long long ncta = std::min<long long>((long long)ctx->opt_wire_ctas_per_sm * ctx->prop.multiProcessorCount, (long long)((2ll << 30) / std::max<size_t>(per_cta, 1)));
    w.n_cta = (int)std::max<long long>(1, std::min<long long>(ncta, w.frames));
    int rc;
    WireCfg c{};
    if ((rc = dalloc((void**)&w.d_edges, (size_t)w.frames * w.edge_cap * 8))) return rc;
    if ((rc = dalloc((void**)&w.d_cnt, (size_t)w.frames * 4))) return rc;
    if ((rc = dalloc((void**)&w.d_sink, sizeof(EdgeSink)))) return rc;
    if ((rc = dalloc((void**)&c.wl_of, (size_t)w.n_cta * n_wat * 4))) return rc;
    if ((rc = dalloc((void**)&c.lw_list, (size_t)w.n_cta * n_wat * 4))) return rc;
    if ((rc = dalloc((void**)&c.ar_of, (size_t)w.n_cta * n_res * 4))) return rc;
    if ((rc = dalloc((void**)&c.ar_list, (size_t)w.n_cta * n_res * 4))) return rc;
    if ((rc = dalloc((void**)&c.wbits, (size_t)w.n_cta * 3 * n_wat * cap_words * 4))) return rc;
    if ((rc = dalloc((void**)&c.tbits, (size_t)w.n_cta * 2 * n_res * cap_words * 4))) return rc;
    if ((rc = dalloc((void**)&c.rlist, (size_t)w.n_cta * w.edge_cap * 8))) return rc;
    CK(cudaFuncSetAttribute(k_wire_bfs, cudaFuncAttributeMaxDynamicSharedMemorySize, WIRE_SMEM_WORDS * 4));
    CK(cudaMemsetAsync(c.wl_of, 0xFF, (size_t)w.n_cta * n_wat * 4, ctx->s_compute));
    CK(cudaMemsetAsync(c.ar_of, 0xFF, (size_t)w.n_cta * n_res * 4, ctx->s_compute));
    c.ent_node = w.d_ent_node;
    c.n_res = n_res;
    c.n_wat = n_wat;
    c.cap_words = cap_words;
    w.cfg = c;
    return HB_OK;
}

// launch the kernels for one batch resident at d_xyz (wire mode: rounds of search + breadth-first search over as
// many frames as the edge lists hold)
static int run_batch(hb_ctx* ctx, const float* d_xyz, long long frame0, int nf, long long atoms_per_frame,
                     size_t xyz_bytes, const float* d_box, bool allow_fused = true) {
    cudaStream_t st = ctx->s_compute;
    cudaEvent_t t0 = get_event(ctx), t1 = get_event(ctx);
    cudaEventRecord(t0, st);
    int rc;
    ctx->bt_run = ctx->bt;
    if (!ctx->wire.on) {
        if ((rc = run_search(ctx, d_xyz, frame0, nf, atoms_per_frame, d_xyz, xyz_bytes, d_box, allow_fused))) return rc;
    } else {
        if ((rc = ensure_wire_buffers(ctx))) return rc;
        hb_ctx::Wire& w = ctx->wire;
        WireCfg c = w.cfg;
        c.ent_node = w.d_ent_node;
        c.direct_node = w.d_direct_node ? w.d_direct_node : w.d_ent_node;
        c.max_water = w.max_water;
        c.allow_direct = w.allow_direct;
        c.nb4 = w.nb4;
        c.status = ctx->bt.n_keys;
        c.smem_words = (int)std::max<long long>(0, std::min<long long>(ctx->opt_wire_smem_words, WIRE_SMEM_WORDS));
        for (int done = 0; done < nf; done += w.frames) {
            const int n = std::min(nf - done, w.frames);
            CK(cudaMemsetAsync(w.d_cnt, 0, (size_t)n * 4, st));
            // the sink descriptor of this round (stream ordered: the previous round's kernels are done with theirs)
            EdgeSink sk{};
            sk.edges = w.d_edges;
            sk.edge_cnt = w.d_cnt;
            sk.edge_cap = (int)w.edge_cap;
            sk.edge_frame0 = frame0 + done;
            sk.wire_overflow = ctx->bt.n_keys + 6;
            k_set_sink<<<1, 1, 0, st>>>(w.d_sink, sk);
            ctx->bt_run = ctx->bt;
            ctx->bt_run.sink = w.d_sink;
            if ((rc = run_search(ctx, d_xyz + (size_t)done * atoms_per_frame * 3, frame0 + done, n, atoms_per_frame, d_xyz,
                                 xyz_bytes, d_box ? d_box + (size_t)done * 9 : nullptr, allow_fused))) return rc;
            if (w.virt_on) {       // convex-hull option: the selection-water and water-water pairs over the host's water lists
                BatchView bv = ctx->bv;
                bv.xyz = d_xyz + (size_t)done * atoms_per_frame * 3;
                bv.box = nullptr;
                bv.frame0 = frame0 + done;
                bv.n_frames = n;
                bv.atoms_per_frame = atoms_per_frame;
                bv.atom_base = 0;
                KTimer t(ctx, KC_PAIRS);
                k_virtual_pairs<<<n, 256, 0, st>>>(bv, ctx->tp, ctx->rp, ctx->bt_run, w.virt);
            }
            {
                KTimer t(ctx, KC_WIRES);
                k_wire_bfs<<<std::min(n, w.n_cta), WIRE_THREADS, WIRE_SMEM_WORDS * 4, st>>>(ctx->bt, c, w.d_edges, w.d_cnt, (int)w.edge_cap, frame0 + done, n);
            }
            CK(cudaGetLastError());
        }
    }
    cudaEventRecord(t1, st);
    ctx->pending.push_back({t0, t1, -1});
    ctx->tm.frames += nf;
    return HB_OK;
}

// ---- batch bookkeeping -------------------------------------------------------------------------------
// A batch stays "in flight" (its coordinates must stay intact) until a status snapshot taken right
// after it on the compute stream shows that the bond table did not overflow while it ran.  On overflow
// the table is grown and every in-flight batch is re-run (idempotent: bits are OR-ed, keys deduplicate).
static int full_verify(hb_ctx* ctx) {
    for (int attempt = 0; attempt < 28; ++attempt) {
        int st[8];        // n_keys, overflow, fused_overflow, pair_overflow, n_hits (64 bit), wire_overflow, pad
        CK(cudaMemcpyAsync(ctx->h_verify, ctx->bt.n_keys, sizeof(st), cudaMemcpyDeviceToHost, ctx->s_compute));
        CK(cudaStreamSynchronize(ctx->s_compute));
        memcpy(st, ctx->h_verify, sizeof(st));
        memcpy(ctx->last_status, st, 6 * sizeof(int));
        bool overflow = st[1] != 0, fused_overflow = st[2] != 0, pair_overflow = st[3] != 0, wire_overflow = st[6] != 0;
        const bool fp_overflow = st[7] != 0;
        if (!overflow && !fused_overflow && !pair_overflow && !wire_overflow && !fp_overflow && (long long)st[0] * 2 <= ctx->table_cap) {
            for (auto& sl : ctx->slots) sl.busy = false;
            return HB_OK;
        }
        int rc;
        if (overflow || (long long)st[0] * 2 > ctx->table_cap) {
            if ((rc = grow_table(ctx))) return rc;
            while ((long long)st[0] * 2 > ctx->table_cap)
                if ((rc = grow_table(ctx))) return rc;
        }
        if (fused_overflow) {      // some frame did not fit in shared memory: retry with doubled capacities
            ctx->fused_fallbacks++;   // while they fit, else this run continues on the general path
            if (ctx->opt_fused_lw_cap || ctx->opt_fused_pair_cap || ctx->fused_grow >= 6 || !plan_fused(ctx, ctx->fused_grow + 1)) {
                ctx->fused_active = false;
                if (!ctx->opt_fused_lw_cap && !ctx->opt_fused_pair_cap && ctx->rp.method == HB_METHOD_WATER_AROUND)
                    ctx->fused_unfit_around = ctx->prm.around_radius;
            } else {
                if (ctx->rp.method == HB_METHOD_WATER_AROUND) {
                    ctx->fused_grow_hint = ctx->fused_grow;
                    ctx->fused_hint_around = ctx->prm.around_radius;
                }
            }
        }
        if (pair_overflow) {       // the sub-batch found more pairs than the pair buffer holds: halve it
            if (ctx->batch_frames > 1) ctx->batch_frames = std::max(1, ctx->batch_frames / 2);
            else ctx->use_pairbuf = false;
        }
        if (wire_overflow) {       // some frame has more H-bond pairs than its edge list holds: double the lists
            CK(cudaStreamSynchronize(ctx->s_compute));
            free_list(ctx->wire.allocs);
            ctx->wire.d_edges = nullptr;
            ctx->wire.edge_cap *= 2;
            CK(cudaMemsetAsync(ctx->bt.n_keys + 6, 0, sizeof(int), ctx->s_compute));
        }
        if (fp_overflow) {         // a frame did not fit k_frame_pairs: back to the multi-kernel pair search
            ctx->fp_level++;
            CK(cudaMemsetAsync(ctx->bt.n_keys + 7, 0, sizeof(int), ctx->s_compute));
        }
        if (overflow || fused_overflow || pair_overflow || wire_overflow || fp_overflow) {
            CK(cudaMemsetAsync(ctx->bt.overflow, 0, 3 * sizeof(int), ctx->s_compute));
            for (auto& sl : ctx->slots)
                if (sl.busy) {
                    long long frames_before = ctx->tm.frames;
                    if (sl.fused) ctx->tm.fused_frames -= sl.nf;    // a re-run is not new work
                    if ((rc = run_batch(ctx, sl.d_xyz, sl.f0, sl.nf, sl.apf, sl.bytes, sl.d_box))) return rc;
                    sl.fused = ctx->last_batch_fused;
                    ctx->tm.frames = frames_before;
                }
        }
    }
    return fail(ctx, HB_ERR_NOMEM, "bond table kept overflowing");
}

// make slot k reusable: its previous batch must be complete and verified
static int acquire_slot(hb_ctx* ctx, int k) {
    hb_ctx::Slot& sl = ctx->slots[k];
    if (!sl.busy) return HB_OK;
    CK(cudaEventSynchronize(ctx->ev_done[k]));
    int nkeys = ctx->h_status[8 * k], overflow = ctx->h_status[8 * k + 1] | ctx->h_status[8 * k + 2] | ctx->h_status[8 * k + 3] | ctx->h_status[8 * k + 6] | ctx->h_status[8 * k + 7];
    if (overflow || (long long)nkeys * 2 > ctx->table_cap) return full_verify(ctx);
    sl.busy = false;
    return HB_OK;
}

static int launch_in_slot(hb_ctx* ctx, int k, const float* d_xyz, long long f0, int nf, long long apf, size_t bytes,
                          const float* d_box) {
    int rc = run_batch(ctx, d_xyz, f0, nf, apf, bytes, d_box);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_status + 8 * k, ctx->bt.n_keys, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->s_compute));
    CK(cudaEventRecord(ctx->ev_done[k], ctx->s_compute));
    ctx->slots[k] = hb_ctx::Slot{true, d_xyz, f0, nf, apf, bytes, d_box, ctx->last_batch_fused};
    return HB_OK;
}

static int check_submit(hb_ctx* ctx, int64_t n_atoms, int64_t n_frames) {
    if (!ctx->running) return fail(ctx, HB_ERR_STATE, "hb_submit_frames before hb_begin");
    if (n_frames < 0 || ctx->next_frame + n_frames > ctx->n_frames_total) return fail(ctx, HB_ERR_ARG, "more frames submitted than announced in hb_begin");
    if (!ctx->prm.structure_batch && n_atoms != ctx->h_sys_atom_off[1] - ctx->h_sys_atom_off[0])
        return fail(ctx, HB_ERR_ARG, "n_atoms does not match the topology");
    return HB_OK;
}

// lattice vectors of host-supplied frames: lower-triangular rows a, b, c with a positive diagonal, and cut-offs
// below half of every brick edge so that the sequential minimum-image reduction finds every pair
static int check_boxes(hb_ctx* ctx, const float* box, long long n_frames) {
    const double rmax = std::max(ctx->prm.distance, (ctx->prm.method == HB_METHOD_WATER_AROUND || ctx->presence) ? ctx->prm.around_radius : 0.0);
    for (long long f = 0; f < n_frames; ++f) {
        const float* b = box + 9 * f;
        if (b[1] != 0.f || b[2] != 0.f || b[5] != 0.f)
            return fail(ctx, HB_ERR_ARG, "box: lattice vectors must be lower-triangular rows (a = (ax,0,0), b = (bx,by,0), c = (cx,cy,cz))");
        if (!(b[0] > 0.f) || !(b[4] > 0.f) || !(b[8] > 0.f) || !std::isfinite(b[0] + b[3] + b[4] + b[6] + b[7] + b[8]))
            return fail(ctx, HB_ERR_ARG, "box: non-positive or non-finite cell edge");
        if (ctx->prm.pbc_mode == HB_PBC_ORTHO && (b[3] != 0.f || b[6] != 0.f || b[7] != 0.f))
            return fail(ctx, HB_ERR_ARG, "box: HB_PBC_ORTHO needs an orthorhombic cell");
        if (!(2.0 * rmax < std::min(b[0], std::min(b[4], b[8]))))
            return fail(ctx, HB_ERR_ARG, "box: cut-off radius must stay below half of the smallest cell edge");
    }
    return HB_OK;
}

static long long batch_atoms(hb_ctx* ctx, long long f0, int nf, long long n_atoms) {
    if (ctx->prm.structure_batch) return ctx->h_sys_atom_off[f0 + nf] - ctx->h_sys_atom_off[f0];
    return n_atoms * nf;
}

static int next_batch_frames(hb_ctx* ctx, long long f0, long long remaining, bool host) {
    long long lim = host ? ctx->host_batch_frames : (ctx->fused_active ? ctx->opt_fused_batch_frames : ctx->batch_frames);
    int nf = (int)std::min<long long>(remaining, std::max<long long>(lim, 1));
    if (host && !ctx->prm.structure_batch) {
        long long fit = (long long)(ctx->in_alloc / std::max<size_t>((size_t)ctx->max_atoms * 12, 1));
        nf = (int)std::max<long long>(1, std::min<long long>(nf, fit));
    }
    if (host && ctx->prm.structure_batch) {   // ragged: stay within the input buffer
        long long budget = (long long)(ctx->in_alloc / 12);
        while (nf > 1 && ctx->h_sys_atom_off[f0 + nf] - ctx->h_sys_atom_off[f0] > budget) --nf;
    }
    return nf;
}

static int ensure_box_buffers(hb_ctx* ctx) {
    if (ctx->box_cap >= (size_t)std::max(ctx->host_batch_frames, 1)) return HB_OK;
    size_t cap = (size_t)std::max(ctx->host_batch_frames, 1);
    for (int k = 0; k < 2; ++k) {
        if (ctx->d_box[k]) cudaFree(ctx->d_box[k]);
        if (ctx->h_box[k]) cudaFreeHost(ctx->h_box[k]);
        ctx->d_box[k] = nullptr; ctx->h_box[k] = nullptr;
    }
    ctx->box_cap = 0;
    for (int k = 0; k < 2; ++k) {
        CK(cudaMalloc((void**)&ctx->d_box[k], cap * 36));
        CK(cudaHostAlloc((void**)&ctx->h_box[k], cap * 36, cudaHostAllocDefault));
    }
    ctx->box_cap = cap;
    return HB_OK;
}

// the two device input buffers (frames as [atom][3]), the two pinned staging buffers for pageable sources and,
// for planar sources, the two raw device buffers: as large as one batch of this call needs, at most the plan
static int ensure_input_buffers(hb_ctx* ctx, size_t call_bytes, bool pinned, size_t raw_bytes_per_aos_byte_num,
                                size_t raw_bytes_per_aos_byte_den) {
    size_t need = std::max<size_t>(std::min(ctx->in_bytes, call_bytes), 16);
    const bool planar = raw_bytes_per_aos_byte_num != 0;
    auto raw_of = [&](size_t aos) { return planar ? aos / raw_bytes_per_aos_byte_den * raw_bytes_per_aos_byte_num + 256 : (size_t)0; };
    const bool stage_small = !pinned && (!ctx->h_stage[0] || ctx->stage_alloc < (planar ? raw_of(std::max(need, ctx->in_alloc)) : std::max(need, ctx->in_alloc)));
    const bool raw_small = planar && (!ctx->d_raw[0] || ctx->raw_alloc < raw_of(std::max(need, ctx->in_alloc)));
    if (need <= ctx->in_alloc && !stage_small && !raw_small) return HB_OK;
    CK(cudaStreamSynchronize(ctx->s_copy));
    int rcv = full_verify(ctx);               // nothing may still read the old buffers
    if (rcv) return rcv;
    need = std::max(need, ctx->in_alloc);
    for (int k = 0; k < 2; ++k) {
        if (need > ctx->in_alloc || !ctx->d_in[k]) {
            if (ctx->d_in[k]) cudaFree(ctx->d_in[k]);
            ctx->d_in[k] = nullptr;
            CK(cudaMalloc((void**)&ctx->d_in[k], need));
        }
        if (raw_small) {
            if (ctx->d_raw[k]) cudaFree(ctx->d_raw[k]);
            ctx->d_raw[k] = nullptr;
            CK(cudaMalloc((void**)&ctx->d_raw[k], raw_of(need)));
        }
        if (stage_small) {
            if (ctx->h_stage[k]) cudaFreeHost(ctx->h_stage[k]);
            ctx->h_stage[k] = nullptr;
            CK(cudaHostAlloc((void**)&ctx->h_stage[k], planar ? raw_of(need) : need, cudaHostAllocDefault));
        }
    }
    if (raw_small) ctx->raw_alloc = raw_of(need);
    if (stage_small) ctx->stage_alloc = planar ? raw_of(need) : need;
    ctx->in_alloc = need;
    return HB_OK;
}

int hb_submit_frames(hb_ctx* ctx, const float* xyz, int64_t n_atoms, int64_t n_frames, const float* box) {
    if (!ctx || (!xyz && n_frames > 0)) return HB_ERR_ARG;
    int rc = check_submit(ctx, n_atoms, n_frames);
    if (rc) return rc;
    const bool pbc = ctx->prm.pbc_mode != HB_PBC_NONE;
    if (pbc) {
        if (!box) return fail(ctx, HB_ERR_ARG, "hb_submit_frames: pbc_mode needs the lattice vectors of every frame");
        if ((rc = check_boxes(ctx, box, n_frames))) return rc;
    }
    CK(cudaSetDevice(ctx->device));
    if (pbc && (rc = ensure_box_buffers(ctx))) return rc;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, xyz) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    {
        size_t call_bytes = (size_t)batch_atoms(ctx, ctx->next_frame, (int)std::min<long long>(n_frames, 1 << 30), n_atoms) * 12;
        if ((rc = ensure_input_buffers(ctx, call_bytes, pinned, 0, 1))) return rc;
    }
    const float* src = xyz;
    long long done = 0;
    while (done < n_frames) {
        long long f0 = ctx->next_frame;
        int nf = next_batch_frames(ctx, f0, n_frames - done, true);
        long long atoms = batch_atoms(ctx, f0, nf, n_atoms);
        size_t bytes = (size_t)atoms * 12;
        int k = (int)(ctx->batch_counter & 1);
        if ((rc = acquire_slot(ctx, k))) return rc;       // buffer k (device + staging) is free again
        const float* h = src;
        const bool ranged = ctx->opt_copy_ranges && !ctx->prm.structure_batch && !ctx->copy_ranges.empty();
        if (!pinned) {
            float* stg = ctx->h_stage[k];
            if (ranged) {
                const auto& ranges = ctx->copy_ranges;
                parallel_frames(nf, (size_t)n_atoms * 12, ctx->opt_stage_threads, [=, &ranges](int fa, int fb2) {
                    for (int f = fa; f < fb2; ++f)
                        for (auto& r : ranges)
                            memcpy(stg + ((size_t)f * n_atoms + r.first) * 3, src + ((size_t)f * n_atoms + r.first) * 3,
                                   (size_t)(r.second - r.first) * 12);
                });
            } else if (ctx->prm.structure_batch) {
                memcpy(stg, src, bytes);
            } else {
                parallel_frames(nf, (size_t)n_atoms * 12, ctx->opt_stage_threads, [=](int fa, int fb2) {
                    memcpy(stg + (size_t)fa * n_atoms * 3, src + (size_t)fa * n_atoms * 3, (size_t)(fb2 - fa) * n_atoms * 12);
                });
            }
            h = stg;
        }
        if (pbc) {
            memcpy(ctx->h_box[k], box + 9 * done, (size_t)nf * 36);
            CK(cudaMemcpyAsync(ctx->d_box[k], ctx->h_box[k], (size_t)nf * 36, cudaMemcpyHostToDevice, ctx->s_copy));
        }
        cudaEvent_t c0 = get_event(ctx), c1 = get_event(ctx);
        cudaEventRecord(c0, ctx->s_copy);
        if (ranged) {      // strided copies: only the atom ranges the search reads, same layout on the device
            const size_t pitch = (size_t)n_atoms * 12;
            for (auto& r : ctx->copy_ranges) {
                CK(cudaMemcpy2DAsync((char*)ctx->d_in[k] + (size_t)r.first * 12, pitch, (const char*)h + (size_t)r.first * 12, pitch,
                                     (size_t)(r.second - r.first) * 12, (size_t)nf, cudaMemcpyHostToDevice, ctx->s_copy));
                ctx->tm.h2d_bytes += (long long)(r.second - r.first) * 12 * nf;
            }
        } else {
            CK(cudaMemcpyAsync(ctx->d_in[k], h, bytes, cudaMemcpyHostToDevice, ctx->s_copy));
            ctx->tm.h2d_bytes += (long long)bytes;
        }
        cudaEventRecord(c1, ctx->s_copy);
        ctx->pending.push_back({c0, c1, -2});
        CK(cudaEventRecord(ctx->ev_copied[k], ctx->s_copy));
        CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_copied[k], 0));
        if ((rc = launch_in_slot(ctx, k, ctx->d_in[k], f0, nf, n_atoms, bytes, pbc ? ctx->d_box[k] : nullptr))) return rc;
        ctx->batch_counter++;
        ctx->next_frame += nf;
        done += nf;
        src += (size_t)atoms * 3;
        if (ctx->pending.size() > 8192) drain_events(ctx);
    }
    if (pinned) CK(cudaStreamSynchronize(ctx->s_copy));   // caller may reuse its pinned buffer on return
    return HB_OK;
}

int hb_submit_frames_planar(hb_ctx* ctx, const void* raw, int64_t frame_stride, int64_t x_off, int64_t y_off, int64_t z_off,
                            int64_t n_atoms, int64_t n_frames, const float* box) {
    if (!ctx || (!raw && n_frames > 0)) return HB_ERR_ARG;
    int rc = check_submit(ctx, n_atoms, n_frames);
    if (rc) return rc;
    if (ctx->prm.structure_batch) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_planar: trajectories only");
    const int64_t plane = n_atoms * 4;
    if ((x_off | y_off | z_off | frame_stride) & 3 || x_off < 0 || y_off < 0 || z_off < 0 || x_off + plane > frame_stride ||
        y_off + plane > frame_stride || z_off + plane > frame_stride)
        return fail(ctx, HB_ERR_ARG, "hb_submit_frames_planar: plane offsets must be 4-byte aligned and inside the frame record");
    const bool pbc = ctx->prm.pbc_mode != HB_PBC_NONE;
    if (pbc) {
        if (!box) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_planar: pbc_mode needs the lattice vectors of every frame");
        if ((rc = check_boxes(ctx, box, n_frames))) return rc;
    }
    CK(cudaSetDevice(ctx->device));
    if (pbc && (rc = ensure_box_buffers(ctx))) return rc;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, raw) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if ((rc = ensure_input_buffers(ctx, (size_t)n_atoms * 12 * (size_t)n_frames, pinned, (size_t)frame_stride, (size_t)n_atoms * 12))) return rc;
    const unsigned char* src = (const unsigned char*)raw;
    const int64_t offs[3] = {x_off, y_off, z_off};
    std::vector<std::pair<long long, long long>> whole{{0, n_atoms}};
    const auto& ranges = (ctx->opt_copy_ranges && !ctx->copy_ranges.empty()) ? ctx->copy_ranges : whole;
    long long done = 0;
    while (done < n_frames) {
        long long f0 = ctx->next_frame;
        int nf = next_batch_frames(ctx, f0, n_frames - done, true);
        nf = (int)std::min<long long>(nf, (long long)(ctx->raw_alloc / (size_t)frame_stride));
        int k = (int)(ctx->batch_counter & 1);
        if ((rc = acquire_slot(ctx, k))) return rc;
        const unsigned char* h = src;
        if (!pinned) {                                   // only the plane pieces that are copied on are staged
            unsigned char* st = (unsigned char*)ctx->h_stage[k];
            parallel_frames(nf, (size_t)frame_stride, ctx->opt_stage_threads, [=, &ranges](int fa, int fb2) {
                for (int f = fa; f < fb2; ++f)
                    for (int p = 0; p < 3; ++p)
                        for (auto& r : ranges)
                            memcpy(st + (size_t)f * frame_stride + offs[p] + r.first * 4,
                                   src + (size_t)f * frame_stride + offs[p] + r.first * 4, (size_t)(r.second - r.first) * 4);
            });
            h = st;
        }
        if (pbc) {
            memcpy(ctx->h_box[k], box + 9 * done, (size_t)nf * 36);
            CK(cudaMemcpyAsync(ctx->d_box[k], ctx->h_box[k], (size_t)nf * 36, cudaMemcpyHostToDevice, ctx->s_copy));
        }
        cudaEvent_t c0 = get_event(ctx), c1 = get_event(ctx);
        cudaEventRecord(c0, ctx->s_copy);
        for (int p = 0; p < 3; ++p)
            for (auto& r : ranges) {
                CK(cudaMemcpy2DAsync(ctx->d_raw[k] + offs[p] + r.first * 4, (size_t)frame_stride, h + offs[p] + r.first * 4, (size_t)frame_stride,
                                     (size_t)(r.second - r.first) * 4, (size_t)nf, cudaMemcpyHostToDevice, ctx->s_copy));
                ctx->tm.h2d_bytes += (long long)(r.second - r.first) * 4 * nf;
            }
        for (auto& r : ranges) {
            const long long len = r.second - r.first;
            dim3 grid((unsigned)((len + 255) / 256), (unsigned)nf);
            k_interleave<<<grid, 256, 0, ctx->s_copy>>>(ctx->d_raw[k], frame_stride, x_off, y_off, z_off, r.first, len, n_atoms, ctx->d_in[k]);
            ctx->tm.n_launches += 1;
        }
        CK(cudaGetLastError());
        cudaEventRecord(c1, ctx->s_copy);
        ctx->pending.push_back({c0, c1, -2});
        CK(cudaEventRecord(ctx->ev_copied[k], ctx->s_copy));
        CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_copied[k], 0));
        if ((rc = launch_in_slot(ctx, k, ctx->d_in[k], f0, nf, n_atoms, (size_t)nf * n_atoms * 12, pbc ? ctx->d_box[k] : nullptr))) return rc;
        ctx->batch_counter++;
        ctx->next_frame += nf;
        done += nf;
        src += (size_t)nf * frame_stride;
        if (ctx->pending.size() > 8192) drain_events(ctx);
    }
    if (pinned) CK(cudaStreamSynchronize(ctx->s_copy));
    return HB_OK;
}

// ---- XTC sources -----------------------------------------------------------------------------------------
static int check_xtc_status(hb_ctx* ctx) {        // after the copy stream has drained
    if (ctx->h_xtc_status && *ctx->h_xtc_status) {
        *ctx->h_xtc_status = 0;
        if (ctx->d_xtc_status) cudaMemsetAsync(ctx->d_xtc_status, 0, sizeof(int), ctx->s_compute);
        return fail(ctx, HB_ERR_ARG, "hb_submit_frames_xtc: malformed XTC frame record (magic, atom count or bit stream)");
    }
    return HB_OK;
}

int hb_xtc_index(const void* buf, int64_t nbytes, int64_t* frame_off, int64_t max_frames, int64_t* n_frames, int64_t* n_atoms) {
    if (!buf || nbytes < 0 || !n_frames) return HB_ERR_ARG;
    const unsigned char* p = (const unsigned char*)buf;
    int64_t off = 0, n = 0;
    int atoms0 = -1;
    while (off < nbytes) {
        int na = 0;
        const size_t len = xtc::frame_bytes(p + off, (size_t)(nbytes - off), &na);
        if (!len) break;                               // truncated tail or foreign bytes: the index ends here
        if (atoms0 < 0) atoms0 = na;
        if (na != atoms0) return HB_ERR_ARG;
        if (frame_off && n < max_frames) frame_off[n] = off;
        ++n;
        off += (int64_t)len;
    }
    if (frame_off && n <= max_frames) frame_off[n] = off;
    *n_frames = n;
    if (n_atoms) *n_atoms = atoms0 < 0 ? 0 : atoms0;
    return HB_OK;
}

int hb_xtc_encode(const float* xyz, int64_t n_atoms, int64_t n_frames, float in_scale, float precision, const float* box9,
                  int64_t first_step, float dt, void* out, int64_t out_cap, int64_t* frame_off, int64_t* bytes_written,
                  int32_t n_threads) {
    if (!xyz || n_atoms < 1 || n_atoms > INT32_MAX / 3 || n_frames < 0 || !bytes_written) return HB_ERR_ARG;
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n_threads > 0 ? n_threads : 1, 64), n_frames));
    std::vector<std::vector<unsigned char>> parts((size_t)nt);
    std::vector<std::vector<int64_t>> lens((size_t)nt);
    std::vector<int> ok((size_t)nt, 1);
    auto work = [&](int t) {
        const int64_t a = n_frames * t / nt, b = n_frames * (t + 1) / nt;
        for (int64_t f = a; f < b; ++f) {
            const size_t before = parts[t].size();
            if (!xtc::encode_frame(xyz + (size_t)f * n_atoms * 3, (int)n_atoms, in_scale, precision, (int)(first_step + f),
                                   dt * (float)(first_step + f), box9 ? box9 + 9 * f : nullptr, parts[t])) { ok[t] = 0; return; }
            lens[t].push_back((int64_t)(parts[t].size() - before));
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    int64_t total = 0, f = 0;
    for (int t = 0; t < nt; ++t) { if (!ok[t]) return HB_ERR_ARG; total += (int64_t)parts[t].size(); }
    *bytes_written = total;
    if (!out || total > out_cap) return out ? HB_ERR_NOMEM : HB_OK;     // out == NULL: size query
    int64_t off = 0;
    for (int t = 0; t < nt; ++t) {
        memcpy((unsigned char*)out + off, parts[t].data(), parts[t].size());
        if (frame_off) for (int64_t l : lens[t]) { frame_off[f++] = off; off += l; }
        else off += (int64_t)parts[t].size();
    }
    if (frame_off) frame_off[n_frames] = total;
    return HB_OK;
}

int hb_xtc_decode_host(const void* rec, int64_t avail, int64_t n_atoms, float scale, float* xyz, float* box9, int64_t* rec_bytes) {
    if (!rec || !xyz || avail < 0 || n_atoms < 0 || n_atoms > INT32_MAX / 3) return HB_ERR_ARG;
    const size_t len = xtc::decode_frame((const unsigned char*)rec, (size_t)avail, (int)n_atoms, scale, xyz, box9, nullptr, nullptr);
    if (!len) return HB_ERR_ARG;
    if (rec_bytes) *rec_bytes = (int64_t)len;
    return HB_OK;
}

int hb_xtc_decode_to_device(hb_ctx* ctx, const void* raw, const int64_t* frame_off, int64_t n_atoms, int64_t n_frames,
                            float scale, float* d_out) {
    if (!ctx || !raw || !frame_off || !d_out || n_frames < 0 || n_atoms < 1 || n_atoms > INT32_MAX / 3) return HB_ERR_ARG;
    if (n_frames == 0) return HB_OK;
    for (int64_t f = 0; f < n_frames; ++f)
        if (frame_off[f + 1] <= frame_off[f] || (frame_off[f] & 3)) return fail(ctx, HB_ERR_ARG, "hb_xtc_decode_to_device: frame offsets must ascend in multiples of 4");
    CK(cudaSetDevice(ctx->device));
    const int64_t b0 = frame_off[0];
    const size_t bytes = (size_t)(frame_off[n_frames] - b0);
    unsigned char* d_raw = nullptr;
    long long* d_off = nullptr;
    int* d_status = nullptr;
    std::vector<long long> rel((size_t)n_frames + 1);
    for (int64_t f = 0; f <= n_frames; ++f) rel[f] = frame_off[f] - b0;
    int status = 0;
    cudaError_t e = cudaMalloc((void**)&d_raw, bytes + 16384);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_off, rel.size() * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_status, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_status, 0, sizeof(int), ctx->s_copy);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_raw + bytes, 0, 16384, ctx->s_copy);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_raw, (const unsigned char*)raw + b0, bytes, cudaMemcpyHostToDevice, ctx->s_copy);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, ctx->s_copy);
    if (e == cudaSuccess) {
        XtcArgs xa{};
        xa.raw = d_raw; xa.frame_off = d_off; xa.n_frames = (int)n_frames; xa.n_atoms = (int)n_atoms; xa.scale = scale;
        xa.out = d_out; xa.status = d_status;
        k_xtc_decode<<<(unsigned)n_frames, XTC_WARPS * 32, 0, ctx->s_copy>>>(xa);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_copy);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_copy);
    if (d_raw) cudaFree(d_raw);
    if (d_off) cudaFree(d_off);
    if (d_status) cudaFree(d_status);
    if (e != cudaSuccess) return fail(ctx, HB_ERR_CUDA, cudaGetErrorString(e));
    if (status) return fail(ctx, HB_ERR_ARG, "hb_xtc_decode_to_device: malformed XTC frame record (magic, atom count or bit stream)");
    return HB_OK;
}

int hb_submit_frames_xtc(hb_ctx* ctx, const void* raw, const int64_t* frame_off, int64_t n_atoms, int64_t n_frames, float scale,
                         const float* box) {
    if (!ctx || ((!raw || !frame_off) && n_frames > 0)) return HB_ERR_ARG;
    int rc = check_submit(ctx, n_atoms, n_frames);
    if (rc) return rc;
    if (ctx->prm.structure_batch) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_xtc: trajectories only");
    if (n_atoms > INT32_MAX / 3) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_xtc: too many atoms");
    size_t max_rec = 0;
    for (int64_t f = 0; f < n_frames; ++f) {
        if (frame_off[f + 1] <= frame_off[f] || (frame_off[f] & 3)) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_xtc: frame offsets must ascend in multiples of 4");
        max_rec = std::max(max_rec, (size_t)(frame_off[f + 1] - frame_off[f]));
    }
    const bool pbc = ctx->prm.pbc_mode != HB_PBC_NONE;
    if (pbc) {
        if (!box) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_xtc: pbc_mode needs the lattice vectors of every frame");
        if ((rc = check_boxes(ctx, box, n_frames))) return rc;
    }
    CK(cudaSetDevice(ctx->device));
    if (pbc && (rc = ensure_box_buffers(ctx))) return rc;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, raw) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    // raw buffers: half the bytes of the decoded frames (a record takes 3 - 5 bytes per atom), at least one record
    if ((rc = ensure_input_buffers(ctx, std::max((size_t)n_atoms * 12 * (size_t)n_frames, 2 * max_rec + 65536), pinned, 1, 2))) return rc;
    if (max_rec + 16384 > ctx->raw_alloc) return fail(ctx, HB_ERR_NOMEM, "hb_submit_frames_xtc: a frame record does not fit the input buffer (raise batch_bytes)");
    if (!ctx->d_xtc_status) {
        CK(cudaMalloc((void**)&ctx->d_xtc_status, sizeof(int)));
        CK(cudaMemset(ctx->d_xtc_status, 0, sizeof(int)));
        CK(cudaHostAlloc((void**)&ctx->h_xtc_status, sizeof(int), cudaHostAllocDefault));
        *ctx->h_xtc_status = 0;
    }
    const size_t want_off = (size_t)std::max(ctx->host_batch_frames, 1) + 1;
    if (ctx->xoff_cap < want_off) {
        CK(cudaStreamSynchronize(ctx->s_copy));
        for (int k = 0; k < 2; ++k) {
            if (ctx->d_xoff[k]) cudaFree(ctx->d_xoff[k]);
            if (ctx->h_xoff[k]) cudaFreeHost(ctx->h_xoff[k]);
            ctx->d_xoff[k] = nullptr; ctx->h_xoff[k] = nullptr;
        }
        ctx->xoff_cap = 0;
        for (int k = 0; k < 2; ++k) {
            CK(cudaMalloc((void**)&ctx->d_xoff[k], want_off * 8));
            CK(cudaHostAlloc((void**)&ctx->h_xoff[k], want_off * 8, cudaHostAllocDefault));
        }
        ctx->xoff_cap = want_off;
    }
    XtcRanges xr{};
    if (ctx->opt_copy_ranges && !ctx->copy_ranges.empty() && ctx->copy_ranges.size() <= 8) {
        xr.n = (int)ctx->copy_ranges.size();
        for (int k = 0; k < xr.n; ++k) { xr.b[k] = (int)ctx->copy_ranges[k].first; xr.e[k] = (int)ctx->copy_ranges[k].second; }
    }
    const unsigned char* src = (const unsigned char*)raw;
    long long done = 0;
    while (done < n_frames) {
        long long f0 = ctx->next_frame;
        int nf = next_batch_frames(ctx, f0, n_frames - done, true);
        nf = (int)std::min<long long>(nf, (long long)ctx->xoff_cap - 1);
        while (nf > 1 && (size_t)(frame_off[done + nf] - frame_off[done]) + 16384 > ctx->raw_alloc) --nf;   // (the decoder reads windows ahead)
        const int64_t b0 = frame_off[done];
        const size_t bytes = (size_t)(frame_off[done + nf] - b0);
        int k = (int)(ctx->batch_counter & 1);
        if ((rc = acquire_slot(ctx, k))) return rc;
        if ((rc = check_xtc_status(ctx))) return rc;
        const unsigned char* h = src + b0;
        if (!pinned) {
            unsigned char* st = (unsigned char*)ctx->h_stage[k];
            const int pieces = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->opt_stage_threads, bytes >> 23));
            parallel_frames(pieces, bytes / (size_t)pieces + 1, ctx->opt_stage_threads, [=](int a, int b2) {
                const size_t lo = bytes * (size_t)a / (size_t)pieces, hi = bytes * (size_t)b2 / (size_t)pieces;
                memcpy(st + lo, h + lo, hi - lo);
            });
            h = st;
        }
        for (int f = 0; f <= nf; ++f) ctx->h_xoff[k][f] = frame_off[done + f] - b0;
        if (pbc) {
            memcpy(ctx->h_box[k], box + 9 * done, (size_t)nf * 36);
            CK(cudaMemcpyAsync(ctx->d_box[k], ctx->h_box[k], (size_t)nf * 36, cudaMemcpyHostToDevice, ctx->s_copy));
        }
        cudaEvent_t c0 = get_event(ctx), c1 = get_event(ctx);
        cudaEventRecord(c0, ctx->s_copy);
        CK(cudaMemcpyAsync(ctx->d_xoff[k], ctx->h_xoff[k], (size_t)(nf + 1) * 8, cudaMemcpyHostToDevice, ctx->s_copy));
        CK(cudaMemcpyAsync(ctx->d_raw[k], h, bytes, cudaMemcpyHostToDevice, ctx->s_copy));
        ctx->tm.h2d_bytes += (long long)bytes;
        XtcArgs xa{};
        xa.raw = ctx->d_raw[k];
        xa.frame_off = ctx->d_xoff[k];
        xa.n_frames = nf;
        xa.n_atoms = (int)n_atoms;
        xa.scale = scale;
        xa.out = ctx->d_in[k];
        xa.status = ctx->d_xtc_status;
        xa.ranges = xr;
        cudaEventRecord(c1, ctx->s_copy);
        ctx->pending.push_back({c0, c1, -2});
        CK(cudaEventRecord(ctx->ev_copied[k], ctx->s_copy));
        CK(cudaStreamWaitEvent(ctx->s_compute, ctx->ev_copied[k], 0));
        // the decode runs on the compute stream, in front of the search of its batch: the copy stream goes straight on
        // with the records of the next batch
        {
            KTimer t(ctx, KC_XTC);
            k_xtc_decode<<<nf, XTC_WARPS * 32, 0, ctx->s_compute>>>(xa);
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_xtc_status, ctx->d_xtc_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_compute));
        if ((rc = launch_in_slot(ctx, k, ctx->d_in[k], f0, nf, n_atoms, (size_t)nf * n_atoms * 12, pbc ? ctx->d_box[k] : nullptr))) return rc;
        ctx->batch_counter++;
        ctx->next_frame += nf;
        done += nf;
        if (ctx->pending.size() > 8192) drain_events(ctx);
    }
    if (pinned) CK(cudaStreamSynchronize(ctx->s_copy));
    return HB_OK;
}

int hb_submit_frames_device(hb_ctx* ctx, const float* d_xyz, int64_t n_atoms, int64_t n_frames, const float* d_box) {
    if (!ctx || (!d_xyz && n_frames > 0)) return HB_ERR_ARG;
    int rc = check_submit(ctx, n_atoms, n_frames);
    if (rc) return rc;
    const bool pbc = ctx->prm.pbc_mode != HB_PBC_NONE;
    if (pbc && !d_box) return fail(ctx, HB_ERR_ARG, "hb_submit_frames_device: pbc_mode needs the lattice vectors of every frame");
    CK(cudaSetDevice(ctx->device));
    const float* src = d_xyz;
    long long done = 0;
    while (done < n_frames) {
        long long f0 = ctx->next_frame;
        int nf = next_batch_frames(ctx, f0, n_frames - done, false);
        long long atoms = batch_atoms(ctx, f0, nf, n_atoms);
        int k = (int)(ctx->batch_counter & 1);
        if ((rc = acquire_slot(ctx, k))) return rc;
        if ((rc = launch_in_slot(ctx, k, src, f0, nf, n_atoms, (size_t)atoms * 12, pbc ? d_box + 9 * done : nullptr))) return rc;
        ctx->batch_counter++;
        ctx->next_frame += nf;
        done += nf;
        src += (size_t)atoms * 3;
        if (ctx->pending.size() > 8192) drain_events(ctx);
    }
    // the caller's buffers must stay valid until the batches are verified, which hb_sync / hb_finalize do
    return HB_OK;
}

int hb_sync(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_copy));
    if (ctx->running) {
        int rc = full_verify(ctx);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(ctx->s_compute));
    drain_events(ctx);
    return check_xtc_status(ctx);
}

// Queue compaction, key sort and row gather behind the search. Small tables (the common case) are handled
// speculatively: nothing is known about the bond count on the host, the kernels read it from the device
// counter, and one status read back at the end verifies the run and delivers the count. Large tables first
// settle (hb_sync) to size the radix passes. Leaves the sorted keys / slots in ctx->fin_sorted_*.
static int queue_finalize(hb_ctx* ctx, bool speculative, int* n_known) {
    cudaStream_t st = ctx->s_compute;
    const int wpr = ctx->bt.wpr;
    int rc, n = 0;
    long long n_max = speculative ? ctx->table_cap / 2 : 0;
    if (!speculative) {
        if ((rc = hb_sync(ctx))) return rc;
        n = ctx->last_status[0];
        n_max = std::max(n, 1);
    }
    const int nblk = (int)((n_max + RS_TILE - 1) / RS_TILE);
    if ((size_t)n_max > ctx->fin_cap) {           // (re)allocate the sort buffers with headroom
        size_t cap2 = std::max<size_t>((size_t)n_max * 2, 4096);
        for (int k = 0; k < 2; ++k) {
            if (ctx->fin_k[k]) cudaFree(ctx->fin_k[k]);
            if (ctx->fin_v[k]) cudaFree(ctx->fin_v[k]);
            ctx->fin_k[k] = nullptr; ctx->fin_v[k] = nullptr;
        }
        ctx->fin_cap = 0;
        for (int k = 0; k < 2; ++k) {
            CK(cudaMalloc((void**)&ctx->fin_k[k], cap2 * 8));
            CK(cudaMalloc((void**)&ctx->fin_v[k], cap2 * 4));
        }
        ctx->fin_cap = cap2;
    }
    size_t hist_need = (size_t)256 * std::max(nblk, 1) * 4;
    if (hist_need > ctx->fin_hist_cap) {
        if (ctx->fin_hist) cudaFree(ctx->fin_hist);
        ctx->fin_hist = nullptr; ctx->fin_hist_cap = 0;
        CK(cudaMalloc((void**)&ctx->fin_hist, hist_need * 2));
        ctx->fin_hist_cap = hist_need * 2;
    }
    if (!ctx->fin_counter) CK(cudaMalloc((void**)&ctx->fin_counter, 4));
    size_t rows_need = (size_t)n_max * (size_t)wpr * 4;
    if (rows_need > ctx->out_rows_cap) {
        if (ctx->d_out_rows) cudaFree(ctx->d_out_rows);
        ctx->d_out_rows = nullptr; ctx->out_rows_cap = 0;
        CK(cudaMalloc((void**)&ctx->d_out_rows, rows_need * 2));
        ctx->out_rows_cap = rows_need * 2;
    }
    unsigned long long *k0 = ctx->fin_k[0], *k1 = ctx->fin_k[1];
    unsigned *v0 = ctx->fin_v[0], *v1 = ctx->fin_v[1];
    unsigned* hist = ctx->fin_hist;
    int* counter = ctx->fin_counter;
    CK(cudaMemsetAsync(counter, 0, 4, st));
    unsigned cap = ctx->bt.mask + 1u;
    k_compact<<<(cap + 255) / 256, 256, 0, st>>>(ctx->bt, k0, v0, counter);
    if (speculative || (n > 1 && n <= SORT_SMALL_MAX && !ctx->opt_force_radix)) {
        int cap = SORT_SPECULATE_MAX;                     // power of two >= the number of keys the sort can meet
        while (cap < (speculative ? (int)std::min<long long>(n_max, SORT_SMALL_MAX) : n)) cap <<= 1;
        if (cap > ctx->sort_smem_cap) {
            CK(cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SORT_SMALL_MAX * 12));
            ctx->sort_smem_cap = SORT_SMALL_MAX;
        }
        k_sort_small<<<1, 1024, (size_t)cap * 12, st>>>(k0, v0, counter, k1, v1, cap);
        std::swap(k0, k1);
        std::swap(v0, v1);
        ctx->tm.n_launches += 3;
    } else if (n > 1) {
        for (int pass = 0; pass < 8; ++pass) {
            int shift = pass * 8;
            k_rs_hist<<<nblk, 256, 0, st>>>(k0, n, shift, hist);
            k_rs_scan<<<1, 1024, 0, st>>>(hist, (size_t)256 * nblk);
            k_rs_scatter<<<nblk, 32, 0, st>>>(k0, v0, n, shift, hist, k1, v1);
            std::swap(k0, k1);
            std::swap(v0, v1);
        }
        ctx->tm.n_launches += 26;
    } else {
        ctx->tm.n_launches += 2;
    }
    k_gather_rows<<<1184, 256, 0, st>>>(ctx->bt.bits, v0, counter, (int)n_max, wpr, ctx->d_out_rows);
    CK(cudaGetLastError());
    ctx->fin_sorted_k = k0;
    ctx->fin_n_max = n_max;
    if (speculative) {   // the read back that verifies the run, and the end-of-run time stamp
        CK(cudaMemcpyAsync(ctx->h_verify, ctx->bt.n_keys, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev_run1, st));
    }
    ctx->fin_queued = true;
    ctx->fin_speculative = speculative;
    *n_known = n;
    return HB_OK;
}

static bool finalize_can_speculate(hb_ctx* ctx) { return ctx->table_cap <= 2ll * SORT_SPECULATE_MAX && !ctx->opt_force_radix; }

int hb_finalize_async(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    if (!ctx->running) return fail(ctx, HB_ERR_STATE, "hb_finalize_async before hb_begin");
    if (ctx->fin_queued || !finalize_can_speculate(ctx)) return HB_OK;    // large tables: hb_finalize does it all
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->s_copy));
    int n;
    return queue_finalize(ctx, true, &n);
}

int hb_pack_keys_device(hb_ctx* ctx, uint64_t* d_block, int32_t cap) {
    if (!ctx || !d_block || cap < 1) return HB_ERR_ARG;
    if (!ctx->fin_queued && !ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_pack_keys_device before hb_finalize_async / hb_finalize");
    CK(cudaSetDevice(ctx->device));
    k_pack_block<<<(cap + 1 + 255) / 256, 256, 0, ctx->s_compute>>>(ctx->fin_sorted_k, ctx->fin_counter, (int)ctx->fin_n_max,
                                                                     (unsigned long long*)d_block, cap);
    CK(cudaGetLastError());
    ctx->tm.n_launches += 1;
    return HB_OK;
}

int hb_finalize(hb_ctx* ctx, int64_t* n_bonds, int64_t* words_per_bond) {
    if (!ctx) return HB_ERR_ARG;
    if (ctx->finalized && !ctx->running) {   // idempotent
        if (n_bonds) *n_bonds = ctx->n_bonds;
        if (words_per_bond) *words_per_bond = ctx->bt.wpr;
        return HB_OK;
    }
    if (!ctx->running) return fail(ctx, HB_ERR_STATE, "hb_finalize before hb_begin");
    Trace tr(ctx, "finalize");
    cudaStream_t st = ctx->s_compute;
    const int wpr = ctx->bt.wpr;
    int rc;
    int n = 0;
    bool settled = false;             // ev_run1 recorded and the stream drained
    CK(cudaStreamSynchronize(ctx->s_copy));
    for (int attempt = 0;; ++attempt) {
        if (attempt > 40) return fail(ctx, HB_ERR_NOMEM, "hb_finalize: the bond table kept overflowing");
        if (!ctx->fin_queued && (rc = queue_finalize(ctx, finalize_can_speculate(ctx), &n))) return rc;
        tr.mark("queued");
        if (!ctx->fin_speculative) break;
        // the one round trip of a speculative finalize
        CK(cudaStreamSynchronize(st));
        memcpy(ctx->last_status, ctx->h_verify, 6 * sizeof(int));
        n = ctx->last_status[0];
        const bool bad = ctx->last_status[1] || ctx->last_status[2] || ctx->last_status[3] || ctx->h_verify[6] || ctx->h_verify[7] ||
                         (long long)n * 2 > ctx->table_cap;
        if (!bad) {
            for (auto& sl : ctx->slots) sl.busy = false;
            drain_events(ctx);
            settled = true;
            break;
        }
        ctx->fin_queued = false;                      // what was queued (and anything built on it) is void
        ctx->tm.finalize_retries++;
        if ((rc = full_verify(ctx))) return rc;       // grows the table / re-runs the in-flight batches, then retry
    }
    {
        unsigned long long hits;
        memcpy(&hits, ctx->last_status + 4, sizeof(hits));
        ctx->tm.pairs_emitted = (long long)hits;
    }
    if (!settled) {
        CK(cudaEventRecord(ctx->ev_run1, st));
        CK(cudaStreamSynchronize(st));
    }
    if ((rc = check_xtc_status(ctx))) {      // a frame record was malformed: the run is void
        ctx->running = false;
        ctx->fin_queued = false;
        return rc;
    }
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev_run0, ctx->ev_run1) == cudaSuccess) ctx->tm.ms_run = ms;
    }
    unsigned long long* k0 = ctx->fin_sorted_k;
    ctx->fin_queued = false;
    ctx->d_out_keys = k0;
    tr.mark("sort+gather");
    if (ctx->trace && ctx->bt.phase_cycles && ctx->wire.on) {
        unsigned long long pc[9];
        if (cudaMemcpy(pc, ctx->bt.phase_cycles + 16, sizeof(pc), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double f = (double)std::max<long long>(ctx->tm.frames, 1);
            fprintf(stderr, "[hbgpu] wire BFS, cycles per frame (thread 0): index+direct %.0f | local ids+rewrite %.0f | zero+seed %.0f | "
                            "acc (residue-water) %.0f | found %.0f | table updates %.0f | water-water step %.0f | frontier update %.0f | reset %.0f\n",
                    pc[0] / f, pc[1] / f, pc[2] / f, pc[3] / f, pc[4] / f, pc[5] / f, pc[6] / f, pc[7] / f, pc[8] / f);
        }
    }
    if (ctx->trace && ctx->bt.phase_cycles && ctx->tm.fused_frames) {
        unsigned long long pc[40];
        if (cudaMemcpy(pc, ctx->bt.phase_cycles, sizeof(pc), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double f = (double)ctx->tm.fused_frames;
            fprintf(stderr, "[hbgpu]   B2 (candidate tests) %.0f\n", pc[9] / f);
            if (ctx->opt_fused_debug & 4)
                fprintf(stderr, "[hbgpu]   stream loop, cycles per frame: producer waits for a free stage %.0f, issues %.0f | consumer warp 0 waits for a tile %.0f, "
                                "loads + releases %.0f, tests %.0f\n", pc[25] / f, pc[26] / f, pc[27] / f, pc[28] / f, pc[29] / f);
            if (ctx->opt_fused_debug & 4)
                fprintf(stderr, "[hbgpu]   B2 (warp 0), cycles per frame: sync after the stream %.0f | gather %.0f | fetch %.0f | 27-cell walk %.0f | append %.0f | "
                                "candidates per frame %.0f\n", pc[30] / f, pc[31] / f, pc[32] / f, pc[33] / f, pc[34] / f, pc[35] / f);
            fprintf(stderr, "[hbgpu] fused phases, cycles per frame (thread 0): A gather+grid %.0f | g1 sort+near %.0f | B stream %.0f | "
                            "B tail sync %.0f | C sort %.0f | C collect %.0f | C sync %.0f | C expand %.0f | flush %.0f\n",
                    pc[0] / f, pc[1] / f, pc[2] / f, pc[3] / f, pc[4] / f, pc[5] / f, pc[6] / f, pc[7] / f, pc[8] / f);
            fprintf(stderr, "[hbgpu]   C expand split: (a) setup %.0f | sync %.0f | (b) angle tasks %.0f | sync %.0f | (c) emit %.0f | sync %.0f\n",
                    pc[10] / f, pc[11] / f, pc[12] / f, pc[13] / f, pc[14] / f, pc[15] / f);
        }
    }
    ctx->n_bonds = n;
    ctx->finalized = true;
    ctx->running = false;
    if (n_bonds) *n_bonds = n;
    if (words_per_bond) *words_per_bond = wpr;
    return HB_OK;
}

int hb_get_keys(hb_ctx* ctx, uint64_t* keys) {
    if (!ctx || !keys) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_keys before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_bonds) CK(cudaMemcpy(keys, ctx->d_out_keys, (size_t)ctx->n_bonds * 8, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_get_bits(hb_ctx* ctx, uint32_t* words) {
    if (!ctx || !words) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_bits before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_bonds) CK(cudaMemcpy(words, ctx->d_out_rows, (size_t)ctx->n_bonds * ctx->bt.wpr * 4, cudaMemcpyDeviceToHost));
    return HB_OK;
}

int hb_get_keys_device(hb_ctx* ctx, uint64_t* d_keys) {
    if (!ctx || !d_keys) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_keys_device before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    // device to device on the compute stream: ordered with the caller's work on that stream, no host round trip
    if (ctx->n_bonds) CK(cudaMemcpyAsync(d_keys, ctx->d_out_keys, (size_t)ctx->n_bonds * 8, cudaMemcpyDeviceToDevice, ctx->s_compute));
    return HB_OK;
}

int hb_get_bits_device(hb_ctx* ctx, uint32_t* d_words) {
    if (!ctx || !d_words) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_bits_device before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    if (ctx->n_bonds) CK(cudaMemcpyAsync(d_words, ctx->d_out_rows, (size_t)ctx->n_bonds * ctx->bt.wpr * 4, cudaMemcpyDeviceToDevice, ctx->s_compute));
    return HB_OK;
}

int hb_merge_keys_device(hb_ctx* ctx, const uint64_t* d_gathered, int32_t world, int32_t cap, uint64_t* d_out, int64_t* d_info) {
    if (!ctx || !d_gathered || !d_out || !d_info || world < 1 || cap < 1) return HB_ERR_ARG;
    long long total = (long long)world * cap;
    int m = 2;
    while (m < total) m <<= 1;
    const int pow2_cap = (cap & (cap - 1)) == 0 && cap >= 2;     // merge-only network needs power-of-two runs
    if (total > 16384) return fail(ctx, HB_ERR_ARG, "hb_merge_keys_device: more than 16384 gathered keys; merge on the host");
    CK(cudaSetDevice(ctx->device));
    if (total > 8192) {      // many ranks: de-duplicate in a hash set first, sort only the distinct keys
        cudaStream_t st = ctx->s_compute;
        if (!ctx->merge_table) {
            CK(cudaMalloc((void**)&ctx->merge_table, (size_t)UNION_SLOTS * 8));
            CK(cudaFuncSetAttribute(k_union_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
        }
        k_fill_u64<<<32, 256, 0, st>>>(ctx->merge_table, (size_t)UNION_SLOTS, EMPTY_KEY);
        CK(cudaMemsetAsync(d_info, 0, 16, st));
        k_union_insert<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const unsigned long long*)d_gathered, world, cap,
                                                                        ctx->merge_table, (long long*)d_info);
        k_union_finish<<<1, 1024, 16384 * 8, st>>>(ctx->merge_table, (unsigned long long*)d_out, (long long*)d_info);
        CK(cudaGetLastError());
        ctx->tm.n_launches += 3;
        return HB_OK;
    }
    size_t smem = (size_t)m * 8;
    CK(cudaFuncSetAttribute(k_merge_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_merge_keys<<<1, 1024, smem, ctx->s_compute>>>((const unsigned long long*)d_gathered, world, cap, m, pow2_cap,
                                                     (unsigned long long*)d_out, (long long*)d_info);
    CK(cudaGetLastError());
    ctx->tm.n_launches += 1;
    return HB_OK;
}

// ---- multi-GPU key exchange inside the C ABI: one NCCL all-gather of the distinct bond keys ------------------
int hb_comm_unique_id(void* id128) {
    if (!id128) return HB_ERR_ARG;
    HbNccl& n = hb_nccl();
    if (!n.ok) { g_create_error = "libnccl.so.2 not found (set HBGPU_NCCL_LIB)"; return HB_ERR_STATE; }
    HbNcclId id;
    if (n.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return HB_ERR_CUDA; }
    memcpy(id128, &id, sizeof(id));
    return HB_OK;
}

int hb_comm_destroy(hb_ctx* ctx) {
    if (!ctx) return HB_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm) { cudaStreamSynchronize(ctx->s_compute); hb_nccl().CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    if (ctx->comm_send) cudaFree(ctx->comm_send);
    if (ctx->comm_recv) cudaFree(ctx->comm_recv);
    ctx->comm_send = ctx->comm_recv = nullptr;
    ctx->comm_cap = ctx->comm_world = 0;
    return HB_OK;
}

int hb_comm_init(hb_ctx* ctx, const void* id128, int32_t world, int32_t rank) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return HB_ERR_ARG;
    HbNccl& n = hb_nccl();
    if (!n.ok) return fail(ctx, HB_ERR_STATE, "hb_comm_init: libnccl.so.2 not found (set HBGPU_NCCL_LIB)");
    hb_comm_destroy(ctx);
    CK(cudaSetDevice(ctx->device));
    HbNcclId id;
    memcpy(&id, id128, sizeof(id));
    const int rc = n.CommInitRank(&ctx->nccl_comm, world, id, rank);
    if (rc != 0) {
        ctx->nccl_comm = nullptr;
        return fail(ctx, HB_ERR_CUDA, std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "failed"));
    }
    ctx->comm_world = world;
    ctx->comm_rank = rank;
    return HB_OK;
}

int hb_exchange_keys(hb_ctx* ctx, int32_t cap, uint64_t* d_out, int64_t* d_info) {
    if (!ctx || cap < 1 || !d_out || !d_info) return HB_ERR_ARG;
    if (!ctx->nccl_comm) return fail(ctx, HB_ERR_STATE, "hb_exchange_keys before hb_comm_init");
    if (!ctx->fin_queued && !ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_exchange_keys before hb_finalize_async / hb_finalize");
    if ((long long)ctx->comm_world * cap > 16384) return fail(ctx, HB_ERR_ARG, "hb_exchange_keys: world * cap must not exceed 16384");
    CK(cudaSetDevice(ctx->device));
    if (cap != ctx->comm_cap) {
        CK(cudaStreamSynchronize(ctx->s_compute));
        if (ctx->comm_send) cudaFree(ctx->comm_send);
        if (ctx->comm_recv) cudaFree(ctx->comm_recv);
        ctx->comm_send = ctx->comm_recv = nullptr;
        ctx->comm_cap = 0;
        CK(cudaMalloc((void**)&ctx->comm_send, ((size_t)cap + 1) * 8));
        CK(cudaMalloc((void**)&ctx->comm_recv, (size_t)ctx->comm_world * ((size_t)cap + 1) * 8));
        ctx->comm_cap = cap;
    }
    int rc = hb_pack_keys_device(ctx, (uint64_t*)ctx->comm_send, cap);
    if (rc) return rc;
    const int nrc = hb_nccl().AllGather(ctx->comm_send, ctx->comm_recv, (size_t)cap + 1, 4 /* ncclInt64 */, ctx->nccl_comm, ctx->s_compute);
    if (nrc != 0) return fail(ctx, HB_ERR_CUDA, std::string("ncclAllGather: ") + (hb_nccl().GetErrorString ? hb_nccl().GetErrorString(nrc) : "failed"));
    return hb_merge_keys_device(ctx, (const uint64_t*)ctx->comm_recv, ctx->comm_world, cap, d_out, d_info);
}

int hb_get_counts(hb_ctx* ctx, int64_t* counts) {
    if (!ctx || !counts) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_counts before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->n_bonds) return HB_OK;
    long long* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)ctx->n_bonds * 8));
    int n = (int)ctx->n_bonds;
    k_row_counts<<<(n + 7) / 8, 256, 0, ctx->s_compute>>>(ctx->d_out_rows, n, ctx->bt.wpr, d);
    cudaError_t e = cudaMemcpyAsync(counts, d, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, HB_ERR_CUDA, cudaGetErrorString(e));
    return HB_OK;
}

int hb_joint_occupancy(hb_ctx* ctx, const uint8_t* select, uint32_t* words, int64_t* n_frames_set) {
    if (!ctx || !words) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_joint_occupancy before hb_finalize");
    CK(cudaSetDevice(ctx->device));
    const int n = (int)ctx->n_bonds, wpr = ctx->bt.wpr;
    const long long nf = ctx->n_frames_total;
    const unsigned tail = (nf & 31) ? ((1u << (nf & 31)) - 1u) : 0xffffffffu;
    unsigned char* d_sel = nullptr;
    unsigned* d_out = nullptr;
    unsigned long long* d_cnt = nullptr;
    cudaStream_t st = ctx->s_compute;
    cudaError_t e = cudaMalloc((void**)&d_out, ((size_t)wpr + 4) * 4);     // words, padding, one 8-byte counter
    if (e == cudaSuccess && select && n) e = cudaMalloc((void**)&d_sel, (size_t)n);
    if (e == cudaSuccess && d_sel) e = cudaMemcpyAsync(d_sel, select, (size_t)n, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        d_cnt = (unsigned long long*)(d_out + wpr + (wpr & 1));
        e = cudaMemsetAsync(d_cnt, 0, 8, st);
    }
    unsigned long long cnt = 0;
    if (e == cudaSuccess) {
        k_joint_rows<<<(wpr + 255) / 256, 256, 0, st>>>(ctx->d_out_rows, d_sel, n, wpr, tail, d_out, d_cnt);
        e = cudaMemcpyAsync(words, d_out, (size_t)wpr * 4, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (d_out) cudaFree(d_out);
    if (d_sel) cudaFree(d_sel);
    if (e != cudaSuccess) return fail(ctx, HB_ERR_CUDA, cudaGetErrorString(e));
    if (n_frames_set) *n_frames_set = (int64_t)cnt;
    return HB_OK;
}

int hb_get_frame_column(hb_ctx* ctx, int64_t frame, uint8_t* present) {
    if (!ctx || !present) return HB_ERR_ARG;
    if (!ctx->finalized) return fail(ctx, HB_ERR_STATE, "hb_get_frame_column before hb_finalize");
    if (frame < 0 || frame >= ctx->n_frames_total) return fail(ctx, HB_ERR_ARG, "hb_get_frame_column: frame out of range");
    CK(cudaSetDevice(ctx->device));
    const int n = (int)ctx->n_bonds;
    if (!n) return HB_OK;
    unsigned char* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)n));
    k_frame_column<<<(n + 255) / 256, 256, 0, ctx->s_compute>>>(ctx->d_out_rows, n, ctx->bt.wpr, frame, d);
    cudaError_t e = cudaMemcpyAsync(present, d, (size_t)n, cudaMemcpyDeviceToHost, ctx->s_compute);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_compute);
    cudaFree(d);
    if (e != cudaSuccess) return fail(ctx, HB_ERR_CUDA, cudaGetErrorString(e));
    return HB_OK;
}

int hb_get_timers(hb_ctx* ctx, hb_timers* out) {
    if (!ctx || !out) return HB_ERR_ARG;
    drain_events(ctx);
    ctx->tm.table_capacity = ctx->table_cap;
    ctx->tm.table_grows = ctx->table_grows;
    ctx->tm.fused_fallbacks = ctx->fused_fallbacks;
    *out = ctx->tm;
    return HB_OK;
}

int hb_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) return HB_ERR_ARG;
    return cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault) == cudaSuccess ? HB_OK : HB_ERR_NOMEM;
}

int hb_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? HB_OK : HB_ERR_CUDA; }

}  // extern "C"


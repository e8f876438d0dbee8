/* hbgpu.h - C ABI of libhbgpu.so: B200 (sm_100a) H-bond detection for cgraphs / Bridge HbondAnalysis.
 *
 * This is the drop-in boundary of the hot path.  The reference has no FFI of its own (it is
 * pure Python); each entry point below names the reference code whose work it takes over
 * (paths relative to the reference checkout, evabertalan/cgraphs):
 *
 *   hb_set_topology   <- tables built once per object by NetworkAnalysis.__init__
 *                        cgraphs/mdhbond/network.py:40-97 (+ helpfunctions.py:44-83,107-117)
 *   hb_set_params     <- distance / cut_angle / check_angle / residuewise kwargs, hbond.py:42-52, and the
 *                        method arguments exclude_backbone_backbone / around_radius, hbond.py:96,232
 *   hb_begin          <- `result = {}` + `frames = self.nb_frames`, hbond.py:98-100 / 234-236
 *   hb_submit_frames  <- the per-frame loop body, hbond.py:103-132 (set_hbonds_in_selection) and
 *                        hbond.py:238-286 (set_hbonds_in_selection_and_water_around): position fetch,
 *                        cKDTree builds and ball queries, check_angle (helpfunctions.py:137-159),
 *                        key building and result[key][frame] = True
 *   hb_finalize / hb_get_keys / hb_get_bits
 *                     <- the `result` dict handed to _set_results, hbond.py:134 / 288 (basic.py:94-98)
 *   hb_submit_frames_planar
 *                     <- the trajectory reader's frame decode, basic.py:46-47 (DCD X / Y / Z records)
 *   hb_submit_frames_xtc / hb_xtc_index / hb_xtc_decode_host / hb_xtc_encode
 *                     <- the same decode for GROMACS .xtc files (MDAnalysis' XTC reader = xdrfile's
 *                        xdrfile_decompress_coord_float + the float32 nm -> Angstrom conversion), basic.py:46-47
 *   hb_set_ions + HB_METHOD_ION_CONTACTS
 *                     <- set_add_ion_interactions, hbond.py:57-94 (ion group: basic.py:53-60)
 *   hb_set_wires      <- WireAnalysis.set_water_wires, waterwire.py:191-317: the same neighbour search
 *                        (waterwire.py:200-248) feeds a per-frame breadth-first search over the water
 *                        network (waterwire.py:250-309); results are wire lengths per residue pair and frame
 *   HB_METHOD_WATER_PRESENCE
 *                     <- HydrationAnalysis.set_presence_around, hydration.py:71-107 (the same local-water test)
 *   hb_get_counts / hb_joint_occupancy / hb_get_frame_column
 *                     <- filter_occupancy, compute_joint_occupancy, _compute_graph_in_frame,
 *                        network.py:99-107, 369-376, 1186-1198 (on the packed matrix)
 *   hb_finalize_async / hb_pack_keys_device / hb_merge_keys_device, hb_comm_init / hb_exchange_keys
 *                     <- no counterpart: the multi-GPU key exchange (frames sharded over ranks), with the caller's
 *                        collective or with the library's own NCCL all-gather
 *
 * Conventions: every function returns 0 on success or a negative hb_status; no C++ exception crosses the
 * ABI; the message of the last failure is available from hb_last_error().  The caller owns every host
 * buffer it passes in or out; the library owns all device memory, pinned staging buffers and streams.
 * A context serves one submitting host thread; contexts are independent.  The library is compiled for
 * sm_100a only and has no CPU fallback: without a usable GPU hb_create fails.
 */
#ifndef HBGPU_H
#define HBGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_ctx hb_ctx;

typedef enum {
    HB_OK = 0,
    HB_ERR_CUDA = -1,       /* a CUDA runtime call failed (message has the CUDA error string) */
    HB_ERR_ARG = -2,        /* invalid argument / call order */
    HB_ERR_NOMEM = -3,      /* host or device allocation failed */
    HB_ERR_NODEVICE = -4,   /* no CUDA device / not sm_100 */
    HB_ERR_STATE = -5       /* e.g. results requested before hb_finalize */
} hb_status;

#define HB_METHOD_IN_SELECTION 0   /* HbondAnalysis.set_hbonds_in_selection, hbond.py:96 */
#define HB_METHOD_WATER_AROUND 1   /* HbondAnalysis.set_hbonds_in_selection_and_water_around, hbond.py:232 */
#define HB_METHOD_ION_CONTACTS 2   /* HbondAnalysis.set_add_ion_interactions, hbond.py:57-94 (needs hb_set_ions) */
#define HB_METHOD_WATER_PRESENCE 3 /* HydrationAnalysis.set_presence_around, hydration.py:71-107 (per_residue=False): the
                                      waters within around_radius of any selection point; one row per water, key =
                                      ent_group of the water's entry (upper 32 bits 0). Selection points need no
                                      entries (sel_don = sel_acc = sel_wat = -1); distance / angle are not used */

#define HB_PBC_NONE 0              /* reference semantics: no periodic images */
#define HB_PBC_ORTHO 1             /* opt-in minimum image, orthorhombic cell */
#define HB_PBC_TRICLINIC 2         /* opt-in minimum image, triclinic cell */

/* Topology tables.  "Entries" are the reference's index space donors ++ acceptors ++ waters
 * (network.py:55-67): entry e < nD is a donor, nD <= e < nD+nA an acceptor, the rest waters.
 * "Points" are the distinct heavy atoms behind those entries: selection points (atoms that are donor
 * and/or acceptor; they may additionally be a water of the water group) and water points (waters of
 * the water group that are not selection points).
 *
 * A topology holds n_systems >= 1 independent systems.  A trajectory uses one system for all frames.
 * A batch of static structures (BASELINE config C5) uses one system per "frame"; entries of all systems
 * share one global numbering, points and atoms are numbered per system through the offset arrays.
 */
typedef struct {
    int32_t n_systems;
    const int64_t* sys_atom_off;  /* [n_systems+1] atoms before each system in a ragged coordinate batch */
    const int32_t* sys_sel_off;   /* [n_systems+1] ranges of selection points */
    const int32_t* sys_wat_off;   /* [n_systems+1] ranges of water points */

    int32_t n_sel;                /* selection points, all systems */
    const int32_t* sel_atom;      /* atom index inside its system */
    const int32_t* sel_don;       /* donor entry (global) or -1 */
    const int32_t* sel_acc;       /* acceptor entry (global) or -1 */
    const int32_t* sel_wat;       /* water entry (global) if the atom is also in the water group, else -1 */
    const uint8_t* sel_bb;        /* 1 if the atom name is 'O' or 'N' (backbone filter, hbond.py:97,111) */

    int32_t n_wat;                /* water points, all systems */
    const int32_t* wat_atom;
    const int32_t* wat_ent;       /* water entry (global) */

    int32_t n_ent;                /* entries, all systems */
    const int32_t* ent_group;     /* key group id: index of the entry's id string among the distinct ids */
    const int32_t* ent_resid;     /* _resids, network.py:89 (orientation rule hbond.py:120-122) */
    const int32_t* ent_residue;   /* id of 'segid-resname-resid' (same-residue skip, hbond.py:125-127) */
    const int32_t* ent_h_off;     /* [n_ent+1] CSR into ent_h_atom = heavy2hydrogen, network.py:86 */
    int32_t n_ent_h;
    const int32_t* ent_h_atom;    /* hydrogen atom index inside its system */
} hb_topology;

typedef struct {
    double distance;              /* heavy-atom cut-off, A (hbond.py:42, default 3.5) */
    double around_radius;         /* water_around only: local-water radius (hbond.py:232,251) */
    float cos_threshold;          /* smallest float32 cos_arg that still satisfies
                                     rad2deg(arccos(c)) <= float32(cut_angle) on the host's numpy
                                     (helpfunctions.py:102-104,155) */
    int32_t check_angle;          /* hbond.py:113,265 */
    int32_t exclude_backbone_backbone; /* hbond.py:96,232 */
    int32_t method;               /* HB_METHOD_* */
    int32_t pbc_mode;             /* HB_PBC_*; HB_PBC_NONE reproduces the reference */
    int32_t structure_batch;      /* 0: trajectory of system 0; 1: frame f is system f (ragged batch) */
} hb_params;

typedef struct {
    double ms_total;              /* device time of all search kernels since hb_begin (CUDA events) */
    double ms_h2d;                /* device time of the host->device frame copies */
    double ms_kernel[16];         /* per kernel class, filled when profiling is on (see hb_kernel_name) */
    int64_t launches[16];
    int64_t n_launches;           /* kernels launched since hb_begin */
    int64_t frames;               /* frames processed since hb_begin */
    int64_t pairs_emitted;        /* H-bond hits (entry pairs passing all criteria), all frames */
    int64_t table_capacity;       /* bond hash-table slots */
    int64_t table_grows;
    double ms_run;                /* device time from hb_begin to the end of hb_finalize (CUDA events) */
    int64_t fused_frames;         /* frames processed by the one-CTA-per-frame kernel */
    int64_t fused_fallbacks;      /* times a frame did not fit in shared memory and the run moved to the
                                     general (global-memory cell list) kernels */
    int64_t h2d_bytes;            /* bytes hb_submit_frames copied host -> device (inert atom ranges are skipped) */
    int64_t finalize_retries;     /* times a speculative finalize found an overflow and was redone (anything a caller
                                     queued on top of hb_finalize_async is void then and must be repeated) */
} hb_timers;

int hb_create(hb_ctx** out, int device_ordinal);
int hb_destroy(hb_ctx* ctx);
const char* hb_last_error(hb_ctx* ctx);            /* ctx may be NULL: message of the last failed hb_create */
const char* hb_version(void);
const char* hb_kernel_name(int kernel_class);      /* names for hb_timers.ms_kernel[]; NULL past the end */

int hb_set_topology(hb_ctx* ctx, const hb_topology* topo);
int hb_set_entry_groups(hb_ctx* ctx, const int32_t* ent_group, int32_t n_ent); /* swap key ids only */
int hb_set_params(hb_ctx* ctx, const hb_params* params);
/* Ions for HB_METHOD_ION_CONTACTS (basic.py:53-60): atom index inside the frame and key group id (an index into the
 * same id-string table ent_group refers to) of every ion; include_water: also pair the waters of the water group
 * with the ions (hbond.py:78-83). Call after hb_set_topology. */
int hb_set_ions(hb_ctx* ctx, const int32_t* ion_atom, const int32_t* ion_group, int32_t n_ions, int32_t include_water);
/* Water wires (WireAnalysis.set_water_wires, waterwire.py:191-317). ent_node[e]: graph node of entry e - for a
 * selection entry the residue node (`da_trans`, waterwire.py:60-65: dense rank, in entry order, of the first entry
 * with the same residue-level id; < n_res), for the entry of water w the value n_res + w. Use with
 * HB_METHOD_WATER_AROUND, around_radius = (max_water + 1) * distance / 2, exclude_backbone_backbone = 0 on a
 * trajectory. The run then yields one row per residue pair that is ever connected: key = (node_lo << 32) | node_hi,
 * words_per_bond = nb4 + 2 with nb4 = (ceil(n_frames / 4) + 1) & ~1; byte f of a row (little endian inside the words)
 * = waters of the shortest wire in frame f (1..max_water, 0 = none) | 0x80 if the residues are directly H-bonded;
 * words nb4, nb4 + 1 = a 64-bit value v, 0 if there never is a direct bond, else ~v = (first frame with a direct
 * bond << 32) | (its smallest acceptor entry << 1) | (1 if the donor is node_hi) - what a caller needs to orient the
 * key string like the reference's first discovery (waterwire.py:277-279, 299-301). ent_node = NULL switches the
 * mode off. The bit-matrix post-processing calls (hb_get_counts, ...) do not apply to such rows. */
int hb_set_wires(hb_ctx* ctx, const int32_t* ent_node, int32_t n_ent, int32_t n_res, int32_t max_water,
                 int32_t allow_direct_bonds);
/* residuewise = False (waterwire.py:60-65, 277-301): wires stay booked on the node of the residue's first entry
 * (ent_node of hb_set_wires, which then holds ATOM nodes: the dense rank of the entry's atom-level id), direct H-bonds
 * on the node of the entry's own atom: direct_node[e]. Call after hb_set_wires; NULL = direct bonds use ent_node. */
int hb_set_wire_direct_nodes(hb_ctx* ctx, const int32_t* direct_node, int32_t n_ent);
/* Convex-hull option of the water wires (set_water_wires(water_in_convex_hull=...), waterwire.py:218-228; the only form
 * cgraphs itself calls, proteingraphanalyser.py:484). The Delaunay test stays on the host (scipy / Qhull, as in the
 * reference); the host then hands over, per frame, the waters whose POSITIONS enter the distance searches (true_wat:
 * the local waters inside the hull) and the waters the pairs are ATTRIBUTED to (label_wat: the reference keeps naming
 * them by the unfiltered local-water list, so the j-th in-hull water is called like the j-th local water; the angle
 * criterion is evaluated with the labelled water's coordinates and hydrogens). Indices are water points;
 * frame_off[n_frames + 1] indexes both lists by frame. Use with hb_set_wires and HB_METHOD_IN_SELECTION (the
 * acceptor-donor pairs come from the ordinary search, the selection-water and water-water pairs from these lists),
 * exclude_backbone_backbone = 0, no pbc_mode. frame_off = NULL switches the mode off. */
int hb_set_wire_virtual_waters(hb_ctx* ctx, const int64_t* frame_off, const int32_t* true_wat, const int32_t* label_wat,
                               int64_t n_frames);
int hb_set_option(hb_ctx* ctx, const char* name, int64_t value);
        /* "batch_bytes": device budget per frame batch; "table_capacity": initial bond-table slots;
           "profile": 1 = time every kernel class with CUDA events; "max_batch_frames";
           "fused": 0 = never use the one-CTA-per-frame kernel; "fused_lw_cap" / "fused_cell_cap": its
           shared-memory capacities (0 = automatic); "fused_pair_cap", "fused_stages" (stages of its water-tile ring =
           warps that stream, 2..8), "fused_prefetch" (tiles pulled into L2 ahead of the ring), "fused_wait_ns",
           "fused_batch_frames", "fused_stagger" (cycles over which the first wave of its CTAs spreads its start, 0 = off),
           "fused_sel_prefetch" (0 = no bulk L2 prefetch of the selection's atom block), "fused_max_streams" (frames
           of one SM that may stream at the same time, 0 = no limit); "pair_buf_bytes": batch-wide pair buffer of the general kernels (0 = per-point kernel);
           "frame_pairs": 0 = never run the general path's pair search as one CTA per frame (hb_framepairs.cuh);
           "lw_smem": 0 = keep the selection grid of the general path's local-water test in global memory;
           "wire_edge_cap": H-bond pairs per frame the edge lists of the water-wire mode start with (doubles on
           overflow), "wire_smem_words": shared-memory budget of its bit sets (test hook);
           "copy_ranges": 0 = always copy whole frames; "stage_threads": host threads that move pageable frames
           into the pinned staging buffers (default min(16, cores)); "force_radix_sort" (test hook);
           "compute_stream": a cudaStream_t of the caller on which every kernel of the context is launched, so
           that the caller's own stream-ordered work (events, collectives) brackets the search (0 = library stream) */

/* Start an analysis over n_frames_total frames (the width of the occupancy matrix). */
int hb_begin(hb_ctx* ctx, int64_t n_frames_total);

/* Feed the next n_frames frames.  xyz is float32 [n_frames][n_atoms][3] in host memory (pinned memory is
 * copied directly, pageable memory goes through the library's two pinned staging buffers).  With
 * structure_batch the frames are systems [next, next + n_frames) concatenated raggedly.
 * box: NULL, or 9 floats per frame (lattice vectors as rows) when pbc_mode != HB_PBC_NONE.
 * Asynchronous: returns once the last batch is queued; buffers passed in may be reused after return. */
int hb_submit_frames(hb_ctx* ctx, const float* xyz, int64_t n_atoms, int64_t n_frames, const float* box);

/* Same, for planar frame records as trajectory files store them (DCD: per frame X[n_atoms], Y[n_atoms], Z[n_atoms]
 * as float32 inside a record of frame_stride bytes; x_off / y_off / z_off = byte offsets of the three planes
 * inside a frame record, all multiples of 4). The records are copied to the device as they are (inert atom ranges
 * skipped) and interleaved there - the decode MDAnalysis' reader does for the reference (basic.py:46-47). */
int hb_submit_frames_planar(hb_ctx* ctx, const void* raw, int64_t frame_stride, int64_t x_off, int64_t y_off,
                            int64_t z_off, int64_t n_atoms, int64_t n_frames, const float* box);

/* Same, for GROMACS XTC frame records (compressed, about 3.5 bytes per atom): raw + frame_off[f] is the first byte of
 * frame f's record (magic 1995), frame_off[n_frames] the end of the last one; offsets are multiples of 4 (XDR). The
 * records are copied to the device as they are and decoded there, one warp per frame, with the arithmetic of the
 * reader the reference uses through MDAnalysis: x = fl32(fl32(float(int) * fl32(1 / precision)) * scale); scale = 10
 * gives Angstrom (MDAnalysis' convert_units), 1 leaves nm. box as in hb_submit_frames (the caller reads the cell of a
 * frame from its record header, see hb_xtc_decode_host). A malformed record makes this or a later call (at the latest
 * hb_finalize) fail with HB_ERR_ARG. */
int hb_submit_frames_xtc(hb_ctx* ctx, const void* raw, const int64_t* frame_off, int64_t n_atoms, int64_t n_frames,
                         float scale, const float* box);
/* Decode XTC records (host memory) straight into a caller-provided DEVICE buffer float32 [n_frames][n_atoms][3], e.g.
 * to keep a decoded trajectory resident for hb_submit_frames_device. Synchronous. */
int hb_xtc_decode_to_device(hb_ctx* ctx, const void* raw, const int64_t* frame_off, int64_t n_atoms, int64_t n_frames,
                            float scale, float* d_out);
/* Host helpers for XTC files (no context, no GPU needed).
 * hb_xtc_index: walk the frame records in buf[0, nbytes): frame_off[0 .. min(n, max_frames)] (may be NULL to count),
 * *n_frames = n complete records found, *n_atoms their atom count (HB_ERR_ARG if it changes between frames).
 * hb_xtc_decode_host: decode ONE record on the CPU (frame 0 for the constructor's bonded-hydrogen detection,
 * network.py:54-86; spot checks): xyz [n_atoms][3] with the arithmetic above, box9 = lattice rows in nm (unscaled).
 * hb_xtc_encode: write records (test and benchmark inputs): xyz [n_frames][n_atoms][3] times in_scale are nm (0.1 for
 * Angstrom input), precision 1000 is GROMACS' default, box9 NULL or 9 floats per frame (nm), frame f gets step
 * first_step + f and time dt * step. out = NULL asks for the size only; frame_off (n_frames + 1, may be NULL) receives
 * the record offsets; n_threads host threads encode frame blocks in parallel. */
int hb_xtc_index(const void* buf, int64_t nbytes, int64_t* frame_off, int64_t max_frames, int64_t* n_frames, int64_t* n_atoms);
int hb_xtc_decode_host(const void* rec, int64_t avail, int64_t n_atoms, float scale, float* xyz, float* box9, int64_t* rec_bytes);
int hb_xtc_encode(const float* xyz, int64_t n_atoms, int64_t n_frames, float in_scale, float precision, const float* box9,
                  int64_t first_step, float dt, void* out, int64_t out_cap, int64_t* frame_off, int64_t* bytes_written,
                  int32_t n_threads);

/* Same, for coordinates already resident in device memory (the caller guarantees they are complete, and keeps
 * d_xyz / d_box valid and unchanged until hb_sync or hb_finalize returns: a batch may be re-run on overflow). */
int hb_submit_frames_device(hb_ctx* ctx, const float* d_xyz, int64_t n_atoms, int64_t n_frames,
                            const float* d_box);

/* Optional: queue the finalisation (compaction, key sort, row gather) behind the search WITHOUT waiting, so that
 * the caller can queue more stream-ordered work on the compute stream (hb_pack_keys_device, a collective,
 * hb_merge_keys_device) before the single host wait in hb_finalize. A no-op for large bond tables. */
int hb_finalize_async(hb_ctx* ctx);
/* Wait for all queued work, sort the distinct bond keys, gather the occupancy rows in key order. */
int hb_finalize(hb_ctx* ctx, int64_t* n_bonds, int64_t* words_per_bond);

/* keys[i] = (group_first << 32) | group_second, ascending; bits = [n_bonds][words_per_bond] uint32, bit
 * (f & 31) of word (f >> 5) of row i is set iff bond i exists in frame f. */
int hb_get_keys(hb_ctx* ctx, uint64_t* keys);
int hb_get_bits(hb_ctx* ctx, uint32_t* words);
/* Device-side access for collectives: copy keys / rows into caller-provided DEVICE buffers. The copies are
 * queued on the context's compute stream (see "compute_stream"); use hb_sync or that stream to wait for them. */
int hb_get_keys_device(hb_ctx* ctx, uint64_t* d_keys);
int hb_get_bits_device(hb_ctx* ctx, uint32_t* d_words);
/* Multi-GPU exchange, send side: write this rank's block [count, key_0 .. key_{cap-1}] (uint64, keys ascending,
 * padding after them) to d_block (cap + 1 words). After hb_finalize_async or hb_finalize; stream-ordered. */
int hb_pack_keys_device(hb_ctx* ctx, uint64_t* d_block, int32_t cap);
/* Multi-GPU exchange, merge step: d_gathered = all-gather of one [count, key_0 .. key_{cap-1}] block
 * per rank (uint64, keys sorted ascending); writes the global sorted distinct keys to d_out (room for world*cap)
 * and {n_out, overflow flag} to d_info. world*cap <= 16384. Queued on the compute stream, no host round trip. */
int hb_merge_keys_device(hb_ctx* ctx, const uint64_t* d_gathered, int32_t world, int32_t cap, uint64_t* d_out,
                         int64_t* d_info);
/* The same exchange without Python: the library loads NCCL at run time (libnccl.so.2, or $HBGPU_NCCL_LIB) and runs the
 * one collective of the path itself, so that any host language can drive several GPUs (one process or thread per GPU,
 * one context each).
 *   hb_comm_unique_id : rank 0 creates the 128-byte NCCL id; the caller hands it to every rank (file, socket, MPI ...)
 *   hb_comm_init      : ncclCommInitRank on the context's device; collective over all ranks
 *   hb_exchange_keys  : after hb_finalize_async (or hb_finalize): pack this rank's distinct keys, ONE ncclAllGather of
 *                       (cap + 1) 64-bit words per rank, merge - all queued on the compute stream, nothing waits on the
 *                       host. d_out (device, world * cap words) receives the global sorted distinct keys, d_info
 *                       (device, 2 words) {n_keys, overflow}: overflow != 0 means some rank had more than cap keys -
 *                       repeat with a larger cap. world * cap <= 16384. Global bond id = rank of a key in d_out.
 *   hb_comm_destroy   : frees the communicator (hb_destroy does it too). */
int hb_comm_unique_id(void* id128);
int hb_comm_init(hb_ctx* ctx, const void* id128, int32_t world, int32_t rank);
int hb_exchange_keys(hb_ctx* ctx, int32_t cap, uint64_t* d_out, int64_t* d_info);
int hb_comm_destroy(hb_ctx* ctx);
/* Per-bond popcount over frames (occupancy numerator; network.py:99-107 filter_occupancy). */
int hb_get_counts(hb_ctx* ctx, int64_t* counts);
/* AND of the occupancy rows with select[i] != 0 (NULL: all rows): words = uint32 [words_per_bond] (host),
 * n_frames_set = frames in which every selected bond exists (network.py:369-376 compute_joint_occupancy). */
int hb_joint_occupancy(hb_ctx* ctx, const uint8_t* select, uint32_t* words, int64_t* n_frames_set);
/* present[i] = 1 iff bond i exists in `frame` (network.py:1186-1198 _compute_graph_in_frame). */
int hb_get_frame_column(hb_ctx* ctx, int64_t frame, uint8_t* present);

int hb_get_timers(hb_ctx* ctx, hb_timers* out);
int hb_sync(hb_ctx* ctx);

/* Pinned host memory helpers for callers that want zero-staging H2D. */
int hb_host_alloc(void** ptr, uint64_t bytes);
int hb_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* HBGPU_H */

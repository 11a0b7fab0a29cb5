/* dipgenie_cuda.h — C ABI of libdipgenie_cuda.so, the B200 (sm_100a) device side of DipGenie's hot path.
 *
 * The reference (gsc74/DipGenie) has no plugin/FFI layer: its hot path is a handful of C++ member
 * functions called from one place each.  Every entry point below names the reference seam it replaces
 * (file:line into the reference tree); INTEGRATION.md shows the few lines a maintainer adds at that
 * seam to call it.  Conventions: plain pointers and sizes only; caller-owned host buffers unless a
 * parameter is documented as library-allocated (release those with dg_free); every function returns
 * 0 on success or a negative error code, with text available from dg_last_error(); no exceptions and
 * no C++ types cross the boundary; one dg_ctx per GPU, used by one host thread at a time.
 * There is no CPU fallback: dg_create fails if no CUDA device is usable.
 */
#ifndef DIPGENIE_CUDA_H
#define DIPGENIE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dg_ctx dg_ctx;

#define DG_OK 0
#define DG_ERR_ARG (-1)       /* malformed input (see dg_last_error) */
#define DG_ERR_CUDA (-2)      /* CUDA runtime error */
#define DG_ERR_NOMEM (-3)     /* device or host allocation failed */
#define DG_ERR_CAPACITY (-4)  /* an output did not fit the documented capacity */

/* ---- context -------------------------------------------------------------------------------- */
dg_ctx* dg_create(int device);               /* NULL if the device cannot be initialised */
void dg_destroy(dg_ctx* ctx);
const char* dg_last_error(dg_ctx* ctx);      /* valid until the next call on ctx */
void dg_free(void* p);                       /* releases library-allocated host arrays */
/* Device buffers of destroyed problems stay in the context's stream-ordered pool for reuse (a cudaFree of a GB costs
 * tens of ms).  After a phase that held far more device memory than what follows (e.g. a hundred resident problems),
 * hand the cached blocks back: waits for the device to be idle. */
int dg_release_cached_memory(dg_ctx* ctx);
int dg_device_info(dg_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes);

/* ---- diploid DP: Approximator::diploid_dp_approximation_solver, src/approximator.cpp:362-785 ---
 * (the DP sweep :532-716 and the recombination-edge lists it returns :757-785; the caller keeps the
 *  sequence stitching :787-930, which needs node_seq/paths).
 *
 * Input: the levelized ExpandedGraph (src/ExpandedGraph.hpp:16-26 after
 * strict_bfs_levelize_and_reorder, :269-409), flattened:
 *   level_off[L+1]   vertices are numbered in (level,id) order; level l owns [level_off[l], level_off[l+1])
 *                    (= g.vertices_in_level, whose entries are consecutive ids after the reorder)
 *   adj_off[V+1], adj_dst[E], adj_w[E]   g.adj_list as CSR, adjacency order preserved, weights 0/1
 *   col_off[V+1], col_val[]              g.color as CSR (colour ids)
 *   colour_is_hom[n_colours]             color_homo_bv (src/approximator.cpp:1283-1290)
 *   R                                    recombination_limit
 * Output (bit-exact with the reference, same tie-breaking :657-659):
 *   sink_value   dp value of cell (r=R,0,0) of the last level ("DP value:", :774-776)
 *   sink_s_het   dp_entry::s_het of that cell
 *   p1_edges / p2_edges   weighted_p1_edges / weighted_p2_edges (:781-782) as (from,to) vertex-id pairs,
 *                oldest first; capacity 2*(R+2) int32 each; n_p1 / n_p2 = number of pairs
 * Refused with DG_ERR_ARG, before anything is launched: null arrays, level 0 with more than one vertex, an empty
 * level, offsets that are not monotone, an edge that does not go to the next level, a colour id outside
 * [0, n_colours), an edge weight above 1, and parallel edges of differing weight between one vertex pair (the
 * reference resolves their ties by the timing of its racing relax loop, :627-701 — there is no result to be exact
 * with; edges of equal weight may repeat).
 */
int dg_dp_diploid(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off,
                  const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val,
                  const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                  int32_t* sink_value, int32_t* sink_s_het,
                  int32_t* p1_edges, int32_t* n_p1, int32_t* p2_edges, int32_t* n_p2);

/* The same computation split so that the graph can stay resident in HBM between runs
 * (bench.py times dg_dip_run alone; tests read the per-level checksums). */
typedef struct dg_dip dg_dip;

int dg_dip_create(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off,
                  const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val,
                  const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, dg_dip** out);
/* flags: bit0 = also fold every DP layer into per-level checksums (slower; tests only);
 *        bit1 = record phase cycle counters of CTA 0 (read back with dg_dip_profile; diagnostics). */
int dg_dip_run(dg_ctx* ctx, dg_dip* d, uint32_t flags);
int dg_dip_result(dg_ctx* ctx, dg_dip* d, int32_t* sink_value, int32_t* sink_s_het,
                  int32_t* p1_edges, int32_t* n_p1, int32_t* p2_edges, int32_t* n_p2);
/* level_checksum/level_live: [n_levels]; entry l (l>=1) folds (flat index, value, pred_i, pred_j) of the
 * live cells of level l exactly like oracle/ref_hook.h does inside the reference. */
int dg_dip_checksums(dg_ctx* ctx, dg_dip* d, uint64_t* level_checksum, uint64_t* level_live);
/* Work and traffic accounting: U = (R+1)*sum E_l^2 cell-updates, C = (R+1)*sum k_l^2
 * destination cells, B = (R+1)*sum(4 k_l^2 + 5 k_{l+1}^2) algorithmic bytes; plus kernel launches and
 * device milliseconds (CUDA events on the context's stream) of the last dg_dip_run (valid after
 * dg_dip_result). */
typedef struct {
    uint64_t cell_updates, cells, algo_bytes;
    uint64_t device_bytes;       /* HBM held by this problem */
    int32_t n_levels, n_vertices, max_width, max_indegree, mask_words_max, grid_ctas, pred_bytes;
    int32_t launches;
    float sweep_ms, traceback_ms;       /* dip_sweep_kernel; the four traceback kernels */
    float delta_ms;                     /* dip_delta_kernel (pair-score matrices) */
    float plan_ms, upload_ms;           /* host planning and H2D inside dg_dip_create */
    int32_t n_narrow, n_wide;           /* transitions run by CTA 0 alone in shared memory / spread over CTAs */
    int64_t n_tasks;                    /* tasks over all CTA streams */
    uint64_t delta_bytes;               /* HBM held by the pair-score matrices */
    uint64_t prog_bytes;                /* level-program engine: HBM held by the level programs */
    uint64_t code_bytes;                /* HBM held by the predecessor codes */
    int32_t engine;                     /* 4 = level programs (dp_prog.h), 3 = task streams */
    float build_ms;                     /* prog_fill_kernel inside dg_dip_create (engine 4) */
    uint64_t cells_written;             /* engine 4: destination cells the sweep writes (in-place layers skip the passive pairs) */
    int32_t n_relocate, pad_;           /* engine 4: transitions that move the level between the shared-memory and the HBM tile */
    uint64_t h2d_bytes;                 /* host arrays dg_dip_create copied to the device (the program itself is built there) */
} dg_dip_stats_t;
int dg_dip_stats(dg_ctx* ctx, dg_dip* d, dg_dip_stats_t* out);
/* out24: for each of {shared-memory layers, HBM/L2 layers} six counters {tasks, slot-wait, grid-wait,
 * cell-loop, barrier+arrive, unused} in SM clock cycles of CTA 0 / thread 0; the rest is zero. */
int dg_dip_profile(dg_ctx* ctx, dg_dip* d, uint64_t* out24);
/* Diagnostics/tests: the level programs the device builder wrote for this problem (engine 4; *bytes = 0 otherwise).
 * out may be NULL to query the size. */
int dg_dip_debug_program(dg_ctx* ctx, dg_dip* d, uint8_t* out, uint64_t cap, uint64_t* bytes);
void dg_dip_destroy(dg_ctx* ctx, dg_dip* d);

/* Independent samples on one GPU at the same time.  In the reference every sample is its own process
 * (data/run_DipGenie_batch.sh:21-39: one DipGenie run per sample of the 22-sample leave-one-out study); the
 * DP of one small-panel sample is a chain of dependent levels that keeps only a few SMs busy, so a batch is
 * spread over the GPU: sample i runs as its own persistent sweep of `ctas_per_sample` CTAs (0 = 4) on its own
 * stream, up to `max_concurrent` samples resident together (0 = min(SM count / ctas_per_sample, 32)), while
 * host threads plan the next samples.  No data-path exchange between samples.  Results are identical to n calls of
 * dg_dp_diploid.  A sample whose level programs and predecessor codes do not fit the free HBM (wide panels need
 * gigabytes) waits for earlier samples to finish instead of failing its allocation.  out[i].status = DG_OK or that sample's error code; returns the first error, if any. */
#define DG_BATCH_MAX_EDGES 64        /* capacity of the edge lists below: needs R + 2 <= 64 */
typedef struct {
    int32_t n_levels; const int32_t* level_off;
    const int64_t* adj_off; const int32_t* adj_dst; const uint8_t* adj_w;
    const int64_t* col_off; const int32_t* col_val;
    const uint8_t* colour_is_hom; int32_t n_colours; int32_t R;
} dg_dip_input_t;
typedef struct {
    int32_t status, sink_value, sink_s_het, n_p1, n_p2;
    int32_t p1_edges[2 * DG_BATCH_MAX_EDGES], p2_edges[2 * DG_BATCH_MAX_EDGES];
} dg_dip_output_t;
int dg_dp_diploid_batch(dg_ctx* ctx, int32_t n, const dg_dip_input_t* in, dg_dip_output_t* out, int32_t max_concurrent,
                        int32_t ctas_per_sample);

/* ---- Row-sharded diploid DP over the GPUs of one node (north_star: "h1-row tiles of the diploid matrix" for large
 * panels).  One process per GPU.  Every rank passes the SAME graph; wide level transitions are split by
 * destination row (P1 axis) over world x ctas CTAs, each CTA writes its rows of the layer and of the predecessor
 * codes into every peer's buffers over NVLink and all grid barriers span the GPUs; narrow transitions run redundantly
 * on every rank.  Every rank ends with the complete result (dg_dip_result).  Protocol, per rank:
 *   dg_dip_create_sharded -> dg_dip_ipc_export -> (all-gather the world x 4 handles, e.g. torch.distributed)
 *   -> dg_dip_ipc_attach -> { dg_dip_shard_arm -> host barrier over the ranks -> dg_dip_run -> dg_dip_result }*
 * The sweeps of all ranks must be launched concurrently (they spin on each other).  ctas = 0: one CTA per SM. */
#define DG_IPC_HANDLE_BYTES 64
#define DG_MAX_SHARD_RANKS 8
int dg_dip_create_sharded(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off, const int64_t* adj_off,
                          const int32_t* adj_dst, const uint8_t* adj_w, const int64_t* col_off, const int32_t* col_val,
                          const uint8_t* colour_is_hom, int32_t n_colours, int32_t R, int32_t rank, int32_t world,
                          int32_t ctas, dg_dip** out);
int dg_dip_ipc_export(dg_ctx* ctx, dg_dip* d, uint8_t* handles /* [4 * DG_IPC_HANDLE_BYTES] */);
int dg_dip_ipc_attach(dg_ctx* ctx, dg_dip* d, const uint8_t* all_handles /* [world * 4 * DG_IPC_HANDLE_BYTES], rank-major */);
int dg_dip_shard_arm(dg_ctx* ctx, dg_dip* d);
/* The same protocol with all `world` ranks as sibling problems of one process on one GPU (no IPC; tests and
 * single-GPU boxes): replaces export/attach; then arm every sibling, dg_dip_run every sibling (asynchronous, own
 * streams), dg_dip_result. */
int dg_dip_attach_in_process(dg_ctx* ctx, dg_dip** all, int32_t world);

/* The same with the graphs resident in HBM between runs (bench.py times dg_dip_run_many alone): a problem
 * created in `slot` owns that slot's stream and a sweep grid of `ctas` CTAs (0 = 8); dg_dip_run_many starts the
 * n problems together (distinct slots run concurrently), waits for all of them and returns the device time of
 * the whole group (CUDA events on the context's stream around fork and join).  dg_dip_result / dg_dip_stats /
 * dg_dip_destroy apply as for dg_dip_create. */
int dg_dip_create_slot(dg_ctx* ctx, int32_t n_levels, const int32_t* level_off,
                       const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                       const int64_t* col_off, const int32_t* col_val,
                       const uint8_t* colour_is_hom, int32_t n_colours, int32_t R,
                       int32_t slot, int32_t ctas, dg_dip** out);
int dg_dip_run_many(dg_ctx* ctx, dg_dip** problems, int32_t n, float* device_ms);

/* ---- haploid DP: Approximator::dp_approximation_solver, src/approximator.cpp:44-168 ------------
 * (the push-relaxation loop :50-67, the R+1 tracebacks with their distinct-colour counts :70-102 and
 *  the final traceback :141-153; the caller keeps the floating-point best_r rule :116-136, the
 *  original-vertex expansion and de-duplication :154-167 and the FASTA writer :1266-1277).
 *
 * Input: the ExpandedGraph after topologically_reorder (src/ExpandedGraph.hpp:29-102), flattened:
 *   n                                     vertices in Kahn order (every edge goes to a larger id; sink = n-1)
 *   adj_off[n+1], adj_dst[E], adj_w[E]    g.adj_list as CSR, adjacency order preserved, weights 0/1
 *   col_off[n+1], col_val[]               g.color as CSR (colour ids < n_colours)
 *   R                                     recombination_limit
 * Output (bit-exact with the reference, same first-writer tie-breaking :60):
 *   colours_by_r[R+1]   distinct colours on the traceback path of layer r ("true score", :100)
 *   path_off[R+2], *paths   the R+1 traceback paths, source first (:153), as vertex ids, concatenated;
 *                       *paths is library-allocated (release with dg_free)
 */
int dg_dp_haploid(dg_ctx* ctx, int32_t n, const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val, int32_t n_colours, int32_t R,
                  int32_t* colours_by_r, int64_t* path_off, int32_t** paths);

/* The same computation with the graph resident in HBM between runs. */
typedef struct dg_hap dg_hap;
int dg_hap_create(dg_ctx* ctx, int32_t n, const int64_t* adj_off, const int32_t* adj_dst, const uint8_t* adj_w,
                  const int64_t* col_off, const int32_t* col_val, int32_t n_colours, int32_t R, dg_hap** out);
int dg_hap_run(dg_ctx* ctx, dg_hap* d);
int dg_hap_result(dg_ctx* ctx, dg_hap* d, int32_t* colours_by_r /* [R+1] */, int32_t* path_len /* [R+1] */);
/* path of layer r, source first, vertex ids; DG_ERR_CAPACITY (with *len set) when cap is too small */
int dg_hap_path(dg_ctx* ctx, dg_hap* d, int32_t r, int32_t* path, int32_t cap, int32_t* len);
/* U = (R+1)*nE cell-updates, C = (R+1)*n cells, B = (R+1)*(8n + 8nE) algorithmic bytes */
typedef struct {
    uint64_t cell_updates, cells, algo_bytes, device_bytes;
    int32_t n_levels, n_vertices, max_width, max_indegree;
    int32_t launches;
    float sweep_ms, traceback_ms;
} dg_hap_stats_t;
int dg_hap_stats(dg_ctx* ctx, dg_hap* d, dg_hap_stats_t* out);
void dg_hap_destroy(dg_ctx* ctx, dg_hap* d);

/* ---- minimizer sketch / read spectrum / walk-index join ------------------------------------------
 * Sequences are passed concatenated: sequence s = bases[seq_off[s] .. seq_off[s+1]) (any case, any bytes;
 * the comparison order is the reference's: lexicographic on the upper-cased canonical ASCII k-mer,
 * rightmost minimum in a window, emit when the hash changes).  hash = MurmurHash3_x64_128
 * (seed 0) h1^h2 of the k canonical bytes (src/solver.cpp:16-24).  Arrays returned through T** are
 * library-allocated host arrays (release with dg_free).
 */

/* Raw minimizer lists: Solver::index_kmers' (hash, start) sequence (src/solver.cpp:303-361) and the
 * pre-set content of Solver::compute_hashes (:376-409) for every sequence, in sequence order.
 * seq_count[n_seq] (nullable) = minimizers per sequence; starts = k-mer start within its sequence. */
int dg_sketch_minimizers(dg_ctx* ctx, const uint8_t* bases, const uint64_t* seq_off, uint32_t n_seq, int k, int w,
                         uint64_t* seq_count, uint64_t** hashes, uint64_t** starts);

/* Solver::compute_hashes over all reads (src/solver.cpp:366-412, :528-532) + the read spectrum Sp_R
 * (:534-555: distinct hashes ascending, id = index) + kmer_count (:711-732: reads containing the hash). */
int dg_sketch_reads(dg_ctx* ctx, const uint8_t* bases, const uint64_t* read_off, uint32_t n_reads, int k, int w,
                    uint64_t** spectrum, uint32_t** read_count, uint64_t* n_spectrum);

/* Solver::index_kmers for every walk (src/solver.cpp:277-363) joined with the spectrum like
 * Solver::compute_anchors (:415-446, :560-576).  Panel as flat arrays: segment v = seg_bases[seg_off[v] ..
 * seg_off[v+1]) (node_seq), walk h = walk_vtx[walk_off[h] .. walk_off[h+1]) (paths), top_order_map[n_seg].
 * Output, walk by walk in walk order: n_minimizers[n_walks] (all minimizers of the walk, before the join:
 * the log line of :474), hit_off[n_walks+1], hit_sid[n_hits] (spectrum id), hit_vtx_off[n_hits+1] and
 * hit_vtx (the unique vertices under bases [start,start+k) ordered by top_order_map, :343-357). */
int dg_index_walks(dg_ctx* ctx, const uint8_t* seg_bases, const uint64_t* seg_off, uint32_t n_seg,
                   const int32_t* walk_vtx, const uint64_t* walk_off, uint32_t n_walks,
                   const int32_t* top_order_map, int k, int w,
                   const uint64_t* spectrum, uint64_t n_spectrum,
                   uint64_t* n_minimizers, uint64_t** hit_off, uint32_t** hit_sid,
                   uint64_t** hit_vtx_off, int32_t** hit_vtx);

/* Accounting of the last sketch call on ctx: input bases, emitted minimizers, spectrum size, join hits,
 * device milliseconds (CUDA events on the context's stream around the kernels) and kernel launches. */
typedef struct {
    uint64_t bases, minimizers, spectrum, hits;
    float kernel_ms;
    int32_t launches;
} dg_sketch_stats_t;
int dg_sketch_last_stats(dg_ctx* ctx, dg_sketch_stats_t* out);

#ifdef __cplusplus
}
#endif
#endif /* DIPGENIE_CUDA_H */

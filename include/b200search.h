/*
 * b200search.h -- C ABI of libb200search.so (sm_100a).
 *
 * Drop-in boundary for ONE hot path of Axionis47/semantic-search-kd: exact
 * inner-product / cosine top-k of 384-d embeddings against a corpus matrix.
 * The reference has no FFI of its own for this path: it is a duck-typed Python
 * class (FAISSIndexBuilder, source absent from the tree) that forwards to
 * faiss-cpu.  Each entry point below names the reference call it replaces
 * (paths relative to /root/reference).  All functions are extern "C", take
 * plain pointers and sizes, never throw, never abort; they return B2S_OK (0)
 * or a negative B2S_ERR_* code with a thread-local message in b2s_last_error().
 *
 * Inputs are expected to be finite.  A NaN score never passes a threshold test, so rows (or queries)
 * containing NaN simply do not appear in results; nothing crashes.
 *
 * Threading and streams: one in-flight call per index handle, matching the reference's
 * single-worker, event-loop-thread use of .search (src/serve/app.py:258,293;
 * src/config.py:213).  Host-side enqueueing is serialised by an internal mutex.  The handle's
 * workspaces are shared by all of its searches: a device-API call enqueued on ANOTHER stream
 * than the handle's previous search first waits (event) for everything enqueued on that
 * previous stream, so searches of one handle never overlap on the device whatever streams they
 * are issued on; host-API calls run on a private stream under the same rule.  (A search
 * captured into a CUDA graph is exempt: whoever replays the graph orders it.)
 */
#ifndef B200SEARCH_H
#define B200SEARCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B2S_API __attribute__((visibility("default")))
#else
#define B2S_API
#endif

typedef struct b2s_index b2s_index;

enum {
    B2S_OK = 0,
    B2S_ERR_INVALID = -1,     /* bad argument (NULL, dim mismatch, k < 0 ...)        */
    B2S_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed                  */
    B2S_ERR_NOMEM = -3,       /* device or host allocation failed                     */
    B2S_ERR_UNSUPPORTED = -4, /* shape outside what the kernels are built for         */
    B2S_ERR_NO_DEVICE = -5    /* no sm_100 device / CUDA driver not present           */
};

enum { B2S_METRIC_INNER_PRODUCT = 0, B2S_METRIC_COSINE = 1 };
enum { B2S_DTYPE_F32 = 0, B2S_DTYPE_BF16 = 1 };
enum { B2S_PATH_AUTO = 0, B2S_PATH_SCAN = 1, B2S_PATH_TENSOR = 2 };

/*
 * Flags of the device-buffer searches.
 * B2S_SEARCH_STABLE_QUERIES: the caller guarantees that the query buffer was completely written
 *   before the PREVIOUS operation on cuda_stream was enqueued (a pre-encoded query set, a ring of
 *   query slots filled ahead, ...).  The batch-1/2 scan kernel may then start streaming the corpus
 *   while the previous search's tail (candidate merge, cross-GPU exchange) is still running
 *   (programmatic dependent launch with a deferred wait).  Without the flag the kernel only
 *   prefetches immutable corpus rows into L2 before it waits.  Results are identical either way.
 */
enum { B2S_SEARCH_STABLE_QUERIES = 1u };

/* Per-call counters of the last search on a handle (b2s_last_stats). */
typedef struct b2s_stats {
    int32_t path;            /* B2S_PATH_SCAN or B2S_PATH_TENSOR actually taken                 */
    int32_t kernel_launches; /* kernels of this library launched by the last search call        */
    int32_t seeded;          /* 1 if a threshold-seeding pre-pass ran                            */
    int32_t reserved;
    int64_t corpus_bytes;    /* algorithmic bytes of one pass over the local shard              */
    int64_t passes;          /* how many times the dominant kernel streamed the shard           */
    float dominant_ms;       /* device time of the dominant kernel(s), only if timing enabled    */
    float total_ms;          /* device time of the whole call, only if timing enabled            */
} b2s_stats;

/* Library / build identification. */
B2S_API int b2s_version(void);
B2S_API const char* b2s_last_error(void);
B2S_API int b2s_device_count(void);

/*
 * Replaces FAISSIndexBuilder(embedding_dim, index_type, metric)
 *   scripts/build_faiss_index.py:49-53, src/serve/app.py:427-429, and
 *   faiss.IndexFlatIP(384) (tests/conftest.py:184).
 * metric COSINE L2-normalises rows at add time and queries at search time and
 * then uses the inner product (configs/index.yaml:11,30).
 */
B2S_API int b2s_create(int dim, int metric, int device, b2s_index** out);
B2S_API int b2s_destroy(b2s_index* idx);

/* Pre-size the device corpus buffer for n_rows rows (optional). */
B2S_API int b2s_reserve(b2s_index* idx, int64_t n_rows);

/*
 * Replaces index.add(x) (tests/conftest.py:185) and the add loop inside
 * FAISSIndexBuilder.build_from_parquet (scripts/build_faiss_index.py:55-62).
 * rows: row-major [n, dim]; is_device != 0 means a device pointer on the
 * index's device.  Rows are stored as bf16 (round-to-nearest-even).
 */
B2S_API int b2s_add_f32(b2s_index* idx, const float* rows, int64_t n, int is_device);
B2S_API int b2s_add_bf16(b2s_index* idx, const void* rows, int64_t n, int is_device);
/*
 * Append rows VERBATIM (dtype B2S_DTYPE_F32 -> rounded to bf16, B2S_DTYPE_BF16 -> copied bit for
 * bit), without the cosine metric's normalisation: the rows are what the index is to hold.  This
 * is what FAISSIndexBuilder.load (src/serve/app.py:430-433) needs: rows saved by save() come back
 * bit-identical however often an index is saved and loaded.
 */
B2S_API int b2s_add_prepared(b2s_index* idx, const void* rows, int dtype, int64_t n, int is_device);

/* index.ntotal (scripts/build_faiss_index.py:72). */
B2S_API int64_t b2s_ntotal(const b2s_index* idx);
B2S_API int b2s_dim(const b2s_index* idx);
B2S_API int b2s_reset(b2s_index* idx); /* drop all rows, keep the allocation */

/* Global id of local row 0 (row-sharded corpus: SURVEY.md 8e). */
B2S_API int b2s_set_id_offset(b2s_index* idx, int64_t offset);

/*
 * Tuning / behaviour switches, by name:
 *   "path"        B2S_PATH_*           force a kernel family (default AUTO)
 *   "seed"        -1 auto | 0 off | 1 on   threshold-seeding pre-pass
 *   "scan_ctas_per_sm"  CTAs per SM of the scan kernel (default 2)
 *   "keep_f32"    1: keep an fp32 copy of added rows for exact re-scoring
 *   "rescore_pad" extra candidates re-scored in fp32 when keep_f32 is on
 *   "timing"      N > 0: every N-th search records CUDA events around its dominant kernel and the
 *                 whole call (b2s_last_stats / b2s_read_timings); 0 = off
 *   "pdl"         0 off | 1 (default) programmatic dependent launch between this handle's kernels
 *                 | 2 treat every device call as if B2S_SEARCH_STABLE_QUERIES were set (per handle)
 *   "pdl_early"   1 (default): the scan kernel triggers its dependent launch at its start, so the next search's
 *                 CTAs take over SM slots as this one's retire (no launch gap); 0 = after the scan
 *   "host_spin"   1 (default): a host-buffer call of 1-2 queries waits on a completion word the kernel writes
 *                 into mapped pinned memory after its last output (a few microseconds earlier than the stream
 *                 reports completion); 0 = cudaStreamSynchronize
 *   "grid_spare"  CTA slots a fused-tail scan launch leaves free for its predecessor's last CTA (default 1)
 *   "cascade_min_units" static iterations per warp below which the cascade select is not used (default 32)
 *   "prefetch_iters" scan kernel: iterations per warp prefetched into L2 before the PDL wait (default 6)
 *   "dynamic_tail"   scan kernel: units per CTA dealt by ticket at the end of the scan (default 3, 0 = static)
 *   "cascade"     1 (default): k <= 16 on large shards keeps the running global top-k in k sorted
 *                 device slots (lock-free atomicMax insertion) instead of per-CTA lists + merge
 *   "phase_a"     cascade: iterations against the CTA-local lists before the slots take over (0 = auto)
 *   "phase_a_stagger" cascade: CTA b switches b % this iterations later (default 64, capped by the shard size)
 *   "transition_mode" cascade A/B switch: 0 (default) CTA-wide phase switch behind a barrier, 1 warp by warp
 *   "peek_every"  cascade A/B switch: warp 0 re-reads the slots' k-th key every this many iterations (default 0 =
 *                 only after its own offers)
 *   "trace"       1: the scan kernel stamps %globaltimer at its phase boundaries (b2s_read_trace)
 *   "exchange_timeout_ms" bound of a sharded search's wait for a peer rank (default 10000)
 *   "tc_min_nq"   smallest query batch that takes the tensor path (default 3)
 *   "tc_sample_div"  threshold pre-pass samples 1/div of the corpus tiles (0 = chosen by k)
 *   "tc_chunk_tiles" tiles per work item when several query blocks share the corpus
 *   "tc_shared_thr"  1 (default): thresholds tighten GPU-wide through a per-query survivor histogram
 *   "tc_thr_period_ns" refresh period of the histogram's bound-updater warp (default 10000)
 *   "tc_single_cta"  1 (default): batches of <= 128 queries use single-CTA MMAs (M = 128)
 *   "fused_tail"  1 (default): the scan kernel's last CTA merges (and exchanges) in the same launch
 *   "exchange_ll" 1 (default): fused exchange uses sequence-tagged 8-byte words (no fence, no flag)
 *   "host_inline" 1 (default): host-buffer calls of 1-2 queries pass the query in the kernel
 *                 parameters and receive the answer in mapped pinned memory (no copy operations)
 */
B2S_API int b2s_set_option(b2s_index* idx, const char* name, int64_t value);
B2S_API int64_t b2s_get_option(const b2s_index* idx, const char* name);

/*
 * Replaces FAISSIndexBuilder.search(query_emb, k) -> (distances, indices)
 *   call site src/serve/app.py:293-295, consumer :299-317; faiss Index.search.
 * HOST buffers: queries fp32 row-major [nq, dim]; out_scores [nq, k] float32
 * descending; out_ids [nq, k] int64; unfilled slots (-FLT_MAX, -1).
 * Host<->device copies happen inside the call.
 */
B2S_API int b2s_search(b2s_index* idx, const float* queries, int64_t nq, int k, float* out_scores,
                       int64_t* out_ids);

/*
 * Same search with DEVICE buffers on the index's device, enqueued on
 * cuda_stream (a cudaStream_t; NULL = the legacy default stream).  q_dtype is
 * B2S_DTYPE_F32 or B2S_DTYPE_BF16; flags = 0 or B2S_SEARCH_*.  Returns after enqueueing; no host sync.
 * For a row-sharded corpus this is the per-shard "local top-k" whose outputs
 * feed the all-gather and b2s_merge_device.
 */
B2S_API int b2s_search_device(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k,
                              float* out_scores, int64_t* out_ids, void* cuda_stream, unsigned flags);

/*
 * Merge G sorted candidate lists per query into the global top-k (after the
 * NCCL all-gather of a row-sharded search).  scores/ids: device, [G, nq, k];
 * lists must be ordered by ascending id range (rank order) so that ties keep
 * the lower id.  Outputs device [nq, k].
 */
B2S_API int b2s_merge_device(int device, const float* scores, const int64_t* ids, int g, int64_t nq,
                             int k, float* out_scores, int64_t* out_ids, void* cuda_stream);

/*
 * Same merge for the PACKED candidate block one all-gather moves per rank:
 *   [ids int64 nq*k][scores float32 nq*k], padded to b2s_packed_bytes(nq, k).
 * packed: device [g][b2s_packed_bytes].  A per-shard search writes straight into the views
 * (ids at offset 0, scores at offset nq*k*8) of its own block.
 */
B2S_API int64_t b2s_packed_bytes(int64_t nq, int k);
B2S_API int b2s_merge_packed_device(int device, const void* packed, int g, int64_t nq, int k,
                                    float* out_scores, int64_t* out_ids, void* cuda_stream);

/*
 * Replaces StudentModel.compute_similarity(q, d) -> [nq, nd]
 *   tests/test_student_model.py:104-124; src/mining/miners.py:228-233;
 *   src/kd/eval.py:75.  HOST fp32 buffers, out [nq, nd] row-major.
 */
B2S_API int b2s_similarity(int device, const float* q, int64_t nq, const float* d, int64_t nd,
                           int dim, float* out);

/* Read rows [start, start+n) back as fp32 into a host buffer (save()). */
B2S_API int b2s_read_rows_f32(b2s_index* idx, int64_t start, int64_t n, float* out_host);

/* Device pointer of the bf16 corpus [ntotal, dim] (borrowed; valid until the next add). */
B2S_API const void* b2s_rows_device(const b2s_index* idx);

/*
 * Inner products of each query with chosen corpus rows: out[q, j] = <queries[q], row ids[q, j]>
 * (ids are global, i.e. include the id offset; ids outside this shard give -FLT_MAX).  DEVICE
 * buffers.  round_q_bf16 != 0 rounds the query to bf16 first (the tensor path's operand type).
 * Used to score a query's known positives for the ANCE margin
 * (src/mining/miners.py:228-236: pos_scores = compute_similarity(q, pos_embs); max()).
 */
B2S_API int b2s_score_rows_device(b2s_index* idx, const void* queries, int q_dtype, int round_q_bf16,
                                  int64_t nq, const int64_t* ids, int m, float* out_scores,
                                  void* cuda_stream);

/*
 * Replaces the selection loop of ANCEMiner.mine (src/mining/miners.py:237-247) for corpus-wide
 * candidates: per query keep, in order, the first top_k of its k_in sorted candidates
 * (cand_scores descending, cand_ids; -1 padded) that are not among its positives
 * (pos_ids [nq, n_pos], -1 padded) and whose score >= max(pos_scores) - margin (0.0 - margin when
 * a query has no positive).  DEVICE buffers; out_ids / out_scores [nq, top_k] padded with
 * (-1, -FLT_MAX); out_counts [nq] optional.
 */
B2S_API int b2s_ance_filter_device(int device, const float* cand_scores, const int64_t* cand_ids,
                                   int64_t nq, int k_in, const int64_t* pos_ids, const float* pos_scores,
                                   int n_pos, float margin, int top_k, int64_t* out_ids,
                                   float* out_scores, int32_t* out_counts, void* cuda_stream);

/*
 * Row-sharded search with the candidate exchange fused into the merge kernel (no NCCL on the data
 * path): every rank's merge CTA stores its local top-k straight into every peer's exchange buffer
 * over NVLink (peer-mapped memory + a release flag per (rank, query)), waits for the peers' flags
 * in its own memory and merges G*k candidates.  Realises the reference's "shard, fan out, merge"
 * prose (docs/operations/scaling-and-performance.md:154-172) on one 8-GPU box.
 *
 *   b2s_exchange_create   allocate this rank's buffer (slot_bytes >= b2s_packed_bytes(nq, k) of the
 *                         largest call, max_nq flags per rank); writes a 64-byte CUDA IPC handle
 *   b2s_exchange_connect  handles = world * 64 bytes, every rank's handle in rank order (exchanged
 *                         by the host over its own channel); raw_pointers != 0:
 *                         handles is instead an array of `world` device pointers (ranks emulated
 *                         inside one process: b2s_exchange_local of each)
 *   b2s_search_sharded_device   like b2s_search_device, outputs are the GLOBAL top-k (ids include
 *                         each shard's id offset).  All ranks must issue the same sequence of
 *                         calls with the same (nq, k).  phase 0 = whole call; phase 1 = local
 *                         search + push only and phase 2 = wait + merge only (lets a test run
 *                         several emulated ranks one after the other on a single GPU).
 *   b2s_exchange_status   0, or the sequence number of a call whose wait timed out (synchronises).
 *                         A timeout is never silent: the host-buffer call that hit it returns
 *                         B2S_ERR_CUDA, and so does every later sharded call on the handle.
 * world * k is not limited by the merge buffer: above 4096 candidates per query the ranks' sorted
 * blocks are merged by binary search (8 ranks x k = 2048 works).  k itself is <= 2048 everywhere
 * (the reference's serving schema allows <= 100 / 200, src/serve/schemas.py:12-16; its
 * SearchConfig.max_top_k = 10000, src/config.py:227, is refused with B2S_ERR_UNSUPPORTED).
 */
B2S_API int b2s_exchange_create(b2s_index* idx, int world, int rank, int64_t slot_bytes, int max_nq,
                                void* ipc_handle_out);
B2S_API void* b2s_exchange_local(const b2s_index* idx);
B2S_API int b2s_exchange_connect(b2s_index* idx, const void* handles, int raw_pointers);
B2S_API int b2s_exchange_status(b2s_index* idx);
B2S_API int b2s_search_sharded_device(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k,
                                      float* out_scores, int64_t* out_ids, void* cuda_stream, int phase,
                                      unsigned flags);
/* HOST-buffer variant of the whole sharded call (phase 0): H2D + exchange + D2H inside; every rank
 * gets the global top-k.  This is what ShardedFlatIPIndex.search(np.ndarray, k) calls. */
B2S_API int b2s_search_sharded(b2s_index* idx, const float* queries, int64_t nq, int k, float* out_scores,
                               int64_t* out_ids);

/*
 * Replaces maxsim_aggregation (src/utils/chunk.py:123-148: max chunk score per document) for search
 * results: scores/ids [nq, k_in] are a query's chunk hits sorted by descending score; chunk_to_doc
 * [n_chunks] maps a chunk (row) id to its document id.  Keeps each document's best (= first) hit,
 * in order: out_scores / out_doc_ids [nq, k_out] padded with (-FLT_MAX, -1); out_counts optional.
 * DEVICE buffers.
 */
B2S_API int b2s_maxsim_device(int device, const float* scores, const int64_t* ids, int64_t nq, int k_in,
                              const int64_t* chunk_to_doc, int64_t n_chunks, int k_out, float* out_scores,
                              int64_t* out_doc_ids, int32_t* out_counts, void* cuda_stream);

B2S_API int b2s_last_stats(const b2s_index* idx, b2s_stats* out);

/*
 * With option "timing" = 1 every search records CUDA events around its dominant kernel(s) and
 * around the whole call on the stream it runs on.  After synchronising that stream, read the
 * last (up to max_n, ring of 4096) calls' device times in ms, oldest first.  Returns the count.
 */
B2S_API int b2s_read_timings(const b2s_index* idx, float* dominant_ms, float* total_ms, int max_n);

/*
 * With option "trace" = 1 the scan kernel of a batch-1/2 search stamps the GPU-wide nanosecond timer:
 * word 1 = the last CTA to finish holds the ticket (every CTA's scan is done), 2 = the local top-k is
 * ready, 3 = pushed to the peer ranks, 4 = every peer's candidates have arrived, 5 = outputs written,
 * word 0 = G (CTAs); then six arrays of 512 words, entry b = CTA b: start (after the dependency
 * wait), end of scan, cascade transition begin / end, end of the static part, SM id.
 * Two such blocks are kept and used alternately (consecutive launches may overlap): the block of the LAST
 * traced launch comes first, then that of the launch before it.
 * Synchronises the device and copies up to max_words words (2 * (16 + 6 * 512) for all); returns the count.
 */
B2S_API int b2s_read_trace(b2s_index* idx, uint64_t* out, int max_words);

#ifdef __cplusplus
}
#endif
#endif /* B200SEARCH_H */

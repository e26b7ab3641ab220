// b2s_api.cu -- C-ABI entry points of libb200search.so (see include/b200search.h).
//
// Host side of the exact inner-product top-k path: owns the bf16 corpus shard in HBM, sizes the
// workspace, picks the kernel family (K1 scan for tiny batches, K2 tcgen05 GEMM for real
// batches), launches the final per-query merge (K3) and moves host buffers for the host API.
// There is NO CPU fallback: without an sm_100 device every compute entry point returns
// B2S_ERR_NO_DEVICE.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define B2S_CPU_RELAX() _mm_pause()
#else
#define B2S_CPU_RELAX() ((void)0)
#endif

#include "../../include/b200search.h"
#include "ance_filter.cuh"
#include "exchange.cuh"
#include "merge_topk.cuh"
#include "rescore.cuh"
#include "scan_topk.cuh"
#include "select.cuh"
#include "util_kernels.cuh"
#ifndef B2S_NO_TENSOR_PATH
#include "gemm_topk_tc.cuh"
#endif

using namespace b2s;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            int _code = (_e == cudaErrorMemoryAllocation) ? B2S_ERR_NOMEM : B2S_ERR_CUDA;        \
            if (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) _code = B2S_ERR_NO_DEVICE; \
            return fail(_code, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
        }                                                                                       \
    } while (0)

constexpr int kMaxK = 2048;
constexpr int kScanQueryChunk = 64;   // queries per workspace round on the scan path
constexpr int kSeedUnitStride = 64;   // the seeding pre-pass reads every 64th unit (~1.6% of rows)
constexpr int kTimingSlots = 4096;    // search calls remembered by the timing ring

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return B2S_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, need);
            want = need;
        }
        if (e != cudaSuccess) {
            p = nullptr;
            cudaGetLastError();
            return fail(B2S_ERR_NOMEM, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e));
        }
        bytes = want;
        return B2S_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

}  // namespace

struct b2s_index {
    int dim = 0;
    int metric = 0;
    int device = 0;
    int num_sms = 0;
    __nv_bfloat16* rows = nullptr;  // [cap_rows, dim]
    float* rows_f32 = nullptr;      // optional fp32 copy (keep_f32)
    int64_t n = 0;
    int64_t cap_rows = 0;
    int64_t id_offset = 0;
    // options
    int opt_path = B2S_PATH_AUTO;
    int opt_seed = -1;
    int opt_scan_ctas_per_sm = 2;
    int opt_keep_f32 = 0;
    int opt_rescore_pad = 32;
    int opt_timing = 0;           // N > 0: every N-th search call records CUDA events
    bool time_this = false;       // the current call is one of them
    int64_t timing_seq = 0;
    int opt_tc_min_nq = 3;
    int opt_pdl = 1;              // 1: programmatic dependent launch hides launch latencies (always safe);
                                  // 2: also overlap the scan of call i+1 with the merge of call i -- only
                                  //    valid when the query buffer is not written by the kernel enqueued
                                  //    immediately before the search on the same stream
    int opt_tc_sample_div = 0;    // the threshold pre-pass samples 1 / this of the full tiles (0 = by k)
    int opt_tc_shared_thr = 1;    // tensor path: tighten thresholds through a per-query survivor histogram
    int opt_tc_thr_period_ns = 10000;   // refresh period of the bound-updater warp
    int opt_tc_single_cta = 1;    // tensor path: single-CTA MMAs (M = 128) for batches of <= 128 queries
    int opt_tc_chunk_lo = 48;     // tiles per work item when several query blocks share the corpus
    int opt_tc_chunk_hi = 96;
    // workspace
    DevBuf ws_lists, ws_counts, ws_thr, ws_gmax, ws_hist, ws_hcfg, ws_rs_scores, ws_rs_ids, ws_seed, ws_qf32, ws_qbf16, ws_io_q, ws_io_ids, ws_tmp;
    void* pin_q = nullptr;
    // host ingest (b2s_add_* with host rows): two pinned staging buffers + their "DMA done" events
    void* pin_in[2] = {nullptr, nullptr};
    cudaEvent_t pin_in_ev[2] = {nullptr, nullptr};
    void* pin_out = nullptr;
    size_t pin_q_bytes = 0, pin_out_bytes = 0;
    cudaStream_t stream = nullptr;
    // timing ring: per search call, events [total begin, dominant begin, dominant end, total end]
    cudaEvent_t* ring = nullptr;   // kTimingSlots * 4 events, created when "timing" is switched on
    cudaEvent_t* ev = nullptr;     // the 4 events of the current call
    int64_t ring_calls = 0;        // search calls recorded since timing was switched on
    bool ev_valid = false;
    b2s_stats stats;
    std::mutex mu;
    // cross-GPU candidate exchange (exchange.cuh); ex_call != nullptr only inside a sharded search
    struct Exchange {
        unsigned char* local = nullptr;
        size_t bytes = 0;
        int world = 0, rank = 0, max_nq = 0;
        long long slot_stride = 0, flags_off = 0, ll_off = 0, ll_entries = 0;
        unsigned char** peers_dev = nullptr;   // device array [world]
        std::vector<void*> opened;             // cudaIpcOpenMemHandle'd peer mappings
        unsigned* status = nullptr;            // mapped pinned host word: the kernels store the sequence number of
                                               // a call whose wait timed out, the host reads it without a copy
        unsigned seq = 0;
        bool connected = false;
        int max_fused_nq = 0;                  // CTAs of the fused merge+exchange kernel that are co-resident
        long long timeout_ms = 10000;
    } ex;
    const ExchangeArgs* ex_call = nullptr;
    int ex_fused = 0;
    // device control block of the scan kernel: done ticket, dynamic-tail work counters, the cascade select's
    // global slots [kScanMaxNQ][kCascadeMaxK] u64 -- all zero between launches
    ScanCtl* ctl = nullptr;             // two sets, used alternately by consecutive scan launches
    unsigned ctl_launches = 0;          // counted scan launches so far: launch n uses set n & 1 and expects epoch n >> 1
    bool ctl_prev_counted = false;      // the previous scan launch took part in the epoch protocol (not graph-captured)
    unsigned long long* trace_buf = nullptr;   // option "trace": globaltimer stamps of the last scan launch
    int opt_trace = 0;
    unsigned trace_seq = 0;                    // the buffer has two halves used alternately (overlapping launches)
    int opt_prefetch_iters = 6;         // iterations per warp prefetched into L2 before the PDL wait (0 = off)
    int opt_dynamic_tail = 3;           // units per CTA dealt dynamically at the end of the scan (0 = all static)
    int opt_cascade = 1;                // k <= 16 on large shards: global sorted slots instead of per-CTA lists
    int opt_phase_a = 0;                // cascade: iterations against the local lists first (0 = auto)
    int opt_phase_a_stagger = 64;       // cascade: CTA b switches to the global slots b % this iterations later
    int opt_transition_mode = 0;        // cascade: see ScanParams::transition_mode
    int opt_grid_spare = 1;             // CTA slots left free by a fused-tail scan launch
    int opt_cascade_min_units = 32;     // static iterations per warp below which the cascade select is not used
    int opt_pdl_early = 1;              // scan kernel triggers its dependent launch at its start (see ScanParams::early_trigger)
    int opt_peek_every = 0;             // cascade: iterations between re-reads of the slots' k-th key (0 = only after inserts)
    unsigned call_flags = 0;            // B2S_SEARCH_* of the call in flight
    // stream of the previous search: a call on another stream first waits for everything enqueued there
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    cudaEvent_t xs_event = nullptr;
    const float* host_q = nullptr;      // host-buffer call in flight: its queries may ride in the kernel parameters
    unsigned* host_flag = nullptr;      // ... and this mapped pinned word may be used as its completion flag
    unsigned host_calls = 0;            // tiny host calls so far (flag values are never reused)
    unsigned host_seq = 0;              // value the flag takes when the call in flight has written its outputs
    bool host_flag_armed = false;       // a launch of the call in flight carries the flag
    int opt_fused_tail = 1;
    int opt_host_spin = 1;              // tiny host calls: wait on the kernel's completion flag instead of the stream
    int opt_host_inline = 1;            // host-buffer calls of 1-2 queries: query in the kernel parameters, answer
                                        // written straight to pinned host memory (no copy-engine operations)
    int opt_exchange_ll = 1;            // fused exchange: tagged 8-byte words instead of payload + fence + flag
#ifndef B2S_NO_TENSOR_PATH
    TensorPathState tc;
#endif
};

namespace {

int use_device(const b2s_index* idx) {
    CUDA_TRY(cudaSetDevice(idx->device));
    return B2S_OK;
}

int grow_rows(b2s_index* idx, int64_t need_rows) {
    if (need_rows <= idx->cap_rows) return B2S_OK;
    int64_t new_cap = std::max<int64_t>(need_rows, idx->cap_rows + idx->cap_rows / 2);
    __nv_bfloat16* nr = nullptr;
    const size_t row_bytes = (size_t)idx->dim * sizeof(__nv_bfloat16);
    cudaError_t e = cudaMalloc((void**)&nr, (size_t)new_cap * row_bytes);
    if (e != cudaSuccess && new_cap > need_rows) {
        cudaGetLastError();
        new_cap = need_rows;
        e = cudaMalloc((void**)&nr, (size_t)new_cap * row_bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B2S_ERR_NOMEM, std::string("cudaMalloc corpus: ") + cudaGetErrorString(e));
    }
    if (idx->n > 0) CUDA_TRY(cudaMemcpy(nr, idx->rows, (size_t)idx->n * row_bytes, cudaMemcpyDeviceToDevice));
    if (idx->rows) cudaFree(idx->rows);
    idx->rows = nr;
    if (idx->opt_keep_f32) {
        float* nf = nullptr;
        e = cudaMalloc((void**)&nf, (size_t)new_cap * idx->dim * sizeof(float));
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B2S_ERR_NOMEM, std::string("cudaMalloc fp32 corpus: ") + cudaGetErrorString(e));
        }
        if (idx->n > 0 && idx->rows_f32)
            CUDA_TRY(cudaMemcpy(nf, idx->rows_f32, (size_t)idx->n * idx->dim * sizeof(float),
                                cudaMemcpyDeviceToDevice));
        if (idx->rows_f32) cudaFree(idx->rows_f32);
        idx->rows_f32 = nf;
    }
    idx->cap_rows = new_cap;
#ifndef B2S_NO_TENSOR_PATH
    idx->tc.corpus_map_valid = false;
#endif
    return B2S_OK;
}

// ---------------------------------------------------------------------------------------------
// K1 dispatch
// ---------------------------------------------------------------------------------------------

// Launch with programmatic dependent launch allowed: the kernel may begin while its predecessor in
// the stream is still running and synchronises on it with grid_dep_wait() (select.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

std::mutex g_attr_mu;   // guards the once-per-device function attribute flags below

template <int CPL, int NQ, int U>
int launch_scan_t(const ScanParams& p, int grid, cudaStream_t s, bool pdl) {
    const size_t smem = (size_t)NQ * p.cap * sizeof(u64);
    auto kern = scan_topk_kernel<CPL, NQ, U>;
    // dynamic candidate lists sit next to ~44 KB of static shared memory (the fused merge tail): opt in
    // once per instantiation for the largest list set (k = 2048: NQ * 4096 keys)
    static bool attr_set[64] = {};   // per device: function attributes belong to the device's context
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64) {
        std::lock_guard<std::mutex> g(g_attr_mu);
        if (!attr_set[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, NQ * 4096 * (int)sizeof(u64)));
            attr_set[dev] = true;
        }
    }
    CUDA_TRY(launch_pdl(pdl, kern, dim3(grid), dim3(kScanThreads), smem, s, p));
    return B2S_OK;
}

// Largest register-resident query count the scan kernel is built for at this dim.
int scan_max_nq(int dim) {
    const int cpl = dim / 128;
    if (cpl <= 3) return 4;
    if (cpl == 4) return 2;
    return 1;
}

int scan_rows_per_iter(int dim) { return (dim / 128 <= 3) ? 8 : 4; }

int launch_scan(int dim, int nq_group, const ScanParams& p, int grid, cudaStream_t s, bool pdl) {
    const int cpl = dim / 128;
#define B2S_SCAN_CASE(C, Q, UU) \
    if (cpl == C && nq_group == Q) return launch_scan_t<C, Q, UU>(p, grid, s, pdl);
    B2S_SCAN_CASE(1, 1, 4) B2S_SCAN_CASE(1, 2, 4) B2S_SCAN_CASE(1, 4, 4)
    B2S_SCAN_CASE(2, 1, 4) B2S_SCAN_CASE(2, 2, 4) B2S_SCAN_CASE(2, 4, 4)
    B2S_SCAN_CASE(3, 1, 4) B2S_SCAN_CASE(3, 2, 4) B2S_SCAN_CASE(3, 4, 4)
    B2S_SCAN_CASE(4, 1, 2) B2S_SCAN_CASE(4, 2, 2)
    B2S_SCAN_CASE(6, 1, 2)
    B2S_SCAN_CASE(8, 1, 2)
#undef B2S_SCAN_CASE
    return fail(B2S_ERR_UNSUPPORTED, "scan kernel: unsupported (dim, query group)");
}

bool dim_supported(int dim) {
    return dim == 128 || dim == 256 || dim == 384 || dim == 512 || dim == 768 || dim == 1024;
}

// Final (or seeding) per-query merge of `nq` queries starting at query q_offset of the call.  Inside a
// sharded search the final merge also pushes the local top-k to every rank (and, fused, waits and
// merges the ranks' candidates): exchange.cuh.
int launch_merge(const b2s_index* idx, const MergeParams& mp, int nq, int q_offset, cudaStream_t s) {
    if (mp.num_lists > kMergeMaxLists) return fail(B2S_ERR_UNSUPPORTED, "too many candidate lists per query");
    if (idx->ex_call != nullptr && mp.out_kth_key == nullptr) {
        ExchangeArgs ex = *idx->ex_call;
        ex.q_offset = q_offset;
        const bool pdl = idx->opt_pdl != 0;
        if (idx->ex_fused) CUDA_TRY(launch_pdl(pdl, merge_exchange_kernel<true>, dim3(nq), dim3(kMergeThreads), 0, s, mp, ex));
        else CUDA_TRY(launch_pdl(pdl, merge_exchange_kernel<false>, dim3(nq), dim3(kMergeThreads), 0, s, mp, ex));
    } else {
        CUDA_TRY(launch_pdl(idx->opt_pdl != 0, merge_topk_kernel, dim3(nq), dim3(kMergeThreads), 0, s, mp));
    }
    return B2S_OK;
}

// Scan path for queries [0, nq) already in fp32 on the device.
int search_scan(b2s_index* idx, const float* q_f32, int64_t nq, int k, float* out_scores,
                int64_t* out_ids, cudaStream_t s, bool seed, bool late_wait_ok, const float* inline_q) {
    const int cap = list_capacity(k);
    const int rpi = scan_rows_per_iter(idx->dim);
    const int unit = rpi * kScanWarps;
    const int max_group = scan_max_nq(idx->dim);
    // One scan launch covers the whole call and nothing is seeded: the last CTA of the scan produces the
    // final top-k (and, sharded with co-resident CTAs, the exchange) itself -- see scan_topk.cuh.
    const bool fuse_tail = idx->opt_fused_tail && !seed && nq <= max_group && idx->ctl != nullptr &&
                           (idx->ex_call == nullptr || idx->ex_fused);
    int grid = idx->num_sms * std::max(1, idx->opt_scan_ctas_per_sm);
    // one CTA slot of the device stays free: the last CTA of a launch is still busy with the tail (read-out,
    // NVLink exchange) when the next launch's CTAs move in, and none of those has to wait for its slot
    if (fuse_tail && grid > idx->num_sms) grid -= std::max(0, std::min(idx->opt_grid_spare, idx->num_sms));
    const int64_t units = (idx->n + unit - 1) / unit;
    if ((int64_t)grid > units) grid = (int)units;
    grid = std::min(grid, kMergeMaxLists);
    const bool pdl = idx->opt_pdl != 0;
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap_status) != cudaSuccess) cudaGetLastError();
    const bool capturing = cap_status != cudaStreamCaptureStatusNone;

    const int chunk = (int)std::min<int64_t>(nq, kScanQueryChunk);
    int rc;
    if ((rc = idx->ws_lists.ensure((size_t)grid * chunk * cap * sizeof(u64))) != B2S_OK) return rc;
    if ((rc = idx->ws_counts.ensure((size_t)grid * chunk * sizeof(int))) != B2S_OK) return rc;
    if (seed && (rc = idx->ws_seed.ensure((size_t)chunk * sizeof(u64))) != B2S_OK) return rc;

    // Dynamic tail: the first S units of every CTA are static, the rest of the shard is dealt by ticket.
    const int64_t full_units = idx->n / unit;
    const int64_t static_units = std::max<int64_t>(0, full_units / grid - idx->opt_dynamic_tail);   // S
    const bool dyn_ok = idx->opt_dynamic_tail > 0 && idx->ctl != nullptr;
    const long long dyn_begin = dyn_ok ? (long long)(static_units * grid * unit) : (long long)idx->n;
    // Cascade select: small k, one fused launch, enough static iterations for a meaningful phase A.
    const bool cascade = idx->opt_cascade && fuse_tail && dyn_ok && k <= kCascadeMaxK &&
                         static_units >= idx->opt_cascade_min_units;
    for (int64_t c0 = 0; c0 < nq; c0 += chunk) {
        const int cn = (int)std::min<int64_t>(chunk, nq - c0);
        for (int pass = seed ? 0 : 1; pass < 2; ++pass) {
            for (int g0 = 0; g0 < cn;) {
                int group = 1;
                while (group * 2 <= max_group && g0 + group * 2 <= cn) group *= 2;
                ScanParams p;
                memset(&p, 0, offsetof(ScanParams, q_inline));
                p.corpus = reinterpret_cast<const uint4*>(idx->rows);
                p.queries = q_f32 + (size_t)c0 * idx->dim;
                p.n_rows = idx->n;
                p.q_begin = g0;
                p.nq_valid = group;
                p.k = k;
                p.cap = cap;
                p.unit_stride = pass == 0 ? kSeedUnitStride : 1;
                p.seed_keys = (pass == 1 && seed) ? reinterpret_cast<const u64*>(idx->ws_seed.p) : nullptr;
                p.lists = reinterpret_cast<u64*>(idx->ws_lists.p);
                p.counts = reinterpret_cast<int*>(idx->ws_counts.p);
                p.nq_lists = chunk;
                // The scan only READS the corpus and the caller's queries unless a kernel of THIS call
                // ran before it (query prep, seeding pass, an earlier group writing the same workspace).
                p.pdl_late_wait = (late_wait_ok && !seed && nq <= max_group) ? 1 : 0;
                p.dyn_begin = pass == 1 ? dyn_begin : (long long)idx->n;   // the sampling pre-pass is all static
                // control set: launches that use one (fused tail or dynamic tail) alternate between the two sets
                const bool uses_ctl = idx->ctl != nullptr && (fuse_tail || p.dyn_begin < (long long)idx->n);
                if (idx->ctl != nullptr) {
                    p.ctl = idx->ctl + (idx->ctl_launches & 1u);
                    p.ctl_expect = idx->ctl_launches >> 1;
                    p.ctl_bump = (uses_ctl && !capturing) ? 1 : 0;
                }
                // overlap with the predecessor only inside an unbroken chain of counted launches
                if (!idx->ctl_prev_counted || capturing || !uses_ctl) p.pdl_late_wait = 0;
                p.prefetch_iters = pass == 1 ? std::min(32, std::max(0, idx->opt_prefetch_iters)) : 0;
                p.trace = (idx->opt_trace && pass == 1)
                              ? idx->trace_buf + (size_t)(idx->trace_seq++ & 1u) * (kTraceWords + kTraceArrays * kTraceStride)
                              : nullptr;
                // early trigger unless this launch would park its CTAs on a full-grid wait in the middle of the scan
                p.early_trigger = (idx->opt_pdl_early && (!p.pdl_late_wait || cascade)) ? 1 : 0;
                if (inline_q != nullptr && fuse_tail && idx->host_flag != nullptr && g0 + group >= cn && c0 + cn >= nq) {
                    p.host_flag = idx->host_flag;   // the launch that writes the call's last outputs
                    p.host_seq = idx->host_seq;
                    idx->host_flag_armed = true;
                }
                if (inline_q != nullptr) {
                    p.use_inline = 1;
                    memcpy(p.q_inline, inline_q, (size_t)nq * idx->dim * sizeof(float));
                }
                if (fuse_tail) {
                    p.fused_tail = idx->ex_call ? 2 : 1;
                    p.mp.lists = p.lists;
                    p.mp.counts = p.counts;
                    p.mp.num_lists = grid;
                    p.mp.nq_lists = chunk;
                    p.mp.lists_sorted = 1;
                    p.mp.cap = cap;
                    p.mp.k = k;
                    p.mp.id_offset = idx->id_offset;
                    p.mp.out_scores = out_scores;
                    p.mp.out_ids = reinterpret_cast<long long*>(out_ids);
                    if (idx->ex_call) p.ex = *idx->ex_call;
                    if (cascade) {
                        p.select_mode = 1;
                        p.peek_every = idx->opt_peek_every;
                        p.transition_mode = idx->opt_transition_mode;
                        // phase A ends (and, with stable queries, the PDL wait sits) after a0 + blockIdx % stagger
                        // iterations: early enough that the last CTA switches over in the first half of the scan
                        int a = idx->opt_phase_a > 0 ? idx->opt_phase_a : std::max(2, std::min((int)(static_units / 8), 8));
                        p.phase_a_iters = (int)std::min<int64_t>(a, static_units);
                        p.phase_a_stagger = (int)std::max<int64_t>(1, std::min<int64_t>(idx->opt_phase_a_stagger,
                                                                                         static_units / 2 - p.phase_a_iters));
                    }
                }
                if ((rc = launch_scan(idx->dim, group, p, grid, s, pdl)) != B2S_OK) return rc;
                if (p.ctl_bump) ++idx->ctl_launches;
                idx->ctl_prev_counted = p.ctl_bump != 0;
                idx->stats.kernel_launches++;
                if (pass == 1) idx->stats.passes++;
                g0 += group;
            }
            if (fuse_tail) {
                if (idx->time_this) cudaEventRecord(idx->ev[2], s);
                continue;   // finished (and exchanged) by the scan kernel's last CTA
            }
            MergeParams mp;
            memset(&mp, 0, sizeof(mp));
            mp.lists = reinterpret_cast<const u64*>(idx->ws_lists.p);
            mp.counts = reinterpret_cast<const int*>(idx->ws_counts.p);
            mp.num_lists = grid;
            mp.nq_lists = chunk;
            mp.lists_sorted = 1;   // scan_topk_kernel publishes sorted lists
            mp.cap = cap;
            mp.k = k;
            mp.id_offset = idx->id_offset;
            if (pass == 0) {
                mp.out_kth_key = reinterpret_cast<u64*>(idx->ws_seed.p);
            } else {
                mp.out_scores = out_scores + (size_t)c0 * k;
                mp.out_ids = reinterpret_cast<long long*>(out_ids) + (size_t)c0 * k;
            }
            if (pass == 1 && idx->time_this && c0 == 0) cudaEventRecord(idx->ev[2], s);
            if ((rc = launch_merge(idx, mp, cn, (int)c0, s)) != B2S_OK) return rc;
            idx->stats.kernel_launches++;
        }
    }
    return B2S_OK;
}

#ifndef B2S_NO_TENSOR_PATH
int search_tensor(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                  int64_t* out_ids, cudaStream_t s, bool seed, bool normalize);
void tensor_path_release(b2s_index* idx);
#endif

// kernel family of a search
int choose_path(const b2s_index* idx, int64_t nq, int k) {
    int path = idx->opt_path;
#ifdef B2S_NO_TENSOR_PATH
    path = B2S_PATH_SCAN;
#else
    // batches go to the tensor path; so do 1-2 queries with a large k (the scan kernel's shared-memory lists
    // are tuned for small k: 3.0 ms at k = 1000 vs 1.2 ms on the single-CTA tensor variant)
    if (path == B2S_PATH_AUTO) path = (nq >= idx->opt_tc_min_nq || k >= 256) ? B2S_PATH_TENSOR : B2S_PATH_SCAN;
    if (path == B2S_PATH_TENSOR && !tensor_path_supported(idx->dim)) path = B2S_PATH_SCAN;
#endif
    return path;
}

int search_core(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                int64_t* out_ids, cudaStream_t s) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "nq and k must be >= 0");
    if (nq == 0 || k == 0) return B2S_OK;
    if (!queries || !out_scores || !out_ids) return fail(B2S_ERR_INVALID, "null buffer");
    if (q_dtype != B2S_DTYPE_F32 && q_dtype != B2S_DTYPE_BF16) return fail(B2S_ERR_INVALID, "bad q_dtype");
    if (k > kMaxK) return fail(B2S_ERR_UNSUPPORTED, "k > 2048 is not supported by the fused select");
    if (nq > (int64_t)1 << 24) return fail(B2S_ERR_UNSUPPORTED, "nq too large for one call");
    int rc;
    if ((rc = use_device(idx)) != B2S_OK) return rc;

    memset(&idx->stats, 0, sizeof(idx->stats));
    idx->stats.corpus_bytes = idx->n * (int64_t)idx->dim * 2;
    // "timing" = N: every N-th search records its events (events between kernels switch programmatic
    // dependent launch off for that call, so sampling keeps the steady state undisturbed)
    idx->time_this = idx->opt_timing > 0 && idx->ring && (idx->timing_seq++ % idx->opt_timing) == 0;
    if (idx->time_this) {
        idx->ev = idx->ring + 4 * (idx->ring_calls % kTimingSlots);
        idx->ring_calls++;
        cudaEventRecord(idx->ev[0], s);
        idx->ev_valid = true;
    } else {
        idx->ev_valid = false;
    }

    if (idx->n == 0 && idx->ex_call != nullptr) {
        // empty shard of a sharded search: contribute "no candidates" for every query
        MergeParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.k = k;
        mp.cap = 64;
        mp.nq_lists = (int)nq;
        mp.out_scores = out_scores;
        mp.out_ids = reinterpret_cast<long long*>(out_ids);
        if ((rc = launch_merge(idx, mp, (int)nq, 0, s)) != B2S_OK) return rc;
        idx->stats.kernel_launches++;
        if (idx->time_this) {
            cudaEventRecord(idx->ev[1], s);
            cudaEventRecord(idx->ev[2], s);
            cudaEventRecord(idx->ev[3], s);
        }
        return B2S_OK;
    }
    if (idx->n == 0) {
        const long long cnt = (long long)nq * k;
        fill_empty_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(out_scores, reinterpret_cast<long long*>(out_ids), cnt);
        CUDA_TRY(cudaGetLastError());
        idx->stats.kernel_launches++;
        if (idx->time_this) {
            cudaEventRecord(idx->ev[1], s);
            cudaEventRecord(idx->ev[2], s);
            cudaEventRecord(idx->ev[3], s);
        }
        return B2S_OK;
    }

    const int path = choose_path(idx, nq, k);
    idx->stats.path = path;
    const bool normalize = idx->metric == B2S_METRIC_COSINE;
    // Threshold seeding: on the scan path a pre-pass over every 64th unit pays for itself once the
    // per-CTA cold start matters (nq * k > 128 on a large shard); the tensor path always bounds the
    // k-th best score from a row sample first (its lists are private to one thread).
    // (measured with the fused merge tail, 8.8M rows: one unseeded launch wins up to nq * k = 128 --
    //  k = 100 at batch 1: 0.984 ms vs 0.993 ms seeded; k = 200: 1.037 vs 1.010)
    bool seed = idx->opt_seed == 1 ||
                (idx->opt_seed < 0 && (path == B2S_PATH_TENSOR || (idx->n >= (int64_t)1 << 20 && nq * k > 128)));
    idx->stats.seeded = seed ? 1 : 0;

    if (path == B2S_PATH_SCAN) {
        const float* qf = reinterpret_cast<const float*>(queries);
        // host-buffer call of 1-2 queries: they travel in the scan kernel's parameters (normalised here for cosine)
        float inline_buf[kScanInlineFloats];
        const float* inline_q = nullptr;
        if (idx->host_q != nullptr && q_dtype == B2S_DTYPE_F32 && nq * idx->dim <= kScanInlineFloats &&
            nq <= scan_max_nq(idx->dim)) {
            inline_q = idx->host_q;
            if (normalize) {
                for (int64_t qi = 0; qi < nq; ++qi) {
                    const float* src = idx->host_q + qi * idx->dim;
                    float ss = 0.f;
                    for (int i = 0; i < idx->dim; ++i) ss = fmaf(src[i], src[i], ss);
                    const float scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
                    for (int i = 0; i < idx->dim; ++i) inline_buf[qi * idx->dim + i] = src[i] * scale;
                }
                inline_q = inline_buf;
            }
        }
        if (inline_q == nullptr && (q_dtype != B2S_DTYPE_F32 || normalize)) {
            if ((rc = idx->ws_qf32.ensure((size_t)nq * idx->dim * sizeof(float))) != B2S_OK) return rc;
            const int warps = 8;
            prep_queries_kernel<<<(unsigned)((nq + warps - 1) / warps), warps * 32, 0, s>>>(
                queries, q_dtype == B2S_DTYPE_BF16, nq, nq, idx->dim, normalize ? 1 : 0,
                reinterpret_cast<float*>(idx->ws_qf32.p), nullptr);
            CUDA_TRY(cudaGetLastError());
            idx->stats.kernel_launches++;
            qf = reinterpret_cast<const float*>(idx->ws_qf32.p);
        }
        if (idx->time_this) cudaEventRecord(idx->ev[1], s);
        // Late PDL wait (the scan overlaps the tail of the previous kernel of the stream) only when nothing the
        // scan reads early can have been produced by that kernel: the query rides in the parameters (host
        // call), or the caller vouches for its buffer (B2S_SEARCH_STABLE_QUERIES / option pdl = 2) and no
        // kernel of this call precedes the scan -- see scan_topk.cuh
        const bool stable = (idx->call_flags & B2S_SEARCH_STABLE_QUERIES) != 0 || idx->opt_pdl == 2;
        const bool late_wait_ok = idx->opt_pdl != 0 &&
                                  (inline_q != nullptr || (stable && qf == reinterpret_cast<const float*>(queries)));
        rc = search_scan(idx, qf, nq, k, out_scores, out_ids, s, seed, late_wait_ok, inline_q);
        if (rc != B2S_OK) return rc;
    } else {
#ifndef B2S_NO_TENSOR_PATH
        rc = search_tensor(idx, queries, q_dtype, nq, k, out_scores, out_ids, s, seed, normalize);
        if (rc != B2S_OK) return rc;
#endif
    }
    if (idx->time_this) cudaEventRecord(idx->ev[3], s);
    return B2S_OK;
}

// Search, then (option keep_f32) re-rank k + pad bf16 candidates with the fp32 copy of the rows: rescore.cuh.
int search_impl(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                int64_t* out_ids, cudaStream_t s) {
    const bool rescore = idx && idx->opt_keep_f32 && idx->rows_f32 && idx->n > 0 && idx->ex_call == nullptr &&
                         nq > 0 && k > 0 && k <= kMaxK && queries && out_scores && out_ids;
    if (!rescore) return search_core(idx, queries, q_dtype, nq, k, out_scores, out_ids, s);
    const int k2 = std::min(kRescoreCap, k + idx->opt_rescore_pad);
    int rc;
    if ((rc = use_device(idx)) != B2S_OK) return rc;
    if ((rc = idx->ws_rs_scores.ensure((size_t)nq * k2 * sizeof(float))) != B2S_OK) return rc;
    if ((rc = idx->ws_rs_ids.ensure((size_t)nq * k2 * sizeof(int64_t))) != B2S_OK) return rc;
    float* cs = reinterpret_cast<float*>(idx->ws_rs_scores.p);
    int64_t* ci = reinterpret_cast<int64_t*>(idx->ws_rs_ids.p);
    if ((rc = search_core(idx, queries, q_dtype, nq, k2, cs, ci, s)) != B2S_OK) return rc;
    rescore_f32_kernel<<<(unsigned)nq, kRescoreThreads, 0, s>>>(
        idx->rows_f32, idx->n, idx->dim, queries, q_dtype == B2S_DTYPE_BF16, idx->metric == B2S_METRIC_COSINE ? 1 : 0, cs,
        reinterpret_cast<const long long*>(ci), k2, idx->id_offset, k, out_scores, reinterpret_cast<long long*>(out_ids));
    CUDA_TRY(cudaGetLastError());
    idx->stats.kernel_launches++;
    if (idx->time_this && idx->ev_valid) cudaEventRecord(idx->ev[3], s);
    return B2S_OK;
}

// The workspaces (lists, counters, slots) belong to the handle, not to a stream: a search enqueued on
// another stream than the previous one first waits for everything enqueued on that stream so far.
int guard_stream(b2s_index* idx, cudaStream_t s) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) cudaGetLastError();
    // A search captured into a CUDA graph is ordered by whoever launches the graph: nothing outside the
    // capture may be waited for here (and the graph may be replayed on any stream later).
    if (cap != cudaStreamCaptureStatusNone) return B2S_OK;
    if (idx->last_stream_valid && idx->last_stream != s) {
        cudaError_t e = cudaSuccess;
        if (!idx->xs_event) e = cudaEventCreateWithFlags(&idx->xs_event, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(idx->xs_event, idx->last_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(s, idx->xs_event, 0);
        if (e != cudaSuccess) {   // e.g. the previous stream was destroyed by its owner
            cudaGetLastError();
            CUDA_TRY(cudaDeviceSynchronize());
        }
    }
    idx->last_stream = s;
    idx->last_stream_valid = true;
    return B2S_OK;
}

int ensure_pinned(void** p, size_t* have, size_t need) {
    if (need <= *have) return B2S_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *have = 0;
    size_t want = std::max<size_t>(need, 1 << 16);
    cudaError_t e = cudaMallocHost(p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        return fail(B2S_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    *have = want;
    return B2S_OK;
}

}  // namespace

#ifndef B2S_NO_TENSOR_PATH
#include "gemm_topk_host.inl"
#endif

// =============================================================================================
// extern "C"
// =============================================================================================

static void exchange_release(b2s_index* idx) {
    auto& ex = idx->ex;
    for (void* m : ex.opened) cudaIpcCloseMemHandle(m);
    ex.opened.clear();
    if (ex.peers_dev) cudaFree(ex.peers_dev);
    if (ex.status) cudaFreeHost(ex.status);
    if (ex.local) cudaFree(ex.local);
    const long long keep_timeout = ex.timeout_ms;
    ex = b2s_index::Exchange();
    ex.timeout_ms = keep_timeout;
    cudaGetLastError();
}


extern "C" {

B2S_API int b2s_version(void) { return 200; }

B2S_API const char* b2s_last_error(void) { return g_err.c_str(); }

B2S_API int b2s_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

B2S_API int b2s_create(int dim, int metric, int device, b2s_index** out) {
    if (!out) return fail(B2S_ERR_INVALID, "out is null");
    *out = nullptr;
    if (!dim_supported(dim))
        return fail(B2S_ERR_UNSUPPORTED, "dim must be one of 128, 256, 384, 512, 768, 1024");
    if (metric != B2S_METRIC_INNER_PRODUCT && metric != B2S_METRIC_COSINE)
        return fail(B2S_ERR_INVALID, "metric must be B2S_METRIC_INNER_PRODUCT or B2S_METRIC_COSINE");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2S_ERR_NO_DEVICE, "no CUDA device: libb200search has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(B2S_ERR_INVALID, "device ordinal out of range");
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(B2S_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                           ", this library is built for sm_100a only");
    CUDA_TRY(cudaSetDevice(device));
    b2s_index* idx = new (std::nothrow) b2s_index();
    if (!idx) return fail(B2S_ERR_NOMEM, "host allocation failed");
    idx->dim = dim;
    idx->metric = metric;
    idx->device = device;
    idx->num_sms = prop.multiProcessorCount;
    memset(&idx->stats, 0, sizeof(idx->stats));
    if (cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete idx;
        return fail(B2S_ERR_CUDA, "cudaStreamCreate failed");
    }
    // two control sets of the scan kernel (done ticket, dynamic-tail counters, cascade slots, epoch): all zero
    if (cudaMalloc((void**)&idx->ctl, 2 * sizeof(ScanCtl)) != cudaSuccess ||
        cudaMemset(idx->ctl, 0, 2 * sizeof(ScanCtl)) != cudaSuccess) {
        cudaGetLastError();
        idx->ctl = nullptr;   // the fused tail, the dynamic tail and the cascade select are simply not used
    }
    *out = idx;
    return B2S_OK;
}

B2S_API int b2s_destroy(b2s_index* idx) {
    if (!idx) return B2S_OK;
    cudaSetDevice(idx->device);
    cudaDeviceSynchronize();
    if (idx->rows) cudaFree(idx->rows);
    if (idx->rows_f32) cudaFree(idx->rows_f32);
    idx->ws_lists.release();
    idx->ws_counts.release();
    idx->ws_seed.release();
    idx->ws_thr.release();
    idx->ws_gmax.release();
    idx->ws_hist.release();
    idx->ws_rs_scores.release();
    idx->ws_rs_ids.release();
    idx->ws_hcfg.release();
    idx->ws_qf32.release();
    idx->ws_qbf16.release();
    idx->ws_io_q.release();
    idx->ws_io_ids.release();
    idx->ws_tmp.release();
#ifndef B2S_NO_TENSOR_PATH
    tensor_path_release(idx);
#endif
    exchange_release(idx);
    if (idx->ctl) cudaFree(idx->ctl);
    if (idx->trace_buf) cudaFree(idx->trace_buf);
    if (idx->xs_event) cudaEventDestroy(idx->xs_event);
    if (idx->pin_q) cudaFreeHost(idx->pin_q);
    for (int i = 0; i < 2; ++i) {
        if (idx->pin_in[i]) cudaFreeHost(idx->pin_in[i]);
        if (idx->pin_in_ev[i]) cudaEventDestroy(idx->pin_in_ev[i]);
    }
    if (idx->pin_out) cudaFreeHost(idx->pin_out);
    if (idx->ring) {
        for (int i = 0; i < 4 * kTimingSlots; ++i)
            if (idx->ring[i]) cudaEventDestroy(idx->ring[i]);
        delete[] idx->ring;
    }
    if (idx->stream) cudaStreamDestroy(idx->stream);
    cudaGetLastError();
    delete idx;
    return B2S_OK;
}

B2S_API int b2s_reserve(b2s_index* idx, int64_t n_rows) {
    if (!idx || n_rows < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    return grow_rows(idx, n_rows);
}

// Host rows reach the device through two pinned staging buffers: while the copy engine moves chunk c and the
// convert kernel runs, a few host threads already copy chunk c+1 into the other buffer (a pageable
// cudaMemcpyAsync would serialise the staging copy, the DMA and the kernel, chunk after chunk).
constexpr size_t kIngestChunkBytes = (size_t)64 << 20;
static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const int nt = (int)std::min<size_t>(std::min(8u, hw), bytes >> 22);   // >= 4 MB per thread
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (int t = 1; t < nt; ++t) {
        const size_t off = (size_t)t * per;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((unsigned char*)dst + off, (const unsigned char*)src + off, len); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto& t : th) t.join();
}

static int add_impl(b2s_index* idx, const void* rows, int64_t n, int is_device, bool is_bf16, bool prepared = false) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (n < 0) return fail(B2S_ERR_INVALID, "n < 0");
    if (n == 0) return B2S_OK;
    if (!rows) return fail(B2S_ERR_INVALID, "rows is null");
    if (idx->n + n > 0x7fffffff) return fail(B2S_ERR_UNSUPPORTED, "more than 2^31-1 rows in one shard");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    if ((rc = grow_rows(idx, idx->n + n)) != B2S_OK) return rc;
    const int dim = idx->dim;
    const size_t esz = is_bf16 ? 2 : 4;
    const bool normalize = idx->metric == B2S_METRIC_COSINE && !prepared;
    const size_t chunk_bytes = is_device ? ((size_t)256 << 20) : kIngestChunkBytes;
    const int64_t chunk_rows = std::max<int64_t>(1, (int64_t)chunk_bytes / ((int64_t)dim * (int64_t)esz));
    const size_t stage_bytes = (size_t)chunk_rows * dim * esz;
    if (!is_device) {
        if ((rc = idx->ws_tmp.ensure(2 * stage_bytes)) != B2S_OK) return rc;   // two device halves as well
        for (int i = 0; i < 2; ++i) {
            if (!idx->pin_in[i]) {
                // (stage_bytes <= kIngestChunkBytes for either element size: chunk_rows is a floor)
                if (cudaHostAlloc(&idx->pin_in[i], std::max(kIngestChunkBytes, (size_t)dim * 4), cudaHostAllocDefault) != cudaSuccess) {
                    cudaGetLastError();
                    return fail(B2S_ERR_NOMEM, "cudaHostAlloc failed (host ingest staging)");
                }
            }
            if (!idx->pin_in_ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&idx->pin_in_ev[i], cudaEventDisableTiming));
        }
    }
    int64_t chunk_no = 0;
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows, ++chunk_no) {
        const int64_t rn = std::min<int64_t>(chunk_rows, n - r0);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(rows) + (size_t)r0 * dim * esz;
        const void* dsrc = src;
        if (!is_device) {
            const int b = (int)(chunk_no & 1);
            const size_t bytes = (size_t)rn * dim * esz;
            if (chunk_no >= 2) CUDA_TRY(cudaEventSynchronize(idx->pin_in_ev[b]));   // its previous DMA has finished
            parallel_memcpy(idx->pin_in[b], src, bytes);
            unsigned char* dtmp = reinterpret_cast<unsigned char*>(idx->ws_tmp.p) + (size_t)b * stage_bytes;
            CUDA_TRY(cudaMemcpyAsync(dtmp, idx->pin_in[b], bytes, cudaMemcpyHostToDevice, idx->stream));
            CUDA_TRY(cudaEventRecord(idx->pin_in_ev[b], idx->stream));
            dsrc = dtmp;
        }
        __nv_bfloat16* dst = idx->rows + (size_t)(idx->n + r0) * dim;
        const int warps = 8;
        const unsigned blocks = (unsigned)((rn + warps - 1) / warps);
        if (is_bf16) {
            if (normalize) {
                rows_bf16_normalize_kernel<<<blocks, warps * 32, 0, idx->stream>>>(
                    reinterpret_cast<const __nv_bfloat16*>(dsrc), dst, rn, dim);
                CUDA_TRY(cudaGetLastError());
            } else {
                CUDA_TRY(cudaMemcpyAsync(dst, dsrc, (size_t)rn * dim * 2, cudaMemcpyDeviceToDevice, idx->stream));
            }
            if (idx->opt_keep_f32 && idx->rows_f32) {
                rows_bf16_to_f32_kernel<<<1024, 256, 0, idx->stream>>>(dst, idx->rows_f32 + (size_t)(idx->n + r0) * dim,
                                                                      (long long)rn * dim);
                CUDA_TRY(cudaGetLastError());
            }
        } else {
            rows_f32_to_bf16_kernel<<<blocks, warps * 32, 0, idx->stream>>>(
                reinterpret_cast<const float*>(dsrc), dst, rn, dim, normalize ? 1 : 0);
            CUDA_TRY(cudaGetLastError());
            if (idx->opt_keep_f32 && idx->rows_f32) {
                // fp32 copy keeps the caller's values (unit-normalised for the cosine metric)
                CUDA_TRY(cudaMemcpyAsync(idx->rows_f32 + (size_t)(idx->n + r0) * dim, dsrc,
                                         (size_t)rn * dim * sizeof(float), cudaMemcpyDeviceToDevice, idx->stream));
                if (normalize) {
                    rows_f32_normalize_kernel<<<blocks, warps * 32, 0, idx->stream>>>(
                        idx->rows_f32 + (size_t)(idx->n + r0) * dim, rn, dim);
                    CUDA_TRY(cudaGetLastError());
                }
            }
        }
        if (is_device) CUDA_TRY(cudaStreamSynchronize(idx->stream));   // the caller's buffer may be reused right away
    }
    CUDA_TRY(cudaStreamSynchronize(idx->stream));
    idx->n += n;
#ifndef B2S_NO_TENSOR_PATH
    idx->tc.corpus_map_valid = false;
#endif
    return B2S_OK;
}

B2S_API int b2s_add_f32(b2s_index* idx, const float* rows, int64_t n, int is_device) {
    return add_impl(idx, rows, n, is_device, false);
}
B2S_API int b2s_add_bf16(b2s_index* idx, const void* rows, int64_t n, int is_device) {
    return add_impl(idx, rows, n, is_device, true);
}

B2S_API int b2s_add_prepared(b2s_index* idx, const void* rows, int q_dtype, int64_t n, int is_device) {
    if (q_dtype != B2S_DTYPE_F32 && q_dtype != B2S_DTYPE_BF16) return fail(B2S_ERR_INVALID, "bad dtype");
    return add_impl(idx, rows, n, is_device, q_dtype == B2S_DTYPE_BF16, true);
}

B2S_API int64_t b2s_ntotal(const b2s_index* idx) { return idx ? idx->n : 0; }
B2S_API int b2s_dim(const b2s_index* idx) { return idx ? idx->dim : 0; }

B2S_API int b2s_reset(b2s_index* idx) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    std::lock_guard<std::mutex> g(idx->mu);
    idx->n = 0;
#ifndef B2S_NO_TENSOR_PATH
    idx->tc.corpus_map_valid = false;
#endif
    return B2S_OK;
}

B2S_API int b2s_set_id_offset(b2s_index* idx, int64_t offset) {
    if (!idx || offset < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    idx->id_offset = offset;
    return B2S_OK;
}

B2S_API int b2s_set_option(b2s_index* idx, const char* name, int64_t value) {
    if (!idx || !name) return fail(B2S_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(idx->mu);
    const std::string s(name);
    if (s == "path") {
        if (value < 0 || value > 2) return fail(B2S_ERR_INVALID, "path must be 0, 1 or 2");
        idx->opt_path = (int)value;
    } else if (s == "seed") {
        idx->opt_seed = (int)value;
    } else if (s == "scan_ctas_per_sm") {
        if (value < 1 || value > 8) return fail(B2S_ERR_INVALID, "scan_ctas_per_sm must be in [1, 8]");
        idx->opt_scan_ctas_per_sm = (int)value;
    } else if (s == "keep_f32") {
        if (idx->n > 0 && value && !idx->opt_keep_f32)
            return fail(B2S_ERR_INVALID, "keep_f32 must be set before the first add");
        idx->opt_keep_f32 = value ? 1 : 0;
    } else if (s == "rescore_pad") {
        idx->opt_rescore_pad = (int)std::max<int64_t>(0, value);
    } else if (s == "timing") {
        if (value && !idx->ring) {
            int rc = use_device(idx);
            if (rc != B2S_OK) return rc;
            idx->ring = new (std::nothrow) cudaEvent_t[4 * kTimingSlots]();
            if (!idx->ring) return fail(B2S_ERR_NOMEM, "host allocation failed");
            for (int i = 0; i < 4 * kTimingSlots; ++i) CUDA_TRY(cudaEventCreate(&idx->ring[i]));
        }
        idx->opt_timing = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(0, value));
        idx->timing_seq = 0;
        idx->ring_calls = 0;
        idx->ev_valid = false;
    } else if (s == "tc_min_nq") {
        idx->opt_tc_min_nq = (int)std::max<int64_t>(1, value);
    } else if (s == "exchange_ll") {
        idx->opt_exchange_ll = value ? 1 : 0;
    } else if (s == "host_inline") {
        idx->opt_host_inline = value ? 1 : 0;
    } else if (s == "fused_tail") {
        idx->opt_fused_tail = value ? 1 : 0;
    } else if (s == "pdl") {
        if (value < 0 || value > 2) return fail(B2S_ERR_INVALID, "pdl must be 0, 1 or 2");
        idx->opt_pdl = (int)value;
    } else if (s == "exchange_timeout_ms") {
        if (value < 1 || value > 600000) return fail(B2S_ERR_INVALID, "exchange_timeout_ms must be in [1, 600000]");
        idx->ex.timeout_ms = value;
    } else if (s == "prefetch_iters") {
        if (value < 0 || value > 32) return fail(B2S_ERR_INVALID, "prefetch_iters must be in [0, 32]");
        idx->opt_prefetch_iters = (int)value;
    } else if (s == "dynamic_tail") {
        if (value < 0 || value > 64) return fail(B2S_ERR_INVALID, "dynamic_tail must be in [0, 64]");
        idx->opt_dynamic_tail = (int)value;
    } else if (s == "cascade") {
        idx->opt_cascade = value ? 1 : 0;
    } else if (s == "phase_a") {
        if (value < 0 || value > 4096) return fail(B2S_ERR_INVALID, "phase_a must be in [0, 4096]");
        idx->opt_phase_a = (int)value;
    } else if (s == "transition_mode") {
        if (value < 0 || value > 2) return fail(B2S_ERR_INVALID, "transition_mode must be 0, 1 or 2");
        idx->opt_transition_mode = (int)value;
    } else if (s == "peek_every") {
        if (value < 0 || value > 4096) return fail(B2S_ERR_INVALID, "peek_every must be in [0, 4096]");
        idx->opt_peek_every = (int)value;
    } else if (s == "phase_a_stagger") {
        if (value < 1 || value > 4096) return fail(B2S_ERR_INVALID, "phase_a_stagger must be in [1, 4096]");
        idx->opt_phase_a_stagger = (int)value;
    } else if (s == "host_spin") {
        idx->opt_host_spin = value ? 1 : 0;
    } else if (s == "grid_spare") {
        if (value < 0 || value > 64) return fail(B2S_ERR_INVALID, "grid_spare must be in [0, 64]");
        idx->opt_grid_spare = (int)value;
    } else if (s == "cascade_min_units") {
        if (value < 4 || value > 1 << 20) return fail(B2S_ERR_INVALID, "cascade_min_units must be in [4, 2^20]");
        idx->opt_cascade_min_units = (int)value;
    } else if (s == "pdl_early") {
        idx->opt_pdl_early = value ? 1 : 0;
    } else if (s == "trace") {
        if (value && !idx->trace_buf) {
            int rc = use_device(idx);
            if (rc != B2S_OK) return rc;
            const size_t bytes = (size_t)2 * (kTraceWords + kTraceArrays * kTraceStride) * sizeof(unsigned long long);
            CUDA_TRY(cudaMalloc((void**)&idx->trace_buf, bytes));
            CUDA_TRY(cudaMemset(idx->trace_buf, 0, bytes));
        }
        idx->opt_trace = value ? 1 : 0;
    } else if (s == "tc_sample_div") {
        idx->opt_tc_sample_div = (int)std::min<int64_t>(1 << 20, std::max<int64_t>(0, value));
    } else if (s == "tc_shared_thr") {
        idx->opt_tc_shared_thr = value ? 1 : 0;
    } else if (s == "tc_thr_period_ns") {
        idx->opt_tc_thr_period_ns = (int)std::min<int64_t>(1000000, std::max<int64_t>(100, value));
    } else if (s == "tc_single_cta") {
        idx->opt_tc_single_cta = value ? 1 : 0;
    } else if (s == "tc_chunk_tiles") {
        if (value < 1 || value > 4096) return fail(B2S_ERR_INVALID, "tc_chunk_tiles must be in [1, 4096]");
        idx->opt_tc_chunk_lo = idx->opt_tc_chunk_hi = (int)value;
    } else {
        return fail(B2S_ERR_INVALID, "unknown option: " + s);
    }
    return B2S_OK;
}

B2S_API int64_t b2s_get_option(const b2s_index* idx, const char* name) {
    if (!idx || !name) return -1;
    const std::string s(name);
    if (s == "path") return idx->opt_path;
    if (s == "seed") return idx->opt_seed;
    if (s == "scan_ctas_per_sm") return idx->opt_scan_ctas_per_sm;
    if (s == "keep_f32") return idx->opt_keep_f32;
    if (s == "rescore_pad") return idx->opt_rescore_pad;
    if (s == "timing") return idx->opt_timing;
    if (s == "tc_min_nq") return idx->opt_tc_min_nq;
    if (s == "pdl") return idx->opt_pdl;
    if (s == "prefetch_iters") return idx->opt_prefetch_iters;
    if (s == "dynamic_tail") return idx->opt_dynamic_tail;
    if (s == "cascade") return idx->opt_cascade;
    if (s == "phase_a") return idx->opt_phase_a;
    if (s == "phase_a_stagger") return idx->opt_phase_a_stagger;
    if (s == "trace") return idx->opt_trace;
    if (s == "pdl_early") return idx->opt_pdl_early;
    if (s == "grid_spare") return idx->opt_grid_spare;
    if (s == "host_spin") return idx->opt_host_spin;
    if (s == "cascade_min_units") return idx->opt_cascade_min_units;
    if (s == "transition_mode") return idx->opt_transition_mode;
    if (s == "peek_every") return idx->opt_peek_every;
    if (s == "exchange_timeout_ms") return idx->ex.timeout_ms;
    if (s == "fused_tail") return idx->opt_fused_tail;
    if (s == "tc_sample_div") return idx->opt_tc_sample_div;
    if (s == "tc_shared_thr") return idx->opt_tc_shared_thr;
    if (s == "tc_chunk_tiles") return idx->opt_tc_chunk_lo == idx->opt_tc_chunk_hi ? idx->opt_tc_chunk_lo : 0;
    if (s == "num_sms") return idx->num_sms;
    return -1;
}

B2S_API int b2s_search_device(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k,
                              float* out_scores, int64_t* out_ids, void* cuda_stream, unsigned flags) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (flags & ~(unsigned)B2S_SEARCH_STABLE_QUERIES) return fail(B2S_ERR_INVALID, "unknown search flag");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    if ((rc = guard_stream(idx, s)) != B2S_OK) return rc;
    idx->call_flags = flags;
    rc = search_impl(idx, queries, q_dtype, nq, k, out_scores, out_ids, s);
    idx->call_flags = 0;
    return rc;
}

// Host-buffer search shared by b2s_search and b2s_search_sharded.  Three regimes:
//  * tiny (1-2 queries on the scan path, <= 4 KB of results): the query rides in the kernel parameters and the
//    kernel writes the answer straight into mapped pinned host memory -- no copy-engine operation at all;
//  * small (<= 1 MB each way): staged through pinned memory, one H2D and one D2H copy;
//  * large: copies straight from / to the caller's (pageable) buffers.
static int search_sharded_impl(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                               int64_t* out_ids, void* cuda_stream, int phase);

static int search_host(b2s_index* idx, const float* queries, int64_t nq, int k, float* out_scores, int64_t* out_ids,
                       bool sharded) {
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    if ((rc = guard_stream(idx, idx->stream)) != B2S_OK) return rc;
    const size_t qbytes = (size_t)nq * idx->dim * sizeof(float);
    const size_t sbytes = (size_t)nq * k * sizeof(float);
    const size_t ibytes = (size_t)nq * k * sizeof(int64_t);
    if ((rc = idx->ws_io_q.ensure(qbytes)) != B2S_OK) return rc;
    if ((rc = idx->ws_io_ids.ensure(ibytes + sbytes)) != B2S_OK) return rc;   // [ids | scores]: one D2H copy
    int64_t* io_ids = reinterpret_cast<int64_t*>(idx->ws_io_ids.p);
    float* io_scores = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(idx->ws_io_ids.p) + ibytes);
    const bool small = qbytes <= ((size_t)1 << 20) && (sbytes + ibytes) <= ((size_t)1 << 20);
    // (not with fp32 re-ranking: it searches k + pad candidates, possibly on another path, and re-reads the query)
    const bool tiny = idx->opt_host_inline && idx->n > 0 && !(idx->opt_keep_f32 && idx->rows_f32) &&
                      choose_path(idx, nq, k) == B2S_PATH_SCAN &&
                      nq <= scan_max_nq(idx->dim) && nq * idx->dim <= kScanInlineFloats && (sbytes + ibytes) <= 4096;
    unsigned char* po = nullptr;
    if (small) {
        // (+ one 64-byte line behind the outputs: the completion flag of the tiny path)
        if ((rc = ensure_pinned(&idx->pin_out, &idx->pin_out_bytes, ((sbytes + ibytes + 63) & ~(size_t)63) + 64)) != B2S_OK) return rc;
        po = reinterpret_cast<unsigned char*>(idx->pin_out);
    }
    volatile unsigned* flag = nullptr;
    if (tiny) {
        flag = reinterpret_cast<volatile unsigned*>(po + ((sbytes + ibytes + 63) & ~(size_t)63));
        idx->host_flag = const_cast<unsigned*>(flag);
        idx->host_seq = ++idx->host_calls;
        if (idx->host_seq == 0u) idx->host_seq = ++idx->host_calls;   // (wrap-around: 0 means "not yet")
        *flag = 0u;   // the word moves with (nq, k) and pinned memory is recycled: never trust what it holds
        idx->host_flag_armed = false;
        idx->host_q = queries;                                     // read at launch time, inside this call
        io_ids = reinterpret_cast<int64_t*>(po);                   // pinned host memory is device-accessible (UVA)
        io_scores = reinterpret_cast<float*>(po + ibytes);
    } else if (small) {
        // stage through pinned memory so that both copies are truly asynchronous DMA
        if ((rc = ensure_pinned(&idx->pin_q, &idx->pin_q_bytes, qbytes)) != B2S_OK) return rc;
        memcpy(idx->pin_q, queries, qbytes);
        CUDA_TRY(cudaMemcpyAsync(idx->ws_io_q.p, idx->pin_q, qbytes, cudaMemcpyHostToDevice, idx->stream));
    } else {
        CUDA_TRY(cudaMemcpyAsync(idx->ws_io_q.p, queries, qbytes, cudaMemcpyHostToDevice, idx->stream));
    }
    rc = sharded ? search_sharded_impl(idx, idx->ws_io_q.p, B2S_DTYPE_F32, nq, k, io_scores, io_ids, idx->stream, 0)
                 : search_impl(idx, idx->ws_io_q.p, B2S_DTYPE_F32, nq, k, io_scores, io_ids, idx->stream);
    idx->host_q = nullptr;
    idx->host_flag = nullptr;
    if (rc != B2S_OK) return rc;
    if (small) {
        if (!tiny) CUDA_TRY(cudaMemcpyAsync(po, idx->ws_io_ids.p, ibytes + sbytes, cudaMemcpyDeviceToHost, idx->stream));
        bool seen = false;
        if (tiny && idx->host_flag_armed && idx->opt_host_spin) {
            // spin on the kernel's flag; every few hundred polls ask the stream, so that a failed or (impossibly)
            // flag-less launch ends the wait through the ordinary synchronise below
            const unsigned want = idx->host_seq;
            for (unsigned spins = 1; !(seen = (*flag == want)); ++spins) {
                if ((spins & 511u) == 0u && cudaStreamQuery(idx->stream) != cudaErrorNotReady) break;
                B2S_CPU_RELAX();
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            seen = seen || *flag == want;
        }
        if (!seen) CUDA_TRY(cudaStreamSynchronize(idx->stream));
        memcpy(out_ids, po, ibytes);
        memcpy(out_scores, po + ibytes, sbytes);
    } else {
        CUDA_TRY(cudaMemcpyAsync(out_ids, idx->ws_io_ids.p, ibytes, cudaMemcpyDeviceToHost, idx->stream));
        CUDA_TRY(cudaMemcpyAsync(out_scores, io_scores, sbytes, cudaMemcpyDeviceToHost, idx->stream));
        CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    if (sharded && idx->ex.status && *(volatile unsigned*)idx->ex.status != 0u)
        return fail(B2S_ERR_CUDA, "sharded search: timed out waiting for a peer rank's candidates (call #" +
                                      std::to_string(*(volatile unsigned*)idx->ex.status) + "); the results are invalid");
    return B2S_OK;
}

B2S_API int b2s_search(b2s_index* idx, const float* queries, int64_t nq, int k, float* out_scores,
                       int64_t* out_ids) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "nq and k must be >= 0");
    if (nq == 0 || k == 0) return B2S_OK;
    if (!queries || !out_scores || !out_ids) return fail(B2S_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    return search_host(idx, queries, nq, k, out_scores, out_ids, false);
}

B2S_API int b2s_merge_device(int device, const float* scores, const int64_t* ids, int g, int64_t nq,
                             int k, float* out_scores, int64_t* out_ids, void* cuda_stream) {
    if (g < 1 || nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (nq == 0 || k == 0) return B2S_OK;
    if (!scores || !ids || !out_scores || !out_ids) return fail(B2S_ERR_INVALID, "null buffer");
    if ((int64_t)g * k > (int64_t)1 << 24) return fail(B2S_ERR_UNSUPPORTED, "g * k too large");
    CUDA_TRY(cudaSetDevice(device));
    MergeParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.k = k;
    mp.in_scores = scores;
    mp.in_ids = reinterpret_cast<const long long*>(ids);
    mp.g = g;
    mp.k_in = k;
    mp.nq = nq;
    mp.g_stride_ids = (long long)nq * k;
    mp.g_stride_scores = (long long)nq * k;
    mp.out_scores = out_scores;
    mp.out_ids = reinterpret_cast<long long*>(out_ids);
    merge_pairs_kernel<<<(unsigned)nq, kMergeThreads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(mp);
    CUDA_TRY(cudaGetLastError());
    return B2S_OK;
}

// compute_similarity has no index handle: its device buffers, pinned staging and stream are cached per device
// (ANCEMiner.mine calls it twice per query with a 1 x <= 20 problem: no allocation in steady state).
namespace {
struct SimWorkspace {
    DevBuf q, d, out;
    void* pin = nullptr;
    size_t pin_bytes = 0;
    cudaStream_t stream = nullptr;
    bool attr_set = false;
};
std::mutex g_sim_mu;
SimWorkspace g_sim[64];
}  // namespace

B2S_API int b2s_similarity(int device, const float* q, int64_t nq, const float* d, int64_t nd, int dim,
                           float* out) {
    if (nq < 0 || nd < 0 || dim <= 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (nq == 0 || nd == 0) return B2S_OK;
    if (!q || !d || !out) return fail(B2S_ERR_INVALID, "null buffer");
    if ((size_t)dim * 8 * sizeof(float) > 96 * 1024) return fail(B2S_ERR_UNSUPPORTED, "dim too large");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2S_ERR_NO_DEVICE, "no CUDA device: libb200search has no CPU fallback");
    }
    if (device < 0 || device >= ndev || device >= 64) return fail(B2S_ERR_INVALID, "device ordinal out of range");
    CUDA_TRY(cudaSetDevice(device));
    std::lock_guard<std::mutex> g(g_sim_mu);
    SimWorkspace& w = g_sim[device];
    if (!w.stream) CUDA_TRY(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    const size_t qb = (size_t)nq * dim * 4, db = (size_t)nd * dim * 4, ob = (size_t)nq * nd * 4;
    int rc;
    if ((rc = w.q.ensure(qb)) != B2S_OK || (rc = w.d.ensure(db)) != B2S_OK || (rc = w.out.ensure(ob)) != B2S_OK) return rc;
    const size_t smem = (size_t)dim * 8 * sizeof(float);
    if (smem > 48 * 1024 && !w.attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(similarity_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        w.attr_set = true;
    }
    // small problems go through pinned staging (true DMA, one synchronise); large ones copy from the caller's buffers
    const bool small = qb + db + ob <= ((size_t)1 << 20);
    unsigned char* pin = nullptr;
    if (small) {
        if ((rc = ensure_pinned(&w.pin, &w.pin_bytes, qb + db + ob)) != B2S_OK) return rc;
        pin = reinterpret_cast<unsigned char*>(w.pin);
        memcpy(pin, q, qb);
        memcpy(pin + qb, d, db);
    }
    CUDA_TRY(cudaMemcpyAsync(w.q.p, small ? (const void*)pin : (const void*)q, qb, cudaMemcpyHostToDevice, w.stream));
    CUDA_TRY(cudaMemcpyAsync(w.d.p, small ? (const void*)(pin + qb) : (const void*)d, db, cudaMemcpyHostToDevice, w.stream));
    similarity_f32_kernel<<<(unsigned)((nd + 7) / 8), 256, smem, w.stream>>>(
        reinterpret_cast<const float*>(w.q.p), nq, reinterpret_cast<const float*>(w.d.p), nd, dim,
        reinterpret_cast<float*>(w.out.p));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(small ? (void*)(pin + qb + db) : (void*)out, w.out.p, ob, cudaMemcpyDeviceToHost, w.stream));
    CUDA_TRY(cudaStreamSynchronize(w.stream));
    if (small) memcpy(out, pin + qb + db, ob);
    return B2S_OK;
}

B2S_API int b2s_read_rows_f32(b2s_index* idx, int64_t start, int64_t n, float* out_host) {
    if (!idx || start < 0 || n < 0 || start + n > idx->n) return fail(B2S_ERR_INVALID, "row range out of bounds");
    if (n == 0) return B2S_OK;
    if (!out_host) return fail(B2S_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    const int64_t chunk_rows = std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)idx->dim * 4));
    for (int64_t r0 = 0; r0 < n; r0 += chunk_rows) {
        const int64_t rn = std::min<int64_t>(chunk_rows, n - r0);
        if ((rc = idx->ws_tmp.ensure((size_t)rn * idx->dim * 4)) != B2S_OK) return rc;
        rows_bf16_to_f32_kernel<<<1024, 256, 0, idx->stream>>>(idx->rows + (size_t)(start + r0) * idx->dim,
                                                              reinterpret_cast<float*>(idx->ws_tmp.p),
                                                              (long long)rn * idx->dim);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_host + (size_t)r0 * idx->dim, idx->ws_tmp.p, (size_t)rn * idx->dim * 4,
                                 cudaMemcpyDeviceToHost, idx->stream));
        CUDA_TRY(cudaStreamSynchronize(idx->stream));
    }
    return B2S_OK;
}

B2S_API int b2s_read_timings(const b2s_index* idx, float* dominant_ms, float* total_ms, int max_n) {
    if (!idx || max_n < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (!idx->ring) return 0;
    const int64_t have = std::min<int64_t>(idx->ring_calls, kTimingSlots);
    const int n = (int)std::min<int64_t>(have, max_n);
    for (int i = 0; i < n; ++i) {
        const int64_t call = idx->ring_calls - n + i;
        cudaEvent_t* e = idx->ring + 4 * (call % kTimingSlots);
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, e[1], e[2]) != cudaSuccess) a = -1.f;
        if (cudaEventElapsedTime(&b, e[0], e[3]) != cudaSuccess) b = -1.f;
        cudaGetLastError();
        if (dominant_ms) dominant_ms[i] = a;
        if (total_ms) total_ms[i] = b;
    }
    return n;
}

B2S_API int b2s_read_trace(b2s_index* idx, uint64_t* out, int max_words) {
    if (!idx || !out || max_words < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (!idx->trace_buf) return 0;
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    constexpr int kHalf = kTraceWords + kTraceArrays * kTraceStride;
    const int n = std::min(max_words, 2 * kHalf);
    CUDA_TRY(cudaDeviceSynchronize());
    // the half of the LAST traced launch first, then the launch before it
    const unsigned last = (idx->trace_seq + 1u) & 1u;
    const int n0 = std::min(n, kHalf);
    CUDA_TRY(cudaMemcpy(out, idx->trace_buf + (size_t)last * kHalf, (size_t)n0 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (n > kHalf)
        CUDA_TRY(cudaMemcpy(out + kHalf, idx->trace_buf + (size_t)(last ^ 1u) * kHalf, (size_t)(n - kHalf) * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost));
    return n;
}

B2S_API int b2s_merge_packed_device(int device, const void* packed, int g, int64_t nq, int k, float* out_scores,
                                    int64_t* out_ids, void* cuda_stream) {
    if (!packed) return fail(B2S_ERR_INVALID, "null buffer");
    // per rank: [ids int64 nq*k][scores fp32 nq*k], padded to a multiple of 16 bytes
    const size_t per_rank = b2s_packed_bytes(nq, k);
    if (g < 1 || nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (nq == 0 || k == 0) return B2S_OK;
    if (!out_scores || !out_ids) return fail(B2S_ERR_INVALID, "null buffer");
    if ((int64_t)g * k > (int64_t)1 << 24) return fail(B2S_ERR_UNSUPPORTED, "g * k too large");
    CUDA_TRY(cudaSetDevice(device));
    MergeParams mp;
    memset(&mp, 0, sizeof(mp));
    mp.k = k;
    mp.in_ids = reinterpret_cast<const long long*>(packed);
    mp.in_scores = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(packed) + (size_t)nq * k * 8);
    mp.g = g;
    mp.k_in = k;
    mp.nq = nq;
    mp.g_stride_ids = (long long)(per_rank / 8);
    mp.g_stride_scores = (long long)(per_rank / 4);
    mp.out_scores = out_scores;
    mp.out_ids = reinterpret_cast<long long*>(out_ids);
    merge_pairs_kernel<<<(unsigned)nq, kMergeThreads, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(mp);
    CUDA_TRY(cudaGetLastError());
    return B2S_OK;
}

B2S_API int64_t b2s_packed_bytes(int64_t nq, int k) {
    const int64_t raw = nq * (int64_t)k * 12;
    return (raw + 15) / 16 * 16;
}

B2S_API const void* b2s_rows_device(const b2s_index* idx) { return idx ? idx->rows : nullptr; }

B2S_API int b2s_score_rows_device(b2s_index* idx, const void* queries, int q_dtype, int round_q_bf16, int64_t nq,
                                  const int64_t* ids, int m, float* out_scores, void* cuda_stream) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (nq < 0 || m < 0) return fail(B2S_ERR_INVALID, "nq and m must be >= 0");
    if (nq == 0 || m == 0) return B2S_OK;
    if (!queries || !ids || !out_scores) return fail(B2S_ERR_INVALID, "null buffer");
    if (q_dtype != B2S_DTYPE_F32 && q_dtype != B2S_DTYPE_BF16) return fail(B2S_ERR_INVALID, "bad q_dtype");
    if (idx->metric == B2S_METRIC_COSINE)
        return fail(B2S_ERR_UNSUPPORTED, "b2s_score_rows_device expects pre-normalised queries: use an inner-product index");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    const long long pairs = (long long)nq * m;
    const int warps = 8;
    score_rows_kernel<<<(unsigned)((pairs + warps - 1) / warps), warps * 32, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        idx->rows, idx->n, idx->dim, queries, q_dtype == B2S_DTYPE_BF16, round_q_bf16 ? 1 : 0, nq, m,
        reinterpret_cast<const long long*>(ids), idx->id_offset, out_scores);
    CUDA_TRY(cudaGetLastError());
    return B2S_OK;
}

B2S_API int b2s_ance_filter_device(int device, const float* cand_scores, const int64_t* cand_ids, int64_t nq, int k_in,
                                   const int64_t* pos_ids, const float* pos_scores, int n_pos, float margin, int top_k,
                                   int64_t* out_ids, float* out_scores, int32_t* out_counts, void* cuda_stream) {
    if (nq < 0 || k_in < 0 || n_pos < 0 || top_k < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (nq == 0 || top_k == 0) return B2S_OK;
    if (!out_ids || !out_scores || (k_in > 0 && (!cand_scores || !cand_ids)) || (n_pos > 0 && (!pos_ids || !pos_scores)))
        return fail(B2S_ERR_INVALID, "null buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2S_ERR_NO_DEVICE, "no CUDA device: libb200search has no CPU fallback");
    }
    CUDA_TRY(cudaSetDevice(device));
    ance_filter_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        cand_scores, reinterpret_cast<const long long*>(cand_ids), k_in, reinterpret_cast<const long long*>(pos_ids),
        pos_scores, n_pos, margin, top_k, nq, reinterpret_cast<long long*>(out_ids), out_scores, out_counts);
    CUDA_TRY(cudaGetLastError());
    return B2S_OK;
}

B2S_API int b2s_maxsim_device(int device, const float* scores, const int64_t* ids, int64_t nq, int k_in,
                              const int64_t* chunk_to_doc, int64_t n_chunks, int k_out, float* out_scores,
                              int64_t* out_doc_ids, int32_t* out_counts, void* cuda_stream) {
    if (nq < 0 || k_in < 0 || k_out < 0 || n_chunks < 0) return fail(B2S_ERR_INVALID, "bad arguments");
    if (nq == 0 || k_out == 0) return B2S_OK;
    if (!out_scores || !out_doc_ids || (k_in > 0 && (!scores || !ids)) || (n_chunks > 0 && !chunk_to_doc))
        return fail(B2S_ERR_INVALID, "null buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2S_ERR_NO_DEVICE, "no CUDA device: libb200search has no CPU fallback");
    }
    CUDA_TRY(cudaSetDevice(device));
    maxsim_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
        scores, reinterpret_cast<const long long*>(ids), nq, k_in, reinterpret_cast<const long long*>(chunk_to_doc), n_chunks,
        k_out, out_scores, reinterpret_cast<long long*>(out_doc_ids), out_counts);
    CUDA_TRY(cudaGetLastError());
    return B2S_OK;
}

// ---------------------------------------------------------------------------------------------
// cross-GPU exchange (exchange.cuh)
// ---------------------------------------------------------------------------------------------
B2S_API int b2s_exchange_create(b2s_index* idx, int world, int rank, int64_t slot_bytes, int max_nq,
                                void* ipc_handle_out) {
    if (!idx || world < 1 || rank < 0 || rank >= world || slot_bytes < 16 || max_nq < 1)
        return fail(B2S_ERR_INVALID, "bad arguments");
    if (world > 64) return fail(B2S_ERR_UNSUPPORTED, "at most 64 ranks");
    std::lock_guard<std::mutex> g(idx->mu);
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    exchange_release(idx);
    auto& ex = idx->ex;
    ex.world = world;
    ex.rank = rank;
    ex.max_nq = max_nq;
    ex.slot_stride = (slot_bytes + 255) / 256 * 256;
    ex.flags_off = 2ll * world * ex.slot_stride;
    ex.ll_off = ((ex.flags_off + (long long)2 * world * max_nq * (long long)sizeof(unsigned)) + 255) / 256 * 256;
    ex.ll_entries = std::min<long long>(ex.slot_stride / 12, 131072);   // fused calls are small: <= #SMs queries
    ex.bytes = (size_t)ex.ll_off + (size_t)2 * world * (size_t)ex.ll_entries * 24;
    CUDA_TRY(cudaMalloc((void**)&ex.local, ex.bytes));
    CUDA_TRY(cudaMemset(ex.local, 0, ex.bytes));
    CUDA_TRY(cudaHostAlloc((void**)&ex.status, sizeof(unsigned), cudaHostAllocMapped | cudaHostAllocPortable));
    *ex.status = 0u;
    {
        // the fused kernel's CTAs push before they wait, but a CTA that is not resident cannot push: keep the
        // fused grid within what the device holds at once (and never above one CTA per SM)
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_exchange_kernel<true>, kMergeThreads, 0));
        ex.max_fused_nq = std::min(idx->num_sms, per_sm * idx->num_sms);
    }
    CUDA_TRY(cudaMalloc((void**)&ex.peers_dev, sizeof(unsigned char*) * world));
    CUDA_TRY(cudaDeviceSynchronize());
    if (ipc_handle_out) {
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, ex.local));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        memcpy(ipc_handle_out, &h, 64);
    }
    return B2S_OK;
}

B2S_API void* b2s_exchange_local(const b2s_index* idx) { return idx ? idx->ex.local : nullptr; }

B2S_API int b2s_exchange_connect(b2s_index* idx, const void* handles, int raw_pointers) {
    if (!idx || !handles) return fail(B2S_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> g(idx->mu);
    auto& ex = idx->ex;
    if (!ex.local) return fail(B2S_ERR_INVALID, "b2s_exchange_create was not called");
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    std::vector<unsigned char*> peers(ex.world, nullptr);
    for (int r = 0; r < ex.world; ++r) {
        if (r == ex.rank) {
            peers[r] = ex.local;
        } else if (raw_pointers) {
            peers[r] = reinterpret_cast<unsigned char* const*>(handles)[r];   // same-process ranks (tests)
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, reinterpret_cast<const unsigned char*>(handles) + (size_t)r * 64, 64);
            void* m = nullptr;
            CUDA_TRY(cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
            ex.opened.push_back(m);
            peers[r] = reinterpret_cast<unsigned char*>(m);
        }
        if (!peers[r]) return fail(B2S_ERR_INVALID, "null peer buffer");
    }
    CUDA_TRY(cudaMemcpy(ex.peers_dev, peers.data(), sizeof(unsigned char*) * ex.world, cudaMemcpyHostToDevice));
    ex.connected = true;
    return B2S_OK;
}

B2S_API int b2s_exchange_status(b2s_index* idx) {
    if (!idx || !idx->ex.status) return 0;
    cudaSetDevice(idx->device);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (int)*(volatile unsigned*)idx->ex.status;
}

static int search_sharded_impl(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k,
                               float* out_scores, int64_t* out_ids, void* cuda_stream, int phase) {
    auto& ex = idx->ex;
    if (!ex.connected) return fail(B2S_ERR_INVALID, "exchange is not connected");
    if (nq > ex.max_nq || b2s_packed_bytes(nq, k) > ex.slot_stride)
        return fail(B2S_ERR_UNSUPPORTED, "sharded search: (nq, k) exceeds the exchange buffer; use the all-gather path");
    if (idx->opt_keep_f32 && idx->rows_f32)
        return fail(B2S_ERR_UNSUPPORTED, "sharded search: fp32 re-ranking (keep_f32) is not applied before the peer "
                                         "exchange; use the all-gather path");
    if (const unsigned bad = *(volatile unsigned*)ex.status)
        return fail(B2S_ERR_CUDA, "sharded search: call #" + std::to_string(bad) + " timed out waiting for a peer rank; "
                                  "its results were invalid (re-create the exchange)");
    int rc = use_device(idx);
    if (rc != B2S_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    if ((rc = guard_stream(idx, s)) != B2S_OK) return rc;
    ExchangeArgs a;
    memset(&a, 0, sizeof(a));
    a.peer_base = ex.peers_dev;
    a.local_base = ex.local;
    a.world = ex.world;
    a.rank = ex.rank;
    a.slot_stride = ex.slot_stride;
    a.flags_off = ex.flags_off;
    a.max_nq = ex.max_nq;
    a.nq = nq;
    a.timeout_cycles = ex.timeout_ms * 2000000ll;   // SM clock ~2 GHz: ranks may be skewed by host work, but a
                                                    // missing peer must not hang the GPU (the status word reports it)
    a.status = ex.status;
    a.out_scores = out_scores;
    a.out_ids = reinterpret_cast<long long*>(out_ids);
    if (phase != 2) {
        a.seq = ++ex.seq;
        // one kernel when every CTA of the merge is co-resident on every rank, else push / wait split
        idx->ex_fused = (phase == 0 && nq <= ex.max_fused_nq) ? 1 : 0;
        a.use_ll = (idx->ex_fused && idx->opt_exchange_ll && nq * (int64_t)k <= ex.ll_entries &&
                    (int64_t)ex.world * k <= kMergeSortCap) ? 1 : 0;
        a.ll_off = ex.ll_off;
        a.ll_entries = ex.ll_entries;
        idx->ex_call = &a;
        rc = search_impl(idx, queries, q_dtype, nq, k, out_scores, out_ids, s);
        idx->ex_call = nullptr;
        if (rc != B2S_OK) {
            if (idx->stats.kernel_launches == 0) --ex.seq;   // nothing was enqueued: the ranks stay in step
            return rc;
        }
        if (idx->ex_fused || phase == 1) return B2S_OK;
    } else {
        a.seq = ex.seq;
    }
    exchange_wait_merge_kernel<<<(unsigned)nq, kMergeThreads, 0, s>>>(a, k);
    CUDA_TRY(cudaGetLastError());
    idx->stats.kernel_launches++;
    return B2S_OK;
}

B2S_API int b2s_search_sharded_device(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k,
                                      float* out_scores, int64_t* out_ids, void* cuda_stream, int phase,
                                      unsigned flags) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "nq and k must be >= 0");
    if (nq == 0 || k == 0) return B2S_OK;
    if (phase < 0 || phase > 2) return fail(B2S_ERR_INVALID, "phase must be 0 (whole call), 1 (push) or 2 (wait+merge)");
    if (flags & ~(unsigned)B2S_SEARCH_STABLE_QUERIES) return fail(B2S_ERR_INVALID, "unknown search flag");
    std::lock_guard<std::mutex> g(idx->mu);
    idx->call_flags = phase == 0 ? flags : 0u;
    const int rc = search_sharded_impl(idx, queries, q_dtype, nq, k, out_scores, out_ids, cuda_stream, phase);
    idx->call_flags = 0;
    return rc;
}

B2S_API int b2s_search_sharded(b2s_index* idx, const float* queries, int64_t nq, int k, float* out_scores,
                               int64_t* out_ids) {
    if (!idx) return fail(B2S_ERR_INVALID, "null index");
    if (nq < 0 || k < 0) return fail(B2S_ERR_INVALID, "nq and k must be >= 0");
    if (nq == 0 || k == 0) return B2S_OK;
    if (!queries || !out_scores || !out_ids) return fail(B2S_ERR_INVALID, "null buffer");
    std::lock_guard<std::mutex> g(idx->mu);
    return search_host(idx, queries, nq, k, out_scores, out_ids, true);
}

B2S_API int b2s_last_stats(const b2s_index* idx, b2s_stats* out) {
    if (!idx || !out) return fail(B2S_ERR_INVALID, "bad arguments");
    *out = idx->stats;
    if (idx->ev_valid) {
        // the caller must have synchronised the stream the search ran on
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, idx->ev[1], idx->ev[2]) == cudaSuccess) out->dominant_ms = ms;
        if (cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[3]) == cudaSuccess) out->total_ms = ms;
        cudaGetLastError();
    }
    return B2S_OK;
}

}  // extern "C"

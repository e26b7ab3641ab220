// scan_topk.cuh -- K1: small-batch (1..4 queries) exact inner-product scan with fused top-k.
//
// Replaces faiss Index.search at nq = 1 behind FAISSIndexBuilder.search
// (/root/reference/src/serve/app.py:293-295) and the per-query full-row score + argsort of
// /root/reference/src/kd/eval.py:75,86.
//
// HBM-bound: one pass over the bf16 corpus (dim*2 bytes per row), nothing written back but the
// candidates (<= k keys per query).  Mapping: a HALF-WARP owns one row; lane l (0..15) reads the
// 16-byte chunks l, l+16, l+32, ... of that row with 128-bit ld.global.nc.L1::no_allocate loads, so a
// warp instruction covers two rows x 256 contiguous bytes.  Each warp keeps U row-pairs (2U rows,
// U*CPL independent 16-byte loads per lane) in flight per iteration.  bf16 -> fp32 is a shift / mask,
// products are accumulated in fp32 against the query held in registers as fp32 (the query is NOT
// rounded to bf16 on this path), a 4-step xor-shuffle finishes the dot product inside the half-warp.
// Scores that reach the current k-th best are kept, everything else is dropped in registers, so the
// [N] score vector never exists in memory.
//
// Work distribution: the first S*G units (G = grid, a unit = 8 warps x 2U rows = 48 KB at dim 384) are
// dealt statically, CTA b takes units b, b + G, ... so that the whole grid reads one compact ~14 MB
// window at any moment; the LAST few units per CTA are handed out dynamically, 2U rows per atomic
// ticket (fetched one iteration ahead), so every warp of the GPU finishes within one iteration
// (~2 us) of every other -- the end of a 0.12 ms shard scan at 8 GPUs is not a ragged tail.
//
// Two ways of keeping the survivors (select_mode):
//   0  per-CTA shared-memory lists (select.cuh), published sorted; the last CTA to finish (or a
//      separate merge kernel) merges the G lists.  Any k <= 2048.
//   1  "cascade" (k <= 16, large shards): after a short phase A with the shared-memory lists every CTA
//      offers its local top-k to ONE global array of k sorted slots per query (lock-free insertion:
//      a chain of 64-bit atomicMax that carries the smaller key downwards), and from then on the
//      slots' k-th key IS every CTA's threshold: a row that beats it is inserted straight away
//      (~k ln(1/phase A fraction) insertions per query over the whole GPU).  When the scan ends the
//      answer already sits sorted in the slots: the last CTA only reads k keys -- no merge.
//      The switch from phase A to the slots happens behind one CTA barrier: warp q sorts query q's list,
//      adopts the better of its k-th key and the slots' k-th key, and walks the insertion chain while the
//      other warps stream on (a barrier-free warp-by-warp switch exists as an A/B variant: no faster, and
//      it sends 2.5x as many offers to the slots).
//
// Consecutive launches: the state launches share (done ticket, dynamic-tail counters, cascade slots)
// exists TWICE (ScanCtl[2]) and is used alternately; the last CTA of a launch resets its set and
// publishes that in the set's epoch word.  With B2S_SEARCH_STABLE_QUERIES and the cascade select a
// launch therefore depends on its predecessor in ONE place only: its last CTA waits for the
// predecessor grid before it writes outputs / talks to the peer ranks.  Everything else -- the whole
// scan -- overlaps the predecessor's ragged end, top-k read-out and NVLink exchange.
#pragma once
#include "exchange.cuh"
#include "merge_topk.cuh"
#include "select.cuh"

namespace b2s {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanInlineFloats = 768;   // queries of a host call travel inside the kernel parameters (<= 3 KB)
constexpr int kScanMaxNQ = 4;            // queries one launch holds in registers
constexpr int kCascadeMaxK = 16;         // slots per query of the cascade select
constexpr int kCascadeStormOffers = 12;  // offers to the slots one warp may make in phase B before its CTA falls back to lists
constexpr unsigned kCascadeStormGrid = 1024u;   // ... and offers of the whole grid (random data: ~200 per search)
constexpr int kWorkCounterStride = 32;   // words between the dynamic tail's ticket counters (one 128-byte line each)
constexpr int kTraceWords = 16;          // header of the trace buffer, then kTraceArrays arrays of kTraceStride per-CTA stamps:
constexpr int kTraceStride = 512;        //   0 start, 1 scan end, 2 transition begin, 3 transition end, 4 static part end, 5 SM id
constexpr int kTraceArrays = 6;

// State shared by consecutive launches; two sets, launch n of a handle uses set n & 1.
struct ScanCtl {
    unsigned done;                                      // CTAs that have finished (ticket); the last one resets the set
    unsigned epoch;                                     // launches that have used AND reset this set (release store)
    unsigned offers;                                    // cascade: phase B offers of the whole grid so far (storm guard)
    unsigned pad[29];
    unsigned work[kScanWarps * kWorkCounterStride];     // dynamic-tail ticket counters, one 128-byte line each
    u64 gslots[kScanMaxNQ * kCascadeMaxK];              // cascade select: running global top-k, sorted descending
};

struct ScanParams {
    const uint4* corpus;   // bf16 rows, row-major, dim*2 bytes each (16-byte aligned)
    const float* queries;  // fp32 [nq_total, dim]; this launch uses rows q_begin .. q_begin+NQ-1
    long long n_rows;      // rows in the shard
    int q_begin;
    int nq_valid;          // how many of the NQ register queries are real (others masked)
    int k;
    int cap;               // list capacity (power of two)
    int unit_stride;       // 1 = every unit; S = every S-th unit (threshold-seeding pre-pass)
    const u64* seed_keys;  // optional [nq_total] initial thresholds (0 = none), or nullptr
    u64* lists;            // out [gridDim.x, nq_lists, cap]
    int* counts;           // out [gridDim.x, nq_lists]
    int nq_lists;          // stride (in queries) of the list arrays
    int pdl_late_wait;     // 1: the caller guarantees that the query buffer was complete before the PREVIOUS
                           // kernel of the stream was enqueued (B2S_SEARCH_STABLE_QUERIES), so nothing this
                           // kernel reads early is produced by its predecessor: the PDL wait is deferred to
                           // the first point where state shared with the predecessor is touched (cascade:
                           // the phase A -> B transition; lists: before the dynamic tail / the write-out)
    // Fused tail: the LAST CTA to finish (ticket from ctl->done) produces the final top-k itself -- no
    // separate merge kernel, no kernel boundary.  1 = write the final top-k (mp.out_*); 2 = sharded
    // search: push to the peers, wait, merge (ex).
    int fused_tail;
    ScanCtl* ctl;             // this launch's control set (zero when its epoch reads ctl_expect)
    unsigned ctl_expect;      // epoch value that says the set's previous user has reset it
    int ctl_bump;             // 1: publish expect + 1 when the set has been reset.  0 (launch captured into a CUDA graph,
                              // replayed any number of times): the launch is ordered by full grid dependencies on both
                              // sides, resets the set and leaves the epoch alone
    MergeParams mp;
    ExchangeArgs ex;
    // dynamic tail: rows [dyn_begin, n_rows) are handed out 2U rows per ticket (dyn_begin == n_rows: all static)
    long long dyn_begin;      // (ctl->work: kScanWarps counters, one per region of the dynamic tail -- 2400 warps on
                              // one address would be bound by L2 atomic throughput)
    // cascade select
    int select_mode;          // 0 lists, 1 cascade
    int phase_a_iters;        // iterations every warp runs against the shared-memory lists first ...
    int phase_a_stagger;      // ... plus blockIdx % phase_a_stagger: the CTAs reach the slots a few at a time, so
                              // that the peek of a late CTA already sees (and is filtered by) the early ones' keys
    int transition_mode;      // 0 (default): CTA-wide switch behind a barrier; 1: warp-by-warp switch without a barrier
                              // (A/B variant: no better in steady state, and more offers to the slots)
    int peek_every;           // phase B: warp 0 re-reads the slots' k-th key every this many iterations (0 = never)
    // L2 prefetch of this warp's first iterations, issued BEFORE the PDL wait: the corpus is immutable
    // while searches are in flight, so the idle DRAM time of the predecessor's tail is put to use
    int prefetch_iters;
    // 1 (default): the dependent launch is triggered at the START of the kernel, so the next kernel of the stream
    // is already queued behind this one and its CTAs take over SM slots the moment this grid's CTAs retire (no
    // launch latency after the last CTA's scan).  Safe: every CTA of this grid is resident by then, and a
    // dependent kernel touches state shared with this one only behind its own griddepcontrol.wait, which returns
    // when this grid has completed.  0: trigger after the scan (A/B switch).
    int early_trigger;
    unsigned long long* trace;   // optional (option "trace"): globaltimer stamps, see kTraceWords
    // Host-buffer searches of 1-2 queries: the query rides in the kernel parameters (constant bank) -- no
    // staging copy, no H2D operation in front of the kernel.
    int use_inline;
    // Host-buffer calls whose outputs go straight to mapped pinned host memory: the last CTA stores host_seq here
    // (system-scope fence first) after the last output, and the host thread, which is spinning on the word, has
    // the answer a few microseconds before the grid has been torn down and the stream reports completion.
    unsigned* host_flag;
    unsigned host_seq;
    alignas(16) float q_inline[kScanInlineFloats];
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// pinned in program order among the (volatile) corpus loads: issued where it is written, consumed an iteration later
__device__ __forceinline__ u64 ldcg_pinned_u64(const u64* p) {
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ float dot8(const uint4& w, const float* q, float acc) {
    acc = fmaf(__uint_as_float(w.x << 16), q[0], acc);
    acc = fmaf(__uint_as_float(w.x & 0xffff0000u), q[1], acc);
    acc = fmaf(__uint_as_float(w.y << 16), q[2], acc);
    acc = fmaf(__uint_as_float(w.y & 0xffff0000u), q[3], acc);
    acc = fmaf(__uint_as_float(w.z << 16), q[4], acc);
    acc = fmaf(__uint_as_float(w.z & 0xffff0000u), q[5], acc);
    acc = fmaf(__uint_as_float(w.w << 16), q[6], acc);
    acc = fmaf(__uint_as_float(w.w & 0xffff0000u), q[7], acc);
    return acc;
}

// if (pred) old = atomicMax(addr, v); else old = v;   (no branch: see the note on lane-uniform control flow in select.cuh)
__device__ __forceinline__ u64 atom_max_u64_if(u64* addr, u64 v, bool pred) {
    u64 old = v;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %3, 0;\n\t"
        "@p atom.global.max.u64 %0, [%1], %2;\n\t"
        "}\n"
        : "+l"(old)
        : "l"(addr), "l"(v), "r"((uint32_t)pred)
        : "memory");
    return old;
}

// Lock-free insertion of (unique) keys into k global slots kept sorted descending -- warp-collective, every
// lane may bring one key (`active`), control flow is warp-uniform throughout.  The peek finds, per lane, the
// first slot below its key (slots only ever grow, so the slots above it stay above it for ever); from there
// atomicMax leaves the larger key in the slot and the smaller one is carried to the next slot.  At quiescence
// slot j holds the (j+1)-th largest key ever offered, whatever the interleaving of lanes, warps and CTAs.
__device__ __forceinline__ void cascade_insert_warp(u64* slots, int k, bool active, u64 key) {
    constexpr unsigned kFull = 0xffffffffu;
    const ulonglong2* sv = reinterpret_cast<const ulonglong2*>(slots);
    ulonglong2 v[kCascadeMaxK / 2];
#pragma unroll
    for (int i = 0; i < kCascadeMaxK / 2; ++i) v[i] = __ldcg(sv + i);
    int j0 = k;   // first slot this lane's key has to visit
#pragma unroll
    for (int i = kCascadeMaxK / 2 - 1; i >= 0; --i) {
        j0 = (2 * i + 1 < k && v[i].y < key) ? 2 * i + 1 : j0;
        j0 = (2 * i < k && v[i].x < key) ? 2 * i : j0;
    }
    active = active && j0 < k;
    j0 = active ? j0 : k;
    int jmin = j0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) jmin = min(jmin, __shfl_xor_sync(kFull, jmin, off));
    for (int j = jmin; j < k; ++j) {   // uniform bounds
        const bool doit = active && j >= j0;
        const u64 old = atom_max_u64_if(slots + j, key, doit);
        const bool displaced = doit && old < key;
        active = active && !(displaced && old == 0ull);   // an empty slot took it: done
        key = displaced ? old : key;                        // carry the smaller key down
        if (!__any_sync(kFull, active)) break;              // uniform
    }
}

// CPL = 16-byte chunks per lane = dim / 128; NQ = queries held in registers; U = row pairs in flight.
template <int CPL, int NQ, int U>
__global__ void __launch_bounds__(kScanThreads, (NQ <= 2 ? 2 : 1))
scan_topk_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* entries = reinterpret_cast<u64*>(smem_raw);  // [NQ][cap]
    __shared__ u64 s_thr_key[NQ];
    __shared__ float s_thr[NQ];
    __shared__ int s_count[NQ];
    __shared__ int s_lock[NQ];
    __shared__ int s_storm;                      // cascade: this CTA has fallen back to its shared-memory lists (see offer_global)
    __shared__ int s_arrived;                    // cascade: warps of this CTA that have left phase A
    __shared__ u64 s_casc[NQ][kCascadeMaxK];     // cascade: the last CTA's copy of the slots
    auto list_of = [&](int q) { return ListRef{&s_thr_key[q], &s_thr[q], &s_count[q], &s_lock[q]}; };

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int half = lane >> 4;
    const int hl = lane & 15;
    constexpr int kChunksPerRow = CPL * 16;  // uint4 per row
    constexpr int kRowsPerIter = 2 * U;
    constexpr long long kUnitRows = (long long)kScanWarps * kRowsPerIter;
    constexpr unsigned kFull = 0xffffffffu;

    const long long row_end = p.n_rows;
    const long long last_row = p.n_rows - 1;
    const long long static_end = p.dyn_begin;     // == row_end when nothing is dealt dynamically
    const bool dynamic = static_end < row_end;
    const bool cascade = p.select_mode == 1;
    const long long step = kUnitRows * (long long)gridDim.x * p.unit_stride;
    long long base = kUnitRows * (long long)blockIdx.x * p.unit_stride + (long long)warp * kRowsPerIter;

    // L2 prefetch before the PDL wait (lane i: this warp's i-th static iteration)
    if (lane < p.prefetch_iters) {
        const long long b = base + (long long)lane * step;
        if (b + kRowsPerIter <= static_end)
            prefetch_l2_bulk(p.corpus + b * kChunksPerRow, (uint32_t)(kRowsPerIter * kChunksPerRow * 16));
    }
    if (p.early_trigger) grid_dep_launch();
    if (!p.pdl_late_wait) grid_dep_wait();
    auto stamp = [&](int which) {
        if (p.trace != nullptr && tid == 0) p.trace[kTraceWords + which * kTraceStride + blockIdx.x] = globaltimer_ns();
    };
    stamp(0);
    if (p.trace != nullptr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[kTraceWords + 5 * kTraceStride + blockIdx.x] = smid;
    }
    if (tid == 0) {
        s_arrived = 0;
        s_storm = 0;
    }
    if (tid < NQ) {
        u64 seed = 0ull;
        if (p.seed_keys != nullptr && tid < p.nq_valid) seed = p.seed_keys[p.q_begin + tid];
        if (tid < p.nq_valid) list_init(list_of(tid), seed);
        else list_disable(list_of(tid));  // masked query slot: nothing can pass
    }

    // query -> registers (fp32): lane hl keeps elements [(hl + 16 j) * 8, +8) for j < CPL
    float qreg[NQ][CPL][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int qi = p.q_begin + (q < p.nq_valid ? q : 0);
        const float* qbase = p.use_inline ? p.q_inline : p.queries;   // parameter space or global memory
        const float4* qp = reinterpret_cast<const float4*>(qbase + (size_t)qi * (CPL * 128));
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float4 a = qp[(hl + 16 * j) * 2];
            float4 b = qp[(hl + 16 * j) * 2 + 1];
            qreg[q][j][0] = a.x; qreg[q][j][1] = a.y; qreg[q][j][2] = a.z; qreg[q][j][3] = a.w;
            qreg[q][j][4] = b.x; qreg[q][j][5] = b.y; qreg[q][j][6] = b.z; qreg[q][j][7] = b.w;
        }
    }
    __syncthreads();

    // Rare path, kept out of line of the hot loop: pick the row this lane speaks for (lane hl < U of each
    // half-warp owns row base + 2*hl + half) and re-test it.  List mode: append under the list lock.
    auto offer_list = [&](long long b, const float (&acc)[NQ][U]) {
        const long long myrow = b + 2 * hl + half;
        const bool row_ok = (hl < U) && (myrow < row_end);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float s = acc[q][0];
#pragma unroll
            for (int i = 1; i < U; ++i) s = (hl == i) ? acc[q][i] : s;
            const bool pass = row_ok && (s >= *(volatile float*)&s_thr[q]);
            if (__any_sync(kFull, pass)) {
                list_append_warp(list_of(q), entries + (size_t)q * p.cap, p.cap, p.k, pass,
                                 make_key(s, (uint32_t)myrow), lane);
            }
        }
    };
    // every lane stores the same values: no lane predicate, no branch
    auto adopt_bound = [&](int q, u64 kth) {
        if (kth > *(volatile u64*)&s_thr_key[q]) {   // uniform (all lanes hold the same kth)
            *(volatile u64*)&s_thr_key[q] = kth;
            *(volatile float*)&s_thr[q] = key_score(kth);
        }
    };
    // Cascade mode (phase B): straight into the global slots, then pick up the slots' new k-th key.
    // Guard against an insertion storm (a corpus whose scores RISE along the scan order makes every row beat the
    // running k-th: each offer is a chain of atomics on the same k words for the whole GPU -- measured 3.0 ms
    // instead of 0.13 ms on a shard sorted by score): a warp that has made more than kCascadeStormOffers offers,
    // or that sees the whole grid past kCascadeStormGrid offers (random data needs ~200), switches its CTA back
    // to the shared-memory lists for the rest of the scan.  The lists are emptied first
    // (their old content went to the slots at the phase switch, keys in the slots must stay unique) and are
    // offered to the slots once more when the CTA has finished scanning.
    int my_offers = 0;
    auto offer_global = [&](long long b, const float (&acc)[NQ][U]) {
        if (*(volatile int*)&s_storm) {   // uniform
            offer_list(b, acc);
            return;
        }
        const long long myrow = b + 2 * hl + half;
        const bool row_ok = (hl < U) && (myrow < row_end);
        bool offered = false;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float s = acc[q][0];
#pragma unroll
            for (int i = 1; i < U; ++i) s = (hl == i) ? acc[q][i] : s;
            const bool pass = row_ok && (s >= *(volatile float*)&s_thr[q]);
            if (__any_sync(kFull, pass)) {
                u64* slots = p.ctl->gslots + q * kCascadeMaxK;
                const u64 key = make_key(s, (uint32_t)myrow);
                const bool ins = pass && key > *(volatile u64*)&s_thr_key[q];
                cascade_insert_warp(slots, p.k, ins, key);
                offered = offered || __any_sync(kFull, ins);
                if (p.trace != nullptr) {   // diagnostics: warp-level offers / keys sent to the slots in phase B
                    const unsigned m = __ballot_sync(kFull, ins);
                    if (lane == 0) {
                        atomicAdd(p.trace + 8, 1ull);
                        atomicAdd(p.trace + 9, (unsigned long long)__popc(m));
                    }
                }
                __syncwarp();
                adopt_bound(q, __ldcg(slots + p.k - 1));
                __syncwarp();
            }
        }
        unsigned grid_offers = 0u;
        if (offered) {
            if (lane == 0) grid_offers = atomicAdd(&p.ctl->offers, 1u);
            grid_offers = __shfl_sync(kFull, grid_offers, 0);
        }
        if (offered && (++my_offers > kCascadeStormOffers || grid_offers >= kCascadeStormGrid) &&
            !*(volatile int*)&s_storm) {   // uniform
            for (int q = 0; q < p.nq_valid; ++q) {
                const ListRef st = list_of(q);
                spin_acquire(st.lock, lane);
                *(volatile int*)st.count = 0;        // every lane stores the same value
                spin_release(st.lock, lane);
            }
            __threadfence_block();
            *(volatile int*)&s_storm = 1;
            __syncwarp();
        }
    };

    // One iteration on 2U full rows: rp = this lane's pointer to chunk hl of row (b + half).
    auto scan_rows = [&](const uint4* rp, long long b, auto&& offer) {
        uint4 w[U][CPL];
#pragma unroll
        for (int i = 0; i < U; ++i)
#pragma unroll
            for (int j = 0; j < CPL; ++j) w[i][j] = ldg_stream(rp + (2 * i) * kChunksPerRow + 16 * j);
        float acc[NQ][U];
#pragma unroll
        for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int i = 0; i < U; ++i) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < CPL; ++j) a = dot8(w[i][j], qreg[q][j], a);
                acc[q][i] = a;
            }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1)
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int i = 0; i < U; ++i) acc[q][i] += __shfl_xor_sync(kFull, acc[q][i], off);
        // after the butterfly every lane of a half-warp holds all U sums of its half
        bool hot = false;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float t = *(volatile float*)&s_thr[q];
#pragma unroll
            for (int i = 0; i < U; ++i) hot = hot || (acc[q][i] >= t);
        }
        if (__any_sync(kFull, hot)) offer(b, acc);
        __syncwarp();
    };
    // The (at most one) partial group of rows at the end of the shard: loads clamped to the last row.
    auto scan_rows_clamped = [&](long long b, auto&& offer) {
        float acc[NQ][U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
            long long r = b + 2 * i + half;
            if (r > last_row) r = last_row;
            const uint4* tp = p.corpus + r * kChunksPerRow + hl;
            uint4 w[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) w[j] = ldg_stream(tp + 16 * j);
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float a = 0.f;
#pragma unroll
                for (int j = 0; j < CPL; ++j) a = dot8(w[j], qreg[q][j], a);
                acc[q][i] = a;
            }
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1)
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int i = 0; i < U; ++i) acc[q][i] += __shfl_xor_sync(kFull, acc[q][i], off);
        offer(b, acc);
        __syncwarp();
    };
    // Cascade phase B: warp 0 re-reads the slots' k-th key once per iteration (other CTAs raise it); the
    // load flies during the iteration's corpus loads.
    // (all lanes load the same word -- query `peek_q`, rotating -- and later store the same values: uniform)
    int peek_in = p.peek_every;   // iterations until warp 0 peeks again (0: never -- inserts refresh the bound anyway)
    int peek_q = 0, peeked_q = 0;
    auto bound_peek = [&]() -> u64 {
        u64 g = 0ull;
        if (warp == 0 && p.peek_every > 0 && --peek_in == 0) {   // uniform
            peek_in = p.peek_every;
            peeked_q = peek_q;
            g = ldcg_pinned_u64(p.ctl->gslots + peek_q * kCascadeMaxK + p.k - 1);
            peek_q = peek_q + 1 < p.nq_valid ? peek_q + 1 : 0;
        }
        return g;
    };
    auto bound_apply = [&](u64 g) {
        if (g != 0ull) adopt_bound(peeked_q, g);   // uniform: every lane of warp 0 holds the same g
    };

    // lane's pointer to chunk hl of row (base + half); advanced by `step` rows per iteration
    const uint4* rp = p.corpus + (base + half) * kChunksPerRow + hl;
    const long long rp_step = step * kChunksPerRow;

    // The control set is ours once its epoch says that the launch before the previous one has reset it (with
    // the early PDL wait the predecessor grid has completed and this passes at once).  Warp-collective.
    auto ctl_ready = [&]() {
        if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(&p.ctl->epoch) != p.ctl_expect && clock64() - t0 < 4000000000ll) __nanosleep(64);
        }
        __syncwarp();
    };

    // ---- static part, against the shared-memory lists (cascade: only the first phase_a_iters) ------
    {
        int it = 0;
        const int it_end = cascade ? p.phase_a_iters + (int)(blockIdx.x % (unsigned)p.phase_a_stagger) : 0x7fffffff;
        for (; base + kRowsPerIter <= static_end && it < it_end; base += step, rp += rp_step, ++it)
            scan_rows(rp, base, offer_list);
    }
    u64 g_bound = 0ull;   // cascade: the slots' k-th key as peeked one iteration ago (warp 0, lane = query)
    if (cascade) {
        // ---- phase A -> B, warp by warp, no CTA barrier: a warp that has done its phase A iterations goes on
        // against the global slots at once (its bound: the CTA's list threshold, raised to the slots' k-th key
        // whenever it looks).  The LAST warp of the CTA to arrive knows that nobody appends to the shared-memory
        // lists any more: it alone sorts them and walks the chains of atomics that offer the CTA's local top-k to
        // the slots, while the other seven warps keep streaming.
        if (p.transition_mode == 0) {
            // CTA-wide switch: barrier, warp q sorts query q's list and adopts the better of its k-th
            // key and the slots' k-th key, barrier, the other warps go on while warp q walks the insertion chain
            __syncthreads();
            stamp(2);
            if (p.pdl_late_wait) ctl_ready();
            int c_mine = 0;
            if (warp < p.nq_valid) {
                c_mine = list_compact_warp(list_of(warp), entries + (size_t)warp * p.cap, p.cap, p.k, lane);
                __syncwarp();
                adopt_bound(warp, __ldcg(p.ctl->gslots + warp * kCascadeMaxK + p.k - 1));
                __syncwarp();
            }
            __syncthreads();
            if (warp < p.nq_valid) {
                u64* slots = p.ctl->gslots + warp * kCascadeMaxK;
                cascade_insert_warp(slots, p.k, lane < c_mine, entries[(size_t)warp * p.cap + (lane < c_mine ? lane : 0)]);
                if (p.trace != nullptr && lane == 0) atomicAdd(p.trace + 10, (unsigned long long)c_mine);
                __syncwarp();
                adopt_bound(warp, __ldcg(slots + p.k - 1));
                __syncwarp();
                if (warp == 0) stamp(3);
            }
        } else {
        if (warp == 0) stamp(2);
        if (p.pdl_late_wait) ctl_ready();
        __threadfence_block();
        int arrived = 0;
        if (lane == 0) arrived = atomicAdd(&s_arrived, 1);
        arrived = __shfl_sync(kFull, arrived, 0);
        if (arrived == kScanWarps - 1) {        // warp-uniform
            __threadfence_block();
            for (int q = 0; q < p.nq_valid; ++q) {
                u64* slots = p.ctl->gslots + q * kCascadeMaxK;
                u64* mine = entries + (size_t)q * p.cap;
                const int c_mine = list_compact_warp(list_of(q), mine, p.cap, p.k, lane);
                adopt_bound(q, __ldcg(slots + p.k - 1));
                __syncwarp();
                const u64 key = mine[lane < c_mine ? lane : 0];
                cascade_insert_warp(slots, p.k, lane < c_mine && key > *(volatile u64*)&s_thr_key[q], key);
                if (p.trace != nullptr && lane == 0) atomicAdd(p.trace + 10, (unsigned long long)c_mine);
                __syncwarp();
                adopt_bound(q, __ldcg(slots + p.k - 1));
                __syncwarp();
            }
            if (p.trace != nullptr && lane == 0) p.trace[kTraceWords + 3 * kTraceStride + blockIdx.x] = globaltimer_ns();
        } else {
            for (int q = 0; q < p.nq_valid; ++q) adopt_bound(q, __ldcg(p.ctl->gslots + q * kCascadeMaxK + p.k - 1));
            __syncwarp();
        }
        }
        for (; base + kRowsPerIter <= static_end; base += step, rp += rp_step) {
            const u64 g = bound_peek();     // lands during this iteration's loads, applied at the start of the next
            bound_apply(g_bound);
            scan_rows(rp, base, offer_global);
            g_bound = g;
        }
    } else {
        // static tail (only when nothing is dealt dynamically): this warp's partial group, if any
        if (!dynamic && base < row_end) scan_rows_clamped(base, offer_list);
        // lists mode: the candidate lists / counts workspace is read by the predecessor's merge until it completes
        if (p.pdl_late_wait && dynamic) grid_dep_wait();
        if (p.pdl_late_wait && dynamic) ctl_ready();
    }

    if (p.trace != nullptr && warp == 0 && lane == 0) p.trace[kTraceWords + 4 * kTraceStride + blockIdx.x] = globaltimer_ns();
    // ---- dynamic tail: 2U rows per ticket, the next ticket is fetched before the current rows are read.  The
    // chunks are split into kScanWarps regions with a counter each; a warp starts in the region of its own
    // index and moves on to the next one when a region is drained.
    if (dynamic) {
        const unsigned n_chunks = (unsigned)((row_end - static_end + kRowsPerIter - 1) / kRowsPerIter);
        const unsigned per_region = (n_chunks + kScanWarps - 1) / kScanWarps;
        for (int r = 0; r < kScanWarps; ++r) {
            const unsigned region = (unsigned)((warp + r) % kScanWarps);
            const unsigned c0 = region * per_region;
            if (c0 >= n_chunks) continue;
            const unsigned cn = min(per_region, n_chunks - c0);
            unsigned* ctr = p.ctl->work + region * kWorkCounterStride;
            unsigned nxt = 0u;
            if (lane == 0) nxt = atomicAdd(ctr, 1u);
            nxt = __shfl_sync(kFull, nxt, 0);
            while (nxt < cn) {
                const long long b = static_end + (long long)(c0 + nxt) * kRowsPerIter;
                unsigned t = 0u;
                if (lane == 0) t = atomicAdd(ctr, 1u);
                const uint4* dp = p.corpus + (b + half) * kChunksPerRow + hl;
                if (cascade) {
                    const u64 g = bound_peek();
                    bound_apply(g_bound);
                    if (b + kRowsPerIter <= row_end) scan_rows(dp, b, offer_global);
                    else scan_rows_clamped(b, offer_global);
                    g_bound = g;
                } else {
                    if (b + kRowsPerIter <= row_end) scan_rows(dp, b, offer_list);
                    else scan_rows_clamped(b, offer_list);
                }
                nxt = __shfl_sync(kFull, t, 0);
            }
        }
    }
    __syncthreads();
    if (cascade && *(volatile int*)&s_storm) {
        // this CTA fell back to its lists (insertion storm): their content goes to the slots before the ticket
        if (warp < p.nq_valid) {
            u64* slots = p.ctl->gslots + warp * kCascadeMaxK;
            u64* mine = entries + (size_t)warp * p.cap;
            const int c_mine = list_compact_warp(list_of(warp), mine, p.cap, p.k, lane);
            __syncwarp();
            cascade_insert_warp(slots, p.k, lane < c_mine, mine[lane < c_mine ? lane : 0]);
        }
        __syncthreads();
    }
    stamp(1);
    if (!p.early_trigger) grid_dep_launch();   // the next kernel may be scheduled from here on

    if (!cascade) {
        if (p.pdl_late_wait && !dynamic) grid_dep_wait();   // the previous call's merge has finished reading the lists
        // final: leave exactly min(count, k) best keys per query, SORTED descending (the merge relies on
        // it: a list's first key is its maximum), then publish
        for (int q = warp; q < NQ; q += kScanWarps)
            list_compact_warp(list_of(q), entries + (size_t)q * p.cap, p.cap, p.k, lane);
        __syncthreads();
        for (int q = 0; q < p.nq_valid; ++q) {
            const int c = s_count[q];
            u64* dst = p.lists + ((size_t)blockIdx.x * p.nq_lists + (p.q_begin + q)) * p.cap;
            for (int i = tid; i < c; i += kScanThreads) dst[i] = entries[(size_t)q * p.cap + i];
            if (tid == 0) p.counts[(size_t)blockIdx.x * p.nq_lists + (p.q_begin + q)] = c;
        }
    }
    if (p.fused_tail == 0 && !dynamic) return;

    // ---- last CTA done (threadfence reduction pattern): resets the control set, runs the fused tail -----
    __shared__ MergeSmem sm;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        if (p.pdl_late_wait && !cascade && !dynamic) {   // (a set nobody has touched yet in this launch)
            const long long t0 = clock64();
            while (ld_acquire_gpu_u32(&p.ctl->epoch) != p.ctl_expect && clock64() - t0 < 4000000000ll) __nanosleep(64);
        }
        const unsigned ticket = atomicAdd(&p.ctl->done, 1u);
        s_last = ticket == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    unsigned long long* tr = (p.trace != nullptr && tid == 0) ? p.trace : nullptr;
    if (tr) {
        tr[0] = gridDim.x;
        tr[1] = globaltimer_ns();
        tr[11] = tr[8], tr[12] = tr[9], tr[13] = tr[10];   // this launch's diagnostics counters
        tr[8] = tr[9] = tr[10] = 0ull;
    }
    // Take what the set holds (cascade: the answer), reset it and hand it on: from the epoch store on, the
    // launch after the next one may use the set -- before this CTA has written a single output.
    if (cascade) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            u64* slots = p.ctl->gslots + q * kCascadeMaxK;
            if (q < p.nq_valid && tid < kCascadeMaxK) {
                s_casc[q][tid] = tid < p.k ? __ldcg(slots + tid) : 0ull;
                slots[tid] = 0ull;
            }
        }
    }
    if (tid == 0) {
        p.ctl->done = 0u;
        p.ctl->offers = 0u;
    }
    if (dynamic && tid < kScanWarps) p.ctl->work[tid * kWorkCounterStride] = 0u;
    __threadfence();
    __syncthreads();
    if (tid == 0 && p.ctl_bump) st_release_gpu_u32(&p.ctl->epoch, p.ctl_expect + 1u);
    if (p.fused_tail == 0) return;
    // The one dependency on the predecessor launch under B2S_SEARCH_STABLE_QUERIES: outputs are written and the
    // peer ranks are talked to in call order.
    if (p.pdl_late_wait) grid_dep_wait();
    for (int q = 0; q < p.nq_valid; ++q) {
        const int lq = p.q_begin + q;     // list / output index of this query inside the call
        int kk;
        if (cascade) {
            if (tid < kCascadeMaxK) sm.buf[tid] = s_casc[q][tid];
            __syncthreads();
            kk = 0;
            for (int i = 0; i < p.k; ++i) kk += s_casc[q][i] != 0ull;   // sorted: the non-empty slots are a prefix
        } else {
            const int m_sorted = merge_lists_sorted<kScanThreads>(p.mp, lq, sm);
            kk = m_sorted < p.mp.k ? m_sorted : p.mp.k;
        }
        if (tr) tr[2] = globaltimer_ns();
        if (p.fused_tail == 1) {
            for (int i = tid; i < p.mp.k; i += kScanThreads) {
                float s = -FLT_MAX;
                long long id = -1;
                if (i < kk) {
                    const u64 key = sm.buf[i];
                    s = key_score(key);
                    id = (long long)key_row(key) + p.mp.id_offset;
                }
                p.mp.out_scores[(size_t)lq * p.mp.k + i] = s;
                p.mp.out_ids[(size_t)lq * p.mp.k + i] = id;
            }
        } else {
            const long long gq = (long long)p.ex.q_offset + lq;
            exchange_fused<kScanThreads>(p.ex, p.mp, sm.buf, kk, gq, tr);
        }
        __syncthreads();
    }
    if (p.host_flag != nullptr && tid == 0) {
        __threadfence_system();                         // every output above (seen through the barrier) before the flag
        *(volatile unsigned*)p.host_flag = p.host_seq;
    }
    if (tr) tr[5] = globaltimer_ns();
}

}  // namespace b2s

// scan_topk.cuh -- K1: small-batch (1..4 queries) exact inner-product scan with fused top-k.
//
// Replaces faiss Index.search at nq = 1 behind FAISSIndexBuilder.search
// (/root/reference/src/serve/app.py:293-295) and the per-query full-row score + argsort of
// /root/reference/src/kd/eval.py:75,86.
//
// HBM-bound: one pass over the bf16 corpus (dim*2 bytes per row), nothing written back but the
// per-CTA candidate lists (<= capacity keys per query).  Mapping: a HALF-WARP owns one row; lane
// l (0..15) reads the 16-byte chunks l, l+16, l+32, ... of that row with 128-bit
// ld.global.nc.L1::no_allocate loads, so a warp instruction covers two rows x 256 contiguous bytes.
// Each warp keeps U row-pairs (2U rows, U*CPL independent 16-byte loads per lane) in flight per
// iteration.  bf16 -> fp32 is a shift / mask, products are accumulated in fp32 against the query
// held in registers as fp32 (the query is NOT rounded to bf16 on this path), a 4-step xor-shuffle
// finishes the dot product inside the half-warp.  Scores that reach the CTA's current k-th best
// are appended to the shared-memory candidate list (select.cuh); everything else is dropped in
// registers, so the [N] score vector never exists in memory.
#pragma once
#include "exchange.cuh"
#include "merge_topk.cuh"
#include "select.cuh"

namespace b2s {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanInlineFloats = 768;   // queries of a host call travel inside the kernel parameters (<= 3 KB)

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
    int pdl_late_wait;     // 1: nothing this kernel READS is produced by the preceding kernel of the
                           // stream, so the PDL wait is deferred to just before the list write-out
                           // (the scan of query i+1 then overlaps the merge of query i)
    // Fused tail: the LAST CTA to publish its lists (ticket from *done_counter) merges all lists of
    // the launch's queries itself -- no separate merge kernel, no kernel boundary.  1 = write the
    // final top-k (mp.out_*); 2 = sharded search: push to the peers, wait, merge (ex).
    int fused_tail;
    unsigned* done_counter;   // zero before the launch; the last CTA resets it
    MergeParams mp;
    ExchangeArgs ex;
    // Host-buffer searches of 1-2 queries: the query rides in the kernel parameters (constant bank) -- no
    // staging copy, no H2D operation in front of the kernel.
    int use_inline;
    alignas(16) float q_inline[kScanInlineFloats];
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
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
    auto list_of = [&](int q) { return ListRef{&s_thr_key[q], &s_thr[q], &s_count[q], &s_lock[q]}; };

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int half = lane >> 4;
    const int hl = lane & 15;
    constexpr int kChunksPerRow = CPL * 16;  // uint4 per row
    constexpr int kRowsPerIter = 2 * U;

    if (!p.pdl_late_wait) grid_dep_wait();
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

    // Grid-stride schedule: a "unit" is kScanWarps * 2U consecutive rows (48 KB at dim 384); CTA b
    // takes units b, b + G, b + 2G, ...  At any moment the whole grid therefore reads one compact
    // window of G units (~14 MB), which keeps DRAM pages and the 2 MB-page TLB hot -- unlike a
    // contiguous per-CTA partition, where G distant regions stream at once.
    const long long row_end = p.n_rows;
    const long long last_row = p.n_rows - 1;

    // Rare path, kept out of line of the hot loop: pick the row this lane speaks for (lane hl < U
    // of each half-warp owns row base + 2*hl + half), re-test it and append under the list lock.
    auto offer = [&](long long base, const float (&acc)[NQ][U]) {
        const long long myrow = base + 2 * hl + half;
        const bool row_ok = (hl < U) && (myrow < row_end);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float s = acc[q][0];
#pragma unroll
            for (int i = 1; i < U; ++i) s = (hl == i) ? acc[q][i] : s;
            const bool pass = row_ok && (s >= *(volatile float*)&s_thr[q]);
            if (__any_sync(0xffffffffu, pass)) {
                list_append_warp(list_of(q), entries + (size_t)q * p.cap, p.cap, p.k, pass,
                                 make_key(s, (uint32_t)myrow), lane);
            }
        }
    };

    constexpr long long kUnitRows = (long long)kScanWarps * kRowsPerIter;
    const long long step = kUnitRows * (long long)gridDim.x * p.unit_stride;
    long long base = kUnitRows * (long long)blockIdx.x * p.unit_stride + (long long)warp * kRowsPerIter;
    // lane's pointer to chunk hl of row (base + half); advanced by `step` rows per iteration
    const uint4* rp = p.corpus + (base + half) * kChunksPerRow + hl;
    const long long rp_step = step * kChunksPerRow;

    // ---- hot loop: all 2U rows of the unit exist; no clamps, no per-lane selects ----------------
    for (; base + kRowsPerIter <= row_end; base += step, rp += rp_step) {
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
                for (int i = 0; i < U; ++i) acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], off);
        // after the butterfly every lane of a half-warp holds all U sums of its half
        bool hot = false;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float t = *(volatile float*)&s_thr[q];
#pragma unroll
            for (int i = 0; i < U; ++i) hot = hot || (acc[q][i] >= t);
        }
        if (__any_sync(0xffffffffu, hot)) offer(base, acc);
        __syncwarp();
    }

    // ---- tail: the (at most one) partial unit of this warp, loads clamped to the last row --------
    if (base < row_end) {
        float acc[NQ][U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
            long long r = base + 2 * i + half;
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
                for (int i = 0; i < U; ++i) acc[q][i] += __shfl_xor_sync(0xffffffffu, acc[q][i], off);
        offer(base, acc);
        __syncwarp();
    }
    __syncthreads();
    grid_dep_launch();                          // the merge kernel may be scheduled: its launch latency hides here
    if (p.pdl_late_wait) grid_dep_wait();       // the previous call's merge has finished reading the lists

    // final: leave exactly min(count, k) best keys per query, SORTED descending (the merge kernel
    // relies on it: a list's first key is its maximum), then publish
    for (int q = warp; q < NQ; q += kScanWarps)
        list_compact_warp(list_of(q), entries + (size_t)q * p.cap, p.cap, p.k, lane);
    __syncthreads();
    for (int q = 0; q < p.nq_valid; ++q) {
        const int c = s_count[q];
        u64* dst = p.lists + ((size_t)blockIdx.x * p.nq_lists + (p.q_begin + q)) * p.cap;
        for (int i = tid; i < c; i += kScanThreads) dst[i] = entries[(size_t)q * p.cap + i];
        if (tid == 0) p.counts[(size_t)blockIdx.x * p.nq_lists + (p.q_begin + q)] = c;
    }
    if (p.fused_tail == 0) return;

    // ---- fused tail: last CTA done merges (threadfence reduction pattern) ------------------------
    __shared__ MergeSmem sm;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned ticket = atomicAdd(p.done_counter, 1u);
        s_last = ticket == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) *p.done_counter = 0u;   // ready for the next launch (stream-ordered after this kernel)
    for (int q = 0; q < p.nq_valid; ++q) {
        const int lq = p.q_begin + q;     // list / output index of this query inside the call
        const int m_sorted = merge_lists_sorted<kScanThreads>(p.mp, lq, sm);
        const int kk = m_sorted < p.mp.k ? m_sorted : p.mp.k;
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
            exchange_fused<kScanThreads>(p.ex, p.mp, sm.buf, kk, gq);
        }
        __syncthreads();
    }
}

}  // namespace b2s

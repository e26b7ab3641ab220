// exchange.cuh -- cross-GPU candidate exchange fused with the per-query merge (row-sharded corpus).
//
// The reference has no multi-device search; its scaling prose ("shard across instances, fan out,
// merge": /root/reference/docs/operations/scaling-and-performance.md:154-172) becomes: every rank
// finds the local top-k of its shard, the ranks exchange k candidates per query, every rank merges
// G*k candidates.  Instead of local-merge kernel -> NCCL all-gather -> merge kernel, ONE kernel does
// all three: the CTA that has merged query q's local candidates stores them straight into every
// peer's exchange buffer over NVLink (peer-mapped memory, plain st.global + a release flag per
// (source rank, query)), waits for the peers' flags for q in its OWN memory, and merges.
//
// Exchange buffer of one rank (cudaMalloc'd, exported with cudaIpcGetMemHandle, mapped by peers):
//   slots [2 parities][world source ranks][slot_stride bytes] : packed block of the source rank
//          ([ids int64 nq*k][scores f32 nq*k]), written by that rank
//   flags [2 parities][world][max_nq] u32                      : sequence number of the call
// Parity = seq & 1.  A rank can only start call seq+2 after it has seen every peer's flags of call
// seq+1, which a peer publishes after it finished reading the slots of call seq, so two slot sets
// suffice.  All ranks must issue the same sequence of sharded calls (same nq, k).
//
// FUSED = true needs all nq CTAs of every rank co-resident (a CTA pushes before it waits, but a
// not-yet-scheduled CTA cannot push): the host uses it for nq <= #SMs and otherwise launches the
// push (FUSED = false) and the wait+merge (exchange_wait_merge_kernel) as two kernels, which is
// deadlock-free for any nq because push kernels never wait.
#pragma once
#include "merge_topk.cuh"

namespace b2s {

struct ExchangeArgs {
    unsigned char* const* peer_base;   // device array [world]: exchange buffer base of every rank as mapped here
    unsigned char* local_base;         // == peer_base[rank]
    int world;
    int rank;
    long long slot_stride;             // bytes per (parity, source rank) slot
    long long flags_off;               // byte offset of the flag region
    int max_nq;                        // flags per (parity, source rank)
    unsigned seq;                      // sequence number of this call (>= 1)
    long long nq;                      // queries of the whole call
    int q_offset;                      // global index of this launch's query 0
    long long timeout_cycles;          // spin budget; on expiry *status = seq and the result is garbage
    unsigned* status;                  // mapped pinned host word, 0 = ok
    float* out_scores;                 // [nq, k] final
    long long* out_ids;
    // low-latency protocol (fused mode): candidates travel as three 8-byte words, each tagged with the
    // call's sequence number in its high half -- data IS the flag, so there is no fence and no second
    // message: one NVLink one-way trip instead of write + system fence + flag.
    int use_ll;
    long long ll_off;                  // byte offset of the LL region: [2 parities][world][ll_entries][3] u64
    long long ll_entries;              // candidates per (parity, source rank)
};

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Store this rank's sorted top-k of global query gq (keys in buf[0..kk)) into every rank's slot and
// publish the flag.  Block-wide.
template <int NT = 512>
__device__ __forceinline__ void exchange_push(const ExchangeArgs& ex, const MergeParams& p, const u64* buf, int kk,
                                              long long gq) {
    const int tid = threadIdx.x;
    const int parity = (int)(ex.seq & 1u);
    const long long slot = ((long long)parity * ex.world + ex.rank) * ex.slot_stride;
    const long long ids_off = slot + (gq * p.k) * 8;
    const long long sc_off = slot + ex.nq * p.k * 8 + (gq * p.k) * 4;
    for (int e = tid; e < p.k * ex.world; e += NT) {
        const int peer = e / p.k, i = e - peer * p.k;
        float s = -FLT_MAX;
        long long id = -1;
        if (i < kk) {
            const u64 key = buf[i];
            s = key_score(key);
            id = (long long)key_row(key) + p.id_offset;
        }
        unsigned char* base = ex.peer_base[peer];
        *reinterpret_cast<long long*>(base + ids_off + (long long)i * 8) = id;
        *reinterpret_cast<float*>(base + sc_off + (long long)i * 4) = s;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < ex.world) {
        unsigned* flag = reinterpret_cast<unsigned*>(ex.peer_base[tid] + ex.flags_off) +
                         ((long long)parity * ex.world + ex.rank) * ex.max_nq + gq;
        st_release_sys_u32(flag, ex.seq);
    }
}

// Wait for every rank's candidates of global query gq, merge world*k of them, write the final top-k.
template <int NT = 512>
__device__ __forceinline__ void exchange_wait_merge(const ExchangeArgs& ex, int k, u64* buf, long long gq) {
    const int tid = threadIdx.x;
    const int parity = (int)(ex.seq & 1u);
    if (tid < ex.world) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(ex.local_base + ex.flags_off) +
                               ((long long)parity * ex.world + tid) * ex.max_nq + gq;
        const long long t0 = clock64();
        while (ld_acquire_sys_u32(flag) != ex.seq) {
            if (clock64() - t0 > ex.timeout_cycles) {
                *(volatile unsigned*)ex.status = ex.seq;   // mapped host word
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const int total = ex.world * k;
    auto slot_of = [&](int r) { return ex.local_base + ((long long)parity * ex.world + r) * ex.slot_stride; };
    if (total > kMergeSortCap) {
        // large k on many ranks (e.g. 8 x 1000): every rank's block is sorted, so a candidate's global rank is
        // its own position plus, per other rank, a binary search -- no shared-memory staging at all
        merge_runs_bsearch<NT>(
            ex.world, k, k,
            [&](int r, int i) { return __ldcg(reinterpret_cast<const float*>(slot_of(r) + ex.nq * k * 8 + (gq * k + i) * 4)); },
            [&](int r, int i) { return __ldcg(reinterpret_cast<const long long*>(slot_of(r) + (gq * k + i) * 8)); },
            ex.out_scores + gq * k, ex.out_ids + gq * k);
        return;
    }
    for (int e = tid; e < total; e += NT) {
        const int r = e / k, i = e - r * k;
        const unsigned char* slot = slot_of(r);
        const long long id = __ldcg(reinterpret_cast<const long long*>(slot + (gq * k + i) * 8));
        const float s = __ldcg(reinterpret_cast<const float*>(slot + ex.nq * k * 8 + (gq * k + i) * 4));
        // position e = rank-major, then local order: keeps (score desc, id asc) across shards
        buf[e] = id < 0 ? 0ull : make_key(s, (uint32_t)e);
    }
    __syncthreads();
    block_sort_desc<NT>(buf, total > 0 ? total : 1, tid);
    for (int i = tid; i < k; i += NT) {
        float s = -FLT_MAX;
        long long id = -1;
        if (i < total && buf[i] != 0ull) {
            const int pos = (int)key_row(buf[i]);
            const int r = pos / k, j = pos - r * k;
            const unsigned char* slot = slot_of(r);
            id = __ldcg(reinterpret_cast<const long long*>(slot + (gq * k + j) * 8));
            s = __ldcg(reinterpret_cast<const float*>(slot + ex.nq * k * 8 + (gq * k + j) * 4));
        }
        ex.out_scores[gq * k + i] = s;
        ex.out_ids[gq * k + i] = id;
    }
}

__device__ __forceinline__ void st_relaxed_sys_u64(u64* p, u64 v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_sys_u64(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// LL push: entry = {seq|score bits, seq|id low, seq|id high}; 8-byte stores are single-copy atomic, so a
// reader that sees the tag sees the payload of that word.
template <int NT = 512>
__device__ __forceinline__ void exchange_push_ll(const ExchangeArgs& ex, const MergeParams& p, const u64* buf, int kk,
                                                 long long gq) {
    const int tid = threadIdx.x;
    const int parity = (int)(ex.seq & 1u);
    const long long entry0 = ((long long)parity * ex.world + ex.rank) * ex.ll_entries + gq * p.k;
    const u64 tag = (u64)ex.seq << 32;
    for (int e = tid; e < p.k * ex.world; e += NT) {
        const int peer = e / p.k, i = e - peer * p.k;
        float s = -FLT_MAX;
        long long id = -1;
        if (i < kk) {
            const u64 key = buf[i];
            s = key_score(key);
            id = (long long)key_row(key) + p.id_offset;
        }
        u64* dst = reinterpret_cast<u64*>(ex.peer_base[peer] + ex.ll_off) + (entry0 + i) * 3;
        st_relaxed_sys_u64(dst, tag | (u64)__float_as_uint(s));
        st_relaxed_sys_u64(dst + 1, tag | (u64)(uint32_t)id);
        st_relaxed_sys_u64(dst + 2, tag | (u64)(uint32_t)((unsigned long long)id >> 32));
    }
}

// LL wait + merge: every thread polls its own candidate's three words in local memory (ONE deadline for
// the whole call: a missing peer costs timeout_cycles once, not once per candidate).
constexpr int kExchangeRankSortMax = 1024;
template <int NT = 512>
__device__ __forceinline__ void exchange_wait_merge_ll(const ExchangeArgs& ex, int k, u64* buf, long long gq,
                                                       unsigned long long* tr = nullptr) {
    __shared__ int s_valid;
    const int tid = threadIdx.x;
    const int parity = (int)(ex.seq & 1u);
    const int total = ex.world * k;   // host guarantees total <= kMergeSortCap
    const u64* ll = reinterpret_cast<const u64*>(ex.local_base + ex.ll_off);
    if (tid == 0) s_valid = 0;
    const long long t0 = clock64();
    float s_first = -FLT_MAX;         // this thread's first candidate stays in registers (total <= NT: the only one)
    long long id_first = -1;
    for (int e = tid; e < total; e += NT) {
        const int r = e / k, i = e - r * k;
        const u64* src = ll + (((long long)parity * ex.world + r) * ex.ll_entries + gq * k + i) * 3;
        u64 w0, w1, w2;
        while (true) {
            w0 = ld_relaxed_sys_u64(src);
            w1 = ld_relaxed_sys_u64(src + 1);
            w2 = ld_relaxed_sys_u64(src + 2);
            if ((uint32_t)(w0 >> 32) == ex.seq && (uint32_t)(w1 >> 32) == ex.seq && (uint32_t)(w2 >> 32) == ex.seq) break;
            if (clock64() - t0 > ex.timeout_cycles) {
                *(volatile unsigned*)ex.status = ex.seq;   // mapped host word
                w1 = w2 = 0xffffffffull;   // id -1: ignored
                break;
            }
        }
        const long long id = (long long)(((u64)(uint32_t)w2 << 32) | (u64)(uint32_t)w1);
        const float sc = __uint_as_float((uint32_t)w0);
        buf[e] = id < 0 ? 0ull : make_key(sc, (uint32_t)e);
        if (e == tid) {
            s_first = sc;
            id_first = id;
        }
    }
    __syncthreads();
    if (tr) tr[4] = globaltimer_ns();
    if (total <= kExchangeRankSortMax) {
        // rank sort: keys are unique (position is the tie-break), a valid key's rank is its output slot
        int mine_valid = 0;
        for (int e = tid; e < total; e += NT) {
            const u64 key = buf[e];
            if (key == 0ull) continue;
            ++mine_valid;
            int rank = 0;
            for (int j = 0; j < total; ++j) rank += buf[j] > key;
            if (rank < k) {
                float sc = s_first;
                long long id = id_first;
                if (e != tid) {
                    const int r = e / k, i = e - r * k;
                    const u64* src = ll + (((long long)parity * ex.world + r) * ex.ll_entries + gq * k + i) * 3;
                    sc = __uint_as_float((uint32_t)__ldcg(src));
                    id = (long long)(((u64)(uint32_t)__ldcg(src + 2) << 32) | (u64)(uint32_t)__ldcg(src + 1));
                }
                ex.out_scores[gq * k + rank] = sc;
                ex.out_ids[gq * k + rank] = id;
            }
        }
        if (mine_valid) atomicAdd(&s_valid, mine_valid);
        __syncthreads();
        for (int i = s_valid + tid; i < k; i += NT) {
            ex.out_scores[gq * k + i] = -FLT_MAX;
            ex.out_ids[gq * k + i] = -1;
        }
        return;
    }
    block_sort_desc<NT>(buf, total > 0 ? total : 1, tid);
    for (int i = tid; i < k; i += NT) {
        float s = -FLT_MAX;
        long long id = -1;
        if (i < total && buf[i] != 0ull) {
            const int pos = (int)key_row(buf[i]);
            const int r = pos / k, j = pos - r * k;
            const u64* src = ll + (((long long)parity * ex.world + r) * ex.ll_entries + gq * k + j) * 3;
            s = __uint_as_float((uint32_t)__ldcg(src));
            id = (long long)(((u64)(uint32_t)__ldcg(src + 2) << 32) | (u64)(uint32_t)__ldcg(src + 1));
        }
        ex.out_scores[gq * k + i] = s;
        ex.out_ids[gq * k + i] = id;
    }
}

// push + wait + merge of one query inside a block that holds its sorted local top-k in buf
template <int NT = 512>
__device__ __forceinline__ void exchange_fused(const ExchangeArgs& ex, const MergeParams& p, u64* buf, int kk, long long gq,
                                               unsigned long long* tr = nullptr) {
    if (ex.use_ll) {
        exchange_push_ll<NT>(ex, p, buf, kk, gq);
        __syncthreads();
        if (tr) tr[3] = globaltimer_ns();
        exchange_wait_merge_ll<NT>(ex, p.k, buf, gq, tr);
    } else {
        exchange_push<NT>(ex, p, buf, kk, gq);
        __syncthreads();
        exchange_wait_merge<NT>(ex, p.k, buf, gq);
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(kMergeThreads) merge_exchange_kernel(const MergeParams p, const ExchangeArgs ex) {
    __shared__ MergeSmem sm;
    const int q = blockIdx.x;
    const long long gq = (long long)ex.q_offset + q;
    grid_dep_launch();
    grid_dep_wait();
    const int m_sorted = merge_lists_sorted(p, q, sm);
    const int kk = m_sorted < p.k ? m_sorted : p.k;
    if (FUSED) exchange_fused(ex, p, sm.buf, kk, gq);
    else exchange_push(ex, p, sm.buf, kk, gq);
}

__global__ void __launch_bounds__(kMergeThreads) exchange_wait_merge_kernel(const ExchangeArgs ex, int k) {
    __shared__ u64 buf[kMergeSortCap];
    exchange_wait_merge(ex, k, buf, (long long)ex.q_offset + blockIdx.x);
}

}  // namespace b2s

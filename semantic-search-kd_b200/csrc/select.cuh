// select.cuh -- candidate keys, threshold-filtered candidate lists and bitonic sorting.
//
// Every score that can still belong to a query's top-k is packed into ONE sortable 64-bit key:
//      key = (order_preserving_u32(score) << 32) | ~local_row_id
// so "larger key" == "higher score, then LOWER id" -- the (score desc, id asc) order of the
// oracle (oracle/flat_ip.c: hit_better) and the earlier-id-survives rule of faiss' strict '>'
// heap replacement.  All selection below is integer work on these keys and is bit-exact.
//
// A candidate list is a small array of keys owned by one CTA for one query, guarded by a spin
// lock in shared memory.  Rows whose score reaches the list's threshold are appended by the
// whole warp (ballot-compacted); when the list is full one warp sorts it, keeps the best k and
// raises the threshold to the k-th key.  On unit-norm embedding data the survivor rate after
// warm-up is ~k/rows_seen, so this path is cold; it is written for correctness on adversarial
// inputs (sorted scores, all-equal scores), not for speed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2s {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t score_to_ord(float s) {
    s = s + 0.0f;  // -0.0 -> +0.0 so that equal scores have equal keys
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_score(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ u64 make_key(float s, uint32_t row) {
    return ((u64)score_to_ord(s) << 32) | (u64)(~row);
}
__device__ __forceinline__ float key_score(u64 key) { return ord_to_score((uint32_t)(key >> 32)); }
__device__ __forceinline__ uint32_t key_row(u64 key) { return ~(uint32_t)key; }

// Smallest power of two >= max(2k, k + 32), at least 64: room for one full warp of appends after
// a compaction, and at least a doubling of rows seen between compactions.  Small lists keep the
// compaction (a single-warp bitonic sort of `cap` keys) cheap and the thresholds fresh.
__host__ __device__ inline int list_capacity(int k) {
    int need = 2 * k > k + 32 ? 2 * k : k + 32;
    int c = 64;
    while (c < need) c <<= 1;
    return c;
}

// ---------------------------------------------------------------------------------------------
// Lane-uniform building blocks.  Everything below is written so that NO branch depends on the
// lane id or on per-lane data: per-lane effects are predicated instructions or selects, loop trip
// counts are warp-uniform.  A warp that never diverges never needs to reconverge -- which matters
// twice: the hot loops stay on the converged fast path of every warp collective, and the tensor
// path's tcgen05.ld.sync.aligned is only defined for a converged warp.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// if (pred) *ptr = v;   (generic address, no branch)
__device__ __forceinline__ void st_u64_if(u64* ptr, u64 v, bool pred) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p st.u64 [%0], %1;\n\t"
        "}\n"
        ::"l"(ptr), "l"(v), "r"((uint32_t)pred)
        : "memory");
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; it must execute
// grid_dep_wait() before touching anything the predecessor writes (or, for buffers the predecessor
// READS, before overwriting them).  grid_dep_launch() lets the successor start early.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// L2 load pinned in program order (volatile asm): issued where it is written, so that a value wanted one loop
// iteration later is in flight for a whole iteration instead of being sunk to its use by the compiler
__device__ __forceinline__ unsigned ldcg_pinned_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// nanoseconds of the GPU-wide timer (comparable across SMs; ~32 ns granularity): option "trace" stamps
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// In-place bitonic sort, DESCENDING, of a[0..n) (n a power of two, n/2 a multiple of nthreads)
// by `nthreads` cooperating threads with ids tid in [0, nthreads).  `sync` is __syncwarp or
// __syncthreads.  Compare-exchange is select + unconditional stores.
template <typename SyncFn>
__device__ __forceinline__ void bitonic_sort_desc(u64* a, int n, int tid, int nthreads, SyncFn sync) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                const int lo = 2 * i - (i & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const u64 x = a[lo], y = a[hi];
                const bool sw = (x < y) == desc;
                a[lo] = sw ? y : x;
                a[hi] = sw ? x : y;
            }
            sync();
        }
    }
}

struct WarpSync {
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// Per-(CTA, query) list state, kept in shared memory.  Held as separate arrays by the kernels
// (the tensor-core epilogue reads thresholds as broadcast vectors), referenced through ListRef.
struct ListRef {
    u64* thr_key;   // a candidate must have key > *thr_key
    float* thr;     // fast test: score >= *thr  (score of thr_key, -inf when unset)
    int* count;     // entries currently in the list
    int* lock;      // 0 free, 1 held (shared memory)
};

// seed_key (optional) is the k-th best key of a row SAMPLE of the same shard: every final top-k
// key is >= it, so the threshold is made inclusive (seed - 1) -- the sampled row itself is
// scanned again by the main pass and must still be accepted.
__device__ __forceinline__ void list_init(const ListRef& st, u64 seed_key) {
    *st.thr_key = seed_key ? seed_key - 1ull : 0ull;
    *st.thr = seed_key ? key_score(seed_key) : -INFINITY;
    *st.count = 0;
    *st.lock = 0;
}
__device__ __forceinline__ void list_disable(const ListRef& st) {  // masked query slot
    *st.thr_key = ~0ull;
    *st.thr = INFINITY;
    *st.count = 0;
    *st.lock = 0;
}

// Warp-collective spin lock in shared memory: lane 0 issues the CAS under a predicate, the result
// is broadcast, the retry loop is warp-uniform.
__device__ __forceinline__ void spin_acquire(int* lock, int lane) {
    const uint32_t addr = smem_addr_u32(lock);
    uint32_t old;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.eq.u32 p, %2, 0;\n\t"
            "mov.u32 %0, 1;\n\t"
            "@p atom.shared.cas.b32 %0, [%1], 0, 1;\n\t"
            "}\n"
            : "=r"(old)
            : "r"(addr), "r"((uint32_t)lane)
            : "memory");
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old != 0) __nanosleep(32);
    } while (old != 0);
    __threadfence_block();
}
__device__ __forceinline__ void spin_release(int* lock, int lane) {
    (void)lane;
    __threadfence_block();
    __syncwarp();
    *(volatile int*)lock = 0;   // every lane stores the same value: no lane predicate, no branch
    __syncwarp();
}

// One warp keeps the best k of entries[0..count) and publishes the new threshold.  The sort runs
// in `work` (shared memory, cap slots): for a shared-memory list work == entries; for a
// global-memory list the keys are staged through the CTA's scratch buffer (scratch_lock).
// cap is a power of two >= 64.  Caller holds the list lock.  Returns the new count.
__device__ __forceinline__ int list_compact_warp(const ListRef& st, u64* entries, int cap, int k, int lane,
                                                 u64* scratch = nullptr, int* scratch_lock = nullptr) {
    const int c = *(volatile int*)st.count;
    u64* work = entries;
    if (scratch != nullptr) {   // uniform
        spin_acquire(scratch_lock, lane);
        work = scratch;
    }
    for (int i = lane; i < cap; i += 32) {
        const u64 v = entries[i];          // slots >= c hold stale keys: masked to 0
        work[i] = (i < c) ? v : 0ull;
    }
    __syncwarp();
    bitonic_sort_desc(work, cap, lane, 32, WarpSync());
    const int keep = c < k ? c : k;
    if (scratch != nullptr) {
        for (int i = lane; i < cap; i += 32) st_u64_if(entries + i, work[i], i < keep);
    }
    __syncwarp();
    if (c >= k) {   // uniform; every lane stores the same values
        const u64 kth = work[k - 1];
        if (kth > *(volatile u64*)st.thr_key) {   // never below a bound adopted from outside (scan kernel: the global slots)
            *(volatile u64*)st.thr_key = kth;
            *(volatile float*)st.thr = key_score(kth);
        }
    }
    *(volatile int*)st.count = keep;
    __syncwarp();
    if (scratch != nullptr) spin_release(scratch_lock, lane);
    return keep;
}

// Warp-collective append.  Every lane of the warp calls it (converged); `pass` says whether this
// lane offers `key`.  entries has `cap` slots (shared memory, or global memory with a scratch).
__device__ __forceinline__ void list_append_warp(const ListRef& st, u64* entries, int cap, int k, bool pass,
                                                 u64 key, int lane, u64* scratch = nullptr,
                                                 int* scratch_lock = nullptr) {
    spin_acquire(st.lock, lane);
    u64 tk = *(volatile u64*)st.thr_key;
    pass = pass && (key > tk);
    unsigned m = __ballot_sync(0xffffffffu, pass);
    int n = __popc(m);
    if (n) {   // uniform
        int c = *(volatile int*)st.count;
        if (c + n > cap) {   // uniform
            c = list_compact_warp(st, entries, cap, k, lane, scratch, scratch_lock);
            tk = *(volatile u64*)st.thr_key;
            pass = pass && (key > tk);
            m = __ballot_sync(0xffffffffu, pass);
            n = __popc(m);
        }
        const int slot = c + __popc(m & ((1u << lane) - 1u));
        st_u64_if(entries + (pass ? slot : 0), key, pass);
        __syncwarp();
        *(volatile int*)st.count = c + n;
    }
    spin_release(st.lock, lane);
}

}  // namespace b2s

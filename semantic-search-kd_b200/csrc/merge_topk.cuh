// merge_topk.cuh -- K3/K4: per-query selection of the final top-k from candidate keys.
//
// Replaces faiss' heap_reorder / the reference's `np.argsort(scores)[::-1][:k]`
// (/root/reference/src/kd/eval.py:86, scripts/simple_eval.py:35) on the (tiny) candidate set
// that survives the fused per-CTA filters, and -- as K4 -- the cross-GPU merge of G sorted
// local top-k lists after the all-gather.
//
// One CTA per query.  Candidates arrive as L segments of 64-bit keys (select.cuh).  If they fit
// the shared-memory sort buffer they are gathered and bitonic-sorted; otherwise an MSB-first
// 11-bit radix select over the segments narrows them to (winners + one bucket) that fits, then
// the same sort finishes.  Keys are unique (row ids are unique), so the result is exact and
// deterministic: scores descending, ties by ascending id, unfilled slots (-FLT_MAX, -1).
#pragma once
#include <float.h>
#include "select.cuh"

namespace b2s {

constexpr int kMergeThreads = 512;
constexpr int kMergeSortCap = 4096;  // keys (32 KB of shared memory)
constexpr int kRadixBits = 10;
constexpr int kRadixBins = 1 << kRadixBits;

struct MergeParams {
    const u64* lists;    // [L, nq_lists, cap]
    const int* counts;   // [L, nq_lists]
    int num_lists;       // L
    int nq_lists;        // list stride in queries
    int lists_sorted;    // 1: every list is sorted descending (scan path) -> cheap merge without a gather
    int cap;
    int k;
    long long id_offset;  // added to decoded row ids
    float* out_scores;    // [nq, k] or nullptr
    long long* out_ids;   // [nq, k] or nullptr
    u64* out_kth_key;     // optional [nq]: k-th best key (0 if fewer than k candidates) -- seeding
    // K4 mode: candidates are (score, id) pairs instead of keys
    const float* in_scores;    // [G, nq, k_in] or nullptr
    const long long* in_ids;   // [G, nq, k_in]
    int g;
    int k_in;
    long long nq;
    long long g_stride_scores;  // elements between consecutive ranks' score blocks
    long long g_stride_ids;     // elements between consecutive ranks' id blocks
};

__device__ __forceinline__ int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

// Block-wide: sort buf[0..m) descending (m <= kMergeSortCap), padding with 0 keys.
template <int NT = 512>
__device__ __forceinline__ void block_sort_desc(u64* buf, int m, int tid) {
    const int n = next_pow2(m);
    for (int i = m + tid; i < n; i += NT) buf[i] = 0ull;
    __syncthreads();
    bitonic_sort_desc(buf, n, tid, NT, BlockSync());
}

// Merge G runs of k_in (score, id) candidates each, every run sorted by (score desc, id asc) and run r
// holding smaller ids than run r+1 (rank order): the global position of candidate (r, i) is i plus, for
// every other run, the number of its entries that precede (r, i) -- one binary search per run, straight
// from global memory, so G * k_in is not bounded by the shared-memory sort buffer (8 ranks x k = 2048).
// Padding entries (-FLT_MAX, -1) sort last and are copied through like any other entry.
template <int NT, typename ScoreAt, typename IdAt>
__device__ __forceinline__ void merge_runs_bsearch(int G, int k_in, int k_out, ScoreAt score_at, IdAt id_at,
                                                   float* out_scores, long long* out_ids) {
    const int total = G * k_in;
    for (int e = threadIdx.x; e < total; e += NT) {
        const int r = e / k_in, i = e - r * k_in;
        const float s = score_at(r, i);
        int rank = i;
        for (int r2 = 0; r2 < G && rank < k_out; ++r2) {
            if (r2 == r) continue;
            int lo = 0, hi = k_in;   // first entry of run r2 that does NOT precede (s, r)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float s2 = score_at(r2, mid);
                const bool prec = (s2 > s) || (s2 == s && r2 < r);
                if (prec) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k_out) {
            out_scores[rank] = s;
            out_ids[rank] = id_at(r, i);
        }
    }
}

constexpr int kMergeMaxLists = 512;   // candidate lists per query (CTAs of the producing kernel)
constexpr int kMergeFastCap = 512;    // survivors the rank-sort fast path can hold

// Position e of the concatenation of all lists -> its key.  offs[] = exclusive prefix sums of the
// list counts (shared memory); a binary search finds the list, so every thread's load is
// independent of every other (one DRAM/L2 round trip for the whole gather instead of one per list).
__device__ __forceinline__ u64 fetch_candidate(const MergeParams& p, const int* offs, int L, int q, int e,
                                               int* list_out = nullptr) {
    int lo = 0, hi = L;  // largest l with offs[l] <= e
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (offs[mid] <= e) lo = mid;
        else hi = mid;
    }
    if (list_out) *list_out = lo;
    return __ldcg(p.lists + ((size_t)lo * p.nq_lists + q) * p.cap + (e - offs[lo]));
}

struct MergeSmem {
    u64 buf[kMergeSortCap];
    int offs[kMergeMaxLists + 1];
    __align__(16) int hist[kRadixBins];
    int s_fill;
    u64 s_prefix;      // selected high bits so far
    int s_bits_done;   // number of high bits fixed in s_prefix
    int s_k_rem;       // rank still to find inside the current bucket
    int s_bucket_cnt;  // candidates inside the current bucket
    u64 sel[kMergeFastCap];
    u64 s_floor;
    int s_nonempty;
};

// Block-wide: leave the best min(m, k..) candidate keys of query q sorted descending in sm.buf and
// return how many of them are valid (>= min(k, total candidates)).
template <int NT = 512>
__device__ __forceinline__ int merge_lists_sorted(const MergeParams& p, const int q, MergeSmem& sm) {
    u64* buf = sm.buf;
    int* offs = sm.offs;
    int* hist = sm.hist;
    int& s_fill = sm.s_fill;
    u64& s_prefix = sm.s_prefix;
    int& s_bits_done = sm.s_bits_done;
    int& s_k_rem = sm.s_k_rem;
    int& s_bucket_cnt = sm.s_bucket_cnt;
    u64* sel = sm.sel;
    u64& s_floor = sm.s_floor;
    int& s_nonempty = sm.s_nonempty;

    const int tid = threadIdx.x;
    const int L = p.num_lists;    // host guarantees L <= kMergeMaxLists

    if (p.lists_sorted) {
        // ---- sorted lists (K1): the k-th largest list HEAD bounds the answer from below, and a sorted
        // list can be cut at the first key under that floor: two dependent memory round trips in total,
        // no prefix sum, no gather of every candidate.
        u64* heads = reinterpret_cast<u64*>(hist);
        // the first four keys of every list come in with the counts, in ONE memory round trip: a list rarely
        // has more than a few keys above the floor, so the walk below seldom touches memory again
        constexpr int R = (kMergeMaxLists + NT - 1) / NT;
        u64 pk[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int l = tid + r * NT;
            int c = 0;
            ulonglong2 a = make_ulonglong2(0ull, 0ull), b = a;
            if (l < L) {
                c = __ldcg(p.counts + (size_t)l * p.nq_lists + q);
                const ulonglong2* lp2 = reinterpret_cast<const ulonglong2*>(p.lists + ((size_t)l * p.nq_lists + q) * p.cap);
                a = __ldcg(lp2);       // (garbage beyond the list's count, masked below; cap >= 64 keys)
                b = __ldcg(lp2 + 1);
            }
            pk[r][0] = c > 0 ? a.x : 0ull;
            pk[r][1] = c > 1 ? a.y : 0ull;
            pk[r][2] = c > 2 ? b.x : 0ull;
            pk[r][3] = c > 3 ? b.y : 0ull;
            if (l < kMergeMaxLists) {
                offs[l] = c;
                heads[l] = pk[r][0];
            }
        }
        if (tid == 0) {
            s_floor = 0ull;
            s_nonempty = 0;
            s_fill = 0;
        }
        __syncthreads();
        {
            int mine = 0;
            for (int l = tid; l < L; l += NT) mine += heads[l] != 0ull;
            if (mine) atomicAdd(&s_nonempty, mine);
        }
        __syncthreads();
        if (s_nonempty >= p.k) {
            for (int l = tid; l < L; l += NT) {
                const u64 h = heads[l];
                int rank = 0;
                for (int j = 0; j < L; ++j) rank += heads[j] > h;
                if (rank == p.k - 1 && h != 0ull) s_floor = h;   // keys are unique: exactly one writer
            }
        }
        __syncthreads();
        const u64 floor_key = s_floor;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int l = tid + r * NT;
            if (l >= L) continue;
            const int c = offs[l];
            const u64* lp = p.lists + ((size_t)l * p.nq_lists + q) * p.cap;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < c && pk[r][i] >= floor_key && (i == 0 || pk[r][i - 1] >= floor_key)) {
                    const int pos = atomicAdd(&s_fill, 1);
                    if (pos < kMergeFastCap) sel[pos] = pk[r][i];
                }
            }
            if (c > 4 && pk[r][3] >= floor_key) {   // rare: more than four keys of one list above the floor
                for (int i = 4; i < c; ++i) {
                    const u64 key = __ldcg(lp + i);
                    if (key < floor_key) break;
                    const int pos = atomicAdd(&s_fill, 1);
                    if (pos < kMergeFastCap) sel[pos] = key;
                }
            }
        }
        __syncthreads();
        const int C = s_fill;
        if (C <= kMergeFastCap) {
            for (int t = tid; t < C; t += NT) {
                const u64 key = sel[t];
                int rank = 0;
                for (int j = 0; j < C; ++j) rank += sel[j] > key;
                buf[rank] = key;
            }
            __syncthreads();
            return C;
        }
        __syncthreads();   // too many survivors (many ties / tiny k-th bound): the general path below
    }

    // exclusive prefix sums of the per-list counts (Hillis-Steele in shared memory)
    for (int l = tid; l < L; l += NT) offs[l + 1] = __ldcg(p.counts + (size_t)l * p.nq_lists + q);
    if (tid == 0) {
        offs[0] = 0;
        s_fill = 0;
    }
    __syncthreads();
    for (int d = 1; d < L; d <<= 1) {
        int add[(kMergeMaxLists + NT - 1) / NT];
#pragma unroll
        for (int r = 0; r < (kMergeMaxLists + NT - 1) / NT; ++r) {
            const int l = tid + r * NT + 1;
            add[r] = (l <= L && l - d >= 1) ? offs[l - d] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < (kMergeMaxLists + NT - 1) / NT; ++r) {
            const int l = tid + r * NT + 1;
            if (l <= L) offs[l] += add[r];
        }
        __syncthreads();
    }
    const int M = offs[L];
    int m_sorted;  // number of valid keys in buf[0..) sorted descending at the end
    bool sorted_done = false;

    // ---- fast path ---------------------------------------------------------------------------
    // The k-th largest of the per-list maxima ("heads") is a lower bound of the k-th largest key
    // overall (the k largest heads are k distinct candidates), so only keys >= that floor can be
    // in the answer -- typically ~2k of the M candidates.  They are compacted and rank-sorted
    // (each thread counts the keys greater than its own): no 4096-wide bitonic network.
    u64* heads = reinterpret_cast<u64*>(hist);   // [L] aliases the radix histogram (used later only)
    for (int l = tid; l < L; l += NT) heads[l] = 0ull;
    if (tid == 0) {
        s_floor = 0ull;
        s_nonempty = 0;
    }
    __syncthreads();
    for (int e = tid; e < M; e += NT) {
        int l;
        const u64 key = fetch_candidate(p, offs, L, q, e, &l);
        atomicMax(&heads[l], key);
        if (M <= kMergeSortCap) buf[e] = key;
    }
    __syncthreads();
    {
        int mine = 0;
        for (int l = tid; l < L; l += NT) mine += heads[l] != 0ull;
        if (mine) atomicAdd(&s_nonempty, mine);
    }
    __syncthreads();
    if (s_nonempty >= p.k) {
        for (int l = tid; l < L; l += NT) {
            const u64 h = heads[l];
            int rank = 0;
            for (int j = 0; j < L; ++j) rank += heads[j] > h;
            if (rank == p.k - 1 && h != 0ull) s_floor = h;   // keys are unique: exactly one writer
        }
    }
    __syncthreads();
    const u64 floor_key = s_floor;
    for (int e = tid; e < M; e += NT) {
        const u64 key = (M <= kMergeSortCap) ? buf[e] : fetch_candidate(p, offs, L, q, e);
        if (key >= floor_key) {
            const int pos = atomicAdd(&s_fill, 1);
            if (pos < kMergeFastCap) sel[pos] = key;
        }
    }
    __syncthreads();
    const int C = s_fill;
    if (C <= kMergeFastCap) {
        __syncthreads();   // everyone has read s_fill / buf before buf is overwritten
        for (int t = tid; t < C; t += NT) {
            const u64 key = sel[t];
            int rank = 0;
            for (int j = 0; j < C; ++j) rank += sel[j] > key;
            buf[rank] = key;
        }
        __syncthreads();
        m_sorted = C;
        sorted_done = true;
    } else {
        // Too many survivors for the rank sort: MSB-first radix select of the k-th largest key, then
        // only the winners and the bucket that holds the k-th are sorted (~k keys, not all M).
        const bool in_smem = M <= kMergeSortCap;   // buf[0..M) holds every candidate
        int stop_cap = 2 * next_pow2(p.k);
        stop_cap = stop_cap < 512 ? 512 : (stop_cap > kMergeSortCap ? kMergeSortCap : stop_cap);
        if (in_smem && M <= stop_cap) {
            m_sorted = M;
        } else {
            if (tid == 0) {
                s_fill = 0;
                s_prefix = 0ull;
                s_bits_done = 0;
                s_k_rem = p.k;
                s_bucket_cnt = M;
            }
            __syncthreads();
            while (true) {
                const int bits_done = s_bits_done;
                const int k_rem = s_k_rem;
                // stop when winners (k - k_rem) + bucket are few enough to sort, or all bits are fixed
                if ((p.k - k_rem) + s_bucket_cnt <= stop_cap || bits_done >= 64) break;
                const int nbits = (64 - bits_done) < kRadixBits ? (64 - bits_done) : kRadixBits;
                const int shift = 64 - bits_done - nbits;
                const u64 prefix = s_prefix;
                for (int i = tid; i < kRadixBins; i += NT) hist[i] = 0;
                __syncthreads();
                for (int e = tid; e < M; e += NT) {
                    const u64 key = in_smem ? buf[e] : fetch_candidate(p, offs, L, q, e);
                    const bool in_bucket = bits_done == 0 || (key >> (64 - bits_done)) == prefix;
                    if (in_bucket) atomicAdd(&hist[(int)((key >> shift) & ((1u << nbits) - 1u))], 1);
                }
                __syncthreads();
                if (tid < 32) {
                    // warp 0: lane l owns bins [32 l, 32 l + 32); suffix sums over lanes find the lane
                    // in which the count from the top reaches k_rem, that lane walks its 32 bins
                    int mine = 0;
                    for (int b = 0; b < 32; ++b) mine += hist[tid * 32 + b];
                    int above = 0;   // candidates in bins of higher lanes
                    for (int l = 31; l >= 0; --l) {
                        const int v = __shfl_sync(0xffffffffu, mine, l);
                        if (l > tid) above += v;
                    }
                    const bool here = above < k_rem && above + mine >= k_rem;   // exactly one lane
                    if (here) {
                        int acc = above;
                        int b = tid * 32 + 31;
                        for (; b > tid * 32; --b) {
                            if (acc + hist[b] >= k_rem) break;
                            acc += hist[b];
                        }
                        // bucket b holds the k_rem-th largest of the current bucket
                        s_prefix = (prefix << nbits) | (u64)b;
                        s_bits_done = bits_done + nbits;
                        s_k_rem = k_rem - acc;
                        s_bucket_cnt = hist[b];
                    }
                }
                __syncthreads();
            }
            // gather winners (prefix bits above the bucket) and the bucket itself
            const int bits_done = s_bits_done;
            const u64 prefix = s_prefix;
            for (int e = tid; e < M; e += NT) {
                const u64 key = fetch_candidate(p, offs, L, q, e);
                const bool take = bits_done == 0 || (key >> (64 - bits_done)) >= prefix;
                if (take) {
                    const int pos = atomicAdd(&s_fill, 1);
                    if (pos < kMergeSortCap) buf[pos] = key;
                }
            }
            __syncthreads();
            m_sorted = s_fill < kMergeSortCap ? s_fill : kMergeSortCap;
        }
    }

    if (!sorted_done) block_sort_desc<NT>(buf, m_sorted > 0 ? m_sorted : 1, tid);
    return m_sorted;
}

__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const MergeParams p) {
    __shared__ MergeSmem sm;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    grid_dep_launch();   // PDL: the next kernel of the stream may start; it waits for us where it must
    grid_dep_wait();     // the producer of the candidate lists has completed
    const int m_sorted = merge_lists_sorted(p, q, sm);
    const u64* buf = sm.buf;

    const int kk = m_sorted < p.k ? m_sorted : p.k;
    for (int i = tid; i < p.k; i += kMergeThreads) {
        float s = -FLT_MAX;
        long long id = -1;
        if (i < kk) {
            const u64 key = buf[i];
            s = key_score(key);
            id = (long long)key_row(key) + p.id_offset;
        }
        if (p.out_scores) p.out_scores[(size_t)q * p.k + i] = s;
        if (p.out_ids) p.out_ids[(size_t)q * p.k + i] = id;
    }
    if (p.out_kth_key && tid == 0) p.out_kth_key[q] = (m_sorted >= p.k) ? buf[p.k - 1] : 0ull;
}

// K4: merge G sorted (score, id) lists of length k_in per query.  Position g*k_in + j is the
// tie-break (lists are in ascending id-range order and internally ordered by ascending id among
// equal scores), so the packed key keeps the global (score desc, id asc) order without 64-bit ids.
__global__ void __launch_bounds__(kMergeThreads) merge_pairs_kernel(const MergeParams p) {
    __shared__ u64 buf[kMergeSortCap];
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = p.g * p.k_in;
    if (total > kMergeSortCap) {
        const size_t qoff = (size_t)q * p.k_in;
        merge_runs_bsearch<kMergeThreads>(
            p.g, p.k_in, p.k, [&](int g, int j) { return p.in_scores[(size_t)g * p.g_stride_scores + qoff + j]; },
            [&](int g, int j) { return p.in_ids[(size_t)g * p.g_stride_ids + qoff + j]; },
            p.out_scores + (size_t)q * p.k, p.out_ids + (size_t)q * p.k);
        return;
    }
    for (int i = tid; i < total; i += kMergeThreads) {
        const int g = i / p.k_in, j = i - g * p.k_in;
        const size_t off = (size_t)q * p.k_in + j;
        const long long id = p.in_ids[(size_t)g * p.g_stride_ids + off];
        buf[i] = id < 0 ? 0ull : make_key(p.in_scores[(size_t)g * p.g_stride_scores + off], (uint32_t)i);
    }
    __syncthreads();
    block_sort_desc(buf, total > 0 ? total : 1, tid);
    for (int i = tid; i < p.k; i += kMergeThreads) {
        float s = -FLT_MAX;
        long long id = -1;
        if (i < total && buf[i] != 0ull) {
            const int pos = (int)key_row(buf[i]);
            const int g = pos / p.k_in, j = pos - g * p.k_in;
            const size_t off = (size_t)q * p.k_in + j;
            s = p.in_scores[(size_t)g * p.g_stride_scores + off];
            id = p.in_ids[(size_t)g * p.g_stride_ids + off];
        }
        p.out_scores[(size_t)q * p.k + i] = s;
        p.out_ids[(size_t)q * p.k + i] = id;
    }
}

}  // namespace b2s

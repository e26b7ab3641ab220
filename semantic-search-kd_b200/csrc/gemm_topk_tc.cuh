// gemm_topk_tc.cuh -- K2: batched exact inner-product search on the 5th-gen tensor cores with a
// fused top-k epilogue (the [nq x N] score matrix never exists outside TMEM).
//
// Replaces the reference's dense contraction + per-row sort for query batches:
//   np.matmul(query_embs, corpus_embs.T) + argsort   (/root/reference/scripts/simple_eval.py:25,35)
//   compute_similarity(q, corpus)[0] + argsort        (/root/reference/src/kd/eval.py:75,86)
//   faiss IndexFlatIP's blocked sgemm + heap path behind FAISSIndexBuilder.search.
//
// Shape of one MMA (tcgen05.mma.cta_group::2.kind::f16, M = 256, N = 256, K = 16): a CTA PAIR
// (2-CTA cluster = one TPC) multiplies a resident block of 256 queries (A operand: 128 queries per
// CTA, bf16, loaded once per work item and kept in shared memory) with a streamed tile of 256
// corpus rows (B operand: each CTA TMA-loads ITS 128 rows, SWIZZLE_128B boxes of 128 rows x 64
// elements = 16 KB per pipeline stage).  Every corpus byte fetched from L2/HBM therefore meets 256
// queries -- the bf16 machine balance of a B200 (flop/byte) -- while each SM only stages half of
// the tile.  The fp32 accumulator [128 queries x 256 rows] of each CTA lives in its TMEM
// (query -> lane, corpus row -> column), double buffered (2 x 256 = all 512 columns) so the MMAs
// of tile t+1 overlap the epilogue of tile t.
//
// Epilogue = the fused select.  An epilogue thread owns ONE query (its TMEM lane): its threshold
// sits in a register, the test of a score is one FSETP against that register, and the rare
// survivors are appended to a list that only this thread writes -- no locks, no atomics, no
// shared-memory state.  Thresholds come from a pre-pass over a row sample (PREPASS = true: the
// same pipeline, but the epilogue only reduces every 32-row group to its maximum; the k-th
// largest group maximum is a lower bound of the k-th best score: seed_select_kernel).  If a list
// still fills up (adversarial data, or no sample) its owner keeps the best k with a private
// quickselect and raises its threshold, so the result is exact for any input.
//
// Work items = (corpus chunk, query block), chunk-major, dealt round-robin to the 74 pairs: pairs
// that hold different query blocks walk the same chunk at the same time, so a chunk is fetched
// from HBM once and re-read from the 126 MB L2 by the others.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only, one
// elected lane), warp 2 = TMEM allocator, warps 4-11 = epilogue (warp w reads TMEM lanes
// 32*(w%4).., column half (w-4)/4).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "select.cuh"

namespace b2s {

constexpr int kTcThreads = 384;
constexpr int kTcEpiWarps = 8;
constexpr int kTcQueriesPerCta = 128;   // UMMA M per CTA
constexpr int kTcQueriesPerPair = 256;  // UMMA M
constexpr int kTcRowsPerCta = 128;      // B rows staged by one CTA
constexpr int kTcTileRows = 256;        // UMMA N = corpus rows per pair tile
constexpr int kTcKBlock = 64;           // bf16 elements per 128-byte swizzle row
constexpr int kTcStageBytes = kTcRowsPerCta * kTcKBlock * 2;     // 16 KB
constexpr int kTcQBlockBytes = kTcQueriesPerCta * kTcKBlock * 2;  // 16 KB per k-block
constexpr int kTcMaxStages = 10;
constexpr int kTcAccStages = 2;
constexpr int kTcGroupsPerTile = kTcTileRows / 32;   // 32-row groups whose maxima the pre-pass records
constexpr int kTcQbDone = -3;                        // s_cur_qb value: the epilogue has finished
constexpr int kTcHistBins = 64;                      // score bins of the shared per-query survivor histogram

struct TcParams {
    uint32_t n_rows;        // rows in the shard
    int tiles_total;        // tiles this launch iterates over (pre-pass: sampled tiles)
    int tile_mul;           // actual tile = index * tile_mul (pre-pass sample stride; 1 otherwise)
    int chunk_tiles;        // tiles per work item
    int num_chunks;         // ceil(tiles_total / chunk_tiles)
    int qblocks;            // query blocks of 256
    int kblocks;            // dim / 64
    int stages;             // B ring depth
    int nq;                 // real queries (rows >= nq of the padded query matrix are zero / masked)
    int nq_pad;             // qblocks * 256
    int k;
    int cap;                // candidate list capacity (>= 2k)
    unsigned long long policy;   // L2 cache hint for corpus tiles
    const u64* seed_keys;   // optional [nq_pad]: a candidate must have key > seed (0 = none)
    u64* lists;             // [2 * pairs, nq_pad, cap]
    int* counts;            // [2 * pairs, nq_pad]   (zeroed by the host before the main pass)
    u64* thr_keys;          // [2 * pairs, nq_pad]   (zeroed by the host before the main pass)
    float* gmax;            // pre-pass out: [nq_pad, groups] maxima of 32-row groups
    int groups;             // tiles_total * 8 (pre-pass)
    // shared per-query threshold tightening (main pass, only with a pre-pass): see hist_publish()
    const uint2* hcfg;      // [nq_pad] {base ord, shift} from seed_select_kernel (shift > 31: disabled)
    unsigned* hist;         // [nq_pad, kTcHistBins] survivors per score bin     (zeroed by the host)
    unsigned* gthr;         // [nq_pad] ord of the best proven lower bound of the k-th score (zeroed)
    int hstep;              // refresh period of the bound-updater warp, ns
};

// dynamic shared memory carve-up (all offsets from a 1024-byte aligned base; identical in both CTAs)
struct TcSmemLayout {
    uint32_t q_off, b_off, bar_off, total;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int kblocks, int stages, int stage_bytes = kTcStageBytes) {
    TcSmemLayout L;
    L.q_off = 0;
    L.b_off = (uint32_t)kblocks * kTcQBlockBytes;
    L.bar_off = L.b_off + (uint32_t)stages * (uint32_t)stage_bytes;
    L.total = L.bar_off + 256;   // 2*10 + 2 + 4 barriers of 8 bytes + the TMEM base address
    return L;
}

// Keep the k largest keys of a[0..n) in a[0..k) (any order) and return the smallest of them.
// Thread-private Hoare quickselect on unique keys; runs only when a list overflows.
__device__ __noinline__ u64 list_keep_top_k(u64* a, int n, int k) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const u64 x = a[lo], y = a[(lo + hi) >> 1], z = a[hi];
        const u64 pivot = x > y ? (y > z ? y : (x > z ? z : x)) : (x > z ? x : (y > z ? z : y));
        int i = lo, j = hi;
        while (i <= j) {
            while (a[i] > pivot) ++i;
            while (a[j] < pivot) --j;
            if (i <= j) {
                const u64 t = a[i];
                a[i] = a[j];
                a[j] = t;
                ++i;
                --j;
            }
        }
        // a[lo..j] >= pivot >= a[i..hi]; positions j+1..i-1 (if any) hold the pivot
        if (k - 1 <= j) hi = j;
        else if (k - 1 >= i) lo = i;
        else break;
    }
    u64 m = a[0];
    for (int i = 1; i < k; ++i) m = a[i] < m ? a[i] : m;
    return m;
}

// Shared per-query threshold tightening.  Every survivor of query q (appended by some thread of some
// CTA pair) is counted in the query's global histogram: bins of 2^shift score images above the
// pre-pass bound `base` (one fire-and-forget RED per survivor).  The lower edge of the highest bin
// b such that bins >= b hold at least k rows is a proven lower bound of the k-th best score (k
// distinct corpus rows reach it); an otherwise idle warp of every CTA re-derives it for the
// queries its CTA currently works on and publishes it with atomicMax, the epilogue threads pick it
// up once per tile.  Counts only grow and stale reads under-count, so the bound is always valid;
// thresholds then track the k-th best of ALL rows seen so far by the whole GPU instead of staying
// at the pre-pass sample's.
__device__ __forceinline__ void hist_count(unsigned* h, uint32_t ord, uint2 cfg) {
    uint32_t bin = (ord - cfg.x) >> cfg.y;
    bin = bin > (uint32_t)(kTcHistBins - 1) ? (uint32_t)(kTcHistBins - 1) : bin;
    atomicAdd(h + bin, 1u);   // result unused: RED
}
// 0 = no bound better than the pre-pass one
__device__ __forceinline__ uint32_t hist_bound(const unsigned* h, unsigned k, uint2 cfg) {
    const uint4* hv = reinterpret_cast<const uint4*>(h);
    unsigned acc = 0;
    int bsel = 0;
    bool found = false;
#pragma unroll
    for (int i = kTcHistBins / 4 - 1; i >= 0; --i) {
        const uint4 v = __ldcg(hv + i);
        acc += v.w; if (!found && acc >= k) { found = true; bsel = 4 * i + 3; }
        acc += v.z; if (!found && acc >= k) { found = true; bsel = 4 * i + 2; }
        acc += v.y; if (!found && acc >= k) { found = true; bsel = 4 * i + 1; }
        acc += v.x; if (!found && acc >= k) { found = true; bsel = 4 * i; }
    }
    return (found && bsel > 0) ? cfg.x + ((uint32_t)bsel << cfg.y) : 0u;
}

// PAIR = true : launched as 2-CTA clusters, one cta_group::2 MMA (M = 256 queries) per pair, each CTA
//               stages 128 of the tile's 256 rows.  For batches above 128 queries.
// PAIR = false: every CTA on its own (cta_group::1, M = 128 queries, it stages all 256 rows of its
//               tiles): for <= 128 queries the second CTA of a pair would multiply padding, and the
//               batch is HBM-bound anyway -- twice the tiles in flight per MMA cycle spent.
template <bool PREPASS, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_corpus, const __grid_constant__ CUtensorMap map_queries,
                 const TcParams p) {
    constexpr int kStageBytes = PAIR ? kTcStageBytes : 2 * kTcStageBytes;
    constexpr int kQueriesPerBlock = PAIR ? kTcQueriesPerPair : kTcQueriesPerCta;
    constexpr uint32_t kCtas = PAIR ? 2u : 1u;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // SWIZZLE_128B tiles need a 1024-byte aligned base; the launch adds 1024 bytes of slack
    unsigned char* smem = tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const TcSmemLayout L = tc_smem_layout(p.kblocks, p.stages, kStageBytes);
    unsigned char* smem_q = smem + L.q_off;
    unsigned char* smem_b = smem + L.b_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* full_bar = bars;                        // [stages]  used in the leader CTA
    uint64_t* empty_bar = bars + kTcMaxStages;        // [stages]  per CTA (multicast commit)
    uint64_t* q_full = bars + 2 * kTcMaxStages;       // [1]       leader
    uint64_t* q_empty = q_full + 1;                   // [1]       per CTA (multicast commit)
    uint64_t* acc_full = q_empty + 1;                 // [2]       per CTA (multicast commit)
    uint64_t* acc_empty = acc_full + kTcAccStages;    // [2]       leader: 16 epilogue warps arrive
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + kTcAccStages);
    int* s_cur_qb = reinterpret_cast<int*>(s_tmem + 1);   // query block the epilogue works on (updater warp)

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;     // 0 = leader
    const int pair = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // index of this work unit (pair or CTA)
    const int num_pairs = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int items_total = p.num_chunks * p.qblocks;
    const bool reload_q = p.qblocks > 1;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_corpus);
        ptx::prefetch_tensormap(&map_queries);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            ptx::mbar_init(&full_bar[i], 1);
            ptx::mbar_init(&empty_bar[i], 1);
        }
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 1);
        for (int i = 0; i < kTcAccStages; ++i) {
            ptx::mbar_init(&acc_full[i], 1);
            ptx::mbar_init(&acc_empty[i], kCtas * kTcEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    if (tid == 0) *s_cur_qb = -1;
    if (warp == 2) {
        if constexpr (PAIR) {
            ptx::tmem_alloc_pair(s_tmem, 512);
            ptx::tmem_relinquish_pair();
        } else {
            ptx::tmem_alloc(s_tmem, 512);
            ptx::tmem_relinquish();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) ptx::cluster_sync_all();   // the peer's barriers exist before anything is signalled on them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on
    // this kernel reads what its predecessors wrote (queries, seeds, zeroed list state)
    grid_dep_wait();

    if (warp == 0) {
        // ================= TMA producer (both CTAs; bytes are counted on the leader's barriers) ====
        if (lane == 0) {
            const uint32_t q_full_leader = PAIR ? ptx::mapa_u32(ptx::smem_u32(q_full), 0) : ptx::smem_u32(q_full);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t q_loads = 0;
            for (int item = pair; item < items_total; item += num_pairs) {
                const int chunk = item / p.qblocks;
                const int qb = item - chunk * p.qblocks;
                if (reload_q || q_loads == 0) {
                    ptx::mbar_wait(q_empty, (q_loads & 1u) ^ 1u);   // MMAs that read the old block are done
                    if (rank == 0) ptx::mbar_expect_tx(q_full, kCtas * (uint32_t)p.kblocks * kTcQBlockBytes);
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        ptx::tma_load_2d_on<PAIR>(smem_q + (size_t)kb * kTcQBlockBytes, &map_queries, q_full_leader,
                                                  kb * kTcKBlock, qb * kQueriesPerBlock + (int)rank * kTcQueriesPerCta,
                                                  ptx::kEvictLast);
                    ++q_loads;
                }
                const int t0 = chunk * p.chunk_tiles;
                const int t1 = min(t0 + p.chunk_tiles, p.tiles_total);
                for (int t = t0; t < t1; ++t) {
                    const int row0 = t * p.tile_mul * kTcTileRows + (int)rank * kTcRowsPerCta;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], kCtas * kTcStageBytes * (PAIR ? 1u : 2u));
                        const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
                        ptx::tma_load_2d_on<PAIR>(smem_b + (size_t)stage * kStageBytes, &map_corpus,
                                                  PAIR ? ptx::mapa_u32(fb, 0) : fb, kb * kTcKBlock, row0, p.policy);
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
            // the leader's last multicast commit (q_empty) must have landed here before this CTA may exit
            if (reload_q && q_loads > 0) ptx::mbar_wait(q_empty, (q_loads & 1u) ^ 1u);
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA, one lane) =================
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = ptx::idesc_bf16_f32(kQueriesPerBlock, kTcTileRows);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            uint32_t q_loads = 0;
            for (int item = pair; item < items_total; item += num_pairs) {
                const int chunk = item / p.qblocks;
                if (reload_q || q_loads == 0) {
                    ptx::mbar_wait(q_full, q_loads & 1u);
                    ptx::tc_fence_after();
                    ++q_loads;
                }
                const int t0 = chunk * p.chunk_tiles;
                const int t1 = min(t0 + p.chunk_tiles, p.tiles_total);
                for (int t = t0; t < t1; ++t) {
                    ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                    ptx::tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * kTcTileRows);
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint64_t da = ptx::smem_desc_sw128(ptx::smem_u32(smem_q + (size_t)kb * kTcQBlockBytes));
                        const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(smem_b + (size_t)stage * kStageBytes));
#pragma unroll
                        for (int kk = 0; kk < kTcKBlock / 16; ++kk) {
                            // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 (16-byte units)
                            ptx::umma_bf16_on<PAIR>(tmem_d, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc,
                                                    (uint32_t)((kb | kk) != 0));
                        }
                        ptx::umma_commit_on<PAIR>(&empty_bar[stage]);   // frees the stage (in both CTAs of a pair)
                        if (++stage == p.stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    ptx::umma_commit_on<PAIR>(&acc_full[acc]);      // accumulator tile complete
                    if (++acc == kTcAccStages) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
                if (reload_q) ptx::umma_commit_on<PAIR>(q_empty);   // the query block may be overwritten
            }
        }
    } else if (warp == 3) {
        // ================= bound updater (see hist_bound) =================
        if constexpr (!PREPASS) {
            if (p.hcfg != nullptr) {
                int last_qb = -1;
                uint2 hc[4];
                uint32_t pub[4];
                while (true) {
                    const int qb = *(volatile int*)s_cur_qb;
                    if (qb == kTcQbDone) break;
                    if (qb < 0) {
                        __nanosleep(2000);
                        continue;
                    }
                    if (qb != last_qb) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int q = qb * kQueriesPerBlock + (int)rank * kTcQueriesPerCta + j * 32 + lane;
                            hc[j] = q < p.nq ? p.hcfg[q] : make_uint2(0u, 0xffffffffu);
                            pub[j] = 0u;
                        }
                        last_qb = qb;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int q = qb * kQueriesPerBlock + (int)rank * kTcQueriesPerCta + j * 32 + lane;
                        if (hc[j].y < 32u) {
                            const uint32_t b = hist_bound(p.hist + (size_t)q * kTcHistBins, (unsigned)p.k, hc[j]);
                            if (b > pub[j]) {
                                atomicMax(p.gthr + q, b);
                                pub[j] = b;
                            }
                        }
                    }
                    __nanosleep(p.hstep);   // refresh period (ns)
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: one thread <-> one query =================
        const int qt = warp & 3;                  // TMEM lane quarter this warp may read
        const int half = (warp - 4) >> 2;         // column half of the accumulator
        const uint32_t lane_addr = (uint32_t)(qt * 32) << 16;
        const uint32_t acc_empty_leader0 = PAIR ? ptx::mapa_u32(ptx::smem_u32(&acc_empty[0]), 0) : ptx::smem_u32(&acc_empty[0]);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = pair; item < items_total; item += num_pairs) {
            const int chunk = item / p.qblocks;
            const int qb = item - chunk * p.qblocks;
            const int q = qb * kQueriesPerBlock + (int)rank * kTcQueriesPerCta + qt * 32 + lane;
            const size_t list_id = (size_t)(pair * 2 + half) * p.nq_pad + q;
            // list state of (this pair, this column half, query q)
            u64* my_list = nullptr;
            u64 tk = 0ull;
            float thr = -INFINITY;
            int cnt = 0;
            if constexpr (!PREPASS) {
                my_list = p.lists + list_id * p.cap;
                cnt = p.counts[list_id];
                tk = p.thr_keys[list_id];
                if (tk == 0ull && p.seed_keys != nullptr) tk = p.seed_keys[q];
                if (tk != 0ull) thr = key_score(tk);
                if (q >= p.nq) {   // padding query: nothing passes
                    thr = INFINITY;
                    tk = ~0ull;
                }
            }
            uint2 hc = make_uint2(0u, 0xffffffffu);
            if constexpr (!PREPASS) {
                if (p.hcfg != nullptr && q < p.nq) hc = p.hcfg[q];
            }
            const bool shared_thr = hc.y < 32u;
            unsigned* gthr_q = shared_thr ? p.gthr + q : nullptr;
            unsigned* hist_q = shared_thr ? p.hist + (size_t)q * kTcHistBins : nullptr;
            if (warp == 4 && lane == 0) *(volatile int*)s_cur_qb = qb;   // tell the bound-updater warp
            const int t0 = chunk * p.chunk_tiles;
            const int t1 = min(t0 + p.chunk_tiles, p.tiles_total);
            // the query's shared bound (other CTA pairs raise it) is fetched ONE TILE AHEAD: when the epilogue is
            // the busy side of the pipeline the accumulator is already waiting, and an L2 round trip per tile in
            // front of the first compare was the largest single stall of the kernel at k = 1000 (ncu, r02)
            unsigned g_next = 0u;
            if constexpr (!PREPASS) {
                if (shared_thr) g_next = ldcg_pinned_u32(gthr_q);
            }
            for (int t = t0; t < t1; ++t) {
                unsigned g = 0u;
                if constexpr (!PREPASS) {
                    g = g_next;
                    if (shared_thr) g_next = ldcg_pinned_u32(gthr_q);   // for the next tile
                }
                ptx::mbar_wait(&acc_full[acc], acc_phase);
                ptx::tc_fence_after();
                if constexpr (!PREPASS) {
                    if (g != 0u) {
                        const u64 gk = ((u64)g << 32) - 1ull;   // every score whose image is >= g passes
                        if (gk > tk) {
                            tk = gk;
                            thr = key_score(tk);
                        }
                    }
                }
                const uint32_t col0 = (uint32_t)(acc * kTcTileRows + half * 128);
                const uint32_t row_base = (uint32_t)(t * p.tile_mul) * kTcTileRows + (uint32_t)half * 128u;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    ptx::tmem_ld_32x32(tmem_base + lane_addr + col0 + (uint32_t)(c * 32), v);
                    if constexpr (PREPASS) {
                        float m = v[0];
#pragma unroll
                        for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
                        p.gmax[(size_t)q * p.groups + (size_t)t * kTcGroupsPerTile + half * 4 + c] = m;
                    } else {
                        // group maxima first: one compare per 8 rows on the hot path, and the
                        // (rare, per-lane) slow path only walks the groups that hold a survivor
                        float g8[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float m = fmaxf(fmaxf(v[8 * g], v[8 * g + 1]), v[8 * g + 2]);
                            m = fmaxf(fmaxf(m, v[8 * g + 3]), v[8 * g + 4]);
                            m = fmaxf(fmaxf(m, v[8 * g + 5]), v[8 * g + 6]);
                            g8[g] = fmaxf(m, v[8 * g + 7]);
                        }
                        if (fmaxf(fmaxf(g8[0], g8[1]), fmaxf(g8[2], g8[3])) >= thr) {
                            const uint32_t row_c = row_base + (uint32_t)(c * 32);
                            // one survivor: append, count, keep the list bounded
                            auto keep = [&](float sc, uint32_t row) {
                                const u64 key = make_key(sc, row);
                                if (key > tk && row < p.n_rows) {
                                    my_list[cnt++] = key;
                                    if (shared_thr) hist_count(hist_q, (uint32_t)(key >> 32), hc);
                                    if (cnt == p.cap) {
                                        tk = list_keep_top_k(my_list, cnt, p.k);
                                        thr = key_score(tk);
                                        cnt = p.k;
                                    }
                                }
                            };
                            // Common case, branch-free up to the append: exactly ONE of this lane's four 8-row groups
                            // reaches the threshold and exactly one of its rows does -- that row is the group maximum.
                            // (The slow path is entered by a third of the warp-chunks at k = 1000; walking 4 x 8
                            // unrolled per-row bodies cost ~260 warp instructions per entry and made the slowest of a
                            // pair's 16 epilogue warps pace the MMA pipe: ncu, profiles/r02_k2_ncu_summary.md.)
                            const bool h0 = g8[0] >= thr, h1 = g8[1] >= thr, h2 = g8[2] >= thr, h3 = g8[3] >= thr;
                            bool handled = false;
                            if ((int)h0 + (int)h1 + (int)h2 + (int)h3 == 1) {
                                float w8[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float a = h0 ? v[i] : v[8 + i];
                                    const float b = h2 ? v[16 + i] : v[24 + i];
                                    w8[i] = (h0 || h1) ? a : b;
                                }
                                int n_hit = 0, idx = 0;
#pragma unroll
                                for (int i = 7; i >= 0; --i) {
                                    const bool hit = w8[i] >= thr;
                                    n_hit += hit ? 1 : 0;
                                    idx = hit ? i : idx;
                                }
                                if (n_hit == 1) {
                                    const int gi = h0 ? 0 : (h1 ? 1 : (h2 ? 2 : 3));
                                    const float sc = h0 ? g8[0] : (h1 ? g8[1] : (h2 ? g8[2] : g8[3]));
                                    keep(sc, row_c + (uint32_t)(gi * 8 + idx));
                                    handled = true;
                                }
                            }
                            if (!handled) {   // several survivors in this lane's chunk: walk the groups that hold one
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    if (g8[g] >= thr) {
#pragma unroll
                                        for (int j = 8 * g; j < 8 * g + 8; ++j) {
                                            if (v[j] >= thr) keep(v[j], row_c + (uint32_t)j);
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(acc_empty_leader0 + (uint32_t)(acc * 8));
                if (++acc == kTcAccStages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
            if constexpr (!PREPASS) {
                if (q < p.nq) {
                    p.counts[list_id] = cnt;
                    p.thr_keys[list_id] = tk;
                }
            }
        }
    }

    if (warp == 4 && lane == 0) *(volatile int*)s_cur_qb = kTcQbDone;
    grid_dep_launch();   // the next kernel of the call (seed select / merge) may be scheduled

    // ---- teardown ---------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) ptx::cluster_sync_all();   // the peer no longer reads this CTA's shared memory / TMEM / barriers
    if (warp == 2) {
        ptx::tc_fence_after();
        if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem_base, 512);
        else ptx::tmem_dealloc(tmem_base, 512);
    }
}

// k-th largest of each query's group maxima -> the seed threshold of the main pass.
// gmax [nq_pad, groups]; out seed_keys[q]: a candidate must have key > seed (0 = no threshold).
// One CTA per query: MSB-first 8-bit radix select on the order-preserving integer image.
constexpr int kSeedThreads = 256;
__global__ void __launch_bounds__(kSeedThreads) seed_select_kernel(const float* __restrict__ gmax, int groups, int k,
                                                                   u64* __restrict__ seed_keys,
                                                                   uint2* __restrict__ hcfg, int in_smem) {
    extern __shared__ uint32_t s_ord[];   // [groups] score images, when they fit (in_smem)
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_rem;
    __shared__ uint32_t s_red[2 * (kSeedThreads / 32)];
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const float* g = gmax + (size_t)q * groups;
    grid_dep_launch();
    grid_dep_wait();     // the pre-pass has written the group maxima
    if (groups < k) {
        if (tid == 0) {
            seed_keys[q] = 0ull;
            if (hcfg) hcfg[q] = make_uint2(0u, 0xffffffffu);
        }
        return;
    }
    // one sweep over global memory (independent loads), min / max on the way
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll 8
    for (int i = tid; i < groups; i += kSeedThreads) {
        const uint32_t o = score_to_ord(g[i]);
        if (in_smem) s_ord[i] = o;
        lo = o < lo ? o : lo;
        hi = o > hi ? o : hi;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint32_t a = __shfl_xor_sync(0xffffffffu, lo, off), b = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((tid & 31) == 0) {
        s_red[tid >> 5] = lo;
        s_red[kSeedThreads / 32 + (tid >> 5)] = hi;
    }
    __syncthreads();
    lo = s_red[0];
    hi = s_red[kSeedThreads / 32];
    for (int w = 1; w < kSeedThreads / 32; ++w) {
        lo = s_red[w] < lo ? s_red[w] : lo;
        hi = s_red[kSeedThreads / 32 + w] > hi ? s_red[kSeedThreads / 32 + w] : hi;
    }
    // MSB-first radix select of the k-th largest image, starting at the highest bit in which the
    // candidates differ (the bits above it are common to all of them: nothing to histogram there)
    int undecided = lo == hi ? 0 : 32 - __clz(lo ^ hi);      // low bits still to fix
    if (tid == 0) {
        s_prefix = undecided >= 32 ? 0u : (hi >> undecided) << undecided;
        s_rem = k;
    }
    __syncthreads();
    while (undecided > 0) {
        const int nb = undecided < 8 ? undecided : 8;
        const int shift = undecided - nb;
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (int i = tid; i < groups; i += kSeedThreads) {
            const uint32_t o = in_smem ? s_ord[i] : score_to_ord(g[i]);
            const bool in = undecided >= 32 || (o >> undecided) == (prefix >> undecided);
            if (in) atomicAdd(&hist[(o >> shift) & ((1u << nb) - 1u)], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int rem = s_rem, b = (1 << nb) - 1;
            for (; b > 0; --b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            s_prefix = prefix | ((uint32_t)b << shift);
            s_rem = rem;
        }
        __syncthreads();
        undecided = shift;
    }
    // s_prefix = image of the k-th largest maximum T; every score >= T must pass: key > (T << 32) - 1
    if (tid == 0) {
        const uint32_t t = s_prefix;
        seed_keys[q] = t ? (((u64)t << 32) - 1ull) : 0ull;
        if (hcfg) {
            // histogram geometry for the main pass: bins of 2^shift images from T up, the sample's best
            // score lands in bin <= 61; anything above goes to the last bin
            const uint32_t range = hi - t;
            uint32_t sh = 0;
            while ((range >> sh) > (uint32_t)(kTcHistBins - 3)) ++sh;
            hcfg[q] = t ? make_uint2(t, sh) : make_uint2(0u, 0xffffffffu);
        }
    }
}

// Host-side state of the tensor path kept in the index handle.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TensorPathState {
    CUtensorMap corpus_map;        // boxes of 128 rows (pair variant: each CTA stages half a tile)
    CUtensorMap corpus_map_full;   // boxes of 256 rows (single-CTA variant)
    bool corpus_map_valid = false;
    PFN_encodeTiled encode = nullptr;
    int max_smem_optin = 0;
    bool attr_set = false;
};

// dims whose resident query block (dim * 256 bytes per CTA) leaves room for >= 4 pipeline stages
inline bool tensor_path_supported(int dim) { return dim % kTcKBlock == 0 && dim >= 64 && dim <= 512; }

}  // namespace b2s

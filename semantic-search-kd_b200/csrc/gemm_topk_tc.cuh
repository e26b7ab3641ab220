// gemm_topk_tc.cuh -- K2: batched exact inner-product search on the 5th-gen tensor cores with a
// fused top-k epilogue (the [N x B] score matrix never exists outside TMEM).
//
// Replaces the reference's dense contraction + per-row sort for query batches:
//   np.matmul(query_embs, corpus_embs.T) + argsort   (/root/reference/scripts/simple_eval.py:25,35)
//   compute_similarity(q, corpus)[0] + argsort        (/root/reference/src/kd/eval.py:75,86)
//   faiss IndexFlatIP's blocked sgemm + heap path behind FAISSIndexBuilder.search.
//
// Layout ("swap-AB"): the 128 corpus rows of a tile are the MMA M dimension (A operand, streamed
// HBM -> shared memory by TMA in 128-row x 64-element SWIZZLE_128B boxes, 16 KB per pipeline
// stage); the CTA's query block (n_tile <= 128 queries, bf16) is the N dimension (B operand),
// loaded once and kept resident in shared memory; K = dim is consumed 64 elements per stage as 4
// tcgen05.mma (K = 16) instructions.  The fp32 accumulator tile [128 rows x n_tile queries] lives
// in TMEM (row -> lane, query -> column), double buffered so the MMA of tile t+1 overlaps the
// epilogue of tile t.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue: thread <-> corpus row; tcgen05.ld 32 query
// columns at a time, compare against the per-query running thresholds (shared memory, broadcast
// reads) and append the rare survivors to the CTA's per-query candidate lists (select.cuh; the
// lists live in global memory / L2, compaction is staged through a shared-memory scratch).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "select.cuh"

namespace b2s {

constexpr int kTcThreads = 256;
constexpr int kTcTileRows = 128;   // UMMA M
constexpr int kTcKBlock = 64;      // bf16 elements per 128-byte swizzle row
constexpr int kTcStageBytes = kTcTileRows * kTcKBlock * 2;  // 16 KB
constexpr int kTcMaxStages = 10;
constexpr int kTcMaxNTile = 128;
constexpr int kTcAccStages = 2;

struct TcParams {
    long long n_rows;      // rows in the shard
    int num_tiles;         // ceil(n_rows / 128)
    int tiles_per_slice;   // contiguous tiles owned by one slice (blockIdx.x)
    int tile_stride;       // 1 = every tile, S = every S-th tile (threshold-seeding pre-pass)
    int kblocks;           // dim / 64
    int n_tile;            // queries per CTA (32 | 64 | 128) = UMMA N
    int stages;            // A-ring depth
    int nq;                // real queries (columns >= nq are masked)
    int nq_pad;            // gridDim.y * n_tile
    int k;
    int cap;               // candidate list capacity (power of two)
    const u64* seed_keys;  // optional [nq_pad] initial thresholds
    u64* lists;            // [gridDim.x, nq_pad, cap]
    int* counts;           // [gridDim.x, nq_pad]
};

// dynamic shared memory carve-up (all offsets from a 1024-byte aligned base)
struct TcSmemLayout {
    uint32_t q_off, a_off, scratch_off, bar_off, state_off, total;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(int kblocks, int n_tile, int stages, int cap) {
    TcSmemLayout L;
    L.q_off = 0;
    L.a_off = (uint32_t)kblocks * n_tile * 128;                  // n_tile*128 is a multiple of 1024
    L.scratch_off = L.a_off + (uint32_t)stages * kTcStageBytes;
    L.bar_off = L.scratch_off + (uint32_t)cap * 8;
    L.state_off = L.bar_off + 256;                              // <= 2*10 + 1 + 4 barriers of 8 bytes
    L.total = L.state_off + (uint32_t)n_tile * (8 + 4 + 4 + 4) + 16;
    return L;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_corpus, const __grid_constant__ CUtensorMap map_queries,
                 const TcParams p) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // SWIZZLE_128B tiles need a 1024-byte aligned base; the launch adds 1024 bytes of slack
    unsigned char* smem = tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const TcSmemLayout L = tc_smem_layout(p.kblocks, p.n_tile, p.stages, p.cap);
    unsigned char* smem_q = smem + L.q_off;
    unsigned char* smem_a = smem + L.a_off;
    u64* scratch = reinterpret_cast<u64*>(smem + L.scratch_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* full_bar = bars;                        // [stages]
    uint64_t* empty_bar = bars + kTcMaxStages;        // [stages]
    uint64_t* q_bar = bars + 2 * kTcMaxStages;        // [1]
    uint64_t* acc_full = q_bar + 1;                   // [2]
    uint64_t* acc_empty = acc_full + kTcAccStages;    // [2]
    u64* s_thr_key = reinterpret_cast<u64*>(smem + L.state_off);           // [n_tile]
    float* s_thr = reinterpret_cast<float*>(s_thr_key + p.n_tile);          // [n_tile]
    int* s_count = reinterpret_cast<int*>(s_thr + p.n_tile);                // [n_tile]
    int* s_lock = s_count + p.n_tile;                                       // [n_tile]
    int* s_scratch_lock = s_lock + p.n_tile;                                // [1]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_scratch_lock + 1);     // [1]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int q0 = blockIdx.y * p.n_tile;             // first query of this CTA's block

    // this CTA's tiles: first_tile, first_tile + stride, ... < tile_end
    const int first_tile = blockIdx.x * p.tiles_per_slice;
    int tile_end = first_tile + p.tiles_per_slice;
    if (tile_end > p.num_tiles) tile_end = p.num_tiles;
    const int my_tiles = first_tile < tile_end ? (tile_end - first_tile + p.tile_stride - 1) / p.tile_stride : 0;

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(kTcAccStages * p.n_tile)) tmem_cols <<= 1;

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
        ptx::mbar_init(q_bar, 1);
        for (int i = 0; i < kTcAccStages; ++i) {
            ptx::mbar_init(&acc_full[i], 1);
            ptx::mbar_init(&acc_empty[i], 4);   // one arrive per epilogue warp
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(s_tmem, tmem_cols);
        ptx::tmem_relinquish();
    }
    for (int j = tid; j < p.n_tile; j += kTcThreads) {
        const ListRef st{&s_thr_key[j], &s_thr[j], &s_count[j], &s_lock[j]};
        if (q0 + j < p.nq) list_init(st, p.seed_keys ? p.seed_keys[q0 + j] : 0ull);
        else list_disable(st);
    }
    if (tid == 0) *s_scratch_lock = 0;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            ptx::mbar_expect_tx(q_bar, (uint32_t)p.kblocks * p.n_tile * 128);
            for (int kb = 0; kb < p.kblocks; ++kb)
                ptx::tma_load_2d(smem_q + (size_t)kb * p.n_tile * 128, &map_queries, q_bar, kb * kTcKBlock, q0,
                                 ptx::kEvictLast);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = first_tile + i * p.tile_stride;
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_expect_tx(&full_bar[stage], kTcStageBytes);
                    ptx::tma_load_2d(smem_a + (size_t)stage * kTcStageBytes, &map_corpus, &full_bar[stage],
                                     kb * kTcKBlock, tile * kTcTileRows, ptx::kEvictFirst);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = ptx::idesc_bf16_f32(kTcTileRows, (uint32_t)p.n_tile);
            ptx::mbar_wait(q_bar, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                ptx::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * p.n_tile);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t da = ptx::smem_desc_sw128(ptx::smem_u32(smem_a + (size_t)stage * kTcStageBytes));
                    const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(smem_q + (size_t)kb * p.n_tile * 128));
#pragma unroll
                    for (int kk = 0; kk < kTcKBlock / 16; ++kk) {
                        // advance 16 elements = 32 bytes inside the 128-byte swizzle row: +2 (16-byte units)
                        ptx::umma_bf16(tmem_d, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc,
                                       (uint32_t)((kb | kk) != 0));
                    }
                    ptx::umma_commit(&empty_bar[stage]);   // frees the smem stage when these MMAs retire
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                ptx::umma_commit(&acc_full[acc]);          // accumulator tile complete
                if (++acc == kTcAccStages) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: threshold filter + candidate lists =================
        const int ew = warp - 4;                           // TMEM lane quarter this warp may read
        u64* my_lists = p.lists + ((size_t)blockIdx.x * p.nq_pad + q0) * p.cap;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = first_tile + i * p.tile_stride;
            const long long row = (long long)tile * kTcTileRows + ew * 32 + lane;
            const bool row_ok = row < p.n_rows;
            ptx::mbar_wait(&acc_full[acc], acc_phase);
            ptx::tc_fence_after();
            for (int c = 0; c < p.n_tile; c += 32) {
                float v[32];
                ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * p.n_tile + c), v);
                unsigned passmask = 0;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 t = ptx::lds_volatile_f4(&s_thr[c + 4 * j4]);
                    passmask |= (v[4 * j4 + 0] >= t.x ? 1u : 0u) << (4 * j4 + 0);
                    passmask |= (v[4 * j4 + 1] >= t.y ? 1u : 0u) << (4 * j4 + 1);
                    passmask |= (v[4 * j4 + 2] >= t.z ? 1u : 0u) << (4 * j4 + 2);
                    passmask |= (v[4 * j4 + 3] >= t.w ? 1u : 0u) << (4 * j4 + 3);
                }
                if (!row_ok) passmask = 0;
                if (__any_sync(0xffffffffu, passmask != 0)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const bool pj = (passmask >> j) & 1u;
                        if (__any_sync(0xffffffffu, pj)) {
                            const int q = c + j;
                            const ListRef st{&s_thr_key[q], &s_thr[q], &s_count[q], &s_lock[q]};
                            list_append_warp(st, my_lists + (size_t)q * p.cap, p.cap, p.k, pj,
                                             make_key(v[j], (uint32_t)row), lane, scratch, s_scratch_lock);
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
            if (++acc == kTcAccStages) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    // ---- teardown ---------------------------------------------------------------------------
    ptx::tc_fence_before();
    __syncthreads();
    for (int j = tid; j < p.n_tile; j += kTcThreads)
        p.counts[(size_t)blockIdx.x * p.nq_pad + q0 + j] = (q0 + j < p.nq) ? s_count[j] : 0;
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, tmem_cols);
    }
}

// Host-side state of the tensor path kept in the index handle.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TensorPathState {
    CUtensorMap corpus_map;
    bool corpus_map_valid = false;
    PFN_encodeTiled encode = nullptr;
    int max_smem_optin = 0;
    bool attr_set = false;
};

inline bool tensor_path_supported(int dim) { return dim % kTcKBlock == 0 && dim >= 64 && dim <= 1024; }

}  // namespace b2s

// rescore.cuh -- optional exact fp32 re-ranking of the bf16 search result ("keep_f32").
//
// The corpus is searched in bf16 (half the HBM bytes).  bf16 rounding can swap neighbours whose
// fp32 scores differ by less than ~1e-3; the north star allows exactly that, but a caller that
// wants ids IDENTICAL to faiss.IndexFlatIP on the fp32 embeddings
// (/root/reference/tests/conftest.py:184-185) keeps an fp32 copy of the rows: the search then asks
// for k + pad bf16 candidates, re-scores them in fp32 (query un-rounded) and re-sorts.  One CTA per
// query; k + pad <= 2048.
#pragma once
#include "select.cuh"

namespace b2s {

constexpr int kRescoreThreads = 256;
constexpr int kRescoreCap = 2048;

__global__ void __launch_bounds__(kRescoreThreads) rescore_f32_kernel(
    const float* __restrict__ rows_f32, long long n_rows, int dim, const void* __restrict__ queries, int q_is_bf16,
    int normalize_q, const float* __restrict__ cand_scores, const long long* __restrict__ cand_ids, int k_in,
    long long id_offset, int k, float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    __shared__ u64 keys[kRescoreCap];
    __shared__ float s_scale;
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* qf = reinterpret_cast<const float*>(queries) + (size_t)q * dim;
    const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(queries) + (size_t)q * dim;
    if (warp == 0) {
        float ss = 0.f;
        if (normalize_q) {
            for (int i = lane; i < dim; i += 32) {
                const float v = q_is_bf16 ? __bfloat162float(qh[i]) : qf[i];
                ss = fmaf(v, v, ss);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        }
        if (lane == 0) s_scale = normalize_q ? (ss > 0.f ? 1.0f / sqrtf(ss) : 0.f) : 1.f;
    }
    int n2 = 2;
    while (n2 < k_in) n2 <<= 1;
    for (int i = k_in + tid; i < n2; i += kRescoreThreads) keys[i] = 0ull;
    __syncthreads();
    const float scale = s_scale;
    for (int j = warp; j < k_in; j += kRescoreThreads / 32) {
        const long long id = cand_ids[(size_t)q * k_in + j];
        const long long row = id - id_offset;
        u64 key = 0ull;
        if (id >= 0 && row >= 0 && row < n_rows) {
            const float* r = rows_f32 + row * dim;
            float acc = 0.f;
            for (int i = lane; i < dim; i += 32) {
                const float v = (q_is_bf16 ? __bfloat162float(qh[i]) : qf[i]) * scale;
                acc = fmaf(v, r[i], acc);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            key = make_key(acc, (uint32_t)row);
        }
        if (lane == 0) keys[j] = key;
    }
    __syncthreads();
    bitonic_sort_desc(keys, n2, tid, kRescoreThreads, BlockSync());
    for (int i = tid; i < k; i += kRescoreThreads) {
        float s = -FLT_MAX;
        long long id = -1;
        if (i < k_in && keys[i] != 0ull) {
            s = key_score(keys[i]);
            id = (long long)key_row(keys[i]) + id_offset;
        }
        out_scores[(size_t)q * k + i] = s;
        out_ids[(size_t)q * k + i] = id;
    }
}

// fp32 rows -> unit-norm fp32 rows in place (cosine metric keeps a normalised fp32 copy). One warp per row.
__global__ void rows_f32_normalize_kernel(float* __restrict__ rows, long long n, int dim) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    float* r = rows + row * dim;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) ss = fmaf(r[i], r[i], ss);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    for (int i = lane; i < dim; i += 32) r[i] *= scale;
}

}  // namespace b2s

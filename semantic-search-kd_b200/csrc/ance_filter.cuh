// ance_filter.cuh -- the selection step of ANCE hard-negative mining on the device, plus the
// row re-scoring it needs.
//
// Replaces the per-query Python loop of ANCEMiner.mine
// (/root/reference/src/mining/miners.py:237-247):
//     max_pos_score = pos_scores.max() if len(pos_scores) > 0 else 0.0
//     adversarial   = [(doc, s) for doc, s in zip(cand_ids, cand_scores) if s >= max_pos_score - margin]
//     adversarial.sort(key=score, reverse=True); hard_negatives = adversarial[:top_k]
// for candidates that are the query's exact top-`k_in` of the WHOLE corpus (already sorted by the
// search), with the query's own positives removed (the reference's candidate lists never contain
// them; a corpus-wide search does).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace b2s {

// out[q, j] = <query q, corpus row ids[q, j]>  (fp32 accumulate over the bf16 row); -FLT_MAX for
// ids outside [0, n_rows).  One warp per (q, j).  round_q: round the query to bf16 first, to
// reproduce the scores of the tensor path bit for bit in the products.
__global__ void score_rows_kernel(const __nv_bfloat16* __restrict__ rows, long long n_rows, int dim,
                                  const void* __restrict__ queries, int q_is_bf16, int round_q, long long nq, int m,
                                  const long long* __restrict__ ids, long long id_offset, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nq * m) return;
    const long long q = w / m;
    const long long id = ids[w] - id_offset;
    if (id < 0 || id >= n_rows) {
        if (lane == 0) out[w] = -FLT_MAX;
        return;
    }
    const __nv_bfloat16* r = rows + id * dim;
    const float* qf = reinterpret_cast<const float*>(queries) + q * dim;
    const __nv_bfloat16* qh = reinterpret_cast<const __nv_bfloat16*>(queries) + q * dim;
    float acc = 0.f;
    for (int i = lane; i < dim; i += 32) {
        float qv = q_is_bf16 ? __bfloat162float(qh[i]) : qf[i];
        if (round_q) qv = __bfloat162float(__float2bfloat16_rn(qv));
        acc = fmaf(qv, __bfloat162float(r[i]), acc);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[w] = acc;
}

// One thread per query walks its sorted candidates.
__global__ void ance_filter_kernel(const float* __restrict__ cand_scores, const long long* __restrict__ cand_ids,
                                   int k_in, const long long* __restrict__ pos_ids,
                                   const float* __restrict__ pos_scores, int n_pos, float margin, int top_k,
                                   long long nq, long long* __restrict__ out_ids, float* __restrict__ out_scores,
                                   int* __restrict__ out_counts) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const long long* pid = pos_ids + q * n_pos;
    float max_pos = 0.f;
    bool any_pos = false;
    for (int j = 0; j < n_pos; ++j) {
        if (pid[j] >= 0) {
            const float s = pos_scores[q * n_pos + j];
            max_pos = any_pos ? fmaxf(max_pos, s) : s;
            any_pos = true;
        }
    }
    const float thr = (any_pos ? max_pos : 0.f) - margin;
    int c = 0;
    for (int j = 0; j < k_in && c < top_k; ++j) {
        const long long id = cand_ids[q * k_in + j];
        if (id < 0) break;
        const float s = cand_scores[q * k_in + j];
        if (!(s >= thr)) break;   // sorted descending: nothing further can pass
        bool is_pos = false;
        for (int t = 0; t < n_pos; ++t) is_pos = is_pos || (pid[t] == id);
        if (is_pos) continue;
        out_ids[q * top_k + c] = id;
        out_scores[q * top_k + c] = s;
        ++c;
    }
    if (out_counts) out_counts[q] = c;
    for (; c < top_k; ++c) {
        out_ids[q * top_k + c] = -1;
        out_scores[q * top_k + c] = -FLT_MAX;
    }
}

// MaxSim aggregation of chunk hits into document hits
// (/root/reference/src/utils/chunk.py:123-148: max chunk score per document).  The chunk hits of a
// query arrive sorted by descending score, so a document's maximum is its FIRST occurrence: keep
// the first hit of every document, in order.  One thread per query.
__global__ void maxsim_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long nq, int k_in,
                              const long long* __restrict__ chunk_to_doc, long long n_chunks, int k_out,
                              float* __restrict__ out_scores, long long* __restrict__ out_docs,
                              int* __restrict__ out_counts) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    long long* od = out_docs + q * k_out;
    float* os = out_scores + q * k_out;
    int c = 0;
    for (int j = 0; j < k_in && c < k_out; ++j) {
        const long long id = ids[q * k_in + j];
        if (id < 0) break;
        const long long doc = (id < n_chunks) ? chunk_to_doc[id] : id;
        bool seen = false;
        for (int t = 0; t < c; ++t) seen = seen || (od[t] == doc);
        if (seen) continue;
        od[c] = doc;
        os[c] = scores[q * k_in + j];
        ++c;
    }
    if (out_counts) out_counts[q] = c;
    for (; c < k_out; ++c) {
        od[c] = -1;
        os[c] = -FLT_MAX;
    }
}

}  // namespace b2s

// util_kernels.cuh -- storage conversion, query preparation, dense similarity, fills.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace b2s {

// fp32 rows -> bf16 rows (round-to-nearest-even), optional L2 normalisation per row
// ("cosine" metric: /root/reference/configs/index.yaml:11,30).  One warp per row.
__global__ void rows_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                        long long n, int dim, int normalize) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* src = in + row * dim;
    float scale = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int i = lane; i < dim; i += 32) ss = fmaf(src[i], src[i], ss);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    __nv_bfloat16* dst = out + row * dim;
    for (int i = lane; i < dim; i += 32) dst[i] = __float2bfloat16_rn(src[i] * scale);
}

// bf16 rows -> bf16 rows with L2 normalisation (cosine metric, bf16 input)
__global__ void rows_bf16_normalize_kernel(const __nv_bfloat16* __restrict__ in,
                                           __nv_bfloat16* __restrict__ out, long long n, int dim) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const __nv_bfloat16* src = in + row * dim;
    float ss = 0.f;
    for (int i = lane; i < dim; i += 32) {
        float v = __bfloat162float(src[i]);
        ss = fmaf(v, v, ss);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    __nv_bfloat16* dst = out + row * dim;
    for (int i = lane; i < dim; i += 32) dst[i] = __float2bfloat16_rn(__bfloat162float(src[i]) * scale);
}

__global__ void rows_bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                        long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < count; i += stride) out[i] = __bfloat162float(in[i]);
}

// Queries (fp32 or bf16) -> fp32 [nq, dim] (scan path) and bf16 [nq_pad, dim] (tensor path; rows
// >= nq are zero), optionally L2-normalised.  One warp per query row.
__global__ void prep_queries_kernel(const void* __restrict__ in, int in_is_bf16, long long nq,
                                    long long nq_pad, int dim, int normalize, float* __restrict__ out_f32,
                                    __nv_bfloat16* __restrict__ out_bf16) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // PDL: the kernel that follows (scan or tensor pre-pass) may be scheduled now; it waits for this grid
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (row >= nq_pad) return;
    if (row >= nq) {
        if (out_bf16)
            for (int i = lane; i < dim; i += 32) out_bf16[row * dim + i] = __float2bfloat16_rn(0.f);
        return;
    }
    const float* f = reinterpret_cast<const float*>(in) + row * dim;
    const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(in) + row * dim;
    float scale = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int i = lane; i < dim; i += 32) {
            float v = in_is_bf16 ? __bfloat162float(h[i]) : f[i];
            ss = fmaf(v, v, ss);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        scale = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    }
    for (int i = lane; i < dim; i += 32) {
        float v = (in_is_bf16 ? __bfloat162float(h[i]) : f[i]) * scale;
        if (out_f32) out_f32[row * dim + i] = v;
        if (out_bf16) out_bf16[row * dim + i] = __float2bfloat16_rn(v);
    }
}

__global__ void fill_empty_kernel(float* scores, long long* ids, long long count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        scores[i] = -FLT_MAX;
        ids[i] = -1;
    }
}

// Dense similarity S[nq, nd] = Q D^T in fp32 (StudentModel.compute_similarity,
// /root/reference/tests/test_student_model.py:104-124).  One warp per document row, queries
// staged through shared memory 8 at a time; products and sums in fp32.
__global__ void similarity_f32_kernel(const float* __restrict__ Q, long long nq,
                                      const float* __restrict__ D, long long nd, int dim,
                                      float* __restrict__ S) {
    extern __shared__ float qs[];  // [8][dim]
    const int lane = threadIdx.x & 31;
    const int warps = blockDim.x >> 5;
    const long long row = (long long)blockIdx.x * warps + (threadIdx.x >> 5);
    for (long long q0 = 0; q0 < nq; q0 += 8) {
        const int qb = (int)((nq - q0) < 8 ? (nq - q0) : 8);
        __syncthreads();
        for (int i = threadIdx.x; i < qb * dim; i += blockDim.x) qs[i] = Q[q0 * dim + i];
        __syncthreads();
        if (row < nd) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const float* d = D + row * dim;
            for (int i = lane; i < dim; i += 32) {
                const float v = d[i];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < qb) acc[j] = fmaf(v, qs[j * dim + i], acc[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
                if (j < qb && lane == 0) S[(q0 + j) * nd + row] = acc[j];
            }
        }
    }
}

}  // namespace b2s

// gemm_topk_host.inl -- host side of the tensor path (included by b2s_api.cu after b2s_index).
// Builds the TMA tensor maps, sizes the pipeline to the shared-memory budget, launches K2 (and
// the optional threshold-seeding pre-pass) and the per-query merge K3.

namespace {

constexpr int kTcQueryChunk = 4096;   // queries per workspace round on the tensor path
constexpr int kTcSeedStride = 64;

int tc_init(b2s_index* idx) {
    TensorPathState& tc = idx->tc;
    if (!tc.encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return fail(B2S_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        tc.encode = reinterpret_cast<PFN_encodeTiled>(fn);
        CUDA_TRY(cudaDeviceGetAttribute(&tc.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, idx->device));
    }
    if (!tc.attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc.max_smem_optin));
        tc.attr_set = true;
    }
    return B2S_OK;
}

int tc_encode_rows(b2s_index* idx, CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)idx->dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)idx->dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)kTcKBlock, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = idx->tc.encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B2S_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return B2S_OK;
}

// Largest query block (32 | 64 | 128) whose resident bf16 copy leaves room for >= 4 pipeline stages.
int tc_pick_ntile(const b2s_index* idx, int64_t nq, int cap) {
    const int kblocks = idx->dim / kTcKBlock;
    int nt = nq <= 32 ? 32 : (nq <= 64 ? 64 : 128);
    while (nt > 32) {
        const TcSmemLayout L = tc_smem_layout(kblocks, nt, 4, cap);
        if ((int)L.total + 1024 <= idx->tc.max_smem_optin) break;
        nt >>= 1;
    }
    return nt;
}

int search_tensor(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                  int64_t* out_ids, cudaStream_t s, bool seed, bool normalize) {
    int rc;
    if ((rc = tc_init(idx)) != B2S_OK) return rc;
    TensorPathState& tc = idx->tc;
    const int cap = list_capacity(k);
    const int kblocks = idx->dim / kTcKBlock;
    if (!tc.corpus_map_valid) {
        if ((rc = tc_encode_rows(idx, &tc.corpus_map, idx->rows, (uint64_t)idx->n, kTcTileRows)) != B2S_OK) return rc;
        tc.corpus_map_valid = true;
    }
    const int num_tiles = (int)((idx->n + kTcTileRows - 1) / kTcTileRows);

    for (int64_t c0 = 0; c0 < nq; c0 += kTcQueryChunk) {
        const int cn = (int)std::min<int64_t>(kTcQueryChunk, nq - c0);
        const int n_tile = tc_pick_ntile(idx, cn, cap);
        const int nqb = (cn + n_tile - 1) / n_tile;
        const int nq_pad = nqb * n_tile;
        // pipeline depth from the shared-memory budget
        int stages = kTcMaxStages;
        while (stages > 2 && (int)tc_smem_layout(kblocks, n_tile, stages, cap).total + 1024 > tc.max_smem_optin) --stages;
        const TcSmemLayout L = tc_smem_layout(kblocks, n_tile, stages, cap);
        if ((int)L.total + 1024 > tc.max_smem_optin)
            return fail(B2S_ERR_UNSUPPORTED, "tensor path: shared memory budget exceeded for this (dim, k)");
        // slices of the corpus: fill the SMs once the query blocks are accounted for
        int slices = std::max(1, idx->num_sms / nqb);
        if (nqb > 1 && nqb < idx->num_sms) slices = std::max(1, (2 * idx->num_sms) / nqb);   // two waves
        slices = std::min(slices, num_tiles);
        const int tiles_per_slice = (num_tiles + slices - 1) / slices;
        slices = (num_tiles + tiles_per_slice - 1) / tiles_per_slice;

        // queries -> bf16 [nq_pad, dim], zero padded, optionally normalised
        if ((rc = idx->ws_qbf16.ensure((size_t)nq_pad * idx->dim * 2)) != B2S_OK) return rc;
        {
            const int warps = 8;
            const unsigned char* qsrc = reinterpret_cast<const unsigned char*>(queries) +
                                        (size_t)c0 * idx->dim * (q_dtype == B2S_DTYPE_BF16 ? 2 : 4);
            prep_queries_kernel<<<(unsigned)((nq_pad + warps - 1) / warps), warps * 32, 0, s>>>(
                qsrc, q_dtype == B2S_DTYPE_BF16, cn, nq_pad, idx->dim, normalize ? 1 : 0, nullptr,
                reinterpret_cast<__nv_bfloat16*>(idx->ws_qbf16.p));
            CUDA_TRY(cudaGetLastError());
            idx->stats.kernel_launches++;
        }
        CUtensorMap qmap;
        if ((rc = tc_encode_rows(idx, &qmap, idx->ws_qbf16.p, (uint64_t)nq_pad, (uint32_t)n_tile)) != B2S_OK) return rc;

        if ((rc = idx->ws_lists.ensure((size_t)slices * nq_pad * cap * sizeof(u64))) != B2S_OK) return rc;
        if ((rc = idx->ws_counts.ensure((size_t)slices * nq_pad * sizeof(int))) != B2S_OK) return rc;
        if (seed && (rc = idx->ws_seed.ensure((size_t)nq_pad * sizeof(u64))) != B2S_OK) return rc;

        if (idx->opt_timing && c0 == 0) cudaEventRecord(idx->ev[1], s);
        for (int pass = seed ? 0 : 1; pass < 2; ++pass) {
            TcParams p;
            p.n_rows = idx->n;
            p.num_tiles = num_tiles;
            p.tiles_per_slice = tiles_per_slice;
            p.tile_stride = pass == 0 ? kTcSeedStride : 1;
            p.kblocks = kblocks;
            p.n_tile = n_tile;
            p.stages = stages;
            p.nq = cn;
            p.nq_pad = nq_pad;
            p.k = k;
            p.cap = cap;
            p.seed_keys = (pass == 1 && seed) ? reinterpret_cast<const u64*>(idx->ws_seed.p) : nullptr;
            p.lists = reinterpret_cast<u64*>(idx->ws_lists.p);
            p.counts = reinterpret_cast<int*>(idx->ws_counts.p);
            dim3 grid((unsigned)slices, (unsigned)nqb, 1);
            gemm_topk_kernel<<<grid, kTcThreads, L.total + 1024, s>>>(tc.corpus_map, qmap, p);
            CUDA_TRY(cudaGetLastError());
            idx->stats.kernel_launches++;
            if (pass == 1) idx->stats.passes += nqb;
            if (pass == 1 && idx->opt_timing && c0 == 0) cudaEventRecord(idx->ev[2], s);

            MergeParams mp;
            memset(&mp, 0, sizeof(mp));
            mp.lists = reinterpret_cast<const u64*>(idx->ws_lists.p);
            mp.counts = reinterpret_cast<const int*>(idx->ws_counts.p);
            mp.num_lists = slices;
            mp.nq_lists = nq_pad;
            mp.cap = cap;
            mp.k = k;
            mp.id_offset = idx->id_offset;
            if (pass == 0) {
                mp.out_kth_key = reinterpret_cast<u64*>(idx->ws_seed.p);
            } else {
                mp.out_scores = out_scores + (size_t)c0 * k;
                mp.out_ids = reinterpret_cast<long long*>(out_ids) + (size_t)c0 * k;
            }
            if ((rc = launch_merge(mp, cn, s)) != B2S_OK) return rc;
            idx->stats.kernel_launches++;
        }
    }
    return B2S_OK;
}

void tensor_path_release(b2s_index* idx) { idx->tc.corpus_map_valid = false; }

}  // namespace

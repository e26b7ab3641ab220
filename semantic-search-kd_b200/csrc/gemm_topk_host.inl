// gemm_topk_host.inl -- host side of the tensor path (included by b2s_api.cu after b2s_index).
// Builds the TMA tensor maps, sizes the pipeline to the shared-memory budget, plans the work
// items, launches the threshold pre-pass (K2 in PREPASS mode + seed_select), the main pass (K2)
// and the per-query merge K3.

namespace {

constexpr int kTcQueryChunk = 4096;   // queries per workspace round on the tensor path

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
        CUDA_TRY(cudaFuncSetAttribute(gemm_topk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      tc.max_smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(gemm_topk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      tc.max_smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(gemm_topk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      tc.max_smem_optin));
        CUDA_TRY(cudaFuncSetAttribute(gemm_topk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      tc.max_smem_optin));
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

// K2 launch: 2-CTA clusters for the pair variant, plain grid for the single-CTA variant.
template <bool PREPASS, bool PAIR>
cudaError_t tc_launch(bool pdl, int ctas, size_t smem, cudaStream_t s, const CUtensorMap& corpus, const CUtensorMap& queries,
                      const TcParams& p) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = PAIR ? 2 : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the kernel calls grid_dep_wait() itself
    at[1].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, gemm_topk_kernel<PREPASS, PAIR>, corpus, queries, p);
}

// Tiles per work item in [lo, hi] that minimises the makespan (rounds x tiles) of dealing
// ceil(tiles / c) * qblocks equal items round-robin to `pairs` CTA pairs.
int tc_pick_chunk(int tiles, int qblocks, int pairs, int lo, int hi) {
    int best = lo;
    long long best_cost = -1;
    for (int c = lo; c <= hi; ++c) {
        const long long items = (long long)((tiles + c - 1) / c) * qblocks;
        const long long rounds = (items + pairs - 1) / pairs;
        const long long cost = rounds * c;
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && c > best)) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

int search_tensor(b2s_index* idx, const void* queries, int q_dtype, int64_t nq, int k, float* out_scores,
                  int64_t* out_ids, cudaStream_t s, bool seed, bool normalize) {
    int rc;
    if ((rc = tc_init(idx)) != B2S_OK) return rc;
    TensorPathState& tc = idx->tc;
    const int cap = list_capacity(k);
    const int kblocks = idx->dim / kTcKBlock;
    const bool pdl = idx->opt_pdl != 0;
    if (!tc.corpus_map_valid) {
        if ((rc = tc_encode_rows(idx, &tc.corpus_map, idx->rows, (uint64_t)idx->n, kTcRowsPerCta)) != B2S_OK) return rc;
        if ((rc = tc_encode_rows(idx, &tc.corpus_map_full, idx->rows, (uint64_t)idx->n, kTcTileRows)) != B2S_OK) return rc;
        tc.corpus_map_valid = true;
    }
    // <= 128 queries: every CTA on its own (M = 128); above: CTA pairs (M = 256).  "units" = work units.
    const bool pair_mode = nq > kTcQueriesPerCta || idx->opt_tc_single_cta == 0;
    const int qblock = pair_mode ? kTcQueriesPerPair : kTcQueriesPerCta;
    const int pairs = pair_mode ? std::max(1, idx->num_sms / 2) : idx->num_sms;
    const int ctas = pair_mode ? 2 * pairs : pairs;
    const int stage_bytes = pair_mode ? kTcStageBytes : 2 * kTcStageBytes;
    const CUtensorMap& corpus_map = pair_mode ? tc.corpus_map : tc.corpus_map_full;
    const int num_lists = 2 * pairs;
    if (num_lists > kMergeMaxLists) return fail(B2S_ERR_UNSUPPORTED, "too many SMs for the merge kernel");
    const int tiles_all = (int)((idx->n + kTcTileRows - 1) / kTcTileRows);
    const int tiles_full = (int)(idx->n / kTcTileRows);

    // pipeline depth from the shared-memory budget
    int stages = kTcMaxStages;
    while (stages > 2 && (int)tc_smem_layout(kblocks, stages, stage_bytes).total + 1024 > tc.max_smem_optin) --stages;
    const TcSmemLayout L = tc_smem_layout(kblocks, stages, stage_bytes);
    if ((int)L.total + 1024 > tc.max_smem_optin)
        return fail(B2S_ERR_UNSUPPORTED, "tensor path: shared memory budget exceeded for this dim");

    // threshold pre-pass: a sample of full tiles whose 32-row group maxima bound the k-th best score
    int sample_tiles = 0, sample_stride = 1;
    if (seed && tiles_full > 0) {
        // Sample density by k (measured at 1024 queries x 8.8M rows, sweep in profiles/): the pre-pass costs
        // 1/div of a main pass, the survivors that take the epilogue's slow path number ~ k * div.
        const int div = idx->opt_tc_sample_div > 0 ? idx->opt_tc_sample_div
                                                   : (idx->opt_tc_shared_thr ? 64 : (k <= 25 ? 64 : (k <= 50 ? 32 : (k <= 150 ? 16 : 8))));
        const int want_groups = std::max(1024, 8 * k);
        int ts = std::max((tiles_full + div - 1) / div,
                          (want_groups + kTcGroupsPerTile - 1) / kTcGroupsPerTile);
        ts = std::min(ts, tiles_full);
        sample_stride = std::max(1, tiles_full / ts);
        sample_tiles = (tiles_full + sample_stride - 1) / sample_stride;
        if (sample_tiles * kTcGroupsPerTile < k) sample_tiles = 0;
    }
    const int groups = sample_tiles * kTcGroupsPerTile;
    idx->stats.seeded = sample_tiles > 0 ? 1 : 0;

    // queries per workspace round: the candidate lists take 2*pairs * round * cap * 8 bytes (2.5 GB at
    // 4096 queries and k <= 256); keep that bound for larger k by shrinking the round
    int round_q = kTcQueryChunk;
    while (round_q > qblock && (int64_t)round_q * cap > (int64_t)kTcQueryChunk * 512) round_q >>= 1;
    for (int64_t c0 = 0; c0 < nq; c0 += round_q) {
        const int cn = (int)std::min<int64_t>(round_q, nq - c0);
        const int qblocks = (cn + qblock - 1) / qblock;
        const int nq_pad = qblocks * qblock;

        // workspace first, so that the list state can be zeroed BEFORE the kernels of this round: they then
        // follow one another without a memset in between and chain through programmatic dependent launch
        const size_t n_state = (size_t)num_lists * nq_pad;
        if ((rc = idx->ws_lists.ensure(n_state * cap * sizeof(u64))) != B2S_OK) return rc;
        if ((rc = idx->ws_counts.ensure(n_state * sizeof(int))) != B2S_OK) return rc;
        if ((rc = idx->ws_thr.ensure(n_state * sizeof(u64))) != B2S_OK) return rc;
        if ((rc = idx->ws_seed.ensure((size_t)nq_pad * sizeof(u64))) != B2S_OK) return rc;
        if (sample_tiles > 0 && (rc = idx->ws_gmax.ensure((size_t)nq_pad * groups * sizeof(float))) != B2S_OK) return rc;
        // shared-threshold state: [hist nq_pad*64 | gthr nq_pad] u32 (zeroed per call) + hcfg
        const bool shared_thr = sample_tiles > 0 && idx->opt_tc_shared_thr != 0;
        const size_t hist_words = (size_t)nq_pad * (kTcHistBins + 1);
        if (shared_thr && (rc = idx->ws_hist.ensure(hist_words * sizeof(unsigned))) != B2S_OK) return rc;
        if (shared_thr && (rc = idx->ws_hcfg.ensure((size_t)nq_pad * sizeof(uint2))) != B2S_OK) return rc;
        CUDA_TRY(cudaMemsetAsync(idx->ws_counts.p, 0, n_state * sizeof(int), s));
        CUDA_TRY(cudaMemsetAsync(idx->ws_thr.p, 0, n_state * sizeof(u64), s));
        if (shared_thr) CUDA_TRY(cudaMemsetAsync(idx->ws_hist.p, 0, hist_words * sizeof(unsigned), s));

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
        if ((rc = tc_encode_rows(idx, &qmap, idx->ws_qbf16.p, (uint64_t)nq_pad, kTcQueriesPerCta)) != B2S_OK) return rc;

        TcParams p;
        memset(&p, 0, sizeof(p));
        p.n_rows = (uint32_t)idx->n;
        p.qblocks = qblocks;
        p.kblocks = kblocks;
        p.stages = stages;
        p.nq = cn;
        p.nq_pad = nq_pad;
        p.k = k;
        p.cap = cap;
        p.policy = qblocks > 1 ? ptx::kEvictNormal : ptx::kEvictFirst;
        p.lists = reinterpret_cast<u64*>(idx->ws_lists.p);
        p.counts = reinterpret_cast<int*>(idx->ws_counts.p);
        p.thr_keys = reinterpret_cast<u64*>(idx->ws_thr.p);

        if (sample_tiles > 0) {
            TcParams pp = p;
            pp.tiles_total = sample_tiles;
            pp.tile_mul = sample_stride;
            pp.chunk_tiles = tc_pick_chunk(sample_tiles, qblocks, pairs, 1, 16);
            pp.num_chunks = (sample_tiles + pp.chunk_tiles - 1) / pp.chunk_tiles;
            pp.policy = ptx::kEvictNormal;
            pp.gmax = reinterpret_cast<float*>(idx->ws_gmax.p);
            pp.groups = groups;
            if (pair_mode) CUDA_TRY((tc_launch<true, true>(pdl, ctas, L.total + 1024, s, corpus_map, qmap, pp)));
            else CUDA_TRY((tc_launch<true, false>(pdl, ctas, L.total + 1024, s, corpus_map, qmap, pp)));
            const int sel_smem = (size_t)groups * 4 <= 40 * 1024 ? groups * 4 : 0;   // images staged in shared memory
            CUDA_TRY(launch_pdl(pdl, seed_select_kernel, dim3((unsigned)nq_pad), dim3(kSeedThreads), (size_t)sel_smem, s,
                                (const float*)pp.gmax, groups, k, reinterpret_cast<u64*>(idx->ws_seed.p),
                                shared_thr ? reinterpret_cast<uint2*>(idx->ws_hcfg.p) : (uint2*)nullptr, sel_smem ? 1 : 0));
            idx->stats.kernel_launches += 2;
            p.seed_keys = reinterpret_cast<const u64*>(idx->ws_seed.p);
        }

        if (shared_thr) {
            p.hcfg = reinterpret_cast<const uint2*>(idx->ws_hcfg.p);
            p.hist = reinterpret_cast<unsigned*>(idx->ws_hist.p);
            p.gthr = p.hist + (size_t)nq_pad * kTcHistBins;
            p.hstep = idx->opt_tc_thr_period_ns;
        }
        p.tiles_total = tiles_all;
        p.tile_mul = 1;
        p.chunk_tiles = qblocks > 1 ? tc_pick_chunk(tiles_all, qblocks, pairs, idx->opt_tc_chunk_lo, idx->opt_tc_chunk_hi)
                                    : tc_pick_chunk(tiles_all, 1, pairs, 8, 32);
        p.num_chunks = (tiles_all + p.chunk_tiles - 1) / p.chunk_tiles;
        if (idx->time_this && c0 == 0) cudaEventRecord(idx->ev[1], s);
        if (pair_mode) CUDA_TRY((tc_launch<false, true>(pdl, ctas, L.total + 1024, s, corpus_map, qmap, p)));
        else CUDA_TRY((tc_launch<false, false>(pdl, ctas, L.total + 1024, s, corpus_map, qmap, p)));
        idx->stats.kernel_launches++;
        idx->stats.passes += qblocks;
        if (idx->time_this && c0 == 0) cudaEventRecord(idx->ev[2], s);

        MergeParams mp;
        memset(&mp, 0, sizeof(mp));
        mp.lists = reinterpret_cast<const u64*>(idx->ws_lists.p);
        mp.counts = reinterpret_cast<const int*>(idx->ws_counts.p);
        mp.num_lists = num_lists;
        mp.nq_lists = nq_pad;
        mp.cap = cap;
        mp.k = k;
        mp.id_offset = idx->id_offset;
        mp.out_scores = out_scores + (size_t)c0 * k;
        mp.out_ids = reinterpret_cast<long long*>(out_ids) + (size_t)c0 * k;
        if ((rc = launch_merge(idx, mp, cn, (int)c0, s)) != B2S_OK) return rc;
        idx->stats.kernel_launches++;
    }
    return B2S_OK;
}

void tensor_path_release(b2s_index* idx) { idx->tc.corpus_map_valid = false; }

}  // namespace

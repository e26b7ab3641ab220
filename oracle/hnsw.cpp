/*
 * oracle/hnsw.cpp -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the approximate index the reference serves from:
 * faiss.IndexHNSWFlat(d, M, METRIC_INNER_PRODUCT) with the reference's
 * parameters M=32, efConstruction=200, efSearch=64
 *   /root/reference/src/config.py:126-139          (FAISSConfig defaults)
 *   /root/reference/configs/index.yaml:7-11,51-56  (HNSW knobs; recall@10 gate)
 *   /root/reference/scripts/build_faiss_index.py:49-62
 *
 * faiss is not vendored in /root/reference and not installable here, so this
 * is written from the published HNSW algorithm (Malkov & Yashunin) with faiss'
 * documented choices: 2*M links on level 0 and M above, level drawn with
 * multiplier 1/ln(M), greedy descent through the upper levels, ef-bounded
 * best-first search on each level <= the node's level during insertion,
 * neighbour selection by the "keep v only if it is closer to the new point
 * than to any already-kept neighbour" heuristic, reverse links re-pruned on
 * overflow, inner product handled as distance = -<q,x>.
 * Numbers produced with it are labelled "HNSW restatement", never "faiss".
 *
 * Used only by tests/, bench.py's reference/cpu_baseline legs and the
 * recall@10 report.  The product path never links or loads it.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <queue>
#include <random>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

struct SpinLock {
    volatile char f = 0;
    void lock() {
        while (__atomic_exchange_n(&f, 1, __ATOMIC_ACQUIRE)) {
            while (f) __builtin_ia32_pause();
        }
    }
    void unlock() { __atomic_store_n(&f, 0, __ATOMIC_RELEASE); }
};

struct Hnsw {
    const float* X = nullptr;  // borrowed, row-major [n, d]
    int64_t n = 0;
    int d = 0;
    int M = 32;
    int efC = 200;
    std::vector<int> level;           // level of each node
    std::vector<int64_t> off0;        // level-0 links: node * 2M
    std::vector<int32_t> links0;      // [n * 2M], -1 = empty
    std::vector<std::vector<int32_t>> linksUp;  // per node: (level) * M slots, -1 = empty
    std::vector<SpinLock> locks;
    int32_t entry = -1;
    int maxLevel = -1;
    std::mutex entryMu;
    int64_t ndis = 0;

    inline float dist(const float* q, int32_t id) const {
        const float* x = X + (int64_t)id * d;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int i = 0;
        for (; i + 8 <= d; i += 8)
            for (int j = 0; j < 8; ++j) acc[j] += q[i + j] * x[i + j];
        float s = 0.f;
        for (; i < d; ++i) s += q[i] * x[i];
        s += ((acc[0] + acc[4]) + (acc[2] + acc[6])) + ((acc[1] + acc[5]) + (acc[3] + acc[7]));
        return -s;
    }
    inline int32_t* nbrs(int32_t id, int lev, int& cap) {
        if (lev == 0) {
            cap = 2 * M;
            return links0.data() + (int64_t)id * 2 * M;
        }
        cap = M;
        return linksUp[id].data() + (int64_t)(lev - 1) * M;
    }
};

typedef std::pair<float, int32_t> DI;  // (distance, id)

// ef-bounded best-first search on one level.  Returns up to ef nearest as a max-heap.
static void search_level(Hnsw& h, const float* q, int32_t ep, float epd, int lev, int ef,
                         std::vector<uint32_t>& visited, uint32_t tag,
                         std::priority_queue<DI>& top, bool locked) {
    std::priority_queue<DI, std::vector<DI>, std::greater<DI>> cand;
    cand.push(DI(epd, ep));
    top.push(DI(epd, ep));
    visited[ep] = tag;
    std::vector<int32_t> buf;
    while (!cand.empty()) {
        DI c = cand.top();
        if (c.first > top.top().first && (int)top.size() >= ef) break;
        cand.pop();
        int cap;
        buf.clear();
        if (locked) h.locks[c.second].lock();
        int32_t* nb = h.nbrs(c.second, lev, cap);
        for (int i = 0; i < cap; ++i) {
            if (nb[i] < 0) break;
            buf.push_back(nb[i]);
        }
        if (locked) h.locks[c.second].unlock();
        for (int32_t v : buf) {
            if (visited[v] == tag) continue;
            visited[v] = tag;
            float dv = h.dist(q, v);
            if ((int)top.size() < ef || dv < top.top().first) {
                cand.push(DI(dv, v));
                top.push(DI(dv, v));
                if ((int)top.size() > ef) top.pop();
            }
        }
    }
}

// neighbour-selection heuristic: input sorted nearest-first
static void shrink(Hnsw& h, std::vector<DI>& in, int maxSize, std::vector<DI>& out) {
    out.clear();
    for (const DI& v : in) {
        bool good = true;
        const float* xv = h.X + (int64_t)v.second * h.d;
        for (const DI& u : out) {
            float duv = h.dist(xv, u.second);
            if (duv < v.first) {
                good = false;
                break;
            }
        }
        if (good) {
            out.push_back(v);
            if ((int)out.size() >= maxSize) return;
        }
    }
}

static void add_link(Hnsw& h, int32_t src, int32_t dst, int lev) {
    int cap;
    int32_t* nb = h.nbrs(src, lev, cap);
    if (nb[cap - 1] < 0) {
        int i = cap - 1;
        while (i > 0 && nb[i - 1] < 0) --i;
        nb[i] = dst;
        return;
    }
    // full: re-select among the existing neighbours plus dst
    const float* xs = h.X + (int64_t)src * h.d;
    std::vector<DI> all;
    all.reserve(cap + 1);
    all.push_back(DI(h.dist(xs, dst), dst));
    for (int i = 0; i < cap; ++i) all.push_back(DI(h.dist(xs, nb[i]), nb[i]));
    std::sort(all.begin(), all.end());
    std::vector<DI> keep;
    shrink(h, all, cap, keep);
    int i = 0;
    for (; i < (int)keep.size(); ++i) nb[i] = keep[i].second;
    for (; i < cap; ++i) nb[i] = -1;
}

static void insert(Hnsw& h, int32_t id, std::vector<uint32_t>& visited, uint32_t& tag) {
    const float* q = h.X + (int64_t)id * h.d;
    int lv = h.level[id];
    int32_t ep;
    int maxL;
    {
        std::lock_guard<std::mutex> g(h.entryMu);
        ep = h.entry;
        maxL = h.maxLevel;
        if (ep < 0) {
            h.entry = id;
            h.maxLevel = lv;
            return;
        }
    }
    float epd = h.dist(q, ep);
    // greedy descent through levels above lv
    for (int l = maxL; l > lv; --l) {
        bool moved = true;
        while (moved) {
            moved = false;
            int cap;
            h.locks[ep].lock();
            int32_t* nb = h.nbrs(ep, l, cap);
            int32_t loc[128];
            int cnt = 0;
            for (int i = 0; i < cap && nb[i] >= 0; ++i) loc[cnt++] = nb[i];
            h.locks[ep].unlock();
            for (int i = 0; i < cnt; ++i) {
                float dv = h.dist(q, loc[i]);
                if (dv < epd) {
                    epd = dv;
                    ep = loc[i];
                    moved = true;
                }
            }
        }
    }
    h.locks[id].lock();
    for (int l = std::min(lv, maxL); l >= 0; --l) {
        std::priority_queue<DI> top;
        ++tag;
        search_level(h, q, ep, epd, l, h.efC, visited, tag, top, true);
        std::vector<DI> cands;
        cands.reserve(top.size());
        while (!top.empty()) {
            if (top.top().second != id) cands.push_back(top.top());
            top.pop();
        }
        std::sort(cands.begin(), cands.end());
        int cap;
        int32_t* mine = h.nbrs(id, l, cap);
        std::vector<DI> keep;
        shrink(h, cands, cap, keep);
        for (int i = 0; i < (int)keep.size(); ++i) mine[i] = keep[i].second;
        // reverse links
        h.locks[id].unlock();
        for (const DI& v : keep) {
            h.locks[v.second].lock();
            add_link(h, v.second, id, l);
            h.locks[v.second].unlock();
        }
        h.locks[id].lock();
        if (!cands.empty()) {
            ep = cands[0].second;
            epd = cands[0].first;
        }
    }
    h.locks[id].unlock();
    if (lv > maxL) {
        std::lock_guard<std::mutex> g(h.entryMu);
        if (lv > h.maxLevel) {
            h.maxLevel = lv;
            h.entry = id;
        }
    }
}

}  // namespace

ORC_API void* orc_hnsw_build(const float* X, int64_t n, int d, int M, int efC, int nthreads,
                             uint64_t seed) {
    if (!X || n <= 0 || d <= 0 || M < 2 || M > 64 || n > 0x7fffffff) return nullptr;
    Hnsw* h = new Hnsw();
    h->X = X;
    h->n = n;
    h->d = d;
    h->M = M;
    h->efC = efC;
    h->level.resize(n);
    h->links0.assign((size_t)n * 2 * M, -1);
    h->linksUp.resize(n);
    h->locks = std::vector<SpinLock>(n);
    // level ~ floor(-ln(U) / ln(M)), fixed-seed generator
    std::mt19937 rng((uint32_t)(seed ? seed : 12345));
    double mult = 1.0 / std::log((double)M);
    for (int64_t i = 0; i < n; ++i) {
        double u = (rng() + 1.0) / 4294967297.0;
        int lv = (int)std::floor(-std::log(u) * mult);
        if (lv > 12) lv = 12;
        h->level[i] = lv;
        if (lv > 0) h->linksUp[i].assign((size_t)lv * M, -1);
    }
#ifdef _OPENMP
    int T = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    int T = 1;
    (void)nthreads;
#endif
    // the first nodes are inserted serially so that the parallel phase starts on a connected graph
    int64_t serial = std::min<int64_t>(n, 1024);
    {
        std::vector<uint32_t> visited(n, 0);
        uint32_t tag = 0;
        for (int64_t i = 0; i < serial; ++i) insert(*h, (int32_t)i, visited, tag);
    }
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
        std::vector<uint32_t> visited(n, 0);
        uint32_t tag = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int64_t i = serial; i < n; ++i) insert(*h, (int32_t)i, visited, tag);
    }
    return h;
}

ORC_API int orc_hnsw_search(void* hv, const float* Q, int64_t nq, int k, int efS, float* D,
                            int64_t* I, int nthreads) {
    Hnsw* h = (Hnsw*)hv;
    if (!h || !Q || nq < 0 || k <= 0) return -1;
#ifdef _OPENMP
    int T = nthreads > 0 ? nthreads : omp_get_max_threads();
#else
    int T = 1;
    (void)nthreads;
#endif
    int ef = std::max(efS, k);
#ifdef _OPENMP
#pragma omp parallel num_threads(T)
#endif
    {
        std::vector<uint32_t> visited(h->n, 0);
        uint32_t tag = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t qi = 0; qi < nq; ++qi) {
            const float* q = Q + qi * (int64_t)h->d;
            int32_t ep = h->entry;
            float epd = h->dist(q, ep);
            for (int l = h->maxLevel; l > 0; --l) {
                bool moved = true;
                while (moved) {
                    moved = false;
                    int cap;
                    int32_t* nb = h->nbrs(ep, l, cap);
                    for (int i = 0; i < cap && nb[i] >= 0; ++i) {
                        float dv = h->dist(q, nb[i]);
                        if (dv < epd) {
                            epd = dv;
                            ep = nb[i];
                            moved = true;
                        }
                    }
                }
            }
            std::priority_queue<DI> top;
            ++tag;
            if (tag == 0) {
                std::fill(visited.begin(), visited.end(), 0u);
                tag = 1;
            }
            search_level(*h, q, ep, epd, 0, ef, visited, tag, top, false);
            std::vector<DI> res;
            while (!top.empty()) {
                res.push_back(top.top());
                top.pop();
            }
            std::sort(res.begin(), res.end());
            for (int i = 0; i < k; ++i) {
                if (i < (int)res.size()) {
                    D[qi * k + i] = -res[i].first;
                    I[qi * k + i] = res[i].second;
                } else {
                    D[qi * k + i] = -3.402823466e+38f;
                    I[qi * k + i] = -1;
                }
            }
        }
    }
    return 0;
}

ORC_API void orc_hnsw_free(void* hv) { delete (Hnsw*)hv; }

ORC_API int orc_hnsw_max_level(void* hv) { return hv ? ((Hnsw*)hv)->maxLevel : -1; }

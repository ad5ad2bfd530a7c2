// iter.cu -- host side of hnsw.iterative_scan (pgvector 0.8 hnswscan.c [RECALL; the reference mount
// has no source, /root/reference/README.md:1]): scan state that survives the first ef_search
// results, batched over many queries.  Kernels: iter_kernel.cuh.
#include "index.h"
#include "iter_kernel.cuh"

#include <cmath>
#include <vector>

struct hb_iter {
    hb_index *ix = nullptr;
    int64_t nq = 0;
    int ef = 0;
    long long max_tuples = 0;
    bool started = false;
    int grid = 1;
    hb::DevBuf q, bits, disc_d, disc_id, disc_n, tuples, gwd, gwi, elem, dist, cnt, misc;
    int gwords = 0, dcap = 0, gcap = 0;
};

namespace hb {
#define HB_DECLI(name) cudaError_t iter_##name(const IterParams &, int, cudaStream_t);
HB_DECLI(f32_l2) HB_DECLI(f32_ip) HB_DECLI(f16_l2) HB_DECLI(f16_ip) HB_DECLI(f32_l1) HB_DECLI(f16_l1)
#undef HB_DECLI
static cudaError_t launch_iter(const hb_index *ix, const IterParams &p, int grid, cudaStream_t s)
{
    const int kind = metric_kind(ix->metric);
    if (ix->dtype == HB_F32) return kind == 2 ? iter_f32_l1(p, grid, s) : kind == 1 ? iter_f32_ip(p, grid, s) : iter_f32_l2(p, grid, s);
    return kind == 2 ? iter_f16_l1(p, grid, s) : kind == 1 ? iter_f16_ip(p, grid, s) : iter_f16_l2(p, grid, s);
}
}   // namespace hb

using namespace hb;

static void iter_release(hb_iter *it)
{
    hb::DevBuf *b[] = { &it->q, &it->bits, &it->disc_d, &it->disc_id, &it->disc_n, &it->tuples, &it->gwd, &it->gwi,
                        &it->elem, &it->dist, &it->cnt, &it->misc };
    for (auto x : b) x->release();
}

static int iter_setup(hb_iter *it, const void *host_queries)
{
    hb_index *ix = it->ix;
    HB_CK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    const int64_t nq = it->nq;
    const size_t qbytes = (size_t) nq * ix->dim * ix->esize;
    it->gwords = (int) ((ix->n + 31) / 32 + 1);
    // room for everything max_scan_tuples can visit plus what ONE more search may evaluate before the limit is noticed
    // (on iid data a resumed ef_search = 1000 search evaluates tens of thousands of rows)
    it->dcap = (int) std::min<long long>(it->max_tuples, ix->n) + std::max(16384, 64 * it->ef);
    if ((long long) it->dcap > ix->n + 64) it->dcap = (int) ix->n + 64;      // every element is discarded at most once
    it->gcap = it->ef + HB_TIE_LIMIT;
    it->grid = (int) std::min<int64_t>((nq + SCAN_WARPS - 1) / SCAN_WARPS, (int64_t) ix->num_sms * 4);
    const size_t state = (size_t) nq * ((size_t) it->gwords * 4 + (size_t) it->dcap * 8);
    size_t free_b = 0, total_b = 0;
    HB_CK(cudaMemGetInfo(&free_b, &total_b));
    if (state > free_b / 2) {
        set_error("hb_iter_begin: %lld resumable scans need %.1f GB of scan state (visited bitmaps + discarded lists); use smaller batches",
                  (long long) nq, state / 1e9);
        return HB_ENOMEM;
    }
    HB_CK(it->q.ensure(qbytes));
    HB_CK(it->bits.ensure((size_t) nq * it->gwords * 4));
    HB_CK(it->disc_d.ensure((size_t) nq * it->dcap * 4));
    HB_CK(it->disc_id.ensure((size_t) nq * it->dcap * 4));
    HB_CK(it->disc_n.ensure((size_t) nq * 4));
    HB_CK(it->tuples.ensure((size_t) nq * 8));
    HB_CK(it->gwd.ensure((size_t) it->grid * SCAN_WARPS * it->gcap * 4));
    HB_CK(it->gwi.ensure((size_t) it->grid * SCAN_WARPS * it->gcap * 4));
    HB_CK(it->elem.ensure((size_t) nq * it->ef * 4));
    HB_CK(it->dist.ensure((size_t) nq * it->ef * 4));
    HB_CK(it->cnt.ensure((size_t) nq * 4));
    HB_CK(it->misc.ensure(64));
    HB_CK(cudaMemsetAsync(it->bits.p, 0, (size_t) nq * it->gwords * 4, s));
    HB_CK(cudaMemsetAsync(it->disc_n.p, 0, (size_t) nq * 4, s));
    HB_CK(cudaMemsetAsync(it->tuples.p, 0, (size_t) nq * 8, s));
    if (ix->metric == HB_COSINE) {
        HB_CK(ix->ws_q.ensure(qbytes));
        HB_CK(cudaMemcpyAsync(ix->ws_q.p, host_queries, qbytes, cudaMemcpyHostToDevice, s));
        const int rc = normalize_dev(ix, ix->ws_q.p, nq, it->q.p, s);
        if (rc) return rc;
    } else {
        HB_CK(cudaMemcpyAsync(it->q.p, host_queries, qbytes, cudaMemcpyHostToDevice, s));
    }
    HB_CK(cudaStreamSynchronize(s));
    return HB_OK;
}

extern "C" {

hb_iter *hb_iter_begin(hb_index *ix, const void *host_queries, int64_t nq, int ef_search, int64_t max_scan_tuples)
{
    if (!ix || !host_queries || nq < 1 || ef_search < 1 || ef_search > 1000 || max_scan_tuples < 1) {
        set_error("hb_iter_begin: bad argument (nq=%lld ef_search=%d max_scan_tuples=%lld)", (long long) nq, ef_search,
                  (long long) max_scan_tuples);
        return nullptr;
    }
    hb_iter *it = new hb_iter();
    it->ix = ix; it->nq = nq; it->ef = ef_search; it->max_tuples = max_scan_tuples;
    if (iter_setup(it, host_queries) != HB_OK) { iter_release(it); delete it; return nullptr; }
    return it;
}

int64_t hb_iter_next(hb_iter *it, int32_t *elem, float *dist, int32_t *cnt)
{
    if (!it || !elem || !dist || !cnt) { set_error("hb_iter_next: NULL argument"); return HB_EINVAL; }
    hb_index *ix = it->ix;
    HB_CK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    IterParams p;
    memset(&p, 0, sizeof p);
    p.g = ix->view();
    p.queries = it->q.p; p.nq = it->nq; p.ef = it->ef; p.mode = it->started ? 1 : 0;
    // visited table of the ef = 1 descent through the upper layers: a hop adds up to m keys, so large m needs more than 1024 slots
    p.upper_slots = 1024;
    while (p.upper_slots < 64 * ix->m && p.upper_slots < 8192) p.upper_slots <<= 1;
    p.max_tuples = it->max_tuples;
    p.bits = it->bits.as<uint32_t>(); p.gwords = it->gwords;
    p.disc_d = it->disc_d.as<float>(); p.disc_id = it->disc_id.as<uint32_t>(); p.dcap = it->dcap;
    p.disc_n = it->disc_n.as<int32_t>(); p.tuples = it->tuples.as<long long>();
    p.gwd = it->gwd.as<float>(); p.gwi = it->gwi.as<uint32_t>(); p.gcap = it->gcap;
    p.out_elem = it->elem.as<int32_t>(); p.out_dist = it->dist.as<float>(); p.out_cnt = it->cnt.as<int32_t>();
    unsigned int *misc = it->misc.as<unsigned int>();
    HB_CK(cudaMemsetAsync(misc, 0, 16, s));
    p.work = misc; p.err = reinterpret_cast<int32_t *>(misc + 1);
    p.totals = ix->d_totals;
    HB_CK(launch_iter(ix, p, it->grid, s));
    it->started = true;
    int32_t err = 0;
    HB_CK(cudaMemcpyAsync(elem, p.out_elem, (size_t) it->nq * it->ef * 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaMemcpyAsync(dist, p.out_dist, (size_t) it->nq * it->ef * 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaMemcpyAsync(cnt, p.out_cnt, (size_t) it->nq * 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaMemcpyAsync(&err, p.err, 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaStreamSynchronize(s));
    if (err == 1) { set_error("a query has more than %d candidates tying exactly at the ef boundary", HB_TIE_LIMIT); return HB_ELIMIT; }
    if (err == 2) { set_error("hb_iter_next: discarded-candidate list full"); return HB_ELIMIT; }
    if (err == 3) { set_error("hb_iter_next: visited table of the upper-layer descent full (m = %d)", ix->m); return HB_ELIMIT; }
    int64_t total = 0;
    for (int64_t i = 0; i < it->nq; i++) total += cnt[i];
    return total;
}

int hb_iter_tuples(hb_iter *it, int64_t *tuples)
{
    if (!it || !tuples) return HB_EINVAL;
    HB_CK(cudaSetDevice(it->ix->device));
    HB_CK(cudaMemcpy(tuples, it->tuples.p, (size_t) it->nq * 8, cudaMemcpyDeviceToHost));
    return HB_OK;
}

// `ORDER BY col <op> $1 LIMIT k` under a WHERE clause the index cannot evaluate, the case
// hnsw.iterative_scan exists for: scans are resumed until k tuples pass the filter (a bitmap over heap
// TIDs: bit t set = tuple t qualifies), the index is exhausted or max_scan_tuples is reached.
// relaxed_order semantics: tuples are kept in the order the scan produces them.
int hb_search_batch_filtered(hb_index *ix, const void *host_queries, int64_t nq, int ef_search, int k,
                             const uint8_t *allowed_bits, int64_t n_bits, int64_t max_scan_tuples, int64_t *out_tids,
                             float *out_dist, int32_t *out_cnt)
{
    if (!ix || !host_queries || !allowed_bits || !out_tids || !out_dist || !out_cnt || k < 1 || nq < 1) {
        set_error("hb_search_batch_filtered: bad argument");
        return HB_EINVAL;
    }
    hb_iter *it = hb_iter_begin(ix, host_queries, nq, ef_search, max_scan_tuples);
    if (!it) return HB_ECUDA;
    std::vector<int32_t> elem((size_t) nq * ef_search), cnt(nq);
    std::vector<float> dist((size_t) nq * ef_search);
    for (int64_t i = 0; i < nq; i++) out_cnt[i] = 0;
    for (int64_t i = 0; i < nq * k; i++) { out_tids[i] = -1; out_dist[i] = INFINITY; }
    int rc = HB_OK;
    for (;;) {
        const int64_t got = hb_iter_next(it, elem.data(), dist.data(), cnt.data());
        if (got < 0) { rc = (int) got; break; }
        if (got == 0) break;
        bool all_full = true;
        for (int64_t q = 0; q < nq; q++) {
            for (int j = 0; j < cnt[q] && out_cnt[q] < k; j++) {
                const int32_t e = elem[(size_t) q * ef_search + j];
                for (int t = ix->h_ntids[e] - 1; t >= 0 && out_cnt[q] < k; t--) {       // heaptids[--heaptidsLength]
                    const int64_t tid = ix->h_tids[(size_t) e * HB_HEAPTIDS + t];
                    if (tid < 0 || tid >= n_bits || !((allowed_bits[tid >> 3] >> (tid & 7)) & 1)) continue;
                    out_tids[q * k + out_cnt[q]] = tid;
                    out_dist[q * k + out_cnt[q]] = dist[(size_t) q * ef_search + j];
                    out_cnt[q]++;
                }
            }
            if (out_cnt[q] < k) all_full = false;
        }
        if (all_full) break;
    }
    hb_iter_end(it);
    return rc;
}

void hb_iter_end(hb_iter *it)
{
    if (!it) return;
    cudaSetDevice(it->ix->device);
    iter_release(it);
    delete it;
}

}   // extern "C"

// iter_kernel.cuh -- hnsw.iterative_scan (pgvector 0.8), batched: scans that can be resumed after
// their first ef_search results.
//
// Role [RECALL; the reference mount has no source, /root/reference/README.md:1]: hnswscan.c
// GetScanItems with a `discarded` heap + ResumeScanItems, hnswutils.c HnswSearchLayer's
// discarded / initVisited = false / tuples arguments.  Per query the state that survives between
// calls lives in HBM: the visited set (one bit per element), the discarded candidates (unsorted
// list; the ef_search nearest are selected when a scan resumes) and the `tuples` count.
//   mode 0 (first):   entry point -> greedy descent (shared-memory visited hash, as scan_kernel)
//                     -> layer-0 search that keeps what it discards
//   mode 1 (resume):  tuples < max_scan_tuples: the ef_search nearest discarded candidates become
//                     the entry list of another layer-0 search over the same visited set;
//                     otherwise the nearest discarded candidate alone is returned
// One warp per query; W lives in a per-warp HBM slice (capacity ef + HB_TIE_LIMIT) so that exact
// ties can never overflow it.  Results are those of oracle/hnsw_oracle.c orc_iter_next.
#pragma once
#include "scan_kernel.cuh"

namespace hb {

struct IterParams {
    GraphView g;
    const void *queries;          // nq x dim, index dtype, normalised when cosine
    int64_t nq;
    int ef, mode, upper_slots;
    long long max_tuples;
    // per-query state
    uint32_t *bits; int gwords;   // nq x gwords
    float *disc_d; uint32_t *disc_id; int dcap;   // nq x dcap
    int32_t *disc_n;              // nq
    long long *tuples;            // nq
    // per-warp scratch: W
    float *gwd; uint32_t *gwi; int gcap;
    // results
    int32_t *out_elem; float *out_dist; int32_t *out_cnt;   // nq x ef, nq
    int32_t *err;                 // 1: tie tail beyond HB_TIE_LIMIT, 2: discarded list full
    unsigned long long *totals;
    unsigned int *work;
};

__host__ __device__ inline size_t iter_warp_smem(int qfloats, int upper_slots)
{
    return ((size_t) qfloats * 4 + (size_t) upper_slots * 4 + 15) & ~(size_t) 15;
}

// remove and return the nearest entry of the discarded list (key: distance, then id)
__device__ __forceinline__ void disc_pop_min(DiscList &ds, int lane, float &out_d, uint32_t &out_id)
{
    float bd = __int_as_float(0x7f800000);
    uint32_t bi = 0xffffffffu;
    int bx = -1;
    for (int i = lane; i < ds.n; i += 32) {
        const float d = ds.d[i];
        const uint32_t id = ds.id[i];
        if (bx < 0 || d < bd || (d == bd && id < bi)) { bd = d; bi = id; bx = i; }
    }
    for (int b = 16; b >= 1; b >>= 1) {
        const float od = __shfl_xor_sync(FULL, bd, b);
        const uint32_t oi = __shfl_xor_sync(FULL, bi, b);
        const int ox = __shfl_xor_sync(FULL, bx, b);
        if (ox >= 0 && (bx < 0 || od < bd || (od == bd && oi < bi))) { bd = od; bi = oi; bx = ox; }
    }
    __syncwarp();
    if (lane == 0 && bx >= 0) { ds.d[bx] = ds.d[ds.n - 1]; ds.id[bx] = ds.id[ds.n - 1]; }
    ds.n--;
    __syncwarp();
    out_d = bd; out_id = bi;
}

template <typename T, int IP, int NV, int G>
__global__ void __launch_bounds__(SCAN_WARPS * 32) iter_scan_kernel(const IterParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const GraphView &g = p.g;
    const int qfloats = g.nvec * Vec<T>::VEC;
    unsigned char *base = smem + iter_warp_smem(qfloats, p.upper_slots) * warp;
    float *q = reinterpret_cast<float *>(base);
    VisitedHash vh;
    vh.tab = reinterpret_cast<uint32_t *>(base + (size_t) qfloats * 4);
    vh.set_overflow(nullptr, 0);
    vh.configure(p.upper_slots);
    const size_t gw = (size_t) blockIdx.x * SCAN_WARPS + warp;
    WList w;
    w.d = p.gwd + gw * p.gcap;
    w.id = p.gwi + gw * p.gcap;
    w.cap = p.gcap;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(p.work, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= (unsigned) p.nq) break;
        const int64_t qi = item;
        __syncwarp();
        stage_query<T>(reinterpret_cast<const T *>(p.queries) + qi * g.dim, g.dim, g.nvec, q, lane);
        __syncwarp();

        VisitedBitmap vb;
        vb.bits = p.bits + (size_t) qi * p.gwords;
        vb.words = p.gwords;
        vb.count = 0;
        DiscList ds;
        ds.d = p.disc_d + (size_t) qi * p.dcap;
        ds.id = p.disc_id + (size_t) qi * p.dcap;
        ds.cap = p.dcap;
        ds.overflow = false;
        ds.n = p.mode == 0 ? 0 : p.disc_n[qi];
        long long tuples = p.mode == 0 ? 0 : p.tuples[qi];
        QueryCounters ctr = { 0, 0, 0 };
        int st = ST_OK, cnt = 0;
        w.L = 0;

        if (p.mode == 0) {
            if (g.entry >= 0) {
                const float d0 = one_distance<T, IP, NV>(g, q, g.entry, lane);
                ctr.n_dist = 1;
                if (lane == 0) { w.d[0] = d0; w.id[0] = (uint32_t) g.entry; }
                w.L = 1;
                __syncwarp();
                for (int lc = g.entry_level; lc >= 1 && st == ST_OK; lc--) {
                    st = wlist_as_entries(w, vh, 1, lane);
                    if (st == ST_OK) st = search_layer<T, IP, NV, G>(g, w, vh, q, 1, lc, lane, ctr);
                }
                if (st == ST_OK) {
                    // layer 0 on the persistent bitmap (zeroed by the host); the entry point counts as a tuple
                    w.L = 1;
                    if (lane == 0) { w.id[0] &= ID_MASK; vb.insert(w.id[0], false); }
                    __syncwarp();
                    tuples = 1;
                    const int before = ctr.n_dist;
                    st = search_layer<T, IP, NV, G, VisitedBitmap, DiscList>(g, w, vb, q, p.ef, 0, lane, ctr, ds);
                    tuples += ctr.n_dist - before;
                }
            }
        } else if (ds.n > 0) {
            if (tuples >= p.max_tuples) {
                // hnsw.max_scan_tuples reached: hand out the remaining candidates one at a time
                float d; uint32_t id;
                disc_pop_min(ds, lane, d, id);
                if (lane == 0) { w.d[0] = d; w.id[0] = id | EXP_BIT; }
                w.L = 1;
                __syncwarp();
            } else {
                const int nep = min(p.ef, ds.n);
                for (int i = 0; i < nep; i++) {
                    float d; uint32_t id;
                    disc_pop_min(ds, lane, d, id);
                    if (lane == 0) { w.d[i] = d; w.id[i] = id; }      // ascending: W stays sorted, unexpanded
                }
                w.L = nep;
                __syncwarp();
                const int before = ctr.n_dist;
                st = search_layer<T, IP, NV, G, VisitedBitmap, DiscList>(g, w, vb, q, p.ef, 0, lane, ctr, ds);
                tuples += ctr.n_dist - before;
            }
        }
        if (st == ST_OK) {
            cnt = min(w.L, p.ef);
            // what stayed behind entry ef-1 (exact ties at the boundary) was evicted too
            for (int b0 = p.ef; b0 < w.L; b0 += 32) {
                const int i = b0 + lane;
                const bool act = i < w.L;
                ds.push_mask(__ballot_sync(FULL, act), act ? w.d[i] : 0.f, act ? (w.id[i] & ID_MASK) : 0u, lane);
            }
        }
        for (int j = lane; j < p.ef; j += 32) {
            p.out_elem[qi * p.ef + j] = j < cnt ? (int32_t) (w.id[j] & ID_MASK) : -1;
            p.out_dist[qi * p.ef + j] = j < cnt ? w.d[j] : __int_as_float(0x7f800000);
        }
        if (lane == 0) {
            p.out_cnt[qi] = cnt;
            p.disc_n[qi] = min(ds.n, ds.cap);
            p.tuples[qi] = tuples;
            if (st != ST_OK) atomicExch(p.err, st == ST_TABLE ? 3 : 1);
            else if (ds.overflow) atomicExch(p.err, 2);
            atomicAdd(p.totals + 0, (unsigned long long) ctr.n_dist);
            atomicAdd(p.totals + 1, (unsigned long long) ctr.n_hop0);
            atomicAdd(p.totals + 2, (unsigned long long) ctr.n_hopu);
        }
    }
}

template <typename T, int IP>
cudaError_t launch_iter_t(const IterParams &p, int grid, cudaStream_t stream)
{
    cudaError_t err = cudaSuccess;
    const size_t smem = iter_warp_smem(p.g.nvec * Vec<T>::VEC, p.upper_slots) * SCAN_WARPS;
#define HB_ICALL(NVV, GG)                                                                          \
    {                                                                                              \
        auto kern = iter_scan_kernel<T, IP, NVV, GG>;                                              \
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        if (err == cudaSuccess) {                                                                  \
            kern<<<grid, SCAN_WARPS * 32, smem, stream>>>(p);                                      \
            err = cudaGetLastError();                                                              \
        }                                                                                          \
    }
    switch (nv_of(p.g.nvec)) {
    case 1: HB_ICALL(1, 8) break;
    case 2: HB_ICALL(2, 8) break;
    case 3: HB_ICALL(3, 4) break;
    case 4: HB_ICALL(4, 4) break;
    case 6: HB_ICALL(6, 4) break;
    case 8: HB_ICALL(8, 2) break;
    default: HB_ICALL(0, 2) break;
    }
#undef HB_ICALL
    return err;
}

}   // namespace hb

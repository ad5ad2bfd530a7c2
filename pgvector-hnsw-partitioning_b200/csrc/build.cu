// build.cu -- host orchestration of the batched insert pipeline (hnswbuild.c InsertTupleInMemory /
// UpdateGraphInMemory / UpdateNeighborsInMemory, hnswutils.c HnswFindElementNeighbors
// [RECALL; reference mount empty, /root/reference/README.md:1]).
//
// Per batch:  upload rows -> build_search_kernel (candidates per layer) -> build_select_kernel
// (heuristic selection + duplicate detection) -> host: fold duplicates into existing elements,
// assign ids, group reverse links by (layer, target) -> build_commit_kernel (AddConnections) ->
// build_link_kernel (HnswUpdateConnection).  The host only moves bookkeeping integers; every
// distance is evaluated on the GPU.  A batch is at most 1/16 of the current graph, and an element
// that would raise the entry level is inserted alone.
#include "index.h"
#include "build_kernel.cuh"

#include <algorithm>
#include <cstring>
#include <tuple>

namespace hb {

#define HB_DECLB(name)                                                                                   \
    cudaError_t build_search_##name(const BuildSearchParams &, int, int, cudaStream_t, bool slow);       \
    cudaError_t build_select_##name(const BuildSelectParams &, int, cudaStream_t);                       \
    cudaError_t build_link_##name(const BuildLinkParams &, int, cudaStream_t);                           \
    cudaError_t nbr_dist_##name(const NbrDistParams &, int, cudaStream_t);
HB_DECLB(f32_l2) HB_DECLB(f32_ip) HB_DECLB(f16_l2) HB_DECLB(f16_ip)
#undef HB_DECLB

#define HB_PICK(fn, ix)                                                                                   \
    ((ix)->dtype == HB_F32 ? ((ix)->metric != HB_L2 ? fn##_f32_ip : fn##_f32_l2)                          \
                           : ((ix)->metric != HB_L2 ? fn##_f16_ip : fn##_f16_l2))

template <typename T>
__global__ void normalize_rows_kernel(const T *__restrict__ in, char *__restrict__ out_rows, size_t row_bytes,
                                      int64_t n, int dim);

// l2_normalize (canonical order, see api.cu normalize_kernel) writing into padded rows
template <typename T>
__global__ void normalize_rows_kernel(const T *__restrict__ in, char *__restrict__ out_rows, size_t row_bytes,
                                      int64_t n, int dim)
{
    constexpr int VEC = Vec<T>::VEC;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const T *src = in + row * dim;
    double acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc[k] = 0.0;
    const int nch = (dim + VEC - 1) / VEC;
    for (int ch = lane; ch < nch; ch += 32) {
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const int e = ch * VEC + k;
            if (e < dim) {
                const double x = (double) (float) src[e];
                acc[k] = acc[k] + x * x;
            }
        }
    }
    double s;
    if constexpr (VEC == 4) s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    else s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    for (int b = 16; b >= 1; b >>= 1) s = s + __shfl_xor_sync(FULL, s, b);
    const double norm = sqrt(s);
    T *dst = reinterpret_cast<T *>(out_rows + row * row_bytes);
    const int padded = (int) (row_bytes / sizeof(T));
    for (int e = lane; e < padded; e += 32) {
        float y = 0.f;
        if (e < dim) {
            const float x = (float) src[e];
            y = norm > 0.0 ? (float) ((double) x / norm) : x;
        }
        dst[e] = (T) y;
    }
}

static int ensure_build_arrays(hb_index *ix, cudaStream_t s)
{
    if (ix->d_nbr0d) return HB_OK;
    const int m2 = 2 * ix->m;
    HB_CK(cudaMalloc(&ix->d_nbr0d, sizeof(float) * ix->cap * m2));
    HB_CK(cudaMalloc(&ix->d_nbrud, sizeof(float) * ix->upper_cap * ix->m));
    if (ix->n == 0) return HB_OK;
    // the graph was loaded: recompute the cached owner->neighbour distances
    NbrDistParams p;
    p.g = ix->view(); p.rows = ix->n; p.deg = m2; p.owner_of_row = nullptr; p.nbr = ix->d_nbr0; p.nbrd = ix->d_nbr0d;
    HB_CK(HB_PICK(nbr_dist, ix)(p, ix->num_sms, s));
    if (ix->upper_rows > 0) {
        std::vector<int32_t> uoff(ix->n), owner(ix->upper_rows, 0);
        HB_CK(cudaMemcpy(uoff.data(), ix->d_uoff, sizeof(int32_t) * ix->n, cudaMemcpyDeviceToHost));
        for (int64_t e = 0; e < ix->n; e++)
            for (int l = 0; l < ix->h_level[e]; l++) owner[uoff[e] + l] = (int32_t) e;
        HB_CK(ix->ws_build[11].ensure(sizeof(int32_t) * ix->upper_rows));
        HB_CK(cudaMemcpyAsync(ix->ws_build[11].p, owner.data(), sizeof(int32_t) * ix->upper_rows, cudaMemcpyHostToDevice, s));
        p.rows = ix->upper_rows; p.deg = ix->m; p.owner_of_row = ix->ws_build[11].as<int32_t>();
        p.nbr = ix->d_nbru; p.nbrd = ix->d_nbrud;
        HB_CK(HB_PICK(nbr_dist, ix)(p, ix->num_sms, s));
        HB_CK(cudaStreamSynchronize(s));
    }
    return HB_OK;
}

static bool all_zero(const char *row, size_t bytes, int esize)
{
    // a vector has zero norm iff every component is +-0 (squares of fp32/fp16 values do not
    // underflow in double)
    if (esize == 4) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
        for (size_t i = 0; i < bytes / 4; i++) if (w[i] & 0x7fffffffu) return false;
    } else {
        const uint16_t *w = reinterpret_cast<const uint16_t *>(row);
        for (size_t i = 0; i < bytes / 2; i++) if (w[i] & 0x7fffu) return false;
    }
    return true;
}

struct Edge { int32_t layer, target, src; float d; };

int64_t build_insert(hb_index *ix, const void *host_vecs, int64_t n_in, const int64_t *heap_tids)
{
    if (n_in == 0) return 0;
    HB_CK(cudaSetDevice(ix->device));
    cudaStream_t s = ix->stream;
    const size_t src_row = (size_t) ix->dim * ix->esize;
    const int m = ix->m, m2 = 2 * m, efc = ix->efc;
    const bool ip = ix->metric != HB_L2;
    (void) ip;

    // tuples to index, in order (HnswCheckNorm drops zero-norm vectors under the cosine opclass)
    std::vector<int64_t> todo;
    todo.reserve(n_in);
    for (int64_t i = 0; i < n_in; i++) {
        if (ix->metric == HB_COSINE && all_zero((const char *) host_vecs + i * src_row, src_row, ix->esize)) continue;
        todo.push_back(i);
    }
    if (ix->n + (int64_t) todo.size() > ix->cap) {
        set_error("index capacity %lld exceeded (%lld + %lld)", (long long) ix->cap, (long long) ix->n, (long long) todo.size());
        return HB_ENOMEM;
    }
    int rc = ensure_build_arrays(ix, s);
    if (rc) return rc;
    ix->h_level.reserve(ix->n + todo.size());
    ix->h_ntids.reserve(ix->n + todo.size());
    ix->h_tids.reserve((ix->n + todo.size()) * HB_HEAPTIDS);

    const int max_batch = ix->opt_build_batch > 0 ? ix->opt_build_batch : 4096;
    int64_t indexed = 0;
    size_t pos = 0;
    std::vector<char> stage;
    std::vector<uint8_t> levels;
    std::vector<int32_t> ucand_row, final_id, dest_urow, sel0_id, sel0_cnt, selu_id, selu_cnt, dup, status, new_uoff;
    std::vector<float> sel0_d, selu_d;
    std::vector<Edge> edges;
    std::vector<int32_t> seg_off, seg_target, seg_layer, edge_src;
    std::vector<float> edge_d;

    while (pos < todo.size()) {
        const int64_t cur = ix->n;
        int64_t b = std::max<int64_t>(1, std::min<int64_t>(max_batch, cur / 16));
        b = std::min<int64_t>(b, (int64_t) todo.size() - pos);
        levels.resize(b);
        for (int64_t i = 0; i < b; i++) levels[i] = (uint8_t) level_for(ix->seed, ix->seq + i, m);
        // an element above the current entry level becomes the entry point: insert it alone
        if (cur > 0) {
            if (levels[0] > ix->entry_level) b = 1;
            else for (int64_t i = 1; i < b; i++) if (levels[i] > ix->entry_level) { b = i; break; }
        } else b = 1;
        levels.resize(b);

        // ---- rows into HBM at [cur, cur + b)
        stage.resize((size_t) b * src_row);
        for (int64_t i = 0; i < b; i++) memcpy(&stage[i * src_row], (const char *) host_vecs + todo[pos + i] * src_row, src_row);
        char *rows = ix->d_vecs + (size_t) cur * ix->row_bytes;
        if (ix->metric == HB_COSINE) {
            HB_CK(ix->ws_build[0].ensure((size_t) b * src_row));
            HB_CK(cudaMemcpyAsync(ix->ws_build[0].p, stage.data(), (size_t) b * src_row, cudaMemcpyHostToDevice, s));
            const int wpb = 8, grid = (int) ((b + wpb - 1) / wpb);
            if (ix->dtype == HB_F32)
                normalize_rows_kernel<float><<<grid, wpb * 32, 0, s>>>(ix->ws_build[0].as<float>(), rows, ix->row_bytes, b, ix->dim);
            else
                normalize_rows_kernel<__half><<<grid, wpb * 32, 0, s>>>(ix->ws_build[0].as<__half>(), rows, ix->row_bytes, b, ix->dim);
            HB_CK(cudaGetLastError());
        } else if (src_row == ix->row_bytes) {
            HB_CK(cudaMemcpyAsync(rows, stage.data(), src_row * b, cudaMemcpyHostToDevice, s));
        } else {
            HB_CK(cudaMemsetAsync(rows, 0, ix->row_bytes * b, s));
            HB_CK(cudaMemcpy2DAsync(rows, ix->row_bytes, stage.data(), src_row, src_row, b, cudaMemcpyHostToDevice, s));
        }

        auto tid_of = [&](int64_t i) { return heap_tids ? heap_tids[todo[pos + i]] : todo[pos + i]; };

        if (cur == 0) {
            // first element: becomes the entry point, no neighbours
            const int lv = levels[0];
            if (lv > ix->upper_cap) { set_error("upper layer table full"); return HB_ENOMEM; }
            int32_t uo = lv > 0 ? 0 : -1;
            HB_CK(cudaMemcpyAsync(ix->d_uoff, &uo, sizeof uo, cudaMemcpyHostToDevice, s));
            HB_CK(cudaStreamSynchronize(s));
            ix->h_level.push_back((uint8_t) lv);
            ix->h_ntids.push_back(1);
            ix->h_tids.resize(HB_HEAPTIDS, 0);
            ix->h_tids[0] = tid_of(0);
            ix->upper_rows = lv; ix->n = 1; ix->entry = 0; ix->entry_level = lv; ix->seq += 1;
            // stale lists from a previous life of this slot
            HB_CK(cudaMemset(ix->d_nbr0, 0xff, sizeof(int32_t) * m2));
            if (lv > 0) HB_CK(cudaMemset(ix->d_nbru, 0xff, sizeof(int32_t) * (size_t) lv * m));
            // heap TIDs of the first element
            {
                int64_t t0 = ix->h_tids[0];
                uint8_t one = 1;
                HB_CK(cudaMemcpy(ix->d_tid0, &t0, sizeof t0, cudaMemcpyHostToDevice));
                HB_CK(cudaMemcpy(ix->d_ntids, &one, 1, cudaMemcpyHostToDevice));
            }
            pos += 1; indexed += 1;
            continue;
        }

        // ---- candidate search
        const int EL = ix->entry_level;
        ucand_row.assign(b, -1);
        int UR = 0;
        for (int64_t i = 0; i < b; i++) {
            const int l = std::min<int>(levels[i], EL);
            if (l > 0) { ucand_row[i] = UR; UR += l; }
        }
        hb::DevBuf *W = ix->ws_build;
        HB_CK(W[1].ensure(b));                                           // level
        HB_CK(W[2].ensure(sizeof(int32_t) * b * 3));                     // ucand_row | final_id | new uoff
        HB_CK(W[3].ensure((size_t) b * efc * 8 + sizeof(int32_t) * b));  // cand0 id | d | cnt
        HB_CK(W[4].ensure((size_t) std::max(UR, 1) * efc * 8 + sizeof(int32_t) * std::max(UR, 1)));
        HB_CK(W[5].ensure((size_t) b * m2 * 8 + sizeof(int32_t) * b));   // sel0 id | d | cnt
        HB_CK(W[6].ensure((size_t) std::max(UR, 1) * m * 8 + sizeof(int32_t) * std::max(UR, 1) * 2));   // selu id | d | cnt | dest_urow
        HB_CK(W[7].ensure(sizeof(int32_t) * b * DUP_SLOTS));
        HB_CK(W[8].ensure(sizeof(int32_t) * b * 2 + 64));                // status | slow list
        HB_CK(ix->ws_misc.ensure(256));
        HB_CK(cudaMemcpyAsync(W[1].p, levels.data(), b, cudaMemcpyHostToDevice, s));
        HB_CK(cudaMemcpyAsync(W[2].p, ucand_row.data(), sizeof(int32_t) * b, cudaMemcpyHostToDevice, s));
        unsigned int *misc = ix->ws_misc.as<unsigned int>();
        HB_CK(cudaMemsetAsync(misc, 0, 16, s));

        BuildSearchParams sp;
        memset(&sp, 0, sizeof sp);
        sp.g = ix->view();
        sp.first = cur; sp.B = (int) b;
        sp.level = W[1].as<uint8_t>();
        sp.ucand_row = W[2].as<int32_t>();
        sp.efc = efc;
        sp.capW = ((efc + 16 + 3) / 4) * 4;
        {
            int slots = 1; while (slots < efc * 16) slots <<= 1;
            if (slots < 1024) slots = 1024;
            const size_t fixed = (size_t) ix->nvec * (ix->dtype == HB_F32 ? 4 : 8) * 4 + (size_t) sp.capW * 8 + 16;
            while (slots > 256 && (fixed + (size_t) slots * 4) * BUILD_WARPS > 200 * 1024) slots >>= 1;
            sp.slots = slots;
            sp.upper_slots = std::min(slots, 1024);
        }
        sp.cand0_id = W[3].as<int32_t>();
        sp.cand0_d = reinterpret_cast<float *>(sp.cand0_id + (size_t) b * efc);
        sp.cand0_cnt = reinterpret_cast<int32_t *>(sp.cand0_d + (size_t) b * efc);
        sp.candu_id = W[4].as<int32_t>();
        sp.candu_d = reinterpret_cast<float *>(sp.candu_id + (size_t) std::max(UR, 1) * efc);
        sp.candu_cnt = reinterpret_cast<int32_t *>(sp.candu_d + (size_t) std::max(UR, 1) * efc);
        sp.status = W[8].as<int32_t>();
        sp.slow_list = sp.status + b;
        sp.slow_count = reinterpret_cast<int32_t *>(misc + 2);
        sp.totals = ix->d_totals;
        sp.work = misc + 0;
        const int slow_grid = 16;
        const int64_t slow_warps = (int64_t) slow_grid * BUILD_WARPS;
        sp.gwords = (int) ((cur + b + 31) / 32 + 1);
        sp.gcap = efc + HB_TIE_LIMIT;
        HB_CK(ix->ws_gbits.ensure(sizeof(uint32_t) * slow_warps * sp.gwords));
        HB_CK(ix->ws_gwd.ensure(sizeof(float) * slow_warps * sp.gcap));
        HB_CK(ix->ws_gwi.ensure(sizeof(uint32_t) * slow_warps * sp.gcap));
        sp.gbits = ix->ws_gbits.as<uint32_t>(); sp.gwd = ix->ws_gwd.as<float>(); sp.gwi = ix->ws_gwi.as<uint32_t>();
        sp.oslots = 4096;
        HB_CK(ix->ws_ovf.ensure(sizeof(uint32_t) * (size_t) ix->num_sms * MAX_CTAS_PER_SM * BUILD_WARPS * sp.oslots));
        sp.ovf = ix->ws_ovf.as<uint32_t>();
        HB_CK(HB_PICK(build_search, ix)(sp, ix->num_sms, slow_grid, s, false));
        BuildSearchParams sps = sp;
        sps.work = misc + 1; sps.qlist = sp.slow_list; sps.qcount = sp.slow_count;
        HB_CK(HB_PICK(build_search, ix)(sps, ix->num_sms, slow_grid, s, true));

        // ---- neighbour selection
        BuildSelectParams lp;
        memset(&lp, 0, sizeof lp);
        lp.g = sp.g; lp.first = cur; lp.B = (int) b; lp.UR = UR; lp.efc = efc;
        lp.cand0_id = sp.cand0_id; lp.cand0_d = sp.cand0_d; lp.cand0_cnt = sp.cand0_cnt;
        lp.candu_id = sp.candu_id; lp.candu_d = sp.candu_d; lp.candu_cnt = sp.candu_cnt;
        lp.sel0_id = W[5].as<int32_t>();
        lp.sel0_d = reinterpret_cast<float *>(lp.sel0_id + (size_t) b * m2);
        lp.sel0_cnt = reinterpret_cast<int32_t *>(lp.sel0_d + (size_t) b * m2);
        lp.selu_id = W[6].as<int32_t>();
        lp.selu_d = reinterpret_cast<float *>(lp.selu_id + (size_t) std::max(UR, 1) * m);
        lp.selu_cnt = reinterpret_cast<int32_t *>(lp.selu_d + (size_t) std::max(UR, 1) * m);
        int32_t *d_dest_urow = lp.selu_cnt + std::max(UR, 1);
        lp.dup = W[7].as<int32_t>();
        lp.totals = ix->d_totals;
        HB_CK(HB_PICK(build_select, ix)(lp, ix->num_sms, s));

        // ---- bookkeeping on the host
        sel0_id.resize((size_t) b * m2); sel0_d.resize((size_t) b * m2); sel0_cnt.resize(b);
        selu_id.resize((size_t) std::max(UR, 1) * m); selu_d.resize((size_t) std::max(UR, 1) * m); selu_cnt.resize(std::max(UR, 1));
        dup.resize((size_t) b * DUP_SLOTS); status.resize(b);
        HB_CK(cudaMemcpyAsync(sel0_id.data(), lp.sel0_id, sizeof(int32_t) * b * m2, cudaMemcpyDeviceToHost, s));
        HB_CK(cudaMemcpyAsync(sel0_d.data(), lp.sel0_d, sizeof(float) * b * m2, cudaMemcpyDeviceToHost, s));
        HB_CK(cudaMemcpyAsync(sel0_cnt.data(), lp.sel0_cnt, sizeof(int32_t) * b, cudaMemcpyDeviceToHost, s));
        if (UR > 0) {
            HB_CK(cudaMemcpyAsync(selu_id.data(), lp.selu_id, sizeof(int32_t) * UR * m, cudaMemcpyDeviceToHost, s));
            HB_CK(cudaMemcpyAsync(selu_d.data(), lp.selu_d, sizeof(float) * UR * m, cudaMemcpyDeviceToHost, s));
            HB_CK(cudaMemcpyAsync(selu_cnt.data(), lp.selu_cnt, sizeof(int32_t) * UR, cudaMemcpyDeviceToHost, s));
        }
        HB_CK(cudaMemcpyAsync(dup.data(), lp.dup, sizeof(int32_t) * b * DUP_SLOTS, cudaMemcpyDeviceToHost, s));
        HB_CK(cudaMemcpyAsync(status.data(), sp.status, sizeof(int32_t) * b, cudaMemcpyDeviceToHost, s));
        HB_CK(cudaStreamSynchronize(s));
        for (int64_t i = 0; i < b; i++)
            if (status[i] < 0) {
                set_error("insert: more than %d candidates tie exactly at the ef_construction boundary", HB_TIE_LIMIT);
                return HB_ELIMIT;
            }

        // FindDuplicateInMemory: fold a tuple into the first byte-identical neighbour with room
        final_id.assign(b, -1);
        dest_urow.assign(std::max(UR, 1), -1);
        new_uoff.assign(b, -1);
        int64_t next = cur;
        int64_t urows = ix->upper_rows;
        bool any_dup = false;
        std::vector<int32_t> dirty;
        for (int64_t i = 0; i < b; i++) {
            int32_t into = -1;
            for (int k = 0; k < DUP_SLOTS; k++) {
                const int32_t c = dup[i * DUP_SLOTS + k];
                if (c < 0) break;
                if (ix->h_ntids[c] < HB_HEAPTIDS) { into = c; break; }
            }
            if (into >= 0) {
                ix->h_tids[(size_t) into * HB_HEAPTIDS + ix->h_ntids[into]++] = tid_of(i);
                dirty.push_back(into);
                any_dup = true;
                continue;
            }
            final_id[i] = (int32_t) next++;
            ix->h_level.push_back(levels[i]);
            ix->h_ntids.push_back(1);
            ix->h_tids.resize(ix->h_tids.size() + HB_HEAPTIDS, 0);
            ix->h_tids[(size_t) final_id[i] * HB_HEAPTIDS] = tid_of(i);
            if (levels[i] > 0) {
                new_uoff[i] = (int32_t) urows;
                const int l = std::min<int>(levels[i], EL);
                for (int r = 0; r < l; r++) dest_urow[ucand_row[i] + r] = (int32_t) (urows + r);
                urows += levels[i];
            }
        }
        if (urows > ix->upper_cap) { set_error("upper layer table full (%lld rows)", (long long) urows); return HB_ENOMEM; }
        const int64_t nb_new = next - cur;

        if (any_dup && nb_new > 0) {
            // close the gaps the folded tuples left in [cur, cur + b)
            HB_CK(ix->ws_build[0].ensure((size_t) b * ix->row_bytes));
            HB_CK(cudaMemcpyAsync(ix->ws_build[0].p, rows, (size_t) b * ix->row_bytes, cudaMemcpyDeviceToDevice, s));
            for (int64_t i = 0; i < b; i++)
                if (final_id[i] >= 0 && final_id[i] != cur + i)
                    HB_CK(cudaMemcpyAsync(ix->d_vecs + (size_t) final_id[i] * ix->row_bytes,
                                          ix->ws_build[0].as<char>() + (size_t) i * ix->row_bytes, ix->row_bytes,
                                          cudaMemcpyDeviceToDevice, s));
        }

        if (nb_new > 0) {
            // uoff of the new elements, clean list rows, then AddConnections
            std::vector<int32_t> uo(nb_new);
            for (int64_t i = 0; i < b; i++) if (final_id[i] >= 0) uo[final_id[i] - cur] = new_uoff[i];
            HB_CK(cudaMemcpyAsync(ix->d_uoff + cur, uo.data(), sizeof(int32_t) * nb_new, cudaMemcpyHostToDevice, s));
            HB_CK(cudaMemsetAsync(ix->d_nbr0 + (size_t) cur * m2, 0xff, sizeof(int32_t) * nb_new * m2, s));
            if (urows > ix->upper_rows)
                HB_CK(cudaMemsetAsync(ix->d_nbru + (size_t) ix->upper_rows * m, 0xff, sizeof(int32_t) * (urows - ix->upper_rows) * m, s));
            int32_t *d_final = W[2].as<int32_t>() + b;
            HB_CK(cudaMemcpyAsync(d_final, final_id.data(), sizeof(int32_t) * b, cudaMemcpyHostToDevice, s));
            HB_CK(cudaMemcpyAsync(d_dest_urow, dest_urow.data(), sizeof(int32_t) * std::max(UR, 1), cudaMemcpyHostToDevice, s));
            BuildCommitParams cp;
            cp.B = (int) b; cp.UR = UR; cp.m = m;
            cp.final_id = d_final; cp.dest_urow = d_dest_urow;
            cp.sel0_id = lp.sel0_id; cp.sel0_d = lp.sel0_d; cp.selu_id = lp.selu_id; cp.selu_d = lp.selu_d;
            cp.nbr0 = ix->d_nbr0; cp.nbr0d = ix->d_nbr0d; cp.nbru = ix->d_nbru; cp.nbrud = ix->d_nbrud;
            const int64_t threads = (int64_t) b * m2 + (int64_t) UR * m;
            build_commit_kernel<<<(int) ((threads + 255) / 256), 256, 0, s>>>(cp);
            HB_CK(cudaGetLastError());

            // reverse links grouped by (layer, target), sources ascending
            edges.clear();
            for (int64_t i = 0; i < b; i++) {
                if (final_id[i] < 0) continue;
                for (int j = 0; j < sel0_cnt[i]; j++)
                    edges.push_back({ 0, sel0_id[i * m2 + j], final_id[i], sel0_d[i * m2 + j] });
                const int l = std::min<int>(levels[i], EL);
                for (int lc = 1; lc <= l; lc++) {
                    const int r = ucand_row[i] + (lc - 1);
                    for (int j = 0; j < selu_cnt[r]; j++)
                        edges.push_back({ lc, selu_id[(size_t) r * m + j], final_id[i], selu_d[(size_t) r * m + j] });
                }
            }
            std::sort(edges.begin(), edges.end(), [](const Edge &a, const Edge &c) {
                return std::tie(a.layer, a.target, a.src) < std::tie(c.layer, c.target, c.src);
            });
            seg_off.clear(); seg_target.clear(); seg_layer.clear();
            edge_src.resize(edges.size()); edge_d.resize(edges.size());
            for (size_t e = 0; e < edges.size(); e++) {
                if (e == 0 || edges[e].layer != edges[e - 1].layer || edges[e].target != edges[e - 1].target) {
                    seg_off.push_back((int32_t) e);
                    seg_target.push_back(edges[e].target);
                    seg_layer.push_back(edges[e].layer);
                }
                edge_src[e] = edges[e].src; edge_d[e] = edges[e].d;
            }
            seg_off.push_back((int32_t) edges.size());
            const int S = (int) seg_target.size();
            if (S > 0) {
                const size_t E = edges.size();
                HB_CK(W[9].ensure(sizeof(int32_t) * (3 * (size_t) S + 1) + 8 * E));
                int32_t *d_seg_off = W[9].as<int32_t>();
                int32_t *d_seg_target = d_seg_off + S + 1;
                int32_t *d_seg_layer = d_seg_target + S;
                int32_t *d_edge_src = d_seg_layer + S;
                float *d_edge_d = reinterpret_cast<float *>(d_edge_src + E);
                HB_CK(cudaMemcpyAsync(d_seg_off, seg_off.data(), sizeof(int32_t) * (S + 1), cudaMemcpyHostToDevice, s));
                HB_CK(cudaMemcpyAsync(d_seg_target, seg_target.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice, s));
                HB_CK(cudaMemcpyAsync(d_seg_layer, seg_layer.data(), sizeof(int32_t) * S, cudaMemcpyHostToDevice, s));
                HB_CK(cudaMemcpyAsync(d_edge_src, edge_src.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
                HB_CK(cudaMemcpyAsync(d_edge_d, edge_d.data(), sizeof(float) * E, cudaMemcpyHostToDevice, s));
                BuildLinkParams kp;
                memset(&kp, 0, sizeof kp);
                // the link kernel reads rows and uoff of old and new elements alike
                ix->n = next; ix->upper_rows = urows;
                kp.g = ix->view();
                kp.S = S; kp.seg_off = d_seg_off; kp.seg_target = d_seg_target; kp.seg_layer = d_seg_layer;
                kp.edge_src = d_edge_src; kp.edge_d = d_edge_d;
                kp.nbr0 = ix->d_nbr0; kp.nbr0d = ix->d_nbr0d; kp.nbru = ix->d_nbru; kp.nbrud = ix->d_nbrud;
                kp.totals = ix->d_totals;
                HB_CK(HB_PICK(build_link, ix)(kp, ix->num_sms, s));
            }
            // entry point: the first element whose level exceeds the current entry level
            for (int64_t i = 0; i < b; i++)
                if (final_id[i] >= 0 && levels[i] > ix->entry_level) { ix->entry = final_id[i]; ix->entry_level = levels[i]; }
        }
        ix->n = next; ix->upper_rows = urows;
        ix->seq += b;
        HB_CK(cudaStreamSynchronize(s));   // host vectors above are reused by the next batch

        // heap TIDs on the device
        if (nb_new > 0 || !dirty.empty()) {
            std::vector<int64_t> t0(nb_new);
            std::vector<uint8_t> nt(nb_new, 1);
            for (int64_t e = 0; e < nb_new; e++) t0[e] = ix->h_tids[(size_t) (cur + e) * HB_HEAPTIDS];
            if (nb_new > 0) {
                HB_CK(cudaMemcpy(ix->d_tid0 + cur, t0.data(), sizeof(int64_t) * nb_new, cudaMemcpyHostToDevice));
                HB_CK(cudaMemcpy(ix->d_ntids + cur, nt.data(), nb_new, cudaMemcpyHostToDevice));
            }
            if (!dirty.empty()) {
                if (!ix->d_tidx) {
                    HB_CK(cudaMalloc(&ix->d_tidx, sizeof(int64_t) * ix->cap * (HB_HEAPTIDS - 1)));
                    HB_CK(cudaMemset(ix->d_tidx, 0, sizeof(int64_t) * ix->cap * (HB_HEAPTIDS - 1)));
                    ix->has_dups = true;
                }
                for (int32_t c : dirty) {
                    HB_CK(cudaMemcpy(ix->d_tidx + (size_t) c * (HB_HEAPTIDS - 1), &ix->h_tids[(size_t) c * HB_HEAPTIDS + 1],
                                     sizeof(int64_t) * (HB_HEAPTIDS - 1), cudaMemcpyHostToDevice));
                    HB_CK(cudaMemcpy(ix->d_ntids + c, &ix->h_ntids[c], 1, cudaMemcpyHostToDevice));
                }
            }
        }
        pos += b;
        indexed += b;
    }
    return indexed;
}

}   // namespace hb

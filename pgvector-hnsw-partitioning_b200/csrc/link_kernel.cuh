// link_kernel.cuh -- HnswUpdateConnection for a whole batch: the reverse links of the new elements.
//
// Role [RECALL; the reference mount has no source, /root/reference/README.md:1]: hnswutils.c
// HnswUpdateConnection.  Appending to a list with room is free; a full list is shrunk by
// SelectNeighbors over list + {new element} sorted by (distance, id), and the element it prunes is
// replaced in place.  In steady state nearly every list is full, so this is where a build spends
// its distance evaluations: ~200 CheckElementCloser pairs per reverse link, 32 links per insert.
//
// Input: the batch's edges sorted by key = layer | target | source (build.cu generates and sorts
// them on the device) and the list of segment heads (first edge of each distinct (layer, target));
// a segment's edges are applied in ascending source order, segments are independent.
//
// Two kernels, same results:
//   link_pipe_kernel  one persistent CTA per SM, warp-specialised.  A producer warp reads the lists
//                     of 32 segments ahead, applies appends itself, and for a full list stages the
//                     lm+1 candidate rows ONCE into a shared-memory stage by 1-D bulk async copies
//                     (cp.async.bulk -> mbarrier; 2 stages at 3 kB rows, up to 8 for short rows).
//                     Twelve consumer warps evaluate the distance matrix of the stage's rows from
//                     shared memory in 4x4 (fp32) / 2x4 (fp16) register tiles in the canonical
//                     summation order (tiles handed out by an atomic counter); the warp that
//                     finishes the last tile replays the sequential selection on the matrix with
//                     bit masks, and handles the segment's further edges incrementally: the matrix
//                     is kept by list slot, so another new element costs one row and lm pairs.
//                     HBM traffic per segment = lm+1 rows (+1 per extra edge) instead of ~125 rows
//                     per edge (pairs re-fetched per CheckElementCloser).
//   link_warp_kernel  one warp per segment, pairs fetched from HBM as CheckElementCloser asks for
//                     them; used when a stage does not fit shared memory (very wide rows) or m > 32.
// n_pair is counted as the sequential algorithm evaluates pairs (stop at the first selected
// neighbour that is at least as close), so both kernels and the oracle report the same figure.
#pragma once
#include "scan_kernel.cuh"
#include <cuda_runtime.h>

namespace hb {

constexpr int LINK_KEY_SRC_BITS = 16;   // key = layer << 48 | target << 16 | (source - first)
constexpr unsigned long long LINK_KEY_INVALID = ~0ull;

struct LinkParams {
    GraphView g;
    int64_t first;                 // id of the batch's first new element (sources are first + low key bits)
    int E;                         // edge slots (valid edges sort first)
    const int32_t *nseg;           // device: number of segments
    const int32_t *seg_start;      // per segment: index of its first edge
    const unsigned long long *edge_key;   // E, sorted
    const float *edge_d;           // E, distance source <-> target
    int32_t *nbr0; float *nbr0d; int32_t *nbru; float *nbrud;
    unsigned long long *totals;
    const int32_t *flag;           // device: non-zero = the batch is being redone by the host path, do nothing
    // pair cache (link_memo_kernel): per list, the distances among its members by slot, strict lower
    // triangle, and a byte saying whether it has been filled.  NULL when the cache is not allocated.
    float *pc0; uint8_t *pv0;      // n x lm0(lm0-1)/2, n
    float *pcu; uint8_t *pvu;      // upper_rows x m(m-1)/2, upper_rows
    // fill mode of link_pipe_kernel: instead of processing segments, fill the pair cache of the
    // lists named here (key = layer << 32 | target; full lists whose triangle is not filled yet)
    const unsigned long long *fill_list;
    const int32_t *nfill;
    // evaluated-distance tables of the batch's candidate searches (search_core.cuh EvalTable): table
    // of new element first + i at et_key + i * et_slots.  NULL = not available (distances are computed)
    const uint32_t *et_key; const float *et_val; int et_slots;
};

__device__ __forceinline__ uint32_t link_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

// ---- register-tiled pair distances from shared memory ------------------------------------------
template <typename T> __device__ __forceinline__ void chunk_to_float(const uint4 &raw, float (&v)[Vec<T>::VEC])
{
    if constexpr (sizeof(T) == 4) {
        v[0] = __uint_as_float(raw.x); v[1] = __uint_as_float(raw.y);
        v[2] = __uint_as_float(raw.z); v[3] = __uint_as_float(raw.w);
    } else {
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float2 h1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
        const float2 h2 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.z));
        const float2 h3 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.w));
        v[0] = h0.x; v[1] = h0.y; v[2] = h1.x; v[3] = h1.y;
        v[4] = h2.x; v[5] = h2.y; v[6] = h3.x; v[7] = h3.y;
    }
}

// distances of the TA x TB pairs (ia[x], ib[y]) of rows resident in shared memory; pair c = x*TB+y
// is returned in lane c * (32 / (TA*TB)).  Same additions as group_distance (distance.cuh).
template <typename T, int IP, int TA, int TB>
__device__ __forceinline__ float pair_tile(const char *rows, uint32_t row_bytes, int nvec, const int (&ia)[TA],
                                           const int (&ib)[TB], int lane)
{
    constexpr int VEC = Vec<T>::VEC;
    constexpr int NP = TA * TB;
    float acc[NP][VEC];
#pragma unroll
    for (int c = 0; c < NP; c++)
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[c][k] = 0.0f;
    const char *pa[TA], *pb[TB];
#pragma unroll
    for (int x = 0; x < TA; x++) pa[x] = rows + (size_t) ia[x] * row_bytes + 16 * lane;
#pragma unroll
    for (int y = 0; y < TB; y++) pb[y] = rows + (size_t) ib[y] * row_bytes + 16 * lane;
    for (int off = 0; off + lane < nvec; off += 32) {
        float a[TA][VEC], b[TB][VEC];
#pragma unroll
        for (int x = 0; x < TA; x++) chunk_to_float<T>(*reinterpret_cast<const uint4 *>(pa[x] + 16 * off), a[x]);
#pragma unroll
        for (int y = 0; y < TB; y++) chunk_to_float<T>(*reinterpret_cast<const uint4 *>(pb[y] + 16 * off), b[y]);
#pragma unroll
        for (int x = 0; x < TA; x++)
#pragma unroll
            for (int y = 0; y < TB; y++)
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    if constexpr (IP == 1) acc[x * TB + y][k] = fmaf(a[x][k], b[y][k], acc[x * TB + y][k]);
                    else if constexpr (IP == 2) acc[x * TB + y][k] = acc[x * TB + y][k] + fabsf(a[x][k] - b[y][k]);
                    else {
                        const float t = a[x][k] - b[y][k];
                        acc[x * TB + y][k] = fmaf(t, t, acc[x * TB + y][k]);
                    }
                }
    }
    float part[NP];
#pragma unroll
    for (int c = 0; c < NP; c++) part[c] = fold_lane<VEC>(acc[c]);
    const float s = XReduce<NP>::run(part, lane, 16);
    return IP == 1 ? -s : s;
}

// -DHB_LINK_PROFILE: cycle accounting of the pipeline roles into totals[6..13] (experiments only)
#ifdef HB_LINK_PROFILE
#define LINK_CLK() clock64()
#else
#define LINK_CLK() 0ll
#endif

constexpr int LINK_CW = 12;                          // consumer warps
constexpr int LINK_THREADS = (LINK_CW + 1) * 32;     // + the producer warp
constexpr int LINK_MAX_STAGES = 8;
constexpr int LINK_PREFETCH = 32;                    // segments whose lists the producer reads ahead
constexpr int LINK_MAX_LM = 63;                      // lm + 1 candidates: two per lane, 64-bit selection masks
constexpr int LINK_MAX_TILES = 352;
constexpr int LINK_TB = 4;

struct LinkStageMeta {
    int kind;                    // 1 = segment, 2 = fill the pair cache of a list, 0 = no more work
    int lm;                      // list capacity of this layer; candidates = lm + 1
    int e;                       // index of the edge whose new element sits in slot lm
    int ntiles;
    int tile_next, tiles_done;   // tiles are handed out dynamically; the last finisher finalises
    int table;                   // which tile table (0: layer 0, 1: upper layers)
    int pad;
    unsigned long long seg_key;  // key >> LINK_KEY_SRC_BITS of the segment
    int32_t *gl;                 // the list in HBM
    float *gld;
    float *pc;                   // fill mode (kind 2): where the triangle goes
    uint8_t *pv;
};

template <typename T> struct LinkTile { static constexpr int TA = sizeof(T) == 4 ? 4 : 2; };

// column blocks a row block needs to cover the strict lower triangle of lm x lm
__host__ __device__ inline int link_tile_cols(int rb, int TA, int lm)
{
    int imax = rb * TA + TA - 1;
    if (imax > lm - 1) imax = lm - 1;
    return (imax + LINK_TB - 1) / LINK_TB;
}
__host__ __device__ inline int link_num_tiles(int TA, int lm)
{
    int n = 0;
    for (int rb = 0; rb * TA < lm; rb++) n += link_tile_cols(rb, TA, lm);
    return n + (lm + LINK_TB - 1) / LINK_TB;        // + the new element's row
}

__host__ __device__ inline size_t link_stage_bytes(size_t row_bytes, int lm0)
{
    const int cap = lm0 + 1;
    size_t b = (size_t) cap * row_bytes;            // candidate rows, by list slot; slot lm = the new element
    b += (size_t) cap * cap * 4;                    // distance matrix, symmetric, by slot
    b += (size_t) cap * 8;                          // l_id, l_d
    b += (size_t) ((cap + 7) & ~7);                 // ord: sorted position -> slot
    b = (b + 15) & ~(size_t) 15;
    b += sizeof(LinkStageMeta) + 16;                // meta, full + empty mbarriers
    return (b + 127) & ~(size_t) 127;
}
__host__ __device__ inline size_t link_shared_bytes(int lm0)
{
    return (size_t) LINK_PREFETCH * (lm0 + 1) * 8 + 2 * LINK_MAX_TILES * 2 + 64;
}

__device__ __forceinline__ void link_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "LINK_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra LINK_DONE;\n\t"
        "bra LINK_WAIT;\n\t"
        "LINK_DONE:\n\t"
        "}" ::"r"(link_smem_u32(bar)), "r"(parity) : "memory");
}

struct LinkStage {
    char *rows; float *D; int32_t *l_id; float *l_d; uint8_t *ord; LinkStageMeta *meta; uint64_t *full, *empty;
};
__device__ __forceinline__ LinkStage link_stage_at(unsigned char *base, size_t row_bytes, int lm0)
{
    const int cap = lm0 + 1;
    LinkStage st;
    st.rows = reinterpret_cast<char *>(base);
    st.D = reinterpret_cast<float *>(base + (size_t) cap * row_bytes);
    st.l_id = reinterpret_cast<int32_t *>(st.D + cap * cap);
    st.l_d = reinterpret_cast<float *>(st.l_id + cap);
    st.ord = reinterpret_cast<uint8_t *>(st.l_d + cap);
    size_t off = (size_t) cap * row_bytes + (size_t) cap * cap * 4 + (size_t) cap * 8 + ((cap + 7) & ~7);
    off = (off + 15) & ~(size_t) 15;
    st.meta = reinterpret_cast<LinkStageMeta *>(base + off);
    st.full = reinterpret_cast<uint64_t *>(base + off + sizeof(LinkStageMeta));
    st.empty = st.full + 1;
    return st;
}

// distances new element (slot lm) <-> list slots [cb*4, cb*4+4), written to both halves of D
template <typename T, int IP>
__device__ __forceinline__ void link_new_row_tile(const LinkStage &st, uint32_t row_bytes, int nvec, int lm, int ld,
                                                  int cb, int lane)
{
    const int ia[1] = { lm };
    int ib[LINK_TB];
#pragma unroll
    for (int y = 0; y < LINK_TB; y++) ib[y] = min(cb * LINK_TB + y, lm - 1);
    const float s = pair_tile<T, IP, 1, LINK_TB>(st.rows, row_bytes, nvec, ia, ib, lane);
    const int j = cb * LINK_TB + lane / (32 / LINK_TB);
    if ((lane % (32 / LINK_TB)) == 0 && j < lm) { st.D[lm * ld + j] = s; st.D[j * ld + lm] = s; }
}

// One warp: the candidates are the list slots 0..lm-1 plus the new element in slot lm, with distances
// to the owner l_d and pairwise distances D (symmetric, leading dimension ld).  Sorts them by
// (distance, id) [sortCandidates = true], replays SelectNeighbors and returns the SLOT of the
// candidate upstream reports as pruned.  n_pair advances by what the sequential loop evaluates.
__device__ __forceinline__ int link_select_slot(const float *D, int ld, const int32_t *l_id, const float *l_d,
                                                uint8_t *ord, int lm, int lane, unsigned long long &npair)
{
    const int nc = lm + 1;
    const bool has1 = lane + 32 < nc;
    // candidates in (distance, id) order [sortCandidates = true]
    const bool has0 = lane < nc;
    const float d0 = has0 ? l_d[lane] : 0.f;
    const int32_t i0 = has0 ? l_id[lane] : 0;
    const float d1 = has1 ? l_d[lane + 32] : 0.f;
    const int32_t i1 = has1 ? l_id[lane + 32] : 0;
    int rank0 = 0, rank1 = 0;
    for (int b = 0; b < nc; b++) {
        const float db = __shfl_sync(FULL, b < 32 ? d0 : d1, b & 31);
        const int32_t ib = __shfl_sync(FULL, b < 32 ? i0 : i1, b & 31);
        rank0 += (db < d0 || (db == d0 && ib < i0)) ? 1 : 0;
        rank1 += (db < d1 || (db == d1 && ib < i1)) ? 1 : 0;
    }
    if (has0) ord[rank0] = (uint8_t) lane;
    if (has1) ord[rank1] = (uint8_t) (lane + 32);
    __syncwarp();
    // lane j takes sorted positions j and j+32: slot and distance to the owner
    const int s0 = has0 ? ord[lane] : 0, s1 = has1 ? ord[lane + 32] : 0;
    const float e0 = l_d[s0], e1 = l_d[s1];
    // SelectNeighbors.  Candidate i is kept iff no kept candidate j < i is at least as close to it
    // as the owner is: mk = ballot over j of D[i][j] <= d_i, chained through `sel`.  The loads and
    // ballots do not depend on the chain, so they run ahead of it.  With lm + 1 candidates the
    // loop's other exit (lm kept) can only trigger at the last candidate with nothing rejected,
    // and the pruned candidate is the last rejected one, or the last candidate if none was.
    unsigned long long sel = 0ull;
    int pairs0 = 0, pairs1 = 0;                // this lane's positions: pairs the sequential loop evaluates
#pragma unroll 4
    for (int i = 0; i < nc; i++) {
        const int si = __shfl_sync(FULL, i < 32 ? s0 : s1, i & 31);
        const float di = __shfl_sync(FULL, i < 32 ? e0 : e1, i & 31);
        const float v0 = D[si * ld + s0];
        const unsigned lo = __ballot_sync(FULL, lane < i && v0 <= di);
        unsigned hi = 0u;
        if (nc > 33) {                     // position 32 can only fail candidates after it: none when nc = 33
            const float v1 = D[si * ld + s1];
            hi = __ballot_sync(FULL, lane + 32 < i && v1 <= di);
        }
        const unsigned long long f = ((unsigned long long) hi << 32 | lo) & sel;
        const int before = __popcll(sel);
        if (before < lm) {
            int ev;
            if (f == 0ull) { sel |= 1ull << i; ev = before; }
            else ev = __popcll(sel & ((2ull << (__ffsll((long long) f) - 1)) - 1ull));
            if ((i & 31) == lane) { if (i < 32) pairs0 = ev; else pairs1 = ev; }
        }
    }
    {
        int ev = pairs0 + pairs1;
        for (int b = 16; b >= 1; b >>= 1) ev += __shfl_xor_sync(FULL, ev, b);
        npair += ev;
    }
    const unsigned long long all = nc >= 64 ? ~0ull : (1ull << nc) - 1ull;
    const unsigned long long rej = all & ~sel;
    const int nkept = __popcll(sel);
    int pruned_pos = nc - 1;
    if (rej != 0ull && !(nkept == lm && rej == 1ull << (nc - 1))) pruned_pos = 63 - __clzll((long long) rej);
    return __shfl_sync(FULL, pruned_pos < 32 ? s0 : s1, pruned_pos & 31);
}

// One warp: the sequential SelectNeighbors replayed on the stage's distance matrix for the edge in
// slot lm, the replacement of the pruned element, then the segment's remaining edges one by one.
// The shared-memory pipe is saturated by the tile loads of the other warps, so this avoids chains
// of dependent shared-memory reads: lane l owns candidate slots l and l+32 in registers, ranks come
// from shuffles, matrix rows are read with independent loads.
template <typename T, int IP>
__device__ __forceinline__ void link_finalize(const LinkParams &p, const LinkStage &st, int lane,
                                              unsigned long long &npair)
{
    const GraphView &g = p.g;
    const uint32_t row_bytes = (uint32_t) g.row_bytes;
    const int lm = st.meta->lm, ld = 2 * g.m + 1;
    const unsigned long long seg_key = st.meta->seg_key;
    int e = st.meta->e;
    long long t_extra = 0;
    for (;;) {
        // the key of the following edge is needed only at the end of the iteration: ask for it now
        const unsigned long long next_key = e + 1 < p.E ? __ldcg(p.edge_key + e + 1) : LINK_KEY_INVALID;
        const float next_d = e + 1 < p.E ? __ldcg(p.edge_d + e + 1) : 0.f;
        const int ps = link_select_slot(st.D, ld, st.l_id, st.l_d, st.ord, lm, lane, npair);
        if (ps != lm) {
            // the new element takes the pruned element's slot: list entry, matrix row/column, row
            for (int b = lane; b < lm; b += 32) {
                if (b != ps) { const float v = st.D[lm * ld + b]; st.D[ps * ld + b] = v; st.D[b * ld + ps] = v; }
            }
            if (lane == 0) { st.l_id[ps] = st.l_id[lm]; st.l_d[ps] = st.l_d[lm]; }
            const uint4 *from = reinterpret_cast<const uint4 *>(st.rows + (size_t) lm * row_bytes);
            uint4 *to = reinterpret_cast<uint4 *>(st.rows + (size_t) ps * row_bytes);
            for (int ch = lane; ch < g.nvec; ch += 32) to[ch] = from[ch];
        }
        __syncwarp();
        // the segment's next edge: the list stays full, so it is another shrink
        e++;
        if (e >= p.E || (next_key >> LINK_KEY_SRC_BITS) != seg_key) break;
        const long long tx0 = LINK_CLK();
        const int32_t src = (int32_t) (p.first + (int64_t) (next_key & ((1u << LINK_KEY_SRC_BITS) - 1)));
        if (lane == 0) { st.l_id[lm] = src; st.l_d[lm] = next_d; }
        {
            const char *from = g.vecs + (size_t) src * row_bytes;
            uint4 *to = reinterpret_cast<uint4 *>(st.rows + (size_t) lm * row_bytes);
            for (int ch = lane; ch < g.nvec; ch += 32) to[ch] = ldg_stream(from + 16 * ch);
        }
        __syncwarp();
        for (int cb = 0; cb * LINK_TB < lm; cb++) link_new_row_tile<T, IP>(st, row_bytes, g.nvec, lm, ld, cb, lane);
        __syncwarp();
        t_extra += LINK_CLK() - tx0;
    }
#ifdef HB_LINK_PROFILE
    if (lane == 0) {
        atomicAdd(p.totals + 9, (unsigned long long) t_extra);
    }
#endif
    int32_t *gl = st.meta->gl;
    float *gld = st.meta->gld;
    for (int a = lane; a < lm; a += 32) { gl[a] = st.l_id[a]; gld[a] = st.l_d[a]; }
}

template <typename T, int IP>
__global__ void __launch_bounds__(LINK_THREADS, 1) link_pipe_kernel(const LinkParams p, const int nstages)
{
    constexpr int TA = LinkTile<T>::TA, TB = LINK_TB, NP = TA * TB;
    extern __shared__ __align__(16) unsigned char smem[];
    if (*p.flag) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GraphView &g = p.g;
    const int lm0 = 2 * g.m, cap = lm0 + 1, ld = cap;
    const uint32_t row_bytes = (uint32_t) g.row_bytes;
    const size_t stage_bytes = link_stage_bytes(row_bytes, lm0);
    unsigned char *shared = smem + stage_bytes * nstages;
    int32_t *cache_id = reinterpret_cast<int32_t *>(shared);                       // LINK_PREFETCH x cap
    float *cache_d = reinterpret_cast<float *>(cache_id + LINK_PREFETCH * cap);    // LINK_PREFETCH x cap
    uint16_t *tiles = reinterpret_cast<uint16_t *>(cache_d + LINK_PREFETCH * cap); // 2 x LINK_MAX_TILES

    // tile tables: big tiles (rb << 8 | cb) over the list's lower triangle, then the new element's row (0x8000 | cb)
    for (int tb = 0; tb < 2; tb++) {
        const int lm = tb == 0 ? lm0 : g.m;
        const int nbig = link_num_tiles(TA, lm) - (lm + TB - 1) / TB;
        for (int t = tid; t < link_num_tiles(TA, lm); t += LINK_THREADS) {
            uint16_t v;
            if (t >= nbig) v = (uint16_t) (0x8000 | (t - nbig));
            else {
                int rb = 0, rest = t;
                for (;; rb++) {
                    const int c = link_tile_cols(rb, TA, lm);
                    if (rest < c) break;
                    rest -= c;
                }
                v = (uint16_t) (rb << 8 | rest);
            }
            tiles[tb * LINK_MAX_TILES + t] = v;
        }
    }
    if (tid == 0) {
        for (int s = 0; s < nstages; s++) {
            const LinkStage st = link_stage_at(smem + stage_bytes * s, row_bytes, lm0);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(link_smem_u32(st.full)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(link_smem_u32(st.empty)), "r"(LINK_CW));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool fill = p.fill_list != nullptr;
    const int S = fill ? *p.nfill : *p.nseg;

    if (warp == LINK_CW) {
        // ================= producer =================
        int stage = 0;
        uint32_t phase = 0;
        long long tp_wait = 0;
        const long long tp0 = LINK_CLK();
        (void) tp0;
        for (int base = blockIdx.x * LINK_PREFETCH; base < S; base += gridDim.x * LINK_PREFETCH) {
            // lane k reads ahead for segment base + k: key, list location, the list itself
            const int seg = base + lane;
            int e0 = 0, lm = lm0;
            unsigned long long key0 = 0ull;
            int32_t *gl = nullptr;
            float *gld = nullptr;
            float *pc = nullptr;
            uint8_t *pv = nullptr;
            if (seg < S) {
                if (fill) key0 = p.fill_list[seg];
                else {
                    e0 = p.seg_start[seg];
                    key0 = p.edge_key[e0] >> LINK_KEY_SRC_BITS;
                }
                const int lc = (int) (key0 >> 32);
                const int32_t target = (int32_t) (key0 & 0xffffffffu);
                if (lc == 0) {
                    gl = p.nbr0 + (size_t) target * lm0; gld = p.nbr0d + (size_t) target * lm0;
                    if (fill) { pc = p.pc0 + (size_t) target * (lm0 * (lm0 - 1) / 2); pv = p.pv0 + target; }
                } else {
                    lm = g.m;
                    const size_t row = (size_t) g.uoff[target] + (lc - 1);
                    gl = p.nbru + row * g.m; gld = p.nbrud + row * g.m;
                    if (fill) { pc = p.pcu + row * (g.m * (g.m - 1) / 2); pv = p.pvu + row; }
                }
                int32_t *ci = cache_id + lane * cap;
                float *cd = cache_d + lane * cap;
                if ((lm & 3) == 0) {
                    for (int j = 0; j < lm; j += 4) {
                        const int4 v = __ldcg(reinterpret_cast<const int4 *>(gl + j));
                        const float4 w = __ldcg(reinterpret_cast<const float4 *>(gld + j));
                        ci[j] = v.x; ci[j + 1] = v.y; ci[j + 2] = v.z; ci[j + 3] = v.w;
                        cd[j] = w.x; cd[j + 1] = w.y; cd[j + 2] = w.z; cd[j + 3] = w.w;
                    }
                } else {
                    for (int j = 0; j < lm; j++) { ci[j] = __ldcg(gl + j); cd[j] = __ldcg(gld + j); }
                }
            }
            __syncwarp();
            const int nk = min(LINK_PREFETCH, S - base);
            for (int k = 0; k < nk; k++) {
                const int lm_k = __shfl_sync(FULL, lm, k);
                const unsigned long long key_k = __shfl_sync(FULL, key0, k);
                int e = __shfl_sync(FULL, e0, k);
                int32_t *gl_k = reinterpret_cast<int32_t *>(__shfl_sync(FULL, (unsigned long long) gl, k));
                float *gld_k = reinterpret_cast<float *>(__shfl_sync(FULL, (unsigned long long) gld, k));
                int32_t *ci = cache_id + k * cap;
                float *cd = cache_d + k * cap;
                int cnt = 0;
                for (int jb = 0; jb < lm_k; jb += 32) cnt += __popc(__ballot_sync(FULL, jb + lane < lm_k && ci[jb + lane] >= 0));
                if (fill) {
                    // stage the members' rows; the consumers compute their triangle
                    float *pc_k = reinterpret_cast<float *>(__shfl_sync(FULL, (unsigned long long) pc, k));
                    uint8_t *pv_k = reinterpret_cast<uint8_t *>(__shfl_sync(FULL, (unsigned long long) pv, k));
                    if (cnt < lm_k) continue;
                    const LinkStage st = link_stage_at(smem + stage_bytes * stage, row_bytes, lm0);
                    link_mbar_wait(st.empty, phase ^ 1u);
                    if (lane == 0) {
                        LinkStageMeta *mt = st.meta;
                        mt->kind = 2; mt->lm = lm_k; mt->e = 0; mt->table = lm_k == lm0 ? 0 : 1;
                        mt->ntiles = link_num_tiles(TA, lm_k) - (lm_k + TB - 1) / TB; mt->tile_next = 0; mt->tiles_done = 0;
                        mt->seg_key = key_k; mt->gl = gl_k; mt->gld = gld_k; mt->pc = pc_k; mt->pv = pv_k;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                     ::"r"(link_smem_u32(st.full)), "r"((uint32_t) lm_k * row_bytes) : "memory");
                    }
                    __syncwarp();
                    for (int a = lane; a < lm_k; a += 32) {
                        const char *srcp = g.vecs + (size_t) ci[a] * row_bytes;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(link_smem_u32(st.rows + (size_t) a * row_bytes)), "l"(srcp), "r"(row_bytes),
                                       "r"(link_smem_u32(st.full)) : "memory");
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                    continue;
                }
                bool handed = false;
                for (; e < p.E; e++) {
                    const unsigned long long key = p.edge_key[e];
                    if ((key >> LINK_KEY_SRC_BITS) != key_k) break;
                    const int32_t src = (int32_t) (p.first + (int64_t) (key & ((1u << LINK_KEY_SRC_BITS) - 1)));
                    const float d = p.edge_d[e];
                    __syncwarp();
                    if (lane == 0) { ci[cnt] = src; cd[cnt] = d; }      // slot cnt <= lm: append, or the new element of a shrink
                    __syncwarp();
                    if (cnt < lm_k) { cnt++; continue; }
                    // full list: stage it for the consumers
                    const LinkStage st = link_stage_at(smem + stage_bytes * stage, row_bytes, lm0);
                    const long long tw0 = LINK_CLK();
                    link_mbar_wait(st.empty, phase ^ 1u);
                    tp_wait += LINK_CLK() - tw0;
                    for (int a = lane; a <= lm_k; a += 32) { st.l_id[a] = ci[a]; st.l_d[a] = cd[a]; }
                    if (lane == 0) {
                        LinkStageMeta *mt = st.meta;
                        mt->kind = 1; mt->lm = lm_k; mt->e = e; mt->table = lm_k == lm0 ? 0 : 1;
                        mt->ntiles = link_num_tiles(TA, lm_k); mt->tile_next = 0; mt->tiles_done = 0;
                        mt->seg_key = key_k; mt->gl = gl_k; mt->gld = gld_k;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                     ::"r"(link_smem_u32(st.full)), "r"((uint32_t) (lm_k + 1) * row_bytes) : "memory");
                    }
                    __syncwarp();
                    for (int a = lane; a <= lm_k; a += 32) {
                        const char *srcp = g.vecs + (size_t) ci[a] * row_bytes;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(link_smem_u32(st.rows + (size_t) a * row_bytes)), "l"(srcp), "r"(row_bytes),
                                       "r"(link_smem_u32(st.full)) : "memory");
                    }
                    if (++stage == nstages) { stage = 0; phase ^= 1u; }
                    handed = true;
                    break;
                }
                if (!handed) {
                    for (int a = lane; a < cnt; a += 32) { gl_k[a] = ci[a]; gld_k[a] = cd[a]; }
                }
                __syncwarp();
            }
        }
        // no more segments
        const LinkStage st = link_stage_at(smem + stage_bytes * stage, row_bytes, lm0);
        link_mbar_wait(st.empty, phase ^ 1u);
        if (lane == 0) {
            st.meta->kind = 0;
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(link_smem_u32(st.full)) : "memory");
#ifdef HB_LINK_PROFILE
            atomicAdd(p.totals + 10, (unsigned long long) tp_wait);
            atomicAdd(p.totals + 11, (unsigned long long) (LINK_CLK() - tp0));
#endif
        }
        return;
    }

    // ================= consumers =================
    unsigned long long npair = 0;
    int stage = 0;
    uint32_t phase = 0;
    long long tc_wait = 0, tc_tiles = 0, tc_final = 0, n_final = 0;
    const long long tc0 = LINK_CLK();
    (void) tc0;
    for (;;) {
        const LinkStage st = link_stage_at(smem + stage_bytes * stage, row_bytes, lm0);
        const long long tw0 = LINK_CLK();
        link_mbar_wait(st.full, phase);
        const long long tw1 = LINK_CLK();
        tc_wait += tw1 - tw0;
        if (st.meta->kind == 0) break;
        const int lm = st.meta->lm, ntiles = st.meta->ntiles;
        const uint16_t *tab = tiles + st.meta->table * LINK_MAX_TILES;
        bool last = false;
        for (;;) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&st.meta->tile_next, 1);
            t = __shfl_sync(FULL, t, 0);
            if (t >= ntiles) break;
            const int code = tab[t];
            if (code & 0x8000) link_new_row_tile<T, IP>(st, row_bytes, g.nvec, lm, ld, code & 0xff, lane);
            else {
                const int rb = code >> 8, cb = code & 0xff;
                int ia[TA], ib[TB];
#pragma unroll
                for (int x = 0; x < TA; x++) ia[x] = min(rb * TA + x, lm - 1);
#pragma unroll
                for (int y = 0; y < TB; y++) ib[y] = min(cb * TB + y, lm - 1);
                const float s = pair_tile<T, IP, TA, TB>(st.rows, row_bytes, g.nvec, ia, ib, lane);
                const int c = lane / (32 / NP);
                const int i = rb * TA + c / TB, j = cb * TB + c % TB;
                if ((lane % (32 / NP)) == 0 && i < lm && j < i) { st.D[i * ld + j] = s; st.D[j * ld + i] = s; }
            }
            __syncwarp();
            int done = 0;
            if (lane == 0) { __threadfence_block(); done = atomicAdd(&st.meta->tiles_done, 1) + 1; }
            done = __shfl_sync(FULL, done, 0);
            if (done == ntiles) { last = true; break; }
        }
        const long long tw2 = LINK_CLK();
        tc_tiles += tw2 - tw1;
        if (last) {
            __threadfence_block();
            if (st.meta->kind == 2) {
                float *pc = st.meta->pc;
                for (int a = 1; a < lm; a++)
                    for (int b = lane; b < a; b += 32) __stcs(pc + a * (a - 1) / 2 + b, st.D[a * ld + b]);
                if (lane == 0) *st.meta->pv = 1;
            } else link_finalize<T, IP>(p, st, lane, npair);
            __syncwarp();
            tc_final += LINK_CLK() - tw2;
            n_final++;
        }
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(link_smem_u32(st.empty)) : "memory");
        if (++stage == nstages) { stage = 0; phase ^= 1u; }
    }
    if (lane == 0 && npair) atomicAdd(p.totals + 4, npair);
#ifdef HB_LINK_PROFILE
    if (lane == 0) {
        atomicAdd(p.totals + 6, (unsigned long long) tc_wait);
        atomicAdd(p.totals + 7, (unsigned long long) tc_tiles);
        atomicAdd(p.totals + 8, (unsigned long long) tc_final);
        atomicAdd(p.totals + 12, (unsigned long long) (LINK_CLK() - tc0));
        atomicAdd(p.totals + 13, (unsigned long long) n_final);
    }
#endif
}

}   // namespace hb

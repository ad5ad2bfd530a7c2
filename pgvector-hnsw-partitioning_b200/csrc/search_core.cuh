// search_core.cuh -- one warp runs one query's HnswSearchLayer beam search.
//
// Takes the role of hnswutils.c HnswSearchLayer [RECALL / arXiv:1603.09320 Alg. 2; the reference
// mount has no source, /root/reference/README.md:1].  The CPU formulation keeps two heaps, C
// (candidates, nearest first) and W (results, furthest first), and a visited hash.  Here:
//   * W and C are ONE array in shared memory sorted by the key (distance, id), each entry
//     carrying an "expanded" bit.  C = the unexpanded entries.  The array keeps ef entries plus
//     the run of entries whose distance EQUALS entry ef-1 (the only evicted candidates that the
//     CPU loop could still expand: `c->distance > f->distance` is false for them), so the
//     expansion sequence -- and therefore every id returned -- is identical to the two-heap loop
//     with (distance, id) tie order (oracle/hnsw_oracle.c search_layer).
//   * the visited set is an open-addressing hash table in shared memory (exact: when it would
//     exceed its load limit the query is handed to the large-visited-set path, a bitmap in HBM).
//   * per expansion the warp reads the neighbour list with one coalesced load, filters it through
//     the visited set, evaluates the new candidates G rows at a time (distance.cuh) and inserts the
//     admitted ones in neighbour order, re-testing `d < furthest` before each insertion exactly as
//     the sequential loop does.
#pragma once
#include "distance.cuh"

namespace hb {

constexpr uint32_t EXP_BIT = 0x80000000u;
constexpr uint32_t ID_MASK = 0x7fffffffu;
constexpr uint32_t EMPTY = 0xffffffffu;

enum { ST_OK = 0, ST_TABLE = 1, ST_TAIL = 2 };

struct GraphView {
    const char *vecs;        // n rows, row_bytes each (16-byte multiple, zero padded)
    size_t row_bytes;
    int dim, nvec;           // nvec = 16-byte chunks per row
    const int32_t *nbr0;     // n x 2m, -1 padded
    const int32_t *uoff;     // n: first row of the element in nbru, or -1
    const int32_t *nbru;     // upper_rows x m
    int m;
    int32_t entry;
    int entry_level;
    int64_t n;
};

struct WList {               // sorted (distance, id) array; warp-uniform bookkeeping
    float *d;
    uint32_t *id;
    int cap, L;
};

// ---- visited sets ---------------------------------------------------------------------------
struct VisitedHash {         // shared-memory table + per-warp overflow table in HBM (both exact)
    uint32_t *tab;           // shared memory, open addressing, linear probing
    int slots, count, limit; // any multiple of 4 slots: the home slot is mulhi(hash, slots), not a mask
    uint32_t *otab;          // overflow: used only once the shared table reached its load limit
    int oslots, ocount, olimit;
    bool odirty;
    __device__ __forceinline__ void configure(int s)
    {
        slots = s;
        limit = s - (s >> 2);             // 75 % load
    }
    __device__ __forceinline__ void set_overflow(uint32_t *t, int s)
    {
        otab = t; oslots = s; odirty = true;   // first clear() wipes it
        olimit = s - (s >> 2);
        ocount = 0;
    }
    __device__ __forceinline__ void clear(int lane)
    {
        uint4 *t = reinterpret_cast<uint4 *>(tab);
        for (int i = lane; i < slots / 4; i += 32) t[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        count = 0;
        if (odirty) {
            uint4 *o = reinterpret_cast<uint4 *>(otab);
            for (int i = lane; i < oslots / 4; i += 32) o[i] = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
            odirty = false;
            ocount = 0;
        }
        __syncwarp();
    }
    // room for `incoming` more keys?  (warp-uniform; mirrors spill(): once a batch went to the overflow
    // table every later one does, so from then on only the overflow table's load limit counts)
    __device__ __forceinline__ bool room(int incoming) const
    {
        return ocount > 0 ? ocount + incoming <= olimit : (count + incoming <= limit || incoming <= olimit);
    }
    // warp-uniform: where does this batch of `incoming` keys go
    // (sticky: once a batch spilled, every later batch of this layer search goes to the overflow
    // table, so a key lives in exactly one of the two tables)
    __device__ __forceinline__ bool spill(int incoming) const { return ocount > 0 || count + incoming > limit; }

    __device__ __forceinline__ static bool probe_insert(uint32_t *t, uint32_t nslots, uint32_t h, uint32_t key)
    {
        // most probes find the key already present (neighbour lists overlap heavily): a plain
        // load answers those; only an empty slot needs the CAS.  Slots only ever go EMPTY -> key.
        for (;;) {
            const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(t + h);
            if (cur == key) return false;
            if (cur == EMPTY) {
                const uint32_t old = atomicCAS(t + h, EMPTY, key);
                if (old == EMPTY) return true;
                if (old == key) return false;
            }
            h = h + 1 == nslots ? 0u : h + 1;
        }
    }
    __device__ __forceinline__ bool contains(uint32_t key) const
    {
        uint32_t h = __umulhi(key * 0x9E3779B1u, (uint32_t) slots);
        for (;;) {
            const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(tab + h);
            if (cur == key) return true;
            if (cur == EMPTY) return false;
            h = h + 1 == (uint32_t) slots ? 0u : h + 1;
        }
    }
    // insert one key; `to_overflow` is the warp-uniform decision of spill() for this batch.
    // Returns true when the key was new.  May be called by a subset of lanes.
    __device__ __forceinline__ bool insert(uint32_t key, bool to_overflow)
    {
        const uint32_t hk = key * 0x9E3779B1u;
        if (!to_overflow) return probe_insert(tab, (uint32_t) slots, __umulhi(hk, (uint32_t) slots), key);
        if (contains(key)) return false;
        return probe_insert(otab, (uint32_t) oslots, __umulhi(hk, (uint32_t) oslots), key);
    }
    // account for `n_new` keys inserted by the last batch
    __device__ __forceinline__ void added(int n_new, bool to_overflow)
    {
        if (to_overflow) { ocount += n_new; odirty = true; }
        else count += n_new;
    }
};

struct VisitedBitmap {       // HBM, one bit per element (large-visited-set path)
    uint32_t *bits;
    int words, count;
    __device__ __forceinline__ void configure(int) {}
    __device__ __forceinline__ void clear(int lane)
    {
        for (int i = lane; i < words; i += 32) bits[i] = 0u;
        count = 0;
        __syncwarp();
    }
    __device__ __forceinline__ bool room(int) const { return true; }
    __device__ __forceinline__ bool spill(int) const { return false; }
    __device__ __forceinline__ bool insert(uint32_t key, bool)
    {
        const uint32_t bit = 1u << (key & 31);
        return (atomicOr(bits + (key >> 5), bit) & bit) == 0;
    }
    __device__ __forceinline__ void added(int n_new, bool) { count += n_new; }
};

// ---- discarded candidates (hnsw.iterative_scan) ------------------------------------------------
// pgvector 0.8's HnswSearchLayer can keep what it throws away -- candidates evicted from W and
// candidates that were evaluated but not admitted -- so that the scan can be resumed from them
// (hnswscan.c ResumeScanItems [RECALL]).  NoDiscard compiles to nothing.
struct NoDiscard {
    static constexpr bool enabled = false;
    __device__ __forceinline__ void push1(float, uint32_t, int) {}
    __device__ __forceinline__ void push_mask(unsigned, float, uint32_t, int) {}
};
struct DiscList {            // unsorted list in HBM, one per query; n is warp-uniform
    static constexpr bool enabled = true;
    float *d;
    uint32_t *id;
    int n, cap;
    bool overflow;
    __device__ __forceinline__ void push1(float dd, uint32_t i, int lane)
    {
        if (n < cap) { if (lane == 0) { d[n] = dd; id[n] = i; } }
        else overflow = true;
        n++;
    }
    // every lane flagged in `mask` contributes its (dd, i)
    __device__ __forceinline__ void push_mask(unsigned mask, float dd, uint32_t i, int lane)
    {
        const int pos = n + __popc(mask & ((1u << lane) - 1u));
        if ((mask >> lane) & 1u) {
            if (pos < cap) { d[pos] = dd; id[pos] = i; }
        }
        n += __popc(mask);
        if (n > cap) overflow = true;
    }
};

// ---- evaluated distances (build) ----------------------------------------------------------------
// The build keeps every query<->element distance its candidate search evaluates in a per-element
// hash table in HBM: when the search of a layer ends, every member of W has been expanded, so the
// new element already knows its distance to each neighbour of each neighbour it will link to --
// exactly the distances HnswUpdateConnection asks for next (link_memo_kernel).  NoEvalSink
// compiles to nothing.
struct NoEvalSink {
    static constexpr bool enabled = false;
    __device__ __forceinline__ void put_mask(unsigned, uint32_t, float, int) {}
};
// what the search writes: an append-only log (two coalesced stores per batch of evaluations, nothing
// to wait for); eval_table_build_kernel turns the logs into the hash tables afterwards
struct EvalLog {
    static constexpr bool enabled = true;
    uint32_t *id;
    float *d;
    int n, cap;              // warp-uniform
    __device__ __forceinline__ void put_mask(unsigned mask, uint32_t i, float dd, int lane)
    {
        const int pos = n + __popc(mask & ((1u << lane) - 1u));
        if (((mask >> lane) & 1u) && pos < cap) { id[pos] = i; d[pos] = dd; }
        n += __popc(mask);
    }
};
struct EvalTable {           // open addressing, linear probing, at most EVAL_PROBES probes; EMPTY = free
    static constexpr bool enabled = true;
    static constexpr int EVAL_PROBES = 16;
    uint32_t *key;
    float *val;
    uint32_t mask;           // slots - 1
    __device__ __forceinline__ void put(uint32_t id, float d, bool active)
    {
        if (!active) return;
        uint32_t h = (id * 0x9E3779B1u) >> 8 & mask;
        for (int probe = 0; probe < EVAL_PROBES; probe++) {
            const uint32_t old = atomicCAS(key + h, EMPTY, id);
            if (old == EMPTY || old == id) { val[h] = d; return; }
            h = (h + 1) & mask;
        }
        // table crowded: the reader computes this one again
    }
    __device__ __forceinline__ bool get(uint32_t id, float &d) const
    {
        uint32_t h = (id * 0x9E3779B1u) >> 8 & mask;
        for (int probe = 0; probe < EVAL_PROBES; probe++) {
            const uint32_t k = __ldcg(key + h);
            if (k == id) { d = __ldcg(val + h); return true; }
            if (k == EMPTY) return false;
            h = (h + 1) & mask;
        }
        return false;
    }
};

// ---- W list ---------------------------------------------------------------------------------
// insert (ed, eid) keeping key order; then trim to ef + boundary ties.  `low` = every entry below
// it is expanded.
template <typename DS>
__device__ __forceinline__ int wlist_insert(WList &w, float ed, uint32_t eid, int ef, int lane, int &low, DS &ds)
{
    if (w.L + 1 > w.cap) return ST_TAIL;
    int pos = 0;
    for (int base = 0; base < w.L; base += 32) {
        const int i = base + lane;
        bool lt = false;
        if (i < w.L) {
            const float d = w.d[i];
            lt = d < ed || (d == ed && (w.id[i] & ID_MASK) < eid);
        }
        pos += __popc(__ballot_sync(FULL, lt));
    }
    for (int hi = w.L; hi > pos; hi -= 32) {
        const int lo = max(pos, hi - 32);
        const int i = lo + lane;
        const bool act = i < hi;
        float d = 0.f;
        uint32_t x = 0u;
        if (act) { d = w.d[i]; x = w.id[i]; }
        __syncwarp();
        if (act) { w.d[i + 1] = d; w.id[i + 1] = x; }
        __syncwarp();
    }
    if (lane == 0) { w.d[pos] = ed; w.id[pos] = eid; }
    __syncwarp();
    w.L++;
    if (w.L > ef) {
        const float f = w.d[ef - 1];
        int keep = 0;
        for (int base = ef; base < w.L; base += 32) {
            const int i = base + lane;
            const bool eq = i < w.L && w.d[i] == f;
            const unsigned b = __ballot_sync(FULL, eq);
            const int run = (b == FULL) ? 32 : (__ffs(~b) - 1);
            keep += run;
            if (run < 32) break;
        }
        if constexpr (DS::enabled) {
            // evicted from W for good: they go to the discarded list
            for (int base = ef + keep; base < w.L; base += 32) {
                const int i = base + lane;
                const bool act = i < w.L;
                ds.push_mask(__ballot_sync(FULL, act), act ? w.d[i] : 0.f, act ? (w.id[i] & ID_MASK) : 0u, lane);
            }
        }
        w.L = ef + keep;
    }
    if (pos < low) low = pos;
    return ST_OK;
}
__device__ __forceinline__ int wlist_insert(WList &w, float ed, uint32_t eid, int ef, int lane, int &low)
{
    NoDiscard nd;
    return wlist_insert(w, ed, eid, ef, lane, low, nd);
}

// make the first min(L, keep) entries the entry list of the next HnswSearchLayer call: drop the
// tie tail, clear expanded bits, reset the visited set and mark the entries visited.
// Returns ST_TABLE when neither visited table can hold the entries (the caller hands the query to the
// large-visited-set path).
template <typename VS>
__device__ __forceinline__ int wlist_as_entries(WList &w, VS &vs, int keep, int lane)
{
    if (w.L > keep) w.L = keep;
    vs.clear(lane);
    if (!vs.room(w.L)) return ST_TABLE;
    const bool sp = vs.spill(w.L);
    for (int base = 0; base < w.L; base += 32) {
        const int i = base + lane;
        if (i < w.L) {
            const uint32_t id = w.id[i] & ID_MASK;
            w.id[i] = id;
            vs.insert(id, sp);
        }
    }
    vs.added(w.L, sp);
    __syncwarp();
    return ST_OK;
}

struct QueryCounters { int n_dist, n_hop0, n_hopu; };

// distances of the lanes flagged in `mask` (each flagged lane holds a candidate id in nb);
// result returned in the flagged lane, +inf elsewhere.
template <typename T, int IP, int NV, int G>
__device__ __forceinline__ float eval_candidates(const GraphView &g, const float *q, int32_t nb, unsigned mask,
                                                 int lane)
{
    float myd = __int_as_float(0x7f800000);
    unsigned rem = mask;
#define HB_EVAL_GROUP(GG)                                                                          \
    {                                                                                              \
        int32_t ids[GG];                                                                           \
        int src[GG];                                                                               \
        _Pragma("unroll") for (int c = 0; c < GG; c++)                                             \
        {                                                                                          \
            src[c] = __ffs(rem) - 1;                                                               \
            rem &= rem - 1;                                                                        \
            ids[c] = __shfl_sync(FULL, nb, src[c]);                                                \
        }                                                                                          \
        const float s = group_distance<T, IP, NV, GG>(g.vecs, (uint32_t) g.row_bytes, g.nvec, q, ids, lane); \
        _Pragma("unroll") for (int c = 0; c < GG; c++)                                             \
        {                                                                                          \
            const float v = __shfl_sync(FULL, s, c * (32 / GG));                                   \
            if (lane == src[c]) myd = v;                                                           \
        }                                                                                          \
    }
    if constexpr (G >= 8) while (__popc(rem) >= 8) HB_EVAL_GROUP(8)
    if constexpr (G >= 4) while (__popc(rem) >= 4) HB_EVAL_GROUP(4)
    if constexpr (G >= 2) while (__popc(rem) >= 2) HB_EVAL_GROUP(2)
    while (rem) HB_EVAL_GROUP(1)
#undef HB_EVAL_GROUP
    return myd;
}

// as eval_candidates, against two staged queries: each candidate row is fetched once
template <typename T, int IP, int NV, int G>
__device__ __forceinline__ void eval_candidates2(const GraphView &g, const float *q0, const float *q1, int32_t nb,
                                                 unsigned mask, int lane, float &d0, float &d1)
{
    d0 = d1 = __int_as_float(0x7f800000);
    unsigned rem = mask;
#define HB_EVAL_GROUP2(GG)                                                                         \
    {                                                                                              \
        int32_t ids[GG];                                                                           \
        int src[GG];                                                                               \
        _Pragma("unroll") for (int c = 0; c < GG; c++)                                             \
        {                                                                                          \
            src[c] = __ffs(rem) - 1;                                                               \
            rem &= rem - 1;                                                                        \
            ids[c] = __shfl_sync(FULL, nb, src[c]);                                                \
        }                                                                                          \
        float s0, s1;                                                                              \
        group_distance2<T, IP, NV, GG>(g.vecs, (uint32_t) g.row_bytes, g.nvec, q0, q1, ids, lane, s0, s1); \
        _Pragma("unroll") for (int c = 0; c < GG; c++)                                             \
        {                                                                                          \
            const float v0 = __shfl_sync(FULL, s0, c * (32 / GG));                                 \
            const float v1 = __shfl_sync(FULL, s1, c * (32 / GG));                                 \
            if (lane == src[c]) { d0 = v0; d1 = v1; }                                              \
        }                                                                                          \
    }
    if constexpr (G >= 4) while (__popc(rem) >= 4) HB_EVAL_GROUP2(4)
    if constexpr (G >= 2) while (__popc(rem) >= 2) HB_EVAL_GROUP2(2)
    while (rem) HB_EVAL_GROUP2(1)
#undef HB_EVAL_GROUP2
}

// HnswSearchLayer.  Precondition: w holds the entry candidates (sorted, unexpanded) and vs holds
// exactly their ids.  Postcondition: w[0 .. min(L, ef)) = the result, nearest first.
template <typename T, int IP, int NV, int G, typename VS, typename DS, typename ES>
__device__ __forceinline__ int search_layer(const GraphView &g, WList &w, VS &vs, const float *q, int ef, int lc,
                                            int lane, QueryCounters &ctr, DS &ds, ES &es)
{
    const int deg = lc == 0 ? 2 * g.m : g.m;
    int low = 0;
    for (;;) {
        int idx = -1;
        for (int base = low; base < w.L; base += 32) {
            const int i = base + lane;
            const bool un = i < w.L && !(w.id[i] & EXP_BIT);
            const unsigned b = __ballot_sync(FULL, un);
            if (b) {
                idx = base + __ffs(b) - 1;
                // the candidate after this one is the likeliest next expansion: start its neighbour list
                // towards L2 now, a whole hop ahead of the load that will want it
                const unsigned b2 = b & (b - 1);
                if (lc == 0 && b2 && lane == 0) {
                    const uint32_t nid = w.id[base + __ffs(b2) - 1] & ID_MASK;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(g.nbr0 + (size_t) nid * deg));
                }
                break;
            }
        }
        if (idx < 0) break;
        const uint32_t cid = w.id[idx];
        __syncwarp();
        if (lane == 0) w.id[idx] = cid | EXP_BIT;
        __syncwarp();
        low = idx + 1;
        if (lc == 0) ctr.n_hop0++; else ctr.n_hopu++;
        if (!vs.room(deg)) return ST_TABLE;
        const int32_t *list = lc == 0 ? g.nbr0 + (size_t) cid * deg
                                      : g.nbru + ((size_t) g.uoff[cid] + (lc - 1)) * g.m;
        for (int cb = 0; cb < deg; cb += 32) {
            const int i = cb + lane;
            const int32_t nb = i < deg ? list[i] : -1;
            const bool sp = vs.spill(min(32, deg - cb));
            bool isnew = false;
            if (nb >= 0) isnew = vs.insert((uint32_t) nb, sp);
            const unsigned nmask = __ballot_sync(FULL, isnew);
            if (nmask == 0) continue;
            vs.added(__popc(nmask), sp);
            ctr.n_dist += __popc(nmask);
            const float myd = eval_candidates<T, IP, NV, G>(g, q, nb, nmask, lane);
            if constexpr (ES::enabled) es.put_mask(nmask, (uint32_t) nb, myd, lane);
            const bool full = w.L >= ef;
            const float f = full ? w.d[ef - 1] : 0.f;
            unsigned amask = __ballot_sync(FULL, isnew && (!full || myd < f));
            if constexpr (DS::enabled) ds.push_mask(nmask & ~amask, myd, (uint32_t) nb, lane);   // evaluated, not admitted
            while (amask) {
                const int s = __ffs(amask) - 1;
                amask &= amask - 1;
                const float ed = __shfl_sync(FULL, myd, s);
                const uint32_t eid = (uint32_t) __shfl_sync(FULL, nb, s);
                if (w.L >= ef && !(ed < w.d[ef - 1])) {
                    if constexpr (DS::enabled) ds.push1(ed, eid, lane);
                    continue;
                }
                const int st = wlist_insert(w, ed, eid, ef, lane, low, ds);
                if (st) return st;
            }
        }
    }
    return ST_OK;
}

template <typename T, int IP, int NV, int G, typename VS, typename DS>
__device__ __forceinline__ int search_layer(const GraphView &g, WList &w, VS &vs, const float *q, int ef, int lc,
                                            int lane, QueryCounters &ctr, DS &ds)
{
    NoEvalSink ne;
    return search_layer<T, IP, NV, G, VS, DS, NoEvalSink>(g, w, vs, q, ef, lc, lane, ctr, ds, ne);
}
template <typename T, int IP, int NV, int G, typename VS>
__device__ __forceinline__ int search_layer(const GraphView &g, WList &w, VS &vs, const float *q, int ef, int lc,
                                            int lane, QueryCounters &ctr)
{
    NoDiscard nd;
    NoEvalSink ne;
    return search_layer<T, IP, NV, G, VS, NoDiscard, NoEvalSink>(g, w, vs, q, ef, lc, lane, ctr, nd, ne);
}

// distance of the single element `e` (warp-uniform) to the staged query
template <typename T, int IP, int NV>
__device__ __forceinline__ float one_distance(const GraphView &g, const float *q, int32_t e, int lane)
{
    const int32_t ids[1] = { e };
    return group_distance<T, IP, NV, 1>(g.vecs, (uint32_t) g.row_bytes, g.nvec, q, ids, lane);
}

}   // namespace hb

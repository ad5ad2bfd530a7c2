// pgpages.cu -- reader for the pages of a pgvector HNSW index relation: the
// route by which a graph built by real pgvector can be loaded here and its search results compared.
//
// LAYOUT IS RECALLED, NOT VERIFIED: the reference mount has no source (/root/reference/README.md:1)
// and the image has no PostgreSQL, so no index file exists here to read.  What is parsed is the
// layout of upstream pgvector 0.7 / 0.8 hnsw.h and PostgreSQL's bufpage.h / itemptr.h / itemid.h as
// remembered (little-endian, BLCKSZ 8192):
//   page      PageHeaderData 24 B (pd_lower at +12, pd_upper +14, pd_special +16), ItemIdData[] of
//             4 B (lp_off:15 | lp_flags:2 | lp_len:15), tuples, special = HnswPageOpaqueData
//             { uint32 nextblkno; uint16 unused; uint16 page_id = 0xFF90 }
//   block 0   HnswMetaPageData at +24 { uint32 magic 0xA953A953; uint32 version 1; uint32 dimensions;
//             uint16 m; uint16 efConstruction; uint32 entryBlkno; uint16 entryOffno; int16 entryLevel;
//             uint32 insertPage }
//   element   { uint8 type = 1; uint8 level; uint8 deleted; uint8 version; ItemPointerData
//             heaptids[10]; ItemPointerData neighbortid; uint16 unused; varlena vector { int32
//             vl_len_ (4-byte header, length << 2); int16 dim; int16 unused; float|half x[dim] } }
//   neighbour { uint8 type = 2; uint8 version; uint16 count; ItemPointerData indextids[count] },
//             count = (level + 2) * m: layers level..1 with m slots each, then layer 0 with 2m
//   ItemPointerData { uint16 bi_hi; uint16 bi_lo; uint16 ip_posid }, invalid = posid 0
// Elements are numbered in (block, offset) order; heap TIDs become (block << 16) | offset.
// tests/test_pgpages.py checks the reader against a writer of the same recalled layout -- that
// pins the two to each other, not to pgvector.
#include "index.h"

#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

constexpr int BLCKSZ = 8192;
constexpr uint32_t HNSW_MAGIC = 0xA953A953u;
constexpr uint16_t HNSW_PAGE_ID = 0xFF90;

template <typename V> V rd(const uint8_t *p) { V v; memcpy(&v, p, sizeof v); return v; }

struct Tid { uint32_t blk; uint16_t off; };
Tid rd_tid(const uint8_t *p)
{
    const uint16_t hi = rd<uint16_t>(p), lo = rd<uint16_t>(p + 2);
    return { ((uint32_t) hi << 16) | lo, rd<uint16_t>(p + 4) };
}
inline uint64_t tid_key(Tid t) { return ((uint64_t) t.blk << 16) | t.off; }

struct Elem { const uint8_t *tup; uint32_t blk; uint16_t off; };

}   // namespace

// metapage fields and tuple counts, so that the caller can size hb_index_create (no device needed)
extern "C" int hb_pgvector_pages_info(const void *pages_v, int64_t n_pages, int *dim, int *m, int *ef_construction,
                                      int64_t *n_elements, int64_t *upper_rows)
{
    using hb::set_error;
    if (!pages_v || n_pages < 1) { set_error("hb_pgvector_pages_info: bad argument"); return HB_EINVAL; }
    const uint8_t *pages = (const uint8_t *) pages_v;
    const uint8_t *meta = pages + 24;
    if (rd<uint32_t>(meta) != HNSW_MAGIC) { set_error("not a pgvector hnsw index: magic %08x", rd<uint32_t>(meta)); return HB_EINVAL; }
    if (dim) *dim = (int) rd<uint32_t>(meta + 8);
    if (m) *m = rd<uint16_t>(meta + 12);
    if (ef_construction) *ef_construction = rd<uint16_t>(meta + 14);
    int64_t ne = 0, ur = 0;
    for (int64_t b = 1; b < n_pages; b++) {
        const uint8_t *pg = pages + b * BLCKSZ;
        const int lower = rd<uint16_t>(pg + 12);
        if (lower < 24 || lower > BLCKSZ) { set_error("block %lld: corrupt page header", (long long) b); return HB_EINVAL; }
        for (int i = 0; i < (lower - 24) / 4; i++) {
            const uint32_t lp = rd<uint32_t>(pg + 24 + 4 * i);
            const int off = lp & 0x7fff, flags = (lp >> 15) & 3, len = lp >> 17;
            if (flags != 1 || len < 4 || off + len > BLCKSZ) continue;
            if (pg[off] == 1) { ne++; ur += pg[off + 1]; }
        }
    }
    if (n_elements) *n_elements = ne;
    if (upper_rows) *upper_rows = ur;
    return HB_OK;
}

extern "C" int hb_index_load_pgvector_pages(hb_index *ix, const void *pages_v, int64_t n_pages)
{
    using hb::set_error;
    if (!ix || !pages_v || n_pages < 1) { set_error("hb_index_load_pgvector_pages: bad argument"); return HB_EINVAL; }
    const uint8_t *pages = (const uint8_t *) pages_v;
    const uint8_t *meta = pages + 24;
    if (rd<uint32_t>(meta) != HNSW_MAGIC) { set_error("not a pgvector hnsw index: magic %08x", rd<uint32_t>(meta)); return HB_EINVAL; }
    if (rd<uint32_t>(meta + 4) != 1) { set_error("unsupported hnsw index version %u", rd<uint32_t>(meta + 4)); return HB_EINVAL; }
    const uint32_t dims = rd<uint32_t>(meta + 8);
    const int m = rd<uint16_t>(meta + 12);
    const Tid entry_tid = { rd<uint32_t>(meta + 16), rd<uint16_t>(meta + 20) };
    if ((int) dims != ix->dim || m != ix->m) {
        set_error("index pages hold dimensions=%u m=%d, the handle was created with dim=%d m=%d", dims, m, ix->dim, ix->m);
        return HB_EINVAL;
    }
    // pass 1: element tuples in (block, offset) order
    std::vector<Elem> elems;
    std::unordered_map<uint64_t, int32_t> id_of;
    std::unordered_map<uint64_t, const uint8_t *> nbr_tuple;
    for (int64_t b = 1; b < n_pages; b++) {
        const uint8_t *pg = pages + b * BLCKSZ;
        const int lower = rd<uint16_t>(pg + 12), special = rd<uint16_t>(pg + 16);
        if (lower < 24 || lower > BLCKSZ || special > BLCKSZ) { set_error("block %lld: corrupt page header", (long long) b); return HB_EINVAL; }
        if (special + 8 <= BLCKSZ && rd<uint16_t>(pg + special + 6) != HNSW_PAGE_ID) { set_error("block %lld is not an hnsw page", (long long) b); return HB_EINVAL; }
        const int nitems = (lower - 24) / 4;
        for (int i = 0; i < nitems; i++) {
            const uint32_t lp = rd<uint32_t>(pg + 24 + 4 * i);
            const int off = lp & 0x7fff, flags = (lp >> 15) & 3, len = lp >> 17;
            if (flags != 1 || len < 4 || off + len > BLCKSZ) continue;          // LP_NORMAL only
            const uint8_t *tup = pg + off;
            const Tid self = { (uint32_t) b, (uint16_t) (i + 1) };
            if (tup[0] == 1) {
                // HnswElementTupleData: 72 bytes of header, then the vector datum (8 + dim * esize): a tuple that is
                // shorter cannot be this handle's type (e.g. a halfvec index loaded into a vector_* handle)
                if ((size_t) len < 80 + (size_t) ix->dim * ix->esize) {
                    set_error("block %lld item %d: element tuple of %d bytes is too short for %d dimensions of %d bytes (wrong operator class?)",
                              (long long) b, i + 1, len, ix->dim, ix->esize);
                    return HB_EINVAL;
                }
                id_of[tid_key(self)] = (int32_t) elems.size();
                elems.push_back({ tup, (uint32_t) b, (uint16_t) (i + 1) });
            } else if (tup[0] == 2) {
                if (len < 4 + 6 * (int) rd<uint16_t>(tup + 2)) {
                    set_error("block %lld item %d: neighbour tuple of %d bytes cannot hold its %d slots", (long long) b, i + 1, len,
                              (int) rd<uint16_t>(tup + 2));
                    return HB_EINVAL;
                }
                nbr_tuple[tid_key(self)] = tup;
            }
        }
    }
    const int64_t n = (int64_t) elems.size();
    if (n > ix->cap && !ix->opt_auto_grow) { set_error("index pages hold %lld elements, capacity is %lld", (long long) n, (long long) ix->cap); return HB_ENOMEM; }
    const int m2 = 2 * m;
    const size_t rowb = (size_t) ix->dim * ix->esize;
    std::vector<char> vecs((size_t) n * rowb);
    std::vector<uint8_t> level(n), ntids(n);
    std::vector<int64_t> tids((size_t) n * HB_HEAPTIDS, 0);
    std::vector<int32_t> nbr0((size_t) n * m2, -1), uoff(n, -1), nbru;
    int64_t urows = 0;
    for (int64_t e = 0; e < n; e++) {
        const uint8_t *t = elems[e].tup;
        level[e] = t[1];
        int nt = 0;
        for (int k = 0; k < HB_HEAPTIDS; k++) {
            const Tid h = rd_tid(t + 4 + 6 * k);
            if (h.off == 0) break;                                               // ItemPointerIsValid
            tids[(size_t) e * HB_HEAPTIDS + nt++] = (int64_t) tid_key(h);
        }
        ntids[e] = t[2] ? 0 : (uint8_t) nt;                                      // a deleted element returns nothing
        const uint8_t *v = t + 72;                                               // 4 + 60 + 6 + 2
        const int vdim = rd<int16_t>(v + 4);
        if (vdim != ix->dim) { set_error("element %lld has %d dimensions", (long long) e, vdim); return HB_EINVAL; }
        memcpy(&vecs[(size_t) e * rowb], v + 8, rowb);
        if (level[e] > 0) { uoff[e] = (int32_t) urows; urows += level[e]; }
    }
    if (urows > ix->upper_cap && !ix->opt_auto_grow) { set_error("index pages hold %lld upper-layer rows, capacity is %lld", (long long) urows, (long long) ix->upper_cap); return HB_ENOMEM; }
    nbru.assign((size_t) std::max<int64_t>(urows, 1) * m, -1);
    // pass 2: neighbour tuples
    for (int64_t e = 0; e < n; e++) {
        const Tid nt = rd_tid(elems[e].tup + 64);
        auto it = nbr_tuple.find(tid_key(nt));
        if (it == nbr_tuple.end()) { set_error("element %lld: neighbour tuple (%u,%u) not found", (long long) e, nt.blk, nt.off); return HB_EINVAL; }
        const uint8_t *t = it->second;
        const int count = rd<uint16_t>(t + 2), lv = level[e];
        if (count != (lv + 2) * m) { set_error("element %lld: neighbour tuple holds %d slots, expected %d", (long long) e, count, (lv + 2) * m); return HB_EINVAL; }
        int slot = 0;
        for (int lc = lv; lc >= 0; lc--) {
            const int lm = lc == 0 ? m2 : m;
            int32_t *dst = lc == 0 ? &nbr0[(size_t) e * m2] : &nbru[((size_t) uoff[e] + (lc - 1)) * m];
            int w = 0;
            for (int j = 0; j < lm; j++, slot++) {
                const Tid x = rd_tid(t + 4 + 6 * slot);
                if (x.off == 0) continue;
                auto f = id_of.find(tid_key(x));
                if (f != id_of.end()) dst[w++] = f->second;                      // lists stay compact
            }
        }
    }
    int32_t entry = -1;
    if (n > 0 && entry_tid.off != 0) {
        auto f = id_of.find(tid_key(entry_tid));
        if (f == id_of.end()) { set_error("entry point (%u,%u) is not an element", entry_tid.blk, entry_tid.off); return HB_EINVAL; }
        entry = f->second;
    }
    return hb_index_load(ix, n, urows, entry, vecs.data(), level.data(), nbr0.data(), uoff.data(), nbru.data(), ntids.data(),
                         tids.data());
}

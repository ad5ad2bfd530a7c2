/*
 * hnsw_b200_am.c -- PostgreSQL index access method glue over libhnsw_b200.so (include/hnsw_b200.h).
 *
 * What a maintainer of the reference extension adds to route the HNSW hot path to the GPU library: an
 * IndexAmRoutine whose build / insert / vacuum / scan callbacks forward to the C ABI.  It takes the roles of
 * upstream pgvector's hnsw.c (hnswhandler), hnswbuild.c (hnswbuild, hnswbuildempty), hnswinsert.c
 * (hnswinsert), hnswvacuum.c (hnswbulkdelete, hnswvacuumcleanup) and hnswscan.c (hnswbeginscan, hnswrescan,
 * hnswgettuple, hnswendscan) [RECALL: the reference mount has no source, /root/reference/README.md:1].
 *
 * Built only with -DHNSW_B200_WITH_POSTGRES, against the server headers, and linked with -lhnsw_b200.  This
 * image has no PostgreSQL: the CPU test suite compile-checks the file against tests/c/pgstub/ (a stand-in that
 * declares the index-AM API with the real PostgreSQL 16 names and signatures), so a drifting signature breaks
 * a test; it has never run inside a server, and says so.
 *
 * Not here (the extension keeps its own): reloption parsing (hnswoptions), cost estimation (hnswcostestimate),
 * opclass validation (hnswvalidate), and durability -- the GPU library holds the graph in HBM; persisting it
 * (hb_index_export at checkpoint, hb_index_load / hb_index_load_pgvector_pages at first use) is the storage
 * engine's side of the boundary and out of scope of the hot path.
 */
#ifdef HNSW_B200_WITH_POSTGRES

#include "postgres.h"
#include "access/amapi.h"

#include "hnsw_b200.h"

/* ---- supplied by the extension (unchanged upstream code) ------------------------------------------------ */
/* reloptions of the index: upstream's HnswOptions { int32 vl_len_; int m; int efConstruction; } */
typedef struct HnswB200Options { int32 vl_len_; int m; int efConstruction; } HnswB200Options;
#define HNSW_DEFAULT_M 16
#define HNSW_DEFAULT_EF_CONSTRUCTION 64
/* operator class of the indexed column -> HB_L2 / HB_IP / HB_COSINE / HB_L1, HB_F32 / HB_F16, and typmod dimensions */
extern void hnsw_b200_opclass(Relation index, int *metric, int *dtype, int *dim);
extern bytea *hnswoptions(Datum reloptions, bool validate);
extern void hnswcostestimate(struct PlannerInfo *root, struct IndexPath *path, double loop_count, Cost *indexStartupCost,
                             Cost *indexTotalCost, Selectivity *indexSelectivity, double *indexCorrelation, double *indexPages);
extern bool hnswvalidate(Oid opclassoid);

/* ---- GUCs (hnsw.c _PG_init upstream) ------------------------------------------------------------------- */
static int hnsw_ef_search = 40;                 /* hnsw.ef_search, 1..1000 */
static int hnsw_iterative_scan = HB_ITER_OFF;   /* hnsw.iterative_scan */
static int hnsw_max_scan_tuples = 20000;        /* hnsw.max_scan_tuples */
static int hnsw_b200_device = 0;                /* which GPU this backend's indexes live on */
static const struct config_enum_entry hnsw_iterative_scan_options[] = {
    { "off", HB_ITER_OFF, false }, { "relaxed_order", HB_ITER_RELAXED, false }, { "strict_order", HB_ITER_STRICT, false },
    { NULL, 0, false } };

void hnsw_b200_init_gucs(void)
{
    DefineCustomIntVariable("hnsw.ef_search", "Sets the size of the dynamic candidate list for search", "Valid range is 1..1000.",
                            &hnsw_ef_search, 40, 1, 1000, PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomEnumVariable("hnsw.iterative_scan", "Sets the mode for iterative scans", NULL, &hnsw_iterative_scan, HB_ITER_OFF,
                             hnsw_iterative_scan_options, PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomIntVariable("hnsw.max_scan_tuples", "Sets the max number of tuples to visit for iterative scans", NULL,
                            &hnsw_max_scan_tuples, 20000, 1, 0x7fffffff, PGC_USERSET, 0, NULL, NULL, NULL);
    DefineCustomIntVariable("hnsw_b200.device", "CUDA device holding this backend's HNSW indexes", NULL, &hnsw_b200_device, 0, 0,
                            63, PGC_USERSET, 0, NULL, NULL, NULL);
}

/* ---- error convention: return codes + hb_last_error() become ereport(ERROR) ----------------------------- */
static void b200_check(int64 rc)
{
    if (rc < 0)
        ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("hnsw_b200: %s", hb_last_error())));
}

/* heap TIDs cross the ABI as int64 = (block << 16) | offset */
static int64 b200_tid(ItemPointer t) { return ((int64) ItemPointerGetBlockNumber(t) << 16) | ItemPointerGetOffsetNumber(t); }

/* the column datum: vector { int32 vl_len_; int16 dim; int16 unused; float x[] } / halfvec likewise with half x[] */
static const void *b200_datum_values(Datum d) { return VARDATA_ANY(PG_DETOAST_DATUM(d)) + 4; }

/* ---- one hb_index per index relation, kept for the life of the backend ---------------------------------- */
#define B200_MAX_OPEN 64
static struct { Oid relid; hb_index *ix; } b200_open[B200_MAX_OPEN];
static int b200_n_open = 0;

static hb_index *b200_index_for(Relation index, int64 capacity_hint)
{
    int i, metric, dtype, dim, m = HNSW_DEFAULT_M, efc = HNSW_DEFAULT_EF_CONSTRUCTION;
    HnswB200Options *opts = (HnswB200Options *) index->rd_options;
    hb_index *ix;

    for (i = 0; i < b200_n_open; i++)
        if (b200_open[i].relid == RelationGetRelid(index))
            return b200_open[i].ix;
    if (b200_n_open == B200_MAX_OPEN)
        elog(ERROR, "hnsw_b200: too many open indexes");
    if (opts) { m = opts->m; efc = opts->efConstruction; }
    hnsw_b200_opclass(index, &metric, &dtype, &dim);
    /* capacity is an initial reservation only: inserts beyond it grow the index (hb_index_reserve) */
    ix = hb_index_create(hnsw_b200_device, dim, m, efc, metric, dtype, capacity_hint > 0 ? capacity_hint : 1 << 20,
                         (uint64_t) RelationGetRelid(index));
    if (ix == NULL)
        b200_check(HB_ECUDA);
    b200_open[b200_n_open].relid = RelationGetRelid(index);
    b200_open[b200_n_open].ix = ix;
    b200_n_open++;
    return ix;
}

/* ---- ambuild / ambuildempty / aminsert ----------------------------------------------------------------- */
#define B200_CHUNK 65536
typedef struct B200BuildState {
    hb_index *ix;
    size_t row_bytes;
    char *rows;           /* B200_CHUNK x row_bytes */
    int64 *tids;
    int64 n, indexed;
    double reltuples;
} B200BuildState;

static void b200_flush(B200BuildState *bs)
{
    int64 rc;
    if (bs->n == 0) return;
    rc = hb_insert(bs->ix, bs->rows, bs->n, bs->tids);      /* zero-norm vectors are skipped inside under cosine */
    b200_check(rc);
    bs->indexed += rc;
    bs->n = 0;
}

/* BuildCallback of hnswbuild.c: NULL vectors are skipped, everything else is appended to the chunk */
static void b200_build_callback(Relation index, ItemPointer tid, Datum *values, bool *isnull, bool tupleIsAlive, void *state)
{
    B200BuildState *bs = (B200BuildState *) state;
    (void) index; (void) tupleIsAlive;
    bs->reltuples += 1;
    if (isnull[0]) return;
    memcpy(bs->rows + (size_t) bs->n * bs->row_bytes, b200_datum_values(values[0]), bs->row_bytes);
    bs->tids[bs->n] = b200_tid(tid);
    if (++bs->n == B200_CHUNK) b200_flush(bs);
}

static IndexBuildResult *b200_build(Relation heap, Relation index, IndexInfo *indexInfo)
{
    B200BuildState bs;
    IndexBuildResult *result;
    int metric, dtype, dim;

    hnsw_b200_opclass(index, &metric, &dtype, &dim);
    memset(&bs, 0, sizeof bs);
    bs.ix = b200_index_for(index, 0);
    if (hb_index_size(bs.ix) != 0)
        elog(ERROR, "hnsw_b200: index is not empty");
    bs.row_bytes = (size_t) dim * (dtype == HB_F16 ? 2 : 4);
    bs.rows = palloc(bs.row_bytes * B200_CHUNK);
    bs.tids = palloc(sizeof(int64) * B200_CHUNK);
    (void) table_index_build_scan(heap, index, indexInfo, true, true, b200_build_callback, &bs, NULL);
    b200_flush(&bs);
    b200_check(hb_index_trim(bs.ix));           /* the in-memory build state is freed when CREATE INDEX ends */
    pfree(bs.rows);
    pfree(bs.tids);
    result = (IndexBuildResult *) palloc0(sizeof(IndexBuildResult));
    result->heap_tuples = bs.reltuples;
    result->index_tuples = (double) bs.indexed;
    return result;
}

static void b200_buildempty(Relation index)
{
    /* the init fork of an unlogged index is an empty graph: creating the handle is all there is to do */
    (void) b200_index_for(index, 0);
}

static bool b200_insert(Relation index, Datum *values, bool *isnull, ItemPointer heap_tid, Relation heap,
                        IndexUniqueCheck checkUnique, bool indexUnchanged, IndexInfo *indexInfo)
{
    int64 tid = b200_tid(heap_tid);
    (void) heap; (void) checkUnique; (void) indexUnchanged; (void) indexInfo;
    if (isnull[0]) return false;
    b200_check(hb_insert(b200_index_for(index, 0), b200_datum_values(values[0]), 1, &tid));
    return false;
}

/* ---- ambulkdelete / amvacuumcleanup -------------------------------------------------------------------- */
static IndexBulkDeleteResult *b200_bulkdelete(IndexVacuumInfo *info, IndexBulkDeleteResult *stats, IndexBulkDeleteCallback callback,
                                              void *callback_state)
{
    hb_index *ix = b200_index_for(info->index, 0);
    int64 n = hb_index_size(ix), i, n_dead = 0, removed;
    int64 *tids, *dead;
    uint8_t *ntids;
    int k;

    if (stats == NULL) stats = (IndexBulkDeleteResult *) palloc0(sizeof(IndexBulkDeleteResult));
    if (n <= 0) return stats;
    /* pass 1 (RemoveHeapTids): ask the callback about every heap TID the index holds */
    tids = palloc(sizeof(int64) * n * HB_HEAPTIDS);
    ntids = palloc(n);
    dead = palloc(sizeof(int64) * n * HB_HEAPTIDS);
    b200_check(hb_index_export(ix, NULL, NULL, NULL, NULL, NULL, ntids, tids));
    for (i = 0; i < n; i++)
        for (k = 0; k < ntids[i]; k++) {
            ItemPointerData ip;
            const int64 t = tids[i * HB_HEAPTIDS + k];
            ItemPointerSet(&ip, (BlockNumber) (t >> 16), (OffsetNumber) (t & 0xffff));
            if (callback(&ip, callback_state)) dead[n_dead++] = t;
        }
    removed = hb_bulk_delete(ix, dead, n_dead);
    b200_check(removed);
    /* passes 2 and 3 (RepairGraph, MarkDeleted): emptied elements leave the graph, their neighbours are re-linked */
    b200_check(hb_vacuum_repair(ix, NULL));
    stats->tuples_removed += (double) removed;
    stats->num_index_tuples = (double) hb_index_size(ix);
    pfree(tids); pfree(ntids); pfree(dead);
    return stats;
}

static IndexBulkDeleteResult *b200_vacuumcleanup(IndexVacuumInfo *info, IndexBulkDeleteResult *stats)
{
    if (info->analyze_only) return stats;
    if (stats == NULL) stats = (IndexBulkDeleteResult *) palloc0(sizeof(IndexBulkDeleteResult));
    stats->num_index_tuples = (double) hb_index_size(b200_index_for(info->index, 0));
    return stats;
}

/* ---- ambeginscan / amrescan / amgettuple / amendscan --------------------------------------------------- */
typedef struct B200ScanOpaque { hb_scan *scan; bool first; } B200ScanOpaque;

static IndexScanDesc b200_beginscan(Relation index, int nkeys, int norderbys)
{
    IndexScanDesc scan = RelationGetIndexScan(index, nkeys, norderbys);
    B200ScanOpaque *so = (B200ScanOpaque *) palloc0(sizeof(B200ScanOpaque));
    so->scan = hb_beginscan(b200_index_for(index, 0));
    if (so->scan == NULL) b200_check(HB_ECUDA);
    if (hnsw_iterative_scan != HB_ITER_OFF)
        b200_check(hb_scan_set_iterative(so->scan, hnsw_iterative_scan, hnsw_max_scan_tuples));
    so->first = true;
    scan->opaque = so;
    return scan;
}

static void b200_rescan(IndexScanDesc scan, ScanKey keys, int nkeys, ScanKey orderbys, int norderbys)
{
    B200ScanOpaque *so = (B200ScanOpaque *) scan->opaque;
    (void) nkeys; (void) norderbys;
    if (keys && scan->numberOfKeys > 0)
        memmove(scan->keyData, keys, scan->numberOfKeys * sizeof(ScanKeyData));
    if (orderbys && scan->numberOfOrderBys > 0)
        memmove(scan->orderByData, orderbys, scan->numberOfOrderBys * sizeof(ScanKeyData));
    so->first = true;
}

static bool b200_gettuple(IndexScanDesc scan, ScanDirection dir)
{
    B200ScanOpaque *so = (B200ScanOpaque *) scan->opaque;
    int64_t tid;
    float dist;
    int rc;

    if (dir != ForwardScanDirection)
        elog(ERROR, "hnsw_b200: index scans are forward only");
    if (so->first) {
        if (scan->orderByData == NULL)
            elog(ERROR, "cannot scan hnsw index without order");
        if (scan->orderByData->sk_flags & SK_ISNULL)
            return false;                         /* ORDER BY col <-> NULL returns nothing, as upstream */
        b200_check(hb_rescan(so->scan, b200_datum_values(scan->orderByData->sk_argument), hnsw_ef_search));
        so->first = false;
    }
    rc = hb_gettuple(so->scan, &tid, &dist);
    b200_check(rc);
    if (rc == 0) return false;
    ItemPointerSet(&scan->xs_heaptid, (BlockNumber) (tid >> 16), (OffsetNumber) (tid & 0xffff));
    scan->xs_recheck = false;
    scan->xs_recheckorderby = false;
    return true;
}

static void b200_endscan(IndexScanDesc scan)
{
    B200ScanOpaque *so = (B200ScanOpaque *) scan->opaque;
    hb_endscan(so->scan);
    pfree(so);
    scan->opaque = NULL;
}

/* ---- the handler: CREATE ACCESS METHOD hnsw TYPE INDEX HANDLER hnsw_b200_handler ------------------------- */
PGDLLEXPORT PG_FUNCTION_INFO_V1(hnsw_b200_handler);
Datum hnsw_b200_handler(PG_FUNCTION_ARGS)
{
    IndexAmRoutine *amroutine = makeNode(IndexAmRoutine);
    (void) fcinfo;
    amroutine->amstrategies = 0;
    amroutine->amsupport = 3;            /* HNSW_DISTANCE_PROC, HNSW_NORM_PROC, HNSW_TYPE_INFO_PROC */
    amroutine->amoptsprocnum = 0;
    amroutine->amcanorder = false;
    amroutine->amcanorderbyop = true;    /* ORDER BY col <-> $1 */
    amroutine->amcanbackward = false;
    amroutine->amcanunique = false;
    amroutine->amcanmulticol = false;
    amroutine->amoptionalkey = true;
    amroutine->amsearcharray = false;
    amroutine->amsearchnulls = false;
    amroutine->amstorage = false;
    amroutine->amclusterable = false;
    amroutine->ampredlocks = false;
    amroutine->amcanparallel = false;
    amroutine->amcaninclude = false;
    amroutine->amusemaintenanceworkmem = false;
    amroutine->amsummarizing = false;
    amroutine->amparallelvacuumoptions = 0;
    amroutine->amkeytype = InvalidOid;

    amroutine->ambuild = b200_build;
    amroutine->ambuildempty = b200_buildempty;
    amroutine->aminsert = b200_insert;
    amroutine->ambulkdelete = b200_bulkdelete;
    amroutine->amvacuumcleanup = b200_vacuumcleanup;
    amroutine->amcanreturn = NULL;
    amroutine->amcostestimate = hnswcostestimate;
    amroutine->amoptions = hnswoptions;
    amroutine->amproperty = NULL;
    amroutine->ambuildphasename = NULL;
    amroutine->amvalidate = hnswvalidate;
    amroutine->amadjustmembers = NULL;
    amroutine->ambeginscan = b200_beginscan;
    amroutine->amrescan = b200_rescan;
    amroutine->amgettuple = b200_gettuple;
    amroutine->amgetbitmap = NULL;
    amroutine->amendscan = b200_endscan;
    amroutine->ammarkpos = NULL;
    amroutine->amrestrpos = NULL;
    amroutine->amestimateparallelscan = NULL;
    amroutine->aminitparallelscan = NULL;
    amroutine->amparallelrescan = NULL;
    PG_RETURN_POINTER(amroutine);
}

#endif /* HNSW_B200_WITH_POSTGRES */

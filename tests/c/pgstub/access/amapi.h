/* stand-in for access/amapi.h + genam.h + relscan.h + itemptr.h + skey.h (PostgreSQL 16), see ../postgres.h */
#ifndef PGSTUB_AMAPI_H
#define PGSTUB_AMAPI_H
#include "postgres.h"

typedef struct ItemPointerData { uint16 bi_hi, bi_lo; OffsetNumber ip_posid; } ItemPointerData;
typedef ItemPointerData *ItemPointer;
static inline BlockNumber ItemPointerGetBlockNumber(const ItemPointerData *p) { return ((BlockNumber) p->bi_hi << 16) | p->bi_lo; }
static inline OffsetNumber ItemPointerGetOffsetNumber(const ItemPointerData *p) { return p->ip_posid; }
static inline void ItemPointerSet(ItemPointerData *p, BlockNumber b, OffsetNumber o) { p->bi_hi = (uint16) (b >> 16); p->bi_lo = (uint16) b; p->ip_posid = o; }

typedef struct RelationData { Oid rd_id; void *rd_options; struct { int natts; } *rd_att; Oid *rd_opfamily; } RelationData;
typedef RelationData *Relation;
#define RelationGetRelid(r) ((r)->rd_id)

typedef struct ScanKeyData { int sk_flags; int16_t sk_attno; uint16 sk_strategy; Oid sk_subtype; Oid sk_collation; Datum sk_argument; } ScanKeyData;
typedef ScanKeyData *ScanKey;
#define SK_ISNULL 0x0001

typedef enum ScanDirection { BackwardScanDirection = -1, NoMovementScanDirection = 0, ForwardScanDirection = 1 } ScanDirection;

typedef struct IndexScanDescData {
    Relation heapRelation, indexRelation;
    int numberOfKeys, numberOfOrderBys;
    ScanKeyData *keyData, *orderByData;
    bool xs_want_itup;
    void *opaque;
    ItemPointerData xs_heaptid;
    bool xs_recheck, xs_recheckorderby;
} IndexScanDescData;
typedef IndexScanDescData *IndexScanDesc;
IndexScanDesc RelationGetIndexScan(Relation indexRelation, int nkeys, int norderbys);

typedef struct IndexInfo { NodeTag type; int ii_NumIndexAttrs; bool ii_Concurrent; int ii_ParallelWorkers; } IndexInfo;
typedef struct IndexBuildResult { double heap_tuples, index_tuples; } IndexBuildResult;
typedef struct IndexVacuumInfo { Relation index; Relation heaprel; bool analyze_only, report_progress, estimated_count; int message_level; double num_heap_tuples; void *strategy; } IndexVacuumInfo;
typedef struct IndexBulkDeleteResult { BlockNumber num_pages; bool estimated_count; double num_index_tuples, tuples_removed; BlockNumber pages_newly_deleted, pages_deleted, pages_free; } IndexBulkDeleteResult;
typedef bool (*IndexBulkDeleteCallback) (ItemPointer itemptr, void *state);
typedef enum IndexUniqueCheck { UNIQUE_CHECK_NO, UNIQUE_CHECK_YES, UNIQUE_CHECK_PARTIAL, UNIQUE_CHECK_EXISTING } IndexUniqueCheck;
struct PlannerInfo; struct IndexPath; typedef double Cost; typedef double Selectivity; typedef struct bytea bytea;

/* heap scan of CREATE INDEX (tableam.h) */
typedef void (*IndexBuildCallback) (Relation index, ItemPointer tid, Datum *values, bool *isnull, bool tupleIsAlive, void *state);
double table_index_build_scan(Relation table_rel, Relation index_rel, IndexInfo *index_info, bool allow_sync, bool progress,
                              IndexBuildCallback callback, void *callback_state, void *scan);

/* the callback types of IndexAmRoutine, PostgreSQL 16 signatures */
typedef IndexBuildResult *(*ambuild_function) (Relation heapRelation, Relation indexRelation, IndexInfo *indexInfo);
typedef void (*ambuildempty_function) (Relation indexRelation);
typedef bool (*aminsert_function) (Relation indexRelation, Datum *values, bool *isnull, ItemPointer heap_tid, Relation heapRelation,
                                   IndexUniqueCheck checkUnique, bool indexUnchanged, IndexInfo *indexInfo);
typedef IndexBulkDeleteResult *(*ambulkdelete_function) (IndexVacuumInfo *info, IndexBulkDeleteResult *stats,
                                                         IndexBulkDeleteCallback callback, void *callback_state);
typedef IndexBulkDeleteResult *(*amvacuumcleanup_function) (IndexVacuumInfo *info, IndexBulkDeleteResult *stats);
typedef void (*amcostestimate_function) (struct PlannerInfo *root, struct IndexPath *path, double loop_count, Cost *indexStartupCost,
                                         Cost *indexTotalCost, Selectivity *indexSelectivity, double *indexCorrelation, double *indexPages);
typedef bytea *(*amoptions_function) (Datum reloptions, bool validate);
typedef bool (*amvalidate_function) (Oid opclassoid);
typedef IndexScanDesc (*ambeginscan_function) (Relation indexRelation, int nkeys, int norderbys);
typedef void (*amrescan_function) (IndexScanDesc scan, ScanKey keys, int nkeys, ScanKey orderbys, int norderbys);
typedef bool (*amgettuple_function) (IndexScanDesc scan, ScanDirection direction);
typedef void (*amendscan_function) (IndexScanDesc scan);

typedef struct IndexAmRoutine {
    NodeTag type;
    uint16 amstrategies, amsupport, amoptsprocnum;
    bool amcanorder, amcanorderbyop, amcanbackward, amcanunique, amcanmulticol, amoptionalkey, amsearcharray, amsearchnulls,
         amstorage, amclusterable, ampredlocks, amcanparallel, amcaninclude, amusemaintenanceworkmem, amsummarizing;
    uint8 amparallelvacuumoptions;
    Oid amkeytype;
    ambuild_function ambuild;
    ambuildempty_function ambuildempty;
    aminsert_function aminsert;
    ambulkdelete_function ambulkdelete;
    amvacuumcleanup_function amvacuumcleanup;
    void *amcanreturn;
    amcostestimate_function amcostestimate;
    amoptions_function amoptions;
    void *amproperty, *ambuildphasename;
    amvalidate_function amvalidate;
    void *amadjustmembers;
    ambeginscan_function ambeginscan;
    amrescan_function amrescan;
    amgettuple_function amgettuple;
    void *amgetbitmap;
    amendscan_function amendscan;
    void *ammarkpos, *amrestrpos, *amestimateparallelscan, *aminitparallelscan, *amparallelrescan;
} IndexAmRoutine;
#endif

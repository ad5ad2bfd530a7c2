/* tests/c/pgstub -- a MINIMAL STAND-IN for the PostgreSQL 16 server headers, written from the documented
 * index access method API (amapi.h / genam.h / relscan.h) so that pg/hnsw_b200_am.c can be compile-checked in
 * an image that has no PostgreSQL.  It declares only what the glue uses, with the real names and signatures;
 * it is NOT PostgreSQL and nothing here is linked.  A maintainer builds the glue against the real headers. */
#ifndef PGSTUB_POSTGRES_H
#define PGSTUB_POSTGRES_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef uintptr_t Datum;
typedef unsigned int Oid;
typedef uint16_t uint16;
typedef uint32_t uint32;
typedef int32_t int32;
typedef int64_t int64;
typedef uint8_t uint8;
typedef size_t Size;
typedef uint32 BlockNumber;
typedef uint16 OffsetNumber;
typedef int NodeTag;
#define T_IndexAmRoutine 1
#define InvalidOid ((Oid) 0)
#define PGDLLEXPORT
#define PG_MODULE_MAGIC extern int pgstub_module_magic
#define PG_FUNCTION_INFO_V1(f) extern Datum f(struct FunctionCallInfoBaseData *fcinfo)
struct FunctionCallInfoBaseData;
typedef struct FunctionCallInfoBaseData *FunctionCallInfo;
#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_RETURN_POINTER(x) return (Datum) (x)
#define PointerGetDatum(x) ((Datum) (x))
#define DatumGetPointer(x) ((char *) (x))

void *palloc(Size size);
void *palloc0(Size size);
void *repalloc(void *p, Size size);
void pfree(void *p);

/* varlena */
struct varlena { char vl_len_[4]; char vl_dat[1]; };
struct varlena *pg_detoast_datum(struct varlena *datum);
#define PG_DETOAST_DATUM(d) pg_detoast_datum((struct varlena *) DatumGetPointer(d))
#define VARDATA_ANY(p) (((struct varlena *) (p))->vl_dat)

/* elog / ereport, reduced to a call that does not return for ERROR */
#define ERROR 21
#define ERRCODE_INTERNAL_ERROR 1
#define ERRCODE_FEATURE_NOT_SUPPORTED 2
void pgstub_ereport(int level, int code, const char *fmt, ...) __attribute__((format(printf, 3, 4)));
#define elog(level, ...) pgstub_ereport(level, ERRCODE_INTERNAL_ERROR, __VA_ARGS__)
#define ereport(level, rest) pgstub_ereport_struct rest
#define errcode(c) (c)
#define errmsg(...) __VA_ARGS__
#define pgstub_ereport_struct(code, ...) pgstub_ereport(ERROR, code, __VA_ARGS__)

#define makeNode(T) ((T *) pgstub_make_node(sizeof(T), T_##T))
void *pgstub_make_node(Size size, NodeTag tag);

/* GUCs */
typedef enum { PGC_USERSET = 1 } GucContext;
struct config_enum_entry { const char *name; int val; bool hidden; };
void DefineCustomIntVariable(const char *name, const char *short_desc, const char *long_desc, int *valueAddr, int bootValue,
                             int minValue, int maxValue, GucContext context, int flags, void *check_hook, void *assign_hook,
                             void *show_hook);
void DefineCustomEnumVariable(const char *name, const char *short_desc, const char *long_desc, int *valueAddr, int bootValue,
                              const struct config_enum_entry *options, GucContext context, int flags, void *check_hook,
                              void *assign_hook, void *show_hook);
#endif

/* The C ABI from plain C (C99, -pedantic): what the Postgres-side glue of INTEGRATION.md compiles against.
 * Takes the address of every entry point (so a missing or mis-declared symbol fails the link), then walks the
 * error paths that need no GPU.  Exit code 0 = fine.  Built and run by tests/test_cabi.py. */
#include <stdio.h>
#include <string.h>
#include "hnsw_b200.h"

#define REF(f) do { void (*p)(void) = (void (*)(void)) f; if (!p) return 2; n_syms++; } while (0)

int main(void)
{
    int n_syms = 0;
    REF(hb_last_error); REF(hb_version); REF(hb_device_count);
    REF(hb_index_create); REF(hb_index_free); REF(hb_index_size); REF(hb_index_entry);
    REF(hb_build); REF(hb_insert); REF(hb_index_reserve); REF(hb_bulk_delete); REF(hb_vacuum_repair); REF(hb_index_trim); REF(hb_set_build_batch); REF(hb_level_for); REF(hb_set_option);
    REF(hb_index_load); REF(hb_index_load_pgvector_pages); REF(hb_pgvector_pages_info); REF(hb_index_upper_rows); REF(hb_index_export);
    REF(hb_beginscan); REF(hb_rescan); REF(hb_gettuple); REF(hb_endscan); REF(hb_scan_set_iterative);
    REF(hb_iter_begin); REF(hb_iter_next); REF(hb_iter_tuples); REF(hb_iter_end); REF(hb_search_batch_filtered);
    REF(hb_search_batch); REF(hb_search_batch_async); REF(hb_search_batch_wait); REF(hb_search_batch_elements); REF(hb_search_batch_dev); REF(hb_search_batch_status);
    REF(hb_distance_batch); REF(hb_distance_batch_dev); REF(hb_normalize); REF(hb_bruteforce);
    REF(hb_partition_of); REF(hb_partition_route); REF(hb_merge_topk_dev); REF(hb_elements_to_tids_dev);
    REF(hb_part_unique_id); REF(hb_part_create); REF(hb_part_free); REF(hb_part_owned); REF(hb_part_index); REF(hb_part_size);
    REF(hb_part_set_option); REF(hb_part_get_counters); REF(hb_part_build); REF(hb_part_search_async); REF(hb_part_search_wait); REF(hb_part_search);
    REF(hb_get_counters); REF(hb_get_per_query_counters); REF(hb_last_search_ms); REF(hb_search_layer); REF(hb_bruteforce_ex);

    printf("version %s, %d entry points, %d device(s)\n", hb_version(), n_syms, hb_device_count());
    /* argument checking happens before any device work */
    if (hb_index_create(0, 0, 16, 64, HB_L2, HB_F32, 10, 1) != NULL) return 3;           /* dim 0 */
    if (hb_index_create(0, 8, 16, 16, HB_L2, HB_F32, 10, 1) != NULL) return 4;           /* ef_construction < 2m */
    if (strlen(hb_last_error()) == 0) return 5;
    if (hb_rescan(NULL, NULL, 40) != HB_EINVAL) return 6;
    if (hb_level_for(1, 0, 16) < 0 || hb_level_for(1, 0, 1) != HB_EINVAL) return 7;
    if (hb_partition_of(12345, 8) < 0 || hb_partition_of(12345, 8) > 7) return 8;
    {
        unsigned char page[8192];
        memset(page, 0, sizeof page);
        if (hb_pgvector_pages_info(page, 1, NULL, NULL, NULL, NULL, NULL) != HB_EINVAL) return 9;   /* wrong magic */
    }
    if (hb_part_create(0, 8, 16, 64, HB_L2, HB_F32, 8, 10, 1, 1, 2, NULL) != NULL) return 12;   /* world 2 needs the communicator id */
    if (hb_part_search_wait(NULL, 0) != HB_EINVAL) return 13;
    if (hb_device_count() <= 0) {
        /* no CUDA device: there is no CPU fallback, creation must fail and say why */
        if (hb_index_create(0, 8, 16, 64, HB_L2, HB_F32, 10, 1) != NULL) return 10;
        if (strstr(hb_last_error(), "CUDA") == NULL && strstr(hb_last_error(), "device") == NULL) return 11;
    }
    return 0;
}

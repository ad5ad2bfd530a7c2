/* examples/knn.c -- the C ABI end to end from plain C: CREATE INDEX (hb_build), ordered scans through the
 * amgettuple mirror (hb_rescan / hb_gettuple) and through the batched call, on synthetic clustered rows.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/knn.c pgvector-hnsw-partitioning_b200/libhnsw_b200.so \
 *       -Wl,-rpath,$PWD/pgvector-hnsw-partitioning_b200 -lm -o knn && ./knn
 * Needs a CUDA device at run time (there is no CPU fallback); tests/test_cabi.py only checks that it builds. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "hnsw_b200.h"

static float frand(unsigned *s) { *s = *s * 1664525u + 1013904223u; return (float) (*s >> 8) / 16777216.0f; }

int main(void)
{
    const int dim = 64, n = 20000, nq = 8, k = 5, ef_search = 40;
    unsigned seed = 7;
    float *x = malloc(sizeof(float) * (size_t) n * dim), *q = malloc(sizeof(float) * (size_t) nq * dim);
    int64_t *tids = malloc(sizeof(int64_t) * nq * k), tid;
    float *dist = malloc(sizeof(float) * nq * k), d;
    int32_t *cnt = malloc(sizeof(int32_t) * nq);
    int i, j;
    for (i = 0; i < n; i++)
        for (j = 0; j < dim; j++) x[(size_t) i * dim + j] = (float) ((i % 50) * ((j % 7) - 3)) * 0.1f + frand(&seed);
    for (i = 0; i < nq; i++)
        for (j = 0; j < dim; j++) q[(size_t) i * dim + j] = x[(size_t) (i * 997) * dim + j] + 0.01f * frand(&seed);

    if (hb_device_count() <= 0) { fprintf(stderr, "no CUDA device: %s\n", hb_last_error()); return 1; }
    hb_index *ix = hb_index_create(0, dim, 16, 64, HB_L2, HB_F32, n, 1);
    if (!ix) { fprintf(stderr, "hb_index_create: %s\n", hb_last_error()); return 1; }
    if (hb_build(ix, x, n, NULL) != n) { fprintf(stderr, "hb_build: %s\n", hb_last_error()); return 1; }
    hb_index_trim(ix);                                   /* CREATE INDEX is over: free the build-only memory */

    /* one backend: amrescan + amgettuple */
    hb_scan *scan = hb_beginscan(ix);
    hb_rescan(scan, q, ef_search);
    printf("query 0 through hb_gettuple:");
    for (i = 0; i < k && hb_gettuple(scan, &tid, &d) == 1; i++) printf(" %lld (%.3f)", (long long) tid, sqrt(d));
    printf("\n");
    hb_endscan(scan);

    /* many backends at once */
    if (hb_search_batch(ix, q, nq, ef_search, k, tids, dist, cnt) != HB_OK) { fprintf(stderr, "%s\n", hb_last_error()); return 1; }
    for (i = 0; i < nq; i++) printf("query %d nearest heap TID %lld (row it was derived from: %d)\n", i, (long long) tids[i * k], i * 997);
    hb_index_free(ix);
    free(x); free(q); free(tids); free(dist); free(cnt);
    return 0;
}

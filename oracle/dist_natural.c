/*
 * dist_natural.c -- pgvector's scalar distance loops, restated (TEST INFRASTRUCTURE, see
 * hnsw_oracle.h).  Restates upstream vector.c VectorL2SquaredDistance / VectorInnerProduct and
 * halfutils.c HalfvecL2SquaredDistanceDefault / HalfvecInnerProductDefault [RECALL -- the mount
 * has no source, /root/reference/README.md:1 is the whole reference].
 *
 * This translation unit alone is compiled with the optimisation flags of pgvector's Makefile
 * (-ftree-vectorize -fassociative-math -fno-signed-zeros -fno-trapping-math), so its summation
 * order is whatever the compiler picks, exactly as for the real extension.  It is the arithmetic
 * the CPU baseline is timed with.
 */
#include <stdint.h>
#include <immintrin.h>

float orc_nat_l2_f32(int dim, const float *ax, const float *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++) {
        float diff = ax[i] - bx[i];
        distance += diff * diff;
    }
    return distance;
}

float orc_nat_ip_f32(int dim, const float *ax, const float *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++)
        distance += ax[i] * bx[i];
    return distance;
}

/* halfvec: fp16 storage, fp32 convert + accumulate. ax is the already-converted query. */
float orc_nat_l2_f16(int dim, const float *ax, const uint16_t *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++) {
        float diff = ax[i] - _cvtsh_ss(bx[i]);
        distance += diff * diff;
    }
    return distance;
}

float orc_nat_ip_f16(int dim, const float *ax, const uint16_t *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++)
        distance += ax[i] * _cvtsh_ss(bx[i]);
    return distance;
}

/* VectorL1Distance / HalfvecL1Distance (pgvector 0.7+) */
float orc_nat_l1_f32(int dim, const float *ax, const float *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++)
        distance += __builtin_fabsf(ax[i] - bx[i]);
    return distance;
}

float orc_nat_l1_f16(int dim, const float *ax, const uint16_t *bx)
{
    float distance = 0.0f;
    for (int i = 0; i < dim; i++)
        distance += __builtin_fabsf(ax[i] - _cvtsh_ss(bx[i]));
    return distance;
}

/* vector_norm / l2_normalize accumulate in double */
double orc_nat_sqnorm_f32(int dim, const float *ax)
{
    double norm = 0.0;
    for (int i = 0; i < dim; i++)
        norm += (double) ax[i] * (double) ax[i];
    return norm;
}

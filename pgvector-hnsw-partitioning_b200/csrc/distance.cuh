// distance.cuh -- warp-cooperative distance evaluation in the canonical summation order.
// Template parameter IP selects the opclass FUNCTION 1: 0 = squared L2, 1 = negative inner product
// (also cosine, on normalised rows), 2 = L1 (vector_l1_ops, pgvector 0.7 VectorL1Distance).
//
// Takes the role of pgvector's opclass FUNCTION 1 support functions: VectorL2SquaredDistance /
// VectorInnerProduct (vector.c) and HalfvecL2SquaredDistance / HalfvecInnerProduct (halfutils.c)
// [RECALL; the reference mount has no source, /root/reference/README.md:1].
//
// Canonical order (restated on the CPU by oracle/hnsw_oracle.c canon_l2/canon_ip so results can
// be compared bit-for-bit):
//   * a row is a sequence of 16-byte chunks (VEC = 4 fp32 or 8 fp16 components);
//   * lane l of the warp owns chunks l, l+32, l+64, ... and keeps VEC fp32 accumulators, one per
//     component slot, each an FMA chain in increasing chunk order;
//   * the lane's accumulators fold pairwise: (a0+a1)+(a2+a3) [+ the same for a4..a7];
//   * the 32 lane partials fold by the xor-butterfly 16, 8, 4, 2, 1.
// Several candidates are evaluated per pass (G rows in flight per lane as 128-bit loads) and
// reduced together by a transposed butterfly, which performs the same additions in the same order.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace hb {

constexpr unsigned FULL = 0xffffffffu;

// 128-bit streaming load of vector data: read-only path, do not allocate in L1 (rows are touched
// once per query; L2 keeps the hubs).
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int VEC = 4; };
template <> struct Vec<__half> { static constexpr int VEC = 8; };

// Query layout in shared memory: fp32 rows: q[4*ch + k]. fp16 rows: two planes so that each
// lane's 128-bit read is bank-conflict free: q[4*ch + k] (k<4) and q[plane + 4*ch + (k-4)].
template <typename T> __device__ __forceinline__ int query_floats(int nvec) { return nvec * Vec<T>::VEC; }

// ---- packed fp32 pairs --------------------------------------------------------------------------
// sm_100 has two-wide fp32 instructions (PTX add/sub/fma.rn.f32x2 -> FADD2 / FFMA2): two independent IEEE
// operations per issue slot, so results are bit-identical to the scalar forms.  Accumulators are kept as
// pairs: pair p of a lane holds component slots 2p and 2p+1.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pack2u(uint32_t lo, uint32_t hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 abs2(f32x2 a) { return a & 0x7fffffff7fffffffull; }

template <int IP> __device__ __forceinline__ f32x2 accum_pair(f32x2 acc, f32x2 q, f32x2 v)
{
    if constexpr (IP == 1) return fma2(q, v, acc);
    else if constexpr (IP == 2) return add2(acc, abs2(sub2(q, v)));
    else {
        const f32x2 t = sub2(q, v);
        return fma2(t, t, acc);
    }
}

// one 16-byte chunk of a row against the matching query components (already in registers)
template <typename T, int IP>
__device__ __forceinline__ void accum_chunk(f32x2 (&acc)[Vec<T>::VEC / 2], const uint4 &raw, const float4 &qa,
                                            const float4 &qb)
{
    if constexpr (sizeof(T) == 4) {
        acc[0] = accum_pair<IP>(acc[0], pack2(qa.x, qa.y), pack2u(raw.x, raw.y));
        acc[1] = accum_pair<IP>(acc[1], pack2(qa.z, qa.w), pack2u(raw.z, raw.w));
    } else {
        const float2 h0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
        const float2 h1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
        const float2 h2 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.z));
        const float2 h3 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.w));
        acc[0] = accum_pair<IP>(acc[0], pack2(qa.x, qa.y), pack2(h0.x, h0.y));
        acc[1] = accum_pair<IP>(acc[1], pack2(qa.z, qa.w), pack2(h1.x, h1.y));
        acc[2] = accum_pair<IP>(acc[2], pack2(qb.x, qb.y), pack2(h2.x, h2.y));
        acc[3] = accum_pair<IP>(acc[3], pack2(qb.z, qb.w), pack2(h3.x, h3.y));
    }
}

template <int VEC> __device__ __forceinline__ float fold_lane(const float (&a)[VEC])
{
    if constexpr (VEC == 4) return (a[0] + a[1]) + (a[2] + a[3]);
    else return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}
// the same fold over packed accumulators
template <int VEC> __device__ __forceinline__ float fold_lane2(const f32x2 (&p)[VEC / 2])
{
    float a[VEC];
#pragma unroll
    for (int k = 0; k < VEC / 2; k++) unpack2(p[k], a[2 * k], a[2 * k + 1]);
    return fold_lane<VEC>(a);
}

// Lane partials of G candidate rows against the query in shared memory.
// NV > 0: the row is exactly 32*NV chunks (every lane owns NV chunks; loads are issued
// back-to-back from one base pointer per row with immediate offsets, G*NV 128-bit loads in flight
// per lane); NV == 0: run-time loop for any row length.
// SMEM: `vecs` is an array of rows staged in shared memory and ids are its slot numbers (same arithmetic, other loads).
template <bool SMEM> __device__ __forceinline__ uint4 ld_row_chunk(const void *p)
{
    if constexpr (SMEM) return *reinterpret_cast<const uint4 *>(p);
    else return ldg_stream(p);
}

template <typename T, int IP, int NV, int G, bool SMEM = false>
__device__ __forceinline__ void group_partials(const char *__restrict__ vecs, uint32_t row_bytes, int nvec,
                                               const float *q, const int32_t (&ids)[G], int lane,
                                               float (&part)[G])
{
    constexpr int VEC = Vec<T>::VEC;
    constexpr bool HALF = sizeof(T) == 2;
    f32x2 acc[G][VEC / 2];
#pragma unroll
    for (int c = 0; c < G; c++)
#pragma unroll
        for (int k = 0; k < VEC / 2; k++) acc[c][k] = 0ull;
    // row address = (base + this lane's chunk) + id * row_bytes: one multiply-add with a 64-bit addend per row
    const char *lane_base = vecs + 16 * lane;
    const char *rp[G];
#pragma unroll
    for (int c = 0; c < G; c++) rp[c] = lane_base + (uint64_t) (uint32_t) ids[c] * row_bytes;
    const float *qp = q + 4 * lane;

    if constexpr (NV > 0) {
        uint4 raw[G][NV];
#pragma unroll
        for (int c = 0; c < G; c++)
#pragma unroll
            for (int j = 0; j < NV; j++) raw[c][j] = ld_row_chunk<SMEM>(rp[c] + 512 * j);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const float4 qa = *reinterpret_cast<const float4 *>(qp + 128 * j);
            float4 qb = qa;
            if constexpr (HALF) qb = *reinterpret_cast<const float4 *>(qp + 128 * (NV + j));
#pragma unroll
            for (int c = 0; c < G; c++) accum_chunk<T, IP>(acc[c], raw[c][j], qa, qb);
        }
    } else {
        const int plane = 4 * nvec;
        for (int ch = lane; ch < nvec; ch += 32) {
            uint4 raw[G];
#pragma unroll
            for (int c = 0; c < G; c++) raw[c] = ld_row_chunk<SMEM>(rp[c] + 16 * (ch - lane));
            const float4 qa = *reinterpret_cast<const float4 *>(q + 4 * ch);
            float4 qb = qa;
            if constexpr (HALF) qb = *reinterpret_cast<const float4 *>(q + plane + 4 * ch);
#pragma unroll
            for (int c = 0; c < G; c++) accum_chunk<T, IP>(acc[c], raw[c], qa, qb);
        }
    }
#pragma unroll
    for (int c = 0; c < G; c++) part[c] = fold_lane2<VEC>(acc[c]);
}

// Transposed butterfly: reduces G per-lane partials across the warp with the additions of the
// canonical butterfly (16, 8, 4, 2, 1).  On return every lane holds the total of candidate
// c = lane / (32 / G).
template <int G> struct XReduce {
    __device__ __forceinline__ static float run(const float (&p)[G], int lane, int bit)
    {
        constexpr int H = G / 2;
        const bool hi = (lane & bit) != 0;
        float keep[H];
#pragma unroll
        for (int i = 0; i < H; i++) {
            const float send = hi ? p[i] : p[i + H];
            const float mine = hi ? p[i + H] : p[i];
            keep[i] = mine + __shfl_xor_sync(FULL, send, bit);
        }
        return XReduce<H>::run(keep, lane, bit >> 1);
    }
};
template <> struct XReduce<1> {
    __device__ __forceinline__ static float run(const float (&p)[1], int lane, int bit)
    {
        float s = p[0];
        for (int b = bit; b >= 1; b >>= 1) s = s + __shfl_xor_sync(FULL, s, b);
        return s;
    }
};

// distances of G candidates; result of candidate c is returned in every lane via out[c]
template <typename T, int IP, int NV, int G, bool SMEM = false>
__device__ __forceinline__ float group_distance(const char *__restrict__ vecs, uint32_t row_bytes, int nvec,
                                                const float *q, const int32_t (&ids)[G], int lane)
{
    float part[G];
    group_partials<T, IP, NV, G, SMEM>(vecs, row_bytes, nvec, q, ids, lane, part);
    const float s = XReduce<G>::run(part, lane, 16);
    return IP == 1 ? -s : s;   // lane (c * 32/G) .. hold candidate c
}

// Two queries at once: each candidate row is fetched once and accumulated against both staged
// queries (same additions per (query, row) pair as group_distance).  Lane c*(32/G).. holds candidate
// c's distance to q0 in out0 and to q1 in out1.
template <typename T, int IP, int NV, int G>
__device__ __forceinline__ void group_distance2(const char *__restrict__ vecs, uint32_t row_bytes, int nvec,
                                                const float *q0, const float *q1, const int32_t (&ids)[G], int lane,
                                                float &out0, float &out1)
{
    constexpr int VEC = Vec<T>::VEC;
    constexpr bool HALF = sizeof(T) == 2;
    f32x2 acc0[G][VEC / 2], acc1[G][VEC / 2];
#pragma unroll
    for (int c = 0; c < G; c++)
#pragma unroll
        for (int k = 0; k < VEC / 2; k++) { acc0[c][k] = 0ull; acc1[c][k] = 0ull; }
    const char *rp[G];
#pragma unroll
    for (int c = 0; c < G; c++) rp[c] = vecs + (uint64_t) (uint32_t) ids[c] * row_bytes + 16 * lane;
    const int plane = 4 * nvec;
    if constexpr (NV > 0) {
        uint4 raw[G][NV];
#pragma unroll
        for (int c = 0; c < G; c++)
#pragma unroll
            for (int j = 0; j < NV; j++) raw[c][j] = ldg_stream(rp[c] + 512 * j);
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int ch = lane + 32 * j;
            const float4 a0 = *reinterpret_cast<const float4 *>(q0 + 4 * ch);
            const float4 a1 = *reinterpret_cast<const float4 *>(q1 + 4 * ch);
            float4 b0 = a0, b1 = a1;
            if constexpr (HALF) {
                b0 = *reinterpret_cast<const float4 *>(q0 + plane + 4 * ch);
                b1 = *reinterpret_cast<const float4 *>(q1 + plane + 4 * ch);
            }
#pragma unroll
            for (int c = 0; c < G; c++) {
                accum_chunk<T, IP>(acc0[c], raw[c][j], a0, b0);
                accum_chunk<T, IP>(acc1[c], raw[c][j], a1, b1);
            }
        }
    } else {
        for (int ch = lane; ch < nvec; ch += 32) {
            uint4 raw[G];
#pragma unroll
            for (int c = 0; c < G; c++) raw[c] = ldg_stream(rp[c] + 16 * (ch - lane));
            const float4 a0 = *reinterpret_cast<const float4 *>(q0 + 4 * ch);
            const float4 a1 = *reinterpret_cast<const float4 *>(q1 + 4 * ch);
            float4 b0 = a0, b1 = a1;
            if constexpr (HALF) {
                b0 = *reinterpret_cast<const float4 *>(q0 + plane + 4 * ch);
                b1 = *reinterpret_cast<const float4 *>(q1 + plane + 4 * ch);
            }
#pragma unroll
            for (int c = 0; c < G; c++) {
                accum_chunk<T, IP>(acc0[c], raw[c], a0, b0);
                accum_chunk<T, IP>(acc1[c], raw[c], a1, b1);
            }
        }
    }
    float part0[G], part1[G];
#pragma unroll
    for (int c = 0; c < G; c++) { part0[c] = fold_lane2<VEC>(acc0[c]); part1[c] = fold_lane2<VEC>(acc1[c]); }
    const float s0 = XReduce<G>::run(part0, lane, 16), s1 = XReduce<G>::run(part1, lane, 16);
    out0 = IP == 1 ? -s0 : s0;
    out1 = IP == 1 ? -s1 : s1;
}

// distance between two rows staged in shared memory (stage_row layout), canonical order
template <typename T, int IP>
__device__ __forceinline__ float staged_pair_distance(const float *q0, const float *q1, int nvec, int lane)
{
    constexpr int VEC = Vec<T>::VEC;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; k++) acc[k] = 0.0f;
    const int plane = 4 * nvec;
    for (int ch = lane; ch < nvec; ch += 32) {
        float a[VEC], b[VEC];
        const float4 a0 = *reinterpret_cast<const float4 *>(q0 + 4 * ch), b0 = *reinterpret_cast<const float4 *>(q1 + 4 * ch);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        if constexpr (VEC == 8) {
            const float4 a1 = *reinterpret_cast<const float4 *>(q0 + plane + 4 * ch), b1 = *reinterpret_cast<const float4 *>(q1 + plane + 4 * ch);
            a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            if constexpr (IP == 1) acc[k] = fmaf(a[k], b[k], acc[k]);
            else if constexpr (IP == 2) acc[k] = acc[k] + fabsf(a[k] - b[k]);
            else { const float t = a[k] - b[k]; acc[k] = fmaf(t, t, acc[k]); }
        }
    }
    float s = fold_lane<VEC>(acc);
    for (int b = 16; b >= 1; b >>= 1) s = s + __shfl_xor_sync(FULL, s, b);
    return IP == 1 ? -s : s;
}

// stage a query (global, index dtype, `dim` components) into shared memory as fp32 in the
// layout accum_chunk expects; pads with zeros up to the chunk boundary.
template <typename T>
__device__ __forceinline__ void stage_query(const T *__restrict__ src, int dim, int nvec, float *q, int lane)
{
    constexpr int VEC = Vec<T>::VEC;
    const int total = nvec * VEC;
    for (int e = lane; e < total; e += 32) {
        float v = 0.0f;
        if (e < dim) {
            if constexpr (sizeof(T) == 4) v = src[e];
            else v = __half2float(src[e]);
        }
        if constexpr (sizeof(T) == 4) q[e] = v;
        else {
            const int ch = e >> 3, k = e & 7;
            q[(k < 4 ? 0 : 4 * nvec) + 4 * ch + (k & 3)] = v;
        }
    }
}

// stage a row that lives in the index (padded to whole 16-byte chunks, 16-byte aligned) as the query:
// 128-bit loads, same shared-memory layout as stage_query
template <typename T>
__device__ __forceinline__ void stage_row(const char *__restrict__ row, int nvec, float *q, int lane)
{
    for (int ch = lane; ch < nvec; ch += 32) {
        const uint4 raw = *reinterpret_cast<const uint4 *>(row + 16 * ch);
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<uint4 *>(q + 4 * ch) = raw;
        } else {
            const float2 h0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.x));
            const float2 h1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.y));
            const float2 h2 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.z));
            const float2 h3 = __half22float2(*reinterpret_cast<const __half2 *>(&raw.w));
            *reinterpret_cast<float4 *>(q + 4 * ch) = make_float4(h0.x, h0.y, h1.x, h1.y);
            *reinterpret_cast<float4 *>(q + 4 * nvec + 4 * ch) = make_float4(h2.x, h2.y, h3.x, h3.y);
        }
    }
}

// same, with the row's 128-bit loads issued back to back before any store (row of exactly 32*NV chunks)
template <typename T, int NV>
__device__ __forceinline__ void stage_row_nv(const char *__restrict__ row, int nvec, float *q, int lane)
{
    if constexpr (NV == 0) stage_row<T>(row, nvec, q, lane);
    else {
        uint4 raw[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) raw[j] = *reinterpret_cast<const uint4 *>(row + 16 * (lane + 32 * j));
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const int ch = lane + 32 * j;
            if constexpr (sizeof(T) == 4) *reinterpret_cast<uint4 *>(q + 4 * ch) = raw[j];
            else {
                const float2 h0 = __half22float2(*reinterpret_cast<const __half2 *>(&raw[j].x));
                const float2 h1 = __half22float2(*reinterpret_cast<const __half2 *>(&raw[j].y));
                const float2 h2 = __half22float2(*reinterpret_cast<const __half2 *>(&raw[j].z));
                const float2 h3 = __half22float2(*reinterpret_cast<const __half2 *>(&raw[j].w));
                *reinterpret_cast<float4 *>(q + 4 * ch) = make_float4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<float4 *>(q + 4 * nvec + 4 * ch) = make_float4(h2.x, h2.y, h3.x, h3.y);
            }
        }
    }
}

// ask L2 for `bytes` (multiple of 16) at a 16-byte aligned address; returns immediately
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

}   // namespace hb

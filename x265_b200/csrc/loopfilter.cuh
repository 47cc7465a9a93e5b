// x265_b200/csrc/loopfilter.cuh -- in-loop filter class: deblocking edge filters and SAO.
// Semantics: /root/reference/source/common/loopfilter.cpp:39-200 (calSign, processSaoCUE0/E1/E1_2Rows/E2/E3/B0,
// pelFilterLumaStrong_c, pelFilterChroma_c) and /root/reference/source/encoder/sao.cpp:1762-1927 (saoCuStatsBO/E0..E3).
// The reference walks every row left to right and carries the edge sign of the previous sample / row in a scalar or in the
// upBuff arrays; every sign is a function of ORIGINAL samples only (a sample is modified after its last use as a neighbour),
// so all kernels read an unmodified input block and write a separate output block: one thread per sample, rows that pass
// signs to the next row (E1 / E2 / E3 and the statistics) synchronise once per row.
// Blocks are staged compactly by the per-call thunks (thunks.cuh); each kernel documents its layout.
#pragma once
#include "common.cuh"

namespace lf {

__device__ __forceinline__ int sgn(int x) { return (x > 0) - (x < 0); }
template <typename P> __device__ __forceinline__ int clipP(int v) { return min(max(v, 0), PixTraits<P>::maxv); }

// calSign (loopfilter.cpp:39-43)
template <typename P>
__global__ void k_sign(int8_t* __restrict__ dst, const P* __restrict__ a, const P* __restrict__ b, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (int8_t)sgn((int)a[i] - (int)b[i]);
}

// processSaoCUE0 (:45-63).  in: 2 rows x (width + 1) samples, pitch = width + 1.  out: 2 rows x width.
template <typename P>
__global__ void k_sao_e0(const P* __restrict__ in, P* __restrict__ out, const int8_t* __restrict__ off, int width, const int8_t* __restrict__ signLeft)
{
    for (int i = threadIdx.x; i < 2 * width; i += blockDim.x)
    {
        const int y = i / width, x = i - y * width;
        const P* r = in + y * (width + 1);
        const int c = r[x];
        const int sr = sgn(c - (int)r[x + 1]);
        const int sl = x ? sgn(c - (int)r[x - 1]) : (int)signLeft[y];       // signLeft0 = -signRight of the sample before
        out[y * width + x] = (P)clipP<P>(c + off[sr + sl + 2]);
    }
}

// processSaoCUE1 / E1_2Rows (:65-98).  in: (rows + 1) x width.  out: rows x width; up: upBuff1 in/out (width).
template <typename P>
__global__ void k_sao_e1(const P* __restrict__ in, P* __restrict__ out, int8_t* __restrict__ up, const int8_t* __restrict__ off, int width, int rows)
{
    for (int x = threadIdx.x; x < width; x += blockDim.x)
    {
        int u = up[x];
        for (int y = 0; y < rows; y++)
        {
            const int c = in[y * width + x];
            const int sd = sgn(c - (int)in[(y + 1) * width + x]);
            out[y * width + x] = (P)clipP<P>(c + off[sd + u + 2]);
            u = -sd;
        }
        up[x] = (int8_t)u;
    }
}

// processSaoCUE2 (:100-110).  in: row 0 = rec[0 .. width), row 1 = rec[stride + 1 .. stride + 1 + width).  buff1 read,
// bufft[1 .. width] written (passed as bufft + 1).
template <typename P>
__global__ void k_sao_e2(const P* __restrict__ in, P* __restrict__ out, int8_t* __restrict__ bufft1, const int8_t* __restrict__ buff1,
                         const int8_t* __restrict__ off, int width)
{
    for (int x = threadIdx.x; x < width; x += blockDim.x)
    {
        const int c = in[x];
        const int sd = sgn(c - (int)in[width + x]);
        bufft1[x] = (int8_t)(-sd);
        out[x] = (P)clipP<P>(c + off[sd + buff1[x] + 2]);
    }
}

// processSaoCUE3 (:112-124).  n = endX - startX - 1 samples x = startX + 1 ..: in row 0 = rec[x], row 1 = rec[x + stride];
// up = upBuff1 + startX (n + 1 entries): reads up[1 + i], writes up[i].
template <typename P>
__global__ void k_sao_e3(const P* __restrict__ in, P* __restrict__ out, const int8_t* __restrict__ upIn, int8_t* __restrict__ upOut,
                         const int8_t* __restrict__ off, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const int c = in[i];
        const int sd = sgn(c - (int)in[n + i]);
        out[i] = (P)clipP<P>(c + off[sd + upIn[1 + i] + 2]);
        upOut[i] = (int8_t)(-sd);
    }
}

// processSaoCUB0 (:126-138).  w x h block, compact.
template <typename P>
__global__ void k_sao_b0(const P* __restrict__ in, P* __restrict__ out, const int8_t* __restrict__ off, int n)
{
    constexpr int boShift = PixTraits<P>::depth - 5;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int c = in[i];
        out[i] = (P)clipP<P>(c + off[c >> boShift]);
    }
}

// pelFilterLumaStrong_c (:140-160) / pelFilterChroma_c (:167-181).  in / out: 4 lines x 8 samples m0..m7 (m4 = src[0]).
template <typename P>
__global__ void k_deblock_luma_strong(const P* __restrict__ in, P* __restrict__ out, int tcP, int tcQ)
{
    const int i = threadIdx.x;
    if (i >= 4) return;
    int m[8];
#pragma unroll
    for (int k = 0; k < 8; k++) m[k] = (int)(int16_t)in[i * 8 + k];
    P* o = out + i * 8;
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = (P)m[k];
    o[1] = (P)(clip3i(-tcP, tcP, ((2 * m[0] + 3 * m[1] + m[2] + m[3] + m[4] + 4) >> 3) - m[1]) + m[1]);
    o[2] = (P)(clip3i(-tcP, tcP, ((m[1] + m[2] + m[3] + m[4] + 2) >> 2) - m[2]) + m[2]);
    o[3] = (P)(clip3i(-tcP, tcP, ((m[1] + 2 * m[2] + 2 * m[3] + 2 * m[4] + m[5] + 4) >> 3) - m[3]) + m[3]);
    o[4] = (P)(clip3i(-tcQ, tcQ, ((m[2] + 2 * m[3] + 2 * m[4] + 2 * m[5] + m[6] + 4) >> 3) - m[4]) + m[4]);
    o[5] = (P)(clip3i(-tcQ, tcQ, ((m[3] + m[4] + m[5] + m[6] + 2) >> 2) - m[5]) + m[5]);
    o[6] = (P)(clip3i(-tcQ, tcQ, ((m[3] + m[4] + m[5] + 3 * m[6] + 2 * m[7] + 4) >> 3) - m[6]) + m[6]);
}
template <typename P>
__global__ void k_deblock_chroma(const P* __restrict__ in, P* __restrict__ out, int tc, int maskP, int maskQ)
{
    const int i = threadIdx.x;
    if (i >= 4) return;
    const int m2 = (int)(int16_t)in[i * 8 + 2], m3 = (int)(int16_t)in[i * 8 + 3], m4 = (int)(int16_t)in[i * 8 + 4], m5 = (int)(int16_t)in[i * 8 + 5];
    const int delta = clip3i(-tc, tc, ((((m4 - m3) * 4) + m2 - m5 + 4) >> 3));
#pragma unroll
    for (int k = 0; k < 8; k++) out[i * 8 + k] = in[i * 8 + k];
    out[i * 8 + 3] = (P)clipP<P>(m3 + (delta & maskP));
    out[i * 8 + 4] = (P)clipP<P>(m4 - (delta & maskQ));
}

// ---- SAO statistics (sao.cpp:1762-1927).  stats / count: in/out (32 classes for BO, 5 for the edge types through s_eoTable).
__constant__ int c_eoTable[5] = { 1, 2, 0, 3, 4 };                 // SAO::s_eoTable (sao.cpp)

// block accumulation helper: every thread adds into shared bins, thread 0..n-1 flush
template <int NB>
__device__ __forceinline__ void stats_flush(int* s_st, int* s_ct, int32_t* stats, int32_t* count, bool viaEoTable)
{
    __syncthreads();
    if (threadIdx.x < NB)
    {
        const int dst = viaEoTable ? c_eoTable[threadIdx.x] : threadIdx.x;
        stats[dst] += s_st[threadIdx.x]; count[dst] += s_ct[threadIdx.x];
    }
}

// saoCuStatsBO: diff pitch 64, rec compact pitch `rp`
template <typename P>
__global__ void k_sao_stats_bo(const int16_t* __restrict__ diff, const P* __restrict__ rec, int rp, int endX, int endY, int32_t* stats, int32_t* count)
{
    constexpr int boShift = PixTraits<P>::depth - 5;
    __shared__ int s_st[32], s_ct[32];
    if (threadIdx.x < 32) { s_st[threadIdx.x] = 0; s_ct[threadIdx.x] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < endX * endY; i += blockDim.x)
    {
        const int y = i / endX, x = i - y * endX;
        const int cls = rec[y * rp + x] >> boShift;
        atomicAdd(&s_st[cls], (int)diff[y * 64 + x]); atomicAdd(&s_ct[cls], 1);
    }
    stats_flush<32>(s_st, s_ct, stats, count, false);
}

// saoCuStatsE0: rec rows with one sample to the left and one to the right: compact pitch rp, row origin at column 1
template <typename P>
__global__ void k_sao_stats_e0(const int16_t* __restrict__ diff, const P* __restrict__ rec, int rp, int endX, int endY, int32_t* stats, int32_t* count)
{
    __shared__ int s_st[5], s_ct[5];
    if (threadIdx.x < 5) { s_st[threadIdx.x] = 0; s_ct[threadIdx.x] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < endX * endY; i += blockDim.x)
    {
        const int y = i / endX, x = i - y * endX;
        const P* r = rec + y * rp + 1;
        const int c = r[x];
        const int et = sgn(c - (int)r[x + 1]) + sgn(c - (int)r[x - 1]) + 2;
        atomicAdd(&s_st[et], (int)diff[y * 64 + x]); atomicAdd(&s_ct[et], 1);
    }
    stats_flush<5>(s_st, s_ct, stats, count, true);
}

// saoCuStatsE1: rec (endY + 1) rows x endX compact; up = upBuff1 in/out
template <typename P>
__global__ void k_sao_stats_e1(const int16_t* __restrict__ diff, const P* __restrict__ rec, int8_t* __restrict__ up, int endX, int endY, int32_t* stats, int32_t* count)
{
    __shared__ int s_st[5], s_ct[5];
    if (threadIdx.x < 5) { s_st[threadIdx.x] = 0; s_ct[threadIdx.x] = 0; }
    __syncthreads();
    for (int x = threadIdx.x; x < endX; x += blockDim.x)
    {
        int u = up[x];
        for (int y = 0; y < endY; y++)
        {
            const int sd = sgn((int)rec[y * endX + x] - (int)rec[(y + 1) * endX + x]);
            atomicAdd(&s_st[sd + u + 2], (int)diff[y * 64 + x]); atomicAdd(&s_ct[sd + u + 2], 1);
            u = -sd;
        }
        up[x] = (int8_t)u;
    }
    stats_flush<5>(s_st, s_ct, stats, count, true);
}

// saoCuStatsE2: rec block rows 0 .. endY, columns -1 .. endX (pitch rp = endX + 2, origin at column 1).
// bufA / bufB: the caller's upBuff1 / upBufft INCLUDING their [-1] element (so index 1 + x); the reference swaps the two
// pointers after every row: row y reads from buffer (y & 1 ? B : A) and writes into the other one.
template <typename P>
__global__ void k_sao_stats_e2(const int16_t* __restrict__ diff, const P* __restrict__ rec, int rp, int8_t* __restrict__ bufA, int8_t* __restrict__ bufB,
                               int endX, int endY, int32_t* stats, int32_t* count)
{
    __shared__ int s_st[5], s_ct[5];
    __shared__ int8_t s_b[2][80];
    if (threadIdx.x < 5) { s_st[threadIdx.x] = 0; s_ct[threadIdx.x] = 0; }
    for (int i = threadIdx.x; i < endX + 2; i += blockDim.x) { s_b[0][i] = bufA[i]; s_b[1][i] = bufB[i]; }
    __syncthreads();
    for (int y = 0; y < endY; y++)
    {
        int8_t* b1 = s_b[y & 1] + 1;       // upBuff1 of this row
        int8_t* bt = s_b[(y & 1) ^ 1] + 1; // upBufft of this row
        const P* r = rec + y * rp + 1;
        if (threadIdx.x == 0) bt[0] = (int8_t)sgn((int)r[rp] - (int)r[-1]);
        for (int x = threadIdx.x; x < endX; x += blockDim.x)
        {
            const int sd = sgn((int)r[x] - (int)r[x + rp + 1]);
            const int et = sd + b1[x] + 2;
            atomicAdd(&s_st[et], (int)diff[y * 64 + x]); atomicAdd(&s_ct[et], 1);
            // bt[x + 1] is written here and bt[0] above: distinct entries; b1 is only read in this row
            bt[x + 1] = (int8_t)(-sd);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < endX + 2; i += blockDim.x) { bufA[i] = s_b[0][i]; bufB[i] = s_b[1][i]; }
    stats_flush<5>(s_st, s_ct, stats, count, true);
}

// saoCuStatsE3: rec block rows 0 .. endY, columns -1 .. endX (pitch rp, origin at column 1); up = upBuff1 INCLUDING [-1].
template <typename P>
__global__ void k_sao_stats_e3(const int16_t* __restrict__ diff, const P* __restrict__ rec, int rp, int8_t* __restrict__ up, int endX, int endY,
                               int32_t* stats, int32_t* count)
{
    __shared__ int s_st[5], s_ct[5];
    __shared__ int8_t s_u[2][80];
    if (threadIdx.x < 5) { s_st[threadIdx.x] = 0; s_ct[threadIdx.x] = 0; }
    for (int i = threadIdx.x; i < endX + 1; i += blockDim.x) { s_u[0][i] = up[i]; s_u[1][i] = up[i]; }
    __syncthreads();
    for (int y = 0; y < endY; y++)
    {
        const int8_t* cur = s_u[y & 1] + 1;
        int8_t* nxt = s_u[(y & 1) ^ 1] + 1;
        const P* r = rec + y * rp + 1;
        for (int x = threadIdx.x; x < endX; x += blockDim.x)
        {
            const int sd = sgn((int)r[x] - (int)r[x + rp - 1]);
            const int et = sd + cur[x] + 2;
            atomicAdd(&s_st[et], (int)diff[y * 64 + x]); atomicAdd(&s_ct[et], 1);
            nxt[x - 1] = (int8_t)(-sd);                            // upBuff1[x - 1] for the next row
        }
        if (threadIdx.x == 0) nxt[endX - 1] = (int8_t)sgn((int)r[endX - 1 + rp] - (int)r[endX]);
        __syncthreads();
    }
    for (int i = threadIdx.x; i < endX + 1; i += blockDim.x) up[i] = s_u[endY & 1][i];
    stats_flush<5>(s_st, s_ct, stats, count, true);
}

} // namespace lf

// x265_b200/csrc/pixelcmp.cuh -- pixel-compare class: SAD / SATD / SA8D / SSE / SSD_S / VAR / PSY.
// Semantics: /root/reference/source/common/pixel.cpp:40-377 (sad, satd_4x4/8x4, sa8d), :167-186 (sse),
// :379-391 (ssd_s), :703-720 (var), :726-757 (psyCost_pp).  Bit-exact integer arithmetic.
#pragma once
#include "common.cuh"

// ---- Hadamard cores (per thread, registers only) -------------------------------------------
__device__ __forceinline__ void had4(int& a, int& b, int& c, int& d)
{
    int s01 = a + b, d01 = a - b, s23 = c + d, d23 = c - d;
    a = s01 + s23; c = s01 - s23; b = d01 + d23; d = d01 - d23;
}

// sum |H4 D H4^T| of one 4x4 tile, un-normalised (pixel.cpp:210-236 before the >>1)
template <typename PA, typename PB>
__device__ __forceinline__ int had4x4_abs(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb)
{
    int m[4][4];
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
#pragma unroll
        for (int x = 0; x < 4; x++) m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
        had4(m[y][0], m[y][1], m[y][2], m[y][3]);
    }
    int acc = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        had4(m[0][x], m[1][x], m[2][x], m[3][x]);
        acc += abs(m[0][x]) + abs(m[1][x]) + abs(m[2][x]) + abs(m[3][x]);
    }
    return acc;
}

// same, but plane b is implicit zero (psy cost energy term)
template <typename PA>
__device__ __forceinline__ int had4x4_abs_zero(const PA* __restrict__ a, int sa)
{
    int m[4][4];
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
#pragma unroll
        for (int x = 0; x < 4; x++) m[y][x] = (int)a[y * sa + x];
        had4(m[y][0], m[y][1], m[y][2], m[y][3]);
    }
    int acc = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        had4(m[0][x], m[1][x], m[2][x], m[3][x]);
        acc += abs(m[0][x]) + abs(m[1][x]) + abs(m[2][x]) + abs(m[3][x]);
    }
    return acc;
}

// un-normalised 8x8 Hadamard abs-sum (pixel.cpp:299-334); `zero_b` skips plane b
template <typename PA, typename PB, bool ZERO_B>
__device__ __forceinline__ int had8x8_abs(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb)
{
    int m[8][8];
#pragma unroll
    for (int y = 0; y < 8; y++)
    {
#pragma unroll
        for (int x = 0; x < 8; x++) m[y][x] = ZERO_B ? (int)a[y * sa + x] : (int)a[y * sa + x] - (int)b[y * sb + x];
        had4(m[y][0], m[y][1], m[y][2], m[y][3]);
        had4(m[y][4], m[y][5], m[y][6], m[y][7]);
#pragma unroll
        for (int k = 0; k < 4; k++) { int p = m[y][k], q = m[y][k + 4]; m[y][k] = p + q; m[y][k + 4] = p - q; }
    }
    int acc = 0;
#pragma unroll
    for (int x = 0; x < 8; x++)
    {
        had4(m[0][x], m[1][x], m[2][x], m[3][x]);
        had4(m[4][x], m[5][x], m[6][x], m[7][x]);
#pragma unroll
        for (int k = 0; k < 4; k++) acc += abs(m[k][x] + m[k + 4][x]) + abs(m[k][x] - m[k + 4][x]);
    }
    return acc;
}

// ---- warp-cooperative block costs (all 32 lanes call; result valid in every lane) ----------

template <typename PA, typename PB>
__device__ __forceinline__ int warp_sad(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb, int w, int h, int lane)
{
    int acc = 0;
    const int n = w * h;
    if ((w & (w - 1)) == 0)
    {
        const int lg = 31 - __clz(w);
        for (int i = lane; i < n; i += 32)
        {
            int y = i >> lg, x = i & (w - 1);
            acc += abs((int)a[y * sa + x] - (int)b[y * sb + x]);
        }
    }
    else
    {
        for (int i = lane; i < n; i += 32)
        {
            int y = i / w, x = i - y * w;
            acc += abs((int)a[y * sa + x] - (int)b[y * sb + x]);
        }
    }
    return warp_sum(acc);
}

// SATD with the reference's tiling (pixel.cpp:263-297, table :1134-1158): width % 8 == 0 -> 8x4
// tiles halved once per tile; otherwise 4x4 tiles halved per tile.
template <typename PA, typename PB>
__device__ __forceinline__ int warp_satd(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb, int w, int h, int lane)
{
    int acc = 0;
    if ((w & 7) == 0)
    {
        const int tw = w >> 3, nt = tw * (h >> 2);
        for (int t = lane; t < nt; t += 32)
        {
            int ty = t / tw, tx = t - ty * tw;
            const PA* pa = a + (ty * 4) * sa + tx * 8; const PB* pb = b + (ty * 4) * sb + tx * 8;
            acc += (had4x4_abs(pa, sa, pb, sb) + had4x4_abs(pa + 4, sa, pb + 4, sb)) >> 1;
        }
    }
    else
    {
        const int tw = w >> 2, nt = tw * (h >> 2);
        for (int t = lane; t < nt; t += 32)
        {
            int ty = t / tw, tx = t - ty * tw;
            acc += had4x4_abs(a + (ty * 4) * sa + tx * 4, sa, b + (ty * 4) * sb + tx * 4, sb) >> 1;
        }
    }
    return warp_sum(acc);
}

// SA8D (pixel.cpp:336-377, table :1166-1170 and the chroma aliases): <8 -> satd; multiples of 16 ->
// 16x16 tiles rounded once; otherwise 8x8 tiles each rounded.
template <typename PA, typename PB>
__device__ __forceinline__ int warp_sa8d(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb, int w, int h, int lane)
{
    if (w < 8 || h < 8) return warp_satd(a, sa, b, sb, w, h, lane);
    const int tw = w >> 3, nt = tw * (h >> 3);
    int acc = 0;
    if (((w | h) & 15) == 0)
    {
        // one lane per 8x8; the 4 lanes of a 16x16 tile are combined before rounding
        for (int t0 = 0; t0 < nt; t0 += 32)
        {
            int t = t0 + lane;
            int v = 0;
            // order 8x8 tiles so that 4 consecutive t form one 16x16: t = (tile16 * 4 + sub)
            if (t < nt)
            {
                int t16 = t >> 2, sub = t & 3;
                int tw16 = w >> 4;
                int y16 = t16 / tw16, x16 = t16 - y16 * tw16;
                int y = y16 * 16 + (sub >> 1) * 8, x = x16 * 16 + (sub & 1) * 8;
                v = had8x8_abs<PA, PB, false>(a + y * sa + x, sa, b + y * sb + x, sb);
            }
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if ((lane & 3) == 0 && t < nt) acc += (v + 2) >> 2;
        }
    }
    else
    {
        for (int t = lane; t < nt; t += 32)
        {
            int ty = t / tw, tx = t - ty * tw;
            acc += (had8x8_abs<PA, PB, false>(a + ty * 8 * sa + tx * 8, sa, b + ty * 8 * sb + tx * 8, sb) + 2) >> 2;
        }
    }
    return warp_sum(acc);
}

template <typename PA, typename PB>
__device__ __forceinline__ unsigned long long warp_sse(const PA* __restrict__ a, int sa, const PB* __restrict__ b, int sb, int w, int h, int lane)
{
    unsigned long long acc = 0;
    const int n = w * h;
    for (int i = lane; i < n; i += 32)
    {
        int y = i / w, x = i - y * w;
        int d = (int)a[y * sa + x] - (int)b[y * sb + x];
        acc += (unsigned)(d * d);
    }
    return warp_sum64(acc);
}

// psyCost_pp (pixel.cpp:726-757)
template <typename P>
__device__ __forceinline__ int warp_psy(const P* __restrict__ s, int ss, const P* __restrict__ r, int rs, int n, int lane)
{
    if (n == 4)
    {
        int v = 0;
        if (lane == 0)
        {
            int sadS = 0, sadR = 0;
            for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) { sadS += s[y * ss + x]; sadR += r[y * rs + x]; }
            int es = (had4x4_abs_zero(s, ss) >> 1) - (sadS >> 2);
            int er = (had4x4_abs_zero(r, rs) >> 1) - (sadR >> 2);
            v = abs(es - er);
        }
        return __shfl_sync(0xffffffffu, v, 0);
    }
    const int tw = n >> 3, nt = tw * tw;
    int acc = 0;
    for (int t = lane; t < nt; t += 32)
    {
        int ty = t / tw, tx = t - ty * tw;
        const P* ps = s + ty * 8 * ss + tx * 8; const P* pr = r + ty * 8 * rs + tx * 8;
        int sadS = 0, sadR = 0;
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) { sadS += ps[y * ss + x]; sadR += pr[y * rs + x]; }
        int es = ((had8x8_abs<P, P, true>(ps, ss, ps, ss) + 2) >> 2) - (sadS >> 2);
        int er = ((had8x8_abs<P, P, true>(pr, rs, pr, rs) + 2) >> 2) - (sadR >> 2);
        acc += abs(es - er);
    }
    return warp_sum(acc);
}

// ---- batched generic kernel: one warp per job ---------------------------------------------
template <typename P>
__global__ void __launch_bounds__(256) k_pixelcmp(int op, const P* __restrict__ A, const P* __restrict__ B,
                                                  const x265cu_cmp_job* __restrict__ jobs, int n, uint64_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb)
    {
        const x265cu_cmp_job job = jobs[j];
        const int w = job.w, h = job.h, sa = job.a_stride, sb = job.b_stride;
        uint64_t res = 0;
        switch (op)
        {
        case X265CU_SAD:  res = (uint32_t)warp_sad(A + job.a_off, sa, B + job.b_off, sb, w, h, lane); break;
        case X265CU_SATD: res = (uint32_t)warp_satd(A + job.a_off, sa, B + job.b_off, sb, w, h, lane); break;
        case X265CU_SA8D: res = (uint32_t)warp_sa8d(A + job.a_off, sa, B + job.b_off, sb, w, h, lane); break;
        case X265CU_SSE_PP:
        {
            unsigned long long v = warp_sse(A + job.a_off, sa, B + job.b_off, sb, w, h, lane);
            res = PixTraits<P>::depth == 8 ? (uint64_t)(uint32_t)v : v;     // sse_t width (common.h:144-148)
            break;
        }
        case X265CU_SSE_SS:
        {
            const int16_t* a = (const int16_t*)A + job.a_off; const int16_t* b = (const int16_t*)B + job.b_off;
            unsigned long long v = warp_sse(a, sa, b, sb, w, h, lane);
            res = PixTraits<P>::depth == 8 ? (uint64_t)(uint32_t)v : v;
            break;
        }
        case X265CU_SSD_S:
        {
            const int16_t* a = (const int16_t*)A + job.a_off;
            unsigned long long acc = 0;
            for (int i = lane; i < w * h; i += 32) { int y = i / w, x = i - y * w; int v = a[y * sa + x]; acc += (unsigned)(v * v); }
            acc = warp_sum64(acc);
            res = PixTraits<P>::depth == 8 ? (uint64_t)(uint32_t)acc : acc;
            break;
        }
        case X265CU_VAR:
        {
            const P* a = A + job.a_off;
            unsigned s = 0, q = 0;
            for (int i = lane; i < w * h; i += 32) { int y = i / w, x = i - y * w; unsigned v = a[y * sa + x]; s += v; q += v * v; }
            s = (unsigned)warp_sum((int)s); q = (unsigned)warp_sum((int)q);
            res = (uint64_t)s + ((uint64_t)q << 32);
            break;
        }
        case X265CU_PSY: res = (uint32_t)warp_psy(A + job.a_off, sa, B + job.b_off, sb, w, lane); break;
        }
        if (lane == 0) out[j] = res;
    }
}

static int launch_pixelcmp(x265cu_ctx* ctx, int depth, int op, const void* A, const void* B,
                           const x265cu_cmp_job* jobs, int n, uint64_t* out)
{
    if (n <= 0) return 0;
    const int threads = 256, wpb = threads / 32;
    int blocks = (n + wpb - 1) / wpb;
    int maxb = ctx->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (depth == 8)
        k_pixelcmp<uint8_t><<<blocks, threads, 0, ctx->stream>>>(op, (const uint8_t*)A, (const uint8_t*)B, jobs, n, out);
    else
        k_pixelcmp<uint16_t><<<blocks, threads, 0, ctx->stream>>>(op, (const uint16_t*)A, (const uint16_t*)B, jobs, n, out);
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

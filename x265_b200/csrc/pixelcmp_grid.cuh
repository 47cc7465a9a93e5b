// x265_b200/csrc/pixelcmp_grid.cuh -- pixel-compare class over a REGULAR GRID of equal blocks
// (plane A vs plane B, B's displacement baked into its base pointer): the shape the lookahead and the
// frame-level cost passes have.  Same semantics as pixelcmp.cuh (pixel.cpp:40-377), no job list.
//
// Mapping: one lane owns a 16-byte wide column segment ("unit": 16 px at 8 bit, 8 px at 16 bit) of one
// block row, so consecutive lanes read consecutive 16-byte chunks of a plane row (fully coalesced),
// walks the block's rows keeping everything in registers (a unit holds whole 4x4 / 8x4 / 8x8 Hadamard
// tiles, so SATD / SA8D need no shuffles at 8 bit), and the lanes of one block fold with xor-shuffles.
// Algorithmic bytes per block: 2*bw*bh*sizeof(pixel) read + 8 written.
#pragma once
#include "pixelcmp.cuh"

template <bool ALIGNED>
__device__ __forceinline__ uint4 grid_load16(const uint8_t* p)
{
    if (ALIGNED) return __ldg((const uint4*)p);
    const uint32_t* ap = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
    const unsigned sh = ((unsigned)(uintptr_t)p & 3u) * 8u;
    const uint32_t w0 = __ldg(ap), w1 = __ldg(ap + 1), w2 = __ldg(ap + 2), w3 = __ldg(ap + 3);
    const uint32_t w4 = sh ? __ldg(ap + 4) : 0u;
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// pixel x (0..3 at 8 bit, 0..1 at 16 bit) of one 32-bit word
template <typename P> __device__ __forceinline__ int grid_px(uint32_t w, int x);
template <> __device__ __forceinline__ int grid_px<uint8_t>(uint32_t w, int x) { return (int)((w >> (8 * x)) & 255u); }
template <> __device__ __forceinline__ int grid_px<uint16_t>(uint32_t w, int x) { return (int)((w >> (16 * x)) & 65535u); }

__device__ __forceinline__ uint32_t u4w(const uint4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// difference of the 4 pixels starting at pixel column c (multiple of 4) of a unit
template <typename P>
__device__ __forceinline__ void grid_diff4(const uint4& a, const uint4& b, int c, int d[4])
{
    if (sizeof(P) == 1)
    {
        const uint32_t wa = u4w(a, c >> 2), wb = u4w(b, c >> 2);
#pragma unroll
        for (int x = 0; x < 4; x++) d[x] = grid_px<P>(wa, x) - grid_px<P>(wb, x);
    }
    else
    {
#pragma unroll
        for (int x = 0; x < 4; x++) d[x] = grid_px<P>(u4w(a, (c + x) >> 1), x & 1) - grid_px<P>(u4w(b, (c + x) >> 1), x & 1);
    }
}

// un-normalised 4x4 Hadamard abs-sum of rows r[0..3], pixel columns c..c+3
template <typename P>
__device__ __forceinline__ int grid_had4x4(const uint4 (&ra)[4], const uint4 (&rb)[4], int c)
{
    int m[4][4];
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        grid_diff4<P>(ra[y], rb[y], c, m[y]);
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

// un-normalised 8x8 Hadamard abs-sum of 8 rows, pixel columns c..c+7
template <typename P>
__device__ __forceinline__ int grid_had8x8(const uint4 (&ra)[8], const uint4 (&rb)[8], int c)
{
    int m[8][8];
#pragma unroll
    for (int y = 0; y < 8; y++)
    {
        grid_diff4<P>(ra[y], rb[y], c, &m[y][0]);
        grid_diff4<P>(ra[y], rb[y], c + 4, &m[y][4]);
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

// BPL = blocks per lane (2 only for 8-px-wide blocks at 8 bit); lgl = log2(lanes per block) otherwise.
template <typename P, int OP, int BPL, bool BAL>
__global__ void __launch_bounds__(256) k_pixelcmp_grid(const P* __restrict__ A, int64_t sa, const P* __restrict__ B, int64_t sb,
                                                       int bw, int bh, int nbx, int ncols, int64_t nunits, int lgl, uint64_t* __restrict__ out)
{
    constexpr int PX = 16 / (int)sizeof(P);
    const int lane = threadIdx.x & 31;
    const int64_t u0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = u0 < nunits;
    const int64_t u = valid ? u0 : nunits - 1;
    const int by = (int)(u / ncols), col = (int)(u - (int64_t)by * ncols);
    const uint8_t* pa = (const uint8_t*)(A + (int64_t)by * bh * sa + (int64_t)col * PX);
    const uint8_t* pb = (const uint8_t*)(B + (int64_t)by * bh * sb + (int64_t)col * PX);
    const int64_t ba = sa * (int64_t)sizeof(P), bb = sb * (int64_t)sizeof(P);
    unsigned long long acc0 = 0, acc1 = 0;

    if (OP == X265CU_SAD || OP == X265CU_SSE_PP)
    {
        unsigned s0 = 0, s1 = 0;
        for (int y = 0; y < bh; y++, pa += ba, pb += bb)
        {
            const uint4 a = grid_load16<true>(pa), b = grid_load16<BAL>(pb);
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const uint32_t wa = u4w(a, k), wb = u4w(b, k);
                unsigned v;
                if (sizeof(P) == 1)
                {
                    if (OP == X265CU_SAD) v = __vsadu4(wa, wb);
                    else { const unsigned d = __vabsdiffu4(wa, wb); v = __dp4a(d, d, 0u); }
                }
                else
                {
                    const int d0 = (int)(wa & 65535u) - (int)(wb & 65535u), d1 = (int)(wa >> 16) - (int)(wb >> 16);
                    v = OP == X265CU_SAD ? (unsigned)(abs(d0) + abs(d1)) : (unsigned)(d0 * d0 + d1 * d1);
                }
                if (BPL == 2 && k >= 2) s1 += v; else s0 += v;
            }
            if (OP == X265CU_SSE_PP && sizeof(P) == 2 && (y & 7) == 7) { acc0 += s0; s0 = 0; }    // 8 rows x 8 px x 1023^2 < 2^32
        }
        acc0 += s0; acc1 += s1;
    }
    else if (OP == X265CU_SATD)
    {
        unsigned s0 = 0, s1 = 0;
        for (int y = 0; y < bh; y += 4)
        {
            uint4 ra[4], rb[4];
#pragma unroll
            for (int r = 0; r < 4; r++, pa += ba, pb += bb) { ra[r] = grid_load16<true>(pa); rb[r] = grid_load16<BAL>(pb); }
#pragma unroll
            for (int t = 0; t < PX / 8; t++)
            {
                const unsigned v = (unsigned)(grid_had4x4<P>(ra, rb, 8 * t) + grid_had4x4<P>(ra, rb, 8 * t + 4)) >> 1;
                if (BPL == 2 && t == 1) s1 += v; else s0 += v;
            }
        }
        acc0 = s0; acc1 = s1;
    }
    else    // SA8D
    {
        const bool tile16 = ((bw | bh) & 15) == 0;
        unsigned s0 = 0, s1 = 0, pend = 0;
        for (int y = 0; y < bh; y += 8)
        {
            uint4 ra[8], rb[8];
#pragma unroll
            for (int r = 0; r < 8; r++, pa += ba, pb += bb) { ra[r] = grid_load16<true>(pa); rb[r] = grid_load16<BAL>(pb); }
            unsigned v0 = (unsigned)grid_had8x8<P>(ra, rb, 0), v1 = 0;
            if (PX == 16) v1 = (unsigned)grid_had8x8<P>(ra, rb, 8);
            if (tile16)
            {
                // pixel.cpp:361-377: the four 8x8 sums of a 16x16 tile are added, then rounded once
                pend += v0 + v1;
                if (y & 8)
                {
                    if (PX == 8) { pend += __shfl_xor_sync(0xffffffffu, pend, 1); if (lane & 1) pend = 0; }
                    s0 += PX == 8 && (lane & 1) ? 0u : (pend + 2) >> 2;
                    pend = 0;
                }
            }
            else
            {
                s0 += (v0 + 2) >> 2;
                if (PX == 16) { if (BPL == 2) s1 += (v1 + 2) >> 2; else s0 += (v1 + 2) >> 2; }
            }
        }
        acc0 = s0; acc1 = s1;
    }

    for (int o = 1; o < (1 << lgl); o <<= 1) acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
    if (valid && (lane & ((1 << lgl) - 1)) == 0)
    {
        if (OP == X265CU_SSE_PP && sizeof(P) == 1) { acc0 = (uint32_t)acc0; acc1 = (uint32_t)acc1; }     // sse_t is 32 bit at 8 bit depth
        const int64_t idx = (int64_t)by * nbx + (BPL == 2 ? col * 2 : col >> lgl);
        out[idx] = acc0;
        if (BPL == 2) out[idx + 1] = acc1;
    }
}

// generic grid fallback: one warp per block through the job kernel's evaluators
template <typename P>
__global__ void __launch_bounds__(256) k_pixelcmp_grid_generic(int op, const P* __restrict__ A, int64_t sa, const P* __restrict__ B, int64_t sb,
                                                               int bw, int bh, int nbx, int64_t nblocks, uint64_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int64_t j = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); j < nblocks; j += (int64_t)gridDim.x * wpb)
    {
        const int by = (int)(j / nbx), bx = (int)(j - (int64_t)by * nbx);
        const P* a = A + (int64_t)by * bh * sa + (int64_t)bx * bw;
        const P* b = B + (int64_t)by * bh * sb + (int64_t)bx * bw;
        uint64_t res = 0;
        switch (op)
        {
        case X265CU_SAD:  res = (uint32_t)warp_sad(a, (int)sa, b, (int)sb, bw, bh, lane); break;
        case X265CU_SATD: res = (uint32_t)warp_satd(a, (int)sa, b, (int)sb, bw, bh, lane); break;
        case X265CU_SA8D: res = (uint32_t)warp_sa8d(a, (int)sa, b, (int)sb, bw, bh, lane); break;
        default:
        {
            unsigned long long v = warp_sse(a, (int)sa, b, (int)sb, bw, bh, lane);
            res = PixTraits<P>::depth == 8 ? (uint64_t)(uint32_t)v : v;
        }
        }
        if (lane == 0) out[j] = res;
    }
}

template <typename P, int OP>
static void launch_grid_fast(x265cu_ctx* ctx, const P* A, int64_t sa, const P* B, int64_t sb, int bw, int bh, int nbx, int nby, bool bal, uint64_t* out)
{
    constexpr int PX = 16 / (int)sizeof(P);
    const int ncols = nbx * bw / PX;
    const int64_t nunits = (int64_t)ncols * nby;
    const int blocks = (int)((nunits + 255) / 256);
    if (bw < PX)
    {
        if (bal) k_pixelcmp_grid<P, OP, 2, true><<<blocks, 256, 0, ctx->stream>>>(A, sa, B, sb, bw, bh, nbx, ncols, nunits, 0, out);
        else     k_pixelcmp_grid<P, OP, 2, false><<<blocks, 256, 0, ctx->stream>>>(A, sa, B, sb, bw, bh, nbx, ncols, nunits, 0, out);
    }
    else
    {
        int lgl = 0; while ((PX << lgl) < bw) lgl++;
        if (bal) k_pixelcmp_grid<P, OP, 1, true><<<blocks, 256, 0, ctx->stream>>>(A, sa, B, sb, bw, bh, nbx, ncols, nunits, lgl, out);
        else     k_pixelcmp_grid<P, OP, 1, false><<<blocks, 256, 0, ctx->stream>>>(A, sa, B, sb, bw, bh, nbx, ncols, nunits, lgl, out);
    }
}

template <typename P>
static int launch_pixelcmp_grid_t(x265cu_ctx* ctx, int op, const P* A, int64_t sa, const P* B, int64_t sb, int bw, int bh, int nbx, int nby, uint64_t* out)
{
    constexpr int PX = 16 / (int)sizeof(P);
    const bool size_ok = (bw == 8 || bw == 16 || bw == 32 || bw == 64) && bh > 0 &&
                         (op == X265CU_SATD ? (bh & 3) == 0 : op == X265CU_SA8D ? (bh & 7) == 0 : true);
    const bool a_ok = ((uintptr_t)A & 15) == 0 && ((sa * (int64_t)sizeof(P)) & 15) == 0 && ((int64_t)nbx * bw) % PX == 0;
    const bool b_ok = ((sb * (int64_t)sizeof(P)) & 3) == 0;
    const bool fast_op = op == X265CU_SAD || op == X265CU_SATD || op == X265CU_SA8D || op == X265CU_SSE_PP;
    if (fast_op && size_ok && a_ok && b_ok)
    {
        const bool bal = ((uintptr_t)B & 15) == 0 && ((sb * (int64_t)sizeof(P)) & 15) == 0;
        switch (op)
        {
        case X265CU_SAD:  launch_grid_fast<P, X265CU_SAD>(ctx, A, sa, B, sb, bw, bh, nbx, nby, bal, out); break;
        case X265CU_SATD: launch_grid_fast<P, X265CU_SATD>(ctx, A, sa, B, sb, bw, bh, nbx, nby, bal, out); break;
        case X265CU_SA8D: launch_grid_fast<P, X265CU_SA8D>(ctx, A, sa, B, sb, bw, bh, nbx, nby, bal, out); break;
        default:          launch_grid_fast<P, X265CU_SSE_PP>(ctx, A, sa, B, sb, bw, bh, nbx, nby, bal, out); break;
        }
    }
    else
    {
        if (!fast_op || sa > 0x7fffffff || sb > 0x7fffffff) { x265cu_set_error("pixelcmp_grid: unsupported op / stride", cudaErrorInvalidValue, __FILE__, __LINE__); return -1; }
        const int64_t nblocks = (int64_t)nbx * nby;
        int64_t blocks = (nblocks + 7) / 8;
        if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
        k_pixelcmp_grid_generic<P><<<(int)blocks, 256, 0, ctx->stream>>>(op, A, sa, B, sb, bw, bh, nbx, nblocks, out);
    }
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

static int launch_pixelcmp_grid(x265cu_ctx* ctx, int depth, int op, const void* A, int64_t sa, const void* B, int64_t sb,
                                int bw, int bh, int nbx, int nby, uint64_t* out)
{
    if (nbx <= 0 || nby <= 0) return 0;
    if (bw <= 0 || bh <= 0) { x265cu_set_error("pixelcmp_grid: bad block size", cudaErrorInvalidValue, __FILE__, __LINE__); return -1; }
    return depth == 8 ? launch_pixelcmp_grid_t<uint8_t>(ctx, op, (const uint8_t*)A, sa, (const uint8_t*)B, sb, bw, bh, nbx, nby, out)
                      : launch_pixelcmp_grid_t<uint16_t>(ctx, op, (const uint16_t*)A, sa, (const uint16_t*)B, sb, bw, bh, nbx, nby, out);
}

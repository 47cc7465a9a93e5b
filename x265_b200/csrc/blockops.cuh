// x265_b200/csrc/blockops.cuh -- elementwise block-op class.
// Semantics: /root/reference/source/common/pixel.cpp:393-491 (blockfill, cpy1D/2D, residual,
// transpose), :493-557 (weight, pixelavg), :584-602 (scale2D), :759-862 (copies, sub_ps, add_ps,
// addAvg), ipfilter.cpp:40-57 (p2s), dct.cpp:612-634 (dequant_normal).
#pragma once
#include "common.cuh"

template <typename P>
__device__ __forceinline__ void blockop_elem(int op, const x265cu_blk_job& jb, void* __restrict__ D, const void* __restrict__ A,
                                             const void* __restrict__ Bv, int x, int y)
{
    constexpr int depth = PixTraits<P>::depth;
    constexpr int maxv = PixTraits<P>::maxv;
    const int64_t di = jb.d_off + (int64_t)y * jb.d_stride + x;
    const int64_t ai = jb.a_off + (int64_t)y * jb.a_stride + x;
    const int64_t bi = jb.b_off + (int64_t)y * jb.b_stride + x;
    switch (op)
    {
    case X265CU_COPY_PP: ((P*)D)[di] = ((const P*)A)[ai]; break;
    case X265CU_COPY_SS: ((int16_t*)D)[di] = ((const int16_t*)A)[ai]; break;
    case X265CU_COPY_SP: ((P*)D)[di] = (P)((const int16_t*)A)[ai]; break;
    case X265CU_COPY_PS: ((int16_t*)D)[di] = (int16_t)((const P*)A)[ai]; break;
    case X265CU_SUB_PS:  ((int16_t*)D)[di] = (int16_t)((int)((const P*)A)[ai] - (int)((const P*)Bv)[bi]); break;
    case X265CU_ADD_PS:  ((P*)D)[di] = (P)clip3i(0, maxv, (int)((const P*)A)[ai] + (int)((const int16_t*)Bv)[bi]); break;
    case X265CU_PIXELAVG_PP: ((P*)D)[di] = (P)(((int)((const P*)A)[ai] + (int)((const P*)Bv)[bi] + 1) >> 1); break;
    case X265CU_ADDAVG:
    {
        constexpr int shift = 14 + 1 - depth;
        constexpr int offset = (1 << (shift - 1)) + 2 * 8192;
        ((P*)D)[di] = (P)clip3i(0, maxv, ((int)((const int16_t*)A)[ai] + (int)((const int16_t*)Bv)[bi] + offset) >> shift);
        break;
    }
    case X265CU_P2S: ((int16_t*)D)[di] = (int16_t)((int)(int16_t)(((const P*)A)[ai] << (14 - depth)) - 8192); break;
    case X265CU_TRANSPOSE: ((P*)D)[jb.d_off + (int64_t)y * jb.w + x] = ((const P*)A)[jb.a_off + (int64_t)x * jb.a_stride + y]; break;
    case X265CU_BLOCKFILL_S: ((int16_t*)D)[di] = (int16_t)jb.p0; break;
    case X265CU_CPY2DTO1D_SHL: case X265CU_CPY1DTO2D_SHL:
        ((int16_t*)D)[di] = (int16_t)(((const int16_t*)A)[ai] << jb.p0); break;
    case X265CU_CPY2DTO1D_SHR: case X265CU_CPY1DTO2D_SHR:
        ((int16_t*)D)[di] = (int16_t)((((const int16_t*)A)[ai] + (int16_t)(1 << (jb.p0 - 1))) >> jb.p0); break;
    case X265CU_WEIGHT_PP:
    {
        int16_t v = (int16_t)(((const P*)A)[ai] << (14 - depth));
        ((P*)D)[di] = (P)clip3i(0, maxv, ((jb.p0 * v + jb.p1) >> jb.p2) + jb.p3);
        break;
    }
    case X265CU_WEIGHT_SP:
        ((P*)D)[di] = (P)clip3i(0, maxv, ((jb.p0 * (((const int16_t*)A)[ai] + 8192) + jb.p1) >> jb.p2) + jb.p3); break;
    case X265CU_SCALE2D_64TO32:
    {
        const P* p = (const P*)A + jb.a_off + (int64_t)(2 * y) * jb.a_stride + 2 * x;
        ((P*)D)[jb.d_off + y * 32 + x] = (P)((p[0] + p[1] + p[jb.a_stride] + p[jb.a_stride + 1] + 2) >> 2);
        break;
    }
    case X265CU_DEQUANT_NORMAL:
        ((int16_t*)D)[di] = (int16_t)clip16((((const int16_t*)A)[ai] * jb.p0 + (1 << (jb.p1 - 1))) >> jb.p1); break;
    }
}

// one block per job (grid-stride over jobs), threads stride over the w*h elements
template <typename P>
__global__ void __launch_bounds__(256) k_blockop(int op, void* __restrict__ D, const void* __restrict__ A, const void* __restrict__ B,
                                                 const x265cu_blk_job* __restrict__ jobs, int n)
{
    for (int j = blockIdx.x; j < n; j += gridDim.x)
    {
        const x265cu_blk_job jb = jobs[j];
        const int w = jb.w, cnt = jb.w * jb.h;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x)
        {
            int y = i / w, x = i - y * w;
            blockop_elem<P>(op, jb, D, A, B, x, y);
        }
    }
}

static int launch_blockop(x265cu_ctx* ctx, int depth, int op, void* D, const void* A, const void* B, const x265cu_blk_job* jobs, int n)
{
    if (n <= 0) return 0;
    int blocks = n < ctx->sm_count * 16 ? n : ctx->sm_count * 16;
    if (depth == 8) k_blockop<uint8_t><<<blocks, 256, 0, ctx->stream>>>(op, D, A, B, jobs, n);
    else            k_blockop<uint16_t><<<blocks, 256, 0, ctx->stream>>>(op, D, A, B, jobs, n);
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

// ---- frame_init_lowres (pixel.cpp:604-628) + border extension (pixel.cpp:1027-1041) ---------
// Each thread produces 4 adjacent lowres pixels of all 4 planes from a 3x9 full-res window.
template <typename P>
__global__ void __launch_bounds__(256) k_lowres_init(const P* __restrict__ src, int sstride, P* __restrict__ d0, P* __restrict__ dh,
                                                     P* __restrict__ dv, P* __restrict__ dc, int dstride, int width, int height)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= width || y >= height) return;
    const P* r0 = src + (int64_t)(2 * y) * sstride + 2 * x0;
    const P* r1 = r0 + sstride;
    const P* r2 = r1 + sstride;
    int a[9], b[9], c[9];
    const int nload = min(9, 2 * (width - x0) + 1);
#pragma unroll
    for (int i = 0; i < 9; i++)
    {
        int k = i < nload ? i : nload - 1;
        a[i] = r0[k]; b[i] = r1[k]; c[i] = r2[k];
    }
    // vertical pair averages: v01[i] = avg(r0,r1), v12[i] = avg(r1,r2)
    int v01[9], v12[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { v01[i] = (a[i] + b[i] + 1) >> 1; v12[i] = (b[i] + c[i] + 1) >> 1; }
    const int nout = min(4, width - x0);
    const int64_t o = (int64_t)y * dstride + x0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        if (k < nout)
        {
            d0[o + k] = (P)((v01[2 * k] + v01[2 * k + 1] + 1) >> 1);
            dh[o + k] = (P)((v01[2 * k + 1] + v01[2 * k + 2] + 1) >> 1);
            dv[o + k] = (P)((v12[2 * k] + v12[2 * k + 1] + 1) >> 1);
            dc[o + k] = (P)((v12[2 * k + 1] + v12[2 * k + 2] + 1) >> 1);
        }
    }
}

// replicate left/right edges (rows 0..height-1), then copy first/last rows into the top/bottom margins
template <typename P>
__global__ void k_extend_lr(P* __restrict__ pic, int stride, int width, int height, int marginX)
{
    int y = blockIdx.x;
    if (y >= height) return;
    P* row = pic + (int64_t)y * stride;
    P l = row[0], r = row[width - 1];
    for (int x = threadIdx.x; x < marginX; x += blockDim.x) { row[-marginX + x] = l; row[width + x] = r; }
}
template <typename P>
__global__ void k_extend_tb(P* __restrict__ pic, int stride, int width, int height, int marginX, int marginY)
{
    int m = blockIdx.x;                  // 0..2*marginY-1
    const P* srcrow = (m < marginY) ? pic - marginX : pic - marginX + (int64_t)(height - 1) * stride;
    P* dst = (m < marginY) ? pic - marginX - (int64_t)(m + 1) * stride : pic - marginX + (int64_t)(height + (m - marginY)) * stride;
    for (int x = threadIdx.x; x < stride; x += blockDim.x) dst[x] = srcrow[x];
}

template <typename P>
static int extend_border_t(x265cu_ctx* ctx, P* pic, int stride, int width, int height, int mx, int my)
{
    k_extend_lr<P><<<height, 64, 0, ctx->stream>>>(pic, stride, width, height, mx);
    CU_LAUNCH_CHECK(ctx);
    if (my > 0)
    {
        k_extend_tb<P><<<2 * my, 256, 0, ctx->stream>>>(pic, stride, width, height, mx, my);
        CU_LAUNCH_CHECK(ctx);
    }
    return 0;
}

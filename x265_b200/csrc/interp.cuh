// x265_b200/csrc/interp.cuh -- sub-pel interpolation class (8-tap luma / 4-tap chroma).
// Semantics: /root/reference/source/common/ipfilter.cpp:79-118 (hpp), :120-162 (hps, rowExt),
// :164-203 (vpp), :205-239 (vps), :241-282 (vsp), :284-317 (vss), :362-369 (hvpp = hps+vsp).
// IF_INTERNAL_PREC 14, IF_FILTER_PREC 6, IF_INTERNAL_OFFS 8192 (common/constants.h:66-70).
#pragma once
#include "common.cuh"

__device__ __forceinline__ const int16_t* interp_taps(int ntaps, int idx)
{
    return ntaps == 8 ? c_lumaFilter[idx] : c_chromaFilter[idx];
}

// rounding / output stage shared by all variants.  `kind`: 0 = pp (pixel in, pixel out),
// 1 = ps (pixel in, short out), 2 = sp (short in, pixel out), 3 = ss (short in, short out)
template <int DEPTH>
__device__ __forceinline__ int interp_finish(int sum, int kind)
{
    constexpr int headroom = 14 - DEPTH;
    constexpr int maxv = (1 << DEPTH) - 1;
    switch (kind)
    {
    case 0:  return clip3i(0, maxv, (int)(int16_t)((sum + 32) >> 6));
    case 1:  { constexpr int shift = 6 - headroom; return (int)(int16_t)((sum + (int)((unsigned)-8192 << shift)) >> shift); }
    case 2:  { constexpr int shift = 6 + headroom; return clip3i(0, maxv, (int)(int16_t)((sum + (1 << (shift - 1)) + (8192 << 6)) >> shift)); }
    default: return (int)(int16_t)(sum >> 6);
    }
}

// One CTA per job.  The source window (block + filter halo) is staged once in shared memory as int16,
// the horizontal pass writes a second shared tile, the vertical pass reads it: every source sample is
// fetched from HBM/L2 exactly once per job.
template <typename P>
__global__ void __launch_bounds__(256) k_interp(int op, const void* __restrict__ srcv, void* __restrict__ dstv,
                                                const x265cu_interp_job* __restrict__ jobs, int n)
{
    constexpr int DEPTH = PixTraits<P>::depth;
    __shared__ int16_t win[72 * 72];
    __shared__ int16_t mid[72 * 64];
    for (int j = blockIdx.x; j < n; j += gridDim.x)
    {
        const x265cu_interp_job jb = jobs[j];
        const int w = jb.w, h = jb.h, nt = jb.ntaps;
        const bool srcShort = (op == X265CU_VSP || op == X265CU_VSS);
        const bool doH = (op == X265CU_HPP || op == X265CU_HPS || op == X265CU_HVPP);
        const bool doV = (op == X265CU_VPP || op == X265CU_VPS || op == X265CU_VSP || op == X265CU_VSS || op == X265CU_HVPP);
        const bool vHalo = doV || (op == X265CU_HPS && jb.rowExt);
        const int hl = doH ? nt / 2 - 1 : 0, hr = doH ? nt / 2 : 0;         // left / right halo
        const int vt = vHalo ? nt / 2 - 1 : 0, vb = vHalo ? nt / 2 : 0;      // top / bottom halo
        const int ww = w + hl + hr, wh = h + vt + vb;
        // stage window
        for (int i = threadIdx.x; i < ww * wh; i += blockDim.x)
        {
            int r = i / ww, c = i - r * ww;
            int64_t si = jb.s_off + (int64_t)(r - vt) * jb.s_stride + (c - hl);
            win[r * 72 + c] = srcShort ? ((const int16_t*)srcv)[si] : (int16_t)((const P*)srcv)[si];
        }
        __syncthreads();
        const int16_t* cx = interp_taps(nt, jb.idxX);
        const int16_t* cy = interp_taps(nt, jb.idxY);
        if (op == X265CU_HPP || op == X265CU_HPS)
        {
            const int16_t* c = cx;
            const int kind = (op == X265CU_HPP) ? 0 : 1;
            for (int i = threadIdx.x; i < w * wh; i += blockDim.x)
            {
                int r = i / w, x = i - r * w;
                int sum = 0;
                for (int k = 0; k < nt; k++) sum += (int)win[r * 72 + x + k] * c[k];
                int v = interp_finish<DEPTH>(sum, kind);
                int64_t di = jb.d_off + (int64_t)r * jb.d_stride + x;      // rowExt: dst row 0 = 3 rows above the block
                if (kind == 0) ((P*)dstv)[di] = (P)v; else ((int16_t*)dstv)[di] = (int16_t)v;
            }
        }
        else if (op == X265CU_HVPP)
        {
            for (int i = threadIdx.x; i < w * wh; i += blockDim.x)
            {
                int r = i / w, x = i - r * w;
                int sum = 0;
                for (int k = 0; k < nt; k++) sum += (int)win[r * 72 + x + k] * cx[k];
                mid[r * 64 + x] = (int16_t)interp_finish<DEPTH>(sum, 1);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < w * h; i += blockDim.x)
            {
                int y = i / w, x = i - y * w;
                int sum = 0;
                for (int k = 0; k < nt; k++) sum += (int)mid[(y + k) * 64 + x] * cy[k];
                ((P*)dstv)[jb.d_off + (int64_t)y * jb.d_stride + x] = (P)interp_finish<DEPTH>(sum, 2);
            }
        }
        else
        {
            // vertical-only variants; coefficient index is idxX for the single-index signatures
            const int16_t* c = cx;
            const int kind = op == X265CU_VPP ? 0 : (op == X265CU_VPS ? 1 : (op == X265CU_VSP ? 2 : 3));
            for (int i = threadIdx.x; i < w * h; i += blockDim.x)
            {
                int y = i / w, x = i - y * w;
                int sum = 0;
                for (int k = 0; k < nt; k++) sum += (int)win[(y + k) * 72 + x] * c[k];
                int v = interp_finish<DEPTH>(sum, kind);
                int64_t di = jb.d_off + (int64_t)y * jb.d_stride + x;
                if (kind == 0 || kind == 2) ((P*)dstv)[di] = (P)v; else ((int16_t*)dstv)[di] = (int16_t)v;
            }
        }
        __syncthreads();
    }
}

static int launch_interp(x265cu_ctx* ctx, int depth, int op, const void* src, void* dst, const x265cu_interp_job* jobs, int n)
{
    if (n <= 0) return 0;
    int blocks = n < ctx->sm_count * 8 ? n : ctx->sm_count * 8;
    if (depth == 8) k_interp<uint8_t><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, jobs, n);
    else            k_interp<uint16_t><<<blocks, 256, 0, ctx->stream>>>(op, src, dst, jobs, n);
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

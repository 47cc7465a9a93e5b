// x265_b200/csrc/predcost.cuh -- prediction costs of predInterSearch around the motion search, batched:
//   Search::selectMVP        (/root/reference/source/encoder/search.cpp:1992-2023)  SAD of the luma prediction at an AMVP candidate
//   Search::mergeEstimation  (search.cpp:1901-1960)  SATD (+ Cb/Cr SATD) of a merge candidate's motion-compensated prediction
//   bidir of predInterSearch (search.cpp:2474-2607)  SATD of the bi-prediction: addAvg of the two 14-bit intermediates
//                                                    (+ chroma) when bChromaSATD, else pixelavg_pp of two pixel predictions
// The prediction is Predict::motionCompensation's unweighted path (common/predict.cpp:76-240): predInterLumaPixel /
// predInterLumaShort / predInterChromaPixel / predInterChromaShort (predict.cpp:242-407), Yuv::addAvg (yuv.cpp:189-211);
// the costs are MotionEstimate::bufSAD / bufSATD / bufChromaSATD (encoder/motion.h:87-95).  4:2:0, vectors clipped by the
// caller (CUData::clipMv).  One job = one (PU, candidate); nothing is written but the cost.
//
// Four launches over the same list: one-list jobs on the motion search's sub-pel code (k_pred_uni, small / large PUs), two-list
// jobs staged through shared memory (k_pred_cost): PUs up to 16x16 take one WARP each (four per CTA, private tiles, warp
// barriers only), larger PUs one 128-thread CTA each; the prediction lives in shared memory from the source window to the cost.
#pragma once
#include "common.cuh"
#include "interp.cuh"
#include "pixelcmp.cuh"

template <int TPJ> __device__ __forceinline__ void pc_sync() { if (TPJ == 32) __syncwarp(); else __syncthreads(); }

// i -> (row, column) of a block `w` wide: a shift for the power-of-two widths (every symmetric PU), a division for 12 / 24 / 48
struct PcDiv { int w, lg; };
__device__ __forceinline__ PcDiv pc_div(int width) { PcDiv d; d.w = width; d.lg = (width & (width - 1)) ? -1 : 31 - __clz(width); return d; }
__device__ __forceinline__ void pc_split(const PcDiv& d, int i, int& y, int& x) { y = d.lg >= 0 ? i >> d.lg : i / d.w; x = i - y * d.w; }

// One plane's prediction of a w x h block into dst (int16, pitch DP): pixel domain (SHORT = false: copy / hpp / vpp / hvpp)
// or the 14-bit intermediate (SHORT = true: p2s / hps / vps / hps + vss).  NT = 8 (luma, quarter-pel) or 4 (chroma, eighth).
template <typename P, int TPJ, int MAXS, int NT, bool SHORT>
__device__ __forceinline__ void pc_predict(const P* __restrict__ ref, int stride, int mvx, int mvy, int w, int h,
                                           int16_t* __restrict__ dst, int16_t* __restrict__ s_win, int16_t* __restrict__ s_mid, int t)
{
    constexpr int DEPTH = PixTraits<P>::depth;
    constexpr int FB = NT == 8 ? 2 : 3, HL = NT / 2 - 1;          // fraction bits; taps left of / above the sample
    constexpr int WP = MAXS + 8, DP = MAXS;                       // window / destination pitch
    const int xf = mvx & ((1 << FB) - 1), yf = mvy & ((1 << FB) - 1);
    const P* src = ref + (mvx >> FB) + (ptrdiff_t)(mvy >> FB) * stride;
    if (!(xf | yf))
    {
        const PcDiv dw = pc_div(w);
        for (int i = t; i < w * h; i += TPJ)
        {
            int y, x; pc_split(dw, i, y, x);
            const int p = src[(ptrdiff_t)y * stride + x];
            dst[y * DP + x] = (int16_t)(SHORT ? (p << (14 - DEPTH)) - 8192 : p);
        }
        return;
    }
    const int hl = xf ? HL : 0, vt = yf ? HL : 0;
    const int ww = w + (xf ? NT - 1 : 0), wh = h + (yf ? NT - 1 : 0);
    const PcDiv dw = pc_div(w);
    for (int i = t; i < ww * wh; i += TPJ)
    {
        const int r = i / ww, c = i - r * ww;
        s_win[r * WP + c] = (int16_t)src[(ptrdiff_t)(r - vt) * stride + (c - hl)];
    }
    pc_sync<TPJ>();
    const int16_t* cx = NT == 8 ? c_lumaFilter[xf] : c_chromaFilter[xf];
    const int16_t* cy = NT == 8 ? c_lumaFilter[yf] : c_chromaFilter[yf];
    if (!yf || !xf)
    {
        const int step = !yf ? 1 : WP;
        const int16_t* cf = !yf ? cx : cy;
        for (int i = t; i < w * h; i += TPJ)
        {
            int y, x; pc_split(dw, i, y, x);
            int sum = 0;
#pragma unroll
            for (int k = 0; k < NT; k++) sum += (int)s_win[y * WP + x + k * step] * cf[k];
            dst[y * DP + x] = (int16_t)interp_finish<DEPTH>(sum, SHORT ? 1 : 0);
        }
    }
    else
    {
        for (int i = t; i < w * wh; i += TPJ)
        {
            int r, x; pc_split(dw, i, r, x);
            int sum = 0;
#pragma unroll
            for (int k = 0; k < NT; k++) sum += (int)s_win[r * WP + x + k] * cx[k];
            s_mid[r * DP + x] = (int16_t)interp_finish<DEPTH>(sum, 1);
        }
        pc_sync<TPJ>();
        for (int i = t; i < w * h; i += TPJ)
        {
            int y, x; pc_split(dw, i, y, x);
            int sum = 0;
#pragma unroll
            for (int k = 0; k < NT; k++) sum += (int)s_mid[(y + k) * DP + x] * cy[k];
            dst[y * DP + x] = (int16_t)interp_finish<DEPTH>(sum, SHORT ? 3 : 2);
        }
    }
}

// prediction of one plane for the job's lists into s_p0 (pixel values), per motionCompensation / the bidir estimate
template <typename P, int TPJ, int MAXS, int NT>
__device__ __forceinline__ void pc_plane(const P* r0, const P* r1, int stride, const x265cu_pred_job& jb, int w, int h,
                                         int16_t* s_p0, int16_t* s_p1, int16_t* s_win, int16_t* s_mid, int t)
{
    constexpr int DEPTH = PixTraits<P>::depth, maxv = PixTraits<P>::maxv, DP = MAXS;
    const bool bi = r0 && r1;
    pc_sync<TPJ>();                                               // the previous plane's cost has been read
    if (!bi)
    {
        pc_predict<P, TPJ, MAXS, NT, false>(r0 ? r0 : r1, stride, r0 ? jb.mv0[0] : jb.mv1[0], r0 ? jb.mv0[1] : jb.mv1[1], w, h, s_p0, s_win, s_mid, t);
        pc_sync<TPJ>();
        return;
    }
    if (jb.flags & X265CU_PRED_AVG_PP)
    {
        pc_predict<P, TPJ, MAXS, NT, false>(r0, stride, jb.mv0[0], jb.mv0[1], w, h, s_p0, s_win, s_mid, t);
        pc_sync<TPJ>();
        pc_predict<P, TPJ, MAXS, NT, false>(r1, stride, jb.mv1[0], jb.mv1[1], w, h, s_p1, s_win, s_mid, t);
        pc_sync<TPJ>();
        const PcDiv dw = pc_div(w);
        for (int i = t; i < w * h; i += TPJ)
        {
            int y, x; pc_split(dw, i, y, x);
            s_p0[y * DP + x] = (int16_t)(((int)s_p0[y * DP + x] + (int)s_p1[y * DP + x] + 1) >> 1);       // pixelavg_pp (pixel.cpp:545-557)
        }
    }
    else
    {
        pc_predict<P, TPJ, MAXS, NT, true>(r0, stride, jb.mv0[0], jb.mv0[1], w, h, s_p0, s_win, s_mid, t);
        pc_sync<TPJ>();
        pc_predict<P, TPJ, MAXS, NT, true>(r1, stride, jb.mv1[0], jb.mv1[1], w, h, s_p1, s_win, s_mid, t);
        pc_sync<TPJ>();
        constexpr int shift = 15 - DEPTH, offset = (1 << (shift - 1)) + 2 * 8192;                       // addAvg (pixel.cpp:842-862)
        const PcDiv dw = pc_div(w);
        for (int i = t; i < w * h; i += TPJ)
        {
            int y, x; pc_split(dw, i, y, x);
            s_p0[y * DP + x] = (int16_t)clip3i(0, maxv, ((int)s_p0[y * DP + x] + (int)s_p1[y * DP + x] + offset) >> shift);
        }
    }
    pc_sync<TPJ>();
}

// SAD / SATD of (source block in global memory, prediction in shared memory); the SATD tiling is the reference's per size
// (8x4 tiles halved once when the width is a multiple of 8, else 4x4 tiles: pixel.cpp:263-297, 1131-1155)
template <typename P, int TPJ, int MAXS>
__device__ __forceinline__ int pc_cost(const P* __restrict__ f, int fs, const int16_t* __restrict__ pr, int w, int h, bool satd, int t)
{
    constexpr int DP = MAXS;
    int acc = 0;
    if (!satd)
    {
        const PcDiv dw = pc_div(w);
        for (int i = t; i < w * h; i += TPJ) { int y, x; pc_split(dw, i, y, x); acc += abs((int)f[(ptrdiff_t)y * fs + x] - (int)pr[y * DP + x]); }
    }
    else if (!(w & 7))
    {
        const int tw = w >> 3, nt = tw * (h >> 2);
        for (int q = t; q < nt; q += TPJ)
        {
            const int ty = q / tw, tx = q - ty * tw;
            const P* pf = f + (ptrdiff_t)(ty * 4) * fs + tx * 8; const int16_t* pp = pr + (ty * 4) * DP + tx * 8;
            acc += (had4x4_abs(pf, fs, pp, DP) + had4x4_abs(pf + 4, fs, pp + 4, DP)) >> 1;
        }
    }
    else
    {
        const int tw = w >> 2, nt = tw * (h >> 2);
        for (int q = t; q < nt; q += TPJ)
        {
            const int ty = q / tw, tx = q - ty * tw;
            acc += had4x4_abs(f + (ptrdiff_t)(ty * 4) * fs + tx * 4, fs, pr + (ty * 4) * DP + tx * 4, DP) >> 1;
        }
    }
    return acc;
}

struct PredChroma { const void* fcb; const void* fcr; const void* const* rcb; const void* const* rcr; int cstride; };

// TPJ = 32: a warp per job (PUs up to MAXS = 16); TPJ = 128: a CTA per job (MAXS = 64)
template <typename P, int TPJ, int MAXS>
__global__ void __launch_bounds__(128, TPJ == 32 ? 8 : 4) k_pred_cost(const P* __restrict__ fenc, int fstride, const P* const* __restrict__ refs, int rstride,
                                                   PredChroma ch, const x265cu_pred_job* __restrict__ jobs, int n, int32_t* __restrict__ out, int uniElsewhere)
{
    constexpr int JPB = 128 / TPJ;
    constexpr int WIN = (MAXS + 7) * (MAXS + 8), MID = (MAXS + 7) * MAXS, PRD = MAXS * MAXS;
    __shared__ int16_t s_all[JPB][WIN + MID + 2 * PRD];
    __shared__ int s_red[4];
    const int grp = threadIdx.x / TPJ, t = threadIdx.x % TPJ;
    int16_t* s_win = s_all[grp]; int16_t* s_mid = s_win + WIN; int16_t* s_p0 = s_mid + MID; int16_t* s_p1 = s_p0 + PRD;
    for (int j = blockIdx.x * JPB + grp; j < n; j += gridDim.x * JPB)
    {
        const x265cu_pred_job jb = jobs[j];
        const int w = jb.pw, h = jb.ph;
        const bool small = w <= 16 && h <= 16;
        if (small != (MAXS == 16)) continue;                     // the sibling launch owns this job
        if (uniElsewhere && (jb.ref0 < 0 || jb.ref1 < 0)) continue;   // one-list jobs: k_pred_uni (the motion search's sub-pel code)
        const P* r0 = jb.ref0 >= 0 ? refs[jb.ref0] + jb.offset : nullptr;
        const P* r1 = jb.ref1 >= 0 ? refs[jb.ref1] + jb.offset : nullptr;
        pc_plane<P, TPJ, MAXS, 8>(r0, r1, rstride, jb, w, h, s_p0, s_p1, s_win, s_mid, t);
        int acc = pc_cost<P, TPJ, MAXS>(fenc + jb.offset, fstride, s_p0, w, h, jb.cost == X265CU_PRED_SATD, t);
        const int cw = w >> 1, chh = h >> 1;
        const bool bi = r0 && r1;
        if (jb.cost == X265CU_PRED_SATD && (jb.flags & X265CU_PRED_CHROMA) && !(bi && (jb.flags & X265CU_PRED_AVG_PP)) && ch.fcb && !((cw | chh) & 3))
        {
            const int py = jb.offset / fstride, px = jb.offset - py * fstride;
            const ptrdiff_t coff = (ptrdiff_t)(py >> 1) * ch.cstride + (px >> 1);
            for (int p = 0; p < 2; p++)
            {
                const P* const* tab = (const P* const*)(p ? ch.rcr : ch.rcb);
                const P* c0 = jb.ref0 >= 0 ? tab[jb.ref0] + coff : nullptr;
                const P* c1 = jb.ref1 >= 0 ? tab[jb.ref1] + coff : nullptr;
                pc_plane<P, TPJ, MAXS, 4>(c0, c1, ch.cstride, jb, cw, chh, s_p0, s_p1, s_win, s_mid, t);
                acc += pc_cost<P, TPJ, MAXS>((const P*)(p ? ch.fcr : ch.fcb) + coff, ch.cstride, s_p0, cw, chh, true, t);
            }
        }
        acc = warp_sum(acc);
        if (TPJ == 32) { if (t == 0) out[j] = acc; }
        else
        {
            __syncthreads();
            if ((t & 31) == 0) s_red[t >> 5] = acc;
            __syncthreads();
            if (t == 0) out[j] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        }
    }
}

// One-list jobs (AMVP candidates, uni-directional merge candidates) are exactly what MotionEstimate::subpelCompare computes for
// one candidate (motion.cpp:1571-1664: the same copy / hpp / vpp / hvpp prediction, SAD or SATD, + the chroma-SATD term), so
// they run on the motion search's own sub-pel code (me.cuh: me_subpel_batch, me_chroma_batch -- lane-per-row-segment DP4A
// interpolation straight from the planes) instead of the staged generic path above: a warp per job, small-PU and large-PU
// instantiations as in the search.  Same arithmetic, ~5x fewer instructions per job.
template <typename P, int CLS>
__global__ void __launch_bounds__(256, CLS == 1 ? ME_BIG_BLOCKS : ME_MIN_BLOCKS) k_pred_uni(const P* __restrict__ fenc, int fstride, const P* const* __restrict__ refs, int rstride,
                                                                                        MeChromaArgs ch, int haveChroma, const x265cu_pred_job* __restrict__ jobs, int n,
                                                                                        int32_t* __restrict__ out, int* __restrict__ counter)
{
    extern __shared__ unsigned char me_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MeShared* sm = (MeShared*)me_smem + warp;
    for (;;)
    {
        int jid = 0;
        if (lane == 0) jid = atomicAdd(counter, 1);
        jid = __shfl_sync(0xffffffffu, jid, 0);
        if (jid >= n) break;
        const x265cu_pred_job pj = jobs[jid];
        if (pj.ref0 >= 0 && pj.ref1 >= 0) continue;                // two lists: k_pred_cost
        const bool l0 = pj.ref0 >= 0;
        x265cu_me_job j;
        j.offset = pj.offset; j.ref = (int16_t)(l0 ? pj.ref0 : pj.ref1); j.pw = (int8_t)pj.pw; j.ph = (int8_t)pj.ph;
        j.mvmin[0] = j.mvmin[1] = j.mvmax[0] = j.mvmax[1] = 0; j.qmvp[0] = j.qmvp[1] = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) j.mvc[k] = 0;
        j.numCand = 0; j.method = 3; j.merange = 0;
        j.subme = (int8_t)((pj.flags & X265CU_PRED_CHROMA) ? 3 : 2);
        MeCtx<P> c;
        me_make_ctx<P>(c, j, fenc, fstride, refs, rstride, 0, nullptr, lane, sm);
        if (me_subpel_class(c) != CLS) continue;                   // the sibling launch owns this job
        const int qx = l0 ? pj.mv0[0] : pj.mv1[0], qy = l0 ? pj.mv0[1] : pj.mv1[1];
        const bool satd = pj.cost == X265CU_PRED_SATD;
        int cost = me_subpel_batch<P, CLS>(c, 1, qx, qy, satd);
        if (satd && haveChroma && (pj.flags & X265CU_PRED_CHROMA))
        {
            MeChromaCtx<P> cc;
            me_set_chroma<P>(cc, c, j, ch, fstride);
            if (cc.on) cost += me_chroma_batch<P, CLS>(c, cc, 1, qx, qy, 1u);
        }
        if (lane == 0) out[jid] = cost;
        __syncwarp();
    }
}

template <typename P>
static int launch_pred_cost_t(x265cu_ctx* ctx, const void* fenc, int fstride, const void* const* refs, int rstride, const PredChroma& ch,
                              const x265cu_pred_job* jobs, int n, int32_t* out)
{
    // one-list jobs
    {
        CU_CHECK(cudaMemsetAsync(ctx->d_counter + 12, 0, 2 * sizeof(int), ctx->stream));
        MeChromaArgs a; a.fencCb = ch.fcb; a.fencCr = ch.fcr; a.refCb = ch.rcb; a.refCr = ch.rcr; a.cstride = ch.cstride;
        const int threads = 256, warps = threads / 32;
        const size_t smem = sizeof(MeShared) * warps;
        const int need = (n + warps - 1) / warps;
        int b0 = ctx->sm_count * ME_MIN_BLOCKS, b1 = ctx->sm_count * ME_BIG_BLOCKS;
        if (b0 > need) b0 = need;
        if (b1 > need) b1 = need;
        k_pred_uni<P, 0><<<b0, threads, smem, ctx->stream>>>((const P*)fenc, fstride, (const P* const*)refs, rstride, a, ch.fcb != NULL, jobs, n, out, ctx->d_counter + 12);
        CU_LAUNCH_CHECK(ctx);
        k_pred_uni<P, 1><<<b1, threads, smem, ctx->stream>>>((const P*)fenc, fstride, (const P* const*)refs, rstride, a, ch.fcb != NULL, jobs, n, out, ctx->d_counter + 13);
        CU_LAUNCH_CHECK(ctx);
    }
    // two-list jobs
    const int maxb = ctx->sm_count * 16;
    int b1 = (n + 3) / 4; if (b1 > maxb) b1 = maxb;
    int b2 = n < maxb ? n : maxb;
    k_pred_cost<P, 32, 16><<<b1, 128, 0, ctx->stream>>>((const P*)fenc, fstride, (const P* const*)refs, rstride, ch, jobs, n, out, 1);
    CU_LAUNCH_CHECK(ctx);
    k_pred_cost<P, 128, 64><<<b2, 128, 0, ctx->stream>>>((const P*)fenc, fstride, (const P* const*)refs, rstride, ch, jobs, n, out, 1);
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

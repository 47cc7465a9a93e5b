// x265_b200/csrc/me.cuh -- motion-estimation class: one warp replays one
// MotionEstimate::motionEstimate() call (/root/reference/source/encoder/motion.cpp:739-1569)
// bit-exactly: MVP / zero / candidate pre-checks (:771-814), DIA (:822-846), HEX (:848-945),
// STAR (:362-604, :1132-1240 incl. the raster refinement and its `tmv << 3` quirk), sub-pel
// refinement by workload[subme] (:48-58, :1449-1558) through subpelCompare (:1571-1598, luma) or
// the lowres qpel path (common/lowres.h:94-120).  Costs are warp-uniform after each reduction, so
// all 32 lanes take identical branches.  Parallelism = (PUs x refs x CTUs) jobs per launch.
#pragma once
#include "common.cuh"
#include "pixelcmp.cuh"
#include "interp.cuh"

#define ME_BAND 16                         // rows of prediction staged per pass
#define ME_MID_ROWS (ME_BAND + 7)

struct MeShared                            // per-warp scratch
{
    int16_t mid[ME_MID_ROWS * 64];         // hps(rowExt) intermediate of one band
    uint16_t pred[ME_BAND * 64];           // predicted band (pixel values)
};

template <typename P>
struct MeCtx
{
    const P* fenc;  int fstride;           // fenc block origin (plane + offset)
    const P* ref[4]; int rstride;          // ref plane(s) + offset
    const uint16_t* mvc;                   // centred mvcost table
    int mvpx, mvpy;
    int minx, miny, maxx, maxy;            // full-pel bounds
    int w, h, lane, lowres;
    MeShared* sm;
};

template <typename P>
__device__ __forceinline__ int me_mvcost(const MeCtx<P>& c, int qx, int qy)
{
    return (uint16_t)(c.mvc[qx - c.mvpx] + c.mvc[qy - c.mvpy]);
}

// SAD of the fenc block against ref at an arbitrary element pointer (rows may be unaligned).
// Lanes walk 4-byte words; unaligned ref words are assembled from two aligned loads (funnel shift).
template <typename P>
__device__ __forceinline__ int me_sad_direct(const MeCtx<P>& c, const P* __restrict__ r)
{
    const int rowBytes = c.w * (int)sizeof(P);
    const int wpr = rowBytes >> 2;                     // words per row
    const int rpp = 32 / wpr;                          // rows per pass (wpr <= 32)
    const int myrow = c.lane / wpr, myword = c.lane - myrow * wpr;
    const bool active = myrow < rpp;
    const uintptr_t rbase = (uintptr_t)r;
    int acc = 0;
    for (int y0 = 0; y0 < c.h; y0 += rpp)
    {
        int y = y0 + myrow;
        if (active && y < c.h)
        {
            uint32_t f = *(const uint32_t*)((const uint8_t*)c.fenc + (size_t)y * c.fstride * sizeof(P) + myword * 4);
            uintptr_t a = rbase + (size_t)y * c.rstride * sizeof(P) + myword * 4;
            const uint32_t* ap = (const uint32_t*)(a & ~(uintptr_t)3);
            uint32_t sh = (uint32_t)(a & 3) * 8;
            uint32_t lo = ap[0];
            uint32_t v = lo;
            if (sh) v = __funnelshift_r(lo, ap[1], sh);
            if (sizeof(P) == 1) acc = __vsadu4(f, v) + acc;
            else                acc = __vsadu2(f, v) + acc;
        }
    }
    return warp_sum(acc);
}

// cost of a band of prediction held in c.sm->pred (stride 64) against fenc rows [y0, y0+rows)
template <typename P>
__device__ __forceinline__ int me_band_cost(const MeCtx<P>& c, int y0, int rows, bool satd)
{
    const uint16_t* pr = c.sm->pred;
    const P* f = c.fenc + (size_t)y0 * c.fstride;
    int acc = 0;
    if (!satd)
    {
        const int n = c.w * rows;
        for (int i = c.lane; i < n; i += 32)
        {
            int y = i / c.w, x = i - y * c.w;
            acc += abs((int)f[y * c.fstride + x] - (int)pr[y * 64 + x]);
        }
    }
    else if ((c.w & 7) == 0)
    {
        const int tw = c.w >> 3, nt = tw * (rows >> 2);
        for (int t = c.lane; t < nt; t += 32)
        {
            int ty = t / tw, tx = t - ty * tw;
            const P* pf = f + (ty * 4) * c.fstride + tx * 8; const uint16_t* pp = pr + (ty * 4) * 64 + tx * 8;
            acc += (had4x4_abs(pf, c.fstride, pp, 64) + had4x4_abs(pf + 4, c.fstride, pp + 4, 64)) >> 1;
        }
    }
    else
    {
        const int tw = c.w >> 2, nt = tw * (rows >> 2);
        for (int t = c.lane; t < nt; t += 32)
        {
            int ty = t / tw, tx = t - ty * tw;
            acc += had4x4_abs(f + (ty * 4) * c.fstride + tx * 4, c.fstride, pr + (ty * 4) * 64 + tx * 4, 64) >> 1;
        }
    }
    return acc;           // un-reduced partial (caller reduces once)
}

// subpelCompare (motion.cpp:1571-1598): luma_hpp / luma_vpp / luma_hvpp into a scratch, then cmp.
template <typename P>
__device__ int me_subpel_compare(const MeCtx<P>& c, int qx, int qy, bool satd)
{
    constexpr int DEPTH = PixTraits<P>::depth;
    const P* r = c.ref[0] + (qx >> 2) + (ptrdiff_t)(qy >> 2) * c.rstride;
    const int xf = qx & 3, yf = qy & 3;
    if (!(xf | yf))
    {
        if (!satd) return me_sad_direct(c, r);
        return warp_satd(c.fenc, c.fstride, r, c.rstride, c.w, c.h, c.lane);
    }
    const int16_t* cx = c_lumaFilter[xf];
    const int16_t* cy = c_lumaFilter[yf];
    int acc = 0;
    for (int y0 = 0; y0 < c.h; y0 += ME_BAND)
    {
        const int rows = min(ME_BAND, c.h - y0);
        __syncwarp();
        if (!yf)
        {   // horizontal only: pp rounding
            for (int i = c.lane; i < c.w * rows; i += 32)
            {
                int y = i / c.w, x = i - y * c.w;
                const P* s = r + (ptrdiff_t)(y0 + y) * c.rstride + x - 3;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)s[k] * cx[k];
                c.sm->pred[y * 64 + x] = (uint16_t)interp_finish<DEPTH>(sum, 0);
            }
        }
        else if (!xf)
        {   // vertical only
            for (int i = c.lane; i < c.w * rows; i += 32)
            {
                int y = i / c.w, x = i - y * c.w;
                const P* s = r + (ptrdiff_t)(y0 + y - 3) * c.rstride + x;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)s[(ptrdiff_t)k * c.rstride] * cy[k];
                c.sm->pred[y * 64 + x] = (uint16_t)interp_finish<DEPTH>(sum, 0);
            }
        }
        else
        {   // hps with row extension into mid, then vsp
            const int mrows = rows + 7;
            for (int i = c.lane; i < c.w * mrows; i += 32)
            {
                int y = i / c.w, x = i - y * c.w;
                const P* s = r + (ptrdiff_t)(y0 + y - 3) * c.rstride + x - 3;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)s[k] * cx[k];
                c.sm->mid[y * 64 + x] = (int16_t)interp_finish<DEPTH>(sum, 1);
            }
            __syncwarp();
            for (int i = c.lane; i < c.w * rows; i += 32)
            {
                int y = i / c.w, x = i - y * c.w;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)c.sm->mid[(y + k) * 64 + x] * cy[k];
                c.sm->pred[y * 64 + x] = (uint16_t)interp_finish<DEPTH>(sum, 2);
            }
        }
        __syncwarp();
        acc += me_band_cost(c, y0, rows, satd);
    }
    return warp_sum(acc);
}

// lowresQPelCost (lowres.h:94-120): qpel = rounded average of the two nearest hpel planes; 8x8 blocks
template <typename P>
__device__ int me_lowres_cost(const MeCtx<P>& c, int qx, int qy, bool satd)
{
    if ((qx | qy) & 1)
    {
        int ha = (qy & 2) | ((qx & 2) >> 1);
        const P* a = c.ref[ha] + (qx >> 2) + (ptrdiff_t)(qy >> 2) * c.rstride;
        int rx = qx + (qx & 1), ry = qy + (qy & 1);
        int hb = (ry & 2) | ((rx & 2) >> 1);
        const P* b = c.ref[hb] + (rx >> 2) + (ptrdiff_t)(ry >> 2) * c.rstride;
        __syncwarp();
        for (int i = c.lane; i < 64; i += 32)
        {
            int y = i >> 3, x = i & 7;
            c.sm->pred[y * 64 + x] = (uint16_t)(((int)a[(ptrdiff_t)y * c.rstride + x] + (int)b[(ptrdiff_t)y * c.rstride + x] + 1) >> 1);
        }
        __syncwarp();
        // lowres PUs are 8x8 (common.h:217-218); the compare uses the job's w,h like the reference's comp()
        return warp_sum(me_band_cost(c, 0, c.h, satd));
    }
    int hp = (qy & 2) | ((qx & 2) >> 1);
    const P* r = c.ref[hp] + (qx >> 2) + (ptrdiff_t)(qy >> 2) * c.rstride;
    if (!satd) return me_sad_direct(c, r);
    return warp_satd(c.fenc, c.fstride, r, c.rstride, c.w, c.h, c.lane);
}

template <typename P>
__device__ __forceinline__ int me_qpel_cost(const MeCtx<P>& c, int qx, int qy, bool satd)
{
    return c.lowres ? me_lowres_cost(c, qx, qy, satd) : me_subpel_compare(c, qx, qy, satd);
}

template <typename P>
__device__ __forceinline__ int me_cost_fpel(const MeCtx<P>& c, int x, int y)
{
    return me_sad_direct(c, c.ref[0] + x + (ptrdiff_t)y * c.rstride) + me_mvcost(c, x * 4, y * 4);
}

struct MeStar { int bx, by, bcost, point, dist; };

template <typename P>
__device__ __forceinline__ void me_star_try(const MeCtx<P>& c, MeStar& s, int x, int y, int point, int dist)
{
    int cost = me_cost_fpel(c, x, y);
    if (cost < s.bcost) { s.bcost = cost; s.bx = x; s.by = y; s.point = point; s.dist = dist; }
}

// StarPatternSearch (motion.cpp:362-604).  The reference's x4 fast path visits the same points in
// the same order as its bounds-checked path, so one checked walk reproduces both.
template <typename P>
__device__ void me_star_pattern(const MeCtx<P>& c, MeStar& s, int earlyExitIters, int merange)
{
    const int ox = s.bx, oy = s.by;
    int saved = s.bcost, rounds = 0;
    {
        if (oy - 1 >= c.miny) me_star_try(c, s, ox, oy - 1, 2, 1);
        if (ox - 1 >= c.minx) me_star_try(c, s, ox - 1, oy, 4, 1);
        if (ox + 1 <= c.maxx) me_star_try(c, s, ox + 1, oy, 5, 1);
        if (oy + 1 <= c.maxy) me_star_try(c, s, ox, oy + 1, 7, 1);
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int d = 2; d <= 8; d <<= 1)
    {
        const int top = oy - d, bot = oy + d, lft = ox - d, rgt = ox + d, h2 = d >> 1;
        const int top2 = oy - h2, bot2 = oy + h2, lft2 = ox - h2, rgt2 = ox + h2;
        saved = s.bcost;
        if (top >= c.miny) me_star_try(c, s, ox, top, 2, d);
        if (top2 >= c.miny)
        {
            if (lft2 >= c.minx) me_star_try(c, s, lft2, top2, 1, h2);
            if (rgt2 <= c.maxx) me_star_try(c, s, rgt2, top2, 3, h2);
        }
        if (lft >= c.minx) me_star_try(c, s, lft, oy, 4, d);
        if (rgt <= c.maxx) me_star_try(c, s, rgt, oy, 5, d);
        if (bot2 <= c.maxy)
        {
            if (lft2 >= c.minx) me_star_try(c, s, lft2, bot2, 6, h2);
            if (rgt2 <= c.maxx) me_star_try(c, s, rgt2, bot2, 8, h2);
        }
        if (bot <= c.maxy) me_star_try(c, s, ox, bot, 7, d);
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int d = 16; d <= merange; d <<= 1)
    {
        const int top = oy - d, bot = oy + d, lft = ox - d, rgt = ox + d, q = d >> 2;
        saved = s.bcost;
        if (top >= c.miny) me_star_try(c, s, ox, top, 0, d);
        if (lft >= c.minx) me_star_try(c, s, lft, oy, 0, d);
        if (rgt <= c.maxx) me_star_try(c, s, rgt, oy, 0, d);
        if (bot <= c.maxy) me_star_try(c, s, ox, bot, 0, d);
        for (int k = 1; k < 4; k++)
        {
            const int yt = top + q * k, yb = bot - q * k, xl = ox - q * k, xr = ox + q * k;
            if (yt >= c.miny)
            {
                if (xl >= c.minx) me_star_try(c, s, xl, yt, 0, d);
                if (xr <= c.maxx) me_star_try(c, s, xr, yt, 0, d);
            }
            if (yb <= c.maxy)
            {
                if (xl >= c.minx) me_star_try(c, s, xl, yb, 0, d);
                if (xr <= c.maxx) me_star_try(c, s, xr, yb, 0, d);
            }
        }
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
}

__constant__ int8_t c_hex2[8][2]    = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
__constant__ uint8_t c_mod6m1[8]    = { 5, 0, 1, 2, 3, 4, 5, 0 };
__constant__ int8_t c_square1[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
__constant__ int8_t c_star_off[16][2] = { {-1,0},{0,-1}, {-1,-1},{1,-1}, {-1,0},{1,0}, {-1,1},{-1,-1},
                                          {1,-1},{1,1}, {-1,0},{0,1}, {-1,1},{1,1}, {1,0},{0,1} };
// motion.cpp:48-58 {hpel_iters, hpel_dirs, qpel_iters, qpel_dirs, hpel_satd}
__constant__ int8_t c_workload[8][5] = { {1,4,0,4,0}, {1,4,1,4,0}, {1,4,1,4,1}, {2,4,1,4,1}, {2,4,2,4,1}, {1,8,1,8,1}, {2,8,1,8,1}, {2,8,2,8,1} };

template <typename P>
__device__ void me_run_job(MeCtx<P>& c, const x265cu_me_job& j, int32_t* __restrict__ out)
{
    const int qminx = c.minx * 4, qminy = c.miny * 4, qmaxx = c.maxx * 4, qmaxy = c.maxy * 4;
    const int merange = j.merange;
#define ME_YOK(y) (((y) >= c.miny) & ((y) <= c.maxy))
#define ME_INRANGE(x, y) ((x) >= c.minx && (x) <= c.maxx && (y) >= c.miny && (y) <= c.maxy)
    int pmvx = min(max(c.mvpx, qminx), qmaxx) , pmvy = min(max(c.mvpy, qminy), qmaxy);
    // NB: clipped() = min with max first, then max with min (mv.h:100-105); identical when min <= max
    pmvx = max(min(c.mvpx, qmaxx), qminx); pmvy = max(min(c.mvpy, qmaxy), qminy);
    int bestprex = pmvx, bestprey = pmvy;
    int bprecost = me_qpel_cost(c, pmvx, pmvy, false);
    int bmx = (pmvx + 2) >> 2, bmy = (pmvy + 2) >> 2;
    int bcost = bprecost;
    if ((pmvx & 3) | (pmvy & 3)) bcost = me_cost_fpel(c, bmx, bmy);
    if (pmvx | pmvy)
    {
        int cost = me_sad_direct(c, c.ref[0]) + me_mvcost(c, 0, 0);
        if (cost < bcost) { bcost = cost; bmx = 0; bmy = max(min(0, c.maxy), c.miny); }
    }
    for (int i = 0; i < j.numCand; i++)
    {
        int mx = max(min((int)j.mvc[2 * i], qmaxx), qminx), my = max(min((int)j.mvc[2 * i + 1], qmaxy), qminy);
        if ((mx | my) && !(mx == pmvx && my == pmvy) && !(mx == bestprex && my == bestprey))
        {
            int cost = me_subpel_compare(c, mx, my, false) + me_mvcost(c, mx, my);
            if (cost < bprecost) { bprecost = cost; bestprex = mx; bestprey = my; }
        }
    }

    if (j.method == 0)
    {   // DIA (motion.cpp:822-846)
        bcost <<= 4;
        int i = merange;
        do
        {
            int c0 = me_cost_fpel(c, bmx, bmy - 1), c1 = me_cost_fpel(c, bmx, bmy + 1);
            int c2 = me_cost_fpel(c, bmx - 1, bmy), c3 = me_cost_fpel(c, bmx + 1, bmy);
            if (ME_YOK(bmy - 1) && (c0 << 4) + 1 < bcost) bcost = (c0 << 4) + 1;
            if (ME_YOK(bmy + 1) && (c1 << 4) + 3 < bcost) bcost = (c1 << 4) + 3;
            if ((c2 << 4) + 4 < bcost) bcost = (c2 << 4) + 4;
            if ((c3 << 4) + 12 < bcost) bcost = (c3 << 4) + 12;
            if (!(bcost & 15)) break;
            bmx -= (int)((unsigned)bcost << 28) >> 30;
            bmy -= (int)((unsigned)bcost << 30) >> 30;
            bcost &= ~15;
        }
        while (--i && ME_INRANGE(bmx, bmy));
        bcost >>= 4;
    }
    else if (j.method == 1)
    {   // HEX (motion.cpp:848-945)
        int c0 = me_cost_fpel(c, bmx - 2, bmy), c1 = me_cost_fpel(c, bmx - 1, bmy + 2), c2 = me_cost_fpel(c, bmx + 1, bmy + 2);
        bcost <<= 3;
        if (ME_YOK(bmy) && (c0 << 3) + 2 < bcost) bcost = (c0 << 3) + 2;
        if (ME_YOK(bmy + 2))
        {
            if ((c1 << 3) + 3 < bcost) bcost = (c1 << 3) + 3;
            if ((c2 << 3) + 4 < bcost) bcost = (c2 << 3) + 4;
        }
        c0 = me_cost_fpel(c, bmx + 2, bmy); c1 = me_cost_fpel(c, bmx + 1, bmy - 2); c2 = me_cost_fpel(c, bmx - 1, bmy - 2);
        if (ME_YOK(bmy) && (c0 << 3) + 5 < bcost) bcost = (c0 << 3) + 5;
        if (ME_YOK(bmy - 2))
        {
            if ((c1 << 3) + 6 < bcost) bcost = (c1 << 3) + 6;
            if ((c2 << 3) + 7 < bcost) bcost = (c2 << 3) + 7;
        }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (ME_YOK(bmy + c_hex2[dir + 1][1]))
            {
                bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
                for (int i = (merange >> 1) - 1; i > 0 && ME_INRANGE(bmx, bmy); i--)
                {
                    c0 = me_cost_fpel(c, bmx + c_hex2[dir + 0][0], bmy + c_hex2[dir + 0][1]);
                    c1 = me_cost_fpel(c, bmx + c_hex2[dir + 1][0], bmy + c_hex2[dir + 1][1]);
                    c2 = me_cost_fpel(c, bmx + c_hex2[dir + 2][0], bmy + c_hex2[dir + 2][1]);
                    bcost &= ~7;
                    if (ME_YOK(bmy + c_hex2[dir + 0][1]) && (c0 << 3) + 1 < bcost) bcost = (c0 << 3) + 1;
                    if (ME_YOK(bmy + c_hex2[dir + 1][1]) && (c1 << 3) + 2 < bcost) bcost = (c1 << 3) + 2;
                    if (ME_YOK(bmy + c_hex2[dir + 2][1]) && (c2 << 3) + 3 < bcost) bcost = (c2 << 3) + 3;
                    if (!(bcost & 7)) break;
                    dir += (bcost & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    bmx += c_hex2[dir + 1][0]; bmy += c_hex2[dir + 1][1];
                }
            }
        }
        bcost >>= 3;
        int dir = 0;
        int s0 = me_cost_fpel(c, bmx, bmy - 1), s1 = me_cost_fpel(c, bmx, bmy + 1);
        int s2 = me_cost_fpel(c, bmx - 1, bmy), s3 = me_cost_fpel(c, bmx + 1, bmy);
        if (ME_YOK(bmy - 1) && s0 < bcost) { bcost = s0; dir = 1; }
        if (ME_YOK(bmy + 1) && s1 < bcost) { bcost = s1; dir = 2; }
        if (s2 < bcost) { bcost = s2; dir = 3; }
        if (s3 < bcost) { bcost = s3; dir = 4; }
        s0 = me_cost_fpel(c, bmx - 1, bmy - 1); s1 = me_cost_fpel(c, bmx - 1, bmy + 1);
        s2 = me_cost_fpel(c, bmx + 1, bmy - 1); s3 = me_cost_fpel(c, bmx + 1, bmy + 1);
        if (ME_YOK(bmy - 1) && s0 < bcost) { bcost = s0; dir = 5; }
        if (ME_YOK(bmy + 1) && s1 < bcost) { bcost = s1; dir = 6; }
        if (ME_YOK(bmy - 1) && s2 < bcost) { bcost = s2; dir = 7; }
        if (ME_YOK(bmy + 1) && s3 < bcost) { bcost = s3; dir = 8; }
        bmx += c_square1[dir][0]; bmy += c_square1[dir][1];
    }
    else
    {   // STAR (motion.cpp:1132-1240)
        MeStar s; s.bx = bmx; s.by = bmy; s.bcost = bcost; s.point = 0; s.dist = 0;
        me_star_pattern(c, s, 3, merange);
        bool done = false;
        if (s.dist == 1)
        {
            if (s.point)
            {
                const int saved = s.bcost;
                const int x1 = s.bx + c_star_off[(s.point - 1) * 2][0], y1 = s.by + c_star_off[(s.point - 1) * 2][1];
                const int x2 = s.bx + c_star_off[(s.point - 1) * 2 + 1][0], y2 = s.by + c_star_off[(s.point - 1) * 2 + 1][1];
                if (ME_INRANGE(x1, y1)) { int cost = me_cost_fpel(c, x1, y1); if (cost < s.bcost) { s.bcost = cost; s.bx = x1; s.by = y1; } }
                if (ME_INRANGE(x2, y2)) { int cost = me_cost_fpel(c, x2, y2); if (cost < s.bcost) { s.bcost = cost; s.bx = x2; s.by = y2; } }
                if (s.bcost == saved) done = true;
            }
            else done = true;
        }
        if (!done)
        {
            const int RD = 5;
            if (s.dist > RD)
            {
                for (int ty = c.miny; ty <= c.maxy; ty += RD)
                    for (int tx = c.minx; tx <= c.maxx; tx += RD)
                    {
                        if (tx + RD * 3 <= c.maxx)
                        {
                            for (int k = 0; k < 4; k++)
                            {
                                int sad = me_sad_direct(c, c.ref[0] + tx + (ptrdiff_t)ty * c.rstride);
                                int cost = sad + (k < 3 ? me_mvcost(c, tx * 4, ty * 4) : me_mvcost(c, tx * 8, ty * 8));
                                if (cost < s.bcost) { s.bcost = cost; s.bx = tx; s.by = ty; }
                                if (k < 3) tx += RD;
                            }
                        }
                        else
                        {
                            int cost = me_cost_fpel(c, tx, ty);
                            if (cost < s.bcost) { s.bcost = cost; s.bx = tx; s.by = ty; }
                        }
                    }
            }
            while (s.dist > 0)
            {
                s.dist = 0; s.point = 0;
                me_star_pattern(c, s, 32, merange);
                if (s.dist == 1)
                {
                    if (!s.point) break;
                    const int x1 = s.bx + c_star_off[(s.point - 1) * 2][0], y1 = s.by + c_star_off[(s.point - 1) * 2][1];
                    const int x2 = s.bx + c_star_off[(s.point - 1) * 2 + 1][0], y2 = s.by + c_star_off[(s.point - 1) * 2 + 1][1];
                    if (ME_INRANGE(x1, y1)) { int cost = me_cost_fpel(c, x1, y1); if (cost < s.bcost) { s.bcost = cost; s.bx = x1; s.by = y1; } }
                    if (ME_INRANGE(x2, y2)) { int cost = me_cost_fpel(c, x2, y2); if (cost < s.bcost) { s.bcost = cost; s.bx = x2; s.by = y2; } }
                    break;
                }
            }
        }
        bmx = s.bx; bmy = s.by; bcost = s.bcost;
    }

    int bx, by;
    if (bprecost < bcost) { bx = bestprex; by = bestprey; bcost = bprecost; }
    else { bx = bmx * 4; by = bmy * 4; }
    const int8_t* wl = c_workload[j.subme];
    if (!bcost)
        bcost = me_mvcost(c, bx, by);
    else if (c.lowres)
    {
        int bdir = 0;
        for (int i = 1; i <= wl[1]; i++)
        {
            int qx = bx + c_square1[i][0] * 2, qy = by + c_square1[i][1] * 2;
            if ((qy < qminy) | (qy > qmaxy)) continue;
            int cost = me_lowres_cost(c, qx, qy, false) + me_mvcost(c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bx += c_square1[bdir][0] * 2; by += c_square1[bdir][1] * 2;
        bcost = me_lowres_cost(c, bx, by, true) + me_mvcost(c, bx, by);
        bdir = 0;
        for (int i = 1; i <= wl[3]; i++)
        {
            int qx = bx + c_square1[i][0], qy = by + c_square1[i][1];
            if ((qy < qminy) | (qy > qmaxy)) continue;
            int cost = me_lowres_cost(c, qx, qy, true) + me_mvcost(c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bx += c_square1[bdir][0]; by += c_square1[bdir][1];
    }
    else
    {
        bool hsatd = false;
        if (wl[4]) { bcost = me_subpel_compare(c, bx, by, true) + me_mvcost(c, bx, by); hsatd = true; }
        for (int it = 0; it < wl[0]; it++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl[1]; i++)
            {
                int qx = bx + c_square1[i][0] * 2, qy = by + c_square1[i][1] * 2;
                if ((qy < qminy) | (qy > qmaxy)) continue;
                int cost = me_subpel_compare(c, qx, qy, hsatd) + me_mvcost(c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bx += c_square1[bdir][0] * 2; by += c_square1[bdir][1] * 2; }
            else break;
        }
        if (!wl[4]) bcost = me_subpel_compare(c, bx, by, true) + me_mvcost(c, bx, by);
        for (int it = 0; it < wl[2]; it++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl[3]; i++)
            {
                int qx = bx + c_square1[i][0], qy = by + c_square1[i][1];
                if ((qy < qminy) | (qy > qmaxy)) continue;
                int cost = me_subpel_compare(c, qx, qy, true) + me_mvcost(c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bx += c_square1[bdir][0]; by += c_square1[bdir][1]; }
            else break;
        }
    }
    if (c.lane == 0) { out[0] = bcost; out[1] = bx; out[2] = by; out[3] = 0; }
#undef ME_YOK
#undef ME_INRANGE
}

// persistent warps, dynamic job fetch (jobs differ by up to 64x in work)
template <typename P>
__global__ void __launch_bounds__(256) k_me(const P* __restrict__ fenc, int fstride, const P* const* __restrict__ refs, int rstride, int lowres,
                                            const uint16_t* __restrict__ mvcost, const x265cu_me_job* __restrict__ jobs, int n,
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
        const x265cu_me_job j = jobs[jid];
        MeCtx<P> c;
        c.fenc = fenc + j.offset; c.fstride = fstride;
        if (lowres) { for (int i = 0; i < 4; i++) c.ref[i] = refs[j.ref * 4 + i] + j.offset; }
        else        { c.ref[0] = refs[j.ref] + j.offset; c.ref[1] = c.ref[2] = c.ref[3] = c.ref[0]; }
        c.rstride = rstride; c.mvc = mvcost;
        c.mvpx = j.qmvp[0]; c.mvpy = j.qmvp[1];
        c.minx = j.mvmin[0]; c.miny = j.mvmin[1]; c.maxx = j.mvmax[0]; c.maxy = j.mvmax[1];
        c.w = j.pw; c.h = j.ph; c.lane = lane; c.lowres = lowres; c.sm = sm;
        me_run_job<P>(c, j, out + (size_t)jid * 4);
        __syncwarp();
    }
}

static int launch_me(x265cu_ctx* ctx, int depth, const void* fenc, int fstride, const void* const* refs, int rstride, int lowres,
                     const uint16_t* mvcost, const x265cu_me_job* jobs, int n, int32_t* out, int* counter_dev)
{
    if (n <= 0) return 0;
    CU_CHECK(cudaMemsetAsync(counter_dev, 0, sizeof(int), ctx->stream));
    const int threads = 256, warps = threads / 32;
    const size_t smem = sizeof(MeShared) * warps;
    int blocks = ctx->sm_count * 4;
    int need = (n + warps - 1) / warps;
    if (blocks > need) blocks = need;
    if (depth == 8)
        k_me<uint8_t><<<blocks, threads, smem, ctx->stream>>>((const uint8_t*)fenc, fstride, (const uint8_t* const*)refs, rstride, lowres, mvcost, jobs, n, out, counter_dev);
    else
        k_me<uint16_t><<<blocks, threads, smem, ctx->stream>>>((const uint16_t*)fenc, fstride, (const uint16_t* const*)refs, rstride, lowres, mvcost, jobs, n, out, counter_dev);
    CU_LAUNCH_CHECK(ctx);
    return 0;
}

/* oracle/oracle_me.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates the callers whose decisions the batched device kernels must reproduce:
 *   BitCost tables            encoder/bitcost.cpp:33-110
 *   MotionEstimate::motionEstimate  encoder/motion.cpp:739-1569 (DIA, HEX, STAR; luma only)
 *   StarPatternSearch         encoder/motion.cpp:362-604
 *   subpelCompare             encoder/motion.cpp:1571-1664 (luma part)
 *   lowresQPelCost            common/lowres.h:94-120
 * Pinned against the real MotionEstimate through oracle/_ref (x265ref_motion_estimate).
 */
#include <stdlib.h>
#include "oracle.h"
#include <math.h>
#include <string.h>

/* bitcost.cpp:95-109 + :49-55: s_bitsizes is float, lambda double, result capped at 2^15-1 */
void orc_mvcost_table(double lambda, int range, uint16_t* out)
{
    float log2_2 = 2.0f / logf(2.0f);
    for (int i = 0; i <= range; i++)
    {
        float bits = i ? logf((float)(i + 1)) * log2_2 + 1.718f : 0.718f;
        double v = bits * lambda + 0.5f;
        if (v > 32767.0) v = 32767.0;
        uint16_t c = (uint16_t)v;
        out[range + i] = out[range - i] = c;
    }
}

/* call counters for workload characterisation (not thread safe; tests/bench never rely on them) */
long long g_orc_cnt[4];   /* fpel SAD evaluations, sub-pel compares, raster points, jobs */

typedef struct { int x, y; } mv_t;
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
typedef int (*cmp_fn)(const pixel*, intptr_t, const pixel*, intptr_t, int, int);

typedef struct {
    const orc_me_job* j;
    pixel fenc[64 * 64];        /* FENC_STRIDE cache (motion.cpp:189) */
    const pixel* fref;          /* ref plane + blockOffset */
    intptr_t stride;
    mv_t mvp;                   /* qpel */
    mv_t mvmin, mvmax;
    int chroma;                 /* bChromaSATD (motion.cpp:212) */
    pixel fencCb[32 * 32], fencCr[32 * 32];   /* fencPUYuv chroma, stride m_csize = 32 (yuv.cpp:126-140) */
    const pixel* refCb; const pixel* refCr;    /* ref chroma planes + the PU's chroma offset */
    intptr_t cstride;
} me_ctx;

static inline int mvcost(const me_ctx* c, int qx, int qy)
{
    return (uint16_t)(c->j->mvcost[qx - c->mvp.x] + c->j->mvcost[qy - c->mvp.y]);
}
static inline int in_range(const me_ctx* c, int x, int y)
{ return x >= c->mvmin.x && x <= c->mvmax.x && y >= c->mvmin.y && y <= c->mvmax.y; }
static inline int fpel_sad(const me_ctx* c, int x, int y)
{ g_orc_cnt[0]++; return orc_sad(c->fenc, 64, c->fref + x + (intptr_t)y * c->stride, c->stride, c->j->pw, c->j->ph); }
static inline int cost_fpel(const me_ctx* c, int x, int y)
{ return fpel_sad(c, x, y) + mvcost(c, x * 4, y * 4); }

/* lowres.h:94-120 */
static int lowres_qpel_cost(const me_ctx* c, int qx, int qy, cmp_fn cmp)
{
    const orc_me_job* j = c->j;
    intptr_t st = j->refStride;
    if ((qx | qy) & 1)
    {
        pixel buf[8 * 8];
        int ha = (qy & 2) | ((qx & 2) >> 1);
        const pixel* a = j->ref[ha] + j->offset + (qx >> 2) + (intptr_t)(qy >> 2) * st;
        int rx = qx + (qx & 1), ry = qy + (qy & 1);
        int hb = (ry & 2) | ((rx & 2) >> 1);
        const pixel* b = j->ref[hb] + j->offset + (rx >> 2) + (intptr_t)(ry >> 2) * st;
        orc_pixelavg_pp(buf, 8, a, st, b, st, 8, 8);
        return cmp(c->fenc, 64, buf, 8, j->pw, j->ph);
    }
    int hp = (qy & 2) | ((qx & 2) >> 1);
    const pixel* r = j->ref[hp] + j->offset + (qx >> 2) + (intptr_t)(qy >> 2) * st;
    return cmp(c->fenc, 64, r, st, j->pw, j->ph);
}

/* motion.cpp:1601-1661, 4:2:0 (hshift = vshift = 1: mvx = qmv.x, eighth-pel): Cb + Cr SATD of the chroma block.
 * chromaSatd = the luma SATD of the chroma-sized block (primitives.cpp:139-158), i.e. orc_satd's tiling of (w/2, h/2);
 * hv = hps(rowExt) into a (h/2 + 3)-row intermediate, then vsp from its second row (ipfilter.cpp:362-369 pattern
 * with the 4-tap filter: motion.cpp:1648-1658). */
static int chroma_cost(const me_ctx* c, int qx, int qy)
{
    const int cw = c->j->pw >> 1, ch = c->j->ph >> 1;
    const int xf = qx & 7, yf = qy & 7;
    const intptr_t off = (qx >> 3) + (intptr_t)(qy >> 3) * c->cstride;
    const pixel* ref[2] = { c->refCb + off, c->refCr + off };
    const pixel* fenc[2] = { c->fencCb, c->fencCr };
    int cost = 0;
    for (int p = 0; p < 2; p++)
    {
        if (!(xf | yf)) { cost += orc_satd(fenc[p], 32, ref[p], c->cstride, cw, ch); continue; }
        pixel buf[32 * 32];
        if (!yf)      orc_interp_hpp(ref[p], c->cstride, buf, cw, xf, 4, cw, ch);
        else if (!xf) orc_interp_vpp(ref[p], c->cstride, buf, cw, yf, 4, cw, ch);
        else
        {
            int16_t immed[32 * (32 + 3)];
            orc_interp_hps(ref[p], c->cstride, immed, cw, xf, 1, 4, cw, ch);
            orc_interp_vsp(immed + cw, cw, buf, cw, yf, 4, cw, ch);
        }
        cost += orc_satd(fenc[p], 32, buf, cw, cw, ch);
    }
    return cost;
}

/* motion.cpp:1571-1664 */
static int subpel_compare(const me_ctx* c, int qx, int qy, cmp_fn cmp)
{
    g_orc_cnt[1]++;
    const orc_me_job* j = c->j;
    const pixel* r = c->fref + (qx >> 2) + (intptr_t)(qy >> 2) * c->stride;
    int xf = qx & 3, yf = qy & 3;
    int cost;
    if (!(xf | yf))
        cost = cmp(c->fenc, 64, r, c->stride, j->pw, j->ph);
    else
    {
        pixel buf[64 * 64];
        if (!yf)      orc_interp_hpp(r, c->stride, buf, j->pw, xf, 8, j->pw, j->ph);
        else if (!xf) orc_interp_vpp(r, c->stride, buf, j->pw, yf, 8, j->pw, j->ph);
        else          orc_interp_hvpp(r, c->stride, buf, j->pw, xf, yf, 8, j->pw, j->ph);
        cost = cmp(c->fenc, 64, buf, j->pw, j->pw, j->ph);
    }
    if (c->chroma) cost += chroma_cost(c, qx, qy);
    return cost;
}

static int qpel_cost(const me_ctx* c, int qx, int qy, cmp_fn cmp)
{ return c->j->lowres ? lowres_qpel_cost(c, qx, qy, cmp) : subpel_compare(c, qx, qy, cmp); }

typedef struct { mv_t bmv; int bcost, point, dist; } star_t;

static inline void star_try(const me_ctx* c, star_t* s, int x, int y, int point, int dist)
{
    int cost = cost_fpel(c, x, y);
    if (cost < s->bcost) { s->bcost = cost; s->bmv.x = x; s->bmv.y = y; s->point = point; s->dist = dist; }
}

/* motion.cpp:362-604.  The reference's x4 fast path evaluates the same points in the same order
 * as the bounds-checked path, so a single checked walk reproduces both. */
static void star_pattern(const me_ctx* c, star_t* s, int earlyExitIters, int merange)
{
    const mv_t o = s->bmv;
    int saved = s->bcost, rounds = 0;
    {
        const int d = 1;
        int top = o.y - d, bot = o.y + d, lft = o.x - d, rgt = o.x + d;
        if (top >= c->mvmin.y) star_try(c, s, o.x, top, 2, d);
        if (lft >= c->mvmin.x) star_try(c, s, lft, o.y, 4, d);
        if (rgt <= c->mvmax.x) star_try(c, s, rgt, o.y, 5, d);
        if (bot <= c->mvmax.y) star_try(c, s, o.x, bot, 7, d);
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int d = 2; d <= 8; d <<= 1)
    {
        int top = o.y - d, bot = o.y + d, lft = o.x - d, rgt = o.x + d;
        int top2 = o.y - (d >> 1), bot2 = o.y + (d >> 1), lft2 = o.x - (d >> 1), rgt2 = o.x + (d >> 1);
        saved = s->bcost;
        if (top >= c->mvmin.y && lft >= c->mvmin.x && rgt <= c->mvmax.x && bot <= c->mvmax.y)
        {
            /* order of the two x4 bursts (motion.cpp:456-463) */
            star_try(c, s, o.x, top, 2, d);   star_try(c, s, lft2, top2, 1, d >> 1);
            star_try(c, s, rgt2, top2, 3, d >> 1); star_try(c, s, lft, o.y, 4, d);
            star_try(c, s, rgt, o.y, 5, d);   star_try(c, s, lft2, bot2, 6, d >> 1);
            star_try(c, s, rgt2, bot2, 8, d >> 1); star_try(c, s, o.x, bot, 7, d);
        }
        else
        {
            if (top >= c->mvmin.y) star_try(c, s, o.x, top, 2, d);
            if (top2 >= c->mvmin.y)
            {
                if (lft2 >= c->mvmin.x) star_try(c, s, lft2, top2, 1, d >> 1);
                if (rgt2 <= c->mvmax.x) star_try(c, s, rgt2, top2, 3, d >> 1);
            }
            if (lft >= c->mvmin.x) star_try(c, s, lft, o.y, 4, d);
            if (rgt <= c->mvmax.x) star_try(c, s, rgt, o.y, 5, d);
            if (bot2 <= c->mvmax.y)
            {
                if (lft2 >= c->mvmin.x) star_try(c, s, lft2, bot2, 6, d >> 1);
                if (rgt2 <= c->mvmax.x) star_try(c, s, rgt2, bot2, 8, d >> 1);
            }
            if (bot <= c->mvmax.y) star_try(c, s, o.x, bot, 7, d);
        }
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int d = 16; d <= merange; d <<= 1)
    {
        int top = o.y - d, bot = o.y + d, lft = o.x - d, rgt = o.x + d;
        saved = s->bcost;
        int inside = top >= c->mvmin.y && lft >= c->mvmin.x && rgt <= c->mvmax.x && bot <= c->mvmax.y;
        if (inside || top >= c->mvmin.y) star_try(c, s, o.x, top, 0, d);
        if (inside || lft >= c->mvmin.x) star_try(c, s, lft, o.y, 0, d);
        if (inside || rgt <= c->mvmax.x) star_try(c, s, rgt, o.y, 0, d);
        if (inside || bot <= c->mvmax.y) star_try(c, s, o.x, bot, 0, d);
        for (int k = 1; k < 4; k++)
        {
            int yt = top + (d >> 2) * k, yb = bot - (d >> 2) * k;
            int xl = o.x - (d >> 2) * k, xr = o.x + (d >> 2) * k;
            if (inside || yt >= c->mvmin.y)
            {
                if (inside || xl >= c->mvmin.x) star_try(c, s, xl, yt, 0, d);
                if (inside || xr <= c->mvmax.x) star_try(c, s, xr, yt, 0, d);
            }
            if (inside || yb <= c->mvmax.y)
            {
                if (inside || xl >= c->mvmin.x) star_try(c, s, xl, yb, 0, d);
                if (inside || xr <= c->mvmax.x) star_try(c, s, xr, yb, 0, d);
            }
        }
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
}

static const mv_t k_hex2[8] = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
static const uint8_t k_mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };
static const mv_t k_square1[9] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
static const mv_t k_offsets[16] = { {-1,0},{0,-1}, {-1,-1},{1,-1}, {-1,0},{1,0}, {-1,1},{-1,-1},
                                    {1,-1},{1,1}, {-1,0},{0,1}, {-1,1},{1,1}, {1,0},{0,1} };
static const struct { int hpel_iters, hpel_dirs, qpel_iters, qpel_dirs, hpel_satd; } k_workload[8] = {
    {1,4,0,4,0}, {1,4,1,4,0}, {1,4,1,4,1}, {2,4,1,4,1}, {2,4,2,4,1}, {1,8,1,8,1}, {2,8,1,8,1}, {2,8,2,8,1} };

static inline int yok(const me_ctx* c, int y) { return (y >= c->mvmin.y) & (y <= c->mvmax.y); }

/* ---- X265_UMH_SEARCH helpers (motion.cpp:226-360 macros as functions over a small state) ---- */
typedef struct { const me_ctx* c; mv_t bmv, omv; int bcost; } umh_t;
static void umh_cost_mv(umh_t* u, int x, int y)                       /* COST_MV */
{
    int cost = cost_fpel(u->c, x, y);
    if (cost < u->bcost) { u->bcost = cost; u->bmv.x = x; u->bmv.y = y; }
}
static void umh_x4(umh_t* u, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3)   /* COST_MV_X4: around omv, y range only */
{
    const int o[8] = { x0, y0, x1, y1, x2, y2, x3, y3 };
    int costs[4];
    for (int k = 0; k < 4; k++) costs[k] = cost_fpel(u->c, u->omv.x + o[2 * k], u->omv.y + o[2 * k + 1]);
    for (int k = 0; k < 4; k++)
        if (yok(u->c, u->omv.y + o[2 * k + 1]) && costs[k] < u->bcost)
        { u->bcost = costs[k]; u->bmv.x = u->omv.x + o[2 * k]; u->bmv.y = u->omv.y + o[2 * k + 1]; }
}
static void umh_cross(umh_t* u, int start, int x_max, int y_max)     /* CROSS (motion.cpp:336-360) */
{
    const me_ctx* c = u->c;
    int16_t i = (int16_t)start;
    if (x_max <= imin(c->mvmax.x - u->omv.x, u->omv.x - c->mvmin.x))
        for (; i < x_max - 2; i += 4) umh_x4(u, i, 0, -i, 0, i + 2, 0, -i - 2, 0);
    for (; i < x_max; i += 2)
    {
        if (u->omv.x + i <= c->mvmax.x) umh_cost_mv(u, u->omv.x + i, u->omv.y);
        if (u->omv.x - i >= c->mvmin.x) umh_cost_mv(u, u->omv.x - i, u->omv.y);
    }
    i = (int16_t)start;
    if (y_max <= imin(c->mvmax.y - u->omv.y, u->omv.y - c->mvmin.y))
        for (; i < y_max - 2; i += 4) umh_x4(u, 0, i, 0, -i, 0, i + 2, 0, -i - 2);
    for (; i < y_max; i += 2)
    {
        if (u->omv.y + i <= c->mvmax.y) umh_cost_mv(u, u->omv.x, u->omv.y + i);
        if (u->omv.y - i >= c->mvmin.y) umh_cost_mv(u, u->omv.x, u->omv.y - i);
    }
}

int orc_motion_estimate(const orc_me_job* j, int* outQMv)
{
    me_ctx ctx; me_ctx* c = &ctx;
    c->j = j;
    g_orc_cnt[3]++;
    /* SEA sums source sub-blocks at offsets that can lie OUTSIDE a narrow PU (e.g. fenc + 8 for an 8x32 PU, motion.cpp:1304-1311):
     * the reference reads whatever its 64-stride PU cache holds there.  The pinned model is a cache that was zero before the PU
     * was copied in (the shim zeroes MotionEstimate::fencPUYuv before setSourcePU); only method 4 ever looks there. */
    if (j->method == 4) memset(c->fenc, 0, sizeof(c->fenc));
    orc_copy_pp(c->fenc, 64, j->fenc + j->offset, j->fencStride, j->pw, j->ph);
    c->fref = j->ref[0] + j->offset;
    c->stride = j->refStride;
    /* motion.cpp:204-212: chroma residual cost if subme > 2 and the chroma block has a SATD primitive */
    c->chroma = j->chroma && !j->lowres && j->subme > 2 && !(j->pw & 7) && !(j->ph & 7);
    if (c->chroma)
    {
        const intptr_t py = j->offset / j->fencStride, px = j->offset % j->fencStride;
        const intptr_t coff = (py >> 1) * j->cstride + (px >> 1);
        orc_copy_pp(c->fencCb, 32, j->fencC[0] + coff, j->cstride, j->pw >> 1, j->ph >> 1);
        orc_copy_pp(c->fencCr, 32, j->fencC[1] + coff, j->cstride, j->pw >> 1, j->ph >> 1);
        c->refCb = j->refC[0] + coff; c->refCr = j->refC[1] + coff; c->cstride = j->cstride;
    }
    c->mvp.x = j->qmvp[0]; c->mvp.y = j->qmvp[1];
    c->mvmin.x = j->mvmin[0]; c->mvmin.y = j->mvmin[1];
    c->mvmax.x = j->mvmax[0]; c->mvmax.y = j->mvmax[1];
    const int qminx = c->mvmin.x * 4, qminy = c->mvmin.y * 4, qmaxx = c->mvmax.x * 4, qmaxy = c->mvmax.y * 4;
    const int merange = j->merange;
#define CLIPQ(vx, vy) do { if (vx > qmaxx) vx = qmaxx; if (vy > qmaxy) vy = qmaxy; if (vx < qminx) vx = qminx; if (vy < qminy) vy = qminy; } while (0)

    /* motion.cpp:771-792: cost at the clipped qpel MVP, then at its full-pel rounding */
    int pmvx = c->mvp.x, pmvy = c->mvp.y;
    CLIPQ(pmvx, pmvy);
    int bestprex = pmvx, bestprey = pmvy;
    int bprecost = qpel_cost(c, pmvx, pmvy, orc_sad);      /* NB: no mvcost added (motion.cpp:776-779) */
    mv_t bmv = { (pmvx + 2) >> 2, (pmvy + 2) >> 2 };
    int bcost = bprecost;
    if ((pmvx & 3) | (pmvy & 3))
        bcost = cost_fpel(c, bmv.x, bmv.y);
    /* :787-797 MV(0) */
    if (pmvx | pmvy)
    {
        int cost = fpel_sad(c, 0, 0) + mvcost(c, 0, 0);
        if (cost < bcost)
        {
            bcost = cost;
            bmv.x = 0;
            int zy = 0 < c->mvmax.y ? 0 : c->mvmax.y;
            bmv.y = zy > c->mvmin.y ? zy : c->mvmin.y;
        }
    }
    /* :801-814 qpel candidates */
    for (int i = 0; i < j->numCand; i++)
    {
        int mx = j->mvc[2 * i], my = j->mvc[2 * i + 1];
        CLIPQ(mx, my);
        if ((mx | my) && !(mx == pmvx && my == pmvy) && !(mx == bestprex && my == bestprey))
        {
            int cost = subpel_compare(c, mx, my, orc_sad) + mvcost(c, mx, my);
            if (cost < bprecost) { bprecost = cost; bestprex = mx; bestprey = my; }
        }
    }

    int hexRange = merange;                 /* UMH hands its adapted range to the hexagon search it falls into */
    switch (j->method)
    {
    case 0: /* X265_DIA_SEARCH motion.cpp:822-846 */
    {
        bcost <<= 4;
        int i = merange;
        do
        {
            int c0 = cost_fpel(c, bmv.x, bmv.y - 1), c1 = cost_fpel(c, bmv.x, bmv.y + 1);
            int c2 = cost_fpel(c, bmv.x - 1, bmv.y), c3 = cost_fpel(c, bmv.x + 1, bmv.y);
            if (yok(c, bmv.y - 1) && (c0 << 4) + 1 < bcost) bcost = (c0 << 4) + 1;
            if (yok(c, bmv.y + 1) && (c1 << 4) + 3 < bcost) bcost = (c1 << 4) + 3;
            if ((c2 << 4) + 4 < bcost) bcost = (c2 << 4) + 4;
            if ((c3 << 4) + 12 < bcost) bcost = (c3 << 4) + 12;
            if (!(bcost & 15)) break;
            bmv.x -= (int)((unsigned)bcost << 28) >> 30;
            bmv.y -= (int)((unsigned)bcost << 30) >> 30;
            bcost &= ~15;
        }
        while (--i && in_range(c, bmv.x, bmv.y));
        bcost >>= 4;
        break;
    }
    case 2: /* X265_UMH_SEARCH motion.cpp:946-1130: predictor refinement, early termination, cross, 5x5, hexagon grid, then me_hex2 */
    {
        static const mv_t hex4[16] = { {0,-4}, {0,4}, {-2,-3}, {2,-3}, {-4,-2}, {4,-2}, {-4,-1}, {4,-1}, {-4,0}, {4,0}, {-4,1}, {4,1}, {-4,2}, {4,2}, {-2,3}, {2,3} };
        const int scale = (j->ph * j->ph) >> 4;                         /* sizeScale[partEnum] = (H * H) >> 4 (motion.cpp:123-152) */
#define SAD_THRESH(v) (u.bcost < (((v) >> 4) * scale))
        const mv_t pmv = { (pmvx + 2) >> 2, (pmvy + 2) >> 2 };          /* pmv.roundToFPel() (motion.cpp:816) */
        umh_t u; u.c = c; u.bmv = bmv; u.bcost = bcost;
        int um = merange, cross_start = 1, done = 0;
        u.omv = u.bmv;
        const int ucost1 = u.bcost;
        u.omv = pmv; umh_x4(&u, 0, -1, 0, 1, -1, 0, 1, 0);              /* DIA1_ITER(pmv) */
        if (pmv.x | pmv.y) { u.omv.x = 0; u.omv.y = 0; umh_x4(&u, 0, -1, 0, 1, -1, 0, 1, 0); }
        const int ucost2 = u.bcost;
        if ((u.bmv.x | u.bmv.y) && !(u.bmv.x == pmv.x && u.bmv.y == pmv.y)) { u.omv = u.bmv; umh_x4(&u, 0, -1, 0, 1, -1, 0, 1, 0); }
        if (u.bcost == ucost2) cross_start = 3;
        u.omv = u.bmv;
        if (u.bcost == ucost2 && SAD_THRESH(2000))
        {
            umh_x4(&u, 0, -2, -1, -1, 1, -1, -2, 0);
            umh_x4(&u, 2, 0, -1, 1, 1, 1, 0, 2);
            if (u.bcost == ucost1 && SAD_THRESH(500)) done = 1;
            else if (u.bcost == ucost2)
            {
                const int range = (int16_t)(um >> 1) | 1;
                umh_cross(&u, 3, range, range);
                umh_x4(&u, -1, -2, 1, -2, -2, -1, 2, -1);
                umh_x4(&u, -2, 1, 2, 1, -1, 2, 1, 2);
                if (u.bcost == ucost2) done = 1;
                cross_start = range + 2;
            }
        }
        if (done) { bmv = u.bmv; bcost = u.bcost; break; }
        if (j->numCand)
        {   /* adaptive search range (motion.cpp:986-1037) */
            static const uint8_t range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };
            const int is64 = j->pw == 64 && j->ph == 64;
            int mvd, denom = 1;
            if (j->numCand == 1)
                mvd = is64 ? 25 : abs(j->qmvp[0] - j->mvc[0]) + abs(j->qmvp[1] - j->mvc[1]);
            else
            {
                denom = j->numCand - 1;
                mvd = 0;
                if (!is64) { mvd = abs(j->qmvp[0] - j->mvc[0]) + abs(j->qmvp[1] - j->mvc[1]); denom++; }
                for (int i = 0; i < j->numCand - 1; i++)                /* predictorDifference (motion.cpp:87-98) */
                    mvd += abs(j->mvc[2 * i] - j->mvc[2 * i + 2]) + abs(j->mvc[2 * i + 1] - j->mvc[2 * i + 3]);
            }
            const int sad_ctx = SAD_THRESH(1000) ? 0 : SAD_THRESH(2000) ? 1 : SAD_THRESH(4000) ? 2 : 3;
            const int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            um = (um * range_mul[mvd_ctx][sad_ctx]) >> 2;
        }
        umh_cross(&u, cross_start, um, um >> 1);
        umh_x4(&u, -2, -2, -2, 2, 2, -2, 2, 2);
        /* hexagon grid (motion.cpp:1045-1124) */
        u.omv = u.bmv;
        uint16_t i = 1;
        do
        {
            const int lim = imin(imin(c->mvmax.x - u.omv.x, u.omv.x - c->mvmin.x), imin(c->mvmax.y - u.omv.y, u.omv.y - c->mvmin.y));
            if (4 * i > lim)
            {
                for (int k = 0; k < 16; k++)
                {
                    const int x = u.omv.x + hex4[k].x * i, y = u.omv.y + hex4[k].y * i;
                    if (in_range(c, x, y)) umh_cost_mv(&u, x, y);
                }
            }
            else
            {
                int dir = -1;
                for (int k = 0; k < 16; k++)
                {
                    const int cost = cost_fpel(c, u.omv.x + hex4[k].x * i, u.omv.y + hex4[k].y * i);
                    if (yok(c, u.omv.y + hex4[k].y) && cost < u.bcost) { u.bcost = cost; dir = k; }   /* MIN_MV: the y test uses the UNSCALED offset */
                }
                if (dir >= 0) { u.bmv.x = u.omv.x + i * hex4[dir].x; u.bmv.y = u.omv.y + i * hex4[dir].y; }
            }
        }
        while (++i <= um >> 2);
#undef SAD_THRESH
        bmv = u.bmv; bcost = u.bcost; hexRange = um;
        if (!in_range(c, bmv.x, bmv.y)) break;
    }
    /* fall through: goto me_hex2 (motion.cpp:1125-1127) */
    case 1: /* X265_HEX_SEARCH motion.cpp:848-945 */
    {
        int c0 = cost_fpel(c, bmv.x - 2, bmv.y), c1 = cost_fpel(c, bmv.x - 1, bmv.y + 2), c2 = cost_fpel(c, bmv.x + 1, bmv.y + 2);
        bcost <<= 3;
        if (yok(c, bmv.y) && (c0 << 3) + 2 < bcost) bcost = (c0 << 3) + 2;
        if (yok(c, bmv.y + 2))
        {
            if ((c1 << 3) + 3 < bcost) bcost = (c1 << 3) + 3;
            if ((c2 << 3) + 4 < bcost) bcost = (c2 << 3) + 4;
        }
        c0 = cost_fpel(c, bmv.x + 2, bmv.y); c1 = cost_fpel(c, bmv.x + 1, bmv.y - 2); c2 = cost_fpel(c, bmv.x - 1, bmv.y - 2);
        if (yok(c, bmv.y) && (c0 << 3) + 5 < bcost) bcost = (c0 << 3) + 5;
        if (yok(c, bmv.y - 2))
        {
            if ((c1 << 3) + 6 < bcost) bcost = (c1 << 3) + 6;
            if ((c2 << 3) + 7 < bcost) bcost = (c2 << 3) + 7;
        }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (yok(c, bmv.y + k_hex2[dir + 1].y))
            {
                bmv.x += k_hex2[dir + 1].x; bmv.y += k_hex2[dir + 1].y;
                for (int i = (hexRange >> 1) - 1; i > 0 && in_range(c, bmv.x, bmv.y); i--)
                {
                    c0 = cost_fpel(c, bmv.x + k_hex2[dir + 0].x, bmv.y + k_hex2[dir + 0].y);
                    c1 = cost_fpel(c, bmv.x + k_hex2[dir + 1].x, bmv.y + k_hex2[dir + 1].y);
                    c2 = cost_fpel(c, bmv.x + k_hex2[dir + 2].x, bmv.y + k_hex2[dir + 2].y);
                    bcost &= ~7;
                    if (yok(c, bmv.y + k_hex2[dir + 0].y) && (c0 << 3) + 1 < bcost) bcost = (c0 << 3) + 1;
                    if (yok(c, bmv.y + k_hex2[dir + 1].y) && (c1 << 3) + 2 < bcost) bcost = (c1 << 3) + 2;
                    if (yok(c, bmv.y + k_hex2[dir + 2].y) && (c2 << 3) + 3 < bcost) bcost = (c2 << 3) + 3;
                    if (!(bcost & 7)) break;
                    dir += (bcost & 7) - 2;
                    dir = k_mod6m1[dir + 1];
                    bmv.x += k_hex2[dir + 1].x; bmv.y += k_hex2[dir + 1].y;
                }
            }
        }
        bcost >>= 3;
        /* square refine :921-941 */
        int dir = 0;
        int s0 = cost_fpel(c, bmv.x, bmv.y - 1), s1 = cost_fpel(c, bmv.x, bmv.y + 1);
        int s2 = cost_fpel(c, bmv.x - 1, bmv.y), s3 = cost_fpel(c, bmv.x + 1, bmv.y);
        if (yok(c, bmv.y - 1) && s0 < bcost) { bcost = s0; dir = 1; }
        if (yok(c, bmv.y + 1) && s1 < bcost) { bcost = s1; dir = 2; }
        if (s2 < bcost) { bcost = s2; dir = 3; }
        if (s3 < bcost) { bcost = s3; dir = 4; }
        s0 = cost_fpel(c, bmv.x - 1, bmv.y - 1); s1 = cost_fpel(c, bmv.x - 1, bmv.y + 1);
        s2 = cost_fpel(c, bmv.x + 1, bmv.y - 1); s3 = cost_fpel(c, bmv.x + 1, bmv.y + 1);
        if (yok(c, bmv.y - 1) && s0 < bcost) { bcost = s0; dir = 5; }
        if (yok(c, bmv.y + 1) && s1 < bcost) { bcost = s1; dir = 6; }
        if (yok(c, bmv.y - 1) && s2 < bcost) { bcost = s2; dir = 7; }
        if (yok(c, bmv.y + 1) && s3 < bcost) { bcost = s3; dir = 8; }
        bmv.x += k_square1[dir].x; bmv.y += k_square1[dir].y;
        break;
    }
    case 4: /* X265_SEA motion.cpp:1242-1395: successive elimination on the integral planes, literally (incl. its quirks: the row
             * cost p_cost_mvy[tmv.y] << 2 indexes the quarter-pel table with a full-pel value; the scanned width is rounded up to 4) */
    {
        const int w = j->pw, h = j->ph;
        const int minX = imax(bmv.x - merange, c->mvmin.x), minY = imax(bmv.y - merange, c->mvmin.y);      /* omv = bmv */
        const int maxX = imin(bmv.x + merange, c->mvmax.x), maxY = imin(bmv.y + merange, c->mvmax.y);
        const int meRangeWidth = (maxX - minX + 3) & ~3;
        int16_t* scratch = (int16_t*)calloc((size_t)(merange * 2 + 4 > meRangeWidth + 4 ? merange * 2 + 4 : meRangeWidth + 4), sizeof(int16_t));
        int deltaX = (w <= 8) ? w : (w >> 1), deltaY = (h <= 8) ? h : (h >> 1);
        const int smallRect = (w == 4 && h == 4) || (w == 16 && h == 12) || (w == 12 && h == 16) || (w == 16 && h == 4) || (w == 4 && h == 16);
        const int verticalRect = (w == 32 && h == 64) || (w == 16 && h == 32) || (w == 8 && h == 16) || (w == 4 && h == 8);
        const int horizontalRect = (w == 64 && h == 32) || (w == 32 && h == 16) || (w == 16 && h == 8) || (w == 8 && h == 4);
        const int asymV = (w == 12 && h == 16) || (w == 4 && h == 16) || (w == 24 && h == 32) || (w == 8 && h == 32) || (w == 48 && h == 64) || (w == 16 && h == 64);
        const int asymH = (w == 16 && h == 12) || (w == 16 && h == 4) || (w == 32 && h == 24) || (w == 32 && h == 8) || (w == 64 && h == 48) || (w == 64 && h == 16);
        int tw, th;                                        /* the sub-block whose four sums are the source's "DC"s */
        if (verticalRect) { tw = w; th = h >> 1; }
        else if (horizontalRect) { tw = w >> 1; th = h; }
        else if (asymV || asymH) { tw = smallRect ? w : w >> 1; th = smallRect ? h : h >> 1; }
        else { tw = (w <= 8) ? w : w >> 1; th = (w <= 8) ? h : h >> 1; }
        int encDC[4];
        {   /* sad_x4 against a zero block = the four sub-block sums of the source (motion.cpp:1304-1311) */
            const pixel* f[4] = { c->fenc, c->fenc + deltaX, c->fenc + deltaY * 64, c->fenc + deltaX + deltaY * 64 };
            for (int k = 0; k < 4; k++)
            {
                int sum = 0;
                for (int y = 0; y < th; y++) for (int x = 0; x < tw; x++) sum += f[k][y * 64 + x];
                encDC[k] = sum;
            }
        }
        int plane;
        switch (deltaX)
        {
        case 32: plane = (deltaY % 24 == 0) ? 1 : (deltaY == 8 ? 2 : 0); break;
        case 24: plane = 3; break;
        case 16: plane = (deltaY % 12 == 0) ? 5 : (deltaY == 4 ? 6 : 4); break;
        case 12: plane = 7; break;
        case 8:  plane = (deltaY == 32) ? 8 : 9; break;
        case 4:  plane = (deltaY == 16) ? 10 : 11; break;
        default: plane = 11; break;
        }
        const uint32_t* sumsBase = j->integral[plane] + j->offset;
        if ((w == 64 && h == 64) || (w == 32 && h == 32) || (w == 16 && h == 16) || verticalRect || asymV)
            deltaY *= (int)c->stride;
        if (verticalRect) encDC[1] = encDC[2];
        if (horizontalRect) deltaY = deltaX;
        /* m_cost_mvx / m_cost_mvy are already centred on the MVP (bitcost.h:42: m_cost - mvp); the SEA branch subtracts qmvp AGAIN
         * (motion.cpp:1250-1251), so its p_cost tables are centred on 2 * mvp -- restated as written.  The ADS column costs come from
         * m_fpelMvCosts[-qmvp.x & 3] + (-qmvp.x >> 2) (motion.cpp:1264; bitcost.cpp:68-75: [j][i] = cost[4 i + j]), centred on mvp. */
        const uint16_t* p_cost_mvx = j->mvcost - 2 * c->mvp.x;
        const uint16_t* p_cost_mvy = j->mvcost - 2 * c->mvp.y;
        uint16_t* costX = (uint16_t*)malloc((size_t)(meRangeWidth + 4) * sizeof(uint16_t));
        for (int i = 0; i < meRangeWidth + 4; i++) costX[i] = j->mvcost[((minX + i) << 2) - c->mvp.x];
        for (int ty = minY; ty <= maxY; ty++)
        {
            const int ycost = p_cost_mvy[ty] << 2;
            if (bcost <= ycost) continue;
            bcost -= ycost;
            const int xn = orc_ads(encDC, sumsBase + minX + (intptr_t)ty * c->stride, deltaY, costX, scratch, meRangeWidth, bcost, w, h);
            int i;
            for (i = 0; i < xn - 2; i += 3)
                for (int k = 0; k < 3; k++)
                {   /* COST_MV_X3_ABS: SAD + the x cost only, against the row-relative best */
                    const int tx = minX + scratch[i + k];
                    const int cost = fpel_sad(c, tx, ty) + p_cost_mvx[tx << 2];
                    if (cost < bcost) { bcost = cost; bmv.x = tx; bmv.y = ty; }
                }
            bcost += ycost;
            for (; i < xn; i++)
            {
                const int tx = minX + scratch[i];
                const int cost = cost_fpel(c, tx, ty);
                if (cost < bcost) { bcost = cost; bmv.x = tx; bmv.y = ty; }
            }
        }
        free(costX); free(scratch);
        break;
    }
    case 5: /* X265_FULL_SEARCH motion.cpp:1397-1440: every full-pel position of [mvmin, mvmax], raster order, strict '<' */
    {
        for (int ty = c->mvmin.y; ty <= c->mvmax.y; ty++)
            for (int tx = c->mvmin.x; tx <= c->mvmax.x; tx++)
            {
                int cost = cost_fpel(c, tx, ty);
                if (cost < bcost) { bcost = cost; bmv.x = tx; bmv.y = ty; }
            }
        break;
    }
    case 3: /* X265_STAR_SEARCH motion.cpp:1132-1240 */
    {
        star_t s; s.bmv = bmv; s.bcost = bcost; s.point = 0; s.dist = 0;
        star_pattern(c, &s, 3, merange);
        int done = 0;
        if (s.dist == 1)
        {
            if (s.point)
            {
                int saved = s.bcost;
                mv_t m1 = { s.bmv.x + k_offsets[(s.point - 1) * 2].x, s.bmv.y + k_offsets[(s.point - 1) * 2].y };
                mv_t m2 = { s.bmv.x + k_offsets[(s.point - 1) * 2 + 1].x, s.bmv.y + k_offsets[(s.point - 1) * 2 + 1].y };
                if (in_range(c, m1.x, m1.y)) { int cost = cost_fpel(c, m1.x, m1.y); if (cost < s.bcost) { s.bcost = cost; s.bmv = m1; } }
                if (in_range(c, m2.x, m2.y)) { int cost = cost_fpel(c, m2.x, m2.y); if (cost < s.bcost) { s.bcost = cost; s.bmv = m2; } }
                if (s.bcost == saved) done = 1;
            }
            else
                done = 1;
        }
        if (!done)
        {
            const int RD = 5;
            if (s.dist > RD)
            {
                /* raster refinement :1171-1201; NB the 4th point's mvcost uses (tmv << 3) as written */
                for (int ty = c->mvmin.y; ty <= c->mvmax.y; ty += RD)
                    for (int tx = c->mvmin.x; tx <= c->mvmax.x; tx += RD)
                    {
                        if (tx + RD * 3 <= c->mvmax.x)
                        {
                            for (int k = 0; k < 4; k++)
                            {
                                int sad = fpel_sad(c, tx, ty);
                                g_orc_cnt[2]++;
                                int cost = sad + (k < 3 ? mvcost(c, tx * 4, ty * 4) : mvcost(c, tx * 8, ty * 8));
                                if (cost < s.bcost) { s.bcost = cost; s.bmv.x = tx; s.bmv.y = ty; }
                                if (k < 3) tx += RD;
                            }
                        }
                        else
                        {
                            int cost = cost_fpel(c, tx, ty);
                            if (cost < s.bcost) { s.bcost = cost; s.bmv.x = tx; s.bmv.y = ty; }
                        }
                    }
            }
            while (s.dist > 0)
            {
                s.dist = 0; s.point = 0;
                star_pattern(c, &s, 32, merange);
                if (s.dist == 1)
                {
                    if (!s.point) break;
                    mv_t m1 = { s.bmv.x + k_offsets[(s.point - 1) * 2].x, s.bmv.y + k_offsets[(s.point - 1) * 2].y };
                    mv_t m2 = { s.bmv.x + k_offsets[(s.point - 1) * 2 + 1].x, s.bmv.y + k_offsets[(s.point - 1) * 2 + 1].y };
                    if (in_range(c, m1.x, m1.y)) { int cost = cost_fpel(c, m1.x, m1.y); if (cost < s.bcost) { s.bcost = cost; s.bmv = m1; } }
                    if (in_range(c, m2.x, m2.y)) { int cost = cost_fpel(c, m2.x, m2.y); if (cost < s.bcost) { s.bcost = cost; s.bmv = m2; } }
                    break;
                }
            }
        }
        bmv = s.bmv; bcost = s.bcost;
        break;
    }
    default:
        return -1;
    }

    /* motion.cpp:1440-1447 */
    int bx, by;
    if (bprecost < bcost) { bx = bestprex; by = bestprey; bcost = bprecost; }
    else { bx = bmv.x * 4; by = bmv.y * 4; }

    const int sub = j->subme;
    if (!bcost)
        bcost = mvcost(c, bx, by);
    else if (j->lowres)
    {
        /* :1462-1493 */
        int bdir = 0;
        for (int i = 1; i <= k_workload[sub].hpel_dirs; i++)
        {
            int qx = bx + k_square1[i].x * 2, qy = by + k_square1[i].y * 2;
            if ((qy < qminy) | (qy > qmaxy)) continue;
            int cost = lowres_qpel_cost(c, qx, qy, orc_sad) + mvcost(c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bx += k_square1[bdir].x * 2; by += k_square1[bdir].y * 2;
        bcost = lowres_qpel_cost(c, bx, by, orc_satd) + mvcost(c, bx, by);
        bdir = 0;
        for (int i = 1; i <= k_workload[sub].qpel_dirs; i++)
        {
            int qx = bx + k_square1[i].x, qy = by + k_square1[i].y;
            if ((qy < qminy) | (qy > qmaxy)) continue;
            int cost = lowres_qpel_cost(c, qx, qy, orc_satd) + mvcost(c, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bx += k_square1[bdir].x; by += k_square1[bdir].y;
    }
    else
    {
        /* :1495-1558 */
        cmp_fn hcmp = orc_sad;
        if (k_workload[sub].hpel_satd)
        {
            bcost = subpel_compare(c, bx, by, orc_satd) + mvcost(c, bx, by);
            hcmp = orc_satd;
        }
        for (int it = 0; it < k_workload[sub].hpel_iters; it++)
        {
            int bdir = 0;
            for (int i = 1; i <= k_workload[sub].hpel_dirs; i++)
            {
                int qx = bx + k_square1[i].x * 2, qy = by + k_square1[i].y * 2;
                if ((qy < qminy) | (qy > qmaxy)) continue;
                int cost = subpel_compare(c, qx, qy, hcmp) + mvcost(c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bx += k_square1[bdir].x * 2; by += k_square1[bdir].y * 2; }
            else break;
        }
        if (!k_workload[sub].hpel_satd)
            bcost = subpel_compare(c, bx, by, orc_satd) + mvcost(c, bx, by);
        for (int it = 0; it < k_workload[sub].qpel_iters; it++)
        {
            int bdir = 0;
            for (int i = 1; i <= k_workload[sub].qpel_dirs; i++)
            {
                int qx = bx + k_square1[i].x, qy = by + k_square1[i].y;
                if ((qy < qminy) | (qy > qmaxy)) continue;
                int cost = subpel_compare(c, qx, qy, orc_satd) + mvcost(c, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bx += k_square1[bdir].x; by += k_square1[bdir].y; }
            else break;
        }
    }
    outQMv[0] = bx; outQMv[1] = by;
    return bcost;
#undef CLIPQ
}

/* FrameFilter::computeMEIntegral (encoder/framefilter.cpp:725-822) for a whole picture: per plane (W x H), row y + 1 = the running
 * W-wide row sums of picture row y added to row y (integral_inith), and once H rows exist the row H above is turned into box sums
 * (integral_initv).  Rows run from -padY to picHeight + padY - 2, columns from -padX over `stride` elements, padX = 96, padY = 80. */
void orc_build_integral(const pixel* picOrg, intptr_t stride, int picHeightCtu, uint32_t* const* planes)
{
    static const int W[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 };
    static const int H[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
    const int padX = 96, padY = 80, maxHeight = picHeightCtu * 64;
    for (int k = 0; k < 12; k++) memset(planes[k] - padY * stride - padX, 0, (size_t)stride * sizeof(uint32_t));
    for (int y = -padY; y < maxHeight + padY - 1; y++)
    {
        const pixel* pix = picOrg + (intptr_t)y * stride - padX;
        for (int k = 0; k < 12; k++)
        {
            uint32_t* sum = planes[k] + (intptr_t)(y + 1) * stride - padX;
            orc_integral_inith(sum, pix, stride, W[k]);
            if (y >= H[k] - padY) orc_integral_initv(sum - (intptr_t)H[k] * stride, stride, H[k]);
        }
    }
}

/* oracle/oracle_pixel.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates /root/reference/source/common/pixel.cpp.  Block sizes are runtime arguments and the
 * Hadamard transforms are written as plain int32 butterflies (the reference packs two 16-bit
 * lanes per word, pixel.cpp:188-208; the packed form is exact, so the values are identical --
 * pinned by tests/test_oracle_vs_ref.py incl. the all-max fixtures).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define PIXEL_MAX ((1 << ORC_DEPTH) - 1)
static inline int clip_pixel(int v) { return v < 0 ? 0 : (v > PIXEL_MAX ? PIXEL_MAX : v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

int orc_depth(void) { return ORC_DEPTH; }

/* pixel.cpp:40-55 */
int orc_sad(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int w, int h)
{
    int acc = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
            acc += iabs((int)a[x] - (int)b[x]);
    return acc;
}

/* pixel.cpp:74-119: fenc stride is FENC_STRIDE = 64 (common.h:70) */
void orc_sad_x3(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t rs, int32_t* res, int w, int h)
{
    res[0] = orc_sad(fenc, 64, r0, rs, w, h);
    res[1] = orc_sad(fenc, 64, r1, rs, w, h);
    res[2] = orc_sad(fenc, 64, r2, rs, w, h);
}

void orc_sad_x4(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t rs, int32_t* res, int w, int h)
{
    res[0] = orc_sad(fenc, 64, r0, rs, w, h);
    res[1] = orc_sad(fenc, 64, r1, rs, w, h);
    res[2] = orc_sad(fenc, 64, r2, rs, w, h);
    res[3] = orc_sad(fenc, 64, r3, rs, w, h);
}

/* pixel.cpp:121-165 + the per-size x1/x2/x4 choice of setupPixelPrimitives_c (pixel.cpp:1108-1132).
 * Number of DC terms: 4 when the PU is "4-quadrant" coded, 2 for the 2:1 shapes, else 1. */
static int ads_terms(int w, int h)
{
    if (w == h) return (w >= 16) ? 4 : 1;                    /* 16,32,64 -> x4 ; 4,8 -> x1 */
    if (w == 2 * h || h == 2 * w) return 2;                  /* 8x4 .. 64x32 -> x2 */
    if (w <= 16 && h <= 16) return 1;                        /* 16x12 12x16 16x4 4x16 -> x1 */
    return 4;                                                /* 32x24 24x32 32x8 8x32 64x48 48x64 64x16 16x64 */
}

int orc_ads(const int* encDC, const uint32_t* sums, int delta, const uint16_t* costMvX, int16_t* mvs, int width, int thresh, int w, int h)
{
    int terms = ads_terms(w, h), n = 0;
    int half = w >> 1;
    for (int i = 0; i < width; i++, sums++)
    {
        long v;
        if (terms == 4)
            v = labs((long)encDC[0] - (long)sums[0]) + labs((long)encDC[1] - (long)sums[half])
              + labs((long)encDC[2] - (long)sums[delta]) + labs((long)encDC[3] - (long)sums[delta + half]);
        else if (terms == 2)
            v = labs((long)encDC[0] - (long)sums[0]) + labs((long)encDC[1] - (long)sums[delta]);
        else
            v = labs((long)encDC[0] - (long)sums[0]);
        int ads = (int)v + costMvX[i];
        if (ads < thresh)
            mvs[n++] = (int16_t)i;
    }
    return n;
}

/* 4-point Hadamard on 4 ints, in place */
static inline void had4(int* v0, int* v1, int* v2, int* v3)
{
    int s01 = *v0 + *v1, d01 = *v0 - *v1, s23 = *v2 + *v3, d23 = *v2 - *v3;
    *v0 = s01 + s23; *v2 = s01 - s23; *v1 = d01 + d23; *v3 = d01 - d23;
}

/* sum |H4 D H4^T| over one 4x4 tile, un-normalised (pixel.cpp:210-236 before the >>1) */
static int had4x4_abs(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int m[4][4];
    for (int y = 0; y < 4; y++)
    {
        for (int x = 0; x < 4; x++)
            m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
        had4(&m[y][0], &m[y][1], &m[y][2], &m[y][3]);
    }
    int acc = 0;
    for (int x = 0; x < 4; x++)
    {
        had4(&m[0][x], &m[1][x], &m[2][x], &m[3][x]);
        acc += iabs(m[0][x]) + iabs(m[1][x]) + iabs(m[2][x]) + iabs(m[3][x]);
    }
    return acc;
}

/* pixel.cpp:263-297 + setup table :1134-1158: PUs whose width is a multiple of 8 are tiled 8x4
 * (one >>1 per 8x4 tile, :239-261), the others (w = 4 or 12) are tiled 4x4 (one >>1 per tile). */
int orc_satd(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int w, int h)
{
    int total = 0;
    if ((w & 7) == 0)
    {
        for (int y = 0; y < h; y += 4)
            for (int x = 0; x < w; x += 8)
                total += (had4x4_abs(a + y * sa + x, sa, b + y * sb + x, sb)
                        + had4x4_abs(a + y * sa + x + 4, sa, b + y * sb + x + 4, sb)) >> 1;
    }
    else
    {
        for (int y = 0; y < h; y += 4)
            for (int x = 0; x < w; x += 4)
                total += had4x4_abs(a + y * sa + x, sa, b + y * sb + x, sb) >> 1;
    }
    return total;
}

/* un-normalised 8x8 Hadamard abs-sum (pixel.cpp:299-334) */
static int had8x8_abs(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int m[8][8];
    for (int y = 0; y < 8; y++)
    {
        for (int x = 0; x < 8; x++)
            m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
        /* 8-pt Hadamard = two 4-pt + one combining stage */
        had4(&m[y][0], &m[y][1], &m[y][2], &m[y][3]);
        had4(&m[y][4], &m[y][5], &m[y][6], &m[y][7]);
        for (int k = 0; k < 4; k++)
        {
            int p = m[y][k], q = m[y][k + 4];
            m[y][k] = p + q; m[y][k + 4] = p - q;
        }
    }
    int acc = 0;
    for (int x = 0; x < 8; x++)
    {
        had4(&m[0][x], &m[1][x], &m[2][x], &m[3][x]);
        had4(&m[4][x], &m[5][x], &m[6][x], &m[7][x]);
        for (int k = 0; k < 4; k++)
            acc += iabs(m[k][x] + m[k + 4][x]) + iabs(m[k][x] - m[k + 4][x]);
    }
    return acc;
}

/* pixel.cpp:336-377 + table :1166-1170 (and the chroma/alias variants): 4x4 -> satd_4x4;
 * multiples of 16 -> 16x16 tiles rounded once; otherwise 8x8 tiles each rounded. */
int orc_sa8d(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int w, int h)
{
    if (w < 8 || h < 8)
        return orc_satd(a, sa, b, sb, w, h);
    int total = 0;
    if (((w | h) & 15) == 0)
    {
        for (int y = 0; y < h; y += 16)
            for (int x = 0; x < w; x += 16)
            {
                const pixel* pa = a + y * sa + x; const pixel* pb = b + y * sb + x;
                int s = had8x8_abs(pa, sa, pb, sb) + had8x8_abs(pa + 8, sa, pb + 8, sb)
                      + had8x8_abs(pa + 8 * sa, sa, pb + 8 * sb, sb) + had8x8_abs(pa + 8 * sa + 8, sa, pb + 8 * sb + 8, sb);
                total += (s + 2) >> 2;
            }
    }
    else
    {
        for (int y = 0; y < h; y += 8)
            for (int x = 0; x < w; x += 8)
                total += (had8x8_abs(a + y * sa + x, sa, b + y * sb + x, sb) + 2) >> 2;
    }
    return total;
}

/* pixel.cpp:167-186 */
sse_t orc_sse_pp(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int w, int h)
{
    sse_t acc = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
        {
            int d = (int)a[x] - (int)b[x];
            acc += (sse_t)(d * d);
        }
    return acc;
}

sse_t orc_sse_ss(const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb, int w, int h)
{
    sse_t acc = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
        {
            int d = (int)a[x] - (int)b[x];
            acc += (sse_t)(d * d);   /* int product converted to sse_t as in the reference */
        }
    return acc;
}

/* pixel.cpp:379-391 */
sse_t orc_ssd_s(const int16_t* a, intptr_t sa, int n)
{
    sse_t acc = 0;
    for (int y = 0; y < n; y++, a += sa)
        for (int x = 0; x < n; x++)
            acc += (sse_t)((int)a[x] * (int)a[x]);
    return acc;
}

/* pixel.cpp:703-720: both accumulators are uint32 */
uint64_t orc_var(const pixel* a, intptr_t sa, int n)
{
    uint32_t s = 0, q = 0;
    for (int y = 0; y < n; y++, a += sa)
        for (int x = 0; x < n; x++)
        {
            s += a[x];
            q += (uint32_t)a[x] * a[x];
        }
    return (uint64_t)s + ((uint64_t)q << 32);
}

/* pixel.cpp:726-757: AC energy = sa8d(block, 0) - (sad(block, 0) >> 2) per 8x8 (4x4: satd) */
int orc_psy_cost_pp(const pixel* src, intptr_t ss, const pixel* rec, intptr_t rs, int n)
{
    static const pixel zeros[8] = { 0 };
    if (n == 4)
    {
        int es = (had4x4_abs(src, ss, zeros, 0) >> 1) - (orc_sad(src, ss, zeros, 0, 4, 4) >> 2);
        int er = (had4x4_abs(rec, rs, zeros, 0) >> 1) - (orc_sad(rec, rs, zeros, 0, 4, 4) >> 2);
        return iabs(es - er);
    }
    uint32_t tot = 0;
    for (int y = 0; y < n; y += 8)
        for (int x = 0; x < n; x += 8)
        {
            const pixel* ps = src + y * ss + x; const pixel* pr = rec + y * rs + x;
            int es = ((had8x8_abs(ps, ss, zeros, 0) + 2) >> 2) - (orc_sad(ps, ss, zeros, 0, 8, 8) >> 2);
            int er = ((had8x8_abs(pr, rs, zeros, 0) + 2) >> 2) - (orc_sad(pr, rs, zeros, 0, 8, 8) >> 2);
            tot += (uint32_t)iabs(es - er);
        }
    return (int)tot;
}

/* ---- block ops ---- */
void orc_copy_pp(pixel* d, intptr_t ds, const pixel* s, intptr_t ss, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, s += ss) for (int x = 0; x < w; x++) d[x] = s[x]; }
void orc_copy_ss(int16_t* d, intptr_t ds, const int16_t* s, intptr_t ss, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, s += ss) for (int x = 0; x < w; x++) d[x] = s[x]; }
void orc_copy_sp(pixel* d, intptr_t ds, const int16_t* s, intptr_t ss, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, s += ss) for (int x = 0; x < w; x++) d[x] = (pixel)s[x]; }
void orc_copy_ps(int16_t* d, intptr_t ds, const pixel* s, intptr_t ss, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, s += ss) for (int x = 0; x < w; x++) d[x] = (int16_t)s[x]; }

/* pixel.cpp:814-840 */
void orc_sub_ps(int16_t* d, intptr_t ds, const pixel* a, const pixel* b, intptr_t sa, intptr_t sb, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, a += sa, b += sb) for (int x = 0; x < w; x++) d[x] = (int16_t)((int)a[x] - (int)b[x]); }
void orc_add_ps(pixel* d, intptr_t ds, const pixel* a, const int16_t* r, intptr_t sa, intptr_t sr, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, a += sa, r += sr) for (int x = 0; x < w; x++) d[x] = (pixel)clip_pixel((int)a[x] + (int)r[x]); }
/* pixel.cpp:471-483: one stride for all three buffers */
void orc_calcresidual(const pixel* fenc, const pixel* pred, int16_t* resi, intptr_t stride, int n)
{ for (int y = 0; y < n; y++, fenc += stride, pred += stride, resi += stride) for (int x = 0; x < n; x++) resi[x] = (int16_t)((int)fenc[x] - (int)pred[x]); }
/* pixel.cpp:545-557 */
void orc_pixelavg_pp(pixel* d, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int w, int h)
{ for (int y = 0; y < h; y++, d += ds, a += sa, b += sb) for (int x = 0; x < w; x++) d[x] = (pixel)(((int)a[x] + (int)b[x] + 1) >> 1); }
/* pixel.cpp:842-862 */
void orc_addAvg(const int16_t* a, const int16_t* b, pixel* d, intptr_t sa, intptr_t sb, intptr_t ds, int w, int h)
{
    const int shift = 14 + 1 - ORC_DEPTH;
    const int offset = (1 << (shift - 1)) + 2 * 8192;
    for (int y = 0; y < h; y++, a += sa, b += sb, d += ds)
        for (int x = 0; x < w; x++)
            d[x] = (pixel)clip_pixel(((int)a[x] + (int)b[x] + offset) >> shift);
}
/* pixel.cpp:485-491 */
void orc_transpose(pixel* d, const pixel* s, intptr_t stride, int n)
{ for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) d[r * n + c] = s[c * stride + r]; }
void orc_blockfill_s(int16_t* d, intptr_t ds, int16_t val, int n)
{ for (int y = 0; y < n; y++) for (int x = 0; x < n; x++) d[y * ds + x] = val; }
/* pixel.cpp:401-469: the shr forms add an int16 rounding term before the arithmetic shift */
void orc_cpy2Dto1D_shl(int16_t* d, const int16_t* s, intptr_t ss, int shift, int n)
{ for (int y = 0; y < n; y++, s += ss, d += n) for (int x = 0; x < n; x++) d[x] = (int16_t)(s[x] << shift); }
void orc_cpy2Dto1D_shr(int16_t* d, const int16_t* s, intptr_t ss, int shift, int n)
{ int16_t rnd = (int16_t)(1 << (shift - 1)); for (int y = 0; y < n; y++, s += ss, d += n) for (int x = 0; x < n; x++) d[x] = (int16_t)((s[x] + rnd) >> shift); }
void orc_cpy1Dto2D_shl(int16_t* d, const int16_t* s, intptr_t ds, int shift, int n)
{ for (int y = 0; y < n; y++, s += n, d += ds) for (int x = 0; x < n; x++) d[x] = (int16_t)(s[x] << shift); }
void orc_cpy1Dto2D_shr(int16_t* d, const int16_t* s, intptr_t ds, int shift, int n)
{ int16_t rnd = (int16_t)(1 << (shift - 1)); for (int y = 0; y < n; y++, s += n, d += ds) for (int x = 0; x < n; x++) d[x] = (int16_t)((s[x] + rnd) >> shift); }
/* dct.cpp:728-742 */
uint32_t orc_copy_cnt(int16_t* coeff, const int16_t* resi, intptr_t rs, int n)
{
    uint32_t nz = 0;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
        {
            int16_t v = resi[y * rs + x];
            coeff[y * n + x] = v;
            nz += (v != 0);
        }
    return nz;
}
/* dct.cpp:714-726 */
int orc_count_nonzero(const int16_t* q, int n)
{ int c = 0; for (int i = 0; i < n * n; i++) c += (q[i] != 0); return c; }

/* pixel.cpp:584-602 */
void orc_scale2D_64to32(pixel* d, const pixel* s, intptr_t stride)
{
    for (int y = 0; y < 32; y++)
        for (int x = 0; x < 32; x++)
        {
            const pixel* p = s + 2 * y * stride + 2 * x;
            d[y * 32 + x] = (pixel)((p[0] + p[1] + p[stride] + p[stride + 1] + 2) >> 2);
        }
}
/* pixel.cpp:559-582: two rows of 128 -> two rows of 64 */
void orc_scale1D_128to64(pixel* d, const pixel* s)
{
    for (int x = 0; x < 64; x++)
    {
        d[x]      = (pixel)((s[2 * x] + s[2 * x + 1] + 1) >> 1);
        d[64 + x] = (pixel)((s[128 + 2 * x] + s[128 + 2 * x + 1] + 1) >> 1);
    }
}
/* pixel.cpp:518-543 */
void orc_weight_pp(const pixel* s, pixel* d, intptr_t stride, int width, int height, int w0, int round, int shift, int offset)
{
    const int corr = 14 - ORC_DEPTH;
    for (int y = 0; y < height; y++, s += stride, d += stride)
        for (int x = 0; x < width; x++)
        {
            int16_t v = (int16_t)(s[x] << corr);
            d[x] = (pixel)clip_pixel(((w0 * v + round) >> shift) + offset);
        }
}
/* pixel.cpp:493-516 */
void orc_weight_sp(const int16_t* s, pixel* d, intptr_t ss, intptr_t ds, int width, int height, int w0, int round, int shift, int offset)
{
    for (int y = 0; y < height; y++, s += ss, d += ds)
        for (int x = 0; x < width; x++)
            d[x] = (pixel)clip_pixel(((w0 * (s[x] + 8192) + round) >> shift) + offset);
}

/* pixel.cpp:604-628: nested rounding averages, reads one row/col beyond 2*width x 2*height */
static inline int avg2(int a, int b) { return (a + b + 1) >> 1; }
void orc_frame_init_lowres(const pixel* src0, pixel* dst0, pixel* dsth, pixel* dstv, pixel* dstc,
                           intptr_t src_stride, intptr_t dst_stride, int width, int height)
{
    for (int y = 0; y < height; y++)
    {
        const pixel* r0 = src0 + (intptr_t)2 * y * src_stride;
        const pixel* r1 = r0 + src_stride;
        const pixel* r2 = r1 + src_stride;
        for (int x = 0; x < width; x++)
        {
            int c0 = 2 * x, c1 = 2 * x + 1, c2 = 2 * x + 2;
            dst0[y * dst_stride + x] = (pixel)avg2(avg2(r0[c0], r1[c0]), avg2(r0[c1], r1[c1]));
            dsth[y * dst_stride + x] = (pixel)avg2(avg2(r0[c1], r1[c1]), avg2(r0[c2], r1[c2]));
            dstv[y * dst_stride + x] = (pixel)avg2(avg2(r1[c0], r2[c0]), avg2(r1[c1], r2[c1]));
            dstc[y * dst_stride + x] = (pixel)avg2(avg2(r1[c1], r2[c1]), avg2(r1[c2], r2[c2]));
        }
    }
}

/* pixel.cpp:1027-1041 + ipfilter.cpp:59-77: replicate edges into the margins */
void orc_extend_pic_border(pixel* pic, intptr_t stride, int width, int height, int marginX, int marginY)
{
    for (int y = 0; y < height; y++)
    {
        pixel* row = pic + y * stride;
        for (int x = 0; x < marginX; x++)
        {
            row[-marginX + x] = row[0];
            row[width + x] = row[width - 1];
        }
    }
    pixel* top = pic - marginX;
    pixel* bot = pic - marginX + (intptr_t)(height - 1) * stride;
    for (int y = 1; y <= marginY; y++)
    {
        memcpy(top - y * stride, top, (size_t)stride * sizeof(pixel));
        memcpy(bot + y * stride, bot, (size_t)stride * sizeof(pixel));
    }
}

/* ---- SEA integral planes (encoder/framefilter.cpp:39-143): integral_init{4,8,12,16,24,32}{h,v}, width n as a run-time argument.
 * inith: running n-wide row sums added to the row above (sum[x - stride]); initv: sum[x] = sum[x + n * stride] - sum[x]. ---- */
void orc_integral_inith(uint32_t* sum, const pixel* pix, intptr_t stride, int n)
{
    int32_t v = 0;
    for (int k = 0; k < n; k++) v += pix[k];
    for (int16_t x = 0; x < stride - n; x++)
    {
        sum[x] = v + sum[x - stride];
        v += pix[x + n] - pix[x];
    }
}

void orc_integral_initv(uint32_t* sum, intptr_t stride, int n)
{
    for (int x = 0; x < stride; x++) sum[x] = sum[x + n * stride] - sum[x];
}

/* oracle/oracle_ipfilter.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates /root/reference/source/common/ipfilter.cpp:40-369 with one generic tap loop.
 * Constants: IF_INTERNAL_PREC 14, IF_FILTER_PREC 6, IF_INTERNAL_OFFS 8192 (common/constants.h:66-70).
 */
#include "oracle.h"

#define PIXEL_MAX ((1 << ORC_DEPTH) - 1)
#define HEADROOM  (14 - ORC_DEPTH)

/* HEVC interpolation taps (spec tables 8-11/8-12; same values as constants.cpp:250-268) */
static const int16_t k_luma[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
    { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int16_t k_chroma[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

static inline const int16_t* taps(int ntaps, int idx) { return ntaps == 8 ? k_luma[idx] : k_chroma[idx]; }
static inline int clip_pixel(int v) { return v < 0 ? 0 : (v > PIXEL_MAX ? PIXEL_MAX : v); }

/* ipfilter.cpp:40-57 */
void orc_p2s(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int w, int h)
{
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
            d[x] = (int16_t)((int16_t)(s[x] << HEADROOM) - (int16_t)8192);
}

/* generic FIR along `step` (1 = horizontal, stride = vertical) on pixel input */
static inline int fir_p(const pixel* p, intptr_t step, const int16_t* c, int n)
{
    int acc = 0;
    for (int k = 0; k < n; k++) acc += (int)p[k * step] * c[k];
    return acc;
}
static inline int fir_s(const int16_t* p, intptr_t step, const int16_t* c, int n)
{
    int acc = 0;
    for (int k = 0; k < n; k++) acc += (int)p[k * step] * c[k];
    return acc;
}

/* ipfilter.cpp:79-118: (sum + 32) >> 6, through int16, clamp */
void orc_interp_hpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int coeffIdx, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    s -= ntaps / 2 - 1;
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
        {
            int16_t v = (int16_t)((fir_p(s + x, 1, c, ntaps) + 32) >> 6);
            d[x] = (pixel)clip_pixel(v);
        }
}

/* ipfilter.cpp:120-162: shift = 6 - headroom, offset = -8192 << shift; rowExt adds ntaps-1 rows */
void orc_interp_hps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int coeffIdx, int isRowExt, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    const int shift = 6 - HEADROOM;
    const int offset = (int)((unsigned)-8192 << shift);
    int rows = h;
    s -= ntaps / 2 - 1;
    if (isRowExt)
    {
        s -= (ntaps / 2 - 1) * ss;
        rows += ntaps - 1;
    }
    for (int y = 0; y < rows; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
            d[x] = (int16_t)((fir_p(s + x, 1, c, ntaps) + offset) >> shift);
}

/* ipfilter.cpp:164-203 */
void orc_interp_vpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int coeffIdx, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    s -= (ntaps / 2 - 1) * ss;
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
        {
            int16_t v = (int16_t)((fir_p(s + x, ss, c, ntaps) + 32) >> 6);
            d[x] = (pixel)clip_pixel(v);
        }
}

/* ipfilter.cpp:205-239 */
void orc_interp_vps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int coeffIdx, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    const int shift = 6 - HEADROOM;
    const int offset = (int)((unsigned)-8192 << shift);
    s -= (ntaps / 2 - 1) * ss;
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
            d[x] = (int16_t)((fir_p(s + x, ss, c, ntaps) + offset) >> shift);
}

/* ipfilter.cpp:241-282: shift = 6 + headroom, offset = half + (8192 << 6) */
void orc_interp_vsp(const int16_t* s, intptr_t ss, pixel* d, intptr_t ds, int coeffIdx, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    const int shift = 6 + HEADROOM;
    const int offset = (1 << (shift - 1)) + (8192 << 6);
    s -= (ntaps / 2 - 1) * ss;
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
        {
            int16_t v = (int16_t)((fir_s(s + x, ss, c, ntaps) + offset) >> shift);
            d[x] = (pixel)clip_pixel(v);
        }
}

/* ipfilter.cpp:284-317: >> 6, no rounding, no clamp */
void orc_interp_vss(const int16_t* s, intptr_t ss, int16_t* d, intptr_t ds, int coeffIdx, int ntaps, int w, int h)
{
    const int16_t* c = taps(ntaps, coeffIdx);
    s -= (ntaps / 2 - 1) * ss;
    for (int y = 0; y < h; y++, s += ss, d += ds)
        for (int x = 0; x < w; x++)
            d[x] = (int16_t)(fir_s(s + x, ss, c, ntaps) >> 6);
}

/* ipfilter.cpp:362-369: hps(rowExt) into a w-stride scratch, then vsp from row ntaps/2-1 */
void orc_interp_hvpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int idxX, int idxY, int ntaps, int w, int h)
{
    int16_t tmp[64 * (64 + 7)];
    orc_interp_hps(s, ss, tmp, w, idxX, 1, ntaps, w, h);
    orc_interp_vsp(tmp + (ntaps / 2 - 1) * w, w, d, ds, idxY, ntaps, w, h);
}

/* oracle/oracle_intra.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Restates /root/reference/source/common/intrapred.cpp:31-234.  Neighbour layout (srcPix):
 * [0] = top-left, [1 .. 2N] = top + top-right, [2N+1 .. 4N] = left + bottom-left.
 * Written per output pixel (closed form) rather than with the reference's flip/transposes.
 */
#include "oracle.h"

#define PIXEL_MAX ((1 << ORC_DEPTH) - 1)
static inline int clip_pixel(int v) { return v < 0 ? 0 : (v > PIXEL_MAX ? PIXEL_MAX : v); }

/* intrapred.cpp:31-51 */
void orc_intra_filter(const pixel* s, pixel* f, int n)
{
    const int n2 = 2 * n, n4 = 4 * n;
    f[0] = (pixel)((2 * s[0] + s[1] + s[n2 + 1] + 2) >> 2);
    for (int i = 1; i < n2; i++)
        f[i] = (pixel)((2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2);
    f[n2] = s[n2];
    f[n2 + 1] = (pixel)((2 * s[n2 + 1] + s[0] + s[n2 + 2] + 2) >> 2);
    for (int i = n2 + 2; i < n4; i++)
        f[i] = (pixel)((2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2);
    f[n4] = s[n4];
}

/* intrapred.cpp:53-85 */
static void pred_dc(pixel* dst, intptr_t ds, const pixel* src, int bFilter, int n)
{
    const pixel* top = src + 1;
    const pixel* left = src + 2 * n + 1;
    int sum = n;
    for (int i = 0; i < n; i++) sum += top[i] + left[i];
    int dc = sum / (2 * n);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
        {
            int v = dc;
            if (bFilter)
            {
                if (x == 0 && y == 0) v = (top[0] + left[0] + 2 * dc + 2) >> 2;
                else if (y == 0)      v = (top[x] + 3 * dc + 2) >> 2;
                else if (x == 0)      v = (left[y] + 3 * dc + 2) >> 2;
            }
            dst[y * ds + x] = (pixel)v;
        }
}

/* intrapred.cpp:87-100 */
static void pred_planar(pixel* dst, intptr_t ds, const pixel* src, int n)
{
    const pixel* top = src + 1;
    const pixel* left = src + 2 * n + 1;
    int lg = 0; while ((1 << lg) < n) lg++;
    int tr = top[n], bl = left[n];
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            dst[y * ds + x] = (pixel)(((n - 1 - x) * left[y] + (n - 1 - y) * top[x] + (x + 1) * tr + (y + 1) * bl + n) >> (lg + 1));
}

static const int8_t  k_angle[17]   = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
static const int16_t k_invAngle[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };

/* Angular prediction in the "vertical family" frame: `main` = neighbours along the prediction
 * axis (index 1..2N), `side` = the other edge; returns pred(row r, col c) where r walks away from
 * the main edge.  Horizontal modes (<18) swap the edges and transpose (intrapred.cpp:102-204). */
static void pred_ang_vfamily(pixel* out /* n*n, row-major in the v-frame */, const pixel* src, int horMode, int angleOffset, int bFilter, int n)
{
    const int n2 = 2 * n;
    pixel nb[129];
    nb[0] = src[0];
    for (int i = 0; i < n2; i++)
    {
        nb[1 + i]      = horMode ? src[n2 + 1 + i] : src[1 + i];
        nb[n2 + 1 + i] = horMode ? src[1 + i]      : src[n2 + 1 + i];
    }
    int angle = k_angle[8 + angleOffset];
    if (angle == 0)
    {
        for (int r = 0; r < n; r++)
            for (int c = 0; c < n; c++)
                out[r * n + c] = nb[1 + c];
        if (bFilter)
            for (int r = 0; r < n; r++)
                out[r * n] = (pixel)clip_pixel((int16_t)(nb[1] + ((nb[n2 + 1 + r] - nb[0]) >> 1)));
        return;
    }
    pixel line[64 + 33];
    const pixel* ref;
    if (angle < 0)
    {
        int nproj = -((n * angle) >> 5) - 1;
        pixel* base = line + nproj + 1;                 /* base[-1] = top-left */
        int inv = k_invAngle[-angleOffset - 1], acc = 128;
        for (int i = 0; i < nproj; i++)
        {
            acc += inv;
            base[-2 - i] = nb[n2 + (acc >> 8)];
        }
        for (int i = 0; i < n + 1; i++) base[-1 + i] = nb[i];
        ref = base;
    }
    else
        ref = nb + 1;
    int pos = 0;
    for (int r = 0; r < n; r++)
    {
        pos += angle;
        int off = pos >> 5, frac = pos & 31;
        for (int c = 0; c < n; c++)
            out[r * n + c] = frac ? (pixel)(((32 - frac) * ref[off + c] + frac * ref[off + c + 1] + 16) >> 5)
                                  : ref[off + c];
    }
}

void orc_intra_pred(pixel* dst, intptr_t ds, const pixel* src, int mode, int bFilter, int n)
{
    if (mode == 0) { pred_planar(dst, ds, src, n); return; }
    if (mode == 1) { pred_dc(dst, ds, src, bFilter, n); return; }
    pixel tmp[32 * 32];
    int hor = mode < 18;
    pred_ang_vfamily(tmp, src, hor, hor ? 10 - mode : mode - 26, bFilter, n);
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
            dst[y * ds + x] = hor ? tmp[x * n + y] : tmp[y * n + x];
}

/* g_intraFilterFlags (constants.cpp:561): which TU sizes use filtered neighbours for a mode.
 * HEVC 8.4.4.2.3: filter when min(|mode-26|,|mode-10|) > threshold(size): 8->7, 16->1, 32->0; never for 4, DC. */
static int use_filtered(int mode, int n)
{
    if (mode == 1 || n == 4) return 0;
    if (mode == 0) return n >= 8;
    int d1 = mode > 26 ? mode - 26 : 26 - mode, d2 = mode > 10 ? mode - 10 : 10 - mode;
    int d = d1 < d2 ? d1 : d2;
    int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    return d > thr;
}

/* intrapred.cpp:206-234: 33 blocks of n*n; horizontal modes are left in the v-frame (untransposed) */
void orc_intra_pred_allangs(pixel* dst, const pixel* refPix, const pixel* filtPix, int bLuma, int n)
{
    for (int mode = 2; mode <= 34; mode++)
    {
        const pixel* src = use_filtered(mode, n) ? filtPix : refPix;
        int hor = mode < 18;
        pred_ang_vfamily(dst + (mode - 2) * n * n, src, hor, hor ? 10 - mode : mode - 26, bLuma, n);
    }
}

/* oracle/frame_spec.h -- TEST INFRASTRUCTURE ONLY.
 * Shared by the oracle (oracle_frame.c) and the reference shim (ref_shim.cpp): the frame-level
 * "CTU analysis" workload definition used by tests and bench.py (DESIGN.md, section "Frame analysis
 * workload").  It restates how x265's callers drive the primitives for one inter frame:
 *   - PU enumeration per 64x64 CTU: 2Nx2N, 2NxN, Nx2N at CU sizes 64/32/16/8 (preset slow: rect
 *     on, AMP off; param.cpp:478-492), PUs not fully inside the picture are skipped
 *     (analysis.cpp compressInterCU_* only visits CUs inside the picture),
 *   - per PU and reference: Search::setSearchRange (search.cpp:2724-2768, via CUData::clipMv
 *     cudata.cpp:1915-1928) then MotionEstimate::motionEstimate (motion.cpp:739),
 *   - MVP / candidates come from a caller-supplied 16x16-granular predictor field (what AMVP /
 *     the lookahead's lowres MVs provide in the encoder: search.cpp:1968-2023, 2406-2410).
 * The product restates the same rules independently in x265_b200/csrc/analyser.cuh; tests compare
 * the two job lists element by element.
 */
#ifndef X265_FRAME_SPEC_H
#define X265_FRAME_SPEC_H
#include <stdint.h>

#define FS_CTU 64
#define FS_MARGIN_X 96   /* picyuv.cpp:87: maxCUSize + 32 */
#define FS_MARGIN_Y 80   /* picyuv.cpp:88: maxCUSize + 16 */

typedef struct {
    int width, height;      /* luma picture size */
    int stride;             /* plane stride in elements (same for fenc and refs) */
    int numRefs;
    int method, subme, merange;
    int rect;               /* evaluate 2NxN / Nx2N */
    int qp;                 /* luma QP for quant/dequant (per = qp/6, rem = qp%6) */
    int amp;                /* evaluate the AMP partitions 2NxnU / 2NxnD / nLx2N / nRx2N of CUs >= 16 (param.cpp:494-520: slower+) */
} fs_params;

typedef struct {
    int32_t offset;         /* element offset of the PU origin from the plane origin pixel (0,0) */
    int16_t ref;
    int8_t pw, ph;
    int16_t mvmin[2], mvmax[2];
    int16_t qmvp[2];
    int16_t mvc[8];
    int8_t numCand, method, subme, merange;
} fs_me_job;               /* layout == x265cu_me_job */

static inline int fs_stride(int width) { return (width + 2 * FS_MARGIN_X + 63) / 64 * 64; }
static inline int fs_field_w(int width) { return (width + 15) / 16; }
static inline int fs_field_h(int height) { return (height + 15) / 16; }

/* PU table of one CTU: (x, y, w, h, cuX, cuY, cuSize); returns count (<= 85 * 5 + 21 * 8 = 593) */
#define FS_MAX_CTU_PUS 593
static inline int fs_ctu_pus(int rect, int amp, int16_t out[][7])
{
    int n = 0;
    for (int size = 64; size >= 8; size >>= 1)
        for (int cy = 0; cy < 64; cy += size)
            for (int cx = 0; cx < 64; cx += size)
            {
                int16_t p[5][4] = { { (int16_t)cx, (int16_t)cy, (int16_t)size, (int16_t)size },
                                    { (int16_t)cx, (int16_t)cy, (int16_t)size, (int16_t)(size / 2) },
                                    { (int16_t)cx, (int16_t)(cy + size / 2), (int16_t)size, (int16_t)(size / 2) },
                                    { (int16_t)cx, (int16_t)cy, (int16_t)(size / 2), (int16_t)size },
                                    { (int16_t)(cx + size / 2), (int16_t)cy, (int16_t)(size / 2), (int16_t)size } };
                int np = rect ? 5 : 1;
                for (int k = 0; k < np; k++)
                {
                    out[n][0] = p[k][0]; out[n][1] = p[k][1]; out[n][2] = p[k][2]; out[n][3] = p[k][3];
                    out[n][4] = (int16_t)cx; out[n][5] = (int16_t)cy; out[n][6] = (int16_t)size;
                    n++;
                }
                if (amp && size >= 16)
                {   /* AMP (not at the minimum CU size): 2NxnU, 2NxnD, nLx2N, nRx2N -- the quarter / three-quarter splits */
                    const int16_t q = (int16_t)(size / 4), t = (int16_t)(size - size / 4), S = (int16_t)size;
                    const int16_t a[8][4] = { { 0, 0, S, q }, { 0, q, S, t }, { 0, 0, S, t }, { 0, t, S, q },
                                              { 0, 0, q, S }, { q, 0, t, S }, { 0, 0, t, S }, { t, 0, q, S } };
                    for (int k = 0; k < 8; k++)
                    {
                        out[n][0] = (int16_t)(cx + a[k][0]); out[n][1] = (int16_t)(cy + a[k][1]); out[n][2] = a[k][2]; out[n][3] = a[k][3];
                        out[n][4] = (int16_t)cx; out[n][5] = (int16_t)cy; out[n][6] = (int16_t)size;
                        n++;
                    }
                }
            }
    return n;
}

static inline int fs_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Enumerate all ME jobs of a frame in canonical order (CTU raster, ref, PU table order).
 * field: int16 [numRefs][fh][fw][2] qpel predictors.  jobs may be NULL to count only. */
static inline int fs_build_me_jobs(const fs_params* p, const int16_t* field, fs_me_job* jobs)
{
    int16_t pus[FS_MAX_CTU_PUS][7];
    const int npu = fs_ctu_pus(p->rect, p->amp, pus);
    const int fw = fs_field_w(p->width), fh = fs_field_h(p->height);
    const int ctuW = (p->width + 63) / 64, ctuH = (p->height + 63) / 64;
    int n = 0;
    for (int cty = 0; cty < ctuH; cty++)
        for (int ctx = 0; ctx < ctuW; ctx++)
            for (int r = 0; r < p->numRefs; r++)
                for (int k = 0; k < npu; k++)
                {
                    int x = ctx * 64 + pus[k][0], y = cty * 64 + pus[k][1], w = pus[k][2], h = pus[k][3];
                    /* the reference never evaluates partitions of a CU that crosses the picture edge (analysis.cpp visits
                     * only CUs inside the picture): filter by the CU, not by the PU */
                    (void)w; (void)h;
                    if (ctx * 64 + pus[k][4] + pus[k][6] > p->width || cty * 64 + pus[k][5] + pus[k][6] > p->height) continue;
                    if (jobs)
                    {
                        fs_me_job* j = &jobs[n];
                        const int cuX = ctx * 64 + pus[k][4], cuY = cty * 64 + pus[k][5];
                        const int bx = x >> 4, by = y >> 4;
                        const int16_t* f = field + ((size_t)(r * fh + by) * fw + bx) * 2;
                        const int16_t* fr = field + ((size_t)(r * fh + by) * fw + fs_clampi(bx + 1, 0, fw - 1)) * 2;
                        const int16_t* fb = field + ((size_t)(r * fh + fs_clampi(by + 1, 0, fh - 1)) * fw + bx) * 2;
                        j->offset = y * p->stride + x;
                        j->ref = (int16_t)r; j->pw = (int8_t)w; j->ph = (int8_t)h;
                        j->qmvp[0] = f[0]; j->qmvp[1] = f[1];
                        j->mvc[0] = fr[0]; j->mvc[1] = fr[1]; j->mvc[2] = fb[0]; j->mvc[3] = fb[1];
                        j->mvc[4] = j->mvc[5] = j->mvc[6] = j->mvc[7] = 0;
                        j->numCand = 2;
                        j->method = (int8_t)p->method; j->subme = (int8_t)p->subme; j->merange = (int8_t)p->merange;
                        /* setSearchRange + clipMv (CU position, offset 8, maxCUSize 64) */
                        int dist = p->merange << 2;
                        int mnx = f[0] - dist, mny = f[1] - dist, mxx = f[0] + dist, mxy = f[1] + dist;
                        int xmax = (p->width + 8 - cuX - 1) << 2, xmin = -((64 + 8 + cuX - 1) << 2);
                        int ymax = (p->height + 8 - cuY - 1) << 2, ymin = -((64 + 8 + cuY - 1) << 2);
                        mnx = fs_clampi(mnx, xmin, xmax); mxx = fs_clampi(mxx, xmin, xmax);
                        mny = fs_clampi(mny, ymin, ymax); mxy = fs_clampi(mxy, ymin, ymax);
                        mnx >>= 2; mny >>= 2; mxx >>= 2; mxy >>= 2;
                        if (mxy < mny) mxy = mny;
                        j->mvmin[0] = (int16_t)mnx; j->mvmin[1] = (int16_t)mny; j->mvmax[0] = (int16_t)mxx; j->mvmax[1] = (int16_t)mxy;
                    }
                    n++;
                }
    return n;
}

/* CU list of a frame for the residual / intra stages: all CUs (64,32,16,8) fully inside the picture,
 * canonical order = CTU raster, then size 64..8, then raster inside the CTU.  out: (x, y, size). */
static inline int fs_build_cus(int width, int height, int16_t (*out)[3])
{
    const int ctuW = (width + 63) / 64, ctuH = (height + 63) / 64;
    int n = 0;
    for (int cty = 0; cty < ctuH; cty++)
        for (int ctx = 0; ctx < ctuW; ctx++)
            for (int size = 64; size >= 8; size >>= 1)
                for (int cy = 0; cy < 64; cy += size)
                    for (int cx = 0; cx < 64; cx += size)
                    {
                        int x = ctx * 64 + cx, y = cty * 64 + cy;
                        if (x + size > width || y + size > height) continue;
                        if (out) { out[n][0] = (int16_t)x; out[n][1] = (int16_t)y; out[n][2] = (int16_t)size; }
                        n++;
                    }
    return n;
}

#endif

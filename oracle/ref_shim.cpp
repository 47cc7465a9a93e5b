/* oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiled by oracle/Makefile.ref TOGETHER WITH the unmodified reference sources under
 * /root/reference/source into oracle/_ref/libx265ref{8,10}.so.  It exposes, through a flat
 * extern "C" surface that Python/ctypes and the C tests can bind:
 *   - the reference's post-alias C primitive table (primitives.cpp:63-73 setupCPrimitives +
 *     :88-209 setupAliasPrimitives) as raw function pointers looked up by name,
 *   - the reference's MotionEstimate::motionEstimate (encoder/motion.cpp:739) on caller planes,
 *   - BitCost tables (encoder/bitcost.cpp:33-88), lambda tables (common/constants.cpp),
 *   - the lookahead helpers used to pin the frame-level restatements.
 * It contains no algorithm of its own: every call forwards to reference code.
 */
#include "common.h"
#include "primitives.h"
#include "constants.h"
#include "lowres.h"
#include "motion.h"
#include "bitcost.h"
#include "mv.h"
#include "yuv.h"
#include "picyuv.h"
#include <string.h>
#include <stdio.h>

using namespace X265_NS;

namespace X265_NS {
void setupCPrimitives(EncoderPrimitives &p);
void setupAliasPrimitives(EncoderPrimitives &p);
}

static EncoderPrimitives g_c;
static bool g_init = false;

static void ensure_init()
{
    if (g_init) return;
    memset(&g_c, 0, sizeof(g_c));
    setupCPrimitives(g_c);
    setupAliasPrimitives(g_c);
    /* the global table is what lowresQPelCost / subpelCompare / MotionEstimate use */
    memcpy(&primitives, &g_c, sizeof(g_c));
    g_init = true;
}

#define PU_FIELD(f)   if (!strcmp(name, "pu." #f))  return (void*)g_c.pu[i].f;
#define PU_FIELD2(f)  if (!strcmp(name, "pu." #f))  return (void*)g_c.pu[i].f[j];
#define CU_FIELD(f)   if (!strcmp(name, "cu." #f))  return (void*)g_c.cu[i].f;
#define CU_FIELD2(f)  if (!strcmp(name, "cu." #f))  return (void*)g_c.cu[i].f[j];
#define TOP_FIELD(f)  if (!strcmp(name, #f))        return (void*)g_c.f;
#define CPU_FIELD(f)  if (!strcmp(name, "chroma.pu." #f)) return (void*)g_c.chroma[k].pu[i].f;
#define CPU_FIELD2(f) if (!strcmp(name, "chroma.pu." #f)) return (void*)g_c.chroma[k].pu[i].f[j];
#define CCU_FIELD(f)  if (!strcmp(name, "chroma.cu." #f)) return (void*)g_c.chroma[k].cu[i].f;
#define CCU_FIELD2(f) if (!strcmp(name, "chroma.cu." #f)) return (void*)g_c.chroma[k].cu[i].f[j];

extern "C" {

int x265ref_depth(void) { return X265_DEPTH; }

/* name: "pu.sad", "cu.dct", "quant", "chroma.pu.filter_hpp" ...; i = LumaPU / LumaCU index,
 * j = second index (alignment / intra mode), k = csp for chroma entries. */
void* x265ref_get(const char* name, int i, int j, int k)
{
    ensure_init();
    PU_FIELD(sad) PU_FIELD(sad_x3) PU_FIELD(sad_x4) PU_FIELD(ads) PU_FIELD(satd)
    PU_FIELD(luma_hpp) PU_FIELD(luma_hps) PU_FIELD(luma_vpp) PU_FIELD(luma_vps)
    PU_FIELD(luma_vsp) PU_FIELD(luma_vss) PU_FIELD(luma_hvpp)
    PU_FIELD2(pixelavg_pp) PU_FIELD2(addAvg) PU_FIELD(copy_pp) PU_FIELD2(convert_p2s)

    CU_FIELD(dct) CU_FIELD(idct) CU_FIELD2(calcresidual) CU_FIELD(sub_ps) CU_FIELD2(add_ps)
    CU_FIELD2(blockfill_s) CU_FIELD(copy_cnt) CU_FIELD(count_nonzero)
    CU_FIELD(cpy2Dto1D_shl) CU_FIELD(cpy2Dto1D_shr) CU_FIELD2(cpy1Dto2D_shl) CU_FIELD(cpy1Dto2D_shr)
    CU_FIELD(copy_sp) CU_FIELD(copy_ps) CU_FIELD(copy_ss) CU_FIELD(copy_pp)
    CU_FIELD(var) CU_FIELD(sse_pp) CU_FIELD(sse_ss) CU_FIELD(psy_cost_pp) CU_FIELD2(ssd_s)
    CU_FIELD(sa8d) CU_FIELD(transpose) CU_FIELD(intra_pred_allangs) CU_FIELD(intra_filter)
    CU_FIELD2(intra_pred)

    TOP_FIELD(dst4x4) TOP_FIELD(idst4x4) TOP_FIELD(quant) TOP_FIELD(nquant)
    TOP_FIELD(dequant_scaling) TOP_FIELD(dequant_normal) TOP_FIELD(denoiseDct)
    TOP_FIELD(scale2D_64to32) TOP_FIELD(frameInitLowres) TOP_FIELD(frameInitLowerRes)
    TOP_FIELD(extendRowBorder) TOP_FIELD(weight_pp) TOP_FIELD(weight_sp) TOP_FIELD(propagateCost)
    if (!strcmp(name, "scale1D_128to64")) return (void*)g_c.scale1D_128to64[j];
    if (!strcmp(name, "integral_initv")) return (void*)g_c.integral_initv[i];
    if (!strcmp(name, "integral_inith")) return (void*)g_c.integral_inith[i];

    CPU_FIELD(satd) CPU_FIELD(filter_vpp) CPU_FIELD(filter_vps) CPU_FIELD(filter_vsp)
    CPU_FIELD(filter_vss) CPU_FIELD(filter_hpp) CPU_FIELD(filter_hps) CPU_FIELD2(addAvg)
    CPU_FIELD(copy_pp) CPU_FIELD2(p2s)
    CCU_FIELD(sa8d) CCU_FIELD(sse_pp) CCU_FIELD(sub_ps) CCU_FIELD2(add_ps)
    CCU_FIELD(copy_ps) CCU_FIELD(copy_sp) CCU_FIELD(copy_ss) CCU_FIELD(copy_pp)
    return NULL;
}

/* constants (common/constants.cpp) so the oracle's regenerated tables can be pinned */
const int16_t* x265ref_dct_matrix(int n)
{
    switch (n) { case 4: return &g_t4[0][0]; case 8: return &g_t8[0][0];
                 case 16: return &g_t16[0][0]; case 32: return &g_t32[0][0]; }
    return NULL;
}
const int16_t* x265ref_luma_filter(void)   { return &g_lumaFilter[0][0]; }
const int16_t* x265ref_chroma_filter(void) { return &g_chromaFilter[0][0]; }
const uint8_t* x265ref_intra_filter_flags(void) { return g_intraFilterFlags; }
double x265ref_lambda(int qp)  { return x265_lambda_tab[qp]; }
double x265ref_lambda2(int qp) { return x265_lambda2_tab[qp]; }
int x265ref_partition_from_sizes(int w, int h) { ensure_init(); return partitionFromSizes(w, h); }

void x265ref_extend_pic_border(pixel* pic, intptr_t stride, int width, int height, int marginX, int marginY)
{
    ensure_init();
    extendPicBorder(pic, stride, width, height, marginX, marginY);
}

/* BitCost (encoder/bitcost.cpp:33-88): mvcost table for one qp, centred: out[i + range] = cost[i]
 * for i in [-range, range]. */
struct BitCostPeek : public BitCost { uint16_t* table() { return m_cost; } };
void x265ref_mvcost_table(int qp, int range, uint16_t* out)
{
    BitCostPeek bc;
    bc.setQP(qp);
    uint16_t* c = bc.table();
    for (int i = -range; i <= range; i++) out[i + range] = c[i];
}

/* MotionEstimate::motionEstimate (encoder/motion.cpp:739-1569) on a caller-supplied full-res plane
 * pair, driven through the luma-only setSourcePU (motion.cpp:167-192).  All MVs as int32 pairs.
 * lowres != 0: `planes` are the 4 lowres hpel planes (lowres.h:67-120 path).
 * Returns the cost, writes outQMv[2]. */
int x265ref_motion_estimate(pixel* fencPlane, intptr_t fencStride, intptr_t offset,
                            pixel* const* refPlanes, intptr_t refStride, int lowres,
                            int pw, int ph, int method, int subme, int qp,
                            const int* mvmin, const int* mvmax, const int* qmvp,
                            int numCand, const int* mvc, int merange, int* outQMv)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setQP(qp);
    me.setSourcePU(fencPlane, fencStride, offset, pw, ph, method, method, method, subme);
    ReferencePlanes ref;
    ref.lumaStride = refStride;
    ref.isLowres = !!lowres;
    if (lowres)
    {
        for (int i = 0; i < 4; i++) ref.lowresPlane[i] = refPlanes[i];
        ref.fpelPlane[0] = refPlanes[0];
    }
    else
        ref.fpelPlane[0] = refPlanes[0];
    MV mn(mvmin[0], mvmin[1]), mx(mvmax[0], mvmax[1]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[32];
    for (int i = 0; i < numCand && i < 32; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mn, mx, mvp, numCand, cands, merange, out, 1, NULL);
    outQMv[0] = out.x; outQMv[1] = out.y;
    return cost;
}

/* The 12 SEA integral planes of a picture exactly as FrameFilter::computeMEIntegral produces them (encoder/framefilter.cpp:725-822),
 * driven row by row through the REAL integral_inith / integral_initv primitives; planes[k] = pixel (0, 0) of plane k in a buffer of
 * the picture plane's geometry (pins oracle/oracle_me.c: orc_build_integral). */
void x265ref_build_integral(const pixel* picOrg, intptr_t stride, int picHeightCtu, uint32_t* const* planes)
{
    ensure_init();
    static const int W[12] = { INTEGRAL_32, INTEGRAL_32, INTEGRAL_32, INTEGRAL_24, INTEGRAL_16, INTEGRAL_16, INTEGRAL_16, INTEGRAL_12, INTEGRAL_8, INTEGRAL_8, INTEGRAL_4, INTEGRAL_4 };
    static const int Hk[12] = { INTEGRAL_32, INTEGRAL_24, INTEGRAL_8, INTEGRAL_32, INTEGRAL_16, INTEGRAL_12, INTEGRAL_4, INTEGRAL_16, INTEGRAL_32, INTEGRAL_8, INTEGRAL_16, INTEGRAL_4 };
    static const int Hn[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
    const int padX = 64 + 32, padY = 64 + 16, maxHeight = picHeightCtu * 64;
    for (int k = 0; k < INTEGRAL_PLANE_NUM; k++) memset(planes[k] - padY * stride - padX, 0, stride * sizeof(uint32_t));
    for (int y = -padY; y < maxHeight + padY - 1; y++)
    {
        pixel* pix = (pixel*)picOrg + y * stride - padX;
        for (int k = 0; k < INTEGRAL_PLANE_NUM; k++)
        {
            uint32_t* sum = planes[k] + (y + 1) * stride - padX;
            primitives.integral_inith[W[k]](sum, pix, stride);
            if (y >= Hn[k] - padY) primitives.integral_initv[Hk[k]](sum - Hn[k] * stride, stride);
        }
    }
}

/* motionEstimate with X265_SEA: MotionEstimate::integral[] set as Search::predInterSearch sets it (search.cpp:2264): plane + PU offset */
int x265ref_motion_estimate_sea(pixel* fencPlane, intptr_t fencStride, intptr_t offset, pixel* refPlane, intptr_t refStride,
                                uint32_t* const* integral, int pw, int ph, int subme, int qp,
                                const int* mvmin, const int* mvmax, const int* qmvp, int numCand, const int* mvc, int merange, int* outQMv)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    MotionEstimate me;
    me.init(X265_CSP_I400);
    me.setQP(qp);
    /* SEA reads source sub-blocks beyond a narrow PU's width (motion.cpp:1304-1311 with deltaX = w for w <= 8): make the PU cache
     * deterministic (zero) outside the PU instead of leaving malloc's content there */
    memset(me.fencPUYuv.m_buf[0], 0, sizeof(pixel) * me.fencPUYuv.m_size * me.fencPUYuv.m_size);
    me.setSourcePU(fencPlane, fencStride, offset, pw, ph, X265_SEA, X265_SEA, X265_SEA, subme);
    for (int k = 0; k < INTEGRAL_PLANE_NUM; k++) me.integral[k] = integral[k] + offset;
    ReferencePlanes ref;
    ref.lumaStride = refStride;
    ref.isLowres = false;
    ref.fpelPlane[0] = refPlane;
    MV mn(mvmin[0], mvmin[1]), mx(mvmax[0], mvmax[1]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[32];
    for (int i = 0; i < numCand && i < 32; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mn, mx, mvp, numCand, cands, merange, out, 1, NULL);
    outQMv[0] = out.x; outQMv[1] = out.y;
    return cost;
}

/* The same call with the chroma-SATD term of subpelCompare active (motion.cpp:204-212, 1601-1661): driven
 * through the Yuv variant of setSourcePU (motion.cpp:194-222) with bChroma = true, 4:2:0.  The PU is handed
 * over as a Yuv whose top-left block is the PU (puPartIdx 0) and the reference as planes already offset to the
 * PU, with a PicYuv whose CTU / partition offset tables are a single zero (so blockOffset = 0, motion.cpp:753,
 * and getCbAddr(0, 0) = fpelPlane[1], lowres.h:62-63).  Chroma origin of the PU: (x >> 1, y >> 1). */
int x265ref_motion_estimate_chroma(pixel* fencPlane, intptr_t fencStride, intptr_t offset,
                                   pixel* fencCb, pixel* fencCr, intptr_t cstride,
                                   pixel* refPlane, intptr_t refStride, pixel* refCb, pixel* refCr,
                                   int pw, int ph, int method, int subme, int qp,
                                   const int* mvmin, const int* mvmax, const int* qmvp,
                                   int numCand, const int* mvc, int merange, int* outQMv)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    const intptr_t py = offset / fencStride, px = offset % fencStride;
    const intptr_t coff = (py >> 1) * cstride + (px >> 1);
    Yuv src;
    if (!src.create(64, X265_CSP_I420)) return -1;
    for (int y = 0; y < ph; y++)
        memcpy(src.m_buf[0] + y * src.m_size, fencPlane + offset + y * fencStride, pw * sizeof(pixel));
    for (int y = 0; y < ph / 2; y++)
    {
        memcpy(src.m_buf[1] + y * src.m_csize, fencCb + coff + y * cstride, (pw / 2) * sizeof(pixel));
        memcpy(src.m_buf[2] + y * src.m_csize, fencCr + coff + y * cstride, (pw / 2) * sizeof(pixel));
    }
    MotionEstimate me;
    me.init(X265_CSP_I420);
    me.setQP(qp);
    me.setSourcePU(src, 0, 0, 0, pw, ph, method, subme, true);
    PicYuv pic;
    intptr_t zero = 0;
    pic.m_cuOffsetY = &zero; pic.m_cuOffsetC = &zero; pic.m_buOffsetY = &zero; pic.m_buOffsetC = &zero;
    pic.m_picOrg[0] = refPlane + offset; pic.m_picOrg[1] = refCb + coff; pic.m_picOrg[2] = refCr + coff;
    pic.m_stride = refStride; pic.m_strideC = cstride;
    ReferencePlanes ref;
    ref.reconPic = &pic;
    ref.lumaStride = refStride;
    ref.isLowres = false;
    ref.fpelPlane[0] = refPlane + offset; ref.fpelPlane[1] = refCb + coff; ref.fpelPlane[2] = refCr + coff;
    MV mn(mvmin[0], mvmin[1]), mx(mvmax[0], mvmax[1]), mvp(qmvp[0], qmvp[1]), out;
    MV cands[32];
    for (int i = 0; i < numCand && i < 32; i++) cands[i] = MV(mvc[2 * i], mvc[2 * i + 1]);
    int cost = me.motionEstimate(&ref, mn, mx, mvp, numCand, cands, merange, out, 1, NULL);
    outQMv[0] = out.x; outQMv[1] = out.y;
    src.destroy();
    pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL;
    pic.m_picOrg[0] = pic.m_picOrg[1] = pic.m_picOrg[2] = NULL;
    return cost;
}

} // extern "C"

/* ------------------------------------------------------------------------------------------
 * Frame-level CTU-analysis workload over the REAL reference code: MotionEstimate class for the
 * ME stage, the post-alias C primitive table for the residual and intra stages.  Same driver
 * source as the oracle (frame_driver.h); this is bench.py's `--impl reference` arm.
 * ------------------------------------------------------------------------------------------ */
#define DRV_PIXEL pixel
#define DRV_DEPTH X265_DEPTH
#include "frame_spec.h"

static int ref_drv_me(const void* fv, const fs_me_job* j, int* qmv);
static inline int lg2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }
#define DRV_ME(f, j, qmv) ref_drv_me(f, j, qmv)
#define DRV_MC(src, ss, dst, ds, S, xf, yf) do { int part_ = partitionFromSizes(S, S); \
        if (!((xf) | (yf))) g_c.pu[part_].copy_pp(dst, ds, src, ss); \
        else if (!(yf)) g_c.pu[part_].luma_hpp(src, ss, dst, ds, xf); \
        else if (!(xf)) g_c.pu[part_].luma_vpp(src, ss, dst, ds, yf); \
        else g_c.pu[part_].luma_hvpp(src, ss, dst, ds, xf, yf); } while (0)
#define DRV_SUB_PS(d, ds, a, b, sa, sb, T) g_c.cu[lg2(T) - 2].sub_ps(d, ds, a, b, sa, sb)
#define DRV_DCT(src, dst, stride, T) g_c.cu[lg2(T) - 2].dct(src, dst, stride)
#define DRV_IDCT(src, dst, stride, T) g_c.cu[lg2(T) - 2].idct(src, dst, stride)
#define DRV_QUANT(c, qc, du, q, qbits, add, n) g_c.quant(c, qc, du, q, qbits, add, n)
#define DRV_DEQUANT(q, c, n, scale, shift) g_c.dequant_normal(q, c, n, scale, shift)
#define DRV_BLOCKFILL(d, ds, v, T) g_c.cu[lg2(T) - 2].blockfill_s[0](d, ds, v)
#define DRV_ADD_PS(d, ds, a, r, sa, sr, T) g_c.cu[lg2(T) - 2].add_ps[0](d, ds, a, r, sa, sr)
#define DRV_COPY_PP(d, ds, s, ss, T) g_c.cu[lg2(T) - 2].copy_pp(d, ds, s, ss)
#define DRV_SSE(a, sa, b, sb, T) g_c.cu[lg2(T) - 2].sse_pp(a, sa, b, sb)
#define DRV_INTRA_FILTER(nb, f, S) g_c.cu[lg2(S) - 2].intra_filter(nb, f)
#define DRV_INTRA_PRED(d, ds, nb, mode, bf, S) g_c.cu[lg2(S) - 2].intra_pred[mode](d, ds, nb, mode, bf)
#define DRV_USE_FILTERED(mode, S) ((g_intraFilterFlags[mode] & (S)) != 0)
#define DRV_SA8D(a, sa, b, sb, S) g_c.cu[lg2(S) - 2].sa8d(a, sa, b, sb)
#include "frame_driver.h"

static int ref_drv_me(const void* fv, const fs_me_job* j, int* qmv)
{
    const drv_frame* f = (const drv_frame*)fv;
    if (f->chroma)
    {
        /* the encoder's 4:2:0 call: Yuv variant of setSourcePU with bChroma = true (motion.cpp:194-222), which switches the
         * chroma-SATD term of subpelCompare on for the PUs that have a chroma SATD (see x265ref_motion_estimate_chroma) */
        static thread_local MotionEstimate* mec = NULL;
        static thread_local Yuv* src = NULL;
        static thread_local int mecqp = -1;
        if (!mec) { mec = new MotionEstimate(); mec->init(X265_CSP_I420); src = new Yuv(); src->create(64, X265_CSP_I420); }
        if (mecqp != f->p.qp) { mec->setQP(f->p.qp); mecqp = f->p.qp; }
        const intptr_t stride = f->p.stride, cstride = f->cstride;
        const intptr_t py = j->offset / stride, px = j->offset % stride;
        const intptr_t coff = (py >> 1) * cstride + (px >> 1);
        const pixel* fy = (const pixel*)f->fenc + j->offset;
        for (int y = 0; y < j->ph; y++) memcpy(src->m_buf[0] + y * src->m_size, fy + y * stride, j->pw * sizeof(pixel));
        for (int y = 0; y < j->ph / 2; y++)
        {
            memcpy(src->m_buf[1] + y * src->m_csize, (const pixel*)f->fencC[0] + coff + y * cstride, (j->pw / 2) * sizeof(pixel));
            memcpy(src->m_buf[2] + y * src->m_csize, (const pixel*)f->fencC[1] + coff + y * cstride, (j->pw / 2) * sizeof(pixel));
        }
        mec->setSourcePU(*src, 0, 0, 0, j->pw, j->ph, j->method, j->subme, true);
        PicYuv pic;
        intptr_t zero = 0;
        pic.m_cuOffsetY = &zero; pic.m_cuOffsetC = &zero; pic.m_buOffsetY = &zero; pic.m_buOffsetC = &zero;
        pic.m_stride = stride; pic.m_strideC = cstride;
        ReferencePlanes ref;
        ref.reconPic = &pic;
        ref.lumaStride = stride;
        ref.isLowres = false;
        ref.fpelPlane[0] = (pixel*)f->refs[j->ref] + j->offset;
        ref.fpelPlane[1] = (pixel*)f->refC[j->ref][0] + coff; ref.fpelPlane[2] = (pixel*)f->refC[j->ref][1] + coff;
        pic.m_picOrg[0] = ref.fpelPlane[0]; pic.m_picOrg[1] = ref.fpelPlane[1]; pic.m_picOrg[2] = ref.fpelPlane[2];
        MV mn(j->mvmin[0], j->mvmin[1]), mx(j->mvmax[0], j->mvmax[1]), mvp(j->qmvp[0], j->qmvp[1]), out;
        MV cands[4];
        for (int i = 0; i < j->numCand && i < 4; i++) cands[i] = MV(j->mvc[2 * i], j->mvc[2 * i + 1]);
        int cost = mec->motionEstimate(&ref, mn, mx, mvp, j->numCand, cands, j->merange, out, 1, NULL);
        qmv[0] = out.x; qmv[1] = out.y;
        pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL;
        pic.m_picOrg[0] = pic.m_picOrg[1] = pic.m_picOrg[2] = NULL;
        return cost;
    }
    static thread_local MotionEstimate* me = NULL;
    static thread_local int meqp = -1;
    if (!me) { me = new MotionEstimate(); me->init(X265_CSP_I400); }
    if (meqp != f->p.qp) { me->setQP(f->p.qp); meqp = f->p.qp; }
    me->setSourcePU((pixel*)f->fenc, f->p.stride, j->offset, j->pw, j->ph, j->method, j->method, j->method, j->subme);
    ReferencePlanes ref;
    ref.lumaStride = f->p.stride;
    ref.isLowres = false;
    ref.fpelPlane[0] = (pixel*)f->refs[j->ref];
    MV mn(j->mvmin[0], j->mvmin[1]), mx(j->mvmax[0], j->mvmax[1]), mvp(j->qmvp[0], j->qmvp[1]), out;
    MV cands[4];
    for (int i = 0; i < j->numCand && i < 4; i++) cands[i] = MV(j->mvc[2 * i], j->mvc[2 * i + 1]);
    int cost = me->motionEstimate(&ref, mn, mx, mvp, j->numCand, cands, j->merange, out, 1, NULL);
    qmv[0] = out.x; qmv[1] = out.y;
    return cost;
}

extern "C" int x265ref_analyse_frame(drv_frame* f, int stages)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    { BitCost bc; bc.setQP(f->p.qp); }     /* build the shared cost tables before threads start */
    drv_prepare(f);
    if (stages & 1) drv_run_stage(f, 0);
    if (stages & 2) drv_run_stage(f, 1);
    if (stages & 4) drv_run_stage(f, 2);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Lookahead (BASELINE configs[1]): the REAL Lowres::init, LookaheadTLD::lowresIntraEstimate
 * (slicetype.cpp:696-805) and CostEstimateGroup::singleCost -> estimateFrameCost -> estimateCUCost
 * (slicetype.cpp:3021-3388) on caller frames, serial (non-coop) path, no weightp / AQ / HME.
 * ------------------------------------------------------------------------------------------ */
#include <vector>
#include "slicetype.h"
#include "picyuv.h"
#include "x265.h"

struct RefLookahead
{
    x265_param* param;
    Lookahead* la;
    int n;
    PicYuv** pics;
    Lowres** frames;
    std::vector<Frame*> aqFrames;
};

extern "C" {

static void* la_create(int width, int height, int nframes, const pixel* const* luma, intptr_t stride, int bframes, int lslices,
                       int aqMode = 0, const pixel* const* cb = NULL, const pixel* const* cr = NULL, intptr_t strideC = 0);
void* x265ref_la_create(int width, int height, int nframes, const pixel* const* luma, intptr_t stride, int bframes)
{
    return la_create(width, height, nframes, luma, stride, bframes, 0);
}
/* the same with cooperative lookahead slices (param.cpp:173 medium = 8, :492 slow = 4; slicetype.cpp:1016-1040 clamps): the
 * Lookahead gets a ThreadPool with no running workers, so estimateFrameCost takes the cooperative path
 * (slicetype.cpp:3143-3173) and the calling thread processes every slice itself (tryBondPeers finds no sleeping peer) */
void* x265ref_la_create_slices(int width, int height, int nframes, const pixel* const* luma, intptr_t stride, int bframes, int lslices)
{
    return la_create(width, height, nframes, luma, stride, bframes, lslices);
}
/* the same with adaptive quantisation on (param.cpp:268 default aqMode = AUTO_VARIANCE, strength 1.0): every frame is a real
 * Frame (fenc PicYuv with chroma + Lowres); LookaheadTLD::calcAdaptiveQuantFrame (slicetype.cpp:444-694) fills
 * Lowres::invQscaleFactor / qpAqOffset / wp_sum / wp_ssd before lowresIntraEstimate, as PreLookaheadGroup::processTasks does
 * (slicetype.cpp:1390-1397).  4:2:0 chroma planes are needed (acEnergyCu adds the chroma energy). */
void* x265ref_la_create_aq(int width, int height, int nframes, const pixel* const* luma, const pixel* const* cb, const pixel* const* cr,
                           intptr_t stride, intptr_t strideC, int bframes, int lslices, int aqMode)
{
    return la_create(width, height, nframes, luma, stride, bframes, lslices, aqMode, cb, cr, strideC);
}
}
#include "frame.h"
#include <thread>
#include "predict.h"
#include "cudata.h"
#include "slice.h"
#include "framedata.h"
static void* la_create(int width, int height, int nframes, const pixel* const* luma, intptr_t stride, int bframes, int lslices,
                       int aqMode, const pixel* const* cb, const pixel* const* cr, intptr_t strideC)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    RefLookahead* h = new RefLookahead();
    x265_param* p = x265_param_alloc();
    x265_param_default(p);
    p->sourceWidth = width; p->sourceHeight = height; p->internalCsp = X265_CSP_I420;
    p->bframes = bframes; p->bEnableWeightedPred = 0; p->bEnableWeightedBiPred = 0;
    p->rc.aqMode = aqMode; p->rc.cuTree = 0; p->rc.hevcAq = 0; p->bAQMotion = 0; p->lookaheadSlices = lslices; p->bEnableHME = 0;
    p->rc.qgSize = 32; p->maxCUSize = 64; p->rc.vbvBufferSize = 0; p->bFrameAdaptive = 0;
    h->param = p; h->n = nframes;
    h->pics = new PicYuv*[nframes]; h->frames = new Lowres*[nframes + 2];
    for (int i = 0; i < nframes && aqMode; i++)
    {
        /* AQ path: a real Frame per picture */
        Frame* fr = new Frame();
        fr->m_param = p;
        fr->create(p, NULL);
        PicYuv* pic = fr->m_fencPic;
        for (int y = 0; y < height; y++)
            memcpy(pic->m_picOrg[0] + (intptr_t)y * pic->m_stride, luma[i] + (intptr_t)y * stride, (size_t)width * sizeof(pixel));
        for (int y = 0; y < height / 2; y++)
        {
            memcpy(pic->m_picOrg[1] + (intptr_t)y * pic->m_strideC, cb[i] + (intptr_t)y * strideC, (size_t)(width / 2) * sizeof(pixel));
            memcpy(pic->m_picOrg[2] + (intptr_t)y * pic->m_strideC, cr[i] + (intptr_t)y * strideC, (size_t)(width / 2) * sizeof(pixel));
        }
        extendPicBorder(pic->m_picOrg[0], pic->m_stride, width, height, pic->m_lumaMarginX, pic->m_lumaMarginY);
        extendPicBorder(pic->m_picOrg[1], pic->m_strideC, width / 2, height / 2, pic->m_chromaMarginX, pic->m_chromaMarginY);
        extendPicBorder(pic->m_picOrg[2], pic->m_strideC, width / 2, height / 2, pic->m_chromaMarginX, pic->m_chromaMarginY);
        fr->m_lowres.init(pic, i);
        h->pics[i] = pic; h->frames[i] = &fr->m_lowres;
        h->aqFrames.push_back(fr);
    }
    for (int i = 0; i < nframes && !aqMode; i++)
    {
        PicYuv* pic = new PicYuv();
        pic->create(p, true, NULL);
        for (int y = 0; y < height; y++)
            memcpy(pic->m_picOrg[0] + (intptr_t)y * pic->m_stride, luma[i] + (intptr_t)y * stride, (size_t)width * sizeof(pixel));
        extendPicBorder(pic->m_picOrg[0], pic->m_stride, width, height, pic->m_lumaMarginX, pic->m_lumaMarginY);
        Lowres* lr = new Lowres();
        memset((void*)lr, 0, sizeof(Lowres));
        lr->create(p, pic, p->rc.qgSize);
        lr->init(pic, i);
        h->pics[i] = pic; h->frames[i] = lr;
    }
    ThreadPool* pool = NULL;
    if (lslices > 0)
    {
        pool = new ThreadPool();
        pool->create(1, 1, 0);                 /* never started: no worker ever sleeps, so no peer can be bonded */
    }
    h->la = new Lookahead(p, pool);
    const int ntld = 1 + (pool ? pool->m_numWorkers : 0);
    h->la->m_tld = new LookaheadTLD[ntld];
    for (int t = 0; t < ntld; t++)
        h->la->m_tld[t].init(h->la->m_8x8Width, h->la->m_8x8Height, h->la->m_cuCount);
    for (int i = 0; i < nframes; i++)
    {
        if (aqMode) h->la->m_tld[0].calcAdaptiveQuantFrame(h->aqFrames[i], p);
        h->la->m_tld[0].lowresIntraEstimate(*h->frames[i], p->rc.qgSize);
    }
    return h;
}
extern "C" {
/* out = {numCoopSlices, numRowsPerSlice} as the Lookahead constructor settled them (slicetype.cpp:1016-1040) */
void x265ref_la_slices(void* hv, int* out)
{
    RefLookahead* h = (RefLookahead*)hv;
    out[0] = h->la->m_numCoopSlices; out[1] = h->la->m_numRowsPerSlice;
}


int64_t x265ref_la_cost(void* hv, int p0, int p1, int b)
{
    RefLookahead* h = (RefLookahead*)hv;
    CostEstimateGroup est(*h->la, h->frames);
    return est.singleCost(p0, p1, b, false);
}

/* LookaheadTLD::weightsAnalyse (slicetype.cpp:860-961, with weightCostLuma :807-840) on frames b (fenc) and p0 (ref), with the
 * picture statistics Lowres::wp_sum[0] / wp_ssd[0] supplied by the caller (the reference accumulates them in
 * calcAdaptiveQuantFrame, slicetype.cpp:49-57, 462-480, 665-676: inputs of the analysis, not part of it).
 * stats = {fenc sum, fenc ssd, ref sum, ref ssd}.  Returns weightedRef[b - p0].isWeighted; out = {planesize (pixels),
 * paddedLines}; when weighted and wplanes != NULL, the 4 re-weighted planes (4 * planesize pixels) are copied out. */
int x265ref_la_weights(void* hv, int b, int p0, const uint64_t* stats, int64_t* out, pixel* wplanes)
{
    RefLookahead* h = (RefLookahead*)hv;
    Lowres& fenc = *h->frames[b];
    Lowres& ref = *h->frames[p0];
    fenc.wp_sum[0] = stats[0]; fenc.wp_ssd[0] = stats[1]; ref.wp_sum[0] = stats[2]; ref.wp_ssd[0] = stats[3];
    LookaheadTLD& tld = h->la->m_tld[0];
    fenc.weightedRef[b - p0].isWeighted = false;
    tld.weightsAnalyse(fenc, ref);
    const intptr_t planesize = fenc.buffer[1] - fenc.buffer[0];
    out[0] = (int64_t)planesize; out[1] = tld.paddedLines;
    const int w = fenc.weightedRef[b - p0].isWeighted ? 1 : 0;
    if (w && wplanes) memcpy(wplanes, tld.wbuffer[0], sizeof(pixel) * 4 * planesize);
    return w;
}

/* geometry: out = {width8, height8, lumaStride, lowres width, lowres lines, marginX, marginY} */
void x265ref_la_geometry(void* hv, int* out)
{
    RefLookahead* h = (RefLookahead*)hv;
    Lowres* f = h->frames[0];
    out[0] = h->la->m_8x8Width; out[1] = h->la->m_8x8Height; out[2] = (int)f->lumaStride; out[3] = f->width; out[4] = f->lines;
    out[5] = h->pics[0]->m_lumaMarginX; out[6] = h->pics[0]->m_lumaMarginY;
}

/* what: 0 intraCost[int32], 1 intraMode[u8], 2 lowresCosts[d0][d1][u16], 3 rowSatds[d0][d1][int32 per row],
 * 4 lowresMvs[list][dist] as int32 pairs, 5 lowresMvCosts[list][dist][int32], 6 {costEst, costEstAq, intraMbs[d0]} as int64 x3,
 * 7 lowres plane `d0` (0..3) rows [-marginY, lines+marginY) x lumaStride */
int x265ref_la_get(void* hv, int frame, int what, int d0, int d1, void* out)
{
    RefLookahead* h = (RefLookahead*)hv;
    Lowres* f = h->frames[frame];
    const int ncu = h->la->m_cuCount, rows = h->la->m_8x8Height;
    switch (what)
    {
    case 0: memcpy(out, f->intraCost, sizeof(int32_t) * ncu); break;
    case 1: memcpy(out, f->intraMode, ncu); break;
    case 2: memcpy(out, f->lowresCosts[d0][d1], sizeof(uint16_t) * ncu); break;
    case 3: memcpy(out, f->rowSatds[d0][d1], sizeof(int32_t) * rows); break;
    case 4: for (int i = 0; i < ncu; i++) { ((int32_t*)out)[2 * i] = f->lowresMvs[d0][d1][i].x; ((int32_t*)out)[2 * i + 1] = f->lowresMvs[d0][d1][i].y; } break;
    case 5: memcpy(out, f->lowresMvCosts[d0][d1], sizeof(int32_t) * ncu); break;
    case 6: ((int64_t*)out)[0] = f->costEst[d0][d1]; ((int64_t*)out)[1] = f->costEstAq[d0][d1]; ((int64_t*)out)[2] = f->intraMbs[d0]; break;
    case 7:
    {
        const int my = h->pics[0]->m_lumaMarginY, mx = h->pics[0]->m_lumaMarginX;
        memcpy(out, f->lowresPlane[d0] - (intptr_t)my * f->lumaStride - mx, sizeof(pixel) * f->lumaStride * (f->lines + 2 * my));
        break;
    }
    case 8:     /* Lowres::invQscaleFactor as int32 per lowres CU (rc.qgSize 32 -> one per 8x8 lowres CU), 0 words when AQ is off */
        if (!f->invQscaleFactor) return -1;
        for (int i = 0; i < ncu; i++) ((int32_t*)out)[i] = f->invQscaleFactor[i];
        break;
    case 9:     /* {wp_sum[0], wp_ssd[0]} of luma as calcAdaptiveQuantFrame leaves them */
        ((uint64_t*)out)[0] = f->wp_sum[0]; ((uint64_t*)out)[1] = f->wp_ssd[0];
        break;
    default: return -1;
    }
    return 0;
}


/* Prediction costs around the motion search on the REAL classes (pins oracle/oracle_pred.c):
 * Predict::motionCompensation (common/predict.cpp:76-240) / predInterLumaPixel for the prediction, then
 * MotionEstimate::bufSAD / bufSATD / bufChromaSATD (encoder/motion.h:87-95) exactly as Search::selectMVP
 * (search.cpp:1992-2023), Search::mergeEstimation (:1901-1960) and the bidir block of predInterSearch (:2474-2607) call
 * them.  All planes are handed over at the PU origin (PicYuv offset tables = a single zero, as in
 * x265ref_motion_estimate_chroma); the slice is a B slice without weighted prediction; the picture is made large
 * enough that CUData::clipMv leaves the vectors alone (the batched entry point takes clipped vectors).
 * ref0 / ref1 == NULL: list unused.  cost: 0 = SAD, 1 = SATD; chroma: + bufChromaSATD; biAvgPP: search.cpp:2499-2510. */
struct RefPredCtx
{
    Yuv src, predYuv, bidirYuv[2];
    MotionEstimate me;
    Predict pred;
    PicYuv pic[2];
    x265_param param;
    SPS* sps; PPS* pps; FrameData* fd; Slice* slice; CUData* cu;
    intptr_t zero;
    bool ok;
    RefPredCtx()
    {
        ok = src.create(64, X265_CSP_I420) && predYuv.create(64, X265_CSP_I420) && bidirYuv[0].create(64, X265_CSP_I420) && bidirYuv[1].create(64, X265_CSP_I420);
        me.init(X265_CSP_I420);
        me.setQP(30);
        pred.allocBuffers(X265_CSP_I420);
        zero = 0;
        memset(&param, 0, sizeof(param)); param.maxCUSize = 64;
        sps = (SPS*)calloc(1, sizeof(SPS)); pps = (PPS*)calloc(1, sizeof(PPS)); fd = (FrameData*)calloc(1, sizeof(FrameData));
        slice = (Slice*)calloc(1, sizeof(Slice)); cu = (CUData*)calloc(1, sizeof(CUData));
        sps->picWidthInLumaSamples = 1 << 15; sps->picHeightInLumaSamples = 1 << 15;
        pps->bUseWeightPred = false; pps->bUseWeightedBiPred = false;
        fd->m_param = &param;
        slice->m_sps = sps; slice->m_pps = pps; slice->m_sliceType = B_SLICE;
        slice->m_numRefIdx[0] = slice->m_numRefIdx[1] = 1;
        slice->m_refReconPicList[0][0] = &pic[0]; slice->m_refReconPicList[1][0] = &pic[1];
        cu->m_slice = slice; cu->m_encData = fd;
        cu->m_cuPelX = 1 << 14; cu->m_cuPelY = 1 << 14;
        for (int l = 0; l < 2; l++)
        {
            pic[l].m_cuOffsetY = &zero; pic[l].m_cuOffsetC = &zero; pic[l].m_buOffsetY = &zero; pic[l].m_buOffsetC = &zero;
        }
    }
    ~RefPredCtx()
    {
        for (int l = 0; l < 2; l++)
        {
            pic[l].m_cuOffsetY = pic[l].m_cuOffsetC = pic[l].m_buOffsetY = pic[l].m_buOffsetC = NULL;
            pic[l].m_picOrg[0] = pic[l].m_picOrg[1] = pic[l].m_picOrg[2] = NULL;
        }
        src.destroy(); predYuv.destroy(); bidirYuv[0].destroy(); bidirYuv[1].destroy();
        free(sps); free(pps); free(fd); free(slice); free(cu);
    }
    int run(const pixel* const* fenc, const pixel* const* ref0, const pixel* const* ref1, intptr_t stride, intptr_t cstride,
            int pw, int ph, const int* mv0, const int* mv1, int cost, int chroma, int biAvgPP)
    {
        for (int y = 0; y < ph; y++) memcpy(src.m_buf[0] + y * src.m_size, fenc[0] + y * stride, pw * sizeof(pixel));
        for (int y = 0; y < ph / 2; y++)
        {
            memcpy(src.m_buf[1] + y * src.m_csize, fenc[1] + y * cstride, (pw / 2) * sizeof(pixel));
            memcpy(src.m_buf[2] + y * src.m_csize, fenc[2] + y * cstride, (pw / 2) * sizeof(pixel));
        }
        me.setSourcePU(src, 0, 0, 0, pw, ph, X265_STAR_SEARCH, chroma ? 3 : 2, !!chroma);
        const bool bChromaSATD = me.bChromaSATD;
        for (int l = 0; l < 2; l++)
        {
            const pixel* const* r = l ? ref1 : ref0;
            pic[l].m_stride = stride; pic[l].m_strideC = cstride;
            for (int k = 0; k < 3; k++) pic[l].m_picOrg[k] = r ? (pixel*)r[k] : NULL;
        }
        int8_t refIdx[2] = { (int8_t)(ref0 ? 0 : -1), (int8_t)(ref1 ? 0 : -1) };
        MV mvs[2] = { MV(mv0[0], mv0[1]), MV(mv1[0], mv1[1]) };
        cu->m_refIdx[0] = &refIdx[0]; cu->m_refIdx[1] = &refIdx[1];
        cu->m_mv[0] = &mvs[0]; cu->m_mv[1] = &mvs[1];
        alignas(PredictionUnit) char pubuf[sizeof(PredictionUnit)];
        PredictionUnit& pu = *(PredictionUnit*)pubuf;
        pu.ctuAddr = 0; pu.cuAbsPartIdx = 0; pu.puAbsPartIdx = 0; pu.width = pw; pu.height = ph;
        int out;
        if (ref0 && ref1 && biAvgPP)
        {   /* search.cpp:2499-2510 */
            pred.predInterLumaPixel(pu, bidirYuv[0], pic[0], mvs[0]);
            pred.predInterLumaPixel(pu, bidirYuv[1], pic[1], mvs[1]);
            primitives.pu[me.partEnum].pixelavg_pp[(predYuv.m_size % 64 == 0) && (bidirYuv[0].m_size % 64 == 0) && (bidirYuv[1].m_size % 64 == 0)](
                predYuv.m_buf[0], predYuv.m_size, bidirYuv[0].getLumaAddr(0), bidirYuv[0].m_size, bidirYuv[1].getLumaAddr(0), bidirYuv[1].m_size, 32);
            out = me.bufSATD(predYuv.m_buf[0], predYuv.m_size);
        }
        else if (!cost)
        {   /* search.cpp:2017-2018 */
            pred.predInterLumaPixel(pu, predYuv, ref0 ? pic[0] : pic[1], ref0 ? mvs[0] : mvs[1]);
            out = me.bufSAD(predYuv.getLumaAddr(0), predYuv.m_size);
        }
        else
        {   /* search.cpp:1944-1948, 2489-2493 */
            pred.motionCompensation(*cu, pu, predYuv, true, bChromaSATD);
            out = me.bufSATD(predYuv.getLumaAddr(0), predYuv.m_size);
            if (bChromaSATD) out += me.bufChromaSATD(predYuv, 0);
        }
        cu->m_refIdx[0] = cu->m_refIdx[1] = NULL; cu->m_mv[0] = cu->m_mv[1] = NULL;
        return out;
    }
};

int x265ref_pred_cost(const pixel* const* fenc, const pixel* const* ref0, const pixel* const* ref1, intptr_t stride, intptr_t cstride,
                      int pw, int ph, const int* mv0, const int* mv1, int cost, int chroma, int biAvgPP)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    RefPredCtx c;
    if (!c.ok) return -1;
    return c.run(fenc, ref0, ref1, stride, cstride, pw, ph, mv0, mv1, cost, chroma, biAvgPP);
}

/* The same over a job list in the batched entry point's layout (include/x265_b200.h: x265cu_pred_job), `nthreads` host
 * threads, one context (Predict, MotionEstimate, Yuv buffers) per thread: bench.py's CPU arm of the `pred_cost` leg.
 * Planes are given at the picture origin (pixel (0,0)); refs / refCb / refCr are tables indexed by job.ref0 / ref1. */
struct ref_pred_job { int32_t offset; int16_t pw, ph; int8_t ref0, ref1; uint8_t cost, flags; int16_t mv0[2], mv1[2]; };
int x265ref_pred_cost_batch(const pixel* fenc, const pixel* fencCb, const pixel* fencCr, const pixel* const* refs, const pixel* const* refCb,
                            const pixel* const* refCr, intptr_t stride, intptr_t cstride, const void* jobsv, int n, int32_t* out, int nthreads)
{
    ensure_init();
    static bool scales = false;
    if (!scales) { MotionEstimate::initScales(); scales = true; }
    const ref_pred_job* jobs = (const ref_pred_job*)jobsv;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=]() {
            RefPredCtx c;
            for (int i = t; i < n; i += nthreads)
            {
                const ref_pred_job& j = jobs[i];
                const intptr_t py = j.offset / stride, px = j.offset % stride;
                const intptr_t coff = (py >> 1) * cstride + (px >> 1);
                const pixel* f[3] = { fenc + j.offset, fencCb + coff, fencCr + coff };
                const pixel* r0[3]; const pixel* r1[3];
                if (j.ref0 >= 0) { r0[0] = refs[j.ref0] + j.offset; r0[1] = refCb[j.ref0] + coff; r0[2] = refCr[j.ref0] + coff; }
                if (j.ref1 >= 0) { r1[0] = refs[j.ref1] + j.offset; r1[1] = refCb[j.ref1] + coff; r1[2] = refCr[j.ref1] + coff; }
                const int mv0[2] = { j.mv0[0], j.mv0[1] }, mv1[2] = { j.mv1[0], j.mv1[1] };
                out[i] = c.run(f, j.ref0 >= 0 ? r0 : NULL, j.ref1 >= 0 ? r1 : NULL, stride, cstride, j.pw, j.ph, mv0, mv1, j.cost, j.flags & 1, (j.flags >> 1) & 1);
            }
        });
    for (auto& x : th) x.join();
    return 0;
}

} // extern "C"

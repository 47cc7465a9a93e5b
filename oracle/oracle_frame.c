/* oracle/oracle_frame.c -- TEST INFRASTRUCTURE ONLY (see oracle.h, frame_spec.h, frame_driver.h).
 * Instantiates the frame-level CTU-analysis driver over the oracle's plain-C primitives. */
#include "oracle.h"
#define DRV_PIXEL pixel
#define DRV_DEPTH ORC_DEPTH
#include "frame_spec.h"

struct drv_frame_s;
static int orc_drv_me(const void* fv, const fs_me_job* j, int* qmv);
#define DRV_ME(f, j, qmv) orc_drv_me(f, j, qmv)
#define DRV_MC(src, ss, dst, ds, S, xf, yf) do { \
        if (!((xf) | (yf))) orc_copy_pp(dst, ds, src, ss, S, S); \
        else if (!(yf)) orc_interp_hpp(src, ss, dst, ds, xf, 8, S, S); \
        else if (!(xf)) orc_interp_vpp(src, ss, dst, ds, yf, 8, S, S); \
        else orc_interp_hvpp(src, ss, dst, ds, xf, yf, 8, S, S); } while (0)
#define DRV_SUB_PS(d, ds, a, b, sa, sb, T) orc_sub_ps(d, ds, a, b, sa, sb, T, T)
#define DRV_DCT(src, dst, stride, T) orc_dct(src, dst, stride, T)
#define DRV_IDCT(src, dst, stride, T) orc_idct(src, dst, stride, T)
#define DRV_QUANT(c, qc, du, q, qbits, add, n) orc_quant(c, qc, du, q, qbits, add, n)
#define DRV_DEQUANT(q, c, n, scale, shift) orc_dequant_normal(q, c, n, scale, shift)
#define DRV_BLOCKFILL(d, ds, v, T) orc_blockfill_s(d, ds, v, T)
#define DRV_ADD_PS(d, ds, a, r, sa, sr, T) orc_add_ps(d, ds, a, r, sa, sr, T, T)
#define DRV_COPY_PP(d, ds, s, ss, T) orc_copy_pp(d, ds, s, ss, T, T)
#define DRV_SSE(a, sa, b, sb, T) orc_sse_pp(a, sa, b, sb, T, T)
#define DRV_INTRA_FILTER(nb, f, S) orc_intra_filter(nb, f, S)
#define DRV_INTRA_PRED(d, ds, nb, mode, bf, S) orc_intra_pred(d, ds, nb, mode, bf, S)
#define DRV_USE_FILTERED(mode, S) orc_intra_use_filtered(mode, S)
#define DRV_SA8D(a, sa, b, sb, S) orc_sa8d(a, sa, b, sb, S, S)

int orc_intra_use_filtered(int mode, int n)
{
    if (mode == 1 || n == 4) return 0;
    if (mode == 0) return n >= 8;
    int d1 = mode > 26 ? mode - 26 : 26 - mode, d2 = mode > 10 ? mode - 10 : 10 - mode;
    int d = d1 < d2 ? d1 : d2;
    return d > (n == 8 ? 7 : (n == 16 ? 1 : 0));
}

#include "frame_driver.h"

static int orc_drv_me(const void* fv, const fs_me_job* j, int* qmv)
{
    const drv_frame* f = (const drv_frame*)fv;
    orc_me_job job;
    int mvc[8];
    job.chroma = f->chroma;
    if (f->chroma)
    {
        job.fencC[0] = f->fencC[0]; job.fencC[1] = f->fencC[1];
        job.refC[0] = f->refC[j->ref][0]; job.refC[1] = f->refC[j->ref][1];
        job.cstride = f->cstride;
    }
    job.fenc = f->fenc; job.fencStride = f->p.stride; job.offset = j->offset;
    job.ref[0] = job.ref[1] = job.ref[2] = job.ref[3] = f->refs[j->ref];
    job.refStride = f->p.stride; job.lowres = 0; job.pw = j->pw; job.ph = j->ph;
    job.method = j->method; job.subme = j->subme;
    job.mvmin[0] = j->mvmin[0]; job.mvmin[1] = j->mvmin[1]; job.mvmax[0] = j->mvmax[0]; job.mvmax[1] = j->mvmax[1];
    job.qmvp[0] = j->qmvp[0]; job.qmvp[1] = j->qmvp[1];
    for (int i = 0; i < 8; i++) mvc[i] = j->mvc[i];
    job.numCand = j->numCand; job.mvc = mvc; job.merange = j->merange; job.mvcost = f->mvcost;
    return orc_motion_estimate(&job, qmv);
}

int orc_frame_counts(const fs_params* p, int* njobs, int* ncu, int64_t* ncoef)
{
    *njobs = fs_build_me_jobs(p, NULL, NULL);
    *ncu = fs_build_cus(p->width, p->height, NULL);
    int16_t (*cus)[3] = (int16_t (*)[3])malloc(sizeof(int16_t) * 3 * (size_t)*ncu);
    fs_build_cus(p->width, p->height, cus);
    int64_t n = 0;
    for (int c = 0; c < *ncu; c++) n += (int64_t)cus[c][2] * cus[c][2];
    free(cus);
    *ncoef = n;
    return 0;
}

/* stages: bit 0 ME, bit 1 residual, bit 2 intra (prepare always runs) */
int orc_analyse_frame(drv_frame* f, int stages)
{
    drv_prepare(f);
    if (stages & 1) drv_run_stage(f, 0);
    if (stages & 2) drv_run_stage(f, 1);
    if (stages & 4) drv_run_stage(f, 2);
    return 0;
}

/* x265_b200/plugin/setup_cuda_primitives.cpp -- the drop-in boundary.
 *
 *   void setupCudaPrimitives(EncoderPrimitives& p, int cpuMask)
 *
 * is the sibling of setupAssemblyPrimitives() (/root/reference/source/common/primitives.h:470) and is
 * meant to be called at the same place, between setupCPrimitives() and setupAliasPrimitives()
 * (/root/reference/source/common/primitives.cpp:252-276), so the alias pass propagates our pointers.
 * It overwrites exactly the fields libx265cu.so provides (x265cu_get_primitive) and leaves every other
 * field untouched (the C entry stays).  This file is compiled against the reference's own headers
 * (never copied into this repo); see INTEGRATION.md.  If no CUDA device is present
 * x265cu_get_primitive() returns NULL for everything and the table is left as it was -- the caller can
 * test x265cu_device_count() and abort, there is no CPU fallback inside the CUDA primitives.
 */
#include "common.h"
#include "primitives.h"
#include "x265_b200.h"

namespace X265_NS {

#define CU_GET(name, i, j, k) x265cu_get_primitive(X265_DEPTH, name, i, j, k)
#define SET(field, type, name, i, j, k) do { void* f_ = CU_GET(name, i, j, k); if (f_) field = (type)f_; } while (0)

void setupCudaPrimitives(EncoderPrimitives& p, int /*cpuMask*/)
{
    for (int i = 0; i < NUM_PU_SIZES; i++)
    {
        SET(p.pu[i].sad, pixelcmp_t, "pu.sad", i, 0, 0);
        SET(p.pu[i].sad_x3, pixelcmp_x3_t, "pu.sad_x3", i, 0, 0);
        SET(p.pu[i].sad_x4, pixelcmp_x4_t, "pu.sad_x4", i, 0, 0);
        SET(p.pu[i].ads, pixelcmp_ads_t, "pu.ads", i, 0, 0);
        SET(p.pu[i].satd, pixelcmp_t, "pu.satd", i, 0, 0);
        SET(p.pu[i].luma_hpp, filter_pp_t, "pu.luma_hpp", i, 0, 0);
        SET(p.pu[i].luma_hps, filter_hps_t, "pu.luma_hps", i, 0, 0);
        SET(p.pu[i].luma_vpp, filter_pp_t, "pu.luma_vpp", i, 0, 0);
        SET(p.pu[i].luma_vps, filter_ps_t, "pu.luma_vps", i, 0, 0);
        SET(p.pu[i].luma_vsp, filter_sp_t, "pu.luma_vsp", i, 0, 0);
        SET(p.pu[i].luma_vss, filter_ss_t, "pu.luma_vss", i, 0, 0);
        SET(p.pu[i].luma_hvpp, filter_hv_pp_t, "pu.luma_hvpp", i, 0, 0);
        SET(p.pu[i].copy_pp, copy_pp_t, "pu.copy_pp", i, 0, 0);
        for (int a = 0; a < NUM_ALIGNMENT_TYPES; a++)
        {
            SET(p.pu[i].pixelavg_pp[a], pixelavg_pp_t, "pu.pixelavg_pp", i, a, 0);
            SET(p.pu[i].addAvg[a], addAvg_t, "pu.addAvg", i, a, 0);
            SET(p.pu[i].convert_p2s[a], filter_p2s_t, "pu.convert_p2s", i, a, 0);
        }
        /* 4:2:0 chroma interpolation (indexed by the luma PU, block = W/2 x H/2); only where the C table has one */
        EncoderPrimitives::Chroma::PUChroma& c = p.chroma[X265_CSP_I420].pu[i];
        if (c.filter_hpp) SET(c.filter_hpp, filter_pp_t, "chroma.pu.filter_hpp", i, 0, X265_CSP_I420);
        if (c.filter_hps) SET(c.filter_hps, filter_hps_t, "chroma.pu.filter_hps", i, 0, X265_CSP_I420);
        if (c.filter_vpp) SET(c.filter_vpp, filter_pp_t, "chroma.pu.filter_vpp", i, 0, X265_CSP_I420);
        if (c.filter_vps) SET(c.filter_vps, filter_ps_t, "chroma.pu.filter_vps", i, 0, X265_CSP_I420);
        if (c.filter_vsp) SET(c.filter_vsp, filter_sp_t, "chroma.pu.filter_vsp", i, 0, X265_CSP_I420);
        if (c.filter_vss) SET(c.filter_vss, filter_ss_t, "chroma.pu.filter_vss", i, 0, X265_CSP_I420);
    }
    for (int i = 0; i < NUM_CU_SIZES; i++)
    {
        SET(p.cu[i].dct, dct_t, "cu.dct", i, 0, 0);
        SET(p.cu[i].idct, idct_t, "cu.idct", i, 0, 0);
        SET(p.cu[i].sub_ps, pixel_sub_ps_t, "cu.sub_ps", i, 0, 0);
        SET(p.cu[i].copy_cnt, copy_cnt_t, "cu.copy_cnt", i, 0, 0);
        SET(p.cu[i].count_nonzero, count_nonzero_t, "cu.count_nonzero", i, 0, 0);
        SET(p.cu[i].cpy2Dto1D_shl, cpy2Dto1D_shl_t, "cu.cpy2Dto1D_shl", i, 0, 0);
        SET(p.cu[i].cpy2Dto1D_shr, cpy2Dto1D_shr_t, "cu.cpy2Dto1D_shr", i, 0, 0);
        SET(p.cu[i].cpy1Dto2D_shr, cpy1Dto2D_shr_t, "cu.cpy1Dto2D_shr", i, 0, 0);
        SET(p.cu[i].copy_sp, copy_sp_t, "cu.copy_sp", i, 0, 0);
        SET(p.cu[i].copy_ps, copy_ps_t, "cu.copy_ps", i, 0, 0);
        SET(p.cu[i].copy_ss, copy_ss_t, "cu.copy_ss", i, 0, 0);
        SET(p.cu[i].var, var_t, "cu.var", i, 0, 0);
        SET(p.cu[i].sse_pp, pixel_sse_t, "cu.sse_pp", i, 0, 0);
        SET(p.cu[i].sse_ss, pixel_sse_ss_t, "cu.sse_ss", i, 0, 0);
        SET(p.cu[i].psy_cost_pp, pixelcmp_t, "cu.psy_cost_pp", i, 0, 0);
        SET(p.cu[i].sa8d, pixelcmp_t, "cu.sa8d", i, 0, 0);
        SET(p.cu[i].transpose, transpose_t, "cu.transpose", i, 0, 0);
        SET(p.cu[i].intra_filter, intra_filter_t, "cu.intra_filter", i, 0, 0);
        SET(p.cu[i].intra_pred_allangs, intra_allangs_t, "cu.intra_pred_allangs", i, 0, 0);
        for (int m = 0; m < NUM_INTRA_MODE; m++)
            SET(p.cu[i].intra_pred[m], intra_pred_t, "cu.intra_pred", i, m, 0);
        for (int a = 0; a < NUM_ALIGNMENT_TYPES; a++)
        {
            SET(p.cu[i].calcresidual[a], calcresidual_t, "cu.calcresidual", i, a, 0);
            SET(p.cu[i].add_ps[a], pixel_add_ps_t, "cu.add_ps", i, a, 0);
            SET(p.cu[i].blockfill_s[a], blockfill_s_t, "cu.blockfill_s", i, a, 0);
            SET(p.cu[i].cpy1Dto2D_shl[a], cpy1Dto2D_shl_t, "cu.cpy1Dto2D_shl", i, a, 0);
            SET(p.cu[i].ssd_s[a], pixel_ssd_s_t, "cu.ssd_s", i, a, 0);
        }
    }
    SET(p.dst4x4, dct_t, "dst4x4", 0, 0, 0);
    SET(p.idst4x4, idct_t, "idst4x4", 0, 0, 0);
    SET(p.quant, quant_t, "quant", 0, 0, 0);
    SET(p.nquant, nquant_t, "nquant", 0, 0, 0);
    SET(p.dequant_normal, dequant_normal_t, "dequant_normal", 0, 0, 0);
    SET(p.dequant_scaling, dequant_scaling_t, "dequant_scaling", 0, 0, 0);
    SET(p.denoiseDct, denoiseDct_t, "denoiseDct", 0, 0, 0);
    SET(p.propagateCost, cutree_propagate_cost, "propagateCost", 0, 0, 0);
    for (int i = 0; i < NUM_INTEGRAL_SIZE; i++)
    {   // SEA integral planes (framefilter.cpp:39-143, driven row by row from FrameFilter::computeMEIntegral)
        SET(p.integral_inith[i], integralh_t, "integral_inith", i, 0, 0);
        SET(p.integral_initv[i], integralv_t, "integral_initv", i, 0, 0);
    }
    /* in-loop filters: deblocking edge filters and SAO (loopfilter.cpp:184-200, sao.cpp:1927-1935) */
    SET(p.sign, sign_t, "sign", 0, 0, 0);
    SET(p.saoCuOrgE0, saoCuOrgE0_t, "saoCuOrgE0", 0, 0, 0);
    SET(p.saoCuOrgE1, saoCuOrgE1_t, "saoCuOrgE1", 0, 0, 0);
    SET(p.saoCuOrgE1_2Rows, saoCuOrgE1_t, "saoCuOrgE1_2Rows", 0, 0, 0);
    SET(p.saoCuOrgB0, saoCuOrgB0_t, "saoCuOrgB0", 0, 0, 0);
    SET(p.saoCuStatsBO, saoCuStatsBO_t, "saoCuStatsBO", 0, 0, 0);
    SET(p.saoCuStatsE0, saoCuStatsE0_t, "saoCuStatsE0", 0, 0, 0);
    SET(p.saoCuStatsE1, saoCuStatsE1_t, "saoCuStatsE1", 0, 0, 0);
    SET(p.saoCuStatsE2, saoCuStatsE2_t, "saoCuStatsE2", 0, 0, 0);
    SET(p.saoCuStatsE3, saoCuStatsE3_t, "saoCuStatsE3", 0, 0, 0);
    for (int e = 0; e < 2; e++)
    {
        SET(p.saoCuOrgE2[e], saoCuOrgE2_t, "saoCuOrgE2", e, 0, 0);
        SET(p.saoCuOrgE3[e], saoCuOrgE3_t, "saoCuOrgE3", e, 0, 0);
        SET(p.pelFilterLumaStrong[e], pelFilterLumaStrong_t, "pelFilterLumaStrong", e, 0, 0);
        SET(p.pelFilterChroma[e], pelFilterChroma_t, "pelFilterChroma", e, 0, 0);
    }
    SET(p.scale2D_64to32, scale2D_t, "scale2D_64to32", 0, 0, 0);
    SET(p.weight_pp, weightp_pp_t, "weight_pp", 0, 0, 0);
    SET(p.weight_sp, weightp_sp_t, "weight_sp", 0, 0, 0);
    SET(p.frameInitLowres, downscale_t, "frameInitLowres", 0, 0, 0);
    SET(p.frameInitLowerRes, downscale_t, "frameInitLowerRes", 0, 0, 0);
}

} // namespace

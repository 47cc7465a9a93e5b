// tests/geometry_capi.cpp -- CPU-only test shim: exposes the analyser's host-side geometry builder
// (x265_b200/csrc/geometry.h, the code x265cu_analyser_create runs) and the ABI struct layout to tests/test_geometry.py.
#include "../x265_b200/csrc/geometry.h"
#include "../include/x265_b200.h"
#include <stddef.h>
#include <string.h>

static FrameGeometry g_geo;

extern "C" {

int geo_build(int W, int H, int stride, int nref, int rect, int amp, int64_t* counts /* njobs, ncu, ntu, ctuRows, ncoef */)
{
    g_geo = FrameGeometry();
    geometry_build(W, H, stride, nref, rect, amp, g_geo);
    counts[0] = (int64_t)g_geo.pus.size(); counts[1] = (int64_t)g_geo.cus.size(); counts[2] = (int64_t)g_geo.tus.size();
    counts[3] = g_geo.ctuRows; counts[4] = g_geo.ncoef;
    return 0;
}

/* what: 0 pus (int32 x 6: offset, cuX, cuY, pw, ph, ref), 1 cus (int64 x 4: x, y, size, coef_off), 2 tus (int32 x 3: cu, tx, ty),
 * 3 cu_jobs (int32), 4 rowJob, 5 rowCu, 6 rowTu (int32, ctuRows + 1 each) */
int geo_get(int what, void* out)
{
    switch (what)
    {
    case 0: { int32_t* o = (int32_t*)out; for (size_t i = 0; i < g_geo.pus.size(); i++) { const PuDesc& d = g_geo.pus[i];
              o[6 * i] = d.offset; o[6 * i + 1] = d.cuX; o[6 * i + 2] = d.cuY; o[6 * i + 3] = d.pw; o[6 * i + 4] = d.ph; o[6 * i + 5] = d.ref; } break; }
    case 1: { int64_t* o = (int64_t*)out; for (size_t i = 0; i < g_geo.cus.size(); i++) { const CuDesc& c = g_geo.cus[i];
              o[4 * i] = c.x; o[4 * i + 1] = c.y; o[4 * i + 2] = c.size; o[4 * i + 3] = c.coef_off; } break; }
    case 2: { int32_t* o = (int32_t*)out; for (size_t i = 0; i < g_geo.tus.size(); i++) { const TuDesc& t = g_geo.tus[i];
              o[3 * i] = t.cu; o[3 * i + 1] = t.tx; o[3 * i + 2] = t.ty; } break; }
    case 3: memcpy(out, g_geo.cu_jobs.data(), sizeof(int32_t) * g_geo.cu_jobs.size()); break;
    case 4: memcpy(out, g_geo.rowJob.data(), sizeof(int) * g_geo.rowJob.size()); break;
    case 5: memcpy(out, g_geo.rowCu.data(), sizeof(int) * g_geo.rowCu.size()); break;
    case 6: memcpy(out, g_geo.rowTu.data(), sizeof(int) * g_geo.rowTu.size()); break;
    /* job groups of the shared-memory-window search, class k = 0 (CU 64), 1 (CU 32) and 2 (16x16 cells):
     * 10 + 4k: counts {ngroups, njobs}; 11 + 4k: first|count pairs; 12 + 4k: job indices; 13 + 4k: rowGrp */
    case 10: case 14: case 18: { int k = (what - 10) / 4; ((int32_t*)out)[0] = (int32_t)g_geo.grpFirst[k].size(); ((int32_t*)out)[1] = (int32_t)g_geo.grpJobs[k].size(); break; }
    case 11: case 15: case 19: { int k = (what - 11) / 4; int32_t* o = (int32_t*)out; for (size_t i = 0; i < g_geo.grpFirst[k].size(); i++) { o[2 * i] = g_geo.grpFirst[k][i]; o[2 * i + 1] = g_geo.grpCount[k][i]; } break; }
    case 12: case 16: case 20: { int k = (what - 12) / 4; memcpy(out, g_geo.grpJobs[k].data(), sizeof(int32_t) * g_geo.grpJobs[k].size()); break; }
    case 13: case 17: case 21: { int k = (what - 13) / 4; memcpy(out, g_geo.rowGrp[k].data(), sizeof(int) * g_geo.rowGrp[k].size()); break; }
    case 22: memcpy(out, g_geo.order.data(), sizeof(int32_t) * g_geo.order.size()); break;   /* shape-sorted job order per CTU row */
    default: return -1;
    }
    return 0;
}

/* layout of the ABI structs the Python binding mirrors: sizeof / offsetof pairs */
int geo_abi_layout(int* out)
{
    out[0] = (int)sizeof(x265cu_analysis_params); out[1] = (int)offsetof(x265cu_analysis_params, qp);
    out[2] = (int)offsetof(x265cu_analysis_params, lambda); out[3] = (int)offsetof(x265cu_analysis_params, amp);
    out[4] = (int)sizeof(x265cu_me_job); out[5] = (int)sizeof(x265cu_me_chroma); out[6] = (int)offsetof(x265cu_me_chroma, cstride);
    out[7] = (int)sizeof(x265cu_analysis_out); out[8] = (int)sizeof(PuDesc); out[9] = (int)sizeof(CuDesc); out[10] = (int)sizeof(TuDesc);
    return 11;
}

} // extern "C"

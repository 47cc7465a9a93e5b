// x265_b200/csrc/geometry.h -- host-only (no CUDA): the static geometry tables of the frame analyser.
// PU / CU / TU lists of a picture in canonical order (CTU raster; per CTU and reference the CUs 64..8 in raster order, each
// with its 2Nx2N PU, then -- rect -- 2NxN x2, Nx2N x2, then -- amp, CU >= 16 -- 2NxnU, 2NxnD, nLx2N, nRx2N; preset slow has rect
// on and AMP off, slower / veryslow have both: common/param.cpp:478-520).  A CU must lie fully inside the picture
// (analysis.cpp only visits CUs inside it); no partition of a CU that crosses the edge is emitted.  Kept free of CUDA so that tests/test_geometry.py can check it on the CPU
// against the oracle's independent enumeration (oracle/frame_spec.h) at every BASELINE picture size.
#pragma once
#include <stdint.h>
#include <vector>
#include <utility>

struct PuDesc { int32_t offset; int16_t cuX, cuY; int8_t pw, ph; int16_t ref; };   // static per geometry
struct CuDesc { int16_t x, y, size, pad; int64_t coef_off; };
struct TuDesc { int32_t cu; int16_t tx, ty; };                                        // TU origin inside the CU

struct FrameGeometry
{
    std::vector<PuDesc> pus; std::vector<CuDesc> cus; std::vector<TuDesc> tus;
    std::vector<int32_t> cu_jobs;                 // [cu][ref] -> index of the CU's 2Nx2N PU job
    // first PU job / CU / TU of every CTU row (+ one past the end): the lists are in CTU raster order, so a CTU-row
    // range [r0, r1) is one contiguous slice of each list (what a WPP row shard owns, frameencoder.cpp:850-868)
    int ctuRows; std::vector<int> rowJob, rowCu, rowTu;
    int64_t ncoef;
    // job groups of the shared-memory-window integer search (me_window.cuh): per (CTU, ref) the 64x64 CU's PUs, each 32x32
    // CU's PUs, and per 16x16 cell the PUs of the 16x16 CU and of its four 8x8 CUs.  grpJobs holds absolute job indices,
    // largest PU first inside a group; groups are in CTU raster order (rowGrp = first group of every CTU row + end).
    // Three classes, launched separately (different shared-memory budgets): [0] = CU 64 groups, [1] = CU 32 groups, [2] = 16x16 cells.
    std::vector<int32_t> grpFirst[3], grpCount[3], grpJobs[3]; std::vector<int> rowGrp[3];
    // the jobs of every CTU row sorted by PU shape (largest first; absolute job indices; row r = order[rowJob[r] .. rowJob[r + 1])):
    // the pre-check and sub-pel launches hand out jobs in this order so that all warps run the same code paths together
    std::vector<int32_t> order;
};

// partitions of a CU of `size`: returns the count and fills part[k] = {x, y, w, h} relative to the CU origin
static inline int geometry_cu_parts(int size, int rect, int amp, int part[13][4])
{
    const int h2 = size / 2, q = size / 4;
    int n = 0;
#define GEO_PART(x, y, w, h) do { part[n][0] = (x); part[n][1] = (y); part[n][2] = (w); part[n][3] = (h); n++; } while (0)
    GEO_PART(0, 0, size, size);
    if (rect)
    {
        GEO_PART(0, 0, size, h2); GEO_PART(0, h2, size, h2);
        GEO_PART(0, 0, h2, size); GEO_PART(h2, 0, h2, size);
    }
    if (amp && size >= 16)
    {   // 2NxnU, 2NxnD, nLx2N, nRx2N (AMP is not allowed at the minimum CU size)
        GEO_PART(0, 0, size, q);          GEO_PART(0, q, size, size - q);
        GEO_PART(0, 0, size, size - q);   GEO_PART(0, size - q, size, q);
        GEO_PART(0, 0, q, size);          GEO_PART(q, 0, size - q, size);
        GEO_PART(0, 0, size - q, size);   GEO_PART(size - q, 0, q, size);
    }
#undef GEO_PART
    return n;
}

static inline void geometry_build(int W, int H, int stride, int nref, int rect, int amp, FrameGeometry& g)
{
    // CU list: CTU raster, sizes 64..8, raster inside the CTU; CUs must lie fully inside the picture
    const int ctuW = (W + 63) / 64, ctuH = (H + 63) / 64;
    int64_t coefOff = 0;
    g.ctuRows = ctuH;
    for (int cty = 0; cty < ctuH; cty++)
    {
        g.rowJob.push_back((int)g.pus.size()); g.rowCu.push_back((int)g.cus.size()); g.rowTu.push_back((int)g.tus.size());
        for (int k = 0; k < 3; k++) g.rowGrp[k].push_back((int)g.grpFirst[k].size());
        for (int ctx = 0; ctx < ctuW; ctx++)
        {
            int local[85]; int nl = 0;
            for (int size = 64; size >= 8; size >>= 1)
                for (int cy = 0; cy < 64; cy += size)
                    for (int cx = 0; cx < 64; cx += size, nl++)
                    {
                        const int x = ctx * 64 + cx, y = cty * 64 + cy;
                        local[nl] = -1;
                        if (x + size > W || y + size > H) continue;
                        local[nl] = (int)g.cus.size();
                        CuDesc c; c.x = (int16_t)x; c.y = (int16_t)y; c.size = (int16_t)size; c.pad = 0; c.coef_off = coefOff;
                        coefOff += (int64_t)size * size;
                        const int T = size > 32 ? 32 : size;
                        for (int ty = 0; ty < size; ty += T)
                            for (int tx = 0; tx < size; tx += T)
                            {
                                TuDesc t; t.cu = (int32_t)g.cus.size(); t.tx = (int16_t)tx; t.ty = (int16_t)ty;
                                g.tus.push_back(t);
                            }
                        g.cus.push_back(c);
                    }
            g.cu_jobs.resize(g.cus.size() * nref, -1);
            // PU jobs of this CTU: per ref, per CU (same order), the CU's partitions in geometry_cu_parts order
            for (int r = 0; r < nref; r++)
            {
                int li = 0;
                int jobStart[85], jobCount[85];
                for (int q = 0; q < 85; q++) { jobStart[q] = 0; jobCount[q] = 0; }
                for (int size = 64; size >= 8; size >>= 1)
                    for (int cy = 0; cy < 64; cy += size)
                        for (int cx = 0; cx < 64; cx += size, li++)
                        {
                            if (local[li] < 0) continue;          // CU crosses the picture edge: no partition of it is evaluated
                            int part[13][4];
                            const int np = geometry_cu_parts(size, rect, amp, part);
                            jobStart[li] = (int)g.pus.size(); jobCount[li] = np;
                            for (int k = 0; k < np; k++)
                            {
                                const int x = ctx * 64 + cx + part[k][0], y = cty * 64 + cy + part[k][1], w = part[k][2], h = part[k][3];
                                PuDesc d; d.offset = y * stride + x; d.cuX = (int16_t)(ctx * 64 + cx); d.cuY = (int16_t)(cty * 64 + cy);
                                d.pw = (int8_t)w; d.ph = (int8_t)h; d.ref = (int16_t)r;
                                if (k == 0) g.cu_jobs[(size_t)local[li] * nref + r] = (int32_t)g.pus.size();
                                g.pus.push_back(d);
                            }
                        }
                // groups of this (CTU, ref): CU 64, CUs 32, then the 16x16 cells (CU 16 + its four CUs 8)
                for (int gi = 0; gi < 21; gi++)
                {
                    int members[5], nm = 0;
                    members[nm++] = gi;
                    if (gi >= 5)
                    {
                        const int k = gi - 5, cy16 = k >> 2, cx16 = k & 3;
                        for (int a8 = 0; a8 < 2; a8++)
                            for (int b8 = 0; b8 < 2; b8++) members[nm++] = 21 + (2 * cy16 + a8) * 8 + (2 * cx16 + b8);
                    }
                    const int cl = gi == 0 ? 0 : (gi < 5 ? 1 : 2);
                    std::vector<int32_t>& gj = g.grpJobs[cl];
                    const int first = (int)gj.size();
                    for (int m = 0; m < nm; m++)
                        for (int q = 0; q < jobCount[members[m]]; q++) gj.push_back(jobStart[members[m]] + q);
                    const int count = (int)gj.size() - first;
                    if (!count) continue;
                    // largest PU first (stable insertion sort: groups hold at most 33 jobs)
                    for (int a = first + 1; a < first + count; a++)
                    {
                        const int32_t v = gj[a];
                        const int av = g.pus[v].pw * g.pus[v].ph;
                        int b = a - 1;
                        while (b >= first && g.pus[gj[b]].pw * g.pus[gj[b]].ph < av) { gj[b + 1] = gj[b]; b--; }
                        gj[b + 1] = v;
                    }
                    g.grpFirst[cl].push_back(first); g.grpCount[cl].push_back(count);
                }
            }
        }
    }
    for (int k = 0; k < 3; k++) g.rowGrp[k].push_back((int)g.grpFirst[k].size());
    g.rowJob.push_back((int)g.pus.size()); g.rowCu.push_back((int)g.cus.size()); g.rowTu.push_back((int)g.tus.size());
    g.ncoef = coefOff;
    // shape-sorted job order per CTU row: counting sort on (w, h) keys, stable (area descending, then width descending)
    g.order.resize(g.pus.size());
    for (int r = 0; r < ctuH; r++)
    {
        const int j0 = g.rowJob[r], j1 = g.rowJob[r + 1];
        // key = (area, w): collect distinct shapes
        std::vector<std::pair<int, int> > shapes;
        for (int j = j0; j < j1; j++)
        {
            const std::pair<int, int> sh(g.pus[j].pw * g.pus[j].ph, g.pus[j].pw);
            bool seen = false;
            for (size_t q = 0; q < shapes.size(); q++) if (shapes[q] == sh) { seen = true; break; }
            if (!seen) shapes.push_back(sh);
        }
        for (size_t a = 1; a < shapes.size(); a++)          // insertion sort, descending
        {
            const std::pair<int, int> v = shapes[a]; size_t b = a;
            while (b > 0 && shapes[b - 1] < v) { shapes[b] = shapes[b - 1]; b--; }
            shapes[b] = v;
        }
        int pos = j0;
        for (size_t q = 0; q < shapes.size(); q++)
            for (int j = j0; j < j1; j++)
                if (g.pus[j].pw * g.pus[j].ph == shapes[q].first && g.pus[j].pw == shapes[q].second) g.order[pos++] = j;
    }
}

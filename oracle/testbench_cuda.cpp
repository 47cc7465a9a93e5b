/* oracle/testbench_cuda.cpp -- TEST INFRASTRUCTURE ONLY.
 * The reference's own parity gate, driven against our table: builds the C table exactly as
 * source/test/testbench.cpp:155-158 does, builds the CUDA table with setupCudaPrimitives() +
 * setupAliasPrimitives(), and runs the four reference harnesses' testCorrectness(cprim, cudaprim)
 * (PixelHarness, MBDstHarness, IPFilterHarness, IntraPredHarness; source/test/*harness.cpp, compiled
 * unmodified from /root/reference into oracle/_ref/libx265harness{8,10}.a).  A field is tested iff the
 * CUDA table has it (pixelharness.cpp:2325-2359).  Exit code 0 = every check passed.
 */
#include "common.h"
#include "primitives.h"
#include "pixelharness.h"
#include "mbdstharness.h"
#include "ipfilterharness.h"
#include "intrapredharness.h"
#include "x265_b200.h"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

using namespace X265_NS;
namespace X265_NS {
void setupCPrimitives(EncoderPrimitives& p);
void setupAliasPrimitives(EncoderPrimitives& p);
void setupCudaPrimitives(EncoderPrimitives& p, int cpuMask);
}
/* the harness sources refer to these tables (testbench.cpp:33-73); content is only used in messages */
const char* lumaPartStr[NUM_PU_SIZES] = { "4x4", "8x8", "16x16", "32x32", "64x64", "8x4", "4x8", "16x8", "8x16", "32x16", "16x32", "64x32", "32x64",
    "16x12", "12x16", "16x4", "4x16", "32x24", "24x32", "32x8", "8x32", "64x48", "48x64", "64x16", "16x64" };
static const char* c420[NUM_PU_SIZES] = { "2x2", "4x4", "8x8", "16x16", "32x32", "4x2", "2x4", "8x4", "4x8", "16x8", "8x16", "32x16", "16x32",
    "8x6", "6x8", "8x2", "2x8", "16x12", "12x16", "16x4", "4x16", "32x24", "24x32", "32x8", "8x32" };
static const char* c422[NUM_PU_SIZES] = { "2x4", "4x8", "8x16", "16x32", "32x64", "4x4", "2x8", "8x8", "4x16", "16x16", "8x32", "32x32", "16x64",
    "8x12", "6x16", "8x4", "2x16", "16x24", "12x32", "16x8", "4x32", "32x48", "24x64", "32x16", "8x64" };
const char* const* chromaPartStr[X265_CSP_COUNT] = { lumaPartStr, c420, c422, lumaPartStr };

int main(int argc, char** argv)
{
    int seed = argc > 1 ? atoi(argv[1]) : 265;
    printf("x265 TestBench harnesses vs CUDA table, seed %d, %d-bit, CUDA devices %d\n", seed, X265_DEPTH, x265cu_device_count());
    if (x265cu_device_count() <= 0) { fprintf(stderr, "no CUDA device: nothing to test (no CPU fallback)\n"); return 2; }
    srand(seed);
    PixelHarness hPixel; MBDstHarness hMBDist; IPFilterHarness hIPFilter; IntraPredHarness hIPred;
    TestHarness* harness[] = { &hPixel, &hMBDist, &hIPFilter, &hIPred };
    EncoderPrimitives cprim, cuprim;
    memset(&cprim, 0, sizeof(cprim)); memset(&cuprim, 0, sizeof(cuprim));
    setupCPrimitives(cprim); setupAliasPrimitives(cprim);
    /* start from an all-NULL table so that only CUDA-backed fields are exercised; the chroma interpolation
     * entries are installed only where the C table has them, so seed those slots first */
    for (int i = 0; i < NUM_PU_SIZES; i++) cuprim.chroma[X265_CSP_I420].pu[i] = cprim.chroma[X265_CSP_I420].pu[i];
    setupCudaPrimitives(cuprim, 0);
    for (int i = 0; i < NUM_PU_SIZES; i++)
    {   /* drop the seeded C pointers that the CUDA setup did not replace */
        EncoderPrimitives::Chroma::PUChroma& c = cuprim.chroma[X265_CSP_I420].pu[i];
        const EncoderPrimitives::Chroma::PUChroma& r = cprim.chroma[X265_CSP_I420].pu[i];
        if (c.filter_hpp == r.filter_hpp) c.filter_hpp = NULL;
        if (c.filter_hps == r.filter_hps) c.filter_hps = NULL;
        if (c.filter_vpp == r.filter_vpp) c.filter_vpp = NULL;
        if (c.filter_vps == r.filter_vps) c.filter_vps = NULL;
        if (c.filter_vsp == r.filter_vsp) c.filter_vsp = NULL;
        if (c.filter_vss == r.filter_vss) c.filter_vss = NULL;
        c.satd = NULL; c.addAvg[0] = c.addAvg[1] = NULL; c.copy_pp = NULL; c.p2s[0] = c.p2s[1] = NULL;
    }
    setupAliasPrimitives(cuprim);
    memcpy(&primitives, &cprim, sizeof(EncoderPrimitives));   /* hybrid C helpers use the global table */
    int nset = 0; void** raw = (void**)&cuprim;
    for (size_t i = 0; i < sizeof(cuprim) / sizeof(void*); i++) nset += raw[i] != NULL;
    printf("CUDA table: %d non-NULL entries (post-alias)\n", nset);
    for (size_t h = 0; h < sizeof(harness) / sizeof(harness[0]); h++)
    {
        printf("== %s\n", harness[h]->getName()); fflush(stdout);
        if (!harness[h]->testCorrectness(cprim, cuprim))
        {
            fflush(stdout);
            fprintf(stderr, "\nx265: CUDA primitive has failed in harness %s\n", harness[h]->getName());
            return 1;
        }
    }
    printf("\nall reference harnesses passed against the CUDA table\n");
    return 0;
}

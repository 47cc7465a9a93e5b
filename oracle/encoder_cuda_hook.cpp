/* oracle/encoder_cuda_hook.cpp -- TEST INFRASTRUCTURE ONLY.
 * Links the UNMODIFIED reference encoder + CLI (oracle/Makefile.ref) with the CUDA primitive table at the position
 * INTEGRATION.md names: common/primitives.cpp is compiled for this binary with -DENABLE_ASSEMBLY=1, so that
 * x265_setup_primitives() (primitives.cpp:248-285) calls setupInstrinsicPrimitives() and setupAssemblyPrimitives()
 * between setupCPrimitives() and setupAliasPrimitives() -- and this file IS those two functions:
 *   X265_PRIMITIVES=cuda   -> setupCudaPrimitives(p, mask)  (x265_b200/plugin/setup_cuda_primitives.cpp over libx265cu.so)
 *   anything else          -> nothing: the C table stays (= x265 --no-asm)
 * The encoder above the table (x265_encoder_* C ABI, x265.h:2070-2129) is the reference's own code, untouched.
 * At exit the process reports how many per-call primitives ran on the device and whether any failed. */
#include "common.h"
#include "primitives.h"
#include "x265_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace X265_NS {
void setupCudaPrimitives(EncoderPrimitives& p, int cpuMask);

static void report_at_exit()
{
    fprintf(stderr, "x265cu: per-call primitives executed on the device: %llu, error flag: %d %s\n",
            (unsigned long long)x265cu_primitive_calls(), x265cu_primitive_error(), x265cu_primitive_error_string());
}

void setupInstrinsicPrimitives(EncoderPrimitives&, int) {}

void setupAssemblyPrimitives(EncoderPrimitives& p, int cpuMask)
{
    const char* sel = getenv("X265_PRIMITIVES");
    if (sel && !strcmp(sel, "cuda"))
    {
        if (x265cu_device_count() <= 0)
        {
            fprintf(stderr, "x265cu: X265_PRIMITIVES=cuda but no CUDA device: refusing to run (no CPU fallback on the CUDA path)\n");
            exit(3);
        }
        setupCudaPrimitives(p, cpuMask);
        atexit(report_at_exit);
        fprintf(stderr, "x265cu: EncoderPrimitives table = CUDA (setupCudaPrimitives)\n");
    }
}
}

/* primitives.cpp leaves these to the assembly when ENABLE_ASSEMBLY is defined (primitives.cpp:288-303) */
extern "C" {
int PFX(cpu_cpuid_test)(void) { return 0; }
void PFX(cpu_emms)(void) {}
void PFX(cpu_cpuid)(uint32_t, uint32_t* eax, uint32_t*, uint32_t*, uint32_t*) { *eax = 0; }
void PFX(cpu_xgetbv)(uint32_t, uint32_t*, uint32_t*) {}
void PFX(cpu_neon_test)(void) {}
int PFX(cpu_fast_neon_mrc_test)(void) { return 0; }
}

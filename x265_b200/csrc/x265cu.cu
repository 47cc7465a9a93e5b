// x265_b200/csrc/x265cu.cu -- the single translation unit of libx265cu.so: context, memory,
// batched C-ABI entry points (include/x265_b200.h) and the per-call primitive table.
#include "common.cuh"
#include "pixelcmp.cuh"
#include "pixelcmp_grid.cuh"
#include "blockops.cuh"
#include "interp.cuh"
#include "transform.cuh"
#include "intra.cuh"
#include "me.cuh"
#include "predcost.cuh"
#include "frame.cuh"
#include "lookahead.cuh"
#include <math.h>
#include <mutex>

static thread_local char g_err[512] = "";
unsigned long long g_x265cu_launches = 0;
void x265cu_set_error(const char* what, cudaError_t e, const char* file, int line)
{
    snprintf(g_err, sizeof(g_err), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    fprintf(stderr, "x265cu: %s\n", g_err);
}

extern "C" {

const char* x265cu_last_error(void) { return g_err; }

int x265cu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static void ctx_release(x265cu_ctx* c)
{
    if (c->h_stage) cudaFreeHost(c->h_stage);
    cudaFree(c->d_stage); cudaFree(c->d_counter); cudaFree(c->d_me_state);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 4; i++) if (c->me_ev[i]) cudaEventDestroy(c->me_ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

x265cu_ctx* x265cu_create(int device)
{
    if (device < 0 || device >= x265cu_device_count())
    {
        snprintf(g_err, sizeof(g_err), "no CUDA device %d (libx265cu has no CPU fallback)", device);
        return NULL;
    }
    if (cudaSetDevice(device) != cudaSuccess) return NULL;
    x265cu_ctx* c = new x265cu_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->sm_count = prop.multiProcessorCount;
    c->stage_bytes = 4u << 20;
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess;
    for (int i = 0; i < 4 && ok; i++) ok = cudaEventCreate(&c->me_ev[i]) == cudaSuccess;
    ok = ok && build_dct_tables(device) == 0;
    ok = ok && cudaMallocHost((void**)&c->h_stage, c->stage_bytes) == cudaSuccess &&
         cudaMalloc((void**)&c->d_stage, c->stage_bytes) == cudaSuccess &&
         cudaMalloc((void**)&c->d_counter, 64) == cudaSuccess;
    if (!ok)
    {
        x265cu_set_error("ctx create", cudaGetLastError(), __FILE__, __LINE__);
        ctx_release(c);
        return NULL;
    }
    return c;
}

void x265cu_destroy(x265cu_ctx* c)
{
    cudaSetDevice(c->device);
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    ctx_release(c);
}

int x265cu_sync(x265cu_ctx* c) { cudaSetDevice(c->device); CU_CHECK(cudaStreamSynchronize(c->stream)); return 0; }
void* x265cu_stream(x265cu_ctx* c) { cudaSetDevice(c->device); return (void*)c->stream; }
void* x265cu_malloc(x265cu_ctx* c, size_t bytes)
{
    void* p = NULL;
    cudaSetDevice(c->device);
    if (cudaMalloc(&p, bytes) != cudaSuccess) { x265cu_set_error("cudaMalloc", cudaGetLastError(), __FILE__, __LINE__); return NULL; }
    return p;
}
void x265cu_free(x265cu_ctx* c, void* dev) { cudaSetDevice(c->device); cudaFree(dev); }
void* x265cu_host_alloc(size_t bytes) { void* p = NULL; if (cudaMallocHost(&p, bytes) != cudaSuccess) return NULL; return p; }
void x265cu_host_free(void* p) { cudaFreeHost(p); }
int x265cu_h2d(x265cu_ctx* c, void* dev, const void* host, size_t bytes) { cudaSetDevice(c->device); CU_CHECK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c->stream)); return 0; }
int x265cu_d2h(x265cu_ctx* c, void* host, const void* dev, size_t bytes) { cudaSetDevice(c->device); CU_CHECK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream)); return 0; }
int x265cu_copy2d(x265cu_ctx* c, void* dst, size_t dpitch, const void* src, size_t spitch, size_t widthBytes, size_t rows, int kind)
{
    cudaSetDevice(c->device);
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CU_CHECK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, widthBytes, rows, k, c->stream));
    return 0;
}
int x265cu_memset(x265cu_ctx* c, void* dev, int value, size_t bytes) { cudaSetDevice(c->device); CU_CHECK(cudaMemsetAsync(dev, value, bytes, c->stream)); return 0; }
int x265cu_timer_begin(x265cu_ctx* c) { cudaSetDevice(c->device); CU_CHECK(cudaEventRecord(c->ev0, c->stream)); return 0; }
float x265cu_timer_end(x265cu_ctx* c)
{
    cudaSetDevice(c->device);
    float ms = -1.f;
    if (cudaEventRecord(c->ev1, c->stream) != cudaSuccess) return -1.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.f;
    return ms;
}
uint64_t x265cu_launch_count(x265cu_ctx* c) { (void)c; return __atomic_load_n(&g_x265cu_launches, __ATOMIC_RELAXED); }
int x265cu_me_phase_ms(x265cu_ctx* c, float* ms)
{
    cudaSetDevice(c->device);
    CU_CHECK(cudaEventSynchronize(c->me_ev[3]));
    for (int i = 0; i < 3; i++) CU_CHECK(cudaEventElapsedTime(&ms[i], c->me_ev[i], c->me_ev[i + 1]));
    return 0;
}

// ---------------- batched API ----------------
int x265cu_pixelcmp_batch(x265cu_ctx* c, int depth, int op, const void* A, const void* B, const x265cu_cmp_job* jobs, int n, uint64_t* out)
{ cudaSetDevice(c->device); return launch_pixelcmp(c, depth, op, A, B, jobs, n, out); }
int x265cu_pixelcmp_grid(x265cu_ctx* c, int depth, int op, const void* A, int64_t a_stride, const void* B, int64_t b_stride,
                         int bw, int bh, int nbx, int nby, uint64_t* out)
{ cudaSetDevice(c->device); return launch_pixelcmp_grid(c, depth, op, A, a_stride, B, b_stride, bw, bh, nbx, nby, out); }

int x265cu_blockop_batch(x265cu_ctx* c, int depth, int op, void* D, const void* A, const void* B, const x265cu_blk_job* jobs, int n)
{ cudaSetDevice(c->device); return launch_blockop(c, depth, op, D, A, B, jobs, n); }

int x265cu_interp_batch(x265cu_ctx* c, int depth, int op, const void* src, void* dst, const x265cu_interp_job* jobs, int n)
{ cudaSetDevice(c->device); return launch_interp(c, depth, op, src, dst, jobs, n); }

int x265cu_transform_batch(x265cu_ctx* c, int depth, int op, int size, const int16_t* src, int16_t* dst, int stride, int64_t tu_pitch, int n)
{ cudaSetDevice(c->device); return launch_transform(c, depth, op, size, src, dst, stride, tu_pitch, n); }

int x265cu_quant_batch(x265cu_ctx* c, const int16_t* coef, const int32_t* qc, int32_t* deltaU, int16_t* qCoef, int qBits, int add,
                       int numCoeff, int n, int nquant, uint32_t* numSig)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    int blocks = (n + 7) / 8; if (blocks > c->sm_count * 8) blocks = c->sm_count * 8;
    k_quant<<<blocks, 256, 0, c->stream>>>(coef, qc, deltaU, qCoef, qBits, add, numCoeff, n, nquant, numSig);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_dequant_normal_batch(x265cu_ctx* c, const int16_t* q, int16_t* coef, int64_t num, int scale, int shift)
{
    cudaSetDevice(c->device);
    if (num <= 0) return 0;
    int64_t blocks = (num + 255) / 256; if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    k_dequant_normal<<<(int)blocks, 256, 0, c->stream>>>(q, coef, num, scale, shift);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_dequant_scaling_batch(x265cu_ctx* c, const int16_t* q, const int32_t* dq, int16_t* coef, int numCoeff, int n, int per, int shift)
{
    cudaSetDevice(c->device);
    int64_t total = (int64_t)numCoeff * n;
    if (total <= 0) return 0;
    int64_t blocks = (total + 255) / 256; if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    k_dequant_scaling<<<(int)blocks, 256, 0, c->stream>>>(q, dq, coef, numCoeff, total, per, shift);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_intra_pred_batch(x265cu_ctx* c, int depth, int size, const void* nb, int64_t nb_pitch, void* dst, int64_t dst_pitch, int dst_stride,
                            const x265cu_intra_job* jobs, int n)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    int blocks = n < c->sm_count * 8 ? n : c->sm_count * 8;
    int threads = size * size < 256 ? (size * size < 32 ? 32 : size * size) : 256;
    if (depth == 8) k_intra_pred<uint8_t><<<blocks, threads, 0, c->stream>>>(size, (const uint8_t*)nb, nb_pitch, (uint8_t*)dst, dst_pitch, dst_stride, jobs, n);
    else            k_intra_pred<uint16_t><<<blocks, threads, 0, c->stream>>>(size, (const uint16_t*)nb, nb_pitch, (uint16_t*)dst, dst_pitch, dst_stride, jobs, n);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_intra_filter_batch(x265cu_ctx* c, int depth, int size, const void* nb, void* filt, int64_t pitch, int n)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    int blocks = (n + 7) / 8 < c->sm_count * 8 ? (n + 7) / 8 : c->sm_count * 8;       // a warp per neighbour array
    if (depth == 8) k_intra_filter<uint8_t><<<blocks, 256, 0, c->stream>>>(size, (const uint8_t*)nb, (uint8_t*)filt, pitch, n);
    else            k_intra_filter<uint16_t><<<blocks, 256, 0, c->stream>>>(size, (const uint16_t*)nb, (uint16_t*)filt, pitch, n);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_intra_allangs_batch(x265cu_ctx* c, int depth, int size, const void* ref, const void* filt, int64_t nb_pitch, void* dst, int bLuma, int n)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    if ((((uintptr_t)dst) & 7) == 0 && size >= 4 && size <= 32)
    {   // vector-store kernel: one CTA per block, all 33 modes
        int blocks = n < c->sm_count * 16 ? n : c->sm_count * 16;
        if (!((uintptr_t)dst & 15))
        {   // packed-pixel kernel (16-byte row pieces): words per item = min(row bytes, 16) / 4
            const int rowBytes = size * (depth == 8 ? 1 : 2), uw = (rowBytes < 16 ? rowBytes : 16) / 4;
            blocks = n < c->sm_count * 8 ? n : c->sm_count * 8;
            if (depth == 8)
            {
                if (uw == 4)      k_intra_allangs_packed<uint8_t, 4><<<blocks, 256, 0, c->stream>>>(size, (const uint8_t*)ref, (const uint8_t*)filt, nb_pitch, (uint8_t*)dst, bLuma, n);
                else if (uw == 2) k_intra_allangs_packed<uint8_t, 2><<<blocks, 256, 0, c->stream>>>(size, (const uint8_t*)ref, (const uint8_t*)filt, nb_pitch, (uint8_t*)dst, bLuma, n);
                else              k_intra_allangs_packed<uint8_t, 1><<<blocks, 256, 0, c->stream>>>(size, (const uint8_t*)ref, (const uint8_t*)filt, nb_pitch, (uint8_t*)dst, bLuma, n);
            }
            else
            {
                if (uw == 4)      k_intra_allangs_packed<uint16_t, 4><<<blocks, 256, 0, c->stream>>>(size, (const uint16_t*)ref, (const uint16_t*)filt, nb_pitch, (uint16_t*)dst, bLuma, n);
                else              k_intra_allangs_packed<uint16_t, 2><<<blocks, 256, 0, c->stream>>>(size, (const uint16_t*)ref, (const uint16_t*)filt, nb_pitch, (uint16_t*)dst, bLuma, n);
            }
            CU_LAUNCH_CHECK(c);
            return 0;
        }
        if (depth == 8) k_intra_allangs_cta<uint8_t><<<blocks, 256, 0, c->stream>>>(size, (const uint8_t*)ref, (const uint8_t*)filt, nb_pitch, (uint8_t*)dst, bLuma, n);
        else            k_intra_allangs_cta<uint16_t><<<blocks, 256, 0, c->stream>>>(size, (const uint16_t*)ref, (const uint16_t*)filt, nb_pitch, (uint16_t*)dst, bLuma, n);
        CU_LAUNCH_CHECK(c);
        return 0;
    }
    dim3 grid(33, n < 4096 ? n : 4096);
    int threads = size * size < 256 ? (size * size < 32 ? 32 : size * size) : 256;
    if (depth == 8) k_intra_allangs<uint8_t><<<grid, threads, 0, c->stream>>>(size, (const uint8_t*)ref, (const uint8_t*)filt, nb_pitch, (uint8_t*)dst, bLuma, n);
    else            k_intra_allangs<uint16_t><<<grid, threads, 0, c->stream>>>(size, (const uint16_t*)ref, (const uint16_t*)filt, nb_pitch, (uint16_t*)dst, bLuma, n);
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_extend_border(x265cu_ctx* c, int depth, void* plane, int stride, int width, int height, int mx, int my)
{
    cudaSetDevice(c->device);
    if (depth == 8) return extend_border_t<uint8_t>(c, (uint8_t*)plane, stride, width, height, mx, my);
    return extend_border_t<uint16_t>(c, (uint16_t*)plane, stride, width, height, mx, my);
}

int x265cu_frame_init_lowres(x265cu_ctx* c, int depth, const void* src, int sstride, void* d0, void* dh, void* dv, void* dc,
                             int dstride, int width, int height, int mx, int my)
{
    cudaSetDevice(c->device);
    dim3 block(64, 4), grid((width / 4 + 63) / 64 + 1, (height + 3) / 4);
    const uintptr_t dal = (uintptr_t)d0 | (uintptr_t)dh | (uintptr_t)dv | (uintptr_t)dc | (uintptr_t)dstride;
    if (depth == 8 && (width & 7) == 0 && (((uintptr_t)src | (uintptr_t)sstride) & 15) == 0 && (dal & 7) == 0)
    {
        dim3 g8((width / 8 + 63) / 64, (height + 3) / 4);
        k_lowres_init_u8x8<<<g8, block, 0, c->stream>>>((const uint8_t*)src, sstride, (uint8_t*)d0, (uint8_t*)dh, (uint8_t*)dv, (uint8_t*)dc, dstride, width / 8, height);
    }
    else if (depth != 8 && (width & 3) == 0 && (((uintptr_t)src | (uintptr_t)(sstride * 2)) & 15) == 0 && (((uintptr_t)d0 | (uintptr_t)dh | (uintptr_t)dv | (uintptr_t)dc | (uintptr_t)(dstride * 2)) & 7) == 0)
    {
        dim3 g4((width / 4 + 63) / 64, (height + 3) / 4);
        k_lowres_init_u16x4<<<g4, block, 0, c->stream>>>((const uint16_t*)src, sstride, (uint16_t*)d0, (uint16_t*)dh, (uint16_t*)dv, (uint16_t*)dc, dstride, width / 4, height);
    }
    else if (depth == 8)
        k_lowres_init<uint8_t><<<grid, block, 0, c->stream>>>((const uint8_t*)src, sstride, (uint8_t*)d0, (uint8_t*)dh, (uint8_t*)dv, (uint8_t*)dc, dstride, width, height);
    else
        k_lowres_init<uint16_t><<<grid, block, 0, c->stream>>>((const uint16_t*)src, sstride, (uint16_t*)d0, (uint16_t*)dh, (uint16_t*)dv, (uint16_t*)dc, dstride, width, height);
    CU_LAUNCH_CHECK(c);
    if (mx > 0 || my > 0)
    {
        void* planes[4] = { d0, dh, dv, dc };                    // all four planes in one launch
        if (depth == 8) { if (extend_border_n<uint8_t>(c, (uint8_t* const*)planes, 4, dstride, width, height, mx, my)) return -1; }
        else            { if (extend_border_n<uint16_t>(c, (uint16_t* const*)planes, 4, dstride, width, height, mx, my)) return -1; }
    }
    return 0;
}

// bitcost.cpp:95-109 + :49-55 (host helper; float log, double lambda, cap 2^15-1)
void x265cu_mvcost_table(double lambda, int range, uint16_t* out)
{
    float log2_2 = 2.0f / logf(2.0f);
    for (int i = 0; i <= range; i++)
    {
        float bits = i ? logf((float)(i + 1)) * log2_2 + 1.718f : 0.718f;
        double v = bits * lambda + 0.5f;
        if (v > 32767.0) v = 32767.0;
        out[range + i] = out[range - i] = (uint16_t)v;
    }
}

int x265cu_me_batch(x265cu_ctx* c, int depth, const void* fenc, int fencStride, const void* const* refs, int refStride, int lowres,
                    const uint16_t* mvcost, int mvcost_range, const x265cu_me_job* jobs, int n, int32_t* out)
{
    cudaSetDevice(c->device);
    (void)mvcost_range;
    return launch_me(c, depth, fenc, fencStride, refs, refStride, lowres, mvcost, jobs, n, out, c->d_counter);
}

int x265cu_me_batch_chroma(x265cu_ctx* c, int depth, const void* fenc, int fencStride, const void* const* refs, int refStride,
                           const x265cu_me_chroma* chroma, const uint16_t* mvcost, int mvcost_range, const x265cu_me_job* jobs, int n, int32_t* out)
{
    cudaSetDevice(c->device);
    (void)mvcost_range;
    if (!chroma || !chroma->fencCb_dev || !chroma->fencCr_dev || !chroma->refCb_dev || !chroma->refCr_dev || chroma->cstride <= 0)
    {
        x265cu_set_error("x265cu_me_batch_chroma: chroma planes missing", cudaErrorInvalidValue, __FILE__, __LINE__);
        return -1;
    }
    MeChromaArgs a;
    a.fencCb = chroma->fencCb_dev; a.fencCr = chroma->fencCr_dev; a.refCb = chroma->refCb_dev; a.refCr = chroma->refCr_dev; a.cstride = chroma->cstride;
    return launch_me(c, depth, fenc, fencStride, refs, refStride, 0, mvcost, jobs, n, out, c->d_counter, &a);
}

int x265cu_pred_cost_batch(x265cu_ctx* c, int depth, const void* fenc, int fencStride, const void* const* refs, int refStride,
                           const x265cu_me_chroma* chroma, const x265cu_pred_job* jobs, int n, int32_t* out)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    PredChroma ch = { NULL, NULL, NULL, NULL, 0 };
    if (chroma && chroma->fencCb_dev && chroma->fencCr_dev && chroma->refCb_dev && chroma->refCr_dev && chroma->cstride > 0)
    {
        ch.fcb = chroma->fencCb_dev; ch.fcr = chroma->fencCr_dev; ch.rcb = chroma->refCb_dev; ch.rcr = chroma->refCr_dev; ch.cstride = chroma->cstride;
    }
    if (depth == 8) return launch_pred_cost_t<uint8_t>(c, fenc, fencStride, refs, refStride, ch, jobs, n, out);
    return launch_pred_cost_t<uint16_t>(c, fenc, fencStride, refs, refStride, ch, jobs, n, out);
}

__global__ void k_la_intra_zero(const x265cu_la_intra_job* jobs, int h8)
{
    const x265cu_la_intra_job jb = jobs[blockIdx.x];
    for (int i = threadIdx.x; i < h8; i += blockDim.x) jb.rowSatds[i] = 0;
    if (threadIdx.x < 2) jb.out[threadIdx.x] = 0;
}

int x265cu_lowres_intra_batch(x265cu_ctx* c, int depth, const x265cu_la_intra_job* jobs, int n, int stride, int w8, int h8, int lambda)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    k_la_intra_zero<<<n, 128, 0, c->stream>>>(jobs, h8);
    CU_LAUNCH_CHECK(c);
    int bx = (w8 * h8 + 31) / 32; if (bx > c->sm_count * 4) bx = c->sm_count * 4;
    dim3 grid(bx, n);
    if (depth == 8) k_lowres_intra<uint8_t><<<grid, 256, 0, c->stream>>>(jobs, stride, w8, h8, lambda);
    else            k_lowres_intra<uint16_t><<<grid, 256, 0, c->stream>>>(jobs, stride, w8, h8, lambda);
    CU_LAUNCH_CHECK(c);
    return 0;
}

// frame totals of every triple of the batch: zeroed once before the launch (the slices of a triple accumulate into them)
__global__ void k_la_zero_out(const x265cu_la_job* jobs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const x265cu_la_job jb = jobs[i];
    if (jb.rows == 0 || (jb.rows & 0xffff) == 0) { jb.out[0] = 0; jb.out[1] = 0; jb.out[2] = 0; }
}

int x265cu_lookahead_cost_batch(x265cu_ctx* c, int depth, const x265cu_la_job* jobs, int n, int stride, int w8, int h8, const uint16_t* mvcost)
{
    cudaSetDevice(c->device);
    if (n <= 0) return 0;
    k_la_zero_out<<<(n + 127) / 128, 128, 0, c->stream>>>(jobs, n);
    CU_LAUNCH_CHECK(c);
    const size_t smem = sizeof(MeShared) * LA_WARPS;
    // cluster size: enough CTAs (of LA_WARPS warps) for the longest anti-diagonal in one round, at most 4
    const int maxdiag = h8 < (w8 + 1) / 2 ? h8 : (w8 + 1) / 2;
    int csize = (maxdiag + LA_WARPS - 1) / LA_WARPS;
    csize = csize >= 4 ? 4 : (csize >= 2 ? 2 : 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)n * csize); cfg.blockDim = dim3(LA_WARPS * 32); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (depth == 8)
    {
        CU_CHECK(cudaFuncSetAttribute(k_lookahead_cost<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_CHECK(cudaLaunchKernelEx(&cfg, k_lookahead_cost<uint8_t>, jobs, stride, w8, h8, mvcost));
    }
    else
    {
        CU_CHECK(cudaFuncSetAttribute(k_lookahead_cost<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_CHECK(cudaLaunchKernelEx(&cfg, k_lookahead_cost<uint16_t>, jobs, stride, w8, h8, mvcost));
    }
    CU_LAUNCH_CHECK(c);
    return 0;
}

int x265cu_propagate_cost_batch(x265cu_ctx* c, int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                                const int32_t* invQscales, double fpsFactor, int64_t len)
{
    cudaSetDevice(c->device);
    if (len <= 0) return 0;
    int64_t blocks = (len + 255) / 256; if (blocks > c->sm_count * 16) blocks = c->sm_count * 16;
    k_propagate_cost<<<(int)blocks, 256, 0, c->stream>>>(dst, propagateIn, intraCosts, interCosts, invQscales, fpsFactor, len);
    CU_LAUNCH_CHECK(c);
    return 0;
}

} // extern "C"

#include "analyser.cuh"
#include "thunks.cuh"
#include "la_weights.cuh"

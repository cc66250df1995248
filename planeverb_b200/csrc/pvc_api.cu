// pvc_api.cu -- the C-ABI entry points declared in include/planeverb_cuda.h: device memory, the
// coefficient-plane voxeliser, the free-field normaliser and the run/fetch calls.  No CPU fallback:
// every entry point fails with PVC_ERR_NO_DEVICE / PVC_ERR_CUDA when the GPU is not usable.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <float.h>
#include "pvc_internal.h"

namespace pvc
{
    static thread_local char g_error[512] = "";

    void setError(const char* fmt, ...)
    {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g_error, sizeof(g_error), fmt, ap);
        va_end(ap);
    }

    #define PVC_CUDA(call)                                                                         \
        do { cudaError_t e_ = (call);                                                               \
             if (e_ != cudaSuccess) { setError("%s: %s", #call, cudaGetErrorString(e_));            \
                                      return (e_ == cudaErrorMemoryAllocation) ? PVC_ERR_MEMORY : PVC_ERR_CUDA; } } while (0)

    static int roundUp(int v, int m) { return (v + m - 1) / m * m; }

    static Layout makeLayout(const pvc_config& c, int historySteps = 0)        // historySteps: samples the history holds (0: c.T)
    {
        const int histT = historySteps > 0 ? historySteps : c.T;
        Layout L;
        L.gx = c.gx; L.gy = c.gy;
        L.rows = c.gx + 1; L.cols = c.gy + 1;
        L.tile_rows = fusedTileRows(c.reserved);
        L.warp_rows = fusedWarpRows(c.reserved);
        L.valid_rows = L.tile_rows - 2 * kTileK;
        L.tiles_x = (L.cols + kValidCols - 1) / kValidCols;
        // resident tiles only have to OWN the gx interior rows: the padding row gx is always current in the halo of the tile above it
        // (it depends on row gx - 1 of the same step only) and is recorded from there (pvc_step_res.cu)
        L.tiles_y = (variantKind(c.reserved) == 6) ? (L.gx + L.valid_rows - 1) / L.valid_rows : (L.rows + L.valid_rows - 1) / L.valid_rows;
        L.pitch = roundUp(L.tiles_x * kValidCols + 2 * kGuardCols, 32);
        L.rows_alloc = L.tiles_y * L.valid_rows + 2 * kGuardRows;
        if (L.rows_alloc < L.rows + kGuardRows + 1) L.rows_alloc = L.rows + kGuardRows + 1;
        L.plane = (size_t)L.rows_alloc * L.pitch;
        L.T = histT;
        L.hist_chunk = fusedHistChunk(c.reserved);
        L.hist_chunks = (L.cols + L.hist_chunk - 1) / L.hist_chunk;
        L.hist_row = (size_t)L.hist_chunks * histT * L.hist_chunk;
        L.hist_source = (size_t)L.rows * L.hist_row;
        return L;
    }

    // variant 0 = auto: the kernel with the smallest ESTIMATED time per pass (4 time steps of every source of the batch).  The
    // estimates are the measured pass periods of profiles/r02_resident_variants.txt and r02_small_tilings.txt (B200, microseconds):
    //   * resident kernel (pvc_step_res.cu; state in registers for the whole solve, only halo strips through the L2): a launch holds
    //     as many sources as fit co-resident and advances them one pass per period -- 3.0 (4-warp tiles, x 1.3 when two CTAs share an SM),
    //     3.95 (8-warp tiles; x 1.45 when two CTAs share an SM), 4.3 / 4.7 / 5.15 / 5.5 / 6.3 / 6.8 (10 / 12 / 14 / 16 / 18 /
    //     20 warps, one CTA per SM), all but the 4-warp tiles with the barrier-free row exchange -- a good third of it the
    //     neighbour hand-over, so the period barely depends on how full the GPU is, and the smallest tiling whose tiles all fit
    //     co-resident wins;
    //   * generational kernel (pvc_step_ws2.cu, TMA-staged tiles pulled from a work queue): 5.56 us per work item and SM (variant 47,
    //     56-row tiles), 3.9 (variant 50, 32-row tiles), and never less than the publish -> acquire -> TMA chain between generations
    //     (11.5 / 9.1 us).
    // So: one listener up to 1024^2 and batches of small grids (the plugin's case) run resident; a batch whose sources each fill the GPU
    // (four 1024^2 sources: 4 x 6.3 against 792 items x 5.56 / 148 = 29.8) runs resident too; batches that leave a resident
    // launch half empty (four 768^2 sources) and grids beyond the register files (2048^2) take the work queue.  Without
    // cuTensorMapEncodeTiled in the driver the generational kernels cannot run: resident, else the plain 8 x 6 kernel (18).
    static long residentTiles(const pvc_config& c, int v)
    {
        const int vr = fusedTileRows(v) - 2 * kTileK;
        return (long)((c.gx + vr - 1) / vr) * ((c.gy + 1 + kValidCols - 1) / kValidCols);
    }
    static int bestResident(const pvc_config& c, int sms, double* passUs)
    {
        // measured pass periods (profiles/r02_small_tilings.txt); `shared` = the factor when two CTAs sit on one SM
        static const struct { int v; double period, shared; } cand[] = { {67, 3.0, 1.3}, {69, 3.95, 1.45}, {72, 4.3, 1.0}, {70, 4.7, 1.0}, {71, 5.15, 1.0},
                                                                         {63, 5.5, 1.0}, {65, 6.3, 1.0}, {64, 6.8, 1.0} };
        int best = 0;
        double bestUs = 0;
        for (const auto& k : cand)
        {
            if (!variantAvailable(k.v)) continue;
            const long tiles = residentTiles(c, k.v), cap = (long)sms * variantMinBlocks(k.v);
            if (tiles > cap) continue;
            if (k.v == 67 && tiles * c.max_sources > 2L * sms) continue;       // measured up to two CTAs per SM (256^2 .. 384^2, batches of 128^2)
            const long perLaunch = cap / tiles < c.max_sources ? cap / tiles : c.max_sources;
            const long launches = (c.max_sources + perLaunch - 1) / perLaunch;
            const double us = (double)launches * k.period * (tiles * perLaunch > sms ? k.shared : 1.0);
            if (!best || us < bestUs) { best = k.v; bestUs = us; }
        }
        *passUs = bestUs;
        return best;
    }
    static int resolveVariant(const pvc_config& c, bool streamed = false)     // streamed: the plain fallback kernel (18) has no chunked form
    {
        if (c.reserved != 0 || c.step_kernel != 0) return c.reserved;
        int sms = 148;
        { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, c.device) == cudaSuccess && v > 0) sms = v; else cudaGetLastError(); }
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        const bool tma = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn != nullptr;
        if (!tma) cudaGetLastError();
        double residentUs = 0;
        const int resident = bestResident(c, sms, &residentUs);
        if (!tma) return resident ? resident : (streamed ? 0 : 18);
        const long cols = (c.gy + 1 + kValidCols - 1) / kValidCols;
        const double items47 = (double)((c.gx + 1 + 47) / 48) * cols * c.max_sources, items50 = (double)((c.gx + 1 + 23) / 24) * cols * c.max_sources;
        const double us47 = items47 * 5.56 / sms > 11.5 ? items47 * 5.56 / sms : 11.5;
        const double us50 = items50 * 3.9 / sms > 9.1 ? items50 * 3.9 / sms : 9.1;
        if (resident && residentUs <= us47 && residentUs <= us50) return resident;
        return us50 < us47 ? 50 : 47;
    }

    static bool validConfig(const pvc_config* c)
    {
        return c && c->gx >= 2 && c->gy >= 2 && c->T >= 1 && c->fs > 0 && c->resolution > 0 && c->dx > 0.f &&
               c->max_sources >= 1 && c->flux_samples >= 0 && c->dry_samples >= c->flux_samples &&
               c->wet_samples >= 0 && c->tail_samples >= 0 && (c->step_kernel == 0 || c->step_kernel == 1);
    }

    // ---- geometry kernels: Grid ctor field init (Grid.cpp:84-108) and AddAABB/RemoveAABB (:229-296) ----
    __global__ void clearGeometryKernel(Layout L, float* __restrict__ w)
    {
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= L.plane) return;
        const int r = (int)(i / L.pitch) - kGuardRows, c = (int)(i % L.pitch) - kGuardCols;
        float v;
        if (r < 0 || c < 0 || r > L.gx || c > L.gy) v = 0.f;                 // guard band: inert wall, Y = 0
        else if (r == L.gx || c == L.gy) v = 1.f;                            // padding row/col: b = 0, R = 0 -> Y = 1
        else v = __uint_as_float(kAirBits);
        w[i] = v;
    }

    // every alloc cell replays the ordered edit list and keeps the last edit covering it -- identical to
    // applying the rectangles one after another (GeometryManager.cpp:130-143), in one launch
    __global__ void applyRectsKernel(Layout L, float* __restrict__ w, const pvc_rect* __restrict__ rects, int n)
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        const int r = blockIdx.y * blockDim.y + threadIdx.y;
        if (r > L.gx || c > L.gy) return;                                    // clip 0..gx / 0..gy inclusive (Grid.cpp:231,235)
        const size_t i = cellIndex(L, r, c);
        float v = w[i];
        bool touched = false;
        for (int k = 0; k < n; ++k)
        {
            const pvc_rect q = rects[k];
            if (r >= q.r0 && r < q.r1 && c >= q.c0 && c < q.c1)
            {
                touched = true;
                if (q.add) v = q.admittance;
                else v = (r == L.gx || c == L.gy) ? 1.f : __uint_as_float(kAirBits);
            }
        }
        if (touched) w[i] = v;
    }

    __global__ void fetchCoefKernel(Layout L, const float* __restrict__ w, short* __restrict__ b, float* __restrict__ y)
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        const int r = blockIdx.y;
        if (c > L.gy) return;
        const float v = w[cellIndex(L, r, c)];
        const bool air = __float_as_uint(v) == kAirBits;
        b[(size_t)r * L.cols + c] = air ? 1 : 0;
        y[(size_t)r * L.cols + c] = air ? 1.f : v;       // air has R = 0 -> Y = 1 in the reference's terms
    }

    __global__ void gatherProbeKernel(const float* __restrict__ hist, size_t offset, int chunk, int n, float* __restrict__ out)
    {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        if (t < n) out[t] = hist[offset + (size_t)t * chunk];
    }

    // t < 0: a guarded state plane; t >= 0: sample t of a source's pressure history
    __global__ void unpackPlaneKernel(Layout L, const float* __restrict__ plane, int t, float* __restrict__ out)
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        const int r = blockIdx.y;
        if (c > L.gy) return;
        out[(size_t)r * L.cols + c] = (t < 0) ? plane[cellIndex(L, r, c)] : plane[histCell(L, r, c) + (size_t)t * L.hist_chunk];
    }

    int launchClearGeometry(pvc_solver* s)
    {
        const Layout& L = s->L;
        clearGeometryKernel<<<(unsigned)((L.plane + 255) / 256), 256, 0, s->stream>>>(L, s->w);
        s->slowMaskDirty = 1;
        PVC_CUDA(cudaGetLastError());
        return PVC_OK;
    }

    int launchApplyRects(pvc_solver* s, const pvc_rect* rects_host, int n)
    {
        if (n <= 0) return PVC_OK;
        const Layout& L = s->L;
        if (n > s->rectCapacity)
        {   // persistent edit-list buffers (stream-ordered allocation here cost up to tens of ms per frame in pool trimming)
            PVC_CUDA(cudaStreamSynchronize(s->stream));
            if (s->rects) cudaFree(s->rects);
            if (s->rectsHost) cudaFreeHost(s->rectsHost);
            s->rects = nullptr; s->rectsHost = nullptr; s->rectCapacity = 0;
            const int cap = n < 64 ? 64 : 2 * n;
            PVC_CUDA(cudaMalloc(&s->rects, sizeof(pvc_rect) * (size_t)cap));
            PVC_CUDA(cudaMallocHost(&s->rectsHost, sizeof(pvc_rect) * (size_t)cap * 2));
            s->rectCapacity = cap;
        }
        // the caller's list may be a temporary: stage it in pinned memory instead of waiting for the copy
        const unsigned slot = s->rectSlot++ & 1u;
        if (!s->rectCopied[slot]) PVC_CUDA(cudaEventCreateWithFlags(&s->rectCopied[slot], cudaEventDisableTiming));
        PVC_CUDA(cudaEventSynchronize(s->rectCopied[slot]));
        pvc_rect* staged = s->rectsHost + (size_t)slot * s->rectCapacity;
        memcpy(staged, rects_host, sizeof(pvc_rect) * (size_t)n);
        PVC_CUDA(cudaMemcpyAsync(s->rects, staged, sizeof(pvc_rect) * n, cudaMemcpyHostToDevice, s->stream));
        PVC_CUDA(cudaEventRecord(s->rectCopied[slot], s->stream));
        dim3 block(32, 8), grid((L.cols + 31) / 32, (L.rows + 7) / 8);
        applyRectsKernel<<<grid, block, 0, s->stream>>>(L, s->w, s->rects, n);
        PVC_CUDA(cudaGetLastError());
        s->slowMaskDirty = 1;
        return PVC_OK;
    }

    static int zeroState(pvc_solver* s, int nsrc)
    {
        s->cur = 0;
        // the resident kernel starts from zero in its registers and never reads the state planes: it only stores the final state of
        // the cells its tiles own, the same cells in every solve, so whatever else the planes hold stays the zero of pvc_create
        // (six memsets are 12 us of a 0.4 ms contract-grid frame)
        if (s->cfg.step_kernel == 0 && variantKind(s->cfg.reserved) == 6) return PVC_OK;
        for (int b = 0; b < 2; ++b)
            for (int f = 0; f < 3; ++f)
                PVC_CUDA(cudaMemsetAsync(s->state[b][f], 0, sizeof(float) * s->L.plane * nsrc, s->stream));
        s->cur = 0;
        return PVC_OK;
    }

    // The T/4 fused launches of a full solve are captured once per batch size into a CUDA graph and replayed:
    // every kernel argument (ping-pong pointers, t0, work-counter slot) is a pure function of the launch index, and
    // the buffers live as long as the solver.  The first solve of a batch size runs eagerly (it also performs the
    // one-time function-attribute setup), the second captures, later ones replay.
    static int runSteps(pvc_solver* s, int nsrc, int T, int* launches)
    {
        if (s->cfg.step_kernel == 1) return launchBaselineSteps(s, nsrc, 0, T, s->hist, launches);
        if (s->slowMaskDirty) { int rc = rebuildSlowMask(s); if (rc) return rc; }
        const bool graphable = s->useGraphs && T == s->cfg.T && nsrc >= 1 && nsrc <= kMaxGraphBatch && s->cur == 0;
        if (!graphable) return launchFusedSteps(s, nsrc, 0, T, s->hist, launches);
        GraphSlot& g = s->graphs[nsrc];
        if (g.exec)
        {
            cudaError_t e = cudaGraphLaunch(g.exec, s->stream);
            if (e != cudaSuccess) { setError("cudaGraphLaunch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            s->cur = g.finalCur;
            *launches += g.launches;
            return PVC_OK;
        }
        if (g.seen++ == 0) return launchFusedSteps(s, nsrc, 0, T, s->hist, launches);
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess)
        { cudaGetLastError(); s->useGraphs = 0; return launchFusedSteps(s, nsrc, 0, T, s->hist, launches); }
        int captured = 0;
        int rc = launchFusedSteps(s, nsrc, 0, T, s->hist, &captured);
        cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
        if (rc || e != cudaSuccess || !graph)
        {   // capture failed: fall back to eager launches for good
            cudaGetLastError(); s->useGraphs = 0; s->cur = 0;
            if (graph) cudaGraphDestroy(graph);
            return launchFusedSteps(s, nsrc, 0, T, s->hist, launches);
        }
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { cudaGetLastError(); g.exec = nullptr; s->useGraphs = 0; s->cur = 0; return launchFusedSteps(s, nsrc, 0, T, s->hist, launches); }
        g.finalCur = s->cur; g.launches = captured;
        s->cur = 0;
        e = cudaGraphLaunch(g.exec, s->stream);
        if (e != cudaSuccess) { setError("cudaGraphLaunch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        s->cur = g.finalCur;
        *launches += g.launches;
        return PVC_OK;
    }
}

namespace pvc
{
    __global__ void initCarryKernel(float* __restrict__ carry, size_t cstride, size_t n)
    {
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        carry[i] = __int_as_float(-1);                                       // plane 0: onset, none yet
        #pragma unroll
        for (int k = 1; k < kCarryPlanes; ++k) carry[k * cstride + i] = 0.f;
    }

    // Streamed solve (pvc_create_streamed): the response in chunks of s->chunkT samples through a history that holds one chunk.
    //   forward sweep   chunk 0 .. K-1: [checkpoint the state] -> time steps -> causal analyzer sums (launchStreamForward)
    //   backward sweep  chunk K-1 (its history is still resident), then K-2 .. 0: restore the checkpoint -> the SAME time steps again
    //                   -> backward Schroeder sums (launchStreamBackward; chunk 0 finishes the regression and writes the results)
    // Re-running a chunk from its checkpoint reproduces its history bit for bit (the kernels are deterministic: no atomics or
    // order-dependent arithmetic in the field update), so the outputs equal the full-history solve's exactly, at (2 - 1/K) x
    // the time steps and 1/K of the history memory.  Checkpoints: the three state planes at the start of chunks 1 .. K-2 (chunk 0
    // starts from zero, chunk K-1 is never recomputed).
    static int runStreamed(pvc_solver* s, int nsrc, int analyze, int* launches)
    {
        const Layout& L = s->L;
        const int T = s->cfg.T, C = s->chunkT, K = (T + C - 1) / C;
        const size_t S = (size_t)s->cfg.max_sources, cells = (size_t)L.gx * L.gy;
        if (s->slowMaskDirty) { int rc = rebuildSlowMask(s); if (rc) return rc; }
        if (analyze)
        {
            const size_t n = S * cells;
            initCarryKernel<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->carry, n, n);
            *launches += 1;
        }
        auto copyState = [&](float* const dst[3], float* const src[3]) -> int {
            for (int f = 0; f < 3; ++f)
                PVC_CUDA(cudaMemcpyAsync(dst[f], src[f], sizeof(float) * L.plane * (size_t)nsrc, cudaMemcpyDeviceToDevice, s->stream));
            return PVC_OK;
        };
        auto slot = [&](int k, float* out[3]) { for (int f = 0; f < 3; ++f) out[f] = s->ckpt + ((size_t)(k - 1) * 3 + f) * S * L.plane; };
        for (int k = 0; k < K; ++k)
        {
            const int base = k * C, end = (base + C < T) ? base + C : T;
            if (s->cur != 0) { setError("streamed solve: chunk %d does not start on ping-pong buffer 0", k); return PVC_ERR_CUDA; }
            if (analyze && k >= 1 && k <= K - 2) { float* q[3]; slot(k, q); int rc = copyState(q, s->state[0]); if (rc) return rc; }
            s->finalPass = (k == K - 1);
            s->abortSticky = (k > 0);
            // activity hints of this chunk (first generation, counted from the chunk's start, in which a warp's block recorded
            // anything but zeros): the forward analysis skips the exact zeros in front of the wave
            s->hintsValid = analyze ? 1 : 0;
            if (s->hintsValid)
                PVC_CUDA(cudaMemsetAsync(s->firstActive, 0x7f, sizeof(int) * (size_t)nsrc * L.tiles_x * L.tiles_y * 32, s->stream));
            int rc = launchFusedSteps(s, nsrc, base, end, s->hist, launches);
            if (!rc && analyze) rc = launchStreamForward(s, nsrc, base, end - base, launches);
            s->hintsValid = 0;
            if (rc) return rc;
        }
        s->stateStale = 0;
        s->lastStepLaunches = *launches;
        if (!analyze) { s->abortSticky = 0; return PVC_OK; }
        // a pipelined fetch of the previous run's grids must finish before the results are overwritten
        if (s->copyPending) PVC_CUDA(cudaStreamWaitEvent(s->stream, s->evCopied, 0));
        int rc = launchStreamBackward(s, nsrc, (K - 1) * C, T - (K - 1) * C, launches);
        if (rc) return rc;
        for (int k = K - 2; k >= 0; --k)
        {
            if (k == 0) { rc = zeroState(s, nsrc); if (rc) return rc; }
            else { float* q[3]; slot(k, q); rc = copyState(s->state[0], q); if (rc) return rc; s->cur = 0; }
            s->finalPass = 0;
            s->stateStale = 1;                      // from here on the state planes hold the end of chunk k, not of the response
            rc = launchFusedSteps(s, nsrc, k * C, (k + 1) * C, s->hist, launches);
            if (!rc) rc = launchStreamBackward(s, nsrc, k * C, C, launches);
            if (rc) return rc;
        }
        s->finalPass = 1;
        s->abortSticky = 0;
        return launchListenerDirection(s, nsrc, launches);
    }
}

using namespace pvc;

extern "C" {

int pvc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* pvc_last_error(void) { return g_error; }

int pvc_device_memory(int device, size_t* free_bytes, size_t* total_bytes)
{
    if (!free_bytes || !total_bytes) { setError("pvc_device_memory: null argument"); return PVC_ERR_INVALID; }
    if (device < 0 || device >= pvc_device_count()) { setError("pvc_device_memory: device %d", device); return PVC_ERR_NO_DEVICE; }
    PVC_CUDA(cudaSetDevice(device));
    PVC_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return PVC_OK;
}

// chunk length of a streamed solve as the device uses it: a multiple of 8 samples (an even number of 4-step generations, so every
// chunk starts on ping-pong buffer 0), at least 8, and no chunking at all (0) when the whole response fits
static int streamChunk(const pvc_config* cfg, int history_steps)
{
    if (history_steps <= 0) return 0;
    int c = history_steps / 8 * 8;
    if (c < 8) c = 8;
    return c;
}

static size_t memoryRequirement(const pvc_config* cfg, int history_steps)
{
    if (!validConfig(cfg)) return 0;
    const int chunk = streamChunk(cfg, history_steps);
    pvc_config r = *cfg; r.reserved = resolveVariant(*cfg, chunk > 0);
    const Layout L = makeLayout(r, chunk);
    const size_t S = (size_t)cfg->max_sources, cells = (size_t)cfg->gx * cfg->gy;
    size_t floats = 6 * S * L.plane + (variantKind(r.reserved) == 6 ? 7 + 32 * S : (variantKind(r.reserved) == 5 ? 7 : 4)) * L.plane + S * L.hist_source + (size_t)cfg->T +
                    S * cells * 11 + 3 * (size_t)cfg->T;
    if (chunk > 0)
    {
        const int K = (cfg->T + chunk - 1) / chunk;
        floats += (size_t)kCarryPlanes * S * cells + (K > 2 ? (size_t)(K - 2) * 3 * S * L.plane : 0);
    }
    return sizeof(float) * floats + (size_t)L.tiles_x * L.tiles_y * 64;
}

size_t pvc_memory_requirement(const pvc_config* cfg) { return memoryRequirement(cfg, 0); }
size_t pvc_memory_requirement_streamed(const pvc_config* cfg, int history_steps) { return memoryRequirement(cfg, history_steps); }

static int createSolver(const pvc_config* cfg, int history_steps, pvc_solver** out);
int pvc_create(const pvc_config* cfg, pvc_solver** out) { return createSolver(cfg, 0, out); }
int pvc_create_streamed(const pvc_config* cfg, int history_steps, pvc_solver** out)
{
    if (history_steps < 1) { setError("pvc_create_streamed: history_steps must be positive"); if (out) *out = nullptr; return PVC_ERR_INVALID; }
    return createSolver(cfg, history_steps, out);
}

static int createSolver(const pvc_config* cfg, int history_steps, pvc_solver** out)
{
    if (!out) { setError("pvc_create: null out"); return PVC_ERR_INVALID; }
    *out = nullptr;
    if (!validConfig(cfg)) { setError("pvc_create: invalid config"); return PVC_ERR_INVALID; }
    int ndev = pvc_device_count();
    if (ndev <= 0) { setError("pvc_create: no CUDA device (this library has no CPU path)"); return PVC_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { setError("pvc_create: device %d of %d", cfg->device, ndev); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(cfg->device));

    pvc_solver* s = new pvc_solver();
    memset(s, 0, sizeof(*s));
    s->cfg = *cfg;
    s->chunkT = streamChunk(cfg, history_steps);
    s->finalPass = 1;
    s->cfg.reserved = resolveVariant(*cfg, s->chunkT > 0);
    if (s->chunkT && (cfg->step_kernel != 0 || !((s->cfg.reserved == 47 || s->cfg.reserved == 50 || variantKind(s->cfg.reserved) == 6) && variantAvailable(s->cfg.reserved))))
    {
        setError("pvc_create_streamed: a streamed solve needs a step kernel that can continue from a stored state (the generational variants 47 / 50 or a resident tiling); got step_kernel %d, variant %d",
                 cfg->step_kernel, s->cfg.reserved);
        delete s;
        return PVC_ERR_INVALID;
    }
    if (cfg->step_kernel == 0 && !variantAvailable(s->cfg.reserved))
    {
        setError("pvc_create: step-kernel variant %d is not compiled into this build (make EXTRA=-DPVC_ALL_VARIANTS)", s->cfg.reserved);
        delete s;
        return PVC_ERR_INVALID;
    }
    s->device = cfg->device;
    s->L = makeLayout(s->cfg, s->chunkT);
    { int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device); s->numSMs = sms > 0 ? sms : 148; }
    const Layout& L = s->L;
    const size_t S = (size_t)cfg->max_sources, cells = (size_t)cfg->gx * cfg->gy;
    #define PVC_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                              \
        setError("%s: %s", #call, cudaGetErrorString(e_)); pvc_destroy(s);                                     \
        return (e_ == cudaErrorMemoryAllocation) ? PVC_ERR_MEMORY : PVC_ERR_CUDA; } } while (0)
    PVC_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) PVC_TRY(cudaEventCreate(&s->ev[i]));
    for (int i = 0; i < 2; ++i) PVC_TRY(cudaEventCreate(&s->mark[i]));
    PVC_TRY(cudaStreamCreateWithFlags(&s->copyStream, cudaStreamNonBlocking));
    PVC_TRY(cudaEventCreateWithFlags(&s->evAnalyzed, cudaEventDisableTiming));
    PVC_TRY(cudaEventCreateWithFlags(&s->evCopied, cudaEventDisableTiming));
    PVC_TRY(cudaMallocHost(&s->hostAbort, sizeof(int)));
    *s->hostAbort = 0;
    // the six state arrays in ONE allocation, so that a single L2 access-policy window can cover the ping-pong state
    PVC_TRY(cudaMalloc(&s->stateBlock, sizeof(float) * 6 * S * L.plane));
    PVC_TRY(cudaMemsetAsync(s->stateBlock, 0, sizeof(float) * 6 * S * L.plane, s->stream));
    for (int b = 0; b < 2; ++b)
        for (int f = 0; f < 3; ++f) s->state[b][f] = s->stateBlock + (size_t)(b * 3 + f) * S * L.plane;
    PVC_TRY(cudaMalloc(&s->w, sizeof(float) * L.plane));
    for (int f = 0; f < 3; ++f) PVC_TRY(cudaMalloc(&s->coef[f], sizeof(float) * L.plane));
    PVC_TRY(cudaMalloc(&s->slowMask, sizeof(uint32_t) * (size_t)L.tiles_x * L.tiles_y * 32));
    PVC_TRY(cudaMemsetAsync(s->slowMask, 0, sizeof(uint32_t) * (size_t)L.tiles_x * L.tiles_y * 32, s->stream));
    PVC_TRY(cudaMalloc(&s->tileOrder, sizeof(int) * (size_t)L.tiles_x * L.tiles_y));
    PVC_TRY(cudaMalloc(&s->bpMask, sizeof(uint32_t) * (size_t)L.tiles_x * L.tiles_y * 32 * 32));
    PVC_TRY(cudaMalloc(&s->firstActive, sizeof(int) * S * (size_t)L.tiles_x * L.tiles_y * 32));
    PVC_TRY(cudaMemsetAsync(s->firstActive, 0x7f, sizeof(int) * S * (size_t)L.tiles_x * L.tiles_y * 32, s->stream));
    s->tileCounterCount = cfg->T / kTileK + 2;
    PVC_TRY(cudaMalloc(&s->tileCounters, sizeof(int) * (size_t)s->tileCounterCount));
    PVC_TRY(cudaMalloc(&s->doneGen, sizeof(int) * S * (size_t)L.tiles_x * L.tiles_y));
    if (variantKind(s->cfg.reserved) >= 5)
        for (int f = 0; f < 3; ++f) PVC_TRY(cudaMalloc(&s->lin[f], sizeof(float) * L.plane));
    if (variantKind(s->cfg.reserved) == 6)
    {
        PVC_TRY(cudaMalloc(&s->resXchg, sizeof(float) * 4 * 8 * S * L.plane));
        PVC_TRY(cudaMemsetAsync(s->resXchg, 0, sizeof(float) * 4 * 8 * S * L.plane, s->stream));
    }
    PVC_TRY(cudaMalloc(&s->hist, sizeof(float) * S * L.hist_source));
    if (s->chunkT)
    {
        const int K = (cfg->T + s->chunkT - 1) / s->chunkT;
        PVC_TRY(cudaMalloc(&s->carry, sizeof(float) * (size_t)kCarryPlanes * S * cells));
        if (K > 2) PVC_TRY(cudaMalloc(&s->ckpt, sizeof(float) * (size_t)(K - 2) * 3 * S * L.plane));
    }
    PVC_TRY(cudaMalloc(&s->pulse, sizeof(float) * (size_t)cfg->T));
    PVC_TRY(cudaMemsetAsync(s->pulse, 0, sizeof(float) * (size_t)cfg->T, s->stream));
    PVC_TRY(cudaMalloc(&s->results, sizeof(float) * S * cells * 8));
    PVC_TRY(cudaMemsetAsync(s->results, 0, sizeof(float) * S * cells * 8, s->stream));
    PVC_TRY(cudaMalloc(&s->delay, sizeof(float) * S * cells));
    PVC_TRY(cudaMalloc(&s->walkDelay, sizeof(float) * S * cells));
    PVC_TRY(cudaMalloc(&s->walkNext, sizeof(int) * S * cells));
    PVC_TRY(cudaMalloc(&s->scratch, sizeof(float) * 3 * (size_t)cfg->T));
    PVC_TRY(cudaMalloc(&s->src, sizeof(SourceParams) * S));
    PVC_TRY(cudaMallocHost(&s->srcHost, sizeof(SourceParams) * S * 4));
    for (int i = 0; i < 4; ++i) PVC_TRY(cudaEventCreateWithFlags(&s->srcCopied[i], cudaEventDisableTiming));
    #undef PVC_TRY
    s->efree = 1.f;
    s->useGraphs = variantKind(s->cfg.reserved) == 0 ? 1 : 0;      // only the one-launch-per-4-steps kernels have enough launches to replay
#ifdef PVC_TUNING
    if (getenv("PVC_NO_GRAPHS")) s->useGraphs = 0;
#endif
    buildTensorMaps(s);
    int rc = launchClearGeometry(s);
    if (rc) { pvc_destroy(s); return rc; }
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) { setError("pvc_create: sync failed"); pvc_destroy(s); return PVC_ERR_CUDA; }
    *out = s;
    return PVC_OK;
}

void pvc_destroy(pvc_solver* s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->stateBlock); cudaFree(s->ckpt); cudaFree(s->carry);
    cudaFree(s->w); for (int f = 0; f < 3; ++f) { cudaFree(s->coef[f]); cudaFree(s->lin[f]); } cudaFree(s->resXchg); cudaFree(s->rects); cudaFree(s->slowMask); cudaFree(s->bpMask); cudaFree(s->tileOrder); cudaFree(s->firstActive); cudaFree(s->tileCounters); cudaFree(s->doneGen); cudaFree(s->hist); cudaFree(s->pulse);
    if (s->copyStream) { cudaStreamSynchronize(s->copyStream); cudaStreamDestroy(s->copyStream); }
    if (s->evAnalyzed) cudaEventDestroy(s->evAnalyzed);
    if (s->evCopied) cudaEventDestroy(s->evCopied);
    if (s->hostAbort) cudaFreeHost(s->hostAbort);
    if (s->srcHost) cudaFreeHost(s->srcHost);
    for (int i = 0; i < 2; ++i) if (s->gathered[i]) cudaEventDestroy(s->gathered[i]);
    if (s->rectsHost) cudaFreeHost(s->rectsHost);
    for (int i = 0; i < 2; ++i) if (s->rectCopied[i]) cudaEventDestroy(s->rectCopied[i]);
    for (int i = 0; i < 4; ++i) if (s->srcCopied[i]) cudaEventDestroy(s->srcCopied[i]);
    cudaFree(s->results); cudaFree(s->delay); cudaFree(s->walkDelay); cudaFree(s->walkNext); cudaFree(s->scratch); cudaFree(s->src);
    for (int i = 0; i <= kMaxGraphBatch; ++i) if (s->graphs[i].exec) cudaGraphExecDestroy(s->graphs[i].exec);
    for (int i = 0; i < 4; ++i) if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    for (int i = 0; i < 2; ++i) if (s->mark[i]) cudaEventDestroy(s->mark[i]);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int pvc_set_pulse(pvc_solver* s, const float* pulse, int n)
{
    if (!s || !pulse || n < 0) { setError("pvc_set_pulse: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    if (n > s->cfg.T) n = s->cfg.T;
    PVC_CUDA(cudaMemsetAsync(s->pulse, 0, sizeof(float) * (size_t)s->cfg.T, s->stream));
    PVC_CUDA(cudaMemcpyAsync(s->pulse, pulse, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    return PVC_OK;
}

int pvc_clear_geometry(pvc_solver* s)
{
    if (!s) { setError("pvc_clear_geometry: null solver"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    return launchClearGeometry(s);
}

int pvc_apply_geometry(pvc_solver* s, const pvc_rect* rects, int n)
{
    if (!s || (n > 0 && !rects) || n < 0) { setError("pvc_apply_geometry: bad argument"); return PVC_ERR_INVALID; }
    // a wall's admittance Y = (1-R)/(1+R) lies in [0, 1] for every reflection coefficient R in [0, 1]; a negative one (R > 1, an
    // amplifying wall) has no physical meaning and the coefficient planes of pvc_step_res.cu keep a flag in its sign
    for (int i = 0; i < n; ++i)
        if (rects[i].add && !(rects[i].admittance >= 0.f && rects[i].admittance <= 1.f))
        { setError("pvc_apply_geometry: edit %d has admittance %g outside [0, 1] (absorption must lie in [0, 1])", i, (double)rects[i].admittance); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    return launchApplyRects(s, rects, n);
}

int pvc_fetch_coefficients(pvc_solver* s, int16_t* b, float* admittance)
{
    if (!s || !b || !admittance) { setError("pvc_fetch_coefficients: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const Layout& L = s->L;
    const size_t n = (size_t)L.rows * L.cols;
    short* db = nullptr; float* dy = nullptr;
    PVC_CUDA(cudaMalloc(&db, sizeof(short) * n));
    PVC_CUDA(cudaMalloc(&dy, sizeof(float) * n));
    fetchCoefKernel<<<dim3((L.cols + 127) / 128, L.rows), 128, 0, s->stream>>>(L, s->w, db, dy);
    PVC_CUDA(cudaMemcpyAsync(b, db, sizeof(short) * n, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaMemcpyAsync(admittance, dy, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(db); cudaFree(dy);
    return PVC_OK;
}

int pvc_set_efree(pvc_solver* s, float efree)
{
    if (!s) { setError("pvc_set_efree: null solver"); return PVC_ERR_INVALID; }
    s->efree = efree;
    return PVC_OK;
}

int pvc_compute_efree(pvc_solver* s, int lr, int lc, int er, int ec, int n, float r, float* efree)
{
    if (!s || n < 1 || n > s->cfg.T || lr < 0 || lc < 0 || lr > s->cfg.gx || lc > s->cfg.gy ||
        er < 0 || ec < 0 || er > s->cfg.gx || ec > s->cfg.gy)
    { setError("pvc_compute_efree: bad argument (n=%d, T=%d)", n, s ? s->cfg.T : -1); return PVC_ERR_INVALID; }
    if (s->chunkT && n > s->chunkT) { setError("pvc_compute_efree: the %d-sample probe does not fit the %d-sample history of this streamed solver", n, s->chunkT); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const Layout& L = s->L;
    // the free field is a second, empty coefficient plane (FreeGrid.cpp:11-18): swap it in for n steps
    float* wScene = s->w;
    float* wFree = nullptr;
    PVC_CUDA(cudaMalloc(&wFree, sizeof(float) * L.plane));
    s->w = wFree;
    int rc = launchClearGeometry(s);
    SourceParams sp{ lr, lc, lr, lc, 0.f, 0.f, 0 };
    if (!rc && cudaMemcpyAsync(s->src, &sp, sizeof(sp), cudaMemcpyHostToDevice, s->stream) != cudaSuccess) rc = PVC_ERR_CUDA;
    if (!rc) rc = markDeadSources(s, 1);
    if (!rc) rc = zeroState(s, 1);
    int launches = 0;
    if (!rc) rc = runSteps(s, 1, n, &launches);
    std::vector<float> probe((size_t)n);
    if (!rc)
    {
        gatherProbeKernel<<<(n + 127) / 128, 128, 0, s->stream>>>(s->hist, histCell(L, er, ec), L.hist_chunk, n, s->scratch);
        if (cudaMemcpyAsync(probe.data(), s->scratch, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
            cudaStreamSynchronize(s->stream) != cudaSuccess)
        { setError("pvc_compute_efree: %s", cudaGetErrorString(cudaGetLastError())); rc = PVC_ERR_CUDA; }
    }
    s->w = wScene;
    s->slowMaskDirty = 1;
    cudaFree(wFree);
    if (rc) return rc;
    // FreeGrid::CalculateEFree (FreeGrid.cpp:96-110): sequential fp32 sum of squares, then times r (:89-91)
    volatile float e = 0.f;
    for (int i = 0; i < n; ++i) { volatile float sq = probe[(size_t)i] * probe[(size_t)i]; e = e + sq; }
    volatile float scaled = e * r;
    s->efree = scaled;
    if (efree) *efree = scaled;
    return PVC_OK;
}

int pvc_run(pvc_solver* s, const pvc_listener* listeners, int n, int analyze)
{
    if (!s || !listeners || n < 1 || n > s->cfg.max_sources) { setError("pvc_run: bad argument (n=%d)", n); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    for (int i = 0; i < n; ++i)
    {
        const pvc_listener& l = listeners[i];
        if (l.cell_r < 0 || l.cell_c < 0 || l.cell_r > s->cfg.gx || l.cell_c > s->cfg.gy)
        { setError("pvc_run: listener %d cell (%d,%d) outside the grid", i, l.cell_r, l.cell_c); return PVC_ERR_INVALID; }
    }
    // staged through a pinned ring so that a frame loop can enqueue run k+1 while run k is still on the device
    const unsigned slot = s->srcSlot++ & 3u;
    PVC_CUDA(cudaEventSynchronize(s->srcCopied[slot]));
    SourceParams* sp = s->srcHost + (size_t)slot * s->cfg.max_sources;
    for (int i = 0; i < n; ++i)
    {
        const pvc_listener& l = listeners[i];
        sp[i] = SourceParams{ l.cell_r, l.cell_c, l.efree_r, l.efree_c, l.x, l.z, 0 };
    }
    PVC_CUDA(cudaMemcpyAsync(s->src, sp, sizeof(SourceParams) * n, cudaMemcpyHostToDevice, s->stream));
    PVC_CUDA(cudaEventRecord(s->srcCopied[slot], s->stream));
    { const int rcDead = markDeadSources(s, n); if (rcDead) return rcDead; }
    int launches = 0;
    PVC_CUDA(cudaEventRecord(s->ev[0], s->stream));
    int rc = zeroState(s, n);
    if (rc) return rc;
    s->hintsValid = (s->cfg.step_kernel == 0);
    if (s->hintsValid)
        PVC_CUDA(cudaMemsetAsync(s->firstActive, 0x7f, sizeof(int) * (size_t)n * s->L.tiles_x * s->L.tiles_y * 32, s->stream));
    if (s->chunkT)
    {   // streamed solve: time steps and analysis interleave chunk by chunk (the step/analyzer split of pvc_last_timing is not
        // meaningful: everything is reported as step time)
        s->hintsValid = 0;
        rc = runStreamed(s, n, analyze, &launches);
        s->abortSticky = 0; s->finalPass = 1;          // also after a failed sweep
        if (rc) return rc;
        PVC_CUDA(cudaEventRecord(s->ev[1], s->stream));
        PVC_CUDA(cudaEventRecord(s->ev[2], s->stream));
        s->lastSources = n;
        s->lastLaunches = launches;
        return PVC_OK;
    }
    rc = runSteps(s, n, s->cfg.T, &launches);
    if (rc) return rc;
    s->lastStepLaunches = launches;
    PVC_CUDA(cudaEventRecord(s->ev[1], s->stream));
    if (analyze)
    {
        // a pipelined fetch of the previous run's grids must finish before the analyzer overwrites them
        if (s->copyPending) PVC_CUDA(cudaStreamWaitEvent(s->stream, s->evCopied, 0));
        rc = launchAnalyzer(s, n, &launches);
        if (rc) return rc;
    }
    PVC_CUDA(cudaEventRecord(s->ev[2], s->stream));
    s->lastSources = n;
    s->lastLaunches = launches;
    return PVC_OK;
}

static int checkAbort(pvc_solver* s)
{
    if (!s->checkAbort) return PVC_OK;
    s->checkAbort = 0;
    int flag = 0;
    PVC_CUDA(cudaMemcpyAsync(&flag, s->tileCounters, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    if (flag) { setError("generational step kernel: a tile dependency wait timed out (results invalid)"); return PVC_ERR_CUDA; }
    return PVC_OK;
}

int pvc_synchronize(pvc_solver* s)
{
    if (!s) { setError("pvc_synchronize: null solver"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    return checkAbort(s);
}

int pvc_clear_results(pvc_solver* s, int source)
{
    if (!s || source < 0 || source >= s->cfg.max_sources) { setError("pvc_clear_results: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const size_t cells = (size_t)s->cfg.gx * s->cfg.gy;
    PVC_CUDA(cudaMemsetAsync(s->results + (size_t)source * cells * 8, 0, sizeof(float) * cells * 8, s->stream));
    return PVC_OK;
}

int pvc_fetch_results(pvc_solver* s, int source, float* results, float* delay)
{
    if (!s || source < 0 || source >= s->cfg.max_sources) { setError("pvc_fetch_results: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const size_t cells = (size_t)s->cfg.gx * s->cfg.gy;
    if (results) PVC_CUDA(cudaMemcpyAsync(results, s->results + (size_t)source * cells * 8, sizeof(float) * cells * 8, cudaMemcpyDeviceToHost, s->stream));
    if (delay) PVC_CUDA(cudaMemcpyAsync(delay, s->delay + (size_t)source * cells, sizeof(float) * cells, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    return checkAbort(s);
}

int pvc_fetch_results_async(pvc_solver* s, int n, float* results, float* delay)
{
    if (!s || n < 1 || n > s->cfg.max_sources) { setError("pvc_fetch_results_async: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    if (s->copyPending)
    {   // one copy in flight at a time; a run whose step kernel gave up on a dependency wait must not pass for a good frame
        PVC_CUDA(cudaEventSynchronize(s->evCopied));
        s->copyPending = 0;
        if (*s->hostAbort) { *s->hostAbort = 0; setError("step kernel: a tile dependency wait timed out in the previous frame (its results are invalid)"); return PVC_ERR_CUDA; }
    }
    const size_t cells = (size_t)s->cfg.gx * s->cfg.gy;
    *s->hostAbort = 0;
    if (s->checkAbort)
    {
        // the abort word is read on the solver's own stream: the next pvc_run resets it there
        PVC_CUDA(cudaMemcpyAsync(s->hostAbort, s->tileCounters, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
        s->checkAbort = 0;
    }
    PVC_CUDA(cudaEventRecord(s->evAnalyzed, s->stream));
    PVC_CUDA(cudaStreamWaitEvent(s->copyStream, s->evAnalyzed, 0));
    if (results) PVC_CUDA(cudaMemcpyAsync(results, s->results, sizeof(float) * cells * 8 * (size_t)n, cudaMemcpyDeviceToHost, s->copyStream));
    if (delay) PVC_CUDA(cudaMemcpyAsync(delay, s->delay, sizeof(float) * cells * (size_t)n, cudaMemcpyDeviceToHost, s->copyStream));
    PVC_CUDA(cudaEventRecord(s->evCopied, s->copyStream));
    s->copyPending = 1;
    return PVC_OK;
}

int pvc_fetch_wait(pvc_solver* s)
{
    if (!s) { setError("pvc_fetch_wait: null solver"); return PVC_ERR_INVALID; }
    if (!s->copyPending) return PVC_OK;
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaEventSynchronize(s->evCopied));
    s->copyPending = 0;
    if (*s->hostAbort) { *s->hostAbort = 0; setError("generational step kernel: a tile dependency wait timed out (results invalid)"); return PVC_ERR_CUDA; }
    return PVC_OK;
}

int pvc_fetch_result_at(pvc_solver* s, int source, int r, int c, float* out8)
{
    if (!s || !out8 || source < 0 || source >= s->cfg.max_sources || r < 0 || c < 0 || r >= s->cfg.gx || c >= s->cfg.gy)
    { setError("pvc_fetch_result_at: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const size_t cells = (size_t)s->cfg.gx * s->cfg.gy;
    PVC_CUDA(cudaMemcpyAsync(out8, s->results + ((size_t)source * cells + (size_t)r * s->cfg.gy + c) * 8, sizeof(float) * 8, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    return PVC_OK;
}

int pvc_gather_results_async(pvc_solver* s, int n, const int* cells, int n_cells, float* out, int* ticket)
{
    if (!s || !cells || !out || !ticket || n < 1 || n > s->cfg.max_sources || n_cells < 0)
    { setError("pvc_gather_results_async: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const size_t total = (size_t)s->cfg.gx * s->cfg.gy;
    for (int i = 0; i < n_cells; ++i)
        if (cells[i] >= 0 && (size_t)cells[i] >= total) { setError("pvc_gather_results_async: cell %d outside the lattice", cells[i]); return PVC_ERR_INVALID; }
    const unsigned slot = s->gatherSlot++ & 1u;
    if (!s->gathered[slot]) PVC_CUDA(cudaEventCreateWithFlags(&s->gathered[slot], cudaEventDisableTiming));
    for (int src = 0; src < n; ++src)
        for (int i = 0; i < n_cells; ++i)
            if (cells[i] >= 0)
            PVC_CUDA(cudaMemcpyAsync(out + ((size_t)src * n_cells + i) * 8, s->results + ((size_t)src * total + (size_t)cells[i]) * 8,
                                     sizeof(float) * 8, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaEventRecord(s->gathered[slot], s->stream));
    *ticket = (int)slot;
    return PVC_OK;
}

int pvc_gather_wait(pvc_solver* s, int ticket)
{
    if (!s || ticket < 0 || ticket > 1) { setError("pvc_gather_wait: bad argument"); return PVC_ERR_INVALID; }
    if (!s->gathered[ticket]) return PVC_OK;
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaEventSynchronize(s->gathered[ticket]));
    return PVC_OK;
}

int pvc_fetch_ir(pvc_solver* s, int source, int r, int c, float* out3T)
{
    if (!s || !out3T || source < 0 || source >= s->cfg.max_sources || r < 0 || c < 0 || r > s->cfg.gx || c > s->cfg.gy)
    { setError("pvc_fetch_ir: bad argument"); return PVC_ERR_INVALID; }
    if (s->chunkT) { setError("pvc_fetch_ir: a streamed solver keeps no full history"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    int rc = launchIrRebuild(s, source, r, c, s->scratch);
    if (rc) return rc;
    PVC_CUDA(cudaMemcpyAsync(out3T, s->scratch, sizeof(float) * 3 * (size_t)s->cfg.T, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    return PVC_OK;
}

static int fetchPlane(pvc_solver* s, const float* dev, int t, float* host)
{
    const Layout& L = s->L;
    const size_t n = (size_t)L.rows * L.cols;
    float* tmp = nullptr;
    PVC_CUDA(cudaMalloc(&tmp, sizeof(float) * n));
    unpackPlaneKernel<<<dim3((L.cols + 127) / 128, L.rows), 128, 0, s->stream>>>(L, dev, t, tmp);
    PVC_CUDA(cudaMemcpyAsync(host, tmp, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
    PVC_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(tmp);
    return PVC_OK;
}

int pvc_fetch_pressure(pvc_solver* s, int source, int t, float* plane)
{
    if (!s || !plane || source < 0 || source >= s->cfg.max_sources || t < 0 || t >= s->cfg.T)
    { setError("pvc_fetch_pressure: bad argument"); return PVC_ERR_INVALID; }
    if (s->chunkT) { setError("pvc_fetch_pressure: a streamed solver keeps no full history"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    return fetchPlane(s, s->hist + (size_t)source * s->L.hist_source, t, plane);
}

int pvc_fetch_state(pvc_solver* s, int source, float* p, float* vx, float* vy)
{
    if (!s || source < 0 || source >= s->cfg.max_sources) { setError("pvc_fetch_state: bad argument"); return PVC_ERR_INVALID; }
    if (s->chunkT && s->stateStale) { setError("pvc_fetch_state: after an analyzed streamed solve the state planes hold the end of its first chunk (run with analyze = 0 for the final state)"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    float* host[3] = { p, vx, vy };
    for (int f = 0; f < 3; ++f)
        if (host[f]) { int rc = fetchPlane(s, s->state[s->cur][f] + (size_t)source * s->L.plane, -1, host[f]); if (rc) return rc; }
    return PVC_OK;
}

int pvc_last_timing(pvc_solver* s, float* out3, int* launches)
{
    if (!s) { setError("pvc_last_timing: null solver"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaEventSynchronize(s->ev[2]));
    float a = 0.f, b = 0.f;
    PVC_CUDA(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
    PVC_CUDA(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
    if (out3) { out3[0] = a; out3[1] = b; out3[2] = a + b; }
    if (launches) *launches = s->lastLaunches;
    return PVC_OK;
}

int pvc_last_launch_counts(pvc_solver* s, int* step_launches, int* analyzer_launches)
{
    if (!s) { setError("pvc_last_launch_counts: null solver"); return PVC_ERR_INVALID; }
    if (step_launches) *step_launches = s->lastStepLaunches;
    if (analyzer_launches) *analyzer_launches = s->lastLaunches - s->lastStepLaunches;
    return PVC_OK;
}

int pvc_mark(pvc_solver* s, int which)
{
    if (!s || which < 0 || which > 1) { setError("pvc_mark: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaEventRecord(s->mark[which], s->stream));
    return PVC_OK;
}

int pvc_mark_elapsed(pvc_solver* s, float* ms)
{
    if (!s || !ms) { setError("pvc_mark_elapsed: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    PVC_CUDA(cudaEventSynchronize(s->mark[1]));
    PVC_CUDA(cudaEventElapsedTime(ms, s->mark[0], s->mark[1]));
    return PVC_OK;
}

const float* pvc_results_dev(pvc_solver* s, int source)
{
    if (!s || source < 0 || source >= s->cfg.max_sources) return nullptr;
    return s->results + (size_t)source * s->cfg.gx * s->cfg.gy * 8;
}

int pvc_debug_timeline(pvc_solver* s, int nsrc, unsigned long long* out, int maxBlocks)
{
    if (!s || !out || nsrc < 1 || nsrc > s->cfg.max_sources || s->cfg.step_kernel != 0) { setError("pvc_debug_timeline: bad argument"); return PVC_ERR_INVALID; }
    PVC_CUDA(cudaSetDevice(s->device));
    const int blocks = s->L.tiles_x * s->L.tiles_y * nsrc;
    if (blocks > maxBlocks) { setError("pvc_debug_timeline: need room for %d blocks", blocks); return PVC_ERR_INVALID; }
    unsigned long long* d = nullptr;
    PVC_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * 8 * blocks));
    PVC_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 8 * blocks, s->stream));
    if (s->slowMaskDirty) { int rc = rebuildSlowMask(s); if (rc) return rc; }
    int launches = 0;
    // a few untimed launches to reach steady state, then one stamped launch (state keeps evolving, harmless)
    int rc = launchFusedSteps(s, nsrc, 8, 20, s->hist, &launches);
    s->timeline = d;
    if (!rc) rc = launchFusedSteps(s, nsrc, 20, 24, s->hist, &launches);
    s->timeline = nullptr;
    if (!rc && (cudaMemcpyAsync(out, d, sizeof(unsigned long long) * 8 * blocks, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess)) { setError("pvc_debug_timeline: copy failed"); rc = PVC_ERR_CUDA; }
    cudaFree(d);
    return rc ? rc : blocks;
}

int pvc_debug_ws2_item(int w, int gen_chunk, int src_group, int num_gen, int nsrc, int tiles_per_source, int* out3)
{
    if (!out3 || gen_chunk < 1 || src_group < 1 || src_group > nsrc || num_gen < 1 || nsrc < 1 || tiles_per_source < 1 ||
        w < 0 || (long long)w >= (long long)num_gen * tiles_per_source * nsrc) { setError("pvc_debug_ws2_item: bad argument"); return PVC_ERR_INVALID; }
    const pvc::Ws2Order ord = { gen_chunk, src_group, num_gen, nsrc, tiles_per_source, tiles_per_source * nsrc, 0, 0 };
    const pvc::Ws2Item it = pvc::ws2DecodeItem(w, ord);
    out3[0] = it.s; out3[1] = it.gen; out3[2] = it.o;
    return PVC_OK;
}

int pvc_debug_ws2_item_banded(int w, int gen_chunk, int src_group, int num_gen, int nsrc, int tiles_x, int tiles_y, int band, int* out3)
{
    if (!out3 || gen_chunk < 1 || src_group < 1 || src_group > nsrc || num_gen < 1 || nsrc < 1 || tiles_x < 1 || tiles_y < 1 || band < 1 ||
        w < 0 || (long long)w >= (long long)num_gen * tiles_x * tiles_y * nsrc) { setError("pvc_debug_ws2_item_banded: bad argument"); return PVC_ERR_INVALID; }
    const pvc::Ws2Order ord = { gen_chunk, src_group, num_gen, nsrc, tiles_x * tiles_y, tiles_x * tiles_y * nsrc, band, tiles_x };
    const pvc::Ws2Item it = pvc::ws2DecodeItem(w, ord);
    out3[0] = it.s; out3[1] = it.gen; out3[2] = it.o;
    return PVC_OK;
}

int pvc_set_walk_mode(pvc_solver* s, int sequential)
{
    if (!s) { setError("pvc_set_walk_mode: null solver"); return PVC_ERR_INVALID; }
    s->walkSequential = sequential ? 1 : 0;
    return PVC_OK;
}

int pvc_step_variant(pvc_solver* s) { return s ? s->cfg.reserved : -1; }

void* pvc_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { setError("pvc_host_alloc: %s", cudaGetErrorString(cudaGetLastError())); return nullptr; }
    return p;
}

void pvc_host_free(void* p) { if (p) cudaFreeHost(p); }

void* pvc_stream(pvc_solver* s) { return s ? (void*)s->stream : nullptr; }

} // extern "C"

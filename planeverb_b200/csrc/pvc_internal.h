// pvc_internal.h -- device data layout shared by the kernels of the thin CUDA layer.
//
// HBM layout (all fp32, one allocation per kind):
//   state planes   p, vx, vy   [2 ping-pong][source][rows_alloc][pitch]
//       Row r / column c of the reference's alloc grid (FDTD.cpp:99) lives at
//       (r + kGuardRows) * pitch + (c + kGuardCols).  The guard band (kGuardRows rows above,
//       kGuardCols columns to the left, tile overrun below/right) is zero and never written, so the
//       temporally blocked step kernel loads whole tiles + halo without bounds checks.
//   coefficient plane  w  [rows_alloc][pitch], shared by all sources:
//       bit pattern kAirBits  -> air cell (reference b = 1)
//       anything else         -> wall cell (b = 0) and the float is its admittance
//                                Y = (1-R)/(1+R) (FDTD.cpp:153,160); guard cells are walls with Y = 0.
//   pressure history   hist [source][row][chunk][t][128]: sample t of alloc cell (r, c) lives at
//       ((r*hist_chunks + c/W)*T + t)*W + c%W -- time-major inside each W-column strip (W = 128; 120 for the
//       TMA-store step variant, where a strip is the owned columns of one tile), so the
//       step kernel appends 512 contiguous bytes per warp-row per step and the analyzer, which walks
//       every cell through time, reads each strip as ONE sequential stream.  It is the only per-step
//       record kept (4 B per cell-step instead of the reference's 16-byte Cell, FDTD.cpp:226-231);
//       vx/vy of any sample are rebuilt from it.
//   results  [source][gx*gy][8], delay [source][gx*gy]   (Analyzer.h:13-21, Analyzer.cpp:40)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/planeverb_cuda.h"

namespace pvc
{
    constexpr int kGuardRows = 4;          // = max temporal block depth K
    constexpr int kGuardCols = 4;          // one float4
    constexpr uint32_t kAirBits = 0xFFFFFFFFu;

    // tile geometry of the fused step kernel (pvc_step_fused.cu)
    constexpr int kTileK = 4;              // time steps per launch
    constexpr int kTileCols = 128;         // one warp wide: 32 lanes x float4
    constexpr int kValidCols = kTileCols - 2 * kGuardCols;   // 120 output columns per tile

    struct Layout
    {
        int gx, gy;            // interior cells
        int rows, cols;        // alloc grid: gx+1, gy+1
        int pitch;             // floats per state row (multiple of 32)
        int rows_alloc;        // state rows incl. guards
        size_t plane;          // rows_alloc * pitch
        int T;                 // samples per impulse response
        int hist_chunk;        // columns per history strip (128, or 120 for the TMA-store step variant)
        int hist_chunks;       // strips per row: ceil(cols / hist_chunk)
        size_t hist_row;       // floats between consecutive rows of the history: hist_chunks * T * hist_chunk
        size_t hist_source;    // floats between sources: rows * hist_row
        int tiles_x, tiles_y;  // fused-kernel tile grid
        int tile_rows;         // rows per tile incl. halo (warps * rows per thread)
        int valid_rows;        // tile_rows - 2*kTileK
        int warp_rows;         // rows per warp (R) of the fused-kernel variant in use
    };

    __host__ __device__ inline size_t cellIndex(const Layout& L, int r, int c)
    {
        return (size_t)(r + kGuardRows) * L.pitch + (c + kGuardCols);
    }

    // history strip width (Layout::hist_chunk): 128 columns (one aligned 512-byte row per sample and strip) for every kernel
    // that records with per-thread stores; 120 = the owned columns of one tile for the TMA-store variant, whose record of one
    // sample must be one dense box
    constexpr int kHistChunkDefault = 128;
    constexpr int kNeverActive = 0x7f7f7f7f;     // memset(0x7f) pattern of firstActive
    constexpr int kCarryPlanes = 10;             // streamed solve: onset, edry, fx, fy, vx, vy, wet, edc, xysum, ysum per cell

    // float offset of sample 0 of alloc cell (r, c) inside one source's history; sample t is + t*hist_chunk
    __host__ __device__ inline size_t histCell(const Layout& L, int r, int c)
    {
        return (size_t)r * L.hist_row + (size_t)(c / L.hist_chunk) * L.T * L.hist_chunk + (c % L.hist_chunk);
    }

    struct SourceParams       // one listener, device copy of pvc_listener plus derived indices
    {
        int cell_r, cell_c;
        int efree_r, efree_c;
        float x, z;
        int dead;             // the pulse cell is not an interior air cell (wall, padding row / column): filled on the device by
                              // markDeadSourcesKernel after the geometry of the frame has been applied.  The reference zeroes a
                              // sample injected there in the next pressure sub-step before anything reads it (FDTD.cpp:125-141,
                              // :234); the resident kernel's multiply-free wall rule relies on p == 0 in such cells and skips it
    };
}

namespace pvc
{
    constexpr int kMaxGraphBatch = 16;
    struct GraphSlot { cudaGraphExec_t exec; int finalCur, launches, seen; };
}

struct pvc_solver
{
    pvc_config cfg;
    pvc::Layout L;
    int device;
    int numSMs;
    cudaStream_t stream;
    cudaEvent_t ev[4];
    cudaEvent_t mark[2];
    cudaStream_t copyStream; // pipelined result fetch (pvc_fetch_results_async)
    cudaEvent_t evAnalyzed, evCopied;
    int copyPending;
    int* hostAbort;          // pinned: abort flag of the run whose results are being fetched

    float* stateBlock;       // the six state arrays, one allocation
    float* state[2][3];      // [pingpong][p,vx,vy] each max_sources * plane floats (slices of stateBlock)
    float* w;                // wall plane (air flag / admittance), the geometry's source of truth
    float* coef[3];          // general-path coefficient planes bp, gx, gy derived from w (pvc_step_fused.cu)
    float* lin[3];           // linear-form coefficient planes cP, sX, sY derived from w (pvc_step_res.cu); null until a resident variant needs them
    void* resXchg;           // resident kernel: mailbox planes [8 slots][max_sources][rows_alloc][pitch] of 16-byte {p, vx, vy, tag} words
    int resEpoch;            // resident kernel: tag epoch of the last solve (15 bits)
    int resSourcesPerLaunch; // resident kernel: sources solved concurrently by one launch (co-residency limit), 0 = not computed yet
    uint32_t* slowMask;      // per (tile, warp): lanes that must take the general (wall/edge) path
    uint32_t* bpMask;        // per (tile, warp, lane): air flags of the thread's cells (pvc_step_ws2.cu)
    int slowMaskDirty;
    int* tileOrder;          // tiles_x*tiles_y tile ids: most expensive first, or row-major (tileOrderNatural)
    int tileOrderNatural;
    int* firstActive;        // [source][tile][32]: activity hints written by the fused kernels, read by the analyzer
    int hintsValid;          // the last run's step kernel filled firstActive
    int* tileCounters;       // one work counter per launch of the persistent TMA variants (slot 0: abort flag of the generational one)
    int* doneGen;            // generational variant: completed generations per (source, tile)
    int checkAbort;
    int tileCounterCount;
    alignas(64) unsigned char tensorMaps[6 * 128];   // CUtensorMap[2 ping-pong][3 fields] (128 B each)
    int tmaReady, tmaTileRows;
    float* hist;             // max_sources * T * hist_plane
    float* pulse;            // T floats
    float* results;          // max_sources * gx*gy*8
    float* delay;            // max_sources * gx*gy
    float* walkDelay;        // max_sources * gx*gy   (delay of selectable cells, FLT_MAX otherwise)
    int* walkNext;           // max_sources * gx*gy   (links of the listener-direction walk, pvc_analyze.cu)
    float* scratch;          // small device scratch (IR fetch)
    pvc_rect* rects;         // device copy of the current geometry edit list
    int rectCapacity;
    pvc_rect* rectsHost;     // pinned staging, 2 x rectCapacity (alternating halves): applying edits never waits for the stream
    cudaEvent_t rectCopied[2];
    unsigned rectSlot;
    pvc::SourceParams* src;  // max_sources
    cudaEvent_t gathered[2];        // pvc_gather_results_async tickets
    unsigned gatherSlot;
    pvc::SourceParams* srcHost;     // pinned ring of kSrcRing x max_sources staging slots: pvc_run never waits for the stream
    cudaEvent_t srcCopied[4];
    unsigned srcSlot;
    float efree;
    int cur;                 // ping-pong index holding the latest state
    int lastSources;
    float lastMs[3];
    int lastLaunches;
    int lastStepLaunches;
    unsigned long long* timeline;   // debug only
    int useGraphs;
    int walkSequential;      // pvc_set_walk_mode: 1 = listener direction by the reference's sequential walk (cross-check)
    pvc::GraphSlot graphs[pvc::kMaxGraphBatch + 1];   // captured step-launch sequences, by batch size
    // streamed solve (pvc_create_streamed): the history holds chunkT samples only.  Forward sweep chunk by chunk (state checkpointed
    // at every chunk start, causal analyzer sums carried per cell), then the chunks are RECOMPUTED from their checkpoints in
    // reverse order for the backward Schroeder pass -- the kernels are deterministic, so every output stays bit-exact.
    int chunkT;              // samples of history kept (0: the whole response, L.T == cfg.T); L.T == chunkT otherwise
    int finalPass;           // the chunk being stepped ends the response
    int stateStale;          // the state planes hold the end of chunk 0, not of the response (after the backward sweep)
    int abortSticky;         // step launches leave the abort flag alone (every chunk of a streamed solve but the first)
    float* ckpt;             // state at the start of chunks 1 .. K-2: [K-2][3 fields][max_sources * plane]
    float* carry;            // analyzer state carried across chunks: [pvc::kCarryPlanes][max_sources * gx*gy]
};

namespace pvc
{
    // ---- work-item order of the generational step kernel (pvc_step_ws2.cu) --------------------------------------------
    // One launch covers numGen generations of every (source, tile).  Work item w of 0 .. numGen * tps * nsrc - 1 is:
    // chunks of genChunk generations; inside a chunk one source group (srcGroup sources whose ping-pong state fits the L2)
    // after the other; inside a group generation-major, then position in the tile order, then source.  Every dependency of
    // an item -- same source, the tile and its up-to-8 neighbours, previous generation -- precedes it in this order, which
    // is what makes the kernel's "pull the next item, wait for its dependencies" loop deadlock-free with co-resident CTAs.
    // Shared by the kernel and the host (pvc_debug_ws2_item -> tests/test_abi.py checks the invariant on the CPU).
    // Grids whose ping-pong state does not fit the L2 even for one source (2048^2: 100 MB) add a level: inside a chunk the tile rows
    // are split into BANDS of `band` rows that shift up by one tile row per generation (band k, generation g of the chunk: rows
    // k*band - g .. (k+1)*band - g - 1, clipped to the grid, the last band reaching its end), and a band runs ALL the chunk's
    // generations before the next band starts -- a band's state then stays in the L2 for the whole chunk instead of streaming
    // through HBM every generation.  The skew is what keeps it a dependency-respecting order: item (g, row r) needs (g-1, r-1 .. r+1),
    // which lie in the same band one generation earlier or in an earlier band.  band == 0: no bands (tx unused).
    struct Ws2Order { int genChunk, srcGroup, numGen, nsrc, tps, numTiles, band, tx; };      // numTiles = tps * nsrc; tx = tiles per tile row
    struct Ws2Item { int s, gen, o; };                  // source, generation relative to the launch, position in the tile order
#if defined(__CUDACC__)
    __host__ __device__
#endif
    inline Ws2Item ws2DecodeItem(int w, const Ws2Order& P)
    {
        const int chunkItems = P.genChunk * P.numTiles;
        const int c = w / chunkItems;
        int rem = w - c * chunkItems;
        const int left = P.numGen - c * P.genChunk;
        const int gc = P.genChunk < left ? P.genChunk : left;
        const int groupItems = gc * P.tps * P.srcGroup;
        const int q = rem / groupItems;
        rem -= q * groupItems;
        const int rest = P.nsrc - q * P.srcGroup;
        const int sq = P.srcGroup < rest ? P.srcGroup : rest;
        if (P.band > 0)
        {
            const int ty = P.tps / P.tx, nb = (ty + P.band - 1) / P.band, per = P.tx * sq;      // per: items of one tile row
            int k = 0, g = 0, lo = 0;
            bool found = false;
            for (k = 0; k < nb && !found; ++k)
                for (g = 0; g < gc; ++g)
                {
                    int a = k * P.band - g, b = (k == nb - 1) ? ty : (k + 1) * P.band - g;
                    a = a < 0 ? 0 : (a > ty ? ty : a);
                    b = b < 0 ? 0 : (b > ty ? ty : b);
                    const int n = (b - a) * per;
                    if (rem < n) { lo = a; found = true; break; }
                    rem -= n;
                }
            const int r = lo + rem / per;
            rem -= (rem / per) * per;
            const int col = rem / sq;
            Ws2Item bi;
            bi.s = q * P.srcGroup + (rem - col * sq);
            bi.gen = c * P.genChunk + g;
            bi.o = r * P.tx + col;
            return bi;
        }
        const int g = rem / (sq * P.tps);
        rem -= g * (sq * P.tps);
        const int o = rem / sq;
        Ws2Item it;
        it.s = q * P.srcGroup + (rem - o * sq);
        it.gen = c * P.genChunk + g;
        it.o = o;
        return it;
    }


#ifdef __CUDACC__
    // ---- barrier-free row exchange between the warps of a tile (both register-tiled step kernels) ---------------------------
    // A warp hands its boundary row (4 floats per lane) to the warp above / below through shared memory.  Instead of meeting
    // the neighbour on a barrier, the row travels with a tag in every 8-byte unit -- {v0, tag, v1, tag}, {v2, tag, v3, tag}: an
    // aligned 8-byte shared-memory access is indivisible -- and the reader polls the row itself (the low-latency protocol of the
    // resident kernel's mailbox, between warps).  The reader first updates the rows that do not need the neighbour's data, so
    // a neighbour that is a little late costs nothing; with barriers every sub-step waited for the slower of three warps twice.
    // Tags count the rows a warp has published; every warp of a CTA publishes the same sequence, so the expected tag is the
    // reader's own count.  A row buffer is rewritten only after the writer has consumed a row the reader published AFTER
    // reading it (the two directions alternate), except where noted at the call sites.
    namespace flow
    {
        constexpr int kRowFloat4 = 64;        // one exchanged row: [2][32] float4
        __device__ __forceinline__ void publish(float4* row, const int lane, const float v0, const float v1, const float v2, const float v3, const int tag)
        {
            const unsigned a = (unsigned)__cvta_generic_to_shared(row + lane);
            const float t = __int_as_float(tag);
            asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v0), "f"(t), "f"(v1), "f"(t) : "memory");
            asm volatile("st.volatile.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a + 512u), "f"(v2), "f"(t), "f"(v3), "f"(t) : "memory");
        }
        __device__ __forceinline__ float4 poll(const float4* row, const int lane, const int tag)
        {
            const unsigned a = (unsigned)__cvta_generic_to_shared(row + lane);
            float4 lo, hi;
            bool ok;
            do
            {
                asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w) : "r"(a) : "memory");
                asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "r"(a + 512u) : "memory");
                ok = ((__float_as_int(lo.y) ^ tag) | (__float_as_int(lo.w) ^ tag) | (__float_as_int(hi.y) ^ tag) | (__float_as_int(hi.w) ^ tag)) == 0;
                if (__all_sync(0xffffffffu, ok)) break;
                __nanosleep(20);          // a spinning warp takes issue slots from the warps it waits for (1.5 % of the 1024^2 step phase)
            } while (true);
            return make_float4(lo.x, lo.z, hi.x, hi.z);
        }
    }
#endif
    // step kernels (pvc_step.cu / pvc_step_fused.cu)
    int launchBaselineSteps(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches);
    int launchFusedSteps(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches);
    int rebuildSlowMask(pvc_solver* s);
    int launchWs2Steps(pvc_solver* s, int variant, int nsrc, int t0, int t1, float* hist, int* launches);
    int rebuildWs2Descriptors(pvc_solver* s, int variant);
    int launchResidentSteps(pvc_solver* s, int variant, int nsrc, int t0, int t1, float* hist, int* launches);
    int rebuildResidentDescriptors(pvc_solver* s, int variant);
    int residentCapacity(int variant, int device);     // CTAs of a resident variant that can be co-resident on the device (0: unknown / does not fit)
    int markDeadSources(pvc_solver* s, int n);
    bool variantAvailable(int variant);
    int variantKind(int variant);
    int variantMinBlocks(int variant);
    int variantWarps(int variant);
    int buildTensorMaps(pvc_solver* s);
    int fusedTileRows(int variant);
    int fusedWarpRows(int variant);
    int fusedHistChunk(int variant);
    // analyzer kernels (pvc_analyze.cu)
    int launchAnalyzer(pvc_solver* s, int nsrc, int* launches);
    // streamed solve: causal sums over samples base .. base+len-1 (in the history as samples 0 .. len-1), carried in s->carry;
    // backward Schroeder sums over the same chunk (chunk 0 comes last and writes the results); listener-direction kernels
    int launchStreamForward(pvc_solver* s, int nsrc, int base, int len, int* launches);
    int launchStreamBackward(pvc_solver* s, int nsrc, int base, int len, int* launches);
    int launchListenerDirection(pvc_solver* s, int nsrc, int* launches);
    int launchIrRebuild(pvc_solver* s, int source, int r, int c, float* out_dev);
    // geometry kernels (pvc_geometry.cu)
    int launchClearGeometry(pvc_solver* s);
    int launchApplyRects(pvc_solver* s, const pvc_rect* rects_host, int n);
    void setError(const char* fmt, ...);
}

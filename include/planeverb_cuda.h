/*
 * planeverb_cuda.h -- the thin C-ABI CUDA layer behind Planeverb's hot path.
 *
 * This is the internal seam SURVEY.md section 8(b) names: the reference's
 *     Grid::GenerateResponse(listener)      ProjectPlaneverb/src/FDTD/Grid.h:32-34, FDTD.cpp:244-254
 *     Grid::GenerateResponseGPU (throws)    ProjectPlaneverb/src/FDTD/FDTD.cpp:238-242   <- the stub this implements
 *     Analyzer::AnalyzeResponses(listener)  ProjectPlaneverb/src/DSP/Analyzer.h:30, Analyzer.cpp:48-104
 *     Grid::AddAABB / RemoveAABB            ProjectPlaneverb/src/FDTD/Grid.cpp:229-296
 *     FreeGrid::SimulateFreeFieldEnergy     ProjectPlaneverb/src/FDTD/FreeGrid.cpp:71-94
 *     Grid::GetResponse                     ProjectPlaneverb/src/FDTD/FDTD.cpp:74-79
 * selected when PlaneverbConfig::threadExecutionType == 1 (the commented-out pv_GPU,
 * ProjectPlaneverb/include/PvTypes.h:13-17).
 *
 * Plain C: opaque handle, plain pointers and sizes, int status codes (0 = ok).  All pointers are
 * HOST pointers unless the name ends in _dev.  Every scalar that the reference derives through a
 * float->int truncation (grid size, response length, listener cell, AABB cell rectangle, analysis
 * window lengths) is computed by the HOST caller with the reference's own expressions
 * (planeverb_b200/csrc/pv_params.h) and handed in here as integers, so the device never re-derives
 * an index.  Host-side consumers: planeverb_b200/csrc/planeverb_api.cpp (the Planeverb C++ API +
 * Unity C ABI) and planeverb_b200/pvcuda.py (ctypes, tests and bench).
 *
 * Layout contract: a "plane" is (gx+1) rows of (gy+1) floats, row-major, index r*(gy+1)+c -- the
 * alloc-grid indexing of FDTD.cpp:99.  Results use the analyzer's interior indexing s = r*gy + c (the reference
 * strides by gx, PvDefinitions.h:23-24 / Analyzer.cpp:79 -- identical on the square grids it supports), 8 floats per cell in AnalyzerResult order
 * (Analyzer.h:13-21): occlusion, wetGain, rt60, lowpass, direction.x, direction.y,
 * sourceDirectivity.x, sourceDirectivity.y.
 */
#ifndef PLANEVERB_CUDA_H
#define PLANEVERB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PVC_API __declspec(dllexport)
#else
#define PVC_API __attribute__((visibility("default")))
#endif

typedef struct pvc_solver pvc_solver;

enum pvc_status
{
    PVC_OK = 0,
    PVC_ERR_INVALID = 1,    /* bad argument            (reference: throw pv_InvalidConfig)   */
    PVC_ERR_MEMORY = 2,     /* device allocation failed (reference: throw pv_NotEnoughMemory) */
    PVC_ERR_CUDA = 3,       /* CUDA runtime error, see pvc_last_error()                       */
    PVC_ERR_NO_DEVICE = 4   /* no usable CUDA device: the product has NO CPU fallback         */
};

/* Everything Grid's constructor and Analyzer's constructor derive (Grid.cpp:46-55, Analyzer.cpp:20-27,
 * 170-171,237,284), already evaluated on the host. */
typedef struct pvc_config
{
    int   gx, gy;            /* interior cells (int)m_gridSize.x/.y                               */
    int   T;                 /* response length in samples (m_responseLength)                     */
    int   fs;                /* sampling rate (m_samplingRate)                                    */
    int   resolution;        /* gridResolution                                                    */
    float dx;                /* metres per cell                                                   */
    float courant;           /* PV_C*dt/dx as evaluated at FDTD.cpp:90                            */
    int   flux_samples;      /* (int)(PV_DRY_DIRECTION_ANALYSIS_LENGTH*fs)   Analyzer.cpp:171     */
    int   dry_samples;       /* (int)(PV_DRY_GAIN_ANALYSIS_LENGTH*fs)        Analyzer.cpp:170     */
    int   wet_samples;       /* (int)(PV_WET_GAIN_ANALYSIS_LENGTH*fs)        Analyzer.cpp:237     */
    int   tail_samples;      /* (int)(PV_SCHROEDER_OFFSET_S*fs)              Analyzer.cpp:284     */
    int   max_sources;       /* listener positions solved per pvc_run call (batched on one GPU)   */
    int   device;            /* CUDA device ordinal                                               */
    int   step_kernel;       /* 0 = default (fused, temporally blocked), 1 = two-launch baseline  */
    int   reserved;          /* fused-kernel tile variant (0 = default); tuning knob, see pvc_step_fused.cu */
} pvc_config;

/* One queued geometry edit, already voxelised to a half-open cell rectangle rows [r0,r1) x cols
 * [c0,c1) by the host with the truncations of Grid.cpp:139-142 / 252-255 (clipping to 0..gx / 0..gy
 * inclusive happens on the device as in Grid.cpp:231,235). add != 0: wall with admittance
 * Y = (1-R)/(1+R) (FDTD.cpp:153,160); add == 0: back to air (Grid.cpp:260-295). Applied in order. */
typedef struct pvc_rect
{
    int   r0, r1, c0, c1;
    int   add;
    float admittance;
} pvc_rect;

/* One listener ("source" in BASELINE.json): pulse cell (FDTD.cpp:97-99), the analyzer's own listener
 * cell (Analyzer.cpp:201-202 -- multiply-by-reciprocal, can differ by one) and its world position. */
typedef struct pvc_listener
{
    int   cell_r, cell_c;
    int   efree_r, efree_c;
    float x, z;
} pvc_listener;

PVC_API int  pvc_device_count(void);
/* free / total bytes of device memory (cudaMemGetInfo) */
PVC_API int  pvc_device_memory(int device, size_t* free_bytes, size_t* total_bytes);
PVC_API const char* pvc_last_error(void);

PVC_API int  pvc_create(const pvc_config* cfg, pvc_solver** out);
PVC_API void pvc_destroy(pvc_solver* s);

/* bytes of device memory pvc_create would allocate for cfg (pressure history dominates: 4*T*plane per source) */
PVC_API size_t pvc_memory_requirement(const pvc_config* cfg);

/* Streamed solve for responses whose pressure history (4 bytes x cells x T per source) does not fit the device: the history
 * holds history_steps samples (rounded down to a multiple of 8) and the response is solved in K = ceil(T / history_steps)
 * chunks.  Forward sweep: every chunk's time steps followed by the causal part of the analysis (onset, dry energy, flux, wet
 * energy: Analyzer.cpp:146-247) with the running sums carried per cell; the state is checkpointed at every chunk start.
 * Backward sweep: the chunks are recomputed from their checkpoints in reverse order for the anti-causal Schroeder integral and
 * regression (Analyzer.cpp:282-326).  Same results as pvc_create's solver, bit for bit, at (2 - 1/K) x the time steps and 1/K
 * of the history memory.  Runs on the resident and the generational step kernels (not on the plain fallback 18); pvc_fetch_ir /
 * pvc_fetch_pressure are not available (there is no full history), pvc_fetch_state only after a run with analyze == 0. */
PVC_API int  pvc_create_streamed(const pvc_config* cfg, int history_steps, pvc_solver** out);
PVC_API size_t pvc_memory_requirement_streamed(const pvc_config* cfg, int history_steps);

/* Gaussian source pulse, T floats (Grid.cpp:12-27, evaluated by the host) */
PVC_API int  pvc_set_pulse(pvc_solver* s, const float* pulse, int n);

/* reset the coefficient plane to the empty grid of Grid.cpp:84-108 */
PVC_API int  pvc_clear_geometry(pvc_solver* s);
/* apply n queued edits in order on the device (GeometryManager::PushGeometryChanges, GeometryManager.cpp:123-152) */
PVC_API int  pvc_apply_geometry(pvc_solver* s, const pvc_rect* rects, int n);
/* read back the coefficient plane as the reference's two fields: b (0/1) and admittance Y (tests) */
PVC_API int  pvc_fetch_coefficients(pvc_solver* s, int16_t* b, float* admittance);

/* FreeGrid (FreeGrid.cpp:71-110): n-step free-field run on an empty grid from (lr,lc); returns
 * r * sum_{i<n} p_i^2 at (er,ec). The result is also installed as the solver's EFree. */
PVC_API int  pvc_compute_efree(pvc_solver* s, int lr, int lc, int er, int ec, int n, float r, float* efree);
PVC_API int  pvc_set_efree(pvc_solver* s, float efree);

/* GenerateResponse + AnalyzeResponses for n listeners (n <= max_sources), asynchronous on the
 * solver's stream; results stay on the device until fetched. analyze == 0 skips the analyzer. */
PVC_API int  pvc_run(pvc_solver* s, const pvc_listener* listeners, int n, int analyze);
PVC_API int  pvc_synchronize(pvc_solver* s);

/* zero the persistent results of one source slot (Context's memset, PvContext.cpp:132) */
PVC_API int  pvc_clear_results(pvc_solver* s, int source);
/* results: gx*gy*8 floats, delay: gx*gy floats (FLT_MAX = no onset); either may be NULL. Synchronous. */
PVC_API int  pvc_fetch_results(pvc_solver* s, int source, float* results, float* delay);
/* Pipelined form for frame loops: enqueue the copy of the result grids of sources 0..n-1 of the LAST pvc_run (results:
 * n*gx*gy*8 floats, delay: n*gx*gy floats, either may be NULL; use pvc_host_alloc'd memory) on the solver's copy stream and
 * return at once.  The copy starts when that run's analyzer has finished and overlaps the time steps of the next pvc_run, whose
 * analyzer in turn waits for the copy before it overwrites the device grids.  pvc_fetch_wait blocks until the host buffers
 * are filled (and reports a failed run).  One copy in flight at a time. */
PVC_API int  pvc_fetch_results_async(pvc_solver* s, int n, float* results, float* delay);
PVC_API int  pvc_fetch_wait(pvc_solver* s);
/* the 8 floats of one interior cell (Analyzer::GetResponseResult, Analyzer.cpp:106-116) */
PVC_API int  pvc_fetch_result_at(pvc_solver* s, int source, int r, int c, float* out8);
/* Frame-loop form of pvc_fetch_result_at: enqueue on the solver's stream -- after the last pvc_run, before the next -- the
 * copy of the 8 floats of n_cells interior cells (cells[i] = r * gy + c) of sources 0..n-1 into out (n * n_cells * 8
 * floats, source-major; pvc_host_alloc'd memory; a negative cell index leaves its 8 floats untouched) and return at once
 * with a ticket (0 or 1).  pvc_gather_wait blocks until
 * that copy has landed.  Two gathers may be in flight, so frame k's emitter outputs can be consumed (and exchanged between
 * GPUs) while frame k+1 is being solved. */
PVC_API int  pvc_gather_results_async(pvc_solver* s, int n, const int* cells, int n_cells, float* out, int* ticket);
PVC_API int  pvc_gather_wait(pvc_solver* s, int ticket);
/* impulse response of alloc cell (r,c): T x {p, vx, vy} (Grid::GetResponse). vx/vy are rebuilt on the
 * device from the pressure history with the solver's own update rules. */
PVC_API int  pvc_fetch_ir(pvc_solver* s, int source, int r, int c, float* out3T);
/* pressure plane recorded for sample t (tests): (gx+1)*(gy+1) floats */
PVC_API int  pvc_fetch_pressure(pvc_solver* s, int source, int t, float* plane);
/* final p/vx/vy planes after the last step, before the last injection (tests) */
PVC_API int  pvc_fetch_state(pvc_solver* s, int source, float* p, float* vx, float* vy);

/* timing of the last pvc_run in milliseconds, measured with CUDA events on the solver's stream:
 * out[0] = step kernels, out[1] = analyzer kernels, out[2] = total; launches = kernels launched */
PVC_API int  pvc_last_timing(pvc_solver* s, float* out3, int* launches);
/* kernels launched by the last pvc_run, split into the time-step phase (zeroing + step kernels) and the analyzer phase */
PVC_API int  pvc_last_launch_counts(pvc_solver* s, int* step_launches, int* analyzer_launches);
/* bracket any sequence of calls with two CUDA events on the solver's stream (which = 0 start, 1 stop);
 * pvc_mark_elapsed waits for the stop event and returns the milliseconds between them */
PVC_API int  pvc_mark(pvc_solver* s, int which);
PVC_API int  pvc_mark_elapsed(pvc_solver* s, float* ms);
/* device-resident raw pointers for zero-copy consumers (results_dev: gx*gy*8 floats of a source) */
PVC_API const float* pvc_results_dev(pvc_solver* s, int source);
/* debug/profiling aid: runs a few 4-step launches and returns, per CTA of the last one, 8 %globaltimer stamps
 * (start, tile loaded, after each of the 4 steps, stores issued); returns the number of CTAs or a negative... >= 0 ok */
PVC_API int  pvc_debug_timeline(pvc_solver* s, int nsrc, unsigned long long* out, int maxBlocks);
/* host-side mirror of the generational step kernel's work-item order (pvc_internal.h::ws2DecodeItem, the same inline function
 * the kernel calls): item w of 0 .. num_gen * tiles_per_source * nsrc - 1 -> out3 = { source, generation, position in the tile
 * order }.  Host arithmetic, needs no GPU; tests/test_abi.py checks that the order is a bijection in which every dependency of
 * an item (same source, previous generation) precedes it -- the kernel's deadlock-freedom argument.  PVC_ERR_INVALID if out of range. */
PVC_API int  pvc_debug_ws2_item(int w, int gen_chunk, int src_group, int num_gen, int nsrc, int tiles_per_source, int* out3);
/* the same with the tile rows of a chunk split into bands of `band` rows (the order used on grids whose state exceeds the L2,
 * pvc_internal.h::Ws2Order); out3[2] = tile_row * tiles_x + tile_column */
PVC_API int  pvc_debug_ws2_item_banded(int w, int gen_chunk, int src_group, int num_gen, int nsrc, int tiles_x, int tiles_y, int band, int* out3);
/* Listener-direction algorithm of the analyzer (Analyzer::EncodeListenerDirection, Analyzer.cpp:340-431): 0 (default) = pointer
 * jumping over a link array, 1 = the reference's walk, one thread per start cell.  Bit-identical results; the second exists as the
 * cross-check of the first (tests/test_gpu_parity.py). */
PVC_API int  pvc_set_walk_mode(pvc_solver* s, int sequential);
/* the step-kernel variant this solver runs (pvc_config::reserved after "0 = auto" has been resolved; see pvc_api.cu::resolveVariant) */
PVC_API int  pvc_step_variant(pvc_solver* s);
/* page-locked host buffers for the result grids (plain malloc'd memory works too, just slower to copy) */
PVC_API void* pvc_host_alloc(size_t bytes);
PVC_API void  pvc_host_free(void* p);
PVC_API void* pvc_stream(pvc_solver* s);

#ifdef __cplusplus
}
#endif
#endif

/*
 * planeverb_ext.h -- "scene solver" entry points: the reference's Grid + FreeGrid + Analyzer trio
 * driven directly (no Context background thread), with HOST buffers in and out.
 *
 * Mirrors, call for call, what a host program does with the reference classes
 *     Grid::Grid / AddAABB / RemoveAABB      ProjectPlaneverb/src/FDTD/Grid.cpp:30-117,136-296
 *     FreeGrid::FreeGrid                     ProjectPlaneverb/src/FDTD/FreeGrid.cpp:6-34
 *     Grid::GenerateResponse                 ProjectPlaneverb/src/FDTD/FDTD.cpp:244-254
 *     Analyzer::AnalyzeResponses             ProjectPlaneverb/src/DSP/Analyzer.cpp:48-104
 *     Analyzer::GetResponseResult            ProjectPlaneverb/src/DSP/Analyzer.cpp:106-116
 * plus the two things BASELINE.json's configs need that the reference's Context does not offer:
 * an explicit response length (500/2000/4000 steps instead of the derived fs*0.3015 s) and several
 * listener positions ("sources") solved as one batch.  The Planeverb C++ API / Unity C ABI
 * (include/Planeverb.h, include/PlaneverbUnity.h) sits on the same code for the drop-in path; this
 * header is what tests/ and bench.py bind (planeverb_b200/pvcuda.py).  Positions are world metres on
 * the (x, z) plane exactly as in Planeverb::vec3 usage (FDTD.cpp:97-98).
 */
#ifndef PLANEVERB_EXT_H
#define PLANEVERB_EXT_H

#include "planeverb_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvx_scene pvx_scene;

/* responseLength <= 0: the reference's derived length (Grid.cpp:55). efree < 0: run the free-field
 * simulation on the device (FreeGrid.cpp:71-94). stepKernel/variant: see pvc_config. */
PVC_API int  pvx_create(float sizeX, float sizeY, int resolution, int responseLength, float efree,
                        int maxSources, int device, int stepKernel, int variant, pvx_scene** out);
/* the same scene on a streamed solver (pvc_create_streamed): the pressure history holds historySteps samples and the response is
 * solved in chunks, for grids / response lengths whose full history does not fit the device.  historySteps == 0: pvx_create;
 * historySteps < 0: automatic -- the full history if it fits 90 % of the device's free memory, else the longest that does. */
PVC_API int  pvx_create_streamed(float sizeX, float sizeY, int resolution, int responseLength, float efree,
                                 int maxSources, int device, int stepKernel, int variant, int historySteps, pvx_scene** out);
/* samples the scene's history holds: 0 = the whole response (full-history solver), > 0 = streamed solver */
PVC_API int  pvx_history_steps(pvx_scene* sc);
PVC_API void pvx_destroy(pvx_scene* sc);

/* ints[10] = gx, gy, T, fs, fluxSamples, drySamples, wetSamples, tailSamples, freeSamples, maxSources
 * floats[4] = dx, dt, courant, efree */
PVC_API int  pvx_info(pvx_scene* sc, int* ints, float* floats);
PVC_API int  pvx_pulse(pvx_scene* sc, float* out, int n);

/* queue an AddAABB / RemoveAABB (Grid.cpp:136,249); flushed in order before the next solve, like
 * GeometryManager::PushGeometryChanges (GeometryManager.cpp:123-152) */
PVC_API int  pvx_add_aabb(pvx_scene* sc, float posX, float posY, float width, float height, float absorption);
PVC_API int  pvx_remove_aabb(pvx_scene* sc, float posX, float posY, float width, float height, float absorption);
PVC_API int  pvx_flush_geometry(pvx_scene* sc);

/* GenerateResponse + AnalyzeResponses for n listeners given as xyz triples (y ignored).
 * results (n*gx*gy*8 floats) / delay (n*gx*gy floats) may be NULL to leave them on the device.
 * Synchronous: returns when the host buffers are filled. */
PVC_API int  pvx_solve(pvx_scene* sc, const float* listenersXYZ, int n, int analyze, float* results, float* delay);
/* asynchronous halves for overlapped multi-device use */
PVC_API int  pvx_solve_async(pvx_scene* sc, const float* listenersXYZ, int n, int analyze);
PVC_API int  pvx_wait(pvx_scene* sc);
/* frame-loop form: solve + enqueue the copy of this solve's result grids into (results, delay) without waiting; the copy
 * overlaps the next solve.  pvx_fetch_wait returns when the buffers of the last pvx_solve_pipelined are filled. */
PVC_API int  pvx_solve_pipelined(pvx_scene* sc, const float* listenersXYZ, int n, float* results, float* delay);
PVC_API int  pvx_fetch_wait(pvx_scene* sc);
/* frame-loop form of pvx_lookup: the outputs of n_emitters world-space emitter positions (x, y, z triples) for sources
 * 0..n-1 of the last solve, copied into out (n * n_emitters * 8 floats, pvc_host_alloc'd) in stream order, without waiting.
 * An emitter outside the grid yields eight -1 (Analyzer::GetResponseResult returns nullptr there and FDTD.cpp:33-44 reports
 * occlusion -1).  pvx_lookup_wait(ticket) blocks until the copy has landed. */
PVC_API int  pvx_lookup_async(pvx_scene* sc, int n, const float* emittersXYZ, int n_emitters, float* out, int* ticket);
PVC_API int  pvx_lookup_wait(pvx_scene* sc, int ticket);

/* Analyzer::GetResponseResult for a world-space emitter position: 0 = ok and out8 filled,
 * PVC_ERR_INVALID when the reference would return nullptr (outside the grid) */
PVC_API int  pvx_lookup(pvx_scene* sc, int source, float x, float y, float z, float* out8);
/* Planeverb::GetImpulseResponse (FDTD.cpp:60-70): T x {p, vx, vy} of the cell containing the position */
PVC_API int  pvx_impulse_response(pvx_scene* sc, int source, float x, float y, float z, float* out3T);

PVC_API pvc_solver* pvx_solver(pvx_scene* sc);

/* ---- multi-GPU form: listener positions are independent simulations, so a list of them is sharded contiguously
 * over the devices -- one host thread and one stream per device, no data-path collective -- and the per-emitter outputs of every
 * source are gathered into ONE host table.  The reference has no counterpart (one Context, one listener, PvContext.cpp:63-94);
 * this is BASELINE.json's "independent source positions shard embarrassingly across the 8 GPUs" behind the same C ABI. ---- */
typedef struct pvx_multi pvx_multi;
/* devices[n_devices]: CUDA ordinals (an ordinal may repeat: several scenes on one GPU).  maxSources: the longest listener list a
 * solve may carry; each device gets the contiguous shard pvx_shard_bounds(maxSources, n_devices, k) and solves it maxBatch sources
 * at a time (0 = as many as fit 90 % of the device's free memory).  Other arguments as pvx_create. */
PVC_API int  pvx_create_multi(const int* devices, int n_devices, float sizeX, float sizeY, int resolution, int responseLength,
                              float efree, int maxSources, int maxBatch, int maxEmitters, pvx_multi** out);
PVC_API void pvx_destroy_multi(pvx_multi* m);
PVC_API int  pvx_multi_devices(pvx_multi* m);
PVC_API pvx_scene* pvx_multi_scene(pvx_multi* m, int k);
PVC_API int  pvx_multi_batch(pvx_multi* m, int k);
/* geometry edits go to every device's queue (each flushes before its next solve) */
PVC_API int  pvx_multi_add_aabb(pvx_multi* m, float posX, float posY, float width, float height, float absorption);
PVC_API int  pvx_multi_remove_aabb(pvx_multi* m, float posX, float posY, float width, float height, float absorption);
/* GenerateResponse + AnalyzeResponses for n listeners (xyz triples) on all devices at once, then Analyzer::GetResponseResult for
 * n_emitters world positions per listener: out[(listener * n_emitters + emitter) * 8 + field], eight -1 for an emitter outside
 * the grid.  Synchronous; the first failing device's status is returned (pvx_multi_last_error names it). */
PVC_API int  pvx_multi_solve(pvx_multi* m, const float* listenersXYZ, int n, const float* emittersXYZ, int n_emitters, float* out);
PVC_API const char* pvx_multi_last_error(pvx_multi* m);
/* the sharding rule, host arithmetic: contiguous balanced shard [lo, hi) of n_items for part `part` of `parts` */
PVC_API int  pvx_shard_bounds(int n_items, int parts, int part, int* lo, int* hi);

/* ---- host-only helpers (no device needed): the index/scalar derivations the scene solver feeds the
 * CUDA layer, exported so they can be checked against the reference on a machine without a GPU ---- */
/* cfg is filled for (resolution, size, responseLength override); floats[2] = dt, FreeGrid probe radius;
 * ints[5] = FreeGrid listener r,c, emitter r,c, sample count (FreeGrid.cpp:78-99) */
PVC_API int  pvx_derive(int resolution, float sizeX, float sizeY, int responseLength, pvc_config* cfg, float* floats, int* ints);
PVC_API int  pvx_derive_pulse(int resolution, int fs, float* out, int n);
PVC_API int  pvx_derive_rect(int resolution, float posX, float posY, float width, float height, float absorption, int add, pvc_rect* out);
PVC_API int  pvx_derive_listener(int resolution, float x, float z, pvc_listener* out);
/* emitter cell of Analyzer::GetResponseResult; returns PVC_ERR_INVALID where the reference returns nullptr */
PVC_API int  pvx_derive_emitter_cell(int resolution, float sizeX, float sizeY, float x, float z, int* rc);

#ifdef __cplusplus
}
#endif
#endif

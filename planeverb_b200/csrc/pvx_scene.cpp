// pvx_scene.cpp -- host C++ above the C-ABI CUDA layer: one "scene" = the reference's Grid + FreeGrid +
// Analyzer trio (ProjectPlaneverb/src/FDTD/Grid.cpp, FreeGrid.cpp, src/DSP/Analyzer.cpp) with all
// index/scalar derivation done here on the host (pv_params.h) and all field work done on the device.
// Entry points are declared in include/planeverb_ext.h.  There is no CPU solve path in this file: if
// the device layer fails, the error is returned to the caller.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>
#include "pv_params.h"
#include "../../include/planeverb_ext.h"

struct pvx_scene
{
    pvhost::GridParams params;
    pvc_solver* solver = nullptr;
    std::vector<float> pulse;
    std::vector<pvc_rect> pending;     // queued geometry edits, in call order
    std::mutex pendingMutex;           // AddObject/UpdateObject race with the solve thread (GeometryManager.h:46)
    float efree = 0.f;
    int maxSources = 1;
    int lastSources = 0;
    int historySteps = 0;              // > 0: streamed solver with a history of this many samples
};

namespace
{
    // historySteps < 0: the full history if it fits 90 % of the device's free memory, else the longest history (a multiple of 8
    // samples) that does -- a streamed solver, same results at up to twice the time steps (pvc_create_streamed)
    int autoHistory(const pvc_config& cfg, int freeSamples)
    {
        size_t freeB = 0, totalB = 0;
        if (pvc_device_memory(cfg.device, &freeB, &totalB) != PVC_OK) return 0;
        const size_t budget = (size_t)(0.9 * (double)freeB);
        const size_t full = pvc_memory_requirement(&cfg);
        if (full && full <= budget) return 0;
        const int floor8 = std::max(64, (freeSamples + 7) / 8 * 8);
        int lo = floor8, hi = cfg.T / 8 * 8;                    // the requirement grows with the history length: bisect
        if (hi <= lo) return lo;
        if (pvc_memory_requirement_streamed(&cfg, lo) > budget) return lo;      // will fail with PVC_ERR_MEMORY, as it should
        while (hi - lo > 8)
        {
            const int mid = (lo + (hi - lo) / 2) / 8 * 8;
            if (pvc_memory_requirement_streamed(&cfg, mid) <= budget) lo = mid; else hi = mid;
        }
        return lo;
    }
}

extern "C" {

int pvx_create(float sizeX, float sizeY, int resolution, int responseLength, float efree,
               int maxSources, int device, int stepKernel, int variant, pvx_scene** out)
{
    return pvx_create_streamed(sizeX, sizeY, resolution, responseLength, efree, maxSources, device, stepKernel, variant, 0, out);
}

int pvx_create_streamed(float sizeX, float sizeY, int resolution, int responseLength, float efree,
                        int maxSources, int device, int stepKernel, int variant, int historySteps, pvx_scene** out)
{
    if (!out) return PVC_ERR_INVALID;
    *out = nullptr;
    // Context's validation (PvContext.cpp:101-107)
    if (resolution < 275 || sizeX == 0.f || sizeY == 0.f || maxSources < 1) return PVC_ERR_INVALID;
    pvx_scene* sc = new pvx_scene();
    sc->params = pvhost::derive(resolution, sizeX, sizeY, responseLength);
    sc->maxSources = maxSources;
    const pvhost::GridParams& g = sc->params;
    if (g.gx < 2 || g.gy < 2) { delete sc; return PVC_ERR_INVALID; }
    pvc_config cfg = pvhost::configFor(g, maxSources, device, stepKernel);
    cfg.reserved = variant;
    if (historySteps < 0) historySteps = autoHistory(cfg, g.freeSamples);
    sc->historySteps = historySteps;
    int rc = historySteps > 0 ? pvc_create_streamed(&cfg, historySteps, &sc->solver) : pvc_create(&cfg, &sc->solver);
    if (rc) { delete sc; return rc; }
    pvhost::gaussianPulse(resolution, g.fs, sc->pulse, g.T);
    rc = pvc_set_pulse(sc->solver, sc->pulse.data(), g.T);
    if (!rc)
    {
        if (efree >= 0.f) { sc->efree = efree; rc = pvc_set_efree(sc->solver, efree); }
        else
        {
            if (g.freeSamples >= g.T || g.freeEmitterR > g.gx) rc = PVC_ERR_INVALID;   // FreeGrid.cpp:100 asserts this
            else rc = pvc_compute_efree(sc->solver, g.freeListenerR, g.freeListenerC, g.freeEmitterR, g.freeEmitterC,
                                        g.freeSamples, g.freeRadius, &sc->efree);
        }
    }
    if (rc) { pvc_destroy(sc->solver); delete sc; return rc; }
    *out = sc;
    return PVC_OK;
}

int pvx_history_steps(pvx_scene* sc) { return sc ? sc->historySteps : -1; }

void pvx_destroy(pvx_scene* sc)
{
    if (!sc) return;
    pvc_destroy(sc->solver);
    delete sc;
}

int pvx_info(pvx_scene* sc, int* ints, float* floats)
{
    if (!sc) return PVC_ERR_INVALID;
    const pvhost::GridParams& g = sc->params;
    if (ints)
    {
        ints[0] = g.gx; ints[1] = g.gy; ints[2] = g.T; ints[3] = (int)g.fs;
        ints[4] = g.fluxSamples; ints[5] = g.drySamples; ints[6] = g.wetSamples; ints[7] = g.tailSamples;
        ints[8] = g.freeSamples; ints[9] = sc->maxSources;
    }
    if (floats) { floats[0] = g.dx; floats[1] = g.dt; floats[2] = g.courant; floats[3] = sc->efree; }
    return PVC_OK;
}

int pvx_pulse(pvx_scene* sc, float* out, int n)
{
    if (!sc || !out || n < 0) return PVC_ERR_INVALID;
    if (n > (int)sc->pulse.size()) n = (int)sc->pulse.size();
    std::memcpy(out, sc->pulse.data(), sizeof(float) * (size_t)n);
    return PVC_OK;
}

int pvx_add_aabb(pvx_scene* sc, float posX, float posY, float width, float height, float absorption)
{
    if (!sc) return PVC_ERR_INVALID;
    std::lock_guard<std::mutex> lock(sc->pendingMutex);
    sc->pending.push_back(pvhost::rectFor(sc->params, posX, posY, width, height, absorption, true));
    return PVC_OK;
}

int pvx_remove_aabb(pvx_scene* sc, float posX, float posY, float width, float height, float absorption)
{
    if (!sc) return PVC_ERR_INVALID;
    std::lock_guard<std::mutex> lock(sc->pendingMutex);
    sc->pending.push_back(pvhost::rectFor(sc->params, posX, posY, width, height, absorption, false));
    return PVC_OK;
}

int pvx_flush_geometry(pvx_scene* sc)
{
    if (!sc) return PVC_ERR_INVALID;
    std::vector<pvc_rect> batch;
    {
        std::lock_guard<std::mutex> lock(sc->pendingMutex);
        batch.swap(sc->pending);
    }
    if (batch.empty()) return PVC_OK;
    return pvc_apply_geometry(sc->solver, batch.data(), (int)batch.size());
}

int pvx_solve_async(pvx_scene* sc, const float* listenersXYZ, int n, int analyze)
{
    if (!sc || !listenersXYZ || n < 1 || n > sc->maxSources) return PVC_ERR_INVALID;
    int rc = pvx_flush_geometry(sc);
    if (rc) return rc;
    std::vector<pvc_listener> ls((size_t)n);
    for (int i = 0; i < n; ++i)
        ls[(size_t)i] = pvhost::listenerFor(sc->params, listenersXYZ[3 * i], listenersXYZ[3 * i + 2]);
    sc->lastSources = n;
    return pvc_run(sc->solver, ls.data(), n, analyze);
}

int pvx_wait(pvx_scene* sc)
{
    if (!sc) return PVC_ERR_INVALID;
    return pvc_synchronize(sc->solver);
}

int pvx_solve(pvx_scene* sc, const float* listenersXYZ, int n, int analyze, float* results, float* delay)
{
    int rc = pvx_solve_async(sc, listenersXYZ, n, analyze);
    if (rc) return rc;
    const size_t cells = (size_t)sc->params.gx * sc->params.gy;
    if (results || delay)
    {
        for (int i = 0; i < n && !rc; ++i)
            rc = pvc_fetch_results(sc->solver, i, results ? results + (size_t)i * cells * 8 : nullptr,
                                   delay ? delay + (size_t)i * cells : nullptr);
        return rc;
    }
    return pvc_synchronize(sc->solver);
}

int pvx_solve_pipelined(pvx_scene* sc, const float* listenersXYZ, int n, float* results, float* delay)
{
    int rc = pvx_solve_async(sc, listenersXYZ, n, 1);
    if (rc) return rc;
    return pvc_fetch_results_async(sc->solver, n, results, delay);
}

int pvx_fetch_wait(pvx_scene* sc)
{
    if (!sc) return PVC_ERR_INVALID;
    return pvc_fetch_wait(sc->solver);
}

int pvx_lookup(pvx_scene* sc, int source, float x, float y, float z, float* out8)
{
    (void)y;
    if (!sc || !out8) return PVC_ERR_INVALID;
    int r, c;
    if (!pvhost::emitterCell(sc->params, x, z, r, c)) return PVC_ERR_INVALID;
    return pvc_fetch_result_at(sc->solver, source, r, c, out8);
}

int pvx_lookup_async(pvx_scene* sc, int n, const float* emittersXYZ, int n_emitters, float* out, int* ticket)
{
    if (!sc || !emittersXYZ || !out || !ticket || n < 1 || n > sc->maxSources || n_emitters < 0) return PVC_ERR_INVALID;
    std::vector<int> cells((size_t)n_emitters, -1);
    for (int e = 0; e < n_emitters; ++e)
    {
        int r, c;
        if (pvhost::emitterCell(sc->params, emittersXYZ[3 * e], emittersXYZ[3 * e + 2], r, c)) cells[(size_t)e] = r * sc->params.gy + c;
        else
            for (int src = 0; src < n; ++src)
                for (int k = 0; k < 8; ++k) out[((size_t)src * n_emitters + e) * 8 + k] = -1.f;
    }
    return pvc_gather_results_async(sc->solver, n, cells.data(), n_emitters, out, ticket);
}

int pvx_lookup_wait(pvx_scene* sc, int ticket)
{
    if (!sc) return PVC_ERR_INVALID;
    return pvc_gather_wait(sc->solver, ticket);
}

int pvx_impulse_response(pvx_scene* sc, int source, float x, float y, float z, float* out3T)
{
    (void)y;
    if (!sc || !out3T) return PVC_ERR_INVALID;
    const float fx = x / sc->params.dx, fz = z / sc->params.dx;     // FDTD.cpp:64-68
    const int r = (int)fx, c = (int)fz;
    if (r < 0 || c < 0 || r > sc->params.gx || c > sc->params.gy) return PVC_ERR_INVALID;
    return pvc_fetch_ir(sc->solver, source, r, c, out3T);
}

pvc_solver* pvx_solver(pvx_scene* sc) { return sc ? sc->solver : nullptr; }

int pvx_derive(int resolution, float sizeX, float sizeY, int responseLength, pvc_config* cfg, float* floats, int* ints)
{
    if (resolution <= 0 || !cfg) return PVC_ERR_INVALID;
    const pvhost::GridParams g = pvhost::derive(resolution, sizeX, sizeY, responseLength);
    *cfg = pvhost::configFor(g, 1, 0);
    if (floats) { floats[0] = g.dt; floats[1] = g.freeRadius; }
    if (ints) { ints[0] = g.freeListenerR; ints[1] = g.freeListenerC; ints[2] = g.freeEmitterR; ints[3] = g.freeEmitterC; ints[4] = g.freeSamples; }
    return PVC_OK;
}

int pvx_derive_pulse(int resolution, int fs, float* out, int n)
{
    if (resolution <= 0 || fs <= 0 || !out || n < 0) return PVC_ERR_INVALID;
    std::vector<float> v;
    pvhost::gaussianPulse(resolution, (unsigned)fs, v, n);
    std::memcpy(out, v.data(), sizeof(float) * (size_t)n);
    return PVC_OK;
}

int pvx_derive_rect(int resolution, float posX, float posY, float width, float height, float absorption, int add, pvc_rect* out)
{
    if (resolution <= 0 || !out) return PVC_ERR_INVALID;
    const pvhost::GridParams g = pvhost::derive(resolution, 1.f, 1.f);
    *out = pvhost::rectFor(g, posX, posY, width, height, absorption, add != 0);
    return PVC_OK;
}

int pvx_derive_listener(int resolution, float x, float z, pvc_listener* out)
{
    if (resolution <= 0 || !out) return PVC_ERR_INVALID;
    const pvhost::GridParams g = pvhost::derive(resolution, 1.f, 1.f);
    *out = pvhost::listenerFor(g, x, z);
    return PVC_OK;
}

int pvx_derive_emitter_cell(int resolution, float sizeX, float sizeY, float x, float z, int* rc)
{
    if (resolution <= 0 || !rc) return PVC_ERR_INVALID;
    const pvhost::GridParams g = pvhost::derive(resolution, sizeX, sizeY);
    return pvhost::emitterCell(g, x, z, rc[0], rc[1]) ? PVC_OK : PVC_ERR_INVALID;
}

} // extern "C"

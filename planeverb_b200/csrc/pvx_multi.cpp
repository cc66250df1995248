// pvx_multi.cpp -- multi-GPU form of the scene solver, in host C++ above the C-ABI CUDA layer:
// the independent units of the hot path are listener positions ("sources"), so a list of n listeners is sharded contiguously
// over the devices, ONE HOST THREAD AND ONE STREAM PER DEVICE, no data-path collective; every device thread solves its shard
// (in batches that fit its memory), looks its emitter outputs up and writes them into the caller's one host table -- the
// "single gather of scalar acoustic parameters at the end" of BASELINE.json's north_star, done with plain D2H copies.
// Entry points are declared in include/planeverb_ext.h.  No CPU solve path: a failing device fails the call.
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "pv_params.h"
#include "../../include/planeverb_ext.h"

struct pvx_multi
{
    std::vector<int> devices;
    std::vector<pvx_scene*> scenes;      // one per entry of devices (the same ordinal may appear twice: two scenes on one GPU)
    std::vector<int> batch;              // sources per solve call on that scene
    std::vector<float*> staging;         // pinned, batch * maxEmitters * 8 floats per device
    int maxEmitters = 0;
    int maxSources = 0;
    std::string error;
};

namespace
{
    // contiguous, balanced shard [lo, hi) of n items for part k of parts (the first parts take the remainder)
    void shardBounds(int n, int parts, int k, int& lo, int& hi)
    {
        const int base = n / parts, extra = n % parts;
        lo = k * base + std::min(k, extra);
        hi = lo + base + (k < extra ? 1 : 0);
    }
}

extern "C" {

int pvx_shard_bounds(int n_items, int parts, int part, int* lo, int* hi)
{
    if (n_items < 0 || parts < 1 || part < 0 || part >= parts || !lo || !hi) return PVC_ERR_INVALID;
    shardBounds(n_items, parts, part, *lo, *hi);
    return PVC_OK;
}

int pvx_create_multi(const int* devices, int n_devices, float sizeX, float sizeY, int resolution, int responseLength, float efree,
                     int maxSources, int maxBatch, int maxEmitters, pvx_multi** out)
{
    if (!out) return PVC_ERR_INVALID;
    *out = nullptr;
    if (!devices || n_devices < 1 || maxSources < 1 || maxEmitters < 1 || maxBatch < 0) return PVC_ERR_INVALID;
    std::unique_ptr<pvx_multi> m(new pvx_multi());
    m->devices.assign(devices, devices + n_devices);
    m->maxEmitters = maxEmitters;
    m->maxSources = maxSources;
    const pvhost::GridParams g = pvhost::derive(resolution, sizeX, sizeY, responseLength);
    int rc = PVC_OK;
    for (int k = 0; k < n_devices && !rc; ++k)
    {
        int lo, hi;
        shardBounds(maxSources, n_devices, k, lo, hi);
        int batch = std::max(1, hi - lo);
        if (maxBatch > 0) batch = std::min(batch, maxBatch);
        else
        {
            // as many sources per solve as fit 90 % of the device's free memory (the pressure history dominates: 4 B x cells x T each)
            size_t freeB = 0, totalB = 0;
            if (pvc_device_memory(devices[k], &freeB, &totalB) == PVC_OK)
            {
                pvc_config cfg = pvhost::configFor(g, 1, devices[k], 0);
                while (batch > 1)
                {
                    cfg.max_sources = batch;
                    const size_t need = pvc_memory_requirement(&cfg);
                    if (need && need <= (size_t)(0.9 * (double)freeB)) break;
                    --batch;
                }
            }
        }
        pvx_scene* sc = nullptr;
        // history length automatic: a batch of one whose full history does not fit the device still runs, streamed
        rc = pvx_create_streamed(sizeX, sizeY, resolution, responseLength, efree, batch, devices[k], 0, 0, -1, &sc);
        if (rc) break;
        m->scenes.push_back(sc);
        m->batch.push_back(batch);
        float* st = static_cast<float*>(pvc_host_alloc(sizeof(float) * 8 * (size_t)batch * maxEmitters));
        if (!st) { rc = PVC_ERR_MEMORY; break; }
        m->staging.push_back(st);
    }
    if (rc) { pvx_multi* raw = m.release(); pvx_destroy_multi(raw); return rc; }
    *out = m.release();
    return PVC_OK;
}

void pvx_destroy_multi(pvx_multi* m)
{
    if (!m) return;
    for (pvx_scene* sc : m->scenes) pvx_destroy(sc);
    for (float* st : m->staging) pvc_host_free(st);
    delete m;
}

int pvx_multi_devices(pvx_multi* m) { return m ? (int)m->scenes.size() : 0; }
pvx_scene* pvx_multi_scene(pvx_multi* m, int k) { return (m && k >= 0 && k < (int)m->scenes.size()) ? m->scenes[(size_t)k] : nullptr; }
int pvx_multi_batch(pvx_multi* m, int k) { return (m && k >= 0 && k < (int)m->batch.size()) ? m->batch[(size_t)k] : 0; }

int pvx_multi_add_aabb(pvx_multi* m, float posX, float posY, float width, float height, float absorption)
{
    if (!m) return PVC_ERR_INVALID;
    int rc = PVC_OK;
    for (pvx_scene* sc : m->scenes) { const int r = pvx_add_aabb(sc, posX, posY, width, height, absorption); if (r && !rc) rc = r; }
    return rc;
}

int pvx_multi_remove_aabb(pvx_multi* m, float posX, float posY, float width, float height, float absorption)
{
    if (!m) return PVC_ERR_INVALID;
    int rc = PVC_OK;
    for (pvx_scene* sc : m->scenes) { const int r = pvx_remove_aabb(sc, posX, posY, width, height, absorption); if (r && !rc) rc = r; }
    return rc;
}

int pvx_multi_solve(pvx_multi* m, const float* listenersXYZ, int n, const float* emittersXYZ, int n_emitters, float* out)
{
    if (!m || !listenersXYZ || !emittersXYZ || !out || n < 1 || n > m->maxSources || n_emitters < 1 || n_emitters > m->maxEmitters)
        return PVC_ERR_INVALID;
    const int parts = (int)m->scenes.size();
    std::vector<int> status((size_t)parts, PVC_OK);
    std::vector<std::string> text((size_t)parts);
    auto work = [&](int k) {
        int lo, hi;
        shardBounds(n, parts, k, lo, hi);
        pvx_scene* sc = m->scenes[(size_t)k];
        const int batch = m->batch[(size_t)k];
        for (int at = lo; at < hi; at += batch)
        {
            const int cnt = std::min(batch, hi - at);
            // A result slot serves a different listener in every batch and every call, so the reference's "a cell without an onset
            // keeps the previous frame's values" (Analyzer.cpp:161-165) would leak one listener's outputs into another's: the
            // slots start from zero like the first frame after Init (PvContext.cpp:132)
            int rc = PVC_OK;
            for (int i = 0; i < cnt && !rc; ++i) rc = pvc_clear_results(pvx_solver(sc), i);
            if (!rc) rc = pvx_solve_async(sc, listenersXYZ + 3 * (size_t)at, cnt, 1);
            int ticket = 0;
            if (!rc) rc = pvx_lookup_async(sc, cnt, emittersXYZ, n_emitters, m->staging[(size_t)k], &ticket);
            if (!rc) rc = pvx_lookup_wait(sc, ticket);
            if (!rc) rc = pvx_wait(sc);                    // reports a failed run (dependency time-out) of this batch
            if (rc) { status[(size_t)k] = rc; text[(size_t)k] = pvc_last_error(); return; }
            std::memcpy(out + (size_t)at * n_emitters * 8, m->staging[(size_t)k], sizeof(float) * 8 * (size_t)cnt * n_emitters);
        }
    };
    std::vector<std::thread> threads;
    for (int k = 1; k < parts; ++k) threads.emplace_back(work, k);
    work(0);                                               // the calling thread drives the first device
    for (std::thread& t : threads) t.join();
    for (int k = 0; k < parts; ++k)
        if (status[(size_t)k]) { m->error = "device " + std::to_string(m->devices[(size_t)k]) + ": " + text[(size_t)k]; return status[(size_t)k]; }
    return PVC_OK;
}

const char* pvx_multi_last_error(pvx_multi* m) { return m ? m->error.c_str() : ""; }

} // extern "C"

// planeverb_api.cpp -- the Planeverb C++ API (include/Planeverb.h) and the Unity C ABI
// (include/PlaneverbUnity.h) on top of the CUDA scene solver.
//
// Re-implements, thinly and in host C++, the callers around the hot path so the Unity plugin and the
// Sandbox link unchanged:
//   Context + BackgroundProcessor   ProjectPlaneverb/src/Context/PvContext.cpp:25-57,63-94,97-179
//   GeometryManager                 ProjectPlaneverb/src/Geometry/GeometryManager.cpp:67-152
//   EmissionManager                 ProjectPlaneverb/src/Emissions/EmissionManager.cpp:37-75
//   GetOutput / GetImpulseResponse  ProjectPlaneverb/src/FDTD/FDTD.cpp:16-70
//   C ABI                           ProjectPlaneverb/PlaneverbUnityPluginAPI/PlaneverbUnity.cpp:12-135
// The solve itself (GenerateResponse + AnalyzeResponses) is the device path of pvx_scene.cpp; there is
// no CPU solver here.  Differences from the reference, all deliberate:
//   * results are published through page-locked host grids swapped atomically after every frame (three of them:
//     the loop runs one frame deep, a frame's grid is copied out while the next frame is solved), so GetOutput on
//     the game/audio thread never reads a half-written frame (the reference races by design);
//   * the listener position and running flag are atomics / mutex-protected;
//   * errors thrown inside the C ABI are caught there (the reference lets enums escape extern "C").
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/Planeverb.h"
#include "../../include/PlaneverbUnity.h"
#include "../../include/planeverb_ext.h"
#include "pv_params.h"

namespace Planeverb
{
    namespace
    {
        struct Context
        {
            PlaneverbConfig config;
            pvhost::GridParams params;
            pvx_scene* scene = nullptr;

            // background solve loop
            std::thread worker;
            std::atomic<bool> running{ true };
            std::atomic<unsigned long long> frames{ 0 };
            std::mutex solverMutex;               // serialises device access (worker vs GetImpulseResponse)
            std::atomic<int> solverWaiters{ 0 };  // API threads waiting for the device: the worker yields to them

            // listener (written by the game thread, read by the worker)
            std::mutex listenerMutex;
            vec3 listener;

            // published results: three page-locked host grids of gx*gy*8 floats, index of the readable one.  The worker
            // runs one frame deep: while frame k+1 is solved, frame k's grid is copied into the second buffer
            // (pvx_solve_pipelined) and the game / audio threads read frame k-1 from the third.
            float* grid[3] = { nullptr, nullptr, nullptr };
            std::atomic<int> readable{ 0 };
            std::mutex publishMutex;              // held only for the pointer swap / a 32-byte copy

            // emitters: id -> position with free-list reuse (EmissionManager.cpp:37-54)
            std::mutex emitterMutex;
            std::vector<vec3> emitters;
            std::vector<EmissionID> freeEmitters;

            // geometry: id -> AABB with free-list reuse (GeometryManager.cpp:67-93)
            std::mutex geometryMutex;
            std::vector<AABB> objects;
            std::vector<PlaneObjectID> freeObjects;

            // GetImpulseResponse scratch
            std::vector<Cell> irCells;
            std::vector<float> irFloats;
        };

        std::mutex g_contextMutex;
        std::unique_ptr<Context> g_context;
        thread_local std::string g_lastError;

        void workerLoop(Context* ctx)
        {
            int filling = -1;                     // buffer the copy in flight writes to (-1: none)
            int next = 1;                         // buffer the next frame will be copied into
            while (ctx->running.load(std::memory_order_acquire))
            {
                vec3 l;
                {
                    std::lock_guard<std::mutex> lock(ctx->listenerMutex);
                    l = ctx->listener;
                }
                const float xyz[3] = { l.x, l.y, l.z };
                int rc;
                while (ctx->solverWaiters.load(std::memory_order_acquire) > 0) std::this_thread::yield();
                {
                    std::lock_guard<std::mutex> lock(ctx->solverMutex);
                    // queued geometry edits are flushed inside the solve call before the time steps (the reference
                    // applies them after the previous frame's analysis, PvContext.cpp:86: same ordering).  The call
                    // enqueues this frame and returns once the PREVIOUS frame's grid has landed in `filling`.
                    rc = pvx_solve_pipelined(ctx->scene, xyz, 1, ctx->grid[next], nullptr);
                }
                if (rc != PVC_OK)
                {   // device failure: stop publishing; GetOutput keeps serving the last good frame
                    ctx->running.store(false, std::memory_order_release);
                    break;
                }
                if (filling >= 0)
                {
                    int old;
                    {
                        std::lock_guard<std::mutex> lock(ctx->publishMutex);
                        old = ctx->readable.load(std::memory_order_acquire);
                        ctx->readable.store(filling, std::memory_order_release);
                    }
                    // every grid is a full snapshot of the device's persistent result grid, so the stale-result
                    // semantics of the reference (cells without an onset keep their previous values,
                    // Analyzer.cpp:161-165) carry over whichever buffer is reused
                    filling = next;
                    next = old;
                    ctx->frames.fetch_add(1, std::memory_order_release);
                }
                else
                {
                    filling = next;
                    next = 2;
                }
            }
            if (filling >= 0)
            {   // drain the copy in flight (Exit joins this thread before it frees the buffers)
                std::lock_guard<std::mutex> lock(ctx->solverMutex);
                pvx_fetch_wait(ctx->scene);
            }
        }

        Context* current() { return g_context.get(); }
    } // namespace

    void Init(const PlaneverbConfig* config)
    {
        std::lock_guard<std::mutex> guard(g_contextMutex);
        if (g_context)
        {
            g_context->running.store(false);
            if (g_context->worker.joinable()) g_context->worker.join();
            pvx_destroy(g_context->scene);
            for (float* g : g_context->grid) pvc_host_free(g);
            g_context.reset();
        }
        // PvContext.cpp:101-107
        if (config == nullptr || config->gridResolution < pv_LowResolution ||
            config->gridSizeInMeters.x == 0 || config->gridSizeInMeters.y == 0 ||
            config->tempFileDirectory == nullptr)
        {
            throw pv_InvalidConfig;
        }
        std::unique_ptr<Context> ctx(new Context());
        std::memcpy(&ctx->config, config, sizeof(PlaneverbConfig));
        ctx->params = pvhost::derive(config->gridResolution, config->gridSizeInMeters.x, config->gridSizeInMeters.y);
        int device = 0;
        if (const char* env = std::getenv("PLANEVERB_CUDA_DEVICE")) device = std::atoi(env);
        const int rc = pvx_create(config->gridSizeInMeters.x, config->gridSizeInMeters.y, config->gridResolution,
                                  0, -1.f, 1, device, 0, 0, &ctx->scene);
        if (rc != PVC_OK)
        {
            g_lastError = pvc_last_error();
            throw (rc == PVC_ERR_MEMORY) ? pv_NotEnoughMemory : pv_InvalidConfig;
        }
        const size_t cells = (size_t)ctx->params.gx * ctx->params.gy;
        for (int i = 0; i < 3; ++i)
        {
            ctx->grid[i] = static_cast<float*>(pvc_host_alloc(sizeof(float) * cells * 8));
            if (!ctx->grid[i])
            {
                g_lastError = pvc_last_error();
                for (int k = 0; k < i; ++k) pvc_host_free(ctx->grid[k]);
                pvx_destroy(ctx->scene);
                throw pv_NotEnoughMemory;
            }
            std::memset(ctx->grid[i], 0, sizeof(float) * cells * 8);      // Context's memset (PvContext.cpp:132)
        }
        ctx->worker = std::thread(workerLoop, ctx.get());
        g_context = std::move(ctx);
    }

    void Exit()
    {
        std::lock_guard<std::mutex> guard(g_contextMutex);
        if (!g_context) return;
        g_context->running.store(false, std::memory_order_release);
        if (g_context->worker.joinable()) g_context->worker.join();
        pvx_destroy(g_context->scene);
        for (float* g : g_context->grid) pvc_host_free(g);
        g_context.reset();
    }

    void ChangeSettings(const PlaneverbConfig* newConfig)
    {
        Exit();
        Init(newConfig);
    }

    void SetListenerPosition(const vec3& listenerPosition)
    {
        Context* ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->listenerMutex);
        ctx->listener = listenerPosition;
    }

    EmissionID Emit(const vec3& emitterPosition)
    {
        Context* ctx = current();
        if (!ctx) return PV_INVALID_EMISSION_ID;
        std::lock_guard<std::mutex> lock(ctx->emitterMutex);
        if (!ctx->freeEmitters.empty())
        {
            const EmissionID id = ctx->freeEmitters.back();
            ctx->freeEmitters.pop_back();
            ctx->emitters[id] = emitterPosition;
            return id;
        }
        ctx->emitters.push_back(emitterPosition);
        return ctx->emitters.size() - 1;
    }

    void UpdateEmission(EmissionID id, const vec3& position)
    {
        Context* ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->emitterMutex);
        if (id < ctx->emitters.size()) ctx->emitters[id] = position;
    }

    void EndEmission(EmissionID id)
    {
        Context* ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->emitterMutex);
        if (id < ctx->emitters.size()) ctx->freeEmitters.push_back(id);
    }

    PlaneverbOutput GetOutput(EmissionID emitter)
    {
        PlaneverbOutput out{};
        out.occlusion = PV_INVALID_DRY_GAIN;             // FDTD.cpp:23-47: every failure path
        Context* ctx = current();
        if (!ctx) return out;
        vec3 pos;
        {
            std::lock_guard<std::mutex> lock(ctx->emitterMutex);
            if (emitter >= ctx->emitters.size()) return out;
            pos = ctx->emitters[emitter];
        }
        int r, c;
        if (!pvhost::emitterCell(ctx->params, pos.x, pos.z, r, c,
                                 ctx->config.gridWorldOffset.x, ctx->config.gridWorldOffset.y))
            return out;
        float v[8];
        {
            std::lock_guard<std::mutex> lock(ctx->publishMutex);
            const float* g = ctx->grid[ctx->readable.load(std::memory_order_acquire)];
            std::memcpy(v, g + ((size_t)r * ctx->params.gy + c) * 8, sizeof(v));
        }
        out.occlusion = v[0];
        out.wetGain = v[1];
        out.rt60 = v[2];
        out.lowpass = v[3];
        out.direction = vec2(v[4], v[5]);
        out.sourceDirectivity = vec2(v[6], v[7]);
        return out;
    }

    namespace
    {
        PlaneObjectID addObject(Context* ctx, const AABB& box)
        {
            PlaneObjectID id;
            if (ctx->freeObjects.empty()) { ctx->objects.push_back(box); id = ctx->objects.size() - 1; }
            else { id = ctx->freeObjects.back(); ctx->freeObjects.pop_back(); ctx->objects[id] = box; }
            return id;
        }
    }

    PlaneObjectID AddGeometry(const AABB* transform)
    {
        Context* ctx = current();
        if (!ctx || !transform) return PV_INVALID_PLANE_OBJECT_ID;
        std::lock_guard<std::mutex> lock(ctx->geometryMutex);
        const PlaneObjectID id = addObject(ctx, *transform);
        pvx_add_aabb(ctx->scene, transform->position.x, transform->position.y, transform->width, transform->height, transform->absorption);
        return id;
    }

    void UpdateGeometry(PlaneObjectID id, const AABB* newTransform)
    {
        Context* ctx = current();
        if (!ctx || !newTransform) return;
        std::lock_guard<std::mutex> lock(ctx->geometryMutex);
        if (id >= ctx->objects.size()) return;
        const AABB old = ctx->objects[id];
        // remove(old) then add(new), in that order, in one queue (GeometryManager.cpp:112-121)
        pvx_remove_aabb(ctx->scene, old.position.x, old.position.y, old.width, old.height, old.absorption);
        ctx->objects[id] = *newTransform;
        pvx_add_aabb(ctx->scene, newTransform->position.x, newTransform->position.y, newTransform->width, newTransform->height, newTransform->absorption);
    }

    void RemoveGeometry(PlaneObjectID id)
    {
        Context* ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->geometryMutex);
        if (id >= ctx->objects.size()) return;
        const AABB old = ctx->objects[id];
        pvx_remove_aabb(ctx->scene, old.position.x, old.position.y, old.width, old.height, old.absorption);
        ctx->objects[id] = AABB();
        ctx->freeObjects.push_back(id);
    }

    std::pair<const Cell*, unsigned> GetImpulseResponse(const vec3& position)
    {
        Context* ctx = current();
        if (!ctx) return std::make_pair((const Cell*)nullptr, 0u);
        const unsigned T = (unsigned)ctx->params.T;
        ctx->solverWaiters.fetch_add(1, std::memory_order_acq_rel);
        std::lock_guard<std::mutex> lock(ctx->solverMutex);
        ctx->solverWaiters.fetch_sub(1, std::memory_order_acq_rel);
        ctx->irFloats.assign((size_t)T * 3, 0.f);
        ctx->irCells.assign(T, Cell());
        if (ctx->frames.load() > 0 &&
            pvx_impulse_response(ctx->scene, 0, position.x, position.y, position.z, ctx->irFloats.data()) == PVC_OK)
        {
            for (unsigned t = 0; t < T; ++t)
            {
                ctx->irCells[t].pr = ctx->irFloats[3 * t];
                ctx->irCells[t].vx = ctx->irFloats[3 * t + 1];
                ctx->irCells[t].vy = ctx->irFloats[3 * t + 2];
            }
        }
        return std::make_pair((const Cell*)ctx->irCells.data(), T);
    }
} // namespace Planeverb

// ------------------------------------------------------------------------------------------------
// Unity C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

void PVU_CC UnityPluginLoad(void* unityInterfaces) { (void)unityInterfaces; }
void PVU_CC UnityPluginUnload(void) {}

void PVU_CC PlaneverbInit(float gridSizeX, float gridSizeY, int gridResolution, int gridBoundaryType,
                          char* tempFileDir, int maxThreadUsage, int threadExecutionType)
{
    Planeverb::PlaneverbConfig config;
    config.gridSizeInMeters.x = gridSizeX;
    config.gridSizeInMeters.y = gridSizeY;
    config.gridResolution = gridResolution;
    config.gridBoundaryType = (Planeverb::PlaneverbBoundaryType)gridBoundaryType;
    config.tempFileDirectory = tempFileDir;
    config.maxThreadUsage = (unsigned)maxThreadUsage;
    config.threadExecutionType = (Planeverb::PlaneverbExecutionType)threadExecutionType;
    try { Planeverb::Init(&config); }
    catch (Planeverb::PlaneverbErrorCode code)
    {
        Planeverb::g_lastError = (code == Planeverb::pv_NotEnoughMemory ? "pv_NotEnoughMemory: " : "pv_InvalidConfig: ") + std::string(pvc_last_error());
    }
}

void PVU_CC PlaneverbExit(void) { Planeverb::Exit(); }

int PVU_CC PlaneverbEmit(float x, float y, float z) { return (int)Planeverb::Emit(Planeverb::vec3(x, y, z)); }

void PVU_CC PlaneverbUpdateEmission(int id, float x, float y, float z)
{
    Planeverb::UpdateEmission((Planeverb::EmissionID)id, Planeverb::vec3(x, y, z));
}

void PVU_CC PlaneverbEndEmission(int id) { Planeverb::EndEmission((Planeverb::EmissionID)id); }

PlaneverbUnityOutput PVU_CC PlaneverbGetOutput(int emissionID)
{
    const Planeverb::PlaneverbOutput o = Planeverb::GetOutput((Planeverb::EmissionID)emissionID);
    PlaneverbUnityOutput out;
    out.occlusion = o.occlusion;
    out.wetGain = o.wetGain;
    out.rt60 = o.rt60;
    out.lowpass = o.lowpass;
    out.directionX = o.direction.x;
    out.directionY = o.direction.y;
    out.sourceDirectionX = o.sourceDirectivity.x;
    out.sourceDirectionY = o.sourceDirectivity.y;
    return out;
}

int PVU_CC PlaneverbAddGeometry(float posX, float posY, float width, float height, float absorption)
{
    Planeverb::AABB box;
    box.position = Planeverb::vec2(posX, posY);
    box.width = width; box.height = height; box.absorption = absorption;
    return (int)Planeverb::AddGeometry(&box);
}

void PVU_CC PlaneverbUpdateGeometry(int id, float posX, float posY, float width, float height, float absorption)
{
    Planeverb::AABB box;
    box.position = Planeverb::vec2(posX, posY);
    box.width = width; box.height = height; box.absorption = absorption;
    Planeverb::UpdateGeometry((Planeverb::PlaneObjectID)id, &box);
}

void PVU_CC PlaneverbRemoveGeometry(int id) { Planeverb::RemoveGeometry((Planeverb::PlaneObjectID)id); }

void PVU_CC PlaneverbSetListenerPosition(float x, float y, float z)
{
    Planeverb::SetListenerPosition(Planeverb::vec3(x, y, z));
}

unsigned long long PVU_CC PlaneverbFramesCompleted(void)
{
    Planeverb::Context* ctx = Planeverb::current();
    return ctx ? ctx->frames.load(std::memory_order_acquire) : 0ull;
}

const char* PVU_CC PlaneverbLastError(void) { return Planeverb::g_lastError.c_str(); }

} // extern "C"

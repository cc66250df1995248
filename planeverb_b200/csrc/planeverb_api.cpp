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
//   * every API call works on a shared_ptr snapshot of the context taken under the context mutex, so Exit() on the game
//     thread cannot free the result grids under a GetOutput() running on the audio thread (AudioCore.cpp:95 pattern): the
//     last holder frees them;
//   * a listener outside the grid (or NaN) skips the frame -- GetOutput keeps serving the last good one -- instead of
//     indexing outside the grid as the reference does (FDTD.cpp:97-99); a DEVICE failure stops the worker, and both are
//     reported through the process-wide PlaneverbLastError() / PlaneverbWorkerState();
//   * PlaneverbConfig::gridWorldOffset != 0 is rejected with pv_InvalidConfig (the reference's header marks it
//     "!!! Not supported !!!", PvTypes.h:58, and applies it inconsistently, Grid.cpp:139-142 vs 252-255);
//   * errors thrown inside the C ABI are caught there (the reference lets enums escape extern "C").
#include <atomic>
#include <chrono>
#include <cmath>
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
            std::atomic<unsigned long long> skipped{ 0 };   // frames skipped because the listener was outside the grid
            std::atomic<bool> listenerWarned{ false };

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

            std::atomic<int> workerState{ 0 };    // 0 not started, 1 running, 2 stopped by Exit, -1 stopped by a device failure

            void shutdown()                       // stop and join the worker (idempotent; called by Exit / Init, never by the worker)
            {
                running.store(false, std::memory_order_release);
                if (worker.joinable()) worker.join();
            }
            ~Context()
            {   // runs in whichever thread drops the last reference: the worker has been joined by then (Exit / Init)
                shutdown();
                if (scene) pvx_destroy(scene);
                for (float* g : grid) pvc_host_free(g);
            }
        };

        std::mutex g_contextMutex;
        std::shared_ptr<Context> g_context;

        // process-wide error text: the worker thread's failures must reach whoever asks (PlaneverbLastError)
        std::mutex g_errorMutex;
        std::string g_lastError;
        void setLastError(const std::string& text)
        {
            std::lock_guard<std::mutex> lock(g_errorMutex);
            g_lastError = text;
        }

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
                {
                    // FDTD.cpp:97-99 turns the listener position into a cell without any check; outside the grid (or NaN) the
                    // reference indexes out of bounds.  Here such a frame is skipped and the last good results stay published.
                    bool inside = std::isfinite(l.x) && std::isfinite(l.z);
                    if (inside)
                    {
                        const float fx = l.x / ctx->params.dx, fz = l.z / ctx->params.dx;
                        inside = fx >= 0.f && fz >= 0.f && fx < (float)(ctx->params.gx + 1) && fz < (float)(ctx->params.gy + 1);
                    }
                    if (!inside)
                    {
                        if (!ctx->listenerWarned.exchange(true))
                            setLastError("listener position outside the grid: frames are skipped until it returns (last good results stay published)");
                        ctx->skipped.fetch_add(1, std::memory_order_release);
                        std::this_thread::sleep_for(std::chrono::milliseconds(1));
                        continue;
                    }
                    ctx->listenerWarned.store(false);
                }
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
                {   // device failure: stop publishing (GetOutput keeps serving the last good frame) and say so
                    setLastError(std::string("acoustics worker stopped: ") + pvc_last_error());
                    ctx->workerState.store(-1, std::memory_order_release);
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
            {   // drain the copy in flight (Exit joins this thread before the buffers can be freed)
                std::lock_guard<std::mutex> lock(ctx->solverMutex);
                if (pvx_fetch_wait(ctx->scene) != PVC_OK && ctx->workerState.load() != -1)
                {
                    setLastError(std::string("acoustics worker: last frame failed: ") + pvc_last_error());
                    ctx->workerState.store(-1, std::memory_order_release);
                }
            }
            int expected = 1;
            ctx->workerState.compare_exchange_strong(expected, 2);
        }

        // every API call works on a snapshot: Exit() may reset g_context while another thread is inside GetOutput()
        std::shared_ptr<Context> current()
        {
            std::lock_guard<std::mutex> guard(g_contextMutex);
            return g_context;
        }
    } // namespace

    void Init(const PlaneverbConfig* config)
    {
        Exit();                                        // PvContext.cpp:27-30
        // PvContext.cpp:101-107
        if (config == nullptr || config->gridResolution < pv_LowResolution ||
            config->gridSizeInMeters.x == 0 || config->gridSizeInMeters.y == 0 ||
            config->tempFileDirectory == nullptr)
        {
            setLastError("Init: invalid PlaneverbConfig (PvContext.cpp:101-107)");
            throw pv_InvalidConfig;
        }
        if (config->gridWorldOffset.x != 0.f || config->gridWorldOffset.y != 0.f)
        {
            setLastError("Init: gridWorldOffset is not supported (PvTypes.h:58)");
            throw pv_InvalidConfig;
        }
        std::shared_ptr<Context> ctx = std::make_shared<Context>();
        std::memcpy(&ctx->config, config, sizeof(PlaneverbConfig));
        ctx->params = pvhost::derive(config->gridResolution, config->gridSizeInMeters.x, config->gridSizeInMeters.y);
        int device = 0;
        if (const char* env = std::getenv("PLANEVERB_CUDA_DEVICE")) device = std::atoi(env);
        // history length automatic (-1): a grid whose full pressure history does not fit the device runs on the streamed solver
        // instead of failing with pv_NotEnoughMemory (only GetImpulseResponse is unavailable then).  PLANEVERB_HISTORY_STEPS forces
        // a history length (operators of memory-shared devices; the plugin test of the streamed path).
        int historySteps = -1;
        if (const char* env = std::getenv("PLANEVERB_HISTORY_STEPS")) historySteps = std::atoi(env);
        const int rc = pvx_create_streamed(config->gridSizeInMeters.x, config->gridSizeInMeters.y, config->gridResolution,
                                           0, -1.f, 1, device, 0, 0, historySteps, &ctx->scene);
        if (rc != PVC_OK)
        {
            setLastError(std::string("Init: ") + pvc_last_error());
            throw (rc == PVC_ERR_MEMORY) ? pv_NotEnoughMemory : pv_InvalidConfig;
        }
        const size_t cells = (size_t)ctx->params.gx * ctx->params.gy;
        for (int i = 0; i < 3; ++i)
        {
            ctx->grid[i] = static_cast<float*>(pvc_host_alloc(sizeof(float) * cells * 8));
            if (!ctx->grid[i])
            {
                setLastError(std::string("Init: ") + pvc_last_error());
                throw pv_NotEnoughMemory;             // ~Context frees what was allocated
            }
            std::memset(ctx->grid[i], 0, sizeof(float) * cells * 8);      // Context's memset (PvContext.cpp:132)
        }
        setLastError("");
        ctx->workerState.store(1, std::memory_order_release);
        ctx->worker = std::thread(workerLoop, ctx.get());
        std::lock_guard<std::mutex> guard(g_contextMutex);
        g_context = std::move(ctx);
    }

    void Exit()
    {
        std::shared_ptr<Context> ctx;
        {
            std::lock_guard<std::mutex> guard(g_contextMutex);
            ctx.swap(g_context);
        }
        if (!ctx) return;
        ctx->shutdown();        // join the worker here; the buffers go when the last snapshot (a GetOutput in flight) is dropped
    }

    void ChangeSettings(const PlaneverbConfig* newConfig)
    {
        Exit();
        Init(newConfig);
    }

    void SetListenerPosition(const vec3& listenerPosition)
    {
        const std::shared_ptr<Context> ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->listenerMutex);
        ctx->listener = listenerPosition;
    }

    EmissionID Emit(const vec3& emitterPosition)
    {
        const std::shared_ptr<Context> ctx = current();
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
        const std::shared_ptr<Context> ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->emitterMutex);
        if (id < ctx->emitters.size()) ctx->emitters[id] = position;
    }

    void EndEmission(EmissionID id)
    {
        const std::shared_ptr<Context> ctx = current();
        if (!ctx) return;
        std::lock_guard<std::mutex> lock(ctx->emitterMutex);
        if (id < ctx->emitters.size()) ctx->freeEmitters.push_back(id);
    }

    PlaneverbOutput GetOutput(EmissionID emitter)
    {
        PlaneverbOutput out{};
        out.occlusion = PV_INVALID_DRY_GAIN;             // FDTD.cpp:23-47: every failure path
        const std::shared_ptr<Context> ctx = current();
        if (!ctx) return out;
        vec3 pos;
        {
            std::lock_guard<std::mutex> lock(ctx->emitterMutex);
            if (emitter >= ctx->emitters.size()) return out;
            pos = ctx->emitters[emitter];
        }
        int r, c;
        if (!pvhost::emitterCell(ctx->params, pos.x, pos.z, r, c)) return out;
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
        const std::shared_ptr<Context> ctx = current();
        if (!ctx || !transform) return PV_INVALID_PLANE_OBJECT_ID;
        std::lock_guard<std::mutex> lock(ctx->geometryMutex);
        const PlaneObjectID id = addObject(ctx.get(), *transform);
        pvx_add_aabb(ctx->scene, transform->position.x, transform->position.y, transform->width, transform->height, transform->absorption);
        return id;
    }

    void UpdateGeometry(PlaneObjectID id, const AABB* newTransform)
    {
        const std::shared_ptr<Context> ctx = current();
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
        const std::shared_ptr<Context> ctx = current();
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
        const std::shared_ptr<Context> ctx = current();
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
    catch (Planeverb::PlaneverbErrorCode)
    {
        // Init has recorded the reason (PlaneverbLastError); the reference lets the enum escape extern "C" (PlaneverbUnity.cpp:39)
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
    const std::shared_ptr<Planeverb::Context> ctx = Planeverb::current();
    return ctx ? ctx->frames.load(std::memory_order_acquire) : 0ull;
}

int PVU_CC PlaneverbHistorySteps(void)
{
    const std::shared_ptr<Planeverb::Context> ctx = Planeverb::current();
    return (ctx && ctx->scene) ? pvx_history_steps(ctx->scene) : -1;
}

int PVU_CC PlaneverbWorkerState(void)
{
    const std::shared_ptr<Planeverb::Context> ctx = Planeverb::current();
    return ctx ? ctx->workerState.load(std::memory_order_acquire) : 0;
}

const char* PVU_CC PlaneverbLastError(void)
{
    thread_local std::string copy;                 // the caller's own copy: the worker may overwrite the shared text at any time
    std::lock_guard<std::mutex> lock(Planeverb::g_errorMutex);
    copy = Planeverb::g_lastError;
    return copy.c_str();
}

} // extern "C"

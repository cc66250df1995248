// Planeverb.h -- the client C++ API, source-compatible with ProjectPlaneverb/include/Planeverb.h:8-49
// (same namespace, names, signatures and error behaviour), implemented on the B200 CUDA path in
// planeverb_b200/csrc/planeverb_api.cpp.
#pragma once
#include <utility>
#include "PvDefinitions.h"
#include "PvTypes.h"

namespace Planeverb
{
    // Start the acoustics module: allocates device memory, runs the free-field normalisation and starts
    // the background solve loop.  Throws pv_InvalidConfig / pv_NotEnoughMemory (bare enum values).
    PV_API void Init(const PlaneverbConfig* config);
    PV_API void Exit();
    // Exit followed by Init with the new config
    PV_API void ChangeSettings(const PlaneverbConfig* newConfig);

    // emitters: id -> position table; GetOutput looks the emitter's cell up in the latest analysed frame
    PV_API EmissionID Emit(const vec3& emitterPosition);
    PV_API void UpdateEmission(EmissionID id, const vec3& position);
    PV_API void EndEmission(EmissionID id);
    PV_API PlaneverbOutput GetOutput(EmissionID emitter);

    // geometry edits are queued and reach the grid between two solves, in call order
    PV_API PlaneObjectID AddGeometry(const AABB* transform);
    PV_API void UpdateGeometry(PlaneObjectID id, const AABB* newTransform);
    PV_API void RemoveGeometry(PlaneObjectID id);

    PV_API void SetListenerPosition(const vec3& listenerPosition);

    // debugging aid: impulse response (T samples) of the cell containing `position`; the buffer is owned
    // by the library and valid until the next call or Exit
    PV_API std::pair<const Cell*, unsigned> GetImpulseResponse(const vec3& position);
} // namespace Planeverb

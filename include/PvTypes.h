// PvTypes.h -- source-compatible with ProjectPlaneverb/include/PvTypes.h: same enums, struct layouts,
// id types and constants (they are the public contract Unity's C# mirror and the Sandbox compile
// against), written for this library.
#pragma once
#include <cstddef>
#include "PvMathTypes.h"

namespace Planeverb
{
    // thrown (as the bare enum) by Init/ChangeSettings, as in the reference (Planeverb.h:11)
    enum PlaneverbErrorCode
    {
        pv_NotEnoughMemory,   // host or device allocation failed
        pv_InvalidConfig,     // null/invalid config, or no usable CUDA device
    };

    // The reference reserves value 1 for a GPU path it never shipped (PvTypes.h:13-17; the C# enum
    // already exposes it, PlaneverbConfig.cs:23-29).  This library implements that path; both values
    // run on the GPU because there is no CPU solver here.
    enum PlaneverbExecutionType
    {
        pv_CPU,
        pv_GPU,
    };

    // maximum frequency resolved by the grid; cell size = (c / resolution) / 3.5
    enum PlaneverbResolution
    {
        pv_LowResolution = 275,
        pv_MidResolution = 375,
        pv_HighResolution = 500,
        pv_ExtremeResolution = 750,
        pv_DefaultResolution = pv_MidResolution
    };

    enum PlaneverbBoundaryType
    {
        pv_AbsorbingBoundary,    // grid edges absorb (the only behaviour implemented, as in the reference)
        pv_ReflectingBoundary,
    };

    struct PlaneverbConfig
    {
        vec2 gridSizeInMeters = { 10.f, 10.f };
        int gridResolution = pv_DefaultResolution;
        PlaneverbBoundaryType gridBoundaryType = pv_AbsorbingBoundary;
        const char* tempFileDirectory;                         // must be non-null (validated, never used)
        unsigned maxThreadUsage = 0;                           // accepted for compatibility; unused on the GPU
        PlaneverbExecutionType threadExecutionType = pv_CPU;
        vec2 gridWorldOffset = { 0.f, 0.f };                   // unsupported, as in the reference
    };

    // acoustic parameters of one emitter (32 bytes)
    struct PlaneverbOutput
    {
        float occlusion;
        float wetGain;
        float rt60;
        float lowpass;
        vec2 direction;
        vec2 sourceDirectivity;
    };

    using EmissionID = size_t;
    using PlaneObjectID = size_t;

    const constexpr PlaneObjectID PV_INVALID_PLANE_OBJECT_ID = (PlaneObjectID)(-1);
    const constexpr EmissionID PV_INVALID_EMISSION_ID = (EmissionID)(-1);
    const constexpr Real PV_INVALID_DRY_GAIN = (Real)-1.f;

    // physical / analysis constants (values are part of the contract; PvTypes.h:83-101)
    const constexpr Real PV_PI = (Real)3.141593f;
    const constexpr Real PV_RHO = (Real)1.2041f;
    const constexpr Real PV_C = (Real)343.21f;
    const constexpr Real PV_Z_AIR = PV_RHO * PV_C;
    const constexpr Real PV_INV_Z_AIR = (Real)1.f / PV_Z_AIR;
    const constexpr Real PV_INV_Z_REFLECT = (Real)0.0f;
    const constexpr Real PV_AUDIBLE_THRESHOLD_GAIN = (Real)0.00000316f;
    const constexpr Real PV_DRY_DIRECTION_ANALYSIS_LENGTH = (Real)0.005f;
    const constexpr Real PV_DRY_GAIN_ANALYSIS_LENGTH = (Real)0.01f;
    const constexpr Real PV_WET_GAIN_ANALYSIS_LENGTH = (Real)0.080f;
    const constexpr Real PV_SQRT_2 = (Real)1.4142136f;
    const constexpr Real PV_SQRT_3 = (Real)1.7320508f;
    const constexpr Real PV_MAX_AUDIBLE_FREQ = (Real)20000.f;
    const constexpr Real PV_MIN_AUDIBLE_FREQ = (Real)20.f;
    const constexpr Real PV_POINTS_PER_WAVELENGTH = (Real)3.5f;
    const constexpr Real PV_SCHROEDER_OFFSET_S = (Real)0.01f;
    const constexpr Real PV_DISTANCE_GAIN_THRESHOLD = (Real)0.891251f;
    const constexpr Real PV_DELAY_CLOSE_THRESHOLD = (Real)5.f;
    const constexpr Real PV_IMPULSE_RESPONSE_S = PV_SQRT_2 * Real(12.5) / PV_C + Real(0.25);

    // one impulse-response sample as GetImpulseResponse hands it out (16 bytes)
    struct Cell
    {
        Real pr;
        Real vx;
        Real vy;
        short b;
        short by;
        Cell(Real pr_ = 0.f, Real vx_ = 0.f, Real vy_ = 0.f, int b_ = 1, int by_ = 1)
            : pr(pr_), vx(vx_), vy(vy_), b((short)b_), by((short)by_) {}
    };
} // namespace Planeverb

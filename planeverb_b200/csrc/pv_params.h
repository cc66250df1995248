// pv_params.h -- host-side derivation of every scalar and index the solver needs, evaluated with the
// SAME single-precision expressions (operand order, intermediate types, truncation) as the reference,
// because a one-ulp difference before a float->int truncation moves a wall or the listener by a cell
// (SURVEY.md App. D.3).  Compile with -ffp-contract=off and without fast-math.
//
// Reference expressions restated here:
//   CalculateGridParameters   ProjectPlaneverb/src/FDTD/Grid.cpp:390-396
//   grid size / IR length     ProjectPlaneverb/src/FDTD/Grid.cpp:48-55, include/PvTypes.h:101
//   Gaussian pulse            ProjectPlaneverb/src/FDTD/Grid.cpp:12-27
//   Courant, listener cell    ProjectPlaneverb/src/FDTD/FDTD.cpp:90,97-99
//   AABB cell rectangle       ProjectPlaneverb/src/FDTD/Grid.cpp:139-142,252-255
//   analysis windows          ProjectPlaneverb/src/DSP/Analyzer.cpp:170-171,201-202,237,284
//   free-field probe          ProjectPlaneverb/src/FDTD/FreeGrid.cpp:78-91,99
//   emitter lookup cell       ProjectPlaneverb/src/DSP/Analyzer.cpp:110-113
#pragma once
#include <cmath>
#include <vector>
#include "../../include/planeverb_cuda.h"

namespace pvhost
{
    // physical constants that are part of Planeverb's contract (include/PvTypes.h:83-101)
    constexpr float kSpeedOfSound = 343.21f;
    constexpr float kPointsPerWavelength = 3.5f;
    constexpr float kFluxWindowS = 0.005f;
    constexpr float kDryWindowS = 0.01f;
    constexpr float kWetWindowS = 0.080f;
    constexpr float kSchroederTailS = 0.01f;
    constexpr float kSqrt2 = 1.4142136f;
    constexpr float kImpulseResponseS = kSqrt2 * 12.5f / kSpeedOfSound + 0.25f;

    struct GridParams
    {
        float dx = 0, dt = 0, courant = 0;
        unsigned fs = 0;
        int gx = 0, gy = 0, T = 0;
        int resolution = 0;
        int fluxSamples = 0, drySamples = 0, wetSamples = 0, tailSamples = 0;
        // FreeGrid probe
        int freeListenerR = 0, freeListenerC = 0, freeEmitterR = 0, freeEmitterC = 0, freeSamples = 0;
        float freeRadius = 0;
    };

    inline GridParams derive(int resolution, float sizeX, float sizeY, int responseLengthOverride = 0)
    {
        GridParams g;
        g.resolution = resolution;
        const float minWavelength = kSpeedOfSound / (float)resolution;
        g.dx = minWavelength / kPointsPerWavelength;
        g.dt = g.dx / (kSpeedOfSound * 1.5f);
        g.fs = (unsigned)(1.0f / g.dt);
        const float cellsX = (1.f / g.dx) * sizeX;
        const float cellsY = (1.f / g.dx) * sizeY;
        g.gx = (int)cellsX;
        g.gy = (int)cellsY;
        g.T = (int)(unsigned)((float)g.fs * kImpulseResponseS);
        if (responseLengthOverride > 0) g.T = responseLengthOverride;
        g.courant = kSpeedOfSound * g.dt / g.dx;
        const float fsf = (float)g.fs;
        g.fluxSamples = (int)(kFluxWindowS * fsf);
        g.drySamples = (int)(kDryWindowS * fsf);
        g.wetSamples = (int)(kWetWindowS * fsf);
        g.tailSamples = (int)(kSchroederTailS * fsf);
        // free-field probe: source at the grid centre re-derived through world metres, probe ~1 m along +x
        const int cr = g.gx / 2, cc = g.gy / 2;
        g.freeEmitterR = cr + (int)(1.f / g.dx);
        g.freeEmitterC = cc;
        g.freeListenerR = (int)(((float)cr * g.dx + 0.f) / g.dx);
        g.freeListenerC = (int)(((float)cc * g.dx + 0.f) / g.dx);
        g.freeSamples = (int)(kDryWindowS * (float)(int)g.fs) + (int)((1.f / kSpeedOfSound) * (float)(int)g.fs);
        g.freeRadius = (float)(g.freeEmitterR - cr) * g.dx;
        return g;
    }

    inline void gaussianPulse(int resolution, unsigned fs, std::vector<float>& out, int n)
    {
        out.resize((size_t)n);
        const float maxFreq = (float)resolution;
        const float pi = (float)std::acos(-1.0);
        const float sigma = (float)(1.0f / (0.5 * (double)pi * (double)maxFreq));
        const float delay = 2 * sigma;
        const float dt = 1.0f / (float)fs;
        for (int i = 0; i < n; ++i)
        {
            const float t = (float)(unsigned)i * dt;
            out[(size_t)i] = std::exp(-(t - delay) * (t - delay) / (sigma * sigma));
        }
    }

    inline pvc_listener listenerFor(const GridParams& g, float x, float z, float offX = 0.f, float offY = 0.f)
    {
        pvc_listener l;
        const float lx = x + offX, lz = z + offY;
        l.cell_r = (int)(lx / g.dx);
        l.cell_c = (int)(lz / g.dx);
        l.efree_r = (int)(lx * (1.f / g.dx));
        l.efree_c = (int)(lz * (1.f / g.dx));
        l.x = lx;
        l.z = lz;
        return l;
    }

    // rows come from position.x -/+ width/2, columns from position.y -/+ height/2
    inline pvc_rect rectFor(const GridParams& g, float px, float py, float width, float height,
                            float absorption, bool add, float offX = 0.f, float offY = 0.f)
    {
        pvc_rect q;
        const float inv = 1.f / g.dx;
        // the reference adds offset.x to the y extent and offset.y to the x extent in AddAABB and the
        // other way round in RemoveAABB (Grid.cpp:139-142 vs 252-255); offsets are unsupported (always 0)
        const float oy = add ? offX : offY;
        const float ox = add ? offY : offX;
        q.c0 = (int)((py - height / 2.f + oy) * inv);
        q.r0 = (int)((px - width / 2.f + ox) * inv);
        q.c1 = (int)((py + height / 2.f + oy) * inv);
        q.r1 = (int)((px + width / 2.f + ox) * inv);
        q.add = add ? 1 : 0;
        q.admittance = add ? (1.f - absorption) / (1.f + absorption) : 0.f;
        return q;
    }

    // Analyzer::GetResponseResult's cell; false when the reference returns nullptr.  The reference
    // accepts pos == gridSize and then indexes outside the lattice; we reject it.
    inline bool emitterCell(const GridParams& g, float x, float z, int& r, int& c, float offX = 0.f, float offY = 0.f)
    {
        const float fx = (x + offX) / g.dx, fz = (z + offY) / g.dx;
        if (!(fx >= 0.f) || !(fz >= 0.f)) return false;      // (unsigned) of a negative is UB in the reference
        if (fx >= 4294967296.f || fz >= 4294967296.f) return false;
        const unsigned ur = (unsigned)fx, uc = (unsigned)fz;
        if (ur >= (unsigned)g.gx || uc >= (unsigned)g.gy) return false;
        r = (int)ur; c = (int)uc;
        return true;
    }

    inline pvc_config configFor(const GridParams& g, int maxSources, int device, int stepKernel = 0)
    {
        pvc_config c{};
        c.gx = g.gx; c.gy = g.gy; c.T = g.T; c.fs = (int)g.fs; c.resolution = g.resolution;
        c.dx = g.dx; c.courant = g.courant;
        c.flux_samples = g.fluxSamples; c.dry_samples = g.drySamples;
        c.wet_samples = g.wetSamples; c.tail_samples = g.tailSamples;
        c.max_sources = maxSources; c.device = device; c.step_kernel = stepKernel;
        return c;
    }
} // namespace pvhost

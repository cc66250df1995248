// PvMathTypes.h -- source-compatible with ProjectPlaneverb/include/PvMathTypes.h (same type names,
// member names, layouts and macro values), written for this library.  The struct layouts are part of
// the drop-in contract: PlaneverbConfig is copied bytewise by clients (PvContext.cpp:110) and AABB /
// vec3 cross the API by pointer/reference (Planeverb.h:25-44).
#pragma once

// single precision throughout (PvMathTypes.h:4 of the reference); the CUDA path computes in fp32
#define Real float

namespace Planeverb
{
    struct vec3
    {
        Real x, y, z;
        vec3(Real x_ = 0.f, Real y_ = 0.f, Real z_ = 0.f) : x(x_), y(y_), z(z_) {}
    };

    struct vec2
    {
        union
        {
            struct { Real x, y; };
            Real m[2];
        };
        vec2(Real x_ = 0.f, Real y_ = 0.f) : x(x_), y(y_) {}
    };

    // axis-aligned box on the (x, z) plane: centre, full extents, and reflection coefficient
    // R = sqrt(1 - alpha).  width spans the first grid axis (rows), height the second (columns).
    struct AABB
    {
        vec2 position;
        Real width;
        Real height;
        Real absorption;
    };
} // namespace Planeverb

// Material presets, R = sqrt(1 - alpha).  Values are the public constants of the reference API
// (PvMathTypes.h:52-90) grouped here by absorption coefficient alpha.
#define PV_ABSORPTION_FREE_SPACE                ((Real)(0.000000000))   // alpha 1.00 (no wall)
// alpha 0.01
#define PV_ABSORPTION_TILE_GLAZED               ((Real)(0.994987437))
#define PV_ABSORPTION_WATER_SURFACE             ((Real)(0.994987437))
#define PV_ABSORPTION_MARBLE                    ((Real)(0.994987437))
#define PV_ABSORPTION_ICE                       ((Real)(0.994987437))
#define PV_ABSORPTION_SNOW_PACKED               ((Real)(0.994987437))
// alpha 0.02
#define PV_ABSORPTION_DEFAULT                   ((Real)(0.989949494))
#define PV_ABSORPTION_BRICK_PAINTED             ((Real)(0.989949494))
#define PV_ABSORPTION_CONCRETE_PAINTED          ((Real)(0.989949494))
// alpha 0.03
#define PV_ABSORPTION_GLASS_HEAVY               ((Real)(0.984885780))
#define PV_ABSORPTION_PLASTER_BRICK             ((Real)(0.984885780))
#define PV_ABSORPTION_WOOD_VARNISHED            ((Real)(0.984885780))
// alpha 0.04
#define PV_ABSORPTION_BRICK_UNGLAZED            ((Real)(0.979795897))
#define PV_ABSORPTION_CONCRETE                  ((Real)(0.979795897))
// alpha 0.05 .. 0.07
#define PV_ABSORPTION_PLASTER_CONCRETE_BLOCK    ((Real)(0.974679434))
#define PV_ABSORPTION_CONCRETE_ROUGH            ((Real)(0.969535971))
#define PV_ABSORPTION_GLASS                     ((Real)(0.969535971))
#define PV_ABSORPTION_CONCRETE_BLOCK_PAINTED    ((Real)(0.964365076))
#define PV_ABSORPTION_WOOD                      ((Real)(0.964365076))
// alpha 0.09 .. 0.17
#define PV_ABSORPTION_WOOD_PANEL                ((Real)(0.953939201))
#define PV_ABSORPTION_WOOD_PLYWOOD_PANEL        ((Real)(0.948683298))
#define PV_ABSORPTION_STEEL                     ((Real)(0.948683298))
#define PV_ABSORPTION_METAL                     ((Real)(0.948683298))
#define PV_ABSORPTION_GLASS_WINDOW              ((Real)(0.938083152))
#define PV_ABSORPTION_DRAPERY_LIGHT             ((Real)(0.921954446))
#define PV_ABSORPTION_DRAPERY                   ((Real)(0.921954446))
#define PV_ABSORPTION_CLOTH                     ((Real)(0.921954446))
#define PV_ABSORPTION_AWNING                    ((Real)(0.921954446))
#define PV_ABSORPTION_WOOD_TREE                 ((Real)(0.911043358))
#define PV_ABSORPTION_FOLIAGE                   ((Real)(0.911043358))
// alpha 0.35 .. 0.90
#define PV_ABSORPTION_CONCRETE_BLOCK_COARSE     ((Real)(0.806225775))
#define PV_ABSORPTION_CARPET_HEAVY              ((Real)(0.806225775))
#define PV_ABSORPTION_SOIL_ROUGH                ((Real)(0.741619849))
#define PV_ABSORPTION_DRAPERY_MEDIUM            ((Real)(0.670820393))
#define PV_ABSORPTION_DRAPERY_HEAVY             ((Real)(0.632455532))
#define PV_ABSORPTION_FIBERBOARD_SHREDDED_WOOD  ((Real)(0.632455532))
#define PV_ABSORPTION_GRAVEL                    ((Real)(0.547722558))
#define PV_ABSORPTION_GRASS                     ((Real)(0.547722558))
#define PV_ABSORPTION_SNOW_FRESH                ((Real)(0.316227766))

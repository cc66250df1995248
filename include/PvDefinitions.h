// PvDefinitions.h -- source-compatible stand-in for ProjectPlaneverb/include/PvDefinitions.h.
// Unlike the reference (which #errors outside Windows, PvDefinitions.h:4-6) this header builds on Linux.
#pragma once

#if defined(_WIN32)
  #if defined(PV_BUILD)
    #define PV_API __declspec(dllexport)
  #else
    #define PV_API __declspec(dllimport)
  #endif
  #define PV_FORCEINLINE __forceinline
#else
  #define PV_API __attribute__((visibility("default")))
  #define PV_FORCEINLINE inline __attribute__((always_inline))
#endif
#define PV_INLINE inline

#if defined(_DEBUG) && defined(_WIN32)
  #define PV_ASSERT(cond) do { if (!(cond)) __debugbreak(); } while (0)
#else
  #define PV_ASSERT(cond) ((void)0)
#endif

// row-major helpers with the reference's semantics: the stride is the x extent of `dim`
#define INDEX(row, col, dim)              ((row) * ((unsigned)(dim).x) + (col))
#define INDEX_TO_POS(ISET, JSET, i, dim)  (ISET) = (i) / (unsigned)(dim).x; (JSET) = (i) % (unsigned)(dim).x
#define INDEX3(row, col, t, dim, maxT)    ((t) + (maxT) * (INDEX((row), (col), (dim))))

// MSVC's <cmath> puts the C float functions into namespace std (PvDSPContext.cpp:287 calls std::atan2f); libstdc++ 13 does
// not.  Forced into the reference's translation units by the Makefile (-include); nothing of the reference is edited.
#pragma once
#include <cmath>
namespace std { using ::atan2f; using ::sqrtf; using ::powf; using ::cosf; using ::sinf; using ::fabsf; }

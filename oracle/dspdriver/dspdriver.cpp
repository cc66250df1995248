// dspdriver.cpp -- C entry points around the UNMODIFIED reference consumer of the hot path's outputs,
// PlaneverbDSP (PlaneverbDSP/src/PvDSPContext.cpp: Context::SubmitSource :250-425 with its validity gates :258-262
// and FindGainA/B/C :165-229; DSP/Lowpass.cpp), compiled in place from /root/reference by the Makefile beside this
// file into oracle/_ref/libpvdspref.so.  TEST INFRASTRUCTURE ONLY: the tests feed the device's
// PlaneverbOutput values and the reference's own golden outputs through this same code and compare what comes out.
// Nothing of the reference is copied here; this file only calls its public API (PlaneverbDSP/include/PlaneverbDSP.h).
#include <cstring>
#include "PlaneverbDSP.h"

extern "C" {

// One emitter through a fresh DSP context: `calls` audio callbacks of numFrames stereo frames with the same acoustic
// parameters out8 = {obstructionGain, wetGain, rt60, lowpass, direction.xy, sourceDirectivity.xy} (PvDSPTypes.h:69-77,
// the layout of Planeverb's PlaneverbOutput).  Writes the four interleaved stereo output buffers of the LAST callback
// (dry, reverb A / B / C, each 2 * numFrames floats).  Returns 0, or -1 on a bad argument / failed Init.
int pvdsp_render(const float* out8, float emitterX, float emitterZ, float listenerX, float listenerZ,
                 unsigned samplingRate, const float* audioIn, unsigned numFrames, int calls,
                 float* dry, float* revA, float* revB, float* revC)
{
    if (!out8 || !audioIn || !dry || !revA || !revB || !revC || numFrames == 0 || numFrames > 4096 || calls < 1) return -1;
    PlaneverbDSP::PlaneverbDSPConfig config;
    config.samplingRate = samplingRate;
    config.maxCallbackLength = (unsigned short)numFrames;
    try { PlaneverbDSP::Init(&config); }
    catch (...) { return -1; }
    PlaneverbDSP::SetListenerTransform(listenerX, 0.f, listenerZ, 1.f, 0.f, 0.f);
    PlaneverbDSP::UpdateEmitter(0, emitterX, 0.f, emitterZ, 1.f, 0.f, 0.f);
    PlaneverbDSP::PlaneverbDSPInput in;
    in.obstructionGain = out8[0]; in.wetGain = out8[1]; in.rt60 = out8[2]; in.lowpass = out8[3];
    in.direction = PlaneverbDSP::vec2(out8[4], out8[5]);
    in.sourceDirectivity = PlaneverbDSP::vec2(out8[6], out8[7]);
    float *d = nullptr, *a = nullptr, *b = nullptr, *c = nullptr;
    for (int k = 0; k < calls; ++k)
    {
        PlaneverbDSP::GetOutput(&d, &a, &b, &c);            // swaps and clears the double buffers (PvDSPContext.cpp:427-)
        PlaneverbDSP::SendSource(0, &in, audioIn, numFrames);
    }
    // the buffers SendSource just wrote are the ones GetOutput hands out next
    PlaneverbDSP::GetOutput(&d, &a, &b, &c);
    const size_t bytes = sizeof(float) * 2 * (size_t)numFrames;
    std::memcpy(dry, d, bytes); std::memcpy(revA, a, bytes); std::memcpy(revB, b, bytes); std::memcpy(revC, c, bytes);
    PlaneverbDSP::Exit();
    return 0;
}

}

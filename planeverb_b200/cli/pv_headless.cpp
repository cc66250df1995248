// pv_headless -- a command-line stand-in for the PlaneverbSandbox editor: loads a .pv scene
// (text: count, then `id posX posY width height absorption` per object -- the format PlaneverbSandbox/src/Editor/
// Editor.cpp:219-281 reads and writes), drives the acoustics module ONLY through the public C++ API of
// include/Planeverb.h exactly as the Sandbox does (Init, AddGeometry, SetListenerPosition, Emit, GetOutput,
// GetImpulseResponse: main.cpp:14-21, Editor.cpp:36-48,408,457), waits for analysed frames and prints the outputs.
// It is the link-unchanged demonstration for the C++ API; nothing here touches CUDA or the thin C layer underneath.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/Planeverb.h"
#include "../../include/PlaneverbUnity.h"       // only for the PlaneverbFramesCompleted extension

int main(int argc, char** argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s scene.pv [--size metres] [--res 275|375|500|750] [--listener x z] [--emitter x z]... "
                             "[--frames n] [--move id dx dy] [--ir] [--save out.pv]\n", argv[0]);
        return 2;
    }
    float size = 25.f;                                  // Sandbox world (main.cpp:17)
    int resolution = Planeverb::pv_LowResolution;       // Sandbox default (main.cpp:15)
    Planeverb::vec3 listener(5.f, 0.f, 4.f);            // Editor.cpp:36
    std::vector<Planeverb::vec3> emitters;
    int frames = 2, moveId = -1;
    float moveDx = 0.f, moveDy = 0.f;
    bool printIr = false;
    const char* savePath = nullptr;
    for (int i = 2; i < argc; ++i)
    {
        if (!std::strcmp(argv[i], "--size") && i + 1 < argc) size = (float)std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--res") && i + 1 < argc) resolution = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--listener") && i + 2 < argc) { listener.x = (float)std::atof(argv[++i]); listener.z = (float)std::atof(argv[++i]); }
        else if (!std::strcmp(argv[i], "--emitter") && i + 2 < argc) { float x = (float)std::atof(argv[++i]); float z = (float)std::atof(argv[++i]); emitters.push_back(Planeverb::vec3(x, 0.f, z)); }
        else if (!std::strcmp(argv[i], "--frames") && i + 1 < argc) frames = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--move") && i + 3 < argc) { moveId = std::atoi(argv[++i]); moveDx = (float)std::atof(argv[++i]); moveDy = (float)std::atof(argv[++i]); }
        else if (!std::strcmp(argv[i], "--ir")) printIr = true;
        else if (!std::strcmp(argv[i], "--save") && i + 1 < argc) savePath = argv[++i];
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (emitters.empty()) emitters.push_back(Planeverb::vec3(5.f, 0.f, 6.f));

    std::ifstream in(argv[1]);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    int count = 0;
    in >> count;
    std::vector<Planeverb::AABB> boxes;
    for (int i = 0; i < count; ++i)
    {
        int id; Planeverb::AABB b;
        in >> id >> b.position.x >> b.position.y >> b.width >> b.height >> b.absorption;
        if (!in) { std::fprintf(stderr, "%s: malformed object %d\n", argv[1], i); return 1; }
        boxes.push_back(b);
    }

    Planeverb::PlaneverbConfig config;
    config.gridSizeInMeters = Planeverb::vec2(size, size);
    config.gridResolution = resolution;
    config.gridBoundaryType = Planeverb::pv_AbsorbingBoundary;
    config.tempFileDirectory = ".";
    config.maxThreadUsage = 0;
    config.threadExecutionType = Planeverb::pv_GPU;
    try { Planeverb::Init(&config); }
    catch (Planeverb::PlaneverbErrorCode e)
    {
        std::fprintf(stderr, "Planeverb::Init threw %s: %s\n", e == Planeverb::pv_InvalidConfig ? "pv_InvalidConfig" : "pv_NotEnoughMemory", PlaneverbLastError());
        return 1;
    }
    Planeverb::SetListenerPosition(listener);
    std::vector<Planeverb::PlaneObjectID> ids;
    for (const auto& b : boxes) ids.push_back(Planeverb::AddGeometry(&b));
    std::vector<Planeverb::EmissionID> eids;
    for (const auto& e : emitters) eids.push_back(Planeverb::Emit(e));

    // wait for n more frames; gives up when the acoustics thread has stopped (device failure) or makes no progress for 20 s
    // (e.g. a listener outside the grid: its frames are skipped)
    auto waitFrames = [](unsigned long long n) {
        const unsigned long long start = PlaneverbFramesCompleted();
        unsigned long long seen = start;
        auto lastProgress = std::chrono::steady_clock::now();
        while (PlaneverbFramesCompleted() < start + n)
        {
            const unsigned long long now = PlaneverbFramesCompleted();
            if (now != seen) { seen = now; lastProgress = std::chrono::steady_clock::now(); }
            const bool stalled = std::chrono::steady_clock::now() - lastProgress > std::chrono::seconds(20);
            if (PlaneverbWorkerState() != 1 || stalled)
            {
                std::fprintf(stderr, "pv_headless: no frames (%s): %s\n", stalled ? "no progress for 20 s" : "acoustics thread stopped", PlaneverbLastError());
                Planeverb::Exit();
                std::exit(2);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
    };
    const auto t0 = std::chrono::steady_clock::now();
    waitFrames((unsigned long long)frames + 1);          // geometry queued now reaches the grid before the next solve
    if (moveId >= 0 && moveId < (int)boxes.size())
    {
        Planeverb::AABB b = boxes[(size_t)moveId];
        b.position.x += moveDx; b.position.y += moveDy;
        Planeverb::UpdateGeometry(ids[(size_t)moveId], &b);
        waitFrames(3);
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::printf("scene %s objects %d size %.3f res %d listener %.4f %.4f frames %llu (%.1f ms/frame)\n", argv[1], count, size, resolution,
                listener.x, listener.z, PlaneverbFramesCompleted(), 1e3 * secs / (double)PlaneverbFramesCompleted());
    for (size_t i = 0; i < eids.size(); ++i)
    {
        const Planeverb::PlaneverbOutput o = Planeverb::GetOutput(eids[i]);
        std::printf("emitter %zu at %.4f %.4f : occlusion %.9g wetGain %.9g rt60 %.9g lowpass %.9g direction %.9g %.9g sourceDirectivity %.9g %.9g\n",
                    i, emitters[i].x, emitters[i].z, o.occlusion, o.wetGain, o.rt60, o.lowpass, o.direction.x, o.direction.y,
                    o.sourceDirectivity.x, o.sourceDirectivity.y);
    }
    if (savePath)
    {
        // the Sandbox's scene writer (PlaneverbSandbox/src/Editor/Editor.cpp:219-243): object count, then one line per object
        // "id posX posY width height absorption" with the ids AddGeometry handed out -- after the optional --move, so the file
        // holds the scene as it was last solved
        std::ofstream out(savePath);
        if (!out) { std::fprintf(stderr, "cannot write %s\n", savePath); Planeverb::Exit(); return 1; }
        out << boxes.size() << std::endl;
        for (size_t i = 0; i < boxes.size(); ++i)
        {
            Planeverb::AABB b = boxes[i];
            if ((int)i == moveId) { b.position.x += moveDx; b.position.y += moveDy; }
            out << ids[i] << " " << b.position.x << " " << b.position.y << " " << b.width << " " << b.height << " " << b.absorption << std::endl;
        }
    }
    if (printIr)
    {
        const auto ir = Planeverb::GetImpulseResponse(emitters[0]);
        std::printf("ir samples %u\n", ir.second);
        for (unsigned t = 0; t < ir.second && t < 64; ++t) std::printf("ir %u %.9g %.9g %.9g\n", t, ir.first[t].pr, ir.first[t].vx, ir.first[t].vy);
    }
    Planeverb::Exit();
    return 0;
}

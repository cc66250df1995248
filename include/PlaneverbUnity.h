/* PlaneverbUnity.h -- the plain-C ABI Unity P/Invokes ("ProjectPlaneverbUnityPlugin",
 * ProjectPlaneverb/PlaneverbUnityPluginAPI/PlaneverbContext.cs:23-60), declared here for C/C++/ctypes
 * clients.  Same names, argument order and return types as
 * ProjectPlaneverb/PlaneverbUnityPluginAPI/PlaneverbUnity.cpp:12-135. */
#ifndef PLANEVERB_UNITY_H
#define PLANEVERB_UNITY_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PVU_CC __stdcall
#define PVU_EXPORT __declspec(dllexport)
#else
#define PVU_CC
#define PVU_EXPORT __attribute__((visibility("default")))
#endif

typedef struct PlaneverbUnityOutput
{
    float occlusion;
    float wetGain;
    float rt60;
    float lowpass;
    float directionX;
    float directionY;
    float sourceDirectionX;
    float sourceDirectionY;
} PlaneverbUnityOutput;

PVU_EXPORT void PVU_CC UnityPluginLoad(void* unityInterfaces);
PVU_EXPORT void PVU_CC UnityPluginUnload(void);

PVU_EXPORT void PVU_CC PlaneverbInit(float gridSizeX, float gridSizeY, int gridResolution, int gridBoundaryType,
                                     char* tempFileDir, int maxThreadUsage, int threadExecutionType);
PVU_EXPORT void PVU_CC PlaneverbExit(void);
PVU_EXPORT int  PVU_CC PlaneverbEmit(float x, float y, float z);
PVU_EXPORT void PVU_CC PlaneverbUpdateEmission(int id, float x, float y, float z);
PVU_EXPORT void PVU_CC PlaneverbEndEmission(int id);
PVU_EXPORT PlaneverbUnityOutput PVU_CC PlaneverbGetOutput(int emissionID);
PVU_EXPORT int  PVU_CC PlaneverbAddGeometry(float posX, float posY, float width, float height, float absorption);
PVU_EXPORT void PVU_CC PlaneverbUpdateGeometry(int id, float posX, float posY, float width, float height, float absorption);
PVU_EXPORT void PVU_CC PlaneverbRemoveGeometry(int id);
PVU_EXPORT void PVU_CC PlaneverbSetListenerPosition(float x, float y, float z);

/* extensions (not in the reference):
 *   PlaneverbFramesCompleted  completed background solve+analyse frames since Init, so a host can wait for the first frame
 *                             instead of polling GetOutput;
 *   PlaneverbWorkerState      0 no context, 1 the acoustics thread is running, 2 stopped by Exit, -1 stopped by a device
 *                             failure (GetOutput then keeps serving the last good frame);
 *   PlaneverbLastError        process-wide text of the last failure, including the acoustics thread's (empty if none); the
 *                             pointer stays valid until the calling thread asks again;
 *   PlaneverbHistorySteps     samples of pressure history the context's solver keeps: 0 = the whole response, > 0 = the streamed
 *                             solver (a grid whose full history does not fit the device, or PLANEVERB_HISTORY_STEPS), -1 no context. */
PVU_EXPORT unsigned long long PVU_CC PlaneverbFramesCompleted(void);
PVU_EXPORT int PVU_CC PlaneverbHistorySteps(void);
PVU_EXPORT int PVU_CC PlaneverbWorkerState(void);
PVU_EXPORT const char* PVU_CC PlaneverbLastError(void);

#ifdef __cplusplus
}
#endif
#endif

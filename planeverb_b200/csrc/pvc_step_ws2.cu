// pvc_step_ws2.cu -- second-generation warp-specialised generational step kernel (variants 39/40).
//
// Same numerics, work-item order and dependency protocol as fusedStepWsKernel (pvc_step_fused.cu): K = 4 time
// steps of Grid::GenerateResponseCPU (ProjectPlaneverb/src/FDTD/FDTD.cpp:122-235) per tile pass with the state
// in registers, tiles of (NW*R) x 128 cells incl. a 4-cell halo, one persistent CTA per SM pulling (generation,
// tile) items from a global counter, per-tile completed-generation counters instead of a grid-wide barrier.
// What changed, all of it read off the ncu source-level stall profile of the first kernel
// (profiles/r01_ws_stalls.txt: 19 % of the compute warps' time was long-scoreboard waits in the per-tile
// prologue/epilogue, 4 % CTA barriers outside the step loop):
//
//   * The compute warps issue NO global loads.  Everything a tile needs besides its state -- path mode per warp,
//     activity hint per warp, the source cell, the four pulse samples, the output buffer parity -- is fetched by
//     the producer warp one tile ahead and handed over in a small shared-memory record (double-buffered by tile
//     parity) that becomes visible with the TMA "full" barrier.
//   * The general (wall) path reads its coefficients from shared memory that the PRODUCER fills with TMA: two 2-D
//     tensor copies (gx, gy planes) plus one 1-D bulk copy of a per-(tile, warp, lane) bit mask of the air flag bp,
//     on their own mbarrier, CB-deep buffered (CB = 2: the coefficients of the next wall tile land while the
//     current one computes).  bp travels as 16 bits per thread instead of a third fp32 plane.
//   * No kernel parameter is indexed dynamically (the first kernel's G.state[(gen+1)&1][f] went through local
//     memory); buffers are picked with selects.
//   * Per tile the only CTA-wide synchronisation left is the two named barriers per time step.  The stage drain and
//     the "tile stored" hand-offs are per-warp mbarrier arrivals (count NW), the first vertical halo row of a pass is
//     read straight from the TMA stage instead of being exchanged, and the last step's exchange write is skipped.
//   * One step loop per path (fast / edge / general) instead of a three-way branch inside one loop; the activity
//     OR is skipped once a warp's block is known to be active.
//
// Roofline: HBM; algorithmic bytes 28 B per cell-update.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "pvc_internal.h"

namespace pvc
{
    namespace ws2
    {
        // ---------------------------------------------------------------- small PTX helpers
        __device__ __forceinline__ uint32_t smemAddr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
        __device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
        }
        __device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
        }
        __device__ __forceinline__ void mbarArrive(uint64_t* bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
        }
        // bounded wait: returns false (and raises the abort flag) instead of hanging the GPU on a protocol error
        __device__ __forceinline__ bool mbarWaitBounded(uint64_t* bar, uint32_t parity, int* abortFlag)
        {
            for (unsigned spins = 0;; ++spins)
            {
                uint32_t ready;
                asm volatile("{\n.reg .pred r;\nmbarrier.try_wait.parity.shared::cta.b64 r, [%1], %2;\nselp.u32 %0, 1, 0, r;\n}\n"
                             : "=r"(ready) : "r"(smemAddr(bar)), "r"(parity) : "memory");
                if (ready) return true;
                if ((spins & 0x3ffu) == 0x3ffu && (spins > (1u << 24) || *(volatile int*)abortFlag)) { atomicExch(abortFlag, 1); return false; }
            }
        }
        __device__ __forceinline__ void tmaLoad3d(void* dstSmem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
        {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smemAddr(dstSmem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smemAddr(bar)) : "memory");
        }
        __device__ __forceinline__ void tmaLoad2d(void* dstSmem, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
        {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smemAddr(dstSmem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smemAddr(bar)) : "memory");
        }
        __device__ __forceinline__ void bulkLoad1d(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
        {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
        }
        // ---- TMA stores (shared -> global tensor tiles, completion tracked in per-thread bulk async-groups)
        __device__ __forceinline__ void tmaStore3d(const CUtensorMap* map, const void* srcSmem, int c0, int c1, int c2)
        {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smemAddr(srcSmem)) : "memory");
        }
        __device__ __forceinline__ void tmaStore5dHint(const CUtensorMap* map, const void* srcSmem, int c0, int c1, int c2, int c3, int c4, uint64_t policy)
        {
            asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4, %5}], [%6], %7;"
                         ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smemAddr(srcSmem)), "l"(policy) : "memory");
        }
        __device__ __forceinline__ void bulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
        template <int N> __device__ __forceinline__ void bulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
        template <int N> __device__ __forceinline__ void bulkWait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
        __device__ __forceinline__ void fenceAsyncShared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
        __device__ __forceinline__ uint64_t evictFirstPolicy()
        {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            return pol;
        }
        __device__ __forceinline__ bool mbarTest(uint64_t* bar, uint32_t parity)
        {
            uint32_t ready;
            asm volatile("{\n.reg .pred r;\nmbarrier.test_wait.parity.shared::cta.b64 r, [%1], %2;\nselp.u32 %0, 1, 0, r;\n}\n"
                         : "=r"(ready) : "r"(smemAddr(bar)), "r"(parity) : "memory");
            return ready != 0u;
        }
        __device__ __forceinline__ int loadAcquire(const int* p)
        {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
            return v;
        }
        __device__ __forceinline__ void storeRelease(int* p, int v)
        {
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
        }
        template <int NW>
        __device__ __forceinline__ void computeBarrier()
        {
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
        }
        // Neighbour-only synchronisation of the step loop (PVC_PAIR_SYNC): a warp exchanges halo rows only with the warp above
        // and the warp below, so instead of a CTA-wide barrier per sub-step it meets each neighbour on the named barrier of
        // their common edge (id = upper warp + 1; 64 threads).  Even warps take the lower edge first, odd warps the upper
        // one, so every edge is the first meeting of both its warps or the second of both: no cycle.  Warps more than k
        // edges away from a slow (wall-path) warp may run k sub-steps ahead of it, finish their pass early and issue their
        // state stores / drain the next stage while the slow warps still compute.
    #ifndef PVC_PAIR_SYNC
        #define PVC_PAIR_SYNC 1
    #endif
        __device__ __forceinline__ void pairBarrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
        template <int NW, bool TS>
        __device__ __forceinline__ void phaseSync(int wp)
        {
            if (PVC_PAIR_SYNC && !TS && NW <= 16)
            {
                if (wp & 1) { pairBarrier(wp); if (wp + 1 < NW) pairBarrier(wp + 1); }
                else { if (wp + 1 < NW) pairBarrier(wp + 1); if (wp > 0) pairBarrier(wp); }
            }
            else computeBarrier<NW>();
        }
        __device__ __forceinline__ bool isAirF(float w) { return __float_as_uint(w) == kAirBits; }
        // 1.0f where x < 0 (one FSET): the k of the linear-form velocity rule, folded into the sign of its coefficient
        __device__ __forceinline__ float signFlag(float x)
        {
            float r;
            asm("set.lt.f32.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
            return r;
        }

        // ---------------------------------------------------------------- kernel arguments (never indexed dynamically)
        struct Args
        {
            float* p0; float* vx0; float* vy0;     // ping-pong buffer 0 (generation g reads buffer g & 1, writes the other)
            float* p1; float* vx1; float* vy1;
            float* hist;                           // pressure history (null: no record)
            const uint32_t* mode;                  // [tile][32] path mode per warp (slowMaskKernel)
            const uint32_t* bpMask;                // [tile][NW][32] bit j*4+k = cell (j,k) of the thread is an interior air cell
            const int* tileOrder;
            int* firstActive;                      // [source][tile][32]
            const SourceParams* src;
            const float* pulse;
            int* doneGen;                          // [source][tile] generations completed
            int* workCounter;
            int* abortFlag;
            int tilesPerSource, nsrc, numTiles;    // numTiles = tilesPerSource * nsrc
            int gen0, numGen, T;
            int earlyFetch;
            int lateRelease;                       // release a tile's generation counter after the next TMA has been issued (1, default)
            int fenceMode;                         // cross-proxy fence: 0 reader side after the publish, 1 reader side right after the early probe, 2 writer side (with the release)
            int slowPathPoll;                      // dependency miss: poll deps and own tile together (1, default) or publish own tile first (0)
            int tsDebug;                           // debug: bit 0 no history TMA, bit 1 no state TMA, bit 2 never defer the done arrival
            unsigned long long* debug;             // optional counters (PVC_DEBUG_COUNTERS): tiles, slow-path hand-overs, cycles waiting / total
            int srcGroup, genChunk;                // item order: sources per L2-resident group, generations per chunk
            float courant;
            int finalPass;                         // the launch ends the response (0: a chunk of a streamed solve that another chunk follows)
            int band;                              // tile rows per L2-resident band of the item order (0: none; pvc_internal.h::Ws2Order)
        };
        // state: loads, [buffer][p,vx,vy], box 128 x tile rows; coef: gx, gy, bp; store: [buffer][p,vx,vy], box 120 x (tile's owned rows), clipped to the
        // alloc grid; hist: 5-D {120 columns, T, strips, rows, sources}, box 120 x 1 x 1 x owned rows x 1 (TS variants only)
        struct Maps { CUtensorMap state[6]; CUtensorMap coef[3]; CUtensorMap store[6]; CUtensorMap hist; };

        // producer -> compute hand-off record of one tile
        struct Meta
        {
            int valid, s, tx, ty, gen;
            int coefBuf, coefParity;               // coefBuf < 0: no general-path warp in this tile
            int srcR, srcC;
            int srcDead;                           // the pulse cell is not an interior air cell (SourceParams::dead)
            int pad[2];
            float pulse[4];
            int mode[32];
            int hint[32];
        };

        enum { kFast = 0, kEdge = 1, kGeneral = 2 };
        // Timeline / counter instrumentation (PVC_DEBUG_COUNTERS) is compiled in only with -DPVC_WS2_TRACE: even as dead
        // branches it costs the step loop ~3 % (registers, predicates).
    #ifdef PVC_WS2_TRACE
        #define PVC_DBG(A) ((A).debug)
    #else
        #define PVC_DBG(A) ((unsigned long long*)nullptr)
    #endif
        constexpr int kTraceTiles = 64, kTraceSlots = 8, kTraceCtas = 148;
        __device__ __forceinline__ void trace(unsigned long long* dbg, int seq, int slot)
        {
            if (dbg && seq >= 200 && seq < 200 + kTraceTiles && blockIdx.x < kTraceCtas)
            {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                dbg[4 + ((size_t)blockIdx.x * kTraceTiles + (seq - 200)) * kTraceSlots + slot] = t;
            }
        }

        // One time step of one warp's R x 128 block.  MODE is the warp-uniform path (see slowMaskKernel):
        //   fast     every cell of the warp, and each one's up/left neighbour, is interior air: 11 fp32 ops per cell
        //   edge     no wall, but the warp touches the grid edge / padding / guard band: position-only overwrites
        //   general  some cell or neighbour is a wall: coefficient planes gx, gy from shared memory, bp from a bit mask
        // All arithmetic is explicit round-to-nearest mul/add/sub in the reference's operation order (no FMA).
        // BPF (three staged coefficient planes, every product variant): the general path is the LINEAR FORM of pvc_step_res.cu --
        // p' = p - cP*div, v' = fma(k, v, -(c*(p - p_prev))) with k folded into the sign of c -- which costs what the fast path
        // costs plus one FSET per velocity component; the planes are sX, sY, cP (buildLinearKernel).  It relies on p == 0 in every
        // cell that is not interior air, so nothing may be injected there (Meta::srcDead).  Without BPF (two planes + a per-thread
        // bit mask of the air flag; experiment variants whose shared memory cannot hold three): the select form over gx, gy.
        template <int NW, int R, int MODE, bool BPF = false>
        struct Stepper
        {
            const Layout& L;
            const float C;
            const int lane, wp, rBase;
            const uint32_t colOut, colPad, colLeft;      // edge path: column classes of the thread's 4 cells
            const uint32_t bpBits;                       // general path: air flags of the thread's R x 4 cells
            const float4* cGx; const float4* cGy;        // general path: this thread's coefficient float4s, row stride 32
            const float4* cBp;                           // BPF: likewise for the air-flag plane

            __device__ __forceinline__ void pressure(float (&p)[R][4], const float (&vx)[R][4], const float (&vy)[R][4], const float4 vxBelow) const
            {
                const float vb[4] = { vxBelow.x, vxBelow.y, vxBelow.z, vxBelow.w };
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                    if (MODE == kEdge)
                    {
                        // Rows outside the lattice (and the padding row r == gx) hold p == 0 for ever: nothing to do (warp-uniform).
                        // Columns outside the lattice / the padding column are kept at zero by a per-column Courant factor of 0
                        // instead of a select per cell: 0 - 0 * div == +0 (div is finite), and the live columns see the same C.
                        const int r = rBase + j;
                        if (r < 0 || r >= L.gx) continue;
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float ck = (((colOut | colPad) >> k) & 1u) ? 0.f : C;
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            p[j][k] = __fsub_rn(p[j][k], __fmul_rn(ck, div));
                        }
                    }
                    else
                    {
                        float ck[4] = { C, C, C, C };
                        if (MODE == kGeneral && BPF) { const float4 c4 = cBp[j * 32]; ck[0] = c4.x; ck[1] = c4.y; ck[2] = c4.z; ck[3] = c4.w; }
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            const float pn = __fsub_rn(p[j][k], __fmul_rn(ck[k], div));     // cP = 0 where the cell is not interior air: p stays 0
                            if (MODE == kFast || BPF) p[j][k] = pn;
                            else p[j][k] = ((bpBits >> (j * 4 + k)) & 1u) ? pn : 0.f;
                        }
                    }
                }
            }

            __device__ __forceinline__ void velocity(const float (&p)[R][4], float (&vx)[R][4], float (&vy)[R][4], const float4 pAbove) const
            {
                const float pa[4] = { pAbove.x, pAbove.y, pAbove.z, pAbove.w };
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                    if (MODE == kFast)
                    {
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            vx[j][k] = __fsub_rn(vx[j][k], __fmul_rn(C, __fsub_rn(p[j][k], pu)));
                            vy[j][k] = __fsub_rn(vy[j][k], __fmul_rn(C, __fsub_rn(p[j][k], pl)));
                        }
                    }
                    else if (MODE == kEdge)
                    {
                        const int r = rBase + j;
                        const bool rowOut = (r < 0) || (r > L.gx);
                        const bool rowTop = (r == 0), rowPad = (r == L.gx);
                        const uint32_t colDeadX = colOut | colPad;          // vx: the padding column is never driven
                        if (rowOut) continue;                               // outside the alloc grid: zero for ever (warp-uniform)
                        if (!rowTop && !rowPad)
                        {
                            // interior row of an edge tile (warp-uniform branch): the fast-path arithmetic with the per-column
                            // Courant factor (dead and padding columns stay +0), then the two position-only overwrites of
                            // FDTD.cpp:220-221 on the ONE lane that holds column 0 / column gy
                            #pragma unroll
                            for (int k = 0; k < 4; ++k)
                            {
                                const float ck = ((colDeadX >> k) & 1u) ? 0.f : C;
                                const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                                const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                                vx[j][k] = __fsub_rn(vx[j][k], __fmul_rn(ck, __fsub_rn(p[j][k], pu)));
                                vy[j][k] = __fsub_rn(vy[j][k], __fmul_rn(ck, __fsub_rn(p[j][k], pl)));
                            }
                            if (colLeft & 1u) vy[j][0] = -p[j][0];                                       // column 0 is always a thread's first cell
                            if (colPad)
                            {
                                #pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if ((colPad >> k) & 1u) vy[j][k] = (k > 0) ? p[j][k - 1] : pLeft;
                            }
                            continue;
                        }
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            const float pt = p[j][k];
                            float nx = __fsub_rn(vx[j][k], __fmul_rn(C, __fsub_rn(pt, pu)));
                            float ny = __fsub_rn(vy[j][k], __fmul_rn(C, __fsub_rn(pt, pl)));
                            nx = rowTop ? -pt : nx;                                  // FDTD.cpp:208
                            nx = rowPad ? pu : nx;                                   // FDTD.cpp:209
                            nx = (rowOut || ((colDeadX >> k) & 1u)) ? 0.f : nx;
                            ny = ((colLeft >> k) & 1u) ? -pt : ny;                   // FDTD.cpp:220
                            ny = ((colPad >> k) & 1u) ? pl : ny;                     // FDTD.cpp:221
                            ny = (rowOut || rowPad || ((colOut >> k) & 1u)) ? 0.f : ny;
                            vx[j][k] = nx; vy[j][k] = ny;
                        }
                    }
                    else
                    {
                        // per-cell coefficients (buildCoefficientsKernel): walls, the absorbing grid edge, padding and the
                        // guard band are all data here, the code is straight-line
                        const float4 x4 = cGx[j * 32];
                        const float4 y4 = cGy[j * 32];
                        const float ga[4] = { x4.x, x4.y, x4.z, x4.w };
                        const float ha[4] = { y4.x, y4.y, y4.z, y4.w };
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            const float pt = p[j][k];
                            if (BPF)
                            {
                                // fma(k, v, -(|s| * d)): k * v is exact (k is 0 or 1), so this is the reference's v - c*d (k = 1) or -(c*d)
                                const float tx = __fmul_rn(fabsf(ga[k]), __fsub_rn(pt, pu));
                                const float ty = __fmul_rn(fabsf(ha[k]), __fsub_rn(pt, pl));
                                vx[j][k] = __fmaf_rn(signFlag(ga[k]), vx[j][k], -tx);
                                vy[j][k] = __fmaf_rn(signFlag(ha[k]), vy[j][k], -ty);
                            }
                            else
                            {
                                const bool air = ((bpBits >> (j * 4 + k)) & 1u) != 0u;
                                const float airX = __fsub_rn(vx[j][k], __fmul_rn(C, __fsub_rn(pt, pu)));
                                const float airY = __fsub_rn(vy[j][k], __fmul_rn(C, __fsub_rn(pt, pl)));
                                const float wallX = __fmul_rn(ga[k], air ? pt : pu);
                                const float wallY = __fmul_rn(ha[k], air ? pt : pl);
                                vx[j][k] = (air && isAirF(ga[k])) ? airX : wallX;
                                vy[j][k] = (air && isAirF(ha[k])) ? airY : wallY;
                            }
                        }
                    }
                }
            }
        };

        // per-thread, per-tile constants of the record / inject / exchange part of a step
        struct TileCtx
        {
            float* hist;              // sample t0 of the thread's 4 cells in row rBase (null: no record)
            size_t histRow;           // floats between rows of the history
            uint32_t ownRows;         // bit j: row j of this thread is owned (stored), see computeTile
            int sj, sk;               // pulse cell inside the thread's block (sj < 0: not here)
            bool onlyLast;            // pulse cell on the padding row / column: inject only the solve's last sample
            const float* pulse;       // 4 samples of this generation (shared memory)
            bool track;               // this warp's block is not yet known to be active: accumulate the activity OR
            // TMA-store variants (TS): the tile's stores go through three [VR][120] staging planes in shared memory and ONE
            // storer thread (warp 0, lane 0) that issues a tensor store per plane; see stepKernel
            float* out;               // staging planes (all warps see the same pointer)
            int outRow;               // first staging row of this warp (< 0: the warp owns no rows)
            bool storer;
            const Maps* maps;
            int tx, ty, s, t0, gen;
            bool histOn;
            uint64_t* deferredDone;   // storer: "tile stored" barrier of the PREVIOUS tile, still to be arrived on (null: none)
            uint64_t policy;
            int tsDebug;
        };

        // storer thread: one sample of the tile's VR owned rows.  A history strip is exactly the 120 owned columns of a tile
        // (Layout::hist_chunk == kValidCols for this variant), so the record is one dense [VR][120] box: strip tx, sample t, rows from the tile's first
        // owned row; rows past the grid are clipped by the tensor map.
        __device__ __forceinline__ void storeHistoryPlane(const TileCtx& X, const Layout& L, const float* plane, int t)
        {
            tmaStore5dHint(&X.maps->hist, plane, 0, t, X.tx, X.ty * L.valid_rows, X.s, X.policy);
        }
        __device__ __forceinline__ void storeStatePlane(const TileCtx& X, const Layout& L, const float* plane, int field)
        {
            const CUtensorMap* m = &X.maps->store[((X.gen & 1) ? 0 : 3) + field];    // generation g writes buffer (g + 1) & 1
            tmaStore3d(m, plane, X.tx * kValidCols + kGuardCols, X.ty * L.valid_rows + kGuardRows, X.s);
        }
        // an owning warp: its R rows of one field into a staging plane
        template <int R>
        __device__ __forceinline__ void stageRows(float* plane, int outRow, int lane, const float (&v)[R][4])
        {
            if (lane >= 1 && lane <= 30)
            {
                #pragma unroll
                for (int j = 0; j < R; ++j)
                    *reinterpret_cast<float4*>(plane + (outRow + j) * kValidCols + (lane - 1) * 4) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
            }
            fenceAsyncShared();          // generic-proxy writes above -> async-proxy reads of the TMA store (issued after a barrier)
        }
        template <int R>
        __device__ __forceinline__ void injectPulse(const TileCtx& X, float (&p)[R][4], float add)
        {
            // adding +0 to the three other cells of the row is exact (it can only turn -0 into +0)
            const float a0 = (X.sk == 0) ? add : 0.f, a1 = (X.sk == 1) ? add : 0.f;
            const float a2 = (X.sk == 2) ? add : 0.f, a3 = (X.sk == 3) ? add : 0.f;
            #pragma unroll
            for (int j = 0; j < R; ++j)
                if (j == X.sj)
                {
                    p[j][0] = __fadd_rn(p[j][0], a0); p[j][1] = __fadd_rn(p[j][1], a1);
                    p[j][2] = __fadd_rn(p[j][2], a2); p[j][3] = __fadd_rn(p[j][3], a3);
                }
        }

        template <int NW, int R, int MODE, bool TS, bool BPF = false>
        __device__ __forceinline__ void stepLoop(const Stepper<NW, R, MODE, BPF>& S, TileCtx& X, const int nsteps,
                                                 float (&p)[R][4], float (&vx)[R][4], float (&vy)[R][4], float4 vxBelow,
                                                 float4 (*sVxTop)[32], float4 (*sPBot)[32], uint32_t& activity)
        {
            const int lane = S.lane, wp = S.wp;
            constexpr int kPlane = (NW - 2) * R * kValidCols;          // floats per staging plane (TS)
            // TS: the pulse sample of a pass's LAST step is injected at the start of the next pass instead (pulse[0] of the
            // record is that pending sample): the stored state pressure is then identical to the last recorded sample and
            // both leave through the same staging plane.  Same additions on the same operands in the same order.
            if (TS && X.sj >= 0 && X.t0 > 0) injectPulse<R>(X, p, X.pulse[0]);
            #pragma unroll 1
            for (int step = 0; step < nsteps; ++step)
            {
                // ---- pressure sub-step (FDTD.cpp:125-141)
                if (step > 0) vxBelow = sVxTop[wp + 1][lane];
                S.pressure(p, vx, vy, vxBelow);
                sPBot[wp + 1][lane] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
                if (TS && X.storer)
                {
                    // the staging plane(s) written after this barrier must have been read by their previous TMA store.
                    // Groups in issue order: per pass g0..g2 (samples 0..2, planes 0..2), g3 (sample 3 = state p from plane 0,
                    // vx from plane 1), g4 (vy from plane 2); see the epilogue in stepKernel.
                    if (nsteps != kTileK) bulkWaitRead<0>();
                    else if (step == 0) bulkWaitRead<1>();            // plane 0: previous pass's g3; its g4 may still read
                    else if (step == 3) bulkWaitRead<1>();            // planes 0, 1: g0, g1; g2 may still read
                    else bulkWaitRead<2>();                           // plane 1 / 2: previous pass's g3 / g4
                }
                phaseSync<NW, TS>(wp);
                // ---- velocity sub-steps + edge overrides (FDTD.cpp:144-223)
                S.velocity(p, vx, vy, sPBot[wp][lane]);
                // ---- record sample t0 + step (FDTD.cpp:226-231), then inject (FDTD.cpp:234)
                const bool last = step + 1 == nsteps;
                if (TS)
                {
                    if (X.outRow >= 0)
                    {
                        stageRows<R>(X.out + ((step == 3) ? 0 : step) * kPlane, X.outRow, lane, p);
                        if (last) stageRows<R>(X.out + (((step == 3) ? 0 : step) + 1) % 3 * kPlane, X.outRow, lane, vx);
                    }
                }
                if (TS ? X.histOn : (X.hist != nullptr))
                {
                    if (!TS)
                    {
                        #pragma unroll
                        for (int j = 0; j < R; ++j)
                            if ((X.ownRows >> j) & 1u)
                                __stcs(reinterpret_cast<float4*>(X.hist + (size_t)j * X.histRow), make_float4(p[j][0], p[j][1], p[j][2], p[j][3]));
                        X.hist += kHistChunkDefault;           // compile-time stride (the launcher checks Layout::hist_chunk)
                    }
                    if (X.track)
                    {
                        #pragma unroll
                        for (int j = 0; j < R; ++j)
                        {
                            activity |= __float_as_uint(p[j][0]) | __float_as_uint(p[j][1]);
                            activity |= __float_as_uint(p[j][2]) | __float_as_uint(p[j][3]);
                        }
                    }
                }
                if (X.sj >= 0 && !(TS && last) && (!X.onlyLast || last)) injectPulse<R>(X, p, X.pulse[TS ? step + 1 : step]);
                if (!last) sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
                if (TS && X.storer && last) bulkWaitRead<0>();        // plane for vy (staged after the barrier): g2, one step old
                // also the write-after-read fence of sPBot for the next pass's first pressure sub-step
                phaseSync<NW, TS>(wp);
                if (TS && X.storer)
                {
                    const float* plane = X.out + ((step == 3) ? 0 : step) * kPlane;
                    if (X.histOn && !(X.tsDebug & 1)) storeHistoryPlane(X, S.L, plane, X.t0 + step);
                    if (last && !(X.tsDebug & 2))
                    {
                        storeStatePlane(X, S.L, plane, 0);                                             // the new state's pressure IS the last sample
                        storeStatePlane(X, S.L, X.out + (((step == 3) ? 0 : step) + 1) % 3 * kPlane, 1);
                    }
                    bulkCommit();
                    if (step == 0 && X.deferredDone)
                    {   // the previous pass's stores: everything but the group just committed has completed
                        bulkWait<1>();
                        mbarArrive(X.deferredDone);
                    }
                }
            }
        }

        template <int NW, int R, int CB, bool TS = false, bool SO = false>
        struct Smem
        {
            static constexpr int TR = NW * R;
            static constexpr uint32_t kPlaneBytes = TR * kTileCols * sizeof(float);
            static constexpr uint32_t kMaskBytes = NW * 32 * sizeof(uint32_t);
            // coefficient planes staged for a wall tile: gx, gy and -- where 227 KB of shared memory allow -- the air-flag plane
            // bp (otherwise the per-thread bit mask of bpMaskKernel)
            static constexpr size_t kOther = 3 * (size_t)TR * kTileCols * sizeof(float) + 2 * (size_t)(NW + 1) * 32 * sizeof(float4)
                                           + ((TS || SO) ? (size_t)3 * (NW - 2) * R * kValidCols * sizeof(float) : 0) + 2 * sizeof(Meta) + 256;
            static constexpr int NP = (kOther + (size_t)CB * (3 * (size_t)TR * kTileCols * sizeof(float) + NW * 32 * sizeof(uint32_t)) <= 232448) ? 3 : 2;
            static constexpr size_t offStage = 0;
            static constexpr size_t offCoef = offStage + 3 * (size_t)kPlaneBytes;
            static constexpr size_t offMask = offCoef + (size_t)CB * NP * kPlaneBytes;
            static constexpr size_t offVxTop = offMask + (size_t)CB * kMaskBytes;
            static constexpr size_t offPBot = offVxTop + (size_t)(NW + 1) * 32 * sizeof(float4);
            static constexpr size_t offOut = offPBot + (size_t)(NW + 1) * 32 * sizeof(float4);      // [3][(NW - 2) * R][120] (TS only)
            static constexpr uint32_t kOutPlaneBytes = (NW - 2) * R * kValidCols * sizeof(float);
            static constexpr size_t offMeta = offOut + ((TS || SO) ? (size_t)3 * kOutPlaneBytes : 0);
            static constexpr size_t offBars = offMeta + 2 * sizeof(Meta);
            static constexpr size_t offPub = offBars + (5 + CB) * sizeof(uint64_t);                  // publisher warp: ring[4][4], observed, total
            static constexpr size_t total = offPub + 24 * sizeof(int);
        };

        // SO ("state out", needs PUB): the new state of a tile leaves through three [owned rows][120] staging planes in shared
        // memory and three TMA tensor stores issued by the PUBLISHER warp, instead of 12 STG.128 per compute thread.  The
        // LSU path from an SM into the crossbar carries 32 B/clk (ncu: l1tex__m_l1tex2xbar_write_bytes peak); at that rate
        // the 69-86 KB of state per tile took 1.0-1.4 us per 6 us tile with the FP pipes idle, and competed with the
        // history stores inside the step loop.  The compute warps now spend ~0.3 us on STS and go on to the next tile.
        template <int NW, int R, int CB, bool TS, bool PUB, bool SO>
        __global__ void __launch_bounds__((NW + 1 + (PUB ? 1 : 0)) * 32, 1)
        stepKernel(const Layout L, const Args A, const __grid_constant__ Maps maps)
        {
            using SM = Smem<NW, R, CB, TS, SO>;
            static_assert(!SO || (PUB && !TS && R == kTileK), "state-out needs the publisher warp; halo rows = first and last warp");
            static_assert(!TS || R == kTileK, "TMA-store variants: the halo rows must be exactly the first and the last warp");
            constexpr int TR = SM::TR;
            static_assert(NW <= 31 && R * 4 <= 32, "Meta holds 32 warps; bpMask holds 32 cells per thread");
            extern __shared__ __align__(128) unsigned char smemRaw[];
            float* stage = reinterpret_cast<float*>(smemRaw + SM::offStage);                       // [3][TR][128]
            float4* sCoef = reinterpret_cast<float4*>(smemRaw + SM::offCoef);                      // [CB][NP][TR][32]: gx, gy(, bp)
            uint32_t* sMask = reinterpret_cast<uint32_t*>(smemRaw + SM::offMask);                  // [CB][NW][32]
            float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(smemRaw + SM::offVxTop);       // [w]   = vx of warp w's first row
            float4 (*sPBot)[32] = reinterpret_cast<float4 (*)[32]>(smemRaw + SM::offPBot);         // [w+1] = p of warp w's last row
            Meta* meta = reinterpret_cast<Meta*>(smemRaw + SM::offMeta);                           // [2]
            uint64_t* full = reinterpret_cast<uint64_t*>(smemRaw + SM::offBars);
            uint64_t* empty = full + 1;
            uint64_t* done = full + 2;                                                              // [2], by tile parity
            uint64_t* fullCoef = full + 4;                                                          // [CB]
            uint64_t* outFree = full + 4 + CB;                                                      // SO: the staging planes have been read by the TMA stores
            // PUB: a third role, the publisher warp (warp NW + 1).  It owns the "done" barriers: for every tile, in hand-over
            // order, it observes done(k), bumps pubObserved, executes the cross-proxy fence and releases the tile's generation
            // counter -- 2-4 us that no longer sit between a drained stage and the producer's next TMA.
            volatile int* pubRing = reinterpret_cast<volatile int*>(smemRaw + SM::offPub);          // [4][4] = {counter slot, generation, source, tile}
            volatile int* pubObserved = pubRing + 16;                                                // tiles whose done barrier has been consumed
            volatile int* pubTotal = pubRing + 17;                                                   // tiles handed out in all (-1: still running)

            const int lane = threadIdx.x & 31;
            const int wp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // warp-uniform and known to be (uniform registers, no divergence guards)
            const int total = A.numGen * A.numTiles;
            const int tps = A.tilesPerSource;

            if (threadIdx.x == 0)
            {
                mbarInit(full, 1); mbarInit(empty, NW); mbarInit(done, TS ? 1 : NW); mbarInit(done + 1, TS ? 1 : NW);
                for (int b = 0; b < CB; ++b) mbarInit(fullCoef + b, 1);
                mbarInit(outFree, 1);
                *pubObserved = 0; *pubTotal = -1;
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            if (wp == 0)
            {
                sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();

            if (PUB && wp == NW + 1)
            {
                // ================= publisher warp =================
                uint32_t par0 = 0u, par1 = 0u;
                for (int k = 0;; ++k)
                {
                    int st = 0;                                  // 1: done(k) observed, 2: no tile k (the producer has finished), -1: abort
                    if (lane == 0)
                    {
                        uint64_t* bar = done + (k & 1);
                        const uint32_t par = (k & 1) ? par1 : par0;
                        for (unsigned spins = 0; st == 0; ++spins)
                        {
                            if (mbarTest(bar, par)) { st = 1; break; }
                            const int tot = *pubTotal;
                            if (tot >= 0 && k >= tot) { st = 2; break; }
                            __nanosleep(20);
                            if ((spins & 0x3ffu) == 0x3ffu && (spins > (1u << 24) || *(volatile int*)A.abortFlag)) { atomicExch(A.abortFlag, 1); st = -1; }
                        }
                    }
                    st = __shfl_sync(0xffffffffu, st, 0);
                    if (st != 1) break;
                    if (k & 1) par1 ^= 1u; else par0 ^= 1u;
                    if (lane == 0)
                    {
                        const int slot = pubRing[(k & 3) * 4], g = pubRing[(k & 3) * 4 + 1];
                        *pubObserved = k + 1;
                        if (SO)
                        {
                            // "done" here means "staged": every compute warp has written its rows of the new state into the
                            // staging planes and executed the generic->async proxy fence.  Generation g writes buffer (g + 1) & 1.
                            const int s = pubRing[(k & 3) * 4 + 2], id = pubRing[(k & 3) * 4 + 3];
                            const int ty = id / L.tiles_x, tx = id - ty * L.tiles_x;
                            const float* planes = reinterpret_cast<const float*>(smemRaw + SM::offOut);
                            constexpr int kPlane = (NW - 2) * R * kValidCols;
                            const CUtensorMap* m = &maps.store[(g & 1) ? 0 : 3];
                            tmaStore3d(m + 0, planes, tx * kValidCols + kGuardCols, ty * L.valid_rows + kGuardRows, s);
                            tmaStore3d(m + 1, planes + kPlane, tx * kValidCols + kGuardCols, ty * L.valid_rows + kGuardRows, s);
                            tmaStore3d(m + 2, planes + 2 * kPlane, tx * kValidCols + kGuardCols, ty * L.valid_rows + kGuardRows, s);
                            bulkCommit();
                            bulkWaitRead<0>();
                            mbarArrive(outFree);                                  // the compute warps may stage the next tile
                            bulkWait<0>();                                        // writes complete: the release below publishes them
                        }
                        else asm volatile("fence.proxy.async;" ::: "memory");    // writer-side cross-proxy fence (see fenceMode 2)
                        storeRelease(A.doneGen + slot, g + 1);
                    }
                    __syncwarp();
                }
                return;
            }
            if (wp == NW)
            {
                // ================= producer warp =================
                auto waitObserved = [&](int need) -> bool {      // PUB: until the publisher has consumed the done barrier of `need` tiles
                    int ok = 1;
                    if (lane == 0)
                        for (unsigned spins = 0; *pubObserved < need; ++spins)
                        {
                            __nanosleep(20);
                            if ((spins & 0x3ffu) == 0x3ffu && (spins > (1u << 24) || *(volatile int*)A.abortFlag)) { atomicExch(A.abortFlag, 1); ok = 0; break; }
                        }
                    return __shfl_sync(0xffffffffu, ok, 0) != 0;
                };
                auto depsReady = [&](int s, int tx, int ty, int gen, bool block) -> bool {
                    if (gen == 0) return true;                 // generation 0 reads the zeroed state: written by a memset before the launch
                    const int dx = lane % 3 - 1, dy = lane / 3 - 1;
                    const int nx = tx + dx, ny = ty + dy;
                    const bool mine = lane < 9 && nx >= 0 && ny >= 0 && nx < L.tiles_x && ny < L.tiles_y;
                    const int* slot = A.doneGen + (size_t)s * tps + (mine ? ny * L.tiles_x + nx : 0);
                    unsigned spins = 0;
                    while (true)
                    {
                        const bool ok = !mine || loadAcquire(slot) >= gen;
                        if (__all_sync(0xffffffffu, ok)) return true;     // the caller still owes a proxy fence before the TMA (fenceOwed)
                        if (!block) return false;
                        __nanosleep(32);
                        ++spins;
                        bool giveUp = false;
                        if ((spins & 0xffu) == 0u) giveUp = spins > (1u << 22) || *(volatile int*)A.abortFlag;
                        if (__any_sync(0xffffffffu, giveUp)) { if (lane == 0) atomicExch(A.abortFlag, 1); return false; }
                    }
                };
                // "done" alternates between two mbarriers by tile parity: the producer looks at done(k-1) only after it has
                // fetched the item after k, and with a single barrier a fast tile k could complete a second phase before the
                // first one was observed (the parity wait would then never succeed).  Tile k+1, the next user of the same
                // barrier, is handed out only after done(k-1) has been consumed, so with two barriers no phase can be skipped.
                uint32_t emptyParity = 0, doneParity0 = 0u, doneParity1 = 0u;
                uint32_t coefPhase = 0;                       // wall tiles handed out so far: buffer = phase % CB, parity = (phase / CB) & 1
                bool stageBusy = false;                       // a tile has been handed to the compute warps and not yet drained
                bool alive = true;
                int seq = 0;
                // Tiles handed to the compute warps whose completion has not been published yet, oldest first (at most two: the
                // one being computed and the one before it).  Publishing = (a) observe the tile's "done" mbarrier, (b) release
                // its generation counter.  (b) waits out the CTA's outstanding stores (1.5-2.5 us), so by default (lateRelease)
                // it is taken off the hand-over chain: when the compute warps have drained tile n the producer consumes
                // done(n-1) -- complete by then, the compute warps arrive on it before they touch tile n -- issues the TMA of
                // tile n+1 at once and only then releases n-1.  (a) always precedes the hand-out of the next user of the same
                // barrier, so no phase of the two alternating barriers can be skipped.
                struct Pend { int slot, gen, seq; bool usedCoef, seen; };
                Pend q0 = { -1, 0, 0, false, false }, q1 = { -1, 0, 0, false, false };
                int nq = 0;
                auto observeOldest = [&](bool block) -> int {      // 1: done observed, 0: not yet (non-blocking only), -1: abort
                    if (q0.seen) return 1;
                    int r = 0;
                    if (lane == 0)
                    {
                        uint64_t* bar = done + (q0.seq & 1);
                        const uint32_t par = (q0.seq & 1) ? doneParity1 : doneParity0;
                        if (block) r = mbarWaitBounded(bar, par, A.abortFlag) ? 1 : -1;
                        else r = mbarTest(bar, par) ? 1 : 0;
                    }
                    r = __shfl_sync(0xffffffffu, r, 0);
                    if (r == 1) { if (q0.seq & 1) doneParity1 ^= 1u; else doneParity0 ^= 1u; q0.seen = true; }
                    return r;
                };
                auto releaseOldest = [&]() {                       // q0.seen must hold
                    // fenceMode 2: the generic-proxy -> async-proxy fence of the consumers' TMA reads is executed HERE, on the
                    // writer side of the causality chain (tile stores -> "done" mbarrier -> this fence -> release -> a
                    // consumer's acquire -> its TMA), once per tile and back to back with the release, which waits out the same
                    // outstanding stores; the consumers then issue their TMA straight after the acquire.
                    if (A.fenceMode == 2) asm volatile("fence.proxy.async;" ::: "memory");
                    if (lane == 0) storeRelease(A.doneGen + q0.slot, q0.gen + 1);
                    q0 = q1; q1.slot = -1; q1.seen = false; q1.usedCoef = false;
                    --nq;
                };
                auto publishOldest = [&](bool block) -> int {
                    if (nq == 0) return 1;
                    const int r = observeOldest(block);
                    if (r == 1) releaseOldest();
                    return r;
                };
                auto publishAll = [&]() -> bool {                  // blocking
                    while (nq > 0) if (publishOldest(true) != 1) return false;
                    return true;
                };
                // Fetching a work item costs two dependent L2 round trips: the counter atomic, then -- all in flight
                // together -- the record loads and the dependency probe.  By default (earlyFetch = 2) the next item is
                // fetched right after the TMA of the current one has been issued, before waiting for the tile in flight, so
                // those round trips never sit between a drained stage and the next TMA; the probe is repeated just before
                // the hand-over if it failed early.  earlyFetch = 0 (PVC_EARLY_FETCH) fetches at the top of the loop.
                struct Item { int valid, s, id, tx, ty, gen, mode, hint, misc; bool anySlow, ready, fenced; } nx;
                auto fetchNext = [&]() {
                    int w = 0;
                    if (lane == 0) w = atomicAdd(A.workCounter, 1);
                    w = __shfl_sync(0xffffffffu, w, 0);
                    nx.valid = w < total;
                    if (!nx.valid) return;
                    // item order (pvc_internal.h::ws2DecodeItem): every dependency of an item precedes it
                    const Ws2Order ord = { A.genChunk, A.srcGroup, A.numGen, A.nsrc, tps, A.numTiles, A.band, L.tiles_x };
                    const Ws2Item wi = ws2DecodeItem(w, ord);
                    const int o = wi.o;
                    nx.s = wi.s;
                    nx.gen = A.gen0 + wi.gen;
                    nx.id = A.tileOrder ? A.tileOrder[o] : o;          // null: row-major order
                    nx.ty = nx.id / L.tiles_x; nx.tx = nx.id - nx.ty * L.tiles_x;
                    // the tile's hand-off record, gathered lane-parallel (loads overlap the dependency probe below)
                    nx.mode = 0; nx.hint = 0; nx.misc = 0;
                    if (lane < NW)
                    {
                        nx.mode = (int)A.mode[(size_t)nx.id * 32 + lane];
                        nx.hint = (A.firstActive && A.hist) ? A.firstActive[((size_t)nx.s * tps + nx.id) * 32 + lane] : 0;
                    }
                    if (lane == 0) nx.misc = A.src[nx.s].cell_r;
                    else if (lane == 1) nx.misc = A.src[nx.s].cell_c;
                    else if (lane == 2) nx.misc = A.src[nx.s].dead;
                    else if (lane >= 4 && lane < 8)
                    {
                        const int t = nx.gen * kTileK + (lane - 4) - (TS ? 1 : 0);          // TS: [pending sample of the previous pass, samples 0..2]
                        nx.misc = __float_as_int((t >= 0 && t < A.T) ? __ldg(A.pulse + t) : 0.f);
                    }
                    nx.ready = (A.earlyFetch == 1) ? false : depsReady(nx.s, nx.tx, nx.ty, nx.gen, false);
                    nx.fenced = false;
                    nx.anySlow = __ballot_sync(0xffffffffu, lane < NW && nx.mode == kGeneral) != 0u;
                };
                if (A.earlyFetch) fetchNext();
                while (alive)
                {
                    if (!A.earlyFetch) fetchNext();
                    if (!nx.valid) break;
                    const Item it = nx;
                    const int s = it.s, id = it.id, tx = it.tx, ty = it.ty, gen = it.gen;
                    const bool anySlow = it.anySlow;
                    bool ready = it.ready;
                    bool fenced = it.fenced;
                    if (!ready && A.earlyFetch) ready = depsReady(s, tx, ty, gen, false);
                    if (PVC_DBG(A) && lane == 0) { atomicAdd(PVC_DBG(A) + 0, 1ull); if (!ready) atomicAdd(PVC_DBG(A) + 1, 1ull); }
                    if (!ready)
                    {
                        if (A.slowPathPoll)
                        {
                            // Poll the dependencies AND our own tiles in flight: a dependency is usually a tile of a lagging CTA
                            // that completes within a microsecond or two, long before our compute warps finish theirs, so the
                            // TMA of this item still goes out a good part of a tile ahead.  (Blocking on our own tile first --
                            // the older path below -- put a whole TMA latency in front of the compute warps on every miss.)
                            // The item may also depend on our own tiles: publish them as soon as they complete.
                            unsigned spins = 0;
                            while (true)
                            {
                                ready = depsReady(s, tx, ty, gen, false);
                                if (ready) break;
                                if (!PUB)
                                {
                                    int r = publishOldest(false);
                                    if (r == 1 && nq > 0) r = publishOldest(false);
                                    if (r < 0) break;
                                }
                                __nanosleep(64);
                                ++spins;
                                bool giveUp = false;
                                if ((spins & 0xffu) == 0u) giveUp = spins > (1u << 22) || *(volatile int*)A.abortFlag;
                                if (__any_sync(0xffffffffu, giveUp)) { if (lane == 0) atomicExch(A.abortFlag, 1); break; }
                            }
                        }
                        else
                        {   // it may depend on the tiles our own compute warps are working on: publish them first, then wait for real
                            if (!PUB && !publishAll()) { alive = false; break; }
                            ready = depsReady(s, tx, ty, gen, true);
                        }
                        if (!ready) { alive = false; break; }
                    }
                    if (stageBusy)
                    {
                        bool ok = true;
                        if (lane == 0) ok = mbarWaitBounded(empty, emptyParity, A.abortFlag);
                        ok = __shfl_sync(0xffffffffu, ok, 0);
                        emptyParity ^= 1u;
                        if (!ok) { alive = false; break; }
                    }
                    if (lane == 0) trace(PVC_DBG(A), seq, 7);
                    // the stage is drained: the compute warps work on the newest pending tile and have arrived on the "done"
                    // barrier of the one before it.  Observe that one now -- this item is the next user of its barrier.
                    if (PUB)
                    {
                        // this item is the next user of done[seq & 1]: the publisher must have consumed tile seq-2's phase (it has,
                        // within nanoseconds of the compute warps' arrival, unless it is still busy with an older release).
                        // Single coefficient buffer: the tile in compute (seq-1) must be finished if both use it.
                        const bool coefClash = CB == 1 && anySlow && q0.usedCoef;
                        if (!waitObserved(coefClash ? seq : seq - 1)) { alive = false; break; }
                    }
                    else
                    {
                        if (nq == 2 && observeOldest(true) != 1) { alive = false; break; }
                        // single coefficient buffer: the tile being computed may still be reading it
                        if (CB == 1 && anySlow && ((nq == 2) ? q1.usedCoef : (nq == 1 && q0.usedCoef))) { if (!publishAll()) { alive = false; break; } }
                    }

                    Meta* m = meta + (seq & 1);
                    const int coefBuf = anySlow ? (int)(coefPhase % CB) : -1;
                    const int coefParity = (int)((coefPhase / CB) & 1u);
                    if (lane < NW) { m->mode[lane] = it.mode; m->hint[lane] = it.hint; }
                    if (lane == 0) m->srcR = it.misc;
                    else if (lane == 1) m->srcC = it.misc;
                    else if (lane == 2) m->srcDead = it.misc;
                    else if (lane >= 4 && lane < 8) m->pulse[lane - 4] = __int_as_float(it.misc);
                    else if (lane == 8) { m->valid = 1; m->s = s; m->tx = tx; m->ty = ty; m->gen = gen; m->coefBuf = coefBuf; m->coefParity = coefParity; }
                    else if (PUB && lane == 9) { pubRing[(seq & 3) * 4] = s * tps + id; pubRing[(seq & 3) * 4 + 1] = gen; pubRing[(seq & 3) * 4 + 2] = s; pubRing[(seq & 3) * 4 + 3] = id; }
                    __syncwarp();
                    // order the dependency acquires before the async-proxy (TMA) reads of the neighbours' cells.  The fence waits
                    // out the CTA's outstanding stores (~1-2 us measured), so on the fast path it has already been executed a
                    // tile ahead (after the early probe AND after the previous tile was published, see below)
                    if (!PUB && !fenced && A.fenceMode != 2) asm volatile("fence.proxy.async;" ::: "memory");
                    if (lane == 0)
                    {
                        if (anySlow)
                        {
                            uint64_t* bar = fullCoef + coefBuf;
                            float4* dst = sCoef + (size_t)coefBuf * SM::NP * TR * 32;
                            mbarExpectTx(bar, (SM::NP == 3) ? 3u * SM::kPlaneBytes : 2u * SM::kPlaneBytes + SM::kMaskBytes);
                            tmaLoad2d(dst, &maps.coef[0], tx * kValidCols, ty * L.valid_rows, bar);
                            tmaLoad2d(dst + (size_t)TR * 32, &maps.coef[1], tx * kValidCols, ty * L.valid_rows, bar);
                            if (SM::NP == 3) tmaLoad2d(dst + (size_t)2 * TR * 32, &maps.coef[2], tx * kValidCols, ty * L.valid_rows, bar);
                            else bulkLoad1d(sMask + (size_t)coefBuf * NW * 32, A.bpMask + (size_t)id * NW * 32, SM::kMaskBytes, bar);
                        }
                        const CUtensorMap* mp = (gen & 1) ? &maps.state[3] : &maps.state[0];
                        mbarExpectTx(full, 3u * SM::kPlaneBytes);
                        tmaLoad3d(stage, mp + 0, tx * kValidCols, ty * L.valid_rows, s, full);
                        tmaLoad3d(stage + (size_t)TR * kTileCols, mp + 1, tx * kValidCols, ty * L.valid_rows, s, full);
                        tmaLoad3d(stage + (size_t)2 * TR * kTileCols, mp + 2, tx * kValidCols, ty * L.valid_rows, s, full);
                    }
                    __syncwarp();
                    if (lane == 0) trace(PVC_DBG(A), seq, 0);
                    if (anySlow) ++coefPhase;
                    ++seq;
                    stageBusy = true;
                    // the tile before the one in compute has been observed above: release it now, after the TMA has gone out
                    if (PUB) q0.usedCoef = anySlow;               // only "the tile in compute reads the coefficient buffer" is tracked
                    else
                    {
                        if (nq == 2) releaseOldest();
                        const Pend me = { s * tps + id, gen, seq - 1, anySlow, false };
                        if (nq == 0) q0 = me; else q1 = me;
                        ++nq;
                    }
                    if (A.earlyFetch) fetchNext();                // the next item's fetch overlaps the tile in flight
                    if (lane == 0) trace(PVC_DBG(A), seq - 1, 1);
                    // fenceMode 1: the proxy fence right after the early probe, before waiting for the tile in compute
                    if (!PUB && A.fenceMode == 1 && A.earlyFetch && nx.valid && nx.ready) { asm volatile("fence.proxy.async;" ::: "memory"); nx.fenced = true; }
                    // older order (PVC_LATE_RELEASE=0): the tile handed over before this one is published here, blocking
                    if (!PUB && !A.lateRelease && nq == 2 && publishOldest(true) != 1) { alive = false; break; }
                    if (!PUB && A.fenceMode == 0 && A.earlyFetch && nx.valid && nx.ready) { asm volatile("fence.proxy.async;" ::: "memory"); nx.fenced = true; }
                    if (lane == 0) trace(PVC_DBG(A), seq - 1, 2);
                }
                // drain: publish the last tile, then tell the compute warps to stop
                if (PUB) { if (lane == 0) *pubTotal = seq; }
                else if (alive) alive = publishAll();
                if (stageBusy && alive)
                {
                    bool ok = true;
                    if (lane == 0) ok = mbarWaitBounded(empty, emptyParity, A.abortFlag);
                    (void)ok;
                }
                if (lane == 0) { meta[seq & 1].valid = 0; mbarArrive(full); }      // wake the compute warps with "no more work"
                return;
            }

            // ================= compute warps =================
            uint32_t fullParity = 0;
            int seq = 0;
            long long tStart = PVC_DBG(A) ? clock64() : 0;
            const bool owner = TS && wp >= 1 && wp <= NW - 2;
            const bool storer = TS && threadIdx.x == 0;
            float* const outBuf = reinterpret_cast<float*>(smemRaw + SM::offOut);
            uint64_t* deferredDone = nullptr;
            const uint64_t histPolicy = TS ? evictFirstPolicy() : 0ull;
            while (true)
            {
                long long tw0 = 0;
                if (PVC_DBG(A)) tw0 = clock64();
                if (threadIdx.x == 32 * 7) trace(PVC_DBG(A), seq, 3);
                if (!mbarWaitBounded(full, fullParity, A.abortFlag)) break;
                if (PVC_DBG(A) && threadIdx.x == 0) { const long long t1 = clock64(); atomicAdd(PVC_DBG(A) + 2, (unsigned long long)(t1 - tw0)); atomicAdd(PVC_DBG(A) + 3, (unsigned long long)(t1 - tStart)); tStart = t1; }
                fullParity ^= 1u;
                if (threadIdx.x == 32 * 7) trace(PVC_DBG(A), seq, 4);
                const Meta* m = meta + (seq & 1);
                ++seq;
                if (m->valid == 0) break;
                const int s = m->s, tx = m->tx, ty = m->ty, gen = m->gen;
                const int mode = m->mode[wp];
                const int hintKnown = m->hint[wp];
                const int coefBuf = m->coefBuf;
                const uint32_t coefParity = (uint32_t)m->coefParity;

                float p[R][4], vx[R][4], vy[R][4];
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    const int row = wp * R + j;
                    const float4 a = *reinterpret_cast<const float4*>(stage + ((size_t)(0 * TR + row)) * kTileCols + lane * 4);
                    const float4 b = *reinterpret_cast<const float4*>(stage + ((size_t)(1 * TR + row)) * kTileCols + lane * 4);
                    const float4 c = *reinterpret_cast<const float4*>(stage + ((size_t)(2 * TR + row)) * kTileCols + lane * 4);
                    p[j][0] = a.x; p[j][1] = a.y; p[j][2] = a.z; p[j][3] = a.w;
                    vx[j][0] = b.x; vx[j][1] = b.y; vx[j][2] = b.z; vx[j][3] = b.w;
                    vy[j][0] = c.x; vy[j][1] = c.y; vy[j][2] = c.z; vy[j][3] = c.w;
                }
                // first vertical halo row of the pass: vx of the first row of the warp below, straight from the stage
                float4 vxBelow = make_float4(0.f, 0.f, 0.f, 0.f);
                if (wp + 1 < NW) vxBelow = *reinterpret_cast<const float4*>(stage + ((size_t)(1 * TR + (wp + 1) * R)) * kTileCols + lane * 4);
                __syncwarp();
                if (lane == 0) mbarArrive(empty);                  // this warp has drained the stage and read the record's scalars

                // ---- per-thread tile constants
                const int rBase = ty * L.valid_rows - kTileK + wp * R;
                const int cBase = tx * kValidCols - kGuardCols + lane * 4;
                const size_t src0 = (size_t)s * L.plane + (size_t)(rBase + kGuardRows) * L.pitch + (cBase + kGuardCols);
                uint32_t colOut = 0u, colPad = 0u, colLeft = 0u;
                #pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const int cc = cBase + k;
                    if (cc < 0 || cc > L.gy) colOut |= 1u << k;
                    if (cc == L.gy) colPad |= 1u << k;
                    if (cc == 0) colLeft |= 1u << k;
                }
                // owned (stored) rows of this thread: not halo, inside the alloc grid; none for the halo lanes 0 and 31
                // and for columns past the grid
                int jLo = kTileK - wp * R, jHi = NW * R - kTileK - wp * R;
                jLo = max(jLo, 0);
                jHi = min(min(jHi, R), L.rows - rBase);
                if (lane == 0 || lane == 31 || cBase >= L.cols) jHi = 0;
                uint32_t ownRows = 0u;
                #pragma unroll
                for (int j = 0; j < R; ++j) if (j >= jLo && j < jHi) ownRows |= 1u << j;

                const int t0 = gen * kTileK;
                const int nsteps = min(kTileK, A.T - t0);
                TileCtx X;
                X.hist = nullptr;
                if (A.hist)
                    X.hist = A.hist + (size_t)s * L.hist_source + (ptrdiff_t)rBase * (ptrdiff_t)L.hist_row
                           + ((ptrdiff_t)(cBase >> 7) * L.T + t0) * kHistChunkDefault + (cBase & 127);
                X.histRow = L.hist_row;
                X.ownRows = ownRows;
                {
                    const int sj = m->srcR - rBase, sk = m->srcC - cBase;
                    const bool hasSrc = (sj >= 0) && (sj < R) && (sk >= 0) && (sk < 4);
                    X.sj = hasSrc ? sj : -1; X.sk = sk;
                    // A listener on the padding row / column or in a wall cell (b == 0): the reference zeroes the injected sample in the next
                    // pressure sub-step before anything reads it (FDTD.cpp:125-141, :234), so the cell records 0 for ever and
                    // only the very last sample survives in the final state.  The edge path never recomputes such cells, so
                    // the injection is dropped instead -- except for that last sample.
                    X.onlyLast = false;
                    if (hasSrc && (m->srcR >= L.gx || m->srcC >= L.gy || m->srcDead))
                    {
                        if (TS || t0 + nsteps < A.T || !A.finalPass) X.sj = -1;
                        else X.onlyLast = true;
                    }
                }
                X.pulse = m->pulse;
                const bool hints = (A.firstActive != nullptr) && (A.hist != nullptr);
                X.track = hints && hintKnown > gen;
                uint32_t activity = 0u;
                X.out = outBuf; X.outRow = owner ? (wp - 1) * R : -1; X.storer = storer;
                X.maps = &maps; X.tx = tx; X.ty = ty; X.s = s; X.t0 = t0; X.gen = gen;
                X.histOn = A.hist != nullptr;
                X.deferredDone = deferredDone; X.policy = histPolicy; X.tsDebug = A.tsDebug;
                deferredDone = nullptr;

                if (mode == kFast)
                {
                    const Stepper<NW, R, kFast> S{ L, A.courant, lane, wp, rBase, colOut, colPad, colLeft, 0u, nullptr, nullptr, nullptr };
                    stepLoop<NW, R, kFast, TS>(S, X, nsteps, p, vx, vy, vxBelow, sVxTop, sPBot, activity);
                }
                else if (mode == kEdge)
                {
                    const Stepper<NW, R, kEdge> S{ L, A.courant, lane, wp, rBase, colOut, colPad, colLeft, 0u, nullptr, nullptr, nullptr };
                    stepLoop<NW, R, kEdge, TS>(S, X, nsteps, p, vx, vy, vxBelow, sVxTop, sPBot, activity);
                }
                else
                {
                    // the coefficient buffer of this tile (usually landed long ago: it was requested one tile ahead)
                    bool ok = mbarWaitBounded(fullCoef + coefBuf, coefParity, A.abortFlag);
                    (void)ok;
                    const float4* cg = sCoef + (size_t)coefBuf * SM::NP * TR * 32 + (size_t)(wp * R) * 32 + lane;
                    if (SM::NP == 3)
                    {
                        const Stepper<NW, R, kGeneral, true> S{ L, A.courant, lane, wp, rBase, colOut, colPad, colLeft, 0u, cg, cg + (size_t)TR * 32, cg + (size_t)2 * TR * 32 };
                        stepLoop<NW, R, kGeneral, TS, true>(S, X, nsteps, p, vx, vy, vxBelow, sVxTop, sPBot, activity);
                    }
                    else
                    {
                        const uint32_t bpBits = sMask[(size_t)coefBuf * NW * 32 + wp * 32 + lane];
                        const Stepper<NW, R, kGeneral> S{ L, A.courant, lane, wp, rBase, colOut, colPad, colLeft, bpBits, cg, cg + (size_t)TR * 32, nullptr };
                        stepLoop<NW, R, kGeneral, TS>(S, X, nsteps, p, vx, vy, vxBelow, sVxTop, sPBot, activity);
                    }
                }

                if (threadIdx.x == 32 * 7) trace(PVC_DBG(A), seq - 1, 5);
                // ---- store the owned cells of the new state (generation g reads buffer g & 1, writes the other)
                if (TS)
                {
                    // p (= the last sample) and vx left with the last step's group; vy goes through the plane the storer freed
                    // before the step's closing barrier
                    constexpr int kPlane = (NW - 2) * R * kValidCols;
                    const int bLast = (nsteps == kTileK) ? 0 : nsteps - 1;
                    float* planeVy = outBuf + (bLast + 2) % 3 * kPlane;
                    if (owner) stageRows<R>(planeVy, X.outRow, lane, vy);
                    computeBarrier<NW>();
                    if (storer)
                    {
                        if (!(A.tsDebug & 2)) storeStatePlane(X, L, planeVy, 2);
                        bulkCommit();
                        if (t0 + nsteps == A.T && A.T > 0)
                        {
                            // the very last pass has no successor to inject its last pulse sample: if this tile owns the pulse
                            // cell, add it to the stored state here (after the stores have completed) so that the final state
                            // equals the other kernels'
                            const int sr = m->srcR - (ty * L.valid_rows), sc = m->srcC - tx * kValidCols;
                            if (sr >= 0 && sr < (NW - 2) * R && sc >= 0 && sc < kValidCols)
                            {
                                bulkWait<0>();
                                float* q = ((gen & 1) ? A.p0 : A.p1) + (size_t)s * L.plane + cellIndex(L, m->srcR, m->srcC);
                                *q = __fadd_rn(*q, __ldg(A.pulse + (A.T - 1)));
                            }
                        }
                    }
                }
                else if (SO)
                {
                    // the TMA stores of the previous tile must have read the planes (they have, a whole tile ago)
                    if (seq > 1) { bool ok = mbarWaitBounded(outFree, (uint32_t)(seq & 1), A.abortFlag); (void)ok; }
                    if (wp >= 1 && wp <= NW - 2 && lane >= 1 && lane <= 30)
                    {
                        constexpr int kPlane = (NW - 2) * R * kValidCols;
                        float* o = outBuf + ((wp - 1) * R) * kValidCols + (lane - 1) * 4;
                        #pragma unroll
                        for (int j = 0; j < R; ++j)
                        {
                            *reinterpret_cast<float4*>(o + j * kValidCols) = make_float4(p[j][0], p[j][1], p[j][2], p[j][3]);
                            *reinterpret_cast<float4*>(o + kPlane + j * kValidCols) = make_float4(vx[j][0], vx[j][1], vx[j][2], vx[j][3]);
                            *reinterpret_cast<float4*>(o + 2 * kPlane + j * kValidCols) = make_float4(vy[j][0], vy[j][1], vy[j][2], vy[j][3]);
                        }
                    }
                    fenceAsyncShared();          // generic-proxy writes above -> async-proxy reads of the publisher's TMA stores
                }
                else
                {
                    float* gp = ((gen & 1) ? A.p0 : A.p1) + src0;
                    float* gx = ((gen & 1) ? A.vx0 : A.vx1) + src0;
                    float* gy = ((gen & 1) ? A.vy0 : A.vy1) + src0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        if ((ownRows >> j) & 1u)
                        {
                            *reinterpret_cast<float4*>(gp + (size_t)j * L.pitch) = make_float4(p[j][0], p[j][1], p[j][2], p[j][3]);
                            *reinterpret_cast<float4*>(gx + (size_t)j * L.pitch) = make_float4(vx[j][0], vx[j][1], vx[j][2], vx[j][3]);
                            *reinterpret_cast<float4*>(gy + (size_t)j * L.pitch) = make_float4(vy[j][0], vy[j][1], vy[j][2], vy[j][3]);
                        }
                    }
                }
                if (X.track)
                {
                    // activity hint for the analyzer: first generation in which this warp's block recorded anything but
                    // zeros (conservative: halo rows and -0 count as activity)
                    const bool hot = ((activity & 0x7fffffffu) != 0u) && lane >= 1 && lane <= 30;
                    const unsigned any = __ballot_sync(0xffffffffu, hot);
                    if (lane == 0 && any)
                        atomicMin(A.firstActive + ((size_t)s * tps + (size_t)ty * L.tiles_x + tx) * 32 + wp, gen);
                }
                __syncwarp();
                if (threadIdx.x == 32 * 7) trace(PVC_DBG(A), seq - 1, 6);
                uint64_t* doneBar = done + ((seq - 1) & 1);
                if (TS)
                {
                    if (storer)
                    {
                        // The tile counts as stored once the bulk stores have COMPLETED.  If the next tile has already landed,
                        // do not wait here: arrive after its first step (stepLoop), when only the newest group may still be in
                        // flight.  If it has not, the producer may be waiting for exactly this arrival before it can hand over
                        // more work (dependency slow path), so complete now.
                        if (!(A.tsDebug & 4) && mbarTest(full, fullParity)) deferredDone = doneBar;
                        else { bulkWait<0>(); mbarArrive(doneBar); }
                    }
                }
                else if (lane == 0) mbarArrive(doneBar);             // every store of this warp for this tile has been issued
            }
            if (storer)
            {
                bulkWait<0>();
                if (deferredDone) mbarArrive(deferredDone);
            }
        }

        // bit j*4+k of word [tile][warp][lane]: cell (j, k) of that thread's block is an interior air cell (reference b = 1)
        template <int NW, int R>
        __global__ void bpMaskKernel(const Layout L, const float* __restrict__ w, uint32_t* __restrict__ mask)
        {
            const int lane = threadIdx.x & 31;
            const int wp = threadIdx.x >> 5;
            const int tx = blockIdx.x, ty = blockIdx.y;
            const int rBase = ty * L.valid_rows - kTileK + wp * R;
            const int cBase = tx * kValidCols - kGuardCols + lane * 4;
            uint32_t bits = 0u;
            for (int j = 0; j < R; ++j)
                for (int k = 0; k < 4; ++k)
                {
                    const int r = rBase + j, c = cBase + k;
                    const bool interior = (r >= 0) && (r < L.gx) && (c >= 0) && (c < L.gy);
                    if (interior && __float_as_uint(w[cellIndex(L, r, c)]) == kAirBits) bits |= 1u << (j * 4 + k);
                }
            mask[(((size_t)ty * L.tiles_x + tx) * NW + wp) * 32 + lane] = bits;
        }

        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

        // 2-D tensor maps {pitch, rows_alloc} of the gx / gy coefficient planes, box 128 x tileRows
        static int buildCoefMaps(pvc_solver* s, CUtensorMap* out, bool linear)
        {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            { cudaGetLastError(); setError("ws2 step kernel: cuTensorMapEncodeTiled unavailable"); return PVC_ERR_CUDA; }
            const Layout& L = s->L;
            for (int f = 0; f < 3; ++f)
            {
                const cuuint64_t dims[2] = { (cuuint64_t)L.pitch, (cuuint64_t)L.rows_alloc };
                const cuuint64_t strides[1] = { (cuuint64_t)L.pitch * sizeof(float) };
                const cuuint32_t box[2] = { (cuuint32_t)kTileCols, (cuuint32_t)L.tile_rows };
                const cuuint32_t estr[2] = { 1u, 1u };
                // staged plane order: select form gx, gy, bp (s->coef = bp, gx, gy); linear form sX, sY, cP (s->lin = cP, sX, sY)
                float* plane = linear ? s->lin[(1 + f) % 3] : s->coef[(1 + f) % 3];
                const CUresult r = ((EncodeFn)fn)(&out[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { setError("ws2 step kernel: coefficient tensor map %d failed (%d)", f, (int)r); return PVC_ERR_CUDA; }
            }
            return PVC_OK;
        }

        // TS variants: store maps of the six state planes (dims clipped to the alloc grid, so the hardware drops what the
        // per-thread predicates of the plain variant drop) and the 5-D history map
        static int buildStoreMaps(pvc_solver* s, float* hist, int boxRows, CUtensorMap* store, CUtensorMap* histMap)
        {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
            { cudaGetLastError(); setError("ws2 step kernel: cuTensorMapEncodeTiled unavailable"); return PVC_ERR_CUDA; }
            const Layout& L = s->L;
            for (int b = 0; b < 2; ++b)
                for (int f = 0; f < 3; ++f)
                {
                    const cuuint64_t dims[3] = { (cuuint64_t)(kGuardCols + L.cols), (cuuint64_t)(kGuardRows + L.rows), (cuuint64_t)s->cfg.max_sources };
                    const cuuint64_t strides[2] = { (cuuint64_t)L.pitch * sizeof(float), (cuuint64_t)L.plane * sizeof(float) };
                    const cuuint32_t box[3] = { (cuuint32_t)kValidCols, (cuuint32_t)boxRows, 1u };
                    const cuuint32_t estr[3] = { 1u, 1u, 1u };
                    const CUresult r = ((EncodeFn)fn)(&store[b * 3 + f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->state[b][f], dims, strides, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { setError("ws2 step kernel: state store tensor map failed (%d)", (int)r); return PVC_ERR_CUDA; }
                }
            memset(histMap, 0, sizeof(*histMap));
            if (hist)
            {
                const cuuint64_t dims[5] = { (cuuint64_t)L.hist_chunk, (cuuint64_t)L.T, (cuuint64_t)L.hist_chunks, (cuuint64_t)L.rows, (cuuint64_t)s->cfg.max_sources };
                const cuuint64_t strides[4] = { (cuuint64_t)L.hist_chunk * sizeof(float), (cuuint64_t)L.T * L.hist_chunk * sizeof(float),
                                                (cuuint64_t)L.hist_row * sizeof(float), (cuuint64_t)L.hist_source * sizeof(float) };
                const cuuint32_t box[5] = { (cuuint32_t)kValidCols, 1u, 1u, (cuuint32_t)boxRows, 1u };
                const cuuint32_t estr[5] = { 1u, 1u, 1u, 1u, 1u };
                const CUresult r = ((EncodeFn)fn)(histMap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, hist, dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { setError("ws2 step kernel: history tensor map failed (%d)", (int)r); return PVC_ERR_CUDA; }
            }
            return PVC_OK;
        }

        template <int NW, int R, int CB, bool TS = false, bool PUB = false, bool SO = false>
        static int launch(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
        {
            using SM = Smem<NW, R, CB, TS, SO>;
            const Layout& L = s->L;
            if (!s->tmaReady || s->tmaTileRows != NW * R) { setError("ws2 step kernel: tensor maps not built for %d-row tiles", NW * R); return PVC_ERR_INVALID; }
            // a chunk of a streamed solve (pvc_create_streamed) is a launch of its own that starts from the state planes as they are:
            // samples t0 .. t1 - 1 of the response, recorded as samples 0 .. t1 - t0 - 1 of the (chunk-sized) history
            if ((t0 != 0 && !s->chunkT) || s->cur != 0) { setError("ws2 step kernel: must start at step 0"); return PVC_ERR_INVALID; }
            if (s->chunkT && (TS || SO)) { setError("ws2 step kernel: variant not usable for a streamed solve"); return PVC_ERR_INVALID; }
            if (!s->bpMask) { setError("ws2 step kernel: descriptor buffer missing"); return PVC_ERR_INVALID; }
            if (L.hist_chunk != (TS ? kValidCols : kHistChunkDefault)) { setError("ws2 step kernel: history strip width %d does not match the variant", L.hist_chunk); return PVC_ERR_INVALID; }
            const size_t smem = SM::total;
            static bool configured[64] = {};
            if (!configured[s->device & 63])
            {
                cudaError_t e = cudaFuncSetAttribute(stepKernel<NW, R, CB, TS, PUB, SO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { setError("ws2 step kernel smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
                configured[s->device & 63] = true;
            }
            Maps maps;
            memcpy(maps.state, s->tensorMaps, sizeof(maps.state));
            if (SM::NP == 3 && !s->lin[0]) { setError("ws2 step kernel: linear coefficient planes missing"); return PVC_ERR_INVALID; }
            int rc = buildCoefMaps(s, maps.coef, SM::NP == 3);
            if (rc) return rc;
            memset(maps.store, 0, sizeof(maps.store)); memset(&maps.hist, 0, sizeof(maps.hist));
            const int numTiles = L.tiles_x * L.tiles_y * nsrc;
            const int grid = numTiles < s->numSMs ? numTiles : s->numSMs;          // all CTAs must be co-resident (1 CTA per SM)
            const int gens = (t1 - t0 + kTileK - 1) / kTileK;
            const int perLaunch = 256;                                             // generations per launch (bounds kernel time)
            cudaMemsetAsync(s->doneGen, 0, sizeof(int) * (size_t)numTiles, s->stream);
            {   // slot 0 is the abort flag: a streamed solve clears it with its first chunk only, so that a time-out in any chunk is still there
                // when the host looks (after the last one)
                const int keep = s->abortSticky ? 1 : 0;
                cudaMemsetAsync(s->tileCounters + keep, 0, sizeof(int) * (size_t)((gens + perLaunch - 1) / perLaunch + 1 - keep), s->stream);
            }
            Args A;
            A.p0 = s->state[0][0]; A.vx0 = s->state[0][1]; A.vy0 = s->state[0][2];
            A.p1 = s->state[1][0]; A.vx1 = s->state[1][1]; A.vy1 = s->state[1][2];
            A.hist = hist;
#ifdef PVC_TUNING
            { static const char* dbg = getenv("PVC_DEBUG_NOHIST"); if (dbg) A.hist = nullptr; }      // debug: memory-floor probe (results invalid)
#endif
            if (TS || SO) { rc = buildStoreMaps(s, TS ? A.hist : nullptr, (NW - 2) * R, maps.store, &maps.hist); if (rc) return rc; }
            A.mode = s->slowMask; A.bpMask = s->bpMask; A.tileOrder = s->tileOrderNatural ? nullptr : s->tileOrder;
            A.firstActive = (s->chunkT && !s->hintsValid) ? nullptr : s->firstActive;   // streamed solve: hints per chunk, forward sweep only
            A.src = s->src; A.pulse = s->pulse + t0; A.finalPass = s->chunkT ? s->finalPass : 1;
            A.doneGen = s->doneGen; A.abortFlag = s->tileCounters;                  // slot 0 of the pool is the abort flag
            A.tilesPerSource = L.tiles_x * L.tiles_y; A.nsrc = nsrc; A.numTiles = numTiles;
            A.T = t1 - t0; A.courant = s->cfg.courant;
            {
                // sources per group: as many as keep both ping-pong copies of the group's state (2 x 12 B per cell) inside
                // ~85 MB of the 126 MB L2, groups balanced; the history stream is written evict-first and does not compete
                const double perSource = 24.0 * (double)L.plane;
                int maxSg = (int)(85.0e6 / perSource);
                A.band = 0;
                A.genChunk = 16;
                // (Bands of tile rows for grids whose state exceeds the L2 even for one source -- pvc_internal.h::Ws2Order -- were built and
                // measured on config 4's grid: a band small enough to stay in the L2 (15 tile rows at 2048^2) leaves an item only 270 items
                // away from its dependencies, fewer than the CTAs in flight, and the polling costs 14 %; a band that keeps the distance
                // (22 rows) no longer fits and changes nothing.  profiles/r02_l2_bands.txt.  A.band stays 0 outside tuning builds.)
                if (maxSg < 1) maxSg = 1;
                const int groups = (nsrc + maxSg - 1) / maxSg;
                A.srcGroup = (nsrc + groups - 1) / groups;
                // protocol knobs, fixed in the release build (the measured optimum, profiles/r01_variants.txt); a tuning build
                // (make EXTRA=-DPVC_TUNING) reads them from the environment
                A.earlyFetch = 2; A.slowPathPoll = 1; A.lateRelease = 0; A.fenceMode = 0; A.tsDebug = 0;
                A.debug = nullptr;
#ifdef PVC_TUNING
                static const char* eg = getenv("PVC_GROUP_SRC"); static const char* ec = getenv("PVC_GROUP_GENS");
                { static const char* ef = getenv("PVC_EARLY_FETCH"); if (ef) A.earlyFetch = atoi(ef); }
                { static const char* sp = getenv("PVC_SLOW_POLL"); if (sp) A.slowPathPoll = atoi(sp); }
                { static const char* lr = getenv("PVC_LATE_RELEASE"); if (lr) A.lateRelease = atoi(lr); }
                { static const char* fm = getenv("PVC_FENCE_MODE"); if (fm) A.fenceMode = atoi(fm); }
                { static const char* td = getenv("PVC_TS_DEBUG"); if (td) A.tsDebug = atoi(td); }
                if (eg && atoi(eg) > 0) A.srcGroup = atoi(eg);
                if (ec && atoi(ec) > 0) A.genChunk = atoi(ec);
                { static const char* eb = getenv("PVC_BAND_ROWS"); if (eb) A.band = (atoi(eb) > 0 && s->tileOrderNatural) ? atoi(eb) : 0; }
#endif
#ifdef PVC_WS2_TRACE
                {
                    static const char* dbg = getenv("PVC_DEBUG_COUNTERS");
                    static unsigned long long* counters = nullptr;
                    if (dbg)
                    {
                        const size_t nWords = 4 + (size_t)kTraceCtas * kTraceTiles * kTraceSlots;
                        if (!counters) { cudaMalloc(&counters, nWords * sizeof(unsigned long long)); }
                        else
                        {
                            std::vector<unsigned long long> hv(nWords);
                            cudaMemcpy(hv.data(), counters, nWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                            const unsigned long long* h = hv.data();
                            {
                                // per-tile phases (ns), averaged over the traced tiles: k = tile, producer slots 0..2, compute (warp 7) slots 3..6
                                double sum[8] = {}; long cnt = 0;
                                for (int c = 0; c < kTraceCtas; ++c)
                                    for (int k = 1; k + 1 < kTraceTiles; ++k)
                                    {
                                        const unsigned long long* a = h + 4 + ((size_t)c * kTraceTiles + k) * kTraceSlots;
                                        const unsigned long long* prev = a - kTraceSlots;
                                        if (!a[0] || !a[4] || !a[6] || !prev[6] || !a[3] || !a[5] || !a[2]) continue;
                                        sum[0] += (double)a[4] - (double)a[0];        // TMA issue -> compute sees full
                                        sum[1] += (double)a[4] - (double)a[3];        // compute wait on full
                                        sum[2] += (double)a[5] - (double)a[4];        // drain + 4 steps
                                        sum[3] += (double)a[6] - (double)a[5];        // state stores issued
                                        sum[4] += (double)a[0] - (double)prev[4];     // previous tile ready -> this TMA issued
                                        sum[5] += (double)a[1] - (double)a[0];        // fetch of the next item
                                        sum[6] += (double)a[2] - (double)a[1];        // wait for + publish the previous tile
                                        sum[7] += (double)a[6] - (double)prev[6];     // tile period
                                        ++cnt;
                                    }
                                if (cnt) fprintf(stderr, "[ws2 trace] n=%ld  issue->ready %.0f  wait-on-full %.0f  drain+steps %.0f  stores %.0f | prevReady->issue %.0f  fetchNext %.0f  publishPrev %.0f | period %.0f ns\n",
                                                 cnt, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, sum[4] / cnt, sum[5] / cnt, sum[6] / cnt, sum[7] / cnt);
                            }
                            fprintf(stderr, "[ws2 counters] tiles %llu  slow-path hand-overs %llu (%.2f%%)  compute-warp wait on full: %.1f%% of cycles\n",
                                    h[0], h[1], h[0] ? 100.0 * h[1] / h[0] : 0.0, h[3] ? 100.0 * h[2] / h[3] : 0.0);
                        }
                        cudaMemset(counters, 0, nWords * sizeof(unsigned long long));
                        A.debug = counters;
                    }
                }
#endif
                if (A.srcGroup > nsrc) A.srcGroup = nsrc;
            }
#ifdef PVC_TUNING
            {
                // experiment (PVC_L2_PERSIST=<MB>): pin the ping-pong state of the batch in the L2 with an access-policy window
                static const char* lp = getenv("PVC_L2_PERSIST");
                if (lp && atoi(lp) > 0)
                {
                    static bool limitSet = false;
                    if (!limitSet) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)atoi(lp) << 20); limitSet = true; }
                    int maxWin = 0; cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, s->device);
                    cudaStreamAttrValue av; memset(&av, 0, sizeof(av));
                    size_t bytes = sizeof(float) * 6 * (size_t)s->cfg.max_sources * L.plane;
                    if (maxWin > 0 && bytes > (size_t)maxWin) bytes = (size_t)maxWin;
                    av.accessPolicyWindow.base_ptr = s->stateBlock;
                    av.accessPolicyWindow.num_bytes = bytes;
                    av.accessPolicyWindow.hitRatio = (float)(((double)((size_t)atoi(lp) << 20)) / (double)bytes > 1.0 ? 1.0 : ((double)((size_t)atoi(lp) << 20)) / (double)bytes);
                    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    const cudaError_t e = cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av);
                    static bool said = false;
                    if (!said) { fprintf(stderr, "[ws2] L2 window: %zu MB of state, set-aside %d MB, hit ratio %.2f, max window %d MB: %s\n", bytes >> 20, atoi(lp), av.accessPolicyWindow.hitRatio, maxWin >> 20, cudaGetErrorString(e)); said = true; }
                }
            }
#endif
            int k = 1;
            for (int g0 = 0; g0 < gens; g0 += perLaunch, ++k)
            {
                A.gen0 = g0; A.numGen = (gens - g0 < perLaunch) ? (gens - g0) : perLaunch;
                A.workCounter = s->tileCounters + k;
                stepKernel<NW, R, CB, TS, PUB, SO><<<grid, (NW + 1 + (PUB ? 1 : 0)) * 32, smem, s->stream>>>(L, A, maps);
                *launches += 1;
            }
            s->cur = gens & 1;
            s->checkAbort = 1;
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { setError("ws2 step kernel launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            return PVC_OK;
        }

        template <int NW, int R>
        static int buildMask(pvc_solver* s)
        {
            const Layout& L = s->L;
            bpMaskKernel<NW, R><<<dim3(L.tiles_x, L.tiles_y), NW * 32, 0, s->stream>>>(L, s->w, s->bpMask);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { setError("bp mask launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            return PVC_OK;
        }
    }

    // ---- entry points used by pvc_step_fused.cu's variant table ----
    int launchWs2Steps(pvc_solver* s, int variant, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        switch (variant)
        {
            case 47: return ws2::launch<14, 4, 1, false, true>(s, nsrc, t0, t1, hist, launches);
            case 50: return ws2::launch<8, 4, 1, false, true>(s, nsrc, t0, t1, hist, launches);
#ifdef PVC_ALL_VARIANTS
            case 39: return ws2::launch<14, 4, 2>(s, nsrc, t0, t1, hist, launches);
            case 40: return ws2::launch<15, 4, 1>(s, nsrc, t0, t1, hist, launches);
            case 41: return ws2::launch<30, 2, 1>(s, nsrc, t0, t1, hist, launches);
            case 42: return ws2::launch<20, 3, 1>(s, nsrc, t0, t1, hist, launches);
            case 43: return ws2::launch<14, 4, 1, true>(s, nsrc, t0, t1, hist, launches);
            case 44: return ws2::launch<10, 8, 1>(s, nsrc, t0, t1, hist, launches);
            case 45: return ws2::launch<11, 6, 1>(s, nsrc, t0, t1, hist, launches);
            case 46: return ws2::launch<12, 6, 1>(s, nsrc, t0, t1, hist, launches);
            case 48: return ws2::launch<15, 4, 1, false, true>(s, nsrc, t0, t1, hist, launches);
            case 49: return ws2::launch<14, 4, 1, false, true, true>(s, nsrc, t0, t1, hist, launches);
            case 51: return ws2::launch<10, 4, 1, false, true>(s, nsrc, t0, t1, hist, launches);
            case 52: return ws2::launch<12, 4, 1, false, true>(s, nsrc, t0, t1, hist, launches);
            case 53: return ws2::launch<12, 5, 1, false, true>(s, nsrc, t0, t1, hist, launches);
#endif
            default: setError("ws2 step kernel: variant %d is not compiled into this build", variant); return PVC_ERR_INVALID;
        }
    }
    int rebuildWs2Descriptors(pvc_solver* s, int variant)
    {
        switch (variant)
        {
            case 47: return ws2::buildMask<14, 4>(s);
            case 50: return ws2::buildMask<8, 4>(s);
#ifdef PVC_ALL_VARIANTS
            case 39: return ws2::buildMask<14, 4>(s);
            case 40: return ws2::buildMask<15, 4>(s);
            case 41: return ws2::buildMask<30, 2>(s);
            case 42: return ws2::buildMask<20, 3>(s);
            case 43: return ws2::buildMask<14, 4>(s);
            case 44: return ws2::buildMask<10, 8>(s);
            case 45: return ws2::buildMask<11, 6>(s);
            case 46: return ws2::buildMask<12, 6>(s);
            case 48: return ws2::buildMask<15, 4>(s);
            case 49: return ws2::buildMask<14, 4>(s);
            case 51: return ws2::buildMask<10, 4>(s);
            case 52: return ws2::buildMask<12, 4>(s);
            case 53: return ws2::buildMask<12, 5>(s);
#endif
            default: return PVC_OK;
        }
    }
}

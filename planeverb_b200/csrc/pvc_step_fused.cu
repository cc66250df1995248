// pvc_step_fused.cu -- the hot kernel: K = 4 full FDTD time steps per launch, state in registers.
//
// What it replaces: the per-time-step body of Grid::GenerateResponseCPU
// (ProjectPlaneverb/src/FDTD/FDTD.cpp:122-235): pressure sweep, vx sweep, vy sweep, grid-edge
// absorbing overrides, IR record, pulse injection -- three full-grid sweeps plus a 16-byte scatter per
// cell per step on the CPU.  Here all of it is one launch per FOUR steps:
//
//   * A CTA owns a tile of (NW*R) rows x 128 columns of the alloc grid INCLUDING a 4-cell halo on
//     every side; warp w owns R consecutive rows, lane l owns 4 consecutive columns (one float4), so a
//     thread keeps p, vx, vy of R x 4 cells in registers for the whole launch (R*12 registers).
//   * Loads/stores are 128-bit, 512 contiguous bytes per warp per row (guard band in the plane layout
//     makes every tile load in-bounds, no predication).
//   * Horizontal neighbours (vy of the cell to the right for the pressure update, p of the cell to the
//     left for the vy update) come from the adjacent lane by warp shuffle; vertical neighbours are the
//     thread's own registers except across warp boundaries, which exchange one row per sub-step through
//     2 KB of shared memory.  Lanes 0/31 and the top/bottom 4 rows are halo: their values go stale by
//     one cell per step, which is exactly the 4-cell halo budget, and are never stored.
//   * Each of the 4 sub-steps appends the freshly updated pressure of the tile's owned cells to the
//     time-major pressure history (evict-first stores; the 4 samples of a warp-row are 2 KB contiguous) -- the only per-step HBM traffic: 4 B per cell-step
//     instead of the 28 B of a one-step-per-launch formulation (read+write p,vx,vy + coefficient).
//   * Walls, the padding row/column and the grid-edge overrides take a general per-cell path; a
//     per-(tile, warp) lane mask precomputed after every geometry edit tells a thread whether all its
//     cells and their up/left neighbours are plain interior air, in which case it runs the branch-free
//     11-flop update.  Both paths use explicit round-to-nearest mul/add/sub (never contracted to FMA),
//     in the reference's operation order, so planes are bit-identical to the strict-fp32 CPU build.
//
// Roofline: with the state L2-resident between launches the kernel is bound by instruction issue and by
// the history write stream to HBM; see DESIGN.md for the byte accounting.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <utility>
#include <vector>
#include "pvc_internal.h"

namespace pvc
{
    __device__ __forceinline__ bool isAirF(float w) { return __float_as_uint(w) == kAirBits; }

    struct FusedArgs
    {
        const float* inP; const float* inVx; const float* inVy;
        float* outP; float* outVx; float* outVy;
        const float* coefBp; const float* coefGx; const float* coefGy;   // general-path coefficient planes
        const uint32_t* slowMask;
        const int* tileOrder;      // tiles of one source sorted by estimated cost, most expensive first
        int* firstActive;          // [source][tile][32 warps]: first launch in which the warp's block recorded a non-zero pressure
        int tilesPerSource, nsrc;
        float* hist;               // pressure history of source 0 (null: no record)
        const SourceParams* src;
        const float* pulse;
        int t0, nsteps;
        float courant;
        unsigned long long* timeline;   // debug: 8 globaltimer stamps per CTA (pvc_debug_timeline), else null
    };

    // Everything after the tile's state is in registers: K sub-steps, history append, injection, state store.
    // sVxTop / sPBot: (NW+1) x 32 float4 each; row NW of sVxTop and row 0 of sPBot stay zero (the tile's bottom /
    // top neighbours, halo of the halo).
    __device__ __forceinline__ void stamp(const FusedArgs& A, int slot, int idx = -1)
    {
        if (A.timeline && threadIdx.x == 0)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            A.timeline[(size_t)(idx < 0 ? blockIdx.x : idx) * 8 + slot] = t;
        }
    }

    // what differs between the tiles a persistent CTA processes: the time window and the output buffers
    struct TileRun
    {
        int t0, nsteps;
        float* outP; float* outVx; float* outVy;
    };
    __device__ __forceinline__ TileRun runOf(const FusedArgs& A)
    {
        TileRun r; r.t0 = A.t0; r.nsteps = A.nsteps; r.outP = A.outP; r.outVx = A.outVx; r.outVy = A.outVy;
        return r;
    }

    // barrier over the NW compute warps of a CTA (named barrier 1), so that a kernel may run extra, non-compute warps
    template <int NW>
    __device__ __forceinline__ void computeBarrier()
    {
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    }

    template <int NW, int R, bool CS>
    __device__ __forceinline__ void computeTile(const Layout& L, const FusedArgs& A, const int tx, const int ty, const int s,
                                                float (&p)[R][4], float (&vx)[R][4], float (&vy)[R][4],
                                                float4 (*sVxTop)[32], float4 (*sPBot)[32], float4* sCoefArg, const int stampIdx, const TileRun run)
    {
        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        const int rBase = ty * L.valid_rows - kTileK + wp * R;
        const int cBase = tx * kValidCols - kGuardCols + lane * 4;
        const size_t cell0 = (size_t)(rBase + kGuardRows) * L.pitch + (cBase + kGuardCols);
        const size_t src0 = (size_t)s * L.plane + cell0;

        // warp-uniform path choice (so the shuffles below never sit in divergent code):
        //   0 fast    every cell of the warp, and each one's up/left neighbour, is interior air
        //   1 edge    no wall, but the warp touches the grid edge / padding / guard band: fast arithmetic plus
        //             position-only overwrites (the absorbing-edge overrides and the zeros outside the interior)
        //   2 general some cell or neighbour is a wall: per-cell coefficient loads and the full rule
        const uint32_t mode = A.slowMask[((size_t)ty * L.tiles_x + tx) * 32 + wp];
        const bool slow = mode == 2u;
        const bool edge = mode == 1u;
        const float C = A.courant;

        // General-path warps stage their coefficient rows (bp, gx, gy) in shared memory once per launch with
        // cp.async (no registers, all 3R copies in flight together); every thread later reads back only the
        // float4s it copied itself, so cp.async.wait_group is the only synchronisation needed.
        // sCoef layout: [plane][tile row][lane] float4.
        constexpr int TR = NW * R;
        float4* const sCoef = CS ? sCoefArg : nullptr;      // compile-time null without CS: the staging code folds away
        if (CS && slow)
        {
            #pragma unroll
            for (int f = 0; f < 3; ++f)
            {
                const float* plane = (f == 0 ? A.coefBp : (f == 1 ? A.coefGx : A.coefGy)) + cell0;
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sCoef + ((size_t)f * TR + wp * R + j) * 32 + lane);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(plane + (size_t)j * L.pitch) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }

        // column classes of this thread's 4 cells, one bit per k (only consulted on the edge path)
        uint32_t colOut = 0u, colPad = 0u, colLeft = 0u;
        #pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int cc = cBase + k;
            if (cc < 0 || cc > L.gy) colOut |= 1u << k;
            if (cc == L.gy) colPad |= 1u << k;
            if (cc == 0) colLeft |= 1u << k;
        }

        // owned (stored) rows of this thread: j in [jLo, jHi) -- not halo, inside the alloc grid; empty
        // for the halo lanes 0 and 31 and for columns past the grid
        int jLo = kTileK - wp * R, jHi = NW * R - kTileK - wp * R;
        jLo = max(jLo, 0);
        jHi = min(min(jHi, R), L.rows - rBase);
        if (lane == 0 || lane == 31 || cBase >= L.cols) jHi = 0;

        // does this thread hold the pulse cell of its source?
        const SourceParams sp = A.src[s];
        const int sj = sp.cell_r - rBase, sk = sp.cell_c - cBase;
        const bool hasSrc = (sj >= 0) && (sj < R) && (sk >= 0) && (sk < 4);

        // sample t0 of this thread's 4 cells in row rBase; only dereferenced for owned rows/columns, where
        // rBase + j >= 0 and cBase >= 0.  Rows are hist_row floats apart, consecutive samples 128 floats.
        float* hist = nullptr;
        if (A.hist)
            hist = A.hist + (size_t)s * L.hist_source + (ptrdiff_t)rBase * (ptrdiff_t)L.hist_row
                 + ((ptrdiff_t)(cBase >> 7) * L.T + run.t0) * kHistChunkDefault + (cBase & 127);      // these kernels record 128-column strips

        uint32_t activity = 0u;
        // activity-hint slot of this warp's block, fetched now so its latency hides behind the steps
        int* const hintSlot = (A.firstActive && A.hist)
            ? A.firstActive + ((size_t)s * A.tilesPerSource + (size_t)ty * L.tiles_x + tx) * 32 + wp : nullptr;
        int hintKnown = 0;
        if (hintSlot && lane == 0) hintKnown = *hintSlot;
        if (CS && slow) asm volatile("cp.async.wait_group 0;" ::: "memory");
        sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
        computeBarrier<NW>();
        stamp(A, 1, stampIdx);

        #pragma unroll 1
        for (int step = 0; step < run.nsteps; ++step)
        {
            // ---------------- pressure sub-step (FDTD.cpp:125-141) ----------------
            {
                const float4 vxBelow = sVxTop[wp + 1][lane];
                const float vb[4] = { vxBelow.x, vxBelow.y, vxBelow.z, vxBelow.w };
                if (!slow && !edge)
                {
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            p[j][k] = __fsub_rn(p[j][k], __fmul_rn(C, div));
                        }
                    }
                }
                else if (edge)
                {
                    const uint32_t colDead = colOut | colPad;          // b = 0 there: pressure stays 0
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                        const int r = rBase + j;
                        const bool rowDead = (r < 0) || (r >= L.gx);
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            const float pn = __fsub_rn(p[j][k], __fmul_rn(C, div));
                            p[j][k] = (rowDead || ((colDead >> k) & 1u)) ? 0.f : pn;
                        }
                    }
                }
                else
                {
                    const float* bpRow = A.coefBp + cell0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                        const float4 b4 = CS ? sCoef[((size_t)0 * TR + wp * R + j) * 32 + lane]
                                                : __ldg(reinterpret_cast<const float4*>(bpRow + (size_t)j * L.pitch));
                        const float ba[4] = { b4.x, b4.y, b4.z, b4.w };
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            p[j][k] = (ba[k] != 0.f) ? __fsub_rn(p[j][k], __fmul_rn(C, div)) : 0.f;
                        }
                        if (!CS && (j & 1)) asm volatile("" ::: "memory");   // global fallback: bound load hoisting (register pressure)
                    }
                }
            }
            sPBot[wp + 1][lane] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
            computeBarrier<NW>();

            // ---------------- velocity sub-steps + edge overrides (FDTD.cpp:144-223) ----------------
            {
                const float4 pAbove = sPBot[wp][lane];
                const float pa[4] = { pAbove.x, pAbove.y, pAbove.z, pAbove.w };
                if (!slow && !edge)
                {
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            vx[j][k] = __fsub_rn(vx[j][k], __fmul_rn(C, __fsub_rn(p[j][k], pu)));
                            vy[j][k] = __fsub_rn(vy[j][k], __fmul_rn(C, __fsub_rn(p[j][k], pl)));
                        }
                    }
                }
                else if (edge)
                {
                    const uint32_t colDeadX = colOut | colPad;         // vx: padding column is never driven
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                        const int r = rBase + j;
                        const bool rowOut = (r < 0) || (r > L.gx);
                        const bool rowTop = (r == 0), rowPad = (r == L.gx);
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
                }
                else
                {
                    // per-cell coefficient planes (buildCoefficientsKernel): every special case -- walls, the
                    // absorbing grid edge, padding, guard band -- is data, the code is straight-line
                    const float* bpRow = A.coefBp + cell0;
                    const float* gxRow = A.coefGx + cell0;
                    const float* gyRow = A.coefGy + cell0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                        float4 b4, x4, y4;
                        if (CS)
                        {
                            b4 = sCoef[((size_t)0 * TR + wp * R + j) * 32 + lane];
                            x4 = sCoef[((size_t)1 * TR + wp * R + j) * 32 + lane];
                            y4 = sCoef[((size_t)2 * TR + wp * R + j) * 32 + lane];
                        }
                        else
                        {
                            b4 = __ldg(reinterpret_cast<const float4*>(bpRow + (size_t)j * L.pitch));
                            x4 = __ldg(reinterpret_cast<const float4*>(gxRow + (size_t)j * L.pitch));
                            y4 = __ldg(reinterpret_cast<const float4*>(gyRow + (size_t)j * L.pitch));
                        }
                        const float ba[4] = { b4.x, b4.y, b4.z, b4.w };
                        const float ga[4] = { x4.x, x4.y, x4.z, x4.w };
                        const float ha[4] = { y4.x, y4.y, y4.z, y4.w };
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            const float pt = p[j][k];
                            const bool air = ba[k] != 0.f;
                            const float airX = __fsub_rn(vx[j][k], __fmul_rn(C, __fsub_rn(pt, pu)));
                            const float airY = __fsub_rn(vy[j][k], __fmul_rn(C, __fsub_rn(pt, pl)));
                            const float wallX = __fmul_rn(ga[k], air ? pt : pu);
                            const float wallY = __fmul_rn(ha[k], air ? pt : pl);
                            vx[j][k] = (air && isAirF(ga[k])) ? airX : wallX;
                            vy[j][k] = (air && isAirF(ha[k])) ? airY : wallY;
                        }
                        if (!CS && (j & 1)) asm volatile("" ::: "memory");
                    }
                }
            }

            // ---------------- record sample t0+step (FDTD.cpp:226-231), then inject (FDTD.cpp:234) ----------------
            if (hist)
            {
                #pragma unroll
                for (int j = 0; j < R; ++j)
                    if (j >= jLo && j < jHi)
                        __stcs(reinterpret_cast<float4*>(hist + (size_t)j * L.hist_row), make_float4(p[j][0], p[j][1], p[j][2], p[j][3]));
                hist += kHistChunkDefault;
                // activity hint for the analyzer: has this warp's block recorded anything but zeros yet?  (one 3-input
                // OR per two cells; conservative -- halo rows and -0 count as activity)
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    activity |= __float_as_uint(p[j][0]) | __float_as_uint(p[j][1]);
                    activity |= __float_as_uint(p[j][2]) | __float_as_uint(p[j][3]);
                }
            }
            if (hasSrc)
            {
                // adding +0 to the three other cells of the row is exact (it can only turn -0 into +0)
                const float add = __ldg(A.pulse + run.t0 + step);
                const float a0 = (sk == 0) ? add : 0.f, a1 = (sk == 1) ? add : 0.f;
                const float a2 = (sk == 2) ? add : 0.f, a3 = (sk == 3) ? add : 0.f;
                #pragma unroll
                for (int j = 0; j < R; ++j)
                    if (j == sj)
                    {
                        p[j][0] = __fadd_rn(p[j][0], a0); p[j][1] = __fadd_rn(p[j][1], a1);
                        p[j][2] = __fadd_rn(p[j][2], a2); p[j][3] = __fadd_rn(p[j][3], a3);
                    }
            }
            sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
            computeBarrier<NW>();
            stamp(A, 2 + step, stampIdx);
        }

        // ---------------- store the owned cells of the new state ----------------
        {
            float* gp = run.outP + src0;
            float* gx = run.outVx + src0;
            float* gy = run.outVy + src0;
            #pragma unroll
            for (int j = 0; j < R; ++j)
            {
                if (j >= jLo && j < jHi)
                {
                    *reinterpret_cast<float4*>(gp + (size_t)j * L.pitch) = make_float4(p[j][0], p[j][1], p[j][2], p[j][3]);
                    *reinterpret_cast<float4*>(gx + (size_t)j * L.pitch) = make_float4(vx[j][0], vx[j][1], vx[j][2], vx[j][3]);
                    *reinterpret_cast<float4*>(gy + (size_t)j * L.pitch) = make_float4(vy[j][0], vy[j][1], vy[j][2], vy[j][3]);
                }
            }
        }
        if (hintSlot)
        {
            const bool hot = ((activity & 0x7fffffffu) != 0u) && lane >= 1 && lane <= 30;
            const unsigned any = __ballot_sync(0xffffffffu, hot);
            const int launchIndex = run.t0 / kTileK;
            if (lane == 0 && any && hintKnown > launchIndex) atomicMin(hintSlot, launchIndex);
        }
        if (A.timeline) { computeBarrier<NW>(); stamp(A, 6, stampIdx); }      // debug only: no barrier at the end of a tile otherwise
    }

    // One tile per CTA, state loaded straight from global memory into registers.
    template <int NW, int R, int MINB, bool CS>
    __global__ void __launch_bounds__(NW * 32, MINB)
    fusedStepKernel(const Layout L, const FusedArgs A)
    {
        __shared__ float4 sVxTop[NW + 1][32];   // [w]   = vx of warp w's first row (read by warp w-1)
        __shared__ float4 sPBot[NW + 1][32];    // [w+1] = p of warp w's last row  (read by warp w+1)

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        // 1-D grid in longest-first order: CTAs are dispatched in blockIdx order, so the expensive tiles (walls:
        // general path) start in the first wave instead of forming the kernel's tail
        const int s = blockIdx.x % A.nsrc;
        const int tileId = A.tileOrder[blockIdx.x / A.nsrc];
        const int ty = tileId / L.tiles_x, tx = tileId - ty * L.tiles_x;
        const int rBase = ty * L.valid_rows - kTileK + wp * R;
        const int cBase = tx * kValidCols - kGuardCols + lane * 4;
        const size_t src0 = (size_t)s * L.plane + (size_t)(rBase + kGuardRows) * L.pitch + (cBase + kGuardCols);
        stamp(A, 0);

        float p[R][4], vx[R][4], vy[R][4];
        {
            const float* gp = A.inP + src0;
            const float* gx = A.inVx + src0;
            const float* gy = A.inVy + src0;
            #pragma unroll
            for (int j = 0; j < R; ++j)
            {
                const float4 a = __ldg(reinterpret_cast<const float4*>(gp + (size_t)j * L.pitch));
                const float4 b = __ldg(reinterpret_cast<const float4*>(gx + (size_t)j * L.pitch));
                const float4 c = __ldg(reinterpret_cast<const float4*>(gy + (size_t)j * L.pitch));
                p[j][0] = a.x; p[j][1] = a.y; p[j][2] = a.z; p[j][3] = a.w;
                vx[j][0] = b.x; vx[j][1] = b.y; vx[j][2] = b.z; vx[j][3] = b.w;
                vy[j][0] = c.x; vy[j][1] = c.y; vy[j][2] = c.z; vy[j][3] = c.w;
            }
        }

        if (wp == 0)
        {
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        extern __shared__ __align__(16) float4 sCoefDyn[];        // [3][NW*R][32] float4
        computeTile<NW, R, CS>(L, A, tx, ty, s, p, vx, vy, sVxTop, sPBot, sCoefDyn, -1, runOf(A));
    }

#ifdef PVC_ALL_VARIANTS      // superseded experiments (profiles/r01_variants.txt): make EXTRA=-DPVC_ALL_VARIANTS
    // ---- persistent variant: TMA bulk-copy prefetch of the next tile through shared memory ----------------
    // One CTA per SM walks tiles b, b+G, b+2G, ...  While it steps tile i in registers, the TMA engine
    // (cp.async.bulk global->shared, mbarrier complete_tx) lands the 3*TR 512-byte rows of tile i+1 in a
    // shared-memory stage; switching tiles is then 3R conflict-free LDS.128 per thread instead of a cold
    // round trip to L2/HBM, and the stores of tile i drain while tile i+1 computes.  Without this every CTA of
    // a launch loads, computes and stores in lockstep and the memory system idles during the compute phase.
    __device__ __forceinline__ uint32_t smemAddr(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

    __device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
    }
    __device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
    }
    __device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
    {
        asm volatile(
            "{\n"
            ".reg .pred ready;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 ready, [%0], %1;\n"
            "@ready bra WAIT_DONE;\n"
            "bra WAIT_LOOP;\n"
            "WAIT_DONE:\n"
            "}\n" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
    }
    __device__ __forceinline__ void bulkLoadRow(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
    {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
    }

    template <int NW, int R>
    __global__ void __launch_bounds__(NW * 32, 1)
    fusedStepPersistentKernel(const Layout L, const FusedArgs A, const int numTiles)
    {
        constexpr int TR = NW * R;                          // tile rows incl. halo
        constexpr uint32_t kRowBytes = kTileCols * sizeof(float);
        extern __shared__ __align__(128) unsigned char smemRaw[];
        float* stage = reinterpret_cast<float*>(smemRaw);                                   // [3][TR][128]
        float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(smemRaw + 3 * TR * kRowBytes);
        float4 (*sPBot)[32] = sVxTop + (NW + 1);
        uint64_t* full = reinterpret_cast<uint64_t*>(sPBot + (NW + 1));

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        // issue the 3*TR row copies of a tile; thread i copies row i%TR of field i/TR
        auto prefetch = [&](int tile) {
            const int s = tile % A.nsrc, rem = A.tileOrder[tile / A.nsrc];
            const int ty = rem / L.tiles_x, tx = rem - ty * L.tiles_x;
            if (threadIdx.x == 0) mbarExpectTx(full, 3u * TR * kRowBytes);
            for (int i = threadIdx.x; i < 3 * TR; i += NW * 32)
            {
                const int f = i / TR, row = i - f * TR;
                const float* plane = (f == 0) ? A.inP : (f == 1 ? A.inVx : A.inVy);
                const float* src = plane + (size_t)s * L.plane + (size_t)(ty * L.valid_rows + row) * L.pitch + (size_t)tx * kValidCols;
                bulkLoadRow(stage + (size_t)i * kTileCols, src, kRowBytes, full);
            }
        };

        if (threadIdx.x == 0) mbarInit(full, 1);
        if (wp == 0)
        {
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();

        int tile = blockIdx.x;
        if (tile < numTiles) prefetch(tile);
        uint32_t parity = 0;
        while (tile < numTiles)
        {
            const int s = tile % A.nsrc, rem = A.tileOrder[tile / A.nsrc];
            const int ty = rem / L.tiles_x, tx = rem - ty * L.tiles_x;

            stamp(A, 0, tile);
            mbarWait(full, parity);
            parity ^= 1u;
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
            __syncthreads();                                   // every thread has drained the stage: refill it
            const int next = tile + gridDim.x;
            if (next < numTiles) prefetch(next);

            computeTile<NW, R, false>(L, A, tx, ty, s, p, vx, vy, sVxTop, sPBot, nullptr, tile, runOf(A));
            tile = next;
        }
    }

    // ---- persistent, TMA-fed variant -------------------------------------------------------------------------
    // One CTA per SM pulls tiles from a global counter (cost-sorted order, dynamic balancing).  While the CTA steps
    // tile i in registers, ONE elected thread has already asked the TMA engine for tile i+1: three 3-D tensor copies
    // (cp.async.bulk.tensor, box 128 x TR x 1 of the p / vx / vy planes) that complete on an mbarrier.  Switching
    // tiles is then 3R conflict-free LDS.128 per thread.  General-path warps stage their coefficient rows in shared
    // memory with cp.async (computeTile<.., CS = true>).
    __device__ __forceinline__ void tmaLoadTile3d(void* dstSmem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar)
    {
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smemAddr(dstSmem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smemAddr(bar)) : "memory");
    }

    template <int NW, int R, bool CS, int NS>
    __global__ void __launch_bounds__(NW * 32, 1)
    fusedStepTmaKernel(const Layout L, const FusedArgs A, const int numTiles, int* __restrict__ tileCounter,
                       const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapVx,
                       const __grid_constant__ CUtensorMap mapVy)
    {
        // NS = number of shared-memory stages = how many tiles ahead the TMA engine runs (1 or 2)
        constexpr int TR = NW * R;
        constexpr uint32_t kPlaneBytes = TR * kTileCols * sizeof(float);
        extern __shared__ __align__(128) unsigned char smemRaw[];
        float* stageBase = reinterpret_cast<float*>(smemRaw);                                    // [NS][3][TR][128]
        float4* sCoef = reinterpret_cast<float4*>(smemRaw + NS * 3 * kPlaneBytes);             // [3][TR][32] (CS only)
        unsigned char* tail = smemRaw + NS * 3 * kPlaneBytes + (CS ? 3 * kPlaneBytes : 0);
        float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(tail);
        float4 (*sPBot)[32] = sVxTop + (NW + 1);
        uint64_t* full = reinterpret_cast<uint64_t*>(sPBot + (NW + 1));                         // [NS]
        volatile int* sQueue = reinterpret_cast<volatile int*>(full + NS);                      // [NS] tile index held by each stage
        volatile int* sCoord = sQueue + NS;                                                     // [NS][3] its (source, tx, ty)

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;

        auto decode = [&](int order, int& s, int& tx, int& ty) {
            s = order % A.nsrc;
            const int id = A.tileOrder[order / A.nsrc];
            ty = id / L.tiles_x; tx = id - ty * L.tiles_x;
        };
        auto issue = [&](int order, int st) {                   // thread 0 only
            int s, tx, ty;
            decode(order, s, tx, ty);
            sCoord[st * 3 + 0] = s; sCoord[st * 3 + 1] = tx; sCoord[st * 3 + 2] = ty;
            float* stage = stageBase + (size_t)st * 3 * TR * kTileCols;
            mbarExpectTx(full + st, 3u * kPlaneBytes);
            tmaLoadTile3d(stage, &mapP, tx * kValidCols, ty * L.valid_rows, s, full + st);
            tmaLoadTile3d(stage + (size_t)TR * kTileCols, &mapVx, tx * kValidCols, ty * L.valid_rows, s, full + st);
            tmaLoadTile3d(stage + (size_t)2 * TR * kTileCols, &mapVy, tx * kValidCols, ty * L.valid_rows, s, full + st);
        };

        if (threadIdx.x == 0)
        {
            #pragma unroll
            for (int st = 0; st < NS; ++st) mbarInit(full + st, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            #pragma unroll
            for (int st = 0; st < NS; ++st)
            {
                const int t = atomicAdd(tileCounter, 1);
                sQueue[st] = t;
                if (t < numTiles) issue(t, st);
            }
        }
        if (wp == 0)
        {
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        int pending = 0;                                        // thread 0: next tile index, fetched one tile early
        if (threadIdx.x == 0) pending = atomicAdd(tileCounter, 1);
        int st = 0;
        uint32_t parity = 0;                                    // parity of the stage-0 barrier; stage 1 flips on wrap too
        while (true)
        {
            const int tile = sQueue[st];
            if (tile >= numTiles) break;                        // indices only grow: once a stage is empty every later one is
            const int s = sCoord[st * 3 + 0], tx = sCoord[st * 3 + 1], ty = sCoord[st * 3 + 2];
            stamp(A, 0, tile);
            mbarWait(full + st, parity);
            const float* stage = stageBase + (size_t)st * 3 * TR * kTileCols;
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
            __syncthreads();                                   // stage drained by every thread (and sQueue[st] read): refill it
            if (threadIdx.x == 0)
            {
                sQueue[st] = pending;
                if (pending < numTiles) issue(pending, st);
                pending = atomicAdd(tileCounter, 1);
            }
            computeTile<NW, R, CS>(L, A, tx, ty, s, p, vx, vy, sVxTop, sPBot, sCoef, tile, runOf(A));   // its barriers publish sQueue[st]
            if (++st == NS) { st = 0; parity ^= 1u; }
        }
    }

    // ---- generational persistent variant: the whole time loop in one launch, no grid-wide synchronisation ------------
    // Work items are (generation g, tile): generation g advances every tile from step 4g to 4g+4.  CTAs (one per SM)
    // pull items from a global counter in (g, cost-sorted tile) order.  An item may start as soon as the tile itself and
    // its up-to-8 neighbours of the same source have finished generation g-1 (their outputs are this item's halo, and
    // that also covers the write-after-read on the ping-pong buffer), which is tracked with one completed-generation
    // counter per tile (st.release / ld.acquire at gpu scope).  Nothing ever waits for "everybody": the expensive wall
    // tiles of generation g overlap the cheap tiles of g+1, there is no launch ramp or tail every 4 steps, and the TMA
    // prefetch of the next item runs straight across generation boundaries.  Deadlock-free because all CTAs are
    // co-resident and an item only depends on items handed out before it; a bounded spin sets an abort flag otherwise.
    struct GenArgs
    {
        float* state[2][3];       // ping-pong planes: generation g reads [g & 1], writes [(g + 1) & 1]
        int* doneGen;             // [nsrc * tilesPerSource] generations completed by each tile
        int* workCounter;         // next work item of this launch
        int* abortFlag;           // set if a dependency wait times out (host reports PVC_ERR_CUDA)
        int gen0, numGen;         // generations covered by this launch
        int T;                    // total number of steps (the last generation may be short)
    };
    struct TensorMaps6 { CUtensorMap m[6]; };     // [buffer][field]

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

    template <int NW, int R, bool CS>
    __global__ void __launch_bounds__(NW * 32, 1)
    fusedStepGenKernel(const Layout L, const FusedArgs A0, const GenArgs G, const int numTiles,
                       const __grid_constant__ TensorMaps6 maps)
    {
        constexpr int TR = NW * R;
        constexpr uint32_t kPlaneBytes = TR * kTileCols * sizeof(float);
        extern __shared__ __align__(128) unsigned char smemRaw[];
        float* stage = reinterpret_cast<float*>(smemRaw);                                        // [3][TR][128]
        float4* sCoef = reinterpret_cast<float4*>(smemRaw + 3 * kPlaneBytes);                  // [3][TR][32] (CS only)
        unsigned char* tail = smemRaw + 3 * kPlaneBytes + (CS ? 3 * kPlaneBytes : 0);
        float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(tail);
        float4 (*sPBot)[32] = sVxTop + (NW + 1);
        uint64_t* full = reinterpret_cast<uint64_t*>(sPBot + (NW + 1));
        volatile int* sItem = reinterpret_cast<volatile int*>(full + 1);     // [0] work item, [1] source, [2] tx, [3] ty, [4] generation

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        const int total = G.numGen * numTiles;
        const int tps = A0.tilesPerSource;

        // ---- thread-0 helpers ----
        // executed by ALL lanes of warp 0: lane k < 9 probes neighbour k (one L2 round trip for the whole stencil of tiles)
        auto depsReady = [&](int s, int tx, int ty, int gen, bool block) -> bool {
            if (gen == 0) return true;
            const int dx = lane % 3 - 1, dy = lane / 3 - 1;
            const int nx = tx + dx, ny = ty + dy;
            const bool mine = lane < 9 && nx >= 0 && ny >= 0 && nx < L.tiles_x && ny < L.tiles_y;
            const int* slot = G.doneGen + (size_t)s * tps + (mine ? ny * L.tiles_x + nx : 0);
            unsigned spins = 0;
            while (true)
            {
                const bool ok = !mine || loadAcquire(slot) >= gen;
                if (__all_sync(0xffffffffu, ok)) return true;
                if (!block) return false;
                __nanosleep(64);
                ++spins;
                bool giveUp = false;
                if ((spins & 0xffu) == 0u) giveUp = spins > (1u << 22) || *(volatile int*)G.abortFlag;
                if (__any_sync(0xffffffffu, giveUp)) { if (lane == 0) atomicExch(G.abortFlag, 1); return false; }
            }
        };
        auto issue = [&](int s, int tx, int ty, int gen) {
            asm volatile("fence.proxy.async;" ::: "memory");     // order the acquires above before the async-proxy reads below
            const CUtensorMap* m = maps.m + 3 * (gen & 1);
            mbarExpectTx(full, 3u * kPlaneBytes);
            tmaLoadTile3d(stage, m + 0, tx * kValidCols, ty * L.valid_rows, s, full);
            tmaLoadTile3d(stage + (size_t)TR * kTileCols, m + 1, tx * kValidCols, ty * L.valid_rows, s, full);
            tmaLoadTile3d(stage + (size_t)2 * TR * kTileCols, m + 2, tx * kValidCols, ty * L.valid_rows, s, full);
        };
        auto decodeItem = [&](int w, int& s, int& tx, int& ty, int& gen) {
            const int order = w % numTiles;
            gen = G.gen0 + w / numTiles;
            s = order % A0.nsrc;
            const int id = A0.tileOrder[order / A0.nsrc];
            ty = id / L.tiles_x; tx = id - ty * L.tiles_x;
        };
        auto publishItem = [&](int w, int s, int tx, int ty, int gen) {
            sItem[0] = w; sItem[1] = s; sItem[2] = tx; sItem[3] = ty; sItem[4] = gen;
        };

        // warp-0 state (uniform across its lanes)
        int pending = 0;                 // next work item index, fetched one tile early
        bool nextIssued = false;         // TMA for the published next item already in flight
        int prevSlot = -1, prevGen = 0;  // finished tile whose completion is not published yet
        auto fetchItem = [&]() -> int {                         // warp 0: one atomic, broadcast
            int v = 0;
            if (lane == 0) v = atomicAdd(G.workCounter, 1);
            return __shfl_sync(0xffffffffu, v, 0);
        };

        if (wp == 0)
        {
            if (lane == 0)
            {
                mbarInit(full, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            const int w = fetchItem();
            int s = 0, tx = 0, ty = 0, gen = 0;
            bool ok = false;
            if (w < total)
            {
                decodeItem(w, s, tx, ty, gen);
                ok = depsReady(s, tx, ty, gen, true);
            }
            if (lane == 0)
            {
                if (ok) { publishItem(w, s, tx, ty, gen); issue(s, tx, ty, gen); }
                else publishItem(total, 0, 0, 0, 0);
            }
            pending = fetchItem();
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();

        uint32_t parity = 0;
        while (true)
        {
            const int w = sItem[0];
            if (w >= total) break;
            const int s = sItem[1], tx = sItem[2], ty = sItem[3], gen = sItem[4];
            mbarWait(full, parity);
            parity ^= 1u;
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
            __syncthreads();                                   // stage drained; sItem consumed; the previous tile's stores were
                                                               // all issued before the barrier that ended it
            if (wp == 0)
            {
                // release the previous tile (its stores were issued about a microsecond ago, so the release is cheap by
                // now): one st.release.gpu by lane 0 after the CTA barrier orders every thread's stores before it
                if (prevSlot >= 0) { if (lane == 0) storeRelease(G.doneGen + prevSlot, prevGen + 1); prevSlot = -1; }
                nextIssued = true;
                if (pending < total)
                {
                    int ns, ntx, nty, ngen;
                    decodeItem(pending, ns, ntx, nty, ngen);
                    nextIssued = depsReady(ns, ntx, nty, ngen, false);
                    if (lane == 0)
                    {
                        publishItem(pending, ns, ntx, nty, ngen);
                        if (nextIssued) issue(ns, ntx, nty, ngen);
                    }
                }
                else if (lane == 0) publishItem(total, 0, 0, 0, 0);
            }

            TileRun run;
            run.t0 = gen * kTileK;
            run.nsteps = min(kTileK, G.T - gen * kTileK);
            run.outP = G.state[(gen + 1) & 1][0]; run.outVx = G.state[(gen + 1) & 1][1]; run.outVy = G.state[(gen + 1) & 1][2];
            computeTile<NW, R, CS>(L, A0, tx, ty, s, p, vx, vy, sVxTop, sPBot, sCoef, w, run);

            const bool needSlowPath = __syncthreads_or(wp == 0 && !nextIssued);      // also: every store of this tile is issued
            if (wp == 0) { prevSlot = s * tps + ty * L.tiles_x + tx; prevGen = gen; }
            if (needSlowPath)
            {   // the next item waits on tiles that were not finished when we probed (possibly on this very tile):
                // publish our completion now, then wait for real
                if (wp == 0)
                {
                    if (lane == 0) storeRelease(G.doneGen + prevSlot, prevGen + 1);
                    prevSlot = -1;
                    const int ns = sItem[1], ntx = sItem[2], nty = sItem[3], ngen = sItem[4];
                    const bool ok = depsReady(ns, ntx, nty, ngen, true);
                    if (lane == 0) { if (ok) issue(ns, ntx, nty, ngen); else publishItem(total, 0, 0, 0, 0); }
                }
                __syncthreads();
            }
            if (wp == 0) pending = fetchItem();
        }
        // publish the last tile
        __syncthreads();
        if (wp == 0 && lane == 0 && prevSlot >= 0) storeRelease(G.doneGen + prevSlot, prevGen + 1);
    }

    // ---- warp-specialised generational variant -------------------------------------------------------------------
    // Same work-item / dependency scheme as fusedStepGenKernel, but all scheduling lives in ONE extra producer warp so
    // the NW compute warps never touch a global counter or flag:
    //   producer: fetch item -> probe the 9 dependency counters (one lane each) -> wait "stage empty" -> publish the
    //             item's coordinates in shared memory -> expect_tx + 3 TMA tensor copies;  then wait "tile done" of
    //             the tile being computed and st.release its completion counter (the producer, not the compute warps,
    //             absorbs the store-drain latency of the release).
    //   compute : wait "full" (TMA landed) -> LDS stage into registers -> barrier, arrive "empty" -> 4 steps ->
    //             barrier, arrive "done".
    // Three mbarriers (full / empty / done), one shared-memory stage.
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

    template <int NW, int R, bool CS>
    __global__ void __launch_bounds__((NW + 1) * 32, 1)
    fusedStepWsKernel(const Layout L, const FusedArgs A0, const GenArgs G, const int numTiles,
                      const __grid_constant__ TensorMaps6 maps)
    {
        constexpr int TR = NW * R;
        constexpr uint32_t kPlaneBytes = TR * kTileCols * sizeof(float);
        extern __shared__ __align__(128) unsigned char smemRaw[];
        float* stage = reinterpret_cast<float*>(smemRaw);                                        // [3][TR][128]
        float4* sCoef = reinterpret_cast<float4*>(smemRaw + 3 * kPlaneBytes);                  // [3][TR][32] (CS only)
        unsigned char* tail = smemRaw + 3 * kPlaneBytes + (CS ? 3 * kPlaneBytes : 0);
        float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(tail);
        float4 (*sPBot)[32] = sVxTop + (NW + 1);
        uint64_t* full = reinterpret_cast<uint64_t*>(sPBot + (NW + 1));
        uint64_t* empty = full + 1;
        uint64_t* done = full + 2;
        volatile int* sItem = reinterpret_cast<volatile int*>(full + 3);     // [0] valid, [1] source, [2] tx, [3] ty, [4] generation

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        const int total = G.numGen * numTiles;
        const int tps = A0.tilesPerSource;

        if (threadIdx.x == 0)
        {
            mbarInit(full, 1); mbarInit(empty, 1); mbarInit(done, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (wp == 0)
        {
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();

        if (wp == NW)
        {
            // ================= producer warp =================
            auto depsReady = [&](int s, int tx, int ty, int gen, bool block) -> bool {
                if (gen == 0) return true;
                const int dx = lane % 3 - 1, dy = lane / 3 - 1;
                const int nx = tx + dx, ny = ty + dy;
                const bool mine = lane < 9 && nx >= 0 && ny >= 0 && nx < L.tiles_x && ny < L.tiles_y;
                const int* slot = G.doneGen + (size_t)s * tps + (mine ? ny * L.tiles_x + nx : 0);
                unsigned spins = 0;
                while (true)
                {
                    const bool ok = !mine || loadAcquire(slot) >= gen;
                    if (__all_sync(0xffffffffu, ok)) return true;
                    if (!block) return false;
                    __nanosleep(32);
                    ++spins;
                    bool giveUp = false;
                    if ((spins & 0xffu) == 0u) giveUp = spins > (1u << 22) || *(volatile int*)G.abortFlag;
                    if (__any_sync(0xffffffffu, giveUp)) { if (lane == 0) atomicExch(G.abortFlag, 1); return false; }
                }
            };
            uint32_t emptyParity = 0, doneParity = 0;
            bool stageBusy = false;                       // a tile has been handed to the compute warps and not yet drained
            int prevSlot = -1, prevGen = 0;               // tile being computed, completion not yet published
            bool alive = true;
            auto publishPrev = [&]() -> bool {            // wait for the compute warps to finish it, then release its counter
                if (prevSlot < 0) return true;
                bool ok = true;
                if (lane == 0) ok = mbarWaitBounded(done, doneParity, G.abortFlag);
                ok = __shfl_sync(0xffffffffu, ok, 0);
                doneParity ^= 1u;
                if (ok && lane == 0) storeRelease(G.doneGen + prevSlot, prevGen + 1);
                prevSlot = -1;
                return ok;
            };
            while (alive)
            {
                int w = 0;
                if (lane == 0) w = atomicAdd(G.workCounter, 1);
                w = __shfl_sync(0xffffffffu, w, 0);
                if (w >= total) break;
                const int order = w % numTiles;
                const int gen = G.gen0 + w / numTiles;
                const int s = order % A0.nsrc;
                const int id = A0.tileOrder[order / A0.nsrc];
                const int ty = id / L.tiles_x, tx = id - ty * L.tiles_x;

                bool ready = depsReady(s, tx, ty, gen, false);
                if (!ready)
                {   // it may depend on the tile our own compute warps are working on: publish that first, then wait for real
                    if (!publishPrev()) { alive = false; break; }
                    ready = depsReady(s, tx, ty, gen, true);
                    if (!ready) { alive = false; break; }
                }
                if (stageBusy)
                {
                    bool ok = true;
                    if (lane == 0) ok = mbarWaitBounded(empty, emptyParity, G.abortFlag);
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    emptyParity ^= 1u;
                    if (!ok) { alive = false; break; }
                }
                if (lane == 0)
                {
                    sItem[0] = 1; sItem[1] = s; sItem[2] = tx; sItem[3] = ty; sItem[4] = gen;
                    asm volatile("fence.proxy.async;" ::: "memory");
                    const CUtensorMap* m = maps.m + 3 * (gen & 1);
                    mbarExpectTx(full, 3u * kPlaneBytes);
                    tmaLoadTile3d(stage, m + 0, tx * kValidCols, ty * L.valid_rows, s, full);
                    tmaLoadTile3d(stage + (size_t)TR * kTileCols, m + 1, tx * kValidCols, ty * L.valid_rows, s, full);
                    tmaLoadTile3d(stage + (size_t)2 * TR * kTileCols, m + 2, tx * kValidCols, ty * L.valid_rows, s, full);
                }
                __syncwarp();
                stageBusy = true;
                // the tile handed over before this one is (or was) being computed: publish it once the compute warps are done
                if (!publishPrev()) { alive = false; break; }
                prevSlot = s * tps + ty * L.tiles_x + tx; prevGen = gen;
            }
            // drain: publish the last tile, then tell the compute warps to stop
            if (alive) alive = publishPrev();
            if (stageBusy && alive)
            {
                bool ok = true;
                if (lane == 0) ok = mbarWaitBounded(empty, emptyParity, G.abortFlag);
                (void)ok;
            }
            if (lane == 0) { sItem[0] = 0; mbarArrive(full); }          // wake the compute warps with "no more work"
            return;
        }

        // ================= compute warps =================
        uint32_t fullParity = 0;
        while (true)
        {
            if (!mbarWaitBounded(full, fullParity, G.abortFlag)) break;
            fullParity ^= 1u;
            if (sItem[0] == 0) break;
            const int s = sItem[1], tx = sItem[2], ty = sItem[3], gen = sItem[4];
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
            computeBarrier<NW>();                               // every compute thread has drained the stage and read sItem
            if (threadIdx.x == 0) mbarArrive(empty);

            TileRun run;
            run.t0 = gen * kTileK;
            run.nsteps = min(kTileK, G.T - gen * kTileK);
            run.outP = G.state[(gen + 1) & 1][0]; run.outVx = G.state[(gen + 1) & 1][1]; run.outVy = G.state[(gen + 1) & 1][2];
            computeTile<NW, R, CS>(L, A0, tx, ty, s, p, vx, vy, sVxTop, sPBot, sCoef, -1, run);

            computeBarrier<NW>();                               // every store of this tile has been issued
            if (threadIdx.x == 0) mbarArrive(done);
        }
    }

    // Per-cell coefficients of the general path, rebuilt after every geometry edit from the wall plane w.
#endif // PVC_ALL_VARIANTS

    // With bp = 1 for an interior air cell (reference b = 1) and 0 otherwise, the reference's three update
    // rules (FDTD.cpp:125-223 incl. the grid-edge overrides) collapse to data:
    //   p  <- bp ? p - C*div : 0
    //   vx <- bp ? (gx == AIR ? vx - C*(p - p_up) : gx * p) : gx * p_up        (same for vy with p_left, gy)
    // where gx is  AIR       interior air cell with an interior air cell above it
    //              -Y_up     air cell under a wall            (reference: -(Y_n * p))
    //              -1        air cell in row 0                (vx = -p, FDTD.cpp:208)
    //              +Y_self   wall cell under an air cell      (reference: Y * p_prev)
    //              +1        padding row gx                   (vx = p_up, FDTD.cpp:209)
    //              0         everything else (wall under wall, padding column, guard band)
    // Multiplying by +-1 or 0 and negating a product are exact, so the planes reproduce the branch form bit for bit.
    __global__ void buildCoefficientsKernel(const Layout L, const float* __restrict__ w,
                                            float* __restrict__ bp, float* __restrict__ gx, float* __restrict__ gy)
    {
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= L.plane) return;
        const int r = (int)(i / L.pitch) - kGuardRows, c = (int)(i % L.pitch) - kGuardCols;
        const float AIR = __uint_as_float(kAirBits);
        const bool inRows = (r >= 0) && (r < L.gx), inCols = (c >= 0) && (c < L.gy);
        const float wSelf = w[i];
        const bool air = inRows && inCols && isAirF(wSelf);
        float cx = 0.f, cy = 0.f;
        if (inCols && r >= 0 && r <= L.gx)
        {
            if (r == 0) cx = air ? -1.f : 0.f;
            else if (r == L.gx) cx = 1.f;
            else
            {
                const float wUp = w[i - L.pitch];
                const bool airUp = isAirF(wUp);
                cx = air ? (airUp ? AIR : -wUp) : (airUp ? wSelf : 0.f);
            }
        }
        if (inRows && c >= 0 && c <= L.gy)
        {
            if (c == 0) cy = air ? -1.f : 0.f;
            else if (c == L.gy) cy = 1.f;
            else
            {
                const float wLeft = w[i - 1];
                const bool airLeft = isAirF(wLeft);
                cy = air ? (airLeft ? AIR : -wLeft) : (airLeft ? wSelf : 0.f);
            }
        }
        bp[i] = air ? 1.f : 0.f;
        gx[i] = cx;
        gy[i] = cy;
    }

    // path mode per (tile, warp), see fusedStepKernel: 0 fast, 1 edge, 2 general.  Only INTERIOR cells can be
    // walls that matter: the padding row/column is b = 0 whatever an AABB wrote there (Grid.cpp:94-97,276-280)
    // and its admittance is never used (every velocity next to it is an edge override).
    template <int NW, int R, int MINB>
    __global__ void slowMaskKernel(const Layout L, const float* __restrict__ w, uint32_t* __restrict__ mask)
    {
        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        const int tx = blockIdx.x, ty = blockIdx.y;
        const int rBase = ty * L.valid_rows - kTileK + wp * R;
        const int cBase = tx * kValidCols - kGuardCols + lane * 4;
        bool wall = false, edge = false;
        for (int j = -1; j < R; ++j)
            for (int k = -1; k < 4; ++k)
            {
                const int r = rBase + j, c = cBase + k;
                const bool interior = (r >= 0) && (r < L.gx) && (c >= 0) && (c < L.gy);
                if (interior)
                {
                    if (__float_as_uint(w[cellIndex(L, r, c)]) != kAirBits) wall = true;
                    if ((j >= 0 && r == 0) || (k >= 0 && c == 0)) edge = true;
                }
                else if (j >= 0 && k >= 0) edge = true;       // an own cell outside the interior
            }
        const uint32_t anyWall = __ballot_sync(0xffffffffu, wall);
        const uint32_t anyEdge = __ballot_sync(0xffffffffu, edge);
        if (lane == 0) mask[((size_t)ty * L.tiles_x + tx) * 32 + wp] = anyWall ? 2u : (anyEdge ? 1u : 0u);
    }

    // tile variants (pvc_config::reserved): warps per CTA, rows per thread, min CTAs per SM, kind
    //   kind 0 one launch per 4 steps (fusedStepKernel)      1 persistent        2 TMA persistent   3 generational
    //        4 first warp-specialised generational           5 ws2 (pvc_step_ws2.cu)                6 resident (pvc_step_res.cu)
    // The default build carries what the product selects -- 47 / 50 (ws2), 63..67 and 69..72 (resident), 18 (fallback without the TMA
    // driver entry point) -- plus step_kernel = 1 (two-launch baseline, pvc_step.cu).  Everything else documents the
    // search (profiles/r01_variants.txt) and is compiled only with make EXTRA=-DPVC_ALL_VARIANTS.
    struct Variant { int nw, r, minBlocks, persistent, builtin; };
    static const Variant kVariants[] = { {8, 6, 2, 0, 1}, {8, 8, 2, 0, 0}, {16, 4, 2, 0, 0}, {16, 8, 1, 0, 0}, {8, 4, 4, 0, 0}, {8, 8, 1, 0, 0},
                                         {16, 4, 1, 0, 0}, {12, 8, 1, 0, 0}, {8, 8, 1, 1, 0}, {14, 8, 1, 1, 0},
                                         {24, 4, 1, 0, 0}, {16, 6, 1, 0, 0}, {20, 4, 1, 0, 0}, {24, 4, 1, 1, 0}, {16, 6, 1, 1, 0}, {12, 8, 1, 1, 0},
                                         {10, 4, 2, 0, 0}, {12, 4, 2, 0, 0}, {8, 6, 2, 0, 1}, {10, 6, 2, 0, 0}, {8, 6, 2, 0, 0}, {20, 4, 1, 0, 0},
                                         {16, 4, 1, 2, 0}, {12, 6, 1, 2, 0}, {20, 4, 1, 2, 0}, {16, 4, 1, 2, 0}, {12, 8, 1, 2, 0},
                                         {16, 4, 1, 2, 0}, {12, 4, 1, 2, 0}, {10, 4, 1, 2, 0}, {16, 4, 1, 3, 0}, {16, 4, 1, 3, 0}, {12, 6, 1, 3, 0}, {16, 4, 1, 4, 0}, {16, 4, 1, 4, 0}, {12, 6, 1, 4, 0}, {15, 4, 1, 4, 0}, {15, 4, 1, 4, 0}, {11, 6, 1, 4, 0},
                                         {14, 4, 1, 5, 0}, {15, 4, 1, 5, 0}, {30, 2, 1, 5, 0}, {20, 3, 1, 5, 0}, {14, 4, 1, 5, 0},
                                         {10, 8, 1, 5, 0}, {11, 6, 1, 5, 0}, {12, 6, 1, 5, 0},         // 39..46: pvc_step_ws2.cu experiments
                                         {14, 4, 1, 5, 1}, {15, 4, 1, 5, 0}, {14, 4, 1, 5, 0},           // 47 (default for large batches), 48, 49
                                         {8, 4, 1, 5, 1}, {10, 4, 1, 5, 0},                              // 50 (default when few work items), 51
                                         {12, 4, 1, 5, 0}, {12, 5, 1, 5, 0},                             // 52, 53
                                         {8, 6, 2, 0, 0}, {8, 6, 2, 0, 0}, {8, 6, 2, 0, 0}, {8, 6, 2, 0, 0}, {8, 6, 2, 0, 0}, {8, 6, 2, 0, 0},   // 54..59 unused
                                         {8, 4, 2, 6, 0}, {10, 4, 2, 6, 0}, {12, 4, 2, 6, 0}, {16, 4, 1, 6, 1}, {20, 4, 1, 6, 1}, {18, 4, 1, 6, 1},    // 60..65: resident (pvc_step_res.cu); 60..62 (CTA barrier per sub-step, two CTAs per SM) superseded by 69 / 72 / 70
                                         {16, 5, 1, 6, 1}, {4, 4, 4, 6, 1},         // 66: resident, 5 rows per warp; 67: resident, 4-warp tiles for tiny grids
                                         {10, 4, 2, 6, 0}, {8, 4, 2, 6, 1},         // 68 / 69: resident, 10 / 8 warps, two CTAs per SM, barrier-free row exchange (68 superseded by 72)
                                         {12, 4, 1, 6, 1}, {14, 4, 1, 6, 1}, {10, 4, 1, 6, 1} };      // 70 / 71 / 72: resident, 12 / 14 / 10 warps, one CTA per SM, barrier-free row exchange
    static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

    bool variantAvailable(int variant)
    {
        if (variant < 0 || variant >= kNumVariants || (variant >= 54 && variant <= 59)) return false;
#ifdef PVC_ALL_VARIANTS
        return true;
#else
        return kVariants[variant].builtin != 0;
#endif
    }
    int variantKind(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].persistent;
    }
    int variantMinBlocks(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].minBlocks;
    }
    int variantWarps(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].nw;
    }

    int fusedTileRows(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].nw * kVariants[variant].r;
    }

    static FusedArgs makeArgs(pvc_solver* s, float* hist, int t, int t1)
    {
        FusedArgs A;
        float** in = s->state[s->cur];
        float** out = s->state[s->cur ^ 1];
        A.inP = in[0]; A.inVx = in[1]; A.inVy = in[2];
        A.outP = out[0]; A.outVx = out[1]; A.outVy = out[2];
        A.coefBp = s->coef[0]; A.coefGx = s->coef[1]; A.coefGy = s->coef[2]; A.slowMask = s->slowMask; A.tileOrder = s->tileOrder; A.tilesPerSource = s->L.tiles_x * s->L.tiles_y; A.firstActive = s->firstActive;
        A.hist = hist;
        A.src = s->src; A.pulse = s->pulse;
        A.t0 = t; A.nsteps = (t1 - t < kTileK) ? (t1 - t) : kTileK;
        A.courant = s->cfg.courant;
        A.timeline = s->timeline;
#ifdef PVC_TUNING
        if (s->timeline) { static const char* dbg = getenv("PVC_DEBUG_NSTEPS"); if (dbg) A.nsteps = atoi(dbg); }   // debug: memory-floor probe (results invalid)
#endif
        return A;
    }

    int fusedHistChunk(int variant)
    {
        return variant == 43 ? kValidCols : kHistChunkDefault;      // the TMA-store variant records dense 120-column boxes
    }

    int fusedWarpRows(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].r;
    }

    template <int NW, int R, int MINB, bool CS = false>
    static int launchVariant(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        dim3 grid(L.tiles_x * L.tiles_y * nsrc), block(NW * 32);
        const size_t smem = CS ? (size_t)3 * NW * R * 32 * sizeof(float4) : 0;
        static bool configured[64] = {};
        if (!configured[s->device & 63])
        {
            cudaError_t e = cudaFuncSetAttribute(fusedStepKernel<NW, R, MINB, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { setError("fused kernel smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            configured[s->device & 63] = true;
        }
        for (int t = t0; t < t1; t += kTileK)
        {
            FusedArgs A = makeArgs(s, hist, t, t1);
            A.nsrc = nsrc;
            fusedStepKernel<NW, R, MINB, CS><<<grid, block, smem, s->stream>>>(L, A);
            s->cur ^= 1;
            *launches += 1;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("fused step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

#ifdef PVC_ALL_VARIANTS
    template <int NW, int R>
    static int launchPersistent(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        constexpr int TR = NW * R;
        const size_t smem = (size_t)3 * TR * kTileCols * sizeof(float) + (size_t)2 * (NW + 1) * 32 * sizeof(float4) + 16;
        static bool configured[64] = {};
        if (!configured[s->device & 63])
        {
            cudaError_t e = cudaFuncSetAttribute(fusedStepPersistentKernel<NW, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { setError("persistent kernel smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            configured[s->device & 63] = true;
        }
        const int numTiles = L.tiles_x * L.tiles_y * nsrc;
        const int grid = numTiles < s->numSMs ? numTiles : s->numSMs;
        for (int t = t0; t < t1; t += kTileK)
        {
            FusedArgs A = makeArgs(s, hist, t, t1);
            A.nsrc = nsrc;
            fusedStepPersistentKernel<NW, R><<<grid, NW * 32, smem, s->stream>>>(L, A, numTiles);
            s->cur ^= 1;
            *launches += 1;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("persistent fused step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    template <int NW, int R, bool CS, int NS = 1>
    static int launchTma(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        constexpr int TR = NW * R;
        if (!s->tmaReady || s->tmaTileRows != TR) { setError("TMA variant: tensor maps not built for %d-row tiles", TR); return PVC_ERR_INVALID; }
        const size_t plane = (size_t)TR * kTileCols * sizeof(float);
        const size_t smem = NS * 3 * plane + (CS ? 3 * plane : 0) + (size_t)2 * (NW + 1) * 32 * sizeof(float4) + 128;
        static bool configured[64] = {};
        if (!configured[s->device & 63])
        {
            cudaError_t e = cudaFuncSetAttribute(fusedStepTmaKernel<NW, R, CS, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { setError("TMA kernel smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            configured[s->device & 63] = true;
        }
        const int numTiles = L.tiles_x * L.tiles_y * nsrc;
        const int grid = numTiles < s->numSMs ? numTiles : s->numSMs;
        const int nLaunch = (t1 - t0 + kTileK - 1) / kTileK;
        if (nLaunch > s->tileCounterCount) { setError("TMA variant: %d launches exceed the counter pool", nLaunch); return PVC_ERR_INVALID; }
        cudaMemsetAsync(s->tileCounters, 0, sizeof(int) * (size_t)nLaunch, s->stream);
        int k = 0;
        for (int t = t0; t < t1; t += kTileK, ++k)
        {
            FusedArgs A = makeArgs(s, hist, t, t1);
            A.nsrc = nsrc;
            const CUtensorMap* m = reinterpret_cast<const CUtensorMap*>(s->tensorMaps) + 3 * s->cur;
            fusedStepTmaKernel<NW, R, CS, NS><<<grid, NW * 32, smem, s->stream>>>(L, A, numTiles, s->tileCounters + k, m[0], m[1], m[2]);
            s->cur ^= 1;
            *launches += 1;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("TMA fused step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    template <int NW, int R, bool CS, bool WS = false>
    static int launchGen(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        constexpr int TR = NW * R;
        if (!s->tmaReady || s->tmaTileRows != TR) { setError("generational variant: tensor maps not built for %d-row tiles", TR); return PVC_ERR_INVALID; }
        if (t0 != 0 || s->cur != 0) { setError("generational variant: must start at step 0"); return PVC_ERR_INVALID; }
        const size_t plane = (size_t)TR * kTileCols * sizeof(float);
        const size_t smem = 3 * plane + (CS ? 3 * plane : 0) + (size_t)2 * (NW + 1) * 32 * sizeof(float4) + 128;
        static bool configured[64] = {};
        if (!configured[s->device & 63])
        {
            cudaError_t e;
            if constexpr (WS) e = cudaFuncSetAttribute(fusedStepWsKernel<NW, R, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            else e = cudaFuncSetAttribute(fusedStepGenKernel<NW, R, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { setError("generational kernel smem opt-in (%zu B): %s", smem, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
            configured[s->device & 63] = true;
        }
        int maxCtas = 0;
        if constexpr (WS) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxCtas, fusedStepWsKernel<NW, R, CS>, (NW + 1) * 32, smem);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxCtas, fusedStepGenKernel<NW, R, CS>, NW * 32, smem);
        if (maxCtas < 1) { setError("generational kernel does not fit an SM"); return PVC_ERR_CUDA; }
        const int numTiles = L.tiles_x * L.tiles_y * nsrc;
        const int grid = numTiles < s->numSMs ? numTiles : s->numSMs;          // all CTAs must be co-resident
        const int gens = (t1 + kTileK - 1) / kTileK;
        const int perLaunch = 256;                                             // generations per launch (bounds kernel time)
        cudaMemsetAsync(s->doneGen, 0, sizeof(int) * (size_t)numTiles, s->stream);
        cudaMemsetAsync(s->tileCounters, 0, sizeof(int) * (size_t)((gens + perLaunch - 1) / perLaunch + 1), s->stream);
        FusedArgs A = makeArgs(s, hist, 0, t1);
        A.nsrc = nsrc;
        GenArgs G;
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 3; ++f) G.state[b][f] = s->state[b][f];
        G.doneGen = s->doneGen; G.abortFlag = s->tileCounters;                  // slot 0 of the pool is the abort flag
        G.T = t1;
        TensorMaps6 maps;
        memcpy(&maps, s->tensorMaps, sizeof(maps));
        int k = 1;
        for (int g0 = 0; g0 < gens; g0 += perLaunch, ++k)
        {
            G.gen0 = g0; G.numGen = (gens - g0 < perLaunch) ? (gens - g0) : perLaunch;
            G.workCounter = s->tileCounters + k;
            if constexpr (WS) fusedStepWsKernel<NW, R, CS><<<grid, (NW + 1) * 32, smem, s->stream>>>(L, A, G, numTiles, maps);
            else fusedStepGenKernel<NW, R, CS><<<grid, NW * 32, smem, s->stream>>>(L, A, G, numTiles, maps);
            *launches += 1;
        }
        s->cur = gens & 1;
        s->checkAbort = 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("generational fused step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

#endif // PVC_ALL_VARIANTS

    // 3-D tensor maps {pitch, rows_alloc, sources} of the six state planes, box 128 x tileRows x 1 (driver entry point
    // fetched through the runtime, no libcuda link)
    int buildTensorMaps(pvc_solver* s)
    {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
        { cudaGetLastError(); s->tmaReady = 0; return PVC_OK; }           // TMA variants unavailable; the default kernel does not need them
        const Layout& L = s->L;
        CUtensorMap* maps = reinterpret_cast<CUtensorMap*>(s->tensorMaps);
        for (int b = 0; b < 2; ++b)
            for (int f = 0; f < 3; ++f)
            {
                const cuuint64_t dims[3] = { (cuuint64_t)L.pitch, (cuuint64_t)L.rows_alloc, (cuuint64_t)s->cfg.max_sources };
                const cuuint64_t strides[2] = { (cuuint64_t)L.pitch * sizeof(float), (cuuint64_t)L.plane * sizeof(float) };
                const cuuint32_t box[3] = { (cuuint32_t)kTileCols, (cuuint32_t)L.tile_rows, 1u };
                const cuuint32_t estr[3] = { 1u, 1u, 1u };
                const CUresult r = ((EncodeFn)fn)(&maps[b * 3 + f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->state[b][f], dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { s->tmaReady = 0; return PVC_OK; }
            }
        s->tmaReady = 1;
        s->tmaTileRows = L.tile_rows;
        return PVC_OK;
    }

    template <int NW, int R, int MINB>
    static int maskVariant(pvc_solver* s)
    {
        const Layout& L = s->L;
        slowMaskKernel<NW, R, MINB><<<dim3(L.tiles_x, L.tiles_y), NW * 32, 0, s->stream>>>(L, s->w, s->slowMask);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("slow mask launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        // The generational / resident kernels keep the row-major tile order: every dependency of an item is then about one
        // generation old, whereas "expensive tiles first" makes the wall tiles at the head of generation g wait for
        // neighbours at the tail of generation g-1 (measured: 7 % slower).  The one-launch-per-4-steps kernels sort by
        // cost (general-path warps ~3x, edge-path ~1.3x a fast warp) so the expensive tiles do not form the launch's tail.
        const int tiles = L.tiles_x * L.tiles_y;
        bool natural = variantKind(s->cfg.reserved) >= 4;
#ifdef PVC_TUNING
        { static const char* orderEnv = getenv("PVC_TILE_ORDER"); if (orderEnv) natural = orderEnv[0] == 'n'; }
#endif
        // row-major order needs no read-back: a frame loop that edits geometry every frame then never waits for the
        // device here.  The identity order is uploaded once.
        if (natural && s->tileOrderNatural == 1) { s->slowMaskDirty = 0; return PVC_OK; }
        std::vector<std::pair<int, int>> cost((size_t)tiles);
        if (natural) { for (int t = 0; t < tiles; ++t) cost[(size_t)t] = std::make_pair(0, t); }
        else
        {
            std::vector<uint32_t> modes((size_t)tiles * 32);
            if (cudaMemcpyAsync(modes.data(), s->slowMask, sizeof(uint32_t) * modes.size(), cudaMemcpyDeviceToHost, s->stream) != cudaSuccess ||
                cudaStreamSynchronize(s->stream) != cudaSuccess)
            { setError("slow mask readback: %s", cudaGetErrorString(cudaGetLastError())); return PVC_ERR_CUDA; }
            for (int t = 0; t < tiles; ++t)
            {
                int c = 0;
                for (int wIdx = 0; wIdx < NW; ++wIdx) { const uint32_t m = modes[(size_t)t * 32 + wIdx]; c += (m == 2u) ? 30 : (m == 1u ? 13 : 10); }
                cost[(size_t)t] = std::make_pair(-c, t);
            }
            std::sort(cost.begin(), cost.end());
        }
        s->tileOrderNatural = natural ? 1 : 0;
        std::vector<int> order((size_t)tiles);
        for (int t = 0; t < tiles; ++t) order[(size_t)t] = cost[(size_t)t].second;
        if (cudaMemcpyAsync(s->tileOrder, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
            cudaStreamSynchronize(s->stream) != cudaSuccess)
        { setError("tile order upload: %s", cudaGetErrorString(cudaGetLastError())); return PVC_ERR_CUDA; }
        s->slowMaskDirty = 0;
        return PVC_OK;
    }

    int launchFusedSteps(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const int v = s->cfg.reserved;
        if (!variantAvailable(v)) { setError("step-kernel variant %d is not compiled into this build (make EXTRA=-DPVC_ALL_VARIANTS)", v); return PVC_ERR_INVALID; }
        if (variantKind(v) == 5) return launchWs2Steps(s, v, nsrc, t0, t1, hist, launches);
        if (variantKind(v) == 6) return launchResidentSteps(s, v, nsrc, t0, t1, hist, launches);
        switch (v)
        {
            case 0: case 18: return launchVariant<8, 6, 2>(s, nsrc, t0, t1, hist, launches);
#ifdef PVC_ALL_VARIANTS
            case 1: return launchVariant<8, 8, 2>(s, nsrc, t0, t1, hist, launches);
            case 2: return launchVariant<16, 4, 2>(s, nsrc, t0, t1, hist, launches);
            case 3: return launchVariant<16, 8, 1>(s, nsrc, t0, t1, hist, launches);
            case 4: return launchVariant<8, 4, 4>(s, nsrc, t0, t1, hist, launches);
            case 5: return launchVariant<8, 8, 1>(s, nsrc, t0, t1, hist, launches);
            case 6: return launchVariant<16, 4, 1>(s, nsrc, t0, t1, hist, launches);
            case 7: return launchVariant<12, 8, 1>(s, nsrc, t0, t1, hist, launches);
            case 8: return launchPersistent<8, 8>(s, nsrc, t0, t1, hist, launches);
            case 9: return launchPersistent<14, 8>(s, nsrc, t0, t1, hist, launches);
            case 10: return launchVariant<24, 4, 1>(s, nsrc, t0, t1, hist, launches);
            case 11: return launchVariant<16, 6, 1>(s, nsrc, t0, t1, hist, launches);
            case 12: return launchVariant<20, 4, 1>(s, nsrc, t0, t1, hist, launches);
            case 13: return launchPersistent<24, 4>(s, nsrc, t0, t1, hist, launches);
            case 14: return launchPersistent<16, 6>(s, nsrc, t0, t1, hist, launches);
            case 15: return launchPersistent<12, 8>(s, nsrc, t0, t1, hist, launches);
            case 16: return launchVariant<10, 4, 2>(s, nsrc, t0, t1, hist, launches);
            case 17: return launchVariant<12, 4, 2>(s, nsrc, t0, t1, hist, launches);
            case 19: return launchVariant<10, 6, 2>(s, nsrc, t0, t1, hist, launches);
            case 20: return launchVariant<8, 6, 2, true>(s, nsrc, t0, t1, hist, launches);
            case 21: return launchVariant<20, 4, 1, true>(s, nsrc, t0, t1, hist, launches);
            case 22: return launchTma<16, 4, true>(s, nsrc, t0, t1, hist, launches);
            case 23: return launchTma<12, 6, false>(s, nsrc, t0, t1, hist, launches);
            case 24: return launchTma<20, 4, false>(s, nsrc, t0, t1, hist, launches);
            case 25: return launchTma<16, 4, false>(s, nsrc, t0, t1, hist, launches);
            case 26: return launchTma<12, 8, false>(s, nsrc, t0, t1, hist, launches);
            case 27: return launchTma<16, 4, false, 2>(s, nsrc, t0, t1, hist, launches);
            case 28: return launchTma<12, 4, false, 2>(s, nsrc, t0, t1, hist, launches);
            case 29: return launchTma<10, 4, true, 2>(s, nsrc, t0, t1, hist, launches);
            case 30: return launchGen<16, 4, true>(s, nsrc, t0, t1, hist, launches);
            case 31: return launchGen<16, 4, false>(s, nsrc, t0, t1, hist, launches);
            case 32: return launchGen<12, 6, false>(s, nsrc, t0, t1, hist, launches);
            case 33: return launchGen<16, 4, true, true>(s, nsrc, t0, t1, hist, launches);
            case 34: return launchGen<16, 4, false, true>(s, nsrc, t0, t1, hist, launches);
            case 35: return launchGen<12, 6, false, true>(s, nsrc, t0, t1, hist, launches);
            case 36: return launchGen<15, 4, true, true>(s, nsrc, t0, t1, hist, launches);
            case 37: return launchGen<15, 4, false, true>(s, nsrc, t0, t1, hist, launches);
            case 38: return launchGen<11, 6, true, true>(s, nsrc, t0, t1, hist, launches);
#endif
            default: setError("step-kernel variant %d: no launcher", v); return PVC_ERR_INVALID;
        }
    }

    int rebuildSlowMask(pvc_solver* s)
    {
        const int v = s->cfg.reserved;
        if (!variantAvailable(v)) { setError("step-kernel variant %d is not compiled into this build (make EXTRA=-DPVC_ALL_VARIANTS)", v); return PVC_ERR_INVALID; }
        buildCoefficientsKernel<<<(unsigned)((s->L.plane + 255) / 256), 256, 0, s->stream>>>(s->L, s->w, s->coef[0], s->coef[1], s->coef[2]);
        if (variantKind(v) == 5) { int rc = rebuildWs2Descriptors(s, v); if (!rc) rc = rebuildResidentDescriptors(s, v); if (rc) return rc; }     // + the linear-form planes
        if (variantKind(v) == 6) { const int rc = rebuildResidentDescriptors(s, v); if (rc) return rc; }
        switch (kVariants[v].nw * 100 + kVariants[v].r)
        {
            case 806: return maskVariant<8, 6, 1>(s);
            case 804: return maskVariant<8, 4, 1>(s);
            case 404: return maskVariant<4, 4, 1>(s);
            case 1004: return maskVariant<10, 4, 1>(s);
            case 1204: return maskVariant<12, 4, 1>(s);
            case 1404: return maskVariant<14, 4, 1>(s);
            case 1604: return maskVariant<16, 4, 1>(s);
            case 2004: return maskVariant<20, 4, 1>(s);
            case 1804: return maskVariant<18, 4, 1>(s);
            case 1605: return maskVariant<16, 5, 1>(s);
#ifdef PVC_ALL_VARIANTS
            case 1008: return maskVariant<10, 8, 1>(s);
            case 3002: return maskVariant<30, 2, 1>(s);
            case 2003: return maskVariant<20, 3, 1>(s);
            case 808: return maskVariant<8, 8, 1>(s);
            case 1608: return maskVariant<16, 8, 1>(s);
            case 1408: return maskVariant<14, 8, 1>(s);
            case 2404: return maskVariant<24, 4, 1>(s);
            case 1606: return maskVariant<16, 6, 1>(s);
            case 1205: return maskVariant<12, 5, 1>(s);
            case 1006: return maskVariant<10, 6, 1>(s);
            case 1206: return maskVariant<12, 6, 1>(s);
            case 1504: return maskVariant<15, 4, 1>(s);
            case 1106: return maskVariant<11, 6, 1>(s);
            case 1208: return maskVariant<12, 8, 1>(s);
#endif
            default: setError("step-kernel variant %d: no mask kernel for %d x %d tiles", v, kVariants[v].nw, kVariants[v].r); return PVC_ERR_INVALID;
        }
    }
}

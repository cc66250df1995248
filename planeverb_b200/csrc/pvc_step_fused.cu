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
#include "pvc_internal.h"

namespace pvc
{
    __device__ __forceinline__ bool isAirF(float w) { return __float_as_uint(w) == kAirBits; }

    // general velocity rule (FDTD.cpp:149-168 in branch form), written as selects so the general path stays
    // straight-line code
    __device__ __forceinline__ float ruleF(float v, float pThis, float pPrev, float wThis, float wPrev, float courant)
    {
        const bool aThis = isAirF(wThis), aPrev = isAirF(wPrev);
        const float airAir = __fsub_rn(v, __fmul_rn(courant, __fsub_rn(pThis, pPrev)));
        const float wallThis = __fmul_rn(wThis, pPrev);
        const float wallPrev = -__fmul_rn(wPrev, pThis);
        const float ifAirThis = aPrev ? airAir : wallPrev;
        const float ifWallThis = aPrev ? wallThis : 0.f;
        return aThis ? ifAirThis : ifWallThis;
    }

    struct FusedArgs
    {
        const float* inP; const float* inVx; const float* inVy;
        float* outP; float* outVx; float* outVy;
        const float* w;
        const uint32_t* slowMask;
        float* hist;               // pressure history of source 0 (null: no record)
        const SourceParams* src;
        const float* pulse;
        int t0, nsteps;
        float courant;
    };

    template <int NW, int R, int MINB>
    __global__ void __launch_bounds__(NW * 32, MINB)
    fusedStepKernel(const Layout L, const FusedArgs A)
    {
        // row NW of sVxTop and row 0 of sPBot stay zero: the tile's bottom / top neighbours (halo of the halo)
        __shared__ float4 sVxTop[NW + 1][32];   // [w]   = vx of warp w's first row (read by warp w-1)
        __shared__ float4 sPBot[NW + 1][32];    // [w+1] = p of warp w's last row  (read by warp w+1)

        const int lane = threadIdx.x & 31;
        const int wp = threadIdx.x >> 5;
        const int tx = blockIdx.x, ty = blockIdx.y, s = blockIdx.z;
        // domain coordinates of this thread's first cell (may be negative / beyond the grid: guard band)
        const int rBase = ty * L.valid_rows - kTileK + wp * R;
        const int cBase = tx * kValidCols - kGuardCols + lane * 4;
        const size_t cell0 = (size_t)(rBase + kGuardRows) * L.pitch + (cBase + kGuardCols);
        const size_t src0 = (size_t)s * L.plane + cell0;

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

        // warp-uniform path choice (so the shuffles below never sit in divergent code):
        //   0 fast    every cell of the warp, and each one's up/left neighbour, is interior air
        //   1 edge    no wall, but the warp touches the grid edge / padding / guard band: fast arithmetic plus
        //             position-only overwrites (the absorbing-edge overrides and the zeros outside the interior)
        //   2 general some cell or neighbour is a wall: per-cell coefficient loads and the full rule
        const uint32_t mode = A.slowMask[((size_t)ty * L.tiles_x + tx) * 32 + wp];
        const bool slow = mode == 2u;
        const bool edge = mode == 1u;
        const float C = A.courant;

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
                 + ((ptrdiff_t)(cBase >> 7) * L.T + A.t0) * kHistChunk + (cBase & 127);

        if (wp == 0)
        {
            sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
        __syncthreads();

        #pragma unroll 1
        for (int step = 0; step < A.nsteps; ++step)
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
                    const float* wrow = A.w + cell0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)j * L.pitch));
                        const float wa[4] = { w4.x, w4.y, w4.z, w4.w };
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                            const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                            const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                            p[j][k] = isAirF(wa[k]) ? __fsub_rn(p[j][k], __fmul_rn(C, div)) : 0.f;
                        }
                        asm volatile("" ::: "memory");      // keep the general path row-by-row: low register pressure
                    }
                }
            }
            sPBot[wp + 1][lane] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
            __syncthreads();

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
                    // the row above / column left of the very first tile row / column lies outside the allocation
                    const bool haveUp = (rBase + kGuardRows) > 0, haveLeft = (cBase + kGuardCols) > 0;
                    const float* wrow = A.w + cell0;
                    float4 wPrevRow = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (haveUp) wPrevRow = __ldg(reinterpret_cast<const float4*>(wrow - L.pitch));
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + (size_t)j * L.pitch));
                        const float wLeft = haveLeft ? __ldg(wrow + (size_t)j * L.pitch - 1) : 0.f;
                        const float wa[5] = { wLeft, w4.x, w4.y, w4.z, w4.w };
                        const float wu[4] = { wPrevRow.x, wPrevRow.y, wPrevRow.z, wPrevRow.w };
                        const int r = rBase + j;
                        const bool rowDead = (r < 0) || (r > L.gx);
                        #pragma unroll
                        for (int k = 0; k < 4; ++k)
                        {
                            const int cc = cBase + k;
                            const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                            const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                            const float pt = p[j][k];
                            float nx = ruleF(vx[j][k], pt, pu, wa[k + 1], wu[k], C);
                            float ny = ruleF(vy[j][k], pt, pl, wa[k + 1], wa[k], C);
                            if (r == 0) nx = -pt;                                   // FDTD.cpp:208
                            if (r == L.gx) nx = pu;                                 // FDTD.cpp:209
                            if (cc == 0) ny = -pt;                                  // FDTD.cpp:220
                            if (cc == L.gy) ny = pl;                                // FDTD.cpp:221
                            if (rowDead || cc < 0 || cc >= L.gy) nx = 0.f;          // padding column / guard band
                            if (rowDead || r == L.gx || cc < 0 || cc > L.gy) ny = 0.f;   // padding row / guard band
                            vx[j][k] = nx; vy[j][k] = ny;
                        }
                        wPrevRow = w4;
                        asm volatile("" ::: "memory");
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
                hist += kHistChunk;
            }
            if (hasSrc)
            {
                // adding +0 to the three other cells of the row is exact (it can only turn -0 into +0)
                const float add = __ldg(A.pulse + A.t0 + step);
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
            __syncthreads();
        }

        // ---------------- store the owned cells of the new state ----------------
        {
            float* gp = A.outP + src0;
            float* gx = A.outVx + src0;
            float* gy = A.outVy + src0;
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

    struct Variant { int nw, r, minBlocks; };
    static const Variant kVariants[] = { {12, 8, 1}, {8, 8, 2}, {16, 4, 2}, {16, 8, 1}, {8, 4, 4}, {8, 8, 1}, {16, 4, 1} };
    static const int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

    int fusedTileRows(int variant)
    {
        if (variant < 0 || variant >= kNumVariants) variant = 0;
        return kVariants[variant].nw * kVariants[variant].r;
    }

    template <int NW, int R, int MINB>
    static int launchVariant(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        dim3 grid(L.tiles_x, L.tiles_y, nsrc), block(NW * 32);
        for (int t = t0; t < t1; t += kTileK)
        {
            FusedArgs A;
            float** in = s->state[s->cur];
            float** out = s->state[s->cur ^ 1];
            A.inP = in[0]; A.inVx = in[1]; A.inVy = in[2];
            A.outP = out[0]; A.outVx = out[1]; A.outVy = out[2];
            A.w = s->w; A.slowMask = s->slowMask;
            A.hist = hist;
            A.src = s->src; A.pulse = s->pulse;
            A.t0 = t; A.nsteps = (t1 - t < kTileK) ? (t1 - t) : kTileK;
            A.courant = s->cfg.courant;
            fusedStepKernel<NW, R, MINB><<<grid, block, 0, s->stream>>>(L, A);
            s->cur ^= 1;
            *launches += 1;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("fused step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    template <int NW, int R, int MINB>
    static int maskVariant(pvc_solver* s)
    {
        const Layout& L = s->L;
        slowMaskKernel<NW, R, MINB><<<dim3(L.tiles_x, L.tiles_y), NW * 32, 0, s->stream>>>(L, s->w, s->slowMask);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("slow mask launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        s->slowMaskDirty = 0;
        return PVC_OK;
    }

    #define PVC_DISPATCH(fn, ...)                                             \
        switch (v) {                                                          \
            case 1: return fn<8, 8, 2>(__VA_ARGS__);                          \
            case 2: return fn<16, 4, 2>(__VA_ARGS__);                         \
            case 3: return fn<16, 8, 1>(__VA_ARGS__);                         \
            case 4: return fn<8, 4, 4>(__VA_ARGS__);                          \
            case 5: return fn<8, 8, 1>(__VA_ARGS__);                          \
            case 6: return fn<16, 4, 1>(__VA_ARGS__);                         \
            default: return fn<12, 8, 1>(__VA_ARGS__);                        \
        }

    int launchFusedSteps(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        int v = s->cfg.reserved; if (v < 0 || v >= kNumVariants) v = 0;
        PVC_DISPATCH(launchVariant, s, nsrc, t0, t1, hist, launches)
    }

    int rebuildSlowMask(pvc_solver* s)
    {
        int v = s->cfg.reserved; if (v < 0 || v >= kNumVariants) v = 0;
        PVC_DISPATCH(maskVariant, s)
    }
}

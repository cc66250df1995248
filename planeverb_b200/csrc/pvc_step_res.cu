// pvc_step_res.cu -- "resident" step kernel: the whole solve of a source in ONE launch with the state in registers.
//
// Same numerics as every other step kernel here: Grid::GenerateResponseCPU's time step
// (ProjectPlaneverb/src/FDTD/FDTD.cpp:122-235) in the reference's operation order, explicit round-to-nearest fp32
// mul / add / sub, K = 4 steps per pass over tiles of (NW*R) x 128 cells incl. a 4-cell halo, a warp owning R rows x 128
// columns and a lane one float4 of each row.  What differs from the generational kernel (pvc_step_ws2.cu):
//
//   * A tile belongs to ONE CTA for the whole solve and its p / vx / vy never leave the registers.  Between passes a
//     CTA writes only the 4-cell-wide strips its neighbours need (first / last four owned rows, lanes 1 and 30 of the
//     others) into a mailbox plane and reads its own halo ring back from it: ~50 KB per tile and pass through the L2
//     instead of 155 KB (full tile in through TMA, owned cells out), no TMA stage to drain, no store burst, no producer /
//     publisher warps.  Hand-over WITHOUT fences, flags or CTA barriers: a strip cell travels as ONE 16-byte word
//     {p, vx, vy, tag} (tag = solve epoch << 16 | pass), two words per 256-bit strong access, so a reader that sees the
//     tag it expects in a word has the data that came with it (the low-latency protocol of collective libraries; it
//     relies on an aligned 16-byte unit being read and written indivisibly, which tools/micro/vec16_atomicity.cu stresses
//     on the device for 128- and 256-bit accesses and every parity test would expose).  The halo loads ARE the poll:
//     eight loads in flight per thread, repeated until every tag matches.  Measured against the first version of this
//     kernel (one counter per tile, CTA barrier + st.release, acquire polls, then reloads: a hand-over chain of 4.2 us in
//     an 8.3 us pass, profiles/r02_resident_trace.txt): ~1.9 us of a 6.3 us pass (profiles/r02_flow_exchange.txt).
//   * Inside a tile the warps exchange their boundary rows WITHOUT barriers as well (pvc_internal.h, namespace flow): tagged
//     rows in shared memory, polled by the reader after it has updated the rows that do not need them.
//     Eight mailbox slots (pass & 7) keep a writer from overwriting cells a lagging warp of a neighbour still has to read.
//     All CTAs are co-resident (cooperative launch); a batch of sources that does not fit is solved a few sources per
//     launch, one launch after the other.
//   * Walls, the absorbing grid edge, the padding row / column and the guard band are DATA in a form that costs what
//     the interior-air path costs.  With p == 0 in every cell that is not interior air (an invariant of the scheme:
//     such a cell's pressure coefficient is 0 and nothing is ever injected there), the four cases of the reference's
//     velocity rule and its edge overrides collapse to
//         vx' = k * vx - c * (p - p_up)        k in {0, 1}, c in {Courant, Y_wall, 1, 0}
//         p'  = p - cP * div                    cP in {Courant, 0}
//     -- (air, air): k = 1, c = Courant; air under a wall: p_up == 0, c = Y_up -> -(Y_up * p); wall under air: p == 0,
//     c = Y_self -> Y_self * p_up; row 0: c = 1 -> -p; padding row: c = 1 -> p_up; everything else 0.  k * vx is exact
//     and -(c * d) is an exactly negated product, so fma(k, vx, -(c * d)) rounds once, exactly like the reference's
//     separate multiply and subtract: three instructions per velocity component, as in the air path.  Three coefficient
//     planes (cP, and k folded into the sign of c: -Courant marks k = 1) are staged once per solve in shared memory.
//     Admittances must be >= 0 (absorption in [0, 1]); pvc_apply_geometry rejects anything else.
//
// Selected automatically (pvc_api.cu::resolveVariant, by estimated pass time) when every tile of at least one source fits the
// GPU at once and a launch is reasonably full: one listener up to 1024 x 1024, batches of small grids, batches of 1024^2
// sources (the headline bench).  The reference's own contract (70^2 .. 191^2 cells, one listener) runs on 4-warp tiles.
//
// Roofline: HBM by the 28 B / cell-update accounting of SURVEY.md 8d; physically the kernel
// moves the 4-byte history record per cell-step and the strips (4.2 B of DRAM traffic per cell-update), and is bound by
// instruction issue inside a pass and by the neighbour hand-over between passes.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <utility>
#include <vector>
#include "pvc_internal.h"

namespace pvc
{
    namespace res
    {
        // one mailbox word: 16 bytes, one strong 128-bit access each way (see the header)
        __device__ __forceinline__ void storeWord(float4* q, float a, float b, float c, int tag)
        {
            asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(q), "f"(a), "f"(b), "f"(c), "f"(__int_as_float(tag)) : "memory");
        }
        __device__ __forceinline__ float4 loadWord(const float4* q)
        {
            float4 v;
            asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(q) : "memory");
            return v;
        }
        // two neighbouring words in one 256-bit access (sm_100: LDG / STG .256).  Every 16-byte half still carries its own tag, so the
        // protocol needs no more than 16-byte indivisibility; cells 0 / 1 and 2 / 3 of a thread's row start 32-byte aligned
        // (column base a multiple of 4, pitch a multiple of 32).  Halves the instructions of a ring fetch (16 -> 8 per thread).
        struct WordPair { float4 a, b; };
        __device__ __forceinline__ void storeWordPair(float4* q, float a0, float b0, float c0, float a1, float b1, float c1, int tag)
        {
            const float t = __int_as_float(tag);
            asm volatile("st.relaxed.gpu.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(q), "f"(a0), "f"(b0), "f"(c0), "f"(t), "f"(a1), "f"(b1), "f"(c1), "f"(t) : "memory");
        }
        __device__ __forceinline__ WordPair loadWordPair(const float4* q)
        {
            WordPair v;
            asm volatile("ld.relaxed.gpu.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w) : "l"(q) : "memory");
            return v;
        }
        constexpr int kSlots = 8;
        // A mailbox word is a register QUAD {p, vx, vy, tag} of ONE cell, while the step loop wants the four p (vx, vy) of a ROW in
        // a quad (128-bit shared / history accesses).  Left to itself the register allocator coalesces the two and pays with ~30
        // moves per time step; copying through an integer op with a run-time zero keeps the mailbox quads temporaries.
        __device__ __forceinline__ float opaqueCopy(float v, int zero) { return __int_as_float(__float_as_int(v) ^ zero); }
        // Row exchange between the warps of a tile in the step loop (a warp needs one row of the warp above and one of the warp below
        // per time step):
        //   kSyncCta   rows through shared memory, one CTA barrier per sub-step (the small tiles that run two CTAs per SM: their pass
        //              is bound by the neighbour hand-over, not by the step loop)
        //   kSyncFlow  no barrier at all: the exchanged rows carry tags and the reader polls the row itself -- after it has updated the
        //              rows that do not need it (pvc_internal.h, namespace flow).  1024^2, one listener, 18-warp tiles: 2.13 -> 1.85 ms
        //              per 1000 steps against named barriers per shared edge (which needed a barrier shared by the warps 15 .. NW-1
        //              beyond 16 warps; one shared-memory mbarrier per edge was 1.8x slower than the CTA barrier: try_wait costs ~90
        //              cycles even when the phase is over; profiles/r02_resident_variants.txt)
        enum { kSyncCta = 0, kSyncFlow = 3 };
        template <int NW, int SYNC>
        __device__ __forceinline__ void phaseSync(int)
        {
            if (SYNC == kSyncCta) __syncthreads();          // (kSyncFlow never gets here: its step loop has no barriers)
        }
        // 1.0f where x < 0 (one FSET): the k of the linear-form velocity rule, folded into the sign of its coefficient
        __device__ __forceinline__ float signFlag(float x)
        {
            float r;
            asm("set.lt.f32.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(x));
            return r;
        }

        struct Args
        {
            float* p0; float* vx0; float* vy0;     // ping-pong buffer 0 (pass g reads buffer g & 1, writes the other)
            float* p1; float* vx1; float* vy1;
            float* hist;
            const uint32_t* mode;                  // [tile][32] path mode per warp (slowMaskKernel): 0 = interior air, else coefficient path
            const float* cP; const float* sX; const float* sY;
            int* firstActive;                      // [source][tile][32]
            const SourceParams* src;
            const float* pulse;
            float4* xchg;                          // mailbox [kSlots][max_sources][rows_alloc][pitch] of {p, vx, vy, tag}
            size_t xchgSlot;                       // float4s between slots: max_sources * plane
            int tagBase;                           // solve epoch << 16
            int zero;                              // 0, unknown to the compiler (see opaqueCopy)
            int* abortFlag;
            int tilesPerSource, s0, nsrc;          // this launch solves sources s0 .. s0 + nsrc - 1
            int numGen, T;
            float courant;
            int loadState;                         // start from the state planes (buffer 0) instead of from zero: a later chunk of a streamed solve
            int finalPass;                         // the launch ends the response (0: a chunk of a streamed solve that another chunk follows)
#ifdef PVC_TUNING
            int dbg;                               // tuning builds only: bit 0 no history stores, bit 1 no neighbour wait / halo reload (both: results invalid), bit 2 no nanosleep in the poll
            unsigned long long* trace;             // per CTA x traced pass x 8 %globaltimer stamps (null: off)
#endif
        };
#ifdef PVC_TUNING
        constexpr int kTracePass0 = 40, kTracePasses = 32, kTraceSlots = 8;
        __device__ __forceinline__ void stamp(const Args& A, int g, int slot)
        {
            if (A.trace && threadIdx.x == 0 && g >= kTracePass0 && g < kTracePass0 + kTracePasses)
            {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                A.trace[((size_t)blockIdx.x * kTracePasses + (g - kTracePass0)) * kTraceSlots + slot] = t;
            }
        }
        __device__ __forceinline__ void stampAny(const Args& A, int g, int slot)
        {
            if (A.trace && g >= kTracePass0 && g < kTracePass0 + kTracePasses)
            {
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                A.trace[((size_t)blockIdx.x * kTracePasses + (g - kTracePass0)) * kTraceSlots + slot] = t;
            }
        }
        #define PVC_STAMP(A, g, slot) stamp(A, g, slot)
        #define PVC_STAMP_IF(cond, A, g, slot) do { if (cond) stampAny(A, g, slot); } while (0)
        #define PVC_DBG(A, bit) (((A).dbg >> (bit)) & 1)
#else
        #define PVC_STAMP(A, g, slot) do {} while (0)
        #define PVC_STAMP_IF(cond, A, g, slot) do {} while (0)
        #define PVC_DBG(A, bit) 0
#endif

        // ---- one time step of a warp's R x 128 block; GEN = coefficient path (walls / edges / guard band as data) ----
        // rows J0 .. J1 - 1 of the block (the whole block by default; kSyncFlow updates the row that needs the neighbour's data last)
        template <int R, bool GEN, int J0 = 0, int J1 = R>
        __device__ __forceinline__ void pressureStep(float (&p)[R][4], const float (&vx)[R][4], const float (&vy)[R][4], const float4 vxBelow,
                                                     const float C, const float4* __restrict__ cP)
        {
            const float vb[4] = { vxBelow.x, vxBelow.y, vxBelow.z, vxBelow.w };
            #pragma unroll
            for (int j = J0; j < J1; ++j)
            {
                const float vyRight = __shfl_down_sync(0xffffffffu, vy[j][0], 1);
                float ck[4] = { C, C, C, C };
                if (GEN) { const float4 c4 = cP[j * 32]; ck[0] = c4.x; ck[1] = c4.y; ck[2] = c4.z; ck[3] = c4.w; }
                #pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const float vxd = (j + 1 < R) ? vx[j + 1][k] : vb[k];
                    const float vyr = (k < 3) ? vy[j][k + 1] : vyRight;
                    const float div = __fadd_rn(__fsub_rn(vxd, vx[j][k]), __fsub_rn(vyr, vy[j][k]));
                    p[j][k] = __fsub_rn(p[j][k], __fmul_rn(ck[k], div));
                }
            }
        }
        template <int R, bool GEN, int J0 = 0, int J1 = R>
        __device__ __forceinline__ void velocityStep(const float (&p)[R][4], float (&vx)[R][4], float (&vy)[R][4], const float4 pAbove,
                                                     const float C, const float4* __restrict__ sX, const float4* __restrict__ sY)
        {
            const float pa[4] = { pAbove.x, pAbove.y, pAbove.z, pAbove.w };
            #pragma unroll
            for (int j = J0; j < J1; ++j)
            {
                const float pLeft = __shfl_up_sync(0xffffffffu, p[j][3], 1);
                if (!GEN)
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
                else
                {
                    const float4 x4 = sX[j * 32], y4 = sY[j * 32];
                    const float gx[4] = { x4.x, x4.y, x4.z, x4.w };
                    const float gy[4] = { y4.x, y4.y, y4.z, y4.w };
                    #pragma unroll
                    for (int k = 0; k < 4; ++k)
                    {
                        const float pu = (j > 0) ? p[j - 1][k] : pa[k];
                        const float pl = (k > 0) ? p[j][k - 1] : pLeft;
                        // fma(k, v, -(|s| * d)): k * v is exact (k is 0 or 1), so this is the reference's v - c*d (k = 1) or -(c*d)
                        const float tx = __fmul_rn(fabsf(gx[k]), __fsub_rn(p[j][k], pu));
                        const float ty = __fmul_rn(fabsf(gy[k]), __fsub_rn(p[j][k], pl));
                        vx[j][k] = __fmaf_rn(signFlag(gx[k]), vx[j][k], -tx);
                        vy[j][k] = __fmaf_rn(signFlag(gy[k]), vy[j][k], -ty);
                    }
                }
            }
        }

        template <int NW, int R>
        struct Smem
        {
            static constexpr int TR = NW * R;
            static constexpr size_t offVxTop = 0;
            static constexpr size_t offPBot = offVxTop + (size_t)(NW + 1) * 64 * sizeof(float4);       // kSyncFlow: rows of 2 x 32 float4 (value, tag pairs)
            static constexpr size_t offCoef = offPBot + (size_t)(NW + 1) * 64 * sizeof(float4);        // [3][TR][32] float4: cP, sX, sY
            static constexpr size_t total = offCoef + (size_t)3 * TR * 32 * sizeof(float4);
        };

        // everything of a thread that is fixed for the solve
        struct Ctx
        {
            int lane, wp;
            float C;
            float* hist;              // sample 0 of the thread's 4 cells in its row 0 (advanced by the step loop)
            size_t histRow;
            uint32_t ownRows;         // bit j: row j of the thread is recorded / stored (owned, inside the alloc grid, lanes 1..30)
            int sj, sk;               // pulse cell inside the thread's block (sj < 0: not here)
            const float* pulse;
            const float4* cP; const float4* sX; const float4* sY;       // this thread's coefficient float4s (row stride 32), shared memory
            int zero;
            int phase;                // kSyncFlow: number of rows this warp has published (the tag of the row its neighbours expect next)
            bool warpHasSource;       // some lane of this warp holds the pulse cell (warp-uniform: the other 17 warps branch around the injection)
        };

        // K (<= 4) time steps; the caller has published this warp's first vx row in sVxTop and synchronised
        // record sample (FDTD.cpp:226-231): the pressure of the thread's owned rows into the time-major history, streaming stores
        template <int R>
        __device__ __forceinline__ void recordSample(Ctx& X, const float (&p)[R][4])
        {
            #pragma unroll
            for (int j = 0; j < R; ++j)
                if ((X.ownRows >> j) & 1u)
                {
#ifdef PVC_RES_COPYREC
                    // experiment: store from copies, so that the next pressure update does not wait for the store to have read p
                    __stcs(reinterpret_cast<float4*>(X.hist + (size_t)j * X.histRow),
                           make_float4(opaqueCopy(p[j][0], X.zero), opaqueCopy(p[j][1], X.zero), opaqueCopy(p[j][2], X.zero), opaqueCopy(p[j][3], X.zero)));
#else
                    __stcs(reinterpret_cast<float4*>(X.hist + (size_t)j * X.histRow), make_float4(p[j][0], p[j][1], p[j][2], p[j][3]));
#endif
                }
            X.hist += kHistChunkDefault;
        }
        // inject (FDTD.cpp:234): pulse sample t into the source cell, if this thread holds it
        template <int R>
        __device__ __forceinline__ void injectSample(const Ctx& X, float (&p)[R][4], const int t)
        {
            if (X.sj >= 0)
            {
                // adding +0 to the three other cells of the row is exact (it can only turn -0 into +0)
                const float add = __ldg(X.pulse + t);
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
        }

        // K (<= 4) time steps; the caller has published this warp's first vx row in sVxTop and synchronised.  The record + inject of
        // the LAST step are left to the caller, which mails the strips first: the neighbours wait for those, nobody for the history.
        template <int NW, int R, int SYNC, bool GEN, bool TRACK>
        __device__ __forceinline__ void stepLoop(Ctx& X, const int t0, const int nsteps, float (&p)[R][4], float (&vx)[R][4], float (&vy)[R][4],
                                                 float4* sVxTopRaw, float4* sPBotRaw, uint32_t& activity)
        {
            const int lane = X.lane, wp = X.wp;
            if (SYNC == kSyncFlow)
            {
                // rows of the exchange arrays: [w][2][32] float4
                const float4* vxFromBelow = sVxTopRaw + (size_t)(wp + 1) * 64;
                const float4* pFromAbove = sPBotRaw + (size_t)wp * 64;
                float4* myVxTop = sVxTopRaw + (size_t)wp * 64;
                float4* myPBot = sPBotRaw + (size_t)(wp + 1) * 64;
                const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                #pragma unroll 1
                for (int step = 0; step < nsteps; ++step)
                {
                    // ---- pressure sub-step (FDTD.cpp:125-141): the rows that need only this warp's own vx first
                    pressureStep<R, GEN, 0, R - 1>(p, vx, vy, zero4, X.C, X.cP);
                    const float4 vxBelow = (wp + 1 < NW) ? flow::poll(vxFromBelow, lane, X.phase) : zero4;
                    pressureStep<R, GEN, R - 1, R>(p, vx, vy, vxBelow, X.C, X.cP);
                    X.phase += 1;
                    flow::publish(myPBot, lane, p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3], X.phase);
                    if (step + 1 < nsteps) recordSample<R>(X, p);
                    // ---- velocity sub-steps + edge overrides (FDTD.cpp:144-223): row 0 needs the p row of the warp above
                    velocityStep<R, GEN, 1, R>(p, vx, vy, zero4, X.C, X.sX, X.sY);
                    const float4 pAbove = (wp > 0) ? flow::poll(pFromAbove, lane, X.phase) : zero4;
                    velocityStep<R, GEN, 0, 1>(p, vx, vy, pAbove, X.C, X.sX, X.sY);
                    if (TRACK)
                    {
                        #pragma unroll
                        for (int j = 0; j < R; ++j)
                        {
                            activity |= __float_as_uint(p[j][0]) | __float_as_uint(p[j][1]);
                            activity |= __float_as_uint(p[j][2]) | __float_as_uint(p[j][3]);
                        }
                    }
                    if (step + 1 < nsteps)
                    {
                        if (X.warpHasSource) injectSample<R>(X, p, t0 + step);
                        X.phase += 1;
                        flow::publish(myVxTop, lane, vx[0][0], vx[0][1], vx[0][2], vx[0][3], X.phase);
                    }
                }
                return;
            }
            float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(sVxTopRaw);
            float4 (*sPBot)[32] = reinterpret_cast<float4 (*)[32]>(sPBotRaw);
            #pragma unroll 1
            for (int step = 0; step < nsteps; ++step)
            {
                // ---- pressure sub-step (FDTD.cpp:125-141)
                pressureStep<R, GEN>(p, vx, vy, sVxTop[wp + 1][lane], X.C, X.cP);
                sPBot[wp + 1][lane] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
                // ---- record sample t0 + step (FDTD.cpp:226-231) NOW: the velocity sub-steps do not touch p, so this is the value the
                //      reference records after them, and the stores get the whole velocity phase to read their registers and leave the SM
                //      (recorded after it, they were still queued when the next pressure update -- or the strips -- needed the path)
                if (step + 1 < nsteps) recordSample<R>(X, p);
                phaseSync<NW, SYNC>(wp);
                // ---- velocity sub-steps + edge overrides (FDTD.cpp:144-223)
                velocityStep<R, GEN>(p, vx, vy, sPBot[wp][lane], X.C, X.sX, X.sY);
                if (TRACK)
                {
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                    {
                        activity |= __float_as_uint(p[j][0]) | __float_as_uint(p[j][1]);
                        activity |= __float_as_uint(p[j][2]) | __float_as_uint(p[j][3]);
                    }
                }
                if (step + 1 < nsteps)
                {
                    // ---- inject (FDTD.cpp:234)
                    if (X.warpHasSource) injectSample<R>(X, p, t0 + step);
                    sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
                    phaseSync<NW, SYNC>(wp);            // also the write-after-read fence of sPBot
                }
            }
        }

        template <int NW, int R, int MINB, int SYNC>
        __global__ void __launch_bounds__(NW * 32, MINB)
        residentKernel(const Layout L, const Args A)
        {
            using SM = Smem<NW, R>;
            constexpr int TR = SM::TR;
            extern __shared__ __align__(128) unsigned char smemRaw[];
            float4* sVxTopRaw = reinterpret_cast<float4*>(smemRaw + SM::offVxTop);
            float4* sPBotRaw = reinterpret_cast<float4*>(smemRaw + SM::offPBot);
            float4 (*sVxTop)[32] = reinterpret_cast<float4 (*)[32]>(sVxTopRaw);                    // [w]   = vx of warp w's first row
            float4 (*sPBot)[32] = reinterpret_cast<float4 (*)[32]>(sPBotRaw);                      // [w+1] = p of warp w's last row
            float4* sCoef = reinterpret_cast<float4*>(smemRaw + SM::offCoef);

            const int lane = threadIdx.x & 31;
            // broadcast from lane 0: the compiler then KNOWS the warp index is warp-uniform (uniform registers, plain branches instead of
            // divergence guards around every shuffle / vote of the step loop)
            const int wp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
            const int tps = A.tilesPerSource;
            const int sLocal = blockIdx.x / tps;
            const int tile = blockIdx.x - sLocal * tps;
            const int s = A.s0 + sLocal;
            const int ty = tile / L.tiles_x, tx = tile - ty * L.tiles_x;
            const int rBase = ty * L.valid_rows - kTileK + wp * R;
            const int cBase = tx * kValidCols - kGuardCols + lane * 4;
            const size_t cell0 = (size_t)(rBase + kGuardRows) * L.pitch + (cBase + kGuardCols);
            const size_t src0 = (size_t)s * L.plane + cell0;

            if (SYNC == kSyncFlow)
            {
                // tag 0 = nothing published yet (the first published row carries tag 1)
                sVxTopRaw[(size_t)wp * 64 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); sVxTopRaw[(size_t)wp * 64 + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                sPBotRaw[(size_t)(wp + 1) * 64 + lane] = make_float4(0.f, 0.f, 0.f, 0.f); sPBotRaw[(size_t)(wp + 1) * 64 + 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            else if (wp == 0)
            {
                sVxTop[NW][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
                sPBot[0][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const bool general = A.mode[(size_t)tile * 32 + wp] != 0u;
            if (general)
            {
                // this thread's coefficient float4s, read back only by itself: no synchronisation needed
                #pragma unroll
                for (int f = 0; f < 3; ++f)
                {
                    const float* plane = (f == 0 ? A.cP : (f == 1 ? A.sX : A.sY)) + cell0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                        sCoef[((size_t)f * TR + wp * R + j) * 32 + lane] = __ldg(reinterpret_cast<const float4*>(plane + (size_t)j * L.pitch));
                }
            }
            __syncthreads();

            // ---- row / lane classes of this thread, fixed for the solve
            const bool up = ty > 0, down = ty + 1 < L.tiles_y, left = tx > 0, right = tx + 1 < L.tiles_x;
            const bool haloLane = lane == 0 || lane == 31;
            uint32_t ownRows = 0u, loadRows = 0u, sendRows = 0u;
            #pragma unroll
            for (int j = 0; j < R; ++j)
            {
                const int tr = wp * R + j, r = rBase + j;
                const bool top = tr < kTileK, bottom = tr >= TR - kTileK;
                // the padding row r == gx may be the first halo row of the last tile row (tiles cover gx rows): it depends only on the
                // owned row gx - 1 of the same step, so it is always current there and is recorded / stored with the owned rows
                const bool owned = (!top && !bottom && r < L.rows) || (bottom && r == L.gx);
                if (owned) ownRows |= 1u << j;
                // A halo cell is reloaded iff the tile that owns it exists; everything else outside the tile is guard band (zero for
                // ever).  The padding row in the bottom halo is LIVE (vx = p of the row above it): in the halo lanes it is refreshed
                // from the left / right tile, which keeps it current the same way, or its error would creep into the halo columns.
                const bool liveRow = bottom && r == L.gx;
                const bool rowOk = top ? up : (bottom ? (down || liveRow) : true);
                const bool colOk = (lane == 0) ? left : (lane == 31 ? right : true);
                if ((top || (bottom && !liveRow) || haloLane) && rowOk && colOk) loadRows |= 1u << j;
                // an owned cell is mailed iff it lies within 4 cells of an edge behind which a tile exists (corners go with the rows)
                const bool inRows = (!top && !bottom) || liveRow;
                const bool nearRow = !liveRow && ((tr < 2 * kTileK && up) || (tr >= TR - 2 * kTileK && down));
                const bool nearCol = (lane == 1 && left) || (lane == 30 && right);
                if (inRows && !haloLane && (nearRow || nearCol)) sendRows |= 1u << j;
            }
            if (haloLane || cBase >= L.cols) ownRows = 0u;

            Ctx X;
            X.lane = lane; X.wp = wp; X.C = A.courant;
            X.hist = A.hist + (size_t)s * L.hist_source + (ptrdiff_t)rBase * (ptrdiff_t)L.hist_row
                   + ((ptrdiff_t)(cBase >> 7) * L.T) * kHistChunkDefault + (cBase & 127);
            X.histRow = L.hist_row;
            X.ownRows = PVC_DBG(A, 0) ? 0u : ownRows;
            const SourceParams sp = A.src[s];
            {
                const int sj = sp.cell_r - rBase, sk = sp.cell_c - cBase;
                const bool hasSrc = (sj >= 0) && (sj < R) && (sk >= 0) && (sk < 4);
                X.sj = (hasSrc && !sp.dead) ? sj : -1; X.sk = sk;
                X.warpHasSource = __any_sync(0xffffffffu, X.sj >= 0);
            }
            X.pulse = A.pulse; X.zero = A.zero; X.phase = 0;
            X.cP = sCoef + (size_t)(wp * R) * 32 + lane;
            X.sX = X.cP + (size_t)TR * 32;
            X.sY = X.cP + (size_t)2 * TR * 32;

            float p[R][4], vx[R][4], vy[R][4];
            #pragma unroll
            for (int j = 0; j < R; ++j)
                #pragma unroll
                for (int k = 0; k < 4; ++k) { p[j][k] = 0.f; vx[j][k] = 0.f; vy[j][k] = 0.f; }
            if (A.loadState)
            {
                // A later chunk of a streamed solve (pvc_create_streamed) continues from what the previous chunk's final store -- or the
                // restored checkpoint -- holds in buffer 0: the owned cells of every tile, i.e. this tile's own cells AND its halo (the
                // neighbours' owned cells: what the mailbox would have delivered); whatever no tile owns is guard band or inert padding
                // and zero in the planes as it is in the registers of an unbroken solve.
                #pragma unroll
                for (int j = 0; j < R; ++j)
                {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(A.p0 + src0 + (size_t)j * L.pitch));
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(A.vx0 + src0 + (size_t)j * L.pitch));
                    const float4 c = __ldcg(reinterpret_cast<const float4*>(A.vy0 + src0 + (size_t)j * L.pitch));
                    p[j][0] = a.x; p[j][1] = a.y; p[j][2] = a.z; p[j][3] = a.w;
                    vx[j][0] = b.x; vx[j][1] = b.y; vx[j][2] = b.z; vx[j][3] = b.w;
                    vy[j][0] = c.x; vy[j][1] = c.y; vy[j][2] = c.z; vy[j][3] = c.w;
                }
            }

            int firstGen = kNeverActive;          // first pass in which this warp's block recorded anything but zeros
            #pragma unroll 1
            for (int g = 0; g < A.numGen; ++g)
            {
                const int t0 = g * kTileK;
                const int nsteps = min(kTileK, A.T - t0);
                PVC_STAMP(A, g, 0);
                if (g > 0 && loadRows && !PVC_DBG(A, 1))
                {
                    // ---- reload the halo ring from the mailbox: the words the neighbours wrote at the end of pass g - 1 carry tag g
                    const float4* q0 = A.xchg + (size_t)(g & (kSlots - 1)) * A.xchgSlot + src0;
                    const int tag = A.tagBase + g;
                    // The ring fetch IS the poll: every iteration loads all the words (eight 256-bit loads in flight per thread) and
                    // checks every tag, so the pass starts one L2 round trip after the last word has landed.  (With 128-bit loads --
                    // sixteen per thread and iteration -- polling ONE word first and fetching the ring in a second round trip was
                    // the faster scheme; with 256-bit loads it is the slower one: 1024^2 x 4 7.30 -> 7.10 ms per 1000 steps,
                    // 512^2 2.28 -> 2.14, profiles/r02_flow_exchange.txt.)
                    unsigned spins = 0;
                    while (true)
                    {
                        {
                            // The loads are volatile asm statements: a load followed by its use, word after word, would make the fetch one
                            // L2 round trip per word.  A thread that reloads ALL its rows (halo warps; lanes 0 / 31 of the others)
                            // overwrites its whole state, so the registers for all its words in flight are there: one round trip.
                            // Partial reloads (the live padding row) go row by row.
                            int bad = 0;
                            if (loadRows == (1u << R) - 1u)
                            {
                                float4 v[R][4];
                                #pragma unroll
                                for (int j = 0; j < R; ++j)
                                    #pragma unroll
                                    for (int k = 0; k < 4; k += 2)
                                    {
                                        const WordPair w = loadWordPair(q0 + (size_t)j * L.pitch + k);
                                        v[j][k] = w.a; v[j][k + 1] = w.b;
                                    }
                                #pragma unroll
                                for (int j = 0; j < R; ++j)
                                    #pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                    {
                                        p[j][k] = opaqueCopy(v[j][k].x, A.zero); vx[j][k] = opaqueCopy(v[j][k].y, A.zero); vy[j][k] = opaqueCopy(v[j][k].z, A.zero);
                                        bad |= __float_as_int(v[j][k].w) ^ tag;
                                    }
                            }
                            else
                            {
                                #pragma unroll
                                for (int j = 0; j < R; ++j)
                                    if ((loadRows >> j) & 1u)
                                    {
                                        const float4* q = q0 + (size_t)j * L.pitch;
                                        const WordPair w01 = loadWordPair(q), w23 = loadWordPair(q + 2);
                                        const float4 v0 = w01.a, v1 = w01.b, v2 = w23.a, v3 = w23.b;
                                        p[j][0] = opaqueCopy(v0.x, A.zero); vx[j][0] = opaqueCopy(v0.y, A.zero); vy[j][0] = opaqueCopy(v0.z, A.zero);
                                        p[j][1] = opaqueCopy(v1.x, A.zero); vx[j][1] = opaqueCopy(v1.y, A.zero); vy[j][1] = opaqueCopy(v1.z, A.zero);
                                        p[j][2] = opaqueCopy(v2.x, A.zero); vx[j][2] = opaqueCopy(v2.y, A.zero); vy[j][2] = opaqueCopy(v2.z, A.zero);
                                        p[j][3] = opaqueCopy(v3.x, A.zero); vx[j][3] = opaqueCopy(v3.y, A.zero); vy[j][3] = opaqueCopy(v3.z, A.zero);
                                        bad |= (__float_as_int(v0.w) ^ tag) | (__float_as_int(v1.w) ^ tag) | (__float_as_int(v2.w) ^ tag) | (__float_as_int(v3.w) ^ tag);
                                    }
                            }
                            if (bad == 0) break;
                        }
                        // not there yet (or a neighbour died): bounded, and a raised abort flag ends every later wait at once
                        if ((spins & 0x3fu) == 0u && (spins > (1u << 18) || *(volatile int*)A.abortFlag)) { atomicExch(A.abortFlag, 1); break; }
                        ++spins;
                        if (!PVC_DBG(A, 2)) __nanosleep(20);
                    }
                }
                PVC_STAMP_IF(wp == 0 && lane == 1, A, g, 7);
                __syncwarp();            // the lanes that polled rejoin the others before the (warp-aligned) barriers below
                PVC_STAMP(A, g, 1);
                if (SYNC == kSyncFlow)
                {
                    X.phase += 1;
                    flow::publish(sVxTopRaw + (size_t)wp * 64, lane, vx[0][0], vx[0][1], vx[0][2], vx[0][3], X.phase);
                }
                else
                {
                    sVxTop[wp][lane] = make_float4(vx[0][0], vx[0][1], vx[0][2], vx[0][3]);
                    phaseSync<NW, SYNC>(wp);
                }
                PVC_STAMP(A, g, 2);

                uint32_t activity = 0u;
                if (firstGen == kNeverActive)
                {
                    if (general) stepLoop<NW, R, SYNC, true, true>(X, t0, nsteps, p, vx, vy, sVxTopRaw, sPBotRaw, activity);
                    else stepLoop<NW, R, SYNC, false, true>(X, t0, nsteps, p, vx, vy, sVxTopRaw, sPBotRaw, activity);
                    const bool hot = ((activity & 0x7fffffffu) != 0u) && !haloLane;
                    if (__any_sync(0xffffffffu, hot)) firstGen = g;
                }
                else
                {
                    if (general) stepLoop<NW, R, SYNC, true, false>(X, t0, nsteps, p, vx, vy, sVxTopRaw, sPBotRaw, activity);
                    else stepLoop<NW, R, SYNC, false, false>(X, t0, nsteps, p, vx, vy, sVxTopRaw, sPBotRaw, activity);
                }

                PVC_STAMP(A, g, 3);
                const bool lastPass = g + 1 == A.numGen;
                // the last step's record + inject: the thread that holds the source cell (one per tile, its mailed value includes the
                // injection) does them now, every other thread after the strips have left
                const int tLast = t0 + nsteps - 1;
                if (X.sj >= 0) { recordSample<R>(X, p); injectSample<R>(X, p, tLast); }
                if (!lastPass)
                {
                    // ---- mail the cells the neighbours need for pass g + 1
                    if (sendRows)
                    {
                        float4* q0 = A.xchg + (size_t)((g + 1) & (kSlots - 1)) * A.xchgSlot + src0;
                        const int tag = A.tagBase + g + 1;
                        #pragma unroll
                        for (int j = 0; j < R; ++j)
                            if ((sendRows >> j) & 1u)
                            {
                                #pragma unroll
                                for (int k = 0; k < 4; k += 2)
                                    storeWordPair(q0 + (size_t)j * L.pitch + k, opaqueCopy(p[j][k], A.zero), opaqueCopy(vx[j][k], A.zero), opaqueCopy(vy[j][k], A.zero),
                                                  opaqueCopy(p[j][k + 1], A.zero), opaqueCopy(vx[j][k + 1], A.zero), opaqueCopy(vy[j][k + 1], A.zero), tag);
                            }
                    }
                }
                else if (ownRows)
                {
                    // ---- the final state of every owned cell into the state planes (Grid's m_grid after the last step)
                    if (sp.dead && A.T == t0 + nsteps && A.finalPass)
                    {
                        // the reference injects the last pulse sample even into a wall / padding cell, where nothing ever reads it
                        // again (FDTD.cpp:234): it only shows in the final state
                        const int sj = sp.cell_r - rBase, sk = sp.cell_c - cBase;
                        if (sj >= 0 && sj < R && sk >= 0 && sk < 4)
                        {
                            const float add = __ldg(A.pulse + A.T - 1);
                            #pragma unroll
                            for (int j = 0; j < R; ++j)
                                #pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (j == sj && k == sk) p[j][k] = __fadd_rn(p[j][k], add);
                        }
                    }
                    float* gp = ((g & 1) ? A.p0 : A.p1) + src0;
                    float* gx = ((g & 1) ? A.vx0 : A.vx1) + src0;
                    float* gy = ((g & 1) ? A.vy0 : A.vy1) + src0;
                    #pragma unroll
                    for (int j = 0; j < R; ++j)
                        if ((ownRows >> j) & 1u)
                        {
                            __stcg(reinterpret_cast<float4*>(gp + (size_t)j * L.pitch), make_float4(p[j][0], p[j][1], p[j][2], p[j][3]));
                            __stcg(reinterpret_cast<float4*>(gx + (size_t)j * L.pitch), make_float4(vx[j][0], vx[j][1], vx[j][2], vx[j][3]));
                            __stcg(reinterpret_cast<float4*>(gy + (size_t)j * L.pitch), make_float4(vy[j][0], vy[j][1], vy[j][2], vy[j][3]));
                        }
                }
                PVC_STAMP_IF(wp == NW - 2 && lane == 1, A, g, 5);
                if (X.sj < 0) recordSample<R>(X, p);
                PVC_STAMP(A, g, 4);
            }
            if (lane == 0 && firstGen != kNeverActive && A.firstActive)
                A.firstActive[((size_t)s * tps + tile) * 32 + wp] = firstGen;
        }

        // ---- linear-form coefficient planes from the wall plane (see the header) ------------------------------------
        __global__ void buildLinearKernel(const Layout L, const float courant, const float* __restrict__ w,
                                          float* __restrict__ cP, float* __restrict__ sX, float* __restrict__ sY)
        {
            const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= L.plane) return;
            const int r = (int)(i / L.pitch) - kGuardRows, c = (int)(i % L.pitch) - kGuardCols;
            const bool inRows = (r >= 0) && (r < L.gx), inCols = (c >= 0) && (c < L.gy);
            const float wSelf = w[i];
            const bool air = inRows && inCols && __float_as_uint(wSelf) == kAirBits;
            float x = 0.f, y = 0.f;
            if (inCols && r >= 0 && r <= L.gx)
            {
                if (r == 0 || r == L.gx) x = 1.f;                        // vx = -p (FDTD.cpp:208) / vx = p_up (FDTD.cpp:209)
                else
                {
                    const float wUp = w[i - L.pitch];
                    const bool airUp = __float_as_uint(wUp) == kAirBits;
                    x = air ? (airUp ? -courant : wUp) : (airUp ? wSelf : 0.f);
                }
            }
            if (inRows && c >= 0 && c <= L.gy)
            {
                if (c == 0 || c == L.gy) y = 1.f;                        // vy = -p (FDTD.cpp:220) / vy = p_left (FDTD.cpp:221)
                else
                {
                    const float wLeft = w[i - 1];
                    const bool airLeft = __float_as_uint(wLeft) == kAirBits;
                    y = air ? (airLeft ? -courant : wLeft) : (airLeft ? wSelf : 0.f);
                }
            }
            cP[i] = air ? courant : 0.f;
            sX[i] = x;
            sY[i] = y;
        }

        __global__ void markDeadSourcesKernel(const Layout L, const float* __restrict__ w, SourceParams* __restrict__ src, int n)
        {
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= n) return;
            const int r = src[i].cell_r, c = src[i].cell_c;
            const bool interior = r >= 0 && c >= 0 && r < L.gx && c < L.gy;
            src[i].dead = (interior && __float_as_uint(w[cellIndex(L, r, c)]) == kAirBits) ? 0 : 1;
        }

        template <int NW, int R, int MINB, int SYNC>
        static int capacity(int device)
        {
            static int cached[64] = {};
            int& c = cached[device & 63];
            if (c) return c;
            const size_t smem = Smem<NW, R>::total;
            if (cudaFuncSetAttribute(residentKernel<NW, R, MINB, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
            int perSm = 0, sms = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, residentKernel<NW, R, MINB, SYNC>, NW * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) { cudaGetLastError(); return 0; }
            c = perSm * sms;
            return c;
        }

        template <int NW, int R, int MINB, int SYNC>
        static int launch(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
        {
            const Layout& L = s->L;
            // a chunk of a streamed solve (pvc_create_streamed): samples t0 .. t1 - 1 of the response, recorded as samples 0 .. t1 - t0 - 1
            // of the chunk-sized history, continued from the state planes when t0 > 0
            if ((t0 != 0 && !s->chunkT) || s->cur != 0) { setError("resident step kernel: must start at step 0"); return PVC_ERR_INVALID; }
            if (!hist || !s->lin[0] || !s->resXchg) { setError("resident step kernel: history / coefficient planes / mailbox missing"); return PVC_ERR_INVALID; }
            if (L.hist_chunk != kHistChunkDefault || L.tile_rows != NW * R) { setError("resident step kernel: layout does not match the variant"); return PVC_ERR_INVALID; }
            const int tps = L.tiles_x * L.tiles_y;
            const int cap = capacity<NW, R, MINB, SYNC>(s->device);
            if (cap < tps) { setError("resident step kernel: %d tiles per source exceed the %d co-resident CTAs of this device", tps, cap); return PVC_ERR_INVALID; }
            const int perLaunch = cap / tps;
            const int gens = (t1 - t0 + kTileK - 1) / kTileK;
            if (gens >= 65536) { setError("resident step kernel: %d passes exceed the 16-bit pass field of the mailbox tag", gens); return PVC_ERR_INVALID; }
            // every solve gets its own tag epoch, so words left in the mailbox by earlier solves can never match; when the 15-bit
            // epoch wraps the mailbox is cleared once
            s->resEpoch = (s->resEpoch + 1) & 0x7fff;
            if (s->resEpoch == 0)
            {
                s->resEpoch = 1;
                cudaMemsetAsync(s->resXchg, 0, sizeof(float4) * (size_t)res::kSlots * s->cfg.max_sources * L.plane, s->stream);
            }
            cudaMemsetAsync(s->tileCounters, 0, sizeof(int), s->stream);            // slot 0 of the pool is the abort flag
            Args A;
            A.p0 = s->state[0][0]; A.vx0 = s->state[0][1]; A.vy0 = s->state[0][2];
            A.p1 = s->state[1][0]; A.vx1 = s->state[1][1]; A.vy1 = s->state[1][2];
            A.hist = hist; A.mode = s->slowMask; A.cP = s->lin[0]; A.sX = s->lin[1]; A.sY = s->lin[2];
            A.firstActive = (s->chunkT && !s->hintsValid) ? nullptr : s->firstActive;      // streamed solve: hints per chunk, forward sweep only
            A.src = s->src; A.pulse = s->pulse + t0;
            A.loadState = (s->chunkT && t0 > 0) ? 1 : 0; A.finalPass = s->chunkT ? s->finalPass : 1;
            A.xchg = reinterpret_cast<float4*>(s->resXchg); A.xchgSlot = (size_t)s->cfg.max_sources * L.plane; A.tagBase = s->resEpoch << 16;
            A.abortFlag = s->tileCounters; A.zero = 0;
            A.tilesPerSource = tps; A.numGen = gens; A.T = t1 - t0; A.courant = s->cfg.courant;
#ifdef PVC_TUNING
            { static const char* d = getenv("PVC_RES_DEBUG"); A.dbg = d ? atoi(d) : 0; }
            A.trace = nullptr;
            static unsigned long long* traceBuf = nullptr;
            const size_t traceWords = (size_t)cap * kTracePasses * kTraceSlots;
            if (getenv("PVC_RES_TRACE"))
            {
                if (!traceBuf) cudaMalloc(&traceBuf, traceWords * sizeof(unsigned long long));
                cudaMemsetAsync(traceBuf, 0, traceWords * sizeof(unsigned long long), s->stream);
                A.trace = traceBuf;
            }
#endif
            Layout Lc = L;
            for (int s0 = 0; s0 < nsrc; s0 += perLaunch)
            {
                A.s0 = s0; A.nsrc = (nsrc - s0 < perLaunch) ? (nsrc - s0) : perLaunch;
                void* params[2] = { &Lc, &A };
                // cooperative launch: fails instead of deadlocking if the CTAs could not all be resident
                const cudaError_t e = cudaLaunchCooperativeKernel((const void*)residentKernel<NW, R, MINB, SYNC>, dim3((unsigned)(tps * A.nsrc)), dim3(NW * 32), params,
                                                                  Smem<NW, R>::total, s->stream);
                if (e != cudaSuccess) { setError("resident step kernel launch (%d CTAs): %s", tps * A.nsrc, cudaGetErrorString(e)); return PVC_ERR_CUDA; }
                *launches += 1;
            }
#ifdef PVC_TUNING
            if (A.trace && gens >= kTracePass0 + kTracePasses)
            {
                cudaStreamSynchronize(s->stream);
                std::vector<unsigned long long> h(traceWords);
                cudaMemcpy(h.data(), traceBuf, traceWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                const int ctas = tps * A.nsrc;
                double sum[5] = {}; long cnt = 0;
                std::vector<std::pair<double, int>> busy;      // per CTA: mean (steps + mail) and its index
                std::vector<uint32_t> modes((size_t)tps * 32);
                cudaMemcpy(modes.data(), s->slowMask, sizeof(uint32_t) * modes.size(), cudaMemcpyDeviceToHost);
                for (int c = 0; c < ctas; ++c)
                {
                    double b = 0; long n = 0;
                    for (int k = 1; k + 1 < kTracePasses; ++k)
                    {
                        const unsigned long long* a = h.data() + ((size_t)c * kTracePasses + k) * kTraceSlots;
                        const unsigned long long* nx = a + kTraceSlots;
                        if (!a[0] || !a[1] || !a[2] || !a[3] || !a[4] || !nx[0]) continue;
                        sum[0] += (double)a[1] - (double)a[0];       // halo reload (the poll)
                        sum[1] += (double)a[2] - (double)a[1];       // first exchange row + sync
                        sum[2] += (double)a[3] - (double)a[2];       // 4 steps
                        sum[3] += (double)a[4] - (double)a[3];       // strips mailed
                        sum[4] += (double)nx[0] - (double)a[0];      // pass period
                        b += (double)a[4] - (double)a[1]; ++n;
                        ++cnt;
                    }
                    if (n) busy.push_back(std::make_pair(b / n, c));
                }
                {
                    // hand-over latency seen by thread 0 (warp 0, lane 0: its halo words come from the upper-left tile, or the left one in
                    // the top tile row): its reload of pass k completes this long after that tile finished the steps of pass k - 1
                    double lat = 0; long nl = 0;
                    for (int c = 0; c < ctas; ++c)
                    {
                        const int tile = c % tps, tx = tile % L.tiles_x, ty = tile / L.tiles_x;
                        if (tx == 0) continue;
                        const int nb = (c - tile) + (ty > 0 ? (ty - 1) : ty) * L.tiles_x + (tx - 1);
                        for (int k = 2; k + 1 < kTracePasses; ++k)
                        {
                            const unsigned long long* a = h.data() + ((size_t)c * kTracePasses + k) * kTraceSlots;
                            const unsigned long long* b = h.data() + ((size_t)nb * kTracePasses + (k - 1)) * kTraceSlots;
                            if (!a[1] || !b[3]) continue;
                            lat += (double)a[1] - (double)b[3]; ++nl;
                        }
                    }
                    if (nl) fprintf(stderr, "[res trace] hand-over: reload of pass k done %.0f ns after the writer tile finished the steps of pass k-1 (n=%ld)\n", lat / nl, nl);
                    // vertical pairs, thread to thread: (warp NW-2, lane 1) of the upper tile mails the words that (warp 0, lane 1) of this
                    // tile polls: mail issued -> first word seen -> all 16 words loaded; and how long that reader had been polling
                    double seen = 0, loaded = 0, polled = 0, skew = 0; long nv = 0;
                    for (int c = 0; c < ctas; ++c)
                    {
                        const int tile = c % tps, ty = tile / L.tiles_x;
                        if (ty == 0) continue;
                        const int nb = c - L.tiles_x;
                        for (int k = 2; k + 1 < kTracePasses; ++k)
                        {
                            const unsigned long long* a = h.data() + ((size_t)c * kTracePasses + k) * kTraceSlots;
                            const unsigned long long* b = h.data() + ((size_t)nb * kTracePasses + (k - 1)) * kTraceSlots;
                            if (!a[6] || !a[7] || !b[5] || !a[0] || !b[3]) continue;
                            seen += (double)a[6] - (double)b[5]; loaded += (double)a[7] - (double)a[6]; polled += (double)a[6] - (double)a[0];
                            skew += (double)b[5] - (double)b[3]; ++nv;
                        }
                    }
                    if (nv) fprintf(stderr, "[res trace] vertical pairs (n=%ld): writer thread mails %.0f ns after its tile's thread 0 finished the steps; word seen %.0f ns after the mail; ring loaded %.0f ns later; the reader had polled for %.0f ns\n",
                                    nv, skew / nv, seen / nv, loaded / nv, polled / nv);
                }
                if (cnt) fprintf(stderr, "[res trace] NW=%d ctas=%d n=%ld  reload %.0f  exchange+sync %.0f  steps %.0f  mail %.0f | period %.0f ns (thread 0 of every CTA)\n",
                                 NW, ctas, cnt, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, sum[4] / cnt);
                if (!busy.empty())
                {
                    std::sort(busy.begin(), busy.end());
                    fprintf(stderr, "[res trace] busy time per pass (exchange+steps+mail) over CTAs: min %.0f  median %.0f  max %.0f ns; slowest:", busy.front().first, busy[busy.size() / 2].first, busy.back().first);
                    for (size_t k = 0; k < 6 && k < busy.size(); ++k)
                    {
                        const int c = busy[busy.size() - 1 - k].second, tile = c % tps;
                        int gen = 0; for (int w = 0; w < NW; ++w) gen += modes[(size_t)tile * 32 + w] != 0u;
                        fprintf(stderr, " tile(%d,%d) %.0f ns %d/%d general;", tile % L.tiles_x, tile / L.tiles_x, busy[busy.size() - 1 - k].first, gen, NW);
                    }
                    fprintf(stderr, " fastest:");
                    for (size_t k = 0; k < 3 && k < busy.size(); ++k)
                    {
                        const int c = busy[k].second, tile = c % tps;
                        int gen = 0; for (int w = 0; w < NW; ++w) gen += modes[(size_t)tile * 32 + w] != 0u;
                        fprintf(stderr, " tile(%d,%d) %.0f ns %d/%d general;", tile % L.tiles_x, tile / L.tiles_x, busy[k].first, gen, NW);
                    }
                    fprintf(stderr, "\n");
                }
            }
#endif
            s->cur = gens & 1;
            s->checkAbort = 1;
            return PVC_OK;
        }
    }

    // tilings, all 4 rows per warp unless noted.  Barrier-free row exchange, one CTA per SM: 72 / 70 / 71 / 63 / 65 / 64 = 10 / 12 / 14 / 16 /
    // 18 / 20 warps, 66 = 16 warps x 5 rows; two CTAs per SM: 69 = 8 warps.  CTA barrier per sub-step: 67 = 4 warps (8 owned rows: the
    // reference's own 70^2 .. 191^2 contract and everything else whose 8-row tiles fit the GPU at most two to an SM -- a pass there
    // is pure latency and more SMs help; with the barrier-free exchange these tiles are 7 % SLOWER, profiles/r02_small_tilings.txt).  Superseded, -DPVC_ALL_VARIANTS only:
    // 60..62 = 8 / 10 / 12 warps with the CTA barrier, 68 = 10 warps, two CTAs per SM.
#ifdef PVC_ALL_VARIANTS
    #define PVC_RES_OLD_VARIANTS(X) X(60, 8, 4, 2, res::kSyncCta) X(61, 10, 4, 2, res::kSyncCta) X(62, 12, 4, 2, res::kSyncCta) X(68, 10, 4, 2, res::kSyncFlow)
#else
    #define PVC_RES_OLD_VARIANTS(X)
#endif
    #define PVC_RES_VARIANTS(X) \
        PVC_RES_OLD_VARIANTS(X) \
        X(63, 16, 4, 1, res::kSyncFlow) X(64, 20, 4, 1, res::kSyncFlow) X(65, 18, 4, 1, res::kSyncFlow) X(66, 16, 5, 1, res::kSyncFlow) \
        X(67, 4, 4, 4, res::kSyncCta) X(69, 8, 4, 2, res::kSyncFlow) \
        X(70, 12, 4, 1, res::kSyncFlow) X(71, 14, 4, 1, res::kSyncFlow) X(72, 10, 4, 1, res::kSyncFlow)
    int launchResidentSteps(pvc_solver* s, int variant, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        switch (variant)
        {
    #define X(v, nw, r, minb, sync) case v: return res::launch<nw, r, minb, sync>(s, nsrc, t0, t1, hist, launches);
            PVC_RES_VARIANTS(X)
    #undef X
            default: setError("resident step kernel: unknown variant %d", variant); return PVC_ERR_INVALID;
        }
    }

    int residentCapacity(int variant, int device)
    {
        switch (variant)
        {
    #define X(v, nw, r, minb, sync) case v: return res::capacity<nw, r, minb, sync>(device);
            PVC_RES_VARIANTS(X)
    #undef X
            default: return 0;
        }
    }

    int rebuildResidentDescriptors(pvc_solver* s, int variant)
    {
        (void)variant;
        if (!s->lin[0]) { setError("resident step kernel: coefficient planes not allocated"); return PVC_ERR_INVALID; }
        res::buildLinearKernel<<<(unsigned)((s->L.plane + 255) / 256), 256, 0, s->stream>>>(s->L, s->cfg.courant, s->w, s->lin[0], s->lin[1], s->lin[2]);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("linear coefficient launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    int markDeadSources(pvc_solver* s, int n)
    {
        res::markDeadSourcesKernel<<<(n + 63) / 64, 64, 0, s->stream>>>(s->L, s->w, s->src, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("dead-source launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }
}

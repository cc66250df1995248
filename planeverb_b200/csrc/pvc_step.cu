// pvc_step.cu -- two-launch baseline time step (step_kernel = 1).
//
// One launch for the pressure sub-step and one for both velocity sub-steps, in place on the current
// state planes, exactly the sub-step order of Grid::GenerateResponseCPU
// (ProjectPlaneverb/src/FDTD/FDTD.cpp:122-235): pressure -> vx, vy -> grid-edge overrides -> record ->
// inject.  It exists as the simple, obviously-correct device path: the fused temporally blocked
// kernel (pvc_step_fused.cu) must reproduce its planes bit for bit at sizes the CPU oracle cannot
// reach.  All arithmetic uses explicit round-to-nearest intrinsics so no multiply-add is ever
// contracted (the reference build is strict fp32).
#include "pvc_internal.h"

namespace pvc
{
    __device__ __forceinline__ bool isAir(float w) { return __float_as_uint(w) == kAirBits; }

    // general velocity rule for a cell and its "previous" neighbour (FDTD.cpp:149-168 in branch form)
    __device__ __forceinline__ float velocityRule(float v, float pThis, float pPrev, float wThis, float wPrev, float courant)
    {
        const bool aThis = isAir(wThis), aPrev = isAir(wPrev);
        if (aThis && aPrev) return __fsub_rn(v, __fmul_rn(courant, __fsub_rn(pThis, pPrev)));
        if (!aThis && aPrev) return __fmul_rn(wThis, pPrev);
        if (aThis && !aPrev) return -__fmul_rn(wPrev, pThis);
        return 0.f;
    }

    // pressure sub-step of sample t for 4 consecutive cells per thread, then record, with the previous
    // sample's injection folded in front (p[li] += pulse[t-1], FDTD.cpp:234)
    __global__ void __launch_bounds__(256)
    baselinePressureKernel(Layout L, float* __restrict__ p, const float* __restrict__ vx, const float* __restrict__ vy,
                           const float* __restrict__ w, float* __restrict__ hist,
                           const SourceParams* __restrict__ src, const float* __restrict__ pulse, int t, float courant)
    {
        const int q = blockIdx.x * blockDim.x + threadIdx.x;       // column quad
        const int r = blockIdx.y * blockDim.y + threadIdx.y;
        const int c = q * 4;
        if (c >= L.cols || r >= L.rows) return;
        const int s = blockIdx.z;
        const size_t base = (size_t)s * L.plane;
        const size_t i = cellIndex(L, r, c);

        float4 p4 = *reinterpret_cast<const float4*>(p + base + i);
        const float4 vx4 = *reinterpret_cast<const float4*>(vx + base + i);
        const float4 vxd = *reinterpret_cast<const float4*>(vx + base + i + L.pitch);
        const float4 vy4 = *reinterpret_cast<const float4*>(vy + base + i);
        const float vyn = vy[base + i + 4];
        const float4 w4 = *reinterpret_cast<const float4*>(w + i);

        float pv[4] = { p4.x, p4.y, p4.z, p4.w };
        const float vxa[4] = { vx4.x, vx4.y, vx4.z, vx4.w };
        const float vda[4] = { vxd.x, vxd.y, vxd.z, vxd.w };
        const float vya[5] = { vy4.x, vy4.y, vy4.z, vy4.w, vyn };
        const float wa[4] = { w4.x, w4.y, w4.z, w4.w };

        if (t > 0)
        {
            const SourceParams sp = src[s];
            if (sp.cell_r == r && sp.cell_c >= c && sp.cell_c < c + 4)
            {
                const float add = pulse[t - 1];
                #pragma unroll
                for (int k = 0; k < 4; ++k) if (sp.cell_c == c + k) pv[k] = __fadd_rn(pv[k], add);
            }
        }
        #pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const float div = __fadd_rn(__fsub_rn(vda[k], vxa[k]), __fsub_rn(vya[k + 1], vya[k]));
            pv[k] = isAir(wa[k]) ? __fsub_rn(pv[k], __fmul_rn(courant, div)) : 0.f;
        }
        p4 = make_float4(pv[0], pv[1], pv[2], pv[3]);
        *reinterpret_cast<float4*>(p + base + i) = p4;
        if (hist)
            __stcs(reinterpret_cast<float4*>(hist + (size_t)s * L.hist_source + histCell(L, r, c) + (size_t)t * L.hist_chunk), p4);
    }

    // both velocity sub-steps + the grid-edge absorbing overrides (FDTD.cpp:144-223)
    __global__ void __launch_bounds__(256)
    baselineVelocityKernel(Layout L, const float* __restrict__ p, float* __restrict__ vx, float* __restrict__ vy,
                           const float* __restrict__ w, float courant)
    {
        const int q = blockIdx.x * blockDim.x + threadIdx.x;
        const int r = blockIdx.y * blockDim.y + threadIdx.y;
        const int c = q * 4;
        if (c >= L.cols || r >= L.rows) return;
        const size_t base = (size_t)blockIdx.z * L.plane;
        const size_t i = cellIndex(L, r, c);

        const float4 p4 = *reinterpret_cast<const float4*>(p + base + i);
        const float4 pu4 = *reinterpret_cast<const float4*>(p + base + i - L.pitch);
        const float pl = p[base + i - 1];
        const float4 vx4 = *reinterpret_cast<const float4*>(vx + base + i);
        const float4 vy4 = *reinterpret_cast<const float4*>(vy + base + i);
        const float4 w4 = *reinterpret_cast<const float4*>(w + i);
        const float4 wu4 = *reinterpret_cast<const float4*>(w + i - L.pitch);
        const float wl = w[i - 1];

        const float pa[5] = { pl, p4.x, p4.y, p4.z, p4.w };
        const float pua[4] = { pu4.x, pu4.y, pu4.z, pu4.w };
        const float wa[5] = { wl, w4.x, w4.y, w4.z, w4.w };
        const float wua[4] = { wu4.x, wu4.y, wu4.z, wu4.w };
        float vxa[4] = { vx4.x, vx4.y, vx4.z, vx4.w };
        float vya[4] = { vy4.x, vy4.y, vy4.z, vy4.w };

        #pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const int cc = c + k;
            const float pt = pa[k + 1];
            float nx, ny;
            if (cc >= L.gy) nx = 0.f;                       // padding column: never driven
            else if (r == 0) nx = -pt;                      // FDTD.cpp:208
            else if (r == L.gx) nx = pua[k];                // FDTD.cpp:209
            else nx = velocityRule(vxa[k], pt, pua[k], wa[k + 1], wua[k], courant);
            if (r >= L.gx) ny = 0.f;                        // padding row
            else if (cc == 0) ny = -pt;                     // FDTD.cpp:220
            else if (cc == L.gy) ny = pa[k];                // FDTD.cpp:221
            else if (cc > L.gy) ny = 0.f;
            else ny = velocityRule(vya[k], pt, pa[k], wa[k + 1], wa[k], courant);
            vxa[k] = nx; vya[k] = ny;
        }
        *reinterpret_cast<float4*>(vx + base + i) = make_float4(vxa[0], vxa[1], vxa[2], vxa[3]);
        *reinterpret_cast<float4*>(vy + base + i) = make_float4(vya[0], vya[1], vya[2], vya[3]);
    }

    // steps t0..t1-1 on state[s->cur]; hist (may be null) receives one plane per step
    int launchBaselineSteps(pvc_solver* s, int nsrc, int t0, int t1, float* hist, int* launches)
    {
        const Layout& L = s->L;
        const int quads = (L.cols + 3) / 4;
        dim3 block(32, 8, 1);
        dim3 grid((quads + 31) / 32, (L.rows + 7) / 8, nsrc);
        float** st = s->state[s->cur];
        for (int t = t0; t < t1; ++t)
        {
            baselinePressureKernel<<<grid, block, 0, s->stream>>>(L, st[0], st[1], st[2], s->w, hist,
                                                                 s->src, s->pulse, t, s->cfg.courant);
            baselineVelocityKernel<<<grid, block, 0, s->stream>>>(L, st[0], st[1], st[2], s->w, s->cfg.courant);
            *launches += 2;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("baseline step launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }
}

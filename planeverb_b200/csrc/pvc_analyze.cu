// pvc_analyze.cu -- per-cell impulse-response analyzer on the device.
//
// Restates Analyzer::EncodeResponse (ProjectPlaneverb/src/DSP/Analyzer.cpp:139-328) and
// Analyzer::EncodeListenerDirection (:340-431) over the pressure history the step kernels record
// (one fp32 plane per sample instead of the reference's 16-byte Cell per sample per cell).
//
// One thread per interior cell.  Every sum runs in the reference's order in fp32 without
// contraction, so the results equal the CPU reference's up to libm differences in log10f/powf:
//   * causal part (onset, Edry, flux): ascending t from sample 0 (Analyzer.cpp:146-195).  The flux
//     needs vx, vy of the cell, which are rebuilt on the fly from the cell's and its up/left
//     neighbours' recorded pressures with the solver's own velocity rules (FDTD.cpp:144-223) -- an
//     exact recurrence because a velocity only ever depends on its own previous value and on the
//     recorded pressures;
//   * wet energy: ascending t over its window (Analyzer.cpp:235-247);
//   * RT60: backward Schroeder integral, descending t, with the running log10 and the two regression
//     sums of Analyzer.cpp:303-319 -- this anti-causal pass is why a pressure history exists at all.
// A block is the cells of one history strip (128 columns): its 4 warps walk one contiguous 512-byte-per-sample stream.
#include <float.h>
#include "pvc_internal.h"

namespace pvc
{
    __device__ __forceinline__ bool isAirA(float w) { return __float_as_uint(w) == kAirBits; }

    constexpr float kAudibleThreshold = 0.00000316f;   // PvTypes.h:89
    constexpr float kGainThreshold = 0.891251f;        // PvTypes.h:99
    constexpr float kDelayClose = 5.f;                 // PvTypes.h:100
    constexpr float kSpeedOfSoundA = 343.21f;          // PvTypes.h:85

    // ---- log10f exactly as the reference's libm computes it -------------------------------------------------
    // RT60 is a regression over y_i = 10*log10f(E_i) accumulated in fp32 (Analyzer.cpp:309-319); with few
    // regression points a 1-ulp difference in y_i moves RT60 by more than 1e-4, so the device reproduces the
    // libm the oracle links (glibc 2.39, x86-64) bit for bit instead of calling CUDA's log10f (<= 2 ulp):
    //   log10f(x) = (k*log10_2lo + ivln10*logf(m)) + k*log10_2hi   with x = m*2^k, m in [~0.7, 1.4)   (fdlibm e_log10f)
    //   logf(m)   = the table-driven double-precision kernel of glibc's e_logf.c (16-entry {1/c, log c} table,
    //               degree-3 polynomial in r = m/c - 1), rounded once to float.
    // tests/test_oracle.py::test_device_log10f_recipe_matches_libm checks this recipe against libm on the CPU.
    struct LogfEntry { double invc, logc; };
    __device__ const LogfEntry kLogfTable[16] = {
        { 0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2 }, { 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2 },
        { 0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2 }, { 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3 },
        { 0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3 }, { 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3 },
        { 0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4 }, { 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4 },
        { 0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5 }, { 0x1.0000000000000p+0, 0x0.0p+0 },
        { 0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5 }, { 0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4 },
        { 0x1.b2036576afce6p-1, 0x1.526e57720db08p-3 }, { 0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3 },
        { 0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2 }, { 0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2 } };

    // logf for a normal positive float (the only inputs log10fExact hands it)
    __device__ __forceinline__ float logfExact(float x, const LogfEntry* __restrict__ tab)
    {
        const uint32_t ix = __float_as_uint(x);
        if (ix == 0x3f800000u) return 0.f;
        const uint32_t tmp = ix - 0x3f330000u;
        const int i = (tmp >> 19) & 15;
        const int k = (int)tmp >> 23;
        const uint32_t iz = ix - (tmp & 0xff800000u);
        const LogfEntry e = tab[i];
        const double z = (double)__uint_as_float(iz);
        const double r = __dsub_rn(__dmul_rn(z, e.invc), 1.0);
        const double y0 = __dadd_rn(e.logc, __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
        const double r2 = __dmul_rn(r, r);
        double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
        y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
        y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
        return (float)y;
    }

    __device__ __forceinline__ float log10fExact(float x, const LogfEntry* __restrict__ tab)
    {
        int hx = __float_as_int(x);
        int k = 0;
        if (hx < 0x00800000)
        {
            if ((hx & 0x7fffffff) == 0) return __int_as_float(0xff800000);       // log10(0) = -inf
            if (hx < 0) return __int_as_float(0x7fc00000);                       // negative: NaN
            k -= 25; x = __fmul_rn(x, 3.3554432000e+07f);                        // subnormal: scale by 2^25
            hx = __float_as_int(x);
        }
        if (hx >= 0x7f800000) return __fadd_rn(x, x);
        k += (hx >> 23) - 127;
        const int i = (int)((unsigned)k >> 31);
        hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
        const float y = (float)(k + i);
        const float m = __int_as_float(hx);
        const float z = __fadd_rn(__fmul_rn(y, 7.9034151668e-07f), __fmul_rn(4.3429449201e-01f, logfExact(m, tab)));
        return __fadd_rn(z, __fmul_rn(y, 3.0102920532e-01f));
    }

    // 10*log10f(E) of the Schroeder curve (Analyzer.cpp:313).  -DPVC_FAST_LOG10 swaps in the hardware base-2
    // logarithm (MUFU.LG2; a few ulp off, RT60 then agrees only to ~1e-5 on well-conditioned cells).
    __device__ __forceinline__ float decibels(float e, const LogfEntry* __restrict__ tab)
    {
    #ifdef PVC_FAST_LOG10
        return __fmul_rn(__log2f(e), 3.0102999566398120f);
    #else
        return __fmul_rn(10.f, log10fExact(e, tab));
    #endif
    }

    // The same value as decibels() for a NORMAL, positive, finite e (the caller checks a whole batch with one
    // predicate), as straight-line code with the logf kernel in a cheaper, EXHAUSTIVELY verified form
    // (tools/micro/logf_fma_recipe.c runs all 2^24 floats of [0.5, 2), the only inputs log10f hands to logf, against
    // the host libm: zero mismatches; tests/test_oracle.py::test_device_logf_fma_form_is_exhaustively_exact):
    //   * a 33-entry table indexed by (bits(m) - 0x3f330000) >> 19 (arithmetic, -7..25) that holds, per logf interval AND
    //     exponent k in {-1, 0, 1}, { invc * 2^-k, logc + k*ln2 } (buildLogfTable33: the reference's own y0 = logc + k*ln2,
    //     one rounding, and an exact power-of-two scaling), so neither the reduced mantissa z = m * 2^-k nor k is formed;
    //   * r = fma(m, invc', -1) and the whole of logc + r + r^2 (A2 + A1 r + A0 r^2) as ONE Horner chain in r,
    //     y = fma(fma(fma(fma(A0, r, A1), r, A2), r, 1), r, logc'): 5 double-precision instructions instead of e_logf.c's 11 (the
    //     Estrin-like form of e_logf.c in FMAs, which glibc's own __logf_fma variant runs on FMA-capable x86-64 hosts, needs 6).
    //     The roundings differ from e_logf.c's only far below the final rounding to float -- on no input at all, which is not
    //     argued but checked on every one of the 2^24 inputs;
    //   * float -> double of m by integer ops (exact for normal floats).
    __device__ __forceinline__ bool isNormalPositive(float e)
    {
        return (__float_as_uint(e) - 0x00800000u) < 0x7f000000u;
    }
    constexpr int kLogf33 = 33, kLogf33Bias = 7;
    __device__ __forceinline__ LogfEntry buildLogfTable33(int idx, const LogfEntry* __restrict__ tab16)
    {
        const int j = idx - kLogf33Bias, ti = j & 15, tk = j >> 4;                       // tk = -1, 0 or 1
        const LogfEntry e = tab16[ti];
        LogfEntry o;
        o.invc = __dmul_rn(e.invc, tk < 0 ? 2.0 : (tk > 0 ? 0.5 : 1.0));                 // exact
        o.logc = __dadd_rn(e.logc, __dmul_rn((double)tk, 0x1.62e42fefa39efp-1));         // y0 of e_logf.c
        return o;
    }
    // Per-exponent part of fdlibm's log10f for a normal float with exponent field E = 1..254: k = E - 127, i = (k < 0),
    // K = k + i, and the two products that depend on K alone, y*log10_2lo and y*log10_2hi with y = (float)K; z and w hold the integer
    // offsets that turn bits(e) into the logf table index and into the double of the mantissa re-biased to [0.5, 2).  One 16-byte shared-memory load (neighbouring cells have
    // neighbouring energies: mostly a broadcast) replaces 10 integer / conversion / multiply instructions per sample.
    constexpr int kExpEntries = 256;
    __device__ __forceinline__ float4 buildExponentEntry(int E)
    {
        const int k = E - 127;
        const int i = (int)((unsigned)k >> 31);
        const float yk = (float)(k + i);
        // .z: bits(e) - z = the logf table index << 19 (re-bias by K << 23, the table's origin 0x3f330000 and its bias folded in);
        // .w: (bits(e) >> 3) + w = high word of the re-biased mantissa m as a double (0x38000000 = the exponent re-bias of float -> double)
        return make_float4(__fmul_rn(yk, 7.9034151668e-07f), __fmul_rn(yk, 3.0102920532e-01f),
                           __int_as_float(((k + i) << 23) + 0x3f330000 - (kLogf33Bias << 19)), __int_as_float(0x38000000 - ((k + i) << 20)));
    }
    __device__ __forceinline__ float decibelsNormal(float e, const LogfEntry* __restrict__ tab33, const float4* __restrict__ tabExp)
    {
    #ifdef PVC_FAST_LOG10
        return __fmul_rn(__log2f(e), 3.0102999566398120f);
    #else
        const int hx = __float_as_int(e);
        const float4 ex = tabExp[(unsigned)hx >> 23];
        // m = e * 2^-K in [1,2) (e >= 1) or [0.5,1) is never formed as a float: its table index and its double come straight from bits(e)
        const int idx = (hx - __float_as_int(ex.z)) >> 19;                           // == ((bits(m) - 0x3f330000u) >> 19) + kLogf33Bias
        const LogfEntry en = tab33[idx];
        const double md = __hiloint2double((int)(((uint32_t)hx >> 3) + (uint32_t)__float_as_int(ex.w)), (int)((uint32_t)hx << 29));
        const double r = __fma_rn(md, en.invc, -1.0);
        double q = __fma_rn(-0x1.00ea348b88334p-2, r, 0x1.5575b0be00b6ap-2);
        q = __fma_rn(q, r, -0x1.ffffef20a4123p-2);
        q = __fma_rn(q, r, 1.0);
        const double y = __fma_rn(q, r, en.logc);
        const float lf = (float)y;
        const float zz = __fadd_rn(ex.x, __fmul_rn(4.3429449201e-01f, lf));
        return __fmul_rn(10.f, __fadd_rn(zz, ex.y));
    #endif
    }

    struct AnalyzeParams
    {
        int T, fs;
        int fluxSamples, drySamples, wetSamples, tailSamples;
        float dx, courant, efree;
        int resolution;
    };

    // MINB: blocks per SM the register allocation aims at.  0 = the compiler's own choice (64 registers, no spills) is
    // best where the kernel is a chain of dependent loads (small grids: 70^2 0.077 ms against 0.083); 10 (48 registers, 16 bytes of
    // spill) trades instruction-level for thread-level parallelism and is 4-6 % faster once the grid fills the GPU (HugeRoom
    // 2 x 2048^2: 26.1 -> 24.9 ms; 12 blocks: 24.8 but slower on everything smaller; 6 / 4 blocks = 80 / 124 registers: 29.2 / 32.1 ms).
#ifndef PVC_AN_DENSE_MINB
#define PVC_AN_DENSE_MINB 10
#endif
    template <int HC, int MINB, int BATCH>          // HC: history strip width (Layout::hist_chunk) as a compile-time stride
    __global__ void __launch_bounds__(128, MINB)
    encodeResponseKernel(Layout L, AnalyzeParams A, const float* __restrict__ hist, const float* __restrict__ w,
                         const SourceParams* __restrict__ src, float* __restrict__ results,
                         float* __restrict__ delay, float* __restrict__ walkDelay, const int* __restrict__ firstActive)
    {
        __shared__ LogfEntry sTab[16];
        __shared__ LogfEntry sTab33[kLogf33];
        if (threadIdx.x < 16) sTab[threadIdx.x] = kLogfTable[threadIdx.x];
        if (threadIdx.x < kLogf33) sTab33[threadIdx.x] = buildLogfTable33(threadIdx.x, kLogfTable);
        __shared__ float4 sTabExp[kExpEntries];
        for (int E = threadIdx.x; E < kExpEntries; E += blockDim.x) sTabExp[E] = buildExponentEntry(E);
        __syncthreads();

        const int c = blockIdx.x * HC + threadIdx.x;              // block = one history strip of one row (spare threads if the strip is narrower)
        const int r = blockIdx.y;
        const int s = blockIdx.z;
        if ((int)threadIdx.x >= HC || c >= L.gy) return;
        const size_t cells = (size_t)L.gx * L.gy;
        // interior cell (r, c) -> r*gy + c.  The reference strides by the x extent (INDEX_TO_POS, PvDefinitions.h:23-24),
        // which is the same thing on the square grids it supports and self-overlapping on others.
        const size_t serial = (size_t)r * L.gy + c;
        float* out = results + ((size_t)s * cells + serial) * 8;
        const int T = A.T;
        // sample t of this cell is H[t * hist_chunk]: the 4 warps of the block walk one contiguous stream
        const float* H = hist + (size_t)s * L.hist_source + histCell(L, r, c);
        constexpr ptrdiff_t hs = HC;

        const size_t wi = cellIndex(L, r, c);
        const float wSelf = w[wi];
        if (!isAirA(wSelf))
        {   // a wall cell's pressure is identically zero: no onset, results left untouched (Analyzer.cpp:161-165)
            delay[(size_t)s * cells + serial] = FLT_MAX;
            walkDelay[(size_t)s * cells + serial] = FLT_MAX;
            return;
        }

        // ---- activity hints from the fused step kernel: every recorded sample of this cell before launch `mine` (4 steps
        //      per launch) is exactly zero, likewise for the up / left neighbours; skipping exact zeros changes no sum ----
        int onsetBegin = 0, causalBegin = 0;
        if (firstActive)
        {
            const int* fa = firstActive + (size_t)s * L.tiles_x * L.tiles_y * 32;
            auto blockFirst = [&](int rr, int cc) {
                const int ty = rr / L.valid_rows, tx = cc / kValidCols;
                const int wIdx = (rr - ty * L.valid_rows + kTileK) / L.warp_rows;
                return fa[((size_t)ty * L.tiles_x + tx) * 32 + wIdx];
            };
            const int mine = blockFirst(r, c);
            if (mine >= kNeverActive)
            {   // never anything but zeros: no onset (Analyzer.cpp:161-165)
                delay[(size_t)s * cells + serial] = FLT_MAX;
                walkDelay[(size_t)s * cells + serial] = FLT_MAX;
                return;
            }
            const int up = (r > 0) ? blockFirst(r - 1, c) : mine, left = (c > 0) ? blockFirst(r, c - 1) : mine;
            onsetBegin = min(mine * kTileK, T);
            causalBegin = min(min(mine, min(up, left)) * kTileK, T);
        }

        // ---- onset: first sample with |p| > threshold (Analyzer.cpp:146-154); kBatch loads in flight ----
    #ifndef PVC_AN_BATCH
    #define PVC_AN_BATCH 16
    #endif
        // streaming loads in flight per thread: 16 (A/B 8 / 16 / 32 on grids that fill the GPU), 32 in the instantiation for tiny grids --
        // there a thread's chain of dependent load rounds IS the kernel's run time, and registers are no concern
        constexpr int kBatch = BATCH;
        int onset = -1;
        for (int t0 = onsetBegin; t0 < T && onset < 0; t0 += kBatch)
        {
            float v[kBatch];
            #pragma unroll
            for (int u = 0; u < kBatch; ++u) v[u] = __ldg(H + (ptrdiff_t)min(t0 + u, T - 1) * hs);
            #pragma unroll
            for (int u = kBatch - 1; u >= 0; --u)
                if (t0 + u < T && fabsf(v[u]) > kAudibleThreshold) onset = t0 + u;
        }
        if (onset < 0)
        {
            delay[(size_t)s * cells + serial] = FLT_MAX;
            walkDelay[(size_t)s * cells + serial] = FLT_MAX;
            return;
        }
        const int directEnd = onset + A.drySamples;
        const int dryEnd = min(directEnd, T), fluxEnd = min(onset + A.fluxSamples, T);

        // ---- Edry over [0, onset+D), flux over [0, onset+Sd), both from sample 0 (Analyzer.cpp:182-195).
        //      vx, vy of the cell are rebuilt from the recorded pressures of the cell and of its up / left
        //      neighbours (the neighbouring threads' lines: L1 hits) with the solver's own rules. ----
        float edry = 0.f, fx = 0.f, fy = 0.f;
        {
            const float wUp = w[wi - L.pitch], wLeft = w[wi - 1];
            const bool topEdge = (r == 0), leftEdge = (c == 0);
            const bool upAir = isAirA(wUp), leftAir = isAirA(wLeft);
            const ptrdiff_t upOff = topEdge ? 0 : -(ptrdiff_t)L.hist_row;
            const ptrdiff_t leftOff = leftEdge ? 0 : (((c % HC) != 0) ? -1 : -(ptrdiff_t)T * hs + (hs - 1));
            float vx = 0.f, vy = 0.f;
            constexpr int kCausalBatch = BATCH / 4;
            for (int t0 = causalBegin; t0 < fluxEnd; t0 += kCausalBatch)
            {
                float bp[kCausalBatch], bu[kCausalBatch], bl[kCausalBatch];
                #pragma unroll
                for (int u = 0; u < kCausalBatch; ++u)
                {
                    const float* q = H + (ptrdiff_t)min(t0 + u, T - 1) * hs;
                    bp[u] = __ldg(q);
                    bu[u] = __ldg(q + upOff);
                    bl[u] = __ldg(q + leftOff);
                }
                #pragma unroll
                for (int u = 0; u < kCausalBatch; ++u)
                {
                    if (t0 + u >= fluxEnd) break;
                    const float p = bp[u];
                    vx = topEdge ? -p : (upAir ? __fsub_rn(vx, __fmul_rn(A.courant, __fsub_rn(p, bu[u]))) : -__fmul_rn(wUp, p));
                    vy = leftEdge ? -p : (leftAir ? __fsub_rn(vy, __fmul_rn(A.courant, __fsub_rn(p, bl[u]))) : -__fmul_rn(wLeft, p));
                    edry = __fadd_rn(edry, __fmul_rn(p, p));
                    fx = __fadd_rn(fx, __fmul_rn(p, vx));
                    fy = __fadd_rn(fy, __fmul_rn(p, vy));
                }
            }
            const float* q = H + (ptrdiff_t)fluxEnd * hs;
            for (int t = fluxEnd; t < dryEnd; t += kBatch, q += kBatch * hs)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldg(q + min(u, dryEnd - 1 - t) * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) if (t + u < dryEnd) edry = __fadd_rn(edry, __fmul_rn(v[u], v[u]));
            }
        }

        // ---- obstruction gain and source directivity (Analyzer.cpp:199-220, FreeGrid.cpp:41-59) ----
        const SourceParams sp = src[s];
        float efreePr;
        {
            const float lX = __fmul_rn((float)sp.efree_r, A.dx), lY = __fmul_rn((float)sp.efree_c, A.dx);
            const float eX = __fmul_rn((float)r, A.dx), eY = __fmul_rn((float)c, A.dx);
            const float ddx = __fsub_rn(eX, lX), ddy = __fsub_rn(eY, lY);
            const float rr = __fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
            efreePr = (rr == 0.f) ? A.efree : __fdiv_rn(A.efree, rr);
        }
        const float occ = __fsqrt_rn(__fdiv_rn(edry, efreePr));
        float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
        norm = __fdiv_rn(-1.0f, (norm > 0.0f ? norm : 1.0f));
        const float sdx = __fmul_rn(norm, fx), sdy = __fmul_rn(norm, fy);

        // ---- low-pass cutoff (Analyzer.cpp:227-230); powf evaluated in double then rounded: equals libm's
        //      powf except on rare cells where that one is 1 ulp off the correctly rounded value ----
        const float rinv = __fdiv_rn(1.0f, fmaxf(0.001f, occ));
        const float pw = (float)pow((double)__fdiv_rn(rinv, 12.f), (double)0.8f);
        const float lowpass = __fadd_rn(-147.f, __fdiv_rn(18390.f, __fadd_rn(1.f, pw)));

        // ---- wet gain over [directEnd+1, min(directEnd+1+W, T)) (Analyzer.cpp:235-247) ----
        float wet = 0.f;
        {
            const int end = min(directEnd + 1 + A.wetSamples, T);
            const float* q = H + (ptrdiff_t)(directEnd + 1) * hs;
            int j = directEnd + 1;
            for (; j + kBatch <= end; j += kBatch, q += kBatch * hs)        // kBatch loads in flight: on small grids a dependent load per few samples IS the frame time
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldg(q + u * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) wet = __fadd_rn(wet, __fmul_rn(v[u], v[u]));
            }
            if (j < end)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldg(q + min(u, end - 1 - j) * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) if (j + u < end) wet = __fadd_rn(wet, __fmul_rn(v[u], v[u]));
            }
        }
        const float wetGain = __fsqrt_rn(__fdiv_rn(wet, A.efree));

        // ---- RT60: backward Schroeder integration + closed-form regression (Analyzer.cpp:282-326) ----
        const int start = directEnd + 1;
        const int endPoint = T - A.tailSamples;
        const int regressN = endPoint - start;
        const float rn = (float)regressN;
        const float xmean = __fmul_rn(__fsub_rn(rn, 1.0f), 0.5f);
        const float xsum = __fmul_rn(rn, xmean);
        const float denominator = __fmul_rn(__fmul_rn(1.0f / 12.0f, rn), __fsub_rn(__fmul_rn(rn, rn), 1.0f));
        float edc = 0.f, xysum = 0.f, ysum = 0.f;
        // both backward loops walk one pointer down the strip and keep kBatch streaming loads in flight per
        // thread (one outstanding 4-byte load per thread leaves the pass latency-bound far below HBM speed)
        {
            int i = T - 1;
            const int stop = max(endPoint, 0);
            const float* q = H + (ptrdiff_t)i * hs;
            for (; i - (kBatch - 1) >= stop; i -= kBatch, q -= kBatch * hs)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - u * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
            }
            if (i >= stop)
            {   // the last, partial batch: its loads in flight together as well (clamped into the range, extra ones unused)
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - min(u, i - stop) * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) if (i - u >= stop) edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
            }
        }
        if (endPoint - 1 >= start)
        {
            int i = endPoint - 1;
            float x = (float)(i - start);              // exact; decremented by 1.0f per sample (< 2^24)
            const float* q = H + (ptrdiff_t)i * hs;
            for (; i - (kBatch - 1) >= start; i -= kBatch, q -= kBatch * hs)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - u * hs);
                // the kBatch Schroeder values first (one dependent add each), then their logarithms, which are
                // independent of each other: straight-line when all are normal floats (all but a vanishing few batches)
                float e[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u)
                {
                    edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
                    e[u] = edc;
                }
                // the running sum of squares never decreases (a NaN or an infinity sticks): the first and the last value of
                // the batch bracket the others
                const bool normal = isNormalPositive(e[0]) && isNormalPositive(e[kBatch - 1]);
                float y[kBatch];
                if (normal)
                {
                    #pragma unroll
                    for (int u = 0; u < kBatch; ++u) y[u] = decibelsNormal(e[u], sTab33, sTabExp);
                }
                else
                {
                    #pragma unroll
                    for (int u = 0; u < kBatch; ++u) y[u] = decibels(e[u], sTab);
                }
                #pragma unroll
                for (int u = 0; u < kBatch; ++u)
                {
                    xysum = __fadd_rn(xysum, __fmul_rn(y[u], x));
                    ysum = __fadd_rn(ysum, y[u]);
                    x = __fsub_rn(x, 1.0f);
                }
            }
            if (i >= start)
            {   // the last, partial batch
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - min(u, i - start) * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    if (i - u >= start)
                    {
                        edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
                        const float y = decibels(edc, sTab);
                        xysum = __fadd_rn(xysum, __fmul_rn(y, x));
                        ysum = __fadd_rn(ysum, y);
                        x = __fsub_rn(x, 1.0f);
                    }
            }
        }
        const float ymean = __fdiv_rn(ysum, rn);
        float numerator = __fsub_rn(xysum, __fmul_rn(ymean, xsum));
        numerator = __fsub_rn(numerator, __fmul_rn(xmean, ysum));
        numerator = __fadd_rn(numerator, __fmul_rn(__fmul_rn(rn, xmean), ymean));
        const float slopePerSample = __fdiv_rn(numerator, denominator);
        const float slopePerSec = __fmul_rn(slopePerSample, (float)A.fs);
        const float rt60 = __fdiv_rn(-60.f, slopePerSec);

        out[0] = occ; out[1] = wetGain; out[2] = rt60; out[3] = lowpass;
        out[6] = sdx; out[7] = sdy;
        delay[(size_t)s * cells + serial] = (float)onset;
        walkDelay[(size_t)s * cells + serial] = (occ > 0.f) ? (float)onset : FLT_MAX;
    }

    // ---- streamed solve: the same analysis over a history that holds one chunk of the response at a time ----------------
    // EncodeResponse (Analyzer.cpp:139-328) reads a cell's response twice: ascending from sample 0 (onset, Edry, flux, wet
    // energy) and descending from the last sample (Schroeder integral + regression).  With a bounded history the ascending part
    // runs chunk by chunk during the forward sweep (forwardChunkKernel) and the descending part chunk by chunk over the
    // RECOMPUTED chunks in reverse order (backwardChunkKernel; chunk 0 comes last and writes the results).  Every sum is
    // carried per cell between chunks in fp32 and continues in the reference's order with the reference's operations, so the
    // outputs are bit-identical to encodeResponseKernel's on the full history (tests/test_gpu_streamed.py).
    // Carry planes ([kCarryPlanes][sources * cells]): 0 onset (int bits, -1 none), 1 edry, 2 fx, 3 fy, 4 vx, 5 vy, 6 wet,
    // 7 edc, 8 xysum, 9 ysum.  Sample t of the response is sample t - base of the history; L.T is the history's length.
    template <int HC>
    __global__ void __launch_bounds__(128)
    forwardChunkKernel(Layout L, AnalyzeParams A, const float* __restrict__ hist, const float* __restrict__ w,
                       float* __restrict__ carry, size_t cstride, int base, int len, const int* __restrict__ firstActive)
    {
        const int c = blockIdx.x * HC + threadIdx.x;
        const int r = blockIdx.y;
        const int s = blockIdx.z;
        if ((int)threadIdx.x >= HC || c >= L.gy) return;
        const size_t cells = (size_t)L.gx * L.gy;
        const size_t ci = (size_t)s * cells + (size_t)r * L.gy + c;
        const size_t wi = cellIndex(L, r, c);
        const float wSelf = w[wi];
        if (!isAirA(wSelf)) return;                          // a wall cell never has an onset: its carry keeps onset = -1
        const int T = A.T;
        const int end = base + len;                          // one past the chunk's last sample
        int onset = __float_as_int(carry[ci]);
        if (onset >= 0 && base >= min(onset + A.drySamples + 1 + A.wetSamples, T)) return;      // every causal window is closed
        const float* H = hist + (size_t)s * L.hist_source + histCell(L, r, c) - (ptrdiff_t)base * HC;     // sample t at H[t * HC]
        constexpr ptrdiff_t hs = HC;
        constexpr int kBatch = 16;
        // activity hints of THIS chunk's step launch (generations counted from the chunk's first sample): everything this cell,
        // its up and its left neighbour recorded before them is exactly zero -- no onset there, every sum unchanged, the rebuilt
        // velocities unchanged up to the sign of a zero that only ever multiplies those zero pressures
        int onsetBegin = base, causalBegin = base;
        if (firstActive)
        {
            const int* fa = firstActive + (size_t)s * L.tiles_x * L.tiles_y * 32;
            auto blockFirst = [&](int rr, int cc) {
                const int ty = rr / L.valid_rows, tx = cc / kValidCols;
                const int wIdx = (rr - ty * L.valid_rows + kTileK) / L.warp_rows;
                const int g = fa[((size_t)ty * L.tiles_x + tx) * 32 + wIdx];
                return g >= kNeverActive ? len : min(g * kTileK, len);
            };
            const int mine = blockFirst(r, c);
            const int up = (r > 0) ? blockFirst(r - 1, c) : mine, left = (c > 0) ? blockFirst(r, c - 1) : mine;
            onsetBegin = base + mine;
            causalBegin = base + min(mine, min(up, left));
            if (causalBegin >= end) return;                  // nothing but zeros in this chunk (then there is no onset behind it either)
        }
        if (onset < 0)
        {   // Analyzer.cpp:146-154, continued where the previous chunk stopped
            for (int t0 = onsetBegin; t0 < end && onset < 0; t0 += kBatch)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldg(H + (ptrdiff_t)min(t0 + u, end - 1) * hs);
                #pragma unroll
                for (int u = kBatch - 1; u >= 0; --u)
                    if (t0 + u < end && fabsf(v[u]) > kAudibleThreshold) onset = t0 + u;
            }
            if (onset >= 0) carry[ci] = __int_as_float(onset);
        }
        // no onset yet: every sample of the chunk precedes it, and so lies inside both dry windows (they start at sample 0)
        const int directEnd = onset >= 0 ? onset + A.drySamples : T;
        const int dryEnd = min(min(directEnd, T), end);
        const int fluxEnd = min(onset >= 0 ? min(onset + A.fluxSamples, T) : T, end);
        float edry = carry[cstride + ci];
        if (base < fluxEnd)
        {
            float fx = carry[2 * cstride + ci], fy = carry[3 * cstride + ci], vx = carry[4 * cstride + ci], vy = carry[5 * cstride + ci];
            const float wUp = w[wi - L.pitch], wLeft = w[wi - 1];
            const bool topEdge = (r == 0), leftEdge = (c == 0);
            const bool upAir = isAirA(wUp), leftAir = isAirA(wLeft);
            const ptrdiff_t upOff = topEdge ? 0 : -(ptrdiff_t)L.hist_row;
            const ptrdiff_t leftOff = leftEdge ? 0 : (((c % HC) != 0) ? -1 : -(ptrdiff_t)L.T * hs + (hs - 1));
            constexpr int kCausalBatch = 4;
            for (int t0 = causalBegin; t0 < fluxEnd; t0 += kCausalBatch)
            {
                float bp[kCausalBatch], bu[kCausalBatch], bl[kCausalBatch];
                #pragma unroll
                for (int u = 0; u < kCausalBatch; ++u)
                {
                    const float* q = H + (ptrdiff_t)min(t0 + u, end - 1) * hs;
                    bp[u] = __ldg(q);
                    bu[u] = __ldg(q + upOff);
                    bl[u] = __ldg(q + leftOff);
                }
                #pragma unroll
                for (int u = 0; u < kCausalBatch; ++u)
                {
                    if (t0 + u >= fluxEnd) break;
                    const float p = bp[u];
                    vx = topEdge ? -p : (upAir ? __fsub_rn(vx, __fmul_rn(A.courant, __fsub_rn(p, bu[u]))) : -__fmul_rn(wUp, p));
                    vy = leftEdge ? -p : (leftAir ? __fsub_rn(vy, __fmul_rn(A.courant, __fsub_rn(p, bl[u]))) : -__fmul_rn(wLeft, p));
                    edry = __fadd_rn(edry, __fmul_rn(p, p));
                    fx = __fadd_rn(fx, __fmul_rn(p, vx));
                    fy = __fadd_rn(fy, __fmul_rn(p, vy));
                }
            }
            carry[2 * cstride + ci] = fx; carry[3 * cstride + ci] = fy; carry[4 * cstride + ci] = vx; carry[5 * cstride + ci] = vy;
        }
        for (int t = max(base, fluxEnd); t < dryEnd; ++t)
        {
            const float p = __ldg(H + (ptrdiff_t)t * hs);
            edry = __fadd_rn(edry, __fmul_rn(p, p));
        }
        carry[cstride + ci] = edry;
        if (onset >= 0)
        {   // wet energy over [directEnd + 1, min(directEnd + 1 + W, T)) (Analyzer.cpp:235-247)
            const int wetEnd = min(min(directEnd + 1 + A.wetSamples, T), end);
            const int wetBegin = max(directEnd + 1, base);
            if (wetBegin < wetEnd)
            {
                float wet = carry[6 * cstride + ci];
                const float* q = H + (ptrdiff_t)wetBegin * hs;
                #pragma unroll 4
                for (int j = wetBegin; j < wetEnd; ++j, q += hs)
                {
                    const float p = __ldg(q);
                    wet = __fadd_rn(wet, __fmul_rn(p, p));
                }
                carry[6 * cstride + ci] = wet;
            }
        }
    }

    template <int HC>
    __global__ void __launch_bounds__(128)
    backwardChunkKernel(Layout L, AnalyzeParams A, const float* __restrict__ hist, const float* __restrict__ w,
                        const SourceParams* __restrict__ src, float* __restrict__ carry, size_t cstride, int base, int len,
                        float* __restrict__ results, float* __restrict__ delay, float* __restrict__ walkDelay)
    {
        __shared__ LogfEntry sTab[16];
        __shared__ LogfEntry sTab33[kLogf33];
        if (threadIdx.x < 16) sTab[threadIdx.x] = kLogfTable[threadIdx.x];
        if (threadIdx.x < kLogf33) sTab33[threadIdx.x] = buildLogfTable33(threadIdx.x, kLogfTable);
        __shared__ float4 sTabExp[kExpEntries];
        for (int E = threadIdx.x; E < kExpEntries; E += blockDim.x) sTabExp[E] = buildExponentEntry(E);
        __syncthreads();

        const int c = blockIdx.x * HC + threadIdx.x;
        const int r = blockIdx.y;
        const int s = blockIdx.z;
        if ((int)threadIdx.x >= HC || c >= L.gy) return;
        const size_t cells = (size_t)L.gx * L.gy;
        const size_t serial = (size_t)r * L.gy + c;
        const size_t ci = (size_t)s * cells + serial;
        const bool last = (base == 0);                       // chunk 0 is processed last: finish the regression, write the results
        const int onset = isAirA(w[cellIndex(L, r, c)]) ? __float_as_int(carry[ci]) : -1;
        if (onset < 0)
        {   // wall cell or no onset: results left untouched (Analyzer.cpp:161-165)
            if (last) { delay[ci] = FLT_MAX; walkDelay[ci] = FLT_MAX; }
            return;
        }
        const int T = A.T;
        const int end = base + len;
        const int directEnd = onset + A.drySamples;
        const int start = directEnd + 1;
        const int endPoint = T - A.tailSamples;
        const float* H = hist + (size_t)s * L.hist_source + histCell(L, r, c) - (ptrdiff_t)base * HC;
        constexpr ptrdiff_t hs = HC;
        constexpr int kBatch = PVC_AN_BATCH;
        float edc = carry[7 * cstride + ci], xysum = carry[8 * cstride + ci], ysum = carry[9 * cstride + ci];
        bool touched = false;
        {   // tail of the curve: energy only (Analyzer.cpp:297-301)
            int i = end - 1;
            const int stop = max(max(endPoint, 0), base);
            if (i >= stop) touched = true;
            const float* q = H + (ptrdiff_t)i * hs;
            for (; i - (kBatch - 1) >= stop; i -= kBatch, q -= kBatch * hs)
            {
                float v[kBatch];
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - u * hs);
                #pragma unroll
                for (int u = 0; u < kBatch; ++u) edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
            }
            for (; i >= stop; --i, q -= hs)
            {
                const float p = __ldcs(q);
                edc = __fadd_rn(edc, __fmul_rn(p, p));
            }
        }
        {   // regression part (Analyzer.cpp:303-319)
            int i = min(endPoint, end) - 1;
            const int stop = max(start, base);
            if (i >= stop)
            {
                touched = true;
                float x = (float)(i - start);              // exact; decremented by 1.0f per sample (< 2^24)
                const float* q = H + (ptrdiff_t)i * hs;
                for (; i - (kBatch - 1) >= stop; i -= kBatch, q -= kBatch * hs)
                {
                    float v[kBatch];
                    #pragma unroll
                    for (int u = 0; u < kBatch; ++u) v[u] = __ldcs(q - u * hs);
                    float e[kBatch];
                    #pragma unroll
                    for (int u = 0; u < kBatch; ++u)
                    {
                        edc = __fadd_rn(edc, __fmul_rn(v[u], v[u]));
                        e[u] = edc;
                    }
                    const bool normal = isNormalPositive(e[0]) && isNormalPositive(e[kBatch - 1]);
                    float y[kBatch];
                    if (normal)
                    {
                        #pragma unroll
                        for (int u = 0; u < kBatch; ++u) y[u] = decibelsNormal(e[u], sTab33, sTabExp);
                    }
                    else
                    {
                        #pragma unroll
                        for (int u = 0; u < kBatch; ++u) y[u] = decibels(e[u], sTab);
                    }
                    #pragma unroll
                    for (int u = 0; u < kBatch; ++u)
                    {
                        xysum = __fadd_rn(xysum, __fmul_rn(y[u], x));
                        ysum = __fadd_rn(ysum, y[u]);
                        x = __fsub_rn(x, 1.0f);
                    }
                }
                for (; i >= stop; --i, q -= hs)
                {
                    const float p = __ldcs(q);
                    edc = __fadd_rn(edc, __fmul_rn(p, p));
                    const float y = decibels(edc, sTab);
                    xysum = __fadd_rn(xysum, __fmul_rn(y, x));
                    ysum = __fadd_rn(ysum, y);
                    x = __fsub_rn(x, 1.0f);
                }
            }
        }
        if (!last)
        {
            if (touched) { carry[7 * cstride + ci] = edc; carry[8 * cstride + ci] = xysum; carry[9 * cstride + ci] = ysum; }
            return;
        }
        // ---- the outputs, exactly as encodeResponseKernel forms them (Analyzer.cpp:199-230, 250, 321-326) ----
        const float edry = carry[cstride + ci], fx = carry[2 * cstride + ci], fy = carry[3 * cstride + ci], wet = carry[6 * cstride + ci];
        const SourceParams sp = src[s];
        float efreePr;
        {
            const float lX = __fmul_rn((float)sp.efree_r, A.dx), lY = __fmul_rn((float)sp.efree_c, A.dx);
            const float eX = __fmul_rn((float)r, A.dx), eY = __fmul_rn((float)c, A.dx);
            const float ddx = __fsub_rn(eX, lX), ddy = __fsub_rn(eY, lY);
            const float rr = __fsqrt_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)));
            efreePr = (rr == 0.f) ? A.efree : __fdiv_rn(A.efree, rr);
        }
        const float occ = __fsqrt_rn(__fdiv_rn(edry, efreePr));
        float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)));
        norm = __fdiv_rn(-1.0f, (norm > 0.0f ? norm : 1.0f));
        const float sdx = __fmul_rn(norm, fx), sdy = __fmul_rn(norm, fy);
        const float rinv = __fdiv_rn(1.0f, fmaxf(0.001f, occ));
        const float pw = (float)pow((double)__fdiv_rn(rinv, 12.f), (double)0.8f);
        const float lowpass = __fadd_rn(-147.f, __fdiv_rn(18390.f, __fadd_rn(1.f, pw)));
        const float wetGain = __fsqrt_rn(__fdiv_rn(wet, A.efree));
        const int regressN = endPoint - start;
        const float rn = (float)regressN;
        const float xmean = __fmul_rn(__fsub_rn(rn, 1.0f), 0.5f);
        const float xsum = __fmul_rn(rn, xmean);
        const float denominator = __fmul_rn(__fmul_rn(1.0f / 12.0f, rn), __fsub_rn(__fmul_rn(rn, rn), 1.0f));
        const float ymean = __fdiv_rn(ysum, rn);
        float numerator = __fsub_rn(xysum, __fmul_rn(ymean, xsum));
        numerator = __fsub_rn(numerator, __fmul_rn(xmean, ysum));
        numerator = __fadd_rn(numerator, __fmul_rn(__fmul_rn(rn, xmean), ymean));
        const float slopePerSample = __fdiv_rn(numerator, denominator);
        const float slopePerSec = __fmul_rn(slopePerSample, (float)A.fs);
        const float rt60 = __fdiv_rn(-60.f, slopePerSec);
        float* out = results + ci * 8;
        out[0] = occ; out[1] = wetGain; out[2] = rt60; out[3] = lowpass;
        out[6] = sdx; out[7] = sdy;
        delay[ci] = (float)onset;
        walkDelay[ci] = (occ > 0.f) ? (float)onset : FLT_MAX;
    }

    // ---- Analyzer::EncodeListenerDirection (Analyzer.cpp:340-431) by pointer jumping ----------------------------------
    // walkDelay holds the onset of every cell a walk may step onto (has an onset and occlusion > 0, Analyzer.cpp:372-374)
    // and FLT_MAX elsewhere.  The reference walks, from every cell, to the 8-neighbour with the strictly smallest delay
    // (first wins in the scan order of :332-337) until the delay is <= 5 samples, the loudness >= 0.891, there is no
    // strictly earlier neighbour (the walk then ends POINTING AT that neighbour, :374-386) or the geodesic matches the
    // Euclidean distance (:392-407).  Walking costs O(path) per cell -- 25 ms for two 2048 x 2048 sources, more than
    // the impulse-response analysis itself.  But once a walk stands on a cell u != start, everything that follows is a
    // pure function of u (its current delay is wd[u], its loudness occ[u]); only the first hop differs (delay = FLT_MAX:
    // no "no improvement" exit, and the start itself is tested for loudness only).  So:
    //   1. walkNextKernel    next[u] = the cell the walk ends on if it ends at/after u without another move (DONE bit),
    //                        else the cell it moves to;
    //   2. walkJumpKernel    next[u] <- next[next[next[next[u]]]] in place, ceil(log4 T) + 1 passes (delays are integral
    //                        sample indices that strictly decrease along a walk, so no path is longer than T);
    //   3. walkResolveKernel the first hop from every start cell, the end cell from next[], the unit vector (:413-430).
    // Same exits, same tie-breaking, same fp32 expressions as the sequential walk (which remains the cross-check in
    // tests/test_gpu_parity.py through PVC_WALK=sequential).
    constexpr int kWalkDone = (int)0x80000000;

    struct WalkConsts { float samplingRate, thresholdDist; };
    __device__ __forceinline__ WalkConsts walkConsts(const AnalyzeParams& A)
    {
        WalkConsts k;
        k.samplingRate = (float)A.fs;
        k.thresholdDist = __fmul_rn(0.3f, __fdiv_rn(kSpeedOfSoundA, (float)A.resolution));
        return k;
    }
    // strictly smallest walkDelay among the 8 neighbours of (r, c) inside the lattice, first wins; -1 if none is selectable
    __device__ __forceinline__ int walkArgmin(const float* __restrict__ wd, int r, int c, int gx, int gy, float& best)
    {
        int arg = -1;
        best = FLT_MAX;
        #pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            const int dr = (k < 3) ? -1 : (k < 5 ? 0 : 1);
            const int dc = (k == 0 || k == 3 || k == 5) ? -1 : ((k == 1 || k == 6) ? 0 : 1);
            const int nr = r + dr, nc = c + dc;
            if (nr < 0 || nc < 0 || nr >= gx || nc >= gy) continue;
            const int ni = nr * gy + nc;
            const float d = wd[ni];
            if (d < best) { arg = ni; best = d; }
        }
        return arg;
    }
    // line-of-sight exit of a walk standing on (r, c) with delay d (Analyzer.cpp:392-407)
    __device__ __forceinline__ bool walkLineOfSight(const AnalyzeParams& A, const WalkConsts& K, const SourceParams& sp, int r, int c, float d)
    {
        const float geodesic = __fdiv_rn(__fmul_rn(kSpeedOfSoundA, d), K.samplingRate);
        const float tx = __fsub_rn(__fmul_rn((float)r, A.dx), sp.x);
        const float ty = __fsub_rn(__fmul_rn((float)c, A.dx), sp.z);
        const float eu = __fsqrt_rn(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)));
        return fabsf(__fsub_rn(geodesic, eu)) < K.thresholdDist;
    }

    // link of cell u = (r, c): the cell a walk standing on u ends on without another move (kWalkDone set) or moves to
    __device__ __forceinline__ int walkLink(const AnalyzeParams& A, const SourceParams& sp, const float* __restrict__ res, const float* __restrict__ wd,
                                            int r, int c, int gx, int gy)
    {
        const int u = r * gy + c;
        const float d = wd[u];
        int out = u | kWalkDone;
        if (d != FLT_MAX)                    // cells no walk can stand on keep a harmless self link
        {
            const WalkConsts K = walkConsts(A);
            const float loud = res[(size_t)u * 8];
            const bool goOn = (d > kDelayClose && loud < kGainThreshold) && !walkLineOfSight(A, K, sp, r, c, d);
            if (goOn)
            {
                float best;
                const int v = walkArgmin(wd, r, c, gx, gy, best);
                if (v >= 0) out = (best >= d) ? (v | kWalkDone) : v;
            }
        }
        return out;
    }

    __global__ void __launch_bounds__(128)
    walkNextKernel(Layout L, AnalyzeParams A, const SourceParams* __restrict__ src, const float* __restrict__ results,
                   const float* __restrict__ walkDelay, int* __restrict__ next)
    {
        const int c = blockIdx.x * blockDim.x + threadIdx.x;
        const int r = blockIdx.y;
        const int s = blockIdx.z;
        if (c >= L.gy) return;
        const size_t cells = (size_t)L.gx * L.gy;
        next[(size_t)s * cells + (size_t)r * L.gy + c] = walkLink(A, src[s], results + (size_t)s * cells * 8, walkDelay + (size_t)s * cells, r, c, L.gx, L.gy);
    }

    __global__ void __launch_bounds__(256)
    walkJumpKernel(size_t cells, int jumps, int* __restrict__ nextAll)
    {
        const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (u >= cells) return;
        int* next = nextAll + (size_t)blockIdx.y * cells;
        int n = next[u];
        if (n < 0) return;
        // in place: a link read here may already have been shortened by another thread -- it then only points
        // further along the same walk
        for (int k = 0; k < jumps && n >= 0; ++k) n = *((volatile int*)next + n);
        next[u] = n;
    }

    // the walk from start cell (r0, c0) through the shortened links, and the direction it yields (Analyzer.cpp:357-386, 409-431)
    __device__ __forceinline__ void walkResolve(const AnalyzeParams& A, const SourceParams& sp, float* __restrict__ res, const float* __restrict__ wd,
                                                const int* next, int r0, int c0, int gx, int gy)
    {
        int target = r0 * gy + c0;
        // first hop: delay = FLT_MAX, so only the start's loudness can stop the walk before it moves, and any selectable
        // neighbour is an improvement (Analyzer.cpp:357-386)
        if (res[(size_t)target * 8] < kGainThreshold)
        {
            float best;
            const int v = walkArgmin(wd, r0, c0, gx, gy, best);
            if (v >= 0)
            {
                int n = next[v];
                while (n >= 0) n = next[n];            // whatever the jump rounds left (see launchListenerDirection)
                target = n & 0x7fffffff;
            }
        }
        const int r = target / gy, c = target % gy;
        float ox = __fsub_rn(__fmul_rn((float)r, A.dx), sp.x);
        float oy = __fsub_rn(__fmul_rn((float)c, A.dx), sp.z);
        float len = __fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy));
        if (len != 0.f)
        {
            len = __fsqrt_rn(len);
            ox = __fdiv_rn(ox, len);
            oy = __fdiv_rn(oy, len);
        }
        float* out = res + ((size_t)r0 * gy + c0) * 8;
        out[4] = ox; out[5] = oy;
    }

    __global__ void __launch_bounds__(128)
    walkResolveKernel(Layout L, AnalyzeParams A, const SourceParams* __restrict__ src, float* __restrict__ results,
                      const float* __restrict__ walkDelay, const int* __restrict__ nextAll)
    {
        const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
        const int r0 = blockIdx.y;
        const int s = blockIdx.z;
        if (c0 >= L.gy) return;
        const size_t cells = (size_t)L.gx * L.gy;
        walkResolve(A, src[s], results + (size_t)s * cells * 8, walkDelay + (size_t)s * cells, nextAll + (size_t)s * cells, r0, c0, L.gx, L.gy);
    }

    // Small grids (the reference's own contract: 70 x 70 .. 127 x 127 cells): links, every jump round and the resolve in ONE launch,
    // one CTA per source, block barriers between the phases -- eight dependent launches cost more than the work they carry there.
    constexpr int kWalkSmallCells = 128 * 128;
    __global__ void __launch_bounds__(1024)
    walkSmallKernel(Layout L, AnalyzeParams A, const SourceParams* __restrict__ src, float* __restrict__ results,
                    const float* __restrict__ walkDelay, int rounds, int hops)
    {
        extern __shared__ int next[];                        // the links live in shared memory: a hop costs 30 cycles instead of an L2 round trip
        const int s = blockIdx.x;
        const int gx = L.gx, gy = L.gy, cells = gx * gy;
        float* res = results + (size_t)s * cells * 8;
        const float* wd = walkDelay + (size_t)s * cells;
        const SourceParams sp = src[s];
        for (int u = threadIdx.x; u < cells; u += blockDim.x) next[u] = walkLink(A, sp, res, wd, u / gy, u % gy, gx, gy);
        __syncthreads();
        for (int k = 0; k < rounds; ++k)
        {
            for (int u = threadIdx.x; u < cells; u += blockDim.x)
            {
                int n = *((volatile int*)next + u);
                if (n < 0) continue;
                for (int h = 0; h < hops && n >= 0; ++h) n = *((volatile int*)next + n);
                *((volatile int*)next + u) = n;
            }
            __syncthreads();
        }
        for (int u = threadIdx.x; u < cells; u += blockDim.x) walkResolve(A, sp, res, wd, next, u / gy, u % gy, gx, gy);
    }

    // The reference's walk, one thread per start cell (cross-check of the pointer-jumping kernels: PVC_WALK=sequential)
    __global__ void __launch_bounds__(128)
    listenerDirectionKernel(Layout L, AnalyzeParams A, const SourceParams* __restrict__ src,
                            float* __restrict__ results, const float* __restrict__ walkDelay)
    {
        const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
        const int r0 = blockIdx.y;
        const int s = blockIdx.z;
        if (c0 >= L.gy) return;
        const int gx = L.gx, gy = L.gy;
        const size_t cells = (size_t)gx * gy;
        float* res = results + (size_t)s * cells * 8;
        const float* wd = walkDelay + (size_t)s * cells;
        const SourceParams sp = src[s];
        const WalkConsts K = walkConsts(A);

        int nextIndex = r0 * gy + c0;
        float loudness = res[(size_t)nextIndex * 8];
        float delay = FLT_MAX;
        while (delay > kDelayClose && loudness < kGainThreshold)
        {
            const int r = nextIndex / gy, c = nextIndex % gy;
            float nextDelay;
            const int v = walkArgmin(wd, r, c, gx, gy, nextDelay);
            if (v >= 0) nextIndex = v;
            if (nextDelay == FLT_MAX || nextDelay >= delay) break;
            delay = nextDelay;
            loudness = res[(size_t)nextIndex * 8];
            if (walkLineOfSight(A, K, sp, nextIndex / gy, nextIndex % gy, nextDelay)) break;
        }
        const int r = nextIndex / gy, c = nextIndex % gy;
        float ox = __fsub_rn(__fmul_rn((float)r, A.dx), sp.x);
        float oy = __fsub_rn(__fmul_rn((float)c, A.dx), sp.z);
        float len = __fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy));
        if (len != 0.f)
        {
            len = __fsqrt_rn(len);
            ox = __fdiv_rn(ox, len);
            oy = __fdiv_rn(oy, len);
        }
        float* out = res + ((size_t)r0 * gy + c0) * 8;
        out[4] = ox; out[5] = oy;
    }

    // Grid::GetResponse for one alloc cell: T x {p, vx, vy}, velocities rebuilt from the history
    __global__ void rebuildIrKernel(Layout L, int T, float courant, const float* __restrict__ hist,
                                    const float* __restrict__ w, int r, int c, float* __restrict__ out)
    {
        if (blockIdx.x != 0 || threadIdx.x != 0) return;
        const float* H = hist + histCell(L, r, c);
        const float* Hu = (r > 0) ? hist + histCell(L, r - 1, c) : H;
        const float* Hl = (c > 0) ? hist + histCell(L, r, c - 1) : H;
        const size_t wi = cellIndex(L, r, c);
        const float wSelf = w[wi], wUp = w[wi - L.pitch], wLeft = w[wi - 1];
        const bool aSelf = isAirA(wSelf), aUp = isAirA(wUp), aLeft = isAirA(wLeft);
        float vx = 0.f, vy = 0.f;
        for (int t = 0; t < T; ++t)
        {
            const float p = H[(size_t)t * L.hist_chunk];
            const float pu = (r > 0) ? Hu[(size_t)t * L.hist_chunk] : 0.f;
            const float pl = (c > 0) ? Hl[(size_t)t * L.hist_chunk] : 0.f;
            if (c >= L.gy) vx = 0.f;
            else if (r == 0) vx = -p;
            else if (r == L.gx) vx = pu;
            else if (aSelf && aUp) vx = __fsub_rn(vx, __fmul_rn(courant, __fsub_rn(p, pu)));
            else if (!aSelf && aUp) vx = __fmul_rn(wSelf, pu);
            else if (aSelf && !aUp) vx = -__fmul_rn(wUp, p);
            else vx = 0.f;
            if (r >= L.gx) vy = 0.f;
            else if (c == 0) vy = -p;
            else if (c == L.gy) vy = pl;
            else if (aSelf && aLeft) vy = __fsub_rn(vy, __fmul_rn(courant, __fsub_rn(p, pl)));
            else if (!aSelf && aLeft) vy = __fmul_rn(wSelf, pl);
            else if (aSelf && !aLeft) vy = -__fmul_rn(wLeft, p);
            else vy = 0.f;
            out[3 * t] = p; out[3 * t + 1] = vx; out[3 * t + 2] = vy;
        }
    }

    static AnalyzeParams paramsOf(const pvc_solver* s)
    {
        AnalyzeParams A;
        A.T = s->cfg.T; A.fs = s->cfg.fs;
        A.fluxSamples = s->cfg.flux_samples; A.drySamples = s->cfg.dry_samples;
        A.wetSamples = s->cfg.wet_samples; A.tailSamples = s->cfg.tail_samples;
        A.dx = s->cfg.dx; A.courant = s->cfg.courant; A.efree = s->efree;
        A.resolution = s->cfg.resolution;
        return A;
    }

    int launchAnalyzer(pvc_solver* s, int nsrc, int* launches)
    {
        const Layout& L = s->L;
        const AnalyzeParams A = paramsOf(s);
        dim3 block(128, 1, 1);
        dim3 stripGrid((L.gy + L.hist_chunk - 1) / L.hist_chunk, L.gx, nsrc);     // one block per history strip and row
        const int* hints = s->hintsValid ? s->firstActive : nullptr;
        const size_t threads = (size_t)L.gx * L.gy * nsrc;
        const bool dense = threads >= (size_t)1 << 20;          // enough threads to fill the GPU several times over
        const bool tiny = threads <= (size_t)96 << 10;          // fewer than the GPU holds at once: latency-bound (70^2 .. 256^2: 5-13 % faster with 32 loads in flight; 512^2: 33 % slower)
        if (L.hist_chunk == kHistChunkDefault && dense)
            encodeResponseKernel<kHistChunkDefault, PVC_AN_DENSE_MINB, PVC_AN_BATCH><<<stripGrid, block, 0, s->stream>>>(L, A, s->hist, s->w, s->src, s->results, s->delay, s->walkDelay, hints);
        else if (L.hist_chunk == kHistChunkDefault && tiny)
            encodeResponseKernel<kHistChunkDefault, 0, 2 * PVC_AN_BATCH><<<stripGrid, block, 0, s->stream>>>(L, A, s->hist, s->w, s->src, s->results, s->delay, s->walkDelay, hints);
        else if (L.hist_chunk == kHistChunkDefault)
            encodeResponseKernel<kHistChunkDefault, 0, PVC_AN_BATCH><<<stripGrid, block, 0, s->stream>>>(L, A, s->hist, s->w, s->src, s->results, s->delay, s->walkDelay, hints);
        else if (L.hist_chunk == kValidCols)
            encodeResponseKernel<kValidCols, 0, PVC_AN_BATCH><<<stripGrid, block, 0, s->stream>>>(L, A, s->hist, s->w, s->src, s->results, s->delay, s->walkDelay, hints);
        else { setError("analyzer: unsupported history strip width %d", L.hist_chunk); return PVC_ERR_INVALID; }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("analyzer launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        *launches += 1;
        return launchListenerDirection(s, nsrc, launches);
    }

    int launchListenerDirection(pvc_solver* s, int nsrc, int* launches)
    {
        const Layout& L = s->L;
        const AnalyzeParams A = paramsOf(s);
        dim3 block(128, 1, 1);
        dim3 grid((L.gy + 127) / 128, L.gx, nsrc);
        if (s->walkSequential)                              // pvc_set_walk_mode: the reference's walk, the cross-check of the tests
        {
            listenerDirectionKernel<<<grid, block, 0, s->stream>>>(L, A, s->src, s->results, s->walkDelay);
            *launches += 1;
        }
        else
        {
            const size_t cells = (size_t)L.gx * L.gy;
            if (cells <= (size_t)kWalkSmallCells)
            {
                int rounds = 1;
                for (long span = 4; span < (long)A.T; span *= 4) ++rounds;
                static bool configured[64] = {};
                if (!configured[s->device & 63])
                {
                    if (cudaFuncSetAttribute(walkSmallKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWalkSmallCells * (int)sizeof(int)) != cudaSuccess)
                    { setError("listener direction: shared-memory opt-in failed: %s", cudaGetErrorString(cudaGetLastError())); return PVC_ERR_CUDA; }
                    configured[s->device & 63] = true;
                }
                walkSmallKernel<<<nsrc, 1024, cells * sizeof(int), s->stream>>>(L, A, s->src, s->results, s->walkDelay, rounds, 3);
                *launches += 1;
                cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) { setError("listener direction launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
                return PVC_OK;
            }
            walkNextKernel<<<grid, block, 0, s->stream>>>(L, A, s->src, s->results, s->walkDelay, s->walkNext);
            // A pass follows up to kHops further links from every cell, so it multiplies the length every link spans by at
            // least kHops + 1 (in place: a link read here may already be longer); delays are integral sample indices that
            // strictly decrease along a walk, so no walk is longer than T hops: ceil(log_(kHops+1) T) passes resolve them all
            // (walkResolveKernel follows whatever could be left, so the count is a matter of speed, not of correctness).
            constexpr int kHops = 3;
            int rounds = 1;
            for (long span = kHops + 1; span < (long)A.T; span *= kHops + 1) ++rounds;
            for (int k = 0; k < rounds; ++k)
                walkJumpKernel<<<dim3((unsigned)((cells + 255) / 256), nsrc), 256, 0, s->stream>>>(cells, kHops, s->walkNext);
            walkResolveKernel<<<grid, block, 0, s->stream>>>(L, A, s->src, s->results, s->walkDelay, s->walkNext);
            *launches += 2 + rounds;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("listener direction launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    static int streamChecks(const pvc_solver* s)
    {
        if (!s->chunkT || !s->carry) { setError("streamed analyzer: the solver was not created with pvc_create_streamed"); return PVC_ERR_INVALID; }
        if (s->L.hist_chunk != kHistChunkDefault) { setError("streamed analyzer: unsupported history strip width %d", s->L.hist_chunk); return PVC_ERR_INVALID; }
        return PVC_OK;
    }

    int launchStreamForward(pvc_solver* s, int nsrc, int base, int len, int* launches)
    {
        const int rc = streamChecks(s);
        if (rc) return rc;
        const Layout& L = s->L;
        const AnalyzeParams A = paramsOf(s);
        const dim3 stripGrid((L.gy + L.hist_chunk - 1) / L.hist_chunk, L.gx, nsrc);
        forwardChunkKernel<kHistChunkDefault><<<stripGrid, 128, 0, s->stream>>>(L, A, s->hist, s->w, s->carry, (size_t)s->cfg.max_sources * L.gx * L.gy, base, len,
                                                                               s->hintsValid ? s->firstActive : nullptr);
        *launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("streamed analyzer (forward) launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    int launchStreamBackward(pvc_solver* s, int nsrc, int base, int len, int* launches)
    {
        const int rc = streamChecks(s);
        if (rc) return rc;
        const Layout& L = s->L;
        const AnalyzeParams A = paramsOf(s);
        const dim3 stripGrid((L.gy + L.hist_chunk - 1) / L.hist_chunk, L.gx, nsrc);
        backwardChunkKernel<kHistChunkDefault><<<stripGrid, 128, 0, s->stream>>>(L, A, s->hist, s->w, s->src, s->carry, (size_t)s->cfg.max_sources * L.gx * L.gy, base, len,
                                                                                s->results, s->delay, s->walkDelay);
        *launches += 1;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("streamed analyzer (backward) launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }

    int launchIrRebuild(pvc_solver* s, int source, int r, int c, float* out_dev)
    {
        const Layout& L = s->L;
        const float* h = s->hist + (size_t)source * L.hist_source;
        rebuildIrKernel<<<1, 32, 0, s->stream>>>(L, s->cfg.T, s->cfg.courant, h, s->w, r, c, out_dev);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { setError("ir rebuild launch: %s", cudaGetErrorString(e)); return PVC_ERR_CUDA; }
        return PVC_OK;
    }
}

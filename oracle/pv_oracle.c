/*
 * pv_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported or called by the product
 * (planeverb_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker.
 *
 * A plain-C, single-precision CPU restatement of the reference's hot path
 *   Grid::GenerateResponseCPU        /root/reference/ProjectPlaneverb/src/FDTD/FDTD.cpp:87-236
 *   Grid ctor / GaussianPulse / AABB /root/reference/ProjectPlaneverb/src/FDTD/Grid.cpp:12-117,136-296,390-396
 *   FreeGrid                         /root/reference/ProjectPlaneverb/src/FDTD/FreeGrid.cpp:41-110
 *   Analyzer                         /root/reference/ProjectPlaneverb/src/DSP/Analyzer.cpp:48-431
 * written from the semantics tables in SURVEY.md App. A/B, not from the reference text.
 *
 * PARITY PIN: tests/test_oracle.py checks every function here bit-for-bit (fields,
 * delays, all eight outputs) against the UNMODIFIED reference compiled in place into
 * oracle/_ref/libpvref.so (oracle/refdriver/Makefile) and against the committed golden vectors in
 * tests/golden/ that were captured from that build (tools/make_golden.py).  The reference itself
 * ships no tests or golden vectors for this path.
 *
 * Differences from the reference that do NOT change values:
 *   - SoA planes instead of AoS Cell; the wall blend of FDTD.cpp:162-168 is written in branch form
 *     (value-identical; only the sign of exact zeros can differ);
 *   - instead of a 16-byte IR per cell per step it keeps one pressure plane per step (hist[t][i])
 *     and accumulates the causal analyzer sums (onset, Edry, flux, wet) while stepping, in the same
 *     ascending-t fp32 order as Analyzer.cpp:182-195,239-243; the anti-causal RT60 integral
 *     (Analyzer.cpp:303-319) walks the pressure history backwards exactly as the reference does;
 *   - analysis windows that would run past the end of the IR (reference reads out of bounds,
 *     Analyzer.cpp:183-195) are clamped to T; such cells are flagged so tests can exclude them.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  No fast-math: all
 * float->int truncations and every rounding step must match the strict x86-64 SSE reference build.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* PvTypes.h:83-101 constants (values are part of the public contract) */
#define PVO_C            343.21f
#define PVO_PPW          3.5f
#define PVO_THRESH       0.00000316f
#define PVO_DRY_DIR_S    0.005f
#define PVO_DRY_GAIN_S   0.01f
#define PVO_WET_S        0.080f
#define PVO_SCHROEDER_S  0.01f
#define PVO_GAIN_THRESH  0.891251f
#define PVO_DELAY_CLOSE  5.f
#define PVO_SQRT2        1.4142136f
static const float PVO_IR_SECONDS = PVO_SQRT2 * 12.5f / PVO_C + 0.25f;   /* PvTypes.h:101 */

/* ---- Grid.cpp:390-396 ---- */
void pvo_grid_params(int resolution, float* dx, float* dt, unsigned* fs)
{
    float minWavelength = PVO_C / (float)resolution;
    *dx = minWavelength / PVO_PPW;
    *dt = *dx / (PVO_C * 1.5f);
    *fs = (unsigned)(1.0f / *dt);
}

/* Grid.cpp:48-49,55 and FDTD.cpp:90 : derived sizes.  out = {gx, gy, T}; courant returned */
float pvo_derived(int resolution, float sizeX, float sizeY, int* out)
{
    float dx, dt; unsigned fs;
    pvo_grid_params(resolution, &dx, &dt, &fs);
    float gsx = (1.f / dx) * sizeX;
    float gsy = (1.f / dx) * sizeY;
    out[0] = (int)gsx;
    out[1] = (int)gsy;
    out[2] = (int)(unsigned)((float)fs * PVO_IR_SECONDS);
    return PVO_C * dt / dx;
}

/* ---- Grid.cpp:12-27 (mixed float/double expression kept as written there) ---- */
void pvo_gaussian_pulse(int resolution, float samplingRate, float* out, unsigned n)
{
    const float maxFreq = (float)resolution;
    const float pi = (float)acos(-1.0);
    float sigma = (float)(1.0f / (0.5 * (double)pi * (double)maxFreq));
    const float delay = 2 * sigma;
    const float dt = 1.0f / samplingRate;
    for (unsigned i = 0; i < n; ++i)
    {
        float t = (float)i * dt;
        out[i] = expf(-(t - delay) * (t - delay) / (sigma * sigma));
    }
}

/* ---- Grid.cpp:84-108 : b = 0 on the padding row/col, absorption R = 0 everywhere ---- */
void pvo_coef_init(int gx, int gy, int16_t* b, float* R)
{
    int S = gy + 1;
    for (int r = 0; r <= gx; ++r)
        for (int c = 0; c <= gy; ++c)
        {
            b[r * S + c] = (r == gx || c == gy) ? 0 : 1;
            R[r * S + c] = 0.f;
        }
}

/* rectangle of Grid.cpp:139-142 / 252-255: rows from (pos.x -/+ width/2), cols from (pos.y -/+ height/2),
 * multiply-by-reciprocal then truncate toward zero */
static void aabb_span(float dx, float px, float py, float w, float h, int* r0, int* r1, int* c0, int* c1)
{
    float inv = 1.f / dx;
    *c0 = (int)((py - h / 2.f + 0.f) * inv);
    *r0 = (int)((px - w / 2.f + 0.f) * inv);
    *c1 = (int)((py + h / 2.f + 0.f) * inv);
    *r1 = (int)((px + w / 2.f + 0.f) * inv);
}

/* Grid.cpp:229-246 (square grids; the reference mixes strides otherwise) */
void pvo_add_aabb(int gx, int gy, float dx, int16_t* b, float* R,
                  float px, float py, float w, float h, float absorption)
{
    int r0, r1, c0, c1, S = gy + 1;
    aabb_span(dx, px, py, w, h, &r0, &r1, &c0, &c1);
    for (int c = c0; c < c1; ++c)
    {
        if (c < 0 || c > gy) continue;
        for (int r = r0; r < r1; ++r)
        {
            if (r < 0 || r > gx) continue;
            b[r * S + c] = 0;
            R[r * S + c] = absorption;
        }
    }
}

/* Grid.cpp:249-296: no overlap ref-counting; padding row/col stays b = 0 */
void pvo_remove_aabb(int gx, int gy, float dx, int16_t* b, float* R,
                     float px, float py, float w, float h)
{
    int r0, r1, c0, c1, S = gy + 1;
    aabb_span(dx, px, py, w, h, &r0, &r1, &c0, &c1);
    for (int c = c0; c < c1; ++c)
    {
        if (c < 0 || c > gy) continue;
        for (int r = r0; r < r1; ++r)
        {
            if (r < 0 || r > gx) continue;
            R[r * S + c] = 0.f;
            b[r * S + c] = (c == gx || r == gy) ? 0 : 1;   /* Grid.cpp:276 compares col with size.x, row with size.y */
        }
    }
}

/* FDTD.cpp:97-99 */
void pvo_listener_cell(float dx, float lx, float lz, int* lr, int* lc)
{
    *lr = (int)((lx + 0.f) / dx);
    *lc = (int)((lz + 0.f) / dx);
}

/* ---- one velocity update in branch form (FDTD.cpp:149-168 / 178-197) ---- */
static inline float vel_update(float v, float p_this, float p_prev, int b_this, int b_prev,
                               float R_this, float R_prev, float courant)
{
    if (b_this && b_prev) return v - courant * (p_this - p_prev);
    if (!b_this && b_prev) { float Y = (1.f - R_this) / (1.f + R_this); return Y * p_prev; }
    if (b_this && !b_prev) { float Yn = (1.f - R_prev) / (1.f + R_prev); return -(Yn * p_this); }
    return 0.f;
}

/* per-cell causal analyzer accumulators kept while stepping (Analyzer.cpp:146-195,235-243) */
typedef struct
{
    int*   onset;   /* -1 until found */
    float* edry;
    float* fx;
    float* fy;
    float* wet;
} pvo_causal;

/*
 * FDTD.cpp:110-235.  State planes p,vx,vy of N=(gx+1)*(gy+1) floats are zeroed here.
 *   hist   : optional T*N pressure planes, hist[t*N+i] = recorded p of sample t (FDTD.cpp:226-231)
 *   hvx,hvy: optional T*N velocity planes (small cases only; for IR parity)
 *   acc    : optional causal accumulators over ALL alloc cells (arrays of N), windows in samples
 *            Sd (flux), D (dry), W (wet) as in Analyzer.cpp:170-173,237
 */
static void simulate_impl(int gx, int gy, const int16_t* b, const float* R, float courant,
                  int li, const float* pulse, int T,
                  float* p, float* vx, float* vy,
                  float* hist, size_t histI0, size_t histCount, float* hvx, float* hvy,
                  int* onset, float* edry, float* fx, float* fy, float* wet,
                  int Sd, int D, int W)
{
    const int S = gy + 1;
    const int N = (gx + 1) * S;
    memset(p, 0, sizeof(float) * N);
    memset(vx, 0, sizeof(float) * N);
    memset(vy, 0, sizeof(float) * N);
    if (onset)
        for (int i = 0; i < N; ++i) { onset[i] = -1; edry[i] = fx[i] = fy[i] = wet[i] = 0.f; }

    for (int t = 0; t < T; ++t)
    {
        /* 1. pressure (FDTD.cpp:125-141).  b = 0 wherever i+S or the wrapped i+1 would matter. */
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i)
        {
            if (b[i])
            {
                float div = (vx[i + S] - vx[i]) + (vy[i + 1] - vy[i]);
                p[i] = p[i] - courant * div;
            }
            else p[i] = 0.f;
        }
        /* 2. vx, i >= S, previous = one row up (FDTD.cpp:144-170) */
        #pragma omp parallel for schedule(static)
        for (int i = S; i < N; ++i)
            vx[i] = vel_update(vx[i], p[i], p[i - S], b[i], b[i - S], R[i], R[i - S], courant);
        /* 3. vy, i >= 1, previous = i-1 incl. the wrap onto the previous row's padding cell (FDTD.cpp:173-199) */
        #pragma omp parallel for schedule(static)
        for (int i = 1; i < N; ++i)
            vy[i] = vel_update(vy[i], p[i], p[i - 1], b[i], b[i - 1], R[i], R[i - 1], courant);
        /* 4. grid-edge absorbing overrides (FDTD.cpp:202-223) */
        for (int c = 0; c < gy; ++c)
        {
            vx[c] = -p[c];
            vx[gx * S + c] = p[(gx - 1) * S + c];
        }
        for (int r = 0; r < gx; ++r)
        {
            vy[r * S] = -p[r * S];
            vy[r * S + gy] = p[r * S + gy - 1];
        }
        /* 5. record sample t (FDTD.cpp:226-231) */
        if (hist) memcpy(hist + (size_t)t * histCount, p + histI0, sizeof(float) * histCount);
        if (hvx)  memcpy(hvx + (size_t)t * N, vx, sizeof(float) * N);
        if (hvy)  memcpy(hvy + (size_t)t * N, vy, sizeof(float) * N);
        if (onset)
        {
            #pragma omp parallel for schedule(static)
            for (int i = 0; i < N; ++i)
            {
                float pr = p[i];
                if (onset[i] < 0 && fabsf(pr) > PVO_THRESH) onset[i] = t;      /* Analyzer.cpp:146-154 */
                int on = onset[i];
                /* sums start at sample 0, not at the onset (Analyzer.cpp:182-195) */
                if (on < 0 || t < on + Sd) { fx[i] += pr * vx[i]; fy[i] += pr * vy[i]; }
                if (on < 0 || t < on + D) edry[i] += pr * pr;
                else if (t > on + D && t < on + D + 1 + W) wet[i] += pr * pr;  /* Analyzer.cpp:239-243 */
            }
        }
        /* 6. inject after the record (FDTD.cpp:234) */
        p[li] += pulse[t];
    }
}

void pvo_simulate(int gx, int gy, const int16_t* b, const float* R, float courant,
                  int li, const float* pulse, int T,
                  float* p, float* vx, float* vy,
                  float* hist, float* hvx, float* hvy,
                  int* onset, float* edry, float* fx, float* fy, float* wet,
                  int Sd, int D, int W)
{
    simulate_impl(gx, gy, b, R, courant, li, pulse, T, p, vx, vy, hist, 0, (size_t)(gx + 1) * (gy + 1), hvx, hvy,
                  onset, edry, fx, fy, wet, Sd, D, W);
}

/* Same simulation, keeping the pressure history of alloc cells [histI0, histI0 + histCount) only (hist[t*histCount + i - histI0]):
 * sizes whose full T*N history does not fit the host (2048 x 2048 x 4000 steps = 67 GB).  The causal accumulators cover every
 * cell as before; pvo_encode_band then computes the anti-causal RT60 for the cells of the band. */
void pvo_simulate_band(int gx, int gy, const int16_t* b, const float* R, float courant,
                       int li, const float* pulse, int T,
                       float* p, float* vx, float* vy,
                       float* hist, long long histI0, long long histCount,
                       int* onset, float* edry, float* fx, float* fy, float* wet,
                       int Sd, int D, int W)
{
    simulate_impl(gx, gy, b, R, courant, li, pulse, T, p, vx, vy, hist, (size_t)histI0, (size_t)histCount, NULL, NULL,
                  onset, edry, fx, fy, wet, Sd, D, W);
}

/* FreeGrid.cpp:71-110: sum p^2 of the first n samples at the probe cell, times r = (int)(1/dx)*dx.
 * Runs its own n-step free-field simulation on an empty (gx, gy) grid (only the first n samples of
 * the probe are ever read, and they do not depend on later steps). */
float pvo_efree(int resolution, int gx, int gy)
{
    float dx, dt; unsigned fs;
    pvo_grid_params(resolution, &dx, &dt, &fs);
    const float courant = PVO_C * dt / dx;
    const int S = gy + 1, N = (gx + 1) * S;
    int lX = gx / 2, lY = gy / 2;
    int eX = lX + (int)(1.f / dx), eY = lY;
    int n = (int)(PVO_DRY_GAIN_S * (float)(int)fs) + (int)((1.f / PVO_C) * (float)(int)fs);
    int lr, lc;
    pvo_listener_cell(dx, (float)lX * dx, (float)lY * dx, &lr, &lc);   /* FreeGrid.cpp:84 -> FDTD.cpp:97-98 */

    int16_t* b = malloc(sizeof(int16_t) * N);
    float* R = malloc(sizeof(float) * N);
    float* p = malloc(sizeof(float) * (N + S + 1));
    float* vx = calloc(N + S + 1, sizeof(float));
    float* vy = calloc(N + S + 1, sizeof(float));
    float* pulse = malloc(sizeof(float) * n);
    float* hist = malloc(sizeof(float) * (size_t)n * N);
    pvo_coef_init(gx, gy, b, R);
    pvo_gaussian_pulse(resolution, (float)fs, pulse, n);
    pvo_simulate(gx, gy, b, R, courant, lr * S + lc, pulse, n, p, vx, vy, hist, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    float e = 0.f;
    int ei = eX * S + eY;
    for (int i = 0; i < n; ++i) { float v = hist[(size_t)i * N + ei]; e += v * v; }
    float r = (float)(eX - lX) * dx;
    e *= r;
    free(b); free(R); free(p); free(vx); free(vy); free(pulse); free(hist);
    return e;
}

/* FreeGrid.cpp:41-59 */
float pvo_efree_per_r(float efree, float dx, int lX, int lY, int eX, int eY)
{
    float lx = (float)lX * dx, ly = (float)lY * dx, ex = (float)eX * dx, ey = (float)eY * dx;
    float r = sqrtf((ex - lx) * (ex - lx) + (ey - ly) * (ey - ly));
    if (r == 0.f) return efree;
    return efree / r;
}

/*
 * Analyzer.cpp:139-328 from the causal accumulators + pressure history.
 * results: gx*gy*8 floats {occlusion, wetGain, rt60, lowpass, dir.x, dir.y, srcDir.x, srcDir.y}; only
 * cells with an onset are written (no-onset cells keep what the caller put there, Analyzer.cpp:161-165).
 * delay: gx*gy floats. clamped: optional gx*gy bytes, 1 where onset+D >= T (reference reads out of bounds).
 */
static void encode_impl(int gx, int gy, int T, int fs, float dx, float efree, float lx, float lz,
                const float* histBand, size_t histI0, size_t histCount, const int* onset, const float* edry, const float* fx,
                const float* fy, const float* wet, float* results, float* delay, uint8_t* clamped)
{
    const int S = gy + 1;
    const size_t N = histCount;
    const int D = (int)(PVO_DRY_GAIN_S * (float)fs);
    const int listenerX = (int)(lx * (1.f / dx));      /* Analyzer.cpp:201-202 */
    const int listenerY = (int)(lz * (1.f / dx));
    const int cells = gx * gy;

    #pragma omp parallel for schedule(dynamic, 64)
    for (int s = 0; s < cells; ++s)
    {
        int r = s / gx, c = s % gx;            /* INDEX_TO_POS, PvDefinitions.h:24 */
        int i = r * S + c;                     /* FDTD.cpp:76-77 */
        float* out = results + (size_t)s * 8;
        int on = onset[i];
        if (clamped) clamped[s] = 0;
        if (on < 0) { delay[s] = FLT_MAX; continue; }
        delay[s] = (float)on;
        int directEnd = on + D;
        if (clamped && directEnd >= T) clamped[s] = 1;

        /* obstruction + source directivity (Analyzer.cpp:199-220) */
        float efreePr = pvo_efree_per_r(efree, dx, listenerX, listenerY, r, c);
        float E = edry[i] / efreePr;
        float occ = sqrtf(E);
        float rx = fx[i], ry = fy[i];
        float norm = sqrtf(rx * rx + ry * ry);
        norm = -1.0f / (norm > 0.0f ? norm : 1.0f);
        out[0] = occ;
        out[6] = norm * rx;
        out[7] = norm * ry;

        /* low-pass cutoff (Analyzer.cpp:227-230) */
        float rr = 1.0f / fmaxf(0.001f, occ);
        out[3] = -147.f + 18390.f / (1.f + powf(rr / 12.f, 0.8f));

        /* wet gain (Analyzer.cpp:247) */
        out[1] = sqrtf(wet[i] / efree);

        /* RT60 (Analyzer.cpp:282-326); outside the recorded band: NaN = "not computed" */
        if ((size_t)i < histI0 || (size_t)i >= histI0 + histCount) { out[2] = NAN; continue; }
        const float* hist = histBand + ((size_t)i - histI0) - (size_t)i;      /* hist[k*N + i] below addresses the band */
        int start = directEnd + 1;
        int end = T - (int)(PVO_SCHROEDER_S * fs);
        int regressN = end - start;
        float rn = (float)regressN;
        float xmean = (rn - 1.0f) * 0.5f;
        float xsum = rn * xmean;
        float denominator = (1.0f / 12.0f) * rn * (rn * rn - 1.0f);
        float edc = 0.f, xysum = 0.f, ysum = 0.f;
        for (int k = T - 1; k >= end && k >= 0; --k) { float v = hist[(size_t)k * N + i]; edc += v * v; }
        for (int k = end - 1; k >= start; --k)
        {
            float v = hist[(size_t)k * N + i];
            edc += v * v;
            float y = 10.f * log10f(edc);
            xysum += y * (k - start);
            ysum += y;
        }
        float ymean = ysum / rn;
        float numerator = xysum - ymean * xsum - xmean * ysum + rn * xmean * ymean;
        float slopeDBperSample = numerator / denominator;
        float slopeDBperSec = slopeDBperSample * fs;
        out[2] = -60.f / slopeDBperSec;
    }
}

void pvo_encode(int gx, int gy, int T, int fs, float dx, float efree, float lx, float lz,
                const float* hist, const int* onset, const float* edry, const float* fx,
                const float* fy, const float* wet, float* results, float* delay, uint8_t* clamped)
{
    encode_impl(gx, gy, T, fs, dx, efree, lx, lz, hist, 0, (size_t)(gx + 1) * (gy + 1), onset, edry, fx, fy, wet, results, delay, clamped);
}

/* pvo_encode over a banded history (pvo_simulate_band): every output but RT60 for every cell, RT60 for the cells of the band,
 * NaN elsewhere */
void pvo_encode_band(int gx, int gy, int T, int fs, float dx, float efree, float lx, float lz,
                     const float* hist, long long histI0, long long histCount, const int* onset, const float* edry, const float* fx,
                     const float* fy, const float* wet, float* results, float* delay, uint8_t* clamped)
{
    encode_impl(gx, gy, T, fs, dx, efree, lx, lz, hist, (size_t)histI0, (size_t)histCount, onset, edry, fx, fy, wet, results, delay, clamped);
}

/* Analyzer.cpp:340-431 for every interior cell (needs all occlusion/delay values first) */
void pvo_directions(int gx, int gy, int T, int fs, int resolution, float dx, float lx, float lz,
                    float* results, const float* delaySamples)
{
    static const int NB[8][2] = { {-1,-1},{-1,0},{-1,1},{0,-1},{0,1},{1,-1},{1,0},{1,1} };
    const int cells = gx * gy;
    const float samplingRate = (float)fs;
    const float wavelength = PVO_C / (float)resolution;
    const float thresholdDist = 0.3f * wavelength;
    float* dirs = malloc(sizeof(float) * 2 * (size_t)cells);

    #pragma omp parallel for schedule(dynamic, 64)
    for (int s = 0; s < cells; ++s)
    {
        float loudness = results[(size_t)s * 8];
        int nextIndex = s;
        float delay = FLT_MAX;
        while (delay > PVO_DELAY_CLOSE && loudness < PVO_GAIN_THRESH)
        {
            int r = nextIndex / gx, c = nextIndex % gx;
            float nextLoudness = 0.f, nextDelay = FLT_MAX;
            for (int k = 0; k < 8; ++k)
            {
                int nr = r + NB[k][0], nc = c + NB[k][1];
                if (nr < 0 || nc < 0 || nr >= gx || nc >= gy) continue;
                int ni = nr * gx + nc;
                float occ = results[(size_t)ni * 8];
                float d = delaySamples[ni];
                /* Analyzer.cpp:372: (unsigned)delay == numSamples can never hold for a recorded onset
                 * (< T) and FLT_MAX->unsigned is not T on x86-64; occlusion == 0 is the live test */
                if (occ == 0.f) continue;
                if (d < nextDelay && occ > 0.f) { nextLoudness = occ; nextIndex = ni; nextDelay = d; }
            }
            if (nextDelay == FLT_MAX || nextDelay >= delay) break;
            delay = nextDelay;
            loudness = nextLoudness;
            float geodesic = PVO_C * nextDelay / samplingRate;
            int r2 = nextIndex / gx, c2 = nextIndex % gx;
            float ex = (float)r2 * dx, ey = (float)c2 * dx;
            float tx = ex - lx, ty = ey - lz;
            float eu = sqrtf(tx * tx + ty * ty);
            if (fabsf(geodesic - eu) < thresholdDist) break;
        }
        int r = nextIndex / gx, c = nextIndex % gx;
        float ox = (float)r * dx - lx, oy = (float)c * dx - lz;
        float len = ox * ox + oy * oy;
        if (len != 0.f) { len = sqrtf(len); ox /= len; oy /= len; }
        dirs[2 * (size_t)s] = ox; dirs[2 * (size_t)s + 1] = oy;
    }
    /* the reference writes direction in place while later cells only read occlusion/delay (Analyzer.cpp:91-103) */
    for (int s = 0; s < cells; ++s) { results[(size_t)s * 8 + 4] = dirs[2 * (size_t)s]; results[(size_t)s * 8 + 5] = dirs[2 * (size_t)s + 1]; }
    free(dirs);
}

/* analysis window lengths in samples (Analyzer.cpp:170-171,237,284): out = {Sd, D, W, tail} */
void pvo_windows(int fs, int* out)
{
    out[0] = (int)(PVO_DRY_DIR_S * (float)fs);
    out[1] = (int)(PVO_DRY_GAIN_S * (float)fs);
    out[2] = (int)(PVO_WET_S * (float)fs);
    out[3] = (int)(PVO_SCHROEDER_S * fs);
}

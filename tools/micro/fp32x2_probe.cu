// Probe: issue / pipe throughput of packed fp32x2 arithmetic (FADD2 / FFMA2) against scalar FADD / FMUL on sm_100a, and
// bit-equality of  __ffma2_rn(a, b, -0)  with  __fmul_rn  (ptxas 12.9 contracts mul.rn.f32x2 + sub.rn.f32x2 into FFMA2 even
// under --fmad false, so the packed multiply of the step kernel is written as an FMA with a -0 addend).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3 -o fp32x2_probe fp32x2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    float2 r;
    asm("{ .reg .b64 ra, rb, rr; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rr, ra, rb; mov.b64 {%0, %1}, rr; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// negz = -0.0f passed at run time: ptxas must not see the constant, or it rewrites fma(a, b, -0) as a multiply and contracts it
// with the subtraction that follows into one FFMA2 (single rounding) -- even under --fmad false
template <int MODE>
__global__ void __launch_bounds__(512, 1) spin(float* out, int iters, float c, float negz)
{
    float2 a[8], b[8];
    uint32_t z[8];
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); b[i] = make_float2(1e-3f * i, 2e-3f); z[i] = threadIdx.x + i; }
    const float2 cc = make_float2(c, c), nz = make_float2(negz, negz);
    for (int it = 0; it < iters; ++it)
    {
        #pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            if (MODE == 0 || MODE == 2)      // scalar: sub, mul, sub per element (6 instructions per pair)
            {
                const float dx = __fsub_rn(a[i].x, b[i].x), dy = __fsub_rn(a[i].y, b[i].y);
                b[i].x = __fsub_rn(b[i].x, __fmul_rn(c, dx)); b[i].y = __fsub_rn(b[i].y, __fmul_rn(c, dy));
                a[i].x = __fadd_rn(a[i].x, b[(i + 1) & 7].x); a[i].y = __fadd_rn(a[i].y, b[(i + 1) & 7].y);
            }
            else                             // packed: 4 instructions per pair
            {
                const float2 d = sub2(a[i], b[i]);
                const float2 m = __ffma2_rn(cc, d, nz);
                b[i] = sub2(b[i], m);
                a[i] = __fadd2_rn(a[i], b[(i + 1) & 7]);
            }
            if (MODE >= 2) { z[i] = (z[i] ^ (z[(i + 3) & 7] >> 3)) + 0x9e37u; z[(i + 5) & 7] ^= z[i] << 1; }   // integer side work
        }
    }
    float s = 0.f; uint32_t zz = 0;
    for (int i = 0; i < 8; ++i) { s += a[i].x + a[i].y + b[i].x + b[i].y; zz ^= z[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(zz & 0xff);
}

__global__ void exactness(const float* x, const float* y, uint32_t* bad, int n, float negz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 >= n) return;
    const float2 a = make_float2(x[2 * i], x[2 * i + 1]), b = make_float2(y[2 * i], y[2 * i + 1]);
    const float2 m = __ffma2_rn(a, b, make_float2(negz, negz));
    const float2 s = sub2(a, b);
    const float2 f = sub2(a, __ffma2_rn(b, b, make_float2(negz, negz)));     // the step kernel's  x - C * d  shape
    const float2 t = __fadd2_rn(a, b);
    uint32_t e = 0;
    e += __float_as_uint(m.x) != __float_as_uint(__fmul_rn(a.x, b.x));
    e += __float_as_uint(m.y) != __float_as_uint(__fmul_rn(a.y, b.y));
    e += __float_as_uint(s.x) != __float_as_uint(__fsub_rn(a.x, b.x));
    e += __float_as_uint(s.y) != __float_as_uint(__fsub_rn(a.y, b.y));
    e += __float_as_uint(t.x) != __float_as_uint(__fadd_rn(a.x, b.x));
    e += __float_as_uint(t.y) != __float_as_uint(__fadd_rn(a.y, b.y));
    e += __float_as_uint(f.x) != __float_as_uint(__fsub_rn(a.x, __fmul_rn(b.x, b.x)));
    e += __float_as_uint(f.y) != __float_as_uint(__fsub_rn(a.y, __fmul_rn(b.y, b.y)));
    if (e) atomicAdd(bad, e);
}

template <int MODE> float run(float* out, int iters)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spin<MODE><<<148, 512>>>(out, 1000, 0.66f, -0.f);
    cudaEventRecord(e0);
    spin<MODE><<<148, 512>>>(out, iters, 0.66f, -0.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 512 * 4);
    const int iters = 200000;
    const char* names[4] = { "scalar FADD/FMUL", "packed FADD2/FFMA2", "scalar + int work", "packed + int work" };
    float ms[4] = { run<0>(out, iters), run<1>(out, iters), run<2>(out, iters), run<3>(out, iters) };
    for (int m = 0; m < 4; ++m)
        printf("%-22s %8.3f ms  %.2f pair-updates/clk/SM (at 1.965 GHz)\n", names[m], ms[m], 8.0 * iters * 512 / (ms[m] * 1e-3 * 1.965e9));
    // exactness over random bit patterns, small / denormal / zero / signed-zero operands
    const int n = 1 << 24;
    float* hx = new float[n]; float* hy = new float[n];
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < n; ++i)
    {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17; uint32_t u = (uint32_t)s, v = (uint32_t)(s >> 32);
        if ((i & 7) == 0) u &= 0x807fffffu;                       // denormal
        if ((i & 15) == 1) u &= 0x80000000u;                      // signed zero
        if ((i & 7) == 2) v = (v & 0x807fffffu) | 0x00800000u;    // tiny normal
        if ((i & 3) == 3) { u = (u & 0x80ffffffu) | 0x3f000000u; v = (v & 0x80ffffffu) | 0x3e000000u; }   // O(1) values
        memcpy(hx + i, &u, 4); memcpy(hy + i, &v, 4);
    }
    float *dx, *dy; uint32_t* bad; cudaMalloc(&dx, n * 4); cudaMalloc(&dy, n * 4); cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
    cudaMemcpy(dx, hx, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dy, hy, n * 4, cudaMemcpyHostToDevice);
    exactness<<<(n / 2 + 255) / 256, 256>>>(dx, dy, bad, n, -0.f);
    uint32_t hb = 1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("exactness: %u mismatching results over %d pairs (NaN payloads included)  err=%s\n", hb, n / 2, cudaGetErrorString(cudaGetLastError()));
    return 0;
}

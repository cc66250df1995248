// dp_pipe_probe.cu -- issue cost of the analyzer's expensive instructions on sm_100a: DFMA, F2F.F32.F64 (double -> float), and the
// integer round-to-nearest-even that could replace the conversion.  8 independent chains per thread, 1024 threads per SM, all SMs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dp_pipe_probe dp_pipe_probe.cu && ./dp_pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(double* out, int iters, double seed)
{
    double a[8]; float f[8]; unsigned u[8];
    #pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = seed + k + threadIdx.x * 1e-3; f[k] = (float)a[k]; u[k] = threadIdx.x + k; }
    for (int i = 0; i < iters; ++i)
    {
        #pragma unroll
        for (int k = 0; k < 8; ++k)
        {
            if (MODE == 0) a[k] = __fma_rn(a[k], 1.0000001, 1e-9);                       // DFMA
            if (MODE == 1) { f[k] = (float)a[k]; a[k] = __hiloint2double(__double2hiint(a[k]) ^ (__float_as_int(f[k]) & 1), __double2loint(a[k])); }   // F2F.F32.F64 + 2 int
            if (MODE == 2) f[k] = __fmaf_rn(f[k], 1.0000001f, 1e-9f);                    // FFMA
            if (MODE == 3)
            {   // integer RNE of a normal double to float (sign / exponent re-bias / 29 dropped bits)
                const unsigned hi = (unsigned)__double2hiint(a[k]), lo = (unsigned)__double2loint(a[k]);
                unsigned m = ((hi & 0x7fffffffu) - 0x38000000u) << 3 | (lo >> 29);
                const unsigned rem = lo & 0x1fffffffu;
                m += (rem > 0x10000000u) || (rem == 0x10000000u && (m & 1u));
                u[k] = m | (hi & 0x80000000u);
                a[k] = __hiloint2double((int)(hi ^ (u[k] & 1)), (int)lo);
            }
            if (MODE == 4) { a[k] = __fma_rn(a[k], 1.0000001, 1e-9); f[k] = __fmaf_rn(f[k], 1.0000001f, 1e-9f); f[k] = __fmaf_rn(f[k], 1.0000001f, 1e-9f); }  // 1 DFMA + 2 FFMA
        }
    }
    double s = 0;
    #pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] + f[k] + u[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double perIter)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms, 1024>>>(out, 64, 1.0);
    cudaEventRecord(e0);
    probe<MODE><<<sms, 1024>>>(out, iters, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double warpInstr = 32.0 * iters * 8 * perIter;              // per SM: 32 warps
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-44s %.3f ms  %.2f issue cycles per warp instruction per scheduler (%.1f lanes/clk/SM)\n", name, ms, cycles / (warpInstr / 4), warpInstr * 32 / cycles);
    cudaFree(out);
}

int main()
{
    run<0>("DFMA", 1);
    run<1>("F2F.F32.F64 (+2 int, counted as 1)", 1);
    run<2>("FFMA", 1);
    run<3>("integer RNE double->float (counted as 1)", 1);
    run<4>("1 DFMA + 2 FFMA (counted as 3)", 3);
    return 0;
}

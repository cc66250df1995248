// vec16_atomicity.cu -- stress test of the assumption behind the resident step kernel's mailbox (pvc_step_res.cu): an aligned
// 16-byte st.relaxed.gpu.global.v4 is observed by an aligned 16-byte ld.relaxed.gpu.global.v4 either entirely or not at all.
// Writer CTAs keep overwriting an array of 16-byte words with {n, n, n, n} for increasing n, reader CTAs on OTHER SMs keep loading
// them and count words whose four lanes disagree ("torn").  The PTX memory model only promises this for scalar accesses; the
// hardware performs an aligned 128-bit access of one thread as one request inside one 32-byte sector.
// Second pass: the kernel now mails two words per 256-bit access (st / ld .v8, sm_100).  It needs no more than 16-byte
// indivisibility of each half: writers store {n, n, n, n, m, m, m, m} with m = ~n, readers check each half by itself.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/vec16_atomicity tools/micro/vec16_atomicity.cu
//   tools/micro/vec16_atomicity [seconds]
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cuda_runtime.h>

template <bool PAIR>
__global__ void hammer(float4* words, int nWords, volatile int* stop, unsigned long long* torn, unsigned long long* reads, unsigned long long* fresh)
{
    const bool writer = (blockIdx.x & 1) == 0;
    const int tid = (blockIdx.x >> 1) * blockDim.x + threadIdx.x;
    const int stride = (gridDim.x >> 1) * blockDim.x;
    unsigned long long myTorn = 0, myReads = 0, myFresh = 0;
    unsigned n = 1, last = 0;
    while (!*stop)
    {
        for (int i = tid; i < nWords; i += stride)
        {
            if (PAIR)
            {
                if (i & 1) continue;
                if (writer)
                {
                    const float f = __uint_as_float(n), g = __uint_as_float(~n);
                    asm volatile("st.relaxed.gpu.global.v8.f32 [%0], {%1, %1, %1, %1, %2, %2, %2, %2};" ::"l"(words + i), "f"(f), "f"(g) : "memory");
                }
                else
                {
                    float4 v, w;
                    asm volatile("ld.relaxed.gpu.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w), "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "l"(words + i) : "memory");
                    const unsigned a = __float_as_uint(v.x), b = __float_as_uint(v.y), c = __float_as_uint(v.z), d = __float_as_uint(v.w);
                    const unsigned e = __float_as_uint(w.x), f = __float_as_uint(w.y), g = __float_as_uint(w.z), h = __float_as_uint(w.w);
                    if (a != b || a != c || a != d) ++myTorn;
                    if (e != f || e != g || e != h) ++myTorn;
                    if (a != last) { ++myFresh; last = a; }
                    myReads += 2;
                }
                continue;
            }
            if (writer)
            {
                const float f = __uint_as_float(n);
                asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(words + i), "f"(f) : "memory");
            }
            else
            {
                float4 v;
                asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(words + i) : "memory");
                const unsigned a = __float_as_uint(v.x), b = __float_as_uint(v.y), c = __float_as_uint(v.z), d = __float_as_uint(v.w);
                if (a != b || a != c || a != d) ++myTorn;
                if (a != last) { ++myFresh; last = a; }
                ++myReads;
            }
        }
        ++n;
    }
    if (!writer) { atomicAdd(torn, myTorn); atomicAdd(reads, myReads); atomicAdd(fresh, myFresh); }
}

int main(int argc, char** argv)
{
    const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
    const int nWords = 1 << 14;                         // 256 KB: L2 resident, every word contended
    float4* words; int* stop; unsigned long long* counters;
    cudaMalloc(&words, sizeof(float4) * nWords); cudaMemset(words, 0, sizeof(float4) * nWords);
    cudaMallocHost(&stop, sizeof(int)); *stop = 0;
    cudaMalloc(&counters, 3 * sizeof(unsigned long long)); cudaMemset(counters, 0, 3 * sizeof(unsigned long long));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    for (int pass = 0; pass < 2; ++pass)
    {
        *stop = 0;
        cudaMemset(counters, 0, 3 * sizeof(unsigned long long));
        cudaMemset(words, 0, sizeof(float4) * nWords);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        if (pass == 0) hammer<false><<<sms & ~1, 256>>>(words, nWords, stop, counters, counters + 1, counters + 2);
        else hammer<true><<<sms & ~1, 256>>>(words, nWords, stop, counters, counters + 1, counters + 2);
        cudaEventRecord(e1);
        double waited = 0.0;
        while (cudaEventQuery(e1) == cudaErrorNotReady)
        {
            struct timespec ts = { 0, 20000000 };
            nanosleep(&ts, nullptr);
            waited += 0.02;
            if (waited >= seconds) *stop = 1;
        }
        unsigned long long h[3];
        cudaMemcpy(h, counters, sizeof(h), cudaMemcpyDeviceToHost);
        const cudaError_t err = cudaGetLastError();
        printf("vec16 atomicity (%s accesses): %d SMs (%d writer / %d reader CTAs), %.1f s: %llu 16-byte words loaded, %llu saw a new value, %llu TORN%s\n",
               pass == 0 ? "128-bit" : "256-bit", sms, (sms & ~1) / 2, (sms & ~1) / 2, seconds, h[1], h[2], h[0], err == cudaSuccess ? "" : "  (CUDA error!)");
        if (!(h[0] == 0 && err == cudaSuccess && h[1] > 0)) rc = 1;
    }
    return rc;
}

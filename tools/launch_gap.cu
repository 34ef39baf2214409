// Development micro-benchmark: time per back-to-back launch of a one-block kernel on one stream, as a function of its
// dynamic shared memory, register count, and whether it touches mapped host memory.  nvcc -arch=sm_100a -O3 launch_gap.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 1) k_small(unsigned long long* out, volatile unsigned long long* host, int work)
{
    extern __shared__ unsigned char sm[];
    unsigned long long acc = 0;
    for (int i = threadIdx.x; i < work; i += blockDim.x) { sm[i & 1023] = (unsigned char)i; acc += sm[(i * 7) & 1023]; }
    if (host && threadIdx.x == 0) { acc += *host; __threadfence_system(); *host = acc + 1; }
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}
int main()
{
    unsigned long long *d, *h;
    cudaMalloc(&d, 1024);
    cudaHostAlloc(&h, 64, cudaHostAllocMapped);
    *h = 0;
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int smem : {1024, 48 * 1024, 102 * 1024, 200 * 1024})
        for (int host = 0; host < 2; ++host)
            for (int work : {0, 100000}) {
                cudaFuncSetAttribute(k_small, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                for (int w = 0; w < 20; ++w) k_small<<<1, 128, smem, st>>>(d, host ? h : nullptr, work);
                cudaStreamSynchronize(st);
                const int N = 500;
                cudaEventRecord(a, st);
                for (int i = 0; i < N; ++i) k_small<<<1, 128, smem, st>>>(d, host ? h : nullptr, work);
                cudaEventRecord(b, st);
                cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b);
                printf("smem %6d host %d work %6d: %.2f us per launch\n", smem, host, work, 1e3f * ms / N);
            }
    return 0;
}

// Microbenchmark: how are single-warp CTAs distributed over the 4 sub-partitions (SMSPs) of an SM?
// Each CTA = 1 warp running 4 independent FADD chains (one warp alone nearly saturates one scheduler's issue port).
// K CTAs per SM are forced to be co-resident through their shared-memory size. If the K warps land on different SMSPs
// the run time stays flat up to K = 4; if they pile up on one SMSP it grows linearly with K.
// Also prints the %warpid (hardware warp slot) histogram modulo 4.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chains(float* out, int iters, int* warpids, int threads) {
    extern __shared__ float sm[];
    float a = threadIdx.x, b = a + 1, c = a + 2, d = a + 3;
    for (int i = 0; i < iters; ++i) { a += 1.0f; b += 1.0f; c += 1.0f; d += 1.0f; a *= 0.999f; b *= 0.999f; c *= 0.999f; d *= 0.999f; }
    unsigned wid, smid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((threadIdx.x & 31) == 0) warpids[blockIdx.x * (threads / 32) + threadIdx.x / 32] = (int)(wid | (smid << 16));
    if (a + b + c + d == 12345.0f) out[0] = a;
    if (iters < 0) sm[threadIdx.x] = a;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4);
    int* wids; cudaMalloc(&wids, sizeof(int) * sms * 64);
    cudaFuncSetAttribute(chains, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(chains, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int iters = 200000;
    for (int threads : {32, 128}) {
        for (int K : {1, 2, 3, 4, 6, 8, 12, 16}) {
            if (threads == 128 && K > 4) continue;
            int smem = (220 * 1024) / K - 1024;           // K CTAs fill the SM's shared memory
            smem &= ~127;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            chains<<<sms * K, threads, smem>>>(out, 1000, wids, threads);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            chains<<<sms * K, threads, smem>>>(out, iters, wids, threads);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            static int h[148 * 64 + 1024];
            const int nw = sms * K * (threads / 32);
            cudaMemcpy(h, wids, sizeof(int) * nw, cudaMemcpyDeviceToHost);
            int hist[4] = {0, 0, 0, 0};
            int sm0[64], n0 = 0;
            for (int i = 0; i < nw; ++i) { hist[(h[i] & 0xffff) & 3]++; if ((h[i] >> 16) == 0 && n0 < 64) sm0[n0++] = h[i] & 0xffff; }
            printf("threads/CTA %3d  CTAs/SM %2d  warps/SM %2d: %7.2f ms  warpid%%4 histogram %d %d %d %d   warp slots on SM0:", threads, K,
                   K * threads / 32, ms, hist[0], hist[1], hist[2], hist[3]);
            for (int i = 0; i < n0; ++i) printf(" %d", sm0[i]);
            printf("  (%s)\n", cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}

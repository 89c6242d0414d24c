// Micro-benchmark: throughput of red.global.add.f64 into 72-byte slots with ~8 contributions per slot spread over a
// window of the input (the access pattern of accumulating 3x3 Hessian blocks into a table instead of sort + reduce).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_f64 red_f64.cu && ./red_f64
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(288) k_red(double* __restrict__ table, const double* __restrict__ val, long nBlocks, long nSlots, long window)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long b = t / 9;
    const int comp = (int)(t - b * 9);
    if (b >= nBlocks) return;
    // 8 contributions per slot, spaced `window` blocks apart
    const long grp = b / (8 * window), r = b % (8 * window);
    const long slot = (grp * window + r % window) % nSlots;
    atomicAdd(table + 9 * slot + comp, val[9 * b + comp]);
}
__global__ void __launch_bounds__(288) k_red_hash(double* __restrict__ table, unsigned long long* __restrict__ keys, const double* __restrict__ val,
    long nBlocks, long cap, long window)
{
    __shared__ long sslot[32];
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long b = t / 9;
    const int comp = (int)(t - b * 9);
    if (b < nBlocks && comp == 0) {
        const long grp = b / (8 * window), r = b % (8 * window);
        const unsigned long long key = (unsigned long long)(grp * window + r % window) * 2654435761ull + 12345ull;
        unsigned long long h = (key * 0x9E3779B97F4A7C15ull) >> 20;
        long s = (long)(h % (unsigned long long)cap);
        while (true) {
            unsigned long long k = keys[s];
            if (k == key) break;
            if (k == ~0ull) { k = atomicCAS(keys + s, ~0ull, key); if (k == ~0ull || k == key) break; }
            s = (s + 1 == cap) ? 0 : s + 1;
        }
        sslot[threadIdx.x / 9] = s;
    }
    __syncthreads();
    if (b >= nBlocks) return;
    atomicAdd(table + 9 * sslot[threadIdx.x / 9] + comp, val[9 * b + comp]);
}
int main()
{
    const long nBlocks = 200000000L, nSlots = nBlocks / 8, cap = 1L << 26;
    double *table, *val; unsigned long long* keys;
    cudaMalloc(&table, (size_t)cap * 72); cudaMalloc(&val, (size_t)nBlocks * 72); cudaMalloc(&keys, (size_t)cap * 8);
    cudaMemset(val, 0, (size_t)nBlocks * 72);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (long window : {256L, 4096L, 65536L}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(table, 0, (size_t)cap * 72);
            cudaEventRecord(e0);
            k_red<<<(unsigned)((nBlocks * 9 + 287) / 288), 288>>>(table, val, nBlocks, nSlots, window);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("direct  window %6ld: %.2f ms (%.1f G red/s) %s\n", window, ms, nBlocks * 9 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(table, 0, (size_t)cap * 72); cudaMemset(keys, 0xff, (size_t)cap * 8);
            cudaEventRecord(e0);
            k_red_hash<<<(unsigned)((nBlocks * 9 + 287) / 288), 288>>>(table, keys, val, nBlocks, cap, window);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("hashed  window %6ld: %.2f ms (%.1f G red/s) %s\n", window, ms, nBlocks * 9 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    cudaEventRecord(e0); cudaMemsetAsync(table, 0, (size_t)cap * 72); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("memset %.1f GB: %.2f ms\n", cap * 72 / 1e9, ms);
    return 0;
}

// microbench.cu -- raw B200 rates that bound the counting kernels (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench.bin tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// random global RED into a region of `mask+1` words
__global__ void redg_kernel(uint32_t* hist, uint32_t mask, int per_thread) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = hash32(t * 2654435761u + 12345u);
    for (int i = 0; i < per_thread; i++) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = hash32(x) & mask;
        asm volatile("red.global.add.u32 [%0], 1;" ::"l"(hist + idx) : "memory");
    }
}

// k-mer like access: consecutive "windows" share all but 2 bits (kmer = kmer*4+b)
__global__ void redg_kmer_kernel(uint32_t* hist, uint32_t mask, int per_thread) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = hash32(t * 2654435761u + 12345u);
    uint32_t kmer = hash32(x);
    for (int i = 0; i < per_thread; i++) {
        if ((i & 15) == 0) x = hash32(x + i);
        kmer = (kmer << 2) | ((x >> (2 * (i & 15))) & 3u);
        asm volatile("red.global.add.u32 [%0], 1;" ::"l"(hist + (kmer & mask)) : "memory");
    }
}

template <int RET>
__global__ void reds_kernel(uint32_t* out, uint32_t mask, int per_thread) {
    extern __shared__ uint32_t sh[];
    const uint32_t nwords = RET == 2 ? (mask + 1) / 2 : mask + 1;
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = hash32(t * 2654435761u + 12345u);
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh);
    uint32_t acc = 0;
    for (int i = 0; i < per_thread; i++) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = (x >> 8) & mask;
        if (RET == 0) {
            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sbase + idx * 4) : "memory");
        } else if (RET == 1) {
            uint32_t old;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(sbase + idx * 4) : "memory");
            acc += old;
        } else {   // packed u16: increment half-word via 32-bit add, watch the old value
            uint32_t old;
            uint32_t inc = 1u << (16 * (idx & 1));
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(sbase + (idx >> 1) * 4), "r"(inc) : "memory");
            if (((old >> (16 * (idx & 1))) & 0xFFFFu) == 0x3FFFu) acc++;
        }
    }
    __syncthreads();
    uint32_t s = acc;
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) s += sh[i];
    if (s == 0xdeadbeef) out[0] = s;
}

// conflict-free variant: each lane owns a bank
__global__ void reds_nobank_kernel(uint32_t* out, uint32_t mask, int per_thread) {
    extern __shared__ uint32_t sh[];
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t x = hash32(t * 2654435761u + 12345u);
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh);
    uint32_t lane = threadIdx.x & 31;
    for (int i = 0; i < per_thread; i++) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = (((x >> 8) & mask) & ~31u) | lane;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sbase + idx * 4) : "memory");
    }
    __syncthreads();
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i <= mask; i += blockDim.x) s += sh[i];
    if (s == 0xdeadbeef) out[0] = s;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s SMs %d L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    uint32_t* hist;
    size_t max_bytes = 512ull << 20;
    CK(cudaMalloc(&hist, max_bytes));
    CK(cudaMemset(hist, 0, max_bytes));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, per = 256;
    const double n_ops = (double)blocks * threads * per;
    for (int kind = 0; kind < 2; kind++) {
        for (size_t mb = 64; mb <= 64; mb *= 2) {
            uint32_t mask = (uint32_t)((mb << 20) / 4 - 1);
            float best = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                CK(cudaMemsetAsync(hist, 0, mb << 20));
                CK(cudaEventRecord(a));
                if (kind == 0) redg_kernel<<<blocks, threads>>>(hist, mask, per);
                else redg_kmer_kernel<<<blocks, threads>>>(hist, mask, per);
                CK(cudaEventRecord(b));
                CK(cudaEventSynchronize(b));
                float ms; CK(cudaEventElapsedTime(&ms, a, b));
                if (ms < best) best = ms;
            }
            printf("REDG %s region %4zu MB : %7.1f G atomics/s  (%.3f ms)\n", kind ? "kmer  " : "random", mb, n_ops / best / 1e6, best);
        }
    }
    // shared atomics
    CK(cudaFuncSetAttribute(reds_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
    CK(cudaFuncSetAttribute(reds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
    CK(cudaFuncSetAttribute(reds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
    CK(cudaFuncSetAttribute(reds_nobank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
    uint32_t* out = hist;
    for (int kb : {16, 64, 128}) {
        for (int th : {256, 512, 1024}) {
            uint32_t words = kb * 1024 / 4;
            int bl = prop.multiProcessorCount * (kb <= 64 ? (1024 / th) * (kb <= 16 ? 2 : 1) : 1);
            const int per_s = 2048;
            double ops = (double)bl * th * per_s;
            for (int variant = 0; variant < 4; variant++) {
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(a));
                    if (variant == 0) reds_kernel<0><<<bl, th, kb * 1024>>>(out, words - 1, per_s);
                    if (variant == 1) reds_kernel<1><<<bl, th, kb * 1024>>>(out, words - 1, per_s);
                    if (variant == 2) reds_kernel<2><<<bl, th, kb * 1024>>>(out, 2 * words - 1, per_s);
                    if (variant == 3) reds_nobank_kernel<<<bl, th, kb * 1024>>>(out, words - 1, per_s);
                    CK(cudaEventRecord(b));
                    CK(cudaEventSynchronize(b));
                    CK(cudaGetLastError());
                    float ms; CK(cudaEventElapsedTime(&ms, a, b));
                    if (ms < best) best = ms;
                }
                const char* nm[] = {"red.shared", "atom.shared(ret)", "packed-u16(ret)", "red.shared no-bank-conflict"};
                printf("SMEM %3d KB hist, %4d thr x %d CTA : %-28s %7.1f G atomics/s\n", kb, th, bl, nm[variant], ops / best / 1e6);
            }
        }
    }
    return 0;
}

// gram_tc.cu -- exact integer Gram matrix of count rows on the 5th-generation tensor cores.
//
// The genome x genome distance matrix is the only dense contraction of the
// path.  The 1e-6 relative tolerance on 1 - cos(x, y) rules out bf16/tf32 products and an fp32
// accumulator over 65 536 terms, but the feature rows ARE integer counts, so the Gram matrix
//      G = C C^T,   C = sum_d 256^d D_d   (D_d = d-th byte of every count, uint8)
// is computed EXACTLY as   G = sum_{a,b} 256^(a+b) D_a D_b^T   with
//      tcgen05.mma.cta_group::1.kind::i8   (u8 x u8 -> s32 accumulators in TMEM),
// at most 8192 features per accumulation (255*255*8192 < 2^31), then summed in int64.
//
// One CTA (128 threads) owns a 128 x 128 tile of one (a, b) digit pair and one K split:
//   global uint8 rows --16 B loads--> registers --> shared memory in the UMMA canonical K-major,
//   no-swizzle core-matrix layout (8 rows x 16 B per core matrix) --> two MMAs (K = 32 each) per
//   64-byte K block, issued by one thread, tracked by an mbarrier (tcgen05.commit); the next
//   block's global loads are in flight while the tensor core works.  Epilogue: tcgen05.ld
//   (32 lanes x 32 columns per warp) and int64 atomicAdd of scale * acc into G (and its transpose
//   for a != b).
#include "internal.h"

namespace km {

constexpr int GT_M = 128, GT_N = 128, GT_KB = 64;       // tile and K block (bytes = uint8 elements)
constexpr int GT_KSPLIT = 8192;                         // features per accumulation: 255^2 * 8192 < 2^31
constexpr int GT_THREADS = 128;

__global__ void split_digits_kernel(const uint32_t* __restrict__ counts, uint64_t stride, int n, uint64_t m,
                                    uint8_t* __restrict__ planes /* [4][n][m] */, unsigned int* max_count) {
    const uint64_t total = (uint64_t)n * m;
    unsigned int mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / m, c = i % m;
        const uint32_t v = counts[r * stride + c];
        mx = max(mx, v);
        planes[i] = (uint8_t)v;
        planes[total + i] = (uint8_t)(v >> 8);
        planes[2 * total + i] = (uint8_t)(v >> 16);
        planes[3 * total + i] = (uint8_t)(v >> 24);
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(max_count, mx);
}

__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14) | LBO >> 4 [16,30) | SBO >> 4 [32,46) | version 1 [46,48) | layout 0
    const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14);
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(GT_THREADS)
gram_i8_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int n, uint64_t m, int symmetric_pair,
               unsigned long long scale, unsigned long long* __restrict__ G, int row_tile0, int rows_only) {
    // rows_only (multi-GPU row block): tile rows row_tile0 .. row_tile0 + gridDim.y - 1 against ALL columns, every
    // (a, b) digit pair launched separately, nothing mirrored: G holds rows [128 row_tile0, ...) only, at their
    // global row index.
    // shared: A tile and B tile in core-matrix layout: [kg (4)][mg (16)][8 rows][16 B]
    __shared__ __align__(128) uint8_t sA[GT_M * GT_KB];
    __shared__ __align__(128) uint8_t sB[GT_N * GT_KB];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = (blockIdx.y + row_tile0) * GT_M, col0 = blockIdx.x * GT_N;
    if (!rows_only && symmetric_pair && col0 < row0) return;   // D_a D_a^T: the upper triangle is enough
    const uint64_t k_begin = (uint64_t)blockIdx.z * GT_KSPLIT;
    const uint64_t k_end = min(m, k_begin + GT_KSPLIT);

    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {                                        // 128 TMEM columns: the 128 x 128 s32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = tmem_base_s;

    const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA);
    const uint32_t sB_addr = (uint32_t)__cvta_generic_to_shared(sB);
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 [4,6), a/b format u8 = 0,
    // a/b K-major = 0, n_dim = N >> 3 [17,23), m_dim = M >> 4 [24,29)
    const uint32_t idesc = (2u << 4) | ((uint32_t)(GT_N >> 3) << 17) | ((uint32_t)(GT_M >> 4) << 24);

    // each thread moves 4 + 4 16-byte pieces per K block: piece q -> row q / 4, 16-byte column q % 4
    uint4 ra[4], rb[4];
    auto load_block = [&](uint64_t k0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = tid + GT_THREADS * i, r = q >> 2, p = q & 3;
            const uint64_t kk = k0 + 16 * p;
            const bool kin = kk < k_end;
            ra[i] = (row0 + r < n && kin) ? *reinterpret_cast<const uint4*>(A + (uint64_t)(row0 + r) * m + kk) : make_uint4(0, 0, 0, 0);
            rb[i] = (col0 + r < n && kin) ? *reinterpret_cast<const uint4*>(B + (uint64_t)(col0 + r) * m + kk) : make_uint4(0, 0, 0, 0);
        }
    };
    auto store_block = [&]() {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q = tid + GT_THREADS * i, r = q >> 2, p = q & 3;
            const uint32_t off = (uint32_t)((p * 16 + (r >> 3)) * 128 + (r & 7) * 16);
            *reinterpret_cast<uint4*>(sA + off) = ra[i];
            *reinterpret_cast<uint4*>(sB + off) = rb[i];
        }
    };

    uint32_t phase = 0;
    bool first = true;
    load_block(k_begin);
    for (uint64_t k0 = k_begin; k0 < k_end; k0 += GT_KB) {
        if (!first) { mbar_wait(bar_addr, phase); phase ^= 1; }   // previous MMAs have consumed the tiles
        store_block();
        if (k0 + GT_KB < k_end) load_block(k0 + GT_KB);           // next block in flight during the MMAs
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < GT_KB / 32; j++) {
                // MMA j reads K bytes [32 j, 32 j + 32): core-matrix columns 2j and 2j+1
                const uint64_t da = umma_desc_kmajor_noswizzle(sA_addr + (uint32_t)(2 * j) * 16 * 128, 16 * 128, 128);
                const uint64_t db = umma_desc_kmajor_noswizzle(sB_addr + (uint32_t)(2 * j) * 16 * 128, 16 * 128, 128);
                const uint32_t accumulate = (first && j == 0) ? 0u : 1u;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
                    "}\n" ::"r"(tmem_acc), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
        }
        first = false;
    }
    mbar_wait(bar_addr, phase);                              // the last MMAs are done
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // epilogue: warp w owns TMEM lanes (= tile rows) 32 w .. 32 w + 31
    const int gi = row0 + warp * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < GT_N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (gi < n) {
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int gj = col0 + c0 + c;
                if (gj >= n || !v[c]) continue;
                const unsigned long long add = (unsigned long long)v[c] * scale;
                if (rows_only) {
                    atomicAdd(G + (uint64_t)gi * n + gj, add);
                } else if (symmetric_pair) {
                    if (gj < gi) continue;                                    // upper triangle (tiles on the diagonal)
                    atomicAdd(G + (uint64_t)gi * n + gj, add);
                    if (gj != gi) atomicAdd(G + (uint64_t)gj * n + gi, add);
                } else {                                                       // D_a D_b^T + (D_a D_b^T)^T
                    atomicAdd(G + (uint64_t)gi * n + gj, add);
                    atomicAdd(G + (uint64_t)gj * n + gi, add);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_acc));
}

__global__ void gram_to_double_kernel(const unsigned long long* __restrict__ G, double* __restrict__ out, uint64_t nn) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) out[i] = (double)G[i];
}

size_t gram_tc_workspace(int n, uint64_t m) { return (size_t)4 * n * m + (size_t)n * n * 8 + 512; }

// Exact Gram matrix (as doubles, exact below 2^53) of uint32 count rows.  m must be a multiple of 64.
int launch_gram_tc(const uint32_t* d_counts, uint64_t stride, int n, uint64_t m, void* workspace, double* d_gram,
                   cudaStream_t s) {
    uint8_t* planes = (uint8_t*)workspace;
    const size_t plane = (size_t)n * m;
    unsigned long long* G = (unsigned long long*)(planes + ((4 * plane + 255) / 256) * 256);
    unsigned int* d_max = (unsigned int*)(G + (size_t)n * n);
    KM_CUDA(cudaMemsetAsync(G, 0, (size_t)n * n * 8 + 8, s));
    split_digits_kernel<<<148 * 8, 256, 0, s>>>(d_counts, stride, n, m, planes, d_max);
    KM_CUDA(cudaGetLastError());
    unsigned int h_max = 0;
    KM_CUDA(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, s));
    KM_CUDA(cudaStreamSynchronize(s));
    int nd = h_max >= (1u << 24) ? 4 : h_max >= (1u << 16) ? 3 : h_max >= (1u << 8) ? 2 : 1;
    const unsigned tiles = (unsigned)((n + GT_M - 1) / GT_M);
    const unsigned ksplits = (unsigned)((m + GT_KSPLIT - 1) / GT_KSPLIT);
    for (int a = 0; a < nd; a++) {
        for (int b = a; b < nd; b++) {
            if (8 * (a + b) >= 64) continue;                 // would not fit 64 bits anyway (counts^2 sums < 2^63 assumed)
            const unsigned long long scale = 1ull << (8 * (a + b));
            gram_i8_kernel<<<dim3(tiles, tiles, ksplits), GT_THREADS, 0, s>>>(planes + a * plane, planes + b * plane, n, m,
                                                                             a == b ? 1 : 0, scale, G, 0, 0);
            KM_CUDA(cudaGetLastError());
        }
    }
    const uint64_t nn = (uint64_t)n * n;
    gram_to_double_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(G, d_gram, nn);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

// Exact squared norms of the count rows (uint64, then double): the Gram diagonal every rank of a row-block
// distance needs for ALL genomes.  One warp per row.
__global__ void row_sumsq_kernel(const uint32_t* __restrict__ counts, uint64_t stride, int n, uint64_t m, double* __restrict__ out) {
    const int row = (int)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const uint32_t* c = counts + (uint64_t)row * stride;
    unsigned long long acc = 0;
    for (uint64_t i = lane; i < m; i += 32) {
        const unsigned long long v = c[i];
        acc += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = (double)acc;
}

// rows [row_begin, row_end) of the distance matrix from the row block of G and the squared norms of all rows
__global__ void distance_rows_kernel(const unsigned long long* __restrict__ G, const double* __restrict__ norm2, int n,
                                     int row_begin, int row_end, int metric, float* D32, double* D64) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)(row_end - row_begin) * n) return;
    const int i = row_begin + (int)(idx / n), j = (int)(idx % n);
    const double gii = norm2[i], gjj = norm2[j], gij = (double)G[(uint64_t)i * n + j];
    double d;
    if (i == j) {
        d = 0.0;
    } else if (metric == 0) {
        const double den = sqrt(gii) * sqrt(gjj);
        d = den > 0.0 ? 1.0 - gij / den : nan("");
    } else {
        const double d2 = gii + gjj - 2.0 * gij;
        d = d2 > 0.0 ? sqrt(d2) : 0.0;
    }
    if (D32) D32[idx] = (float)d;
    if (D64) D64[idx] = d;
}

size_t gram_rows_workspace(int n, uint64_t m) { return gram_tc_workspace(n, m) + (size_t)n * 8 + 256; }

// Rows [row_begin, row_end) of the n x n distance matrix of uint32 count rows (all n rows resident): the unit one
// rank computes when the genomes were counted on several GPUs and the rows gathered.  The Gram
// entries are exact (tcgen05 kind::i8), so the block is bit-identical to the same rows of the single-GPU matrix.
int launch_distance_rows_tc(const uint32_t* d_counts, uint64_t stride, int n, uint64_t m, int row_begin, int row_end,
                            int metric, void* workspace, float* d_out32, double* d_out64, cudaStream_t s) {
    if (row_begin >= row_end) return KMERML_OK;
    uint8_t* planes = (uint8_t*)workspace;
    const size_t plane = (size_t)n * m;
    unsigned long long* G = (unsigned long long*)(planes + ((4 * plane + 255) / 256) * 256);
    unsigned int* d_max = (unsigned int*)(G + (size_t)n * n);
    double* d_norm = (double*)((uint8_t*)workspace + gram_tc_workspace(n, m) / 256 * 256 + 256);
    const int t0 = row_begin / GT_M, t1 = (row_end + GT_M - 1) / GT_M;
    KM_CUDA(cudaMemsetAsync(G + (size_t)t0 * GT_M * n, 0, (size_t)(std::min(t1 * GT_M, n) - t0 * GT_M) * n * 8, s));
    KM_CUDA(cudaMemsetAsync(d_max, 0, 8, s));
    split_digits_kernel<<<148 * 8, 256, 0, s>>>(d_counts, stride, n, m, planes, d_max);
    row_sumsq_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, s>>>(d_counts, stride, n, m, d_norm);
    KM_CUDA(cudaGetLastError());
    unsigned int h_max = 0;
    KM_CUDA(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, s));
    KM_CUDA(cudaStreamSynchronize(s));
    const int nd = h_max >= (1u << 24) ? 4 : h_max >= (1u << 16) ? 3 : h_max >= (1u << 8) ? 2 : 1;
    const unsigned tiles = (unsigned)((n + GT_N - 1) / GT_N);
    const unsigned ksplits = (unsigned)((m + GT_KSPLIT - 1) / GT_KSPLIT);
    for (int a = 0; a < nd; a++) {
        for (int b = 0; b < nd; b++) {
            if (8 * (a + b) >= 64) continue;
            gram_i8_kernel<<<dim3(tiles, (unsigned)(t1 - t0), ksplits), GT_THREADS, 0, s>>>(
                planes + a * plane, planes + b * plane, n, m, 0, 1ull << (8 * (a + b)), G, t0, 1);
            KM_CUDA(cudaGetLastError());
        }
    }
    const uint64_t cells = (uint64_t)(row_end - row_begin) * n;
    distance_rows_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(G, d_norm, n, row_begin, row_end, metric, d_out32, d_out64);
    KM_CUDA(cudaGetLastError());
    return KMERML_OK;
}

}  // namespace km

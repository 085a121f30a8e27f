// Micro-benchmark behind the write-side design of loss_backward / detect_forward: how fast can 135 MB of zeros be
// written on a B200 by (a) cudaMemsetAsync, (b) 16-byte register stores from a persistent grid, (c) bulk asynchronous
// stores (cp.async.bulk.global.shared::cta) from one constant block of zeros in shared memory.  CUDA events, a 512 MiB
// memset (L2 flush) before every timed call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zero_bw tools/zero_bw.cu && ./zero_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

__global__ void stg_kernel(float4* __restrict__ dst, size_t n16) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// every CTA owns a contiguous span; each warp writes 512-byte lines, 4 stores in flight per thread
__global__ void stg_span_kernel(float4* __restrict__ dst, size_t n16) {
    const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
    const size_t lo = per * blockIdx.x, hi = min(n16, lo + per);
    for (size_t i = lo + threadIdx.x; i < hi; i += 4 * blockDim.x) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i + u * blockDim.x < hi) dst[i + u * blockDim.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int kChunk>
__global__ void bulk_kernel(unsigned char* __restrict__ dst, size_t bytes, int strided) {
    __shared__ __align__(128) unsigned char z[kChunk];
    for (int i = threadIdx.x; i < kChunk / 16; i += blockDim.x) reinterpret_cast<float4*>(z)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    if (strided) {
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const size_t off = c * kChunk;
            const unsigned n = (unsigned)min((size_t)kChunk, bytes - off);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off),
                         "r"((unsigned)__cvta_generic_to_shared(z)), "r"(n) : "memory");
        }
    } else {
        const size_t per = (nchunks + gridDim.x - 1) / gridDim.x;
        const size_t lo = per * blockIdx.x, hi = min(nchunks, lo + per);
        for (size_t c = lo; c < hi; ++c) {
            const size_t off = c * kChunk;
            const unsigned n = (unsigned)min((size_t)kChunk, bytes - off);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off),
                         "r"((unsigned)__cvta_generic_to_shared(z)), "r"(n) : "memory");
        }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float timed(F fn, void* flush, size_t flush_bytes, int n = 15) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    std::vector<float> ms;
    for (int i = 0; i < n + 3; ++i) {
        cudaMemsetAsync(flush, 1, flush_bytes);
        cudaEventRecord(a);
        fn();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float t; cudaEventElapsedTime(&t, a, b);
        if (i >= 3) ms.push_back(t);
    }
    std::sort(ms.begin(), ms.end());
    return ms[ms.size() / 2];
}

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? (size_t)atoll(argv[1]) : (size_t)32 * 16320 * 81 * 4);
    unsigned char *dst, *flush;
    const size_t flush_bytes = (size_t)512 << 20;
    cudaMalloc(&dst, bytes); cudaMalloc(&flush, flush_bytes);
    auto rep = [&](const char* name, float ms) { printf("%-40s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / (ms * 1e-3) / 1e9); };
    rep("cudaMemsetAsync", timed([&] { cudaMemsetAsync(dst, 0, bytes); }, flush, flush_bytes));
    for (int per_sm : {2, 4, 8}) {
        char nm[64];
        snprintf(nm, 64, "stg128 grid-stride %d x148 x256", per_sm);
        rep(nm, timed([&] { stg_kernel<<<148 * per_sm, 256>>>((float4*)dst, bytes / 16); }, flush, flush_bytes));
        snprintf(nm, 64, "stg128 span        %d x148 x256", per_sm);
        rep(nm, timed([&] { stg_span_kernel<<<148 * per_sm, 256>>>((float4*)dst, bytes / 16); }, flush, flush_bytes));
    }
    rep("stg128 one thread per 16 B", timed([&] { stg_kernel<<<(unsigned)((bytes / 16 + 255) / 256), 256>>>((float4*)dst, bytes / 16); }, flush, flush_bytes));
    for (int strided : {1, 0})
        for (int per_sm : {1, 2, 4}) {
            char nm[64];
            snprintf(nm, 64, "bulk  4 KB %s %d x148", strided ? "strided" : "span   ", per_sm);
            rep(nm, timed([&] { bulk_kernel<4096><<<148 * per_sm, 32>>>(dst, bytes, strided); }, flush, flush_bytes));
            snprintf(nm, 64, "bulk 16 KB %s %d x148", strided ? "strided" : "span   ", per_sm);
            rep(nm, timed([&] { bulk_kernel<16384><<<148 * per_sm, 32>>>(dst, bytes, strided); }, flush, flush_bytes));
            snprintf(nm, 64, "bulk 32 KB %s %d x148", strided ? "strided" : "span   ", per_sm);
            rep(nm, timed([&] { bulk_kernel<32768><<<148 * per_sm, 32>>>(dst, bytes, strided); }, flush, flush_bytes));
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

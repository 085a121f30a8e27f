// Micro-benchmark behind the read-side roofline of conf_loss: how fast can 174 MB be READ once on a B200 by (a) 16-byte
// loads from a persistent grid (sum into a register, one store per thread), (b) a ring of bulk asynchronous loads
// (cp.async.bulk.shared::cluster.global + mbarrier) with a trivial consumer.  CUDA events, a 512 MiB memset (L2 flush)
// before every timed call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw tools/read_bw.cu && ./read_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

__global__ void ldg_kernel(const float4* __restrict__ src, size_t n16, float* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        float4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(src + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(src + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(src + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(src + i + 3 * stride));
        acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
    }
    for (; i < n16; i += stride) { const float4 a = src[i]; acc += a.x + a.y + a.z + a.w; }
    if (acc == 123.456f) out[0] = acc;
}

template <int kTile, int kStages>
__global__ void bulk_kernel(const unsigned char* __restrict__ src, size_t bytes, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) unsigned long long bar[kStages];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t ntiles = bytes / kTile;
    auto issue = [&](size_t tile, int s) {
        const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(kTile) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(ring + (size_t)s * kTile)),
                     "l"(src + tile * kTile), "r"(kTile), "r"(b) : "memory");
    };
    if (tid == 0)
        for (int k = 0; k < kStages - 1; ++k) {
            const size_t t = blockIdx.x + (size_t)k * gridDim.x;
            if (t < ntiles) issue(t, k);
        }
    float acc = 0.f;
    int it = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % kStages;
        if (tid == 0) {
            const size_t nxt = tile + (size_t)(kStages - 1) * gridDim.x;
            if (nxt < ntiles) issue(nxt, (it + kStages - 1) % kStages);
        }
        const unsigned b = (unsigned)__cvta_generic_to_shared(&bar[s]);
        const unsigned parity = (it / kStages) & 1;
        unsigned ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        const float4* t4 = reinterpret_cast<const float4*>(ring + (size_t)s * kTile);
        for (int i = tid; i < kTile / 16; i += blockDim.x) { const float4 a = t4[i]; acc += a.x + a.y + a.z + a.w; }
        __syncthreads();
    }
    if (acc == 123.456f) out[0] = acc;
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

template <int kTile, int kStages>
static void run_bulk(const unsigned char* src, size_t bytes, float* out, void* flush, size_t fb, int per_sm, int threads) {
    const size_t smem = (size_t)kTile * kStages;
    cudaFuncSetAttribute(bulk_kernel<kTile, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const float ms = timed([&] { bulk_kernel<kTile, kStages><<<148 * per_sm, threads, smem>>>(src, bytes, out); }, flush, fb);
    printf("bulk ring %3d KB x %d stages, %d x148 x%-4d   %8.2f us  %7.1f GB/s\n", kTile / 1024, kStages, per_sm, threads, ms * 1e3,
           bytes / (ms * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
    size_t bytes = (argc > 1 ? (size_t)atoll(argv[1]) : (size_t)32 * 16320 * 81 * 4);
    bytes = bytes / 65536 * 65536;
    unsigned char *src, *flush;
    float* out;
    const size_t fb = (size_t)512 << 20;
    cudaMalloc(&src, bytes); cudaMalloc(&flush, fb); cudaMalloc(&out, 4);
    cudaMemset(src, 0, bytes);
    for (int per_sm : {2, 4, 8, 16}) {
        const float ms = timed([&] { ldg_kernel<<<148 * per_sm, 256>>>((const float4*)src, bytes / 16, out); }, flush, fb);
        printf("ldg128 x4 grid-stride %2d x148 x256              %8.2f us  %7.1f GB/s\n", per_sm, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    }
    {
        const float ms = timed([&] { ldg_kernel<<<(unsigned)(bytes / 16 / 256 / 4), 256>>>((const float4*)src, bytes / 16, out); }, flush, fb);
        printf("ldg128 x4 one pass (grid = n/1024)               %8.2f us  %7.1f GB/s\n", ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    }
    run_bulk<16384, 4>(src, bytes, out, flush, fb, 2, 256);
    run_bulk<32768, 3>(src, bytes, out, flush, fb, 2, 256);
    run_bulk<32768, 2>(src, bytes, out, flush, fb, 3, 256);
    run_bulk<16384, 4>(src, bytes, out, flush, fb, 3, 256);
    run_bulk<16384, 3>(src, bytes, out, flush, fb, 4, 128);
    run_bulk<65536, 3>(src, bytes, out, flush, fb, 1, 512);
    run_bulk<8192, 4>(src, bytes, out, flush, fb, 6, 128);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

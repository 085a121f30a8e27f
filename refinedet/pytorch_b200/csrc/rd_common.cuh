// rd_common.cuh — shared device math for the RefineDet hot-path kernels (sm_100a).
//
// All arithmetic is fp32 in the reference's operation order; the library is built
// with -fmad=false so no mul/add pair is contracted.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include "../../../include/refinedet_b200.h"

namespace rd {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- launch accounting (rd_launch_count) ------------------------------------
void note_launch(int n = 1);

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, so
// the high-water mark is kept per device: a process that drives several GPUs sets it on each of them.
constexpr int kMaxDevices = 64;
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&high_water)[kMaxDevices]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t& hw = high_water[(unsigned)dev % kMaxDevices];
    // The 48 KB that need no opt-in cover STATIC + dynamic shared memory; the kernels here hold at most a few
    // KB of static arrays, so anything above 40 KB of dynamic memory opts in (once per kernel and device).
    if (hw == 0) hw = 40 * 1024;
    if (bytes <= hw) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) hw = bytes;
    return e;
}

// ---- NVTX range around the launches of one C-ABI entry (header-only NVTX 3: a no-op unless a tool is attached;
// Nsight Systems / Compute then show "rd_detect_fused", "rd_multibox_criterion", ... around their kernels)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- launchers shared between translation units (the fused criterion in rd_loss.cu chains kernels of rd_match.cu)
// `pdl`: launch with programmatic stream serialisation (every kernel of the chain waits with grid_dependency_wait)
int match_launch(const float4* truths, const float* labels, const int* gt_count, const float4* priors,
                 const float4* arm_loc, int B, int P, int Gmax, float threshold, float v0, float v1, int label_mode,
                 unsigned long long* best_prior, float* bt_overlap, int* bt_idx, float4* loc_t, long long* conf_t,
                 cudaStream_t st, bool pdl);
int hnm_launch(const float* loss_c, const unsigned char* pos, int B, int P, int negpos_ratio, unsigned char* neg_out,
               int* num_pos_out, cudaStream_t st, bool pdl);

#define RD_CHECK_LAUNCH()                                   \
    do {                                                    \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

// ---- programmatic dependent launch (sm_90+) ----------------------------------
// wait until the preceding kernel of the stream has completed and its memory is visible / allow the next kernel
// of the stream to start launching.  A kernel launched WITHOUT the attribute passes the wait at once.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with programmatic stream serialisation: the kernel may begin launching while its predecessor in
// the stream drains; kernels launched this way call grid_dependency_wait() before touching its results
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- streaming loads ---------------------------------------------------------
// read-once data: bypass L1 allocation (guide: Guideline 13/14)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
                 : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// ---- layers/box_utils.py ----------------------------------------------------
// decode, box_utils.py:200-204: cxcy = p.xy + (loc.xy*v0)*p.wh ; wh = p.wh*exp(loc.wh*v1)
//                               x1y1 = cxcy - wh/2 ; x2y2 = wh + x1y1
__device__ __forceinline__ float4 decode_box(float4 loc, float4 p, float v0, float v1) {
    float cx = p.x + (loc.x * v0) * p.z;
    float cy = p.y + (loc.y * v0) * p.w;
    float w = p.z * expf(loc.z * v1);
    float h = p.w * expf(loc.w * v1);
    float4 o;
    o.x = cx - w / 2.0f;
    o.y = cy - h / 2.0f;
    o.z = w + o.x;
    o.w = h + o.y;
    return o;
}
// center_size, box_utils.py:25-26
__device__ __forceinline__ float4 center_size_box(float4 b) {
    float4 o;
    o.x = (b.z + b.x) / 2.0f;
    o.y = (b.w + b.y) / 2.0f;
    o.z = b.z - b.x;
    o.w = b.w - b.y;
    return o;
}
// point_form, box_utils.py:13-14
__device__ __forceinline__ float4 point_form_box(float4 b) {
    float4 o;
    float hw = b.z / 2.0f, hh = b.w / 2.0f;
    o.x = b.x - hw;
    o.y = b.y - hh;
    o.z = b.x + hw;
    o.w = b.y + hh;
    return o;
}
// encode, box_utils.py:175-183
__device__ __forceinline__ float4 encode_box(float4 m, float4 p, float v0, float v1) {
    float4 o;
    o.x = ((m.x + m.z) / 2.0f - p.x) / (v0 * p.z);
    o.y = ((m.y + m.w) / 2.0f - p.y) / (v0 * p.w);
    o.z = logf((m.z - m.x) / p.z + 1e-5f) / v1;
    o.w = logf((m.w - m.y) / p.w + 1e-5f) / v1;
    return o;
}
// Detect_RefineDet two-stage decode, detection_refinedet.py:57-59
__device__ __forceinline__ float4 refine_decode(float4 arm, float4 odm, float4 prior,
                                                float v0, float v1) {
    float4 r = center_size_box(decode_box(arm, prior, v0, v1));
    return decode_box(odm, r, v0, v1);
}
// intersect / jaccard for one pair, box_utils.py:42-47, 62-68 (a = truth, b = prior box)
__device__ __forceinline__ float intersect_pair(float4 a, float4 b) {
    float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.0f);
    float h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.0f);
    return w * h;
}
__device__ __forceinline__ float jaccard_pair(float4 a, float4 b) {
    float inter = intersect_pair(a, b);
    float area_a = (a.z - a.x) * (a.w - a.y);
    float area_b = (b.z - b.x) * (b.w - b.y);
    float uni = area_a + area_b - inter;
    return inter / uni;
}

// ---- score keys ---------------------------------------------------------------
// order-preserving float -> uint32 (larger float = larger uint); keys are
// (score_bits << 32) | (0xffffffff - index): sorting keys descending gives score
// descending, lower index first on ties (the documented tie rule).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t index) {
    return ((unsigned long long)float_to_ordered(score) << 32) | (unsigned long long)(0xffffffffu - index);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
    return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
    return ordered_to_float((uint32_t)(k >> 32));
}

}  // namespace rd

// rd_detect.cu — detect stage of RefineDet on B200 (sm_100a).
//
// Replaces (reference paths): layers/functions/detection_refinedet.py:27-113,
// eval_refinedet_coco.py:205-232, utils/nms_wrapper.py:23-31, utils/nms/nms_kernel.cu.
//
// Kernels
//   detect_forward_kernel   a3: ARM filter + two-stage decode, dense boxes/scores, in-place zeroing
//   collect_kernel          K1: one CTA per (image, slice of 1024 anchors).  ARM filter (probabilities, or
//                           logits with the model's softmax folded in); the ARM-passing anchors of an
//                           image become its NODES, numbered in anchor order.  Per node: decoded + scaled
//                           box, anchor, bin range and bin marks (graph input), and its odm_conf row
//                           transposed through shared memory into the class-major score matrix
//                           nsc[image][class][node].  odm_conf / loc rows of ARM-filtered anchors are never
//                           fetched: traffic scales with the pass rate.  No atomics on the data path.
//                           Host-mapped odm_conf (zero-copy over PCIe) is fetched as whole 128-byte lines.
//   graph_kernel            KG: 16 CTAs per image: exact suppression graph between the nodes of an image
//                           (class independent, up to 4096 nodes in blocks of 1024), adjacency lists per node
//   nms_small_kernel        K2: one CTA per (image, class): scan the class's score row, sort the candidates
//                           (beside graph_kernel: programmatic dependent launch), then resolve the
//                           suppression through the graph and emit rows.  <256 candidates, 128 threads> or,
//                           when there are few problems, <1024, 256>
//   nms_large_kernel        persistent CTAs drawing queued problems: graph resolve up to 1024 candidates for
//                           what K2 could not hold, else radix select when n > top_k + own bin tables
//   nms_single_kernel       stand-alone problem (rd_nms / rd_nms_host)
//   pack kernels            slot layout -> packed rows
#include "rd_nms_core.cuh"

#include <atomic>
#include <cmath>
#include <mutex>

namespace rd {

static std::atomic<unsigned long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

constexpr int kCollectThreads = 256;
constexpr int kSliceAnchors = 1024;   // anchors per collect CTA
#ifndef RD_LARGE_THREADS
#define RD_LARGE_THREADS 512
#endif
constexpr int kLargeThreads = RD_LARGE_THREADS;
#ifndef RD_LARGE_PER_SM
#define RD_LARGE_PER_SM 3          // resident large-problem CTAs per SM (also the kernel's register cap: 40)
#endif

// ---------------------------------------------------------------------------------------
// workspace of the fused stage.  The control block (header, gtab) must be zero when a call
// starts: rd_detect_workspace_reset zeroes it once, every call leaves it zero again (collect clears
// the queue header, nms_small_kernel's class-0 CTAs clear gtab of their image).
//   header u32 [64]               : [0] = number of queued large problems, [1] = next ticket
//   gtab   u32 [B][4][4][32][33]  : start/end bin marks of the first 4096 nodes, one table per block of 1024
//                                   nodes (OR-ed in by collect)
//   nnodes int [B]                : nodes (= ARM-passing anchors) of every image (written by collect)
//   flag   int [B]                : 1 = the image has no suppression graph (too many nodes / degree overflow)
//   queue  int [B*C]              : (image,class) problems routed to nms_large_kernel
//   nsc    f32 [B][C][Pn]         : class-major scores of the nodes, Pn = P rounded up to 32
//   nbox   f4  [B][P], nanc int [B][P] : node box (scaled) / anchor
//   ncr    u32 [B][4096]          : bin range of the first 4096 nodes
//   adjn   int [B][4096]          : graph degree of every node (collect zeroes, graph counts)
//   adj    u16 [B][4096][8]       : adjacency lists (node indices), suppressors 0 .. 7
//   adj2   u16 [B][4096][24]      : suppressors 8 .. 31 of the nodes that have them
//   cand   u64 [B*C][P]           : candidate keys of the problems nms_large_kernel handles
// ---------------------------------------------------------------------------------------
constexpr int kBlockNodes = 1024;       // suppressor nodes graph_kernel holds in shared memory at a time
constexpr int kBlockW = kBlockNodes / 32;
constexpr int kBlockWS = kBlockW + 1;   // padded row stride of the prefix-OR tables
constexpr int kBlockTab = 4 * kCols * kBlockWS;                       // words of one block's mark table
constexpr int kGtabWords = (kGraphNodes / kBlockNodes) * kBlockTab;   // block-major: the first block stays compact

struct DetectWs {
    uint32_t* header;
    uint32_t* gtab;
    int* nnodes;
    int* flag;
    int* queue;
    float* nsc;
    float4* nbox;
    int* nanc;
    uint32_t* ncr;
    int* adjn;
    uint4* adj;
    unsigned short* adj2;
    unsigned long long* cand;
    int S, Pn;
    size_t ctrl_bytes;
    size_t total;
};
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static DetectWs carve_ws(void* base, int B, int P, int C) {
    DetectWs w;
    size_t o = 0;
    unsigned char* p = static_cast<unsigned char*>(base);
    w.S = (P + kSliceAnchors - 1) / kSliceAnchors;
    w.Pn = (P + 31) & ~31;
    w.header = reinterpret_cast<uint32_t*>(p + o);             o += 256;
    w.gtab = reinterpret_cast<uint32_t*>(p + o);               o += align_up((size_t)B * kGtabWords * 4, 256);
    w.ctrl_bytes = o;
    w.nnodes = reinterpret_cast<int*>(p + o);                  o += align_up((size_t)B * 4, 256);
    w.flag = reinterpret_cast<int*>(p + o);                    o += align_up((size_t)B * 4, 256);
    w.queue = reinterpret_cast<int*>(p + o);                   o += align_up((size_t)B * C * 4, 256);
    w.nsc = reinterpret_cast<float*>(p + o);                   o += align_up((size_t)B * C * w.Pn * 4, 256);
    w.nbox = reinterpret_cast<float4*>(p + o);                 o += align_up((size_t)B * P * 16, 256);
    w.nanc = reinterpret_cast<int*>(p + o);                    o += align_up((size_t)B * P * 4, 256);
    w.ncr = reinterpret_cast<uint32_t*>(p + o);                o += align_up((size_t)B * kGraphNodes * 4, 256);
    w.adjn = reinterpret_cast<int*>(p + o);                    o += align_up((size_t)B * kGraphNodes * 4, 256);
    w.adj = reinterpret_cast<uint4*>(p + o);                   o += align_up((size_t)B * kGraphNodes * 16, 256);
    w.adj2 = reinterpret_cast<unsigned short*>(p + o);         o += align_up((size_t)B * kGraphNodes * kAdjDeg2 * 2, 256);
    w.cand = reinterpret_cast<unsigned long long*>(p + o);     o += align_up((size_t)B * C * P * 8, 256);
    w.total = o;
    return w;
}

// ---------------------------------------------------------------------------------------
// a3: Detect_RefineDet.forward (detection_refinedet.py:27-65)
// One warp per 32 consecutive (image*P + anchor) rows of the flattened batch.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCollectThreads)
detect_forward_kernel(const float4* __restrict__ arm_loc, const float2* __restrict__ arm_conf,
                      const float4* __restrict__ odm_loc, float* odm_conf,
                      const float4* __restrict__ priors, long long total, int P, int C,
                      float obj_thre, float v0, float v1, float4* __restrict__ boxes_out,
                      float* __restrict__ scores_out) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * kCollectThreads + threadIdx.x) >> 5;
    const long long g0 = warp_global * 32;
    if (g0 >= total) return;
    const long long g = g0 + lane;
    const bool valid = g < total;
    bool pass = false;
    if (valid) {
        float2 ac = ldg_stream2(arm_conf + g);
        pass = !(ac.y <= obj_thre);            // reference zeroes where arm_conf[...,1] <= thre (:41)
        const int a = (int)(g % P);
        float4 box = refine_decode(ldg_stream4(arm_loc + g), ldg_stream4(odm_loc + g), __ldg(priors + a), v0, v1);
        boxes_out[g] = box;
    }
    const unsigned mask = __ballot_sync(kFullMask, pass);
    const int nvalid = (int)min((long long)32, total - g0);
    const long long e_base = g0 * C;
    const int nelem = nvalid * C;
    float* conf = odm_conf + e_base;
    float* sc = scores_out + e_base;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(conf) | reinterpret_cast<uintptr_t>(sc)) & 15) == 0;
    const int nvec = vec_ok ? (nelem >> 2) : 0;
    // per-lane running (anchor, class) of element 4*q, advanced by 128 elements per iteration
    int e0 = lane * 4;
    int al = e0 / C, c = e0 - al * C;
    const int dal = 128 / C, dc = 128 - dal * C;
    for (int q = lane; q < nvec; q += 32) {
        // anchors of the 4 elements
        int a0 = al, c0 = c;
        unsigned pm = 0;       // pass bit per element
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pm |= ((mask >> a0) & 1u) << k;
            if (++c0 == C) { c0 = 0; ++a0; }
        }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pm) {
            v = *reinterpret_cast<const float4*>(conf + 4 * q);
            if (!(pm & 1u)) v.x = 0.f;
            if (!(pm & 2u)) v.y = 0.f;
            if (!(pm & 4u)) v.z = 0.f;
            if (!(pm & 8u)) v.w = 0.f;
        }
        *reinterpret_cast<float4*>(sc + 4 * q) = v;
        if (pm != 15u) *reinterpret_cast<float4*>(conf + 4 * q) = v;   // in-place zeroing (:42)
        al += dal; c += dc;
        if (c >= C) { c -= C; ++al; }
    }
    // scalar tail (and the whole region when the pointers are not 16-byte aligned)
    for (int e = nvec * 4 + lane; e < nelem; e += 32) {
        int a1 = e / C;
        float v = 0.f;
        if ((mask >> a1) & 1u) v = conf[e]; else conf[e] = 0.f;
        sc[e] = v;
    }
}

// The in-place half of a3 on its own (detection_refinedet.py:79-81 inside forward_python_nms): rows of odm_conf
// whose anchor fails the ARM gate become all-zero.  Warp per 32 rows, lanes stride over the rows' contiguous span.
__global__ void __launch_bounds__(kCollectThreads)
zero_filtered_rows_kernel(const float2* __restrict__ arm_conf, float* odm_conf, long long total, int C, float obj_thre) {
    const int lane = threadIdx.x & 31;
    const long long g0 = (((long long)blockIdx.x * kCollectThreads + threadIdx.x) >> 5) * 32;
    if (g0 >= total) return;
    const long long g = g0 + lane;
    const bool pass = g < total && !(ldg_stream2(arm_conf + g).y <= obj_thre);
    const unsigned mask = __ballot_sync(kFullMask, pass);
    if (mask == kFullMask) return;
    const int nelem = (int)min((long long)32, total - g0) * C;
    float* conf = odm_conf + g0 * C;
    for (int e = lane; e < nelem; e += 32)
        if (!((mask >> (e / C)) & 1u)) conf[e] = 0.f;
}

// ---------------------------------------------------------------------------------------
// K1: ARM filter + decode + node registration + class-major score matrix
// grid = (S, B); CTA (s, b) owns anchors [s*1024, s*1024 + 1024) of image b: 8 warps x 128 anchors.
// The node index of a passing anchor is its rank among the passing anchors of the image in anchor
// order (so "lower node first" == "lower anchor first", the documented tie rule): the CTA counts the
// passing anchors of the preceding slices itself (arm_conf is 8 B / anchor and L2 resident), which
// keeps the kernel free of atomics, inter-CTA ordering and state to reset.
// ---------------------------------------------------------------------------------------
constexpr int kChunks = kSliceAnchors / (kCollectThreads / 32) / 32;   // 32-anchor chunks per warp = 4
#ifndef RD_ROW_BATCH
#define RD_ROW_BATCH 4
#endif
constexpr int kRowBatch = RD_ROW_BATCH;                                // odm_conf rows in flight per warp
constexpr int kMaxClasses = 128;
constexpr int kTileRows = 64;                                          // nodes per transposition tile
constexpr int kRowsPerWarp = kTileRows / (kCollectThreads / 32);       // 8

struct GraphOut {            // what collect contributes to the per-image suppression graph
    int* nnodes;             // [B]
    uint32_t* gtab;          // [B][kGtabWords]
    float4* nbox;            // [B][P] scaled boxes
    int* nanc;               // [B][P]
    uint32_t* ncr;           // [B][kGraphNodes]
    int* adjn;               // [B][kGraphNodes]
    int* img_flag;           // [B]
    const float* img_scale;  // [B,4] or null
    float thr;
    int flags;
};

// ARM objectness gate.  Probabilities in: pass unless arm_conf[...,1] <= thre (detection_refinedet.py:41).
// Logits in (RD_INPUT_LOGITS; the softmax of models/refinedet.py:143-145 folded in): p1 =
// e1 / (e0 + e1) with e_k = exp(l_k - max), evaluated only when the logit gap d = l1 - l0 is within
// `margin` of logit(thre) — outside that band the outcome of the fp32 formula is certain.
struct ArmGate {
    float thre;
    float gap_lo, gap_hi;     // d < gap_lo: fails for sure; d > gap_hi: passes for sure (logits only)
    int admit_all;            // conf_thresh < 0: the reference zeroes the scores of ARM-filtered anchors
                              // (detection_refinedet.py:40-42) and then tests `score > conf_thresh`
                              // (eval_refinedet_coco.py:214), so filtered anchors ARE candidates, with score 0:
                              // every anchor becomes a node, filtered ones carry an all-zero score row
};
constexpr unsigned short kFlatFiltered = 0x8000;   // s_flat entry: node of an ARM-filtered anchor (admit_all only)
constexpr unsigned short kFlatMask = 0x7fff;
template <bool kLogits>
__device__ __forceinline__ bool arm_pass(float2 ac, const ArmGate& g) {
    if (!kLogits) return !(ac.y <= g.thre);
    const float d = ac.y - ac.x;
    if (d > g.gap_hi) return true;
    if (d < g.gap_lo) return false;
    const float m = fmaxf(ac.x, ac.y);
    const float e0 = expf(ac.x - m), e1 = expf(ac.y - m);
    return !(e1 / (e0 + e1) <= g.thre);
}

template <bool kLogits, bool kLines>
__global__ void __launch_bounds__(kCollectThreads)
collect_kernel(const float4* __restrict__ arm_loc, const float2* __restrict__ arm_conf,
               const float4* __restrict__ odm_loc, const float* __restrict__ odm_conf,
               const float4* __restrict__ priors, int P, int C, int S, int Pn, ArmGate gate,
               float v0, float v1, float* __restrict__ nsc, uint32_t* header, GraphOut GO) {
    // [class][node of the tile], padded; C rows of dynamic shared memory (21 KB at C = 81, not the 33 KB of
    // kMaxClasses rows: the CTAs of the other kernels of the batches in flight share the SM with this one)
    extern __shared__ __align__(16) float s_tile[];
    constexpr int kTileStride = kTileRows + 1;
    __shared__ unsigned short s_flat[kSliceAnchors];                   // passing anchors of the slice, anchor order
    __shared__ int s_wpass[kCollectThreads / 32];
    __shared__ int s_before[kCollectThreads / 32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int s = blockIdx.x, b = blockIdx.y;
    if (s == 0 && b == 0 && threadIdx.x == 0) { header[0] = 0; header[1] = 0; }   // queue of the large-NMS kernel: length, next ticket
    if (s == 0 && threadIdx.x == 0) GO.img_flag[b] = 0;
    const int a0 = s * kSliceAnchors + wib * (32 * kChunks);          // first anchor of this warp
    const size_t img = (size_t)b * P;
    // 1. ARM filter for 32*kChunks anchors; all loads issued before the first use
    float2 obj[kChunks];
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        const int a = a0 + ch * 32 + lane;
        obj[ch] = (a < P) ? ldg_stream2(arm_conf + img + a) : make_float2(0.f, 0.f);
    }
    //    ... and the number of passing anchors in the preceding slices of the image
    int before = 0;
    {
        const int nprev = s * kSliceAnchors;                           // multiple of kCollectThreads * 4
        if (gate.admit_all) {
            before = threadIdx.x == 0 ? nprev : 0;                     // every anchor is a node
        } else {
            for (int a = threadIdx.x; a < nprev; a += kCollectThreads * 4) {
                float2 o4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) o4[k] = __ldg(arm_conf + img + a + k * kCollectThreads);
#pragma unroll
                for (int k = 0; k < 4; ++k) before += arm_pass<kLogits>(o4[k], gate) ? 1 : 0;
            }
        }
        before = __reduce_add_sync(kFullMask, before);
    }
    unsigned pmask[kChunks];
    unsigned filt = 0;                       // bit ch: this lane's anchor of chunk ch is a node only because of admit_all
    int npass = 0;
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) {
        const int a = a0 + ch * 32 + lane;
        const bool arm_ok = arm_pass<kLogits>(obj[ch], gate);
        const bool pass = (a < P) && (arm_ok || gate.admit_all);
        if (pass && !arm_ok) filt |= 1u << ch;
        pmask[ch] = __ballot_sync(kFullMask, pass);
        npass += __popc(pmask[ch]);
    }
    if (lane == 0) { s_wpass[wib] = npass; s_before[wib] = before; }
    __syncthreads();
    int woff = 0, tot = 0, base = 0;
#pragma unroll
    for (int w = 0; w < kCollectThreads / 32; ++w) {
        if (w < wib) woff += s_wpass[w];
        tot += s_wpass[w];
        base += s_before[w];
    }
    {
        int pos = woff;
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
            if ((pmask[ch] >> lane) & 1u)
                s_flat[pos + __popc(pmask[ch] & ((1u << lane) - 1u))] =
                    (unsigned short)((wib * (32 * kChunks) + ch * 32 + lane) | (((filt >> ch) & 1u) ? kFlatFiltered : 0));
            pos += __popc(pmask[ch]);
        }
    }
    if (s == S - 1 && threadIdx.x == 0) GO.nnodes[b] = base + tot;
    __syncthreads();
    // decode inputs of the first kCollectThreads nodes: issue the loads now, use them after the row phase
    float4 pre_al = make_float4(0.f, 0.f, 0.f, 0.f), pre_ol = pre_al, pre_pr = pre_al;
    if ((int)threadIdx.x < tot) {
        const int a = s * kSliceAnchors + (s_flat[threadIdx.x] & kFlatMask);
        pre_al = ldg_stream4(arm_loc + img + a);
        pre_ol = ldg_stream4(odm_loc + img + a);
        pre_pr = __ldg(priors + a);
    }
    // 2. odm_conf rows of the nodes: lane = class, kRowBatch rows in flight per warp, transposed through
    //    shared memory so that nsc[b][c][base + t ...] is written in runs of consecutive nodes
    const int nseg = (C + 31) >> 5;          // <= 4 (C <= 128)
    float* nsc_b = nsc + (size_t)b * C * Pn + base;
    for (int t0 = 0; t0 < tot; t0 += kTileRows) {
        const int rows_here = min(kTileRows, tot - t0);
#pragma unroll
        for (int k0 = 0; k0 < kRowsPerWarp; k0 += kRowBatch) {
            const int rl0 = wib * kRowsPerWarp + k0;               // row of the tile
            if (rl0 >= rows_here) break;
            float v[kRowBatch][4];
            if (kLines) {
                // host-mapped odm_conf (zero-copy over PCIe): one 16-byte load per lane fetches the (up to) four
                // whole 128-byte lines that cover the row — PCIe read throughput is set by the request size
                // (measured: 42 GB/s useful this way against 31 GB/s for 4-byte lane = class loads, which
                // arrive as sector-sized requests) — and shuffles bring the values to lane = class
                float4 w[kRowBatch];
                int off[kRowBatch];
#pragma unroll
                for (int k = 0; k < kRowBatch; ++k) {
                    const unsigned short fe = rl0 + k < rows_here ? s_flat[t0 + rl0 + k] : kFlatFiltered;
                    const bool rv = !(fe & kFlatFiltered);               // filtered nodes: all-zero row, nothing fetched
                    const int a = s * kSliceAnchors + (fe & kFlatMask);
                    const float* row = odm_conf + (img + a) * C;
                    const float* line0 = reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(row) & ~(uintptr_t)127);
                    off[k] = (int)(row - line0);                               // 0..31
                    const bool need = rv && lane * 4 < off[k] + C;             // lanes past the row's last line stay idle
                    w[k] = need ? ldg_stream4(reinterpret_cast<const float4*>(line0) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < kRowBatch; ++k) {
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) {
                        const int i = off[k] + sgm * 32 + lane;                // float index inside the 512-byte span
                        const int src = (i >> 2) & 31;
                        const float x = __shfl_sync(kFullMask, w[k].x, src), y = __shfl_sync(kFullMask, w[k].y, src);
                        const float z = __shfl_sync(kFullMask, w[k].z, src), q = __shfl_sync(kFullMask, w[k].w, src);
                        const int comp = i & 3;
                        v[k][sgm] = comp == 0 ? x : comp == 1 ? y : comp == 2 ? z : q;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRowBatch; ++k) {
                    const unsigned short fe = rl0 + k < rows_here ? s_flat[t0 + rl0 + k] : kFlatFiltered;
                    const bool rv = !(fe & kFlatFiltered);
                    const int a = s * kSliceAnchors + (fe & kFlatMask);
                    const float* row = odm_conf + (img + a) * C;
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) {
                        const int c = sgm * 32 + lane;
                        v[k][sgm] = (rv && sgm < nseg && c < C) ? ldg_stream1(row + c) : 0.f;
                    }
                }
            }
            if (kLogits) {
                // softmax over the class dimension (models/refinedet.py:146-147): max, exp(x - max), sum, divide;
                // xor-butterfly reductions, so every lane holds the same max / sum and the order is fixed
#pragma unroll
                for (int k = 0; k < kRowBatch; ++k) {
                    float m = -INFINITY;
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm)
                        if (sgm < nseg && sgm * 32 + lane < C) m = fmaxf(m, v[k][sgm]);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, d));
                    float sum = 0.f;
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) {
                        const bool on = sgm < nseg && sgm * 32 + lane < C;
                        v[k][sgm] = on ? expf(v[k][sgm] - m) : 0.f;
                        sum += v[k][sgm];
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(kFullMask, sum, d);
                    const bool zero_row = gate.admit_all && rl0 + k < rows_here && (s_flat[t0 + rl0 + k] & kFlatFiltered);
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) v[k][sgm] = zero_row ? 0.f : v[k][sgm] / sum;
                }
            }
#pragma unroll
            for (int k = 0; k < kRowBatch; ++k) {
#pragma unroll
                for (int sgm = 0; sgm < 4; ++sgm) {
                    const int c = sgm * 32 + lane;
                    if (sgm < nseg && c < C) s_tile[c * kTileStride + rl0 + k] = v[k][sgm];
                }
            }
        }
        __syncthreads();
        // class 0 = background, never evaluated (eval_refinedet_coco.py:213)
        for (int task = 2 + wib; task < 2 * C; task += kCollectThreads / 32) {
            const int c = task >> 1, r = (task & 1) * 32 + lane;
            if (r < rows_here) nsc_b[c * Pn + t0 + r] = s_tile[c * kTileStride + r];
        }
        __syncthreads();
    }
    // 3. decode the nodes, one per thread: scaled box, anchor, bin range (fixed bins over the image
    //    extent: any monotone binning keeps the cull conservative) and the start / end marks OR-ed
    //    into the image's table
    const bool pixel = (GO.flags & RD_NMS_PIXEL_PLUS1) != 0;
    const bool has_scale = GO.img_scale != nullptr;
    const float4 scale = has_scale ? __ldg(reinterpret_cast<const float4*>(GO.img_scale) + b) : make_float4(1.f, 1.f, 1.f, 1.f);
    const bool force_full = cull_disabled(GO.thr, GO.flags);
    const float invx = scale.x > 0.f ? (float)kCols / scale.x : 0.f;
    const float invy = scale.y > 0.f ? (float)kCols / scale.y : 0.f;
    for (int t = threadIdx.x; t < tot; t += kCollectThreads) {
        const int a = s * kSliceAnchors + (s_flat[t] & kFlatMask);
        float4 bx = t < kCollectThreads ? refine_decode(pre_al, pre_ol, pre_pr, v0, v1)
                                        : refine_decode(ldg_stream4(arm_loc + img + a), ldg_stream4(odm_loc + img + a),
                                                        __ldg(priors + a), v0, v1);
        if (has_scale) { bx.x *= scale.x; bx.y *= scale.y; bx.z *= scale.z; bx.w *= scale.w; }
        const int i = base + t;
        GO.nbox[img + i] = bx;
        GO.nanc[img + i] = a;
        if (i < kGraphNodes) {
            const uint32_t cr = bin_range(bx.x, bx.y, bx.z, bx.w, pixel, force_full, 0.f, invx, 0.f, invy);
            GO.ncr[(size_t)b * kGraphNodes + i] = cr;
            GO.adjn[(size_t)b * kGraphNodes + i] = 0;            // degree counter of the suppression graph
            const int ax = cr & 255u, bxx = (cr >> 8) & 255u, ay = (cr >> 16) & 255u, by = cr >> 24;
            const uint32_t bit = 1u << (i & 31);
            const int w = (i >> 5) & (kBlockW - 1);
            uint32_t* tab = GO.gtab + (size_t)b * kGtabWords + (i / kBlockNodes) * kBlockTab;
            atomicOr(&tab[(0 * kCols + ax) * kBlockWS + w], bit);
            if (bxx + 1 < kCols) atomicOr(&tab[(1 * kCols + bxx + 1) * kBlockWS + w], bit);
            atomicOr(&tab[(2 * kCols + ay) * kBlockWS + w], bit);
            if (by + 1 < kCols) atomicOr(&tab[(3 * kCols + by + 1) * kBlockWS + w], bit);
        }
    }
}

// ---------------------------------------------------------------------------------------
// KG: suppression graph of one image.  grid = (kGraphSplit, B).  The nodes are processed in blocks of
// kBlockNodes potential suppressors: every CTA loads the block's boxes and the block's columns of the
// mark table (one round of coalesced L2 loads), finishes the prefix-OR (S[c] = starts <= c, E[c] =
// ends < c), and — for the nodes it owns (every kGraphSplit-th) that come AFTER a node of the block —
// lists the pairs (i < j) that survive the bin cull and tests them exactly.
//   adj[node j] = { node u : suppresses(kept = u, candidate = j) }    (exact fp32 test, boxes scaled)
// One block (<= 1024 nodes) is the common case; images with more than kGraphNodes nodes, a node of
// degree > kAdjDeg are flagged and handled by the per-problem bin path instead.
// ---------------------------------------------------------------------------------------
#ifndef RD_GRAPH_THREADS
#define RD_GRAPH_THREADS 256
#endif
constexpr int kGraphThreads = RD_GRAPH_THREADS;
constexpr int kGraphSplit = 16;         // CTAs per image
constexpr int kGraphPairCap = 2048;

struct GraphSmem {
    float x1[kBlockNodes], y1[kBlockNodes], x2[kBlockNodes], y2[kBlockNodes];
    // (no array of box areas: the area-ratio pre-test of the multi-block path recomputes them.  With it a CTA took
    //  49.7 + 1 KB and only THREE of the kernel's four CTAs per SM fit the 196 KB carve-out the chain runs best at;
    //  at 45.6 + 1 KB four fit: 27.1 -> 25.8 us per batch with four batches in flight, 49 -> 45 us for one batch alone,
    //  for + 2 - 5 % on images of more than 1024 nodes)
    uint32_t cr[kBlockNodes];
    uint32_t tab[4 * kCols * kBlockWS];
    uint32_t pairs[kGraphPairCap];
    int overflow;
    int npairs;
};

// an edge is added in two steps so that the atomics of several edges can be in flight before their stores
__device__ __forceinline__ void graph_store_edge(unsigned short* __restrict__ adj, unsigned short* __restrict__ adj2,
                                                 int to, int from, int slot, int* overflow) {
    if (slot < kAdjDeg) adj[to * kAdjDeg + slot] = (unsigned short)from;
    else if (slot < kAdjMax) {
        adj2[to * kAdjDeg2 + slot - kAdjDeg] = (unsigned short)from;
        if (slot == kAdjDeg) atomicOr(overflow, kFlagWideDeg);
    } else if (slot == kAdjMax) atomicOr(overflow, kFlagNoGraph);
}

// exact test of the unordered pair (block node i, own node n0 + jl), both directions
__device__ __forceinline__ void graph_test_pair_inline(GraphSmem& G, int i, int jl, int nb, const float4* __restrict__ boxes_n0,
                                                       float thr, int flags, int* __restrict__ adjn,
                                                       unsigned short* __restrict__ adj, unsigned short* __restrict__ adj2, int n0) {
    float4 bj;
    if (jl < nb) bj = make_float4(G.x1[jl], G.y1[jl], G.x2[jl], G.y2[jl]);
    else bj = __ldg(boxes_n0 + jl);
    const bool i_sup_j = suppresses(G.x1[i], G.y1[i], G.x2[i], G.y2[i], bj.x, bj.y, bj.z, bj.w, thr, flags);
    // the pixel(+1) IoU is symmetric in fp32 (ai + aj commutes); the normalised one is not ((aj - inter) + ai)
    const bool j_sup_i = (flags & RD_NMS_PIXEL_PLUS1)
                             ? i_sup_j
                             : suppresses(bj.x, bj.y, bj.z, bj.w, G.x1[i], G.y1[i], G.x2[i], G.y2[i], thr, flags);
    int s_j = 0, s_i = 0;
    if (i_sup_j) s_j = atomicAdd(&adjn[n0 + jl], 1);
    if (j_sup_i) s_i = atomicAdd(&adjn[n0 + i], 1);
    if (i_sup_j) graph_store_edge(adj, adj2, n0 + jl, n0 + i, s_j, &G.overflow);
    if (j_sup_i) graph_store_edge(adj, adj2, n0 + i, n0 + jl, s_i, &G.overflow);
}
// out-of-line twin for the (rare) pairs found after the list has filled up: keeps the listing loop lean
__device__ __noinline__ void graph_test_pair(GraphSmem& G, int i, int jl, int nb, const float4* __restrict__ boxes_n0,
                                             float thr, int flags, int* __restrict__ adjn,
                                             unsigned short* __restrict__ adj, unsigned short* __restrict__ adj2, int n0) {
    graph_test_pair_inline(G, i, jl, nb, boxes_n0, thr, flags, adjn, adj, adj2, n0);
}

// the block loop of graph_kernel.  kSingle: the image has at most kBlockNodes nodes (the common case) — one
// block, every own node inside it, nothing read from global memory after the staging.
template <bool kSingle>
__device__ __forceinline__ void graph_image(GraphSmem& G, int N, int g, int b, int tid, const uint32_t* __restrict__ gtab,
                                            const float4* __restrict__ nbox, const uint32_t* __restrict__ ncr, int P,
                                            float thr, int flags, uint4* __restrict__ adj_all,
                                            unsigned short* __restrict__ adj2_all, int* __restrict__ adjn_all) {
    int* adjn = adjn_all + (size_t)b * kGraphNodes;
    unsigned short* adj = reinterpret_cast<unsigned short*>(adj_all + (size_t)b * kGraphNodes);
    unsigned short* adj2 = adj2_all + (size_t)b * kGraphNodes * kAdjDeg2;
    const float4* boxes = nbox + (size_t)b * P;
    const uint32_t* crs = ncr + (size_t)b * kGraphNodes;
    const uint32_t* gt = gtab + (size_t)b * kGtabWords;
    const uint32_t* Sx = G.tab;
    const uint32_t* Ex = G.tab + 1 * kCols * kBlockWS;
    const uint32_t* Sy = G.tab + 2 * kCols * kBlockWS;
    const uint32_t* Ey = G.tab + 3 * kCols * kBlockWS;
    static_assert(kGraphSplit == 16, "item enumeration assumes 2 own nodes per 32-node word");
    const bool pixel = (flags & RD_NMS_PIXEL_PLUS1) != 0;
    // Area-ratio pre-test.  The intersection of two boxes is at most the smaller area and their union at least the
    // larger one, so IoU <= min(a_i, a_j) / max(a_i, a_j): a bin-surviving pair whose ratio is below the threshold
    // (with a 2^-12 margin, far above the rounding of the fp32 IoU) cannot be an edge and is neither listed nor
    // tested.  Anchors come in scales a factor 2 apart (areas a factor 4), so most pairs a large box forms with the
    // small boxes under it go this way.  Only for positive finite areas; never when the cull is disabled.  Used for
    // images of more than one block of nodes only: measured, it takes 10 % off graph_kernel at 1.9 - 3.9 k nodes per
    // image (208 -> 185 us) but ADDS 1 us at the 700 nodes of the headline workload, where few pairs survive the bins.
    const float ratio_thr = cull_disabled(thr, flags) ? -1.0f : thr * (1.0f - 2.44140625e-4f);
    if (tid == 0) G.overflow = 0;
    for (int n0 = 0; n0 < (kSingle ? 1 : N); n0 += kBlockNodes) {
        const int nb = kSingle ? N : min(kBlockNodes, N - n0);             // suppressor candidates i = n0 .. n0 + nb - 1
        const int Wb = (nb + 31) >> 5;
        if (n0 > 0) __syncthreads();                         // the previous block's tables are no longer read
        // 1. the block's boxes + mark-table columns -> shared memory (independent loads, one round trip)
        if (tid == 0) G.npairs = 0;
        for (int i = tid; i < nb; i += kGraphThreads) {
            const float4 bx = boxes[n0 + i];
            G.x1[i] = bx.x; G.y1[i] = bx.y; G.x2[i] = bx.z; G.y2[i] = bx.w;
            G.cr[i] = crs[n0 + i];
        }
        // 2. inclusive prefix-OR over the bins, straight from the global marks.  One (table, word) column per
        //    thread, 32 independent loads each.
        for (int task = tid; task < 4 * Wb; task += kGraphThreads) {
            const int t = task / Wb, w = task - t * Wb;
            uint32_t v[kCols];
#pragma unroll
            for (int c = 0; c < kCols; ++c) v[c] = __ldg(gt + (n0 / kBlockNodes) * kBlockTab + (t * kCols + c) * kBlockWS + w);
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < kCols; ++c) { acc |= v[c]; G.tab[(t * kCols + c) * kBlockWS + w] = acc; }
        }
        __syncthreads();
        // 3. pairs (i in the block) < (own node j) that survive the bin cull.  Work item = (own node, mask word).
        //    Own nodes inside the block, jo = 0 .. nin-1 (j = n0 + g + 16 jo), only need the words 0 .. jo >> 1:
        //    the pair of own nodes (2m, 2m + 1) has m + 1 words each and m (m + 1) items precede it.  Own nodes
        //    of later blocks, jo = nin .., need all Wb words.
        const int nin = nb > g ? (nb - g + kGraphSplit - 1) / kGraphSplit : 0;
        const int nall = (N - n0 - g + kGraphSplit - 1) / kGraphSplit;          // own nodes with j >= n0 (N - n0 > g here)
        const int mfull = nin >> 1;
        const int ntri = mfull * (mfull + 1) + ((nin & 1) ? mfull + 1 : 0);
        const int nitems = ntri + ((!kSingle && N - n0 > g) ? (nall - nin) * Wb : 0);
        for (int q0 = 0; q0 < nitems; q0 += kGraphThreads) {
            const int q = q0 + tid;
            uint32_t h = 0;
            int jo = 0, w = 0;
            if (q < nitems) {
                uint32_t cr;
                if (kSingle || q < ntri) {
                    int m = (int)((sqrtf((float)(4 * q + 1)) - 1.0f) * 0.5f);
                    while (m * (m + 1) > q) --m;
                    while ((m + 1) * (m + 2) <= q) ++m;
                    const int rem = q - m * (m + 1);
                    const int second = rem > m ? 1 : 0;
                    jo = 2 * m + second;
                    w = rem - second * (m + 1);
                    cr = G.cr[g + jo * kGraphSplit];
                } else {
                    const int r = q - ntri;
                    jo = nin + r / Wb;
                    w = r - (jo - nin) * Wb;
                    cr = __ldg(crs + n0 + g + jo * kGraphSplit);
                }
                h = Sx[((cr >> 8) & 255u) * kBlockWS + w] & ~Ex[(cr & 255u) * kBlockWS + w] &
                    Sy[((cr >> 24) & 255u) * kBlockWS + w] & ~Ey[((cr >> 16) & 255u) * kBlockWS + w];
                const int jl = g + jo * kGraphSplit;                             // j - n0
                if (w == (jl >> 5)) h &= (1u << (jl & 31)) - 1u;                 // predecessors only
                if (!kSingle && h && ratio_thr > 0.0f) {
                    float aj;
                    if (jl < nb) aj = box_area(G.x1[jl], G.y1[jl], G.x2[jl], G.y2[jl], pixel);
                    else { const float4 bj = __ldg(boxes + n0 + jl); aj = box_area(bj.x, bj.y, bj.z, bj.w, pixel); }
                    uint32_t rest = h;
                    while (rest) {
                        const int bit = __ffs(rest) - 1;
                        rest &= rest - 1;
                        const int ii = (w << 5) + bit;
                        const float ai = box_area(G.x1[ii], G.y1[ii], G.x2[ii], G.y2[ii], pixel);
                        const float lo = fminf(ai, aj), hi = fmaxf(ai, aj);
                        if (lo > 0.0f && lo < ratio_thr * hi) h &= ~(1u << bit);
                    }
                }
            }
            // reserve list slots: one shared-memory atomic per item that has pairs (list order is irrelevant);
            // pairs that do not fit the list any more are tested on the spot
            int off = h ? atomicAdd(&G.npairs, __popc(h)) : 0;
            const uint32_t tag = (uint32_t)jo << 16;
            while (h) {
                const int i = (w << 5) + __ffs(h) - 1;
                h &= h - 1;
                if (off < kGraphPairCap) G.pairs[off] = tag | (uint32_t)i;
                else graph_test_pair(G, i, g + jo * kGraphSplit, kSingle ? kBlockNodes : nb, boxes + n0, thr, flags, adjn, adj, adj2, n0);
                ++off;
            }
            // dense neighbourhoods (dozens of overlapping boxes per object): drain the list between rounds of
            // items instead of letting it fill up and testing the excess pairs one thread at a time
            //  The decision is taken CTA-wide by the barrier itself (__syncthreads_or): a thread that read npairs
            //  on its own could skip the branch and append pairs of the next round before a slower warp has
            //  looked at the counter, and the barriers inside the branch would pair up at different places.
#ifndef RD_GRAPH_DRAIN_EVERY
#define RD_GRAPH_DRAIN_EVERY 4     // 1: 57.4 us at 1871 nodes per image, 4: 51.6 us; 8 and 16 let the list overflow at 3.9 k nodes (185 / 168 / 188 / 203 us)
#endif
            // (multi-block images: the check -- a CTA-wide barrier -- every RD_GRAPH_DRAIN_EVERY-th round of items)
            if (q0 + kGraphThreads < nitems &&
                (kSingle || RD_GRAPH_DRAIN_EVERY == 1 || ((q0 / kGraphThreads) % RD_GRAPH_DRAIN_EVERY) == RD_GRAPH_DRAIN_EVERY - 1)) {
                if (__syncthreads_or(G.npairs >= kGraphPairCap / 2)) {
                    const int cnt = min(G.npairs, kGraphPairCap);
                    for (int p = tid; p < cnt; p += kGraphThreads) {
                        const uint32_t e = G.pairs[p];
                        graph_test_pair_inline(G, (int)(e & 0xffffu), g + (int)(e >> 16) * kGraphSplit, kSingle ? kBlockNodes : nb,
                                               boxes + n0, thr, flags, adjn, adj, adj2, n0);
                    }
                    __syncthreads();
                    if (tid == 0) G.npairs = 0;
                    __syncthreads();
                }
            }
        }
        __syncthreads();
        // 4. exact tests of the listed pairs, both directions
        const int cnt = min(G.npairs, kGraphPairCap);
        for (int p = tid; p < cnt; p += kGraphThreads) {
            const uint32_t e = G.pairs[p];
            graph_test_pair_inline(G, (int)(e & 0xffffu), g + (int)(e >> 16) * kGraphSplit, kSingle ? kBlockNodes : nb,
                                   boxes + n0, thr, flags, adjn, adj, adj2, n0);
        }
    }
}

__global__ void __launch_bounds__(kGraphThreads)
graph_kernel(const int* __restrict__ nnodes, const uint32_t* __restrict__ gtab, const float4* __restrict__ nbox,
             const uint32_t* __restrict__ ncr, int P, float thr, int flags,
             uint4* __restrict__ adj_all, unsigned short* __restrict__ adj2_all, int* __restrict__ adjn_all,
             int* __restrict__ img_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GraphSmem& G = *reinterpret_cast<GraphSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = blockIdx.x, b = blockIdx.y;
    grid_dependency_wait();          // collect_kernel has completed (this kernel is launched early)
    grid_launch_dependents();        // nms_small_kernel's scan + sort (graph independent) may run beside this kernel
    const int N = nnodes[b];
    if (N > kGraphNodes) {
        if (g == 0 && tid == 0) img_flag[b] = kFlagNoGraph;
        return;
    }
    // own nodes: j = g, g + kGraphSplit, ... (interleaved, so the few large boxes that overlap hundreds
    // of others, and the later nodes that have more predecessors, are spread over the CTAs of the image)
    if (N <= g) return;
    if (N <= kBlockNodes) graph_image<true>(G, N, g, b, tid, gtab, nbox, ncr, P, thr, flags, adj_all, adj2_all, adjn_all);
    else graph_image<false>(G, N, g, b, tid, gtab, nbox, ncr, P, thr, flags, adj_all, adj2_all, adjn_all);
    __syncthreads();
    if (tid == 0 && G.overflow) atomicOr(&img_flag[b], G.overflow);
}

// ---------------------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------------------
struct FusedNmsArgs {
    const float* nsc;                // [B][C][Pn] node scores
    const float4* nbox;              // [B][P] node boxes (scaled)
    const int* nanc;                 // [B][P] node anchors
    const int* nnodes;               // [B]
    uint32_t* gtab;                  // [B][kGtabWords]   control block, cleared by the class-0 CTA of every image
    const int* img_flag;             // [B] kFlagNoGraph | kFlagWideDeg
    const uint4* adj;                // [B][kGraphNodes]
    const unsigned short* adj2;      // [B][kGraphNodes][kAdjDeg2]
    const int* adjn;                 // [B][kGraphNodes]
    int* queue;
    uint32_t* header;
    unsigned long long* cand;        // [B*C][P] key scratch of the large problems
    int nbc, C, P, Pn;
    int large_grid, large_mcap, large_smem;   // launch shape of nms_large_kernel
    float conf_thresh;
    float thr;
    int top_k, max_out, flags, row_layout;
    int* out_counts;
    float* out_dets;
    int* out_anchor;
};

// K2: one CTA per (image, class).  Launched with programmatic stream serialisation behind graph_kernel
// (which triggers at its start): the scan of the class's score row and the sort need collect_kernel's
// results only and run beside graph_kernel; the CTA then waits for the graph and resolves.
#ifndef RD_SMALL_MINBLOCKS
#define RD_SMALL_MINBLOCKS (1536 / RD_SMALL_THREADS)
#endif
// Two instantiations: <256 candidates, 128 threads> for the many-class case (thousands of problems, occupancy
// matters) and <1024, 256> when the grid cannot fill the GPU anyway (few classes: every problem holds a large
// share of the image's nodes, e.g. the 2-class SAR-ship configuration).
template <int kCap, int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
nms_small_kernel(FusedNmsArgs A) {
    __shared__ SmallSmem<kCap> S;
    constexpr int kPerT = (kBlockNodes + kThreads - 1) / kThreads;     // register-resident scan: N <= 1024
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x, b = blockIdx.y;
    const int bc = b * A.C + c;
    if (c == 0) {
        // this CTA has no problem (background is never evaluated, eval_refinedet_coco.py:213): it leaves the
        // graph control block of its image zero for the next call, once graph_kernel is done with it
        grid_dependency_wait();
        uint32_t* gt = A.gtab + (size_t)b * kGtabWords;
        const int nblk = (min(A.nnodes[b], kGraphNodes) + kBlockNodes - 1) / kBlockNodes;   // tables collect_kernel marked
        for (int i = tid; i < nblk * kBlockTab; i += kThreads) gt[i] = 0;
        if (tid == 0) A.out_counts[bc] = 0;
        return;
    }
    const int N = A.nnodes[b];
    if (N > SmallSmem<kCap>::kMaxNodes) {    // more nodes than this variant's rank table covers (or no graph at all):
        if (tid == 0) A.queue[atomicAdd(&A.header[0], 1u)] = bc;      // nms_large_kernel resolves or bins it
        return;
    }
    // candidates: nodes whose score exceeds the threshold.  Two passes (count, then place) with one barrier
    // in between: no atomics.  Up to 1024 nodes the scores stay in registers between the passes.
    const float* row = A.nsc + (size_t)bc * A.Pn;
    const int nq = (N + kThreads - 1) / kThreads;
    const unsigned lt = (1u << lane) - 1u;
    if (SmallSmem<kCap>::kMaxNodes <= kPerT * kThreads || nq <= kPerT) {      // always, for the common variant
        float v[kPerT];
        unsigned bal[kPerT];
#pragma unroll
        for (int q = 0; q < kPerT; ++q) {
            const int i = q * kThreads + tid;
            v[q] = (q < nq && i < N) ? __ldg(row + i) : -INFINITY;
        }
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < kPerT; ++q) {
            bal[q] = 0;
            if (q < nq) {
                bal[q] = __ballot_sync(kFullMask, v[q] > A.conf_thresh);
                cnt += __popc(bal[q]);
            }
        }
        if (lane == 0) S.wsum[warp] = cnt;
        __syncthreads();
        int slot = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            if (w < warp) slot += S.wsum[w];
            tot += S.wsum[w];
        }
        if (tid == 0) S.n = tot;
        if (tot <= kCap) {
#pragma unroll
            for (int q = 0; q < kPerT; ++q) {
                if (q < nq && bal[q]) {
                    if ((bal[q] >> lane) & 1u)
                        S.u.runs[slot + __popc(bal[q] & lt)] = make_key(v[q], (uint32_t)(q * kThreads + tid));
                    slot += __popc(bal[q]);
                }
            }
        }
    } else {                                                  // 1024 < N <= kGraphNodes: the row is read twice (L2)
        int cnt = 0;
        for (int q = 0; q < nq; ++q) {
            const int i = q * kThreads + tid;
            const float v = i < N ? __ldg(row + i) : -INFINITY;
            cnt += __popc(__ballot_sync(kFullMask, v > A.conf_thresh));
        }
        if (lane == 0) S.wsum[warp] = cnt;
        __syncthreads();
        int slot = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) {
            if (w < warp) slot += S.wsum[w];
            tot += S.wsum[w];
        }
        if (tid == 0) S.n = tot;
        if (tot <= kCap) {
            for (int q = 0; q < nq; ++q) {
                const int i = q * kThreads + tid;
                const float v = i < N ? __ldg(row + i) : -INFINITY;
                const unsigned bal = __ballot_sync(kFullMask, v > A.conf_thresh);
                if ((bal >> lane) & 1u) S.u.runs[slot + __popc(bal & lt)] = make_key(v, (uint32_t)i);
                slot += __popc(bal);
            }
        }
    }
    __syncthreads();
    const int n = S.n;
    if (n == 0) {
        if (tid == 0) A.out_counts[bc] = 0;
        return;
    }
    if (n > A.top_k || n > kCap) {                           // needs the top-k select: large kernel
        if (tid == 0) A.queue[atomicAdd(&A.header[0], 1u)] = bc;
        return;
    }
    cta_sort_small<kThreads, kCap>(S, n, A.conf_thresh);
    grid_dependency_wait();                                       // graph_kernel has completed
    const int fl = A.img_flag[b];                                 // degree overflow: no graph (for this variant) after all
    if ((fl & kFlagNoGraph) || ((fl & kFlagWideDeg) && !SmallSmem<kCap>::kBitRows)) {
        if (tid == 0) A.queue[atomicAdd(&A.header[0], 1u)] = bc;
        return;
    }
    RowSink sink;
    sink.rows = A.out_dets + (size_t)bc * A.max_out * 5;
    sink.anchors = A.out_anchor ? A.out_anchor + (size_t)bc * A.max_out : nullptr;
    sink.keep64 = nullptr; sink.keep32 = nullptr; sink.idx_map = nullptr;
    sink.row_layout = A.row_layout;
    GraphView G;
    G.adj = A.adj + (size_t)b * kGraphNodes;
    G.adj2 = A.adj2 + (size_t)b * kGraphNodes * kAdjDeg2;
    G.adjn = A.adjn + (size_t)b * kGraphNodes;
    G.nbox = A.nbox + (size_t)b * A.P;
    G.nanc = A.nanc + (size_t)b * A.P;
    const int kept = cta_nms_graph<kThreads, kCap>(S, n, N, A.max_out, sink, G, (fl & kFlagWideDeg) != 0);
    if (tid == 0) A.out_counts[bc] = kept;
}

// A queued problem of an image that has a suppression graph and only outgrew nms_small_kernel's 256 candidates
// (or its 1024-node rank table): the same sort + graph resolve, 1024 candidates wide, on the large kernel's
// shared memory.  Out of line, so that the register allocation of the bin path (nms_process) is unaffected.
__device__ __noinline__ int large_graph_resolve(unsigned char* smem, const unsigned long long* __restrict__ keys, int n,
                                                float conf_thresh, int N, int max_out, const uint4* adj, const int* adjn,
                                                const float4* nbox, const int* nanc, float* rows, int* anchors,
                                                int row_layout) {
    SmallSmem<kWideCap>& S = *reinterpret_cast<SmallSmem<kWideCap>*>(smem);
    for (int i = threadIdx.x; i < n; i += kLargeThreads) S.u.runs[i] = keys[i];
    __syncthreads();
    cta_sort_small<kLargeThreads, kWideCap>(S, n, conf_thresh);
    GraphView G;
    G.adj = adj; G.adj2 = nullptr; G.adjn = adjn; G.nbox = nbox; G.nanc = nanc;    // kWideCap: kDeps == kAdjDeg
    RowSink sink;
    sink.rows = rows; sink.anchors = anchors; sink.keep64 = nullptr; sink.keep32 = nullptr; sink.idx_map = nullptr;
    sink.row_layout = row_layout;
    return cta_nms_graph<kLargeThreads, kWideCap>(S, n, N, max_out, sink, G);
}

// The common kind of queued problem -- an image with a graph whose 1025 .. 4096 nodes outgrew nms_small_kernel's rank
// table, at most kMidCap candidates: candidates are compacted from the score row straight into shared memory (no
// round trip through the global key list), bucket-sorted and resolved through the graph, one candidate per thread.
// Returns -1 (nothing written) when the problem has more than kMidCap / top_k candidates.
__device__ __noinline__ int large_mid_resolve(unsigned char* smem, const float* __restrict__ row, int N, float conf_thresh,
                                              int top_k, int max_out, const uint4* adj, const int* adjn, const float4* nbox,
                                              const int* nanc, float* rows, int* anchors, int row_layout) {
    SmallSmem<kMidCap>& S = *reinterpret_cast<SmallSmem<kMidCap>*>(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) S.n = 0;
    __syncthreads();
    constexpr int kU = 4;
    for (int i0 = 0; i0 < N; i0 += kLargeThreads * kU) {
        float v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int i = i0 + u * kLargeThreads + tid;
            v[u] = i < N ? __ldg(row + i) : -INFINITY;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const bool pass = v[u] > conf_thresh;
            const unsigned bal = __ballot_sync(kFullMask, pass);
            if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&S.n, __popc(bal));
                base = __shfl_sync(kFullMask, base, 0) + __popc(bal & ((1u << lane) - 1u));
                if (pass && base < kMidCap) S.u.runs[base] = make_key(v[u], (uint32_t)(i0 + u * kLargeThreads + tid));
            }
        }
    }
    __syncthreads();
    const int n = S.n;
    __syncthreads();                                      // S.n is rewritten by the next problem of this CTA
    if (n > kMidCap || n > top_k) return -1;
    if (n == 0) return 0;
    cta_sort_small<kLargeThreads, kMidCap>(S, n, conf_thresh);
    GraphView G;
    G.adj = adj; G.adj2 = nullptr; G.adjn = adjn; G.nbox = nbox; G.nanc = nanc;
    RowSink sink;
    sink.rows = rows; sink.anchors = anchors; sink.keep64 = nullptr; sink.keep32 = nullptr; sink.idx_map = nullptr;
    sink.row_layout = row_layout;
    return cta_nms_graph<kLargeThreads, kMidCap>(S, n, N, max_out, sink, G);
}

__global__ void __launch_bounds__(kLargeThreads, RD_LARGE_PER_SM)      // 3 x 512 threads: 40 registers
nms_large_kernel(FusedNmsArgs A, int mcap) {
    extern __shared__ __align__(16) unsigned char smem[];
    grid_dependency_wait();          // nms_small_kernel (and everything before it) has completed
    const uint32_t nq = A.header[0];
    if (nq == 0) return;             // the common case: nothing queued — leave before any set-up work
    int* s_cnt;
    bool wide_fits;                  // the 1024-candidate graph resolve fits this launch's shared memory
    {
        const NmsSmemLayout L0 = nms_layout(mcap);
        s_cnt = reinterpret_cast<int*>(smem + L0.off_misc) + 15;     // misc[15]: unused by nms_process
        wide_fits = sizeof(SmallSmem<kWideCap>) <= L0.total;
    }
    const int tid = threadIdx.x, lane = tid & 31;
    // dynamic tickets: problems differ a lot in cost (select passes, pairs), a static stride leaves CTAs idle
    for (;;) {
        __shared__ uint32_t s_ticket;
        if (tid == 0) s_ticket = atomicAdd(&A.header[1], 1u);
        __syncthreads();
        const uint32_t q = s_ticket;
        __syncthreads();
        if (q >= nq) break;
        const int bc = A.queue[q];
        const int b = bc / A.C;
        const int N = A.nnodes[b];
#ifndef RD_NO_LARGE_GRAPH
        if (wide_fits && N <= kGraphNodes && A.img_flag[b] == 0) {
            const int kept = large_mid_resolve(smem, A.nsc + (size_t)bc * A.Pn, N, A.conf_thresh, A.top_k, A.max_out,
                                               A.adj + (size_t)b * kGraphNodes, A.adjn + (size_t)b * kGraphNodes,
                                               A.nbox + (size_t)b * A.P, A.nanc + (size_t)b * A.P,
                                               A.out_dets + (size_t)bc * A.max_out * 5,
                                               A.out_anchor ? A.out_anchor + (size_t)bc * A.max_out : nullptr, A.row_layout);
            if (kept >= 0) {
                if (tid == 0) A.out_counts[bc] = kept;
                __syncthreads();
                continue;
            }
        }
#endif
        // candidate keys of the problem -> its slot of the global scratch list
        unsigned long long* keys = A.cand + (size_t)bc * A.P;
        const float* row = A.nsc + (size_t)bc * A.Pn;
        if (tid == 0) s_cnt[0] = 0;
        __syncthreads();
        constexpr int kScanUnroll = 4;                            // independent loads in flight per thread
        for (int i0 = 0; i0 < N; i0 += kLargeThreads * kScanUnroll) {
            float v[kScanUnroll];
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                const int i = i0 + u * kLargeThreads + tid;
                v[u] = i < N ? __ldg(row + i) : -INFINITY;
            }
#pragma unroll
            for (int u = 0; u < kScanUnroll; ++u) {
                const int i = i0 + u * kLargeThreads + tid;
                const bool pass = v[u] > A.conf_thresh;
                const unsigned bal = __ballot_sync(kFullMask, pass);
                if (bal) {
                    int wbase = 0;
                    if (lane == 0) wbase = atomicAdd(&s_cnt[0], __popc(bal));
                    wbase = __shfl_sync(kFullMask, wbase, 0);
                    if (pass) keys[wbase + __popc(bal & ((1u << lane) - 1u))] = make_key(v[u], (uint32_t)i);
                }
            }
        }
        __syncthreads();
        const int n = s_cnt[0];
        __syncthreads();
        RowSink sink;
        sink.rows = A.out_dets + (size_t)bc * A.max_out * 5;
        sink.anchors = A.out_anchor ? A.out_anchor + (size_t)bc * A.max_out : nullptr;
        sink.keep64 = nullptr; sink.keep32 = nullptr;
        sink.idx_map = A.nanc + (size_t)b * A.P;
        sink.row_layout = A.row_layout;
        int kept;
#ifndef RD_NO_LARGE_GRAPH
        if (wide_fits && N <= kGraphNodes && A.img_flag[b] == 0 && n <= kWideCap && n <= A.top_k) {
            kept = large_graph_resolve(smem, keys, n, A.conf_thresh, N, A.max_out, A.adj + (size_t)b * kGraphNodes,
                                       A.adjn + (size_t)b * kGraphNodes, A.nbox + (size_t)b * A.P, A.nanc + (size_t)b * A.P,
                                       sink.rows, sink.anchors, sink.row_layout);
        } else
#endif
        {
            NmsProblem pb;
            pb.cl.base = keys; pb.cl.n = n;
            pb.boxes = A.nbox + (size_t)b * A.P;
            pb.has_scale = 0;                                     // node boxes are already scaled
            pb.scale = make_float4(1.f, 1.f, 1.f, 1.f);
            pb.thr = A.thr; pb.top_k = A.top_k; pb.max_out = A.max_out; pb.flags = A.flags;
            pb.dets = nullptr; pb.dets_dim = 0;
            const NmsSmemLayout L = nms_layout(mcap);               // recomputed here: not kept live across the loop
            kept = nms_process(smem, L, pb, sink);
        }
        if (tid == 0) A.out_counts[bc] = kept;
        __syncthreads();
    }
}

// stand-alone problem
__global__ void make_keys_kernel(const float* __restrict__ scores, int n, unsigned long long* keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = make_key(scores[i], (uint32_t)i);
}
__global__ void __launch_bounds__(kLargeThreads)
nms_single_kernel(const unsigned long long* cand, int n, const float4* boxes, float thr, int top_k, int max_out,
                  int flags, int mcap, long long* keep_out, int* keep_out32, int* count_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const NmsSmemLayout L = nms_layout(mcap);
    NmsProblem pb;
    pb.cl.base = cand; pb.cl.n = n;
    pb.boxes = boxes; pb.has_scale = 0;
    pb.scale = make_float4(1.f, 1.f, 1.f, 1.f);
    pb.thr = thr; pb.top_k = top_k; pb.max_out = max_out; pb.flags = flags;
    pb.dets = nullptr; pb.dets_dim = 0;
    RowSink sink;
    sink.rows = nullptr; sink.anchors = nullptr; sink.keep64 = keep_out; sink.keep32 = keep_out32; sink.row_layout = 0;
    sink.idx_map = nullptr;
    const int kept = nms_process(smem, L, pb, sink);
    if (threadIdx.x == 0) *count_out = kept;
}

// rd_nms_host: `io` is the call's pinned, mapped host buffer  [n*dim floats, padded to 16 B | count | keep[n]] --
// the rows (already score-descending, the contract of _nms) are read and the kept positions written straight
// over PCIe, so the whole call is one launch + one stream synchronisation
constexpr int kHostThreads = 1024;      // one CTA owns the SM: the walk is issue-bound, more warps hide its latencies
__global__ void __launch_bounds__(kHostThreads)
nms_host_kernel(const float* dets, int n, int dim, float thr, int flags, int* count_out, int* keep_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    const NmsSmemLayout L = nms_layout(n);
    NmsProblem pb;
    pb.cl.base = nullptr; pb.cl.n = n;
    pb.boxes = nullptr; pb.has_scale = 0;
    pb.scale = make_float4(1.f, 1.f, 1.f, 1.f);
    pb.thr = thr; pb.top_k = n; pb.max_out = n; pb.flags = flags;
    pb.dets = dets; pb.dets_dim = dim;
    RowSink sink;
    sink.rows = nullptr; sink.anchors = nullptr; sink.keep64 = nullptr; sink.keep32 = keep_out; sink.row_layout = 0;
    sink.idx_map = nullptr;
    const int kept = nms_process(smem, L, pb, sink);
    if (threadIdx.x == 0) *count_out = kept;
}

// ---------------------------------------------------------------------------------------
// packing: [B,C,max_out,5] slots -> packed rows + offsets
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
pack_offsets_kernel(const int* __restrict__ counts, int nbc, int* __restrict__ offsets) {
    // single CTA exclusive scan (nbc is a few thousand)
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nbc; base += 1024) {
        int i = base + tid;
        int v = i < nbc ? counts[i] : 0;
        int x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(kFullMask, x, d); if (lane >= d) x += o; }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int s = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(kFullMask, s, d); if (lane >= d) s += o; }
            warp_sums[lane] = s;
        }
        __syncthreads();
        int prefix = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + x - v;
        if (i < nbc) offsets[i] = prefix;
        __syncthreads();
        if (tid == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (tid == 0) offsets[nbc] = carry;
}

__global__ void pack_rows_kernel(const int* __restrict__ counts, const int* __restrict__ offsets,
                                 const float* __restrict__ dets, int max_out, float* __restrict__ packed,
                                 int capacity) {
    const int bc = blockIdx.x;
    const int n = counts[bc];
    const int off = offsets[bc];
    const float* src = dets + (size_t)bc * max_out * 5;
    for (int t = threadIdx.x; t < n * 5; t += blockDim.x) {
        int row = off + t / 5;
        if (row < capacity) packed[(size_t)off * 5 + t] = src[t];
    }
}


// ---------------------------------------------------------------------------------------
// multi-GPU exchange: packing fused with the gather.  Every rank owns one slot in every
// peer's exchange buffer ([header 64 x i32 | counts nbc x i32, padded to 256 B | rows capacity x 5 x f32]);
// the pack kernel stores this rank's counts and packed rows straight into its slot on EVERY peer (P2P
// stores over NVLink / NVSwitch, peer order rotated by rank so the links are used evenly) -- no staging copy
// and no collective call; a barrier between the ranks afterwards is all that is needed.
// ---------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerSlots { unsigned char* p[kMaxPeers]; };
constexpr int kPackThreads = 256;
constexpr int kPackPer = kPackThreads / 32;       // (image, class) slots per CTA: one warp each

// phase 1: pack this rank's rows into ITS OWN slot of its own exchange buffer (local HBM), write counts + header.
// The exclusive prefix of the counts is computed by the CTA itself (a few thousand ints from L2, once per 8 slots):
// no separate scan kernel in front of the copy.
__global__ void __launch_bounds__(kPackThreads)
pack_local_kernel(const int* __restrict__ counts, const float* __restrict__ dets, int max_out, int nbc, int B, int C,
                  unsigned char* slot, int capacity, size_t rows_off, int* __restrict__ offsets_out) {
    __shared__ int s_red[kPackPer];
    __shared__ int s_cnt[kPackPer];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bc0 = blockIdx.x * kPackPer;
    int acc = 0;
    for (int i = threadIdx.x; i < bc0; i += kPackThreads) acc += __ldg(counts + i);
    acc = __reduce_add_sync(kFullMask, acc);
    const int bc = bc0 + warp;
    const int n = bc < nbc ? __ldg(counts + bc) : 0;
    if (lane == 0) { s_red[warp] = acc; s_cnt[warp] = n; }
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < kPackPer; ++w) off += s_red[w] + (w < warp ? s_cnt[w] : 0);
    if (bc < nbc) {
        if (offsets_out && lane == 0) offsets_out[bc] = off;
        const float* src = dets + (size_t)bc * max_out * 5;
        int len = n * 5;
        if (off + n > capacity) len = max(0, capacity - off) * 5;
        float* dst = reinterpret_cast<float*>(slot + rows_off) + (size_t)off * 5;
        for (int t = lane; t < len; t += 32) dst[t] = src[t];
    }
    // counts of this CTA's slots; the CTA that owns the last slot also writes the header (total = its running sum)
    int* cdst = reinterpret_cast<int*>(slot + 256);
    if (bc < nbc && lane == 0) cdst[bc] = n;
    if (bc == nbc - 1 && lane == 0) {
        const int total = off + n;
        int* hdr = reinterpret_cast<int*>(slot);
        hdr[0] = min(total, capacity); hdr[1] = B; hdr[2] = C; hdr[3] = total;
        if (offsets_out) offsets_out[nbc] = total;
    }
}

// ---------------------------------------------------------------------------------------
// result wire format (data/sarship_coco.py:293-336): the rows of the slot layout (or of packed / gathered rows)
// as COCO result records, in the reference's order -- classes ascending (background and unmapped classes skipped),
// images ascending inside a class, rows score-descending -- with bbox = [x, y, x2 - x + 1, y2 - y + 1] computed in
// float64 exactly as `dets.astype(np.float)` does (:296-303).  One warp per (class, image); the exclusive prefix of
// the counts in that (class-major) order is computed by the CTA itself.
//   out_ids  [capacity, 2] int32   (image index b, class c)      -- the host maps them to image_id / category_id
//   out_vals [capacity, 5] float64 (x, y, w, h, score)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPackThreads)
coco_records_kernel(const int* __restrict__ counts, const float* __restrict__ dets, int B, int C, int max_out,
                    const int* __restrict__ src_offsets, const int* __restrict__ class_to_cat, int* __restrict__ out_ids,
                    double* __restrict__ out_vals, int capacity, int* __restrict__ out_total) {
    __shared__ int s_red[kPackPer];
    __shared__ int s_cnt[kPackPer];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nbc = B * C;
    const int t0 = blockIdx.x * kPackPer;                 // slots in class-major order: t = c * B + b
    auto count_of = [&](int t) -> int {
        const int c = t / B, b = t - c * B;
        return (c == 0 || (class_to_cat && __ldg(class_to_cat + c) < 0)) ? 0 : __ldg(counts + b * C + c);
    };
    int acc = 0;
    for (int t = threadIdx.x; t < t0; t += kPackThreads) acc += count_of(t);
    acc = __reduce_add_sync(kFullMask, acc);
    const int t = t0 + warp;
    const int n = t < nbc ? count_of(t) : 0;
    if (lane == 0) { s_red[warp] = acc; s_cnt[warp] = n; }
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < kPackPer; ++w) off += s_red[w] + (w < warp ? s_cnt[w] : 0);
    if (t == nbc - 1 && lane == 0) *out_total = off + n;
    if (t >= nbc || n == 0) return;
    const int c = t / B, b = t - c * B;
    const int bc = b * C + c;
    const float* src = max_out > 0 ? dets + (size_t)bc * max_out * 5 : dets + (size_t)__ldg(src_offsets + bc) * 5;
    for (int k = lane; k < n; k += 32) {
        const int o = off + k;
        if (o >= capacity) break;
        const float* r = src + (size_t)k * 5;
        const double x = (double)r[0], y = (double)r[1];
        double* v = out_vals + (size_t)o * 5;
        v[0] = x; v[1] = y;
        v[2] = (double)r[2] - x + 1.0;                    // :302
        v[3] = (double)r[3] - y + 1.0;
        v[4] = (double)r[4];
        out_ids[2 * o] = b;
        out_ids[2 * o + 1] = c;
    }
}

__device__ __forceinline__ void st_multimem_v4(void* mc_addr, uint4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}

// phase 2: the local slot [header | counts | rows stored] -> the same slot of every peer's buffer, as a flat stream
// of 16-byte words (the slot is 256-byte aligned, its regions 256-byte padded: whole words, no tails).
//   grid = (chunks, targets).  targets = world - 1 unicast destinations (peer order rotated by rank so that the
//   ranks do not all write to the same peer at once), or ONE multicast destination (kMulticast: the address of the
//   slot in the multicast mapping of the symmetric buffer -- a single store is replicated to every rank by the
//   NVSwitch, so the rows leave this GPU once instead of world - 1 times).
// kUnroll independent 16-byte loads in flight per thread before the stores: NVLink stores are fire-and-forget, the
// kernel is bound by how fast the SMs can issue them.
constexpr int kCopyThreads = 256;
constexpr int kCopyUnroll = 8;
template <bool kMulticast>
__global__ void __launch_bounds__(kCopyThreads)
slot_copy_kernel(const unsigned char* __restrict__ local_slot, PeerSlots peers, unsigned char* mc_slot, int world, int rank,
                 size_t rows_off) {
    const int* hdr = reinterpret_cast<const int*>(local_slot);
    const int stored = hdr[0];
    const size_t head_words = rows_off / 16;                                  // header + counts region
    const size_t words = head_words + ((size_t)stored * 20 + 15) / 16;
    const uint4* src = reinterpret_cast<const uint4*>(local_slot);
    uint4* dst;
    if (kMulticast) dst = reinterpret_cast<uint4*>(mc_slot);
    else dst = reinterpret_cast<uint4*>(peers.p[(rank + 1 + blockIdx.y) % world]);
    const size_t stride = (size_t)gridDim.x * kCopyThreads;
    for (size_t w0 = (size_t)blockIdx.x * kCopyThreads + threadIdx.x; w0 < words; w0 += stride * kCopyUnroll) {
        uint4 v[kCopyUnroll];
#pragma unroll
        for (int u = 0; u < kCopyUnroll; ++u) {
            const size_t w = w0 + u * stride;
            if (w < words) v[u] = __ldg(src + w);
        }
#pragma unroll
        for (int u = 0; u < kCopyUnroll; ++u) {
            const size_t w = w0 + u * stride;
            if (w < words) {
                if (kMulticast) st_multimem_v4(dst + w, v[u]);
                else dst[w] = v[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// The exchange with its synchronisation folded in (rd_exchange_*): no barrier kernels, ONE cross-GPU rendezvous per
// round.  Every rank's symmetric buffer is [ExchangeCtrl | slots of parity 0 | slots of parity 1]; round e (the
// epoch, kept on the device so that the launch pair can be replayed from a CUDA graph) uses the slots of parity
// e & 1.  pack: rows -> this rank's own slot (local HBM).  copy: the slot's used prefix -> the same slot of every
// peer (P2P or multicast stores); the CTA that finishes last publishes the epoch in flags[rank] of every rank and
// waits until every flag of its OWN control block has reached the epoch: when the kernel completes, the rows of
// every rank have landed here.  The second parity makes a "buffer free" rendezvous unnecessary: a rank can only
// start writing round e + 2 after round e + 1 has completed everywhere, i.e. after every peer has, in ITS stream,
// passed the reads of round e that it enqueued before its round e + 1.
// ---------------------------------------------------------------------------------------
struct ExchangeCtrl {
    unsigned int flags[kMaxPeers];     // flags[q] = latest epoch published by rank q (written by rank q)
    unsigned int pad0[64 - kMaxPeers];
    unsigned int epoch;                // rounds completed by this rank (local)
    unsigned int done;                 // CTAs of the running copy kernel that have finished their stores (local)
    unsigned int error;                // 1 = a wait timed out
    unsigned int pad1[256 - 64 - 3];
};
static_assert(sizeof(ExchangeCtrl) == 1024, "control block is 1 KB");
struct PeerBases { unsigned char* p[kMaxPeers]; };

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kPackThreads)
exchange_pack_kernel(const int* __restrict__ counts, const float* __restrict__ dets, int max_out, int nbc, int B, int C,
                     unsigned char* local_base, int world, int rank, size_t slot_bytes, int capacity, size_t rows_off) {
    __shared__ int s_red[kPackPer];
    __shared__ int s_cnt[kPackPer];
    const ExchangeCtrl* ctrl = reinterpret_cast<const ExchangeCtrl*>(local_base);
    grid_dependency_wait();                                  // the producer of counts / dets
    const unsigned int e = *reinterpret_cast<const volatile unsigned int*>(&ctrl->epoch) + 1u;
    unsigned char* slot = local_base + sizeof(ExchangeCtrl) + ((size_t)(e & 1u) * world + rank) * slot_bytes;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bc0 = blockIdx.x * kPackPer;
    int acc = 0;
    for (int i = threadIdx.x; i < bc0; i += kPackThreads) acc += __ldg(counts + i);
    acc = __reduce_add_sync(kFullMask, acc);
    const int bc = bc0 + warp;
    const int n = bc < nbc ? __ldg(counts + bc) : 0;
    if (lane == 0) { s_red[warp] = acc; s_cnt[warp] = n; }
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int w = 0; w < kPackPer; ++w) off += s_red[w] + (w < warp ? s_cnt[w] : 0);
    if (bc < nbc) {
        const float* src = dets + (size_t)bc * max_out * 5;
        int len = n * 5;
        if (off + n > capacity) len = max(0, capacity - off) * 5;
        float* dst = reinterpret_cast<float*>(slot + rows_off) + (size_t)off * 5;
        for (int t = lane; t < len; t += 32) dst[t] = src[t];
        if (lane == 0) reinterpret_cast<int*>(slot + 256)[bc] = n;
    }
    if (bc == nbc - 1 && lane == 0) {
        const int total = off + n;
        int* hdr = reinterpret_cast<int*>(slot);
        hdr[0] = min(total, capacity); hdr[1] = B; hdr[2] = C; hdr[3] = total; hdr[4] = (int)e;
    }
}

template <bool kMulticast>
__global__ void __launch_bounds__(kCopyThreads)
exchange_copy_kernel(PeerBases bases, unsigned char* mc_base, int world, int rank, size_t slot_bytes, size_t rows_off,
                     unsigned long long timeout_ns) {
    __shared__ int s_last;
    ExchangeCtrl* ctrl = reinterpret_cast<ExchangeCtrl*>(bases.p[rank]);
    grid_dependency_wait();                                  // exchange_pack_kernel
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&ctrl->epoch) + 1u;
    const size_t slot_off = sizeof(ExchangeCtrl) + ((size_t)(e & 1u) * world + rank) * slot_bytes;
    const unsigned char* local_slot = bases.p[rank] + slot_off;
    const int stored = reinterpret_cast<const int*>(local_slot)[0];
    const size_t words = rows_off / 16 + ((size_t)stored * 20 + 15) / 16;      // header + counts + rows, 16-byte words
    const uint4* src = reinterpret_cast<const uint4*>(local_slot);
    if (world > 1) {
        uint4* dst = kMulticast ? reinterpret_cast<uint4*>(mc_base + slot_off)
                                : reinterpret_cast<uint4*>(bases.p[(rank + 1 + blockIdx.y) % world] + slot_off);
        const size_t stride = (size_t)gridDim.x * kCopyThreads;
        for (size_t w0 = (size_t)blockIdx.x * kCopyThreads + threadIdx.x; w0 < words; w0 += stride * kCopyUnroll) {
            uint4 v[kCopyUnroll];
#pragma unroll
            for (int u = 0; u < kCopyUnroll; ++u) {
                const size_t w = w0 + u * stride;
                if (w < words) v[u] = __ldg(src + w);
            }
#pragma unroll
            for (int u = 0; u < kCopyUnroll; ++u) {
                const size_t w = w0 + u * stride;
                if (w < words) {
                    if (kMulticast) st_multimem_v4(dst + w, v[u]);
                    else dst[w] = v[u];
                }
            }
        }
    }
    // ---- completion: the last CTA publishes the epoch everywhere, then waits for everybody's ------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                              // this CTA's stores, system-wide, before its ticket
        const unsigned int total = gridDim.x * gridDim.y;
        s_last = atomicAdd(&ctrl->done, 1u) == total - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        ctrl->done = 0u;
        ctrl->epoch = e;
        __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < world) {                          // thread q: publish on rank q, then wait for rank q here
        ExchangeCtrl* peer = reinterpret_cast<ExchangeCtrl*>(bases.p[threadIdx.x]);
        *reinterpret_cast<volatile unsigned int*>(&peer->flags[rank]) = e;
        const volatile unsigned int* mine = &ctrl->flags[threadIdx.x];
        const unsigned long long t0 = global_timer_ns();
        while ((int)(*mine - e) < 0) {
            if (global_timer_ns() - t0 > timeout_ns) { ctrl->error = 1u; break; }
            __nanosleep(100);
        }
        __threadfence_system();                              // the peers' rows are visible to whatever follows in the stream
    }
}


// ---- the copy phase on the TMA engines (unicast).  Stores to a peer's memory issued by the SMs' load/store units
// hold up the whole SM's memory pipeline while NVLink drains them (measured at 2 GPUs: with the LSU copy on 24 - 48 SMs
// the exchange ADDED its 12 - 16 us to the 26 us stage of the other lanes instead of hiding behind them), so the rows
// never pass through registers here: one thread per CTA moves 16 KB chunks local slot -> shared memory
// (cp.async.bulk, mbarrier) -> the peer's slot (cp.async.bulk.global.shared::cta), a ring of kXStages chunks in
// flight.  grid = (CTAs per peer, world - 1), 32 threads; the completion protocol is that of exchange_copy_kernel.
constexpr int kXChunk = 16 * 1024;
constexpr int kXStages = 4;

__device__ __forceinline__ void x_mbar_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void x_load(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void x_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XDONE_%=;\n"
        "bra XWAIT_%=;\n"
        "XDONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void x_store(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(32)
exchange_copy_tma_kernel(PeerBases bases, int world, int rank, size_t slot_bytes, size_t rows_off, unsigned long long timeout_ns) {
    extern __shared__ __align__(128) unsigned char x_smem[];          // kXStages chunks
    __shared__ __align__(8) unsigned long long s_bar[kXStages];
    __shared__ int s_last;
    ExchangeCtrl* ctrl = reinterpret_cast<ExchangeCtrl*>(bases.p[rank]);
    grid_dependency_wait();                                  // exchange_pack_kernel
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&ctrl->epoch) + 1u;
    const size_t slot_off = sizeof(ExchangeCtrl) + ((size_t)(e & 1u) * world + rank) * slot_bytes;
    const unsigned char* src = bases.p[rank] + slot_off;
    if (threadIdx.x == 0 && world > 1) {
        const int stored = reinterpret_cast<const int*>(src)[0];
        const size_t bytes = rows_off + (((size_t)stored * 20 + 15) & ~(size_t)15);    // header + counts + rows
        unsigned char* dst = bases.p[(rank + 1 + blockIdx.y) % world] + slot_off;
        const size_t nchunks = (bytes + kXChunk - 1) / kXChunk;
#pragma unroll
        for (int st = 0; st < kXStages; ++st) x_mbar_init(&s_bar[st]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        auto chunk_bytes = [&](size_t c) { return (unsigned)min((size_t)kXChunk, bytes - c * kXChunk); };
        // prologue: the first kXStages - 1 chunks of this CTA are requested
        size_t c_load = blockIdx.x;
        int n_load = 0;
        for (; n_load < kXStages - 1 && c_load < nchunks; ++n_load, c_load += gridDim.x)
            x_load(x_smem + (size_t)(n_load % kXStages) * kXChunk, src + c_load * kXChunk, chunk_bytes(c_load), &s_bar[n_load % kXStages]);
        int it = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int st = it % kXStages;
            x_wait(&s_bar[st], (unsigned)((it / kXStages) & 1));
            x_store(dst + c * kXChunk, x_smem + (size_t)st * kXChunk, chunk_bytes(c));
            // refill: load number n_load goes to the stage chunk it - 1 was stored from; that store (every bulk group
            // but the one just committed) must have finished reading shared memory
            if (c_load < nchunks) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                x_load(x_smem + (size_t)(n_load % kXStages) * kXChunk, src + c_load * kXChunk, chunk_bytes(c_load),
                       &s_bar[n_load % kXStages]);
                ++n_load;
                c_load += gridDim.x;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // every store of this CTA has been performed
    }
    // ---- completion: the last CTA publishes the epoch everywhere, then waits for everybody's ------------------
    __syncwarp();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y;
        s_last = atomicAdd(&ctrl->done, 1u) == total - 1u;
    }
    __syncwarp();
    if (!s_last) return;
    if (threadIdx.x == 0) {
        ctrl->done = 0u;
        ctrl->epoch = e;
        __threadfence_system();
    }
    __syncwarp();
    if ((int)threadIdx.x < world) {
        ExchangeCtrl* peer = reinterpret_cast<ExchangeCtrl*>(bases.p[threadIdx.x]);
        *reinterpret_cast<volatile unsigned int*>(&peer->flags[rank]) = e;
        const volatile unsigned int* mine = &ctrl->flags[threadIdx.x];
        const unsigned long long t0 = global_timer_ns();
        while ((int)(*mine - e) < 0) {
            if (global_timer_ns() - t0 > timeout_ns) { ctrl->error = 1u; break; }
            __nanosleep(100);
        }
        __threadfence_system();
    }
}

}  // namespace rd

using namespace rd;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

unsigned long long rd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }


int rd_detect_forward(const float* arm_loc, const float* arm_conf, const float* odm_loc, float* odm_conf,
                      const float* priors, int B, int P, int C, float objectness_thre, float v0, float v1,
                      float* boxes_out, float* scores_out, void* stream) {
    NvtxRange nvtx_range("rd_detect_forward");
    if (!arm_loc || !arm_conf || !odm_loc || !odm_conf || !priors || !boxes_out || !scores_out) return RD_ERR_BAD_ARG;
    if (B <= 0 || P <= 0 || C <= 0) return RD_ERR_BAD_ARG;
    if (C > kMaxClasses) return RD_ERR_UNSUPPORTED;
    if ((((uintptr_t)arm_loc | (uintptr_t)odm_loc | (uintptr_t)priors | (uintptr_t)boxes_out) & 15) ||
        ((uintptr_t)arm_conf & 7))
        return RD_ERR_ALIGNMENT;
    const long long total = (long long)B * P;
    const long long warps = (total + 31) / 32;
    const int blocks = (int)((warps * 32 + kCollectThreads - 1) / kCollectThreads);
    detect_forward_kernel<<<blocks, kCollectThreads, 0, (cudaStream_t)stream>>>(
        (const float4*)arm_loc, (const float2*)arm_conf, (const float4*)odm_loc, odm_conf, (const float4*)priors,
        total, P, C, objectness_thre, v0, v1, (float4*)boxes_out, scores_out);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_arm_zero_rows(const float* arm_conf, float* odm_conf, long long rows, int C, float objectness_thre, void* stream) {
    if (!arm_conf || !odm_conf || rows <= 0 || C <= 0) return RD_ERR_BAD_ARG;
    if ((uintptr_t)arm_conf & 7) return RD_ERR_ALIGNMENT;
    const long long blocks = (rows + kCollectThreads - 1) / kCollectThreads;
    zero_filtered_rows_kernel<<<(unsigned)blocks, kCollectThreads, 0, (cudaStream_t)stream>>>(
        (const float2*)arm_conf, odm_conf, rows, C, objectness_thre);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

size_t rd_detect_workspace_bytes(int B, int P, int C) {
    if (B <= 0 || P <= 0 || C <= 0) return 0;
    return carve_ws(nullptr, B, P, C).total;
}

int rd_detect_workspace_reset(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace) return RD_ERR_BAD_ARG;
    // the control block (header, nnodes, gtab) must start zero; its size depends on B, so clear everything
    cudaError_t e = cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream);
    return (int)e;
}

}  // extern "C"

// Shared-memory carve-out of the stage's kernels, set once per device instead of left to the driver's per-kernel
// choice (measured on B200, cfg 3 sparse; `RD_CARVEOUT="collect,graph,small,large"` in percent overrides it for
// experiments).  The carve-out is a state of the whole SM, so the choice of one kernel is also the L1 size of every CTA
// of the other batches in flight that shares the SM with it.  With the driver's choice -- 228 KB for graph_kernel, 28 KB
// of L1 left -- four batches in flight took 28.3 us per batch; with 196 KB for graph_kernel and 164 KB for
// collect_kernel / nms_small_kernel<256,128>: 27.1 us, and 25.8 us once graph_kernel's footprint had shrunk to 46.6 KB
// per CTA so that its four CTAs per SM fit the 196 KB (with 50.7 KB only three did, which cost one batch alone 4 us).
static void set_stage_carveouts() {
    static bool s_done[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    bool& done = s_done[(unsigned)dev % kMaxDevices];
    if (done) return;
    done = true;
    int pc = 72, pg = 85, ps = 72, pl = -1;
    if (const char* env = getenv("RD_CARVEOUT")) {
        int a = -1, b = -1, c = -1, d = -1;
        const int nf = sscanf(env, "%d,%d,%d,%d", &a, &b, &c, &d);
        if (nf >= 1) { pc = a; pg = (nf >= 2 ? b : a); ps = (nf >= 3 ? c : a); pl = (nf >= 4 ? d : a); }
    }
    const auto attr = cudaFuncAttributePreferredSharedMemoryCarveout;
    if (pc >= 0) {
        cudaFuncSetAttribute(collect_kernel<false, false>, attr, pc);
        cudaFuncSetAttribute(collect_kernel<false, true>, attr, pc);
        cudaFuncSetAttribute(collect_kernel<true, false>, attr, pc);
        cudaFuncSetAttribute(collect_kernel<true, true>, attr, pc);
    }
    if (pg >= 0) cudaFuncSetAttribute(graph_kernel, attr, pg);
    if (ps >= 0) cudaFuncSetAttribute(nms_small_kernel<kSmallCap, kSmallThreads, RD_SMALL_MINBLOCKS>, attr, ps);
    if (pl >= 0) cudaFuncSetAttribute(nms_large_kernel, attr, pl);
    (void)cudaGetLastError();
}

static int detect_fused_impl(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                             const float* priors, int B, int P, int C, float objectness_thre, float conf_thresh,
                             float nms_thresh, int top_k, int max_out, const float* img_scale, int nms_flags,
                             int row_layout, float v0, float v1, void* workspace, size_t workspace_bytes,
                             int* out_counts, float* out_dets, int* out_anchor, void* stream, cudaEvent_t* ev) {
    NvtxRange nvtx_range("rd_detect_fused");
    if (!arm_loc || !arm_conf || !odm_loc || !odm_conf || !priors || !workspace || !out_counts || !out_dets)
        return RD_ERR_BAD_ARG;
    if (B <= 0 || P <= 0 || C <= 0 || top_k <= 0 || max_out <= 0) return RD_ERR_BAD_ARG;
    if (C > kMaxClasses || B > 65535 || (long long)B * C > (1 << 30)) return RD_ERR_UNSUPPORTED;
    if (P > 65535 * kSliceAnchors / 64) return RD_ERR_UNSUPPORTED;
    if (top_k > RD_MAX_NMS_BOXES) return RD_ERR_UNSUPPORTED;
    if ((((uintptr_t)arm_loc | (uintptr_t)odm_loc | (uintptr_t)priors | (uintptr_t)workspace) & 15) ||
        ((uintptr_t)arm_conf & 7) || (img_scale && ((uintptr_t)img_scale & 15)))
        return RD_ERR_ALIGNMENT;
    DetectWs ws = carve_ws(workspace, B, P, C);
    if (workspace_bytes < ws.total) return RD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    set_stage_carveouts();

    GraphOut GO;
    GO.nnodes = ws.nnodes; GO.gtab = ws.gtab; GO.nbox = ws.nbox; GO.nanc = ws.nanc; GO.ncr = ws.ncr;
    GO.adjn = ws.adjn; GO.img_flag = ws.flag; GO.img_scale = img_scale; GO.thr = nms_thresh; GO.flags = nms_flags;
    if (ev) cudaEventRecord(ev[0], st);
    ArmGate gate;
    gate.thre = objectness_thre;
    gate.admit_all = conf_thresh < 0.f ? 1 : 0;
    gate.gap_lo = INFINITY; gate.gap_hi = -INFINITY;                  // always evaluate the formula ...
    if (objectness_thre > 0.f && objectness_thre < 1.f) {             // ... unless logit(thre) exists
        const double lg = std::log((double)objectness_thre / (1.0 - (double)objectness_thre));
        const double margin = 1e-3 * (1.0 + std::fabs(lg));           // >> the fp32 error of the formula (~1e-6)
        gate.gap_lo = (float)(lg - margin);
        gate.gap_hi = (float)(lg + margin);
    }
    // host-mapped (pinned) odm_conf: fetch whole 128-byte lines (see collect_kernel); needs rows that fit four
    // lines from any offset, a line-aligned base and a tensor end that no 16-byte load can straddle
    bool lines = false;
    if (C <= 97 && ((uintptr_t)odm_conf & 127) == 0 && ((size_t)B * P * C) % 32 == 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, odm_conf) == cudaSuccess) lines = pa.type == cudaMemoryTypeHost;
        else (void)cudaGetLastError();
    }
    {
        auto kern = (nms_flags & RD_INPUT_LOGITS) ? (lines ? collect_kernel<true, true> : collect_kernel<true, false>)
                                                  : (lines ? collect_kernel<false, true> : collect_kernel<false, false>);
        kern<<<dim3(ws.S, B), kCollectThreads, (size_t)C * (kTileRows + 1) * sizeof(float), st>>>(
            (const float4*)arm_loc, (const float2*)arm_conf, (const float4*)odm_loc, odm_conf, (const float4*)priors,
            P, C, ws.S, ws.Pn, gate, v0, v1, ws.nsc, ws.header, GO);
    }
    note_launch();
    RD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[1], st);
    {
        static size_t s_graph_smem[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(graph_kernel, sizeof(GraphSmem), s_graph_smem);
        if (e != cudaSuccess) return (int)e;
        e = launch_pdl(graph_kernel, dim3(kGraphSplit, B), dim3(kGraphThreads), sizeof(GraphSmem), st,
                       (const int*)ws.nnodes, (const uint32_t*)ws.gtab, (const float4*)ws.nbox,
                       (const uint32_t*)ws.ncr, P, nms_thresh, nms_flags, ws.adj, ws.adj2, ws.adjn, ws.flag);
        if (e != cudaSuccess) return (int)e;
        note_launch();
        RD_CHECK_LAUNCH();
    }

    FusedNmsArgs A;
    A.nsc = ws.nsc; A.nbox = ws.nbox; A.nanc = ws.nanc; A.nnodes = ws.nnodes; A.gtab = ws.gtab;
    A.img_flag = ws.flag; A.adj = ws.adj; A.adj2 = ws.adj2; A.adjn = ws.adjn; A.queue = ws.queue; A.header = ws.header;
    A.cand = ws.cand;
    A.nbc = B * C; A.C = C; A.P = P; A.Pn = ws.Pn;
    A.conf_thresh = conf_thresh;
    A.thr = nms_thresh; A.top_k = top_k; A.max_out = max_out; A.flags = nms_flags; A.row_layout = row_layout;
    A.out_counts = out_counts; A.out_dets = out_dets; A.out_anchor = out_anchor;

    {
        const int mcap = top_k < P ? top_k : P;
        const NmsSmemLayout Ll = nms_layout(mcap);
        static size_t s_large_smem[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(nms_large_kernel, Ll.total, s_large_smem);
        if (e != cudaSuccess) return (int)e;
        static int s_sms[kMaxDevices];
        int dev = 0;
        cudaGetDevice(&dev);
        int& s_dev_sms = s_sms[(unsigned)dev % kMaxDevices];
        if (s_dev_sms == 0) {
            cudaDeviceGetAttribute(&s_dev_sms, cudaDevAttrMultiProcessorCount, dev);
            if (s_dev_sms <= 0) s_dev_sms = 148;
        }
        int per_sm = (int)((220 * 1024) / (Ll.total + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > RD_LARGE_PER_SM) per_sm = RD_LARGE_PER_SM;
        A.large_grid = s_dev_sms * per_sm;
        A.large_mcap = mcap;
        A.large_smem = (int)Ll.total;
    }
    if (ev) cudaEventRecord(ev[2], st);
    {   // programmatic dependent launch: the scan + sort of nms_small_kernel run beside graph_kernel
        // few problems (the grid cannot fill the GPU: footprint is free): the <1024, 256> instance, else <256, 128>.
        // (A separate <512, 256> kernel for the images with 1025 .. 4096 nodes was measured: 0.086 against 0.104 ms at
        // 1185 nodes per image, but its empty launch costs the sparse workload 1.7 us per batch with four batches in
        // flight and 3 us alone; those images are resolved by nms_large_kernel's mid path instead.)
        bool wide = (long long)B * (C - 1) <= 2 * 148;
        const int forced = nms_flags & RD_DEBUG_INSTANCE_MASK;     // test-only override (parity tests of every instance)
        if (forced == RD_DEBUG_INSTANCE_256) wide = false;
        else if (forced == RD_DEBUG_INSTANCE_1024) wide = true;
        cudaError_t e = wide ? launch_pdl(nms_small_kernel<kWideCap, kWideThreads, 2>, dim3(C, B), dim3(kWideThreads), 0, st, A)
                             : launch_pdl(nms_small_kernel<kSmallCap, kSmallThreads, RD_SMALL_MINBLOCKS>, dim3(C, B),
                                          dim3(kSmallThreads), 0, st, A);
        if (e != cudaSuccess) return (int)e;
        note_launch();
    }
    RD_CHECK_LAUNCH();
    {
        cudaError_t e = launch_pdl(nms_large_kernel, dim3(A.large_grid < B * C ? A.large_grid : B * C),
                                   dim3(kLargeThreads), (size_t)A.large_smem, st, A, A.large_mcap);
        if (e != cudaSuccess) return (int)e;
    }
    note_launch();
    RD_CHECK_LAUNCH();
    if (ev) cudaEventRecord(ev[3], st);
    if (ev) cudaEventRecord(ev[4], st);
    return 0;
}

extern "C" {

int rd_detect_fused(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                    const float* priors, int B, int P, int C, float objectness_thre, float conf_thresh,
                    float nms_thresh, int top_k, int max_out, const float* img_scale, int nms_flags,
                    int row_layout, float v0, float v1, void* workspace, size_t workspace_bytes,
                    int* out_counts, float* out_dets, int* out_anchor, void* stream) {
    return detect_fused_impl(arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C, objectness_thre, conf_thresh,
                             nms_thresh, top_k, max_out, img_scale, nms_flags, row_layout, v0, v1, workspace,
                             workspace_bytes, out_counts, out_dets, out_anchor, stream, nullptr);
}

int rd_detect_fused_timed(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                          const float* priors, int B, int P, int C, float objectness_thre, float conf_thresh,
                          float nms_thresh, int top_k, int max_out, const float* img_scale, int nms_flags,
                          int row_layout, float v0, float v1, void* workspace, size_t workspace_bytes,
                          int* out_counts, float* out_dets, int* out_anchor, void* stream, float* stage_ms_host) {
    if (!stage_ms_host) return RD_ERR_BAD_ARG;
    cudaEvent_t ev[5];
    for (int i = 0; i < 5; ++i) {
        cudaError_t e = cudaEventCreate(&ev[i]);
        if (e != cudaSuccess) return (int)e;
    }
    int rc = detect_fused_impl(arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C, objectness_thre, conf_thresh,
                               nms_thresh, top_k, max_out, img_scale, nms_flags, row_layout, v0, v1, workspace,
                               workspace_bytes, out_counts, out_dets, out_anchor, stream, ev);
    if (rc == 0) {
        cudaError_t e = cudaEventSynchronize(ev[4]);
        if (e != cudaSuccess) rc = (int)e;
        for (int i = 0; i < 4 && rc == 0; ++i) {
            e = cudaEventElapsedTime(&stage_ms_host[i], ev[i], ev[i + 1]);
            if (e != cudaSuccess) rc = (int)e;
        }
    }
    for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
    return rc;
}

}  // extern "C"

struct rd_detect_plan {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int launches;          // kernels per replay
};

extern "C" {

int rd_detect_plan_create(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                          const float* priors, int B, int P, int C, float objectness_thre, float conf_thresh,
                          float nms_thresh, int top_k, int max_out, const float* img_scale, int nms_flags,
                          int row_layout, float v0, float v1, void* workspace, size_t workspace_bytes,
                          int* out_counts, float* out_dets, int* out_anchor, rd_detect_plan** plan_out) {
    if (!plan_out) return RD_ERR_BAD_ARG;
    *plan_out = nullptr;
    cudaStream_t cs = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cs); return (int)e; }
    const unsigned long long l0 = g_launches.load(std::memory_order_relaxed);
    int rc = detect_fused_impl(arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C, objectness_thre, conf_thresh,
                               nms_thresh, top_k, max_out, img_scale, nms_flags, row_layout, v0, v1, workspace,
                               workspace_bytes, out_counts, out_dets, out_anchor, cs, nullptr);
    const unsigned long long l1 = g_launches.load(std::memory_order_relaxed);
    g_launches.fetch_sub(l1 - l0, std::memory_order_relaxed);      // captured, not executed
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(cs, &graph);                          // always end the capture, even on error
    cudaStreamDestroy(cs);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return (int)e; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return (int)e; }
    rd_detect_plan* p = new rd_detect_plan;
    p->graph = graph;
    p->exec = exec;
    p->launches = (int)(l1 - l0);
    *plan_out = p;
    return 0;
}

// Generic form: capture whatever the caller enqueues on `stream` between begin and end -- the stage's launch chain
// followed by the exchange (device barriers of the symmetric-memory handle, rd_pack_scatter_ex), say -- into one
// plan.  Thread-local capture mode: other threads' CUDA calls are unaffected.
static thread_local unsigned long long t_capture_l0 = 0;
int rd_plan_capture_begin(void* stream) {
    if (!stream) return RD_ERR_BAD_ARG;                       // the legacy default stream cannot be captured
    t_capture_l0 = g_launches.load(std::memory_order_relaxed);
    return (int)cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal);
}

int rd_plan_capture_end(void* stream, rd_detect_plan** plan_out) {
    if (!stream || !plan_out) return RD_ERR_BAD_ARG;
    *plan_out = nullptr;
    const unsigned long long l1 = g_launches.load(std::memory_order_relaxed);
    g_launches.fetch_sub(l1 - t_capture_l0, std::memory_order_relaxed);      // captured, not executed
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &graph);
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return (int)e; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(graph); return (int)e; }
    rd_detect_plan* p = new rd_detect_plan;
    p->graph = graph;
    p->exec = exec;
    p->launches = (int)(l1 - t_capture_l0);
    *plan_out = p;
    return 0;
}

int rd_detect_plan_launch(rd_detect_plan* plan, void* stream) {
    NvtxRange nvtx_range("rd_detect_plan_launch");
    if (!plan || !plan->exec) return RD_ERR_BAD_ARG;
    cudaError_t e = cudaGraphLaunch(plan->exec, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    note_launch(plan->launches);
    return 0;
}

int rd_detect_plan_destroy(rd_detect_plan* plan) {
    if (!plan) return 0;
    if (plan->exec) cudaGraphExecDestroy(plan->exec);
    if (plan->graph) cudaGraphDestroy(plan->graph);
    delete plan;
    return 0;
}

int rd_pack_detections(const int* counts, const float* dets, int B, int C, int max_out, int* out_offsets,
                       float* packed, int packed_capacity, void* stream) {
    NvtxRange nvtx_range("rd_pack_detections");
    if (!counts || !dets || !out_offsets || !packed || B <= 0 || C <= 0 || max_out <= 0) return RD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    pack_offsets_kernel<<<1, 1024, 0, st>>>(counts, B * C, out_offsets);
    note_launch();
    RD_CHECK_LAUNCH();
    pack_rows_kernel<<<B * C, 128, 0, st>>>(counts, out_offsets, dets, max_out, packed, packed_capacity);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_coco_records(const int* counts, const float* dets, int B, int C, int max_out, const int* src_offsets,
                    const int* class_to_cat, int* out_ids, double* out_vals, int capacity, int* out_total, void* stream) {
    NvtxRange nvtx_range("rd_coco_records");
    if (!counts || !dets || !out_ids || !out_vals || !out_total || B <= 0 || C <= 0 || capacity < 0) return RD_ERR_BAD_ARG;
    if (max_out <= 0 && !src_offsets) return RD_ERR_BAD_ARG;
    if ((uintptr_t)out_vals & 7) return RD_ERR_ALIGNMENT;
    coco_records_kernel<<<(B * C + kPackPer - 1) / kPackPer, kPackThreads, 0, (cudaStream_t)stream>>>(
        counts, dets, B, C, max_out, src_offsets, class_to_cat, out_ids, out_vals, capacity, out_total);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

size_t rd_exchange_slot_bytes(int B, int C, int capacity_rows) {
    if (B <= 0 || C <= 0 || capacity_rows < 0) return 0;
    return 256 + align_up((size_t)B * C * 4, 256) + align_up((size_t)capacity_rows * 20, 256);
}

int rd_pack_scatter_ex(const int* counts, const float* dets, int B, int C, int max_out, int* scratch_offsets,
                       void* const* peer_slots_host, int world, int rank, int slot_B, int capacity_rows,
                       void* multicast_slot, int copy_ctas, void* stream) {
    NvtxRange nvtx_range("rd_pack_scatter");
    if (slot_B < B) return RD_ERR_BAD_ARG;
    if (!counts || !dets || !peer_slots_host || B <= 0 || C <= 0 || max_out <= 0) return RD_ERR_BAD_ARG;
    if (world <= 0 || world > kMaxPeers || rank < 0 || rank >= world || capacity_rows < 0) return RD_ERR_BAD_ARG;
    PeerSlots ps;
    for (int k = 0; k < kMaxPeers; ++k) ps.p[k] = k < world ? static_cast<unsigned char*>(peer_slots_host[k]) : nullptr;
    for (int k = 0; k < world; ++k) if (!ps.p[k] || ((uintptr_t)ps.p[k] & 255)) return RD_ERR_ALIGNMENT;
    if (multicast_slot && ((uintptr_t)multicast_slot & 255)) return RD_ERR_ALIGNMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rows_off = 256 + align_up((size_t)slot_B * C * 4, 256);
    unsigned char* local = ps.p[rank];
    pack_local_kernel<<<(B * C + kPackPer - 1) / kPackPer, kPackThreads, 0, st>>>(counts, dets, max_out, B * C, B, C, local,
                                                                                  capacity_rows, rows_off, scratch_offsets);
    note_launch();
    RD_CHECK_LAUNCH();
    if (world == 1) return 0;
    if (copy_ctas <= 0) copy_ctas = multicast_slot ? 96 : 24;        // CTAs per destination
    if (multicast_slot)
        slot_copy_kernel<true><<<dim3(copy_ctas, 1), kCopyThreads, 0, st>>>(local, ps, (unsigned char*)multicast_slot,
                                                                             world, rank, rows_off);
    else
        slot_copy_kernel<false><<<dim3(copy_ctas, world - 1), kCopyThreads, 0, st>>>(local, ps, nullptr, world, rank,
                                                                                      rows_off);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_pack_scatter(const int* counts, const float* dets, int B, int C, int max_out, int* scratch_offsets,
                    void* const* peer_slots_host, int world, int rank, int slot_B, int capacity_rows, void* stream) {
    return rd_pack_scatter_ex(counts, dets, B, C, max_out, scratch_offsets, peer_slots_host, world, rank, slot_B,
                              capacity_rows, nullptr, 0, stream);
}

size_t rd_exchange_ctrl_bytes(void) { return sizeof(ExchangeCtrl); }

int rd_exchange_round(const int* counts, const float* dets, int B, int C, int max_out, void* const* peer_bases_host,
                      void* multicast_base, int world, int rank, int slot_B, int capacity_rows, int copy_ctas,
                      int timeout_ms, void* stream) {
    NvtxRange nvtx_range("rd_exchange_round");
    if (slot_B < B) return RD_ERR_BAD_ARG;
    if (!counts || !dets || !peer_bases_host || B <= 0 || C <= 0 || max_out <= 0) return RD_ERR_BAD_ARG;
    if (world <= 0 || world > kMaxPeers || rank < 0 || rank >= world || capacity_rows < 0) return RD_ERR_BAD_ARG;
    PeerBases pb;
    for (int k = 0; k < kMaxPeers; ++k) pb.p[k] = k < world ? static_cast<unsigned char*>(peer_bases_host[k]) : nullptr;
    for (int k = 0; k < world; ++k) if (!pb.p[k] || ((uintptr_t)pb.p[k] & 255)) return RD_ERR_ALIGNMENT;
    if (multicast_base && ((uintptr_t)multicast_base & 255)) return RD_ERR_ALIGNMENT;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rows_off = 256 + align_up((size_t)slot_B * C * 4, 256);
    const size_t slot_bytes = rd_exchange_slot_bytes(slot_B, C, capacity_rows);
    cudaError_t e = launch_pdl(exchange_pack_kernel, dim3((B * C + kPackPer - 1) / kPackPer), dim3(kPackThreads), 0, st, counts, dets,
                               max_out, B * C, B, C, pb.p[rank], world, rank, slot_bytes, capacity_rows, rows_off);
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    const unsigned long long timeout_ns = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ull;
    const bool mc = multicast_base != nullptr && world > 1;
    // unicast: the TMA engines move the rows (copy_ctas > 0: that many single-warp CTAs per peer, default 16);
    // copy_ctas < 0 selects the load/store-unit kernel with -copy_ctas CTAs per peer (kept for comparison);
    // multicast: multimem.st is an LSU instruction, |copy_ctas| CTAs (default 96)
    if (mc) {
        const int n = copy_ctas == 0 ? 96 : (copy_ctas < 0 ? -copy_ctas : copy_ctas);
        e = launch_pdl(exchange_copy_kernel<true>, dim3(n, 1), dim3(kCopyThreads), 0, st, pb, (unsigned char*)multicast_base, world,
                       rank, slot_bytes, rows_off, timeout_ns);
    } else if (copy_ctas < 0) {
        e = launch_pdl(exchange_copy_kernel<false>, dim3(-copy_ctas, world == 1 ? 1 : world - 1), dim3(kCopyThreads), 0, st, pb,
                       (unsigned char*)nullptr, world, rank, slot_bytes, rows_off, timeout_ns);
    } else {
        const int n = copy_ctas == 0 ? 16 : copy_ctas;
        static size_t s_x_smem[kMaxDevices];
        e = ensure_dynamic_smem(exchange_copy_tma_kernel, (size_t)kXStages * kXChunk, s_x_smem);
        if (e != cudaSuccess) return (int)e;
        e = launch_pdl(exchange_copy_tma_kernel, dim3(n, world == 1 ? 1 : world - 1), dim3(32), (size_t)kXStages * kXChunk, st, pb,
                       world, rank, slot_bytes, rows_off, timeout_ns);
    }
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

// ---- stand-alone NMS --------------------------------------------------------------------
size_t rd_nms_workspace_bytes(int n) {
    if (n <= 0) return 256;
    return align_up((size_t)n * 8, 256) + align_up((size_t)n * 16, 256);
}

static int launch_single(const unsigned long long* keys, int n, const float4* boxes, float thresh, int top_k,
                         int nms_flags, long long* keep64, int* keep32, int* count_out, cudaStream_t st) {
    int m = top_k < n ? top_k : n;
    if (m > RD_MAX_NMS_BOXES) return RD_ERR_UNSUPPORTED;
    const NmsSmemLayout L = nms_layout(m);
    static size_t s_single_smem[kMaxDevices];
    cudaError_t e = ensure_dynamic_smem(nms_single_kernel, L.total, s_single_smem);
    if (e != cudaSuccess) return (int)e;
    nms_single_kernel<<<1, kLargeThreads, L.total, st>>>(keys, n, boxes, thresh, top_k, m, nms_flags, m, keep64,
                                                          keep32, count_out);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_nms(const float* boxes, const float* scores, int n, float thresh, int top_k, int nms_flags,
           void* workspace, size_t workspace_bytes, long long* keep_out, int* count_out, void* stream) {
    NvtxRange nvtx_range("rd_nms");
    if (!count_out || n < 0 || top_k <= 0) return RD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return (int)cudaMemsetAsync(count_out, 0, sizeof(int), st);
    if (!boxes || !scores || !workspace || !keep_out) return RD_ERR_BAD_ARG;
    if (((uintptr_t)boxes | (uintptr_t)workspace) & 15) return RD_ERR_ALIGNMENT;
    if (workspace_bytes < align_up((size_t)n * 8, 256)) return RD_ERR_WORKSPACE;
    unsigned long long* keys = (unsigned long long*)workspace;
    make_keys_kernel<<<(n + 255) / 256, 256, 0, st>>>(scores, n, keys);
    note_launch();
    RD_CHECK_LAUNCH();
    return launch_single(keys, n, (const float4*)boxes, thresh, top_k, nms_flags, keep_out, nullptr, count_out, st);
}

int rd_nms_host_ex(int* keep_out_host, int* num_out_host, const float* boxes_host, int boxes_num, int boxes_dim,
                   float nms_overlap_thresh, int device_id, int nms_flags) {
    NvtxRange nvtx_range("rd_nms_host");
    if (!keep_out_host || !num_out_host || boxes_num < 0 || boxes_dim < 5) return RD_ERR_BAD_ARG;
    if (boxes_num == 0) { *num_out_host = 0; return 0; }
    if (!boxes_host) return RD_ERR_BAD_ARG;
    if (boxes_num > RD_MAX_NMS_BOXES) return RD_ERR_UNSUPPORTED;
    cudaError_t e = cudaSetDevice(device_id);
    if (e != cudaSuccess) return (int)e;
    const int n = boxes_num;
    const size_t b_dets = align_up((size_t)n * boxes_dim * 4, 256);
    const size_t need = b_dets + 256 + align_up((size_t)n * 4, 256);
    // grow-only pinned + mapped staging buffer per device (the reference malloc'd and freed device scratch per
    // call, nms_kernel.cu:100-108,142-143: driver round trips that cost more than the NMS itself).  The kernel
    // reads the rows and writes the result through the mapping, so there is no copy call at all; the call is
    // synchronous, so the lock is simply held for its duration
    static struct { unsigned char* ptr; size_t bytes; cudaStream_t st; } s_stage[kMaxDevices];
    static std::mutex s_mu;
    std::lock_guard<std::mutex> lock(s_mu);
    auto& sc = s_stage[(unsigned)device_id % kMaxDevices];
    if (sc.bytes < need) {
        if (sc.ptr) cudaFreeHost(sc.ptr);
        sc.ptr = nullptr; sc.bytes = 0;
        const size_t grow = need < (256u << 10) ? (256u << 10) : need;
        e = cudaHostAlloc((void**)&sc.ptr, grow, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) return (int)e;
        sc.bytes = grow;
    }
    float* h_dets = (float*)sc.ptr;
    int* h_cnt = (int*)(sc.ptr + b_dets);
    int* h_keep = (int*)(sc.ptr + b_dets + 256);
    unsigned char* dptr = nullptr;
    e = cudaHostGetDevicePointer((void**)&dptr, sc.ptr, 0);
    if (e != cudaSuccess) return (int)e;
    memcpy(h_dets, boxes_host, (size_t)n * boxes_dim * 4);
    const NmsSmemLayout L = nms_layout(n);
    static size_t s_host_smem[kMaxDevices];
    e = ensure_dynamic_smem(nms_host_kernel, L.total, s_host_smem);
    if (e != cudaSuccess) return (int)e;
    if (!sc.st) {     // private stream: no implicit ordering against the caller's other streams (the data is host data)
        e = cudaStreamCreateWithFlags(&sc.st, cudaStreamNonBlocking);
        if (e != cudaSuccess) return (int)e;
    }
    cudaStream_t st = sc.st;
    nms_host_kernel<<<1, kHostThreads, L.total, st>>>((const float*)dptr, n, boxes_dim, nms_overlap_thresh, nms_flags,
                                                       (int*)(dptr + b_dets), (int*)(dptr + b_dets + 256));
    note_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    const int kept = *h_cnt;
    *num_out_host = kept;
    memcpy(keep_out_host, h_keep, (size_t)kept * 4);
    return 0;
}

int rd_nms_host(int* keep_out_host, int* num_out_host, const float* boxes_host, int boxes_num, int boxes_dim,
                float nms_overlap_thresh, int device_id) {
    return rd_nms_host_ex(keep_out_host, num_out_host, boxes_host, boxes_num, boxes_dim, nms_overlap_thresh,
                          device_id, RD_NMS_PIXEL_PLUS1);
}

}  // extern "C"

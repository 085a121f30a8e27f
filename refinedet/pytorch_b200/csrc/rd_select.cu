// rd_select.cu — stand-alone threshold + top-k select (SURVEY §8b `rd_select_topk`) and the §8b names of the
// entry points that rd_detect.cu implements under longer names.
//
//   select_topk_kernel     one CTA per (image, group of G <= 8 adjacent classes): `scores[b, :, c] > conf_thresh`
//                          (eval_refinedet_coco.py:214, detection_refinedet.py:98), the top_k highest in
//                          score-descending order (eval :222 `argsort()[::-1][:top_k]`, box_utils.py:242-244).
//
// The fused stage never materialises this list (its per-class CTAs select, sort and resolve in one pass);
// the kernel here serves callers that want the candidate lists themselves.  No workspace.
//
// Memory behaviour: `scores` is [B,P,C] row-major, so one class is a column with a stride of C floats — a
// CTA per class with lane = row pulls a whole 32-byte sector per 4-byte score and spends 32 L1 wavefronts per
// warp load (first version: 0.36 ms per batch at config 3, four to five strided passes of a radix select).  A
// CTA therefore owns G adjacent classes and maps lane = (row, class): a warp load covers 32/G rows x 4*G
// contiguous bytes.  ONE scan appends the candidate ANCHORS of every class to its own shared-memory list
// (16-bit entries when P <= 65536; capacity `cap` >= top_k, normally 2*top_k), so 8 classes and the sort buffer
// fit 48 KB and all CTAs of a config-3 batch are resident at once.  Keys (score bits, ~anchor) are then built
// from a gather of the listed scores (L2-resident), several classes at a time when their lists are short,
// sorted with one bitonic network and the first min(n, top_k) entries leave.  A class with more than top_k
// candidates first finds its top_k-th key with an MSB-first 8-bit radix select — over the listed scores staged
// in shared memory when the list held them all, over its column (re-scans) when the list overflowed — so that
// only top_k keys are sorted.
#include "rd_common.cuh"

namespace rd {

constexpr int kSelectThreads = 512;
constexpr int kSelectMaxGroup = 8;

__host__ __device__ inline int select_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Radix select of the top_k-th largest of the keys `key_at(0 .. n-1)` (0 = not a candidate, skipped; more than
// top_k candidates); CTA-wide, returns the key such that exactly top_k candidate keys are >= it (keys are
// unique: the anchor index is part of the key).
template <typename KeyAt>
__device__ __forceinline__ unsigned long long select_threshold_key(int n, KeyAt key_at, int top_k, uint32_t* hist,
                                                                   uint32_t* misc) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long prefix = 0;
    int need = top_k;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kSelectThreads) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kSelectThreads) {
            const unsigned long long k = key_at(i);
            if (k != 0ull && (shift == 56 || (k >> (shift + 8)) == prefix))
                atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns digits [8l, 8l+8): find the digit where the count from the top crosses `need`
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; sum += loc[q]; }
            uint32_t v = sum;         // inclusive suffix sum over lanes
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_down_sync(kFullMask, v, d);
                if (lane + d < 32) v += o;
            }
            uint32_t cum = v - sum;   // keys in the digits owned by higher lanes
            int found = -1;
            uint32_t new_need = 0;
#pragma unroll
            for (int q = 7; q >= 0; --q) {
                if (found < 0 && cum < (uint32_t)need && cum + loc[q] >= (uint32_t)need) {
                    found = lane * 8 + q;
                    new_need = (uint32_t)need - cum;
                }
                cum += loc[q];
            }
            if (found >= 0) {
                misc[1] = (uint32_t)found;
                misc[2] = new_need;
                misc[3] = (hist[found] == new_need) ? 1u : 0u;
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | misc[1];
        need = (int)misc[2];
        const bool done = misc[3] != 0;
        __syncthreads();
        if (done || shift == 0) return prefix << shift;
    }
    return 0ull;
}

// Stages k = 2 .. 32 of the bitonic network below for the 32 keys a warp holds (one per lane, i = index of the
// lane's key in the buffer): 15 compare-exchange steps over shuffles instead of shared memory + CTA barriers.
__device__ __forceinline__ unsigned long long warp_bitonic32(unsigned long long a, int i) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool desc = (i & k) == 0;
            const bool lower = (i & j) == 0;
            const unsigned long long b = __shfl_xor_sync(kFullMask, a, j);
            a = (lower == desc) ? (a > b ? a : b) : (a < b ? a : b);
        }
    }
    return a;
}

// Sorts `nseg` key segments of Kp keys (stride Kp) in `buf`, descending, with one bitonic network.
__device__ __forceinline__ void sort_segments_desc(unsigned long long* buf, int nseg, int Kp) {
    const int half = Kp >> 1, work = nseg * half, hshift = __ffs(half) - 1;      // Kp is a power of two
    int kstart = 2;
    if (Kp >= 64) {                   // runs of 32 in registers (segment offsets are multiples of 64: i & k unchanged)
        const int total = nseg * Kp, lane = threadIdx.x & 31;
        for (int base = (threadIdx.x >> 5) * 32; base < total; base += (kSelectThreads >> 5) * 32)
            buf[base + lane] = warp_bitonic32(buf[base + lane], base + lane);
        __syncthreads();
        kstart = 64;
    }
    for (int k = kstart; k <= Kp; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int w = threadIdx.x; w < work; w += kSelectThreads) {
                const int q = w >> hshift, t = w & (half - 1);
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i + j;
                const bool desc = (i & k) == 0;
                unsigned long long* kg = buf + (size_t)q * Kp;
                const unsigned long long a = kg[i], bb = kg[l];
                if ((a < bb) == desc) { kg[i] = bb; kg[l] = a; }
            }
            __syncthreads();
        }
    }
}

// IdxT: anchor index type of the candidate lists (uint16_t when P <= 65536: 2 bytes per candidate keep
// 8 classes x 2*top_k candidates + the sort buffer within 48 KB, so every CTA of a config-3 batch is resident).
template <typename IdxT>
__global__ void __launch_bounds__(kSelectThreads)
select_topk_kernel(const float* __restrict__ scores, int P, int C, float conf_thresh, int top_k, int first_class,
                   int slot_stride, int G, int cap, int sb, int* __restrict__ idx_out, float* __restrict__ score_out,
                   int* __restrict__ count_out) {      // top_k here is min(top_k, P); slot_stride the caller's top_k
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[4];
    __shared__ uint32_t cnt[kSelectMaxGroup];
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(smem_raw);            // [sb] keys, sb >= cap
    IdxT* lists = reinterpret_cast<IdxT*>(smem_raw + (size_t)sb * 8);                          // [G][cap] anchors
    const int cbase = blockIdx.x * G, b = blockIdx.y;
    const int tid = threadIdx.x;
    const int nact = (C - cbase) < G ? (C - cbase) : G;
    if (tid < kSelectMaxGroup) cnt[tid] = 0;
    __syncthreads();

    // ---- one scan: append the candidate anchors of every class of the group to its list ---------------------
    // lane = (row r, class g): a warp load covers 32/G consecutive rows x the group's G adjacent classes, i.e.
    // 32/G pieces of 4*G contiguous bytes (1-2 sectors each) instead of 32 sectors for 32 scores
    const float* rows = scores + (size_t)b * P * C + cbase;
    {
        const int lane = tid & 31, warp = tid >> 5;
        const int gshift = __ffs(G) - 1;                        // G is a power of two
        const int g = lane & (G - 1), r = lane >> gshift;
        const int rpw = 32 >> gshift;                           // rows per warp load
        const int stride = (kSelectThreads >> 5) * rpw;         // rows per CTA step
        // classes below first_class (background) are never evaluated (eval :213, detection_refinedet.py:97)
        const bool cls_ok = g < nact && cbase + g >= first_class;
        IdxT* list = lists + (size_t)g * cap;
        unsigned cmask = 0;                                     // lanes of this lane's class
        for (int l = g; l < 32; l += G) cmask |= 1u << l;
        const unsigned lt_mask = (1u << lane) - 1u;
        constexpr int kUnroll = 8;                              // independent loads in flight per thread
        const size_t step = (size_t)stride * C;                 // floats between two rows of this lane
        const float* q = rows + g + (size_t)(warp * rpw + r) * C;
        // the loop bound is warp-uniform (first row of the warp's piece): the ballots below need all 32 lanes
        for (int pw = warp * rpw; pw < P; pw += kUnroll * stride, q += kUnroll * step) {
            const int p0 = pw + r;
            float s[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)                   // == thresh: not a candidate
                s[u] = (cls_ok && p0 + u * stride < P) ? __ldg(q + u * step) : conf_thresh;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const bool cand = s[u] > conf_thresh;           // NaN fails the compare, like the reference's mask
                const unsigned bal = __ballot_sync(kFullMask, cand);
                if (bal == 0) continue;                         // warp-uniform
                // the lanes of one class reserve their list slots with ONE shared-memory atomic
                const unsigned peers = bal & cmask;
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (cand && lane == leader) base = atomicAdd(&cnt[g], (uint32_t)__popc(peers));
                base = __shfl_sync(kFullMask, base, leader < 0 ? 0 : leader);
                if (cand) {
                    const uint32_t pos = base + (uint32_t)__popc(peers & lt_mask);
                    if (pos < (uint32_t)cap) list[pos] = (IdxT)(p0 + u * stride);
                }
            }
        }
    }
    __syncthreads();

    // ---- keys, sort, emit.  Three kinds of classes (n = candidates, all decisions uniform over the CTA):
    //   n <= limit        keys built from a gather of the listed scores, several classes packed into the sort
    //                     buffer at a time, sorted, first min(n, top_k) emitted
    //   limit < n <= cap  (only when the buffer has room: cap >= 2 * pow2(top_k), limit = top_k) the listed
    //                     scores are staged in the upper half of the buffer, a shared-memory radix select finds
    //                     the top_k-th key, the top_k keys are compacted into the lower half: a 1024-key sort
    //                     instead of a 2048-key one at top_k = 1000
    //   n > cap           the list overflowed: radix select over the class column (re-scans), alone
    const int kp_top = select_pow2(top_k);
    const bool smem_select = top_k < cap && cap >= 2 * kp_top;
    const int limit = smem_select ? top_k : cap;
    int nmax = 1;
    for (int g = 0; g < nact; ++g) {
        const int n = (int)cnt[g];
        if (n <= limit && n > nmax) nmax = n;
    }
    const int Kp = select_pow2(nmax);                           // segment size of the packed classes (<= cap)
    const int nper = sb / Kp;                                   // segments the sort buffer holds
    int g0 = 0;
    while (g0 < nact) {
        int nseg = 1, kp = kp_top;
        const int n0 = (int)cnt[g0];
        if (n0 > cap) {
            const float* col = rows + g0;
            auto key_at = [&](int p) -> unsigned long long {
                const float s = __ldg(col + (size_t)p * C);
                return s > conf_thresh ? make_key(s, (uint32_t)p) : 0ull;
            };
            const unsigned long long thresh_key = select_threshold_key(P, key_at, top_k, hist, misc);
            for (int i = top_k + tid; i < kp; i += kSelectThreads) sortbuf[i] = 0ull;     // top_k <= kp <= cap here
            if (tid == 0) misc[0] = 0;
            __syncthreads();
            for (int p = tid; p < P; p += kSelectThreads) {
                const unsigned long long k = key_at(p);
                if (k >= thresh_key && k != 0ull) {             // exactly top_k keys are
                    const uint32_t pos = atomicAdd(&misc[0], 1u);
                    if (pos < (uint32_t)top_k) sortbuf[pos] = k;
                }
            }
            __syncthreads();
            if (tid == 0) cnt[g0] = (uint32_t)top_k;
        } else if (n0 > limit) {
            uint32_t* sc32 = reinterpret_cast<uint32_t*>(sortbuf + (cap >> 1));           // upper half: [cap] words
            const IdxT* lst = lists + (size_t)g0 * cap;
            for (int i = tid; i < n0; i += kSelectThreads)
                sc32[i] = float_to_ordered(__ldg(rows + (size_t)lst[i] * C + g0));
            __syncthreads();
            auto key_at = [&](int i) -> unsigned long long {
                return ((unsigned long long)sc32[i] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)lst[i]);
            };
            const unsigned long long thresh_key = select_threshold_key(n0, key_at, top_k, hist, misc);
            for (int i = top_k + tid; i < kp; i += kSelectThreads) sortbuf[i] = 0ull;     // kp <= cap / 2: lower half
            if (tid == 0) misc[0] = 0;
            __syncthreads();
            for (int i = tid; i < n0; i += kSelectThreads) {
                const unsigned long long k = key_at(i);
                if (k >= thresh_key) {
                    const uint32_t pos = atomicAdd(&misc[0], 1u);
                    if (pos < (uint32_t)top_k) sortbuf[pos] = k;
                }
            }
            __syncthreads();
            if (tid == 0) cnt[g0] = (uint32_t)top_k;
        } else {
            while (nseg < nper && g0 + nseg < nact && (int)cnt[g0 + nseg] <= limit) ++nseg;
            kp = Kp;
            const int kshift = __ffs(kp) - 1;
            for (int w = tid; w < nseg * kp; w += kSelectThreads) {
                const int qs = w >> kshift, i = w & (kp - 1);
                unsigned long long k = 0ull;                    // padding: below every real key
                if (i < (int)cnt[g0 + qs]) {
                    const uint32_t p = (uint32_t)lists[(size_t)(g0 + qs) * cap + i];
                    k = make_key(__ldg(rows + (size_t)p * C + g0 + qs), p);
                }
                sortbuf[w] = k;
            }
        }
        __syncthreads();
        sort_segments_desc(sortbuf, nseg, kp);
        const int kshift = __ffs(kp) - 1;
        for (int w = tid; w < nseg * kp; w += kSelectThreads) {
            const int qs = w >> kshift, t = w & (kp - 1);
            const int n = (int)cnt[g0 + qs];
            if (t < (n < top_k ? n : top_k)) {
                const unsigned long long k = sortbuf[w];
                const size_t slot = ((size_t)b * C + cbase + g0 + qs) * (size_t)slot_stride;
                idx_out[slot + t] = (int)key_index(k);
                if (score_out) score_out[slot + t] = key_score(k);
            }
        }
        if (tid < nseg) {
            const int n = (int)cnt[g0 + tid];
            count_out[b * C + cbase + g0 + tid] = n < top_k ? n : top_k;
        }
        __syncthreads();                                        // the sort buffer is reused by the next pack
        g0 += nseg;
    }
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_select_topk(const float* scores, int B, int P, int C, float conf_thresh, int top_k, int first_class,
                   int* idx_out, float* score_out, int* count_out, void* stream) {
    NvtxRange nvtx_range("rd_select_topk");
    if (!scores || !idx_out || !count_out || B <= 0 || P <= 0 || C <= 0 || top_k <= 0 || first_class < 0)
        return RD_ERR_BAD_ARG;
    if (B > 65535) return RD_ERR_UNSUPPORTED;
    const int need = top_k < P ? top_k : P;                       // what a class can yield; the kernel's top_k
    if (need > 4 * RD_MAX_NMS_BOXES) return RD_ERR_UNSUPPORTED;   // 128 KB of keys in shared memory
    const bool narrow = P <= 65536;                               // anchor indices fit 16 bits
    const size_t isz = narrow ? 2 : 4;
    // shared memory per CTA = sb * 8 [sort buffer, sb >= cap keys] + cap * G * isz [lists].  List capacity:
    // twice top_k when 8 classes then still fit the 64 KB budget (48 KB at top_k = 1000: 4 CTAs/SM); never more
    // than the P candidates a class can have.
    int cap = select_pow2(need);
    const int roomy = select_pow2(2 * need < P ? 2 * need : P);
    if (roomy > cap && (size_t)roomy * (8 + kSelectMaxGroup * isz) <= (64u << 10)) cap = roomy;
    int G = kSelectMaxGroup;
    while (G > 1 && (size_t)cap * (8 + G * isz) > (64u << 10)) G >>= 1;
    if (G > C) G = select_pow2(C);
    if (G > kSelectMaxGroup) G = kSelectMaxGroup;
    // a sort buffer of 2*cap keys (two full lists sorted at a time) was measured: 64 KB instead of 48 KB per CTA
    // costs more (cfg 3 sparse 0.092 -> 0.150 ms, dense 0.45 -> 0.54 ms) than the halved number of sort rounds saves
    const int sb = cap;
    const size_t smem = (size_t)sb * 8 + (size_t)cap * G * isz;
    static size_t s_smem16[kMaxDevices], s_smem32[kMaxDevices];
    cudaError_t e = narrow ? ensure_dynamic_smem(select_topk_kernel<uint16_t>, smem, s_smem16)
                           : ensure_dynamic_smem(select_topk_kernel<uint32_t>, smem, s_smem32);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)((C + G - 1) / G), (unsigned)B);
    if (narrow)
        select_topk_kernel<uint16_t><<<grid, kSelectThreads, smem, (cudaStream_t)stream>>>(
            scores, P, C, conf_thresh, need, first_class, top_k, G, cap, sb, idx_out, score_out, count_out);
    else
        select_topk_kernel<uint32_t><<<grid, kSelectThreads, smem, (cudaStream_t)stream>>>(
            scores, P, C, conf_thresh, need, first_class, top_k, G, cap, sb, idx_out, score_out, count_out);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

// ---- SURVEY §8b names ------------------------------------------------------------------------------------
int rd_decode_filter(const float* arm_loc, const float* arm_conf, const float* odm_loc, float* odm_conf,
                     const float* priors, int B, int P, int C, float objectness_thre, float v0, float v1,
                     float* boxes_out, float* scores_out, void* stream) {
    return rd_detect_forward(arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C, objectness_thre, v0, v1,
                             boxes_out, scores_out, stream);
}

int rd_detect(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
              const float* priors, int B, int P, int C, float objectness_thre, float conf_thresh, float nms_thresh,
              int top_k, int max_out, const float* img_scale, int nms_flags, int row_layout, float v0, float v1,
              void* workspace, size_t workspace_bytes, int* out_counts, float* out_dets, int* out_anchor,
              void* stream) {
    return rd_detect_fused(arm_loc, arm_conf, odm_loc, odm_conf, priors, B, P, C, objectness_thre, conf_thresh,
                           nms_thresh, top_k, max_out, img_scale, nms_flags, row_layout, v0, v1, workspace,
                           workspace_bytes, out_counts, out_dets, out_anchor, stream);
}

size_t rd_workspace_bytes(int B, int P, int C) { return rd_detect_workspace_bytes(B, P, C); }

}  // extern "C"

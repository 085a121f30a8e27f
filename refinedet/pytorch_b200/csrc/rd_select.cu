// rd_select.cu — stand-alone threshold + top-k select (SURVEY §8b `rd_select_topk`) and the §8b names of the
// entry points that rd_detect.cu implements under longer names.
//
//   select_topk_kernel     one CTA per (image, class): `scores[b, :, c] > conf_thresh`
//                          (eval_refinedet_coco.py:214, detection_refinedet.py:98), the top_k highest in
//                          score-descending order (eval :222 `argsort()[::-1][:top_k]`, box_utils.py:242-244).
//
// The fused stage never materialises this list (its per-class CTAs select, sort and resolve in one pass);
// the kernel here serves callers that want the candidate lists themselves.  No workspace: the MSB-first
// 8-bit radix select re-scans the class column (81 CTAs of an image read the same 5.3 MB, L2-resident)
// instead of keeping an n-entry key list, the ≤ top_k selected keys are sorted in shared memory.
#include "rd_common.cuh"

namespace rd {

constexpr int kSelectThreads = 256;

__host__ __device__ inline int select_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__global__ void __launch_bounds__(kSelectThreads)
select_topk_kernel(const float* __restrict__ scores, int P, int C, float conf_thresh, int top_k, int first_class,
                   int* __restrict__ idx_out, float* __restrict__ score_out, int* __restrict__ count_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t misc[4];      // 0 fill counter, 1 digit, 2 need, 3 done
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int c = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t slot = ((size_t)b * C + c) * (size_t)top_k;
    if (c < first_class) {            // background is never evaluated (eval :213, detection_refinedet.py:97)
        if (tid == 0) count_out[b * C + c] = 0;
        return;
    }
    const float* col = scores + (size_t)b * P * C + c;

    unsigned long long prefix = 0, thresh_key = 0;
    int need = top_k, n = 0;
    bool take_all = false;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kSelectThreads) hist[i] = 0;
        __syncthreads();
        for (int p = tid; p < P; p += kSelectThreads) {
            const float s = __ldg(col + (size_t)p * C);
            if (s > conf_thresh) {    // NaN fails the compare, like the reference's mask
                const unsigned long long k = make_key(s, (uint32_t)p);
                if (shift == 56 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
            }
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
            const uint32_t total = __shfl_sync(kFullMask, v, 0);
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
            if (shift == 56 && lane == 0) misc[0] = total;
            if (found >= 0) {
                misc[1] = (uint32_t)found;
                misc[2] = new_need;
                misc[3] = (hist[found] == new_need) ? 1u : 0u;
            }
        }
        __syncthreads();
        if (shift == 56) {
            n = (int)misc[0];
            if (n <= top_k) { take_all = true; __syncthreads(); break; }
        }
        prefix = (prefix << 8) | misc[1];
        need = (int)misc[2];
        const bool done = misc[3] != 0;
        __syncthreads();
        if (done || shift == 0) { thresh_key = prefix << shift; break; }
    }
    const int m = n < top_k ? n : top_k;
    if (m == 0) {
        if (tid == 0) count_out[b * C + c] = 0;
        return;
    }
    if (take_all) thresh_key = 0;
    if (tid == 0) misc[0] = 0;
    __syncthreads();
    for (int p = tid; p < P; p += kSelectThreads) {
        const float s = __ldg(col + (size_t)p * C);
        if (s > conf_thresh) {
            const unsigned long long k = make_key(s, (uint32_t)p);
            if (k >= thresh_key) {
                const uint32_t pos = atomicAdd(&misc[0], 1u);
                if (pos < (uint32_t)m) keys[pos] = k;
            }
        }
    }
    const int Kp = select_pow2(m);
    for (int i = m + tid; i < Kp; i += kSelectThreads) keys[i] = 0ull;      // below every real key
    __syncthreads();
    for (int k = 2; k <= Kp; k <<= 1) {           // bitonic sort, descending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (Kp >> 1); t += kSelectThreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i + j;
                const bool desc = (i & k) == 0;
                const unsigned long long a = keys[i], bb = keys[l];
                if ((a < bb) == desc) { keys[i] = bb; keys[l] = a; }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < m; t += kSelectThreads) {
        const unsigned long long k = keys[t];
        idx_out[slot + t] = (int)key_index(k);
        if (score_out) score_out[slot + t] = key_score(k);
    }
    if (tid == 0) count_out[b * C + c] = m;
}

}  // namespace rd

using namespace rd;

extern "C" {

int rd_select_topk(const float* scores, int B, int P, int C, float conf_thresh, int top_k, int first_class,
                   int* idx_out, float* score_out, int* count_out, void* stream) {
    if (!scores || !idx_out || !count_out || B <= 0 || P <= 0 || C <= 0 || top_k <= 0 || first_class < 0)
        return RD_ERR_BAD_ARG;
    if (B > 65535) return RD_ERR_UNSUPPORTED;
    const int cap = top_k < P ? top_k : P;
    if (cap > 4 * RD_MAX_NMS_BOXES) return RD_ERR_UNSUPPORTED;          // 128 KB of keys in shared memory
    const size_t smem = (size_t)select_pow2(cap) * 8;
    static size_t s_smem[kMaxDevices];
    cudaError_t e = ensure_dynamic_smem(select_topk_kernel, smem, s_smem);
    if (e != cudaSuccess) return (int)e;
    select_topk_kernel<<<dim3((unsigned)C, (unsigned)B), kSelectThreads, smem, (cudaStream_t)stream>>>(
        scores, P, C, conf_thresh, top_k, first_class, idx_out, score_out, count_out);
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

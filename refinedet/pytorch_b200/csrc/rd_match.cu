// rd_match.cu — training-side anchor matching and hard-negative mining on B200 (sm_100a).
//
// Replaces (reference paths): layers/box_utils.py:29-160 (intersect, jaccard, match,
// refine_match, encode) as driven by layers/modules/refinedet_multibox_loss.py:73-86, and
// the double-sort hard-negative ranking of refinedet_multibox_loss.py:117-123.
//
// Kernels
//   match_pass1_kernel   per (image, anchor): anchor box (point_form(prior) or decode(arm_loc,
//                        prior)), running best truth (first index on ties) over the ground
//                        truths staged in shared memory; per-truth best prior through a
//                        CTA-level 64-bit max in shared memory and one global atomicMax per
//                        (CTA, truth).  The [G,P] overlap matrix never exists in memory.
//   match_pass2_kernel   forced matches (last truth wins, box_utils.py:146-150), labels,
//                        threshold, encode -> loc_t, conf_t
//   hnm_select_kernel    one CTA per image row: count positives, MSB-first radix select of the
//                        num_neg-th largest (loss, ~index) key, write the neg mask; the row lives in
//                        registers (16 values per thread) when P <= 16,384
//   elementwise kernels  point_form / center_size / decode / encode / intersect / jaccard
#include <cooperative_groups.h>

#include "rd_common.cuh"

namespace rd {

constexpr int kMatchThreads = 256;
constexpr int kMatchDirect = 96;       // up to this many truths per image: no CTA-level truth list (see match_pass1_kernel)

// key of (overlap, prior): larger overlap wins, then the LOWER prior index (torch.max returns
// the first maximal index)
__device__ __forceinline__ unsigned long long prior_key(float iou, uint32_t p) {
    return ((unsigned long long)float_to_ordered(iou) << 32) | (unsigned long long)(0xffffffffu - p);
}

__device__ __forceinline__ float4 anchor_point_box(const float4* priors, const float4* arm_loc, int b, int p, int P,
                                                   float v0, float v1) {
    float4 pr = __ldg(priors + p);
    if (arm_loc) return decode_box(ldg_stream4(arm_loc + (size_t)b * P + p), pr, v0, v1);   // box_utils.py:135
    return point_form_box(pr);                                                                // box_utils.py:133
}

__global__ void __launch_bounds__(kMatchThreads)
match_pass1_kernel(const float4* __restrict__ truths, const int* __restrict__ gt_count,
                   const float4* __restrict__ priors, const float4* __restrict__ arm_loc, int P, int Gmax,
                   float v0, float v1, unsigned long long* __restrict__ best_prior,   // [B,Gmax]
                   float* __restrict__ bt_overlap, int* __restrict__ bt_idx) {       // [B,P] temporaries
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_truth = reinterpret_cast<float4*>(smem_raw);
    unsigned short* s_list = reinterpret_cast<unsigned short*>(smem_raw + (size_t)Gmax * 16);   // truths to visit
    __shared__ uint32_t s_bb[4];          // bounding box of this CTA's anchor boxes (ordered-uint min/max)
    __shared__ int s_wcnt[kMatchThreads / 32];
    __shared__ int s_nlist;
    const int b = blockIdx.y;
    grid_dependency_wait();                 // padded targets / cleared best_prior of the launches before this one
    const int G = gt_count[b];
    const int p = blockIdx.x * kMatchThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (G <= 0) return;
    for (int g = threadIdx.x; g < G; g += kMatchThreads) {
        s_truth[g] = __ldg(truths + (size_t)b * Gmax + g);
    }
    // Few truths (G <= kMatchDirect): the warp-level cull below tests every truth with one lane each in a round or
    // three; building the CTA-level list first (bounding box of the CTA, ordered compaction: five barriers) costs
    // more than it saves.  The list is then the identity.
    const bool direct = G <= kMatchDirect;                               // CTA-uniform
    if (threadIdx.x < 4) s_bb[threadIdx.x] = (threadIdx.x < 2) ? 0xffffffffu : 0u;   // min x, min y, max x, max y
    if (threadIdx.x == 0) s_nlist = direct ? G : 0;
    if (direct)
        for (int g = threadIdx.x; g < G; g += kMatchThreads) s_list[g] = (unsigned short)g;
    __syncthreads();
    const bool valid = p < P;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) box = anchor_point_box(priors, arm_loc, b, p, P, v0, v1);
    float wx1, wy1, wx2, wy2;
    {
        uint32_t mnx = valid ? float_to_ordered(box.x) : 0xffffffffu, mny = valid ? float_to_ordered(box.y) : 0xffffffffu;
        uint32_t mxx = valid ? float_to_ordered(box.z) : 0u, mxy = valid ? float_to_ordered(box.w) : 0u;
        mnx = __reduce_min_sync(kFullMask, mnx); mny = __reduce_min_sync(kFullMask, mny);
        mxx = __reduce_max_sync(kFullMask, mxx); mxy = __reduce_max_sync(kFullMask, mxy);
        if (!direct && lane == 0) { atomicMin(&s_bb[0], mnx); atomicMin(&s_bb[1], mny); atomicMax(&s_bb[2], mxx); atomicMax(&s_bb[3], mxy); }
        wx1 = ordered_to_float(mnx); wy1 = ordered_to_float(mny);      // bounding box of this WARP's anchor boxes
        wx2 = ordered_to_float(mxx); wy2 = ordered_to_float(mxy);
    }
    // Truths that cannot intersect ANY anchor box of this CTA have IoU exactly 0 with all of them
    // (min(t.x2, a.x2) - max(t.x1, a.x1) <= min(t.x2, bb.x2) - max(t.x1, bb.x1) <= 0, clamped): they can
    // neither become an anchor's best truth (strict >, ascending g, start at (0, g = 0)) nor win a best
    // prior, so they are skipped.  The first CTA of an image visits every truth: it supplies prior 0
    // for truths that overlap nothing (torch.max returns the first index).
    if (!direct) {
        __syncthreads();
        const float bx1 = ordered_to_float(s_bb[0]), by1 = ordered_to_float(s_bb[1]);
        const float bx2 = ordered_to_float(s_bb[2]), by2 = ordered_to_float(s_bb[3]);
        const bool all = blockIdx.x == 0 || !(bx1 <= bx2) || !(by1 <= by2);
        for (int g0 = 0; g0 < G; g0 += kMatchThreads) {        // ordered compaction (ascending g)
            const int g = g0 + threadIdx.x;
            bool take = false;
            if (g < G) {
                const float4 t = s_truth[g];
                const float w = fminf(t.z, bx2) - fmaxf(t.x, bx1);
                const float h = fminf(t.w, by2) - fmaxf(t.y, by1);
                take = all || (w > 0.0f && h > 0.0f) || !(w == w) || !(h == h);
            }
            const unsigned bal = __ballot_sync(kFullMask, take);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            int off = s_nlist + __popc(bal & ((1u << lane) - 1u));
            int tot = 0;
#pragma unroll
            for (int w2 = 0; w2 < kMatchThreads / 32; ++w2) { if (w2 < warp) off += s_wcnt[w2]; tot += s_wcnt[w2]; }
            if (take) s_list[off] = (unsigned short)g;
            __syncthreads();
            if (threadIdx.x == 0) s_nlist += tot;
            __syncthreads();
        }
    }
    const int nlist = s_nlist;
    const float area_b = (box.z - box.x) * (box.w - box.y);
    float best = 0.f;           // IoU with every skipped truth is 0; ties keep the lowest index
    int best_g = 0;
    const bool first_warp_of_row = (blockIdx.x == 0) && (threadIdx.x < 32);
    // second, finer cull with the same argument per warp (32 consecutive anchors = a few cells of one row of
    // the feature map): a truth that cannot intersect any anchor box of the warp is skipped by the warp
    const bool warp_all = first_warp_of_row || !(wx1 <= wx2) || !(wy1 <= wy2);
    unsigned long long* bp_row = best_prior + (size_t)b * Gmax;
    for (int c0 = 0; c0 < nlist; c0 += 32) {
        // lane = one listed truth: test it against the warp's bounding box, then walk the survivors in
        // ascending g (list order)
        const int li = c0 + lane;
        int g_l = 0;
        bool hit = false;
        if (li < nlist) {
            g_l = s_list[li];
            hit = warp_all;
            if (!hit) {
                const float4 t = s_truth[g_l];
                const float cw = fminf(t.z, wx2) - fmaxf(t.x, wx1);
                const float ch = fminf(t.w, wy2) - fmaxf(t.y, wy1);
                hit = (cw > 0.0f && ch > 0.0f) || !(cw == cw) || !(ch == ch);
            }
        }
        // Walk the survivors in ascending g (list order); the loop is warp-uniform.  Per truth: the anchors' IoUs,
        // each anchor's running best truth, and the warp's best prior for the truth by two warp reductions (max of
        // the ordered IoU, then the lowest prior among the lanes that reach it -- torch.max returns the first
        // index); the lane that OWNS the truth in this round keeps the pair and issues ONE 64-bit global max for it
        // after the walk.  No ballots, branches or atomics inside the walk.
        unsigned todo = __ballot_sync(kFullMask, hit);
        uint32_t my_mx = 0u, my_who = 0xffffffffu;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int g = __shfl_sync(kFullMask, g_l, src);
            const float4 t = s_truth[g];
            // box_utils.py:42-47,62-68 (truth = box_a, anchor = box_b)
            float w = fmaxf(fminf(t.z, box.z) - fmaxf(t.x, box.x), 0.0f);
            float h = fmaxf(fminf(t.w, box.w) - fmaxf(t.y, box.y), 0.0f);
            float inter = w * h;
            float area_t = (t.z - t.x) * (t.w - t.y);
            float iou = inter / (area_t + area_b - inter);
            if (!valid) iou = -1.0f;
            if (iou > best) { best = iou; best_g = g; }        // first maximal index; (0, g = 0) when nothing overlaps
            const uint32_t ord = valid ? float_to_ordered(iou) : 0u;
            const uint32_t mx = __reduce_max_sync(kFullMask, ord);
            const uint32_t who = __reduce_min_sync(kFullMask, (valid && ord == mx) ? (uint32_t)p : 0xffffffffu);
            if (lane == src) { my_mx = mx; my_who = who; }
        }
        // per-truth best prior: only overlapping anchors (IoU > 0 or NaN) can win -- or the row's first warp, which
        // supplies prior 0 for a truth that overlaps nothing
        if (hit && my_who != 0xffffffffu && (my_mx > float_to_ordered(0.0f) || first_warp_of_row))
            atomicMax(bp_row + g_l, ((unsigned long long)my_mx << 32) | (unsigned long long)(0xffffffffu - my_who));
    }
    if (valid) {
        bt_overlap[(size_t)b * P + p] = best;
        bt_idx[(size_t)b * P + p] = best_g;
    }
}

__global__ void __launch_bounds__(kMatchThreads)
match_pass2_kernel(const float4* __restrict__ truths, const float* __restrict__ labels,
                   const int* __restrict__ gt_count, const float4* __restrict__ priors,
                   const float4* __restrict__ arm_loc, int P, int Gmax, float threshold, float v0, float v1,
                   int label_mode, const unsigned long long* __restrict__ best_prior,
                   const float* bt_overlap_tmp, const int* bt_idx_tmp,
                   float4* __restrict__ loc_t, long long* __restrict__ conf_t,
                   int* out_bt_idx, float* out_bt_overlap) {
    // forced[t] = the LAST truth (ascending j, box_utils.py:146-150) whose best prior is anchor p0 + t, or -1: the
    // truths scatter themselves (one shared-memory max each) instead of every anchor scanning every truth
    __shared__ int s_forced[kMatchThreads];
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * kMatchThreads;
    const int p = p0 + threadIdx.x;
    s_forced[threadIdx.x] = -1;
    grid_dependency_wait();                 // match_pass1_kernel
    const int G = gt_count[b];
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += kMatchThreads) {
        const int bp = (int)key_index(best_prior[(size_t)b * Gmax + g]) - p0;
        if (bp >= 0 && bp < kMatchThreads) atomicMax(&s_forced[bp], g);
    }
    __syncthreads();
    if (p >= P) return;
    const size_t o = (size_t)b * P + p;
    if (G <= 0) {
        loc_t[o] = make_float4(0.f, 0.f, 0.f, 0.f);
        conf_t[o] = 0;
        if (out_bt_idx) out_bt_idx[o] = 0;
        if (out_bt_overlap) out_bt_overlap[o] = 0.f;
        return;
    }
    float ov = bt_overlap_tmp[o];
    int idx = bt_idx_tmp[o];
    // box_utils.py:146-150: best_truth_overlap[best_prior_idx] = 2; ascending j, last j wins
    if (s_forced[threadIdx.x] >= 0) { idx = s_forced[threadIdx.x]; ov = 2.0f; }
    const float4 m = __ldg(truths + (size_t)b * Gmax + idx);
    const float lab = __ldg(labels + (size_t)b * Gmax + idx);
    long long conf;
    if (label_mode == 1) conf = (lab >= 0.0f) ? 1 : 0;          // refinedet_multibox_loss.py:78-79
    else if (label_mode == 2) conf = (long long)(lab + 1.0f);   // box_utils.py:107
    else conf = (long long)lab;                                 // box_utils.py:152-156
    if (ov < threshold) conf = 0;                               // box_utils.py:158
    float4 pr = __ldg(priors + p);
    float4 ref = pr;
    if (arm_loc) ref = center_size_box(decode_box(ldg_stream4(arm_loc + o), pr, v0, v1));   // :157
    loc_t[o] = encode_box(m, ref, v0, v1);
    conf_t[o] = conf;
    if (out_bt_idx) out_bt_idx[o] = idx;
    if (out_bt_overlap) out_bt_overlap[o] = ov;
}

// ---------------------------------------------------------------------------------------
// hard-negative mining: neg = the num_neg largest loss_c of the row (positives count as 0)
// ---------------------------------------------------------------------------------------
constexpr int kHnmThreads = 1024;
constexpr int kHnmPerT = 16;            // register-resident rows: P <= 16 * 1024 anchors (RefineDet512: 16,320)

// kRegs: every thread keeps its kHnmPerT (ordered) losses in registers, so the row is read from memory once
// and the select passes touch shared memory only; otherwise every pass re-reads the row (L2).
template <bool kRegs>
__global__ void __launch_bounds__(kHnmThreads)
hnm_select_kernel(const float* __restrict__ loss_c, const unsigned char* __restrict__ pos, int P, int ratio,
                  unsigned char* __restrict__ neg_out, int* __restrict__ num_pos_out) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_misc[4];   // 0 digit, 1 need, 2 done, 3 num_pos
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = loss_c + (size_t)b * P;
    const unsigned char* prow = pos + (size_t)b * P;
    unsigned char* nrow = neg_out + (size_t)b * P;
    if (tid == 0) s_misc[3] = 0;
    __syncthreads();
    uint32_t ord[kRegs ? kHnmPerT : 1];
    int local = 0;
    if (kRegs) {
#pragma unroll
        for (int j = 0; j < kHnmPerT; ++j) {
            const int i = j * kHnmThreads + tid;
            const bool p = i < P && prow[i] != 0;
            local += p ? 1 : 0;
            ord[j] = float_to_ordered((i < P && !p) ? row[i] : 0.0f);      // loss_c[pos] = 0, :117
        }
    } else {
        for (int i = tid; i < P; i += kHnmThreads) local += prow[i] ? 1 : 0;
    }
    local = __reduce_add_sync(kFullMask, local);
    if (lane == 0 && local) atomicAdd(&s_misc[3], (uint32_t)local);
    __syncthreads();
    const int num_pos = (int)s_misc[3];
    if (tid == 0 && num_pos_out) num_pos_out[b] = num_pos;
    long long want = (long long)ratio * num_pos;
    if (want > P - 1) want = P - 1;                        // torch.clamp(max=P-1), :122
    int need = (int)(want < 0 ? 0 : want);
    if (need == 0) {
        for (int i = tid; i < P; i += kHnmThreads) nrow[i] = 0;
        return;
    }
    auto key_of = [&](int j, int i) -> unsigned long long {
        if (kRegs) return ((unsigned long long)ord[j] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
        return make_key(prow[i] ? 0.0f : row[i], (uint32_t)i);
    };
    const int nper = kRegs ? kHnmPerT : (P + kHnmThreads - 1) / kHnmThreads;
    unsigned long long prefix = 0, thresh_key = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kHnmThreads) hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (kRegs ? kHnmPerT : 1); ++j) {
            for (int jj = kRegs ? j : 0; jj < (kRegs ? j + 1 : nper); ++jj) {
                const int i = jj * kHnmThreads + tid;
                if (i < P) {
                    const unsigned long long k = key_of(kRegs ? j : 0, i);
                    if (shift == 56 || (k >> (shift + 8)) == prefix) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t loc[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; s += loc[q]; }
            uint32_t v = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_down_sync(kFullMask, v, d);
                if (lane + d < 32) v += o;
            }
            uint32_t cum = v - s;
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
                s_misc[0] = (uint32_t)found;
                s_misc[1] = new_need;
                s_misc[2] = (hist[found] == new_need) ? 1u : 0u;
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | s_misc[0];
        need = (int)s_misc[1];
        const bool done = s_misc[2] != 0;
        __syncthreads();
        if (done || shift == 0) { thresh_key = prefix << shift; break; }
    }
#pragma unroll
    for (int j = 0; j < (kRegs ? kHnmPerT : 1); ++j) {
        for (int jj = kRegs ? j : 0; jj < (kRegs ? j + 1 : nper); ++jj) {
            const int i = jj * kHnmThreads + tid;
            if (i < P) nrow[i] = key_of(kRegs ? j : 0, i) >= thresh_key ? 1 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Hard-negative mining, P <= 16,384: one thread-block CLUSTER of kHnmSplit CTAs per image row (4 x 256 threads, every
// thread keeps 16 (loss, ~index) keys in registers; 128 CTAs for a batch of 32 instead of 32).  The num_neg-th
// largest key is found by narrowing a key interval [lo, hi]: one histogram round over kHnmBins buckets LINEAR in
// the key ((key - lo) >> shift; lo / hi start at the row's min / max key, so the buckets adapt to the range the
// losses really occupy -- a fixed 8-bit digit of the float pattern would put a whole row of cross-entropy values,
// which share their exponent, into two or three contended bins), per-CTA histograms in shared memory, combined
// by every CTA of the cluster over distributed shared memory.  The bucket where the count from the top crosses
// num_neg becomes the new interval (2048 x narrower per round); once it holds few keys they are listed and
// ranked exactly.  Typically ONE round + the list.  Exact for any input: ties in the loss are ordered by index
// (the low word of the key), degenerate rows (all losses equal) narrow on the index bits.
// ---------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kHnmSplit = 4;
constexpr int kHnmCT = 256;
constexpr int kHnmBins = 2048;
constexpr int kHnmBinsPerT = kHnmBins / kHnmCT;     // 8
constexpr int kHnmList = 512;                       // keys of the final interval, whole cluster

struct HnmClusterSmem {
    uint32_t hist[kHnmBins];
    unsigned long long list[kHnmList];              // this CTA's keys of the final interval
    unsigned long long all[kHnmList];               // the cluster's
    unsigned long long kmin, kmax, kmin_neg;        // this CTA's min / max key, min key of a non-positive anchor
    unsigned long long thresh;
    uint32_t npos, nlist;
    uint32_t wsum[kHnmCT / 32];
    uint32_t found_bin, found_need, found_cnt, total;
};

__global__ void __cluster_dims__(kHnmSplit, 1, 1) __launch_bounds__(kHnmCT)
hnm_cluster_kernel(const float* __restrict__ loss_c, const unsigned char* __restrict__ pos, int P, int ratio,
                   unsigned char* __restrict__ neg_out, int* __restrict__ num_pos_out) {
    __shared__ HnmClusterSmem S;
    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x / kHnmSplit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = loss_c + (size_t)b * P;
    const unsigned char* prow = pos + (size_t)b * P;
    unsigned char* nrow = neg_out + (size_t)b * P;
    const int i0 = r * (kHnmCT * kHnmPerT);
    grid_dependency_wait();                                   // the producer of loss_c / pos (conf_loss kernel)
    if (tid == 0) { S.npos = 0; S.nlist = 0; S.kmin = ~0ull; S.kmax = 0ull; S.kmin_neg = ~0ull; }
    __syncthreads();
    // ---- ordered losses into registers (the key of element i is (ord << 32) | ~i, rebuilt where needed);
    //      positives count as loss 0 (refinedet_multibox_loss.py:117).  Element (j, tid) of this CTA is anchor
    //      i0 + (j / 4) * 1024 + 4 tid + (j % 4): four consecutive anchors per thread and round, so that the losses
    //      arrive as one 16-byte load and the flags as one 4-byte load, all sixteen independent of each other (a load
    //      of the loss that waits for the flag's value costs a second trip to DRAM per element). ---------------------
    uint32_t ord[kHnmPerT];
    auto index_of = [&](int j) -> int { return i0 + (j >> 2) * (4 * kHnmCT) + 4 * tid + (j & 3); };
    auto key_of = [&](int j) -> unsigned long long {          // 0 = not an element of the row
        const int i = index_of(j);
        return i < P ? (((unsigned long long)ord[j] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i)) : 0ull;
    };
    float lossv[kHnmPerT];
    unsigned char posv[kHnmPerT];
    const bool vec = (P & 3) == 0 && (reinterpret_cast<uintptr_t>(loss_c) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 3) == 0;
    if (vec) {
#pragma unroll
        for (int j4 = 0; j4 < kHnmPerT / 4; ++j4) {
            const int i = index_of(4 * j4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            uchar4 f = make_uchar4(0, 0, 0, 0);
            if (i < P) {                                          // P % 4 == 0: the four are in or out together
                v = ldg_stream4(reinterpret_cast<const float4*>(row + i));
                f = __ldg(reinterpret_cast<const uchar4*>(prow + i));
            }
            lossv[4 * j4] = v.x; lossv[4 * j4 + 1] = v.y; lossv[4 * j4 + 2] = v.z; lossv[4 * j4 + 3] = v.w;
            posv[4 * j4] = f.x; posv[4 * j4 + 1] = f.y; posv[4 * j4 + 2] = f.z; posv[4 * j4 + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kHnmPerT; ++j) {
            const int i = index_of(j);
            lossv[j] = i < P ? __ldg(row + i) : 0.f;
            posv[j] = i < P ? __ldg(prow + i) : (unsigned char)0;
        }
    }
    int local = 0;
    unsigned long long mn = ~0ull, mx = 0ull, mn_neg = ~0ull;   // mn_neg: over the non-positive anchors only
#pragma unroll
    for (int j = 0; j < kHnmPerT; ++j) {
        const int i = index_of(j);
        ord[j] = 0u;
        if (i < P) {
            const bool p = posv[j] != 0;
            local += p ? 1 : 0;
            ord[j] = float_to_ordered(p ? 0.0f : lossv[j]);
            const unsigned long long k = key_of(j);
            mn = k < mn ? k : mn;
            mx = k > mx ? k : mx;
            if (!p) mn_neg = k < mn_neg ? k : mn_neg;
        }
    }
    local = __reduce_add_sync(kFullMask, local);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(kFullMask, mn, d), c = __shfl_xor_sync(kFullMask, mx, d);
        const unsigned long long e = __shfl_xor_sync(kFullMask, mn_neg, d);
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
        mn_neg = e < mn_neg ? e : mn_neg;
    }
    if (lane == 0) {
        if (local) atomicAdd(&S.npos, (uint32_t)local);
        atomicMin(&S.kmin, mn);
        atomicMax(&S.kmax, mx);
        atomicMin(&S.kmin_neg, mn_neg);
    }
    cluster.sync();
    int num_pos = 0;
    unsigned long long lo_all = ~0ull, lo = ~0ull, hi = 0ull;
#pragma unroll
    for (int q = 0; q < kHnmSplit; ++q) {
        const HnmClusterSmem* R = cluster.map_shared_rank(&S, q);
        num_pos += (int)R->npos;
        lo_all = R->kmin < lo_all ? R->kmin : lo_all;
        lo = R->kmin_neg < lo ? R->kmin_neg : lo;
        hi = R->kmax > hi ? R->kmax : hi;
    }
    // The first interval starts at the smallest NEGATIVE's key: the positives sit at loss 0, far below the losses
    // of the negatives, and would stretch the first round's buckets over a range that holds nothing (all negatives
    // of a row of cross-entropy values in a dozen contended bins).  Should num_neg reach below that key (rows with
    // more positives than a third of the anchors), the search moves on to [min key, lo - 1] -- see `below`.
    if (lo > hi) lo = lo_all;                                  // no negative at all
    if (r == 0 && tid == 0 && num_pos_out) num_pos_out[b] = num_pos;
    long long want = (long long)ratio * num_pos;
    if (want > P - 1) want = P - 1;                           // torch.clamp(max=P-1), :122
    int need = (int)(want < 0 ? 0 : want);
    unsigned long long thresh_key = ~0ull;                    // need == 0: nothing selected
    bool listed = false;
    while (need > 0) {
        // ---- one narrowing round over [lo, hi] ---------------------------------------------------------------
        const unsigned long long range = hi - lo;
        int shift = (64 - __clzll((long long)(range | 1ull))) - 11;      // (range >> shift) < kHnmBins
        if (shift < 0) shift = 0;
        for (int i = tid; i < kHnmBins; i += kHnmCT) S.hist[i] = 0;
        if (tid == 0) S.found_bin = 0xffffffffu;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kHnmPerT; ++j) {
            const unsigned long long k = key_of(j);
            if (k >= lo && k <= hi && k != 0ull) atomicAdd(&S.hist[(uint32_t)((k - lo) >> shift)], 1u);
        }
        cluster.sync();
        // combined histogram: thread t owns bins [8 t, 8 t + 8); suffix sums from the top
        uint32_t c[kHnmBinsPerT];
        uint32_t s = 0;
#pragma unroll
        for (int q = 0; q < kHnmBinsPerT; ++q) c[q] = 0;
#pragma unroll
        for (int rr = 0; rr < kHnmSplit; ++rr) {
            const uint4* h4 = reinterpret_cast<const uint4*>(cluster.map_shared_rank(S.hist, rr)) + tid * (kHnmBinsPerT / 4);
#pragma unroll
            for (int q4 = 0; q4 < kHnmBinsPerT / 4; ++q4) {
                const uint4 v = h4[q4];
                c[q4 * 4 + 0] += v.x; c[q4 * 4 + 1] += v.y; c[q4 * 4 + 2] += v.z; c[q4 * 4 + 3] += v.w;
            }
        }
#pragma unroll
        for (int q = 0; q < kHnmBinsPerT; ++q) s += c[q];
        uint32_t incl = s;                                    // inclusive suffix sum over the lanes (higher lane = higher keys)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_down_sync(kFullMask, incl, d);
            if (lane + d < 32) incl += o;
        }
        if (lane == 0) S.wsum[warp] = incl;
        __syncthreads();
        uint32_t above = incl - s;                            // keys in bins owned by higher threads
#pragma unroll
        for (int w = 0; w < kHnmCT / 32; ++w) if (w > warp) above += S.wsum[w];
#pragma unroll
        for (int q = kHnmBinsPerT - 1; q >= 0; --q) {
            if (above < (uint32_t)need && above + c[q] >= (uint32_t)need) {
                S.found_bin = (uint32_t)(tid * kHnmBinsPerT + q);
                S.found_need = (uint32_t)need - above;
                S.found_cnt = c[q];
            }
            above += c[q];
        }
        if (tid == 0) S.total = above;                        // keys inside [lo, hi]
        __syncthreads();
        const uint32_t fb = S.found_bin, fcnt = S.found_cnt;
        if (fb == 0xffffffffu) {                              // below: num_neg reaches under the interval (see above)
            need -= (int)S.total;
            hi = lo - 1ull;
            lo = lo_all;
            cluster.sync();                                   // every CTA has read every histogram
            continue;
        }
        need = (int)S.found_need;
        const unsigned long long nlo = lo + ((unsigned long long)fb << shift);
        unsigned long long nhi = nlo + ((1ull << shift) - 1ull);
        if (nhi > hi || nhi < nlo) nhi = hi;
        lo = nlo; hi = nhi;
        cluster.sync();                                       // every CTA has read every histogram: they may be reused
        if (fcnt == (uint32_t)need) { thresh_key = lo; break; }       // the whole bucket is selected
        if (fcnt <= (uint32_t)kHnmList) { listed = true; break; }
    }
    if (listed) {
        // ---- the final interval holds few keys: list them, rank exactly ---------------------------------------
#pragma unroll
        for (int j = 0; j < kHnmPerT; ++j) {
            const unsigned long long k = key_of(j);
            if (k >= lo && k <= hi && k != 0ull) S.list[atomicAdd(&S.nlist, 1u)] = k;
        }
        cluster.sync();
        int tot = 0;
#pragma unroll
        for (int rr = 0; rr < kHnmSplit; ++rr) {
            const HnmClusterSmem* R = cluster.map_shared_rank(&S, rr);
            const int n_r = (int)R->nlist;
            for (int i = tid; i < n_r; i += kHnmCT) S.all[tot + i] = R->list[i];
            tot += n_r;
        }
        __syncthreads();
        for (int i = tid; i < tot; i += kHnmCT) {
            const unsigned long long k = S.all[i];
            int rank = 0;
            for (int q = 0; q < tot; ++q) rank += S.all[q] > k ? 1 : 0;
            if (rank == need - 1) S.thresh = k;
        }
        __syncthreads();
        thresh_key = S.thresh;
    }
    if (vec && (reinterpret_cast<uintptr_t>(neg_out) & 3) == 0) {
#pragma unroll
        for (int j4 = 0; j4 < kHnmPerT / 4; ++j4) {
            const int i = index_of(4 * j4);
            if (i < P) {
                uchar4 o;
                o.x = (need > 0 && key_of(4 * j4) >= thresh_key) ? 1 : 0;
                o.y = (need > 0 && key_of(4 * j4 + 1) >= thresh_key) ? 1 : 0;
                o.z = (need > 0 && key_of(4 * j4 + 2) >= thresh_key) ? 1 : 0;
                o.w = (need > 0 && key_of(4 * j4 + 3) >= thresh_key) ? 1 : 0;
                *reinterpret_cast<uchar4*>(nrow + i) = o;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < kHnmPerT; ++j) {
            const int i = index_of(j);
            if (i < P) nrow[i] = (need > 0 && key_of(j) >= thresh_key) ? 1 : 0;
        }
    }
    cluster.sync();                                           // nobody leaves while a peer may still read its shared memory
}

// ---------------------------------------------------------------------------------------
// target ingestion: the B ragged [G_i, 5] target tensors of a step (data/__init__.py:9-27 detection_collate),
// concatenated, -> padded truths[B,Gmax,4], labels[B,Gmax], gt_count[B] in one pass
// ---------------------------------------------------------------------------------------
__global__ void pad_targets_kernel(const float* __restrict__ flat, const int* __restrict__ offs, int B, int Gmax,
                                   float4* __restrict__ truths, float* __restrict__ labels, int* __restrict__ gt_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Gmax) return;
    const int b = i / Gmax, g = i - b * Gmax;
    const int o = offs[b], n = offs[b + 1] - o;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    float lab = 0.f;
    if (g < n) {
        const float* r = flat + (size_t)(o + g) * 5;
        t = make_float4(r[0], r[1], r[2], r[3]);
        lab = r[4];
    }
    truths[i] = t;
    labels[i] = lab;
    if (g == 0) gt_count[b] = n;
}

// ---------------------------------------------------------------------------------------
// element-wise box_utils functions
// ---------------------------------------------------------------------------------------
enum { OP_POINT_FORM = 0, OP_CENTER_SIZE = 1 };
template <int OP>
__global__ void unary_box_kernel(const float4* __restrict__ in, float4* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = in[i];
    out[i] = OP == OP_POINT_FORM ? point_form_box(b) : center_size_box(b);
}
template <bool ENCODE>
__global__ void coder_kernel(const float4* __restrict__ a, const float4* __restrict__ priors, float v0, float v1,
                             float4* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = ENCODE ? encode_box(a[i], priors[i], v0, v1) : decode_box(a[i], priors[i], v0, v1);
}
template <bool IOU>
__global__ void pairwise_kernel(const float4* __restrict__ box_a, const float4* __restrict__ box_b,
                                float* __restrict__ out, int A, int Bn) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= Bn || i >= A) return;
    float4 a = __ldg(box_a + i), b = box_b[j];
    out[(size_t)i * Bn + j] = IOU ? jaccard_pair(a, b) : intersect_pair(a, b);
}


int hnm_launch(const float* loss_c, const unsigned char* pos, int B, int P, int negpos_ratio, unsigned char* neg_out,
               int* num_pos_out, cudaStream_t st, bool pdl) {
    cudaError_t e = cudaSuccess;
    if (P <= kHnmSplit * kHnmCT * kHnmPerT) {
        // one cluster of kHnmSplit CTAs per image (the cluster shape is a compile-time attribute of the kernel)
        if (pdl) e = launch_pdl(hnm_cluster_kernel, dim3(B * kHnmSplit), dim3(kHnmCT), 0, st, loss_c, pos, P, negpos_ratio,
                                neg_out, num_pos_out);
        else hnm_cluster_kernel<<<B * kHnmSplit, kHnmCT, 0, st>>>(loss_c, pos, P, negpos_ratio, neg_out, num_pos_out);
    } else {
        hnm_select_kernel<false><<<B, kHnmThreads, 0, st>>>(loss_c, pos, P, negpos_ratio, neg_out, num_pos_out);
    }
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}


int match_launch(const float4* truths, const float* labels, const int* gt_count, const float4* priors,
                 const float4* arm_loc, int B, int P, int Gmax, float threshold, float v0, float v1, int label_mode,
                 unsigned long long* best_prior, float* bt_overlap, int* bt_idx, float4* loc_t, long long* conf_t,
                 cudaStream_t st, bool pdl) {
    const dim3 grid((P + kMatchThreads - 1) / kMatchThreads, B);
    cudaError_t e;
    // pass 1 follows the memset of best_prior (not a kernel): plain stream order
    match_pass1_kernel<<<grid, kMatchThreads, (size_t)Gmax * 18, st>>>(truths, gt_count, priors, arm_loc, P, Gmax, v0, v1,
                                                                         best_prior, bt_overlap, bt_idx);
    e = cudaSuccess;
    note_launch();
    RD_CHECK_LAUNCH();
    if (pdl) e = launch_pdl(match_pass2_kernel, grid, dim3(kMatchThreads), (size_t)Gmax * 4, st, truths, labels, gt_count, priors,
                            arm_loc, P, Gmax, threshold, v0, v1, label_mode, (const unsigned long long*)best_prior,
                            (const float*)bt_overlap, (const int*)bt_idx, loc_t, conf_t, bt_idx, bt_overlap);
    else
        match_pass2_kernel<<<grid, kMatchThreads, (size_t)Gmax * 4, st>>>(truths, labels, gt_count, priors, arm_loc, P, Gmax,
                                                                            threshold, v0, v1, label_mode, best_prior, bt_overlap,
                                                                            bt_idx, loc_t, conf_t, bt_idx, bt_overlap);
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
}  // namespace rd

using namespace rd;

extern "C" {

#define RD_ALIGNED16(p) ((((uintptr_t)(p)) & 15) == 0)

int rd_point_form(const float* boxes, float* out, int n, void* stream) {
    if (n == 0) return 0;
    if (!boxes || !out || n < 0) return RD_ERR_BAD_ARG;
    if (!RD_ALIGNED16(boxes) || !RD_ALIGNED16(out)) return RD_ERR_ALIGNMENT;
    unary_box_kernel<OP_POINT_FORM><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, (float4*)out, n);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
int rd_center_size(const float* boxes, float* out, int n, void* stream) {
    if (n == 0) return 0;
    if (!boxes || !out || n < 0) return RD_ERR_BAD_ARG;
    if (!RD_ALIGNED16(boxes) || !RD_ALIGNED16(out)) return RD_ERR_ALIGNMENT;
    unary_box_kernel<OP_CENTER_SIZE><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)boxes, (float4*)out, n);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
int rd_decode(const float* loc, const float* priors, float v0, float v1, float* out, int n, void* stream) {
    if (n == 0) return 0;
    if (!loc || !priors || !out || n < 0) return RD_ERR_BAD_ARG;
    if (!RD_ALIGNED16(loc) || !RD_ALIGNED16(priors) || !RD_ALIGNED16(out)) return RD_ERR_ALIGNMENT;
    coder_kernel<false><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)loc, (const float4*)priors, v0, v1, (float4*)out, n);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
int rd_encode(const float* matched, const float* priors, float v0, float v1, float* out, int n, void* stream) {
    if (n == 0) return 0;
    if (!matched || !priors || !out || n < 0) return RD_ERR_BAD_ARG;
    if (!RD_ALIGNED16(matched) || !RD_ALIGNED16(priors) || !RD_ALIGNED16(out)) return RD_ERR_ALIGNMENT;
    coder_kernel<true><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)matched, (const float4*)priors, v0, v1, (float4*)out, n);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
static int pairwise(bool iou, const float* box_a, const float* box_b, float* out, int A, int Bn, void* stream) {
    if (A == 0 || Bn == 0) return 0;
    if (!box_a || !box_b || !out || A < 0 || Bn < 0) return RD_ERR_BAD_ARG;
    if (A > 65535) return RD_ERR_UNSUPPORTED;
    if (!RD_ALIGNED16(box_a) || !RD_ALIGNED16(box_b)) return RD_ERR_ALIGNMENT;
    dim3 grid((Bn + 255) / 256, A);
    if (iou) pairwise_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)box_a, (const float4*)box_b, out, A, Bn);
    else pairwise_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)box_a, (const float4*)box_b, out, A, Bn);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}
int rd_intersect(const float* box_a, const float* box_b, float* out, int A, int Bn, void* stream) {
    return pairwise(false, box_a, box_b, out, A, Bn, stream);
}
int rd_jaccard(const float* box_a, const float* box_b, float* out, int A, int Bn, void* stream) {
    return pairwise(true, box_a, box_b, out, A, Bn, stream);
}

size_t rd_match_workspace_bytes(int B, int Gmax) {
    if (B <= 0 || Gmax <= 0) return 256;
    return ((size_t)B * Gmax * 8 + 255) / 256 * 256;
}

int rd_refine_match(const float* truths, const float* labels, const int* gt_count, const float* priors,
                    const float* arm_loc, int B, int P, int Gmax, float threshold, float v0, float v1,
                    int label_mode, void* workspace, size_t workspace_bytes, float* loc_t, long long* conf_t,
                    int* best_truth_idx, float* best_truth_overlap, void* stream) {
    NvtxRange nvtx_range("rd_refine_match");
    if (!truths || !labels || !gt_count || !priors || !workspace || !loc_t || !conf_t) return RD_ERR_BAD_ARG;
    if (B <= 0 || P <= 0 || Gmax <= 0 || label_mode < 0 || label_mode > 2) return RD_ERR_BAD_ARG;
    if (Gmax > RD_MAX_GT || B > 65535) return RD_ERR_UNSUPPORTED;
    if (!RD_ALIGNED16(truths) || !RD_ALIGNED16(priors) || !RD_ALIGNED16(loc_t) || (arm_loc && !RD_ALIGNED16(arm_loc)) ||
        !RD_ALIGNED16(workspace) || (((uintptr_t)conf_t) & 7))
        return RD_ERR_ALIGNMENT;
    if (workspace_bytes < rd_match_workspace_bytes(B, Gmax)) return RD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* best_prior = (unsigned long long*)workspace;
    cudaError_t e = cudaMemsetAsync(best_prior, 0, (size_t)B * Gmax * 8, st);
    if (e != cudaSuccess) return (int)e;
    // best_truth_idx / best_truth_overlap are outputs AND the scratch between the two passes
    if (!best_truth_idx || !best_truth_overlap) return RD_ERR_BAD_ARG;
    return match_launch((const float4*)truths, labels, gt_count, (const float4*)priors, (const float4*)arm_loc, B, P, Gmax,
                        threshold, v0, v1, label_mode, best_prior, best_truth_overlap, best_truth_idx, (float4*)loc_t, conf_t,
                        st, false);
}

int rd_pad_targets(const float* flat, const int* offsets, int B, int Gmax, float* truths, float* labels,
                   int* gt_count, void* stream) {
    NvtxRange nvtx_range("rd_pad_targets");
    if (!flat || !offsets || !truths || !labels || !gt_count || B <= 0 || Gmax <= 0) return RD_ERR_BAD_ARG;
    if ((uintptr_t)truths & 15) return RD_ERR_ALIGNMENT;
    const int n = B * Gmax;
    pad_targets_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flat, offsets, B, Gmax, (float4*)truths, labels,
                                                                        gt_count);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_hnm_select(const float* loss_c, const unsigned char* pos, int B, int P, int negpos_ratio,
                  unsigned char* neg_out, int* num_pos_out, void* stream) {
    NvtxRange nvtx_range("rd_hnm_select");
    if (!loss_c || !pos || !neg_out || B <= 0 || P <= 0 || negpos_ratio < 0) return RD_ERR_BAD_ARG;
    return hnm_launch(loss_c, pos, B, P, negpos_ratio, neg_out, num_pos_out, (cudaStream_t)stream, false);
}

int rd_abi_version(void) { return RD_ABI_VERSION; }

const char* rd_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case RD_ERR_BAD_ARG: return "refinedet_b200: bad argument (null pointer or non-positive size)";
        case RD_ERR_ALIGNMENT: return "refinedet_b200: tensor is not 16-byte aligned";
        case RD_ERR_UNSUPPORTED: return "refinedet_b200: size beyond the supported limit";
        case RD_ERR_WORKSPACE: return "refinedet_b200: workspace too small";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "refinedet_b200: unknown error";
}

}  // extern "C"

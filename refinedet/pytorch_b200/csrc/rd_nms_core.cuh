// rd_nms_core.cuh — one NMS problem per CTA: top-k select, sort, spatially-culled
// exact greedy NMS.  Shared by the fused detect stage (one problem per
// (image, class)) and the stand-alone rd_nms / rd_nms_host entry points.
//
// Reference semantics reproduced:
//   layers/box_utils.py:222-286 (normalised flavour), utils/nms/py_cpu_nms.py:10-38 ==
//   utils/nms/nms_kernel.cu:24-32,124-140 (pixel +1 flavour), utils/nms/cpu_nms.pyx:65
//   (suppress-on-equal variant), eval_refinedet_coco.py:222,231 (top_k, per-class cap).
//
// Algorithm (B200-first, not the reference's dense n x n bitmask):
//   1. keys = (ordered score bits << 32) | ~index; if n > top_k an MSB-first 8-bit
//      radix select over the L2-resident key list finds the exact top_k set.
//   2. bitonic sort of the <= top_k keys in shared memory (descending).
//   3. boxes gathered into shared memory (SoA), each box is binned into 32 columns
//      and 32 rows of the problem's own bounding extent; prefix-OR tables
//      S[c] = {boxes starting at or before column c}, E[c] = {boxes ending before c}
//      turn "which earlier boxes can intersect box j at all" into a few word-wide
//      ANDs:  X_j = S[b_j] & ~E[a_j].  Only those pairs get the exact fp32 IoU test,
//      so the work is O(n^2/32) word operations + O(#intersecting pairs) IoUs
//      instead of O(n^2) IoUs.  The cull is conservative (monotone binning), so the
//      kept set is exactly the greedy NMS result.
//   4. one warp walks the score order 32 boxes at a time: candidates are tested only
//      against KEPT earlier boxes, in-block dependencies are resolved with ballots.
#pragma once
#include "rd_common.cuh"

namespace rd {

constexpr int kNmsThreads = 128;
constexpr int kCols = 32;  // spatial bins per axis

struct NmsSmemLayout {
    int mcap;   // max boxes held (multiple of 32)
    int W;      // mcap / 32
    int WS;     // padded row stride of the column tables (odd -> conflict-free)
    int Kp;     // power of two >= mcap (bitonic sort buffer)
    size_t off_keys, off_x1, off_y1, off_x2, off_y2, off_area, off_cr, off_tab, off_keptbits,
        off_keptidx, off_hist, off_misc, total;
};

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

__host__ __device__ inline NmsSmemLayout nms_layout(int mcap_req) {
    NmsSmemLayout L;
    int mcap = ((mcap_req < 1 ? 1 : mcap_req) + 31) & ~31;
    L.mcap = mcap;
    L.W = mcap / 32;
    L.WS = L.W | 1;
    L.Kp = next_pow2(mcap);
    size_t o = 0;
    L.off_keys = o;      o += (size_t)L.Kp * 8;
    L.off_x1 = o;        o += (size_t)mcap * 4;
    L.off_y1 = o;        o += (size_t)mcap * 4;
    L.off_x2 = o;        o += (size_t)mcap * 4;
    L.off_y2 = o;        o += (size_t)mcap * 4;
    L.off_area = o;      o += (size_t)mcap * 4;
    L.off_cr = o;        o += (size_t)mcap * 4;
    L.off_tab = o;       o += (size_t)4 * kCols * L.WS * 4;
    L.off_keptbits = o;  o += (size_t)L.W * 4;
    L.off_keptidx = o;   o += (size_t)mcap * 4;
    L.off_hist = o;      o += 256 * 4;
    L.off_misc = o;      o += 16 * 4;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

struct NmsProblem {
    const unsigned long long* cand;  // n keys in global memory
    int n;
    const float4* boxes;             // boxes[key index] (point form)
    float4 scale;                    // multiplied into the boxes when has_scale
    int has_scale;
    float thr;
    int top_k;
    int max_out;
    int flags;
};

// exact suppression test: does kept box i (higher score) suppress candidate j ?
__device__ __forceinline__ bool suppresses(float x1i, float y1i, float x2i, float y2i, float ai,
                                           float x1j, float y1j, float x2j, float y2j, float aj,
                                           float thr, int flags) {
    float iou;
    if (flags & RD_NMS_PIXEL_PLUS1) {
        // py_cpu_nms.py:25-33 / nms_kernel.cu:24-32
        float w = fmaxf(0.0f, fminf(x2i, x2j) - fmaxf(x1i, x1j) + 1.0f);
        float h = fmaxf(0.0f, fminf(y2i, y2j) - fmaxf(y1i, y1j) + 1.0f);
        float inter = w * h;
        iou = inter / (ai + aj - inter);
    } else {
        // box_utils.py:268-283: union = (rem_areas - inter) + area[i]
        float w = fmaxf(fminf(x2j, x2i) - fmaxf(x1j, x1i), 0.0f);
        float h = fmaxf(fminf(y2j, y2i) - fmaxf(y1j, y1i), 0.0f);
        float inter = w * h;
        iou = inter / ((aj - inter) + ai);
    }
    return (flags & RD_NMS_SUPPRESS_EQ) ? (iou >= thr) : !(iou <= thr);
}

__device__ __forceinline__ int col_of(float v, float lo, float inv) {
    float f = floorf((v - lo) * inv);
    f = fminf(fmaxf(f, 0.0f), (float)(kCols - 1));
    return (int)f;   // NaN -> fmaxf/fminf drop it -> 0
}

// Runs one problem on the calling CTA (kNmsThreads threads).  Returns the number of kept
// boxes (uniform over the CTA); kept sorted positions are in keptidx[0..count), and
// keys/x1..y2 hold the sorted candidates so the caller can emit rows.
__device__ inline int nms_process(unsigned char* smem, const NmsSmemLayout& L, const NmsProblem& pb) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.off_keys);
    float* sx1 = reinterpret_cast<float*>(smem + L.off_x1);
    float* sy1 = reinterpret_cast<float*>(smem + L.off_y1);
    float* sx2 = reinterpret_cast<float*>(smem + L.off_x2);
    float* sy2 = reinterpret_cast<float*>(smem + L.off_y2);
    float* sarea = reinterpret_cast<float*>(smem + L.off_area);
    uint32_t* scr = reinterpret_cast<uint32_t*>(smem + L.off_cr);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + L.off_tab);
    uint32_t* keptbits = reinterpret_cast<uint32_t*>(smem + L.off_keptbits);
    int* keptidx = reinterpret_cast<int*>(smem + L.off_keptidx);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L.off_hist);
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L.off_misc);
    // misc: 0 select counter, 1 digit, 2 need, 3 done, 4..7 ordered min/max, 8 kept count

    const int n = pb.n;
    const int top_k = pb.top_k < L.mcap ? pb.top_k : L.mcap;
    const int m = n < top_k ? n : top_k;
    if (m <= 0) return 0;

    // ---- 1. load or select the m highest keys --------------------------------
    if (n <= top_k) {
        for (int i = tid; i < n; i += kNmsThreads) keys[i] = pb.cand[i];
    } else {
        unsigned long long prefix = 0;   // known high bits of the threshold key
        int need = top_k;                // how many keys of the current bucket are wanted
        unsigned long long thresh_key = 0;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += kNmsThreads) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += kNmsThreads) {
                unsigned long long k = pb.cand[i];
                bool match = (shift == 56) || ((k >> (shift + 8)) == prefix);
                if (match) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                // lane l owns digits [8l, 8l+8); find the digit where the count from the
                // top crosses `need`
                uint32_t loc[8];
                uint32_t s = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; s += loc[q]; }
                // suffix sum over lanes above me (exclusive)
                uint32_t above = 0;
                uint32_t v = s;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t o = __shfl_down_sync(kFullMask, v, d);
                    if (lane + d < 32) v += o;
                }
                above = v - s;  // keys in digits of higher lanes
                int found = -1; uint32_t new_need = 0;
                uint32_t cum = above;
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
            bool done = misc[3] != 0;
            __syncthreads();
            if (done || shift == 0) { thresh_key = prefix << shift; break; }
        }
        if (tid == 0) misc[0] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kNmsThreads) {
            unsigned long long k = pb.cand[i];
            if (k >= thresh_key) {
                uint32_t pos = atomicAdd(&misc[0], 1u);
                if (pos < (uint32_t)top_k) keys[pos] = k;
            }
        }
    }
    const int Kp = next_pow2(m);
    for (int i = m + tid; i < Kp; i += kNmsThreads) keys[i] = 0ull;
    __syncthreads();

    // ---- 2. bitonic sort, descending -------------------------------------------
    for (int k = 2; k <= Kp; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (Kp >> 1); t += kNmsThreads) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int l = i + j;
                bool desc = (i & k) == 0;
                unsigned long long a = keys[i], b = keys[l];
                if ((a < b) == desc) { keys[i] = b; keys[l] = a; }
            }
            __syncthreads();
        }
    }

    // ---- 3. gather boxes, areas, extent ----------------------------------------
    if (tid < 4) misc[4 + tid] = (tid & 1) ? 0u : 0xffffffffu;   // 4:min x, 5:max x, 6:min y, 7:max y
    if (tid == 0) misc[8] = 0;
    __syncthreads();
    const bool pixel = (pb.flags & RD_NMS_PIXEL_PLUS1) != 0;
    {
        uint32_t mnx = 0xffffffffu, mxx = 0, mny = 0xffffffffu, mxy = 0;
        for (int i = tid; i < m; i += kNmsThreads) {
            float4 b = pb.boxes[key_index(keys[i])];
            if (pb.has_scale) { b.x *= pb.scale.x; b.y *= pb.scale.y; b.z *= pb.scale.z; b.w *= pb.scale.w; }
            float area = pixel ? (b.z - b.x + 1.0f) * (b.w - b.y + 1.0f) : (b.z - b.x) * (b.w - b.y);
            sx1[i] = b.x; sy1[i] = b.y; sx2[i] = b.z; sy2[i] = b.w; sarea[i] = area;
            // binning extent = range of the box CENTRES (a few huge boxes must not coarsen the bins);
            // boxes reaching beyond it clamp to the edge bins, which stays conservative
            float cx = 0.5f * b.x + 0.5f * b.z, cy = 0.5f * b.y + 0.5f * b.w;
            if (isfinite(cx)) { mnx = min(mnx, float_to_ordered(cx)); mxx = max(mxx, float_to_ordered(cx)); }
            if (isfinite(cy)) { mny = min(mny, float_to_ordered(cy)); mxy = max(mxy, float_to_ordered(cy)); }
        }
        mnx = __reduce_min_sync(kFullMask, mnx); mxx = __reduce_max_sync(kFullMask, mxx);
        mny = __reduce_min_sync(kFullMask, mny); mxy = __reduce_max_sync(kFullMask, mxy);
        if (lane == 0) {
            atomicMin(&misc[4], mnx); atomicMax(&misc[5], mxx);
            atomicMin(&misc[6], mny); atomicMax(&misc[7], mxy);
        }
    }
    for (int i = tid; i < 4 * kCols * L.WS; i += kNmsThreads) tab[i] = 0;
    __syncthreads();

    // ---- 4. column ranges + start/end marks -------------------------------------
    const int Wm = (m + 31) >> 5;   // words in use
    {
        float lox = 0.f, invx = 0.f, loy = 0.f, invy = 0.f;
        if (misc[4] <= misc[5]) {
            lox = ordered_to_float(misc[4]);
            float hi = ordered_to_float(misc[5]);
            invx = hi > lox ? (float)kCols / (hi - lox) : 0.f;
        }
        if (misc[6] <= misc[7]) {
            loy = ordered_to_float(misc[6]);
            float hi = ordered_to_float(misc[7]);
            invy = hi > loy ? (float)kCols / (hi - loy) : 0.f;
        }
        if (!isfinite(invx)) invx = 0.f;
        if (!isfinite(invy)) invy = 0.f;
        const bool eq = (pb.flags & RD_NMS_SUPPRESS_EQ) != 0;
        // culling assumes "no intersection => not suppressed"; false for these thresholds
        const bool force_full = eq ? !(pb.thr > 0.0f) : !(pb.thr >= 0.0f);
        const float eps = 9.5367431640625e-07f;   // 2^-20 relative nudge (conservative binning)
        for (int i = tid; i < m; i += kNmsThreads) {
            float x1 = sx1[i], y1 = sy1[i], x2 = sx2[i], y2 = sy2[i], ar = sarea[i];
            float hx = pixel ? x2 + 1.0f : x2, hy = pixel ? y2 + 1.0f : y2;
            int ax, bx, ay, by;
            bool ok = isfinite(x1) && isfinite(y1) && isfinite(hx) && isfinite(hy) && (ar > 0.0f) &&
                      isfinite(ar) && !force_full;
            if (ok) {
                ax = col_of(x1 - fabsf(x1) * eps, lox, invx);
                bx = col_of(hx + fabsf(hx) * eps, lox, invx);
                ay = col_of(y1 - fabsf(y1) * eps, loy, invy);
                by = col_of(hy + fabsf(hy) * eps, loy, invy);
            } else {
                ax = 0; bx = kCols - 1; ay = 0; by = kCols - 1;   // always tested exactly
            }
            scr[i] = (uint32_t)ax | ((uint32_t)bx << 8) | ((uint32_t)ay << 16) | ((uint32_t)by << 24);
            const uint32_t bit = 1u << (i & 31);
            const int w = i >> 5;
            atomicOr(&tab[(0 * kCols + ax) * L.WS + w], bit);
            if (bx + 1 < kCols) atomicOr(&tab[(1 * kCols + bx + 1) * L.WS + w], bit);
            atomicOr(&tab[(2 * kCols + ay) * L.WS + w], bit);
            if (by + 1 < kCols) atomicOr(&tab[(3 * kCols + by + 1) * L.WS + w], bit);
        }
    }
    __syncthreads();
    // inclusive prefix-OR over columns: S[c] = starts at <= c ; E[c] = ends (b) < c
    for (int task = tid; task < 4 * Wm; task += kNmsThreads) {
        int t = task / Wm, w = task - t * Wm;
        uint32_t acc = 0;
        uint32_t* p = tab + (size_t)t * kCols * L.WS + w;
#pragma unroll 8
        for (int c = 0; c < kCols; ++c) { acc |= p[c * L.WS]; p[c * L.WS] = acc; }
    }
    __syncthreads();

    // ---- 5. greedy walk, one warp -------------------------------------------------
    if (warp == 0) {
        const float thr = pb.thr;
        const int flags = pb.flags;
        const int max_out = pb.max_out;
        int kept_total = 0;
        const uint32_t* Sx = tab;
        const uint32_t* Ex = tab + 1 * kCols * L.WS;
        const uint32_t* Sy = tab + 2 * kCols * L.WS;
        const uint32_t* Ey = tab + 3 * kCols * L.WS;
        for (int ib = 0; ib < Wm; ++ib) {
            const int j = ib * 32 + lane;
            const bool valid = j < m;
            bool alive = valid;
            float x1 = 0, y1 = 0, x2 = 0, y2 = 0, ar = 0;
            uint32_t cr = 0;
            if (valid) { x1 = sx1[j]; y1 = sy1[j]; x2 = sx2[j]; y2 = sy2[j]; ar = sarea[j]; cr = scr[j]; }
            const uint32_t* rSx = Sx + ((cr >> 8) & 255u) * L.WS;    // S_x[b_j]
            const uint32_t* rEx = Ex + (cr & 255u) * L.WS;           // E_x[a_j]
            const uint32_t* rSy = Sy + ((cr >> 24) & 255u) * L.WS;   // S_y[b_j]
            const uint32_t* rEy = Ey + ((cr >> 16) & 255u) * L.WS;   // E_y[a_j]
            // earlier blocks: only KEPT boxes that can intersect
            for (int w = 0; w < ib; ++w) {
                uint32_t h = rSx[w] & ~rEx[w] & rSy[w] & ~rEy[w] & keptbits[w];
                while (alive && h) {
                    int i = (w << 5) + __ffs(h) - 1;
                    h &= h - 1;
                    if (suppresses(sx1[i], sy1[i], sx2[i], sy2[i], sarea[i], x1, y1, x2, y2, ar, thr, flags))
                        alive = false;
                }
            }
            // same block: exact bits for earlier lanes, then ordered resolution
            uint32_t tin = 0;
            if (alive) {
                uint32_t h = rSx[ib] & ~rEx[ib] & rSy[ib] & ~rEy[ib] & ((1u << lane) - 1u);
                while (h) {
                    int k = __ffs(h) - 1;
                    h &= h - 1;
                    int i = (ib << 5) + k;
                    if (suppresses(sx1[i], sy1[i], sx2[i], sy2[i], sarea[i], x1, y1, x2, y2, ar, thr, flags))
                        tin |= 1u << k;
                }
            }
            uint32_t u = __reduce_or_sync(kFullMask, tin);
            while (u) {
                int k = __ffs(u) - 1;
                u &= u - 1;
                uint32_t al = __ballot_sync(kFullMask, alive);
                if (((al >> k) & 1u) && ((tin >> k) & 1u)) alive = false;
            }
            uint32_t keptw = __ballot_sync(kFullMask, alive);
            int room = max_out - kept_total;
            int cnt = __popc(keptw);
            if (cnt > room) {   // keep only the first `room` set bits
                uint32_t t = keptw, keep = 0;
                for (int r = 0; r < room; ++r) { uint32_t low = t & (0u - t); keep |= low; t ^= low; }
                keptw = keep;
                cnt = room;
            }
            if ((keptw >> lane) & 1u) keptidx[kept_total + __popc(keptw & ((1u << lane) - 1u))] = j;
            if (lane == 0) keptbits[ib] = keptw;
            __syncwarp();
            kept_total += cnt;
            if (kept_total >= max_out) break;
        }
        if (lane == 0) misc[8] = (uint32_t)kept_total;
    }
    __syncthreads();
    return (int)misc[8];
}


// =========================================================================================
// warp-per-problem path: n <= 256 candidates, no select needed.  One warp
//   1. sorts the keys 32 at a time in registers (bitonic network over shuffles) and merges
//      the sorted runs by rank (binary search of every key in the other runs),
//   2. stages boxes and the bin tables in its private slice of shared memory,
//   3. walks the score order 32 candidates at a time; the (candidate, kept-earlier-box) pairs
//      that survive the bin cull are flattened into a list and tested 32 pairs per step, so a
//      few large boxes that overlap many others do not serialise the warp,
//   4. emits rows straight from the walk.
// No CTA-wide barrier anywhere.
// =========================================================================================
constexpr int kWarpCap = 256;
constexpr int kWarpW = kWarpCap / 32;      // 8 mask words
constexpr int kWarpWS = kWarpW + 1;        // padded row stride
constexpr int kPairCap = 384;              // flattened (candidate, earlier box) pairs per 32-candidate block
struct WarpSmem {
    unsigned long long keys[kWarpCap];
    float x1[kWarpCap], y1[kWarpCap], x2[kWarpCap], y2[kWarpCap], area[kWarpCap];   // x1,y1 double as the
                                                                                    // sorted-runs buffer
    uint32_t tab[4 * kCols * kWarpWS];
    uint32_t pairs[kPairCap];
    uint32_t keptbits[kWarpW];
    uint32_t tin[32];
    uint32_t sup;
    uint32_t pad[7];
};

// 32 keys, one per lane, sorted descending across lanes
__device__ __forceinline__ unsigned long long warp_sort32_desc(unsigned long long a, int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const bool desc = (lane & k2) == 0;
            const bool lower = (lane & j) == 0;
            const unsigned long long b = __shfl_xor_sync(kFullMask, a, j);
            const bool take_max = (lower == desc);
            a = take_max ? (a > b ? a : b) : (a < b ? a : b);
        }
    }
    return a;
}

struct RowSink {           // where the fused stage writes its rows
    float* rows;           // [max_out,5] slot of this (image, class)
    int* anchors;          // [max_out] or null
    int row_layout;
};

// returns kept count (uniform over the warp)
__device__ inline int warp_nms_small(WarpSmem& S, const NmsProblem& pb, const RowSink& sink) {
    const int lane = threadIdx.x & 31;
    const int m = pb.n;                                   // caller guarantees n <= min(top_k, kWarpCap)
    const int Wm = (m + 31) >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // ---- 1. sort: runs of 32 in registers, then merge by rank -------------------------------
    unsigned long long* runs = reinterpret_cast<unsigned long long*>(S.x1);   // 256 x 8 B = x1 + y1
    for (int blk = 0; blk < Wm; ++blk) {
        const int e = blk * 32 + lane;
        unsigned long long k = e < m ? pb.cand[e] : 0ull;
        k = warp_sort32_desc(k, lane);
        if (Wm == 1) S.keys[lane] = k; else runs[e] = k;
    }
    for (int i = lane; i < 4 * kCols * kWarpWS; i += 32) S.tab[i] = 0;
    if (lane == 0) S.sup = 0;
    S.tin[lane] = 0;
    __syncwarp();
    if (Wm > 1) {
        for (int blk = 0; blk < Wm; ++blk) {
            const unsigned long long k = runs[blk * 32 + lane];
            int pos = lane;
            for (int ob = 0; ob < Wm; ++ob) {
                if (ob == blk) continue;
                const unsigned long long* r = runs + ob * 32;
                int lo = 0;                         // number of keys of run `ob` greater than k
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (r[lo + step - 1] > k) lo += step;
                if (lo == 31 && r[31] > k) lo = 32;
                pos += lo;
            }
            if (k != 0ull) S.keys[pos] = k;         // zero = padding, sorts last
        }
        __syncwarp();
    }

    // ---- 2. boxes, bin tables ----------------------------------------------------------------
    const bool pixel = (pb.flags & RD_NMS_PIXEL_PLUS1) != 0;
    uint32_t mnx = 0xffffffffu, mxx = 0, mny = 0xffffffffu, mxy = 0;
    float bx1[kWarpW], by1[kWarpW], bx2[kWarpW], by2[kWarpW], bar[kWarpW];
#pragma unroll
    for (int ib = 0; ib < kWarpW; ++ib) {
        const int j = ib * 32 + lane;
        bx1[ib] = by1[ib] = bx2[ib] = by2[ib] = bar[ib] = 0.f;
        if (ib < Wm && j < m) {
            float4 b = pb.boxes[key_index(S.keys[j])];
            if (pb.has_scale) { b.x *= pb.scale.x; b.y *= pb.scale.y; b.z *= pb.scale.z; b.w *= pb.scale.w; }
            bx1[ib] = b.x; by1[ib] = b.y; bx2[ib] = b.z; by2[ib] = b.w;
            bar[ib] = pixel ? (b.z - b.x + 1.0f) * (b.w - b.y + 1.0f) : (b.z - b.x) * (b.w - b.y);
            float cx = 0.5f * b.x + 0.5f * b.z, cy = 0.5f * b.y + 0.5f * b.w;
            if (isfinite(cx)) { mnx = min(mnx, float_to_ordered(cx)); mxx = max(mxx, float_to_ordered(cx)); }
            if (isfinite(cy)) { mny = min(mny, float_to_ordered(cy)); mxy = max(mxy, float_to_ordered(cy)); }
        }
    }
    __syncwarp();                                   // all lanes are done reading `runs`
#pragma unroll
    for (int ib = 0; ib < kWarpW; ++ib) {
        const int j = ib * 32 + lane;
        if (ib < Wm && j < m) { S.x1[j] = bx1[ib]; S.y1[j] = by1[ib]; S.x2[j] = bx2[ib]; S.y2[j] = by2[ib]; S.area[j] = bar[ib]; }
    }
    mnx = __reduce_min_sync(kFullMask, mnx); mxx = __reduce_max_sync(kFullMask, mxx);
    mny = __reduce_min_sync(kFullMask, mny); mxy = __reduce_max_sync(kFullMask, mxy);
    float lox = 0.f, invx = 0.f, loy = 0.f, invy = 0.f;
    if (mnx <= mxx) {
        lox = ordered_to_float(mnx);
        float hi = ordered_to_float(mxx);
        invx = hi > lox ? (float)kCols / (hi - lox) : 0.f;
    }
    if (mny <= mxy) {
        loy = ordered_to_float(mny);
        float hi = ordered_to_float(mxy);
        invy = hi > loy ? (float)kCols / (hi - loy) : 0.f;
    }
    if (!isfinite(invx)) invx = 0.f;
    if (!isfinite(invy)) invy = 0.f;
    const bool eq = (pb.flags & RD_NMS_SUPPRESS_EQ) != 0;
    const bool force_full = eq ? !(pb.thr > 0.0f) : !(pb.thr >= 0.0f);
    const float eps = 9.5367431640625e-07f;
    uint32_t crr[kWarpW];
#pragma unroll
    for (int ib = 0; ib < kWarpW; ++ib) {
        const int i = ib * 32 + lane;
        crr[ib] = 0;
        if (ib < Wm && i < m) {
            const float x1 = bx1[ib], y1 = by1[ib], x2 = bx2[ib], y2 = by2[ib], ar = bar[ib];
            const float hx = pixel ? x2 + 1.0f : x2, hy = pixel ? y2 + 1.0f : y2;
            int ax, bx, ay, by;
            const bool ok = isfinite(x1) && isfinite(y1) && isfinite(hx) && isfinite(hy) && (ar > 0.0f) &&
                            isfinite(ar) && !force_full;
            if (ok) {
                ax = col_of(x1 - fabsf(x1) * eps, lox, invx);
                bx = col_of(hx + fabsf(hx) * eps, lox, invx);
                ay = col_of(y1 - fabsf(y1) * eps, loy, invy);
                by = col_of(hy + fabsf(hy) * eps, loy, invy);
            } else {
                ax = 0; bx = kCols - 1; ay = 0; by = kCols - 1;
            }
            crr[ib] = (uint32_t)ax | ((uint32_t)bx << 8) | ((uint32_t)ay << 16) | ((uint32_t)by << 24);
            const uint32_t bit = 1u << lane;
            atomicOr(&S.tab[(0 * kCols + ax) * kWarpWS + ib], bit);
            if (bx + 1 < kCols) atomicOr(&S.tab[(1 * kCols + bx + 1) * kWarpWS + ib], bit);
            atomicOr(&S.tab[(2 * kCols + ay) * kWarpWS + ib], bit);
            if (by + 1 < kCols) atomicOr(&S.tab[(3 * kCols + by + 1) * kWarpWS + ib], bit);
        }
    }
    __syncwarp();
    for (int task = lane; task < 4 * Wm; task += 32) {
        int t = task / Wm, w = task - t * Wm;
        uint32_t acc = 0;
        uint32_t* p = S.tab + t * kCols * kWarpWS + w;
#pragma unroll 8
        for (int c = 0; c < kCols; ++c) { acc |= p[c * kWarpWS]; p[c * kWarpWS] = acc; }
    }
    __syncwarp();

    // ---- 3. walk -----------------------------------------------------------------------------
    const float thr = pb.thr;
    const int flags = pb.flags;
    const int max_out = pb.max_out;
    int kept_total = 0;
    const uint32_t* Sx = S.tab;
    const uint32_t* Ex = S.tab + 1 * kCols * kWarpWS;
    const uint32_t* Sy = S.tab + 2 * kCols * kWarpWS;
    const uint32_t* Ey = S.tab + 3 * kCols * kWarpWS;
#pragma unroll
    for (int ib = 0; ib < kWarpW; ++ib) {
        if (ib >= Wm || kept_total >= max_out) break;
        const int j = ib * 32 + lane;
        const bool valid = j < m;
        bool alive = valid;
        const float x1 = bx1[ib], y1 = by1[ib], x2 = bx2[ib], y2 = by2[ib], ar = bar[ib];
        const uint32_t cr = crr[ib];
        const uint32_t* rSx = Sx + ((cr >> 8) & 255u) * kWarpWS;
        const uint32_t* rEx = Ex + (cr & 255u) * kWarpWS;
        const uint32_t* rSy = Sy + ((cr >> 24) & 255u) * kWarpWS;
        const uint32_t* rEy = Ey + ((cr >> 16) & 255u) * kWarpWS;
        // candidate pairs after the bin cull: kept boxes of earlier blocks, earlier lanes of this block
        uint32_t h[kWarpW];
        int nh = 0;
#pragma unroll
        for (int w = 0; w < kWarpW; ++w) {
            h[w] = 0;
            if (w <= ib && valid) {
                h[w] = rSx[w] & ~rEx[w] & rSy[w] & ~rEy[w] & (w < ib ? S.keptbits[w] : lt_mask);
                nh += __popc(h[w]);
            }
        }
        int off = nh;                                // inclusive scan over lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(kFullMask, off, d); if (lane >= d) off += o; }
        const int total = __shfl_sync(kFullMask, off, 31);
        off -= nh;
        uint32_t tin = 0;
        if (total > 0 && total <= kPairCap) {
#pragma unroll
            for (int w = 0; w < kWarpW; ++w) {
                uint32_t hw = h[w];
                while (hw) {
                    const int i = (w << 5) + __ffs(hw) - 1;
                    hw &= hw - 1;
                    S.pairs[off++] = ((uint32_t)lane << 16) | (uint32_t)i;
                }
            }
            __syncwarp();
            for (int p = lane; p < total; p += 32) {
                const uint32_t e = S.pairs[p];
                const int jl = (int)(e >> 16), i = (int)(e & 0xffffu);
                const int jj = ib * 32 + jl;
                if (suppresses(S.x1[i], S.y1[i], S.x2[i], S.y2[i], S.area[i], S.x1[jj], S.y1[jj], S.x2[jj], S.y2[jj],
                               S.area[jj], thr, flags)) {
                    if (i < ib * 32) atomicOr(&S.sup, 1u << jl);
                    else atomicOr(&S.tin[jl], 1u << (i - ib * 32));
                }
            }
            __syncwarp();
            if ((S.sup >> lane) & 1u) alive = false;
            tin = S.tin[lane];
            __syncwarp();
            S.tin[lane] = 0;
            if (lane == 0) S.sup = 0;
        } else if (total > 0) {                      // pair list would overflow: test in place
#pragma unroll
            for (int w = 0; w < kWarpW; ++w) {
                uint32_t hw = h[w];
                while (hw && (alive || w == ib)) {
                    const int i = (w << 5) + __ffs(hw) - 1;
                    hw &= hw - 1;
                    if (suppresses(S.x1[i], S.y1[i], S.x2[i], S.y2[i], S.area[i], x1, y1, x2, y2, ar, thr, flags)) {
                        if (w < ib) alive = false; else tin |= 1u << (i - ib * 32);
                    }
                }
            }
        }
        // in-block resolution in score order
        uint32_t u = __reduce_or_sync(kFullMask, alive ? tin : 0u);
        while (u) {
            const int k = __ffs(u) - 1;
            u &= u - 1;
            const uint32_t al = __ballot_sync(kFullMask, alive);
            if (((al >> k) & 1u) && ((tin >> k) & 1u)) alive = false;
        }
        uint32_t keptw = __ballot_sync(kFullMask, alive);
        const int room = max_out - kept_total;
        int cnt = __popc(keptw);
        if (cnt > room) {
            uint32_t t = keptw, keep = 0;
            for (int r = 0; r < room; ++r) { uint32_t low = t & (0u - t); keep |= low; t ^= low; }
            keptw = keep;
            cnt = room;
        }
        if ((keptw >> lane) & 1u) {
            const int t = kept_total + __popc(keptw & lt_mask);
            const unsigned long long key = S.keys[j];
            const float sc = key_score(key);
            float* r = sink.rows + (size_t)t * 5;
            if (sink.row_layout == RD_ROW_SCORE_BOX) { r[0] = sc; r[1] = x1; r[2] = y1; r[3] = x2; r[4] = y2; }
            else { r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = sc; }
            if (sink.anchors) sink.anchors[t] = (int)key_index(key);
        }
        if (lane == 0) S.keptbits[ib] = keptw;
        __syncwarp();
        kept_total += cnt;
    }
    return kept_total;
}

}  // namespace rd

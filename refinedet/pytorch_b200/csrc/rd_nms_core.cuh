// rd_nms_core.cuh — one NMS problem per CTA: top-k select, sort, spatially-culled exact greedy
// NMS.  Shared by the fused detect stage (one problem per (image, class)) and the stand-alone
// rd_nms / rd_nms_host entry points.
//
// Reference semantics reproduced:
//   layers/box_utils.py:222-286 (normalised flavour), utils/nms/py_cpu_nms.py:10-38 ==
//   utils/nms/nms_kernel.cu:24-32,124-140 (pixel +1 flavour), utils/nms/cpu_nms.pyx:65
//   (suppress-on-equal variant), eval_refinedet_coco.py:222,231 (top_k, per-class cap).
//
// Algorithm (B200-first, not the reference's dense n x n bitmask + host scan):
//   keys = (ordered score bits << 32) | ~index, so one unsigned compare gives "score descending,
//   lower index first".  Boxes are binned into 32 columns x 32 rows of the problem's own extent;
//   prefix-OR tables  S[c] = {boxes starting at or before bin c},  E[c] = {boxes ending before
//   bin c}  turn "which earlier boxes can intersect box j at all" into word-wide ANDs
//   (X_j = S[b_j] & ~E[a_j]).  Only those pairs get the exact fp32 IoU test, so the work is
//   O(n^2/32) word operations + O(#intersecting pairs) IoUs.  The cull is conservative (monotone
//   binning), so the kept set is exactly the greedy result.  The walk handles 32 candidates per
//   step; the surviving (candidate, kept earlier box) pairs are flattened into a list and tested
//   by all threads, in-block dependencies are resolved with ballots.
//
//   small problems (n <= 256, no select) of images whose suppression graph exists: cta_sort_small
//     (sorted runs of 32 in registers, bitonic over shuffles, merged by rank) + cta_nms_graph (direct
//     node -> rank table, graph look-ups, dependency rounds)
//   everything else: nms_process — radix select over the L2-resident key list when n > top_k,
//     shared-memory bitonic sort, per-problem bin tables, walk with the tests done in place
#pragma once
#include <type_traits>

#include "rd_common.cuh"

namespace rd {

constexpr int kCols = 32;  // spatial bins per axis

// Candidate keys of one problem: n keys at `base`, any order.
struct CandList {
    const unsigned long long* base;
    int n;
};

struct NmsProblem {
    CandList cl;
    const float4* boxes;             // boxes[key index] (point form)
    float4 scale;                    // multiplied into the boxes when has_scale
    int has_scale;
    float thr;
    int top_k;
    int max_out;
    int flags;
    // alternative source (rd_nms_host): n rows [x1,y1,x2,y2,score,...] of dets_dim floats, ALREADY in
    // score-descending order (the contract of _nms, utils/nms/gpu_nms.pyx:26-29), 16-byte aligned; when
    // set, cl/boxes are unused, nothing is selected or sorted and the key index of a row is its position
    const float* dets;
    int dets_dim;
};

struct RowSink {           // where rows are written
    float* rows;           // [max_out,5] slot, or null
    int* anchors;          // [max_out] key index per row, or null
    long long* keep64;     // stand-alone: kept key indices, or null
    int* keep32;
    int row_layout;
    const int* idx_map;    // anchors[t] = idx_map[key index] when set (fused stage: key index = node)
};

__device__ __forceinline__ void sink_emit(const RowSink& sink, int t, unsigned long long key, float x1, float y1,
                                          float x2, float y2) {
    const uint32_t idx = key_index(key);
    if (sink.rows) {
        const float sc = key_score(key);
        float* r = sink.rows + (size_t)t * 5;
        if (sink.row_layout == RD_ROW_SCORE_BOX) { r[0] = sc; r[1] = x1; r[2] = y1; r[3] = x2; r[4] = y2; }
        else { r[0] = x1; r[1] = y1; r[2] = x2; r[3] = y2; r[4] = sc; }
    }
    if (sink.anchors) sink.anchors[t] = sink.idx_map ? sink.idx_map[idx] : (int)idx;
    if (sink.keep64) sink.keep64[t] = (long long)idx;
    if (sink.keep32) sink.keep32[t] = (int)idx;
}

// exact suppression test: does kept box i (higher score) suppress candidate j ?
__device__ __forceinline__ bool suppresses(float x1i, float y1i, float x2i, float y2i,
                                           float x1j, float y1j, float x2j, float y2j, float thr, int flags) {
    float iou;
    if (flags & RD_NMS_PIXEL_PLUS1) {
        // py_cpu_nms.py:18-33 / nms_kernel.cu:24-32
        const float ai = (x2i - x1i + 1.0f) * (y2i - y1i + 1.0f);
        const float aj = (x2j - x1j + 1.0f) * (y2j - y1j + 1.0f);
        const float w = fmaxf(0.0f, fminf(x2i, x2j) - fmaxf(x1i, x1j) + 1.0f);
        const float h = fmaxf(0.0f, fminf(y2i, y2j) - fmaxf(y1i, y1j) + 1.0f);
        const float inter = w * h;
        iou = inter / (ai + aj - inter);
    } else {
        // box_utils.py:241,268-283: union = (rem_areas - inter) + area[i]
        const float ai = (x2i - x1i) * (y2i - y1i);
        const float aj = (x2j - x1j) * (y2j - y1j);
        const float w = fmaxf(fminf(x2j, x2i) - fmaxf(x1j, x1i), 0.0f);
        const float h = fmaxf(fminf(y2j, y2i) - fmaxf(y1j, y1i), 0.0f);
        const float inter = w * h;
        iou = inter / ((aj - inter) + ai);
    }
    return (flags & RD_NMS_SUPPRESS_EQ) ? (iou >= thr) : !(iou <= thr);
}

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2, bool pixel) {
    return pixel ? (x2 - x1 + 1.0f) * (y2 - y1 + 1.0f) : (x2 - x1) * (y2 - y1);
}

__device__ __forceinline__ int col_of(float v, float lo, float inv) {
    float f = floorf((v - lo) * inv);
    f = fminf(fmaxf(f, 0.0f), (float)(kCols - 1));
    return (int)f;   // NaN -> fmaxf/fminf drop it -> 0
}

// bin range of one box: (ax | bx<<8 | ay<<16 | by<<24).  Boxes the cull cannot reason about
// (non-finite, non-positive area) and thresholds for which "no intersection => not suppressed"
// does not hold get the full range, i.e. are always tested exactly.
__device__ __forceinline__ uint32_t bin_range(float x1, float y1, float x2, float y2, bool pixel, bool force_full,
                                              float lox, float invx, float loy, float invy) {
    const float hx = pixel ? x2 + 1.0f : x2, hy = pixel ? y2 + 1.0f : y2;
    const float ar = box_area(x1, y1, x2, y2, pixel);
    const bool ok = isfinite(x1) && isfinite(y1) && isfinite(hx) && isfinite(hy) && (ar > 0.0f) && isfinite(ar) &&
                    !force_full;
    int ax = 0, bx = kCols - 1, ay = 0, by = kCols - 1;
    if (ok) {
        const float eps = 9.5367431640625e-07f;   // 2^-20 relative nudge keeps the binning conservative
        ax = col_of(x1 - fabsf(x1) * eps, lox, invx);
        bx = col_of(hx + fabsf(hx) * eps, lox, invx);
        ay = col_of(y1 - fabsf(y1) * eps, loy, invy);
        by = col_of(hy + fabsf(hy) * eps, loy, invy);
    }
    return (uint32_t)ax | ((uint32_t)bx << 8) | ((uint32_t)ay << 16) | ((uint32_t)by << 24);
}

__device__ __forceinline__ void extent_to_scale(uint32_t mn, uint32_t mx, float& lo, float& inv) {
    lo = 0.f; inv = 0.f;
    if (mn <= mx) {
        lo = ordered_to_float(mn);
        const float hi = ordered_to_float(mx);
        inv = hi > lo ? (float)kCols / (hi - lo) : 0.f;
        if (!isfinite(inv)) inv = 0.f;
    }
}

__device__ __forceinline__ bool cull_disabled(float thr, int flags) {
    return (flags & RD_NMS_SUPPRESS_EQ) ? !(thr > 0.0f) : !(thr >= 0.0f);
}

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

// =========================================================================================
// small problems: n <= kSmallCap candidates, no select (nms_small_kernel)
// =========================================================================================
#ifndef RD_SMALL_THREADS
#define RD_SMALL_THREADS 128
#endif
constexpr int kSmallThreads = RD_SMALL_THREADS;   // common variant: <= 256 candidates, 128 threads, 6.7 KB smem
constexpr int kSmallCap = 256;
constexpr int kWideThreads = 256;                 // wide variant (few problems, e.g. C = 2): <= 1024 candidates
constexpr int kWideCap = 1024;
constexpr int kMidCap = 512;                      // nms_large_kernel's mid path: images with 1025 .. 4096 nodes, <= 512 candidates

#ifndef RD_GRAPH_NODES
#define RD_GRAPH_NODES 4096
#endif
constexpr int kGraphNodes = RD_GRAPH_NODES;   // images with more ARM-passing anchors have no suppression graph
constexpr int kAdjDeg = 8;            // adjacency slots in a node's primary row (one 16-byte load)
constexpr int kAdjDeg2 = 56;          // further slots in the overflow row: real detector output puts dozens of
                                      // mutually overlapping boxes on every object
constexpr int kAdjMax = kAdjDeg + kAdjDeg2;   // an image with a node of higher degree has no graph (flag bit 0)
constexpr int kFlagNoGraph = 1;       // img_flag bits: no usable graph / some node has more than kAdjDeg suppressors
constexpr int kFlagWideDeg = 2;

// Dependencies of a problem's candidates: lists of at most kAdjDeg ranks per candidate (every node of the image
// has at most kAdjDeg suppressors: the common case, one 16-byte adjacency row per candidate).  For images flagged
// kFlagWideDeg the kCap <= 256 instance (kBitRows) switches to one bit row per candidate over the EARLIER ranks
// -- candidate r of 32-block k = r >> 5 owns k + 1 words at tri_row(r), 4.6 KB in all -- so the degree of a node
// does not matter; wider instances leave those images to the bin path.
__device__ __forceinline__ int tri_row(int r) {
    const int k = r >> 5;
    return 16 * k * (k + 1) + (r & 31) * (k + 1);
}
// kNodes = largest node index + 1 the rank table covers (images with more nodes are not handled by this instance)
#ifndef RD_SMALL_NODES
#define RD_SMALL_NODES 1024     // 2048 keeps images of 1025 - 2048 nodes in this kernel (1185 nodes: 0.100 -> 0.092 ms) but costs the headline 3 % (28.3 -> 29.1 us: the two-pass scan becomes live code, 32 bytes of spills)
#endif
template <int kCap, int kNodes = (kCap > 256 ? RD_GRAPH_NODES : RD_SMALL_NODES)>
struct SmallSmem {
    static constexpr bool kBitRows = kCap <= 256; // false: images flagged kFlagWideDeg are not handled by this instance
    static constexpr int kDeps = kAdjDeg;         // list mode: ranks kept per candidate
    static constexpr int kTriWords = 16 * (kCap / 32) * (kCap / 32 + 1);
    using rank_t = typename std::conditional<(kCap > 256), unsigned short, unsigned char>::type;
    unsigned long long keys[kCap];                // sorted keys
    union {
        unsigned long long runs[kCap];            // unsorted candidates, then sorted runs of 32 (during the sort)
        struct {
            unsigned short rank[kNodes];          // node -> rank in this problem, 0xffff = not a candidate
            union {
                rank_t deps[kCap * kDeps];                     // list mode: ranks of the dependencies of every candidate
                unsigned int bits[kBitRows ? kTriWords : 1];   // bit-row mode (images flagged kFlagWideDeg)
                unsigned int hist[kCap + 1];                   // during the sort: bucket counts, then bucket starts
            };
        } g;
    } u;
    unsigned char depn[kCap];
    unsigned char state[kCap];                    // 0 undecided, 1 kept, 2 suppressed
    unsigned int keptw[kCap / 32], deadw[kCap / 32];
    int wsum[32];
    int n;
    static constexpr int kMaxNodes = kNodes;
};

// Sort of a small problem: S.u.runs[0..m) holds the candidate keys in any order; on return (after a CTA barrier)
// S.keys[0..m) holds them in descending order.
//   m <= 32: one warp, bitonic over shuffles.
//   otherwise a bucket sort with an exact in-bucket rank.  The bucket of a key is a monotone function of its score
//   word -- (ordered score - ordered threshold) >> shift, kCap buckets linear in the float's bit pattern (= 32 per
//   octave at threshold 0.01, scores up to 1: detection scores are spread roughly log-uniformly over (threshold, 1]) --
//   so the order BETWEEN buckets is the key order; one shared-memory atomic per key counts the buckets and hands out
//   a slot, a scan turns counts into starts, keys are regrouped by bucket, and every key ranks itself among the keys
//   of its own bucket (1 - 3 of them, typically) with the full 64-bit compare -- score descending, lower node first,
//   whatever the order the atomics were served in.  Any input is sorted correctly (everything in one bucket is an
//   O(m^2 / threads) rank); the bucket function only has to be monotone.  ~4 x fewer instructions than sorted runs of
//   32 merged by binary-search rank, which this replaces (round 1: 44 % of nms_small_kernel's instructions).
template <int kThreads, int kCap>
__device__ __forceinline__ void cta_sort_small(SmallSmem<kCap>& S, int m, float thresh) {
    constexpr int kPerT = (kCap + kThreads - 1) / kThreads;
    constexpr int kBuckets = kCap;                       // power of two
    constexpr int kPerB = (kBuckets + kThreads - 1) / kThreads;
    constexpr int kWarps = kThreads / 32;
    static_assert((kBuckets & (kBuckets - 1)) == 0, "bucket count must be a power of two");
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    unsigned long long* runs = S.u.runs;
    if (m <= 32) {
        if (warp == 0) {
            unsigned long long k = lane < m ? runs[lane] : 0ull;
            k = warp_sort32_desc(k, lane);
            S.keys[lane] = k;
        }
        __syncthreads();
        return;
    }
    unsigned int* hist = S.u.g.hist;                     // does not alias `runs` (it lies behind the rank table)
    for (int i = tid; i < kBuckets; i += kThreads) hist[i] = 0;
    // bucket = min((ordered(score) - ordered(thresh)) >> shift, kBuckets - 1), reversed (bucket 0 = highest scores)
    const uint32_t base = float_to_ordered(thresh);
    const uint32_t range = float_to_ordered(thresh < 1.0f ? 1.0f : INFINITY) - base;
    int shift = (32 - __clz((int)range)) - (31 - __clz(kBuckets));
    if (shift < 0) shift = 0;
    __syncthreads();
    unsigned long long key[kPerT];
    int bk[kPerT], slot[kPerT];
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        const int e = q * kThreads + tid;
        key[q] = 0ull; bk[q] = 0; slot[q] = 0;
        if (e < m) {
            key[q] = runs[e];
            const uint32_t d = ((uint32_t)(key[q] >> 32) - base) >> shift;
            bk[q] = kBuckets - 1 - (int)min(d, (uint32_t)(kBuckets - 1));
            slot[q] = (int)atomicAdd(&hist[bk[q]], 1u);
        }
    }
    __syncthreads();
    // exclusive scan of the bucket counts, in place: thread t owns buckets [t * kPerB, (t + 1) * kPerB)
    {
        unsigned int c[kPerB];
        unsigned int sum = 0;
#pragma unroll
        for (int i = 0; i < kPerB; ++i) {
            const int b = tid * kPerB + i;
            c[i] = b < kBuckets ? hist[b] : 0u;
            sum += c[i];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int o = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) S.wsum[warp] = (int)incl;
        __syncthreads();
        unsigned int run = incl - sum;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) if (w < warp) run += (unsigned int)S.wsum[w];
#pragma unroll
        for (int i = 0; i < kPerB; ++i) {
            const int b = tid * kPerB + i;
            if (b < kBuckets) hist[b] = run;
            run += c[i];
        }
        if (tid == kThreads - 1) hist[kBuckets] = run;                 // = m
    }
    __syncthreads();
    // regroup by bucket (the unsorted input is dead: every key is in a register)
    int lo[kPerT], hi[kPerT];
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        lo[q] = 0; hi[q] = 0;
        if (q * kThreads + tid < m) {
            lo[q] = (int)hist[bk[q]];
            hi[q] = (int)hist[bk[q] + 1];
            runs[lo[q] + slot[q]] = key[q];
        }
    }
    __syncthreads();
    // exact rank inside the bucket
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        if (q * kThreads + tid < m) {
            int pos = lo[q];
            for (int i = lo[q]; i < hi[q]; ++i) pos += runs[i] > key[q] ? 1 : 0;
            S.keys[pos] = key[q];
        }
    }
    __syncthreads();
}

// =========================================================================================
// small problems, graph mode: the suppression relation between the ARM-passing anchors (nodes) of an
// image does not depend on the class, so it is computed once per image (graph_kernel in
// rd_detect.cu) as adjacency lists  adj[node] = {u : box u suppresses box node when u is kept}.
// A problem then only sorts its keys, marks the rank of every candidate node in a direct table and
// resolves the dependencies; no boxes, bins or IoUs per class.
// =========================================================================================

struct GraphView {
    const uint4* adj;                 // [kGraphNodes] node indices of the suppressors, 8 x u16 per node
    const unsigned short* adj2;       // [kGraphNodes][kAdjDeg2] suppressors 8 .. 31 of the nodes that have them
    const int* adjn;                  // [kGraphNodes] degree (<= kAdjMax when the image is not flagged kFlagNoGraph)
    const float4* nbox;               // [N] node boxes, already multiplied by the image scale
    const int* nanc;                  // [N] anchor index of every node
};

// S.keys[0..m) = the problem's sorted keys (key index = node < N).  One round of global loads: adjacency row,
// box and anchor of every candidate.
template <int kThreads, int kCap>
__device__ inline int cta_nms_graph(SmallSmem<kCap>& S, int m, int N, int max_out, const RowSink& sink,
                                    const GraphView& G, bool wide_deg = false) {
    using rank_t = typename SmallSmem<kCap>::rank_t;
    constexpr int kDeps = SmallSmem<kCap>::kDeps;
    constexpr int kSmallWarps = kThreads / 32;
    constexpr int kPerT = (kCap + kThreads - 1) / kThreads;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // candidate r = q * kThreads + tid (striped); everything about it stays in registers
    unsigned long long key[kPerT];
    float4 box[kPerT];
    uint4 row[kPerT];
    int dn[kPerT], anc[kPerT];
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        const int r = q * kThreads + tid;
        key[q] = r < m ? S.keys[r] : 0ull;
        dn[q] = 0;
        if (r < m) {
            const uint32_t u = key_index(key[q]);
            dn[q] = G.adjn[u];
            row[q] = __ldg(G.adj + u);
            box[q] = G.nbox[u];                              // most candidates are kept: fetch the row data now
            anc[q] = G.nanc[u];
        }
    }
    // rank table (aliases the sort buffer: every thread is past the merge, see the barrier in cta_sort_small)
    {
        uint4* t = reinterpret_cast<uint4*>(S.u.g.rank);
        constexpr int kMaxNodes = SmallSmem<kCap>::kMaxNodes;
        const int nn = kMaxNodes <= 1024 ? kMaxNodes : N;         // small table: constant bound (one store per thread)
        for (int i = tid; i < (nn * 2 + 15) / 16; i += kThreads) t[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        const int r = q * kThreads + tid;
        if (r < m) S.u.g.rank[key_index(key[q])] = (unsigned short)r;
    }
    __syncthreads();
    const bool bit_rows = SmallSmem<kCap>::kBitRows && wide_deg;       // CTA-uniform
    if (bit_rows) {
        // ---- bit rows: bit v of row r = candidate of rank v < r suppresses candidate r when kept ----------
        constexpr int kW = kCap / 32;
        uint32_t* bm = S.u.g.bits;
        unsigned openm = 0;                                          // bit q: candidate q of this thread is undecided
#pragma unroll
        for (int q = 0; q < kPerT; ++q) {
            const int r = q * kThreads + tid;
            if (r < m) {
                uint32_t* drow = bm + tri_row(r);                    // only this thread touches row r here; a row
                // (a row exists once it has a bit)
                const unsigned long long lo = ((unsigned long long)row[q].y << 32) | row[q].x;
                const unsigned long long hi = ((unsigned long long)row[q].w << 32) | row[q].z;
                bool any = false;
#pragma unroll
                for (int k = 0; k < kAdjDeg; ++k) {
                    if (k < dn[q]) {
                        const int rv = S.u.g.rank[(uint32_t)((k < 4 ? lo : hi) >> ((k & 3) * 16)) & 0xffffu];
                        if (rv < r) {
                            if (!any) {
#pragma unroll
                                for (int w = 0; w < kW; ++w) if (w <= (r >> 5)) drow[w] = 0;
                                any = true;
                            }
                            drow[rv >> 5] |= 1u << (rv & 31);
                        }
                    }
                }
                if (dn[q] > kAdjDeg) {                               // overflow row (L2): suppressors 8 .. dn - 1
                    const unsigned short* r2 = G.adj2 + (size_t)key_index(key[q]) * kAdjDeg2;
                    for (int k = 0; k < dn[q] - kAdjDeg; ++k) {
                        const int rv = S.u.g.rank[__ldg(r2 + k)];
                        if (rv < r) {
                            if (!any) {
#pragma unroll
                                for (int w = 0; w < kW; ++w) if (w <= (r >> 5)) drow[w] = 0;
                                any = true;
                            }
                            drow[rv >> 5] |= 1u << (rv & 31);
                        }
                    }
                }
                if (any) openm |= 1u << q;
            }
        }
#pragma unroll
        for (int q = 0; q < kPerT; ++q) {
            const int r = q * kThreads + tid;
            const unsigned free_ = __ballot_sync(kFullMask, r < m && !((openm >> q) & 1u));    // no earlier suppressor: kept
            if (lane == 0 && (r >> 5) < kW) { S.keptw[r >> 5] = free_; S.deadw[r >> 5] = 0; }   // the warp's own word
        }
        __syncthreads();
        // ---- resolve in rounds: a candidate is dead once a dependency is kept, kept once all are dead ------
        for (int round = 0; round < kCap; ++round) {
            int undecided = 0;
#pragma unroll
            for (int q = 0; q < kPerT; ++q) {
                const int r = q * kThreads + tid;
                if (!((openm >> q) & 1u)) continue;
                uint32_t hit = 0, live = 0;
                const uint32_t* drow = bm + tri_row(r);
#pragma unroll
                for (int w = 0; w < kW; ++w) {
                    if (w <= (r >> 5)) {
                        const uint32_t d = drow[w];
                        hit |= d & S.keptw[w];
                        live |= d & ~S.deadw[w];
                    }
                }
                if (hit) { atomicOr(&S.deadw[r >> 5], 1u << (r & 31)); openm &= ~(1u << q); }
                else if (!live) { atomicOr(&S.keptw[r >> 5], 1u << (r & 31)); openm &= ~(1u << q); }
                else undecided = 1;
            }
            if (!__syncthreads_or(undecided)) break;
        }
    } else {
    // ---- dependencies: graph neighbours that are candidates of this class and rank earlier --------
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        const int r = q * kThreads + tid;
        if (r < m) {
            const uint32_t nb[4] = {row[q].x, row[q].y, row[q].z, row[q].w};
            int nd = 0;
#pragma unroll
            for (int k = 0; k < kAdjDeg; ++k) {
                if (k < dn[q]) {
                    const uint32_t v = (nb[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                    const int rv = S.u.g.rank[v];
                    if (rv < r) S.u.g.deps[r * kDeps + nd++] = (rank_t)rv;
                }
            }
            S.depn[r] = (unsigned char)nd;
            S.state[r] = nd == 0 ? 1 : 0;
        }
    }
    __syncthreads();
    // ---- resolve in rounds (dependencies always point to earlier ranks: terminates) ---------------
    for (int round = 0; round < kCap; ++round) {
        int undecided = 0;
#pragma unroll
        for (int q = 0; q < kPerT; ++q) {
            const int r = q * kThreads + tid;
            if (r >= m || S.state[r] != 0) continue;
            const int nd = S.depn[r];
            bool any_kept = false, all_sup = true;
            for (int k = 0; k < nd; ++k) {
                const unsigned char st = S.state[S.u.g.deps[r * kDeps + k]];
                any_kept |= (st == 1);
                all_sup &= (st == 2);
            }
            if (any_kept) S.state[r] = 2;
            else if (all_sup) S.state[r] = 1;
            else undecided = 1;
        }
        if (!__syncthreads_or(undecided)) break;
    }
    }
    // ---- emit kept rows in rank order, first max_out: scan of the kept flags chunk by chunk -----------
    int carry = 0;
#pragma unroll
    for (int q = 0; q < kPerT; ++q) {
        const int r = q * kThreads + tid;
        const bool kept = r < m && (bit_rows ? ((S.keptw[(r >> 5) % (kCap / 32)] >> (r & 31)) & 1u) != 0 : S.state[r] == 1);
        const unsigned bal = __ballot_sync(kFullMask, kept);
        if (lane == 0) S.wsum[warp] = __popc(bal);
        __syncthreads();
        int base = carry + __popc(bal & ((1u << lane) - 1u));
        int chunk = 0;
#pragma unroll
        for (int w = 0; w < kSmallWarps; ++w) {
            if (w < warp) base += S.wsum[w];
            chunk += S.wsum[w];
        }
        if (kept && base < max_out) {
            const float sc = key_score(key[q]);
            float* o = sink.rows + (size_t)base * 5;
            const float4 bx = box[q];
            if (sink.row_layout == RD_ROW_SCORE_BOX) { o[0] = sc; o[1] = bx.x; o[2] = bx.y; o[3] = bx.z; o[4] = bx.w; }
            else { o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w; o[4] = sc; }
            if (sink.anchors) sink.anchors[base] = anc[q];
        }
        carry += chunk;
        __syncthreads();
    }
    return carry < max_out ? carry : max_out;
}

// =========================================================================================
// large problems: one CTA (any multiple of 32 threads), n arbitrary, m = min(n, top_k) <= mcap
// =========================================================================================
constexpr int kLargePairCap = 2048;        // pair-list words of a CTA, split evenly between its warps
constexpr int kChunkBlocks = 8;            // the greedy walk settles up to 8 x 32 candidates per CTA-wide step
constexpr int kTinStride = kChunkBlocks + 1;   // odd row stride of the in-chunk suppressor masks

struct NmsSmemLayout {
    int mcap;   // max boxes held (multiple of 32)
    int W;      // mcap / 32
    int WS;     // padded row stride of the bin tables (odd -> conflict-free)
    int Kp;     // power of two >= mcap (bitonic sort buffer)
    size_t off_keys, off_x1, off_y1, off_x2, off_y2, off_cr, off_tab, off_keptbits, off_hist, off_misc,
        off_pairs, off_tin, total;
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
    L.off_cr = o;        o += (size_t)mcap * 4;
    L.off_tab = o;       o += (size_t)4 * kCols * L.WS * 4;
    L.off_keptbits = o;  o += (size_t)L.W * 4;
    L.off_hist = o;      o += 256 * 4;
    L.off_misc = o;      o += 16 * 4;
    L.off_tin = o;       o += (size_t)(kChunkBlocks * 32 * kTinStride + kChunkBlocks) * 4;
    L.off_pairs = o;     o += (size_t)kLargePairCap * 4;
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// Descending sort of keys[0..m) (64-bit, pairwise distinct) by the whole CTA, any thread count: the bucket sort of
// cta_sort_small for the large path.  Buckets are linear in the key over [min key, max key] of the problem (the
// keys that survive a top-k select span a few octaves of score), nb of them; `scratch` (8-byte aligned) provides
// 256 bytes (reductions) + 8 m bytes (keys regrouped by bucket) + 4 (nb + 1) bytes (counts, then starts) + 2 m bytes
// (slot of every key inside its bucket).  Seven CTA barriers instead of the 55 of a 1024-key bitonic network, ~10 x fewer
// instructions; an adversarial input (everything in one bucket) degrades to an O(m^2 / threads) exact rank.
__device__ __forceinline__ void cta_bucket_sort(unsigned long long* keys, int m, unsigned char* scratch, int nb) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
    unsigned long long* red = reinterpret_cast<unsigned long long*>(scratch);          // 2 words + 32 warp totals: 256 bytes
    scratch += 256;
    unsigned long long* tmp = reinterpret_cast<unsigned long long*>(scratch);
    unsigned int* hist = reinterpret_cast<unsigned int*>(scratch + (size_t)m * 8);
    unsigned short* slots = reinterpret_cast<unsigned short*>(scratch + (size_t)m * 8 + (size_t)(nb + 1) * 4);
    if (tid == 0) { red[0] = ~0ull; red[1] = 0ull; }
    for (int i = tid; i <= nb; i += nthr) hist[i] = 0;
    __syncthreads();
    unsigned long long mn = ~0ull, mx = 0ull;
    for (int e = tid; e < m; e += nthr) {
        const unsigned long long k = keys[e];
        mn = k < mn ? k : mn;
        mx = k > mx ? k : mx;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_xor_sync(kFullMask, mn, d), c = __shfl_xor_sync(kFullMask, mx, d);
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
    }
    if (lane == 0) { atomicMin(&red[0], mn); atomicMax(&red[1], mx); }
    __syncthreads();
    const unsigned long long lo = red[0], range = red[1] - lo;
    int lg = 0;
    while ((1 << lg) < nb) ++lg;
    int shift = (64 - __clzll((long long)(range | 1ull))) - lg;        // (range >> shift) < nb
    if (shift < 0) shift = 0;
    for (int e = tid; e < m; e += nthr) {
        const int b = nb - 1 - (int)((keys[e] - lo) >> shift);         // bucket 0 = the highest keys
        slots[e] = (unsigned short)atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    // exclusive scan of hist[0..nb) in place (nb <= 1024: one pass of the first nb threads' chunks)
    {
        const int per = (nb + nthr - 1) / nthr;
        unsigned int sum = 0;
        for (int i = 0; i < per; ++i) { const int b = tid * per + i; sum += b < nb ? hist[b] : 0u; }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned int o = __shfl_up_sync(kFullMask, incl, d); if (lane >= d) incl += o; }
        unsigned int* wtot = reinterpret_cast<unsigned int*>(red) + 4;  // up to 32 warp totals behind the two words
        if (lane == 31) wtot[tid >> 5] = incl;
        __syncthreads();
        unsigned int run = incl - sum;
        for (int w = 0; w < (tid >> 5); ++w) run += wtot[w];
        for (int i = 0; i < per; ++i) {
            const int b = tid * per + i;
            if (b < nb) { const unsigned int c = hist[b]; hist[b] = run; run += c; }
        }
        if (tid == nthr - 1) hist[nb] = (unsigned int)m;
    }
    __syncthreads();
    for (int e = tid; e < m; e += nthr) {
        const unsigned long long k = keys[e];
        const int b = nb - 1 - (int)((k - lo) >> shift);
        tmp[hist[b] + slots[e]] = k;
    }
    __syncthreads();
    for (int p = tid; p < m; p += nthr) {
        const unsigned long long k = tmp[p];
        const int b = nb - 1 - (int)((k - lo) >> shift);
        const int blo = (int)hist[b], bhi = (int)hist[b + 1];
        int pos = blo;
        for (int i = blo; i < bhi; ++i) pos += tmp[i] > k ? 1 : 0;
        keys[pos] = k;
    }
    __syncthreads();
}

// Runs one problem on the calling CTA.  Rows are emitted through `sink`.  Returns the kept
// count (uniform over the CTA).
__device__ inline int nms_process(unsigned char* smem, const NmsSmemLayout& L, const NmsProblem& pb,
                                  const RowSink& sink) {
    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.off_keys);
    float* sx1 = reinterpret_cast<float*>(smem + L.off_x1);
    float* sy1 = reinterpret_cast<float*>(smem + L.off_y1);
    float* sx2 = reinterpret_cast<float*>(smem + L.off_x2);
    float* sy2 = reinterpret_cast<float*>(smem + L.off_y2);
    uint32_t* scr = reinterpret_cast<uint32_t*>(smem + L.off_cr);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + L.off_tab);
    uint32_t* keptbits = reinterpret_cast<uint32_t*>(smem + L.off_keptbits);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L.off_hist);
    uint32_t* misc = reinterpret_cast<uint32_t*>(smem + L.off_misc);
    // misc: 0 select counter, 1 digit, 2 need, 3 done, 4..7 ordered min/max, 8 kept count, 9 sup, 10 pairs, 11 stop

    const CandList& cl = pb.cl;
    const int n = cl.n;
    const int top_k = pb.top_k < L.mcap ? pb.top_k : L.mcap;
    const int m = n < top_k ? n : top_k;
    if (m <= 0) return 0;
    const bool presorted = pb.dets != nullptr;

    // ---- 1. load or select the m highest keys --------------------------------------------------
    if (presorted) {
        for (int i = tid; i < m; i += nthr)
            keys[i] = ((unsigned long long)(0xffffffffu - (uint32_t)i) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    } else if (n <= top_k) {
        for (int i = tid; i < n; i += nthr) keys[i] = cl.base[i];
    } else {
        unsigned long long prefix = 0;   // known high bits of the threshold key
        int need = top_k;                // how many keys of the current bucket are wanted
        unsigned long long thresh_key = 0;
        for (int shift = 56; shift >= 0; shift -= 8) {
            for (int i = tid; i < 256; i += nthr) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += nthr) {
                const unsigned long long k = cl.base[i];
                const bool match = (shift == 56) || ((k >> (shift + 8)) == prefix);
                if (match) atomicAdd(&hist[(unsigned)(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                // lane l owns digits [8l, 8l+8); find the digit where the count from the top crosses `need`
                uint32_t loc[8];
                uint32_t sum = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) { loc[q] = hist[lane * 8 + q]; sum += loc[q]; }
                uint32_t v = sum;           // inclusive suffix sum over lanes
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_down_sync(kFullMask, v, d);
                    if (lane + d < 32) v += o;
                }
                uint32_t cum = v - sum;     // keys in digits owned by higher lanes
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
            if (done || shift == 0) { thresh_key = prefix << shift; break; }
        }
        if (tid == 0) misc[0] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += nthr) {
            const unsigned long long k = cl.base[i];
            if (k >= thresh_key) {
                const uint32_t pos = atomicAdd(&misc[0], 1u);
                if (pos < (uint32_t)top_k) keys[pos] = k;
            }
        }
    }
    const bool bucket_sort = !presorted && m > 64;
    const int Kp = (presorted || bucket_sort) ? 0 : next_pow2(m);
    for (int i = m + tid; i < Kp; i += nthr) keys[i] = 0ull;
    __syncthreads();

    // ---- 2. sort, descending: bucket sort (the box arrays are not in use yet: they are its scratch), a bitonic
    //         network for the few-key problems ------------------------------------------------------------------
    if (bucket_sort) {
        int nb = 1024;
        while (nb > L.mcap) nb >>= 1;                          // scratch: 256 + 10 m + 4 nb + 4 <= 20 mcap bytes (m > 64)
        cta_bucket_sort(keys, m, smem + L.off_x1, nb);
    }
    for (int k = 2; k <= Kp; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (Kp >> 1); t += nthr) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i + j;
                const bool desc = (i & k) == 0;
                const unsigned long long a = keys[i], b = keys[l];
                if ((a < b) == desc) { keys[i] = b; keys[l] = a; }
            }
            __syncthreads();
        }
    }

    // ---- 3. gather boxes, extent -------------------------------------------------------------------
    if (tid < 4) misc[4 + tid] = (tid & 1) ? 0u : 0xffffffffu;
    if (tid == 0) misc[8] = 0;
    __syncthreads();
    const bool pixel = (pb.flags & RD_NMS_PIXEL_PLUS1) != 0;
    if (presorted) {
        // the rows as one flat array, 16 bytes per load (the buffer may be host memory read over PCIe)
        const int dim = pb.dets_dim, total = m * dim;
        const float4* src = reinterpret_cast<const float4*>(pb.dets);
        for (int e4 = tid; e4 * 4 < total; e4 += nthr) {
            const float4 v = src[e4];
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = e4 * 4 + q, row = e / dim, col = e - row * dim;
                if (e < total && col < 4) (col == 0 ? sx1 : col == 1 ? sy1 : col == 2 ? sx2 : sy2)[row] = vv[q];
            }
        }
        __syncthreads();
    }
    {
        uint32_t mnx = 0xffffffffu, mxx = 0, mny = 0xffffffffu, mxy = 0;
        for (int i = tid; i < m; i += nthr) {
            float4 b;
            if (presorted) {
                b = make_float4(sx1[i], sy1[i], sx2[i], sy2[i]);
            } else {
                b = pb.boxes[key_index(keys[i])];
                if (pb.has_scale) { b.x *= pb.scale.x; b.y *= pb.scale.y; b.z *= pb.scale.z; b.w *= pb.scale.w; }
                sx1[i] = b.x; sy1[i] = b.y; sx2[i] = b.z; sy2[i] = b.w;
            }
            const float cx = 0.5f * b.x + 0.5f * b.z, cy = 0.5f * b.y + 0.5f * b.w;
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
    for (int i = tid; i < 4 * kCols * L.WS; i += nthr) tab[i] = 0;
    __syncthreads();

    // ---- 4. bin ranges + start/end marks -------------------------------------------------------------
    const int Wm = (m + 31) >> 5;   // words in use
    {
        float lox, invx, loy, invy;
        extent_to_scale(misc[4], misc[5], lox, invx);
        extent_to_scale(misc[6], misc[7], loy, invy);
        const bool force_full = cull_disabled(pb.thr, pb.flags);
        for (int i = tid; i < m; i += nthr) {
            const uint32_t cr = bin_range(sx1[i], sy1[i], sx2[i], sy2[i], pixel, force_full, lox, invx, loy, invy);
            scr[i] = cr;
            const int ax = cr & 255u, bx = (cr >> 8) & 255u, ay = (cr >> 16) & 255u, by = cr >> 24;
            const uint32_t bit = 1u << (i & 31);
            const int w = i >> 5;
            atomicOr(&tab[(0 * kCols + ax) * L.WS + w], bit);
            if (bx + 1 < kCols) atomicOr(&tab[(1 * kCols + bx + 1) * L.WS + w], bit);
            atomicOr(&tab[(2 * kCols + ay) * L.WS + w], bit);
            if (by + 1 < kCols) atomicOr(&tab[(3 * kCols + by + 1) * L.WS + w], bit);
        }
    }
    __syncthreads();
    // inclusive prefix-OR over bins: S[c] = starts at <= c ; E[c] = ends (b) < c
    for (int task = tid; task < 4 * Wm; task += nthr) {
        const int t = task / Wm, w = task - t * Wm;
        uint32_t acc = 0;
        uint32_t* p = tab + (size_t)t * kCols * L.WS + w;
#pragma unroll 8
        for (int c = 0; c < kCols; ++c) { acc |= p[c * L.WS]; p[c * L.WS] = acc; }
    }
    __syncthreads();

    // ---- 5. greedy walk in chunks of up to kChunkBlocks x 32 candidates.
    //   (a) every warp takes items (mask word w, block bl of the chunk): lane = candidate of the block, the
    //       candidates' bin-surviving partners in word w -- KEPT boxes for words before the chunk, every
    //       earlier candidate inside it -- are flattened into the warp's own pair list and tested exactly by
    //       all 32 lanes (no CTA barrier inside a chunk; a box that overlaps many others does not serialise
    //       its lane).  A hit of a kept box kills the candidate; a hit inside the chunk sets a bit of its
    //       in-chunk suppressor mask tin[candidate][word];
    //   (b) warp 0 settles the blocks of the chunk in score order from those masks and emits the kept rows.
    uint32_t* pairs = reinterpret_cast<uint32_t*>(smem + L.off_pairs);
    uint32_t* s_tin = reinterpret_cast<uint32_t*>(smem + L.off_tin);
    uint32_t* s_dead = s_tin + kChunkBlocks * 32 * kTinStride;
    // misc: 8 kept count, 11 stop
    if (tid == 0) misc[11] = 0;
    const float thr = pb.thr;
    const int flags = pb.flags;
    const int max_out = pb.max_out;
    const uint32_t* Sx = tab;
    const uint32_t* Ex = tab + 1 * kCols * L.WS;
    const uint32_t* Sy = tab + 2 * kCols * L.WS;
    const uint32_t* Ey = tab + 3 * kCols * L.WS;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int nwarps = nthr >> 5;
    const int pw_cap = kLargePairCap / nwarps;
    uint32_t* pw = pairs + warp * pw_cap;
    int kept_total = 0;                                   // uniform over the CTA (read back from misc[8])
    for (int cb = 0; cb < Wm;) {
        int nb = Wm - cb;
        if (nb > kChunkBlocks) nb = kChunkBlocks;
        {   // rarely more than room + a few candidates are needed
            const int want = ((max_out - kept_total + 31) >> 5) + 1;
            if (nb > want) nb = want;
        }
        for (int i = tid; i < nb * 32 * kTinStride; i += nthr) s_tin[i] = 0;
        if (tid < kChunkBlocks) s_dead[tid] = 0;
        if (tid == 0) misc[12] = 0;                       // next item of the chunk
        __syncthreads();
        // items are drawn from a counter, not dealt round robin: their cost (pairs that survive the bins) varies by
        // an order of magnitude, and a fifth of the kernel's stall samples sat at the barrier behind this loop
        for (;;) {
            int it = 0;
            if (lane == 0) it = (int)atomicAdd(&misc[12], 1u);
            it = __shfl_sync(kFullMask, it, 0);
            // only the items that exist: every (earlier word, block) pair, then (word wi of the chunk, block bl >= wi)
            if (it >= cb * nb + (nb * (nb + 1)) / 2) break;
            int w, bl;
            if (it < cb * nb) {
                w = it / nb; bl = it - w * nb;
            } else {
                int t = it - cb * nb;
                bl = 0;
                while (t > bl) { t -= bl + 1; ++bl; }          // nb <= kChunkBlocks rounds
                w = cb + t;
            }
            const int jb = (cb + bl) * 32;
            const int j = jb + lane;
            uint32_t h = 0;
            if (j < m) {
                const uint32_t cr = scr[j];
                h = Sx[((cr >> 8) & 255u) * L.WS + w] & ~Ex[(cr & 255u) * L.WS + w] &
                    Sy[((cr >> 24) & 255u) * L.WS + w] & ~Ey[((cr >> 16) & 255u) * L.WS + w] &
                    (w < cb ? keptbits[w] : (w < cb + bl ? 0xffffffffu : lt_mask));
            }
            if (!__ballot_sync(kFullMask, h != 0u)) continue;
            const bool early = w < cb;
            uint32_t* trow = s_tin + (bl * 32) * kTinStride + (w - cb);
            const int nh = __popc(h);
            int off = nh;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(kFullMask, off, d); if (lane >= d) off += o; }
            const int wtot = __shfl_sync(kFullMask, off, 31);
            off -= nh;
            if (wtot <= pw_cap) {
                while (h) {
                    const int i = (w << 5) + __ffs(h) - 1;
                    h &= h - 1;
                    pw[off++] = ((uint32_t)lane << 16) | (uint32_t)i;
                }
                __syncwarp();
                for (int p = lane; p < wtot; p += 32) {
                    const uint32_t e = pw[p];
                    const int jl = (int)(e >> 16), i = (int)(e & 0xffffu);
                    const int jj = jb + jl;
                    if (suppresses(sx1[i], sy1[i], sx2[i], sy2[i], sx1[jj], sy1[jj], sx2[jj], sy2[jj], thr, flags)) {
                        if (early) atomicOr(&s_dead[bl], 1u << jl);
                        else atomicOr(&trow[jl * kTinStride], 1u << (i & 31));
                    }
                }
                __syncwarp();
            } else {                                      // more pairs than the warp's list holds: test in place
                const float x1 = sx1[j < m ? j : 0], y1 = sy1[j < m ? j : 0], x2 = sx2[j < m ? j : 0], y2 = sy2[j < m ? j : 0];
                uint32_t hit = 0;
                while (h) {
                    const int i = (w << 5) + __ffs(h) - 1;
                    h &= h - 1;
                    if (suppresses(sx1[i], sy1[i], sx2[i], sy2[i], x1, y1, x2, y2, thr, flags)) hit |= 1u << (i & 31);
                }
                if (hit) {
                    if (early) atomicOr(&s_dead[bl], 1u << lane);
                    else atomicOr(&trow[lane * kTinStride], hit);
                }
            }
        }
        __syncthreads();
        const int kept_before = kept_total;               // rows kept before this chunk (warp 0 advances kept_total below)
        if (warp == 0) {
            uint32_t kw[kChunkBlocks];                    // kept words of this chunk (uniform)
#pragma unroll
            for (int bl = 0; bl < kChunkBlocks; ++bl) {
                kw[bl] = 0;
                if (bl < nb && kept_total < max_out) {
                    const int j = (cb + bl) * 32 + lane;
                    const uint32_t* trow = s_tin + (bl * 32 + lane) * kTinStride;
                    bool alive = j < m && !((s_dead[bl] >> lane) & 1u);
#pragma unroll
                    for (int q = 0; q < kChunkBlocks; ++q)
                        if (q < bl && (trow[q] & kw[q])) alive = false;
                    // in-block resolution in score order: lane i survives iff no SURVIVING earlier lane of the block
                    // suppresses it.  The recursion runs over strictly lower lanes, so it has one solution and the
                    // iteration M <- { i : alive_i and tin_i & M == 0 } reaches it from any start -- bits 0 .. t-1 are
                    // final after t rounds -- in as many rounds as the longest suppression chain is deep (2 - 4),
                    // instead of one round per suppressor bit.
                    const uint32_t tin = trow[bl];
                    uint32_t keptw = __ballot_sync(kFullMask, alive);
                    if (__any_sync(kFullMask, alive && (tin & keptw))) {
                        for (;;) {
                            const uint32_t nm = __ballot_sync(kFullMask, alive && !(tin & keptw));
                            if (nm == keptw) break;
                            keptw = nm;
                        }
                    }
                    const int room = max_out - kept_total;
                    int cnt = __popc(keptw);
                    if (cnt > room) {   // keep only the first `room` set bits
                        uint32_t t = keptw, keep = 0;
                        for (int r = 0; r < room; ++r) { const uint32_t low = t & (0u - t); keep |= low; t ^= low; }
                        keptw = keep;
                        cnt = room;
                    }
                    kept_total += cnt;                        // (the rows are emitted by the whole CTA, below)
                    kw[bl] = keptw;
                    if (lane == 0) keptbits[cb + bl] = keptw;
                } else if (bl < nb && lane == 0) {
                    keptbits[cb + bl] = 0u;               // the output is full: nothing of this block is kept (or emitted)
                }
            }
            if (lane == 0) {
                misc[8] = (uint32_t)kept_total;
                if (kept_total >= max_out) misc[11] = 1;
            }
        }
        __syncthreads();
        // the chunk's kept rows, one candidate per thread: rank = rows kept before the chunk + kept bits before it
        // (warp 0 only decides; 5 - 7 scattered stores per kept row no longer sit in its serial walk)
        for (int t = tid; t < nb * 32; t += nthr) {
            const int q = t >> 5;
            const uint32_t kwq = keptbits[cb + q];
            if ((kwq >> (t & 31)) & 1u) {
                int rank = kept_before + __popc(kwq & ((1u << (t & 31)) - 1u));
                for (int q2 = 0; q2 < q; ++q2) rank += __popc(keptbits[cb + q2]);
                const int j = cb * 32 + t;
                sink_emit(sink, rank, keys[j], sx1[j], sy1[j], sx2[j], sy2[j]);
            }
        }
        kept_total = (int)misc[8];
        if (misc[11]) break;
        cb += nb;
    }
    return (int)misc[8];
}

}  // namespace rd

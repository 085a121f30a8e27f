// rd_loss.cu — tail of RefineDetMultiBoxLoss on B200 (sm_100a), forward and backward.
//
// Replaces (reference paths): layers/modules/refinedet_multibox_loss.py:96-139 — the ARM-theta gate
// of the positives (:96-101), SmoothL1 over the positives (:105-110), the per-anchor confidence loss
// log_sum_exp(conf) - conf.gather(conf_t) (:113-114, layers/box_utils.py:208-216), the cross-entropy
// over pos | neg (:126-130), the division by N (:134-138) — and their autograd backward.
//
// Kernels
//   conf_loss_kernel      one pass over conf[B*P,C]: per row  lse = log(sum exp(x - max)) + max,
//                         ce = lse - x[conf_t]  (the mining loss of :114 AND the cross-entropy term of
//                         :130 — F.cross_entropy(x, t) is the same quantity), pos = conf_t > 0 gated by
//                         softmax(arm_conf)[1] > theta.  Rows staged in shared memory, thread per row.
//                         The reference subtracts the GLOBAL max of the tensor (box_utils.py:215); the row
//                         max used here is the same value mathematically and at least as accurate in fp32.
//   loss_reduce_kernel    per image: sum of ce over pos | neg, SmoothL1 over pos (fp64 accumulation,
//                         fixed reduction tree -> deterministic)
//   loss_final_kernel     sums the per-image partials in order, N = sum(num_pos), divides
//   loss_backward_kernel  d loss_c / d conf = (softmax(x) - onehot(t)) * g_c / N on pos | neg rows,
//                         d loss_l / d loc = clamp(loc - loc_t, -1, 1) * g_l / N on pos rows, zeros
//                         elsewhere; every element of both gradients is written (selected rows twice)
#include <cstdlib>

#include "rd_common.cuh"

namespace rd {

constexpr int kLossThreads = 256;
constexpr int kLossMaxClasses = 128;
constexpr float kLog2e = 1.4426950408889634f;

// 2^x for x <= 0 (max-subtracted logits, log-probabilities): the bare MUFU.EX2.  exp2f() wraps the same instruction in
// a range test and two predicated multiplies that only matter for results below 2^-126 -- here those are terms of a
// sum that is >= 1 (or gradients below 1e-38), and the wrapper was 40 % of conf_loss_tma_kernel's instructions.
__device__ __forceinline__ float exp2_nonpos(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softmax(arm_conf)[1] <= theta (refinedet_multibox_loss.py:98-101), fp32, max-subtracted like F.softmax
__device__ __forceinline__ bool arm_filtered(float2 a, float theta) {
    const float m = fmaxf(a.x, a.y);
    const float e0 = expf(a.x - m), e1 = expf(a.y - m);
    return e1 / (e0 + e1) <= theta;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// one row staged in shared memory: lse = log(sum exp(x - max)) + max, ce = lse - x[t], pos (gated by the ARM theta)
__device__ __forceinline__ void conf_row(const float* __restrict__ x, int C, long long t, float2 arm, bool has_arm,
                                         float theta, float* ce_out, float* lse_out, unsigned char* pos_out) {
    // four independent chains each (a thread owns a whole row: the loop is latency-bound otherwise)
    float m0 = x[0], m1 = m0, m2 = m0, m3 = m0;
    int c = 0;
    for (; c + 4 <= C; c += 4) {
        m0 = fmaxf(m0, x[c]); m1 = fmaxf(m1, x[c + 1]); m2 = fmaxf(m2, x[c + 2]); m3 = fmaxf(m3, x[c + 3]);
    }
    for (; c < C; ++c) m0 = fmaxf(m0, x[c]);
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // exp(x - m) = exp2((x - m) log2 e): one MUFU.EX2, rel. error < 2e-7
    for (c = 0; c + 4 <= C; c += 4) {
        s0 += exp2_nonpos((x[c] - m) * kLog2e); s1 += exp2_nonpos((x[c + 1] - m) * kLog2e);
        s2 += exp2_nonpos((x[c + 2] - m) * kLog2e); s3 += exp2_nonpos((x[c + 3] - m) * kLog2e);
    }
    for (; c < C; ++c) s0 += exp2_nonpos((x[c] - m) * kLog2e);
    const float sum = (s0 + s1) + (s2 + s3);
    const float lse = logf(sum) + m;
    // a label outside [0, C) (dataset / num_classes mismatch) makes the reference's gather raise; here the
    // row's loss becomes NaN, which poisons loss_c visibly instead of reading past the staged row
    *ce_out = (t >= 0 && t < C) ? lse - x[(int)t] : __int_as_float(0x7fc00000);
    *lse_out = lse;
    bool pos = t > 0;
    if (pos && has_arm && arm_filtered(arm, theta)) pos = false;
    *pos_out = pos ? 1 : 0;
}

// generic C (3..128): a CTA stages kLossRows consecutive rows (one contiguous, coalesced float4 stream)
// in shared memory with an odd row stride (conflict-free), then one thread per row makes two passes
// over its row: max, sum of exp.  ~16 thread-instructions per element, against ~60 for a
// lane-per-class layout with shuffle reductions (the kernel is issue-bound, not bandwidth-bound).
constexpr int kLossRows = 128;            // rows (= threads) per CTA
constexpr int kLossLoads = 5;             // float4 loads in flight per thread while staging a tile

__global__ void __launch_bounds__(kLossRows)
conf_loss_kernel(const float* __restrict__ conf, const long long* __restrict__ conf_t,
                 const float2* __restrict__ arm_conf, float theta, long long rows, int C,
                 float* __restrict__ ce_out, float* __restrict__ lse_out, unsigned char* __restrict__ pos_out) {
    extern __shared__ __align__(16) float s_x[];     // [kLossRows][Cp]
    const int Cp = C | 1;
    const int tid = threadIdx.x;
    bool waited = false;                             // conf_t comes from the kernel before this one (match_pass2_kernel)
    for (long long r0 = (long long)blockIdx.x * kLossRows; r0 < rows; r0 += (long long)gridDim.x * kLossRows) {
        const int nrows = (int)min((long long)kLossRows, rows - r0);
        const int nelem = nrows * C;
        const float* src = conf + r0 * C;
        const long long r = r0 + tid;
        long long t = 0;
        float2 arm = make_float2(0.f, 0.f);
        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        const bool async_tile = aligned && Cp == C;
        if (async_tile) {
            // odd C: the tile is a verbatim copy (row stride C is already conflict-free): asynchronous 16-byte
            // global -> shared copies, all of a thread's ~20 in flight at once, no register staging.  conf is an
            // input of the whole chain, so the first tile is requested BEFORE waiting for the preceding kernel
            const int nvec = nelem >> 2;
            for (int q = tid; q < nvec; q += kLossRows)
                cp_async16(reinterpret_cast<float4*>(s_x) + q, reinterpret_cast<const float4*>(src) + q);
            for (int e2 = nvec * 4 + tid; e2 < nelem; e2 += kLossRows) s_x[e2] = src[e2];
        }
        if (!waited) { grid_dependency_wait(); waited = true; }
        // the scalar inputs of this thread's row
        if (tid < nrows) {
            t = conf_t[r];
            if (arm_conf && t > 0) arm = __ldg(arm_conf + r);
        }
        if (async_tile) {
            cp_async_wait_all();
        } else if (aligned) {
            // even C: one padding element per row, (row, class) of every element tracked incrementally
            const int nvec = nelem >> 2;
            int e = tid * 4;
            int rr = e / C, c = e - rr * C;
            const int drr = (4 * kLossRows) / C, dc = 4 * kLossRows - drr * C;
            for (int q0 = tid; q0 < nvec; q0 += kLossLoads * kLossRows) {
                float4 v[kLossLoads];                              // kLossLoads independent 16-byte loads in flight
#pragma unroll
                for (int u = 0; u < kLossLoads; ++u) {
                    const int q = q0 + u * kLossRows;
                    v[u] = q < nvec ? ldg_stream4(reinterpret_cast<const float4*>(src) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < kLossLoads; ++u) {
                    if (q0 + u * kLossRows < nvec) {
                        const float vv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                        int r1 = rr, c1 = c;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            s_x[r1 * Cp + c1] = vv[k];
                            if (++c1 == C) { c1 = 0; ++r1; }
                        }
                    }
                    rr += drr; c += dc;
                    if (c >= C) { c -= C; ++rr; }
                }
            }
            for (int e2 = nvec * 4 + tid; e2 < nelem; e2 += kLossRows) s_x[(e2 / C) * Cp + e2 % C] = src[e2];
        } else {
            for (int e2 = tid; e2 < nelem; e2 += kLossRows) s_x[(e2 / C) * Cp + e2 % C] = src[e2];
        }
        __syncthreads();
        if (tid < nrows)
            conf_row(s_x + tid * Cp, C, t, arm, arm_conf != nullptr, theta, ce_out + r, lse_out + r, pos_out + r);
        __syncthreads();
    }
}

// ---- TMA variant (odd C, 16-byte aligned conf): persistent CTAs, a ring of kTmaStages tiles of kTmaRows rows in
// shared memory.  A tile is one contiguous span of global memory (rows are consecutive, the odd row stride is already
// conflict-free), so ONE bulk asynchronous copy (cp.async.bulk, the 1-D TMA: SASS UBLKCP) per tile fetches it and
// signals the stage's mbarrier with the byte count; thread 0 issues the copy of tile i + kTmaStages - 1 before the CTA
// computes tile i.  No per-thread copy instructions, no register staging, loads of the next tile always in flight.
// Tile and ring are sized by two needs that compete for the 227 KB of shared memory: bytes in flight per SM (the
// tiles being fetched: Little's law against the ~3 us a tile takes to arrive under load) and warps per SM (the exp
// work).  Measured at C = 81, 174 MB: 128 rows x 2 stages (two CTAs per SM, 16 warps, 83 KB in flight) 42.0 us;
// 64 rows x 4 stages (8 warps, 124 KB) 47.1 us; 112 and 96 rows x 3 stages (14 / 12 warps, 145 / 124 KB) 44.0 us: neither more
// bytes in flight nor more warps move it -- the kernel sits at ~4.2 TB/s of pure reads like its cp.async predecessor.
#ifndef RD_TMA_ROWS
#define RD_TMA_ROWS 128
#endif
#ifndef RD_TMA_STAGES
#define RD_TMA_STAGES 2
#endif
constexpr int kTmaRows = RD_TMA_ROWS;                      // even, a multiple of 4 (tile bytes = 4 rows C: whole 16-byte words)
constexpr int kTmaStages = RD_TMA_STAGES;
constexpr int kTmaSplit = 2;                               // threads per row: lane pair (2 i, 2 i + 1) shares row i
constexpr int kTmaThreads = kTmaRows * kTmaSplit;

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}

// one row shared by kTmaSplit = 2 adjacent lanes: lane half h takes the classes h, h + 2, ... (with the odd row stride
// the 32 lanes of a warp -- 16 rows x 2 halves -- hit 32 different banks), max and sum are combined with one shuffle
// each.  Twice the warps for the same tile: the exp work of a tile takes half as long, and 16 warps per SM instead
// of 8 keep the issue slots busy while the next tile streams in.
__device__ __forceinline__ void conf_row_pair(const float* __restrict__ x, int C, int h, long long t, float2 arm, bool has_arm,
                                              float theta, float* ce_out, float* lse_out, unsigned char* pos_out) {
    const int n = (C - h + 1) >> 1;                        // classes of this half
    const float* xh = x + h;
    float m0 = xh[0], m1 = m0, m2 = m0, m3 = m0;
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        m0 = fmaxf(m0, xh[2 * k]); m1 = fmaxf(m1, xh[2 * k + 2]); m2 = fmaxf(m2, xh[2 * k + 4]); m3 = fmaxf(m3, xh[2 * k + 6]);
    }
    for (; k < n; ++k) m0 = fmaxf(m0, xh[2 * k]);
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    m = fmaxf(m, __shfl_xor_sync(kFullMask, m, 1));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (k = 0; k + 4 <= n; k += 4) {
        s0 += exp2_nonpos((xh[2 * k] - m) * kLog2e); s1 += exp2_nonpos((xh[2 * k + 2] - m) * kLog2e);
        s2 += exp2_nonpos((xh[2 * k + 4] - m) * kLog2e); s3 += exp2_nonpos((xh[2 * k + 6] - m) * kLog2e);
    }
    for (; k < n; ++k) s0 += exp2_nonpos((xh[2 * k] - m) * kLog2e);
    float sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(kFullMask, sum, 1);
    if (h != 0) return;
    const float lse = logf(sum) + m;
    *ce_out = (t >= 0 && t < C) ? lse - x[(int)t] : __int_as_float(0x7fc00000);     // label outside [0, C): NaN (see conf_row)
    *lse_out = lse;
    bool pos = t > 0;
    if (pos && has_arm && arm_filtered(arm, theta)) pos = false;
    *pos_out = pos ? 1 : 0;
}

// The same with the class count known at compile time: the half row (<= 41 values at C = 81) is read from shared
// memory ONCE into registers, both passes run on registers, and (x - m) log2 e is one fused multiply-add against the
// pre-scaled maximum -- 5 instructions per element instead of 8 (the kernel is bound by instruction issue, not by
// the tile stream: 17.3 M warp-instructions for 174 MB at C = 81 before this).
template <int kC>
__device__ __forceinline__ void conf_row_pair_regs(const float* __restrict__ x, int h, long long t, float2 arm, bool has_arm,
                                                   float theta, float* ce_out, float* lse_out, unsigned char* pos_out) {
    constexpr int kN = (kC + 1) / 2;                       // classes of half 0; half 1 has kC / 2
    const int n = (kC - h + 1) >> 1;
    const float* xh = x + h;
    float v[kN];
#pragma unroll
    for (int k = 0; k < kN; ++k) v[k] = (k < kC / 2 || k < n) ? xh[2 * k] : -INFINITY;
    float m0 = v[0], m1 = v[0], m2 = v[0], m3 = v[0];
#pragma unroll
    for (int k = 0; k < kN; ++k) {
        if ((k & 3) == 0) m0 = fmaxf(m0, v[k]);
        else if ((k & 3) == 1) m1 = fmaxf(m1, v[k]);
        else if ((k & 3) == 2) m2 = fmaxf(m2, v[k]);
        else m3 = fmaxf(m3, v[k]);
    }
    float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    m = fmaxf(m, __shfl_xor_sync(kFullMask, m, 1));
    const float nm2 = -m * kLog2e;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int k = 0; k < kN; ++k) {
        const float e = exp2_nonpos(__fmaf_rn(v[k], kLog2e, nm2));             // a padded -inf contributes exp2(-inf) = 0
        if ((k & 3) == 0) s0 += e;
        else if ((k & 3) == 1) s1 += e;
        else if ((k & 3) == 2) s2 += e;
        else s3 += e;
    }
    float sum = (s0 + s1) + (s2 + s3);
    sum += __shfl_xor_sync(kFullMask, sum, 1);
    if (h != 0) return;
    const float lse = logf(sum) + m;
    *ce_out = (t >= 0 && t < kC) ? lse - x[(int)t] : __int_as_float(0x7fc00000);    // label outside [0, C): NaN (see conf_row)
    *lse_out = lse;
    bool pos = t > 0;
    if (pos && has_arm && arm_filtered(arm, theta)) pos = false;
    *pos_out = pos ? 1 : 0;
}

template <int kC>
__global__ void __launch_bounds__(kTmaThreads)
conf_loss_tma_kernel(const float* __restrict__ conf, const long long* __restrict__ conf_t,
                     const float2* __restrict__ arm_conf, float theta, long long rows, int C,
                     float* __restrict__ ce_out, float* __restrict__ lse_out, unsigned char* __restrict__ pos_out) {
    extern __shared__ __align__(128) float s_x[];            // [kTmaStages][kTmaRows * C], each stage 16-byte aligned
    __shared__ __align__(8) unsigned long long s_bar[kTmaStages];
    const int tid = threadIdx.x;
    const int rit = tid >> 1, h = tid & 1;                   // row in the tile, half of the row
    const int tile_floats = kTmaRows * C;                    // multiple of 4
    const long long ntiles = (rows + kTmaRows - 1) / kTmaRows;
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < kTmaStages; ++st) mbar_init(&s_bar[st], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long tile, int st) {               // thread 0 only
        const long long r0 = tile * kTmaRows;
        const int nrows = (int)min((long long)kTmaRows, rows - r0);
        const unsigned bytes = (unsigned)(((size_t)nrows * C * 4) & ~(size_t)15);      // whole 16-byte words; tail below
        mbar_expect_tx(&s_bar[st], bytes);                   // (zero bytes: the arrival alone completes the phase)
        if (bytes) tma_load_1d(s_x + (size_t)st * tile_floats, conf + r0 * C, bytes, &s_bar[st]);
    };
    // conf is an input of the whole chain: the first tiles are requested before waiting for the preceding kernel
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kTmaStages - 1; ++k) {
            const long long tile = (long long)blockIdx.x + (long long)k * gridDim.x;
            if (tile < ntiles) issue(tile, k);
        }
    }
    grid_dependency_wait();                                   // conf_t comes from the kernel before this one
    // the row's scalar inputs are fetched one tile AHEAD, like the tile itself
    auto row_inputs = [&](long long tile, long long& t, float2& arm) {
        t = 0; arm = make_float2(0.f, 0.f);
        const long long r = tile * kTmaRows + rit;
        if (tile < ntiles && r < rows && h == 0) {
            t = conf_t[r];
            if (arm_conf) arm = __ldg(arm_conf + r);
        }
    };
    long long t_cur, t_nxt;
    float2 arm_cur, arm_nxt;
    row_inputs(blockIdx.x, t_cur, arm_cur);
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int st = it % kTmaStages;
        const unsigned parity = (unsigned)((it / kTmaStages) & 1);
        const long long r0 = tile * kTmaRows;
        const int nrows = (int)min((long long)kTmaRows, rows - r0);
        // keep the ring full: the stage refilled here was released by the barrier at the end of the previous tile
        if (tid == 0) {
            const long long nxt = tile + (long long)(kTmaStages - 1) * gridDim.x;
            if (nxt < ntiles) issue(nxt, (it + kTmaStages - 1) % kTmaStages);
        }
        row_inputs(tile + gridDim.x, t_nxt, arm_nxt);
        float* x_tile = s_x + (size_t)st * tile_floats;
        const int nelem = nrows * C;
        for (int e2 = (nelem & ~3) + tid; e2 < nelem; e2 += kTmaThreads) x_tile[e2] = conf[r0 * C + e2];   // < 4 floats, last tile only
        mbar_wait(&s_bar[st], parity);
        if ((nelem & 3) != 0) __syncthreads();
        const long long r = r0 + rit;
        // both lanes of a pair take part (shuffles); a pair past the last row works on row 0 of the tile and writes nothing
        const bool live = rit < nrows;
        if (kC > 0)
            conf_row_pair_regs<(kC > 0 ? kC : 3)>(x_tile + (live ? rit : 0) * C, live ? h : 1, t_cur, arm_cur, arm_conf != nullptr,
                                                  theta, ce_out + r, lse_out + r, pos_out + r);
        else
            conf_row_pair(x_tile + (live ? rit : 0) * C, C, live ? h : 1, t_cur, arm_cur, arm_conf != nullptr, theta, ce_out + r,
                          lse_out + r, pos_out + r);
        __syncthreads();                                      // every thread is done with the stage: it may be refilled
        t_cur = t_nxt; arm_cur = arm_nxt;
    }
}

// C == 2 (the ARM criterion): thread per row
__global__ void __launch_bounds__(kLossThreads)
conf_loss2_kernel(const float2* __restrict__ conf, const long long* __restrict__ conf_t,
                  const float2* __restrict__ arm_conf, float theta, long long rows,
                  float* __restrict__ ce_out, float* __restrict__ lse_out, unsigned char* __restrict__ pos_out) {
    const long long r = (long long)blockIdx.x * kLossThreads + threadIdx.x;
    if (r >= rows) return;
    const float2 x = ldg_stream2(conf + r);
    grid_dependency_wait();                          // conf_t comes from the kernel before this one
    const long long t = conf_t[r];
    const float m = fmaxf(x.x, x.y);
    const float s = expf(x.x - m) + expf(x.y - m);
    const float lse = logf(s) + m;
    ce_out[r] = t == 1 ? lse - x.y : t == 0 ? lse - x.x : __int_as_float(0x7fc00000);   // label outside {0, 1}: NaN
    lse_out[r] = lse;
    bool pos = t > 0;
    if (pos && arm_conf && arm_filtered(__ldg(arm_conf + r), theta)) pos = false;
    pos_out[r] = pos ? 1 : 0;
}

// SmoothL1 (beta = 1, reduction = sum), F.smooth_l1_loss: 0.5 d^2 if |d| < 1 else |d| - 0.5
__device__ __forceinline__ double smooth_l1(float p, float t) {
    const float d = p - t;
    const float a = fabsf(d);
    return a < 1.0f ? 0.5 * (double)d * (double)d : (double)a - 0.5;
}

// grid = (kReduceSplit, B): partial[b][g] = { sum_{pos} SmoothL1(loc - loc_t), sum_{pos|neg} ce } over the
// anchors i = g*256 + tid (mod kReduceSplit*256) of image b
constexpr int kReduceSplit = 8;

// the final step, one warp: lane l sums the partials q = l, l + 32, ... in order, then a fixed xor tree ->
// deterministic.  N = sum(num_pos) (:134); N < 1 -> zeros (:135-136)
__device__ __forceinline__ void loss_final_warp(const double* __restrict__ partial, const int* __restrict__ num_pos, int B,
                                                float* loss_l, float* loss_c, float* n_out) {
    const int lane = threadIdx.x & 31;
    double tl = 0.0, tc = 0.0;
    long long n = 0;
    for (int q = lane; q < B * kReduceSplit; q += 32) { tl += partial[2 * q]; tc += partial[2 * q + 1]; }
    for (int b = lane; b < B; b += 32) n += num_pos[b];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        tl += __shfl_xor_sync(kFullMask, tl, d);
        tc += __shfl_xor_sync(kFullMask, tc, d);
        n += __shfl_xor_sync(kFullMask, n, d);
    }
    if (lane == 0) {
        const float N = (float)n;
        *n_out = N;
        *loss_l = n > 0 ? (float)tl / N : 0.f;
        *loss_c = n > 0 ? (float)tc / N : 0.f;
    }
}

// kFinal: the CTA that finishes last (a ticket counter, zero before the launch and left zero) also runs the final
// step, in the same fixed order as the separate kernel -- one launch and ~6 us of pure latency less per criterion.
template <bool kFinal>
__global__ void __launch_bounds__(kLossThreads)
loss_reduce_kernel(const float4* __restrict__ loc, const float4* __restrict__ loc_t, const float* __restrict__ ce,
                   const unsigned char* __restrict__ pos, const unsigned char* __restrict__ neg, int P,
                   double* __restrict__ partial, unsigned int* ticket, const int* __restrict__ num_pos, int B,
                   float* loss_l, float* loss_c, float* n_out) {
    __shared__ double s_l[kLossThreads / 32], s_c[kLossThreads / 32];
    __shared__ int s_last;
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t img = (size_t)b * P;
    grid_dependency_wait();                          // neg / num_pos of the mining kernel (and everything before it)
    double al = 0.0, ac = 0.0;
    // the masks and ce are loaded UNCONDITIONALLY (a ce load that waits for the masks' values is a second trip to
    // DRAM per anchor), four consecutive anchors per thread when the row allows 16-byte loads; only the loc rows of
    // the positives (~1.5 %) are fetched on demand
    const bool vec = (P & 3) == 0 && (reinterpret_cast<uintptr_t>(ce) & 15) == 0 &&
                     ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(neg)) & 3) == 0;
    if (vec) {
        for (int i = 4 * (g * kLossThreads + tid); i < P; i += 4 * kReduceSplit * kLossThreads) {
            const float4 c4 = ldg_stream4(reinterpret_cast<const float4*>(ce + img + i));
            const uchar4 p4 = __ldg(reinterpret_cast<const uchar4*>(pos + img + i));
            const uchar4 n4 = __ldg(reinterpret_cast<const uchar4*>(neg + img + i));
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
            const unsigned char pv[4] = {p4.x, p4.y, p4.z, p4.w}, nv[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (pv[k] | nv[k]) ac += (double)cv[k];
                if (pv[k]) {
                    const float4 x = loc[img + i + k], t = loc_t[img + i + k];
                    al += smooth_l1(x.x, t.x) + smooth_l1(x.y, t.y) + smooth_l1(x.z, t.z) + smooth_l1(x.w, t.w);
                }
            }
        }
    } else {
        for (int i = g * kLossThreads + tid; i < P; i += kReduceSplit * kLossThreads) {
            const bool p = pos[img + i] != 0;
            const bool n = neg[img + i] != 0;
            const float c1 = ce[img + i];
            if (p | n) ac += (double)c1;
            if (p) {
                const float4 x = loc[img + i], t = loc_t[img + i];
                al += smooth_l1(x.x, t.x) + smooth_l1(x.y, t.y) + smooth_l1(x.z, t.z) + smooth_l1(x.w, t.w);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        al += __shfl_xor_sync(kFullMask, al, d);
        ac += __shfl_xor_sync(kFullMask, ac, d);
    }
    if (lane == 0) { s_l[warp] = al; s_c[warp] = ac; }
    __syncthreads();
    if (tid == 0) {
        double tl = 0.0, tc = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) { tl += s_l[w]; tc += s_c[w]; }
        partial[2 * (b * kReduceSplit + g)] = tl;
        partial[2 * (b * kReduceSplit + g) + 1] = tc;
        if (kFinal) {
            __threadfence();                         // the partials are visible before the ticket is
            s_last = atomicAdd(ticket, 1u) == (unsigned int)(gridDim.x * gridDim.y) - 1u;
        }
    }
    if (!kFinal) return;
    __syncthreads();
    if (!s_last) return;
    if (warp == 0) {
        __threadfence();
        loss_final_warp(partial, num_pos, B, loss_l, loss_c, n_out);
        if (lane == 0) *ticket = 0u;
    }
}

__global__ void loss_final_kernel(const double* __restrict__ partial, const int* __restrict__ num_pos, int B,
                                  float* loss_l, float* loss_c, float* n_out) {
    loss_final_warp(partial, num_pos, B, loss_l, loss_c, n_out);
}

// gradients.  One warp per 32 rows; conf rows of unselected anchors are never read.
__global__ void __launch_bounds__(kLossThreads)
loss_backward_kernel(const float4* __restrict__ loc, const float4* __restrict__ loc_t,
                     const float* __restrict__ conf, const long long* __restrict__ conf_t,
                     const float* __restrict__ lse, const unsigned char* __restrict__ pos,
                     const unsigned char* __restrict__ neg, const float* __restrict__ g_l,
                     const float* __restrict__ g_c, const float* __restrict__ n_dev, long long rows, int C,
                     float4* __restrict__ grad_loc, float* __restrict__ grad_conf) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = ((long long)blockIdx.x * kLossThreads + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * kLossThreads) >> 5;
    const float N = *n_dev;
    const float sl = (g_l && N > 0.f) ? *g_l / N : 0.f;
    const float sc = (g_c && N > 0.f) ? *g_c / N : 0.f;
    const int nseg = (C + 31) >> 5;
    for (long long r0 = warp_global * 32; r0 < rows; r0 += nwarps * 32) {
        // 32 rows per step: lane l owns the flags (and the loc gradient) of row r0 + l
        const long long r = r0 + lane;
        bool p = false, sel = false;
        if (r < rows) {
            p = pos[r] != 0;
            sel = p || neg[r] != 0;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grad_loc) {
                if (p) {
                    const float4 x = loc[r], t = loc_t[r];
                    g.x = fminf(fmaxf(x.x - t.x, -1.f), 1.f) * sl;      // d SmoothL1 / d x = clamp(x - t, -1, 1)
                    g.y = fminf(fmaxf(x.y - t.y, -1.f), 1.f) * sl;
                    g.z = fminf(fmaxf(x.z - t.z, -1.f), 1.f) * sl;
                    g.w = fminf(fmaxf(x.w - t.w, -1.f), 1.f) * sl;
                }
                grad_loc[r] = g;
            }
        }
        if (!grad_conf) continue;
        const unsigned selmask = __ballot_sync(kFullMask, sel);
        const int nvalid = (int)min((long long)32, rows - r0);
        // the first kPre selected rows (pos | neg, ~2 of 32) are fetched BEFORE the zero stores are issued, lane =
        // class: their trip to DRAM overlaps the stores instead of following them one row after the other
        constexpr int kPre = 4;
        float xs[kPre][4];
        float ls[kPre];
        int ts[kPre], rrs[kPre];
        unsigned m = selmask;
#pragma unroll
        for (int u = 0; u < kPre; ++u) {
            rrs[u] = -1;
            if (m) {
                rrs[u] = __ffs(m) - 1;
                m &= m - 1;
                const long long row = r0 + rrs[u];
                ls[u] = lse[row];
                ts[u] = (int)conf_t[row];
                const float* x = conf + row * C;
#pragma unroll
                for (int sgm = 0; sgm < 4; ++sgm) {
                    const int c = sgm * 32 + lane;
                    xs[u][sgm] = (sgm < nseg && c < C) ? x[c] : 0.f;
                }
            }
        }
        // zeros over the whole 32*C-element span (contiguous 16-byte stores, no per-element bookkeeping); the
        // few selected rows are then overwritten by the same warp, ordered by __syncwarp
        float* base = grad_conf + r0 * C;
        const int nelem = nvalid * C;
        const int nvec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) ? (nelem >> 2) : 0;
        for (int q = lane; q < nvec; q += 32) reinterpret_cast<float4*>(base)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = nvec * 4 + lane; e < nelem; e += 32) base[e] = 0.f;
        __syncwarp();
        // selected rows: softmax(x) - onehot(t), lane = class
#pragma unroll
        for (int u = 0; u < kPre; ++u) {
            if (rrs[u] >= 0) {
                float* g = grad_conf + (r0 + rrs[u]) * C;
#pragma unroll
                for (int sgm = 0; sgm < 4; ++sgm) {
                    const int c = sgm * 32 + lane;
                    if (sgm < nseg && c < C) g[c] = (exp2_nonpos((xs[u][sgm] - ls[u]) * kLog2e) - (c == ts[u] ? 1.f : 0.f)) * sc;
                }
            }
        }
        while (m) {                                           // more than kPre selected rows in the group: one by one
            const int rr = __ffs(m) - 1;
            m &= m - 1;
            const long long row = r0 + rr;
            const float l = lse[row];
            const int t = (int)conf_t[row];
            const float* x = conf + row * C;
            float* g = grad_conf + row * C;
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                if (sgm < nseg && c < C) g[c] = (exp2_nonpos((x[c] - l) * kLog2e) - (c == t ? 1.f : 0.f)) * sc;
            }
        }
    }
}


// ---- TMA variant of the backward (odd C, 16-byte aligned grad_conf): the gradient of a tile of kBwdRows rows is
// BUILT in shared memory -- zeros, with the few selected rows (pos | neg, ~6 %) computed into place -- and leaves as ONE
// bulk asynchronous store (cp.async.bulk.global.shared::cta, the 1-D TMA store: SASS UBLKCP), two tiles in flight per
// CTA.  The 127 MB of zeros no longer pass through the load/store units as 16-byte stores of every lane (the
// register kernel spends a fifth of its stall samples in lg_throttle); only the rows selected in a buffer's previous
// use are re-zeroed.  Persistent CTAs of four warps over 64-row tiles (41 KB of shared memory: five CTAs per SM, so
// that the trips to DRAM of one CTA -- flags, then the selected rows -- hide behind the others'); thread = row for
// the flags (fetched one tile ahead) and the loc gradient, warp = selected row (lane = class) for the conf gradient.
constexpr int kBwdRows = 64;
constexpr int kBwdThreads = 128;

__global__ void __launch_bounds__(kBwdThreads)
loss_backward_tma_kernel(const float4* __restrict__ loc, const float4* __restrict__ loc_t,
                         const float* __restrict__ conf, const long long* __restrict__ conf_t,
                         const float* __restrict__ lse, const unsigned char* __restrict__ pos,
                         const unsigned char* __restrict__ neg, const float* __restrict__ g_l,
                         const float* __restrict__ g_c, const float* __restrict__ n_dev, long long rows, int C,
                         float4* __restrict__ grad_loc, float* __restrict__ grad_conf) {
    extern __shared__ __align__(128) float s_g[];            // [2][kBwdRows * C]
    __shared__ unsigned char s_list[2][kBwdRows];            // rows of the tile that hold a non-zero gradient
    __shared__ int s_nsel[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_floats = kBwdRows * C;
    const long long ntiles = (rows + kBwdRows - 1) / kBwdRows;
    const float N = *n_dev;
    const float sl = (g_l && N > 0.f) ? *g_l / N : 0.f;
    const float sc = (g_c && N > 0.f) ? *g_c / N : 0.f;
    const int nseg = (C + 31) >> 5;
    for (int i = tid; i < 2 * tile_floats / 4; i += kBwdThreads) reinterpret_cast<float4*>(s_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 2) s_nsel[tid] = 0;
    // flags of this thread's row, one tile ahead
    auto flags_of = [&](long long tile, unsigned char& pf, unsigned char& nf) {
        pf = 0; nf = 0;
        const long long r = tile * kBwdRows + tid;
        if (tile < ntiles && tid < kBwdRows && r < rows) { pf = pos[r]; nf = neg[r]; }
    };
    unsigned char pf, nf, pf_n, nf_n;
    flags_of(blockIdx.x, pf, nf);
    __syncthreads();
    int it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = it & 1;
        float* g_tile = s_g + (size_t)b * tile_floats;
        const long long r0 = tile * kBwdRows;
        const int nrows = (int)min((long long)kBwdRows, rows - r0);
        flags_of(tile + gridDim.x, pf_n, nf_n);
        // the store that last read this buffer (two tiles ago) is done with it: all groups but the latest have read
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        // back to zeros: only the rows written in the buffer's previous use
        const int nprev = s_nsel[b];
        for (int k = warp; k < nprev; k += kBwdThreads / 32) {
            float* g = g_tile + (int)s_list[b][k] * C;
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                if (sgm < nseg && c < C) g[c] = 0.f;
            }
        }
        __syncthreads();
        if (tid == 0) s_nsel[b] = 0;
        __syncthreads();
        // selected rows of this tile -> list; loc gradient: thread = row
        const long long r = r0 + tid;
        const bool in_tile = tid < nrows;                        // nrows <= kBwdRows <= kBwdThreads
        const bool sel = in_tile && (pf | nf) != 0;
        if (in_tile && grad_loc) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pf) {
                const float4 x = loc[r], t = loc_t[r];
                g.x = fminf(fmaxf(x.x - t.x, -1.f), 1.f) * sl;      // d SmoothL1 / d x = clamp(x - t, -1, 1)
                g.y = fminf(fmaxf(x.y - t.y, -1.f), 1.f) * sl;
                g.z = fminf(fmaxf(x.z - t.z, -1.f), 1.f) * sl;
                g.w = fminf(fmaxf(x.w - t.w, -1.f), 1.f) * sl;
            }
            grad_loc[r] = g;
        }
        const unsigned bal = __ballot_sync(kFullMask, sel);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_nsel[b], __popc(bal));
            base = __shfl_sync(kFullMask, base, 0);
            if (sel) s_list[b][base + __popc(bal & ((1u << lane) - 1u))] = (unsigned char)tid;
        }
        __syncthreads();
        // selected rows: softmax(x) - onehot(t), warp = row, lane = class; two rows in flight per warp
        const int nsel = s_nsel[b];
        for (int k = warp; k < nsel; k += 2 * (kBwdThreads / 32)) {
            const int k2 = k + kBwdThreads / 32;
            const int ra = s_list[b][k], rb = k2 < nsel ? s_list[b][k2] : -1;
            const long long rowa = r0 + ra, rowb = r0 + (rb >= 0 ? rb : ra);
            const float la = lse[rowa], lb = lse[rowb];
            const int ta = (int)conf_t[rowa], tb = (int)conf_t[rowb];
            float xa[4], xb[4];
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                const bool on = sgm < nseg && c < C;
                xa[sgm] = on ? conf[rowa * C + c] : 0.f;
                xb[sgm] = (on && rb >= 0) ? conf[rowb * C + c] : 0.f;
            }
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                if (sgm < nseg && c < C) {
                    g_tile[ra * C + c] = (exp2_nonpos((xa[sgm] - la) * kLog2e) - (c == ta ? 1.f : 0.f)) * sc;
                    if (rb >= 0) g_tile[rb * C + c] = (exp2_nonpos((xb[sgm] - lb) * kLog2e) - (c == tb ? 1.f : 0.f)) * sc;
                }
            }
        }
        // shared-memory writes of the generic proxy -> visible to the bulk-copy engine, then one store for the tile
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const int nelem = nrows * C;
        if (tid == 0) {
            const unsigned bytes = (unsigned)(((size_t)nelem * 4) & ~(size_t)15);
            if (bytes)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grad_conf + r0 * C),
                             "r"((unsigned)__cvta_generic_to_shared(g_tile)), "r"(bytes)
                             : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        for (int e2 = (nelem & ~3) + tid; e2 < nelem; e2 += kBwdThreads) grad_conf[r0 * C + e2] = g_tile[e2];   // < 4 floats, last tile only
        pf = pf_n; nf = nf_n;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory stays alive until the stores have read it
}

// ---- zero-stream variant of the backward (any C, 16-byte aligned grad_conf; the default).  grad_conf is 94 % zeros:
// they leave as bulk asynchronous stores (cp.async.bulk.global.shared::cta, SASS UBLKCP) from ONE constant 16 KB
// block of zeros in shared memory -- nothing is built, re-zeroed or waited for on that stream: lane 0 of the last warp
// keeps kZsDepth + 1 tiles in flight and publishes in `s_done` how many of this CTA's tiles have LANDED (wait_group
// without .read).  135 MB of zeros alone take 33 us this way (tools/zero_bw.cu: memset 29 us, 16-byte stores 31-33 us).
// The other eight warps do every READ of the CTA up front, while the write queues are still short (a load issued
// behind 100 MB of queued zeros takes several us -- the first version, which fetched flags and rows tile by tile,
// spent 53 us, 60 % of its stall samples waiting for those loads):
//   1. flags of all rows of the CTA's tiles (thread = row, four independent rows per thread), loc gradient, list of
//      the selected rows (pos | neg, ~6 %) in shared memory;
//   2. softmax - onehot of the listed rows (warp = row, lane = class, four rows in flight per warp) into a staging
//      area in shared memory;
//   3. staged rows written over the zeros of their tile once it has landed.
// Rows beyond the staging area (a CTA with more than kZsStageBytes of selected rows) are fetched and written one by
// one after their tile has landed.  (The tile variant above builds every 64-row tile in shared memory between four
// barriers: 38 us under ncu, IPC 0.78.)
constexpr int kZsRowWarps = 8;
constexpr int kZsRowThreads = 32 * kZsRowWarps;
constexpr int kZsThreads = kZsRowThreads + 32;
constexpr int kZsZeroBytes = 16 * 1024;
constexpr int kZsDepth = 1;                  // + the tile being issued: with three CTAs per SM >= 330 KB of stores in flight per SM
constexpr int kZsList = 2048;                // rows of one round (a round = as many tiles as fit)
constexpr int kZsStageBytes = 40 * 1024;

__device__ __forceinline__ void zs_row_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kZsRowThreads) : "memory"); }

__global__ void __launch_bounds__(kZsThreads)
loss_backward_zs_kernel(const float4* __restrict__ loc, const float4* __restrict__ loc_t,
                        const float* __restrict__ conf, const long long* __restrict__ conf_t,
                        const float* __restrict__ lse, const unsigned char* __restrict__ pos,
                        const unsigned char* __restrict__ neg, const float* __restrict__ g_l,
                        const float* __restrict__ g_c, const float* __restrict__ n_dev, long long rows, int C,
                        int groups_per_tile, float4* __restrict__ grad_loc, float* __restrict__ grad_conf) {
    extern __shared__ __align__(128) unsigned char s_zs[];   // [zeros kZsZeroBytes | staging kZsStageBytes]
    __shared__ unsigned int s_list[kZsList];                 // (tile of this CTA << 16) | row in the tile
    __shared__ int s_n, s_done;                              // listed rows; tiles of this CTA whose zeros have landed
    unsigned char* s_zero = s_zs;
    float* s_stage = reinterpret_cast<float*>(s_zs + kZsZeroBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_rows = 32 * groups_per_tile;              // <= kZsList
    const long long ntiles = (rows + tile_rows - 1) / tile_rows;
    const int T = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);       // tiles of this CTA (grid <= ntiles)
    for (int i = tid; i < kZsZeroBytes / 16; i += kZsThreads) reinterpret_cast<float4*>(s_zero)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) s_done = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the zeros -> visible to the bulk-copy engine
    __syncthreads();
    volatile int* done = &s_done;
    if (warp == kZsRowWarps) {
        if (lane != 0 || !grad_conf) return;
        for (int it = 0; it < T; ++it) {
            const long long r0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * tile_rows;
            const long long nelem = min((long long)tile_rows, rows - r0) * C;
            const size_t bytes = ((size_t)nelem * 4) & ~(size_t)15;
            unsigned char* dst = reinterpret_cast<unsigned char*>(grad_conf + r0 * C);
            for (size_t off = 0; off < bytes; off += kZsZeroBytes) {
                const unsigned n = (unsigned)min((size_t)kZsZeroBytes, bytes - off);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off),
                             "r"((unsigned)__cvta_generic_to_shared(s_zero)), "r"(n)
                             : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            for (long long e = (long long)(bytes >> 2); e < nelem; ++e) grad_conf[r0 * C + e] = 0.f;   // < 4 floats, last tile only
            if (it >= kZsDepth) {
                asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kZsDepth) : "memory");
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __threadfence_block();
                *done = it - kZsDepth + 1;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // (also: shared memory stays alive until it was read)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        __threadfence_block();
        *done = T;
        return;
    }
    const float N = *n_dev;
    const float sl = (g_l && N > 0.f) ? *g_l / N : 0.f;
    const float sc = (g_c && N > 0.f) ? *g_c / N : 0.f;
    const int nseg = (C + 31) >> 5;
    const int cap = min(kZsList, kZsStageBytes / (4 * C));   // rows the staging area holds
    const int R = kZsList / tile_rows;                       // tiles per round: the list cannot overflow
    auto row_of = [&](unsigned int e) {
        return ((long long)blockIdx.x + (long long)(e >> 16) * gridDim.x) * tile_rows + (long long)(e & 0xffffu);
    };
    for (int it0 = 0; it0 < T; it0 += R) {
        const int it1 = min(T, it0 + R);
        zs_row_barrier();                                    // the previous round is done with the list and the staging area
        if (tid == 0) s_n = 0;
        zs_row_barrier();
        // 1. flags, loc gradient, list
        const int round_rows = (it1 - it0) * tile_rows;
        for (int f0 = 0; f0 < round_rows; f0 += 4 * kZsRowThreads) {
            unsigned char pf[4], nf[4];
            long long rr[4];
            unsigned int ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int f = f0 + u * kZsRowThreads + tid;
                const int itl = f / tile_rows, rit = f - itl * tile_rows;
                ee[u] = ((unsigned int)(it0 + itl) << 16) | (unsigned int)rit;
                rr[u] = f < round_rows ? row_of(ee[u]) : rows;
                pf[u] = 0; nf[u] = 0;
                if (rr[u] < rows) { pf[u] = pos[rr[u]]; nf[u] = neg[rr[u]]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long r = rr[u];
                if (r < rows && grad_loc) {
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pf[u]) {
                        const float4 x = loc[r], t = loc_t[r];
                        g.x = fminf(fmaxf(x.x - t.x, -1.f), 1.f) * sl;      // d SmoothL1 / d x = clamp(x - t, -1, 1)
                        g.y = fminf(fmaxf(x.y - t.y, -1.f), 1.f) * sl;
                        g.z = fminf(fmaxf(x.z - t.z, -1.f), 1.f) * sl;
                        g.w = fminf(fmaxf(x.w - t.w, -1.f), 1.f) * sl;
                    }
                    grad_loc[r] = g;
                }
                const bool sel = grad_conf && (pf[u] | nf[u]) != 0;
                const unsigned bal = __ballot_sync(kFullMask, sel);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
                    base = __shfl_sync(kFullMask, base, 0);
                    if (sel) s_list[base + __popc(bal & ((1u << lane) - 1u))] = ee[u];
                }
            }
        }
        zs_row_barrier();
        const int n = s_n, staged = min(n, cap);
        // 2. the listed rows -> staging area; four rows in flight per warp
        for (int k0 = 4 * warp; k0 < staged; k0 += 4 * kZsRowWarps) {
            float xs[4][4], ls[4];
            int ts[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + u < staged) {
                    const long long row = row_of(s_list[k0 + u]);
                    ls[u] = lse[row];
                    ts[u] = (int)conf_t[row];
                    const float* x = conf + row * C;
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) {
                        const int c = sgm * 32 + lane;
                        xs[u][sgm] = (sgm < nseg && c < C) ? x[c] : 0.f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k0 + u < staged) {
                    float* g = s_stage + (size_t)(k0 + u) * C;
#pragma unroll
                    for (int sgm = 0; sgm < 4; ++sgm) {
                        const int c = sgm * 32 + lane;
                        if (sgm < nseg && c < C) g[c] = (exp2_nonpos((xs[u][sgm] - ls[u]) * kLog2e) - (c == ts[u] ? 1.f : 0.f)) * sc;
                    }
                }
            }
        }
        __syncwarp();
        // 3. staged rows (written by this warp) over the zeros of their tile, once it has landed
        for (int k0 = 4 * warp; k0 < staged; k0 += 4 * kZsRowWarps) {
            for (int u = 0; u < 4 && k0 + u < staged; ++u) {
                const unsigned int e = s_list[k0 + u];
                while (*done <= (int)(e >> 16)) {}
                __threadfence_block();
                float* g = grad_conf + row_of(e) * C;
                const float* sg = s_stage + (size_t)(k0 + u) * C;
#pragma unroll
                for (int sgm = 0; sgm < 4; ++sgm) {
                    const int c = sgm * 32 + lane;
                    if (sgm < nseg && c < C) g[c] = sg[c];
                }
            }
        }
        // 4. rows the staging area could not hold
        for (int k = cap + warp; k < n; k += kZsRowWarps) {
            const unsigned int e = s_list[k];
            const long long row = row_of(e);
            const float l = lse[row];
            const int t = (int)conf_t[row];
            const float* x = conf + row * C;
            float xv[4];
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                xv[sgm] = (sgm < nseg && c < C) ? x[c] : 0.f;
            }
            while (*done <= (int)(e >> 16)) {}
            __threadfence_block();
            float* g = grad_conf + row * C;
#pragma unroll
            for (int sgm = 0; sgm < 4; ++sgm) {
                const int c = sgm * 32 + lane;
                if (sgm < nseg && c < C) g[c] = (exp2_nonpos((xv[sgm] - l) * kLog2e) - (c == t ? 1.f : 0.f)) * sc;
            }
        }
    }
}

}  // namespace rd

using namespace rd;

// launches the confidence-loss kernel that fits (C == 2: thread per row; odd C and aligned rows: the TMA ring;
// otherwise the cp.async / register-staged tiles)
static int conf_loss_launch(const float* conf, const long long* conf_t, const float* arm_conf, float theta, long long rows,
                            int C, float* ce_out, float* lse_out, unsigned char* pos_out, cudaStream_t st, bool pdl) {
    cudaError_t e = cudaSuccess;
    if (C == 2) {
        const long long blocks = (rows + kLossThreads - 1) / kLossThreads;
        if (pdl) e = launch_pdl(conf_loss2_kernel, dim3((unsigned)blocks), dim3(kLossThreads), 0, st, (const float2*)conf, conf_t,
                                (const float2*)arm_conf, theta, rows, ce_out, lse_out, pos_out);
        else conf_loss2_kernel<<<(unsigned)blocks, kLossThreads, 0, st>>>((const float2*)conf, conf_t, (const float2*)arm_conf,
                                                                          theta, rows, ce_out, lse_out, pos_out);
    } else if ((C & 1) && ((uintptr_t)conf & 15) == 0 && !getenv("RD_NO_TMA")) {
        const size_t smem = (size_t)kTmaStages * kTmaRows * C * sizeof(float);
        const long long ntiles = (rows + kTmaRows - 1) / kTmaRows;
        long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
        if (per_sm < 1) per_sm = 1;
        long long blocks = 148 * per_sm;                      // persistent: one wave
        if (blocks > ntiles) blocks = ntiles;
        auto kern = C == 81 ? conf_loss_tma_kernel<81> : C == 21 ? conf_loss_tma_kernel<21> : conf_loss_tma_kernel<0>;
        static size_t s_tma_smem[3][kMaxDevices];
        e = ensure_dynamic_smem(kern, smem, s_tma_smem[C == 81 ? 0 : C == 21 ? 1 : 2]);
        if (e != cudaSuccess) return (int)e;
        if (pdl) e = launch_pdl(kern, dim3((unsigned)blocks), dim3(kTmaThreads), smem, st, conf, conf_t,
                                (const float2*)arm_conf, theta, rows, C, ce_out, lse_out, pos_out);
        else kern<<<(unsigned)blocks, kTmaThreads, smem, st>>>(conf, conf_t, (const float2*)arm_conf, theta, rows, C, ce_out,
                                                               lse_out, pos_out);
    } else {
        const size_t smem = (size_t)kLossRows * (C | 1) * sizeof(float);
        static size_t s_conf_smem[kMaxDevices];
        e = ensure_dynamic_smem(conf_loss_kernel, smem, s_conf_smem);
        if (e != cudaSuccess) return (int)e;
        long long blocks = (rows + kLossRows - 1) / kLossRows;
        if (blocks > 148 * 64) blocks = 148 * 64;
        if (pdl) e = launch_pdl(conf_loss_kernel, dim3((unsigned)blocks), dim3(kLossRows), smem, st, conf, conf_t,
                                (const float2*)arm_conf, theta, rows, C, ce_out, lse_out, pos_out);
        else conf_loss_kernel<<<(unsigned)blocks, kLossRows, smem, st>>>(conf, conf_t, (const float2*)arm_conf, theta, rows, C,
                                                                         ce_out, lse_out, pos_out);
    }
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

extern "C" {

size_t rd_multibox_loss_workspace_bytes(int B) { return B > 0 ? (size_t)B * kReduceSplit * 2 * sizeof(double) : 0; }

int rd_conf_loss(const float* conf, const long long* conf_t, const float* arm_conf, float theta, long long rows,
                 int C, float* ce_out, float* lse_out, unsigned char* pos_out, void* stream) {
    NvtxRange nvtx_range("rd_conf_loss");
    if (!conf || !conf_t || !ce_out || !lse_out || !pos_out || rows <= 0 || C < 2) return RD_ERR_BAD_ARG;
    if (C > kLossMaxClasses) return RD_ERR_UNSUPPORTED;
    if (arm_conf && ((uintptr_t)arm_conf & 7)) return RD_ERR_ALIGNMENT;
    if (C == 2 && ((uintptr_t)conf & 7)) return RD_ERR_ALIGNMENT;
    return conf_loss_launch(conf, conf_t, arm_conf, theta, rows, C, ce_out, lse_out, pos_out, (cudaStream_t)stream, false);
}

int rd_multibox_loss_reduce(const float* loc, const float* loc_t, const float* ce, const unsigned char* pos,
                            const unsigned char* neg, const int* num_pos, int B, int P, void* workspace,
                            size_t workspace_bytes, float* loss_l, float* loss_c, float* n_out, void* stream) {
    NvtxRange nvtx_range("rd_multibox_loss_reduce");
    if (!loc || !loc_t || !ce || !pos || !neg || !num_pos || !workspace || !loss_l || !loss_c || !n_out || B <= 0 ||
        P <= 0)
        return RD_ERR_BAD_ARG;
    if (((uintptr_t)loc | (uintptr_t)loc_t) & 15) return RD_ERR_ALIGNMENT;
    if ((uintptr_t)workspace & 7) return RD_ERR_ALIGNMENT;
    if (workspace_bytes < rd_multibox_loss_workspace_bytes(B)) return RD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    loss_reduce_kernel<false><<<dim3(kReduceSplit, B), kLossThreads, 0, st>>>((const float4*)loc, (const float4*)loc_t, ce, pos, neg,
                                                                              P, (double*)workspace, nullptr, num_pos, B, loss_l,
                                                                              loss_c, n_out);
    note_launch();
    RD_CHECK_LAUNCH();
    loss_final_kernel<<<1, 32, 0, st>>>((const double*)workspace, num_pos, B, loss_l, loss_c, n_out);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

// ---- the whole criterion forward as one call ---------------------------------------------------------------------
// workspace: [ticket u32, padded to 256 B | best_prior u64 [B][Gmax] | bt_idx i32 [B][P] | bt_overlap f32 [B][P] |
//             partial f64 [B][kReduceSplit][2]]; ticket + best_prior are cleared by one memset
struct CriterionWs {
    unsigned int* ticket;
    unsigned long long* best_prior;
    int* bt_idx;
    float* bt_ov;
    double* partial;
    size_t clear_bytes, total;
};
static CriterionWs carve_criterion(void* base, int B, int P, int Gmax) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    CriterionWs w;
    unsigned char* p = static_cast<unsigned char*>(base);
    size_t o = 0;
    w.ticket = reinterpret_cast<unsigned int*>(p + o);             o += 256;
    w.best_prior = reinterpret_cast<unsigned long long*>(p + o);   o += up((size_t)B * Gmax * 8);
    w.clear_bytes = o;
    w.bt_idx = reinterpret_cast<int*>(p + o);                      o += up((size_t)B * P * 4);
    w.bt_ov = reinterpret_cast<float*>(p + o);                     o += up((size_t)B * P * 4);
    w.partial = reinterpret_cast<double*>(p + o);                  o += up((size_t)B * kReduceSplit * 2 * sizeof(double));
    w.total = o;
    return w;
}

size_t rd_multibox_criterion_workspace_bytes(int B, int P, int Gmax) {
    if (B <= 0 || P <= 0 || Gmax <= 0) return 0;
    return carve_criterion(nullptr, B, P, Gmax).total;
}

int rd_multibox_criterion(const float* truths, const float* labels, const int* gt_count, const float* priors,
                          const float* arm_loc, const float* loc_data, const float* conf_data, const float* arm_conf_gate,
                          int B, int P, int C, int Gmax, float threshold, float v0, float v1, int label_mode, float theta,
                          int negpos_ratio, void* workspace, size_t workspace_bytes, float* loc_t, long long* conf_t,
                          float* ce, float* lse, unsigned char* pos, unsigned char* neg, int* num_pos, float* losses,
                          void* stream) {
    NvtxRange nvtx_range("rd_multibox_criterion");
    if (!truths || !labels || !gt_count || !priors || !loc_data || !conf_data || !workspace || !loc_t || !conf_t || !ce ||
        !lse || !pos || !neg || !num_pos || !losses)
        return RD_ERR_BAD_ARG;
    if (B <= 0 || P <= 0 || Gmax <= 0 || C < 2 || label_mode < 0 || label_mode > 2 || negpos_ratio < 0) return RD_ERR_BAD_ARG;
    if (Gmax > RD_MAX_GT || B > 65535 || C > kLossMaxClasses) return RD_ERR_UNSUPPORTED;
    if ((((uintptr_t)truths | (uintptr_t)priors | (uintptr_t)loc_t | (uintptr_t)loc_data | (uintptr_t)workspace) & 15) ||
        (arm_loc && ((uintptr_t)arm_loc & 15)) || ((uintptr_t)conf_t & 7) || (arm_conf_gate && ((uintptr_t)arm_conf_gate & 7)) ||
        (C == 2 && ((uintptr_t)conf_data & 7)))
        return RD_ERR_ALIGNMENT;
    const CriterionWs w = carve_criterion(workspace, B, P, Gmax);
    if (workspace_bytes < w.total) return RD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, w.clear_bytes, st);
    if (e != cudaSuccess) return (int)e;
    // refine_match / match (box_utils.py:70-160), batched: pass 1 + pass 2
    int rc = match_launch((const float4*)truths, labels, gt_count, (const float4*)priors, (const float4*)arm_loc, B, P, Gmax,
                          threshold, v0, v1, label_mode, w.best_prior, w.bt_ov, w.bt_idx, (float4*)loc_t, conf_t, st, true);
    if (rc != 0) return rc;
    // per-anchor confidence loss, ARM-theta gate of the positives (:96-101, :113-114)
    rc = conf_loss_launch(conf_data, conf_t, arm_conf_gate, theta, (long long)B * P, C, ce, lse, pos, st, true);
    if (rc != 0) return rc;
    // hard-negative mining (:117-123)
    rc = hnm_launch(ce, pos, B, P, negpos_ratio, neg, num_pos, st, true);
    if (rc != 0) return rc;
    // SmoothL1 over pos, cross-entropy over pos | neg, / N (:105-110, :126-138); the last CTA finishes
    e = launch_pdl(loss_reduce_kernel<true>, dim3(kReduceSplit, B), dim3(kLossThreads), 0, st, (const float4*)loc_data,
                   (const float4*)loc_t, (const float*)ce, (const unsigned char*)pos, (const unsigned char*)neg, P, w.partial,
                   w.ticket, (const int*)num_pos, B, losses, losses + 1, losses + 2);
    if (e != cudaSuccess) return (int)e;
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

int rd_multibox_loss_backward(const float* loc, const float* loc_t, const float* conf, const long long* conf_t,
                              const float* lse, const unsigned char* pos, const unsigned char* neg,
                              const float* grad_loss_l, const float* grad_loss_c, const float* n_dev,
                              long long rows, int C, float* grad_loc, float* grad_conf, void* stream) {
    NvtxRange nvtx_range("rd_multibox_loss_backward");
    if (!loc || !loc_t || !conf || !conf_t || !lse || !pos || !neg || !n_dev || rows <= 0 || C < 2)
        return RD_ERR_BAD_ARG;
    if (!grad_loc && !grad_conf) return 0;
    if (C > kLossMaxClasses) return RD_ERR_UNSUPPORTED;
    if (((uintptr_t)loc | (uintptr_t)loc_t | (uintptr_t)grad_loc) & 15) return RD_ERR_ALIGNMENT;
    // wide rows (C >= 48, the COCO head): the zero-stream kernel, 43 us against 45 us for the tile kernel at C = 81; at
    // C = 21 / C = 2 its three phases per CTA are a longer chain than the whole job (33 / 29 us against 26 / 20 us)
    static const char* s_variant = getenv("RD_BWD");          // "tile" / "regs" / "zs": force a variant (A/B runs)
    const bool want_zs = s_variant ? s_variant[0] == 'z' : C >= 48;
    if (((uintptr_t)grad_conf & 15) == 0 && want_zs && !getenv("RD_NO_TMA")) {
        // tiles of about 80 KB: 8 groups of 32 rows at C = 81, 64 at C = 2
        long long gpt = (81920 + 64ll * C) / (128ll * C);
        gpt = gpt < 1 ? 1 : gpt > 64 ? 64 : gpt;
        const long long ntiles = (rows + 32 * gpt - 1) / (32 * gpt);
        static const char* s_per_sm = getenv("RD_BWD_PER_SM");
        const long long per_sm = s_per_sm ? atoll(s_per_sm) : 3;          // resident CTAs per SM (64 KB of shared memory each)
        const long long waves = (ntiles + 148 * per_sm - 1) / (148 * per_sm);
        const long long nb = (ntiles + waves - 1) / waves;                // the same number of tiles for every CTA
        const size_t smem = kZsZeroBytes + kZsStageBytes;
        static size_t s_zs_smem[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(loss_backward_zs_kernel, smem, s_zs_smem);
        if (e != cudaSuccess) return (int)e;
        loss_backward_zs_kernel<<<(unsigned)nb, kZsThreads, smem, (cudaStream_t)stream>>>(
            (const float4*)loc, (const float4*)loc_t, conf, conf_t, lse, pos, neg, grad_loss_l, grad_loss_c, n_dev, rows, C,
            (int)gpt, (float4*)grad_loc, grad_conf);
        note_launch();
        RD_CHECK_LAUNCH();
        return 0;
    }
    if (grad_conf && (C & 1) && ((uintptr_t)grad_conf & 15) == 0 && !getenv("RD_NO_TMA") &&
        !(s_variant && s_variant[0] == 'r')) {
        const size_t smem = (size_t)2 * kBwdRows * C * sizeof(float);
        static size_t s_bwd_smem[kMaxDevices];
        cudaError_t e = ensure_dynamic_smem(loss_backward_tma_kernel, smem, s_bwd_smem);
        if (e != cudaSuccess) return (int)e;
        const long long ntiles = (rows + kBwdRows - 1) / kBwdRows;
        long long per_sm = (long long)(220 * 1024) / (long long)(smem + 1024);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        long long nb = 148 * per_sm;
        if (nb > ntiles) nb = ntiles;
        loss_backward_tma_kernel<<<(unsigned)nb, kBwdThreads, smem, (cudaStream_t)stream>>>(
            (const float4*)loc, (const float4*)loc_t, conf, conf_t, lse, pos, neg, grad_loss_l, grad_loss_c, n_dev, rows, C,
            (float4*)grad_loc, grad_conf);
        note_launch();
        RD_CHECK_LAUNCH();
        return 0;
    }
    long long blocks = ((rows + 31) / 32 * 32 + kLossThreads - 1) / kLossThreads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    loss_backward_kernel<<<(unsigned)blocks, kLossThreads, 0, (cudaStream_t)stream>>>(
        (const float4*)loc, (const float4*)loc_t, conf, conf_t, lse, pos, neg, grad_loss_l, grad_loss_c, n_dev, rows, C,
        (float4*)grad_loc, grad_conf);
    note_launch();
    RD_CHECK_LAUNCH();
    return 0;
}

// ---- both criteria of a training step as one call ----------------------------------------------------------------
struct CriterionState {
    float* loc_t; long long* conf_t; float* ce; float* lse; unsigned char* pos; unsigned char* neg; int* num_pos;
    float* losses; void* ws;
    size_t ws_bytes, total;
    size_t off[9];
};
static CriterionState carve_state(void* base, int B, int P, int Gmax) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t n = (size_t)B * P;
    const size_t widths[6] = {16, 8, 4, 4, 1, 1};
    CriterionState s;
    size_t o = 0;
    for (int i = 0; i < 6; ++i) { s.off[i] = o; o += up(n * widths[i]); }
    s.off[6] = o; o += up((size_t)B * 4);
    s.off[7] = o; o += 256;
    s.off[8] = o;
    s.ws_bytes = carve_criterion(nullptr, B, P, Gmax).total;
    s.total = o + s.ws_bytes;
    unsigned char* p = static_cast<unsigned char*>(base);
    s.loc_t = reinterpret_cast<float*>(p + s.off[0]);
    s.conf_t = reinterpret_cast<long long*>(p + s.off[1]);
    s.ce = reinterpret_cast<float*>(p + s.off[2]);
    s.lse = reinterpret_cast<float*>(p + s.off[3]);
    s.pos = p + s.off[4];
    s.neg = p + s.off[5];
    s.num_pos = reinterpret_cast<int*>(p + s.off[6]);
    s.losses = reinterpret_cast<float*>(p + s.off[7]);
    s.ws = p + s.off[8];
    return s;
}

// fork / join of the side stream: two events per (thread, device), created once.  Re-recording an event does not
// disturb a wait that was enqueued on its earlier record, so the pair is reused by every call of the thread.
struct ForkJoin {
    cudaEvent_t fork = nullptr, join = nullptr;
};
static cudaError_t fork_join_events(ForkJoin** out) {
    static thread_local ForkJoin tl[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    ForkJoin& fj = tl[(unsigned)dev % kMaxDevices];
    if (!fj.fork) {
        e = cudaEventCreateWithFlags(&fj.fork, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
        e = cudaEventCreateWithFlags(&fj.join, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    *out = &fj;
    return cudaSuccess;
}

size_t rd_criterion_state_bytes(int B, int P, int Gmax) {
    if (B <= 0 || P <= 0 || Gmax <= 0) return 0;
    return carve_state(nullptr, B, P, Gmax).total;
}

int rd_criterion_state_layout(int B, int P, int Gmax, size_t* offsets) {
    if (B <= 0 || P <= 0 || Gmax <= 0 || !offsets) return RD_ERR_BAD_ARG;
    const CriterionState s = carve_state(nullptr, B, P, Gmax);
    for (int i = 0; i < 9; ++i) offsets[i] = s.off[i];
    return 0;
}

int rd_multibox_criterion_pair(const float* truths, const float* labels, const int* gt_count, const float* priors,
                               const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                               int B, int P, int C, int Gmax, float arm_threshold, float odm_threshold, float v0, float v1,
                               int arm_label_mode, float theta, int arm_negpos_ratio, int odm_negpos_ratio,
                               void* arm_state, void* odm_state, size_t state_bytes, void* stream, void* side_stream) {
    NvtxRange nvtx_range("rd_multibox_criterion_pair");
    if (!arm_loc || !arm_conf || !odm_loc || !odm_conf || !arm_state || !odm_state || B <= 0 || P <= 0 || Gmax <= 0)
        return RD_ERR_BAD_ARG;
    if (((uintptr_t)arm_state | (uintptr_t)odm_state) & 255) return RD_ERR_ALIGNMENT;
    const CriterionState a = carve_state(arm_state, B, P, Gmax), o = carve_state(odm_state, B, P, Gmax);
    if (state_bytes < a.total) return RD_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream, side = (cudaStream_t)side_stream;
    const bool two = side && side != st;
    ForkJoin* fj = nullptr;
    if (two) {
        cudaError_t e = fork_join_events(&fj);
        if (e == cudaSuccess) e = cudaEventRecord(fj->fork, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fj->fork, 0);
        if (e != cudaSuccess) return (int)e;
    }
    // the ODM chain is the long one (the C-class confidence loss): it is issued first
    int rc = rd_multibox_criterion(truths, labels, gt_count, priors, arm_loc, odm_loc, odm_conf, arm_conf, B, P, C, Gmax,
                                   odm_threshold, v0, v1, 0 /* labels as they are */, theta, odm_negpos_ratio, o.ws, o.ws_bytes,
                                   o.loc_t, o.conf_t, o.ce, o.lse, o.pos, o.neg, o.num_pos, o.losses, two ? side : st);
    int rc2 = rd_multibox_criterion(truths, labels, gt_count, priors, nullptr, arm_loc, arm_conf, nullptr, B, P, 2, Gmax,
                                    arm_threshold, v0, v1, arm_label_mode, theta, arm_negpos_ratio, a.ws, a.ws_bytes, a.loc_t,
                                    a.conf_t, a.ce, a.lse, a.pos, a.neg, a.num_pos, a.losses, st);
    if (two) {                                               // joined even after an error: `stream` stays ordered behind `side`
        cudaError_t e = cudaEventRecord(fj->join, side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, fj->join, 0);
        if (e != cudaSuccess && rc == 0 && rc2 == 0) return (int)e;
    }
    return rc != 0 ? rc : rc2;
}

int rd_multibox_loss_backward_pair(const float* arm_loc, const float* arm_conf, const float* odm_loc, const float* odm_conf,
                                   const void* arm_state, const void* odm_state, int B, int P, int C, int Gmax,
                                   const float* g_arm_l, const float* g_arm_c, const float* g_odm_l, const float* g_odm_c,
                                   float* grad_arm_loc, float* grad_arm_conf, float* grad_odm_loc, float* grad_odm_conf,
                                   void* stream, void* side_stream) {
    NvtxRange nvtx_range("rd_multibox_loss_backward_pair");
    if (!arm_loc || !arm_conf || !odm_loc || !odm_conf || !arm_state || !odm_state || B <= 0 || P <= 0 || Gmax <= 0)
        return RD_ERR_BAD_ARG;
    const CriterionState a = carve_state(const_cast<void*>(arm_state), B, P, Gmax);
    const CriterionState o = carve_state(const_cast<void*>(odm_state), B, P, Gmax);
    cudaStream_t st = (cudaStream_t)stream, side = (cudaStream_t)side_stream;
    const bool need_arm = grad_arm_loc || grad_arm_conf, need_odm = grad_odm_loc || grad_odm_conf;
    const bool two = side && side != st && need_arm && need_odm;
    ForkJoin* fj = nullptr;
    if (two) {
        cudaError_t e = fork_join_events(&fj);
        if (e == cudaSuccess) e = cudaEventRecord(fj->fork, st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(side, fj->fork, 0);
        if (e != cudaSuccess) return (int)e;
    }
    const long long rows = (long long)B * P;
    int rc = 0, rc2 = 0;
    if (need_odm)
        rc = rd_multibox_loss_backward(odm_loc, o.loc_t, odm_conf, o.conf_t, o.lse, o.pos, o.neg, g_odm_l, g_odm_c, o.losses + 2,
                                       rows, C, grad_odm_loc, grad_odm_conf, st);
    if (need_arm)
        rc2 = rd_multibox_loss_backward(arm_loc, a.loc_t, arm_conf, a.conf_t, a.lse, a.pos, a.neg, g_arm_l, g_arm_c, a.losses + 2,
                                        rows, 2, grad_arm_loc, grad_arm_conf, two ? side : st);
    if (two) {
        cudaError_t e = cudaEventRecord(fj->join, side);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, fj->join, 0);
        if (e != cudaSuccess && rc == 0 && rc2 == 0) return (int)e;
    }
    return rc != 0 ? rc : rc2;
}

}  // extern "C"

// ce_tc2.cu — full-catalog soft-max cross-entropy with BOTH contractions on the tensor cores
// (engine PCV_CE_ENGINE_TF32 of pcv_ce_fwd_bwd / pcv_ce_partials; dense mask, D = 8).
//
// Replaces pivotcvae.py:274 (p = mm(prox, table.t())) + train_generative.py:59 (CrossEntropyLoss) and their
// backward, like ce.cu.  ce_tc.cu (the first tensor-core engine) computed the logits with tcgen05 but accumulated
// d loss / d q = sum_j softmax_j w_j with 8 FFMA per logit on the CUDA cores: FMA-pipe-bound at ~23 cycles per 32
// logits per scheduler.  Here the gradient is a second MMA, flash-attention style:
//
//   S[128 x 128]  = Q[128 x 8] . W_tile^T                    tcgen05.mma kind::tf32, A/B from shared memory
//   P             = exp(S - c_row)                            epilogue warps: TMEM -> registers -> TMEM (in place)
//   dq[128 x 16] += P[128 x 128] . Wt_tile^T                 tcgen05.mma, A = P from TENSOR MEMORY, B = the
//                                                            transposed tile [16 (8 dims + 8 zero rows) x 128 items]
//
// so an epilogue thread spends 3 instructions per logit (FFMA, MUFU.EX2, FADD) instead of 13.  Measured at C3
// (100 k items x 20480 rows): 0.70 ms vs 1.41 ms for ce_tc.cu; the MUFU pipe (one ex2 per logit, 16 / clk / SM) would
// allow 0.44 ms — what is left is the single MMA-issuing warp (17 tcgen05.mma per 128-item tile, see umma_*_elect).  There is no running maximum to rescale the TMEM accumulator by: every row uses a
// FIXED reference exponent c_row chosen by ce_ref_kernel from rigorous bounds —
//   s_j <= ub = |q| max|w| (Cauchy-Schwarz),   lb = max(target logit, logits of the first 32 items) <= max_j s_j,
//   c = max(lb, ub - 60)   =>   s_j - c <= 60 (no overflow: e^60 * N << 3e38) and max_j s_j - c >= lb - ub + 60,
// which cannot underflow when lb >= ub - 140.  Rows that fail that test (never seen in practice: it needs
// |q| max|w| > 70 with a far-off target) are flagged and redone exactly by ce_rowfix_kernel (warp per row).
// The per-(split, row) records (c, l, count, acc[8]) have the layout of ce.cu and are merged by the same
// ce_finalize_kernel (target logit re-derived with the exact fp32 FMA chain).
#include <mutex>

#include "tc_common.cuh"

namespace pcv {

constexpr int C2_BN = 128;                      // items per tile (= TMEM columns of one S/P buffer)
constexpr int C2_BUFS = 3;                      // S/P buffers in flight (columns 0 .. 383)
constexpr int C2_DQ_COL = C2_BUFS * C2_BN;      // dq accumulators: C2_DQ_ACCS x 16 columns from 384
constexpr int C2_DQ_ACCS = 4;                   // N = 16 gradient MMAs are 8 cycles of work each: four independent
                                                // accumulators keep the tensor pipe from waiting on its own accumulate latency
constexpr int C2_EPI_WARPS = 16;                // 4 per TMEM lane quarter, each a 32-column slice
constexpr int C2_THREADS = 64 + 32 * C2_EPI_WARPS;
constexpr int C2_STAGES = 12;
constexpr int C2_W_FLOATS = C2_BN * 8;          // W tile: 128 rows x 32 B (SWIZZLE_32B image), 4 KB
constexpr int C2_WT_FLOATS = 16 * 16 * 8;       // transposed tile: 16 k-atoms x [16 rows x 32 B], 8 KB
constexpr int C2_REC = 3 + 8;

struct __align__(1024) Ce2Smem {
  float w[C2_STAGES][C2_W_FLOATS];
  float wt[C2_STAGES][C2_WT_FLOATS];
  float a[2][TC_BM * 8];
  unsigned long long full[C2_STAGES], empty[C2_STAGES], tfull[C2_BUFS], pfull[C2_BUFS], tempty[C2_BUFS];
  unsigned long long afull[2], aempty[2], dqfull, dqempty;
  float lsum[2][TC_BM];      // per-row sum of l over the column slices 1..3 of a work item (double-buffered by item parity)
  uint32_t tmem_base;
};

// kind::tf32, fp32 accumulate, K-major operands: M = 128, N = 128 (scores) / N = 16 (gradient)
constexpr uint32_t C2_IDESC_S = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C2_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
constexpr uint32_t C2_IDESC_G = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// D[tmem] (+)= A[tmem] . B[smem]^T : the A operand is read from tensor memory (lane = row, one 32-bit column per k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// (the MMA warp runs converged: see umma_tf32_elect in tc_common.cuh)
__device__ __forceinline__ void umma_tf32_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
#define C2_ST32(taddr, v)                                                                                   \
  asm volatile(                                                                                             \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                       \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                            \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                    \
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),  \
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),        \
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),      \
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])       \
      : "memory")

__device__ __forceinline__ float c2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- per-row reference exponent + safety flag (see the header)
__global__ void __launch_bounds__(256)
ce_ref_kernel(const float *__restrict__ W, int64_t n_rows, int64_t row_offset, float max_row_norm, const float *__restrict__ Q,
              const int64_t *__restrict__ targets, int64_t M, float *__restrict__ cref, unsigned char *__restrict__ unsafe) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4 q0 = __ldg(reinterpret_cast<const float4 *>(Q + row * 8));
  const float4 q1 = __ldg(reinterpret_cast<const float4 *>(Q + row * 8) + 1);
  auto dot = [&](int64_t j) {
    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(W + j * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(W + j * 8) + 1);
    float s = 0.f;
    s = fmaf(q0.x, w0.x, s); s = fmaf(q0.y, w0.y, s); s = fmaf(q0.z, w0.z, s); s = fmaf(q0.w, w0.w, s);
    s = fmaf(q1.x, w1.x, s); s = fmaf(q1.y, w1.y, s); s = fmaf(q1.z, w1.z, s); s = fmaf(q1.w, w1.w, s);
    return s;
  };
  float lb = (lane < n_rows) ? dot(lane) : -INFINITY;          // lane = one of the first 32 items
  const int64_t t = targets[row] - row_offset;
  if (lane == 0 && t >= 0 && t < n_rows) lb = fmaxf(lb, dot(t));
  lb = warp_max(lb);
  if (lane == 0) {
    float ss = 0.f;
    ss = fmaf(q0.x, q0.x, ss); ss = fmaf(q0.y, q0.y, ss); ss = fmaf(q0.z, q0.z, ss); ss = fmaf(q0.w, q0.w, ss);
    ss = fmaf(q1.x, q1.x, ss); ss = fmaf(q1.y, q1.y, ss); ss = fmaf(q1.z, q1.z, ss); ss = fmaf(q1.w, q1.w, ss);
    const float ub = sqrtf(ss) * max_row_norm * 1.01f + 1e-30f;   // >= every tf32 logit (truncation only shrinks |s|)
    lb -= 0.01f * ub;                                              // the tf32 logit of that item may sit a little lower
    const bool ok = lb >= ub - 140.f && ub < 1.0e30f;
    cref[row] = ok ? fmaxf(lb, ub - 60.f) : 0.f;
    unsafe[row] = ok ? 0 : 1;
  }
}

// ---- exact redo of a flagged row over this table shard (one warp per row; normally every warp exits at once):
// stream 0 of the row gets the exact partial (running max m, l, acc), every other stream an empty record
__global__ void __launch_bounds__(256)
ce_rowfix_kernel(const float *__restrict__ W, int64_t n_rows, const float *__restrict__ Q, int64_t M,
                 const unsigned char *__restrict__ unsafe, int n_streams, float *__restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M || !unsafe[row]) return;
  float q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) q[k] = __ldg(Q + row * 8 + k);
  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int64_t j = lane; j < n_rows; j += 32) {
    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(W + j * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(W + j * 8) + 1);
    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s = fmaf(q[k], w[k], s);
    if (s > m) {
      const float sc = __expf(m - s);
      l *= sc;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= sc;
      m = s;
    }
    const float p = __expf(s - m);
    l += p;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(p, w[k], acc[k]);
  }
  const float mw = warp_max(m);
  const float sc = (m == -INFINITY) ? 0.f : __expf(m - mw);
  l = warp_sum(l * sc);
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = warp_sum(acc[k] * sc);
  if (lane == 0) {
    for (int s = 0; s < n_streams; ++s) {
      float *rec = part + ((int64_t)s * M + row) * C2_REC;
      rec[0] = s == 0 ? mw : -INFINITY;
      rec[1] = s == 0 ? l : 0.f;
      rec[2] = __int_as_float(s == 0 ? (int)n_rows : 0);
#pragma unroll
      for (int k = 0; k < 8; ++k) rec[3 + k] = s == 0 ? acc[k] : 0.f;
    }
  }
}

// Transposed, pre-swizzled image of the table for the gradient MMA's B operand: per 128-item tile, 16 k-atoms (8 items
// each) of [16 rows x 32 B]: row r < 8 = embedding dimension r, rows 8..15 zero; the two 16-byte chunks of a row are
// swapped when bit 2 of r is set (SWIZZLE_32B).  The buffer is zero-filled first (padding rows, tail of the last tile).
__global__ void pack_wt_kernel(const float *__restrict__ W, int64_t n_rows, float *__restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_rows) return;
  const int64_t tile = j / C2_BN;
  const int i = (int)(j % C2_BN), a = i >> 3, c = (i >> 2) & 1, e = i & 3;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    out[((tile * 16 + a) * 16 + r) * 8 + ((c ^ ((r >> 2) & 1)) << 2) + e] = W[j * 8 + r];
}

__global__ void __launch_bounds__(C2_THREADS, 1)
ce_tc2_kernel(const float *__restrict__ Wsw, const float *__restrict__ Wt, int64_t n_rows, const float *__restrict__ Q, int64_t M,
              const float *__restrict__ cref, int64_t items_per_split, int n_split, int n_work, float *__restrict__ part) {
  extern __shared__ unsigned char smem_raw[];
  Ce2Smem &S = *reinterpret_cast<Ce2Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int lane = threadIdx.x & 31;
  // roles: hardware warps 0..15 = epilogue, 16 = producer, 17 = MMA issuer.  The scheduler arbitrates highest warp id
  // first, so the two single-thread control warps are never starved by the MUFU-bound epilogue warps of their
  // sub-partition.  (`warp` keeps the logical numbering used below: 0 = producer, 1 = MMA, 2.. = epilogue.)
  const int hw_warp = threadIdx.x >> 5;
  const int warp = hw_warp >= C2_EPI_WARPS ? hw_warp - C2_EPI_WARPS : hw_warp + 2;

  if (threadIdx.x < 2 * TC_BM) (&S.lsum[0][0])[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C2_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int b = 0; b < C2_BUFS; ++b) {
      mbar_init(&S.tfull[b], 1); mbar_init(&S.pfull[b], C2_EPI_WARPS / 2); mbar_init(&S.tempty[b], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.afull[b], 1); mbar_init(&S.aempty[b], 1); }
    mbar_init(&S.dqfull, 1);
    mbar_init(&S.dqempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  const int w_begin = (int)((int64_t)blockIdx.x * n_work / gridDim.x);
  const int w_end = (int)((int64_t)(blockIdx.x + 1) * n_work / gridDim.x);
  auto item_rows = [&](int w) { return (int64_t)(w / n_split) * TC_BM; };
  auto item_split = [&](int w) { return (w % n_split + (w / n_split) * 3) % n_split; };
  auto item_jb = [&](int w) { return (int64_t)item_split(w) * items_per_split; };
  auto item_tiles = [&](int w) {
    const int64_t jb = item_jb(w), je = min(n_rows, jb + items_per_split);
    return (int)((je - jb + C2_BN - 1) / C2_BN);
  };

  if (warp == 0) {
    // ---------------- producer: query tile of the work item, then its W / Wt tiles ----------------
    uint32_t gt = 0;
    int it = 0;
    for (int w = w_begin; w < w_end; ++w, ++it) {
      const int ab = it & 1;
      mbar_wait(&S.aempty[ab], ((it >> 1) & 1) ^ 1);
      {
        const int64_t row_base = item_rows(w);
#pragma unroll
        for (int i = 0; i < TC_BM / 32; ++i) {
          const int t = lane + 32 * i;
          const int64_t row = row_base + t;
          float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
          if (row < M) {
            q0 = __ldg(reinterpret_cast<const float4 *>(Q + row * 8));
            q1 = __ldg(reinterpret_cast<const float4 *>(Q + row * 8) + 1);
          }
          const int sw = (t >> 2) & 1;
          float4 *dst = reinterpret_cast<float4 *>(S.a[ab] + t * 8);
          dst[0 ^ sw] = q0;
          dst[1 ^ sw] = q1;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.afull[ab]);
      }
      if (lane == 0) {
        const int64_t j_begin = item_jb(w);
        const int n_tiles = item_tiles(w);
        for (int t = 0; t < n_tiles; ++t, ++gt) {
          const int s = gt % C2_STAGES;
          mbar_wait(&S.empty[s], ((gt / C2_STAGES) & 1) ^ 1);
          const int64_t j0 = j_begin + (int64_t)t * C2_BN;
          const uint32_t wbytes = (uint32_t)min((int64_t)(C2_W_FLOATS * 4), (n_rows - j0) * (int64_t)32);
          mbar_expect_tx(&S.full[s], wbytes + C2_WT_FLOATS * 4);
          tma_bulk_load(S.w[s], Wsw + j0 * 8, wbytes, &S.full[s]);            // rows past the table end stay stale: masked
          tma_bulk_load(S.wt[s], Wt + (j0 / C2_BN) * C2_WT_FLOATS, C2_WT_FLOATS * 4, &S.full[s]);   // zero-padded
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: scores of tile t, then the gradient MMAs of tile t - 2 (the two epilogue groups work
    // on tiles t - 1 and t - 2 meanwhile; three S/P buffers) ----------------
    {   // all 32 lanes run this block converged (see umma_*_elect)
      uint32_t gt = 0;
      int it = 0;
      auto grad_mma = [&](uint32_t g, bool first_of_item) {   // dq (+)= P(tile g) . Wt(tile g)^T
        const int buf = g % C2_BUFS, s = g % C2_STAGES;
        mbar_wait(&S.pfull[buf], (g / C2_BUFS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors of the 16 k-atoms differ only in the start address (512 B apart): one base, constant offsets
        const uint64_t bd0 = umma_desc_sw32(S.wt[s]);
        const uint32_t pa0 = tmem + buf * C2_BN;
        const uint32_t acc0 = first_of_item ? 0u : 1u;
#pragma unroll
        for (int a = 0; a < C2_BN / 8; ++a)
          umma_tf32_ts_elect(tmem + C2_DQ_COL + (a % C2_DQ_ACCS) * 16, pa0 + a * 8, bd0 + (uint64_t)(a * (512 >> 4)), C2_IDESC_G,
                             a < C2_DQ_ACCS ? acc0 : 1u);
        umma_commit_elect(&S.tempty[buf]);
        umma_commit_elect(&S.empty[s]);
      };
      for (int w = w_begin; w < w_end; ++w, ++it) {
        const int ab = it & 1;
        const int n_tiles = item_tiles(w);
        mbar_wait(&S.afull[ab], (it >> 1) & 1);
        if (it > 0) mbar_wait(&S.dqempty, (it - 1) & 1);     // the previous item's gradient has been read out
        const uint64_t adesc = umma_desc_sw32(S.a[ab]);
        for (int t = 0; t < n_tiles; ++t, ++gt) {
          const int buf = gt % C2_BUFS, s = gt % C2_STAGES;
          mbar_wait(&S.tempty[buf], ((gt / C2_BUFS) & 1) ^ 1);
          mbar_wait(&S.full[s], (gt / C2_STAGES) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          umma_tf32_elect(tmem + buf * C2_BN, adesc, umma_desc_sw32(S.w[s]), C2_IDESC_S, 0);
          umma_commit_elect(&S.tfull[buf]);
          if (t > 1) grad_mma(gt - 2, t == 2);
        }
        umma_commit_elect(&S.aempty[ab]);          // every score MMA of the item has read the query tile
        if (n_tiles > 1) grad_mma(gt - 2, n_tiles == 2);
        grad_mma(gt - 1, n_tiles == 1);
        umma_commit_elect(&S.dqfull);
      }
    }
  } else {
    // ---------------- epilogue: P = exp(S - c) in place, l += row sums ----------------
    // Two groups of 8 warps ping-pong over the tiles (group g takes the tiles with gt % 2 == g), so one group's TMEM
    // loads / stores / barrier round trips overlap the other group's exponentials.  A warp owns one TMEM lane quarter
    // and one 64-column half of its tiles.
    const int quarter = hw_warp & 3;          // TMEM lane quarter = hardware warp id % 4
    const int slice = hw_warp >> 2;           // record stream of this warp (0..3)
    const int group = slice & 1, half = slice >> 1;
    const int trow = quarter * 32 + lane;
    const float LOG2E = 1.4426950408889634f;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    uint32_t gt = 0;
    int it = 0;
    for (int w = w_begin; w < w_end; ++w, ++it) {
      const int64_t row = item_rows(w) + trow;
      const int64_t j_begin = item_jb(w);
      const int64_t j_end = min(n_rows, j_begin + items_per_split);
      const int n_tiles = item_tiles(w);
      const float nc = (row < M) ? -__ldg(cref + row) * LOG2E : 0.f;
      float l = 0.f;
      int cnt = 0;
      for (int t = 0; t < n_tiles; ++t, ++gt) {
        if ((int)(gt & 1) != group) continue;
        const int buf = gt % C2_BUFS;
        mbar_wait(&S.tfull[buf], (gt / C2_BUFS) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t col0 = j_begin + (int64_t)t * C2_BN + half * 64;
        const int n_valid = (int)max((int64_t)0, min((int64_t)64, j_end - col0));
        cnt += n_valid;
        const uint32_t taddr = lane_base + (uint32_t)(buf * C2_BN + half * 64);
        uint32_t va[32], vb[32];
        TC_LD32(va, taddr);
        TC_LD32(vb, taddr + 32);
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        auto chunk = [&](uint32_t (&v)[32], int nv) {
          if (nv == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float p0 = c2_ex2(fmaf(__uint_as_float(v[i]), LOG2E, nc));
              const float p1 = c2_ex2(fmaf(__uint_as_float(v[i + 1]), LOG2E, nc));
              const float p2 = c2_ex2(fmaf(__uint_as_float(v[i + 2]), LOG2E, nc));
              const float p3 = c2_ex2(fmaf(__uint_as_float(v[i + 3]), LOG2E, nc));
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              v[i] = __float_as_uint(p0); v[i + 1] = __float_as_uint(p1);
              v[i + 2] = __float_as_uint(p2); v[i + 3] = __float_as_uint(p3);
            }
          } else {     // last tile of the table: stale scores beyond its end contribute nothing
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float p0 = c2_ex2(fmaf(__uint_as_float(v[i]), LOG2E, nc));
              if (i >= nv) p0 = 0.f;
              l0 += p0;
              v[i] = __float_as_uint(p0);
            }
          }
        };
        TC_WAIT_LD(va);
        chunk(va, min(32, n_valid));
        C2_ST32(taddr, va);
        TC_WAIT_LD(vb);
        chunk(vb, max(0, n_valid - 32));
        C2_ST32(taddr + 32, vb);
        l += (l0 + l1) + (l2 + l3);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.pfull[buf]);
      }
      // the item's gradient: the slice-0 warps read the 128 x 16 accumulator (8 real columns) once every MMA landed
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      if (slice == 0) {
        mbar_wait(&S.dqfull, it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int j = 0; j < C2_DQ_ACCS; ++j) {
          uint32_t g[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]), "=r"(g[4]), "=r"(g[5]), "=r"(g[6]), "=r"(g[7])
                       : "r"(lane_base + (uint32_t)(C2_DQ_COL + j * 16)));
          asm volatile("tcgen05.wait::ld.sync.aligned;"
                       : "+r"(g[0]), "+r"(g[1]), "+r"(g[2]), "+r"(g[3]), "+r"(g[4]), "+r"(g[5]), "+r"(g[6]), "+r"(g[7])::"memory");
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += __uint_as_float(g[k]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.dqempty);
      }
      // ONE record per (split, row): the other three column slices add their l into shared memory, a named barrier
      // over the four warps of this TMEM lane quarter orders it, the slice-0 warp writes the record
      float *ls = &S.lsum[it & 1][trow];
      if (slice != 0) atomicAdd(ls, l);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + quarter) : "memory");
      if (slice == 0) {
        l += *ls;
        *ls = 0.f;             // this buffer is used again two items later, behind the next item's barrier
        if (row < M) {
          float *rec = part + ((int64_t)item_split(w) * M + row) * C2_REC;
          rec[0] = __ldg(cref + row);
          rec[1] = l;
          rec[2] = __int_as_float((int)(j_end - j_begin));
#pragma unroll
          for (int k = 0; k < 8; ++k) rec[3 + k] = acc[k];
        }
      }
      (void)cnt;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
  }
}

// ------------------------------------------------------------------ host side
struct Ce2Plan {
  int n_split;
  int64_t items_per_split;
  size_t rec_bytes, ws_bytes;   // records; records + cref + flags
};

static void ce2_plan(const Table *t, int64_t M, Ce2Plan *p) {
  const int row_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int64_t tiles = (t->n_rows + C2_BN - 1) / C2_BN;
  int64_t max_split = tiles / 16;
  if (max_split < 1) max_split = 1;
  if (max_split > 32) max_split = 32;
  int64_t best_ns = 1;
  double best_eff = -1.0;
  for (int64_t ns = 1; ns <= max_split; ++ns) {
    const int64_t tps = (tiles + ns - 1) / ns;
    const int64_t real_ns = (tiles + tps - 1) / tps;
    const int64_t ctas = (int64_t)row_tiles * real_ns;
    const int64_t waves = (ctas + t->sm_count - 1) / t->sm_count;
    const double eff = (double)ctas / (double)(waves * t->sm_count);
    if (eff > best_eff + 0.03) { best_eff = eff; best_ns = real_ns; }
  }
  const int64_t tps = (tiles + best_ns - 1) / best_ns;
  p->n_split = (int)((tiles + tps - 1) / tps);
  p->items_per_split = tps * C2_BN;
  p->rec_bytes = ((size_t)p->n_split * (size_t)M * C2_REC * sizeof(float) + 255) & ~(size_t)255;   // one record per (split, row)
  p->ws_bytes = p->rec_bytes + (((size_t)M * 4 + 255) & ~(size_t)255) + (((size_t)M + 255) & ~(size_t)255);
}

bool ce_tc_supported(const Table *t) { return t->dim == TC_D && t->tmap_valid; }

size_t ce_tc_workspace(const Table *t, int64_t M) {
  if (!ce_tc_supported(t)) return 0;
  Ce2Plan p;
  ce2_plan(t, M, &p);
  return p.ws_bytes;
}

// The transposed image is built on first use (one-off, like the table handle's other packed copy) and owned by it.
static int ensure_packed_t(Table *t) {
  static std::mutex mu;                 // two host threads may hit the first tensor-core CE call of a handle together
  std::lock_guard<std::mutex> lock(mu);
  if (t->packed_t) return PCV_OK;
  const int64_t tiles = (t->n_rows + C2_BN - 1) / C2_BN;
  const size_t bytes = (size_t)tiles * C2_WT_FLOATS * sizeof(float);
  float *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("ce(tf32): cudaMalloc of the transposed table image (%zu bytes) -> %s", bytes, cudaGetErrorString(e));
    return PCV_ERR_CUDA;
  }
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) {
    pack_wt_kernel<<<(unsigned)((t->n_rows + 255) / 256), 256>>>(t->W, t->n_rows, p);
    count_launch();
    e = cudaDeviceSynchronize();
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    set_error("ce(tf32): packing the transposed table image -> %s", cudaGetErrorString(e));
    return PCV_ERR_CUDA;
  }
  t->packed_t = p;
  return PCV_OK;
}

void table_free_ce(Table *t) {
  if (t->packed_t) cudaFree(t->packed_t);
  t->packed_t = nullptr;
}

// -> number of partial records per row (streams), < 0 on error
int ce_tc_launch(const Table *tc, const float *Q, const int64_t *targets, int64_t M, float *part, size_t ws_bytes, cudaStream_t st) {
  Table *t = const_cast<Table *>(tc);
  Ce2Plan p;
  ce2_plan(t, M, &p);
  if (ws_bytes < p.ws_bytes) {
    set_error("ce(tf32): workspace too small (%zu < %zu)", ws_bytes, p.ws_bytes);
    return PCV_ERR_WORKSPACE;
  }
  if (!t->packed_t) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (cs != cudaStreamCaptureStatusNone) {
      set_error("ce(tf32): the first call on a table builds its transposed image and can not run inside a CUDA graph "
                "capture; run one eager step first");
      return PCV_ERR_CUDA;
    }
    const int rc = ensure_packed_t(t);
    if (rc != PCV_OK) return rc;
  }
  float *cref = reinterpret_cast<float *>(reinterpret_cast<char *>(part) + p.rec_bytes);
  unsigned char *unsafe = reinterpret_cast<unsigned char *>(cref) + (((size_t)M * 4 + 255) & ~(size_t)255);
  const int n_streams = p.n_split;
  ce_ref_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(t->W, t->n_rows, t->row_offset, t->max_row_norm, Q, targets, M, cref, unsafe);
  PCV_LAUNCH_CHECK();
  const size_t smem = sizeof(Ce2Smem) + 1024;
  static bool attr_set[64] = {false};
  if (!attr_set[t->device & 63]) {
    PCV_CUDA(cudaFuncSetAttribute(ce_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[t->device & 63] = true;
  }
  const int64_t n_work = ((M + TC_BM - 1) / TC_BM) * (int64_t)p.n_split;
  const unsigned grid = (unsigned)(n_work < t->sm_count ? n_work : t->sm_count);
  ce_tc2_kernel<<<grid, C2_THREADS, smem, st>>>(t->packed, t->packed_t, t->n_rows, Q, M, cref, p.items_per_split, p.n_split,
                                                 (int)n_work, part);
  PCV_LAUNCH_CHECK();
  ce_rowfix_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(t->W, t->n_rows, Q, M, unsafe, n_streams, part);
  PCV_LAUNCH_CHECK();
  return n_streams;
}

}  // namespace pcv

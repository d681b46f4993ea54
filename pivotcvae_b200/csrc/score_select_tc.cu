// score_select_tc.cu — greedy score+select on the 5th-gen tensor cores (sm_100a).
//
// Same contract as score_select.cu (cvae.py:97-101, pivotcvae.py:191): bit-exact
// arg-max of the fp32 sequential-k FMA chain, ties -> lowest index.  The tensor
// cores only FILTER:
//
//   1. TMA (cp.async.bulk.tensor, SWIZZLE_32B) streams 256-item table tiles
//      ([item][8 x fp32] = 32 B rows, K-major) into a 4-stage shared-memory ring.
//   2. One elected thread issues tcgen05.mma.kind::tf32, M=128 (query rows, staged
//      once per CTA), N=256, K=8 per tile; accumulators ping-pong between two
//      256-column TMEM buffers (512 columns = all of TMEM).
//   3. Eight epilogue warps read their TMEM lanes (tcgen05.ld 32x32b.x32: thread =
//      query row), keep a running max r of the APPROXIMATE scores and record every
//      item whose approximate score >= r - band in a per-row shared-memory list.
//      band = 2*eps, eps = 1.25 * 2^-9 * |q|_2 * max_j |w_j|_2 bounds |tf32 - fp32 chain|
//      (both operands truncated to 10 mantissa bits: relative 2^-10 each), so the
//      exact arg-max (and every exact tie) is always in the list.
//   4. Whenever a list fills up, and at the end, the thread COLLAPSES it: it
//      re-scores the listed items with the exact fp32 FMA chain from the fp32
//      table and keeps the exact winner (ties -> lowest index) in registers.
//      The shared finalize kernel merges the per-split winners.
//
// The (M x N) logits never leave TMEM, and the result is exact for any input
// (duplicated rows and exact ties just collapse more often).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "pcv_common.cuh"

namespace pcv {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_D = 8;
constexpr int TC_STAGES = 4;
constexpr int TC_CAP = 16;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;  // warp0 TMA, warp1 MMA/TMEM, 8 epilogue warps
constexpr uint32_t TC_TILE_BYTES = TC_BN * TC_D * 4;

struct __align__(1024) TcSmem {
  float b[TC_STAGES][TC_BN * TC_D];            // SWIZZLE_32B tiles written by TMA
  float a[TC_BM * TC_D];                       // query tile, same layout
  int32_t cand[2][TC_CAP][TC_BM];              // candidate item indices per (column half, row)
  unsigned long long full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, void *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// K-major, SWIZZLE_32B shared-memory matrix descriptor: rows of 32 B, 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw32(const void *smem) {
  uint64_t d = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(256 >> 4) << 32; // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)6 << 61;          // LayoutType::SWIZZLE_32B
  return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=256
constexpr uint32_t TC_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(void *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

#define TC_LD32(v, taddr)                                                                                   \
  asm volatile(                                                                                             \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                             \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),            \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),          \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),          \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                               \
      : "r"(taddr))

// The wait names the destination registers as in/out operands so the compiler
// cannot move any use of them above the wait.
#define TC_WAIT_LD(v)                                                                                       \
  asm volatile("tcgen05.wait::ld.sync.aligned;"                                                             \
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),        \
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]),    \
                 "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), \
                 "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), \
                 "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])::"memory")

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
  float m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  float a = max3(m[0], m[1], m[2]), b = max3(m[3], m[4], m[5]), c = max3(m[6], m[7], m[8]);
  return max3(max3(a, b, c), m[9], m[10]);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
score_select_tc_kernel(const __grid_constant__ CUtensorMap tmapW, const float *__restrict__ W, int64_t n_rows,
                       const float *__restrict__ Q, int64_t M, int64_t items_per_split, float band_scale,
                       float *__restrict__ part_val, int32_t *__restrict__ part_idx) {
  extern __shared__ unsigned char smem_raw[];
  TcSmem &S = *reinterpret_cast<TcSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row_base = (int64_t)blockIdx.x * TC_BM;
  const int64_t j_begin = (int64_t)blockIdx.y * items_per_split;
  const int64_t j_end = min(n_rows, j_begin + items_per_split);
  const int n_tiles = (int)((j_end - j_begin + TC_BN - 1) / TC_BN);

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&S.tfull[b], 1); mbar_init(&S.tempty[b], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 2 && warp < 6) {
    // stage the 128 query rows: row t at byte t*32, 16-byte chunk c at ((c ^ bit2(t)) * 16)
    const int t = threadIdx.x - 64;
    const int64_t row = row_base + t;
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
    if (row < M) {
      q0 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D));
      q1 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D) + 1);
    }
    const int sw = (t >> 2) & 1;
    float4 *dst = reinterpret_cast<float4 *>(S.a + t * TC_D);
    dst[0 ^ sw] = q0;
    dst[1 ^ sw] = q1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % TC_STAGES;
        const uint32_t ph = (t / TC_STAGES) & 1;
        mbar_wait(&S.empty[s], ph ^ 1);
        mbar_expect_tx(&S.full[s], TC_TILE_BYTES);
        tma_load_2d(S.b[s], &tmapW, 0, (int)(j_begin + (int64_t)t * TC_BN), &S.full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint64_t adesc = umma_desc_sw32(S.a);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % TC_STAGES;
        const uint32_t ph = (t / TC_STAGES) & 1;
        const int buf = t & 1;
        const uint32_t bph = (t >> 1) & 1;
        mbar_wait(&S.tempty[buf], bph ^ 1);
        mbar_wait(&S.full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        umma_tf32(tmem + buf * TC_BN, adesc, umma_desc_sw32(S.b[s]), TC_IDESC, 0);
        umma_commit(&S.empty[s]);
        umma_commit(&S.tfull[buf]);
      }
    }
  } else {
    // ---------------- epilogue: thread = (query row, column half) ----------------
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;         // which 128 columns of every 256-column tile
    const int trow = quarter * 32 + lane;
    const int64_t row = row_base + trow;
    const bool live = row < M;
    float q[TC_D];
    {
      float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
      if (live) {
        q0 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D));
        q1 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D) + 1);
      }
      q[0] = q0.x; q[1] = q0.y; q[2] = q0.z; q[3] = q0.w; q[4] = q1.x; q[5] = q1.y; q[6] = q1.z; q[7] = q1.w;
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < TC_D; ++k) ss = fmaf(q[k], q[k], ss);
    const float band = band_scale * sqrtf(ss) * 1.0001f + 1e-37f;
    float r = live ? -INFINITY : INFINITY;   // dummy rows never record anything
    float thr = r;
    int cnt = 0;
    float best = -INFINITY;                   // exact winner so far
    int32_t bidx = 0x7fffffff;
    int32_t *list = &S.cand[half][0][trow];   // entry e at list[e * TC_BM]

    // exact fp32 re-score of the listed candidates (sequential-k FMA chain)
    auto collapse = [&]() {
      for (int e = 0; e < cnt; ++e) {
        const int32_t j = list[e * TC_BM];
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(W + (int64_t)j * TC_D));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(W + (int64_t)j * TC_D) + 1);
        float s = 0.f;
        s = fmaf(q[0], w0.x, s); s = fmaf(q[1], w0.y, s); s = fmaf(q[2], w0.z, s); s = fmaf(q[3], w0.w, s);
        s = fmaf(q[4], w1.x, s); s = fmaf(q[5], w1.y, s); s = fmaf(q[6], w1.z, s); s = fmaf(q[7], w1.w, s);
        if (s > best || (s == best && j < bidx)) { best = s; bidx = j; }
      }
      cnt = 0;
    };

    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const uint32_t bph = (t >> 1) & 1;
      mbar_wait(&S.tfull[buf], bph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int64_t tile_j0 = j_begin + (int64_t)t * TC_BN + half * (TC_BN / 2);
      const int n_valid = (int)min((int64_t)(TC_BN / 2), j_end - tile_j0);  // may be <= 0
      const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN + half * (TC_BN / 2));
#pragma unroll 1
      for (int c = 0; c < (TC_BN / 2) / 32; ++c) {
        uint32_t v[32];
        TC_LD32(v, taddr + c * 32);
        TC_WAIT_LD(v);
        const int col0 = c * 32;
        if (col0 + 32 > n_valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i >= n_valid) v[i] = 0xff800000u;  // -inf: zero-filled / foreign columns never win
        }
        const float m = chunk_max(v);
        if (m >= thr && m > -INFINITY) {
          r = fmaxf(r, m);
          thr = r - band;
          uint32_t pass = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) pass |= (__uint_as_float(v[i]) >= thr) ? (1u << i) : 0u;
          const int32_t jb = (int32_t)(tile_j0 + col0);
          while (pass) {
            const int i = __ffs(pass) - 1;
            pass &= pass - 1;
            if (cnt == TC_CAP) collapse();
            list[cnt * TC_BM] = jb + i;
            ++cnt;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tempty[buf]);
    }
    collapse();
    if (live) {
      const int64_t slot = ((int64_t)blockIdx.y * 2 + half) * M + row;
      part_val[slot] = best;
      part_idx[slot] = bidx;
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
struct TcPlan {
  int row_tiles, n_split;
  int64_t items_per_split;
  size_t ws_bytes;
};

static void tc_plan(const Table *t, int64_t M, TcPlan *p) {
  p->row_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int64_t tiles = (t->n_rows + TC_BN - 1) / TC_BN;
  int64_t max_split = tiles / 8;  // keep >= 8 tiles per CTA to amortise the prologue
  if (max_split < 1) max_split = 1;
  if (max_split > 64) max_split = 64;
  int64_t best_ns = 1;
  double best_eff = -1.0;
  for (int64_t ns = 1; ns <= max_split; ++ns) {
    const int64_t tps = (tiles + ns - 1) / ns;
    const int64_t real_ns = (tiles + tps - 1) / tps;
    const int64_t ctas = (int64_t)p->row_tiles * real_ns;
    const int64_t waves = (ctas + t->sm_count - 1) / t->sm_count;
    const double eff = (double)ctas / (double)(waves * t->sm_count);
    if (eff > best_eff + 0.03) { best_eff = eff; best_ns = real_ns; }
  }
  const int64_t tps = (tiles + best_ns - 1) / best_ns;
  p->n_split = (int)((tiles + tps - 1) / tps);
  p->items_per_split = tps * TC_BN;
  p->ws_bytes = (size_t)2 * p->n_split * (size_t)M * (sizeof(float) + sizeof(int32_t));
}

bool score_select_tc_supported(const Table *t) { return t->dim == TC_D && t->tmap_valid; }

size_t score_select_tc_workspace(const Table *t, int64_t M) {
  if (t->dim != TC_D) return 0;
  TcPlan p;
  tc_plan(t, M, &p);
  return p.ws_bytes;
}

// Build the TMA descriptor for the table ([n_rows][8] fp32, 32 B rows, SWIZZLE_32B, box 8 x 256).
int table_init_tc(Table *t) {
  t->tmap_valid = 0;
  if (t->dim != TC_D) return PCV_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    cudaGetLastError();
    return PCV_OK;  // engine stays unavailable; SIMT engine serves
  }
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUtensorMap *map = reinterpret_cast<CUtensorMap *>(t->tmap);
  cuuint64_t gdim[2] = {(cuuint64_t)TC_D, (cuuint64_t)t->n_rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(TC_D * sizeof(float))};
  cuuint32_t box[2] = {(cuuint32_t)TC_D, (cuuint32_t)TC_BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(t->W), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) t->tmap_valid = 1;
  return PCV_OK;
}

void launch_select_finalize(const float *pv, const int32_t *pi, int n_parts, int64_t M, int64_t row_offset,
                            int64_t *out_idx, float *out_val, cudaStream_t st);

int score_select_tc(const Table *t, const float *Q, int64_t M, int64_t *out_idx, float *out_val, void *ws,
                    size_t ws_bytes, cudaStream_t st) {
  TcPlan p;
  tc_plan(t, M, &p);
  if (ws == nullptr || ws_bytes < p.ws_bytes) {
    set_error("score_select(tcgen05): workspace too small (%zu < %zu)", ws_bytes, p.ws_bytes);
    return PCV_ERR_WORKSPACE;
  }
  float *pv = reinterpret_cast<float *>(ws);
  int32_t *pi = reinterpret_cast<int32_t *>(pv + (size_t)2 * p.n_split * M);
  const size_t smem = sizeof(TcSmem) + 1024;
  static bool attr_set[64] = {false};
  if (!attr_set[t->device & 63]) {
    PCV_CUDA(cudaFuncSetAttribute(score_select_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[t->device & 63] = true;
  }
  // |tf32 chain - fp32 chain| <= 1.25 * 2^-9 * |q| * max|w|; the band is twice that
  const float band_scale = 2.0f * 1.25f * 0.001953125f * t->max_row_norm;
  dim3 grid((unsigned)p.row_tiles, (unsigned)p.n_split);
  score_select_tc_kernel<<<grid, TC_THREADS, smem, st>>>(*reinterpret_cast<const CUtensorMap *>(t->tmap), t->W,
                                                         t->n_rows, Q, M, p.items_per_split, band_scale, pv, pi);
  PCV_LAUNCH_CHECK();
  launch_select_finalize(pv, pi, 2 * p.n_split, M, t->row_offset, out_idx, out_val, st);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // namespace pcv

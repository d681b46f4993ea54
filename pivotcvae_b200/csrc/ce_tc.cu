// ce_tc.cu — full-catalog soft-max cross-entropy with the logits on the tensor cores
// (opt-in engine PCV_CE_ENGINE_TF32 of pcv_ce_fwd_bwd; dense mask, D = 8).
//
// Replaces, like ce.cu: pivotcvae.py:274 (p = mm(prox, table.t())) + train_generative.py:59
// (CrossEntropyLoss) and their backward.  Same persistent pipeline as the score+select
// filter (TMA bulk ring of pre-swizzled 256-item tiles -> tcgen05.mma.kind::tf32 M=128 x
// N=256 x K=8 -> two TMEM buffers), but the epilogue is a streaming soft-max:
//   per thread (= query row, column slice): running max m, sum l = sum e^{s-m} and the
//   un-normalised gradient acc[k] = sum e^{s-m} * w_j[k]; the w rows are re-read from the
//   very shared-memory tile the MMA consumed (warp-broadcast LDS.128), so a table stage is
//   released only when both the MMA and all epilogue warps are done with it.
// The per-(split, slice) partial records (m, l, count, acc[8]) have the layout of ce.cu and
// are merged by the same ce_finalize_kernel (target logit re-derived with the exact fp32
// FMA chain).  Logits are tf32 products (|err| <= 2^-9 |q||w|): loss within ~1e-4 relative,
// dq within ~1e-3 — the north_star's reduced-precision tolerance (1e-2), hence opt-in; the
// exact-fp32 SIMT kernel stays the default / parity engine.
#include "tc_common.cuh"

namespace pcv {

struct __align__(1024) CeTcSmem {
  float b[TC_STAGES][TC_BN * TC_D];            // SWIZZLE_32B tiles written by TMA (exact fp32 bits of W)
  float a[2][TC_BM * TC_D];                    // query tiles (double-buffered across work items)
  unsigned long long full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2], afull[2], aempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// explicit shared-space 128-bit load (a generic-pointer load would go through the global LSU path)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__device__ __forceinline__ float chunk_max32(const uint32_t (&v)[32]) {
  float g[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    g[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
  g[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
  const float a = max3(g[0], g[1], g[2]), b = max3(g[3], g[4], g[5]), c = max3(g[6], g[7], g[8]);
  return max3(max3(a, b, c), g[9], g[10]);
}

constexpr int CE_TC_REC = 3 + TC_D;   // m, l, count, acc[D]  (== ce.cu's partial record)

__global__ void __launch_bounds__(TC_THREADS, 1)
ce_tc_kernel(const float *__restrict__ Wsw, int64_t n_rows, const float *__restrict__ Q, int64_t M,
             int64_t items_per_split, int n_split, int n_work, float *__restrict__ part) {
  extern __shared__ unsigned char smem_raw[];
  CeTcSmem &S = *reinterpret_cast<CeTcSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1 + TC_EPI_WARPS);   // the MMA's commit + every epilogue warp (they re-read the tile)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.tfull[b], 1); mbar_init(&S.tempty[b], TC_EPI_WARPS);
      mbar_init(&S.afull[b], 1); mbar_init(&S.aempty[b], 1);
    }
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

  if (warp == 0) {
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
            q0 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D));
            q1 = __ldg(reinterpret_cast<const float4 *>(Q + row * TC_D) + 1);
          }
          const int sw = (t >> 2) & 1;
          float4 *dst = reinterpret_cast<float4 *>(S.a[ab] + t * TC_D);
          dst[0 ^ sw] = q0;
          dst[1 ^ sw] = q1;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.afull[ab]);
      }
      if (lane == 0) {
        const int64_t j_begin = item_jb(w);
        const int64_t j_end = min(n_rows, j_begin + items_per_split);
        const int n_tiles = (int)((j_end - j_begin + TC_BN - 1) / TC_BN);
        for (int t = 0; t < n_tiles; ++t, ++gt) {
          const int s = gt % TC_STAGES;
          const uint32_t ph = (gt / TC_STAGES) & 1;
          mbar_wait(&S.empty[s], ph ^ 1);
          const int64_t j0 = j_begin + (int64_t)t * TC_BN;
          const uint32_t bytes = (uint32_t)min((int64_t)TC_TILE_BYTES, (n_rows - j0) * (int64_t)(TC_D * 4));
          mbar_expect_tx(&S.full[s], bytes);
          tma_bulk_load(S.b[s], Wsw + j0 * TC_D, bytes, &S.full[s]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t gt = 0;
      int it = 0;
      for (int w = w_begin; w < w_end; ++w, ++it) {
        const int ab = it & 1;
        const int64_t j_begin = item_jb(w);
        const int64_t j_end = min(n_rows, j_begin + items_per_split);
        const int n_tiles = (int)((j_end - j_begin + TC_BN - 1) / TC_BN);
        mbar_wait(&S.afull[ab], (it >> 1) & 1);
        const uint64_t adesc = umma_desc_sw32(S.a[ab]);
        for (int t = 0; t < n_tiles; ++t, ++gt) {
          const int s = gt % TC_STAGES;
          const uint32_t ph = (gt / TC_STAGES) & 1;
          const int buf = gt & 1;
          const uint32_t bph = (gt >> 1) & 1;
          mbar_wait(&S.tempty[buf], bph ^ 1);
          mbar_wait(&S.full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          umma_tf32(tmem + buf * TC_BN, adesc, umma_desc_sw32(S.b[s]), TC_IDESC, 0);
          umma_commit(&S.empty[s]);
          umma_commit(&S.tfull[buf]);
        }
        umma_commit(&S.aempty[ab]);
      }
    }
  } else {
    // ---------------- epilogue: streaming soft-max + dq accumulation ----------------
    const int quarter = warp & 3;
    const int slice = (warp - 2) >> 2;
    const int trow = quarter * 32 + lane;
    const float LOG2E = 1.4426950408889634f;
    uint32_t gt = 0;
    for (int w = w_begin; w < w_end; ++w) {
      const int64_t row = item_rows(w) + trow;
      const int64_t j_begin = item_jb(w);
      const int64_t j_end = min(n_rows, j_begin + items_per_split);
      const int n_tiles = (int)((j_end - j_begin + TC_BN - 1) / TC_BN);
      float m = -3.0e38f, l = 0.f, acc[TC_D];
#pragma unroll
      for (int k = 0; k < TC_D; ++k) acc[k] = 0.f;
      int cnt = 0;

      // one chunk of 32 logits: rescale on a new maximum (rare), then 32 x (ex2, l +=, 8 FFMA)
      auto process = [&](uint32_t (&v)[32], uint32_t tile, int col0, int lim) {
        const float cm = chunk_max32(v);
        if (cm > m) {
          const float sc = ex2_approx((m - cm) * LOG2E);
          l *= sc;
#pragma unroll
          for (int k = 0; k < TC_D; ++k) acc[k] *= sc;
          m = cm;
        }
        const float nm = -m * LOG2E;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= lim) break;                          // masked tail: never touch stale shared-memory rows
          const float p = ex2_approx(fmaf(__uint_as_float(v[i]), LOG2E, nm));
          l += p;
          const int j = col0 + i;                       // row of the tile (compile-time offset within the slice)
          const int sw = (j >> 2) & 1;
          const float4 w0 = lds128(tile + (uint32_t)(j * TC_D + ((0 ^ sw) << 2)) * 4u);
          const float4 w1 = lds128(tile + (uint32_t)(j * TC_D + ((1 ^ sw) << 2)) * 4u);
          acc[0] = fmaf(p, w0.x, acc[0]); acc[1] = fmaf(p, w0.y, acc[1]); acc[2] = fmaf(p, w0.z, acc[2]);
          acc[3] = fmaf(p, w0.w, acc[3]); acc[4] = fmaf(p, w1.x, acc[4]); acc[5] = fmaf(p, w1.y, acc[5]);
          acc[6] = fmaf(p, w1.z, acc[6]); acc[7] = fmaf(p, w1.w, acc[7]);
        }
      };

      constexpr int NCH = TC_SW / 32;
      for (int t = 0; t < n_tiles; ++t, ++gt) {
        const int buf = gt & 1;
        const uint32_t bph = (gt >> 1) & 1;
        const int s = gt % TC_STAGES;
        mbar_wait(&S.tfull[buf], bph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t tile_j0 = j_begin + (int64_t)t * TC_BN + slice * TC_SW;
        const int n_valid = (int)max((int64_t)0, min((int64_t)TC_SW, j_end - tile_j0));
        cnt += n_valid;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TC_BN + slice * TC_SW);
        const uint32_t tile = smem_u32(S.b[s] + slice * TC_SW * TC_D);
        uint32_t va[32];
        TC_LD32(va, taddr);
        TC_WAIT_LD(va);
        if (n_valid == TC_SW) {
          // (one register buffer: the math of a chunk is ~40x its TMEM load, two warps per scheduler hide it;
          //  a second buffer pushed the kernel into local-memory spills)
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            if (c > 0) { TC_LD32(va, taddr + c * 32); TC_WAIT_LD(va); }
            process(va, tile + (uint32_t)(c * 32 * TC_D * 4), 0, 32);
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            if (c > 0) { TC_LD32(va, taddr + c * 32); TC_WAIT_LD(va); }
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= n_valid) va[i] = 0xff800000u;   // -inf: e^{-inf} = 0, stale rows contribute nothing
            if (c * 32 < n_valid) process(va, tile + (uint32_t)(c * 32 * TC_D * 4), 0, min(32, n_valid - c * 32));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&S.tempty[buf]);
          mbar_arrive(&S.empty[s]);   // this warp no longer reads the table stage
        }
      }
      if (row < M) {
        const int64_t stream = (int64_t)item_split(w) * TC_SLICES + slice;
        float *rec = part + (stream * M + row) * CE_TC_REC;
        rec[0] = (cnt > 0) ? m : -INFINITY;
        rec[1] = l;
        rec[2] = __int_as_float(cnt);
#pragma unroll
        for (int k = 0; k < TC_D; ++k) rec[3 + k] = acc[k];
      }
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
struct CeTcPlan {
  int n_split;
  int64_t items_per_split;
  size_t ws_bytes;
};

static void ce_tc_plan(const Table *t, int64_t M, CeTcPlan *p) {
  const int row_tiles = (int)((M + TC_BM - 1) / TC_BM);
  const int64_t tiles = (t->n_rows + TC_BN - 1) / TC_BN;
  int64_t max_split = tiles / 8;
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
  p->items_per_split = tps * TC_BN;
  p->ws_bytes = (size_t)p->n_split * TC_SLICES * (size_t)M * CE_TC_REC * sizeof(float);
}

bool ce_tc_supported(const Table *t) { return t->dim == TC_D && t->tmap_valid; }

size_t ce_tc_workspace(const Table *t, int64_t M) {
  if (!ce_tc_supported(t)) return 0;
  CeTcPlan p;
  ce_tc_plan(t, M, &p);
  return p.ws_bytes;
}

// -> number of partial records per row (streams), < 0 on error
int ce_tc_launch(const Table *t, const float *Q, int64_t M, float *part, size_t ws_bytes, cudaStream_t st) {
  CeTcPlan p;
  ce_tc_plan(t, M, &p);
  if (ws_bytes < p.ws_bytes) {
    set_error("ce(tf32): workspace too small (%zu < %zu)", ws_bytes, p.ws_bytes);
    return PCV_ERR_WORKSPACE;
  }
  const size_t smem = sizeof(CeTcSmem) + 1024;
  static bool attr_set[64] = {false};
  if (!attr_set[t->device & 63]) {
    PCV_CUDA(cudaFuncSetAttribute(ce_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[t->device & 63] = true;
  }
  const int64_t n_work = ((M + TC_BM - 1) / TC_BM) * (int64_t)p.n_split;
  const unsigned grid = (unsigned)(n_work < t->sm_count ? n_work : t->sm_count);
  ce_tc_kernel<<<grid, TC_THREADS, smem, st>>>(t->packed, t->n_rows, Q, M, p.items_per_split, p.n_split, (int)n_work,
                                               part);
  PCV_LAUNCH_CHECK();
  return p.n_split * TC_SLICES;
}

}  // namespace pcv

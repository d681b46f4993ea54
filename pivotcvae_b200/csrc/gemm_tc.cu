// gemm_tc.cu — small dense GEMMs of the MLP blocks on the 5th-gen tensor cores (sm_100a):
//
//     C[M, N] (+ epilogue) = A[M, K] . B[N, K]^T          A, B row-major fp32, K contiguous ("TN")
//
// It serves the BACKWARD of the fused MLP blocks (train_generative.py:133 `loss.backward()` over
// pivotcvae.py:159-174, 204-240; the reference runs it as cuBLAS sgemm + elementwise autograd kernels):
//   dX   = G  . (W^T)^T        A = G [B, n_out],        B = W^T [n_in, n_out]     epilogue: * act'(saved input), + transposed copy
//   dW   = G^T . (X^T)^T       A = G^T [n_out, B],      B = X^T [n_in, B]         split-K partials, reduced by wgrad_reduce_kernel
// and the forward Linear+bias+activation (A = X, B = W: nn.Linear's own layout).
//
// Structure (one output tile of 128 x 128 per CTA, K walked in 32-float blocks through a 4-stage ring):
//   warp 0   : TMA producer — cp.async.bulk.tensor.2d through driver tensor maps (SWIZZLE_128B, box 32 floats x 128
//              rows = the canonical K-major UMMA atom; out-of-bounds rows / K tail are zero-filled by the TMA unit,
//              so ragged M, N, K need no special code);
//   warp 1   : one thread issues tcgen05.mma.kind::tf32 (M=128, N=128, K=8), 4 k-steps per stage, accumulating in
//              TMEM (128 columns); tcgen05.commit releases the stage / publishes the accumulator;
//   warps 2-5: epilogue — tcgen05.ld (thread = output row), bias / activation / activation-gradient, row-major store,
//              optional transposed store (coalesced: a warp writes 32 consecutive rows of one column).
// Precision: kind::tf32 truncates the operands to 10 mantissa bits.  With the optional residual operands
// (A_lo = A - tf32(A), B_lo likewise) every k-step issues three MMAs, hi*hi + lo*hi + hi*lo ("3xTF32"): the dropped
// terms are ~2^-22 relative, i.e. fp32-grade products with fp32 accumulation.
#include <cuda.h>

#include "tc_common.cuh"

namespace pcv {

constexpr int GM_BM = 128, GM_BN = 128, GM_BK = 32;   // 32 fp32 = one 128-byte swizzle row
static_assert(GM_BM == GM_BN, "one tensor-map box shape (32 floats x 128 rows) serves both operands");
constexpr int GM_MAX_STAGES = 4;
constexpr int GM_THREADS = 192;
constexpr uint32_t GM_TILE_BYTES = GM_BM * GM_BK * 4;   // 16 KB per operand tile
constexpr int GM_TILE_FLOATS = GM_BM * GM_BK;

// dynamic shared memory: [A | B] tiles per stage, then (3xTF32 only) [A_lo | B_lo] per stage; 4 stages x 32 KB, or
// 3 stages x 64 KB with the residual operands
template <bool SPLIT3>
struct GemmCfg {
  static constexpr int STAGES = SPLIT3 ? 3 : 4;
  // + epilogue staging (4 warps x 32 x 33 floats) and, in the one-pass mode, room to prefetch the 128 x 128 tile of
  // saved activations (act' epilogue) while the main loop runs
  static constexpr size_t SMEM = (size_t)STAGES * (SPLIT3 ? 4 : 2) * GM_TILE_BYTES + 4 * 32 * 33 * 4 +
                                 (SPLIT3 ? 0 : 4 * 128 * 33 * 4) + 1024;
};
struct GemmBars {
  unsigned long long full[GM_MAX_STAGES], empty[GM_MAX_STAGES], tfull;
  uint32_t tmem_base;
};

struct GemmParams {
  int64_t M, N, K;
  int k_blocks_per_split;        // 32-float blocks per z-slice
  float *C; int64_t ldc, c_split_stride;
  float *C_lo;
  float *Ct, *Ct_lo; int64_t ldct;
  const float *bias;
  int act;
  const float *dact_src; int64_t ld_dact; int dact;
};

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // LayoutType::SWIZZLE_128B
  return d;
}

constexpr uint32_t GM_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GM_BN >> 3) << 17) |
                              ((uint32_t)(GM_BM >> 4) << 24);

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1, void *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float gm_act(float v, int act) {
  if (act == PCV_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == PCV_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}
__device__ __forceinline__ float gm_dact(float saved, int act) {   // derivative from the saved post-activation output
  if (act == PCV_ACT_LEAKY) return saved > 0.f ? 1.f : 0.01f;
  if (act == PCV_ACT_RELU) return saved > 0.f ? 1.f : 0.f;
  return 1.f;
}
__device__ __forceinline__ float tf32_lo(float v) {   // v - (v truncated to 10 mantissa bits): exact in fp32
  return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
}

template <bool SPLIT3>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
                  const GemmParams P) {
  constexpr int GM_STAGES = GemmCfg<SPLIT3>::STAGES;
  extern __shared__ unsigned char smem_raw[];
  float *tiles = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto tile_a = [&](int s) { return tiles + (size_t)s * GM_TILE_FLOATS; };
  auto tile_b = [&](int s) { return tiles + (size_t)(GM_STAGES + s) * GM_TILE_FLOATS; };
  auto tile_alo = [&](int s) { return tiles + (size_t)(2 * GM_STAGES + s) * GM_TILE_FLOATS; };
  auto tile_blo = [&](int s) { return tiles + (size_t)(3 * GM_STAGES + s) * GM_TILE_FLOATS; };
  __shared__ GemmBars Bq;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * GM_BN;
  const int kb_total = (int)((P.K + GM_BK - 1) / GM_BK);
  const int kb0 = blockIdx.z * P.k_blocks_per_split;
  const int nkb = max(0, min(P.k_blocks_per_split, kb_total - kb0));

  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_STAGES; ++s) { mbar_init(&Bq.full[s], 1); mbar_init(&Bq.empty[s], 1); }
    mbar_init(&Bq.tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&Bq.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = Bq.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % GM_STAGES;
        mbar_wait(&Bq.empty[s], ((i / GM_STAGES) & 1) ^ 1);
        mbar_expect_tx(&Bq.full[s], (SPLIT3 ? 4u : 2u) * GM_TILE_BYTES);   // OOB parts are zero-filled and still counted
        const int k = (kb0 + i) * GM_BK;
        tma_load_2d(tile_a(s), &tmA, k, m0, &Bq.full[s]);
        tma_load_2d(tile_b(s), &tmB, k, n0, &Bq.full[s]);
        if (SPLIT3) {
          tma_load_2d(tile_alo(s), &tmAlo, k, m0, &Bq.full[s]);
          tma_load_2d(tile_blo(s), &tmBlo, k, n0, &Bq.full[s]);
        }
      }
    }
  } else if (warp == 1) {
    {   // the whole warp, converged; one elected lane issues (tc_common.cuh: umma_tf32_elect)
      for (int i = 0; i < nkb; ++i) {
        const int s = i % GM_STAGES;
        mbar_wait(&Bq.full[s], (i / GM_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(tile_a(s)), sb = smem_u32(tile_b(s));
#pragma unroll
        for (int j = 0; j < GM_BK / 8; ++j) {   // one k-step = 8 fp32 = 32 B inside the 128-byte swizzle row
          const uint64_t da = umma_desc_sw128(sa + j * 32), db = umma_desc_sw128(sb + j * 32);
          umma_tf32_elect(tmem, da, db, GM_IDESC, (i | j) != 0);
          if (SPLIT3) {
            umma_tf32_elect(tmem, umma_desc_sw128(smem_u32(tile_alo(s)) + j * 32), db, GM_IDESC, 1);
            umma_tf32_elect(tmem, da, umma_desc_sw128(smem_u32(tile_blo(s)) + j * 32), GM_IDESC, 1);
          }
        }
        umma_commit_elect(&Bq.empty[s]);
      }
      umma_commit_elect(&Bq.tfull);
    }
  } else {
    // ---------------- epilogue: TMEM lane = output row; global traffic is made coalesced through a per-warp
    // 32 x 33 staging tile: row-major data (C, the saved activation) moves with lane = column, the transposed copy
    // (Ct) with lane = row
    const int quarter = warp & 3;
    const int ew = warp - 2;
    float (*stage)[33] = reinterpret_cast<float (*)[33]>(tiles + (size_t)GM_STAGES * (SPLIT3 ? 4 : 2) * GM_TILE_FLOATS) + ew * 32;
    const int64_t mb = (int64_t)m0 + quarter * 32;      // first row of this warp
    const int64_t m = mb + lane;
    const bool row_ok = m < P.M;
    float (*dpre)[33] = stage + (4 - ew) * 32 + ew * 128;   // per-warp [4 chunks x 32 rows][33], behind the staging tiles
    const bool prefetched = !SPLIT3 && P.dact_src != nullptr;
    if (prefetched) {
      for (int c = 0; c < GM_BN / 32; ++c) {
        const int nb = n0 + c * 32;
        if (nb >= P.N || mb >= P.M) break;
        const bool col_ok = nb + lane < P.N;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr)
          dpre[c * 32 + rr][lane] = (mb + rr < P.M && col_ok) ? __ldg(P.dact_src + (mb + rr) * P.ld_dact + nb + lane) : 0.f;
      }
      __syncwarp();
    }
    mbar_wait(&Bq.tfull, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float *Cz = P.C ? P.C + (int64_t)blockIdx.z * P.c_split_stride : nullptr;
#pragma unroll 1
    for (int c = 0; c < GM_BN / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= P.N || mb >= P.M) break;          // warp-uniform
      const bool col_ok = nb + lane < P.N;
      uint32_t v[32];
      if (nkb > 0) {
        TC_LD32(v, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32));
        TC_WAIT_LD(v);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;   // empty K slice: the accumulator was never written
      }
      float o[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = __uint_as_float(v[i]);
        if (P.bias && nb + i < P.N) x += __ldg(P.bias + nb + i);
        o[i] = gm_act(x, P.act);
      }
      if (prefetched) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= gm_dact(dpre[c * 32 + lane][i], P.dact);
      } else if (P.dact_src) {
        // saved[m, nb + lane]: one coalesced 128-byte row per instruction, turned to lane = row through the stage
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr)
          stage[rr][lane] = (mb + rr < P.M && col_ok) ? __ldg(P.dact_src + (mb + rr) * P.ld_dact + nb + lane) : 0.f;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= gm_dact(stage[lane][i], P.dact);
        __syncwarp();
      }
      const bool direct = Cz && !P.Ct && !P.dact_src && !P.C_lo && (P.ldc % 4 == 0) && nb + 32 <= P.N &&
                          ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0);
      // full, aligned 32 x 32 sub-tile: both layouts leave through the staging tile as 128-bit stores
      const bool vec_tile = !direct && nb + 32 <= P.N && mb + 32 <= P.M &&
                            (!Cz || ((P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cz) & 15) == 0) &&
                                     (!P.C_lo || (reinterpret_cast<uintptr_t>(P.C_lo) & 15) == 0))) &&
                            (!P.Ct || ((P.ldct % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.Ct) & 15) == 0) &&
                                       (!P.Ct_lo || (reinterpret_cast<uintptr_t>(P.Ct_lo) & 15) == 0)));
      if (direct) {
        // plain tile (split-K partials): 128 contiguous bytes per thread, vector stores straight from the registers
        if (row_ok) {
          float *dst = Cz + m * P.ldc + nb;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(dst + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      } else if (vec_tile) {
#pragma unroll
        for (int i = 0; i < 32; ++i) stage[lane][i] = o[i];      // stage[row][col]
        __syncwarp();
        const int sub = lane >> 3, q4 = (lane & 7) * 4;           // 4 rows (or columns) per instruction, 4 elements per lane
        if (Cz) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub;
            const float4 x = make_float4(stage[rr][q4], stage[rr][q4 + 1], stage[rr][q4 + 2], stage[rr][q4 + 3]);
            *reinterpret_cast<float4 *>(Cz + (mb + rr) * P.ldc + nb + q4) = x;
            if (P.C_lo)
              *reinterpret_cast<float4 *>(P.C_lo + (mb + rr) * P.ldc + nb + q4) =
                  make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
          }
        }
        if (P.Ct) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int cc = it * 4 + sub;                            // column of the tile = row of the transposed copy
            const float4 x = make_float4(stage[q4][cc], stage[q4 + 1][cc], stage[q4 + 2][cc], stage[q4 + 3][cc]);
            *reinterpret_cast<float4 *>(P.Ct + (int64_t)(nb + cc) * P.ldct + mb + q4) = x;
            if (P.Ct_lo)
              *reinterpret_cast<float4 *>(P.Ct_lo + (int64_t)(nb + cc) * P.ldct + mb + q4) =
                  make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
          }
        }
        __syncwarp();
      } else {
        if (P.Ct && row_ok) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i < P.N) {
              P.Ct[(int64_t)(nb + i) * P.ldct + m] = o[i];
              if (P.Ct_lo) P.Ct_lo[(int64_t)(nb + i) * P.ldct + m] = tf32_lo(o[i]);
            }
        }
        if (Cz) {
#pragma unroll
          for (int i = 0; i < 32; ++i) stage[lane][i] = o[i];
          __syncwarp();
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            if (mb + rr < P.M && col_ok) {
              const float x = stage[rr][lane];
              Cz[(mb + rr) * P.ldc + nb + lane] = x;
              if (P.C_lo) P.C_lo[(mb + rr) * P.ldc + nb + lane] = tf32_lo(x);
            }
          }
          __syncwarp();
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
  }
}

// ---- batched 2-D transpose (+ tf32 residuals): dst[c, r] = src[r, c], optionally dst_lo = residual of the transposed
// copy and src_lo = residual of the source itself.  One launch serves every saved activation / weight of an MLP block.
struct TrJob {
  const float *src; int64_t ld_src; int rows, cols;
  float *dst; int64_t ld_dst;
  float *dst_lo;
  float *src_lo;      // [rows, ld_src] residual of the untransposed matrix
  int tile0;          // first tile id of this job in the launch
};
constexpr int TR_MAX_JOBS = 12;
struct TrBatch {
  TrJob job[TR_MAX_JOBS];
  int n_jobs, n_tiles;
};

__global__ void __launch_bounds__(256)
transpose_batch_kernel(const TrBatch Bt) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < Bt.n_jobs && (int)blockIdx.x >= Bt.job[j + 1].tile0) ++j;
  const TrJob &J = Bt.job[j];
  const int t = blockIdx.x - J.tile0;
  const int tiles_c = (J.cols + 31) / 32;
  const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    float v = 0.f;
    if (r < J.rows && c < J.cols) {
      v = J.src[(int64_t)r * J.ld_src + c];
      if (J.src_lo) J.src_lo[(int64_t)r * J.ld_src + c] = tf32_lo(v);
    }
    tile[ty + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < J.cols && r < J.rows) {
      const float v = tile[tx][ty + i];
      J.dst[(int64_t)c * J.ld_dst + r] = v;
      if (J.dst_lo) J.dst_lo[(int64_t)c * J.ld_dst + r] = tf32_lo(v);
    }
  }
}

// ---- weight-gradient reduce: dW[o, i] = sum_z part[z][o][i] (fixed order: deterministic), db[o] = sum_b Gt[o, b]
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float *__restrict__ part, int n_split, int64_t split_stride, int64_t ld_part, int n_out, int n_in,
                    float *__restrict__ dW, int64_t ld_dw, const float *__restrict__ Gt, int64_t ld_gt, int64_t Bsz,
                    float *__restrict__ db) {
  const int o = blockIdx.x;
  if (o >= n_out) return;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    float a = 0.f;
    for (int z = 0; z < n_split; ++z) a += part[(int64_t)z * split_stride + (int64_t)o * ld_part + i];
    dW[(int64_t)o * ld_dw + i] = a;
  }
  if (db) {
    float a = 0.f;
    for (int64_t b = threadIdx.x; b < Bsz; b += blockDim.x) a += Gt[(int64_t)o * ld_gt + b];
    __shared__ float red[8];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
      db[o] = t;
    }
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, K] fp32 row-major, leading dimension ld floats -> 2-D map, box = 32 floats x 128 rows, SWIZZLE_128B, zero OOB fill
static int make_operand_map(CUtensorMap *tm, const float *ptr, int64_t rows, int64_t K, int64_t ld, const char *what) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("gemm: cuTensorMapEncodeTiled is not available from the driver");
    return PCV_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld % 4) || ld < K) {
    set_error("gemm: operand %s must be 16-byte aligned with a leading dimension that is a multiple of 4 floats (ld %lld, K %lld)",
              what, (long long)ld, (long long)K);
    return PCV_ERR_ARG;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)GM_BK, (cuuint32_t)GM_BM};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
    return PCV_ERR_CUDA;
  }
  return PCV_OK;
}

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_gemm_tn(const pcv_gemm_desc *d, pcv_stream_t stream) {
  PCV_CHECK_ARG(d && d->A && d->B, "NULL operand");
  PCV_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "bad shape");
  PCV_CHECK_ARG(d->C || d->Ct, "no output");
  PCV_CHECK_ARG((d->A_lo == nullptr) == (d->B_lo == nullptr), "3xTF32 needs both residual operands (A_lo and B_lo)");
  PCV_CHECK_ARG(d->split_k >= 1, "split_k must be >= 1");
  PCV_CHECK_ARG(d->split_k == 1 || (!d->bias && d->act == PCV_ACT_NONE && !d->dact_src && !d->Ct && !d->C_lo),
                "split-K slices are raw partial sums: no epilogue");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  const bool split3 = d->A_lo != nullptr;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  if ((rc = make_operand_map(&tmA, d->A, d->M, d->K, d->lda, "A")) != PCV_OK) return rc;
  if ((rc = make_operand_map(&tmB, d->B, d->N, d->K, d->ldb, "B")) != PCV_OK) return rc;
  tmAlo = tmA;
  tmBlo = tmB;
  if (split3) {
    if ((rc = make_operand_map(&tmAlo, d->A_lo, d->M, d->K, d->lda, "A_lo")) != PCV_OK) return rc;
    if ((rc = make_operand_map(&tmBlo, d->B_lo, d->N, d->K, d->ldb, "B_lo")) != PCV_OK) return rc;
  }
  GemmParams P;
  P.M = d->M; P.N = d->N; P.K = d->K;
  const int kb_total = (int)((d->K + GM_BK - 1) / GM_BK);
  P.k_blocks_per_split = (kb_total + d->split_k - 1) / d->split_k;
  P.C = d->C; P.ldc = d->ldc; P.c_split_stride = d->c_split_stride;
  P.C_lo = d->C_lo;
  P.Ct = d->Ct; P.Ct_lo = d->Ct_lo; P.ldct = d->ldct;
  P.bias = d->bias; P.act = d->act;
  P.dact_src = d->dact_src; P.ld_dact = d->ld_dact; P.dact = d->dact;
  const size_t smem = split3 ? GemmCfg<true>::SMEM : GemmCfg<false>::SMEM;
  int dev = 0;
  cudaGetDevice(&dev);
  static bool attr_set[2][64] = {{false}};
  if (!attr_set[split3][dev & 63]) {
    if (split3) PCV_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else PCV_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[split3][dev & 63] = true;
  }
  dim3 grid((unsigned)((d->M + GM_BM - 1) / GM_BM), (unsigned)((d->N + GM_BN - 1) / GM_BN), (unsigned)d->split_k);
  cudaStream_t st = (cudaStream_t)stream;
  if (split3) gemm_tn_tc_kernel<true><<<grid, GM_THREADS, smem, st>>>(tmA, tmB, tmAlo, tmBlo, P);
  else gemm_tn_tc_kernel<false><<<grid, GM_THREADS, smem, st>>>(tmA, tmB, tmAlo, tmBlo, P);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_transpose_batch(const pcv_transpose_job *jobs, int n_jobs, pcv_stream_t stream) {
  PCV_CHECK_ARG(jobs && n_jobs >= 1 && n_jobs <= TR_MAX_JOBS, "1..12 jobs per launch");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  TrBatch Bt;
  int tiles = 0;
  for (int j = 0; j < n_jobs; ++j) {
    PCV_CHECK_ARG(jobs[j].src && jobs[j].dst && jobs[j].rows > 0 && jobs[j].cols > 0, "bad transpose job");
    Bt.job[j].src = jobs[j].src; Bt.job[j].ld_src = jobs[j].ld_src; Bt.job[j].rows = jobs[j].rows; Bt.job[j].cols = jobs[j].cols;
    Bt.job[j].dst = jobs[j].dst; Bt.job[j].ld_dst = jobs[j].ld_dst; Bt.job[j].dst_lo = jobs[j].dst_lo;
    Bt.job[j].src_lo = jobs[j].src_lo;
    Bt.job[j].tile0 = tiles;
    tiles += ((jobs[j].rows + 31) / 32) * ((jobs[j].cols + 31) / 32);
  }
  Bt.n_jobs = n_jobs;
  Bt.n_tiles = tiles;
  transpose_batch_kernel<<<(unsigned)tiles, 256, 0, (cudaStream_t)stream>>>(Bt);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_wgrad_reduce(const float *part, int n_split, int64_t split_stride, int64_t ld_part, int n_out, int n_in, float *dW,
                     int64_t ld_dw, const float *Gt, int64_t ld_gt, int64_t B, float *db, pcv_stream_t stream) {
  PCV_CHECK_ARG(part && dW && n_split >= 1 && n_out > 0 && n_in > 0, "bad arguments");
  PCV_CHECK_ARG(db == nullptr || (Gt != nullptr && B > 0), "db needs the transposed output gradient");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  wgrad_reduce_kernel<<<(unsigned)n_out, 256, 0, (cudaStream_t)stream>>>(part, n_split, split_stride, ld_part, n_out, n_in, dW,
                                                                       ld_dw, Gt, ld_gt, B, db);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

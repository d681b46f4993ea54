// mlp.cu — fused MLP blocks: gather/concat/one-hot prologue -> Linear+activation
// chain -> optional Gaussian reparameterisation epilogue, one kernel per block.
//
// Replaces (reference file:line)
//   cvae.py:85-92          get_condition (one-hot of the click count)      -> PCV_SEG_ONEHOT
//   pivotcvae.py:253,258   nn.Embedding gathers                            -> PCV_SEG_GATHER
//   pivotcvae.py:159-174   encode   [emb, c, u] -> enc_i (LeakyReLU) -> mu | logvar
//   pivotcvae.py:229-240   get_prior [c, u]     -> prior_i (LeakyReLU) -> mu | logvar
//   pivotcvae.py:204-210   PSM      [z, c, u]   -> psm_i (LeakyReLU except last)
//   pivotcvae.py:214-224   SCM      [z, c, pivot, u] -> scm_i; pivot row to slot 0 of rx
//   listcvae.py:106-119    decoder  [z, c, u]
//   cvae.py:79-83          reparametrize: z = eps * exp(0.5*logvar) + mu
//   env/response_model.py:76-87  response MLP (whole-vector L2 normalise, ReLU)
//
// Arithmetic contract: y[n] = (fma-chain over k ascending of x[k]*W[n][k], from 0) + b[n],
// i.e. the K loop is never split, so results are bit-identical to the CPU oracle.
//
// Layout: a CTA owns BM = 8/16/32 batch rows; activations ping-pong between two
// transposed shared-memory buffers actT[k][row]; each layer's weights stream through
// a double-buffered [32][256] shared-memory stage (transposed on the fly from
// nn.Linear's [n_out][n_in], prefetched one chunk ahead in registers); a thread
// accumulates an 8-row x CT-column register tile.
#include "pcv_common.cuh"

namespace pcv {

constexpr int MLP_THREADS = 256;
constexpr int MLP_KC = 32;    // k-chunk staged per step
constexpr int MLP_NB = 256;   // output columns per pass
constexpr int MLP_WLD_MAX = MLP_NB + 4;   // stage row stride: 257 (CT=1, conflict-free transposing stores) or 260

struct MlpParams {
  pcv_mlp_desc d;
  int n_in0;  // assembled input width
  int ld;     // activation row stride in shared memory (floats)
  int seg_off[PCV_MAX_SEGMENTS + 1];
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PCV_ACT_LEAKY) return v > 0.f ? v : 0.01f * v;
  if (act == PCV_ACT_RELU) return v > 0.f ? v : 0.f;
  return v;
}

// Philox normals for the reparameterisation (throughput mode): one call gives the
// four eps of latent columns 4c..4c+3 of a row (Box-Muller on two uniform pairs).
__device__ __forceinline__ void normal4(uint64_t seed, uint64_t offset, int64_t row, int c4,
                                        float n[4]) {
  uint64_t r = (uint64_t)row + offset;
  Philox4 p = philox4x32_10((uint32_t)c4, (uint32_t)r, (uint32_t)(r >> 32), PCV_STREAM_NORMAL,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  float r0 = sqrtf(-2.0f * __logf(pcv_u01(p.x)));
  float r1 = sqrtf(-2.0f * __logf(pcv_u01(p.z)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * pcv_u01(p.y), &s0, &c0);
  __sincosf(6.283185307179586f * pcv_u01(p.w), &s1, &c1);
  n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1; n[3] = r1 * s1;
}

// Thread mapping: THREADS = (256/CT column groups) x (row groups of RT rows); a thread owns
// CT adjacent output columns x RT batch rows.  Small batches use <CT=1, RT=4, 512 threads>
// (8 rows per CTA, 16 warps per SM to hide the shared-memory latency); larger ones
// <CT=2|4, RT=8, 256 threads> (16 / 32 rows per CTA).
// Activations live TRANSPOSED in shared memory (actT[k][row]) so a thread's 8 rows
// are two broadcast LDS.128; the weight stage wst[kk][n] is read conflict-free.
// Weight chunks are prefetched into registers one chunk ahead (across column blocks
// and layers) and double-buffered in shared memory: one __syncthreads per chunk.
template <int CT, int RT, int THREADS>
__device__ __forceinline__ void mlp_block(const MlpParams &P, int64_t B, float *smem) {
  constexpr int CG = MLP_NB / CT;          // column groups (threads along n)
  constexpr int BM = RT * (THREADS / CG);  // batch rows per CTA
  constexpr int ALD = BM + 4;              // actT row stride (floats)
  constexpr int MLP_WLD = (CT == 1) ? MLP_NB + 1 : MLP_NB + 4;
  constexpr int NV4 = MLP_NB * MLP_KC / 4 / THREADS;   // float4 pieces of a weight chunk staged per thread
  constexpr int NV1 = MLP_NB * MLP_KC / THREADS;       // scalar pieces (ragged chunks)
  const int ld = P.ld;              // max width (multiple of 4)
  float *actA = smem;               // [ld][ALD]
  float *actB = actA + ld * ALD;
  float *wst = actB + ld * ALD;     // [2][MLP_KC][MLP_WLD]  (actB end is 16B-aligned: ld, ALD multiples of 4)

  const int tid = threadIdx.x;
  const int cg = tid % CG, rg = tid / CG;
  const int64_t b0 = (int64_t)blockIdx.x * BM;
  const pcv_mlp_desc &d = P.d;

  // ---------------- weight-chunk stream (layer, column block, k-chunk) ----------------
  int pl = 0, pnb = 0, pkc = 0;     // position of the chunk held in wv
  // Coalesced staging: MLP_KC/4 consecutive threads read one contiguous row segment W[n][kc..kc+MLP_KC-1]
  // (a per-thread-row pattern costs 32 L1 wavefronts per warp load: it was the bottleneck).
  float wv[NV1];
  bool wvec = true;   // layout of wv: float4 pieces (full, 16B-aligned chunk) or scalars (ragged chunk)
  auto fetch = [&](int l, int nb, int kc) {
    const int K = d.layer[l].n_in, NO = d.layer[l].n_out;
    wvec = (kc + MLP_KC <= K) && ((K & 3) == 0);
    if (wvec) {
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int f = tid + THREADS * i;
        const int n = nb + f / (MLP_KC / 4), c = f % (MLP_KC / 4);
        float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < NO) t4 = __ldg(reinterpret_cast<const float4 *>(d.layer[l].W + (int64_t)n * K + kc) + c);
        wv[4 * i] = t4.x; wv[4 * i + 1] = t4.y; wv[4 * i + 2] = t4.z; wv[4 * i + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV1; ++i) {
        const int e = tid + THREADS * i;
        const int n = nb + e / MLP_KC, k = kc + e % MLP_KC;
        wv[i] = (n < NO && k < K) ? __ldg(d.layer[l].W + (int64_t)n * K + k) : 0.f;
      }
    }
  };
  auto stage = [&](float *ws) {   // registers -> transposed stage ws[kk][n]
    if (wvec) {
#pragma unroll
      for (int i = 0; i < NV4; ++i) {
        const int f = tid + THREADS * i;
        const int n = f / (MLP_KC / 4), k = (f % (MLP_KC / 4)) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) ws[(k + j) * MLP_WLD + n] = wv[4 * i + j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV1; ++i) {
        const int e = tid + THREADS * i;
        ws[(e % MLP_KC) * MLP_WLD + e / MLP_KC] = wv[i];
      }
    }
  };
  auto advance = [&]() {  // -> false when the stream is exhausted
    pkc += MLP_KC;
    if (pkc >= d.layer[pl].n_in) {
      pkc = 0;
      pnb += MLP_NB;
      if (pnb >= d.layer[pl].n_out) { pnb = 0; ++pl; }
    }
    return pl < d.n_layers;
  };
  fetch(0, 0, 0);  // in flight during the prologue
  // The whole weight set of the block (a few hundred KB) is requested into L2 up front, one
  // 128-byte line per prefetch, so the chunk-by-chunk loads below hit L2 instead of paying the
  // HBM latency once per chunk when the weights are cold (e.g. after an L2 flush).
  for (int l = 0; l < d.n_layers; ++l) {
    const char *wb = reinterpret_cast<const char *>(d.layer[l].W);
    const int64_t bytes = (int64_t)d.layer[l].n_in * d.layer[l].n_out * 4;
    for (int64_t o = (int64_t)tid * 128; o < bytes; o += (int64_t)THREADS * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(wb + o));
  }

  // ---------------- prologue: assemble x0 into actA (transposed) ----------------
  {
    constexpr int TPR = THREADS / BM;  // threads per row
    const int row = tid / TPR, sub = tid % TPR;
    const int64_t b = b0 + row;
    float *x = actA + row;                 // element e at x[e * ALD]
    if (b < B) {
      for (int s = 0; s < d.n_segments; ++s) {
        const pcv_segment &sg = d.seg[s];
        float *xs = x + P.seg_off[s] * ALD;
        if (sg.kind == PCV_SEG_DENSE) {
          const float *src = (const float *)sg.ptr + b * sg.width;
          for (int e = sub; e < sg.width; e += TPR) xs[e * ALD] = src[e];
        } else if (sg.kind == PCV_SEG_ONEHOT) {
          const float *r = (const float *)sg.ptr + b * sg.count;
          float sum = 0.f;
          for (int l = 0; l < sg.count; ++l) sum += r[l];
          const int hot = (int)sum;  // .to(torch.long) truncates (cvae.py:91)
          for (int e = sub; e <= sg.count; e += TPR) xs[e * ALD] = (e == hot) ? 1.f : 0.f;
        } else {  // GATHER
          const float *tab = (const float *)sg.ptr;
          const int n = sg.count * sg.width;
          for (int e = sub; e < n; e += TPR) {
            const int c = e / sg.width, k = e - c * sg.width;
            xs[e * ALD] = tab[sg.idx[b * sg.count + c] * (int64_t)sg.width + k];
          }
        }
      }
    } else {
      for (int e = sub; e < P.n_in0; e += TPR) x[e * ALD] = 0.f;
    }
    __syncthreads();
    // segment-wide L2 normalisation (F.normalize eps=1e-12), sequential sum order
    for (int s = 0; s < d.n_segments; ++s) {
      if (d.seg[s].norm != PCV_NORM_SEGMENT) continue;
      const int w = P.seg_off[s + 1] - P.seg_off[s];
      float *xs = x + P.seg_off[s] * ALD;
      float ss = 0.f;
      for (int e = 0; e < w; ++e) ss = fmaf(xs[e * ALD], xs[e * ALD], ss);  // every thread of the row: same value
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      __syncthreads();
      for (int e = sub; e < w; e += TPR) xs[e * ALD] = xs[e * ALD] / nrm;
      __syncthreads();
    }
    if (b < B) {
      if (d.x0) {
        float *dst = d.x0 + b * P.n_in0;
        for (int e = sub; e < P.n_in0; e += TPR) dst[e] = x[e * ALD];
      }
      if (d.copy_seg >= 0) {
        const int w = P.seg_off[d.copy_seg + 1] - P.seg_off[d.copy_seg];
        const float *xs = x + P.seg_off[d.copy_seg] * ALD;
        float *dst = d.out + b * d.out_ld;
        for (int e = sub; e < w; e += TPR) dst[e] = xs[e * ALD];
      }
    }
  }

  // ---------------- layers ----------------
  float *cur = actA, *nxt = actB;
  int buf = 0;
  bool more = true;
  for (int l = 0; l < d.n_layers; ++l) {
    const pcv_linear L = d.layer[l];
    const bool last = (l == d.n_layers - 1);
    const int K = L.n_in;
    for (int nb = 0; nb < L.n_out; nb += MLP_NB) {
      float acc[RT][CT];
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[r][c] = 0.f;

      for (int kc = 0; kc < K; kc += MLP_KC) {
        float *ws = wst + buf * (MLP_KC * MLP_WLD);
        stage(ws);
        __syncthreads();
        if (more) {
          more = advance();
          if (more) fetch(pl, pnb, pkc);   // next chunk's loads fly during this chunk's FMAs
        }
        const int kmax = (nb + cg * CT < L.n_out) ? min(MLP_KC, K - kc) : 0;   // idle column groups skip the FMAs
        const float *ap = cur + (size_t)kc * ALD + rg * RT;
        // operands of step kk: RT activations (broadcast LDS.128) + CT weights
        auto load_ops = [&](int kk, float (&a)[RT], float (&w)[CT]) {
#pragma unroll
          for (int v = 0; v < RT / 4; ++v) {
            const float4 t4 = *reinterpret_cast<const float4 *>(ap + kk * ALD + 4 * v);
            a[4 * v] = t4.x; a[4 * v + 1] = t4.y; a[4 * v + 2] = t4.z; a[4 * v + 3] = t4.w;
          }
          if (CT == 4) {
            const float4 t4 = *reinterpret_cast<const float4 *>(ws + kk * MLP_WLD + cg * 4);
            w[0] = t4.x; w[1 % CT] = t4.y; w[2 % CT] = t4.z; w[3 % CT] = t4.w;
          } else if (CT == 2) {
            const float2 t2 = *reinterpret_cast<const float2 *>(ws + kk * MLP_WLD + cg * 2);
            w[0] = t2.x; w[1 % CT] = t2.y;
          } else {
            w[0] = ws[kk * MLP_WLD + cg];
          }
        };
        auto fma_ops = [&](const float (&a)[RT], const float (&w)[CT]) {
#pragma unroll
          for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int c = 0; c < CT; ++c) acc[r][c] = fmaf(a[r], w[c], acc[r][c]);
        };
        if (kmax == MLP_KC) {
          // full chunk: explicit software pipeline, the operands of steps kk+1 and kk+2 are in flight
          // while step kk's FMAs issue (shared-memory latency ~30 cycles, only 2-4 warps per scheduler)
          float a0[RT], a1[RT], a2[RT], w0[CT], w1[CT], w2[CT];
          load_ops(0, a0, w0);
          load_ops(1, a1, w1);
#pragma unroll
          for (int kk = 0; kk < MLP_KC; kk += 3) {
            if (kk + 2 < MLP_KC) load_ops(kk + 2, a2, w2);
            fma_ops(a0, w0);
            if (kk + 3 < MLP_KC) load_ops(kk + 3, a0, w0);
            if (kk + 1 < MLP_KC) fma_ops(a1, w1);
            if (kk + 4 < MLP_KC) load_ops(kk + 4, a1, w1);
            if (kk + 2 < MLP_KC) fma_ops(a2, w2);
          }
        } else {
          for (int kk = 0; kk < kmax; ++kk) {
            float a[RT], w[CT];
            load_ops(kk, a, w);
            fma_ops(a, w);
          }
        }
        buf ^= 1;
      }
      // epilogue of this column block: bias + activation -> nxt (+ HBM)
#pragma unroll
      for (int c = 0; c < CT; ++c) {
        const int n = nb + cg * CT + c;
        if (n < L.n_out) {
          const float bias = __ldg(L.b + n);
          float v[RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) v[r] = apply_act(acc[r][c] + bias, L.act);
          float *np = nxt + (size_t)n * ALD + rg * RT;
#pragma unroll
          for (int q4 = 0; q4 < RT / 4; ++q4)
            *reinterpret_cast<float4 *>(np + 4 * q4) = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const int64_t b = b0 + rg * RT + r;
            if (b < B) {
              if (last) d.out[b * d.out_ld + d.out_col0 + n] = v[r];
              else if (d.acts[l]) d.acts[l][b * L.n_out + n] = v[r];
            }
          }
        }
      }
    }
    float *t = cur; cur = nxt; nxt = t;
  }
  __syncthreads();

  // ---------------- reparameterisation epilogue ----------------
  if (d.latent > 0) {
    const int Z = d.latent;
    const uint64_t rng_off = d.offset + (d.offset_dev ? *d.offset_dev : 0ull);
    for (int e = tid; e < BM * Z; e += THREADS) {
      const int row = e / Z, j = e - row * Z;
      const int64_t b = b0 + row;
      if (b >= B) continue;
      const float mu = cur[j * ALD + row];
      const float lv = cur[(Z + j) * ALD + row];
      float eps;
      if (d.eps) {
        eps = d.eps[b * Z + j];
      } else {
        float n4[4];
        normal4(d.seed, rng_off, b, j >> 2, n4);
        eps = n4[j & 3];
      }
      const float sd = pcv_expf(lv * 0.5f);
      d.z[b * Z + j] = eps * sd + mu;
      if (d.eps_out) d.eps_out[b * Z + j] = eps;
    }
  }
}

template <int CT, int RT, int THREADS>
__global__ void __launch_bounds__(THREADS)
mlp_fwd_kernel(const MlpParams P, int64_t B) {
  extern __shared__ __align__(16) float smem[];
  mlp_block<CT, RT, THREADS>(P, B, smem);
}

// Two chained blocks in one launch (e.g. prior -> reparameterise -> PSM, pivotcvae.py:279-291 +
// 204-210): block B's prologue reads what block A just wrote for the SAME batch rows (z), so a
// CTA-level fence + barrier is all the ordering it needs.
struct MlpParams2 {
  MlpParams a, b;
};
template <int CT, int RT, int THREADS>
__global__ void __launch_bounds__(THREADS)
mlp_fwd2_kernel(const MlpParams2 P, int64_t B) {
  extern __shared__ __align__(16) float smem[];
  mlp_block<CT, RT, THREADS>(P.a, B, smem);
  __threadfence_block();
  __syncthreads();
  mlp_block<CT, RT, THREADS>(P.b, B, smem);
}

// KL(q||p) summed (train_generative.py:61) with analytic gradients; one CTA,
// fixed reduction order (deterministic).
__global__ void __launch_bounds__(1024)
kl_kernel(const float *__restrict__ mu, const float *__restrict__ lv, const float *__restrict__ pmu,
          const float *__restrict__ plv, int64_t n, float *__restrict__ kl_out,
          float *__restrict__ dmu, float *__restrict__ dlv, float *__restrict__ dpmu,
          float *__restrict__ dplv) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float m = mu[i], v = lv[i], pm = pmu[i], pv = plv[i];
    const float ev = expf(v), ipv = expf(-pv), dm = m - pm;
    const float ratio = (ev + dm * dm) * ipv;
    acc += 1.f + v - pv - ratio;
    if (dmu) dmu[i] = dm * ipv;
    if (dlv) dlv[i] = -0.5f * (1.f - ev * ipv);
    if (dpmu) dpmu[i] = -dm * ipv;
    if (dplv) dplv[i] = -0.5f * (ratio - 1.f);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *kl_out = -0.5f * v;
  }
}

}  // namespace pcv

using namespace pcv;

extern "C" {

static int mlp_prepare(const pcv_mlp_desc *d, MlpParams *Pp, int *maxw_out, int *bias_total) {
  MlpParams &P = *Pp;
  PCV_CHECK_ARG(d != nullptr, "desc is NULL");
  PCV_CHECK_ARG(d->n_segments >= 1 && d->n_segments <= PCV_MAX_SEGMENTS, "bad n_segments");
  PCV_CHECK_ARG(d->n_layers >= 1 && d->n_layers <= PCV_MAX_LAYERS, "bad n_layers");
  PCV_CHECK_ARG(d->out != nullptr, "out is NULL");
  P.d = *d;
  int off = 0;
  for (int s = 0; s < d->n_segments; ++s) {
    const pcv_segment &sg = d->seg[s];
    PCV_CHECK_ARG(sg.ptr != nullptr, "segment pointer is NULL");
    P.seg_off[s] = off;
    if (sg.kind == PCV_SEG_DENSE) {
      PCV_CHECK_ARG(sg.width > 0, "dense segment width");
      off += sg.width;
    } else if (sg.kind == PCV_SEG_ONEHOT) {
      PCV_CHECK_ARG(sg.count > 0, "one-hot segment count");
      off += sg.count + 1;
    } else if (sg.kind == PCV_SEG_GATHER) {
      PCV_CHECK_ARG(sg.idx != nullptr && sg.width > 0 && sg.count > 0, "gather segment");
      off += sg.width * sg.count;
    } else {
      PCV_CHECK_ARG(false, "unknown segment kind");
    }
  }
  P.seg_off[d->n_segments] = off;
  P.n_in0 = off;
  int maxw = off;
  int prev = off;
  int bt = 0;
  for (int l = 0; l < d->n_layers; ++l) {
    const pcv_linear &L = d->layer[l];
    PCV_CHECK_ARG(L.W && L.b, "layer weights are NULL");
    if (L.n_in != prev) {
      set_error("pcv_mlp_fwd: layer %d expects n_in=%d but receives %d", l, L.n_in, prev);
      return PCV_ERR_ARG;
    }
    PCV_CHECK_ARG(L.n_out > 0, "layer n_out");
    PCV_CHECK_ARG(((uintptr_t)L.W & 15) == 0, "layer weight must be 16-byte aligned");
    if (L.n_out > maxw) maxw = L.n_out;
    prev = L.n_out;
    bt += L.n_out;
  }
  PCV_CHECK_ARG(maxw <= PCV_MAX_WIDTH, "layer wider than PCV_MAX_WIDTH");
  PCV_CHECK_ARG(d->copy_seg < d->n_segments, "copy_seg out of range");
  PCV_CHECK_ARG(d->out_ld >= d->out_col0 + prev, "out_ld too small");
  if (d->latent > 0) {
    PCV_CHECK_ARG(prev == 2 * d->latent, "reparam needs the last layer to emit 2*latent");
    PCV_CHECK_ARG(d->z != nullptr, "z is NULL");
  }
  *maxw_out = maxw;
  *bias_total = bt;
  return PCV_OK;
}

static int mlp_launch(const MlpParams *Pa, const MlpParams *Pb, int64_t B, cudaStream_t st) {
  int sm_count = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  // rows per CTA: 8 / 16 / 32 — small batches spread over more SMs (and use 16 warps per CTA)
  const int CT = (B <= (int64_t)sm_count * 16) ? 1 : (B <= (int64_t)sm_count * 64 ? 2 : 4);
  const int BM = 8 * CT;
  const int ld = Pb ? (Pa->ld > Pb->ld ? Pa->ld : Pb->ld) : Pa->ld;
  size_t smem = (size_t)(2 * ld * (BM + 4) + 2 * MLP_KC * MLP_WLD_MAX) * sizeof(float);
  int64_t blocks = (B + BM - 1) / BM;
#define PCV_MLP_LAUNCH(CTv, RTv, THv)                                                                              \
  if (Pb) {                                                                                                        \
    MlpParams2 P2;                                                                                                 \
    P2.a = *Pa; P2.b = *Pb;                                                                                        \
    P2.a.ld = ld; P2.b.ld = ld;                                                                                    \
    PCV_CUDA(cudaFuncSetAttribute(mlp_fwd2_kernel<CTv, RTv, THv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mlp_fwd2_kernel<CTv, RTv, THv><<<(unsigned)blocks, THv, smem, st>>>(P2, B);                                  \
  } else {                                                                                                         \
    PCV_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<CTv, RTv, THv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mlp_fwd_kernel<CTv, RTv, THv><<<(unsigned)blocks, THv, smem, st>>>(*Pa, B);                                  \
  }
  if (CT == 1) { PCV_MLP_LAUNCH(1, 4, 512) } else if (CT == 2) { PCV_MLP_LAUNCH(2, 8, 256) } else { PCV_MLP_LAUNCH(4, 8, 256) }
#undef PCV_MLP_LAUNCH
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_mlp_fwd(const pcv_mlp_desc *d, int64_t B, pcv_stream_t stream) {
  PCV_CHECK_ARG(B > 0, "B must be > 0");
  MlpParams P;
  int maxw = 0, bt = 0;
  int rc = mlp_prepare(d, &P, &maxw, &bt);
  if (rc != PCV_OK) return rc;
  rc = check_arch();
  if (rc != PCV_OK) return rc;
  P.ld = (maxw + 3) & ~3;
  return mlp_launch(&P, nullptr, B, (cudaStream_t)stream);
}

int pcv_mlp_fwd2(const pcv_mlp_desc *a, const pcv_mlp_desc *b, int64_t B, pcv_stream_t stream) {
  PCV_CHECK_ARG(B > 0, "B must be > 0");
  MlpParams Pa, Pb;
  int wa = 0, wb = 0, bt = 0;
  int rc = mlp_prepare(a, &Pa, &wa, &bt);
  if (rc != PCV_OK) return rc;
  rc = mlp_prepare(b, &Pb, &wb, &bt);
  if (rc != PCV_OK) return rc;
  rc = check_arch();
  if (rc != PCV_OK) return rc;
  Pa.ld = (wa + 3) & ~3;
  Pb.ld = (wb + 3) & ~3;
  return mlp_launch(&Pa, &Pb, B, (cudaStream_t)stream);
}

int pcv_kl_fwd_bwd(const float *mu, const float *logvar, const float *pmu, const float *plogvar,
                   int64_t n, float *kl_out, float *dmu, float *dlogvar, float *dpmu,
                   float *dplogvar, pcv_stream_t stream) {
  PCV_CHECK_ARG(mu && logvar && pmu && plogvar && kl_out, "NULL pointer");
  PCV_CHECK_ARG(n > 0, "n must be > 0");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  kl_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mu, logvar, pmu, plogvar, n, kl_out, dmu, dlogvar,
                                                  dpmu, dplogvar);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

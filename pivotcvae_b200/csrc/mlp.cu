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
// Layout: a CTA owns BM batch rows; activations ping-pong between two shared
// memory buffers [BM][ld]; each layer's weights stream through a [16][256]
// shared-memory stage (transposed on the fly from nn.Linear's [n_out][n_in]);
// a thread accumulates a (BM/8) x 8 register tile.
#include "pcv_common.cuh"

namespace pcv {

constexpr int MLP_THREADS = 256;
constexpr int MLP_KC = 16;    // k-chunk staged per step
constexpr int MLP_NB = 256;   // output columns per pass
constexpr int MLP_WLD = MLP_NB + 4;

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

template <int BM>
__global__ void __launch_bounds__(MLP_THREADS)
mlp_fwd_kernel(const MlpParams P, int64_t B) {
  constexpr int RT = BM / 8;  // rows per thread (warp w owns rows w*RT .. w*RT+RT-1)
  extern __shared__ __align__(16) float smem[];
  const int ld = P.ld;
  float *actA = smem;
  float *actB = actA + BM * ld;
  float *wst = actB + BM * ld;  // [MLP_KC][MLP_WLD]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * BM;
  const pcv_mlp_desc &d = P.d;

  // ---------------- prologue: assemble x0 into actA ----------------
  {
    constexpr int TPR = MLP_THREADS / BM;  // threads per row
    const int row = tid / TPR, sub = tid % TPR;
    const int64_t b = b0 + row;
    float *x = actA + row * ld;
    if (b < B) {
      for (int s = 0; s < d.n_segments; ++s) {
        const pcv_segment &sg = d.seg[s];
        float *xs = x + P.seg_off[s];
        if (sg.kind == PCV_SEG_DENSE) {
          const float *src = (const float *)sg.ptr + b * sg.width;
          for (int e = sub; e < sg.width; e += TPR) xs[e] = src[e];
        } else if (sg.kind == PCV_SEG_ONEHOT) {
          const float *r = (const float *)sg.ptr + b * sg.count;
          float sum = 0.f;
          for (int l = 0; l < sg.count; ++l) sum += r[l];
          const int hot = (int)sum;  // .to(torch.long) truncates (cvae.py:91)
          for (int e = sub; e <= sg.count; e += TPR) xs[e] = (e == hot) ? 1.f : 0.f;
        } else {  // GATHER
          const float *tab = (const float *)sg.ptr;
          const int n = sg.count * sg.width;
          for (int e = sub; e < n; e += TPR) {
            const int c = e / sg.width, k = e - c * sg.width;
            xs[e] = tab[sg.idx[b * sg.count + c] * (int64_t)sg.width + k];
          }
        }
      }
    } else {
      for (int e = sub; e < P.n_in0; e += TPR) x[e] = 0.f;
    }
    __syncthreads();
    // segment-wide L2 normalisation (F.normalize eps=1e-12), sequential sum order
    for (int s = 0; s < d.n_segments; ++s) {
      if (d.seg[s].norm != PCV_NORM_SEGMENT) continue;
      const int w = P.seg_off[s + 1] - P.seg_off[s];
      float *xs = x + P.seg_off[s];
      float ss = 0.f;
      for (int e = 0; e < w; ++e) ss = fmaf(xs[e], xs[e], ss);  // every thread of the row: same value
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      __syncthreads();
      for (int e = sub; e < w; e += TPR) xs[e] = xs[e] / nrm;
      __syncthreads();
    }
    if (b < B) {
      if (d.x0) {
        float *dst = d.x0 + b * P.n_in0;
        for (int e = sub; e < P.n_in0; e += TPR) dst[e] = x[e];
      }
      if (d.copy_seg >= 0) {
        const int w = P.seg_off[d.copy_seg + 1] - P.seg_off[d.copy_seg];
        const float *xs = x + P.seg_off[d.copy_seg];
        float *dst = d.out + b * d.out_ld;
        for (int e = sub; e < w; e += TPR) dst[e] = xs[e];
      }
    }
  }

  // ---------------- layers ----------------
  float *cur = actA, *nxt = actB;
  for (int l = 0; l < d.n_layers; ++l) {
    const pcv_linear L = d.layer[l];
    const bool last = (l == d.n_layers - 1);
    const int K = L.n_in;
    for (int nb = 0; nb < L.n_out; nb += MLP_NB) {
      float acc[RT][8];
#pragma unroll
      for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

      for (int kc = 0; kc < K; kc += MLP_KC) {
        __syncthreads();  // previous chunk consumed (and cur fully written)
        {
          // stage W[nb + n][kc .. kc+15] -> wst[kk][n]; thread n = tid
          const int n = nb + tid;
          float wv[MLP_KC];
          if (n < L.n_out) {
            const float *wrow = L.W + (int64_t)n * K + kc;
            if (kc + MLP_KC <= K && ((K & 3) == 0)) {
#pragma unroll
              for (int v = 0; v < MLP_KC / 4; ++v) {
                float4 t4 = __ldg(reinterpret_cast<const float4 *>(wrow) + v);
                wv[4 * v] = t4.x; wv[4 * v + 1] = t4.y; wv[4 * v + 2] = t4.z; wv[4 * v + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < MLP_KC; ++kk) wv[kk] = (kc + kk < K) ? __ldg(wrow + kk) : 0.f;
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < MLP_KC; ++kk) wv[kk] = 0.f;
          }
#pragma unroll
          for (int kk = 0; kk < MLP_KC; ++kk) wst[kk * MLP_WLD + tid] = wv[kk];
        }
        __syncthreads();
        const int kmax = min(MLP_KC, K - kc);
        for (int kk = 0; kk < kmax; ++kk) {
          const float4 w0 = *reinterpret_cast<const float4 *>(wst + kk * MLP_WLD + lane * 4);
          const float4 w1 = *reinterpret_cast<const float4 *>(wst + kk * MLP_WLD + 128 + lane * 4);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const float a = cur[(warp * RT + r) * ld + kc + kk];
            acc[r][0] = fmaf(a, w0.x, acc[r][0]);
            acc[r][1] = fmaf(a, w0.y, acc[r][1]);
            acc[r][2] = fmaf(a, w0.z, acc[r][2]);
            acc[r][3] = fmaf(a, w0.w, acc[r][3]);
            acc[r][4] = fmaf(a, w1.x, acc[r][4]);
            acc[r][5] = fmaf(a, w1.y, acc[r][5]);
            acc[r][6] = fmaf(a, w1.z, acc[r][6]);
            acc[r][7] = fmaf(a, w1.w, acc[r][7]);
          }
        }
      }
      // epilogue of this column block: bias + activation -> nxt (+ HBM)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int n = nb + (c < 4 ? lane * 4 + c : 128 + lane * 4 + (c - 4));
        if (n < L.n_out) {
          const float bias = __ldg(L.b + n);
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const int row = warp * RT + r;
            const float v = apply_act(acc[r][c] + bias, L.act);
            nxt[row * ld + n] = v;
            const int64_t b = b0 + row;
            if (b < B) {
              if (last) d.out[b * d.out_ld + d.out_col0 + n] = v;
              else if (d.acts[l]) d.acts[l][b * L.n_out + n] = v;
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
    for (int e = tid; e < BM * Z; e += MLP_THREADS) {
      const int row = e / Z, j = e - row * Z;
      const int64_t b = b0 + row;
      if (b >= B) continue;
      const float mu = cur[row * ld + j];
      const float lv = cur[row * ld + Z + j];
      float eps;
      if (d.eps) {
        eps = d.eps[b * Z + j];
      } else {
        float n4[4];
        normal4(d.seed, d.offset, b, j >> 2, n4);
        eps = n4[j & 3];
      }
      const float sd = pcv_expf(lv * 0.5f);
      d.z[b * Z + j] = eps * sd + mu;
      if (d.eps_out) d.eps_out[b * Z + j] = eps;
    }
  }
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

int pcv_mlp_fwd(const pcv_mlp_desc *d, int64_t B, pcv_stream_t stream) {
  PCV_CHECK_ARG(d != nullptr, "desc is NULL");
  PCV_CHECK_ARG(B > 0, "B must be > 0");
  PCV_CHECK_ARG(d->n_segments >= 1 && d->n_segments <= PCV_MAX_SEGMENTS, "bad n_segments");
  PCV_CHECK_ARG(d->n_layers >= 1 && d->n_layers <= PCV_MAX_LAYERS, "bad n_layers");
  PCV_CHECK_ARG(d->out != nullptr, "out is NULL");
  MlpParams P;
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
  }
  PCV_CHECK_ARG(maxw <= PCV_MAX_WIDTH, "layer wider than PCV_MAX_WIDTH");
  PCV_CHECK_ARG(d->copy_seg < d->n_segments, "copy_seg out of range");
  PCV_CHECK_ARG(d->out_ld >= d->out_col0 + prev, "out_ld too small");
  if (d->latent > 0) {
    PCV_CHECK_ARG(prev == 2 * d->latent, "reparam needs the last layer to emit 2*latent");
    PCV_CHECK_ARG(d->z != nullptr, "z is NULL");
  }
  int rc = check_arch();
  if (rc != PCV_OK) return rc;

  P.ld = ((maxw + 3) & ~3) + 4;
  int sm_count = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  const bool small = B <= (int64_t)sm_count * 32;
  const int BM = small ? 16 : 32;
  size_t smem = (size_t)(2 * BM * P.ld + MLP_KC * MLP_WLD) * sizeof(float);
  int64_t blocks = (B + BM - 1) / BM;
  cudaStream_t st = (cudaStream_t)stream;
  if (small) {
    PCV_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_fwd_kernel<16><<<(unsigned)blocks, MLP_THREADS, smem, st>>>(P, B);
  } else {
    PCV_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_fwd_kernel<32><<<(unsigned)blocks, MLP_THREADS, smem, st>>>(P, B);
  }
  PCV_LAUNCH_CHECK();
  return PCV_OK;
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

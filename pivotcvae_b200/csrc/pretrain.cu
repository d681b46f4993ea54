// pretrain.cu — the trainable half of the response model (pretrain_env.py:25-139).
//
// The reference trains UserResponseModel_MLP with BCELoss(sigmoid(pred), responses) and Adam; its embeddings ARE
// trainable there (env/response_model.py:29-37), so the backward needs what the inference path never does:
//   gather_norm_fwd : x0[b] = [ normalize(concat_l doc[slates[b, l]]) | normalize(usr[users[b]]) ]   (response_model.py:76-83)
//   gather_norm_bwd : d raw = (g - xhat (xhat . g)) / max(|raw|, eps), scatter-added into the two table gradients
//   bce_sigmoid     : mean BCE of sigmoid(pred) with torch's log clamp (-100), and d loss / d pred
// The MLP in between runs on the fused forward block and the tcgen05 backward GEMMs like every other block.
#include "pcv_common.cuh"

namespace pcv {

constexpr float GN_EPS = 1e-12f;   // F.normalize eps

// one warp per sample: lanes stride over the segment's elements
__global__ void __launch_bounds__(256)
gather_norm_fwd_kernel(const float *__restrict__ doc, const float *__restrict__ usr, const int64_t *__restrict__ slates,
                       const int64_t *__restrict__ users, int64_t B, int L, int D, float *__restrict__ x0, int64_t ld,
                       float *__restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = L * D;
  float ss = 0.f;
  for (int e = lane; e < n; e += 32) {
    const float v = doc[slates[b * L + e / D] * D + e % D];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), GN_EPS);
  for (int e = lane; e < n; e += 32) x0[b * ld + e] = doc[slates[b * L + e / D] * D + e % D] * inv;
  float uinv = 0.f;
  if (usr) {
    float us = 0.f;
    for (int e = lane; e < D; e += 32) {
      const float v = usr[users[b] * D + e];
      us = fmaf(v, v, us);
    }
    us = warp_sum(us);
    uinv = 1.0f / fmaxf(sqrtf(us), GN_EPS);
    for (int e = lane; e < D; e += 32) x0[b * ld + n + e] = usr[users[b] * D + e] * uinv;
  }
  if (lane == 0) {
    inv_norm[b * 2] = inv;
    inv_norm[b * 2 + 1] = uinv;
  }
}

__global__ void __launch_bounds__(256)
gather_norm_bwd_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ x0, int64_t ld,
                       const float *__restrict__ inv_norm, const int64_t *__restrict__ slates,
                       const int64_t *__restrict__ users, int64_t B, int L, int D, float *__restrict__ d_doc,
                       float *__restrict__ d_usr) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = L * D;
  float dot = 0.f;
  for (int e = lane; e < n; e += 32) dot = fmaf(x0[b * ld + e], g[b * ldg + e], dot);
  dot = warp_sum(dot);
  const float inv = inv_norm[b * 2];
  for (int e = lane; e < n; e += 32) {
    const float d = (g[b * ldg + e] - x0[b * ld + e] * dot) * inv;
    atomicAdd(d_doc + slates[b * L + e / D] * D + e % D, d);
  }
  if (d_usr) {
    float ud = 0.f;
    for (int e = lane; e < D; e += 32) ud = fmaf(x0[b * ld + n + e], g[b * ldg + n + e], ud);
    ud = warp_sum(ud);
    const float uinv = inv_norm[b * 2 + 1];
    for (int e = lane; e < D; e += 32) {
      const float d = (g[b * ldg + n + e] - x0[b * ld + n + e] * ud) * uinv;
      atomicAdd(d_usr + users[b] * D + e, d);
    }
  }
}

// mean over n of -(t log s + (1 - t) log(1 - s)), s = sigmoid(p), logs clamped at -100 like nn.BCELoss;
// dp = d loss / d p = (s - t) s (1 - s) / max(s (1 - s), 1e-12) / n.  One block, fixed order: deterministic.
__global__ void __launch_bounds__(1024)
bce_sigmoid_kernel(const float *__restrict__ p, const float *__restrict__ t, int64_t n, float *__restrict__ loss,
                   float *__restrict__ dp) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv_n = 1.0f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = 1.0f / (1.0f + expf(-p[i]));
    const float ti = t[i];
    acc -= ti * fmaxf(logf(s), -100.f) + (1.f - ti) * fmaxf(logf(1.f - s), -100.f);
    if (dp) {
      const float v = s * (1.f - s);
      dp[i] = (s - ti) * v / fmaxf(v, 1e-12f) * inv_n;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    *loss = tot * inv_n;
  }
}

// Backward of the reparameterisation epilogue (cvae.py:79-83: z = eps * exp(0.5 logvar) + mu) folded into the gradient of
// the block's [mu | logvar] output: g[b, j] = d_out[b, j] + d_z[b, j]            (j < Z)
//                                 g[b, Z + j] = d_out[b, Z + j] + d_z[b, j] * eps[b, j] * 0.5 * exp(0.5 * logvar[b, j])
__global__ void reparam_bwd_kernel(const float *__restrict__ d_out, int64_t ld_dout, const float *__restrict__ d_z,
                                   const float *__restrict__ eps, const float *__restrict__ out, int64_t ld_out, int64_t B, int Z,
                                   float *__restrict__ g) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B * 2 * Z) return;
  const int64_t b = f / (2 * Z);
  const int j = (int)(f % (2 * Z));
  float v = d_out ? d_out[b * ld_dout + j] : 0.f;
  if (d_z) {
    if (j < Z) v += d_z[b * Z + j];
    else v += d_z[b * Z + j - Z] * eps[b * Z + j - Z] * (0.5f * expf(0.5f * out[b * ld_out + j]));
  }
  g[f] = v;
}

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_gather_norm_fwd(const float *doc, const float *usr, const int64_t *slates, const int64_t *users, int64_t B, int L,
                        int D, float *x0, int64_t ld, float *inv_norm, pcv_stream_t stream) {
  PCV_CHECK_ARG(doc && slates && x0 && inv_norm, "NULL pointer");
  PCV_CHECK_ARG(usr == nullptr || users != nullptr, "users is NULL");
  PCV_CHECK_ARG(B > 0 && L > 0 && D > 0 && ld >= (int64_t)(L + (usr ? 1 : 0)) * D, "bad shape");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  gather_norm_fwd_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(doc, usr, slates, users, B, L, D, x0, ld, inv_norm);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_gather_norm_bwd(const float *g, int64_t ldg, const float *x0, int64_t ld, const float *inv_norm, const int64_t *slates,
                        const int64_t *users, int64_t B, int L, int D, float *d_doc, float *d_usr, pcv_stream_t stream) {
  PCV_CHECK_ARG(g && x0 && inv_norm && slates && d_doc, "NULL pointer");
  PCV_CHECK_ARG(d_usr == nullptr || users != nullptr, "users is NULL");
  PCV_CHECK_ARG(B > 0 && L > 0 && D > 0, "bad shape");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  gather_norm_bwd_kernel<<<(unsigned)((B + 7) / 8), 256, 0, (cudaStream_t)stream>>>(g, ldg, x0, ld, inv_norm, slates, users, B, L, D,
                                                                                  d_doc, d_usr);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_reparam_bwd(const float *d_out, int64_t ld_dout, const float *d_z, const float *eps, const float *out, int64_t ld_out,
                    int64_t B, int Z, float *g, pcv_stream_t stream) {
  PCV_CHECK_ARG(g && out && B > 0 && Z > 0, "bad arguments");
  PCV_CHECK_ARG(d_z == nullptr || eps != nullptr, "d_z needs eps");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  const int64_t n = B * 2 * Z;
  reparam_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_out, ld_dout, d_z, eps, out, ld_out, B, Z, g);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_bce_sigmoid(const float *pred, const float *target, int64_t n, float *loss, float *dpred, pcv_stream_t stream) {
  PCV_CHECK_ARG(pred && target && loss && n > 0, "bad arguments");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  bce_sigmoid_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, n, loss, dpred);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

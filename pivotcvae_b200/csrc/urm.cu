// urm.cu — simulator response models as one gather-plus-dot kernel.
//
// Replaces (reference file:line)
//   env/response_model.py:129-150  URM.core_forward       sigmoid(<norm(d_l), u> + b_item + b_user)
//   env/response_model.py:286-295  URM_P.core_forward     + u @ posDependentBias.view(D, L) + posBias
//   env/response_model.py:315-323  URM_P_MR.core_forward  + mr * <d_l, sigmoid(mean_l d_l)>
// Quirks preserved: the user row is used UN-normalised (:145 overwrites :144);
// posDependentBias (L, D) is reinterpreted as a (D, L) view (:292), not transposed;
// URM_P adds its positional terms AFTER the sigmoid (:294).
//
// One thread per slate: L rows of D floats (32 B each at D=8) are fetched with
// float4 loads; everything else lives in registers.  HBM/L2-gather bound.
#include "pcv_common.cuh"

namespace pcv {

constexpr int URM_MAX_L = 16;

template <int D>
__global__ void __launch_bounds__(128)
urm_kernel(const pcv_urm_desc P, const int64_t *__restrict__ slates,
           const int64_t *__restrict__ users, int64_t B, float *__restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int L = P.L;
  const int64_t u = users[b];
  float ue[D];
#pragma unroll
  for (int c = 0; c < D / 4; ++c) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(P.user_table + u * D) + c);
    ue[4 * c] = v.x; ue[4 * c + 1] = v.y; ue[4 * c + 2] = v.z; ue[4 * c + 3] = v.w;
  }
  const float ub = __ldg(P.user_bias + u);
  float mean[D];
#pragma unroll
  for (int k = 0; k < D; ++k) mean[k] = 0.f;
  float raw[URM_MAX_L];

  for (int l = 0; l < L; ++l) {
    const int64_t it = slates[b * L + l];
    float de[D];
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 v = __ldg(reinterpret_cast<const float4 *>(P.doc_table + it * D) + c);
      de[4 * c] = v.x; de[4 * c + 1] = v.y; de[4 * c + 2] = v.z; de[4 * c + 3] = v.w;
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) ss = fmaf(de[k], de[k], ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      de[k] = de[k] / nrm;
      dot = fmaf(de[k], ue[k], dot);
      mean[k] += de[k];
    }
    float v = dot + __ldg(P.item_bias + it);
    v = v + ub;
    raw[l] = 1.0f / (1.0f + expf(-v));
  }
  if (P.variant >= PCV_URM_P) {
    for (int l = 0; l < L; ++l) {
      float pb = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) pb = fmaf(ue[k], __ldg(P.pos_dep + k * L + l), pb);
      pb = pb + __ldg(P.pos_bias + l);
      raw[l] = raw[l] + pb;
    }
  }
  if (P.variant == PCV_URM_P_MR) {
    float att[D];
#pragma unroll
    for (int k = 0; k < D; ++k) att[k] = 1.0f / (1.0f + expf(-(mean[k] / (float)L)));
    for (int l = 0; l < L; ++l) {
      const int64_t it = slates[b * L + l];
      float de[D];
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(P.doc_table + it * D) + c);
        de[4 * c] = v.x; de[4 * c + 1] = v.y; de[4 * c + 2] = v.z; de[4 * c + 3] = v.w;
      }
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) ss = fmaf(de[k], de[k], ss);
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      float rel = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) rel = fmaf(de[k] / nrm, att[k], rel);
      raw[l] = raw[l] + rel * P.mr_factor;
    }
  }
  for (int l = 0; l < L; ++l) out[b * L + l] = raw[l];
}

// ---------------------------------------------------------------------------
// Slate metrics of the variation-control evaluation:
//   analysis.py:13-30  get_ILS: mean pairwise cosine similarity inside a slate,
//                      (sum_{i,j} <e_i, e_j> - L) / (L (L-1)), e = row-normalised embeddings
//   analysis.py:5-11   get_coverage: |unique(slates)| / N  (bitmap + popcount here)
// One thread per slate; the coverage bitmap is OR-ed with atomics (N/32 words).
// ---------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
slate_metrics_kernel(const float *__restrict__ table, const int64_t *__restrict__ slates, int64_t B, int L,
                     float *__restrict__ ils, unsigned int *__restrict__ bitmap) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sum[D];
#pragma unroll
  for (int k = 0; k < D; ++k) sum[k] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int64_t it = slates[b * L + l];
    if (bitmap) atomicOr(bitmap + (it >> 5), 1u << (it & 31));
    float e[D];
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4 *>(table + it * D) + c);
      e[4 * c] = v.x; e[4 * c + 1] = v.y; e[4 * c + 2] = v.z; e[4 * c + 3] = v.w;
    }
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) ss = fmaf(e[k], e[k], ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int k = 0; k < D; ++k) sum[k] += e[k] / nrm;
  }
  // sum_{i,j} <e_i, e_j> = |sum_i e_i|^2
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) tot = fmaf(sum[k], sum[k], tot);
  if (ils) ils[b] = (tot - (float)L) / (float)(L * (L - 1));
}

__global__ void popcount_kernel(const unsigned int *__restrict__ bitmap, int64_t words, unsigned long long *out) {
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x)
    c += __popc(bitmap[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_slate_metrics(const float *table, int D, const int64_t *slates, int64_t B, int L, float *ils,
                      uint32_t *bitmap, pcv_stream_t stream) {
  PCV_CHECK_ARG(table && slates, "NULL pointer");
  PCV_CHECK_ARG(B > 0 && L >= 2, "bad shape (slates need >= 2 items)");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((B + 127) / 128);
  switch (D) {
    case 4: slate_metrics_kernel<4><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    case 8: slate_metrics_kernel<8><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    case 16: slate_metrics_kernel<16><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    case 32: slate_metrics_kernel<32><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    case 64: slate_metrics_kernel<64><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    case 128: slate_metrics_kernel<128><<<blocks, 128, 0, st>>>(table, slates, B, L, ils, bitmap); break;
    default:
      set_error("pcv_slate_metrics: dim %d unsupported (use 4, 8, 16, 32, 64 or 128)", D);
      return PCV_ERR_UNSUPPORTED;
  }
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_popcount(const uint32_t *bitmap, int64_t words, uint64_t *count, pcv_stream_t stream) {
  PCV_CHECK_ARG(bitmap && count && words > 0, "bad arguments");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PCV_CUDA(cudaMemsetAsync(count, 0, sizeof(uint64_t), st));
  int blocks = (int)((words + 255) / 256 > 1024 ? 1024 : (words + 255) / 256);
  popcount_kernel<<<blocks, 256, 0, st>>>(bitmap, words, reinterpret_cast<unsigned long long *>(count));
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_urm_fwd(const pcv_urm_desc *d, const int64_t *slates, const int64_t *users, int64_t B,
                float *out, pcv_stream_t stream) {
  PCV_CHECK_ARG(d && slates && users && out, "NULL pointer");
  PCV_CHECK_ARG(B > 0, "B must be > 0");
  PCV_CHECK_ARG(d->variant >= PCV_URM && d->variant <= PCV_URM_P_MR, "bad variant");
  PCV_CHECK_ARG(d->doc_table && d->user_table && d->item_bias && d->user_bias, "NULL table");
  PCV_CHECK_ARG(d->L >= 1 && d->L <= URM_MAX_L, "slate size must be in [1,16]");
  if (d->variant >= PCV_URM_P) PCV_CHECK_ARG(d->pos_bias && d->pos_dep, "NULL positional bias");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = (unsigned)((B + 127) / 128);
  switch (d->D) {
    case 4: urm_kernel<4><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    case 8: urm_kernel<8><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    case 16: urm_kernel<16><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    case 32: urm_kernel<32><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    case 64: urm_kernel<64><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    case 128: urm_kernel<128><<<blocks, 128, 0, st>>>(*d, slates, users, B, out); break;
    default:
      set_error("pcv_urm_fwd: dim %d unsupported (use 4, 8, 16, 32, 64 or 128)", d->D);
      return PCV_ERR_UNSUPPORTED;
  }
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

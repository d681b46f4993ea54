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
constexpr int URM_MAX_D = 32;

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

}  // namespace pcv

using namespace pcv;

extern "C" {

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
    default:
      set_error("pcv_urm_fwd: dim %d unsupported (use 4, 8, 16 or 32)", d->D);
      return PCV_ERR_UNSUPPORTED;
  }
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

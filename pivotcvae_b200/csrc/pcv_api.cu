// pcv_api.cu — handles, error reporting, table construction.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <new>

#include "pcv_common.cuh"

namespace pcv {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_arch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device is current (libpcv_b200 has no CPU path)");
    return PCV_ERR_CUDA;
  }
  return pcv_device_ok(dev);
}

// F.normalize(p=2, dim=1, eps=1e-12): x / max(||x||_2, eps)   (cvae.py:31)
// One thread per row; the sum of squares is a sequential-k FMA chain (the
// oracle restates the same order).
__global__ void normalize_rows_kernel(const float *__restrict__ W, int64_t n, int dim,
                                      float *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *w = W + i * dim;
  float ss = 0.f;
  for (int k = 0; k < dim; ++k) ss = fmaf(w[k], w[k], ss);
  float nrm = fmaxf(sqrtf(ss), 1e-12f);
  for (int k = 0; k < dim; ++k) out[i * dim + k] = w[k] / nrm;
}

// max_j |w_j|_2^2 as an ordered-uint atomicMax (norms are >= 0)
__global__ void max_row_norm2_kernel(const float *__restrict__ W, int64_t n, int dim, unsigned int *out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float ss = 0.f;
  if (i < n)
    for (int k = 0; k < dim; ++k) ss = fmaf(W[i * dim + k], W[i * dim + k], ss);
  ss = warp_max(ss);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(ss));
}

__global__ void counter_add_kernel(unsigned long long *c, unsigned long long inc) { *c += inc; }

int table_init_tc(Table *t);  // score_select_tc.cu
void table_free_tc(Table *t);
void table_free_ce(Table *t);   // ce_tc2.cu

}  // namespace pcv

using namespace pcv;

extern "C" {

int pcv_abi_version(void) { return PCV_ABI_VERSION; }
const char *pcv_last_error(void) { return g_err; }
int64_t pcv_launch_count(void) { return g_launches.load(); }

int pcv_counter_add(uint64_t *counter, uint64_t inc, pcv_stream_t stream) {
  PCV_CHECK_ARG(counter != nullptr, "counter is NULL");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long *>(counter), inc);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

int pcv_device_ok(int device) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device) != cudaSuccess) {
    cudaGetLastError();
    set_error("pcv_device_ok: cannot query device %d (no CUDA device? this library has no CPU path)", device);
    return PCV_ERR_CUDA;
  }
  if (major != 10) {
    set_error("pcv_device_ok: device %d is sm_%d%d; libpcv_b200 is built for sm_100a only", device, major, minor);
    return PCV_ERR_ARCH;
  }
  return PCV_OK;
}

int pcv_table_create(const float *W, int64_t n_rows, int dim, int64_t row_offset,
                     pcv_table **out) {
  PCV_CHECK_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  PCV_CHECK_ARG(W != nullptr, "W is NULL");
  PCV_CHECK_ARG(n_rows > 0, "n_rows must be > 0");
  PCV_CHECK_ARG(dim >= 4 && dim <= 128 && dim % 4 == 0, "dim must be a multiple of 4 in [4,128]");
  PCV_CHECK_ARG(((uintptr_t)W & 15) == 0, "W must be 16-byte aligned");
  PCV_CHECK_ARG(row_offset >= 0, "row_offset must be >= 0");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  Table *t = new (std::nothrow) Table();
  PCV_CHECK_ARG(t != nullptr, "out of host memory");
  t->W = W;
  t->n_rows = n_rows;
  t->dim = dim;
  t->row_offset = row_offset;
  t->tmap_valid = 0;
  t->packed = nullptr;
  t->packed_t = nullptr;
  t->packed_h = nullptr;
  cudaGetDevice(&t->device);
  cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, t->device);
  // one-off: max row norm (error bound of the tf32 filter) + TMA descriptor
  {
    unsigned int *d_max = nullptr;
    unsigned int h_max = 0;
    cudaError_t e = cudaDeviceSynchronize();  // the table may still be being written on another stream
    if (e == cudaSuccess) e = cudaMalloc(&d_max, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(d_max, 0, sizeof(unsigned int));
    if (e == cudaSuccess) {
      max_row_norm2_kernel<<<(unsigned)((n_rows + 255) / 256), 256>>>(W, n_rows, dim, d_max);
      count_launch();
      e = cudaMemcpy(&h_max, d_max, sizeof(unsigned int), cudaMemcpyDeviceToHost);
    }
    if (d_max) cudaFree(d_max);
    if (e != cudaSuccess) {
      set_error("pcv_table_create: %s", cudaGetErrorString(e));
      delete t;
      return PCV_ERR_CUDA;
    }
    float m2;
    memcpy(&m2, &h_max, sizeof(float));
    t->max_row_norm = sqrtf(m2) * 1.000001f;
  }
  rc = table_init_tc(t);
  if (rc != PCV_OK) {
    delete t;
    return rc;
  }
  *out = reinterpret_cast<pcv_table *>(t);
  return PCV_OK;
}

void pcv_table_destroy(pcv_table *th) {
  Table *t = reinterpret_cast<Table *>(th);
  if (!t) return;
  table_free_tc(t);
  table_free_ce(t);
  delete t;
}

int pcv_normalize_rows(const float *W, int64_t n_rows, int dim, float *out,
                       pcv_stream_t stream) {
  PCV_CHECK_ARG(W && out, "NULL pointer");
  PCV_CHECK_ARG(n_rows > 0 && dim > 0, "bad shape");
  int rc = check_arch();
  if (rc != PCV_OK) return rc;
  int threads = 256;
  int64_t blocks = (n_rows + threads - 1) / threads;
  normalize_rows_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(W, n_rows, dim, out);
  PCV_LAUNCH_CHECK();
  return PCV_OK;
}

}  // extern "C"

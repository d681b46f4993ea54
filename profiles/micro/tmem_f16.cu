// Micro-benchmark (not part of the library): how do f16 accumulators of tcgen05.mma.kind::f16 live in tensor memory, and
// does tcgen05.ld ... .pack::16b read them at twice the fp32 element rate?  (Decides whether an f16-accumulator variant
// of the select filter can beat the TMEM-read bound of the fp32 one.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_f16 tmem_f16.cu && ./tmem_f16
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw32(const void *smem) {   // K-major, SWIZZLE_32B: rows of 32 B, 8-row groups 256 B apart
  uint64_t d = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
#define LD32(v, taddr, MOD)                                                                                 \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32" MOD ".b32 "                                             \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                             \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"             \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),     \
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),            \
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),          \
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),          \
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                               \
      : "r"(taddr))
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

struct __align__(1024) Smem {
  __half a[128 * 16];     // [128 rows][16 k] fp16, SWIZZLE_32B image
  __half b[256 * 16];     // [256 rows][16 k]
  unsigned long long bar;
  uint32_t tmem_base;
};

// mode 0: f32 accumulators, mode 1: f16 accumulators
__global__ void __launch_bounds__(512) k_layout(int mode, uint32_t *out_raw, uint32_t *out_pack, long long *cyc, int reps) {
  __shared__ Smem S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // A[m][0] = 1, B[n][0] = n (exact in fp16 up to 2048), everything else 0 -> D[m][n] = n
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) S.a[i] = __float2half(0.f);
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) S.b[i] = __float2half(0.f);
  __syncthreads();
  // element (row r, k) of the SWIZZLE_32B image: 16-byte half h = k / 8 swapped when row bit 2 is set
  for (int r = threadIdx.x; r < 128; r += blockDim.x) S.a[r * 16 + ((0 ^ ((r >> 2) & 1)) * 8)] = __float2half(1.f);
  for (int r = threadIdx.x; r < 256; r += blockDim.x) S.b[r * 16 + ((0 ^ ((r >> 2) & 1)) * 8)] = __float2half((float)r);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;
  if (threadIdx.x == 0) {
    // kind::f16: a/b format F16 (0), c format F16 (0) or F32 (1); K-major; N = 256, M = 128
    const uint32_t idesc = ((mode == 0 ? 1u : 0u) << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t da = desc_sw32(S.a), db = desc_sw32(S.b);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar)) : "memory");
  }
  {
    uint32_t addr = smem_u32(&S.bar);
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(addr), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t v[32];
  if (warp < 4) {
    for (int c = 0; c < 8; ++c) {       // raw 32-bit view of columns 0..255
      LD32(v, lane_base + c * 32, "");
      WAITLD();
      if (threadIdx.x == 5) for (int i = 0; i < 32; ++i) out_raw[c * 32 + i] = v[i];
    }
    for (int c = 0; c < 4; ++c) {       // pack::16b view: 64 columns per load
      LD32(v, lane_base + c * 64, ".pack::16b");
      WAITLD();
      if (threadIdx.x == 5) for (int i = 0; i < 32; ++i) out_pack[c * 32 + i] = v[i];
    }
  }
  __syncthreads();
  // throughput: 16 warps (4 per lane quarter, like the filter's epilogue), each reads its 64-column slice `reps` times
  uint32_t sink = 0;
  const uint32_t slice = (uint32_t)(warp >> 2) * 64;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    LD32(v, lane_base + slice, "");
    WAITLD();
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= v[i];
    LD32(v, lane_base + slice + 32, "");
    WAITLD();
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= v[i];
  }
  __syncthreads();
  long long t1 = clock64();
  for (int r = 0; r < reps; ++r) {     // the same 64 columns through one packed load
    LD32(v, lane_base + slice, ".pack::16b");
    WAITLD();
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= v[i];
  }
  __syncthreads();
  long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
  if (sink == 0x12345678u) out_raw[0] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  uint32_t *raw, *pack;
  long long *cyc;
  cudaMallocManaged(&raw, 256 * 4);
  cudaMallocManaged(&pack, 128 * 4);
  cudaMallocManaged(&cyc, 16);
  const int reps = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(raw, 0, 256 * 4);
    cudaMemset(pack, 0, 128 * 4);
    k_layout<<<1, 512>>>(mode, raw, pack, cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== accumulators %s: %s\n", mode ? "f16" : "f32", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    printf("raw  (row 5, 32-bit TMEM columns 0..15):");
    for (int i = 0; i < 16; ++i) printf(" %08x", raw[i]);
    printf("\nraw  (columns 128..135):");
    for (int i = 128; i < 136; ++i) printf(" %08x", raw[i]);
    printf("\npack (row 5, first 8 registers of the .pack::16b load at column 0):");
    for (int i = 0; i < 8; ++i) printf(" %08x", pack[i]);
    printf("\n16 warps x %d reps over a 64-column slice: 2 x (x32) = %lld cycles, 1 x (x32.pack::16b) = %lld cycles\n", reps, cyc[0], cyc[1]);
    printf("   -> %.1f / %.1f cycles per 128 rows x 256 columns\n", (double)cyc[0] / reps, (double)cyc[1] / reps);
  }
  return 0;
}

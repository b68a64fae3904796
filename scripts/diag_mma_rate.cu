// Issue-rate probe for tcgen05.mma on sm_100a: cycles per MMA instruction as a function of kind (tf32 / f16=bf16),
// N (32..256) and M (128, 64), cta_group::1, operands = zeros in shared memory (SWIZZLE_128B K-major canonical tiles),
// accumulator in TMEM.  One CTA per SM, one issuing thread, `iters` back-to-back MMAs, one commit, clock64 around.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_variants/diag_mma_rate scripts/diag_mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}
__host__ __device__ inline uint32_t make_idesc(int kind, int m, int n) {
  uint32_t d = 0;
  d |= 1u << 4;                                     // D = f32
  d |= (uint32_t)(kind == 0 ? 2 : 1) << 7;          // A: tf32 (2) / bf16 (1)
  d |= (uint32_t)(kind == 0 ? 2 : 1) << 10;         // B
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(int m, int n, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += 128)
    reinterpret_cast<float4*>(smem + (base - smem_u32(smem)))[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(KIND, m, n);
    const uint64_t da = make_desc(base), db = make_desc(base + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(i) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  printf("kind  M    N   grid  issue cyc/MMA  complete cyc/MMA   MAC/clk/SM\n");
  for (int kind = 0; kind < 2; ++kind)
    for (int m : {128, 64})
      for (int n : {32, 64, 96, 128, 192, 256})
        for (int grid : {1, 148}) {
          if (kind == 0) probe<0><<<grid, 128, 100 * 1024>>>(m, n, iters, out);
          else probe<1><<<grid, 128, 100 * 1024>>>(m, n, iters, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s (kind %d m %d n %d)\n", cudaGetErrorString(e), kind, m, n); return 1; }
          long long h[2];
          cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
          const double per = (double)h[1] / iters;
          const int k = kind == 0 ? 8 : 16;
          printf("%-5s %3d  %3d  %4d   %10.1f   %12.1f   %10.0f\n", kind == 0 ? "tf32" : "bf16", m, n, grid, (double)h[0] / iters, per,
                 (double)m * n * k / per);
        }
  return 0;
}

// mma_rate.cu -- microbenchmark: cycles per tcgen05.mma (kind::tf32, K = 8; kind::f16 bf16, K = 16) as a
// function of the instruction shape, for one CTA (M = 128) and for a CTA pair (cta_group::2, M = 256).
// Operands are whatever is in shared memory; only the issue/execute rate matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t a)
{
   return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mwait(uint64_t *bar, uint32_t parity)
{
   uint32_t done, addr = s32(bar);
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
   } while (!done);
}

// kind: 0 = tf32, 1 = bf16
template <int PAIR, int ND>
__global__ void __launch_bounds__(128, 1) rate_kernel(int kind, int N, int reps, int nDist, long long *out, int ats)
{
   extern __shared__ uint8_t raw[];
   uint8_t *base = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
   __shared__ uint64_t bar;
   __shared__ uint32_t slot;
   uint32_t rank = 0;
   if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
   for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((float *)base)[i] = 0.001f * (i % 97);
   if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (threadIdx.x < 32) {
      if (PAIR) {
         asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512) : "memory");
         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
      } else {
         asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(512) : "memory");
         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
      }
   }
   asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
   else __syncthreads();
   asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
   const uint32_t tmem = slot;
   const int M = PAIR ? 256 : 128;
   const uint32_t fmt = kind == 0 ? 2u : 1u;            // a/b format: TF32 = 2, BF16 = 1 (kind::f16)
   const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
   if (threadIdx.x < 32 && rank == 0) {
      // the whole warp runs the loop (descriptor arithmetic stays on the uniform datapath); one elected
      // lane issues the tcgen05 instructions
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
      const uint32_t a0 = s32(base), b0 = s32(base + 64 * 1024);
      long long best = 1ll << 60;
      for (int trial = 0; trial < 5; trial++) {
         long long t0 = clock64();
         for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
               // 8 operand slices in rotation (kk offsets as in a real K loop), two accumulators
               const uint64_t da = sdesc(a0 + (j & 3) * 32 + (j >> 2) * 16384), db = sdesc(b0 + (j & 3) * 32 + (j >> 2) * 16384);
               const uint32_t d = tmem + (uint32_t)(((j / ND) & 1) * (ats ? 128 : 256));   // nDist = MMAs in a row into one accumulator
               const uint32_t ta = tmem + 384 + (uint32_t)(j * 8);      // A operand in TMEM: K = 16 halfs = 8 columns per MMA
               if (elected && ats) {
                  if (PAIR) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(1u) : "memory");
                  else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(1u) : "memory");
               } else
               if (elected) {
                  if (PAIR) {
                     if (kind == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                     else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                  } else {
                     if (kind == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                     else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
                  }
               }
            }
         }
         long long t1 = clock64();
         if (elected) {
         if (PAIR) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
         else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
         }
         __syncwarp();
         mwait(&bar, trial & 1);
         long long t2 = clock64();
         if (elected && t2 - t0 < best) { best = t2 - t0; out[0] = t1 - t0; out[1] = t2 - t0; }
      }
   }
   asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
   if (PAIR) { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
   else __syncthreads();
   if (threadIdx.x < 32) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
      else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
   }
}

int main()
{
   long long *d, h[2];
   cudaMalloc(&d, 16);
   const int smem = 200 * 1024, reps = 2000;
   cudaFuncSetAttribute(rate_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
   cudaFuncSetAttribute(rate_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
   cudaFuncSetAttribute(rate_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
   cudaFuncSetAttribute(rate_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
   for (int grid = 148; grid <= 148; grid += 147)
      for (int pair = 0; pair < 2; pair++)
         for (int kind = 0; kind < 2; kind++)
          for (int ats = 0; ats < 2; ats++)
           for (int nDist = 1; nDist <= 8; nDist *= 8)
            for (int N = 64; N <= 256; N *= 2) {
               if (ats && (kind == 0 || N > 128)) continue;
               if (pair) {
                  cudaLaunchConfig_t cfg = {};
                  cfg.gridDim = dim3(grid == 1 ? 2 : 148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                  cudaLaunchAttribute at[1];
                  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                  cfg.attrs = at; cfg.numAttrs = 1;
                  if (nDist == 1) cudaLaunchKernelEx(&cfg, rate_kernel<1, 1>, kind, N, reps, nDist, d, ats);
                  else cudaLaunchKernelEx(&cfg, rate_kernel<1, 8>, kind, N, reps, nDist, d, ats);
               } else if (nDist == 1) rate_kernel<0, 1><<<grid, 128, smem>>>(kind, N, reps, nDist, d, ats);
               else rate_kernel<0, 8><<<grid, 128, smem>>>(kind, N, reps, nDist, d, ats);
               cudaError_t e = cudaDeviceSynchronize();
               cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
               const int M = pair ? 256 : 128, K = kind == 0 ? 8 : 16;
               const double cyc = (double)h[1] / reps;
               printf("grid %3d  %s  %s  %s  run of %d per accumulator  M=%3d N=%3d K=%2d : issue %.1f cyc/MMA, issue+drain %.1f cyc/MMA -> %.0f MAC/clk/SM  (%s)\n",
                      grid, pair ? "cta_group::2" : "cta_group::1", kind == 0 ? "tf32" : "bf16", ats ? "A in TMEM" : "A in smem", nDist, M, N, K, (double)h[0] / reps, cyc,
                      (double)(M / (pair ? 2 : 1)) * N * K / cyc, cudaGetErrorString(e));
            }
   return 0;
}

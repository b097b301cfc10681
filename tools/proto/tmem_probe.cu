// tmem_probe.cu - is tensor memory usable as warp-private scratch?  Round trip check and the
// throughput of tcgen05.ld (32x32b.x16: four 16-byte pairs per lane) against LDS.128 under the
// trajectory kernel's occupancy (2 CTAs x 4 warps per SM).
// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define NCOLS 128

__device__ __forceinline__ void tm_st4(unsigned taddr, const unsigned (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                  "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_ld4(unsigned taddr, unsigned (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 2) probe(int iters, int* errors, long long* cyc_tm, long long* cyc_lds,
                                                unsigned* sink) {
  __shared__ unsigned tm_base_s;
  extern __shared__ uint4 smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((unsigned)__cvta_generic_to_shared(&tm_base_s)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tbase = tm_base_s + ((unsigned)(32 * warp) << 16);
  // ---- round trip: every 16-column group gets a distinct pattern ----
  int bad = 0;
  for (int g = 0; g < NCOLS / 16; ++g) {
    unsigned r[16];
    for (int k = 0; k < 16; ++k) r[k] = (blockIdx.x << 20) ^ (warp << 16) ^ (lane << 8) ^ (g * 16 + k) ^ 0x5a000000u;
    tm_st4(tbase + g * 16, r);
  }
  tm_wait_st();
  for (int g = 0; g < NCOLS / 16; ++g) {
    unsigned r[16];
    tm_ld4(tbase + g * 16, r);
    tm_wait_ld();
    for (int k = 0; k < 16; ++k)
      if (r[k] != ((blockIdx.x << 20) ^ (warp << 16) ^ (lane << 8) ^ (g * 16 + k) ^ 0x5a000000u)) ++bad;
  }
  if (bad) atomicAdd(errors, bad);
  // ---- throughput: `iters` x (8 loads of 4 pairs), all 8 warps of the SM busy ----
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      unsigned r[16];
      tm_ld4(tbase + g * 16, r);
      tm_wait_ld();
#pragma unroll
      for (int k = 0; k < 16; ++k) acc ^= r[k];
    }
  }
  long long t1 = clock64();
  // same bytes with LDS.128 (4 x 16 B per lane per group, conflict-free)
  uint4* my = smem + warp * (32 * 32);
  for (int p = 0; p < 32; ++p) my[p * 32 + lane] = make_uint4(lane, p, warp, 7);
  __syncwarp();
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 v;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(&my[(g * 4 + q) * 32 + lane]);
        asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sa));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  long long t3 = clock64();
  if (lane == 0) {
    cyc_tm[blockIdx.x * 4 + warp] = t1 - t0;
    cyc_lds[blockIdx.x * 4 + warp] = t3 - t2;
  }
  if (acc == 0x12345u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm_base_s), "n"(NCOLS) : "memory");
}

int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  const int grid = pr.multiProcessorCount * 2, iters = 2000;
  int* err; long long *ctm, *clds; unsigned* sink;
  cudaMallocManaged(&err, 4); cudaMallocManaged(&ctm, grid * 4 * 8); cudaMallocManaged(&clds, grid * 4 * 8);
  cudaMallocManaged(&sink, 4);
  *err = 0;
  const size_t smem = 4 * 32 * 32 * 16;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<grid, 128, smem>>>(iters, err, ctm, clds, sink);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s, round-trip errors %d\n", cudaGetErrorString(e), *err);
  double a = 0, b = 0;
  for (int i = 0; i < grid * 4; ++i) { a += ctm[i]; b += clds[i]; }
  a /= grid * 4; b /= grid * 4;
  const double bytes = (double)iters * 8 * 4 * 16 * 32;     // per warp
  printf("per warp: tcgen05.ld %.0f cycles (%.1f B/cyc/SM with 8 warps), LDS.128 %.0f cycles (%.1f B/cyc/SM)\n",
         a, 8 * bytes / a, b, 8 * bytes / b);
  return (*err != 0) || e != cudaSuccess;
}

// bw_probe.cu — read-only HBM bandwidth ceiling on this GPU for the access
// shapes the scan kernel can use (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint4 ldnc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// (a) grid-stride LDG.128, UNROLL independent loads per thread per iteration; each warp reads
// UNROLL contiguous 512-byte chunks (warp-contiguous) — the scan kernel's pattern.
template <int UNROLL, bool NOALLOC>
__global__ void __launch_bounds__(512) ldg_kernel(const uint4* __restrict__ p, size_t n_vec, uint32_t* out) {
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t acc = 0;
  const size_t chunk = 32 * UNROLL;  // vectors per warp iteration
  for (size_t base = warp_global * chunk; base + chunk <= n_vec; base += n_warps * chunk) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      v[u] = NOALLOC ? ldnc(p + base + u * 32 + lane) : __ldg(p + base + u * 32 + lane);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

// (b) TMA 1-D bulk copies into a shared-memory ring, one producer thread, consumers sum from smem.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((uint32_t)__cvta_generic_to_shared(bar)));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(phase));
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
                  "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

template <int STAGES, int STAGE_BYTES>
__global__ void __launch_bounds__(288) tma_kernel(const uint8_t* __restrict__ p, size_t n_bytes, uint32_t* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const int consumers = blockDim.x - 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], consumers / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  const size_t n_tiles = n_bytes / STAGE_BYTES;
  uint32_t acc = 0;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      int s = 0; uint32_t ph = 0;
      for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        bulk_g2s(smem + (size_t)s * STAGE_BYTES, p + t * STAGE_BYTES, STAGE_BYTES, &full[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    const int ct = threadIdx.x - 32;
    int s = 0; uint32_t ph = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      mbar_wait(&full[s], ph);
      const uint4* sp = (const uint4*)(smem + (size_t)s * STAGE_BYTES);
      for (int i = ct; i < STAGE_BYTES / 16; i += consumers) { uint4 v = sp[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static void timeit(const char* name, size_t bytes, F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f, tot = 0;
  const int reps = 20;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best; tot += ms;
  }
  CK(cudaGetLastError());
  printf("%-44s best %8.1f GB/s  avg %8.1f GB/s  (%.1f us)\n", name, bytes / best / 1e6, bytes / (tot / reps) / 1e6, best * 1e3);
}

int main() {
  const size_t bytes = 3072000000ull;
  uint8_t* d; uint32_t* out;
  CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(d, 1, bytes));
  const size_t n_vec = bytes / 16;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d, buffer %.3f GB\n", sms, bytes / 1e9);
  for (int mult = 1; mult <= 4; mult *= 2) {
    char nm[96];
    snprintf(nm, 96, "ldg.nc.noalloc u6  grid %dxSM x512", mult);
    timeit(nm, bytes, [&] { ldg_kernel<6, true><<<sms * mult, 512>>>((const uint4*)d, n_vec, out); });
    snprintf(nm, 96, "ldg.nc.noalloc u12 grid %dxSM x512", mult);
    timeit(nm, bytes, [&] { ldg_kernel<12, true><<<sms * mult, 512>>>((const uint4*)d, n_vec, out); });
    snprintf(nm, 96, "ldg (L1 alloc) u12 grid %dxSM x512", mult);
    timeit(nm, bytes, [&] { ldg_kernel<12, false><<<sms * mult, 512>>>((const uint4*)d, n_vec, out); });
  }
  timeit("ldg.nc.noalloc u24 grid 1xSM x512", bytes, [&] { ldg_kernel<24, true><<<sms, 512>>>((const uint4*)d, n_vec, out); });
  {
    auto k = tma_kernel<4, 32768>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
    timeit("tma bulk 4 x 32KB  grid 1xSM", bytes, [&] { k<<<sms, 288, 4 * 32768>>>(d, bytes, out); });
  }
  {
    auto k = tma_kernel<6, 32768>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768));
    timeit("tma bulk 6 x 32KB  grid 1xSM", bytes, [&] { k<<<sms, 288, 6 * 32768>>>(d, bytes, out); });
  }
  {
    auto k = tma_kernel<8, 16384>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384));
    timeit("tma bulk 8 x 16KB  grid 1xSM", bytes, [&] { k<<<sms, 288, 8 * 16384>>>(d, bytes, out); });
  }
  {
    auto k = tma_kernel<4, 16384>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384));
    timeit("tma bulk 4 x 16KB  grid 2xSM", bytes, [&] { k<<<sms * 2, 288, 4 * 16384>>>(d, bytes, out); });
  }
  {
    auto k = tma_kernel<12, 16384>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384));
    timeit("tma bulk 12 x 16KB grid 1xSM", bytes, [&] { k<<<sms, 288, 12 * 16384>>>(d, bytes, out); });
  }
  // plain cudaMemcpy D2D for reference (read + write)
  uint8_t* d2;
  CK(cudaMalloc(&d2, bytes / 2));
  timeit("cudaMemcpyAsync D2D (r+w bytes)", bytes, [&] { cudaMemcpyAsync(d2, d, bytes / 2, cudaMemcpyDeviceToDevice); });
  return 0;
}

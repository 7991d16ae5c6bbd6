// Read-bandwidth probe for the access patterns of the MokA streaming kernels (B200, sm_100a).
//   mode 0: plain coalesced 128-bit loads, grid-stride (the copy-kernel pattern, read only)
//   mode 1: TMA {64 cols x 128 rows} boxes, CTA = (token tile, column split), `kbox` boxes per stage,
//           `stages` stages in flight -- the k_down_tc / k_dgrad_tc pattern without any MMA
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_plain(const uint4* __restrict__ src, size_t n, unsigned* sink) {
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + i));
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

struct P { int T, cols, ksplit, kbox, stages, order; };

__global__ void __launch_bounds__(64) k_tma(const __grid_constant__ CUtensorMap tm, P p, unsigned* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t stage_bytes = 16384u * p.kbox;
  uint64_t* full = (uint64_t*)(ring + (size_t)p.stages * stage_bytes);
  uint64_t* empty = full + p.stages;
  const int tile = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int nkb = p.cols / (64 * p.kbox);
  const int kb0 = (int)((long long)split * nkb / p.ksplit), kb1 = (int)((long long)(split + 1) * nkb / p.ksplit);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&empty[s])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto wait = [](uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  };
  if (threadIdx.x == 0) {
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % p.stages;
      wait(&empty[s], (((uint32_t)(it / p.stages)) & 1u) ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
      for (int bx = 0; bx < p.kbox; ++bx) {
        const int col = (kb * p.kbox + bx) * 64;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                     :: "r"(smem_u32(ring + (size_t)s * stage_bytes + bx * 16384)), "l"((uint64_t)&tm), "r"(smem_u32(&full[s])),
                        "r"(col), "r"(tile * 128), "l"(0x12F0000000000000ull) : "memory");
      }
    }
  } else if (threadIdx.x == 32) {
    for (int kb = kb0; kb < kb1; ++kb) {
      const int it = kb - kb0, s = it % p.stages;
      wait(&full[s], ((uint32_t)(it / p.stages)) & 1u);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
  if (ring[threadIdx.x] == 0x77 && ring[1000] == 0x13) *sink = 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 8192, cols = 4096;
  const size_t bytes = (size_t)T * cols * 2;
  const int NB = 3;
  std::vector<void*> bufs(NB);
  for (auto& b : bufs) { CK(cudaMalloc(&b, bytes)); CK(cudaMemset(b, 1, bytes)); }
  unsigned* sink; CK(cudaMalloc(&sink, 4));
  void* fn_p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn_p, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn_p;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](auto launch, const char* name) {
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    const int iters = 30;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch(i);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / iters;
    printf("%-52s %8.2f us  %8.1f GB/s\n", name, us, bytes / us / 1e3);
  };
  for (int grid : {148, 296, 592, 1184}) {
    char nm[96]; snprintf(nm, 96, "plain ld.v4 grid=%d x 512", grid);
    timeit([&](int i) { k_plain<<<grid, 512>>>((const uint4*)bufs[i % NB], bytes / 16, sink); }, nm);
  }
  for (int promo : {0, 2}) {                     // L2 promotion none / 256B
    for (int ksplit : {1, 2, 4}) {
      for (int kbox : {1, 2, 4}) {
        for (int stages : {4, 8}) {
          if (16384 * kbox * stages > 200 * 1024) continue;
          std::vector<CUtensorMap> tms(NB);
          for (int i = 0; i < NB; ++i) {
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)T};
            cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
            cuuint32_t box[2] = {64, 128};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = enc(&tms[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bufs[i], dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
          }
          P p{T, cols, ksplit, kbox, stages, 0};
          const size_t smem = (size_t)stages * kbox * 16384 + 1024 + 256;
          CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          char nm[96];
          snprintf(nm, 96, "tma promo=%d ksplit=%d (grid %d) kbox=%d stages=%d", promo, ksplit, T / 128 * ksplit, kbox, stages);
          timeit([&](int i) { k_tma<<<T / 128 * ksplit, 64, smem>>>(tms[i % NB], p, sink); }, nm);
        }
      }
    }
  }
  return 0;
}

// tcgen05 + TMA path of the MokA adapter layer (bf16 storage), sm_100a.
//
// Every kernel here is a skinny GEMM that streams one [tokens x features] bf16 matrix through
// shared memory with TMA (128-byte swizzle), feeds it to tcgen05.mma with the tiny adapter
// operand and keeps fp32 accumulators in TMEM.  The kernels are HBM-bound by design; the tensor
// pipe is used because CUDA-core FMA cannot keep up with 6.5 TB/s once r >= 16
// (6 r (d_in+d_out) flop per token vs 2 (4 d_in + 3 d_out) bytes).
//
//   k_down_tc : H[128 tok, M*r] = X[128 tok, d_in] . [A_0;A_1;A_2]^T     (row-select by modality)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (TMEM lane quadrant = warp % 4).
#include <mutex>

#include "moka_attn.cuh"
#include "moka_common.cuh"
#include "moka_kernels.h"
#include "moka_ptx.cuh"

namespace moka {

// ---------------------------------------------------------------------------- host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Row-major bf16 matrix [rows, cols] -> tensor map with a {64 cols, box_rows} box, 128B swizzle.
static bool make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static inline uint32_t tmem_cols_for(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

constexpr int kTcThreads = 192;

// ---------------------------------------------------------------------------- k_down_tc
struct DownTcParams {
  int T, d_in, r, rp, M, NP;       // rp = r rounded to 8 (rows per modality block), NP = padded N
  int ksplit, kblocks;             // split of the d_in reduction across CTAs; kblocks = d_in / 64
  int stages;
  uint32_t stage_bytes, w_bytes, tmem_cols;
  float scale;                     // pre * dropout scale
};

__global__ void __launch_bounds__(kTcThreads)
k_down_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_a0,
          const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_a2,
          DownTcParams p, const uint8_t* __restrict__ tokmeta, float* __restrict__ h_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tmem_full = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int t0 = tile * 128;
  const int kb0 = (int)((long long)split * p.kblocks / p.ksplit);
  const int kb1 = (int)((long long)(split + 1) * p.kblocks / p.ksplit);

  // rows of the weight tiles that TMA never writes (padding between modality blocks) must be 0
  for (int s = 0; s < p.stages; ++s) {
    uint4* w = reinterpret_cast<uint4*>(ring + (size_t)s * p.stage_bytes + 16384);
    for (uint32_t i = threadIdx.x; i < p.w_bytes / 16; i += kTcThreads) w[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_x); ptx::prefetch_tmap(&tm_a0);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, p.tmem_cols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = 16384u + (uint32_t)p.M * (uint32_t)p.r * 128u;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait_or_trap(&empty[s], ph ^ 1u);
        uint8_t* xs = ring + (size_t)s * p.stage_bytes;
        uint8_t* ws = xs + 16384;
        ptx::mbar_arrive_expect_tx(&full[s], tx);
        ptx::tma_load_2d(xs, &tm_x, &full[s], kb * 64, t0, ptx::kEvictFirst);
        ptx::tma_load_2d(ws, &tm_a0, &full[s], kb * 64, 0, ptx::kEvictLast);
        if (p.M > 1) ptx::tma_load_2d(ws + (size_t)p.rp * 128, &tm_a1, &full[s], kb * 64, 0, ptx::kEvictLast);
        if (p.M > 2) ptx::tma_load_2d(ws + (size_t)2 * p.rp * 128, &tm_a2, &full[s], kb * 64, 0, ptx::kEvictLast);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // N > 256 does not fit one instruction: two halves (each a multiple of 16)
      const int nparts = p.NP > 256 ? 2 : 1;
      const int npart = p.NP / nparts;
      const uint32_t idesc = ptx::idesc_bf16(128, (uint32_t)npart, 0, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait_or_trap(&full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t xa = ptx::smem_u32(ring + (size_t)s * p.stage_bytes);
        const uint32_t wa = xa + 16384u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adesc = ptx::smem_desc_sw128(xa + k * 32, 1024, 0);
          for (int part = 0; part < nparts; ++part) {
            const uint64_t bdesc = ptx::smem_desc_sw128(wa + (uint32_t)part * npart * 128u + k * 32, 1024, 0);
            ptx::umma_bf16(tmem_base + (uint32_t)part * npart, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(&empty[s]);
      }
      ptx::umma_commit(tmem_full);
    }
  } else {
    // ---- epilogue: thread <-> token row (TMEM lane); pick the columns of the token's modality
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const int code = (t < p.T) ? (tokmeta[t] & MOKA_MOD_MASK) : -1;
    ptx::mbar_wait_or_trap(tmem_full, 0);
    ptx::tc_fence_after_sync();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int m = 0; m < p.M; ++m) {
      for (int c8 = 0; c8 < p.rp; c8 += 8) {
        float v[8];
        ptx::tmem_ld8(lane_base + (uint32_t)(m * p.rp + c8), v);
        ptx::tmem_ld_wait();
        if (code == m + 1) {
          float* dst = h_out + (size_t)t * p.r + c8;
          if (p.ksplit == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c8 + i < p.r) dst[i] = v[i] * p.scale;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c8 + i < p.r) red_add_f32(dst + i, v[i] * p.scale);
          }
        }
      }
    }
    if (code == 0 && p.ksplit == 1)
      for (int j = 0; j < p.r; ++j) h_out[(size_t)t * p.r + j] = 0.f;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

static int check_launch_tc() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

static int launch_down_tc(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st) {
  DownTcParams p;
  p.T = d.T; p.d_in = d.d_in; p.r = d.r; p.rp = round_up(d.r, 8); p.M = d.M;
  p.NP = round_up(p.M * p.rp, 16);
  if (p.NP > 256) p.NP = round_up(p.NP, 32);
  p.kblocks = d.d_in / 64;
  const int ntiles = (d.T + 127) / 128;
  p.ksplit = (ntiles * 2 <= num_sms * 2 && p.kblocks >= 8 && ntiles < num_sms) ? 2 : 1;
  p.w_bytes = (uint32_t)p.NP * 128u;
  p.stage_bytes = 16384u + p.w_bytes;
  const size_t budget = 100 * 1024;
  p.stages = (int)(budget / p.stage_bytes);
  if (p.stages > 8) p.stages = 8;
  if (p.stages < 2) p.stages = 2;
  p.tmem_cols = tmem_cols_for(p.NP);
  p.scale = d.pre * d.drop_scale;
  CUtensorMap tx, ta[3];
  if (!make_tmap(&tx, a.x, (uint64_t)d.T, (uint64_t)d.d_in, 128)) return MOKA_EUNSUPPORTED;
  for (int m = 0; m < 3; ++m) {
    const void* base = a.A[m < d.M ? m : 0];
    if (!make_tmap(&ta[m], base, (uint64_t)d.r, (uint64_t)d.d_in, (uint32_t)d.r)) return MOKA_EUNSUPPORTED;
  }
  if (p.ksplit > 1) {
    cudaError_t e = cudaMemsetAsync(a.h_stash, 0, (size_t)d.T * d.r * sizeof(float), st);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  cudaError_t e = cudaFuncSetAttribute(k_down_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  k_down_tc<<<ntiles * p.ksplit, kTcThreads, smem, st>>>(tx, ta[0], ta[1], ta[2], p, a.tokmeta, a.h_stash);
  return check_launch_tc();
}

// ---------------------------------------------------------------------------- dispatch
int tc_stage_mask(const DevDesc& d, int dtype) {
  if (dtype != MOKA_BF16) return 0;                 // fp32 storage keeps exact fp32 FMA (1e-5 bar)
  if (d.drop_thresh) return 0;                      // dropout is applied on the warp-FMA path
  if (d.d_in % 64 || d.d_out % 64) return 0;
  if (d.r % 8 || d.r > 128) return 0;
  if (!get_encode_fn()) return 0;
  return kStageDown;
}

int tc_forward(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  if (stages & kStageDown) {
    if (int e = launch_down_tc(d, a, num_sms, st)) return e;
  }
  if (stages & kStageUp) return MOKA_EUNSUPPORTED;
  return MOKA_OK;
}

int tc_backward(const DevDesc&, const SimtArgs&, int, cudaStream_t, int) { return MOKA_EUNSUPPORTED; }

}  // namespace moka

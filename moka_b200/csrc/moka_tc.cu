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
#include <algorithm>
#include <cstdlib>
#include <cstring>
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

// Same for a column window of a wider matrix: `cols` columns starting at `base`, rows `row_stride` elements apart.
static bool make_tmap_strided(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride,
                              uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride * 2};
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
constexpr int kWorkers = 128;                                   // worker threads of a group (one per TMEM lane)
constexpr int kWorkerWarps = kWorkers / 32;

// One mbarrier arrival per WARP instead of one per thread (a stage barrier with 128-512 per-thread arrivals is
// 128-512 serialised shared-memory atomics on one word): every lane first fences its own shared-memory writes /
// TMEM reads, __syncwarp() orders them before lane 0's release-arrive.  Barriers signalled this way are
// initialised with the number of WARPS.
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(bar);
}

// Launch with programmatic stream serialization (see moka_ptx.cuh: pdl_wait / pdl_trigger).
// MOKA_NO_PDL=1 falls back to plain stream order (A/B measurements).
static int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MOKA_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess) (void)cudaGetLastError();               // reported through the return value: do not leave it behind
  return e;
}

// Optional in-kernel timestamps (cycle counter of the SM) for tuning: slot layout is per kernel.
// Written only when MOKA_DEBUG_TIMING=1; read back through moka_debug_read().
constexpr int kDbgSlots = 128;                                  // per CTA: 0..63 cycle stamps, 100.. wall-clock timeline
constexpr int kDbgCtas = 148 * 4;
__device__ unsigned long long g_dbg[kDbgCtas * kDbgSlots];
__device__ __forceinline__ void dbg_stamp(int enabled, int cta, int slot) {
  if ((enabled & 1) && cta < kDbgCtas && slot < 64) g_dbg[cta * kDbgSlots + slot] = clock64();
}
// wall-clock (globaltimer, ns) stamps: comparable across CTAs; slots 60 (entry) / 61 (exit) by convention
__device__ __forceinline__ void dbg_stamp_g(int enabled, int cta, int slot) {
  if ((enabled & 1) && cta < kDbgCtas && slot < kDbgSlots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_dbg[cta * kDbgSlots + slot] = t;
  }
}
static int dbg_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MOKA_DEBUG_TIMING"); v = (e && e[0] == '1') ? 1 : 0; }
  return v;
}
int tc_dbg_enabled() { return dbg_enabled(); }
int debug_set_trap_buffer(unsigned long long* mapped) {
  return cudaMemcpyToSymbol(ptx::g_trap_buf, &mapped, sizeof(mapped)) == cudaSuccess ? MOKA_OK : MOKA_ECUDA;
}
// tuning aid: a one-thread kernel that stamps globaltimer into slot `slot` of the LAST debug row (stream-ordered marker)
__global__ void k_dbg_mark(int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_dbg[(kDbgCtas - 1) * kDbgSlots + (slot & (kDbgSlots - 1))] = t;
}
int debug_mark(int slot, cudaStream_t st) {
  k_dbg_mark<<<1, 1, 0, st>>>(slot);
  return cudaGetLastError() == cudaSuccess ? MOKA_OK : MOKA_ECUDA;
}
int debug_read(unsigned long long* host, int n) {
  if (n > kDbgCtas * kDbgSlots) n = kDbgCtas * kDbgSlots;
  return cudaMemcpyFromSymbol(host, g_dbg, (size_t)n * 8) == cudaSuccess ? MOKA_OK : MOKA_ECUDA;
}

// ---------------------------------------------------------------------------- k_down_tc
// Column-order rotation of the streaming kernels: every CTA walks its column blocks starting at a different one
// (wrapping around), so that the small weight tiles all CTAs need (A_m boxes in k_down_tc, B^T blocks in
// k_dgrad_tc) are not requested by every SM from the same L2 lines at the same moment.  MOKA_NO_ROT=1: off.
// Early keep bits: the masking warps generate the keep bits of a stage BEFORE they wait for its data (the bits
// depend on (row, column group) only), leaving just the sparse zero stores and the fence after the TMA arrival.
// Same bits, same result.  Timed in round 2: on by default in both kernels (inside the step the masking sits on the
// x-slot turnaround of k_wgrad_tc); MOKA_DROP_EARLY=0/1 forces both.
static int drop_early_bits(int dflt) {
  static int v = -2;                               // -1: not set in the environment -> per-kernel default
  if (v == -2) { const char* e = getenv("MOKA_DROP_EARLY"); v = e ? ((e[0] == '1') ? 1 : 0) : -1; }
  return v < 0 ? dflt : v;
}

static int rot_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MOKA_NO_ROT"); v = (e && e[0] == '1') ? 0 : 1; }
  return v;
}

// (i + rot) mod n for 0 <= i, rot < n without the integer division a `%` by a run-time n costs in every stage
__device__ __forceinline__ int rot_index(int i, int rot, int n) {
  const int k = i + rot;
  return k >= n ? k - n : k;
}

struct DownTcParams {
  int T, d_in, r, rp, M, NP;       // rp = r rounded to 8 (rows per modality block), NP = padded N
  int ksplit, kblocks;             // split of the d_in reduction across CTAs; kblocks = d_in / (64 kbox)
  int kbox;                        // 64-column TMA boxes per pipeline stage (1 or 2)
  int rot;                         // rotate the column-block order per CTA (see rot_enabled)
  int early_bits;                  // DROP: generate a stage's keep bits before waiting for its data (see drop_early_bits)
  int stages;
  uint32_t stage_bytes, w_bytes, tmem_cols;
  float scale;                     // pre * dropout scale
  int dbg;
};

// DROP: lora_dropout > 0.  The worker warps (idle until the epilogue otherwise) apply the keep mask to every x
// tile in shared memory between its TMA arrival (full) and its use by the tensor core (ready): thread <-> token
// row, one 16-byte chunk = one 8-column group of the generator; the 1/(1-p) factor rides in p.scale.
constexpr int kDownDropThreads = kTcThreads + 384;              // + 12 warps that only apply the dropout mask
template <bool DROP>
__global__ void __launch_bounds__(DROP ? kDownDropThreads : kTcThreads)
k_down_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_a0,
          const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_a2,
          DownTcParams p, DevDesc d, const uint8_t* __restrict__ tokmeta, float* __restrict__ h_out0,
          float* __restrict__ h_out1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* tmem_full = bars + 2 * p.stages;
  uint64_t* ready = bars + 2 * p.stages + 1;                  // [stages] tile masked (DROP only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 1);

  const int warp = ptx::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int t0 = tile * 128;
  if (threadIdx.x == 0) { dbg_stamp_g(p.dbg, blockIdx.x, 100); dbg_stamp(p.dbg, blockIdx.x, 0); }
  const int kb0 = (int)((long long)split * p.kblocks / p.ksplit);
  const int kb1 = (int)((long long)(split + 1) * p.kblocks / p.ksplit);
  const int nkb = kb1 - kb0;
  const int krot = p.rot ? (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nkb) : 0;   // first block of this CTA

  // rows of the weight tiles that TMA never writes (padding between modality blocks) must be 0
  const uint32_t xbytes = 16384u * (uint32_t)p.kbox;
  for (int s = 0; s < p.stages; ++s) {
    uint4* w = reinterpret_cast<uint4*>(ring + (size_t)s * p.stage_bytes + xbytes);
    for (uint32_t i = threadIdx.x; i < p.w_bytes * p.kbox / 16; i += blockDim.x) w[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); ptx::mbar_init(&ready[s], 4 * kWorkerWarps);
    }
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
  if (threadIdx.x == 0) dbg_stamp(p.dbg, blockIdx.x, 1);
  // x may come straight from the kernel launched before this one: wait, then let k_up_tc start
  ptx::pdl_wait();
  ptx::pdl_trigger();
  if (threadIdx.x == 0) dbg_stamp(p.dbg, blockIdx.x, 2);

  if (warp == 0) {
    // ---- TMA producer: whole warp, warp-uniform values, one elected lane issues (see moka_ptx.cuh)
    {
      const uint32_t tx = (16384u + (uint32_t)p.M * (uint32_t)p.r * 128u) * (uint32_t)p.kbox;
      const uint32_t ring_a = ptx::uniform_u32(ptx::smem_u32(ring));
      for (int it = 0; it < nkb; ++it) {
        const int kb = kb0 + rot_index(it, krot, nkb), s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait_or_trap(&empty[s], ph ^ 1u, __LINE__);
        const uint32_t xs = ring_a + (uint32_t)s * p.stage_bytes;
        ptx::mbar_arrive_expect_tx_w(&full[s], tx);
        for (int bx = 0; bx < p.kbox; ++bx) {
          const int col = (kb * p.kbox + bx) * 64;
          const uint32_t ws = xs + xbytes + (uint32_t)bx * p.w_bytes;
          ptx::tma_load_2d_w(xs + (uint32_t)bx * 16384u, &tm_x, &full[s], col, t0, ptx::kEvictFirst);
          ptx::tma_load_2d_w(ws, &tm_a0, &full[s], col, 0, ptx::kEvictLast);
          if (p.M > 1) ptx::tma_load_2d_w(ws + (uint32_t)p.rp * 128u, &tm_a1, &full[s], col, 0, ptx::kEvictLast);
          if (p.M > 2) ptx::tma_load_2d_w(ws + 2u * (uint32_t)p.rp * 128u, &tm_a2, &full[s], col, 0, ptx::kEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp, one elected lane issues
    {
      const uint32_t tb = ptx::uniform_u32(tmem_base);
      const uint32_t ring_a = ptx::uniform_u32(ptx::smem_u32(ring));
      // N > 256 does not fit one instruction: two halves (each a multiple of 16)
      const int nparts = p.NP > 256 ? 2 : 1;
      const int npart = p.NP / nparts;
      const uint32_t idesc = ptx::idesc_bf16(128, (uint32_t)npart, 0, 0);
      const uint64_t dbase = ptx::smem_desc_sw128(0, 1024, 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int it = kb - kb0, s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        ptx::mbar_wait_or_trap(&full[s], ph, __LINE__);                    // weights + x landed
        if (lane == 0 && it < 8) dbg_stamp(p.dbg, blockIdx.x, 8 + it * 3);
        if (DROP) ptx::mbar_wait_or_trap(&ready[s], ph, __LINE__);           // ... and x masked
        if (lane == 0) { if (it < 8) dbg_stamp(p.dbg, blockIdx.x, 9 + it * 3); if (it == 0) dbg_stamp(p.dbg, blockIdx.x, 3); }
        ptx::tc_fence_after_sync();
        const uint32_t xa0 = ring_a + (uint32_t)s * p.stage_bytes;
        for (int bx = 0; bx < p.kbox; ++bx) {
          const uint32_t xa = xa0 + (uint32_t)bx * 16384u;
          const uint32_t wa = xa0 + xbytes + (uint32_t)bx * p.w_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = dbase + (uint64_t)((xa + k * 32) >> 4);
            for (int part = 0; part < nparts; ++part) {
              const uint64_t bdesc = dbase + (uint64_t)((wa + (uint32_t)part * npart * 128u + k * 32) >> 4);
              ptx::umma_bf16_w(tb + (uint32_t)part * npart, adesc, bdesc, idesc, (it > 0 || bx > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        ptx::umma_commit_w(&empty[s]);
        if (lane == 0 && it < 8) dbg_stamp(p.dbg, blockIdx.x, 10 + it * 3);
      }
      ptx::umma_commit_w(tmem_full);
      if (lane == 0) dbg_stamp(p.dbg, blockIdx.x, 4);
    }
  } else {
    // ---- epilogue: thread <-> token row (TMEM lane); pick the columns of the token's modality
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const int code = (t < p.T) ? (tokmeta[t] & MOKA_MOD_MASK) : -1;
    if (DROP) {
      // dropout on the x tiles, in place in shared memory.  16 warps (the 4 epilogue warps + 12 that only do
      // this): a single warp per scheduler runs these dependent integer chains at ~5 cycles per instruction,
      // four of them hide the latencies.  thread <-> (token row, every 4th 32-column super-group of the stage)
      const int mt = threadIdx.x - 64;                         // 0 .. 511
      const int mrow = mt & 127, part = mt >> 7;
      const unsigned long long rowbase = dropout_row_base(d, t0 + mrow);
      const int nsg = 2 * p.kbox;                              // super-groups per stage (<= 4: at most one per thread)
      const bool has_item = part < nsg;
      for (int it = 0; it < nkb; ++it) {
        const int kb = kb0 + rot_index(it, krot, nkb), s = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        // the keep bits depend on (row, column group) only, not on the data: with early_bits they are generated
        // BEFORE the wait for the stage, so that what follows the TMA arrival is just the zero stores and the fence
        uint32_t k32 = 0xffffffffu;
        if (has_item && p.early_bits) k32 = dropout_keep32(d, rowbase + (unsigned long long)(kb * nsg + part));
        ptx::mbar_sleep_or_trap(&full[s], ph, __LINE__);
        if (mt == 0 && it < 8) dbg_stamp(p.dbg, blockIdx.x, 32 + it * 3);
        if (has_item && !p.early_bits) k32 = dropout_keep32(d, rowbase + (unsigned long long)(kb * nsg + part));
        const uint32_t xs = ptx::smem_u32(ring + (size_t)s * p.stage_bytes);
        if (has_item)
          dropout_zero_dropped(k32, xs + (uint32_t)(part >> 1) * 16384u + (uint32_t)mrow * 128u, (uint32_t)mrow & 7u,
                               (uint32_t)(part & 1));
        if (mt == 0 && it < 8) dbg_stamp(p.dbg, blockIdx.x, 33 + it * 3);
        ptx::fence_proxy_async_smem();                           // generic writes -> visible to the tensor core
        warp_arrive(&ready[s]);
        if (mt == 0 && it < 8) dbg_stamp(p.dbg, blockIdx.x, 34 + it * 3);
      }
    }
    if (warp < 6) {                                              // (the extra masking warps have no epilogue role)
    ptx::mbar_wait_or_trap(tmem_full, 0, __LINE__);
    if (threadIdx.x == 64) dbg_stamp(p.dbg, blockIdx.x, 5);
    ptx::tc_fence_after_sync();
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // each split of the d_in reduction writes its own partial buffer with plain stores (the consumer,
    // k_up_tc, adds the two): no zero-fill, no atomics, bit-reproducible
    float* h_out = split == 0 ? h_out0 : h_out1;
    for (int m = 0; m < p.M; ++m) {
      for (int c8 = 0; c8 < p.rp; c8 += 8) {
        float v[8];
        ptx::tmem_ld8(lane_base + (uint32_t)(m * p.rp + c8), v);
        ptx::tmem_ld_wait();
        if (code == m + 1) {
          float* dst = h_out + (size_t)t * p.r + c8;
          if (c8 + 8 <= p.r) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0] * p.scale, v[1] * p.scale, v[2] * p.scale, v[3] * p.scale);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * p.scale, v[5] * p.scale, v[6] * p.scale, v[7] * p.scale);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c8 + i < p.r) dst[i] = v[i] * p.scale;
          }
        }
      }
    }
    if (code == 0)
      for (int j = 0; j < p.r; ++j) h_out[(size_t)t * p.r + j] = 0.f;
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) { dbg_stamp(p.dbg, blockIdx.x, 6); dbg_stamp_g(p.dbg, blockIdx.x, 101); }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------- k_down_ts
// Second form of the down-projection (round 2, last session): the x tile never feeds the tensor core from shared
// memory.  Sixteen worker warps read the TMA-written rows (thread <-> token row x 32-column super-group, 4 x 16 bytes),
// apply the dropout keep bits IN REGISTERS and store the rows into TMEM (two bf16 per 32-bit cell); the product takes
// its A operand from there (tcgen05.mma TS form), B = the weight tile in shared memory as before.  What that buys:
//   * no generic-proxy write of the tile and therefore no fence.proxy.async on the TMA -> MMA path;
//   * the x slot is handed back to the TMA producer as soon as the workers have READ it - its life no longer
//     includes the masking hand-off, the products and the commit round trip (~2.5k of ~4.5k cycles per stage inside
//     the training step, where the SM clock is 1.5 GHz) - so the same shared memory holds ~2x the bytes in flight;
//   * x and the weight tiles live in two rings of their own depth, and the staged A operands in a third (TMEM).
// Measured (profiles/r02_dropout_cost.md): alone 52.4 -> 49.8 us at d_in = 4096 and 128.6 -> 123.1 us at 11008
// (T = 32768); inside the 7B step 1-2 us per forward call - there the dropout path is bound by the instruction issue
// of the keep-bit generator and the masking arithmetic, not by this hand-off.
struct DownTsParams {
  int T, d_in, r, rp, M, NP;
  int ksplit, kblocks;             // kblocks = d_in / 128 (a stage is always two 64-column boxes)
  int rot;
  int xs, ws, ab;                  // ring depths: x stages, weight stages, TMEM A buffers (64 columns each)
  int drop;                        // 0: keep everything (lora_dropout = 0)
  int a_col0;                      // first TMEM column of the A buffers
  uint32_t w_bytes, tmem_cols;     // w_bytes = NP * 128 (one 64-column weight box)
  float scale;
  int dbg;
};
constexpr int kDownTsThreads = 64 + 512;

__global__ void __launch_bounds__(kDownTsThreads)
k_down_ts(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_a0,
          const __grid_constant__ CUtensorMap tm_a1, const __grid_constant__ CUtensorMap tm_a2,
          DownTsParams p, DevDesc d, const uint8_t* __restrict__ tokmeta, float* __restrict__ h_out0,
          float* __restrict__ h_out1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* xring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wring = xring + (size_t)p.xs * 32768;
  const uint32_t wstage = 2u * p.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wring + (size_t)p.ws * wstage);
  uint64_t* full_x = bars;
  uint64_t* x_empty = full_x + p.xs;
  uint64_t* full_w = x_empty + p.xs;
  uint64_t* w_empty = full_w + p.ws;
  uint64_t* ready = w_empty + p.ws;                             // [ab] A operand staged in TMEM
  uint64_t* a_empty = ready + p.ab;                             // [ab] ... and consumed
  uint64_t* tmem_full = a_empty + p.ab;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = ptx::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int t0 = tile * 128;
  if (threadIdx.x == 0) { dbg_stamp_g(p.dbg, blockIdx.x, 100); dbg_stamp(p.dbg, blockIdx.x, 0); }
  const int kb0 = (int)((long long)split * p.kblocks / p.ksplit);
  const int kb1 = (int)((long long)(split + 1) * p.kblocks / p.ksplit);
  const int nkb = kb1 - kb0;
  const int krot = p.rot ? (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nkb) : 0;

  // rows of the weight tiles that TMA never writes (padding between modality blocks) must be 0
  {
    uint4* w = reinterpret_cast<uint4*>(wring);
    for (uint32_t i = threadIdx.x; i < (uint32_t)p.ws * wstage / 16; i += blockDim.x) w[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.xs; ++s) { ptx::mbar_init(&full_x[s], 1); ptx::mbar_init(&x_empty[s], 4 * kWorkerWarps); }
    for (int s = 0; s < p.ws; ++s) { ptx::mbar_init(&full_w[s], 1); ptx::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < p.ab; ++s) { ptx::mbar_init(&ready[s], 4 * kWorkerWarps); ptx::mbar_init(&a_empty[s], 1); }
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
  ptx::pdl_wait();
  ptx::pdl_trigger();
  if (threadIdx.x == 0) dbg_stamp(p.dbg, blockIdx.x, 2);

  if (warp == 0) {
    // ---- TMA producer: weights of block `it` as soon as their stage is free, then x
    const uint32_t wtx = (uint32_t)p.M * (uint32_t)p.r * 128u * 2u;
    const uint32_t xr_a = ptx::uniform_u32(ptx::smem_u32(xring));
    const uint32_t wr_a = ptx::uniform_u32(ptx::smem_u32(wring));
    int sx = 0, sw = 0; uint32_t xph = 0, wph = 0;
    for (int it = 0; it < nkb; ++it) {
      const int kb = kb0 + rot_index(it, krot, nkb);
      ptx::mbar_wait_or_trap(&w_empty[sw], wph ^ 1u, __LINE__);
      ptx::mbar_arrive_expect_tx_w(&full_w[sw], wtx);
      for (int bx = 0; bx < 2; ++bx) {
        const int col = (kb * 2 + bx) * 64;
        const uint32_t ws_a = wr_a + (uint32_t)sw * wstage + (uint32_t)bx * p.w_bytes;
        ptx::tma_load_2d_w(ws_a, &tm_a0, &full_w[sw], col, 0, ptx::kEvictLast);
        if (p.M > 1) ptx::tma_load_2d_w(ws_a + (uint32_t)p.rp * 128u, &tm_a1, &full_w[sw], col, 0, ptx::kEvictLast);
        if (p.M > 2) ptx::tma_load_2d_w(ws_a + 2u * (uint32_t)p.rp * 128u, &tm_a2, &full_w[sw], col, 0, ptx::kEvictLast);
      }
      ptx::mbar_wait_or_trap(&x_empty[sx], xph ^ 1u, __LINE__);
      ptx::mbar_arrive_expect_tx_w(&full_x[sx], 32768u);
      const uint32_t xs_a = xr_a + (uint32_t)sx * 32768u;
      ptx::tma_load_2d_w(xs_a, &tm_x, &full_x[sx], kb * 128, t0, ptx::kEvictFirst);
      ptx::tma_load_2d_w(xs_a + 16384u, &tm_x, &full_x[sx], kb * 128 + 64, t0, ptx::kEvictFirst);
      if (++sx == p.xs) { sx = 0; xph ^= 1u; }
      if (++sw == p.ws) { sw = 0; wph ^= 1u; }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: A from TMEM, B = weight tile
    const uint32_t tb = ptx::uniform_u32(tmem_base);
    const uint32_t wr_a = ptx::uniform_u32(ptx::smem_u32(wring));
    const int nparts = p.NP > 256 ? 2 : 1;
    const int npart = p.NP / nparts;
    const uint32_t idesc = ptx::idesc_bf16(128, (uint32_t)npart, 0, 0);
    const uint64_t dbase = ptx::smem_desc_sw128(0, 1024, 0);
    int sw = 0, ab = 0; uint32_t wph = 0, aph = 0;
    for (int it = 0; it < nkb; ++it) {
      ptx::mbar_wait_or_trap(&full_w[sw], wph, __LINE__);
      ptx::mbar_wait_or_trap(&ready[ab], aph, __LINE__);
      if (lane == 0 && it == 0) dbg_stamp(p.dbg, blockIdx.x, 3);
      ptx::tc_fence_after_sync();
      const uint32_t ta = tb + (uint32_t)p.a_col0 + (uint32_t)ab * 64u;
      for (int bx = 0; bx < 2; ++bx) {
        const uint32_t wa = wr_a + (uint32_t)sw * wstage + (uint32_t)bx * p.w_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          for (int part = 0; part < nparts; ++part) {
            const uint64_t bdesc = dbase + (uint64_t)((wa + (uint32_t)part * npart * 128u + k * 32) >> 4);
            ptx::umma_bf16_ts_w(tb + (uint32_t)part * npart, ta + (uint32_t)(bx * 4 + k) * 8u, bdesc, idesc,
                                (it > 0 || bx > 0 || k > 0) ? 1u : 0u);
          }
        }
      }
      ptx::umma_commit_w(&w_empty[sw]);
      ptx::umma_commit_w(&a_empty[ab]);
      if (++sw == p.ws) { sw = 0; wph ^= 1u; }
      if (++ab == p.ab) { ab = 0; aph ^= 1u; }
    }
    ptx::umma_commit_w(tmem_full);
    if (lane == 0) dbg_stamp(p.dbg, blockIdx.x, 4);
  } else {
    // ---- workers: warp w <-> TMEM lane quadrant w % 4 (the only lanes it may write), super-group (w - 2) / 4
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int t = t0 + row;
    const int code = (t < p.T) ? (tokmeta[t] & MOKA_MOD_MASK) : -1;
    const int part = (warp - 2) >> 2;                            // 0..3
    const unsigned long long rowbase = dropout_row_base(d, t);
    const uint32_t row7 = (uint32_t)row & 7u;
    const uint32_t xrow0 = ptx::smem_u32(xring) + (uint32_t)(part >> 1) * 16384u + (uint32_t)row * 128u;
    const uint32_t ta_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)p.a_col0 + (uint32_t)part * 16u;
    int sx = 0, ab = 0; uint32_t xph = 0, aph = 0;
    for (int it = 0; it < nkb; ++it) {
      const int kb = kb0 + rot_index(it, krot, nkb);
      uint32_t k32 = 0xffffffffu;
      if (p.drop) k32 = dropout_keep32(d, rowbase + (unsigned long long)(kb * 4 + part));
      ptx::mbar_sleep_or_trap(&full_x[sx], xph, __LINE__);    // (polling instead: no difference, measured)
      const uint32_t xrow = xrow0 + (uint32_t)sx * 32768u;
      uint4 v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = ptx::lds128(xrow + ((((uint32_t)(part & 1) * 4u + (uint32_t)c) ^ row7) << 4));
      uint32_t w16[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t W[4];
        dropout_expand8((k32 >> (8 * c)) & 0xffu, W);
        w16[c * 4 + 0] = v[c].x & W[0]; w16[c * 4 + 1] = v[c].y & W[1];
        w16[c * 4 + 2] = v[c].z & W[2]; w16[c * 4 + 3] = v[c].w & W[3];
      }
      warp_arrive(&x_empty[sx]);                                 // the rows are in registers: the slot may be reloaded
      ptx::mbar_wait_or_trap(&a_empty[ab], aph ^ 1u, __LINE__);
      ptx::tc_fence_after_sync();
      ptx::tmem_st16(ta_row + (uint32_t)ab * 64u, w16);
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      warp_arrive(&ready[ab]);
      if (++sx == p.xs) { sx = 0; xph ^= 1u; }
      if (++ab == p.ab) { ab = 0; aph ^= 1u; }
    }
    if (warp < 6) {
      ptx::mbar_wait_or_trap(tmem_full, 0, __LINE__);
      if (threadIdx.x == 64) dbg_stamp(p.dbg, blockIdx.x, 5);
      ptx::tc_fence_after_sync();
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      float* h_out = split == 0 ? h_out0 : h_out1;
      for (int m = 0; m < p.M; ++m) {
        for (int c8 = 0; c8 < p.rp; c8 += 8) {
          float v8[8];
          ptx::tmem_ld8(lane_base + (uint32_t)(m * p.rp + c8), v8);
          ptx::tmem_ld_wait();
          if (code == m + 1) {
            float* dst = h_out + (size_t)t * p.r + c8;
            if (c8 + 8 <= p.r) {
              *reinterpret_cast<float4*>(dst) = make_float4(v8[0] * p.scale, v8[1] * p.scale, v8[2] * p.scale, v8[3] * p.scale);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(v8[4] * p.scale, v8[5] * p.scale, v8[6] * p.scale, v8[7] * p.scale);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (c8 + i < p.r) dst[i] = v8[i] * p.scale;
            }
          }
        }
      }
      if (code == 0)
        for (int j = 0; j < p.r; ++j) h_out[(size_t)t * p.r + j] = 0.f;
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) { dbg_stamp(p.dbg, blockIdx.x, 6); dbg_stamp_g(p.dbg, blockIdx.x, 101); }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

static int check_launch_tc() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

// Split of the d_in reduction of k_down_tc across two CTAs (when the token tiles alone do not fill the
// SMs).  k_up_tc must agree: with a split it reads two partial buffers.
static int down_ksplit(const DevDesc& d, int num_sms, bool allow_partials) {
  const int ntiles = (d.T + 127) / 128;
  return (allow_partials && ntiles < num_sms && d.d_in / 128 >= 8) ? 2 : 1;
}

static int launch_down_tc(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, bool allow_partials) {
  DownTcParams p;
  p.T = d.T; p.d_in = d.d_in; p.r = d.r; p.rp = round_up(d.r, 8); p.M = d.M;
  p.NP = round_up(p.M * p.rp, 16);
  if (p.NP > 256) p.NP = round_up(p.NP, 32);
  static int kbox_env = -1;
  if (kbox_env < 0) { const char* e = getenv("MOKA_DOWN_KBOX"); kbox_env = e ? atoi(e) : 2; }
  const int ntiles = (d.T + 127) / 128;
  p.ksplit = down_ksplit(d, num_sms, allow_partials && a.hpart != nullptr);
  p.w_bytes = (uint32_t)p.NP * 128u;
  // two 64-column boxes per stage (256 contiguous bytes per token row and request: measured faster
  // than one) as long as the ring still holds >= 4 stages
  static int down_budget_env = -1;                 // tuning aid: ring budget in KB (200: one CTA per SM, 100: two)
  if (down_budget_env < 0) { const char* e = getenv("MOKA_DOWN_BUDGET"); down_budget_env = e ? atoi(e) : 0; }
  const bool one_per_sm = down_budget_env ? down_budget_env > 110 : (ntiles * p.ksplit <= num_sms);
  const size_t budget0 = one_per_sm ? 200 * 1024 : 100 * 1024;
  // (two CTAs per SM share the SM's in-flight bytes: two stages each are enough there)
  const size_t min_stages = one_per_sm ? 4 : 2;
  p.kbox = (kbox_env == 2 && d.d_in % 128 == 0 && budget0 / ((16384u + p.w_bytes) * 2u) >= min_stages) ? 2 : 1;
  p.kblocks = d.d_in / (64 * p.kbox);
  p.rot = rot_enabled();
  p.early_bits = drop_early_bits(1);               // measured round 2: k_down_tc 64.9 -> 60.8 us (T = 32768, d = 4096)
  p.stage_bytes = (16384u + p.w_bytes) * (uint32_t)p.kbox;
  // one CTA per SM when the grid does not exceed the SM count: spend the whole shared memory on the
  // ring (tcgen05.commit frees a stage ~1.5k cycles after issue, so >= 6 stages keep HBM busy)
  const size_t budget = budget0;
  p.stages = (int)(budget / p.stage_bytes);
  if (p.stages > 10) p.stages = 10;
  if (p.stages < 2) p.stages = 2;
  if ((size_t)p.stages * p.stage_bytes + 1024 + 256 > 227 * 1024) return MOKA_EUNSUPPORTED;
  p.tmem_cols = tmem_cols_for(p.NP);
  p.scale = d.pre * d.drop_scale;
  p.dbg = dbg_enabled();
  CUtensorMap tx, ta[3];
  if (!make_tmap(&tx, a.x, (uint64_t)d.T, (uint64_t)d.d_in, 128)) return MOKA_EUNSUPPORTED;
  for (int m = 0; m < 3; ++m) {
    const void* base = a.A[m < d.M ? m : 0];
    if (!make_tmap(&ta[m], base, (uint64_t)d.r, (uint64_t)d.d_in, (uint32_t)d.r)) return MOKA_EUNSUPPORTED;
  }
  // split reduction: partials go to the forward workspace (hpart0 / hpart1), k_up_tc adds them and
  // writes the final h into the stash; unsplit: straight into the stash
  float* h0 = p.ksplit > 1 ? a.hpart : a.h_stash;
  float* h1 = p.ksplit > 1 ? a.hpart + (size_t)d.T * d.r : nullptr;
  cudaError_t e;
  // TS form (k_down_ts): MOKA_DOWN_TS = 0 off, 1 with dropout only, 2 always
  static int down_ts_env = -1;
  if (down_ts_env < 0) { const char* ev = getenv("MOKA_DOWN_TS"); down_ts_env = ev ? atoi(ev) : 1; }
  if (d.d_in % 128 == 0 && (down_ts_env == 2 || (down_ts_env == 1 && d.drop_thresh))) {
    DownTsParams q;
    q.T = p.T; q.d_in = p.d_in; q.r = p.r; q.rp = p.rp; q.M = p.M; q.NP = p.NP;
    q.ksplit = p.ksplit; q.kblocks = d.d_in / 128; q.rot = p.rot; q.drop = d.drop_thresh ? 1 : 0;
    q.w_bytes = p.w_bytes; q.scale = p.scale; q.dbg = p.dbg;
    q.a_col0 = round_up(p.NP, 64);
    const size_t wstage = 2u * (size_t)q.w_bytes;
    const size_t budget_ts = (one_per_sm ? 216 : 108) * 1024;   // ring bytes per CTA
    const int tmem_budget = one_per_sm ? 512 : 256;
    static int ts_xs_env = -1, ts_ws_env = -1;
    if (ts_xs_env < 0) { const char* ev = getenv("MOKA_DOWN_TS_XS"); ts_xs_env = ev ? atoi(ev) : 0; }
    if (ts_ws_env < 0) { const char* ev = getenv("MOKA_DOWN_TS_WS"); ts_ws_env = ev ? atoi(ev) : 0; }
    // x stages first (they hold the HBM stream), at least as many weight stages + 1 (a weight stage lives until the
    // products that read it have retired), A buffers from what TMEM is left
    q.xs = one_per_sm ? 4 : 2;
    if (ts_xs_env > 0) q.xs = ts_xs_env;
    while (q.xs > 1 && (size_t)q.xs * 32768 + (size_t)(q.xs + 1) * wstage > budget_ts) --q.xs;
    q.ws = (int)((budget_ts - (size_t)q.xs * 32768) / wstage);
    if (q.ws > q.xs + 2) q.ws = q.xs + 2;
    if (ts_ws_env > 0) q.ws = ts_ws_env;
    q.ab = (tmem_budget - q.a_col0) / 64;
    if (q.ab > q.ws) q.ab = q.ws;
    const size_t smem_ts = (size_t)q.xs * 32768 + (size_t)q.ws * wstage + 1024 + 512;
    if (q.ws >= 2 && q.ab >= 2 && smem_ts <= 227 * 1024 &&
        (size_t)q.xs * 32768 + (size_t)q.ws * wstage <= budget_ts + 1024) {
      q.tmem_cols = tmem_cols_for(q.a_col0 + q.ab * 64);
      e = cudaFuncSetAttribute(k_down_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ts);
      if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      e = launch_pdl(k_down_ts, dim3(ntiles * p.ksplit), dim3(kDownTsThreads), smem_ts, st, tx, ta[0], ta[1], ta[2], q, d,
                     a.tokmeta, h0, h1);
      if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      return check_launch_tc();
    }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + 256;
  if (d.drop_thresh) {
    e = cudaFuncSetAttribute(k_down_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    e = launch_pdl(k_down_tc<true>, dim3(ntiles * p.ksplit), dim3(kDownDropThreads), smem, st, tx, ta[0], ta[1], ta[2], p, d,
                   a.tokmeta, h0, h1);
  } else {
    e = cudaFuncSetAttribute(k_down_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    e = launch_pdl(k_down_tc<false>, dim3(ntiles * p.ksplit), dim3(kTcThreads), smem, st, tx, ta[0], ta[1], ta[2], p, d,
                   a.tokmeta, h0, h1);
  }
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  return check_launch_tc();
}

// ---------------------------------------------------------------------------- k_attn_fwd
// Rank-r cross-modal attention of the forward, between k_down_tc and k_up_tc.  Inside k_up_tc the
// attention sat on the critical path of every streaming CTA (measured: ~10 us of a ~36 us kernel
// during which only the 4-slot y0 ring was in flight, and twice per tile when the columns are split
// over two CTAs).  As its own small kernel it runs once per token on all SMs while k_up_tc - launched
// early by programmatic dependent launch - already prefetches its y0 tiles.
// A CTA owns 32 tokens of one sample: sums the partial h buffers of a split down-projection, writes
// the final h, g (= h for tokens that are not queries) and the per-token log-sum-exp.  One warp per
// query token, lane <-> key for the scores, lane <-> rank index for the weighted sum.
__global__ void __launch_bounds__(256)
k_attn_fwd(DevDesc d, int kcap, int gps, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
           const int32_t* __restrict__ klen, const float* __restrict__ h0, const float* __restrict__ h1,
           float* __restrict__ h_final, float* __restrict__ g_out, float* __restrict__ lse_out) {
  constexpr int TT = 32;
  extern __shared__ __align__(16) float sm_af[];
  const int r = d.r, ks = r + 1;
  float* Ks = sm_af;                                 // [kcap][r + 1]
  float* hs = Ks + (size_t)kcap * ks + ((kcap * ks) & 1 ? 1 : 0) + 2;
  hs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(hs) + 15) & ~uintptr_t(15));   // [TT][r]
  float* gs = hs + TT * r;                           // [TT][r]
  __shared__ int s_code[TT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / gps;
  const int s0 = (blockIdx.x % gps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const int L = min(klen[b], kcap);
  int myq = 0;
  if (tid < TT) {
    const int code = (tid < nt) ? (tokmeta[tbase + tid] & MOKA_MOD_MASK) : 0;
    s_code[tid] = code;
    myq = code >= MOKA_MOD_VIS;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  ptx::pdl_trigger();                                // (see k_attn_fwd_split)
  ptx::pdl_wait();                                   // h comes from k_down_tc
  const int r4 = r >> 2;                             // r % 4 == 0 on this path
  for (int i = tid; i < nt * r4; i += 256) {
    float4 v = reinterpret_cast<const float4*>(h0 + tbase * r)[i];
    if (h1) {
      const float4 w = reinterpret_cast<const float4*>(h1 + tbase * r)[i];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    reinterpret_cast<float4*>(hs)[i] = v;
    reinterpret_cast<float4*>(gs)[i] = v;
    if (h_final) reinterpret_cast<float4*>(h_final + tbase * r)[i] = v;
  }
  if (anyq) {
    for (int idx = tid; idx < L * r4; idx += 256) {
      const int k = idx / r4, j = (idx - k * r4) * 4;
      const int pos = kidx[(size_t)b * d.Lk + k];
      const size_t t = (size_t)b * d.S + pos;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tokmeta[t] & MOKA_KEY_VALID) {
        v = *reinterpret_cast<const float4*>(h0 + t * r + j);
        if (h1) {
          const float4 w = *reinterpret_cast<const float4*>(h1 + t * r + j);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
      float* dst = Ks + k * ks + j;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  }
  __syncthreads();
  if (anyq) {
    AttnCtx c{tokmeta, kidx, klen, h0, nullptr, Ks, nullptr, ks, kcap, nullptr};
    for (int ti = 0; ti < TT / 8; ++ti) {
      const int tt = warp * (TT / 8) + ti;
      if (tt >= nt || s_code[tt] < MOKA_MOD_VIS) continue;       // warp-uniform
      attn_token_fast<false>(d, c, L, hs + tt * r, gs + tt * r, nullptr, lane, lse_out + tbase + tt);
    }
  }
  __syncthreads();
  for (int i = tid; i < nt * r4; i += 256) reinterpret_cast<float4*>(g_out + tbase * r)[i] = reinterpret_cast<float4*>(gs)[i];
}

// k_attn_fwd_split: the fast form for r <= 32.  Eight threads share one query token: thread `sub`
// takes the keys k = sub (mod 8) with an online softmax over its share (two keys per step, h and
// the running sum in registers, key rows read as 128-bit shared loads: row stride r + 4 floats puts
// the eight rows a warp touches per load on disjoint banks), then the eight partial (max, sum,
// weighted-sum) triples are merged with three xor-shuffle rounds.  32 tokens per CTA, 256 threads.
template <int R>
__global__ void __launch_bounds__(256)
k_attn_fwd_split(DevDesc d, int kcap, int gps, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
                 const int32_t* __restrict__ klen, const float* __restrict__ h0, const float* __restrict__ h1,
                 float* __restrict__ h_final, float* __restrict__ g_out, float* __restrict__ lse_out) {
  constexpr int TT = 32, SUB = 8, KS = R + 4;
  extern __shared__ __align__(16) float sm_af[];
  const int r = d.r;
  float* Ks = sm_af;                                 // [kcap][KS]
  float* hs = Ks + (size_t)kcap * KS;                // [TT][R]  (h in, g out)
  __shared__ int s_code[TT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / gps;
  const int s0 = (blockIdx.x % gps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const int L = min(klen[b], kcap);
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 102);
  int myq = 0;
  if (tid < TT) {
    const int code = (tid < nt) ? (tokmeta[tbase + tid] & MOKA_MOD_MASK) : 0;
    s_code[tid] = code;
    myq = code >= MOKA_MOD_VIS;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  // key positions / validity do not depend on the previous kernel: look them up before the wait
  constexpr int R4 = R / 4;
  const int r4 = r >> 2;
  int kpos[2] = {-1, -1};
  if (anyq) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + u * 256;
      const int k = idx / R4;
      if (k < L) {
        const int pos = kidx[(size_t)b * d.Lk + k];
        if (tokmeta[(size_t)b * d.S + pos] & MOKA_KEY_VALID) kpos[u] = pos;
      }
    }
  }
  // k_up_tc may start right away: before its own wait it only touches tensors older than this call
  // (y0 tiles, B, token metadata) - and k_down_tc, which triggered after ITS wait, vouches for those
  ptx::pdl_trigger();
  ptx::pdl_wait();                                   // h comes from k_down_tc
  for (int i = tid; i < TT * R4; i += 256) {
    const int tt = i / R4, j4 = i - tt * R4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt < nt && j4 < r4) {
      const size_t off = (tbase + tt) * r + j4 * 4;
      v = *reinterpret_cast<const float4*>(h0 + off);
      if (h1) {
        const float4 w = *reinterpret_cast<const float4*>(h1 + off);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      if (h_final) *reinterpret_cast<float4*>(h_final + off) = v;
    }
    reinterpret_cast<float4*>(hs)[i] = v;
  }
  if (anyq) {
    // kcap * R4 <= 512 items is guaranteed by the launcher for the first two rounds; the rest loops
    for (int u = 0, idx = tid; idx < L * R4; idx += 256, ++u) {
      const int k = idx / R4, j4 = idx - k * R4;
      int pos;
      if (u < 2) pos = kpos[u];
      else {
        pos = kidx[(size_t)b * d.Lk + k];
        if (!(tokmeta[(size_t)b * d.S + pos] & MOKA_KEY_VALID)) pos = -1;
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pos >= 0 && j4 < r4) {
        const size_t off = ((size_t)b * d.S + pos) * r + j4 * 4;
        v = *reinterpret_cast<const float4*>(h0 + off);
        if (h1) {
          const float4 w = *reinterpret_cast<const float4*>(h1 + off);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
      *reinterpret_cast<float4*>(Ks + k * KS + j4 * 4) = v;
    }
  }
  __syncthreads();
  if (anyq) {
    const int tt = warp * 4 + (lane >> 3), sub = lane & 7;
    const bool isq = tt < nt && s_code[tt] >= MOKA_MOD_VIS;
    const int Lq = isq ? L : 0;
    float h[R], acc[R];
#pragma unroll
    for (int j = 0; j < R; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(hs + tt * R + j);
      h[j] = v.x; h[j + 1] = v.y; h[j + 2] = v.z; h[j + 3] = v.w;
      acc[j] = acc[j + 1] = acc[j + 2] = acc[j + 3] = 0.f;
    }
    const float sl = d.sm_scale * kLog2e;
    float m = -INFINITY, l = 0.f;
    const uint32_t ks_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
    for (int k0 = sub; k0 < Lq; k0 += 2 * SUB) {
      const int k1 = k0 + SUB;
      const bool has1 = k1 < Lq;
      float kv0[R], kv1[R];
#pragma unroll
      for (int j = 0; j < R; j += 4) {
        float4 a, c2;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(ks_addr + (uint32_t)(k0 * KS + j) * 4u));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(c2.x), "=f"(c2.y), "=f"(c2.z), "=f"(c2.w)
                     : "r"(ks_addr + (uint32_t)((has1 ? k1 : k0) * KS + j) * 4u));
        kv0[j] = a.x; kv0[j + 1] = a.y; kv0[j + 2] = a.z; kv0[j + 3] = a.w;
        kv1[j] = c2.x; kv1[j + 1] = c2.y; kv1[j + 2] = c2.z; kv1[j + 3] = c2.w;
      }
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
      for (int j = 0; j < R; j += 2) {
        a0 = fmaf(h[j], kv0[j], a0); a1 = fmaf(h[j + 1], kv0[j + 1], a1);
        b0 = fmaf(h[j], kv1[j], b0); b1 = fmaf(h[j + 1], kv1[j + 1], b1);
      }
      const float sc0 = (a0 + a1) * sl;
      const float sc1 = has1 ? (b0 + b1) * sl : -INFINITY;
      const float smax = fmaxf(sc0, sc1);
      if (smax > m) {
        const float corr = exp2f(m - smax);
        l *= corr;
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] *= corr;
        m = smax;
      }
      const float e0 = exp2f(sc0 - m), e1 = exp2f(sc1 - m);
      l += e0 + e1;
#pragma unroll
      for (int j = 0; j < R; ++j) acc[j] = fmaf(e0, kv0[j], fmaf(e1, kv1[j], acc[j]));
    }
    // merge the eight partials of a token (lanes differing in bits 0..2)
#pragma unroll
    for (int bit = 1; bit < SUB; bit <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, bit);
      const float lo = __shfl_xor_sync(0xffffffffu, l, bit);
      const float mn = fmaxf(m, mo);
      const float ca = (m == -INFINITY) ? 0.f : exp2f(m - mn);
      const float cb = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
      l = l * ca + lo * cb;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[j], bit);
        acc[j] = acc[j] * ca + ao * cb;
      }
      m = mn;
    }
    if (isq && sub == 0) {
      const float sc = d.attn_w / l;
      lse_out[tbase + tt] = m + log2f(l);
#pragma unroll
      for (int j = 0; j < R; ++j) hs[tt * R + j] = fmaf(sc, acc[j], h[j]);
    }
  }
  __syncthreads();
  for (int i = tid; i < TT * R4; i += 256) {
    const int tt = i / R4, j4 = i - tt * R4;
    if (tt < nt && j4 < r4)
      *reinterpret_cast<float4*>(g_out + (tbase + tt) * r + j4 * 4) = reinterpret_cast<const float4*>(hs)[i];
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 103);
}

// k_attn_fwd_rank: forward attention for 32 < r <= 128.  One warp per query token, lane <-> C = R/32
// consecutive rank indices (h and the running weighted sum in registers, key rows read as aligned
// vectors), the score of a key is a warp reduction; two keys per step.  Same tile / staging scheme as
// k_attn_fwd_split.
template <int C>
__global__ void __launch_bounds__(256)
k_attn_fwd_rank(DevDesc d, int kcap, int gps, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
                const int32_t* __restrict__ klen, const float* __restrict__ h0, const float* __restrict__ h1,
                float* __restrict__ h_final, float* __restrict__ g_out, float* __restrict__ lse_out) {
  constexpr int TT = 32, R = 32 * C, R4 = R / 4;
  extern __shared__ __align__(16) float sm_af[];
  const int r = d.r;
  const int kcap2 = (kcap + 1) & ~1;                 // rows are read in pairs
  float* Ks = sm_af;                                 // [kcap2][R]
  float* hs = Ks + (size_t)kcap2 * R;                // [TT][R]  (h in, g out)
  __shared__ int s_code[TT];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / gps;
  const int s0 = (blockIdx.x % gps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const int L = min(klen[b], kcap);
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 102);
  int myq = 0;
  if (tid < TT) {
    const int code = (tid < nt) ? (tokmeta[tbase + tid] & MOKA_MOD_MASK) : 0;
    s_code[tid] = code;
    myq = code >= MOKA_MOD_VIS;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  ptx::pdl_trigger();                                // (see k_attn_fwd_split)
  ptx::pdl_wait();                                   // h comes from k_down_tc
  const int r4 = r >> 2;
  for (int i = tid; i < TT * R4; i += 256) {
    const int tt = i / R4, j4 = i - tt * R4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt < nt && j4 < r4) {
      const size_t off = (tbase + tt) * r + j4 * 4;
      v = *reinterpret_cast<const float4*>(h0 + off);
      if (h1) {
        const float4 w = *reinterpret_cast<const float4*>(h1 + off);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      if (h_final) *reinterpret_cast<float4*>(h_final + off) = v;
    }
    reinterpret_cast<float4*>(hs)[i] = v;
  }
  if (anyq) {
    for (int idx = tid; idx < kcap2 * R4; idx += 256) {
      const int k = idx / R4, j4 = idx - k * R4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < L && j4 < r4) {
        const int pos = kidx[(size_t)b * d.Lk + k];
        const size_t t = (size_t)b * d.S + pos;
        if (tokmeta[t] & MOKA_KEY_VALID) {
          v = *reinterpret_cast<const float4*>(h0 + t * r + j4 * 4);
          if (h1) {
            const float4 w = *reinterpret_cast<const float4*>(h1 + t * r + j4 * 4);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
          }
        }
      }
      reinterpret_cast<float4*>(Ks)[idx] = v;
    }
  }
  __syncthreads();
  if (anyq) {
    const float sl = d.sm_scale * kLog2e;
    for (int ti = 0; ti < TT / 8; ++ti) {
      const int tt = warp * (TT / 8) + ti;
      if (tt >= nt || s_code[tt] < MOKA_MOD_VIS) continue;       // warp-uniform
      float h[C], acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { h[c] = hs[tt * R + lane * C + c]; acc[c] = 0.f; }
      float m = -INFINITY, l = 0.f;
      for (int k0 = 0; k0 < L; k0 += 2) {
        float K0[C], K1[C];
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          K0[c] = Ks[k0 * R + lane * C + c];
          K1[c] = Ks[(k0 + 1) * R + lane * C + c];
          a0 = fmaf(h[c], K0[c], a0);
          a1 = fmaf(h[c], K1[c], a1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          a0 += __shfl_xor_sync(0xffffffffu, a0, o);
          a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        const float sc0 = a0 * sl;
        const float sc1 = (k0 + 1 < L) ? a1 * sl : -INFINITY;
        const float smax = fmaxf(sc0, sc1);
        if (smax > m) {                                // uniform
          const float corr = exp2f(m - smax);
          l *= corr;
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] *= corr;
          m = smax;
        }
        const float e0 = exp2f(sc0 - m), e1 = exp2f(sc1 - m);
        l += e0 + e1;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(e0, K0[c], fmaf(e1, K1[c], acc[c]));
      }
      const float scw = d.attn_w / l;
      if (lane == 0) lse_out[tbase + tt] = m + log2f(l);
#pragma unroll
      for (int c = 0; c < C; ++c) hs[tt * R + lane * C + c] = fmaf(scw, acc[c], h[c]);
    }
  }
  __syncthreads();
  for (int i = tid; i < TT * R4; i += 256) {
    const int tt = i / R4, j4 = i - tt * R4;
    if (tt < nt && j4 < r4)
      *reinterpret_cast<float4*>(g_out + (tbase + tt) * r + j4 * 4) = reinterpret_cast<const float4*>(hs)[i];
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 103);
}

// ---------------------------------------------------------------------------- k_attn_fwd_mma
// Forward attention on the (warp-level) tensor cores: mma.sync.m16n8k16 bf16 with fp32 accumulation, every
// fp32 operand split into hi + lo bf16 halves and three products per tile (hi*hi + lo*hi + hi*lo), which keeps
// the scores and the weighted sums at fp32-level accuracy.  A warp owns 16 tokens (one m-tile) and ALL keys of
// their sample, so the softmax is warp-local (quad shuffles), as in flash attention:
//   S[16, LP]  = H[16, R] K^T        (A = H rows, B = K rows, k = rank)
//   O[16, R]   = P[16, LP] K         (A = P straight from the S accumulator registers, B = K^T rows, k = key)
// K is staged once per CTA as bf16 hi/lo in both orientations; 64 tokens per CTA (4 warps).  This replaces the
// SIMT forms above whenever r % 8 == 0: ~5x fewer instructions at r = 16, ~50x at r = 64.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 xh = __float2bfloat16_rn(x), yh = __float2bfloat16_rn(y);
  hi = pack_bf16(__bfloat162float(xh), __bfloat162float(yh));
  lo = pack_bf16(x - __bfloat162float(xh), y - __bfloat162float(yh));
}

template <int R, int LP>      // R: rank padded to 16; LP: keys padded to 16 (<= 128)
__global__ void __launch_bounds__(256)
k_attn_fwd_mma(DevDesc d, int gps, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
               const int32_t* __restrict__ klen, const float* __restrict__ h0, const float* __restrict__ h1,
               float* __restrict__ h_final, float* __restrict__ g_out, float* __restrict__ lse_out) {
  constexpr int TT = 64, HS = R + 4;                 // fp32 row stride of the token rows
  constexpr int KS = R + 8, TS = LP + 8;             // bf16 row strides: K [LP][KS], K^T [R][TS] (conflict-free fragment loads)
  extern __shared__ __align__(16) uint8_t sm_mma[];
  float* hs = reinterpret_cast<float*>(sm_mma);                         // [TT][HS]
  __nv_bfloat16* Khi = reinterpret_cast<__nv_bfloat16*>(hs + TT * HS);  // [LP][KS]
  __nv_bfloat16* Klo = Khi + LP * KS;
  __nv_bfloat16* Thi = Klo + LP * KS;                                   // [R][TS]
  __nv_bfloat16* Tlo = Thi + R * TS;
  __shared__ int s_code[TT];
  __shared__ int s_kpos[LP];                         // in-sample position of key k, -1 = zero row
  constexpr int NTH = 256;                           // 4 compute warps (one m-tile each) + 4 that only help staging
  const int r = d.r;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / gps;
  const int s0 = (blockIdx.x % gps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const int L = min(klen[b], LP);
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 102);
  int myq = 0;
  if (tid < TT) {
    const int code = (tid < nt) ? (tokmeta[tbase + tid] & MOKA_MOD_MASK) : 0;
    s_code[tid] = code;
    myq = code >= MOKA_MOD_VIS;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  ptx::pdl_trigger();                                // (see k_attn_fwd_split)
  // key positions / validity are older than this call: resolve the two dependent lookups before the wait
  if (anyq && tid < LP) {
    int pos = -1;
    if (tid < L) {
      pos = kidx[(size_t)b * d.Lk + tid];
      if (!(tokmeta[(size_t)b * d.S + pos] & MOKA_KEY_VALID)) pos = -1;
    }
    s_kpos[tid] = pos;
  }
  ptx::pdl_wait();                                   // h comes from k_down_tc
  __syncthreads();
  constexpr int R4 = R / 4;
  const int r4 = r >> 2;
  for (int i = tid; i < TT * R4; i += NTH) {
    const int tt = i / R4, j4 = i - tt * R4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt < nt && j4 < r4) {
      const size_t off = (tbase + tt) * r + j4 * 4;
      v = *reinterpret_cast<const float4*>(h0 + off);
      if (h1) {
        const float4 w = *reinterpret_cast<const float4*>(h1 + off);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      if (h_final) *reinterpret_cast<float4*>(h_final + off) = v;
      if (!anyq) *reinterpret_cast<float4*>(g_out + off) = v;          // no query in this tile: g == h
    }
    *reinterpret_cast<float4*>(hs + tt * HS + j4 * 4) = v;
  }
  if (!anyq) { if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 103); return; }
  // keys -> bf16 hi / lo, both orientations (rows beyond L and invalid keys are zero); four loads in flight
  // per thread before the first use
  for (int base = 0; base < LP * R4; base += 4 * NTH) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + tid + u * NTH;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < LP * R4) {
        const int k = idx / R4, j = (idx - k * R4) * 4;
        const int pos = s_kpos[k];
        if (pos >= 0 && j < r) {
          const size_t t = (size_t)b * d.S + pos;
          v[u] = *reinterpret_cast<const float4*>(h0 + t * r + j);
          if (h1) {
            const float4 w = *reinterpret_cast<const float4*>(h1 + t * r + j);
            v[u].x += w.x; v[u].y += w.y; v[u].z += w.z; v[u].w += w.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = base + tid + u * NTH;
      if (idx >= LP * R4) continue;
      const int k = idx / R4, j = (idx - k * R4) * 4;
      const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(f[e]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(f[e] - __bfloat162float(hi));
        Khi[k * KS + j + e] = hi; Klo[k * KS + j + e] = lo;
        Thi[(j + e) * TS + k] = hi; Tlo[(j + e) * TS + k] = lo;
      }
    }
  }
  __syncthreads();
  // ---- one m-tile (16 tokens) per compute warp
  const int m0 = (warp & 3) * 16;
  bool any = false;
  if (warp < 4)
    for (int i = 0; i < 16; ++i) any |= (s_code[m0 + i] >= MOKA_MOD_VIS);
  const int g = lane >> 2, t = lane & 3;
  if (any) {
    constexpr int NT = LP / 8;
    float S[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { S[n][0] = S[n][1] = S[n][2] = S[n][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < R / 16; ++ks) {
      uint32_t ahi[4], alo[4];
      {
        const float2 x0 = *reinterpret_cast<const float2*>(hs + (m0 + g) * HS + ks * 16 + 2 * t);
        const float2 x1 = *reinterpret_cast<const float2*>(hs + (m0 + g + 8) * HS + ks * 16 + 2 * t);
        const float2 x2 = *reinterpret_cast<const float2*>(hs + (m0 + g) * HS + ks * 16 + 2 * t + 8);
        const float2 x3 = *reinterpret_cast<const float2*>(hs + (m0 + g + 8) * HS + ks * 16 + 2 * t + 8);
        split_bf16x2(x0.x, x0.y, ahi[0], alo[0]); split_bf16x2(x1.x, x1.y, ahi[1], alo[1]);
        split_bf16x2(x2.x, x2.y, ahi[2], alo[2]); split_bf16x2(x3.x, x3.y, ahi[3], alo[3]);
      }
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const int krow = n * 8 + g;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(Khi + krow * KS + ks * 16 + 2 * t);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(Khi + krow * KS + ks * 16 + 2 * t + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(Klo + krow * KS + ks * 16 + 2 * t);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(Klo + krow * KS + ks * 16 + 2 * t + 8);
        mma_bf16_16816(S[n], ahi, bh0, bh1);
        mma_bf16_16816(S[n], alo, bh0, bh1);
        mma_bf16_16816(S[n], ahi, bl0, bl1);
      }
    }
    // softmax over the L keys of rows g (c0, c1) and g + 8 (c2, c3); key index of c0 / c2 = n*8 + 2t, c1 / c3 = +1
    const float sl = d.sm_scale * kLog2e;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = n * 8 + 2 * t + e < L;
        S[n][e] = ok ? S[n][e] * sl : -INFINITY;
        S[n][2 + e] = ok ? S[n][2 + e] * sl : -INFINITY;
        mx0 = fmaxf(mx0, S[n][e]); mx1 = fmaxf(mx1, S[n][2 + e]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        S[n][e] = exp2f(S[n][e] - mx0); l0 += S[n][e];
        S[n][2 + e] = exp2f(S[n][2 + e] - mx1); l1 += S[n][2 + e];
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P K : the S accumulators of two adjacent n-tiles are exactly one A fragment (16 keys) of P
    constexpr int NO = R / 8;
    float O[NO][4];
#pragma unroll
    for (int n = 0; n < NO; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk) {
      if (kk * 16 >= L) break;                       // uniform
      uint32_t phi[4], plo[4];
      split_bf16x2(S[2 * kk][0], S[2 * kk][1], phi[0], plo[0]);
      split_bf16x2(S[2 * kk][2], S[2 * kk][3], phi[1], plo[1]);
      split_bf16x2(S[2 * kk + 1][0], S[2 * kk + 1][1], phi[2], plo[2]);
      split_bf16x2(S[2 * kk + 1][2], S[2 * kk + 1][3], phi[3], plo[3]);
#pragma unroll
      for (int n = 0; n < NO; ++n) {
        const int rrow = n * 8 + g;                  // rank index = n of the B operand
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(Thi + rrow * TS + kk * 16 + 2 * t);
        const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(Thi + rrow * TS + kk * 16 + 2 * t + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(Tlo + rrow * TS + kk * 16 + 2 * t);
        const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(Tlo + rrow * TS + kk * 16 + 2 * t + 8);
        mma_bf16_16816(O[n], phi, bh0, bh1);
        mma_bf16_16816(O[n], plo, bh0, bh1);
        mma_bf16_16816(O[n], phi, bl0, bl1);
      }
    }
    // g = h + w / l * O for query rows (others keep h); O[n]: rows g / g+8, rank columns n*8 + 2t, +1
    const int row0 = m0 + g, row1 = m0 + g + 8;
    const bool q0 = s_code[row0] >= MOKA_MOD_VIS, q1 = s_code[row1] >= MOKA_MOD_VIS;
    const float sc0 = q0 ? d.attn_w / l0 : 0.f, sc1 = q1 ? d.attn_w / l1 : 0.f;
    if (t == 0) {
      if (q0) lse_out[tbase + row0] = mx0 + log2f(l0);
      if (q1) lse_out[tbase + row1] = mx1 + log2f(l1);
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < NO; ++n) {
      const int c = n * 8 + 2 * t;
      float2* p0 = reinterpret_cast<float2*>(hs + row0 * HS + c);
      float2* p1 = reinterpret_cast<float2*>(hs + row1 * HS + c);
      float2 v0 = *p0, v1 = *p1;
      v0.x = fmaf(sc0, O[n][0], v0.x); v0.y = fmaf(sc0, O[n][1], v0.y);
      v1.x = fmaf(sc1, O[n][2], v1.x); v1.y = fmaf(sc1, O[n][3], v1.y);
      *p0 = v0; *p1 = v1;
    }
  }
  __syncthreads();
  for (int i = tid; i < TT * R4; i += NTH) {
    const int tt = i / R4, j4 = i - tt * R4;
    if (tt < nt && j4 < r4)
      *reinterpret_cast<float4*>(g_out + (tbase + tt) * r + j4 * 4) = *reinterpret_cast<const float4*>(hs + tt * HS + j4 * 4);
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 103);
}

// ---------------------------------------------------------------------------- k_up_tc
//   Y[128 tok, 128 cols] = Y0 + (post_m g)[128 tok, r] . B[128 cols, r]^T   per column chunk
// Warp 0 streams the y0 tiles in with TMA, the 128 worker threads (warps 2..5) run the rank-r
// cross-modal attention for the token tile while those loads are in flight, stage the bf16
// operand tiles (manual 128-byte swizzle), and per chunk add the TMEM accumulator onto the
// y0 tile in shared memory, which one thread TMA-stores back.  Warp 1 issues the MMAs.
struct UpTcParams {
  int T, d_out, r, rk;             // rk = r rounded up to 16 (UMMA K step)
  int nsplit, chunks;              // chunks = d_out / 128, split across nsplit CTAs per token tile
  int ystages, kcap;
  int write_g;
  int attn_in_kernel;              // 0: g rows come precomputed from k_attn_fwd (h_in0 = g)
  int bpg;                         // B-tile buffers / TMEM accumulators per worker group: 2, or 1 when r > 64
  int delta;                       // MOKA_FLAG_DELTA_OUT: y receives the adapter term only (no y0 read)
  int rot;                         // rotate the chunk order per CTA (see rot_enabled)
  int streamk;                     // 1: persistent grid, CTA c owns the units [c U / G, (c+1) U / G) of the (token tile,
                                   //    128-column chunk) list (see the stream-K note); 0: one CTA per (tile, column split)
  int grid;                        // G
  long long units;                 // U = ntiles * chunks
  int dbg;
};


__device__ __forceinline__ void named_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory");
}

// Epilogue of the read-modify-write kernels: add the fp32 TMEM accumulator [128 lanes x 128 cols]
// onto a bf16 tile held in shared memory as two {64 cols x 128 rows} 128B-swizzled TMA boxes.
// One thread owns one row (TMEM lane).  All shared loads of a half-row are issued before the
// arithmetic and all stores after it, so nothing waits on a load -> store -> load chain.
// DROP (k_wgrad_tc with lora_dropout > 0): the accumulator is the gradient with respect to the DROPPED input,
// so it is masked with the row's keep bits (km16: 16 bytes, one per 8-column chunk of the 128-column tile,
// produced when the x tile of the same rows was masked) before it is added.
// DELTA (MOKA_FLAG_DELTA_OUT): the tile is write-only - the slot holds no data, the accumulator is rounded to
// bf16 and stored; the frozen GEMM adds its product onto it afterwards (beta = 1).
template <bool DROP = false, bool DELTA = false>
__device__ __forceinline__ void add_acc_to_tile(uint32_t tmem_row_addr, uint32_t tile_smem, uint32_t row,
                                                uint64_t* tmem_free = nullptr, const uint32_t* km16 = nullptr) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t box = tile_smem + (uint32_t)half * 16384u;
    uint4 yv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      yv[c] = DELTA ? make_uint4(0u, 0u, 0u, 0u) : ptx::lds128(box + ptx::sw128_offset(row, (uint32_t)c));
    float acc[2][32];
    ptx::tmem_ld32(tmem_row_addr + (uint32_t)(half * 64), acc[0]);
    ptx::tmem_ld32(tmem_row_addr + (uint32_t)(half * 64 + 32), acc[1]);
    ptx::tmem_ld_wait();
    if (half == 1 && tmem_free) {              // accumulator fully read: hand the TMEM buffer back early
      ptx::tc_fence_before_sync();
      warp_arrive(tmem_free);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float* a = &acc[c >> 2][(c & 3) * 8];
      uint4 v = yv[c];
      if (DROP) {
        // (a test + select per element - two alu-pipe instructions - costs ~3 us of a 300 us backward call inside the
        // step; folding the keep bit into the add as an fp32 factor, (km & 2^i) * (0x3F800000 >> i) and an FMA, moves
        // the work to the fma pipe and was measured 27 us SLOWER: profiles/r02_dropout_cost.md)
        const uint32_t km = (km16[half * 2 + (c >> 2)] >> ((c & 3) * 8)) & 0xffu;   // keep bits of chunk (half, c)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (!((km >> i) & 1u)) a[i] = 0.f;
      }
      v.x = pack_bf16(bf16lo(v.x) + a[0], bf16hi(v.x) + a[1]);
      v.y = pack_bf16(bf16lo(v.y) + a[2], bf16hi(v.y) + a[3]);
      v.z = pack_bf16(bf16lo(v.z) + a[4], bf16hi(v.z) + a[5]);
      v.w = pack_bf16(bf16lo(v.w) + a[6], bf16hi(v.w) + a[7]);
      ptx::sts128(box + ptx::sw128_offset(row, (uint32_t)c), v);
    }
  }
}

constexpr int kUpThreads = 352;        // warps 2..9: two worker groups (even / odd chunks); warp 10: TMA store issuer

__global__ void __launch_bounds__(kUpThreads)
k_up_tc(const __grid_constant__ CUtensorMap tm_y0, const __grid_constant__ CUtensorMap tm_y, DevDesc d,
        UpTcParams p, const __nv_bfloat16* __restrict__ Bw, const uint8_t* __restrict__ tokmeta,
        const int32_t* __restrict__ kidx, const int32_t* __restrict__ klen, const float* __restrict__ h_in0,
        const float* __restrict__ h_in1, float* __restrict__ h_final, float* __restrict__ g_out) {
  // h_in0 (+ h_in1 when k_down_tc split its reduction) = pre-attention rows; h_final: where the summed
  // rows are written for the backward (null when h_in0 already is the stash)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kblocks = (p.rk + 63) / 64;                       // 64-wide K blocks of the operand tiles
  uint8_t* ytiles = base;                                     // [ystages][2 boxes][128 x 128 B]
  uint8_t* gtile = ytiles + (size_t)p.ystages * 32768;        // [kblocks][128 x 128 B]  A operand
  uint8_t* btiles = gtile + (size_t)kblocks * 16384;          // [4][kblocks][128 x 128 B] B operand
  float* gs = reinterpret_cast<float*>(btiles + (size_t)2 * p.bpg * kblocks * 16384);   // [128, r] (in-kernel attention only)
  float* Ks = gs + (p.attn_in_kernel ? 128 * d.r : 0);        // [kcap, r+1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ks + (size_t)p.kcap * (d.r + 1) + ((p.kcap * (d.r + 1)) & 1));
  uint64_t* y_full = bars;                                    // [ystages]
  uint64_t* y_empty = y_full + p.ystages;                     // [ystages]
  uint64_t* b_full = y_empty + p.ystages;                     // [4]
  uint64_t* t_full = b_full + 4;                              // [4]
  uint64_t* t_empty = t_full + 4;                             // [4]
  uint64_t* g_full = t_empty + 4;                             // [1] A operand (post*g) staged
  uint64_t* y_done = g_full + 1;                              // [ystages] tile updated in smem, ready to store
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_done + p.ystages);
  int* or_slot = reinterpret_cast<int*>(tmem_slot + 1);
  uint8_t* meta_s = reinterpret_cast<uint8_t*>(or_slot + 1);  // [128]

  const int warp = ptx::uniform_warp_idx(), lane = threadIdx.x & 31;
  // work of this CTA: `nch` (tile, chunk) units.  Stream-K: a contiguous range of the tile-major unit list, possibly
  // spanning several token tiles (segments); otherwise the column split `split` of one tile, in a rotated order.
  long long u_beg = 0;
  int tile, split, c_beg, nch, crot;
  if (p.streamk) {
    u_beg = (long long)blockIdx.x * p.units / p.grid;
    nch = (int)((long long)(blockIdx.x + 1) * p.units / p.grid - u_beg);
    tile = (int)(u_beg / p.chunks); split = 0; c_beg = 0; crot = 0;
  } else {
    tile = blockIdx.x / p.nsplit; split = blockIdx.x % p.nsplit;
    c_beg = (int)((long long)split * p.chunks / p.nsplit);
    nch = (int)((long long)(split + 1) * p.chunks / p.nsplit) - c_beg;
    crot = p.rot ? (int)(((unsigned)tile * 7u + (unsigned)split * 3u) % (unsigned)nch) : 0;
  }
  const int t0 = tile * 128;                                  // first token tile of this CTA
  const int nt = min(128, p.T - t0);
  auto unit_of = [&](int j, int& tl, int& ch) {               // j-th unit this CTA processes -> (token tile, column chunk)
    if (p.streamk) {
      const long long u = u_beg + j;
      tl = (int)(u / p.chunks);
      ch = (int)(u - (long long)tl * p.chunks);
    } else {
      tl = tile;
      ch = c_beg + rot_index(j, crot, nch);
    }
  };
  auto chunk_of = [&](int j) { int tl, ch; unit_of(j, tl, ch); return ch; };
  if (threadIdx.x == 0) dbg_stamp_g(p.dbg, blockIdx.x, 104);
  // last kernel of the forward chain: whatever follows (the first kernel of the next libmoka call waits
  // before it touches memory; anything else is a normal stream-ordered launch) may start its prologue now
  if (threadIdx.x == 0) ptx::pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.ystages; ++s) {
      ptx::mbar_init(&y_full[s], 1); ptx::mbar_init(&y_empty[s], 1); ptx::mbar_init(&y_done[s], kWorkerWarps);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&b_full[i], kWorkerWarps); ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], kWorkerWarps);
    }
    ptx::mbar_init(g_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_y0); ptx::prefetch_tmap(&tm_y);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: y0 tiles, as far ahead as the ring allows (whole warp, one elected lane issues)
    {
      const uint32_t yt_a = ptx::uniform_u32(ptx::smem_u32(ytiles));
      for (int j = 0; j < nch; ++j) {
        const int s = j % p.ystages;
        const uint32_t ph = (uint32_t)(j / p.ystages) & 1u;
        ptx::mbar_wait_or_trap(&y_empty[s], ph ^ 1u, __LINE__);
        const uint32_t yt = yt_a + (uint32_t)s * 32768u;
        if (p.delta) {                                           // write-only output: the free slot is all the workers need
          ptx::mbar_arrive_w(&y_full[s]);
          if (lane == 0) ptx::prog_mark(0, j + 1);
          continue;
        }
        ptx::mbar_arrive_expect_tx_w(&y_full[s], 32768u);
        int tl, ch;
        unit_of(j, tl, ch);
        const int col = ch * 128;
        ptx::tma_load_2d_w(yt, &tm_y0, &y_full[s], col, tl * 128, ptx::kEvictFirst);
        ptx::tma_load_2d_w(yt + 16384u, &tm_y0, &y_full[s], col + 64, tl * 128, ptx::kEvictFirst);
        if (lane == 0) ptx::prog_mark(0, j + 1);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t tb = ptx::uniform_u32(tmem_base);
      const uint32_t idesc = ptx::idesc_bf16(128, 128, 0, 0);
      const uint32_t ga = ptx::uniform_u32(ptx::smem_u32(gtile));
      const uint32_t bt_a = ptx::uniform_u32(ptx::smem_u32(btiles));
      const uint64_t dbase = ptx::smem_desc_sw128(0, 1024, 0);
      ptx::mbar_wait_or_trap(g_full, 0, __LINE__);
      int cur_tile = tile;
      uint32_t g_ph = 0;
      for (int j = 0; j < nch; ++j) {
        if (p.streamk) {
          int tl, ch;
          unit_of(j, tl, ch);
          if (tl != cur_tile) {                                  // next segment: its A operand (post * g of that tile)
            cur_tile = tl; g_ph ^= 1u;
            ptx::mbar_wait_or_trap(g_full, g_ph, __LINE__);
          }
        }
        // chunk j belongs to worker group (j & 1); each group rotates over two accumulator / B buffers
        const int i = (j & 1) * 2 + ((j >> 1) & (p.bpg - 1));
        const uint32_t ph = (uint32_t)((j >> 1) / p.bpg) & 1u;
        ptx::mbar_wait_or_trap(&b_full[i], ph, __LINE__);
        if (lane == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 32 + j * 3);
        ptx::mbar_wait_or_trap(&t_empty[i], ph ^ 1u, __LINE__);
        if (lane == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 33 + j * 3);
        ptx::tc_fence_after_sync();
        const uint32_t ba = bt_a + (uint32_t)((j & 1) * p.bpg + ((j >> 1) & (p.bpg - 1))) * (uint32_t)kblocks * 16384u;
        int kdone = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          const int ksteps = min(4, (p.rk - kb * 64) / 16);
          for (int k = 0; k < ksteps; ++k, ++kdone) {
            const uint64_t adesc = dbase + (uint64_t)((ga + kb * 16384 + k * 32) >> 4);
            const uint64_t bdesc = dbase + (uint64_t)((ba + kb * 16384 + k * 32) >> 4);
            ptx::umma_bf16_w(tb + (uint32_t)i * 128u, adesc, bdesc, idesc, kdone > 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit_w(&t_full[i]);
        if (lane == 0) { ptx::prog_mark(1, j + 1); if (j < 6) dbg_stamp(p.dbg, blockIdx.x, 34 + j * 3); }
      }
    }
  } else if (warp == 10) {
    // ================= TMA store issuer: tile j leaves as soon as the workers have updated it.  One lane: the bulk
    // async-group (commit / wait_group) is per THREAD state
    if (lane == 0) {
      for (int j = 0; j < nch; ++j) {
        const int s = j % p.ystages;
        const uint32_t ph = (uint32_t)(j / p.ystages) & 1u;
        ptx::mbar_wait_or_trap(&y_done[s], ph, __LINE__);
        uint8_t* yt = ytiles + (size_t)s * 32768;
        int tl, ch;
        unit_of(j, tl, ch);
        const int col = ch * 128;
        ptx::tma_store_2d(&tm_y, yt, col, tl * 128);
        ptx::tma_store_2d(&tm_y, yt + 16384, col + 64, tl * 128);
        ptx::tma_store_commit();
        ptx::tma_store_wait_read<0>();                         // smem read out: the slot can take the next load now
        ptx::mbar_arrive(&y_empty[s]);
        ptx::prog_mark(2, j + 1);
      }
      ptx::tma_store_wait_all<0>();
    }
  } else {
    // ================= workers: attention, operand staging, epilogue.  Two groups of 128 threads;
    // group g owns the chunks j = g (mod 2), TMEM accumulator g and B-tile buffer g, so the epilogue
    // of one chunk overlaps the waits / staging of the next.
    const int wt = threadIdx.x - 64;                 // 0..255
    const int grp_id = wt >> 7;                      // worker group
    const int wl = wt & 127;                         // thread within the group
    const int q = warp & 3;                          // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;                   // token row handled in the epilogue
    const int r = d.r;
    const SyncGroup grp{wt, 2 * kWorkers, 1, or_slot};
    AttnCtx c{tokmeta, kidx, klen, h_in0, nullptr, Ks, nullptr, r + 1, p.kcap,
              (p.write_g && split == 0) ? g_out + (size_t)d.T * d.r : nullptr, h_in1};

    // B operand tile of a chunk: rows = 128 output columns, K = rank (bf16, 128B swizzle).
    // Split in a global-load half (issued early, results parked in registers) and a
    // shared-store half, so the L2 latency hides behind the epilogue arithmetic.
    const int cpr_all = p.rk / 8;                               // 16-byte chunks per B row (all K blocks)
    const int per_thread = cpr_all;                             // 128 rows * cpr_all chunks / 128 threads
    // load_b: cp.async straight into the swizzled tile (no registers, latency hidden behind the
    // epilogue); store_b: wait for the copies, publish the tile to the tensor core
    auto load_b = [&](int j, int i) {
      for (int qq = 0; qq < per_thread; ++qq) {
        const int idx = wl + qq * kWorkers;
        const int n = idx / cpr_all, ch = idx - n * cpr_all;
        const int k = ch * 8;
        const int kb = ch >> 3, chb = ch & 7;                   // 64-wide K block and chunk inside it
        const int slot = (i >> 1) * p.bpg + (i & 1);              // barrier index i = group * 2 + buffer -> smem slot
        uint8_t* dst = btiles + ((size_t)slot * kblocks + kb) * 16384 + ptx::sw128_offset((uint32_t)n, (uint32_t)chb);
        const __nv_bfloat16* row_src = Bw + (size_t)(chunk_of(j) * 128 + n) * r;
        if (k + 8 <= r) {
          ptx::cp_async16(ptx::smem_u32(dst), row_src + k, true);
        } else if (k < r) {
          // r % 8 == 4: the last four rank values of the row (8 bytes, 8-byte aligned), the rest of the chunk is zero
          ptx::cp_async8(ptx::smem_u32(dst), row_src + k);
          *reinterpret_cast<uint2*>(dst + 8) = make_uint2(0u, 0u);
        } else {
          ptx::cp_async16(ptx::smem_u32(dst), Bw, false);          // beyond the rank: zero-fill (no bytes are read)
        }
      }
      ptx::cp_async_commit();
    };
    auto store_b = [&](int i) {
      ptx::cp_async_wait_all();
      ptx::fence_proxy_async_smem();
      warp_arrive(&b_full[i]);
    };

    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 0);
    if (grp_id < nch) load_b(grp_id, grp_id * 2);               // first two B tiles of each group, in flight early
    if (p.bpg == 2 && grp_id + 2 < nch) load_b(grp_id + 2, grp_id * 2 + 1);
    // g rows come finished from k_attn_fwd: build the A operand (post_m * g, bf16) of token tile [t0s, t0s + nts)
    // straight from global memory - every load of a thread is issued before its first use, so it costs one memory
    // round trip instead of a chain of them
    const int cpr_t = p.rk / 8;                                  // 16-byte chunks per token row
    constexpr int kMaxItems = 4;                                 // 128 * cpr_t / 256 items per thread (rk <= 64)
    auto stage_a = [&](int t0s, int nts, const int* pre_codes) {
      for (int base = 0; base < 128 * cpr_t; base += kMaxItems * 2 * kWorkers) {
        float4 ga[kMaxItems], gb[kMaxItems];
#pragma unroll
        for (int u = 0; u < kMaxItems; ++u) {
          const int idx = base + wt + u * 2 * kWorkers;
          const int n = idx / cpr_t, k = (idx - n * cpr_t) * 8;
          ga[u] = gb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx < 128 * cpr_t && n < nts) {
            if (k < r) ga[u] = *reinterpret_cast<const float4*>(h_in0 + (size_t)(t0s + n) * r + k);
            if (k + 4 < r) gb[u] = *reinterpret_cast<const float4*>(h_in0 + (size_t)(t0s + n) * r + k + 4);
          }
        }
#pragma unroll
        for (int u = 0; u < kMaxItems; ++u) {
          const int idx = base + wt + u * 2 * kWorkers;
          if (idx >= 128 * cpr_t) continue;
          const int n = idx / cpr_t, chf = idx - n * cpr_t;      // chunk over the full rank
          const int kb = chf >> 3, ch = chf & 7;
          int code;
          if (pre_codes && base == 0) code = pre_codes[u];
          else code = (n < nts) ? (tokmeta[t0s + n] & MOKA_MOD_MASK) : 0;
          const float ps = code ? d.post[code] : 0.f;
          uint4 v;
          v.x = pack_bf16(ga[u].x * ps, ga[u].y * ps); v.y = pack_bf16(ga[u].z * ps, ga[u].w * ps);
          v.z = pack_bf16(gb[u].x * ps, gb[u].y * ps); v.w = pack_bf16(gb[u].z * ps, gb[u].w * ps);
          *reinterpret_cast<uint4*>(gtile + (size_t)kb * 16384 + ptx::sw128_offset((uint32_t)n, (uint32_t)ch)) = v;
        }
      }
    };
    if (!p.attn_in_kernel) {
      int codes[kMaxItems];
#pragma unroll
      for (int u = 0; u < kMaxItems; ++u) {
        const int idx = wt + u * 2 * kWorkers;
        const int n = idx / cpr_t;
        codes[u] = (idx < 128 * cpr_t && n < nt) ? (tokmeta[t0 + n] & MOKA_MOD_MASK) : 0;
      }
      ptx::pdl_wait();                                           // g comes from k_attn_fwd
      if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 6);
      stage_a(t0, nt, codes);
      if (grp_id < nch) store_b(grp_id * 2);
      if (p.bpg == 2 && grp_id + 2 < nch) store_b(grp_id * 2 + 1);
      if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 2);
    } else {
    if (wt < 128) meta_s[wt] = (wt < nt) ? tokmeta[t0 + wt] : 0;
    ptx::pdl_wait();                                             // h comes from k_down_tc
    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 6);
    for (int i2 = wt; i2 < 128 * r; i2 += 2 * kWorkers) {
      float v = 0.f;
      if (i2 < nt * r) {
        v = h_in0[(size_t)t0 * r + i2];
        if (h_in1) v += h_in1[(size_t)t0 * r + i2];
        if (h_final && split == 0) h_final[(size_t)t0 * r + i2] = v;
      }
      gs[i2] = v;
    }
    grp.sync();
    if (grp_id < nch) store_b(grp_id * 2);
    if (p.bpg == 2 && grp_id + 2 < nch) store_b(grp_id * 2 + 1);
    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 1);
    if (p.attn_in_kernel) {
      attention_tile<false, 32>(d, c, grp, t0, nt, meta_s, gs, gs, nullptr);
      grp.sync();
      if (p.write_g && split == 0)
        for (int i2 = wt; i2 < nt * r; i2 += 2 * kWorkers) g_out[(size_t)t0 * r + i2] = gs[i2];
    }
    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 2);
    // A operand tile: post_m * g, bf16, rows = tokens
    for (int kb = 0; kb < kblocks; ++kb) {
      const int kw = min(64, p.rk - kb * 64);
      const int cpr = kw / 8;
      for (int idx = wt; idx < 128 * cpr; idx += 2 * kWorkers) {
        const int n = idx / cpr, ch = idx - n * cpr;
        const int k = kb * 64 + ch * 8;
        const int code = meta_s[n] & MOKA_MOD_MASK;
        const float ps = code ? d.post[code] : 0.f;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = (k + e < r) ? gs[n * r + k + e] * ps : 0.f;
        uint4 v;
        v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
        v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(gtile + (size_t)kb * 16384 + ptx::sw128_offset((uint32_t)n, (uint32_t)ch)) = v;
      }
    }
    }
    ptx::fence_proxy_async_smem();
    named_sync(2, 2 * kWorkers);
    if (wt == 0) ptx::mbar_arrive(g_full);
    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 3);

    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // segments: the runs of this CTA's units that share a token tile (exactly one unless stream-K).  BOTH worker
    // groups walk every segment (one of them may own none of its chunks), so that they meet at the boundaries.
    for (int j0 = 0, seg = 0; j0 < nch; ++seg) {
      int j1 = nch;
      if (p.streamk) {
        const long long u0 = u_beg + j0;
        const int tl = (int)(u0 / p.chunks);
        j1 = (int)min((long long)nch, (long long)(tl + 1) * p.chunks - u_beg);
        if (seg > 0) {
          // every epilogue of the previous tile is done once both groups are here, hence all of its MMAs have
          // retired: the A operand tile may be rewritten for the next token tile
          named_sync(2, 2 * kWorkers);
          stage_a(tl * 128, min(128, p.T - tl * 128), nullptr);
          ptx::fence_proxy_async_smem();
          named_sync(2, 2 * kWorkers);
          if (wt == 0) ptx::mbar_arrive(g_full);
        }
      }
    for (int j = j0 + ((grp_id ^ j0) & 1); j < j1; j += 2) {
      const int n = j >> 1;                                      // this group's chunk counter
      const int i = grp_id * 2 + (n & (p.bpg - 1));              // accumulator / B buffer
      const uint32_t ph = (uint32_t)(n / p.bpg) & 1u;
      const int ahead = 2 * p.bpg;                               // chunks until this buffer is used again
      const int s = j % p.ystages;
      const uint32_t yph = (uint32_t)(j / p.ystages) & 1u;
      ptx::mbar_wait_or_trap(&t_full[i], ph, __LINE__);
      if (wl == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 8 + j * 4);
      // MMA j has consumed B buffer i: restage it right away (async copies fly during the epilogue)
      if (j + ahead < nch) load_b(j + ahead, i);
      ptx::mbar_wait_or_trap(&y_full[s], yph, __LINE__);
      if (wl == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 9 + j * 4);
      ptx::tc_fence_after_sync();
      uint8_t* yt = ytiles + (size_t)s * 32768;
      if (p.delta)
        add_acc_to_tile<false, true>(lane_base + (uint32_t)(i * 128), ptx::smem_u32(yt), (uint32_t)row, &t_empty[i]);
      else
        add_acc_to_tile(lane_base + (uint32_t)(i * 128), ptx::smem_u32(yt), (uint32_t)row, &t_empty[i]);
      if (wl == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 10 + j * 4);
      ptx::fence_proxy_async_smem();                            // generic writes -> visible to the TMA store
      warp_arrive(&y_done[s]);
      if (wl == 0) ptx::prog_mark(3 + grp_id, j + 1);
      if (j + ahead < nch) store_b(i);
      if (wl == 0) ptx::prog_mark(5 + grp_id, j + 1);
      if (wl == 0 && j < 6) dbg_stamp(p.dbg, blockIdx.x, 11 + j * 4);
    }
      j0 = j1;
    }
    if (wt == 0) dbg_stamp(p.dbg, blockIdx.x, 4);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) { dbg_stamp(p.dbg, blockIdx.x, 5); dbg_stamp_g(p.dbg, blockIdx.x, 105); }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
}

static int launch_up_tc(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, bool partials) {
  UpTcParams p;
  p.T = d.T; p.d_out = d.d_out; p.r = d.r; p.rk = round_up(d.r, 16);
  p.chunks = d.d_out / 128;
  const int ntiles = (d.T + 127) / 128;
  p.nsplit = 1;
  while (ntiles * p.nsplit * 2 <= num_sms + num_sms / 4 && p.chunks / (p.nsplit * 2) >= 4) p.nsplit *= 2;
  {
    static int up_nsplit_env = -1;                 // tuning aid: force the column split of k_up_tc
    if (up_nsplit_env < 0) { const char* e = getenv("MOKA_UP_NSPLIT"); up_nsplit_env = e ? atoi(e) : 0; }
    if (up_nsplit_env > 0 && p.chunks / up_nsplit_env >= 2) p.nsplit = up_nsplit_env;
  }
  const int kblocks = (p.rk + 63) / 64;
  p.kcap = min(max(a.klen_hint, 1), 32 * kKeyBlocksMax);
  // attention as its own kernel (k_attn_fwd) whenever the sample's keys fit its shared memory
  const size_t attn_smem = ((size_t)p.kcap * (d.r + 1) + 8 + 2 * 32 * d.r) * 4 + 32;
  static int sep_env = -1;
  if (sep_env < 0) { const char* e = getenv("MOKA_ATTN_IN_UP"); sep_env = (e && e[0] == '1') ? 0 : 1; }
  const bool sep_attn = sep_env && a.klen_hint <= 32 * kKeyBlocksMax && attn_smem <= 160 * 1024 && d.r % 4 == 0;
  p.attn_in_kernel = sep_attn ? 0 : 1;
  if (sep_attn) p.kcap = 1;
  const int attn_kcap = min(max(a.klen_hint, 1), 32 * kKeyBlocksMax);
  // (the fp32 row staging `gs` is only needed when the attention runs inside this kernel)
  const size_t gs_bytes = sep_attn ? 0 : (size_t)128 * d.r * 4;
  // two B-tile buffers per worker group when they fit next to a 4-slot y ring, else one (r > 64)
  p.bpg = ((size_t)5 * kblocks * 16384 + gs_bytes + 1024 + 512 + 2 * 32768 <= 224 * 1024) ? 2 : 1;
  const size_t fixed = (size_t)(1 + 2 * p.bpg) * kblocks * 16384 + gs_bytes + 1024 + 512;
  const size_t limit = 224 * 1024;
  while (p.kcap > 1 && fixed + (size_t)p.kcap * (d.r + 1) * 4 + 2 * 32768 > limit) p.kcap /= 2;
  const size_t key_bytes = ((size_t)p.kcap * (d.r + 1) + 1) * 4;
  p.ystages = (int)((limit - fixed - key_bytes) / 32768);
  if (p.ystages > 4) p.ystages = 4;
  // The two worker groups alternate chunks, so with an EVEN ring every slot (and its y_full / y_done barriers)
  // belongs to one group.  With 3 slots the groups take turns on a slot and a group can ask for a phase two ahead
  // of the barrier; mbarrier parity waits cannot tell that from "done" (seen as a rare hang at r = 32 / 64).
  if (p.ystages == 3) p.ystages = 2;
  if (p.ystages < 2) return MOKA_EUNSUPPORTED;
  p.write_g = 1;
  p.delta = d.delta;
  p.rot = rot_enabled();
  p.dbg = dbg_enabled();
  // stream-K (persistent, balanced unit ranges; needs the attention as its own kernel).  ALONE it is slower (T = 32768:
  // 106.4 vs 100.9 us at d_out = 4096, 275 vs 253 us at 11008): this kernel reads and re-writes the same rows, and
  // with 148 CTAs in flight it already sits at what HBM gives that traffic mix (~5.5 TB/s), so balancing the tail buys
  // nothing and the per-segment restaging of the A operand costs.  INSIDE the training step (SM clock 1.5 GHz, every
  // CTA's hand-off chain slower) the second, 73 %-full wave of 256 tiles on 148 SMs is what costs: 204.7 -> 200.7 us
  // per forward call at d_out = 4096, 382.0 -> 374.3 us at 11008 (same box, twice; profiles/r02_dropout_cost.md).
  // Measured worse at one wave or less (T = 8192: 66.3 -> 68.3 us; T = 16384, r = 64: 167 -> 185 us in the step).
  // Hence: on when the token tiles make more than one wave with a last wave under 90 % full and the operand is
  // narrow (r <= 32); MOKA_UP_STREAMK=0 / 1 forces it.
  static int up_sk_env = -2;
  if (up_sk_env == -2) { const char* e = getenv("MOKA_UP_STREAMK"); up_sk_env = e ? ((e[0] == '1') ? 1 : 0) : -1; }
  p.units = (long long)ntiles * p.chunks;
  const int waves = (ntiles * p.nsplit + num_sms - 1) / num_sms;
  const bool ragged = waves > 1 && (double)(ntiles * p.nsplit) / ((double)waves * num_sms) < 0.9;
  const int want_sk = up_sk_env >= 0 ? up_sk_env : ((ragged && p.rk <= 32 && !d.delta) ? 1 : 0);
  p.streamk = (want_sk && sep_attn && p.units >= 2) ? 1 : 0;
  p.grid = p.streamk ? (int)std::min<long long>(num_sms, p.units) : ntiles * p.nsplit;
  CUtensorMap ty0, ty;
  if (!make_tmap(&ty0, d.delta ? a.y : a.y0, (uint64_t)d.T, (uint64_t)d.d_out, 128)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&ty, a.y, (uint64_t)d.T, (uint64_t)d.d_out, 128)) return MOKA_EUNSUPPORTED;
  const size_t smem = (size_t)p.ystages * 32768 + fixed + key_bytes;
  cudaError_t e = cudaFuncSetAttribute(k_up_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  const float* h0 = partials ? a.hpart : a.h_stash;
  const float* h1 = partials ? a.hpart + (size_t)d.T * d.r : nullptr;
  float* g_stash = a.h_stash + (size_t)d.T * d.r;
  if (sep_attn) {
    if (!(d.dbg_skip & 64)) {
      if (attn_smem > 40 * 1024) {
        e = cudaFuncSetAttribute(k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem);
        if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      }
      const int gps = (d.S + 31) / 32;
      float* hfin = partials ? a.h_stash : (float*)nullptr;
      float* lse = g_stash + (size_t)d.T * d.r;
      static int attn_simt = -1;
      if (attn_simt < 0) { const char* ev = getenv("MOKA_ATTN_SIMT"); attn_simt = (ev && ev[0] == '1') ? 1 : 0; }
      if (!attn_simt && attn_kcap <= 128 && d.r <= 128) {
        // warp-level tensor cores (mma.sync, hi/lo split operands)
        const int R = round_up(d.r, 16) <= 16 ? 16 : (d.r <= 32 ? 32 : (d.r <= 64 ? 64 : 128));
        const int LP = attn_kcap <= 64 ? 64 : 128;
        const size_t smm = (size_t)64 * (R + 4) * 4 + (size_t)2 * LP * (R + 8) * 2 + (size_t)2 * R * (LP + 8) * 2;
        const int gps64 = (d.S + 63) / 64;
#define MOKA_LAUNCH_ATTN_MMA(RR, LL)                                                                                   \
        do {                                                                                                           \
          if (smm > 40 * 1024) cudaFuncSetAttribute(k_attn_fwd_mma<RR, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smm); \
          e = launch_pdl(k_attn_fwd_mma<RR, LL>, dim3(d.B * gps64), dim3(256), smm, st, d, gps64, a.tokmeta, a.kidx, a.klen, \
                         h0, h1, hfin, g_stash, lse);                                                                  \
        } while (0)
        if (R == 16) { if (LP == 64) MOKA_LAUNCH_ATTN_MMA(16, 64); else MOKA_LAUNCH_ATTN_MMA(16, 128); }
        else if (R == 32) { if (LP == 64) MOKA_LAUNCH_ATTN_MMA(32, 64); else MOKA_LAUNCH_ATTN_MMA(32, 128); }
        else if (R == 64) { if (LP == 64) MOKA_LAUNCH_ATTN_MMA(64, 64); else MOKA_LAUNCH_ATTN_MMA(64, 128); }
        else { if (LP == 64) MOKA_LAUNCH_ATTN_MMA(128, 64); else MOKA_LAUNCH_ATTN_MMA(128, 128); }
#undef MOKA_LAUNCH_ATTN_MMA
      } else if (d.r <= 32) {
        const int R = d.r <= 8 ? 8 : (d.r <= 16 ? 16 : 32);
        const size_t sm2 = ((size_t)attn_kcap * (R + 4) + 32 * R) * 4;
        if (R == 8) e = launch_pdl(k_attn_fwd_split<8>, dim3(d.B * gps), dim3(256), sm2, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen, h0, h1, hfin, g_stash, lse);
        else if (R == 16) e = launch_pdl(k_attn_fwd_split<16>, dim3(d.B * gps), dim3(256), sm2, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen, h0, h1, hfin, g_stash, lse);
        else e = launch_pdl(k_attn_fwd_split<32>, dim3(d.B * gps), dim3(256), sm2, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen, h0, h1, hfin, g_stash, lse);
      } else if (d.r <= 128) {
        const int C = d.r <= 64 ? 2 : 4;
        const size_t sm3 = ((size_t)((attn_kcap + 1) & ~1) * 32 * C + 32 * 32 * C) * 4;
        if (sm3 > 40 * 1024) {                   // (static shared memory counts towards the 48 KB default too)
          cudaFuncSetAttribute(k_attn_fwd_rank<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
          cudaFuncSetAttribute(k_attn_fwd_rank<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
        }
        if (C == 2) e = launch_pdl(k_attn_fwd_rank<2>, dim3(d.B * gps), dim3(256), sm3, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen, h0, h1, hfin, g_stash, lse);
        else e = launch_pdl(k_attn_fwd_rank<4>, dim3(d.B * gps), dim3(256), sm3, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen, h0, h1, hfin, g_stash, lse);
      } else
      e = launch_pdl(k_attn_fwd, dim3(d.B * gps), dim3(256), attn_smem, st, d, attn_kcap, gps, a.tokmeta, a.kidx, a.klen,
                     h0, h1, hfin, g_stash, lse);
      if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      if (int rc = check_launch_tc()) return rc;
    }
    if (d.dbg_skip & 2) return MOKA_OK;
    e = launch_pdl(k_up_tc, dim3(p.grid), dim3(kUpThreads), smem, st, ty0, ty, d, p,
                   static_cast<const __nv_bfloat16*>(a.Bw), a.tokmeta, a.kidx, a.klen, (const float*)g_stash,
                   (const float*)nullptr, (float*)nullptr, g_stash);
  } else {
    if (d.dbg_skip & 2) return MOKA_OK;
    e = launch_pdl(k_up_tc, dim3(p.grid), dim3(kUpThreads), smem, st, ty0, ty, d, p,
                   static_cast<const __nv_bfloat16*>(a.Bw), a.tokmeta, a.kidx, a.klen, h0, h1,
                   partials ? a.h_stash : (float*)nullptr, g_stash);
  }
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  return check_launch_tc();
}

// ---------------------------------------------------------------------------- k_dgrad_tc
// Streams dY once.  Per 128-token tile and per pair of 64-column TMA boxes (128 columns):
//   MMA1  dG[128 tok, rk]   += dY[128 tok, 128 cols] . B[128 cols, rk]          (dY K-major)
//   MMA2  dB[128 cols, rk]   = dY^T[128 cols, 128 tok] . (post g)[128 tok, rk]  (dY MN-major, the
//         same shared-memory tile read through a transposing descriptor; g split hi+lo bf16)
// dB leaves through fp32 vector reductions; dG as plain stores into per-CTA PARTIAL buffers that the consumer
// (k_attn_pack) sums.
//
// Persistent, stream-K style (round 2): the work is the list of U = ntiles x P (tile, 128-column pair) units in
// tile-major order; CTA c of a one-CTA-per-SM grid of G owns the contiguous unit range
// [floor(c U / G), floor((c+1) U / G)), so every SM gets the same number of bytes whatever the shape (256 tiles on
// 148 SMs used to be two waves of one CTA per SM, the second 27 % empty) and the TMA ring never drains at a tile
// boundary: the producer simply walks on into the next tile.  A range covers the tail of one tile, whole tiles and
// the head of another; each (tile, CTA) SEGMENT has its own dG accumulator (two in TMEM, so the next segment's MMA1
// starts while the workers drain the last one) and writes partial number (c - first CTA of the tile).  The G^T
// operand of the next segment is fetched into registers while the current one streams and is written to shared
// memory once the current segment's MMAs have retired.
struct DgradTcParams {
  int T, d_out, r, rk;
  int pairs;                       // P = d_out / 128 column pairs per token tile
  int grid;                        // G persistent CTAs
  long long units;                 // U = ntiles * P
  int stages;
  uint32_t stage_bytes, bt_bytes;  // stage = [pair 0: box0 16K, box1 16K]..[pair W-1][Bt blocks: 2 per pair]
  int dbg;
  int dg_mode;                     // 0: plain stores into partial buffer (c - first CTA of the tile); 2: reduce into zeroed dG0
  int nbuf;                        // dB accumulator sets in TMEM (2, or 1 when rk = 128)
  int ndg;                         // dG accumulators in TMEM (2, or 1 when TMEM is short)
  int wide;                        // 128-column pairs per pipeline stage (2 when TMEM / smem allow: halves the
                                   // per-stage hand-offs of the single MMA-issuing thread)
  uint32_t tmem_cols, dg_cols, d2_col;   // dG buffers at [0, ndg*dg_cols), D2 sets from d2_col
  long long part_stride;           // floats between two dG partial buffers
};

// Which CTAs of the persistent k_dgrad_tc grid touch token tile `tile` (see the kernel): shared with the consumer
// of the partial dG buffers.  start(c) = floor(c U / G); the CTA of unit u is floor(((u + 1) G - 1) / U).
struct DgParts {
  long long U;
  long long stride;                // floats between partial buffers
  int P, G;                        // G == 0: one complete buffer (no partials)
};
__host__ __device__ __forceinline__ int dg_cta_of_unit(long long U, int G, long long u) {
  return (int)(((u + 1) * (long long)G - 1) / U);
}
__host__ __device__ __forceinline__ int dg_nparts(const DgParts& q, int tile) {
  if (q.G == 0) return 1;
  const long long u0 = (long long)tile * q.P;
  return dg_cta_of_unit(q.U, q.G, u0 + q.P - 1) - dg_cta_of_unit(q.U, q.G, u0) + 1;
}

static int dgrad_grid(int T, int d_out, int num_sms) {
  const long long U = (long long)((T + 127) / 128) * (d_out / 128);
  return (int)std::min<long long>(num_sms, std::max<long long>(U, 1));
}
// upper bound on dg_nparts over all tiles (workspace sizing)
int tc_dgrad_parts_max(int T, int d_out, int num_sms) {
  const int P = d_out / 128, ntiles = (T + 127) / 128;
  if (P <= 0) return 2;
  DgParts q{(long long)ntiles * P, 0, P, dgrad_grid(T, d_out, num_sms)};
  int mx = 2;
  for (int t = 0; t < ntiles; ++t) mx = std::max(mx, dg_nparts(q, t));
  return mx;
}

// the (tile, pair range) segments of a CTA's unit range, in order
struct DgSegIter {
  long long u, u_end;
  int P;
  int tile, pb, pe;                // current segment: pairs [pb, pe) of `tile`
  __device__ __forceinline__ bool next() {
    if (u >= u_end) return false;
    tile = (int)(u / P);
    const long long t0u = (long long)tile * P;
    pb = (int)(u - t0u);
    const long long e = (t0u + P < u_end) ? t0u + P : u_end;
    pe = (int)(e - t0u);
    u = e;
    return true;
  }
};

// warps: 0 TMA producer, 1 issues the dG products (and owns TMEM), 2 issues the dB products, 3..6 workers.  Two
// issuing warps (round 2): the per-stage waits / fence / commits of one overlap the tensor-pipe time of the other's
// products - with a single issuer the pipe idled ~25 % of every stage, and that share grows as the SM clock drops
// under the power cap of the surrounding GEMMs.
constexpr int kDgradThreads = 224;
__global__ void __launch_bounds__(kDgradThreads)
k_dgrad_tc(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_bt, DevDesc d,
           DgradTcParams p, const uint8_t* __restrict__ tokmeta,
           const float* __restrict__ g_stash, float* __restrict__ dG0, float* __restrict__ dB,
           uint32_t* __restrict__ zero_ptr, int zero_words) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t gt_block = (uint32_t)p.rk * 128u;            // one 64-column K block of B^T (rk rows)
  const uint32_t g2_block = 2u * gt_block;                    // one 64-token K block of [G^T hi ; G^T lo] (2 rk rows)
  uint8_t* gt = ring + (size_t)p.stages * p.stage_bytes;      // [2 kblocks][2 rk x 128 B]: rows 0..rk-1 hi, rk..2rk-1 lo
  uint64_t* bars = reinterpret_cast<uint64_t*>(gt + 2 * g2_block);
  uint64_t* full = bars;                       // TMA landed            [stages]
  uint64_t* empty = full + p.stages;           // MMAs done with stage  [stages]
  uint64_t* d2_full = empty + p.stages;        // [2]
  uint64_t* d2_empty = d2_full + 2;            // [2]
  uint64_t* dg_full = d2_empty + 2;            // [2] dG accumulator of a segment final
  uint64_t* dg_empty = dg_full + 2;            // [2] ... read out by the workers
  uint64_t* gt_full = dg_empty + 2;            // [1] G^T operand of the next segment staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gt_full + 1);

  const int warp = ptx::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const long long u_beg = (long long)cta * p.units / p.grid;
  const long long u_end = (long long)(cta + 1) * p.units / p.grid;
  const int W = p.wide;
  const uint32_t dy_bytes = 32768u * (uint32_t)W;             // dY part of a stage; the B^T blocks follow
  const int r = p.r;
  if (threadIdx.x == 0) { dbg_stamp_g(p.dbg, blockIdx.x, 108); dbg_stamp(p.dbg, blockIdx.x, 0); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 2); }   // (both issuers commit)
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&d2_full[i], 1); ptx::mbar_init(&d2_empty[i], kWorkerWarps);
      ptx::mbar_init(&dg_full[i], 1); ptx::mbar_init(&dg_empty[i], kWorkerWarps);
    }
    ptx::mbar_init(gt_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_dy); ptx::prefetch_tmap(&tm_bt);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, p.tmem_cols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // successors may start early: before their waits they only read the forward's stash, the token
  // metadata, x / dx and the adapter weights - all older than k_bwd_prep, which triggered after its wait
  if (threadIdx.x == 0) ptx::pdl_trigger();

  if (warp == 0) {
    // ---- TMA producer (whole warp, one elected lane issues)
    {
      // dy is older than this call's first kernel (k_bwd_prep waited for it): its tiles stream in
      // right away; the B^T blocks are written by k_bwd_prep, so they follow the dependency wait
      const uint32_t ring_a = ptx::uniform_u32(ptx::smem_u32(ring));
      auto load_dy = [&](int tile, int q0, int nq, uint64_t* bar, uint32_t st) {
        ptx::mbar_arrive_expect_tx_w(bar, (uint32_t)nq * (32768u + p.bt_bytes));
        for (int q = 0; q < nq; ++q) {
          const int col = (q0 + q) * 128;
          ptx::tma_load_2d_w(st + (uint32_t)q * 32768u, &tm_dy, bar, col, tile * 128, ptx::kEvictFirst);
          ptx::tma_load_2d_w(st + (uint32_t)q * 32768u + 16384u, &tm_dy, bar, col + 64, tile * 128, ptx::kEvictFirst);
        }
      };
      auto load_bt = [&](int q0, int nq, uint64_t* bar, uint32_t st) {   // B^T blocks [rk rows x 64 cols] (L2-resident)
        for (int q = 0; q < nq; ++q) {
          const int col = (q0 + q) * 128;
          ptx::tma_load_2d_w(st + dy_bytes + (uint32_t)(2 * q) * gt_block, &tm_bt, bar, col, 0, ptx::kEvictLast);
          ptx::tma_load_2d_w(st + dy_bytes + (uint32_t)(2 * q + 1) * gt_block, &tm_bt, bar, col + 64, 0, ptx::kEvictLast);
        }
      };
      // first fill of the ring: dY before the dependency wait, the B^T blocks of the same stages after it
      {
        DgSegIter it{u_beg, u_end, p.pairs, 0, 0, 0};
        int j = 0;
        while (j < p.stages && it.next())
          for (int q0 = it.pb; q0 < it.pe && j < p.stages; q0 += W, ++j)
            load_dy(it.tile, q0, min(W, it.pe - q0), &full[j], ring_a + (uint32_t)j * p.stage_bytes);
      }
      ptx::pdl_wait();
      DgSegIter it{u_beg, u_end, p.pairs, 0, 0, 0};
      int j = 0;
      while (it.next()) {
        for (int q0 = it.pb; q0 < it.pe; q0 += W, ++j) {
          const int nq = min(W, it.pe - q0);
          const int s = j % p.stages;
          const uint32_t st = ring_a + (uint32_t)s * p.stage_bytes;
          if (j >= p.stages) {
            ptx::mbar_wait_or_trap(&empty[s], (((uint32_t)(j / p.stages)) & 1u) ^ 1u, __LINE__);
            load_dy(it.tile, q0, nq, &full[s], st);
          }
          load_bt(q0, nq, &full[s], st);
        }
      }
    }
  } else if (warp == 1) {
    // ---- issuer of MMA1, dG[128 tok, rk] += dY . B.  The WHOLE warp runs this loop on warp-uniform values, one
    // elected lane issues (see moka_ptx.cuh).  Needs only the TMA data of the stage.
    {
      const uint32_t tb = ptx::uniform_u32(tmem_base);
      const uint32_t idesc1 = ptx::idesc_bf16(128, (uint32_t)p.rk, 0, 0);     // dG: A K-major, B K-major
      const uint32_t ring_a = ptx::uniform_u32(ptx::smem_u32(ring));
      // descriptors differ only in their 14-bit start-address field: build once, add offsets (>> 4)
      const uint64_t a1_base = ptx::smem_desc_sw128(0, 1024, 0);
      const uint32_t gtb = gt_block >> 4;
      DgSegIter it{u_beg, u_end, p.pairs, 0, 0, 0};
      int j = 0;
      int s = 0; uint32_t s_ph = 0;                   // ring slot of stage j, parity of its `full` phase
      int dgb = 0; uint32_t dg_ph = 0;                // dG accumulator of the segment
      while (it.next()) {
        const uint32_t dg_acc = tb + (uint32_t)dgb * p.dg_cols;
        ptx::mbar_wait_or_trap(&dg_empty[dgb], dg_ph ^ 1u, __LINE__);          // accumulator read out (two segments ago)
        ptx::tc_fence_after_sync();
        for (int q0 = it.pb; q0 < it.pe; q0 += W, ++j) {
          const int nq = min(W, it.pe - q0);
          ptx::mbar_wait_or_trap(&full[s], s_ph, __LINE__);
          if (lane == 0) { if (j == 0) dbg_stamp(p.dbg, blockIdx.x, 3); if (j < 8) dbg_stamp(p.dbg, blockIdx.x, 8 + j * 3); }
          const uint32_t sa0 = ring_a + (uint32_t)s * p.stage_bytes;
          for (int q = 0; q < nq; ++q) {
            const uint32_t sa = sa0 + (uint32_t)q * 32768u;
            const uint32_t bta = sa0 + dy_bytes + (uint32_t)(2 * q) * gt_block;
            // K = 128 columns of this pair (two boxes, four 16-wide steps each)
            const uint64_t a1 = a1_base + (uint64_t)(sa >> 4);
            const uint64_t b1 = a1_base + (uint64_t)(bta >> 4);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int box = k >> 2, kk = k & 3;
              if (!(p.dbg & 4))
              ptx::umma_bf16_w(dg_acc, a1 + (uint64_t)((box * 16384 + kk * 32) >> 4),
                               b1 + (uint64_t)(box * gtb + ((kk * 32) >> 4)), idesc1, (q0 > it.pb || q > 0 || k > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit_w(&empty[s]);
          if (++s == p.stages) { s = 0; s_ph ^= 1u; }
        }
        ptx::umma_commit_w(&dg_full[dgb]);
        if (++dgb == p.ndg) { dgb = 0; dg_ph ^= 1u; }
      }
      if (lane == 0) dbg_stamp(p.dbg, blockIdx.x, 4);
    }
  } else if (warp == 2) {
    // ---- issuer of MMA2, dB[128 cols, 2 rk] = dY^T . [g_hi | g_lo]: A MN-major (transposed view of the dY tile);
    // hi and lo halves of g stacked along N, so one pass over the tile yields both products (the epilogue adds them).
    // Needs the G^T operand of the segment and a free set of dB accumulators.
    {
      const uint32_t tb = ptx::uniform_u32(tmem_base);
      const uint32_t idesc2 = ptx::idesc_bf16(128, 2u * (uint32_t)p.rk, 1, 0);
      const uint32_t gta = ptx::uniform_u32(ptx::smem_u32(gt));
      const uint32_t ring_a = ptx::uniform_u32(ptx::smem_u32(ring));
      const uint64_t a2_base = ptx::smem_desc_sw128(0, 1024, 16384);
      const uint64_t g_base = ptx::smem_desc_sw128(0, 1024, 0) + (uint64_t)(gta >> 4);
      const uint32_t g2b = g2_block >> 4;
      DgSegIter it{u_beg, u_end, p.pairs, 0, 0, 0};
      int j = 0, seg = 0;
      int s = 0; uint32_t s_ph = 0;
      int i2 = 0; uint32_t i2_ph = 0;                 // dB accumulator set of stage j, parity of its k-th use
      while (it.next()) {
        ptx::mbar_wait_or_trap(gt_full, (uint32_t)seg & 1u, __LINE__);
        if (lane == 0 && seg == 0) dbg_stamp(p.dbg, blockIdx.x, 2);
        for (int q0 = it.pb; q0 < it.pe; q0 += W, ++j) {
          const int nq = min(W, it.pe - q0);
          ptx::mbar_wait_or_trap(&full[s], s_ph, __LINE__);
          ptx::mbar_wait_or_trap(&d2_empty[i2], i2_ph ^ 1u, __LINE__);
          if (lane == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 9 + j * 3);
          ptx::tc_fence_after_sync();
          const uint32_t sa0 = ring_a + (uint32_t)s * p.stage_bytes;
          // M = 128 columns (two 64-wide atoms 16 KB apart), K = 128 tokens in 8 steps of 16
          for (int q = 0; q < nq; ++q) {
            const uint32_t sa = sa0 + (uint32_t)q * 32768u;
            const uint64_t a2 = a2_base + (uint64_t)(sa >> 4);
            const uint32_t d2 = tb + p.d2_col + (uint32_t)(i2 * W + q) * 2u * (uint32_t)p.rk;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (!(p.dbg & 8))
              ptx::umma_bf16_w(d2, a2 + (uint64_t)((k * 2048) >> 4), g_base + (uint64_t)((k >> 2) * g2b + (((k & 3) * 32) >> 4)),
                               idesc2, k > 0 ? 1u : 0u);
          }
          ptx::umma_commit_w(&empty[s]);
          ptx::umma_commit_w(&d2_full[i2]);
          if (lane == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 10 + j * 3);
          if (++s == p.stages) { s = 0; s_ph ^= 1u; }
          if (++i2 == p.nbuf) { i2 = 0; i2_ph ^= 1u; }
        }
        ++seg;
      }
    }
  } else {
    const int wt = threadIdx.x - 96;                           // workers: warps 3..6 (their warp % 4 covers every TMEM lane quadrant)
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    // G^T operand (post_m * g as hi + lo bf16): rows = rank index, K = tokens.  `gq` = 16 rank values of this
    // thread's token row starting at j0 (fp32, already scaled by post).
    const uint32_t kblk = (uint32_t)(row >> 6) * g2_block, ch = (uint32_t)((row & 63) >> 3), in = (row & 7) * 2;
    auto load_g16 = [&](int tile, int j0, float4 (&gq)[4], uint8_t& mb) {
      const int t = tile * 128 + row;
      const bool live = t < p.T;
      mb = live ? tokmeta[t] : (uint8_t)0;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        gq[u] = (live && j0 + 4 * u < r) ? *reinterpret_cast<const float4*>(g_stash + (size_t)t * r + j0 + 4 * u)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto store_g16 = [&](int j0, const float4 (&gq)[4], uint8_t mb) {
      const int code = mb & MOKA_MOD_MASK;
      const float ps = code ? d.post[code] : 0.f;
      const float gvals[16] = {gq[0].x, gq[0].y, gq[0].z, gq[0].w, gq[1].x, gq[1].y, gq[1].z, gq[1].w,
                               gq[2].x, gq[2].y, gq[2].z, gq[2].w, gq[3].x, gq[3].y, gq[3].z, gq[3].w};
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int jj = j0 + u;
        const float gv = gvals[u] * ps;
        const __nv_bfloat16 hi = __float2bfloat16_rn(gv);
        const __nv_bfloat16 lo = __float2bfloat16_rn(gv - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16*>(gt + kblk + ptx::sw128_offset((uint32_t)jj, ch) + in) = hi;
        *reinterpret_cast<__nv_bfloat16*>(gt + kblk + ptx::sw128_offset((uint32_t)(p.rk + jj), ch) + in) = lo;
      }
    };
    auto stage_gt_direct = [&](int tile) {
      for (int j0 = 0; j0 < p.rk; j0 += 16) {
        float4 gq[4];
        uint8_t mb;
        load_g16(tile, j0, gq, mb);
        store_g16(j0, gq, mb);
      }
    };
    auto publish_gt = [&]() {
      ptx::fence_proxy_async_smem();
      named_sync(2, kWorkers);
      if (wt == 0) ptx::mbar_arrive(gt_full);
    };
    DgSegIter it{u_beg, u_end, p.pairs, 0, 0, 0};
    bool have = it.next();
    if (have) { stage_gt_direct(it.tile); publish_gt(); }
    // dB / dG are written below: everything older than k_bwd_prep (earlier calls that shared the workspace)
    // must be complete
    ptx::pdl_wait();
    // zero-fill of the small reduction targets of the NEXT kernel (per-sample dK, arrival counters of
    // k_attn_pack): that kernel's dependency wait covers this whole grid
    for (int i2 = cta * kWorkers + wt; i2 < zero_words; i2 += p.grid * kWorkers) zero_ptr[i2] = 0u;

    const uint32_t lane_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const bool preload = (p.rk == 16);                           // next segment's g rows parked in registers
    int j = 0, seg = 0;
    while (have) {
      const int tile = it.tile, pb = it.pb, pe = it.pe;
      DgSegIter nx = it;
      const bool has_next = nx.next();
      float4 gqn[4];
      uint8_t mbn = 0;
      if (has_next && preload) load_g16(nx.tile, 0, gqn, mbn);
      for (int q0 = pb; q0 < pe; q0 += W, ++j) {
        const int nq = min(W, pe - q0);
        const int i = j % p.nbuf;
        const uint32_t k2 = (uint32_t)(j / p.nbuf);
        ptx::mbar_wait_or_trap(&d2_full[i], k2 & 1u, __LINE__);
        if (wt == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 32 + j * 2);
        ptx::tc_fence_after_sync();
        for (int q = 0; q < nq; ++q) {
          const int col = (q0 + q) * 128 + row;                  // TMEM lane <-> output column
          float* dst = dB + (size_t)col * r;
          const uint32_t acc = lane_base + p.d2_col + (uint32_t)((i * W + q) * 2 * p.rk);
          for (int c8 = 0; c8 < r; c8 += 8) {
            float v[8], w[8];
            ptx::tmem_ld8(acc + (uint32_t)c8, v);                 // hi product
            ptx::tmem_ld8(acc + (uint32_t)(p.rk + c8), w);        // lo product
            ptx::tmem_ld_wait();
            if (p.dbg & 2) continue;                              // timing experiment: no dB reductions
            if (c8 + 8 <= r) {
              red_add_v4_f32(dst + c8, v[0] + w[0], v[1] + w[1], v[2] + w[2], v[3] + w[3]);
              red_add_v4_f32(dst + c8 + 4, v[4] + w[4], v[5] + w[5], v[6] + w[6], v[7] + w[7]);
            } else {                                              // r % 8 == 4: the last half group
              red_add_v4_f32(dst + c8, v[0] + w[0], v[1] + w[1], v[2] + w[2], v[3] + w[3]);
            }
          }
        }
        ptx::tc_fence_before_sync();
        warp_arrive(&d2_empty[i]);
        if (wt == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 33 + j * 2);
      }
      // ---- end of the segment: its MMAs must retire before the G^T buffer is rewritten
      const int dgb = seg % p.ndg;
      ptx::mbar_wait_or_trap(&dg_full[dgb], ((uint32_t)(seg / p.ndg)) & 1u, __LINE__);
      if (wt == 0 && seg == 0) dbg_stamp(p.dbg, blockIdx.x, 5);
      ptx::tc_fence_after_sync();
      if (has_next) {
        if (preload) store_g16(0, gqn, mbn); else stage_gt_direct(nx.tile);
        publish_gt();
      }
      {
        // dG partial of this segment.  tcgen05.ld is warp-collective: every lane issues it, only live rows store
        const int t = tile * 128 + row;
        const bool live = t < p.T;
        const int code = live ? (tokmeta[t] & MOKA_MOD_MASK) : 0;
        const float ps = code ? d.post[code] : 0.f;
        const int part = cta - dg_cta_of_unit(p.units, p.grid, (long long)tile * p.pairs);
        float* dGp = dG0 + (p.dg_mode == 2 ? 0 : (size_t)part * (size_t)p.part_stride);
        const uint32_t acc = lane_base + (uint32_t)dgb * p.dg_cols;
        for (int c8 = 0; c8 < p.rk; c8 += 8) {
          float v[8];
          ptx::tmem_ld8(acc + (uint32_t)c8, v);
          ptx::tmem_ld_wait();
          if (c8 < r && live) {
            float* dst = dGp + (size_t)t * r + c8;
            const bool full8 = c8 + 8 <= r;
            if (p.dg_mode == 2) {
              if (code) {
                red_add_v4_f32(dst, v[0] * ps, v[1] * ps, v[2] * ps, v[3] * ps);
                if (full8) red_add_v4_f32(dst + 4, v[4] * ps, v[5] * ps, v[6] * ps, v[7] * ps);
              }
            } else {
              // plain stores (zeros for tokens in no mask): nothing has to be zero-filled beforehand
              *reinterpret_cast<float4*>(dst) = make_float4(v[0] * ps, v[1] * ps, v[2] * ps, v[3] * ps);
              if (full8) *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] * ps, v[5] * ps, v[6] * ps, v[7] * ps);
            }
          }
        }
        ptx::tc_fence_before_sync();
        warp_arrive(&dg_empty[dgb]);
      }
      it = nx;
      have = has_next;
      ++seg;
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) { dbg_stamp(p.dbg, blockIdx.x, 6); dbg_stamp_g(p.dbg, blockIdx.x, 109); }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------------------- k_attn_bwd
// dG [T, r] -> dh [T, r] (+ dK into the per-sample scratch): the attention backward of
// SURVEY.md A.3 on 32-token tiles, all data O(T r).  Runs between k_dgrad_tc and k_wgrad_tc.
__global__ void __launch_bounds__(256)
k_attn_bwd(DevDesc d, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
           const int32_t* __restrict__ klen, const float* __restrict__ h_stash, float* __restrict__ dh,
           float* __restrict__ dK_glob, int kcap) {
  extern __shared__ __align__(16) float smem_f[];
  const int r = d.r;
  const int TT = d.tile_tokens;
  float* hs = smem_f;
  float* dgs = hs + (size_t)TT * r;
  float* Ks = dgs + (size_t)TT * r;
  float* dKs = Ks + (size_t)kcap * (r + 1);
  uint8_t* meta_s = reinterpret_cast<uint8_t*>(dKs + (size_t)kcap * (r + 1));
  __shared__ int s_or;
  const SyncGroup grp{(int)threadIdx.x, 256, 0, &s_or};
  AttnCtx c{tokmeta, kidx, klen, h_stash, dK_glob, Ks, dKs, r + 1, kcap, nullptr};
  const int tid = threadIdx.x;
  for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
    const int t0 = tile * TT;
    const int nt = min(TT, d.T - t0);
    int anyq = 0;
    for (int i = tid; i < nt; i += 256) {
      const uint8_t mb = tokmeta[t0 + i];
      meta_s[i] = mb;
      anyq |= (mb & MOKA_MOD_MASK) >= MOKA_MOD_VIS;
    }
    if (!grp.sync_or(anyq)) continue;                        // dh == dG for tiles without query tokens
    for (int i = tid; i < nt * r; i += 256) { hs[i] = h_stash[(size_t)t0 * r + i]; dgs[i] = dh[(size_t)t0 * r + i]; }
    __syncthreads();
    attention_tile<true>(d, c, grp, t0, nt, meta_s, hs, nullptr, dgs);
    __syncthreads();
    for (int i = tid; i < nt * r; i += 256)
      if ((meta_s[i / r] & MOKA_MOD_MASK) >= MOKA_MOD_VIS) dh[(size_t)t0 * r + i] = dgs[i];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- k_attn_bwd_tile
// Fast attention backward for r <= 32 and <= 128 keys per sample.  A CTA owns 32 consecutive tokens
// of one sample: their h / g / dG rows and all key rows of the sample are staged in shared memory
// with coalesced loads; a warp takes 4 tokens and, per 32-key block, maps lane <-> key.  Scores are
// recomputed from the stashed log-sum-exp, <p, dp> is dg.(g - h), the per-token dh contribution is
// reduced over the keys with a shuffle butterfly and written back IN PLACE (the CTA owns the rows),
// the per-key dK sums live in registers across the warp's tokens, are combined across warps in
// shared memory and leave with one reduction per (key, rank) element per CTA.
template <int R, int KB>
__global__ void __launch_bounds__(256)
k_attn_bwd_tile(DevDesc d, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
                const int32_t* __restrict__ klen, const float* __restrict__ h_stash, float* __restrict__ dh,
                float* __restrict__ dK_glob, int groups_per_sample) {
  constexpr int TT = 32;
  extern __shared__ __align__(16) float sm[];
  const int r = d.r;
  constexpr int RS = R + 1;                         // odd row stride: lane <-> key reads are conflict-free
  float* Ks = sm;                                   // [KB*32][RS]   (zero rows for invalid keys)
  float* dKs = Ks + KB * 32 * RS;                   // [KB*32][RS]
  float* hs = dKs + KB * 32 * RS;                   // [TT][R]
  float* gs = hs + TT * R;
  float* dgs = gs + TT * R;
  __shared__ int s_code[TT];
  __shared__ float s_lse[TT];
  __shared__ int s_anyq;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / groups_per_sample;
  const int s0 = (blockIdx.x % groups_per_sample) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const float* g_stash = h_stash + (size_t)d.T * r;
  const float* lse = g_stash + (size_t)d.T * r;
  const int L = klen[b];
  if (tid == 0) s_anyq = 0;
  __syncthreads();
  if (tid < TT) {
    const int code = (tid < nt) ? (tokmeta[tbase + tid] & MOKA_MOD_MASK) : 0;
    s_code[tid] = code;
    s_lse[tid] = (tid < nt && code >= MOKA_MOD_VIS) ? lse[tbase + tid] : 0.f;
    if (code >= MOKA_MOD_VIS) s_anyq = 1;
  }
  __syncthreads();
  if (!s_anyq || L <= 0) return;                    // no query token here: dh == dG already
  // stage rows (coalesced) and keys
  for (int i = tid; i < TT * R; i += 256) {
    const int tt = i / R, j = i - tt * R;
    const bool ok = tt < nt && j < r;
    hs[i] = ok ? h_stash[(tbase + tt) * r + j] : 0.f;
    gs[i] = ok ? g_stash[(tbase + tt) * r + j] : 0.f;
    dgs[i] = ok ? dh[(tbase + tt) * r + j] : 0.f;
  }
  for (int k = tid; k < KB * 32; k += 256) {
    bool valid = false;
    size_t tk = 0;
    if (k < L) {
      const int pos = kidx[(size_t)b * d.Lk + k];
      tk = (size_t)b * d.S + pos;
      valid = (tokmeta[tk] & MOKA_KEY_VALID) != 0;
    }
    for (int j = 0; j < R; ++j) {
      Ks[k * RS + j] = (valid && j < r) ? h_stash[tk * r + j] : 0.f;
      dKs[k * RS + j] = 0.f;
    }
  }
  __syncthreads();
  const float sl = d.sm_scale * kLog2e;
  float dKacc[KB][R];
#pragma unroll
  for (int kb = 0; kb < KB; ++kb)
#pragma unroll
    for (int j = 0; j < R; ++j) dKacc[kb][j] = 0.f;
  const int comp = (R == 32) ? lane : (R == 16 ? (lane >> 1) : (lane >> 2));
  const bool writer = (R == 32) || (R == 16 && !(lane & 1)) || (R == 8 && !(lane & 3));
  for (int ti = 0; ti < 4; ++ti) {
    const int tt = warp * 4 + ti;
    if (s_code[tt] < MOKA_MOD_VIS) continue;        // warp-uniform
    float hq[R], dg[R];
    float Dsum = 0.f;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      hq[j] = hs[tt * R + j];
      dg[j] = dgs[tt * R + j];
      Dsum = fmaf(dg[j], gs[tt * R + j] - hq[j], Dsum);
    }
    const float lse_t = s_lse[tt];
    float dh_tot = 0.f;                             // component `comp` of the dh contribution (writer lanes)
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      if (kb * 32 >= L) break;                      // uniform
      const int k = kb * 32 + lane;
      const float* Kr = Ks + k * RS;
      float Kv[R];
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) { Kv[j] = Kr[j]; sc = fmaf(hq[j], Kv[j], sc); dp = fmaf(dg[j], Kv[j], dp); }
      const float p = (k < L) ? exp2f(sc * sl - lse_t) : 0.f;
      dp *= d.attn_w;
      const float ds = p * (dp - Dsum);
      const float pw = p * d.attn_w, dss = ds * d.sm_scale;
      float part[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        dKacc[kb][j] = fmaf(pw, dg[j], fmaf(dss, hq[j], dKacc[kb][j]));
        part[j] = dss * Kv[j];
      }
#pragma unroll
      for (int half = R / 2, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send = upper ? part[i] : part[i + half];
          const float keep = upper ? part[i + half] : part[i];
          part[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      float tot = part[0];
      if (R <= 16) tot += __shfl_xor_sync(0xffffffffu, tot, 1);
      if (R <= 8) tot += __shfl_xor_sync(0xffffffffu, tot, 2);
      dh_tot += tot;
    }
    if (writer && comp < r) dh[(tbase + tt) * r + comp] = dgs[tt * R + comp] + dh_tot;   // in place: rows are ours
  }
  // combine the per-warp dK partials in shared memory, then one reduction per element
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    if (kb * 32 >= L) break;
    const int k = kb * 32 + lane;
#pragma unroll
    for (int j = 0; j < R; ++j) atomicAdd(&dKs[k * RS + j], dKacc[kb][j]);
  }
  __syncthreads();
  for (int i = tid; i < L * R; i += 256) {
    const int k = i / R, j = i - k * R;
    if (j < r) {
      const float v = dKs[k * RS + j];
      if (v != 0.f) red_add_f32(&dK_glob[((size_t)b * d.Lk + k) * r + j], v);
    }
  }
}

// First kernel of a backward call: B [d_out, r] -> B^T [rk, d_out] bf16 (rows r..rk zero), the K-major
// operand of dG = dY . B that k_dgrad_tc TMA-loads, plus the zero-fill of the small reduction targets
// (per-sample dK, arrival counters).  Replaces a cudaMemsetAsync + transpose pair; being a kernel it
// can chain with programmatic dependent launch.
__global__ void __launch_bounds__(256)
k_bwd_prep(const __nv_bfloat16* __restrict__ Bw, __nv_bfloat16* __restrict__ Bt, int d_out, int r, int rk,
           uint32_t* __restrict__ zero_ptr, int zero_words, int dbg) {
  __shared__ __nv_bfloat16 tile[64][136];                        // 64 columns x up to 128 ranks (+pad)
  if (threadIdx.x == 0) dbg_stamp_g(dbg, blockIdx.x, 106);
  ptx::pdl_wait();                                               // B (optimizer), workspace (earlier calls)
  ptx::pdl_trigger();
  if (threadIdx.x == 0) dbg_stamp_g(dbg, blockIdx.x, 99);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < zero_words; i += gridDim.x * 256) zero_ptr[i] = 0u;
  const int c0 = blockIdx.x * 64;
  for (int idx = threadIdx.x; idx < 64 * r; idx += 256) {
    const int c = idx / r, j = idx - c * r;
    tile[c][j] = (c0 + c < d_out) ? Bw[(size_t)(c0 + c) * r + j] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < rk * 64; idx += 256) {
    const int j = idx >> 6, c = idx & 63;
    if (c0 + c < d_out) Bt[(size_t)j * d_out + c0 + c] = (j < r) ? tile[c][j] : __float2bfloat16_rn(0.f);
  }
  if (threadIdx.x == 0) dbg_stamp_g(dbg, blockIdx.x, 107);
}

static int launch_attn_nopack(const DevDesc& d, const SimtArgs& a, cudaStream_t st);   // k_attn_pack without the packing

// fused = the merged attention-backward + pack kernel follows (partial dG buffers, nothing zero-filled
// but dK and the counters); otherwise the legacy kernels run on a complete dG in a.dh (reduced into the zeroed buffer).
static DgParts dgrad_parts(const DevDesc& d, int num_sms, bool fused, long long stride) {
  DgParts q;
  q.P = d.d_out / 128;
  q.U = (long long)((d.T + 127) / 128) * q.P;
  q.G = fused ? dgrad_grid(d.T, d.d_out, num_sms) : 0;
  q.stride = stride;
  return q;
}

static int launch_dgrad_tc(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, bool fused) {
  DgradTcParams p;
  p.T = d.T; p.d_out = d.d_out; p.r = d.r; p.rk = round_up(d.r, 16);
  p.pairs = d.d_out / 128;
  const int ntiles = (d.T + 127) / 128;
  p.units = (long long)ntiles * p.pairs;
  p.grid = dgrad_grid(d.T, d.d_out, num_sms);
  p.dg_mode = fused ? 0 : 2;
  p.part_stride = (long long)(a.dh_att - a.dh);
  p.bt_bytes = 2u * (uint32_t)p.rk * 128u;
  const size_t fixed = (size_t)4 * p.rk * 128 + 1024 + 512;
  // TMEM: [dG x ndg][dB sets: nbuf x wide x 2 rk]
  p.dg_cols = tmem_cols_for(p.rk);
  static int dg_wide_env = -1;
  if (dg_wide_env < 0) { const char* e = getenv("MOKA_DGRAD_WIDE"); dg_wide_env = e ? atoi(e) : 2; }
  const size_t limit = 226 * 1024;
  {
    // two pairs (256 columns) per stage when two sets of that many dB accumulators fit next to two dG accumulators
    // and the ring still gets three stages
    const size_t st2 = 2 * (32768u + (size_t)p.bt_bytes);
    p.wide = (dg_wide_env >= 2 && 2 * p.dg_cols + 8 * p.rk <= 512 && 3 * st2 + fixed <= limit) ? 2 : 1;
  }
  p.nbuf = (p.dg_cols + 4 * p.wide * p.rk <= 512) ? 2 : 1;
  p.ndg = (2 * p.dg_cols + p.nbuf * p.wide * 2 * p.rk <= 512) ? 2 : 1;
  p.d2_col = (uint32_t)p.ndg * p.dg_cols;
  p.tmem_cols = tmem_cols_for((int)p.d2_col + p.nbuf * p.wide * 2 * p.rk);
  p.stage_bytes = (uint32_t)p.wide * (32768u + p.bt_bytes);
  static int dg_stages_env = -1;
  if (dg_stages_env < 0) { const char* e = getenv("MOKA_DGRAD_STAGES"); dg_stages_env = e ? atoi(e) : 0; }
  p.stages = (int)((limit - fixed) / p.stage_bytes);
  if (p.stages > 6) p.stages = 6;
  if (dg_stages_env > 0 && p.stages > dg_stages_env) p.stages = dg_stages_env;
  if (p.stages < 2) return MOKA_EUNSUPPORTED;
  p.dbg = dbg_enabled();
  { static int nodb = -1; if (nodb < 0) { const char* e = getenv("MOKA_DBG_NO_DB"); nodb = (e && e[0] == '1') ? 2 : 0; } p.dbg |= nodb; }
  // timing experiments (results invalid): MOKA_DBG_DGRAD_SKIP=1 drops the dG product, =2 the dB product
  { static int sk = -1; if (sk < 0) { const char* e = getenv("MOKA_DBG_DGRAD_SKIP"); sk = e ? (atoi(e) & 3) << 2 : 0; } p.dbg |= sk; }
  // pre-transposed B in the (otherwise unused until the pack step) packed-operand workspace
  __nv_bfloat16* Bt = static_cast<__nv_bfloat16*>(a.pack);
  CUtensorMap tdy, tbt;
  if (!make_tmap(&tdy, a.dy, (uint64_t)d.T, (uint64_t)d.d_out, 128)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&tbt, Bt, (uint64_t)p.rk, (uint64_t)d.d_out, (uint32_t)p.rk)) return MOKA_EUNSUPPORTED;
  const size_t smem = (size_t)p.stages * p.stage_bytes + fixed;
  cudaError_t e = cudaFuncSetAttribute(k_dgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  if (!(d.dbg_skip & 4)) {
    // fused: dK [B, Lk, r] and the counters [B] sit next to each other in the workspace
    uint32_t* zp = fused ? reinterpret_cast<uint32_t*>(a.dK) : nullptr;
    const int zw = fused ? (int)((reinterpret_cast<char*>(a.counters) - reinterpret_cast<char*>(a.dK)) / 4) + d.B : 0;
    e = launch_pdl(k_bwd_prep, dim3((d.d_out + 63) / 64), dim3(256), 0, st, static_cast<const __nv_bfloat16*>(a.Bw), Bt,
                   d.d_out, d.r, p.rk, zp, zw, p.dbg);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    if (int rc = check_launch_tc()) return rc;
  }
  if (!(d.dbg_skip & 8)) {
    e = launch_pdl(k_dgrad_tc, dim3(p.grid), dim3(kDgradThreads), smem, st, tdy, tbt, d, p, a.tokmeta,
                   (const float*)(a.h_stash + (size_t)d.T * d.r), a.dh, a.dB, (uint32_t*)nullptr, 0);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    if (int rc = check_launch_tc()) return rc;
  }
  if (fused) return MOKA_OK;
  // not fused (the wgrad stage cannot use the packed operands): attention backward on the small [T, r]
  // tensors, in place in a.dh, dK left in a.dK for the consumer
  if (d.r <= 128 && a.klen_hint <= 128 && d.r % 4 == 0) {
    if (d.dbg_skip & 16) return MOKA_OK;
    return launch_attn_nopack(d, a, st);
  }
  if (d.r <= 32 && a.klen_hint <= 128 && (d.r <= 16 || a.klen_hint <= 64)) {
    const int kblocks = (max(a.klen_hint, 1) + 31) / 32;
    const int gps = (d.S + 31) / 32;
    const int grid = d.B * gps;
    const int R = d.r <= 8 ? 8 : (d.r <= 16 ? 16 : 32);
    const int KBs = kblocks <= 1 ? 1 : (kblocks <= 2 ? 2 : 4);
    const size_t sm = ((size_t)2 * KBs * 32 * (R + 1) + 3 * 32 * R) * 4;
#define MOKA_LAUNCH_ATTN_BWD(RR, KK)                                                                     \
    do {                                                                                                \
      if (sm > 40 * 1024) cudaFuncSetAttribute(k_attn_bwd_tile<RR, KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
      k_attn_bwd_tile<RR, KK><<<grid, 256, sm, st>>>(d, a.tokmeta, a.kidx, a.klen, a.h_stash, a.dh, a.dK, gps); \
    } while (0)
    if (R == 8) { if (KBs == 1) MOKA_LAUNCH_ATTN_BWD(8, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_BWD(8, 2); else MOKA_LAUNCH_ATTN_BWD(8, 4); }
    else if (R == 16) { if (KBs == 1) MOKA_LAUNCH_ATTN_BWD(16, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_BWD(16, 2); else MOKA_LAUNCH_ATTN_BWD(16, 4); }
    else { if (KBs == 1) MOKA_LAUNCH_ATTN_BWD(32, 1); else MOKA_LAUNCH_ATTN_BWD(32, 2); }
#undef MOKA_LAUNCH_ATTN_BWD
    return check_launch_tc();
  }
  DevDesc dd = d;
  dd.tile_tokens = 32;
  dd.num_tiles = (d.T + 31) / 32;
  int kcap = min(max(a.klen_hint, 1), 32 * kKeyBlocksMax);
  const size_t per_key = (size_t)2 * (d.r + 1) * 4;
  const size_t row_bytes = (size_t)2 * 32 * d.r * 4;
  while (kcap > 1 && row_bytes + kcap * per_key > 96 * 1024) kcap /= 2;
  const size_t smem2 = row_bytes + kcap * per_key + 64;
  if (smem2 > 40 * 1024) {
    e = cudaFuncSetAttribute(k_attn_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  }
  k_attn_bwd<<<min(dd.num_tiles, num_sms * 2), 256, smem2, st>>>(dd, a.tokmeta, a.kidx, a.klen, a.h_stash, a.dh,
                                                                  a.dK, kcap);
  return check_launch_tc();
}

// ---------------------------------------------------------------------------- k_pack_dh
// dh (+ the per-sample dK rows folded back onto their key tokens) -> the bf16 operand tensors
// k_wgrad_tc loads with TMA:
//   dHc  [T, KP]   row t is zero except the rank block of the token's modality (K-major operand
//                  of  dx = dHc . [A_0;A_1;A_2]),
//   dHcT [2, NPW, T]  its transpose, split into hi + lo bf16 so that the weight gradient
//                  dA = dHcT . X keeps fp32-level accuracy.
struct PackLayout { int rp, NPW, KP, lo_off; size_t hc_bytes, hct_bytes; };

// One 16-byte chunk (columns n0 .. n0 + 7) of a packed dHc row.  Columns [0, NPW) hold bf16(dh); when the padded K
// axis has room (2 NPW <= KP: every r <= 8, and r = 16 with two modalities) columns [lo_off, 2 lo_off) hold the bf16
// RESIDUAL of the same values and k_wgrad_tc runs its dx product over both halves against a duplicated weight tile:
// dx = dh A then sees dh at fp32-level accuracy.  It matters exactly where the adapter term is large next to
// W0^T dy - small ranks, whose alpha / r scale is big (r = 4: 4x the r = 16 term; its dx missed the 1e-3 bar by 20 %).
template <typename F>
__device__ __forceinline__ uint4 hc_chunk(F val, int base, int r, int n0, int lo_off, float scale) {
  float o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int col = n0 + e;
    const bool islo = lo_off > 0 && col >= lo_off;
    const int j = (islo ? col - lo_off : col) - base;
    const float x = (base >= 0 && j >= 0 && j < r) ? val(j) * scale : 0.f;
    const float hi = __bfloat162float(__float2bfloat16_rn(x));
    o[e] = islo ? x - hi : hi;
  }
  uint4 q;
  q.x = pack_bf16(o[0], o[1]); q.y = pack_bf16(o[2], o[3]); q.z = pack_bf16(o[4], o[5]); q.w = pack_bf16(o[6], o[7]);
  return q;
}

static PackLayout pack_layout(int T, int r, int M) {
  PackLayout L;
  L.rp = round_up(r, 8);
  L.NPW = round_up(M * L.rp, 16);
  L.KP = round_up(L.NPW, 64);
  L.lo_off = (2 * L.NPW <= L.KP) ? L.NPW : 0;
  L.hc_bytes = ((size_t)T * L.KP * 2 + 255) & ~size_t(255);
  L.hct_bytes = ((size_t)L.NPW * T * 2 + 255) & ~size_t(255);
  return L;
}
size_t tc_pack_bytes(int T, int r, int M) {
  PackLayout L = pack_layout(T, r, M);
  return L.hc_bytes + 2 * L.hct_bytes;                           // also hosts B^T [rk, d_out] (see moka_workspace_bytes)
}

__device__ __forceinline__ int find_key_slot(const int32_t* __restrict__ row, int L, int pos) {
  int lo = 0, hi = L - 1;                                     // key positions are strictly increasing
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = row[mid];
    if (v == pos) return mid;
    if (v < pos) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// One CTA per 64 tokens: rows are gathered into shared memory once (coalesced), then written as
// dHc rows and as dHcT columns (64 consecutive tokens = 128 contiguous bytes per packed row).
__global__ void __launch_bounds__(256)
k_pack_dh(DevDesc d, int rp, int NPW, int KP, int lo_off, const uint8_t* __restrict__ tokmeta,
          const int32_t* __restrict__ kidx, const int32_t* __restrict__ klen, const float* __restrict__ dh,
          const float* __restrict__ dh_att, const float* __restrict__ dK, __nv_bfloat16* __restrict__ hc,
          __nv_bfloat16* __restrict__ hct_hi, __nv_bfloat16* __restrict__ hct_lo, float scale) {
  extern __shared__ __align__(16) float vs[];                    // [64][r + 1]
  __shared__ int s_base[64];                                     // first packed row of the token's rank block, -1 = none
  const int r = d.r, tid = threadIdx.x;
  const int t0 = blockIdx.x * 64;
  const int nt = min(64, d.T - t0);
  if (tid < 64) {
    int base = -1;
    if (tid < nt) {
      const int code = tokmeta[t0 + tid] & MOKA_MOD_MASK;
      if (code) base = (code - 1) * rp;
    }
    s_base[tid] = base;
  }
  for (int i = tid; i < 64 * r; i += 256) {
    const int tt = i / r, j = i - tt * r;
    float v = 0.f;
    if (tt < nt) {
      const size_t t = (size_t)t0 + tt;
      v = dh[t * r + j] + dh_att[t * r + j];
    }
    vs[tt * (r + 1) + j] = v;
  }
  __syncthreads();
  // fold dK onto the key tokens of this tile
  if (tid < nt) {
    const int t = t0 + tid;
    const uint8_t mb = tokmeta[t];
    if ((mb & MOKA_KEY_VALID) && (mb & MOKA_MOD_MASK)) {
      const int b = t / d.S, pos = t - b * d.S;
      const int slot = find_key_slot(kidx + (size_t)b * d.Lk, klen[b], pos);
      if (slot >= 0)
        for (int j = 0; j < r; ++j) vs[tid * (r + 1) + j] += dK[((size_t)b * d.Lk + slot) * r + j];
    }
  }
  __syncthreads();
  // dHc rows
  const int cpr = KP / 8;
  if (hc)
  for (int i = tid; i < nt * cpr; i += 256) {
    const int tt = i / cpr, n0 = (i - tt * cpr) * 8;
    const int base = s_base[tt];
    const float* rowp = vs + tt * (r + 1);
    *reinterpret_cast<uint4*>(hc + ((size_t)t0 + tt) * KP + n0) =
        hc_chunk([&](int j) { return rowp[j]; }, base, r, n0, lo_off, scale);
  }
  // dHcT columns (hi + lo); T % 8 == 0 so 8-token chunks never straddle the end
  const int tch = (nt + 7) / 8;
  for (int i = tid; i < NPW * tch; i += 256) {
    const int n = i / tch, c = i - n * tch;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int tt = c * 8 + e;
      const int base = (tt < nt) ? s_base[tt] : -1;
      const int j = n - base;
      const float v = (base >= 0 && j >= 0 && j < r) ? vs[tt * (r + 1) + j] * scale : 0.f;
      hi[e] = __bfloat162float(__float2bfloat16_rn(v));
      lo[e] = v - hi[e];
    }
    uint4 o;
    o.x = pack_bf16(hi[0], hi[1]); o.y = pack_bf16(hi[2], hi[3]); o.z = pack_bf16(hi[4], hi[5]); o.w = pack_bf16(hi[6], hi[7]);
    *reinterpret_cast<uint4*>(hct_hi + (size_t)n * d.T + t0 + c * 8) = o;
    o.x = pack_bf16(lo[0], lo[1]); o.y = pack_bf16(lo[2], lo[3]); o.z = pack_bf16(lo[4], lo[5]); o.w = pack_bf16(lo[6], lo[7]);
    *reinterpret_cast<uint4*>(hct_lo + (size_t)n * d.T + t0 + c * 8) = o;
  }
}

// ---------------------------------------------------------------------------- k_attn_pack
// Attention backward + operand packing in one kernel (r <= 32, <= 128 keys per sample).
// A CTA owns 64 consecutive tokens of one sample (16 warps x 4 tokens):
//   1. dG rows (the two partial buffers of k_dgrad_tc summed) -> shared memory; for tiles with query
//      tokens the attention backward of k_attn_bwd_tile runs on them in place (dh), the per-sample dK
//      sums leave through one reduction per element per CTA;
//   2. tiles WITHOUT key tokens are packed straight from shared memory into the bf16 operands of
//      k_wgrad_tc (dHc, dHcT hi/lo) - dh never goes back to HBM as fp32;
//   3. tiles WITH key tokens need the finished dK of their sample: they park their dh rows in the dG
//      buffer, every CTA bumps the sample's arrival counter, and the CTA that arrives last packs the
//      parked tiles with dK folded onto the key rows (threadfence + counter: the classic last-block
//      pattern, no second launch and no grid-wide barrier).
struct AttnPackParams {
  int rp, NPW, KP, bps;             // bps = 32-token blocks per sample
  int lo_off;                       // residual half of the dHc rows (see hc_chunk), 0 = none
  float scale;
  int pack;                         // 1: emit the packed wgrad operands (fused chain); 0: only write dh back to dG0 and
                                    // reduce dK (the consumer - k_pack_dh or the warp-FMA wgrad - folds dK itself)
};

template <int R, int TT>
__device__ __forceinline__ void pack_rows(const DevDesc& d, const AttnPackParams& p, const float* __restrict__ vs,
                                          const int* __restrict__ s_base, size_t tbase, int nt,
                                          __nv_bfloat16* __restrict__ hc, __nv_bfloat16* __restrict__ hct_hi,
                                          __nv_bfloat16* __restrict__ hct_lo, int tid, int nthr) {
  constexpr int VS = R + 1;
  const int r = d.r;
  const int cpr = p.KP / 8;
  if (hc)
  for (int i = tid; i < nt * cpr; i += nthr) {
    const int tt = i / cpr, n0 = (i - tt * cpr) * 8;
    const int base = s_base[tt];
    const float* rowp = vs + tt * VS;
    *reinterpret_cast<uint4*>(hc + (tbase + tt) * p.KP + n0) =
        hc_chunk([&](int j) { return rowp[j]; }, base, r, n0, p.lo_off, p.scale);
  }
  const int tch = (nt + 7) / 8;                     // S % 8 == 0: 8-token chunks never straddle a sample
  for (int i = tid; i < p.NPW * tch; i += nthr) {
    const int n = i / tch, c = i - n * tch;
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int tt = c * 8 + e;
      const int base = (tt < nt) ? s_base[tt] : -1;
      const int j = n - base;
      const float v = (base >= 0 && j >= 0 && j < r) ? vs[tt * VS + j] * p.scale : 0.f;
      hi[e] = __bfloat162float(__float2bfloat16_rn(v));
      lo[e] = v - hi[e];
    }
    uint4 o;
    o.x = pack_bf16(hi[0], hi[1]); o.y = pack_bf16(hi[2], hi[3]); o.z = pack_bf16(hi[4], hi[5]); o.w = pack_bf16(hi[6], hi[7]);
    *reinterpret_cast<uint4*>(hct_hi + (size_t)n * d.T + tbase + c * 8) = o;
    o.x = pack_bf16(lo[0], lo[1]); o.y = pack_bf16(lo[2], lo[3]); o.z = pack_bf16(lo[4], lo[5]); o.w = pack_bf16(lo[6], lo[7]);
    *reinterpret_cast<uint4*>(hct_lo + (size_t)n * d.T + tbase + c * 8) = o;
  }
}

// R: padded rank (8/16/32); KB: 32-key blocks; NT threads; a CTA owns TT = 32 tokens, a warp TT/(NT/32).
//   phase A  warp <-> tokens, lane <-> key: scores from the stashed log-sum-exp, ds, the dh update of the
//            query rows (butterfly over keys), and the per-(token, key) factors p w and ds/sqrt(r) to smem;
//   phase B  thread <-> (key, 4 rank indices): dK[k] = sum_t (p w)[t,k] dg_t + (ds/sqrt r)[t,k] h_t as an
//            outer-product accumulation over the CTA's tokens - every dK element has one owner, so there is
//            no cross-warp reduction (shared-memory float atomics are CAS loops on this hardware).
// MMA = true: phases A and B on the warp-level tensor cores (mma.sync m16n8k16 bf16, hi/lo split operands, see
// k_attn_fwd_mma).  Phase A: warp <-> 16 tokens x all keys, in chunks of 32 keys:  S = H K^T and dP = dG K^T,
// P = exp2(S - lse), dS = P (w dP - D), dh += (dS/sqrt r) K with dS taken straight from the accumulator
// registers; the factors p w and dS/sqrt(r) go to shared memory as bf16 hi/lo, key-major.  Phase B: every warp
// owns [16 keys x 8 ranks] tiles of dK = (p w)^T dG + (dS/sqrt r)^T H over the CTA's 32 tokens.
template <int R, int KB, int NT, bool MMA = false>
__global__ void __launch_bounds__(NT)
k_attn_pack(DevDesc d, AttnPackParams p, const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
            const int32_t* __restrict__ klen, const float* __restrict__ h_stash, float* __restrict__ dG0,
            DgParts parts, float* __restrict__ dK_glob, int* __restrict__ counters,
            __nv_bfloat16* __restrict__ hc, __nv_bfloat16* __restrict__ hct_hi, __nv_bfloat16* __restrict__ hct_lo) {
  constexpr int TT = 32;
  constexpr int TPW = TT / (NT / 32);               // tokens per warp
  constexpr int RS = R + 1;                         // odd row stride of the pack staging rows
  constexpr int KS = R <= 32 ? R + 1 : R;           // key rows: odd stride for lane <-> key reads, aligned for lane <-> rank
  constexpr int KP32 = KB * 32;
  constexpr int R4 = R / 4;
  extern __shared__ __align__(16) float sm[];
  const int r = d.r;
  float* hs = sm;                                   // [TT][R]
  float* gs = hs + TT * R;
  float* dgs = gs + TT * R;                         // dG rows (kept: phase B needs the original dg)
  float* dhs = dgs + TT * R;                        // dh rows: dG + the attention-backward update of query rows
  float* pws = dhs + TT * R;                        // [TT][KP32]  p * w
  float* dss = pws + TT * KP32;                     // [TT][KP32]  ds / sqrt(r)
  float* Ks = dss + TT * KP32;                      // [KP32][KS]   (zero rows for invalid keys)
  float* vs = Ks + KP32 * KS;                       // [TT][RS] staging of the pack step
  __shared__ int s_code[TT];
  __shared__ int s_base[TT];
  __shared__ float s_lse[TT];
  __shared__ uint32_t s_blkmask[32];
  __shared__ int s_kidx[KP32];                      // the sample's key positions (finisher)
  __shared__ int s_last;
  __shared__ int s_np[TT];                          // partial dG buffers holding a share of this token's row
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / p.bps;
  const int s0 = (blockIdx.x % p.bps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const float* g_stash = h_stash + (size_t)d.T * r;
  const float* lse = g_stash + (size_t)d.T * r;
  const int L = klen[b];
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 110);
  ptx::pdl_trigger();                               // k_wgrad_tc: prologue + x / dx prefetch may start (see k_dgrad_tc)
  int myq = 0, mykey = 0;
  if (tid < 32) s_blkmask[tid] = 0u;
  if (tid < TT) {
    const uint8_t mb = (tid < nt) ? tokmeta[tbase + tid] : (uint8_t)0;
    const int code = mb & MOKA_MOD_MASK;
    s_code[tid] = code;
    s_base[tid] = code ? (code - 1) * p.rp : -1;
    s_lse[tid] = (code >= MOKA_MOD_VIS) ? lse[tbase + tid] : 0.f;
    s_np[tid] = dg_nparts(parts, (int)((tbase + tid) >> 7));
    myq = code >= MOKA_MOD_VIS;
    mykey = (mb & MOKA_KEY_VALID) && code;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  const int anykey = __syncthreads_or(mykey);
  if (anyq) {
    // h, g rows and the sample's keys come from the forward: older than this call
    for (int i = tid; i < TT * R4; i += NT) {
      const int tt = i / R4, j = (i - tt * R4) * 4;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), gv = hv;
      if (tt < nt && j < r) {
        hv = *reinterpret_cast<const float4*>(h_stash + (tbase + tt) * r + j);
        gv = *reinterpret_cast<const float4*>(g_stash + (tbase + tt) * r + j);
      }
      reinterpret_cast<float4*>(hs)[i] = hv;
      reinterpret_cast<float4*>(gs)[i] = gv;
    }
    for (int idx = tid; idx < KP32 * R4; idx += NT) {
      const int k = idx / R4, j = (idx - k * R4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < L && j < r) {
        const int pos = kidx[(size_t)b * d.Lk + k];
        const size_t tk = (size_t)b * d.S + pos;
        if (j == 0) {                                 // for the finisher: key positions + candidate-tile mask
          s_kidx[k] = pos;
          atomicOr(&s_blkmask[(pos >> 5) >> 5], 1u << ((pos >> 5) & 31));
        }
        if (tokmeta[tk] & MOKA_KEY_VALID) v = *reinterpret_cast<const float4*>(h_stash + tk * r + j);
      }
      float* dst = Ks + k * KS + j;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 114);
  ptx::pdl_wait();                                  // dG from k_dgrad_tc, zeroed dK / counters from k_bwd_prep
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 115);
  for (int i = tid; i < TT * R4; i += NT) {
    const int tt = i / R4, j = (i - tt * R4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt < nt && j < r) {
      const float* src = dG0 + (tbase + tt) * r + j;
      v = *reinterpret_cast<const float4*>(src);
      const int np = s_np[tt];                          // partials of the persistent k_dgrad_tc (fixed order: bit-reproducible)
      for (int q = 1; q < np; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(src + (size_t)q * (size_t)parts.stride);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
    }
    reinterpret_cast<float4*>(dgs)[i] = v;
    reinterpret_cast<float4*>(dhs)[i] = v;
  }
  __syncthreads();
  if constexpr (MMA) {
    if (anyq) {
      constexpr int R16 = R < 16 ? 16 : R;
      constexpr int LPp = KP32 + 8, HP = R16 + 8, TP = 40;
      __nv_bfloat16* Thi = reinterpret_cast<__nv_bfloat16*>(vs + TT * RS);   // [R16][LPp]  K^T (rank-major)
      __nv_bfloat16* Tlo = Thi + R16 * LPp;
      __nv_bfloat16* Hh = Tlo + R16 * LPp;                                   // [TT][HP] token rows of h / dG
      __nv_bfloat16* Hl = Hh + TT * HP;
      __nv_bfloat16* Gh = Hl + TT * HP;
      __nv_bfloat16* Gl = Gh + TT * HP;
      __nv_bfloat16* HTh = Gl + TT * HP;                                     // [R16][TP] the same, rank-major
      __nv_bfloat16* HTl = HTh + R16 * TP;
      __nv_bfloat16* GTh = HTl + R16 * TP;
      __nv_bfloat16* GTl = GTh + R16 * TP;
      __nv_bfloat16* PWh = GTl + R16 * TP;                                   // [KP32][TP] p w, key-major
      __nv_bfloat16* PWl = PWh + KP32 * TP;
      __nv_bfloat16* DSh = PWl + KP32 * TP;                                  // [KP32][TP] ds / sqrt(r)
      __nv_bfloat16* DSl = DSh + KP32 * TP;
      // ---- operand tiles from the fp32 rows staged above
      for (int i = tid; i < KP32 * R16; i += NT) {
        const int k = i / R16, j = i - k * R16;
        const float v = (j < R) ? Ks[k * KS + j] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        Thi[j * LPp + k] = hi;
        Tlo[j * LPp + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
      for (int i = tid; i < TT * R16; i += NT) {
        const int tt = i / R16, j = i - tt * R16;
        const float hv = (j < R) ? hs[tt * R + j] : 0.f, gv = (j < R) ? dgs[tt * R + j] : 0.f;
        const __nv_bfloat16 hh = __float2bfloat16_rn(hv), gh = __float2bfloat16_rn(gv);
        const __nv_bfloat16 hl = __float2bfloat16_rn(hv - __bfloat162float(hh));
        const __nv_bfloat16 gl = __float2bfloat16_rn(gv - __bfloat162float(gh));
        Hh[tt * HP + j] = hh; Hl[tt * HP + j] = hl; Gh[tt * HP + j] = gh; Gl[tt * HP + j] = gl;
        HTh[j * TP + tt] = hh; HTl[j * TP + tt] = hl; GTh[j * TP + tt] = gh; GTl[j * TP + tt] = gl;
      }
      __syncthreads();
      const int g = lane >> 2, t = lane & 3;
      auto ld32 = [](const __nv_bfloat16* q) { return *reinterpret_cast<const uint32_t*>(q); };
      auto pk16 = [](const __nv_bfloat16* a, const __nv_bfloat16* b2) {
        return (uint32_t)__bfloat16_as_ushort(*a) | ((uint32_t)__bfloat16_as_ushort(*b2) << 16);
      };
      if (warp < 2) {
        // ---- phase A: this warp's 16 tokens against all keys
        const int m0 = warp * 16;
        const int r0 = m0 + g, r1 = m0 + g + 8;
        const bool q0 = s_code[r0] >= MOKA_MOD_VIS, q1 = s_code[r1] >= MOKA_MOD_VIS;
        float Dl = 0.f;
        if (lane < 16) {
          const int row = m0 + lane;
          for (int j = 0; j < R; ++j) Dl = fmaf(dgs[row * R + j], gs[row * R + j] - hs[row * R + j], Dl);
        }
        const float D0 = __shfl_sync(0xffffffffu, Dl, g), D1 = __shfl_sync(0xffffffffu, Dl, g + 8);
        const float lse0 = s_lse[r0], lse1 = s_lse[r1];
        const float sl = d.sm_scale * kLog2e;
        constexpr int NO = R16 / 8;
        float O[NO][4];
#pragma unroll
        for (int n = 0; n < NO; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
#pragma unroll 1
        for (int c = 0; c < KB; ++c) {
          if (c * 32 >= L) break;
          float S[4][4], dP[4][4];
#pragma unroll
          for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) { S[n][e] = 0.f; dP[n][e] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < R16 / 16; ++ks) {
            uint32_t ah[4], al[4], gh4[4], gl4[4];
            const int c0 = ks * 16 + 2 * t;
            ah[0] = ld32(Hh + r0 * HP + c0); ah[1] = ld32(Hh + r1 * HP + c0);
            ah[2] = ld32(Hh + r0 * HP + c0 + 8); ah[3] = ld32(Hh + r1 * HP + c0 + 8);
            al[0] = ld32(Hl + r0 * HP + c0); al[1] = ld32(Hl + r1 * HP + c0);
            al[2] = ld32(Hl + r0 * HP + c0 + 8); al[3] = ld32(Hl + r1 * HP + c0 + 8);
            gh4[0] = ld32(Gh + r0 * HP + c0); gh4[1] = ld32(Gh + r1 * HP + c0);
            gh4[2] = ld32(Gh + r0 * HP + c0 + 8); gh4[3] = ld32(Gh + r1 * HP + c0 + 8);
            gl4[0] = ld32(Gl + r0 * HP + c0); gl4[1] = ld32(Gl + r1 * HP + c0);
            gl4[2] = ld32(Gl + r0 * HP + c0 + 8); gl4[3] = ld32(Gl + r1 * HP + c0 + 8);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int key = c * 32 + n * 8 + g;
              const uint32_t b0h = pk16(Thi + (c0) * LPp + key, Thi + (c0 + 1) * LPp + key);
              const uint32_t b1h = pk16(Thi + (c0 + 8) * LPp + key, Thi + (c0 + 9) * LPp + key);
              const uint32_t b0l = pk16(Tlo + (c0) * LPp + key, Tlo + (c0 + 1) * LPp + key);
              const uint32_t b1l = pk16(Tlo + (c0 + 8) * LPp + key, Tlo + (c0 + 9) * LPp + key);
              mma_bf16_16816(S[n], ah, b0h, b1h); mma_bf16_16816(S[n], al, b0h, b1h); mma_bf16_16816(S[n], ah, b0l, b1l);
              mma_bf16_16816(dP[n], gh4, b0h, b1h); mma_bf16_16816(dP[n], gl4, b0h, b1h); mma_bf16_16816(dP[n], gh4, b0l, b1l);
            }
          }
          // p, ds and the two factors; S is overwritten with ds / sqrt(r) (the A operand of the dh product)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = c * 32 + n * 8 + 2 * t + e;
              const bool ok = key < L;
              const float p0 = (ok && q0) ? exp2f(S[n][e] * sl - lse0) : 0.f;
              const float p1 = (ok && q1) ? exp2f(S[n][2 + e] * sl - lse1) : 0.f;
              const float ds0 = p0 * (dP[n][e] * d.attn_w - D0) * d.sm_scale;
              const float ds1 = p1 * (dP[n][2 + e] * d.attn_w - D1) * d.sm_scale;
              const float pw0 = p0 * d.attn_w, pw1 = p1 * d.attn_w;
              S[n][e] = ds0; S[n][2 + e] = ds1;
              __nv_bfloat16 hi = __float2bfloat16_rn(pw0);
              PWh[key * TP + r0] = hi; PWl[key * TP + r0] = __float2bfloat16_rn(pw0 - __bfloat162float(hi));
              hi = __float2bfloat16_rn(pw1);
              PWh[key * TP + r1] = hi; PWl[key * TP + r1] = __float2bfloat16_rn(pw1 - __bfloat162float(hi));
              hi = __float2bfloat16_rn(ds0);
              DSh[key * TP + r0] = hi; DSl[key * TP + r0] = __float2bfloat16_rn(ds0 - __bfloat162float(hi));
              hi = __float2bfloat16_rn(ds1);
              DSh[key * TP + r1] = hi; DSl[key * TP + r1] = __float2bfloat16_rn(ds1 - __bfloat162float(hi));
            }
          }
          // dh += (ds / sqrt r) K : two k-steps of 16 keys, A fragments straight from the registers
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            uint32_t dh4[4], dl4[4];
            split_bf16x2(S[2 * kk][0], S[2 * kk][1], dh4[0], dl4[0]);
            split_bf16x2(S[2 * kk][2], S[2 * kk][3], dh4[1], dl4[1]);
            split_bf16x2(S[2 * kk + 1][0], S[2 * kk + 1][1], dh4[2], dl4[2]);
            split_bf16x2(S[2 * kk + 1][2], S[2 * kk + 1][3], dh4[3], dl4[3]);
#pragma unroll
            for (int n = 0; n < NO; ++n) {
              const __nv_bfloat16* th = Thi + (n * 8 + g) * LPp + c * 32 + kk * 16 + 2 * t;
              const __nv_bfloat16* tl = Tlo + (n * 8 + g) * LPp + c * 32 + kk * 16 + 2 * t;
              const uint32_t b0h = ld32(th), b1h = ld32(th + 8), b0l = ld32(tl), b1l = ld32(tl + 8);
              mma_bf16_16816(O[n], dh4, b0h, b1h); mma_bf16_16816(O[n], dl4, b0h, b1h); mma_bf16_16816(O[n], dh4, b0l, b1l);
            }
          }
        }
#pragma unroll
        for (int n = 0; n < NO; ++n) {
          const int col = n * 8 + 2 * t;
          if (col < r) {
            if (q0) { dhs[r0 * R + col] += O[n][0]; dhs[r0 * R + col + 1] += O[n][1]; }
            if (q1) { dhs[r1 * R + col] += O[n][2]; dhs[r1 * R + col + 1] += O[n][3]; }
          }
        }
      }
      __syncthreads();
      if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 116);
      // ---- phase B: dK tiles [16 keys x 8 ranks], one owner each
      {
        constexpr int NRK = R16 / 8;
        const int ntk = (min(L, KP32) + 15) / 16;
        for (int tile = warp; tile < ntk * NRK; tile += NT / 32) {
          const int mk = tile / NRK, nr = tile - mk * NRK;
          float C[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const int c0 = ks * 16 + 2 * t;
            const int k0 = (mk * 16 + g) * TP, k1 = (mk * 16 + g + 8) * TP, nb = (nr * 8 + g) * TP;
            uint32_t ah[4], al[4];
            ah[0] = ld32(PWh + k0 + c0); ah[1] = ld32(PWh + k1 + c0); ah[2] = ld32(PWh + k0 + c0 + 8); ah[3] = ld32(PWh + k1 + c0 + 8);
            al[0] = ld32(PWl + k0 + c0); al[1] = ld32(PWl + k1 + c0); al[2] = ld32(PWl + k0 + c0 + 8); al[3] = ld32(PWl + k1 + c0 + 8);
            uint32_t b0h = ld32(GTh + nb + c0), b1h = ld32(GTh + nb + c0 + 8), b0l = ld32(GTl + nb + c0), b1l = ld32(GTl + nb + c0 + 8);
            mma_bf16_16816(C, ah, b0h, b1h); mma_bf16_16816(C, al, b0h, b1h); mma_bf16_16816(C, ah, b0l, b1l);
            ah[0] = ld32(DSh + k0 + c0); ah[1] = ld32(DSh + k1 + c0); ah[2] = ld32(DSh + k0 + c0 + 8); ah[3] = ld32(DSh + k1 + c0 + 8);
            al[0] = ld32(DSl + k0 + c0); al[1] = ld32(DSl + k1 + c0); al[2] = ld32(DSl + k0 + c0 + 8); al[3] = ld32(DSl + k1 + c0 + 8);
            b0h = ld32(HTh + nb + c0); b1h = ld32(HTh + nb + c0 + 8); b0l = ld32(HTl + nb + c0); b1l = ld32(HTl + nb + c0 + 8);
            mma_bf16_16816(C, ah, b0h, b1h); mma_bf16_16816(C, al, b0h, b1h); mma_bf16_16816(C, ah, b0l, b1l);
          }
          const int key0 = mk * 16 + g, key1 = key0 + 8, col = nr * 8 + 2 * t;
          if (col < r) {
            if (key0 < L && (C[0] != 0.f || C[1] != 0.f))
              red_add_v2_f32(&dK_glob[((size_t)b * d.Lk + key0) * r + col], C[0], C[1]);
            if (key1 < L && (C[2] != 0.f || C[3] != 0.f))
              red_add_v2_f32(&dK_glob[((size_t)b * d.Lk + key1) * r + col], C[2], C[3]);
          }
        }
      }
      __syncthreads();                                 // the operand tiles share nothing with `vs`, but keep the phases apart
    }
  }
  if (!MMA && anyq && R > 32) {
    // ---- phase A, rank-parallel form (R = 64 / 128): lane <-> R/32 consecutive rank indices, the dot
    // products over the rank are warp reductions, one key at a time (two in flight)
    if constexpr (R > 32) {
      constexpr int C = R / 32;
      const float sl = d.sm_scale * kLog2e;
#pragma unroll 1
      for (int ti = 0; ti < TPW; ++ti) {
        const int tt = warp * TPW + ti;
        if (s_code[tt] < MOKA_MOD_VIS) {              // warp-uniform
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) { pws[tt * KP32 + kb * 32 + lane] = 0.f; dss[tt * KP32 + kb * 32 + lane] = 0.f; }
          continue;
        }
        float hq[C], dg[C], dacc[C];
        float Dsum = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          hq[c] = hs[tt * R + lane * C + c];
          dg[c] = dgs[tt * R + lane * C + c];
          dacc[c] = 0.f;
          Dsum = fmaf(dg[c], gs[tt * R + lane * C + c] - hq[c], Dsum);
        }
        Dsum = warp_sum(Dsum);
        const float lse_t = s_lse[tt];
        for (int k0 = 0; k0 < KP32; k0 += 2) {
          if (k0 >= L) {                              // uniform: beyond the key list
            if (lane < 2) { pws[tt * KP32 + k0 + lane] = 0.f; dss[tt * KP32 + k0 + lane] = 0.f; }
            continue;
          }
          float K0[C], K1[C];
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            K0[c] = Ks[k0 * KS + lane * C + c];
            K1[c] = Ks[(k0 + 1) * KS + lane * C + c];          // KP32 is even: row k0 + 1 exists (zero beyond L)
            a0 = fmaf(hq[c], K0[c], a0); b0 = fmaf(dg[c], K0[c], b0);
            a1 = fmaf(hq[c], K1[c], a1); b1 = fmaf(dg[c], K1[c], b1);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o); b0 += __shfl_xor_sync(0xffffffffu, b0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o); b1 += __shfl_xor_sync(0xffffffffu, b1, o);
          }
          const float p0 = exp2f(a0 * sl - lse_t);
          const float p1 = (k0 + 1 < L) ? exp2f(a1 * sl - lse_t) : 0.f;
          const float ds0 = p0 * (b0 * d.attn_w - Dsum) * d.sm_scale;
          const float ds1 = p1 * (b1 * d.attn_w - Dsum) * d.sm_scale;
#pragma unroll
          for (int c = 0; c < C; ++c) dacc[c] = fmaf(ds0, K0[c], fmaf(ds1, K1[c], dacc[c]));
          if (lane == 0) {
            pws[tt * KP32 + k0] = p0 * d.attn_w; pws[tt * KP32 + k0 + 1] = p1 * d.attn_w;
            dss[tt * KP32 + k0] = ds0; dss[tt * KP32 + k0 + 1] = ds1;
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (lane * C + c < r) dhs[tt * R + lane * C + c] += dacc[c];
      }
    }
  }
  if (!MMA && anyq && R <= 32) {
    // ---- phase A
    const float sl = d.sm_scale * kLog2e;
    const int comp = (R == 32) ? lane : (R == 16 ? (lane >> 1) : (lane >> 2));
    const bool writer = (R == 32) || (R == 16 && !(lane & 1)) || (R == 8 && !(lane & 3));
#pragma unroll 1
    for (int ti = 0; ti < TPW; ++ti) {
      const int tt = warp * TPW + ti;
      if (s_code[tt] < MOKA_MOD_VIS) {                // warp-uniform: not a query, contributes nothing to dK
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) { pws[tt * KP32 + kb * 32 + lane] = 0.f; dss[tt * KP32 + kb * 32 + lane] = 0.f; }
        continue;
      }
      float hq[R], dg[R];
      float Dsum = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        hq[j] = hs[tt * R + j];
        dg[j] = dgs[tt * R + j];
        Dsum = fmaf(dg[j], gs[tt * R + j] - hq[j], Dsum);
      }
      const float lse_t = s_lse[tt];
      float part[R];                                  // this lane's keys' share of the dh update: sum_k ds_k K_k
#pragma unroll
      for (int j = 0; j < R; ++j) part[j] = 0.f;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        const int k = kb * 32 + lane;
        float pw = 0.f, dsv = 0.f;
        if (kb * 32 < L) {                            // uniform
          const float* Kr = Ks + k * KS;
          float Kv[R];
          float sc = 0.f, dp = 0.f;
#pragma unroll
          for (int j = 0; j < R; ++j) { Kv[j] = Kr[j]; sc = fmaf(hq[j], Kv[j], sc); dp = fmaf(dg[j], Kv[j], dp); }
          const float pr = (k < L) ? exp2f(sc * sl - lse_t) : 0.f;
          dp *= d.attn_w;
          const float ds = pr * (dp - Dsum);
          pw = pr * d.attn_w;
          dsv = ds * d.sm_scale;
#pragma unroll
          for (int j = 0; j < R; ++j) part[j] = fmaf(dsv, Kv[j], part[j]);
        }
        pws[tt * KP32 + k] = pw;
        dss[tt * KP32 + k] = dsv;
      }
      // one butterfly per token: lane <-> key partials -> lane <-> rank index
#pragma unroll
      for (int half = R / 2, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float send = upper ? part[i] : part[i + half];
          const float keep = upper ? part[i + half] : part[i];
          part[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
        }
      }
      float dh_tot = part[0];
      if (R <= 16) dh_tot += __shfl_xor_sync(0xffffffffu, dh_tot, 1);
      if (R <= 8) dh_tot += __shfl_xor_sync(0xffffffffu, dh_tot, 2);
      if (writer && comp < r) dhs[tt * R + comp] += dh_tot;
    }
  }
  if (!MMA && anyq) {
    __syncthreads();
    if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 116);
    // ---- phase B: dK[k, 4 jq..] = sum over the CTA's tokens, one owner per element
    for (int idx = tid; idx < L * R4; idx += NT) {
      const int k = idx / R4, j = (idx - k * R4) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int tt = 0; tt < TT; ++tt) {
        const float pw = pws[tt * KP32 + k], dsv = dss[tt * KP32 + k];
        const float4 dgv = *reinterpret_cast<const float4*>(dgs + tt * R + j);
        const float4 hv = *reinterpret_cast<const float4*>(hs + tt * R + j);
        acc.x = fmaf(pw, dgv.x, fmaf(dsv, hv.x, acc.x));
        acc.y = fmaf(pw, dgv.y, fmaf(dsv, hv.y, acc.y));
        acc.z = fmaf(pw, dgv.z, fmaf(dsv, hv.z, acc.z));
        acc.w = fmaf(pw, dgv.w, fmaf(dsv, hv.w, acc.w));
      }
      if (j < r && (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f || acc.w != 0.f))
        red_add_v4_f32(&dK_glob[((size_t)b * d.Lk + k) * r + j], acc.x, acc.y, acc.z, acc.w);
    }
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 117);
  if (!p.pack) {
    if (anyq)                                         // rows without queries are unchanged: dh == dG already
      for (int i = tid; i < nt * R; i += NT) {
        const int tt = i / R, j = i - tt * R;
        if (j < r) dG0[(tbase + tt) * r + j] = dhs[i];
      }
    return;
  }
  if (!anykey) {
    for (int i = tid; i < TT * R; i += NT) { const int tt = i / R, j = i - tt * R; vs[tt * RS + j] = dhs[i]; }
    __syncthreads();
    pack_rows<R, TT>(d, p, vs, s_base, tbase, nt, hc, hct_hi, hct_lo, tid, NT);
  } else {
    // parked: the finisher of this sample reads these rows back (L2) once dK is complete
    for (int i = tid; i < nt * R; i += NT) {
      const int tt = i / R, j = i - tt * R;
      if (j < r) dG0[(tbase + tt) * r + j] = dhs[i];
    }
  }
  // ---- arrival: the last CTA of the sample packs the parked tiles
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 118);
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&counters[b], 1) == p.bps - 1);
  __syncthreads();
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 119);
  if (!s_last) { if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 111); return; }
  __threadfence();
  // ---- finisher.  The key index row and the mask of candidate tiles (tiles holding a key position) were
  // built while staging the keys (CTAs without queries fetch them now); the finished dK rows and the
  // parked tiles' dh rows / token codes then arrive in ONE round of independent loads, and folding and
  // packing run out of shared memory, all tiles of a round at once.
  constexpr int kSmFloats = 4 * TT * R + 2 * TT * KP32 + KP32 * KS + TT * RS;
  constexpr int kPBfit = (kSmFloats - KP32 * R) / (TT * RS);
  constexpr int PB = kPBfit < 6 ? kPBfit : 6;       // parked tiles per round
  static_assert(PB >= 1, "finisher scratch does not fit");
  float* dKs = sm;                                  // [L][R]   (the whole dynamic shared memory is free again)
  float* vs_all = sm + KP32 * R;                    // [PB][TT][RS]
  __shared__ int s_codes[PB * TT];
  __shared__ int s_base_all[PB * TT];
  __shared__ int s_list[PB];
  __shared__ int s_parked[PB];
  __shared__ int s_nlist;
  if (!anyq) {
    for (int i = tid; i < L; i += NT) {
      const int pos = kidx[(size_t)b * d.Lk + i];
      s_kidx[i] = pos;
      if (pos >= 0) atomicOr(&s_blkmask[(pos >> 5) >> 5], 1u << ((pos >> 5) & 31));
    }
    __syncthreads();
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 120);
  int word = 0;                                      // candidate tiles are walked 32 at a time
  uint32_t bits = s_blkmask[0];
  bool first_round = true;
  while (true) {
    __syncthreads();                                 // previous round done with the scratch
    if (tid == 0) {
      int n = 0, w = word;
      uint32_t m = bits;
      const int nwords = (p.bps + 31) >> 5;              // mask words that can have bits at all
      while (n < PB) {
        if (m == 0u) { if (++w >= nwords) { w = 32; break; } m = s_blkmask[w]; continue; }
        const int bit = __ffs(m) - 1;
        m &= m - 1u;
        s_list[n] = w * 32 + bit;
        s_parked[n] = 0;
        ++n;
      }
      s_nlist = n;
      s_last = w;                                    // resume point (word) ...
      s_blkmask[w < 32 ? w : 31] = (w < 32) ? m : 0u;   // ... and the bits still to do in it
    }
    __syncthreads();
    const int nl = s_nlist;
    word = s_last;
    bits = (word < 32) ? s_blkmask[word] : 0u;
    if (nl == 0) break;
    if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 121);
    if (first_round) {
      for (int i = tid; i < L * R4; i += NT) {
        const int k = i / R4, j = (i - k * R4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < r) v = __ldcg(reinterpret_cast<const float4*>(&dK_glob[((size_t)b * d.Lk + k) * r + j]));
        *reinterpret_cast<float4*>(dKs + k * R + j) = v;
      }
      first_round = false;
    }
    for (int i = tid; i < nl * TT; i += NT) {
      const int q = i / TT, tt = i - q * TT;
      const int pos = s_list[q] * TT + tt;
      const int mb = (pos < d.S) ? tokmeta[(size_t)b * d.S + pos] : 0;
      const int code = mb & MOKA_MOD_MASK;
      s_codes[i] = mb;
      s_base_all[i] = code ? (code - 1) * p.rp : -1;
      if ((mb & MOKA_KEY_VALID) && code) s_parked[q] = 1;        // benign race: every writer stores 1
    }
    for (int i = tid; i < nl * TT * R4; i += NT) {
      const int q = i / (TT * R4), rem = i - q * TT * R4;
      const int tt = rem / R4, j = (rem - tt * R4) * 4;
      const int pos = s_list[q] * TT + tt;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pos < d.S && j < r) v = __ldcg(reinterpret_cast<const float4*>(&dG0[((size_t)b * d.S + pos) * r + j]));
      float* dst = vs_all + (q * TT + tt) * RS + j;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 122);
    // fold dK onto the key rows: thread <-> (parked row, 4 rank indices)
    for (int i = tid; i < nl * TT * R4; i += NT) {
      const int q = i / (TT * R4), rem = i - q * TT * R4;
      const int tt = rem / R4, j = (rem - tt * R4) * 4;
      const int mb = s_codes[q * TT + tt];
      if ((mb & MOKA_KEY_VALID) && (mb & MOKA_MOD_MASK) && j < r) {
        const int slot = find_key_slot(s_kidx, L, s_list[q] * TT + tt);
        if (slot >= 0) {
          float* dst = vs_all + (q * TT + tt) * RS + j;
          const float4 v = *reinterpret_cast<const float4*>(dKs + slot * R + j);
          dst[0] += v.x; dst[1] += v.y; dst[2] += v.z; dst[3] += v.w;
        }
      }
    }
    __syncthreads();
    if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 123);
    // pack every parked tile of the round (a candidate without a valid key was packed by its own CTA)
    const int cpr = p.KP / 8;
    if (hc)
    for (int i = tid; i < nl * TT * cpr; i += NT) {
      const int q = i / (TT * cpr), rem = i - q * TT * cpr;
      const int tt = rem / cpr, n0 = (rem - tt * cpr) * 8;
      const int pos = s_list[q] * TT + tt;
      if (!s_parked[q] || pos >= d.S) continue;
      const int base = s_base_all[q * TT + tt];
      const float* row = vs_all + (q * TT + tt) * RS;
      *reinterpret_cast<uint4*>(hc + ((size_t)b * d.S + pos) * p.KP + n0) =
          hc_chunk([&](int j) { return row[j]; }, base, r, n0, p.lo_off, p.scale);
    }
    constexpr int TCH = TT / 8;
    for (int i = tid; i < nl * p.NPW * TCH; i += NT) {
      const int q = i / (p.NPW * TCH), rem = i - q * p.NPW * TCH;
      const int n = rem / TCH, c = rem - n * TCH;
      const int pos0 = s_list[q] * TT + c * 8;
      if (!s_parked[q] || pos0 >= d.S) continue;      // S % 8 == 0: a chunk is inside the sample or outside
      float hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int tt = c * 8 + e;
        const int base = s_base_all[q * TT + tt];
        const int j = n - base;
        const float v = (base >= 0 && j >= 0 && j < r) ? vs_all[(q * TT + tt) * RS + j] * p.scale : 0.f;
        hi[e] = __bfloat162float(__float2bfloat16_rn(v));
        lo[e] = v - hi[e];
      }
      uint4 o;
      o.x = pack_bf16(hi[0], hi[1]); o.y = pack_bf16(hi[2], hi[3]); o.z = pack_bf16(hi[4], hi[5]); o.w = pack_bf16(hi[6], hi[7]);
      *reinterpret_cast<uint4*>(hct_hi + (size_t)n * d.T + (size_t)b * d.S + pos0) = o;
      o.x = pack_bf16(lo[0], lo[1]); o.y = pack_bf16(lo[2], lo[3]); o.z = pack_bf16(lo[4], lo[5]); o.w = pack_bf16(lo[6], lo[7]);
      *reinterpret_cast<uint4*>(hct_lo + (size_t)n * d.T + (size_t)b * d.S + pos0) = o;
    }
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 111);
}

// ---------------------------------------------------------------------------- k_wgrad_tc
// A CTA owns 128 columns of d_in and walks token tiles:
//   MMA-A  dAcat[128 cols, NPW] += X^T[128 cols, 128 tok] . dHc[128 tok, NPW]   (X tile MN-major,
//          accumulator resident in TMEM for the CTA's whole token range: one flush, no per-tile atomics)
//   MMA-X  dXadd[128 tok, 128 cols] = dHc[128 tok, KP] . Acat[KP, 128 cols]      (weights MN-major, loaded once)
// and adds dXadd onto the dx tile in shared memory before the TMA store (read-modify-write of dx).
struct WgradTcParams {
  int T, d_in, r, rp, M, NPW, KP;
  int lo_off;                      // dHc rows carry a residual half at this column (0: none): the dx product runs over both
  int tsplit, ntiles;
  int xs, ds, nop;                 // ring depths: X pairs / dX pairs (32 KB slots), operand buffers
  int has_dx, dbg;
  int delta;                       // MOKA_FLAG_DELTA_OUT: dx tiles are write-only in this pass (no dx read)
  int early_bits;                  // DROP: keep bits generated before the tile's waits (see drop_early_bits)
  int dbg_drop;                    // timing aid (MOKA_DBG_WGRAD_DROP; results NOT valid): 1 no generator, 2 no epilogue masking
  int mod_filter;                  // token tiles this pass visits: -1 all; 0 those holding a token of ANY modality; m > 0
                                   // those holding a token of modality code m (a per-modality pass of a rank-chunked call)
  int prefetch_dx;                 // dx tiles may be loaded before the dependency wait (false for the later passes of a
                                   // rank-chunked call: the pass in front is still writing dx)
  uint32_t w_box_bytes;            // one 64-column box of the weight slice: (NPW, or 2 NPW with the residual half) x 128 B
  uint32_t hct_blk;                // one [NPW rows x 64 tokens] block of the transposed operand
  int l2pf;                        // token tiles of L2 prefetch distance for the x / dx streams (0: off)
  int order;                       // CTA order, see the kernel
  int use_hc;                      // 0: dx = dHcT^T . A with dHcT (hi + lo) as an MN-major A operand - no dHc tile
                                   //    in shared memory (16 KB per operand buffer -> one more ring slot), no dHc tensor
                                   //    to pack, dh at fp32-level accuracy in the dx product.  1: the round-1 form
  int stack;                       // 1: hi and lo halves of the transposed operand stacked along N ([hi ; lo] per 64-token
                                   //    block, N = 2 NPW): ONE pass over the x tile per token tile yields both dA products
                                   //    (8 MMAs instead of 16; the flush adds the two accumulator halves); needs 2 NPW <= 256
  uint32_t op_bytes;               // operand buffer: dHc tile + dHcT hi/lo tiles
  uint32_t tmem_cols;
};

constexpr int kWgradThreads = 224;                              // + warp 6: TMA store issuer

// DROP: lora_dropout > 0.  dA needs u = keep (.) x and dx needs keep (.) (A^T dh): the workers mask every x tile
// in shared memory before the tensor core reads it (x_ready) and mask the dx accumulator in the epilogue; the
// 1/(1-p) factor is already in the packed dh operands.
constexpr int kWgradDropThreads = kWgradThreads + 256;          // + 8 warps that only apply the dropout mask
template <bool DROP>
__global__ void __launch_bounds__(DROP ? kWgradDropThreads : kWgradThreads)
k_wgrad_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dx,
           const __grid_constant__ CUtensorMap tm_a0, const __grid_constant__ CUtensorMap tm_a1,
           const __grid_constant__ CUtensorMap tm_a2, const __grid_constant__ CUtensorMap tm_hc,
           const __grid_constant__ CUtensorMap tm_hct_hi, const __grid_constant__ CUtensorMap tm_hct_lo,
           WgradTcParams p, DevDesc d, float* __restrict__ dA0, float* __restrict__ dA1, float* __restrict__ dA2,
           const uint8_t* __restrict__ tokmeta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t hct_blk = p.hct_blk;                         // [NPW rows x 64 tokens] block
  // (round 2: the token-major dHc tile is gone - the dx product reads the transposed operand dHcT, hi and lo, as an
  //  MN-major A operand; kbx = 0 K blocks of dHc in the operand buffers)
  const int kbx = p.use_hc ? p.KP / 64 : 0;
  uint8_t* xring = base;                                      // [xs][2 boxes][128 x 128 B]
  uint8_t* dring = xring + (size_t)p.xs * 32768;              // [ds][2 boxes][128 x 128 B]
  uint8_t* wtile = dring + (size_t)p.ds * 32768;              // [2 boxes][NPW x 128 B]  weights (MN-major B operand)
  uint8_t* op = wtile + 2 * p.w_box_bytes;                    // [nop] x { dHc [kbx][128x128B], dHcT hi [2][blk], lo [2][blk] }
  uint64_t* bars = reinterpret_cast<uint64_t*>(op + (size_t)p.nop * p.op_bytes);
  uint64_t* x_full = bars;                     // [xs]
  uint64_t* x_empty = x_full + p.xs;           // [xs]
  uint64_t* d_full = x_empty + p.xs;           // [ds]  dx tile landed
  uint64_t* d_empty = d_full + p.ds;           // [ds]  dx tile stored, slot reusable
  uint64_t* d_done = d_empty + p.ds;           // [ds]  dx tile updated by the workers
  uint64_t* op_full = d_done + p.ds;           // [2]
  uint64_t* op_empty = op_full + 2;            // [2]
  uint64_t* w_full = op_empty + 2;             // [1]
  uint64_t* t_full = w_full + 1;               // [2] dx accumulator ready
  uint64_t* t_empty = t_full + 2;              // [2]
  uint64_t* a_full = t_empty + 2;              // [1] dA accumulator final
  uint64_t* x_ready = a_full + 1;              // [xs] x tile masked (DROP only)
  uint64_t* km_free = x_ready + p.xs;          // [xs] keep bits of the slot consumed by the epilogue (DROP only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(km_free + p.xs);
  uint32_t* km_s = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + 512);   // [xs][128 rows][4 words] (DROP)

  const int warp = ptx::uniform_warp_idx(), lane = threadIdx.x & 31;
  // CTA order (MOKA_WGRAD_ORDER): 0 = token range fastest (round 1), 1 = column slice fastest: the CTAs in flight then
  // cover whole token rows (all column slices of a few token ranges) instead of a few 256-byte stripes of every row
  const int nslices = (int)gridDim.x / p.tsplit;
  const int slice = p.order ? (int)blockIdx.x % nslices : (int)blockIdx.x / p.tsplit;
  const int split = p.order ? (int)blockIdx.x / nslices : (int)blockIdx.x % p.tsplit;
  const int col0 = slice * 128;
  const int tt_beg = (int)((long long)split * p.ntiles / p.tsplit);
  const int tt_end = (int)((long long)(split + 1) * p.ntiles / p.tsplit);
  // The token tiles of the range that this pass has to visit.  A token belongs to one modality, so a per-modality
  // pass of a rank-chunked call (r = 64 with three modalities, r = 128) only changes dx rows / adds to dA from tiles
  // that hold a token of ITS modality: visiting just those turns "x and dx streamed once per modality" (3x) into
  // ~1.25x on the AVT layouts; tiles holding only padding are skipped by every pass.  The token metadata is older
  // than the call, so the list is built before the dependency wait.
  constexpr int kWgList = 256;
  __shared__ unsigned short s_tile[kWgList];
  __shared__ int s_nl;
  {
    const int nall = tt_end - tt_beg;
    const bool filter = p.mod_filter >= 0 && nall <= kWgList;
    if (filter) {
      for (int j = warp; j < nall; j += (int)(blockDim.x >> 5)) {
        const long long tb = (long long)(tt_beg + j) * 128 + lane * 4;
        bool hit = false;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int code = (tb + e < p.T) ? (tokmeta[tb + e] & MOKA_MOD_MASK) : 0;
          hit |= p.mod_filter ? (code == p.mod_filter) : (code != 0);
        }
        const unsigned any = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_tile[j] = any ? 1 : 0;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int n = 0;
      if (nall <= kWgList) {
        for (int j = 0; j < nall; ++j)
          if (!filter || s_tile[j]) s_tile[n++] = (unsigned short)j;
      } else {
        n = nall;                                           // (identity: the list is not used)
      }
      s_nl = n;
    }
    __syncthreads();
  }
  const int ntt = s_nl;
  const bool use_list = (tt_end - tt_beg) <= kWgList;
  auto tile_of = [&](int j) { return tt_beg + (use_list ? (int)s_tile[j] : j); };
  // (walking the token range in a per-slice rotated order was measured 4 % SLOWER: the column slices of a
  // range share their packed dh operand tiles through L2 when they walk it together)
  if (threadIdx.x == 0) { dbg_stamp_g(p.dbg, blockIdx.x, 112); dbg_stamp(p.dbg, blockIdx.x, 1); }
  if (threadIdx.x == 0) ptx::pdl_trigger();                   // last kernel of the backward chain (see k_up_tc)

  // weight rows that TMA never writes (padding of the rank blocks) must read as zero
  for (uint32_t i = threadIdx.x; i < 2 * p.w_box_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(wtile)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.xs; ++s) {
      ptx::mbar_init(&x_full[s], 1); ptx::mbar_init(&x_empty[s], 1);
      ptx::mbar_init(&x_ready[s], 2 * kWorkerWarps); ptx::mbar_init(&km_free[s], kWorkerWarps);
    }
    for (int s = 0; s < p.ds; ++s) {
      ptx::mbar_init(&d_full[s], 1); ptx::mbar_init(&d_empty[s], 1); ptx::mbar_init(&d_done[s], kWorkerWarps);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&op_full[i], 1); ptx::mbar_init(&op_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1); ptx::mbar_init(&t_empty[i], kWorkerWarps);
    }
    ptx::mbar_init(w_full, 1);
    ptx::mbar_init(a_full, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&tm_x); ptx::prefetch_tmap(&tm_dx); ptx::prefetch_tmap(&tm_hc);
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, p.tmem_cols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t d1_col = 256;                                // dx accumulators at [0,256), dA at [256, 256+NPW)

  if (warp == 0) {
    // ---- TMA producer (whole warp on warp-uniform values, one elected lane issues: see moka_ptx.cuh)
    {
      const uint32_t w_a = ptx::uniform_u32(ptx::smem_u32(wtile));
      const uint32_t x_a = ptx::uniform_u32(ptx::smem_u32(xring));
      const uint32_t d_a = ptx::uniform_u32(ptx::smem_u32(dring));
      const uint32_t op_a = ptx::uniform_u32(ptx::smem_u32(op));
      // weights of this column slice, once
      const int wcopies = p.lo_off ? 2 : 1;                    // second copy: rows lo_off.. (the residual half of dHc)
      ptx::mbar_arrive_expect_tx_w(w_full, (uint32_t)wcopies * 2u * (uint32_t)p.M * (uint32_t)p.r * 128u);
      for (int cp = 0; cp < wcopies; ++cp)
        for (int box = 0; box < 2; ++box) {
          const uint32_t wb = w_a + (uint32_t)box * p.w_box_bytes + (uint32_t)cp * (uint32_t)p.lo_off * 128u;
          ptx::tma_load_2d_w(wb, &tm_a0, w_full, col0 + box * 64, 0, ptx::kEvictLast);
          if (p.M > 1) ptx::tma_load_2d_w(wb + (uint32_t)p.rp * 128u, &tm_a1, w_full, col0 + box * 64, 0, ptx::kEvictLast);
          if (p.M > 2) ptx::tma_load_2d_w(wb + 2u * (uint32_t)p.rp * 128u, &tm_a2, w_full, col0 + box * 64, 0, ptx::kEvictLast);
        }
      const uint32_t op_tx = (uint32_t)kbx * 16384u + 4u * hct_blk;
      // x and dx are older than this call (the first kernel of the call waited for them): fill their
      // rings before the dependency wait; the packed dh operands come from the kernel just before
      const int npre_x = min(ntt, p.xs), npre_d = (p.has_dx && p.prefetch_dx) ? min(ntt, p.ds) : 0;
      for (int j = 0; j < npre_x; ++j) {
        const uint32_t xt = x_a + (uint32_t)j * 32768u;
        ptx::mbar_arrive_expect_tx_w(&x_full[j], 32768u);
        ptx::tma_load_2d_w(xt, &tm_x, &x_full[j], col0, tile_of(j) * 128, ptx::kEvictFirst);
        ptx::tma_load_2d_w(xt + 16384u, &tm_x, &x_full[j], col0 + 64, tile_of(j) * 128, ptx::kEvictFirst);
      }
      for (int j = 0; j < npre_d; ++j) {
        const uint32_t dt = d_a + (uint32_t)j * 32768u;
        if (p.delta) { ptx::mbar_arrive_w(&d_full[j]); continue; }   // write-only dx: a free slot is enough
        ptx::mbar_arrive_expect_tx_w(&d_full[j], 32768u);
        ptx::tma_load_2d_w(dt, &tm_dx, &d_full[j], col0, tile_of(j) * 128, ptx::kEvictFirst);
        ptx::tma_load_2d_w(dt + 16384u, &tm_dx, &d_full[j], col0 + 64, tile_of(j) * 128, ptx::kEvictFirst);
      }
      // L2 prefetch of the x / dx tiles the rings have no room for yet (see tma_prefetch_2d_w)
      auto prefetch_tile = [&](int jj) {
        if (jj >= ntt) return;
        const int tp = tile_of(jj) * 128;
        ptx::tma_prefetch_2d_w(&tm_x, col0, tp);
        ptx::tma_prefetch_2d_w(&tm_x, col0 + 64, tp);
        if (p.has_dx && !p.delta) { ptx::tma_prefetch_2d_w(&tm_dx, col0, tp); ptx::tma_prefetch_2d_w(&tm_dx, col0 + 64, tp); }
      };
      if (p.l2pf)
        for (int jj = min(npre_x, npre_d ? npre_d : npre_x); jj < p.xs + p.l2pf; ++jj) prefetch_tile(jj);
      ptx::pdl_wait();
      for (int j = 0; j < ntt; ++j) {
        const int t0 = tile_of(j) * 128;
        if (p.l2pf) prefetch_tile(j + p.xs + p.l2pf);
        if (j >= npre_x) {   // the big streams first: they have their own ring slots and must never wait on the MMAs
          const int s = j % p.xs;
          const uint32_t ph = (uint32_t)(j / p.xs) & 1u;
          ptx::mbar_wait_or_trap(&x_empty[s], ph ^ 1u, __LINE__);
          const uint32_t xt = x_a + (uint32_t)s * 32768u;
          ptx::mbar_arrive_expect_tx_w(&x_full[s], 32768u);
          ptx::tma_load_2d_w(xt, &tm_x, &x_full[s], col0, t0, ptx::kEvictFirst);
          ptx::tma_load_2d_w(xt + 16384u, &tm_x, &x_full[s], col0 + 64, t0, ptx::kEvictFirst);
        }
        if (p.has_dx && j >= npre_d) {
          const int s = j % p.ds;
          const uint32_t ph = (uint32_t)(j / p.ds) & 1u;
          ptx::mbar_wait_or_trap(&d_empty[s], ph ^ 1u, __LINE__);
          const uint32_t dt = d_a + (uint32_t)s * 32768u;
          if (p.delta) {
            ptx::mbar_arrive_w(&d_full[s]);
          } else {
            ptx::mbar_arrive_expect_tx_w(&d_full[s], 32768u);
            ptx::tma_load_2d_w(dt, &tm_dx, &d_full[s], col0, t0, ptx::kEvictFirst);
            ptx::tma_load_2d_w(dt + 16384u, &tm_dx, &d_full[s], col0 + 64, t0, ptx::kEvictFirst);
          }
        }
        {   // packed dh operands of tile j (L2-resident; nop buffers, recycled when their MMAs retire)
          const int ob = j % p.nop;
          const uint32_t hc_t = op_a + (uint32_t)ob * p.op_bytes;
          const uint32_t hct_t = hc_t + (uint32_t)kbx * 16384u;
          ptx::mbar_wait_or_trap(&op_empty[ob], ((uint32_t)(j / p.nop) & 1u) ^ 1u, __LINE__);
          ptx::mbar_arrive_expect_tx_w(&op_full[ob], op_tx);
          for (int kb = 0; kb < kbx; ++kb) ptx::tma_load_2d_w(hc_t + (uint32_t)kb * 16384u, &tm_hc, &op_full[ob], kb * 64, t0, ptx::kEvictLast);
          for (int kb = 0; kb < 2; ++kb) {
            // stacked: [hi ; lo] of token block kb are adjacent (one B operand of 2 NPW rows); else hi blocks, then lo blocks
            const uint32_t hi_off = p.stack ? (uint32_t)(2 * kb) * hct_blk : (uint32_t)kb * hct_blk;
            const uint32_t lo_off = p.stack ? (uint32_t)(2 * kb + 1) * hct_blk : (uint32_t)(2 + kb) * hct_blk;
            ptx::tma_load_2d_w(hct_t + hi_off, &tm_hct_hi, &op_full[ob], t0 + kb * 64, 0, ptx::kEvictLast);
            ptx::tma_load_2d_w(hct_t + lo_off, &tm_hct_lo, &op_full[ob], t0 + kb * 64, 0, ptx::kEvictLast);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (whole warp, one elected lane issues)
    {
      const uint32_t tb = ptx::uniform_u32(tmem_base);
      const uint32_t idesc_a = ptx::idesc_bf16(128, (uint32_t)p.NPW, 1, 0);   // X^T (MN-major) . dHcT (K-major)
      const uint32_t idesc_a2 = ptx::idesc_bf16(128, 2u * (uint32_t)p.NPW, 1, 0);   // ... . [dHcT hi ; dHcT lo]
      const uint32_t idesc_x = ptx::idesc_bf16(128, 128, 0, 1);               // dHc (K-major) . Acat (MN-major)
      const uint32_t idesc_x2 = ptx::idesc_bf16(128, 128, 1, 1);              // dHcT^T (MN-major) . Acat (MN-major)
      const uint32_t wa = ptx::uniform_u32(ptx::smem_u32(wtile));
      const uint32_t x_a = ptx::uniform_u32(ptx::smem_u32(xring));
      const uint32_t op_a = ptx::uniform_u32(ptx::smem_u32(op));
      const uint64_t d_k = ptx::smem_desc_sw128(0, 1024, 0);                  // K-major operands
      const uint64_t d_w = ptx::smem_desc_sw128(0, 1024, p.w_box_bytes);      // weights, MN-major (two boxes)
      const uint64_t d_x = ptx::smem_desc_sw128(0, 1024, 16384);              // X tile, MN-major (two boxes)
      ptx::mbar_wait_or_trap(w_full, 0, __LINE__);
      for (int j = 0; j < ntt; ++j) {
        const int s = j % p.xs;
        const uint32_t ph = (uint32_t)(j / p.xs) & 1u;
        const int i = j & 1;
        const uint32_t ph2 = (uint32_t)(j >> 1) & 1u;
        const int ob = j % p.nop;
        const uint32_t hca = op_a + (uint32_t)ob * p.op_bytes;
        const uint32_t hcta = hca + (uint32_t)kbx * 16384u;
        ptx::mbar_wait_or_trap(&op_full[ob], (uint32_t)(j / p.nop) & 1u, __LINE__);
        if (lane == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 40 + j * 3);
        ptx::mbar_wait_or_trap(&x_full[s], ph, __LINE__);
        if (DROP) ptx::mbar_wait_or_trap(&x_ready[s], ph, __LINE__);
        if (lane == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 41 + j * 3);
        if (p.has_dx) ptx::mbar_wait_or_trap(&t_empty[i], ph2 ^ 1u, __LINE__);
        if (lane == 0 && j < 8) dbg_stamp(p.dbg, blockIdx.x, 42 + j * 3);
        ptx::tc_fence_after_sync();
        const uint32_t xa = x_a + (uint32_t)s * 32768u;
        if (p.has_dx && !p.use_hc) {
          // MMA-X first (the epilogue is waiting for it): dXadd[128 tok, 128 cols] = dHcT^T . Acat, the transposed
          // operand read as an MN-major A (two 64-token atoms), hi then lo; K = NPW rank rows in steps of 16
          const int ksteps = p.NPW / 16;
          const uint32_t lbo = (p.stack ? 2u : 1u) * hct_blk;                  // distance of the two 64-token atoms
          const uint64_t d_ha = ptx::smem_desc_sw128(0, 1024, lbo);
          for (int part = 0; part < 2; ++part) {
            const uint32_t part_off = p.stack ? (uint32_t)part * hct_blk : (uint32_t)part * 2u * hct_blk;
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adesc = d_ha + (uint64_t)((hcta + part_off + (uint32_t)k * 2048u) >> 4);
              const uint64_t bdesc = d_w + (uint64_t)((wa + (uint32_t)k * 2048u) >> 4);
              ptx::umma_bf16_w(tb + (uint32_t)i * 128u, adesc, bdesc, idesc_x2, (part > 0 || k > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit_w(&t_full[i]);
        } else if (p.has_dx) {
          // (round-1 form: token-major dHc tile as a K-major A operand)
          const int ksteps = (p.lo_off ? 2 * p.NPW : p.NPW) / 16;   // (+ the residual half of the dHc rows)
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc = d_k + (uint64_t)((hca + (uint32_t)(k >> 2) * 16384u + (k & 3) * 32) >> 4);
            const uint64_t bdesc = d_w + (uint64_t)((wa + (uint32_t)k * 2048u) >> 4);
            ptx::umma_bf16_w(tb + (uint32_t)i * 128u, adesc, bdesc, idesc_x, k > 0 ? 1u : 0u);
          }
          ptx::umma_commit_w(&t_full[i]);
        }
        // MMA-A: K = 128 tokens (8 steps)
        if (p.stack) {
          // hi and lo stacked along N: one pass over the x tile (the A operand fetch, 4 KB per MMA, is what an MMA costs)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t adesc = d_x + (uint64_t)((xa + k * 2048) >> 4);
            const uint64_t bdesc = d_k + (uint64_t)((hcta + (uint32_t)(2 * (k >> 2)) * hct_blk + (k & 3) * 32) >> 4);
            ptx::umma_bf16_w(tb + d1_col, adesc, bdesc, idesc_a2, (j > 0 || k > 0) ? 1u : 0u);
          }
        } else {
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t adesc = d_x + (uint64_t)((xa + k * 2048) >> 4);
            const uint64_t bdesc = d_k + (uint64_t)((hcta + (uint32_t)(2 * part + (k >> 2)) * hct_blk + (k & 3) * 32) >> 4);
            ptx::umma_bf16_w(tb + d1_col, adesc, bdesc, idesc_a, (j > 0 || part > 0 || k > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit_w(&x_empty[s]);
        ptx::umma_commit_w(&op_empty[ob]);
      }
      ptx::umma_commit_w(a_full);
    }
  } else if (warp >= 7) {
    // ================= dropout masking warps (DROP only): thread <-> (token row, 64-column box of the x tile).
    // Mask the tile in place before the tensor core reads it and leave the keep bits for the dx epilogue.
    if constexpr (DROP) {
      const int mt = threadIdx.x - 224;                        // 0 .. 255
      const int mrow = mt & 127, box = mt >> 7;
      for (int j = 0; j < ntt; ++j) {
        const int s = j % p.xs;
        const uint32_t k = (uint32_t)(j / p.xs);
        // keep bits of this thread's two 32-column super-groups: data-independent (early_bits: generated before the waits)
        const unsigned long long rowbase = dropout_row_base(d, (long long)tile_of(j) * 128 + mrow);
        uint32_t k32v[2] = {0xffffffffu, 0xffffffffu};
        if (p.early_bits && !(p.dbg_drop & 1)) {
#pragma unroll
          for (int hs = 0; hs < 2; ++hs)
            k32v[hs] = dropout_keep32(d, rowbase + (unsigned long long)((col0 >> 5) + box * 2 + hs));
        }
        ptx::mbar_sleep_or_trap(&x_full[s], k & 1u, __LINE__);
        if (p.has_dx && k > 0) ptx::mbar_sleep_or_trap(&km_free[s], (k - 1u) & 1u, __LINE__);
        if (!p.early_bits && !(p.dbg_drop & 1)) {
#pragma unroll
          for (int hs = 0; hs < 2; ++hs)
            k32v[hs] = dropout_keep32(d, rowbase + (unsigned long long)((col0 >> 5) + box * 2 + hs));
        }
        const uint32_t boxaddr = ptx::smem_u32(xring + (size_t)s * 32768) + (uint32_t)box * 16384u;
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {                         // a box = two 32-column super-groups
          km_s[((size_t)s * 128 + mrow) * 4 + box * 2 + hs] = k32v[hs];
          dropout_zero_dropped(k32v[hs], boxaddr + (uint32_t)mrow * 128u, (uint32_t)mrow & 7u, (uint32_t)hs);
        }
        ptx::fence_proxy_async_smem();
        warp_arrive(&x_ready[s]);
      }
    }
  } else if (warp == 6) {
    // TMA store issuer for the updated dx tiles
    if (lane == 0 && p.has_dx) {
      for (int j = 0; j < ntt; ++j) {
        const int s = j % p.ds;
        const uint32_t ph = (uint32_t)(j / p.ds) & 1u;
        const int t0 = tile_of(j) * 128;
        ptx::mbar_wait_or_trap(&d_done[s], ph, __LINE__);
        uint8_t* dt = dring + (size_t)s * 32768;
        ptx::tma_store_2d(&tm_dx, dt, col0, t0);
        ptx::tma_store_2d(&tm_dx, dt + 16384, col0 + 64, t0);
        ptx::tma_store_commit();
        ptx::tma_store_wait_read<0>();                         // smem read out: the slot can take the next load now
        ptx::mbar_arrive(&d_empty[s]);
      }
      ptx::tma_store_wait_all<0>();
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t km_cur[4] = {0, 0, 0, 0};                           // keep bits of this row in the current tile (DROP)
    if (p.has_dx) {
      for (int j = 0; j < ntt; ++j) {
        const int i = j & 1;
        const uint32_t ph2 = (uint32_t)(j >> 1) & 1u;
        const int s = j % p.ds;
        const uint32_t dph = (uint32_t)(j / p.ds) & 1u;
        const bool st = (threadIdx.x == 64) && j < 8;
        if (st && j == 0) dbg_stamp(p.dbg, blockIdx.x, 0);
        ptx::mbar_wait_or_trap(&t_full[i], ph2, __LINE__);
        if (DROP) {
          // the keep bits of tile j were left in shared memory by the masking warps (published by x_ready)
          const int sx = j % p.xs;
          ptx::mbar_wait_or_trap(&x_ready[sx], (uint32_t)(j / p.xs) & 1u, __LINE__);
          const uint4 kv = ptx::lds128(ptx::smem_u32(km_s + ((size_t)sx * 128 + row) * 4));
          km_cur[0] = kv.x; km_cur[1] = kv.y; km_cur[2] = kv.z; km_cur[3] = kv.w;
          warp_arrive(&km_free[sx]);
        }
        if (st) dbg_stamp(p.dbg, blockIdx.x, 8 + j * 4);
        ptx::mbar_wait_or_trap(&d_full[s], dph, __LINE__);
        if (st) dbg_stamp(p.dbg, blockIdx.x, 9 + j * 4);
        ptx::tc_fence_after_sync();
        uint8_t* dt = dring + (size_t)s * 32768;
        if (p.delta) {
          if (DROP)
            add_acc_to_tile<true, true>(lane_base + (uint32_t)(i * 128), ptx::smem_u32(dt), (uint32_t)row, &t_empty[i], km_cur);
          else
            add_acc_to_tile<false, true>(lane_base + (uint32_t)(i * 128), ptx::smem_u32(dt), (uint32_t)row, &t_empty[i]);
        } else if (DROP && !(p.dbg_drop & 2))
          add_acc_to_tile<true>(lane_base + (uint32_t)(i * 128), ptx::smem_u32(dt), (uint32_t)row, &t_empty[i], km_cur);
        else
          add_acc_to_tile(lane_base + (uint32_t)(i * 128), ptx::smem_u32(dt), (uint32_t)row, &t_empty[i]);
        if (st) dbg_stamp(p.dbg, blockIdx.x, 10 + j * 4);
        ptx::fence_proxy_async_smem();
        warp_arrive(&d_done[s]);
        if (st) dbg_stamp(p.dbg, blockIdx.x, 11 + j * 4);
      }
    }
    if (ntt > 0) {                                           // (no tile visited: the accumulator was never written)
    // ---- dA of this column slice: TMEM lane <-> column, TMEM column <-> (modality, rank index)
    if (threadIdx.x == 64) dbg_stamp(p.dbg, blockIdx.x, 4);
    ptx::mbar_wait_or_trap(a_full, 0, __LINE__);
    if (threadIdx.x == 64) dbg_stamp(p.dbg, blockIdx.x, 5);
    ptx::tc_fence_after_sync();
    // stage the accumulator as [NPW][128 cols] fp32 in the (now idle) X ring so that each thread can
    // reduce four consecutive columns of one (modality, rank) row with a single red.v4
    float* stage = reinterpret_cast<float*>(xring);
    // NPW x 512 bytes can be more than the X ring (one 32 KB slot at r = 128) and then runs into the dx ring,
    // whose last tiles the TMA store engine may still be reading: wait until those slots have been read out
    // (d_empty is what the store warp signals after cp.async.bulk.wait_group.read)
    if (p.has_dx && (size_t)p.NPW * 512 > (size_t)p.xs * 32768)
      for (int j = max(0, ntt - p.ds); j < ntt; ++j)
        ptx::mbar_wait_or_trap(&d_empty[j % p.ds], (uint32_t)(j / p.ds) & 1u, __LINE__);
    for (int n8 = 0; n8 < p.NPW; n8 += 8) {
      float v[8];
      ptx::tmem_ld8(lane_base + d1_col + (uint32_t)n8, v);
      if (p.stack) {                                             // + the lo half of the stacked accumulator
        float w[8];
        ptx::tmem_ld8(lane_base + d1_col + (uint32_t)(p.NPW + n8), w);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += w[e];
      } else {
        ptx::tmem_ld_wait();
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) stage[(n8 + e) * 128 + row] = v[e];
    }
    named_sync(2, kWorkers);
    const int wtid = threadIdx.x - 64;
    for (int idx = wtid; idx < p.NPW * 32; idx += kWorkers) {
      const int n = idx >> 5, c4 = (idx & 31) * 4;
      const int m = n / p.rp, jj = n - m * p.rp;
      if (m < p.M && jj < p.r) {
        float* dAm = m == 0 ? dA0 : (m == 1 ? dA1 : dA2);
        const float4 v = *reinterpret_cast<const float4*>(stage + n * 128 + c4);
        red_add_v4_f32(dAm + (size_t)jj * p.d_in + col0 + c4, v.x, v.y, v.z, v.w);
      }
    }
    }
    if (threadIdx.x == 64) dbg_stamp(p.dbg, blockIdx.x, 6);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) { dbg_stamp(p.dbg, blockIdx.x, 7); dbg_stamp_g(p.dbg, blockIdx.x, 113); }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// Tensor map over a row-major bf16 matrix with an explicit box (cols multiple of 8 elements).
// Can the merged attention-backward + pack kernel (k_attn_pack) serve this call?
static bool wgrad_plan(const DevDesc& d, bool has_dx, int num_sms, WgradTcParams& p, PackLayout& L, size_t& fixed,
                       int& nchunks);

bool tc_backward_fused_ok(const DevDesc& d, int klen_hint, bool has_dx) {
  WgradTcParams p;
  PackLayout L;
  size_t fixed;
  int nchunks;
  return d.r <= 128 && klen_hint <= 128 && d.S % 8 == 0 && d.S <= 32768 &&
         wgrad_plan(d, has_dx, 148, p, L, fixed, nchunks);
}

}  // namespace moka
#include "moka_attn_bwd2.cuh"
namespace moka {

static int launch_attn_pack(const DevDesc& d, const SimtArgs& a, const PackLayout& L, const DgParts& parts,
                            __nv_bfloat16* hc, __nv_bfloat16* hct_hi, __nv_bfloat16* hct_lo, cudaStream_t st,
                            int pack = 1) {
  AttnPackParams p;
  p.pack = pack;
  p.rp = L.rp; p.NPW = L.NPW; p.KP = L.KP; p.lo_off = L.lo_off; p.bps = (d.S + 31) / 32;
  p.scale = d.pre * d.drop_scale;
  const int kblocks = (max(a.klen_hint, 1) + 31) / 32;
  const int R = d.r <= 8 ? 8 : (d.r <= 16 ? 16 : (d.r <= 32 ? 32 : (d.r <= 64 ? 64 : 128)));
  const int KBs = kblocks <= 1 ? 1 : (kblocks <= 2 ? 2 : 4);
  const int KS = R <= 32 ? R + 1 : R;
  const size_t sm = ((size_t)4 * 32 * R + 2 * 32 * KBs * 32 + (size_t)KBs * 32 * KS + 32 * (R + 1)) * 4;
  cudaError_t e = cudaSuccess;
  static int attn_simt = -1;
  if (attn_simt < 0) { const char* ev = getenv("MOKA_ATTN_SIMT"); attn_simt = (ev && ev[0] == '1') ? 1 : 0; }
  static int attn_v1 = -1;                         // MOKA_ATTN_BWD_V1=1: the first-generation k_attn_pack (A/B measurements)
  if (attn_v1 < 0) { const char* ev = getenv("MOKA_ATTN_BWD_V1"); attn_v1 = (ev && ev[0] == '1') ? 1 : 0; }
  if (!attn_simt && !attn_v1 && pack && R <= 64) {
    // second generation (moka_attn_bwd2.cuh): 64 / 128 tokens per CTA, every warp on the tensor cores, dK straight
    // from transposed accumulators, key tiles finished by dedicated CTAs
    const int R16 = R < 16 ? 16 : R;
    int num_sms = 148;
    { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    int TT = 128;
    if (R16 == 64 || d.B * ((d.S + 127) / 128) < num_sms - num_sms / 8) TT = 64;
    const size_t smm = attn_bwd2_smem(R16, KBs, TT);
    if (smm <= 220 * 1024) {
      p.bps = (d.S + TT - 1) / TT;
      const int nq = d.B * p.bps;
#define MOKA_LAUNCH_ATTN_BWD2(RR, KK, TTT)                                                                   \
      do {                                                                                                   \
        cudaFuncSetAttribute(k_attn_bwd2<RR, KK, TTT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smm); \
        e = launch_pdl(k_attn_bwd2<RR, KK, TTT>, dim3(2 * nq), dim3(256), smm, st, d, p, nq, a.tokmeta, a.kidx, \
                       a.klen, (const float*)a.h_stash, a.dh, parts, a.dK, a.counters, hc, hct_hi, hct_lo);   \
      } while (0)
#define MOKA_LAUNCH_ATTN_BWD2_K(RR, TTT)                                                                     \
      do {                                                                                                   \
        if (KBs == 1) MOKA_LAUNCH_ATTN_BWD2(RR, 1, TTT); else if (KBs == 2) MOKA_LAUNCH_ATTN_BWD2(RR, 2, TTT); \
        else MOKA_LAUNCH_ATTN_BWD2(RR, 4, TTT);                                                              \
      } while (0)
      if (R16 == 16) { if (TT == 128) MOKA_LAUNCH_ATTN_BWD2_K(16, 128); else MOKA_LAUNCH_ATTN_BWD2_K(16, 64); }
      else if (R16 == 32) { if (TT == 128) MOKA_LAUNCH_ATTN_BWD2_K(32, 128); else MOKA_LAUNCH_ATTN_BWD2_K(32, 64); }
      else MOKA_LAUNCH_ATTN_BWD2_K(64, 64);
#undef MOKA_LAUNCH_ATTN_BWD2_K
#undef MOKA_LAUNCH_ATTN_BWD2
      if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      return check_launch_tc();
    }
  }
  if (!attn_simt && R <= 64) {
    // phases A / B on the warp-level tensor cores: bf16 hi/lo operand tiles behind the fp32 staging arrays
    const int R16 = R < 16 ? 16 : R, KP32 = KBs * 32;
    const size_t extra = ((size_t)2 * R16 * (KP32 + 8) + (size_t)4 * 32 * (R16 + 8) + (size_t)4 * R16 * 40 +
                          (size_t)4 * KP32 * 40) * 2;
    const size_t smm = sm + extra;
#define MOKA_LAUNCH_ATTN_PACK_MMA(RR, KK)                                                                    \
    do {                                                                                                     \
      cudaFuncSetAttribute(k_attn_pack<RR, KK, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smm); \
      e = launch_pdl(k_attn_pack<RR, KK, 256, true>, dim3(d.B * p.bps), dim3(256), smm, st, d, p, a.tokmeta, a.kidx, \
                     a.klen, (const float*)a.h_stash, a.dh, parts, a.dK, a.counters, hc, hct_hi, hct_lo);     \
    } while (0)
    if (R == 8) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK_MMA(8, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK_MMA(8, 2); else MOKA_LAUNCH_ATTN_PACK_MMA(8, 4); }
    else if (R == 16) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK_MMA(16, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK_MMA(16, 2); else MOKA_LAUNCH_ATTN_PACK_MMA(16, 4); }
    else if (R == 32) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK_MMA(32, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK_MMA(32, 2); else MOKA_LAUNCH_ATTN_PACK_MMA(32, 4); }
    else { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK_MMA(64, 1); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK_MMA(64, 2); else MOKA_LAUNCH_ATTN_PACK_MMA(64, 4); }
#undef MOKA_LAUNCH_ATTN_PACK_MMA
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    return check_launch_tc();
  }
#define MOKA_LAUNCH_ATTN_PACK(RR, KK, NN)                                                                    \
  do {                                                                                                       \
    if (sm > 40 * 1024)                                                                                      \
      cudaFuncSetAttribute(k_attn_pack<RR, KK, NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);    \
    e = launch_pdl(k_attn_pack<RR, KK, NN>, dim3(d.B * p.bps), dim3(NN), sm, st, d, p, a.tokmeta, a.kidx, a.klen, \
                   (const float*)a.h_stash, a.dh, parts, a.dK, a.counters, hc, hct_hi, hct_lo);                \
  } while (0)
  if (R == 8) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK(8, 1, 512); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK(8, 2, 512); else MOKA_LAUNCH_ATTN_PACK(8, 4, 512); }
  // (R = 16 at 512 threads needs 128 registers: one CTA per SM and a second wave; 256 threads fit two)
  else if (R == 16) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK(16, 1, 256); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK(16, 2, 256); else MOKA_LAUNCH_ATTN_PACK(16, 4, 256); }
  else if (R == 32) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK(32, 1, 256); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK(32, 2, 256); else MOKA_LAUNCH_ATTN_PACK(32, 4, 256); }
  else if (R == 64) { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK(64, 1, 256); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK(64, 2, 256); else MOKA_LAUNCH_ATTN_PACK(64, 4, 256); }
  else { if (KBs == 1) MOKA_LAUNCH_ATTN_PACK(128, 1, 256); else if (KBs == 2) MOKA_LAUNCH_ATTN_PACK(128, 2, 256); else MOKA_LAUNCH_ATTN_PACK(128, 4, 256); }
#undef MOKA_LAUNCH_ATTN_PACK
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  return check_launch_tc();
}

static int launch_attn_nopack(const DevDesc& d, const SimtArgs& a, cudaStream_t st) {
  const PackLayout L0 = pack_layout(d.T, d.r, d.M);
  return launch_attn_pack(d, a, L0, dgrad_parts(d, 0, false, 0), nullptr, nullptr, nullptr, st, 0);
}

// Shape checks and ring sizing of k_wgrad_tc (false: the kernel cannot tile this shape).
static bool wgrad_plan1(const DevDesc& d, bool has_dx, int num_sms, WgradTcParams& p, PackLayout& L, size_t& fixed);

// nchunks > 1: the packed rank axis (M r columns) does not fit the kernel's shared-memory tiles in one piece
// (r = 64 with three modalities, r = 128): the kernel then runs once per MODALITY - a "one modality" problem on
// sub-matrices of the packed operands and on A_m / dA_m.  dx is read-modify-written by every pass, but a token
// belongs to one modality, so its row gets a non-zero update (and a bf16 rounding that changes it) in exactly
// one of them: results are identical to a single pass.
static bool wgrad_plan(const DevDesc& d, bool has_dx, int num_sms, WgradTcParams& p, PackLayout& L, size_t& fixed,
                       int& nchunks) {
  nchunks = 1;
  const bool single = wgrad_plan1(d, has_dx, num_sms, p, L, fixed);
  // a single pass whose tiles leave room for only one operand buffer or one x slot (r = 64 with two modalities) runs
  // its load -> MMA -> reload chain unpipelined; per-modality passes have half the operand width and, since each
  // pass only visits the token tiles that hold its modality (mod_filter), stream little more than a single pass
  const bool degraded = single && (p.nop < 2 || p.xs < 2 || (has_dx && p.ds < 2));
  if (single && !degraded) return true;
  if (d.r % 64 || d.T % 8 || d.M < 2) return single;
  WgradTcParams p1 = p;
  PackLayout L1;
  size_t fixed1 = 0;
  DevDesc d1 = d;
  d1.M = 1;                                                   // one pass per modality (see launch_wgrad_tc)
  if (!wgrad_plan1(d1, has_dx, num_sms, p1, L1, fixed1)) return single;
  if (single && (p1.nop < 2 || p1.xs < 2)) return true;       // no better: keep the single pass
  p = p1; fixed = fixed1;
  L = pack_layout(d.T, d.r, d.M);                             // the pack step still produces the full operands
  nchunks = d.M;
  return true;
}

static bool wgrad_plan1(const DevDesc& d, bool has_dx, int num_sms, WgradTcParams& p, PackLayout& L, size_t& fixed) {
  if (d.T % 8) return false;                                  // transposed operand rows must be 16-byte multiples
  L = pack_layout(d.T, d.r, d.M);
  if (L.NPW > 256) return false;
  p.prefetch_dx = 1;
  p.T = d.T; p.d_in = d.d_in; p.r = d.r; p.rp = L.rp; p.M = d.M; p.NPW = L.NPW; p.KP = L.KP;
  p.ntiles = (d.T + 127) / 128;
  const int slices = d.d_in / 128;
  // token split: a CTA owns (column slice, token range).  Pick the split that minimises the makespan in tile
  // steps: waves of CTAs (one CTA per SM) x (tiles per CTA + ~1.2 tile-equivalents of ramp and dA flush).
  // (4096 columns, 64 tiles -> 4; 11008 columns, 256 tiles -> 5 instead of the 86-CTA single wave.)
  {
    int best = 1;
    double best_cost = 1e30;
    for (int ts = 1; ts <= min(p.ntiles, 24); ++ts) {
      const int waves = (slices * ts + num_sms - 1) / num_sms;
      const double cost = waves * ((p.ntiles + ts - 1) / ts + 1.2);
      if (cost < best_cost - 1e-9) { best_cost = cost; best = ts; }
    }
    p.tsplit = best;
  }
  p.has_dx = has_dx;
  p.delta = (has_dx && d.delta) ? 1 : 0;
  p.early_bits = drop_early_bits(1);               // (k_wgrad_tc alone: 160.3 -> 161.3 us; inside the 7B step: 193 -> 188 us)
  {
    static int dbg_drop_env = -1;
    if (dbg_drop_env < 0) { const char* ev = getenv("MOKA_DBG_WGRAD_DROP"); dbg_drop_env = ev ? atoi(ev) : 0; }
    p.dbg_drop = dbg_drop_env;
  }
  p.dbg = dbg_enabled();
  // MOKA_WGRAD_USE_HC=0: the dx product reads the transposed operand (MN-major A, hi + lo) and the token-major dHc
  // tile / tensor disappears.  Measured round 2: equal when timed alone, but 207 vs 177 us inside the 7B step (SM clock
  // 1.5 GHz): the MN-major A fetch of those 6 MMAs sits on the critical path of every dx epilogue.  Default: 1.
  { static int hc_env = -1; if (hc_env < 0) { const char* e = getenv("MOKA_WGRAD_USE_HC"); hc_env = (e && e[0] == '0') ? 0 : 1; }
    p.use_hc = hc_env; }
  // (L2 prefetch ahead of the rings, MOKA_L2_PREFETCH=n tiles: measured slower inside the step - 195 / 204 / 216 us at
  //  n = 0 / 3 / 6 - so off by default)
  { static int pf_env = -1; if (pf_env < 0) { const char* e = getenv("MOKA_L2_PREFETCH"); pf_env = e ? atoi(e) : 0; }
    p.l2pf = pf_env; }
  // (CTA order: no measurable difference alone or inside the step - 184 vs 187 us; the round-1 order stays)
  { static int ord_env = -1; if (ord_env < 0) { const char* e = getenv("MOKA_WGRAD_ORDER"); ord_env = e ? atoi(e) : 0; }
    p.order = ord_env; }
  p.lo_off = p.use_hc ? L.lo_off : 0;
  p.hct_blk = (uint32_t)p.NPW * 128u;
  p.w_box_bytes = (uint32_t)(p.lo_off ? 2 * p.NPW : p.NPW) * 128u;
  p.op_bytes = (p.use_hc ? (uint32_t)(p.KP / 64) * 16384u : 0u) + 4u * p.hct_blk;
  { static int st_env = -1; if (st_env < 0) { const char* e = getenv("MOKA_WGRAD_STACK"); st_env = (e && e[0] == '0') ? 0 : 1; }
    p.stack = (st_env && 2 * p.NPW <= 256) ? 1 : 0; }
  p.tmem_cols = 512;
  const size_t limit = 226 * 1024;
  // prefer two operand buffers (breaks the MMA -> reload -> MMA chain); the dx ring then gets 2 slots
  p.nop = 2;
  // dropout: the masking warps hand the keep bits of each x slot to the epilogue through shared memory
  static int xs_env = -1;                                       // tuning aid: x slots (the dx ring gets the rest)
  if (xs_env < 0) { const char* e = getenv("MOKA_WGRAD_XS"); xs_env = e ? atoi(e) : 0; }
  // (one 2 KB keep-bit block per x slot: two slots next to a dx ring unless more are forced, four without dx)
  const size_t km_bytes = d.drop_thresh ? (size_t)((has_dx && xs_env <= 2) ? 2 : 4) * 2048 : 0;
  fixed = (size_t)2 * p.w_box_bytes + (size_t)p.nop * p.op_bytes + 1024 + 512 + km_bytes;
  if (fixed + 4 * 32768 > limit) { p.nop = 1; fixed -= p.op_bytes; }
  if (fixed + 2 * 32768 > limit) return false;
  int slots = (int)((limit - fixed) / 32768);
  if (p.has_dx) {
    p.xs = 2;
    p.ds = min(3, slots - p.xs);
    if (p.ds < 2) { p.xs = 1; p.ds = min(3, slots - 1); }
    if (p.ds < 1) return false;                               // (one dx slot: unpipelined, still far ahead of the warp-FMA kernel)
    if (xs_env > 0 && slots - xs_env >= 1 && xs_env <= 4) { p.xs = xs_env; p.ds = min(3, slots - xs_env); }
  } else {
    p.xs = min(4, slots); p.ds = 0;
  }
  // the dA flush stages [NPW][128] fp32 over the X ring and, past its end, the dx ring
  if ((size_t)p.NPW * 512 > (size_t)(p.xs + p.ds) * 32768) return false;
  return true;
}

static int launch_wgrad_tc(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, bool fused) {
  WgradTcParams p;
  PackLayout L;
  size_t fixed;
  int nchunks;
  if (!wgrad_plan(d, a.dx != nullptr, num_sms, p, L, fixed, nchunks)) return MOKA_EUNSUPPORTED;
  const int slices = d.d_in / 128;
  char* pk = static_cast<char*>(a.pack);
  __nv_bfloat16* hc = p.use_hc ? reinterpret_cast<__nv_bfloat16*>(pk) : nullptr;   // (token-major operand: round-1 form only)
  __nv_bfloat16* hct_hi = reinterpret_cast<__nv_bfloat16*>(pk + L.hc_bytes);
  __nv_bfloat16* hct_lo = reinterpret_cast<__nv_bfloat16*>(pk + L.hc_bytes + L.hct_bytes);
  if (fused) {
    if (!(d.dbg_skip & 16)) {
      const int ntiles128 = (d.T + 127) / 128;
      if (int rc = launch_attn_pack(d, a, L, dgrad_parts(d, num_sms, true, (long long)(a.dh_att - a.dh)), hc, hct_hi,
                                    hct_lo, st)) return rc;
    }
  } else {
    k_pack_dh<<<(d.T + 63) / 64, 256, (size_t)64 * (d.r + 1) * 4, st>>>(d, L.rp, L.NPW, L.KP, L.lo_off, a.tokmeta, a.kidx, a.klen, a.dh,
                                                                       a.dh_att, a.dK, hc, hct_hi, hct_lo,
                                                                       d.pre * d.drop_scale);
    if (int rc = check_launch_tc()) return rc;
  }
  if (d.dbg_skip & 32) return MOKA_OK;
  CUtensorMap tx, tdx, ta[3], thc, thh, thl;
  if (!make_tmap(&tx, a.x, (uint64_t)d.T, (uint64_t)d.d_in, 128)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&tdx, p.has_dx ? a.dx : a.x, (uint64_t)d.T, (uint64_t)d.d_in, 128)) return MOKA_EUNSUPPORTED;
  const size_t smem = (size_t)(p.xs + p.ds) * 32768 + fixed;
  if (nchunks > 1) {
    cudaError_t e = d.drop_thresh
                        ? cudaFuncSetAttribute(k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                        : cudaFuncSetAttribute(k_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
    const __nv_bfloat16* const* Aw = reinterpret_cast<const __nv_bfloat16* const*>(a.A);
    for (int c = 0; c < nchunks; ++c) {                        // c = modality; r % 64 == 0, so rp == r
      const __nv_bfloat16* Ac = Aw[c];
      float* dAc = a.dA[c];
      const uint64_t w = (uint64_t)L.rp;                       // packed columns of one modality
      CUtensorMap tac, thc_c, thh_c, thl_c;
      if (!make_tmap(&tac, Ac, w, (uint64_t)d.d_in, (uint32_t)w)) return MOKA_EUNSUPPORTED;
      // (the token-major operand map is only dereferenced by the round-1 form; any valid map will do otherwise)
      if (!make_tmap_strided(&thc_c, (hc ? hc : reinterpret_cast<__nv_bfloat16*>(pk)) + (size_t)c * w, (uint64_t)d.T, w, (uint64_t)L.KP, 128)) return MOKA_EUNSUPPORTED;
      if (!make_tmap(&thh_c, hct_hi + (size_t)c * w * d.T, w, (uint64_t)d.T, (uint32_t)w)) return MOKA_EUNSUPPORTED;
      if (!make_tmap(&thl_c, hct_lo + (size_t)c * w * d.T, w, (uint64_t)d.T, (uint32_t)w)) return MOKA_EUNSUPPORTED;
      p.prefetch_dx = (c == 0);
      p.delta = (c == 0 && p.has_dx && d.delta) ? 1 : 0;       // later passes add onto what the first one wrote
      p.mod_filter = p.delta ? -1 : c + 1;                     // (a write-only first pass must store every tile)
      if (d.drop_thresh)
        e = launch_pdl(k_wgrad_tc<true>, dim3(slices * p.tsplit), dim3(kWgradDropThreads), smem, st, tx, tdx, tac, tac, tac,
                       thc_c, thh_c, thl_c, p, d, dAc, dAc, dAc, a.tokmeta);
      else
        e = launch_pdl(k_wgrad_tc<false>, dim3(slices * p.tsplit), dim3(kWgradThreads), smem, st, tx, tdx, tac, tac, tac,
                       thc_c, thh_c, thl_c, p, d, dAc, dAc, dAc, a.tokmeta);
      if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
      if (int rc = check_launch_tc()) return rc;
    }
    return MOKA_OK;
  }
  for (int m = 0; m < 3; ++m)
    if (!make_tmap(&ta[m], a.A[m < d.M ? m : 0], (uint64_t)d.r, (uint64_t)d.d_in, (uint32_t)d.r)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&thc, hc ? hc : reinterpret_cast<__nv_bfloat16*>(pk), (uint64_t)d.T, (uint64_t)L.KP, 128)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&thh, hct_hi, (uint64_t)L.NPW, (uint64_t)d.T, (uint32_t)L.NPW)) return MOKA_EUNSUPPORTED;
  if (!make_tmap(&thl, hct_lo, (uint64_t)L.NPW, (uint64_t)d.T, (uint32_t)L.NPW)) return MOKA_EUNSUPPORTED;
  cudaError_t e = d.drop_thresh
                      ? cudaFuncSetAttribute(k_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(k_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  p.mod_filter = p.delta ? -1 : 0;                            // single pass: only tiles made of padding are skipped
  if (d.drop_thresh)
    e = launch_pdl(k_wgrad_tc<true>, dim3(slices * p.tsplit), dim3(kWgradDropThreads), smem, st, tx, tdx, ta[0], ta[1], ta[2],
                   thc, thh, thl, p, d, a.dA[0], a.dA[d.M > 1 ? 1 : 0], a.dA[d.M > 2 ? 2 : 0], a.tokmeta);
  else
    e = launch_pdl(k_wgrad_tc<false>, dim3(slices * p.tsplit), dim3(kWgradThreads), smem, st, tx, tdx, ta[0], ta[1], ta[2],
                   thc, thh, thl, p, d, a.dA[0], a.dA[d.M > 1 ? 1 : 0], a.dA[d.M > 2 ? 2 : 0], a.tokmeta);
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  return check_launch_tc();
}

// ---------------------------------------------------------------------------- dispatch
int tc_stage_mask(const DevDesc& d, int dtype) {
  if (dtype != MOKA_BF16) return 0;                 // fp32 storage keeps exact fp32 FMA (1e-5 bar)
  if (d.d_in % 64 || d.d_out % 64) return 0;
  if (d.r % 4 || d.r > 128) return 0;         // (r = 4, the reference's shipped rank, pads to 8 / 16 inside the tiles)
  if (!get_encode_fn()) return 0;
  int mask = kStageDown;
  if (d.d_out % 128 == 0) mask |= kStageUp | kStageDgrad;
  if (d.d_in % 128 == 0) mask |= kStageWgrad;
  return mask;
}

int tc_forward(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  // the split down-projection hands two partial buffers to k_up_tc: only when both stages run here
  const bool both = (stages & kStageDown) && (stages & kStageUp) && a.hpart != nullptr;
  if ((stages & kStageDown) && !(d.dbg_skip & 1)) {
    if (int e = launch_down_tc(d, a, num_sms, st, both)) return e;
  }
  if ((stages & kStageUp) && (d.dbg_skip & (2 | 64)) != (2 | 64)) {
    if (int e = launch_up_tc(d, a, num_sms, st, both && down_ksplit(d, num_sms, true) > 1)) return e;
  }
  return MOKA_OK;
}

int tc_backward(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages, bool fused) {
  // fused (both stages here, k_attn_pack applicable): prep -> dgrad -> attn+pack -> wgrad, all chained
  // with programmatic dependent launch; nothing but dK and the counters is zero-filled
  if (stages & kStageDgrad) {
    if (int e = launch_dgrad_tc(d, a, num_sms, st, fused)) return e;
  }
  if (stages & kStageWgrad) {
    if (int e = launch_wgrad_tc(d, a, num_sms, st, fused)) return e;
  }
  return MOKA_OK;
}

}  // namespace moka

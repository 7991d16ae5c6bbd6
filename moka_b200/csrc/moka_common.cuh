// Shared device helpers for libmoka.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "moka.h"

namespace moka {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;

// Launch bookkeeping shared by every translation unit (moka_launch_count()).
void note_launch(int n = 1);
void set_cuda_error(cudaError_t e);

struct DevDesc {            // by-value kernel argument (POD)
  int B, S, T;
  int d_in, d_out, r, M, Lk;
  float pre, post[4];       // post[code], code 1..3 (post[0] unused)
  float attn_w, sm_scale;   // sm_scale = 1/sqrt(r)
  float drop_scale;         // 1/(1-p) or 1
  uint32_t drop_thresh;     // drop probability = thresh / 4096 ; 0 = dropout off
  uint32_t drop_sel[12];    // all-ones where the corresponding bit of drop_thresh is set
  unsigned long long seed, offset;
  uint32_t key_lo, key_odd; // dropout generator keys (splitmix64 of seed, offset; key_odd is odd)
  int tile_tokens;          // tokens per CTA tile
  int num_tiles;
  int dbg_skip;             // MOKA_FLAG_DBG_SKIP bits (timing aid)
  int dbg;                  // in-kernel time stamps enabled (MOKA_DEBUG_TIMING=1)
  int delta;                // MOKA_FLAG_DELTA_OUT: y / dx receive the adapter term only (write-only outputs)
};

// ---------------------------------------------------------------------------- vectors
template <typename T> struct Vec8;                  // 8 consecutive elements
template <> struct Vec8<__nv_bfloat16> { uint4 v; };
template <> struct Vec8<float> { float4 a, b; };

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// plain (coherent) streaming load for buffers that this kernel also writes (dx, y in place)
__device__ __forceinline__ uint4 ldg_rw(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_rw(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// load 8 consecutive elements as fp32.  STREAM: read-only, bypass L1.  RW: coherent.
template <typename T, bool RW> struct Load8;
template <bool RW> struct Load8<__nv_bfloat16, RW> {
  static __device__ __forceinline__ void run(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 v = RW ? ldg_rw(reinterpret_cast<const uint4*>(p)) : ldg_stream(reinterpret_cast<const uint4*>(p));
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
  }
};
template <bool RW> struct Load8<float, RW> {
  static __device__ __forceinline__ void run(const float* p, float (&f)[8]) {
    float4 a = RW ? ldg_rw(reinterpret_cast<const float4*>(p)) : ldg_stream(reinterpret_cast<const float4*>(p));
    float4 b = RW ? ldg_rw(reinterpret_cast<const float4*>(p) + 1) : ldg_stream(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
// cached (L1-allocating) load of 8 weight elements
template <typename T> struct LoadW8;
template <> struct LoadW8<__nv_bfloat16> {
  static __device__ __forceinline__ void run(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
  }
};
template <> struct LoadW8<float> {
  static __device__ __forceinline__ void run(const float* p, float (&f)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};
template <typename T> struct Store8;
template <> struct Store8<__nv_bfloat16> {
  static __device__ __forceinline__ void run(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
    v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
    stg_stream(reinterpret_cast<uint4*>(p), v);
  }
};
template <> struct Store8<float> {
  static __device__ __forceinline__ void run(float* p, const float (&f)[8]) {
    stg_stream(reinterpret_cast<float4*>(p), make_float4(f[0], f[1], f[2], f[3]));
    stg_stream(reinterpret_cast<float4*>(p) + 1, make_float4(f[4], f[5], f[6], f[7]));
  }
};
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---------------------------------------------------------------------------- dropout keep masks
// Bit-sliced Bernoulli generator: one evaluation yields the 32 drop decisions of a 32-column super-group
// (token t, columns 32 sg .. 32 sg + 31), each bit set with probability T / 2^12, T = drop_thresh =
// round(p * 4096).  Twelve pseudo-random words R_i are folded from the least significant bit of T upwards,
//     x <- bit_i(T) ? (R_i | x) : (R_i & x),
// which halves-and-shifts the per-bit probability exactly like long division of T by 2^12.  The words are
// R_i = hi16(y C_i) : hi16(y C'_i) (odd constants C_i, C'_i; see the loop) drawn from one full avalanche
// y = mix32((idx_lo ^ key_lo) + idx_hi * golden) ^ key_odd of the super-group index idx = t * ceil(d_in / 32) + sg
// (mix32 = the 2-multiply finaliser "lowbias32", C. Wellons); (key_lo, key_odd) come from (seed, offset) by
// splitmix64 on the host.  ~2 integer ops per element (Philox4x32-10 + compares: ~17), cheap enough to mask
// the shared-memory tiles of the tensor-core kernels close to HBM rate.  Every kernel and the test hook
// moka_dropout_keep_mask use these functions, so forward, backward and the oracle see one mask.
constexpr int kDropBits = 12;
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu;
  x ^= x >> 15; x *= 0x735a2d97u;
  x ^= x >> 15;
  return x;
}
// 32 KEEP bits (bit j = column 32 sg + j is kept) of super-group `idx`.
__device__ __forceinline__ uint32_t dropout_keep32(const DevDesc& d, unsigned long long idx) {
  constexpr uint32_t C[16] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du,
                              0xFD7046C5u, 0xB55A4F09u, 0x94D049BBu, 0xBF58476Du, 0x2545F491u, 0x6C8E9CF5u,
                              0xE7037ED1u, 0x8EBC6AF1u, 0x589965CDu, 0x1D8E4E27u};
  constexpr uint32_t C2[12] = {0xA54FF53Bu, 0x72BE5D75u, 0x3C6EF373u, 0xF1BBCDCBu, 0x510E527Fu, 0x9B05688Du,
                               0x1F83D9ABu, 0x5BE0CD19u, 0xCBBB9D5Du, 0x629A292Bu, 0x9159015Bu, 0x152FECD9u};
  const uint32_t y = mix32(((uint32_t)idx ^ d.key_lo) + (uint32_t)(idx >> 32) * 0x9E3779B1u) ^ d.key_odd;
  uint32_t x = 0u;
#pragma unroll
  for (int i = 0; i < kDropBits; ++i) {
    // word i of the group: the HIGH halves of two 32-bit products (the well-mixed bits of a multiplicative hash),
    // joined by one PRMT: 2 fma-pipe + 2 alu-pipe instructions per word with the fold below.  (Round 1 used one
    // product, z ^ (z >> 15): 1 + 3, bound by the alu pipe - and its top 17 bits were raw product bits, which left the
    // joint drop rate of some position pairs inside a group off by ~50 sigma; tests: the pair check of
    // test_dropout_mask_statistics.  A 64-bit product (IMAD.WIDE, hi ^ lo) is statistically fine but slower.)
    const uint32_t za = y * C[i], zb = y * C2[i];
    uint32_t R;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(R) : "r"(zb), "r"(za));
    const uint32_t m = d.drop_sel[i];                          // all ones where bit i of T is set (constant bank)
    x = (R & x) | (m & (R | x));                               // one LOP3: m ? (R | x) : (R & x)
  }
  return ~x;
}
__device__ __forceinline__ unsigned long long dropout_row_base(const DevDesc& d, long long t) {
  return (unsigned long long)t * (unsigned long long)((d.d_in + 31) >> 5);
}
// keep mask for 8 consecutive columns (group g = col/8) of token t: bit i set = element kept.
__device__ __forceinline__ uint32_t dropout_keep8(const DevDesc& d, int t, int group) {
  const uint32_t k32 = dropout_keep32(d, dropout_row_base(d, t) + (unsigned long long)(group >> 2));
  return (k32 >> ((group & 3) * 8)) & 0xffu;
}
// Zero the dropped bf16 elements of one 32-column super-group of a token row inside a {64 cols x rows}
// 128B-swizzled shared-memory box: one 2-byte store per dropped element (5 % of them at p = 0.05) instead of
// a read-modify-write of every 16-byte chunk.  `box_row_addr` = shared address of the row's 128-byte line,
// `row7` = row & 7 (swizzle phase), `sg_in_box` = 0 / 1 (which half of the 64 columns).
__device__ __forceinline__ void dropout_zero_dropped(uint32_t keep32, uint32_t box_row_addr, uint32_t row7,
                                                     uint32_t sg_in_box) {
  uint32_t m = ~keep32;
  while (m) {
    const uint32_t e = (uint32_t)__ffs((int)m) - 1u;           // element 0..31 of the super-group
    m &= m - 1u;
    const uint32_t chunk = sg_in_box * 4u + (e >> 3);
    const uint32_t addr = box_row_addr + (((chunk ^ row7) << 4) | ((e & 7u) << 1));
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"((unsigned short)0) : "memory");
  }
}
// 8 keep bits -> AND-masks for the four bf16 pairs of the group (0xFFFF per kept element): a multiply
// spreads four bits onto the four byte sign positions, prmt with sign replication fills the halfwords.
__device__ __forceinline__ void dropout_expand8(uint32_t km, uint32_t (&W)[4]) {
  const uint32_t lo = ((km & 0xfu) * 0x10204080u) & 0x80808080u;
  const uint32_t hi = ((km >> 4) * 0x10204080u) & 0x80808080u;
  asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(W[0]) : "r"(lo));
  asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(W[1]) : "r"(lo));
  asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(W[2]) : "r"(hi));
  asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(W[3]) : "r"(hi));
}

// ---------------------------------------------------------------------------- warp utils
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 32 per-lane partials v[0..31] -> lane L returns sum over lanes of v[L] (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      float send = upper ? v[i] : v[i + half];
      float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace moka

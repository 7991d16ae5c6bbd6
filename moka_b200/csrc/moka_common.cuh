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
  uint32_t drop_thresh;     // keep iff u16 >= thresh ; 0 = dropout off
  unsigned long long seed, offset;
  int tile_tokens;          // tokens per CTA tile
  int num_tiles;
  int dbg_skip;             // MOKA_FLAG_DBG_SKIP bits (timing aid)
  int dbg;                  // in-kernel time stamps enabled (MOKA_DEBUG_TIMING=1)
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

// ---------------------------------------------------------------------------- Philox
// Philox4x32-10 (Salmon et al. 2011), counter = (group index, offset), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// keep mask for 8 consecutive columns (group g = col/8) of token t: bit i set = element kept.
__device__ __forceinline__ uint32_t dropout_keep8(const DevDesc& d, int t, int group) {
  unsigned long long idx = (unsigned long long)t * (unsigned long long)(d.d_in >> 3) + (unsigned long long)group;
  uint4 c = make_uint4((uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)d.offset, (uint32_t)(d.offset >> 32));
  uint4 o = philox4x32_10(c, make_uint2((uint32_t)d.seed, (uint32_t)(d.seed >> 32)));
  uint32_t w[4] = {o.x, o.y, o.z, o.w};
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m |= ((w[i] & 0xffffu) >= d.drop_thresh ? 1u : 0u) << (2 * i);
    m |= ((w[i] >> 16) >= d.drop_thresh ? 1u : 0u) << (2 * i + 1);
  }
  return m;
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
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace moka

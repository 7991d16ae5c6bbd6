// Warp-FMA ("SIMT") kernels of the MokA adapter layer: the path for fp32 storage and for
// small ranks (r < 16), and the always-available fallback for every shape.
//
// Four streaming kernels, each a single pass over one large tensor:
//   k_down : x  -> h  = pre * A_m drop(x)                          (reads x)
//   k_up   : y0 -> y  = y0 + post_m * B (h + w attn(h, K_b))       (reads y0, writes y)
//   k_dgrad: dy -> dg = post_m B^T dy, dB += post_m dy (x) g, attention backward -> dh, dK
//   k_wgrad: x, dx -> dx += pre drop^T(A_m^T dh), dA_m += pre dh (x) drop(x)
// Layout: a CTA owns a contiguous token tile; each warp owns 256-column chunks of the
// feature dimension (8 columns = one 128-bit load per lane) and keeps the matching weight
// fragment in registers for the whole tile, so the weights cost L2 traffic once per tile.
//
// Reference semantics: AudioVisualText/peft_hyper/tuners/lora.py:460-532,
// VisualText/modified_peft/tuners/lora/layer.py:589-671 (restated in oracle/moka_oracle.py).
#include "moka_attn.cuh"
#include "moka_common.cuh"
#include "moka_kernels.h"

namespace moka {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct APtrs { const void* p[3]; };
struct GPtrs { float* p[3]; };

// ======================================================================== k_down (x -> h)
// A warp owns a group of U consecutive tokens and walks the whole d_in for them, so every
// h element is produced by one warp in a fixed order: the forward is bit-reproducible
// (needed for checkpoint recomputation and for the bitwise span-indexing tests).
template <typename T, int RB>
__global__ void __launch_bounds__(kThreads, 2)
k_down(DevDesc d, const T* __restrict__ x, APtrs A, const uint8_t* __restrict__ tokmeta,
       float* __restrict__ h_stash) {
  constexpr int U = 32 / RB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = d.r;
  const int nchunks = (d.d_in + 255) >> 8;
  const int nrb = (r + RB - 1) / RB;
  const float out_scale = d.pre * d.drop_scale;
  const int ngroups = (d.T + U - 1) / U;

  for (int grp = blockIdx.x * kWarps + warp; grp < ngroups; grp += gridDim.x * kWarps) {
    const int g0 = grp * U;
    int code[U];
    int present = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      code[u] = (g0 + u < d.T) ? (tokmeta[g0 + u] & MOKA_MOD_MASK) : -1;
      if (code[u] > 0) present |= 1 << code[u];
    }
    // tokens in no modality: zero rows (they may still be listed as key rows)
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (code[u] == 0)
        for (int j = lane; j < r; j += 32) h_stash[(size_t)(g0 + u) * r + j] = 0.f;

    for (int m = 1; m <= d.M; ++m) {
      if (!((present >> m) & 1)) continue;
      const T* Am = static_cast<const T*>(A.p[m - 1]);
      for (int rb = 0; rb < nrb; ++rb) {
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll 2
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          const int col = (chunk << 8) + lane * 8;
          if (col >= d.d_in) continue;
          float xv[U][8];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (code[u] == m) {
              Load8<T, false>::run(x + (size_t)(g0 + u) * d.d_in + col, xv[u]);
            } else {
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) xv[u][cidx] = 0.f;
            }
          }
          if (d.drop_thresh) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (code[u] == m) {
                const uint32_t km = dropout_keep8(d, g0 + u, col >> 3);
#pragma unroll
                for (int cidx = 0; cidx < 8; ++cidx)
                  if (!((km >> cidx) & 1)) xv[u][cidx] = 0.f;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            const int row = rb * RB + j;
            float fr[8];
            if (row < r) {
              LoadW8<T>::run(Am + (size_t)row * d.d_in + col, fr);
            } else {
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) fr[cidx] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              float s = acc[u * RB + j];
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) s = fmaf(xv[u][cidx], fr[cidx], s);
              acc[u * RB + j] = s;
            }
          }
        }
        const float tot = warp_transpose_reduce32(acc, lane);
        const int u = lane / RB, j = lane % RB;
        const int row = rb * RB + j;
        int cu = -1;
#pragma unroll
        for (int uu = 0; uu < U; ++uu) if (uu == u) cu = code[uu];
        if (cu == m && row < r) h_stash[(size_t)(g0 + u) * r + row] = tot * out_scale;
      }
    }
  }
}

// B fragment for 8 output columns x RB rank indices: frag[c][j] = B[col+c][rb*RB+j]
template <typename T, int RB>
__device__ __forceinline__ void load_b_frag(const T* __restrict__ Bw, int col, int rb, int r, bool act,
                                            float (&frag)[8][RB]) {
  if (RB == 8 && (r & 7) == 0) {
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx) {
      if (act) {
        float f[8];
        LoadW8<T>::run(Bw + (size_t)(col + cidx) * r + rb * 8, f);
#pragma unroll
        for (int j = 0; j < RB; ++j) frag[cidx][j] = f[j];
      } else {
#pragma unroll
        for (int j = 0; j < RB; ++j) frag[cidx][j] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int cidx = 0; cidx < 8; ++cidx)
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int row = rb * RB + j;
        frag[cidx][j] = (act && row < r) ? to_f32<T>(Bw[(size_t)(col + cidx) * r + row]) : 0.f;
      }
  }
}

// ======================================================================== k_up (y0 -> y)
template <typename T, int RB, bool MULTI>
__global__ void __launch_bounds__(kThreads, 2)
k_up(DevDesc d, const T* y0, T* y, const T* __restrict__ Bw,
     const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
     const int32_t* __restrict__ klen, const float* __restrict__ h_stash, float* __restrict__ g_out,
     int kcap) {
  constexpr int U = 4;
  extern __shared__ __align__(16) float smem[];
  const int r = d.r;
  float* gs = smem;                                        // [tile_tokens, r]
  float* Ks = gs + (size_t)d.tile_tokens * r;              // [kcap, r+1]
  uint8_t* meta_s = reinterpret_cast<uint8_t*>(Ks + (size_t)kcap * (r + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (d.d_out + 255) >> 8;
  const int nrb = (r + RB - 1) / RB;
  const bool inplace = (y0 == y);
  AttnCtx c{tokmeta, kidx, klen, h_stash, nullptr, Ks, nullptr, r + 1, kcap, g_out + (size_t)d.T * d.r};
  __shared__ int s_or;
  const SyncGroup grp{(int)threadIdx.x, kThreads, 0, &s_or};

  for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
    const int t0 = tile * d.tile_tokens;
    const int nt = min(d.tile_tokens, d.T - t0);
    for (int i = tid; i < nt; i += kThreads) meta_s[i] = tokmeta[t0 + i];
    for (int i = tid; i < nt * r; i += kThreads) gs[i] = h_stash[(size_t)t0 * r + i];
    __syncthreads();
    attention_tile<false>(d, c, grp, t0, nt, meta_s, gs, gs, nullptr);
    __syncthreads();
    for (int i = tid; i < nt * r; i += kThreads) g_out[(size_t)t0 * r + i] = gs[i];   // post-attention stash

    for (int chunk = warp; chunk < nchunks; chunk += kWarps) {
      const int col = (chunk << 8) + lane * 8;
      const bool act = col < d.d_out;
      // r <= RB: the B fragment stays in registers for the whole tile.  r > RB: fragments are
      // re-read (L1 hits) per token group so that y is accumulated in fp32 and rounded ONCE.
      float frag[8][RB];
      if (!MULTI) load_b_frag<T, RB>(Bw, col, 0, r, act, frag);
      for (int g0 = 0; g0 < nt; g0 += U) {
        float yv[U][8];
        int code[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int tt = g0 + u;
          code[u] = (tt < nt) ? (meta_s[tt] & MOKA_MOD_MASK) : -1;
          // untouched tokens: nothing to do in place; plain copy otherwise
          const bool need = act && (code[u] > 0 || (code[u] == 0 && !inplace));
          if (need) {
            Load8<T, true>::run(y0 + (size_t)(t0 + tt) * d.d_out + col, yv[u]);
          } else {
            code[u] = -1;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (code[u] < 0) continue;
          const int tt = g0 + u;
          if (code[u] > 0) {
            const float ps = d.post[code[u]];
            for (int rb = 0; rb < (MULTI ? nrb : 1); ++rb) {
              if (MULTI) load_b_frag<T, RB>(Bw, col, rb, r, act, frag);
              const float* gq = gs + (size_t)tt * r + rb * RB;
              float gv[RB];
#pragma unroll
              for (int j = 0; j < RB; ++j) gv[j] = (rb * RB + j < r) ? gq[j] * ps : 0.f;
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) {
                float sacc = yv[u][cidx];
#pragma unroll
                for (int j = 0; j < RB; ++j) sacc = fmaf(frag[cidx][j], gv[j], sacc);
                yv[u][cidx] = sacc;
              }
            }
          }
          Store8<T>::run(y + (size_t)(t0 + tt) * d.d_out + col, yv[u]);
        }
      }
    }
    __syncthreads();
  }
}

// ======================================================================== k_dgrad (dy -> dh, dB, dK)
template <typename T, int RB>
__global__ void __launch_bounds__(kThreads, 1)
k_dgrad(DevDesc d, const T* __restrict__ dy, const T* __restrict__ Bw,
        const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
        const int32_t* __restrict__ klen, const float* __restrict__ h_stash,
        float* __restrict__ dB, float* __restrict__ dh_out, float* __restrict__ dK_glob, int kcap) {
  constexpr int U = 32 / RB;
  extern __shared__ __align__(16) float smem[];
  const int r = d.r;
  float* hs = smem;                                        // [tile_tokens, r] h (pre-attention)
  float* gs = hs + (size_t)d.tile_tokens * r;              // g (post-attention)
  float* dgs = gs + (size_t)d.tile_tokens * r;             // dg -> dh
  float* Ks = dgs + (size_t)d.tile_tokens * r;             // [kcap, r+1]
  float* dKs = Ks + (size_t)kcap * (r + 1);                // [kcap, r+1]
  uint8_t* meta_s = reinterpret_cast<uint8_t*>(dKs + (size_t)kcap * (r + 1));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (d.d_out + 255) >> 8;
  const int nrb = (r + RB - 1) / RB;
  AttnCtx c{tokmeta, kidx, klen, h_stash, dK_glob, Ks, dKs, r + 1, kcap, nullptr};
  __shared__ int s_or;
  const SyncGroup grp{(int)threadIdx.x, kThreads, 0, &s_or};

  for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
    const int t0 = tile * d.tile_tokens;
    const int nt = min(d.tile_tokens, d.T - t0);
    for (int i = tid; i < nt; i += kThreads) meta_s[i] = tokmeta[t0 + i];
    for (int i = tid; i < nt * r; i += kThreads) {
      const float v = h_stash[(size_t)t0 * r + i];
      hs[i] = v; gs[i] = v; dgs[i] = 0.f;
    }
    __syncthreads();
    attention_tile<false>(d, c, grp, t0, nt, meta_s, hs, gs, nullptr);     // recompute g
    __syncthreads();

    for (int chunk = warp; chunk < nchunks; chunk += kWarps) {
      const int col = (chunk << 8) + lane * 8;
      const bool act = col < d.d_out;
      for (int rb = 0; rb < nrb; ++rb) {
        float frag[8][RB], dBacc[8][RB];
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx)
#pragma unroll
          for (int j = 0; j < RB; ++j) {
            const int row = rb * RB + j;
            frag[cidx][j] = (act && row < r) ? to_f32<T>(Bw[(size_t)(col + cidx) * r + row]) : 0.f;
            dBacc[cidx][j] = 0.f;
          }
        for (int g0 = 0; g0 < nt; g0 += U) {
          float dv[U][8];
          float ps[U];
          bool any = false;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int tt = g0 + u;
            const int code = (tt < nt) ? (meta_s[tt] & MOKA_MOD_MASK) : 0;
            ps[u] = code ? d.post[code] : 0.f;
            any |= code != 0;
            if (code && act) {
              Load8<T, false>::run(dy + (size_t)(t0 + tt) * d.d_out + col, dv[u]);
            } else {
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) dv[u][cidx] = 0.f;
            }
          }
          if (!any) continue;
          float acc[32];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int tt = min(g0 + u, nt - 1);
            const float* gq = gs + (size_t)tt * r + rb * RB;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
              float s = 0.f;
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) s = fmaf(dv[u][cidx], frag[cidx][j], s);
              acc[u * RB + j] = s * ps[u];
              const float gp = (rb * RB + j < r) ? gq[j] * ps[u] : 0.f;
#pragma unroll
              for (int cidx = 0; cidx < 8; ++cidx) dBacc[cidx][j] = fmaf(dv[u][cidx], gp, dBacc[cidx][j]);
            }
          }
          const float tot = warp_transpose_reduce32(acc, lane);
          const int u = lane / RB, j = lane % RB;
          const int tt = g0 + u, row = rb * RB + j;
          if (tt < nt && row < r && (meta_s[tt] & MOKA_MOD_MASK)) atomicAdd(&dgs[(size_t)tt * r + row], tot);
        }
        if (act) {
#pragma unroll
          for (int cidx = 0; cidx < 8; ++cidx)
#pragma unroll
            for (int j = 0; j < RB; ++j) {
              const int row = rb * RB + j;
              if (row < r) red_add_f32(&dB[(size_t)(col + cidx) * r + row], dBacc[cidx][j]);
            }
        }
      }
    }
    __syncthreads();
    attention_tile<true>(d, c, grp, t0, nt, meta_s, hs, nullptr, dgs);      // dg -> dh, dK
    __syncthreads();
    for (int i = tid; i < nt * r; i += kThreads) dh_out[(size_t)t0 * r + i] = dgs[i];
    __syncthreads();
  }
}

// ======================================================================== k_wgrad (x, dx -> dx, dA)
template <typename T, int RB>
__global__ void __launch_bounds__(kThreads, 2)
k_wgrad(DevDesc d, const T* __restrict__ x, T* dx, APtrs A, GPtrs dA,
        const uint8_t* __restrict__ tokmeta, const int32_t* __restrict__ kidx,
        const int32_t* __restrict__ klen, const float* __restrict__ dh_in, const float* __restrict__ dh_att,
        const float* __restrict__ dK_glob) {
  constexpr int U = 4;
  extern __shared__ __align__(16) float smem[];
  const int r = d.r;
  float* dhs = smem;                                       // [tile_tokens, r]
  uint8_t* meta_s = reinterpret_cast<uint8_t*>(dhs + (size_t)d.tile_tokens * r);
  __shared__ int s_present;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nchunks = (d.d_in + 255) >> 8;
  const int nrb = (r + RB - 1) / RB;
  const float scale = d.pre * d.drop_scale;

  for (int tile = blockIdx.x; tile < d.num_tiles; tile += gridDim.x) {
    const int t0 = tile * d.tile_tokens;
    const int nt = min(d.tile_tokens, d.T - t0);
    if (tid == 0) s_present = 0;
    for (int i = tid; i < nt * r; i += kThreads) dhs[i] = dh_in[(size_t)t0 * r + i] + dh_att[(size_t)t0 * r + i];
    __syncthreads();
    for (int i = tid; i < nt; i += kThreads) {
      const uint8_t mb = tokmeta[t0 + i];
      meta_s[i] = mb;
      if (mb & MOKA_MOD_MASK) atomicOr(&s_present, 1 << (mb & MOKA_MOD_MASK));
    }
    // fold dK into the dh rows of this tile's key tokens (each key row maps to one token)
    const int b_first = t0 / d.S, b_last = (t0 + nt - 1) / d.S;
    for (int b = b_first; b <= b_last; ++b) {
      const int L = klen[b];
      for (int k = tid; k < L; k += kThreads) {
        const int pos = kidx[(size_t)b * d.Lk + k];
        const int t = b * d.S + pos;
        if (t < t0 || t >= t0 + nt) continue;
        const uint8_t mb = tokmeta[t];
        if (!(mb & MOKA_KEY_VALID) || !(mb & MOKA_MOD_MASK)) continue;
        const float* src = dK_glob + ((size_t)b * d.Lk + k) * r;
        float* dst = dhs + (size_t)(t - t0) * r;
        for (int j = 0; j < r; ++j) dst[j] += src[j];
      }
    }
    __syncthreads();
    const int present = s_present;

    for (int chunk = warp; chunk < nchunks; chunk += kWarps) {
      const int col = (chunk << 8) + lane * 8;
      const bool act = col < d.d_in;
      // ---- phase A: dA_m += scale * sum_t dh_t (x) u_t   (streams x; accumulators in registers)
      for (int m = 1; m <= d.M; ++m) {
        if (!((present >> m) & 1)) continue;
        float* dAm = dA.p[m - 1];
        for (int rb = 0; rb < nrb; ++rb) {
          float dAacc[RB][8];
#pragma unroll
          for (int j = 0; j < RB; ++j)
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) dAacc[j][cidx] = 0.f;
          for (int g0 = 0; g0 < nt; g0 += U) {
            float xv[U][8];
            bool sel[U];
            bool any = false;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int tt = g0 + u;
              sel[u] = act && (tt < nt) && ((meta_s[min(tt, nt - 1)] & MOKA_MOD_MASK) == m);
              any |= sel[u];
              if (sel[u]) Load8<T, false>::run(x + (size_t)(t0 + tt) * d.d_in + col, xv[u]);
            }
            if (!any) continue;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (!sel[u]) continue;
              const int tt = g0 + u;
              if (d.drop_thresh) {
                const uint32_t km = dropout_keep8(d, t0 + tt, col >> 3);
#pragma unroll
                for (int cidx = 0; cidx < 8; ++cidx)
                  if (!((km >> cidx) & 1)) xv[u][cidx] = 0.f;
              }
              const float* dq = dhs + (size_t)tt * r + rb * RB;
#pragma unroll
              for (int j = 0; j < RB; ++j) {
                const float dv = (rb * RB + j < r) ? dq[j] : 0.f;
#pragma unroll
                for (int cidx = 0; cidx < 8; ++cidx) dAacc[j][cidx] = fmaf(dv, xv[u][cidx], dAacc[j][cidx]);
              }
            }
          }
          if (act) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
              const int row = rb * RB + j;
              if (row < r) {
                float* p = dAm + (size_t)row * d.d_in + col;
                red_add_v4_f32(p, dAacc[j][0] * scale, dAacc[j][1] * scale, dAacc[j][2] * scale, dAacc[j][3] * scale);
                red_add_v4_f32(p + 4, dAacc[j][4] * scale, dAacc[j][5] * scale, dAacc[j][6] * scale, dAacc[j][7] * scale);
              }
            }
          }
        }
      }
      // ---- phase B: dx += scale * keep * A_m^T dh   (read-modify-write of dx, rounded once)
      if (dx != nullptr && act) {
        for (int g0 = 0; g0 < nt; g0 += U) {
          float gx[U][8];
          int code[U];
          int pres = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int tt = g0 + u;
            code[u] = (tt < nt) ? (meta_s[tt] & MOKA_MOD_MASK) : 0;
            if (code[u]) {
              pres |= 1 << code[u];
              Load8<T, true>::run(dx + (size_t)(t0 + tt) * d.d_in + col, gx[u]);
            }
          }
          if (!pres) continue;
          float sacc[U][8];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx) sacc[u][cidx] = 0.f;
          for (int m = 1; m <= d.M; ++m) {
            if (!((pres >> m) & 1)) continue;
            const T* Am = static_cast<const T*>(A.p[m - 1]);
            for (int rb = 0; rb < nrb; ++rb) {
#pragma unroll
              for (int j = 0; j < RB; ++j) {
                const int row = rb * RB + j;
                if (row >= r) continue;
                float fr[8];
                LoadW8<T>::run(Am + (size_t)row * d.d_in + col, fr);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                  if (code[u] != m) continue;
                  const float dv = dhs[(size_t)(g0 + u) * r + row];
#pragma unroll
                  for (int cidx = 0; cidx < 8; ++cidx) sacc[u][cidx] = fmaf(dv, fr[cidx], sacc[u][cidx]);
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (!code[u]) continue;
            const int tt = g0 + u;
            uint32_t km = 0xffu;
            if (d.drop_thresh) km = dropout_keep8(d, t0 + tt, col >> 3);
#pragma unroll
            for (int cidx = 0; cidx < 8; ++cidx)
              if ((km >> cidx) & 1) gx[u][cidx] = fmaf(sacc[u][cidx], scale, gx[u][cidx]);
            Store8<T>::run(dx + (size_t)(t0 + tt) * d.d_in + col, gx[u]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ======================================================================== dropout mask dump (test hook)
__global__ void k_keep_mask(DevDesc d, uint8_t* keep) {
  const size_t groups = (size_t)d.T * (d.d_in >> 3);
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(g / (d.d_in >> 3));
    const int grp = (int)(g % (d.d_in >> 3));
    const uint32_t km = d.drop_thresh ? dropout_keep8(d, t, grp) : 0xffu;
    uint8_t* o = keep + (size_t)t * d.d_in + grp * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (km >> i) & 1;
  }
}

// ======================================================================== host launchers
static inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  }
  return MOKA_OK;
}

static int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

// Choose the token tile so that (a) the smem rows fit, (b) the grid fills every SM twice.
static void plan_tiles(DevDesc& d, int rowbufs, int keybufs, int num_sms, size_t* smem_bytes, int* kcap_out,
                       int* grid_out, int klen_hint) {
  const size_t budget = 100 * 1024;                         // two CTAs per SM
  const int r = d.r;
  int kcap = 0;
  if (keybufs) {
    kcap = min(max(klen_hint, 1), 32 * kKeyBlocksMax);
    const size_t per_key = (size_t)keybufs * (r + 1) * 4;
    const size_t kmax = (budget / 2) / per_key;
    if ((size_t)kcap > kmax) kcap = (int)kmax;                // longer key lists take the slow path
  }
  const size_t key_bytes = (size_t)keybufs * kcap * (r + 1) * 4;
  const size_t per_tok = (size_t)rowbufs * r * 4 + 1;
  int cap = (int)((budget - key_bytes) / per_tok);
  if (cap > 64) cap = 64;
  if (cap < 1) cap = 1;
  const int slots = num_sms * 2;
  const int waves = (d.T + slots * cap - 1) / (slots * cap);
  int tt = (d.T + slots * waves - 1) / (slots * waves);
  if (tt > cap) tt = cap;
  if (tt < 1) tt = 1;
  d.tile_tokens = tt;
  d.num_tiles = (d.T + tt - 1) / tt;
  *smem_bytes = align16((size_t)rowbufs * tt * r * 4 + key_bytes + tt + 16);
  *kcap_out = kcap;
  *grid_out = min(d.num_tiles, slots);
}

template <typename T, int RB>
static int launch_forward_t(DevDesc d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  APtrs A{{a.A[0], a.A[1], a.A[2]}};
  size_t smem; int kcap, grid;
  if (stages & kStageDown) {
    constexpr int U = 32 / RB;
    const int ngroups = (d.T + U - 1) / U;
    grid = min((ngroups + kWarps - 1) / kWarps, num_sms * 2);
    k_down<T, RB><<<grid, kThreads, 0, st>>>(d, static_cast<const T*>(a.x), A, a.tokmeta, a.h_stash);
    if (int e = check_launch()) return e;
  }
  if (stages & kStageUp) {
    DevDesc dd = d;
    plan_tiles(dd, 1, 1, num_sms, &smem, &kcap, &grid, a.klen_hint);
    if (d.r <= RB) {
      if (int e = set_smem(k_up<T, RB, false>, smem)) return e;
      k_up<T, RB, false><<<grid, kThreads, smem, st>>>(dd, static_cast<const T*>(a.y0), static_cast<T*>(a.y),
                                                       static_cast<const T*>(a.Bw), a.tokmeta, a.kidx, a.klen,
                                                       a.h_stash, a.h_stash + (size_t)d.T * d.r, kcap);
    } else {
      if (int e = set_smem(k_up<T, RB, true>, smem)) return e;
      k_up<T, RB, true><<<grid, kThreads, smem, st>>>(dd, static_cast<const T*>(a.y0), static_cast<T*>(a.y),
                                                      static_cast<const T*>(a.Bw), a.tokmeta, a.kidx, a.klen,
                                                      a.h_stash, a.h_stash + (size_t)d.T * d.r, kcap);
    }
    if (int e = check_launch()) return e;
  }
  return MOKA_OK;
}

template <typename T, int RB>
static int launch_backward_t(DevDesc d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  APtrs A{{a.A[0], a.A[1], a.A[2]}};
  GPtrs G{{a.dA[0], a.dA[1], a.dA[2]}};
  size_t smem; int kcap, grid;
  if (stages & kStageDgrad) {
    DevDesc dd = d;
    plan_tiles(dd, 3, 2, num_sms, &smem, &kcap, &grid, a.klen_hint);
    if (int e = set_smem(k_dgrad<T, RB>, smem)) return e;
    k_dgrad<T, RB><<<grid, kThreads, smem, st>>>(dd, static_cast<const T*>(a.dy), static_cast<const T*>(a.Bw),
                                                 a.tokmeta, a.kidx, a.klen, a.h_stash, a.dB, a.dh, a.dK, kcap);
    if (int e = check_launch()) return e;
  }
  if (stages & kStageWgrad) {
    DevDesc dd = d;
    plan_tiles(dd, 1, 0, num_sms, &smem, &kcap, &grid, 0);
    if (int e = set_smem(k_wgrad<T, RB>, smem)) return e;
    k_wgrad<T, RB><<<grid, kThreads, smem, st>>>(dd, static_cast<const T*>(a.x), static_cast<T*>(a.dx), A, G,
                                                 a.tokmeta, a.kidx, a.klen, a.dh, a.dh_att, a.dK);
    if (int e = check_launch()) return e;
  }
  return MOKA_OK;
}

int simt_forward(const DevDesc& d, int dtype, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  const bool small = d.r <= 4;
  if (dtype == MOKA_BF16)
    return small ? launch_forward_t<__nv_bfloat16, 4>(d, a, num_sms, st, stages) : launch_forward_t<__nv_bfloat16, 8>(d, a, num_sms, st, stages);
  return small ? launch_forward_t<float, 4>(d, a, num_sms, st, stages) : launch_forward_t<float, 8>(d, a, num_sms, st, stages);
}

int simt_backward(const DevDesc& d, int dtype, const SimtArgs& a, int num_sms, cudaStream_t st, int stages) {
  const bool small = d.r <= 4;
  if (dtype == MOKA_BF16)
    return small ? launch_backward_t<__nv_bfloat16, 4>(d, a, num_sms, st, stages) : launch_backward_t<__nv_bfloat16, 8>(d, a, num_sms, st, stages);
  return small ? launch_backward_t<float, 4>(d, a, num_sms, st, stages) : launch_backward_t<float, 8>(d, a, num_sms, st, stages);
}

int launch_keep_mask(const DevDesc& d, uint8_t* keep, cudaStream_t st) {
  k_keep_mask<<<592, 256, 0, st>>>(d, keep);
  return check_launch();
}

}  // namespace moka

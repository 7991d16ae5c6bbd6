// k_attn_bwd2: attention backward + operand packing of the fused backward chain, second generation (round 2).
// Included by moka_tc.cu (needs DgParts, AttnPackParams, mma_bf16_16816, split_bf16x2, find_key_slot, dbg stamps).
//
// What was wrong with k_attn_pack (profiles/r01_timeline.md): 32 tokens per CTA, so only 2 of its 8 warps ran the
// tensor-core phase; the per-(token, key) factors made a round trip through shared memory as 2-byte transposed
// stores before the dK phase; and every sample ended in a serial "last CTA packs the parked key tiles" finisher
// (7 us of an 18 us kernel at T = 8192).
//
// Here a CTA owns TT = 64 or 128 consecutive tokens of one sample and EVERY compute warp owns one 16-token m-tile:
//   pass 1 (row form, as before)   S = H K^T, dP = dG K^T per 32-key chunk -> P, dS in registers -> dh += dS K
//   pass 2 (transposed form)       S^T = K H^T, dP^T = K dG^T per chunk (recomputed on the tensor cores: cheaper than
//                                  a shared-memory transpose) -> (P w)^T, (dS/sqrt r)^T are A fragments straight from
//                                  the accumulators (two adjacent token n-tiles = one k-step) ->
//                                  dK[32 keys, r] = (P w)^T dG + (dS/sqrt r)^T H, reduced into the sample's dK with
//                                  red.global.add.v2 (one per element per warp tile: the same number of L2 reductions
//                                  per element as one per 32-token CTA before, no cross-warp stage)
// and the key tiles are finished by DEDICATED CTAs instead of a per-sample finisher: the grid is
// [query-phase CTAs | key-phase CTAs], one of each per tile.  A query-phase CTA never waits; if its tile holds key
// rows it parks its dh rows (fp32, L2) instead of packing them.  The key-phase CTA of a tile without key rows exits
// at once; otherwise it prefetches what it can, spins until the sample's arrival counter shows that every
// query-phase CTA of the sample is done (bounded, traps), folds the finished dK onto the key rows and packs the tile.
// Key-phase CTAs have the HIGHER block indices, so when one of them is resident every query-phase CTA has been
// dispatched already (in-order block dispatch, the same assumption split-K semaphores make) and, since those never
// wait, the spin always ends.
#pragma once

namespace moka {

// dHc rows + dHcT columns (hi / lo) of `nt` tokens from fp32 rows `rows[tt * stride + j]`
__device__ __forceinline__ void pack_rows2(const DevDesc& d, const AttnPackParams& p, const float* __restrict__ rows,
                                           int stride, const int* __restrict__ s_base, size_t tbase, int nt,
                                           __nv_bfloat16* __restrict__ hc, __nv_bfloat16* __restrict__ hct_hi,
                                           __nv_bfloat16* __restrict__ hct_lo, int tid, int nthr) {
  const int r = d.r;
  const int cpr = p.KP / 8;
  if (hc)                                           // (token-major operand: only the round-1 form of k_wgrad_tc reads it)
  for (int i = tid; i < nt * cpr; i += nthr) {
    const int tt = i / cpr, n0 = (i - tt * cpr) * 8;
    const int base = s_base[tt];
    const float* rowp = rows + tt * stride;
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
      const float v = (base >= 0 && j >= 0 && j < r) ? rows[tt * stride + j] * p.scale : 0.f;
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

__device__ __forceinline__ int ld_acquire_i32(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// shared-memory footprint (bytes) of one CTA: must match the carve-up in the kernel
static inline size_t attn_bwd2_smem(int R16, int KB, int TT) {
  const int KP32 = KB * 32, KS = R16 + 8, TS = KP32 + 8, HP = R16 + 8, TP = TT + 8;
  const size_t f32 = (size_t)4 * TT * R16 * 4;
  const size_t bf = ((size_t)2 * KP32 * KS + (size_t)2 * R16 * TS + (size_t)4 * TT * HP + (size_t)4 * R16 * TP) * 2;
  return f32 + bf + 16;
}

template <int R16, int KB, int TT>
__global__ void __launch_bounds__(256)
k_attn_bwd2(DevDesc d, AttnPackParams p, int nq_ctas, const uint8_t* __restrict__ tokmeta,
            const int32_t* __restrict__ kidx, const int32_t* __restrict__ klen, const float* __restrict__ h_stash,
            float* __restrict__ dG0, DgParts parts, float* __restrict__ dK_glob, int* __restrict__ counters,
            __nv_bfloat16* __restrict__ hc, __nv_bfloat16* __restrict__ hct_hi, __nv_bfloat16* __restrict__ hct_lo) {
  constexpr int NT = 256;
  constexpr int KP32 = KB * 32, KS = R16 + 8, TS = KP32 + 8, HP = R16 + 8, TP = TT + 8;
  constexpr int R4 = R16 / 4;
  constexpr int NR = R16 / 8;                       // rank n-tiles
  extern __shared__ __align__(16) uint8_t sm_ab[];
  float* hs = reinterpret_cast<float*>(sm_ab);      // [TT][R16]
  float* gs = hs + TT * R16;
  float* dgs = gs + TT * R16;
  float* dhs = dgs + TT * R16;                      // dh rows: dG + the attention-backward update of query rows
  __nv_bfloat16* Khi = reinterpret_cast<__nv_bfloat16*>(dhs + TT * R16);   // [KP32][KS]
  __nv_bfloat16* Klo = Khi + KP32 * KS;
  __nv_bfloat16* Thi = Klo + KP32 * KS;             // [R16][TS]   K^T
  __nv_bfloat16* Tlo = Thi + R16 * TS;
  __nv_bfloat16* Hh = Tlo + R16 * TS;               // [TT][HP]
  __nv_bfloat16* Hl = Hh + TT * HP;
  __nv_bfloat16* Gh = Hl + TT * HP;
  __nv_bfloat16* Gl = Gh + TT * HP;
  __nv_bfloat16* HTh = Gl + TT * HP;                // [R16][TP]
  __nv_bfloat16* HTl = HTh + R16 * TP;
  __nv_bfloat16* GTh = HTl + R16 * TP;
  __nv_bfloat16* GTl = GTh + R16 * TP;
  __shared__ int s_code[TT];
  __shared__ int s_base[TT];
  __shared__ float s_lse[TT];
  __shared__ float s_D[TT];
  __shared__ int s_np[TT];
  __shared__ int s_kpos[KP32];                      // in-sample position of key k (-1: zero row / beyond L)
  const int r = d.r;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool key_phase = (int)blockIdx.x >= nq_ctas;
  const int tile_id = key_phase ? (int)blockIdx.x - nq_ctas : (int)blockIdx.x;
  const int b = tile_id / p.bps;
  const int s0 = (tile_id % p.bps) * TT;
  const int nt = min(TT, d.S - s0);
  const size_t tbase = (size_t)b * d.S + s0;
  const float* g_stash = h_stash + (size_t)d.T * r;
  const float* lse = g_stash + (size_t)d.T * r;
  const int L = min(klen[b], KP32);
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 110);
  ptx::pdl_trigger();                               // k_wgrad_tc: prologue + x / dx prefetch may start

  int myq = 0, mykey = 0;
  for (int i = tid; i < TT; i += NT) {
    const uint8_t mb = (i < nt) ? tokmeta[tbase + i] : (uint8_t)0;
    const int code = mb & MOKA_MOD_MASK;
    s_code[i] = code | ((mb & MOKA_KEY_VALID) ? 16 : 0);
    s_base[i] = code ? (code - 1) * p.rp : -1;
    s_lse[i] = (code >= MOKA_MOD_VIS) ? lse[tbase + i] : 0.f;
    s_np[i] = dg_nparts(parts, (int)((tbase + i) >> 7));
    myq |= code >= MOKA_MOD_VIS;
    mykey |= ((mb & MOKA_KEY_VALID) && code) ? 1 : 0;
  }
  for (int k = tid; k < KP32; k += NT) {
    int pos = -1;
    if (k < L) pos = kidx[(size_t)b * d.Lk + k];
    s_kpos[k] = pos;
  }
  const int anyq = __syncthreads_or(myq) && L > 0;
  const int anykey = __syncthreads_or(mykey);

  if (key_phase) {
    // =============================================================== key-phase CTA
    if (!anykey) return;
    float* vs = hs;                                  // [TT][R16] parked dh rows (+ dK on the key rows)
    ptx::pdl_wait();                                 // counters / dK zero-filled and dG written by the kernels in front
    if (tid == 0) {
      bool ok = false;
      for (uint32_t it = 0; it < (1u << 22); ++it) {
        if (ld_acquire_i32(&counters[b]) >= p.bps) { ok = true; break; }
        __nanosleep(64);
      }
      if (!ok) asm volatile("trap;");
    }
    __syncthreads();
    for (int i = tid; i < TT * R4; i += NT) {
      const int tt = i / R4, j = (i - tt * R4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tt < nt && j < r) {
        v = __ldcg(reinterpret_cast<const float4*>(dG0 + (tbase + tt) * r + j));
        const int mb = s_code[tt];
        if ((mb & 16) && (mb & MOKA_MOD_MASK)) {
          const int slot = find_key_slot(s_kpos, L, s0 + tt);
          if (slot >= 0) {
            const float4 w = __ldcg(reinterpret_cast<const float4*>(dK_glob + ((size_t)b * d.Lk + slot) * r + j));
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
          }
        }
      }
      *reinterpret_cast<float4*>(vs + tt * R16 + j) = v;
    }
    __syncthreads();
    pack_rows2(d, p, vs, R16, s_base, tbase, nt, hc, hct_hi, hct_lo, tid, NT);
    if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 111);
    return;
  }

  // ================================================================= query-phase CTA
  // ---- everything the forward left behind (older than this call): h, g rows, the sample's key rows
  if (anyq) {
    for (int i = tid; i < TT * R4; i += NT) {
      const int tt = i / R4, j = (i - tt * R4) * 4;
      float4 hv = make_float4(0.f, 0.f, 0.f, 0.f), gv = hv;
      if (tt < nt && j < r) {
        hv = *reinterpret_cast<const float4*>(h_stash + (tbase + tt) * r + j);
        gv = *reinterpret_cast<const float4*>(g_stash + (tbase + tt) * r + j);
      }
      *reinterpret_cast<float4*>(hs + tt * R16 + j) = hv;
      *reinterpret_cast<float4*>(gs + tt * R16 + j) = gv;
      const float f[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(f[e]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(f[e] - __bfloat162float(hi));
        Hh[tt * HP + j + e] = hi; Hl[tt * HP + j + e] = lo;
        HTh[(j + e) * TP + tt] = hi; HTl[(j + e) * TP + tt] = lo;
      }
    }
    for (int base = 0; base < KP32 * R4; base += 4 * NT) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + tid + u * NT;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < KP32 * R4) {
          const int k = idx / R4, j = (idx - k * R4) * 4;
          const int pos = s_kpos[k];
          if (pos >= 0 && j < r) {
            const size_t t = (size_t)b * d.S + pos;
            if (tokmeta[t] & MOKA_KEY_VALID) v[u] = *reinterpret_cast<const float4*>(h_stash + t * r + j);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + tid + u * NT;
        if (idx >= KP32 * R4) continue;
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
  }
  ptx::pdl_wait();                                  // dG from k_dgrad_tc, zeroed dK / counters
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 115);
  for (int i = tid; i < TT * R4; i += NT) {
    const int tt = i / R4, j = (i - tt * R4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt < nt && j < r) {
      const float* src = dG0 + (tbase + tt) * r + j;
      v = *reinterpret_cast<const float4*>(src);
      const int np = s_np[tt];                      // partials of the persistent k_dgrad_tc (fixed order)
      for (int q = 1; q < np; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(src + (size_t)q * (size_t)parts.stride);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
    }
    *reinterpret_cast<float4*>(dgs + tt * R16 + j) = v;
    *reinterpret_cast<float4*>(dhs + tt * R16 + j) = v;
    if (anyq) {
      const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat16 hi = __float2bfloat16_rn(f[e]);
        const __nv_bfloat16 lo = __float2bfloat16_rn(f[e] - __bfloat162float(hi));
        Gh[tt * HP + j + e] = hi; Gl[tt * HP + j + e] = lo;
        GTh[(j + e) * TP + tt] = hi; GTl[(j + e) * TP + tt] = lo;
      }
    }
  }
  __syncthreads();
  if (anyq) {
    // D_t = <p_t, dp_t> = dg_t . (g_t - h_t)
    for (int tt = tid; tt < TT; tt += NT) {
      float D = 0.f;
      for (int j = 0; j < R16; ++j) D = fmaf(dgs[tt * R16 + j], gs[tt * R16 + j] - hs[tt * R16 + j], D);
      s_D[tt] = D;
    }
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    auto ld32 = [](const __nv_bfloat16* q) { return *reinterpret_cast<const uint32_t*>(q); };
    bool any = false;
    const int m0 = warp * 16;
    if (warp < TT / 16)
      for (int i = 0; i < 16; ++i) any |= ((s_code[m0 + i] & MOKA_MOD_MASK) >= MOKA_MOD_VIS);
    if (any) {
      const float sl = d.sm_scale * kLog2e;
      // ---------------------------------------------------------- pass 1: dh += (dS / sqrt r) K
      {
        const int r0 = m0 + g, r1 = m0 + g + 8;
        const bool q0 = (s_code[r0] & MOKA_MOD_MASK) >= MOKA_MOD_VIS, q1 = (s_code[r1] & MOKA_MOD_MASK) >= MOKA_MOD_VIS;
        const float D0 = s_D[r0], D1 = s_D[r1];
        const float lse0 = s_lse[r0], lse1 = s_lse[r1];
        float O[NR][4];
#pragma unroll
        for (int n = 0; n < NR; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
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
              const uint32_t b0h = ld32(Khi + key * KS + c0), b1h = ld32(Khi + key * KS + c0 + 8);
              const uint32_t b0l = ld32(Klo + key * KS + c0), b1l = ld32(Klo + key * KS + c0 + 8);
              mma_bf16_16816(S[n], ah, b0h, b1h); mma_bf16_16816(S[n], al, b0h, b1h); mma_bf16_16816(S[n], ah, b0l, b1l);
              mma_bf16_16816(dP[n], gh4, b0h, b1h); mma_bf16_16816(dP[n], gl4, b0h, b1h); mma_bf16_16816(dP[n], gh4, b0l, b1l);
            }
          }
          // S is overwritten with ds / sqrt(r) (the A operand of the dh product)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = c * 32 + n * 8 + 2 * t + e;
              const bool ok = key < L;
              const float p0 = (ok && q0) ? exp2f(S[n][e] * sl - lse0) : 0.f;
              const float p1 = (ok && q1) ? exp2f(S[n][2 + e] * sl - lse1) : 0.f;
              S[n][e] = p0 * (dP[n][e] * d.attn_w - D0) * d.sm_scale;
              S[n][2 + e] = p1 * (dP[n][2 + e] * d.attn_w - D1) * d.sm_scale;
            }
          }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            uint32_t dh4[4], dl4[4];
            split_bf16x2(S[2 * kk][0], S[2 * kk][1], dh4[0], dl4[0]);
            split_bf16x2(S[2 * kk][2], S[2 * kk][3], dh4[1], dl4[1]);
            split_bf16x2(S[2 * kk + 1][0], S[2 * kk + 1][1], dh4[2], dl4[2]);
            split_bf16x2(S[2 * kk + 1][2], S[2 * kk + 1][3], dh4[3], dl4[3]);
#pragma unroll
            for (int n = 0; n < NR; ++n) {
              const __nv_bfloat16* th = Thi + (n * 8 + g) * TS + c * 32 + kk * 16 + 2 * t;
              const __nv_bfloat16* tl = Tlo + (n * 8 + g) * TS + c * 32 + kk * 16 + 2 * t;
              const uint32_t b0h = ld32(th), b1h = ld32(th + 8), b0l = ld32(tl), b1l = ld32(tl + 8);
              mma_bf16_16816(O[n], dh4, b0h, b1h); mma_bf16_16816(O[n], dl4, b0h, b1h); mma_bf16_16816(O[n], dh4, b0l, b1l);
            }
          }
        }
#pragma unroll
        for (int n = 0; n < NR; ++n) {
          const int col = n * 8 + 2 * t;
          if (col < r) {
            if (q0) { dhs[r0 * R16 + col] += O[n][0]; dhs[r0 * R16 + col + 1] += O[n][1]; }
            if (q1) { dhs[r1 * R16 + col] += O[n][2]; dhs[r1 * R16 + col + 1] += O[n][3]; }
          }
        }
      }
      // ---------------------------------------------------------- pass 2: dK = (P w)^T dG + (dS / sqrt r)^T H
      {
        // this lane's tokens in the transposed accumulators: n-tile n8 (8 tokens), columns 2t, 2t + 1
        float lse4[4], D4[4];
        bool q4[4];
#pragma unroll
        for (int n8 = 0; n8 < 2; ++n8)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int tok = m0 + n8 * 8 + 2 * t + e;
            lse4[n8 * 2 + e] = s_lse[tok];
            D4[n8 * 2 + e] = s_D[tok];
            q4[n8 * 2 + e] = (s_code[tok] & MOKA_MOD_MASK) >= MOKA_MOD_VIS;
          }
#pragma unroll 1
        for (int c = 0; c < KB; ++c) {
          if (c * 32 >= L) break;
          float ST[2][2][4], PT[2][2][4];             // [key m-tile][token n-tile][4]
#pragma unroll
          for (int mk = 0; mk < 2; ++mk)
#pragma unroll
            for (int n8 = 0; n8 < 2; ++n8)
#pragma unroll
              for (int e = 0; e < 4; ++e) { ST[mk][n8][e] = 0.f; PT[mk][n8][e] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < R16 / 16; ++ks) {
            const int c0 = ks * 16 + 2 * t;
            uint32_t bh[2][2], bl[2][2], ch[2][2], cl[2][2];   // B operands: H^T / dG^T, [token n-tile][2]
#pragma unroll
            for (int n8 = 0; n8 < 2; ++n8) {
              const int tok = m0 + n8 * 8 + g;
              bh[n8][0] = ld32(Hh + tok * HP + c0); bh[n8][1] = ld32(Hh + tok * HP + c0 + 8);
              bl[n8][0] = ld32(Hl + tok * HP + c0); bl[n8][1] = ld32(Hl + tok * HP + c0 + 8);
              ch[n8][0] = ld32(Gh + tok * HP + c0); ch[n8][1] = ld32(Gh + tok * HP + c0 + 8);
              cl[n8][0] = ld32(Gl + tok * HP + c0); cl[n8][1] = ld32(Gl + tok * HP + c0 + 8);
            }
#pragma unroll
            for (int mk = 0; mk < 2; ++mk) {
              const int k0 = c * 32 + mk * 16 + g, k1 = k0 + 8;
              uint32_t ah[4], al[4];
              ah[0] = ld32(Khi + k0 * KS + c0); ah[1] = ld32(Khi + k1 * KS + c0);
              ah[2] = ld32(Khi + k0 * KS + c0 + 8); ah[3] = ld32(Khi + k1 * KS + c0 + 8);
              al[0] = ld32(Klo + k0 * KS + c0); al[1] = ld32(Klo + k1 * KS + c0);
              al[2] = ld32(Klo + k0 * KS + c0 + 8); al[3] = ld32(Klo + k1 * KS + c0 + 8);
#pragma unroll
              for (int n8 = 0; n8 < 2; ++n8) {
                mma_bf16_16816(ST[mk][n8], ah, bh[n8][0], bh[n8][1]); mma_bf16_16816(ST[mk][n8], al, bh[n8][0], bh[n8][1]);
                mma_bf16_16816(ST[mk][n8], ah, bl[n8][0], bl[n8][1]);
                mma_bf16_16816(PT[mk][n8], ah, ch[n8][0], ch[n8][1]); mma_bf16_16816(PT[mk][n8], al, ch[n8][0], ch[n8][1]);
                mma_bf16_16816(PT[mk][n8], ah, cl[n8][0], cl[n8][1]);
              }
            }
          }
          // ST <- p w ; PT <- ds / sqrt(r)      element e: key row g (+8 for e >= 2), token column 2t + (e & 1)
#pragma unroll
          for (int mk = 0; mk < 2; ++mk)
#pragma unroll
            for (int n8 = 0; n8 < 2; ++n8)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int key = c * 32 + mk * 16 + g + ((e >> 1) << 3);
                const int ti = n8 * 2 + (e & 1);
                const float pr = (key < L && q4[ti]) ? exp2f(ST[mk][n8][e] * sl - lse4[ti]) : 0.f;
                const float ds = pr * (PT[mk][n8][e] * d.attn_w - D4[ti]) * d.sm_scale;
                ST[mk][n8][e] = pr * d.attn_w;
                PT[mk][n8][e] = ds;
              }
#pragma unroll
          for (int mk = 0; mk < 2; ++mk) {
            if (c * 32 + mk * 16 >= L) continue;      // uniform
            uint32_t pwh[4], pwl[4], dsh[4], dsl[4];   // A fragments [16 keys x 16 tokens]
            split_bf16x2(ST[mk][0][0], ST[mk][0][1], pwh[0], pwl[0]);
            split_bf16x2(ST[mk][0][2], ST[mk][0][3], pwh[1], pwl[1]);
            split_bf16x2(ST[mk][1][0], ST[mk][1][1], pwh[2], pwl[2]);
            split_bf16x2(ST[mk][1][2], ST[mk][1][3], pwh[3], pwl[3]);
            split_bf16x2(PT[mk][0][0], PT[mk][0][1], dsh[0], dsl[0]);
            split_bf16x2(PT[mk][0][2], PT[mk][0][3], dsh[1], dsl[1]);
            split_bf16x2(PT[mk][1][0], PT[mk][1][1], dsh[2], dsl[2]);
            split_bf16x2(PT[mk][1][2], PT[mk][1][3], dsh[3], dsl[3]);
#pragma unroll
            for (int nr = 0; nr < NR; ++nr) {
              float C[4] = {0.f, 0.f, 0.f, 0.f};
              const int nb = (nr * 8 + g) * TP + m0 + 2 * t;
              uint32_t b0h = ld32(GTh + nb), b1h = ld32(GTh + nb + 8), b0l = ld32(GTl + nb), b1l = ld32(GTl + nb + 8);
              mma_bf16_16816(C, pwh, b0h, b1h); mma_bf16_16816(C, pwl, b0h, b1h); mma_bf16_16816(C, pwh, b0l, b1l);
              b0h = ld32(HTh + nb); b1h = ld32(HTh + nb + 8); b0l = ld32(HTl + nb); b1l = ld32(HTl + nb + 8);
              mma_bf16_16816(C, dsh, b0h, b1h); mma_bf16_16816(C, dsl, b0h, b1h); mma_bf16_16816(C, dsh, b0l, b1l);
              const int key0 = c * 32 + mk * 16 + g, key1 = key0 + 8, col = nr * 8 + 2 * t;
              if (col < r) {
                if (key0 < L && (C[0] != 0.f || C[1] != 0.f))
                  red_add_v2_f32(&dK_glob[((size_t)b * d.Lk + key0) * r + col], C[0], C[1]);
                if (key1 < L && (C[2] != 0.f || C[3] != 0.f))
                  red_add_v2_f32(&dK_glob[((size_t)b * d.Lk + key1) * r + col], C[2], C[3]);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) dbg_stamp_g(d.dbg, blockIdx.x, 117);
  if (!anykey) {
    pack_rows2(d, p, dhs, R16, s_base, tbase, nt, hc, hct_hi, hct_lo, tid, NT);
  } else {
    // parked: this tile's key-phase CTA reads the rows back (L2) once the sample's dK is complete
    for (int i = tid; i < TT * R4; i += NT) {
      const int tt = i / R4, j = (i - tt * R4) * 4;
      if (tt < nt && j < r) *reinterpret_cast<float4*>(dG0 + (tbase + tt) * r + j) = *reinterpret_cast<const float4*>(dhs + tt * R16 + j);
    }
  }
  // ---- arrival: dK reductions and parked rows of this CTA are visible before the counter moves
  __threadfence();
  __syncthreads();
  if (tid == 0) { atomicAdd(&counters[b], 1); dbg_stamp_g(d.dbg, blockIdx.x, 111); }
}

}  // namespace moka

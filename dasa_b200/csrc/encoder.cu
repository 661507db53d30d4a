// Encoder-side kernels for DicEncoder / DicModel (r2rmodel.py:2272-2365, vilmodel.py:161-236, 479-506, 1083-1095):
// embedding + LayerNorm, fused dropout/residual/LayerNorm, short-sequence multi-head attention (one CTA per
// (sample, head), everything resident in shared memory), token reversal. The dense projections go through dasa_gemm.
#include <cuda_fp16.h>
#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr int LN_MAXV = 8;  // float4 per lane -> rows up to 1024 wide

struct RowVec { float4 v[LN_MAXV]; };

// mean / rstd over a row held in registers by one warp (two-pass, biased variance: torch.nn.LayerNorm)
__device__ __forceinline__ void warp_row_stats(const RowVec& x, int n4, int Hd, float eps, float& mean, float& rstd) {
  const int lane = threadIdx.x & 31;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (lane + 32 * i < n4) s += (x.v[i].x + x.v[i].y) + (x.v[i].z + x.v[i].w);
  mean = warp_sum(s) / Hd;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (lane + 32 * i < n4) {
      float a;
      a = x.v[i].x - mean; q = fmaf(a, a, q); a = x.v[i].y - mean; q = fmaf(a, a, q);
      a = x.v[i].z - mean; q = fmaf(a, a, q); a = x.v[i].w - mean; q = fmaf(a, a, q);
    }
  rstd = rsqrtf(warp_sum(q) / Hd + eps);
}

__device__ __forceinline__ float4 ln_apply(const float4& x, float mean, float rstd, const float4& g, const float4& b) {
  return make_float4((x.x - mean) * rstd * g.x + b.x, (x.y - mean) * rstd * g.y + b.y, (x.z - mean) * rstd * g.z + b.z,
                     (x.w - mean) * rstd * g.w + b.w);
}

__device__ __forceinline__ float4 mask_f4(const uint8_t* m, float scale) {
  const uchar4 u = *reinterpret_cast<const uchar4*>(m);
  return make_float4(u.x ? scale : 0.f, u.y ? scale : 0.f, u.z ? scale : 0.f, u.w ? scale : 0.f);
}

__global__ void __launch_bounds__(256) embed_layernorm_kernel(const int64_t* __restrict__ ids, int64_t ld_ids, int B, int L, int Hd,
                                                              const float* __restrict__ word, const float* __restrict__ pos,
                                                              const float* __restrict__ type0, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              const uint8_t* __restrict__ mask, float scale, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * L) return;
  const int b = row / L, l = row % L;
  const int64_t id = ids[(int64_t)b * ld_ids + l];
  const int n4 = Hd >> 2;
  RowVec x;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < n4) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(word + id * Hd) + j);
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (int64_t)l * Hd) + j);
      const float4 t = __ldg(reinterpret_cast<const float4*>(type0) + j);
      x.v[i] = make_float4((w.x + p.x) + t.x, (w.y + p.y) + t.y, (w.z + p.z) + t.z, (w.w + p.w) + t.w);
    }
  }
  float mean, rstd;
  warp_row_stats(x, n4, Hd, eps, mean, rstd);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < n4) {
      float4 o = ln_apply(x.v[i], mean, rstd, __ldg(reinterpret_cast<const float4*>(gamma) + j),
                          __ldg(reinterpret_cast<const float4*>(beta) + j));
      if (mask != nullptr) {
        const float4 m = mask_f4(mask + (int64_t)row * Hd + 4 * j, scale);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
      }
      reinterpret_cast<float4*>(out + (int64_t)row * Hd)[j] = o;
    }
  }
}

// In-place dropout draws (rng.cuh) instead of a materialised mask: keep flag of element [row, c] = stream byte row * Hd + c.
struct LnStream { const unsigned long long* seed_dev; unsigned long long seed, base; uint32_t thr; int on; };

// XH: x holds IEEE halves (ldx in halves) - the fp16 output of the preceding dasa_gemm_f16
template <bool XH>
__global__ void __launch_bounds__(128) dropout_residual_layernorm_kernel(
    const void* __restrict__ x_, int64_t ldx, const uint8_t* __restrict__ mask, LnStream st, float scale,
    const float* __restrict__ resid,
    int64_t ldr, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, const uint8_t* __restrict__ post_mask,
    float post_scale, float* __restrict__ out, int64_t ldo, float* __restrict__ stats_out, float* __restrict__ z_out,
    __half* __restrict__ out_half, int R, int Hd) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int n4 = Hd >> 2;
  DropStream ds;
  ds.thr = st.thr; ds.base = st.base;
  ds.mixed = st.on ? mix_seed(st.seed_dev ? st.seed_dev[0] : st.seed) : 0ull;
  RowVec z;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < n4) {
      float4 v;
      if (XH) {
        const uint2 raw = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x_) + (int64_t)row * ldx)[j];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        v = make_float4(lo.x, lo.y, hi.x, hi.y);
      } else {
        v = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x_) + (int64_t)row * ldx)[j];
      }
      if (mask != nullptr) {
        const float4 m = mask_f4(mask + (int64_t)row * Hd + 4 * j, scale);
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
      } else if (st.on) {
        const uint32_t k4 = stream_keep4(ds, (uint64_t)row * n4 + j);
        v.x *= (k4 & 1u) ? scale : 0.f; v.y *= (k4 & 2u) ? scale : 0.f;
        v.z *= (k4 & 4u) ? scale : 0.f; v.w *= (k4 & 8u) ? scale : 0.f;
      }
      if (resid != nullptr) {
        const float4 r = reinterpret_cast<const float4*>(resid + (int64_t)row * ldr)[j];
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      z.v[i] = v;
      if (z_out != nullptr) reinterpret_cast<float4*>(z_out + (int64_t)row * Hd)[j] = v;
    }
  }
  float mean, rstd;
  warp_row_stats(z, n4, Hd, eps, mean, rstd);
  if (stats_out != nullptr && lane == 0) { stats_out[2 * row] = mean; stats_out[2 * row + 1] = rstd; }
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int j = lane + 32 * i;
    if (j < n4) {
      float4 o = ln_apply(z.v[i], mean, rstd, __ldg(reinterpret_cast<const float4*>(gamma) + j),
                          __ldg(reinterpret_cast<const float4*>(beta) + j));
      if (post_mask != nullptr) {
        const float4 m = mask_f4(post_mask + (int64_t)row * Hd + 4 * j, post_scale);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
      }
      reinterpret_cast<float4*>(out + (int64_t)row * ldo)[j] = o;
      if (out_half != nullptr) {         // fp16 copy [R, Hd] contiguous: the A operand of the next fp16 GEMM
        __half2 h[2] = {__floats2half2_rn(o.x, o.y), __floats2half2_rn(o.z, o.w)};
        reinterpret_cast<uint2*>(out_half + (int64_t)row * Hd)[j] = *reinterpret_cast<uint2*>(h);
      }
    }
  }
}

// Forward-only fast path of the frozen stack for Hd = 128 * NV (768 -> NV = 6): exact vector count (no per-vector guards),
// every load of the row requested before the first use, dropout + residual as one FMA per element, the normalisation as
// two FMAs per element ((x * rstd - mean * rstd) * gamma + beta), hash words advanced by addition (rng.cuh: (w + 1) * G).
template <bool XH, int NV>
__global__ void __launch_bounds__(128) ln_fwd_fast_kernel(
    const void* __restrict__ x_, int64_t ldx, const uint8_t* __restrict__ mask, LnStream st, float scale,
    const float* __restrict__ resid, int64_t ldr, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    float* __restrict__ out, int64_t ldo, __half* __restrict__ out_half, int R) {
  constexpr int Hd = 128 * NV, n4 = 32 * NV;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  float4 v[NV], r[NV];
  uint32_t mk[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    if (XH) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(x_) + (int64_t)row * ldx) + j);
      const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      v[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
      v[i] = ldg_stream4(reinterpret_cast<const float*>(x_) + (int64_t)row * ldx + 4 * j);
    }
    r[i] = resid != nullptr ? ldg_stream4(resid + (int64_t)row * ldr + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (mask != nullptr) mk[i] = *reinterpret_cast<const uint32_t*>(mask + (int64_t)row * Hd + 4 * j);
  }
  if (mask != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x = fmaf(v[i].x, (mk[i] & 0xFFu) ? scale : 0.f, r[i].x);
      v[i].y = fmaf(v[i].y, (mk[i] & 0xFF00u) ? scale : 0.f, r[i].y);
      v[i].z = fmaf(v[i].z, (mk[i] & 0xFF0000u) ? scale : 0.f, r[i].z);
      v[i].w = fmaf(v[i].w, (mk[i] & 0xFF000000u) ? scale : 0.f, r[i].w);
    }
  } else if (st.on) {
    const uint64_t mixed = mix_seed(st.seed_dev ? st.seed_dev[0] : st.seed);
    uint64_t zg = (st.base + (uint64_t)row * n4 + lane + 1) * 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint64_t z = mixed ^ zg;
      zg += 32ull * 0x9E3779B97F4A7C15ull;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      const uint32_t lo = (uint32_t)z, hi = (uint32_t)(z >> 32);
      v[i].x = fmaf(v[i].x, ((lo & 0xFFFFu) >= st.thr) ? scale : 0.f, r[i].x);
      v[i].y = fmaf(v[i].y, ((lo >> 16) >= st.thr) ? scale : 0.f, r[i].y);
      v[i].z = fmaf(v[i].z, ((hi & 0xFFFFu) >= st.thr) ? scale : 0.f, r[i].z);
      v[i].w = fmaf(v[i].w, ((hi >> 16) >= st.thr) ? scale : 0.f, r[i].w);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i].x += r[i].x; v[i].y += r[i].y; v[i].z += r[i].z; v[i].w += r[i].w; }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.f / Hd);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a;
    a = v[i].x - mean; q = fmaf(a, a, q); a = v[i].y - mean; q = fmaf(a, a, q);
    a = v[i].z - mean; q = fmaf(a, a, q); a = v[i].w - mean; q = fmaf(a, a, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / Hd) + eps);
  const float nm = -mean * rstd;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = lane + 32 * i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j), b = __ldg(reinterpret_cast<const float4*>(beta) + j);
    float4 o;
    o.x = fmaf(fmaf(v[i].x, rstd, nm), g.x, b.x); o.y = fmaf(fmaf(v[i].y, rstd, nm), g.y, b.y);
    o.z = fmaf(fmaf(v[i].z, rstd, nm), g.z, b.z); o.w = fmaf(fmaf(v[i].w, rstd, nm), g.w, b.w);
    stg_stream4(out + (int64_t)row * ldo + 4 * j, o);
    if (out_half != nullptr) {
      __half2 h[2] = {__floats2half2_rn(o.x, o.y), __floats2half2_rn(o.z, o.w)};
      reinterpret_cast<uint2*>(out_half + (int64_t)row * Hd)[j] = *reinterpret_cast<uint2*>(h);
    }
  }
}

// dz = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)); dgamma += sum_r g*xhat; dbeta += sum_r g
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dout, int64_t lddo, const float* __restrict__ z,
                                                            const float* __restrict__ gamma, const float* __restrict__ stats,
                                                            const uint8_t* __restrict__ mask, float scale,
                                                            const uint8_t* __restrict__ post_mask, float post_scale,
                                                            float* __restrict__ dx, int64_t lddx, float* __restrict__ dresid,
                                                            int64_t lddr, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                            int R, int Hd) {
  extern __shared__ float sm_acc[];  // [2*Hd] per-CTA dgamma/dbeta partials
  for (int i = threadIdx.x; i < 2 * Hd; i += blockDim.x) sm_acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int n4 = Hd >> 2;
  for (int row = blockIdx.x * nwarp + (threadIdx.x >> 5); row < R; row += gridDim.x * nwarp) {
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    RowVec g, xh;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        float4 d = reinterpret_cast<const float4*>(dout + (int64_t)row * lddo)[j];
        if (post_mask != nullptr) {
          const float4 m = mask_f4(post_mask + (int64_t)row * Hd + 4 * j, post_scale);
          d.x *= m.x; d.y *= m.y; d.z *= m.z; d.w *= m.w;
        }
        const float4 zz = reinterpret_cast<const float4*>(z + (int64_t)row * Hd)[j];
        const float4 xhat = make_float4((zz.x - mean) * rstd, (zz.y - mean) * rstd, (zz.z - mean) * rstd, (zz.w - mean) * rstd);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + j);
        atomicAdd(&sm_acc[4 * j + 0], d.x * xhat.x); atomicAdd(&sm_acc[4 * j + 1], d.y * xhat.y);
        atomicAdd(&sm_acc[4 * j + 2], d.z * xhat.z); atomicAdd(&sm_acc[4 * j + 3], d.w * xhat.w);
        atomicAdd(&sm_acc[Hd + 4 * j + 0], d.x); atomicAdd(&sm_acc[Hd + 4 * j + 1], d.y);
        atomicAdd(&sm_acc[Hd + 4 * j + 2], d.z); atomicAdd(&sm_acc[Hd + 4 * j + 3], d.w);
        const float4 dg = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
        g.v[i] = dg; xh.v[i] = xhat;
        s1 += (dg.x + dg.y) + (dg.z + dg.w);
        s2 += (dg.x * xhat.x + dg.y * xhat.y) + (dg.z * xhat.z + dg.w * xhat.w);
      }
    }
    s1 = warp_sum(s1) / Hd;
    s2 = warp_sum(s2) / Hd;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int j = lane + 32 * i;
      if (j < n4) {
        float4 dz = make_float4(rstd * (g.v[i].x - s1 - xh.v[i].x * s2), rstd * (g.v[i].y - s1 - xh.v[i].y * s2),
                                rstd * (g.v[i].z - s1 - xh.v[i].z * s2), rstd * (g.v[i].w - s1 - xh.v[i].w * s2));
        if (dresid != nullptr) reinterpret_cast<float4*>(dresid + (int64_t)row * lddr)[j] = dz;
        if (dx != nullptr) {
          if (mask != nullptr) {
            const float4 m = mask_f4(mask + (int64_t)row * Hd + 4 * j, scale);
            dz.x *= m.x; dz.y *= m.y; dz.z *= m.z; dz.w *= m.w;
          }
          reinterpret_cast<float4*>(dx + (int64_t)row * lddx)[j] = dz;
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Hd; i += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + i, sm_acc[i]);
    if (dbeta) atomicAdd(dbeta + i, sm_acc[Hd + i]);
  }
}

// ---------------------------------------------------------------------------------------- multi-head attention
struct MhaArgs {
  const float *q, *k, *v; int64_t ldq, sq, ldk, sk, ldv, sv;
  const uint8_t* key_pad; int64_t ld_pad; const uint8_t* drop_mask; float drop_scale;
  float* out; int64_t ldo, so; float* probs_out;
  int B, heads, Lq, Lk, dh;
  int out_half;      // `out` holds IEEE halves (ldo / so in halves): the A operand of the fp16 output projection
  // packed (variable-length) operands: rows of sample b start at row q_off[b] / k_off[b] of the packed matrix and there are
  // q_len[b] / k_len[b] of them; Lq / Lk are then the maxima (shared-memory carve-up, mask / probs indexing). NULL = dense.
  const int32_t *q_off, *q_len, *k_off, *k_len;
};

// Register-tiled: a warp owns 4 query rows; lane l owns keys l, l+32, l+64 (scores) / head-dim columns l, l+32 (P.V).
// Q/K rows are padded to dh+4 floats so 128-bit shared loads are conflict-free per quarter-warp.
__global__ void __launch_bounds__(256) mha_fwd_kernel(MhaArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int Lq = a.q_len ? a.q_len[b] : a.Lq, Lk = a.k_len ? a.k_len[b] : a.Lk, dh = a.dh, dhp = dh + 4;
  const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
  const int LqM = (a.Lq + 3) & ~3, LkM = (a.Lk + 3) & ~3;     // carve-up by the maxima
  float* Qs = smem;                      // [LqP][dhp]
  float* Ks = Qs + LqM * dhp;            // [LkP][dhp]
  float* Vs = Ks + LkM * dhp;            // [LkP][dh]
  float* Ss = Vs + LkM * dh;             // [LqP][LkP]
  const float* qb = a.q + (a.q_off ? (int64_t)a.q_off[b] * a.ldq : (int64_t)b * a.sq) + h * dh;
  const float* kb = a.k + (a.k_off ? (int64_t)a.k_off[b] * a.ldk : (int64_t)b * a.sk) + h * dh;
  const float* vb = a.v + (a.k_off ? (int64_t)a.k_off[b] * a.ldv : (int64_t)b * a.sv) + h * dh;
  if (Lq <= 0) return;
  const int d4n = dh >> 2;
  for (int i = threadIdx.x; i < LqP * d4n; i += blockDim.x) {
    const int r = i / d4n, c = i % d4n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < Lq) v = *reinterpret_cast<const float4*>(qb + (int64_t)r * a.ldq + 4 * c);
    *reinterpret_cast<float4*>(Qs + r * dhp + 4 * c) = v;
  }
  for (int i = threadIdx.x; i < LkP * d4n; i += blockDim.x) {
    const int r = i / d4n, c = i % d4n;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (r < Lk) {
      kv = *reinterpret_cast<const float4*>(kb + (int64_t)r * a.ldk + 4 * c);
      vv = *reinterpret_cast<const float4*>(vb + (int64_t)r * a.ldv + 4 * c);
    }
    *reinterpret_cast<float4*>(Ks + r * dhp + 4 * c) = kv;
    *reinterpret_cast<float4*>(Vs + r * dh + 4 * c) = vv;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float scale = rsqrtf((float)dh);
  constexpr int JT = 3;                  // keys per lane (Lk <= 96)
  for (int i0 = wid * 4; i0 < Lq; i0 += nw * 4) {
    float acc[4][JT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int t = 0; t < JT; ++t) acc[r][t] = 0.f;
    for (int c = 0; c < d4n; ++c) {
      float4 q[4], k[JT];
#pragma unroll
      for (int r = 0; r < 4; ++r) q[r] = *reinterpret_cast<const float4*>(Qs + (i0 + r) * dhp + 4 * c);
#pragma unroll
      for (int t = 0; t < JT; ++t) {
        const int j = lane + 32 * t;
        k[t] = (j < LkP) ? *reinterpret_cast<const float4*>(Ks + j * dhp + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < JT; ++t) {
          acc[r][t] = fmaf(q[r].x, k[t].x, acc[r][t]);
          acc[r][t] = fmaf(q[r].y, k[t].y, acc[r][t]);
          acc[r][t] = fmaf(q[r].z, k[t].z, acc[r][t]);
          acc[r][t] = fmaf(q[r].w, k[t].w, acc[r][t]);
        }
    }
    // softmax over keys for the 4 rows (scores / sqrt(dh) + additive -10000 padding mask, vilmodel.py:219-225)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + r;
      float v[JT];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < JT; ++t) {
        const int j = lane + 32 * t;
        float x = -INFINITY;
        if (j < Lk) {
          x = acc[r][t] * scale;
          if (a.key_pad != nullptr && a.key_pad[(int64_t)b * a.ld_pad + j]) x += -10000.0f;
        }
        v[t] = x;
        mx = fmaxf(mx, x);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int t = 0; t < JT; ++t) {
        v[t] = (lane + 32 * t < Lk) ? expf(v[t] - mx) : 0.f;
        sum += v[t];
      }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      if (i < Lq) {
#pragma unroll
        for (int t = 0; t < JT; ++t) {
          const int j = lane + 32 * t;
          if (j < LkP) {
            float p = 0.f;
            if (j < Lk) {
              p = v[t] * inv;
              const int64_t gi = (((int64_t)b * a.heads + h) * a.Lq + i) * a.Lk + j;     // padded-shape mask / probs
              if (a.probs_out != nullptr) a.probs_out[gi] = p;
              if (a.drop_mask != nullptr) p *= a.drop_mask[gi] ? a.drop_scale : 0.f;
            }
            Ss[i * LkP + j] = p;
          }
        }
      }
    }
    __syncwarp();
    // O[i0..i0+3][d] = P V ; lane owns head-dim columns lane, lane+32
    float o[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) { o[r][0] = 0.f; o[r][1] = 0.f; }
    for (int j4 = 0; j4 < LkP; j4 += 4) {
      float4 p[4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
        p[r] = (i0 + r < Lq) ? *reinterpret_cast<const float4*>(Ss + (i0 + r) * LkP + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int d = lane + 32 * c;
        if (d < dh) {
          const float v0 = Vs[(j4 + 0) * dh + d], v1 = Vs[(j4 + 1) * dh + d], v2 = Vs[(j4 + 2) * dh + d], v3 = Vs[(j4 + 3) * dh + d];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            o[r][c] = fmaf(p[r].x, v0, o[r][c]);
            o[r][c] = fmaf(p[r].y, v1, o[r][c]);
            o[r][c] = fmaf(p[r].z, v2, o[r][c]);
            o[r][c] = fmaf(p[r].w, v3, o[r][c]);
          }
        }
      }
    }
    float* ob = a.out + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh;
#pragma unroll
    for (int r = 0; r < 4; ++r)
      if (i0 + r < Lq) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int d = lane + 32 * c;
          if (d < dh) {
            if (a.out_half) reinterpret_cast<__half*>(a.out)[(a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh +
                                                             (int64_t)(i0 + r) * a.ldo + d] = __float2half_rn(o[r][c]);
            else ob[(int64_t)(i0 + r) * a.ldo + d] = o[r][c];
          }
        }
      }
  }
}

// ---- tensor-core variant (precision mode tf32): S = Q K^T and O = P V on mma.sync.m16n8k8 TF32, fp32 accumulate ----------
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int MHA_TC_THREADS = 128;
constexpr int MHA_TC_MAXNT = 12;     // key tiles of 8 -> Lk <= 96

__host__ __device__ inline int mha_tc_pstride(int LkP) { return LkP + ((36 - (LkP & 31)) & 31); }   // stride % 32 == 4

// one CTA per (sample, head); warp w owns query m-tiles w, w+nwarps, ... (16 rows each). NT = compile-time bound on the key
// tiles (6 for Lk <= 48, else 12): the unrolled softmax over unused tiles would still issue (predicated off).
template <int NT>
__global__ void __launch_bounds__(MHA_TC_THREADS) mha_fwd_tc_kernel(MhaArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int Lq = a.q_len ? a.q_len[b] : a.Lq, Lk = a.k_len ? a.k_len[b] : a.Lk, dh = a.dh;
  if (Lq <= 0) return;
  const int LqP = (Lq + 15) & ~15, LkP = (Lk + 7) & ~7;
  const int LqM = (a.Lq + 15) & ~15, LkM = (a.Lk + 7) & ~7;  // carve-up by the maxima
  const int QS = dh + 4, VS = dh + 8, PS = mha_tc_pstride(LkM);
  uint32_t* Qs = reinterpret_cast<uint32_t*>(smem);          // [LqP][QS]  tf32 bit patterns
  uint32_t* Ks = Qs + LqM * QS;                              // [LkP][QS]
  uint32_t* Vs = Ks + LkM * QS;                              // [LkP][VS]
  uint32_t* Ps = Vs + LkM * VS;                              // [LqP][PS]
  const float* qb = a.q + (a.q_off ? (int64_t)a.q_off[b] * a.ldq : (int64_t)b * a.sq) + h * dh;
  const float* kb = a.k + (a.k_off ? (int64_t)a.k_off[b] * a.ldk : (int64_t)b * a.sk) + h * dh;
  const float* vb = a.v + (a.k_off ? (int64_t)a.k_off[b] * a.ldv : (int64_t)b * a.sv) + h * dh;
  const int d4n = dh >> 2;
  // stage Q, K, V of this (sample, head): one fused loop, three independent streaming loads per iteration and the loop
  // unrolled, so ~9 x 16 B per thread are in flight (the separate, rolled loops exposed the full DRAM latency per row)
  const int nthr = blockDim.x;
  const int nit = (max(LqP, LkP) * d4n + nthr - 1) / nthr;
#pragma unroll 3
  for (int it = 0; it < nit; ++it) {
    const int i = threadIdx.x + it * nthr;
    const int r = i / d4n, c = i % d4n;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv;
    if (r < Lq) qv = ldg_stream4(qb + (int64_t)r * a.ldq + 4 * c);
    if (r < Lk) {
      kv = ldg_stream4(kb + (int64_t)r * a.ldk + 4 * c);
      vv = ldg_stream4(vb + (int64_t)r * a.ldv + 4 * c);
    }
    if (r < LqP) *reinterpret_cast<uint4*>(Qs + r * QS + 4 * c) = make_uint4(f2tf32(qv.x), f2tf32(qv.y), f2tf32(qv.z), f2tf32(qv.w));
    if (r < LkP) {
      *reinterpret_cast<uint4*>(Ks + r * QS + 4 * c) = make_uint4(f2tf32(kv.x), f2tf32(kv.y), f2tf32(kv.z), f2tf32(kv.w));
      *reinterpret_cast<uint4*>(Vs + r * VS + 4 * c) = make_uint4(f2tf32(vv.x), f2tf32(vv.y), f2tf32(vv.z), f2tf32(vv.w));
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nkt = LkP >> 3;                                  // key tiles
  const float scale = rsqrtf((float)dh);
  const uint8_t* pad = a.key_pad ? a.key_pad + (int64_t)b * a.ld_pad : nullptr;
  for (int mt = warp; mt * 16 < Lq; mt += (nthr >> 5)) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
    for (int k0 = 0; k0 < dh; k0 += 8) {
      const uint32_t a0 = Qs[r0 * QS + k0 + t], a1 = Qs[r1 * QS + k0 + t], a2 = Qs[r0 * QS + k0 + t + 4], a3 = Qs[r1 * QS + k0 + t + 4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt < nkt) {
          const uint32_t b0 = Ks[(nt * 8 + g) * QS + k0 + t], b1 = Ks[(nt * 8 + g) * QS + k0 + t + 4];
          mma_m16n8k8_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    // scores -> masked softmax per row; a row lives in the 4 lanes sharing g (cols 2t, 2t+1 of every key tile)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          float add = 0.f;
          const bool ok = j < Lk;
          if (ok && pad != nullptr && pad[j]) add = -10000.0f;
          acc[nt][e] = ok ? acc[nt][e] * scale + add : -INFINITY;
          acc[nt][2 + e] = ok ? acc[nt][2 + e] * scale + add : -INFINITY;
          mx0 = fmaxf(mx0, acc[nt][e]);
          mx1 = fmaxf(mx1, acc[nt][2 + e]);
        }
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          acc[nt][e] = (acc[nt][e] == -INFINITY) ? 0.f : __expf(acc[nt][e] - mx0);
          acc[nt][2 + e] = (acc[nt][2 + e] == -INFINITY) ? 0.f : __expf(acc[nt][2 + e] - mx1);
          s0 += acc[nt][e];
          s1 += acc[nt][2 + e];
        }
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float inv0 = 1.f / s0, inv1 = 1.f / s1;
    const int64_t pb0 = (((int64_t)b * a.heads + h) * a.Lq + r0) * a.Lk, pb1 = (((int64_t)b * a.heads + h) * a.Lq + r1) * a.Lk;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          float p0 = acc[nt][e] * inv0, p1 = acc[nt][2 + e] * inv1;
          if (j < Lk) {
            if (r0 < Lq) {
              if (a.probs_out != nullptr) a.probs_out[pb0 + j] = p0;
              if (a.drop_mask != nullptr) p0 *= a.drop_mask[pb0 + j] ? a.drop_scale : 0.f;
            }
            if (r1 < Lq) {
              if (a.probs_out != nullptr) a.probs_out[pb1 + j] = p1;
              if (a.drop_mask != nullptr) p1 *= a.drop_mask[pb1 + j] ? a.drop_scale : 0.f;
            }
          } else { p0 = 0.f; p1 = 0.f; }
          Ps[r0 * PS + j] = f2tf32(p0);
          Ps[r1 * PS + j] = f2tf32(p1);
        }
      }
    }
    __syncwarp();
    // O = P V for this m-tile: 8 head-dim tiles of 8
    float o[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    for (int k0 = 0; k0 < LkP; k0 += 8) {
      const uint32_t a0 = Ps[r0 * PS + k0 + t], a1 = Ps[r1 * PS + k0 + t], a2 = Ps[r0 * PS + k0 + t + 4], a3 = Ps[r1 * PS + k0 + t + 4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt * 8 < dh) {
          const uint32_t b0 = Vs[(k0 + t) * VS + nt * 8 + g], b1 = Vs[(k0 + t + 4) * VS + nt * 8 + g];
          mma_m16n8k8_tf32(o[nt], a0, a1, a2, a3, b0, b1);
        }
      }
    }
    float* ob = a.out + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt * 8 < dh) {
        const int d = nt * 8 + 2 * t;
        if (a.out_half) {
          __half* oh = reinterpret_cast<__half*>(a.out) + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh;
          if (r0 < Lq) *reinterpret_cast<__half2*>(oh + (int64_t)r0 * a.ldo + d) = __floats2half2_rn(o[nt][0], o[nt][1]);
          if (r1 < Lq) *reinterpret_cast<__half2*>(oh + (int64_t)r1 * a.ldo + d) = __floats2half2_rn(o[nt][2], o[nt][3]);
        } else {
          if (r0 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r0 * a.ldo + d) = make_float2(o[nt][0], o[nt][1]);
          if (r1 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r1 * a.ldo + d) = make_float2(o[nt][2], o[nt][3]);
        }
      }
    }
  }
}

// dh == 64 variant with only K and V in shared memory. Q never touches shared memory: a lane loads the float4s of its two
// query rows straight from global memory and uses component j in MMA step j (any bijection of the reduction index is a valid
// dot product as long as both operands use it: "MMA k index t / t+4 of step j" = head-dim 32c + 4t + j / 32c + 16 + 4t + j).
// The softmax probabilities never touch shared memory either: the score accumulator fragment (row g, keys 2t, 2t+1) is reused
// directly as the A fragment of P.V by reading V rows 2t / 2t+1 for MMA k indices t / t+4. Shared memory per CTA drops from
// 93 KB to 47 KB at 80 keys (23 KB at 36 views): 4-9 CTAs per SM instead of 2, which is what this latency-bound kernel needs.
constexpr int MHA2_KS = 80;          // K row stride (floats): == 16 mod 32 -> conflict-free 128-bit fragment loads
constexpr int MHA2_VS = 68;          // V row stride: 2*VS == 8 mod 32 -> rows 2t (t = 0..3) land in distinct bank groups

template <int NT>
__global__ void __launch_bounds__(MHA_TC_THREADS, (NT >= 10 ? 3 : 4)) mha_fwd_tc64_kernel(MhaArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int dh = 64;
  const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int Lq = a.q_len ? a.q_len[b] : a.Lq, Lk = a.k_len ? a.k_len[b] : a.Lk;
  if (Lq <= 0) return;
  const int LkP = (Lk + 7) & ~7, LkM = (a.Lk + 7) & ~7;
  uint32_t* Ks = reinterpret_cast<uint32_t*>(smem);          // [LkP][KS] tf32 bit patterns
  uint32_t* Vs = Ks + LkM * MHA2_KS;                         // [LkP][VS]
  const float* qb = a.q + (a.q_off ? (int64_t)a.q_off[b] * a.ldq : (int64_t)b * a.sq) + h * dh;
  const float* kb = a.k + (a.k_off ? (int64_t)a.k_off[b] * a.ldk : (int64_t)b * a.sk) + h * dh;
  const float* vb = a.v + (a.k_off ? (int64_t)a.k_off[b] * a.ldv : (int64_t)b * a.sv) + h * dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int nwarps = blockDim.x >> 5;
  // the first query tile of this warp is requested BEFORE the K / V staging so that all three streams are in flight together
  float4 qf[2][2][2];                                        // [row r0 / r1][32-wide chunk][k half]
  auto load_q = [&](int mt) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int col = 32 * c + 16 * hf + 4 * t;
        qf[0][c][hf] = (r0 < Lq) ? ldg_stream4(qb + (int64_t)r0 * a.ldq + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        qf[1][c][hf] = (r1 < Lq) ? ldg_stream4(qb + (int64_t)r1 * a.ldq + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };
  // keep bits of the probability dropout mask for this lane's (row, key) pairs: bit 4*nt + 2*e + row. Requested together with Q,
  // i.e. before the K / V staging barrier, so their latency is off the softmax -> P.V critical path.
  const int nkt_early = LkP >> 3;
  uint64_t keep = ~0ull;
  auto load_mask = [&](int mt) {
    keep = ~0ull;
    if (a.drop_mask == nullptr) return;
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const uint8_t* m0 = a.drop_mask + (((int64_t)b * a.heads + h) * a.Lq + r0) * a.Lk;
    const uint8_t* m1 = m0 + (int64_t)8 * a.Lk;
    uint64_t bits = 0;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt_early) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          const bool in = j < Lk;
          const uint64_t k0 = (in && r0 < Lq) ? (m0[j] != 0) : 1;
          const uint64_t k1 = (in && r1 < Lq) ? (m1[j] != 0) : 1;
          bits |= (k0 << (4 * nt + 2 * e)) | (k1 << (4 * nt + 2 * e + 1));
        }
      }
    }
    keep = bits;
  };
  if (warp * 16 < Lq) { load_q(warp); load_mask(warp); }
  const int nit = (LkP * 16 + blockDim.x - 1) / blockDim.x;
#pragma unroll 4
  for (int it = 0; it < nit; ++it) {
    const int i = threadIdx.x + it * blockDim.x;
    const int r = i >> 4, c = i & 15;
    if (r < LkP) {
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (r < Lk) {
        kv = ldg_stream4(kb + (int64_t)r * a.ldk + 4 * c);
        vv = ldg_stream4(vb + (int64_t)r * a.ldv + 4 * c);
      }
      *reinterpret_cast<uint4*>(Ks + r * MHA2_KS + 4 * c) = make_uint4(f2tf32(kv.x), f2tf32(kv.y), f2tf32(kv.z), f2tf32(kv.w));
      *reinterpret_cast<uint4*>(Vs + r * MHA2_VS + 4 * c) = make_uint4(f2tf32(vv.x), f2tf32(vv.y), f2tf32(vv.z), f2tf32(vv.w));
    }
  }
  __syncthreads();
  const int nkt = LkP >> 3;
  const float scale = 0.125f;                                // 1 / sqrt(64)
  const uint8_t* pad = a.key_pad ? a.key_pad + (int64_t)b * a.ld_pad : nullptr;
  for (int mt = warp; mt * 16 < Lq; mt += nwarps) {
    if (mt != warp) { load_q(mt); load_mask(mt); }
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t qa[4][4];                                     // [step j][a0..a3]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        qa[j][0] = f2tf32(reinterpret_cast<const float*>(&qf[0][c][0])[j]);
        qa[j][1] = f2tf32(reinterpret_cast<const float*>(&qf[1][c][0])[j]);
        qa[j][2] = f2tf32(reinterpret_cast<const float*>(&qf[0][c][1])[j]);
        qa[j][3] = f2tf32(reinterpret_cast<const float*>(&qf[1][c][1])[j]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt < nkt) {
          const uint32_t* krow = Ks + (nt * 8 + g) * MHA2_KS + 32 * c + 4 * t;
          const uint4 k0 = *reinterpret_cast<const uint4*>(krow);
          const uint4 k1 = *reinterpret_cast<const uint4*>(krow + 16);
          mma_m16n8k8_tf32(acc[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], k0.x, k1.x);
          mma_m16n8k8_tf32(acc[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], k0.y, k1.y);
          mma_m16n8k8_tf32(acc[nt], qa[2][0], qa[2][1], qa[2][2], qa[2][3], k0.z, k1.z);
          mma_m16n8k8_tf32(acc[nt], qa[3][0], qa[3][1], qa[3][2], qa[3][3], k0.w, k1.w);
        }
      }
    }
    // scores -> masked softmax per row; a row lives in the 4 lanes sharing g (cols 2t, 2t+1 of every key tile)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          float add = 0.f;
          const bool ok = j < Lk;
          if (ok && pad != nullptr && pad[j]) add = -10000.0f;
          acc[nt][e] = ok ? acc[nt][e] * scale + add : -INFINITY;
          acc[nt][2 + e] = ok ? acc[nt][2 + e] * scale + add : -INFINITY;
          mx0 = fmaxf(mx0, acc[nt][e]);
          mx1 = fmaxf(mx1, acc[nt][2 + e]);
        }
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          acc[nt][e] = (acc[nt][e] == -INFINITY) ? 0.f : __expf(acc[nt][e] - mx0);
          acc[nt][2 + e] = (acc[nt][2 + e] == -INFINITY) ? 0.f : __expf(acc[nt][2 + e] - mx1);
          s0 += acc[nt][e];
          s1 += acc[nt][2 + e];
        }
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float inv0 = 1.f / s0, inv1 = 1.f / s1;
    const int64_t pb0 = (((int64_t)b * a.heads + h) * a.Lq + r0) * a.Lk, pb1 = (((int64_t)b * a.heads + h) * a.Lq + r1) * a.Lk;
    float o[8][4];
#pragma unroll
    for (int n8 = 0; n8 < 8; ++n8) { o[n8][0] = o[n8][1] = o[n8][2] = o[n8][3] = 0.f; }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nkt) {
        uint32_t pa[4];                                      // A fragment of P.V: k index t <-> key 2t, k index t+4 <-> key 2t+1
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          float p0 = acc[nt][e] * inv0, p1 = acc[nt][2 + e] * inv1;
          if (j < Lk) {
            if (r0 < Lq) {
              if (a.probs_out != nullptr) a.probs_out[pb0 + j] = p0;
              if (a.drop_mask != nullptr) p0 *= ((keep >> (4 * nt + 2 * e)) & 1) ? a.drop_scale : 0.f;
            }
            if (r1 < Lq) {
              if (a.probs_out != nullptr) a.probs_out[pb1 + j] = p1;
              if (a.drop_mask != nullptr) p1 *= ((keep >> (4 * nt + 2 * e + 1)) & 1) ? a.drop_scale : 0.f;
            }
          } else { p0 = 0.f; p1 = 0.f; }
          pa[2 * e] = f2tf32(p0);                            // e = 0 -> a0 (row g, k t);   e = 1 -> a2 (row g, k t+4)
          pa[2 * e + 1] = f2tf32(p1);                        // e = 0 -> a1 (row g+8, k t); e = 1 -> a3 (row g+8, k t+4)
        }
        const uint32_t* v0 = Vs + (nt * 8 + 2 * t) * MHA2_VS + g;
#pragma unroll
        for (int n8 = 0; n8 < 8; ++n8) mma_m16n8k8_tf32(o[n8], pa[0], pa[1], pa[2], pa[3], v0[n8 * 8], v0[MHA2_VS + n8 * 8]);
      }
    }
    float* ob = a.out + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh;
#pragma unroll
    for (int n8 = 0; n8 < 8; ++n8) {
      const int d = n8 * 8 + 2 * t;
      if (a.out_half) {
        __half* oh = reinterpret_cast<__half*>(a.out) + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * dh;
        if (r0 < Lq) *reinterpret_cast<__half2*>(oh + (int64_t)r0 * a.ldo + d) = __floats2half2_rn(o[n8][0], o[n8][1]);
        if (r1 < Lq) *reinterpret_cast<__half2*>(oh + (int64_t)r1 * a.ldo + d) = __floats2half2_rn(o[n8][2], o[n8][3]);
      } else {
        if (r0 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r0 * a.ldo + d) = make_float2(o[n8][0], o[n8][1]);
        if (r1 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r1 * a.ldo + d) = make_float2(o[n8][2], o[n8][3]);
      }
    }
  }
}

struct MhaBwdArgs {
  const float *q, *k, *v; int64_t ldq, sq, ldk, sk, ldv, sv;
  const float* probs; const uint8_t* drop_mask; float drop_scale;
  const float* dout; int64_t ldo, so;
  float *dq, *dk, *dv; int64_t lddq, sdq, lddk, sdk, lddv, sdv;
  int B, heads, Lq, Lk, dh;
};

__global__ void __launch_bounds__(256) mha_bwd_kernel(MhaBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.x / a.heads, h = blockIdx.x % a.heads;
  const int Lq = a.Lq, Lk = a.Lk, dh = a.dh, dhp = dh + 1;
  float* Qs = smem;                     // [Lq][dh+1]
  float* Ks = Qs + Lq * dhp;            // [Lk][dh+1]
  float* Vs = Ks + Lk * dhp;            // [Lk][dh+1]
  float* Os = Vs + Lk * dhp;            // dOut [Lq][dh+1]
  float* Ps = Os + Lq * dhp;            // [Lq][Lk]  softmax probs -> dS
  float* Pd = Ps + Lq * Lk;             // [Lq][Lk]  dropped probs
  const float* qb = a.q + (int64_t)b * a.sq + h * dh;
  const float* kb = a.k + (int64_t)b * a.sk + h * dh;
  const float* vb = a.v + (int64_t)b * a.sv + h * dh;
  const float* ob = a.dout + (int64_t)b * a.so + h * dh;
  for (int i = threadIdx.x; i < Lq * dh; i += blockDim.x) {
    const int r = i / dh, c = i % dh;
    Qs[r * dhp + c] = qb[(int64_t)r * a.ldq + c];
    Os[r * dhp + c] = ob[(int64_t)r * a.ldo + c];
  }
  for (int i = threadIdx.x; i < Lk * dh; i += blockDim.x) {
    const int r = i / dh, c = i % dh;
    Ks[r * dhp + c] = kb[(int64_t)r * a.ldk + c];
    Vs[r * dhp + c] = vb[(int64_t)r * a.ldv + c];
  }
  const int64_t pbase = ((int64_t)b * a.heads + h) * Lq * Lk;
  for (int i = threadIdx.x; i < Lq * Lk; i += blockDim.x) {
    const float p = a.probs[pbase + i];
    Ps[i] = p;
    Pd[i] = (a.drop_mask != nullptr) ? (a.drop_mask[pbase + i] ? p * a.drop_scale : 0.f) : p;
  }
  __syncthreads();
  // dV[j,d] = sum_i Pd[i,j] dO[i,d]
  float* dvb = a.dv + (int64_t)b * a.sdv + h * dh;
  for (int idx = threadIdx.x; idx < Lk * dh; idx += blockDim.x) {
    const int j = idx / dh, d = idx % dh;
    float acc = 0.f;
    for (int i = 0; i < Lq; ++i) acc = fmaf(Pd[i * Lk + j], Os[i * dhp + d], acc);
    dvb[(int64_t)j * a.lddv + d] = acc;
  }
  __syncthreads();
  // dP[i,j] = (dO[i,:] . V[j,:]) * dropmask ; stored into Pd
  for (int idx = threadIdx.x; idx < Lq * Lk; idx += blockDim.x) {
    const int i = idx / Lk, j = idx % Lk;
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < dh; ++d) acc = fmaf(Os[i * dhp + d], Vs[j * dhp + d], acc);
    if (a.drop_mask != nullptr) acc *= a.drop_mask[pbase + idx] ? a.drop_scale : 0.f;
    Pd[idx] = acc;
  }
  __syncthreads();
  // dS = P * (dP - sum_j P dP), scaled by 1/sqrt(dh); stored into Ps
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float scale = rsqrtf((float)dh);
  for (int i = wid; i < Lq; i += nw) {
    float dot = 0.f;
    for (int j = lane; j < Lk; j += 32) dot = fmaf(Ps[i * Lk + j], Pd[i * Lk + j], dot);
    dot = warp_sum(dot);
    for (int j = lane; j < Lk; j += 32) Ps[i * Lk + j] = Ps[i * Lk + j] * (Pd[i * Lk + j] - dot) * scale;
  }
  __syncthreads();
  float* dqb = a.dq + (int64_t)b * a.sdq + h * dh;
  for (int idx = threadIdx.x; idx < Lq * dh; idx += blockDim.x) {
    const int i = idx / dh, d = idx % dh;
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) acc = fmaf(Ps[i * Lk + j], Ks[j * dhp + d], acc);
    dqb[(int64_t)i * a.lddq + d] = acc;
  }
  float* dkb = a.dk + (int64_t)b * a.sdk + h * dh;
  for (int idx = threadIdx.x; idx < Lk * dh; idx += blockDim.x) {
    const int j = idx / dh, d = idx % dh;
    float acc = 0.f;
    for (int i = 0; i < Lq; ++i) acc = fmaf(Ps[i * Lk + j], Qs[i * dhp + d], acc);
    dkb[(int64_t)j * a.lddk + d] = acc;
  }
}

__global__ void __launch_bounds__(256) reverse_tokens_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                             const int32_t* __restrict__ lengths, int B, int L, int Hd) {
  const int n4 = Hd >> 2;
  const int64_t total = (int64_t)B * L * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % n4);
    const int64_t bl = i / n4;
    const int l = (int)(bl % L), b = (int)(bl / L);
    const int src = lengths[b] - 1 - l;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0) v = reinterpret_cast<const float4*>(x + ((int64_t)b * L + src) * Hd)[j];
    reinterpret_cast<float4*>(out + ((int64_t)b * L + l) * Hd)[j] = v;
  }
}

// packed variant: x holds only the valid tokens, sample b's rows start at off[b]
__global__ void __launch_bounds__(256) reverse_tokens_packed_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                                    const int32_t* __restrict__ off, const int32_t* __restrict__ lengths,
                                                                    int B, int L, int Hd) {
  const int n4 = Hd >> 2;
  const int64_t total = (int64_t)B * L * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % n4);
    const int64_t bl = i / n4;
    const int l = (int)(bl % L), b = (int)(bl / L);
    const int src = lengths[b] - 1 - l;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (src >= 0) v = reinterpret_cast<const float4*>(x + ((int64_t)off[b] + src) * Hd)[j];
    reinterpret_cast<float4*>(out + ((int64_t)b * L + l) * Hd)[j] = v;
  }
}

// dst[r, :] = src[idx[r], :]  (row gather: packs the valid tokens of a padded [B*L, C] matrix)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ idx,
                                                          float* __restrict__ dst, int64_t ld_dst, int R, int C) {
  const int n4 = C >> 2;
  const int64_t total = (int64_t)R * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % n4);
    const int r = (int)(i / n4);
    reinterpret_cast<float4*>(dst + (int64_t)r * ld_dst)[j] = reinterpret_cast<const float4*>(src + (int64_t)idx[r] * ld_src)[j];
  }
}

inline bool ln_shape_ok(int Hd) { return Hd % 4 == 0 && Hd >= 4 && Hd <= 32 * 4 * LN_MAXV; }

}  // namespace

extern "C" int dasa_embed_layernorm(const int64_t* ids, int64_t ld_ids, int B, int L, int Hd, const float* word, const float* pos,
                                    const float* type0, const float* gamma, const float* beta, float eps,
                                    const uint8_t* drop_mask, float drop_scale, float* out, void* stream) {
  if (B <= 0 || L <= 0) return DASA_OK;
  if (!ln_shape_ok(Hd)) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(word) || !dasa_aligned16(pos) || !dasa_aligned16(type0) || !dasa_aligned16(gamma) || !dasa_aligned16(beta) ||
      !dasa_aligned16(out) || (drop_mask && reinterpret_cast<uintptr_t>(drop_mask) % 4))
    return DASA_ERR_BAD_ALIGN;
  embed_layernorm_kernel<<<(unsigned)dasa_cdiv((int64_t)B * L, 8), 256, 0, (cudaStream_t)stream>>>(
      ids, ld_ids, B, L, Hd, word, pos, type0, gamma, beta, eps, drop_mask, drop_scale, out);
  return dasa_check_launch("embed_layernorm_kernel");
}

extern "C" int dasa_dropout_residual_layernorm(const float* x, int64_t ldx, const uint8_t* drop_mask, float drop_scale,
                                               const float* resid, int64_t ldr, const float* gamma, const float* beta, float eps,
                                               const uint8_t* post_mask, float post_scale, float* out, int64_t ldo,
                                               float* stats_out, float* z_out, dasa_half_t* out_half, int R, int Hd, void* stream) {
  if (R <= 0) return DASA_OK;
  if (!ln_shape_ok(Hd)) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(x) || ldx % 4 || (resid && (!dasa_aligned16(resid) || ldr % 4)) || !dasa_aligned16(out) || ldo % 4 ||
      !dasa_aligned16(gamma) || !dasa_aligned16(beta) || (drop_mask && reinterpret_cast<uintptr_t>(drop_mask) % 4) ||
      (post_mask && reinterpret_cast<uintptr_t>(post_mask) % 4) || (z_out && !dasa_aligned16(z_out)))
    return DASA_ERR_BAD_ALIGN;
  dropout_residual_layernorm_kernel<false><<<(unsigned)dasa_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      x, ldx, drop_mask, LnStream{nullptr, 0ull, 0ull, 0u, 0}, drop_scale, resid, ldr, gamma, beta, eps, post_mask, post_scale, out,
      ldo, stats_out, z_out, reinterpret_cast<__half*>(out_half), R, Hd);
  return dasa_check_launch("dropout_residual_layernorm_kernel");
}

extern "C" int dasa_dropout_residual_layernorm_fwd(const void* x, int x_half, int64_t ldx, const uint8_t* drop_mask,
                                                   const uint64_t* drop_seed_dev, uint64_t drop_seed, uint64_t drop_base,
                                                   float drop_p, float drop_scale, const float* resid, int64_t ldr,
                                                   const float* gamma, const float* beta, float eps, float* out, int64_t ldo,
                                                   dasa_half_t* out_half, int R, int Hd, void* stream) {
  if (R <= 0) return DASA_OK;
  if (!ln_shape_ok(Hd)) return DASA_ERR_BAD_SHAPE;
  if (!(x_half ? (reinterpret_cast<uintptr_t>(x) % 8 == 0 && ldx % 4 == 0) : (dasa_aligned16(x) && ldx % 4 == 0)) ||
      (resid && (!dasa_aligned16(resid) || ldr % 4)) || !dasa_aligned16(out) || ldo % 4 || !dasa_aligned16(gamma) ||
      !dasa_aligned16(beta) || (drop_mask && reinterpret_cast<uintptr_t>(drop_mask) % 4) ||
      (out_half && reinterpret_cast<uintptr_t>(out_half) % 8))
    return DASA_ERR_BAD_ALIGN;
  const bool on = drop_mask == nullptr && drop_p > 0.f;
  const LnStream st{reinterpret_cast<const unsigned long long*>(drop_seed_dev), (unsigned long long)drop_seed,
                    (unsigned long long)drop_base, (uint32_t)(drop_p * 65536.0f), on ? 1 : 0};
  const float scale = (drop_mask != nullptr || on) ? drop_scale : 1.f;
  const unsigned grid = (unsigned)dasa_cdiv(R, 4);
  if (Hd == 768) {
    auto kern = x_half ? ln_fwd_fast_kernel<true, 6> : ln_fwd_fast_kernel<false, 6>;
    kern<<<grid, 128, 0, (cudaStream_t)stream>>>(x, ldx, drop_mask, st, scale, resid, ldr, gamma, beta, eps, out, ldo,
                                                 reinterpret_cast<__half*>(out_half), R);
    return dasa_check_launch("ln_fwd_fast_kernel");
  }
  if (x_half)
    dropout_residual_layernorm_kernel<true><<<grid, 128, 0, (cudaStream_t)stream>>>(
        x, ldx, drop_mask, st, scale, resid, ldr, gamma, beta, eps, nullptr, 1.f, out, ldo, nullptr, nullptr,
        reinterpret_cast<__half*>(out_half), R, Hd);
  else
    dropout_residual_layernorm_kernel<false><<<grid, 128, 0, (cudaStream_t)stream>>>(
        x, ldx, drop_mask, st, scale, resid, ldr, gamma, beta, eps, nullptr, 1.f, out, ldo, nullptr, nullptr,
        reinterpret_cast<__half*>(out_half), R, Hd);
  return dasa_check_launch("dropout_residual_layernorm_kernel(fwd)");
}

extern "C" int dasa_layernorm_bwd(const float* dout, int64_t lddo, const float* z, const float* gamma, const float* stats,
                                  const uint8_t* drop_mask, float drop_scale, const uint8_t* post_mask, float post_scale,
                                  float* dx, int64_t lddx, float* dresid, int64_t lddr, float* dgamma, float* dbeta, int R, int Hd,
                                  void* stream) {
  if (R <= 0) return DASA_OK;
  if (!ln_shape_ok(Hd)) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(dout) || lddo % 4 || !dasa_aligned16(z) || !dasa_aligned16(gamma) || (dx && (!dasa_aligned16(dx) || lddx % 4)) ||
      (dresid && (!dasa_aligned16(dresid) || lddr % 4)))
    return DASA_ERR_BAD_ALIGN;
  const unsigned grid = (unsigned)min((int64_t)DASA_NUM_SMS * 2, dasa_cdiv(R, 8));
  layernorm_bwd_kernel<<<grid, 256, 2 * Hd * sizeof(float), (cudaStream_t)stream>>>(
      dout, lddo, z, gamma, stats, drop_mask, drop_scale, post_mask, post_scale, dx, lddx, dresid, lddr, dgamma, dbeta, R, Hd);
  return dasa_check_launch("layernorm_bwd_kernel");
}

static int mha_fwd_impl(const int32_t* q_off, const int32_t* q_len, const int32_t* k_off, const int32_t* k_len,
                        const float* q, int64_t ldq, int64_t sq, const float* k, int64_t ldk, int64_t sk, const float* v,
                            int64_t ldv, int64_t sv, const uint8_t* key_pad, int64_t ld_pad, const uint8_t* drop_mask,
                            float drop_scale, float* out, int64_t ldo, int64_t so, float* probs_out, int B, int heads, int Lq,
                            int Lk, int dh, int precision, int out_half, void* stream) {
  if (B <= 0 || heads <= 0) return DASA_OK;
  if (out_half && probs_out != nullptr) return DASA_ERR_UNSUPPORTED;      // the probability-saving path re-reads `out` as floats
  if (precision == DASA_PREC_TF32 && Lq > 0 && Lk > 0 && Lk <= 8 * MHA_TC_MAXNT && dh % 8 == 0 && dh <= 64 && dasa_aligned16(q) &&
      dasa_aligned16(k) && dasa_aligned16(v) && !(ldq % 4 || ldk % 4 || ldv % 4 || sq % 4 || sk % 4 || sv % 4) &&
      dasa_aligned16(out) && !(ldo % 2 || so % 2)) {
    const int LqP = (Lq + 15) & ~15, LkP = (Lk + 7) & ~7;
    if (dh == 64) {                                            // K / V only in shared memory, Q and P in registers
      const size_t smem64 = sizeof(float) * (size_t)LkP * (MHA2_KS + MHA2_VS);
      // key-tile count as a template parameter: 6 (<= 48 keys: the 36 views), 8 (<= 64 keys: R2R instructions, no register
      // spills), 12 (<= 96 keys)
      void (*kern)(MhaArgs) = mha_fwd_tc64_kernel<MHA_TC_MAXNT>;
      switch (LkP >> 3) {                                      // exact tile counts for the lengths the rollout produces
        case 1: case 2: case 3: case 4: kern = mha_fwd_tc64_kernel<4>; break;
        case 5: kern = mha_fwd_tc64_kernel<5>; break;          // the 36 views
        case 6: kern = mha_fwd_tc64_kernel<6>; break;
        case 7: kern = mha_fwd_tc64_kernel<7>; break;
        case 8: kern = mha_fwd_tc64_kernel<8>; break;
        case 9: case 10: kern = mha_fwd_tc64_kernel<10>; break;
        default: break;
      }
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64);
      if (e2 != cudaSuccess) { dasa_set_error("mha_fwd_tc64 attr", e2); return DASA_ERR_CUDA; }
      MhaArgs at{q, k, v, ldq, sq, ldk, sk, ldv, sv, key_pad, ld_pad, drop_mask, drop_scale, out, ldo, so, probs_out, B, heads, Lq, Lk, dh,
                 out_half, q_off, q_len, k_off, k_len};
      const int warps = (LqP / 16) < 4 ? (LqP / 16) : 4;
      kern<<<(unsigned)(B * heads), 32 * warps, smem64, (cudaStream_t)stream>>>(at);
      return dasa_check_launch("mha_fwd_tc64_kernel");
    }
    const size_t smem_tc = sizeof(float) * ((size_t)LqP * (dh + 4) + (size_t)LkP * (dh + 4) + (size_t)LkP * (dh + 8) +
                                            (size_t)LqP * mha_tc_pstride(LkP));
    if (smem_tc <= 227 * 1024) {
      auto kern = (LkP <= 48) ? mha_fwd_tc_kernel<6> : mha_fwd_tc_kernel<MHA_TC_MAXNT>;
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc);
      if (e2 != cudaSuccess) { dasa_set_error("mha_fwd_tc attr", e2); return DASA_ERR_CUDA; }
      MhaArgs at{q, k, v, ldq, sq, ldk, sk, ldv, sv, key_pad, ld_pad, drop_mask, drop_scale, out, ldo, so, probs_out, B, heads, Lq, Lk, dh,
                 out_half, q_off, q_len, k_off, k_len};
      const int warps = (LqP / 16) < 4 ? (LqP / 16) : 4;       // one warp per 16-row query tile, no idle warps
      kern<<<(unsigned)(B * heads), 32 * warps, smem_tc, (cudaStream_t)stream>>>(at);
      return dasa_check_launch("mha_fwd_tc_kernel");
    }
  }
  if (Lq <= 0 || Lk <= 0 || dh <= 0) return DASA_ERR_BAD_SHAPE;
  if (Lk > 96 || dh > 64 || dh % 4 != 0) return DASA_ERR_UNSUPPORTED;
  if (!dasa_aligned16(q) || !dasa_aligned16(k) || !dasa_aligned16(v) || ldq % 4 || ldk % 4 || ldv % 4 || sq % 4 || sk % 4 || sv % 4)
    return DASA_ERR_BAD_ALIGN;
  const int LqP = (Lq + 3) & ~3, LkP = (Lk + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)LqP * (dh + 4) + (size_t)LkP * (dh + 4) + (size_t)LkP * dh + (size_t)LqP * LkP);
  if (smem > 227 * 1024) return DASA_ERR_BAD_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(mha_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error("mha_fwd attr", e); return DASA_ERR_CUDA; }
  MhaArgs a{q, k, v, ldq, sq, ldk, sk, ldv, sv, key_pad, ld_pad, drop_mask, drop_scale, out, ldo, so, probs_out, B, heads, Lq, Lk, dh,
                 out_half, q_off, q_len, k_off, k_len};
  mha_fwd_kernel<<<(unsigned)(B * heads), 256, smem, (cudaStream_t)stream>>>(a);
  return dasa_check_launch("mha_fwd_kernel");
}

extern "C" int dasa_mha_fwd(const float* q, int64_t ldq, int64_t sq, const float* k, int64_t ldk, int64_t sk, const float* v,
                            int64_t ldv, int64_t sv, const uint8_t* key_pad, int64_t ld_pad, const uint8_t* drop_mask,
                            float drop_scale, float* out, int64_t ldo, int64_t so, float* probs_out, int B, int heads, int Lq,
                            int Lk, int dh, int precision, int out_half, void* stream) {
  return mha_fwd_impl(nullptr, nullptr, nullptr, nullptr, q, ldq, sq, k, ldk, sk, v, ldv, sv, key_pad, ld_pad, drop_mask,
                      drop_scale, out, ldo, so, probs_out, B, heads, Lq, Lk, dh, precision, out_half, stream);
}

extern "C" int dasa_mha_fwd_varlen(const float* q, int64_t ldq, const int32_t* q_off, const int32_t* q_len, const float* k,
                                   int64_t ldk, const float* v, int64_t ldv, const int32_t* k_off, const int32_t* k_len,
                                   int64_t dense_q_stride, int64_t dense_kv_stride, const uint8_t* drop_mask, float drop_scale,
                                   float* out, int64_t ldo, int B, int heads, int max_Lq, int max_Lk, int dh, int precision,
                                   int out_half, void* stream) {
  if ((q_off == nullptr) != (q_len == nullptr) || (k_off == nullptr) != (k_len == nullptr)) return DASA_ERR_BAD_SHAPE;
  return mha_fwd_impl(q_off, q_len, k_off, k_len, q, ldq, dense_q_stride, k, ldk, dense_kv_stride, v, ldv, dense_kv_stride,
                      nullptr, 0, drop_mask, drop_scale, out, ldo, dense_q_stride / (ldq ? ldq : 1) * ldo, nullptr, B, heads,
                      max_Lq, max_Lk, dh, precision, out_half, stream);
}

extern "C" int dasa_mha_bwd(const float* q, int64_t ldq, int64_t sq, const float* k, int64_t ldk, int64_t sk, const float* v,
                            int64_t ldv, int64_t sv, const float* probs, const uint8_t* drop_mask, float drop_scale,
                            const float* dout, int64_t ldo, int64_t so, float* dq, int64_t lddq, int64_t sdq, float* dk,
                            int64_t lddk, int64_t sdk, float* dv, int64_t lddv, int64_t sdv, int B, int heads, int Lq, int Lk,
                            int dh, void* stream) {
  if (B <= 0 || heads <= 0) return DASA_OK;
  if (Lq <= 0 || Lk <= 0 || dh <= 0 || probs == nullptr) return DASA_ERR_BAD_SHAPE;
  const size_t smem = sizeof(float) * (2 * (size_t)Lq * (dh + 1) + 2 * (size_t)Lk * (dh + 1) + 2 * (size_t)Lq * Lk);
  if (smem > 227 * 1024) return DASA_ERR_BAD_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(mha_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error("mha_bwd attr", e); return DASA_ERR_CUDA; }
  MhaBwdArgs a{q, k, v, ldq, sq, ldk, sk, ldv, sv, probs, drop_mask, drop_scale, dout, ldo, so,
               dq, dk, dv, lddq, sdq, lddk, sdk, lddv, sdv, B, heads, Lq, Lk, dh};
  mha_bwd_kernel<<<(unsigned)(B * heads), 256, smem, (cudaStream_t)stream>>>(a);
  return dasa_check_launch("mha_bwd_kernel");
}

extern "C" int dasa_reverse_tokens_packed(const float* x, const int32_t* offsets, const int32_t* lengths, float* out, int B, int L,
                                          int Hd, void* stream) {
  if (B <= 0 || L <= 0) return DASA_OK;
  if (Hd % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(x) || !dasa_aligned16(out)) return DASA_ERR_BAD_ALIGN;
  const int64_t total = (int64_t)B * L * (Hd / 4);
  const unsigned grid = (unsigned)min((int64_t)DASA_NUM_SMS * 8, dasa_cdiv(total, 256));
  reverse_tokens_packed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, offsets, lengths, B, L, Hd);
  return dasa_check_launch("reverse_tokens_packed_kernel");
}

extern "C" int dasa_gather_rows(const float* src, int64_t ld_src, const int32_t* idx, float* dst, int64_t ld_dst, int R, int C,
                                void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  if (C % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(src) || !dasa_aligned16(dst) || ld_src % 4 || ld_dst % 4) return DASA_ERR_BAD_ALIGN;
  const int64_t total = (int64_t)R * (C / 4);
  const unsigned grid = (unsigned)min((int64_t)DASA_NUM_SMS * 8, dasa_cdiv(total, 256));
  gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, idx, dst, ld_dst, R, C);
  return dasa_check_launch("gather_rows_kernel");
}

extern "C" int dasa_reverse_tokens(const float* x, float* out, const int32_t* lengths, int B, int L, int Hd, void* stream) {
  if (B <= 0 || L <= 0) return DASA_OK;
  if (Hd % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(x) || !dasa_aligned16(out)) return DASA_ERR_BAD_ALIGN;
  const int64_t total = (int64_t)B * L * (Hd / 4);
  const unsigned grid = (unsigned)min((int64_t)DASA_NUM_SMS * 8, dasa_cdiv(total, 256));
  reverse_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, lengths, B, L, Hd);
  return dasa_check_launch("reverse_tokens_kernel");
}

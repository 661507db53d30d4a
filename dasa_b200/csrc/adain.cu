// Depth-guided AdaIN family (agent_dg.py:1513-1661, model.py:1822-1840): HBM-bound streaming kernels.
//   gate_modulate / gate_backward : the DGAdaChannel sigmoid gate given the pre-activation (K1 epilogue form)
//   view_stats + channel_modulate : DGAdaStatChannel / DGAdaMeanChannel (stats over the 36 views per channel)
//   adain_rows                    : adaptive_instance_normalization (stats over channels per view), one pass
// All loads/stores are 128-bit and streaming (L1::no_allocate); grids are sized in multiples of the SM count.
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 mask4(const uint8_t* m, float scale) {
  const uchar4 u = *reinterpret_cast<const uchar4*>(m);
  return make_float4(u.x ? scale : 0.f, u.y ? scale : 0.f, u.z ? scale : 0.f, u.w ? scale : 0.f);
}

// out = sigmoid(g) * f (* mask)
template <bool VEC>
__global__ void __launch_bounds__(256) gate_modulate_kernel(const float* __restrict__ g, int64_t ldg,
                                                            const float* __restrict__ f, int64_t ldf,
                                                            float* __restrict__ out, int64_t ldo,
                                                            const uint8_t* __restrict__ mask, float scale, int R, int C) {
  if (VEC) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)R * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / c4), c = (int)(i % c4) << 2;
      const float4 gv = ldg_stream4(g + (int64_t)r * ldg + c);
      const float4 fv = ldg_stream4(f + (int64_t)r * ldf + c);
      float4 o = make_float4(sigmoidf_(gv.x) * fv.x, sigmoidf_(gv.y) * fv.y, sigmoidf_(gv.z) * fv.z, sigmoidf_(gv.w) * fv.w);
      if (mask != nullptr) {
        const float4 m = mask4(mask + (int64_t)r * C + c, scale);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
      }
      stg_stream4(out + (int64_t)r * ldo + c, o);
    }
  } else {
    const int64_t total = (int64_t)R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / C), c = (int)(i % C);
      float o = sigmoidf_(g[(int64_t)r * ldg + c]) * f[(int64_t)r * ldf + c];
      if (mask != nullptr) o *= mask[(int64_t)r * C + c] ? scale : 0.f;
      out[(int64_t)r * ldo + c] = o;
    }
  }
}

// dg = dout * mask * f * s(1-s)
__global__ void __launch_bounds__(256) gate_backward_kernel(const float* __restrict__ dout, int64_t lddo,
                                                            const float* __restrict__ f, int64_t ldf,
                                                            const float* __restrict__ s, int64_t lds,
                                                            const uint8_t* __restrict__ mask, float scale,
                                                            float* __restrict__ dg, int64_t lddg, int R, int C, int vec,
                                                            __half* __restrict__ dg16, float scale16) {
  if (vec) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)R * c4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / c4), c = (int)(i % c4) << 2;
      const float4 d = ldg_stream4(dout + (int64_t)r * lddo + c);
      const float4 fv = ldg_stream4(f + (int64_t)r * ldf + c);
      const float4 sv = ldg_stream4(s + (int64_t)r * lds + c);
      float4 o = make_float4(d.x * fv.x * sv.x * (1.f - sv.x), d.y * fv.y * sv.y * (1.f - sv.y),
                             d.z * fv.z * sv.z * (1.f - sv.z), d.w * fv.w * sv.w * (1.f - sv.w));
      if (mask != nullptr) {
        const float4 m = mask4(mask + (int64_t)r * C + c, scale);
        o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
      }
      if (dg != nullptr) stg_stream4(dg + (int64_t)r * lddg + c, o);
      if (dg16 != nullptr) {                  // scaled, saturating fp16 copy: the dY operand of dW = dg^T d on kind::f16
        uint2 h;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.x) : "f"(o.y * scale16), "f"(o.x * scale16));
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h.y) : "f"(o.w * scale16), "f"(o.z * scale16));
        *reinterpret_cast<uint2*>(dg16 + (int64_t)r * C + c) = h;
      }
    }
  } else {
    const int64_t total = (int64_t)R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / C), c = (int)(i % C);
      const float sv = s[(int64_t)r * lds + c];
      float o = dout[(int64_t)r * lddo + c] * f[(int64_t)r * ldf + c] * sv * (1.f - sv);
      if (mask != nullptr) o *= mask[(int64_t)r * C + c] ? scale : 0.f;
      dg[(int64_t)r * lddg + c] = o;
    }
  }
}

// mean / unbiased std / max / min over V views for each (sample, channel); a thread owns 4 adjacent channels.
// Sums are taken about the first view's value (shifted data) so fp32 cancellation stays harmless.
__global__ void __launch_bounds__(128) view_stats_kernel(const float* __restrict__ d, int64_t ld_row, int64_t ld_sample,
                                                         int N, int V, int C, float* __restrict__ stats) {
  const int c4 = C >> 2;
  const int64_t total = (int64_t)N * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / c4), c = (int)(i % c4) << 2;
    const float* p = d + (int64_t)n * ld_sample + c;
    const float4 x0 = ldg_stream4(p);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s, mx = x0, mn = x0;
#pragma unroll 6
    for (int v = 1; v < V; ++v) {
      const float4 x = ldg_stream4(p + (int64_t)v * ld_row);
      const float dx = x.x - x0.x, dy = x.y - x0.y, dz = x.z - x0.z, dw = x.w - x0.w;
      s.x += dx; s.y += dy; s.z += dz; s.w += dw;
      q.x = fmaf(dx, dx, q.x); q.y = fmaf(dy, dy, q.y); q.z = fmaf(dz, dz, q.z); q.w = fmaf(dw, dw, q.w);
      mx.x = fmaxf(mx.x, x.x); mx.y = fmaxf(mx.y, x.y); mx.z = fmaxf(mx.z, x.z); mx.w = fmaxf(mx.w, x.w);
      mn.x = fminf(mn.x, x.x); mn.y = fminf(mn.y, x.y); mn.z = fminf(mn.z, x.z); mn.w = fminf(mn.w, x.w);
    }
    const float inv = 1.f / V, invm1 = 1.f / (V - 1);
    float4 mean = make_float4(x0.x + s.x * inv, x0.y + s.y * inv, x0.z + s.z * inv, x0.w + s.w * inv);
    float4 sd = make_float4(sqrtf(fmaxf(q.x - s.x * s.x * inv, 0.f) * invm1), sqrtf(fmaxf(q.y - s.y * s.y * inv, 0.f) * invm1),
                            sqrtf(fmaxf(q.z - s.z * s.z * inv, 0.f) * invm1), sqrtf(fmaxf(q.w - s.w * s.w * inv, 0.f) * invm1));
    float* o = stats + (int64_t)n * 4 * C + c;
    *reinterpret_cast<float4*>(o) = mean;
    *reinterpret_cast<float4*>(o + C) = sd;
    *reinterpret_cast<float4*>(o + 2 * C) = mx;
    *reinterpret_cast<float4*>(o + 3 * C) = mn;
  }
}

__global__ void __launch_bounds__(256) channel_modulate_kernel(const float* __restrict__ f, int64_t ldf_row, int64_t ldf_sample,
                                                               const float* __restrict__ a, const float* __restrict__ b,
                                                               float* __restrict__ out, int64_t ldo_row, int64_t ldo_sample,
                                                               int N, int V, int C) {
  const int c4 = C >> 2;
  const int64_t total = (int64_t)N * V * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) << 2;
    const int64_t nv = i / c4;
    const int v = (int)(nv % V), n = (int)(nv / V);
    const float4 fv = ldg_stream4(f + (int64_t)n * ldf_sample + (int64_t)v * ldf_row + c);
    const float4 av = __ldg(reinterpret_cast<const float4*>(a + (int64_t)n * C + c));
    float4 o = make_float4(av.x * fv.x, av.y * fv.y, av.z * fv.z, av.w * fv.w);
    if (b != nullptr) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(b + (int64_t)n * C + c));
      o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
    }
    stg_stream4(out + (int64_t)n * ldo_sample + (int64_t)v * ldo_row + c, o);
  }
}

// backward of out = a[n,c] * f[n,v,c] + b[n,c]: da = sum_v dout*f, db = sum_v dout, df = a*dout (optional). One pass over
// dout and f: a thread owns 4 channels of one sample and walks the V view rows (coalesced 128-bit rows across the warp).
__global__ void __launch_bounds__(128) channel_modulate_bwd_kernel(const float* __restrict__ dout, int64_t ldo_row,
                                                                   int64_t ldo_sample, const float* __restrict__ f,
                                                                   int64_t ldf_row, int64_t ldf_sample,
                                                                   const float* __restrict__ a, float* __restrict__ da,
                                                                   float* __restrict__ db, float* __restrict__ df,
                                                                   int64_t lddf_row, int64_t lddf_sample, int N, int V, int C) {
  const int c4 = C >> 2;
  const int64_t total = (int64_t)N * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / c4), c = (int)(i % c4) << 2;
    const float* po = dout + (int64_t)n * ldo_sample + c;
    const float* pf = f + (int64_t)n * ldf_sample + c;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (df != nullptr) av = __ldg(reinterpret_cast<const float4*>(a + (int64_t)n * C + c));
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
#pragma unroll 6
    for (int v = 0; v < V; ++v) {
      const float4 g = ldg_stream4(po + (int64_t)v * ldo_row);
      const float4 x = ldg_stream4(pf + (int64_t)v * ldf_row);
      sa.x = fmaf(g.x, x.x, sa.x); sa.y = fmaf(g.y, x.y, sa.y); sa.z = fmaf(g.z, x.z, sa.z); sa.w = fmaf(g.w, x.w, sa.w);
      sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
      if (df != nullptr)
        stg_stream4(df + (int64_t)n * lddf_sample + (int64_t)v * lddf_row + c,
                    make_float4(av.x * g.x, av.y * g.y, av.z * g.z, av.w * g.w));
    }
    *reinterpret_cast<float4*>(da + (int64_t)n * C + c) = sa;
    if (db != nullptr) *reinterpret_cast<float4*>(db + (int64_t)n * C + c) = sb;
  }
}

// two block-wide sums at once; `red` is >= 64 floats of shared memory
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  a = warp_sum(a); b = warp_sum(b);
  __syncthreads();
  if (lane == 0) { red[wid] = a; red[32 + wid] = b; }
  __syncthreads();
  a = warp_sum((lane < nw) ? red[lane] : 0.f);
  b = warp_sum((lane < nw) ? red[32 + lane] : 0.f);
}

// adaptive_instance_normalization: one CTA (NT threads) per (sample, view) row; the content row lives in registers,
// the style row only feeds the two reductions. Two-pass variance (mean first) like torch.var. NT grows with C so that a thread
// never holds more than 4 float4 of each row (at VPT = 8 the 64 row registers cut the occupancy: 0.65 of HBM at C >= 3072).
template <int VPT, int NT>
__global__ void __launch_bounds__(NT) adain_rows_kernel(const float* __restrict__ f, int64_t ldf, const float* __restrict__ d,
                                                         int64_t ldd, float* __restrict__ out, int64_t ldo, int R, int C,
                                                         float eps) {
  __shared__ float red[64];
  const int c4 = C >> 2;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    const float* fr = f + (int64_t)r * ldf;
    const float* dr = d + (int64_t)r * ldd;
    float4 fv[VPT], dv[VPT];
    float sf = 0.f, sd = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int j = threadIdx.x + i * NT;
      if (j < c4) {
        fv[i] = ldg_stream4(fr + 4 * j);
        dv[i] = ldg_stream4(dr + 4 * j);
        sf += (fv[i].x + fv[i].y) + (fv[i].z + fv[i].w);
        sd += (dv[i].x + dv[i].y) + (dv[i].z + dv[i].w);
      }
    }
    float mu_f = sf, mu_d = sd;
    block_sum2(mu_f, mu_d, red);                 // both reductions behind one pair of barriers (same summation order as block_sum)
    mu_f /= C; mu_d /= C;
    float qf = 0.f, qd = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int j = threadIdx.x + i * NT;
      if (j < c4) {
        float a;
        a = fv[i].x - mu_f; qf = fmaf(a, a, qf); a = fv[i].y - mu_f; qf = fmaf(a, a, qf);
        a = fv[i].z - mu_f; qf = fmaf(a, a, qf); a = fv[i].w - mu_f; qf = fmaf(a, a, qf);
        a = dv[i].x - mu_d; qd = fmaf(a, a, qd); a = dv[i].y - mu_d; qd = fmaf(a, a, qd);
        a = dv[i].z - mu_d; qd = fmaf(a, a, qd); a = dv[i].w - mu_d; qd = fmaf(a, a, qd);
      }
    }
    block_sum2(qf, qd, red);
    const float sd_f = sqrtf(qf / (C - 1) + eps);
    const float sd_d = sqrtf(qd / (C - 1) + eps);
    float* orow = out + (int64_t)r * ldo;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int j = threadIdx.x + i * NT;
      if (j < c4) {
        float4 o;
        // same operation order as the reference: ((f - mu_f) / sd_f) * sd_d + mu_d
        o.x = (fv[i].x - mu_f) / sd_f * sd_d + mu_d;
        o.y = (fv[i].y - mu_f) / sd_f * sd_d + mu_d;
        o.z = (fv[i].z - mu_f) / sd_f * sd_d + mu_d;
        o.w = (fv[i].w - mu_f) / sd_f * sd_d + mu_d;
        stg_stream4(orow + 4 * j, o);
      }
    }
  }
}

inline unsigned stream_grid(int64_t work_items, int threads, int per_sm) {
  const int64_t want = dasa_cdiv(work_items, threads);
  const int64_t cap = (int64_t)DASA_NUM_SMS * per_sm;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

inline bool vec_ok(const void* p, int64_t ld) { return dasa_aligned16(p) && (ld % 4 == 0); }

}  // namespace

extern "C" int dasa_gate_modulate(const float* g, int64_t ldg, const float* f, int64_t ldf, float* out, int64_t ldo,
                                  const uint8_t* drop_mask, float drop_scale, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (C % 4 == 0) && vec_ok(g, ldg) && vec_ok(f, ldf) && vec_ok(out, ldo) &&
                   (drop_mask == nullptr || (reinterpret_cast<uintptr_t>(drop_mask) % 4 == 0));
  if (vec)
    gate_modulate_kernel<true><<<stream_grid((int64_t)R * C / 4, 256, 8), 256, 0, st>>>(g, ldg, f, ldf, out, ldo, drop_mask,
                                                                                        drop_scale, R, C);
  else
    gate_modulate_kernel<false><<<stream_grid((int64_t)R * C, 256, 8), 256, 0, st>>>(g, ldg, f, ldf, out, ldo, drop_mask,
                                                                                     drop_scale, R, C);
  return dasa_check_launch("gate_modulate_kernel");
}

extern "C" int dasa_gate_backward(const float* dout, int64_t lddo, const float* f, int64_t ldf, const float* s, int64_t lds,
                                  const uint8_t* drop_mask, float drop_scale, float* dg, int64_t lddg, int R, int C,
                                  void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  const bool vec = (C % 4 == 0) && vec_ok(dout, lddo) && vec_ok(f, ldf) && vec_ok(s, lds) && vec_ok(dg, lddg) &&
                   (drop_mask == nullptr || (reinterpret_cast<uintptr_t>(drop_mask) % 4 == 0));
  gate_backward_kernel<<<stream_grid((int64_t)R * C / (vec ? 4 : 1), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      dout, lddo, f, ldf, s, lds, drop_mask, drop_scale, dg, lddg, R, C, vec ? 1 : 0, nullptr, 1.f);
  return dasa_check_launch("gate_backward_kernel");
}

extern "C" int dasa_gate_backward_h(const float* dout, int64_t lddo, const float* f, int64_t ldf, const float* s, int64_t lds,
                                    const uint8_t* drop_mask, float drop_scale, float* dg, int64_t lddg, dasa_half_t* dg16,
                                    float scale16, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  const bool vec = (C % 4 == 0) && vec_ok(dout, lddo) && vec_ok(f, ldf) && vec_ok(s, lds) && (dg == nullptr || vec_ok(dg, lddg)) &&
                   (drop_mask == nullptr || (reinterpret_cast<uintptr_t>(drop_mask) % 4 == 0));
  if (dg16 != nullptr && (!vec || (reinterpret_cast<uintptr_t>(dg16) & 7))) return DASA_ERR_BAD_ALIGN;
  if (dg == nullptr && (dg16 == nullptr || !vec)) return DASA_ERR_BAD_SHAPE;
  gate_backward_kernel<<<stream_grid((int64_t)R * C / (vec ? 4 : 1), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      dout, lddo, f, ldf, s, lds, drop_mask, drop_scale, dg, lddg, R, C, vec ? 1 : 0, reinterpret_cast<__half*>(dg16), scale16);
  return dasa_check_launch("gate_backward_kernel");
}

extern "C" int dasa_view_stats(const float* d, int64_t ld_row, int64_t ld_sample, int N, int V, int C, float* stats,
                               void* stream) {
  if (N <= 0) return DASA_OK;
  if (V < 2 || C % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!vec_ok(d, ld_row) || ld_sample % 4 != 0 || !dasa_aligned16(stats)) return DASA_ERR_BAD_ALIGN;
  view_stats_kernel<<<stream_grid((int64_t)N * C / 4, 128, 16), 128, 0, (cudaStream_t)stream>>>(d, ld_row, ld_sample, N, V, C,
                                                                                               stats);
  return dasa_check_launch("view_stats_kernel");
}

extern "C" int dasa_channel_modulate(const float* f, int64_t ldf_row, int64_t ldf_sample, const float* a, const float* b,
                                     float* out, int64_t ldo_row, int64_t ldo_sample, int N, int V, int C, void* stream) {
  if (N <= 0) return DASA_OK;
  if (C % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!vec_ok(f, ldf_row) || ldf_sample % 4 != 0 || !vec_ok(out, ldo_row) || ldo_sample % 4 != 0 || !dasa_aligned16(a) ||
      (b != nullptr && !dasa_aligned16(b)))
    return DASA_ERR_BAD_ALIGN;
  channel_modulate_kernel<<<stream_grid((int64_t)N * V * C / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(
      f, ldf_row, ldf_sample, a, b, out, ldo_row, ldo_sample, N, V, C);
  return dasa_check_launch("channel_modulate_kernel");
}

extern "C" int dasa_channel_modulate_bwd(const float* dout, int64_t ldo_row, int64_t ldo_sample, const float* f, int64_t ldf_row,
                                         int64_t ldf_sample, const float* a, float* da, float* db, float* df, int64_t lddf_row,
                                         int64_t lddf_sample, int N, int V, int C, void* stream) {
  if (N <= 0) return DASA_OK;
  if (C % 4 != 0 || V <= 0) return DASA_ERR_BAD_SHAPE;
  if (!vec_ok(dout, ldo_row) || ldo_sample % 4 != 0 || !vec_ok(f, ldf_row) || ldf_sample % 4 != 0 || !dasa_aligned16(da) ||
      (db != nullptr && !dasa_aligned16(db)) ||
      (df != nullptr && (!vec_ok(df, lddf_row) || lddf_sample % 4 != 0 || a == nullptr || !dasa_aligned16(a))))
    return DASA_ERR_BAD_ALIGN;
  channel_modulate_bwd_kernel<<<stream_grid((int64_t)N * C / 4, 128, 16), 128, 0, (cudaStream_t)stream>>>(
      dout, ldo_row, ldo_sample, f, ldf_row, ldf_sample, a, da, db, df, lddf_row, lddf_sample, N, V, C);
  return dasa_check_launch("channel_modulate_bwd_kernel");
}

extern "C" int dasa_adain_rows(const float* f, int64_t ldf, const float* d, int64_t ldd, float* out, int64_t ldo, int R,
                               int C, float eps, void* stream) {
  if (R <= 0) return DASA_OK;
  if (C % 4 != 0 || C < 8 || C > 8192) return DASA_ERR_BAD_SHAPE;
  if (!vec_ok(f, ldf) || !vec_ok(d, ldd) || !vec_ok(out, ldo)) return DASA_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)(R < DASA_NUM_SMS * 16 ? R : DASA_NUM_SMS * 16);
  const int c4 = C / 4;
  if (c4 <= 512) adain_rows_kernel<4, 128><<<grid, 128, 0, st>>>(f, ldf, d, ldd, out, ldo, R, C, eps);
  else if (c4 <= 768) adain_rows_kernel<3, 256><<<grid / 2 + 1, 256, 0, st>>>(f, ldf, d, ldd, out, ldo, R, C, eps);
  else if (c4 <= 1024) adain_rows_kernel<4, 256><<<grid / 2 + 1, 256, 0, st>>>(f, ldf, d, ldd, out, ldo, R, C, eps);
  else adain_rows_kernel<4, 512><<<grid / 4 + 1, 512, 0, st>>>(f, ldf, d, ldd, out, ldo, R, C, eps);
  return dasa_check_launch("adain_rows_kernel");
}

// Fused recurrence of the encoder's packed bidirectional LSTM (r2rmodel.py:2339-2357) for small batches (B <= 20).
// One launch per time step handles BOTH directions: every CTA owns 16 hidden units of one direction, keeps the previous
// hidden state (fwd) / the gate gradients (bwd) of the whole batch in shared memory, streams its 64 recurrent-weight rows
// (256 KB, L2-resident across the 80 steps) with 128-bit loads, and applies the LSTM pointwise math for its units in the
// same kernel. Exact fp32 (FFMA). The whole sequence loop is issued from one C call, so the host cost is one call per
// encoder pass instead of ~320.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int UNITS = 16;       // hidden units per CTA
constexpr int THREADS = 256;    // 8 warps

struct SeqFwd {
  const float* xp[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  float* hs[2]; float* cs[2]; float* acts[2]; float* out; const int32_t* lengths;
  int B, L, H;
};

struct SeqBwd {
  const float* w_hh_t[2]; const float* acts[2]; const float* cs[2]; const float* dout;
  const float* dh_fin[2]; const float* dc_fin[2];
  float* dgates[2]; float* dh_pass[2]; float* dc_work[2];   // dh_pass / dc_work: [2 (ping-pong)][B][H]
  const int32_t* lengths;
  int B, L, H;
};

template <int BT>
__global__ void __launch_bounds__(THREADS) bilstm_step_fwd_kernel(SeqFwd p, int s) {
  extern __shared__ __align__(16) float smem[];
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int B = p.B, L = p.L, H = p.H;
  const int l = d == 0 ? s : L - 1 - s;
  float* hbuf = smem;                         // [BT][H]
  float* gbuf = smem + (size_t)BT * H;        // [64][BT]
  const float* h_prev = p.hs[d] + (size_t)s * B * H;
  for (int i = threadIdx.x; i < BT * (H >> 2); i += THREADS) {
    const int b = i / (H >> 2), k4 = i % (H >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < B) v = reinterpret_cast<const float4*>(h_prev + (size_t)b * H)[k4];
    reinterpret_cast<float4*>(hbuf + (size_t)b * H)[k4] = v;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = lane >> 4, kl = lane & 15;
  // this lane's 4 weight rows: unit jj = 2*warp + rg, gates 0..3  -> global row g*H + j
  const int jj = 2 * warp + rg;
  const float* wbase = p.w_hh[d] + (size_t)(j0 + jj) * H;
  float acc[4][BT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[r][b] = 0.f;
  const int iters = H >> 6;                   // float4 index = kl + 16*i
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
    const int k4 = kl + 16 * i;
    float4 w[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) w[r] = __ldg(reinterpret_cast<const float4*>(wbase + (size_t)r * H * H) + k4);
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const float4 h4 = reinterpret_cast<const float4*>(hbuf + (size_t)b * H)[k4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][b] = fmaf(w[r].x, h4.x, acc[r][b]);
        acc[r][b] = fmaf(w[r].y, h4.y, acc[r][b]);
        acc[r][b] = fmaf(w[r].z, h4.z, acc[r][b]);
        acc[r][b] = fmaf(w[r].w, h4.w, acc[r][b]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      float v = acc[r][b];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (kl == 0) gbuf[(jj * 4 + r) * BT + b] = v;
    }
  __syncthreads();

  // pointwise LSTM cell for (b, unit)
  for (int t = threadIdx.x; t < UNITS * B; t += THREADS) {
    const int u = t % UNITS, b = t / UNITS, j = j0 + u;
    const size_t sb = (size_t)b * H + j;
    const float cp = p.cs[d][(size_t)s * B * H + sb];
    float* hn = p.hs[d] + (size_t)(s + 1) * B * H;
    float* cn = p.cs[d] + (size_t)(s + 1) * B * H;
    float* a = p.acts[d] + ((size_t)s * B + b) * 4 * H;
    float* o = p.out + ((size_t)b * L + l) * 2 * H + (size_t)d * H + j;
    if (l >= p.lengths[b]) {                       // packed-sequence semantics: carry state, zero output row
      hn[sb] = hbuf[(size_t)b * H + j];
      cn[sb] = cp;
      *o = 0.f;
      a[j] = 0.f; a[H + j] = 0.f; a[2 * H + j] = 0.f; a[3 * H + j] = 0.f;
      continue;
    }
    const float* xrow = p.xp[d] + ((size_t)b * L + l) * 4 * H;
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      g[q] = gbuf[(u * 4 + q) * BT + b] + xrow[q * H + j] + __ldg(p.b_ih[d] + q * H + j) + __ldg(p.b_hh[d] + q * H + j);
    const float ig = sigmoidf_(g[0]), fg = sigmoidf_(g[1]), gg = tanhf(g[2]), og = sigmoidf_(g[3]);
    const float c1 = fg * cp + ig * gg;
    const float h1 = og * tanhf(c1);
    hn[sb] = h1;
    cn[sb] = c1;
    *o = h1;
    a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
  }
}

template <int BT>
__global__ void __launch_bounds__(THREADS) bilstm_step_bwd_kernel(SeqBwd p, int s) {
  extern __shared__ __align__(16) float smem[];
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int B = p.B, L = p.L, H = p.H, G = 4 * H;
  const int l = d == 0 ? s : L - 1 - s;
  const int CH = G < 1024 ? G : 1024;             // reduction chunk staged in shared memory
  float* dbuf = smem;                             // [BT][CH]
  float* red = smem + (size_t)BT * CH;            // [8 warps][UNITS][BT]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool last = (s == L - 1);

  if (!last) {
    // rec[b, unit] = dgates_{s+1}[b, :] . W_hh^T[unit, :]   (reduction over the 4H gate rows)
    const float* dg_next = p.dgates[d] + (size_t)(s + 1) * B * G;
    const int rg = lane >> 3, kl = lane & 7;      // 4 unit-groups x 8 k-lanes; the warp covers CH/8 of each chunk
    const int wspan4 = (CH >> 3) >> 2;            // float4 per warp per chunk
    float acc[4][BT];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[r][b] = 0.f;
    for (int c0 = 0; c0 < G; c0 += CH) {
      __syncthreads();
      for (int i = threadIdx.x; i < BT * (CH >> 2); i += THREADS) {
        const int b = i / (CH >> 2), k4 = i % (CH >> 2);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) v = reinterpret_cast<const float4*>(dg_next + (size_t)b * G + c0)[k4];
        reinterpret_cast<float4*>(dbuf + (size_t)b * CH)[k4] = v;
      }
      __syncthreads();
      const float* wbase = p.w_hh_t[d] + (size_t)(j0 + 4 * rg) * G + c0;
#pragma unroll 2
      for (int i = kl; i < wspan4; i += 8) {
        const int k4 = warp * wspan4 + i;
        float4 w[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) w[r] = __ldg(reinterpret_cast<const float4*>(wbase + (size_t)r * G) + k4);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 g4 = reinterpret_cast<const float4*>(dbuf + (size_t)b * CH)[k4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            acc[r][b] = fmaf(w[r].x, g4.x, acc[r][b]);
            acc[r][b] = fmaf(w[r].y, g4.y, acc[r][b]);
            acc[r][b] = fmaf(w[r].z, g4.z, acc[r][b]);
            acc[r][b] = fmaf(w[r].w, g4.w, acc[r][b]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        float v = acc[r][b];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (kl == 0) red[((size_t)warp * UNITS + 4 * rg + r) * BT + b] = v;
      }
  }
  __syncthreads();

  const int par = s & 1;                          // ping-pong: read [par^1] (written by step s+1), write [par]
  for (int t = threadIdx.x; t < UNITS * B; t += THREADS) {
    const int u = t % UNITS, b = t / UNITS, j = j0 + u;
    const size_t sb = (size_t)b * H + j;
    float dh, dc;
    if (last) {
      dh = p.dh_fin[d] ? p.dh_fin[d][sb] : 0.f;
      dc = p.dc_fin[d] ? p.dc_fin[d][sb] : 0.f;
    } else {
      float rec = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) rec += red[((size_t)w * UNITS + u) * BT + b];
      dh = rec + p.dh_pass[d][(size_t)(par ^ 1) * B * H + sb];
      dc = p.dc_work[d][(size_t)(par ^ 1) * B * H + sb];
    }
    float* dg = p.dgates[d] + ((size_t)s * B + b) * G;
    float* dh_pass = p.dh_pass[d] + (size_t)par * B * H;
    float* dc_out = p.dc_work[d] + (size_t)par * B * H;
    if (l >= p.lengths[b]) {                       // inactive: state was carried, pass gradients straight through
      dg[j] = 0.f; dg[H + j] = 0.f; dg[2 * H + j] = 0.f; dg[3 * H + j] = 0.f;
      dh_pass[sb] = dh;
      dc_out[sb] = dc;
      continue;
    }
    dh += p.dout[((size_t)b * L + l) * 2 * H + (size_t)d * H + j];
    const float* a = p.acts[d] + ((size_t)s * B + b) * G;
    const float ig = a[j], fg = a[H + j], gg = a[2 * H + j], og = a[3 * H + j];
    const float cp = p.cs[d][(size_t)s * B * H + sb];
    const float tc = tanhf(p.cs[d][(size_t)(s + 1) * B * H + sb]);
    const float dct = dc + dh * og * (1.f - tc * tc);
    dg[j] = dct * gg * ig * (1.f - ig);
    dg[H + j] = dct * cp * fg * (1.f - fg);
    dg[2 * H + j] = dct * ig * (1.f - gg * gg);
    dg[3 * H + j] = dh * tc * og * (1.f - og);
    dh_pass[sb] = 0.f;
    dc_out[sb] = dct * fg;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// TF32 tensor-core variants (precision mode tf32). Same CTA = 16 hidden units decomposition, but the 64 x H (fwd) /
// 16 x 4H (bwd) weight slice is streamed from L2 straight into mma.sync.m16n8k8 A-fragments with 128-bit loads (the k
// index inside each 16-wide group is permuted identically for A and B, so one float4 per row feeds two MMAs), the
// batch side (h / dgates, <= 24 rows) is loaded the same way, the 8 warps split the reduction dimension and combine
// their partial tiles through shared memory. No operand staging, no per-chunk barriers; fp32 accumulate.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int NB = 24;          // padded batch (3 n-tiles of 8)

__global__ void __launch_bounds__(THREADS) bilstm_step_fwd_tc_kernel(SeqFwd p, int s) {
  extern __shared__ __align__(16) float smem_dyn[];
  float (*part)[64][NB + 1] = reinterpret_cast<float (*)[64][NB + 1]>(smem_dyn);   // [8] per-warp partial gate tiles
  float (*gbuf)[NB + 1] = part[0];                 // the reduced tile overwrites warp 0's partials in place
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int B = p.B, L = p.L, H = p.H;
  const int l = d == 0 ? s : L - 1 - s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const float* h_prev = p.hs[d] + (size_t)s * B * H;
  const float* W = p.w_hh[d];
  const int kspan = H >> 3, kw0 = warp * kspan;    // this warp's slice of the reduction dimension
  // local row r = unit*4 + gate  ->  global W row gate*H + j0 + unit ; m-tile mt covers local rows 16mt .. 16mt+15
  const float* rowA[4];
  const float* rowB[4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    const int ra = 16 * mt + g, rb = ra + 8;
    rowA[mt] = W + ((size_t)(ra & 3) * H + j0 + (ra >> 2)) * H + kw0 + 4 * t;
    rowB[mt] = W + ((size_t)(rb & 3) * H + j0 + (rb >> 2)) * H + kw0 + 4 * t;
  }
  const float* hrow[3];
  bool hok[3];
#pragma unroll
  for (int nt = 0; nt < 3; ++nt) {
    const int b = 8 * nt + g;
    hok[nt] = b < B;
    hrow[nt] = h_prev + (size_t)(hok[nt] ? b : 0) * H + kw0 + 4 * t;
  }
  float acc[4][3][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  const int groups = kspan >> 4;
  constexpr int PF = 2;                              // 16-wide k groups of operand loads kept in flight per lane
  float4 wa[PF][4], wb[PF][4], hv[PF][3];
  auto load = [&](int gi, int slot) {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      wa[slot][mt] = __ldg(reinterpret_cast<const float4*>(rowA[mt] + 16 * gi));
      wb[slot][mt] = __ldg(reinterpret_cast<const float4*>(rowB[mt] + 16 * gi));
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
      hv[slot][nt] = hok[nt] ? *reinterpret_cast<const float4*>(hrow[nt] + 16 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
#pragma unroll
  for (int i = 0; i < PF; ++i)
    if (i < groups) load(i, i);
  for (int g0 = 0; g0 < groups; g0 += PF) {
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      const int gi = g0 + i;
      if (gi >= groups) break;
      float4 ca[4], cb[4], ch[3];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { ca[mt] = wa[i][mt]; cb[mt] = wb[i][mt]; }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) ch[nt] = hv[i][nt];
      if (gi + PF < groups) load(gi + PF, i);        // refill this slot while the MMAs below run
      uint32_t b0[3], b1[3], b2[3], b3[3];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) { b0[nt] = to_tf32(ch[nt].x); b1[nt] = to_tf32(ch[nt].y); b2[nt] = to_tf32(ch[nt].z); b3[nt] = to_tf32(ch[nt].w); }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint32_t a0 = to_tf32(ca[mt].x), a1 = to_tf32(cb[mt].x), a2 = to_tf32(ca[mt].y), a3 = to_tf32(cb[mt].y);
        const uint32_t e0 = to_tf32(ca[mt].z), e1 = to_tf32(cb[mt].z), e2 = to_tf32(ca[mt].w), e3 = to_tf32(cb[mt].w);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          mma_tf32(acc[mt][nt], a0, a1, a2, a3, b0[nt], b1[nt]);
          mma_tf32(acc[mt][nt], e0, e1, e2, e3, b2[nt], b3[nt]);
        }
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      part[warp][16 * mt + g][8 * nt + 2 * t] = acc[mt][nt][0];
      part[warp][16 * mt + g][8 * nt + 2 * t + 1] = acc[mt][nt][1];
      part[warp][16 * mt + g + 8][8 * nt + 2 * t] = acc[mt][nt][2];
      part[warp][16 * mt + g + 8][8 * nt + 2 * t + 1] = acc[mt][nt][3];
    }
  __syncthreads();
  for (int o = threadIdx.x; o < 64 * NB; o += THREADS) {
    const int r = o / NB, c = o % NB;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += part[w][r][c];
    gbuf[r][c] = v;
  }
  __syncthreads();
  for (int tt = threadIdx.x; tt < UNITS * B; tt += THREADS) {
    const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
    const size_t sb = (size_t)b * H + j;
    const float cp = p.cs[d][(size_t)s * B * H + sb];
    float* hn = p.hs[d] + (size_t)(s + 1) * B * H;
    float* cn = p.cs[d] + (size_t)(s + 1) * B * H;
    float* a = p.acts[d] + ((size_t)s * B + b) * 4 * H;
    float* o = p.out + ((size_t)b * L + l) * 2 * H + (size_t)d * H + j;
    if (l >= p.lengths[b]) {
      hn[sb] = h_prev[sb];
      cn[sb] = cp;
      *o = 0.f;
      a[j] = 0.f; a[H + j] = 0.f; a[2 * H + j] = 0.f; a[3 * H + j] = 0.f;
      continue;
    }
    const float* xrow = p.xp[d] + ((size_t)b * L + l) * 4 * H;
    float gt[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      gt[q] = gbuf[u * 4 + q][b] + xrow[q * H + j] + __ldg(p.b_ih[d] + q * H + j) + __ldg(p.b_hh[d] + q * H + j);
    const float ig = sigmoidf_(gt[0]), fg = sigmoidf_(gt[1]), gg = tanhf(gt[2]), og = sigmoidf_(gt[3]);
    const float c1 = fg * cp + ig * gg;
    const float h1 = og * tanhf(c1);
    hn[sb] = h1;
    cn[sb] = c1;
    *o = h1;
    a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
  }
}

__global__ void __launch_bounds__(THREADS) bilstm_step_bwd_tc_kernel(SeqBwd p, int s) {
  __shared__ float part[8][UNITS][NB + 1];
  __shared__ float red[UNITS][NB];
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int B = p.B, L = p.L, H = p.H, G = 4 * H;
  const int l = d == 0 ? s : L - 1 - s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bool last = (s == L - 1);
  if (!last) {
    // rec[b, unit] = dgates_{s+1}[b, :] . W_hh^T[unit, :] : one 16-row m-tile, 3 n-tiles, reduction 4H split over the warps
    const float* dg_next = p.dgates[d] + (size_t)(s + 1) * B * G;
    const int kspan = G >> 3, kw0 = warp * kspan;
    const float* rowA = p.w_hh_t[d] + (size_t)(j0 + g) * G + kw0 + 4 * t;
    const float* rowB = rowA + (size_t)8 * G;
    const float* brow[3];
    bool bok[3];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      const int b = 8 * nt + g;
      bok[nt] = b < B;
      brow[nt] = dg_next + (size_t)(bok[nt] ? b : 0) * G + kw0 + 4 * t;
    }
    float acc[3][4];
#pragma unroll
    for (int nt = 0; nt < 3; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    const int groups = kspan >> 4;
    constexpr int PF = 4;                               // groups of loads kept in flight
    float4 wa[PF], wb[PF], bv[PF][3];
    auto load = [&](int gi, int slot) {
      wa[slot] = __ldg(reinterpret_cast<const float4*>(rowA + 16 * gi));
      wb[slot] = __ldg(reinterpret_cast<const float4*>(rowB + 16 * gi));
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
        bv[slot][nt] = bok[nt] ? *reinterpret_cast<const float4*>(brow[nt] + 16 * gi) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
#pragma unroll
    for (int i = 0; i < PF; ++i)
      if (i < groups) load(i, i);
    for (int g0 = 0; g0 < groups; g0 += PF) {
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const int gi = g0 + i;
        if (gi >= groups) break;
        const float4 ca = wa[i], cb = wb[i];
        float4 ch[3];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) ch[nt] = bv[i][nt];
        if (gi + PF < groups) load(gi + PF, i);
        const uint32_t a0 = to_tf32(ca.x), a1 = to_tf32(cb.x), a2 = to_tf32(ca.y), a3 = to_tf32(cb.y);
        const uint32_t e0 = to_tf32(ca.z), e1 = to_tf32(cb.z), e2 = to_tf32(ca.w), e3 = to_tf32(cb.w);
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) {
          mma_tf32(acc[nt], a0, a1, a2, a3, to_tf32(ch[nt].x), to_tf32(ch[nt].y));
          mma_tf32(acc[nt], e0, e1, e2, e3, to_tf32(ch[nt].z), to_tf32(ch[nt].w));
        }
      }
    }
#pragma unroll
    for (int nt = 0; nt < 3; ++nt) {
      part[warp][g][8 * nt + 2 * t] = acc[nt][0];
      part[warp][g][8 * nt + 2 * t + 1] = acc[nt][1];
      part[warp][g + 8][8 * nt + 2 * t] = acc[nt][2];
      part[warp][g + 8][8 * nt + 2 * t + 1] = acc[nt][3];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < UNITS * NB; o += THREADS) {
      const int r = o / NB, c = o % NB;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][r][c];
      red[r][c] = v;
    }
  }
  __syncthreads();
  const int par = s & 1;
  for (int tt = threadIdx.x; tt < UNITS * B; tt += THREADS) {
    const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
    const size_t sb = (size_t)b * H + j;
    float dh, dc;
    if (last) {
      dh = p.dh_fin[d] ? p.dh_fin[d][sb] : 0.f;
      dc = p.dc_fin[d] ? p.dc_fin[d][sb] : 0.f;
    } else {
      dh = red[u][b] + p.dh_pass[d][(size_t)(par ^ 1) * B * H + sb];
      dc = p.dc_work[d][(size_t)(par ^ 1) * B * H + sb];
    }
    float* dg = p.dgates[d] + ((size_t)s * B + b) * G;
    float* dh_pass = p.dh_pass[d] + (size_t)par * B * H;
    float* dc_out = p.dc_work[d] + (size_t)par * B * H;
    if (l >= p.lengths[b]) {
      dg[j] = 0.f; dg[H + j] = 0.f; dg[2 * H + j] = 0.f; dg[3 * H + j] = 0.f;
      dh_pass[sb] = dh;
      dc_out[sb] = dc;
      continue;
    }
    dh += p.dout[((size_t)b * L + l) * 2 * H + (size_t)d * H + j];
    const float* a = p.acts[d] + ((size_t)s * B + b) * G;
    const float ig = a[j], fg = a[H + j], gg = a[2 * H + j], og = a[3 * H + j];
    const float cp = p.cs[d][(size_t)s * B * H + sb];
    const float tc = tanhf(p.cs[d][(size_t)(s + 1) * B * H + sb]);
    const float dct = dc + dh * og * (1.f - tc * tc);
    dg[j] = dct * gg * ig * (1.f - ig);
    dg[H + j] = dct * cp * fg * (1.f - fg);
    dg[2 * H + j] = dct * ig * (1.f - gg * gg);
    dg[3 * H + j] = dh * tc * og * (1.f - og);
    dh_pass[sb] = 0.f;
    dc_out[sb] = dct * fg;
  }
}

// ================================================================================================ persistent form (TF32 mode)
// The whole time loop in ONE cooperative launch each way (B <= 20, H % 128 == 0, H <= 1024): the same 2 x H/16 CTAs stay
// resident, keep their 64 recurrent-weight rows (forward) / their 16 rows of W_hh^T (backward) in SHARED MEMORY as fp16 for all L
// steps (128 KB; streamed from L2 as 256 KB of fp32 in every step before), exchange the state / the gate gradients of a step as
// fp16 rows through L2 (h in (-1, 1); gate gradients scaled by 2^8, saturating: fp16 keeps TF32's 11 significant bits) and
// separate the steps with a device-wide barrier instead of a kernel boundary (11.5 / 16.6 us per step -> see DESIGN.md).
// Products on mma.sync.m16n8k16 f16 x f16 -> f32; pointwise math, state and every saved tensor in fp32 exactly as the per-step
// kernels write them.
constexpr int PS_MAXH = 1024;
__device__ __half g_ps_x16[2][2][NB][PS_MAXH];            // forward: state rows [ping-pong][direction][b][H]
__device__ __half g_ps_g16[2][2][NB][4 * PS_MAXH];        // backward: scaled gate gradients [ping-pong][direction][b][4H]
__device__ unsigned int g_ps_bar[2];                      // barrier counters (forward, backward), zeroed before each launch
constexpr float PS_GSCALE = 256.f;

struct PsBar { unsigned int* ctr; unsigned int target, nblk; };
__device__ __forceinline__ void ps_grid_sync(PsBar& gb) {
  gb.target += gb.nblk;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
    unsigned int v, it = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      if (++it > (1u << 22)) __trap();                   // a protocol bug traps instead of hanging the device
    } while ((int)(v - gb.target) < 0);
  }
  __syncthreads();
}
// The same barrier in two halves: everything a CTA stores BEFORE ps_arrive is visible to the others after their ps_wait; stores
// issued between the two halves (tensors nobody reads inside the launch) drain under the wait instead of in front of the arrive.
__device__ __forceinline__ void ps_arrive(PsBar& gb) {
  gb.target += gb.nblk;
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gb.ctr) : "memory");
}
__device__ __forceinline__ void ps_wait(PsBar& gb) {
  if (threadIdx.x == 0) {
    unsigned int v, it = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gb.ctr) : "memory");
      if (++it > (1u << 22)) __trap();
    } while ((int)(v - gb.target) < 0);
  }
  __syncthreads();
}
__device__ __forceinline__ void ps_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t ps_u4(const uint4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }
__device__ __forceinline__ uint4 ps_ldcg16(const __half* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// Forward. Shared memory: Ws[64][H + 32] halves (row pitch = 64 B mod 128 B: the 8 lanes of a quarter warp hit 32 distinct banks),
// part[8][64][NB + 1] floats. Lane (g, t) owns 8 consecutive k of its rows per 32-wide chunk (k slots of the MMA are a fixed
// bijection of the reduction index, the same for both operands).
__global__ void __launch_bounds__(THREADS, 1) bilstm_persist_fwd_kernel(SeqFwd p) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const int B = p.B, L = p.L, H = p.H;
  const int pitch = H + 32;
  __half* Ws = reinterpret_cast<__half*>(ps_smem);
  float (*part)[64][NB + 1] = reinterpret_cast<float (*)[64][NB + 1]>(ps_smem + (size_t)64 * pitch * sizeof(__half));
  float (*gbuf)[NB + 1] = part[0];
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PsBar gb{&g_ps_bar[0], 0u, gridDim.x * gridDim.y};
  // weights: local row r = unit * 4 + gate <- W_hh row gate * H + j0 + unit, fp32 -> fp16 once
  for (int i = threadIdx.x; i < 64 * (H >> 2); i += THREADS) {
    const int r = i / (H >> 2), k = (i % (H >> 2)) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w_hh[d] + ((size_t)(r & 3) * H + j0 + (r >> 2)) * H + k));
    __half2 h2[2] = {__floats2half2_rn(v.x, v.y), __floats2half2_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(Ws + (size_t)r * pitch + k) = *reinterpret_cast<uint2*>(h2);
  }
  // the state before step 0 is zero (hs / cs index 0 are zero-filled by the caller): its fp16 rows for this CTA's units
  for (int i = threadIdx.x; i < UNITS * NB; i += THREADS) g_ps_x16[0][d][i / UNITS][j0 + i % UNITS] = __float2half_rn(0.f);
  ps_grid_sync(gb);
  const int kspan = H >> 3, kw0 = warp * kspan, nch = kspan >> 5;       // this warp's slice of the reduction: nch chunks of 32
  for (int s = 0; s < L; ++s) {
    const int l = d == 0 ? s : L - 1 - s;
    const __half* x16 = &g_ps_x16[s & 1][d][0][0];
    __half* x16n = &g_ps_x16[(s + 1) & 1][d][0][0];
    // the pointwise operands of this thread's (unit, episode) items do not depend on the GEMM: requested now (xp and the cell
    // state come from DRAM), consumed after the fold
    float pre_x[2][4], pre_c[2], pre_h[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int tt = threadIdx.x + it * THREADS;
      pre_c[it] = 0.f; pre_h[it] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) pre_x[it][q] = 0.f;
      if (tt < UNITS * B) {
        const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
        pre_c[it] = p.cs[d][(size_t)s * B * H + (size_t)b * H + j];
        if (l < p.lengths[b]) {
          const float* xrow = p.xp[d] + ((size_t)b * L + l) * 4 * H;
#pragma unroll
          for (int q = 0; q < 4; ++q) pre_x[it][q] = __ldg(xrow + q * H + j);
        } else {
          pre_h[it] = p.hs[d][(size_t)s * B * H + (size_t)b * H + j];
        }
      }
    }
    float acc[4][3][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c >= nch) break;
      const int k = kw0 + 32 * c + 8 * t;
      uint4 xf[3];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) xf[nt] = (8 * nt + g < B) ? ps_ldcg16(x16 + (size_t)(8 * nt + g) * PS_MAXH + k) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const uint4 wa = *reinterpret_cast<const uint4*>(Ws + (size_t)(16 * mt + g) * pitch + k);
        const uint4 wb = *reinterpret_cast<const uint4*>(Ws + (size_t)(16 * mt + g + 8) * pitch + k);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t af[4] = {ps_u4(wa, 2 * j), ps_u4(wb, 2 * j), ps_u4(wa, 2 * j + 1), ps_u4(wb, 2 * j + 1)};
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) ps_mma(acc[mt][nt], af, ps_u4(xf[nt], 2 * j), ps_u4(xf[nt], 2 * j + 1));
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        part[warp][16 * mt + g][8 * nt + 2 * t] = acc[mt][nt][0];
        part[warp][16 * mt + g][8 * nt + 2 * t + 1] = acc[mt][nt][1];
        part[warp][16 * mt + g + 8][8 * nt + 2 * t] = acc[mt][nt][2];
        part[warp][16 * mt + g + 8][8 * nt + 2 * t + 1] = acc[mt][nt][3];
      }
    __syncthreads();
    for (int o = threadIdx.x; o < 64 * NB; o += THREADS) {
      const int r = o / NB, c = o % NB;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += part[w][r][c];
      gbuf[r][c] = v;
    }
    __syncthreads();
    float r_h[2], r_c[2], r_g[2][4];
    bool r_live[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {                 // phase 1: the cell update; only the fp16 state row other CTAs read is stored
      const int tt = threadIdx.x + it * THREADS;
      r_h[it] = r_c[it] = 0.f; r_live[it] = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) r_g[it][q] = 0.f;
      if (tt >= UNITS * B) break;
      const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
      const float cp = pre_c[it];
      if (l >= p.lengths[b]) {                       // packed-sequence semantics: carry state, zero output row
        r_h[it] = pre_h[it];
        r_c[it] = cp;
      } else {
        float gt[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          gt[q] = gbuf[u * 4 + q][b] + pre_x[it][q] + __ldg(p.b_ih[d] + q * H + j) + __ldg(p.b_hh[d] + q * H + j);
        const float ig = sigmoidf_(gt[0]), fg = sigmoidf_(gt[1]), gg = tanhf(gt[2]), og = sigmoidf_(gt[3]);
        r_c[it] = fg * cp + ig * gg;
        r_h[it] = og * tanhf(r_c[it]);
        r_g[it][0] = ig; r_g[it][1] = fg; r_g[it][2] = gg; r_g[it][3] = og;
        r_live[it] = true;
      }
      x16n[(size_t)b * PS_MAXH + j] = __float2half_rn(r_h[it]);
    }
    if (s + 1 < L) ps_arrive(gb);
#pragma unroll
    for (int it = 0; it < 2; ++it) {                 // phase 2: everything only later kernels (or this thread) read
      const int tt = threadIdx.x + it * THREADS;
      if (tt >= UNITS * B) break;
      const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
      const size_t sb = (size_t)b * H + j;
      p.hs[d][(size_t)(s + 1) * B * H + sb] = r_h[it];
      p.cs[d][(size_t)(s + 1) * B * H + sb] = r_c[it];
      float* a = p.acts[d] + ((size_t)s * B + b) * 4 * H;
      p.out[((size_t)b * L + l) * 2 * H + (size_t)d * H + j] = r_live[it] ? r_h[it] : 0.f;
      a[j] = r_g[it][0]; a[H + j] = r_g[it][1]; a[2 * H + j] = r_g[it][2]; a[3 * H + j] = r_g[it][3];
    }
    if (s + 1 < L) ps_wait(gb);
  }
}

// Backward. Shared memory: Wt[16][4H + 32] halves = this CTA's 16 rows of W_hh^T, part[8][16][NB + 1] + red[16][NB] floats.
__global__ void __launch_bounds__(THREADS, 1) bilstm_persist_bwd_kernel(SeqBwd p) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const int B = p.B, L = p.L, H = p.H, G = 4 * H;
  const int pitch = G + 32;
  __half* Wt = reinterpret_cast<__half*>(ps_smem);
  float (*part)[UNITS][NB + 1] = reinterpret_cast<float (*)[UNITS][NB + 1]>(ps_smem + (size_t)UNITS * pitch * sizeof(__half));
  float (*red)[NB] = reinterpret_cast<float (*)[NB]>(reinterpret_cast<float*>(part) + 8 * UNITS * (NB + 1));
  const int d = blockIdx.y, j0 = blockIdx.x * UNITS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  PsBar gb{&g_ps_bar[1], 0u, gridDim.x * gridDim.y};
  for (int i = threadIdx.x; i < UNITS * (G >> 2); i += THREADS) {
    const int r = i / (G >> 2), k = (i % (G >> 2)) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(p.w_hh_t[d] + (size_t)(j0 + r) * G + k));
    __half2 h2[2] = {__floats2half2_rn(v.x, v.y), __floats2half2_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(Wt + (size_t)r * pitch + k) = *reinterpret_cast<uint2*>(h2);
  }
  __syncthreads();
  const int kspan = G >> 3, kw0 = warp * kspan, nch = kspan >> 5;
  for (int s = L - 1; s >= 0; --s) {
    const int l = d == 0 ? s : L - 1 - s;
    const bool last = (s == L - 1);
    const int par = s & 1;
    // pointwise operands of this thread's (unit, episode) items (saved activations / cell states / output gradient come from
    // DRAM): requested before the GEMM
    float q_a[2][4], q_cp[2], q_cn[2], q_do[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int tt = threadIdx.x + it * THREADS;
      q_cp[it] = q_cn[it] = q_do[it] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) q_a[it][q] = 0.f;
      if (tt < UNITS * B) {
        const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
        if (l < p.lengths[b]) {
          const size_t sb = (size_t)b * H + j;
          const float* a = p.acts[d] + ((size_t)s * B + b) * G;
#pragma unroll
          for (int q = 0; q < 4; ++q) q_a[it][q] = __ldg(a + q * H + j);
          q_cp[it] = __ldg(p.cs[d] + (size_t)s * B * H + sb);
          q_cn[it] = __ldg(p.cs[d] + (size_t)(s + 1) * B * H + sb);
          q_do[it] = __ldg(p.dout + ((size_t)b * L + l) * 2 * H + (size_t)d * H + j);
        }
      }
    }
    if (!last) {
      // rec[unit, b] = dgates_{s+1}[b, :] . W_hh^T[unit, :] over the scaled fp16 copies every CTA published in the previous step
      const __half* x16 = &g_ps_g16[par ^ 1][d][0][0];
      float acc[3][4];
#pragma unroll
      for (int nt = 0; nt < 3; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll 8
      for (int c = 0; c < nch; ++c) {
        const int k = kw0 + 32 * c + 8 * t;
        uint4 xf[3];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
          xf[nt] = (8 * nt + g < B) ? ps_ldcg16(x16 + (size_t)(8 * nt + g) * (4 * PS_MAXH) + k) : make_uint4(0u, 0u, 0u, 0u);
        const uint4 wa = *reinterpret_cast<const uint4*>(Wt + (size_t)g * pitch + k);
        const uint4 wb = *reinterpret_cast<const uint4*>(Wt + (size_t)(g + 8) * pitch + k);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t af[4] = {ps_u4(wa, 2 * j), ps_u4(wb, 2 * j), ps_u4(wa, 2 * j + 1), ps_u4(wb, 2 * j + 1)};
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) ps_mma(acc[nt], af, ps_u4(xf[nt], 2 * j), ps_u4(xf[nt], 2 * j + 1));
        }
      }
#pragma unroll
      for (int nt = 0; nt < 3; ++nt) {
        part[warp][g][8 * nt + 2 * t] = acc[nt][0];
        part[warp][g][8 * nt + 2 * t + 1] = acc[nt][1];
        part[warp][g + 8][8 * nt + 2 * t] = acc[nt][2];
        part[warp][g + 8][8 * nt + 2 * t + 1] = acc[nt][3];
      }
      __syncthreads();
      for (int o = threadIdx.x; o < UNITS * NB; o += THREADS) {
        const int r = o / NB, c = o % NB;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += part[w][r][c];
        red[r][c] = v * (1.f / PS_GSCALE);
      }
    }
    __syncthreads();
    __half* g16 = &g_ps_g16[par][d][0][0];
    float r_d[2][4], r_dh[2], r_dc[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {                 // phase 1: gate gradients; only the scaled fp16 rows other CTAs read are stored
      const int tt = threadIdx.x + it * THREADS;
      r_dh[it] = r_dc[it] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) r_d[it][q] = 0.f;
      if (tt >= UNITS * B) break;
      const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
      const size_t sb = (size_t)b * H + j;
      float dh, dc;
      if (last) {
        dh = p.dh_fin[d] ? p.dh_fin[d][sb] : 0.f;
        dc = p.dc_fin[d] ? p.dc_fin[d][sb] : 0.f;
      } else {
        dh = red[u][b] + p.dh_pass[d][(size_t)(par ^ 1) * B * H + sb];
        dc = p.dc_work[d][(size_t)(par ^ 1) * B * H + sb];
      }
      __half* gh = g16 + (size_t)b * (4 * PS_MAXH);
      if (l >= p.lengths[b]) {                       // inactive: state was carried, pass gradients straight through
        const __half z = __float2half_rn(0.f);
        gh[j] = z; gh[H + j] = z; gh[2 * H + j] = z; gh[3 * H + j] = z;
        r_dh[it] = dh;
        r_dc[it] = dc;
        continue;
      }
      dh += q_do[it];
      const float ig = q_a[it][0], fg = q_a[it][1], gg = q_a[it][2], og = q_a[it][3];
      const float cp = q_cp[it];
      const float tc = tanhf(q_cn[it]);
      const float dct = dc + dh * og * (1.f - tc * tc);
      r_d[it][0] = dct * gg * ig * (1.f - ig);
      r_d[it][1] = dct * cp * fg * (1.f - fg);
      r_d[it][2] = dct * ig * (1.f - gg * gg);
      r_d[it][3] = dh * tc * og * (1.f - og);
      r_dh[it] = 0.f;
      r_dc[it] = dct * fg;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        unsigned short hq;
        asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hq) : "f"(r_d[it][q] * PS_GSCALE));
        gh[q * H + j] = __ushort_as_half(hq);
      }
    }
    if (s > 0) ps_arrive(gb);
#pragma unroll
    for (int it = 0; it < 2; ++it) {                 // phase 2: fp32 gate gradients (weight gradients, later) and this thread's carries
      const int tt = threadIdx.x + it * THREADS;
      if (tt >= UNITS * B) break;
      const int u = tt % UNITS, b = tt / UNITS, j = j0 + u;
      const size_t sb = (size_t)b * H + j;
      float* dg = p.dgates[d] + ((size_t)s * B + b) * G;
      dg[j] = r_d[it][0]; dg[H + j] = r_d[it][1]; dg[2 * H + j] = r_d[it][2]; dg[3 * H + j] = r_d[it][3];
      p.dh_pass[d][(size_t)par * B * H + sb] = r_dh[it];
      p.dc_work[d][(size_t)par * B * H + sb] = r_dc[it];
    }
    if (s > 0) ps_wait(gb);
  }
}

int g_ps_mode = -1;      // -1: read DASA_BILSTM_PERSIST (default 1); 0 = per-step launches; 1 = persistent kernels when they apply

bool ps_enabled(int B, int H) {
  if (g_ps_mode < 0) { const char* e = getenv("DASA_BILSTM_PERSIST"); g_ps_mode = e ? atoi(e) : 1; }
  return g_ps_mode != 0 && B <= 20 && H % 256 == 0 && H <= PS_MAXH;     // each of the 8 warps owns whole 32-wide chunks of H
}

template <typename Kern, typename Args>
int ps_launch(Kern kern, const Args& p, size_t smem, int which, cudaStream_t st, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  static unsigned int* bar = nullptr;
  if (bar == nullptr && cudaGetSymbolAddress(reinterpret_cast<void**>(&bar), g_ps_bar) != cudaSuccess) return DASA_ERR_CUDA;
  if (cudaMemsetAsync(bar + which, 0, sizeof(unsigned int), st) != cudaSuccess) return DASA_ERR_CUDA;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(p.H / UNITS), 2);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;         // all 2 x H/16 CTAs co-resident, or the launch fails (never a deadlock)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, p);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

int run_fwd_persist(const SeqFwd& p, cudaStream_t st) {
  const size_t smem = (size_t)64 * (p.H + 32) * sizeof(__half) + sizeof(float) * 8 * 64 * (NB + 1);
  return ps_launch(bilstm_persist_fwd_kernel, p, smem, 0, st, "bilstm_persist_fwd_kernel");
}
int run_bwd_persist(const SeqBwd& p, cudaStream_t st) {
  const size_t smem = (size_t)UNITS * (4 * p.H + 32) * sizeof(__half) + sizeof(float) * (8 * UNITS * (NB + 1) + UNITS * NB);
  return ps_launch(bilstm_persist_bwd_kernel, p, smem, 1, st, "bilstm_persist_bwd_kernel");
}

int run_fwd_tc(const SeqFwd& p, cudaStream_t st) {
  dim3 grid((unsigned)(p.H / UNITS), 2);
  constexpr size_t smem = sizeof(float) * 8 * 64 * (NB + 1);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(bilstm_step_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dasa_set_error("bilstm fwd tc attr", e); return DASA_ERR_CUDA; }
    attr = true;
  }
  for (int s = 0; s < p.L; ++s) bilstm_step_fwd_tc_kernel<<<grid, THREADS, smem, st>>>(p, s);
  return dasa_check_launch("bilstm_step_fwd_tc_kernel");
}

int run_bwd_tc(const SeqBwd& p, cudaStream_t st) {
  dim3 grid((unsigned)(p.H / UNITS), 2);
  for (int s = p.L - 1; s >= 0; --s) bilstm_step_bwd_tc_kernel<<<grid, THREADS, 0, st>>>(p, s);
  return dasa_check_launch("bilstm_step_bwd_tc_kernel");
}

template <int BT>
int run_fwd(const SeqFwd& p, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)BT * p.H + 64 * BT);
  cudaError_t e = cudaFuncSetAttribute(bilstm_step_fwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error("bilstm fwd attr", e); return DASA_ERR_CUDA; }
  dim3 grid((unsigned)(p.H / UNITS), 2);
  for (int s = 0; s < p.L; ++s) bilstm_step_fwd_kernel<BT><<<grid, THREADS, smem, st>>>(p, s);
  return dasa_check_launch("bilstm_step_fwd_kernel");
}

template <int BT>
int run_bwd(const SeqBwd& p, cudaStream_t st) {
  const int G = 4 * p.H, CH = G < 1024 ? G : 1024;
  const size_t smem = sizeof(float) * ((size_t)BT * CH + 8 * UNITS * BT);
  cudaError_t e = cudaFuncSetAttribute(bilstm_step_bwd_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error("bilstm bwd attr", e); return DASA_ERR_CUDA; }
  dim3 grid((unsigned)(p.H / UNITS), 2);
  for (int s = p.L - 1; s >= 0; --s) bilstm_step_bwd_kernel<BT><<<grid, THREADS, smem, st>>>(p, s);
  return dasa_check_launch("bilstm_step_bwd_kernel");
}

}  // namespace

extern "C" int dasa_bilstm_max_batch(void) { return 20; }

extern "C" int dasa_debug_bilstm_persist(int mode) {
  g_ps_mode = mode;
  return DASA_OK;
}

extern "C" int dasa_bilstm_seq_fwd(const dasa_bilstm_fwd_t* a, int precision, void* stream) {
  if (a == nullptr || a->B <= 0 || a->L <= 0) return DASA_ERR_BAD_SHAPE;
  if (a->B > 20 || a->H % 64 != 0) return DASA_ERR_UNSUPPORTED;
  SeqFwd p;
  for (int d = 0; d < 2; ++d) {
    p.xp[d] = a->xp[d]; p.w_hh[d] = a->w_hh[d]; p.b_ih[d] = a->b_ih[d]; p.b_hh[d] = a->b_hh[d];
    p.hs[d] = a->hs[d]; p.cs[d] = a->cs[d]; p.acts[d] = a->acts[d];
    if (!dasa_aligned16(p.w_hh[d]) || !dasa_aligned16(p.hs[d])) return DASA_ERR_BAD_ALIGN;
  }
  p.out = a->out; p.lengths = a->lengths; p.B = a->B; p.L = a->L; p.H = a->H;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == DASA_PREC_TF32 && ps_enabled(p.B, p.H)) return run_fwd_persist(p, st);
  if (precision == DASA_PREC_TF32 && p.H % 128 == 0) return run_fwd_tc(p, st);
  if (p.B <= 4) return run_fwd<4>(p, st);
  if (p.B <= 8) return run_fwd<8>(p, st);
  if (p.B <= 12) return run_fwd<12>(p, st);
  if (p.B <= 16) return run_fwd<16>(p, st);
  return run_fwd<20>(p, st);
}

extern "C" int dasa_bilstm_seq_bwd(const dasa_bilstm_bwd_t* a, int precision, void* stream) {
  if (a == nullptr || a->B <= 0 || a->L <= 0) return DASA_ERR_BAD_SHAPE;
  if (a->B > 20 || a->H % 64 != 0) return DASA_ERR_UNSUPPORTED;
  SeqBwd p;
  for (int d = 0; d < 2; ++d) {
    p.w_hh_t[d] = a->w_hh_t[d]; p.acts[d] = a->acts[d]; p.cs[d] = a->cs[d];
    p.dh_fin[d] = a->dh_fin[d]; p.dc_fin[d] = a->dc_fin[d];
    p.dgates[d] = a->dgates[d]; p.dh_pass[d] = a->dh_pass[d]; p.dc_work[d] = a->dc_work[d];
    if (!dasa_aligned16(p.w_hh_t[d]) || !dasa_aligned16(p.dgates[d])) return DASA_ERR_BAD_ALIGN;
  }
  p.dout = a->dout; p.lengths = a->lengths; p.B = a->B; p.L = a->L; p.H = a->H;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == DASA_PREC_TF32 && ps_enabled(p.B, p.H)) return run_bwd_persist(p, st);
  if (precision == DASA_PREC_TF32 && p.H % 128 == 0) return run_bwd_tc(p, st);
  if (p.B <= 4) return run_bwd<4>(p, st);
  if (p.B <= 8) return run_bwd<8>(p, st);
  if (p.B <= 12) return run_bwd<12>(p, st);
  if (p.B <= 16) return run_bwd<16>(p, st);
  return run_bwd<20>(p, st);
}

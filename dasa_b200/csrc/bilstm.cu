// Fused recurrence of the encoder's packed bidirectional LSTM (r2rmodel.py:2339-2357) for small batches (B <= 20).
// One launch per time step handles BOTH directions: every CTA owns 16 hidden units of one direction, keeps the previous
// hidden state (fwd) / the gate gradients (bwd) of the whole batch in shared memory, streams its 64 recurrent-weight rows
// (256 KB, L2-resident across the 80 steps) with 128-bit loads, and applies the LSTM pointwise math for its units in the
// same kernel. Exact fp32 (FFMA). The whole sequence loop is issued from one C call, so the host cost is one call per
// encoder pass instead of ~320.
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
  if (precision == DASA_PREC_TF32 && p.H % 128 == 0) return run_bwd_tc(p, st);
  if (p.B <= 4) return run_bwd<4>(p, st);
  if (p.B <= 8) return run_bwd<8>(p, st);
  if (p.B <= 12) return run_bwd<12>(p, st);
  if (p.B <= 16) return run_bwd<16>(p, st);
  return run_bwd<20>(p, st);
}

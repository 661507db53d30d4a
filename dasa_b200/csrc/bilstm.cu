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

extern "C" int dasa_bilstm_seq_fwd(const dasa_bilstm_fwd_t* a, void* stream) {
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
  if (p.B <= 4) return run_fwd<4>(p, st);
  if (p.B <= 8) return run_fwd<8>(p, st);
  if (p.B <= 12) return run_fwd<12>(p, st);
  if (p.B <= 16) return run_fwd<16>(p, st);
  return run_fwd<20>(p, st);
}

extern "C" int dasa_bilstm_seq_bwd(const dasa_bilstm_bwd_t* a, void* stream) {
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
  if (p.B <= 4) return run_bwd<4>(p, st);
  if (p.B <= 8) return run_bwd<8>(p, st);
  if (p.B <= 12) return run_bwd<12>(p, st);
  if (p.B <= 16) return run_bwd<16>(p, st);
  return run_bwd<20>(p, st);
}

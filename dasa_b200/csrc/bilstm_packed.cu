// Padding-free recurrence of the encoder's packed bidirectional LSTM (r2rmodel.py:2339-2357; include/dasa_b200.h
// dasa_bilstm_packed_fwd / _bwd) for the batched teacher-forced schedule (R = T x B = 700 instruction copies at once).
//
// nn.LSTM over a PackedSequence touches only the valid tokens. bilstm_gemm.cu ran every time step over all R rows of an
// [R, L] padded grid (43 % of the row-steps were padding at the benchmark's length distribution). Here the sequences are ranked
// by length (descending) and every per-token array is stored in "position-block" order: block p holds the tokens at position p
// of the n[p] sequences longer than p, rank-major, at rows off[p] .. off[p] + n[p] of a compact [N_tokens, .] array. Then
//   * the forward direction's step s works on block s, the reverse direction's step s on block L-1-s: live rows are always a
//     PREFIX of the ranks, so each step is one grouped tcgen05 GEMM with its own row count per direction (gemm_tc2.cu
//     dasa_gemm_tc_pair_grouped2: M0 = n[s] shrinking, M1 = n[L-1-s] growing) on contiguous operand blocks;
//   * the input projections x W_ih^T, the weight gradients dgates^T [x | h_prev] and dX run over N_tokens rows instead of R x L;
//   * a sequence's final state is handed to h_fin / c_fin by the step that consumes its last token; in the reverse direction
//     a sequence joins at its own last token with a zero state written by the previous step's pointwise kernel.
// One C call issues the whole sequence loop (2 launches per step). TF32 products, fp32 state and pointwise math, the same
// summation order per gate as bilstm_gemm.cu.
#include "common.cuh"
#include "gemm_common.cuh"
#include "rng.cuh"
#include "lstm_epi.cuh"

namespace {

struct PkFwd {
  const float* xp[2]; const float* gh[2]; const float* b_ih[2]; const float* b_hh[2];
  const float* c_prev[2]; float* c_out[2]; float* acts[2];
  float* h_next[2];          // rows of the block the next step reads (nullptr: this step consumes every live sequence's last token)
  float* h_fin[2]; float* c_fin[2];
  float* out; const int32_t* perm;
  int pos[2], n[2], n_next[2];
  int L, H;
  PkDrop drop;
};

// thread = 4 consecutive hidden units of one (direction, rank); all global accesses are 128-bit
__global__ void __launch_bounds__(256) bilstm_packed_pointwise_fwd_kernel(PkFwd p) {
  const int d = blockIdx.y;
  const int H = p.H, H4 = H >> 2;
  const int n = p.n[d], n_next = p.n_next[d];
  const int rows = n > n_next ? n : n_next;
  const int64_t total = (int64_t)rows * H4;
  const int l = p.pos[d];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / H4), j = (int)(idx % H4) * 4;
    const int64_t sb = (int64_t)r * H + j;
    if (r >= n) {   // reverse direction: this sequence joins at the NEXT step with a zero state
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(p.h_next[d] + sb) = z;
      *reinterpret_cast<float4*>(p.c_out[d] + sb) = z;
      continue;
    }
    const float4 cp = *reinterpret_cast<const float4*>(p.c_prev[d] + sb);
    const float* xrow = p.xp[d] + (int64_t)r * 4 * H + j;
    const float* grow = p.gh[d] + (int64_t)r * 4 * H + j;
    float g[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 x = ldg_stream4(xrow + q * H);
      const float4 rr = ldg_stream4(grow + q * H);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b_ih[d] + q * H + j));
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.b_hh[d] + q * H + j));
      g[q][0] = ((x.x + rr.x) + b1.x) + b2.x; g[q][1] = ((x.y + rr.y) + b1.y) + b2.y;
      g[q][2] = ((x.z + rr.z) + b1.z) + b2.z; g[q][3] = ((x.w + rr.w) + b1.w) + b2.w;
    }
    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
    float ig[4], fg[4], gg[4], og[4], c1[4], h1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ig[e] = sigmoidf_(g[0][e]); fg[e] = sigmoidf_(g[1][e]); gg[e] = tanhf(g[2][e]); og[e] = sigmoidf_(g[3][e]);
      c1[e] = fg[e] * cpv[e] + ig[e] * gg[e];
      h1[e] = og[e] * tanhf(c1[e]);
    }
    const float4 h4 = make_float4(h1[0], h1[1], h1[2], h1[3]);
    const float4 c4 = make_float4(c1[0], c1[1], c1[2], c1[3]);
    *reinterpret_cast<float4*>(p.c_out[d] + sb) = c4;
    if (r < n_next) {
      *reinterpret_cast<float4*>(p.h_next[d] + sb) = h4;
    } else {        // last token of this sequence in this direction
      *reinterpret_cast<float4*>(p.h_fin[d] + sb) = h4;
      *reinterpret_cast<float4*>(p.c_fin[d] + sb) = c4;
    }
    const int64_t oe = ((int64_t)__ldg(p.perm + r) * p.L + l) * 2 * H + (int64_t)d * H + j;
    *reinterpret_cast<float4*>(p.out + oe) = pk_drop4(p.drop, oe, h4);
    float* arow = p.acts[d] + (int64_t)r * 4 * H + j;
    *reinterpret_cast<float4*>(arow) = make_float4(ig[0], ig[1], ig[2], ig[3]);
    *reinterpret_cast<float4*>(arow + H) = make_float4(fg[0], fg[1], fg[2], fg[3]);
    *reinterpret_cast<float4*>(arow + 2 * H) = make_float4(gg[0], gg[1], gg[2], gg[3]);
    *reinterpret_cast<float4*>(arow + 3 * H) = make_float4(og[0], og[1], og[2], og[3]);
  }
}

struct PkBwd {
  const float* part[2]; int nparts; int64_t part_stride;   // dh partial sums of the previous GEMM ([nparts][R][H])
  const float* dh_fin[2]; const float* dc_fin[2];          // [R, H] rank order (may be nullptr)
  const float* dc_in[2];                                   // carried dc
  const float* acts[2]; const float* c_prev[2]; const float* c_new[2];
  float* dgates[2]; float* dc_out[2];
  const float* dout; const int32_t* perm;
  int pos[2], n[2], n_carried[2];                          // rows < n_carried continue a gradient; the others start from d*_fin
  int L, H;
  PkDrop drop;
  // fp16 recurrence (optional): dgates also as fp16 scaled by PK_GSCALE (saturating) = the A operand of the next dh GEMM on
  // tcgen05 kind::f16; the partial sums then carry the same factor and are multiplied by part_scale = 1 / PK_GSCALE (exact)
  __half* dg16[2];
  float part_scale;
};
// 2^8: fp16's normal range then covers |dgate| in [2.4e-7, 256) with the 11 significant bits TF32 keeps; smaller values keep an
// absolute step of 2.3e-10, larger ones saturate
constexpr float PK_GSCALE = 256.f;
__device__ __forceinline__ uint2 pk_half4_sat(float a, float b, float c, float d) {
  uint2 r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(b), "f"(a));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(d), "f"(c));
  return r;
}

__global__ void __launch_bounds__(256) bilstm_packed_pointwise_bwd_kernel(PkBwd p) {
  const int d = blockIdx.y;
  const int H = p.H, H4 = H >> 2;
  const int n = p.n[d], nc = p.n_carried[d];
  const int64_t total = (int64_t)n * H4;
  const int l = p.pos[d];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / H4), j = (int)(idx % H4) * 4;
    const int64_t sb = (int64_t)r * H + j;
    float dh[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < nc) {
      for (int k = 0; k < p.nparts; ++k) {                    // fixed summation order: deterministic
        const float4 v = ldg_stream4(p.part[d] + (int64_t)k * p.part_stride + sb);
        dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) dh[e] *= p.part_scale;
      const float4 v = *reinterpret_cast<const float4*>(p.dc_in[d] + sb);
      dc[0] = v.x; dc[1] = v.y; dc[2] = v.z; dc[3] = v.w;
    } else {                                                  // the step that produced this sequence's final state
      if (p.dh_fin[d] != nullptr) {
        const float4 v = *reinterpret_cast<const float4*>(p.dh_fin[d] + sb);
        dh[0] = v.x; dh[1] = v.y; dh[2] = v.z; dh[3] = v.w;
      }
      if (p.dc_fin[d] != nullptr) {
        const float4 v = *reinterpret_cast<const float4*>(p.dc_fin[d] + sb);
        dc[0] = v.x; dc[1] = v.y; dc[2] = v.z; dc[3] = v.w;
      }
    }
    const int64_t oe = ((int64_t)__ldg(p.perm + r) * p.L + l) * 2 * H + (int64_t)d * H + j;
    const float4 go = pk_drop4(p.drop, oe, ldg_stream4(p.dout + oe));
    dh[0] += go.x; dh[1] += go.y; dh[2] += go.z; dh[3] += go.w;
    const float* a = p.acts[d] + (int64_t)r * 4 * H + j;
    const float4 i4 = ldg_stream4(a), f4 = ldg_stream4(a + H), g4 = ldg_stream4(a + 2 * H), o4 = ldg_stream4(a + 3 * H);
    const float4 cp4 = *reinterpret_cast<const float4*>(p.c_prev[d] + sb);
    const float4 cn4 = *reinterpret_cast<const float4*>(p.c_new[d] + sb);
    const float ig[4] = {i4.x, i4.y, i4.z, i4.w}, fg[4] = {f4.x, f4.y, f4.z, f4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w},
                og[4] = {o4.x, o4.y, o4.z, o4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w}, cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w};
    float di[4], df[4], dgg[4], dO[4], dcp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float tc = tanhf(cn[e]);
      const float dct = dc[e] + dh[e] * og[e] * (1.f - tc * tc);
      di[e] = dct * gg[e] * ig[e] * (1.f - ig[e]);
      df[e] = dct * cp[e] * fg[e] * (1.f - fg[e]);
      dgg[e] = dct * ig[e] * (1.f - gg[e] * gg[e]);
      dO[e] = dh[e] * tc * og[e] * (1.f - og[e]);
      dcp[e] = dct * fg[e];
    }
    if (p.dgates[d] != nullptr) {          // fp32 copy: only when somebody reads it (dX in the finetune configuration, TF32 path)
      float* dg = p.dgates[d] + (int64_t)r * 4 * H + j;
      *reinterpret_cast<float4*>(dg) = make_float4(di[0], di[1], di[2], di[3]);
      *reinterpret_cast<float4*>(dg + H) = make_float4(df[0], df[1], df[2], df[3]);
      *reinterpret_cast<float4*>(dg + 2 * H) = make_float4(dgg[0], dgg[1], dgg[2], dgg[3]);
      *reinterpret_cast<float4*>(dg + 3 * H) = make_float4(dO[0], dO[1], dO[2], dO[3]);
    }
    if (p.dg16[d] != nullptr) {
      __half* dh16 = p.dg16[d] + (int64_t)r * 4 * H + j;
      const float s = PK_GSCALE;
      *reinterpret_cast<uint2*>(dh16) = pk_half4_sat(di[0] * s, di[1] * s, di[2] * s, di[3] * s);
      *reinterpret_cast<uint2*>(dh16 + H) = pk_half4_sat(df[0] * s, df[1] * s, df[2] * s, df[3] * s);
      *reinterpret_cast<uint2*>(dh16 + 2 * H) = pk_half4_sat(dgg[0] * s, dgg[1] * s, dgg[2] * s, dgg[3] * s);
      *reinterpret_cast<uint2*>(dh16 + 3 * H) = pk_half4_sat(dO[0] * s, dO[1] * s, dO[2] * s, dO[3] * s);
    }
    *reinterpret_cast<float4*>(p.dc_out[d] + sb) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
  }
}

inline dim3 pk_grid(int rows, int H) {
  int64_t g = dasa_cdiv((int64_t)(rows < 1 ? 1 : rows) * (H / 4), 256);
  const int64_t cap = (int64_t)DASA_NUM_SMS * 4;
  return dim3((unsigned)(g < 1 ? 1 : (g > cap ? cap : g)), 2);
}

constexpr int PK_BWD_SPLITS = 3;

// n_rows must be non-increasing, n_rows[0] == R, off the running sum
int pk_check(int R, int L, int H, const int32_t* n_rows, const int64_t* off) {
  if (R <= 0 || L <= 0 || H <= 0 || n_rows == nullptr || off == nullptr) return DASA_ERR_BAD_SHAPE;
  if (H % 32 != 0) return DASA_ERR_UNSUPPORTED;
  if (n_rows[0] != R || off[0] != 0) return DASA_ERR_BAD_SHAPE;
  for (int p = 0; p < L; ++p) {
    if (n_rows[p] < 0 || (p > 0 && n_rows[p] > n_rows[p - 1]) || off[p + 1] != off[p] + n_rows[p]) return DASA_ERR_BAD_SHAPE;
  }
  return DASA_OK;
}

inline PkDrop pk_drop(const uint8_t* mask, const uint64_t* seed_dev, uint64_t seed, uint64_t base, float p, float scale) {
  PkDrop d;
  d.mask = mask; d.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev); d.seed = seed; d.base = base;
  d.thr = (uint32_t)(p * 65536.0f); d.stream = (mask == nullptr && p > 0.f) ? 1 : 0;
  d.scale = (mask != nullptr || d.stream) ? scale : 1.f;
  return d;
}

// out[c][k] = fp16(W_hh[gate * H + u][k]) with c = nt * 256 + ch * 128 + gate * 32 + ul, u = nt * 64 + ch * 32 + ul (lstm_epi.cuh)
__global__ void __launch_bounds__(256) lstm_whh_interleave_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int H) {
  const int64_t total = (int64_t)4 * H * (H >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i / (H >> 2)), k = (int)(i % (H >> 2)) * 4;
    const int nt = c >> 8, ch = (c >> 7) & 1, gate = (c >> 5) & 3, ul = c & 31;
    const int u = nt * 64 + ch * 32 + ul;
    const float4 v = *reinterpret_cast<const float4*>(w + ((int64_t)gate * H + u) * H + k);
    __half2 h[2] = {__floats2half2_rn(v.x, v.y), __floats2half2_rn(v.z, v.w)};
    *reinterpret_cast<uint2*>(out + (int64_t)c * H + k) = *reinterpret_cast<uint2*>(h);
  }
}

}  // namespace

extern "C" int dasa_lstm_whh_interleave_f16(const float* w_hh, dasa_half_t* out, int H, void* stream) {
  if (w_hh == nullptr || out == nullptr || H < 64 || (H % 64) != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(w_hh) || !dasa_aligned16(out)) return DASA_ERR_BAD_ALIGN;
  const int64_t work = (int64_t)H * H;
  const unsigned grid = (unsigned)(dasa_cdiv(work, 256) < (int64_t)DASA_NUM_SMS * 8 ? dasa_cdiv(work, 256) : (int64_t)DASA_NUM_SMS * 8);
  lstm_whh_interleave_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_hh, reinterpret_cast<__half*>(out), H);
  return dasa_check_launch("lstm_whh_interleave_f16_kernel");
}

extern "C" size_t dasa_bilstm_packed_workspace(int R, int H, int backward) {
  if (R <= 0 || H <= 0) return 0;
  return backward ? (size_t)2 * PK_BWD_SPLITS * R * H * sizeof(float) : (size_t)2 * R * 4 * H * sizeof(float);
}

extern "C" int dasa_bilstm_packed_fwd(const dasa_bilstm_packed_fwd_t* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (a == nullptr) return DASA_ERR_BAD_SHAPE;
  const int R = a->R, L = a->L, H = a->H;
  int rc = pk_check(R, L, H, a->n_rows, a->off);
  if (rc != DASA_OK) return rc;
  if (workspace == nullptr || workspace_bytes < dasa_bilstm_packed_workspace(R, H, 0)) return DASA_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* n = a->n_rows;
  const int64_t* off = a->off;
  int Le = L;                                              // positions beyond the longest sequence hold no token
  while (Le > 1 && n[Le - 1] == 0) --Le;
  float* gh[2] = {static_cast<float*>(workspace), static_cast<float*>(workspace) + (size_t)R * 4 * H};
  const float* Bw[2] = {a->w_hh[0], a->w_hh[1]};
  const size_t RH = (size_t)R * H;
  // fused form (lstm_epi.cuh): fp16 state rows x interleaved fp16 W_hh on tcgen05 kind::f16, LSTM cell on the accumulator:
  // ONE launch per time step, gh never written. h in (-1, 1) and fp16 carries TF32's 10 mantissa bits, so the recurrent
  // products are the TF32 kernel's (below 2^-14 the absolute rounding step stays 6e-8).
  const bool fused = a->w_hh16[0] != nullptr && a->w_hh16[1] != nullptr && a->h16[0] != nullptr && a->h16[1] != nullptr &&
                     H >= 64 && (H % 64) == 0;
  if (fused) {
    __half* h16[2] = {reinterpret_cast<__half*>(a->h16[0]), reinterpret_cast<__half*>(a->h16[1])};
    const __half* W16[2] = {reinterpret_cast<const __half*>(a->w_hh16[0]), reinterpret_cast<const __half*>(a->w_hh16[1])};
    cudaMemsetAsync(a->hprev[0], 0, (size_t)n[0] * H * sizeof(float), st);
    cudaMemsetAsync(h16[0], 0, (size_t)n[0] * H * sizeof(__half), st);
    cudaMemsetAsync(a->cs[0], 0, (size_t)n[0] * H * sizeof(float), st);
    cudaMemsetAsync(a->hprev[1] + off[Le - 1] * H, 0, (size_t)n[Le - 1] * H * sizeof(float), st);
    cudaMemsetAsync(h16[1] + off[Le - 1] * H, 0, (size_t)n[Le - 1] * H * sizeof(__half), st);
    cudaMemsetAsync(a->cs[1], 0, (size_t)n[Le - 1] * H * sizeof(float), st);
    for (int s = 0; s < Le; ++s) {
      const int pos[2] = {s, Le - 1 - s};
      LstmEpi le;
      const __half* A16[2];
      for (int d = 0; d < 2; ++d) {
        const int pp = pos[d];
        const int nx = d == 0 ? pp + 1 : pp - 1;               // block the next step of this direction reads
        const bool has_next = nx >= 0 && nx < Le;
        A16[d] = h16[d] + off[pp] * H;
        le.xp[d] = a->xp[d] + off[pp] * 4 * H; le.b_ih[d] = a->b_ih[d]; le.b_hh[d] = a->b_hh[d];
        le.c_prev[d] = a->cs[d] + s * RH; le.c_out[d] = a->cs[d] + (s + 1) * RH;
        le.acts[d] = a->acts[d] + off[pp] * 4 * H;
        le.h_next[d] = has_next ? a->hprev[d] + off[nx] * H : nullptr;
        le.h16_next[d] = has_next ? h16[d] + off[nx] * H : nullptr;
        le.h_fin[d] = a->h_fin[d]; le.c_fin[d] = a->c_fin[d];
        le.pos[d] = pp; le.n[d] = n[pp]; le.n_next[d] = has_next ? n[nx] : 0;
      }
      le.out = a->out; le.perm = a->perm; le.L = L; le.H = H;
      le.drop = pk_drop(a->out_mask, a->drop_seed_dev, a->drop_seed, a->drop_base, a->drop_p, a->drop_scale);
      rc = dasa_gemm_tc_pair_lstm(A16, W16, le, st);
      if (rc != DASA_OK) return rc;
    }
    return DASA_OK;
  }
  // zero initial state: forward direction block 0 (every sequence), reverse direction block Le-1 (the longest sequences)
  cudaMemsetAsync(a->hprev[0], 0, (size_t)n[0] * H * sizeof(float), st);
  cudaMemsetAsync(a->cs[0], 0, (size_t)n[0] * H * sizeof(float), st);
  cudaMemsetAsync(a->hprev[1] + off[Le - 1] * H, 0, (size_t)n[Le - 1] * H * sizeof(float), st);
  cudaMemsetAsync(a->cs[1], 0, (size_t)n[Le - 1] * H * sizeof(float), st);
  for (int s = 0; s < Le; ++s) {
    const int p0 = s, p1 = Le - 1 - s;
    const float* A[2] = {a->hprev[0] + off[p0] * H, a->hprev[1] + off[p1] * H};
    rc = dasa_gemm_tc_pair_grouped2(n[p0], n[p1], 4 * H, H, A, H, Bw, H, gh, 4 * H, 1, 0, st);
    if (rc < 0) return rc;
    PkFwd p;
    const int pos[2] = {p0, p1};
    for (int d = 0; d < 2; ++d) {
      const int pp = pos[d];
      p.xp[d] = a->xp[d] + off[pp] * 4 * H; p.gh[d] = gh[d]; p.b_ih[d] = a->b_ih[d]; p.b_hh[d] = a->b_hh[d];
      p.c_prev[d] = a->cs[d] + s * RH; p.c_out[d] = a->cs[d] + (s + 1) * RH;
      p.acts[d] = a->acts[d] + off[pp] * 4 * H;
      p.h_fin[d] = a->h_fin[d]; p.c_fin[d] = a->c_fin[d];
      p.pos[d] = pp; p.n[d] = n[pp];
    }
    // forward direction hands h to block p0 + 1 (fewer rows: the others just consumed their last token)
    p.n_next[0] = (p0 + 1 < Le) ? n[p0 + 1] : 0;
    p.h_next[0] = (p0 + 1 < Le) ? a->hprev[0] + off[p0 + 1] * H : nullptr;
    // reverse direction hands h to block p1 - 1 (more rows: the newcomers get a zero state)
    p.n_next[1] = (p1 >= 1) ? n[p1 - 1] : 0;
    p.h_next[1] = (p1 >= 1) ? a->hprev[1] + off[p1 - 1] * H : nullptr;
    p.out = a->out; p.perm = a->perm; p.L = L; p.H = H;
    p.drop = pk_drop(a->out_mask, a->drop_seed_dev, a->drop_seed, a->drop_base, a->drop_p, a->drop_scale);
    const int rows = p.n[0] > p.n_next[1] ? p.n[0] : p.n_next[1];
    bilstm_packed_pointwise_fwd_kernel<<<pk_grid(rows > p.n[1] ? rows : p.n[1], H), 256, 0, st>>>(p);
  }
  return dasa_check_launch("bilstm_packed_pointwise_fwd_kernel");
}

extern "C" int dasa_bilstm_packed_bwd(const dasa_bilstm_packed_bwd_t* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (a == nullptr) return DASA_ERR_BAD_SHAPE;
  const int R = a->R, L = a->L, H = a->H;
  int rc = pk_check(R, L, H, a->n_rows, a->off);
  if (rc != DASA_OK) return rc;
  if (workspace == nullptr || workspace_bytes < dasa_bilstm_packed_workspace(R, H, 1)) return DASA_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* n = a->n_rows;
  const int64_t* off = a->off;
  int Le = L;
  while (Le > 1 && n[Le - 1] == 0) --Le;
  const size_t RH = (size_t)R * H;
  float* part[2] = {static_cast<float*>(workspace), static_cast<float*>(workspace) + (size_t)PK_BWD_SPLITS * RH};
  const float* Bw[2] = {a->w_hh_t[0], a->w_hh_t[1]};          // [H, 4H]: K-major B operand of dh = dgates * W_hh
  // fp16 form of the recurrent GEMM (dgates scaled by 2^8 as the fp16 A operand, W_hh^T as fp16): twice the tensor rate on the
  // 2 x L dependent launches; the fp32 dgates (weight gradients, dX) are written as before
  const bool f16 = a->w_hh_t16[0] != nullptr && a->w_hh_t16[1] != nullptr && a->dg16[0] != nullptr && a->dg16[1] != nullptr &&
                   (H % 64) == 0;
  const __half* Bw16[2] = {reinterpret_cast<const __half*>(a->w_hh_t16[0]), reinterpret_cast<const __half*>(a->w_hh_t16[1])};
  if (!f16 && (a->dgates[0] == nullptr || a->dgates[1] == nullptr)) return DASA_ERR_BAD_SHAPE;   // the TF32 recurrence reads them
  int nparts = 0;
  for (int s = Le - 1; s >= 0; --s) {
    const int cur = (Le - 1 - s) & 1, prv = cur ^ 1;           // ping-pong halves of dc_work
    const int p0 = s, p1 = Le - 1 - s;
    const int pos[2] = {p0, p1};
    PkBwd p;
    for (int d = 0; d < 2; ++d) {
      const int pp = pos[d];
      p.part[d] = part[d];
      p.dh_fin[d] = a->dh_fin[d]; p.dc_fin[d] = a->dc_fin[d];
      p.dc_in[d] = a->dc_work[d] + prv * RH; p.dc_out[d] = a->dc_work[d] + cur * RH;
      p.acts[d] = a->acts[d] + off[pp] * 4 * H;
      p.c_prev[d] = a->cs[d] + s * RH; p.c_new[d] = a->cs[d] + (s + 1) * RH;
      p.dgates[d] = a->dgates[d] != nullptr ? a->dgates[d] + off[pp] * 4 * H : nullptr;
      p.dg16[d] = f16 ? reinterpret_cast<__half*>(a->dg16[d]) + off[pp] * 4 * H : nullptr;
      p.pos[d] = pp; p.n[d] = n[pp];
    }
    p.part_scale = f16 ? 1.f / PK_GSCALE : 1.f;
    // rows that continue a gradient from step s + 1: forward direction the sequences still alive at p0 + 1; reverse direction
    // everything alive now except at the very first backward step (position 0 produced every final state)
    p.n_carried[0] = (p0 + 1 < Le) ? n[p0 + 1] : 0;
    p.n_carried[1] = (s == Le - 1) ? 0 : n[p1];
    p.nparts = nparts; p.part_stride = (int64_t)RH;
    p.dout = a->dout; p.perm = a->perm; p.L = L; p.H = H;
    p.drop = pk_drop(a->out_mask, a->drop_seed_dev, a->drop_seed, a->drop_base, a->drop_p, a->drop_scale);
    bilstm_packed_pointwise_bwd_kernel<<<pk_grid(p.n[0] > p.n[1] ? p.n[0] : p.n[1], H), 256, 0, st>>>(p);
    if (s > 0) {                                              // dh of the states that fed step s (the gradient before step 0 is unused)
      if (f16) {
        const __half* A16[2] = {p.dg16[0], p.dg16[1]};
        nparts = dasa_gemm_tc_pair_grouped2_f16(n[p0], n[p1], H, 4 * H, A16, 4 * H, Bw16, 4 * H, part, H, PK_BWD_SPLITS, (int64_t)RH, st);
      } else {
        const float* A[2] = {p.dgates[0], p.dgates[1]};
        nparts = dasa_gemm_tc_pair_grouped2(n[p0], n[p1], H, 4 * H, A, 4 * H, Bw, 4 * H, part, H, PK_BWD_SPLITS, (int64_t)RH, st);
      }
      if (nparts < 0) return nparts;
    }
  }
  return dasa_check_launch("bilstm_packed_pointwise_bwd_kernel");
}

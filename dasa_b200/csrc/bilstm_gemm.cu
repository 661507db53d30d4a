// Large-batch recurrence of the encoder's packed bidirectional LSTM (r2rmodel.py:2339-2357): the batched teacher-forced
// schedule runs the bi-LSTM of all T x B instruction copies at once (700 sequences), so every time step is a real GEMM.
// One time step = ONE grouped launch of the persistent CTA-pair tcgen05 kernel for BOTH directions (gemm_tc2.cu) plus ONE
// pointwise launch for both directions:
//   forward : gh[d] = hs[d][s] * W_hh[d]^T (M = B, N = 4H, K = H)          -> gates, cell and hidden update (fwd kernel below)
//   backward: dh partials[d][k] = dgates[d][s] * W_hh[d] (M = B, N = H, K = 4H, split along K so the 2 x 3 x 4 output tiles fill
//             the 74 TPCs); the NEXT step's pointwise kernel sums the partials while it forms dgates[s-1], so there is no
//             separate split-K reduction or axpy pass.
// The whole sequence loop is issued from one C call. TF32 products, fp32 state and pointwise math.
#include "common.cuh"
#include "gemm_common.cuh"

namespace {

struct PwFwd {
  const float* xp[2]; const float* gh[2]; const float* b_ih[2]; const float* b_hh[2];
  const float* c_prev[2]; const float* h_prev[2]; float* h_out[2]; float* c_out[2]; float* acts[2];
  float* out; const int32_t* lengths;
  int pos[2];        // token position this step touches, per direction
  int B, L, H;
};

// thread = 4 consecutive hidden units of one (direction, sequence); all global accesses are 128-bit
__global__ void __launch_bounds__(256) bilstm_gemm_pointwise_fwd_kernel(PwFwd p) {
  const int d = blockIdx.y;
  const int H = p.H, H4 = H >> 2;
  const int64_t total = (int64_t)p.B * H4;
  const int l = p.pos[d];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H4), j = (int)(idx % H4) * 4;
    const int64_t sb = (int64_t)b * H + j;
    const bool on = l < p.lengths[b];
    const float4 cp = *reinterpret_cast<const float4*>(p.c_prev[d] + sb);
    float* arow = p.acts[d] + (int64_t)b * 4 * H + j;
    float* orow = p.out + ((int64_t)b * p.L + l) * 2 * H + (int64_t)d * H + j;
    if (!on) {   // packed-sequence semantics: the state is carried, the sequence output row is zero
      *reinterpret_cast<float4*>(p.h_out[d] + sb) = *reinterpret_cast<const float4*>(p.h_prev[d] + sb);
      *reinterpret_cast<float4*>(p.c_out[d] + sb) = cp;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(orow) = z;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(arow + q * H) = z;
      continue;
    }
    const float* xrow = p.xp[d] + ((int64_t)b * p.L + l) * 4 * H + j;
    const float* grow = p.gh[d] + (int64_t)b * 4 * H + j;
    float g[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 x = ldg_stream4(xrow + q * H);
      const float4 r = ldg_stream4(grow + q * H);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b_ih[d] + q * H + j));
      const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.b_hh[d] + q * H + j));
      // same summation order as the per-direction path: ((x + gh) + b_ih) + b_hh
      g[q][0] = ((x.x + r.x) + b1.x) + b2.x; g[q][1] = ((x.y + r.y) + b1.y) + b2.y;
      g[q][2] = ((x.z + r.z) + b1.z) + b2.z; g[q][3] = ((x.w + r.w) + b1.w) + b2.w;
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
    *reinterpret_cast<float4*>(p.h_out[d] + sb) = h4;
    *reinterpret_cast<float4*>(p.c_out[d] + sb) = make_float4(c1[0], c1[1], c1[2], c1[3]);
    *reinterpret_cast<float4*>(orow) = h4;
    *reinterpret_cast<float4*>(arow) = make_float4(ig[0], ig[1], ig[2], ig[3]);
    *reinterpret_cast<float4*>(arow + H) = make_float4(fg[0], fg[1], fg[2], fg[3]);
    *reinterpret_cast<float4*>(arow + 2 * H) = make_float4(gg[0], gg[1], gg[2], gg[3]);
    *reinterpret_cast<float4*>(arow + 3 * H) = make_float4(og[0], og[1], og[2], og[3]);
  }
}

struct PwBwd {
  const float* part[2]; int nparts; int64_t part_stride;   // dh partial sums from the previous GEMM ([nparts][B][H]) or nullptr
  const float* dh_in[2];                                   // carried pass-through dh (or dh_fin at the first step, may be nullptr)
  const float* dc_in[2];                                   // carried dc (or dc_fin, may be nullptr)
  const float* acts[2]; const float* c_prev[2]; const float* c_new[2];
  float* dgates[2]; float* dc_out[2]; float* dh_pass[2];
  const float* dout; const int32_t* lengths;
  int pos[2];
  int B, L, H;
};

__global__ void __launch_bounds__(256) bilstm_gemm_pointwise_bwd_kernel(PwBwd p) {
  const int d = blockIdx.y;
  const int H = p.H, H4 = H >> 2;
  const int64_t total = (int64_t)p.B * H4;
  const int l = p.pos[d];
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H4), j = (int)(idx % H4) * 4;
    const int64_t sb = (int64_t)b * H + j;
    const bool on = l < p.lengths[b];
    float dh[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.part[d] != nullptr) {                               // fixed summation order: deterministic
      for (int k = 0; k < p.nparts; ++k) {
        const float4 v = ldg_stream4(p.part[d] + (int64_t)k * p.part_stride + sb);
        dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
      }
    }
    if (p.dh_in[d] != nullptr) {
      const float4 v = *reinterpret_cast<const float4*>(p.dh_in[d] + sb);
      dh[0] += v.x; dh[1] += v.y; dh[2] += v.z; dh[3] += v.w;
    }
    float dc[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.dc_in[d] != nullptr) {
      const float4 v = *reinterpret_cast<const float4*>(p.dc_in[d] + sb);
      dc[0] = v.x; dc[1] = v.y; dc[2] = v.z; dc[3] = v.w;
    }
    float* dg = p.dgates[d] + (int64_t)b * 4 * H + j;
    if (!on) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(dg + q * H) = z;
      *reinterpret_cast<float4*>(p.dc_out[d] + sb) = make_float4(dc[0], dc[1], dc[2], dc[3]);
      *reinterpret_cast<float4*>(p.dh_pass[d] + sb) = make_float4(dh[0], dh[1], dh[2], dh[3]);
      continue;
    }
    const float4 go = ldg_stream4(p.dout + ((int64_t)b * p.L + l) * 2 * H + (int64_t)d * H + j);   // grad of the sequence output row
    dh[0] += go.x; dh[1] += go.y; dh[2] += go.z; dh[3] += go.w;
    const float* a = p.acts[d] + (int64_t)b * 4 * H + j;
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
    *reinterpret_cast<float4*>(dg) = make_float4(di[0], di[1], di[2], di[3]);
    *reinterpret_cast<float4*>(dg + H) = make_float4(df[0], df[1], df[2], df[3]);
    *reinterpret_cast<float4*>(dg + 2 * H) = make_float4(dgg[0], dgg[1], dgg[2], dgg[3]);
    *reinterpret_cast<float4*>(dg + 3 * H) = make_float4(dO[0], dO[1], dO[2], dO[3]);
    *reinterpret_cast<float4*>(p.dc_out[d] + sb) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    *reinterpret_cast<float4*>(p.dh_pass[d] + sb) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

inline dim3 pw_grid(int B, int H) {
  int64_t g = dasa_cdiv((int64_t)B * (H / 4), 256);
  const int64_t cap = (int64_t)DASA_NUM_SMS * 4;
  return dim3((unsigned)(g < 1 ? 1 : (g > cap ? cap : g)), 2);
}

constexpr int BWD_SPLITS = 3;     // 2 directions x ceil(B/256) x H/256 output tiles x 3 K-splits ~ one wave of the 74 TPCs at B = 700

}  // namespace

extern "C" size_t dasa_bilstm_seq_gemm_workspace(int B, int H, int backward) {
  if (B <= 0 || H <= 0) return 0;
  return backward ? (size_t)2 * BWD_SPLITS * B * H * sizeof(float) : (size_t)2 * B * 4 * H * sizeof(float);
}

extern "C" int dasa_bilstm_seq_gemm_fwd(const dasa_bilstm_fwd_t* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (a == nullptr || a->B <= 0 || a->L <= 0 || a->H <= 0) return DASA_ERR_BAD_SHAPE;
  const int B = a->B, L = a->L, H = a->H;
  if (H % 32 != 0) return DASA_ERR_UNSUPPORTED;
  if (workspace == nullptr || workspace_bytes < dasa_bilstm_seq_gemm_workspace(B, H, 0)) return DASA_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  float* gh[2] = {static_cast<float*>(workspace), static_cast<float*>(workspace) + (size_t)B * 4 * H};
  const float* Bw[2] = {a->w_hh[0], a->w_hh[1]};
  const size_t BH = (size_t)B * H;
  for (int s = 0; s < L; ++s) {
    const float* A[2] = {a->hs[0] + s * BH, a->hs[1] + s * BH};
    int rc = dasa_gemm_tc_pair_grouped(B, 4 * H, H, A, H, Bw, H, gh, 4 * H, 1, 0, st);
    if (rc < 0) return rc;
    PwFwd p;
    for (int d = 0; d < 2; ++d) {
      p.xp[d] = a->xp[d]; p.gh[d] = gh[d]; p.b_ih[d] = a->b_ih[d]; p.b_hh[d] = a->b_hh[d];
      p.c_prev[d] = a->cs[d] + s * BH; p.h_prev[d] = a->hs[d] + s * BH;
      p.h_out[d] = a->hs[d] + (s + 1) * BH; p.c_out[d] = a->cs[d] + (s + 1) * BH;
      p.acts[d] = a->acts[d] + (size_t)s * B * 4 * H;
    }
    p.out = a->out; p.lengths = a->lengths; p.pos[0] = s; p.pos[1] = L - 1 - s; p.B = B; p.L = L; p.H = H;
    bilstm_gemm_pointwise_fwd_kernel<<<pw_grid(B, H), 256, 0, st>>>(p);
  }
  return dasa_check_launch("bilstm_gemm_pointwise_fwd_kernel");
}

extern "C" int dasa_bilstm_seq_gemm_bwd(const dasa_bilstm_bwd_t* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (a == nullptr || a->B <= 0 || a->L <= 0 || a->H <= 0) return DASA_ERR_BAD_SHAPE;
  const int B = a->B, L = a->L, H = a->H;
  if (H % 32 != 0) return DASA_ERR_UNSUPPORTED;
  if (workspace == nullptr || workspace_bytes < dasa_bilstm_seq_gemm_workspace(B, H, 1)) return DASA_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t BH = (size_t)B * H;
  float* part[2] = {static_cast<float*>(workspace), static_cast<float*>(workspace) + (size_t)BWD_SPLITS * BH};
  const float* Bw[2] = {a->w_hh_t[0], a->w_hh_t[1]};          // [H, 4H]: K-major B operand of dh = dgates * W_hh
  int nparts = 0;
  for (int s = L - 1; s >= 0; --s) {
    const int cur = (L - 1 - s) & 1, prv = cur ^ 1;           // ping-pong halves of dh_pass / dc_work
    PwBwd p;
    for (int d = 0; d < 2; ++d) {
      const bool first = (s == L - 1);
      p.part[d] = first ? nullptr : part[d];
      p.dh_in[d] = first ? a->dh_fin[d] : a->dh_pass[d] + prv * BH;
      p.dc_in[d] = first ? a->dc_fin[d] : a->dc_work[d] + prv * BH;
      p.acts[d] = a->acts[d] + (size_t)s * B * 4 * H;
      p.c_prev[d] = a->cs[d] + s * BH; p.c_new[d] = a->cs[d] + (s + 1) * BH;
      p.dgates[d] = a->dgates[d] + (size_t)s * B * 4 * H;
      p.dc_out[d] = a->dc_work[d] + cur * BH; p.dh_pass[d] = a->dh_pass[d] + cur * BH;
    }
    p.nparts = nparts; p.part_stride = (int64_t)BH;
    p.dout = a->dout; p.lengths = a->lengths; p.pos[0] = s; p.pos[1] = L - 1 - s; p.B = B; p.L = L; p.H = H;
    bilstm_gemm_pointwise_bwd_kernel<<<pw_grid(B, H), 256, 0, st>>>(p);
    if (s > 0) {                                              // dh for step s-1; the state gradient before step 0 is not needed
      const float* A[2] = {p.dgates[0], p.dgates[1]};
      nparts = dasa_gemm_tc_pair_grouped(B, H, 4 * H, A, 4 * H, Bw, 4 * H, part, H, BWD_SPLITS, (int64_t)BH, st);
      if (nparts < 0) return nparts;
    }
  }
  return dasa_check_launch("bilstm_gemm_pointwise_bwd_kernel");
}

// Pointwise halves of the LSTM steps: decoder nn.LSTMCell (model.py:437,514) and the encoder's packed bidirectional
// nn.LSTM (r2rmodel.py:2339-2357). The gate pre-activations come from dasa_gemm; gate order is i,f,g,o.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) lstm_pointwise_fwd_kernel(
    const float* __restrict__ ga, int64_t ld_ga, const float* __restrict__ gb, int64_t ld_gb, const float* __restrict__ bias_a,
    const float* __restrict__ bias_b, const float* __restrict__ c_prev, int64_t ld_cp, const float* __restrict__ h_prev,
    int64_t ld_hp, float* __restrict__ h_out, int64_t ld_h, float* __restrict__ c_out, int64_t ld_c, float* __restrict__ seq_out,
    int64_t ld_seq, float* __restrict__ acts_out, int64_t ld_acts, const int32_t* __restrict__ active, int pos, int B, int H,
    const uint8_t* __restrict__ seq_mask, float seq_scale) {
  const int64_t total = (int64_t)B * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H), j = (int)(idx % H);
    const bool on = (active == nullptr) || (pos < active[b]);
    const float cp = c_prev ? c_prev[(int64_t)b * ld_cp + j] : 0.f;
    if (!on) {
      // packed-sequence semantics: state is carried, the sequence output row is zero
      h_out[(int64_t)b * ld_h + j] = h_prev ? h_prev[(int64_t)b * ld_hp + j] : 0.f;
      c_out[(int64_t)b * ld_c + j] = cp;
      if (seq_out) seq_out[(int64_t)b * ld_seq + j] = 0.f;
      if (acts_out) {
        float* a = acts_out + (int64_t)b * ld_acts;
        a[j] = 0.f; a[H + j] = 0.f; a[2 * H + j] = 0.f; a[3 * H + j] = 0.f;
      }
      continue;
    }
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v = ga[(int64_t)b * ld_ga + q * H + j];
      if (gb) v += gb[(int64_t)b * ld_gb + q * H + j];
      if (bias_a) v += __ldg(bias_a + q * H + j);
      if (bias_b) v += __ldg(bias_b + q * H + j);
      g[q] = v;
    }
    const float ig = sigmoidf_(g[0]), fg = sigmoidf_(g[1]), gg = tanhf(g[2]), og = sigmoidf_(g[3]);
    const float c1 = fg * cp + ig * gg;
    const float h1 = og * tanhf(c1);
    h_out[(int64_t)b * ld_h + j] = h1;
    c_out[(int64_t)b * ld_c + j] = c1;
    // second copy of h': the sequence output row, or (seq_mask) the dropped hidden state the next layer reads
    if (seq_out) seq_out[(int64_t)b * ld_seq + j] = seq_mask ? (seq_mask[idx] ? h1 * seq_scale : 0.f) : h1;
    if (acts_out) {
      float* a = acts_out + (int64_t)b * ld_acts;
      a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
    }
  }
}

__global__ void __launch_bounds__(256) lstm_pointwise_bwd_kernel(
    const float* __restrict__ dh, int64_t ld_dh, const float* __restrict__ dh2, int64_t ld_dh2, const float* __restrict__ dc,
    int64_t ld_dc, const float* __restrict__ acts, int64_t ld_acts, const float* __restrict__ c_prev, int64_t ld_cp,
    const float* __restrict__ c_new, int64_t ld_cn, float* __restrict__ dgates, int64_t ld_dg, float* __restrict__ dc_prev,
    int64_t ld_dcp, float* __restrict__ dh_pass, int64_t ld_dhp, const int32_t* __restrict__ active, int pos, int B, int H,
    const uint8_t* __restrict__ dh2_mask, float dh2_scale) {
  const int64_t total = (int64_t)B * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H), j = (int)(idx % H);
    const bool on = (active == nullptr) || (pos < active[b]);
    float dhv = dh ? dh[(int64_t)b * ld_dh + j] : 0.f;
    // dh2 = grad of the second copy of h': the sequence output (a constant 0 row when inactive) or the dropped hidden state
    if (dh2 && on) dhv += dh2_mask ? (dh2_mask[idx] ? dh2[(int64_t)b * ld_dh2 + j] * dh2_scale : 0.f) : dh2[(int64_t)b * ld_dh2 + j];
    const float dcv = dc ? dc[(int64_t)b * ld_dc + j] : 0.f;
    float* dg = dgates + (int64_t)b * ld_dg;
    if (!on) {
      dg[j] = 0.f; dg[H + j] = 0.f; dg[2 * H + j] = 0.f; dg[3 * H + j] = 0.f;
      dc_prev[(int64_t)b * ld_dcp + j] = dcv;
      if (dh_pass) dh_pass[(int64_t)b * ld_dhp + j] = dhv;
      continue;
    }
    const float* a = acts + (int64_t)b * ld_acts;
    const float ig = a[j], fg = a[H + j], gg = a[2 * H + j], og = a[3 * H + j];
    const float cp = c_prev ? c_prev[(int64_t)b * ld_cp + j] : 0.f;
    const float tc = tanhf(c_new[(int64_t)b * ld_cn + j]);
    const float dct = dcv + dhv * og * (1.f - tc * tc);
    dg[j] = dct * gg * ig * (1.f - ig);
    dg[H + j] = dct * cp * fg * (1.f - fg);
    dg[2 * H + j] = dct * ig * (1.f - gg * gg);
    dg[3 * H + j] = dhv * tc * og * (1.f - og);
    dc_prev[(int64_t)b * ld_dcp + j] = dct * fg;
    if (dh_pass) dh_pass[(int64_t)b * ld_dhp + j] = 0.f;
  }
}

inline unsigned ew_grid(int64_t n) {
  int64_t g = dasa_cdiv(n, 256);
  const int64_t cap = (int64_t)DASA_NUM_SMS * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int dasa_lstm_pointwise_fwd(const float* ga, int64_t ld_ga, const float* gb, int64_t ld_gb, const float* bias_a,
                                       const float* bias_b, const float* c_prev, int64_t ld_cp, const float* h_prev,
                                       int64_t ld_hp, float* h_out, int64_t ld_h, float* c_out, int64_t ld_c, float* seq_out,
                                       int64_t ld_seq, float* acts_out, int64_t ld_acts, const int32_t* active, int pos, int B,
                                       int H, const uint8_t* seq_mask, float seq_scale, void* stream) {
  if (B <= 0 || H <= 0) return DASA_OK;
  if (ga == nullptr || h_out == nullptr || c_out == nullptr) return DASA_ERR_BAD_SHAPE;
  lstm_pointwise_fwd_kernel<<<ew_grid((int64_t)B * H), 256, 0, (cudaStream_t)stream>>>(
      ga, ld_ga, gb, ld_gb, bias_a, bias_b, c_prev, ld_cp, h_prev, ld_hp, h_out, ld_h, c_out, ld_c, seq_out, ld_seq, acts_out,
      ld_acts, active, pos, B, H, seq_mask, seq_scale);
  return dasa_check_launch("lstm_pointwise_fwd_kernel");
}

extern "C" int dasa_lstm_pointwise_bwd(const float* dh, int64_t ld_dh, const float* dh2, int64_t ld_dh2, const float* dc,
                                       int64_t ld_dc, const float* acts, int64_t ld_acts, const float* c_prev, int64_t ld_cp,
                                       const float* c_new, int64_t ld_cn, float* dgates, int64_t ld_dg, float* dc_prev,
                                       int64_t ld_dcp, float* dh_pass, int64_t ld_dhp, const int32_t* active, int pos, int B,
                                       int H, const uint8_t* dh2_mask, float dh2_scale, void* stream) {
  if (B <= 0 || H <= 0) return DASA_OK;
  if (acts == nullptr || c_new == nullptr || dgates == nullptr || dc_prev == nullptr) return DASA_ERR_BAD_SHAPE;
  lstm_pointwise_bwd_kernel<<<ew_grid((int64_t)B * H), 256, 0, (cudaStream_t)stream>>>(
      dh, ld_dh, dh2, ld_dh2, dc, ld_dc, acts, ld_acts, c_prev, ld_cp, c_new, ld_cn, dgates, ld_dg, dc_prev, ld_dcp, dh_pass,
      ld_dhp, active, pos, B, H, dh2_mask, dh2_scale);
  return dasa_check_launch("lstm_pointwise_bwd_kernel");
}

// Small elementwise / per-row kernels: dropout plumbing, activation backward, masked cross-entropy + action selection
// (agent_dg.py:832-886), fused RMSprop + gradient clipping (agent_dg.py:1389-1405).
#include <cuda_fp16.h>
#include "common.cuh"
#include "rng.cuh"

namespace {

inline unsigned ew_grid(int64_t n, int threads = 256) {
  int64_t g = dasa_cdiv(n, threads);
  const int64_t cap = (int64_t)DASA_NUM_SMS * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__global__ void __launch_bounds__(256) dropout_apply_kernel(const float* __restrict__ x, int64_t ldx, const uint8_t* __restrict__ mask,
                                                            float scale, float* __restrict__ y, int64_t ldy, int R, int C) {
  const int64_t total = (int64_t)R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    float v = x[(int64_t)r * ldx + c];
    if (mask != nullptr) v *= mask[i] ? scale : 0.f;
    y[(int64_t)r * ldy + c] = v;
  }
}

__global__ void __launch_bounds__(256) act_backward_kernel(int act, const float* __restrict__ dy, int64_t lddy,
                                                           const float* __restrict__ y, int64_t ldy, const uint8_t* __restrict__ mask,
                                                           float scale, float* __restrict__ dx, int64_t lddx, int R, int C) {
  const int64_t total = (int64_t)R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    float g = dy[(int64_t)r * lddy + c];
    if (mask != nullptr) g *= mask[i] ? scale : 0.f;
    const float yv = y[(int64_t)r * ldy + c];
    float o;
    if (act == 0) o = g * (1.f - yv * yv);
    else if (act == 1) o = yv > 0.f ? g : 0.f;
    else o = g * (0.5f * (1.f + erff(yv * 0.70710678118654752f)) + yv * 0.3989422804014327f * expf(-0.5f * yv * yv));   // gelu'(pre-activation)
    dx[(int64_t)r * lddx + c] = o;
  }
}

// y = x * 0.5 * (1 + erf(x / sqrt(2)))  (vilmodel.gelu, vilmodel.py:125-131)
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v * 0.5f * (1.f + erff(v * 0.70710678118654752f));
  }
}

__global__ void __launch_bounds__(256) axpy2d_kernel(float a, const float* __restrict__ x, int64_t ldx, float* __restrict__ y,
                                                     int64_t ldy, int accumulate, int R, int C) {
  const int64_t total = (int64_t)R * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    const float v = a * x[(int64_t)r * ldx + c];
    float* dst = y + (int64_t)r * ldy + c;
    *dst = accumulate ? *dst + v : v;
  }
}

__device__ __forceinline__ void mask_fill(uint8_t* __restrict__ mask, int64_t n, float p, uint64_t raw_seed, uint64_t offset) {
  const uint64_t seed = mix_seed(raw_seed);
  const uint32_t thr = (uint32_t)(p * 65536.0f);                       // keep iff 16-bit uniform >= p
  const int64_t n16 = n >> 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t h = hash64(seed, offset + (uint64_t)(4 * i + q));
      w[q] = ((uint32_t)(h & 0xFFFF) >= thr ? 1u : 0u) | (((uint32_t)((h >> 16) & 0xFFFF) >= thr ? 1u : 0u) << 8) |
             (((uint32_t)((h >> 32) & 0xFFFF) >= thr ? 1u : 0u) << 16) | (((uint32_t)(h >> 48) >= thr ? 1u : 0u) << 24);
    }
    reinterpret_cast<uint4*>(mask)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // tail (n not a multiple of 16)
  for (int64_t i = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = (uint32_t)(hash64(seed ^ 0x5bd1e9955bd1e995ull, offset + (uint64_t)i) & 0xFFFF) >= thr ? 1 : 0;
}

__global__ void __launch_bounds__(256) dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t n, float p, uint64_t seed, uint64_t offset) {
  mask_fill(mask, n, p, seed, offset);
}

constexpr int CE_MAX_ROWS = 1 << 18;
__device__ float g_ce_rows[CE_MAX_ROWS];
__device__ unsigned int g_ce_ticket = 0;

// one warp per sample: log-softmax over the (masked) candidate logits, CE with ignore_index, gradient, argmax
__global__ void __launch_bounds__(128) masked_ce_kernel(const float* __restrict__ logit, int64_t ld, const int64_t* __restrict__ target,
                                                        int ignore_index, int B, int Nc, float grad_scale, float* __restrict__ loss_acc,
                                                        float* __restrict__ dlogit, int64_t* __restrict__ action,
                                                        float* __restrict__ logprob_action, float* __restrict__ entropy) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float* row_loss = g_ce_rows;
  if (b < B) {
  const float* z = logit + (int64_t)b * ld;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = lane; j < Nc; j += 32) {
    const float v = z[j];
    if (v > mx) { mx = v; arg = j; }
  }
  // argmax with first-index tie-break (torch.max semantics)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float sum = 0.f;
  for (int j = lane; j < Nc; j += 32) sum += expf(z[j] - mx);
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  const int64_t t = target ? target[b] : (int64_t)ignore_index;
  const bool valid = target && (t != (int64_t)ignore_index);
  // torch's CrossEntropyLoss raises on a target outside [0, Nc); a kernel inside a captured graph cannot, so the loss is
  // poisoned with NaN instead of silently reading past the row (e.g. a candidate buffer narrower than cand_leng)
  const bool in_range = !valid || (t >= 0 && t < (int64_t)Nc);
  float ent = 0.f;
  for (int j = lane; j < Nc; j += 32) {
    const float lp = z[j] - lse;          // -inf for masked candidates
    const float p = expf(lp);
    if (p > 0.f) ent -= p * lp;
    if (dlogit != nullptr)
      dlogit[(int64_t)b * Nc + j] = !in_range ? NAN : (valid ? (p - (j == (int)t ? 1.f : 0.f)) * grad_scale : 0.f);
  }
  ent = warp_sum(ent);
  if (lane == 0) {
    if (loss_acc != nullptr) row_loss[b] = valid ? (in_range ? lse - z[t] : NAN) : 0.f;
    if (action != nullptr) action[b] = arg;
    if (logprob_action != nullptr) logprob_action[b] = z[arg] - lse;
    if (entropy != nullptr) entropy[b] = ent;
  }
  }
  if (loss_acc == nullptr) return;
  // the per-row losses are folded in row order by the last block to finish (bit-reproducible loss, no float atomics)
  __shared__ bool last;
  __shared__ float red[32];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&g_ce_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  float v = 0.f;
  const int per = (B + blockDim.x - 1) / blockDim.x;          // contiguous slice per thread, fixed order
  for (int i = threadIdx.x * per; i < min(B, (int)(threadIdx.x + 1) * per); ++i) v += __ldcg(row_loss + i);
  v = block_sum(v, red);
  if (threadIdx.x == 0) { loss_acc[0] += v; g_ce_ticket = 0; }
}

// out[c, r] = in[r, c] through a 32x33 shared tile (both sides coalesced)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int64_t ld_in, int rows, int cols,
                                                        float* __restrict__ out, int64_t ld_out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(int64_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(int64_t)c * ld_out + r] = tile[threadIdx.x][i];
  }
}

__global__ void bump_counter_kernel(unsigned long long* ctr, unsigned long long inc) { ctr[0] += inc; }

__global__ void __launch_bounds__(256) dropout_mask_dev_kernel(uint8_t* __restrict__ mask, int64_t n, float p,
                                                               const unsigned long long* __restrict__ seed_dev, uint64_t offset) {
  mask_fill(mask, n, p, seed_dev[0], offset);
}

__global__ void __launch_bounds__(256) rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq,
                                                      int64_t n, float lr_host, float alpha, float eps, float wd,
                                                      const float* __restrict__ clip, const float* __restrict__ lr_scale) {
  const float c = clip ? clip[0] : 1.f;
  const float lr = lr_scale ? lr_host * lr_scale[0] : lr_host;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i] * c;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float s = alpha * sq[i] + (1.f - alpha) * gi * gi;
    sq[i] = s;
    p[i] = pi - lr * gi / (sqrtf(s) + eps);
  }
}

// Two-stage, fixed-order sum of squares (bit-reproducible clip coefficient): every block folds its grid-stride slice, the
// LAST block to finish (ticket counter) adds the block partials in index order and accumulates into out[0].
constexpr int SUMSQ_BLOCKS = DASA_NUM_SMS * 4;
__device__ float g_sumsq_part[SUMSQ_BLOCKS];
__device__ unsigned int g_sumsq_ticket = 0;

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  __shared__ float red[32];
  __shared__ bool last;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fmaf(x[i], x[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    g_sumsq_part[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(&g_sumsq_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float v = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(g_sumsq_part + i);   // fixed slice per thread
  v = block_sum(v, red);
  if (threadIdx.x == 0) { out[0] += v; g_sumsq_ticket = 0; }
}

__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef) {
  const float norm = sqrtf(sumsq[0]);
  coef[0] = fminf(1.f, max_norm / (norm + 1e-6f));
}

}  // namespace

extern "C" int dasa_dropout_apply(const float* x, int64_t ldx, const uint8_t* mask, float scale, float* y, int64_t ldy, int R, int C,
                                  void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  dropout_apply_kernel<<<ew_grid((int64_t)R * C), 256, 0, (cudaStream_t)stream>>>(x, ldx, mask, scale, y, ldy, R, C);
  return dasa_check_launch("dropout_apply_kernel");
}

extern "C" int dasa_act_backward(int act, const float* dy, int64_t lddy, const float* y, int64_t ldy, const uint8_t* mask, float scale,
                                 float* dx, int64_t lddx, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  if (act < 0 || act > 2) return DASA_ERR_BAD_SHAPE;
  act_backward_kernel<<<ew_grid((int64_t)R * C), 256, 0, (cudaStream_t)stream>>>(act, dy, lddy, y, ldy, mask, scale, dx, lddx, R, C);
  return dasa_check_launch("act_backward_kernel");
}

extern "C" int dasa_gelu_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return DASA_OK;
  gelu_fwd_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return dasa_check_launch("gelu_fwd_kernel");
}

extern "C" int dasa_axpy2d(float a, const float* x, int64_t ldx, float* y, int64_t ldy, int accumulate, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  axpy2d_kernel<<<ew_grid((int64_t)R * C), 256, 0, (cudaStream_t)stream>>>(a, x, ldx, y, ldy, accumulate, R, C);
  return dasa_check_launch("axpy2d_kernel");
}

extern "C" int dasa_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  if (n <= 0) return DASA_OK;
  if (reinterpret_cast<uintptr_t>(mask) & 15) return DASA_ERR_BAD_ALIGN;
  dropout_mask_kernel<<<ew_grid(n / 16 + 1), 256, 0, (cudaStream_t)stream>>>(mask, n, p, seed, offset);
  return dasa_check_launch("dropout_mask_kernel");
}

extern "C" int dasa_transpose(const float* in, int64_t ld_in, int rows, int cols, float* out, int64_t ld_out, void* stream) {
  if (rows <= 0 || cols <= 0) return DASA_OK;
  dim3 grid((unsigned)dasa_cdiv(cols, 32), (unsigned)dasa_cdiv(rows, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(in, ld_in, rows, cols, out, ld_out);
  return dasa_check_launch("transpose_kernel");
}

extern "C" int dasa_dropout_mask_dev(uint8_t* mask, int64_t n, float p, const uint64_t* seed_dev, uint64_t offset, void* stream) {
  if (n <= 0) return DASA_OK;
  if (reinterpret_cast<uintptr_t>(mask) & 15) return DASA_ERR_BAD_ALIGN;
  dropout_mask_dev_kernel<<<ew_grid(n / 16 + 1), 256, 0, (cudaStream_t)stream>>>(mask, n, p, reinterpret_cast<const unsigned long long*>(seed_dev), offset);
  return dasa_check_launch("dropout_mask_dev_kernel");
}

extern "C" int dasa_bump_counter(uint64_t* counter, uint64_t inc, void* stream) {
  bump_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(counter), inc);
  return dasa_check_launch("bump_counter_kernel");
}

// fp32 -> fp16 (round to nearest even), 8 elements per thread
__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n) {
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
    reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<uint4*>(h);
  }
  for (int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}

// the same over the leading C columns of R strided rows (C % 8 == 0): out[r, c] = fp16(in[r * ld_in + c])
__global__ void __launch_bounds__(256) f32_to_f16_rows_kernel(const float* __restrict__ in, int64_t ld_in, __half* __restrict__ out,
                                                              int64_t ld_out, int R, int C) {
  const int c8 = C >> 3;
  const int64_t total = (int64_t)R * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c8), c = (int)(i % c8) << 3;
    const float4 a = ldg_stream4(in + (int64_t)r * ld_in + c), b = ldg_stream4(in + (int64_t)r * ld_in + c + 4);
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(out + (int64_t)r * ld_out + c) = *reinterpret_cast<uint4*>(h);
  }
}

extern "C" int dasa_f32_to_f16_rows(const float* in, int64_t ld_in, dasa_half_t* out, int64_t ld_out, int R, int C, void* stream) {
  if (R <= 0 || C <= 0) return DASA_OK;
  if ((C & 7) || (ld_in & 3) || (ld_out & 7)) return DASA_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return DASA_ERR_BAD_ALIGN;
  f32_to_f16_rows_kernel<<<ew_grid((int64_t)R * (C >> 3)), 256, 0, (cudaStream_t)stream>>>(in, ld_in, reinterpret_cast<__half*>(out), ld_out, R, C);
  return dasa_check_launch("f32_to_f16_rows_kernel");
}

extern "C" int dasa_f32_to_f16(const float* in, dasa_half_t* out, int64_t n, void* stream) {
  if (n <= 0) return DASA_OK;
  if (n >= 8 && (!dasa_aligned16(in) || !dasa_aligned16(out))) return DASA_ERR_BAD_ALIGN;
  f32_to_f16_kernel<<<ew_grid(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<__half*>(out), n);
  return dasa_check_launch("f32_to_f16_kernel");
}

extern "C" int dasa_masked_ce(const float* logit, int64_t ld, const int64_t* target, int ignore_index, int B, int Nc, float grad_scale,
                              float* loss_acc, float* dlogit, int64_t* action, float* logprob_action, float* entropy, void* stream) {
  if (B <= 0) return DASA_OK;
  if (Nc <= 0 || B > CE_MAX_ROWS) return DASA_ERR_BAD_SHAPE;
  masked_ce_kernel<<<(unsigned)dasa_cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(logit, ld, target, ignore_index, B, Nc, grad_scale,
                                                                               loss_acc, dlogit, action, logprob_action, entropy);
  return dasa_check_launch("masked_ce_kernel");
}

extern "C" int dasa_rmsprop_step(float* param, const float* grad, float* square_avg, int64_t n, float lr, float alpha, float eps,
                                 float weight_decay, const float* clip_coef, const float* lr_scale, void* stream) {
  if (n <= 0) return DASA_OK;
  rmsprop_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(param, grad, square_avg, n, lr, alpha, eps, weight_decay, clip_coef,
                                                               lr_scale);
  return dasa_check_launch("rmsprop_kernel");
}

// LambdaLR multiplier of agent_dg.py:219-227 from a DEVICE iteration counter (a captured CUDA graph replays the schedule):
// mult[0] = lr_lambda(*iter); then *iter += advance
__global__ void lr_lambda_kernel(int* iter, int warm, int decay_start, int decay_intervals, float lr_decay, float* mult, int advance) {
  const int it = *iter;
  double a = 1.0;
  if (warm > 0 && it < warm) a = (1.0 + (double)it) / (double)warm;
  else if (it >= decay_start) {
    const int n = (it - decay_start) / (decay_intervals > 0 ? decay_intervals : 1);
    for (int i = 0; i < n && a > 1e-30; ++i) a *= (double)lr_decay;
  }
  mult[0] = (float)a;
  *iter = it + advance;
}

extern "C" int dasa_lr_lambda(int* iter, int warm_steps, int decay_start, int decay_intervals, float lr_decay, float* mult, int advance,
                              void* stream) {
  lr_lambda_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(iter, warm_steps, decay_start, decay_intervals, lr_decay, mult, advance);
  return dasa_check_launch("lr_lambda_kernel");
}

extern "C" int dasa_sumsq(const float* x, int64_t n, float* out, void* stream) {
  if (n <= 0) return DASA_OK;
  // launches on one stream are ordered, which is what the shared partial buffer / ticket rely on (one optimizer per process)
  int64_t g = dasa_cdiv(n, 256 * 8);
  g = g < 1 ? 1 : (g > SUMSQ_BLOCKS ? SUMSQ_BLOCKS : g);
  sumsq_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return dasa_check_launch("sumsq_kernel");
}

extern "C" int dasa_clip_coef(const float* sumsq, float max_norm, float* clip_coef, void* stream) {
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, clip_coef);
  return dasa_check_launch("clip_coef_kernel");
}

// Sampled-feedback / A2C pieces of vl_rollout (agent_dg.py:870-999): categorical action sampling with log-prob and
// entropy (and their backward), the per-step reward / mask / ended bookkeeping, and the fused A2C epilogue (discounted
// returns, advantage, policy + value + entropy losses and all their gradients in one launch).
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------- action sampling
// warp per episode. probs = softmax(logit) (candidates past cand_leng carry -inf => p = 0), entropy = -sum p log p,
// action = injected | inverse-CDF sample with the supplied uniform | argmax; logprob = log p[action].
__global__ void __launch_bounds__(256) policy_sample_fwd_kernel(const float* __restrict__ logit, int64_t ld, int B, int Nc,
                                                                const float* __restrict__ u, const int64_t* __restrict__ action_in,
                                                                int64_t* __restrict__ action, float* __restrict__ logprob,
                                                                float* __restrict__ entropy, float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* z = logit + (int64_t)b * ld;
  float mx = -INFINITY;
  for (int j = lane; j < Nc; j += 32) mx = fmaxf(mx, z[j]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < Nc; j += 32) s += (z[j] == -INFINITY) ? 0.f : expf(z[j] - mx);
  s = warp_sum(s);
  const float logs = logf(s), inv = 1.f / s;
  float h = 0.f;
  for (int j = lane; j < Nc; j += 32) {
    const bool live = z[j] != -INFINITY;
    const float p = live ? expf(z[j] - mx) * inv : 0.f;
    if (probs) probs[(int64_t)b * Nc + j] = p;
    if (live && p > 0.f) h -= p * (z[j] - mx - logs);
  }
  h = warp_sum(h);
  int a;
  if (action_in != nullptr) {
    a = (int)action_in[b];
  } else if (u != nullptr) {
    // smallest j with cumsum(p)[j] > u; the last live candidate catches rounding
    const float target = u[b];
    float run = 0.f;
    a = -1;
    int last_live = 0;
    for (int base = 0; base < Nc && a < 0; base += 32) {
      const int j = base + lane;
      const float p = (j < Nc && z[j] != -INFINITY) ? expf(z[j] - mx) * inv : 0.f;
      float c = p;                                   // inclusive scan over the warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) c += v;
      }
      c += run;
      const unsigned hit = __ballot_sync(0xffffffffu, p > 0.f && c > target);
      const unsigned live = __ballot_sync(0xffffffffu, p > 0.f);
      if (live) last_live = base + 31 - __clz(live);
      if (hit) a = base + __ffs(hit) - 1;
      run = __shfl_sync(0xffffffffu, c, 31);
    }
    if (a < 0) a = last_live;
  } else {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = lane; j < Nc; j += 32)
      if (z[j] > best) { best = z[j]; bi = j; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    a = (bi == 0x7fffffff) ? 0 : bi;
  }
  if (lane == 0) {
    if (action) action[b] = a;
    if (entropy) entropy[b] = h;
    if (logprob) logprob[b] = (a >= 0 && a < Nc) ? (z[a] - mx - logs) : 0.f;
  }
}

// dlogit_j = dlogp (delta_aj - p_j) - dent p_j (log p_j + H)
__global__ void __launch_bounds__(256) policy_sample_bwd_kernel(const float* __restrict__ probs, const int64_t* __restrict__ action,
                                                                const float* __restrict__ dlogp, const float* __restrict__ dent,
                                                                const float* __restrict__ entropy, int B, int Nc,
                                                                float* __restrict__ dlogit, int64_t ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Nc) return;
  const int b = i / Nc, j = i % Nc;
  const float p = probs[i];
  float g = 0.f;
  if (dlogp) g = dlogp[b] * (((int64_t)j == action[b] ? 1.f : 0.f) - p);
  if (dent && p > 0.f) g -= dent[b] * p * (logf(p) + entropy[b]);
  dlogit[(int64_t)b * ld + j] = g;
}

// ---------------------------------------------------------------------------------------- reward / mask / ended
// agent_dg.py:890-930 with `dist` = distance to the goal AFTER the action and `last_dist` before it.
__global__ void nav_reward_kernel(const int64_t* __restrict__ action, const int32_t* __restrict__ cand_leng, int ignore_id,
                                  const float* __restrict__ dist, const float* __restrict__ last_dist, uint8_t* __restrict__ ended,
                                  float* __restrict__ reward, float* __restrict__ mask, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t a = action[b];
  const bool is_end = (a == (int64_t)cand_leng[b] - 1) || (a == (int64_t)ignore_id);
  float r = 0.f, m = 1.f;
  if (ended[b]) {
    m = 0.f;
  } else if (is_end) {
    r = (dist[b] < 3.f) ? 2.f : -2.f;
  } else {
    const float d = -(dist[b] - last_dist[b]);
    r = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);     // the reference raises on d == 0 (a move that changes nothing)
  }
  reward[b] = r;
  mask[b] = m;
  ended[b] = (uint8_t)(ended[b] || is_end);
}

// ----------------------------------------------------------------------------------------------- A2C epilogue
// One block. Thread per episode walks t = T-1..0 (agent_dg.py:959-992): R = R*gamma + r_t (two roundings, like numpy),
// a = R - v_t; loss += -logp*a*m + 0.5*a^2*m - ent_coef*ent*m; total += m. Then the normaliser (total | batch | none) and
// the gradients dlogp = -a*m/norm, dvalue = -a*m/norm, dent = -ent_coef*m/norm.
__global__ void __launch_bounds__(1024) a2c_loss_kernel(const float* __restrict__ logp, const float* __restrict__ ent,
                                                         const float* __restrict__ value, const float* __restrict__ last_value,
                                                         const float* __restrict__ reward, const float* __restrict__ mask,
                                                         const uint8_t* __restrict__ ended, float gamma, float ent_coef,
                                                         int normalize, int T, int B, float* __restrict__ loss_out,
                                                         float* __restrict__ total_out, float* __restrict__ dlogp,
                                                         float* __restrict__ dent, float* __restrict__ dvalue) {
  __shared__ float red[64];
  float loss = 0.f, total = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float R = ended[b] ? 0.f : last_value[b];
    for (int t = T - 1; t >= 0; --t) {
      const int i = t * B + b;
      R = __fadd_rn(__fmul_rn(R, gamma), reward[i]);
      const float m = mask[i];
      const float a = R - value[i];
      loss += -logp[i] * a * m + 0.5f * a * a * m;
      if (ent) loss += -ent_coef * ent[i] * m;
      total += m;
      dvalue[i] = a;                                  // advantage, scaled below
    }
  }
  loss = block_sum(loss, red);
  total = block_sum(total, red + 32);
  const float norm = (normalize == 1) ? total : ((normalize == 2) ? (float)B : 1.f);
  const float inv = (norm != 0.f) ? 1.f / norm : 0.f;
  if (threadIdx.x == 0) {
    loss_out[0] = loss * inv;
    if (total_out) total_out[0] = total;
  }
  for (int i = threadIdx.x; i < T * B; i += blockDim.x) {
    const float m = mask[i] * inv;
    const float a = dvalue[i];
    dvalue[i] = -a * m;
    dlogp[i] = -a * m;
    if (dent) dent[i] = -ent_coef * m;
  }
}

}  // namespace

extern "C" int dasa_policy_sample_fwd(const float* logit, int64_t ld, int B, int Nc, const float* u, const int64_t* action_in,
                                      int64_t* action, float* logprob, float* entropy, float* probs, void* stream) {
  if (B <= 0) return DASA_OK;
  if (Nc <= 0 || ld < Nc) return DASA_ERR_BAD_SHAPE;
  policy_sample_fwd_kernel<<<(unsigned)dasa_cdiv(B, 8), 256, 0, (cudaStream_t)stream>>>(logit, ld, B, Nc, u, action_in, action,
                                                                                       logprob, entropy, probs);
  return dasa_check_launch("policy_sample_fwd_kernel");
}

extern "C" int dasa_policy_sample_bwd(const float* probs, const int64_t* action, const float* dlogp, const float* dent,
                                      const float* entropy, int B, int Nc, float* dlogit, int64_t ld, void* stream) {
  if (B <= 0) return DASA_OK;
  if (Nc <= 0 || ld < Nc || (dent != nullptr && entropy == nullptr)) return DASA_ERR_BAD_SHAPE;
  policy_sample_bwd_kernel<<<(unsigned)dasa_cdiv((int64_t)B * Nc, 256), 256, 0, (cudaStream_t)stream>>>(probs, action, dlogp, dent,
                                                                                                      entropy, B, Nc, dlogit, ld);
  return dasa_check_launch("policy_sample_bwd_kernel");
}

// Speaker greedy decode, word selection of one step (speaker.py:318-343): logits[:, unk] = -inf; word = argmax (first index
// on ties, like torch.max); the emitted word is <PAD> for sequences that had already ended; ended |= (emitted == <EOS>).
// One warp per sequence. next_word feeds the next decoder step (the raw argmax, as the reference feeds `word`).
__global__ void __launch_bounds__(128) speaker_select_kernel(const float* __restrict__ logit, int64_t ld, int B, int V, int unk,
                                                             int pad, int eos, uint8_t* __restrict__ ended,
                                                             int64_t* __restrict__ next_word, int64_t* __restrict__ emitted,
                                                             int64_t ld_emitted) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* z = logit + (int64_t)b * ld;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int j = lane; j < V; j += 32) {
    const float v = (j == unk) ? -INFINITY : z[j];
    if (v > mx || (v == mx && j < arg)) { mx = v; arg = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  if (lane == 0) {
    const bool was_ended = ended[b] != 0;
    const int w = was_ended ? pad : arg;
    next_word[b] = arg;
    emitted[(int64_t)b * ld_emitted] = w;
    ended[b] = (uint8_t)(was_ended || w == eos);
  }
}

extern "C" int dasa_speaker_select(const float* logit, int64_t ld, int B, int V, int unk, int pad, int eos, uint8_t* ended,
                                   int64_t* next_word, int64_t* emitted, int64_t ld_emitted, void* stream) {
  if (B <= 0) return DASA_OK;
  if (V <= 0) return DASA_ERR_BAD_SHAPE;
  speaker_select_kernel<<<(unsigned)dasa_cdiv(B, 4), 128, 0, (cudaStream_t)stream>>>(logit, ld, B, V, unk, pad, eos, ended, next_word,
                                                                                    emitted, ld_emitted);
  return dasa_check_launch("speaker_select_kernel");
}

extern "C" int dasa_nav_reward(const int64_t* action, const int32_t* cand_leng, int ignore_id, const float* dist,
                               const float* last_dist, uint8_t* ended, float* reward, float* mask, int B, void* stream) {
  if (B <= 0) return DASA_OK;
  nav_reward_kernel<<<(unsigned)dasa_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(action, cand_leng, ignore_id, dist, last_dist,
                                                                                  ended, reward, mask, B);
  return dasa_check_launch("nav_reward_kernel");
}

extern "C" int dasa_a2c_loss(const float* logp, const float* ent, const float* value, const float* last_value, const float* reward,
                             const float* mask, const uint8_t* ended, float gamma, float ent_coef, int normalize, int T, int B,
                             float* loss, float* total, float* dlogp, float* dent, float* dvalue, void* stream) {
  if (T <= 0 || B <= 0) return DASA_ERR_BAD_SHAPE;
  if (normalize < 0 || normalize > 2 || (ent != nullptr && dent == nullptr)) return DASA_ERR_BAD_SHAPE;
  a2c_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logp, ent, value, last_value, reward, mask, ended, gamma, ent_coef,
                                                        normalize, T, B, loss, total, dlogp, dent, dvalue);
  return dasa_check_launch("a2c_loss_kernel");
}

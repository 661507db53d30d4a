// Device-resident navigation environment: the per-step observation assembly and the state
// transition of the agent_dg rollout as two kernels over tables that live in HBM (dasa_b200/navgraph.py).
//   env_observe : env.py:_get_obs / make_candidate (buffered branch, :299-311) + agent_dg.py:get_input_feat (:313-323),
//                 _candidate_variable (:300-311), _teacher_action (:325-344) -> f_t, d_t, cand_feat, cand_dfeat, input_a_t,
//                 cand_leng, target, dist, written straight into the rollout's [T, B, ...] buffers (no host round trip).
//   env_step    : agent_dg.py:890-935 — the <end>/ignore test, make_equiv_action (:358-391: turn to the candidate's pointId,
//                 move to it), the distance after the move, the quantised reward, the mask and the ended flags.
// Both are HBM-bound gathers: one CTA per output row copies a 2048-float bank row with 128-bit streaming accesses and
// appends the 128-float angle part from the precomputed float32 tables (bit-identical to utils.angle_feature).
#include "common.cuh"

namespace {

struct ObserveArgs {
  const float *rgb_bank, *dep_bank;      // [n_vp, V, C]
  const int32_t *nbr, *nbr_point, *deg;  // [n_vp, dmax], [n_vp, dmax], [n_vp]
  const float *cand_angle;               // [n_vp, dmax, 12, 4]
  const float *view_angle;               // [12, V, 4]
  const float *agent_angle;              // [V, 4]
  const float *dist_tab;                 // [n_vp, n_vp]
  const int32_t *next_hop;               // [n_vp, n_vp]
  const int32_t *vp, *view, *goal;       // episode state [B]
  const uint8_t *ended;                  // [B] or nullptr
  float *f_t, *d_t, *cand, *cand_d;      // [B, V, F], [B, V, F], [B, nc, F] x2
  float *input_a_t;                      // [B, A]
  int32_t *cand_leng;                    // [B]
  int64_t *target;                       // [B]
  float *dist;                           // [B]
  int64_t ld_f_sample, ld_c_sample;
  int n_vp, dmax, B, V, C, A, nc, ignore_id, headings;
};

// rows [0, V): panorama views; rows [V, V + nc): candidate slots (slot deg = END row, zeros; beyond: zero padding)
__global__ void __launch_bounds__(256) env_observe_kernel(const ObserveArgs a) {
  const int rows = a.V + a.nc;
  const int F = a.C + a.A;
  for (int64_t job = blockIdx.x; job < (int64_t)a.B * rows; job += gridDim.x) {
    const int b = (int)(job / rows), r = (int)(job % rows);
    const int vp = a.vp[b], view = a.view[b];
    const int hb = view % a.headings;
    const float *src_f = nullptr, *src_d = nullptr, *ang = nullptr;
    float *dst_f, *dst_d;
    if (r < a.V) {
      src_f = a.rgb_bank + ((int64_t)vp * a.V + r) * a.C;
      src_d = a.dep_bank + ((int64_t)vp * a.V + r) * a.C;
      ang = a.view_angle + ((int64_t)hb * a.V + r) * 4;
      dst_f = a.f_t + (int64_t)b * a.ld_f_sample + (int64_t)r * F;
      dst_d = a.d_t + (int64_t)b * a.ld_f_sample + (int64_t)r * F;
    } else {
      const int k = r - a.V;
      dst_f = a.cand + (int64_t)b * a.ld_c_sample + (int64_t)k * F;
      dst_d = a.cand_d + (int64_t)b * a.ld_c_sample + (int64_t)k * F;
      if (k < a.dmax && k < a.deg[vp]) {
        const int pt = a.nbr_point[(int64_t)vp * a.dmax + k];
        src_f = a.rgb_bank + ((int64_t)vp * a.V + pt) * a.C;
        src_d = a.dep_bank + ((int64_t)vp * a.V + pt) * a.C;
        ang = a.cand_angle + (((int64_t)vp * a.dmax + k) * 12 + hb) * 4;
      }
    }
    const int c4 = a.C >> 2, f4 = F >> 2;
    if (src_f != nullptr) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(ang));
      for (int i = threadIdx.x; i < f4; i += blockDim.x) {
        float4 x, y;
        if (i < c4) {
          x = ldg_stream4(src_f + 4 * i);
          y = ldg_stream4(src_d + 4 * i);
        } else {
          x = av;
          y = av;
        }
        stg_stream4(dst_f + 4 * i, x);
        stg_stream4(dst_d + 4 * i, y);
      }
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int i = threadIdx.x; i < f4; i += blockDim.x) {
        stg_stream4(dst_f + 4 * i, z);
        stg_stream4(dst_d + 4 * i, z);
      }
    }
    if (r == 0) {   // per-episode scalars
      const int g = a.goal[b], dg = a.deg[vp];
      for (int i = threadIdx.x; i < a.A; i += blockDim.x) a.input_a_t[(int64_t)b * a.A + i] = a.agent_angle[view * 4 + (i & 3)];
      if (threadIdx.x == 0) {
        a.cand_leng[b] = dg + 1;
        const int nh = a.next_hop[(int64_t)vp * a.n_vp + g];
        int64_t tgt = (nh < 0) ? (int64_t)dg : (int64_t)nh;      // at the goal: STOP = the END row (agent_dg.py:341-343)
        if (a.ended != nullptr && a.ended[b]) tgt = a.ignore_id;
        if (a.target != nullptr) a.target[b] = tgt;
        if (a.dist != nullptr) a.dist[b] = a.dist_tab[(int64_t)vp * a.n_vp + g];
      }
    }
  }
}

__global__ void env_step_kernel(const int64_t* __restrict__ action, int ignore_id, const int32_t* __restrict__ nbr,
                                const int32_t* __restrict__ nbr_point, const int32_t* __restrict__ deg, int dmax,
                                const float* __restrict__ dist_tab, int n_vp, int32_t* __restrict__ vp, int32_t* __restrict__ view,
                                const int32_t* __restrict__ goal, uint8_t* __restrict__ ended, float* __restrict__ last_dist,
                                float* __restrict__ reward, float* __restrict__ mask, int32_t* __restrict__ traj_vp,
                                int32_t* __restrict__ traj_view, int32_t* __restrict__ err, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t act = action[b];
  int v = vp[b], w = view[b];
  const int dg = deg[v];
  const bool is_end = (act == (int64_t)dg) || (act == (int64_t)ignore_id);   // cand_leng - 1 == deg (agent_dg.py:893)
  if (!is_end) {
    if (act < 0 || act >= dg) {
      atomicOr(err, 1);                      // an action outside the candidate list (masked logits make this impossible)
    } else {
      w = nbr_point[(int64_t)v * dmax + act];                                // turn to the candidate's view ...
      v = nbr[(int64_t)v * dmax + act];                                      // ... and move (make_equiv_action)
    }
  }
  const float d = dist_tab[(int64_t)v * n_vp + goal[b]];
  float r = 0.f, m = 1.f;
  if (ended[b]) {
    m = 0.f;
  } else if (is_end) {
    r = (d < 3.f) ? 2.f : -2.f;
  } else {
    const float delta = -(d - last_dist[b]);
    r = (delta > 0.f) ? 1.f : ((delta < 0.f) ? -1.f : 0.f);
    if (delta == 0.f) atomicOr(err, 2);      // the reference raises NameError("The action doesn't change the move")
  }
  if (reward != nullptr) reward[b] = r;
  if (mask != nullptr) mask[b] = m;
  ended[b] = (uint8_t)(ended[b] || is_end);
  last_dist[b] = d;
  vp[b] = v;
  view[b] = w;
  if (traj_vp != nullptr) traj_vp[b] = v;
  if (traj_view != nullptr) traj_view[b] = w;
}

// --submit (agent_dg.py:834-840, "avoiding cyclic path"): the current viewpoint joins the episode's visited set (a bitmap over
// the viewpoints), every candidate whose viewpoint is in the set is masked: blocked[b, k] = 1 and logit[b, k] = -inf. The END
// slot (k == deg) is not an entry of ob['candidate'] and is never masked. One thread per episode (B x dmax table reads).
__global__ void env_visited_mask_kernel(const int32_t* __restrict__ vp, const int32_t* __restrict__ nbr,
                                        const int32_t* __restrict__ deg, int dmax, uint32_t* __restrict__ visited, int words,
                                        uint8_t* __restrict__ blocked, int nc, float* __restrict__ logit, int64_t ld_logit, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int v = vp[b];
  uint32_t* vis = visited + (int64_t)b * words;
  vis[v >> 5] |= 1u << (v & 31);
  const int dg = deg[v];
  for (int k = 0; k < nc; ++k) {
    bool hit = false;
    if (k < dg && k < dmax) {
      const int u = nbr[(int64_t)v * dmax + k];
      hit = (vis[u >> 5] >> (u & 31)) & 1u;
    }
    if (blocked != nullptr) blocked[(int64_t)b * nc + k] = hit ? 1 : 0;
    if (hit && logit != nullptr) logit[(int64_t)b * ld_logit + k] = -INFINITY;
  }
}

}  // namespace

extern "C" int dasa_env_visited_mask(const int32_t* vp, const int32_t* nbr, const int32_t* deg, int dmax, int n_vp,
                                     uint32_t* visited, int words, uint8_t* blocked, int nc, float* logit, int64_t ld_logit,
                                     int B, void* stream) {
  if (B <= 0) return DASA_OK;
  if (n_vp <= 0 || dmax <= 0 || nc <= 0 || visited == nullptr || words * 32 < n_vp || (logit != nullptr && ld_logit < nc))
    return DASA_ERR_BAD_SHAPE;
  env_visited_mask_kernel<<<(unsigned)dasa_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(vp, nbr, deg, dmax, visited, words, blocked,
                                                                                       nc, logit, ld_logit, B);
  return dasa_check_launch("env_visited_mask_kernel");
}

extern "C" int dasa_env_observe(const float* rgb_bank, const float* dep_bank, const int32_t* nbr, const int32_t* nbr_point,
                                const int32_t* deg, const float* cand_angle, const float* view_angle, const float* agent_angle,
                                const float* dist_tab, const int32_t* next_hop, int n_vp, int dmax, const int32_t* vp,
                                const int32_t* view, const int32_t* goal, const uint8_t* ended, int B, int V, int C, int A, int nc,
                                int headings, int ignore_id, float* f_t, float* d_t, int64_t ld_f_sample, float* cand,
                                float* cand_d, int64_t ld_c_sample, float* input_a_t, int32_t* cand_leng, int64_t* target,
                                float* dist, void* stream) {
  if (B <= 0) return DASA_OK;
  if (C % 4 != 0 || A % 4 != 0 || V <= 0 || nc <= 0 || n_vp <= 0 || dmax <= 0 || headings <= 0 || headings > 12 || V % headings != 0)
    return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(rgb_bank) || !dasa_aligned16(dep_bank) || !dasa_aligned16(f_t) || !dasa_aligned16(d_t) ||
      !dasa_aligned16(cand) || !dasa_aligned16(cand_d) || !dasa_aligned16(cand_angle) || !dasa_aligned16(view_angle) ||
      ld_f_sample % 4 != 0 || ld_c_sample % 4 != 0)
    return DASA_ERR_BAD_ALIGN;
  ObserveArgs a;
  a.rgb_bank = rgb_bank; a.dep_bank = dep_bank; a.nbr = nbr; a.nbr_point = nbr_point; a.deg = deg;
  a.cand_angle = cand_angle; a.view_angle = view_angle; a.agent_angle = agent_angle; a.dist_tab = dist_tab; a.next_hop = next_hop;
  a.vp = vp; a.view = view; a.goal = goal; a.ended = ended; a.f_t = f_t; a.d_t = d_t; a.cand = cand; a.cand_d = cand_d;
  a.input_a_t = input_a_t; a.cand_leng = cand_leng; a.target = target; a.dist = dist;
  a.ld_f_sample = ld_f_sample; a.ld_c_sample = ld_c_sample;
  a.n_vp = n_vp; a.dmax = dmax; a.B = B; a.V = V; a.C = C; a.A = A; a.nc = nc; a.ignore_id = ignore_id; a.headings = headings;
  const int64_t jobs = (int64_t)B * (V + nc);
  const unsigned grid = (unsigned)(jobs < (int64_t)DASA_NUM_SMS * 32 ? jobs : (int64_t)DASA_NUM_SMS * 32);
  env_observe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return dasa_check_launch("env_observe_kernel");
}

extern "C" int dasa_env_step(const int64_t* action, int ignore_id, const int32_t* nbr, const int32_t* nbr_point, const int32_t* deg,
                             int dmax, const float* dist_tab, int n_vp, int32_t* vp, int32_t* view, const int32_t* goal,
                             uint8_t* ended, float* last_dist, float* reward, float* mask, int32_t* traj_vp, int32_t* traj_view,
                             int32_t* err, int B, void* stream) {
  if (B <= 0) return DASA_OK;
  if (n_vp <= 0 || dmax <= 0 || err == nullptr) return DASA_ERR_BAD_SHAPE;
  env_step_kernel<<<(unsigned)dasa_cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(action, ignore_id, nbr, nbr_point, deg, dmax, dist_tab,
                                                                               n_vp, vp, view, goal, ended, last_dist, reward, mask,
                                                                               traj_vp, traj_view, err, B);
  return dasa_check_launch("env_step_kernel");
}

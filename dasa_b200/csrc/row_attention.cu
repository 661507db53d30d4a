// Decoder attention kernels (model.py:253-353): ShiftSoftDotAttention over the 36-view panorama, SoftDotAttention over
// the instruction context, candidate logits. HBM-bound: the context of one sample is staged ONCE in shared memory by
// bulk-async (TMA engine) row copies, split by channel across the CTAs of a thread-block cluster; the per-row partial
// dot products are exchanged through distributed shared memory, every CTA then owns the full 36/80-entry distribution
// and produces its channel slice of the weighted sum from the tile it already holds. One read of ctx, no re-read.
#include <cooperative_groups.h>
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int RA_THREADS = 256;
constexpr int RA_MAX_ROWS = 128;
constexpr int RA_MAX_K = 15;

struct RowAttnSmem {
  float* tile;   // [rows][chunk]
  float* tv;     // [chunk]   target slice
  float* dv;     // [chunk]   (bwd) dwc slice
  float* zpart;  // [CS][rows] partial row dot products from every CTA of the cluster
  float* w;      // [rows] final weights (shifted q or alpha)
  float* p;      // [rows] softmax
  float* aux;    // [rows] scratch (dz in bwd)
  float* red;    // [1024] cross-row-group reduction scratch
  uint64_t* bar;
};

__host__ __device__ inline int ra_pad4(int rows) { return (rows + 3) & ~3; }

__host__ __device__ inline size_t ra_smem_bytes(int rows, int chunk, int cs) {
  const size_t rp = (size_t)ra_pad4(rows);   // keep every sub-array 16-byte aligned (float4 access to red / tv / dv)
  return sizeof(float) * ((size_t)rows * chunk + 2 * (size_t)chunk + (size_t)cs * rp + 3 * rp + 1024) + 16 + 16;
}

__device__ inline RowAttnSmem ra_carve(unsigned char* raw, int rows, int chunk, int cs) {
  RowAttnSmem s;
  s.tile = reinterpret_cast<float*>(raw);
  s.tv = s.tile + (size_t)rows * chunk;
  s.dv = s.tv + chunk;
  s.zpart = s.dv + chunk;
  const int rp = ra_pad4(rows);
  s.w = s.zpart + (size_t)cs * rp;
  s.p = s.w + rp;
  s.aux = s.p + rp;
  s.red = s.aux + rp;
  uintptr_t b = reinterpret_cast<uintptr_t>(s.red + 1024);
  b = (b + 15) & ~uintptr_t(15);
  s.bar = reinterpret_cast<uint64_t*>(b);
  return s;
}

struct RowAttnArgs {
  const float* ctx; int64_t ld_row, ld_sample; int B, rows, D;
  const float* t; int64_t ld_t; const uint8_t* mask; int64_t ld_mask;
  int shift_k, headings; const float* kappa_logits; int64_t ld_kappa;
  float* wc; int64_t ld_wc; float* attn_out; float* q_out; float* kappa_out;
  int chunk;
  int nbox, boxw;        // the channel slice is staged as nbox TMA boxes of [rows x boxw] floats (boxw <= 256, chunk = nbox*boxw)
  // fused DGAdaChannel gate (K1 epilogue -> K3, agent_dg.py:1544-1547 feeding model.py:327-345): the staged tile holds the RAW
  // features f; channels c < gate_C are multiplied in place by sigmoid(gate[b, r, c]) (* chan_scale[c]: the shared drop_env noise,
  // agent_dg.py:656) before the attention runs on it, so the modulated features df_t are never written to HBM
  const float* gate; int64_t ld_grow, ld_gsample; int gate_C; const float* chan_scale;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// stage the rows of this CTA's channel slice: warp 0 issues one bulk copy per unmasked row
__device__ __forceinline__ void ra_issue_loads(const RowAttnSmem& s, const float* ctx_b, int64_t ld_row, int rows, int chunk,
                                               int c0, int cn, const uint8_t* mask_b) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int mine = 0;
    for (int r = lane; r < rows; r += 32) mine += (mask_b == nullptr || mask_b[r] == 0) ? 1 : 0;
    const int nvalid = (int)__reduce_add_sync(0xffffffffu, (unsigned)mine);
    if (lane == 0) mbar_expect_tx(s.bar, (uint32_t)nvalid * (uint32_t)cn * 4u);
    __syncwarp();
    for (int r = lane; r < rows; r += 32)
      if (cn > 0 && (mask_b == nullptr || mask_b[r] == 0))
        bulk_g2s(s.tile + (size_t)r * chunk, ctx_b + (int64_t)r * ld_row + c0, (uint32_t)cn * 4u, s.bar);
  }
}

// z_partial[r] = tile[r,:] . vec  -> pushed into every cluster CTA's zpart[rank][r]
__device__ __forceinline__ void ra_partial_dots(cg::cluster_group& cluster, const RowAttnSmem& s, const float* vec, int rows,
                                                int chunk, int cn, const uint8_t* mask_b, int cs, int rank) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = RA_THREADS / 32;
  const int n4 = cn >> 2;
  for (int r = wid; r < rows; r += nw) {
    float acc = 0.f;
    if (mask_b == nullptr || mask_b[r] == 0) {
      const float4* row = reinterpret_cast<const float4*>(s.tile + (size_t)r * chunk);
      const float4* v4 = reinterpret_cast<const float4*>(vec);
      for (int j = lane; j < n4; j += 32) {
        const float4 a = row[j], b = v4[j];
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      acc = warp_sum(acc);
    }
    if (lane < cs) {
      float* remote = cluster.map_shared_rank(s.zpart, lane);
      remote[rank * rows + r] = acc;
    }
  }
}

template <bool kBackward>
__device__ __forceinline__ void ra_softmax_rows(const RowAttnSmem& s, int rows, int cs, const uint8_t* mask_b) {
  // warp 0: z[r] = sum over cluster partials (fixed order => identical in every CTA), masked softmax over rows
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float z[RA_MAX_ROWS / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      const int r = lane + 32 * i;
      float v = -INFINITY;
      if (r < rows && (mask_b == nullptr || mask_b[r] == 0)) {
        v = 0.f;
        for (int k = 0; k < cs; ++k) v += s.zpart[k * rows + r];
      }
      z[i] = v;
      mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      z[i] = (z[i] == -INFINITY) ? 0.f : expf(z[i] - mx);
      sum += z[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < rows) s.p[r] = z[i] * inv;
    }
  }
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Forward. Per CTA: (1) bulk-async stage the [rows x chunk] context slice and the target slice, (2) partial row dots into
// LOCAL smem, (3) ONE full cluster barrier, (4) pull the peers' partials through DSMEM (fixed rank order => identical sums
// everywhere), (5) arrive on the exit barrier (peers may still read our partials), softmax / shift, weighted sum from the
// resident tile, (6) wait on the exit barrier. Only step (3) sits on the critical path.
constexpr int RA_GATE_REGS = 12;   // float4 gate fragments a thread requests BEFORE it waits for the feature tile

template <int CS, bool SHIFT, bool GATE = false>
__global__ void __launch_bounds__(RA_THREADS) row_attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap, RowAttnArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (CS == 1) ? 0 : (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int rows = a.rows, chunk = a.chunk;
  const RowAttnSmem s = ra_carve(smem_raw, rows, chunk, CS);
  const int c0 = rank * chunk;
  const int cn = max(0, min(chunk, a.D - c0));
  const uint8_t* mask_b = (!SHIFT && a.mask) ? a.mask + (int64_t)b * a.ld_mask : nullptr;
  const float* ctx_b = a.ctx + (int64_t)b * a.ld_sample;
  const float* t_b = a.t + (int64_t)b * a.ld_t + c0;
  const bool t_bulk = ((reinterpret_cast<uintptr_t>(t_b) & 15) == 0) && cn > 0;

  const int nbox = a.nbox, boxw = a.boxw;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    mbar_init(s.bar, 1);
    mbar_fence_init();
    // one TMA box load per [rows x boxw] sub-tile (out-of-range channels are zero-filled) + the target slice
    mbar_expect_tx(s.bar, (uint32_t)nbox * (uint32_t)rows * (uint32_t)boxw * 4u + (t_bulk ? (uint32_t)cn * 4u : 0u));
    for (int sb = 0; sb < nbox; ++sb) tma_load_3d(s.tile + (size_t)sb * rows * boxw, &tmap, c0 + sb * boxw, 0, b, s.bar);
    if (t_bulk) bulk_g2s(s.tv, t_b, (uint32_t)cn * 4u, s.bar);
  }
  __syncthreads();                               // barrier initialised before anybody waits on it
  if (!t_bulk) {
    for (int c = threadIdx.x; c < cn; c += RA_THREADS) s.tv[c] = t_b[c];
    __syncthreads();
  }
  if (GATE) {
    // gate pre-activations of this slice: streamed once from HBM, requested before the tile wait so both streams are in flight
    const int b4g = boxw >> 2, per_box = rows * b4g, total4 = nbox * per_box;
    const float* g_b = a.gate + (int64_t)b * a.ld_gsample;
    auto locate = [&](int idx, int& c, float4*& tp) {
      const int sb = idx / per_box, rem = idx - sb * per_box, r = rem / b4g, cin = rem - r * b4g;
      c = c0 + sb * boxw + 4 * cin;
      tp = reinterpret_cast<float4*>(s.tile + ((size_t)sb * rows + r) * boxw) + cin;
      return r;
    };
    auto apply = [&](float4* tp, int c, const float4& g) {
      float4 sg = make_float4(sigmoidf_(g.x), sigmoidf_(g.y), sigmoidf_(g.z), sigmoidf_(g.w));
      if (a.chan_scale != nullptr) {
        const float4 cs4 = __ldg(reinterpret_cast<const float4*>(a.chan_scale + c));
        sg.x *= cs4.x; sg.y *= cs4.y; sg.z *= cs4.z; sg.w *= cs4.w;
      }
      float4 v = *tp;
      v.x *= sg.x; v.y *= sg.y; v.z *= sg.z; v.w *= sg.w;
      *tp = v;
    };
    float4 gr[RA_GATE_REGS];
#pragma unroll
    for (int i = 0; i < RA_GATE_REGS; ++i) {
      const int idx = threadIdx.x + i * RA_THREADS;
      gr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total4) {
        int c; float4* tp;
        const int r = locate(idx, c, tp);
        if (c < a.gate_C) gr[i] = ldg_stream4(g_b + (int64_t)r * a.ld_grow + c);
      }
    }
    mbar_wait(s.bar, 0);
#pragma unroll
    for (int i = 0; i < RA_GATE_REGS; ++i) {
      const int idx = threadIdx.x + i * RA_THREADS;
      if (idx < total4) {
        int c; float4* tp;
        locate(idx, c, tp);
        if (c < a.gate_C) apply(tp, c, gr[i]);
      }
    }
    for (int idx = threadIdx.x + RA_GATE_REGS * RA_THREADS; idx < total4; idx += RA_THREADS) {
      int c; float4* tp;
      const int r = locate(idx, c, tp);
      if (c < a.gate_C) apply(tp, c, ldg_stream4(g_b + (int64_t)r * a.ld_grow + c));
    }
    __syncthreads();
  } else {
    mbar_wait(s.bar, 0);
  }

  // partial dots of this channel slice -> local smem (aux doubles as the local partial array)
  {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b4 = boxw >> 2;
    for (int r = wid; r < rows; r += RA_THREADS / 32) {
      float acc = 0.f;
      if (mask_b == nullptr || mask_b[r] == 0) {
        for (int sb = 0; sb < nbox; ++sb) {
          const float4* row = reinterpret_cast<const float4*>(s.tile + ((size_t)sb * rows + r) * boxw);
          const float4* v4 = reinterpret_cast<const float4*>(s.tv + sb * boxw);
          const int lim = min(b4, (cn - sb * boxw + 3) >> 2);          // columns past cn: tile is zero-filled but tv is not staged
          for (int j = lane; j < lim; j += 32) {
            const float4 x = row[j], y = v4[j];
            acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
          }
        }
        acc = warp_sum(acc);
      }
      if (lane == 0) s.aux[r] = acc;
    }
  }
  if (CS > 1) cluster.sync(); else __syncthreads();
  // pull: z[r] = sum over ranks (fixed order) of their partials
  for (int r = threadIdx.x; r < rows; r += RA_THREADS) {
    float z = 0.f;
#pragma unroll
    for (int k = 0; k < CS; ++k) z += (CS == 1) ? s.aux[r] : cluster.map_shared_rank(s.aux, k)[r];
    s.zpart[r] = (mask_b != nullptr && mask_b[r]) ? -INFINITY : z;
  }
  if (CS > 1) cluster_arrive();                  // exit barrier, arrive half: done reading the peers' shared memory
  __syncthreads();
  // softmax over rows (warp 0)
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float z[RA_MAX_ROWS / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      const int r = lane + 32 * i;
      z[i] = (r < rows) ? s.zpart[r] : -INFINITY;
      mx = fmaxf(mx, z[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      z[i] = (z[i] == -INFINITY) ? 0.f : expf(z[i] - mx);
      sum += z[i];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int i = 0; i < RA_MAX_ROWS / 32; ++i) {
      const int r = lane + 32 * i;
      if (r < rows) s.p[r] = z[i] * inv;
    }
    if (SHIFT) {
      const int k = a.shift_k;
      float kl = (lane < k) ? a.kappa_logits[(int64_t)b * a.ld_kappa + lane] : -INFINITY;
      const float kmx = warp_max(kl);
      const float e = (lane < k) ? expf(kl - kmx) : 0.f;
      const float ksum = warp_sum(e);
      if (lane < k) {
        s.red[lane] = e / ksum;
        if (rank == 0 && a.kappa_out) a.kappa_out[(int64_t)b * k + lane] = e / ksum;
      }
    }
  }
  __syncthreads();
  if (SHIFT) {
    const int k = a.shift_k, half = k / 2, Hn = a.headings;
    for (int r = threadIdx.x; r < rows; r += RA_THREADS) {
      const int e = r / Hn, l = r % Hn;
      float q = 0.f;
      for (int j = 0; j < k; ++j) {
        int src = (l + j - half) % Hn;
        if (src < 0) src += Hn;
        q = fmaf(s.red[j], s.p[e * Hn + src], q);
      }
      s.w[r] = q;
    }
  } else {
    for (int r = threadIdx.x; r < rows; r += RA_THREADS) s.w[r] = s.p[r];
  }
  __syncthreads();
  if (rank == 0) {
    for (int r = threadIdx.x; r < rows; r += RA_THREADS) {
      if (a.attn_out) a.attn_out[(int64_t)b * rows + r] = s.p[r];
      if (a.q_out) a.q_out[(int64_t)b * rows + r] = s.w[r];
    }
  }

  // weighted sum from the resident tile: thread = (row group, float4 column)
  const int n4 = cn >> 2;
  if (n4 > 0) {
    const int G = max(1, min(RA_THREADS / n4, 8));
    const int passes = (n4 + RA_THREADS - 1) / RA_THREADS;
    for (int pass = 0; pass < passes; ++pass) {
      const int col = (G > 1) ? (int)(threadIdx.x % n4) : (int)threadIdx.x + pass * RA_THREADS;
      const int grp = (G > 1) ? (int)(threadIdx.x / n4) : 0;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const bool act = (col < n4) && (grp < G);
      if (act) {
        const int b4w = boxw >> 2, sbi = col / b4w, cin = col % b4w;
        const float4* tp = reinterpret_cast<const float4*>(s.tile + (size_t)sbi * rows * boxw) + cin;
        const int stride4 = b4w;
#pragma unroll 4
        for (int r = grp; r < rows; r += G) {
          const float wr = s.w[r];
          {
            const float4 v = tp[(size_t)r * stride4];
            acc.x = fmaf(wr, v.x, acc.x); acc.y = fmaf(wr, v.y, acc.y); acc.z = fmaf(wr, v.z, acc.z); acc.w = fmaf(wr, v.w, acc.w);
          }
        }
      }
      if (G > 1) {
        __syncthreads();
        if (act) reinterpret_cast<float4*>(s.red)[grp * n4 + col] = acc;
        __syncthreads();
        if (act && grp == 0) {
          for (int g2 = 1; g2 < G; ++g2) {
            const float4 o = reinterpret_cast<const float4*>(s.red)[g2 * n4 + col];
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
          }
        }
      }
      if (act && grp == 0) {
        float* dst = a.wc + (int64_t)b * a.ld_wc + c0 + 4 * col;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) *reinterpret_cast<float4*>(dst) = acc;
        else { dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w; }
      }
    }
  }
  if (CS > 1) cluster_wait();                    // exit barrier, wait half: nobody is still reading our partials
}

struct RowAttnBwdArgs {
  const float* ctx; int64_t ld_row, ld_sample; int B, rows, D;
  const float* t; int64_t ld_t; const float* attn; const float* q; const float* kappa;
  int shift_k, headings; const float* dwc; int64_t ld_dwc;
  float* dctx; int64_t ldd_row, ldd_sample; int dctx_accumulate;
  float* dt; int64_t ld_dt; float* dkappa_logits; int64_t ld_dkappa;
  int chunk;
};

template <int CS>
__global__ void __launch_bounds__(RA_THREADS) row_attention_bwd_kernel(RowAttnBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (CS == 1) ? 0 : (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int rows = a.rows, chunk = a.chunk;
  const RowAttnSmem s = ra_carve(smem_raw, rows, chunk, CS);
  const int c0 = rank * chunk;
  const int cn = max(0, min(chunk, a.D - c0));
  const float* ctx_b = a.ctx + (int64_t)b * a.ld_sample;

  if (threadIdx.x == 0) { mbar_init(s.bar, 1); mbar_fence_init(); }
  __syncthreads();
  ra_issue_loads(s, ctx_b, a.ld_row, rows, chunk, c0, cn, nullptr);
  for (int c = threadIdx.x; c < cn; c += RA_THREADS) {
    s.tv[c] = a.t[(int64_t)b * a.ld_t + c0 + c];
    s.dv[c] = a.dwc[(int64_t)b * a.ld_dwc + c0 + c];
  }
  for (int r = threadIdx.x; r < rows; r += RA_THREADS) {
    s.p[r] = a.attn[(int64_t)b * rows + r];
    s.w[r] = a.q[(int64_t)b * rows + r];
  }
  if (CS > 1) cluster.sync(); else __syncthreads();
  mbar_wait(s.bar, 0);

  // dq_r = ctx_r . dwc  (cluster-reduced)
  ra_partial_dots(cluster, s, s.dv, rows, chunk, cn, nullptr, CS, rank);
  if (CS > 1) cluster.sync(); else __syncthreads();

  // warp 0: dq -> dp (undo the shift), dkappa, dz = p * (dp - sum p dp)
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    float* dq = s.red;            // [rows]
    float* dp = s.red + RA_MAX_ROWS;
    for (int r = lane; r < rows; r += 32) {
      float v = 0.f;
      for (int k = 0; k < CS; ++k) v += s.zpart[k * rows + r];
      dq[r] = v;
    }
    __syncwarp();
    if (a.shift_k > 0) {
      const int k = a.shift_k, half = k / 2, Hn = a.headings;
      for (int r = lane; r < rows; r += 32) {
        const int e = r / Hn, m = r % Hn;
        float v = 0.f;
        for (int j = 0; j < k; ++j) {
          int src = (m - j + half) % Hn;
          if (src < 0) src += Hn;
          v = fmaf(a.kappa[(int64_t)b * k + j], dq[e * Hn + src], v);
        }
        dp[r] = v;
      }
      // dkappa_j = sum_{e,l} dq[e,l] * p[e,(l+j-half) mod Hn];  dkappa_logit = kappa * (dkappa - sum kappa dkappa)
      float dk_mine = 0.f, dot = 0.f;
      for (int j = 0; j < k; ++j) {
        float part = 0.f;
        for (int r = lane; r < rows; r += 32) {
          const int e = r / Hn, l = r % Hn;
          int src = (l + j - half) % Hn;
          if (src < 0) src += Hn;
          part = fmaf(dq[r], s.p[e * Hn + src], part);
        }
        part = warp_sum(part);
        dot = fmaf(a.kappa[(int64_t)b * k + j], part, dot);
        if (lane == j) dk_mine = part;
      }
      if (lane < k && rank == 0 && a.dkappa_logits)
        a.dkappa_logits[(int64_t)b * a.ld_dkappa + lane] = a.kappa[(int64_t)b * k + lane] * (dk_mine - dot);
    } else {
      for (int r = lane; r < rows; r += 32) dp[r] = dq[r];
    }
    __syncwarp();
    float pd = 0.f;
    for (int r = lane; r < rows; r += 32) pd = fmaf(s.p[r], dp[r], pd);
    pd = warp_sum(pd);
    for (int r = lane; r < rows; r += 32) s.aux[r] = s.p[r] * (dp[r] - pd);
  }
  __syncthreads();

  // dt[c] = sum_r dz_r ctx[r,c];  dctx[r,c] = w_r dwc[c] + dz_r t[c]
  const int n4 = cn >> 2;
  for (int col = threadIdx.x; col < n4; col += RA_THREADS) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rows; ++r) {
      const float dz = s.aux[r];
      if (dz != 0.f) {
        const float4 v = reinterpret_cast<const float4*>(s.tile + (size_t)r * chunk)[col];
        acc.x = fmaf(dz, v.x, acc.x); acc.y = fmaf(dz, v.y, acc.y); acc.z = fmaf(dz, v.z, acc.z); acc.w = fmaf(dz, v.w, acc.w);
      }
    }
    float* dst = a.dt + (int64_t)b * a.ld_dt + c0 + 4 * col;
    dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
  }
  if (a.dctx != nullptr && n4 > 0) {
    const int total = rows * n4;
    for (int i = threadIdx.x; i < total; i += RA_THREADS) {
      const int r = i / n4, col = i % n4;
      const float wr = s.w[r], dz = s.aux[r];
      const float4 dw = reinterpret_cast<const float4*>(s.dv)[col];
      const float4 tv = reinterpret_cast<const float4*>(s.tv)[col];
      float4 o = make_float4(fmaf(wr, dw.x, dz * tv.x), fmaf(wr, dw.y, dz * tv.y), fmaf(wr, dw.z, dz * tv.z),
                             fmaf(wr, dw.w, dz * tv.w));
      float4* dst = reinterpret_cast<float4*>(a.dctx + (int64_t)b * a.ldd_sample + (int64_t)r * a.ldd_row + c0 + 4 * col);
      if (a.dctx_accumulate) {
        const float4 old = *dst;
        o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
      }
      *dst = o;
    }
  }
}

// ------------------------------------------------------------------------------------------- candidate logits
__global__ void __launch_bounds__(256) cand_logits_fwd_kernel(const float* __restrict__ cand, int64_t ld_row, int64_t ld_sample,
                                                              int B, int Nc, int D, const float* __restrict__ t, int64_t ld_t,
                                                              const int32_t* __restrict__ leng, float* __restrict__ logit) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * Nc) return;
  const int b = row / Nc, c = row % Nc;
  if (leng != nullptr && c >= leng[b]) {
    if (lane == 0) logit[row] = -INFINITY;
    return;
  }
  const float4* x = reinterpret_cast<const float4*>(cand + (int64_t)b * ld_sample + (int64_t)c * ld_row);
  const float4* tv = reinterpret_cast<const float4*>(t + (int64_t)b * ld_t);
  float acc = 0.f;
  for (int j = lane; j < (D >> 2); j += 32) {
    const float4 a = x[j], w = __ldg(tv + j);
    acc = fmaf(a.x, w.x, acc); acc = fmaf(a.y, w.y, acc); acc = fmaf(a.z, w.z, acc); acc = fmaf(a.w, w.w, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) logit[row] = acc;
}

// grid (ceil(D/4/128), B): thread owns 4 channels of one sample; loops over that sample's candidates
__global__ void __launch_bounds__(128) cand_logits_bwd_kernel(const float* __restrict__ cand, int64_t ld_row, int64_t ld_sample,
                                                              int B, int Nc, int D, const float* __restrict__ t, int64_t ld_t,
                                                              const int32_t* __restrict__ leng, const float* __restrict__ dlogit,
                                                              float* __restrict__ dcand, int64_t ldd_row, int64_t ldd_sample,
                                                              int Dc, float* __restrict__ dt, int64_t ld_dt) {
  const int b = blockIdx.y;
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c4 >= (D >> 2)) return;
  const int n = leng ? min(Nc, leng[b]) : Nc;
  const float4 tv = *reinterpret_cast<const float4*>(t + (int64_t)b * ld_t + 4 * c4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < Nc; ++c) {
    const float g = (c < n) ? dlogit[(int64_t)b * Nc + c] : 0.f;
    if (c < n) {
      const float4 x = *reinterpret_cast<const float4*>(cand + (int64_t)b * ld_sample + (int64_t)c * ld_row + 4 * c4);
      acc.x = fmaf(g, x.x, acc.x); acc.y = fmaf(g, x.y, acc.y); acc.z = fmaf(g, x.z, acc.z); acc.w = fmaf(g, x.w, acc.w);
    }
    if (dcand != nullptr && 4 * c4 < Dc)
      *reinterpret_cast<float4*>(dcand + (int64_t)b * ldd_sample + (int64_t)c * ldd_row + 4 * c4) =
          make_float4(g * tv.x, g * tv.y, g * tv.z, g * tv.w);
  }
  if (dt != nullptr) *reinterpret_cast<float4*>(dt + (int64_t)b * ld_dt + 4 * c4) = acc;
}

int pick_cluster(int B, int rows, int D, int* chunk_out) {
  int cs = 1;
  auto chunk_of = [&](int c) { return (int)(dasa_cdiv(dasa_cdiv(D, 4), c) * 4); };
  // small channel slices => several CTAs per SM overlap their load / barrier / compute phases (measured: 2 CTAs/SM with
  // 88 KB slices reached 25 % of HBM peak); never below 64 floats per row slice
  while (cs < 8 && (ra_smem_bytes(rows, chunk_of(cs), cs) > 48 * 1024) && chunk_of(cs * 2) >= 64) cs *= 2;
  while (cs < 8 && ra_smem_bytes(rows, chunk_of(cs), cs) > 100 * 1024) cs *= 2;
  while (cs < 8 && (int64_t)B * cs * 2 <= 2 * DASA_NUM_SMS) cs *= 2;   // small batches: more CTAs in flight per sample
  *chunk_out = chunk_of(cs);
  return cs;
}

// batches at or above this size use the persistent pipelined kernel (DASA_RA_PIPE_MIN_B overrides, for the sweep)
int ra_pipe_min_batch() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DASA_RA_PIPE_MIN_B");
    v = e ? atoi(e) : 64;
    if (v < 1) v = 1;
  }
  return v;
}

template <typename Kern, typename... Args>
int launch_cluster(Kern kern, int B, int cs, size_t smem, cudaStream_t st, const char* name, Args... args) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(B * cs));
  cfg.blockDim = dim3(RA_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) { dasa_set_error(name, e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

}  // namespace

static int row_attention_fwd_impl(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                                  const float* t, int64_t ld_t, const uint8_t* mask, int64_t ld_mask, int shift_k,
                                  int headings, const float* kappa_logits, int64_t ld_kappa, float* wc, int64_t ld_wc,
                                  float* attn_out, float* q_out, float* kappa_out, const float* gate, int64_t ld_grow,
                                  int64_t ld_gsample, int gate_C, const float* chan_scale, void* stream) {
  if (B <= 0) return DASA_OK;
  if (rows <= 0 || rows > RA_MAX_ROWS || D <= 0 || D % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (shift_k < 0 || shift_k > RA_MAX_K || (shift_k > 0 && (headings <= 0 || rows % headings != 0 || kappa_logits == nullptr)))
    return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(ctx) || ld_row % 4 != 0 || ld_sample % 4 != 0) return DASA_ERR_BAD_ALIGN;
  if (mask == nullptr && B >= ra_pipe_min_batch()) {
    const int rc = dasa_row_attention_fwd_pipelined(ctx, ld_row, ld_sample, B, rows, D, t, ld_t, shift_k, headings, kappa_logits,
                                                    ld_kappa, wc, ld_wc, attn_out, q_out, kappa_out, (cudaStream_t)stream,
                                                    gate, ld_grow, ld_gsample, gate_C, chan_scale);
    if (rc != DASA_ERR_UNSUPPORTED) return rc;
  }
  RowAttnArgs a{ctx, ld_row, ld_sample, B, rows, D, t, ld_t, mask, ld_mask, shift_k, headings, kappa_logits, ld_kappa,
                wc, ld_wc, attn_out, q_out, kappa_out, 0, 1, 0, gate, ld_grow, ld_gsample, gate_C, chan_scale};
  const int cs = pick_cluster(B, rows, D, &a.chunk);
  // TMA boxes: inner extent <= 256 elements; widen the slice so that it is a whole number of equal boxes
  a.nbox = (int)dasa_cdiv(a.chunk, 256);
  a.boxw = (int)(dasa_cdiv(dasa_cdiv(a.chunk, a.nbox), 4) * 4);
  if (a.nbox > 1) {                              // every sub-tile must start 128-byte aligned: rows*boxw*4 % 128 == 0
    int g = 32;
    while (g > 1 && (rows % g) != 0) g >>= 1;    // g = gcd(rows, 32)
    const int q = 32 / g;
    a.boxw = (int)(dasa_cdiv(a.boxw, q) * q);
    if (a.boxw > 256) return DASA_ERR_BAD_SHAPE;
  }
  a.chunk = a.nbox * a.boxw;
  const size_t smem = ra_smem_bytes(rows, a.chunk, cs);
  if (smem > 227 * 1024) return DASA_ERR_BAD_SHAPE;
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn enc = reinterpret_cast<EncodeFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return DASA_ERR_UNSUPPORTED;
  CUtensorMap tmap;
  {
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld_row * 4, (cuuint64_t)ld_sample * 4};
    cuuint32_t box[3] = {(cuuint32_t)a.boxw, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ctx), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DASA_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
#define DASA_RA_FWD(CSV)                                                                                                  \
  (gate != nullptr ? launch_cluster(row_attention_fwd_kernel<CSV, true, true>, B, CSV, smem, st, "row_attention_fwd<gate,shift>", tmap, a) \
   : shift_k > 0 ? launch_cluster(row_attention_fwd_kernel<CSV, true>, B, CSV, smem, st, "row_attention_fwd<shift>", tmap, a) \
               : launch_cluster(row_attention_fwd_kernel<CSV, false>, B, CSV, smem, st, "row_attention_fwd<softdot>", tmap, a))
  switch (cs) {
    case 1: return DASA_RA_FWD(1);
    case 2: return DASA_RA_FWD(2);
    case 4: return DASA_RA_FWD(4);
    default: return DASA_RA_FWD(8);
  }
#undef DASA_RA_FWD
}

extern "C" int dasa_row_attention_fwd(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                                      const float* t, int64_t ld_t, const uint8_t* mask, int64_t ld_mask, int shift_k,
                                      int headings, const float* kappa_logits, int64_t ld_kappa, float* wc, int64_t ld_wc,
                                      float* attn_out, float* q_out, float* kappa_out, void* stream) {
  return row_attention_fwd_impl(ctx, ld_row, ld_sample, B, rows, D, t, ld_t, mask, ld_mask, shift_k, headings, kappa_logits, ld_kappa,
                                wc, ld_wc, attn_out, q_out, kappa_out, nullptr, 0, 0, 0, nullptr, stream);
}

extern "C" int dasa_gate_shift_attention_fwd(const float* f, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                                             const float* gate_pre, int64_t ld_grow, int64_t ld_gsample, int gate_C,
                                             const float* chan_scale, const float* t, int64_t ld_t, int shift_k, int headings,
                                             const float* kappa_logits, int64_t ld_kappa, float* wc, int64_t ld_wc,
                                             float* attn_out, float* q_out, float* kappa_out, void* stream) {
  if (gate_pre == nullptr || shift_k <= 0) return DASA_ERR_BAD_SHAPE;
  if (gate_C <= 0 || gate_C > D || gate_C % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(gate_pre) || ld_grow % 4 != 0 || ld_gsample % 4 != 0 || (chan_scale && !dasa_aligned16(chan_scale)))
    return DASA_ERR_BAD_ALIGN;
  return row_attention_fwd_impl(f, ld_row, ld_sample, B, rows, D, t, ld_t, nullptr, 0, shift_k, headings, kappa_logits, ld_kappa,
                                wc, ld_wc, attn_out, q_out, kappa_out, gate_pre, ld_grow, ld_gsample, gate_C, chan_scale, stream);
}

extern "C" int dasa_row_attention_bwd(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D,
                                      const float* t, int64_t ld_t, const float* attn, const float* q, const float* kappa,
                                      int shift_k, int headings, const float* dwc, int64_t ld_dwc, float* dctx,
                                      int64_t ldd_row, int64_t ldd_sample, int dctx_accumulate, float* dt, int64_t ld_dt,
                                      float* dkappa_logits, int64_t ld_dkappa, void* stream) {
  if (B <= 0) return DASA_OK;
  if (rows <= 0 || rows > RA_MAX_ROWS || D <= 0 || D % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (shift_k < 0 || shift_k > RA_MAX_K || (shift_k > 0 && (headings <= 0 || rows % headings != 0 || kappa == nullptr)))
    return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(ctx) || ld_row % 4 != 0 || ld_sample % 4 != 0) return DASA_ERR_BAD_ALIGN;
  if (dctx != nullptr && (!dasa_aligned16(dctx) || ldd_row % 4 != 0 || ldd_sample % 4 != 0)) return DASA_ERR_BAD_ALIGN;
  if (q == nullptr) q = attn;
  RowAttnBwdArgs a{ctx, ld_row, ld_sample, B, rows, D, t, ld_t, attn, q, kappa, shift_k, headings, dwc, ld_dwc,
                   dctx, ldd_row, ldd_sample, dctx_accumulate, dt, ld_dt, dkappa_logits, ld_dkappa, 0};
  const int cs = pick_cluster(B, rows, D, &a.chunk);
  const size_t smem = ra_smem_bytes(rows, a.chunk, cs);
  if (smem > 227 * 1024) return DASA_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cs) {
    case 1: return launch_cluster(row_attention_bwd_kernel<1>, B, 1, smem, st, "row_attention_bwd<1>", a);
    case 2: return launch_cluster(row_attention_bwd_kernel<2>, B, 2, smem, st, "row_attention_bwd<2>", a);
    case 4: return launch_cluster(row_attention_bwd_kernel<4>, B, 4, smem, st, "row_attention_bwd<4>", a);
    default: return launch_cluster(row_attention_bwd_kernel<8>, B, 8, smem, st, "row_attention_bwd<8>", a);
  }
}

extern "C" int dasa_cand_logits_fwd(const float* cand, int64_t ld_row, int64_t ld_sample, int B, int Nc, int D, const float* t,
                                    int64_t ld_t, const int32_t* cand_leng, float* logit, void* stream) {
  if (B <= 0 || Nc <= 0) return DASA_OK;
  if (D % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(cand) || ld_row % 4 != 0 || ld_sample % 4 != 0 || !dasa_aligned16(t) || ld_t % 4 != 0) return DASA_ERR_BAD_ALIGN;
  const int rows = B * Nc;
  cand_logits_fwd_kernel<<<(unsigned)dasa_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(cand, ld_row, ld_sample, B, Nc, D, t,
                                                                                        ld_t, cand_leng, logit);
  return dasa_check_launch("cand_logits_fwd_kernel");
}

extern "C" int dasa_cand_logits_bwd(const float* cand, int64_t ld_row, int64_t ld_sample, int B, int Nc, int D, const float* t,
                                    int64_t ld_t, const int32_t* cand_leng, const float* dlogit, float* dcand, int64_t ldd_row,
                                    int64_t ldd_sample, int Dc, float* dt, int64_t ld_dt, void* stream) {
  if (B <= 0 || Nc <= 0) return DASA_OK;
  if (D % 4 != 0 || Dc % 4 != 0) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(cand) || ld_row % 4 != 0 || ld_sample % 4 != 0 || !dasa_aligned16(t) || ld_t % 4 != 0) return DASA_ERR_BAD_ALIGN;
  if (dcand && (!dasa_aligned16(dcand) || ldd_row % 4 != 0 || ldd_sample % 4 != 0)) return DASA_ERR_BAD_ALIGN;
  if (dt && (!dasa_aligned16(dt) || ld_dt % 4 != 0)) return DASA_ERR_BAD_ALIGN;
  dim3 grid((unsigned)dasa_cdiv(D / 4, 128), (unsigned)B);
  cand_logits_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(cand, ld_row, ld_sample, B, Nc, D, t, ld_t, cand_leng, dlogit,
                                                                 dcand, ldd_row, ldd_sample, Dc, dt, ld_dt);
  return dasa_check_launch("cand_logits_bwd_kernel");
}

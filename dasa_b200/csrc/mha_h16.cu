// Multi-head attention of the frozen transformer stack on fp16 operands (include/dasa_b200.h: dasa_mha_fwd_h16):
// BertSelfAttention / BertXAttention forward (vilmodel.py:203-236, 479-506), dh = 64, forward only.
//
// Why a second kernel: mha_fwd_tc64_kernel (encoder.cu) launches one CTA per (sample, head) that loads ~22 KB, synchronises,
// computes for a microsecond and stores - 8400 short-lived CTAs per call, 71 % of the cycles without an eligible warp (ncu,
// profiles/r01_mha_fwd_tc64_ncu_full.txt), 160 us for 250 MB. Here
//   * Q / K / V arrive as fp16 (the QKV projections write them with c_half = 1): half the bytes; fp16 keeps the 10 mantissa
//     bits the TF32 kernel rounded its operands to, so the products are the same inside fp16's normal range;
//   * persistent CTAs walk the (sample, head) units with a 2-stage cp.async ring: the next unit's Q / K / V tiles stream into
//     shared memory while the current one is computed, so the load latency is off the critical path whatever the occupancy;
//   * fragments come from ldmatrix (V through ldmatrix.trans), products on mma.sync.m16n8k16 (half the instructions of the
//     TF32 k8 shape), softmax in fp32 on the accumulator fragment, which is re-used as the A operand of P.V;
//   * the keep flags of the attention-probability dropout are drawn in place from the counter-hash stream (rng.cuh) - no mask
//     tensor is written or read (nothing is back-propagated through these layers, so nobody re-reads it);
//   * the context tile goes back through the warp's own (already consumed) Q rows in shared memory and leaves as 128-bit stores.
#include <cuda_fp16.h>
#include "common.cuh"
#include "rng.cuh"

namespace {

constexpr int H16_DH = 64;
constexpr int H16_STRIDE = 72;        // halves per shared-memory row (144 B): 8 consecutive rows land in 8 distinct 16-byte bank groups
constexpr int H16_THREADS = 64;

struct MhaH16Args {
  const __half *q, *k, *v;
  int64_t ldq, sq, ldk, sk, ldv, sv;                     // halves
  const int32_t *q_off, *q_len, *k_off, *k_len;          // packed operands (NULL = dense)
  const uint8_t* key_pad; int64_t ld_pad;
  const uint8_t* drop_mask;                              // materialised keep mask [B, heads, Lq, Lk] (tests), or
  const unsigned long long* seed_dev; unsigned long long seed, base; uint32_t thr; int use_stream;   // in-place draws
  float drop_scale;
  void* out; int64_t ldo, so; int out_half;
  int B, heads, Lq, Lk;                                  // Lq / Lk: maxima
  int n_units;
  int stages;                                            // 1 or 2 shared-memory stages
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
  const int n = valid ? 16 : 0;                          // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// NKB = 16-key blocks the launch can hold (max_Lk <= 16 * NKB)
template <int NKB>
__global__ void __launch_bounds__(H16_THREADS, 8) mha_fwd_h16_kernel(MhaH16Args a) {
  extern __shared__ __align__(16) __half smem_h[];
  const int LqM = (a.Lq + 15) & ~15, LkM = (a.Lk + 15) & ~15;
  const int stage_halves = (LqM + 2 * LkM) * H16_STRIDE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  constexpr int nwarps = H16_THREADS / 32;
  const int nMT = LqM >> 4, nNT = LkM >> 3;              // tile counts of the dropout stream layout

  DropStream ds;
  ds.thr = a.thr;
  ds.base = a.base;
  ds.mixed = a.use_stream ? mix_seed(a.seed_dev ? a.seed_dev[0] : a.seed) : 0ull;

  auto load_unit = [&](int u, int s) {
    const int b = u / a.heads, h = u - b * a.heads;
    const int Lq = a.q_len ? a.q_len[b] : a.Lq, Lk = a.k_len ? a.k_len[b] : a.Lk;
    const int LqP = (Lq + 15) & ~15, LkP = (Lk + 15) & ~15;
    const __half* qb = a.q + (a.q_off ? (int64_t)a.q_off[b] * a.ldq : (int64_t)b * a.sq) + h * H16_DH;
    const __half* kb = a.k + (a.k_off ? (int64_t)a.k_off[b] * a.ldk : (int64_t)b * a.sk) + h * H16_DH;
    const __half* vb = a.v + (a.k_off ? (int64_t)a.k_off[b] * a.ldv : (int64_t)b * a.sv) + h * H16_DH;
    __half* Qs = smem_h + (size_t)s * stage_halves;
    __half* Ks = Qs + LqM * H16_STRIDE;
    __half* Vs = Ks + LkM * H16_STRIDE;
    // 64 threads = 8 rows x 8 16-byte chunks per pass; a thread keeps its chunk column and walks down the rows
    const int ch8 = (threadIdx.x & 7) * 8, rr = threadIdx.x >> 3;
    {
      const __half* src = qb + (int64_t)rr * a.ldq + ch8;
      __half* dst = Qs + rr * H16_STRIDE + ch8;
      for (int r = rr; r < LqP; r += 8, src += 8 * a.ldq, dst += 8 * H16_STRIDE) cp_async16(dst, r < Lq ? src : qb, r < Lq);
    }
    {
      const __half* srck = kb + (int64_t)rr * a.ldk + ch8;
      const __half* srcv = vb + (int64_t)rr * a.ldv + ch8;
      __half* dk = Ks + rr * H16_STRIDE + ch8;
      __half* dv = Vs + rr * H16_STRIDE + ch8;
      for (int r = rr; r < LkP; r += 8, srck += 8 * a.ldk, srcv += 8 * a.ldv, dk += 8 * H16_STRIDE, dv += 8 * H16_STRIDE) {
        cp_async16(dk, r < Lk ? srck : kb, r < Lk);
        cp_async16(dv, r < Lk ? srcv : vb, r < Lk);
      }
    }
  };

  // two stages (the next unit streams in under the current one) when they fit next to >= 6 resident CTAs, else one stage and the
  // other resident CTAs cover the load latency
  const bool ring = a.stages == 2;
  int u = blockIdx.x, s = 0;
  if (ring) {
    if (u < a.n_units) load_unit(u, 0);
    cp_async_commit();
  }
  for (; u < a.n_units; u += gridDim.x) {
    if (ring) {
      const int un = u + gridDim.x;
      if (un < a.n_units) load_unit(un, s ^ 1);
      cp_async_commit();
      cp_async_wait<1>();                                // everything but the prefetch just issued has landed
    } else {
      load_unit(u, 0);
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();

    const int b = u / a.heads, h = u - b * a.heads;
    const int Lq = a.q_len ? a.q_len[b] : a.Lq, Lk = a.k_len ? a.k_len[b] : a.Lk;
    const int nkb = (Lk + 15) >> 4;
    __half* Qs = smem_h + (size_t)s * stage_halves;
    const __half* Ks = Qs + LqM * H16_STRIDE;
    const __half* Vs = Ks + LkM * H16_STRIDE;
    const uint8_t* pad = a.key_pad ? a.key_pad + (int64_t)b * a.ld_pad : nullptr;

    for (int mt = warp; mt * 16 < Lq; mt += nwarps) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], Qs + (mt * 16 + (lane & 15)) * H16_STRIDE + ks * 16 + (lane >> 4) * 8);
      float acc[2 * NKB][4];
#pragma unroll
      for (int nt = 0; nt < 2 * NKB; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        if (kb < nkb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t kf[4];
            ldsm_x4(kf, Ks + (kb * 16 + (lane & 7) + ((lane >> 4) << 3)) * H16_STRIDE + ks * 16 + ((lane >> 3) & 1) * 8);
            mma_f16(acc[2 * kb], qa[ks], kf[0], kf[1]);
            mma_f16(acc[2 * kb + 1], qa[ks], kf[2], kf[3]);
          }
        }
      }
      // scores -> masked softmax per row; a row lives in the 4 lanes sharing g (cols 2t, 2t+1 of every key tile). Keys past Lk
      // (only in the last 16-key block) get -inf: exp() makes them exact zeros, and Lk >= 1 keeps every row maximum finite.
      const float scale = 0.125f;                          // 1 / sqrt(64)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2 * NKB; ++nt) {
        if (nt < 2 * nkb) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = nt * 8 + 2 * t + e;
            float add = (j < Lk) ? 0.f : -INFINITY;
            if (pad != nullptr && j < Lk && pad[j]) add = -10000.0f;
            acc[nt][e] = fmaf(acc[nt][e], scale, add);
            acc[nt][2 + e] = fmaf(acc[nt][2 + e], scale, add);
            mx0 = fmaxf(mx0, acc[nt][e]);
            mx1 = fmaxf(mx1, acc[nt][2 + e]);
          }
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2 * NKB; ++nt) {
        if (nt < 2 * nkb) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            acc[nt][e] = __expf(acc[nt][e] - mx0);
            acc[nt][2 + e] = __expf(acc[nt][2 + e] - mx1);
            s0 += acc[nt][e];
            s1 += acc[nt][2 + e];
          }
        }
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float inv0 = a.drop_scale / s0, inv1 = a.drop_scale / s1;
      // dropout on the probabilities. Materialised mask: byte [b, h, r, j] of the padded [B, heads, Lq, Lk] tensor. Stream:
      // the 4 accumulator elements a lane holds of score tile (mt, nt) are ONE hash word, stream byte
      //   ((((b * heads + h) * nMT + mt) * nNT + nt) * 32 + lane) * 4 + 2 * e + (row half)
      // (ops.mha_h16_stream_index gives the same map to the tests).
      if (a.use_stream) {
        // hash64(mixed, w) with w = w0 + 32 * nt: the (w + 1) * G term advances by 32 * G per tile
        const uint64_t w0 = ((((uint64_t)b * a.heads + h) * nMT + mt) * nNT) * 32 + lane;
        uint64_t zg = (ds.base + w0 + 1) * 0x9E3779B97F4A7C15ull;
#pragma unroll
        for (int nt = 0; nt < 2 * NKB; ++nt) {
          if (nt < 2 * nkb) {
            uint64_t z = ds.mixed ^ zg;
            zg += 32ull * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            const uint32_t lo = (uint32_t)z, hi = (uint32_t)(z >> 32);
            acc[nt][0] *= ((lo & 0xFFFFu) >= ds.thr) ? inv0 : 0.f;      // e = 0, row g
            acc[nt][2] *= ((lo >> 16) >= ds.thr) ? inv1 : 0.f;         // e = 0, row g + 8
            acc[nt][1] *= ((hi & 0xFFFFu) >= ds.thr) ? inv0 : 0.f;      // e = 1, row g
            acc[nt][3] *= ((hi >> 16) >= ds.thr) ? inv1 : 0.f;         // e = 1, row g + 8
          }
        }
      } else if (a.drop_mask != nullptr) {
        const int64_t mb0 = (((int64_t)b * a.heads + h) * a.Lq + r0) * a.Lk, mb1 = mb0 + (int64_t)8 * a.Lk;
#pragma unroll
        for (int nt = 0; nt < 2 * NKB; ++nt) {
          if (nt < 2 * nkb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = nt * 8 + 2 * t + e;
              const bool k0 = (j < Lk && r0 < Lq) ? (a.drop_mask[mb0 + j] != 0) : true;
              const bool k1 = (j < Lk && r1 < Lq) ? (a.drop_mask[mb1 + j] != 0) : true;
              acc[nt][e] *= k0 ? inv0 : 0.f;
              acc[nt][2 + e] *= k1 ? inv1 : 0.f;
            }
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 2 * NKB; ++nt) {
          if (nt < 2 * nkb) {
            acc[nt][0] *= inv0; acc[nt][1] *= inv0; acc[nt][2] *= inv1; acc[nt][3] *= inv1;
          }
        }
      }
      float o[8][4];
#pragma unroll
      for (int n8 = 0; n8 < 8; ++n8) { o[n8][0] = o[n8][1] = o[n8][2] = o[n8][3] = 0.f; }
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
        if (kb < nkb) {
          uint32_t pa[4];
          pa[0] = pack_h2(acc[2 * kb][0], acc[2 * kb][1]);          // row g,     keys 2t, 2t+1
          pa[1] = pack_h2(acc[2 * kb][2], acc[2 * kb][3]);          // row g + 8
          pa[2] = pack_h2(acc[2 * kb + 1][0], acc[2 * kb + 1][1]);  // row g,     keys 8 + 2t, ..
          pa[3] = pack_h2(acc[2 * kb + 1][2], acc[2 * kb + 1][3]);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t vf[4];
            ldsm_x4_t(vf, Vs + (kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * H16_STRIDE + np * 16 + (lane >> 4) * 8);
            mma_f16(o[2 * np], pa, vf[0], vf[1]);
            mma_f16(o[2 * np + 1], pa, vf[2], vf[3]);
          }
        }
      }
      if (a.out_half) {
        // through this warp's own Q rows (their fragments are in registers): 128-bit coalesced stores
        __half* Os = Qs + (mt * 16) * H16_STRIDE;
        __syncwarp();
#pragma unroll
        for (int n8 = 0; n8 < 8; ++n8) {
          *reinterpret_cast<uint32_t*>(Os + g * H16_STRIDE + n8 * 8 + 2 * t) = pack_h2(o[n8][0], o[n8][1]);
          *reinterpret_cast<uint32_t*>(Os + (g + 8) * H16_STRIDE + n8 * 8 + 2 * t) = pack_h2(o[n8][2], o[n8][3]);
        }
        __syncwarp();
        __half* oh = reinterpret_cast<__half*>(a.out) + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * H16_DH;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = (lane >> 3) + 4 * i, ch = lane & 7;
          if (mt * 16 + rr < Lq)
            *reinterpret_cast<uint4*>(oh + (int64_t)(mt * 16 + rr) * a.ldo + ch * 8) =
                *reinterpret_cast<const uint4*>(Os + rr * H16_STRIDE + ch * 8);
        }
      } else {
        float* ob = reinterpret_cast<float*>(a.out) + (a.q_off ? (int64_t)a.q_off[b] * a.ldo : (int64_t)b * a.so) + h * H16_DH;
#pragma unroll
        for (int n8 = 0; n8 < 8; ++n8) {
          const int d = n8 * 8 + 2 * t;
          if (r0 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r0 * a.ldo + d) = make_float2(o[n8][0], o[n8][1]);
          if (r1 < Lq) *reinterpret_cast<float2*>(ob + (int64_t)r1 * a.ldo + d) = make_float2(o[n8][2], o[n8][3]);
        }
      }
    }
    __syncthreads();                                     // stage s is overwritten by the next iteration's loads
    if (ring) s ^= 1;
  }
  cp_async_wait<0>();
}

}  // namespace

extern "C" int dasa_mha_fwd_h16(const dasa_half_t* q, int64_t ldq, int64_t sq, const int32_t* q_off, const int32_t* q_len,
                                const dasa_half_t* k, int64_t ldk, const dasa_half_t* v, int64_t ldv, int64_t skv,
                                const int32_t* k_off, const int32_t* k_len, const uint8_t* key_pad, int64_t ld_pad,
                                const uint8_t* drop_mask, const uint64_t* drop_seed_dev, uint64_t drop_seed, uint64_t drop_base,
                                float drop_p, float drop_scale, void* out, int64_t ldo, int64_t so, int out_half, int B, int heads,
                                int max_Lq, int max_Lk, int dh, void* stream) {
  if (B <= 0 || heads <= 0) return DASA_OK;
  if (max_Lq <= 0 || max_Lk <= 0) return DASA_ERR_BAD_SHAPE;
  if (dh != H16_DH || max_Lk > 96) return DASA_ERR_UNSUPPORTED;
  if ((q_off == nullptr) != (q_len == nullptr) || (k_off == nullptr) != (k_len == nullptr)) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(q) || !dasa_aligned16(k) || !dasa_aligned16(v) || !dasa_aligned16(out) || ldq % 8 || ldk % 8 || ldv % 8 ||
      sq % 8 || skv % 8 || ldo % (out_half ? 8 : 2) || so % (out_half ? 8 : 2))
    return DASA_ERR_BAD_ALIGN;
  const int LqM = (max_Lq + 15) & ~15, LkM = (max_Lk + 15) & ~15;
  const size_t stage_bytes = (size_t)(LqM + 2 * LkM) * H16_STRIDE * sizeof(__half);
  if (stage_bytes > 227 * 1024) return DASA_ERR_BAD_SHAPE;
  const int stages = (8 * (2 * stage_bytes + 1024) <= 227 * 1024) ? 2 : 1;
  const size_t smem = stages * stage_bytes;
  const bool use_stream = drop_mask == nullptr && drop_p > 0.f;
  MhaH16Args a{reinterpret_cast<const __half*>(q), reinterpret_cast<const __half*>(k), reinterpret_cast<const __half*>(v),
               ldq, sq, ldk, skv, ldv, skv, q_off, q_len, k_off, k_len, key_pad, ld_pad, drop_mask,
               reinterpret_cast<const unsigned long long*>(drop_seed_dev), (unsigned long long)drop_seed,
               (unsigned long long)drop_base, (uint32_t)(drop_p * 65536.0f), use_stream ? 1 : 0,
               (drop_mask != nullptr || use_stream) ? drop_scale : 1.f, out, ldo, so, out_half, B, heads, max_Lq, max_Lk, B * heads, stages};
  void (*kern)(MhaH16Args) = nullptr;
  switch (LkM >> 4) {
    case 1: kern = mha_fwd_h16_kernel<1>; break;
    case 2: kern = mha_fwd_h16_kernel<2>; break;
    case 3: kern = mha_fwd_h16_kernel<3>; break;
    case 4: kern = mha_fwd_h16_kernel<4>; break;
    case 5: kern = mha_fwd_h16_kernel<5>; break;
    default: kern = mha_fwd_h16_kernel<6>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { dasa_set_error("mha_fwd_h16 attr", e); return DASA_ERR_CUDA; }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, H16_THREADS, smem);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  const int grid = a.n_units < DASA_NUM_SMS * per_sm ? a.n_units : DASA_NUM_SMS * per_sm;
  kern<<<(unsigned)grid, H16_THREADS, smem, (cudaStream_t)stream>>>(a);
  return dasa_check_launch("mha_fwd_h16_kernel");
}

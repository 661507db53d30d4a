// Large-batch forward of the view attention (ShiftSoftDotAttention, model.py:307-353; plain soft-dot when shift_k == 0,
// no mask): PERSISTENT thread-block clusters, one sample per cluster per iteration, warp-specialised and pipelined through
// mbarriers so that the HBM loads never stop while earlier samples are being reduced.
//
//   * cluster of CS CTAs (1 CTA / SM); CTA `rank` owns the channel slice [rank*chunk, (rank+1)*chunk) of every sample its
//     cluster processes and keeps a ring of NS shared-memory stages. A stage = the [rows x chunk] context slice, staged as
//     nbox TMA boxes [rows x boxw] (boxw*4 bytes is an odd multiple of 16 when the shape allows it, so that 8 consecutive
//     rows fall into 8 different 16-byte bank groups) + the target slice (bulk copy), completing on the stage's `full`.
//   * producer warp : waits `empty[stage]`, issues the TMA loads of the next sample;
//   * dot warps     : lane = (row of an 8-row group, box), two groups per warp: box-row dot products, 2 shuffles to add the boxes,
//                     partial sums PUSHED into every CTA of the cluster with st.async (data + complete_tx on the receiver's
//                     `zfull` mbarrier - no cluster-wide barrier anywhere in the loop);
//   * softmax warps : (4, round-robin over samples) wait `zfull`, add the CS partials in fixed rank order (all CTAs get
//                     bit-identical weights), softmax over the rows, circular shift along the heading axis, publish the
//                     weights, arrive `wready[stage]`;
//   * weighted-sum warps: warp = box, lane = float4 column: sum_r w_r * slice[r, :] from the still-resident stage, store,
//                     arrive `empty[stage]`.
//
// GATE variant (fused DGAdaChannel epilogue -> shift attention, agent_dg.py:1544-1547 feeding model.py:327-345): the staged slice
// holds the RAW features; the gate pre-activations of the same slice arrive through a second TMA ring (NG stages, released as
// soon as they are consumed), 8 extra `gate warps` multiply the resident slice in place by sigmoid(g) (* chan_scale) for the
// channels below gate_C and hand the stage to the dot warps through `gated[stage]`. The modulated features never reach HBM.
#include <cooperative_groups.h>
#include <cuda.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int RP_DWARPS = 4;          // dot-product warps, 16 rows each (ceil(rows/16) of them are active)
constexpr int RP_MAX_BOX = 4;         // boxes per slice = weighted-sum warps
constexpr int RP_SWARPS = 4;          // softmax warps
constexpr int RP_WARPS = RP_DWARPS + RP_MAX_BOX + RP_SWARPS + 1;
constexpr int RP_THREADS = RP_WARPS * 32;
constexpr int RP_GWARPS = 8;          // gate warps (GATE variant only), placed after the producer warp
constexpr int RP_MAX_ROWS = 64;       // two rows per lane in a softmax warp
constexpr int RP_MAX_K = 15;

struct PipeArgs {
  const float* t; int64_t ld_t;
  const float* kappa_logits; int64_t ld_kappa;
  float* wc; int64_t ld_wc; float* attn_out; float* q_out; float* kappa_out;
  int B, rows, D, chunk, nbox, boxw, box_stride, shift_k, headings, nstages, nclusters;   // box_stride in floats
  long long* trace;      // debug: [CTA][sample][8] SM-clock stamps of the pipeline hand-offs (nullptr in production)
  int gate_C, ngstages;  // GATE: channels [0, gate_C) are modulated; depth of the gate ring
  const float* chan_scale;   // GATE: optional per-channel factor [gate_C] (drop_env noise, agent_dg.py:656)
};

// Shared memory of one CTA. NS stages; the partial-dot exchange ring is 2*NS deep: a peer can push sample j only after this
// CTA's softmax warps have consumed sample j - 2*NS (its load of j waits for its own weighted sum of j-NS, which waited for
// our dots of j-NS, whose load waited for our weighted sum - hence softmax - of j-2*NS).
struct PipeSmem {
  float* tile;      // [NS][nbox][box_stride]
  float* tv;        // [NS][chunk]            target slice
  float* wts;       // [NS][RPAD]             final weights of the stage's sample
  float* zbuf;      // [2*NS][CS][RPAD]       partial dots pushed by every CTA of the cluster
  float* pscr;      // [SWARPS][RPAD + 16]    softmax-warp scratch: p, kappa
  uint8_t* tab;     // [rows][k]              circular-shift source rows
  uint64_t* full;   // [NS]   TMA landed
  uint64_t* empty;  // [NS]   weighted-sum warps are done with the stage
  uint64_t* wready; // [NS]   weights published
  uint64_t* zfull;  // [2*NS] all partial dots of a sample arrived
  float* gtile;     // [NG][nbox][box_stride] GATE: gate pre-activations of the slice
  float* cscale;    // [chunk]                GATE: chan_scale slice (1 where absent)
  uint64_t* gfull;  // [NG]   gate slice landed
  uint64_t* gempty; // [NG]   gate warps are done with the gate slice
  uint64_t* gated;  // [NS]   the stage holds the modulated features
  size_t stage_floats;
};

__host__ __device__ inline int rp_rpad(int rows) { return (rows + 3) & ~3; }

__host__ __device__ inline size_t rp_smem_bytes(int rows, int chunk, int nbox, int box_stride, int cs, int ns, int k, int ng = 0) {
  const size_t rp = (size_t)rp_rpad(rows);
  size_t b = 4 * ((size_t)ns * nbox * box_stride + (size_t)ns * chunk + (size_t)ns * rp + (size_t)2 * ns * cs * rp +
                  (size_t)RP_SWARPS * (rp + 16));
  b += (((size_t)rows * (k > 0 ? k : 1)) + 15) & ~(size_t)15;
  b += 8 * (size_t)(5 * ns) + 128;
  if (ng > 0) b += 4 * ((size_t)ng * nbox * box_stride + (size_t)chunk) + 8 * (size_t)(2 * ng + ns) + 128;
  return b;
}

__device__ inline PipeSmem rp_carve(unsigned char* raw, const PipeArgs& a, int cs) {
  PipeSmem s;
  const int rp = rp_rpad(a.rows), ns = a.nstages, k = a.shift_k;
  s.stage_floats = (size_t)a.nbox * a.box_stride;               // box_stride*4 is a multiple of 128 bytes
  s.tile = reinterpret_cast<float*>(raw);
  s.tv = s.tile + (size_t)ns * s.stage_floats;
  s.wts = s.tv + (size_t)ns * a.chunk;
  s.zbuf = s.wts + (size_t)ns * rp;
  s.pscr = s.zbuf + (size_t)2 * ns * cs * rp;
  s.tab = reinterpret_cast<uint8_t*>(s.pscr + (size_t)RP_SWARPS * (rp + 16));
  uintptr_t b = reinterpret_cast<uintptr_t>(s.tab + (size_t)a.rows * (k > 0 ? k : 1));
  b = (b + 15) & ~uintptr_t(15);
  s.full = reinterpret_cast<uint64_t*>(b);
  s.empty = s.full + ns;
  s.wready = s.empty + ns;
  s.zfull = s.wready + ns;
  s.gtile = nullptr; s.cscale = nullptr; s.gfull = s.gempty = s.gated = nullptr;
  if (a.ngstages > 0) {                                         // GATE: appended after the barriers, 128-byte aligned
    // byte offset from `raw` (not a uintptr_t round trip): keeps the shared-memory address space visible to the compiler (LDS/STS)
    size_t off = (size_t)(reinterpret_cast<unsigned char*>(s.zfull + 2 * ns) - raw);
    off = (off + 127) & ~size_t(127);
    s.gtile = reinterpret_cast<float*>(raw + off);
    s.cscale = s.gtile + (size_t)a.ngstages * s.stage_floats;
    s.gfull = reinterpret_cast<uint64_t*>(s.cscale + a.chunk);
    s.gempty = s.gfull + a.ngstages;
    s.gated = s.gempty + a.ngstages;
  }
  return s;
}

__device__ __forceinline__ void tma_box_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// remote 4-byte store that also signals 4 transaction bytes on the receiver's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
// Waiting warps back off with nanosleep: a bare try_wait loop retried ~35 times per wait (measured) and took a third of the
// SM's issue slots away from the warps that had work. (try_wait with a 10 ms suspend-time hint instead of the sleep: no polling
// instructions at all, but the wake-up is ~80 cycles slower per hand-off and the kernels measure the same or 3 % slower.)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    __nanosleep(32);
    if (mbar_try(bar, parity)) return;
  }
  __trap();
}
__device__ __forceinline__ float warp_max_f32(float v) {       // sm_100a: one REDUX instead of five shuffle steps
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// ROWS / B4W (float4 columns per box row) > 0: compile-time shape (fully unrolled inner loops); 0: taken from the arguments
// KT > 0: compile-time shift taps (source rows hoisted into registers, fully unrolled); 0: runtime k through the table
template <int CS, int ROWS, int B4W, int KT, bool GATE>
__global__ void __launch_bounds__(RP_THREADS + (GATE ? RP_GWARPS * 32 : 0), 1)
row_attention_fwd_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap gmap, PipeArgs a) {
  constexpr int NTHREADS = RP_THREADS + (GATE ? RP_GWARPS * 32 : 0);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int rows = ROWS ? ROWS : a.rows, boxw = B4W ? 4 * B4W : a.boxw;
  const int chunk = a.chunk, NS = a.nstages, NZ = 2 * a.nstages, nbox = a.nbox;
  const int k = KT ? KT : a.shift_k, Hn = a.headings;
  const int rp = rp_rpad(rows);
  const PipeSmem s = rp_carve(smem_raw, a, CS);
  const int c0 = rank * chunk;
  const int cn = max(0, min(chunk, a.D - c0));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = (a.B - cid + a.nclusters - 1) / a.nclusters;          // samples of this cluster: cid, cid + nclusters, ...
  const uint32_t z_tx = (uint32_t)CS * (uint32_t)rows * 4u;
  long long* trace = a.trace ? a.trace + (size_t)blockIdx.x * 8 * ((a.B + a.nclusters - 1) / a.nclusters) : nullptr;
#define RP_STAMP(i_, slot_) do { if (trace && lane == 0) trace[(size_t)(i_) * 8 + (slot_)] = clock64(); } while (0)

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    for (int i = 0; i < NS; ++i) {
      mbar_init(&s.full[i], 1);
      mbar_init(&s.empty[i], (uint32_t)nbox);
      mbar_init(&s.wready[i], 1);
    }
    for (int i = 0; i < NZ; ++i) mbar_init(&s.zfull[i], 1);
    if (GATE) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&gmap) : "memory");
      for (int i = 0; i < a.ngstages; ++i) {
        mbar_init(&s.gfull[i], 1);
        mbar_init(&s.gempty[i], RP_GWARPS);
      }
      for (int i = 0; i < NS; ++i) mbar_init(&s.gated[i], RP_GWARPS);
    }
    mbar_fence_init();
    for (int i = 0; i < NZ; ++i) mbar_expect_tx(&s.zfull[i], z_tx);
  }
  // target-slice tails past D stay zero for the whole kernel (the tile tail is zero-filled by TMA); weight padding too
  for (int i = threadIdx.x; i < NS * chunk; i += NTHREADS) s.tv[i] = 0.f;
  for (int i = threadIdx.x; i < NS * rp; i += NTHREADS) s.wts[i] = 0.f;
  if (GATE)
    for (int i = threadIdx.x; i < chunk; i += NTHREADS)
      s.cscale[i] = (a.chan_scale != nullptr && c0 + i < a.gate_C) ? a.chan_scale[c0 + i] : 1.f;
  for (int i = threadIdx.x; i < rows * k; i += NTHREADS) {          // circular shift along the heading axis (model.py:333-349)
    const int r = i / k, jj = i % k;
    const int e = r / Hn, l = r % Hn;
    int src = (l + jj - k / 2) % Hn;
    if (src < 0) src += Hn;
    s.tab[i] = (uint8_t)(e * Hn + src);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster.sync();                               // every peer's barriers are initialised before any st.async reaches them

  if (wid < RP_DWARPS) {
    // ------------------------- dot warps: lane = (row rl of an 8-row group, box bl), two 8-row groups per warp (one target
    // load feeds two rows: the broadcast target loads cost as many shared-memory wavefronts as the tile loads)
    const int rl = lane & 7, bl = lane >> 3;
    const int pairs = (rows + 15) >> 4;
    if (wid < pairs) {
      constexpr int PP = CS / 4;                 // peers served by one lane
      uint32_t zb_remote[PP], zf_remote[PP];
#pragma unroll
      for (int pp = 0; pp < PP; ++pp) {
        zb_remote[pp] = map_to_rank(smem_u32(s.zbuf), (uint32_t)(bl * PP + pp));
        zf_remote[pp] = map_to_rank(smem_u32(s.zfull), (uint32_t)(bl * PP + pp));
      }
      const int b4w = boxw >> 2;
      const bool box_ok = bl < nbox;
      const int rA = wid * 16 + rl, rB = rA + 8;
      const bool okA = box_ok && rA < rows, okB = box_ok && rB < rows;
      // branch-free: lanes without a row / box re-read a valid address (broadcast, no extra wavefronts) and are masked below,
      // so the unrolled loads are not fenced in by a divergent region and run ahead of the FMAs
      const int offA = min(rA, rows - 1) * boxw, offB = min(rB, rows - 1) * boxw;
      for (int i = 0; i < n; ++i) {
        const int st = i % NS, zs = i % NZ;
        mbar_wait_sleep(GATE ? &s.gated[st] : &s.full[st], (uint32_t)(i / NS) & 1u);
        if (wid == 0) RP_STAMP(i, 1);
        const float* tile = s.tile + (size_t)st * s.stage_floats + (box_ok ? bl * a.box_stride : 0);
        const float4* tv4 = reinterpret_cast<const float4*>(s.tv + (size_t)st * chunk + (box_ok ? bl * boxw : 0));
        const float4* rowA = reinterpret_cast<const float4*>(tile + offA);
        const float4* rowB = reinterpret_cast<const float4*>(tile + offB);
        float4 accA = make_float4(0.f, 0.f, 0.f, 0.f), accB = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wid * 16 + 8 < rows) {               // uniform: the warp's second 8-row group exists
#pragma unroll(B4W ? B4W : 4)
          for (int j = 0; j < b4w; ++j) {
            const float4 y = tv4[j], xa = rowA[j], xb = rowB[j];
            accA.x = fmaf(xa.x, y.x, accA.x); accA.y = fmaf(xa.y, y.y, accA.y);
            accA.z = fmaf(xa.z, y.z, accA.z); accA.w = fmaf(xa.w, y.w, accA.w);
            accB.x = fmaf(xb.x, y.x, accB.x); accB.y = fmaf(xb.y, y.y, accB.y);
            accB.z = fmaf(xb.z, y.z, accB.z); accB.w = fmaf(xb.w, y.w, accB.w);
          }
        } else {
#pragma unroll(B4W ? B4W : 4)
          for (int j = 0; j < b4w; ++j) {
            const float4 y = tv4[j], xa = rowA[j];
            accA.x = fmaf(xa.x, y.x, accA.x); accA.y = fmaf(xa.y, y.y, accA.y);
            accA.z = fmaf(xa.z, y.z, accA.z); accA.w = fmaf(xa.w, y.w, accA.w);
          }
        }
        float vA = okA ? (accA.x + accA.y) + (accA.z + accA.w) : 0.f;
        float vB = okB ? (accB.x + accB.y) + (accB.z + accB.w) : 0.f;
        vA += __shfl_xor_sync(0xffffffffu, vA, 8);
        vB += __shfl_xor_sync(0xffffffffu, vB, 8);
        vA += __shfl_xor_sync(0xffffffffu, vA, 16);
        vB += __shfl_xor_sync(0xffffffffu, vB, 16);
#pragma unroll
        for (int pp = 0; pp < PP; ++pp) {        // peer bl*PP+pp: zbuf[zs][my rank][row]
          const uint32_t dst = zb_remote[pp] + 4u * (uint32_t)((zs * CS + rank) * rp), bar = zf_remote[pp] + 8u * (uint32_t)zs;
          if (rA < rows) st_async_f32(dst + 4u * (uint32_t)rA, vA, bar);
          if (rB < rows) st_async_f32(dst + 4u * (uint32_t)rB, vB, bar);
        }
        if (wid == 0) RP_STAMP(i, 2);
      }
    }
  } else if (wid < RP_DWARPS + RP_MAX_BOX) {
    // ------------------------------------------------- weighted-sum warps: warp = box, lane = float4 column(s) of the box
    const int box = wid - RP_DWARPS;
    if (box < nbox) {
      const int b4w = boxw >> 2;
      const int cA = min(lane, b4w - 1), cB = min(lane + 32, b4w - 1);   // boxw <= 256 floats: at most two float4 columns per lane
      const bool twocol = B4W ? (B4W > 32) : (b4w > 32);           // uniform: no predicated-off second column in the common case
      const bool actA = lane < b4w && box * boxw + 4 * cA < cn;          // inactive lanes re-read a valid column (broadcast) and do not store
      const bool actB = twocol && lane + 32 < b4w && box * boxw + 4 * cB < cn;
      const int rows4 = rows >> 2;
      for (int i = 0; i < n; ++i) {
        const int st = i % NS;
        const int b = cid + i * a.nclusters;
        mbar_wait_sleep(&s.wready[st], (uint32_t)(i / NS) & 1u);
        if (GATE) mbar_wait_sleep(&s.gated[st], (uint32_t)(i / NS) & 1u);   // direct acquire of the gate warps' writes (long complete)
        if (box == 0) RP_STAMP(i, 5);
        const float* tile = s.tile + (size_t)st * s.stage_floats + (size_t)box * a.box_stride;
        const float4* w4 = reinterpret_cast<const float4*>(s.wts + (size_t)st * rp);
        float4 accA = make_float4(0.f, 0.f, 0.f, 0.f), accB = make_float4(0.f, 0.f, 0.f, 0.f);
        auto fma_row = [&](float wr, int r) {
          {
            const float4 x = reinterpret_cast<const float4*>(tile + r * boxw)[cA];
            accA.x = fmaf(wr, x.x, accA.x); accA.y = fmaf(wr, x.y, accA.y); accA.z = fmaf(wr, x.z, accA.z); accA.w = fmaf(wr, x.w, accA.w);
          }
          if (twocol) {
            const float4 x = reinterpret_cast<const float4*>(tile + r * boxw)[cB];
            accB.x = fmaf(wr, x.x, accB.x); accB.y = fmaf(wr, x.y, accB.y); accB.z = fmaf(wr, x.z, accB.z); accB.w = fmaf(wr, x.w, accB.w);
          }
        };
#pragma unroll(ROWS ? ROWS / 4 : 2)
        for (int g = 0; g < rows4; ++g) {
          const float4 wv = w4[g];
          fma_row(wv.x, 4 * g); fma_row(wv.y, 4 * g + 1); fma_row(wv.z, 4 * g + 2); fma_row(wv.w, 4 * g + 3);
        }
        for (int r = 4 * rows4; r < rows; ++r) fma_row(s.wts[(size_t)st * rp + r], r);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[st]);                   // this warp no longer needs the stage
        if (box == 0) RP_STAMP(i, 6);
        float* dst = a.wc + (int64_t)b * a.ld_wc + c0 + box * boxw;
        if (actA) stg_stream4(dst + 4 * cA, accA);
        if (actB) stg_stream4(dst + 4 * cB, accB);
      }
    }
  } else if (wid < RP_DWARPS + RP_MAX_BOX + RP_SWARPS) {
    // --------- softmax warps (round-robin over samples): z_r = sum over ranks (fixed order), softmax, circular shift
    const int sw = wid - RP_DWARPS - RP_MAX_BOX;
    float* pw = s.pscr + (size_t)sw * (rp + 16);
    float* kw = pw + rp;
    const bool r0ok = lane < rows, r1ok = lane + 32 < rows;
    int src0[KT ? KT : 1], src1[KT ? KT : 1];
    if (KT) {
#pragma unroll
      for (int jj = 0; jj < (KT ? KT : 1); ++jj) {
        src0[jj] = r0ok ? s.tab[lane * k + jj] : 0;
        src1[jj] = r1ok ? s.tab[(lane + 32) * k + jj] : 0;
      }
    }
    for (int i = sw; i < n; i += RP_SWARPS) {
      const int st = i % NS, zs = i % NZ;
      const int b = cid + i * a.nclusters;
      const bool writer = (i % CS) == rank;      // one CTA of the cluster stores the per-sample distributions
      float kv = 0.f;
      if (k > 0) {                               // independent of the exchange: done while the partial dots are in flight
        const float kl = (lane < k) ? a.kappa_logits[(int64_t)b * a.ld_kappa + lane] : -INFINITY;
        const float kmx = warp_max_f32(kl);
        const float ke = (lane < k) ? expf(kl - kmx) : 0.f;
        kv = ke / warp_sum(ke);
        if (lane < k) kw[lane] = kv;
      }
      mbar_wait_sleep(&s.zfull[zs], (uint32_t)(i / NZ) & 1u);
      RP_STAMP(i, 3);
      const float* zb = s.zbuf + (size_t)zs * CS * rp;
      float z0 = -INFINITY, z1 = -INFINITY;
      if (r0ok) {
        z0 = 0.f;
#pragma unroll
        for (int c = 0; c < CS; ++c) z0 += zb[c * rp + lane];
      }
      if (r1ok) {
        z1 = 0.f;
#pragma unroll
        for (int c = 0; c < CS; ++c) z1 += zb[c * rp + lane + 32];
      }
      if (lane == 0) mbar_expect_tx(&s.zfull[zs], z_tx);            // re-arm the slot for sample i + 2*NS
      const float mx = warp_max_f32(fmaxf(z0, z1));
      const float e0 = r0ok ? expf(z0 - mx) : 0.f;
      const float e1 = r1ok ? expf(z1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      const float p0 = e0 * inv, p1 = e1 * inv;
      float* w = s.wts + (size_t)st * rp;
      if (k > 0) {
        if (r0ok) pw[lane] = p0;
        if (r1ok) pw[lane + 32] = p1;
        __syncwarp();
        float q0 = 0.f, q1 = 0.f;
        if (KT) {                                // independent loads, no table look-up on the critical path
#pragma unroll
          for (int jj = 0; jj < (KT ? KT : 1); ++jj) {
            const float kj = kw[jj];
            q0 = fmaf(kj, pw[src0[jj]], q0);
            q1 = fmaf(kj, pw[src1[jj]], q1);
          }
          if (r0ok) w[lane] = q0;
          if (r1ok) w[lane + 32] = q1;
        } else {
          if (r0ok) {
            for (int jj = 0; jj < k; ++jj) q0 = fmaf(kw[jj], pw[s.tab[lane * k + jj]], q0);
            w[lane] = q0;
          }
          if (r1ok) {
            for (int jj = 0; jj < k; ++jj) q1 = fmaf(kw[jj], pw[s.tab[(lane + 32) * k + jj]], q1);
            w[lane + 32] = q1;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.wready[st]);                  // release: weights visible to the weighted-sum warps
        RP_STAMP(i, 4);
        if (writer) {
          if (a.q_out) {
            if (r0ok) a.q_out[(int64_t)b * rows + lane] = q0;
            if (r1ok) a.q_out[(int64_t)b * rows + lane + 32] = q1;
          }
          if (a.kappa_out && lane < k) a.kappa_out[(int64_t)b * k + lane] = kv;
        }
      } else {
        if (r0ok) w[lane] = p0;
        if (r1ok) w[lane + 32] = p1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.wready[st]);
        if (writer && a.q_out) {
          if (r0ok) a.q_out[(int64_t)b * rows + lane] = p0;
          if (r1ok) a.q_out[(int64_t)b * rows + lane + 32] = p1;
        }
      }
      if (writer && a.attn_out) {
        if (r0ok) a.attn_out[(int64_t)b * rows + lane] = p0;
        if (r1ok) a.attn_out[(int64_t)b * rows + lane + 32] = p1;
      }
      __syncwarp();                              // pw / kw are rewritten by the next sample of this warp
    }
  } else if (GATE && wid >= RP_WARPS) {
    // ------------- gate warps: slice *= sigmoid(g) (* chan_scale) in place for the channels below gate_C (agent_dg.py:1544-1547).
    // sigmoid on raw MUFU ex2 / rcp: relative error <= ~4e-7 for |g| <= 16 (2-ulp ex2, 1-ulp rcp, argument rounding |g|*2^-24)
    constexpr int GT = RP_GWARPS * 32;
    const int gt = (int)threadIdx.x - RP_THREADS;
    const int NG = a.ngstages;
    const int b4w = boxw >> 2, per_box4 = rows * b4w;
    int ngb = 0;                                                   // boxes of this CTA that hold gated channels
    while (ngb < nbox && c0 + ngb * boxw < a.gate_C) ++ngb;
    const bool has_scale = a.chan_scale != nullptr;
    const int col0 = gt % b4w, dcol = GT % b4w;
    // sigmoid on bare MUFU: e = ex2(min(-g*log2(e), 60)) (the clamp bounds 1+e by 2^60 so that the product of two of them stays
    // finite; sigmoid(g) for g < -41.6 then reads 8.7e-19 instead of a smaller positive number), ONE rcp per channel pair:
    // 1/a = b * rcp(a*b), 1/b = a * rcp(a*b). Relative error <= ~5e-7 for |g| <= 16 (2-ulp ex2, 1-ulp rcp, argument rounding).
    auto ex2n = [](float g) {
      float e;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(g * -1.4426950408889634f, 60.f)));
      return 1.f + e;
    };
    auto rcpa = [](float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; };
    auto modulate = [&](float4& v, const float4& gv) {
      const float a0 = ex2n(gv.x), a1 = ex2n(gv.y), a2 = ex2n(gv.z), a3 = ex2n(gv.w);
      const float r01 = rcpa(a0 * a1), r23 = rcpa(a2 * a3);
      v.x *= a1 * r01; v.y *= a0 * r01; v.z *= a3 * r23; v.w *= a2 * r23;
    };
    // compile-time shape: IT float4 per thread and box, all loads of a box issued before the first MUFU
    constexpr int IT = (ROWS && B4W) ? (ROWS * B4W + GT - 1) / GT : 1;
    int cols[IT];
#pragma unroll
    for (int m = 0; m < IT; ++m) cols[m] = (gt + m * GT) % b4w;
    for (int i = 0; i < n; ++i) {
      const int st = i % NS, gs = i % NG;
      mbar_wait_sleep(&s.full[st], (uint32_t)(i / NS) & 1u);
      if (ngb > 0) mbar_wait_sleep(&s.gfull[gs], (uint32_t)(i / NG) & 1u);
      float4* tile_st = reinterpret_cast<float4*>(s.tile + (size_t)st * s.stage_floats);
      const float4* gate_st = reinterpret_cast<const float4*>(s.gtile + (size_t)gs * s.stage_floats);
      const int bs4 = a.box_stride >> 2;
      if (ROWS && B4W) {
        // all loads of a box are issued before its first MUFU. (A ping-pong prefetch of the next box was tried: slower under
        // the 80-register cap of the 672-thread block, 719 vs 530 us at B=4096.)
        float4 gv[1][IT], fv[1][IT];
        bool on[1][IT];
        auto fetch = [&](int sb, float4 (&g_)[IT], float4 (&f_)[IT], bool (&on_)[IT]) {
          const int climit = (a.gate_C - (c0 + sb * boxw)) >> 2;   // float4 columns of this box below gate_C
#pragma unroll
          for (int m = 0; m < IT; ++m) {
            const int idx = gt + m * GT;
            on_[m] = idx < per_box4 && cols[m] < climit;
            if (on_[m]) { g_[m] = gate_st[sb * bs4 + idx]; f_[m] = tile_st[sb * bs4 + idx]; }
          }
        };
        auto finish = [&](int sb, float4 (&g_)[IT], float4 (&f_)[IT], bool (&on_)[IT]) {
          const float4* cs4 = reinterpret_cast<const float4*>(s.cscale + sb * boxw);
#pragma unroll
          for (int m = 0; m < IT; ++m) {
            if (on_[m]) {
              modulate(f_[m], g_[m]);
              if (has_scale) {
                const float4 sc = cs4[cols[m]];
                f_[m].x *= sc.x; f_[m].y *= sc.y; f_[m].z *= sc.z; f_[m].w *= sc.w;
              }
              tile_st[sb * bs4 + gt + m * GT] = f_[m];
            }
          }
        };
        for (int sb = 0; sb < ngb; ++sb) {
          fetch(sb, gv[0], fv[0], on[0]);
          finish(sb, gv[0], fv[0], on[0]);
        }
      } else {
        for (int sb = 0; sb < ngb; ++sb) {
          float4* tile4 = tile_st + sb * bs4;
          const float4* g4 = gate_st + sb * bs4;
          const float4* cs4 = reinterpret_cast<const float4*>(s.cscale + sb * boxw);
          const int climit = (a.gate_C - (c0 + sb * boxw)) >> 2;
          int col = col0;
          for (int idx = gt; idx < per_box4; idx += GT) {
            if (col < climit) {
              const float4 gv = g4[idx];
              float4 v = tile4[idx];
              modulate(v, gv);
              if (has_scale) {
                const float4 sc = cs4[col];
                v.x *= sc.x; v.y *= sc.y; v.z *= sc.z; v.w *= sc.w;
              }
              tile4[idx] = v;
            }
            col += dcol;
            if (col >= b4w) col -= b4w;
          }
        }
      }
      __syncwarp();
      if (gt == 0) RP_STAMP(i, 7);
      if (lane == 0) {
        mbar_arrive(&s.gated[st]);                                 // release: the modulated slice is visible to the dot warps
        if (ngb > 0) mbar_arrive(&s.gempty[gs]);
      }
    }
  } else if (wid == RP_WARPS - 1 && lane == 0) {
    // ------------------------------------------------------------------------- producer: TMA loads into the stage ring
    const uint32_t stage_tx = (uint32_t)nbox * (uint32_t)rows * (uint32_t)boxw * 4u + (uint32_t)cn * 4u;
    int ngb = 0;
    if (GATE) while (ngb < nbox && c0 + ngb * boxw < a.gate_C) ++ngb;
    const uint32_t gate_tx = (uint32_t)ngb * (uint32_t)rows * (uint32_t)boxw * 4u;
    const int NG = GATE ? a.ngstages : 1;
    for (int j = 0; j < n; ++j) {
      const int st = j % NS;
      const int b = cid + j * a.nclusters;
      if (j >= NS) mbar_wait_sleep(&s.empty[st], (uint32_t)(j / NS - 1) & 1u);
      RP_STAMP(j, 0);
      mbar_expect_tx(&s.full[st], stage_tx);
      float* dst = s.tile + (size_t)st * s.stage_floats;
      for (int sb = 0; sb < nbox; ++sb) tma_box_3d(dst + (size_t)sb * a.box_stride, &tmap, c0 + sb * boxw, 0, b, &s.full[st]);
      if (cn > 0) bulk_g2s(s.tv + (size_t)st * chunk, a.t + (int64_t)b * a.ld_t + c0, (uint32_t)cn * 4u, &s.full[st]);
      if (GATE && ngb > 0) {
        const int gs = j % NG;
        if (j >= NG) mbar_wait_sleep(&s.gempty[gs], (uint32_t)(j / NG - 1) & 1u);
        mbar_expect_tx(&s.gfull[gs], gate_tx);
        float* gdst = s.gtile + (size_t)gs * s.stage_floats;
        for (int sb = 0; sb < ngb; ++sb) tma_box_3d(gdst + (size_t)sb * a.box_stride, &gmap, c0 + sb * boxw, 0, b, &s.gfull[gs]);
      }
    }
  }
#undef RP_STAMP
  __syncwarp();
  cluster.sync();                                // nobody exits while a peer may still push into its shared memory
}

long long* g_trace = nullptr;   // set by dasa_debug_row_attention_trace (profiling scripts only)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS, int ROWS, int B4W, int KT, bool GATE = false>
int launch_pipe(const CUtensorMap& tmap, const CUtensorMap& gmap, PipeArgs a, size_t smem, cudaStream_t st) {
  auto kern = row_attention_fwd_pipe_kernel<CS, ROWS, B4W, KT, GATE>;
  static int max_clusters = -1;                  // per (CS) instantiation; the smem request below is the worst case
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) { dasa_set_error("row_attention_fwd_pipe attr", e); return DASA_ERR_CUDA; }
  if (CS > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { dasa_set_error("row_attention_fwd_pipe cluster attr", e); return DASA_ERR_CUDA; }
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(RP_THREADS + (GATE ? RP_GWARPS * 32 : 0));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    cfg.gridDim = dim3(CS * DASA_NUM_SMS);
    cudaLaunchConfig_t q = cfg;
    q.dynamicSmemBytes = 200 * 1024;             // 1 CTA / SM by construction
    int nc = 0;
    e = cudaOccupancyMaxActiveClusters(&nc, kern, &q);
    if (e != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = DASA_NUM_SMS / CS / 2; }
    max_clusters = nc;
  }
  a.nclusters = a.B < max_clusters ? a.B : max_clusters;
  cfg.gridDim = dim3((unsigned)(a.nclusters * CS));
  e = cudaLaunchKernelEx(&cfg, kern, tmap, gmap, a);
  if (e != cudaSuccess) { dasa_set_error(GATE ? "row_attention_fwd_pipe<gate>" : "row_attention_fwd_pipe", e); return DASA_ERR_CUDA; }
  return DASA_OK;
}

}  // namespace

// Returns DASA_ERR_UNSUPPORTED when the shape does not fit this kernel (the caller then uses the one-shot cluster kernel).
// gate != nullptr: fused DGAdaChannel epilogue (GATE variant; requires shift_k > 0 only because that is the only caller).
int dasa_row_attention_fwd_pipelined(const float* ctx, int64_t ld_row, int64_t ld_sample, int B, int rows, int D, const float* t,
                                     int64_t ld_t, int shift_k, int headings, const float* kappa_logits, int64_t ld_kappa,
                                     float* wc, int64_t ld_wc, float* attn_out, float* q_out, float* kappa_out, cudaStream_t st,
                                     const float* gate, int64_t ld_grow, int64_t ld_gsample, int gate_C, const float* chan_scale) {
  if (rows > RP_MAX_ROWS || shift_k > RP_MAX_K || D % 4 != 0) return DASA_ERR_UNSUPPORTED;
  if (!dasa_aligned16(t) || ld_t % 4 != 0 || !dasa_aligned16(wc) || ld_wc % 4 != 0) return DASA_ERR_UNSUPPORTED;
  const bool gated = gate != nullptr;
  if (gated && (gate_C <= 0 || gate_C > D || gate_C % 4 != 0 || !dasa_aligned16(gate) || ld_grow % 4 != 0 || ld_gsample % 4 != 0))
    return DASA_ERR_UNSUPPORTED;
  // channel slice per CTA, split into <= 4 TMA boxes of equal width (<= 256 floats); prefer an odd number of float4 per box row.
  // Ring depths: plain = the deepest NS in [3, 8] that fits; gated = (NS, NG) from the list below (the gate ring is released
  // right after the modulation, so it can be shallower than the feature ring).
  static const int gate_rings[][2] = {{4, 3}, {4, 2}, {3, 2}};
  int best_cs = 0, best_ns = 0, best_ng = 0, best_nbox = 0, best_boxw = 0, best_stride = 0;
  for (int cs = 8; cs <= 16 && best_cs == 0; cs *= 2) {
    const int c4 = (int)dasa_cdiv(D / 4, cs);                        // float4 columns per CTA
    if ((int64_t)(cs - 1) * c4 * 4 >= D) continue;                   // every rank owns >= 1 column
    for (int pass = 0; pass < 2 && best_cs == 0; ++pass) {           // pass 0: odd box width only
      for (int nbox = 1; nbox <= RP_MAX_BOX; ++nbox) {
        const int b4 = (int)dasa_cdiv(c4, nbox);
        if (b4 * 4 > 256 || (pass == 0 && (b4 & 1) == 0)) continue;
        if ((int64_t)b4 * nbox != c4) continue;                      // equal boxes tile the slice exactly (no overlap with the peer)
        const int boxw = b4 * 4;
        const int stride = (int)(dasa_cdiv((int64_t)rows * boxw, 32) * 32);   // floats; 128-byte aligned box bases
        int ns = 0, ng = 0;
        if (!gated) {
          ns = 8;
          while (ns >= 3 && rp_smem_bytes(rows, nbox * boxw, nbox, stride, cs, ns, shift_k) > 227 * 1024) --ns;
          if (ns < 3) continue;
        } else {
          for (const auto& r : gate_rings)
            if (rp_smem_bytes(rows, nbox * boxw, nbox, stride, cs, r[0], shift_k, r[1]) <= 227 * 1024) { ns = r[0]; ng = r[1]; break; }
          if (ns == 0) continue;
        }
        best_cs = cs; best_ns = ns; best_ng = ng; best_nbox = nbox; best_boxw = boxw; best_stride = stride;
        break;
      }
    }
  }
  if (best_cs == 0) return DASA_ERR_UNSUPPORTED;
  const int cs = best_cs, ns = best_ns, nbox = best_nbox, boxw = best_boxw, chunk = best_nbox * best_boxw;
  EncodeFn enc = reinterpret_cast<EncodeFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return DASA_ERR_UNSUPPORTED;
  CUtensorMap tmap, gmap;
  cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld_row * 4, (cuuint64_t)ld_sample * 4};
  cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ctx), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return DASA_ERR_UNSUPPORTED;
  gmap = tmap;
  if (gated) {
    cuuint64_t gdims[3] = {(cuuint64_t)gate_C, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t gstrides[2] = {(cuuint64_t)ld_grow * 4, (cuuint64_t)ld_gsample * 4};
    if (enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(gate), gdims, gstrides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return DASA_ERR_UNSUPPORTED;
  }
  PipeArgs a{t, ld_t, kappa_logits, ld_kappa, wc, ld_wc, attn_out, q_out, kappa_out,
             B, rows, D, chunk, nbox, boxw, best_stride, shift_k, headings, ns, 0, g_trace, gate_C, best_ng, chan_scale};
  const size_t smem = rp_smem_bytes(rows, chunk, nbox, best_stride, cs, ns, shift_k, best_ng);
  const int b4 = boxw / 4;
  if (gated) {
    if (cs == 8 && rows == 36 && b4 == 17 && shift_k == 5) return launch_pipe<8, 36, 17, 5, true>(tmap, gmap, a, smem, st);
    if (cs == 16 && rows == 36 && b4 == 33 && shift_k == 5) return launch_pipe<16, 36, 33, 5, true>(tmap, gmap, a, smem, st);
    return cs == 8 ? launch_pipe<8, 0, 0, 0, true>(tmap, gmap, a, smem, st) : launch_pipe<16, 0, 0, 0, true>(tmap, gmap, a, smem, st);
  }
  if (cs == 8 && rows == 36 && b4 == 17 && shift_k == 5) return launch_pipe<8, 36, 17, 5>(tmap, gmap, a, smem, st);     // 36 x (2048+128), k=5
  if (cs == 16 && rows == 36 && b4 == 33 && shift_k == 5) return launch_pipe<16, 36, 33, 5>(tmap, gmap, a, smem, st);   // 36 x (4096+128), k=5
  return cs == 8 ? launch_pipe<8, 0, 0, 0>(tmap, gmap, a, smem, st) : launch_pipe<16, 0, 0, 0>(tmap, gmap, a, smem, st);
}

// Debug hook for scripts/: stamps of the pipeline hand-offs go to `buf` ([CTAs][ceil(B/clusters)][8] int64); nullptr = off.
extern "C" int dasa_debug_row_attention_trace(void* buf) {
  g_trace = static_cast<long long*>(buf);
  return DASA_OK;
}

// Persistent CTA-pair (tcgen05 cta_group::2) TF32 GEMM for the many-tile token-major projections:
//   C[M,N] = epi(alpha * A[M,K] * B[N,K]^T + beta*C),  A and B K-major fp32 in HBM.
//
// Why a second kernel: the 128x128 single-CTA kernel (gemm_tc.cu) pulls (128+128) x 32 fp32 per 1 MFLOP through L2 -> SM
// (32 FLOP/B) and ncu shows it pinned at the L2 output cap (3.0 GB in 300 us for M=20300 N=3072 K=768), well below the tensor
// pipe. Here two SMs of a TPC form one 256 x BN tile: each CTA stages ITS 128 rows of A and ITS BN/2 rows of B only, the
// leader's single thread issues tcgen05.mma.cta_group::2 (UMMA 256 x BN x 8) which reads both halves of B from both SMs'
// shared memory, and each CTA's TMEM receives its own 128 accumulator rows: 64 FLOP per L2 byte at BN = 256.
// The kernel is persistent (one pair per TPC, static tile schedule) with a double-buffered TMEM accumulator, so the eight
// epilogue warps drain tile i (tcgen05.ld -> smem transpose -> fused epilogue -> coalesced 128-bit stores) while the MMA
// thread already accumulates tile i+1.
//   warp 0 lane 0 : TMA producer (both CTAs; cp.async.bulk.tensor .cta_group::2 completing on the LEADER's mbarrier)
//   warp 1 lane 0 : MMA issuer (leader CTA only); tcgen05.commit multicast frees the stage / publishes the accumulator in both CTAs
//   warps 2..9    : epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4): two warps per SM sub-partition, so
//                   one warp's tcgen05.ld / shared-memory round trip hides behind the other's activation math
#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "gemm_common.cuh"
#include "lstm_epi.cuh"

namespace {

constexpr int P_BM = 128;            // rows of A per CTA (256 per pair)
constexpr int P_BK = 32;             // 32 fp32 = one 128-byte swizzle row
constexpr int P_EPI_WARPS = 8;       // two warps per TMEM lane quarter, each draining half of the tile's columns
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS;
constexpr int P_EPI_LD = 36;         // staging row stride (floats): conflict-free 128-bit rows

struct PairParams {
  int M, N, K;
  float alpha, beta;
  float* C[2]; int64_t ldc;      // one output per group (grouped launches: the two directions of the bi-LSTM recurrence)
  EpiParams ep;
  int tiles_n, tiles_mn;         // column tiles / tiles of one (group, K split)
  int M2, tiles_mn2;             // grouped launches: rows / tiles per K split of group 1 (its own row count; 0 rows = group absent)
  int c_half;                    // fp16 operand kernels: C holds IEEE halves (ldc in halves) instead of floats
  int splits, kb_per_split;      // split-K: partial sums go to C[g] + split * split_stride, epilogue NONE
  int64_t split_stride;
  int tiles_total, num_pairs;
};

// Internal epilogue id (not part of the C ABI): the LSTM cell of the packed bi-LSTM recurrence on the accumulator (lstm_epi.cuh)
constexpr int P_EPI_LSTM = 100;
template <int EPI> struct PairExt {};                       // per-launch extra parameters of an epilogue kind (none by default)
template <> struct PairExt<P_EPI_LSTM> { LstmEpi le; };

struct PairTile { int g, ks, mt, nt; };
__device__ __forceinline__ PairTile pair_decode(const PairParams& p, int w) {
  PairTile t;
  const int per_group0 = p.tiles_mn * p.splits;
  t.g = w >= per_group0 ? 1 : 0;
  int r = w - t.g * per_group0;
  const int tmn = t.g ? p.tiles_mn2 : p.tiles_mn;
  t.ks = r / tmn;
  r -= t.ks * tmn;
  t.mt = r / p.tiles_n;
  t.nt = r - t.mt * p.tiles_n;
  return t;
}

__device__ __forceinline__ void p_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void p_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t p_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void p_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t p_mapa(uint32_t local, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank)); return r;
}
__device__ __forceinline__ void p_remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair-aware tile load: lands in the executing CTA's shared memory, completes its bytes on the mbarrier at `bar_cluster`
__device__ __forceinline__ void p_tma_load_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_commit_pair(uint64_t* bar) {       // arrives on `bar` at the same offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void p_umma_tf32_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void p_umma_f16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// 4 consecutive outputs of one row: floats, or halves when the kernel writes an fp16 activation for the next fp16 GEMM
__device__ __forceinline__ void p_store4(float* C, int64_t row_off, int n, const float (&v)[4], bool c_half) {
  if (c_half) {
    __half2 h[2] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3])};
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(C) + row_off + n) = *reinterpret_cast<uint2*>(h);
  } else {
    *reinterpret_cast<float4*>(C + row_off + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
__device__ __forceinline__ uint64_t p_desc_sw128(const void* smem_ptr) {   // K-major, 128B swizzle, 8-row atoms 1024 B apart
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (tcgen05 accepts MN-major TF32 operands, unlike wgmma). For 32-bit elements the only MN-major shared-memory
// layout the tensor core reads is the 128-byte swizzle with 32-byte atomicity (descriptor layout type 1; byte-address swizzle
// <2,5,2>: the 32 B chunk index is XORed with the row index mod 4) = what a TMA box {32 mn, P_BK k} written with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces: per 32-wide chunk of the M/N extent, P_BK rows (one per k) of 128 bytes.
// Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units: LBO = distance between 32-wide M/N chunks (P_BK * 128 B),
// SBO = distance between groups of 4 k rows (512 B). One UMMA (K = 8) consumes two 4-row groups of every chunk.
__device__ __forceinline__ uint64_t p_desc_sw128_mn(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((P_BK * 128) >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// MN-major 16-bit operand: the ordinary 128-byte swizzle (16-byte atoms, descriptor layout type 2). A TMA box {64 mn, BKE k}
// written with CU_TENSOR_MAP_SWIZZLE_128B gives, per 64-wide chunk of the M/N extent, BKE rows (one per k) of 128 bytes swizzled in
// groups of 8 rows (1024 B). Canonical form ((8,8,n),(8,k)):((1,8,LBO),(64,SBO)) in elements: LBO = distance between 64-wide M/N
// chunks (BKE * 128 B), SBO = distance between groups of 8 k rows (1024 B). One UMMA (K = 16) consumes two groups of every chunk.
__device__ __forceinline__ uint64_t p_desc_sw128_mn16(const void* smem_ptr, int bke) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((bke * 128) >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void p_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Fused LSTM cell (P_EPI_LSTM): the calling epilogue warp owns rows row0 .. row0+31 and the 128 accumulator columns at t_addr =
// gates i, f, g, o of hidden units unit0 .. unit0+31 (lstm_epi.cuh). The 32 x 128 accumulator block goes through `stg` (a
// [32][P_LSTM_LD] fp32 staging area in the - by then idle - operand ring: the launch gives every CTA pair exactly ONE tile) so
// that every global access is coalesced: a lane owns 4 consecutive units, 8 lanes cover the 32 units of one row (128 contiguous
// bytes of xp / acts / c / h / out per gate), a warp instruction covers 4 rows. (Thread = row straight out of TMEM touched 32
// different lines per instruction: 62 us per step instead of 35 for the two-launch form.) The summation order per gate is the
// pointwise kernel's: ((x + gh) + b_ih) + b_hh. `release()` hands the accumulator buffer back after the last TMEM read.
// Fused LSTM cell (P_EPI_LSTM): the calling epilogue warp owns rows row0 .. row0+31 and the 128 accumulator columns at t_addr =
// gates i, f, g, o of hidden units unit0 .. unit0+31 (lstm_epi.cuh). The 32 x 128 accumulator block goes through `stg` (a
// [32][P_LSTM_LD] fp32 staging area in the - by then idle - operand ring: the launch gives every CTA pair exactly ONE tile) so
// that every global access is coalesced: a lane owns 4 consecutive units, 8 lanes cover the 32 units of one row (128 contiguous
// bytes of xp / acts / c / h / out per gate), a warp instruction covers 4 rows. (Thread = row straight out of TMEM touched 32
// different lines per instruction: 62 us per step instead of 35 for the two-launch form.) Rows go in batches of 2 row groups whose
// xp / c_prev loads are issued one batch AHEAD of the cell update that uses them - the first batch and the bias columns before the
// warp even waits for the accumulator (none of it depends on the GEMM). The summation order per gate is the pointwise kernel's:
// ((x + gh) + b_ih) + b_hh.
constexpr int P_LSTM_LD = 132;       // staging row stride in floats: conflict-free 128-bit rows for both access patterns
constexpr int P_LSTM_NB = 2;         // row groups (of 4 rows) per batch
struct LstmBatch { float4 x[P_LSTM_NB][4]; float4 cp[P_LSTM_NB]; int pr[P_LSTM_NB]; };
// per-lane constants of a tile: the direction's pointers already offset to this lane's 4 units (indexing the kernel parameter
// with the run-time direction costs a constant-bank indirection per access), bias columns, the dropout stream
struct LstmLane {
  float4 bi[4], bh[4];
  const float* xp; const float* c_prev; float* c_out; float* acts; float* h_next; __half* h16_next; float* h_fin; float* c_fin;
  float* out; const int32_t* perm;
  int n, n_next, H, L, col4, rsub;
  int64_t out_off;               // (pos * 2 + d) * H + u: offset of this lane's units inside an output row block
  DropStream ds; const uint8_t* mask; float scale; int drop_mode;    // 0 none, 1 mask tensor, 2 in-place draws
};

__device__ __forceinline__ void p_lstm_lane(LstmLane& ln, const LstmEpi& le, const int d, const int unit0) {
  const int lane = threadIdx.x & 31;
  ln.col4 = lane & 7; ln.rsub = lane >> 3;
  const int u = unit0 + 4 * ln.col4, H = le.H;
  ln.H = H; ln.L = le.L; ln.n = le.n[d]; ln.n_next = le.n_next[d];
  ln.xp = le.xp[d] + u; ln.c_prev = le.c_prev[d] + u; ln.c_out = le.c_out[d] + u; ln.acts = le.acts[d] + u;
  ln.h_next = le.h_next[d] + u; ln.h16_next = le.h16_next[d] + u; ln.h_fin = le.h_fin[d] + u; ln.c_fin = le.c_fin[d] + u;
  ln.out = le.out; ln.perm = le.perm;
  ln.out_off = ((int64_t)le.pos[d] * 2 + d) * H + u;
  ln.mask = le.drop.mask; ln.scale = le.drop.scale;
  ln.drop_mode = le.drop.mask != nullptr ? 1 : (le.drop.stream ? 2 : 0);
  if (ln.drop_mode == 2) {
    ln.ds.mixed = mix_seed(le.drop.seed_dev ? le.drop.seed_dev[0] : le.drop.seed); ln.ds.base = le.drop.base; ln.ds.thr = le.drop.thr;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {          // this lane's 4 units keep their bias columns for every row of the tile
    ln.bi[q] = __ldg(reinterpret_cast<const float4*>(le.b_ih[d] + q * H + u));
    ln.bh[q] = __ldg(reinterpret_cast<const float4*>(le.b_hh[d] + q * H + u));
  }
}
// Pulls the xp / c_prev lines of the warp's 32 rows x 32 units into L2 while the tensor cores work (no registers held): the
// epilogue's loads then are L2 hits - its registers cannot keep the ~100 KB per SM in flight that DRAM latency would need.
__device__ __forceinline__ void p_lstm_prefetch(const LstmEpi& le, const int d, const int row0, const int unit0) {
  const int lane = threadIdx.x & 31;
  const int H = le.H, n = le.n[d];
  const int q = lane & 3;
#pragma unroll
  for (int rr = 0; rr < 32; rr += 8) {
    const int r = row0 + rr + (lane >> 2);
    if (r < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(le.xp[d] + (int64_t)r * 4 * H + q * H + unit0));
  }
  if (row0 + lane < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(le.c_prev[d] + (int64_t)(row0 + lane) * H + unit0));
}
__device__ __forceinline__ void p_lstm_load(LstmBatch& b, const LstmLane& ln, const int row0, const int rb) {
  const int H = ln.H;
#pragma unroll
  for (int i = 0; i < P_LSTM_NB; ++i) {
    const int r = row0 + rb + 4 * i + ln.rsub;
    const int rc = r < ln.n ? r : ln.n - 1;                   // clamped: the loads are unconditional, dead rows discard them
    const float* xrow = ln.xp + (int64_t)rc * 4 * H;
#pragma unroll
    for (int q = 0; q < 4; ++q) b.x[i][q] = ldg_stream4(xrow + q * H);
    b.cp[i] = *reinterpret_cast<const float4*>(ln.c_prev + (int64_t)rc * H);
    b.pr[i] = __ldg(ln.perm + rc);
  }
}
__device__ __forceinline__ void p_lstm_cell(const LstmBatch& b, const LstmLane& ln, const int row0, const int rb, const float* stg) {
  const int H = ln.H;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < P_LSTM_NB; ++i) {
    const int rl = rb + 4 * i + ln.rsub, r = row0 + rl;
    const bool live = r < ln.n, fresh = !live && r < ln.n_next;
    const int64_t rH = (int64_t)r * H;
    if (live) {
      float g[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a4 = *reinterpret_cast<const float4*>(stg + rl * P_LSTM_LD + q * 32 + 4 * ln.col4);
        g[q][0] = ((b.x[i][q].x + a4.x) + ln.bi[q].x) + ln.bh[q].x; g[q][1] = ((b.x[i][q].y + a4.y) + ln.bi[q].y) + ln.bh[q].y;
        g[q][2] = ((b.x[i][q].z + a4.z) + ln.bi[q].z) + ln.bh[q].z; g[q][3] = ((b.x[i][q].w + a4.w) + ln.bi[q].w) + ln.bh[q].w;
      }
      const float cp[4] = {b.cp[i].x, b.cp[i].y, b.cp[i].z, b.cp[i].w};
      float c1[4], h1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[0][e] = lstm_sigmoid(g[0][e]); g[1][e] = lstm_sigmoid(g[1][e]); g[2][e] = lstm_tanh(g[2][e]); g[3][e] = lstm_sigmoid(g[3][e]);
        c1[e] = g[1][e] * cp[e] + g[0][e] * g[2][e];
        h1[e] = g[3][e] * lstm_tanh(c1[e]);
      }
      float* arow = ln.acts + (int64_t)r * 4 * H;
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(arow + q * H) = make_float4(g[q][0], g[q][1], g[q][2], g[q][3]);
      const float4 c4 = make_float4(c1[0], c1[1], c1[2], c1[3]);
      float4 h4 = make_float4(h1[0], h1[1], h1[2], h1[3]);
      *reinterpret_cast<float4*>(ln.c_out + rH) = c4;
      if (r < ln.n_next) {
        *reinterpret_cast<float4*>(ln.h_next + rH) = h4;
        __half2 hh[2] = {__floats2half2_rn(h1[0], h1[1]), __floats2half2_rn(h1[2], h1[3])};
        *reinterpret_cast<uint2*>(ln.h16_next + rH) = *reinterpret_cast<uint2*>(hh);
      } else {        // last token of this sequence in this direction
        *reinterpret_cast<float4*>(ln.h_fin + rH) = h4;
        *reinterpret_cast<float4*>(ln.c_fin + rH) = c4;
      }
      const int64_t oe = (int64_t)b.pr[i] * ln.L * 2 * H + ln.out_off;      // out[perm[r], pos, d * H + u]
      if (ln.drop_mode == 1) {
        const uint32_t m = *reinterpret_cast<const uint32_t*>(ln.mask + oe);
        h4.x *= (m & 0xFFu) ? ln.scale : 0.f; h4.y *= (m & 0xFF00u) ? ln.scale : 0.f;
        h4.z *= (m & 0xFF0000u) ? ln.scale : 0.f; h4.w *= (m & 0xFF000000u) ? ln.scale : 0.f;
      } else if (ln.drop_mode == 2) {
        const uint32_t k4 = stream_keep4(ln.ds, (uint64_t)oe >> 2);
        h4.x *= (k4 & 1u) ? ln.scale : 0.f; h4.y *= (k4 & 2u) ? ln.scale : 0.f;
        h4.z *= (k4 & 4u) ? ln.scale : 0.f; h4.w *= (k4 & 8u) ? ln.scale : 0.f;
      }
      *reinterpret_cast<float4*>(ln.out + oe) = h4;
    } else if (fresh) {   // reverse direction: this sequence joins at the NEXT step with a zero state
      *reinterpret_cast<float4*>(ln.h_next + rH) = z4;
      *reinterpret_cast<uint2*>(ln.h16_next + rH) = make_uint2(0u, 0u);
      *reinterpret_cast<float4*>(ln.c_out + rH) = z4;
    }
  }
}
// accumulator block -> staging (thread = row = TMEM lane); `release()` hands the TMEM buffer back after the last read
template <typename Release>
__device__ __forceinline__ void p_lstm_stage(const uint32_t t_addr, float* stg, Release release) {
  float* srow = stg + (threadIdx.x & 31) * P_LSTM_LD;
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    uint32_t r[32];
    p_tmem_ld32(t_addr + (uint32_t)(q * 32), r);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(srow + q * 32 + 4 * j) =
          make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
  }
  release();
  __syncwarp();
}

// AMN / BMN: the operand is MN-major in HBM (A stored [K][M], B stored [K][N]): dX = dY.W and dW = dY^T.X run without any
// transposed copy of the activations / weights.
// F16: both operands are IEEE fp16 in HBM (K-major, 64 elements = one 128-byte swizzle row per stage row), tcgen05 kind::f16 at
// twice the TF32 rate, fp32 accumulate. fp16 carries TF32's 10 mantissa bits, so for operands inside fp16's normal range the
// products equal the TF32 kernel's: used for the frozen, forward-only transformer stack, whose activations are O(1..100).
// EW: epilogue warps per CTA (8, or 16 = four per TMEM lane quarter, each draining a quarter of the tile's columns: the erf-GELU
// epilogue is FMA-pipe work - ~10 fma-pipe instructions per element - that two warps per scheduler cannot overlap with their own
// TMEM / shared-memory round trips; with four the tile's epilogue drops under its fp16 main loop).
template <int BN, int STAGES, int EPI, bool AMN = false, bool BMN = false, bool F16 = false, int EW = P_EPI_WARPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, PairParams p,
                      const __grid_constant__ PairExt<EPI> ext) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int A_BYTES = P_BM * P_BK * 4, B_BYTES = (BN / 2) * P_BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = 2 * BN;                      // two accumulator buffers of BN fp32 columns
  unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* stage_all = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_all + EW * 32 * P_EPI_LD);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;              // [2]
  uint64_t* tmem_empty = tmem_full + 2;                  // [2], the leader's copy is the one that counts
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = p_ctarank();
  const int pair = blockIdx.x >> 1;
  constexpr int BKE = F16 ? 2 * P_BK : P_BK;             // K elements per pipeline stage (128 bytes per operand row either way)
  const int nkb = (p.K + BKE - 1) / BKE;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 2 * EW); }   // every epilogue warp of both CTAs
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  p_fence_before();
  __syncthreads();
  p_cluster_sync();                                       // the peer's barriers and TMEM exist before anything reaches them
  p_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // -------------------------------------------------------------- TMA producer (both CTAs)
      uint32_t it = 0;
      for (int tile = pair; tile < p.tiles_total; tile += p.num_pairs) {
        const PairTile t = pair_decode(p, tile);
        const int m0 = t.mt * (2 * P_BM) + (int)crank * P_BM;
        const int nb0 = t.nt * BN + (int)crank * (BN / 2);
        const CUtensorMap* ma = t.g ? &tmA1 : &tmA;
        const CUtensorMap* mb = t.g ? &tmB1 : &tmB;
        const int kb0 = t.ks * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);      // both CTAs' boxes complete on the leader's barrier
          const uint32_t fb = p_mapa(smem_u32(&full_bar[s]), 0);
          unsigned char* dst = tiles + s * STAGE_BYTES;
          constexpr int MNC = F16 ? 64 : 32;             // M/N elements per 128-byte row of an MN-major chunk
          if constexpr (AMN) {
#pragma unroll
            for (int c = 0; c < P_BM / MNC; ++c) p_tma_load_2sm(dst + c * (BKE * 128), ma, m0 + MNC * c, kb * BKE, fb);
          } else {
            p_tma_load_2sm(dst, ma, kb * BKE, m0, fb);
          }
          if constexpr (BMN) {
#pragma unroll
            for (int c = 0; c < (BN / 2) / MNC; ++c) p_tma_load_2sm(dst + A_BYTES + c * (BKE * 128), mb, nb0 + MNC * c, kb * BKE, fb);
          } else {
            p_tma_load_2sm(dst + A_BYTES, mb, kb * BKE, nb0, fb);
          }
        }
      }
      // drain: every multicast commit aimed at this CTA's `empty` barriers has landed before the CTA may exit
      const uint32_t last = it;
      for (uint32_t j = (last > (uint32_t)STAGES ? last - STAGES : 0u); j < last; ++j)
        mbar_wait(&empty_bar[j % STAGES], (j / STAGES) & 1);
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {
      // -------------------------------------------------------------- MMA issuer (leader CTA, single thread)
      // instruction descriptor: D = f32, A = B = tf32, both K-major, N = BN, M = 256 (128 rows in each CTA's TMEM)
      constexpr uint32_t fmt = F16 ? 0u : 2u;            // kind::f16: 0 = F16 ; kind::tf32: 2 = TF32
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((AMN ? 1u : 0u) << 15) | ((BMN ? 1u : 0u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * P_BM) >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int tile = pair; tile < p.tiles_total; tile += p.num_pairs, ++tcount) {
        const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty[a], aph ^ 1);               // both CTAs' epilogues have drained this accumulator buffer
        p_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        const PairTile t = pair_decode(p, tile);
        const int kb0 = t.ks * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          p_fence_after();
          unsigned char* src = tiles + s * STAGE_BYTES;
          const uint64_t da = AMN ? (F16 ? p_desc_sw128_mn16(src, BKE) : p_desc_sw128_mn(src)) : p_desc_sw128(src);
          const uint64_t db = BMN ? (F16 ? p_desc_sw128_mn16(src + A_BYTES, BKE) : p_desc_sw128_mn(src + A_BYTES)) : p_desc_sw128(src + A_BYTES);
          // K-major: 8 tf32 / 16 fp16 = 32 B further along the swizzled row; MN-major: the next 8 (tf32) / 16 (fp16) k rows
          constexpr uint64_t ka = AMN ? (F16 ? 128 : 64) : 2, kbs = BMN ? (F16 ? 128 : 64) : 2;
#pragma unroll
          for (int k = 0; k < P_BK / 8; ++k) {             // 4 UMMAs of 32 bytes of K each (8 tf32 / 16 fp16)
            if constexpr (F16) p_umma_f16_pair(d_tmem, da + (uint64_t)k * ka, db + (uint64_t)k * kbs, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
            else p_umma_tf32_pair(d_tmem, da + (uint64_t)k * ka, db + (uint64_t)k * kbs, idesc, (kb > kb0 || k != 0) ? 1u : 0u);
          }
          p_commit_pair(&empty_bar[s]);                    // stage free in both CTAs once these MMAs retire
        }
        p_commit_pair(&tmem_full[a]);                      // accumulator complete: visible to both CTAs' epilogues
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warps 2 .. 2 + EW - 1 (both CTAs)
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int chalf = (warp - 2) >> 2;                     // which slice of the tile's columns
    constexpr int CW = BN / (EW / 4);                      // columns per epilogue warp
    float* stg = stage_all + (warp - 2) * 32 * P_EPI_LD;
    const int col4 = lane & 7, rsub = lane >> 3;           // read-back: 8 lanes x float4 per row, 4 rows per instruction
    const float alpha = p.alpha, beta = p.beta;
    const bool c_vec = ((p.ldc & 3) == 0) && (((reinterpret_cast<uintptr_t>(p.C[0]) | reinterpret_cast<uintptr_t>(p.C[1])) & 15) == 0) &&
                       ((p.split_stride & 3) == 0);
    const bool plain = c_vec && beta == 0.f && alpha == 1.f;
    // sigmoid-gate epilogue (DGAdaChannel): the gate operand f, the keep mask and the saved gate are moved as 128-bit / 32-bit
    // vectors, and f + mask of a whole 32 x 32 chunk are requested BEFORE the accumulator chunk is read from TMEM, so their DRAM
    // latency hides behind the TMEM read and the shared-memory transpose instead of stalling every row of the chunk
    bool gate_vec = false;
    if constexpr (EPI == DASA_EPI_GATE) {
      gate_vec = (p.ep.gate_src != nullptr) && ((reinterpret_cast<uintptr_t>(p.ep.gate_src) & 15) == 0) && ((p.ep.ld_gate & 3) == 0) &&
                 (p.ep.gate_out == nullptr || (((reinterpret_cast<uintptr_t>(p.ep.gate_out) & 15) == 0) && ((p.ep.ld_gate_out & 3) == 0))) &&
                 (p.ep.drop_mask == nullptr || (((reinterpret_cast<uintptr_t>(p.ep.drop_mask) & 3) == 0) && ((p.N & 3) == 0)));
    }
    const uint32_t empty_remote0 = p_mapa(smem_u32(&tmem_empty[0]), 0);
    const uint32_t empty_remote1 = p_mapa(smem_u32(&tmem_empty[1]), 0);
    uint32_t tcount = 0;
    for (int tile = pair; tile < p.tiles_total; tile += p.num_pairs, ++tcount) {
      const PairTile t = pair_decode(p, tile);
      const int m_base = t.mt * (2 * P_BM) + (int)crank * P_BM + q * 32;
      const int n_base = t.nt * BN + chalf * CW;
      float* const Cg = p.C[t.g] + (int64_t)t.ks * p.split_stride;
      const uint32_t a = tcount & 1, aph = (tcount >> 1) & 1;
      LstmLane lln;
      LstmBatch lb[2];
      if constexpr (EPI == P_EPI_LSTM) {                   // nothing here depends on the GEMM: in flight while the tensor cores work
        p_lstm_prefetch(ext.le, t.g, m_base, t.nt * 64 + chalf * 32);
        p_lstm_lane(lln, ext.le, t.g, t.nt * 64 + chalf * 32);
        p_lstm_load(lb[0], lln, m_base, 0);
      }
      mbar_wait(&tmem_full[a], aph);
      p_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + a * BN + chalf * CW;
      const int Mg = t.g ? p.M2 : p.M;
      const bool interior = plain && (m_base + 32 <= Mg) && (n_base + CW <= p.N);   // warp-uniform: no edge checks at all
      if constexpr (EPI == P_EPI_LSTM) {
        static_assert(EPI != P_EPI_LSTM || (BN == 256 && EW == 8), "the LSTM epilogue assumes 128 accumulator columns per epilogue warp");
        (void)interior; (void)Cg;
        static_assert(EPI != P_EPI_LSTM || 8 * 32 * P_LSTM_LD * 4 <= STAGES * STAGE_BYTES, "LSTM staging must fit the operand ring");
        // one tile per pair (host: num_pairs == tiles_total): once tmem_full has fired every operand stage is idle for good
        float* stg = reinterpret_cast<float*>(tiles) + (warp - 2) * 32 * P_LSTM_LD;
        p_lstm_stage(t_addr, stg, [&]() {
          p_fence_before();
          __syncwarp();
          if (lane == 0) p_remote_arrive(a ? empty_remote1 : empty_remote0);
        });
        constexpr int NBATCH = 32 / (4 * P_LSTM_NB);
#pragma unroll
        for (int bi = 0; bi < NBATCH; ++bi) {
          if (bi + 1 < NBATCH) p_lstm_load(lb[(bi + 1) & 1], lln, m_base, (bi + 1) * 4 * P_LSTM_NB);
          p_lstm_cell(lb[bi & 1], lln, m_base, bi * 4 * P_LSTM_NB, stg);
        }
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < CW; c0 += 32) {
        float4 gsrc[8];
        uint32_t gmask[8];
        const bool gate_fast = (EPI == DASA_EPI_GATE) && gate_vec && interior;
        if constexpr (EPI == DASA_EPI_GATE) {
          if (gate_fast) {
            const int ng = n_base + c0 + 4 * col4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int mg = m_base + 4 * i + rsub;
              gsrc[i] = __ldg(reinterpret_cast<const float4*>(p.ep.gate_src + (int64_t)mg * p.ep.ld_gate + ng));
              gmask[i] = (p.ep.drop_mask != nullptr) ? __ldg(reinterpret_cast<const uint32_t*>(p.ep.drop_mask + (int64_t)mg * p.N + ng))
                                                     : 0x01010101u;
            }
          }
        }
        uint32_t r[32];
        p_tmem_ld32(t_addr + (uint32_t)c0, r);
        if (c0 + 32 >= CW) {                               // last TMEM read of this buffer: hand it back to the MMA thread
          p_fence_before();
          __syncwarp();
          if (lane == 0) p_remote_arrive(a ? empty_remote1 : empty_remote0);
        }
        float* srow = stg + lane * P_EPI_LD;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + 4 * j) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int n = n_base + c0 + 4 * col4;
        if (interior) {
          float bias4[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (epi_has_bias<EPI>()) {
            if (p.ep.bias != nullptr) {
#pragma unroll
              for (int e = 0; e < 4; ++e) bias4[e] = __ldg(p.ep.bias + n + e);
            }
          }
          if (gate_fast) {
            if constexpr (EPI == DASA_EPI_GATE) {
              const float ds = (p.ep.drop_mask != nullptr) ? p.ep.drop_scale : 1.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = m_base + 4 * i + rsub;
                const float4 a4 = *reinterpret_cast<const float4*>(stg + (4 * i + rsub) * P_EPI_LD + 4 * col4);
                const float4 sg = make_float4(sigmoidf_(a4.x + bias4[0]), sigmoidf_(a4.y + bias4[1]), sigmoidf_(a4.z + bias4[2]),
                                              sigmoidf_(a4.w + bias4[3]));
                if (p.ep.gate_out != nullptr) *reinterpret_cast<float4*>(p.ep.gate_out + (int64_t)m * p.ep.ld_gate_out + n) = sg;
                const uint32_t mk = gmask[i];
                const float4 f4 = gsrc[i];
                *reinterpret_cast<float4*>(Cg + (int64_t)m * p.ldc + n) =
                    make_float4(sg.x * f4.x * ((mk & 0xFFu) ? ds : 0.f), sg.y * f4.y * ((mk & 0xFF00u) ? ds : 0.f),
                                sg.z * f4.z * ((mk & 0xFF0000u) ? ds : 0.f), sg.w * f4.w * ((mk & 0xFF000000u) ? ds : 0.f));
              }
            }
          } else {
#pragma unroll
          for (int rr = 0; rr < 32; rr += 4) {
            const int m = m_base + rr + rsub;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + (rr + rsub) * P_EPI_LD + 4 * col4);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = apply_activation_t<EPI>(v[e] + bias4[e], m, n + e, p.N, p.ep);
            p_store4(Cg, (int64_t)m * p.ldc, n, v, F16 && p.c_half);
          }
          }
        } else if (n < p.N) {
          const bool vec_ok = c_vec && (n + 3 < p.N);
          float bias4[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (epi_has_bias<EPI>()) {
            if (p.ep.bias != nullptr) {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (n + e < p.N) bias4[e] = __ldg(p.ep.bias + n + e);
            }
          }
#pragma unroll 2
          for (int rr = 0; rr < 32; rr += 4) {
            const int m = m_base + rr + rsub;
            if (m >= Mg) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + (rr + rsub) * P_EPI_LD + 4 * col4);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            float* crow = Cg + (int64_t)m * p.ldc + n;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (n + e < p.N) {
                float x = alpha * v[e];
                if (beta != 0.f) x += beta * crow[e];
                v[e] = apply_activation_t<EPI>(x + bias4[e], m, n + e, p.N, p.ep);
              }
            }
            if (F16 && p.c_half) {                        // fp16 output: N % 4 == 0 and aligned rows are launch requirements
              if (vec_ok) p_store4(Cg, (int64_t)m * p.ldc, n, v, true);
            } else if (vec_ok) *reinterpret_cast<float4*>(crow) = make_float4(v[0], v[1], v[2], v[3]);
            else
#pragma unroll
              for (int e = 0; e < 4; ++e) if (n + e < p.N) crow[e] = v[e];
          }
        }
        __syncwarp();                                      // staging rows are rewritten by the next chunk
      }
    }
  }

  p_fence_before();
  __syncthreads();
  p_cluster_sync();                                        // neither CTA frees TMEM / exits while the other may still signal it
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool pair_make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)P_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major operand stored [K rows][cols] (row stride ld): box = 32 columns (one 128-byte swizzle row) x P_BK k rows
bool pair_make_map_mn(CUtensorMap* map, const float* base, int64_t cols, int64_t K, int64_t ld) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)P_BK};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES, int EPI, bool AMN = false, bool BMN = false, bool F16 = false, int EW = P_EPI_WARPS>
int launch_pair_e(const CUtensorMap* ta, const CUtensorMap* tb, const PairParams& p, cudaStream_t st,
                  const PairExt<EPI>& ext = PairExt<EPI>{}) {
  constexpr size_t smem = (size_t)STAGES * (P_BM * P_BK * 4 + (BN / 2) * P_BK * 4) + EW * 32 * P_EPI_LD * 4 + 256 + 1024;
  static_assert(smem <= 232448, "exceeds the 227 KB of shared memory a CTA may opt into");
  auto kern = gemm_tf32_pair_kernel<BN, STAGES, EPI, AMN, BMN, F16, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { dasa_set_error("gemm_tf32_pair attr", e); return DASA_ERR_CUDA; }
    attr_set = true;
  }
  kern<<<dim3(2u * (unsigned)p.num_pairs), 64 + 32 * EW, smem, st>>>(ta[0], tb[0], ta[1], tb[1], p, ext);     // __cluster_dims__(2,1,1): one pair per TPC
  return dasa_check_launch("gemm_tf32_pair_kernel");
}

template <int BN, int STAGES>
int launch_pair(const CUtensorMap* ta, const CUtensorMap* tb, const PairParams& p, int epilogue, cudaStream_t st) {
  switch (epilogue) {
    case DASA_EPI_BIAS: return launch_pair_e<BN, STAGES, DASA_EPI_BIAS>(ta, tb, p, st);
    case DASA_EPI_BIAS_TANH: return launch_pair_e<BN, STAGES, DASA_EPI_BIAS_TANH>(ta, tb, p, st);
    case DASA_EPI_BIAS_GELU: return launch_pair_e<BN, STAGES, DASA_EPI_BIAS_GELU>(ta, tb, p, st);
    case DASA_EPI_BIAS_RELU: return launch_pair_e<BN, STAGES, DASA_EPI_BIAS_RELU>(ta, tb, p, st);
    case DASA_EPI_GATE: return launch_pair_e<BN, STAGES, DASA_EPI_GATE>(ta, tb, p, st);
    case DASA_EPI_TANH: return launch_pair_e<BN, STAGES, DASA_EPI_TANH>(ta, tb, p, st);
    default: return launch_pair_e<BN, STAGES, DASA_EPI_NONE>(ta, tb, p, st);
  }
}

}  // namespace

static int g_pair_mode = -1;          // -1: read DASA_TC_PAIR (default 1), 0: never, 1: when eligible, 2: always (tests)
extern "C" int dasa_debug_gemm_pair(int mode) {
  g_pair_mode = mode;
  return DASA_OK;
}

// Tile width for the pair kernel, or 0 when the single-CTA kernel should keep the problem. A pair processes 256 x BN tiles; the
// static schedule wants at least one full wave over the 74 TPCs and little tail quantisation.
int dasa_gemm_pair_plan(int M, int N, int K) {
  if (g_pair_mode < 0) { const char* e = getenv("DASA_TC_PAIR"); g_pair_mode = e ? atoi(e) : 1; }
  if (g_pair_mode == 0 || K < P_BK) return 0;
  if (g_pair_mode == 2) return 256;
  // measured on B200 (scripts/pair_gemm.py): 256-wide tiles beat both the 128-wide pair tiles and the single-CTA kernel on every
  // many-tile shape of the rollout (540-770 vs 320-600 TFLOP/s); problems that do not fill one wave of the 74 TPCs, or that
  // leave most of the last wave idle, stay on the single-CTA kernel (which can split K).
  const int pairs = DASA_NUM_SMS / 2;
  const int64_t t = dasa_cdiv(M, 2 * P_BM) * dasa_cdiv(N, 256);
  if (t < pairs) return (2 * t >= pairs && K >= 2048) ? 256 : 0;     // long-K weight-gradient shapes: one partial wave still wins
  const double eff = (double)t / (double)(dasa_cdiv(t, pairs) * pairs);
  return eff >= 0.6 ? 256 : 0;
}

int dasa_gemm_tc_pair(int bn, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta,
                      float* C, int64_t ldc, int epilogue, const EpiParams& ep, cudaStream_t st) {
  if (bn != 256) return DASA_ERR_UNSUPPORTED;
  PairParams p{};
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C[0] = C; p.C[1] = C; p.ldc = ldc; p.ep = ep;
  p.tiles_n = (int)dasa_cdiv(N, bn);
  p.tiles_mn = (int)(dasa_cdiv(M, 2 * P_BM) * p.tiles_n);
  p.splits = 1; p.kb_per_split = (int)dasa_cdiv(K, P_BK); p.split_stride = 0;
  p.tiles_total = p.tiles_mn;
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  if (!pair_make_map(&ta[0], A, M, K, lda, P_BM) || !pair_make_map(&tb[0], B, N, K, ldb, bn / 2)) return DASA_ERR_UNSUPPORTED;
  ta[1] = ta[0]; tb[1] = tb[0];
  return launch_pair<256, 5>(ta, tb, p, epilogue, st);
}

// fp16 operands (both K-major): C = epilogue(A B^T + bias), C fp32 or fp16. Forward-only GEMMs of the frozen transformer stack.
namespace {
bool pair_make_map_f16(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)(2 * P_BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

bool dasa_gemm_f16_pair_supported(int M, int N, int K) {
  return M > 0 && N > 0 && K >= 2 * P_BK && (K % 8) == 0 && dasa_tensormap_encoder() != nullptr && dasa_gemm_pair_plan(M, N, K) == 256;
}

int dasa_gemm_tc_pair_f16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc,
                          int c_half, int epilogue, const EpiParams& ep, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K < 2 * P_BK) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(A) || !dasa_aligned16(B) || !dasa_aligned16(C) || (lda & 7) || (ldb & 7) || (ldc & 3) || (c_half && (N & 3)))
    return DASA_ERR_BAD_ALIGN;
  ++g_gemm_routes[DASA_ROUTE_PAIR_F16];
  PairParams p{};
  p.M = M; p.N = N; p.K = K; p.alpha = 1.f; p.beta = 0.f; p.C[0] = p.C[1] = static_cast<float*>(C); p.ldc = ldc; p.ep = ep;
  p.c_half = c_half ? 1 : 0;
  p.tiles_n = (int)dasa_cdiv(N, 256);
  p.tiles_mn = (int)(dasa_cdiv(M, 2 * P_BM) * p.tiles_n);
  p.splits = 1; p.kb_per_split = (int)dasa_cdiv(K, 2 * P_BK); p.split_stride = 0;
  p.tiles_total = p.tiles_mn;
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  if (!pair_make_map_f16(&ta[0], A, M, K, lda, P_BM) || !pair_make_map_f16(&tb[0], B, N, K, ldb, 128)) return DASA_ERR_UNSUPPORTED;
  ta[1] = ta[0]; tb[1] = tb[0];
  switch (epilogue) {
    case DASA_EPI_NONE: return launch_pair_e<256, 5, DASA_EPI_NONE, false, false, true>(ta, tb, p, st);
    case DASA_EPI_BIAS: return launch_pair_e<256, 5, DASA_EPI_BIAS, false, false, true>(ta, tb, p, st);
    case DASA_EPI_BIAS_GELU: return launch_pair_e<256, 4, DASA_EPI_BIAS_GELU, false, false, true, 16>(ta, tb, p, st);
    case DASA_EPI_GATE: return c_half ? DASA_ERR_UNSUPPORTED : launch_pair_e<256, 5, DASA_EPI_GATE, false, false, true>(ta, tb, p, st);
    default: return DASA_ERR_UNSUPPORTED;
  }
}

// Operand layouts other than (K-major, K-major): C = alpha * op(A) * op(B) + beta * C with an MN-major A ([K][M] in memory)
// and / or B ([K][N]); no fused epilogue (the backward GEMMs dX = dY.W and dW += dY^T.X are the users).
bool dasa_gemm_pair_mn_supported(int a_kmajor, int b_kmajor, int M, int N, int K, const float* A, int64_t lda, const float* B,
                                 int64_t ldb, int epilogue) {
  if (a_kmajor && b_kmajor) return false;
  if (epilogue != DASA_EPI_NONE || M <= 0 || N <= 0 || K < P_BK) return false;
  if (!dasa_aligned16(A) || !dasa_aligned16(B) || (lda & 3) || (ldb & 3)) return false;
  if (dasa_tensormap_encoder() == nullptr) return false;
  return dasa_gemm_pair_plan(M, N, K) == 256;
}

// Few-tile, long-K problems (the weight gradients: 4096 x 768 outputs reduced over 35 700 rows = 48 tiles for 74 CTA pairs): K is
// split so that the tiles fill whole waves; partial sums go to the workspace and one pass folds them into C.
int dasa_gemm_pair_mn_splits(int M, int N, int K) {
  const int pairs = DASA_NUM_SMS / 2;
  const int64_t tiles = dasa_cdiv(M, 2 * P_BM) * dasa_cdiv(N, 256);
  const int64_t nkb = dasa_cdiv(K, P_BK);
  if (tiles >= pairs || nkb < 64) return 1;
  int best = 1;
  double best_t = (double)dasa_cdiv(tiles, pairs);
  for (int s = 2; s <= 4; ++s) {
    const double t = (double)dasa_cdiv(tiles * s, pairs) / s;
    if (t < best_t - 0.15) { best_t = t; best = s; }
  }
  return best;
}

namespace {
__global__ void __launch_bounds__(256) pair_split_reduce_kernel(const float* __restrict__ part, int splits, int64_t stride, float alpha,
                                                                float beta, float* __restrict__ C, int64_t ldc, int M, int N) {
  const int n4 = N >> 2;
  const int64_t total = (int64_t)M * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / n4), n = (int)(i % n4) << 2;
    float4 acc = *reinterpret_cast<const float4*>(part + (int64_t)m * N + n);
    for (int sidx = 1; sidx < splits; ++sidx) {                 // fixed order: deterministic
      const float4 v = *reinterpret_cast<const float4*>(part + sidx * stride + (int64_t)m * N + n);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float* c = C + (int64_t)m * ldc + n;
    if (beta != 0.f) {
      acc.x = alpha * acc.x + beta * c[0]; acc.y = alpha * acc.y + beta * c[1];
      acc.z = alpha * acc.z + beta * c[2]; acc.w = alpha * acc.w + beta * c[3];
    } else {
      acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
    }
    c[0] = acc.x; c[1] = acc.y; c[2] = acc.z; c[3] = acc.w;
  }
}
}  // namespace

int dasa_gemm_tc_pair_mn(int a_kmajor, int b_kmajor, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                         int64_t ldb, float beta, float* C, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  PairParams p{};
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C[0] = C; p.C[1] = C; p.ldc = ldc;
  p.tiles_n = (int)dasa_cdiv(N, 256);
  p.tiles_mn = (int)(dasa_cdiv(M, 2 * P_BM) * p.tiles_n);
  p.splits = 1; p.kb_per_split = (int)dasa_cdiv(K, P_BK); p.split_stride = 0;
  p.tiles_total = p.tiles_mn;
  int splits = dasa_gemm_pair_mn_splits(M, N, K);
  if ((N & 3) != 0 || workspace == nullptr || workspace_bytes < (size_t)splits * M * N * sizeof(float) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    splits = 1;
  if (splits > 1) {
    const int nkb = (int)dasa_cdiv(K, P_BK);
    p.kb_per_split = (int)dasa_cdiv(nkb, splits);
    p.splits = (int)dasa_cdiv(nkb, p.kb_per_split);
    p.split_stride = (int64_t)M * N;
    p.C[0] = p.C[1] = static_cast<float*>(workspace);
    p.ldc = N; p.alpha = 1.f; p.beta = 0.f;
    p.tiles_total = p.tiles_mn * p.splits;
  }
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  const bool oka = a_kmajor ? pair_make_map(&ta[0], A, M, K, lda, P_BM) : pair_make_map_mn(&ta[0], A, M, K, lda);
  const bool okb = b_kmajor ? pair_make_map(&tb[0], B, N, K, ldb, 128) : pair_make_map_mn(&tb[0], B, N, K, ldb);
  if (!oka || !okb) return DASA_ERR_UNSUPPORTED;
  ta[1] = ta[0]; tb[1] = tb[0];
  int rc;
  if (a_kmajor) rc = launch_pair_e<256, 5, DASA_EPI_NONE, false, true>(ta, tb, p, st);
  else if (b_kmajor) rc = launch_pair_e<256, 5, DASA_EPI_NONE, true, false>(ta, tb, p, st);
  else rc = launch_pair_e<256, 5, DASA_EPI_NONE, true, true>(ta, tb, p, st);
  if (rc != DASA_OK || p.splits <= 1) return rc;
  const int64_t work = (int64_t)M * (N >> 2);
  const unsigned grid = (unsigned)(dasa_cdiv(work, 256) < (int64_t)DASA_NUM_SMS * 8 ? dasa_cdiv(work, 256) : (int64_t)DASA_NUM_SMS * 8);
  pair_split_reduce_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(workspace), p.splits, p.split_stride, alpha, beta, C, ldc, M, N);
  return dasa_check_launch("pair_split_reduce_kernel");
}

// Two independent problems of the same shape in ONE launch, optionally split along K (raw partial sums, no epilogue):
//   P[g][s] (M x N, row stride ldc, at C[g] + s * split_stride) = A[g][:, Ks] * B[g][:, Ks]^T.
// The recurrent GEMMs of the two bi-LSTM directions are issued this way, so one time step is one launch that fills the TPCs.
int dasa_gemm_tc_pair_grouped(int M, int N, int K, const float* const A[2], int64_t lda, const float* const B[2], int64_t ldb,
                              float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st) {
  return dasa_gemm_tc_pair_grouped2(M, M, N, K, A, lda, B, ldb, C, ldc, splits, split_stride, st);
}

// The same with a row count per group (M0 rows of A[0] / C[0], M1 rows of A[1] / C[1]; either may be 0): the padding-free
// bi-LSTM recurrence, where the forward direction's live sequences shrink while the reverse direction's grow.
int dasa_gemm_tc_pair_grouped2(int M0, int M1, int N, int K, const float* const A[2], int64_t lda, const float* const B[2], int64_t ldb,
                               float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st) {
  if (M0 < 0 || M1 < 0 || M0 + M1 <= 0 || N <= 0 || K < P_BK || splits < 1) return DASA_ERR_BAD_SHAPE;
  ++g_gemm_routes[DASA_ROUTE_PAIR_GROUPED];
  PairParams p{};
  p.M = M0; p.M2 = M1; p.N = N; p.K = K; p.alpha = 1.f; p.beta = 0.f; p.C[0] = C[0]; p.C[1] = C[1]; p.ldc = ldc;
  p.tiles_n = (int)dasa_cdiv(N, 256);
  p.tiles_mn = (int)(dasa_cdiv(M0, 2 * P_BM) * p.tiles_n);
  p.tiles_mn2 = (int)(dasa_cdiv(M1, 2 * P_BM) * p.tiles_n);
  const int nkb = (int)dasa_cdiv(K, P_BK);
  if (splits > nkb) splits = nkb;
  p.kb_per_split = (int)dasa_cdiv(nkb, splits);
  p.splits = (int)dasa_cdiv(nkb, p.kb_per_split);          // every split owns at least one k-block
  p.split_stride = split_stride;
  p.tiles_total = (p.tiles_mn + p.tiles_mn2) * p.splits;
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  for (int g = 0; g < 2; ++g) {
    const int Mg = g ? M1 : M0;
    if (Mg == 0) continue;
    if (!dasa_aligned16(A[g]) || !dasa_aligned16(B[g]) || (lda & 3) || (ldb & 3)) return DASA_ERR_BAD_ALIGN;
    if (!pair_make_map(&ta[g], A[g], Mg, K, lda, P_BM) || !pair_make_map(&tb[g], B[g], N, K, ldb, 128)) return DASA_ERR_UNSUPPORTED;
  }
  if (M0 == 0) { ta[0] = ta[1]; tb[0] = tb[1]; }            // an absent group has no tiles: its maps are never fetched
  if (M1 == 0) { ta[1] = ta[0]; tb[1] = tb[0]; }
  return launch_pair_e<256, 5, DASA_EPI_NONE>(ta, tb, p, st) == DASA_OK ? p.splits : DASA_ERR_CUDA;
}

// dW-shaped products on fp16 operands: C[M, N] = alpha * A^T B + beta * C with A stored [K][M] and B stored [K][N] as IEEE fp16
// (both MN-major: the row index of both arrays is the reduction index), tcgen05 kind::f16, fp32 accumulate, K split like the
// TF32 MN-major path. The bi-LSTM weight gradients dW_ih = dgates^T x, dW_hh = dgates^T h_prev use it with the fp16 copies the
// recurrence already keeps (dgates * 2^8, the state rows): twice the tensor rate and half the operand bytes of the TF32 form.
namespace {
bool pair_make_map_mn16(CUtensorMap* map, const void* base, int64_t cols, int64_t K, int64_t ld) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(dasa_tensormap_encoder());
  if (enc == nullptr) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)(2 * P_BK)};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

int dasa_gemm_tc_pair_mn_f16(int M, int N, int K, float alpha, const void* A, int64_t lda, const void* B, int64_t ldb, float beta,
                             float* C, int64_t ldc, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K < 2 * P_BK) return DASA_ERR_BAD_SHAPE;
  if (!dasa_aligned16(A) || !dasa_aligned16(B) || (lda & 7) || (ldb & 7)) return DASA_ERR_BAD_ALIGN;
  ++g_gemm_routes[DASA_ROUTE_PAIR_F16];
  PairParams p{};
  p.M = M; p.N = N; p.K = K; p.alpha = alpha; p.beta = beta; p.C[0] = C; p.C[1] = C; p.ldc = ldc;
  p.tiles_n = (int)dasa_cdiv(N, 256);
  p.tiles_mn = (int)(dasa_cdiv(M, 2 * P_BM) * p.tiles_n);
  const int nkb = (int)dasa_cdiv(K, 2 * P_BK);
  p.splits = 1; p.kb_per_split = nkb; p.split_stride = 0;
  p.tiles_total = p.tiles_mn;
  int splits = dasa_gemm_pair_mn_splits(M, N, K);
  if ((N & 3) != 0 || workspace == nullptr || workspace_bytes < (size_t)splits * M * N * sizeof(float) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    splits = 1;
  if (splits > 1) {
    p.kb_per_split = (int)dasa_cdiv(nkb, splits);
    p.splits = (int)dasa_cdiv(nkb, p.kb_per_split);
    p.split_stride = (int64_t)M * N;
    p.C[0] = p.C[1] = static_cast<float*>(workspace);
    p.ldc = N; p.alpha = 1.f; p.beta = 0.f;
    p.tiles_total = p.tiles_mn * p.splits;
  }
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  if (!pair_make_map_mn16(&ta[0], A, M, K, lda) || !pair_make_map_mn16(&tb[0], B, N, K, ldb)) return DASA_ERR_UNSUPPORTED;
  ta[1] = ta[0]; tb[1] = tb[0];
  int rc = launch_pair_e<256, 5, DASA_EPI_NONE, true, true, true>(ta, tb, p, st);
  if (rc != DASA_OK || p.splits <= 1) return rc;
  const int64_t work = (int64_t)M * (N >> 2);
  const unsigned grid = (unsigned)(dasa_cdiv(work, 256) < (int64_t)DASA_NUM_SMS * 8 ? dasa_cdiv(work, 256) : (int64_t)DASA_NUM_SMS * 8);
  pair_split_reduce_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(workspace), p.splits, p.split_stride, alpha, beta, C, ldc, M, N);
  return dasa_check_launch("pair_split_reduce_kernel");
}

// The grouped / split-K launch with fp16 operands (tcgen05 kind::f16, fp32 partial sums): the backward recurrence of the packed
// bi-LSTM, dh = dgates W_hh over scaled fp16 copies of dgates (bilstm_packed.cu). K % 8 == 0, lda / ldb in halves.
int dasa_gemm_tc_pair_grouped2_f16(int M0, int M1, int N, int K, const __half* const A[2], int64_t lda, const __half* const B[2],
                                   int64_t ldb, float* const C[2], int64_t ldc, int splits, int64_t split_stride, cudaStream_t st) {
  if (M0 < 0 || M1 < 0 || M0 + M1 <= 0 || N <= 0 || K < 2 * P_BK || (K & 7) || splits < 1) return DASA_ERR_BAD_SHAPE;
  if ((lda & 7) || (ldb & 7)) return DASA_ERR_BAD_ALIGN;
  ++g_gemm_routes[DASA_ROUTE_PAIR_GROUPED];
  PairParams p{};
  p.M = M0; p.M2 = M1; p.N = N; p.K = K; p.alpha = 1.f; p.beta = 0.f; p.C[0] = C[0]; p.C[1] = C[1]; p.ldc = ldc;
  p.tiles_n = (int)dasa_cdiv(N, 256);
  p.tiles_mn = (int)(dasa_cdiv(M0, 2 * P_BM) * p.tiles_n);
  p.tiles_mn2 = (int)(dasa_cdiv(M1, 2 * P_BM) * p.tiles_n);
  const int nkb = (int)dasa_cdiv(K, 2 * P_BK);
  if (splits > nkb) splits = nkb;
  p.kb_per_split = (int)dasa_cdiv(nkb, splits);
  p.splits = (int)dasa_cdiv(nkb, p.kb_per_split);
  p.split_stride = split_stride;
  p.tiles_total = (p.tiles_mn + p.tiles_mn2) * p.splits;
  p.num_pairs = p.tiles_total < DASA_NUM_SMS / 2 ? p.tiles_total : DASA_NUM_SMS / 2;
  CUtensorMap ta[2], tb[2];
  for (int g = 0; g < 2; ++g) {
    const int Mg = g ? M1 : M0;
    if (Mg == 0) continue;
    if (!dasa_aligned16(A[g]) || !dasa_aligned16(B[g])) return DASA_ERR_BAD_ALIGN;
    if (!pair_make_map_f16(&ta[g], A[g], Mg, K, lda, P_BM) || !pair_make_map_f16(&tb[g], B[g], N, K, ldb, 128)) return DASA_ERR_UNSUPPORTED;
  }
  if (M0 == 0) { ta[0] = ta[1]; tb[0] = tb[1]; }
  if (M1 == 0) { ta[1] = ta[0]; tb[1] = tb[0]; }
  return launch_pair_e<256, 5, DASA_EPI_NONE, false, false, true>(ta, tb, p, st) == DASA_OK ? p.splits : DASA_ERR_CUDA;
}

// One recurrence step of the packed bi-LSTM, both directions, GEMM + cell update in ONE launch (lstm_epi.cuh). Tiles cover
// max(n, n_next) rows per direction (rows that join at the next step get their zero state from the epilogue); the operand maps
// cover the n live rows only, so the rest of a tile's A rows are zero-filled by TMA.
int dasa_gemm_tc_pair_lstm(const __half* const A16[2], const __half* const W16[2], const LstmEpi& le, cudaStream_t st) {
  const int H = le.H, N = 4 * le.H, K = le.H;
  if (H < 64 || (H % 64) != 0) return DASA_ERR_UNSUPPORTED;
  int cover[2];
  for (int g = 0; g < 2; ++g) cover[g] = le.n[g] > le.n_next[g] ? le.n[g] : le.n_next[g];
  if (le.n[0] < 0 || le.n[1] < 0 || cover[0] + cover[1] <= 0) return DASA_ERR_BAD_SHAPE;
  ++g_gemm_routes[DASA_ROUTE_PAIR_GROUPED];
  PairParams p{};
  p.M = cover[0]; p.M2 = cover[1]; p.N = N; p.K = K; p.alpha = 1.f; p.beta = 0.f; p.ldc = N;
  p.tiles_n = N / 256;
  p.tiles_mn = (int)(dasa_cdiv(cover[0], 2 * P_BM) * p.tiles_n);
  p.tiles_mn2 = (int)(dasa_cdiv(cover[1], 2 * P_BM) * p.tiles_n);
  p.splits = 1; p.kb_per_split = K / (2 * P_BK); p.split_stride = 0;
  p.tiles_total = p.tiles_mn + p.tiles_mn2;
  p.num_pairs = p.tiles_total;                                // ONE tile per CTA pair: the epilogue stages through the idle operand ring
  CUtensorMap ta[2], tb[2];
  bool have[2] = {false, false};
  for (int g = 0; g < 2; ++g) {
    if (cover[g] == 0) continue;
    if (!dasa_aligned16(A16[g]) || !dasa_aligned16(W16[g])) return DASA_ERR_BAD_ALIGN;
    // a direction with tiles but no live row (cannot happen for a valid plan) would need a 0-row map: refuse it
    if (le.n[g] <= 0) return DASA_ERR_BAD_SHAPE;
    if (!pair_make_map_f16(&ta[g], A16[g], le.n[g], K, K, P_BM) || !pair_make_map_f16(&tb[g], W16[g], N, K, K, 128)) return DASA_ERR_UNSUPPORTED;
    have[g] = true;
  }
  if (!have[0]) { ta[0] = ta[1]; tb[0] = tb[1]; }            // an absent group has no tiles: its maps are never fetched
  if (!have[1]) { ta[1] = ta[0]; tb[1] = tb[0]; }
  PairExt<P_EPI_LSTM> ext;
  ext.le = le;
  return launch_pair_e<256, 5, P_EPI_LSTM, false, false, true>(ta, tb, p, st, ext);
}

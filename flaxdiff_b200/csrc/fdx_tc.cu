// fdx_tc.cu -- the tcgen05 / TMA / TMEM "tap-GEMM" engine for sm_100a.
//
// Replaces what the reference delegates to XLA for every dense contraction on the
// UNet hot path: flax nn.Conv 3x3 / 1x1 (flaxdiff/models/common.py:166-172,
// 237-244, 324-333), nn.DenseGeneral projections (models/attention.py:132-154)
// and the QK^T / PV products of nn.dot_product_attention (attention.py:170-174),
// forward and backward.
//
// Design (one CTA per SM, persistent over output tiles, 6 warps):
//   warp 0      TMA producer: per K-chunk one 128B-swizzled A box (64 channels x
//               128 pixels, shifted by the tap offset; out-of-image = zero fill
//               = SAME padding) and the matching B box, into a ring of stages.
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN,
//               K=16, bf16 -> f32) into a double-buffered TMEM accumulator and
//               tcgen05.commit's the stage back to the producer.
//   warps 2..5  epilogue: tcgen05.ld the accumulator (one pixel row per thread),
//               fuse bias + per-image timestep vector + residual, store bf16/f32
//               (or f32 atomics for the split-K weight gradient).
#include "fdx_tc.cuh"
#include "fdx_epilogue.cuh"
#include <stdlib.h>

namespace {

constexpr int kThreads = 192;
constexpr int kBK = 64;                 // K elements (or K rows) per pipeline stage
constexpr int kABytes = 128 * kBK * 2;  // 16 KB per stage for A in every mode

template <int BN>
struct TcCfg {
  // K chunks (64 deep) per pipeline stage.  At BN = 64 one chunk is only 4 MMAs = 128 tensor
  // cycles, less than the ~150-200 cycles the single producer / issuer threads spend per stage
  // on mbarrier try_wait + TMA / MMA issue (ncu: tensor pipe 22 % active), so two chunks share
  // a stage there.  Not used by the MN-major-A (weight-gradient) mode.
  static constexpr int kSub = (BN == 64) ? 2 : 1;
  static constexpr int kStages = (BN == 64) ? 4 : (BN == 128) ? 6 : 4;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kSub * (kABytes + kBBytes);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                    16384 /*epilogue staging: 4 warps x 4 KB*/ +
                                    4096 /*GroupNorm-backward column-sum exchange*/;
  // Every M=128,K=16 MMA with a K-major A operand costs ~150 cycles whatever N is (tensor pipe 21 %
  // active at N=64, 43 % at 128, 85 % at 256).  Cycling through kParts independent partial accumulators
  // (summed by the epilogue) was measured NOT to help (4 / 2 parts at BN = 64 / 128: 26.4 vs 25.9 ms per
  // step), so the cost is not an accumulator dependency chain; neither is it B's shared-memory traffic
  // (CTA pairs, below).  The mechanism is kept for experiments.
  static constexpr int kParts = 1;
  static constexpr int kTmemCols = 512;     // 2 buffers x kParts x BN columns
};

struct TcDev {
  int mode;
  int TW, TH, TN;        // pixel box of one A tile (128 pixels for K-major A, 64 for MN-major A)
  int nxb, nyb, nnb;     // pixel blocks along x, y, n
  int W, H, N;
  int es;
  int ntaps;
  int tap_dx[kTcMaxTaps], tap_dy[kTcMaxTaps], tap_b[kTcMaxTaps];
  int es_b, tap_bdx[kTcMaxTaps], tap_bdy[kTcMaxTaps];   // TC_MNMN strided / shifted B view
  int kchunks;           // K chunks per tap (TC_KK/TC_KMN)
  int K;
  int M;                 // TC_MNMN valid rows
  int mblks;             // TC_MNMN: row blocks of 128
  int Ncols, nblks;      // column blocks of BN
  int b_batched;
  int mn_batched;        // TC_MNMN: reduce over x blocks only, (y, n) blocks are batch
  int tpp;               // TC_MNMN: taps per tile (2 when M <= 64: two taps share one M=128 MMA)
  int ntp;               // TC_MNMN: number of tap groups = ceil(ntaps / tpp)
  int splits;
  int ntiles;
  int npairs;            // PAIR kernels: ceil(row tiles / 2) * nblks tile pairs
  // epilogue
  void* out;
  int out_f32, out_atomic;
  long long os_x, os_y, os_n, os_tap, os_m;
  float alpha;
  const float* bias;
  const float* rowvec;
  const void* res;
  long long rs_x, rs_y, rs_n;
  // fused GroupNorm-backward first pass (GN kernels only; see fdx_epilogue.cuh)
  const float* gn_ab;
  float* gn_ws;
  int gn_slots;
  int ws_ld;               // row length of gn_ws (EPI_COLSTATS: channels of the destination buffer)
};

// TC_MNMN tile decode: tile -> (batch block bz, split, nt, mb, tap) and the K range [pb0, pb1)
struct MnTile {
  int bz, sp, nt, mb, t, pb0, pb1;
};
__device__ __forceinline__ MnTile decode_mn(const TcDev& p, int tile) {
  // fastest -> slowest: tap group, column block, row block, batch block, K split.  CTAs that run
  // concurrently therefore work on the same pixel range and share A / B boxes in L2.
  MnTile m;
  int r = tile;
  const int nbz = p.mn_batched ? p.nyb * p.nnb : 1;
  m.t = r % p.ntp;     r /= p.ntp;
  m.nt = r % p.nblks;  r /= p.nblks;
  m.mb = r % p.mblks;  r /= p.mblks;
  m.bz = r % nbz;      r /= nbz;
  m.sp = r;
  const long long kblocks = p.mn_batched ? p.nxb : (long long)p.nxb * p.nyb * p.nnb;
  m.pb0 = (int)((kblocks * m.sp) / p.splits);
  m.pb1 = (int)((kblocks * (m.sp + 1)) / p.splits);
  return m;
}

// PAIR = true: the kernel runs as clusters of two CTAs (one TPC).  Each CTA owns one 128-row tile and
// stages its own A box plus HALF of the B tile's columns; the leader CTA (cluster rank 0) issues
// tcgen05.mma.cta_group::2 (M = 256) that reads both CTAs' shared memory and writes both CTAs' TMEM.
// B's TMA writes and UMMA reads per CTA are halved - the engine's measured bound is shared-memory
// bandwidth.  Barrier topology: every TMA of either CTA completes on the LEADER's full[stage]; the
// leader's commits multicast to empty[stage] / tfull[acc] of both CTAs; both CTAs' epilogue warps
// arrive on the leader's tempty[acc].
template <int BN, int MODE, int EPI, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
fdx_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              const TcDev p) {
  using Cfg = TcCfg<BN>;
  static_assert(!PAIR || MODE != TC_MNMN, "CTA pairs: K-major A only");
  constexpr int BNL = PAIR ? BN / 2 : BN;           // B columns staged by THIS CTA
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  // tile loop of this CTA: index t -> (column block nt, row tile mt)
  const int t_begin = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int t_step = PAIR ? (int)cluster_nctaid_x() : (int)gridDim.x;
  const int t_end = PAIR ? p.npairs : p.ntiles;
  constexpr int S = Cfg::kStages;
  constexpr bool A_MN = (MODE == TC_MNMN);
  constexpr bool B_MN = (MODE != TC_KK);
  constexpr int SUB = A_MN ? 1 : Cfg::kSub;    // K chunks per stage (producer view)
  constexpr int SUBK = SUB;                     // same, issuer view

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* full = bars;            // [S]
  uint64_t* empty = bars + S;       // [S]
  uint64_t* tfull = bars + 2 * S;   // [2]
  uint64_t* tempty = bars + 2 * S + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  uint8_t* epi_stage = smem + S * Cfg::kStageBytes + 256;   // 16 KB, 128-byte aligned
  float* epi_xchg = reinterpret_cast<float*>(epi_stage + 16384);   // 4 KB (GN only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&tfull[0], 1);
    mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], PAIR ? 8 : 4);
    mbar_init(&tempty[1], PAIR ? 8 : 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- tile decode helpers ------------------------------------------------
  // TC_KK / TC_KMN : tile -> (nt, xb, yb, nb);  nk = ntaps * kchunks
  // TC_MNMN        : tile -> (split, nt, mb, tap); nk = pixel blocks of this split

  if (warp == 0) {
    // =========================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        if constexpr (!A_MN) {
          const int nt = tile % p.nblks;
          const int mt = PAIR ? 2 * (tile / p.nblks) + (int)rank : tile / p.nblks;
          const int xb = mt % p.nxb;
          const int yb = (mt / p.nxb) % p.nyb;
          const int nb = mt / (p.nxb * p.nyb);
          const int x0 = xb * p.TW * p.es, y0 = yb * p.TH * p.es, n0 = nb * p.TN;
          const int zb1 = p.b_batched ? yb : 0, zb2 = p.b_batched ? nb : 0;
          const int nk = p.ntaps * p.kchunks;
          for (int q = 0; q < nk; q += SUB) {
            const int nsub = (nk - q) < SUB ? (nk - q) : SUB;
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa0 = smem + stage * Cfg::kStageBytes;
            uint8_t* sb0 = sa0 + SUB * kABytes;
            if constexpr (PAIR) {
              // the leader posts the bytes of BOTH CTAs; the peer's loads complete on the same barrier
              if (leader) mbar_arrive_expect_tx(&full[stage], nsub * (2 * kABytes + Cfg::kBBytes));
            } else {
              mbar_arrive_expect_tx(&full[stage], nsub * (kABytes + Cfg::kBBytes));
            }
            for (int u = 0; u < nsub; ++u) {
              const int t = (q + u) / p.kchunks, kc = (q + u) % p.kchunks;
              uint8_t* sa = sa0 + u * kABytes;
              uint8_t* sb = sb0 + u * Cfg::kBBytes;
              const int ncol0 = nt * BN + (int)rank * BNL;      // first B column staged by this CTA
              if constexpr (PAIR) {
                tma_load_4d_2sm(sa, &mapA, &full[stage], kc * kBK, x0 + p.tap_dx[t], y0 + p.tap_dy[t], n0);
                if constexpr (!B_MN) {
                  tma_load_4d_2sm(sb, &mapB, &full[stage], kc * kBK, ncol0, p.b_batched ? zb1 : p.tap_b[t], zb2);
                } else {
#pragma unroll
                  for (int j = 0; j < BNL / 64; ++j)
                    tma_load_4d_2sm(sb + j * (kBK * 128), &mapB, &full[stage], ncol0 + j * 64,
                                    p.tap_b[t] + kc * kBK, zb1, zb2);
                }
              } else {
              tma_load_4d(sa, &mapA, &full[stage], kc * kBK, x0 + p.tap_dx[t], y0 + p.tap_dy[t], n0);
              if constexpr (!B_MN) {
                // B K-major: box (64 k, BN rows); tap selects z1 (or batched z1/z2)
                tma_load_4d(sb, &mapB, &full[stage], kc * kBK, nt * BN,
                            p.b_batched ? zb1 : p.tap_b[t], zb2);
              } else {
                // B MN-major: BN/64 boxes of (64 n, 64 k rows)
#pragma unroll
                for (int j = 0; j < BN / 64; ++j)
                  tma_load_4d(sb + j * (kBK * 128), &mapB, &full[stage], nt * BN + j * 64,
                              p.tap_b[t] + kc * kBK, zb1, zb2);
              }
              }
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        } else {
          const MnTile mt_ = decode_mn(p, tile);
          const int nt = mt_.nt, mb = mt_.mb, t = mt_.t;
          for (int pb = mt_.pb0; pb < mt_.pb1; ++pb) {
            int xb, yb, nb;
            if (p.mn_batched) {
              xb = pb; yb = mt_.bz % p.nyb; nb = mt_.bz / p.nyb;
            } else {
              xb = pb % p.nxb; yb = (pb / p.nxb) % p.nyb; nb = pb / (p.nxb * p.nyb);
            }
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full[stage], kABytes + Cfg::kBBytes);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              // tpp == 1: the two 64-channel halves of row block mb for tap t
              // tpp == 2: channels [0,64) of taps 2t and 2t+1 (a missing tap reads out of range = 0)
              const int tap = (p.tpp == 2) ? (2 * t + j) : t;
              const bool live = tap < p.ntaps;
              const int tt = live ? tap : 0;
              const int ch = (p.tpp == 2) ? (live ? 0 : (int)p.M + 64) : (mb * 128 + j * 64);
              tma_load_4d(sa + j * (kBK * 128), &mapA, &full[stage], ch,
                          xb * p.TW * p.es + p.tap_dx[tt], yb * p.TH * p.es + p.tap_dy[tt], nb * p.TN);
            }
            const int tb = (p.tpp == 2) ? 0 : t;    // per-tap B views only without tap pairing
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * (kBK * 128), &mapB, &full[stage], nt * BN + j * 64,
                          xb * p.TW * p.es_b + p.tap_bdx[tb], yb * p.TH * p.es_b + p.tap_bdy[tb], nb * p.TN);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ===============================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        int nk;
        if constexpr (!A_MN) {
          nk = p.ntaps * p.kchunks;
        } else {
          const MnTile mt_ = decode_mn(p, tile);
          nk = mt_.pb1 - mt_.pb0;
        }
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * Cfg::kParts * BN);
        for (int q = 0; q < nk; q += SUBK) {
          const int nsub = (nk - q) < SUBK ? (nk - q) : SUBK;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa0 = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb0 = sa0 + SUBK * kABytes;
          for (int u = 0; u < nsub; ++u) {
            const uint32_t sa = sa0 + u * kABytes;
            const uint32_t sb = sb0 + u * Cfg::kBBytes;
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // K-major: advance 16 elements (32 B) inside the 128B swizzle row; SBO = 8 rows.
              // MN-major: advance 16 k-rows (2048 B); LBO = next 64-wide block, SBO = 8 rows.
              const uint64_t da = A_MN ? umma_desc_sw128(sa + k * 2048, kBK * 128, 1024)
                                       : umma_desc_sw128(sa + k * 32, 16, 1024);
              const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, kBK * 128, 1024)
                                       : umma_desc_sw128(sb + k * 32, 16, 1024);
              // MMA number (q+u)*4 + k of this tile goes to partial accumulator (number % kParts);
              // the first MMA of every partial overwrites it
              const int mi = (q + u) * (kBK / 16) + k;
              const uint32_t dpart = d_tmem + (uint32_t)((mi % Cfg::kParts) * BN);
              const uint32_t accf = mi >= Cfg::kParts ? 1u : 0u;
              if constexpr (PAIR) umma_f16_2sm(dpart, da, db, idesc, accf);
              else umma_f16(dpart, da, db, idesc, accf);
            }
          }
          // frees the smem stage (of both CTAs of a pair) when these MMAs retire
          if constexpr (PAIR) umma_commit_2sm(&empty[stage]); else umma_commit(&empty[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if (nk == 0) {
          // nothing accumulated for this tile: epilogue must not read garbage
        }
        if constexpr (PAIR) umma_commit_2sm(&tfull[acc]); else umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================= epilogue =================================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;       // accumulator row owned by this thread
    int acc = 0;
    uint32_t acc_phase = 0;
    int gn_par = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step) {
      int nt;
      bool valid;
      long long obase, rbase = 0;
      int img = 0;
      bool has_acc = true;
      if constexpr (!A_MN) {
        nt = tile % p.nblks;
        const int mt = PAIR ? 2 * (tile / p.nblks) + (int)rank : tile / p.nblks;
        const int xb = mt % p.nxb;
        const int yb = (mt / p.nxb) % p.nyb;
        const int nb = mt / (p.nxb * p.nyb);
        const int tx = row % p.TW, ty = (row / p.TW) % p.TH, tn = row / (p.TW * p.TH);
        const int x = xb * p.TW + tx, y = yb * p.TH + ty, n = nb * p.TN + tn;
        valid = (x < p.W) && (y < p.H) && (n < p.N);
        img = n;
        obase = (long long)n * p.os_n + (long long)y * p.os_y + (long long)x * p.os_x;
        if (p.res) {
          rbase = (long long)n * p.rs_n + (long long)y * p.rs_y + (long long)x * p.rs_x;
          // the side input of this row: request all of its lines now (L2 prefetch), before the accumulator
          // wait - the epilogue's own register prefetch runs only one 64-column chunk ahead
          if (valid && !p.out_f32 && !p.out_atomic) {
            const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(p.res) + rbase + nt * BN;
            const int ncol = min(BN, p.Ncols - nt * BN);
            for (int c = 0; c < ncol; c += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + c));
          }
        }
      } else {
        const MnTile mt_ = decode_mn(p, tile);
        nt = mt_.nt;
        int m, tap;
        if (p.tpp == 2) { m = row & 63; tap = 2 * mt_.t + (row >> 6); }
        else            { m = mt_.mb * 128 + row; tap = mt_.t; }
        valid = (m < p.M) && (tap < p.ntaps);
        obase = (long long)p.tap_b[valid ? tap : 0] * p.os_tap + (long long)m * p.os_m;
        if (p.mn_batched)
          obase += (long long)(mt_.bz % p.nyb) * p.os_y + (long long)(mt_.bz / p.nyb) * p.os_n;
        has_acc = (mt_.pb1 - mt_.pb0) > 0;
      }
      auto wait_acc = [&]() {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
      };
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * Cfg::kParts * BN);
      if (!A_MN && !p.out_f32 && !p.out_atomic) {
        // bf16 output: stage through shared memory so global stores / side-input loads are coalesced
        EpiArgs ea{p.out, p.bias, p.rowvec, p.res, p.Ncols, p.alpha, p.gn_ab, p.gn_ws, p.N, p.ws_ld};
        epilogue_bf16_coalesced<BN, EPI, Cfg::kParts>(ea, epi_stage + q * 4096, t_addr, lane, nt * BN, valid && has_acc,
                                         obase, rbase, img, wait_acc, q, epi_xchg, &gn_par,
                                         EPI != EPI_PLAIN ? tile % p.gn_slots : 0);
      } else {
      wait_acc();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_addr + c0, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
#pragma unroll
        for (int pp = 1; pp < Cfg::kParts; ++pp) {        // sum the partial accumulators
          tmem_ld_32x32(t_addr + pp * BN + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        }
        const int col0 = nt * BN + c0;
        if (valid && has_acc && col0 < p.Ncols) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= p.alpha;
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(p.bias + col0 + j);
          }
          if (p.rowvec) {
            const float* rv = p.rowvec + (long long)img * p.Ncols + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __ldg(rv + j);
          }
          if (p.res) {
            const uint4* rp =
                reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.res) + rbase + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = rp[j];
              float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
                     d = unpack_bf16x2(u.w);
              f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
              f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
            }
          }
          if (p.out_atomic) {
            float* op = static_cast<float*>(p.out) + obase + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(op + j, f[j]);
          } else if (p.out_f32) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + obase + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            uint4* op =
                reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + obase + col0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              u.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
              u.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
              u.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
              u.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
              op[j] = u;
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(&tempty[acc], 0); else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // the leader's MMAs read the peer's smem
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int MODE, int EPI>
int launch_cfg(const CUtensorMap& mA, const CUtensorMap& mB, const TcDev& d, cudaStream_t stream) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    FDX_CUDA(cudaFuncSetAttribute(fdx_tc_kernel<BN, MODE, EPI, false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int grid = fdx_num_sms();
  if (grid <= 0) return FDX_ERR_NO_DEVICE;
  if (d.ntiles < grid) grid = d.ntiles;
  fdx_tc_kernel<BN, MODE, EPI, false><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mA, mB, d);
  fdx_note_kernel(FDX_KERNEL_TC);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

// CTA-pair launch: clusters of two CTAs, an even grid
template <int BN, int MODE>
int launch_pair(const CUtensorMap& mA, const CUtensorMap& mB, const TcDev& d, cudaStream_t stream) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    FDX_CUDA(cudaFuncSetAttribute(fdx_tc_kernel<BN, MODE, EPI_PLAIN, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int sms = fdx_num_sms();
  if (sms <= 0) return FDX_ERR_NO_DEVICE;
  int clusters = sms / 2;
  if (d.npairs < clusters) clusters = d.npairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FDX_CUDA(cudaLaunchKernelEx(&cfg, fdx_tc_kernel<BN, MODE, EPI_PLAIN, true>, mA, mB, d));
  fdx_note_kernel(FDX_KERNEL_TC);
  fdx_count_launch();
  return FDX_OK;
}

template <int BN, int MODE>
int launch_gn(const CUtensorMap& mA, const CUtensorMap& mB, const TcDev& d, cudaStream_t stream) {
  if constexpr (MODE == TC_KK) {
    if (d.gn_ab) return launch_cfg<BN, MODE, EPI_GN_BWD>(mA, mB, d, stream);
  }
  if constexpr (MODE == TC_KMN) {
    if (d.gn_ws && !d.gn_ab) return launch_cfg<BN, MODE, EPI_COLSTATS>(mA, mB, d, stream);
  }
  if constexpr (MODE != TC_MNMN && BN >= 128 && !(MODE == TC_KMN && BN == 192)) {
    if (d.npairs > 0) return launch_pair<BN, MODE>(mA, mB, d, stream);
  }
  return launch_cfg<BN, MODE, EPI_PLAIN>(mA, mB, d, stream);
}

template <int MODE>
int launch_mode(int BN, const CUtensorMap& mA, const CUtensorMap& mB, const TcDev& d,
                cudaStream_t stream) {
  switch (BN) {
    case 64: return launch_gn<64, MODE>(mA, mB, d, stream);
    case 128: return launch_gn<128, MODE>(mA, mB, d, stream);
    case 192: return launch_gn<192, MODE>(mA, mB, d, stream);
    case 256: return launch_gn<256, MODE>(mA, mB, d, stream);
  }
  fdx_set_error("tc: unsupported BN %d", BN);
  return FDX_ERR_UNSUPPORTED;
}

int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

}  // namespace

int fdx_tc_launch(const TcLaunch& L, cudaStream_t stream) {
  FDX_REQUIRE(L.mode >= TC_KK && L.mode <= TC_MNMN, "tc: bad mode %d", L.mode);
  FDX_REQUIRE(L.ntaps >= 1 && L.ntaps <= kTcMaxTaps, "tc: bad ntaps %d", L.ntaps);
  FDX_REQUIRE(L.es == 1 || L.es == 2, "tc: bad element stride %d", L.es);
  FDX_REQUIRE(L.Ncols > 0 && L.Ncols % 32 == 0, "tc: Ncols=%d must be a multiple of 32", L.Ncols);
  FDX_REQUIRE(L.A.strides[0] == 1 && L.B.strides[0] == 1, "tc: innermost strides must be 1");

  // Few output columns (Ncols = 64 / 128): the transposed formulation (fdx_tct.cu) issues M x 256 x 16 MMAs
  // instead of 128 x Ncols x 16 ones; faster on every such layer (profiles/layers_r01_tct.txt).
  if (L.mode != TC_MNMN && !getenv("FDX_NO_TCT")) {
    const int rt = fdx_tct_launch(L, stream);
    if (rt != FDX_ERR_UNSUPPORTED) return rt;
  }
  TcDev d{};
  d.mode = L.mode;
  d.W = L.W; d.H = L.H; d.N = L.N; d.es = L.es; d.ntaps = L.ntaps;
  d.es_b = L.es_b > 1 ? L.es_b : 1;
  FDX_REQUIRE(d.es_b <= 2, "tc: bad B element stride %d", L.es_b);
  bool b_view = d.es_b != 1;
  for (int i = 0; i < kTcMaxTaps; ++i) {
    d.tap_dx[i] = L.tap_dx[i]; d.tap_dy[i] = L.tap_dy[i]; d.tap_b[i] = L.tap_b[i];
    d.tap_bdx[i] = L.tap_bdx[i]; d.tap_bdy[i] = L.tap_bdy[i];
    b_view = b_view || L.tap_bdx[i] != 0 || L.tap_bdy[i] != 0;
  }
  FDX_REQUIRE(!b_view || L.mode == TC_MNMN, "tc: strided / shifted B views exist only in TC_MNMN");
  d.K = L.K; d.M = L.M; d.Ncols = L.Ncols; d.b_batched = L.b_batched; d.mn_batched = L.mn_batched;
  d.out = L.out; d.out_f32 = L.out_f32; d.out_atomic = L.out_atomic;
  d.os_x = L.os_x; d.os_y = L.os_y; d.os_n = L.os_n; d.os_tap = L.os_tap; d.os_m = L.os_m;
  d.alpha = L.alpha; d.bias = L.bias; d.rowvec = L.rowvec; d.res = L.res;
  d.rs_x = L.rs_x; d.rs_y = L.rs_y; d.rs_n = L.rs_n;
  d.gn_ab = L.gn_ab; d.gn_ws = L.gn_ws; d.gn_slots = L.gn_slots > 0 ? L.gn_slots : 1;
  d.ws_ld = L.ws_ld > 0 ? L.ws_ld : L.Ncols;
  if (L.gn_ws && !L.gn_ab) {
    FDX_REQUIRE(L.mode == TC_KMN && !L.gemm_like && !L.out_f32 && !L.out_atomic,
                "tc: output column statistics need a bf16 convolution-forward launch");
  }
  if (L.gn_ab) {
    FDX_REQUIRE(L.mode == TC_KK && !L.gemm_like && L.res && L.gn_ws && !L.out_f32 && !L.out_atomic &&
                    !L.bias && !L.rowvec,
                "tc: GroupNorm-backward fusion needs a plain bf16 data-gradient launch");
  }

  // ---- column tile --------------------------------------------------------
  // cost model per row tile: ntiles * (BN + 64)  (MMA columns incl. padding + A re-read per tile)
  int BN = 64;
  {
    long long best = -1;
    const int cands[4] = {256, 192, 128, 64};
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      const long long nt = (L.Ncols + bn - 1) / bn;
      const long long cost = nt * (bn + 64);
      if (best < 0 || cost < best) { best = cost; BN = bn; }
    }
  }
  const int rows_per_tile = (L.mode == TC_MNMN) ? 64 : 128;
  // ---- pixel box ----------------------------------------------------------
  int TW, TH, TN;
  if (!L.gemm_like) {
    // conv: a 16 x 8 (or 16 x 4) pixel patch; small images pack several per tile
    TW = pow2_floor(L.W < 16 ? L.W : 16);
    int rem = rows_per_tile / TW;
    TH = pow2_floor(L.H < rem ? L.H : rem);
    TN = rem / TH;
  } else {
    // GEMM: rows along x only; (y, n) are batch coordinates. A box taller than the
    // matrix is fine: TMA zero-fills and the epilogue predicates on x < W.
    TW = rows_per_tile;
    TH = 1;
    TN = 1;
  }
  d.TW = TW; d.TH = TH; d.TN = TN;
  if (L.gn_ws && TN != 1) {
    fdx_set_error("tc: fused per-image column sums need at least 128 pixels per image (got %dx%d)", L.W, L.H);
    return FDX_ERR_UNSUPPORTED;
  }
  d.nxb = (L.W + TW - 1) / TW;
  d.nyb = (L.H + TH - 1) / TH;
  d.nnb = (L.N + TN - 1) / TN;
  const long long pix_blocks = (long long)d.nxb * d.nyb * d.nnb;

  // halo-sharing variant: fewer TMA writes, but measured slower than the generic kernel on B200
  // (both are shared-memory-bandwidth bound); kept as an opt-in experiment.
  if (L.mode != TC_MNMN && !L.gn_ws && getenv("FDX_CONV3")) {
    const int r3 = fdx_conv3_launch(L, BN, stream);
    if (r3 != FDX_ERR_UNSUPPORTED) return r3;
  }
  if (L.mode != TC_MNMN) {
    // Round 1 narrowed the column tile until the launch had two waves of tiles.  But a K = 16 MMA of this
    // (pixels-as-M) engine costs ~150 cycles whatever its N, so a tile's duration does not shrink
    // with BN and more, narrower tiles only add rounds: 8x8 images (128 pixel tiles at B = 256) ran 256 -> 256 at
    // BN = 64 in four rounds instead of one (47 us; 407 TF/s).  FDX_TC_BN_V1=1 restores the old rule.
    static const bool bn_v1 = getenv("FDX_TC_BN_V1") != nullptr;
    while (bn_v1 && BN > 64 && pix_blocks * ((L.Ncols + BN - 1) / BN) < 2LL * fdx_num_sms())
      BN = (BN == 192) ? 64 : BN / 2;
    d.nblks = (L.Ncols + BN - 1) / BN;
    d.kchunks = (L.K + kBK - 1) / kBK;
    d.ntiles = (int)(pix_blocks * d.nblks);
    // CTA pairs (cta_group::2): plain bf16 epilogue, BN >= 128 (MN-major B needs whole 64-column blocks
    // per CTA: not BN = 192); opt-in while it is being measured
    d.npairs = 0;
    if (getenv("FDX_PAIR") && !L.gn_ws && BN >= 128 && !(L.mode == TC_KMN && BN == 192) && pix_blocks >= 2)
      d.npairs = (int)(((pix_blocks + 1) / 2) * d.nblks);
    d.splits = 1;
    d.mblks = 1;
    d.tpp = 1;
    d.ntp = L.ntaps;
    FDX_REQUIRE(!L.res || !L.out_atomic, "tc: residual with atomic output unsupported");
  } else {
    d.nblks = (L.Ncols + BN - 1) / BN;
    d.tpp = (L.M <= 64 && L.ntaps > 1 && !b_view) ? 2 : 1;
    d.ntp = (L.ntaps + d.tpp - 1) / d.tpp;
    d.mblks = (L.M + 127) / 128;
    d.kchunks = 0;
    long long base_tiles = (long long)d.ntp * d.mblks * d.nblks;
    const long long kblocks = L.mn_batched ? d.nxb : pix_blocks;
    const long long nbz = L.mn_batched ? (long long)d.nyb * d.nnb : 1;
    int splits = L.splits;
    if (splits <= 0) {
      // wave-aware split-K: pick the split count whose tile total fills w waves of the machine
      // best (w = 1..4); more waves only when they raise the fill by > 3 %.
      const int sms = fdx_num_sms();
      const long long bt = base_tiles * nbz;
      double best_eff = -1.0;
      splits = 1;
      for (int w = 1; w <= 4; ++w) {
        long long sp = ((long long)sms * w) / bt;
        if (sp < 1) sp = 1;
        if (sp > kblocks) sp = kblocks;
        const long long tiles = bt * sp;
        const long long waves = (tiles + sms - 1) / sms;
        const double eff = (double)tiles / (double)(waves * sms);
        if (eff > best_eff + 0.03) { best_eff = eff; splits = (int)sp; }
      }
    }
    if (splits > kblocks) splits = (int)kblocks;
    FDX_REQUIRE(splits == 1 || L.out_atomic, "tc: split-K needs atomic output");
    d.splits = splits;
    d.ntiles = (int)(base_tiles * splits * nbz);
  }

  // ---- tensor maps ----------------------------------------------------------
  CUtensorMap mA, mB;
  {
    uint64_t dims[4], str[3];
    for (int i = 0; i < 4; ++i) dims[i] = L.A.dims[i];
    for (int i = 1; i < 4; ++i) str[i - 1] = L.A.strides[i] * 2;
    uint32_t box[4] = {64, (uint32_t)(TW * L.es), (uint32_t)(TH * L.es), (uint32_t)TN};
    uint32_t est[4] = {1, (uint32_t)L.es, (uint32_t)L.es, 1};
    int s = fdx_make_tmap_bf16(&mA, L.A.ptr, 4, dims, str, box, est, 1);
    if (s != FDX_OK) return s;
  }
  {
    uint64_t dims[4], str[3];
    for (int i = 0; i < 4; ++i) dims[i] = L.B.dims[i];
    for (int i = 1; i < 4; ++i) str[i - 1] = L.B.strides[i] * 2;
    uint32_t box[4];
    uint32_t est[4] = {1, 1, 1, 1};
    if (L.mode == TC_KK) {
      box[0] = 64; box[1] = (uint32_t)(d.npairs > 0 ? BN / 2 : BN); box[2] = 1; box[3] = 1;
    } else if (L.mode == TC_KMN) {
      box[0] = 64; box[1] = 64; box[2] = 1; box[3] = 1;
    } else {
      box[0] = 64; box[1] = (uint32_t)(TW * d.es_b); box[2] = (uint32_t)(TH * d.es_b); box[3] = (uint32_t)TN;
      est[1] = est[2] = (uint32_t)d.es_b;
    }
    int s = fdx_make_tmap_bf16(&mB, L.B.ptr, 4, dims, str, box, est, 1);
    if (s != FDX_OK) return s;
  }

  switch (L.mode) {
    case TC_KK: return launch_mode<TC_KK>(BN, mA, mB, d, stream);
    case TC_KMN: return launch_mode<TC_KMN>(BN, mA, mB, d, stream);
    default: return launch_mode<TC_MNMN>(BN, mA, mB, d, stream);
  }
}

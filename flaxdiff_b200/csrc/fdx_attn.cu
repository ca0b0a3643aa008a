// fdx_attn.cu -- fused scaled-dot-product attention for sm_100a: nn.dot_product_attention as used by
// NormalAttention (flaxdiff/models/attention.py:156-177, softmax(q k^T / sqrt(d)) v, self- and cross-attention
// to the 77 x 768 text context) and its backward (jax.value_and_grad, trainer/general_diffusion_trainer.py:321).
//
// The round-1 path ran QK^T, softmax, PV as separate launches with the f32 logits S and the bf16 probabilities
// P round-tripping HBM ((B, h, L, Lk): 2.1 GB + 1 GB per block at 256x256).  Here S and P never leave the SM:
//
//   forward  (one CTA per 128 queries x head x image; flash-style loop over 128-key blocks)
//     warp 0      TMA: Q tile once, K_j / V_j tiles through a 3-stage mbarrier ring (128B-swizzled boxes of
//                 64 channels; a head narrower than 64 is addressed by a byte offset inside the swizzle row)
//     warp 1      tcgen05.mma: S_j = Q K_j^T into a double-buffered TMEM accumulator (2 x 128 columns), then
//                 O_j = P_j V_j into a second pair (2 x 64 columns) once the softmax warps have published P_j
//     warps 2..5  softmax: one query row per thread, tcgen05.ld of its S row, running max / sum in the log2
//                 domain, P_j written as bf16 straight into the K-major swizzled shared-memory tile that is the
//                 A operand of the PV MMA (fence.proxy.async), O accumulated in registers with the usual
//                 exp2(m_old - m_new) correction, final O / l and the log-sum-exp row for the backward.
//   backward = two kernels with the same skeleton, S and dP recomputed on chip (FlashAttention-2 recipe,
//     7 small GEMMs instead of 5, no atomics, deterministic):
//       fdx_attn_bwd_dq   per 128-query tile, loop over key blocks:  dS = P o (dP - D) ;  dQ += dS K
//       fdx_attn_bwd_dkv  per 128-key block, loop over query tiles (transposed tiles, TMEM lane = key):
//                         dV += P^T dO ;  dK += dS^T Q
//     with P = exp2(s * scale*log2e - lse*log2e), D = rowsum(dO o O), dS including the 1/sqrt(d) factor.
// Tensors are bf16 [B][L][heads * dh] (row stride and batch stride given), dh in {32, 64}; keys beyond Lk and
// query rows beyond L are handled by TMA zero fill + masking, so 77-key cross-attention needs no padded copies.
#include "fdx_common.cuh"
#include "../../include/fdx.h"
#include <math.h>
#include <type_traits>
#include <stdlib.h>

namespace {

constexpr int kThreads = 192;
constexpr int kTile = 128 * 128;          // bytes of one [128 rows][64 ch] bf16 swizzled tile
constexpr int kKV = 3;                    // K/V ring stages (forward, dQ kernel)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnDev {
  int B, heads, L, Lk, dh;
  float scale_log2;        // softmax scale * log2(e)
  float scale;             // softmax scale
  __nv_bfloat16* o; long long o_ld, o_bs;
  float* lse;              // [B][heads][L]
  // backward
  const __nv_bfloat16* d_o; long long do_ld, do_bs;
  const __nv_bfloat16* o_in; long long oin_ld, oin_bs;
  float* dvec;             // [B][heads][L] rowsum(dO o O)
  __nv_bfloat16* dq; long long dq_ld, dq_bs;
  __nv_bfloat16* dk; long long dk_ld, dk_bs;
  __nv_bfloat16* dv; long long dv_ld, dv_bs;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// byte offset of the 16-byte piece (row r, piece p) inside a [rows][128 B] 128B-swizzled tile
__device__ __forceinline__ uint32_t swz(int r, int p) { return (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)); }

// store 32 consecutive bf16 (columns c0 .. c0+31 of row r) of a [128][128-key] K-major tile pair
// (two 64-key chunks of 16 KB each)
__device__ __forceinline__ void store_p_chunk(uint32_t sP, int r, int c0, const uint32_t (&w)[16]) {
  const uint32_t base = sP + (uint32_t)(c0 >> 6) * kTile;
  const int p0 = (c0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + swz(r, p0 + i)), "r"(w[4 * i]),
                 "r"(w[4 * i + 1]), "r"(w[4 * i + 2]), "r"(w[4 * i + 3])
                 : "memory");
}

// ======================================================================================================
// forward
// ======================================================================================================
// SHORT = the keys fit one 128-key block (77 text tokens): no ring, one P tile, 256 TMEM columns and 80 KB of
// shared memory, so TWO CTAs share an SM - with L/128 x heads x B short-lived CTAs (131 072 at 128x128, B = 128)
// the prologue (barrier init, TMEM allocation, first TMA round trip) of one overlaps the math of the other.
// DUAL (long key sequences) = the same economy for the flash loop: S, P and O single-buffered (256 TMEM columns),
// a two-stage K/V ring, 112 KB of shared memory - again two CTAs per SM.  Inside one CTA the chain
// S_j -> softmax_j -> PV_j is then serial (the next S is issued as soon as the softmax warps have consumed the
// current one), and the overlap the double buffers gave comes from the OTHER resident CTA instead - which also
// doubles the softmax warps per scheduler, the actual limiter of the one-CTA arrangement (one warp per
// scheduler cannot hide its tcgen05.ld / MUFU latencies) and hides the per-CTA prologue and epilogue.
template <int DH, bool SHORT, bool DUAL>
__global__ void __launch_bounds__(kThreads, (SHORT || DUAL) ? 2 : 1)
fdx_attn_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapV, const AttnDev p) {
  static_assert(!(SHORT && DUAL), "SHORT has no loop to single-buffer");
  extern __shared__ __align__(1024) uint8_t smem_fwd_raw[];
  uint8_t* const smem_raw = smem_fwd_raw;
  uint8_t* smem;
  if constexpr (DUAL) {
    // no alignment slack in the request (2 x 113 KB + the per-CTA reserve must fit the SM): the dynamic window
    // has to start 1024-byte aligned by itself - fail loudly if it ever does not
    smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  } else {
    smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  }
  constexpr bool ONE = SHORT || DUAL;                // single S / P / O buffers
  constexpr int NKV = SHORT ? 1 : (DUAL ? 2 : kKV);  // K/V ring stages
  constexpr int NPB = ONE ? 1 : 2;                   // P tiles
  constexpr uint32_t O_COL = ONE ? 128u : 256u;      // first TMEM column of the O accumulator(s)
  constexpr uint32_t TMEM_COLS = ONE ? 256u : 512u;
  auto sb = [](int j) { return ONE ? 0 : (j & 1); };                 // S / P / O buffer of block j
  auto sph = [](int j) { return ONE ? (j & 1) : ((j >> 1) & 1); };   // ... and the parity of its barrier phase
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTile;                       // NKV x (K tile, V tile)
  uint8_t* sP = sKV + NKV * 2 * kTile;               // NPB x (two 64-key chunks)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NPB * 2 * kTile);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;          // [kKV]
  uint64_t* kv_empty = bars + 1 + kKV;   // [kKV]
  uint64_t* s_full = bars + 1 + 2 * kKV; // [2]
  uint64_t* p_full = s_full + 2;         // [2]
  uint64_t* o_full = p_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * 128;
  const int nblk = (p.Lk + 127) / 128;
  const int c0 = (head * DH) / 64 * 64;                     // first channel of the 64-wide TMA box
  const uint32_t koff = (uint32_t)((head * DH) % 64) * 2;   // byte offset of this head inside the swizzle row

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ); tma_prefetch_desc(&mapK); tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKV; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTile);
      tma_load_4d(sQ, &mapQ, q_full, c0, q0, b, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % NKV;
        mbar_wait(&kv_empty[s], ((j / NKV) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kTile);
        tma_load_4d(sKV + s * 2 * kTile, &mapK, &kv_full[s], c0, j * 128, b, 0);
        tma_load_4d(sKV + s * 2 * kTile + kTile, &mapV, &kv_full[s], c0, j * 128, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idO = umma_idesc_bf16(128, DH, 0, 1);
      const uint32_t aQ = smem_u32(sQ) + koff;
      auto issue_s = [&](int j) {
        const int s = j % NKV;
        mbar_wait(&kv_full[s], (j / NKV) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sKV + s * 2 * kTile) + koff;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_f16(tmem_base + (uint32_t)(sb(j) * 128), umma_desc_sw128(aQ + k * 32, 16, 1024),
                   umma_desc_sw128(aK + k * 32, 16, 1024), idS, k != 0 ? 1u : 0u);
        umma_commit(&s_full[sb(j)]);
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (!ONE && j + 1 < nblk) issue_s(j + 1);
        mbar_wait(&p_full[sb(j)], sph(j));
        tc_fence_after();
        // single S buffer: the softmax warps have consumed S_j (they publish P_j after their last tcgen05.ld)
        if (ONE && j + 1 < nblk) issue_s(j + 1);
        const int s = j % NKV;
        const uint32_t aP = smem_u32(sP + (j % NPB) * 2 * kTile);
        const uint32_t aV = smem_u32(sKV + s * 2 * kTile + kTile) + koff;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem_base + O_COL + (uint32_t)(sb(j) * 64),
                   umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(aV + k * 2048, 8192, 1024), idO, k != 0 ? 1u : 0u);
        umma_commit(&o_full[sb(j)]);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY, l = 0.f, corr_prev = 0.f;
    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.f;

    auto fold_o = [&](int j, float corr) {      // acc = acc * corr + O_j
      mbar_wait(&o_full[sb(j)], sph(j));
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + O_COL + (uint32_t)(sb(j) * 64 + c), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[c + i] = fmaf(acc[c + i], corr, __uint_as_float(v[i]));
      }
    };

    for (int j = 0; j < nblk; ++j) {
      mbar_wait(&s_full[sb(j)], sph(j));
      tc_fence_after();
      const uint32_t sa = lane_addr + (uint32_t)(sb(j) * 128);
      const int kbase = j * 128;
      // pass A: row maximum over the valid keys (only the last key block can be partial)
      const bool partial = kbase + 128 > p.Lk;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(sa + c, v);
        tmem_ld_wait();
        if (partial) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c + i < p.Lk) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
          float m4[4] = {mx, -INFINITY, -INFINITY, -INFINITY};        // four independent chains
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) m4[u] = fmaxf(m4[u], __uint_as_float(v[i + u]));
          }
          mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
      }
      const float m_new = fmaxf(m, mx * p.scale_log2);
      const float corr = ex2(m - m_new);           // first block: exp2(-inf) = 0
      // single P / O buffers: PV_{j-1} must have retired before P_j overwrites its operand, and O_{j-1} must be
      // read before PV_j overwrites it - fold it in here, between the two passes
      if (ONE && j > 0) fold_o(j - 1, corr_prev);
      // pass B: probabilities -> bf16 -> the swizzled A tile of the PV MMA
      float rs = 0.f;
      const uint32_t sPj = smem_u32(sP + (j % NPB) * 2 * kTile);
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32], w[16];
        tmem_ld_32x32(sa + c, v);
        tmem_ld_wait();
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float pe[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            pe[u] = ex2(fmaf(__uint_as_float(v[i + u]), p.scale_log2, -m_new));
            if (partial && !(kbase + c + i + u < p.Lk)) pe[u] = 0.f;
            rs4[u] += pe[u];
          }
          w[i >> 1] = pack_bf16x2(pe[0], pe[1]);
          w[(i >> 1) + 1] = pack_bf16x2(pe[2], pe[3]);
        }
        rs += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
        store_p_chunk(sPj, row, c, w);
      }
      l = fmaf(l, corr, rs);
      m = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb(j)]);
      if (!ONE && j > 0) fold_o(j - 1, corr_prev);
      corr_prev = corr;
    }
    fold_o(nblk - 1, corr_prev);
    if (q0 + row < p.L) {
      const float inv = 1.f / l;
      __nv_bfloat16* op = p.o + (long long)b * p.o_bs + (long long)(q0 + row) * p.o_ld + head * DH;
#pragma unroll
      for (int c = 0; c < DH; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(acc[c] * inv, acc[c + 1] * inv);
        u.y = pack_bf16x2(acc[c + 2] * inv, acc[c + 3] * inv);
        u.z = pack_bf16x2(acc[c + 4] * inv, acc[c + 5] * inv);
        u.w = pack_bf16x2(acc[c + 6] * inv, acc[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
      if (p.lse) p.lse[((long long)b * p.heads + head) * p.L + q0 + row] = m * kLn2 + logf(l);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ======================================================================================================
// backward, part 0: D[b][h][row] = sum_c dO[row][c] * O[row][c]
// ======================================================================================================
__global__ void attn_dvec_kernel(const AttnDev p) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (b, row, head)
  const long long total = (long long)p.B * p.L * p.heads;
  if (idx >= total) return;
  const int head = (int)(idx % p.heads);
  const long long r = idx / p.heads;
  const int row = (int)(r % p.L), b = (int)(r / p.L);
  const __nv_bfloat16* a = p.d_o + (long long)b * p.do_bs + (long long)row * p.do_ld + head * p.dh;
  const __nv_bfloat16* o = p.o_in + (long long)b * p.oin_bs + (long long)row * p.oin_ld + head * p.dh;
  float s = 0.f;
  for (int c = 0; c < p.dh; c += 8) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + c), uo = *reinterpret_cast<const uint4*>(o + c);
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wo[4] = {uo.x, uo.y, uo.z, uo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fa = unpack_bf16x2(wa[i]), fo = unpack_bf16x2(wo[i]);
      s = fmaf(fa.x, fo.x, s);
      s = fmaf(fa.y, fo.y, s);
    }
  }
  p.dvec[((long long)b * p.heads + head) * p.L + row] = s;
}

// ======================================================================================================
// backward, shared skeleton.  ROWS_ARE_KEYS = false: the dQ kernel (tile rows = queries, loop over key blocks);
// true: the dK / dV kernel (tile rows = keys, loop over query tiles).  Per loop step two 128 x 128 products land
// in TMEM (cols 0..127: S or S^T; 128..255: dP or dP^T), the four softmax warps turn them into the bf16 tiles
// E1 = P (only needed for dV) and E2 = dS in shared memory, and the MMA warp accumulates
//   dQ  += dS  . K_j                (cols 256..319)                        [ROWS_ARE_KEYS = false]
//   dV  += P^T . dO_i  (256..319),  dK += dS^T . Q_i  (320..383)           [ROWS_ARE_KEYS = true]
// across the whole loop (no rescaling: P is recomputed from the saved log-sum-exp).
// Shared memory: R1, R2 = the two row-side operands (Q, dO | K, V), ring of (C1, C2) = the column-side pair
// (K_j, V_j | Q_i, dO_i), E1 / E2 tiles of 2 x 16 KB each.
// ======================================================================================================
// EWG = number of 4-warp groups that turn S / dP into the E tiles.  The element-wise step (two tcgen05.ld, exp2,
// two bf16 packs and swizzled stores per element) is what bounds the kernel - 128 columns per thread with one
// warp per scheduler cannot hide its own latencies - and it has no cross-column dependency (the softmax
// statistics are inputs), so with EWG = 2 the warps (2..5) take columns 0..63 of every row and (6..9) columns
// 64..127: two warps per scheduler on the same TMEM lanes.
template <int DH, bool ROWS_ARE_KEYS, int EWG>
__global__ void __launch_bounds__(64 + 128 * EWG, 1)
fdx_attn_bwd_kernel(const __grid_constant__ CUtensorMap mapR1, const __grid_constant__ CUtensorMap mapR2,
                    const __grid_constant__ CUtensorMap mapC1, const __grid_constant__ CUtensorMap mapC2,
                    const AttnDev p) {
  constexpr int RING = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sR1 = smem;
  uint8_t* sR2 = smem + kTile;
  uint8_t* sC = smem + 2 * kTile;                      // RING x (C1, C2)
  uint8_t* sE1 = sC + RING * 2 * kTile;                // 2 chunks
  uint8_t* sE2 = sE1 + 2 * kTile;                      // 2 chunks
  float* sStat = reinterpret_cast<float*>(sE2 + 2 * kTile);     // RING x (lse[128], dvec[128]) of the column side
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + RING * 256);
  uint64_t* r_full = bars;
  uint64_t* c_full = bars + 1;            // [RING]
  uint64_t* c_empty = c_full + RING;      // [RING]
  uint64_t* s_full = c_empty + RING;      // S and dP of this step are in TMEM
  uint64_t* e_full = s_full + 1;          // E tiles published (count 4)
  uint64_t* e_free = e_full + 1;          // accumulating MMAs that read the E tiles retired
  uint64_t* done = e_free + 1;            // all accumulation finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int r0 = rt * 128;
  const int nrows = ROWS_ARE_KEYS ? p.Lk : p.L;        // extent of the row side
  const int ncols = ROWS_ARE_KEYS ? p.L : p.Lk;        // extent of the looped (column) side
  const int nstep = (ncols + 127) / 128;
  const int c0 = (head * DH) / 64 * 64;
  const uint32_t koff = (uint32_t)((head * DH) % 64) * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapR1); tma_prefetch_desc(&mapR2); tma_prefetch_desc(&mapC1); tma_prefetch_desc(&mapC2);
    mbar_init(r_full, 1);
    for (int i = 0; i < RING; ++i) { mbar_init(&c_full[i], 1); mbar_init(&c_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(e_full, 4 * EWG);
    mbar_init(e_free, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(r_full, 2 * kTile);
      tma_load_4d(sR1, &mapR1, r_full, c0, r0, b, 0);
      tma_load_4d(sR2, &mapR2, r_full, c0, r0, b, 0);
      for (int j = 0; j < nstep; ++j) {
        const int s = j % RING;
        mbar_wait(&c_empty[s], ((j / RING) & 1) ^ 1);
        mbar_arrive_expect_tx(&c_full[s], 2 * kTile);
        tma_load_4d(sC + s * 2 * kTile, &mapC1, &c_full[s], c0, j * 128, b, 0);
        tma_load_4d(sC + s * 2 * kTile + kTile, &mapC2, &c_full[s], c0, j * 128, b, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idA = umma_idesc_bf16(128, DH, 0, 1);
      const uint32_t aR1 = smem_u32(sR1) + koff, aR2 = smem_u32(sR2) + koff;
      mbar_wait(r_full, 0);
      tc_fence_after();
      for (int j = 0; j < nstep; ++j) {
        const int s = j % RING;
        mbar_wait(&c_full[s], (j / RING) & 1);
        tc_fence_after();
        const uint32_t aC1 = smem_u32(sC + s * 2 * kTile) + koff, aC2 = aC1 + kTile;
        // the softmax warps must have finished reading the previous step's S / dP: implied by e_full(j-1),
        // which this thread waited for before issuing the previous accumulation
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          // S (or S^T) = R1 . C1^T ;  dP (or dP^T) = R2 . C2^T
          umma_f16(tmem_base, umma_desc_sw128(aR1 + k * 32, 16, 1024), umma_desc_sw128(aC1 + k * 32, 16, 1024),
                   idS, k != 0 ? 1u : 0u);
          umma_f16(tmem_base + 128u, umma_desc_sw128(aR2 + k * 32, 16, 1024),
                   umma_desc_sw128(aC2 + k * 32, 16, 1024), idS, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        mbar_wait(e_full, j & 1);
        tc_fence_after();
        const uint32_t aE1 = smem_u32(sE1), aE2 = smem_u32(sE2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t eo = (uint32_t)((k >> 2) * kTile + (k & 3) * 32);
          if constexpr (ROWS_ARE_KEYS) {
            // dV += P^T . dO_i  (B = C2 = dO_i, MN-major) ;  dK += dS^T . Q_i  (B = C1 = Q_i)
            umma_f16(tmem_base + 256u, umma_desc_sw128(aE1 + eo, 16, 1024),
                     umma_desc_sw128(aC2 + k * 2048, 8192, 1024), idA, (j | k) != 0 ? 1u : 0u);
            umma_f16(tmem_base + 320u, umma_desc_sw128(aE2 + eo, 16, 1024),
                     umma_desc_sw128(aC1 + k * 2048, 8192, 1024), idA, (j | k) != 0 ? 1u : 0u);
          } else {
            // dQ += dS . K_j  (B = C1 = K_j, MN-major)
            umma_f16(tmem_base + 256u, umma_desc_sw128(aE2 + eo, 16, 1024),
                     umma_desc_sw128(aC1 + k * 2048, 8192, 1024), idA, (j | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(e_free);
        umma_commit(&c_empty[s]);
      }
      umma_commit(done);
    }
  } else {
    const int q = warp & 3;
    const int wg = (warp - 2) >> 2;                    // element-wise warp group: columns [wg * 128 / EWG, ...)
    constexpr int CW = 128 / EWG;                      // columns per thread and step
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const long long stat_base = ((long long)b * p.heads + head) * p.L;
    // dQ kernel: this thread's query row statistics; dK/dV kernel: they belong to the COLUMNS (queries)
    float lse_r = 0.f, d_r = 0.f;
    if constexpr (!ROWS_ARE_KEYS) {
      if (r0 + row < p.L) {
        lse_r = p.lse[stat_base + r0 + row] * kLog2e;
        d_r = p.dvec[stat_base + r0 + row] * p.scale;
      }
    }
    const uint32_t sE1a = smem_u32(sE1), sE2a = smem_u32(sE2);
    for (int j = 0; j < nstep; ++j) {
      const int cbase = j * 128;
      if constexpr (ROWS_ARE_KEYS) {
        // stage the query-side statistics of this step (parity buffer j & 1), 128 threads = 128 columns
        if (wg == 0) {
          float* st = sStat + (j & 1) * 256;
          const bool ok = cbase + row < p.L;
          st[row] = ok ? p.lse[stat_base + cbase + row] * kLog2e : 0.f;
          st[128 + row] = ok ? p.dvec[stat_base + cbase + row] * p.scale : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EWG) : "memory");
      }
      if (j > 0) mbar_wait(e_free, (j - 1) & 1);      // previous step's accumulation has consumed E1 / E2
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const float* st = sStat + (j & 1) * 256;
      const bool row_in = ROWS_ARE_KEYS ? (r0 + row < p.Lk) : (r0 + row < p.L);
      const int col_lim = (ROWS_ARE_KEYS ? p.L : p.Lk) - cbase;      // columns of this step inside the tensor
      // P = exp2(s * scale*log2e - lse*log2e);  dS = P * (dP * scale - D * scale)  (D * scale staged / hoisted).
      // Only the last step / the last row tile needs the bounds masks: MASK = false is the 5-instruction path.
      auto ewise = [&](auto mask_tag) {
        constexpr bool MASK = decltype(mask_tag)::value;
#pragma unroll 1
        for (int c = wg * CW; c < (wg + 1) * CW; c += 32) {
          uint32_t vs[32], vp[32], w1[16], w2[16];
          tmem_ld_32x32(lane_addr + (uint32_t)c, vs);
          tmem_ld_32x32(lane_addr + 128u + (uint32_t)c, vp);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float lq[4] = {lse_r, lse_r, lse_r, lse_r}, dq4[4] = {d_r, d_r, d_r, d_r};
            if constexpr (ROWS_ARE_KEYS) {
              const float4 a4 = *reinterpret_cast<const float4*>(st + c + i);
              const float4 b4 = *reinterpret_cast<const float4*>(st + 128 + c + i);
              lq[0] = a4.x; lq[1] = a4.y; lq[2] = a4.z; lq[3] = a4.w;
              dq4[0] = b4.x; dq4[1] = b4.y; dq4[2] = b4.z; dq4[3] = b4.w;
            }
            float pr[4], ds[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float pv = ex2(fmaf(__uint_as_float(vs[i + u]), p.scale_log2, -lq[u]));
              if (MASK && !(row_in && (c + i + u < col_lim))) pv = 0.f;
              pr[u] = pv;
              ds[u] = pv * fmaf(__uint_as_float(vp[i + u]), p.scale, -dq4[u]);
            }
            w1[i >> 1] = pack_bf16x2(pr[0], pr[1]);
            w1[(i >> 1) + 1] = pack_bf16x2(pr[2], pr[3]);
            w2[i >> 1] = pack_bf16x2(ds[0], ds[1]);
            w2[(i >> 1) + 1] = pack_bf16x2(ds[2], ds[3]);
          }
          if constexpr (ROWS_ARE_KEYS) store_p_chunk(sE1a, row, c, w1);
          store_p_chunk(sE2a, row, c, w2);
        }
      };
      // warp-uniform choice (the TMEM loads inside are warp-collective)
      const bool need_mask = __any_sync(0xffffffffu, !row_in) || col_lim < 128;
      if (need_mask) ewise(std::true_type{}); else ewise(std::false_type{});
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(e_full);
    }
    // epilogue: accumulators -> bf16 rows.  tcgen05.ld is warp-collective (.sync.aligned): EVERY lane executes
    // it, only the global stores are predicated on the row being inside the tensor.
    mbar_wait(done, 0);
    tc_fence_after();
    const bool row_ok = r0 + row < nrows;
    constexpr int NOUT = ROWS_ARE_KEYS ? 2 : 1;
#pragma unroll
    for (int which = 0; which < NOUT; ++which) {
      __nv_bfloat16* base;
      if constexpr (ROWS_ARE_KEYS)
        base = which == 0 ? p.dv + (long long)b * p.dv_bs + (long long)(r0 + row) * p.dv_ld
                          : p.dk + (long long)b * p.dk_bs + (long long)(r0 + row) * p.dk_ld;
      else
        base = p.dq + (long long)b * p.dq_bs + (long long)(r0 + row) * p.dq_ld;
      base += head * DH;
#pragma unroll
      for (int c = 0; c < DH; c += 32) {
        if (EWG == 2 && ((which * (DH / 32) + c / 32) & 1) != wg) continue;   // the 32-column pieces alternate
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + 256u + (uint32_t)(which * 64 + c), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
            *reinterpret_cast<uint4*>(base + c + i) = u;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int make_map(CUtensorMap* m, const void* ptr, int hd, int rows, int B, long long ld, long long bs) {
  uint64_t dims[4] = {(uint64_t)hd, (uint64_t)rows, (uint64_t)B, 1};
  uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)bs * 2, (uint64_t)bs * 2 * (uint64_t)B};
  uint32_t box[4] = {64, 128, 1, 1};
  uint32_t est[4] = {1, 1, 1, 1};
  return fdx_make_tmap_bf16(m, ptr, 4, dims, str, box, est, 1);
}

int check_desc(const fdx_attn_desc* a, const char* who) {
  FDX_REQUIRE(a && a->q && a->k && a->v, "%s: null tensor", who);
  FDX_REQUIRE(a->B > 0 && a->heads > 0 && a->L > 0 && a->Lk > 0, "%s: bad sizes", who);
  FDX_REQUIRE(a->dh == 32 || a->dh == 64, "%s: head width %d (stored) must be 32 or 64 - zero-pad narrower heads", who,
              a->dh);
  FDX_REQUIRE((a->heads * a->dh) % 64 == 0, "%s: heads * dh must be a multiple of 64", who);
  FDX_REQUIRE(a->q_ld % 8 == 0 && a->k_ld % 8 == 0 && a->v_ld % 8 == 0, "%s: row strides must be multiples of 8", who);
  return FDX_OK;
}

}  // namespace

extern "C" {

int fdx_attention_fwd(const fdx_attn_desc* a, void* stream) {
  int s = check_desc(a, "attention_fwd");
  if (s != FDX_OK) return s;
  FDX_REQUIRE(a->o, "attention_fwd: null output");
  const int hd = a->heads * a->dh;
  CUtensorMap mQ, mK, mV;
  if ((s = make_map(&mQ, a->q, hd, a->L, a->B, a->q_ld, a->q_bs)) != FDX_OK) return s;
  if ((s = make_map(&mK, a->k, hd, a->Lk, a->B, a->k_ld, a->k_bs)) != FDX_OK) return s;
  if ((s = make_map(&mV, a->v, hd, a->Lk, a->B, a->v_ld, a->v_bs)) != FDX_OK) return s;
  AttnDev d{};
  d.B = a->B; d.heads = a->heads; d.L = a->L; d.Lk = a->Lk; d.dh = a->dh;
  d.scale = a->scale; d.scale_log2 = a->scale * kLog2e;
  d.o = (__nv_bfloat16*)a->o; d.o_ld = a->o_ld; d.o_bs = a->o_bs;
  d.lse = a->lse;
  const bool short_keys = a->Lk <= 128 && !getenv("FDX_ATTN_NO_SHORT");
  const bool dual = !short_keys && !getenv("FDX_ATTN_NO_DUAL");      // two CTAs per SM for the flash loop
  const int smem = short_keys ? (1 + 2 + 2) * kTile + 1024 + 256
                   : dual     ? (1 + 2 * 2 + 2) * kTile + 256
                              : (1 + 2 * kKV + 4) * kTile + 1024 + 256;
  dim3 grid((a->L + 127) / 128, a->heads, a->B);
  static bool attr[6] = {false, false, false, false, false, false};
#define FDX_ATTN_FWD_LAUNCH(DHV, SH, DU, IDX)                                                                     \
  {                                                                                                               \
    if (!attr[IDX]) {                                                                                             \
      FDX_CUDA(cudaFuncSetAttribute(fdx_attn_fwd_kernel<DHV, SH, DU>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    smem));                                                                       \
      attr[IDX] = true;                                                                                           \
    }                                                                                                             \
    fdx_attn_fwd_kernel<DHV, SH, DU><<<grid, kThreads, smem, (cudaStream_t)stream>>>(mQ, mK, mV, d);               \
  }
  if (a->dh == 64) {
    if (short_keys) FDX_ATTN_FWD_LAUNCH(64, true, false, 0)
    else if (dual) FDX_ATTN_FWD_LAUNCH(64, false, true, 4)
    else FDX_ATTN_FWD_LAUNCH(64, false, false, 1)
  } else {
    if (short_keys) FDX_ATTN_FWD_LAUNCH(32, true, false, 2)
    else if (dual) FDX_ATTN_FWD_LAUNCH(32, false, true, 5)
    else FDX_ATTN_FWD_LAUNCH(32, false, false, 3)
  }
#undef FDX_ATTN_FWD_LAUNCH
  fdx_note_kernel(FDX_KERNEL_ATTN_FWD);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

int fdx_attention_bwd(const fdx_attn_desc* a, void* stream) {
  int s = check_desc(a, "attention_bwd");
  if (s != FDX_OK) return s;
  FDX_REQUIRE(a->o && a->d_o && a->lse && a->dvec_ws && a->dq && a->dk && a->dv, "attention_bwd: null pointer");
  const int hd = a->heads * a->dh;
  CUtensorMap mQ, mK, mV, mDO;
  if ((s = make_map(&mQ, a->q, hd, a->L, a->B, a->q_ld, a->q_bs)) != FDX_OK) return s;
  if ((s = make_map(&mK, a->k, hd, a->Lk, a->B, a->k_ld, a->k_bs)) != FDX_OK) return s;
  if ((s = make_map(&mV, a->v, hd, a->Lk, a->B, a->v_ld, a->v_bs)) != FDX_OK) return s;
  if ((s = make_map(&mDO, a->d_o, hd, a->L, a->B, a->do_ld, a->do_bs)) != FDX_OK) return s;
  AttnDev d{};
  d.B = a->B; d.heads = a->heads; d.L = a->L; d.Lk = a->Lk; d.dh = a->dh;
  d.scale = a->scale; d.scale_log2 = a->scale * kLog2e;
  d.lse = a->lse; d.dvec = a->dvec_ws;
  d.d_o = (const __nv_bfloat16*)a->d_o; d.do_ld = a->do_ld; d.do_bs = a->do_bs;
  d.o_in = (const __nv_bfloat16*)a->o; d.oin_ld = a->o_ld; d.oin_bs = a->o_bs;
  d.dq = (__nv_bfloat16*)a->dq; d.dq_ld = a->dq_ld; d.dq_bs = a->dq_bs;
  d.dk = (__nv_bfloat16*)a->dk; d.dk_ld = a->dk_ld; d.dk_bs = a->dk_bs;
  d.dv = (__nv_bfloat16*)a->dv; d.dv_ld = a->dv_ld; d.dv_bs = a->dv_bs;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const long long total = (long long)a->B * a->L * a->heads;
    attn_dvec_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d);
    FDX_LAUNCH_CHECK();
  }
  const int smem = (2 + 2 * 2 + 4) * kTile + 2 * 256 * 4 + 1024 + 256;
  static bool attr[8] = {false, false, false, false, false, false, false, false};
  const bool one_group = getenv("FDX_ATTN_BWD_EWG1") != nullptr;   // round-2a arrangement: four element-wise warps
#define FDX_ATTN_BWD_LAUNCH1(DHV, RK, EW, IDX, GRIDX, M1, M2, M3, M4)                                             \
  {                                                                                                                \
    if (!attr[IDX]) {                                                                                              \
      FDX_CUDA(cudaFuncSetAttribute(fdx_attn_bwd_kernel<DHV, RK, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    smem));                                                                        \
      attr[IDX] = true;                                                                                            \
    }                                                                                                              \
    dim3 grid((GRIDX + 127) / 128, a->heads, a->B);                                                                \
    fdx_attn_bwd_kernel<DHV, RK, EW><<<grid, 64 + 128 * EW, smem, st>>>(M1, M2, M3, M4, d);                        \
    fdx_note_kernel(FDX_KERNEL_ATTN_BWD);                                                                          \
    FDX_LAUNCH_CHECK();                                                                                            \
  }
#define FDX_ATTN_BWD_LAUNCH(DHV, RK, IDX, GRIDX, M1, M2, M3, M4)                                                  \
  if (one_group) FDX_ATTN_BWD_LAUNCH1(DHV, RK, 1, IDX, GRIDX, M1, M2, M3, M4)                                      \
  else FDX_ATTN_BWD_LAUNCH1(DHV, RK, 2, (IDX + 4), GRIDX, M1, M2, M3, M4)
  if (a->dh == 64) {
    FDX_ATTN_BWD_LAUNCH(64, false, 0, a->L, mQ, mDO, mK, mV)      // dQ: rows = queries (Q, dO), loop (K_j, V_j)
    FDX_ATTN_BWD_LAUNCH(64, true, 1, a->Lk, mK, mV, mQ, mDO)      // dK/dV: rows = keys (K, V), loop (Q_i, dO_i)
  } else {
    FDX_ATTN_BWD_LAUNCH(32, false, 2, a->L, mQ, mDO, mK, mV)
    FDX_ATTN_BWD_LAUNCH(32, true, 3, a->Lk, mK, mV, mQ, mDO)
  }
#undef FDX_ATTN_BWD_LAUNCH
#undef FDX_ATTN_BWD_LAUNCH1
  return FDX_OK;
}

}  // extern "C"

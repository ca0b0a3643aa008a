// fdx_conv3.cu -- 3x3 stride-1 convolution (forward and data gradient) with the activation halo
// staged ONCE per K chunk and shared by all nine taps.
//
// The generic tap-GEMM (fdx_tc.cu) re-fetches a 128-pixel x 64-channel A box from L2 for each of the
// nine taps.  ncu on the Cout<=128 layers of the UNet shows tensor pipe 22-29 % active with neither
// DRAM nor any SM pipe saturated: the L2->SM fabric (~11 TB/s) is the bound at 43-64 FLOP per byte.
// Here, per 64-channel chunk, TMA brings THREE boxes (one per kx shift) of (TH+2) x TW = 10 x 16
// pixels; the three ky shifts of a box are row offsets of ky*16*128 B = ky*2048 B, i.e. 1024-byte
// aligned views of the same swizzled shared memory, so each of the nine taps is a plain K-major
// UMMA descriptor.  A traffic drops from 9 to 3.75 boxes per chunk.  Weights stream through their own
// deep ring (one BN x 64 tile per tap) fed by a second producer warp.
//
// Warp roles (224 threads): 0 = A producer, 1 = MMA issuer, 2..5 = epilogue, 6 = B producer.
// Used by fdx_tc_launch for TC_KK / TC_KMN launches with the full 3x3 tap set when the geometry fits.
#include "fdx_tc.cuh"
#include "fdx_epilogue.cuh"

namespace {

constexpr int kThreads = 224;
constexpr int kTW = 16, kTH = 8;
constexpr int kABox = (kTH + 2) * kTW * 128;   // 20480 B: one kx box of a 64-channel chunk
constexpr int kAStage = 3 * kABox;             // 61440 B
constexpr int kAStages = 2;

template <int BN>
struct C3Cfg {
  static constexpr int kBTile = BN * 128;      // one tap, 64-deep: BN rows x 128 B
  static constexpr int kBStages = (BN == 64) ? 10 : (BN == 128) ? 5 : 3;
  static constexpr int kSmemBytes = kAStages * kAStage + kBStages * kBTile + 1024 + 512 + 16384;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

struct C3Dev {
  int nxb, nyb, nimg, W, H;
  int kchunks, K, Ncols, nblks, ntiles;
  int tap_b[kTcMaxTaps];
  void* out;
  long long os_x, os_y, os_n;
  const float* bias;
  const float* rowvec;
  const void* res;
  long long rs_x, rs_y, rs_n;
};

template <int BN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
fdx_conv3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const C3Dev p) {
  using Cfg = C3Cfg<BN>;
  constexpr int NB = Cfg::kBStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_b = smem + kAStages * kAStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + NB * Cfg::kBTile);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* b_full = a_empty + kAStages;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tfull = b_empty + NB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* epi_stage = smem_b + NB * Cfg::kBTile + 512;     // 16 KB

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int i = 0; i < kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&tfull[0], 1); mbar_init(&tfull[1], 1);
    mbar_init(&tempty[0], 4); mbar_init(&tempty[1], 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= A producer: 3 kx boxes per 64-channel chunk =======================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int mt = tile / p.nblks;
        const int xb = mt % p.nxb, yb = (mt / p.nxb) % p.nyb, nb = mt / (p.nxb * p.nyb);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_empty[st], ph ^ 1);
          uint8_t* sa = smem + st * kAStage;
          mbar_arrive_expect_tx(&a_full[st], kAStage);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            tma_load_4d(sa + kx * kABox, &mapA, &a_full[st], kc * 64, xb * kTW + kx - 1,
                        yb * kTH - 1, nb);
          if (++st == kAStages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ======================= B producer: one weight tile per (chunk, tap) ======================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int nt = tile % p.nblks;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_empty[st], ph ^ 1);
            uint8_t* sb = smem_b + st * Cfg::kBTile;
            mbar_arrive_expect_tx(&b_full[st], Cfg::kBTile);
            if constexpr (!B_MN) {
              tma_load_4d(sb, &mapB, &b_full[st], kc * 64, nt * BN, p.tap_b[t], 0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d(sb + j * 8192, &mapB, &b_full[st], nt * BN + j * 64, p.tap_b[t] + kc * 64,
                            0, 0);
            }
            if (++st == NB) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================== MMA issuer ============================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN, 0, B_MN ? 1 : 0);
      int ast = 0, bst = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&a_full[ast], aph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + ast * kAStage);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&b_full[bst], bph);
            tc_fence_after();
            const uint32_t sb = smem_u32(smem_b + bst * Cfg::kBTile);
            // tap t = (ky, kx): box kx, rows shifted by ky image rows (ky * 16 px * 128 B = ky * 2048 B)
            const uint32_t view = sa + (t % 3) * kABox + (t / 3) * (kTW * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_desc_sw128(view + k * 32, 16, 1024);
              const uint64_t db = B_MN ? umma_desc_sw128(sb + k * 2048, 8192, 1024)
                                       : umma_desc_sw128(sb + k * 32, 16, 1024);
              umma_f16(d_tmem, da, db, idesc, (kc | t | k) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[bst]);
            if (++bst == NB) { bst = 0; bph ^= 1; }
          }
          umma_commit(&a_empty[ast]);
          if (++ast == kAStages) { ast = 0; aph ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else {
    // =================================== epilogue ==============================================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0; uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int nt = tile % p.nblks;
      const int mt = tile / p.nblks;
      const int xb = mt % p.nxb, yb = (mt / p.nxb) % p.nyb, n = mt / (p.nxb * p.nyb);
      const int x = xb * kTW + (row % kTW), y = yb * kTH + (row / kTW);
      const bool valid = (x < p.W) && (y < p.H);
      const long long obase = (long long)n * p.os_n + (long long)y * p.os_y + (long long)x * p.os_x;
      const long long rbase =
          p.res ? (long long)n * p.rs_n + (long long)y * p.rs_y + (long long)x * p.rs_x : 0;
      auto wait_acc = [&]() {
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
      };
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      {
        EpiArgs ea{p.out, p.bias, p.rowvec, p.res, p.Ncols, 1.f, nullptr, nullptr, 0, 0};
        epilogue_bf16_coalesced<BN, EPI_PLAIN, 1>(ea, epi_stage + q * 4096, t_addr, lane, nt * BN, valid, obase, rbase,
                                           n, wait_acc);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

template <int BN, bool B_MN>
int launch_c3(const CUtensorMap& mA, const CUtensorMap& mB, const C3Dev& d, cudaStream_t stream) {
  using Cfg = C3Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    FDX_CUDA(cudaFuncSetAttribute(fdx_conv3_kernel<BN, B_MN>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  int grid = fdx_num_sms();
  if (d.ntiles < grid) grid = d.ntiles;
  fdx_conv3_kernel<BN, B_MN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(mA, mB, d);
  fdx_note_kernel(FDX_KERNEL_CONV3);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

}  // namespace

// FDX_ERR_UNSUPPORTED (no error text) = geometry not covered; caller falls back to the generic engine.
int fdx_conv3_launch(const TcLaunch& L, int BN, cudaStream_t stream) {
  if (L.mode == TC_MNMN || L.ntaps != 9 || L.es != 1 || L.gemm_like || L.b_batched) return FDX_ERR_UNSUPPORTED;
  if (L.out_f32 || L.out_atomic || L.alpha != 1.f) return FDX_ERR_UNSUPPORTED;
  if (L.W % kTW != 0 || L.H % kTH != 0) return FDX_ERR_UNSUPPORTED;
  if (BN > 192) return FDX_ERR_UNSUPPORTED;            // BN = 256 tiles are already tensor-bound
  for (int t = 0; t < 9; ++t)
    if (L.tap_dx[t] != (t % 3) - 1 || L.tap_dy[t] != (t / 3) - 1) return FDX_ERR_UNSUPPORTED;
  C3Dev d{};
  d.nxb = L.W / kTW; d.nyb = L.H / kTH; d.nimg = L.N; d.W = L.W; d.H = L.H;
  d.K = L.K; d.kchunks = (L.K + 63) / 64; d.Ncols = L.Ncols;
  d.nblks = (L.Ncols + BN - 1) / BN;
  d.ntiles = d.nxb * d.nyb * d.nimg * d.nblks;
  for (int t = 0; t < 9; ++t) d.tap_b[t] = L.tap_b[t];
  d.out = L.out; d.os_x = L.os_x; d.os_y = L.os_y; d.os_n = L.os_n;
  d.bias = L.bias; d.rowvec = L.rowvec; d.res = L.res;
  d.rs_x = L.rs_x; d.rs_y = L.rs_y; d.rs_n = L.rs_n;

  CUtensorMap mA, mB;
  {
    uint64_t dims[4], str[3];
    for (int i = 0; i < 4; ++i) dims[i] = L.A.dims[i];
    for (int i = 1; i < 4; ++i) str[i - 1] = L.A.strides[i] * 2;
    uint32_t box[4] = {64, (uint32_t)kTW, (uint32_t)(kTH + 2), 1};
    uint32_t est[4] = {1, 1, 1, 1};
    int s = fdx_make_tmap_bf16(&mA, L.A.ptr, 4, dims, str, box, est, 1);
    if (s != FDX_OK) return s;
  }
  {
    uint64_t dims[4], str[3];
    for (int i = 0; i < 4; ++i) dims[i] = L.B.dims[i];
    for (int i = 1; i < 4; ++i) str[i - 1] = L.B.strides[i] * 2;
    uint32_t box[4] = {64, (uint32_t)(L.mode == TC_KK ? BN : 64), 1, 1};
    uint32_t est[4] = {1, 1, 1, 1};
    int s = fdx_make_tmap_bf16(&mB, L.B.ptr, 4, dims, str, box, est, 1);
    if (s != FDX_OK) return s;
  }
  const bool bmn = (L.mode == TC_KMN);
  switch (BN) {
    case 64: return bmn ? launch_c3<64, true>(mA, mB, d, stream) : launch_c3<64, false>(mA, mB, d, stream);
    case 128: return bmn ? launch_c3<128, true>(mA, mB, d, stream) : launch_c3<128, false>(mA, mB, d, stream);
    case 192: return bmn ? launch_c3<192, true>(mA, mB, d, stream) : launch_c3<192, false>(mA, mB, d, stream);
  }
  return FDX_ERR_UNSUPPORTED;
}

// fdx_epilogue.cuh -- coalesced bf16 epilogue for the tcgen05 kernels.
//
// tcgen05.ld hands every thread of an epilogue warp one accumulator ROW (one output pixel).  Storing
// straight from that layout makes each 16-byte store instruction of a warp touch 32 different cache
// lines; ncu showed the epilogue warps, not the tensor pipe, bounding the short-K (Cin = 64) tiles
// (tensor pipe 22 % active, producer parked on a full ring, issuer never waiting for data).
// Here each warp stages its 32 rows x 64 channels (4 KB, XOR-swizzled 16-byte pieces, conflict free)
// in shared memory and moves it to / from global memory with 8 lanes per pixel: four complete
// 128-byte lines per instruction, for the store and for the fused residual read alike.
#pragma once
#include "fdx_common.cuh"

struct EpiArgs {
  void* out;               // bf16
  const float* bias;       // [Ncols] or null
  const float* rowvec;     // [N][Ncols] or null
  const void* res;         // bf16 or null
  int Ncols;
  float alpha;
};

__device__ __forceinline__ uint32_t epi_swz(int row, int piece) {   // byte offset inside a warp's 4 KB tile
  return (uint32_t)(row * 128 + ((piece ^ (row & 7)) << 4));
}

// One warp, its 32 accumulator rows, all BN columns of the tile.
//   stage   : this warp's 4 KB staging tile (shared memory, 128-byte aligned)
//   t_addr  : TMEM address of (first lane of this warp's quarter, first column of the accumulator)
//   col_base: first global output column of the tile (nt * BN)
//   valid / obase / rbase / img : this LANE's row: in range?, element offsets into out / res, image index
template <int BN>
__device__ __forceinline__ void epilogue_bf16_coalesced(const EpiArgs& e, uint8_t* stage, uint32_t t_addr,
                                                       int lane, int col_base, bool valid, long long obase,
                                                       long long rbase, int img) {
  const uint32_t sbase = smem_u32(stage);
  const int sub = lane >> 3, piece = lane & 7;       // coalesced phase: 4 rows x 8 pieces per instruction
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 64) {
    const int col0 = col_base + c0;
    if (col0 >= e.Ncols) break;                      // warp-uniform
    const bool half2 = (col0 + 32) < e.Ncols;        // second 32-column half present (Ncols % 32 == 0)
    // ---- residual: global -> shared, coalesced ------------------------------------------------
    if (e.res) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = 4 * k + sub;
        const long long rb = __shfl_sync(0xffffffffu, rbase, r);
        const int ok = __shfl_sync(0xffffffffu, (int)valid, r);
        uint4 u = make_uint4(0, 0, 0, 0);
        if (ok && (piece < 4 || half2))
          u = *reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(e.res) + rb + col0 + piece * 8);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + epi_swz(r, piece)), "r"(u.x),
                     "r"(u.y), "r"(u.z), "r"(u.w)
                     : "memory");
      }
      __syncwarp();
    }
    // ---- accumulator rows -> bf16 in shared -----------------------------------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !half2) break;
      uint32_t v[32];
      tmem_ld_32x32(t_addr + c0 + h * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * e.alpha;
      if (e.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
      }
      if (e.rowvec) {
        // rows outside the tensor (tile padding) must not index the per-image vector
        const float4* r4 =
            reinterpret_cast<const float4*>(e.rowvec + (long long)(valid ? img : 0) * e.Ncols + col0 + h * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(r4 + j);
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t addr = sbase + epi_swz(lane, h * 4 + j);
        if (e.res) {
          uint4 u;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                       : "r"(addr)
                       : "memory");
          const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
                       d = unpack_bf16x2(u.w);
          f[8 * j + 0] += a.x; f[8 * j + 1] += a.y; f[8 * j + 2] += b.x; f[8 * j + 3] += b.y;
          f[8 * j + 4] += c.x; f[8 * j + 5] += c.y; f[8 * j + 6] += d.x; f[8 * j + 7] += d.y;
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                     "r"(pack_bf16x2(f[8 * j + 0], f[8 * j + 1])), "r"(pack_bf16x2(f[8 * j + 2], f[8 * j + 3])),
                     "r"(pack_bf16x2(f[8 * j + 4], f[8 * j + 5])), "r"(pack_bf16x2(f[8 * j + 6], f[8 * j + 7]))
                     : "memory");
      }
    }
    __syncwarp();
    // ---- shared -> global, coalesced --------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = 4 * k + sub;
      const long long ob = __shfl_sync(0xffffffffu, obase, r);
      const int ok = __shfl_sync(0xffffffffu, (int)valid, r);
      uint4 u;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                   : "r"(sbase + epi_swz(r, piece))
                   : "memory");
      if (ok && (piece < 4 || half2))
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(e.out) + ob + col0 + piece * 8) = u;
    }
    __syncwarp();
  }
}

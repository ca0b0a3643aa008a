// fdx_norm.cu -- GroupNorm(+SiLU) and RMSNorm, forward and backward, NHWC bf16.
//
// Replaces the XLA lowering of
//   nn.GroupNorm(8, eps) + swish      flaxdiff/models/common.py:273-281,286-288,310-312
//   Unet.conv_out_norm + activation   flaxdiff/models/simple_unet.py:24-30,209-210
//   nn.RMSNorm(eps) in TransformerBlock   flaxdiff/models/attention.py:325-326
// flax semantics kept: statistics in f32, fast variance max(0, E[x^2]-E[x]^2),
// y = (x-mean)*rsqrt(var+eps)*scale+bias.
//
// HBM-bound kernels: every thread owns one 16-byte (8 x bf16) channel vector of a pixel,
// fully coalesced; per-(image,group) partial sums are reduced with warp shuffles,
// then shared-memory atomics, then one global atomic per block.
#include "fdx_common.cuh"
#include "../../include/fdx.h"

namespace {

constexpr int kNT = 256;

struct GnGeom {
  int vpp;      // 16-byte vectors per pixel = C/8
  int rows;     // pixel rows per block iteration = kNT / vpp
  int cpg;      // channels per group
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// stats[n][g] = (sum, sumsq) accumulated with atomics (buffer zeroed by the caller)
constexpr int kU = 4;   // pixels in flight per thread (memory-level parallelism)

__global__ void __launch_bounds__(kNT)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long xps, int HW, int C, int G,
                float* __restrict__ stats) {
  const int vpp = C >> 3, rows = kNT / vpp, cpg = C / G;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  __shared__ float sh[64];
  if (tid < 2 * G) sh[tid] = 0.f;
  __syncthreads();
  if (tid < rows * vpp) {
    const int cv = tid % vpp, r = tid / vpp;
    const int g = (cv * 8) / cpg;
    float s = 0.f, q = 0.f;
    const __nv_bfloat16* base = x + (long long)n * HW * xps + cv * 8;
    const int stride = gridDim.x * rows;
    int p = blockIdx.x * rows + r;
    for (; p + (kU - 1) * stride < HW; p += kU * stride) {
      uint4 u[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k)
        u[k] = *reinterpret_cast<const uint4*>(base + (long long)(p + k * stride) * xps);
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const float2 a = unpack_bf16x2(u[k].x), b = unpack_bf16x2(u[k].y), c = unpack_bf16x2(u[k].z),
                     d = unpack_bf16x2(u[k].w);
        s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
        q += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
    }
    for (; p < HW; p += stride) {
      float f[8];
      load8(base + (long long)p * xps, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s += f[j]; q += f[j] * f[j]; }
    }
    atomicAdd(&sh[2 * g], s);
    atomicAdd(&sh[2 * g + 1], q);
  }
  __syncthreads();
  if (tid < 2 * G) atomicAdd(&stats[(long long)n * 2 * G + tid], sh[tid]);
}

__global__ void __launch_bounds__(kNT)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, long long xps, int HW, int C, int G,
                const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int silu,
                __nv_bfloat16* __restrict__ y, long long yps) {
  const int vpp = C >> 3, rows = kNT / vpp, cpg = C / G;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  if (tid >= rows * vpp) return;
  const int cv = tid % vpp, r = tid / vpp;
  const int g = (cv * 8) / cpg;
  const float cnt = (float)HW * (float)cpg;
  const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
  const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
  const float rstd = rsqrtf(var + eps);
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ga = gamma[cv * 8 + j];
    a[j] = rstd * ga;
    b[j] = beta[cv * 8 + j] - mean * rstd * ga;
  }
  const __nv_bfloat16* xb = x + (long long)n * HW * xps + cv * 8;
  __nv_bfloat16* yb = y + (long long)n * HW * yps + cv * 8;
  const int stride = gridDim.x * rows;
  int p = blockIdx.x * rows + r;
  for (; p + (kU - 1) * stride < HW; p += kU * stride) {
    float f[kU][8];
#pragma unroll
    for (int k = 0; k < kU; ++k) load8(xb + (long long)(p + k * stride) * xps, f[k]);
#pragma unroll
    for (int k = 0; k < kU; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float z = f[k][j] * a[j] + b[j];
        f[k][j] = silu ? silu_f(z) : z;
      }
      store8(yb + (long long)(p + k * stride) * yps, f[k]);
    }
  }
  for (; p < HW; p += stride) {
    float f[8];
    load8(xb + (long long)p * xps, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float z = f[j] * a[j] + b[j];
      f[j] = silu ? silu_f(z) : z;
    }
    store8(yb + (long long)p * yps, f);
  }
}

// ---- GroupNorm(+SiLU) backward ---------------------------------------------------------------
// With a_c = rstd*gamma_c, b_c = beta_c - mean*a_c:  z = a_c*x + b_c,  dz = dy * silu'(z).
// Pass 1 needs only two per-(image, channel) sums, S0 = sum dz and S1 = sum dz*x, because
//   dbeta_c  = sum_n S0                       dgamma_c = sum_n rstd*(S1 - mean*S0)
//   red[n][g] = ( sum_{c in g} gamma_c*S0 ,   sum_{c in g} gamma_c*rstd*(S1 - mean*S0) )
// and pass 2 is  dx = dz*a_c - x*c2 + c3  with per-group c2 = rstd^2*m2, c3 = mean*c2 - rstd*m1.
// Both passes are instruction-issue sensitive (two MUFU per element), hence the reduced algebra.
__device__ __forceinline__ float silu_grad_times(float dy, float z) {
  const float sg = sigmoid_fast(z);
  return dy * sg * (1.f + z * (1.f - sg));
}

__global__ void __launch_bounds__(kNT)
gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ x, long long xps,
                    const __nv_bfloat16* __restrict__ dy, long long dps, int HW, int C, int G,
                    const float* __restrict__ stats, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, int silu, float* __restrict__ ws) {
  const int vpp = C >> 3, rows = kNT / vpp, cpg = C / G;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  extern __shared__ float shm[];   // 2 * [rows*C]
  float* part0 = shm;
  float* part1 = part0 + rows * C;
  const float cnt = (float)HW * (float)cpg;
  if (tid < rows * vpp) {
    const int cv = tid % vpp, r = tid / vpp;
    const int g = (cv * 8) / cpg;
    const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
    const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
    const float rstd = rsqrtf(var + eps);
    float a[8], b[8], s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = rstd * gamma[cv * 8 + j];
      b[j] = beta[cv * 8 + j] - mean * a[j];
      s0[j] = 0.f; s1[j] = 0.f;
    }
    const __nv_bfloat16* xb = x + (long long)n * HW * xps + cv * 8;
    const __nv_bfloat16* db_ = dy + (long long)n * HW * dps + cv * 8;
    const int stride = gridDim.x * rows;
    for (int p0 = blockIdx.x * rows + r; p0 < HW; p0 += 2 * stride) {
      float f[2][8], d[2][8];
      const bool two = (p0 + stride) < HW;
      load8(xb + (long long)p0 * xps, f[0]);
      load8(db_ + (long long)p0 * dps, d[0]);
      if (two) {
        load8(xb + (long long)(p0 + stride) * xps, f[1]);
        load8(db_ + (long long)(p0 + stride) * dps, d[1]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dz = silu ? silu_grad_times(d[k][j], f[k][j] * a[j] + b[j]) : d[k][j];
          s0[j] += dz;
          s1[j] += dz * f[k][j];
        }
      }
    }
    // conflict-free partials: part[r][c] (rows x C floats = 8 KB per quantity), then a column sum
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part0[r * C + cv * 8 + j] = s0[j];
      part1[r * C + cv * 8 + j] = s1[j];
    }
  }
  __syncthreads();
  // per-(image, channel) partial sums: only gridDim.x (<= ~10) atomics ever hit one address
  for (int c = tid; c < C; c += kNT) {
    float t0 = 0.f, t1 = 0.f;
    for (int rr = 0; rr < rows; ++rr) { t0 += part0[rr * C + c]; t1 += part1[rr * C + c]; }
    atomicAdd(&ws[((long long)n * C + c) * 2], t0);
    atomicAdd(&ws[((long long)n * C + c) * 2 + 1], t1);
  }
}

// finalize: blocks [0, cb) reduce over images -> dgamma / dbeta (accumulated into the gradient
// buffers), 32 channels x 8 image lanes per block; blocks [cb, cb + gb) reduce over the channels of
// a group -> red[n][g], one warp per (image, group).
__global__ void __launch_bounds__(256)
gn_bwd_finalize_kernel(const float* __restrict__ ws, const float* __restrict__ stats,
                       const float* __restrict__ gamma, int N, int HW, int C, int G, float eps, int cb,
                       float* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int cpg = C / G;
  const float cnt = (float)HW * (float)cpg;
  __shared__ float sh0[8][33], sh1[8][33];
  if ((int)blockIdx.x < cb) {
    const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float a0 = 0.f, a1 = 0.f;
    if (c < C) {
      const int g = c / cpg;
      for (int n = nl; n < N; n += 8) {
        const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
        const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
        const float rstd = rsqrtf(var + eps);
        const float2 s = *reinterpret_cast<const float2*>(ws + ((long long)n * C + c) * 2);
        a0 += s.x;
        a1 += rstd * (s.y - mean * s.x);
      }
    }
    sh0[nl][cl] = a0;
    sh1[nl][cl] = a1;
    __syncthreads();
    if (nl == 0 && c < C) {
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { t0 += sh0[k][cl]; t1 += sh1[k][cl]; }
      dbeta[c] += t0;
      dgamma[c] += t1;
    }
  } else {
    const int i = (blockIdx.x - cb) * 8 + (threadIdx.x >> 5);     // (n, g), one warp each
    if (i >= N * G) return;
    const int lane = threadIdx.x & 31;
    const int n = i / G, g = i % G;
    const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
    const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
    const float rstd = rsqrtf(var + eps);
    float r0 = 0.f, r1 = 0.f;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const float2 s = *reinterpret_cast<const float2*>(ws + ((long long)n * C + c) * 2);
      const float ga = gamma[c];
      r0 += ga * s.x;
      r1 += ga * rstd * (s.y - mean * s.x);
    }
    r0 = warp_sum(r0);
    r1 = warp_sum(r1);
    if (lane == 0) {
      red[(long long)n * 2 * G + 2 * g] = r0;
      red[(long long)n * 2 * G + 2 * g + 1] = r1;
    }
  }
}

// ab[n][0][c] = a = rstd*gamma_c ; ab[n][1][c] = b = beta_c - mean*a   (consumed by the fused dgrad epilogue)
__global__ void __launch_bounds__(256)
gn_coeffs_kernel(const float* __restrict__ stats, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int N, int HW, int C, int G, float eps,
                 float* __restrict__ ab) {
  const int cpg = C / G;
  const float cnt = (float)HW * (float)cpg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * C;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / C), c = (int)(i % C), g = c / cpg;
    const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
    const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
    const float a = rsqrtf(var + eps) * gamma[c];
    ab[((long long)n * 2) * C + c] = a;
    ab[((long long)n * 2 + 1) * C + c] = beta[c] - mean * a;
  }
}

// sums[n][c] = (S0, S1) <- sum over slots of wsl[slot][n][{0,1}][c]   (fused-epilogue workspace layout)
__global__ void __launch_bounds__(256)
gn_collapse_slots_kernel(const float* __restrict__ wsl, int slots, int N, int C, float* __restrict__ sums) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * C;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / C), c = (int)(i % C);
    float s0 = 0.f, s1 = 0.f;
    for (int sl = 0; sl < slots; ++sl) {
      const float* b = wsl + (((long long)sl * N + n) * 2) * C + c;
      s0 += b[0];
      s1 += b[C];
    }
    sums[i * 2] = s0;
    sums[i * 2 + 1] = s1;
  }
}

// out[c] (+)= sum_r in[r][c]; 32 columns x 8 row lanes per block
__global__ void __launch_bounds__(256)
reduce_rows_kernel(const float* __restrict__ in, int R, int C, float* __restrict__ out, int accumulate) {
  __shared__ float sh[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float a = 0.f;
  if (c < C)
    for (int r = rl; r < R; r += 8) a += in[(long long)r * C + c];
  sh[rl][cl] = a;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][cl];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// pass 2: dx (+= if accumulate); optionally the column sums of dx (csum_img[n][c], csum_tot[c]).
__global__ void __launch_bounds__(kNT)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, long long xps,
                    const __nv_bfloat16* __restrict__ dy, long long dps, int HW, int C, int G,
                    const float* __restrict__ stats, const float* __restrict__ red,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                    int silu, __nv_bfloat16* __restrict__ dx, long long dxps, int accumulate,
                    float* __restrict__ csum_img) {
  const int vpp = C >> 3, rows = kNT / vpp, cpg = C / G;
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  extern __shared__ float sh_cs[];   // [C] + [rows*C] when column sums are requested
  const bool want_cs = (csum_img != nullptr);
  if (want_cs) {
    for (int i = tid; i < C; i += kNT) sh_cs[i] = 0.f;
    __syncthreads();
  }
  if (tid < rows * vpp) {
    const int cv = tid % vpp, r = tid / vpp;
    const int g = (cv * 8) / cpg;
    const float cnt = (float)HW * (float)cpg;
    const float mean = stats[(long long)n * 2 * G + 2 * g] / cnt;
    const float var = fmaxf(0.f, stats[(long long)n * 2 * G + 2 * g + 1] / cnt - mean * mean);
    const float rstd = rsqrtf(var + eps);
    const float m1 = red[(long long)n * 2 * G + 2 * g] / cnt;
    const float m2 = red[(long long)n * 2 * G + 2 * g + 1] / cnt;
    const float c2 = rstd * rstd * m2;
    const float c3 = mean * c2 - rstd * m1;
    float a[8], b[8], cs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = rstd * gamma[cv * 8 + j];
      b[j] = beta[cv * 8 + j] - mean * a[j];
      cs[j] = 0.f;
    }
    const __nv_bfloat16* xb = x + (long long)n * HW * xps + cv * 8;
    const __nv_bfloat16* db_ = dy + (long long)n * HW * dps + cv * 8;
    __nv_bfloat16* ob = dx + (long long)n * HW * dxps + cv * 8;
    const int stride = gridDim.x * rows;
    for (int p0 = blockIdx.x * rows + r; p0 < HW; p0 += 2 * stride) {
      float f[2][8], d[2][8], o[2][8];
      const bool two = (p0 + stride) < HW;
      load8(xb + (long long)p0 * xps, f[0]);
      load8(db_ + (long long)p0 * dps, d[0]);
      if (accumulate) load8(ob + (long long)p0 * dxps, o[0]);
      if (two) {
        load8(xb + (long long)(p0 + stride) * xps, f[1]);
        load8(db_ + (long long)(p0 + stride) * dps, d[1]);
        if (accumulate) load8(ob + (long long)(p0 + stride) * dxps, o[1]);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k == 1 && !two) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = f[k][j];
          const float dz = silu ? silu_grad_times(d[k][j], xv * a[j] + b[j]) : d[k][j];
          const float v = dz * a[j] + (c3 - xv * c2);
          cs[j] += v;
          o[k][j] = accumulate ? o[k][j] + v : v;
        }
        store8(ob + (long long)(p0 + k * stride) * dxps, o[k]);
      }
    }
    if (want_cs) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sh_cs[C + r * C + cv * 8 + j] = cs[j];
    }
  }
  if (want_cs) {
    __syncthreads();
    for (int c = tid; c < C; c += kNT) {
      float v = 0.f;
      for (int rr = 0; rr < rows; ++rr) v += sh_cs[C + rr * C + c];
      atomicAdd(&csum_img[(long long)n * C + c], v);
    }
  }
}

// ---------------------------------------------------------------------------
// RMSNorm over channels, one warp per pixel.
// ---------------------------------------------------------------------------
constexpr int kRmsV = 4;   // up to 4 x 32 x 8 = 1024 channels, one warp per pixel

__global__ void __launch_bounds__(256)
rms_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long xps, long long npix, int C,
               const float* __restrict__ scale, float eps, __nv_bfloat16* __restrict__ y,
               long long yps) {
  const int lane = threadIdx.x & 31;
  const int nv = C >> 3;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * 8;
  for (long long p = wid; p < npix; p += nw) {
    float f[kRmsV][8];
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < kRmsV; ++v) {
      if (v * 32 + lane < nv) {
        load8(x + p * xps + (v * 32 + lane) * 8, f[v]);
#pragma unroll
        for (int j = 0; j < 8; ++j) q += f[v][j] * f[v][j];
      }
    }
    q = warp_sum(q);
    const float r = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int v = 0; v < kRmsV; ++v) {
      if (v * 32 + lane < nv) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = f[v][j] * r * scale[(v * 32 + lane) * 8 + j];
        store8(y + p * yps + (v * 32 + lane) * 8, o);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
rms_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long xps,
               const __nv_bfloat16* __restrict__ dy, long long dps, long long npix, int C,
               const float* __restrict__ scale, float eps, __nv_bfloat16* __restrict__ dx,
               long long dxps, int accumulate, float* __restrict__ dscale) {
  extern __shared__ float sh_ds[];   // [C]
  for (int i = threadIdx.x; i < C; i += 256) sh_ds[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int nv = C >> 3;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * 8;
  float ds[kRmsV][8];
#pragma unroll
  for (int v = 0; v < kRmsV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) ds[v][j] = 0.f;
  for (long long p = wid; p < npix; p += nw) {
    float f[kRmsV][8], d[kRmsV][8];
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < kRmsV; ++v) {
      if (v * 32 + lane < nv) {
        load8(x + p * xps + (v * 32 + lane) * 8, f[v]);
        load8(dy + p * dps + (v * 32 + lane) * 8, d[v]);
#pragma unroll
        for (int j = 0; j < 8; ++j) q += f[v][j] * f[v][j];
      }
    }
    q = warp_sum(q);
    const float r = rsqrtf(q / (float)C + eps);
    float m = 0.f;
#pragma unroll
    for (int v = 0; v < kRmsV; ++v) {
      if (v * 32 + lane < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = f[v][j] * r;
          ds[v][j] += d[v][j] * xh;
          d[v][j] *= scale[(v * 32 + lane) * 8 + j];   // g = dy * scale
          m += d[v][j] * xh;
        }
      }
    }
    m = warp_sum(m) / (float)C;
#pragma unroll
    for (int v = 0; v < kRmsV; ++v) {
      if (v * 32 + lane < nv) {
        float o[8];
        if (accumulate) load8(dx + p * dxps + (v * 32 + lane) * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float val = r * (d[v][j] - f[v][j] * r * m);
          o[j] = accumulate ? o[j] + val : val;
        }
        store8(dx + p * dxps + (v * 32 + lane) * 8, o);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < kRmsV; ++v)
    if (v * 32 + lane < nv)
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sh_ds[(v * 32 + lane) * 8 + j], ds[v][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += 256) atomicAdd(&dscale[i], sh_ds[i]);
}

int gn_check(const fdx_act* x, int groups, const char* what) {
  FDX_REQUIRE(x && x->ptr, "%s: null tensor", what);
  FDX_REQUIRE(groups > 0 && groups <= 32 && x->c % groups == 0, "%s: bad group count %d", what, groups);
  FDX_REQUIRE((x->c / groups) % 8 == 0, "%s: channels per group (%d) must be a multiple of 8", what,
              x->c / groups);
  FDX_REQUIRE(x->c / 8 <= kNT, "%s: C=%d too large", what, x->c);
  FDX_REQUIRE(x->pix_stride % 8 == 0, "%s: pix_stride must be a multiple of 8", what);
  return FDX_OK;
}

dim3 gn_grid(const fdx_act* x, int unroll) {
  const int HW = x->h * x->w;
  const int rows = kNT / (x->c / 8);
  int bx = (HW + rows * unroll - 1) / (rows * unroll);
  // ~16 resident blocks per SM keeps enough 16-byte loads in flight to saturate HBM
  int target = (16 * 148 + x->n - 1) / x->n;
  if (target < 1) target = 1;
  if (bx > target) bx = target;
  if (bx < 1) bx = 1;
  return dim3(bx, x->n);
}

}  // namespace

extern "C" {

int fdx_groupnorm_stats(const fdx_act* x, int groups, float* stats, void* stream) {
  int s = gn_check(x, groups, "groupnorm_stats");
  if (s != FDX_OK) return s;
  cudaStream_t st = (cudaStream_t)stream;
  FDX_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * x->n, st));
  gn_stats_kernel<<<gn_grid(x, kU), kNT, 0, st>>>((const __nv_bfloat16*)x->ptr, x->pix_stride,
                                              x->h * x->w, x->c, groups, stats);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

int fdx_groupnorm_apply(const fdx_act* x, int groups, const float* stats, const float* gamma,
                        const float* beta, float eps, int silu, const fdx_act* y, void* stream) {
  int s = gn_check(x, groups, "groupnorm_apply");
  if (s != FDX_OK) return s;
  FDX_REQUIRE(y && y->ptr && y->n == x->n && y->h == x->h && y->w == x->w && y->c == x->c,
              "groupnorm_apply: output shape mismatch");
  gn_apply_kernel<<<gn_grid(x, kU), kNT, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->pix_stride, x->h * x->w, x->c, groups, stats, gamma, beta,
      eps, silu, (__nv_bfloat16*)y->ptr, y->pix_stride);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

int fdx_groupnorm_bwd(const fdx_act* x, const fdx_act* dy, int groups, const float* stats,
                      const float* gamma, const float* beta, float eps, int silu, float* ws,
                      float* dgamma, float* dbeta, const fdx_act* dx, int accumulate,
                      float* csum_img, float* csum_tot, void* stream) {
  int s = gn_check(x, groups, "groupnorm_bwd");
  if (s != FDX_OK) return s;
  FDX_REQUIRE(dy && dy->ptr && dx && dx->ptr && ws, "groupnorm_bwd: null tensor");
  FDX_REQUIRE(dy->c == x->c && dx->c == x->c && dy->n == x->n && dx->n == x->n &&
                  dy->h == x->h && dy->w == x->w && dx->h == x->h && dx->w == x->w,
              "groupnorm_bwd: shape mismatch");
  FDX_REQUIRE(!csum_tot || csum_img, "groupnorm_bwd: csum_tot needs csum_img");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = x->n, C = x->c, HW = x->h * x->w;
  float* sums = ws;                         // [N][C][2]
  float* red = ws + 2LL * N * C;            // [N][G][2]
  FDX_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * N * C, st));
  const size_t shm = sizeof(float) * 2 * (kNT / (C / 8)) * C;
  gn_bwd_stats_kernel<<<gn_grid(x, 2), kNT, shm, st>>>(
      (const __nv_bfloat16*)x->ptr, x->pix_stride, (const __nv_bfloat16*)dy->ptr, dy->pix_stride, HW, C,
      groups, stats, gamma, beta, eps, silu, sums);
  FDX_LAUNCH_CHECK();
  const int cb = (C + 31) / 32, gb = (N * groups + 7) / 8;
  gn_bwd_finalize_kernel<<<cb + gb, 256, 0, st>>>(sums, stats, gamma, N, HW, C, groups, eps, cb, red, dgamma,
                                                   dbeta);
  FDX_LAUNCH_CHECK();
  if (csum_img) FDX_CUDA(cudaMemsetAsync(csum_img, 0, sizeof(float) * N * C, st));
  const size_t shm2 = csum_img ? sizeof(float) * (C + (kNT / (C / 8)) * C) : 0;
  gn_bwd_apply_kernel<<<gn_grid(x, 2), kNT, shm2, st>>>(
      (const __nv_bfloat16*)x->ptr, x->pix_stride, (const __nv_bfloat16*)dy->ptr, dy->pix_stride, HW, C,
      groups, stats, red, gamma, beta, eps, silu, (__nv_bfloat16*)dx->ptr, dx->pix_stride, accumulate,
      csum_img);
  FDX_LAUNCH_CHECK();
  if (csum_tot) {
    reduce_rows_kernel<<<(C + 31) / 32, 256, 0, st>>>(csum_img, N, C, csum_tot, 0);
    FDX_LAUNCH_CHECK();
  }
  return FDX_OK;
}

int fdx_groupnorm_coeffs(const float* stats, const float* gamma, const float* beta, int N, int HW, int C,
                         int groups, float eps, float* ab, void* stream) {
  FDX_REQUIRE(stats && gamma && beta && ab && N > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0,
              "groupnorm_coeffs: bad arguments");
  long long grid = ((long long)N * C + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  gn_coeffs_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(stats, gamma, beta, N, HW, C, groups, eps, ab);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

int fdx_groupnorm_bwd_dz(const fdx_act* x, const fdx_act* dz, int groups, const float* stats,
                         const float* gamma, float eps, const float* ws_slots, int slots, float* ws,
                         float* dgamma, float* dbeta, const fdx_act* dx, int accumulate, float* csum_img,
                         float* csum_tot, void* stream) {
  int s = gn_check(x, groups, "groupnorm_bwd_dz");
  if (s != FDX_OK) return s;
  FDX_REQUIRE(dz && dz->ptr && dx && dx->ptr && ws && ws_slots && slots > 0 && dgamma && dbeta,
              "groupnorm_bwd_dz: null tensor");
  FDX_REQUIRE(dz->c == x->c && dx->c == x->c && dz->n == x->n && dx->n == x->n && dz->h == x->h &&
                  dz->w == x->w && dx->h == x->h && dx->w == x->w,
              "groupnorm_bwd_dz: shape mismatch");
  FDX_REQUIRE(!csum_tot || csum_img, "groupnorm_bwd_dz: csum_tot needs csum_img");
  cudaStream_t st = (cudaStream_t)stream;
  const int N = x->n, C = x->c, HW = x->h * x->w;
  float* sums = ws;                         // [N][C][2]
  float* red = ws + 2LL * N * C;            // [N][G][2]
  long long cg = ((long long)N * C + 255) / 256;
  if (cg > 148 * 8) cg = 148 * 8;
  gn_collapse_slots_kernel<<<(int)cg, 256, 0, st>>>(ws_slots, slots, N, C, sums);
  FDX_LAUNCH_CHECK();
  const int cb = (C + 31) / 32, gb = (N * groups + 7) / 8;
  gn_bwd_finalize_kernel<<<cb + gb, 256, 0, st>>>(sums, stats, gamma, N, HW, C, groups, eps, cb, red, dgamma,
                                                   dbeta);
  FDX_LAUNCH_CHECK();
  if (csum_img) FDX_CUDA(cudaMemsetAsync(csum_img, 0, sizeof(float) * N * C, st));
  const size_t shm2 = csum_img ? sizeof(float) * (C + (kNT / (C / 8)) * C) : 0;
  // dz already carries silu'(z): the second pass is the activation-free one (beta is not read)
  gn_bwd_apply_kernel<<<gn_grid(x, 2), kNT, shm2, st>>>(
      (const __nv_bfloat16*)x->ptr, x->pix_stride, (const __nv_bfloat16*)dz->ptr, dz->pix_stride, HW, C,
      groups, stats, red, gamma, gamma, eps, 0, (__nv_bfloat16*)dx->ptr, dx->pix_stride, accumulate, csum_img);
  FDX_LAUNCH_CHECK();
  if (csum_tot) {
    reduce_rows_kernel<<<(C + 31) / 32, 256, 0, st>>>(csum_img, N, C, csum_tot, 0);
    FDX_LAUNCH_CHECK();
  }
  return FDX_OK;
}

int fdx_rmsnorm_fwd(const fdx_act* x, const float* scale, float eps, const fdx_act* y,
                    void* stream) {
  FDX_REQUIRE(x && x->ptr && y && y->ptr && scale, "rmsnorm_fwd: null tensor");
  FDX_REQUIRE(x->c % 8 == 0 && x->c <= 1024, "rmsnorm_fwd: C=%d must be a multiple of 8, <= 1024", x->c);
  const long long npix = (long long)x->n * x->h * x->w;
  int grid = (int)((npix + 7) / 8);
  if (grid > 148 * 8) grid = 148 * 8;
  rms_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x->ptr, x->pix_stride, npix,
                                                         x->c, scale, eps, (__nv_bfloat16*)y->ptr,
                                                         y->pix_stride);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

int fdx_rmsnorm_bwd(const fdx_act* x, const fdx_act* dy, const float* scale, float eps,
                    const fdx_act* dx, int accumulate, float* dscale, void* stream) {
  FDX_REQUIRE(x && x->ptr && dy && dy->ptr && dx && dx->ptr && scale && dscale, "rmsnorm_bwd: null tensor");
  FDX_REQUIRE(x->c % 8 == 0 && x->c <= 1024, "rmsnorm_bwd: C=%d must be a multiple of 8, <= 1024", x->c);
  const long long npix = (long long)x->n * x->h * x->w;
  int grid = (int)((npix + 7) / 8);
  if (grid > 148 * 2) grid = 148 * 2;
  rms_bwd_kernel<<<grid, 256, sizeof(float) * x->c, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x->ptr, x->pix_stride, (const __nv_bfloat16*)dy->ptr, dy->pix_stride, npix, x->c,
      scale, eps, (__nv_bfloat16*)dx->ptr, dx->pix_stride, accumulate, dscale);
  FDX_LAUNCH_CHECK();
  return FDX_OK;
}

}  // extern "C"

/* fdx_harness.c -- libfdx driven from PLAIN C, exactly as an XLA FFI handler (or any non-Python host) would:
 * raw device pointers wrapped in fdx_act, a caller-owned stream, int status + fdx_last_error().
 * It is the compiled counterpart of the shim sketched in INTEGRATION.md section 2 for
 * flax nn.Conv 3x3 (flaxdiff/models/common.py:166-172): y = conv3x3_SAME(x, w_hwio) + bias.
 * The expected values come from a scalar loop in this file (SAME padding, f32 accumulation).
 *   build: see tests/c_harness/Makefile       run: tests/c_harness/fdx_harness   (prints HARNESS OK) */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fdx.h"

static uint16_t f2bf(float f) { /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

#define CK(call)                                                             \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e_), #call); \
      return 2;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  const int N = 2, H = 16, W = 16, CIN = 64, COUT = 128;
  const size_t nx = (size_t)N * H * W * CIN, nw = (size_t)9 * CIN * COUT, ny = (size_t)N * H * W * COUT;
  uint16_t* hx = malloc(nx * 2);
  uint16_t* hw = malloc(nw * 2);
  uint16_t* hy = malloc(ny * 2);
  float* hb = malloc(COUT * 4);
  float* fx = malloc(nx * 4);
  float* fw = malloc(nw * 4);
  uint32_t s = 12345u;
  for (size_t i = 0; i < nx; ++i) { s = s * 1664525u + 1013904223u; fx[i] = (float)((int)(s >> 24) - 128) / 64.f; hx[i] = f2bf(fx[i]); fx[i] = bf2f(hx[i]); }
  for (size_t i = 0; i < nw; ++i) { s = s * 1664525u + 1013904223u; fw[i] = (float)((int)(s >> 24) - 128) / 1024.f; hw[i] = f2bf(fw[i]); fw[i] = bf2f(hw[i]); }
  for (int i = 0; i < COUT; ++i) hb[i] = 0.01f * (float)(i % 7);

  printf("libfdx version %d, %d SMs\n", fdx_version(), fdx_device_sm_count());
  void *dx, *dw, *dy, *db;
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  CK(cudaMalloc(&dx, nx * 2)); CK(cudaMalloc(&dw, nw * 2)); CK(cudaMalloc(&dy, ny * 2)); CK(cudaMalloc(&db, COUT * 4));
  CK(cudaMemcpyAsync(dx, hx, nx * 2, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(dw, hw, nw * 2, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(db, hb, COUT * 4, cudaMemcpyHostToDevice, st));

  /* what the FFI handler does with its buffers: */
  fdx_act ax = {dx, N, H, W, CIN, CIN};
  fdx_act ay = {dy, N, H, W, COUT, COUT};
  int rc = fdx_conv3x3_fwd(&ax, dw, (const float*)db, NULL, NULL, &ay, 1, (void*)st);
  if (rc != FDX_OK) { fprintf(stderr, "fdx_conv3x3_fwd failed: %d %s\n", rc, fdx_last_error()); return 3; }
  CK(cudaMemcpyAsync(hy, dy, ny * 2, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  double max_err = 0.0, max_ref = 0.0;
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int co = 0; co < COUT; ++co) {
          float acc = hb[co];
          for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx) {
              const int iy = y + ky - 1, ix = x + kx - 1;
              if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
              const float* xp = fx + (((size_t)n * H + iy) * W + ix) * CIN;
              const float* wp = fw + ((size_t)(ky * 3 + kx) * CIN) * COUT + co;
              for (int ci = 0; ci < CIN; ++ci) acc += xp[ci] * wp[(size_t)ci * COUT];
            }
          const float got = bf2f(hy[(((size_t)n * H + y) * W + x) * COUT + co]);
          const double e = fabs((double)got - (double)acc);
          if (e > max_err) max_err = e;
          if (fabs(acc) > max_ref) max_ref = fabs(acc);
        }
  printf("conv3x3_fwd: max |err| %.5f of max |ref| %.3f\n", max_err, max_ref);
  if (!(max_err <= 0.01 * max_ref)) { fprintf(stderr, "HARNESS FAIL\n"); return 4; }

  /* error behaviour: status < 0 and a message, never a silent fallback */
  fdx_act bad = {dx, N, H, W, 3, 8};
  rc = fdx_conv3x3_fwd(&bad, dw, NULL, NULL, NULL, &ay, 1, (void*)st);
  if (rc >= 0 || strlen(fdx_last_error()) == 0) { fprintf(stderr, "expected an error status\n"); return 5; }
  printf("rejected bad input with status %d: %s\n", rc, fdx_last_error());
  printf("kernels launched by libfdx: %llu (last kind: %s)\n", fdx_launch_count(),
         fdx_kernel_kind_name(fdx_last_kernel_kind()));
  cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(db);
  cudaStreamDestroy(st);
  printf("HARNESS OK\n");
  return 0;
}

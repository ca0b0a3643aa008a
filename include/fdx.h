/* fdx.h -- C-ABI of libfdx.so, the B200 (sm_100a) kernel library under the
 * flaxdiff_b200 Python surface.
 *
 * The reference (AshishKumar4/FlaxDiff) has no FFI / plugin interface: every device
 * op on its UNet hot path is a flax.linen / jax.numpy call lowered by XLA.  Each entry
 * point below names the reference call site (file:line under /root/reference) whose
 * XLA lowering it replaces.  Conventions:
 *   - plain pointers + sizes, caller-owned DEVICE buffers, no hidden allocation;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return 0 on success, <0 on error (fdx_last_error() gives the message);
 *   - activations are NHWC bf16 unless stated, described by fdx_act (a pixel stride
 *     larger than c lets producers write straight into channel slots of a concat
 *     buffer: jnp.concatenate at flaxdiff/models/simple_unet.py:145,196 is free);
 *   - conv kernels are HWIO [3][3][Cin][Cout] exactly as flax stores them.
 */
#ifndef FDX_H_
#define FDX_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDX_OK 0
#define FDX_ERR_INVALID_ARG (-1)
#define FDX_ERR_UNSUPPORTED (-2)
#define FDX_ERR_CUDA (-3)
#define FDX_ERR_NO_DEVICE (-4)

typedef struct {
  void* ptr;            /* device pointer to element (n=0,y=0,x=0,c=0) */
  int n, h, w, c;       /* logical dims */
  long long pix_stride; /* elements between consecutive pixels (>= c, multiple of 8) */
} fdx_act;

/* ---- library ---------------------------------------------------------------- */
const char* fdx_last_error(void);
int fdx_version(void);
int fdx_device_sm_count(void);

/* ---- tensor-core contractions (tcgen05 + TMA + TMEM) ------------------------- */
/* flax nn.Conv 3x3 SAME, stride 1 or 2 (models/common.py:166-172, 237-244).
 * y = conv(x, w) + bias + rowvec[n] + res ; bias [Cout] f32, rowvec [N][Cout] f32 (the
 * ResidualBlock's `out + temb` add, common.py:300-308), res = residual add (common.py:334).
 * w_hwio is bf16.  Cin, Cout multiples of 64. */
int fdx_conv3x3_fwd(const fdx_act* x, const void* w_hwio, const float* bias, const float* rowvec,
                    const fdx_act* res, const fdx_act* y, int stride, void* stream);
/* d(loss)/dx of the conv above (jax.value_and_grad, trainer/general_diffusion_trainer.py:321). */
int fdx_conv3x3_dgrad(const fdx_act* dy, const void* w_hwio, const fdx_act* dx, int stride,
                      int accumulate, void* stream);
/* d(loss)/dw, f32 HWIO, ACCUMULATED into dw_hwio (caller zeroes it). */
int fdx_conv3x3_wgrad(const fdx_act* x, const fdx_act* dy, float* dw_hwio, int stride,
                      void* stream);

#define FDX_GEMM_KK 0   /* D[m][n] = sum_k A[m][k] * B[n][k]   (both k-contiguous)   */
#define FDX_GEMM_KMN 1  /* D[m][n] = sum_k A[m][k] * B[k][n]                          */
#define FDX_GEMM_MNMN 2 /* D[m][n] = sum_k A[k][m] * B[k][n]   (weight-gradient type) */
typedef struct {
  int mode;
  int M, N, K;
  int batch1, batch2;      /* two batch dims (e.g. heads, images); 0/1 = none */
  const void* A; long long a_ld, a_s1, a_s2; /* bf16; leading dim and batch strides, elements */
  const void* B; long long b_ld, b_s1, b_s2; /* bf16; b_s1 = b_s2 = 0 -> B shared by the batch */
  void* D; long long d_ld, d_s1, d_s2;
  int d_f32;               /* output dtype: 0 bf16, 1 f32 */
  int d_atomic;            /* f32 atomicAdd into D (split-K allowed) */
  int reduce_batch;        /* MNMN only: also reduce over the batch dims (dense weight grads) */
  float alpha;
  const float* bias;       /* [N] or NULL */
  const void* res; long long r_ld, r_s1, r_s2; /* bf16 residual added to D, or NULL */
} fdx_gemm_desc;
/* 1x1 conv (common.py:324-333), nn.DenseGeneral (attention.py:132-154, common.py:300-305),
 * and the QK^T / PV / backward products of nn.dot_product_attention (attention.py:170-174). */
int fdx_gemm(const fdx_gemm_desc* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FDX_H_ */

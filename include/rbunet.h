/*
 * rbunet.h — C ABI of librbunet.so: the B200 (sm_100a) Robust U-Net hot path.
 *
 * The reference (UofgCoastline/EUSIPCO-2026-Robust-Unet) has no FFI: its hot path is the set of
 * PyTorch ops dispatched by `RobustUNet.forward` (Main_Final.py:290-321), `nn.BCELoss`
 * (Main_Final.py:551,580) and `ModelEvaluator.calculate_metrics` (Main_Final.py:519-547).
 * Each entry point below replaces the ATen/cuDNN dispatch of the reference lines it cites
 * (SURVEY.md §2.1 maps ATen ops to kernels).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *  - the caller allocates and frees every buffer (including workspaces); the library keeps no tensor state;
 *  - every function enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *  - return value: 0 on success, negative on error (RBU_ERR_*); rbu_last_error() gives the message;
 *    nothing throws or exits across the ABI; unsupported input is an error, there is NO CPU fallback;
 *  - activations are NHWC bf16 "views": pointer to the first channel of a channel slice, pixel stride
 *    `ld` in elements (multiple of 8), `C` channels in the slice (multiple of 8);
 *  - "pixels" P = N*H*W.
 */
#ifndef RBUNET_H_
#define RBUNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBU_OK 0
#define RBU_ERR_INVALID (-1)
#define RBU_ERR_CUDA (-2)
#define RBU_ERR_UNSUPPORTED (-3)

/* ------------------------------------------------------------------ library */
int rbu_version(void);
const char* rbu_last_error(void);
/* 0 if the current device is compute capability 10.x (tcgen05/TMEM/TMA available). */
int rbu_device_check(void);
int rbu_sm_count(void);

/* ------------------------------------------------------------------ tensor-core implicit GEMM
 * Replaces aten::convolution for 3x3 (dilation 1/2/4), 1x1 and ConvTranspose2d(2, stride 2)
 * (Main_Final.py:157,159,172,126,131,205-208,261-270) and, with re-packed weights, their data
 * gradients.  D[pixel, n] = sum over segments, taps, channels of X[pixel + tap, c] * W[n, tap, c].
 */
typedef struct {
  const void* x;   /* bf16 NHWC view of the input (for gather: the 2H x 2W tensor)            */
  int64_t x_ld;    /* pixel stride in elements                                               */
  int C;           /* channels in the slice = per-tap K extent                               */
  const void* w;   /* bf16 packed weights [Ncols][taps][C]                                   */
  int taps;        /* 1 (1x1), 9 (3x3) or 4 (gather=1: 2x2 stride-2 gather, tap = 2*i + j)   */
  int dil;         /* dilation of the 3x3 taps (padding == dilation)                         */
  int gather;      /* 1: X[pixel(n,h,w), tap(i,j)] = x[n, 2h+i, 2w+j] (ConvTranspose dgrad)  */
} rbu_gemm_operand;

typedef struct {
  int N, H, W;              /* pixel grid of the GEMM M dimension                                  */
  int nseg;                 /* 1 or 2 accumulated segments (e.g. conv1 dgrad + shortcut dgrad)     */
  rbu_gemm_operand seg[2];
  int Ncols;                /* GEMM N (Cout, or 4*Cout when scatter=1); multiple of 8              */
  void* y;                  /* bf16 NHWC output view                                               */
  int64_t y_ld;
  int scatter;              /* 1: ConvTranspose2d pixel shuffle: column (2i+j)*Cout+co of pixel    */
  int Cout;                 /*    (n,h,w) goes to y[n, 2h+i, 2w+j, co]; y is [N,2H,2W]; else 0     */
  const float* bias;        /* fp32 [Cout] or NULL                                                 */
  const void* addend;       /* optional bf16 NHWC view added to the result (scatter=0 only)        */
  int64_t addend_ld;
} rbu_conv_gemm_args;

int rbu_conv_gemm(const rbu_conv_gemm_args* args, void* stream);

/* Re-pack fp32 torch-layout weights into the bf16 GEMM operand [Nn][T][K] (K contiguous).
 *  mode 0: Conv2d weight [Nn=Cout][K=Cin][T] for the forward GEMM
 *  mode 1: Conv2d weight [K=Cout][Nn=Cin][T] for the data gradient (taps rotated by 180 degrees)
 *  mode 2: ConvTranspose2d weight [K=Cin][Cout][2][2] for the forward GEMM (Nn = 4*Cout, T = 1)
 *  mode 3: ConvTranspose2d weight [Nn=Cin][K=Cout][2][2] for the data gradient (T = 4)
 * Replaces nothing in the reference (layout change only; Main_Final.py:157-172,261-270 parameters). */
int rbu_pack_weight(const float* src, void* dst, int Nn, int T, int K, int mode, int Cout, void* stream);

/* TEST-ONLY device reference: direct (CUDA-core) convolution, bf16 NHWC in, fp32 torch-layout weights
 * (rounded to bf16 on the fly), fp32 dense NHWC out.  Used by the parity tests at sizes where the CPU
 * oracle is too slow; never called by the product path. */
int rbu_conv_direct_ref(const void* x, int64_t x_ld, int N, int H, int W, int Cin, const float* w,
                        const float* bias, int Cout, int ksz, int dil, float* out, void* stream);

/* ------------------------------------------------------------------ loss + metrics (K9)
 * rbu_loss_forward replaces nn.BCELoss()(outputs, masks) (Main_Final.py:551,580) and, in the same
 * pass, the per-image confusion counts behind ModelEvaluator.calculate_metrics (Main_Final.py:519-547).
 * loss = w_bce * BCE + w_dice * (1 - (2*sum(p*y)+smooth)/(sum(p)+sum(y)+smooth)); defaults (1,0) are
 * reference-exact.  probs/target: fp32 [B*HW]; counts: int64 [B][4] = TP,FP,FN,TN for p > threshold
 * (strict); sums_out: double[4] = sum(bce), sum(p), sum(y), sum(p*y) (kept for the backward). */
size_t rbu_loss_workspace_bytes(int B, int64_t HW);
int rbu_loss_forward(const float* probs, const float* target, int B, int64_t HW, float threshold, float w_bce,
                     float w_dice, float smooth, void* workspace, size_t workspace_bytes, float* loss_out,
                     double* sums_out, int64_t* counts, void* stream);
/* dL/dprobs with torch's binary_cross_entropy_backward formula ((p-y)/max(p(1-p),1e-12)/N), times
 * grad_out[0] (NULL = 1). */
int rbu_loss_backward(const float* probs, const float* target, int64_t total, const float* grad_out,
                      const double* sums, float w_bce, float w_dice, float smooth, float* dprobs, void* stream);
/* Integer TP/FP/FN/TN per image only (Main_Final.py:521-535). */
int rbu_confusion_counts(const float* pred, const float* target, int B, int64_t HW, float threshold,
                         int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RBUNET_H_ */

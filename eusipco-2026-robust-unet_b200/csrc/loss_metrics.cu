// K9: robust BCE(+Dice) loss and IoU/F1/Acc confusion counts, and the 1x1 `outc` head.
//   nn.BCELoss (Main_Final.py:551,580): mean(-(y*max(ln p,-100) + (1-y)*max(ln(1-p),-100)))
//   calculate_metrics (Main_Final.py:519-547): pred > threshold (strict), integer TP/FP/FN/TN per image
//   outc (Main_Final.py:274-277,321): 1x1 conv C->1 + bias + Sigmoid, kept in fp32
// All reductions are two-stage with fixed summation order (no float atomics) so results are
// run-to-run deterministic; counts use integer atomics (exact).
#include "rbu_common.cuh"

namespace {

constexpr int LOSS_THREADS = 256;

__device__ __forceinline__ void loss_accum(float p, float y, float thr, float& bce, float& sp, float& sy, float& spy,
                                           unsigned& tp, unsigned& fp, unsigned& fn) {
  const float lp = fmaxf(logf(p), -100.f);
  const float l1p = fmaxf(logf(1.f - p), -100.f);
  bce -= y * lp + (1.f - y) * l1p;
  sp += p;
  sy += y;
  spy += p * y;
  const bool pb = p > thr;
  const bool tb = y != 0.f;
  tp += (pb && tb);
  fp += (pb && !tb);
  fn += (!pb && tb);
}

// grid (nblk, B): per-image partial sums + integer counts
__global__ void __launch_bounds__(LOSS_THREADS)
loss_partial_kernel(const float* __restrict__ probs, const float* __restrict__ target, long HW, float thr,
                    float* __restrict__ partials, unsigned long long* __restrict__ counts) {
  const int b = blockIdx.y;
  const float* p = probs + (long)b * HW;
  const float* y = target + (long)b * HW;
  float bce = 0.f, sp = 0.f, sy = 0.f, spy = 0.f;
  unsigned tp = 0, fp = 0, fn = 0;
  const long nvec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 ? HW / 4 : 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < nvec; i += (long)gridDim.x * blockDim.x) {
    const float4 pv = __ldg(reinterpret_cast<const float4*>(p) + i);
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y) + i);
    loss_accum(pv.x, yv.x, thr, bce, sp, sy, spy, tp, fp, fn);
    loss_accum(pv.y, yv.y, thr, bce, sp, sy, spy, tp, fp, fn);
    loss_accum(pv.z, yv.z, thr, bce, sp, sy, spy, tp, fp, fn);
    loss_accum(pv.w, yv.w, thr, bce, sp, sy, spy, tp, fp, fn);
  }
  for (long i = nvec * 4 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x)
    loss_accum(p[i], y[i], thr, bce, sp, sy, spy, tp, fp, fn);

  __shared__ float sf[4][LOSS_THREADS / 32];
  __shared__ unsigned su[3][LOSS_THREADS / 32];
  bce = warp_sum(bce); sp = warp_sum(sp); sy = warp_sum(sy); spy = warp_sum(spy);
  tp = __reduce_add_sync(0xffffffffu, tp);
  fp = __reduce_add_sync(0xffffffffu, fp);
  fn = __reduce_add_sync(0xffffffffu, fn);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sf[0][warp] = bce; sf[1][warp] = sp; sf[2][warp] = sy; sf[3][warp] = spy;
    su[0][warp] = tp; su[1][warp] = fp; su[2][warp] = fn;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) {
      a0 += sf[0][w]; a1 += sf[1][w]; a2 += sf[2][w]; a3 += sf[3][w];
      c0 += su[0][w]; c1 += su[1][w]; c2 += su[2][w];
    }
    float* out = partials + ((long)b * gridDim.x + blockIdx.x) * 4;
    out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
    atomicAdd(&counts[b * 4 + 0], c0);
    atomicAdd(&counts[b * 4 + 1], c1);
    atomicAdd(&counts[b * 4 + 2], c2);
  }
}

// single thread: ordered double-precision sum of the partials, final loss
__global__ void loss_finalize_kernel(const float* __restrict__ partials, int B, int nblk, long HW, float w_bce,
                                     float w_dice, float smooth, float* __restrict__ loss_out,
                                     double* __restrict__ sums_out) {
  if (threadIdx.x == 0) {
    double s[4] = {0, 0, 0, 0};
    for (long i = 0; i < (long)B * nblk; ++i)
      for (int k = 0; k < 4; ++k) s[k] += (double)partials[i * 4 + k];
    const double total = (double)B * (double)HW;
    double loss = w_bce * (s[0] / total);
    if (w_dice != 0.f) {
      const double dice = (2.0 * s[3] + smooth) / (s[1] + s[2] + smooth);
      loss += w_dice * (1.0 - dice);
    }
    loss_out[0] = (float)loss;
    for (int k = 0; k < 4; ++k) sums_out[k] = s[k];
  }
}

// dL/dp: torch's binary_cross_entropy_backward, literally: (p - y) / max(p (1-p), 1e-12) * g / N,
// plus the Dice term d(1-dice)/dp_i = -(2 y_i (S+s) - (2I+s)) / (S+s)^2
__global__ void loss_backward_kernel(const float* __restrict__ probs, const float* __restrict__ target, long total,
                                     const float* __restrict__ gout, const double* __restrict__ sums, float w_bce,
                                     float w_dice, float smooth, float* __restrict__ dp) {
  const float g = gout ? gout[0] : 1.f;
  const float inv_n = 1.f / (float)total;
  float dice_a = 0.f, dice_b = 0.f;
  if (w_dice != 0.f) {
    const double den = sums[1] + sums[2] + (double)smooth;
    dice_a = (float)(2.0 / den);                                   // coefficient of y_i
    dice_b = (float)((2.0 * sums[3] + (double)smooth) / (den * den));
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float p = probs[i], y = target[i];
    float d = w_bce * ((p - y) / fmaxf(p * (1.f - p), 1e-12f)) * inv_n;
    if (w_dice != 0.f) d -= w_dice * (dice_a * y - dice_b);
    dp[i] = d * g;
  }
}

__global__ void counts_kernel(const float* __restrict__ pred, const float* __restrict__ target, long HW, float thr,
                              unsigned long long* __restrict__ counts) {
  const int b = blockIdx.y;
  const float* p = pred + (long)b * HW;
  const float* y = target + (long)b * HW;
  unsigned tp = 0, fp = 0, fn = 0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < HW; i += (long)gridDim.x * blockDim.x) {
    const bool pb = p[i] > thr;
    const bool tb = y[i] != 0.f;
    tp += (pb && tb); fp += (pb && !tb); fn += (!pb && tb);
  }
  tp = __reduce_add_sync(0xffffffffu, tp);
  fp = __reduce_add_sync(0xffffffffu, fp);
  fn = __reduce_add_sync(0xffffffffu, fn);
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(&counts[b * 4 + 0], (unsigned long long)tp);
    if (fp) atomicAdd(&counts[b * 4 + 1], (unsigned long long)fp);
    if (fn) atomicAdd(&counts[b * 4 + 2], (unsigned long long)fn);
  }
}
__global__ void counts_tn_kernel(long long* counts, int B, long HW) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) counts[b * 4 + 3] = HW - counts[b * 4] - counts[b * 4 + 1] - counts[b * 4 + 2];
}

int loss_nblk(long HW) {
  long n = (HW + 8191) / 8192;
  return (int)(n < 1 ? 1 : (n > 128 ? 128 : n));
}

}  // namespace

extern "C" size_t rbu_loss_workspace_bytes(int B, int64_t HW) { return (size_t)B * loss_nblk(HW) * 4 * sizeof(float); }

extern "C" int rbu_loss_forward(const float* probs, const float* target, int B, int64_t HW, float threshold,
                                float w_bce, float w_dice, float smooth, void* workspace, size_t workspace_bytes,
                                float* loss_out, double* sums_out, int64_t* counts, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(probs && target && loss_out && sums_out && counts && workspace, "rbu_loss_forward: null pointer");
  RBU_CHECK_ARG(B > 0 && HW > 0, "rbu_loss_forward: empty input (B=%d, HW=%lld)", B, (long long)HW);
  RBU_CHECK_ARG(workspace_bytes >= rbu_loss_workspace_bytes(B, HW), "rbu_loss_forward: workspace too small");
  RBU_CHECK_ARG(B <= 65535, "rbu_loss_forward: batch too large");
  const int nblk = loss_nblk(HW);
  RBU_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 4 * sizeof(int64_t), stream));
  loss_partial_kernel<<<dim3(nblk, B), LOSS_THREADS, 0, stream>>>(probs, target, HW, threshold, (float*)workspace,
                                                                 (unsigned long long*)counts);
  RBU_CHECK_LAUNCH();
  loss_finalize_kernel<<<1, 32, 0, stream>>>((const float*)workspace, B, nblk, HW, w_bce, w_dice, smooth, loss_out,
                                               sums_out);
  RBU_CHECK_LAUNCH();
  counts_tn_kernel<<<rbu_cdiv(B, 256), 256, 0, stream>>>((long long*)counts, B, HW);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_loss_backward(const float* probs, const float* target, int64_t total, const float* grad_out,
                                 const double* sums, float w_bce, float w_dice, float smooth, float* dprobs,
                                 void* stream_) {
  RBU_CHECK_ARG(probs && target && sums && dprobs && total > 0, "rbu_loss_backward: bad arguments");
  const int blocks = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  loss_backward_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(probs, target, total, grad_out, sums, w_bce, w_dice,
                                                                  smooth, dprobs);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_confusion_counts(const float* pred, const float* target, int B, int64_t HW, float threshold,
                                    int64_t* counts, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RBU_CHECK_ARG(pred && target && counts, "rbu_confusion_counts: null pointer");
  RBU_CHECK_ARG(B > 0 && B <= 65535 && HW > 0, "rbu_confusion_counts: empty input (B=%d, HW=%lld)", B, (long long)HW);
  RBU_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)B * 4 * sizeof(int64_t), stream));
  const int nblk = loss_nblk(HW);
  counts_kernel<<<dim3(nblk, B), 256, 0, stream>>>(pred, target, HW, threshold, (unsigned long long*)counts);
  RBU_CHECK_LAUNCH();
  counts_tn_kernel<<<rbu_cdiv(B, 256), 256, 0, stream>>>((long long*)counts, B, HW);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

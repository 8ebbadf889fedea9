// Fused multi-tensor Adam step with coupled L2 weight decay -- torch.optim.Adam(lr, betas, eps, weight_decay) as the
// reference uses it (Main_Final.py:552,582): one launch updates every parameter tensor of the model.
//   g' = g + wd*p;  m = m + (1-b1)(g' - m);  v = b2*v + (1-b2) g'^2;
//   p  = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Tensor math in fp32 in the operation order of torch's single-tensor implementation; the scalar factors (1-beta,
// lr/(1-beta1^t), sqrt(1-beta2^t)) are evaluated in double on the host and rounded to fp32 once, as torch does.
#include "rbu_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_multi_kernel(const rbu_adam_job* __restrict__ jobs, int njobs, float step_size, float one_minus_beta1, float beta2,
                  float one_minus_beta2, float eps, float weight_decay, float bias_corr2_sqrt) {
  int lo = 0, hi = njobs - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const rbu_adam_job j = jobs[lo];
  const long long base = (b - j.first_block) * 1024;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = base + u * 256 + threadIdx.x;
    if (i >= j.numel) return;
    const float p = j.param[i];
    float g = j.grad[i];
    if (weight_decay != 0.f) g = g + weight_decay * p;
    float m = j.exp_avg[i], v = j.exp_avg_sq[i];
    m = m + one_minus_beta1 * (g - m);
    v = beta2 * v + one_minus_beta2 * (g * g);
    const float denom = sqrtf(v) / bias_corr2_sqrt + eps;
    j.exp_avg[i] = m;
    j.exp_avg_sq[i] = v;
    j.param[i] = p - step_size * (m / denom);
  }
}

}  // namespace

extern "C" int rbu_adam_step(const rbu_adam_job* jobs_device, int njobs, long long total_blocks, double lr, double beta1,
                             double beta2, double eps, double weight_decay, int step, void* stream) {
  RBU_CHECK_ARG(jobs_device && njobs > 0 && total_blocks > 0 && total_blocks < (1LL << 31) && step >= 1,
                "rbu_adam_step: bad arguments");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  adam_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(
      jobs_device, njobs, (float)(lr / bc1), (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
      (float)weight_decay, (float)sqrt(bc2));
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

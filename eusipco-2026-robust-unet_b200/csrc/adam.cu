// Fused multi-tensor Adam step with coupled L2 weight decay -- torch.optim.Adam(lr, betas, eps, weight_decay) as the
// reference uses it (Main_Final.py:552,582): one launch updates every parameter tensor of the model.
//   g' = g + wd*p;  m = m + (1-b1)(g' - m);  v = b2*v + (1-b2) g'^2;
//   p  = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// All math in fp32, same operation order as torch's single-tensor implementation (results agree to ~1 ulp).
#include "rbu_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_multi_kernel(const rbu_adam_job* __restrict__ jobs, int njobs, float lr, float beta1, float beta2, float eps,
                  float weight_decay, float bias_corr1, float bias_corr2_sqrt) {
  int lo = 0, hi = njobs - 1;
  const long long b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first_block <= b) lo = mid; else hi = mid - 1;
  }
  const rbu_adam_job j = jobs[lo];
  const long long base = (b - j.first_block) * 1024;
  const float step_size = lr / bias_corr1;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const long long i = base + u * 256 + threadIdx.x;
    if (i >= j.numel) return;
    const float p = j.param[i];
    float g = j.grad[i];
    if (weight_decay != 0.f) g = g + weight_decay * p;
    float m = j.exp_avg[i], v = j.exp_avg_sq[i];
    m = m + (1.f - beta1) * (g - m);
    v = beta2 * v + (1.f - beta2) * (g * g);
    const float denom = sqrtf(v) / bias_corr2_sqrt + eps;
    j.exp_avg[i] = m;
    j.exp_avg_sq[i] = v;
    j.param[i] = p - step_size * (m / denom);
  }
}

}  // namespace

extern "C" int rbu_adam_step(const rbu_adam_job* jobs_device, int njobs, long long total_blocks, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, void* stream) {
  RBU_CHECK_ARG(jobs_device && njobs > 0 && total_blocks > 0 && total_blocks < (1LL << 31) && step >= 1,
                "rbu_adam_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_multi_kernel<<<(unsigned)total_blocks, 256, 0, (cudaStream_t)stream>>>(jobs_device, njobs, lr, beta1, beta2, eps,
                                                                             weight_decay, bc1, sqrtf(bc2));
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

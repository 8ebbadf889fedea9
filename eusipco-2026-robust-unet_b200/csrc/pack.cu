// Weight re-packing: fp32 torch-layout parameters -> bf16 GEMM operands [Ncols][taps][K] (K contiguous),
// in the forward and data-gradient forms the implicit-GEMM kernel consumes, plus the SIMT reference
// convolution used by the parity tests (test-only device reference, never on the product path).
#include "rbu_common.cuh"

namespace {

// mode 0: conv fwd   src [Nn][K][T]      -> dst[n][t][k] = src[n][k][t]
// mode 1: conv dgrad src [K][Nn][T]      -> dst[n][t][k] = src[k][n][T-1-t]        (180-degree rotated taps)
// mode 2: convT fwd  src [K][Cout][4]    -> dst[q*Cout+co][0][k] = src[k][co][q]   (Nn = 4*Cout, T = 1)
// mode 3: convT dgrad src [Nn][K][4]     -> dst[n][q][k] = src[n][k][q]            (T = 4)
__global__ void pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int Nn, int T, int K,
                                   int mode, int Cout) {
  const long total = (long)Nn * T * K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int t = (int)((i / K) % T);
    const int n = (int)(i / ((long)K * T));
    long s;
    if (mode == 0) s = ((long)n * K + k) * T + t;
    else if (mode == 1) s = ((long)k * Nn + n) * T + (T - 1 - t);
    else if (mode == 2) { const int q = n / Cout, co = n % Cout; s = ((long)k * Cout + co) * 4 + q; }
    else s = ((long)n * K + k) * 4 + t;
    dst[i] = __float2bfloat16_rn(src[s]);
  }
}

__global__ void conv_direct_ref_kernel(const bf16* __restrict__ x, long x_ld, int N, int H, int W, int Cin,
                                       const float* __restrict__ w, const float* __restrict__ bias, int Cout,
                                       int ksz, int dil, float* __restrict__ out) {
  const long total = (long)N * H * W * Cout;
  const int pad = dil * (ksz / 2);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long pix = i / Cout;
    const int wq = (int)(pix % W);
    const int hq = (int)((pix / W) % H);
    const int n = (int)(pix / ((long)W * H));
    float acc = bias ? bias[co] : 0.f;
    for (int r = 0; r < ksz; ++r) {
      const int hh = hq + r * dil - pad;
      if (hh < 0 || hh >= H) continue;
      for (int s = 0; s < ksz; ++s) {
        const int ww = wq + s * dil - pad;
        if (ww < 0 || ww >= W) continue;
        const bf16* xp = x + (((long)n * H + hh) * W + ww) * x_ld;
        const float* wp = w + ((long)co * Cin * ksz + r) * ksz + s;
        for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xp[ci]) * bf16_round(wp[(long)ci * ksz * ksz]);
      }
    }
    out[i] = acc;
  }
}

}  // namespace

extern "C" int rbu_pack_weight(const float* src, void* dst, int Nn, int T, int K, int mode, int Cout, void* stream) {
  RBU_CHECK_ARG(src && dst && Nn > 0 && T > 0 && K > 0 && mode >= 0 && mode <= 3, "rbu_pack_weight: bad arguments");
  RBU_CHECK_ARG(mode != 2 || (Cout > 0 && Nn == 4 * Cout && T == 1), "rbu_pack_weight: mode 2 needs Nn == 4*Cout, T == 1");
  RBU_CHECK_ARG(mode != 3 || T == 4, "rbu_pack_weight: mode 3 needs T == 4");
  const long total = (long)Nn * T * K;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, Nn, T, K, mode, Cout);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

extern "C" int rbu_conv_direct_ref(const void* x, int64_t x_ld, int N, int H, int W, int Cin, const float* w,
                                   const float* bias, int Cout, int ksz, int dil, float* out, void* stream) {
  RBU_CHECK_ARG(x && w && out && (ksz == 1 || ksz == 3 || ksz == 7) && dil >= 1, "rbu_conv_direct_ref: bad arguments");
  const long total = (long)N * H * W * Cout;
  const int blocks = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  conv_direct_ref_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, x_ld, N, H, W, Cin, w, bias, Cout,
                                                                    ksz, dil, out);
  RBU_CHECK_LAUNCH();
  return RBU_OK;
}

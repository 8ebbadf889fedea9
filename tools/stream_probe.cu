// Diagnostic: what a plain streaming kernel with R read streams and Wr write streams reaches on this GPU
// (16-byte vectors, grid-stride, 256 threads, U vectors in flight per stream and thread), with default and
// cache-streaming (.cs) stores.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/stream_probe.cu -o /tmp/stream_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int R, int Wr, int U, bool CS>
__global__ void __launch_bounds__(256) probe(const uint4* __restrict__ a, const uint4* __restrict__ b, const uint4* __restrict__ c,
                                             uint4* __restrict__ o1, uint4* __restrict__ o2, long n) {
  const uint4* in[3] = {a, b, c};
  uint4* out[2] = {o1, o2};
  for (long base = (long)blockIdx.x * (256 * U) + threadIdx.x; base < n; base += (long)gridDim.x * (256 * U)) {
    uint4 v[R][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (base + u * 256 < n)
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[r][u].x), "=r"(v[r][u].y), "=r"(v[r][u].z), "=r"(v[r][u].w) : "l"(in[r] + base + u * 256));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * 256 < n) {
        uint4 s = v[0][u];
#pragma unroll
        for (int r = 1; r < R; ++r) { s.x ^= v[r][u].x; s.y += v[r][u].y; s.z ^= v[r][u].z; s.w += v[r][u].w; }
#pragma unroll
        for (int w = 0; w < Wr; ++w) {
          uint4* p = out[w] + base + u * 256;
          if (CS)
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(s.x + w), "r"(s.y), "r"(s.z), "r"(s.w) : "memory");
          else
            *p = make_uint4(s.x + w, s.y, s.z, s.w);
        }
      }
    }
  }
}

static int g_smem = 0;      // dynamic shared memory per block: limits the resident blocks per SM (occupancy experiment)

template <int R, int Wr, int U, bool CS>
void run(const char* name, uint4** buf, long n, int blocks) {
  cudaFuncSetAttribute(probe<R, Wr, U, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 6; ++it) {
    cudaMemsetAsync(buf[5], it, 256 << 20);      // L2 flush
    cudaEventRecord(e0);
    probe<R, Wr, U, CS><<<blocks, 256, g_smem>>>(buf[0], buf[1], buf[2], buf[3], buf[4], n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (it >= 2 && ms < best) best = ms;
  }
  printf("%-28s blocks %6d  %7.3f ms  %7.0f GB/s\n", name, blocks, best, (double)(R + Wr) * n * 16 / best / 1e6);
}

int main() {
  const long n = (512L << 20) / 16;               // 512 MiB per stream
  uint4* buf[6];
  for (int i = 0; i < 5; ++i) { cudaMalloc(&buf[i], n * 16); cudaMemset(buf[i], i + 1, n * 16); }
  cudaMalloc(&buf[5], 256 << 20);
  int sms = 148;
  for (int resident : {8, 4, 3, 2, 1}) {
    g_smem = resident == 8 ? 0 : (220 * 1024) / resident - 2048;
    const int blocks = sms * 32;
    printf("--- at most %d resident blocks of 256 threads per SM (dynamic smem %d B)\n", resident, g_smem);
    run<1, 1, 4, false>("1R 1W U4", buf, n, blocks);
    run<2, 1, 4, false>("2R 1W U4", buf, n, blocks);
    run<2, 0, 4, false>("2R 0W U4 (reduce-like)", buf, n, blocks);
    run<3, 2, 2, false>("3R 2W U2", buf, n, blocks);
    run<3, 2, 2, true>("3R 2W U2 .cs stores", buf, n, blocks);
    run<3, 2, 4, false>("3R 2W U4", buf, n, blocks);
    run<3, 2, 4, true>("3R 2W U4 .cs stores", buf, n, blocks);
    run<3, 1, 4, true>("3R 1W U4 .cs stores", buf, n, blocks);
  }
  return 0;
}

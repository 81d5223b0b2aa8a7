// Micro-benchmark behind profiles/r1_summary.md section 4: what does it cost a warp on B200 to run code that no
// other warp on its SM runs (warp-specialised kernels), compared with 16 warps in the same code?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o icache_probe icache_probe.cu && ./icache_probe
// One CTA of 16 warps per SM (148 CTAs). Every warp runs a loop of ITER iterations whose body is BODY dependent
// "norm" evaluations on shared-memory data (3 LDS, 5 flops, IEEE sqrt, fmax) - the shape of the contact-history terms.
//   mode 0: all warps call the SAME function instance          (one instruction stream per SM)
//   mode 1: every warp calls ITS OWN instance (template on id)  (16 streams, same instruction count)
// Prints cycles per body-element per warp (clock64, averaged over warps and SMs), cold (first launch) and warm.
#include <cstdio>
#include <cuda_runtime.h>

template <int ID, int BODY>
__device__ __noinline__ float chain(const float* h, int iters, float thr) {
  float s = (float)ID * 1e-30f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int b = 0; b < BODY; ++b) {
      const float* f = h + ((it + b) & 63) * 3;
      const float ss = (f[0] * f[0] + f[1] * f[1]) + f[2] * f[2];
      const float n = (ss == 0.f) ? 0.f : sqrtf(ss);
      s += (n > thr) ? 1.f : 0.f;
    }
  }
  return s;
}

template <int BODY>
__global__ void __launch_bounds__(512) probe(int mode, int iters, float* out, long long* cyc) {
  __shared__ float sm[32 * 193];
  for (int i = threadIdx.x; i < 32 * 193; i += blockDim.x) sm[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* h = sm + lane * 193;
  const long long t0 = clock64();
  float r = 0.f;
  if (mode == 0) {
    r = chain<0, BODY>(h, iters, 0.5f);
  } else {
    // which code copy a warp runs: 1 = its own; 2 = shared by the 4 warps of its scheduler (warp % 4: 4 copies per
    // SM, 1 per scheduler); 3 = shared by 4 warps on different schedulers (warp / 4: 4 copies per SM, 4 per
    // scheduler); 4 = 2 copies per scheduler (8 per SM); 5 = 2 copies per SM
    int id = warp;
    if (mode == 2) id = warp % 4;
    if (mode == 3) id = warp / 4;
    if (mode == 4) id = warp % 8;
    if (mode == 5) id = warp % 2;
    switch (id) {
#define C(I) case I: r = chain<I + 1, BODY>(h, iters, 0.5f); break;
      C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

template <int BODY>
void run(int iters, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  long long h[148 * 16];
  for (int mode = 0; mode < 6; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(cyc, 0, 148 * 16 * 8);
      probe<BODY><<<148, warps * 32>>>(mode, iters, out, cyc);
      cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double tot = 0; int n = 0;
      for (int i = 0; i < 148 * 16; ++i) if (h[i] > 0) { tot += (double)h[i]; ++n; }
      printf("body %2d  iters %4d  warps %2d  mode %d (%s)  launch %d: %8.1f cycles per warp, %6.1f per norm\n", BODY, iters, warps, mode,
             mode == 0 ? "same code" : mode == 1 ? "own code per warp" : mode == 2 ? "4 copies, 1 per scheduler" : mode == 3 ? "4 copies, 4 per scheduler" : mode == 4 ? "8 copies, 2 per scheduler" : "2 copies", rep, tot / n, tot / n / ((double)iters * BODY));
    }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<3>(7, 16);     // one "7 bodies x 3 history samples" term per warp: the real task size
  run<3>(200, 16);   // long loop: steady state
  run<3>(200, 8);
  run<3>(200, 4);
  run<24>(1, 16);    // straight-line, executed once
  return 0;
}

// Follow-up to icache_probe.cu: is the ~3.9 B/cycle/SM instruction-fetch rate a property of ONE SM, or of something the
// SMs share (the L2 slices holding the code, a per-GPC instruction cache)? Sixteen warps per CTA, every warp runs its
// own copy of a straight-line block of 24 norms once (the regime of the step kernels), for
//   * grid sizes 1 .. 148 (how many SMs fetch at the same time), and
//   * "rotated" code assignment: warp w of CTA b runs copy (w + b) % 16 - every SM still executes all 16 copies, but
//     at any moment different SMs ask L2 for DIFFERENT lines instead of all for the same one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o icache_grid_probe icache_grid_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ID, int BODY>
__device__ __noinline__ float chain(const float* h, float thr) {
  float s = (float)ID * 1e-30f;
#pragma unroll
  for (int b = 0; b < BODY; ++b) {
    const float* f = h + (b & 63) * 3;
    const float ss = (f[0] * f[0] + f[1] * f[1]) + f[2] * f[2];
    const float n = (ss == 0.f) ? 0.f : sqrtf(ss);
    s += (n > thr) ? 1.f : 0.f;
  }
  return s;
}

template <int BODY>
__global__ void __launch_bounds__(512) probe(int copies, int rotate, float* out, long long* cyc) {
  __shared__ float sm[32 * 193];
  for (int i = threadIdx.x; i < 32 * 193; i += blockDim.x) sm[i] = (float)((i * 2654435761u) >> 20) * 1e-3f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* h = sm + lane * 193;
  const long long t0 = clock64();
  float r = 0.f;
  const int id = ((warp % copies) + (rotate ? blockIdx.x : 0)) % 16;
  switch (id) {
#define C(I) case I: r = chain<I + 1, BODY>(h, 0.5f); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (lane == 0) cyc[blockIdx.x * 16 + warp] = t1 - t0;
}

int main() {
  constexpr int BODY = 24;
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 16 * 8);
  static long long h[148 * 16];
  const int grids[] = {1, 2, 4, 8, 16, 37, 74, 111, 148};
  for (int copies : {16, 4, 1})
    for (int rotate = 0; rotate < 2; ++rotate)
      for (int grid : grids) {
        double warm_mean = 0, warm_max = 0;
        for (int rep = 0; rep < 3; ++rep) {   // rep 0 = cold, report the last (caches as warm as they get)
          cudaMemset(cyc, 0, 148 * 16 * 8);
          probe<BODY><<<grid, 512>>>(copies, rotate, out, cyc);
          cudaDeviceSynchronize();
          cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
          double tot = 0, mx = 0; int n = 0;
          for (int i = 0; i < grid * 16; ++i) if (h[i] > 0) { tot += (double)h[i]; mx = h[i] > mx ? (double)h[i] : mx; ++n; }
          warm_mean = tot / n; warm_max = mx;
        }
        printf("copies %2d  rotate %d  grid %3d: %8.0f cycles per warp (max %8.0f), %6.1f per norm\n", copies, rotate, grid,
               warm_mean, warm_max, warm_mean / BODY);
      }
  cudaFree(out); cudaFree(cyc);
  return 0;
}

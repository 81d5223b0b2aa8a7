// What a kernel launch inside a CUDA graph costs on this GPU, as a function of what the launch carries:
// grid size, parameter bytes, dynamic shared memory, the cluster attribute. Back-to-back dependent kernel nodes
// (stream capture), CUDA events around 20 replays of a 200-node graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_probe launch_probe.cu && ./launch_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

template <int BYTES> struct Params { char b[BYTES]; };

template <int BYTES> __global__ void k_params(const __grid_constant__ Params<BYTES> p, int* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && p.b[BYTES - 1] == 77) out[0] = 1;   // touches the last line only
}
template <int BYTES> __global__ void k_params_all(const __grid_constant__ Params<BYTES> p, int* out) {
  int s = 0;
  for (int i = threadIdx.x * 4; i < BYTES; i += blockDim.x * 4) s += *reinterpret_cast<const int*>(p.b + i);
  if (s == 0x7fffffff) out[0] = s;
}
__global__ void k_work(int* out, int iters) {   // a dependent chain: the CTA lives ~iters * 4 cycles
  float x = threadIdx.x;
  for (int i = 0; i < iters; ++i) x = x * 1.0001f + 0.5f;
  if (x == 12345.f) out[0] = 1;
}

template <class F> float time_graph(F launch, cudaStream_t s, int nodes = 200, int reps = 20) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
  for (int i = 0; i < nodes; ++i) launch();
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / (reps * nodes);
}

template <int BYTES> void probe_params(cudaStream_t s, int* out, int grid, int block) {
  Params<BYTES> p; memset(&p, 0, sizeof(p));
  float a = time_graph([&] { k_params<BYTES><<<grid, block, 0, s>>>(p, out); }, s);
  float b = time_graph([&] { k_params_all<BYTES><<<grid, block, 0, s>>>(p, out); }, s);
  printf("params %5d B, grid %4d x %4d: %.2f us per launch (kernel reads one line) / %.2f us (reads all of them)\n", BYTES, grid, block, a, b);
}

int main() {
  cudaStream_t s; cudaStreamCreate(&s);
  int* out; cudaMalloc(&out, 64);
  for (int grid : {1, 128, 148, 296}) {
    probe_params<64>(s, out, grid, 512);
    probe_params<1024>(s, out, grid, 512);
    probe_params<4096>(s, out, grid, 512);
    probe_params<4864>(s, out, grid, 512);
  }
  // dynamic shared memory and the cluster attribute
  for (int smem : {0, 48 * 1024, 100 * 1024}) {
    cudaFuncSetAttribute(k_work, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cl : {1, 4}) {
      cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      for (int iters : {0, 500, 2000}) {
        float t = time_graph([&] { cudaLaunchKernelEx(&cfg, k_work, out, iters); }, s);
        printf("k_work grid 128 x 512, smem %6d B, cluster %d, chain %4d iterations: %.2f us per launch\n", smem, cl, iters, t);
      }
    }
  }
  // programmatic dependent launch between identical nodes
  {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int iters : {0, 500, 2000}) {
      float t = time_graph([&] { cudaLaunchKernelEx(&cfg, k_work, out, iters); }, s);
      printf("k_work grid 128 x 512 with the PDL attribute (no griddepcontrol in the kernel), chain %4d: %.2f us per launch\n", iters, t);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmark (dev aid): time per node of a CUDA graph that is a chain of N dependent short kernels, with plain stream-order
// edges and with programmatic-dependent-launch edges (griddepcontrol at the top of the kernel).  Decides whether the engine's step
// (about 40 dependent kernels per iteration) can gain from PDL.      nvcc -arch=sm_100a -O3 -o graph_gap graph_gap.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_plain(float* p, int work) {
  float v = p[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  p[threadIdx.x] = v;
}
__global__ void k_pdl(float* p, int work) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = p[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  p[threadIdx.x] = v;
}
static float run(bool pdl, int nodes, int grid, int work, float* buf, cudaStream_t s) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nodes; ++i) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = 256; cfg.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    if (pdl) cudaLaunchKernelEx(&cfg, k_pdl, buf, work); else cudaLaunchKernelEx(&cfg, k_plain, buf, work);
  }
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 5; ++w) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  const int reps = 50;
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / (reps * nodes);
}
int main() {
  float* buf; cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20);
  cudaStream_t s; cudaStreamCreate(&s);
  const int grids[3] = {1, 148, 148 * 8};
  const int works[3] = {0, 2000, 20000};
  for (int gi = 0; gi < 3; ++gi)
    for (int wi = 0; wi < 3; ++wi) {
      float a = run(false, 40, grids[gi], works[wi], buf, s), b = run(true, 40, grids[gi], works[wi], buf, s);
      printf("grid %5d work %6d : plain %.3f us/node   pdl %.3f us/node   cudaerr %d\n", grids[gi], works[wi], a, b, (int)cudaGetLastError());
    }
  return 0;
}

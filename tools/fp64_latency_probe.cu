// Scratch probe: dependent-issue latency and single-warp issue interval of the FP64 instructions K-b is made of
// (DFMA, DMUL->DFMA rotation step, DMMA.884 accumulator chain), measured with clock64 in one warp of one CTA,
// and the same with 1..8 independent chains.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency_probe fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dfma(double *out, long long *cyc, int iters, double a, double b)
{
  double c[CH];
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x * 1e-9 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; ++i) s += c[i];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int CH>
__global__ void k_dmma(double *out, long long *cyc, int iters, double a, double b)
{
  double c[CH][2];
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = 0.0;
  a += threadIdx.x * 1e-9; b -= threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// rotation chain as in kbm_rot: 2 dependent levels per step
__global__ void k_rot(double *out, long long *cyc, int iters, double rc, double rs)
{
  double c = 1.0 + threadIdx.x * 1e-9, s = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double cn = fma(c, rc, -(s * rs)), sn = fma(s, rc, c * rs);
    c = cn; s = sn;
  }
  long long t1 = clock64();
  out[threadIdx.x] = c + s;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8 * 1024 * 64); cudaMalloc(&cyc, 8);
  const int iters = 4096;
#define RUN(name, kern, ch, blocks, threads)                                                            \
  kern<<<blocks, threads>>>(out, cyc, iters, 0.999999, 1e-7); cudaDeviceSynchronize();                   \
  kern<<<blocks, threads>>>(out, cyc, iters, 0.999999, 1e-7); cudaDeviceSynchronize();                   \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                                        \
  printf("{\"probe\": \"%s\", \"chains\": %d, \"warps_per_cta\": %d, \"cycles_per_instr\": %.2f, \"cycles_per_step\": %.2f}\n", name, ch, threads / 32, (double)h / iters / ch, (double)h / iters);
  RUN("dfma dependent", k_dfma<1>, 1, 1, 32)
  RUN("dfma", k_dfma<2>, 2, 1, 32)
  RUN("dfma", k_dfma<4>, 4, 1, 32)
  RUN("dfma", k_dfma<8>, 8, 1, 32)
  RUN("dfma", k_dfma<16>, 16, 1, 32)
  RUN("dmma884 dependent", k_dmma<1>, 1, 1, 32)
  RUN("dmma884", k_dmma<2>, 2, 1, 32)
  RUN("dmma884", k_dmma<4>, 4, 1, 32)
  RUN("dmma884", k_dmma<8>, 8, 1, 32)
  RUN("dmma884", k_dmma<16>, 16, 1, 32)
  RUN("dmma884 4 warps (1/SMSP)", k_dmma<4>, 4, 1, 128)
  RUN("dmma884 8 warps (2/SMSP)", k_dmma<4>, 4, 1, 256)
  RUN("dmma884 8 warps (2/SMSP)", k_dmma<8>, 8, 1, 256)
  RUN("dmma884 16 warps (4/SMSP)", k_dmma<4>, 4, 1, 512)
  k_rot<<<1, 32>>>(out, cyc, iters, 0.999999, 1e-3); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("{\"probe\": \"rotation step (DMUL -> DFMA), dependent\", \"cycles_per_step\": %.2f}\n", (double)h / iters);
  printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

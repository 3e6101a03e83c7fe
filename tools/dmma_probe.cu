// Scratch probe: FP64 tensor-core (DMMA, mma.sync m16n8k16 / m8n8k4) throughput on this GPU against the DFMA peak.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(256) dmma_k16(double *out, int iters, double seed)
{
  double a[8], b[4], c[CH][4];
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
  for (int ch = 0; ch < CH; ++ch) for (int i = 0; i < 4; ++i) c[ch][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[ch][0]), "+d"(c[ch][1]), "+d"(c[ch][2]), "+d"(c[ch][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int ch = 0; ch < CH; ++ch) for (int i = 0; i < 4; ++i) s += c[ch][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dmma_k4(double *out, int iters, double seed)
{
  double a = seed + 1e-9 * threadIdx.x, b = seed - 1e-9 * threadIdx.x, c[CH][2];
  for (int ch = 0; ch < CH; ++ch) c[ch][0] = c[ch][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[ch][0]), "+d"(c[ch][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int ch = 0; ch < CH; ++ch) s += c[ch][0] + c[ch][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dfma_k(double *out, int iters, double seed)
{
  double a = seed + 1e-9 * threadIdx.x, b = 1.0 - 1e-9, c[CH];
  for (int ch = 0; ch < CH; ++ch) c[ch] = ch;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) c[ch] = fma(c[ch], b, a);
  }
  double s = 0;
  for (int ch = 0; ch < CH; ++ch) s += c[ch];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA issued by the same warps: do the two share one datapath?
template <int CH>
__global__ void __launch_bounds__(256) both_k(double *out, int iters, double seed)
{
  double a[8], b[4], c[CH][4], f[2 * CH];
  for (int i = 0; i < 8; ++i) a[i] = seed + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 4; ++i) b[i] = seed - 1e-9 * (threadIdx.x + i);
  for (int ch = 0; ch < CH; ++ch) for (int i = 0; i < 4; ++i) c[ch][i] = 0.0;
  for (int i = 0; i < 2 * CH; ++i) f[i] = i;
  const double fb = 1.0 - 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[ch][0]), "+d"(c[ch][1]), "+d"(c[ch][2]), "+d"(c[ch][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      f[2 * ch] = fma(f[2 * ch], fb, a[0]);
      f[2 * ch + 1] = fma(f[2 * ch + 1], fb, a[1]);
    }
  }
  double s = 0;
  for (int ch = 0; ch < CH; ++ch) for (int i = 0; i < 4; ++i) s += c[ch][i];
  for (int i = 0; i < 2 * CH; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 4096;
  double *out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  const double warps = (double)blocks * threads / 32;
  constexpr int CH = 8;
  float t;
  t = time_ms([&] { dmma_k16<CH><<<blocks, threads>>>(out, iters, 1.0); });
  printf("{\"probe\": \"dmma m16n8k16\", \"ms\": %.4f, \"tflops\": %.2f}\n", t, warps * iters * CH * 4096.0 / (t * 1e-3) / 1e12);
  t = time_ms([&] { dmma_k4<CH><<<blocks, threads>>>(out, iters, 1.0); });
  printf("{\"probe\": \"dmma m8n8k4\", \"ms\": %.4f, \"tflops\": %.2f}\n", t, warps * iters * CH * 512.0 / (t * 1e-3) / 1e12);
  t = time_ms([&] { dfma_k<CH><<<blocks, threads>>>(out, iters, 1.0); });
  printf("{\"probe\": \"dfma\", \"ms\": %.4f, \"tflops\": %.2f}\n", t, warps * iters * CH * 64.0 / (t * 1e-3) / 1e12);
  t = time_ms([&] { both_k<CH><<<blocks, threads>>>(out, iters, 1.0); });
  printf("{\"probe\": \"dmma m16n8k16 + 2 dfma per mma, same warps\", \"ms\": %.4f, \"tflops_mma\": %.2f, \"tflops_fma\": %.2f}\n", t,
         warps * iters * CH * 4096.0 / (t * 1e-3) / 1e12, warps * iters * CH * 128.0 / (t * 1e-3) / 1e12);
  printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

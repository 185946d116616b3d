// Developer micro-benchmark: dependent-issue latencies that bound the per-warp critical path of
// the solver kernels (one warp, clock64 around N dependent operations).
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed + lane; sm[32 + lane] = seed;
  __syncwarp();
  double x = seed, y = seed * 0.5;
  long long t0, t1;
  // dependent DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, seed);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // dependent DADD
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x + y;
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // dependent max (DSETP + 2 FSEL)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double t = x * 0.999 ; x = t > y ? t : y; }
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;   // DMUL + max
  // dependent shuffle (64-bit = 2 SHFL)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_xor_sync(0xffffffffu, x, 16);
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // dependent LDS.64 (pointer chase through values)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = ((int)v) & 31; }
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  x += idx;
  // STS -> syncwarp -> LDS round trip (broadcast read of another lane's value)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // dependent rcp (MUFU.RCP64H + Newton, 5 FMAs)
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); e = fma(e, e, e); r = fma(e, r, r); e = fma(-x, r, 1.0); x = fma(e, r, r) + 1.5;
  }
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // 4 independent DFMA chains (issue rate)
  double a = x, b = y, c = x + 1, d = y + 1;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { a = fma(a, y, seed); b = fma(b, y, seed); c = fma(c, y, seed); d = fma(d, y, seed); }
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // dependent DMMA m8n8k4 chain
  double d0 = x, d1 = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(seed), "d"(y));
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  // LDS.128 dependent
  idx = lane & 15;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { double2 v = *reinterpret_cast<double2*>(&sm[2 * idx]); idx = ((int)(v.x + v.y)) & 15; }
  t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
  out[lane] = x + a + b + c + d + d0 + d1 + idx;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  k<<<1, 32>>>(out, cyc, 1.0000001); cudaDeviceSynchronize();
  k<<<1, 32>>>(out, cyc, 1.0000001); cudaDeviceSynchronize();
  long long h[16]; cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"dependent DFMA", "dependent DADD", "dependent DMUL+max(DSETP+2FSEL)", "dependent 64-bit SHFL", "dependent LDS.64 (+cvt)", "STS->syncwarp->LDS->DADD->syncwarp", "dependent rcp (MUFU+5 DFMA)+DADD", "4 independent DFMA chains (per 4 DFMA)", "dependent DMMA m8n8k4", "dependent LDS.128 (+add,cvt)"};
  for (int i = 0; i < 10; ++i) printf("%-42s %7.1f cycles\n", names[i], (double)h[i] / N);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}

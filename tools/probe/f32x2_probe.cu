// Throughput probe: scalar FADD/FMUL/FFMA vs packed add/mul/fma.f32x2 on sm_100a.  (tools only - not product code)
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float2* out, float2 seed) {
  float2 a[CH], b = seed, c = make_float2(seed.y, seed.x);
  for (int i = 0; i < CH; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) { a[i].x += b.x; a[i].y += b.y; }                                   // 2 FADD
      if (MODE == 1) a[i] = __fadd2_rn(a[i], b);                                          // 1 FADD2
      if (MODE == 2) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }  // 2 FFMA
      if (MODE == 3) a[i] = __ffma2_rn(a[i], b, c);                                       // 1 FFMA2
      if (MODE == 4) { a[i].x *= b.x; a[i].y *= b.y; }                                   // 2 FMUL
      if (MODE == 5) a[i] = __fmul2_rn(a[i], b);                                          // 1 FMUL2
      if (MODE == 6) { const float2 t = a[i]; a[i].x = fmaf(t.x, b.x, -t.y * b.y); a[i].y = fmaf(t.x, b.y, t.y * b.x); }  // scalar complex mul: 2 FMUL + 2 FFMA
      if (MODE == 7) { const float2 t = a[i]; a[i] = __ffma2_rn(make_float2(t.y, t.x), make_float2(-b.y, b.y), __fmul2_rn(t, make_float2(b.x, b.x))); }  // packed complex mul
      if (MODE == 8) { const float2 t = a[i]; a[i] = __fadd2_rn(t, make_float2(-a[(i + 1) % CH].x, -a[(i + 1) % CH].y)); }   // packed subtract
      if (MODE == 9) { const float2 t = a[i]; a[i] = __fadd2_rn(t, make_float2(-a[(i + 1) % CH].y, a[(i + 1) % CH].x)); }   // add of i*z (swap + negate one half)
    }
  }
  float2 s = make_float2(0.f, 0.f);
  for (int i = 0; i < CH; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_elem_scalar_ops) {
  float2* out; cudaMalloc(&out, 148 * 512 * sizeof(float2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148, 512>>>(out, make_float2(1.0001f, 0.9999f));
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<148, 512>>>(out, make_float2(1.0001f, 0.9999f));
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // complex-element updates per second per SM-cycle at 1.965 GHz nominal
  const double upd = 148.0 * 512 * ITERS * CH;
  printf("%-34s %8.3f ms  %7.2f Gupd/s  %6.2f lane-updates/clk/SM (1.965GHz)  scalar-op-equiv %5.1f/clk/SM\n", name, ms, upd / ms * 1e-6,
         upd / (ms * 1e-3) / 148 / 1.965e9, per_elem_scalar_ops * upd / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("2xFADD", 2); run<1>("FADD2", 2); run<2>("2xFFMA", 2); run<3>("FFMA2", 2); run<4>("2xFMUL", 2); run<5>("FMUL2", 2);
  run<6>("cmul scalar (2FMUL+2FFMA)", 4); run<7>("cmul packed (FMUL2+FFMA2+swap)", 4); run<8>("packed sub (neg)", 2); run<9>("packed add i*z (swap+neg)", 2);
  return 0;
}

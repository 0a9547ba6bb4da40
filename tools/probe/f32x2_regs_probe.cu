// Probe 2: does the packed / scalar FP32 rate depend on how many REGISTER operands an instruction reads?  (tools only)
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float2* out, const float2* in) {
  float2 a[CH], b[CH], c[CH];
  for (int i = 0; i < CH; ++i) { a[i] = in[threadIdx.x + 512 * i]; b[i] = in[threadIdx.x + 512 * (i + CH)]; c[i] = in[threadIdx.x + 512 * (i + 2 * CH)]; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) a[i] = __fadd2_rn(a[i], b[i]);                              // FADD2  R, R
      if (MODE == 1) a[i] = __fmul2_rn(a[i], b[i]);                              // FMUL2  R, R
      if (MODE == 2) a[i] = __ffma2_rn(a[i], b[i], c[i]);                        // FFMA2  R, R, R
      if (MODE == 3) a[i] = __ffma2_rn(a[i], b[i], a[i]);                        // FFMA2  R, R, R (2 distinct)
      if (MODE == 4) { a[i].x = fmaf(a[i].x, b[i].x, c[i].x); a[i].y = fmaf(a[i].y, b[i].y, c[i].y); }   // 2 x FFMA R, R, R
      if (MODE == 5) { a[i].x = a[i].x + b[i].x; a[i].y = a[i].y + b[i].y; }     // 2 x FADD R, R
      if (MODE == 6) a[i] = __ffma2_rn(make_float2(a[i].y, a[i].x), make_float2(-b[i].y, b[i].y), __fmul2_rn(a[i], make_float2(b[i].x, b[i].x)));   // register-twiddle complex multiply
      if (MODE == 7) a[i] = __ffma2_rn(a[i], make_float2(0.9238795f, 0.9238795f), c[i]);   // FFMA2 R, imm, R
    }
  }
  float2 s = make_float2(0.f, 0.f);
  for (int i = 0; i < CH; ++i) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double packed_per_update) {
  float2 *out, *in; cudaMalloc(&out, 148 * 512 * sizeof(float2)); cudaMalloc(&in, 512 * 3 * CH * sizeof(float2));
  cudaMemset(in, 0, 512 * 3 * CH * sizeof(float2));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<148, 512>>>(out, in);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<148, 512>>>(out, in);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double upd = 148.0 * 512 * ITERS * CH;       // float2 updates
  printf("%-40s %8.3f ms  %7.2f float2-updates/clk/SM @1.965GHz   (x%.0f instr each)\n", name, ms, upd / (ms * 1e-3) / 148 / 1.965e9, packed_per_update);
  cudaFree(out); cudaFree(in);
}
int main() {
  run<0>("FADD2 R,R", 1); run<1>("FMUL2 R,R", 1); run<2>("FFMA2 R,R,R (3 distinct)", 1); run<3>("FFMA2 R,R,R (2 distinct)", 1);
  run<4>("2 x FFMA R,R,R", 2); run<5>("2 x FADD R,R", 2); run<6>("cmul, register twiddle (FMUL2+FFMA2)", 2); run<7>("FFMA2 R,imm,R", 1);
  return 0;
}

// Micro-benchmark: FP32 FMA issue rate on sm_100a for scalar FFMA vs packed FFMA2 and for
// different register-operand patterns.  Development aid (results recorded in DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// (a) scalar, both multiplicands fixed: 1 RF read per FMA (accumulator)
__global__ void k_scalar_fixed(float* out, float a, float b) {
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = fmaf(a, b, acc[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (b) scalar, one multiplicand varies per FMA (8x8 outer product like SGEMM): a[i] reused, b[j] + c read
__global__ void k_scalar_outer(float* out, const float* in) {
  float acc[64], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i]; b[i] = in[threadIdx.x + 8 + i]; }
  for (int it = 0; it < ITER / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (c) packed: 8x(4 pairs) outer product: acc2[i][jp] += {a[i],a[i]} * {b[2jp], b[2jp+1]}
__global__ void k_packed_outer(float* out, const float* in) {
  float2 acc[32], a2[8], b2[4];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) { float v = in[threadIdx.x + i]; a2[i] = make_float2(v, v); }
#pragma unroll
  for (int i = 0; i < 4; ++i) b2[i] = make_float2(in[threadIdx.x + 8 + 2 * i], in[threadIdx.x + 9 + 2 * i]);
  for (int it = 0; it < ITER / 2; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i * 4 + j] = ffma2(a2[i], b2[j], acc[i * 4 + j]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (d) packed, fixed multiplicands
__global__ void k_packed_fixed(float* out, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = ffma2(a2, b2, acc[i]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double fma_per_thread, int threads, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  double tf = 2.0 * fma_per_thread * threads * blocks / (ms * 1e-3) / 1e12;
  printf("%-18s %8.3f ms  %7.2f TFLOP/s\n", name, ms, tf);
}

int main() {
  float *out, *in;
  const int blocks = 148 * 8, threads = 256;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&in, 4096 * 4);
  cudaMemset(in, 0, 4096 * 4);
  run("scalar_fixed", [&] { k_scalar_fixed<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 32.0 * ITER, threads, blocks);
  run("scalar_outer8x8", [&] { k_scalar_outer<<<blocks, threads>>>(out, in); }, 64.0 * ITER / 2, threads, blocks);
  run("packed_outer8x4", [&] { k_packed_outer<<<blocks, threads>>>(out, in); }, 64.0 * ITER / 2, threads, blocks);
  run("packed_fixed", [&] { k_packed_fixed<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 32.0 * ITER, threads, blocks);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

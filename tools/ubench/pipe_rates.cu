// Micro-benchmark: sustained warp-instruction rate per SM sub-partition of the integer / DPX / packed-float
// instructions the ORB kernels are built from (B200, sm_100a).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// Prints instructions per clock per SM for each op (128 lanes/clk/SM = 4 warp-inst/clk/SM is the issue limit).
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
  if constexpr (OP == 0) return __vimax3_u16x2(a, b, c);
  else if constexpr (OP == 1) return __vmaxu2(a, b) + 0 * c;
  else if constexpr (OP == 2) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
  else if constexpr (OP == 3) return __vabsdiffu4(a, b);
  else if constexpr (OP == 4) return __byte_perm(a, b, c);
  else if constexpr (OP == 5) return __funnelshift_r(a, b, 8);
  else if constexpr (OP == 6) return a * b + c;  // IMAD
  else if constexpr (OP == 7) return a + b;      // IADD3
  else if constexpr (OP == 8) return __vimax3_s32(a, b, c);
  else if constexpr (OP == 9) return max(a, b);
  else if constexpr (OP == 10) return __popc(a) + b;
  else return a;
}

template <int OP>
__global__ void k_int(unsigned* out, unsigned seed, long long* cyc) {
  unsigned v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = seed + threadIdx.x * 17 + i * 3;
  unsigned b = seed ^ 0x1234567u, c = seed * 3 + 1;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = op<OP>(v[i], b, c);
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_ffma2(float2* out, float seed, long long* cyc) {
  unsigned long long v[CHAINS];
  float2 t = make_float2(seed + threadIdx.x, seed * 0.5f);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = *reinterpret_cast<unsigned long long*>(&t) + i;
  unsigned long long b = v[0] ^ 0x11, c = v[1] ^ 0x33;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(b), "l"(c));
  }
  long long t1 = clock64();
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = *reinterpret_cast<float2*>(&s);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void k_ffma(float* out, float seed, long long* cyc) {
  float v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = seed + threadIdx.x + i;
  float b = seed * 0.999f, c = seed * 0.001f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = __fmaf_rn(v[i], b, c);
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, unsigned* d_out, long long* d_cyc, int threads) {
  k_int<OP><<<1, threads>>>(d_out, 12345u, d_cyc);
  k_int<OP><<<1, threads>>>(d_out, 12345u, d_cyc);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  double winst = (double)ITER * CHAINS * (threads / 32);
  printf("%-22s threads=%4d  %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, threads, winst / cyc,
         4.0 * cyc / winst);
}

int main() {
  unsigned* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 4096 * sizeof(float2));
  cudaMalloc(&d_cyc, sizeof(long long));
  for (int threads : {128, 512, 1024}) {
    run<0>("VIMNMX3.U16x2", d_out, d_cyc, threads);
    run<1>("VIMNMX.U16x2", d_out, d_cyc, threads);
    run<8>("VIMNMX3.S32", d_out, d_cyc, threads);
    run<9>("VIMNMX.U32", d_out, d_cyc, threads);
    run<2>("LOP3", d_out, d_cyc, threads);
    run<3>("VABSDIFF4", d_out, d_cyc, threads);
    run<4>("PRMT", d_out, d_cyc, threads);
    run<5>("SHF", d_out, d_cyc, threads);
    run<6>("IMAD", d_out, d_cyc, threads);
    run<7>("IADD3", d_out, d_cyc, threads);
    run<10>("POPC+IADD", d_out, d_cyc, threads);
    long long cyc;
    k_ffma2<<<1, threads>>>((float2*)d_out, 1.0f, d_cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("%-22s threads=%4d  %.3f warp-inst/clk/SM\n", "FFMA2", threads, (double)ITER * CHAINS * (threads / 32) / cyc);
    k_ffma<<<1, threads>>>((float*)d_out, 1.0f, d_cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    printf("%-22s threads=%4d  %.3f warp-inst/clk/SM\n", "FFMA", threads, (double)ITER * CHAINS * (threads / 32) / cyc);
  }
  return 0;
}

// Micro-benchmark: do two instruction kinds share an SM sub-partition pipe?  Two independent register chains per
// thread, one per kind; if a pair costs the sum of the single rates the kinds share a pipe, if it costs the maximum they
// issue to different pipes (B200, sm_100a).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_mix pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
#define CH 6

enum { LOP3 = 0, ADD, IMAD, IMADHI, SHF, VABS, PRMT, VMNMX3, IMADI, NONE };

template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c) {
  unsigned d = a;
  if constexpr (OP == LOP3) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if constexpr (OP == ADD) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  else if constexpr (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if constexpr (OP == IMADI) asm volatile("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(a), "r"(c));
  else if constexpr (OP == IMADHI) asm volatile("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if constexpr (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %1, %2, 8;" : "=r"(d) : "r"(a), "r"(b));
  else if constexpr (OP == VABS) d = __vabsdiffu4(a, b);
  else if constexpr (OP == PRMT) asm volatile("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(d) : "r"(a), "r"(b));
  else if constexpr (OP == VMNMX3) d = __vimax3_u16x2(a, b, c);
  return d;
}

template <int A, int B>
__global__ void k_mix(unsigned* out, unsigned seed, long long* cyc) {
  unsigned v[CH], w[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    v[i] = seed + threadIdx.x * 17 + i * 3;
    w[i] = seed * 7 + threadIdx.x * 5 + i;
  }
  unsigned b = seed ^ 0x1234567u, c = seed * 3 + 1;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      v[i] = op<A>(v[i], b, c);
      if constexpr (B != NONE) w[i] = op<B>(w[i], c, b);
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s ^= v[i] ^ w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int A, int B>
void run(const char* name, unsigned* d_out, long long* d_cyc) {
  const int threads = 1024;
  k_mix<A, B><<<1, threads>>>(d_out, 12345u, d_cyc);
  k_mix<A, B><<<1, threads>>>(d_out, 12345u, d_cyc);
  cudaDeviceSynchronize();
  long long cyc = 0;
  cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double steps = (double)ITER * CH * (threads / 32);  // (A [+ B]) issued per warp
  printf("%-24s %.2f clk per %s per SMSP\n", name, 4.0 * cyc / steps, B == NONE ? "instruction" : "pair");
}

int main() {
  unsigned* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, 4096 * sizeof(unsigned));
  cudaMalloc(&d_cyc, sizeof(long long));
  run<LOP3, NONE>("LOP3", d_out, d_cyc);
  run<ADD, NONE>("ADD", d_out, d_cyc);
  run<IMAD, NONE>("IMAD", d_out, d_cyc);
  run<IMADI, NONE>("IMAD imm", d_out, d_cyc);
  run<IMADHI, NONE>("IMAD.HI", d_out, d_cyc);
  run<SHF, NONE>("SHF", d_out, d_cyc);
  run<LOP3, ADD>("LOP3 + ADD", d_out, d_cyc);
  run<LOP3, IMAD>("LOP3 + IMAD", d_out, d_cyc);
  run<LOP3, IMADI>("LOP3 + IMAD imm", d_out, d_cyc);
  run<LOP3, IMADHI>("LOP3 + IMAD.HI", d_out, d_cyc);
  run<LOP3, SHF>("LOP3 + SHF", d_out, d_cyc);
  run<VABS, IMAD>("VABSDIFF4 + IMAD", d_out, d_cyc);
  run<VMNMX3, IMAD>("VIMNMX3 + IMAD", d_out, d_cyc);
  run<VMNMX3, LOP3>("VIMNMX3 + LOP3", d_out, d_cyc);
  run<PRMT, IMAD>("PRMT + IMAD", d_out, d_cyc);
  run<ADD, IMAD>("ADD + IMAD", d_out, d_cyc);
  return 0;
}

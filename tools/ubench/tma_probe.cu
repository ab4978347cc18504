// Probe: which forms of a u8 cp.async.bulk.tensor.3d tile load work on sm_100a?  (x alignment, box widths, big params)
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/tma_probe tools/ubench/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct Big { char pad[4096]; CUtensorMap m[2]; };
struct Small { CUtensorMap m[2]; };
template <typename PS>
__global__ void k(const __grid_constant__ PS P, int which, int x, int y, int z, int bw, int bh, unsigned char* out) {
  __shared__ __align__(128) unsigned char buf[64 * 40];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned sb = (unsigned)__cvta_generic_to_shared(buf), mb = (unsigned)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(bw * bh)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sb),
                 "l"(&P.m[which]), "r"(x), "r"(y), "r"(z), "r"(mb) : "memory");
  }
  __syncwarp();
  unsigned done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mb) : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += 32) out[i] = buf[i];
}
int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 752, H = 480, N = 4;
  std::vector<unsigned char> img((size_t)W * H * N);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *dout; cudaMalloc(&d, img.size()); cudaMalloc(&dout, 64 * 40);
  cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  auto mk = [&](CUtensorMap* m, unsigned bw, unsigned bh) {
    cuuint64_t dim[3] = {W, H, N}; cuuint64_t st[2] = {W, (cuuint64_t)W * H}; cuuint32_t box[3] = {bw, bh, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dim, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
  };
  static Small S; static Big B;
  printf("encode 48x37: %d, 32x31: %d\n", mk(&S.m[0], 48, 37), mk(&S.m[1], 32, 31));
  memcpy(B.m, S.m, sizeof(S.m));
  auto check = [&](const char* name, int which, int x, int y, int z, int bw, int bh, bool big) {
    cudaMemset(dout, 0xEE, 64 * 40);
    if (big) k<Big><<<1, 32>>>(B, which, x, y, z, bw, bh, dout); else k<Small><<<1, 32>>>(S, which, x, y, z, bw, bh, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s x=%d: CUDA error %s\n", name, x, cudaGetErrorString(e)); return false; }
    std::vector<unsigned char> o(bw * bh); cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c) {
      const int gx = x + c, gy = y + r;
      const unsigned char want = (gx < W && gy < H && gx >= 0 && gy >= 0) ? img[((size_t)z * H + gy) * W + gx] : 0;
      bad += o[r * bw + c] != want;
    }
    printf("%-28s x=%d y=%d z=%d: %s (%d bad)\n", name, x, y, z, bad ? "MISMATCH" : "ok", bad);
    return true;
  };
  if (!check("small params, aligned x", 0, 64, 100, 1, 48, 37, false)) return 1;
  if (!check("small params, unaligned x", 0, 77, 101, 2, 48, 37, false)) return 1;
  if (!check("small params, 32x31 unaligned", 1, 333, 17, 3, 32, 31, false)) return 1;
  if (!check("small params, right edge OOB", 0, 730, 440, 0, 48, 37, false)) return 1;
  if (!check("big params (>4KB), unaligned", 0, 77, 101, 2, 48, 37, true)) return 1;
  return 0;
}

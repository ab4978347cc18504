// Concurrent host<->device copy bandwidth on 1..N GPUs of one box (one process per GPU, started together):
// is the end-to-end frames/s of N ranks capped by each GPU's PCIe link or by the host memory system they share?
//   nvcc -O2 -o tools/pcie_scale tools/pcie_scale.cu ; tools/pcie_scale [max_gpus]
// Modes: H2D only, D2H only, both directions at once (the bench's steady state: 1.44 GB up + 0.39 GB down per step),
// from plain pinned memory (cudaHostAlloc default) and from write-combined pinned memory.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

struct Shared {
  pthread_barrier_t bar;
  double gbs[16][3];  // per rank: h2d, d2h, and (bidir) h2d+d2h bytes over the common interval
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void child(int rank, Shared* sh, bool wc, size_t up, size_t down, int iters) {
  cudaSetDevice(rank);
  void *h_up, *h_down, *d_up, *d_down;
  cudaHostAlloc(&h_up, up, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  cudaHostAlloc(&h_down, down, cudaHostAllocDefault);
  memset(h_up, rank + 1, up);
  memset(h_down, 0, down);
  cudaMalloc(&d_up, up);
  cudaMalloc(&d_down, down);
  cudaStream_t s1, s2;
  cudaStreamCreate(&s1);
  cudaStreamCreate(&s2);
  cudaMemcpyAsync(d_up, h_up, up, cudaMemcpyHostToDevice, s1);
  cudaMemcpyAsync(h_down, d_down, down, cudaMemcpyDeviceToHost, s2);
  cudaDeviceSynchronize();
  for (int mode = 0; mode < 3; ++mode) {
    pthread_barrier_wait(&sh->bar);
    const double t0 = now();
    for (int i = 0; i < iters; ++i) {
      if (mode != 1) cudaMemcpyAsync(d_up, h_up, up, cudaMemcpyHostToDevice, s1);
      if (mode != 0) cudaMemcpyAsync(h_down, d_down, down, cudaMemcpyDeviceToHost, s2);
    }
    cudaDeviceSynchronize();
    const double dt = now() - t0;
    const double bytes = (mode != 1 ? (double)up : 0.0) + (mode != 0 ? (double)down : 0.0);
    sh->gbs[rank][mode] = bytes * iters / dt / 1e9;
    pthread_barrier_wait(&sh->bar);
  }
  _exit(0);
}

int main(int argc, char** argv) {
  int maxg = argc > 1 ? atoi(argv[1]) : 8;
  const size_t up = 2ull * 752 * 480 * 1000, down = 193ull * 1000 * 1000;  // one bench step per 1000 frames
  Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  printf("{\"bytes_up\": %zu, \"bytes_down\": %zu, \"runs\": [\n", up, down);
  bool first = true;
  for (int wc = 0; wc < 2; ++wc)
    for (int n = 1; n <= maxg; n *= 2) {
      pthread_barrierattr_t at;
      pthread_barrierattr_init(&at);
      pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
      pthread_barrier_init(&sh->bar, &at, n);
      memset(sh->gbs, 0, sizeof(sh->gbs));
      for (int r = 0; r < n; ++r)
        if (fork() == 0) child(r, sh, wc != 0, up, down, 6);
      for (int r = 0; r < n; ++r) wait(nullptr);
      double sum[3] = {0, 0, 0}, mn[3] = {1e9, 1e9, 1e9};
      for (int r = 0; r < n; ++r)
        for (int m = 0; m < 3; ++m) {
          sum[m] += sh->gbs[r][m];
          if (sh->gbs[r][m] < mn[m]) mn[m] = sh->gbs[r][m];
        }
      printf("%s {\"gpus\": %d, \"host_memory\": \"%s\", \"h2d_sum_gbs\": %.1f, \"h2d_min_gbs\": %.1f, \"d2h_sum_gbs\": %.1f, "
             "\"d2h_min_gbs\": %.1f, \"bidir_sum_gbs\": %.1f, \"bidir_min_gbs\": %.1f}",
             first ? "" : ",\n", n, wc ? "pinned write-combined" : "pinned", sum[0], mn[0], sum[1], mn[1], sum[2], mn[2]);
      first = false;
      fflush(stdout);
      pthread_barrier_destroy(&sh->bar);
    }
  printf("\n]}\n");
  return 0;
}

// Host-side check of the exact retainBest replay (orb-slam3-rust_b200/csrc/select_serial.h + the lane-parallel
// formulation of the Hoare pass used by warp_partition_pivot in orb_kernels.cuh) against the real libstdc++
// std::nth_element / std::partition / std::__heap_select.  Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "select_serial.h"

template <typename T>
static T make_rec(float resp, int id);
template <>
vrec_t make_rec<vrec_t>(float resp, int id) {
  unsigned u;
  memcpy(&u, &resp, 4);
  return ((vrec_t)u << 32) | (unsigned)id;
}
template <>
uint32_t make_rec<uint32_t>(float resp, int id) {
  return ((uint32_t)resp << 24) | ((uint32_t)id & 0xffffffu);
}

// The formulation the device code uses: stopper lists -> K -> pair swaps -> cut = min(L[K], R[K-1]).
template <typename T>
static int par_partition_pivot(std::vector<T>& v, int first, int last) {
  vsel::median_to_first(v.data(), first, first + 1, first + (last - first) / 2, last - 1);
  const T p = v[first];
  std::vector<int> L, R;
  for (int j = last - 1; j >= first; --j)
    if (!vsel::comp(p, v[j])) R.push_back(j);
  for (int i = first + 1; i < last; ++i)
    if (!vsel::comp(v[i], p)) L.push_back(i);
  int K = 0;
  while (K < (int)L.size() && K < (int)R.size() && L[K] < R[K]) K++;
  int cut = last;
  if (K < (int)L.size()) cut = std::min(cut, L[K]);
  if (K > 0) cut = std::min(cut, R[K - 1]);
  for (int k = 0; k < K; ++k) std::swap(v[L[k]], v[R[k]]);
  return cut;
}

template <typename T>
static int par_retain_best(std::vector<T>& v, int n_points, int depth_override, bool* used_heap) {
  int n = (int)v.size();
  if (n_points < 0 || n <= n_points) return n;
  if (n_points == 0) return 0;
  const int nth = n_points - 1;
  int first = 0, last = n;
  int depth = depth_override >= 0 ? depth_override : 2 * vsel::floor_log2(n);
  bool tail = true;
  while (last - first > 3) {
    if (depth == 0) {
      vsel::heap_select_fallback(v.data(), first, nth, last);
      *used_heap = true;
      tail = false;
      break;
    }
    --depth;
    int cut = par_partition_pivot(v, first, last);
    if (cut <= nth) first = cut;
    else last = cut;
  }
  if (tail) vsel::insertion_sort(v.data(), first, last);
  const T thr = v[nth];
  std::vector<int> A, B;
  for (int j = n - 1; j >= n_points; --j)
    if (!vsel::comp(thr, v[j])) B.push_back(j);
  for (int i = n_points; i < n; ++i)
    if (vsel::comp(thr, v[i])) A.push_back(i);
  for (size_t k = 0; k < std::min(A.size(), B.size()); ++k)
    if (A[k] < B[k]) v[A[k]] = v[B[k]];
  return n_points + (int)B.size();
}

template <typename T>
static int ref_retain_best(std::vector<T>& v, int n_points, int depth_override) {
  int n = (int)v.size();
  if (n_points < 0 || n <= n_points) return n;
  if (n_points == 0) return 0;
  auto greater = [](const T& a, const T& b) { return vsel::comp(a, b); };
  if (depth_override >= 0)
    std::__introselect(v.begin(), v.begin() + n_points - 1, v.end(), (long)depth_override,
                       __gnu_cxx::__ops::__iter_comp_iter(greater));
  else std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), greater);
  const T thr = v[n_points - 1];
  auto e = std::partition(v.begin() + n_points, v.end(), [thr](const T& a) { return !vsel::comp(thr, a); });
  return (int)(e - v.begin());
}

template <typename T>
static int run(const char* name, int iters, unsigned seed) {
  std::mt19937 rng(seed);
  long heap_hits = 0;
  for (int it = 0; it < iters; ++it) {
    int N = 4 + rng() % 6000;
    if (it % 3 == 0) N = 4 + rng() % 70;
    int vals = it % 4 == 0 ? 3 : it % 4 == 1 ? 40 : it % 4 == 2 ? 235 : 250;
    std::vector<T> v(N);
    for (int i = 0; i < N; ++i) {
      float r = (float)(rng() % vals);
      if (sizeof(T) == 8 && it % 5 == 0) r = (float)((int)(rng() % 2000000) - 1000000) * 1e-9f;  // Harris-like, +/-
      v[i] = make_rec<T>(r, i);
    }
    if (it % 7 == 0) std::sort(v.begin(), v.end(), [](const T& a, const T& b) { return vsel::comp(a, b); });
    if (it % 11 == 0) std::sort(v.begin(), v.end(), [](const T& a, const T& b) { return vsel::comp(b, a); });
    int n = 1 + rng() % N;
    if (n >= N) n = N - 1;
    if (n < 1) continue;
    // every 4th test forces a tiny depth limit so the heap-select fallback is exercised
    int depth = it % 4 == 3 ? (int)(rng() % 3) : -1;
    std::vector<T> a = v, b = v;
    bool used_heap = false;
    int na = ref_retain_best(a, n, depth);
    int nb = par_retain_best(b, n, depth, &used_heap);
    heap_hits += used_heap;
    bool ok = na == nb;
    for (int i = 0; ok && i < na; ++i) ok = a[i] == b[i];
    if (!ok) {
      printf("%s MISMATCH it=%d N=%d n=%d depth=%d na=%d nb=%d\n", name, it, N, n, depth, na, nb);
      return 1;
    }
  }
  printf("%s ok iters=%d heap_fallbacks=%ld\n", name, iters, heap_hits);
  return heap_hits > 0 ? 0 : 2;
}

int main() {
  int rc = run<vrec_t>("vrec_t(float response)", 6000, 1);
  rc |= run<uint32_t>("uint32(FAST score)", 6000, 2);
  return rc;
}

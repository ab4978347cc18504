// Serial pieces of the exact replay of cv::KeyPointsFilter::retainBest as built against libstdc++
// (std::nth_element = __introselect).  The O(n) partition passes run warp-parallel in
// orb_kernels.cuh; what is left here is O(1)/rare and runs on one lane: median-of-three pivot move, the final
// <=3-element insertion sort, and the heap-select fallback taken when the depth limit is exhausted.
//
// Compiles for host too (VSLAM_HD) so tests/test_select_replay.py can check it against the real
// std::nth_element on the CPU — that host build is a test, never a product path.
#pragma once
#include "vslam_params.h"

#if defined(__CUDACC__)
#define VSLAM_HD __host__ __device__ __forceinline__
#else
#define VSLAM_HD inline
#endif

namespace vsel {

// Two record types are permuted by the replay:
//   vrec_t   (64 bit)  hi = float bits of the Harris response, lo = (y<<16)|x        -> retainBest(N) on Harris
//   uint32_t (32 bit)  FAST score (S-1, 8 bit) << 24 | y << 12 | x  (w, h <= 4096)   -> retainBest(2N) on FAST score
// comp(a,b) := a.response > b.response  (cv::KeypointResponseGreater)
VSLAM_HD float resp_of(vrec_t r) {
  unsigned u = (unsigned)(r >> 32);
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
#endif
}
VSLAM_HD unsigned resp_of(uint32_t r) { return r >> 24; }
template <typename T>
VSLAM_HD bool comp(T a, T b) {
  return resp_of(a) > resp_of(b);
}

template <typename T>
VSLAM_HD void swap_rec(T* v, int i, int j) {
  T t = v[i];
  v[i] = v[j];
  v[j] = t;
}

// std::__move_median_to_first(result=first, a=first+1, b=mid, c=last-1)
template <typename T>
VSLAM_HD void median_to_first(T* v, int first, int a, int b, int c) {
  if (comp(v[a], v[b])) {
    if (comp(v[b], v[c])) swap_rec(v, first, b);
    else if (comp(v[a], v[c])) swap_rec(v, first, c);
    else swap_rec(v, first, a);
  } else if (comp(v[a], v[c])) swap_rec(v, first, a);
  else if (comp(v[b], v[c])) swap_rec(v, first, c);
  else swap_rec(v, first, b);
}

// std::__insertion_sort on [first,last)
template <typename T>
VSLAM_HD void insertion_sort(T* v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    T val = v[i];
    if (comp(val, v[first])) {
      for (int k = i; k > first; --k) v[k] = v[k - 1];
      v[first] = val;
    } else {
      int hole = i, next = i - 1;
      while (comp(val, v[next])) {
        v[hole] = v[next];
        hole = next;
        --next;
      }
      v[hole] = val;
    }
  }
}

// std::__push_heap
template <typename T>
VSLAM_HD void push_heap(T* v, int hole, int top, T val) {
  int parent = (hole - 1) / 2;
  while (hole > top && comp(v[parent], val)) {
    v[hole] = v[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[hole] = val;
}

// std::__adjust_heap on the heap v[0..len)
template <typename T>
VSLAM_HD void adjust_heap(T* v, int hole, int len, T val) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp(v[child], v[child - 1])) child--;
    v[hole] = v[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[hole] = v[child - 1];
    hole = child - 1;
  }
  push_heap(v, hole, top, val);
}

// std::__heap_select(first, middle, last) followed by iter_swap(first, nth) — the introselect fallback.
template <typename T>
VSLAM_HD void heap_select_fallback(T* v, int first, int nth, int last) {
  T* h = v + first;
  const int middle = nth + 1 - first;  // heap length
  const int n = last - first;
  if (middle >= 2) {  // std::__make_heap
    int parent = (middle - 2) / 2;
    while (true) {
      T val = h[parent];
      adjust_heap(h, parent, middle, val);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < n; ++i) {
    if (comp(h[i], h[0])) {  // std::__pop_heap(first, middle, i)
      T val = h[i];
      h[i] = h[0];
      adjust_heap(h, 0, middle, val);
    }
  }
  swap_rec(v, first, nth);
}

VSLAM_HD int floor_log2(int n) {
  int k = 0;
  while (n > 1) {
    n >>= 1;
    ++k;
  }
  return k;
}

}  // namespace vsel

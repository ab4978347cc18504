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
// comp(a,b) := a.response > b.response  (KeypointResponseGreater)
VSLAM_HD bool comp(vrec_t a, vrec_t b) { return resp_of(a) > resp_of(b); }

VSLAM_HD void swap_rec(vrec_t* v, int i, int j) {
  vrec_t t = v[i];
  v[i] = v[j];
  v[j] = t;
}

// std::__move_median_to_first(result=first, a=first+1, b=mid, c=last-1)
VSLAM_HD void median_to_first(vrec_t* v, int first, int a, int b, int c) {
  if (comp(v[a], v[b])) {
    if (comp(v[b], v[c])) swap_rec(v, first, b);
    else if (comp(v[a], v[c])) swap_rec(v, first, c);
    else swap_rec(v, first, a);
  } else if (comp(v[a], v[c])) swap_rec(v, first, a);
  else if (comp(v[b], v[c])) swap_rec(v, first, c);
  else swap_rec(v, first, b);
}

// std::__insertion_sort on [first,last)
VSLAM_HD void insertion_sort(vrec_t* v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    vrec_t val = v[i];
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
VSLAM_HD void push_heap(vrec_t* v, int hole, int top, vrec_t val) {
  int parent = (hole - 1) / 2;
  while (hole > top && comp(v[parent], val)) {
    v[hole] = v[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[hole] = val;
}

// std::__adjust_heap on the heap v[0..len)
VSLAM_HD void adjust_heap(vrec_t* v, int hole, int len, vrec_t val) {
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
VSLAM_HD void heap_select_fallback(vrec_t* v, int first, int nth, int last) {
  vrec_t* h = v + first;
  const int middle = nth + 1 - first;  // heap length
  const int n = last - first;
  if (middle >= 2) {  // std::__make_heap
    int parent = (middle - 2) / 2;
    while (true) {
      vrec_t val = h[parent];
      adjust_heap(h, parent, middle, val);
      if (parent == 0) break;
      parent--;
    }
  }
  for (int i = middle; i < n; ++i) {
    if (comp(h[i], h[0])) {  // std::__pop_heap(first, middle, i)
      vrec_t val = h[i];
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

// Stand-alone entry points of the C ABI (the pieces of process() and the tracking-side searches).
// Included at the end of vslam_b200.cu.  Inputs come from host memory, are staged into a grow-only device
// arena owned by the handle, processed by the same kernels the batch pipeline uses, and copied back.
namespace {

// Reserve `total` bytes (grow-only); sub-allocations are carved with arena_take.
int arena_begin(vslam_handle h, size_t total) {
  total += 4096;
  if (h->arena_cap < total) {
    if (h->arena_base) cudaFree(h->arena_base);
    h->arena_base = nullptr;
    h->arena_cap = 0;
    size_t want = std::max(total, (size_t)1 << 20);
    if (cudaMalloc(&h->arena_base, want) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, VSLAM_ERR_OUT_OF_MEMORY, "device scratch allocation of %zu bytes failed", want);
    }
    h->arena_cap = want;
  }
  h->arena_used = 0;
  return VSLAM_OK;
}
template <typename T>
T* arena_take(vslam_handle h, size_t count) {
  size_t off = align_up(h->arena_used, 256);
  h->arena_used = off + count * sizeof(T);
  return reinterpret_cast<T*>(h->arena_base + off);
}
inline size_t pad256(size_t b) { return align_up(b, 256) + 256; }

}  // namespace

extern "C" {

int vslam_descriptor_distance(vslam_handle h, const uint8_t* a, const uint8_t* b, int n, uint32_t* out) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || (n > 0 && (!a || !b || !out))) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (n == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  const size_t nb = (size_t)n * 32;
  int rc = arena_begin(h, 2 * pad256(nb) + pad256((size_t)n * 4));
  if (rc) return rc;
  unsigned* da = arena_take<unsigned>(h, (size_t)n * 8);
  unsigned* db = arena_take<unsigned>(h, (size_t)n * 8);
  unsigned* dout = arena_take<unsigned>(h, (size_t)n);
  cudaStream_t st = h->slots[0].stream;
  CU_TRY(h, cudaMemcpyAsync(da, a, nb, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(db, b, nb, cudaMemcpyHostToDevice, st));
  vk::k_hamming<<<(n + 255) / 256, 256, 0, st>>>(da, db, n, dout);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

// Build a one-frame result set (same SoA layout as the pipeline's) around caller-provided features.
static int stage_features(vslam_handle h, const vslam_keypoint* kl, const uint8_t* dl, int nl, const vslam_keypoint* kr,
                          const uint8_t* dr, int nr, FrameLayout* F, int2** ml, cudaStream_t st) {
  const int cap = std::max(std::max(nl, nr), 1);
  FrameLayout L;
  const size_t bytes = carve_results(nullptr, 1, cap, &L);
  int rc = arena_begin(h, pad256(bytes) + pad256((size_t)cap * sizeof(int2)));
  if (rc) return rc;
  unsigned char* base = arena_take<unsigned char>(h, bytes);
  carve_results(base, 1, cap, &L);
  *ml = arena_take<int2>(h, (size_t)cap);
  ResHdr hdr;
  memset(&hdr, 0, sizeof(hdr));
  hdr.n_kp[0] = nl;
  hdr.n_kp[1] = nr;
  CU_TRY(h, cudaMemcpyAsync(L.hdr, &hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  if (nl > 0) CU_TRY(h, cudaMemcpyAsync(L.kp[0], kl, (size_t)nl * 28, cudaMemcpyHostToDevice, st));
  if (nr > 0) CU_TRY(h, cudaMemcpyAsync(L.kp[1], kr, (size_t)nr * 28, cudaMemcpyHostToDevice, st));
  if (nl > 0 && dl) CU_TRY(h, cudaMemcpyAsync(L.desc[0], dl, (size_t)nl * 32, cudaMemcpyHostToDevice, st));
  if (nr > 0 && dr) CU_TRY(h, cudaMemcpyAsync(L.desc[1], dr, (size_t)nr * 32, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaStreamSynchronize(st));  // hdr is a stack variable
  *F = L;
  return VSLAM_OK;
}

int vslam_match_features(vslam_handle h, const vslam_keypoint* kp_left, const uint8_t* desc_left, int n_left,
                         const vslam_keypoint* kp_right, const uint8_t* desc_right, int n_right, vslam_dmatch* matches,
                         int* n_matches) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n_left < 0 || n_right < 0 || !n_matches || (n_left > 0 && (!kp_left || !desc_left || !matches)) ||
      (n_right > 0 && (!kp_right || !desc_right)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (n_left > 65535 || n_right > 65535) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "more than 65535 keypoints");
  *n_matches = 0;
  if (n_left == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  FrameLayout F;
  int2* ml;
  int rc = stage_features(h, kp_left, desc_left, n_left, kp_right, desc_right, n_right, &F, &ml, st);
  if (rc) return rc;
  dim3 grid((F.cap + vk::kMatchLeftPerBlock - 1) / vk::kMatchLeftPerBlock, 1);
  vk::k_match<<<grid, 256, 0, st>>>(F, h->S, ml);
  vk::k_match_finalize<<<1, 256, 0, st>>>(F, h->S, ml);
  h->launches += 2;
  ResHdr hdr;
  CU_TRY(h, cudaMemcpyAsync(&hdr, F.hdr, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  *n_matches = hdr.n_matches;
  if (hdr.n_matches > 0)
    CU_TRY(h, cudaMemcpy(matches, F.matches, (size_t)hdr.n_matches * 16, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

// triangulate: the device kernel is shared with the pipeline (k_match_finalize consumes a per-left table).
int vslam_triangulate(vslam_handle h, const vslam_keypoint* kp_left, int n_left, const vslam_keypoint* kp_right,
                      int n_right, const vslam_dmatch* matches, int n_matches, double* points, uint8_t* valid) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n_left < 0 || n_right < 0 || n_matches < 0 || (n_left > 0 && (!kp_left || !points || !valid)) ||
      (n_right > 0 && !kp_right) || (n_matches > 0 && !matches))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (n_left == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  FrameLayout F;
  int2* ml;
  int rc = stage_features(h, kp_left, nullptr, n_left, kp_right, nullptr, n_right, &F, &ml, st);
  if (rc) return rc;
  // per-left table from the caller's match list.  match_features never emits a left index twice; if a caller does, the
  // reference's loop (stereo.rs:196-213) lets the last entry with |disparity| >= 0.5 win: an entry it skips never
  // overwrites an earlier point, so such entries are dropped here as well.
  std::vector<int2> tab((size_t)F.cap, make_int2(-1, 0));
  for (int i = 0; i < n_matches; ++i) {
    int q = matches[i].query_idx, t = matches[i].train_idx;
    if (q < 0 || q >= n_left || t < 0 || t >= n_right) continue;  // .get() returned None
    const double d = (double)kp_left[q].x - (double)kp_right[t].x;
    if (std::fabs(d) < 0.5) continue;
    tab[q] = make_int2(t, (int)matches[i].distance);
  }
  CU_TRY(h, cudaMemcpyAsync(ml, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  vk::k_match_finalize<<<1, 256, 0, st>>>(F, h->S, ml);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(points, F.pts, (size_t)n_left * 24, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(valid, F.valid, (size_t)n_left, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

static vk::GridP grid_params(double img_w, double img_h) {
  vk::GridP G;
  G.inv_w = (double)VSLAM_GRID_COLS / (img_w - 0.0);
  G.inv_h = (double)VSLAM_GRID_ROWS / (img_h - 0.0);
  return G;
}

int vslam_grid_build(vslam_handle h, const vslam_keypoint* kps, int n, double img_w, double img_h, int32_t* cell_start,
                     int32_t* cell_items) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || !cell_start || (n > 0 && (!kps || !cell_items))) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nn = (size_t)std::max(n, 1);
  int rc = arena_begin(h, pad256(nn * 28) + pad256((vk::kGridCells + 1) * 4) + 2 * pad256(nn * 4));
  if (rc) return rc;
  unsigned char* dk = arena_take<unsigned char>(h, nn * 28);
  int* dcs = arena_take<int>(h, vk::kGridCells + 1);
  int* dci = arena_take<int>(h, nn);
  int* dco = arena_take<int>(h, nn);
  if (n > 0) CU_TRY(h, cudaMemcpyAsync(dk, kps, (size_t)n * 28, cudaMemcpyHostToDevice, st));
  vk::k_grid_build<<<1, 1024, 0, st>>>(dk, n, grid_params(img_w, img_h), dcs, dci, dco);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(cell_start, dcs, (vk::kGridCells + 1) * 4, cudaMemcpyDeviceToHost, st));
  if (n > 0) CU_TRY(h, cudaMemcpyAsync(cell_items, dci, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

int vslam_grid_query(vslam_handle h, const vslam_keypoint* kps, int n, double img_w, double img_h, const double* xyr,
                     int nq, int32_t* out_start, int32_t* out_items, int out_capacity) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || nq < 0 || !out_start || (n > 0 && !kps) || (nq > 0 && !xyr) || out_capacity < 0 ||
      (out_capacity > 0 && !out_items))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nn = (size_t)std::max(n, 1), qq = (size_t)std::max(nq, 1), oc = (size_t)std::max(out_capacity, 1);
  int rc = arena_begin(h, pad256(nn * 28) + pad256((vk::kGridCells + 1) * 4) + 2 * pad256(nn * 4) + pad256(qq * 24) +
                              2 * pad256((qq + 1) * 4) + pad256(oc * 4));
  if (rc) return rc;
  unsigned char* dk = arena_take<unsigned char>(h, nn * 28);
  int* dcs = arena_take<int>(h, vk::kGridCells + 1);
  int* dci = arena_take<int>(h, nn);
  int* dco = arena_take<int>(h, nn);
  double* dq = arena_take<double>(h, qq * 3);
  int* dcnt = arena_take<int>(h, qq + 1);
  int* dstart = arena_take<int>(h, qq + 1);
  int* dout = arena_take<int>(h, oc);
  const vk::GridP G = grid_params(img_w, img_h);
  if (n > 0) CU_TRY(h, cudaMemcpyAsync(dk, kps, (size_t)n * 28, cudaMemcpyHostToDevice, st));
  if (nq > 0) CU_TRY(h, cudaMemcpyAsync(dq, xyr, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
  vk::k_grid_build<<<1, 1024, 0, st>>>(dk, n, G, dcs, dci, dco);
  h->launches++;
  if (nq > 0) {
    vk::k_grid_count<<<(nq + 255) / 256, 256, 0, st>>>(dcs, G, dq, nq, dcnt);
    h->launches++;
  }
  vk::k_scan_counts<<<1, 1024, 0, st>>>(dcnt, nq, dstart);
  h->launches++;
  if (nq > 0) {
    vk::k_grid_fill<<<(nq * 32 + 255) / 256, 256, 0, st>>>(dcs, dci, G, dq, nq, dstart, dout, out_capacity);
    h->launches++;
  }
  CU_TRY(h, cudaMemcpyAsync(out_start, dstart, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  const int total = out_start[nq];
  if (total > out_capacity) return fail(h, VSLAM_ERR_CAPACITY, "grid query needs %d items, capacity %d", total, out_capacity);
  if (total > 0) CU_TRY(h, cudaMemcpy(out_items, dout, (size_t)total * 4, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_search_by_projection(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, const double* uv,
                               const uint8_t* mp_desc, int M, const vslam_search_params* sp, int32_t* out_idx,
                               int32_t* out_dist) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || M < 0 || !sp || (n > 0 && (!kps || !desc)) || (M > 0 && (!uv || !mp_desc || !out_idx || !out_dist)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (sp->mode != VSLAM_SEARCH_LOCAL_MAP && sp->mode != VSLAM_SEARCH_MOTION_MODEL)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown search mode %d", sp->mode);
  if (sp->th_high > 1024) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "th_high out of range");
  if (M == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nn = (size_t)std::max(n, 1), mm = (size_t)M;
  int rc = arena_begin(h, pad256(nn * 28) + pad256(nn * 32) + pad256((vk::kGridCells + 1) * 4) + 2 * pad256(nn * 4) +
                              pad256(mm * 16) + pad256(mm * 32) + 2 * pad256(mm * 4));
  if (rc) return rc;
  unsigned char* dk = arena_take<unsigned char>(h, nn * 28);
  unsigned* dd = arena_take<unsigned>(h, nn * 8);
  int* dcs = arena_take<int>(h, vk::kGridCells + 1);
  int* dci = arena_take<int>(h, nn);
  int* dco = arena_take<int>(h, nn);
  double* duv = arena_take<double>(h, mm * 2);
  unsigned* dmd = arena_take<unsigned>(h, mm * 8);
  int* doi = arena_take<int>(h, mm);
  int* dod = arena_take<int>(h, mm);
  const vk::GridP G = grid_params(sp->img_w, sp->img_h);
  vk::SearchP S;
  S.mode = sp->mode;
  S.radius = sp->radius;
  S.th_high = sp->th_high;
  S.nn_ratio = sp->nn_ratio;
  S.bound_w = sp->bound_w;
  S.bound_h = sp->bound_h;
  if (n > 0) {
    CU_TRY(h, cudaMemcpyAsync(dk, kps, (size_t)n * 28, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  }
  CU_TRY(h, cudaMemcpyAsync(duv, uv, mm * 16, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(dmd, mp_desc, mm * 32, cudaMemcpyHostToDevice, st));
  vk::k_grid_build<<<1, 1024, 0, st>>>(dk, n, G, dcs, dci, dco);
  vk::k_search<<<(M * 32 + 255) / 256, 256, 0, st>>>(dcs, dci, dd, G, S, duv, dmd, M, doi, dod);
  h->launches += 2;
  CU_TRY(h, cudaMemcpyAsync(out_idx, doi, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out_dist, dod, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

int vslam_bf_match_crosscheck(vslam_handle h, const uint8_t* query, int n_query, const uint8_t* train, int n_train,
                              vslam_dmatch* matches, int* n_matches) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n_query < 0 || n_train < 0 || !n_matches || (n_query > 0 && (!query || !matches)) || (n_train > 0 && !train))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *n_matches = 0;
  if (n_query == 0 || n_train == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nq = (size_t)n_query, nt = (size_t)n_train;
  int rc = arena_begin(h, pad256(nq * 32) + pad256(nt * 32) + 2 * pad256(nq * 4) + 2 * pad256(nt * 4) + pad256(nq * 16) + 512);
  if (rc) return rc;
  unsigned* dq = arena_take<unsigned>(h, nq * 8);
  unsigned* dt = arena_take<unsigned>(h, nt * 8);
  int* sidx = arena_take<int>(h, nq);
  int* sdist = arena_take<int>(h, nq);
  int* tidx = arena_take<int>(h, nt);
  int* tdist = arena_take<int>(h, nt);
  int4* dm = arena_take<int4>(h, nq);
  int* dn = arena_take<int>(h, 1);
  CU_TRY(h, cudaMemcpyAsync(dq, query, nq * 32, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(dt, train, nt * 32, cudaMemcpyHostToDevice, st));
  vk::k_nearest<<<(n_query + 7) / 8, 256, 0, st>>>(dq, n_query, dt, n_train, sidx, sdist);
  vk::k_nearest<<<(n_train + 7) / 8, 256, 0, st>>>(dt, n_train, dq, n_query, tidx, tdist);
  vk::k_cross<<<1, 256, 0, st>>>(sidx, sdist, tidx, n_query, dm, dn);
  h->launches += 3;
  int nm = 0;
  CU_TRY(h, cudaMemcpyAsync(&nm, dn, 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  *n_matches = nm;
  if (nm > 0) CU_TRY(h, cudaMemcpy(matches, dm, (size_t)nm * 16, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_fuse_search(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, const double* uvr,
                      const uint8_t* mp_desc, int M, uint32_t desc_threshold, int32_t* out_idx, int32_t* out_dist) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || M < 0 || (n > 0 && (!kps || !desc)) || (M > 0 && (!uvr || !mp_desc || !out_idx || !out_dist)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (M == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nn = (size_t)std::max(n, 1), mm = (size_t)M;
  int rc = arena_begin(h, pad256(nn * 28) + pad256(nn * 32) + pad256(mm * 24) + pad256(mm * 32) + 2 * pad256(mm * 4));
  if (rc) return rc;
  unsigned char* dk = arena_take<unsigned char>(h, nn * 28);
  unsigned* dd = arena_take<unsigned>(h, nn * 8);
  double* dq = arena_take<double>(h, mm * 3);
  unsigned* dmd = arena_take<unsigned>(h, mm * 8);
  int* doi = arena_take<int>(h, mm);
  int* dod = arena_take<int>(h, mm);
  if (n > 0) {
    CU_TRY(h, cudaMemcpyAsync(dk, kps, (size_t)n * 28, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(dd, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  }
  CU_TRY(h, cudaMemcpyAsync(dq, uvr, mm * 24, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(dmd, mp_desc, mm * 32, cudaMemcpyHostToDevice, st));
  vk::k_fuse_search<<<(M * 32 + 255) / 256, 256, 0, st>>>(dk, dd, n, dq, dmd, M, desc_threshold, doi, dod);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(out_idx, doi, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out_dist, dod, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

int vslam_distinctive_descriptors(vslam_handle h, const uint8_t* desc, const int32_t* offsets, int M, int32_t* out_idx) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (M < 0 || (M > 0 && (!offsets || !out_idx))) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (M == 0) return VSLAM_OK;
  const int total = offsets[M];
  if (offsets[0] != 0 || total < 0 || (total > 0 && !desc)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad offsets");
  for (int m = 0; m < M; ++m)
    if (offsets[m + 1] < offsets[m]) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "offsets must be non-decreasing");
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t tt = (size_t)std::max(total, 1), mm = (size_t)M;
  int rc = arena_begin(h, pad256(tt * 32) + pad256((mm + 1) * 4) + pad256(mm * 4));
  if (rc) return rc;
  unsigned* dd = arena_take<unsigned>(h, tt * 8);
  int* dof = arena_take<int>(h, mm + 1);
  int* dout = arena_take<int>(h, mm);
  if (total > 0) CU_TRY(h, cudaMemcpyAsync(dd, desc, (size_t)total * 32, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(dof, offsets, (mm + 1) * 4, cudaMemcpyHostToDevice, st));
  vk::k_distinctive<<<(M * 32 + 255) / 256, 256, 0, st>>>(dd, dof, M, dout);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(out_idx, dout, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

int vslam_vocab_upload(vslam_handle h, int n_nodes, const uint8_t* node_desc, const int32_t* child_begin,
                       const int32_t* child_ids, const uint32_t* parent, const int32_t* word_id) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n_nodes < 1 || !node_desc || !child_begin || !parent || !word_id)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "vocabulary: null argument or no root node");
  const int nch = child_begin[n_nodes];
  if (child_begin[0] != 0 || nch < 0 || (nch > 0 && !child_ids)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad child_begin");
  for (int i = 0; i < n_nodes; ++i)
    if (child_begin[i + 1] < child_begin[i]) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "child_begin must be non-decreasing");
  for (int c = 0; c < nch; ++c)
    if (child_ids[c] < 0 || child_ids[c] >= n_nodes) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "child id %d out of range", child_ids[c]);
  for (int i = 0; i < n_nodes; ++i)
    if (parent[i] != 0xffffffffu && parent[i] >= (uint32_t)n_nodes)
      return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "parent of node %d out of range", i);
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  CU_TRY(h, cudaStreamSynchronize(st));
  cudaFree(h->voc_desc); cudaFree(h->voc_child_begin); cudaFree(h->voc_child_ids); cudaFree(h->voc_parent); cudaFree(h->voc_word);
  h->voc_desc = nullptr; h->voc_child_begin = h->voc_child_ids = h->voc_word = nullptr; h->voc_parent = nullptr;
  h->voc_nodes = 0;
  const size_t nn = (size_t)n_nodes;
  CU_TRY(h, cudaMalloc(&h->voc_desc, nn * 32));
  CU_TRY(h, cudaMalloc(&h->voc_child_begin, (nn + 1) * 4));
  CU_TRY(h, cudaMalloc(&h->voc_child_ids, (size_t)std::max(nch, 1) * 4));
  CU_TRY(h, cudaMalloc(&h->voc_parent, nn * 4));
  CU_TRY(h, cudaMalloc(&h->voc_word, nn * 4));
  CU_TRY(h, cudaMemcpy(h->voc_desc, node_desc, nn * 32, cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->voc_child_begin, child_begin, (nn + 1) * 4, cudaMemcpyHostToDevice));
  if (nch > 0) CU_TRY(h, cudaMemcpy(h->voc_child_ids, child_ids, (size_t)nch * 4, cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->voc_parent, parent, nn * 4, cudaMemcpyHostToDevice));
  CU_TRY(h, cudaMemcpy(h->voc_word, word_id, nn * 4, cudaMemcpyHostToDevice));
  h->voc_nodes = n_nodes;
  return VSLAM_OK;
}

int vslam_bow_quantize(vslam_handle h, const uint8_t* desc, int n, int levels_up, uint32_t* out_word,
                       uint32_t* out_leaf, uint32_t* out_node) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (h->voc_nodes <= 0) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "no vocabulary uploaded (vslam_vocab_upload)");
  if (n < 0 || levels_up < 0 || (n > 0 && (!desc || !out_word || !out_leaf || !out_node)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bow_quantize: bad argument");
  if (n == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nn = (size_t)n;
  int rc = arena_begin(h, pad256(nn * 32) + 3 * pad256(nn * 4));
  if (rc) return rc;
  unsigned* dd = arena_take<unsigned>(h, nn * 8);
  unsigned* dw = arena_take<unsigned>(h, nn);
  unsigned* dl = arena_take<unsigned>(h, nn);
  unsigned* dn = arena_take<unsigned>(h, nn);
  CU_TRY(h, cudaMemcpyAsync(dd, desc, nn * 32, cudaMemcpyHostToDevice, st));
  vk::k_bow_quantize<<<(unsigned)((nn * 8 + 255) / 256), 256, 0, st>>>(dd, n, h->voc_desc, h->voc_child_begin, h->voc_child_ids,
                                                                     h->voc_parent, h->voc_word, levels_up, dw, dl, dn);
  h->launches++;
  CU_TRY(h, cudaMemcpyAsync(out_word, dw, nn * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out_leaf, dl, nn * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out_node, dn, nn * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return VSLAM_OK;
}

// ---- search_for_triangulation / search_for_triangulation_bow (local_mapping/triangulation.rs:401-530, 541-658)
static int tri_search(vslam_handle h, const vslam_keypoint* kp1, const uint8_t* desc1, const uint8_t* flags1, int n1,
                      const vslam_keypoint* kp2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                      const vslam_tri_params* tp, int n_groups, const int32_t* g1_start, const int32_t* g1_items,
                      const int32_t* g2_start, const int32_t* g2_items, int32_t* out_pairs, int* n_pairs) {
  const bool bow = n_groups >= 0;
  if (n1 < 0 || n2 < 0 || !tp || !n_pairs || (n1 > 0 && (!kp1 || !desc1 || !flags1 || !out_pairs)) ||
      (n2 > 0 && (!kp2 || !desc2 || !flags2)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (n1 > 65535 || n2 > 65535) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "more than 65535 keypoints");
  *n_pairs = 0;
  vk::TriP T;
  memcpy(T.F, tp->F, sizeof(T.F));
  T.ex = tp->epipole[0];
  T.ey = tp->epipole[1];
  T.max_dist = tp->max_dist;
  T.radius = tp->radius;
  // grid dimensions: ((image_width as f32 / 32.0).ceil() as usize).min(64)   (triangulation.rs:349-350)
  T.cols = (int)std::min<uint64_t>((uint64_t)std::ceil((float)tp->image_width / 32.0f), 64);
  T.rows = (int)std::min<uint64_t>((uint64_t)std::ceil((float)tp->image_height / 32.0f), 64);
  int n_visits = n1;
  std::vector<int> visit_idx1, visit_group;
  if (bow) {
    if (n_groups > 0 && (!g1_start || !g2_start)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null group offsets");
    if (n_groups > 0 && (g1_start[0] != 0 || g2_start[0] != 0)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad group offsets");
    for (int g = 0; g < n_groups; ++g) {
      if (g1_start[g + 1] < g1_start[g] || g2_start[g + 1] < g2_start[g] || g2_start[g + 1] - g2_start[g] > 65535)
        return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "group offsets must be non-decreasing (<= 65535 candidates per group)");
      for (int a = g1_start[g]; a < g1_start[g + 1]; ++a) {
        visit_idx1.push_back(g1_items[a]);
        visit_group.push_back(g);
      }
    }
    n_visits = (int)visit_idx1.size();
    if ((n_visits > 0 && !g1_items) || (n_groups > 0 && g2_start[n_groups] > 0 && !g2_items))
      return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null group items");
  } else if (T.cols < 1 || T.rows < 1) {
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "image_width / image_height must be positive");
  }
  if (n_visits == 0 || n2 == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t a1 = (size_t)n1, a2 = (size_t)n2, nv = (size_t)n_visits;
  const size_t n_g2 = bow ? (size_t)std::max(g2_start[n_groups], 1) : 1;
  int cap = 64;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const size_t lists = nv * (size_t)cap * sizeof(uint2);
    if (lists > ((size_t)1 << 31)) return fail(h, VSLAM_ERR_OUT_OF_MEMORY, "candidate lists of %zu bytes", 2 * lists);
    int rc = arena_begin(h, pad256(a1 * 28) + pad256(a1 * 32) + pad256(a1) + pad256(a2 * 28) + pad256(a2 * 32) + pad256(a2) +
                                pad256((vk::kTriMaxCells + 1) * 4) + 2 * pad256(a2 * 4) + 2 * pad256(nv * 4) +
                                pad256(((size_t)std::max(n_groups, 0) + 1) * 4) + pad256(n_g2 * 4) + 2 * pad256(lists) +
                                pad256(nv * 4) + pad256(nv * 8) + 512);
    if (rc) return rc;
    unsigned char* dk1 = arena_take<unsigned char>(h, a1 * 28);
    unsigned* dd1 = arena_take<unsigned>(h, a1 * 8);
    unsigned char* df1 = arena_take<unsigned char>(h, a1);
    unsigned char* dk2 = arena_take<unsigned char>(h, a2 * 28);
    unsigned* dd2 = arena_take<unsigned>(h, a2 * 8);
    unsigned char* df2 = arena_take<unsigned char>(h, a2);
    int* dcs = arena_take<int>(h, vk::kTriMaxCells + 1);
    int* dci = arena_take<int>(h, a2);
    int* dco = arena_take<int>(h, a2);
    int* dvi = arena_take<int>(h, nv);
    int* dvg = arena_take<int>(h, nv);
    int* dgs = arena_take<int>(h, (size_t)std::max(n_groups, 0) + 1);
    int* dgi = arena_take<int>(h, n_g2);
    uint2* draw = arena_take<uint2>(h, nv * (size_t)cap);
    uint2* dlist = arena_take<uint2>(h, nv * (size_t)cap);
    int* dcnt = arena_take<int>(h, nv);
    int* dpairs = arena_take<int>(h, nv * 2);
    int* dnp = arena_take<int>(h, 2);
    CU_TRY(h, cudaMemcpyAsync(dk1, kp1, a1 * 28, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(dd1, desc1, a1 * 32, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(df1, flags1, a1, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(dk2, kp2, a2 * 28, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(dd2, desc2, a2 * 32, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(df2, flags2, a2, cudaMemcpyHostToDevice, st));
    if (bow) {
      CU_TRY(h, cudaMemcpyAsync(dvi, visit_idx1.data(), nv * 4, cudaMemcpyHostToDevice, st));
      CU_TRY(h, cudaMemcpyAsync(dvg, visit_group.data(), nv * 4, cudaMemcpyHostToDevice, st));
      CU_TRY(h, cudaMemcpyAsync(dgs, g2_start, ((size_t)n_groups + 1) * 4, cudaMemcpyHostToDevice, st));
      if (g2_start[n_groups] > 0)
        CU_TRY(h, cudaMemcpyAsync(dgi, g2_items, (size_t)g2_start[n_groups] * 4, cudaMemcpyHostToDevice, st));
    } else {
      vk::k_tri_grid_build<<<1, 1024, 0, st>>>(dk2, n2, T.cols, T.rows, dcs, dci, dco);
      h->launches++;
    }
    vk::k_tri_candidates<<<(unsigned)((nv * 32 + 255) / 256), 256, 0, st>>>(
        dk1, dd1, df1, n1, dk2, dd2, df2, n2, T, dcs, dci, bow ? dvi : nullptr, dvg, dgs, dgi, n_visits, cap, draw, dlist, dcnt);
    vk::k_tri_resolve<<<1, 32, (size_t)((n2 + 31) / 32) * 4, st>>>(dlist, dcnt, cap, bow ? dvi : nullptr, n_visits, n2, dpairs, dnp);
    h->launches += 2;
    int np[2] = {0, 0};
    CU_TRY(h, cudaMemcpyAsync(np, dnp, 8, cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    CU_TRY(h, cudaGetLastError());
    if (np[1] > cap) {  // a list was truncated (more than 64 candidates pass every gate): once more with room for all
      cap = np[1];
      continue;
    }
    *n_pairs = np[0];
    if (np[0] > 0) CU_TRY(h, cudaMemcpy(out_pairs, dpairs, (size_t)np[0] * 8, cudaMemcpyDeviceToHost));
    return VSLAM_OK;
  }
  return fail(h, VSLAM_ERR_CUDA, "search_for_triangulation: candidate lists did not converge");
}

int vslam_search_for_triangulation(vslam_handle h, const vslam_keypoint* kp1, const uint8_t* desc1, const uint8_t* flags1,
                                   int n1, const vslam_keypoint* kp2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                   const vslam_tri_params* params, int32_t* out_pairs, int* n_pairs) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  return tri_search(h, kp1, desc1, flags1, n1, kp2, desc2, flags2, n2, params, -1, nullptr, nullptr, nullptr, nullptr,
                    out_pairs, n_pairs);
}

int vslam_search_for_triangulation_bow(vslam_handle h, const vslam_keypoint* kp1, const uint8_t* desc1,
                                       const uint8_t* flags1, int n1, const vslam_keypoint* kp2, const uint8_t* desc2,
                                       const uint8_t* flags2, int n2, const vslam_tri_params* params, int n_groups,
                                       const int32_t* g1_start, const int32_t* g1_items, const int32_t* g2_start,
                                       const int32_t* g2_items, int32_t* out_pairs, int* n_pairs) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n_groups < 0) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "n_groups < 0");
  return tri_search(h, kp1, desc1, flags1, n1, kp2, desc2, flags2, n2, params, n_groups, g1_start, g1_items, g2_start,
                    g2_items, out_pairs, n_pairs);
}

// ---- frames resident in HBM for the tracking-side searches (SURVEY.md 8 b: search_by_projection(handle, frame_id, ..))
static int retained_slot(vslam_handle h, int need, int* id_out) {
  int id = -1;
  for (size_t i = 0; i < h->retained.size() && id < 0; ++i)
    if (!h->retained[i].used) id = (int)i;
  if (id < 0) {
    if (h->retained.size() >= 4096) return fail(h, VSLAM_ERR_OUT_OF_MEMORY, "more than 4096 retained frames");
    h->retained.emplace_back();
    id = (int)h->retained.size() - 1;
  }
  vslam_handle_s::Retained& R = h->retained[id];
  if (R.cap < need) {
    cudaFree(R.kp); cudaFree(R.desc); cudaFree(R.cell_start); cudaFree(R.cell_items); cudaFree(R.cell_of);
    R = vslam_handle_s::Retained();
    const size_t c = (size_t)std::max(need, 1);
    if (cudaMalloc(&R.kp, c * 28) != cudaSuccess || cudaMalloc(&R.desc, c * 32) != cudaSuccess ||
        cudaMalloc(&R.cell_start, (vk::kGridCells + 1) * 4) != cudaSuccess || cudaMalloc(&R.cell_items, c * 4) != cudaSuccess ||
        cudaMalloc(&R.cell_of, c * 4) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, VSLAM_ERR_OUT_OF_MEMORY, "retained frame allocation failed");
    }
    R.cap = (int)c;
  }
  *id_out = id;
  return VSLAM_OK;
}

static vslam_handle_s::Retained* retained_get(vslam_handle h, int id) {
  if (id < 0 || id >= (int)h->retained.size() || !h->retained[id].used) return nullptr;
  return &h->retained[id];
}

int vslam_frame_retain(vslam_handle h, int index, double img_w, double img_h, int* frame_id) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!frame_id || !(img_w > 0.0) || !(img_h > 0.0)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad argument");
  *frame_id = -1;
  Slot* src = nullptr;
  for (int k = 0; k < h->nslot && !src; ++k) {
    Slot& s = h->slots[k];
    if (!s.busy && s.frames > 0 && s.serial == h->call_serial && h->call_serial > 0 && index >= s.first_frame &&
        index < s.first_frame + s.frames)
      src = &s;
  }
  if (!src)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT,
                "frame %d of the last process call is no longer resident (only the last %d chunks of a batch are)", index,
                h->nslot);
  const int f = index - src->first_frame;
  const ResHdr& hd = src->h_lay.hdr[f];
  if (hd.status != VSLAM_OK) return fail(h, hd.status, "frame %d did not complete (status %d)", index, hd.status);
  const int n = hd.n_kp[0];
  DEVICE_SCOPE(h);
  int id;
  int rc = retained_slot(h, std::max(n, h->kp_cap), &id);
  if (rc) return rc;
  vslam_handle_s::Retained& R = h->retained[id];
  cudaStream_t st = h->slots[0].stream;
  const FrameLayout& D = src->d_lay;
  if (n > 0) {
    CU_TRY(h, cudaMemcpyAsync(R.kp, D.kp[0] + (size_t)f * D.kp_stride, (size_t)n * 28, cudaMemcpyDeviceToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(R.desc, D.desc[0] + (size_t)f * D.desc_stride, (size_t)n * 32, cudaMemcpyDeviceToDevice, st));
  }
  vk::k_grid_build<<<1, 1024, 0, st>>>(R.kp, n, grid_params(img_w, img_h), R.cell_start, R.cell_items, R.cell_of);
  h->launches++;
  CU_TRY(h, cudaGetLastError());
  // The copies run on slot 0's stream.  A frame of slot 0 (every vslam_process call) is protected by stream order; the
  // result arrays of another slot are rewritten by that slot's own stream on the next call, so wait here.
  if (src != &h->slots[0]) CU_TRY(h, cudaStreamSynchronize(st));
  R.n = n;
  R.img_w = img_w;
  R.img_h = img_h;
  R.used = true;
  *frame_id = id;
  return VSLAM_OK;
}

int vslam_frame_upload(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, double img_w, double img_h,
                       int* frame_id) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!frame_id || n < 0 || n > 65535 || (n > 0 && (!kps || !desc)) || !(img_w > 0.0) || !(img_h > 0.0))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad argument");
  *frame_id = -1;
  DEVICE_SCOPE(h);
  int id;
  int rc = retained_slot(h, std::max(n, h->kp_cap), &id);
  if (rc) return rc;
  vslam_handle_s::Retained& R = h->retained[id];
  cudaStream_t st = h->slots[0].stream;
  if (n > 0) {
    CU_TRY(h, cudaMemcpyAsync(R.kp, kps, (size_t)n * 28, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(R.desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  }
  vk::k_grid_build<<<1, 1024, 0, st>>>(R.kp, n, grid_params(img_w, img_h), R.cell_start, R.cell_items, R.cell_of);
  h->launches++;
  CU_TRY(h, cudaStreamSynchronize(st));  // the caller's arrays may go away
  R.n = n;
  R.img_w = img_w;
  R.img_h = img_h;
  R.used = true;
  *frame_id = id;
  return VSLAM_OK;
}

int vslam_frame_release(vslam_handle h, int frame_id) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  vslam_handle_s::Retained* R = retained_get(h, frame_id);
  if (!R) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown frame id %d", frame_id);
  R->used = false;  // the device arrays are kept for the next retain; stream order protects pending searches
  return VSLAM_OK;
}

int vslam_frame_num_features(vslam_handle h, int frame_id) {
  vslam_handle_s::Retained* R = h ? retained_get(h, frame_id) : nullptr;
  return R ? R->n : -1;
}

int vslam_frame_search_by_projection(vslam_handle h, int frame_id, const double* uv, const uint8_t* mp_desc, int M,
                                     const vslam_search_params* sp, int32_t* out_idx, int32_t* out_dist) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  vslam_handle_s::Retained* R = retained_get(h, frame_id);
  if (!R) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown frame id %d", frame_id);
  if (M < 0 || !sp || (M > 0 && (!uv || !mp_desc || !out_idx || !out_dist)))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (sp->mode != VSLAM_SEARCH_LOCAL_MAP && sp->mode != VSLAM_SEARCH_MOTION_MODEL)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown search mode %d", sp->mode);
  if (sp->th_high > 1024) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "th_high out of range");
  if (sp->img_w != R->img_w || sp->img_h != R->img_h)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "search grid extent %gx%g differs from the retained frame's %gx%g", sp->img_w,
                sp->img_h, R->img_w, R->img_h);
  if (M == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t mm = (size_t)M;
  int rc = arena_begin(h, pad256(mm * 16) + pad256(mm * 32) + 2 * pad256(mm * 4));
  if (rc) return rc;
  double* duv = arena_take<double>(h, mm * 2);
  unsigned* dmd = arena_take<unsigned>(h, mm * 8);
  int* doi = arena_take<int>(h, mm);
  int* dod = arena_take<int>(h, mm);
  vk::SearchP S;
  S.mode = sp->mode;
  S.radius = sp->radius;
  S.th_high = sp->th_high;
  S.nn_ratio = sp->nn_ratio;
  S.bound_w = sp->bound_w;
  S.bound_h = sp->bound_h;
  CU_TRY(h, cudaMemcpyAsync(duv, uv, mm * 16, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(dmd, mp_desc, mm * 32, cudaMemcpyHostToDevice, st));
  if (h->profiling) cudaEventRecord(h->ev[0], st);
  vk::k_search<<<(M * 32 + 255) / 256, 256, 0, st>>>(R->cell_start, R->cell_items, R->desc, grid_params(R->img_w, R->img_h), S, duv,
                                                      dmd, M, doi, dod);
  h->launches++;
  if (h->profiling) cudaEventRecord(h->ev[1], st);
  CU_TRY(h, cudaMemcpyAsync(out_idx, doi, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaMemcpyAsync(out_dist, dod, mm * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  if (h->profiling) cudaEventElapsedTime(&h->last_search_ms, h->ev[0], h->ev[1]);
  return VSLAM_OK;
}

int vslam_frame_grid_query(vslam_handle h, int frame_id, const double* xyr, int nq, int32_t* out_start, int32_t* out_items,
                           int out_capacity) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  vslam_handle_s::Retained* R = retained_get(h, frame_id);
  if (!R) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown frame id %d", frame_id);
  if (nq < 0 || !out_start || (nq > 0 && !xyr) || out_capacity < 0 || (out_capacity > 0 && !out_items))
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t qq = (size_t)std::max(nq, 1), oc = (size_t)std::max(out_capacity, 1);
  int rc = arena_begin(h, pad256(qq * 24) + 2 * pad256((qq + 1) * 4) + pad256(oc * 4));
  if (rc) return rc;
  double* dq = arena_take<double>(h, qq * 3);
  int* dcnt = arena_take<int>(h, qq + 1);
  int* dstart = arena_take<int>(h, qq + 1);
  int* dout = arena_take<int>(h, oc);
  const vk::GridP G = grid_params(R->img_w, R->img_h);
  if (nq > 0) {
    CU_TRY(h, cudaMemcpyAsync(dq, xyr, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
    vk::k_grid_count<<<(nq + 255) / 256, 256, 0, st>>>(R->cell_start, G, dq, nq, dcnt);
    h->launches++;
  }
  vk::k_scan_counts<<<1, 1024, 0, st>>>(dcnt, nq, dstart);
  h->launches++;
  if (nq > 0) {
    vk::k_grid_fill<<<(nq * 32 + 255) / 256, 256, 0, st>>>(R->cell_start, R->cell_items, G, dq, nq, dstart, dout, out_capacity);
    h->launches++;
  }
  CU_TRY(h, cudaMemcpyAsync(out_start, dstart, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  const int total = out_start[nq];
  if (total > out_capacity) return fail(h, VSLAM_ERR_CAPACITY, "grid query needs %d items, capacity %d", total, out_capacity);
  if (total > 0) CU_TRY(h, cudaMemcpy(out_items, dout, (size_t)total * 4, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_frame_bf_match_crosscheck(vslam_handle h, int query_frame_id, int train_frame_id, vslam_dmatch* matches,
                                    int* n_matches) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  vslam_handle_s::Retained* Q = retained_get(h, query_frame_id);
  vslam_handle_s::Retained* T = retained_get(h, train_frame_id);
  if (!Q || !T) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "unknown frame id");
  if (!n_matches || (Q->n > 0 && !matches)) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *n_matches = 0;
  if (Q->n == 0 || T->n == 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  cudaStream_t st = h->slots[0].stream;
  const size_t nq = (size_t)Q->n, nt = (size_t)T->n;
  int rc = arena_begin(h, 2 * pad256(nq * 4) + 2 * pad256(nt * 4) + pad256(nq * 16) + 512);
  if (rc) return rc;
  int* sidx = arena_take<int>(h, nq);
  int* sdist = arena_take<int>(h, nq);
  int* tidx = arena_take<int>(h, nt);
  int* tdist = arena_take<int>(h, nt);
  int4* dm = arena_take<int4>(h, nq);
  int* dn = arena_take<int>(h, 1);
  vk::k_nearest<<<(Q->n + 7) / 8, 256, 0, st>>>(Q->desc, Q->n, T->desc, T->n, sidx, sdist);
  vk::k_nearest<<<(T->n + 7) / 8, 256, 0, st>>>(T->desc, T->n, Q->desc, Q->n, tidx, tdist);
  vk::k_cross<<<1, 256, 0, st>>>(sidx, sdist, tidx, Q->n, dm, dn);
  h->launches += 3;
  int nm = 0;
  CU_TRY(h, cudaMemcpyAsync(&nm, dn, 4, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  *n_matches = nm;
  if (nm > 0) CU_TRY(h, cudaMemcpy(matches, dm, (size_t)nm * 16, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

float vslam_last_search_ms(vslam_handle h) { return h ? h->last_search_ms : 0.f; }

// ---- classic mode accessors (NOT reference parity; include/vslam_b200.h)
static Slot* slot_of_last_call(vslam_handle h, int index) {
  for (int k = 0; k < h->nslot; ++k) {
    Slot& s = h->slots[k];
    if (!s.busy && s.frames > 0 && s.serial == h->call_serial && h->call_serial > 0 && index >= s.first_frame &&
        index < s.first_frame + s.frames)
      return &s;
  }
  return nullptr;
}

int vslam_classic_stereo_extra(vslam_handle h, int index, float* u_right, float* depth, int capacity) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!h->classic) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "the handle was not created with VSLAM_MODE_CLASSIC");
  if (!u_right || !depth || capacity < 0) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  Slot* s = slot_of_last_call(h, index);
  if (!s) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "frame %d of the last process call is no longer resident", index);
  const int f = index - s->first_frame;
  const int n = std::min(s->h_lay.hdr[f].n_kp[0], h->kp_cap);
  if (capacity < n) return fail(h, VSLAM_ERR_CAPACITY, "capacity %d < %d left keypoints", capacity, n);
  DEVICE_SCOPE(h);
  if (n > 0) {
    CU_TRY(h, cudaMemcpy(u_right, s->d_uright + (size_t)f * h->kp_cap, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CU_TRY(h, cudaMemcpy(depth, s->d_depth + (size_t)f * h->kp_cap, (size_t)n * 4, cudaMemcpyDeviceToHost));
  }
  return VSLAM_OK;
}

int vslam_debug_classic_candidates(vslam_handle h, int image, int level, int32_t* xyr, int capacity, int* n_out) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!h->classic) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "the handle was not created with VSLAM_MODE_CLASSIC");
  if (!xyr || !n_out || level < 0 || level >= h->P.nlevels || image < 0 || image >= h->last_frames * h->last_eyes)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad debug_classic_candidates argument");
  DEVICE_SCOPE(h);
  Slot& s = h->slots[h->last_slot];
  const vk::CLevel& CL = h->CP.lv[level];
  std::vector<int> cnt((size_t)std::max(CL.ncells, 1));
  CU_TRY(h, cudaMemcpy(cnt.data(), s.d_ccell_cnt + (size_t)image * h->CP.total_cells + CL.cell0, sizeof(int) * CL.ncells,
                       cudaMemcpyDeviceToHost));
  std::vector<uint32_t> rec((size_t)std::max(CL.rec_cap, 1));
  CU_TRY(h, cudaMemcpy(rec.data(), s.d_ccell_rec + (size_t)image * (size_t)h->CP.rec_img_stride + CL.rec_off,
                       sizeof(uint32_t) * CL.rec_cap, cudaMemcpyDeviceToHost));
  int n = 0;
  for (int c = 0; c < CL.ncells; ++c) {
    const vk::CCell& cc = h->ccells[CL.cell0 + c];
    for (int i = 0; i < cnt[c]; ++i, ++n) {
      if (n >= capacity) continue;
      const uint32_t r = rec[(size_t)cc.rec_off + i];
      xyr[3 * n] = (int)(r & 0xfffu);
      xyr[3 * n + 1] = (int)((r >> 12) & 0xfffu);
      xyr[3 * n + 2] = (int)(r >> 24);
    }
  }
  *n_out = n;
  return n > capacity ? fail(h, VSLAM_ERR_CAPACITY, "candidate buffer too small (%d > %d)", n, capacity) : VSLAM_OK;
}

}  // extern "C"

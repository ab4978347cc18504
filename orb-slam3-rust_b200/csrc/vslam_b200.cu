// libvslam_b200.so — host pipeline and C ABI (include/vslam_b200.h) of the B200-native stereo ORB front end.
//
// Replaces, for the per-frame stereo front end only, what the reference does at
// /root/reference/src/tracking/frame/stereo.rs:37-216 (StereoProcessor::{new, process, detect_features,
// match_features}, triangulate, descriptor_distance) and the tracking-side searches
// (src/tracking/tracking_frame.rs:52-128, src/tracking/tracker.rs:863-923, 1001-1010, 1093-1158).
//
// Design: a handle owns `nslot` (default 4) pipeline slots, each with its own CUDA stream, device scratch for `max_batch`
// stereo frames and pinned host staging.  A batch is cut into chunks; chunk c runs on slot c % nslot so the
// H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernels of chunk c.  Every kernel takes the
// whole chunk (grid.y = images), so launch latency is amortised over the chunk.  No CPU fallback anywhere.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/vslam_b200.h"
#include "orb_kernels.cuh"
#include "stereo_kernels.cuh"
#include "track_kernels.cuh"
#include "tri_kernels.cuh"
#include "classic_kernels.cuh"
#include "vslam_params.h"

namespace {

constexpr int MAX_SLOTS = 6;
thread_local std::string g_create_error;

inline int cv_roundf(float v) { return (int)nearbyintf(v); }
inline int cv_round(double v) { return (int)nearbyint(v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Helper threads for staging PAGEABLE caller images (the reference hands over borrowed cv::Mat data) into the pinned
// ring: one host thread copies ~10-20 GB/s, a 2000-frame batch needs 1.44 GB per 23 ms.  parallel_for(n, f) runs f(i)
// for i in [0, n) on the pool and the calling thread and returns when all are done.
class StagingPool {
 public:
  explicit StagingPool(int nthreads) {
    for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~StagingPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void parallel_for(int n, const std::function<void(int)>& f) {
    if (n <= 0) return;
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &f;
      next_ = 0;
      total_ = n;
      pending_ = n;
    }
    cv_.notify_all();
    run_some();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run_some() {
    for (;;) {
      int i;
      const std::function<void(int)>* f;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (!fn_ || next_ >= total_) return;
        i = next_++;
        f = fn_;
      }
      (*f)(i);
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || (fn_ && next_ < total_); });
        if (stop_) return;
      }
      run_some();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  bool stop_ = false;
};

struct Slot {
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // blur runs here, concurrently with the FAST -> selection chain
  cudaEvent_t done = nullptr;
  cudaEvent_t ev_pyr = nullptr, ev_blur = nullptr;
  uint8_t* d_in = nullptr;
  uint8_t* d_pyr = nullptr;
  uint8_t* d_blur = nullptr;
  int* d_strip_cnt = nullptr;
  uint32_t* d_strip_rec = nullptr;  // k_fast: per-strip raster-ordered corners (score<<24 | y<<12 | x)
  uint32_t* d_sel1 = nullptr;       // k_select1: retainBest(2N) survivors per level
  int* d_sel1_cnt = nullptr;
  vrec_t* d_cand = nullptr;         // k_harris: survivors with Harris response
  uint32_t* d_glists = nullptr;     // index lists of the global-memory selection fallback
  vk::LvlKp* d_lvl_kp = nullptr;
  int* d_lvl_cnt = nullptr;
  unsigned char* d_res = nullptr;  // one block, carved into the SoA result arrays (d_lay)
  FrameLayout d_lay;
  // classic mode (VSLAM_MODE_CLASSIC) scratch: per-cell candidates, quadtree keys / node ids, refined stereo outputs
  int* d_ccell_cnt = nullptr;
  uint32_t* d_ccell_rec = nullptr;
  unsigned long long* d_ckeys = nullptr;
  uint32_t* d_cnode = nullptr;
  float* d_uright = nullptr;
  float* d_depth = nullptr;
  vk::BriefMaps maps;              // TMA tensor maps of this slot's blurred / pyramid slabs (k_brief_tma)
  int2* d_ml = nullptr;
  uint8_t* h_in = nullptr;         // pinned
  unsigned char* h_res = nullptr;  // pinned mirror of d_res (h_lay)
  FrameLayout h_lay;
  // Two carvings of d_res / h_res: for max_batch frames (batch chunks) and for ONE frame (vslam_process: every array of
  // the frame then sits in one contiguous range, so the whole result is a single D2H copy).  d_lay / h_lay are whichever
  // the current call uses.
  FrameLayout d_layN, h_layN, d_lay1, h_lay1;
  size_t res1_bytes = 0;
  size_t res_bytes = 0;
  bool direct_out = false;         // D2H went straight into the caller's arrays
  // in-flight chunk bookkeeping
  int frames = 0;
  int first_frame = 0;
  bool busy = false;
  bool want_results = false;
  int rows_copied = 0;             // rows of every result array the chunk's D2H copies carried
  int eyes = 2;
  uint64_t serial = 0;             // vslam_handle_s::call_serial of the call whose chunk the slot holds
};

}  // namespace

struct vslam_handle_s {
  vslam_config cfg;
  PyrP P;  // eyes = 2 template; copied and patched per launch
  StereoP S;
  int kp_cap = 0;
  int max_batch = 0;
  float harris_scale4 = 0.f;
  std::vector<uint32_t> coef;
  uint32_t* d_coef = nullptr;
  int resize_rows = VSLAM_RESIZE_ROWS;  // destination rows per k_resize lane item
  // k_resize_tile launch geometry per level (destination rows per block, column blocks, 4-column groups per column
  // block, threads, dynamic shared memory), for big chunks [0] and for single pairs [1] (shorter strips: more blocks)
  struct RTile { int rs, nstr, ncb, cgb, threads; size_t smem; bool wide; };
  RTile rtile[2][VSLAM_MAX_LEVELS];
  // VSLAM_DEBUG_SKIP (profiling aid, results become garbage): bitmask of stages NOT launched once every slot has run a few
  // chunks (their buffers then hold valid data of earlier chunks): 1 pyramid, 2 FAST, 4 select1+Harris+select2, 8 blur,
  // 16 BRIEF, 32 match.  The drop in step time is that stage's marginal cost inside the 4-slot pipeline.
  int debug_skip = 0;
  uint64_t chunks_run = 0;
  // BRIEF / orientation patches through TMA tensor-tile loads (k_brief_tma); VSLAM_BRIEF_TMA=0 selects the per-lane
  // cp.async kernel (k_brief), which is also the fallback when the driver has no cuTensorMapEncodeTiled.  Measured on
  // B200 (profiles/r02_summary.md): mid-round the TMA kernel was 14 % faster alone (4.12 vs 4.79 ms per 2000 frames) but
  // 1 % slower inside the 4-slot pipeline (its 16 KB of shared memory per block against 11 KB cost it two of sixteen
  // blocks per SM); since the packed-steering build both kernels hold 12 blocks per SM by registers and the TMA kernel
  // wins in the pipeline too (95.8k vs 95.3k frames/s, two A/B pairs on one box).
  bool brief_tma = true;
  bool select_fused = true;  // VSLAM_SELECT_FUSED=0: the three-kernel selection chain (k_select1, k_harris, k_select2)
  bool resize_tile = true;   // VSLAM_RESIZE_TILE=0 selects the per-lane cp.async ring kernel (k_resize)
  int rtile_rows = 16;       // VSLAM_RTILE_ROWS
  int blur_rows = VSLAM_BLUR_ROWS;      // max output rows per k_blur lane item
  int blur_blocks_per_sm = 0;           // persistent k_blur blocks per SM (VSLAM_BLUR_BPS)
  int sm_count = 148;
  bool pyr_full = false;                // VSLAM_PYR_FULL=1: compute whole pyramid levels, not only what is read
  bool blur_full = false;               // VSLAM_BLUR_FULL=1: blur whole levels, not only the region BRIEF can sample
  bool blur_after_fast = true;          // fork the blur after k_fast (it then fills the SMs under the latency-bound selection
                                        // chain) instead of right after the pyramid; VSLAM_BLUR_AFTER_FAST=0 restores the latter
  size_t fast_smem = 0;
  size_t fast_smem_limit = 0;  // dynamic shared memory k_fast has been opted in to for this handle
  size_t fast_smem_pad = 0;  // tuning knob VSLAM_FAST_SMEM_PAD: extra bytes per k_fast block (caps its blocks per SM)
  int sel1_cap = 0, sel1_cap_latency = 0, sel2_cap = 0;  // records the selection kernels keep in shared memory
  // vocabulary tree of vslam_vocab_upload (device)
  unsigned* voc_desc = nullptr;
  int *voc_child_begin = nullptr, *voc_child_ids = nullptr, *voc_word = nullptr;
  unsigned* voc_parent = nullptr;
  int voc_nodes = 0;
  int harris_blocks = 1;
  size_t match_rows_smem = 0;  // 0: fall back to k_match + k_match_finalize
  size_t in_img_bytes = 0;
  Slot slots[MAX_SLOTS];
  int nslot = 4;  // pipeline slots in use (VSLAM_SLOTS overrides, 2..6)
  std::string err;
  uint64_t launches = 0;
  bool profiling = false;
  float stage_ms[7] = {0};
  cudaEvent_t ev[8] = {nullptr};
  cudaEvent_t ev_batch[2] = {nullptr, nullptr};
  float last_batch_ms = 0.f;
  int last_slot = 0;       // slot holding the most recent chunk (debug accessors)
  int last_frames = 0;
  // geometry of the most recent chunk (debug accessors re-run k_fast on it)
  const uint8_t* last_in_l = nullptr;
  const uint8_t* last_in_r = nullptr;
  long long last_frame_stride = 0;
  int last_pitch = 0, last_eyes = 2;
  // frames kept in HBM for the tracking-side searches (vslam_frame_*): LEFT keypoints + descriptors + CSR FeatureGrid
  struct Retained {
    unsigned char* kp = nullptr;
    unsigned* desc = nullptr;
    int *cell_start = nullptr, *cell_items = nullptr, *cell_of = nullptr;
    int n = 0, cap = 0;
    bool used = false;
    double img_w = 0, img_h = 0;
  };
  std::vector<Retained> retained;
  StagingPool* pool = nullptr;  // created on the first pageable batch chunk (VSLAM_STAGING_THREADS, default min(8, cores/2))
  int staging_threads = -1;
  uint64_t call_serial = 0;   // incremented by every process* call; slots remember the call that filled them
  // single-pair latency path (vslam_process): H2D, the 14-launch chain with its forked blur stream and the D2H copies of
  // one stereo pair captured once into a CUDA graph over slot 0's fixed staging buffers (VSLAM_GRAPH=0 disables)
  // classic mode tables
  bool classic = false;
  vk::ClassicP CP;
  std::vector<vk::CCell> ccells;
  vk::CCell* d_ccells = nullptr;
  int quad_ncap = 0;
  size_t cstereo_smem = 0;
  int cstereo_band = 9;
  bool use_graph = true;
  cudaGraphExec_t pair_graph = nullptr;
  uint64_t pair_graph_kernels = 0;
  float last_search_ms = 0.f;
  unsigned char* arena_base = nullptr;  // grow-only device scratch of the stand-alone entry points
  size_t arena_cap = 0, arena_used = 0;
};

namespace {

int fail(vslam_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else g_create_error = buf;
  return code;
}

#define CU_TRY(h, expr)                                                                                  \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(h, VSLAM_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                  cudaGetErrorString(_e));                                                               \
  } while (0)

// Makes cfg.device current for the duration of an API call and restores the caller's device on return.
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
#define DEVICE_SCOPE(h)                  \
  DeviceGuard _dev_guard((h)->cfg.device); \
  CU_TRY(h, _dev_guard.err)

// The dynamic shared-memory opt-in is a per-function, per-device attribute shared by every handle of the process:
// only ever raise it, so that handles created for smaller images do not shrink what an older handle's launches need.
enum { SMEM_FAST = 0, SMEM_SEL1, SMEM_SEL2, SMEM_MATCH_ROWS, SMEM_RTILE, SMEM_RTILE_WIDE, SMEM_SELF, SMEM_QUAD, SMEM_CSTEREO, SMEM_NFUNC };
cudaError_t raise_smem_limit(int device, int which, size_t want) {
  static std::mutex mu;
  static size_t cur_dev[64][SMEM_NFUNC] = {{0}};
  const void* fn[SMEM_NFUNC] = {(const void*)vk::k_fast, (const void*)vk::k_select1, (const void*)vk::k_select2,
                                (const void*)vk::k_match_rows, (const void*)vk::k_resize_tile<false>,
                                (const void*)vk::k_resize_tile<true>, (const void*)vk::k_select_fused,
                                (const void*)vk::k_quadtree, (const void*)vk::k_cstereo};
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = cur_dev[device & 63][which];
  if (want <= cur) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn[which], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
  if (e == cudaSuccess) cur = want;
  return e;
}

// cv::resize INTER_LINEAR_EXACT coefficients for one axis, packed (offset | alpha1 << 16).
void linear_exact_coeffs(int src, int dst, uint32_t* out) {
  double inv_scale = (double)dst / (double)src;
  double scale = 1.0 / inv_scale;
  for (int i = 0; i < dst; ++i) {
    double f = scale * ((double)i + 0.5) - 0.5;
    int o = (int)std::floor(f);
    uint32_t ofs = 0, a1 = 0;
    if (o >= 0 && src > 1) {
      if (o < src - 1) {
        ofs = (uint32_t)o;
        a1 = (uint32_t)cv_round((f - (double)o) * 256.0);
      } else {
        ofs = (uint32_t)(src - 1);
      }
    }
    out[i] = ofs | (a1 << 16);
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// u8 tensor (x = bytes of a row, y = rows, z = images) with a box of box_w x box_h x 1; OOB reads fill with zero.
bool encode_u8_map(CUtensorMap* m, const void* base, size_t pitch, size_t rows, size_t images, size_t image_stride, unsigned box_w,
                   unsigned box_h) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t dim[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)std::max<size_t>(images, 1)};
  const cuuint64_t stride[2] = {(cuuint64_t)pitch, (cuuint64_t)image_stride};
  const cuuint32_t box[3] = {box_w, box_h, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dim, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Tables of k_brief (see orb_kernels.cuh): transposed byte-packed rBRIEF pattern, disc masks and packed int8 column offsets.
int upload_tables(vslam_handle h) {
  static const int pattern[1024] = {
#include "brief_pattern.inc"
  };
  static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
  // scratch is per call (vslam_create may run concurrently on several host threads)
  unsigned mask[4][9][32], uu[4][9];
  for (int o = 0; o < 4; ++o) {
    for (int j = 0; j < 9; ++j) {
      unsigned w = 0;
      for (int b = 0; b < 4; ++b) w |= (unsigned)(uint8_t)(int8_t)(4 * j + b - (15 + o)) << (8 * b);
      uu[o][j] = w;
      for (int r = 0; r < 32; ++r) {
        unsigned m = 0;
        if (r < 31) {
          const int d = umax[r < 15 ? 15 - r : r - 15];
          for (int b = 0; b < 4; ++b) {
            const int u = 4 * j + b - (15 + o);
            if (u >= -d && u <= d) m |= 0xffu << (8 * b);
          }
        }
        mask[o][j][r] = m;
      }
    }
  }
  {
    unsigned pat_i8[8][32];
    for (int b = 0; b < 32; ++b)
      for (int q = 0; q < 8; ++q) {
        unsigned w = 0;
        for (int c = 0; c < 4; ++c) w |= (unsigned)(uint8_t)(int8_t)pattern[4 * (8 * b + q) + c] << (8 * c);
#if VSLAM_BRIEF_PACKED
        w ^= 0x80808080u;  // bytes biased by 128 (steered_brief_byte)
#endif
        pat_i8[q][b] = w;
      }
    CU_TRY(h, cudaMemcpyToSymbol(vk::g_brief_pattern_i8, pat_i8, sizeof(pat_i8)));
  }
  {
    unsigned mask_t[8][32];  // k_brief_tma: [k][lane], lane = (j = lane & 7, g = lane >> 3), row r = g + 4 k
    for (int k = 0; k < 8; ++k)
      for (int lane = 0; lane < 32; ++lane) {
        const int j = lane & 7, r = (lane >> 3) + 4 * k;
        unsigned m = 0;
        if (r < 31) {
          const int d = umax[r < 15 ? 15 - r : r - 15];
          for (int b = 0; b < 4; ++b) {
            const int u = 4 * j + b - 15;
            if (u >= -d && u <= d) m |= 0xffu << (8 * b);
          }
        }
        mask_t[k][lane] = m;
      }
    CU_TRY(h, cudaMemcpyToSymbol(vk::g_angle_mask_t, mask_t, sizeof(mask_t)));
  }
  CU_TRY(h, cudaMemcpyToSymbol(vk::g_angle_mask, mask, sizeof(mask)));
  CU_TRY(h, cudaMemcpyToSymbol(vk::g_angle_u, uu, sizeof(uu)));
  return VSLAM_OK;
}

// Level geometry, feature quotas, strip/tile tables, slab layouts.  Mirrors cv::ORB_Impl.
int build_params(vslam_handle h) {
  const vslam_config& c = h->cfg;
  PyrP& P = h->P;
  memset(&P, 0, sizeof(P));
  P.nlevels = c.n_levels;
  P.eyes = 2;
  P.fast_threshold = c.fast_threshold;
  P.in_pitch = (int)align_up((size_t)c.width, 16);
  h->in_img_bytes = (size_t)P.in_pitch * c.height;
  P.kp_cap = h->kp_cap;
  const double sf = (double)c.scale_factor;
  size_t pyr_off = 0, blur_off = 0, ncoef = 0;
  long long rec_off = 0;
  int strip0 = 0, tile0 = 0;
  // classic mode: keypoints live >= 19 px (EDGE_THRESHOLD) from the border, so the rings the dense kernels may skip shrink
  const int eb = h->classic ? vk::kCEdge : std::max(c.edge_threshold, 4);
  for (int l = 0; l < P.nlevels; ++l) {
    LevelP& L = P.lv[l];
    float scale = (float)std::pow(sf, (double)l);
    float inv = 1.f / scale;
    L.scale = scale;
    L.w = cv_roundf((float)c.width * inv);
    L.h = cv_roundf((float)c.height * inv);
    if (L.w < 32 || L.h < 32) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "pyramid level %d is too small (%dx%d)", l, L.w, L.h);
    L.pitch = l == 0 ? P.in_pitch : (int)align_up((size_t)L.w, 16);
    if (l > 0) {
      L.pyr_off = (int)pyr_off;
      pyr_off += align_up((size_t)L.pitch * L.h, 256);
      L.xcoef_off = (int)ncoef;
      ncoef += L.w;
      L.ycoef_off = (int)ncoef;
      ncoef += L.h;
    }
    L.blur_off = (int)blur_off;
    blur_off += align_up((size_t)L.pitch * L.h, 256);
    L.ex0 = eb;
    L.ex1 = L.w - eb;
    L.ey0 = eb;
    L.ey1 = L.h - eb;
    L.fast_strip0 = strip0;
    if (L.ex1 > L.ex0 && L.ey1 > L.ey0) {
      L.fast_nstrips = (L.ey1 - L.ey0 + VSLAM_FAST_ROWS - 1) / VSLAM_FAST_ROWS;
      L.strip_cap = ((VSLAM_FAST_ROWS + 1) / 2) * ((L.ex1 - L.ex0 + 1) / 2);
    } else {
      L.ex1 = L.ex0;
      L.ey1 = L.ey0;
      L.fast_nstrips = 0;
      L.strip_cap = 0;
    }
    if (L.fast_nstrips > 128) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "image too tall: %d FAST strips > 128", L.fast_nstrips);
    strip0 += L.fast_nstrips;
    L.rec_off = (int)rec_off;
    rec_off += (long long)L.fast_nstrips * L.strip_cap;
    {
      // Pyramid levels >= 1: nothing reads the outer eb - 21 px ring (FAST reaches eb - 4, Harris eb - 4, the IC disc
      // eb - 15, the restricted blur eb - 21, the next level's needed ring maps inside this one), so k_resize skips it.
      const int m = (h->blur_full || h->pyr_full) ? 0 : std::max(eb - 21, 0);
      const int px0 = std::min(m, L.w - 1), px1 = std::max(L.w - m, px0 + 1);
      L.pyr_y0 = std::min(m, L.h - 1);
      L.pyr_y1 = std::max(L.h - m, L.pyr_y0 + 1);
      L.pyr_cg0 = px0 / 4;
      L.pyr_ncg = (px1 - 1) / 4 - L.pyr_cg0 + 1;
    }
    {
      // Keypoints lie >= eb px from every border (runByImageBorder) and the rotated BRIEF pattern reaches 18 px, so the
      // blurred level is only ever sampled in [eb - 18, w - eb + 18) x [eb - 18, h - eb + 18): blur just that (whole
      // 8-column groups).  VSLAM_BLUR_FULL=1 blurs the whole level (debug / stage comparisons of the border ring).
      const int m = h->blur_full ? 0 : std::max(eb - 18, 0);
      const int bx0 = std::min(m, L.w - 1), bx1 = std::max(L.w - m, bx0 + 1);
      L.blur_y0 = std::min(m, L.h - 1);
      L.blur_y1 = std::max(L.h - m, L.blur_y0 + 1);
      L.blur_cg0 = bx0 / 8;
      L.blur_ncg = (bx1 - 1) / 8 - L.blur_cg0 + 1;
      const int bh = L.blur_y1 - L.blur_y0;
      L.blur_nstr = (bh + h->blur_rows - 1) / h->blur_rows;
      L.blur_rs = (bh + L.blur_nstr - 1) / L.blur_nstr;
    }
    L.blur_item0 = tile0;
    tile0 += (L.blur_ncg * L.blur_nstr + 31) & ~31;
  }
  P.total_fast_strips = strip0;
  P.total_blur_items = tile0;
  P.pyr_img_stride = (long long)align_up(std::max<size_t>(pyr_off, 256), 256);
  P.blur_img_stride = (long long)align_up(blur_off, 256);
  P.rec_img_stride = (rec_off + 1) & ~1LL;
  // feature quota per level (computeKeyPoints)
  {
    float factor = (float)(1.0 / sf);
    float nd = (float)c.n_features * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)P.nlevels));
    int sum = 0;
    for (int l = 0; l < P.nlevels - 1; ++l) {
      P.lv[l].nfeat = cv_roundf(nd);
      sum += P.lv[l].nfeat;
      nd *= factor;
    }
    P.lv[P.nlevels - 1].nfeat = std::max(c.n_features - sum, 0);
  }
  if (h->classic) {
    // cell tables of ComputeKeyPointsOctTree (oracle/classic_model.cpp: classic_cells) and the quadtree roots
    vk::ClassicP& CP = h->CP;
    memset(&CP, 0, sizeof(CP));
    CP.ini_th = c.classic_ini_fast;
    CP.min_th = c.classic_min_fast;
    h->ccells.clear();
    long long rec = 0;
    int maxn = 0;
    for (int l = 0; l < P.nlevels; ++l) {
      const LevelP& L = P.lv[l];
      vk::CLevel& CL = CP.lv[l];
      const int minB = vk::kCEdge - 3, maxBX = L.w - vk::kCEdge + 3, maxBY = L.h - vk::kCEdge + 3;
      if (maxBX - minB < 8 || maxBY - minB < 8)
        return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "classic mode: pyramid level %d is too small (%dx%d)", l, L.w, L.h);
      const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
      const int nCols = std::max((int)(width / vk::kCCell), 1), nRows = std::max((int)(height / vk::kCCell), 1);
      const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
      CL.cell0 = (int)h->ccells.size();
      CL.rec_off = (int)rec;
      int lrec = 0;
      for (int i = 0; i < nRows; ++i) {
        const int iniY = minB + i * hCell;
        int maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = maxBY;
        for (int j = 0; j < nCols; ++j) {
          const int iniX = minB + j * wCell;
          int maxX = iniX + wCell + 6;
          if (iniX >= maxBX - 6) continue;
          if (maxX > maxBX) maxX = maxBX;
          const int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;
          if (x1 <= x0 || y1 <= y0) continue;
          vk::CCell cc;
          cc.x0 = (short)x0; cc.x1 = (short)x1; cc.y0 = (short)y0; cc.y1 = (short)y1;
          cc.rec_off = lrec;
          cc.cap = ((x1 - x0 + 1) / 2) * ((y1 - y0 + 1) / 2);
          lrec += cc.cap;
          h->ccells.push_back(cc);
        }
      }
      CL.ncells = (int)h->ccells.size() - CL.cell0;
      CL.rec_cap = lrec;
      rec += lrec;
      CL.W = maxBX - minB;
      CL.H = maxBY - minB;
      CL.nini = std::max((int)std::round((float)CL.W / (float)CL.H), 1);
      CL.hx = (float)CL.W / (float)CL.nini;
      maxn = std::max(maxn, P.lv[l].nfeat);
    }
    CP.total_cells = (int)h->ccells.size();
    CP.rec_img_stride = (rec + 1) & ~1LL;
    h->quad_ncap = (maxn + 16 + 7) & ~7;
    h->cstereo_band = (int)std::ceil(2.0f * P.lv[P.nlevels - 1].scale) + 2;
  }
  // resize coefficient table
  h->coef.assign(std::max<size_t>(ncoef, 1), 0);
  for (int l = 1; l < P.nlevels; ++l) {
    LevelP& L = P.lv[l];
    const LevelP& Sx = P.lv[l - 1];
    linear_exact_coeffs(Sx.w, L.w, h->coef.data() + L.xcoef_off);
    linear_exact_coeffs(Sx.h, L.h, h->coef.data() + L.ycoef_off);
  }
  // k_resize_tile geometry: exact bounds of the staged source window from the coefficient tables
  for (int v = 0; v < 2; ++v)
    for (int l = 1; l < P.nlevels; ++l) {
      const LevelP& L = P.lv[l];
      vslam_handle_s::RTile& R = h->rtile[v][l];
      const int ph = L.pyr_y1 - L.pyr_y0;
      R.rs = std::max(2, std::min(v == 0 ? h->rtile_rows : std::min(h->rtile_rows, 8), ph));
      R.nstr = (ph + R.rs - 1) / R.rs;
      R.ncb = (L.pyr_ncg + 255) / 256;
      R.cgb = (L.pyr_ncg + R.ncb - 1) / R.ncb;
      R.threads = 32 * ((R.cgb + 31) / 32);
      const uint32_t* xc = h->coef.data() + L.xcoef_off;
      const uint32_t* yc = h->coef.data() + L.ycoef_off;
      int rows_bound = 1, span_bound = 16;
      for (int st = 0; st < R.nstr; ++st) {
        const int y0 = L.pyr_y0 + st * R.rs, y1 = std::min(y0 + R.rs, L.pyr_y1) - 1;
        rows_bound = std::max(rows_bound, (int)(yc[y1] & 0xffffu) + 1 - (int)(yc[y0] & 0xffffu) + 1);
      }
      for (int cb = 0; cb < R.ncb; ++cb) {
        const int c0 = L.pyr_cg0 + cb * R.cgb, c1 = std::min(c0 + R.cgb, L.pyr_cg0 + L.pyr_ncg) - 1;
        if (c0 > c1) continue;
        const int lo = (int)(xc[4 * c0] & 0xffffu) & ~15;
        const int hi = ((int)(xc[std::min(4 * c1 + 3, L.w - 1)] & 0xffffu) + 2 + 15) & ~15;
        span_bound = std::max(span_bound, hi - lo);
      }
      R.smem = vk::resize_tile_smem_bytes(rows_bound, span_bound);
      R.wide = false;  // some 4-column group spans more than 8 source bytes (scale factors near 2)
      for (int c = L.pyr_cg0; c < L.pyr_cg0 + L.pyr_ncg; ++c)
        if ((int)(xc[std::min(4 * c + 3, L.w - 1)] & 0xffffu) + 1 - (int)(xc[std::min(4 * c, L.w - 1)] & 0xffffu) > 7) R.wide = true;
    }
  if (const char* ev = getenv("VSLAM_FAST_SMEM_PAD")) h->fast_smem_pad = (size_t)std::max(atoi(ev), 0);
  h->fast_smem = vk::fast_smem_bytes(P.in_pitch) + h->fast_smem_pad;
  h->fast_smem_limit = h->fast_smem;
  // selection scratch in shared memory: select1 sized for ~3% corner density on the largest level (typical < 2%),
  // select2 for 2N_l plus ties; anything larger is replayed in global memory.
  {
    long long epx0 = (long long)(P.lv[0].ex1 - P.lv[0].ex0) * (P.lv[0].ey1 - P.lv[0].ey0);
    // Two budgets for k_select1's records in shared memory.  Large chunks: 512 records (4 KB).  The kernel is
    // latency-bound (one warp per problem, SM issue slots 3 % busy) and hidden behind the dense kernels of the other
    // pipeline slots, so what matters is its FOOTPRINT: with a 32 KB budget its 1024 blocks held ~220 KB of shared
    // memory per SM for ~270 us per chunk and kept k_fast / k_brief / k_blur blocks of the other slots from becoming
    // resident (measured: 73.9k -> 77.7k frames/s going from 4096 to 512 records; anything from 256 to 1536 is
    // equivalent).  Problems larger than the budget replay in global memory (L1/L2-resident), same result.
    // Small chunks (single pairs): ~3 % corner density of level 0, so that every problem replays in shared memory
    // (single-pair latency 0.29 ms instead of 0.40 ms).
    h->sel1_cap_latency = (int)std::min<long long>(std::max<long long>(epx0 / 32, 1024), 24000) & ~7;
    h->sel1_cap = std::min(h->sel1_cap_latency, 512);
    if (const char* ev = getenv("VSLAM_SEL1_CAP")) h->sel1_cap = std::max(256, atoi(ev)) & ~7;  // tuning knob
    int max2n = 0;
    for (int l = 0; l < P.nlevels; ++l) max2n = std::max(max2n, 2 * P.lv[l].nfeat);
    h->sel2_cap = std::min((max2n + max2n / 4 + 64 + 7) & ~7, 16000);
    if (const char* ev = getenv("VSLAM_SEL2_CAP")) h->sel2_cap = std::max(64, atoi(ev)) & ~7;  // tuning knob
    h->harris_blocks = std::max(1, (max2n + 127) / 128);
  }
  // HarrisResponses scale: scale = 1/((1<<2)*7*255.f); scale^4 = ((s*s)*s)*s   (f32, one rounding per op)
  {
    volatile float s = 1.f / ((1 << 2) * 7 * 255.f);
    volatile float s2 = s * s;
    volatile float s3 = s2 * s;
    volatile float s4 = s3 * s;
    h->harris_scale4 = s4;
  }
  // stereo constants (stereo.rs:84-90)
  {
    const vslam_camera& cam = c.camera;
    h->S.max_disparity = (float)(cam.fx * cam.baseline / 0.1);
    h->S.min_disparity = (float)(cam.fx * cam.baseline / 40.0);
    h->S.fxb = cam.fx * cam.baseline;
    h->S.fx = cam.fx;
    h->S.fy = cam.fy;
    h->S.cx = cam.cx;
    h->S.cy = cam.cy;
  }
  return VSLAM_OK;
}

// Carve the SoA result arrays for `nfr` frames out of one block; returns the block size.
size_t carve_results(unsigned char* base, size_t nfr, int cap, FrameLayout* F) {
  size_t o = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + o : nullptr;
    o += align_up(bytes, 256);
    return p;
  };
  const size_t kc = (size_t)cap;
  F->cap = cap;
  F->kp_stride = (long long)(kc * 28);
  F->desc_stride = (long long)(kc * 32);
  F->matches_stride = (long long)(kc * 16);
  F->pts_stride = (long long)(kc * 24);
  F->valid_stride = (long long)kc;
  F->hdr = reinterpret_cast<ResHdr*>(take(nfr * sizeof(ResHdr)));
  F->kp[0] = take(nfr * kc * 28);
  F->kp[1] = take(nfr * kc * 28);
  F->desc[0] = take(nfr * kc * 32);
  F->desc[1] = take(nfr * kc * 32);
  F->matches = take(nfr * kc * 16);
  F->pts = take(nfr * kc * 24);
  F->valid = take(nfr * kc);
  return o;
}

int alloc_slot(vslam_handle h, Slot& s) {
  const PyrP& P = h->P;
  const size_t nimg = 2 * (size_t)h->max_batch, nfr = (size_t)h->max_batch;
  {
    // VSLAM_BLUR_PRIO (tuning knob): -1 = blur stream at the highest priority, 1 = lowest, 0 = default
    int lo = 0, hi = 0, p1 = 0, p2 = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (const char* ev = getenv("VSLAM_BLUR_PRIO")) {
      const int v = atoi(ev);
      p2 = v < 0 ? hi : (v > 0 ? lo : 0);
      p1 = v < 0 ? lo : (v > 0 ? hi : 0);
    }
    CU_TRY(h, cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, p1));
    CU_TRY(h, cudaStreamCreateWithPriority(&s.stream2, cudaStreamNonBlocking, p2));
  }
  CU_TRY(h, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  CU_TRY(h, cudaEventCreateWithFlags(&s.ev_pyr, cudaEventDisableTiming));
  CU_TRY(h, cudaEventCreateWithFlags(&s.ev_blur, cudaEventDisableTiming));
  CU_TRY(h, cudaMalloc(&s.d_in, nimg * h->in_img_bytes));
  CU_TRY(h, cudaMalloc(&s.d_pyr, nimg * (size_t)P.pyr_img_stride));
  CU_TRY(h, cudaMalloc(&s.d_blur, nimg * (size_t)P.blur_img_stride));
  if (!h->classic) {
    CU_TRY(h, cudaMalloc(&s.d_strip_cnt, nimg * std::max(P.total_fast_strips, 1) * sizeof(int)));
    CU_TRY(h, cudaMalloc(&s.d_strip_rec, nimg * std::max<size_t>((size_t)P.rec_img_stride, 2) * sizeof(uint32_t)));
    CU_TRY(h, cudaMalloc(&s.d_sel1, nimg * std::max<size_t>((size_t)P.rec_img_stride, 2) * sizeof(uint32_t)));
    CU_TRY(h, cudaMalloc(&s.d_sel1_cnt, nimg * P.nlevels * sizeof(int)));
    CU_TRY(h, cudaMalloc(&s.d_cand, nimg * std::max<size_t>((size_t)P.rec_img_stride, 2) * sizeof(vrec_t)));
    CU_TRY(h, cudaMalloc(&s.d_glists, nimg * std::max<size_t>((size_t)P.rec_img_stride, 2) * 2 * sizeof(uint32_t)));
  } else {
    const size_t rs = std::max<size_t>((size_t)h->CP.rec_img_stride, 2);
    CU_TRY(h, cudaMalloc(&s.d_ccell_cnt, nimg * (size_t)std::max(h->CP.total_cells, 1) * sizeof(int)));
    CU_TRY(h, cudaMalloc(&s.d_ccell_rec, nimg * rs * sizeof(uint32_t)));
    CU_TRY(h, cudaMalloc(&s.d_ckeys, nimg * rs * sizeof(unsigned long long)));
    CU_TRY(h, cudaMalloc(&s.d_cnode, nimg * rs * sizeof(uint32_t)));
    CU_TRY(h, cudaMalloc(&s.d_uright, nfr * (size_t)h->kp_cap * sizeof(float)));
    CU_TRY(h, cudaMalloc(&s.d_depth, nfr * (size_t)h->kp_cap * sizeof(float)));
  }
  CU_TRY(h, cudaMalloc(&s.d_lvl_kp, nimg * P.nlevels * (size_t)h->kp_cap * sizeof(vk::LvlKp)));
  CU_TRY(h, cudaMalloc(&s.d_lvl_cnt, nimg * P.nlevels * sizeof(int)));
  FrameLayout tmp;
  s.res_bytes = carve_results(nullptr, nfr, h->kp_cap, &tmp);
  CU_TRY(h, cudaMalloc(&s.d_res, s.res_bytes));
  CU_TRY(h, cudaMalloc(&s.d_ml, nfr * (size_t)h->kp_cap * sizeof(int2)));
  CU_TRY(h, cudaMallocHost(&s.h_in, nimg * h->in_img_bytes));
  CU_TRY(h, cudaMallocHost(&s.h_res, s.res_bytes));
  CU_TRY(h, cudaMemset(s.d_res, 0, s.res_bytes));
  carve_results(s.d_res, nfr, h->kp_cap, &s.d_layN);
  carve_results(s.h_res, nfr, h->kp_cap, &s.h_layN);
  s.res1_bytes = carve_results(s.d_res, 1, h->kp_cap, &s.d_lay1);
  carve_results(s.h_res, 1, h->kp_cap, &s.h_lay1);
  s.d_lay = s.d_layN;
  s.h_lay = s.h_layN;
  memset(&s.maps, 0, sizeof(s.maps));
  for (int l = 0; l < P.nlevels && h->brief_tma; ++l) {
    const LevelP& L = P.lv[l];
    bool ok = encode_u8_map(&s.maps.blur[l], s.d_blur + L.blur_off, (size_t)L.pitch, (size_t)L.h, nimg, (size_t)P.blur_img_stride,
                            vk::kTmaBriefW, vk::kBriefRows);
    if (ok && l > 0)
      ok = encode_u8_map(&s.maps.pyr[l], s.d_pyr + L.pyr_off, (size_t)L.pitch, (size_t)L.h, nimg, (size_t)P.pyr_img_stride,
                         vk::kTmaAngleW, vk::kAngleRows);
    if (!ok) h->brief_tma = false;
  }
  return VSLAM_OK;
}

void free_slot(Slot& s) {
  if (s.stream) cudaStreamSynchronize(s.stream);
  cudaFree(s.d_in); cudaFree(s.d_pyr); cudaFree(s.d_blur); cudaFree(s.d_strip_cnt); cudaFree(s.d_strip_rec);
  cudaFree(s.d_sel1); cudaFree(s.d_sel1_cnt); cudaFree(s.d_cand);
  cudaFree(s.d_glists); cudaFree(s.d_lvl_kp); cudaFree(s.d_lvl_cnt); cudaFree(s.d_res); cudaFree(s.d_ml);
  cudaFree(s.d_ccell_cnt); cudaFree(s.d_ccell_rec); cudaFree(s.d_ckeys); cudaFree(s.d_cnode); cudaFree(s.d_uright); cudaFree(s.d_depth);
  if (s.h_in) cudaFreeHost(s.h_in);
  if (s.h_res) cudaFreeHost(s.h_res);
  if (s.done) cudaEventDestroy(s.done);
  if (s.ev_pyr) cudaEventDestroy(s.ev_pyr);
  if (s.ev_blur) cudaEventDestroy(s.ev_blur);
  if (s.stream2) cudaStreamDestroy(s.stream2);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Slot();
}

// Blocks of the persistent k_blur launch: blur_blocks_per_sm per SM (0 = as many as the work needs, i.e. not persistent).
static int blur_grid(vslam_handle h, int nimg) {
  const int full = nimg * ((h->P.total_blur_items + 127) / 128);
  return h->blur_blocks_per_sm > 0 ? std::min(full, h->blur_blocks_per_sm * h->sm_count) : full;
}

// Enqueue every kernel of one chunk on the slot's stream.  in_l / in_r: device pointers of the first frame's
// left / right level-0 image; frames `frame_stride` bytes apart, rows `pitch` bytes apart.
int enqueue_chunk(vslam_handle h, Slot& s, int frames, int eyes, const uint8_t* in_l, const uint8_t* in_r,
                  long long frame_stride, int pitch) {
  PyrP P = h->P;
  P.eyes = eyes;
  P.in_pitch = pitch;  // pitch of the level-0 INPUT rows; the blurred level 0 keeps lv[0].pitch (sized at creation)
  P.in_frame_stride = frame_stride;
  const int nimg = frames * eyes;
  cudaStream_t st = s.stream;
  const bool prof = h->profiling;
  const int skip = (h->chunks_run++ >= (uint64_t)(3 * h->nslot)) ? h->debug_skip : 0;
  h->last_in_l = in_l;
  h->last_in_r = in_r;
  h->last_frame_stride = frame_stride;
  h->last_pitch = pitch;
  h->last_eyes = eyes;
  int evi = 0;
  auto mark = [&]() {
    if (prof) cudaEventRecord(h->ev[evi++], st);
  };
  CU_TRY(h, cudaMemsetAsync(s.d_lay.hdr, 0, sizeof(ResHdr) * (size_t)frames, st));
  mark();  // 0
  for (int l = 1; l < P.nlevels && !(skip & 1); ++l) {
    if (h->resize_tile) {
      const vslam_handle_s::RTile& R = h->rtile[nimg >= 32 ? 0 : 1][l];
      if (R.wide)
        vk::k_resize_tile<true><<<dim3(R.nstr * R.ncb, nimg), R.threads, R.smem, st>>>(P, l, R.rs, R.ncb, R.cgb, in_l, in_r,
                                                                                    s.d_pyr, h->d_coef);
      else
        vk::k_resize_tile<false><<<dim3(R.nstr * R.ncb, nimg), R.threads, R.smem, st>>>(P, l, R.rs, R.ncb, R.cgb, in_l, in_r,
                                                                                     s.d_pyr, h->d_coef);
    } else {
      const int ph = P.lv[l].pyr_y1 - P.lv[l].pyr_y0;
      const int rs = h->resize_rows, nstr = (ph + rs - 1) / rs, ncg = P.lv[l].pyr_ncg;
      dim3 grid((ncg * nstr + 127) / 128, nimg);
      vk::k_resize<<<grid, 128, 0, st>>>(P, l, (ph + nstr - 1) / nstr, in_l, in_r, s.d_pyr, h->d_coef);
    }
    h->launches++;
  }
  mark();  // 1
  if (h->classic) {
    // classic mode (NOT reference parity): per-cell FAST -> quadtree -> blur -> BRIEF -> row-band / SAD / parabola stereo.
    // The score maps live in the blurred slab until the quadtree is done, so the blur follows on the same stream.
    for (int l = 0; l < P.nlevels; ++l) {
      const int nwx = ((P.lv[l].w - vk::kCEdge + 3) >> 2) - ((vk::kCEdge & ~3) >> 2), rows = P.lv[l].h - 2 * vk::kCEdge;
      if (nwx <= 0 || rows <= 0) continue;
      const int tiles_x = (nwx + vk::kCfTW - 1) / vk::kCfTW, tiles_y = (rows + vk::kCfTH - 1) / vk::kCfTH;
      vk::k_cfast_score<<<dim3(tiles_x * tiles_y, nimg), 256, 0, st>>>(P, l, h->CP.min_th, tiles_x, in_l, in_r, s.d_pyr, s.d_blur);
      h->launches++;
    }
    vk::k_cell_select<<<dim3((h->CP.total_cells + 3) / 4, nimg), 128, 0, st>>>(P, h->CP, h->d_ccells, s.d_blur, s.d_ccell_cnt,
                                                                         s.d_ccell_rec);
    h->launches++;
    mark();  // 2
    vk::k_quadtree<<<dim3(nimg, P.nlevels), 256, vk::quadtree_smem_bytes(h->quad_ncap), st>>>(
        P, h->CP, h->quad_ncap, h->d_ccells, s.d_ccell_cnt, s.d_ccell_rec, s.d_ckeys, s.d_cnode, s.d_lvl_kp, s.d_lvl_cnt, s.d_lay);
    h->launches++;
    mark();  // 3
    vk::k_blur<<<blur_grid(h, nimg), 128, 0, st>>>(P, nimg, in_l, in_r, s.d_pyr, s.d_blur);
    h->launches++;
    mark();  // 4
    vk::k_brief<<<dim3((P.kp_cap + vk::kBriefWarps - 1) / vk::kBriefWarps, nimg), 32 * vk::kBriefWarps, 0, st>>>(
        P, in_l, in_r, s.d_pyr, s.d_blur, s.d_lvl_kp, s.d_lvl_cnt, s.d_lay);
    h->launches++;
    mark();  // 5
    if (eyes == 2) {
      vk::CStereoP CS;
      CS.mbf = (float)(h->cfg.camera.fx * h->cfg.camera.baseline);
      CS.mb = (float)h->cfg.camera.baseline;
      CS.fx = h->cfg.camera.fx; CS.fy = h->cfg.camera.fy; CS.cx = h->cfg.camera.cx; CS.cy = h->cfg.camera.cy;
      vk::k_cstereo<<<frames, 1024, h->cstereo_smem, st>>>(P, s.d_lay, CS, h->cfg.height, h->cstereo_band, in_l, in_r, s.d_pyr,
                                                           s.d_uright, s.d_depth);
      h->launches++;
    }
    mark();  // 6
    CU_TRY(h, cudaGetLastError());
    if (prof) {
      CU_TRY(h, cudaStreamSynchronize(st));
      float total = 0.f;
      for (int i = 0; i < 6; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
        h->stage_ms[i] += ms;
        total += ms;
      }
      h->stage_ms[6] += total;
    }
    return VSLAM_OK;
  }
  // The blur only needs the pyramid: outside profiling it runs on the slot's second stream, overlapping the
  // latency-bound selection chain (k_select1/k_harris/k_select2 keep the SMs mostly idle).
  const bool fork = !prof;
  cudaStream_t sb = fork ? s.stream2 : st;
  auto launch_blur_forked = [&]() -> int {
    CU_TRY(h, cudaEventRecord(s.ev_pyr, st));
    CU_TRY(h, cudaStreamWaitEvent(sb, s.ev_pyr, 0));
    if (!(skip & 8)) vk::k_blur<<<blur_grid(h, nimg), 128, 0, sb>>>(P, nimg, in_l, in_r, s.d_pyr, s.d_blur);
    h->launches++;
    CU_TRY(h, cudaEventRecord(s.ev_blur, sb));
    return VSLAM_OK;
  };
  if (fork && !h->blur_after_fast) {
    int rcb = launch_blur_forked();
    if (rcb != VSLAM_OK) return rcb;
  }
  if (P.total_fast_strips > 0 && !(skip & 2)) {
    dim3 grid(P.total_fast_strips, nimg);
    vk::k_fast<<<grid, vk::kFastThreads, vk::fast_smem_bytes(std::max(pitch, h->P.in_pitch)) + h->fast_smem_pad, st>>>(P, in_l, in_r, s.d_pyr,
                                                                                      s.d_strip_cnt, s.d_strip_rec);
    h->launches++;
  }
  if (fork && h->blur_after_fast) {  // the blur starts when k_fast is done: it fills the SMs under the selection chain
    int rcb = launch_blur_forked();
    if (rcb != VSLAM_OK) return rcb;
  }
  mark();  // 2
  if (!(skip & 4) && h->select_fused) {
    const int cap1 = nimg >= 32 ? h->sel1_cap : std::max(h->sel1_cap, h->sel1_cap_latency);
    vk::k_select_fused<<<dim3(nimg, P.nlevels), 128, vk::select_fused_smem_bytes(cap1, h->sel2_cap), st>>>(
        P, cap1, h->sel2_cap, in_l, in_r, s.d_pyr, s.d_strip_cnt, s.d_strip_rec, s.d_sel1, s.d_cand, s.d_glists, s.d_sel1_cnt,
        s.d_lvl_kp, s.d_lvl_cnt, h->harris_scale4, s.d_lay);
    h->launches += 1;
  } else if (!(skip & 4)) {
    const int cap1 = nimg >= 32 ? h->sel1_cap : std::max(h->sel1_cap, h->sel1_cap_latency);
    vk::k_select1<<<dim3(nimg, P.nlevels), 128, vk::select1_smem_bytes(cap1), st>>>(
        P, cap1, s.d_strip_cnt, s.d_strip_rec, s.d_sel1, s.d_glists, s.d_sel1_cnt);
    vk::k_harris<<<dim3(h->harris_blocks, P.nlevels, nimg), 128, 0, st>>>(P, in_l, in_r, s.d_pyr, s.d_sel1, s.d_sel1_cnt,
                                                                       s.d_cand, h->harris_scale4);
    vk::k_select2<<<dim3(nimg, P.nlevels), 32, vk::select2_smem_bytes(h->sel2_cap), st>>>(
        P, h->sel2_cap, s.d_cand, s.d_sel1_cnt, s.d_glists, s.d_lvl_kp, s.d_lvl_cnt, s.d_lay);
    h->launches += 3;
  }
  mark();  // 3
  if (!fork) {
    vk::k_blur<<<blur_grid(h, nimg), 128, 0, st>>>(P, nimg, in_l, in_r, s.d_pyr, s.d_blur);
    h->launches++;
  } else {
    CU_TRY(h, cudaStreamWaitEvent(st, s.ev_blur, 0));
  }
  mark();  // 4
  if (!(skip & 16)) {
    dim3 grid((P.kp_cap + vk::kBriefWarps - 1) / vk::kBriefWarps, nimg);
    bool tma = h->brief_tma;
    if (tma) {  // level 0 of the un-blurred pyramid is the caller's / the staged input: its maps follow the chunk
      tma = encode_u8_map(&s.maps.in_l, in_l, (size_t)pitch, (size_t)h->cfg.height, (size_t)frames, (size_t)frame_stride,
                          vk::kTmaAngleW, vk::kAngleRows) &&
            encode_u8_map(&s.maps.in_r, eyes == 2 ? in_r : in_l, (size_t)pitch, (size_t)h->cfg.height, (size_t)frames,
                          (size_t)frame_stride, vk::kTmaAngleW, vk::kAngleRows);
    }
    if (tma) vk::k_brief_tma<<<grid, 32 * vk::kBriefWarps, 0, st>>>(P, s.maps, s.d_lvl_kp, s.d_lvl_cnt, s.d_lay);
    else vk::k_brief<<<grid, 32 * vk::kBriefWarps, 0, st>>>(P, in_l, in_r, s.d_pyr, s.d_blur, s.d_lvl_kp, s.d_lvl_cnt, s.d_lay);
    h->launches++;
  }
  mark();  // 5
  if (eyes == 2 && !(skip & 32)) {
    if (h->match_rows_smem) {
      vk::k_match_rows<<<frames, 1024, h->match_rows_smem, st>>>(s.d_lay, h->S, h->cfg.height);
      h->launches += 1;
    } else {
      dim3 grid((P.kp_cap + vk::kMatchLeftPerBlock - 1) / vk::kMatchLeftPerBlock, frames);
      vk::k_match<<<grid, 256, 0, st>>>(s.d_lay, h->S, s.d_ml);
      vk::k_match_finalize<<<frames, 256, 0, st>>>(s.d_lay, h->S, s.d_ml);
      h->launches += 2;
    }
  }
  mark();  // 6
  CU_TRY(h, cudaGetLastError());
  if (prof) {
    CU_TRY(h, cudaStreamSynchronize(st));
    float total = 0.f;
    for (int i = 0; i < 6; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
      h->stage_ms[i] += ms;
      total += ms;
    }
    h->stage_ms[6] += total;
  }
  return VSLAM_OK;
}

// Hand one frame's results (host-visible SoA mirror `L`, frame f of the chunk) to the caller's buffers.
// direct: the arrays already landed in the caller's memory; only counts and status are filled in.
int unpack_frame(const FrameLayout& L, int f, bool direct, vslam_frame_out* out, uint64_t ts) {
  const ResHdr* hdr = L.hdr + f;
  out->n_left = hdr->n_kp[0];
  out->n_right = hdr->n_kp[1];
  out->n_matches = hdr->n_matches;
  out->timestamp_ns = ts;
  out->status = hdr->status;
  if (hdr->status != VSLAM_OK) return hdr->status;
  const int need = std::max(std::max(hdr->n_kp[0], hdr->n_kp[1]), hdr->n_matches);
  if (out->capacity < need) {
    out->status = VSLAM_ERR_CAPACITY;
    return VSLAM_ERR_CAPACITY;
  }
  if (direct) return VSLAM_OK;
  const size_t nl = (size_t)hdr->n_kp[0], nr = (size_t)hdr->n_kp[1], nm = (size_t)hdr->n_matches;
  if (out->kp_left) memcpy(out->kp_left, L.kp[0] + (size_t)f * L.kp_stride, nl * 28);
  if (out->kp_right) memcpy(out->kp_right, L.kp[1] + (size_t)f * L.kp_stride, nr * 28);
  if (out->desc_left) memcpy(out->desc_left, L.desc[0] + (size_t)f * L.desc_stride, nl * 32);
  if (out->desc_right) memcpy(out->desc_right, L.desc[1] + (size_t)f * L.desc_stride, nr * 32);
  if (out->matches) memcpy(out->matches, L.matches + (size_t)f * L.matches_stride, nm * 16);
  if (out->points_cam) memcpy(out->points_cam, L.pts + (size_t)f * L.pts_stride, nl * 24);
  if (out->points_valid) memcpy(out->points_valid, L.valid + (size_t)f * L.valid_stride, nl);
  return VSLAM_OK;
}

// Wait for a slot's in-flight chunk and hand its results to the caller.  The per-frame headers are always read back
// (2 KB per chunk), so a KEYPOINT_OVERFLOW of any chunk is reported even when the results stay on the device.
int retire_slot(vslam_handle h, Slot& s, vslam_frame_out* outs, const uint64_t* ts, int* first_error) {
  if (!s.busy) return VSLAM_OK;
  CU_TRY(h, cudaEventSynchronize(s.done));
  s.busy = false;
  const FrameLayout& M = s.h_lay;
  const FrameLayout& D = s.d_lay;
  if (!(s.want_results && outs)) {
    for (int f = 0; f < s.frames; ++f)
      if (M.hdr[f].status != VSLAM_OK && *first_error == VSLAM_OK) {
        *first_error = M.hdr[f].status;
        fail(h, M.hdr[f].status, "frame %d: more keypoints than keypoint_capacity (response ties)", s.first_frame + f);
      }
    return VSLAM_OK;
  }
  for (int f = 0; f < s.frames; ++f) {
    const int gf = s.first_frame + f;
    const ResHdr& hd = M.hdr[f];
    // The chunk's D2H copies carried s.rows_copied rows of every array (n_features: the keypoint count unless
    // responses tie at the cut).  Frames that hold more fetch the remaining rows now — rare, synchronous.
    const int nl = hd.n_kp[0], nr = hd.n_kp[1], nm = hd.n_matches, rc0 = s.rows_copied;
    if (hd.status == VSLAM_OK && std::max(std::max(nl, nr), nm) > rc0 && outs[gf].capacity >= std::max(std::max(nl, nr), nm)) {
      const vslam_frame_out& o = outs[gf];
      const size_t k = (size_t)f;
      auto extra = [&](unsigned char* dst_base, const unsigned char* src_base, long long stride, size_t elem, int n) -> cudaError_t {
        if (n <= rc0 || !dst_base) return cudaSuccess;
        return cudaMemcpyAsync(dst_base + (size_t)rc0 * elem, src_base + k * (size_t)stride + (size_t)rc0 * elem,
                               (size_t)(n - rc0) * elem, cudaMemcpyDeviceToHost, s.stream);
      };
      unsigned char* d0 = s.direct_out ? (unsigned char*)o.kp_left : M.kp[0] + k * M.kp_stride;
      unsigned char* d1 = s.direct_out ? (unsigned char*)o.kp_right : M.kp[1] + k * M.kp_stride;
      unsigned char* d2 = s.direct_out ? (unsigned char*)o.desc_left : M.desc[0] + k * M.desc_stride;
      unsigned char* d3 = s.direct_out ? (unsigned char*)o.desc_right : M.desc[1] + k * M.desc_stride;
      unsigned char* d4 = s.direct_out ? (unsigned char*)o.matches : M.matches + k * M.matches_stride;
      unsigned char* d5 = s.direct_out ? (unsigned char*)o.points_cam : M.pts + k * M.pts_stride;
      unsigned char* d6 = s.direct_out ? (unsigned char*)o.points_valid : M.valid + k * M.valid_stride;
      CU_TRY(h, extra(d0, D.kp[0], D.kp_stride, 28, nl));
      CU_TRY(h, extra(d2, D.desc[0], D.desc_stride, 32, nl));
      if (s.eyes == 2) {
        CU_TRY(h, extra(d1, D.kp[1], D.kp_stride, 28, nr));
        CU_TRY(h, extra(d3, D.desc[1], D.desc_stride, 32, nr));
        CU_TRY(h, extra(d4, D.matches, D.matches_stride, 16, nm));
        CU_TRY(h, extra(d5, D.pts, D.pts_stride, 24, nl));
        CU_TRY(h, extra(d6, D.valid, D.valid_stride, 1, nl));
      }
      CU_TRY(h, cudaStreamSynchronize(s.stream));
    }
    int rc = unpack_frame(M, f, s.direct_out, &outs[gf], ts ? ts[gf] : 0);
    if (rc != VSLAM_OK && *first_error == VSLAM_OK) {
      *first_error = rc;
      fail(h, rc, "frame %d: %s", gf,
           rc == VSLAM_ERR_CAPACITY ? "output capacity too small"
                                    : "more keypoints than keypoint_capacity (response ties)");
    }
  }
  return VSLAM_OK;
}

bool is_pinned_host(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// True when the caller's output arrays of frames [f0, f0+frames) are page-locked [frames][kp_cap] arrays, i.e. one
// D2H per array can land in them directly.
bool outputs_are_direct(vslam_handle h, const vslam_frame_out* o, int frames) {
  const size_t cap = (size_t)h->kp_cap;
  const vslam_frame_out& a = o[0];
  if ((size_t)a.capacity != cap) return false;
  if (!a.kp_left || !a.kp_right || !a.desc_left || !a.desc_right || !a.matches || !a.points_cam || !a.points_valid)
    return false;
  for (int f = 1; f < frames; ++f) {
    const vslam_frame_out& b = o[f];
    const size_t k = (size_t)f;
    if ((size_t)b.capacity != cap || (uint8_t*)b.kp_left != (uint8_t*)a.kp_left + k * cap * 28 ||
        (uint8_t*)b.kp_right != (uint8_t*)a.kp_right + k * cap * 28 || b.desc_left != a.desc_left + k * cap * 32 ||
        b.desc_right != a.desc_right + k * cap * 32 || (uint8_t*)b.matches != (uint8_t*)a.matches + k * cap * 16 ||
        (uint8_t*)b.points_cam != (uint8_t*)a.points_cam + k * cap * 24 || b.points_valid != a.points_valid + k * cap)
      return false;
  }
  return is_pinned_host(a.kp_left) && is_pinned_host(a.kp_right) && is_pinned_host(a.desc_left) &&
         is_pinned_host(a.desc_right) && is_pinned_host(a.matches) && is_pinned_host(a.points_cam) &&
         is_pinned_host(a.points_valid);
}

// D2H of one chunk's result arrays: one 2-D copy per array, frame f's rows [0, rows) of the [frames][capacity] arrays.
int enqueue_result_copies(vslam_handle h, Slot& s, int frames, int eyes, void* const dst[7]) {
  const FrameLayout& D = s.d_lay;
  const void* src[7] = {D.kp[0], D.kp[1], D.desc[0], D.desc[1], D.matches, D.pts, D.valid};
  const size_t pitchb[7] = {(size_t)D.kp_stride,      (size_t)D.kp_stride,  (size_t)D.desc_stride, (size_t)D.desc_stride,
                            (size_t)D.matches_stride, (size_t)D.pts_stride, (size_t)D.valid_stride};
  const size_t elem[7] = {28, 28, 32, 32, 16, 24, 1};
  // rows = n_features: cv::ORB returns exactly that many keypoints unless responses tie at the cut (retire_slot
  // fetches the extra rows of such a frame), and matches / points never outnumber the left keypoints — the
  // capacity - n_features tail rows (11 % of every array at the default capacity) never cross PCIe.
  // (classic mode: the quadtree may overshoot a level's quota by up to 3 nodes)
  const int rows = std::min(h->kp_cap, h->cfg.n_features + (h->classic ? 3 * h->P.nlevels : 0));
  s.rows_copied = rows;
  for (int a = 0; a < 7; ++a) {
    if (eyes == 1 && (a == 1 || a == 3 || a >= 4)) continue;
    CU_TRY(h, cudaMemcpy2DAsync(dst[a], pitchb[a], src[a], pitchb[a], (size_t)rows * elem[a], (size_t)frames,
                                cudaMemcpyDeviceToHost, s.stream));
  }
  return VSLAM_OK;
}

// After a CUDA failure inside run_batch: wait for everything in flight (copies may still be writing into the caller's
// arrays) and forget the chunks, so that the next call starts from idle slots.
void drain_slots(vslam_handle h) {
  for (int k = 0; k < h->nslot; ++k) {
    Slot& s = h->slots[k];
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.stream2) cudaStreamSynchronize(s.stream2);
    s.busy = false;
    s.want_results = false;
  }
  cudaGetLastError();
}

int run_batch_impl(vslam_handle h, int n, int eyes, const uint8_t* const* left, const uint8_t* const* right, size_t stride,
                   const uint8_t* left_dev, const uint8_t* right_dev, size_t image_stride, size_t row_stride,
                   const uint64_t* ts, vslam_frame_out* outs, bool device_input, int* first_error) {
  const PyrP& P = h->P;
  const int W = h->cfg.width, H = h->cfg.height;
  if (h->profiling) {
    for (int i = 0; i < 7; ++i) h->stage_ms[i] = 0.f;
  }
  int chunk = 0;
  h->slots[0].d_lay = h->slots[0].d_layN;  // (vslam_process switches slot 0 to its one-frame carving)
  h->slots[0].h_lay = h->slots[0].h_layN;
  CU_TRY(h, cudaEventRecord(h->ev_batch[0], h->slots[0].stream));
  for (int f0 = 0; f0 < n; f0 += h->max_batch, ++chunk) {
    const int frames = std::min(h->max_batch, n - f0);
    Slot& s = h->slots[chunk % h->nslot];
    int rc = retire_slot(h, s, outs, ts, first_error);
    if (rc != VSLAM_OK) return rc;
    const uint8_t *in_l, *in_r;
    long long fstride;
    int pitch;
    if (device_input) {
      in_l = left_dev + (size_t)f0 * image_stride;
      in_r = right_dev ? right_dev + (size_t)f0 * image_stride : nullptr;
      fstride = (long long)image_stride;
      pitch = (int)row_stride;
    } else {
      // planar level-0 staging: all left images of the chunk, then all right images; rows in_pitch apart
      const size_t ib = h->in_img_bytes;
      for (int e = 0; e < eyes; ++e) {
        const uint8_t* const* src = e ? right + f0 : left + f0;
        uint8_t* dst = s.d_in + (size_t)e * frames * ib;
        uint8_t* hst = s.h_in + (size_t)e * frames * ib;
        const bool pinned = is_pinned_host(src[0]) && is_pinned_host(src[frames - 1]);
        bool contiguous = true;
        for (int f = 1; f < frames && contiguous; ++f) contiguous = src[f] == src[0] + (size_t)f * H * stride;
        const bool packed = stride == (size_t)P.in_pitch;  // rows already at the device pitch: plain 1-D copies
        if (pinned && contiguous && packed) {
          CU_TRY(h, cudaMemcpyAsync(dst, src[0], (size_t)frames * ib, cudaMemcpyHostToDevice, s.stream));
        } else if (pinned && contiguous) {
          CU_TRY(h, cudaMemcpy2DAsync(dst, P.in_pitch, src[0], stride, W, (size_t)frames * H, cudaMemcpyHostToDevice,
                                      s.stream));
        } else if (pinned) {
          for (int f = 0; f < frames; ++f) {
            if (packed) CU_TRY(h, cudaMemcpyAsync(dst + (size_t)f * ib, src[f], ib, cudaMemcpyHostToDevice, s.stream));
            else
              CU_TRY(h, cudaMemcpy2DAsync(dst + (size_t)f * ib, P.in_pitch, src[f], stride, W, H, cudaMemcpyHostToDevice,
                                          s.stream));
          }
        } else {
          // pageable images (borrowed cv::Mat data): staged through the slot's pinned buffer, several host threads at once
          const int in_pitch = P.in_pitch;
          auto stage_one = [&, hst, src, ib, in_pitch](int f) {
            if (stride == (size_t)in_pitch) memcpy(hst + (size_t)f * ib, src[f], ib);
            else
              for (int y = 0; y < H; ++y) memcpy(hst + (size_t)f * ib + (size_t)y * in_pitch, src[f] + (size_t)y * stride, W);
          };
          if (frames >= 8 && h->staging_threads != 0) {
            if (!h->pool) {
              int nt = h->staging_threads;
              if (nt < 0) {  // default: 8 threads, at most half of the CPUs this process may run on
                cpu_set_t cs;
                CPU_ZERO(&cs);
                const int ncpu = sched_getaffinity(0, sizeof(cs), &cs) == 0 ? CPU_COUNT(&cs) : 2;
                nt = std::min(8, std::max(1, ncpu / 2));
              }
              h->pool = new StagingPool(std::max(nt - 1, 0));  // the calling thread works too
            }
            h->pool->parallel_for(frames, stage_one);
          } else {
            for (int f = 0; f < frames; ++f) stage_one(f);
          }
          CU_TRY(h, cudaMemcpyAsync(dst, hst, (size_t)frames * ib, cudaMemcpyHostToDevice, s.stream));
        }
      }
      in_l = s.d_in;
      in_r = s.d_in + (size_t)frames * ib;
      fstride = (long long)ib;
      pitch = P.in_pitch;
    }
    rc = enqueue_chunk(h, s, frames, eyes, in_l, in_r, fstride, pitch);
    if (rc != VSLAM_OK) return rc;
    s.want_results = outs != nullptr;
    s.direct_out = false;
    s.eyes = eyes;
    const FrameLayout& D = s.d_lay;
    const FrameLayout& M = s.h_lay;
    const size_t fr = (size_t)frames;
    CU_TRY(h, cudaMemcpyAsync(M.hdr, D.hdr, fr * sizeof(ResHdr), cudaMemcpyDeviceToHost, s.stream));
    if (outs) {
      s.direct_out = eyes == 2 && outputs_are_direct(h, outs + f0, frames);
      const vslam_frame_out& o = outs[f0];
      void* dst[7] = {s.direct_out ? (void*)o.kp_left : (void*)M.kp[0],      s.direct_out ? (void*)o.kp_right : (void*)M.kp[1],
                      s.direct_out ? (void*)o.desc_left : (void*)M.desc[0],  s.direct_out ? (void*)o.desc_right : (void*)M.desc[1],
                      s.direct_out ? (void*)o.matches : (void*)M.matches,    s.direct_out ? (void*)o.points_cam : (void*)M.pts,
                      s.direct_out ? (void*)o.points_valid : (void*)M.valid};
      rc = enqueue_result_copies(h, s, frames, eyes, dst);
      if (rc != VSLAM_OK) return rc;
    }
    CU_TRY(h, cudaEventRecord(s.done, s.stream));
    s.busy = true;
    s.frames = frames;
    s.first_frame = f0;
    s.serial = h->call_serial;
    h->last_slot = chunk % h->nslot;
    h->last_frames = frames;
  }
  // join every slot's stream into slot 0's and close the device-time bracket of this batch
  for (int k = 1; k < h->nslot; ++k)
    if (h->slots[k].busy) CU_TRY(h, cudaStreamWaitEvent(h->slots[0].stream, h->slots[k].done, 0));
  CU_TRY(h, cudaEventRecord(h->ev_batch[1], h->slots[0].stream));
  // drain in submission order
  for (int k = 0; k < h->nslot; ++k) {
    Slot& s = h->slots[(chunk + k) % h->nslot];
    int rc = retire_slot(h, s, outs, ts, first_error);
    if (rc != VSLAM_OK) return rc;
  }
  CU_TRY(h, cudaEventSynchronize(h->ev_batch[1]));
  cudaEventElapsedTime(&h->last_batch_ms, h->ev_batch[0], h->ev_batch[1]);
  return VSLAM_OK;
}

constexpr int kGraphUnavailable = -1000;  // internal: run_single_graph could not capture / instantiate

// vslam_process, latency form: one stereo pair through a CUDA graph captured once per handle.  Everything the graph
// touches has a fixed address (slot 0: device input, scratch and result buffers, pinned result mirror), so the per-call
// work is: the two H2D copies, one cudaGraphLaunch (14 kernels with the forked blur + the D2H copies), one synchronize,
// unpack.
int run_single_graph(vslam_handle h, const uint8_t* left, size_t lstride, const uint8_t* right, size_t rstride,
                     uint64_t ts, vslam_frame_out* out) {
  DEVICE_SCOPE(h);
  Slot& s = h->slots[0];
  const PyrP& P = h->P;
  const int W = h->cfg.width, H = h->cfg.height;
  const size_t ib = h->in_img_bytes;
  int first_error = VSLAM_OK;
  int rc = retire_slot(h, s, nullptr, nullptr, &first_error);  // never busy here; keeps the invariant explicit
  if (rc != VSLAM_OK) return rc;
  h->call_serial++;
  s.d_lay = s.d_lay1;  // one-frame carving: the frame's header and arrays are one contiguous range
  s.h_lay = s.h_lay1;
  // H2D outside the graph: straight from the caller's images when they are page-locked, else through the pinned staging
  for (int e = 0; e < 2; ++e) {
    const uint8_t* src = e ? right : left;
    const size_t st = e ? rstride : lstride;
    uint8_t* dev = s.d_in + (size_t)e * ib;
    if (is_pinned_host(src)) {
      if (st == (size_t)P.in_pitch) CU_TRY(h, cudaMemcpyAsync(dev, src, ib, cudaMemcpyHostToDevice, s.stream));
      else CU_TRY(h, cudaMemcpy2DAsync(dev, P.in_pitch, src, st, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, s.stream));
    } else {
      uint8_t* dst = s.h_in + (size_t)e * ib;
      if (st == (size_t)P.in_pitch) memcpy(dst, src, ib);
      else
        for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * P.in_pitch, src + (size_t)y * st, (size_t)W);
      CU_TRY(h, cudaMemcpyAsync(dev, dst, ib, cudaMemcpyHostToDevice, s.stream));
    }
  }
  if (!h->pair_graph) {
    cudaGraph_t g = nullptr;
    const uint64_t l0 = h->launches;
    CU_TRY(h, cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal));
    rc = VSLAM_OK;
    do {
      rc = enqueue_chunk(h, s, 1, 2, s.d_in, s.d_in + ib, (long long)ib, P.in_pitch);
      if (rc != VSLAM_OK) break;
      // header + every result array of the frame in ONE copy (eight copy nodes cost ~2 us each on the critical path)
      if (cudaMemcpyAsync(s.h_res, s.d_res, s.res1_bytes, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess) rc = VSLAM_ERR_CUDA;
    } while (false);
    cudaError_t ce = cudaStreamEndCapture(s.stream, &g);
    if (rc == VSLAM_OK && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&h->pair_graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc != VSLAM_OK || ce != cudaSuccess || !h->pair_graph) {
      // no graph on this driver / configuration: this handle uses plain stream launches from now on (same kernels);
      // kGraphUnavailable tells vslam_process to run this very call through that path
      cudaGetLastError();
      h->pair_graph = nullptr;
      h->use_graph = false;
      h->launches = l0;
      cudaStreamSynchronize(s.stream);
      return kGraphUnavailable;
    }
    h->pair_graph_kernels = h->launches - l0;
    h->launches = l0;
  }
  CU_TRY(h, cudaEventRecord(h->ev_batch[0], s.stream));
  CU_TRY(h, cudaGraphLaunch(h->pair_graph, s.stream));
  CU_TRY(h, cudaEventRecord(h->ev_batch[1], s.stream));
  CU_TRY(h, cudaEventRecord(s.done, s.stream));
  h->launches += h->pair_graph_kernels;
  s.busy = true;
  s.frames = 1;
  s.first_frame = 0;
  s.serial = h->call_serial;
  s.want_results = true;
  s.direct_out = false;
  s.eyes = 2;
  s.rows_copied = h->kp_cap;  // the single copy carries every row
  h->last_slot = 0;
  h->last_frames = 1;
  h->last_in_l = s.d_in;
  h->last_in_r = s.d_in + ib;
  h->last_frame_stride = (long long)ib;
  h->last_pitch = P.in_pitch;
  h->last_eyes = 2;
  rc = retire_slot(h, s, out, &ts, &first_error);
  if (rc != VSLAM_OK) {
    drain_slots(h);
    return rc;
  }
  cudaEventElapsedTime(&h->last_batch_ms, h->ev_batch[0], h->ev_batch[1]);
  return first_error;
}

int run_batch(vslam_handle h, int n, int eyes, const uint8_t* const* left, const uint8_t* const* right, size_t stride,
              const uint8_t* left_dev, const uint8_t* right_dev, size_t image_stride, size_t row_stride,
              const uint64_t* ts, vslam_frame_out* outs, bool device_input) {
  if (n <= 0) return VSLAM_OK;
  DEVICE_SCOPE(h);
  h->call_serial++;
  int first_error = VSLAM_OK;
  const int rc = run_batch_impl(h, n, eyes, left, right, stride, left_dev, right_dev, image_stride, row_stride, ts, outs,
                                device_input, &first_error);
  if (rc != VSLAM_OK) {  // a CUDA call failed mid-batch: leave no chunk in flight behind the caller's back
    drain_slots(h);
    return rc;
  }
  return first_error;
}

}  // namespace

// =============================================================================================== C ABI
extern "C" {

const char* vslam_last_error(vslam_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vslam_keypoint_capacity(vslam_handle h) { return h ? h->kp_cap : 0; }

uint64_t vslam_kernel_launches(vslam_handle h) { return h ? h->launches : 0; }

int vslam_create(const vslam_config* cfg, vslam_handle* out) {
  if (!cfg || !out) return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  if (cfg->abi_version != VSLAM_ABI_VERSION)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "ABI version mismatch: header %d, caller %d", VSLAM_ABI_VERSION,
                cfg->abi_version);
  vslam_config c = *cfg;
  if (c.n_levels == 0) c.n_levels = 8;
  if (c.scale_factor == 0.f) c.scale_factor = 1.2f;
  if (c.edge_threshold == 0) c.edge_threshold = 31;
  if (c.fast_threshold == 0) c.fast_threshold = 20;
  if (c.max_batch == 0) c.max_batch = 64;
  if (c.mode != VSLAM_MODE_REFERENCE && c.mode != VSLAM_MODE_CLASSIC)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "unknown mode %d", c.mode);
  if (c.classic_ini_fast == 0) c.classic_ini_fast = 20;
  if (c.classic_min_fast == 0) c.classic_min_fast = 7;
  if (c.mode == VSLAM_MODE_CLASSIC &&
      (c.classic_min_fast < 1 || c.classic_ini_fast < c.classic_min_fast || c.classic_ini_fast > 127 || c.n_features > 5000))
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT,
                "classic mode: need 1 <= min_fast <= ini_fast <= 127 and n_features <= 5000");
  if (c.width < 64 || c.height < 64 || c.width > 4096 || c.height > 2100)  // corner records pack x,y in 12 bits
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "unsupported image size %dx%d", c.width, c.height);
  if (c.n_features < 1 || c.n_features > 60000)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "n_features out of range");
  if (c.n_levels < 1 || c.n_levels > VSLAM_MAX_LEVELS)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "n_levels out of range (1..%d)", VSLAM_MAX_LEVELS);
  if (!(c.scale_factor > 1.0f) || c.scale_factor > 2.0f)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "scale_factor must be in (1, 2]");
  if (c.edge_threshold < 22 || c.edge_threshold > 255)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "edge_threshold must be >= 22 (BRIEF radius 18 + aligned patch loads)");
  if (c.fast_threshold < 1 || c.fast_threshold > 127)
    return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "fast_threshold out of range");
  if (c.max_batch < 1 || c.max_batch > 4096) return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "max_batch out of range");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, VSLAM_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (c.device < 0 || c.device >= ndev) return fail(nullptr, VSLAM_ERR_INVALID_ARGUMENT, "device %d out of range", c.device);
  e = cudaSetDevice(c.device);
  if (e != cudaSuccess) return fail(nullptr, VSLAM_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, c.device);
  if (prop.major != 10)
    return fail(nullptr, VSLAM_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                prop.major, prop.minor);

  vslam_handle h = new vslam_handle_s();
  h->cfg = c;
  h->classic = c.mode == VSLAM_MODE_CLASSIC;
  h->max_batch = c.max_batch;
  h->sm_count = prop.multiProcessorCount;
  h->kp_cap = c.keypoint_capacity > 0 ? c.keypoint_capacity : c.n_features + std::max(c.n_features / 8, 64);
  if (h->kp_cap > 65535) h->kp_cap = 65535;
  if (const char* ev = getenv("VSLAM_DEBUG_SKIP")) h->debug_skip = atoi(ev);
  if (const char* ev = getenv("VSLAM_GRAPH")) h->use_graph = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_STAGING_THREADS")) h->staging_threads = std::max(atoi(ev), 0);
  if (const char* ev = getenv("VSLAM_SELECT_FUSED")) h->select_fused = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_BRIEF_TMA")) h->brief_tma = atoi(ev) != 0;
  if (!tensor_map_encoder()) h->brief_tma = false;  // driver without cuTensorMapEncodeTiled: the cp.async kernel
  if (const char* ev = getenv("VSLAM_RESIZE_TILE")) h->resize_tile = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_RTILE_ROWS")) h->rtile_rows = std::min(std::max(atoi(ev), 2), 256);
  if (const char* ev = getenv("VSLAM_RESIZE_ROWS")) h->resize_rows = std::min(std::max(atoi(ev), 4), 4096);  // tuning knobs
  if (const char* ev = getenv("VSLAM_BLUR_FULL")) h->blur_full = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_PYR_FULL")) h->pyr_full = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_BLUR_AFTER_FAST")) h->blur_after_fast = atoi(ev) != 0;
  if (const char* ev = getenv("VSLAM_BLUR_BPS")) h->blur_blocks_per_sm = std::max(atoi(ev), 0);
  if (const char* ev = getenv("VSLAM_BLUR_ROWS")) h->blur_rows = std::min(std::max(atoi(ev), 8), 4096);
  int rc = build_params(h);
  if (rc != VSLAM_OK) {
    g_create_error = h->err;
    delete h;
    return rc;
  }
  auto bail = [&](int code) {
    g_create_error = h->err;
    vslam_destroy(h);
    return code;
  };
  if (cudaMalloc(&h->d_coef, h->coef.size() * sizeof(uint32_t)) != cudaSuccess ||
      cudaMemcpy(h->d_coef, h->coef.data(), h->coef.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
    fail(h, VSLAM_ERR_CUDA, "coefficient upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return bail(VSLAM_ERR_CUDA);
  }
  if (upload_tables(h) != VSLAM_OK) return bail(VSLAM_ERR_CUDA);
  if (h->classic) {
    const size_t nb = std::max<size_t>(h->ccells.size(), 1) * sizeof(vk::CCell);
    if (cudaMalloc(&h->d_ccells, nb) != cudaSuccess ||
        cudaMemcpy(h->d_ccells, h->ccells.data(), h->ccells.size() * sizeof(vk::CCell), cudaMemcpyHostToDevice) != cudaSuccess) {
      fail(h, VSLAM_ERR_CUDA, "classic cell table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      return bail(VSLAM_ERR_CUDA);
    }
    h->cstereo_smem = vk::cstereo_smem_bytes(h->kp_cap, c.height);
  }
  for (int i = 0; i < 8; ++i) cudaEventCreate(&h->ev[i]);
  cudaEventCreate(&h->ev_batch[0]);
  cudaEventCreate(&h->ev_batch[1]);
  // opt in to large dynamic shared memory where needed
  const size_t max_sel1 = vk::select1_smem_bytes(std::max(h->sel1_cap, h->sel1_cap_latency));
  const size_t max_sel = std::max(std::max(max_sel1, vk::select2_smem_bytes(h->sel2_cap)),
                                  vk::select_fused_smem_bytes(std::max(h->sel1_cap, h->sel1_cap_latency), h->sel2_cap));
  const size_t lim = (size_t)prop.sharedMemPerBlockOptin;
  if (h->fast_smem > lim || max_sel > lim) {
    fail(h, VSLAM_ERR_INVALID_ARGUMENT, "image too wide for the shared-memory tiles (%zu/%zu > %zu bytes)",
         h->fast_smem, max_sel, lim);
    return bail(VSLAM_ERR_INVALID_ARGUMENT);
  }
  // row-bucketed fused matcher when its tables fit in shared memory; otherwise the chunked pair of kernels
  {
    const size_t need = vk::match_rows_smem_bytes(h->kp_cap, c.height);
    h->match_rows_smem = (need <= lim && c.height <= 3072 && 2 * h->kp_cap >= std::min(c.height, 3072)) ? need : 0;
  }
  {
    size_t rt = 0;
    for (int v = 0; v < 2; ++v)
      for (int l = 1; l < h->P.nlevels; ++l) rt = std::max(rt, h->rtile[v][l].smem);
    if (rt > lim) h->resize_tile = false;  // extreme widths x scale factors: the ring kernel has no such limit
    const size_t want[SMEM_NFUNC] = {h->fast_smem, max_sel1, vk::select2_smem_bytes(h->sel2_cap), h->match_rows_smem,
                                     h->resize_tile ? rt : 0, h->resize_tile ? rt : 0,
                                     vk::select_fused_smem_bytes(std::max(h->sel1_cap, h->sel1_cap_latency), h->sel2_cap),
                                     h->classic ? vk::quadtree_smem_bytes(h->quad_ncap) : 0, h->classic ? h->cstereo_smem : 0};
    if (h->classic && (want[SMEM_QUAD] > lim || want[SMEM_CSTEREO] > lim)) {
      fail(h, VSLAM_ERR_INVALID_ARGUMENT, "classic mode: %zu / %zu bytes of shared memory exceed the device limit %zu",
           want[SMEM_QUAD], want[SMEM_CSTEREO], lim);
      return bail(VSLAM_ERR_INVALID_ARGUMENT);
    }
    for (int i = 0; i < SMEM_NFUNC; ++i) {
      e = raise_smem_limit(c.device, i, want[i]);
      if (e != cudaSuccess) {
        fail(h, VSLAM_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return bail(VSLAM_ERR_CUDA);
      }
    }
  }
  if (const char* ev = getenv("VSLAM_SLOTS")) h->nslot = std::min(std::max(atoi(ev), 2), MAX_SLOTS);
  for (int k = 0; k < h->nslot; ++k) {
    rc = alloc_slot(h, h->slots[k]);
    if (rc != VSLAM_OK) return bail(rc == VSLAM_ERR_CUDA ? VSLAM_ERR_OUT_OF_MEMORY : rc);
  }
  // The set-up work above (memsets, table and coefficient uploads) went to the legacy default stream; the handle's own
  // streams are non-blocking and do not order after it.  Not cudaDeviceSynchronize(): that call fails with
  // cudaErrorStreamCaptureUnsupported while ANOTHER host thread captures its handle's single-pair graph (measured:
  // tools/stress_two_handles.py, 5 of 60 concurrent creations), although that capture is thread-local.
  e = cudaStreamSynchronize(cudaStreamLegacy);
  if (e != cudaSuccess) {
    fail(h, VSLAM_ERR_CUDA, "device error after setup: %s", cudaGetErrorString(e));
    return bail(VSLAM_ERR_CUDA);
  }
  *out = h;
  return VSLAM_OK;
}

void vslam_destroy(vslam_handle h) {
  if (!h) return;
  DeviceGuard guard(h->cfg.device);
  for (int k = 0; k < MAX_SLOTS; ++k) free_slot(h->slots[k]);
  delete h->pool;
  if (h->pair_graph) cudaGraphExecDestroy(h->pair_graph);
  cudaFree(h->d_ccells);
  cudaFree(h->d_coef);
  cudaFree(h->arena_base);
  for (auto& r : h->retained) {
    cudaFree(r.kp); cudaFree(r.desc); cudaFree(r.cell_start); cudaFree(r.cell_items); cudaFree(r.cell_of);
  }
  cudaFree(h->voc_desc); cudaFree(h->voc_child_begin); cudaFree(h->voc_child_ids); cudaFree(h->voc_parent); cudaFree(h->voc_word);
  for (int i = 0; i < 8; ++i)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 2; ++i)
    if (h->ev_batch[i]) cudaEventDestroy(h->ev_batch[i]);
  delete h;
}

int vslam_process(vslam_handle h, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                  uint64_t timestamp_ns, vslam_frame_out* out) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!left || !right || !out) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (left_stride < (size_t)h->cfg.width || right_stride < (size_t)h->cfg.width)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "row stride smaller than the image width");
  if (h->use_graph && !h->profiling) {
    const int rc = run_single_graph(h, left, left_stride, right, right_stride, timestamp_ns, out);
    if (rc != kGraphUnavailable) return rc;
  }
  if (left_stride != right_stride) {
    // stage the right image with the left stride semantics by processing through the generic path twice is
    // wasteful; copy rows into the pinned staging instead (handled by run_batch via per-image stride)
    std::vector<uint8_t> tmp((size_t)h->cfg.height * left_stride);
    for (int y = 0; y < h->cfg.height; ++y)
      memcpy(tmp.data() + (size_t)y * left_stride, right + (size_t)y * right_stride, (size_t)h->cfg.width);
    const uint8_t* l[1] = {left};
    const uint8_t* r[1] = {tmp.data()};
    return run_batch(h, 1, 2, l, r, left_stride, nullptr, nullptr, 0, 0, &timestamp_ns, out, false);
  }
  const uint8_t* l[1] = {left};
  const uint8_t* r[1] = {right};
  return run_batch(h, 1, 2, l, r, left_stride, nullptr, nullptr, 0, 0, &timestamp_ns, out, false);
}

int vslam_process_batch(vslam_handle h, int n, const uint8_t* const* left, const uint8_t* const* right, size_t stride,
                        const uint64_t* timestamps_ns, vslam_frame_out* outs) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || (n > 0 && (!left || !right || !outs))) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (stride < (size_t)h->cfg.width) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "row stride smaller than the image width");
  for (int i = 0; i < n; ++i)
    if (!left[i] || !right[i]) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null image pointer at frame %d", i);
  return run_batch(h, n, 2, left, right, stride, nullptr, nullptr, 0, 0, timestamps_ns, outs, false);
}

int vslam_process_batch_device(vslam_handle h, int n, const uint8_t* left_dev, const uint8_t* right_dev,
                               size_t image_stride, size_t row_stride, const uint64_t* timestamps_ns,
                               vslam_frame_out* outs) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (n < 0 || (n > 0 && (!left_dev || !right_dev))) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (row_stride % 16 != 0 || row_stride < (size_t)h->cfg.width || ((uintptr_t)left_dev & 15) || ((uintptr_t)right_dev & 15) ||
      image_stride % 16 != 0 || image_stride < row_stride * (size_t)h->cfg.height)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "device images need 16-byte aligned bases and strides");
  if (row_stride > 8192) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "row_stride too large");
  {
    // k_fast stages whole rows: a row stride wider than align16(width) needs a larger shared-memory opt-in
    const size_t need = vk::fast_smem_bytes(std::max((int)row_stride, h->P.in_pitch)) + h->fast_smem_pad;
    if (need > h->fast_smem_limit) {
      cudaDeviceProp prop;
      DEVICE_SCOPE(h);
      CU_TRY(h, cudaGetDeviceProperties(&prop, h->cfg.device));
      if (need > (size_t)prop.sharedMemPerBlockOptin)
        return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "row_stride %zu needs %zu bytes of shared memory per FAST strip (> %zu)",
                    row_stride, need, (size_t)prop.sharedMemPerBlockOptin);
      CU_TRY(h, raise_smem_limit(h->cfg.device, SMEM_FAST, need));
      h->fast_smem_limit = need;
    }
  }
  return run_batch(h, n, 2, nullptr, nullptr, 0, left_dev, right_dev, image_stride, row_stride, timestamps_ns, outs, true);
}

int vslam_detect_and_compute(vslam_handle h, const uint8_t* image, size_t stride, vslam_keypoint* kps, uint8_t* desc,
                             int capacity, int* n_out) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  if (!image || !kps || !desc || !n_out) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "null argument");
  if (stride < (size_t)h->cfg.width) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "row stride smaller than the image width");
  vslam_frame_out o;
  memset(&o, 0, sizeof(o));
  o.capacity = capacity;
  o.kp_left = kps;
  o.desc_left = desc;
  const uint8_t* l[1] = {image};
  int rc = run_batch(h, 1, 1, l, nullptr, stride, nullptr, nullptr, 0, 0, nullptr, &o, false);
  *n_out = o.n_left;
  return rc;
}

int vslam_set_profiling(vslam_handle h, int enabled) {
  if (!h) return VSLAM_ERR_INVALID_ARGUMENT;
  h->profiling = enabled != 0;
  return VSLAM_OK;
}

float vslam_last_batch_ms(vslam_handle h) { return h ? h->last_batch_ms : 0.f; }

int vslam_last_stage_ms(vslam_handle h, float* ms7) {
  if (!h || !ms7) return VSLAM_ERR_INVALID_ARGUMENT;
  memcpy(ms7, h->stage_ms, sizeof(h->stage_ms));
  return VSLAM_OK;
}

int vslam_level_info(vslam_handle h, int level, int* w, int* h_out, float* scale, int* n_features) {
  if (!h || level < 0 || level >= h->P.nlevels) return VSLAM_ERR_INVALID_ARGUMENT;
  const LevelP& L = h->P.lv[level];
  if (w) *w = L.w;
  if (h_out) *h_out = L.h;
  if (scale) *scale = L.scale;
  if (n_features) *n_features = L.nfeat;
  return VSLAM_OK;
}

int vslam_debug_level_image(vslam_handle h, int stage, int image, int level, uint8_t* dst, size_t dst_capacity) {
  if (!h || !dst || level < 0 || level >= h->P.nlevels || image < 0 || image >= 2 * h->max_batch)
    return h ? fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad debug_level_image argument") : VSLAM_ERR_INVALID_ARGUMENT;
  const PyrP& P = h->P;
  const LevelP& L = P.lv[level];
  if (dst_capacity < (size_t)L.w * L.h) return fail(h, VSLAM_ERR_CAPACITY, "dst too small");
  DEVICE_SCOPE(h);
  Slot& s = h->slots[h->last_slot];
  const uint8_t* src;
  int pitch = L.pitch;
  if (stage == VSLAM_STAGE_BLURRED) src = s.d_blur + (size_t)image * (size_t)P.blur_img_stride + L.blur_off;
  else if (level == 0) {
    // planar staging: image index = frame * eyes + eye
    const int f = image / h->last_eyes, e = image % h->last_eyes;
    src = (e ? h->last_in_r : h->last_in_l) + (size_t)f * (size_t)h->last_frame_stride;
    pitch = h->last_pitch;
  }
  else src = s.d_pyr + (size_t)image * (size_t)P.pyr_img_stride + L.pyr_off;
  CU_TRY(h, cudaMemcpy2D(dst, L.w, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
  return VSLAM_OK;
}

int vslam_debug_fast_corners(vslam_handle h, int image, int level, int32_t* xys, int capacity, int* n_out) {
  if (!h || !xys || !n_out || level < 0 || level >= h->P.nlevels || image < 0 || image >= 2 * h->max_batch)
    return h ? fail(h, VSLAM_ERR_INVALID_ARGUMENT, "bad debug_fast_corners argument") : VSLAM_ERR_INVALID_ARGUMENT;
  const PyrP& P = h->P;
  const LevelP& L = P.lv[level];
  if (h->classic) return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "debug_fast_corners is a reference-mode accessor");
  DEVICE_SCOPE(h);
  Slot& s = h->slots[h->last_slot];
  if (h->last_frames <= 0 || image >= h->last_frames * h->last_eyes)
    return fail(h, VSLAM_ERR_INVALID_ARGUMENT, "image %d was not part of the last chunk", image);
  {
    // k_select may compact the strip slab in place (global-memory fallback): re-run k_fast on the last chunk.
    PyrP Pr = h->P;
    Pr.eyes = h->last_eyes;
    Pr.in_pitch = h->last_pitch;
    Pr.in_frame_stride = h->last_frame_stride;
    dim3 grid(Pr.total_fast_strips, h->last_frames * h->last_eyes);
    vk::k_fast<<<grid, vk::kFastThreads, vk::fast_smem_bytes(std::max(h->last_pitch, h->P.in_pitch)), s.stream>>>(
        Pr, h->last_in_l, h->last_in_r, s.d_pyr, s.d_strip_cnt, s.d_strip_rec);
    h->launches++;
    CU_TRY(h, cudaStreamSynchronize(s.stream));
  }
  std::vector<int> cnt(std::max(L.fast_nstrips, 1));
  CU_TRY(h, cudaMemcpy(cnt.data(), s.d_strip_cnt + (size_t)image * P.total_fast_strips + L.fast_strip0,
                       sizeof(int) * L.fast_nstrips, cudaMemcpyDeviceToHost));
  int n = 0;
  std::vector<uint32_t> buf(std::max(L.strip_cap, 1));
  for (int st = 0; st < L.fast_nstrips; ++st) {
    CU_TRY(h, cudaMemcpy(buf.data(),
                         s.d_strip_rec + (size_t)image * (size_t)P.rec_img_stride + L.rec_off + (size_t)st * L.strip_cap,
                         sizeof(uint32_t) * cnt[st], cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt[st]; ++i, ++n) {
      if (n >= capacity) continue;
      xys[3 * n + 0] = (int)(buf[i] & 0xfffu);
      xys[3 * n + 1] = (int)((buf[i] >> 12) & 0xfffu);
      xys[3 * n + 2] = (int)(buf[i] >> 24);
    }
  }
  *n_out = n;
  return n > capacity ? fail(h, VSLAM_ERR_CAPACITY, "corner buffer too small (%d > %d)", n, capacity) : VSLAM_OK;
}

}  // extern "C"

#include "abi_pieces.inl"

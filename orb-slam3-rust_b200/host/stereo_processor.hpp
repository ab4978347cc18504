// C++ host-side mirror of the reference's stereo front-end interface over the C ABI (include/vslam_b200.h).
//
// The reference is Rust (no Rust toolchain in the build image), so the compiled host layer is written in C++ and
// mirrors the reference's public items one to one — same names, argument meaning and error behaviour:
//
//   vslam::CameraModel                    <- CameraModel                 src/tracking/frame/camera.rs:3-21
//   vslam::FeatureSet / StereoFrame       <- FeatureSet / StereoFrame    src/tracking/frame/stereo.rs:15-29
//   vslam::StereoProcessor::create        <- StereoProcessor::new        stereo.rs:37-50   (Result -> Status + message)
//   vslam::StereoProcessor::process       <- StereoProcessor::process    stereo.rs:52-66
//   vslam::StereoProcessor::process_batch <- the caller loop             src/main.rs:66-85 (offline batch form)
//   vslam::descriptor_distance            <- descriptor_distance         stereo.rs:166-175
//   vslam::FeatureGrid                    <- FeatureGrid                 src/tracking/tracking_frame.rs:23-128
//   vslam::Frame::from_stereo             <- Frame::from_stereo          tracking_frame.rs:194-215
//   vslam::search_by_projection           <- Tracker's inline searches   src/tracking/tracker.rs:863-923, 1093-1158
//   vslam::OrbVocabulary::transform       <- OrbVocabulary::transform    src/vocabulary/mod.rs:290-340 (tree descent on the GPU)
//
// Errors: the reference returns anyhow::Result<T>; here every fallible call returns vslam::Result<T> holding either
// the value or (status, message) from vslam_last_error().  Nothing here computes on the CPU: without the CUDA
// library / a B200 every call fails.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vslam_b200.h"

namespace vslam {

constexpr uint32_t TH_HIGH = VSLAM_TH_HIGH;  // stereo.rs:10
constexpr uint32_t TH_LOW = VSLAM_TH_LOW;    // stereo.rs:11
constexpr float NN_RATIO = VSLAM_NN_RATIO;   // stereo.rs:12

struct Error {
  int status = VSLAM_OK;
  std::string message;
};

template <typename T>
class Result {
 public:
  Result(T v) : value_(std::move(v)) {}
  Result(Error e) : error_(std::move(e)) {}
  bool ok() const { return value_.has_value(); }
  explicit operator bool() const { return ok(); }
  T& value() { return *value_; }
  const T& value() const { return *value_; }
  const Error& error() const { return error_; }

 private:
  std::optional<T> value_;
  Error error_;
};

struct CameraModel {  // camera.rs:3-10
  double fx, fy, cx, cy, baseline;
  // camera.rs:13-21 — k is a row-major 3x3 intrinsic matrix
  static CameraModel from_k_and_baseline(const double k[9], double baseline) {
    return CameraModel{k[0], k[4], k[2], k[5], baseline};
  }
};

using KeyPoint = vslam_keypoint;  // cv::KeyPoint layout (28 B)
using DMatch = vslam_dmatch;      // cv::DMatch layout (16 B)
struct Descriptor {
  uint8_t bytes[32];
};
struct Point3 {
  double x, y, z;
};

struct FeatureSet {  // stereo.rs:15-19
  std::vector<KeyPoint> keypoints;
  std::vector<Descriptor> descriptors;  // N x 32, CV_8U
};

struct StereoFrame {  // stereo.rs:21-29
  FeatureSet left_features, right_features;
  std::vector<DMatch> matches_lr;
  std::vector<std::optional<Point3>> points_cam;  // per left keypoint, None without valid depth
  uint64_t timestamp_ns = 0;
};

// A borrowed 8UC1 image (the reference's `&Mat`): rows `stride` bytes apart.
struct ImageView {
  const uint8_t* data;
  size_t stride;
};

class StereoProcessor {
 public:
  // StereoProcessor::new(camera, n_features) — plus the image size and device the GPU pipeline is sized for.
  static Result<StereoProcessor> create(const CameraModel& camera, int n_features, int width = 752, int height = 480,
                                        int device = 0, int max_batch = 64) {
    vslam_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = VSLAM_ABI_VERSION;
    cfg.width = width;
    cfg.height = height;
    cfg.n_features = n_features;
    cfg.camera = vslam_camera{camera.fx, camera.fy, camera.cx, camera.cy, camera.baseline};
    cfg.device = device;
    cfg.max_batch = max_batch;
    vslam_handle h = nullptr;
    int rc = vslam_create(&cfg, &h);
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(nullptr)};
    StereoProcessor p;
    p.h_.reset(h);
    p.camera_ = camera;
    p.cap_ = vslam_keypoint_capacity(h);
    return p;
  }

  // process(&mut self, left, right, timestamp_ns) -> Result<StereoFrame>
  Result<StereoFrame> process(ImageView left, ImageView right, uint64_t timestamp_ns) {
    Buffers b(1, cap_);
    int rc = vslam_process(h_.get(), left.data, left.stride, right.data, right.stride, timestamp_ns, b.outs.data());
    if (rc != VSLAM_OK) return err(rc);
    return b.frame(0);
  }

  // n independent pairs with one common row stride; frames are pipelined on the GPU.
  Result<std::vector<StereoFrame>> process_batch(const std::vector<const uint8_t*>& left,
                                                 const std::vector<const uint8_t*>& right, size_t stride,
                                                 const std::vector<uint64_t>& timestamps_ns = {}) {
    if (left.size() != right.size()) return Error{VSLAM_ERR_INVALID_ARGUMENT, "left/right batch sizes differ"};
    const int n = (int)left.size();
    Buffers b(n, cap_);
    int rc = vslam_process_batch(h_.get(), n, left.data(), right.data(), stride,
                                 timestamps_ns.size() == left.size() ? timestamps_ns.data() : nullptr, b.outs.data());
    if (rc != VSLAM_OK) return err(rc);
    std::vector<StereoFrame> out;
    out.reserve(n);
    for (int i = 0; i < n; ++i) out.push_back(b.frame(i));
    return out;
  }

  // detect_features(&mut self, image) -> Result<FeatureSet>
  Result<FeatureSet> detect_features(ImageView image) {
    FeatureSet f;
    f.keypoints.resize(cap_);
    f.descriptors.resize(cap_);
    int n = 0;
    int rc = vslam_detect_and_compute(h_.get(), image.data, image.stride, f.keypoints.data(),
                                      reinterpret_cast<uint8_t*>(f.descriptors.data()), cap_, &n);
    if (rc != VSLAM_OK) return err(rc);
    f.keypoints.resize(n);
    f.descriptors.resize(n);
    return f;
  }

  // match_features(&self, left, right) -> Result<Vector<DMatch>>
  Result<std::vector<DMatch>> match_features(const FeatureSet& left, const FeatureSet& right) {
    std::vector<DMatch> m(left.keypoints.size() + 1);
    int n = 0;
    int rc = vslam_match_features(h_.get(), left.keypoints.data(), reinterpret_cast<const uint8_t*>(left.descriptors.data()),
                                  (int)left.keypoints.size(), right.keypoints.data(),
                                  reinterpret_cast<const uint8_t*>(right.descriptors.data()), (int)right.keypoints.size(),
                                  m.data(), &n);
    if (rc != VSLAM_OK) return err(rc);
    m.resize(n);
    return m;
  }

  const CameraModel& camera() const { return camera_; }
  vslam_handle raw() const { return h_.get(); }
  int keypoint_capacity() const { return cap_; }

 private:
  struct Deleter {
    void operator()(vslam_handle h) const { vslam_destroy(h); }
  };
  struct Buffers {  // caller-owned, capacity-checked output arrays (vslam_frame_out)
    int n, cap;
    std::vector<KeyPoint> kpl, kpr;
    std::vector<Descriptor> dl, dr;
    std::vector<DMatch> m;
    std::vector<double> pts;
    std::vector<uint8_t> valid;
    std::vector<vslam_frame_out> outs;
    Buffers(int n_, int cap_)
        : n(n_), cap(cap_), kpl((size_t)n_ * cap_), kpr((size_t)n_ * cap_), dl((size_t)n_ * cap_), dr((size_t)n_ * cap_),
          m((size_t)n_ * cap_), pts((size_t)n_ * cap_ * 3), valid((size_t)n_ * cap_), outs(n_) {
      for (int i = 0; i < n; ++i) {
        vslam_frame_out& o = outs[i];
        std::memset(&o, 0, sizeof(o));
        const size_t k = (size_t)i * cap;
        o.capacity = cap;
        o.kp_left = kpl.data() + k;
        o.kp_right = kpr.data() + k;
        o.desc_left = reinterpret_cast<uint8_t*>(dl.data() + k);
        o.desc_right = reinterpret_cast<uint8_t*>(dr.data() + k);
        o.matches = m.data() + k;
        o.points_cam = pts.data() + 3 * k;
        o.points_valid = valid.data() + k;
      }
    }
    StereoFrame frame(int i) const {
      const vslam_frame_out& o = outs[i];
      const size_t k = (size_t)i * cap;
      StereoFrame f;
      f.left_features.keypoints.assign(kpl.begin() + k, kpl.begin() + k + o.n_left);
      f.left_features.descriptors.assign(dl.begin() + k, dl.begin() + k + o.n_left);
      f.right_features.keypoints.assign(kpr.begin() + k, kpr.begin() + k + o.n_right);
      f.right_features.descriptors.assign(dr.begin() + k, dr.begin() + k + o.n_right);
      f.matches_lr.assign(m.begin() + k, m.begin() + k + o.n_matches);
      f.points_cam.resize(o.n_left);
      for (int j = 0; j < o.n_left; ++j)
        if (valid[k + j]) f.points_cam[j] = Point3{pts[3 * (k + j)], pts[3 * (k + j) + 1], pts[3 * (k + j) + 2]};
      f.timestamp_ns = o.timestamp_ns;
      return f;
    }
  };
  Error err(int rc) const { return Error{rc, vslam_last_error(h_.get())}; }

  std::unique_ptr<vslam_handle_s, Deleter> h_;
  CameraModel camera_{};
  int cap_ = 0;
};

// descriptor_distance(&desc1, &desc2) -> Result<u32>   (stereo.rs:166-175)
inline Result<uint32_t> descriptor_distance(StereoProcessor& p, const Descriptor& a, const Descriptor& b) {
  uint32_t d = 0;
  int rc = vslam_descriptor_distance(p.raw(), a.bytes, b.bytes, 1, &d);
  if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
  return d;
}

// FeatureGrid (tracking_frame.rs:23-128): 64 x 48 cells, built and queried on the GPU.
class FeatureGrid {
 public:
  static constexpr int GRID_COLS = VSLAM_GRID_COLS, GRID_ROWS = VSLAM_GRID_ROWS;
  static Result<FeatureGrid> create(StereoProcessor& p, const std::vector<KeyPoint>& keypoints, double img_width,
                                    double img_height) {
    FeatureGrid g;
    g.p_ = &p;
    g.kps_ = keypoints;
    g.w_ = img_width;
    g.h_ = img_height;
    g.cell_start_.resize(GRID_COLS * GRID_ROWS + 1);
    g.cell_items_.resize(keypoints.size() + 1);
    int rc = vslam_grid_build(p.raw(), keypoints.data(), (int)keypoints.size(), img_width, img_height,
                              g.cell_start_.data(), g.cell_items_.data());
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
    g.cell_items_.resize(keypoints.size());
    return g;
  }
  // get_features_in_area(x, y, r) -> Vec<usize>: whole cells, no radius filtering (tracking_frame.rs:104-128)
  Result<std::vector<int32_t>> get_features_in_area(double x, double y, double r) const {
    const double q[3] = {x, y, r};
    std::vector<int32_t> items(kps_.size() + 1);
    int32_t start[2] = {0, 0};
    int rc = vslam_grid_query(p_->raw(), kps_.data(), (int)kps_.size(), w_, h_, q, 1, start, items.data(), (int)items.size());
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p_->raw())};
    items.resize(start[1]);
    return items;
  }
  const std::vector<int32_t>& cell_start() const { return cell_start_; }
  const std::vector<int32_t>& cell_items() const { return cell_items_; }

 private:
  StereoProcessor* p_ = nullptr;
  std::vector<KeyPoint> kps_;
  double w_ = 752.0, h_ = 480.0;
  std::vector<int32_t> cell_start_, cell_items_;
};

// Tracking Frame (tracking_frame.rs:172-226)
struct Frame {
  uint64_t timestamp_ns = 0;
  FeatureSet features;
  std::vector<std::optional<Point3>> points_cam;
  std::vector<std::optional<uint64_t>> map_point_matches;
  std::optional<FeatureGrid> grid;

  // EuRoC dimensions are hard-coded in the reference (tracking_frame.rs:194-197)
  static Result<Frame> from_stereo(StereoProcessor& p, StereoFrame stereo) {
    return from_stereo_with_size(p, std::move(stereo), 752.0, 480.0);
  }
  static Result<Frame> from_stereo_with_size(StereoProcessor& p, StereoFrame stereo, double w, double h) {
    auto g = FeatureGrid::create(p, stereo.left_features.keypoints, w, h);
    if (!g) return g.error();
    Frame f;
    f.timestamp_ns = stereo.timestamp_ns;
    f.map_point_matches.assign(stereo.left_features.keypoints.size(), std::nullopt);
    f.features = std::move(stereo.left_features);
    f.points_cam = std::move(stereo.points_cam);
    f.grid = std::move(g.value());
    return f;
  }
  size_t num_features() const { return features.keypoints.size(); }
};

enum class SearchMode { LocalMap = VSLAM_SEARCH_LOCAL_MAP, MotionModel = VSLAM_SEARCH_MOTION_MODEL };

// The descriptor search of track_local_map (tracker.rs:863-923) / track_with_motion_model (tracker.rs:1093-1158).
// uv: host-projected pixel coordinates of the map points that passed z > 0 (the nalgebra projection stays with the
// caller); returns, per map point, the matched keypoint index or -1.
inline Result<std::vector<int32_t>> search_by_projection(StereoProcessor& p, const FeatureSet& frame_features,
                                                         const std::vector<std::pair<double, double>>& uv,
                                                         const std::vector<Descriptor>& mp_desc, SearchMode mode,
                                                         double radius = 15.0) {
  if (uv.size() != mp_desc.size()) return Error{VSLAM_ERR_INVALID_ARGUMENT, "uv / descriptor count mismatch"};
  vslam_search_params sp;
  sp.mode = (int)mode;
  sp.radius = radius;
  sp.th_high = TH_HIGH;
  sp.nn_ratio = NN_RATIO;
  sp.img_w = 752.0;
  sp.img_h = 480.0;
  sp.bound_w = 2.0 * p.camera().cx;
  sp.bound_h = 2.0 * p.camera().cy;
  std::vector<int32_t> idx(uv.size() + 1), dist(uv.size() + 1);
  int rc = vslam_search_by_projection(p.raw(), frame_features.keypoints.data(),
                                      reinterpret_cast<const uint8_t*>(frame_features.descriptors.data()),
                                      (int)frame_features.keypoints.size(), reinterpret_cast<const double*>(uv.data()),
                                      reinterpret_cast<const uint8_t*>(mp_desc.data()), (int)uv.size(), &sp, idx.data(),
                                      dist.data());
  if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
  idx.resize(uv.size());
  return idx;
}

// A tracking Frame whose left features + FeatureGrid stay in HBM (vslam_frame_retain): the searches of
// track_local_map / track_with_motion_model (tracker.rs:863-923, 1093-1158) and the reference-keyframe BFMatcher
// (tracker.rs:1001-1010) then move only (u, v) + map-point descriptors up and indices down.  RAII: released on destruction.
class ResidentFrame {
 public:
  // Right after p.process(...): keeps frame `index` of that call (EuRoC grid extent by default, tracking_frame.rs:194-197).
  static Result<ResidentFrame> retain(StereoProcessor& p, int index = 0, double img_w = 752.0, double img_h = 480.0) {
    int id = -1;
    int rc = vslam_frame_retain(p.raw(), index, img_w, img_h, &id);
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
    return ResidentFrame(&p, id, img_w, img_h);
  }
  static Result<ResidentFrame> upload(StereoProcessor& p, const FeatureSet& f, double img_w = 752.0, double img_h = 480.0) {
    int id = -1;
    int rc = vslam_frame_upload(p.raw(), f.keypoints.data(), reinterpret_cast<const uint8_t*>(f.descriptors.data()),
                                (int)f.keypoints.size(), img_w, img_h, &id);
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
    return ResidentFrame(&p, id, img_w, img_h);
  }
  ResidentFrame(ResidentFrame&& o) noexcept : p_(o.p_), id_(o.id_), w_(o.w_), h_(o.h_) { o.id_ = -1; }
  ResidentFrame& operator=(ResidentFrame&& o) noexcept {
    if (this != &o) {
      release();
      p_ = o.p_; id_ = o.id_; w_ = o.w_; h_ = o.h_;
      o.id_ = -1;
    }
    return *this;
  }
  ResidentFrame(const ResidentFrame&) = delete;
  ResidentFrame& operator=(const ResidentFrame&) = delete;
  ~ResidentFrame() { release(); }
  int id() const { return id_; }
  int num_features() const { return vslam_frame_num_features(p_->raw(), id_); }

  Result<std::vector<int32_t>> search_by_projection(const std::vector<std::pair<double, double>>& uv,
                                                    const std::vector<Descriptor>& mp_desc, SearchMode mode,
                                                    double radius = 15.0) const {
    if (uv.size() != mp_desc.size()) return Error{VSLAM_ERR_INVALID_ARGUMENT, "uv / descriptor count mismatch"};
    vslam_search_params sp;
    sp.mode = (int)mode;
    sp.radius = radius;
    sp.th_high = TH_HIGH;
    sp.nn_ratio = NN_RATIO;
    sp.img_w = w_;
    sp.img_h = h_;
    sp.bound_w = 2.0 * p_->camera().cx;
    sp.bound_h = 2.0 * p_->camera().cy;
    std::vector<int32_t> idx(uv.size() + 1), dist(uv.size() + 1);
    int rc = vslam_frame_search_by_projection(p_->raw(), id_, reinterpret_cast<const double*>(uv.data()),
                                              reinterpret_cast<const uint8_t*>(mp_desc.data()), (int)uv.size(), &sp, idx.data(),
                                              dist.data());
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p_->raw())};
    idx.resize(uv.size());
    return idx;
  }
  // BFMatcher(NORM_HAMMING, true).train_match(query = this keyframe, train = frame)
  Result<std::vector<DMatch>> bf_match_crosscheck(const ResidentFrame& train) const {
    std::vector<DMatch> m((size_t)std::max(num_features(), 1));
    int n = 0;
    int rc = vslam_frame_bf_match_crosscheck(p_->raw(), id_, train.id_, m.data(), &n);
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p_->raw())};
    m.resize((size_t)n);
    return m;
  }

 private:
  ResidentFrame(StereoProcessor* p, int id, double w, double h) : p_(p), id_(id), w_(w), h_(h) {}
  void release() {
    if (id_ >= 0 && p_) vslam_frame_release(p_->raw(), id_);
    id_ = -1;
  }
  StereoProcessor* p_ = nullptr;
  int id_ = -1;
  double w_ = 752.0, h_ = 480.0;
};

// OrbVocabulary (src/vocabulary/mod.rs:79-96): the tree as the reference's loader builds it (node 0 = root).  upload()
// flattens it once for vslam_vocab_upload; transform() mirrors OrbVocabulary::transform (mod.rs:290-331): the descent of
// every descriptor row runs on the GPU (vslam_bow_quantize), the BowVector / FeatureVector accumulation and the L1
// normalisation stay here (the reference sums f64 weights in HashMap order; std::map order is used instead, which is
// one of the orders the reference can take).
struct VocabNode {  // mod.rs:40-58
  uint32_t parent = 0xFFFFFFFFu;
  std::vector<uint32_t> children;
  Descriptor descriptor{};
  double weight = 0.0;
  int32_t word_id = -1;  // Option<u32>: None -> -1
};
using BowVector = std::map<uint32_t, double>;
using FeatureVector = std::map<uint32_t, std::vector<size_t>>;

class OrbVocabulary {
 public:
  explicit OrbVocabulary(std::vector<VocabNode> nodes) : nodes_(std::move(nodes)) {}

  Result<bool> upload(StereoProcessor& p) const {
    const size_t n = nodes_.size();
    std::vector<uint8_t> desc(n * 32);
    std::vector<int32_t> begin(n + 1, 0), ids, word(n);
    std::vector<uint32_t> parent(n);
    for (size_t i = 0; i < n; ++i) {
      std::memcpy(desc.data() + 32 * i, &nodes_[i].descriptor, 32);
      begin[i + 1] = begin[i] + (int32_t)nodes_[i].children.size();
      for (uint32_t c : nodes_[i].children) ids.push_back((int32_t)c);
      parent[i] = nodes_[i].parent;
      word[i] = nodes_[i].word_id;
    }
    if (ids.empty()) ids.push_back(0);
    int rc = vslam_vocab_upload(p.raw(), (int)n, desc.data(), begin.data(), ids.data(), parent.data(), word.data());
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
    return true;
  }

  // (BowVector, FeatureVector) of the descriptor rows; levels_up as in the reference (typically 4 for L = 6).
  Result<std::pair<BowVector, FeatureVector>> transform(StereoProcessor& p, const std::vector<Descriptor>& descriptors,
                                                       int levels_up) const {
    const size_t n = descriptors.size();
    std::vector<uint32_t> word(n + 1), leaf(n + 1), node(n + 1);
    int rc = vslam_bow_quantize(p.raw(), reinterpret_cast<const uint8_t*>(descriptors.data()), (int)n, levels_up, word.data(),
                                leaf.data(), node.data());
    if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
    BowVector bow;
    FeatureVector feat;
    for (size_t i = 0; i < n; ++i) {
      bow[word[i]] += nodes_[leaf[i]].weight;  // mod.rs:308-310
      feat[node[i]].push_back(i);              // mod.rs:313-314
    }
    double sum = 0.0;
    for (const auto& kv : bow) sum += kv.second;
    if (sum > 0.0)
      for (auto& kv : bow) kv.second /= sum;  // mod.rs:318-324
    return std::make_pair(std::move(bow), std::move(feat));
  }

 private:
  std::vector<VocabNode> nodes_;
};

// search_for_triangulation / search_for_triangulation_bow (local_mapping/triangulation.rs:401-530, 541-658): epipolar-gated
// greedy keyframe <-> keyframe Hamming match.  The caller computes F (row-major, exactly as check_epipolar_constraint,
// :668-680) and the epipole (:421-428) with its own linear algebra; has_mp*/has_stereo1 are the `is_some()` tests of the
// reference's Option slices.  `fv_groups` (optional): the FeatureVector nodes present in both keyframes, in the order the
// caller iterates them, as (indices1, indices2) — selects the _bow variant.  Returns (idx1, idx2) in emission order.
struct TriangulationGeometry {
  double F[9];
  double epipole[2];
};
inline Result<std::vector<std::pair<size_t, size_t>>> search_for_triangulation(
    StereoProcessor& p, const FeatureSet& kf1, const std::vector<bool>& has_mp1, const std::vector<bool>& has_stereo1,
    const FeatureSet& kf2, const std::vector<bool>& has_mp2, const TriangulationGeometry& geo, uint32_t max_dist,
    const std::vector<std::pair<std::vector<size_t>, std::vector<size_t>>>* fv_groups = nullptr) {
  const size_t n1 = kf1.keypoints.size(), n2 = kf2.keypoints.size();
  std::vector<uint8_t> f1(n1 + 1, 0), f2(n2 + 1, 0);
  for (size_t i = 0; i < n1; ++i)  // mp1.get(i).map_or(false, ..): a short slice reads as None
    f1[i] = (uint8_t)((i < has_mp1.size() && has_mp1[i] ? 1 : 0) | (i < has_stereo1.size() && has_stereo1[i] ? 2 : 0));
  for (size_t i = 0; i < n2; ++i) f2[i] = i < has_mp2.size() && has_mp2[i] ? 1 : 0;
  vslam_tri_params tp;
  std::memcpy(tp.F, geo.F, sizeof(tp.F));
  tp.epipole[0] = geo.epipole[0];
  tp.epipole[1] = geo.epipole[1];
  tp.max_dist = max_dist;
  tp.radius = 100.0f;
  tp.image_width = (uint32_t)(p.camera().cx * 2.0);
  tp.image_height = (uint32_t)(p.camera().cy * 2.0);
  int n = 0, rc;
  std::vector<int32_t> pairs;
  if (!fv_groups) {
    pairs.resize(2 * n1 + 2);
    rc = vslam_search_for_triangulation(p.raw(), kf1.keypoints.data(), reinterpret_cast<const uint8_t*>(kf1.descriptors.data()),
                                        f1.data(), (int)n1, kf2.keypoints.data(),
                                        reinterpret_cast<const uint8_t*>(kf2.descriptors.data()), f2.data(), (int)n2, &tp,
                                        pairs.data(), &n);
  } else {
    std::vector<int32_t> s1{0}, i1, s2{0}, i2;
    for (const auto& g : *fv_groups) {
      for (size_t v : g.first) i1.push_back((int32_t)v);
      for (size_t v : g.second) i2.push_back((int32_t)v);
      s1.push_back((int32_t)i1.size());
      s2.push_back((int32_t)i2.size());
    }
    pairs.resize(2 * i1.size() + 2);
    if (i1.empty()) i1.push_back(0);
    if (i2.empty()) i2.push_back(0);
    rc = vslam_search_for_triangulation_bow(p.raw(), kf1.keypoints.data(),
                                            reinterpret_cast<const uint8_t*>(kf1.descriptors.data()), f1.data(), (int)n1,
                                            kf2.keypoints.data(), reinterpret_cast<const uint8_t*>(kf2.descriptors.data()),
                                            f2.data(), (int)n2, &tp, (int)fv_groups->size(), s1.data(), i1.data(), s2.data(),
                                            i2.data(), pairs.data(), &n);
  }
  if (rc != VSLAM_OK) return Error{rc, vslam_last_error(p.raw())};
  std::vector<std::pair<size_t, size_t>> out((size_t)n);
  for (int k = 0; k < n; ++k) out[k] = {(size_t)pairs[2 * k], (size_t)pairs[2 * k + 1]};
  return out;
}

}  // namespace vslam

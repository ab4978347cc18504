// Compiled host-side consumer of the C ABI through the C++ mirror of the reference interface
// (stereo_processor.hpp).  Reads two raw 8-bit images (w*h bytes each), runs StereoProcessor::process and prints a
// checksum line that tests/test_gpu_host_cpp.py compares with the oracle.  Fails loudly without a GPU.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "stereo_processor.hpp"

static std::vector<uint8_t> read_file(const char* path, size_t n) {
  std::vector<uint8_t> v(n);
  std::ifstream f(path, std::ios::binary);
  if (!f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)n)) {
    std::fprintf(stderr, "cannot read %zu bytes from %s\n", n, path);
    std::exit(2);
  }
  return v;
}

static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const uint8_t* b = static_cast<const uint8_t*>(p);
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: host_demo left.raw right.raw width height [n_features]\n");
    return 2;
  }
  const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), nf = argc > 5 ? std::atoi(argv[5]) : 1200;
  auto left = read_file(argv[1], (size_t)w * h), right = read_file(argv[2], (size_t)w * h);
  vslam::CameraModel cam{458.654, 457.296, 367.215, 248.375, 0.110074};
  auto sp = vslam::StereoProcessor::create(cam, nf, w, h, 0, 4);
  if (!sp) {
    std::fprintf(stderr, "StereoProcessor::create failed (%d): %s\n", sp.error().status, sp.error().message.c_str());
    return 1;
  }
  auto fr = sp.value().process({left.data(), (size_t)w}, {right.data(), (size_t)w}, 42);
  if (!fr) {
    std::fprintf(stderr, "process failed (%d): %s\n", fr.error().status, fr.error().message.c_str());
    return 1;
  }
  const vslam::StereoFrame& f = fr.value();
  uint64_t hk = fnv1a(f.left_features.keypoints.data(), f.left_features.keypoints.size() * sizeof(vslam::KeyPoint));
  hk = fnv1a(f.right_features.keypoints.data(), f.right_features.keypoints.size() * sizeof(vslam::KeyPoint), hk);
  uint64_t hd = fnv1a(f.left_features.descriptors.data(), f.left_features.descriptors.size() * 32);
  hd = fnv1a(f.right_features.descriptors.data(), f.right_features.descriptors.size() * 32, hd);
  uint64_t hm = fnv1a(f.matches_lr.data(), f.matches_lr.size() * sizeof(vslam::DMatch));
  size_t valid = 0;
  for (auto& p : f.points_cam) valid += p.has_value();
  auto frame = vslam::Frame::from_stereo(sp.value(), f);
  size_t in_area = 0;
  if (frame) {
    auto q = frame.value().grid->get_features_in_area(376.0, 240.0, 15.0);
    if (q) in_area = q.value().size();
  }
  // BoW: a small deterministic k = 5, L = 2 vocabulary (descriptor bytes from an LCG), OrbVocabulary::transform of the
  // left descriptors with levels_up = 1; the checksum covers the (word id, weight-free) assignment order
  uint64_t hb = 0;
  size_t n_words = 0, n_groups = 0;
  {
    std::vector<vslam::VocabNode> nodes(1 + 5 + 25);
    uint32_t lcg = 12345u, wc = 0;
    for (size_t i = 1; i < nodes.size(); ++i) {
      const uint32_t parent = i <= 5 ? 0u : (uint32_t)((i - 6) / 5 + 1);
      nodes[i].parent = parent;
      nodes[parent].children.push_back((uint32_t)i);
      for (int b = 0; b < 32; ++b) {
        lcg = lcg * 1664525u + 1013904223u;
        nodes[i].descriptor.bytes[b] = (uint8_t)(lcg >> 24);
      }
      if (i > 5) {
        nodes[i].word_id = (int32_t)wc++;
        nodes[i].weight = 1.0 + 0.125 * (double)(i % 7);
      }
    }
    vslam::OrbVocabulary voc(std::move(nodes));
    auto up = voc.upload(sp.value());
    auto tr = up ? voc.transform(sp.value(), f.left_features.descriptors, 1)
                 : vslam::Result<std::pair<vslam::BowVector, vslam::FeatureVector>>(up.error());
    if (!tr) {
      std::fprintf(stderr, "OrbVocabulary failed (%d): %s\n", tr.error().status, tr.error().message.c_str());
      return 1;
    }
    n_words = tr.value().first.size();
    n_groups = tr.value().second.size();
    hb = 1469598103934665603ull;
    for (const auto& kv : tr.value().second)  // node id -> feature indices (ascending ids, ascending indices)
      for (size_t idx : kv.second) {
        const uint32_t rec[2] = {kv.first, (uint32_t)idx};
        hb = fnv1a(rec, sizeof(rec), hb);
      }
  }
  // Resident frame (vslam_frame_retain right after process) + local-map search with the frame's own stereo points as the
  // map (uv = the keypoints' positions shifted by (+3, -2) px) + search_for_triangulation between the two eyes with the
  // fundamental matrix of a rectified pair (F and the epipole are the caller's job, as in the Rust binding)
  uint64_t hs = 0, ht = 0;
  size_t n_found = 0, n_tri = 0;
  {
    auto rf = vslam::ResidentFrame::retain(sp.value(), 0);
    if (!rf) {
      std::fprintf(stderr, "ResidentFrame::retain failed (%d): %s\n", rf.error().status, rf.error().message.c_str());
      return 1;
    }
    std::vector<std::pair<double, double>> uv;
    std::vector<vslam::Descriptor> md;
    for (size_t i = 0; i < f.points_cam.size(); ++i)
      if (f.points_cam[i].has_value()) {
        uv.push_back({(double)f.left_features.keypoints[i].x + 3.0, (double)f.left_features.keypoints[i].y - 2.0});
        md.push_back(f.left_features.descriptors[i]);
      }
    auto idx = rf.value().search_by_projection(uv, md, vslam::SearchMode::LocalMap);
    if (!idx) {
      std::fprintf(stderr, "search_by_projection failed (%d): %s\n", idx.error().status, idx.error().message.c_str());
      return 1;
    }
    hs = fnv1a(idx.value().data(), idx.value().size() * sizeof(int32_t));
    for (int32_t v : idx.value()) n_found += v >= 0;
    vslam::TriangulationGeometry geo{};
    // F = K^-T [t]x K^-1 with t = (baseline, 0, 0):  only F[1][2] = -b / fy and F[2][1] = b / fy ... written out
    const double b = cam.baseline;
    const double Kinv[3][3] = {{1.0 / cam.fx, 0.0, -cam.cx / cam.fx}, {0.0, 1.0 / cam.fy, -cam.cy / cam.fy}, {0.0, 0.0, 1.0}};
    const double tx[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, -b}, {0.0, b, 0.0}};
    double tmp[3][3] = {}, F[3][3] = {};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) tmp[i][j] += tx[i][k] * Kinv[k][j];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        for (int k = 0; k < 3; ++k) F[i][j] += Kinv[k][i] * tmp[k][j];
    for (int i = 0; i < 9; ++i) geo.F[i] = F[i / 3][i % 3];
    geo.epipole[0] = 1e9;
    geo.epipole[1] = cam.cy;
    std::vector<bool> none1(f.left_features.keypoints.size(), false), none2(f.right_features.keypoints.size(), false);
    auto tri = vslam::search_for_triangulation(sp.value(), f.left_features, none1, none1, f.right_features, none2, geo, 50);
    if (!tri) {
      std::fprintf(stderr, "search_for_triangulation failed (%d): %s\n", tri.error().status, tri.error().message.c_str());
      return 1;
    }
    n_tri = tri.value().size();
    ht = 1469598103934665603ull;
    for (const auto& pr : tri.value()) {
      const uint32_t rec[2] = {(uint32_t)pr.first, (uint32_t)pr.second};
      ht = fnv1a(rec, sizeof(rec), ht);
    }
  }
  std::printf("found=%zu search=%016llx n_tri=%zu tri=%016llx ", n_found, (unsigned long long)hs, n_tri, (unsigned long long)ht);
  std::printf("bow_words=%zu bow_groups=%zu bow=%016llx ", n_words, n_groups, (unsigned long long)hb);
  std::printf("n_left=%zu n_right=%zu n_matches=%zu n_points=%zu kp=%016llx desc=%016llx matches=%016llx area=%zu ts=%llu\n",
              f.left_features.keypoints.size(), f.right_features.keypoints.size(), f.matches_lr.size(), valid,
              (unsigned long long)hk, (unsigned long long)hd, (unsigned long long)hm, in_area,
              (unsigned long long)f.timestamp_ns);
  return 0;
}

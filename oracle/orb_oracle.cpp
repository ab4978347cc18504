// oracle/orb_oracle.cpp — TEST INFRASTRUCTURE ONLY (never linked into / called by the product path).
//
// CPU restatement of what the reference's `StereoProcessor::detect_features`
// (/root/reference/src/tracking/frame/stereo.rs:68-78, configured at :37-50) actually computes.
// The reference delegates the arithmetic to the third-party, un-vendored OpenCV `cv::ORB`
// (crate `opencv 0.98.1`, Cargo.lock:5576-5577; OpenCV library version unpinned by the reference).
// This file restates cv::ORB's published algorithm stage by stage and is PINNED
// by running the very same C++ function through Python `cv2` (4.13.0 in this image):
// tests/test_oracle_vs_cv2.py demands bit-identical keypoints (order included) and descriptors.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
//
// Build: g++ -O2 -ffp-contract=off -mfma -shared -fPIC (see oracle/Makefile). -ffp-contract=off is
// REQUIRED: every float op below is individually rounded unless an explicit fmaf() is written.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

// cvRound: round-half-to-even (SSE cvtsd2si / lrint semantics).
inline int cv_round(double v) { return (int)std::nearbyint(v); }
inline int cv_roundf(float v) { return (int)std::nearbyintf(v); }

const int kBriefPattern[1024] = {
#include "brief_pattern.inc"
};

// umax table of the 31-px circular patch (half = 15), SURVEY.md §8.A4.
const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// FAST-9/16 Bresenham circle, OpenCV order (dx,dy) — SURVEY.md §8.A2.
const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                            {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

}  // namespace

extern "C" {

struct OrcLevel {
  int w, h;
  float scale;  // layerScale[l] = (float)pow((double)scaleFactor, l)
  int nfeatures;
};

// 12-byte record permuted by retainBest.
struct OrcCorner {
  int x, y;
  float resp;
};

// cv::KeyPoint memory layout (28 B).
struct OrcKeyPoint {
  float x, y, size, angle, response;
  int octave, class_id;
};

// A0 — level geometry and per-level feature quota (cv::ORB_Impl::detectAndCompute / computeKeyPoints).
void orc_level_geometry(int W, int H, int nfeatures, int nlevels, float scale_factor, OrcLevel* out) {
  const double sf = (double)scale_factor;  // ORB stores scaleFactor as double; the reference passes 1.2f
  for (int l = 0; l < nlevels; ++l) {
    float scale = (float)std::pow(sf, (double)l);
    float inv = 1.f / scale;
    out[l].scale = scale;
    out[l].w = cv_roundf((float)W * inv);
    out[l].h = cv_roundf((float)H * inv);
  }
  float factor = (float)(1.0 / sf);
  float nd = (float)nfeatures * (1.f - factor) / (1.f - (float)std::pow((double)factor, (double)nlevels));
  int sum = 0;
  for (int l = 0; l < nlevels - 1; ++l) {
    out[l].nfeatures = cv_roundf(nd);
    sum += out[l].nfeatures;
    nd *= factor;
  }
  out[nlevels - 1].nfeatures = std::max(nfeatures - sum, 0);
}

// A1 — cv::resize(..., INTER_LINEAR_EXACT) for 8UC1: Q8.8 horizontal pass, Q16.16 vertical pass.
static void linear_exact_coeffs(int src, int dst, std::vector<int>& ofs, std::vector<int>& a1) {
  ofs.resize(dst);
  a1.resize(dst);
  double inv_scale = (double)dst / (double)src;
  double scale = 1.0 / inv_scale;
  for (int i = 0; i < dst; ++i) {
    double f = scale * ((double)i + 0.5) - 0.5;
    int o = (int)std::floor(f);
    if (o >= 0 && src > 1) {
      if (o < src - 1) {
        ofs[i] = o;
        a1[i] = cv_round((f - (double)o) * 256.0);
      } else {  // replicate last sample
        ofs[i] = src - 1;
        a1[i] = 0;
      }
    } else {  // replicate first sample
      ofs[i] = 0;
      a1[i] = 0;
    }
  }
}

void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh,
                             int dstride) {
  std::vector<int> xo, xa, yo, ya;
  linear_exact_coeffs(sw, dw, xo, xa);
  linear_exact_coeffs(sh, dh, yo, ya);
  std::vector<uint32_t> h0(dw), h1(dw);
  for (int y = 0; y < dh; ++y) {
    int o = yo[y];
    int o1 = std::min(o + 1, sh - 1);
    const uint8_t* r0 = src + (size_t)o * sstride;
    const uint8_t* r1 = src + (size_t)o1 * sstride;
    uint32_t b1 = (uint32_t)ya[y], b0 = 256u - b1;
    for (int x = 0; x < dw; ++x) {
      int ox = xo[x];
      int ox1 = std::min(ox + 1, sw - 1);
      uint32_t a1 = (uint32_t)xa[x], a0 = 256u - a1;
      h0[x] = a0 * r0[ox] + a1 * r0[ox1];
      h1[x] = a0 * r1[ox] + a1 * r1[ox1];
      uint32_t v = b0 * h0[x] + b1 * h1[x];
      dst[(size_t)y * dstride + x] = (uint8_t)((v + 32768u) >> 16);
    }
  }
}

// A2 — FAST-9/16 score of one pixel; returns S = max over 9-arcs of the min signed difference, both
// polarities (0 when no arc qualifies at all is not needed: caller thresholds S > t).
static inline int fast_arc_score(const uint8_t* p, int stride) {
  int d[16];
  int c = p[0];
  for (int k = 0; k < 16; ++k) d[k] = c - (int)p[kCircle[k][1] * stride + kCircle[k][0]];
  int best = -255;
  for (int k = 0; k < 16; ++k) {
    int mn = 255, mx = -255;
    for (int j = 0; j < 9; ++j) {
      int v = d[(k + j) & 15];
      mn = std::min(mn, v);
      mx = std::max(mx, v);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

// S of one pixel (>= 3 px inside the image), for the classic-mode model (oracle/classic_model.cpp).
int orc_fast_score_at(const uint8_t* img, int stride, int x, int y) { return fast_arc_score(img + (size_t)y * stride + x, stride); }

// FAST(threshold, nonmaxSuppression=true, TYPE_9_16): writes corners in raster order.
// score map holds (S-1) for corners, 0 otherwise. Returns the number of corners (may exceed cap; only cap written).
int orc_fast_nms(const uint8_t* img, int w, int h, int stride, int threshold, OrcCorner* out, int cap) {
  std::vector<uint8_t> score((size_t)w * h, 0);
  for (int y = 3; y < h - 3; ++y) {
    const uint8_t* row = img + (size_t)y * stride;
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = row + x;
      int c = p[0];
      // cheap necessary condition: every 9-arc contains circle point 0 or 8
      int d0 = c - (int)p[3 * stride], d8 = c - (int)p[-3 * stride];
      if (std::abs(d0) <= threshold && std::abs(d8) <= threshold) continue;
      int s = fast_arc_score(p, stride);
      if (s > threshold) score[(size_t)y * w + x] = (uint8_t)(s - 1);
    }
  }
  int n = 0;
  for (int y = 3; y < h - 3; ++y) {
    for (int x = 3; x < w - 3; ++x) {
      int s = score[(size_t)y * w + x];
      if (!s) continue;
      const uint8_t* sp = &score[(size_t)y * w + x];
      if (s > sp[-1] && s > sp[1] && s > sp[-w - 1] && s > sp[-w] && s > sp[-w + 1] && s > sp[w - 1] &&
          s > sp[w] && s > sp[w + 1]) {
        if (n < cap) {
          out[n].x = x;
          out[n].y = y;
          out[n].resp = (float)s;
        }
        ++n;
      }
    }
  }
  return n;
}

// A3(i) — KeyPointsFilter::runByImageBorder: stable filter to [border, dim-border).
int orc_border_filter(OrcCorner* v, int n, int w, int h, int border) {
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (v[i].x >= border && v[i].x < w - border && v[i].y >= border && v[i].y < h - border) v[m++] = v[i];
  }
  return m;
}

// A3(ii,iv) — KeyPointsFilter::retainBest as built against libstdc++ (std::nth_element + std::partition).
int orc_retain_best(OrcCorner* v, int n, int n_points) {
  if (n_points >= 0 && n > n_points) {
    if (n_points == 0) return 0;
    auto greater = [](const OrcCorner& a, const OrcCorner& b) { return a.resp > b.resp; };
    std::nth_element(v, v + n_points - 1, v + n, greater);
    float thr = v[n_points - 1].resp;
    OrcCorner* e = std::partition(v + n_points, v + n, [thr](const OrcCorner& a) { return a.resp >= thr; });
    return (int)(e - v);
  }
  return n;
}

// A3(iii) — HarrisResponses, blockSize 7, k = 0.04, on the un-blurred level.
void orc_harris(const uint8_t* img, int stride, OrcCorner* v, int n) {
  const int block = 7;
  const float harris_k = 0.04f;
  float scale = 1.f / ((1 << 2) * block * 255.f);
  float scale_sq_sq = scale * scale * scale * scale;
  int r = block / 2;
  for (int i = 0; i < n; ++i) {
    const uint8_t* p0 = img + (size_t)(v[i].y - r) * stride + (v[i].x - r);
    int a = 0, b = 0, c = 0;
    for (int k = 0; k < block * block; ++k) {
      const uint8_t* p = p0 + (k / block) * stride + (k % block);
      int Ix = ((int)p[1] - p[-1]) * 2 + ((int)p[-stride + 1] - p[-stride - 1]) + ((int)p[stride + 1] - p[stride - 1]);
      int Iy = ((int)p[stride] - p[-stride]) * 2 + ((int)p[stride - 1] - p[-stride - 1]) +
               ((int)p[stride + 1] - p[-stride + 1]);
      a += Ix * Ix;
      b += Iy * Iy;
      c += Ix * Iy;
    }
    float fa = (float)a, fb = (float)b, fc = (float)c;
    float t = fa + fb;
    v[i].resp = ((fa * fb - fc * fc) - (harris_k * t) * t) * scale_sq_sq;
  }
}

// cv::fastAtan2 scalar float path (degrees).
float orc_fast_atan2(float y, float x) {
  const float rad2deg = (float)(180.0 / 3.141592653589793238462643383279502884);
  const float p1 = 0.9997878412794807f * rad2deg;
  const float p3 = -0.3258083974640975f * rad2deg;
  const float p5 = 0.1555786518463281f * rad2deg;
  const float p7 = -0.04432655554792128f * rad2deg;
  const float eps = (float)2.2204460492503131e-16;
  float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + eps);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + eps);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// A4 — ICAngles: intensity-centroid orientation over the 31-px disc.
float orc_ic_angle(const uint8_t* img, int stride, int x, int y) {
  const uint8_t* center = img + (size_t)y * stride + x;
  int m01 = 0, m10 = 0;
  for (int u = -15; u <= 15; ++u) m10 += u * center[u];
  for (int v = 1; v <= 15; ++v) {
    int v_sum = 0, d = kUmax[v];
    for (int u = -d; u <= d; ++u) {
      int vp = center[u + v * stride], vm = center[u - v * stride];
      v_sum += vp - vm;
      m10 += u * (vp + vm);
    }
    m01 += v * v_sum;
  }
  return orc_fast_atan2((float)m01, (float)m10);
}

// A5 — the blur cv::ORB applies to each (non-isolated ROI) level: float32 separable 7-tap sigma=2 filter,
// FMA accumulation in OpenCV's AVX2 row/column filters, BORDER_REFLECT_101, round-half-even to u8.
void orc_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
  static const float k[7] = {0x1.1f5f62p-4f, 0x1.0c70fcp-3f, 0x1.869472p-3f, 0x1.ba95c0p-3f,
                             0x1.869472p-3f, 0x1.0c70fcp-3f, 0x1.1f5f62p-4f};
  auto refl = [](int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
  };
  std::vector<float> rows((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* r = src + (size_t)y * sstride;
    for (int x = 0; x < w; ++x) {
      float s = k[0] * (float)r[refl(x - 3, w)];
      for (int j = 1; j < 7; ++j) s = std::fmaf((float)r[refl(x - 3 + j, w)], k[j], s);
      rows[(size_t)y * w + x] = s;
    }
  }
  for (int y = 0; y < h; ++y) {
    for (int x = 0; x < w; ++x) {
      float s = k[3] * rows[(size_t)y * w + x];
      for (int j = 1; j <= 3; ++j) {
        float t = rows[(size_t)refl(y + j, h) * w + x] + rows[(size_t)refl(y - j, h) * w + x];
        s = std::fmaf(t, k[3 + j], s);
      }
      int iv = cv_roundf(s);
      dst[(size_t)y * dstride + x] = (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
    }
  }
}

// A6 — steered BRIEF, WTA_K = 2. (x,y) = integer level coordinate of the keypoint, angle in degrees.
void orc_brief(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc) {
  float angle = angle_deg;
  angle *= (float)(3.141592653589793238462643383279502884 / 180.f);
  float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
  const uint8_t* center = blurred + (size_t)y * stride + x;
  auto value = [&](int idx) -> int {
    float px = (float)kBriefPattern[2 * idx], py = (float)kBriefPattern[2 * idx + 1];
    float fx = px * a - py * b;
    float fy = px * b + py * a;
    int ix = cv_roundf(fx), iy = cv_roundf(fy);
    return center[iy * stride + ix];
  };
  for (int i = 0; i < 32; ++i) {
    int byte = 0;
    for (int bit = 0; bit < 8; ++bit) {
      int t0 = value(2 * (8 * i + bit)), t1 = value(2 * (8 * i + bit) + 1);
      byte |= (t0 < t1) << bit;
    }
    desc[i] = (uint8_t)byte;
  }
}

// Whole cv::ORB::detectAndCompute as the reference configures it (stereo.rs:38-48,72-73).
// Returns the number of keypoints; writes at most cap. nlevels <= 16.
int orc_detect_and_compute(const uint8_t* img, int W, int H, int stride, int nfeatures, int nlevels,
                           float scale_factor, int edge_threshold, int fast_threshold, OrcKeyPoint* kps,
                           uint8_t* desc, int cap) {
  std::vector<OrcLevel> lv(nlevels);
  orc_level_geometry(W, H, nfeatures, nlevels, scale_factor, lv.data());
  std::vector<std::vector<uint8_t>> pyr(nlevels), blr(nlevels);
  pyr[0].resize((size_t)W * H);
  for (int y = 0; y < H; ++y) memcpy(&pyr[0][(size_t)y * W], img + (size_t)y * stride, W);
  for (int l = 1; l < nlevels; ++l) {
    pyr[l].resize((size_t)lv[l].w * lv[l].h);
    orc_resize_linear_exact(pyr[l - 1].data(), lv[l - 1].w, lv[l - 1].h, lv[l - 1].w, pyr[l].data(), lv[l].w, lv[l].h,
                            lv[l].w);
  }
  int n = 0;
  for (int l = 0; l < nlevels; ++l) {
    int w = lv[l].w, h = lv[l].h;
    std::vector<OrcCorner> c((size_t)w * h / 4 + 16);
    int m = orc_fast_nms(pyr[l].data(), w, h, w, fast_threshold, c.data(), (int)c.size());
    m = orc_border_filter(c.data(), m, w, h, edge_threshold);
    m = orc_retain_best(c.data(), m, 2 * lv[l].nfeatures);
    orc_harris(pyr[l].data(), w, c.data(), m);
    m = orc_retain_best(c.data(), m, lv[l].nfeatures);
    blr[l].resize((size_t)w * h);
    orc_blur7(pyr[l].data(), w, h, w, blr[l].data(), w);
    for (int i = 0; i < m; ++i, ++n) {
      if (n >= cap) continue;
      float ang = orc_ic_angle(pyr[l].data(), w, c[i].x, c[i].y);
      OrcKeyPoint& k = kps[n];
      k.x = (float)c[i].x * lv[l].scale;
      k.y = (float)c[i].y * lv[l].scale;
      k.size = 31.f * lv[l].scale;
      k.angle = ang;
      k.response = c[i].resp;
      k.octave = l;
      k.class_id = -1;
      // descriptor centre = cvRound(pt * (1/scale)), which recovers the integer level coordinate
      float inv = 1.f / lv[l].scale;
      orc_brief(blr[l].data(), w, cv_roundf(k.x * inv), cv_roundf(k.y * inv), ang, desc + (size_t)n * 32);
    }
  }
  return n;
}

}  // extern "C"

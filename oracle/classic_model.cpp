// oracle/classic_model.cpp — TEST INFRASTRUCTURE ONLY: the CPU model of the `classic` mode.
//
// NOT REFERENCE PARITY.  jurmy24/orb-slam3-rust extracts with stock cv::ORB and matches with a gated brute-force loop
// (SURVEY.md 0.1); it contains no per-cell FAST, no quadtree, no SAD refinement and no parabola.  `north_star` asks
// for those upstream-ORB-SLAM3 stages, so this file defines them — following the PUBLISHED algorithm of ORB-SLAM3's
// ORBextractor::ComputeKeyPointsOctTree / DistributeOctTree and Frame::ComputeStereoMatches (Campos et al., T-RO 2021;
// Mur-Artal & Tardos, T-RO 2017) — and the GPU `classic` mode is tested against THIS model, nothing else.
// Where upstream leaves the result to implementation accidents (std::list splice order, a sort that breaks ties
// by node POINTER value) the model picks a deterministic rule and says so.  Deviations, all deliberate:
//   * pyramid = the same INTER_LINEAR_EXACT chain as the parity mode (upstream: INTER_LINEAR + a reflected border);
//   * level scales = f32 pow(1.2f, l) as cv::ORB (upstream multiplies iteratively);
//   * quadtree: nodes are processed in creation order, the size-ordered phase breaks ties by creation index
//     (later first), the survivors are emitted in candidate order (cell row-major, raster inside a cell);
//   * parabola with a zero denominator is rejected (upstream would carry a NaN).
// Stages shared with the parity mode (IC angle, 7x7 blur, steered BRIEF, FAST arc score) are the functions of
// orb_oracle.cpp.  Build: oracle/Makefile (-ffp-contract=off).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {
struct OrcLevel {
  int w, h;
  float scale;
  int nfeatures;
};
struct OrcKeyPoint {
  float x, y, size, angle, response;
  int octave, class_id;
};
void orc_level_geometry(int W, int H, int nfeatures, int nlevels, float scale_factor, OrcLevel* out);
void orc_resize_linear_exact(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
float orc_ic_angle(const uint8_t* img, int stride, int x, int y);
void orc_blur7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
void orc_brief(const uint8_t* blurred, int stride, int x, int y, float angle_deg, uint8_t* desc);
int orc_fast_score_at(const uint8_t* img, int stride, int x, int y);
uint32_t orc_descriptor_distance(const uint8_t* a, int cols_a, const uint8_t* b, int cols_b);
}

namespace {

constexpr int kCell = 35;       // ORB-SLAM3: const float W = 35
constexpr int kEdge = 19;       // EDGE_THRESHOLD
constexpr int kIniTh = 20, kMinTh = 7;

struct Cand {
  int x, y;   // level coordinates (absolute)
  int resp;   // cv::FAST response = S - 1
};

// Per-cell FAST with threshold retry (ComputeKeyPointsOctTree): candidates in cell row-major / raster order.
void classic_cells(const uint8_t* img, int w, int h, int stride, int ini_th, int min_th, std::vector<Cand>& out) {
  const int minBX = kEdge - 3, minBY = kEdge - 3, maxBX = w - kEdge + 3, maxBY = h - kEdge + 3;
  const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
  const int nCols = std::max((int)(width / kCell), 1), nRows = std::max((int)(height / kCell), 1);
  const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
  std::vector<int> S;  // S - 1 of the cell's interior, 0 = not a corner at min_th
  for (int i = 0; i < nRows; ++i) {
    const int iniY = minBY + i * hCell;
    int maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = maxBY;
    for (int j = 0; j < nCols; ++j) {
      const int iniX = minBX + j * wCell;
      int maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = maxBX;
      // cv::FAST on the cell image: corners live 3 px inside it; scores outside that interior read as 0 in the NMS
      const int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;
      const int cw = x1 - x0, ch = y1 - y0;
      if (cw <= 0 || ch <= 0) continue;
      S.assign((size_t)(cw + 2) * (ch + 2), 0);
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const int s = orc_fast_score_at(img, stride, x, y);
          if (s > min_th) S[(size_t)(y - y0 + 1) * (cw + 2) + (x - x0 + 1)] = s - 1;
        }
      for (int pass = 0; pass < 2; ++pass) {
        const int th = pass == 0 ? ini_th : min_th;
        const size_t before = out.size();
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) {
            const int* p = &S[(size_t)(y - y0 + 1) * (cw + 2) + (x - x0 + 1)];
            const int s = *p, W2 = cw + 2;
            if (s + 1 <= th) continue;  // S > th
            if (s > p[-1] && s > p[1] && s > p[-W2 - 1] && s > p[-W2] && s > p[-W2 + 1] && s > p[W2 - 1] && s > p[W2] &&
                s > p[W2 + 1])
              out.push_back({x, y, s});
          }
        if (out.size() > before) break;  // the cell has corners at iniThFAST: no retry
      }
    }
  }
}

struct Node {
  int ulx, uly, brx, bry;  // [ulx, brx) x [uly, bry), coordinates relative to (minBX, minBY)
  std::vector<int> keys;   // candidate indices, ascending
  bool alive = true;
};

void divide(const Node& n, const std::vector<Cand>& c, int offx, int offy, Node out[4]) {
  const int halfX = (int)std::ceil((float)(n.brx - n.ulx) / 2), halfY = (int)std::ceil((float)(n.bry - n.uly) / 2);
  const int mx = n.ulx + halfX, my = n.uly + halfY;
  out[0] = {n.ulx, n.uly, mx, my, {}, true};
  out[1] = {mx, n.uly, n.brx, my, {}, true};
  out[2] = {n.ulx, my, mx, n.bry, {}, true};
  out[3] = {mx, my, n.brx, n.bry, {}, true};
  for (int k : n.keys) {
    const int x = c[k].x - offx, y = c[k].y - offy;
    out[(x < mx ? 0 : 1) + (y < my ? 0 : 2)].keys.push_back(k);
  }
}

// DistributeOctTree, deterministic variant (see the header).  Returns the surviving candidate indices, ascending.
void classic_quadtree(const std::vector<Cand>& c, int w, int h, int N, std::vector<int>& keep) {
  keep.clear();
  if (c.empty() || N <= 0) return;
  const int minBX = kEdge - 3, minBY = kEdge - 3, maxBX = w - kEdge + 3, maxBY = h - kEdge + 3;
  const int W = maxBX - minBX, H = maxBY - minBY;
  const int nIni = std::max((int)std::round((float)W / (float)H), 1);
  const float hX = (float)W / nIni;
  std::vector<Node> nodes;  // index = creation index
  for (int i = 0; i < nIni; ++i) nodes.push_back({(int)(hX * (float)i), 0, (int)(hX * (float)(i + 1)), H, {}, true});
  for (int k = 0; k < (int)c.size(); ++k) {
    const int i = std::min((int)((float)(c[k].x - minBX) / hX), nIni - 1);
    nodes[i].keys.push_back(k);
  }
  int live = 0;
  for (auto& n : nodes) {
    if (n.keys.empty()) n.alive = false;
    live += n.alive;
  }
  auto split = [&](int id) {  // replaces node id by its non-empty children (creation order n1..n4); returns new live count
    Node ch[4];
    divide(nodes[id], c, minBX, minBY, ch);
    nodes[id].alive = false;
    --live;
    for (int q = 0; q < 4; ++q)
      if (!ch[q].keys.empty()) {
        nodes.push_back(ch[q]);
        ++live;
      }
  };
  bool finish = false;
  while (!finish) {
    const int prev = live, first_new = (int)nodes.size();
    std::vector<int> todo;
    for (int id = 0; id < first_new; ++id)
      if (nodes[id].alive && nodes[id].keys.size() > 1) todo.push_back(id);
    for (int id : todo) split(id);
    int expandable = 0;
    for (int id = first_new; id < (int)nodes.size(); ++id) expandable += nodes[id].keys.size() > 1;
    if (live >= N || live == prev) {
      finish = true;
    } else if (live + expandable * 3 > N) {
      while (!finish) {  // size-ordered phase: largest nodes first, stop the moment N nodes exist
        const int prev2 = live;
        std::vector<int> order;
        for (int id = 0; id < (int)nodes.size(); ++id)
          if (nodes[id].alive && nodes[id].keys.size() > 1) order.push_back(id);
        std::sort(order.begin(), order.end(), [&](int a, int b) {
          if (nodes[a].keys.size() != nodes[b].keys.size()) return nodes[a].keys.size() > nodes[b].keys.size();
          return a > b;
        });
        for (int id : order) {
          split(id);
          if (live >= N) break;
        }
        if (live >= N || live == prev2) finish = true;
      }
    }
  }
  for (const Node& n : nodes) {
    if (!n.alive) continue;
    int best = n.keys[0];
    for (int k : n.keys)
      if (c[k].resp > c[best].resp) best = k;  // first maximum in candidate order
    keep.push_back(best);
  }
  std::sort(keep.begin(), keep.end());
}

}  // namespace

extern "C" {

// Extraction of the classic mode.  levels_out (optional): per-level u8 images are not returned; keypoints carry octave.
// Returns the keypoint count (writes at most cap).
int orc_classic_detect_and_compute(const uint8_t* img, int W, int H, int stride, int nfeatures, int nlevels, float scale_factor,
                                   int ini_th, int min_th, OrcKeyPoint* kps, uint8_t* desc, int cap) {
  std::vector<OrcLevel> lv(nlevels);
  orc_level_geometry(W, H, nfeatures, nlevels, scale_factor, lv.data());
  std::vector<std::vector<uint8_t>> pyr(nlevels);
  pyr[0].resize((size_t)W * H);
  for (int y = 0; y < H; ++y) memcpy(&pyr[0][(size_t)y * W], img + (size_t)y * stride, W);
  for (int l = 1; l < nlevels; ++l) {
    pyr[l].resize((size_t)lv[l].w * lv[l].h);
    orc_resize_linear_exact(pyr[l - 1].data(), lv[l - 1].w, lv[l - 1].h, lv[l - 1].w, pyr[l].data(), lv[l].w, lv[l].h, lv[l].w);
  }
  int n = 0;
  std::vector<Cand> cand;
  std::vector<int> keep;
  std::vector<uint8_t> blr;
  for (int l = 0; l < nlevels; ++l) {
    const int w = lv[l].w, h = lv[l].h;
    cand.clear();
    classic_cells(pyr[l].data(), w, h, w, ini_th, min_th, cand);
    classic_quadtree(cand, w, h, lv[l].nfeatures, keep);
    blr.resize((size_t)w * h);
    orc_blur7(pyr[l].data(), w, h, w, blr.data(), w);
    for (int k : keep) {
      if (n < cap) {
        const float ang = orc_ic_angle(pyr[l].data(), w, cand[k].x, cand[k].y);
        OrcKeyPoint& o = kps[n];
        o.x = (float)cand[k].x * lv[l].scale;
        o.y = (float)cand[k].y * lv[l].scale;
        o.size = 31.f * lv[l].scale;
        o.angle = ang;
        o.response = (float)cand[k].resp;
        o.octave = l;
        o.class_id = -1;
        orc_brief(blr.data(), w, cand[k].x, cand[k].y, ang, desc + (size_t)n * 32);
      }
      ++n;
    }
  }
  return n;
}

// Debug view of one level: candidates after the per-cell FAST (x, y, resp) and the quadtree survivors' indices.
int orc_classic_level_candidates(const uint8_t* img, int w, int h, int stride, int nfeat, int ini_th, int min_th, int* xyr,
                                 int cap, int* keep_idx, int* n_keep) {
  std::vector<Cand> cand;
  std::vector<int> keep;
  classic_cells(img, w, h, stride, ini_th, min_th, cand);
  classic_quadtree(cand, w, h, nfeat, keep);
  for (int i = 0; i < (int)cand.size() && i < cap; ++i) {
    xyr[3 * i] = cand[i].x;
    xyr[3 * i + 1] = cand[i].y;
    xyr[3 * i + 2] = cand[i].resp;
  }
  for (int i = 0; i < (int)keep.size() && i < cap; ++i) keep_idx[i] = keep[i];
  *n_keep = (int)keep.size();
  return (int)cand.size();
}

// ComputeStereoMatches of the classic mode.  pyr_l / pyr_r: the 8 level images of both eyes (tight, w_l * h_l bytes each,
// concatenated); out per LEFT keypoint: match_idx (-1), orb_dist, u_right (-1.f), depth (-1.f).  Returns the number of
// left keypoints that keep a depth.
int orc_classic_stereo(const OrcKeyPoint* kl, const uint8_t* dl, int nl, const OrcKeyPoint* kr, const uint8_t* dr, int nr,
                       const uint8_t* pyr_l, const uint8_t* pyr_r, const OrcLevel* lv, int nlevels, double fx, double baseline,
                       int* match_idx, int* orb_dist, float* u_right, float* depth) {
  const int TH_HIGH = 100, TH_LOW = 50, thOrbDist = (TH_HIGH + TH_LOW) / 2;
  const int nRows = lv[0].h;
  std::vector<size_t> off(nlevels + 1, 0);
  for (int l = 0; l < nlevels; ++l) off[l + 1] = off[l] + (size_t)lv[l].w * lv[l].h;
  std::vector<std::vector<int>> rows(nRows);
  for (int iR = 0; iR < nr; ++iR) {
    const float kpY = kr[iR].y, r = 2.0f * lv[kr[iR].octave].scale;
    const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
    for (int yi = std::max(minr, 0); yi <= std::min(maxr, nRows - 1); ++yi) rows[yi].push_back(iR);
  }
  const float mbf = (float)(fx * baseline), mb = (float)baseline;
  const float minD = 0.f, maxD = mbf / mb;
  std::vector<std::pair<int, int>> distIdx;
  for (int i = 0; i < nl; ++i) {
    match_idx[i] = -1;
    orb_dist[i] = -1;
    u_right[i] = -1.f;
    depth[i] = -1.f;
  }
  for (int iL = 0; iL < nl; ++iL) {
    const int levelL = kl[iL].octave;
    const float vL = kl[iL].y, uL = kl[iL].x;
    const int row = (int)vL;
    if (row < 0 || row >= nRows) continue;
    const std::vector<int>& cands = rows[row];
    if (cands.empty()) continue;
    const float minU = uL - maxD, maxU = uL - minD;
    if (maxU < 0) continue;
    int bestDist = TH_HIGH, bestIdxR = 0;
    for (int iR : cands) {
      if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
      const float uR = kr[iR].x;
      if (uR >= minU && uR <= maxU) {
        const int d = (int)orc_descriptor_distance(dl + (size_t)iL * 32, 32, dr + (size_t)iR * 32, 32);
        if (d < bestDist) {
          bestDist = d;
          bestIdxR = iR;
        }
      }
    }
    if (bestDist >= thOrbDist) continue;
    // sub-pixel refinement by 11x11 SAD over 11 shifts on the keypoint's pyramid level
    const float uR0 = kr[bestIdxR].x, inv = 1.0f / lv[levelL].scale;
    const int su = (int)std::round(uL * inv), sv = (int)std::round(vL * inv), sr0 = (int)std::round(uR0 * inv);
    const int w = 5, L = 5, lw = lv[levelL].w, lh = lv[levelL].h;
    if (sv - w < 0 || sv + w >= lh || su - w < 0 || su + w >= lw) continue;  // (never: keypoints are >= 19 px inside)
    const int iniu = sr0 + L - w, endu = sr0 + L + w + 1;
    if (iniu < 0 || endu >= lw) continue;
    if (sr0 - L - w < 0) continue;  // upstream reads colRange(sr0 - L - w, ...) unchecked; a negative start would throw
    const uint8_t* IL = pyr_l + off[levelL];
    const uint8_t* IR = pyr_r + off[levelL];
    int best = INT_MAX, bestinc = 0;
    float vd[2 * 5 + 1];
    for (int inc = -L; inc <= L; ++inc) {
      int sad = 0;
      for (int dy = -w; dy <= w; ++dy)
        for (int dx = -w; dx <= w; ++dx)
          sad += std::abs((int)IL[(size_t)(sv + dy) * lw + su + dx] - (int)IR[(size_t)(sv + dy) * lw + sr0 + inc + dx]);
      if (sad < best) {
        best = sad;
        bestinc = inc;
      }
      vd[L + inc] = (float)sad;
    }
    if (bestinc == -L || bestinc == L) continue;
    const float d1 = vd[L + bestinc - 1], d2 = vd[L + bestinc], d3 = vd[L + bestinc + 1];
    const float den = 2.0f * (d1 + d3 - 2.0f * d2);
    if (den == 0.0f) continue;  // model rule (upstream would divide by zero)
    const float deltaR = (d1 - d3) / den;
    if (deltaR < -1 || deltaR > 1) continue;
    float bestuR = lv[levelL].scale * ((float)sr0 + (float)bestinc + deltaR);
    float disparity = uL - bestuR;
    if (disparity >= minD && disparity < maxD) {
      if (disparity <= 0) {
        disparity = 0.01f;
        bestuR = uL - 0.01f;
      }
      match_idx[iL] = bestIdxR;
      orb_dist[iL] = bestDist;
      depth[iL] = mbf / disparity;
      u_right[iL] = bestuR;
      distIdx.push_back({best, iL});
    }
  }
  std::sort(distIdx.begin(), distIdx.end());
  int kept = (int)distIdx.size();
  if (!distIdx.empty()) {
    const float median = (float)distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)distIdx.size() - 1; i >= 0; --i) {
      if ((float)distIdx[i].first < thDist) break;
      const int iL = distIdx[i].second;
      match_idx[iL] = -1;
      orb_dist[iL] = -1;
      u_right[iL] = -1.f;
      depth[iL] = -1.f;
      --kept;
    }
  }
  return kept;
}

}  // extern "C"

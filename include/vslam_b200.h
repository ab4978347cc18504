/*
 * vslam_b200.h — C ABI of the B200-native stereo ORB front end (libvslam_b200.so).
 *
 * Drop-in boundary for ONE path of jurmy24/orb-slam3-rust (crate `rust-vslam`): the per-frame stereo front end.
 * Every entry point below replaces a piece of the reference's Rust API; the citations are file:line under the
 * reference repository root.  A Rust `StereoProcessor` with the reference's exact signature binds these through
 * `extern "C"` (see INTEGRATION.md and orb-slam3-rust_b200/rust/).
 *
 *   vslam_create / vslam_destroy        <- StereoProcessor::new              src/tracking/frame/stereo.rs:37-50
 *   vslam_process                       <- StereoProcessor::process          src/tracking/frame/stereo.rs:52-66
 *   vslam_process_batch[_device]        <- the caller loop around process()  src/main.rs:66-85 (offline batch form)
 *   vslam_detect_and_compute            <- StereoProcessor::detect_features  src/tracking/frame/stereo.rs:68-78
 *   vslam_match_features                <- StereoProcessor::match_features   src/tracking/frame/stereo.rs:80-161
 *   vslam_triangulate                   <- triangulate                       src/tracking/frame/stereo.rs:186-216
 *   vslam_descriptor_distance           <- descriptor_distance               src/tracking/frame/stereo.rs:166-175
 *   vslam_grid_build / vslam_grid_query <- FeatureGrid::{new,get_features_in_area}  src/tracking/tracking_frame.rs:52-128
 *   vslam_search_by_projection          <- the two inline searches in Tracker       src/tracking/tracker.rs:863-923, 1093-1158
 *   vslam_bf_match_crosscheck           <- BFMatcher(NORM_HAMMING,true).train_match src/tracking/tracker.rs:1001-1010
 *   vslam_fuse_search                   <- descriptor search of fuse_points_into_keyframes
 *                                          src/local_mapping/search_in_neighbors.rs:303-340 (+ keyframe.rs:408-446)
 *   vslam_distinctive_descriptors       <- Map::compute_distinctive_descriptors       src/atlas/map/map.rs:917-937
 *   vslam_vocab_upload / vslam_bow_quantize <- OrbVocabulary::{load_from_text, transform_descriptor,
 *                                          get_parent_at_level}                       src/vocabulary/mod.rs:109-206, 230-275
 *   vslam_search_for_triangulation[_bow] <- search_for_triangulation[_bow]
 *                                          src/local_mapping/triangulation.rs:401-530, 541-658 (+ :339-392, 661-699)
 *
 * All pixel/descriptor arithmetic runs in hand-written CUDA kernels for sm_100a.  There is NO CPU fallback:
 * every call fails with VSLAM_ERR_CUDA when no usable device is present.
 *
 * Threading: a handle is single-caller (the reference's `process(&mut self)`); use one handle per host thread /
 * per GPU.  All outputs are plain host memory owned by the caller (no device affinity), so they can be moved into
 * `StereoFrame` / `KeyFrame` and read from other threads exactly like the reference's values.
 *
 * Results are bit-identical to the reference's CPU path (cv::ORB as configured at stereo.rs:38-48 + the Rust loops):
 * keypoint order, coordinates, octaves, angles, responses, descriptors, match pairs, f64 points.  Caveats:
 *   1. Keypoint capacity.  cv::ORB keeps every keypoint whose response ties with the n-th best, so a frame can hold more
 *      than n_features keypoints.  A frame that exceeds cfg.keypoint_capacity is NOT truncated: the call returns
 *      VSLAM_ERR_KEYPOINT_OVERFLOW (frame status set) where the reference would return the frame.  Re-create the handle
 *      with a larger keypoint_capacity to process such input (white noise / synthetic plateaus; never seen on textured
 *      scenes at the default capacity n_features + max(n_features / 8, 64)).
 *   2. Host arithmetic the reference inherits from its platform.  The blur follows OpenCV's sepFilter2D on an x86 host
 *      with FMA (AVX2 path: fused multiply-adds in the order documented in SURVEY.md 8.A5), and BRIEF steering uses
 *      (float)cos/sin of the host libm; the device evaluates both in double precision and rounds once, which equals
 *      glibc's result except when a rotated sample coordinate lands within ~1e-7 of x.5 (probability ~1e-8 per
 *      descriptor bit pair; none observed).  Parity is therefore pinned to cv2 4.x on FMA-capable x86, the platform
 *      the parity tests run on.
 *   3. Every entry point makes cfg.device current for the duration of the call and restores the caller's device.
 */
#ifndef VSLAM_B200_H_
#define VSLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSLAM_ABI_VERSION 1

/* status codes (0 = ok). vslam_last_error() holds the message of the last failing call on the handle. */
enum {
  VSLAM_OK = 0,
  VSLAM_ERR_INVALID_ARGUMENT = 1,
  VSLAM_ERR_CUDA = 2,              /* no device / driver / kernel failure: the path fails loudly, never falls back */
  VSLAM_ERR_CAPACITY = 3,          /* a caller-provided output buffer is too small */
  VSLAM_ERR_KEYPOINT_OVERFLOW = 4, /* more keypoints than cfg.keypoint_capacity (response ties at the cut) */
  VSLAM_ERR_OUT_OF_MEMORY = 5
};

/* Thresholds of the reference (stereo.rs:10-12). */
#define VSLAM_TH_HIGH 100u
#define VSLAM_TH_LOW 50u
#define VSLAM_NN_RATIO 0.75f

/* cv::KeyPoint memory layout (28 bytes) — what `Vector<KeyPoint>` holds in FeatureSet (stereo.rs:16-19). */
typedef struct vslam_keypoint {
  float x, y;     /* pt, level coordinate * level scale (f32) */
  float size;     /* 31 * level scale */
  float angle;    /* degrees, cv::fastAtan2 */
  float response; /* Harris response */
  int32_t octave; /* pyramid level */
  int32_t class_id; /* -1 */
} vslam_keypoint;

/* cv::DMatch memory layout (16 bytes) — `Vector<DMatch>` matches_lr (stereo.rs:25). */
typedef struct vslam_dmatch {
  int32_t query_idx; /* left keypoint index */
  int32_t train_idx; /* right keypoint index */
  int32_t img_idx;   /* 0 */
  float distance;    /* best Hamming distance */
} vslam_dmatch;

/* CameraModel (src/tracking/frame/camera.rs:3-10). */
typedef struct vslam_camera {
  double fx, fy, cx, cy, baseline;
} vslam_camera;

typedef struct vslam_config {
  int32_t abi_version;  /* VSLAM_ABI_VERSION */
  int32_t width, height; /* image size in pixels (8UC1) */
  int32_t n_features;   /* StereoProcessor::new(camera, n_features); the reference passes 1200 (main.rs:53) */
  int32_t n_levels;     /* 0 -> 8   (stereo.rs:41) */
  float scale_factor;   /* 0 -> 1.2f (stereo.rs:40) */
  int32_t edge_threshold; /* 0 -> 31 (stereo.rs:42) */
  int32_t fast_threshold; /* 0 -> 20 (stereo.rs:47) */
  vslam_camera camera;
  int32_t device;       /* CUDA device ordinal */
  int32_t max_batch;    /* stereo frames resident per pipeline stage; 0 -> 64 */
  int32_t keypoint_capacity; /* per image; 0 -> n_features + max(n_features/8, 64) */
  int32_t mode;             /* VSLAM_MODE_REFERENCE (0, default) or VSLAM_MODE_CLASSIC; was reserved[0] */
  int32_t classic_ini_fast; /* classic mode: iniThFAST, 0 -> 20 */
  int32_t classic_min_fast; /* classic mode: minThFAST, 0 -> 7 */
  int32_t reserved[5];
} vslam_config;

/* Extraction / stereo algorithm of a handle.
 *   VSLAM_MODE_REFERENCE  what jurmy24/orb-slam3-rust computes: stock cv::ORB (global FAST threshold, retainBest x2 with
 *                         Harris) + the gated brute-force matcher of stereo.rs:80-161.  Bit-exact with the reference.
 *   VSLAM_MODE_CLASSIC    the upstream-ORB-SLAM3 style front end `north_star` also names and the reference does NOT contain:
 *                         per-cell FAST (35-px cells, iniThFAST with a minThFAST retry), quadtree keypoint distribution,
 *                         row-band stereo matching with an octave-scaled radius and +-1 level consistency, 11x11 SAD
 *                         refinement over 11 shifts, parabolic sub-pixel disparity and the median SAD outlier filter.
 *                         NOT REFERENCE PARITY — there is no reference code for it; it is specified by this repository's
 *                         own CPU model (oracle/classic_model.cpp, published ORB-SLAM3 algorithm, deterministic
 *                         tie-breaks) and tested bit-for-bit against that model only.  Same outputs (keypoints,
 *                         descriptors, matches_lr, points_cam) + vslam_classic_stereo_extra for the refined u_right / depth.
 *                         The tracking-side entry points are mode-independent. */
enum { VSLAM_MODE_REFERENCE = 0, VSLAM_MODE_CLASSIC = 1 };

/* Caller-owned, capacity-checked result of one stereo frame == StereoFrame (stereo.rs:21-29). */
typedef struct vslam_frame_out {
  int32_t capacity;          /* elements available in each array below */
  vslam_keypoint* kp_left;   /* [capacity] */
  vslam_keypoint* kp_right;  /* [capacity] */
  uint8_t* desc_left;        /* [capacity][32] */
  uint8_t* desc_right;       /* [capacity][32] */
  vslam_dmatch* matches;     /* [capacity], ascending query_idx */
  double* points_cam;        /* [capacity][3]; valid where points_valid[i] != 0 (Option<Vector3<f64>>) */
  uint8_t* points_valid;     /* [capacity] */
  int32_t n_left, n_right, n_matches; /* written by the call */
  int32_t status;            /* per-frame status (VSLAM_OK / VSLAM_ERR_*) */
  uint64_t timestamp_ns;     /* copied through */
} vslam_frame_out;

typedef struct vslam_handle_s* vslam_handle;

/* ---- lifetime ------------------------------------------------------------------------------------------- */
int vslam_create(const vslam_config* cfg, vslam_handle* out);
void vslam_destroy(vslam_handle h);
/* Message of the last error on this handle (or of the last failed vslam_create when h == NULL). */
const char* vslam_last_error(vslam_handle h);
/* Per-image keypoint capacity the handle was created with (size caller buffers with this). */
int vslam_keypoint_capacity(vslam_handle h);

/* ---- StereoProcessor::process ------------------------------------------------------------------------------ */
/* One stereo pair from HOST memory (borrowed &Mat: 8UC1, row stride in bytes). */
int vslam_process(vslam_handle h, const uint8_t* left, size_t left_stride, const uint8_t* right, size_t right_stride,
                  uint64_t timestamp_ns, vslam_frame_out* out);

/* n independent stereo pairs from HOST memory; frames are pipelined (H2D / kernels / D2H overlap) in chunks of
 * cfg.max_batch.  left[i]/right[i] point to 8UC1 images with the given row stride.  Pinned host memory gives the
 * best transfer rate; pageable memory is staged. */
int vslam_process_batch(vslam_handle h, int n, const uint8_t* const* left, const uint8_t* const* right, size_t stride,
                        const uint64_t* timestamps_ns, vslam_frame_out* outs);

/* Same, with the images already resident in device memory (HBM) of cfg.device: left_dev/right_dev are DEVICE
 * pointers to n images each, `image_stride` bytes apart, rows `row_stride` bytes apart (row_stride % 16 == 0,
 * row_stride >= width, 16-byte aligned bases; a row_stride wider than align16(width) raises k_fast's shared-memory
 * opt-in on first use and fails with VSLAM_ERR_INVALID_ARGUMENT beyond the device limit).  Results still land in host
 * memory.  If outs == NULL the results stay on the device (throughput measurement of the kernels alone); the per-frame
 * status words of EVERY chunk are still read back, so a keypoint overflow anywhere in the batch is reported. */
int vslam_process_batch_device(vslam_handle h, int n, const uint8_t* left_dev, const uint8_t* right_dev,
                               size_t image_stride, size_t row_stride, const uint64_t* timestamps_ns,
                               vslam_frame_out* outs);

/* Classic mode only: the sub-pixel right coordinate (mvuRight) and depth (mvDepth) of every LEFT keypoint of frame
 * `index` of the most recent process* call, -1 where the keypoint has no stereo match.  u_right[capacity], depth[capacity]. */
int vslam_classic_stereo_extra(vslam_handle h, int index, float* u_right, float* depth, int capacity);
/* Classic mode only (tests): candidates of `image` / `level` after the per-cell FAST as (x, y, response) triples. */
int vslam_debug_classic_candidates(vslam_handle h, int image, int level, int32_t* xyr, int capacity, int* n_out);

/* ---- the pieces of process(), individually --------------------------------------------------------------- */
/* detect_features: ORB keypoints + descriptors of one image.  kps[capacity], desc[capacity][32]. */
int vslam_detect_and_compute(vslam_handle h, const uint8_t* image, size_t stride, vslam_keypoint* kps, uint8_t* desc,
                             int capacity, int* n_out);

/* match_features on caller-provided features.  matches[n_left]. */
int vslam_match_features(vslam_handle h, const vslam_keypoint* kp_left, const uint8_t* desc_left, int n_left,
                         const vslam_keypoint* kp_right, const uint8_t* desc_right, int n_right, vslam_dmatch* matches,
                         int* n_matches);

/* triangulate on caller-provided matches.  points[n_left][3], valid[n_left].  Entries with |disparity| < 0.5 are
 * skipped exactly like stereo.rs:204-207 (they never overwrite an earlier point of the same query_idx); among the
 * remaining entries of one query_idx the last wins, as in the reference's loop. */
int vslam_triangulate(vslam_handle h, const vslam_keypoint* kp_left, int n_left, const vslam_keypoint* kp_right,
                      int n_right, const vslam_dmatch* matches, int n_matches, double* points, uint8_t* valid);

/* descriptor_distance for `n` pairs of 32-byte descriptors: out[i] = popcount(a[i] ^ b[i]). */
int vslam_descriptor_distance(vslam_handle h, const uint8_t* a, const uint8_t* b, int n, uint32_t* out);

/* ---- tracking-side searches ---------------------------------------------------------------------------------- */
#define VSLAM_GRID_COLS 64
#define VSLAM_GRID_ROWS 48

/* FeatureGrid::new: CSR form of the 64x48 grid. cell_start[64*48+1], cell_items[n] (row-major cells, ascending
 * keypoint index inside a cell).  img_w/img_h: the reference hard-codes 752x480 (tracking_frame.rs:194-197). */
int vslam_grid_build(vslam_handle h, const vslam_keypoint* kps, int n, double img_w, double img_h, int32_t* cell_start,
                     int32_t* cell_items);

/* FeatureGrid::get_features_in_area for nq queries (x,y,r as f64 triples).  Results are concatenated in query
 * order: out_start[nq+1] (CSR), out_items[out_capacity].  Bug-compatible with the reference's casts. */
int vslam_grid_query(vslam_handle h, const vslam_keypoint* kps, int n, double img_w, double img_h, const double* xyr,
                     int nq, int32_t* out_start, int32_t* out_items, int out_capacity);

enum {
  VSLAM_SEARCH_LOCAL_MAP = 0,   /* tracker.rs:863-923: best<=th_high, ratio test when >1 candidate */
  VSLAM_SEARCH_MOTION_MODEL = 1 /* tracker.rs:1093-1158: 0<=u<bound_w, 0<=v<bound_h, d<th_high, no ratio */
};

typedef struct vslam_search_params {
  int32_t mode;       /* VSLAM_SEARCH_* */
  double radius;      /* 15.0 */
  uint32_t th_high;   /* 100 */
  float nn_ratio;     /* 0.75 */
  double img_w, img_h;     /* grid extent: 752, 480 */
  double bound_w, bound_h; /* motion-model bounds: 2*cx, 2*cy */
} vslam_search_params;

/* Projection search of M map points against one frame.  uv[M][2]: host-projected pixel coordinates (the nalgebra
 * projection stays in the caller); mp_desc[M][32].  out_idx[M] = matched keypoint index or -1,
 * out_dist[M] = best distance or -1. */
int vslam_search_by_projection(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, const double* uv,
                               const uint8_t* mp_desc, int M, const vslam_search_params* params, int32_t* out_idx,
                               int32_t* out_dist);

/* ---- frames resident in HBM (the `frame_id` form SURVEY.md 8 b proposes) -------------------------------------
 * The Tracker searches the frame it has just extracted (Frame::from_stereo, tracking_frame.rs:194-215, keeps the LEFT
 * features + FeatureGrid).  vslam_frame_retain keeps exactly that on the device: the left keypoints and descriptors of
 * frame `index` of the most recent vslam_process / vslam_process_batch* call are copied device-to-device out of the
 * pipeline's result arrays (no PCIe traffic) and the FeatureGrid is built there; img_w x img_h is the grid extent (the
 * reference hard-codes 752 x 480).  Must be called before the next process* call; of a multi-chunk batch only the last
 * chunks (4 x max_batch frames) are still resident.  vslam_frame_upload makes a resident frame from host features (a
 * keyframe that was not retained).  Searches then move only the map points' (u, v) + descriptors up and the matched
 * indices down.  vslam_frame_release recycles the id.  The vslam_frame_* searches are bit-identical to their
 * host-buffer forms below / above (same kernels). */
int vslam_frame_retain(vslam_handle h, int index, double img_w, double img_h, int* frame_id);
int vslam_frame_upload(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, double img_w, double img_h,
                       int* frame_id);
int vslam_frame_release(vslam_handle h, int frame_id);
int vslam_frame_num_features(vslam_handle h, int frame_id); /* -1 for an unknown id */
/* tracker.rs:863-923 / 1093-1158 against a resident frame; params->img_w/img_h must equal the frame's grid extent. */
int vslam_frame_search_by_projection(vslam_handle h, int frame_id, const double* uv, const uint8_t* mp_desc, int M,
                                     const vslam_search_params* params, int32_t* out_idx, int32_t* out_dist);
/* FeatureGrid::get_features_in_area of a resident frame (same output convention as vslam_grid_query). */
int vslam_frame_grid_query(vslam_handle h, int frame_id, const double* xyr, int nq, int32_t* out_start, int32_t* out_items,
                           int out_capacity);
/* tracker.rs:1001-1010: BFMatcher(NORM_HAMMING, true).train_match(query = keyframe descriptors, train = frame
 * descriptors) between two resident frames; matches[n_query]. */
int vslam_frame_bf_match_crosscheck(vslam_handle h, int query_frame_id, int train_frame_id, vslam_dmatch* matches,
                                    int* n_matches);
/* Device time (ms) of the k_search launch of the last vslam_frame_search_by_projection (profiling enabled only). */
float vslam_last_search_ms(vslam_handle h);

/* BFMatcher(NORM_HAMMING, crossCheck=true).train_match(query, train): matches[n_query], ascending query index. */
int vslam_bf_match_crosscheck(vslam_handle h, const uint8_t* query, int n_query, const uint8_t* train, int n_train,
                              vslam_dmatch* matches, int* n_matches);

/* ---- keyframe-rate Hamming consumers --------------------------------------------------- */
/* Fusion search of local mapping: for each query q = (u, v, radius) (host-projected map point, radius already clamped
 * by the caller) scan ALL n keyframe keypoints, keep those with du*du + dv*dv <= radius*radius (f64, as
 * KeyFrame::get_features_in_area), and return the first minimum Hamming distance below desc_threshold.
 * uvr[M][3] f64; mp_desc[M][32]; out_idx[M] = keypoint index or -1; out_dist[M] = distance or -1. */
int vslam_fuse_search(vslam_handle h, const vslam_keypoint* kps, const uint8_t* desc, int n, const double* uvr,
                      const uint8_t* mp_desc, int M, uint32_t desc_threshold, int32_t* out_idx, int32_t* out_dist);

/* Distinctive ("median") descriptor of M map points: descriptors of map point m are rows offsets[m]..offsets[m+1]-1 of
 * desc; out_idx[m] = position (within the group) of the descriptor whose maximum distance to the others is smallest
 * (first minimum), 0 for a single descriptor, -1 for none. */
int vslam_distinctive_descriptors(vslam_handle h, const uint8_t* desc, const int32_t* offsets, int M, int32_t* out_idx);

/* BoW quantisation against a vocabulary tree (the per-descriptor part of OrbVocabulary::transform, mod.rs:290-318).
 * vslam_vocab_upload keeps the flattened tree on the device: node 0 is the root; node_desc[n_nodes][32];
 * children of node i are child_ids[child_begin[i] .. child_begin[i+1]) in the order of VocabNode::children;
 * parent[i] (0xFFFFFFFF for the root, exactly VocabNode::parent); word_id[i] = Some(w) -> w, None -> -1.
 * vslam_bow_quantize: for each of the n descriptors (desc[n][32]) descend from the root to the closest child (first
 * minimum Hamming distance) until a node without children: out_word = word_id.unwrap_or(0), out_leaf = that node,
 * out_node = get_parent_at_level(leaf, levels_up).  The TF-IDF accumulation / L1 normalisation of BowVector stays on the
 * host: the reference sums f64 weights in HashMap iteration order, which no device order can reproduce. */
int vslam_vocab_upload(vslam_handle h, int n_nodes, const uint8_t* node_desc, const int32_t* child_begin,
                       const int32_t* child_ids, const uint32_t* parent, const int32_t* word_id);
int vslam_bow_quantize(vslam_handle h, const uint8_t* desc, int n, int levels_up, uint32_t* out_word,
                       uint32_t* out_leaf, uint32_t* out_node);

/* Epipolar-gated keyframe <-> keyframe matching for triangulation (local_mapping/triangulation.rs:401-530 and, with
 * FeatureVector groups, :541-658).  The loop-invariant pose algebra stays in the caller (nalgebra, like the projection of
 * the tracking searches): F = K^-T [t12]x R12 K^-1 exactly as check_epipolar_constraint builds it (:668-680, row-major
 * here) and the epipole (:421-428).  flags1[i]: bit 0 = feature i of keyframe 1 already has a map point (skipped),
 * bit 1 = it has stereo depth (no epipole-distance gate); flags2[j]: bit 0 = feature j of keyframe 2 has a map point.
 * out_pairs[][2] = (idx1, idx2) in the order the reference's greedy loop emits them (a feature of keyframe 2 is taken
 * at most once, first come first served); capacity n1 pairs (bow: g1_start[n_groups] pairs). */
typedef struct vslam_tri_params {
  double F[9];
  double epipole[2];
  uint32_t max_dist;                  /* config.max_descriptor_dist */
  float radius;                       /* 100.0f (triangulation.rs:440); grid variant only */
  uint32_t image_width, image_height; /* (camera.cx * 2.0) as u32, (camera.cy * 2.0) as u32 (:432-433); grid variant */
} vslam_tri_params;
int vslam_search_for_triangulation(vslam_handle h, const vslam_keypoint* kp1, const uint8_t* desc1, const uint8_t* flags1,
                                   int n1, const vslam_keypoint* kp2, const uint8_t* desc2, const uint8_t* flags2, int n2,
                                   const vslam_tri_params* params, int32_t* out_pairs, int* n_pairs);
/* FeatureVector variant: group g = the g-th vocabulary node the caller's HashMap iteration visited that is present in
 * both keyframes (the visiting order is an input: the greedy pass depends on it); its keyframe-1 features are
 * g1_items[g1_start[g] .. g1_start[g+1]) and its keyframe-2 features g2_items[g2_start[g] .. g2_start[g+1]). */
int vslam_search_for_triangulation_bow(vslam_handle h, const vslam_keypoint* kp1, const uint8_t* desc1,
                                       const uint8_t* flags1, int n1, const vslam_keypoint* kp2, const uint8_t* desc2,
                                       const uint8_t* flags2, int n2, const vslam_tri_params* params, int n_groups,
                                       const int32_t* g1_start, const int32_t* g1_items, const int32_t* g2_start,
                                       const int32_t* g2_items, int32_t* out_pairs, int* n_pairs);

/* ---- introspection for parity tests and profiling ---------------------------------------------------------- */
enum {
  VSLAM_STAGE_PYRAMID = 0,  /* u8 level image; levels >= 1: the outer edge_threshold - 21 px ring, which no stage
                             * reads, is not computed (whole levels with VSLAM_PYR_FULL=1) */
  VSLAM_STAGE_BLURRED = 1   /* u8 blurred level image; only the part BRIEF can sample is computed: pixels at least
                             * edge_threshold - 18 px from every border (whole levels with VSLAM_BLUR_FULL=1) */
};
/* Geometry of pyramid level `level`: width, height, scale, per-level feature quota. */
int vslam_level_info(vslam_handle h, int level, int* w, int* h_out, float* scale, int* n_features);
/* Copies a level image of image slot `image` (0 = left, 1 = right of the LAST processed frame / last frame of the
 * last batch chunk) into dst (tightly packed w*h bytes). */
int vslam_debug_level_image(vslam_handle h, int stage, int image, int level, uint8_t* dst, size_t dst_capacity);
/* FAST corners (post NMS + border filter, raster order) of `image`/`level` as triples (x, y, score) of int32. */
int vslam_debug_fast_corners(vslam_handle h, int image, int level, int32_t* xys, int capacity, int* n_out);
/* Number of kernels launched by this handle since creation (monotonic). */
uint64_t vslam_kernel_launches(vslam_handle h);
/* Device time (ms, CUDA events on the handle's stream) of the last vslam_process_batch* call, per stage:
 * [0]=pyramid [1]=fast [2]=select+harris+angle [3]=blur [4]=brief [5]=match+triangulate [6]=total kernels.
 * Only filled when profiling was enabled with vslam_set_profiling(h,1) (it serialises the pipeline). */
int vslam_set_profiling(vslam_handle h, int enabled);
/* Device time (ms) of the last vslam_process_batch* call: CUDA events recorded on the handle's own streams, from
 * before the first chunk's first operation to after the last operation of every stream (copies included). */
float vslam_last_batch_ms(vslam_handle h);
int vslam_last_stage_ms(vslam_handle h, float* ms7);

#ifdef __cplusplus
}
#endif
#endif /* VSLAM_B200_H_ */

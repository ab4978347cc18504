//! Raw `extern "C"` bindings of include/vslam_b200.h (ABI version 1).  Plain pointers and sizes only.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

pub const VSLAM_ABI_VERSION: i32 = 1;
pub const VSLAM_OK: c_int = 0;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vslam_keypoint { pub x: f32, pub y: f32, pub size: f32, pub angle: f32, pub response: f32, pub octave: i32, pub class_id: i32 }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vslam_dmatch { pub query_idx: i32, pub train_idx: i32, pub img_idx: i32, pub distance: f32 }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vslam_camera { pub fx: f64, pub fy: f64, pub cx: f64, pub cy: f64, pub baseline: f64 }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct vslam_config {
    pub abi_version: i32, pub width: i32, pub height: i32, pub n_features: i32, pub n_levels: i32,
    pub scale_factor: f32, pub edge_threshold: i32, pub fast_threshold: i32, pub camera: vslam_camera,
    pub device: i32, pub max_batch: i32, pub keypoint_capacity: i32, pub mode: i32, pub classic_ini_fast: i32,
    pub classic_min_fast: i32, pub reserved: [i32; 5],
}

#[repr(C)]
pub struct vslam_frame_out {
    pub capacity: i32,
    pub kp_left: *mut vslam_keypoint, pub kp_right: *mut vslam_keypoint,
    pub desc_left: *mut u8, pub desc_right: *mut u8,
    pub matches: *mut vslam_dmatch, pub points_cam: *mut f64, pub points_valid: *mut u8,
    pub n_left: i32, pub n_right: i32, pub n_matches: i32, pub status: i32, pub timestamp_ns: u64,
}

#[repr(C)]
pub struct vslam_search_params {
    pub mode: i32, pub radius: f64, pub th_high: u32, pub nn_ratio: f32,
    pub img_w: f64, pub img_h: f64, pub bound_w: f64, pub bound_h: f64,
}

/// local_mapping/triangulation.rs: F (row-major, built exactly as check_epipolar_constraint :668-680) and the epipole
/// (:421-428) are computed by the caller with nalgebra; see include/vslam_b200.h.
#[repr(C)]
pub struct vslam_tri_params {
    pub f: [f64; 9], pub epipole: [f64; 2], pub max_dist: u32, pub radius: f32, pub image_width: u32, pub image_height: u32,
}

pub const VSLAM_MODE_REFERENCE: i32 = 0;
/// upstream-ORB-SLAM3 style per-cell FAST / quadtree / SAD + parabola: NOT reference parity (include/vslam_b200.h)
pub const VSLAM_MODE_CLASSIC: i32 = 1;

#[repr(C)]
pub struct vslam_handle_s { _private: [u8; 0] }
pub type vslam_handle = *mut vslam_handle_s;

extern "C" {
    pub fn vslam_create(cfg: *const vslam_config, out: *mut vslam_handle) -> c_int;
    pub fn vslam_destroy(h: vslam_handle);
    pub fn vslam_last_error(h: vslam_handle) -> *const c_char;
    pub fn vslam_keypoint_capacity(h: vslam_handle) -> c_int;
    pub fn vslam_process(h: vslam_handle, left: *const u8, left_stride: usize, right: *const u8, right_stride: usize,
                         timestamp_ns: u64, out: *mut vslam_frame_out) -> c_int;
    pub fn vslam_process_batch(h: vslam_handle, n: c_int, left: *const *const u8, right: *const *const u8, stride: usize,
                               timestamps_ns: *const u64, outs: *mut vslam_frame_out) -> c_int;
    pub fn vslam_descriptor_distance(h: vslam_handle, a: *const u8, b: *const u8, n: c_int, out: *mut u32) -> c_int;
    pub fn vslam_grid_query(h: vslam_handle, kps: *const vslam_keypoint, n: c_int, img_w: f64, img_h: f64,
                            xyr: *const f64, nq: c_int, out_start: *mut i32, out_items: *mut i32, out_capacity: c_int) -> c_int;
    pub fn vslam_search_by_projection(h: vslam_handle, kps: *const vslam_keypoint, desc: *const u8, n: c_int,
                                      uv: *const f64, mp_desc: *const u8, m: c_int, params: *const vslam_search_params,
                                      out_idx: *mut i32, out_dist: *mut i32) -> c_int;
    pub fn vslam_bf_match_crosscheck(h: vslam_handle, query: *const u8, n_query: c_int, train: *const u8, n_train: c_int,
                                     matches: *mut vslam_dmatch, n_matches: *mut c_int) -> c_int;
    pub fn vslam_fuse_search(h: vslam_handle, kps: *const vslam_keypoint, desc: *const u8, n: c_int, uvr: *const f64,
                             mp_desc: *const u8, m: c_int, desc_threshold: u32, out_idx: *mut i32, out_dist: *mut i32) -> c_int;
    pub fn vslam_distinctive_descriptors(h: vslam_handle, desc: *const u8, offsets: *const i32, m: c_int,
                                         out_idx: *mut i32) -> c_int;
    // frames resident in HBM: Frame::from_stereo (tracking_frame.rs:194-215) without re-uploading the features
    pub fn vslam_frame_retain(h: vslam_handle, index: c_int, img_w: f64, img_h: f64, frame_id: *mut c_int) -> c_int;
    pub fn vslam_frame_upload(h: vslam_handle, kps: *const vslam_keypoint, desc: *const u8, n: c_int, img_w: f64, img_h: f64,
                              frame_id: *mut c_int) -> c_int;
    pub fn vslam_frame_release(h: vslam_handle, frame_id: c_int) -> c_int;
    pub fn vslam_frame_num_features(h: vslam_handle, frame_id: c_int) -> c_int;
    pub fn vslam_frame_search_by_projection(h: vslam_handle, frame_id: c_int, uv: *const f64, mp_desc: *const u8, m: c_int,
                                            params: *const vslam_search_params, out_idx: *mut i32, out_dist: *mut i32) -> c_int;
    pub fn vslam_frame_grid_query(h: vslam_handle, frame_id: c_int, xyr: *const f64, nq: c_int, out_start: *mut i32,
                                  out_items: *mut i32, out_capacity: c_int) -> c_int;
    pub fn vslam_frame_bf_match_crosscheck(h: vslam_handle, query_frame_id: c_int, train_frame_id: c_int,
                                           matches: *mut vslam_dmatch, n_matches: *mut c_int) -> c_int;
    // search_for_triangulation / _bow (src/local_mapping/triangulation.rs:401-530, 541-658)
    pub fn vslam_search_for_triangulation(h: vslam_handle, kp1: *const vslam_keypoint, desc1: *const u8, flags1: *const u8,
                                          n1: c_int, kp2: *const vslam_keypoint, desc2: *const u8, flags2: *const u8, n2: c_int,
                                          params: *const vslam_tri_params, out_pairs: *mut i32, n_pairs: *mut c_int) -> c_int;
    pub fn vslam_search_for_triangulation_bow(h: vslam_handle, kp1: *const vslam_keypoint, desc1: *const u8, flags1: *const u8,
                                              n1: c_int, kp2: *const vslam_keypoint, desc2: *const u8, flags2: *const u8,
                                              n2: c_int, params: *const vslam_tri_params, n_groups: c_int, g1_start: *const i32,
                                              g1_items: *const i32, g2_start: *const i32, g2_items: *const i32,
                                              out_pairs: *mut i32, n_pairs: *mut c_int) -> c_int;
    // classic mode only (NOT reference parity): refined right coordinate / depth per left keypoint
    pub fn vslam_classic_stereo_extra(h: vslam_handle, index: c_int, u_right: *mut f32, depth: *mut f32, capacity: c_int) -> c_int;
    // OrbVocabulary (src/vocabulary/mod.rs): tree flattened once after load_from_text, quantisation per keyframe
    pub fn vslam_vocab_upload(h: vslam_handle, n_nodes: c_int, node_desc: *const u8, child_begin: *const i32,
                              child_ids: *const i32, parent: *const u32, word_id: *const i32) -> c_int;
    pub fn vslam_bow_quantize(h: vslam_handle, desc: *const u8, n: c_int, levels_up: c_int, out_word: *mut u32,
                              out_leaf: *mut u32, out_node: *mut u32) -> c_int;
}

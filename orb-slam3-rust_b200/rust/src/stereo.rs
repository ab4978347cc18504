//! Drop-in replacement of `src/tracking/frame/stereo.rs` (reference lines 15-66): same public items, same signatures.
//! `FeatureSet`, `StereoFrame`, `TH_HIGH/TH_LOW/NN_RATIO`, `descriptor_distance` keep their reference definitions; only
//! `StereoProcessor` changes: it owns a `vslam_handle` instead of a `Ptr<features2d::ORB>` and `process` is one FFI call.
//! Source only — the build image has no Rust toolchain; the C ABI underneath is what the test-suite exercises.
use anyhow::{anyhow, Result};
use nalgebra::Vector3;
use opencv::core::{DMatch, KeyPoint, Mat, Vector, CV_8U};
use opencv::prelude::*;
use std::ffi::CStr;

use crate::ffi::*;
use crate::tracking::frame::camera::CameraModel;
use crate::tracking::frame::stereo::{FeatureSet, StereoFrame};

pub struct StereoProcessor {
    handle: vslam_handle,
    camera: CameraModel,
    cap: usize,
    width: i32,
    height: i32,
}

// `process(&mut self)` is exclusive and all outputs are plain host memory: same Send story as the reference.
unsafe impl Send for StereoProcessor {}

impl StereoProcessor {
    /// Same signature as the reference (stereo.rs:37); the image size is EuRoC's, as hard-coded elsewhere in the
    /// reference (tracking_frame.rs:194-197).  Use `with_size` for other datasets.
    pub fn new(camera: CameraModel, n_features: i32) -> Result<Self> {
        Self::with_size(camera, n_features, 752, 480)
    }

    pub fn with_size(camera: CameraModel, n_features: i32, width: i32, height: i32) -> Result<Self> {
        let cfg = vslam_config {
            abi_version: VSLAM_ABI_VERSION, width, height, n_features,
            camera: vslam_camera { fx: camera.fx, fy: camera.fy, cx: camera.cx, cy: camera.cy, baseline: camera.baseline },
            max_batch: 1, ..Default::default()  // n_levels/scale/edge/fast = 0 -> the reference's 8 / 1.2 / 31 / 20
        };
        let mut handle: vslam_handle = std::ptr::null_mut();
        let rc = unsafe { vslam_create(&cfg, &mut handle) };
        if rc != VSLAM_OK {
            let msg = unsafe { CStr::from_ptr(vslam_last_error(std::ptr::null_mut())) }.to_string_lossy().into_owned();
            return Err(anyhow!("vslam_create failed ({rc}): {msg}"));
        }
        let cap = unsafe { vslam_keypoint_capacity(handle) } as usize;
        Ok(Self { handle, camera, cap, width, height })
    }

    /// Same signature and result as the reference's `process` (stereo.rs:52-66).
    pub fn process(&mut self, left: &Mat, right: &Mat, timestamp_ns: u64) -> Result<StereoFrame> {
        for m in [left, right] {
            if m.typ() != CV_8U || m.cols() != self.width || m.rows() != self.height {
                return Err(anyhow!("expected {}x{} 8UC1 images", self.width, self.height));
            }
        }
        let cap = self.cap;
        let mut kpl = vec![vslam_keypoint::default(); cap];
        let mut kpr = vec![vslam_keypoint::default(); cap];
        let (mut dl, mut dr) = (vec![0u8; cap * 32], vec![0u8; cap * 32]);
        let mut m = vec![vslam_dmatch::default(); cap];
        let (mut pts, mut valid) = (vec![0f64; cap * 3], vec![0u8; cap]);
        let mut out = vslam_frame_out {
            capacity: cap as i32, kp_left: kpl.as_mut_ptr(), kp_right: kpr.as_mut_ptr(), desc_left: dl.as_mut_ptr(),
            desc_right: dr.as_mut_ptr(), matches: m.as_mut_ptr(), points_cam: pts.as_mut_ptr(),
            points_valid: valid.as_mut_ptr(), n_left: 0, n_right: 0, n_matches: 0, status: 0, timestamp_ns: 0,
        };
        let rc = unsafe {
            vslam_process(self.handle, left.data(), left.mat_step().get(0), right.data(), right.mat_step().get(0),
                          timestamp_ns, &mut out)
        };
        if rc != VSLAM_OK {
            let msg = unsafe { CStr::from_ptr(vslam_last_error(self.handle)) }.to_string_lossy().into_owned();
            return Err(anyhow!("vslam_process failed ({rc}): {msg}"));
        }
        let feature_set = |kp: &[vslam_keypoint], d: &[u8], n: usize| -> Result<FeatureSet> {
            let mut keypoints = Vector::<KeyPoint>::with_capacity(n);
            for k in &kp[..n] {
                keypoints.push(KeyPoint::new_point((k.x, k.y).into(), k.size, k.angle, k.response, k.octave, k.class_id)?);
            }
            let descriptors = Mat::new_rows_cols_with_data(n as i32, 32, &d[..n * 32])?.try_clone()?;
            Ok(FeatureSet { keypoints, descriptors })
        };
        let (nl, nr, nm) = (out.n_left as usize, out.n_right as usize, out.n_matches as usize);
        let mut matches_lr = Vector::<DMatch>::with_capacity(nm);
        for d in &m[..nm] {
            matches_lr.push(DMatch { query_idx: d.query_idx, train_idx: d.train_idx, img_idx: d.img_idx, distance: d.distance });
        }
        let points_cam = (0..nl)
            .map(|i| (valid[i] != 0).then(|| Vector3::new(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])))
            .collect();
        Ok(StereoFrame {
            left_features: feature_set(&kpl, &dl, nl)?,
            right_features: feature_set(&kpr, &dr, nr)?,
            matches_lr, points_cam, timestamp_ns,
        })
    }
}

impl Drop for StereoProcessor {
    fn drop(&mut self) {
        unsafe { vslam_destroy(self.handle) }
    }
}

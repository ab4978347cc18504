//! GPU tree descent for `src/vocabulary/mod.rs`: `OrbVocabulary::transform` / `transform_bow_only` (reference lines
//! 290-360) keep their signatures; only `transform_descriptor` + `get_parent_at_level` (lines 230-275) move to the
//! device (`vslam_vocab_upload` once after `load_from_text`, `vslam_bow_quantize` per keyframe).  The `BowVector` /
//! `FeatureVector` accumulation and the L1 normalisation stay in Rust, exactly as in the reference.
//! Source only — the build image has no Rust toolchain; the C ABI underneath is what the test-suite exercises.
use anyhow::{anyhow, Result};
use opencv::core::Mat;
use opencv::prelude::*;
use std::ffi::CStr;

use crate::ffi::*;
use crate::vocabulary::{BowVector, FeatureVector, OrbVocabulary};

impl OrbVocabulary {
    /// Flattens `self.nodes` (children lists as CSR, `parent`, `word_id.map_or(-1)`) and keeps it on the device.
    pub fn upload_to_gpu(&self, handle: vslam_handle) -> Result<()> {
        let n = self.nodes.len();
        let mut desc = Vec::with_capacity(n * 32);
        let mut child_begin = Vec::with_capacity(n + 1);
        let mut child_ids: Vec<i32> = Vec::new();
        let mut parent = Vec::with_capacity(n);
        let mut word = Vec::with_capacity(n);
        child_begin.push(0i32);
        for node in &self.nodes {
            desc.extend_from_slice(&node.descriptor);
            child_ids.extend(node.children.iter().map(|&c| c as i32));
            child_begin.push(child_ids.len() as i32);
            parent.push(node.parent);
            word.push(node.word_id.map_or(-1i32, |w| w as i32));
        }
        if child_ids.is_empty() {
            child_ids.push(0);
        }
        let rc = unsafe {
            vslam_vocab_upload(handle, n as i32, desc.as_ptr(), child_begin.as_ptr(), child_ids.as_ptr(), parent.as_ptr(),
                               word.as_ptr())
        };
        check(handle, rc)
    }

    /// `transform` (reference lines 290-331) with the descent on the GPU.  `descriptors`: CV_8U, 32 columns, continuous.
    pub fn transform_gpu(&self, handle: vslam_handle, descriptors: &Mat, levels_up: usize) -> Result<(BowVector, FeatureVector)> {
        let rows = descriptors.rows().max(0) as usize;
        let (mut word, mut leaf, mut node) = (vec![0u32; rows + 1], vec![0u32; rows + 1], vec![0u32; rows + 1]);
        let rc = unsafe {
            vslam_bow_quantize(handle, descriptors.data(), rows as i32, levels_up as i32, word.as_mut_ptr(),
                               leaf.as_mut_ptr(), node.as_mut_ptr())
        };
        check(handle, rc)?;
        let mut bow = BowVector::new();
        let mut feat = FeatureVector::new();
        for i in 0..rows {
            let weight = self.nodes[leaf[i] as usize].weight;
            *bow.entry(word[i]).or_insert(0.0) += weight;
            feat.entry(node[i]).or_default().push(i);
        }
        let sum: f64 = bow.values().sum();
        if sum > 0.0 {
            for v in bow.values_mut() {
                *v /= sum;
            }
        }
        Ok((bow, feat))
    }
}

fn check(handle: vslam_handle, rc: i32) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(vslam_last_error(handle)) }.to_string_lossy().into_owned();
    Err(anyhow!("vslam_b200 error {}: {}", rc, msg))
}

//! `ArrowComputePipeline`: the reference records compute passes into one wgpu CommandEncoder and
//! submits them in `finish()` (crates/array/src/gpu_utils/compute_pipeline.rs:8-22, 259-273).
//! A CUDA stream already is an ordered queue, so by default every `*_op` enqueues at once and
//! `finish()` only closes the scope.  `new_captured` is the literal analogue: the ops recorded until
//! `finish()` are captured into a CUDA graph, `finish()` submits the whole program with one driver
//! call and `replay()` submits it again.  The `apply_*_function` family (:24-256) — the seam every
//! operator crate used to call with a (WGSL source, entry point) pair — is gone: operator crates
//! call the `agpu_*` entry points with an (op id, dtype id) pair instead.
use std::ptr;
use std::sync::Arc;

use super::ffi::*;
use super::GpuDevice;
use crate::array::buffer::ArrowGpuBuffer;

pub struct ArrowComputePipeline {
    pub device: Arc<GpuDevice>,
    label: Option<String>,
    capturing: bool,
    graph: *mut AgpuGraph,
}

impl ArrowComputePipeline {
    /// compute_pipeline.rs:14-22
    pub fn new(device: Arc<GpuDevice>, label: Option<&str>) -> Self {
        Self { device, label: label.map(str::to_owned), capturing: false, graph: ptr::null_mut() }
    }

    /// record-then-submit: everything enqueued until `finish()` becomes one CUDA graph
    pub fn new_captured(device: Arc<GpuDevice>, label: Option<&str>) -> Self {
        check(unsafe { agpu_graph_begin(device.handle()) }, "ArrowComputePipeline::new_captured");
        Self { device, label: label.map(str::to_owned), capturing: true, graph: ptr::null_mut() }
    }

    pub fn label(&self) -> Option<&str> {
        self.label.as_deref()
    }

    /// compute_pipeline.rs:275-282
    pub fn clone_buffer(&mut self, buffer: &ArrowGpuBuffer) -> ArrowGpuBuffer {
        self.device.clone_buffer(buffer)
    }

    /// compute_pipeline.rs:259-273 — never waits
    pub fn finish(&mut self) {
        if self.capturing {
            self.capturing = false;
            check(unsafe { agpu_graph_end(self.device.handle(), &mut self.graph) }, "ArrowComputePipeline::finish");
            self.replay();
        }
    }

    /// submit the recorded program again (same input buffers, outputs overwritten in place)
    pub fn replay(&self) {
        assert!(!self.graph.is_null(), "replay() needs a pipeline made with new_captured and finished");
        check(unsafe { agpu_graph_launch(self.device.handle(), self.graph) }, "ArrowComputePipeline::replay");
    }
}

impl Drop for ArrowComputePipeline {
    fn drop(&mut self) {
        if self.capturing {
            let mut g = ptr::null_mut();
            if unsafe { agpu_graph_end(self.device.handle(), &mut g) } == AGPU_OK && !g.is_null() {
                unsafe { agpu_graph_destroy(g) };
            }
        }
        if !self.graph.is_null() {
            unsafe { agpu_graph_destroy(self.graph) };
        }
    }
}

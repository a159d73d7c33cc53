//! array/src/kernels/broadcast.rs:6-17
use std::sync::Arc;

use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};

pub trait Broadcast<Rhs>: Sized {
    fn broadcast(value: Rhs, len: usize, gpu_device: Arc<GpuDevice>) -> Self {
        let mut pipeline = ArrowComputePipeline::new(gpu_device, Some("broadcast"));
        let out = Self::broadcast_op(value, len, &mut pipeline);
        pipeline.finish();
        out
    }

    fn broadcast_op(value: Rhs, len: usize, pipeline: &mut ArrowComputePipeline) -> Self;
}

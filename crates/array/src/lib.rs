//! arrow_gpu_array — columns in CUDA device memory (drop-in for psvri/arrow-gpu's `array` crate).
//! Public names and fields are the reference's (crates/array/src/lib.rs, array/mod.rs); buffers
//! are stream-ordered CUDA allocations behind `libagpu.so` instead of `Arc<wgpu::Buffer>`.
pub mod array;
pub mod gpu_utils;
pub mod kernels;
pub mod utils;

use std::sync::{Arc, LazyLock};

use gpu_utils::GpuDevice;

/// crates/array/src/lib.rs:10-13
#[derive(Debug)]
pub enum ArrowErrorGPU {
    OperationNotSupported(String),
    CastingNotSupported(String),
}

/// crates/array/src/lib.rs:16-17
pub static GPU_DEVICE: LazyLock<Arc<GpuDevice>> = LazyLock::new(|| Arc::new(GpuDevice::new()));

pub mod compute_pipeline;
pub mod ffi;
pub mod gpu_device;

pub use compute_pipeline::ArrowComputePipeline;
pub use gpu_device::GpuDevice;

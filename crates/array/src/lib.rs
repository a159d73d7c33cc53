//! arrow_gpu_array — columns in CUDA device memory (drop-in for psvri/arrow-gpu's `array` crate).
//! Public names and fields are the reference's (crates/array/src/lib.rs, array/mod.rs); buffers
//! are stream-ordered CUDA allocations behind `libagpu.so` instead of `Arc<wgpu::Buffer>`.
pub mod array;
pub mod gpu_utils;

/// Scalars and the `Broadcast` trait (the reference's `array/src/kernels/`): declarations only, so
/// they live here instead of in one-enum files.
pub mod kernels {
    use crate::array::ArrowArrayGPU;

    /// A scalar of any element type, as `broadcast_dyn` takes it (kernels/mod.rs:7-17)
    #[derive(Debug)]
    pub enum ScalarValue {
        F32(f32),
        U32(u32),
        U16(u16),
        U8(u8),
        I32(i32),
        I16(i16),
        I8(i8),
        BOOL(bool),
    }

    /// Either side of a binary kernel (kernels/mod.rs:20-23)
    #[derive(Debug)]
    pub enum Operand {
        Scalar(ScalarValue),
        Array(ArrowArrayGPU),
    }

    pub mod broadcast {
        use std::sync::Arc;

        use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};

        /// `len` copies of one value (kernels/broadcast.rs:6-17); the eager form opens a pipeline,
        /// records the op and finishes, like every other eager method of the crates
        pub trait Broadcast<Rhs>: Sized {
            fn broadcast(value: Rhs, len: usize, gpu_device: Arc<GpuDevice>) -> Self {
                let mut pipeline = ArrowComputePipeline::new(gpu_device, Some("broadcast"));
                let out = Self::broadcast_op(value, len, &mut pipeline);
                pipeline.finish();
                out
            }

            fn broadcast_op(value: Rhs, len: usize, pipeline: &mut ArrowComputePipeline) -> Self;
        }
    }
}

/// Host-side vectors of any element type, what `ArrowArrayGPU::get_raw_values` returns
/// (the reference's `array/src/utils/mod.rs`)
pub mod utils {
    #[derive(Debug, PartialEq)]
    pub enum ScalarArray {
        F32Vec(Vec<f32>),
        U32Vec(Vec<u32>),
        U16Vec(Vec<u16>),
        U8Vec(Vec<u8>),
        I32Vec(Vec<i32>),
        I16Vec(Vec<i16>),
        I8Vec(Vec<i8>),
        BOOLVec(Vec<bool>),
    }

    macro_rules! into_scalar_array {
        ($($t:ty => $variant:ident),*) => {$(
            impl From<Vec<$t>> for ScalarArray {
                fn from(value: Vec<$t>) -> Self {
                    ScalarArray::$variant(value)
                }
            }
        )*};
    }
    into_scalar_array!(f32 => F32Vec, u32 => U32Vec, u16 => U16Vec, u8 => U8Vec, i32 => I32Vec, i16 => I16Vec, i8 => I8Vec,
                       bool => BOOLVec);
}

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

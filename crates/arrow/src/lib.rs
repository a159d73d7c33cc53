//! arrow_gpu — the umbrella crate (crates/arrow/src/lib.rs, kernels.rs): re-exports only.
pub use arrow_gpu_array::*;

pub mod kernels {
    pub use arrow_gpu_arithmetic::*;
    pub use arrow_gpu_array::kernels::broadcast::*;
    pub use arrow_gpu_cast::*;
    pub use arrow_gpu_compare::*;
    pub use arrow_gpu_logical::*;
    pub use arrow_gpu_math::*;
    pub use arrow_gpu_routines::*;
    pub use arrow_gpu_trigonometry::*;
}

//! array/src/kernels/mod.rs
use crate::array::ArrowArrayGPU;

pub mod broadcast;

/// Enum of scalar values used in kernels (kernels/mod.rs:7-17)
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

/// Enum of operands (kernels/mod.rs:20-23)
#[derive(Debug)]
pub enum Operand {
    Scalar(ScalarValue),
    Array(ArrowArrayGPU),
}

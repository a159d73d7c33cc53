//! Array types (crates/array/src/array/mod.rs): `ArrowType`, `ArrowPrimitiveType`, the typed
//! aliases of `PrimitiveArrayGpu<T>`, `BooleanArrayGPU`, the `ArrowArrayGPU` enum and `broadcast_dyn`.
use std::fmt::Debug;
use std::os::raw::c_int;
use std::sync::Arc;

use crate::gpu_utils::ffi::*;
use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};
use crate::kernels::broadcast::Broadcast;
use crate::kernels::ScalarValue;
use crate::utils::ScalarArray;

pub mod boolean_gpu;
pub mod buffer;
pub mod null_bit_buffer;
pub mod primitive_array_gpu;

pub use boolean_gpu::BooleanArrayGPU;
pub use null_bit_buffer::*;
pub use primitive_array_gpu::PrimitiveArrayGpu;

/// array/mod.rs:40-50
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
#[non_exhaustive]
pub enum ArrowType {
    BooleanType,
    Float32Type,
    UInt32Type,
    UInt16Type,
    UInt8Type,
    Int32Type,
    Int16Type,
    Int8Type,
    Date32Type,
}

impl ArrowType {
    /// the `agpu_dtype` id of include/agpu.h
    pub fn dtype_id(&self) -> c_int {
        match self {
            ArrowType::BooleanType => AGPU_BOOL,
            ArrowType::Float32Type => AGPU_F32,
            ArrowType::UInt32Type => AGPU_U32,
            ArrowType::UInt16Type => AGPU_U16,
            ArrowType::UInt8Type => AGPU_U8,
            ArrowType::Int32Type => AGPU_I32,
            ArrowType::Int16Type => AGPU_I16,
            ArrowType::Int8Type => AGPU_I8,
            ArrowType::Date32Type => AGPU_DATE32,
        }
    }
}

/// array/mod.rs:53-62 (bytemuck::Pod is not needed: values cross the ABI as raw bytes)
pub trait RustNativeType: Copy + Debug + Default + 'static {}
impl RustNativeType for i32 {}
impl RustNativeType for i16 {}
impl RustNativeType for i8 {}
impl RustNativeType for f32 {}
impl RustNativeType for u32 {}
impl RustNativeType for u16 {}
impl RustNativeType for u8 {}

/// marker of `Date32ArrayGPU` (array/date32_gpu.rs): i32 storage
#[derive(Debug, Clone, Copy, Default)]
pub struct Date32Type;

/// array/mod.rs:64-85, plus the ids the C ABI dispatches on.  sub-word columns are native 1/2-byte
/// lanes here (`ITEM_SIZE` of i16 is 2: the reference's `4` is its u32-packed shader workaround)
pub trait ArrowPrimitiveType: Send + Sync + 'static {
    type NativeType: RustNativeType;
    const ITEM_SIZE: u64;
    const DTYPE: c_int;
    const ARROW_TYPE: ArrowType;
}

macro_rules! impl_primitive_type {
    ($marker:ty, $native:ty, $size:expr, $dtype:expr, $arrow:ident) => {
        impl ArrowPrimitiveType for $marker {
            type NativeType = $native;
            const ITEM_SIZE: u64 = $size;
            const DTYPE: c_int = $dtype;
            const ARROW_TYPE: ArrowType = ArrowType::$arrow;
        }
    };
}
impl_primitive_type!(f32, f32, 4, AGPU_F32, Float32Type);
impl_primitive_type!(u32, u32, 4, AGPU_U32, UInt32Type);
impl_primitive_type!(u16, u16, 2, AGPU_U16, UInt16Type);
impl_primitive_type!(u8, u8, 1, AGPU_U8, UInt8Type);
impl_primitive_type!(i32, i32, 4, AGPU_I32, Int32Type);
impl_primitive_type!(i16, i16, 2, AGPU_I16, Int16Type);
impl_primitive_type!(i8, i8, 1, AGPU_I8, Int8Type);
impl_primitive_type!(Date32Type, i32, 4, AGPU_DATE32, Date32Type);

pub type Float32ArrayGPU = PrimitiveArrayGpu<f32>;
pub type UInt32ArrayGPU = PrimitiveArrayGpu<u32>;
pub type UInt16ArrayGPU = PrimitiveArrayGpu<u16>;
pub type UInt8ArrayGPU = PrimitiveArrayGpu<u8>;
pub type Int32ArrayGPU = PrimitiveArrayGpu<i32>;
pub type Int16ArrayGPU = PrimitiveArrayGpu<i16>;
pub type Int8ArrayGPU = PrimitiveArrayGpu<i8>;
pub type Date32ArrayGPU = PrimitiveArrayGpu<Date32Type>;

/// Marker traits grouping element types by their storage (the reference's `array/types.rs`), used
/// by the operator crates' `impl<S: Int32Type> ... for Int32ArrayGPU` blocks (arithmetic/src/i32.rs)
pub mod types {
    use super::Date32Type;

    /// column of i32 rows
    pub trait Int32Type {}
    impl Int32Type for i32 {}
    impl Int32Type for Date32Type {}

    /// column of f32 rows
    pub trait Float32Type {}
    impl Float32Type for f32 {}

    /// column of u32 rows
    pub trait UInt32Type {}
    impl UInt32Type for u32 {}

    /// column of u16 rows
    pub trait UInt16Type {}
    impl UInt16Type for u16 {}
}

/// array/mod.rs:96-99
pub trait ArrayUtils {
    fn get_gpu_device(&self) -> Arc<GpuDevice>;
}

/// array/mod.rs:101-114
#[derive(Debug)]
#[non_exhaustive]
pub enum ArrowArrayGPU {
    Float32ArrayGPU(Float32ArrayGPU),
    UInt32ArrayGPU(UInt32ArrayGPU),
    UInt16ArrayGPU(UInt16ArrayGPU),
    UInt8ArrayGPU(UInt8ArrayGPU),
    Int32ArrayGPU(Int32ArrayGPU),
    Int16ArrayGPU(Int16ArrayGPU),
    Int8ArrayGPU(Int8ArrayGPU),
    Date32ArrayGPU(Date32ArrayGPU),
    BooleanArrayGPU(BooleanArrayGPU),
}

/// run `$body` with `$x` bound to the array inside any variant
#[macro_export]
macro_rules! for_each_array {
    ($value:expr, $x:ident => $body:expr) => {
        match $value {
            $crate::array::ArrowArrayGPU::Float32ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::UInt32ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::UInt16ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::UInt8ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::Int32ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::Int16ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::Int8ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::Date32ArrayGPU($x) => $body,
            $crate::array::ArrowArrayGPU::BooleanArrayGPU($x) => $body,
        }
    };
}

impl ArrowArrayGPU {
    pub fn get_gpu_device(&self) -> Arc<GpuDevice> {
        for_each_array!(self, x => x.gpu_device.clone())
    }

    pub fn get_dtype(&self) -> ArrowType {
        for_each_array!(self, x => x.arrow_type())
    }

    pub fn len(&self) -> usize {
        for_each_array!(self, x => x.len)
    }

    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }

    /// array/mod.rs:146-159
    pub fn get_raw_values(&self) -> ScalarArray {
        for_each_array!(self, x => x.raw_values().unwrap().into())
    }

    /// array/mod.rs:161-175 (`BooleanArrayGPU` is `todo!()` there; cloned here as well)
    pub fn clone_array(&self) -> ArrowArrayGPU {
        for_each_array!(self, x => x.clone_array().into())
    }
}

macro_rules! impl_into_enum {
    ($($variant:ident),*) => {$(
        impl From<$variant> for ArrowArrayGPU {
            fn from(value: $variant) -> Self {
                ArrowArrayGPU::$variant(value)
            }
        }
        impl TryFrom<ArrowArrayGPU> for $variant {
            type Error = crate::ArrowErrorGPU;
            /// f32_gpu.rs:45-57 and friends
            fn try_from(value: ArrowArrayGPU) -> Result<Self, Self::Error> {
                match value {
                    ArrowArrayGPU::$variant(x) => Ok(x),
                    other => Err(crate::ArrowErrorGPU::CastingNotSupported(format!(
                        "could not cast {:?} into {}", other.get_dtype(), stringify!($variant)))),
                }
            }
        }
    )*};
}
impl_into_enum!(Float32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, UInt8ArrayGPU, Int32ArrayGPU, Int16ArrayGPU, Int8ArrayGPU,
                Date32ArrayGPU, BooleanArrayGPU);

/// array/mod.rs:181-192
pub fn broadcast_dyn(value: ScalarValue, len: usize, device: Arc<GpuDevice>) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(device, Some("broadcast"));
    let out = broadcast_op_dyn(value, len, &mut pipeline);
    pipeline.finish();
    out
}

/// array/mod.rs:196-211
pub fn broadcast_op_dyn(value: ScalarValue, len: usize, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    match value {
        ScalarValue::F32(x) => Float32ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::U32(x) => UInt32ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::U16(x) => UInt16ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::U8(x) => UInt8ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::I32(x) => Int32ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::I16(x) => Int16ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::I8(x) => Int8ArrayGPU::broadcast_op(x, len, pipeline).into(),
        ScalarValue::BOOL(x) => BooleanArrayGPU::broadcast_op(x, len, pipeline).into(),
    }
}

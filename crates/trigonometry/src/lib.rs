//! arrow_gpu_trigonometry — `sin cos acos`, `sinh` (drop-in for crates/trigonometry, lib.rs:22-68).
//! Integer columns (i8, u8, i16, u16) give `Float32ArrayGPU`: the int -> f32 cast is fused into the
//! kernel like the reference's `{i8,u8,i16,u16}/trigonometry.wgsl`.
use std::os::raw::c_int;

use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

macro_rules! eager {
    ($self:ident, $op:ident) => {{
        let mut pipeline = ArrowComputePipeline::new($self.get_gpu_device(), None);
        let output = $self.$op(&mut pipeline);
        pipeline.finish();
        output
    }};
}

/// `sinh`, row by row (trigonometry/src/lib.rs:22-33 of the reference)
pub trait Hyperbolic: ArrayUtils {
    type Output;
    fn sinh(&self) -> Self::Output {
        eager!(self, sinh_op)
    }
    fn sinh_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// `cos`, `sin`, `acos`, row by row (trigonometry/src/lib.rs:49-68 of the reference)
pub trait Trigonometric: ArrayUtils {
    type Output;
    fn cos(&self) -> Self::Output {
        eager!(self, cos_op)
    }
    fn sin(&self) -> Self::Output {
        eager!(self, sin_op)
    }
    fn acos(&self) -> Self::Output {
        eager!(self, acos_op)
    }
    fn cos_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::Output;
    fn sin_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::Output;
    fn acos_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// Markers of the element types each family supports (f32_kernel.rs, i8/u8/i16/u16_kernel.rs)
pub trait HyperbolicType {}
pub trait TrigonometricType {}
macro_rules! mark { ($($t:ty),*) => { $(impl HyperbolicType for $t {} impl TrigonometricType for $t {})* }; }
mark!(f32, i8, u8, i16, u16);

fn to_f32_kernel<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, what: &str) -> Float32ArrayGPU {
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref()]);
    let out = Float32ArrayGPU::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_unary(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), out.data.ptr(), a.len, a.validity_ptr(),
                       NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

impl<T: HyperbolicType + ArrowPrimitiveType> Hyperbolic for PrimitiveArrayGpu<T> {
    type Output = Float32ArrayGPU;
    fn sinh_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        to_f32_kernel(AGPU_SINH, self, "sinh_op")
    }
}

impl<T: TrigonometricType + ArrowPrimitiveType> Trigonometric for PrimitiveArrayGpu<T> {
    type Output = Float32ArrayGPU;
    fn cos_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        to_f32_kernel(AGPU_COS, self, "cos_op")
    }
    fn sin_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        to_f32_kernel(AGPU_SIN, self, "sin_op")
    }
    /// acos exists for f32 only in the reference's dyn matrix (lib.rs:191-201): ints panic there
    fn acos_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        to_f32_kernel(AGPU_ACOS, self, "acos_op")
    }
}

/// trigonometry/src/lib.rs:139-202
macro_rules! dyn_unary {
    ($([$(#[$doc:meta])* $dyn:ident, $op_dyn:ident, $method:ident, $($arr:ident),+]),*) => {$(
        $(#[$doc])*
        pub fn $dyn(data: &ArrowArrayGPU) -> ArrowArrayGPU {
            let mut pipeline = ArrowComputePipeline::new(data.get_gpu_device(), None);
            let result = $op_dyn(data, &mut pipeline);
            pipeline.finish();
            result
        }

        pub fn $op_dyn(data: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
            match data {
                $(ArrowArrayGPU::$arr(x) => x.$method(pipeline).into(),)+
                _ => panic!("Operation {} not supported for type {:?}", stringify!($op_dyn), data.get_dtype()),
            }
        }
    )*};
}
dyn_unary!(
    [/// sinh(x) of every row
     sinh_dyn, sinh_op_dyn, sinh_op, Float32ArrayGPU, UInt16ArrayGPU, UInt8ArrayGPU, Int16ArrayGPU, Int8ArrayGPU],
    [/// cos(x) of every row
     cos_dyn, cos_op_dyn, cos_op, Float32ArrayGPU, UInt16ArrayGPU, UInt8ArrayGPU, Int16ArrayGPU, Int8ArrayGPU],
    [/// sin(x) of every row
     sin_dyn, sin_op_dyn, sin_op, Float32ArrayGPU, UInt16ArrayGPU, UInt8ArrayGPU, Int16ArrayGPU, Int8ArrayGPU],
    [/// acos(x) of every row
     acos_dyn, acos_op_dyn, acos_op, Float32ArrayGPU]
);

//! arrow_gpu_math — `abs`, `power`, `sqrt cbrt exp exp2 log log2` (drop-in for crates/math,
//! lib.rs:37-122).  f32 functions are within 2 ULP of the correctly rounded value (sqrt: exact);
//! i32 `power` is bit-identical to the reference's O(|p|) multiply/divide loops
//! (math/compute_shaders/i32/binary.wgsl:13-29) in O(log |p|).
use std::os::raw::c_int;

use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

macro_rules! eager {
    ($self:ident, $op:ident $(, $arg:ident)*) => {{
        let mut pipeline = ArrowComputePipeline::new($self.get_gpu_device(), None);
        let output = $self.$op($($arg,)* &mut pipeline);
        pipeline.finish();
        output
    }};
}

/// Trait for math unary operation on each element of the array
pub trait MathUnary: ArrayUtils {
    type OutputType;
    fn abs(&self) -> Self::OutputType {
        eager!(self, abs_op)
    }
    fn abs_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
}

/// Trait for math binary operation on each pair of elements
pub trait MathBinary: ArrayUtils + Sized {
    type OutputType;
    fn power(&self, other: &Self) -> Self::OutputType {
        eager!(self, power_op, other)
    }
    fn power_op(&self, other: &Self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
}

/// Trait for float-only math functions on each element of the array
pub trait FloatMathUnary: ArrayUtils {
    type OutputType;
    fn sqrt(&self) -> Self::OutputType {
        eager!(self, sqrt_op)
    }
    fn cbrt(&self) -> Self::OutputType {
        eager!(self, cbrt_op)
    }
    fn exp(&self) -> Self::OutputType {
        eager!(self, exp_op)
    }
    fn exp2(&self) -> Self::OutputType {
        eager!(self, exp2_op)
    }
    fn log(&self) -> Self::OutputType {
        eager!(self, log_op)
    }
    fn log2(&self) -> Self::OutputType {
        eager!(self, log2_op)
    }
    fn sqrt_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
    fn cbrt_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
    fn exp_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
    fn exp2_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
    fn log_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
    fn log2_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
}

/// Markers of the element types each family supports (f32 and i32: math/src/f32.rs, i32.rs)
pub trait MathUnaryType {}
pub trait MathBinaryType {}
pub trait FloatMathUnaryType {}
impl MathUnaryType for f32 {}
impl MathUnaryType for i32 {}
impl MathBinaryType for f32 {}
impl MathBinaryType for i32 {}
impl FloatMathUnaryType for f32 {}

fn unary_kernel<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, what: &str) -> PrimitiveArrayGpu<T> {
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_unary(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), out.data.ptr(), a.len, a.validity_ptr(),
                       NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

impl<T: MathUnaryType + ArrowPrimitiveType> MathUnary for PrimitiveArrayGpu<T> {
    type OutputType = Self;
    fn abs_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_ABS, self, "abs_op")
    }
}

impl<T: MathBinaryType + ArrowPrimitiveType> MathBinary for PrimitiveArrayGpu<T> {
    type OutputType = Self;
    fn power_op(&self, other: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        assert_eq!(self.len, other.len, "power_op: length mismatch");
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, self.len, &[self.null_buffer.as_ref(), other.null_buffer.as_ref()]);
        let out = Self::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_binary(self.gpu_device.handle(), AGPU_POW, T::DTYPE, self.values_ptr(), other.values_ptr(), out.data.ptr(), self.len,
                            self.validity_ptr(), other.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "power_op",
        );
        out
    }
}

impl<T: FloatMathUnaryType + ArrowPrimitiveType> FloatMathUnary for PrimitiveArrayGpu<T> {
    type OutputType = Self;
    fn sqrt_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_SQRT, self, "sqrt_op")
    }
    fn cbrt_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_CBRT, self, "cbrt_op")
    }
    fn exp_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_EXP, self, "exp_op")
    }
    fn exp2_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_EXP2, self, "exp2_op")
    }
    fn log_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_LOG, self, "log_op")
    }
    fn log2_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        unary_kernel(AGPU_LOG2, self, "log2_op")
    }
}

/// math/src/lib.rs:239-348
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
    [/// abs(x) of every row
     abs_dyn, abs_op_dyn, abs_op, Float32ArrayGPU, Int32ArrayGPU],
    [/// square_root(x) of every row
     sqrt_dyn, sqrt_op_dyn, sqrt_op, Float32ArrayGPU],
    [/// cube_root(x) of every row
     cbrt_dyn, cbrt_op_dyn, cbrt_op, Float32ArrayGPU],
    [/// e^x of every row
     exp_dyn, exp_op_dyn, exp_op, Float32ArrayGPU],
    [/// 2^x of every row
     exp2_dyn, exp2_op_dyn, exp2_op, Float32ArrayGPU],
    [/// log(x) of every row
     log_dyn, log_op_dyn, log_op, Float32ArrayGPU],
    [/// log_to_base_2(x) of every row
     log2_dyn, log2_op_dyn, log2_op, Float32ArrayGPU]
);

/// `x` to the power `y`, row by row over `self` and `other`
pub fn power_dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(data_1.get_gpu_device(), None);
    let result = power_op_dyn(data_1, data_2, &mut pipeline);
    pipeline.finish();
    result
}

pub fn power_op_dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    match (data_1, data_2) {
        (ArrowArrayGPU::Float32ArrayGPU(a), ArrowArrayGPU::Float32ArrayGPU(b)) => a.power_op(b, pipeline).into(),
        (ArrowArrayGPU::Int32ArrayGPU(a), ArrowArrayGPU::Int32ArrayGPU(b)) => a.power_op(b, pipeline).into(),
        _ => panic!("Operation power_dyn not supported for type {:?} {:?}", data_1.get_dtype(), data_2.get_dtype()),
    }
}

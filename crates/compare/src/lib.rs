//! arrow_gpu_compare — `gt gteq lt lteq eq` -> packed `BooleanArrayGPU`, `min` / `max`
//! (drop-in for crates/compare, lib.rs:41-83).  The compare kernel emits whole bitmap words
//! (lanes merge their predicate bits with warp shuffles) and ANDs both validity bitmaps in the same
//! pass; the reference used workgroup-shared atomicOr + a barrier and a 32x over-allocated,
//! zero-filled output (lib.rs:85-111, compute_shaders/*/cmp.wgsl).
use std::os::raw::c_int;

use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

macro_rules! eager {
    ($self:ident, $op:ident, $operand:ident) => {{
        let mut pipeline = ArrowComputePipeline::new($self.get_gpu_device(), None);
        let output = $self.$op($operand, &mut pipeline);
        pipeline.finish();
        output
    }};
}

/// Marker of the element types that can be compared (the reference's helper carried shader text)
pub trait CompareType {}

/// Row-wise comparisons producing a `BooleanArrayGPU` (compare/src/lib.rs:41-62 of the reference)
pub trait Compare: ArrayUtils {
    fn gt(&self, operand: &Self) -> BooleanArrayGPU {
        eager!(self, gt_op, operand)
    }
    fn gteq(&self, operand: &Self) -> BooleanArrayGPU {
        eager!(self, gteq_op, operand)
    }
    fn lt(&self, operand: &Self) -> BooleanArrayGPU {
        eager!(self, lt_op, operand)
    }
    fn lteq(&self, operand: &Self) -> BooleanArrayGPU {
        eager!(self, lteq_op, operand)
    }
    fn eq(&self, operand: &Self) -> BooleanArrayGPU {
        eager!(self, eq_op, operand)
    }
    fn gt_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU;
    fn gteq_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU;
    fn lt_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU;
    fn lteq_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU;
    fn eq_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU;
}

/// Trait for element-wise min / max of two ArrowArrays
pub trait MinMax: ArrayUtils + Sized {
    fn max(&self, operand: &Self) -> Self {
        eager!(self, max_op, operand)
    }
    fn min(&self, operand: &Self) -> Self {
        eager!(self, min_op, operand)
    }
    fn max_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> Self;
    fn min_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> Self;
}

macro_rules! mark { ($($t:ty),*) => { $(impl CompareType for $t {})* }; }
mark!(f32, u32, u16, u8, i32, i16, i8, Date32Type);

fn compare_kernel<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, b: &PrimitiveArrayGpu<T>, what: &str) -> BooleanArrayGPU {
    assert_eq!(a.len, b.len, "{what}: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), b.null_buffer.as_ref()]);
    let out = BooleanArrayGPU::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_compare(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), b.values_ptr(), out.data.ptr() as *mut u32, a.len,
                         a.validity_ptr(), b.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

fn min_max_kernel<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, b: &PrimitiveArrayGpu<T>, what: &str) -> PrimitiveArrayGpu<T> {
    assert_eq!(a.len, b.len, "{what}: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), b.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_binary(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), b.values_ptr(), out.data.ptr(), a.len,
                        a.validity_ptr(), b.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

impl<T: CompareType + ArrowPrimitiveType> Compare for PrimitiveArrayGpu<T> {
    fn gt_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU {
        compare_kernel(AGPU_GT, self, operand, "gt_op")
    }
    fn gteq_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU {
        compare_kernel(AGPU_GTEQ, self, operand, "gteq_op")
    }
    fn lt_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU {
        compare_kernel(AGPU_LT, self, operand, "lt_op")
    }
    fn lteq_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU {
        compare_kernel(AGPU_LTEQ, self, operand, "lteq_op")
    }
    fn eq_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> BooleanArrayGPU {
        compare_kernel(AGPU_EQ, self, operand, "eq_op")
    }
}

impl<T: CompareType + ArrowPrimitiveType> MinMax for PrimitiveArrayGpu<T> {
    fn max_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        min_max_kernel(AGPU_MAX, self, operand, "max_op")
    }
    fn min_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        min_max_kernel(AGPU_MIN, self, operand, "min_op")
    }
}

/// compare lib.rs:174-334: `$dyn(data_1, data_2)` / `$op_dyn(data_1, data_2, pipeline)` over every
/// numeric variant; mismatched or boolean operands `panic!`
macro_rules! dyn_same_type {
    ($([$(#[$doc:meta])* $dyn:ident, $op_dyn:ident, $method:ident]),*) => {$(
        $(#[$doc])*
        pub fn $dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU) -> ArrowArrayGPU {
            let mut pipeline = ArrowComputePipeline::new(data_1.get_gpu_device(), None);
            let result = $op_dyn(data_1, data_2, &mut pipeline);
            pipeline.finish();
            result
        }

        pub fn $op_dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
            use ArrowArrayGPU::*;
            match (data_1, data_2) {
                (Float32ArrayGPU(a), Float32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (UInt32ArrayGPU(a), UInt32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (UInt16ArrayGPU(a), UInt16ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (UInt8ArrayGPU(a), UInt8ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int32ArrayGPU(a), Int32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int16ArrayGPU(a), Int16ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int8ArrayGPU(a), Int8ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Date32ArrayGPU(a), Date32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                _ => panic!("Operation {} not supported for type {:?} {:?}", stringify!($dyn), data_1.get_dtype(), data_2.get_dtype()),
            }
        }
    )*};
}
dyn_same_type!(
    [/// row-wise predicate x > y into a BooleanArrayGPU
     gt_dyn, gt_op_dyn, gt_op],
    [/// row-wise predicate x >= y into a BooleanArrayGPU
     gteq_dyn, gteq_op_dyn, gteq_op],
    [/// row-wise predicate x < y into a BooleanArrayGPU
     lt_dyn, lt_op_dyn, lt_op],
    [/// row-wise predicate x <= y into a BooleanArrayGPU
     lteq_dyn, lteq_op_dyn, lteq_op],
    [/// row-wise predicate x == y into a BooleanArrayGPU
     eq_dyn, eq_op_dyn, eq_op],
    [/// max(x, y), row by row over both columns
     max_dyn, max_op_dyn, max_op],
    [/// min(x, y), row by row over both columns
     min_dyn, min_op_dyn, min_op]
);

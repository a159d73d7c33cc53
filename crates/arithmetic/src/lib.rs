//! arrow_gpu_arithmetic — `+ - * / %`, `neg`, `sum` (drop-in for crates/arithmetic).
//! Trait and function names are the reference's (arithmetic_kernels.rs:18-75, 178-223, 270-280;
//! aggregate_kernels.rs:7-17); every `*_op` is ONE `agpu_*` call — value kernel and validity
//! bitmap in the same pass — instead of `apply_scalar_function` / `apply_binary_function` with a
//! WGSL source plus a separate null-bitmap dispatch (lib.rs:11-94 of the reference).
use std::os::raw::c_int;

use arrow_gpu_array::array::types::Int32Type;
use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

/// `fn name(&self, value) { new pipeline; name_op; finish }` — the reference's `default_impl!`
macro_rules! eager {
    ($self:ident, $op:ident $(, $arg:ident)*) => {{
        let mut pipeline = ArrowComputePipeline::new($self.get_gpu_device(), None);
        let output = $self.$op($($arg,)* &mut pipeline);
        pipeline.finish();
        output
    }};
}

/// The addition operator ArrowArray + Scalar (the scalar is a one-element array on the device)
pub trait ArrowScalarAdd<Rhs>: ArrayUtils {
    type Output;
    fn add_scalar(&self, value: &Rhs) -> Self::Output {
        eager!(self, add_scalar_op, value)
    }
    fn add_scalar_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// column − one-element column
pub trait ArrowScalarSub<Rhs>: ArrayUtils {
    type Output;
    fn sub_scalar(&self, value: &Rhs) -> Self::Output {
        eager!(self, sub_scalar_op, value)
    }
    fn sub_scalar_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// The multiply operator ArrowArray * Scalar
pub trait ArrowScalarMul<Rhs>: ArrayUtils {
    type Output;
    fn mul_scalar(&self, value: &Rhs) -> Self::Output {
        eager!(self, mul_scalar_op, value)
    }
    fn mul_scalar_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// The division operator ArrowArray / Scalar (x / 0 = x for integers: WGSL semantics)
pub trait ArrowScalarDiv<Rhs>: ArrayUtils {
    type Output;
    fn div_scalar(&self, value: &Rhs) -> Self::Output {
        eager!(self, div_scalar_op, value)
    }
    fn div_scalar_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// The remainder operator ArrowArray % Scalar (x % 0 = 0 for integers)
pub trait ArrowScalarRem<Rhs>: ArrayUtils {
    type Output;
    fn rem_scalar(&self, value: &Rhs) -> Self::Output {
        eager!(self, rem_scalar_op, value)
    }
    fn rem_scalar_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// column + column
pub trait ArrowAdd<Rhs>: ArrayUtils {
    type Output;
    fn add(&self, value: &Rhs) -> Self::Output {
        eager!(self, add_op, value)
    }
    fn add_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// column − column
pub trait ArrowSub<Rhs>: ArrayUtils {
    type Output;
    fn sub(&self, value: &Rhs) -> Self::Output {
        eager!(self, sub_op, value)
    }
    fn sub_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// The multiply operator ArrowArray * ArrowArray
pub trait ArrowMul<Rhs>: ArrayUtils {
    type Output;
    fn mul(&self, value: &Rhs) -> Self::Output {
        eager!(self, mul_op, value)
    }
    fn mul_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// column ÷ column
pub trait ArrowDiv<Rhs>: ArrayUtils {
    type Output;
    fn div(&self, value: &Rhs) -> Self::Output {
        eager!(self, div_op, value)
    }
    fn div_op(&self, value: &Rhs, pipeline: &mut ArrowComputePipeline) -> Self::Output;
}

/// The negation operator -ArrowArray (arithmetic_kernels.rs:270-280)
pub trait Neg: ArrayUtils {
    type OutputType;
    fn neg(&self) -> Self::OutputType {
        eager!(self, neg_op)
    }
    fn neg_op(&self, pipeline: &mut ArrowComputePipeline) -> Self::OutputType;
}

/// Marker of the element types that support negation (the reference's helper carried the shader text)
pub trait NegUnaryType {}
impl NegUnaryType for f32 {}

/// Trait for sum of all elements in the array (aggregate_kernels.rs:7-17)
pub trait Sum: ArrayUtils + Sized {
    fn sum(&self) -> Self {
        eager!(self, sum_op)
    }
    fn sum_op(&self, pipeline: &mut ArrowComputePipeline) -> Self;
}

/// Marker of the 32-bit element types that support sum
pub trait Sum32Bit: ArrowPrimitiveType {}
impl Sum32Bit for f32 {}
impl Sum32Bit for i32 {}
impl Sum32Bit for u32 {}

// ---------------------------------------------------------------------------------------------
// the two launchers every impl below ends in
// ---------------------------------------------------------------------------------------------
/// out[i] = a[i] op *scalar; the scalar is a ONE-element array on the device, like the reference
/// (lib.rs:11-50); validity is copied (lib.rs:35-38)
pub fn scalar_kernel<T: ArrowPrimitiveType, S: ArrowPrimitiveType>(
    op: c_int, a: &PrimitiveArrayGpu<T>, scalar: &PrimitiveArrayGpu<S>, what: &str,
) -> PrimitiveArrayGpu<T> {
    assert_eq!(scalar.len, 1, "{what}: the scalar operand must have exactly one element");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_scalar(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), scalar.data.ptr_on(&a.gpu_device), out.data.ptr(), a.len,
                        a.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

/// out[i] = a[i] op b[i]; validity = AND of both (null_bit_buffer.rs:206-243) in the same kernel
pub fn binary_kernel<T: ArrowPrimitiveType, S: ArrowPrimitiveType>(
    op: c_int, a: &PrimitiveArrayGpu<T>, b: &PrimitiveArrayGpu<S>, what: &str,
) -> PrimitiveArrayGpu<T> {
    assert_eq!(a.len, b.len, "{what}: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), b.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_binary(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), b.data.ptr_on(&a.gpu_device), out.data.ptr(), a.len,
                        a.validity_ptr(), b.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

// ---------------------------------------------------------------------------------------------
// impls.  Reference matrix: scalar + - * / % for f32, u32 and the i32-backed types
// (Int32, Date32 in any combination: arithmetic/src/i32.rs:11-119), u16 + scalar; array + for
// f32 / u32 / i32-backed, f32 - * /.  BASELINE.json config 2 adds i8 / u8 / i16 / u16 (new surface).
// ---------------------------------------------------------------------------------------------
macro_rules! impl_scalar_same_type {
    ($trait:ident, $op_fn:ident, $id:expr, $($t:ty),*) => {$(
        impl $trait<PrimitiveArrayGpu<$t>> for PrimitiveArrayGpu<$t> {
            type Output = Self;
            fn $op_fn(&self, value: &PrimitiveArrayGpu<$t>, _pipeline: &mut ArrowComputePipeline) -> Self {
                scalar_kernel($id, self, value, stringify!($op_fn))
            }
        }
    )*};
}
// Self is concrete and only the right-hand side is generic (`Int32ArrayGPU + Date32ArrayGPU` and
// back), as in arithmetic/src/i32.rs:11-119: impls for different Self types never overlap
macro_rules! impl_scalar_i32_backed {
    ($trait:ident, $op_fn:ident, $id:expr) => {
        impl_scalar_i32_backed!(@one $trait, $op_fn, $id, i32);
        impl_scalar_i32_backed!(@one $trait, $op_fn, $id, Date32Type);
    };
    (@one $trait:ident, $op_fn:ident, $id:expr, $t:ty) => {
        impl<S: Int32Type + ArrowPrimitiveType> $trait<PrimitiveArrayGpu<S>> for PrimitiveArrayGpu<$t> {
            type Output = Self;
            fn $op_fn(&self, value: &PrimitiveArrayGpu<S>, _pipeline: &mut ArrowComputePipeline) -> Self {
                scalar_kernel($id, self, value, stringify!($op_fn))
            }
        }
    };
}
macro_rules! impl_array_same_type {
    ($trait:ident, $op_fn:ident, $id:expr, $($t:ty),*) => {$(
        impl $trait<PrimitiveArrayGpu<$t>> for PrimitiveArrayGpu<$t> {
            type Output = Self;
            fn $op_fn(&self, value: &PrimitiveArrayGpu<$t>, _pipeline: &mut ArrowComputePipeline) -> Self {
                binary_kernel($id, self, value, stringify!($op_fn))
            }
        }
    )*};
}
macro_rules! impl_array_i32_backed {
    ($trait:ident, $op_fn:ident, $id:expr) => {
        impl_array_i32_backed!(@one $trait, $op_fn, $id, i32);
        impl_array_i32_backed!(@one $trait, $op_fn, $id, Date32Type);
    };
    (@one $trait:ident, $op_fn:ident, $id:expr, $t:ty) => {
        impl<S: Int32Type + ArrowPrimitiveType> $trait<PrimitiveArrayGpu<S>> for PrimitiveArrayGpu<$t> {
            type Output = Self;
            fn $op_fn(&self, value: &PrimitiveArrayGpu<S>, _pipeline: &mut ArrowComputePipeline) -> Self {
                binary_kernel($id, self, value, stringify!($op_fn))
            }
        }
    };
}

impl_scalar_same_type!(ArrowScalarAdd, add_scalar_op, AGPU_ADD, f32, u32, u16, i16, i8, u8);
impl_scalar_same_type!(ArrowScalarSub, sub_scalar_op, AGPU_SUB, f32, u32, u16, i16, i8, u8);
impl_scalar_same_type!(ArrowScalarMul, mul_scalar_op, AGPU_MUL, f32, u32, u16, i16, i8, u8);
impl_scalar_same_type!(ArrowScalarDiv, div_scalar_op, AGPU_DIV, f32, u32, u16, i16, i8, u8);
impl_scalar_same_type!(ArrowScalarRem, rem_scalar_op, AGPU_REM, f32, u32, u16, i16, i8, u8);
impl_scalar_i32_backed!(ArrowScalarAdd, add_scalar_op, AGPU_ADD);
impl_scalar_i32_backed!(ArrowScalarSub, sub_scalar_op, AGPU_SUB);
impl_scalar_i32_backed!(ArrowScalarMul, mul_scalar_op, AGPU_MUL);
impl_scalar_i32_backed!(ArrowScalarDiv, div_scalar_op, AGPU_DIV);
impl_scalar_i32_backed!(ArrowScalarRem, rem_scalar_op, AGPU_REM);
impl_array_same_type!(ArrowAdd, add_op, AGPU_ADD, f32, u32, u16, i16, i8, u8);
impl_array_same_type!(ArrowSub, sub_op, AGPU_SUB, f32, u32, u16, i16, i8, u8);
impl_array_same_type!(ArrowMul, mul_op, AGPU_MUL, f32, u32, u16, i16, i8, u8);
impl_array_same_type!(ArrowDiv, div_op, AGPU_DIV, f32, u32, u16, i16, i8, u8);
impl_array_i32_backed!(ArrowAdd, add_op, AGPU_ADD);
impl_array_i32_backed!(ArrowSub, sub_op, AGPU_SUB);
impl_array_i32_backed!(ArrowMul, mul_op, AGPU_MUL);
impl_array_i32_backed!(ArrowDiv, div_op, AGPU_DIV);

impl<T: NegUnaryType + ArrowPrimitiveType> Neg for PrimitiveArrayGpu<T> {
    type OutputType = Self;
    fn neg_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, self.len, &[self.null_buffer.as_ref()]);
        let out = Self::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_unary(self.gpu_device.handle(), AGPU_NEG, T::DTYPE, self.values_ptr(), out.data.ptr(), self.len,
                           self.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "neg_op",
        );
        out
    }
}

impl<T: Sum32Bit> Sum for PrimitiveArrayGpu<T> {
    /// one-element array; validity is ignored like the reference; the f32 result follows the
    /// reference's 256-wide pairwise tree order bit for bit (aggregate.wgsl:28-41)
    fn sum_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        let out = Self::new_empty(&self.gpu_device, 1, None);
        check(unsafe { agpu_sum(self.gpu_device.handle(), T::DTYPE, self.values_ptr(), self.len, out.data.ptr()) }, "sum_op");
        out
    }
}

// ---------------------------------------------------------------------------------------------
// *_dyn dispatchers (arithmetic_kernels.rs:77-267, 322-343): runtime dtype match, `panic!` on
// unsupported pairs like the reference
// ---------------------------------------------------------------------------------------------
macro_rules! dyn_pair {
    ($(#[$doc:meta])* $dyn:ident, $op_dyn:ident, $method:ident, same: [$($same:ident),*], mixed: [$([$x:ident, $y:ident]),*]) => {
        $(#[$doc])*
        pub fn $dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU) -> ArrowArrayGPU {
            let mut pipeline = ArrowComputePipeline::new(data_1.get_gpu_device(), None);
            let result = $op_dyn(data_1, data_2, &mut pipeline);
            pipeline.finish();
            result
        }

        pub fn $op_dyn(data_1: &ArrowArrayGPU, data_2: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
            match (data_1, data_2) {
                $((ArrowArrayGPU::$same(a), ArrowArrayGPU::$same(b)) => a.$method(b, pipeline).into(),)*
                $((ArrowArrayGPU::$x(a), ArrowArrayGPU::$y(b)) => a.$method(b, pipeline).into(),)*
                _ => panic!("Operation {} not supported for type {:?} {:?}", stringify!($dyn), data_1.get_dtype(), data_2.get_dtype()),
            }
        }
    };
}

dyn_pair!(/// column + one-element column
          add_scalar_dyn, add_scalar_op_dyn, add_scalar_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// column - one-element column
          sub_scalar_dyn, sub_scalar_op_dyn, sub_scalar_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// column * one-element column
          mul_scalar_dyn, mul_scalar_op_dyn, mul_scalar_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// column / one-element column
          div_scalar_dyn, div_scalar_op_dyn, div_scalar_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// column % one-element column
          rem_scalar_dyn, rem_scalar_op_dyn, rem_scalar_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// x + y, row by row over both columns
          add_array_dyn, add_array_op_dyn, add_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, Date32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU],
          mixed: [[Int32ArrayGPU, Date32ArrayGPU], [Date32ArrayGPU, Int32ArrayGPU]]);
dyn_pair!(/// x - y, row by row over both columns
          sub_array_dyn, sub_array_op_dyn, sub_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU], mixed: []);
dyn_pair!(/// x * y, row by row over both columns
          mul_array_dyn, mul_array_op_dyn, mul_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU], mixed: []);
dyn_pair!(/// x / y, row by row over both columns
          div_array_dyn, div_array_op_dyn, div_op,
          same: [Float32ArrayGPU, Int32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, Int16ArrayGPU, Int8ArrayGPU, UInt8ArrayGPU], mixed: []);

/// add_dyn & co route by operand length (arithmetic_kernels.rs:101-119): both of length 1 or both
/// longer -> array op; exactly one of length 1 -> scalar op with that operand as the scalar
macro_rules! dyn_by_length {
    ($([$dyn:ident, $op_dyn:ident, $array_op:ident, $scalar_op:ident]),*) => {$(
        pub fn $dyn(input1: &ArrowArrayGPU, input2: &ArrowArrayGPU) -> ArrowArrayGPU {
            let mut pipeline = ArrowComputePipeline::new(input1.get_gpu_device(), None);
            let result = $op_dyn(input1, input2, &mut pipeline);
            pipeline.finish();
            result
        }

        pub fn $op_dyn(input1: &ArrowArrayGPU, input2: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
            match (input1.len() == 1, input2.len() == 1) {
                (true, true) | (false, false) => $array_op(input1, input2, pipeline),
                (false, true) => $scalar_op(input1, input2, pipeline),
                (true, false) => $scalar_op(input2, input1, pipeline),
            }
        }
    )*};
}
dyn_by_length!([add_dyn, add_op_dyn, add_array_op_dyn, add_scalar_op_dyn], [sub_dyn, sub_op_dyn, sub_array_op_dyn, sub_scalar_op_dyn],
               [mul_dyn, mul_op_dyn, mul_array_op_dyn, mul_scalar_op_dyn], [div_dyn, div_op_dyn, div_array_op_dyn, div_scalar_op_dyn]);

/// arithmetic_kernels.rs:322-343
pub fn neg_dyn(data: &ArrowArrayGPU) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(data.get_gpu_device(), None);
    let result = neg_op_dyn(data, &mut pipeline);
    pipeline.finish();
    result
}

pub fn neg_op_dyn(data: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    match data {
        ArrowArrayGPU::Float32ArrayGPU(x) => x.neg_op(pipeline).into(),
        _ => panic!("Operation neg_dyn not supported for type {:?}", data.get_dtype()),
    }
}

// ---------------------------------------------------------------------------------------------
// new surface (BASELINE.json config 3): the recorded chain  gt_op(add_op(mul_op(a, b), c), d)  of
// crates/arrow/examples/simple.rs:45-72 as ONE kernel, bit-identical to the three ops (two
// roundings, no FMA contraction); validity = AND of the four bitmaps
// ---------------------------------------------------------------------------------------------
pub fn fused_mul_add_gt_op(
    a: &Float32ArrayGPU, b: &Float32ArrayGPU, c: &Float32ArrayGPU, d: &Float32ArrayGPU, _pipeline: &mut ArrowComputePipeline,
) -> BooleanArrayGPU {
    assert!(a.len == b.len && a.len == c.len && a.len == d.len, "fused_mul_add_gt_op: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len,
                                          &[a.null_buffer.as_ref(), b.null_buffer.as_ref(), c.null_buffer.as_ref(), d.null_buffer.as_ref()]);
    let out = BooleanArrayGPU::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_fused_mul_add_gt(a.gpu_device.handle(), a.values_ptr() as *const f32, b.values_ptr() as *const f32,
                                  c.values_ptr() as *const f32, d.values_ptr() as *const f32, out.data.ptr() as *mut u32, a.len,
                                  a.validity_ptr(), b.validity_ptr(), c.validity_ptr(), d.validity_ptr(),
                                  NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        "fused_mul_add_gt_op",
    );
    out
}

// ---------------------------------------------------------------------------------------------
// new surface (BASELINE.json config 1): the reference's first benchmark program
//     let s = a.add_op(&b, p);  let g = a.gt_op(&b, p);
// as ONE kernel: `value = a binop b`, `predicate = a cmpop c` (c is usually b) read every column
// once (12.5 B/row instead of 20.875) and are one launch instead of two; both results are
// bit-identical to the separate ops.  binop: AGPU_ADD / SUB / MUL / DIV (dedicated kernel) or
// AGPU_REM / MIN / MAX (chain interpreter); cmpop: AGPU_GT .. AGPU_EQ.  The kernel writes ONE
// validity bitmap (AND of a, b, c), so b and c must either both carry one or both carry none;
// the predicate's bitmap is a device-side copy of it (buffers are uniquely owned here).
// ---------------------------------------------------------------------------------------------
pub fn fused_binary_compare_op(
    a: &Float32ArrayGPU, binop: c_int, b: &Float32ArrayGPU, cmpop: c_int, c: &Float32ArrayGPU, _pipeline: &mut ArrowComputePipeline,
) -> (Float32ArrayGPU, BooleanArrayGPU) {
    assert!(a.len == b.len && a.len == c.len, "fused_binary_compare_op: length mismatch");
    let same_column = std::ptr::eq(b.values_ptr(), c.values_ptr());
    assert!(same_column || b.null_buffer.is_some() == c.null_buffer.is_some(),
            "fused_binary_compare_op: value and predicate would depend on different validity bitmaps");
    let step = |kind: c_int, op: c_int, col: Option<&Float32ArrayGPU>| AgpuChainStep {
        kind, op,
        operand: col.map_or(std::ptr::null(), |x| x.values_ptr()),
        validity: col.map_or(std::ptr::null(), |x| x.validity_ptr()),
        scalar: 0.0,
    };
    let steps = [step(AGPU_STEP_BINARY_COLUMN, binop, Some(b)), step(AGPU_STEP_STORE, 0, None), step(AGPU_STEP_RESET, 0, None),
                 step(AGPU_STEP_COMPARE_COLUMN, cmpop, Some(c))];
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), b.null_buffer.as_ref(), c.null_buffer.as_ref()]);
    let value = Float32ArrayGPU::new_empty(&a.gpu_device, a.len, nb);
    let bits = BooleanArrayGPU::new_empty(&a.gpu_device, a.len, None);
    let rc = unsafe {
        agpu_fused_chain_pair(a.gpu_device.handle(), AGPU_F32, a.values_ptr(), a.validity_ptr(), steps.as_ptr(), steps.len() as c_int,
                              value.data.ptr() as *mut f32, bits.data.ptr() as *mut u32, a.len,
                              NullBitBufferGpu::words_mut(value.null_buffer.as_ref()))
    };
    if rc == AGPU_EUNSUPPORTED || rc == AGPU_EINVAL {
        panic!("fused_binary_compare_op: unsupported operator pair {binop} / {cmpop}");
    }
    check(rc, "fused_binary_compare_op");
    let predicate = BooleanArrayGPU { null_buffer: NullBitBufferGpu::clone_null_bit_buffer(&value.null_buffer), ..bits };
    (value, predicate)
}


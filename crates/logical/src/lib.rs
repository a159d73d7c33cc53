//! arrow_gpu_logical — bitwise `and or xor not shl shr` on integer columns and on
//! `BooleanArrayGPU`, `any` / `all` (drop-in for crates/logical, lib.rs:44-86, boolean.rs).
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

/// Marker of the element types that support logical operations
pub trait LogicalType {}
macro_rules! mark { ($($t:ty),*) => { $(impl LogicalType for $t {})* }; }
mark!(u32, u16, u8, i32, i16, i8);

/// Bitwise logic and per-row shifts, row by row (logical/src/lib.rs:44-86 of the reference)
pub trait Logical: ArrayUtils + Sized {
    fn bitwise_and(&self, operand: &Self) -> Self {
        eager!(self, bitwise_and_op, operand)
    }
    fn bitwise_or(&self, operand: &Self) -> Self {
        eager!(self, bitwise_or_op, operand)
    }
    fn bitwise_xor(&self, operand: &Self) -> Self {
        eager!(self, bitwise_xor_op, operand)
    }
    fn bitwise_not(&self) -> Self {
        eager!(self, bitwise_not_op)
    }
    fn bitwise_shl(&self, operand: &UInt32ArrayGPU) -> Self {
        eager!(self, bitwise_shl_op, operand)
    }
    fn bitwise_shr(&self, operand: &UInt32ArrayGPU) -> Self {
        eager!(self, bitwise_shr_op, operand)
    }
    fn bitwise_and_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> Self;
    fn bitwise_or_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> Self;
    fn bitwise_xor_op(&self, operand: &Self, pipeline: &mut ArrowComputePipeline) -> Self;
    fn bitwise_not_op(&self, pipeline: &mut ArrowComputePipeline) -> Self;
    fn bitwise_shl_op(&self, operand: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self;
    fn bitwise_shr_op(&self, operand: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self;
}

/// Trait for any / all over a boolean array (validity is ignored, like the reference)
pub trait LogicalContains {
    fn any(&self) -> bool;
    fn all(&self) -> bool;
}

fn word_op<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, b: &PrimitiveArrayGpu<T>, what: &str) -> PrimitiveArrayGpu<T> {
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

/// one u32 count PER ROW; widen, shift by `count & 31`, truncate (logical/src/lib.rs:160-186)
fn shift_op<T: ArrowPrimitiveType>(op: c_int, a: &PrimitiveArrayGpu<T>, counts: &UInt32ArrayGPU, what: &str) -> PrimitiveArrayGpu<T> {
    assert_eq!(a.len, counts.len, "{what}: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), counts.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_shift(a.gpu_device.handle(), op, T::DTYPE, a.values_ptr(), counts.values_ptr() as *const u32, out.data.ptr(), a.len,
                       a.validity_ptr(), counts.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

impl<T: LogicalType + ArrowPrimitiveType> Logical for PrimitiveArrayGpu<T> {
    fn bitwise_and_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        word_op(AGPU_AND, self, operand, "bitwise_and_op")
    }
    fn bitwise_or_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        word_op(AGPU_OR, self, operand, "bitwise_or_op")
    }
    fn bitwise_xor_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        word_op(AGPU_XOR, self, operand, "bitwise_xor_op")
    }
    fn bitwise_not_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, self.len, &[self.null_buffer.as_ref()]);
        let out = Self::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_unary(self.gpu_device.handle(), AGPU_NOT, T::DTYPE, self.values_ptr(), out.data.ptr(), self.len,
                           self.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "bitwise_not_op",
        );
        out
    }
    fn bitwise_shl_op(&self, operand: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        shift_op(AGPU_SHL, self, operand, "bitwise_shl_op")
    }
    fn bitwise_shr_op(&self, operand: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        shift_op(AGPU_SHR, self, operand, "bitwise_shr_op")
    }
}

fn bitmap_op(op: c_int, a: &BooleanArrayGPU, b: &BooleanArrayGPU, what: &str) -> BooleanArrayGPU {
    assert_eq!(a.len, b.len, "{what}: length mismatch");
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref(), b.null_buffer.as_ref()]);
    let out = BooleanArrayGPU::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_bitmap_binary(a.gpu_device.handle(), op, a.bits_ptr(), b.bits_ptr(), out.data.ptr() as *mut u32, a.len,
                               a.validity_ptr(), b.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

/// logical/src/boolean.rs:45-104 (shifts of a boolean array are an empty shader there: panic here)
impl Logical for BooleanArrayGPU {
    fn bitwise_and_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        bitmap_op(AGPU_AND, self, operand, "bitwise_and_op")
    }
    fn bitwise_or_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        bitmap_op(AGPU_OR, self, operand, "bitwise_or_op")
    }
    fn bitwise_xor_op(&self, operand: &Self, _pipeline: &mut ArrowComputePipeline) -> Self {
        bitmap_op(AGPU_XOR, self, operand, "bitwise_xor_op")
    }
    fn bitwise_not_op(&self, _pipeline: &mut ArrowComputePipeline) -> Self {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, self.len, &[self.null_buffer.as_ref()]);
        let out = BooleanArrayGPU::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_bitmap_not(self.gpu_device.handle(), self.bits_ptr(), out.data.ptr() as *mut u32, self.len, self.validity_ptr(),
                                NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "bitwise_not_op",
        );
        out
    }
    fn bitwise_shl_op(&self, _operand: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        panic!("bitwise_shl is not defined for BooleanArrayGPU")
    }
    fn bitwise_shr_op(&self, _operand: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        panic!("bitwise_shr is not defined for BooleanArrayGPU")
    }
}

/// logical/src/boolean.rs:106-147; `all` counts only the first `len` bits
impl LogicalContains for BooleanArrayGPU {
    fn any(&self) -> bool {
        reduce_flag(self, true)
    }
    fn all(&self) -> bool {
        reduce_flag(self, false)
    }
}

fn reduce_flag(bits: &BooleanArrayGPU, any: bool) -> bool {
    let flag = bits.gpu_device.create_empty_buffer(4);
    let rc = unsafe {
        if any {
            agpu_any(bits.gpu_device.handle(), bits.bits_ptr(), bits.len, flag.ptr() as *mut u32)
        } else {
            agpu_all(bits.gpu_device.handle(), bits.bits_ptr(), bits.len, flag.ptr() as *mut u32)
        }
    };
    check(rc, if any { "any" } else { "all" });
    bits.gpu_device.retrive_data(&flag)[..4] != [0, 0, 0, 0]
}

/// logical/src/lib.rs:189-349
macro_rules! dyn_binary {
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
                (UInt32ArrayGPU(a), UInt32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (UInt16ArrayGPU(a), UInt16ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (UInt8ArrayGPU(a), UInt8ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int32ArrayGPU(a), Int32ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int16ArrayGPU(a), Int16ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (Int8ArrayGPU(a), Int8ArrayGPU(b)) => a.$method(b, pipeline).into(),
                (BooleanArrayGPU(a), BooleanArrayGPU(b)) => a.$method(b, pipeline).into(),
                _ => panic!("Operation {} not supported for type {:?} {:?}", stringify!($dyn), data_1.get_dtype(), data_2.get_dtype()),
            }
        }
    )*};
}
dyn_binary!(
    [/// x & y, row by row over both columns
     bitwise_and_dyn, bitwise_and_op_dyn, bitwise_and_op],
    [/// x | y, row by row over both columns
     bitwise_or_dyn, bitwise_or_op_dyn, bitwise_or_op],
    [/// x ^ y, row by row over both columns
     bitwise_xor_dyn, bitwise_xor_op_dyn, bitwise_xor_op]
);

macro_rules! dyn_shift {
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
                (UInt32ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                (UInt16ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                (UInt8ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                (Int32ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                (Int16ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                (Int8ArrayGPU(a), UInt32ArrayGPU(c)) => a.$method(c, pipeline).into(),
                _ => panic!("Operation {} not supported for type {:?} {:?}", stringify!($dyn), data_1.get_dtype(), data_2.get_dtype()),
            }
        }
    )*};
}
dyn_shift!(
    [/// x << y, row by row over both columns
     bitwise_shl_dyn, bitwise_shl_op_dyn, bitwise_shl_op],
    [/// x >> y, row by row over both columns
     bitwise_shr_dyn, bitwise_shr_op_dyn, bitwise_shr_op]
);

/// `!x` for every row
pub fn bitwise_not_dyn(data: &ArrowArrayGPU) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(data.get_gpu_device(), None);
    let result = bitwise_not_op_dyn(data, &mut pipeline);
    pipeline.finish();
    result
}

pub fn bitwise_not_op_dyn(data: &ArrowArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    use ArrowArrayGPU::*;
    match data {
        UInt32ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        UInt16ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        UInt8ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        Int32ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        Int16ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        Int8ArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        BooleanArrayGPU(a) => a.bitwise_not_op(pipeline).into(),
        _ => panic!("Operation bitwise_not_dyn not supported for type {:?}", data.get_dtype()),
    }
}

//! arrow_gpu_cast — `Cast<T>` / `BitCast<T>` and their `*_dyn` (drop-in for crates/cast,
//! lib.rs:15-38, matrix :135-161).  One `agpu_cast(src dtype, dst dtype)` call per cast; the
//! validity bitmap is copied by the same kernel.
use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

/// Trait for casting each element of the array to type `T`
pub trait Cast<T>: ArrayUtils {
    fn cast(&self) -> T {
        let mut pipeline = ArrowComputePipeline::new(self.get_gpu_device(), None);
        let output = self.cast_op(&mut pipeline);
        pipeline.finish();
        output
    }
    fn cast_op(&self, pipeline: &mut ArrowComputePipeline) -> T;
}

/// Trait for reinterpreting the bits of each element as type `T`
pub trait BitCast<T>: ArrayUtils {
    fn bitcast(&self) -> T {
        let mut pipeline = ArrowComputePipeline::new(self.get_gpu_device(), None);
        let output = self.bitcast_op(&mut pipeline);
        pipeline.finish();
        output
    }
    fn bitcast_op(&self, pipeline: &mut ArrowComputePipeline) -> T;
}

fn cast_kernel<S: ArrowPrimitiveType, D: ArrowPrimitiveType>(a: &PrimitiveArrayGpu<S>, what: &str) -> PrimitiveArrayGpu<D> {
    let nb = NullBitBufferGpu::for_output(&a.gpu_device, a.len, &[a.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<D>::new_empty(&a.gpu_device, a.len, nb);
    check(
        unsafe {
            agpu_cast(a.gpu_device.handle(), S::DTYPE, D::DTYPE, a.values_ptr(), out.data.ptr(), a.len, a.validity_ptr(),
                      NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
        },
        what,
    );
    out
}

/// the matrix of cast/src/lib.rs:40-87 (i8_cast.rs, i16_cast.rs, u8_cast.rs, u16_cast.rs, f32_cast.rs)
macro_rules! impl_cast {
    ($([$from:ty => $($into:ty),+]),*) => {$($(
        impl Cast<PrimitiveArrayGpu<$into>> for PrimitiveArrayGpu<$from> {
            fn cast_op(&self, _pipeline: &mut ArrowComputePipeline) -> PrimitiveArrayGpu<$into> {
                cast_kernel::<$from, $into>(self, "cast_op")
            }
        }
    )+)*};
}
impl_cast!([i8 => u8, u16, u32, i16, i32, f32], [i16 => i32, u16, u32, f32], [u8 => u16, u32, i8, i16, i32, f32],
           [u16 => u32, i16, i32, f32], [f32 => u8]);

/// cast/src/boolean_cast.rs: bit set -> 1.0 else 0.0
impl Cast<Float32ArrayGPU> for BooleanArrayGPU {
    fn cast_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, self.len, &[self.null_buffer.as_ref()]);
        let out = Float32ArrayGPU::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_cast(self.gpu_device.handle(), AGPU_BOOL, AGPU_F32, self.bits_ptr() as *const _, out.data.ptr(), self.len,
                          self.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "cast_op",
        );
        out
    }
}

/// cast/src/lib.rs:90-108, u32_cast.rs:5 — the one bitcast the reference has
impl BitCast<Float32ArrayGPU> for UInt32ArrayGPU {
    fn bitcast_op(&self, _pipeline: &mut ArrowComputePipeline) -> Float32ArrayGPU {
        cast_kernel::<u32, f32>(self, "bitcast_op")
    }
}

/// every row converted to the element type `T` (cast/src/lib.rs:111-161)
pub fn cast_dyn(from: &ArrowArrayGPU, into: &ArrowType) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(from.get_gpu_device(), None);
    let result = cast_op_dyn(from, into, &mut pipeline);
    pipeline.finish();
    result
}

pub fn cast_op_dyn(from: &ArrowArrayGPU, into: &ArrowType, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    macro_rules! route {
        ($x:ident, $($ty:ident => $arr:ident),+) => {
            match into {
                $(ArrowType::$ty => Cast::<$arr>::cast_op($x, pipeline).into(),)+
                _ => panic!("Casting not supported for type {:?} {:?}", from.get_dtype(), into),
            }
        };
    }
    match from {
        ArrowArrayGPU::Int8ArrayGPU(x) => route!(x, UInt8Type => UInt8ArrayGPU, UInt16Type => UInt16ArrayGPU, UInt32Type => UInt32ArrayGPU,
                                                 Int16Type => Int16ArrayGPU, Int32Type => Int32ArrayGPU, Float32Type => Float32ArrayGPU),
        ArrowArrayGPU::Int16ArrayGPU(x) => route!(x, Int32Type => Int32ArrayGPU, UInt16Type => UInt16ArrayGPU, UInt32Type => UInt32ArrayGPU,
                                                  Float32Type => Float32ArrayGPU),
        ArrowArrayGPU::UInt8ArrayGPU(x) => route!(x, UInt16Type => UInt16ArrayGPU, UInt32Type => UInt32ArrayGPU, Int8Type => Int8ArrayGPU,
                                                  Int16Type => Int16ArrayGPU, Int32Type => Int32ArrayGPU, Float32Type => Float32ArrayGPU),
        ArrowArrayGPU::UInt16ArrayGPU(x) => route!(x, UInt32Type => UInt32ArrayGPU, Int16Type => Int16ArrayGPU, Int32Type => Int32ArrayGPU,
                                                   Float32Type => Float32ArrayGPU),
        ArrowArrayGPU::Float32ArrayGPU(x) => route!(x, UInt8Type => UInt8ArrayGPU),
        ArrowArrayGPU::BooleanArrayGPU(x) => route!(x, Float32Type => Float32ArrayGPU),
        _ => panic!("Casting not supported for type {:?} {:?}", from.get_dtype(), into),
    }
}

/// the same bits of every row read as `T` (cast/src/lib.rs:163-192)
pub fn bitcast_dyn(from: &ArrowArrayGPU, into: &ArrowType) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(from.get_gpu_device(), None);
    let result = bitcast_op_dyn(from, into, &mut pipeline);
    pipeline.finish();
    result
}

pub fn bitcast_op_dyn(from: &ArrowArrayGPU, into: &ArrowType, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    match (from, into) {
        (ArrowArrayGPU::UInt32ArrayGPU(x), ArrowType::Float32Type) => BitCast::<Float32ArrayGPU>::bitcast_op(x, pipeline).into(),
        _ => panic!("Casting not supported for type {:?} {:?}", from.get_dtype(), into),
    }
}

//! arrow_gpu_routines — `Swizzle::{merge, take, put}` (drop-in for crates/routines, lib.rs:28-72)
//! plus `filter` (order-preserving compaction: named by BASELINE.json config 5, absent from the
//! reference).  merge selects values AND builds the validity `((va & m) | (vb & !m)) & vmask` in
//! one kernel (the reference: 1 + up to 4 dispatches, merge.rs:17-86); take gathers values and
//! validity bits in one pass over the indexes; put is bounds-checked like wgpu's robust buffer access.
use arrow_gpu_array::array::*;
use arrow_gpu_array::gpu_utils::ffi::*;
use arrow_gpu_array::gpu_utils::ArrowComputePipeline;

/// Row re-arrangement: merge, take, put — and filter (routines/src/lib.rs:28-72 of the reference)
pub trait Swizzle: ArrayUtils + Sized {
    fn merge(&self, other: &Self, mask: &BooleanArrayGPU) -> Self {
        let mut pipeline = ArrowComputePipeline::new(self.get_gpu_device(), None);
        let result = self.merge_op(other, mask, &mut pipeline);
        pipeline.finish();
        result
    }

    fn take(&self, indexes: &UInt32ArrayGPU) -> Self {
        let mut pipeline = ArrowComputePipeline::new(self.get_gpu_device(), None);
        let result = self.take_op(indexes, &mut pipeline);
        pipeline.finish();
        result
    }

    fn put(&self, src_indexes: &UInt32ArrayGPU, dst: &mut Self, dst_indexes: &UInt32ArrayGPU) {
        let mut pipeline = ArrowComputePipeline::new(self.get_gpu_device(), None);
        self.put_op(src_indexes, dst, dst_indexes, &mut pipeline);
        pipeline.finish();
    }

    /// Elements of self where the mask bit is set, else of other; None in mask results in None
    fn merge_op(&self, other: &Self, mask: &BooleanArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self;

    /// gathers `self[indexes[i]]` into a new column of `indexes.len` rows
    fn take_op(&self, indexes: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self;

    /// Put elements from self using src_indexes into dst using dst_indexes
    fn put_op(&self, src_indexes: &UInt32ArrayGPU, dst: &mut Self, dst_indexes: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline);
}

/// Marker of the element types that support swizzle operations (8/16-bit take and put are
/// `todo!()` in the reference — routines/src/i16.rs:5-6 — and work here)
pub trait SwizzleType {}
macro_rules! mark { ($($t:ty),*) => { $(impl SwizzleType for $t {})* }; }
mark!(f32, u32, u16, u8, i32, i16, i8, Date32Type);

impl<T: SwizzleType + ArrowPrimitiveType> Swizzle for PrimitiveArrayGpu<T> {
    fn merge_op(&self, other: &Self, mask: &BooleanArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self {
        assert!(self.len == other.len && self.len == mask.len, "merge_op: length mismatch");
        let nb = merge_null_buffers_op(&self.null_buffer, &other.null_buffer, mask, pipeline);
        let out = Self::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_merge(self.gpu_device.handle(), T::DTYPE, self.values_ptr(), other.values_ptr(), mask.bits_ptr(), out.data.ptr(), self.len,
                           self.validity_ptr(), other.validity_ptr(), mask.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "merge_op",
        );
        out
    }

    /// result length = indexes.len; an index past the end reads zero (robust buffer access)
    fn take_op(&self, indexes: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, indexes.len, &[self.null_buffer.as_ref()]);
        let out = Self::new_empty(&self.gpu_device, indexes.len, nb);
        check(
            unsafe {
                agpu_take(self.gpu_device.handle(), T::DTYPE, self.values_ptr(), self.len, indexes.values_ptr() as *const u32, out.data.ptr(),
                          indexes.len, self.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "take_op",
        );
        out
    }

    fn put_op(&self, src_indexes: &UInt32ArrayGPU, dst: &mut Self, dst_indexes: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) {
        assert_eq!(src_indexes.len, dst_indexes.len, "put_op: index arrays differ in length");
        if self.null_buffer.is_some() || dst.null_buffer.is_some() {
            todo!("put with validity bitmaps (todo!() in the reference as well: routines/src/lib.rs:164-169)")
        }
        check(
            unsafe {
                agpu_put(self.gpu_device.handle(), T::DTYPE, self.values_ptr(), self.len, src_indexes.values_ptr() as *const u32,
                         dst.data.ptr(), dst.len, dst_indexes.values_ptr() as *const u32, src_indexes.len)
            },
            "put_op",
        );
    }
}

/// routines/src/bool.rs:48-128
impl Swizzle for BooleanArrayGPU {
    fn merge_op(&self, other: &Self, mask: &BooleanArrayGPU, pipeline: &mut ArrowComputePipeline) -> Self {
        assert!(self.len == other.len && self.len == mask.len, "merge_op: length mismatch");
        let nb = merge_null_buffers_op(&self.null_buffer, &other.null_buffer, mask, pipeline);
        let out = BooleanArrayGPU::new_empty(&self.gpu_device, self.len, nb);
        check(
            unsafe {
                agpu_merge(self.gpu_device.handle(), AGPU_BOOL, self.bits_ptr() as *const _, other.bits_ptr() as *const _, mask.bits_ptr(),
                           out.data.ptr(), self.len, self.validity_ptr(), other.validity_ptr(), mask.validity_ptr(),
                           NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "merge_op",
        );
        out
    }

    fn take_op(&self, indexes: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) -> Self {
        let nb = NullBitBufferGpu::for_output(&self.gpu_device, indexes.len, &[self.null_buffer.as_ref()]);
        let out = BooleanArrayGPU::new_empty(&self.gpu_device, indexes.len, nb);
        check(
            unsafe {
                agpu_take(self.gpu_device.handle(), AGPU_BOOL, self.bits_ptr() as *const _, self.len, indexes.values_ptr() as *const u32,
                          out.data.ptr(), indexes.len, self.validity_ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()))
            },
            "take_op",
        );
        out
    }

    fn put_op(&self, src_indexes: &UInt32ArrayGPU, dst: &mut Self, dst_indexes: &UInt32ArrayGPU, _pipeline: &mut ArrowComputePipeline) {
        assert_eq!(src_indexes.len, dst_indexes.len, "put_op: index arrays differ in length");
        check(
            unsafe {
                agpu_put(self.gpu_device.handle(), AGPU_BOOL, self.bits_ptr() as *const _, self.len, src_indexes.values_ptr() as *const u32,
                         dst.data.ptr(), dst.len, dst_indexes.values_ptr() as *const u32, src_indexes.len)
            },
            "put_op",
        );
    }
}

/// routines/src/merge.rs:17-86.  The reference computes the merged validity here with up to four
/// dispatches; `agpu_merge` does it inside the value kernel, so this only ALLOCATES the output
/// bitmap (iff any input has one; a missing bitmap counts as all ones).
pub fn merge_null_buffers_op(
    null_buffer_1: &Option<NullBitBufferGpu>, null_buffer_2: &Option<NullBitBufferGpu>, mask: &BooleanArrayGPU,
    _pipeline: &mut ArrowComputePipeline,
) -> Option<NullBitBufferGpu> {
    NullBitBufferGpu::for_output(&mask.gpu_device, mask.len, &[null_buffer_1.as_ref(), null_buffer_2.as_ref(), mask.null_buffer.as_ref()])
}

/// New surface: keep the rows whose mask bit is set and valid, order preserving.  Two C calls so a
/// sharded caller can exchange the per-shard counts between them (`agpu_exchange_post/wait`).
pub fn filter_op<T: SwizzleType + ArrowPrimitiveType>(data: &PrimitiveArrayGpu<T>, mask: &BooleanArrayGPU, _pipeline: &mut ArrowComputePipeline) -> PrimitiveArrayGpu<T> {
    assert_eq!(data.len, mask.len, "filter_op: length mismatch");
    let dev = &data.gpu_device;
    let scratch = dev.create_empty_buffer(unsafe { agpu_filter_scratch_bytes(data.len) } as u64);
    let total = dev.create_empty_buffer(8);
    check(unsafe { agpu_filter_count(dev.handle(), mask.bits_ptr(), mask.validity_ptr(), data.len, scratch.ptr(), total.ptr() as *mut u64) }, "filter_count");
    let count = u64::from_le_bytes(dev.retrive_data(&total)[..8].try_into().unwrap()) as usize; // the op's one host synchronisation
    let nb = NullBitBufferGpu::for_output(dev, count, &[data.null_buffer.as_ref()]);
    let out = PrimitiveArrayGpu::<T>::new_empty(dev, count, nb);
    check(
        unsafe {
            agpu_filter_scatter(dev.handle(), T::DTYPE, data.values_ptr(), data.validity_ptr(), mask.bits_ptr(), mask.validity_ptr(), data.len,
                                scratch.ptr(), out.data.ptr(), NullBitBufferGpu::words_mut(out.null_buffer.as_ref()), count)
        },
        "filter_scatter",
    );
    out
}

/// routines/src/merge.rs:92-143, take.rs:58-95, put.rs:59-108
macro_rules! route_same {
    ($a:expr, $b:expr, |$x:ident, $y:ident| $body:expr, $what:literal) => {{
        use ArrowArrayGPU::*;
        match ($a, $b) {
            (Float32ArrayGPU($x), Float32ArrayGPU($y)) => $body,
            (UInt32ArrayGPU($x), UInt32ArrayGPU($y)) => $body,
            (UInt16ArrayGPU($x), UInt16ArrayGPU($y)) => $body,
            (UInt8ArrayGPU($x), UInt8ArrayGPU($y)) => $body,
            (Int32ArrayGPU($x), Int32ArrayGPU($y)) => $body,
            (Int16ArrayGPU($x), Int16ArrayGPU($y)) => $body,
            (Int8ArrayGPU($x), Int8ArrayGPU($y)) => $body,
            (Date32ArrayGPU($x), Date32ArrayGPU($y)) => $body,
            (BooleanArrayGPU($x), BooleanArrayGPU($y)) => $body,
            (a, b) => panic!(concat!($what, " Operation not supported between {:?} and {:?}"), a.get_dtype(), b.get_dtype()),
        }
    }};
}

pub fn merge_dyn(operand_1: &ArrowArrayGPU, operand_2: &ArrowArrayGPU, mask: &BooleanArrayGPU) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(operand_1.get_gpu_device(), Some("merge"));
    let result = merge_op_dyn(operand_1, operand_2, mask, &mut pipeline);
    pipeline.finish();
    result
}

pub fn merge_op_dyn(operand_1: &ArrowArrayGPU, operand_2: &ArrowArrayGPU, mask: &BooleanArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    route_same!(operand_1, operand_2, |x, y| x.merge_op(y, mask, pipeline).into(), "Merge")
}

pub fn take_dyn(operand_1: &ArrowArrayGPU, indexes: &UInt32ArrayGPU) -> ArrowArrayGPU {
    let mut pipeline = ArrowComputePipeline::new(operand_1.get_gpu_device(), Some("take"));
    let result = take_op_dyn(operand_1, indexes, &mut pipeline);
    pipeline.finish();
    result
}

pub fn take_op_dyn(operand_1: &ArrowArrayGPU, indexes: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline) -> ArrowArrayGPU {
    arrow_gpu_array::for_each_array!(operand_1, x => x.take_op(indexes, pipeline).into())
}

pub fn put_dyn(src: &ArrowArrayGPU, src_indexes: &UInt32ArrayGPU, dst: &mut ArrowArrayGPU, dst_indexes: &UInt32ArrayGPU) {
    let mut pipeline = ArrowComputePipeline::new(src.get_gpu_device(), Some("put"));
    put_op_dyn(src, src_indexes, dst, dst_indexes, &mut pipeline);
    pipeline.finish();
}

pub fn put_op_dyn(src: &ArrowArrayGPU, src_indexes: &UInt32ArrayGPU, dst: &mut ArrowArrayGPU, dst_indexes: &UInt32ArrayGPU, pipeline: &mut ArrowComputePipeline) {
    route_same!(src, dst, |x, y| x.put_op(src_indexes, y, dst_indexes, pipeline), "Put")
}

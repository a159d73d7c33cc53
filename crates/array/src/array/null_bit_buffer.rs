//! Validity bitmaps (array/null_bit_buffer.rs): `BooleanBufferBuilder` on the host,
//! `NullBitBufferGpu` on the device.  LSB-first bits, 1 = valid, `None` = all valid.
use std::sync::Arc;

use super::buffer::ArrowGpuBuffer;
use crate::gpu_utils::ffi::*;
use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};

/// null_bit_buffer.rs:10-62 (padded to whole u32 words: what the kernels read)
#[derive(Debug)]
pub struct BooleanBufferBuilder {
    pub data: Vec<u8>,
    pub len: usize,
    pub contains_nulls: bool,
}

impl BooleanBufferBuilder {
    pub fn new_with_capacity(size: usize) -> Self {
        Self { data: vec![0; size.div_ceil(32) * 4], len: size, contains_nulls: true }
    }

    pub fn new_set_with_capacity(size: usize) -> Self {
        let mut b = Self::new_with_capacity(size);
        for i in 0..size {
            b.set_bit(i);
        }
        b.contains_nulls = false;
        b
    }

    pub fn set_bit(&mut self, pos: usize) {
        self.data[pos / 8] |= 1 << (pos % 8);
    }

    pub fn unset_bit(&mut self, pos: usize) {
        self.data[pos / 8] &= !(1 << (pos % 8));
    }

    pub fn is_set(&self, pos: usize) -> bool {
        Self::is_set_in_slice(&self.data, pos)
    }

    pub fn is_set_in_slice(data: &[u8], pos: usize) -> bool {
        data[pos / 8] & (1 << (pos % 8)) != 0
    }
}

/// null_bit_buffer.rs:91-96
#[derive(Debug)]
pub struct NullBitBufferGpu {
    pub bit_buffer: ArrowGpuBuffer,
    pub len: usize,
    pub gpu_device: Arc<GpuDevice>,
}

impl NullBitBufferGpu {
    /// null_bit_buffer.rs:99-111: `None` when the builder saw no null
    pub fn new(gpu_device: Arc<GpuDevice>, buffer_builder: &BooleanBufferBuilder) -> Option<Self> {
        if !buffer_builder.contains_nulls {
            return None;
        }
        let bit_buffer = gpu_device.create_gpu_buffer_with_data(&buffer_builder.data);
        Some(Self { bit_buffer, len: buffer_builder.len, gpu_device })
    }

    pub fn new_set_with_capacity(gpu_device: Arc<GpuDevice>, size: usize) -> Self {
        let mut b = BooleanBufferBuilder::new_set_with_capacity(size);
        b.contains_nulls = true;
        Self::new(gpu_device, &b).unwrap()
    }

    /// a fresh, uninitialised bitmap for `len` rows (an op's kernel writes every word of it)
    pub fn new_empty(gpu_device: &Arc<GpuDevice>, len: usize) -> Self {
        Self { bit_buffer: gpu_device.create_empty_buffer((len.div_ceil(32) * 4) as u64), len, gpu_device: gpu_device.clone() }
    }

    /// null_bit_buffer.rs:124-128
    pub fn raw_values(&self) -> Vec<u8> {
        let mut v = self.gpu_device.retrive_data(&self.bit_buffer);
        v.truncate(self.len.div_ceil(8));
        v
    }

    pub fn words(data: Option<&Self>) -> *const u32 {
        data.map_or(std::ptr::null(), |x| x.bit_buffer.ptr() as *const u32)
    }

    pub fn words_mut(data: Option<&Self>) -> *mut u32 {
        data.map_or(std::ptr::null_mut(), |x| x.bit_buffer.ptr() as *mut u32)
    }

    /// the output bitmap of an op: allocated iff at least one input has one
    pub fn for_output(gpu_device: &Arc<GpuDevice>, len: usize, inputs: &[Option<&Self>]) -> Option<Self> {
        inputs.iter().any(Option::is_some).then(|| Self::new_empty(gpu_device, len))
    }

    /// null_bit_buffer.rs:130-166
    pub fn clone_null_bit_buffer(data: &Option<Self>) -> Option<Self> {
        data.as_ref().map(|d| Self { bit_buffer: d.gpu_device.clone_buffer(&d.bit_buffer), len: d.len, gpu_device: d.gpu_device.clone() })
    }

    pub fn clone_null_bit_buffer_pass(data: &Option<Self>, _pipeline: &mut ArrowComputePipeline) -> Option<Self> {
        Self::clone_null_bit_buffer(data)
    }

    pub fn clone_null_bit_buffer_op(data: &Option<Self>, _pipeline: &mut ArrowComputePipeline) -> Option<Self> {
        Self::clone_null_bit_buffer(data)
    }

    /// null_bit_buffer.rs:168-204: AND of both bitmaps; one-sided -> copy; none -> None.
    /// The element-wise kernels do this in their own pass (`vout` argument of `agpu_binary` ...);
    /// this stand-alone form remains for callers of the reference API.
    pub fn merge_null_bit_buffer(left: &Option<Self>, right: &Option<Self>) -> Option<Self> {
        let reference = left.as_ref().or(right.as_ref())?;
        let out = Self::new_empty(&reference.gpu_device, reference.len);
        check(
            unsafe {
                agpu_validity_and(reference.gpu_device.handle(), Self::words(left.as_ref()), Self::words(right.as_ref()),
                                  out.bit_buffer.ptr() as *mut u32, reference.len)
            },
            "merge_null_bit_buffer",
        );
        Some(out)
    }

    /// null_bit_buffer.rs:206-243
    pub fn merge_null_bit_buffer_op(left: &Option<Self>, right: &Option<Self>, _pipeline: &mut ArrowComputePipeline) -> Option<Self> {
        Self::merge_null_bit_buffer(left, right)
    }
}

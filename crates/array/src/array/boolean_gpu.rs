//! `BooleanArrayGPU` (array/boolean_gpu.rs:15-21): `data` is a packed LSB-first bitmap.
use std::sync::Arc;

use super::buffer::ArrowGpuBuffer;
use super::{ArrayUtils, ArrowType, BooleanBufferBuilder, NullBitBufferGpu};
use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};
use crate::kernels::broadcast::Broadcast;

pub struct BooleanArrayGPU {
    pub data: ArrowGpuBuffer,
    pub gpu_device: Arc<GpuDevice>,
    /// Actual len of the array (bits)
    pub len: usize,
    pub null_buffer: Option<NullBitBufferGpu>,
}

impl BooleanArrayGPU {
    /// boolean_gpu.rs:24-47
    pub fn from_optional_slice(value: &[Option<bool>], gpu_device: Arc<GpuDevice>) -> Self {
        let mut bits = BooleanBufferBuilder::new_with_capacity(value.len());
        let mut valid = BooleanBufferBuilder::new_with_capacity(value.len());
        for (i, v) in value.iter().enumerate() {
            if let Some(b) = v {
                valid.set_bit(i);
                if *b {
                    bits.set_bit(i);
                }
            }
        }
        let data = gpu_device.create_gpu_buffer_with_data(&bits.data);
        let null_buffer = NullBitBufferGpu::new(gpu_device.clone(), &valid);
        Self { data, gpu_device, len: value.len(), null_buffer }
    }

    /// boolean_gpu.rs:49-70
    pub fn from_slice(value: &[bool], gpu_device: Arc<GpuDevice>) -> Self {
        let mut bits = BooleanBufferBuilder::new_with_capacity(value.len());
        for (i, v) in value.iter().enumerate() {
            if *v {
                bits.set_bit(i);
            }
        }
        let data = gpu_device.create_gpu_buffer_with_data(&bits.data);
        Self { data, gpu_device, len: value.len(), null_buffer: None }
    }

    /// boolean_gpu.rs:72-82.  The reference sets `len` = number of BYTES; `len_bits`
    /// makes the intent explicit, `None` keeps every bit of the slice.
    pub fn from_bytes_slice(value: &[u8], gpu_device: Arc<GpuDevice>, len_bits: Option<usize>) -> Self {
        let mut padded = value.to_vec();
        padded.resize(value.len().div_ceil(4) * 4, 0);
        let data = gpu_device.create_gpu_buffer_with_data(&padded);
        Self { data, gpu_device, len: len_bits.unwrap_or(value.len() * 8), null_buffer: None }
    }

    pub fn new_empty(gpu_device: &Arc<GpuDevice>, len: usize, null_buffer: Option<NullBitBufferGpu>) -> Self {
        let data = gpu_device.create_empty_buffer((len.div_ceil(32) * 4) as u64);
        Self { data, gpu_device: gpu_device.clone(), len, null_buffer }
    }

    /// boolean_gpu.rs:84-96
    pub fn raw_values(&self) -> Option<Vec<bool>> {
        let bytes = self.gpu_device.retrive_data(&self.data);
        Some((0..self.len).map(|i| BooleanBufferBuilder::is_set_in_slice(&bytes, i)).collect())
    }

    /// boolean_gpu.rs:98-117
    pub fn values(&self) -> Vec<Option<bool>> {
        let raw = self.raw_values().unwrap();
        match &self.null_buffer {
            None => raw.into_iter().map(Some).collect(),
            Some(nb) => {
                let bits = nb.raw_values();
                raw.into_iter().enumerate().map(|(i, v)| BooleanBufferBuilder::is_set_in_slice(&bits, i).then_some(v)).collect()
            }
        }
    }

    pub fn clone_array(&self) -> Self {
        Self {
            data: self.gpu_device.clone_buffer(&self.data),
            gpu_device: self.gpu_device.clone(),
            len: self.len,
            null_buffer: NullBitBufferGpu::clone_null_bit_buffer(&self.null_buffer),
        }
    }

    pub fn arrow_type(&self) -> ArrowType {
        ArrowType::BooleanType
    }

    pub fn bits_ptr(&self) -> *const u32 {
        self.data.ptr_on(&self.gpu_device) as *const u32
    }

    pub fn validity_ptr(&self) -> *const u32 {
        NullBitBufferGpu::words(self.null_buffer.as_ref())
    }
}

impl ArrayUtils for BooleanArrayGPU {
    fn get_gpu_device(&self) -> Arc<GpuDevice> {
        self.gpu_device.clone()
    }
}

impl std::fmt::Debug for BooleanArrayGPU {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "BooleanArrayGPU {{ len: {}, values: {:?} }}", self.len, self.values())
    }
}

/// boolean_gpu.rs:119-135: built on the host like the reference
impl Broadcast<bool> for BooleanArrayGPU {
    fn broadcast_op(value: bool, len: usize, pipeline: &mut ArrowComputePipeline) -> Self {
        let builder = if value { BooleanBufferBuilder::new_set_with_capacity(len) } else { BooleanBufferBuilder::new_with_capacity(len) };
        let data = pipeline.device.create_gpu_buffer_with_data(&builder.data);
        Self { data, gpu_device: pipeline.device.clone(), len, null_buffer: None }
    }
}

//! `PrimitiveArrayGpu<T>` (array/primitive_array_gpu.rs:12-19): same public fields as the reference.
use std::fmt::{Debug, Formatter};
use std::marker::PhantomData;
use std::os::raw::c_void;
use std::sync::Arc;

use super::buffer::ArrowGpuBuffer;
use super::{ArrayUtils, ArrowPrimitiveType, ArrowType, BooleanBufferBuilder, NullBitBufferGpu};
use crate::gpu_utils::ffi::*;
use crate::gpu_utils::{ArrowComputePipeline, GpuDevice};
use crate::kernels::broadcast::Broadcast;

pub struct PrimitiveArrayGpu<T: ArrowPrimitiveType> {
    pub data: ArrowGpuBuffer,
    pub gpu_device: Arc<GpuDevice>,
    pub phantom: PhantomData<T>,
    /// number of rows (not bytes)
    pub len: usize,
    pub null_buffer: Option<NullBitBufferGpu>,
}

impl<T: ArrowPrimitiveType> PrimitiveArrayGpu<T> {
    /// primitive_array_gpu.rs:22-55 — null slots store `T::default()` (:39-41)
    pub fn from_optional_slice(value: &[Option<T::NativeType>], gpu_device: Arc<GpuDevice>) -> Self {
        let mut builder = BooleanBufferBuilder::new_with_capacity(value.len());
        let dense: Vec<T::NativeType> = value
            .iter()
            .enumerate()
            .map(|(i, v)| {
                if v.is_some() {
                    builder.set_bit(i);
                }
                v.unwrap_or_default()
            })
            .collect();
        let data = gpu_device.create_gpu_buffer_with_data(&dense);
        let null_buffer = NullBitBufferGpu::new(gpu_device.clone(), &builder);
        Self { data, gpu_device, phantom: PhantomData, len: value.len(), null_buffer }
    }

    /// primitive_array_gpu.rs:57-68
    pub fn from_slice(value: &[T::NativeType], gpu_device: Arc<GpuDevice>) -> Self {
        let data = gpu_device.create_gpu_buffer_with_data(value);
        Self { data, gpu_device, phantom: PhantomData, len: value.len(), null_buffer: None }
    }

    /// an uninitialised array of `len` rows: the output of an op (every kernel writes all of it)
    pub fn new_empty(gpu_device: &Arc<GpuDevice>, len: usize, null_buffer: Option<NullBitBufferGpu>) -> Self {
        let data = gpu_device.create_empty_buffer(len as u64 * T::ITEM_SIZE);
        Self { data, gpu_device: gpu_device.clone(), phantom: PhantomData, len, null_buffer }
    }

    /// primitive_array_gpu.rs:70-74
    pub fn raw_values(&self) -> Option<Vec<T::NativeType>> {
        let mut out = vec![T::NativeType::default(); self.len];
        let bytes = self.len * T::ITEM_SIZE as usize;
        check(unsafe { agpu_d2h(self.gpu_device.handle(), out.as_mut_ptr() as *mut c_void, self.data.ptr(), bytes) }, "raw_values");
        Some(out)
    }

    /// primitive_array_gpu.rs:76-97
    pub fn values(&self) -> Vec<Option<T::NativeType>> {
        let raw = self.raw_values().unwrap();
        match &self.null_buffer {
            None => raw.into_iter().map(Some).collect(),
            Some(nb) => {
                let bits = nb.raw_values();
                raw.into_iter().enumerate().map(|(i, v)| BooleanBufferBuilder::is_set_in_slice(&bits, i).then_some(v)).collect()
            }
        }
    }

    /// primitive_array_gpu.rs:99-104
    pub fn clone_array(&self) -> Self {
        Self {
            data: self.gpu_device.clone_buffer(&self.data),
            gpu_device: self.gpu_device.clone(),
            phantom: PhantomData,
            len: self.len,
            null_buffer: NullBitBufferGpu::clone_null_bit_buffer(&self.null_buffer),
        }
    }

    pub fn arrow_type(&self) -> ArrowType {
        T::ARROW_TYPE
    }

    pub fn values_ptr(&self) -> *const c_void {
        self.data.ptr_on(&self.gpu_device)
    }

    pub fn validity_ptr(&self) -> *const u32 {
        NullBitBufferGpu::words(self.null_buffer.as_ref())
    }
}

impl<T: ArrowPrimitiveType> ArrayUtils for PrimitiveArrayGpu<T> {
    fn get_gpu_device(&self) -> Arc<GpuDevice> {
        self.gpu_device.clone()
    }
}

impl<T: ArrowPrimitiveType> Debug for PrimitiveArrayGpu<T> {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result {
        write!(f, "{:?} {{ len: {}, values: {:?} }}", T::ARROW_TYPE, self.len, self.values())
    }
}

/// array/src/kernels/broadcast.rs:6-17 + array/compute_shaders/*/broadcast.wgsl
impl<T: ArrowPrimitiveType> Broadcast<T::NativeType> for PrimitiveArrayGpu<T> {
    fn broadcast_op(value: T::NativeType, len: usize, pipeline: &mut ArrowComputePipeline) -> Self {
        let out = Self::new_empty(&pipeline.device, len, None);
        check(unsafe { agpu_broadcast(pipeline.device.handle(), T::DTYPE, &value as *const _ as *const c_void, out.data.ptr(), len) }, "broadcast");
        out
    }
}

//! `GpuDevice` = one CUDA device ordinal + one stream + the caching allocator behind it.
//! Replaces the wgpu Device/Queue pair and the shader-text keyed pipeline cache
//! (crates/array/src/gpu_utils/gpu_device.rs:29-33, 137-168).
use std::os::raw::c_void;
use std::ptr;

use super::ffi::*;
use crate::array::buffer::ArrowGpuBuffer;

pub struct GpuDevice {
    pub(crate) handle: *mut AgpuDevice,
}

// the C side is re-entrant per handle (one mutex around the allocator state): gpu_device.rs:29-33
unsafe impl Send for GpuDevice {}
unsafe impl Sync for GpuDevice {}

impl GpuDevice {
    /// gpu_device.rs:46-85
    pub fn new() -> GpuDevice {
        Self::with_ordinal(0)
    }

    pub fn with_ordinal(ordinal: i32) -> GpuDevice {
        let mut handle = ptr::null_mut();
        check(unsafe { agpu_device_create(ordinal, &mut handle) }, "GpuDevice::new");
        GpuDevice { handle }
    }

    pub fn handle(&self) -> *mut AgpuDevice {
        self.handle
    }

    /// gpu_device.rs:171-181
    pub fn create_gpu_buffer_with_data<T: Copy>(self: &std::sync::Arc<Self>, data: &[T]) -> ArrowGpuBuffer {
        let bytes = std::mem::size_of_val(data);
        let buffer = self.create_empty_buffer(bytes as u64);
        check(unsafe { agpu_h2d(self.handle, buffer.ptr(), data.as_ptr() as *const c_void, bytes) }, "create_gpu_buffer_with_data");
        check(unsafe { agpu_sync(self.handle) }, "create_gpu_buffer_with_data"); // `data` may be a temporary
        buffer
    }

    /// gpu_device.rs:183-192 — NOT zero-filled: every kernel writes its whole output
    pub fn create_empty_buffer(self: &std::sync::Arc<Self>, size: u64) -> ArrowGpuBuffer {
        let mut p = ptr::null_mut();
        check(unsafe { agpu_alloc(self.handle, size.max(1) as usize, &mut p) }, "create_empty_buffer");
        ArrowGpuBuffer::from_raw(self.clone(), p, size)
    }

    /// gpu_device.rs:203-210
    pub fn create_scalar_buffer<T: Copy>(self: &std::sync::Arc<Self>, value: &T) -> ArrowGpuBuffer {
        self.create_gpu_buffer_with_data(std::slice::from_ref(value))
    }

    /// gpu_device.rs:212-222
    pub fn clone_buffer(self: &std::sync::Arc<Self>, buffer: &ArrowGpuBuffer) -> ArrowGpuBuffer {
        let out = self.create_empty_buffer(buffer.size());
        check(unsafe { agpu_d2d(self.handle, out.ptr(), buffer.ptr(), buffer.size() as usize) }, "clone_buffer");
        out
    }

    /// gpu_device.rs:232-265 — the only host synchronisation point
    pub fn retrive_data(&self, buffer: &ArrowGpuBuffer) -> Vec<u8> {
        let mut out = vec![0u8; buffer.size() as usize];
        check(unsafe { agpu_d2h(self.handle, out.as_mut_ptr() as *mut c_void, buffer.ptr(), out.len()) }, "retrive_data");
        out
    }

    pub fn sync(&self) {
        check(unsafe { agpu_sync(self.handle) }, "sync");
    }
}

impl Default for GpuDevice {
    fn default() -> Self {
        Self::new()
    }
}

impl Drop for GpuDevice {
    fn drop(&mut self) {
        unsafe { agpu_device_destroy(self.handle) };
    }
}

impl std::fmt::Debug for GpuDevice {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "GpuDevice(cuda:{})", unsafe { agpu_device_ordinal(self.handle) })
    }
}

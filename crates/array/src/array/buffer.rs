//! `ArrowGpuBuffer` (array/buffer.rs:5-7): one device allocation.  The reference wraps
//! `Arc<wgpu::Buffer>`; here it is {device pointer, size, owning GpuDevice} and `Drop` is a
//! stream-ordered `agpu_free`, so dropping an input right after enqueueing work on it is safe —
//! the guarantee wgpu gives by keeping buffers alive until submitted work is done.
use std::os::raw::c_void;
use std::sync::Arc;

use crate::gpu_utils::ffi::*;
use crate::gpu_utils::GpuDevice;

#[derive(Debug)]
pub struct ArrowGpuBuffer {
    ptr: *mut c_void,
    size: u64,
    device: Arc<GpuDevice>,
}

unsafe impl Send for ArrowGpuBuffer {}
unsafe impl Sync for ArrowGpuBuffer {}

impl ArrowGpuBuffer {
    pub(crate) fn from_raw(device: Arc<GpuDevice>, ptr: *mut c_void, size: u64) -> Self {
        Self { ptr, size, device }
    }

    /// bytes (buffer.rs:22-24)
    pub fn size(&self) -> u64 {
        self.size
    }

    pub fn ptr(&self) -> *mut c_void {
        self.ptr
    }

    /// the pointer for work about to be enqueued on ANOTHER handle of the same GPU: after `Drop` the
    /// block is not handed out again before that handle's stream got there
    pub fn ptr_on(&self, user: &GpuDevice) -> *mut c_void {
        if !std::ptr::eq(user, self.device.as_ref()) {
            check(unsafe { agpu_buffer_record_use(user.handle(), self.ptr) }, "agpu_buffer_record_use");
        }
        self.ptr
    }
}

impl Drop for ArrowGpuBuffer {
    fn drop(&mut self) {
        unsafe { agpu_free(self.device.handle(), self.ptr) };
    }
}

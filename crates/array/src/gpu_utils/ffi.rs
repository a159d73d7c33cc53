//! The `extern "C"` declarations of include/agpu.h.  The text is generated
//! (include/gen_rust_ffi.py) so that it cannot drift from the header; tests/test_abi.py compares it
//! with the header and with the ctypes table the Python tests call through.
include!("../../../../include/agpu_ffi.rs");

//! array/src/utils/mod.rs: host-side vectors of any element type (`ArrowArrayGPU::get_raw_values`)
#[derive(Debug, PartialEq)]
pub enum ScalarArray {
    F32Vec(Vec<f32>),
    U32Vec(Vec<u32>),
    U16Vec(Vec<u16>),
    U8Vec(Vec<u8>),
    I32Vec(Vec<i32>),
    I16Vec(Vec<i16>),
    I8Vec(Vec<i8>),
    BOOLVec(Vec<bool>),
}

macro_rules! into_scalar_array {
    ($($t:ty => $variant:ident),*) => {$(
        impl From<Vec<$t>> for ScalarArray {
            fn from(value: Vec<$t>) -> Self {
                ScalarArray::$variant(value)
            }
        }
    )*};
}
into_scalar_array!(f32 => F32Vec, u32 => U32Vec, u16 => U16Vec, u8 => U8Vec, i32 => I32Vec, i16 => I16Vec, i8 => I8Vec, bool => BOOLVec);

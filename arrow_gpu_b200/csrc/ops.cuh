// ops.cuh — per-row semantics, bit-identical to the reference's WGSL shaders as restated in
// oracle/oracle.c (WGSL spec rules).  Integer lanes narrower than 32
// bits are widened (sign/zero extend), operated on in 32-bit and truncated — natively here,
// where the reference needs its u32-packing helpers (compute_shaders/*/utils.wgsl).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

template <typename T> struct Wide { using type = typename std::conditional<std::is_signed<T>::value, int32_t, uint32_t>::type; };
template <> struct Wide<float> { using type = float; };

template <typename T> __device__ __forceinline__ typename Wide<T>::type widen(T v) { return (typename Wide<T>::type)v; }
template <typename T, typename W> __device__ __forceinline__ T narrow(W v) { return (T)(typename std::make_unsigned<typename std::conditional<std::is_same<T, float>::value, int32_t, T>::type>::type)(uint32_t)v; }

// ---- arithmetic: arithmetic/compute_shaders/{f32,i32,u32}/{array,scalar}.wgsl, u16/scalar.wgsl ----
template <typename T> struct OpAdd {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return __fadd_rn(a, b);
    else return (T)((uint32_t)a + (uint32_t)b);
  }
};
template <typename T> struct OpSub {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return __fsub_rn(a, b);
    else return (T)((uint32_t)a - (uint32_t)b);
  }
};
template <typename T> struct OpMul {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return __fmul_rn(a, b);
    else return (T)((uint32_t)a * (uint32_t)b);
  }
};
// WGSL: x/0 = x, MIN/-1 = MIN (never traps); f32: IEEE division
template <typename T> struct OpDiv {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return __fdiv_rn(a, b);
    else if constexpr (std::is_signed<T>::value) {
      int32_t x = a, y = b;
      if (y == 0 || (x == INT32_MIN && y == -1)) return a;
      return (T)(uint32_t)(x / y);
    } else {
      uint32_t x = a, y = b;
      return y == 0 ? a : (T)(x / y);
    }
  }
};
// WGSL: x%0 = 0, MIN%-1 = 0, sign of the dividend; f32: a - b*trunc(a/b), three roundings
template <typename T> struct OpRem {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) {
      return __fsub_rn(a, __fmul_rn(b, truncf(__fdiv_rn(a, b))));
    } else if constexpr (std::is_signed<T>::value) {
      int32_t x = a, y = b;
      if (y == 0 || (x == INT32_MIN && y == -1)) return (T)0;
      return (T)(uint32_t)(x % y);
    } else {
      uint32_t x = a, y = b;
      return y == 0 ? (T)0 : (T)(x % y);
    }
  }
};
// ---- min/max: compare/compute_shaders/*/min_max.wgsl (f32: NaN-ignoring, -0 < +0; u32 unsigned: Q3) ----
// 8/16-bit lanes: whole 32-bit words through the SIMD-in-word video intrinsics (4 or 2 lanes per
// instruction sequence) instead of extract / compare / insert per lane.
template <typename T> struct OpMin {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return fminf(a, b);
    else return a < b ? a : b;
  }
  static constexpr bool kWord = sizeof(T) < 4;
  static __device__ __forceinline__ uint32_t word(uint32_t a, uint32_t b) {
    if constexpr (std::is_same<T, int8_t>::value) return __vmins4(a, b);
    else if constexpr (std::is_same<T, uint8_t>::value) return __vminu4(a, b);
    else if constexpr (std::is_same<T, int16_t>::value) return __vmins2(a, b);
    else return __vminu2(a, b);
  }
};
template <typename T> struct OpMax {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) return fmaxf(a, b);
    else return a > b ? a : b;
  }
  static constexpr bool kWord = sizeof(T) < 4;
  static __device__ __forceinline__ uint32_t word(uint32_t a, uint32_t b) {
    if constexpr (std::is_same<T, int8_t>::value) return __vmaxs4(a, b);
    else if constexpr (std::is_same<T, uint8_t>::value) return __vmaxu4(a, b);
    else if constexpr (std::is_same<T, int16_t>::value) return __vmaxs2(a, b);
    else return __vmaxu2(a, b);
  }
};
// ---- logical: logical/compute_shaders/{i32,u32}/logical.wgsl ----
template <typename T> struct OpAnd { __device__ __forceinline__ T operator()(T a, T b) const { return (T)(a & b); } };
template <typename T> struct OpOr { __device__ __forceinline__ T operator()(T a, T b) const { return (T)(a | b); } };
template <typename T> struct OpXor { __device__ __forceinline__ T operator()(T a, T b) const { return (T)(a ^ b); } };
// ---- power: math/compute_shaders/f32/floatbinary.wgsl:14-18, i32/binary.wgsl:13-29 ----
// i32: the shader multiplies p times (wrapping) or divides |p| times with WGSL `/`.  Wrapping
// products are associative, so square-and-multiply gives the same bits in O(log p); the
// division chain has a closed form: x == 0 -> 1 (r/0 = r), x == 1 -> 1, x == -1 -> (-1)^|p|,
// |x| >= 2 -> 0 after the first step (p < 0 so at least one step runs).
template <typename T> struct OpPow {
  __device__ __forceinline__ T operator()(T a, T b) const {
    if constexpr (std::is_same<T, float>::value) {
      // WGSL pow(x, y) = exp2(y * log2(x)): negative base -> NaN (pinned by the reference's
      // test_f32_power vector, math/src/f32.rs:210-271)
      return a < 0.0f ? __int_as_float(0x7fc00000) : powf(a, b);
    } else {
      int32_t x = a, p = b;
      if (p >= 0) {
        uint32_t r = 1, base = (uint32_t)x, e = (uint32_t)p;
        while (e) {
          if (e & 1u) r *= base;
          base *= base;
          e >>= 1;
        }
        return (T)r;
      }
      if (x == 0 || x == 1) return (T)1;
      if (x == -1) return (T)((p & 1) ? -1 : 1);
      return (T)0;
    }
  }
};

// ---- unary ----
template <typename T> struct OpNeg { __device__ __forceinline__ T operator()(T a) const { return -a; } };  // f32/neg.wgsl
template <typename T> struct OpAbs {  // floatunary.wgsl:42, i32/unary.wgsl (abs(MIN) = MIN)
  __device__ __forceinline__ T operator()(T a) const {
    if constexpr (std::is_same<T, float>::value) return fabsf(a);
    else return a < 0 ? (T)(0u - (uint32_t)a) : a;
  }
};
template <typename T> struct OpNot { __device__ __forceinline__ T operator()(T a) const { return (T)~a; } };  // not.wgsl

// f32 math: math/compute_shaders/f32/floatunary.wgsl; trig: trigonometry/compute_shaders/*
// The input type TI is converted exactly to f32 first (the reference's fused cast+trig shaders).
template <typename TI> struct FSqrt { __device__ __forceinline__ float operator()(TI a) const { return __fsqrt_rn((float)a); } };
// floatunary.wgsl:46-54: cbrt(x) = sign(x) * pow(|x|, 1.0/3.0) with the f32 constant 1/3 =
// 0.3333333433.  powf costs ~70 instructions, so the same value is computed as
//   |x|^(1/3f) = cbrt(|x|) * |x|^d,  d = 1/3f - 1/3 = 9.934e-9,  |x|^d = 1 + d*ln|x| (+ O(1e-12))
// i.e. cbrtf (1 ULP) plus a sub-ULP correction: half the instructions, and closer to the oracle's
// pow(x, (double)(1/3f)) than powf's own 4 ULP bound (tests: <= 2 ULP).
template <typename TI> struct FCbrt {
  __device__ __forceinline__ float operator()(TI a) const {
    const float x = (float)a;
    const float m = fabsf(x);
    float r = cbrtf(m);
    // the correction is < 1e-6 relative, so ln|x| only needs ~3 digits: the SFU's lg2 is plenty
    if (m > 0.0f && m < __int_as_float(0x7f800000)) r = fmaf(r * (9.934107e-9f * 0.69314718f), __log2f(m), r);
    return x < 0.0f ? -r : r;   // NaN -> NaN, -0.0 takes the non-negative branch like the shader
  }
};
template <typename TI> struct FExp { __device__ __forceinline__ float operator()(TI a) const { return expf((float)a); } };
template <typename TI> struct FExp2 { __device__ __forceinline__ float operator()(TI a) const { return exp2f((float)a); } };
template <typename TI> struct FLog { __device__ __forceinline__ float operator()(TI a) const { return logf((float)a); } };
template <typename TI> struct FLog2 { __device__ __forceinline__ float operator()(TI a) const { return log2f((float)a); } };
template <typename TI> struct FSin { __device__ __forceinline__ float operator()(TI a) const { return sinf((float)a); } };
template <typename TI> struct FCos { __device__ __forceinline__ float operator()(TI a) const { return cosf((float)a); } };
template <typename TI> struct FAcos { __device__ __forceinline__ float operator()(TI a) const { return acosf((float)a); } };
template <typename TI> struct FSinh { __device__ __forceinline__ float operator()(TI a) const { return sinhf((float)a); } };

// ---- shifts: logical/compute_shaders/*/shift.wgsl — widen, shift by (count & 31), truncate ----
template <typename T> struct OpShl {
  __device__ __forceinline__ T operator()(T a, uint32_t c) const {
    return (T)((uint32_t)(typename Wide<T>::type)a << (c & 31u));
  }
};
template <typename T> struct OpShr {
  __device__ __forceinline__ T operator()(T a, uint32_t c) const {
    if constexpr (std::is_signed<T>::value) return (T)(uint32_t)((int32_t)a >> (c & 31u));
    else return (T)((uint32_t)a >> (c & 31u));
  }
};

// ---- casts: cast/compute_shaders/* ----
template <typename TI, typename TO> struct OpCast {
  __device__ __forceinline__ TO operator()(TI a) const {
    if constexpr (std::is_same<TI, float>::value) {
      // f32/cast_u8.wgsl:14-21: u32(f) % 256 with WGSL's saturating u32() (NaN/negative -> 0)
      static_assert(std::is_same<TO, uint8_t>::value, "only f32 -> u8 exists");
      return (TO)(__float2uint_rz(a) & 0xffu);  // cvt.rzi.u32.f32 saturates, NaN -> 0
    } else if constexpr (std::is_same<TO, float>::value) {
      return (float)a;  // exact for <= 16-bit integers
    } else {
      return (TO)(uint32_t)(typename Wide<TI>::type)a;  // sign/zero extend, keep low bits
    }
  }
};

// ---- compare predicates: compare/compute_shaders/*/cmp.wgsl ----
// Each predicate also knows how to compare a packed word of 8- or 16-bit lanes at once (SIMD-in-
// word video intrinsics): the result has 0xFF / 0xFFFF in every lane where the predicate holds.
#define AGPU_PRED(NAME, EXPR, S4, U4, S2, U2)                                                        \
  struct NAME {                                                                                      \
    template <typename T> __device__ __forceinline__ bool operator()(T a, T b) const { return EXPR; } \
    template <typename T> static __device__ __forceinline__ uint32_t lanes(uint32_t a, uint32_t b) {  \
      if constexpr (std::is_same<T, int8_t>::value) return S4(a, b);                                 \
      else if constexpr (std::is_same<T, uint8_t>::value) return U4(a, b);                           \
      else if constexpr (std::is_same<T, int16_t>::value) return S2(a, b);                           \
      else return U2(a, b);                                                                          \
    }                                                                                                \
  };
AGPU_PRED(PGt, a > b, __vcmpgts4, __vcmpgtu4, __vcmpgts2, __vcmpgtu2)
AGPU_PRED(PGe, a >= b, __vcmpges4, __vcmpgeu4, __vcmpges2, __vcmpgeu2)
AGPU_PRED(PLt, a < b, __vcmplts4, __vcmpltu4, __vcmplts2, __vcmpltu2)
AGPU_PRED(PLe, a <= b, __vcmples4, __vcmpleu4, __vcmples2, __vcmpleu2)
AGPU_PRED(PEq, a == b, __vcmpeq4, __vcmpeq4, __vcmpeq2, __vcmpeq2)
#undef AGPU_PRED

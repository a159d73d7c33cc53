/*
 * oracle.c — CPU restatement of psvri/arrow-gpu's compute shaders.  TEST INFRASTRUCTURE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libagpu.so) never links or calls it.
 *
 * What it restates: the WGSL compute shaders of the reference's hot path,
 * one function per shader family, each citing the shader it follows (paths relative to the
 * reference root).  The shaders are interpreted with WGSL-spec semantics:
 *   int + - * wrap mod 2^32; `/` truncates, x/0 = x, MIN/-1 = MIN; `%` has the sign of the
 *   dividend, x%0 = 0, MIN%-1 = 0; shift counts are taken mod 32; i32 >> is arithmetic;
 *   u32(f32) truncates toward zero and saturates (NaN -> 0); f32 compares are IEEE;
 *   min/max on f32 return the non-NaN operand; new storage buffers are zero-initialised and
 *   out-of-bounds accesses do nothing.
 * The arithmetic itself lives in third-party naga 24.0.0 / wgpu 24.0.3 + the Vulkan driver
 * (Cargo.lock:1263-1264, 2002-2003), which are not vendored in the reference tree and cannot
 * be built here (no Rust, no Vulkan ICD).
 *
 * PARITY PINNING: this oracle is checked against every golden vector of the reference's own
 * unit tests (tests/golden/reference_vectors.json, extracted by
 * tests/golden/extract_reference_vectors.py; see tests/test_oracle_golden.py).  Integer,
 * boolean, bitmap, cast and indexing results are pinned exactly; f32 transcendentals are
 * pinned to the reference's own tolerance (abs 0.01).  PARITY UNPINNED (no reference vector
 * exists; the WGSL-spec value is used): integer divide/remainder by zero, shift counts >=
 * width, f32->u8 of negatives/NaN, pow(neg, 0), sub-word arithmetic that the reference does
 * not implement (i8/u8/i16 + - *, u16 - *), filter/compaction.
 *
 * Sub-word types: the reference processes i8/u8/i16/u16 columns as packed u32 words with
 * mask/shift helpers (compute_shaders/{i8,u8,i16,u16}/utils.wgsl).  Those shaders are
 * restated here at word level, helpers included, on zero-padded copies of the inputs, so the
 * oracle follows the reference's algorithm rather than a re-derivation of it.
 *
 * Build: gcc -O2 -fopenmp -fno-fast-math -ffp-contract=off -shared -fPIC oracle.c -lm
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#define OMP_FOR _Pragma("omp parallel for schedule(static)")
#else
#define OMP_FOR
#endif

typedef int64_t idx_t; /* signed loop index for OpenMP */

/* ------------------------------------------------------------------------------------------
 * WGSL scalar semantics
 * ---------------------------------------------------------------------------------------- */
static inline int32_t wrap_i32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }
static inline int32_t i32_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t i32_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t i32_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
/* WGSL: e1 / e2 -> e1 when e2 == 0 or (e1 == MIN and e2 == -1) */
static inline int32_t i32_div(int32_t a, int32_t b) {
  if (b == 0 || (a == INT32_MIN && b == -1)) return a;
  return a / b;
}
/* WGSL: e1 % e2 -> 0 when e2 == 0 or (e1 == MIN and e2 == -1) */
static inline int32_t i32_rem(int32_t a, int32_t b) {
  if (b == 0 || (a == INT32_MIN && b == -1)) return 0;
  return a % b;
}
static inline uint32_t u32_div(uint32_t a, uint32_t b) { return b == 0 ? a : a / b; }
static inline uint32_t u32_rem(uint32_t a, uint32_t b) { return b == 0 ? 0u : a % b; }
static inline int32_t i32_shl(int32_t a, uint32_t c) { return (int32_t)((uint32_t)a << (c & 31u)); }
static inline int32_t i32_shr(int32_t a, uint32_t c) {
  uint32_t s = c & 31u; /* arithmetic shift without relying on implementation-defined >> */
  uint32_t u = (uint32_t)a >> s;
  if (a < 0 && s) u |= ~(0xFFFFFFFFu >> s);
  return (int32_t)u;
}
static inline uint32_t u32_shl(uint32_t a, uint32_t c) { return a << (c & 31u); }
static inline uint32_t u32_shr(uint32_t a, uint32_t c) { return a >> (c & 31u); }
static inline int32_t i32_abs(int32_t a) { return a < 0 ? wrap_i32(-(int64_t)a) : a; }
/* WGSL u32(f32): truncate toward zero, saturate, NaN -> 0 */
static inline uint32_t f32_to_u32(float f) {
  if (!(f == f) || f <= 0.0f) return 0u;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}
/* WGSL max/min on f32: the non-NaN operand if one is NaN; -0 < +0 (fmaxf/fminf of IEEE
 * 754-2019 maximumNumber/minimumNumber; CUDA's fmaxf/fminf agree) */
static inline float f32_max(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return signbit(a) ? b : a;
  return a > b ? a : b;
}
static inline float f32_min(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return signbit(a) ? a : b;
  return a < b ? a : b;
}
/* WGSL f32 `%`: e1 - e2 * trunc(e1 / e2), each operation rounded to f32 */
static inline float f32_rem(float a, float b) {
  volatile float q = a / b;
  volatile float t = truncf(q);
  volatile float p = b * t;
  return a - p;
}

/* ------------------------------------------------------------------------------------------
 * WGSL helper functions of the sub-word shaders
 * ---------------------------------------------------------------------------------------- */
/* compute_shaders/i8/utils.wgsl:11-51 — sign-extending byte extractors on an i32 word */
static inline int32_t i8_get_left_byte(int32_t d) {
  int32_t sign = d & 0x80, value = d & 0x7f;
  return sign == 0 ? value : (value | sign | -256);
}
static inline int32_t i8_get_mid_left_byte(int32_t d) {
  int32_t sign = (d & (0x80 << 8)) >> 8, value = (d & (0x7f << 8)) >> 8;
  return sign == 0 ? value : (value | sign | -256);
}
static inline int32_t i8_get_mid_right_byte(int32_t d) {
  int32_t sign = (d & (0x80 << 16)) >> 16, value = (d & (0x7f << 16)) >> 16;
  return sign == 0 ? value : (value | sign | -256);
}
static inline int32_t i8_get_right_byte(int32_t d) {
  /* (data & sign_extractor) >> 24u is an arithmetic shift of a possibly negative i32 */
  int32_t sign = i32_shr(d & (int32_t)0x80000000u, 24);
  int32_t value = (d & (0x7f << 24)) >> 24;
  return sign == 0 ? value : (value | sign | -256);
}
/* compute_shaders/u8/utils.wgsl:9-23 */
static inline uint32_t u8_get_left_byte(uint32_t d) { return d & 0x000000ffu; }
static inline uint32_t u8_get_mid_left_byte(uint32_t d) { return (d & 0x0000ff00u) >> 8; }
static inline uint32_t u8_get_mid_right_byte(uint32_t d) { return (d & 0x00ff0000u) >> 16; }
static inline uint32_t u8_get_right_byte(uint32_t d) { return (d & 0xff000000u) >> 24; }
/* compute_shaders/i16/utils.wgsl:13-29 */
static inline int32_t i16_get_left_half(int32_t d) {
  int32_t sign = d & 0x00008000;
  return sign == 0 ? (d & 0xffff) : ((d & 0xffff) | -65536);
}
static inline int32_t i16_get_right_half(int32_t d) {
  int32_t sign = i32_shr(d & INT32_MIN, 16);
  int32_t hi = i32_shr(d & -65536, 16);
  return sign == 0 ? hi : (hi | -65536);
}
/* compute_shaders/u16/utils.wgsl:8-18 */
static inline uint32_t u16_get_left_half(uint32_t d) { return d & 0x0000ffffu; }
static inline uint32_t u16_get_right_half(uint32_t d) { return (d & 0xffff0000u) >> 16; }
static inline uint32_t u16_merge(uint32_t l, uint32_t r) { return (l & 0xffffu) | ((r & 0xffffu) << 16); }
/* WGSL builtins unpack4xI8 / unpack4xU8 / pack4xI8 / pack4xU8 */
static inline void unpack4xI8(uint32_t w, int32_t o[4]) {
  for (int k = 0; k < 4; ++k) o[k] = (int32_t)(int8_t)(uint8_t)(w >> (8 * k));
}
static inline void unpack4xU8(uint32_t w, uint32_t o[4]) {
  for (int k = 0; k < 4; ++k) o[k] = (w >> (8 * k)) & 0xffu;
}
static inline uint32_t pack4xI8(const int32_t v[4]) {
  uint32_t w = 0;
  for (int k = 0; k < 4; ++k) w |= ((uint32_t)v[k] & 0xffu) << (8 * k);
  return w;
}
static inline uint32_t pack4xU8(const uint32_t v[4]) {
  uint32_t w = 0;
  for (int k = 0; k < 4; ++k) w |= (v[k] & 0xffu) << (8 * k);
  return w;
}

/* ------------------------------------------------------------------------------------------
 * packed-word scratch: the reference keeps sub-word columns in buffers padded to 4 bytes
 * (primitive_array_gpu.rs:27-31); new buffers are zero-filled (wgpu).
 * ---------------------------------------------------------------------------------------- */
/* Whole-word, 4-byte aligned buffers are used in place (no copy); only a ragged tail needs the
 * zero-padded scratch copy.  WFREE / WOUT_DONE know which case they are in. */
static int word_exact(const void* p, size_t bytes) { return bytes && bytes % 4 == 0 && ((uintptr_t)p & 3u) == 0; }
static uint32_t* words_from(const void* src, size_t bytes, size_t* nwords) {
  size_t nw = (bytes + 3) / 4;
  *nwords = nw;
  if (word_exact(src, bytes)) return (uint32_t*)src;
  uint32_t* w = (uint32_t*)calloc(nw ? nw : 1, 4);
  if (w && bytes) memcpy(w, src, bytes);
  return w;
}
static uint32_t* words_zero(size_t nwords) { return (uint32_t*)calloc(nwords ? nwords : 1, 4); }
static uint32_t* words_out(void* dst, size_t bytes, size_t nwords) {
  if (word_exact(dst, bytes) && nwords * 4 == bytes) return (uint32_t*)dst;
  return words_zero(nwords);
}
#define WFREE(ptr, src) do { if ((const void*)(ptr) != (const void*)(src)) free(ptr); } while (0)
#define WOUT_DONE(ptr, dst, bytes) do { if ((void*)(ptr) != (void*)(dst)) { memcpy((dst), (ptr), (bytes)); free(ptr); } } while (0)

static inline size_t dtype_size(int dtype) {
  switch (dtype) {
    case AGPU_I8: case AGPU_U8: return 1;
    case AGPU_I16: case AGPU_U16: return 2;
    case AGPU_I32: case AGPU_U32: case AGPU_F32: case AGPU_DATE32: return 4;
    default: return 0;
  }
}
static inline int get_bit(const uint32_t* b, size_t i) { return (b[i >> 5] >> (i & 31)) & 1u; }
static inline size_t bit_words(size_t n) { return (n + 31) / 32; }
static inline void mask_tail(uint32_t* w, size_t n_bits) {
  if (n_bits & 31) w[n_bits >> 5] &= (1u << (n_bits & 31)) - 1u;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * validity: NullBitBufferGpu::merge_null_bit_buffer, crates/array/src/array/
 * null_bit_buffer.rs:168-243 (AND kernel = logical/compute_shaders/u32/logical.wgsl:15-17)
 * ---------------------------------------------------------------------------------------- */
int oracle_validity_and(const uint32_t* va, const uint32_t* vb, uint32_t* vout, size_t n_bits) {
  if (!vout || (!va && !vb)) return AGPU_EINVAL; /* (None, None) => None */
  idx_t nw = (idx_t)bit_words(n_bits);
  if (va && vb) {
    OMP_FOR for (idx_t i = 0; i < nw; ++i) vout[i] = va[i] & vb[i];
  } else {
    memcpy(vout, va ? va : vb, (size_t)nw * 4); /* one side: clone_buffer */
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * arithmetic, array ∘ array and array ∘ scalar
 *   f32: arithmetic/compute_shaders/f32/array.wgsl:13-35, f32/scalar.wgsl:13-41
 *   i32: i32/array.wgsl:13-17, i32/scalar.wgsl:13-42   u32: u32/array.wgsl, u32/scalar.wgsl
 *   u16 + scalar: u16/scalar.wgsl:13-23 (word level)
 * min/max: compare/compute_shaders/{f32,i32}/min_max.wgsl:13-23, u32/min_max.wgsl (declared
 *   array<i32>: signed compare — reproduced only when quirks != 0),
 *   u16/min_max.wgsl:13-27 (word level)
 * logical: logical/compute_shaders/{i32,u32}/logical.wgsl:13-29 (word-wise for all widths)
 * power: math/compute_shaders/f32/floatbinary.wgsl:14-18, i32/binary.wgsl:13-29
 * ---------------------------------------------------------------------------------------- */
static int g_quirks = 0;
void oracle_set_ref_quirks(int on) { g_quirks = on; }

static inline float f32_binop(int op, float a, float b, int* ok) {
  switch (op) {
    case AGPU_ADD: return a + b;
    case AGPU_SUB: return a - b;
    case AGPU_MUL: return a * b;
    case AGPU_DIV: return a / b;
    case AGPU_REM: return f32_rem(a, b);
    case AGPU_MIN: return f32_min(a, b);
    case AGPU_MAX: return f32_max(a, b);
    case AGPU_POW:
      /* WGSL defines pow(x, y) through exp2(y * log2(x)): a negative base is NaN whatever the
       * exponent.  Pinned by the reference's own vector test_f32_power (math/src/f32.rs:210-271:
       * (-1)^0 = (-10)^0 = (-inf)^(+-inf) = NaN), which it skips on Linux/macOS where the
       * driver answers differently. */
      if (a < 0.0f) return NAN;
      return (float)pow((double)a, (double)b);
    default: *ok = 0; return 0.0f;
  }
}
/* math/compute_shaders/i32/binary.wgsl:13-29 */
static inline int32_t i32_power(int32_t x, int32_t p) {
  int32_t r = 1;
  if (p >= 0) {
    for (int32_t i = 0; i < p; ++i) r = i32_mul(r, x);
  } else {
    int32_t lim = i32_abs(p);
    for (int32_t i = 0; i < lim; ++i) r = i32_div(r, x);
  }
  return r;
}
static inline int32_t i32_binop(int op, int32_t a, int32_t b, int* ok) {
  switch (op) {
    case AGPU_ADD: return i32_add(a, b);
    case AGPU_SUB: return i32_sub(a, b);
    case AGPU_MUL: return i32_mul(a, b);
    case AGPU_DIV: return i32_div(a, b);
    case AGPU_REM: return i32_rem(a, b);
    case AGPU_MIN: return a < b ? a : b;
    case AGPU_MAX: return a > b ? a : b;
    case AGPU_AND: return a & b;
    case AGPU_OR: return a | b;
    case AGPU_XOR: return a ^ b;
    case AGPU_POW: return i32_power(a, b);
    default: *ok = 0; return 0;
  }
}
static inline uint32_t u32_binop(int op, uint32_t a, uint32_t b, int* ok) {
  switch (op) {
    case AGPU_ADD: return a + b;
    case AGPU_SUB: return a - b;
    case AGPU_MUL: return a * b;
    case AGPU_DIV: return u32_div(a, b);
    case AGPU_REM: return u32_rem(a, b);
    case AGPU_MIN:
      if (g_quirks) return (uint32_t)((int32_t)a < (int32_t)b ? (int32_t)a : (int32_t)b);
      return a < b ? a : b;
    case AGPU_MAX:
      if (g_quirks) return (uint32_t)((int32_t)a > (int32_t)b ? (int32_t)a : (int32_t)b);
      return a > b ? a : b;
    case AGPU_AND: return a & b;
    case AGPU_OR: return a | b;
    case AGPU_XOR: return a ^ b;
    default: *ok = 0; return 0;
  }
}
/* sub-word lanes: widen (sign/zero), operate in 32-bit, keep the low bits — the rule the
 * reference's only sub-word arithmetic shader (u16/scalar.wgsl:15-23) and its u16 min/max
 * follow; i8/u8/i16 arithmetic is new surface (parity unpinned). */
static inline int32_t narrow_signed_binop(int op, int32_t a, int32_t b, int* ok) {
  return i32_binop(op, a, b, ok);
}

int oracle_binary(int op, int dtype, const void* a, const void* b, void* out, size_t n) {
  int ok = 1;
  idx_t N = (idx_t)n;
  switch (dtype) {
    case AGPU_F32: {
      const float *x = a, *y = b; float* o = out;
      if (op == AGPU_AND || op == AGPU_OR || op == AGPU_XOR) return AGPU_EUNSUPPORTED;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = f32_binop(op, x[i], y[i], &k); }
      f32_binop(op, 0, 0, &ok);
      break;
    }
    case AGPU_I32: case AGPU_DATE32: {
      const int32_t *x = a, *y = b; int32_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = i32_binop(op, x[i], y[i], &k); }
      i32_binop(op, 0, 1, &ok);
      break;
    }
    case AGPU_U32: {
      const uint32_t *x = a, *y = b; uint32_t* o = out;
      if (op == AGPU_POW) return AGPU_EUNSUPPORTED;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = u32_binop(op, x[i], y[i], &k); }
      u32_binop(op, 0, 1, &ok);
      break;
    }
    case AGPU_U16: {
      if (op == AGPU_POW) return AGPU_EUNSUPPORTED;
      if (op == AGPU_MIN || op == AGPU_MAX || op == AGPU_AND || op == AGPU_OR || op == AGPU_XOR) {
        /* word level: compare/compute_shaders/u16/min_max.wgsl:13-27, logical u32/logical.wgsl */
        size_t nw; uint32_t* x = words_from(a, n * 2, &nw); uint32_t* y = words_from(b, n * 2, &nw);
        uint32_t* o = words_out(out, n * 2, nw);
        OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
          if (op == AGPU_MIN || op == AGPU_MAX) {
            uint32_t ll = u16_get_left_half(x[i]), lr = u16_get_left_half(y[i]);
            uint32_t rl = u16_get_right_half(x[i]), rr = u16_get_right_half(y[i]);
            uint32_t l = op == AGPU_MIN ? (ll < lr ? ll : lr) : (ll > lr ? ll : lr);
            uint32_t r = op == AGPU_MIN ? (rl < rr ? rl : rr) : (rl > rr ? rl : rr);
            o[i] = u16_merge(l, r);
          } else {
            o[i] = op == AGPU_AND ? (x[i] & y[i]) : op == AGPU_OR ? (x[i] | y[i]) : (x[i] ^ y[i]);
          }
        }
        WOUT_DONE(o, out, n * 2);
        WFREE(x, a); WFREE(y, b);
        break;
      }
      const uint16_t *x = a, *y = b; uint16_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (uint16_t)u32_binop(op, x[i], y[i], &k); }
      u32_binop(op, 0, 1, &ok);
      break;
    }
    case AGPU_U8: {
      if (op == AGPU_POW) return AGPU_EUNSUPPORTED;
      if (op == AGPU_AND || op == AGPU_OR || op == AGPU_XOR) {
        size_t nw; uint32_t* x = words_from(a, n, &nw); uint32_t* y = words_from(b, n, &nw);
        uint32_t* o = words_out(out, n, nw);
        OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i)
          o[i] = op == AGPU_AND ? (x[i] & y[i]) : op == AGPU_OR ? (x[i] | y[i]) : (x[i] ^ y[i]);
        WOUT_DONE(o, out, n);
        WFREE(x, a); WFREE(y, b);
        break;
      }
      const uint8_t *x = a, *y = b; uint8_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (uint8_t)u32_binop(op, x[i], y[i], &k); }
      u32_binop(op, 0, 1, &ok);
      break;
    }
    case AGPU_I16: {
      if (op == AGPU_POW) return AGPU_EUNSUPPORTED;
      const int16_t *x = a, *y = b; int16_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (int16_t)(uint16_t)(uint32_t)narrow_signed_binop(op, x[i], y[i], &k); }
      i32_binop(op, 0, 1, &ok);
      break;
    }
    case AGPU_I8: {
      if (op == AGPU_POW) return AGPU_EUNSUPPORTED;
      const int8_t *x = a, *y = b; int8_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (int8_t)(uint8_t)(uint32_t)narrow_signed_binop(op, x[i], y[i], &k); }
      i32_binop(op, 0, 1, &ok);
      break;
    }
    default: return AGPU_EUNSUPPORTED;
  }
  return ok ? 0 : AGPU_EUNSUPPORTED;
}

int oracle_scalar(int op, int dtype, const void* a, const void* scalar, void* out, size_t n) {
  if (op > AGPU_REM) return AGPU_EUNSUPPORTED; /* the reference has only + - * / % with a scalar */
  idx_t N = (idx_t)n;
  int ok = 1;
  switch (dtype) {
    case AGPU_F32: {
      const float* x = a; float s = *(const float*)scalar; float* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = f32_binop(op, x[i], s, &k); }
      break;
    }
    case AGPU_I32: case AGPU_DATE32: {
      const int32_t* x = a; int32_t s = *(const int32_t*)scalar; int32_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = i32_binop(op, x[i], s, &k); }
      break;
    }
    case AGPU_U32: {
      const uint32_t* x = a; uint32_t s = *(const uint32_t*)scalar; uint32_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = u32_binop(op, x[i], s, &k); }
      break;
    }
    case AGPU_U16: {
      uint16_t s16 = *(const uint16_t*)scalar;
      if (op == AGPU_ADD) {
        /* word level: arithmetic/compute_shaders/u16/scalar.wgsl:15-23.  The scalar buffer is a
         * one-element u16 array padded to a u32 word. */
        size_t nw; uint32_t* x = words_from(a, n * 2, &nw); uint32_t* o = words_out(out, n * 2, nw);
        uint32_t operand = s16;
        OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
          uint32_t operand_u16 = u16_get_left_half(operand);
          uint32_t left = u16_get_left_half(x[i]) + operand_u16;
          uint32_t right = u16_get_right_half(x[i]) + operand_u16;
          o[i] = (left & 0xffffu) + (right << 16);
        }
        WOUT_DONE(o, out, n * 2);
        WFREE(x, a);
        break;
      }
      const uint16_t* x = a; uint16_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (uint16_t)u32_binop(op, x[i], s16, &k); }
      break;
    }
    case AGPU_U8: {
      const uint8_t* x = a; uint8_t s = *(const uint8_t*)scalar; uint8_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (uint8_t)u32_binop(op, x[i], s, &k); }
      break;
    }
    case AGPU_I16: {
      const int16_t* x = a; int16_t s = *(const int16_t*)scalar; int16_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (int16_t)(uint16_t)(uint32_t)i32_binop(op, x[i], s, &k); }
      break;
    }
    case AGPU_I8: {
      const int8_t* x = a; int8_t s = *(const int8_t*)scalar; int8_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = (int8_t)(uint8_t)(uint32_t)i32_binop(op, x[i], s, &k); }
      break;
    }
    default: return AGPU_EUNSUPPORTED;
  }
  return ok ? 0 : AGPU_EUNSUPPORTED;
}

/* ------------------------------------------------------------------------------------------
 * unary
 *   neg: arithmetic/compute_shaders/f32/neg.wgsl:10-14
 *   abs/sqrt/cbrt/exp/exp2/log/log2: math/compute_shaders/f32/floatunary.wgsl:10-54
 *   i32 abs: math/compute_shaders/i32/unary.wgsl:9-13
 *   not: logical/compute_shaders/{i32,u32}/not.wgsl:9-13 (word-wise for all widths)
 *   sin/cos/acos/sinh: trigonometry/compute_shaders/f32/{trigonometry,hyperbolic}.wgsl and the
 *   fused cast+trig shaders {i8,u8,i16,u16}/{trigonometry,hyperbolic}.wgsl (word level)
 * Transcendentals are evaluated in double and rounded once to f32: the "correctly rounded"
 * value the GPU's ULP bound is measured against.
 * ---------------------------------------------------------------------------------------- */
static inline float f32_unop(int op, float x, int* ok) {
  switch (op) {
    case AGPU_NEG: return -x;
    case AGPU_ABS: return fabsf(x);
    case AGPU_SQRT: return sqrtf(x);
    case AGPU_CBRT: { /* floatunary.wgsl:46-54: pow(|x|, 1.0/3.0) with the f32 constant */
      double third = (double)(1.0f / 3.0f);
      if (x < 0.0f) return -(float)pow((double)(-x), third);
      return (float)pow((double)x, third);
    }
    case AGPU_EXP: return (float)exp((double)x);
    case AGPU_EXP2: return (float)exp2((double)x);
    case AGPU_LOG: return (float)log((double)x);
    case AGPU_LOG2: return (float)log2((double)x);
    case AGPU_SIN: return (float)sin((double)x);
    case AGPU_COS: return (float)cos((double)x);
    case AGPU_ACOS: return (float)acos((double)x);
    case AGPU_SINH: return (float)sinh((double)x);
    default: *ok = 0; return 0.0f;
  }
}

int oracle_unary(int op, int dtype, const void* a, void* out, size_t n) {
  idx_t N = (idx_t)n;
  int ok = 1;
  if (op == AGPU_NOT) {
    size_t es = dtype_size(dtype);
    if (!es || dtype == AGPU_F32 || dtype == AGPU_DATE32) return AGPU_EUNSUPPORTED;
    size_t nw; uint32_t* x = words_from(a, n * es, &nw); uint32_t* o = words_out(out, n * es, nw);
    OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) o[i] = ~x[i];
    WOUT_DONE(o, out, n * es);
    WFREE(x, a);
    return 0;
  }
  switch (dtype) {
    case AGPU_F32: {
      const float* x = a; float* o = out;
      f32_unop(op, 0, &ok);
      if (!ok) return AGPU_EUNSUPPORTED;
      OMP_FOR for (idx_t i = 0; i < N; ++i) { int k = 1; o[i] = f32_unop(op, x[i], &k); }
      return 0;
    }
    case AGPU_I32: {
      if (op != AGPU_ABS) return AGPU_EUNSUPPORTED;
      const int32_t* x = a; int32_t* o = out;
      OMP_FOR for (idx_t i = 0; i < N; ++i) o[i] = i32_abs(x[i]);
      return 0;
    }
    case AGPU_I8: case AGPU_U8: {
      if (op != AGPU_SIN && op != AGPU_COS && op != AGPU_SINH) return AGPU_EUNSUPPORTED;
      /* {i8,u8}/trigonometry.wgsl:11-33: unpack4x{I,U}8 -> f32() -> fn -> 4 x f32 */
      size_t nw; uint32_t* x = words_from(a, n, &nw);
      float* o = (n % 4 == 0) ? (float*)out : (float*)calloc(nw * 4 + 1, 4);
      OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
        int k = 1;
        if (dtype == AGPU_I8) {
          int32_t u[4]; unpack4xI8(x[i], u);
          for (int j = 0; j < 4; ++j) o[i * 4 + j] = f32_unop(op, (float)u[j], &k);
        } else {
          uint32_t u[4]; unpack4xU8(x[i], u);
          for (int j = 0; j < 4; ++j) o[i * 4 + j] = f32_unop(op, (float)u[j], &k);
        }
      }
      WOUT_DONE(o, out, n * 4);
      WFREE(x, a);
      return 0;
    }
    case AGPU_I16: case AGPU_U16: {
      if (op != AGPU_SIN && op != AGPU_COS && op != AGPU_SINH) return AGPU_EUNSUPPORTED;
      /* {i16,u16}/trigonometry.wgsl: get_left_half/get_right_half -> f32() -> fn */
      size_t nw; uint32_t* x = words_from(a, n * 2, &nw);
      float* o = (n % 2 == 0) ? (float*)out : (float*)calloc(nw * 2 + 1, 4);
      OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
        int k = 1;
        if (dtype == AGPU_I16) {
          o[i * 2] = f32_unop(op, (float)i16_get_left_half((int32_t)x[i]), &k);
          o[i * 2 + 1] = f32_unop(op, (float)i16_get_right_half((int32_t)x[i]), &k);
        } else {
          o[i * 2] = f32_unop(op, (float)u16_get_left_half(x[i]), &k);
          o[i * 2 + 1] = f32_unop(op, (float)u16_get_right_half(x[i]), &k);
        }
      }
      WOUT_DONE(o, out, n * 4);
      WFREE(x, a);
      return 0;
    }
    default: return AGPU_EUNSUPPORTED;
  }
}

/* ------------------------------------------------------------------------------------------
 * compare -> bitmap
 *   {f32,i32,u32}/cmp.wgsl:22-75: one element per invocation, bit set in a workgroup-shared
 *   word with atomicOr, lane gid%32==0 stores word gid/32 (LSB first)
 *   {i16,u16}/cmp.wgsl:23-91: two halves per invocation;  {i8,u8}/cmp.wgsl:23-111: four bytes
 * The atomics only assemble bits, so the restatement sets bit i = predicate(a[i], b[i]);
 * sub-word operands go through the shader's own extract helpers.  Bits >= n are zero (the
 * reference leaves whatever zero-padded out-of-range lanes produce there: SURVEY.md Q4).
 * ---------------------------------------------------------------------------------------- */
#define CMP(op, x, y) ((op) == AGPU_GT ? (x) > (y) : (op) == AGPU_GTEQ ? (x) >= (y) : \
                       (op) == AGPU_LT ? (x) < (y) : (op) == AGPU_LTEQ ? (x) <= (y) : (x) == (y))

int oracle_compare(int op, int dtype, const void* a, const void* b, uint32_t* out_bits, size_t n) {
  if (op < AGPU_GT || op > AGPU_EQ) return AGPU_EUNSUPPORTED;
  size_t nbw = bit_words(n);
  memset(out_bits, 0, nbw * 4);
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  idx_t NW = (idx_t)nbw;
  if (es == 4) {
    OMP_FOR for (idx_t w = 0; w < NW; ++w) {
      uint32_t bits = 0;
      for (size_t i = (size_t)w * 32; i < (size_t)w * 32 + 32 && i < n; ++i) {
        int p;
        if (dtype == AGPU_F32) { float x = ((const float*)a)[i], y = ((const float*)b)[i]; p = CMP(op, x, y); }
        else if (dtype == AGPU_U32) { uint32_t x = ((const uint32_t*)a)[i], y = ((const uint32_t*)b)[i]; p = CMP(op, x, y); }
        else { int32_t x = ((const int32_t*)a)[i], y = ((const int32_t*)b)[i]; p = CMP(op, x, y); }
        bits |= (uint32_t)p << (i & 31);
      }
      out_bits[w] = bits;
    }
    return 0;
  }
  size_t nw; uint32_t* x = words_from(a, n * es, &nw); uint32_t* y = words_from(b, n * es, &nw);
  size_t per = 4 / es; /* elements per packed word */
  OMP_FOR for (idx_t w = 0; w < NW; ++w) {
    uint32_t bits = 0;
    for (size_t i = (size_t)w * 32; i < (size_t)w * 32 + 32 && i < n; ++i) {
      uint32_t xw = x[i / per], yw = y[i / per];
      size_t lane = i % per;
      int p;
      if (dtype == AGPU_I8) {
        int32_t l = lane == 0 ? i8_get_left_byte((int32_t)xw) : lane == 1 ? i8_get_mid_left_byte((int32_t)xw)
                  : lane == 2 ? i8_get_mid_right_byte((int32_t)xw) : i8_get_right_byte((int32_t)xw);
        int32_t r = lane == 0 ? i8_get_left_byte((int32_t)yw) : lane == 1 ? i8_get_mid_left_byte((int32_t)yw)
                  : lane == 2 ? i8_get_mid_right_byte((int32_t)yw) : i8_get_right_byte((int32_t)yw);
        p = CMP(op, l, r);
      } else if (dtype == AGPU_U8) {
        uint32_t l = lane == 0 ? u8_get_left_byte(xw) : lane == 1 ? u8_get_mid_left_byte(xw)
                   : lane == 2 ? u8_get_mid_right_byte(xw) : u8_get_right_byte(xw);
        uint32_t r = lane == 0 ? u8_get_left_byte(yw) : lane == 1 ? u8_get_mid_left_byte(yw)
                   : lane == 2 ? u8_get_mid_right_byte(yw) : u8_get_right_byte(yw);
        p = CMP(op, l, r);
      } else if (dtype == AGPU_I16) {
        int32_t l = lane == 0 ? i16_get_left_half((int32_t)xw) : i16_get_right_half((int32_t)xw);
        int32_t r = lane == 0 ? i16_get_left_half((int32_t)yw) : i16_get_right_half((int32_t)yw);
        p = CMP(op, l, r);
      } else {
        uint32_t l = lane == 0 ? u16_get_left_half(xw) : u16_get_right_half(xw);
        uint32_t r = lane == 0 ? u16_get_left_half(yw) : u16_get_right_half(yw);
        p = CMP(op, l, r);
      }
      bits |= (uint32_t)p << (i & 31);
    }
    out_bits[w] = bits;
  }
  WFREE(x, a); WFREE(y, b);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * shifts (count = one u32 per row)
 *   {i32,u32}/shift.wgsl:13-23; {i16,u16}/shift.wgsl (halves, i16 `shr` helper :31-41);
 *   {i8,u8}/shift.wgsl:13-45 (unpack4x -> vec4 shift -> pack4x)
 * WGSL takes the count modulo 32 on the 32-bit widened lane.  The i16 `shr` helper is evaluated
 * literally for EVERY count: its `16u - shift_value` wraps for counts > 16 and the shifts take
 * their counts modulo 32 like any other WGSL shift.  (Counts 0..15 are pinned by the reference's
 * vectors; for the rest the helper still equals the arithmetic shift of the sign-extended half by
 * `count & 31` — tests/test_oracle_properties.py checks that over the whole range — but no
 * reference vector covers it: SURVEY.md Q17.)
 * ---------------------------------------------------------------------------------------- */
static inline int32_t i16_shr_helper(int32_t input, uint32_t shift_value) {
  /* logical/compute_shaders/i16/shift.wgsl:31-41 */
  if (input < 0) {
    int32_t result = i32_shr(input, shift_value);
    int32_t other = i32_shl(0xffff, 16u - shift_value);
    other = other | i32_shr(0x00008000, shift_value);
    return other | result;
  }
  return i32_shr(input, shift_value);
}

int oracle_shift(int op, int dtype, const void* a, const uint32_t* counts, void* out, size_t n) {
  if (op != AGPU_SHL && op != AGPU_SHR) return AGPU_EUNSUPPORTED;
  idx_t N = (idx_t)n;
  if (dtype == AGPU_I32) {
    const int32_t* x = a; int32_t* o = out;
    OMP_FOR for (idx_t i = 0; i < N; ++i) o[i] = op == AGPU_SHL ? i32_shl(x[i], counts[i]) : i32_shr(x[i], counts[i]);
    return 0;
  }
  if (dtype == AGPU_U32) {
    const uint32_t* x = a; uint32_t* o = out;
    OMP_FOR for (idx_t i = 0; i < N; ++i) o[i] = op == AGPU_SHL ? u32_shl(x[i], counts[i]) : u32_shr(x[i], counts[i]);
    return 0;
  }
  size_t es = dtype_size(dtype);
  if (es != 1 && es != 2) return AGPU_EUNSUPPORTED;
  size_t per = 4 / es;
  size_t nw; uint32_t* x = words_from(a, n * es, &nw);
  const uint32_t* cp = counts; /* counts padded to whole words of lanes only when the tail is ragged */
  uint32_t* cpad = NULL;
  if (nw * per != n) {
    cpad = (uint32_t*)calloc(nw * per + 1, 4);
    memcpy(cpad, counts, n * 4);
    cp = cpad;
  }
  uint32_t* o = words_out(out, n * es, nw);
  OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
    const uint32_t* r = cp + (size_t)i * per;
    if (dtype == AGPU_I8) {
      int32_t l[4]; unpack4xI8(x[i], l);
      for (int k = 0; k < 4; ++k) l[k] = op == AGPU_SHL ? i32_shl(l[k], r[k]) : i32_shr(l[k], r[k]);
      o[i] = pack4xI8(l);
    } else if (dtype == AGPU_U8) {
      uint32_t l[4]; unpack4xU8(x[i], l);
      for (int k = 0; k < 4; ++k) l[k] = op == AGPU_SHL ? u32_shl(l[k], r[k]) : u32_shr(l[k], r[k]);
      o[i] = pack4xU8(l);
    } else if (dtype == AGPU_I16) {
      int32_t lh = i16_get_left_half((int32_t)x[i]), rh = i16_get_right_half((int32_t)x[i]);
      int32_t lo, hi;
      if (op == AGPU_SHL) { lo = i32_shl(lh, r[0]); hi = i32_shl(rh, r[1]); }
      else { lo = i16_shr_helper(lh, r[0]); hi = i16_shr_helper(rh, r[1]); }
      o[i] = (uint32_t)((lo & 0xffff) | i32_shl(hi & 0xffff, 16));
    } else {
      uint32_t lh = u16_get_left_half(x[i]), rh = u16_get_right_half(x[i]);
      uint32_t lo = op == AGPU_SHL ? u32_shl(lh, r[0]) : u32_shr(lh, r[0]);
      uint32_t hi = op == AGPU_SHL ? u32_shl(rh, r[1]) : u32_shr(rh, r[1]);
      o[i] = (lo & 0x0000ffffu) | ((hi << 16) & 0xffff0000u);
    }
  }
  WOUT_DONE(o, out, n * es);
  WFREE(x, a); free(cpad);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * bitmap logical ops: logical/src/boolean.rs:45-75 on u32/logical.wgsl and u32/not.wgsl.
 * Padding bits of the last word are cleared (the reference's `not` flips them: Q5).
 * ---------------------------------------------------------------------------------------- */
int oracle_bitmap_binary(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n_bits) {
  if (op != AGPU_AND && op != AGPU_OR && op != AGPU_XOR) return AGPU_EUNSUPPORTED;
  idx_t nw = (idx_t)bit_words(n_bits);
  OMP_FOR for (idx_t i = 0; i < nw; ++i)
    out[i] = op == AGPU_AND ? (a[i] & b[i]) : op == AGPU_OR ? (a[i] | b[i]) : (a[i] ^ b[i]);
  if (nw) mask_tail(out, n_bits);
  return 0;
}
int oracle_bitmap_not(const uint32_t* a, uint32_t* out, size_t n_bits) {
  idx_t nw = (idx_t)bit_words(n_bits);
  OMP_FOR for (idx_t i = 0; i < nw; ++i) out[i] = ~a[i];
  if (nw) mask_tail(out, n_bits);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * casts — matrix of cast/src/lib.rs:135-161
 *   i8:  cast_{i32,i16,f32}.wgsl (unpack4xI8)       u8: cast_{u32,u16,f32}.wgsl (unpack4xU8)
 *   i16: cast_{i32,f32}.wgsl:13-21 (halves)          u16: cast_{u32,f32}.wgsl:13-21
 *   f32->u8: f32/cast_u8.wgsl:9-23                   bool->f32: boolean/cast_f32.wgsl:9-20
 *   same-width and the unsigned targets of signed sources reuse the signed shader's words
 *   (cast/src/i8_cast.rs:19-44, i16_cast.rs:24-31); same-width = buffer copy (lib.rs:69-86)
 * ---------------------------------------------------------------------------------------- */
int oracle_cast(int src, int dst, const void* a, void* out, size_t n) {
  idx_t N = (idx_t)n;
  if (src == AGPU_BOOL) {
    if (dst != AGPU_F32) return AGPU_EUNSUPPORTED;
    const uint32_t* bits = a; float* o = out;
    OMP_FOR for (idx_t i = 0; i < N; ++i) {
      uint32_t bit_pos = 1u << ((uint32_t)i % 32u);
      o[i] = (bits[i / 32] & bit_pos) == bit_pos ? 1.0f : 0.0f; /* zero-initialised otherwise */
    }
    return 0;
  }
  if (src == AGPU_F32) {
    if (dst != AGPU_U8) return AGPU_EUNSUPPORTED;
    size_t nw = (n + 3) / 4;
    const float* x = a;
    float* xpad = NULL;
    if (n % 4) { xpad = (float*)calloc(nw * 4 + 1, 4); memcpy(xpad, a, n * 4); x = xpad; }
    uint32_t* o = words_out(out, n, nw);
    OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
      size_t idx = 4 * (size_t)i;
      uint32_t w = 0;
      w |= f32_to_u32(x[idx]) % 256u;
      w |= (f32_to_u32(x[idx + 1]) % 256u) << 8;
      w |= (f32_to_u32(x[idx + 2]) % 256u) << 16;
      w |= (f32_to_u32(x[idx + 3]) % 256u) << 24;
      o[i] = w;
    }
    WOUT_DONE(o, out, n);
    free(xpad);
    return 0;
  }
  size_t ss = dtype_size(src), ds = dtype_size(dst);
  if (!ss || !ds) return AGPU_EUNSUPPORTED;
  int src_signed = src == AGPU_I8 || src == AGPU_I16;
  int src_sub = src == AGPU_I8 || src == AGPU_U8 || src == AGPU_I16 || src == AGPU_U16;
  if (!src_sub) return AGPU_EUNSUPPORTED;
  /* allowed targets (cast/src/lib.rs:139-158) */
  int allowed = 0;
  if (src == AGPU_I8) allowed = dst == AGPU_U8 || dst == AGPU_U16 || dst == AGPU_U32 || dst == AGPU_I16 || dst == AGPU_I32 || dst == AGPU_F32;
  if (src == AGPU_I16) allowed = dst == AGPU_I32 || dst == AGPU_U16 || dst == AGPU_U32 || dst == AGPU_F32;
  if (src == AGPU_U8) allowed = dst == AGPU_U16 || dst == AGPU_U32 || dst == AGPU_I8 || dst == AGPU_I16 || dst == AGPU_I32 || dst == AGPU_F32;
  if (src == AGPU_U16) allowed = dst == AGPU_U32 || dst == AGPU_I16 || dst == AGPU_I32 || dst == AGPU_F32;
  if (!allowed) return AGPU_EUNSUPPORTED;
  if (ss == ds) { memcpy(out, a, n * ss); return 0; } /* impl_cast!($into, $from): clone_buffer */
  size_t nw; uint32_t* x = words_from(a, n * ss, &nw);
  size_t per = 4 / ss;
  size_t out_words = nw * per * ds / 4;
  uint32_t* o = words_out(out, n * ds, out_words);
  OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
    int32_t lanes[4];
    if (ss == 1) {
      if (src_signed) unpack4xI8(x[i], lanes);
      else { uint32_t u[4]; unpack4xU8(x[i], u); for (int k = 0; k < 4; ++k) lanes[k] = (int32_t)u[k]; }
    } else {
      if (src_signed) { lanes[0] = i16_get_left_half((int32_t)x[i]); lanes[1] = i16_get_right_half((int32_t)x[i]); }
      else { lanes[0] = (int32_t)u16_get_left_half(x[i]); lanes[1] = (int32_t)u16_get_right_half(x[i]); }
    }
    if (dst == AGPU_F32) {
      float* fo = (float*)o;
      for (size_t k = 0; k < per; ++k) fo[(size_t)i * per + k] = (float)lanes[k];
    } else if (ds == 4) {
      for (size_t k = 0; k < per; ++k) o[(size_t)i * per + k] = (uint32_t)lanes[k];
    } else { /* 8 -> 16 bit: cast_i16.wgsl / cast_u16.wgsl:13-18 */
      size_t np = (size_t)i * 2;
      o[np] = ((uint32_t)lanes[0] & 0x0000ffffu) | ((uint32_t)lanes[1] << 16);
      o[np + 1] = ((uint32_t)lanes[2] & 0x0000ffffu) | ((uint32_t)lanes[3] << 16);
    }
  }
  WOUT_DONE(o, out, n * ds);
  WFREE(x, a);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * merge (mask select): routines/compute_shaders/{32bit,16bit,8bit,bool}/merge.wgsl and
 * merge validity routines/src/merge.rs:17-86 on u32/merge_null_buffer.wgsl:13-35
 * ---------------------------------------------------------------------------------------- */
int oracle_merge(int dtype, const void* a, const void* b, const uint32_t* mask, void* out, size_t n) {
  if (dtype == AGPU_BOOL) { /* bool/merge.wgsl:17-21 */
    const uint32_t *x = a, *y = b; uint32_t* o = out;
    idx_t nw = (idx_t)bit_words(n);
    OMP_FOR for (idx_t i = 0; i < nw; ++i) o[i] = (x[i] & mask[i]) | (y[i] & ~mask[i]);
    if (nw) mask_tail(o, n);
    return 0;
  }
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  size_t nw; uint32_t* x = words_from(a, n * es, &nw); uint32_t* y = words_from(b, n * es, &nw);
  size_t per = 4 / es;
  size_t mw = bit_words(nw * per);
  const uint32_t* m = mask;
  uint32_t* mpad = NULL;
  if (nw * per != n) { mpad = words_zero(mw + 1); memcpy(mpad, mask, bit_words(n) * 4); m = mpad; }
  uint32_t* o = words_out(out, n * es, nw);
  OMP_FOR for (idx_t i = 0; i < (idx_t)nw; ++i) {
    if (es == 4) { /* 32bit/merge.wgsl:22-30 */
      o[i] = get_bit(m, (size_t)i) ? x[i] : y[i];
    } else if (es == 2) { /* 16bit/merge.wgsl:22-36 */
      size_t p = (size_t)i * 2;
      uint32_t w = get_bit(m, p) ? u16_get_left_half(x[i]) : u16_get_left_half(y[i]);
      w |= get_bit(m, p + 1) ? (x[i] & 0xffff0000u) : (y[i] & 0xffff0000u);
      o[i] = w;
    } else { /* 8bit/merge.wgsl:22-46 */
      size_t p = (size_t)i * 4;
      uint32_t w = get_bit(m, p) ? u8_get_left_byte(x[i]) : u8_get_left_byte(y[i]);
      w |= get_bit(m, p + 1) ? (x[i] & 0x0000ff00u) : (y[i] & 0x0000ff00u);
      w |= get_bit(m, p + 2) ? (x[i] & 0x00ff0000u) : (y[i] & 0x00ff0000u);
      w |= get_bit(m, p + 3) ? (x[i] & 0xff000000u) : (y[i] & 0xff000000u);
      o[i] = w;
    }
  }
  WOUT_DONE(o, out, n * es);
  WFREE(x, a); WFREE(y, b); free(mpad);
  return 0;
}

int oracle_merge_validity(const uint32_t* va, const uint32_t* vb, const uint32_t* mask,
                          const uint32_t* vmask, uint32_t* vout, size_t n) {
  if (!va && !vb && !vmask) return AGPU_EINVAL; /* (None, None, None) => None */
  idx_t nw = (idx_t)bit_words(n);
  OMP_FOR for (idx_t i = 0; i < nw; ++i) {
    uint32_t ones = 0xFFFFFFFFu;
    uint32_t sel, notsel;
    if (g_quirks) {
      /* literal merge.rs:28-68: a missing operand bitmap drops its whole term */
      uint32_t w = 0; int any = 0;
      if (va) { w |= va[i] & mask[i]; any = 1; }
      if (vb) { w |= vb[i] & ~mask[i]; any = 1; }
      if (!any) w = ones;
      vout[i] = vmask ? (w & vmask[i]) : w;
      continue;
    }
    sel = (va ? va[i] : ones) & mask[i];      /* merge_selected */
    notsel = (vb ? vb[i] : ones) & ~mask[i];  /* merge_not_selected */
    vout[i] = (sel | notsel) & (vmask ? vmask[i] : ones); /* merge_or, merge_nulls */
  }
  if (nw) mask_tail(vout, n);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * take / put: routines/compute_shaders/32bit/take.wgsl:13-17, bool/take.wgsl:13-33,
 * 32bit/put.wgsl:17-23, bool/put.wgsl:17-34.  Out-of-range source indices read zero
 * (robust buffer access).  8/16-bit take is new surface ("todo!()" in the reference).
 * ---------------------------------------------------------------------------------------- */
int oracle_take(int dtype, const void* src, size_t src_len, const uint32_t* idx, void* out, size_t m) {
  idx_t M = (idx_t)m;
  if (dtype == AGPU_BOOL) {
    const uint32_t* s = src; uint32_t* o = out;
    idx_t nw = (idx_t)bit_words(m);
    OMP_FOR for (idx_t w = 0; w < nw; ++w) {
      uint32_t start_index = (uint32_t)w * 32u, result = 0;
      for (uint32_t i = 0; i < 32u && (size_t)start_index + i < m; ++i) {
        uint32_t index = idx[start_index + i];
        uint32_t base_src_index = index / 32u, src_index = index % 32u;
        uint32_t word = (size_t)index < src_len ? s[base_src_index] : 0u;
        uint32_t value = word & (1u << src_index);
        if (src_index > i) value >>= (src_index - i); else value <<= (i - src_index);
        result |= value;
      }
      o[w] = result;
    }
    return 0;
  }
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  const uint8_t* s = src; uint8_t* o = out;
  OMP_FOR for (idx_t j = 0; j < M; ++j) {
    if ((size_t)idx[j] < src_len) memcpy(o + (size_t)j * es, s + (size_t)idx[j] * es, es);
    else memset(o + (size_t)j * es, 0, es);
  }
  return 0;
}

int oracle_put(int dtype, const void* src, const uint32_t* src_idx, void* dst, const uint32_t* dst_idx, size_t m) {
  if (dtype == AGPU_BOOL) {
    const uint32_t* s = src; uint32_t* d = dst;
    for (size_t i = 0; i < m; ++i) { /* sequential: last writer wins deterministically */
      uint32_t bs = src_idx[i] / 32u, bd = dst_idx[i] / 32u, si = src_idx[i] % 32u, di = dst_idx[i] % 32u;
      uint32_t value = s[bs] & (1u << si);
      if (si > di) value >>= (si - di); else value <<= (di - si);
      d[bd] &= ~(1u << di);
      d[bd] |= value;
    }
    return 0;
  }
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  const uint8_t* s = src; uint8_t* d = dst;
  for (size_t i = 0; i < m; ++i) memcpy(d + (size_t)dst_idx[i] * es, s + (size_t)src_idx[i] * es, es);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * filter / compaction — NOT in the reference (parity unpinned).  Defined as
 * pyarrow.compute.filter(null_selection_behavior='drop'): keep row i iff mask bit i is 1 and
 * (no mask validity or its bit i is 1); order preserved; validity bits compacted alongside.
 * ---------------------------------------------------------------------------------------- */
int oracle_filter(int dtype, const void* src, const uint32_t* vsrc, const uint32_t* mask,
                  const uint32_t* vmask, size_t n, void* out, uint32_t* vout, uint64_t* count) {
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  const uint8_t* s = src; uint8_t* o = out;
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!get_bit(mask, i) || (vmask && !get_bit(vmask, i))) continue;
    memcpy(o + k * es, s + i * es, es);
    if (vsrc && vout) {
      if ((k & 31) == 0) vout[k >> 5] = 0;
      vout[k >> 5] |= (uint32_t)get_bit(vsrc, i) << (k & 31);
    }
    ++k;
  }
  *count = k;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * broadcast: array/compute_shaders/{f32,i32,u32}/broadcast.wgsl:9-13 (correct for negative
 * i8/i16: SURVEY.md Q2)
 * ---------------------------------------------------------------------------------------- */
int oracle_broadcast(int dtype, const void* scalar, void* out, size_t n) {
  size_t es = dtype_size(dtype);
  if (!es) return AGPU_EUNSUPPORTED;
  uint8_t* o = out;
  for (size_t i = 0; i < n; ++i) memcpy(o + i * es, scalar, es);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * sum: arithmetic/compute_shaders/{f32,i32,u32}/aggregate.wgsl:13-42 driven by
 * arithmetic/src/aggregate_kernels.rs:24-52 — per 256-element workgroup a pairwise tree
 * (s = 1, 2, 4 ... 128: element 2*s*k += element 2*s*k + s), one partial per workgroup,
 * passes repeated until one value is left.  Out-of-range lanes contribute 0.
 * ---------------------------------------------------------------------------------------- */
#define SUM_PASS(T, ADD)                                                        \
  static size_t sum_pass_##T(const T* in, size_t len, T* outp) {               \
    size_t groups = (len + 255) / 256;                                          \
    OMP_FOR for (idx_t g = 0; g < (idx_t)groups; ++g) {                         \
      T sh[256];                                                                \
      for (size_t l = 0; l < 256; ++l) {                                        \
        size_t gi = (size_t)g * 256 + l;                                        \
        sh[l] = gi < len ? in[gi] : (T)0;                                       \
      }                                                                         \
      for (size_t s = 1; s < 256; s *= 2)                                       \
        for (size_t l = 0; l < 256; ++l) {                                      \
          size_t index = 2 * s * l;                                             \
          if (index < 256 && index + s < 256) sh[index] = ADD(sh[index], sh[index + s]); \
        }                                                                       \
      outp[g] = sh[0];                                                          \
    }                                                                           \
    return groups;                                                              \
  }
#define ADD_F(a, b) ((a) + (b))
#define ADD_U(a, b) ((uint32_t)((a) + (b)))
SUM_PASS(float, ADD_F)
SUM_PASS(uint32_t, ADD_U)

int oracle_sum(int dtype, const void* a, size_t n, void* out) {
  if (dtype != AGPU_F32 && dtype != AGPU_I32 && dtype != AGPU_U32) return AGPU_EUNSUPPORTED;
  /* an empty column: the reference's `while new_length != 1` never terminates for len 0
   * (aggregate_kernels.rs:26-44: 0.div_ceil(256) stays 0); the defined result here is 0 */
  if (n == 0) { memset(out, 0, 4); return 0; }
  size_t cap = (n + 255) / 256 + 1;
  void* t0 = calloc(cap, 4); void* t1 = calloc(cap, 4);
  size_t len;
  if (dtype == AGPU_F32) {
    len = sum_pass_float(a, n, t0);
    while (len != 1) { len = sum_pass_float(t0, len, t1); void* t = t0; t0 = t1; t1 = t; }
  } else { /* i32 wraps exactly like u32 */
    len = sum_pass_uint32_t(a, n, t0);
    while (len != 1) { len = sum_pass_uint32_t(t0, len, t1); void* t = t0; t0 = t1; t1 = t; }
  }
  memcpy(out, t0, 4);
  free(t0); free(t1);
  return 0;
}

/* any / all: logical/compute_shaders/u32/any.wgsl:11-21, countbitones.wgsl:9-15 + Sum,
 * logical/src/boolean.rs:106-147.  `all` counts the first n_bits only (Q5). */
int oracle_any(const uint32_t* bits, size_t n_bits, uint32_t* result) {
  size_t nw = bit_words(n_bits); uint32_t r = 0;
  for (size_t i = 0; i < nw; ++i) {
    uint32_t w = bits[i];
    if (i == nw - 1 && (n_bits & 31)) w &= (1u << (n_bits & 31)) - 1u;
    if (w > 0u) r += 1;
  }
  *result = r > 0;
  return 0;
}
int oracle_all(const uint32_t* bits, size_t n_bits, uint32_t* result) {
  size_t nw = bit_words(n_bits); uint64_t total = 0;
  for (size_t i = 0; i < nw; ++i) {
    uint32_t w = bits[i];
    if (i == nw - 1 && (n_bits & 31)) w &= (1u << (n_bits & 31)) - 1u;
    total += (uint64_t)__builtin_popcount(w);
  }
  *result = total == (uint64_t)n_bits;
  return 0;
}

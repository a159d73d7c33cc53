// unary.cu — agpu_unary: neg / abs / not, f32 math, trig (with the int->f32 cast fused).
#include "elementwise.cuh"
#include "ops.cuh"

namespace {

template <typename TI, typename TO, class F>
int run_unary(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  UnaryOp<TI, TO, F> op{(const TI*)a, (TO*)out, F{}};
  return launch_ew(dev, op, n, bm, aligned16(a) && aligned16(out));
}

// sin / cos / sinh on an integer column -> f32 column
template <typename TI>
int trig_int(agpu_device* dev, int op, const void* a, void* out, size_t n, const BmAnd& bm) {
  switch (op) {
    case AGPU_SIN: return run_unary<TI, float, FSin<TI>>(dev, a, out, n, bm);
    case AGPU_COS: return run_unary<TI, float, FCos<TI>>(dev, a, out, n, bm);
    case AGPU_SINH: return run_unary<TI, float, FSinh<TI>>(dev, a, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

template <typename T>
int not_int(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  return run_unary<T, T, OpNot<T>>(dev, a, out, n, bm);
}

}  // namespace

extern "C" int agpu_unary(agpu_device* dev, int op, int dtype, const void* a, void* out, size_t n,
                          const uint32_t* va, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !out)) return AGPU_EINVAL;
  if (vout && !va) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, nullptr, nullptr, nullptr, vout);
  if (op == AGPU_NOT) {
    switch (dtype) {
      case AGPU_I32: return not_int<int32_t>(dev, a, out, n, bm);
      case AGPU_U32: return not_int<uint32_t>(dev, a, out, n, bm);
      case AGPU_I16: return not_int<int16_t>(dev, a, out, n, bm);
      case AGPU_U16: return not_int<uint16_t>(dev, a, out, n, bm);
      case AGPU_I8: return not_int<int8_t>(dev, a, out, n, bm);
      case AGPU_U8: return not_int<uint8_t>(dev, a, out, n, bm);
      default: return AGPU_EUNSUPPORTED;
    }
  }
  switch (dtype) {
    case AGPU_F32:
      switch (op) {
        case AGPU_NEG: return run_unary<float, float, OpNeg<float>>(dev, a, out, n, bm);
        case AGPU_ABS: return run_unary<float, float, OpAbs<float>>(dev, a, out, n, bm);
        case AGPU_SQRT: return run_unary<float, float, FSqrt<float>>(dev, a, out, n, bm);
        case AGPU_CBRT: return run_unary<float, float, FCbrt<float>>(dev, a, out, n, bm);
        case AGPU_EXP: return run_unary<float, float, FExp<float>>(dev, a, out, n, bm);
        case AGPU_EXP2: return run_unary<float, float, FExp2<float>>(dev, a, out, n, bm);
        case AGPU_LOG: return run_unary<float, float, FLog<float>>(dev, a, out, n, bm);
        case AGPU_LOG2: return run_unary<float, float, FLog2<float>>(dev, a, out, n, bm);
        case AGPU_SIN: return run_unary<float, float, FSin<float>>(dev, a, out, n, bm);
        case AGPU_COS: return run_unary<float, float, FCos<float>>(dev, a, out, n, bm);
        case AGPU_ACOS: return run_unary<float, float, FAcos<float>>(dev, a, out, n, bm);
        case AGPU_SINH: return run_unary<float, float, FSinh<float>>(dev, a, out, n, bm);
        default: return AGPU_EUNSUPPORTED;
      }
    case AGPU_I32:
      if (op == AGPU_ABS) return run_unary<int32_t, int32_t, OpAbs<int32_t>>(dev, a, out, n, bm);
      return AGPU_EUNSUPPORTED;
    case AGPU_I16: return trig_int<int16_t>(dev, op, a, out, n, bm);
    case AGPU_U16: return trig_int<uint16_t>(dev, op, a, out, n, bm);
    case AGPU_I8: return trig_int<int8_t>(dev, op, a, out, n, bm);
    case AGPU_U8: return trig_int<uint8_t>(dev, op, a, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

// cdata_shim.cpp — a tiny C entry point over c_data_interface.hpp so that pyarrow (the only Arrow
// producer in this image) can drive import -> device compute -> export through the Arrow C Data
// Interface:  out = op(in_a [, in_b])  with op in {"identity", "add", "gt", "sqrt", "filter"}.
#include <cstring>

#include "c_data_interface.hpp"

using namespace arrow_gpu;

extern "C" int agpu_cdata_apply(const char* op, const ArrowSchema* sa, const ArrowArray* aa, const ArrowSchema* sb,
                                const ArrowArray* ab, ArrowSchema* out_schema, ArrowArray* out_array) {
  try {
    static DevicePtr dev = std::make_shared<GpuDevice>(0);
    ArrowArrayGPU a = import_arrow(sa, aa, dev);
    const std::string o(op);
    if (o == "identity") { export_arrow(a, out_schema, out_array); return 0; }
    if (o == "sqrt") { export_arrow(sqrt_dyn(a), out_schema, out_array); return 0; }
    ArrowArrayGPU b = import_arrow(sb, ab, dev);
    if (o == "add") { export_arrow(add_dyn(a, b), out_schema, out_array); return 0; }
    if (o == "gt") { export_arrow(gt_dyn(a, b), out_schema, out_array); return 0; }
    if (o == "filter") {
      const auto& mask = try_from<BooleanArrayGPU>(b);
      ArrowArrayGPU r = std::visit([&](const auto& x) -> ArrowArrayGPU {
        if constexpr (detail::is_bool<std::decay_t<decltype(x)>>) throw Panic("filter on bool");
        else return x.filter(mask);
      }, a);
      export_arrow(r, out_schema, out_array);
      return 0;
    }
    return -1;
  } catch (const std::exception&) {
    return -2;
  }
}

// The same through the Arrow C Device Data Interface, zero-copy in both directions:
// `in` (CUDA memory of device 0, e.g. exported by arrow_gpu_b200/c_device.py) is MOVED into the
// mirror, out = op(in [, in]) stays on the device and is handed back as an ArrowDeviceArray.
extern "C" int agpu_cdata_device_apply(const char* op, const ArrowSchema* schema, ArrowDeviceArray* in, ArrowSchema* out_schema,
                                       ArrowDeviceArray* out) {
  try {
    static DevicePtr dev = std::make_shared<GpuDevice>(0);
    ArrowArrayGPU a = import_arrow_device(schema, in, dev);
    const std::string o(op);
    if (o == "identity") { export_arrow_device(a, out_schema, out); return 0; }
    if (o == "add") { export_arrow_device(add_dyn(a, a), out_schema, out); return 0; }
    if (o == "gt") { export_arrow_device(gt_dyn(add_dyn(a, a), a), out_schema, out); return 0; }
    return -1;
  } catch (const std::exception&) {
    return -2;
  }
}

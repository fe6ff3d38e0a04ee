// core_scalar_type.cpp — `_core_C.ScalarType`, the sub-byte dtype descriptor that crosses the op boundary
// as the `b_q_type` argument of gptq_marlin_gemm (kernels/torch_bindings.cpp:195-201 of the reference).
//
// Own implementation of the interface the reference exposes from kernels/core/scalar_type.hpp:466-507
// (constructor (exponent, mantissa, bias, signed), read-only properties, predicates, min/max, the
// int_/uint/float_IEEE754/float_ static constructors, __eq__/__str__/__repr__/__len__/__obj_flatten__).
// Built as `_core_C.abi3.so`; it is loaded only when no `_core_C.ScalarType` class is registered yet, so an
// installation that keeps the reference's own `_core_C` extension works unchanged: the marlin op in
// torch_shim.cpp reads `size_bits` / `bias` through the class's registered property getters, never through
// a C++ type.
#include <Python.h>

#include <torch/custom_class.h>
#include <torch/library.h>

#include <cmath>
#include <string>
#include <tuple>

namespace b200 {

struct ScalarTypeTorch : public torch::CustomClassHolder {
  int64_t exponent, mantissa, bias;
  bool is_signed_, finite_values_only;
  int64_t nan_repr;  // 0 none, 1 IEEE-754, 2 extended range (max/min encodings are NaN)

  ScalarTypeTorch(int64_t e, int64_t m, int64_t b, bool s, bool fvo = false, int64_t nr = 1)
      : exponent(e), mantissa(m), bias(b), is_signed_(s), finite_values_only(fvo), nan_repr(nr) {}

  int64_t size_bits() const { return exponent + mantissa + (is_signed_ ? 1 : 0); }
  bool is_signed() const { return is_signed_; }
  bool is_integer() const { return exponent == 0; }
  bool is_floating_point() const { return exponent > 0; }
  bool has_bias() const { return bias != 0; }
  bool has_nans() const { return is_floating_point() && nan_repr != 0; }
  bool has_infs() const { return is_floating_point() && !finite_values_only; }
  bool is_ieee_754() const { return is_floating_point() && !finite_values_only && nan_repr == 1; }

  c10::IValue max() const {
    if (is_integer()) return c10::IValue((int64_t)((1ll << mantissa) - 1 - bias));
    // largest finite value of a (possibly non-IEEE) float format
    int64_t max_mant = (1ll << mantissa) - 1;
    int64_t max_exp = (1ll << exponent) - 1;
    if (has_infs() || nan_repr == 1) max_exp -= 1;            // top exponent reserved
    else if (nan_repr == 2) max_mant -= 1;                    // only the all-ones pattern is NaN
    const int64_t ebias = (1ll << (exponent - 1)) - 1;
    const double v = std::ldexp(1.0 + (double)max_mant / (double)(1ll << mantissa), (int)(max_exp - ebias));
    return c10::IValue(v);
  }
  c10::IValue min() const {
    if (is_integer()) {
      const int64_t lo = is_signed_ ? -(1ll << mantissa) : 0;
      return c10::IValue((int64_t)(lo - bias));
    }
    return c10::IValue(-max().toDouble());
  }
  std::string str() const {
    if (is_floating_point()) {
      std::string r = "float" + std::to_string(size_bits()) + "_e" + std::to_string(exponent) + "m" +
                      std::to_string(mantissa);
      if (!is_ieee_754()) {
        if (finite_values_only) r += "f";
        if (nan_repr != 0 && nan_repr != 1) r += "n";
      }
      return r;
    }
    std::string r = (is_signed_ ? "int" : "uint") + std::to_string(size_bits());
    if (has_bias()) r += "b" + std::to_string(bias);
    return r;
  }
  bool equals(const ScalarTypeTorch& o) const {
    return exponent == o.exponent && mantissa == o.mantissa && bias == o.bias && is_signed_ == o.is_signed_ &&
           finite_values_only == o.finite_values_only && nan_repr == o.nan_repr;
  }
  // packed id: [0,8) exponent, [8,16) mantissa, [16] signed, [17,49) bias, [49] finite-only, [50,58) nan repr
  int64_t id() const {
    return exponent | (mantissa << 8) | ((int64_t)is_signed_ << 16) | ((bias & 0xffffffffll) << 17) |
           ((int64_t)finite_values_only << 49) | (nan_repr << 50);
  }
  static c10::intrusive_ptr<ScalarTypeTorch> from_id(int64_t v) {
    return c10::make_intrusive<ScalarTypeTorch>(v & 0xff, (v >> 8) & 0xff, (int64_t)(int32_t)((v >> 17) & 0xffffffffll),
                                                (bool)((v >> 16) & 1), (bool)((v >> 49) & 1), (v >> 50) & 0xff);
  }
};

using Ptr = c10::intrusive_ptr<ScalarTypeTorch>;

}  // namespace b200

TORCH_LIBRARY(_core_C, lib) {
  using namespace b200;
  lib.class_<ScalarTypeTorch>("ScalarType")
      .def(torch::init([](int64_t e, int64_t m, int64_t b, bool s) { return c10::make_intrusive<ScalarTypeTorch>(e, m, b, s); }))
      .def_property("mantissa", [](const Ptr& s) { return s->mantissa; })
      .def_property("exponent", [](const Ptr& s) { return s->exponent; })
      .def_property("bias", [](const Ptr& s) { return s->bias; })
      .def_property("signed", [](const Ptr& s) { return s->is_signed(); })
      .def_property("size_bits", [](const Ptr& s) { return s->size_bits(); })
      .def("is_signed", [](const Ptr& s) { return s->is_signed(); })
      .def("is_integer", [](const Ptr& s) { return s->is_integer(); })
      .def("is_floating_point", [](const Ptr& s) { return s->is_floating_point(); })
      .def("is_ieee_754", [](const Ptr& s) { return s->is_ieee_754(); })
      .def("has_nans", [](const Ptr& s) { return s->has_nans(); })
      .def("has_infs", [](const Ptr& s) { return s->has_infs(); })
      .def("has_bias", [](const Ptr& s) { return s->has_bias(); })
      .def("max", [](const Ptr& s) { return s->max(); })
      .def("min", [](const Ptr& s) { return s->min(); })
      .def("__len__", [](const Ptr&) -> int64_t { throw c10::TypeError({__func__, __FILE__, static_cast<uint32_t>(__LINE__)}, "__len__ not implemented"); })
      .def("__str__", [](const Ptr& s) { return s->str(); })
      .def("__repr__", [](const Ptr& s) { return "ScalarType." + s->str(); })
      .def("__eq__", [](const Ptr& a, const Ptr& b) { return a->equals(*b); })
      .def("__obj_flatten__", [](const Ptr& s) { return std::tuple<std::tuple<std::string, int64_t>>{{"ScalarType", s->id()}}; })
      .def_static("__obj_unflatten__", [](std::tuple<std::tuple<std::string, int64_t>> const& t) {
        return ScalarTypeTorch::from_id(std::get<1>(std::get<0>(t)));
      })
      .def_static("int_", [](int64_t size_bits, c10::optional<int64_t> bias) {
        return c10::make_intrusive<ScalarTypeTorch>(0, size_bits - 1, bias.value_or(0), true);
      })
      .def_static("uint", [](int64_t size_bits, c10::optional<int64_t> bias) {
        return c10::make_intrusive<ScalarTypeTorch>(0, size_bits, bias.value_or(0), false);
      })
      .def_static("float_IEEE754", [](int64_t e, int64_t m) { return c10::make_intrusive<ScalarTypeTorch>(e, m, 0, true); })
      .def_static("float_", [](int64_t e, int64_t m, bool fvo, int64_t nan_repr) {
        return c10::make_intrusive<ScalarTypeTorch>(e, m, 0, true, fvo, nan_repr);
      });
}

PyMODINIT_FUNC PyInit__core_C() {
  static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_core_C", nullptr, 0, nullptr};
  return PyModule_Create(&module);
}

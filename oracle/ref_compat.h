// Force-included (-include) when compiling the UNMODIFIED reference sources
// /root/reference/lib/model/csrc/{vision.cpp,cpu/*.cpp} against torch 2.x.
// The reference passes `tensor.type()` (DeprecatedTypeProperties) to AT_DISPATCH_FLOATING_TYPES
// (cpu/nms_cpu.cpp:71, cpu/ROIAlign_cpu.cpp:242); current ATen wants a ScalarType.  This shim
// re-defines the macro so that both spellings work, without touching a byte of the reference.
#pragma once
#include <torch/extension.h>
#include <ATen/Dispatch.h>

namespace dana_ref_compat {
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace dana_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(dana_ref_compat::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

// C++ host runtime of tnn_b200 (archiver / wire format, allocators, data loaders,
// tokenizer, profiler).  Bound into the same extension module as the kernels.
#pragma once
#include <pybind11/pybind11.h>
namespace tnn {
void runtime_bind(pybind11::module_& m);
}

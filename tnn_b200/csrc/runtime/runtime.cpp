#include "runtime.hpp"
namespace tnn {
void runtime_bind(pybind11::module_& m) { (void)m; }
}  // namespace tnn

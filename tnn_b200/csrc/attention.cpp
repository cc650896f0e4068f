#include <pybind11/pybind11.h>
namespace tnn {
void attention_bind(pybind11::module_& m) { (void)m; }
}  // namespace tnn

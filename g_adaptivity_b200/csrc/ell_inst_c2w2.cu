// Instantiation unit of the mesh-resident ELL kernels for CE = 2 live channels and 2 neighbour
// slots per row (one translation unit per combination so that nvcc compiles them in parallel).
#include "ell_kernels.cuh"

namespace gad {
namespace ell {
GAD_ELL_INSTANTIATE(2, 2)
}  // namespace ell
}  // namespace gad

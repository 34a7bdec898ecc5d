// Specialised 3D smoothness kernels (placeholder until the tiled kernel lands): returns false so the caller
// falls back to the generic kernel.
#include "internal.hpp"

namespace fi {

template <typename T>
bool stencil_fast_3d(const Geom&, const StencilTables&, const T*, T*, double*, double*, unsigned*, const int*, cudaStream_t)
{
	return false;
}

template bool stencil_fast_3d<float>(const Geom&, const StencilTables&, const float*, float*, double*, double*, unsigned*, const int*, cudaStream_t);
template bool stencil_fast_3d<double>(const Geom&, const StencilTables&, const double*, double*, double*, double*, unsigned*, const int*, cudaStream_t);

}  // namespace fi

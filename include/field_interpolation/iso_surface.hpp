// The two helpers the reference's 2D SDF demo keeps next to its solver calls (src/sdf_field.cpp:555-614, namespace
// sdf_field there) — they sit between the solve and marching squares, so they are part of what a caller of the
// library moves to the GPU together with it.  Same signatures and results (bit-identical fp32 arithmetic), computed
// by libfi_b200.so (fi_bicubic_upsample / fi_marching_squares, include/fi_b200.h).
#pragma once

#include <vector>

namespace field_interpolation {

// src/sdf_field.cpp:555-603.  Catmull-Rom upsampling; *io_width / *io_height are replaced by the upsampled sizes
// (upsample * size - upsample + 1).  upsample must be > 1 (the reference CHECKs); returns {} otherwise or on failure,
// leaving the sizes untouched.
std::vector<float> bicubic_upsample(int* io_width, int* io_height, const float* values, int upsample);

// src/sdf_field.cpp:605-614: emilib::marching_squares of (values - iso).
std::vector<float> iso_surface(int width, int height, const float* values, float iso);

}  // namespace field_interpolation

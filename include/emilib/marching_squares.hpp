// Drop-in for the reference's third_party/emilib/emilib/marching_squares.hpp:17-20 — the two functions its demo
// runs on every solved 2D field (src/sdf_field.cpp:193,613,669-670) — computed on the GPU by libfi_b200.so
// (fi_marching_squares / fi_calc_area, include/fi_b200.h; kernels in csrc/isosurface.cu).  Same namespace, names,
// argument meaning and results: segments bit-identical and in the same order.  A CUDA failure yields an empty
// vector / 0 and leaves the text in field_interpolation::b200::last_error().
#pragma once

#include <cstddef>
#include <vector>

namespace emilib {

// Zero contour of `iso` (row-major, width * height, positive = outside) as directed segments x0, y0, x1, y1.
std::vector<float> marching_squares(std::size_t width, std::size_t height, const float* iso);

// Signed area enclosed by `num_line_segments` segments laid out as marching_squares returns them.
float calc_area(std::size_t num_line_segments, const float* xy);

}  // namespace emilib

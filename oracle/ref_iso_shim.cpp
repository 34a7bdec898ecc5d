// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-linkage wrapper around the reference's iso-surface helpers, built by oracle/Makefile into
// oracle/_ref/libfi_ref.so next to the assembly TU:
//   * emilib::marching_squares / emilib::calc_area — /root/reference/third_party/emilib/emilib/marching_squares.cpp,
//     compiled UNMODIFIED from where it lies (its header uses an unqualified size_t that current libstdc++ no longer
//     leaks from <vector>; the Makefile passes `-include stddef.h`, a compiler flag, not a source change);
//   * bicubic_upsample / iso_surface — these live in the demo's GUI translation unit (src/sdf_field.cpp:555-614), which
//     needs SDL/ImGui and cannot be compiled here.  Their loops are restated below around the reference's OWN
//     emath::catmull_rom and emath::clamp (third_party/emath/emath/math.hpp:69,308-316, header-only, included as is),
//     so the floating-point kernel is the reference's.
#include <cstdint>
#include <cstring>
#include <vector>

#include <emath/math.hpp>               // resolved with -I/root/reference/third_party/emath
#include <emilib/marching_squares.hpp>  // resolved with -I/root/reference/third_party/emilib

extern "C" {

// Returns the number of floats (4 per segment); writes them when out != nullptr.
int64_t ref_marching_squares(int64_t width, int64_t height, const float* iso, float* out)
{
	const std::vector<float> lines = emilib::marching_squares(static_cast<size_t>(width), static_cast<size_t>(height), iso);
	if (out && !lines.empty()) { std::memcpy(out, lines.data(), lines.size() * sizeof(float)); }
	return static_cast<int64_t>(lines.size());
}

float ref_calc_area(int64_t num_segments, const float* xy) { return emilib::calc_area(static_cast<size_t>(num_segments), xy); }

// src/sdf_field.cpp:555-603 (loop structure restated; arithmetic by the reference's emath).
void ref_bicubic_upsample(int width, int height, const float* values, int upsample, float* out)
{
	const int  large_width = upsample * width - upsample + 1, large_height = upsample * height - upsample + 1;
	const auto value_at    = [&](int x, int y) {
        x = emath::clamp(x, 0, width - 1);
        y = emath::clamp(y, 0, height - 1);
        return values[y * width + x];
	};
	size_t k = 0;
	for (int ly = 0; ly < large_height; ++ly) {
		for (int lx = 0; lx < large_width; ++lx) {
			const float tx = static_cast<float>(lx % upsample) / static_cast<float>(upsample);
			const float ty = static_cast<float>(ly % upsample) / static_cast<float>(upsample);
			const int   sx = lx / upsample, sy = ly / upsample;
			float       rows[4];
			for (int j = 0; j < 4; ++j) {
				rows[j] = emath::catmull_rom(tx, value_at(sx - 1, sy - 1 + j), value_at(sx + 0, sy - 1 + j), value_at(sx + 1, sy - 1 + j),
				                             value_at(sx + 2, sy - 1 + j));
			}
			out[k++] = emath::catmull_rom(ty, rows[0], rows[1], rows[2], rows[3]);
		}
	}
}

// src/sdf_field.cpp:605-614.
int64_t ref_iso_surface(int width, int height, const float* values, float iso, float* out)
{
	std::vector<float> iso_at_zero(static_cast<size_t>(width) * height);
	for (size_t i = 0; i < iso_at_zero.size(); ++i) { iso_at_zero[i] = values[i] - iso; }
	return ref_marching_squares(width, height, iso_at_zero.data(), out);
}

}  // extern "C"

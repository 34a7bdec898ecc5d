// Test driver for the C++ drop-in API (include/field_interpolation/*.hpp).  Each scenario is written the way
// the reference's own callers use the library (cited per scenario) and dumps the resulting system and
// solution to a binary file that tests/test_cpp_api.py compares with the CPU oracle.
//
//   api_driver <scenario> <out.bin> [in.bin]
//
// File format (little endian): repeated records { char name[16]; int64 count; int32 dtype (0 f32, 1 i32);
// payload }.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include <emilib/marching_squares.hpp>
#include <field_interpolation/field_interpolation.hpp>
#include <field_interpolation/iso_surface.hpp>

namespace fi = field_interpolation;

static FILE* g_out = nullptr;

static void put(const char* name, const void* data, int64_t count, int32_t dtype)
{
	char n[16] = {0};
	std::strncpy(n, name, 15);
	std::fwrite(n, 1, 16, g_out);
	std::fwrite(&count, sizeof(count), 1, g_out);
	std::fwrite(&dtype, sizeof(dtype), 1, g_out);
	std::fwrite(data, 4, static_cast<size_t>(count), g_out);
}
static void put_f(const char* name, const std::vector<float>& v) { put(name, v.data(), static_cast<int64_t>(v.size()), 0); }
static void put_i(const char* name, const std::vector<int>& v) { put(name, v.data(), static_cast<int64_t>(v.size()), 1); }

static void put_system(const fi::LinearEquation& eq)
{
	std::vector<int>   rows, cols;
	std::vector<float> vals;
	for (const fi::Triplet& t : eq.triplets) {
		rows.push_back(t.row);
		cols.push_back(t.col);
		vals.push_back(t.value);
	}
	put_i("rows", rows);
	put_i("cols", cols);
	put_f("vals", vals);
	put_f("rhs", eq.rhs);
}

static std::vector<float> read_floats(FILE* f)
{
	int64_t n = 0;
	if (std::fread(&n, sizeof(n), 1, f) != 1) { return {}; }
	std::vector<float> v(static_cast<size_t>(n));
	if (n && std::fread(v.data(), 4, v.size(), f) != v.size()) { return {}; }
	return v;
}

// reference src/field_1d.cpp:98-110 — data rows first, then the model, exact solve
static int scenario_field_1d(int resolution)
{
	struct Point { float pos, value, gradient; };
	const Point points[] = {{0.2f, 0.0f, +1.0f}, {0.8f, 0.0f, -1.0f}};  // src/field_1d.cpp:20-29
	fi::Weights weights;
	fi::LatticeField field{{resolution}};
	std::vector<int> accepted;
	for (const Point& point : points) {
		float pos_lattice      = point.pos * (resolution - 1);
		float gradient_lattice = point.gradient / (resolution - 1);
		accepted.push_back(add_value_constraint(&field, &pos_lattice, point.value, weights.data_pos));
		accepted.push_back(add_gradient_constraint(&field, &pos_lattice, &gradient_lattice, weights.data_gradient, weights.gradient_kernel));
	}
	add_field_constraints(&field, weights);
	const size_t num_unknowns = resolution;
	auto interpolated = solve_sparse_linear_exact(field.eq, num_unknowns);
	put_system(field.eq);
	put_i("accepted", accepted);
	put_f("solution", interpolated);
	std::ostringstream os;
	os << field.eq;  // src/field_1d.cpp:83-90 prints the system
	std::vector<int> printed{static_cast<int>(os.str().size())};
	put_i("printed", printed);
	return interpolated.size() == num_unknowns ? 0 : 2;
}

// README.md:11-40 on the real library
static int scenario_readme()
{
	fi::LatticeField field{{6}};
	std::vector<int> accepted;
	const float p0 = 0.0f, p5 = 5.0f, gp = 1.0f, gm = -1.0f;
	accepted.push_back(add_value_constraint(&field, &p0, 4.0f, 1.0f));
	accepted.push_back(add_value_constraint(&field, &p5, 2.0f, 1.0f));
	accepted.push_back(add_gradient_constraint(&field, &p0, &gp, 1.0f, fi::GradientKernel::kNearestNeighbor));
	accepted.push_back(add_gradient_constraint(&field, &p5, &gm, 1.0f, fi::GradientKernel::kNearestNeighbor));
	fi::Weights w;
	w.model_2 = 1.0f;
	add_field_constraints(&field, w);
	put_system(field.eq);
	put_i("accepted", accepted);
	put_f("solution", solve_sparse_linear_exact(field.eq, 6));
	return 0;
}

// reference src/interpolate_2d.cpp:31-47 — model rows first, 4x4 value grid with zero-gradient constraints
static int scenario_interpolate_2d(int resolution)
{
	const float values[16] = {5, 4, 2, 3, 4, 2, 1, 5, 6, 3, 5, 2, 1, 2, 4, 1};
	fi::Weights weights;
	fi::LatticeField field{{resolution, resolution}};
	add_field_constraints(&field, weights);
	for (int y = 0; y < 4; ++y) {
		for (int x = 0; x < 4; ++x) {
			const float pos[2]  = {x / 3.0f * (resolution - 1.0f), y / 3.0f * (resolution - 1.0f)};
			add_value_constraint(&field, pos, values[y * 4 + x], weights.data_pos);
			const float zero[2] = {0, 0};
			add_gradient_constraint(&field, pos, zero, weights.data_gradient, weights.gradient_kernel);
		}
	}
	put_system(field.eq);
	put_f("solution", solve_sparse_linear_exact(field.eq, resolution * resolution));
	return 0;
}

// reference src/sdf_field.cpp:212-249 + :251-304 — sdf_from_points, border-distance rows appended by hand with
// add_equation, exact solve; then the approximate path: coarse solve -> upscale_field -> * factor ->
// solve_tiled_with_guess; and the error heat map (:350).
static int scenario_sdf_2d(const char* in_path)
{
	FILE* in = std::fopen(in_path, "rb");
	if (!in) { return 3; }
	const std::vector<float> dims = read_floats(in);  // width, height, boundary_weight, downscale_factor
	const std::vector<float> unit = read_floats(in);  // unit positions xyxy..
	const std::vector<float> nrm  = read_floats(in);
	std::fclose(in);
	const int   width = static_cast<int>(dims[0]), height = static_cast<int>(dims[1]);
	const float boundary_weight = dims[2];
	const int   factor = static_cast<int>(dims[3]);
	const int   npts = static_cast<int>(unit.size() / 2);
	fi::Weights weights;

	auto generate_sdf_field = [&](int w, int h, const std::vector<float>& pos) {
		auto field = fi::sdf_from_points({w, h}, weights, npts, pos.data(), nrm.data(), nullptr);
		if (boundary_weight > 0) {
			for (int y = 0; y < h; ++y) {
				for (int x = 0; x < w; ++x) {
					if (!(x == 0 || x == w - 1 || y == 0 || y == h - 1)) { continue; }
					float closest = std::numeric_limits<float>::infinity();
					for (int i = 0; i < npts; ++i) {
						const float dx = pos[2 * i] - x, dy = pos[2 * i + 1] - y;
						closest = std::min(closest, dx * dx + dy * dy);
					}
					add_equation(&field.eq, fi::Weight{boundary_weight}, {std::sqrt(closest)}, {{y * w + x, 1.0f}});
				}
			}
		}
		return field;
	};
	auto on_lattice = [&](int w, int h) {
		std::vector<float> p(unit.size());
		for (int i = 0; i < npts; ++i) {
			p[2 * i]     = unit[2 * i] * (w - 1.0f);
			p[2 * i + 1] = unit[2 * i + 1] * (h - 1.0f);
		}
		return p;
	};

	const auto pos   = on_lattice(width, height);
	auto       field = generate_sdf_field(width, height, pos);
	put_system(field.eq);
	const auto exact = solve_sparse_linear_exact(field.eq, width * height);
	put_f("exact", exact);

	const int  ws = (width + factor - 1) / factor, hs = (height + factor - 1) / factor;
	auto       field_small = generate_sdf_field(ws, hs, on_lattice(ws, hs));
	const auto small       = solve_sparse_linear_exact(field_small.eq, ws * hs);
	put_f("small", small);
	auto sdf = fi::upscale_field(small.data(), {ws, hs}, {width, height});
	put_f("upscaled", sdf);
	for (float& v : sdf) { v *= factor; }
	fi::SolveOptions so;
	so.error_tolerance = 1e-4f;
	const auto approx = solve_tiled_with_guess(field.eq, sdf, {width, height}, so);
	put_f("approx", approx);
	put_f("bad_guess", solve_tiled_with_guess(field.eq, small, {width, height}, so));  // wrong size -> {}
	fi::SolveOptions tiles_only;  // the demo's "tile" checkbox with CG off: tile_solver_square alone
	tiles_only.tile = true;
	tiles_only.tile_size = 16;
	tiles_only.cg = false;
	put_f("tiled", solve_tiled_with_guess(field.eq, sdf, {width, height}, tiles_only));
	fi::SolveOptions tiles_cg = so;
	tiles_cg.tile = true;
	put_f("tiled_cg", solve_tiled_with_guess(field.eq, sdf, {width, height}, tiles_cg));
	put_f("heatmap", generate_error_map(field.eq.triplets, exact, field.eq.rhs));
	put_f("jacobi", jacobi_iterations(field.eq, sdf, 5, 0.5f));
	return 0;
}

// a system written entirely by hand (reference src/line_2d.cpp / src/bipolar_2d.cpp style): no lattice at all
static int scenario_hand_rows()
{
	fi::LinearEquation eq;
	const int n = 9;
	for (int i = 0; i + 1 < n; ++i) { add_equation(&eq, fi::Weight{0.5f}, fi::Rhs{1.0f}, {{i, -1.0f}, {i + 1, 1.0f}}); }
	add_equation(&eq, fi::Weight{2.0f}, fi::Rhs{3.0f}, {{0, 1.0f}});
	add_equation(&eq, fi::Weight{0.0f}, fi::Rhs{3.0f}, {{0, 1.0f}});           // dropped: zero weight
	add_equation(&eq, fi::Weight{1.0f}, fi::Rhs{3.0f}, {{4, 0.0f}});           // dropped: all-zero row
	add_equation(&eq, fi::Weight{1.0f}, fi::Rhs{1.0f}, {{3, 1.0f}, {3, 0.5f}, {5, 0.0f}});  // duplicate column
	put_system(eq);
	put_f("exact", solve_sparse_linear_exact(eq, n));
	put_f("fast", solve_sparse_linear_fast(eq, n));
	put_f("guess", solve_sparse_linear_with_guess(eq, std::vector<float>(n, 0.0f), 0, 1e-6f));
	return 0;
}

// deferred mode: nothing of size O(rows) on the host; counts still observable; materialize on demand
static int scenario_deferred_3d()
{
	const int n = 20;
	fi::LatticeField field{{n, n, n}};
	fi::b200::defer_triplets(&field, true);
	std::vector<float> pos, nrm;
	for (int i = 0; i < 500; ++i) {
		const float a = 0.37f * i, b = 0.11f * i;
		const float d[3] = {std::cos(a) * std::sin(b), std::sin(a) * std::sin(b), std::cos(b)};
		for (int k = 0; k < 3; ++k) {
			pos.push_back((0.5f + 0.3f * d[k]) * (n - 1.0f));
			nrm.push_back(d[k]);
		}
	}
	fi::Weights w;
	add_field_constraints(&field, w);
	add_points(&field, w.data_pos, w.value_kernel, w.data_gradient, w.gradient_kernel, 500, pos.data(), nrm.data(), nullptr);
	long long rows = 0, trips = 0;
	fi::b200::counts(field, &rows, &trips);
	std::vector<int> c{static_cast<int>(rows), static_cast<int>(trips), static_cast<int>(field.eq.rhs.size())};
	put_i("counts", c);
	fi::b200::SolveStats st;
	put_f("solution", fi::b200::solve(field.eq, n * n * n, fi::b200::Precision::kDouble, nullptr, 0, 1e-11, &st));
	fi::b200::materialize(&field);
	put_system(field.eq);
	put_f("points", pos);
	put_f("normals", nrm);
	std::vector<int> s{static_cast<int>(st.iterations), st.converged ? 1 : 0};
	put_i("stats", s);
	return 0;
}

// LatticeField / LinearEquation are plain value types in the reference (field_interpolation.hpp:97-114,
// sparse_linear.hpp:18-22): a copy is an independent system, and callers may clear or rewrite eq between calls.
static int scenario_value_semantics()
{
	const int n = 14;
	fi::Weights w;
	fi::LatticeField a{{n, n}};
	add_field_constraints(&a, w);
	// (the second-difference rows leave 1, x, y, xy undetermined: five value rows pin the field)
	const float p0[2] = {3.25f, 4.5f}, p1[2] = {9.75f, 8.125f}, p2[2] = {6.5f, 2.25f};
	const float q0[2] = {1.5f, 11.25f}, q1[2] = {11.0f, 1.75f}, q2[2] = {6.0f, 7.0f};
	add_value_constraint(&a, p0, 1.0f, 1.0f);
	add_value_constraint(&a, p1, -2.0f, 1.0f);
	add_value_constraint(&a, q0, 0.5f, 1.0f);
	add_value_constraint(&a, q1, 3.0f, 1.0f);
	add_value_constraint(&a, q2, -1.0f, 1.0f);
	fi::LatticeField b = a;  // shares nothing observable with a from here on
	add_value_constraint(&b, p2, 5.0f, 2.0f);
	add_equation(&b.eq, fi::Weight{1.5f}, fi::Rhs{0.25f}, {{0, 1.0f}, {n * n - 1, -1.0f}});
	put_i("a_counts", {static_cast<int>(a.eq.rhs.size()), static_cast<int>(a.eq.triplets.size())});
	put_f("a_exact", solve_sparse_linear_exact(a.eq, n * n));
	put_f("b_exact", solve_sparse_linear_exact(b.eq, n * n));
	put_f("a_again", solve_sparse_linear_exact(a.eq, n * n));  // b's solve (which forwarded b's hand-written row) left a alone
	// a copy with pending hand-written rows on the ORIGINAL: the copy must not see them
	fi::LatticeField c = a;
	add_equation(&a.eq, fi::Weight{3.0f}, fi::Rhs{1.0f}, {{5, 1.0f}});
	put_f("a_plus_row", solve_sparse_linear_exact(a.eq, n * n));
	put_f("c_exact", solve_sparse_linear_exact(c.eq, n * n));
	// the caller empties the equation and builds another system in the same field object
	fi::LatticeField d = b;
	d.eq.triplets.clear();
	d.eq.rhs.clear();
	add_field_constraints(&d, w);
	add_value_constraint(&d, p2, 7.0f, 1.0f);
	add_value_constraint(&d, p0, 2.0f, 1.0f);
	add_value_constraint(&d, q0, -3.0f, 1.0f);
	add_value_constraint(&d, q1, 0.25f, 1.0f);
	add_value_constraint(&d, q2, 1.0f, 1.0f);
	put_i("d_counts", {static_cast<int>(d.eq.rhs.size()), static_cast<int>(d.eq.triplets.size())});
	put_f("d_exact", solve_sparse_linear_exact(d.eq, n * n));
	// ... or truncates it (drops the last constraint) and solves without another builder call
	fi::LatticeField e = b;  // b's description already holds its hand-written row (2 triplets): e.eq is no extension of it any more
	e.eq.rhs.pop_back();
	e.eq.triplets.pop_back();
	e.eq.triplets.pop_back();
	put_f("e_exact", solve_sparse_linear_exact(e.eq, n * n));
	return 0;
}

// reference src/sdf_field.cpp:660-670 — what the demo does with a solved 2D field: optional bicubic upsampling, the zero
// contour by marching squares, its area relative to the lattice; plus one off-zero iso line (:701).
static int scenario_iso_2d(const char* in_path)
{
	FILE* in = std::fopen(in_path, "rb");
	if (!in) { return 3; }
	const std::vector<float> dims = read_floats(in);  // width, height, upsampling, iso
	const std::vector<float> sdf  = read_floats(in);
	std::fclose(in);
	int       iso_width = static_cast<int>(dims[0]), iso_height = static_cast<int>(dims[1]);
	const int upsampling = static_cast<int>(dims[2]);
	if (sdf.size() != static_cast<size_t>(iso_width) * iso_height) { return 4; }

	const std::vector<float> plain = emilib::marching_squares(iso_width, iso_height, sdf.data());
	put_f("plain_lines", plain);
	put_f("plain_area", {emilib::calc_area(plain.size() / 4, plain.data())});

	std::vector<float> iso_source = sdf;
	if (upsampling > 1) { iso_source = fi::bicubic_upsample(&iso_width, &iso_height, iso_source.data(), upsampling); }
	put_i("up_size", {iso_width, iso_height});
	put_f("upsampled", iso_source);
	const std::vector<float> zero_lines = fi::iso_surface(iso_width, iso_height, iso_source.data(), 0.0f);
	put_f("zero_lines", zero_lines);
	const float side = static_cast<float>(iso_width - 1);
	put_f("lines_area", {emilib::calc_area(zero_lines.size() / 4, zero_lines.data()) / (side * side)});
	put_f("iso_lines", fi::iso_surface(iso_width, iso_height, iso_source.data(), dims[3]));
	int w1 = 3, h1 = 3;
	put_f("bad_upsample", fi::bicubic_upsample(&w1, &h1, sdf.data(), 1));  // the reference CHECKs upsample > 1: {} here
	return 0;
}

int main(int argc, char** argv)
{
	if (argc < 3) {
		std::fprintf(stderr, "usage: api_driver <scenario> <out.bin> [in.bin]\n");
		return 64;
	}
	const std::string sc = argv[1];
	g_out = std::fopen(argv[2], "wb");
	if (!g_out) { return 65; }
	int rc = 66;
	if (sc == "field_1d_100") { rc = scenario_field_1d(100); }
	else if (sc == "field_1d_12") { rc = scenario_field_1d(12); }
	else if (sc == "readme") { rc = scenario_readme(); }
	else if (sc == "interpolate_2d") { rc = scenario_interpolate_2d(24); }
	else if (sc == "sdf_2d" && argc > 3) { rc = scenario_sdf_2d(argv[3]); }
	else if (sc == "iso_2d" && argc > 3) { rc = scenario_iso_2d(argv[3]); }
	else if (sc == "hand_rows") { rc = scenario_hand_rows(); }
	else if (sc == "deferred_3d") { rc = scenario_deferred_3d(); }
	else if (sc == "value_semantics") { rc = scenario_value_semantics(); }
	std::fclose(g_out);
	if (rc != 0) { std::fprintf(stderr, "scenario %s failed (%d): %s\n", sc.c_str(), rc, fi::b200::last_error()); }
	return rc;
}

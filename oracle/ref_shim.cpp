// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C-linkage wrapper around the *unmodified* reference assembly translation unit
// (/root/reference/field_interpolation/field_interpolation.cpp), compiled where it lies by
// oracle/Makefile into oracle/_ref/libfi_ref.so.  Nothing here is shipped; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load the result.
//
// The reference's solve half (sparse_linear.cpp) needs Eigen, which is neither vendored in the
// reference nor installed in this image, so that TU cannot be built.  The assembly TU needs exactly
// one symbol from it, `add_equation` (sparse_linear.cpp:34-50); it is restated below so the
// reference's own constraint builders link.  Everything else in this file only forwards.
#include <cstdint>
#include <cstring>
#include <vector>

#include <field_interpolation/field_interpolation.hpp>  // resolved with -I/root/reference

namespace field_interpolation {

// Restatement of sparse_linear.cpp:34-50: a weighted row is dropped when the weight is zero, pairs
// with a zero coefficient are dropped, and the rhs entry is pushed only if a pair survived.
void add_equation(LinearEquation* eq, Weight weight, Rhs rhs, std::initializer_list<LinearEquationPair> pairs)
{
	if (weight.value == 0) { return; }
	const int row = static_cast<int>(eq->rhs.size());
	bool kept_any = false;
	for (const LinearEquationPair& p : pairs) {
		if (p.value == 0) { continue; }
		eq->triplets.emplace_back(row, p.column, p.value * weight.value);
		kept_any = true;
	}
	if (kept_any) { eq->rhs.emplace_back(rhs.value * weight.value); }
}

} // namespace field_interpolation

namespace fi = field_interpolation;

namespace {

struct RefWeights  // mirrors the field order of fi::Weights (field_interpolation.hpp:75-95)
{
	float data_pos, data_gradient, model_0, model_1, model_2, model_3, model_4, gradient_smoothness;
	int   value_kernel, gradient_kernel;
};

fi::Weights to_weights(const RefWeights* w)
{
	fi::Weights out;
	out.data_pos            = w->data_pos;
	out.data_gradient       = w->data_gradient;
	out.model_0             = w->model_0;
	out.model_1             = w->model_1;
	out.model_2             = w->model_2;
	out.model_3             = w->model_3;
	out.model_4             = w->model_4;
	out.gradient_smoothness = w->gradient_smoothness;
	out.value_kernel        = static_cast<fi::ValueKernel>(w->value_kernel);
	out.gradient_kernel     = static_cast<fi::GradientKernel>(w->gradient_kernel);
	return out;
}

} // namespace

extern "C" {

void* ref_field_create(int ndim, const int* sizes)
{
	return new fi::LatticeField(std::vector<int>(sizes, sizes + ndim));
}

void ref_field_destroy(void* f) { delete static_cast<fi::LatticeField*>(f); }

void ref_add_field_constraints(void* f, const RefWeights* w)
{
	fi::add_field_constraints(static_cast<fi::LatticeField*>(f), to_weights(w));
}

int ref_add_value_constraint(void* f, const float* pos, float value, float weight)
{
	return fi::add_value_constraint(static_cast<fi::LatticeField*>(f), pos, value, weight) ? 1 : 0;
}

int ref_add_value_constraint_nearest_neighbor(void* f, const float* pos, const float* gradient, float value, float weight)
{
	return fi::add_value_constraint_nearest_neighbor(static_cast<fi::LatticeField*>(f), pos, gradient, value, weight) ? 1 : 0;
}

int ref_add_gradient_constraint(void* f, const float* pos, const float* gradient, float weight, int kernel)
{
	return fi::add_gradient_constraint(
		static_cast<fi::LatticeField*>(f), pos, gradient, weight, static_cast<fi::GradientKernel>(kernel)) ? 1 : 0;
}

void ref_add_points(void* f, float value_weight, int value_kernel, float gradient_weight, int gradient_kernel,
                    int num_points, const float* positions, const float* normals, const float* point_weights)
{
	fi::add_points(static_cast<fi::LatticeField*>(f), value_weight, static_cast<fi::ValueKernel>(value_kernel),
	               gradient_weight, static_cast<fi::GradientKernel>(gradient_kernel), num_points, positions, normals,
	               point_weights);
}

void ref_add_equation(void* f, float weight, float rhs, int num_pairs, const int* columns, const float* values)
{
	// add_equation takes an initializer_list; the demo callers use 1..5 pairs (sdf_field.cpp:239-241).
	fi::LinearEquation* eq = &static_cast<fi::LatticeField*>(f)->eq;
	auto P = [&](int i) { return fi::LinearEquationPair{columns[i], values[i]}; };
	switch (num_pairs) {
		case 1: fi::add_equation(eq, fi::Weight{weight}, fi::Rhs{rhs}, {P(0)}); break;
		case 2: fi::add_equation(eq, fi::Weight{weight}, fi::Rhs{rhs}, {P(0), P(1)}); break;
		case 3: fi::add_equation(eq, fi::Weight{weight}, fi::Rhs{rhs}, {P(0), P(1), P(2)}); break;
		case 4: fi::add_equation(eq, fi::Weight{weight}, fi::Rhs{rhs}, {P(0), P(1), P(2), P(3)}); break;
		case 5: fi::add_equation(eq, fi::Weight{weight}, fi::Rhs{rhs}, {P(0), P(1), P(2), P(3), P(4)}); break;
		default: break;
	}
}

void* ref_sdf_from_points(int ndim, const int* sizes, const RefWeights* w, int num_points, const float* positions,
                          const float* normals, const float* point_weights)
{
	auto* out = new fi::LatticeField();
	*out = fi::sdf_from_points(std::vector<int>(sizes, sizes + ndim), to_weights(w), num_points, positions, normals,
	                           point_weights);
	return out;
}

int64_t ref_num_rows(void* f) { return static_cast<int64_t>(static_cast<fi::LatticeField*>(f)->eq.rhs.size()); }
int64_t ref_num_triplets(void* f) { return static_cast<int64_t>(static_cast<fi::LatticeField*>(f)->eq.triplets.size()); }

// Triplet is {int row, col; float value} = 12 bytes (sparse_linear.hpp:8-15); copied out as SoA.
void ref_copy_system(void* f, int* rows, int* cols, float* values, float* rhs)
{
	const fi::LinearEquation& eq = static_cast<fi::LatticeField*>(f)->eq;
	for (size_t i = 0; i < eq.triplets.size(); ++i) {
		rows[i]   = eq.triplets[i].row;
		cols[i]   = eq.triplets[i].col;
		values[i] = eq.triplets[i].value;
	}
	if (!eq.rhs.empty()) { std::memcpy(rhs, eq.rhs.data(), eq.rhs.size() * sizeof(float)); }
}

void ref_upscale_field(const float* small_field, int ndim, const int* small_sizes, const int* large_sizes, float* out)
{
	const std::vector<float> big = fi::upscale_field(
		small_field, std::vector<int>(small_sizes, small_sizes + ndim), std::vector<int>(large_sizes, large_sizes + ndim));
	std::memcpy(out, big.data(), big.size() * sizeof(float));
}

void ref_generate_error_map(int64_t num_triplets, const int* rows, const int* cols, const float* values,
                            int64_t num_unknowns, const float* solution, int64_t num_rows, const float* rhs, float* out)
{
	std::vector<fi::Triplet> trips;
	trips.reserve(static_cast<size_t>(num_triplets));
	for (int64_t i = 0; i < num_triplets; ++i) { trips.emplace_back(rows[i], cols[i], values[i]); }
	const std::vector<float> sol(solution, solution + num_unknowns);
	const std::vector<float> b(rhs, rhs + num_rows);
	const std::vector<float> heat = fi::generate_error_map(trips, sol, b);
	std::memcpy(out, heat.data(), heat.size() * sizeof(float));
}

} // extern "C"

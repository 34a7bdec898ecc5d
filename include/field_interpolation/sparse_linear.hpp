// field_interpolation/sparse_linear.hpp — B200 drop-in for the header of the same name in
// emilk/field_interpolation (reference field_interpolation/sparse_linear.hpp:8-80).  Same namespace, types,
// function names, argument meaning and failure behaviour (a failed solve returns an empty vector); the
// implementation (field_interpolation_b200/host/sparse_linear.cpp) is host C++ over the C ABI of
// libfi_b200.so (include/fi_b200.h), where the work runs as CUDA kernels for sm_100a.
//
// Differences a caller can observe, all additive:
//   * LinearEquation carries an opaque `structured` handle.  When the equation was built by the LatticeField
//     builders of field_interpolation.hpp it points at the device-resident structured description (model
//     weights, point arrays, extra rows) and the solvers below run matrix-free from it; rows the caller
//     appended by hand after that (add_equation, push_back) are picked up as generic rows.  An equation
//     without the handle (built purely by hand) is solved through the generic-rows path.
//   * The direct Cholesky solves (SimplicialLLT) are replaced by Jacobi-preconditioned conjugate gradients on
//     the same normal equations; see b200::set_exact_tolerance.
#pragma once

#include <initializer_list>
#include <iosfwd>
#include <memory>
#include <vector>

namespace field_interpolation {

namespace b200 {
struct Structured;  // device-side description of a LatticeField's equations (opaque; host/structured.hpp)
}

struct Triplet  // reference sparse_linear.hpp:8-15 — layout-identical to fi_triplet and Eigen::Triplet<float>
{
	int   row, col;
	float value;

	Triplet() {}
	Triplet(int row_, int col_, float value_) : row(row_), col(col_), value(value_) {}
};

/// Sparse Ax=b where A is described by `triplets` and `rhs` is b.  (reference :18-22)
struct LinearEquation
{
	std::vector<Triplet> triplets;
	std::vector<float>   rhs;

	std::shared_ptr<b200::Structured> structured;  ///< B200 extension, see the header comment.
};

std::ostream& operator<<(std::ostream& os, const LinearEquation& eq);  // reference sparse_linear.cpp:11-32

struct LinearEquationPair  // reference :26-30
{
	int   column;
	float value;
};

struct Weight { float value; };  // reference :32
struct Rhs    { float value; };  // reference :33

/// Helper to add a row to the linear equation (reference sparse_linear.cpp:34-50): the row is skipped when
/// weight == 0, pairs with value == 0 are skipped, the right-hand side is pushed only if a pair was kept.
void add_equation(LinearEquation* eq, Weight weight, Rhs rhs, std::initializer_list<LinearEquationPair> pairs);

/// Least-squares solve of A x = rhs through the normal equations.  `num_columns` = number of unknowns.
/// Duplicate elements in triplets are summed.  (reference sparse_linear.cpp:115-152 float Cholesky)
std::vector<float> solve_sparse_linear_fast(const LinearEquation& eq, int num_columns);

/// (reference sparse_linear.cpp:154-184 double Cholesky)  fp64 PCG to b200::exact_tolerance().
std::vector<float> solve_sparse_linear_exact(const LinearEquation& eq, int num_columns);

/// Iterative solve from `guess` (reference sparse_linear.cpp:186-212: float, diagonal preconditioner, stops at
/// |r| <= error_tolerance |A^T b|).  max_iterations 0 = default (2 x problem size), error_tolerance 0 = float epsilon.
std::vector<float> solve_sparse_linear_with_guess(const LinearEquation& eq, const std::vector<float>& guess, int max_iterations,
                                                  float error_tolerance);

/// Jacobi iterations (reference sparse_linear.cpp:214-241): x <- w (A^T b - R x) / D + (1 - w) x.
std::vector<float> jacobi_iterations(const LinearEquation& eq, const std::vector<float>& guess, const int num_iterations,
                                     const float weight);

struct SolveOptions  // reference :66-73, same defaults
{
	bool  tile            = false;
	int   tile_size       = 16;
	bool  cg              = true;
	int   max_iterations  = 0;
	float error_tolerance = 1e-3f;
};

/// Approximate solver: guess (+ tile phase) + conjugate gradients (reference sparse_linear.cpp:392-443).
/// A guess whose size is not the lattice size returns {} as the reference does.
std::vector<float> solve_tiled_with_guess(const LinearEquation& eq, const std::vector<float>& guess, const std::vector<int>& sizes,
                                          const SolveOptions& options);

// ---- B200 extensions ---------------------------------------------------------------------------------------
namespace b200 {

enum class Precision { kFloat = 0, kDouble = 1, kMixed = 2 };  // fi_precision

struct SolveStats  // fi_solve_stats
{
	long long iterations        = 0;
	double    relative_residual = 0, true_residual = 0, initial_residual = 0;
	double    setup_ms = 0, solve_ms = 0;
	bool      converged = false;  ///< by the residual recomputed from the solution, not merely by the recurrence
	long long occupied_cells = 0, generic_rows = 0;
	long long widened_after = -1;  ///< >= 0: a float multigrid solve continued with the fp64 outer CG after this many iterations
};

/// Relative-residual target of solve_sparse_linear_exact / _fast (defaults 1e-10 / 1e-9).
void   set_exact_tolerance(double exact, double fast);
double exact_tolerance();

/// The general entry point behind the reference-named solvers: PCG in the chosen arithmetic from `guess`
/// (nullptr: zeros).  Returns {} on failure; `stats` (nullable) receives the counters of the run.
std::vector<float> solve(const LinearEquation& eq, int num_columns, Precision precision, const std::vector<float>* guess,
                         int max_iterations, double tolerance, SolveStats* stats);

/// Text of the last failure inside libfi_b200 on this thread.
const char* last_error();

}  // namespace b200

}  // namespace field_interpolation

// Host C++ implementation of include/field_interpolation/sparse_linear.hpp over the C ABI of libfi_b200.so.
// Each function cites the reference function whose behaviour it reproduces (field_interpolation/
// sparse_linear.cpp of the reference tree).  The reference solves through Eigen (Cholesky / BiCGSTAB); here
// every solve is the Jacobi-preconditioned CG of csrc/solver.cu on the same normal equations with the same
// stopping rule, failure convention (empty vector) and defaults.
#include <cstring>
#include <ostream>

#include "structured.hpp"

namespace field_interpolation {

// reference sparse_linear.cpp:11-32 (debug print: one line per row, "rhs = v * xc  +  v * xc")
std::ostream& operator<<(std::ostream& os, const LinearEquation& eq)
{
	const size_t rows = eq.rhs.size();
	std::vector<size_t> start(rows + 1, 0);
	for (const Triplet& t : eq.triplets) {
		if (t.row >= 0 && static_cast<size_t>(t.row) < rows) { ++start[t.row + 1]; }
	}
	for (size_t r = 0; r < rows; ++r) { start[r + 1] += start[r]; }
	std::vector<const Triplet*> by_row(start[rows]);
	std::vector<size_t>         at(start.begin(), start.end() - 1);
	for (const Triplet& t : eq.triplets) {
		if (t.row >= 0 && static_cast<size_t>(t.row) < rows) { by_row[at[t.row]++] = &t; }
	}
	for (size_t r = 0; r < rows; ++r) {
		os << eq.rhs[r] << " = ";
		for (size_t k = start[r]; k < start[r + 1]; ++k) {
			os << by_row[k]->value << " * x" << by_row[k]->col;
			if (k + 1 < start[r + 1]) { os << "  +  "; }
		}
		os << "\n";
	}
	return os;
}

// reference sparse_linear.cpp:34-50.  Pure bookkeeping on the host-side triplet list; the row reaches the
// device as a generic row at the next builder or solve call (structured.hpp: forward_tail_rows).
void add_equation(LinearEquation* eq, Weight weight, Rhs rhs, std::initializer_list<LinearEquationPair> pairs)
{
	if (weight.value == 0) { return; }
	const int row  = static_cast<int>(eq->rhs.size());
	bool      kept = false;
	for (const LinearEquationPair& p : pairs) {
		if (p.value == 0) { continue; }
		eq->triplets.emplace_back(row, p.column, p.value * weight.value);
		kept = true;
	}
	if (kept) { eq->rhs.emplace_back(rhs.value * weight.value); }
}

namespace b200 {

static double g_exact_tol = 1e-10, g_fast_tol = 1e-9;

void set_exact_tolerance(double exact, double fast)
{
	g_exact_tol = exact;
	g_fast_tol  = fast;
}
double      exact_tolerance() { return g_exact_tol; }
const char* last_error() { return fi_last_error(); }

// The device description to solve `eq` from: its own structured handle when it has one that is still
// consistent with the triplet list, else a fresh generic-rows description of the whole list.
// `lattice` (nullable): the lattice shape the caller states (solve_tiled_with_guess); a description of another
// shape is not reused, and a fresh one is created with that shape so that tiles mean what the caller means.
static Structured* description_of(const LinearEquation& eq, long long num_columns, std::shared_ptr<Structured>* temp,
                                  const std::vector<int>* lattice = nullptr)
{
	Structured* st = eq.structured.get();
	if (st) {
		long long n = 1;
		for (int s : st->sizes) { n *= s; }
		const bool shape_ok = !lattice || st->sizes == *lattice;
		if (shape_ok && (num_columns <= 0 || n == num_columns)) {
			const bool pending = eq.rhs.size() != st->eq_rows || eq.triplets.size() != st->eq_triplets;
			if (pending && eq.structured.use_count() > 1) {
				// hand-written rows wait to be forwarded and the description is shared with a copy of this equation:
				// forward them into a private clone, the other owner must not see them
				*temp = clone_description(*st);
				st    = temp->get();
				if (!st) { return nullptr; }
			}
			const Forwarded fwd = forward_tail_rows(eq, st);
			if (fwd == Forwarded::kOk) { return st; }
			if (fwd == Forwarded::kError || st->deferred) { return nullptr; }
			// kInconsistent: eq was rewritten since the handle was built; fall through to the triplet list itself
		}
	}
	if (num_columns <= 0) { return nullptr; }
	std::vector<int> sizes = {static_cast<int>(num_columns)};
	if (lattice) { sizes = *lattice; }
	*temp = description_from_triplets(eq, sizes);
	return temp->get();
}

std::vector<float> solve(const LinearEquation& eq, int num_columns, Precision precision, const std::vector<float>* guess, int max_iterations,
                         double tolerance, SolveStats* stats)
{
	std::shared_ptr<Structured> temp;
	Structured*                 st = description_of(eq, num_columns, &temp);
	if (!st) { return {}; }
	size_t n = 1;
	for (int s : st->sizes) { n *= static_cast<size_t>(s); }
	if (guess && guess->size() != n) { return {}; }
	fi_solve_options o;
	fi_solve_options_default(&o);
	o.precision      = static_cast<int32_t>(precision);
	o.max_iterations = max_iterations;
	o.tolerance      = tolerance;
	fi_solve_stats     fs;
	std::vector<float> x(n);
	if (fi_field_solve(st->handle, &o, guess ? guess->data() : nullptr, x.data(), FI_HOST, &fs) != FI_OK) { return {}; }
	if (stats) {
		stats->iterations        = fs.iterations;
		stats->relative_residual = fs.relative_residual;
		stats->true_residual     = fs.true_residual;
		stats->initial_residual  = fs.initial_residual;
		stats->setup_ms          = fs.setup_ms;
		stats->solve_ms          = fs.solve_ms;
		stats->converged         = fs.converged != 0;
		stats->occupied_cells    = fs.occupied_cells;
		stats->generic_rows      = fs.generic_rows;
		stats->widened_after     = fs.widened_after;
	}
	return x;
}

}  // namespace b200

// reference sparse_linear.cpp:115-152 (SimplicialLLT<float>).  A direct float factorisation has no iterative
// counterpart that is both cheaper and as robust on these systems (cond ~ n^4), so this is the fp64 PCG at a
// float-level target.
std::vector<float> solve_sparse_linear_fast(const LinearEquation& eq, int num_columns)
{
	return b200::solve(eq, num_columns, b200::Precision::kDouble, nullptr, 0, b200::g_fast_tol, nullptr);
}

// reference sparse_linear.cpp:154-184 (SimplicialLLT<double>, result cast to float)
std::vector<float> solve_sparse_linear_exact(const LinearEquation& eq, int num_columns)
{
	return b200::solve(eq, num_columns, b200::Precision::kDouble, nullptr, 0, b200::g_exact_tol, nullptr);
}

// reference sparse_linear.cpp:186-212 (BiCGSTAB<float> + diagonal preconditioner, solveWithGuess)
std::vector<float> solve_sparse_linear_with_guess(const LinearEquation& eq, const std::vector<float>& guess, int max_iterations,
                                                  float error_tolerance)
{
	return b200::solve(eq, static_cast<int>(guess.size()), b200::Precision::kFloat, &guess, max_iterations, error_tolerance, nullptr);
}

// reference sparse_linear.cpp:214-241
std::vector<float> jacobi_iterations(const LinearEquation& eq, const std::vector<float>& guess, const int num_iterations, const float weight)
{
	if (num_iterations <= 0) { return guess; }  // :220
	std::shared_ptr<b200::Structured> temp;
	b200::Structured* st = b200::description_of(eq, static_cast<long long>(guess.size()), &temp);
	if (!st) { return {}; }
	std::vector<float> x(guess.size());
	if (fi_field_jacobi(st->handle, guess.data(), num_iterations, weight, x.data()) != FI_OK) { return {}; }
	return x;
}

// reference sparse_linear.cpp:392-443
std::vector<float> solve_tiled_with_guess(const LinearEquation& eq, const std::vector<float>& guess, const std::vector<int>& sizes,
                                          const SolveOptions& options)
{
	size_t n = 1;
	for (int s : sizes) { n *= static_cast<size_t>(s); }
	if (guess.size() != n) { return {}; }  // "Incomplete guess", :402-405
	if (!options.tile && !options.cg) { return guess; }
	std::shared_ptr<b200::Structured> temp;
	b200::Structured* st = b200::description_of(eq, static_cast<long long>(n), &temp, &sizes);
	if (!st) { return {}; }
	fi_solve_options o;
	fi_solve_options_default(&o);
	o.precision      = FI_F32;
	o.max_iterations = options.max_iterations;
	o.tolerance      = options.error_tolerance;
	std::vector<float> x(n);
	// tile phase (:423-425 -> tile_solver_square :246-390), then the CG phase (:427-440), both on the device
	if (fi_field_solve_tiled(st->handle, &o, options.tile ? 1 : 0, options.tile_size, options.cg ? 1 : 0, guess.data(), x.data(), FI_HOST, nullptr,
	                         nullptr) != FI_OK) {
		return {};
	}
	return x;
}

}  // namespace field_interpolation

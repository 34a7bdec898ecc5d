// Shared state of the C++ host layer: the device-resident structured description behind a LinearEquation
// (field_interpolation::b200::Structured) and the bookkeeping that keeps it consistent with the host-side
// triplet mirror.  Not installed; include/field_interpolation/*.hpp only forward-declare Structured.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "../../include/fi_b200.h"
#include "../../include/field_interpolation/field_interpolation.hpp"

namespace field_interpolation {
namespace b200 {

struct Structured
{
	fi_field*        handle = nullptr;  // owns one LatticeField on the current device
	std::vector<int> sizes;
	bool             deferred = false;  // true: builder rows are not mirrored into eq
	// What of eq is accounted for in `handle`: the first eq_rows rows / eq_triplets triplets of eq (rows we
	// mirrored there ourselves, and caller rows already forwarded as generic rows).
	size_t eq_rows = 0, eq_triplets = 0;

	Structured() = default;
	Structured(const Structured&)            = delete;
	Structured& operator=(const Structured&) = delete;
	~Structured()
	{
		if (handle) { fi_field_destroy(handle); }
	}
};

// The structured description of `field`, created on first use; rows the caller appended to field->eq by hand
// since the last builder call are forwarded to the device (as generic rows) so row order is preserved.
// Returns nullptr when the library reports an error (b200::last_error() has the text).
Structured* structured_for_append(LatticeField* field);

// Forwards rows of `eq` beyond what `st` already accounts for (as generic rows).  kInconsistent: eq is no longer an
// extension of what the handle holds (it was cleared, truncated or re-assigned) — nothing was forwarded and the
// handle must not be used for this eq; kError: the library refused the rows (handle unchanged).
enum class Forwarded { kOk, kInconsistent, kError };
Forwarded forward_tail_rows(const LinearEquation& eq, Structured* st);

// An independent deep copy of a description (fi_field_clone); nullptr on error.
std::shared_ptr<Structured> clone_description(const Structured& src);

// A fresh description of the lattice `sizes` holding all of eq as generic rows; nullptr on error.
std::shared_ptr<Structured> description_from_triplets(const LinearEquation& eq, const std::vector<int>& sizes);

// After a builder call on st->handle: mirrors the rows added since `rows_before` into eq (unless deferred).
bool mirror_new_rows(LinearEquation* eq, Structured* st, long long rows_before, long long trips_before);

}  // namespace b200
}  // namespace field_interpolation

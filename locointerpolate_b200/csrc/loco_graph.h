// loco_graph.h -- host-side image of an LCSampledFunction object graph.
//
// This is the boundary object of the build: exactly the information the reference keeps in
//   LCSampledFunction   (LCAdaptiveSampling/LCSampledFunction.h:29-35: root_, ranges_, samples_, borderCells_)
//   LCAdaptiveGridCell  (LCAdaptiveGridCell.h:86-109: ranges_, isLeaf_, children, splitDirection_, splitVal_,
//                        centerShapeInfo_, samples_ = map sample -> per-cell value copy)
//   LCSample            (LCSample.h:60-64: center_, shapeInfo_, basisFunctions_)
//   LCBasisFunction     (LCBasisFunction.h:43-47: center_, support_, weight_; type enum :19)
// and therefore exactly what a PrecomputedParametricShape proto file carries
// (LCFunction/PrecomputedParametricShape.proto, LCProtoConverter.cpp:141-168 / 356-403).
// Pointers become indices with the canonical numbering of SURVEY.md 8(a12): tree cells in pre-order
// (child1 first), samples by position in samples_.  Neighbour lists are NOT stored: like the
// reference's loader, the flattener re-derives them (loco_flatten.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace loco {

enum BasisType : int32_t { LINEAR_BSPLINE = 0, CUBIC_BSPLINE = 1 };  // LCBasisFunction.h:19

struct Graph {
	int32_t N = 0;          // number of parameters
	int32_t D = 1;          // doubles per function value (1 = LCRealFunctionValue, 3V = V-vertex mesh)
	int32_t basis = LINEAR_BSPLINE;
	std::vector<double> domain;  // [2N] min,max per parameter (LCSampledFunction::ranges_)

	// value table: rows of D doubles.  Samples and per-cell copies point into it, so objects that
	// share one LCFunctionValue (cubic border samples, LCAdaptiveGrid.cpp:345-355) share a row.
	// `values` may be left empty with virtual_rows set: the table then lives only on the device and is
	// filled there (synthetic mesh tables of tens of GB are never materialised on the host).
	std::vector<double> values;        // [R*D]
	int64_t virtual_rows = 0;
	int64_t n_rows() const { return virtual_rows ? virtual_rows : (D ? (int64_t)values.size() / D : 0); }

	// samples
	std::vector<double> sample_center;  // [S*N]
	std::vector<int32_t> sample_row;    // [S]
	std::vector<int64_t> bf_off;        // [S+1] CSR into the basis-function arrays
	std::vector<double> bf_center;      // [B*N]
	std::vector<double> bf_support;     // [B*N]
	std::vector<double> bf_weight;      // [B]
	int64_t n_samples() const { return (int64_t)sample_row.size(); }

	// cells: tree cells first in pre-order (index 0 = root), then border cells (all leaves, level -1)
	struct CellSet {
		std::vector<double> ranges;       // [C*2N] lo,hi per dimension
		std::vector<int32_t> split_dir;   // [C] -1 for a leaf
		std::vector<double> split_val;    // [C]
		std::vector<int32_t> child2;      // [C] index of child2 (child1 is always index+1); -1 for a leaf
		std::vector<int32_t> center_row;  // [C] value row of centerShapeInfo_ (-1: none / inner node)
		std::vector<int64_t> ls_off;      // [C+1] CSR: the cell's samples_ map
		std::vector<int32_t> ls_sample;   // sample index
		std::vector<int32_t> ls_row;      // value row of the per-cell copy
		int64_t size() const { return (int64_t)split_dir.size(); }
	};
	CellSet tree;
	CellSet border;
};

// Error convention: like LCError (LCUtils/LCError.h:7-20) -- ok flag + description, returned by value.
struct Status {
	bool ok = true;
	std::string msg = "no error";
	Status() {}
	explicit Status(const std::string &m) : ok(false), msg(m) {}
};
#define LOCO_RETURN_IF_ERROR(expr) do { ::loco::Status _s = (expr); if (!_s.ok) return _s; } while (0)

Status validate_graph(const Graph &g);

} // namespace loco

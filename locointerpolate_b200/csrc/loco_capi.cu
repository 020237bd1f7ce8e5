// loco_capi.cu -- implementation of the C ABI declared in include/loco_b200.h.
//
// Host orchestration only: builds the Graph (from a proto file, a plain-array graph or the bootstrap
// generator), flattens it, uploads the flat model once, and drives the kernels.  There is no CPU
// evaluation path in this library: if no CUDA device is usable every constructor fails.
#include "../../include/loco_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "loco_bootstrap.h"
#include "loco_device.h"
#include "loco_flat.h"
#include "loco_cellpoly.h"
#include "loco_proto.h"

using namespace loco;

namespace loco {
KernelProfile &kernel_profile()
{
	static KernelProfile p;
	return p;
}
void KernelProfile::push(const char *name, cudaEvent_t a, cudaEvent_t b)
{
	if (n == cap) {
		cap = cap ? cap * 2 : 1024;
		recs = static_cast<Rec *>(realloc(recs, sizeof(Rec) * (size_t)cap));
	}
	recs[n++] = Rec{name, a, b};
}
} // namespace loco

namespace {

thread_local std::string g_ctor_error = "no error";

const int kSlots = 3;  // pipeline depth of the host-buffer path

struct Workspace {
	cudaStream_t stream = nullptr;
	void *q = nullptr, *out = nullptr, *leaf = nullptr, *status = nullptr, *W = nullptr, *truth = nullptr, *misc = nullptr, *sorted = nullptr;
	void *sortedx[3] = {nullptr, nullptr, nullptr};
	void *h_q = nullptr, *h_out = nullptr, *h_leaf = nullptr, *h_status = nullptr;  // page-locked staging of the host-buffer path (pageable callers)
	size_t h_q_b = 0, h_out_b = 0, h_leaf_b = 0, h_status_b = 0;
	void *woff = nullptr;  // weight-block offsets of the generic mesh path
	size_t woff_b = 0;  // workspaces of the additional streams of the sorted path
	size_t q_b = 0, out_b = 0, leaf_b = 0, status_b = 0, W_b = 0, truth_b = 0, misc_b = 0, sorted_b = 0, sortedx_b[3] = {0, 0, 0};
};

} // namespace

struct loco_model {
	int device = 0;
	int64_t n_dterms = 0;    // derivative-only terms of the linear basis over all leaves (Flat::dterm_*)
	bool host_only = false;  // device == LOCO_HOST_ONLY: introspection only, every evaluation call fails
	bool initial_upload = false;
	Graph graph;
	Flat flat;
	DeviceModel dm{};
	std::vector<void *> allocs;
	double *d_values = nullptr;
	double *d_term_wv = nullptr;
	double *d_mterm_wv = nullptr;
	double *d_eterm_wv = nullptr;
	const int32_t *d_ecell_off = nullptr;
	const int32_t *d_epoly_idx = nullptr, *d_epoly_cell = nullptr;  // polynomial evaluation cells (option "ecell_poly")
	double *d_epoly = nullptr;
	int32_t n_epoly = 0;
	double *d_tl_block = nullptr;
	double *d_tl_poly = nullptr;  // per-leaf polynomial blocks (option "leaf_poly")
	float *d_tl_vals_f32 = nullptr;
	int64_t n_terms = 0;
	int sm_count = 148;
	int64_t device_bytes = 0;
	mutable std::string err = "no error";
	int64_t launches = 0;
	int64_t opt_path = 0, opt_fp32 = 0, opt_lut = 1, opt_chunk = 0, opt_mesh_simt = 0, opt_overlap = 3, opt_uniform = 1, opt_fused_errgen = 1, opt_leaf_poly = 1, opt_adaptive = 1, opt_aggregate = 1, opt_bulk_eval = 0, opt_fast_scatter = 1;
	const uint32_t *d_lut = nullptr;  // kept so that the "lut" option can switch the locate grid off and on
	Workspace ws[kSlots];
	// auxiliary streams: consecutive chunks of the sorted path run round-robin on them so that the latency-bound sort
	// kernels of one chunk overlap the evaluation of another
	cudaStream_t aux[4] = {nullptr, nullptr, nullptr, nullptr};
	cudaEvent_t ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
	// skew statistics (host-mapped, one entry per internal stream): see SkewStats in loco_device.h
	SkewStats *h_stats = nullptr, *d_stats = nullptr;
	long long seen_overflow = 0, seen_total = 0;
	bool skewed = false;  // the recent batches overflowed their leaf buckets: scalar batches take the exact counting sort
};

namespace {

int fail(const loco_model *m, int code, const std::string &msg)
{
	if (m) m->err = msg;
	else g_ctor_error = msg;
	return code;
}

#define NEED_DEVICE(m) do { if ((m)->host_only) return fail(m, LOCO_ERR_CUDA, "model is host-only (LOCO_HOST_ONLY): evaluation needs a CUDA device, there is no CPU path"); } while (0)

#define CUDA_OK(m, expr)                                                                                   \
	do {                                                                                                   \
		cudaError_t _e = (expr);                                                                           \
		if (_e != cudaSuccess) return fail(m, LOCO_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(_e)); \
	} while (0)

template <typename T>
int upload(loco_model *m, const std::vector<T> &h, const T **dptr, size_t min_elems = 1)
{
	size_t n = std::max(h.size(), min_elems);
	void *d = nullptr;
	CUDA_OK(m, cudaMalloc(&d, n * sizeof(T)));
	m->allocs.push_back(d);
	m->device_bytes += (int64_t)(n * sizeof(T));
	if (!h.empty()) CUDA_OK(m, cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
	*dptr = static_cast<const T *>(d);
	return LOCO_OK;
}

int grow(loco_model *m, void **p, size_t *have, size_t need)
{
	if (need <= *have) return LOCO_OK;
	if (*p) { CUDA_OK(m, cudaFree(*p)); *p = nullptr; *have = 0; }
	need = need + need / 4 + 256;
	CUDA_OK(m, cudaMalloc(p, need));
	*have = need;
	return LOCO_OK;
}

template <typename T>
std::vector<int32_t> narrow(loco_model *m, const std::vector<T> &v, bool *ok)
{
	std::vector<int32_t> r(v.size());
	for (size_t i = 0; i < v.size(); i++) {
		if (v[i] > (T)0x7fffffff) *ok = false;
		r[i] = (int32_t)v[i];
	}
	(void)m;
	return r;
}

// scalar functions: term_wv[t] = weight_t * value(sample of t), used by the leaf-sorted tile path
int refresh_term_wv(loco_model *m)
{
	if (m->dm.D == 1) {
		m->launches += launch_update_term_wv(m->dm, m->d_term_wv, m->n_terms, nullptr);
		m->launches += launch_update_mterm_wv(m->dm, m->d_mterm_wv, nullptr);
		m->launches += launch_update_eterm_wv(m->dm, m->d_eterm_wv, nullptr);
		m->launches += launch_update_epoly(m->dm, m->d_epoly_cell, m->n_epoly, m->d_epoly, nullptr);
		m->launches += launch_update_tensor_vals(m->dm, m->d_tl_block, m->d_tl_vals_f32, nullptr);
		m->launches += launch_update_tensor_poly(m->dm, m->d_tl_poly, m->dm.tp_stride, nullptr);  // from the refreshed leaf values
	}
	CUDA_OK(m, cudaGetLastError());
	CUDA_OK(m, cudaDeviceSynchronize());
	return LOCO_OK;
}

// Host copy of the value table (Graph::values: what loco_model_save_proto serialises and loco_eval_cell_poly reads).
// The device table is the authority once the caller replaces values; tables up to kHostValueBytes keep a host mirror,
// larger ones (mesh tables of several GB) live on the device only and are read back on demand.
const size_t kHostValueBytes = (size_t)1 << 30;

void drop_host_values(loco_model *m)
{
	std::vector<double>().swap(m->graph.values);
	m->graph.virtual_rows = m->dm.R;
}

// make Graph::values current again (after loco_model_fill_values_synthetic or a large loco_model_set_values)
int sync_values_to_host(const loco_model *cm)
{
	loco_model *m = const_cast<loco_model *>(cm);
	if (!m->graph.values.empty() || m->host_only || !m->d_values) return LOCO_OK;
	const DeviceModel &dm = m->dm;
	const size_t n = (size_t)dm.R * dm.D;
	if (n * sizeof(double) > ((size_t)2 << 30))
		return fail(m, LOCO_ERR_RANGE, "value table lives on the device and is larger than 2 GiB: too large for a protobuf message");
	CUDA_OK(m, cudaSetDevice(m->device));
	m->graph.values.resize(n);
	CUDA_OK(m, cudaMemcpy2D(m->graph.values.data(), (size_t)dm.D * sizeof(double), m->d_values, (size_t)dm.value_stride * sizeof(double),
	                        (size_t)dm.D * sizeof(double), (size_t)dm.R, cudaMemcpyDeviceToHost));
	m->graph.virtual_rows = 0;
	return LOCO_OK;
}

int build_device_model(loco_model *m)
{
	if (m->device == LOCO_HOST_ONLY) {
		m->host_only = true;
		m->dm.N = m->flat.N; m->dm.D = m->flat.D; m->dm.L = (int32_t)m->flat.L; m->dm.R = m->flat.R;
		return LOCO_OK;
	}
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return fail(m, LOCO_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(e));
	if (m->device < 0 || m->device >= ndev) return fail(m, LOCO_ERR_CUDA, "device index out of range");
	CUDA_OK(m, cudaSetDevice(m->device));
	Flat &f = m->flat;
	if (f.N > kMaxN) return fail(m, LOCO_ERR_INVALID, "more than 16 parameters are not supported");
	if (f.L > 0x7fffffff || f.C > 0x7fffffff) return fail(m, LOCO_ERR_INVALID, "tree too large");
	DeviceModel &dm = m->dm;
	dm.N = f.N; dm.D = f.D; dm.basis = f.basis; dm.exact_rcp = f.exact_rcp ? 1 : 0;
	dm.L = (int32_t)f.L; dm.C = (int32_t)f.C; dm.n_inner = (int32_t)f.cells.size(); dm.max_support = std::max(f.max_support, 1); dm.max_terms = f.max_terms;
	dm.R = f.R;
	dm.value_stride = f.D == 1 ? 1 : ((f.D + 1) / 2) * 2;  // rows 16-byte aligned for 128-bit loads
	bool ok = true;
	std::vector<int32_t> sup_off = narrow(m, f.sup_off, &ok), term_off = narrow(m, f.term_off, &ok);
	if (!ok) return fail(m, LOCO_ERR_INVALID, "support / term lists exceed 2^31 entries");
	int rc;
	if ((rc = upload(m, f.dom_min, &dm.dom_min))) return rc;
	if ((rc = upload(m, f.dom_max, &dm.dom_max))) return rc;
	if ((rc = upload(m, f.cells, &dm.cells))) return rc;
	if ((rc = upload(m, f.leaf_lo_eps, &dm.leaf_lo_eps))) return rc;
	if ((rc = upload(m, f.leaf_hi_eps, &dm.leaf_hi_eps))) return rc;
	if ((rc = upload(m, f.leaf_lo, &dm.leaf_lo))) return rc;
	if ((rc = upload(m, f.leaf_hi, &dm.leaf_hi))) return rc;
	if ((rc = upload(m, f.leaf_status, &dm.leaf_status))) return rc;
	if ((rc = upload(m, f.leaf_center_row, &dm.leaf_center_row))) return rc;
	if ((rc = upload(m, sup_off, &dm.sup_off))) return rc;
	if ((rc = upload(m, f.sup_row, &dm.sup_row))) return rc;
	if ((rc = upload(m, term_off, &dm.term_off))) return rc;
	if ((rc = upload(m, f.term_cs, &dm.term_cs, 2))) return rc;
	if ((rc = upload(m, f.term_w, &dm.term_w))) return rc;
	if ((rc = upload(m, f.term_slot, &dm.term_slot))) return rc;
	if ((rc = upload(m, f.term_row, &dm.term_row))) return rc;
	dm.mterm_off = nullptr; dm.n_mterms = 0;
	if (!f.mterm_off.empty()) {
		std::vector<int32_t> moff = narrow(m, f.mterm_off, &ok), roff = narrow(m, f.mrun_off, &ok);
		if (!ok) return fail(m, LOCO_ERR_INVALID, "support / term lists exceed 2^31 entries");
		dm.n_mterms = f.mterm_off.back();
		if ((rc = upload(m, moff, &dm.mterm_off))) return rc;
		if ((rc = upload(m, f.mterm_cs, &dm.mterm_cs, 2))) return rc;
		if ((rc = upload(m, roff, &dm.mrun_off))) return rc;
		if ((rc = upload(m, f.mrun_term, &dm.mrun_term))) return rc;
		std::vector<double> zeros((size_t)dm.n_mterms + 2, 0.0);
		const double *dwv = nullptr;
		if ((rc = upload(m, zeros, &dwv))) return rc;
		m->d_mterm_wv = const_cast<double *>(dwv);
		dm.mterm_wv = dwv;
		std::vector<double>().swap(f.mterm_cs);
		std::vector<int32_t>().swap(f.mrun_term);
	}
	dm.ecell_off = nullptr; dm.eterm_src = nullptr; dm.eterm_wv = nullptr; dm.n_ecells = 0; dm.n_eterms = 0;
	dm.epoly_idx = nullptr; dm.epoly = nullptr; dm.epoly_stride = 0;
	if (!f.ecell_off.empty() && dm.mterm_off) {
		std::vector<int32_t> coff = narrow(m, f.ecell_off, &ok), toff = narrow(m, f.eterm_off, &ok);
		if (!ok) return fail(m, LOCO_ERR_INVALID, "support / term lists exceed 2^31 entries");
		dm.n_ecells = (int32_t)f.ecell_off.back();
		dm.n_eterms = f.eterm_off.back();
		if ((rc = upload(m, coff, &dm.ecell_off))) return rc;
		if ((rc = upload(m, f.ecell_bits, &dm.ecell_bits))) return rc;
		if ((rc = upload(m, toff, &dm.eterm_off))) return rc;
		if ((rc = upload(m, f.eterm_cs, &dm.eterm_cs, 2))) return rc;
		if ((rc = upload(m, f.eterm_src, &dm.eterm_src))) return rc;
		std::vector<double> zeros((size_t)dm.n_eterms + 2, 0.0);
		const double *dwv = nullptr;
		if ((rc = upload(m, zeros, &dwv))) return rc;
		m->d_eterm_wv = const_cast<double *>(dwv);
		dm.eterm_wv = dwv;
		if (!f.epoly_cell.empty()) {
			const int F = f.basis == CUBIC_BSPLINE ? 4 : 2;
			int P = 1;
			for (int i = 0; i < dm.N; i++) P *= F;
			dm.epoly_stride = 2 * dm.N + P;
			m->n_epoly = (int32_t)f.epoly_cell.size();
			std::vector<double> blocks((size_t)m->n_epoly * dm.epoly_stride, 0.0);
			for (int32_t i = 0; i < m->n_epoly; i++)
				std::copy(f.epoly_geom.begin() + (size_t)i * 2 * dm.N, f.epoly_geom.begin() + (size_t)(i + 1) * 2 * dm.N, blocks.begin() + (size_t)i * dm.epoly_stride);
			const double *dblk = nullptr;
			if ((rc = upload(m, f.epoly_idx, &m->d_epoly_idx))) return rc;
			if ((rc = upload(m, f.epoly_cell, &m->d_epoly_cell))) return rc;
			if ((rc = upload(m, blocks, &dblk))) return rc;
			m->d_epoly = const_cast<double *>(dblk);
			dm.epoly = dblk;
			dm.epoly_idx = m->d_epoly_idx;  // option "ecell_poly", on by default (round 2): cells no knot crosses evaluate their one polynomial
		}
		// the evaluation-cell term lists stay on the host while they are small: loco_model_update_from_graph reuses them for
		// leaves a refinement pass did not touch (Flat::leaf_sig); beyond 256 MB they are released and an update recomputes
		if (f.eterm_cs.size() * sizeof(double) > ((size_t)256 << 20)) {
			std::vector<double>().swap(f.eterm_cs);
			std::vector<int32_t>().swap(f.eterm_src);
		}
	}
	std::vector<int32_t> dterm_off = narrow(m, f.dterm_off, &ok);
	m->n_dterms = (int64_t)f.dterm_w.size();
	if (!ok) return fail(m, LOCO_ERR_INVALID, "support / term lists exceed 2^31 entries");
	if ((rc = upload(m, dterm_off, &dm.dterm_off))) return rc;
	if ((rc = upload(m, f.dterm_cs, &dm.dterm_cs, 2))) return rc;
	if ((rc = upload(m, f.dterm_w, &dm.dterm_w))) return rc;
	if ((rc = upload(m, f.dterm_row, &dm.dterm_row))) return rc;
	// value table
	size_t vbytes = (size_t)std::max<int64_t>(f.R, 1) * dm.value_stride * sizeof(double);
	CUDA_OK(m, cudaMalloc((void **)&m->d_values, vbytes));
	m->allocs.push_back(m->d_values);
	m->device_bytes += (int64_t)vbytes;
	CUDA_OK(m, cudaMemset(m->d_values, 0, vbytes));
	dm.values = m->d_values;
	m->n_terms = (int64_t)f.term_w.size();
	CUDA_OK(m, cudaMalloc((void **)&m->d_term_wv, (size_t)(m->n_terms + 2) * sizeof(double)));
	m->allocs.push_back(m->d_term_wv);
	m->device_bytes += (int64_t)((m->n_terms + 2) * sizeof(double));
	CUDA_OK(m, cudaMemset(m->d_term_wv, 0, (size_t)(m->n_terms + 2) * sizeof(double)));
	dm.term_wv = m->d_term_wv;
	CUDA_OK(m, cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, m->device));
	if (f.D > 1 && getenv("LOCO_L2_PERSIST_MB")) {
		// EXPERIMENT switch (off by default).  The mesh tile kernels re-read the slice of the value table all CTAs share (TMA
		// copies with an L2::evict_last policy) while the result stream passes through the same L2, and 69-75 % of those
		// re-reads miss (ncu r02h).  A persisting-L2 set-aside for the evict_last lines was measured and made it WORSE
		// (c3: 17.7 ms without, 18.5 ms with 48 MB, 29.6 ms with 96 MB; c4b unchanged): the carve-out is taken from the
		// capacity that buffers the dirty result lines.  See profiles/README.md.
		int max_persist = 0;
		cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, m->device);
		const size_t want = (size_t)atoi(getenv("LOCO_L2_PERSIST_MB")) << 20;
		if (max_persist > 0 && want > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>((size_t)max_persist, want));
		cudaGetLastError();
	}
	// tensor-product leaf blocks
	dm.tensor_F = 0;
	if (f.tensor_F > 0 && f.N <= kMaxTemplateN) {
		const int F = f.tensor_F, K = f.tensor_K, N = f.N;
		dm.tensor_F = F; dm.tensor_K = K;
		dm.tensor_uniform = (f.tensor_uniform && m->opt_uniform) ? 1 : 0;
		dm.tl_voff = (N * F + N + 1) & ~1;
		dm.tl_stride = f.D == 1 ? ((dm.tl_voff + K + 1) & ~1) : dm.tl_voff;
		std::vector<double> blk((size_t)f.L * dm.tl_stride, 0.0);
		for (int64_t l = 0; l < f.L; l++) {
			double *b = &blk[(size_t)l * dm.tl_stride];
			std::copy(f.tl_center.begin() + l * N * F, f.tl_center.begin() + (l + 1) * N * F, b);
			std::copy(f.tl_rcp.begin() + l * N, f.tl_rcp.begin() + (l + 1) * N, b + N * F);
		}
		const double *dblk = nullptr;
		if ((rc = upload(m, blk, &dblk))) return rc;
		m->d_tl_block = const_cast<double *>(dblk);
		dm.tl_block = dblk;
		if ((rc = upload(m, f.tl_row, &dm.tl_row))) return rc;
		dm.tl_vals_f32 = nullptr;
		if (f.D == 1) {  // fp32 copy of the leaf values for the optional fp32 mode
			std::vector<float> zeros((size_t)f.L * K, 0.0f);
			const float *dv = nullptr;
			if ((rc = upload(m, zeros, &dv))) return rc;
			m->d_tl_vals_f32 = const_cast<float *>(dv);
			dm.tl_vals_f32 = dv;
		}
		// per-leaf polynomial form for scalar functions: leaves inside one polynomial piece of every basis function;
		// K <= 256 keeps the monomial form well inside the 1e-12 bar (cubic N <= 4, linear N <= 8)
		dm.tl_poly = nullptr; dm.tp_stride = 0;
		if (f.D == 1 && f.tensor_poly && K <= 256) {
			dm.tp_stride = (2 * N + K + 1) & ~1;
			std::vector<double> zeros((size_t)f.L * dm.tp_stride, 0.0);
			const double *dp = nullptr;
			if ((rc = upload(m, zeros, &dp))) return rc;
			m->d_tl_poly = const_cast<double *>(dp);
			if (m->opt_leaf_poly) dm.tl_poly = dp;
		}
		std::vector<int32_t> fold_off = narrow(m, f.fold_off, &ok);
		if ((rc = upload(m, fold_off, &dm.fold_off))) return rc;
		if ((rc = upload(m, f.fold_row, &dm.fold_row))) return rc;
		if ((rc = upload(m, f.fold_map, &dm.fold_map))) return rc;
		if ((rc = upload(m, f.fold_n, &dm.fold_n))) return rc;
	}
	dm.dom_exact_mask = 0;
	for (int d = 0; d < f.N; d++) {
		dm.dom_lo[d] = f.dom_min[(size_t)d];
		dm.dom_w[d] = f.dom_max[(size_t)d] - f.dom_min[(size_t)d];  // LCFunction.cpp:66: getMaxRange(i) - getMinRange(i)
		int e;
		const bool pow2 = dm.dom_w[d] > 0 && std::isfinite(dm.dom_w[d]) && std::frexp(dm.dom_w[d], &e) == 0.5 && std::isnormal(dm.dom_w[d]) &&
		                  std::isnormal(1.0 / dm.dom_w[d]);
		dm.dom_rcp[d] = pow2 ? 1.0 / dm.dom_w[d] : 0.0;
		if (pow2) dm.dom_exact_mask |= 1u << d;
	}
	dm.lut = nullptr;
	dm.lut_uniform_mask = 0;
	if (!f.lut.empty() && m->opt_lut) {
		if ((rc = upload(m, f.lut, &dm.lut))) return rc;
		if ((rc = upload(m, f.lut_bounds, &dm.lut_bounds))) return rc;
		for (int d = 0; d < f.N; d++) {
			dm.lut_bits[d] = f.lut_bits[d];
			dm.lut_boff[d] = f.lut_boff[d];
			const double lo = f.lut_bounds[(size_t)f.lut_boff[d]], hi = f.lut_bounds[(size_t)f.lut_boff[d] + ((size_t)1 << f.lut_bits[d])];
			dm.lut_lo[d] = lo;
			dm.lut_inv[d] = (double)((int64_t)1 << f.lut_bits[d]) / (hi - lo);
			// uniform dyadic boundaries?  h and 1/h exact powers of two and every stored boundary == fma(i, h, lo) bit for bit
			const double h = (hi - lo) / (double)((int64_t)1 << f.lut_bits[d]);
			int e2;
			bool uni = f.lut_bits[d] > 0 && h > 0 && std::frexp(h, &e2) == 0.5 && std::isnormal(h) && std::isnormal(1.0 / h) && dm.lut_inv[d] == 1.0 / h;
			for (int64_t i = 0; uni && i <= ((int64_t)1 << f.lut_bits[d]); i++)
				uni = f.lut_bounds[(size_t)f.lut_boff[d] + (size_t)i] == std::fma((double)i, h, lo);
			dm.lut_h[d] = h;
			if (uni) dm.lut_uniform_mask |= 1u << d;
		}
	}
	dm.path_consistent = f.path_consistent ? 1 : 0;
	dm.any_leaf_status = f.any_leaf_status ? 1 : 0;
	for (int i = 0; i < f.N; i++) {
		dm.root_lo_eps[i] = m->graph.tree.ranges[2 * i] - kEpsilon;
		dm.root_hi_eps[i] = m->graph.tree.ranges[2 * i + 1] + kEpsilon;
	}
	if (!f.values.empty()) {
		m->initial_upload = true;  // Graph::values already holds exactly these rows
		rc = loco_model_set_values(m, f.values.data());
		m->initial_upload = false;
		if (rc) return rc;
		std::vector<double>().swap(f.values);  // Graph::values is the one host copy
	}
	// the big term arrays are only needed on the device
	std::vector<double>().swap(f.term_cs);
	std::vector<double>().swap(f.term_w);
	std::vector<int32_t>().swap(f.term_slot);
	std::vector<int32_t>().swap(f.term_row);
	std::vector<double>().swap(f.dterm_cs);
	std::vector<double>().swap(f.dterm_w);
	std::vector<int32_t>().swap(f.dterm_row);
	if (m->ws[0].stream) return LOCO_OK;  // re-upload of an existing handle (loco_model_update_from_graph): streams, events, counters stay
	for (int s = 0; s < kSlots; s++) CUDA_OK(m, cudaStreamCreateWithFlags(&m->ws[s].stream, cudaStreamNonBlocking));
	for (int k = 0; k < 4; k++) {
		CUDA_OK(m, cudaStreamCreateWithFlags(&m->aux[k], cudaStreamNonBlocking));
		CUDA_OK(m, cudaEventCreateWithFlags(&m->ev_join[k], cudaEventDisableTiming));
	}
	CUDA_OK(m, cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
	CUDA_OK(m, cudaHostAlloc((void **)&m->h_stats, 4 * sizeof(SkewStats), cudaHostAllocMapped));
	memset(m->h_stats, 0, 4 * sizeof(SkewStats));
	CUDA_OK(m, cudaHostGetDevicePointer((void **)&m->d_stats, m->h_stats, 0));
	return LOCO_OK;
}

int finish_model(loco_model *m, int device, loco_model_t **out)
{
	m->device = device;
	Status s = flatten(m->graph, &m->flat);
	if (!s.ok) { int rc = fail(nullptr, LOCO_ERR_INVALID, s.msg); loco_free(m); return rc; }
	int rc = build_device_model(m);
	if (rc) { g_ctor_error = m->err; loco_free(m); return rc; }
	*out = m;
	return LOCO_OK;
}

// scalar or mesh evaluation of Q queries resident on the device, using workspace slot `w`
// deriv_dir >= 0 (scalar functions only): LCSampledFunction::evalDeriv along that cube direction instead of the value
int eval_device(loco_model *m, Workspace &w, cudaStream_t st, bool in_cube, const double *q, int64_t Q, double *out, int64_t out_stride,
                int32_t *leaf, uint8_t *status, int deriv_dir = -1)
{
	const DeviceModel &dm = m->dm;
	if (Q <= 0) return LOCO_OK;
	if (dm.D == 1) {
		// leaf-sorted tiles pay off once a leaf bucket fills a good part of a 256-query tile
		// and the per-leaf data no longer sits in the L1 anyway
		const int64_t leaf_bytes = dm.tensor_F > 0 ? (int64_t)dm.L * dm.tl_stride * 8 : m->n_terms * (16 * dm.N + 8);
		const bool big = Q >= (int64_t)dm.L * 24 && Q >= (1 << 16) && leaf_bytes > 96 * 1024;
		const bool sorted = !in_cube && (m->opt_path == 4 || (m->opt_path == 0 && big));
		const bool tiles = !in_cube && m->opt_path == 2;
		// (derivatives with the linear basis: only while no leaf has derivative-only terms -- Flat::dterm_*, hats that lie right of
		//  a leaf and still count in the reference's signed comparison; the leaf blocks do not hold them)
		bool bucket = !in_cube && bucket_path_supported(dm) && (m->opt_path == 5 || (m->opt_path == 0 && big && dm.L >= 256)) &&
		              (deriv_dir < 0 || dm.basis == CUBIC_BSPLINE || m->n_dterms == 0);
		if (bucket && m->opt_path == 0 && m->opt_adaptive && m->h_stats) {
			// Automatic dispatch between the single-pass leaf buckets (capacity sized for a uniform batch) and the exact counting
			// sort: the kernels of both paths add {queries beyond a bucket's capacity, queries seen} to host-mapped counters, and
			// the counters of the batches evaluated SO FAR decide this call (no synchronisation: a stale value only delays the
			// switch by a call).  Skewed batches -- optimisation trajectories, lattice sweeps in a few leaves -- would overflow
			// their buckets into the thread-per-query fallback; the counting sort has no capacity and evaluates them in leaf order.
			long long ovf = 0, tot = 0;
			for (int k = 0; k < 4; k++) { ovf += ((volatile SkewStats *)m->h_stats)[k].overflow; tot += ((volatile SkewStats *)m->h_stats)[k].total; }
			const long long d_ovf = ovf - m->seen_overflow, d_tot = tot - m->seen_total;
			if (d_tot >= (1 << 20)) {
				const double frac = (double)d_ovf / (double)d_tot;
				// measured (profiles/README.md, round 2): with half of a batch overflowing (clustered queries) the buckets + overflow list
				// still beat the counting sort (11.0 vs 8.4 G q/s), with ~90 % overflowing (lattice sweeps) the sort wins (23.4 vs 21.3)
				if (!m->skewed && frac > 0.65) m->skewed = true;
				else if (m->skewed && frac < 0.30) m->skewed = false;
				m->seen_overflow = ovf; m->seen_total = tot;
			}
			if (m->skewed) bucket = false;  // `sorted` is set for every batch that qualifies for the buckets
		}
		// derivatives on generic leaves with the cubic basis: the sorted path with the derivative factor in its term lists
		const bool sorted_deriv = deriv_dir >= 0 && !bucket && sorted && dm.basis == CUBIC_BSPLINE && dm.tensor_F == 0;
		if (deriv_dir >= 0 && !bucket && !sorted_deriv) {
			// derivatives outside the bucket path (generic leaves with the linear basis, small batches): one thread per query
			m->launches += launch_eval_deriv_direct(dm, q, Q, deriv_dir, out, leaf, status, st);
			CUDA_OK(m, cudaGetLastError());
			return LOCO_OK;
		}
		if (bucket) {
			// single-pass fixed-capacity buckets: chunks of ~4M queries, at least 32 per leaf on average
			int64_t chunk = m->opt_chunk > 0 ? m->opt_chunk : std::max<int64_t>((int64_t)1 << 22, (int64_t)dm.L * 32);
			chunk = std::min(chunk, Q);
			int rc;
			if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
			if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
			const int64_t n_chunks = (Q + chunk - 1) / chunk;
			const int ns = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)4, m->opt_overlap, n_chunks}));
			BucketWork bw[4];
			cudaStream_t run_on[4] = {st, st, st, st};
			for (int k = 0; k < ns; k++) {
				void **buf = k == 0 ? &w.sorted : &w.sortedx[k - 1];
				size_t *have = k == 0 ? &w.sorted_b : &w.sortedx_b[k - 1];
				if ((rc = grow(m, buf, have, bucket_work_bytes(dm, chunk)))) return rc;
				bw[k] = carve_bucket_work(dm, chunk, *buf);
				if (m->opt_path == 0 && deriv_dir < 0) bw[k].stats = m->d_stats + k;
				bw[k].aggregate = m->opt_aggregate != 0;
				bw[k].bulk = m->opt_bulk_eval != 0;
				bw[k].fast = m->opt_fast_scatter != 0;
			}
			if (ns > 1) {
				CUDA_OK(m, cudaEventRecord(m->ev_fork, st));
				for (int k = 0; k < ns; k++) { CUDA_OK(m, cudaStreamWaitEvent(m->aux[k], m->ev_fork, 0)); run_on[k] = m->aux[k]; }
			}
			for (int k = 0; k < ns; k++) launch_bucket_reset(dm, bw[k], run_on[k]);
			int k = 0, used[4] = {0, 0, 0, 0};
			for (int64_t o = 0; o < Q; o += chunk, k = (k + 1) % ns) {
				const int64_t n = std::min(chunk, Q - o);
				m->launches += launch_eval_scalar_bucket_chunk(dm, q + o * dm.N, n, out + o, leaf + o, status + o, bw[k], m->sm_count, m->opt_fp32 != 0, used[k]++,
				                                               run_on[k], deriv_dir);
			}
			if (ns > 1)
				for (int j = 0; j < ns; j++) { CUDA_OK(m, cudaEventRecord(m->ev_join[j], m->aux[j])); CUDA_OK(m, cudaStreamWaitEvent(st, m->ev_join[j], 0)); }
		} else if (sorted) {
			// chunks sized so that records (32 B/query) + the out slice stay L2 resident; never fewer than ~16 queries per leaf
			int64_t chunk = m->opt_chunk > 0 ? m->opt_chunk : std::max<int64_t>((int64_t)1 << 22, (int64_t)(dm.ecell_off ? dm.n_ecells : dm.L) * 16);
			chunk = std::min(chunk, Q);
			int rc;
			if ((rc = grow(m, &w.sorted, &w.sorted_b, sorted_work_bytes(dm, chunk)))) return rc;
			if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
			if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
			const int64_t n_chunks = (Q + chunk - 1) / chunk;
			const int ns = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)4, m->opt_overlap, n_chunks}));
			SortedWork sw[4];
			cudaStream_t run_on[4] = {st, st, st, st};
			sw[0] = carve_sorted_work(dm, w.sorted);
			for (int k = 1; k < ns; k++) {
				if ((rc = grow(m, &w.sortedx[k - 1], &w.sortedx_b[k - 1], sorted_work_bytes(dm, chunk)))) return rc;
				sw[k] = carve_sorted_work(dm, w.sortedx[k - 1]);
			}
			for (int k = 0; k < ns; k++) sw[k].aggregate = m->opt_aggregate != 0;
			if (m->opt_path == 0 && deriv_dir < 0 && m->d_stats && bucket_path_supported(dm) && !dm.ecell_off) {
				// the counting sort reports what the fixed-capacity buckets WOULD have overflowed, so that a stream of batches that
				// turns uniform again goes back to them
				const int64_t bchunk = std::min<int64_t>(Q, std::max<int64_t>((int64_t)1 << 22, (int64_t)dm.L * 32));
				for (int k = 0; k < ns; k++) { sw[k].stats = m->d_stats + k; sw[k].stats_cap = (int32_t)((double)bucket_capacity(dm, bchunk) * (double)chunk / (double)bchunk); }
			}
			if (ns > 1) {
				// fork: the auxiliary streams start after everything already queued on the caller's stream
				CUDA_OK(m, cudaEventRecord(m->ev_fork, st));
				for (int k = 0; k < ns; k++) { CUDA_OK(m, cudaStreamWaitEvent(m->aux[k], m->ev_fork, 0)); run_on[k] = m->aux[k]; }
			}
			for (int k = 0; k < ns; k++) launch_sorted_reset(dm, sw[k], run_on[k]);
			int k = 0;
			for (int64_t o = 0; o < Q; o += chunk, k = (k + 1) % ns) {
				const int64_t n = std::min(chunk, Q - o);
				m->launches += launch_eval_scalar_sorted_chunk(dm, q + o * dm.N, n, out + o, leaf + o, status + o, sw[k], m->sm_count, m->opt_fp32 != 0, run_on[k], deriv_dir);
			}
			if (ns > 1) {
				// join: the caller's stream continues when all of them are done
				for (int j = 0; j < ns; j++) { CUDA_OK(m, cudaEventRecord(m->ev_join[j], m->aux[j])); CUDA_OK(m, cudaStreamWaitEvent(st, m->ev_join[j], 0)); }
			}
		} else if (tiles) {
			int rc;
			if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
			if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
			if ((rc = grow(m, &w.misc, &w.misc_b, sort_work_bytes(dm, Q)))) return rc;
			m->launches += launch_eval_scalar_tiles(dm, q, Q, out, leaf, status, carve_sort_work(dm, Q, w.misc), m->sm_count, st);
		} else if (dm.tensor_F > 0 && m->opt_path != 3) {
			m->launches += launch_eval_scalar_tensor_direct(dm, in_cube, q, Q, out, leaf, status, st);
		} else {
			m->launches += in_cube ? launch_eval_scalar_in_cube(dm, q, Q, out, leaf, status, st) : launch_eval_scalar_direct(dm, q, Q, out, leaf, status, st);
		}
		CUDA_OK(m, cudaGetLastError());
		return LOCO_OK;
	}
	// mesh-valued on tensor-product leaves: sort by leaf, then the tiled gather-sum kernel (loco_mesh.cu)
	if (dm.tensor_F > 0 && !in_cube && m->opt_path != 1 && m->opt_path != 3) {
		int rc;
		if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
		if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
		if ((rc = grow(m, &w.misc, &w.misc_b, sort_work_bytes(dm, Q)))) return rc;
		const SortWork sw = carve_sort_work(dm, Q, w.misc);
		if (mesh_mma_supported(dm) && !m->opt_mesh_simt) {
			m->launches += launch_sort_by_leaf(dm, q, Q, leaf, status, sw, mesh_mma_tile_queries(), nullptr, st);
			m->launches += launch_mesh_tensor_mma(dm, sw, Q, out, out_stride, m->sm_count, 0, -1, -1, st);
			m->launches += launch_mesh_nan_rows(dm, status, Q, out, out_stride, st);
		} else {
			m->launches += launch_sort_by_leaf(dm, q, Q, leaf, status, sw, mesh_tile_queries(), nullptr, st);
			m->launches += launch_mesh_tensor_tiles(dm, sw, status, Q, out, out_stride, m->sm_count, st);
		}
		CUDA_OK(m, cudaGetLastError());
		return LOCO_OK;
	}
	// mesh-valued on generic leaves (adaptive trees): sort by leaf, weights into a leaf-major matrix, fp64 MMA tiles (loco_mesh_gen.cu)
	if (!in_cube && m->opt_path != 1 && m->opt_path != 3 && !m->opt_mesh_simt && dm.max_support > 0 &&
	    mesh_generic_weight_doubles(dm, Q) * sizeof(double) <= ((size_t)24 << 30)) {
		int rc;
		if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
		if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
		if ((rc = grow(m, &w.misc, &w.misc_b, sort_work_bytes(dm, Q)))) return rc;
		if ((rc = grow(m, &w.W, &w.W_b, mesh_generic_weight_doubles(dm, Q) * sizeof(double)))) return rc;
		if ((rc = grow(m, &w.woff, &w.woff_b, ((size_t)dm.L + 1) * sizeof(int64_t)))) return rc;
		const SortWork sw = carve_sort_work(dm, Q, w.misc);
		m->launches += launch_sort_by_leaf(dm, q, Q, leaf, status, sw, mesh_generic_tile_queries(), nullptr, st);
		m->launches += launch_mesh_generic_mma(dm, sw, leaf, Q, (int64_t *)w.woff, (double *)w.W, out, out_stride, m->sm_count, 0, -1, -1, false, st);
		m->launches += launch_mesh_nan_rows(dm, status, Q, out, out_stride, st);
		CUDA_OK(m, cudaGetLastError());
		return LOCO_OK;
	}
	// generic mesh-valued path: weights, then the weighted gather-sum, in chunks that bound the weight scratch
	const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(Q, ((int64_t)256 << 20) / ((int64_t)dm.max_support * 8)));
	int rc;
	if ((rc = grow(m, &w.W, &w.W_b, (size_t)chunk * dm.max_support * sizeof(double)))) return rc;
	if (!leaf && !in_cube) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
	if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
	for (int64_t o = 0; o < Q; o += chunk) {
		const int64_t n = std::min(chunk, Q - o);
		m->launches += in_cube ? launch_weights_in_cube(dm, q + o * dm.N, n, (double *)w.W, leaf + o, status + o, st)
		                       : launch_weights(dm, q + o * dm.N, n, (double *)w.W, leaf + o, status + o, st);
		m->launches += launch_mesh_gather(dm, (const double *)w.W, leaf + o, status + o, n, out + o * out_stride, out_stride, st);
	}
	CUDA_OK(m, cudaGetLastError());
	return LOCO_OK;
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------------------------
// Three slots overlap H2D, kernels and D2H.  Page-locked caller buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory)
// are used in place.  PAGEABLE caller buffers (std::vector, numpy: what a C++ caller of the facade hands in) would make every
// cudaMemcpyAsync a synchronous, driver-staged copy; they are staged through per-slot page-locked buffers of the handle
// instead, filled and drained by a few host threads, so that the copy engines still run asynchronously beside the kernels.
bool is_page_locked(const void *p)
{
	if (!p) return true;
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
	const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
	const size_t nt = std::min<size_t>(std::min<unsigned>(8, std::max(1u, hw / 2)), bytes / ((size_t)4 << 20) + 1);
	if (nt <= 1) { memcpy(dst, src, bytes); return; }
	std::vector<std::thread> th;
	const size_t part = ((bytes / nt) + 63) & ~(size_t)63;
	for (size_t t = 1; t < nt; t++) {
		const size_t o = t * part;
		if (o >= bytes) break;
		th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, std::min(part, bytes - o)); });
	}
	memcpy(dst, src, std::min(part, bytes));
	for (auto &t : th) t.join();
}

int grow_pinned(loco_model *m, void **p, size_t *have, size_t need)
{
	if (need <= *have) return LOCO_OK;
	if (*p) { CUDA_OK(m, cudaFreeHost(*p)); *p = nullptr; *have = 0; }
	CUDA_OK(m, cudaHostAlloc(p, need, cudaHostAllocDefault));
	*have = need;
	return LOCO_OK;
}

int eval_batch_host(loco_model *m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, int deriv_dir)
{
	const DeviceModel &dm = m->dm;
	const int64_t Dout = deriv_dir >= 0 ? 1 : dm.D;
	// a slot's output is 128 MB for scalar functions; mesh-valued: 1 GB of results per slot, so that a chunk still holds
	// enough queries per leaf for the tile kernels
	const int64_t slot_bytes = Dout == 1 ? ((int64_t)128 << 20) : ((int64_t)1 << 30);
	int64_t chunk = std::max<int64_t>(1, std::min<int64_t>((int64_t)4 << 20, slot_bytes / (Dout * 8)));
	chunk = std::min(chunk, std::max<int64_t>(Q, 1));
	const bool stage_in = !is_page_locked(q);
	const bool stage_out = !is_page_locked(out) || !is_page_locked(leaf) || !is_page_locked(status);
	struct Pending { int64_t o = 0, n = 0; } pending[kSlots];
	auto drain = [&](int slot) {  // results of the slot's previous chunk: page-locked staging -> the caller's pageable buffers
		const Pending &p = pending[slot];
		if (!stage_out || p.n == 0) return;
		Workspace &w = m->ws[slot];
		parallel_memcpy(out + p.o * Dout, w.h_out, (size_t)p.n * Dout * 8);
		if (leaf) memcpy(leaf + p.o, w.h_leaf, (size_t)p.n * 4);
		if (status) memcpy(status + p.o, w.h_status, (size_t)p.n);
		pending[slot].n = 0;
	};
	int rc;
	int slot = 0;
	for (int64_t o = 0; o < Q; o += chunk, slot = (slot + 1) % kSlots) {
		const int64_t n = std::min(chunk, Q - o);
		Workspace &w = m->ws[slot];
		CUDA_OK(m, cudaStreamSynchronize(w.stream));  // the slot's previous chunk must have drained before its buffers are reused
		drain(slot);
		if ((rc = grow(m, &w.q, &w.q_b, (size_t)chunk * dm.N * 8))) return rc;
		if ((rc = grow(m, &w.out, &w.out_b, (size_t)chunk * Dout * 8))) return rc;
		if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)chunk * 4))) return rc;
		if ((rc = grow(m, &w.status, &w.status_b, (size_t)chunk))) return rc;
		const double *src_q = q + o * dm.N;
		if (stage_in) {
			if ((rc = grow_pinned(m, &w.h_q, &w.h_q_b, (size_t)chunk * dm.N * 8))) return rc;
			parallel_memcpy(w.h_q, src_q, (size_t)n * dm.N * 8);
			src_q = (const double *)w.h_q;
		}
		CUDA_OK(m, cudaMemcpyAsync(w.q, src_q, (size_t)n * dm.N * 8, cudaMemcpyHostToDevice, w.stream));
		if ((rc = eval_device(m, w, w.stream, false, (const double *)w.q, n, (double *)w.out, Dout, (int32_t *)w.leaf, (uint8_t *)w.status, deriv_dir))) return rc;
		if (stage_out) {
			if ((rc = grow_pinned(m, &w.h_out, &w.h_out_b, (size_t)chunk * Dout * 8))) return rc;
			if ((rc = grow_pinned(m, &w.h_leaf, &w.h_leaf_b, (size_t)chunk * 4))) return rc;
			if ((rc = grow_pinned(m, &w.h_status, &w.h_status_b, (size_t)chunk))) return rc;
			CUDA_OK(m, cudaMemcpyAsync(w.h_out, w.out, (size_t)n * Dout * 8, cudaMemcpyDeviceToHost, w.stream));
			if (leaf) CUDA_OK(m, cudaMemcpyAsync(w.h_leaf, w.leaf, (size_t)n * 4, cudaMemcpyDeviceToHost, w.stream));
			if (status) CUDA_OK(m, cudaMemcpyAsync(w.h_status, w.status, (size_t)n, cudaMemcpyDeviceToHost, w.stream));
			pending[slot].o = o; pending[slot].n = n;
		} else {
			CUDA_OK(m, cudaMemcpyAsync(out + o * Dout, w.out, (size_t)n * Dout * 8, cudaMemcpyDeviceToHost, w.stream));
			if (leaf) CUDA_OK(m, cudaMemcpyAsync(leaf + o, w.leaf, (size_t)n * 4, cudaMemcpyDeviceToHost, w.stream));
			if (status) CUDA_OK(m, cudaMemcpyAsync(status + o, w.status, (size_t)n, cudaMemcpyDeviceToHost, w.stream));
		}
	}
	for (int k = 0; k < kSlots; k++, slot = (slot + 1) % kSlots) {  // in submission order
		CUDA_OK(m, cudaStreamSynchronize(m->ws[slot].stream));
		drain(slot);
	}
	return LOCO_OK;
}

} // namespace

// =================================================================================================
extern "C" {

int loco_model_from_proto(const char *path, int ascii, int device, loco_model_t **out)
{
	if (!path || !out) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	*out = nullptr;
	loco_model *m = new loco_model();
	Status s = read_proto_file(path, ascii != 0, &m->graph);
	if (!s.ok) {
		delete m;
		bool io = s.msg.rfind("could not open", 0) == 0 || s.msg == "could not parse the file";
		return fail(nullptr, io ? LOCO_ERR_IO : LOCO_ERR_INVALID, s.msg);
	}
	return finish_model(m, device, out);
}

// plain-array graph description -> Graph (shared by loco_model_from_graph and loco_model_update_from_graph)
static int graph_from_desc(const loco_graph_t *gd, Graph &g)
{
	const int N = gd->n_params, D = gd->value_dim;
	if (N < 1 || D < 1) return fail(nullptr, LOCO_ERR_INVALID, "invalid number of parameters");
	if (!gd->domain || !gd->sample_center || !gd->sample_value || !gd->bf_offset || !gd->cell_ranges || !gd->cell_split_dir ||
	    !gd->cell_split_val || !gd->cell_child2 || !gd->cell_sample_offset)
		return fail(nullptr, LOCO_ERR_INVALID, "null array in graph description");
	g.N = N; g.D = D; g.basis = gd->basis_type;
	g.domain.assign(gd->domain, gd->domain + 2 * N);
	const int64_t S = gd->n_samples;
	g.sample_center.assign(gd->sample_center, gd->sample_center + S * N);
	// value rows: one per value object (samples of one group share a row)
	std::unordered_map<int32_t, int32_t> group_row;
	g.sample_row.resize((size_t)S);
	for (int64_t s = 0; s < S; s++) {
		if (gd->sample_value_group) {
			auto it = group_row.find(gd->sample_value_group[s]);
			if (it != group_row.end()) { g.sample_row[(size_t)s] = it->second; continue; }
			group_row[gd->sample_value_group[s]] = (int32_t)g.n_rows();
		}
		g.sample_row[(size_t)s] = (int32_t)g.n_rows();
		g.values.insert(g.values.end(), gd->sample_value + s * D, gd->sample_value + (s + 1) * D);
	}
	g.bf_off.assign(gd->bf_offset, gd->bf_offset + S + 1);
	const int64_t B = S > 0 ? gd->bf_offset[S] : 0;
	if (B > 0 && (!gd->bf_center || !gd->bf_support || !gd->bf_weight)) return fail(nullptr, LOCO_ERR_INVALID, "null basis function arrays");
	g.bf_center.assign(gd->bf_center, gd->bf_center + B * N);
	g.bf_support.assign(gd->bf_support, gd->bf_support + B * N);
	g.bf_weight.assign(gd->bf_weight, gd->bf_weight + B);
	auto fill_cells = [&](Graph::CellSet &cs, int64_t C, const double *ranges, const int32_t *split_dir, const double *split_val,
	                      const int32_t *child2, const double *center_value, const int64_t *off, const int32_t *sample, const double *copy) -> bool {
		cs.ranges.assign(ranges, ranges + C * 2 * N);
		cs.split_dir.resize((size_t)C); cs.split_val.resize((size_t)C); cs.child2.resize((size_t)C); cs.center_row.assign((size_t)C, -1);
		for (int64_t c = 0; c < C; c++) {
			cs.split_dir[(size_t)c] = split_dir ? split_dir[c] : -1;
			cs.split_val[(size_t)c] = split_val ? split_val[c] : 0.0;
			cs.child2[(size_t)c] = child2 ? child2[c] : -1;
			if (center_value && cs.split_dir[(size_t)c] < 0) {
				cs.center_row[(size_t)c] = (int32_t)g.n_rows();
				g.values.insert(g.values.end(), center_value + c * D, center_value + (c + 1) * D);
			}
		}
		cs.ls_off.assign(off, off + C + 1);
		const int64_t E = off[C];
		cs.ls_sample.assign(sample, sample + E);
		cs.ls_row.resize((size_t)E);
		for (int64_t e = 0; e < E; e++) {
			int32_t s = sample[e];
			if (s < 0 || s >= S) return false;
			int32_t sr = g.sample_row[(size_t)s];
			if (!copy || memcmp(copy + e * D, &g.values[(size_t)sr * D], sizeof(double) * D) == 0) cs.ls_row[(size_t)e] = sr;
			else {
				cs.ls_row[(size_t)e] = (int32_t)g.n_rows();
				g.values.insert(g.values.end(), copy + e * D, copy + (e + 1) * D);
			}
		}
		return true;
	};
	bool ok = fill_cells(g.tree, gd->n_cells, gd->cell_ranges, gd->cell_split_dir, gd->cell_split_val, gd->cell_child2, gd->cell_center_value,
	                     gd->cell_sample_offset, gd->cell_sample, gd->cell_sample_value);
	if (ok && gd->n_border_cells > 0) {
		if (!gd->border_ranges || !gd->border_sample_offset) ok = false;
		else ok = fill_cells(g.border, gd->n_border_cells, gd->border_ranges, nullptr, nullptr, nullptr, gd->border_center_value,
		                     gd->border_sample_offset, gd->border_sample, gd->border_sample_value);
	} else if (ok) g.border.ls_off.assign(1, 0);
	if (!ok) return fail(nullptr, LOCO_ERR_INVALID, "could not map the sample from id");
	return LOCO_OK;
}

int loco_model_from_graph(const loco_graph_t *gd, int device, loco_model_t **out)
{
	if (!gd || !out) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	*out = nullptr;
	loco_model *m = new loco_model();
	if (int rc = graph_from_desc(gd, m->graph)) { delete m; return rc; }
	return finish_model(m, device, out);
}

int loco_model_update_from_graph(loco_model_t *m, const loco_graph_t *gd)
{
	if (!m || !gd) return fail(m, LOCO_ERR_INVALID, "null argument");
	Graph ng;
	if (int rc = graph_from_desc(gd, ng)) { m->err = g_ctor_error; return rc; }
	if (ng.N != m->flat.N || ng.D != m->flat.D || ng.basis != m->flat.basis)
		return fail(m, LOCO_ERR_INVALID, "the new state must be the same function (parameters, value dimension, basis type)");
	Flat nf;
	Status s = flatten(ng, &nf, &m->flat);
	if (!s.ok) return fail(m, LOCO_ERR_INVALID, s.msg);
	if (!m->host_only) {
		// the flat arrays move (CSR offsets shift with every new leaf), so the device image is rebuilt; streams, events,
		// workspaces and the skew counters of the handle stay
		CUDA_OK(m, cudaSetDevice(m->device));
		CUDA_OK(m, cudaDeviceSynchronize());
		for (void *p : m->allocs) cudaFree(p);
		m->allocs.clear();
		m->device_bytes = 0;
		m->d_values = nullptr; m->d_term_wv = nullptr; m->d_mterm_wv = nullptr; m->d_eterm_wv = nullptr; m->d_ecell_off = nullptr;
		m->d_epoly_idx = nullptr; m->d_epoly_cell = nullptr; m->d_epoly = nullptr; m->n_epoly = 0; m->d_tl_block = nullptr; m->d_tl_poly = nullptr;
		m->d_tl_vals_f32 = nullptr; m->d_lut = nullptr;
		m->dm = DeviceModel{};
	}
	m->graph = std::move(ng);
	m->flat = std::move(nf);
	return build_device_model(m);
}

int loco_model_bootstrap(int32_t n_params, int32_t basis_type, int32_t level, int32_t value_dim, int32_t eval_corners, const double *domain,
                         loco_value_fn fn, void *user, int device, loco_model_t **out)
{
	if (!domain || !out) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	*out = nullptr;
	loco_model *m = new loco_model();
	Status s = bootstrap_graph(n_params, basis_type, level, value_dim, eval_corners != 0, domain, fn, user, &m->graph);
	if (!s.ok) { delete m; return fail(nullptr, LOCO_ERR_INVALID, s.msg); }
	return finish_model(m, device, out);
}

int loco_model_save_proto(const loco_model_t *m, const char *path, int ascii)
{
	if (!m || !path) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (int rc = sync_values_to_host(m)) return rc;  // values replaced on the device since construction: serialise what the kernels evaluate
	if (m->graph.values.empty()) return fail(m, LOCO_ERR_INVALID, "undefined type of shape info");  // no value table at all (host-only handle built without values)
	Status s = write_proto_file(m->graph, path, ascii != 0);
	if (!s.ok) return fail(m, s.msg.rfind("could not write", 0) == 0 ? LOCO_ERR_IO : LOCO_ERR_INVALID, s.msg);
	return LOCO_OK;
}

void loco_free(loco_model_t *m)
{
	if (!m) return;
	if (!m->host_only && (!m->allocs.empty() || m->ws[0].stream)) cudaSetDevice(m->device);
	for (void *p : m->allocs) cudaFree(p);
	for (int k = 0; k < 4; k++) {
		if (m->aux[k]) cudaStreamDestroy(m->aux[k]);
		if (m->ev_join[k]) cudaEventDestroy(m->ev_join[k]);
	}
	if (m->ev_fork) cudaEventDestroy(m->ev_fork);
	if (m->h_stats) cudaFreeHost(m->h_stats);
	for (int s = 0; s < kSlots; s++) {
		Workspace &w = m->ws[s];
		for (void *p : {w.q, w.out, w.leaf, w.status, w.W, w.truth, w.misc, w.sorted, w.sortedx[0], w.sortedx[1], w.sortedx[2], w.woff}) if (p) cudaFree(p);
		for (void *p : {w.h_q, w.h_out, w.h_leaf, w.h_status}) if (p) cudaFreeHost(p);
		if (w.stream) cudaStreamDestroy(w.stream);
	}
	delete m;
}

int loco_model_info(const loco_model_t *m, loco_info_t *info)
{
	if (!m || !info) return fail(m, LOCO_ERR_INVALID, "null argument");
	const Flat &f = m->flat;
	memset(info, 0, sizeof(*info));
	info->n_params = f.N; info->value_dim = f.D; info->basis_type = f.basis; info->exact_rcp = f.exact_rcp;
	info->depth = f.depth; info->max_support = f.max_support; info->max_terms = f.max_terms; info->device = m->device;
	info->n_samples = f.S; info->n_value_rows = f.R; info->n_leaves = f.L; info->n_cells = f.C;
	info->n_border_cells = m->graph.border.size();
	info->n_basis_functions = m->graph.bf_off.empty() ? 0 : m->graph.bf_off.back();
	info->n_support_entries = f.sup_off.back();
	info->n_terms = f.term_off.back();
	info->device_bytes = m->device_bytes;
	info->n_fold_entries = f.fold_off.empty() ? 0 : f.fold_off.back();
	info->n_eval_cells = f.ecell_off.empty() ? 0 : f.ecell_off.back();
	info->n_eval_terms = f.eterm_off.empty() ? 0 : f.eterm_off.back();
	info->n_poly_cells = (int64_t)f.epoly_cell.size();
	info->n_reused_leaves = f.reused_leaves;
	return LOCO_OK;
}

int loco_support_set(const loco_model_t *m, int32_t leaf, int32_t *ids, int32_t cap, int32_t *k)
{
	if (!m || !k) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (leaf < 0 || leaf >= m->flat.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	const int64_t a = m->flat.sup_off[(size_t)leaf], b = m->flat.sup_off[(size_t)leaf + 1];
	*k = (int32_t)(b - a);
	if (!ids) return LOCO_OK;
	if (cap < b - a) return fail(m, LOCO_ERR_RANGE, "support buffer too small");
	std::copy(m->flat.sup_sample.begin() + a, m->flat.sup_sample.begin() + b, ids);
	return LOCO_OK;
}

int loco_merged_terms(const loco_model_t *m, int32_t leaf, double *centre, double *support, int32_t cap, int32_t *n, int64_t *first_id)
{
	if (!m || !n) return fail(m, LOCO_ERR_INVALID, "null argument");
	const Flat &f = m->flat;
	if (leaf < 0 || leaf >= f.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	if (f.mterm_off.empty()) { *n = 0; if (first_id) *first_id = 0; return LOCO_OK; }
	const int64_t a = f.mterm_off[(size_t)leaf], b = f.mterm_off[(size_t)leaf + 1];
	*n = (int32_t)(b - a);
	if (first_id) *first_id = a;
	if (!centre && !support) return LOCO_OK;
	if ((int64_t)f.mterm_cs.size() < b * 2 * f.N) return fail(m, LOCO_ERR_INVALID, "term lists were released after upload (use a host-only handle)");
	if (cap < b - a) return fail(m, LOCO_ERR_RANGE, "term buffer too small");
	for (int64_t t = a; t < b; t++)
		for (int i = 0; i < f.N; i++) {
			const double c = f.mterm_cs[(size_t)t * 2 * f.N + 2 * i], v = f.mterm_cs[(size_t)t * 2 * f.N + 2 * i + 1];
			if (centre) centre[(t - a) * f.N + i] = c;
			if (support) support[(t - a) * f.N + i] = f.exact_rcp ? 1.0 / v : v;  // stored as the reciprocal when that is exact
		}
	return LOCO_OK;
}

int loco_eval_cells(const loco_model_t *m, int32_t leaf, int32_t *bits, int64_t *first_cell, int64_t *n_cells)
{
	if (!m) return fail(m, LOCO_ERR_INVALID, "null argument");
	const Flat &f = m->flat;
	if (leaf < 0 || leaf >= f.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	const bool have = !f.ecell_off.empty();
	if (bits) for (int i = 0; i < f.N; i++) bits[i] = have ? f.ecell_bits[(size_t)leaf * f.N + i] : 0;
	if (first_cell) *first_cell = have ? f.ecell_off[(size_t)leaf] : leaf;
	if (n_cells) *n_cells = have ? f.ecell_off[(size_t)leaf + 1] - f.ecell_off[(size_t)leaf] : 1;
	return LOCO_OK;
}

int loco_eval_cell_terms(const loco_model_t *m, int64_t cell, int64_t *term_ids, int32_t cap, int32_t *n)
{
	if (!m || !n) return fail(m, LOCO_ERR_INVALID, "null argument");
	const Flat &f = m->flat;
	if (f.ecell_off.empty()) return fail(m, LOCO_ERR_INVALID, "the model has no evaluation cells");
	if (cell < 0 || cell >= f.ecell_off.back()) return fail(m, LOCO_ERR_RANGE, "evaluation cell out of range");
	const int64_t a = f.eterm_off[(size_t)cell], b = f.eterm_off[(size_t)cell + 1];
	*n = (int32_t)(b - a);
	if (!term_ids) return LOCO_OK;
	if ((int64_t)f.eterm_src.size() < b) return fail(m, LOCO_ERR_INVALID, "term lists were released after upload (use a host-only handle)");
	if (cap < b - a) return fail(m, LOCO_ERR_RANGE, "term buffer too small");
	for (int64_t t = a; t < b; t++) term_ids[t - a] = f.eterm_src[(size_t)t];
	return LOCO_OK;
}

int loco_eval_cell_poly(const loco_model_t *m, int64_t cell, double *geom, double *coef, int32_t cap, int32_t *n_coef)
{
	if (!m || !n_coef) return fail(m, LOCO_ERR_INVALID, "null argument");
	const Flat &f = m->flat;
	if (f.ecell_off.empty()) return fail(m, LOCO_ERR_INVALID, "the model has no evaluation cells");
	if (cell < 0 || cell >= f.ecell_off.back()) return fail(m, LOCO_ERR_RANGE, "evaluation cell out of range");
	*n_coef = 0;
	if (f.epoly_idx.empty() || f.epoly_idx[(size_t)cell] < 0) return LOCO_OK;
	const int N = f.N;
	const bool cubic = m->graph.basis == CUBIC_BSPLINE;
	const int F = cubic ? 4 : 2;
	int P = 1;
	for (int i = 0; i < N; i++) P *= F;
	*n_coef = P;
	if (!geom && !coef) return LOCO_OK;
	if (cap < P) return fail(m, LOCO_ERR_RANGE, "coefficient buffer too small");
	const int64_t a = f.eterm_off[(size_t)cell], b = f.eterm_off[(size_t)cell + 1];
	if (int rc = sync_values_to_host(m)) return rc;
	if ((int64_t)f.eterm_src.size() < b || f.mrun_term.empty() || f.term_w.empty() || m->graph.values.empty())
		return fail(m, LOCO_ERR_INVALID, "term lists were released after upload (use a host-only handle)");
	const double *gm = &f.epoly_geom[(size_t)f.epoly_idx[(size_t)cell] * 2 * N];
	if (geom) std::copy(gm, gm + 2 * N, geom);
	if (!coef) return LOCO_OK;
	std::fill(coef, coef + P, 0.0);
	double p[8 * 4];
	for (int64_t t = a; t < b; t++) {
		const int64_t mt = f.eterm_src[(size_t)t];
		double wv = 0.0;  // as mterm_wv_kernel: weight * value summed over the merged run
		for (int64_t e = f.mrun_off[(size_t)mt]; e < f.mrun_off[(size_t)mt + 1]; e++) {
			const int32_t o = f.mrun_term[(size_t)e];
			wv += f.term_w[(size_t)o] * m->graph.values[(size_t)f.term_row[(size_t)o] * m->graph.D];
		}
		for (int i = 0; i < N; i++) {
			const double c = f.eterm_cs[(size_t)t * 2 * N + 2 * i], v = f.eterm_cs[(size_t)t * 2 * N + 2 * i + 1];
			const double sp = f.exact_rcp ? 1.0 / v : v;
			if (cubic) cell_piece_1d<true>(c, sp, gm[2 * i], 1.0 / gm[2 * i + 1], p + i * F);
			else cell_piece_1d<false>(c, sp, gm[2 * i], 1.0 / gm[2 * i + 1], p + i * F);
		}
		if (cubic) cell_poly_accumulate<4>(N, p, wv, coef); else cell_poly_accumulate<2>(N, p, wv, coef);
	}
	return LOCO_OK;
}

int loco_leaf_neighbors(const loco_model_t *m, int32_t leaf, int32_t *n_leaf, int32_t *n_border)
{
	if (!m) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (leaf < 0 || leaf >= m->flat.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	if (n_leaf) *n_leaf = m->flat.n_leaf_neighbors[(size_t)leaf];
	if (n_border) *n_border = m->flat.n_border_neighbors[(size_t)leaf];
	return LOCO_OK;
}

int loco_leaf_parent_split(const loco_model_t *m, int32_t leaf, int32_t *direction)
{
	if (!m || !direction) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (leaf < 0 || leaf >= m->flat.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	// tree cells are in pre-order with child1 first and leaves are numbered in that order: walk down from the root
	const Graph::CellSet &t = m->graph.tree;
	const int64_t C = t.size();
	int32_t rank = 0, found = -1;
	std::vector<int32_t> parent_dir((size_t)C, -1);
	for (int64_t c = 0; c < C && found == -1; c++) {
		if (t.split_dir[(size_t)c] >= 0) {
			if (c + 1 < C) parent_dir[(size_t)c + 1] = t.split_dir[(size_t)c];
			const int32_t c2 = t.child2[(size_t)c];
			if (c2 >= 0 && c2 < C) parent_dir[(size_t)c2] = t.split_dir[(size_t)c];
		} else if (rank++ == leaf) found = parent_dir[(size_t)c];
	}
	*direction = found;
	return LOCO_OK;
}

int loco_sample_value_rows(const loco_model_t *m, int32_t *rows)
{
	if (!m || !rows) return fail(m, LOCO_ERR_INVALID, "null argument");
	std::copy(m->graph.sample_row.begin(), m->graph.sample_row.end(), rows);
	return LOCO_OK;
}

int loco_model_domain(const loco_model_t *m, double *ranges)
{
	if (!m || !ranges) return fail(m, LOCO_ERR_INVALID, "null argument");
	for (int i = 0; i < m->flat.N; i++) { ranges[2 * i] = m->flat.dom_min[(size_t)i]; ranges[2 * i + 1] = m->flat.dom_max[(size_t)i]; }
	return LOCO_OK;
}

int loco_leaf_box(const loco_model_t *m, int32_t leaf, double *box)
{
	if (!m || !box) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (leaf < 0 || leaf >= m->flat.L) return fail(m, LOCO_ERR_RANGE, "leaf index out of range");
	const int N = m->flat.N;
	for (int i = 0; i < N; i++) { box[2 * i] = m->flat.leaf_lo[(size_t)leaf * N + i]; box[2 * i + 1] = m->flat.leaf_hi[(size_t)leaf * N + i]; }
	return LOCO_OK;
}

const char *loco_last_error(const loco_model_t *m) { return m ? m->err.c_str() : g_ctor_error.c_str(); }
const char *loco_status_string(int status) { return status_string(status); }

// ---- value table -------------------------------------------------------------------------------------
int loco_model_set_values(loco_model_t *m, const double *values)
{
	if (!m || !values) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	CUDA_OK(m, cudaMemcpy2D(m->d_values, (size_t)dm.value_stride * sizeof(double), values, (size_t)dm.D * sizeof(double),
	                        (size_t)dm.D * sizeof(double), (size_t)dm.R, cudaMemcpyHostToDevice));
	// keep the host copy (what save_proto writes) in step with what the kernels evaluate
	const size_t n = (size_t)dm.R * dm.D;
	if (!m->initial_upload) {
		if (n * sizeof(double) <= kHostValueBytes) { m->graph.values.assign(values, values + n); m->graph.virtual_rows = 0; }
		else drop_host_values(m);
	}
	return refresh_term_wv(m);
}

int loco_model_fill_values_synthetic(loco_model_t *m, double seed)
{
	if (!m) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	m->launches += launch_fill_values(m->d_values, m->dm.R, m->dm.D, m->dm.value_stride, seed, nullptr);
	CUDA_OK(m, cudaGetLastError());
	drop_host_values(m);  // the host copy is stale now; save_proto / eval_cell_poly read the table back
	return refresh_term_wv(m);
}

int loco_model_get_values(const loco_model_t *m, int64_t row0, int64_t n, double *out)
{
	if (!m || !out) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (row0 < 0 || n < 0 || row0 + n > m->dm.R) return fail(m, LOCO_ERR_RANGE, "value row out of range");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	CUDA_OK(m, cudaMemcpy2D(out, (size_t)dm.D * sizeof(double), m->d_values + row0 * dm.value_stride, (size_t)dm.value_stride * sizeof(double),
	                        (size_t)dm.D * sizeof(double), (size_t)n, cudaMemcpyDeviceToHost));
	return LOCO_OK;
}

int loco_model_get_value_columns(const loco_model_t *m, int64_t col0, int64_t n, double *out)
{
	if (!m || !out) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (col0 < 0 || n < 0 || col0 + n > m->dm.D) return fail(m, LOCO_ERR_RANGE, "value component out of range");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	if (n == 0) return LOCO_OK;
	CUDA_OK(m, cudaMemcpy2D(out, (size_t)n * sizeof(double), m->d_values + col0, (size_t)dm.value_stride * sizeof(double),
	                        (size_t)n * sizeof(double), (size_t)dm.R, cudaMemcpyDeviceToHost));
	return LOCO_OK;
}

// ---- hot path ---------------------------------------------------------------------------------------------
int loco_eval_batch_device(loco_model_t *m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status, void *cuda_stream)
{
	if (!m || (Q > 0 && (!q || !out))) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	return eval_device(m, m->ws[0], (cudaStream_t)cuda_stream, false, q, Q, out, m->dm.D, leaf, status);
}

int loco_eval_batch(loco_model_t *m, const double *q, int64_t Q, double *out, int32_t *leaf, uint8_t *status)
{
	if (!m || (Q > 0 && (!q || !out))) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	return eval_batch_host(m, q, Q, out, leaf, status, -1);
}

int loco_eval_batch_ring_device(loco_model_t *m, const double *q, int64_t Q, double *ring, int64_t ring_queries, double *checksum,
                                int32_t *leaf, uint8_t *status, void *cuda_stream)
{
	if (!m || (Q > 0 && (!q || !ring || !checksum)) || ring_queries < 1) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	int rc;
	if (dm.D > 1 && mesh_mma_supported(dm) && !m->opt_mesh_simt && m->opt_path != 1 && m->opt_path != 3 && Q > 0 && ring_queries >= mesh_mma_tile_queries()) {
		// Sort the WHOLE batch by leaf once, then walk the sorted tiles in ring-sized runs: every run is made of full
		// 64-query tiles of one leaf each (however few queries a ring holds), results sit in the ring in leaf order and the
		// consumer (here: the checksum) maps them back to query indices through the sorted records.
		Workspace &w = m->ws[0];
		if (!leaf) { if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)Q * sizeof(int32_t)))) return rc; leaf = (int32_t *)w.leaf; }
		if (!status) { if ((rc = grow(m, &w.status, &w.status_b, (size_t)Q))) return rc; status = (uint8_t *)w.status; }
		if ((rc = grow(m, &w.misc, &w.misc_b, sort_work_bytes(dm, Q)))) return rc;
		const SortWork sw = carve_sort_work(dm, Q, w.misc);
		const int TQ = mesh_mma_tile_queries();
		m->launches += launch_sort_by_leaf(dm, q, Q, leaf, status, sw, TQ, nullptr, st);
		// bucket and tile offsets to the host (2 * (L+1) ints): the split into runs is host logic
		std::vector<int32_t> h_off((size_t)dm.L + 1), h_tile((size_t)dm.L + 1);
		CUDA_OK(m, cudaMemcpyAsync(h_off.data(), sw.offset, h_off.size() * 4, cudaMemcpyDeviceToHost, st));
		CUDA_OK(m, cudaMemcpyAsync(h_tile.data(), sw.tile_off, h_tile.size() * 4, cudaMemcpyDeviceToHost, st));
		CUDA_OK(m, cudaStreamSynchronize(st));
		const int64_t n_tiles = h_tile[(size_t)dm.L];
		// first sorted position of every tile boundary we may cut at (positions are contiguous in tile order)
		auto tile_pos = [&](int64_t t) -> int64_t {
			if (t >= n_tiles) return h_off[(size_t)dm.L];
			const int32_t l = (int32_t)(std::upper_bound(h_tile.begin(), h_tile.end(), (int32_t)t) - h_tile.begin()) - 1;
			return (int64_t)h_off[(size_t)l] + (t - h_tile[(size_t)l]) * TQ;
		};
		for (int64_t t0 = 0; t0 < n_tiles;) {
			const int64_t p0 = tile_pos(t0);
			int64_t lo = t0 + 1, hi = n_tiles;  // largest t1 with tile_pos(t1) - p0 <= ring_queries
			while (lo < hi) {
				const int64_t mid = (lo + hi + 1) / 2;
				if (tile_pos(mid) - p0 <= ring_queries) lo = mid; else hi = mid - 1;
			}
			const int64_t t1 = lo, n = tile_pos(t1) - p0;
			m->launches += launch_mesh_tensor_mma(dm, sw, n, ring, dm.value_stride, m->sm_count, t0, t1 - t0, p0, st);
			m->launches += launch_checksum_sorted(dm, ring, dm.value_stride, sw, p0, n, checksum, st);
			t0 = t1;
		}
		m->launches += launch_checksum_bad(status, Q, checksum, st);
		CUDA_OK(m, cudaGetLastError());
		return LOCO_OK;
	}
	for (int64_t o = 0; o < Q; o += ring_queries) {
		const int64_t n = std::min(ring_queries, Q - o);
		if ((rc = eval_device(m, m->ws[0], st, false, q + o * dm.N, n, ring, dm.value_stride, leaf ? leaf + o : nullptr, status ? status + o : nullptr))) return rc;
		m->launches += launch_checksum(ring, dm.value_stride, dm.D, n, checksum + o, st);
	}
	CUDA_OK(m, cudaGetLastError());
	return LOCO_OK;
}

int loco_eval_batch_ring(loco_model_t *m, const double *q, int64_t Q, int64_t ring_queries, double *checksum, int32_t *leaf, uint8_t *status)
{
	if (!m || (Q > 0 && (!q || !checksum)) || ring_queries < 0) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	if (Q == 0) return LOCO_OK;
	if (ring_queries == 0) ring_queries = std::max<int64_t>(mesh_mma_tile_queries(), std::min<int64_t>(Q, ((int64_t)16 << 30) / ((int64_t)dm.value_stride * 8)));
	Workspace &w = m->ws[1];  // slot 1: slot 0 is the device-side call's workspace
	cudaStream_t st = w.stream;
	int rc;
	if ((rc = grow(m, &w.q, &w.q_b, (size_t)Q * dm.N * 8))) return rc;
	if ((rc = grow(m, &w.out, &w.out_b, (size_t)ring_queries * dm.value_stride * 8))) return rc;
	if ((rc = grow(m, &w.truth, &w.truth_b, (size_t)Q * 8))) return rc;      // the checksums
	if ((rc = grow(m, &w.W, &w.W_b, (size_t)Q * 5))) return rc;              // leaf + status for the caller
	int32_t *d_leaf = (int32_t *)w.W;
	uint8_t *d_status = (uint8_t *)w.W + (size_t)Q * 4;
	CUDA_OK(m, cudaMemcpyAsync(w.q, q, (size_t)Q * dm.N * 8, cudaMemcpyHostToDevice, st));
	if ((rc = loco_eval_batch_ring_device(m, (const double *)w.q, Q, (double *)w.out, ring_queries, (double *)w.truth, d_leaf, d_status, st))) return rc;
	CUDA_OK(m, cudaMemcpyAsync(checksum, w.truth, (size_t)Q * 8, cudaMemcpyDeviceToHost, st));
	if (leaf) CUDA_OK(m, cudaMemcpyAsync(leaf, d_leaf, (size_t)Q * 4, cudaMemcpyDeviceToHost, st));
	if (status) CUDA_OK(m, cudaMemcpyAsync(status, d_status, (size_t)Q, cudaMemcpyDeviceToHost, st));
	CUDA_OK(m, cudaStreamSynchronize(st));
	return LOCO_OK;
}

// ---- derivative ---------------------------------------------------------------------------------------------------
int loco_eval_deriv_batch_device(loco_model_t *m, const double *q, int64_t Q, int32_t direction, double *out, int32_t *leaf, uint8_t *status,
                                 void *cuda_stream)
{
	if (!m || (Q > 0 && (!q || !out))) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (m->flat.D != 1) return fail(m, LOCO_ERR_INVALID, "Shape info type not specified");  // LCFunctionDeriv.cpp:22-34: Real values only
	if (direction < 0 || direction >= m->flat.N) return fail(m, LOCO_ERR_RANGE, "derivative direction out of range");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	return eval_device(m, m->ws[0], (cudaStream_t)cuda_stream, false, q, Q, out, 1, leaf, status, direction);
}

int loco_eval_deriv_batch(loco_model_t *m, const double *q, int64_t Q, int32_t direction, double *out, int32_t *leaf, uint8_t *status)
{
	if (!m || (Q > 0 && (!q || !out))) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (m->flat.D != 1) return fail(m, LOCO_ERR_INVALID, "Shape info type not specified");
	if (direction < 0 || direction >= m->flat.N) return fail(m, LOCO_ERR_RANGE, "derivative direction out of range");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	return eval_batch_host(m, q, Q, out, leaf, status, direction);
}

// ---- error check ----------------------------------------------------------------------------------------------
int loco_error_check_device(loco_model_t *m, const double *q, const double *truth, int64_t P, double threshold, double *err_max, uint8_t *split,
                            int64_t *n_bad_dev, void *cuda_stream)
{
	if (!m || !err_max || !split || (P > 0 && (!q || !truth))) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	Workspace &w = m->ws[0];
	CUDA_OK(m, cudaMemsetAsync(err_max, 0, (size_t)dm.L * 8, st));
	CUDA_OK(m, cudaMemsetAsync(split, 0, (size_t)dm.L, st));
	if (n_bad_dev) CUDA_OK(m, cudaMemsetAsync(n_bad_dev, 0, 8, st));
	const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>((int64_t)8 << 20, ((int64_t)256 << 20) / ((int64_t)dm.D * 8)));
	int rc;
	for (int64_t o = 0; o < P; o += chunk) {
		const int64_t n = std::min(chunk, P - o);
		if ((rc = grow(m, &w.out, &w.out_b, (size_t)chunk * dm.D * 8))) return rc;
		if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)chunk * 4))) return rc;
		if ((rc = grow(m, &w.status, &w.status_b, (size_t)chunk))) return rc;
		if ((rc = eval_device(m, w, st, false, q + o * dm.N, n, (double *)w.out, dm.D, (int32_t *)w.leaf, (uint8_t *)w.status))) return rc;
		m->launches += launch_error_reduce(dm, (const double *)w.out, truth + o * dm.D, (const int32_t *)w.leaf, (const uint8_t *)w.status, n, threshold,
		                                   err_max, split, (unsigned long long *)n_bad_dev, st);
	}
	CUDA_OK(m, cudaGetLastError());
	return LOCO_OK;
}

int loco_error_check(loco_model_t *m, const double *q, const double *truth, int64_t P, double threshold, double *err_max, uint8_t *split, int64_t *n_bad)
{
	if (!m || !err_max || !split || (P > 0 && (!q || !truth))) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	void *dq = nullptr, *dt = nullptr, *de = nullptr, *ds = nullptr, *db = nullptr;
	int rc = LOCO_OK;
	cudaStream_t st = m->ws[1].stream;
	do {
		if (cudaMalloc(&dq, std::max<size_t>((size_t)P * dm.N * 8, 8)) || cudaMalloc(&dt, std::max<size_t>((size_t)P * dm.D * 8, 8)) ||
		    cudaMalloc(&de, (size_t)dm.L * 8) || cudaMalloc(&ds, (size_t)dm.L) || cudaMalloc(&db, 8)) { rc = fail(m, LOCO_ERR_CUDA, "cudaMalloc failed in loco_error_check"); break; }
		cudaMemcpyAsync(dq, q, (size_t)P * dm.N * 8, cudaMemcpyHostToDevice, st);
		cudaMemcpyAsync(dt, truth, (size_t)P * dm.D * 8, cudaMemcpyHostToDevice, st);
		if ((rc = loco_error_check_device(m, (const double *)dq, (const double *)dt, P, threshold, (double *)de, (uint8_t *)ds, (int64_t *)db, st))) break;
		cudaMemcpyAsync(err_max, de, (size_t)dm.L * 8, cudaMemcpyDeviceToHost, st);
		cudaMemcpyAsync(split, ds, (size_t)dm.L, cudaMemcpyDeviceToHost, st);
		if (n_bad) cudaMemcpyAsync(n_bad, db, 8, cudaMemcpyDeviceToHost, st);
		cudaError_t e = cudaStreamSynchronize(st);
		if (e != cudaSuccess) rc = fail(m, LOCO_ERR_CUDA, cudaGetErrorString(e));
	} while (0);
	for (void *p : {dq, dt, de, ds, db}) if (p) cudaFree(p);
	return rc;
}

int loco_center_error(loco_model_t *m, double threshold, double *err_max, uint8_t *split)
{
	if (!m || !err_max || !split) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	for (int64_t l = 0; l < m->flat.L; l++)
		if (m->flat.leaf_center_row[(size_t)l] < 0) return fail(m, LOCO_ERR_INVALID, "model has no centre values (centerShapeInfo_)");
	const int64_t L = dm.L;
	Workspace &w = m->ws[1];
	cudaStream_t st = w.stream;
	int rc;
	if ((rc = grow(m, &w.q, &w.q_b, (size_t)L * dm.N * 8))) return rc;
	if ((rc = grow(m, &w.out, &w.out_b, (size_t)L * dm.D * 8))) return rc;
	if ((rc = grow(m, &w.truth, &w.truth_b, (size_t)L * dm.D * 8))) return rc;
	if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)L * 4))) return rc;
	if ((rc = grow(m, &w.status, &w.status_b, (size_t)L))) return rc;
	if ((rc = grow(m, &w.misc, &w.misc_b, (size_t)L * 9))) return rc;
	std::vector<int32_t> ids((size_t)L);
	for (int64_t l = 0; l < L; l++) ids[(size_t)l] = (int32_t)l;
	CUDA_OK(m, cudaMemcpyAsync(w.leaf, ids.data(), (size_t)L * 4, cudaMemcpyHostToDevice, st));
	CUDA_OK(m, cudaStreamSynchronize(st));
	m->launches += launch_leaf_centers(dm, (double *)w.q, st);
	if ((rc = eval_device(m, w, st, true, (const double *)w.q, L, (double *)w.out, dm.D, (int32_t *)w.leaf, (uint8_t *)w.status))) return rc;
	m->launches += launch_gather_rows(dm, dm.leaf_center_row, L, (double *)w.truth, st);
	double *de = (double *)w.misc;
	uint8_t *ds = (uint8_t *)w.misc + L * 8;
	CUDA_OK(m, cudaMemsetAsync(w.misc, 0, (size_t)L * 9, st));
	m->launches += launch_error_reduce(dm, (const double *)w.out, (const double *)w.truth, (const int32_t *)w.leaf, (const uint8_t *)w.status, L, threshold, de, ds, nullptr, st);
	CUDA_OK(m, cudaGetLastError());
	CUDA_OK(m, cudaMemcpyAsync(err_max, de, (size_t)L * 8, cudaMemcpyDeviceToHost, st));
	CUDA_OK(m, cudaMemcpyAsync(split, ds, (size_t)L, cudaMemcpyDeviceToHost, st));
	CUDA_OK(m, cudaStreamSynchronize(st));
	return LOCO_OK;
}

int loco_error_check_generated_device(loco_model_t *m, int64_t ppl, uint64_t seed, const double *amp, const double *phase, int32_t n_terms,
                                      double threshold, double *err_max, uint8_t *split, void *cuda_stream)
{
	if (!m || !err_max || !split || !amp || !phase || ppl < 1 || n_terms < 0) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (m->dm.D != 1) return fail(m, LOCO_ERR_INVALID, "generated error check needs a scalar function");
	NEED_DEVICE(m);
	CUDA_OK(m, cudaSetDevice(m->device));
	const DeviceModel &dm = m->dm;
	cudaStream_t st = (cudaStream_t)cuda_stream;
	Workspace &w = m->ws[0];
	CUDA_OK(m, cudaMemsetAsync(err_max, 0, (size_t)dm.L * 8, st));
	CUDA_OK(m, cudaMemsetAsync(split, 0, (size_t)dm.L, st));
	if (m->opt_fused_errgen && error_check_generated_fused_supported(dm, n_terms)) {
		// tensor-product leaves: everything in one kernel, no intermediates in HBM (loco_errgen.cu)
		m->launches += launch_error_check_generated_fused(dm, ppl, seed, amp, phase, n_terms, threshold, err_max, split, m->sm_count, st);
		CUDA_OK(m, cudaGetLastError());
		return LOCO_OK;
	}
	const int64_t total = (int64_t)dm.L * ppl;
	const int64_t chunk = std::min<int64_t>(total, (int64_t)16 << 20);
	int rc;
	if ((rc = grow(m, &w.q, &w.q_b, (size_t)chunk * dm.N * 8))) return rc;
	if ((rc = grow(m, &w.out, &w.out_b, (size_t)chunk * 8))) return rc;
	if ((rc = grow(m, &w.truth, &w.truth_b, (size_t)chunk * 8))) return rc;
	if ((rc = grow(m, &w.leaf, &w.leaf_b, (size_t)chunk * 4))) return rc;
	if ((rc = grow(m, &w.status, &w.status_b, (size_t)chunk))) return rc;
	for (int64_t o = 0; o < total; o += chunk) {
		const int64_t n = std::min(chunk, total - o);
		m->launches += launch_jitter_points(dm, ppl, seed, o, n, amp, phase, n_terms, (double *)w.q, (int32_t *)w.leaf, (double *)w.truth, st);
		if ((rc = eval_device(m, w, st, true, (const double *)w.q, n, (double *)w.out, 1, (int32_t *)w.leaf, (uint8_t *)w.status))) return rc;
		m->launches += launch_error_reduce(dm, (const double *)w.out, (const double *)w.truth, (const int32_t *)w.leaf, (const uint8_t *)w.status, n, threshold,
		                                   err_max, split, nullptr, st);
	}
	CUDA_OK(m, cudaGetLastError());
	return LOCO_OK;
}

// ---- multi-GPU result exchange: peer-mapped buffers + copy-engine pushes ------------------------------------------------
#define CUDA_OK0(expr)                                                                                        \
	do {                                                                                                      \
		cudaError_t _e = (expr);                                                                              \
		if (_e != cudaSuccess) return fail(nullptr, LOCO_ERR_CUDA, std::string(#expr ": ") + cudaGetErrorString(_e)); \
	} while (0)

int loco_peer_buffer_create(int device, size_t bytes, void **dptr, unsigned char handle[64])
{
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
	if (!dptr || !handle || bytes == 0) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	CUDA_OK0(cudaSetDevice(device));
	void *p = nullptr;
	CUDA_OK0(cudaMalloc(&p, bytes));
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) { cudaFree(p); return fail(nullptr, LOCO_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
	memcpy(handle, &h, 64);
	*dptr = p;
	return LOCO_OK;
}

int loco_peer_buffer_open(int device, const unsigned char handle[64], void **dptr)
{
	if (!dptr || !handle) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	CUDA_OK0(cudaSetDevice(device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	CUDA_OK0(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
	return LOCO_OK;
}

int loco_peer_buffer_close(int device, void *dptr)
{
	if (!dptr) return LOCO_OK;
	CUDA_OK0(cudaSetDevice(device));
	CUDA_OK0(cudaIpcCloseMemHandle(dptr));
	return LOCO_OK;
}

int loco_peer_buffer_free(int device, void *dptr)
{
	if (!dptr) return LOCO_OK;
	CUDA_OK0(cudaSetDevice(device));
	CUDA_OK0(cudaFree(dptr));
	return LOCO_OK;
}

int loco_peer_copy_async(void *dst, const void *src, size_t bytes, void *cuda_stream)
{
	if (bytes == 0) return LOCO_OK;
	if (!dst || !src) return fail(nullptr, LOCO_ERR_INVALID, "null argument");
	CUDA_OK0(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
	return LOCO_OK;
}

// ---- options ---------------------------------------------------------------------------------------------------
int loco_set_option(loco_model_t *m, const char *name, int64_t value)
{
	if (!m || !name) return fail(m, LOCO_ERR_INVALID, "null argument");
	if (!strcmp(name, "path")) { m->opt_path = value; return LOCO_OK; }
	if (!strcmp(name, "fp32")) { m->opt_fp32 = value; return LOCO_OK; }
	if (!strcmp(name, "mesh_simt")) { m->opt_mesh_simt = value; return LOCO_OK; }  // 1: DFMA register-tile mesh kernel instead of the fp64 MMA one
	if (!strcmp(name, "uniform")) { m->opt_uniform = value; m->dm.tensor_uniform = (m->flat.tensor_uniform && value) ? 1 : 0; return LOCO_OK; }  // 0: four separate factor evaluations
	if (!strcmp(name, "overlap")) { m->opt_overlap = value; return LOCO_OK; }  // streams the chunks of the sorted path are dealt over (1..4, default 2)
	if (!strcmp(name, "chunk")) { m->opt_chunk = value; return LOCO_OK; }  // queries per L2-resident chunk of the sorted path (0 = auto)
	if (!strcmp(name, "leaf_poly")) {  // 1 (default): scalar functions on single-piece tensor leaves evaluate the leaf's polynomial, 0: factor / contraction form
		m->opt_leaf_poly = value;
		m->dm.tl_poly = (value && m->dm.tp_stride > 0) ? m->d_tl_poly : nullptr;
		return LOCO_OK;
	}
	if (!strcmp(name, "fast_scatter")) { m->opt_fast_scatter = value; return LOCO_OK; }  // 1 (default): scatter kernel specialised for uniform locate grids / power-of-two domains
	if (!strcmp(name, "bulk_eval")) { m->opt_bulk_eval = value; return LOCO_OK; }  // 1: bucket evaluation from TMA-prefetched work items, 0 (default): register prefetch one group ahead
	if (!strcmp(name, "aggregate")) { m->opt_aggregate = value; return LOCO_OK; }  // 0: one slot atomic per query, 1 (default): one per run of equal leaves in a warp
	if (!strcmp(name, "adaptive")) { m->opt_adaptive = value; if (!value) m->skewed = false; return LOCO_OK; }  // 0: path 0 never leaves the leaf buckets
	if (!strcmp(name, "fused_errgen")) { m->opt_fused_errgen = value; return LOCO_OK; }  // 0: the three-kernel form of loco_error_check_generated_device
	if (!strcmp(name, "ecell_poly")) {  // 1: polynomial form of the evaluation cells no knot crosses (see loco_cellpoly.h), 0: term lists
		m->dm.epoly_idx = (value && m->dm.ecell_off) ? m->d_epoly_idx : nullptr;
		return LOCO_OK;
	}
	if (!strcmp(name, "ecells")) {
		// 0: sorted path keyed by leaf, 1: by evaluation cell where the model has them (default)
		if (!value) m->dm.epoly_idx = nullptr;  // the polynomial cells are a refinement of the evaluation cells
		if (m->dm.ecell_off) m->d_ecell_off = m->dm.ecell_off;
		m->dm.ecell_off = value ? m->d_ecell_off : nullptr;
		return LOCO_OK;
	}
	if (!strcmp(name, "lut")) {  // 0: plain k-d descent from the root, 1: locate grid first (default)
		if (m->dm.lut) m->d_lut = m->dm.lut;
		m->opt_lut = value;
		m->dm.lut = value ? m->d_lut : nullptr;
		return LOCO_OK;
	}
	return fail(m, LOCO_ERR_INVALID, std::string("unknown option ") + name);
}

int64_t loco_kernel_launches(const loco_model_t *m) { return m ? m->launches : 0; }

int loco_profile_begin(loco_model_t *m)
{
	if (!m) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	KernelProfile &p = kernel_profile();
	for (int i = 0; i < p.n; i++) { cudaEventDestroy(p.recs[i].a); cudaEventDestroy(p.recs[i].b); }
	p.n = 0;
	p.on = true;
	return LOCO_OK;
}

int loco_profile_end(loco_model_t *m, char *json, size_t cap)
{
	if (!m || !json || cap < 2) return fail(m, LOCO_ERR_INVALID, "null argument");
	NEED_DEVICE(m);
	KernelProfile &p = kernel_profile();
	p.on = false;
	CUDA_OK(m, cudaSetDevice(m->device));
	CUDA_OK(m, cudaDeviceSynchronize());
	struct Agg { const char *name; int64_t n; double ms; };
	std::vector<Agg> agg;
	for (int i = 0; i < p.n; i++) {
		float ms = 0;
		cudaEventElapsedTime(&ms, p.recs[i].a, p.recs[i].b);
		cudaEventDestroy(p.recs[i].a); cudaEventDestroy(p.recs[i].b);
		size_t k = 0;
		for (; k < agg.size(); k++) if (!strcmp(agg[k].name, p.recs[i].name)) break;
		if (k == agg.size()) agg.push_back(Agg{p.recs[i].name, 0, 0.0});
		agg[k].n++; agg[k].ms += ms;
	}
	p.n = 0;
	std::string out = "{";
	for (size_t k = 0; k < agg.size(); k++) {
		char buf[256];
		snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %lld, \"total_ms\": %.6f}", k ? ", " : "", agg[k].name, (long long)agg[k].n, agg[k].ms);
		out += buf;
	}
	out += "}";
	if (out.size() + 1 > cap) return fail(m, LOCO_ERR_RANGE, "profile buffer too small");
	memcpy(json, out.c_str(), out.size() + 1);
	return LOCO_OK;
}

} // extern "C"

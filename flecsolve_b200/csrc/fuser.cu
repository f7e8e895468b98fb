// Deferred-execution queue: groups queued statements into registered fused kernels.
//
// Why: the reference's Krylov loops issue one task per vector operation
// (solvers/cg.hh:95-130: 1 SpMV, 3 axpy, 2 dot, 1 norm, 1 preconditioner apply per
// iteration = 8 passes over HBM).  Keeping that call sequence as the API, the queue
// launches e.g. {x += a p; r -= a w; |r|^2} as ONE kernel when the norm is read.
// Statements are evaluated in program order per element, so fusing never changes a
// value; only reductions see a different (fixed, reproducible) summation order.
#include <chrono>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <unordered_map>

#include "ew_kernels.cuh"
#include "fsb_internal.h"
#include "setup_exchange.h"

namespace fsb {

// ---------------------------------------------------------------- registry

using launcher_t = void (*)(const ew_args &, int want_ctas, cudaStream_t);

// grid: enough CTAs for the data, at most one resident wave (occupancy API, cached per program).
// want_ctas < 0: no launch, only make sure the kernel is loaded (CUDA loads a kernel at its first use, and that load
// waits for everything running in the context: the ranks of an in-process group do it BEFORE they meet for a launch
// whose kernels wait for each other, see setup_exchange::rendezvous)
template<class Kernel>
static void launch_one_wave(Kernel kern, std::atomic<int> & resident, const ew_args & a, int want, cudaStream_t s) {
	if (resident.load(std::memory_order_acquire) == 0) {
		int nb = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, EW_BLOCK, 0) != cudaSuccess || nb < 1)
			nb = 4;
		resident.store(nb * SM_COUNT, std::memory_order_release);
	}
	if (want < 0)
		return;
	const int most = resident.load(std::memory_order_relaxed);
	launch_dependent(kern, dim3(want < most ? want : most), dim3(EW_BLOCK), 0, s, a);
}

template<class PT>
static void launch_program(const ew_args & a, int want, cudaStream_t s) {
	static std::atomic<int> resident{0};
	launch_one_wave(ew_program_kernel<PT>, resident, a, want, s);
}

template<class PT>
static void launch_program_dev(const ew_args & a, int want, cudaStream_t s) {
	static std::atomic<int> resident{0};
	launch_one_wave(ew_program_kernel<PT, true>, resident, a, want, s);
}

static std::string key_of(const program & p) {
	std::string k(reinterpret_cast<const char *>(&p.n), sizeof(int));
	k.append(reinterpret_cast<const char *>(p.st), sizeof(stmt) * p.n);
	return k;
}

struct registry_t {
	std::unordered_map<std::string, launcher_t> map;
	std::unordered_map<std::string, launcher_t> dev_map; // instantiations that read device-resident coefficients
	template<class PT>
	void add() {
		static_assert(PT::ok, "program exceeds slot limits");
		map.emplace(key_of(PT::value), &launch_program<PT>);
	}
	template<class PT>
	void add_dev() {
		static_assert(PT::ok, "program exceeds slot limits");
		dev_map.emplace(key_of(PT::value), &launch_program_dev<PT>);
	}
};

// Programs are written with small vector names; canonicalize() renames them.
#define FSB_PROGRAM(NAME, ...)                                                          \
	struct NAME {                                                                       \
		static constexpr raw_stmt raw[] = {__VA_ARGS__};                               \
		static constexpr canon_result cr = canonicalize(raw, sizeof(raw) / sizeof(raw_stmt)); \
		static constexpr bool ok = cr.ok;                                               \
		static constexpr program value = cr.p;                                          \
	};

// shorthand: {op, z, x, y}
#define S_SET(z) {OP_SET, z, -1, -1}
#define S_SCALE(z, x) {OP_SCALE, z, x, -1}
#define S_LIN2(z, x, y) {OP_LIN2, z, x, y}
#define S_MUL(z, x, y) {OP_MUL, z, x, y}
#define S_DIV(z, x, y) {OP_DIV, z, x, y}
#define S_RECIP(z, x) {OP_RECIP, z, x, -1}
#define S_ABS(z, x) {OP_ABS, z, x, -1}
#define S_ADDS(z, x) {OP_ADDS, z, x, -1}
#define R_DOT(x, y) {RD_DOT, -1, x, y}
#define R_ASUM(x) {RD_ASUM, -1, x, -1}
#define R_AMAX(x) {RD_AMAX, -1, x, -1}
#define R_MIN(x) {RD_MIN, -1, x, -1}
#define R_MAX(x) {RD_MAX, -1, x, -1}
#define R_POWSUM(x) {RD_POWSUM, -1, x, -1}

// ---- every single statement, in every alias pattern that survives normalisation ----
FSB_PROGRAM(p_set, S_SET(0))
FSB_PROGRAM(p_scale, S_SCALE(1, 0))
FSB_PROGRAM(p_scale_self, S_SCALE(0, 0))
FSB_PROGRAM(p_lin2, S_LIN2(2, 0, 1))
FSB_PROGRAM(p_lin2_zy, S_LIN2(1, 0, 1))
FSB_PROGRAM(p_lin2_xx, S_LIN2(1, 0, 0))
FSB_PROGRAM(p_lin2_xxx, S_LIN2(0, 0, 0))
FSB_PROGRAM(p_mul, S_MUL(2, 0, 1))
FSB_PROGRAM(p_mul_zy, S_MUL(1, 0, 1))
FSB_PROGRAM(p_mul_xx, S_MUL(1, 0, 0))
FSB_PROGRAM(p_mul_xxx, S_MUL(0, 0, 0))
FSB_PROGRAM(p_div, S_DIV(2, 0, 1))
FSB_PROGRAM(p_div_zx, S_DIV(0, 0, 1))
FSB_PROGRAM(p_div_zy, S_DIV(1, 0, 1))
FSB_PROGRAM(p_div_xx, S_DIV(1, 0, 0))
FSB_PROGRAM(p_div_xxx, S_DIV(0, 0, 0))
FSB_PROGRAM(p_recip, S_RECIP(1, 0))
FSB_PROGRAM(p_recip_self, S_RECIP(0, 0))
FSB_PROGRAM(p_abs, S_ABS(1, 0))
FSB_PROGRAM(p_abs_self, S_ABS(0, 0))
FSB_PROGRAM(p_adds, S_ADDS(1, 0))
FSB_PROGRAM(p_adds_self, S_ADDS(0, 0))
FSB_PROGRAM(p_dot, R_DOT(0, 1))
FSB_PROGRAM(p_sumsq, R_DOT(0, 0))
FSB_PROGRAM(p_asum, R_ASUM(0))
FSB_PROGRAM(p_amax, R_AMAX(0))
FSB_PROGRAM(p_min, R_MIN(0))
FSB_PROGRAM(p_max, R_MAX(0))
FSB_PROGRAM(p_powsum, R_POWSUM(0))

// ---- fused groups the Krylov loops produce (reference line numbers in comments) ----
// CG update, cg.hh:108-111:  x = a p + x ; r = -a w + r ; |r|^2
FSB_PROGRAM(p_cg_update, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), R_DOT(3, 3))
// same without the norm (diagnostic reads x first) and the tail alone
FSB_PROGRAM(p_axpy2, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3))
FSB_PROGRAM(p_axpy_sumsq, S_LIN2(1, 0, 1), R_DOT(1, 1))
FSB_PROGRAM(p_lin2_sumsq, S_LIN2(2, 0, 1), R_DOT(2, 2))
// Jacobi apply + r.z, cg.hh:124-127:  z = dinv * r ; r.z   (and z.r before the loop, cg.hh:87)
FSB_PROGRAM(p_jacobi_dot_rz, S_MUL(2, 0, 1), R_DOT(1, 2))
FSB_PROGRAM(p_jacobi_dot_zr, S_MUL(2, 0, 1), R_DOT(2, 1))
// identity preconditioner: z = r ; r.z
FSB_PROGRAM(p_copy_dot_rz, S_SCALE(1, 0), R_DOT(0, 1))
FSB_PROGRAM(p_copy_dot_zr, S_SCALE(1, 0), R_DOT(1, 0))
// residual r = b - r then its norm (operators/core.hh:140, cg.hh:71-74)
FSB_PROGRAM(p_resid_sumsq, S_LIN2(1, 0, 1), R_DOT(1, 1))
// GMRES modified Gram-Schmidt, gmres.hh:269-276:  v -= h q_j ; v.q_{j+1}   /  v -= h q_j ; |v|^2
FSB_PROGRAM(p_mgs_step, S_LIN2(1, 0, 1), R_DOT(1, 2))
// GMRES normalise + store + identity precondition, gmres.hh:156-162,143-146
FSB_PROGRAM(p_scale_copy, S_SCALE(0, 0), S_SCALE(1, 0))
FSB_PROGRAM(p_scale_copy_copy, S_SCALE(0, 0), S_SCALE(1, 0), S_SCALE(2, 1))
FSB_PROGRAM(p_scale_copy_mul, S_SCALE(0, 0), S_SCALE(1, 0), S_MUL(3, 2, 1))
FSB_PROGRAM(p_copy_copy, S_SCALE(1, 0), S_SCALE(2, 1))
FSB_PROGRAM(p_copy_mul, S_SCALE(1, 0), S_MUL(3, 2, 1))
// GMRES correction, gmres.hh:249-258: z = 0 ; z += y_i q_i (pairs)
FSB_PROGRAM(p_set_axpy, S_SET(0), S_LIN2(0, 1, 0))
FSB_PROGRAM(p_axpy_axpy_same, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1))
FSB_PROGRAM(p_axpy4_same, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1), S_LIN2(1, 3, 1), S_LIN2(1, 4, 1))
// BiCGStab, bicgstab.hh:120-124: p = -w v + p ; p = b p + res ; p_hat = P p
FSB_PROGRAM(p_bicg_dir_jacobi, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1), S_MUL(4, 3, 1))
FSB_PROGRAM(p_bicg_dir_copy, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1), S_SCALE(3, 1))
FSB_PROGRAM(p_bicg_dir, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1))
// bicgstab.hh:132-134: s = -a v + res ; |s|^2        (p_lin2_sumsq)
// bicgstab.hh:145: s_hat = P s
// bicgstab.hh:153-158: x += a p_hat ; x += w s_hat ; res = -w t + s ; |res|^2
FSB_PROGRAM(p_bicg_update, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1), S_LIN2(5, 3, 4), R_DOT(5, 5))
// two dots sharing an operand (bicgstab.hh:149-150 when both are queued)
FSB_PROGRAM(p_dot2, R_DOT(0, 0), R_DOT(0, 1))
// operator_adapter, time-integrators/operator_adapter.hh:29-35: y = -g y + x
// (single p_lin2_zy after normalisation)
// preconditioned residual pair used by left-preconditioned paths: z = d*r ; |z|^2
FSB_PROGRAM(p_mul_sumsq, S_MUL(2, 0, 1), R_DOT(2, 2))
// two independent norms / dots queued back to back (vec::multi with 2 components)
FSB_PROGRAM(p_sumsq2, R_DOT(0, 0), R_DOT(1, 1))
FSB_PROGRAM(p_dot_dot, R_DOT(0, 1), R_DOT(2, 3))
FSB_PROGRAM(p_lin2_lin2, S_LIN2(2, 0, 1), S_LIN2(5, 3, 4))
FSB_PROGRAM(p_scale2, S_SCALE(1, 0), S_SCALE(3, 2))
FSB_PROGRAM(p_set2, S_SET(0), S_SET(1))
// BiCGStab on a two-component vec::multi (vectors/operations/multi.hh applies every call per component):
// bicgstab.hh:132-134  s = -a v + res per component ; |s|^2 per component
FSB_PROGRAM(p_bicg2_s, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), R_DOT(3, 3), R_DOT(1, 1))
// bicgstab.hh:153-158  x += a p_hat ; x += w s_hat ; res = -w t + s ; |res|^2, both components
FSB_PROGRAM(p_bicg2_update, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(1, 4, 1), S_LIN2(3, 5, 3), S_LIN2(8, 6, 7),
            S_LIN2(11, 9, 10), R_DOT(11, 11), R_DOT(8, 8))
// bicgstab.hh:120-124  p = -w v + p ; p = b p + res ; p_hat = p (identity), both components
FSB_PROGRAM(p_bicg2_dir_copy, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(1, 4, 1), S_LIN2(3, 5, 3), S_SCALE(6, 1),
            S_SCALE(7, 3))

// Groups found by replaying the drivers' call traces through the queue's grouping rules
// (scripts/statement_groups.py -> profiles/r1_statement_groups.txt): once per iteration each.
// bicgstab.hh:132-134 on two components: s = -a v + res per component ; |s|^2 per component
FSB_PROGRAM(p_bicg2_half, S_LIN2(2, 0, 1), S_LIN2(5, 3, 4), R_DOT(5, 5), R_DOT(2, 2))
// cg.hh:108-111 on two components: x += a p ; r -= a w (both) ; |r|^2 (both)
FSB_PROGRAM(p_cg2_update, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(5, 4, 5), S_LIN2(7, 6, 7), R_DOT(7, 7), R_DOT(5, 5))
// cg.hh:124-127 on two components, identity preconditioner: z = r (both) ; r.z (both)
FSB_PROGRAM(p_cg2_copy_dot, S_SCALE(1, 0), S_SCALE(3, 2), R_DOT(2, 3), R_DOT(0, 1))
// fcg (cg.hh:196-203): d = -g/rho d + v ; q = -g/rho q + w ; u += a/rho d ; r -= a/rho q ; |r|^2
FSB_PROGRAM(p_fcg_update, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(4, 1, 4), S_LIN2(5, 3, 5), R_DOT(5, 5))
// bdf.hh source term + residual norm of the inner solve's right-hand side: two aliased axpys and a norm
FSB_PROGRAM(p_axpy2_same_sumsq, S_LIN2(1, 0, 1), S_LIN2(1, 2, 1), R_DOT(1, 1))
// bdf.hh:763-801 scaled truncation-error norm: max | (u - u_pred) / (rtol |u| + atol) |
FSB_PROGRAM(p_bdf_error_norm, S_SCALE(1, 0), S_SCALE(2, 0), S_ADDS(2, 2), S_LIN2(4, 0, 3), S_DIV(4, 4, 2), R_AMAX(4))

// Device-scalar CG (solvers/cg_device.hh): no host read separates the update from the preconditioner, so
//   x = a p + x ; r = -a w + r ; |r|^2 ; z = dinv * r ; r.z     is ONE pass (64 B/row)
FSB_PROGRAM(p_cgdev_update_jacobi, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), R_DOT(3, 3), S_MUL(5, 4, 3), R_DOT(3, 5))
FSB_PROGRAM(p_cgdev_update_copy, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), R_DOT(3, 3), S_SCALE(4, 3), R_DOT(3, 4))

// Single-reduction CG (solvers/cg_sr.hh): the whole vector side of an iteration is ONE pass
//   p = u + b p ; s = w + b s ; x += a p ; r -= a s ; |r|^2 ; u = dinv * r ; r.u        (96 B/row)
FSB_PROGRAM(p_cgsr_update_jacobi, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(4, 1, 4), S_LIN2(5, 3, 5), R_DOT(5, 5), S_MUL(0, 6, 5),
            R_DOT(5, 0))
FSB_PROGRAM(p_cgsr_update_copy, S_LIN2(1, 0, 1), S_LIN2(3, 2, 3), S_LIN2(4, 1, 4), S_LIN2(5, 3, 5), R_DOT(5, 5), S_SCALE(0, 5),
            R_DOT(5, 0))

static registry_t & registry() {
	static registry_t r = [] {
		registry_t g;
		g.add<p_set>(); g.add<p_scale>(); g.add<p_scale_self>();
		g.add<p_lin2>(); g.add<p_lin2_zy>(); g.add<p_lin2_xx>(); g.add<p_lin2_xxx>();
		g.add<p_mul>(); g.add<p_mul_zy>(); g.add<p_mul_xx>(); g.add<p_mul_xxx>();
		g.add<p_div>(); g.add<p_div_zx>(); g.add<p_div_zy>(); g.add<p_div_xx>(); g.add<p_div_xxx>();
		g.add<p_recip>(); g.add<p_recip_self>(); g.add<p_abs>(); g.add<p_abs_self>();
		g.add<p_adds>(); g.add<p_adds_self>();
		g.add<p_dot>(); g.add<p_sumsq>(); g.add<p_asum>(); g.add<p_amax>(); g.add<p_min>(); g.add<p_max>();
		g.add<p_powsum>();
		g.add<p_cg_update>(); g.add<p_axpy2>(); g.add<p_axpy_sumsq>(); g.add<p_lin2_sumsq>();
		g.add<p_jacobi_dot_rz>(); g.add<p_jacobi_dot_zr>(); g.add<p_copy_dot_rz>(); g.add<p_copy_dot_zr>();
		g.add<p_resid_sumsq>(); g.add<p_mgs_step>();
		g.add<p_scale_copy>(); g.add<p_scale_copy_copy>(); g.add<p_scale_copy_mul>();
		g.add<p_copy_copy>(); g.add<p_copy_mul>();
		g.add<p_set_axpy>(); g.add<p_axpy_axpy_same>(); g.add<p_axpy4_same>();
		g.add<p_bicg_dir_jacobi>(); g.add<p_bicg_dir_copy>(); g.add<p_bicg_dir>(); g.add<p_bicg_update>();
		g.add<p_dot2>(); g.add<p_mul_sumsq>(); g.add<p_sumsq2>(); g.add<p_dot_dot>();
		g.add<p_lin2_lin2>(); g.add<p_scale2>(); g.add<p_set2>();
		g.add<p_bicg2_s>(); g.add<p_bicg2_update>(); g.add<p_bicg2_dir_copy>();
		g.add<p_bicg2_half>(); g.add<p_cg2_update>(); g.add<p_cg2_copy_dot>(); g.add<p_fcg_update>();
		g.add<p_axpy2_same_sumsq>(); g.add<p_bdf_error_norm>();
		g.add_dev<p_cgdev_update_jacobi>(); g.add_dev<p_cgdev_update_copy>(); g.add_dev<p_lin2_zy>();
		g.add_dev<p_cg_update>(); g.add_dev<p_axpy2>();
		g.add_dev<p_cgsr_update_jacobi>(); g.add_dev<p_cgsr_update_copy>();
		return g;
	}();
	return r;
}

// diagnostics (fsb_debug_program_info): is there an ahead-of-time kernel for this canonical program?
bool program_is_registered(const program & p, bool dev) {
	const auto & table = dev ? registry().dev_map : registry().map;
	return table.find(key_of(p)) != table.end();
}

// ---------------------------------------------------------------- queue

static int64_t length_of(const pending & p) {
	return p.kind == pending::RED ? p.x->n_owned : p.z->n_owned;
}
// statements fuse only over the same index set: equal length and, for structured-grid vectors, equal box
static const fsb_vec_s * layout_of(const pending & p) { return p.kind == pending::RED ? p.x : p.z; }
static bool same_layout(const pending & a, const pending & b) {
	const fsb_vec_s *u = layout_of(a), *v = layout_of(b);
	return u->n_owned == v->n_owned && u->box == v->box && (!u->box || u->shape == v->shape);
}

static std::string describe(const pending * q, int n) {
	static const char * names[] = {"set", "scale", "lin2", "mul", "div", "recip", "abs", "adds"};
	static const char * rnames[] = {"dot", "asum", "amax", "min", "max", "powsum"};
	std::string s;
	for (int i = 0; i < n; ++i) {
		if (i)
			s += " ; ";
		if (q[i].kind == pending::SPMV) {
			s += "spmv";
			continue;
		}
		s += is_reduction(q[i].op) ? rnames[q[i].op - RD_FIRST] : names[q[i].op];
		auto id = [](fsb_vec_s * v) { return v ? std::to_string(v->id) : std::string("-"); };
		s += "(z" + id(q[i].z) + ",x" + id(q[i].x) + ",y" + id(q[i].y) + ")";
		if (q[i].kind == pending::RED)
			s += "#" + std::to_string(q[i].token);
	}
	return s;
}

struct bound_group {
	launcher_t launch = nullptr; // registered instantiation, or nullptr => run-time compiled / generic kernel with `prog`
	ew_args args{};
	program prog{};
	int nr = 0;
	bool dev = false; // device-resident coefficients or halt flag in play
	bool box = false; // structured-grid layout
};

static void launch_interp(const bound_group & g, int want, cudaStream_t s) { // want < 0: load only, as launch_one_wave
	static std::atomic<int> resident{0};
	if (resident.load(std::memory_order_acquire) == 0) {
		int nb = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ew_interp_kernel<0>, EW_BLOCK, 0) != cudaSuccess || nb < 1)
			nb = 4;
		resident.store(nb * SM_COUNT, std::memory_order_release);
	}
	if (want < 0)
		return;
	interp_args ia;
	ia.a = g.args;
	ia.p = g.prog;
	const int most = resident.load(std::memory_order_relaxed);
	launch_dependent(ew_interp_kernel<0>, dim3(want < most ? want : most), dim3(EW_BLOCK), 0, s, ia);
}

// bind queue[i, i+len) to a registered kernel; with allow_generic, to the generic kernel if there is none
static bool bind_group(fsb_ctx_s * c, const pending * q, int len, bound_group & g, bool allow_generic = false) {
	raw_stmt rs[MAXS];
	fsb_vec_s * vecs[3 * MAXS];
	int nvec = 0;
	// slots are keyed on the device storage, not on the handle: two handles over the same memory (fsb_vec_wrap) are one
	// slot, so a statement reading through one sees what an earlier statement of the group wrote through the other
	auto id_of = [&](fsb_vec_s * v) {
		for (int k = 0; k < nvec; ++k)
			if (vecs[k]->d == v->d)
				return k;
		vecs[nvec] = v;
		return nvec++;
	};
	for (int i = 0; i < len; ++i) {
		rs[i].op = q[i].op;
		rs[i].x = reads_x(q[i].op) ? id_of(q[i].x) : -1;
		rs[i].y = reads_y(q[i].op) ? id_of(q[i].y) : -1;
		rs[i].z = is_reduction(q[i].op) ? -1 : id_of(q[i].z);
	}
	const canon_result cr = canonicalize(rs, len);
	if (!cr.ok)
		return false;
	bool dev = c->halt_armed;
	for (int i = 0; i < len; ++i)
		dev = dev || q[i].a_num >= 0 || q[i].b_num >= 0;
	const auto & table = dev ? registry().dev_map : registry().map;
	auto it = table.find(key_of(cr.p));
	if (it == table.end() && !allow_generic)
		return false;
	g.launch = it == table.end() ? nullptr : it->second;
	g.dev = dev;
	g.box = layout_of(q[0])->box;
	if (g.box)
		g.launch = nullptr; // no ahead-of-time instantiation carries the structured-grid loop
	g.prog = cr.p;
	g.nr = cr.p.nr;
	ew_args & a = g.args;
	for (int k = 0; k < cr.p.nv; ++k)
		a.v[k] = vecs[cr.id_of_slot[k]]->d;
	a.n = length_of(q[0]);
	if (const fsb_vec_s * lay = layout_of(q[0]); lay->box) {
		a.bn0 = lay->shape.n[0];
		a.bn1 = lay->shape.n[1];
		a.bE0 = lay->shape.ext[0];
		a.bE01 = lay->shape.ext[0] * lay->shape.ext[1];
		a.boff = lay->shape.origin();
	}
	a.partials = c->d_partials;
	a.counter = c->d_counter;
	a.partial_stride = MAX_RED_BLOCKS;
	a.xr = c->d_xrank;
	a.sdev = c->d_scalars;
	a.halt = c->halt_armed ? c->d_halt : nullptr;
	for (int k = 0; k < MAXSC; ++k)
		a.snum[k] = a.sden[k] = -1;
	for (int i = 0; i < len; ++i) {
		const stmt & s = cr.p.st[i];
		if (s.a >= 0) {
			a.s[s.a] = q[i].a;
			a.snum[s.a] = static_cast<signed char>(q[i].a_num);
			a.sden[s.a] = static_cast<signed char>(q[i].a_den);
		}
		if (s.b >= 0) {
			a.s[s.b] = q[i].b;
			a.snum[s.b] = static_cast<signed char>(q[i].b_num);
			a.sden[s.b] = static_cast<signed char>(q[i].b_den);
		}
		if (is_reduction(s.op))
			fill_red_out(c, q[i], a.r[s.z]);
	}
	return true;
}

// the scalar statements a reduction carries (fsb_red_opts::post), in the compact form the kernels read
static void fill_post(fsb_ctx_s * c, const pending & red, red_out & r) {
	r.slots = c->d_scalars;
	r.n_post = std::min(red.n_post, FSB_MAX_POST_OPS);
	for (int k = 0; k < FSB_MAX_POST_OPS; ++k) {
		if (k >= r.n_post)
			break;
		r.post[k][0] = static_cast<unsigned char>(red.post[k].op);
		r.post[k][1] = static_cast<unsigned char>(red.post[k].dst);
		r.post[k][2] = static_cast<unsigned char>(red.post[k].a);
		r.post[k][3] = static_cast<unsigned char>(red.post[k].b);
	}
}

// where a finished reduction goes.  One rank / peer-memory all-reduce: the producing kernel has the
// all-rank value, so it also serves the scalar slot and the convergence test; NCCL transport: the
// kernel only leaves the rank-local value in d_results and finish_reduction_nccl() does the rest.
void fill_red_out(fsb_ctx_s * c, const pending & red, red_out & r) {
	const int slot = static_cast<int>(red.token % FSB_RED_RING);
	r = red_out{};
	r.d_value = c->d_results + slot;
	r.token = red.token;
	if (c->nranks == 1 || c->d_xrank) {
		r.h_ll = c->h_ll_dev + slot;
		r.d_extra = red.store >= 0 ? c->d_scalars + red.store : nullptr;
		r.halt = c->d_halt;
		r.halt_thr = red.halt_thr;
		r.halt_mode = red.halt_mode;
		fill_post(c, red, r);
	}
}

__global__ void after_allreduce_kernel(const double * value, double * extra, int * halt, int mode, double thr, red_out post) {
	const double t = *value;
	if (extra)
		*extra = t;
	if (mode != 0 && (mode == 1 ? sqrt(t) : t) < thr)
		*halt = 1;
	run_post_ops(post);
}

void finish_reduction_nccl(fsb_ctx_s * c, const pending & red) {
	const int slot = static_cast<int>(red.token % FSB_RED_RING);
	const int f = fold_of(red.op);
	const ncclRedOp_t op = f == 0 ? ncclSum : (f == 1 ? ncclMax : ncclMin);
	FSB_NCCL(ncclAllReduce(c->d_results + slot, c->d_results + slot, 1, ncclDouble, op, c->nccl, c->stream));
	c->stats[FSB_STAT_ALLREDUCES]++;
	if (red.store >= 0 || red.halt_mode != 0 || red.n_post > 0) {
		red_out post{};
		fill_post(c, red, post);
		after_allreduce_kernel<<<1, 1, 0, c->stream>>>(c->d_results + slot, red.store >= 0 ? c->d_scalars + red.store : nullptr,
		                                               c->d_halt, red.halt_mode, red.halt_thr, post);
		FSB_CUDA(cudaGetLastError());
		c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	}
	FSB_CUDA(cudaMemcpyAsync(c->h_results + slot, c->d_results + slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	FSB_CUDA(cudaEventRecord(c->token_event[slot], c->stream));
}

// After a kernel that produced intra-rank reduction values: finish across ranks.
static void publish_multi_rank(fsb_ctx_s * c, const pending * q, int len) {
	if (c->nranks == 1 || c->d_xrank) // the producing kernel already reduced across ranks
		return;
	for (int i = 0; i < len; ++i)
		if (q[i].kind == pending::RED)
			finish_reduction_nccl(c, q[i]);
}

// ---------------------------------------------------------------- armed launches (FSB_OPT_SPECULATE)
//
// A Krylov loop written against the vector API alternates "wait for a reduction" and "launch the statements whose
// coefficients come from it" (cg.hh:98-131: three such round trips per iteration).  Each costs the result's way to the
// host, the host's arithmetic, a launch call (~3.7 us here) and the launch latency.  The loop is periodic, so the
// queue remembers, per launched group, which group followed the wait after it; the next time the host waits at that
// point, that group's kernel is launched AHEAD of the result with its coefficients left open (ew_kernels.cuh:
// late_wait).  When the host then issues exactly that group -- same kernel, same vectors, the reduction tokens the
// kernel was given -- the coefficients and a GO word in mapped memory replace the launch; anything else, and the kernel
// is told to leave (it has touched nothing).  Plain reductions and immediate coefficients only; registered kernels only.
constexpr int LATE_RING = 8;

struct follower {
	launcher_t launch = nullptr;
	double * v[MAXV] = {};
	int nv = 0, ns = 0, nr = 0, n_red = 0;
	long long n = 0;
	int red_slot[MAXR] = {}; // canonical slot of the j-th reduction statement, in statement order
	int red_op[MAXR] = {};
	bool confirmed = false; // seen twice in a row behind the same group: worth launching ahead
	bool same_group(const follower & o) const {
		if (launch != o.launch || nv != o.nv || n != o.n || n_red != o.n_red)
			return false;
		for (int k = 0; k < nv; ++k)
			if (v[k] != o.v[k])
				return false;
		for (int j = 0; j < n_red; ++j)
			if (red_slot[j] != o.red_slot[j] || red_op[j] != o.red_op[j])
				return false;
		return true;
	}
};

struct speculation {
	late_host * h_late = nullptr; // [LATE_RING], mapped pinned
	late_host * h_late_dev = nullptr;
	late_dev * d_relay = nullptr; // [LATE_RING]
	int * h_error = nullptr; // mapped: an armed kernel gave up waiting
	int * h_error_dev = nullptr;
	unsigned seq = 0;
	std::unordered_map<uint64_t, follower> followers;
	uint64_t last_sig = 0; // the last group the current flush launched
	bool last_valid = false;
	uint64_t learn_sig = 0; // the group the host last waited behind
	bool learning = false;
	bool armed = false;
	follower armed_f;
	unsigned armed_seq = 0;
	int64_t armed_tokens[MAXR] = {};
};

static uint64_t sig_mix(uint64_t h, uint64_t v) {
	h ^= v + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2);
	return h;
}

static speculation * spec_of(fsb_ctx_s * c) {
	if (!c->spec) {
		auto * sp = new speculation;
		FSB_CUDA(cudaHostAlloc(&sp->h_late, sizeof(late_host) * LATE_RING, cudaHostAllocMapped));
		std::memset(sp->h_late, 0, sizeof(late_host) * LATE_RING);
		FSB_CUDA(cudaHostGetDevicePointer(&sp->h_late_dev, sp->h_late, 0));
		FSB_CUDA(cudaMalloc(&sp->d_relay, sizeof(late_dev) * LATE_RING));
		FSB_CUDA(cudaMemset(sp->d_relay, 0, sizeof(late_dev) * LATE_RING));
		FSB_CUDA(cudaHostAlloc(&sp->h_error, sizeof(int), cudaHostAllocMapped));
		*sp->h_error = 0;
		FSB_CUDA(cudaHostGetDevicePointer(&sp->h_error_dev, sp->h_error, 0));
		c->spec = sp;
	}
	return c->spec;
}

static void late_verdict(speculation * sp, unsigned verdict) {
	late_host & h = sp->h_late[sp->armed_seq % LATE_RING];
	__atomic_store_n(&h.word, (static_cast<unsigned long long>(sp->armed_seq) << 2) | verdict, __ATOMIC_RELEASE);
	sp->armed = false;
}

void speculation_release(fsb_ctx_s * c) {
	speculation * sp = c->spec;
	if (!sp)
		return;
	sp->learning = false;
	if (sp->armed) {
		late_verdict(sp, LATE_ABORT);
		c->stats[FSB_STAT_ARMED_MISSES]++;
	}
}

bool speculation_failed(const fsb_ctx_s * c) { return c->spec && *reinterpret_cast<volatile int *>(c->spec->h_error) != 0; }

void speculation_destroy(fsb_ctx_s * c) {
	speculation * sp = c->spec;
	if (!sp)
		return;
	cudaFreeHost(sp->h_late);
	cudaFreeHost(sp->h_error);
	cudaFree(sp->d_relay);
	delete sp;
	c->spec = nullptr;
}

// the group just bound as the template of a launch ahead of time, or false if it does not qualify
static bool follower_of(const fsb_ctx_s * c, const bound_group & g, const pending * q, int len, follower & f) {
	if (!g.launch || g.dev || g.box || c->halt_armed)
		return false;
	f = follower{};
	f.launch = g.launch;
	f.nv = g.prog.nv;
	f.ns = g.prog.ns;
	f.nr = g.prog.nr;
	f.n = g.args.n;
	for (int k = 0; k < f.nv; ++k)
		f.v[k] = g.args.v[k];
	for (int i = 0; i < len; ++i) {
		if (q[i].kind != pending::RED)
			continue;
		if (q[i].store >= 0 || q[i].halt_mode != 0 || q[i].n_post != 0 || f.n_red >= MAXR)
			return false; // reductions that feed device scalars or the halt flag belong to the device-scalar solvers
		f.red_slot[f.n_red] = g.prog.st[i].z;
		f.red_op[f.n_red] = q[i].op;
		++f.n_red;
	}
	return true;
}

static uint64_t sig_of_group(const bound_group & g) {
	uint64_t h = sig_mix(0x45, reinterpret_cast<uint64_t>(g.launch));
	for (int k = 0; k < g.prog.nv; ++k)
		h = sig_mix(h, reinterpret_cast<uint64_t>(g.args.v[k]));
	return sig_mix(h, static_cast<uint64_t>(g.args.n));
}
static uint64_t sig_of_spmv(const pending & sp, const pending * dot) {
	uint64_t h = sig_mix(0x53, reinterpret_cast<uint64_t>(sp.A));
	h = sig_mix(h, reinterpret_cast<uint64_t>(sp.x->d));
	h = sig_mix(h, reinterpret_cast<uint64_t>(sp.y->d));
	return dot ? sig_mix(h, reinterpret_cast<uint64_t>((dot->x == sp.y ? dot->y : dot->x)->d)) : h;
}

// does the group the host issued equal the one that is armed?  (then only its coefficients travel)
static bool matches_armed(const speculation * sp, const bound_group & g, const pending * q, int len) {
	const follower & f = sp->armed_f;
	if (g.launch != f.launch || g.dev || g.box || g.args.n != f.n || g.prog.nv != f.nv || g.prog.nr != f.nr)
		return false;
	for (int k = 0; k < f.nv; ++k)
		if (g.args.v[k] != f.v[k])
			return false;
	int j = 0;
	for (int i = 0; i < len; ++i) {
		if (q[i].kind != pending::RED)
			continue;
		if (j >= f.n_red || q[i].store >= 0 || q[i].halt_mode != 0 || q[i].n_post != 0 || q[i].token != sp->armed_tokens[j] ||
		    g.prog.st[i].z != f.red_slot[j])
			return false;
		++j;
	}
	return j == f.n_red;
}

// wait_token, right after its flush: the host is about to wait behind the group that flush launched last.  Remember to
// learn what it issues next, and if that is known from the last time round, launch it now with open coefficients.
void speculate_after_flush(fsb_ctx_s * c) {
	if (c->trace || c->d_timeline || c->halt_armed || (c->boot && c->boot->in_process()) || (c->nranks > 1 && !c->d_xrank))
		return;
	speculation * sp = spec_of(c);
	if (!sp->last_valid || sp->armed)
		return;
	sp->last_valid = false;
	sp->learn_sig = sp->last_sig;
	sp->learning = true;
	const auto it = sp->followers.find(sp->last_sig);
	if (it == sp->followers.end() || !it->second.confirmed)
		return;
	const follower & f = it->second;
	ew_args a{};
	for (int k = 0; k < f.nv; ++k)
		a.v[k] = f.v[k];
	a.n = f.n;
	a.partials = c->d_partials;
	a.counter = c->d_counter;
	a.partial_stride = MAX_RED_BLOCKS;
	a.xr = c->d_xrank;
	a.sdev = c->d_scalars;
	for (int k = 0; k < MAXSC; ++k)
		a.snum[k] = a.sden[k] = -1;
	for (int j = 0; j < f.n_red; ++j) { // the tokens the host's next reductions will get
		pending red{};
		red.kind = pending::RED;
		red.op = f.red_op[j];
		red.token = c->next_token + j;
		sp->armed_tokens[j] = red.token;
		fill_red_out(c, red, a.r[f.red_slot[j]]);
	}
	sp->armed_seq = ++sp->seq;
	const int slot = static_cast<int>(sp->armed_seq % LATE_RING);
	a.late = sp->h_late_dev + slot;
	a.late_relay = sp->d_relay + slot;
	a.late_error = sp->h_error_dev;
	a.late_seq = sp->armed_seq;
	const long long want = ((f.n + 1) / 2 + EW_BLOCK - 1) / EW_BLOCK;
	const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, MAX_RED_BLOCKS)));
	f.launch(a, grid, c->stream);
	FSB_CUDA(cudaGetLastError());
	c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
	sp->armed_f = f;
	sp->armed = true;
}

// breakdown of the flush time (FSB_STAT_BIND_NS, FSB_STAT_EW_LAUNCH_NS, FSB_STAT_SPMV_LAUNCH_NS)
struct lap {
	int64_t & acc;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	explicit lap(int64_t & a) : acc(a) {}
	~lap() { acc += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(); }
};

void flush(fsb_ctx_s * c, bool waiting) {
	if (c->queue.empty()) {
		if (!waiting) // whoever flushes without statements is about to touch the stream or the device; a host that merely
			speculation_release(c); // reads another result of the kernel it already waited for is not
		if (c->spec)
			c->spec->last_valid = false;
		return;
	}
	speculation * const sp = c->spec;
	if (sp)
		sp->last_valid = false;
	struct stopwatch {
		fsb_ctx_s * c;
		std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
		~stopwatch() {
			c->stats[FSB_STAT_FLUSH_NS] +=
				std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
		}
	} watch{c};
	std::vector<pending> q;
	q.swap(c->queue); // launches below must not re-enter the queue
	const int total = static_cast<int>(q.size());
	int i = 0;
	while (i < total) {
		if (q[i].kind == pending::SPMV) {
			// y = A x, optionally fused with a following dot that involves y
			// (the dot may sit behind other SpMVs and dots of the same burst — vec::multi issues
			// A0 x0, A1 x1, <r0,y0>, <r1,y1> — it commutes with them unless one of them writes an operand)
			const pending * dot = nullptr;
			if (c->fusion) {
				for (int j = i + 1; j < total; ++j) {
					if (q[j].kind == pending::SPMV) {
						if (q[j].y == q[i].y)
							break;
						continue;
					}
					if (q[j].kind != pending::RED || q[j].op != RD_DOT)
						break;
					if (q[j].x != q[i].y && q[j].y != q[i].y)
						continue;
					fsb_vec_s * u = q[j].x == q[i].y ? q[j].y : q[j].x;
					bool clobbered = false;
					for (int k = i + 1; k < j; ++k)
						clobbered |= q[k].kind == pending::SPMV && q[k].y == u;
					if (!clobbered) {
						std::rotate(q.begin() + i + 1, q.begin() + j, q.begin() + j + 1);
						dot = &q[i + 1];
					}
					break;
				}
			}
			if (sp && i == 0) { // the host went on with a product: nothing that could be armed
				if (sp->learning)
					sp->followers.erase(sp->learn_sig);
				speculation_release(c);
			}
			if (sp) {
				sp->last_sig = sig_of_spmv(q[i], dot);
				sp->last_valid = true;
			}
			if (c->trace)
				fprintf(stderr, "[fsb %d] launch: spmv%s%s\n", c->rank, dot ? " + dot #" : "", dot ? std::to_string(dot->token).c_str() : "");
			{
				lap t(c->stats[FSB_STAT_SPMV_LAUNCH_NS]);
				spmv_group(c, q[i], dot);
			}
			if (dot) {
				publish_multi_rank(c, dot, 1);
				c->stats[FSB_STAT_FUSED_STATEMENTS] += 2;
			}
			i += dot ? 2 : 1;
			continue;
		}
		int j = i;
		const int64_t n = length_of(q[i]);
		const int cap = c->fusion ? MAXS : 1;
		// a statement whose coefficient is produced by a reduction of this very run needs the finished
		// (grid-wide, all-rank) value: it starts the next launch
		auto needs_result_of_run = [&](int k) {
			auto used = [&](int slot) {
				return slot > 0 && (slot == q[k].a_num || slot == q[k].a_den || slot == q[k].b_num || slot == q[k].b_den);
			};
			for (int m = i; m < k; ++m) {
				if (q[m].kind != pending::RED)
					continue;
				if (q[m].store >= 0 && used(q[m].store))
					return true;
				for (int t = 0; t < q[m].n_post; ++t)
					if (used(q[m].post[t].dst))
						return true;
			}
			return false;
		};
		// operands that overlap in memory without being the same vector (wrapped sub-ranges) cannot share a launch:
		// element i of one is element i + k of the other
		auto overlaps_partially = [&](int k) {
			fsb_vec_s * mine[3] = {q[k].z, q[k].x, q[k].y};
			for (int m = i; m <= k; ++m) {
				fsb_vec_s * theirs[3] = {q[m].z, q[m].x, q[m].y};
				for (fsb_vec_s * u : mine)
					for (fsb_vec_s * v : theirs) {
						if (!u || !v || u->d == v->d || (u->owns && v->owns))
							continue;
						const double *ub = u->d, *ue = u->d + u->n_owned + u->n_ghost, *vb = v->d, *ve = v->d + v->n_owned + v->n_ghost;
						if (ub < ve && vb < ue)
							return true;
					}
			}
			return false;
		};
		if (overlaps_partially(i))
			throw error(FSB_ERR_ARG, "operands of one statement overlap in memory without being the same vector");
		while (j < total && j - i < cap && q[j].kind != pending::SPMV && same_layout(q[i], q[j]) && !needs_result_of_run(j) &&
		       !overlaps_partially(j))
			++j;
		// the whole run as one launch: a registered instantiation if there is one, else the generic
		// program kernel; a run that exceeds the slot limits is cut at the longest prefix that fits
		bound_group g;
		int len = j - i;
		{
			lap t(c->stats[FSB_STAT_BIND_NS]);
			for (; len >= 1; --len)
				if (bind_group(c, &q[i], len, g, true))
					break;
		}
		if (len < 1)
			throw error(FSB_ERR_STATE, "no kernel for statement: " + describe(&q[i], 1));
		bool ran_armed = false;
		if (sp && i == 0) { // first group after a wait: is it the one that was launched ahead?  what follows that wait?
			if (sp->learning) {
				follower f;
				if (follower_of(c, g, &q[i], len, f)) {
					const auto known = sp->followers.find(sp->learn_sig);
					f.confirmed = known != sp->followers.end() && known->second.same_group(f);
					sp->followers[sp->learn_sig] = f;
				}
				else
					sp->followers.erase(sp->learn_sig);
				sp->learning = false;
			}
			if (sp->armed) {
				if (matches_armed(sp, g, &q[i], len)) {
					late_host & h = sp->h_late[sp->armed_seq % LATE_RING];
					for (int k = 0; k < g.prog.ns; ++k)
						h.s[k] = g.args.s[k];
					late_verdict(sp, LATE_GO); // release store: the coefficients are visible before the word
					c->stats[FSB_STAT_ARMED_HITS]++;
					ran_armed = true;
				}
				else
					speculation_release(c);
			}
		}
		if (sp) {
			sp->last_sig = sig_of_group(g);
			sp->last_valid = g.launch != nullptr;
		}
		if (!g.launch)
			c->stats[FSB_STAT_UNMATCHED_GROUPS]++;
		if (c->trace)
			fprintf(stderr, "[fsb %d] launch%s: %s\n", c->rank, g.launch ? "" : " (generic)", describe(&q[i], len).c_str());
		if (ran_armed) {
			// the kernel is already resident: it picked the coefficients up and is running
		}
		else if (n > 0 || (g.nr > 0 && c->nranks > 1)) {
			long long packets = layout_of(q[i])->box ? n : (n + 1) / 2; // box layout: one element per thread and trip
			long long want = (packets + EW_BLOCK - 1) / EW_BLOCK;
			const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(want, MAX_RED_BLOCKS)));
			g.args.tl = timeline_slot(c, TL_KIND_EW + len); // kind = 16 + number of statements
			int resident = 0;
			const void * jk = (!g.launch && c->jit) ? jit_kernel(g.prog, g.dev, g.box, &resident) : nullptr;
			if (g.nr > 0 && c->d_xrank && c->boot && c->boot->in_process()) {
				// the kernel's last CTA waits for the peers' values: meet them first, with the kernel compiled and loaded
				if (g.launch)
					g.launch(g.args, -1, c->stream);
				else if (!jk)
					launch_interp(g, -1, c->stream);
				c->boot->rendezvous();
			}
			lap t(c->stats[FSB_STAT_EW_LAUNCH_NS]);
			if (g.launch)
				g.launch(g.args, grid, c->stream);
			else if (jk) {
				jit_launch(jk, resident, g.args, grid, c->stream);
				c->stats[FSB_STAT_JIT_GROUPS]++;
			}
			else
				launch_interp(g, grid, c->stream);
			FSB_CUDA(cudaGetLastError());
			c->stats[FSB_STAT_KERNEL_LAUNCHES]++;
			if (g.nr > 0 && c->d_xrank && c->boot && c->boot->in_process())
				c->boot->rendezvous(); // every rank's kernel of this step is launched before any rank goes on (setup_exchange.h)
		}
		else if (g.nr > 0) {
			// empty local part: publish fold identities without a kernel
			for (int k = i; k < i + len; ++k) {
				if (q[k].kind != pending::RED)
					continue;
				const int slot = static_cast<int>(q[k].token % FSB_RED_RING);
				const int f = fold_of(q[k].op);
				const double ident = f == 0 ? 0.0 : (f == 1 ? -HUGE_VAL : HUGE_VAL);
				FSB_CUDA(cudaMemcpyAsync(c->d_results + slot, &ident, sizeof(double), cudaMemcpyHostToDevice, c->stream));
				FSB_CUDA(cudaStreamSynchronize(c->stream));
				c->h_results[slot] = ident;
				host_ll_store(c->h_ll + slot, ident, static_cast<uint32_t>(q[k].token));
			}
		}
		for (int k = i; k < i + len; ++k) // ghost copies of overwritten vectors are stale from here on
			if (q[k].kind == pending::EW)
				q[k].z->halo_valid = false;
		if (len > 1)
			c->stats[FSB_STAT_FUSED_STATEMENTS] += len;
		publish_multi_rank(c, &q[i], len);
		i += len;
	}
}

void enqueue(fsb_ctx_s * c, const pending & p) {
	c->queue.push_back(p);
	if (!c->fusion || c->queue.size() >= 64)
		flush(c);
}

int64_t new_token(fsb_ctx_s * c, int fold) {
	const int64_t t = c->next_token++;
	c->token_op[t % FSB_RED_RING] = fold;
	return t;
}

} // namespace fsb

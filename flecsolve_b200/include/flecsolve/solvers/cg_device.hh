// Preconditioned CG with the Krylov scalars kept on the device.
//
// Same recurrence, same operation order and the same floating-point expressions as op::cg
// (solvers/cg.hh; reference flecsolve/solvers/cg.hh:44-139), but no reduction is read by the host
// inside the iteration:
//   * <p,Ap>, <r,z> stay in device scalars; alpha = rho / <p,Ap> and beta = rho' / rho are
//     evaluated by the kernels that use them (one IEEE division, then the +-1 scaling);
//   * ||r||^2 carries the convergence test  sqrt(.) < terminate_tol : the kernel that finishes
//     the reduction raises the context's halt flag, and from then on every vector update is a
//     no-op, so x stays the iterate of the converging iteration exactly;
//   * the host looks at residual norms `lag` iterations late (diagnostics, status, history) --
//     by then the value is long finished, so nothing ever waits and the GPU always holds `lag`
//     iterations of queued work.
// Without the host read between the update and the preconditioner, the queue fuses
//   { w = A p ; <w,p> }  { x += a p ; r -= a w ; |r|^2 ; z = P r ; <r,z> }  { p = b p + z }
// = 3 kernels and 12 nnz + 108 N bytes per iteration (SURVEY 8(d)'s minimum).
//
// Differences a caller can observe, all consequences of the late look:
//   * the diagnostic is called with the residual norm of iteration j while x may already be up
//     to `lag` iterations further (never past the converged iterate); a `true` return stops the
//     solve with x where it is;
//   * single vectors on the csr topology only (device::linear_sum / device::dot).
#ifndef FLECSOLVE_B200_SOLVERS_CG_DEVICE_HH
#define FLECSOLVE_B200_SOLVERS_CG_DEVICE_HH

#include <array>
#include <cmath>
#include <iostream>

#include "flecsolve/device/scalar.hh"
#include "flecsolve/solvers/cg.hh"

namespace flecsolve::cg_device {
struct settings : solver_settings {
	int lag = 2; // iterations the host runs ahead of the residual norm it inspects
};
struct options : solver_options {
	using settings_type = settings;
	options(const char * pre) : solver_options(pre) {}
	po::options_description operator()(settings_type & s) {
		auto desc = solver_options::operator()(s);
		desc.add_options()(label("lag").c_str(), po::value<int>(&s.lag)->default_value(2),
		                   "iterations issued ahead of the inspected residual norm");
		return desc;
	}
};
}

namespace flecsolve::op {

template<class Params>
struct cg_device : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;
	using real = typename Params::real;
	using scalar = typename Params::scalar;

	cg_device(Params p) : base_t(std::move(p)) {}

	const auto & get_operator() const { return params.A(); }

	template<class DomainVec, class RangeVec>
	solve_info apply(const RangeVec & b, DomainVec & x) const {
		using real = typename DomainVec::real;
		using stop = solve_info::stop_reason;
		constexpr int max_lag = 8;

		solve_info info;
		const auto & A = params.A();
		const auto & P = params.P();
		auto & diagnostic = params.ops.diagnostic;
		const auto & settings = params.settings;
		auto & [r, z, p, w] = params.work;
		const int lag = settings.lag < 0 ? 0 : (settings.lag > max_lag ? max_lag : settings.lag);

		real terminate_tol, current_res;
		if (detail::krylov_start(A, settings, b, x, r, info, terminate_tol, current_res))
			return info;

		fsb_ctx_t ctx = x.data.ctx();
		device::scalar rho_a(ctx), rho_b(ctx), curvature(ctx);
		device::scalar * rho = &rho_a, * rho_next = &rho_b;

		P.apply(r, z);
		device::dot(z, r, rho);
		p.copy(z);

		std::array<device_future, max_lag + 1> res_sq, curv;
		device::halt_scope halting(ctx);
		bool done = false;
		// host's late look at iteration j; true ends the solve
		auto inspect = [&](int j) {
			if (curv[j % (max_lag + 1)].get() <= 0.0)
				std::cerr << "PCG: negative curvature encountered!" << std::endl;
			current_res = std::sqrt(res_sq[j % (max_lag + 1)].get());
			if (diagnostic(x, current_res)) {
				info.iters = j + 1;
				info.status = stop::converged_user;
				return true;
			}
			if (current_res < terminate_tol) {
				info.iters = j + 1;
				info.status = stop::converged_rtol;
				return true;
			}
			return false;
		};

		int issued = 0;
		for (; issued < settings.maxiter && !done; ++issued) {
			const int slot = issued % (max_lag + 1);
			A.apply(p, w);
			curv[slot] = device::dot(w, p, &curvature); // p' A p, rides in the SpMV kernel
			device::linear_sum(x, device::ratio(1.0, *rho, curvature), p, device::number(1.0), x);
			device::linear_sum(r, device::ratio(-1.0, *rho, curvature), w, device::number(1.0), r);
			res_sq[slot] = device::dot(r, r, nullptr, FSB_HALT_IF_SQRT_LT, terminate_tol);
			P.apply(r, z);
			device::dot(r, z, rho_next);
			device::linear_sum(p, device::ratio(1.0, *rho_next, *rho), p, device::number(1.0), z); // p = beta p + z
			std::swap(rho, rho_next);
			if (issued >= lag)
				done = inspect(issued - lag);
		}
		for (int j = issued - lag < 0 ? 0 : issued - lag; j < issued && !done; ++j)
			done = inspect(j);
		halting.release();

		info.res_norm_final = current_res;
		info.sol_norm_final = x.l2norm().get();
		if (info.iters == 0)
			info.status = stop::diverged_iters;
		return info;
	}
};
template<class P>
cg_device(P) -> cg_device<P>;

}

namespace flecsolve::cg_device {
static constexpr std::size_t nwork = 4;
static inline work_factory<nwork> make_work;

template<class Work>
struct solver : krylov_solver<op::cg_device, settings, Work> {
	using base_t = krylov_solver<op::cg_device, settings, Work>;
	template<class W>
	solver(const settings & set, W && w) : base_t{set, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;
}
#endif

// Single-reduction preconditioned CG (Chronopoulos & Gear) with the Krylov scalars kept on the device
// ("single-reduction or pipelined CG").  No counterpart in the reference.
//
// Standard CG (solvers/cg.hh; reference flecsolve/solvers/cg.hh:91-131) has two dependent reductions per iteration
// (<p,Ap>, then <r,z>), each a cross-rank all-reduce the next kernel has to wait for.  Here the recurrence is
// rearranged so that ONE product and ONE point of synchronisation per iteration remain:
//
//     p = u + beta p ;  s = w + beta s ;  x += alpha p ;  r -= alpha s ;  |r|^2 ;  u = P r ;  gamma' = <r,u>     (1 kernel)
//     w = A u ;  delta = <w,u> ;  beta = gamma'/gamma ;  alpha = gamma' / (delta - beta gamma'/alpha) ;  gamma = gamma'
//                                                                                                          (1 kernel)
// with s = A p carried by recurrence instead of a second product.  The scalar statements ride on the reduction that
// completes delta (fsb_red_opts::post), so an iteration is TWO launches, nothing is read by the host, and the ghost
// exchange, the all-reduce and the scalar update all happen inside the SpMV kernel.  The price is one more vector
// (s) and 96 N instead of 88 N bytes of vector traffic per iteration; in exact arithmetic the iterates are CG's, in
// floating point they agree to rounding and the iteration count is CG's +- 1-2 (tests/test_device_scalar_gpu.py).
// As in cg_device.hh the kernel that finishes |r|^2 raises the halt flag when sqrt(.) < terminate_tol, every later
// vector update is then a no-op, and the host inspects residual norms `lag` iterations late.
#ifndef FLECSOLVE_B200_SOLVERS_CG_SR_HH
#define FLECSOLVE_B200_SOLVERS_CG_SR_HH

#include <array>
#include <cmath>

#include "flecsolve/device/scalar.hh"
#include "flecsolve/solvers/cg.hh"
#include "flecsolve/solvers/cg_device.hh"

namespace flecsolve::cg_sr {
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
struct cg_sr : base<Params, typename Params::input_var_t, typename Params::output_var_t> {
	using base_t = base<Params, typename Params::input_var_t, typename Params::output_var_t>;
	using base_t::params;

	cg_sr(Params p) : base_t(std::move(p)) {}

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
		auto & [r, u, w, p, s] = params.work;
		const int lag = settings.lag < 0 ? 0 : (settings.lag > max_lag ? max_lag : settings.lag);

		real terminate_tol, current_res;
		if (detail::krylov_start(A, settings, b, x, r, info, terminate_tol, current_res))
			return info;

		fsb_ctx_t ctx = x.data.ctx();
		device::scalar gamma(ctx), gamma_new(ctx), delta(ctx), alpha(ctx), beta(ctx), t(ctx);
		const fsb_coef one = device::number(1.0);
		// coefficient held in one slot: slot / 1.0
		auto plus = [](const device::scalar & v) { return fsb_coef{1.0, v.id(), 0}; };
		auto minus = [](const device::scalar & v) { return fsb_coef{-1.0, v.id(), 0}; };

		// u = P r, w = A u, gamma = <r,u>, delta = <w,u>, alpha = gamma / delta, beta = 0, p = s = 0
		P.apply(r, u);
		device::dot(r, u, &gamma);
		A.apply(u, w);
		device::scalar_program first;
		first.div(alpha, gamma, delta);
		device::dot(w, u, &delta, FSB_HALT_NEVER, 0, &first);
		beta.set(0.0);
		p.set_scalar(0.0);
		s.set_scalar(0.0);

		// beta = gamma'/gamma ; t = beta gamma' ; t = t / alpha ; t = delta - t ; alpha = gamma'/t ; gamma = gamma'
		device::scalar_program next;
		next.div(beta, gamma_new, gamma).mul(t, beta, gamma_new).div(t, t, alpha).sub(t, delta, t).div(alpha, gamma_new, t).copy(gamma, gamma_new);

		std::array<device_future, max_lag + 1> res_sq;
		device::halt_scope halting(ctx);
		bool done = false;
		auto inspect = [&](int j) { // the host's late look at iteration j; true ends the solve
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
			device::linear_sum(p, one, u, plus(beta), p); // p = u + beta p
			device::linear_sum(s, one, w, plus(beta), s); // s = w + beta s   (= A p)
			device::linear_sum(x, plus(alpha), p, one, x); // x += alpha p
			device::linear_sum(r, minus(alpha), s, one, r); // r -= alpha s
			res_sq[issued % (max_lag + 1)] = device::dot(r, r, nullptr, FSB_HALT_IF_SQRT_LT, terminate_tol);
			P.apply(r, u);
			device::dot(r, u, &gamma_new);
			A.apply(u, w);
			device::dot(w, u, &delta, FSB_HALT_NEVER, 0, &next); // rides in the SpMV kernel, scalar update included
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
cg_sr(P) -> cg_sr<P>;

}

namespace flecsolve::cg_sr {
static constexpr std::size_t nwork = 5;
static inline work_factory<nwork> make_work;

template<class Work>
struct solver : krylov_solver<op::cg_sr, settings, Work> {
	using base_t = krylov_solver<op::cg_sr, settings, Work>;
	template<class W>
	solver(const settings & set, W && w) : base_t{set, std::forward<W>(w)} {}
};
template<class W>
solver(const settings &, W &&) -> solver<std::decay_t<W>>;
}
#endif

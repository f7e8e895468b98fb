// Explicit embedded Runge-Kutta integrators: Bogacki-Shampine 3(2) (rk23) and Fehlberg 4(5) (rk45).
//
// Behavioural parity with flecsolve/time-integrators/rk23.hh:23-143 and rk45.hh:24-148: same stage
// combinations in the same order (so the iterates round identically), error estimate in `z`,
// acceptance on ||z||_2 < atol (or dt at its floor), the same step-size laws.  Every stage is a chain
// of aliased axpys onto one vector, which the C ABI's deferred queue turns into one kernel per stage.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_RK23_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_RK23_HH

#include <algorithm>
#include <cmath>

#include "flecsolve/solvers/solver_settings.hh"
#include "flecsolve/time-integrators/base.hh"

namespace flecsolve::time_integrator::rk23 {

struct settings : base_settings {
	float safety_factor = 0.9f;
	float atol = 1e-9f;
	bool use_fixed_dt = false;
};

template<class Op, class Work>
struct parameters : time_integrator::parameters<settings, Op, Work> {
	using base = time_integrator::parameters<settings, Op, Work>;
	template<class W>
	parameters(const settings & s, op::handle<Op> o, W && w) : base(s, o, std::forward<W>(w)) {}
};
template<class O, class W>
parameters(const settings &, op::handle<O>, W &&) -> parameters<O, W>;

enum workvecs : std::size_t { k1, k2, k3, k4, z, next, nvecs };

static inline work_factory<workvecs::nvecs> make_work;
template<std::size_t Version = 0>
using topo_work = topo_work_base<workvecs::nvecs, Version>;

namespace detail {
// shared by rk23 / rk45: accept when the embedded error estimate is below atol or dt is at its floor
template<class Z, class P>
bool accept(Z & z, const P & params, double current_dt) {
	const double err_est = z.l2norm().get();
	return (err_est < params.atol) || (std::fabs(current_dt - params.min_dt) < 1e-10);
}
}

template<class O, class W>
struct integrator : base<parameters<O, W>> {
	using P = parameters<O, W>;
	using base<P>::params;
	using base<P>::current_dt;
	using base<P>::current_time;
	using base<P>::assert_can_advance;

	integrator(P p) : base<P>(std::move(p)), total_step_rejects(0) {}

	template<class Curr, class Out>
	void advance(double dt, Curr & curr, Out & out) {
		assert_can_advance();
		current_dt = dt;
		auto & F = params.get_operator();
		auto & [k1, k2, k3, k4, z, next] = params.work;

		F.apply(curr, k1); // k1 = f(u_n)
		next.axpy(0.5 * dt, k1, curr);
		F.apply(next, k2); // k2 = f(u_n + dt/2 k1)
		next.axpy(0.75 * dt, k2, curr);
		F.apply(next, k3); // k3 = f(u_n + 3dt/4 k2)

		next.linear_sum(2.0, k1, 3.0, k2); // u_{n+1} = u_n + dt/9 (2 k1 + 3 k2 + 4 k3)
		next.axpy(4.0, k3, next);
		next.axpy(dt / 9.0, next, curr);
		F.apply(next, k4);

		z.linear_sum(-5., k1, 6., k2); // error estimate dt/72 (-5 k1 + 6 k2 + 8 k3 - 9 k4)
		z.axpy(8., k3, z);
		z.axpy(-9., k4, z);
		z.scale(dt / 72.);
		out.copy(next);
	}

	bool check_solution() { return detail::accept(std::get<workvecs::z>(params.work), params, current_dt); }

	double get_next_dt(bool good_solution) {
		if (params.use_fixed_dt)
			return std::min(current_dt, params.final_time - current_time);
		const double est_err = std::get<workvecs::z>(params.work).l2norm().get();
		double next_dt = params.safety_factor * current_dt * std::pow(params.atol / est_err, 1. / 3.);
		next_dt = std::min(std::max(next_dt, params.min_dt), params.max_dt);
		next_dt = std::min(next_dt, params.final_time - current_time);
		if (!good_solution)
			++total_step_rejects;
		return next_dt;
	}

protected:
	int total_step_rejects;
};
template<class O, class W>
integrator(parameters<O, W>) -> integrator<O, W>;

}
#endif

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
#include <utility>

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

// dst += c * k for every following (c, k) pair, one aliased axpy each -- the statement chain the
// reference writes out by hand, and the shape the deferred queue fuses into one kernel
template<class V>
void add_terms(V &) {}
template<class V, class K, class... More>
void add_terms(V & dst, double c, const K & k, More &&... more) {
	dst.axpy(c, k, dst);
	add_terms(dst, std::forward<More>(more)...);
}
// dst = c1 k1 + c2 k2 (+ c3 k3 ...): a linear_sum followed by the chain above
template<class V, class K1, class K2, class... More>
void weighted_sum(V & dst, double c1, const K1 & k1, double c2, const K2 & k2, More &&... more) {
	dst.linear_sum(c1, k1, c2, k2);
	add_terms(dst, std::forward<More>(more)...);
}

// the step-size law both pairs use: safety * dt * (atol / err)^(1/order), clipped to [min_dt, max_dt]
// and to what is left until final_time
template<class P>
double propose_dt(const P & params, double dt, double now, double err, double order) {
	double h = params.safety_factor * dt * std::pow(params.atol / err, 1. / order);
	h = std::min(std::max(h, params.min_dt), params.max_dt);
	return std::min(h, params.final_time - now);
}
}

// Bogacki-Shampine: three stages for the third-order update, a fourth (reusing the new state) for the
// embedded second-order comparison
template<class O, class W>
struct integrator : base<parameters<O, W>> {
	using P = parameters<O, W>;
	using base<P>::params;
	using base<P>::current_dt;
	using base<P>::current_time;
	using base<P>::assert_can_advance;

	integrator(P p) : base<P>(std::move(p)) {}

	template<class Curr, class Out>
	void advance(double dt, Curr & curr, Out & out) {
		assert_can_advance();
		current_dt = dt;
		auto & rhs = params.get_operator();
		auto & stage1 = std::get<workvecs::k1>(params.work);
		auto & stage2 = std::get<workvecs::k2>(params.work);
		auto & stage3 = std::get<workvecs::k3>(params.work);
		auto & stage4 = std::get<workvecs::k4>(params.work);
		auto & err = std::get<workvecs::z>(params.work);
		auto & state = std::get<workvecs::next>(params.work);

		rhs.apply(curr, stage1);
		state.axpy(0.5 * dt, stage1, curr); // u + dt/2 k1
		rhs.apply(state, stage2);
		state.axpy(0.75 * dt, stage2, curr); // u + 3dt/4 k2
		rhs.apply(state, stage3);

		detail::weighted_sum(state, 2.0, stage1, 3.0, stage2, 4.0, stage3);
		state.axpy(dt / 9.0, state, curr); // u + dt/9 (2 k1 + 3 k2 + 4 k3)
		rhs.apply(state, stage4);

		detail::weighted_sum(err, -5., stage1, 6., stage2, 8., stage3, -9., stage4);
		err.scale(dt / 72.); // dt/72 (-5 k1 + 6 k2 + 8 k3 - 9 k4)
		out.copy(state);
	}

	bool check_solution() { return detail::accept(std::get<workvecs::z>(params.work), params, current_dt); }

	double get_next_dt(bool good_solution) {
		if (params.use_fixed_dt)
			return std::min(current_dt, params.final_time - current_time);
		const double err = std::get<workvecs::z>(params.work).l2norm().get();
		const double h = detail::propose_dt(params, current_dt, current_time, err, 3.);
		if (!good_solution)
			++total_step_rejects;
		return h;
	}

protected:
	int total_step_rejects = 0;
};
template<class O, class W>
integrator(parameters<O, W>) -> integrator<O, W>;

}
#endif

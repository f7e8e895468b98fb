// Runge-Kutta-Fehlberg 4(5); see rk23.hh.  Reference: flecsolve/time-integrators/rk45.hh:24-148.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_RK45_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_RK45_HH

#include "flecsolve/time-integrators/rk23.hh"

namespace flecsolve::time_integrator::rk45 {

struct settings : rk23::settings {};

template<class Op, class Work>
struct parameters : rk23::parameters<Op, Work> {
	template<class W>
	parameters(const settings & s, op::handle<Op> o, W && w) : rk23::parameters<Op, Work>(s, o, std::forward<W>(w)) {}
};
template<class O, class W>
parameters(const settings &, op::handle<O>, W &&) -> parameters<O, W>;

enum workvecs : std::size_t { k1, k2, k3, k4, k5, k6, z, next, nvecs };

static inline work_factory<workvecs::nvecs> make_work;
template<std::size_t Version = 0>
using topo_work = topo_work_base<workvecs::nvecs, Version>;

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
		auto & [k1, k2, k3, k4, k5, k6, z, next] = params.work;
		// next = curr + first * a; next += rest...   (each line is one stage argument, left to right)
		auto stage = [&](auto & dst, double c1, auto & v1, std::initializer_list<std::pair<double, decltype(&k1)>> rest) {
			dst.axpy(c1, v1, curr);
			for (const auto & [c, v] : rest)
				dst.axpy(c, *v, dst);
		};

		F.apply(curr, k1);
		stage(next, 0.25 * dt, k1, {});
		F.apply(next, k2);
		stage(next, 3.0 * dt / 32., k1, {{9.0 * dt / 32., &k2}});
		F.apply(next, k3);
		stage(next, 1932. * dt / 2197., k1, {{-7200. * dt / 2197., &k2}, {7296. * dt / 2197., &k3}});
		F.apply(next, k4);
		stage(next, 439. * dt / 216., k1, {{-8. * dt, &k2}, {3680. * dt / 513., &k3}, {-845. * dt / 4104., &k4}});
		F.apply(next, k5);
		stage(next, -8. * dt / 27., k1,
		      {{2. * dt, &k2}, {-3544. * dt / 2565., &k3}, {1859. * dt / 4104., &k4}, {-11. * dt / 40., &k5}});
		F.apply(next, k6);

		// 4th-order solution in z, 5th-order in next; z becomes their difference
		stage(z, 25. * dt / 216., k1, {{1408. * dt / 2565., &k3}, {2197. * dt / 4104., &k4}, {-0.2 * dt, &k5}});
		stage(next, 16. * dt / 135., k1,
		      {{6656. * dt / 12825., &k3}, {28561. * dt / 56430., &k4}, {-9. * dt / 50., &k5}, {2. * dt / 55., &k6}});
		z.subtract(next, z);
		out.copy(next);
	}

	bool check_solution() { return rk23::detail::accept(std::get<workvecs::z>(params.work), params, current_dt); }

	double get_next_dt(bool good_solution) {
		if (params.use_fixed_dt)
			return std::min(current_dt, params.final_time - current_time);
		if (!good_solution) {
			++total_step_rejects;
			return params.safety_factor * current_dt;
		}
		const double err_est = std::get<workvecs::z>(params.work).l2norm().get();
		double next_dt = params.safety_factor * current_dt * std::pow((params.atol / err_est), 1. / 5.);
		next_dt = std::min(std::max(next_dt, params.min_dt), params.max_dt);
		return std::min(next_dt, params.final_time - current_time);
	}

protected:
	int total_step_rejects;
};
template<class O, class W>
integrator(parameters<O, W>) -> integrator<O, W>;

}
#endif

// REFCHECK-RK (test infrastructure): the REFERENCE's explicit integrators
// (flecsolve/time-integrators/rk23.hh, rk45.hh from /root/reference) on x' = lambda x with serial
// vectors, driven like time-integrators/test/explicit.cc:52-68.  Prints the step history as JSON.
//   refcheck_rk <23|45> <initial_dt> <max_dt> <min_dt> <final_time> <safety> <atol> <fixed 0|1> <lambda> <ic> <n>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "flecsolve/time-integrators/rk45.hh"
#include "flecsolve/vectors/seq.hh"

using namespace flecsolve;
using namespace flecsolve::time_integrator;

struct rate_params {
	double lambda;
};
struct rate : op::base<rate_params> {
	rate(double l) : op::base<rate_params>(rate_params{l}) {}
	template<class D, class R>
	void apply(const D & x, R & y) const {
		y.scale(params.lambda, x);
	}
};

template<class TI, class V>
static void drive(TI & ti, V & x, V & xnew, double ic, double lambda) {
	x.set_scalar(ic);
	auto dt = ti.get_current_dt();
	std::printf("{\"steps\": [");
	bool comma = false;
	while (ti.get_current_time() < ti.get_final_time()) {
		ti.advance(dt, x, xnew);
		const bool good = ti.check_solution();
		std::printf("%s[\"%a\", %d, \"%a\"]", comma ? ", " : "", dt, good ? 1 : 0, xnew.max().get());
		comma = true;
		if (good || ti.fixed_dt()) {
			ti.update();
			std::swap(x, xnew);
		}
		dt = ti.get_next_dt(good);
	}
	std::printf("], \"final_time\": \"%a\", \"error\": \"%a\", \"nsteps\": %d, \"value\": \"%a\"}\n", ti.get_current_time(),
	            std::fabs(ic * std::exp(lambda * ti.get_final_time()) - x.max().get()), ti.get_current_step(), x.max().get());
}

int main(int argc, char ** argv) {
	if (argc < 12)
		return 2;
	const int which = std::atoi(argv[1]);
	rk45::settings s{};
	s.initial_time = 0.0;
	s.initial_dt = std::atof(argv[2]);
	s.max_dt = std::atof(argv[3]);
	s.min_dt = std::atof(argv[4]);
	s.final_time = std::atof(argv[5]);
	s.safety_factor = std::strtof(argv[6], nullptr);
	s.atol = std::strtof(argv[7], nullptr);
	s.use_fixed_dt = std::atoi(argv[8]) != 0;
	s.max_steps = 1000;
	const double lambda = std::atof(argv[9]), ic = std::atof(argv[10]);
	const std::size_t n = std::atol(argv[11]);
	op::core<rate> F(lambda);
	vec::seq_vec<double> x{n}, xnew{n};
	if (which == 23) {
		std::array<vec::seq_vec<double>, rk23::workvecs::nvecs> work;
		for (auto & w : work)
			w.data.resize(n);
		rk23::integrator ti(rk23::parameters(static_cast<const rk23::settings &>(s), op::ref(F), std::move(work)));
		drive(ti, x, xnew, ic, lambda);
	}
	else {
		std::array<vec::seq_vec<double>, rk45::workvecs::nvecs> work;
		for (auto & w : work)
			w.data.resize(n);
		rk45::integrator ti(rk45::parameters(s, op::ref(F), std::move(work)));
		drive(ti, x, xnew, ic, lambda);
	}
	return 0;
}

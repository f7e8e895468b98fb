// operator_adapter<Op>: turns a right-hand-side operator F into (I - gamma F), the operator an
// implicit time step inverts.  Reference: flecsolve/time-integrators/operator_adapter.hh:23-52.
// The aliased axpy below is queued behind the SpMV of F by the C ABI.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_OPERATOR_ADAPTER_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_OPERATOR_ADAPTER_HH

#include <utility>

namespace flecsolve::time_integrator {

template<class Op>
struct operator_adapter : Op {
	template<class... Args>
	operator_adapter(Args &&... args) : Op(std::forward<Args>(args)...), gamma{1.} {}

	// y = x - gamma F(x)
	template<class D, class R>
	void apply(const D & x, R & y) const {
		apply_rhs(x, y);
		y.axpy(-gamma, y, x);
	}
	template<class D, class R>
	decltype(auto) apply_rhs(const D & x, R & y) const {
		return static_cast<const Op &>(*this).apply(x, y);
	}
	template<class V>
	bool is_valid(const V &) {
		return true;
	}
	double get_scaling() const { return gamma; }
	void set_scaling(double scaling) { gamma = scaling; }

protected:
	double gamma;
};

}
#endif

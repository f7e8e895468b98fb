// operator_adapter<Rhs>: the operator an implicit time step inverts,  x  ->  x - gamma * F(x),
// built on top of the user's right-hand-side operator F (which it inherits, so F's constructor
// arguments and extra members stay available).  Interface of the reference class
// (flecsolve/time-integrators/operator_adapter.hh:23-52): apply, apply_rhs, is_valid,
// get_scaling / set_scaling; the integrators push a new gamma whenever dt or the method changes.
// On the device the combination below is one queued statement behind F's SpMV.
#ifndef FLECSOLVE_B200_TIME_INTEGRATORS_OPERATOR_ADAPTER_HH
#define FLECSOLVE_B200_TIME_INTEGRATORS_OPERATOR_ADAPTER_HH

#include <utility>

namespace flecsolve::time_integrator {

template<class Rhs>
class operator_adapter : public Rhs {
	double scale_ = 1.0; // gamma

public:
	template<class... RhsArgs>
	operator_adapter(RhsArgs &&... rhs_args) : Rhs(std::forward<RhsArgs>(rhs_args)...) {}

	void set_scaling(double gamma) { scale_ = gamma; }
	double get_scaling() const { return scale_; }

	// out = F(in), untouched by the scaling
	template<class In, class Out>
	decltype(auto) apply_rhs(const In & in, Out & out) const {
		return Rhs::apply(in, out);
	}

	// out = in - gamma F(in); `out` first receives F(in) and is then combined in place
	// (same expression as the reference: (-gamma) * out + in)
	template<class In, class Out>
	void apply(const In & in, Out & out) const {
		this->apply_rhs(in, out);
		out.axpy(-scale_, out, in);
	}

	// every state is admissible for a linear problem
	template<class State>
	bool is_valid(const State &) {
		return true;
	}
};

}
#endif

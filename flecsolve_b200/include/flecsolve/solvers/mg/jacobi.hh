// Point-Jacobi preconditioners on a parallel CSR matrix.
//
//  * mg::bound_jacobi  -- weighted Jacobi relaxation, `nrelax` sweeps of
//        x <- omega D^-1 (b - (L+U) x_old) + (1 - omega) x_old
//    over the diag and offd blocks, x is NOT zeroed first (reference:
//    flecsolve/solvers/mg/jacobi.hh:30-121).  One C-ABI call per apply (fsb_parcsr_jacobi_relax).
//  * op::diagonal_inverse -- y = D^-1 x, the operator the reference's tests build as a diagonal
//    CSR matrix and apply by SpMV (csr_op_gen::Dinv(), flecsolve/util/test/mesh.hh:123-140).
//    Here it is an element-wise multiply by a stored 1/diag vector, which the deferred queue
//    fuses with the dot product that follows it in CG (cg.hh:124-127).
#ifndef FLECSOLVE_B200_SOLVERS_MG_JACOBI_HH
#define FLECSOLVE_B200_SOLVERS_MG_JACOBI_HH

#include "flecsolve/matrices/parcsr.hh"
#include "flecsolve/operators/handle.hh"

namespace flecsolve::mg {

struct jacobi_settings {
	float omega;
	std::size_t nrelax;
};

template<class scalar, class size>
struct bound_jacobi : op::base<> {
	using op_t = op::core<mat::parcsr<scalar, size>>;
	using topo_t = typename mat::parcsr<scalar, size>::topo_t;
	static inline const typename topo_t::template vec_def<topo_t::cols> tmpd;

	op::handle<op_t> A;
	jacobi_settings settings;

	bound_jacobi(op::handle<op_t> h, jacobi_settings s) : A{h}, settings{s} {}

	template<class D, class R>
	void apply(const D & b, R & x) const {
		auto tmp = vec::make(tmpd(x.data.topo()));
		device::check(fsb_parcsr_jacobi_relax(A.get().data.handle(), static_cast<scalar>(settings.omega),
		                                      static_cast<std::int64_t>(settings.nrelax), b.data.handle(), x.data.handle(),
		                                      tmp.data.handle()));
	}
};

struct jacobi {
	using settings = jacobi_settings;
	template<class scalar, class size>
	auto operator()(op::handle<op::core<mat::parcsr<scalar, size>>> A) {
		return op::core<bound_jacobi<scalar, size>>{A, settings_};
	}
	jacobi_settings settings_{2 / 3.f, 1};
};

}

namespace flecsolve::op {

template<class scalar, class size, class ivar = variable_t<anon_var::anonymous>, class ovar = variable_t<anon_var::anonymous>>
struct diagonal_inverse : base<std::nullptr_t, ivar, ovar> {
	using topo_t = typename mat::parcsr<scalar, size>::topo_t;
	using vec_t = decltype(vec::make(std::declval<typename topo_t::template vec_def<topo_t::cols> &>()(
		std::declval<typename topo_t::topology &>())));

private:
	// one definition per operator object (several matrices may each carry their own 1/diag);
	// declared before `dinv`, which is built from it
	typename topo_t::template vec_def<topo_t::cols> def_;

public:
	explicit diagonal_inverse(const op::core<mat::parcsr<scalar, size>> & A)
		: dinv(vec::make(def_(A.data.topo()))) {
		device::check(fsb_parcsr_extract_dinv(A.data.handle(), dinv.data.handle()));
	}

	template<class D, class R>
	void apply(const D & x, R & y) const {
		y.multiply(dinv, x);
	}

	vec_t dinv;
};

template<class scalar, class size>
auto make_diagonal_inverse(const op::core<mat::parcsr<scalar, size>> & A) {
	return core<diagonal_inverse<scalar, size>>(A);
}

}
#endif

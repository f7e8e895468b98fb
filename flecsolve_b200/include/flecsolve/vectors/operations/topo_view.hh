// Ops policy of vec::topo_view: every vector operation becomes one call of the C ABI.
//
// Reference: flecsolve/vectors/operations/topo_view.hh:43-291 dispatches on field-id equality
// to `_self' task variants (topo_tasks.hh:52-324).  The device kernels are alias-safe by
// construction (each element is read before it is written inside one thread), so the dispatch
// disappears; the C ABI queues the statement and fuses it with its neighbours (include/fsb.h).
// scale(alpha, x, y) keeps the reference's requirement that x and y are different fields.
#ifndef FLECSOLVE_B200_VECTORS_OPERATIONS_TOPO_VIEW_HH
#define FLECSOLVE_B200_VECTORS_OPERATIONS_TOPO_VIEW_HH

#include <cmath>
#include <string>
#include <string_view>

#include "flecsolve/util/future.hh"
#include "flecsolve/util/traits.hh"

namespace flecsolve::vec::ops {

template<class Data>
struct topo_view {
	using vec_data = Data;
	using scalar = typename Data::scalar;
	using real = typename num_traits<scalar>::real;
	using len_t = std::size_t;
	static_assert(std::is_same_v<scalar, double>, "the device back end stores fp64 vectors");

	using check_t = void (*)(int);
	static void ok(int rc) { device::check(rc); }

	template<class Other>
	static void copy(const Other & x, vec_data & z) {
		static_assert(std::is_same_v<typename Other::topo_t, typename vec_data::topo_t> && Other::space == vec_data::space,
		              "copy: vectors must live on the same index space");
		ok(fsb_vec_copy(z.handle(), x.handle()));
	}
	static void zero(vec_data & x) { ok(fsb_vec_set(x.handle(), 0.0)); }
	static void set_random(vec_data & x, unsigned seed) { ok(fsb_vec_set_random(x.handle(), seed)); }
	static void set_to_scalar(scalar a, vec_data & x) { ok(fsb_vec_set(x.handle(), a)); }
	static void scale(scalar a, vec_data & x) { ok(fsb_vec_scale(x.handle(), a, x.handle())); }
	static void scale(scalar a, const vec_data & x, vec_data & y) {
		if (x.fid() == y.fid())
			throw device::error(FSB_ERR_ARG, "scale operation: vector data cannot be the same");
		ok(fsb_vec_scale(y.handle(), a, x.handle()));
	}
	static void add(const vec_data & x, const vec_data & y, vec_data & z) { ok(fsb_vec_add(z.handle(), x.handle(), y.handle())); }
	static void subtract(const vec_data & x, const vec_data & y, vec_data & z) {
		ok(fsb_vec_sub(z.handle(), x.handle(), y.handle()));
	}
	static void multiply(const vec_data & x, const vec_data & y, vec_data & z) {
		ok(fsb_vec_mul(z.handle(), x.handle(), y.handle()));
	}
	static void divide(const vec_data & x, const vec_data & y, vec_data & z) {
		ok(fsb_vec_div(z.handle(), x.handle(), y.handle()));
	}
	static void reciprocal(const vec_data & x, vec_data & y) { ok(fsb_vec_recip(y.handle(), x.handle())); }
	// intended math z = alpha x + beta y for every alias pattern (the reference's y-alias branch
	// swaps its task arguments, operations/topo_view.hh:158-160; no caller reaches it)
	static void linear_sum(scalar alpha, const vec_data & x, scalar beta, const vec_data & y, vec_data & z) {
		ok(fsb_vec_linear_sum(z.handle(), alpha, x.handle(), beta, y.handle()));
	}
	static void axpy(scalar alpha, const vec_data & x, const vec_data & y, vec_data & z) {
		ok(fsb_vec_axpy(z.handle(), alpha, x.handle(), y.handle()));
	}
	static void axpby(scalar alpha, scalar beta, const vec_data & x, vec_data & z) {
		ok(fsb_vec_axpby(z.handle(), alpha, beta, x.handle()));
	}
	static void abs(const vec_data & x, vec_data & y) { ok(fsb_vec_abs(y.handle(), x.handle())); }
	static void add_scalar(const vec_data & x, scalar alpha, vec_data & y) {
		ok(fsb_vec_add_scalar(y.handle(), x.handle(), alpha));
	}

	// ---- reductions
	template<class Fn, class... A>
	static device_future queue(const vec_data & x, Fn fn, A... a) {
		device_future f{x.ctx(), 0};
		ok(fn(a..., &f.token));
		return f;
	}
	static auto min(const vec_data & x) { return queue(x, fsb_vec_min, x.handle()); }
	static auto max(const vec_data & x) { return queue(x, fsb_vec_max, x.handle()); }
	static auto inf_norm(const vec_data & x) { return queue(x, fsb_vec_amax, x.handle()); }
	static auto dot(const vec_data & x, const vec_data & y) { return queue(x, fsb_vec_dot, x.handle(), y.handle()); }

	// sum_i |x_i| (p=1), sum_i x_i^2 (p=2), sum_i pow(x_i, p) otherwise -- without the root
	template<unsigned short p>
	static auto lp_norm_local(const vec_data & x) {
		if constexpr (p == 1)
			return queue(x, fsb_vec_asum, x.handle());
		else if constexpr (p == 2)
			return queue(x, fsb_vec_sumsq, x.handle());
		else
			return queue(x, fsb_vec_powsum, x.handle(), static_cast<int>(p));
	}
	template<unsigned short p>
	static auto lp_norm(const vec_data & x) {
		auto fut = lp_norm_local<p>(x);
		if constexpr (p == 1)
			return fut;
		else if constexpr (p == 2)
			return future_transform{fut, [](double v) { return std::sqrt(v); }};
		else
			return future_transform{fut, [](double v) { return std::pow(v, 1. / p); }};
	}

	static auto global_size(const vec_data & x) {
		std::int64_t n = 0;
		ok(fsb_vec_global_size(x.handle(), &n));
		return ready_future<std::size_t>{static_cast<std::size_t>(n)};
	}
	static len_t local_size(const vec_data & x) { return static_cast<len_t>(fsb_vec_local_size(x.handle())); }
	static void dump(std::string_view pre, const vec_data & x) { ok(fsb_vec_dump(x.handle(), std::string(pre).c_str())); }

	template<class F, class... Vecs>
	static constexpr decltype(auto) apply(F && f, Vecs &&... vecs) {
		return std::forward<F>(f)(std::forward<Vecs>(vecs)...);
	}
};

}
#endif

// Result handles of global reductions.
//
// Contract kept from the reference (flecsolve/util/future.hh:24-56): anything a reduction
// returns has get() and wait(); composite futures forward both.  device_future wraps a token of
// the C ABI (include/fsb.h, fsb_red_get): get() is where queued kernels are launched and the
// host waits for the all-rank value -- the counterpart of flecsi::future<T>::get().
#ifndef FLECSOLVE_B200_UTIL_FUTURE_HH
#define FLECSOLVE_B200_UTIL_FUTURE_HH

#include <tuple>
#include <utility>

#include "flecsolve/device/runtime.hh"

namespace flecsolve {

struct device_future {
	fsb_ctx_t ctx = nullptr;
	fsb_token_t token = 0;

	double get() {
		double v = 0;
		device::check(fsb_red_get(ctx, token, &v));
		return v;
	}
	void wait() { device::check(fsb_red_wait(ctx, token)); }
};

// a value that is already on the host (sizes etc.)
template<class T>
struct ready_future {
	T value;
	T get() { return value; }
	void wait() {}
};

template<class... Futures>
struct future_vector {
	std::tuple<Futures...> futures;

	auto get() {
		return std::apply([](Futures &... f) { return std::make_tuple(f.get()...); }, futures);
	}
	void wait() {
		std::apply([](Futures &... f) { (f.wait(), ...); }, futures);
	}
};
template<class... Futures>
future_vector(std::tuple<Futures...>) -> future_vector<Futures...>;

template<class Future, class F>
struct future_transform {
	Future fut;
	F f;

	decltype(auto) get() { return f(fut.get()); }
	void wait() { fut.wait(); }
};
template<class Future, class F>
future_transform(Future, F) -> future_transform<Future, F>;

}
#endif

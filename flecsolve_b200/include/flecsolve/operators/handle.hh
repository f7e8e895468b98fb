// op::handle: how solvers hold operators -- shared ownership or a plain reference, never a copy.
// Reference: flecsolve/operators/handle.hh:11-56.
#ifndef FLECSOLVE_B200_OPERATORS_HANDLE_HH
#define FLECSOLVE_B200_OPERATORS_HANDLE_HH

#include <functional>
#include <memory>
#include <variant>

#include "flecsolve/operators/core.hh"
#include "flecsolve/util/traits.hh"

namespace flecsolve::op {

template<class T>
struct handle {
	using type = T;
	using var_t = std::variant<std::shared_ptr<type>, std::reference_wrapper<type>>;
	var_t store;

	constexpr type & get() const {
		if (auto * sp = std::get_if<std::shared_ptr<type>>(&store))
			return **sp;
		return std::get<std::reference_wrapper<type>>(store).get();
	}
	constexpr operator type &() const { return get(); }

	template<class D, class R>
	decltype(auto) operator()(const D & x, R & y) const {
		return get()(x, y);
	}
};

template<class T>
constexpr auto ref(T & o) {
	return handle<T>{std::ref(o)};
}
template<class T>
constexpr auto cref(const T & o) {
	return handle<const T>{std::ref(o)};
}

template<class T, std::enable_if_t<!is_operator_v<T>, bool> = false, class... Args>
constexpr auto make_shared(Args &&... args) {
	return handle<core<T>>{std::make_shared<core<T>>(std::forward<Args>(args)...)};
}
template<class T>
constexpr auto make_shared(core<T> && o) {
	return handle<core<T>>{std::make_shared<core<T>>(std::move(o))};
}

}
#endif

// op::handle<T>: how solvers, integrators and factories hold operators -- by shared ownership
// (op::make_shared<Policy>(args...)) or by plain reference to an object the caller keeps alive
// (op::ref(A), op::cref(A)) -- never by copy: an operator may own device matrices and work vectors.
// Interface of the reference class (flecsolve/operators/handle.hh:11-56): `type`, get(), conversion to
// T&, call forwarding, and the three factory functions.
#ifndef FLECSOLVE_B200_OPERATORS_HANDLE_HH
#define FLECSOLVE_B200_OPERATORS_HANDLE_HH

#include <memory>
#include <type_traits>
#include <utility>

#include "flecsolve/operators/core.hh"
#include "flecsolve/util/traits.hh"

namespace flecsolve::op {

template<class T>
class handle {
	std::shared_ptr<T> keep_alive_; // empty for borrowed operators
	T * target_ = nullptr;

public:
	using type = T;

	// borrow: the caller guarantees the operator outlives every solver bound to it
	explicit constexpr handle(T & borrowed) : target_(&borrowed) {}
	// share
	explicit handle(std::shared_ptr<T> owned) : keep_alive_(std::move(owned)), target_(keep_alive_.get()) {}

	constexpr T & get() const { return *target_; }
	constexpr operator T &() const { return *target_; }
	bool owns() const { return static_cast<bool>(keep_alive_); }

	// a handle to an operator can be applied like the operator
	template<class X, class Y>
	decltype(auto) operator()(const X & x, Y & y) const {
		return (*target_)(x, y);
	}
};

template<class T>
constexpr handle<T> ref(T & op) {
	return handle<T>(op);
}
template<class T>
constexpr handle<const T> cref(const T & op) {
	return handle<const T>(op);
}

// construct op::core<Policy> from the policy's constructor arguments, shared
template<class Policy, std::enable_if_t<!is_operator_v<Policy>, bool> = false, class... Args>
handle<core<Policy>> make_shared(Args &&... args) {
	return handle<core<Policy>>(std::make_shared<core<Policy>>(std::forward<Args>(args)...));
}
// move an existing operator into shared ownership
template<class Policy>
handle<core<Policy>> make_shared(core<Policy> && op) {
	return handle<core<Policy>>(std::make_shared<core<Policy>>(std::move(op)));
}

}
#endif

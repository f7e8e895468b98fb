// Compile-time variable tags selecting components of multi-vectors
// (same vocabulary as the reference: flecsolve/vectors/variable.hh:26-77).
#ifndef FLECSOLVE_B200_VECTORS_VARIABLE_HH
#define FLECSOLVE_B200_VECTORS_VARIABLE_HH

#include <cstddef>
#include <limits>
#include <type_traits>

namespace flecsolve {

enum class anon_var : std::size_t { anonymous = std::numeric_limits<std::size_t>::max() };

template<auto V>
struct variable_name {
	static constexpr const char * value = "";
};

template<auto V>
struct variable_t {
	static constexpr auto value = V;
	static constexpr const char * name = variable_name<V>::value;
};
template<auto A, auto B>
constexpr bool operator==(const variable_t<A> &, const variable_t<B> &) {
	return A == B;
}
template<auto A, auto B>
constexpr bool operator!=(const variable_t<A> &, const variable_t<B> &) {
	return !(A == B);
}
template<auto V>
inline variable_t<V> variable{};

template<class T>
struct is_variable : std::false_type {};
template<auto V>
struct is_variable<variable_t<V>> : std::true_type {};
template<class T>
inline constexpr bool is_variable_v = is_variable<T>::value;

template<auto... Vs>
struct multivariable_t {};
template<auto... Vs>
inline multivariable_t<Vs...> multivariable{};

template<auto... A, auto... B>
constexpr bool operator==(const multivariable_t<A...> &, const multivariable_t<B...> &) {
	return (... && (A == B));
}
template<auto... A, auto... B>
constexpr bool operator!=(const multivariable_t<A...> &, const multivariable_t<B...> &) {
	return (... || (A != B));
}

}
#endif

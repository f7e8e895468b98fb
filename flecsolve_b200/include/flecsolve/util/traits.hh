// Small type traits shared by the header layer.
// Mirrors the names flecsolve user code relies on (reference: flecsolve/util/traits.hh:25-103).
#ifndef FLECSOLVE_B200_UTIL_TRAITS_HH
#define FLECSOLVE_B200_UTIL_TRAITS_HH

#include <complex>
#include <functional>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>

namespace flecsolve {

template<class T>
struct is_reference_wrapper : std::false_type {};
template<class T>
struct is_reference_wrapper<std::reference_wrapper<T>> : std::true_type {};
template<class T>
inline constexpr bool is_reference_wrapper_v = is_reference_wrapper<T>::value;

template<class T>
struct is_smart_ptr : std::false_type {};
template<class T, class D>
struct is_smart_ptr<std::unique_ptr<T, D>> : std::true_type {};
template<class T>
struct is_smart_ptr<std::shared_ptr<T>> : std::true_type {};
template<class T>
inline constexpr bool is_smart_ptr_v = is_smart_ptr<T>::value;

template<class T>
struct traits {};

// scalar / real pairing; the device back end computes in real fp64 only (the reference itself
// disables complex vectors under CUDA, vectors/test/flecsi_vector.cc:398-400)
template<class T>
struct num_traits {
	using scalar = T;
	using real = T;
	static constexpr bool is_complex = false;
};
template<class T>
struct num_traits<std::complex<T>> {
	using scalar = std::complex<T>;
	using real = T;
	static constexpr bool is_complex = true;
};

template<class Derived>
struct with_derived {
	Derived & derived() { return static_cast<Derived &>(*this); }
	const Derived & derived() const { return static_cast<const Derived &>(*this); }
};

namespace detail {
template<class Struct, class = void, class... T>
struct list_initializable : std::false_type {};
template<class Struct, class... T>
struct list_initializable<Struct, std::void_t<decltype(Struct{std::declval<T>()...})>, T...> : std::true_type {};

template<class Struct, class... T>
struct is_copy_of : std::false_type {};
template<class Struct, class T>
struct is_copy_of<Struct, T> : std::is_same<std::decay_t<T>, Struct> {};
}

template<class Struct, class... T>
using is_direct_list_initializable = detail::list_initializable<Struct, void, T...>;
template<class Struct, class... T>
inline constexpr bool is_direct_list_initializable_v = is_direct_list_initializable<Struct, T...>::value;

// aggregate that can be brace-initialised from T..., excluding the copy case
template<class Struct, class... T>
using is_aggregate_initializable = std::conjunction<std::is_aggregate<Struct>,
                                                    is_direct_list_initializable<Struct, T...>,
                                                    std::negation<detail::is_copy_of<Struct, T...>>>;
template<class Struct, class... T>
inline constexpr bool is_aggregate_initializable_v = is_aggregate_initializable<Struct, T...>::value;

}
#endif

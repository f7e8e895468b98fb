// is_vector_v<T>: T is (derived from) some vec::core instantiation (reference: vectors/traits.hh:22-37)
#ifndef FLECSOLVE_B200_VECTORS_TRAITS_HH
#define FLECSOLVE_B200_VECTORS_TRAITS_HH

#include "flecsolve/vectors/core.hh"

namespace flecsolve {

template<class T>
struct is_vector : decltype(vec::detail::derives_from_core(std::declval<T>())) {};
template<class T>
inline constexpr bool is_vector_v = is_vector<T>::value;

}
#endif

// operator labels (reference: flecsolve/operators/traits.hh:25-33)
#ifndef FLECSOLVE_B200_OPERATORS_TRAITS_HH
#define FLECSOLVE_B200_OPERATORS_TRAITS_HH
namespace flecsolve::op {
enum class label { jacobian };
}
#endif

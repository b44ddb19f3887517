// Shim (test infrastructure only): no-op stand-in for ruy's profiler labels,
// which the reference constructs at every kernel entry. ruy is not vendored.
#ifndef LCE_B200_ORACLE_SHIM_RUY_PROFILER_H_
#define LCE_B200_ORACLE_SHIM_RUY_PROFILER_H_
namespace ruy {
namespace profiler {
class ScopeLabel {
 public:
  template <typename... Args>
  explicit ScopeLabel(Args...) {}
};
}  // namespace profiler
}  // namespace ruy
#endif

// builtin_ops.cc -- float TFLite builtins the three model families need around the
// binary path. Placeholder registration point; the ops are added as
// the graph host widens.
#include "host_graph.h"

namespace lce_b200 {
void RegisterBuiltinOps(OpResolver*) {}
}  // namespace lce_b200

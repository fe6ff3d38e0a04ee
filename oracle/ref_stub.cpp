// TEST INFRASTRUCTURE (oracle/_ref build only).
// The reference's kernels/cpu/utils.cpp provides init_cpu_threads_env() on top of libnuma,
// which this image does not have. kernels/cpu/torch_bindings.cpp references the symbol, so the
// oracle build links this stub instead; thread binding is irrelevant to numerics.
#include <string>

std::string init_cpu_threads_env(const std::string& cpu_ids) {
  return "oracle stub: thread binding not available (" + cpu_ids + ")";
}

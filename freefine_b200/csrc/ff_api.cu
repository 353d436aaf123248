// Library-level entry points: version, thread-local last-error string.
#include "ff_common.cuh"

namespace ff {
char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace ff

extern "C" int ff_version(void) { return FF_VERSION; }
extern "C" const char* ff_last_error(void) { return ff::err_buf(); }

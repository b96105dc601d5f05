// Library-wide state of the C ABI: thread-local error text, launch counter, version.
#include "pvrl_host.h"

namespace pvrl {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}
}  // namespace pvrl

extern "C" const char* pvrl_last_error(void) { return pvrl::last_error_buf(); }
extern "C" int pvrl_abi_version(void) { return PVRL_ABI_VERSION; }
extern "C" int64_t pvrl_launch_count(void) { return pvrl::launch_counter().load(); }

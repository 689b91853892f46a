// Library-wide C-ABI pieces: version, thread-local error string, launch counter.
#include <atomic>

#include "ay2_common.h"

namespace ay2 {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace ay2

#ifndef AY2_SOURCE_HASH_STR
#define AY2_SOURCE_HASH_STR "0000000000000000000000000000000000000000000000000000000000000000"
#endif
// sha256 of the sources + flags this binary was built from (ayolov2_b200/_build.py): the loader refuses a stale binary
static const char g_source_hash[] = "AY2_SOURCE_HASH=" AY2_SOURCE_HASH_STR;

extern "C" int ay2_version(void) { return 200; }
extern "C" const char* ay2_source_hash(void) { return g_source_hash + 16; }
extern "C" const char* ay2_last_error_string(void) { return ay2::g_err; }
extern "C" int64_t ay2_launch_count(void) { return ay2::g_launches.load(std::memory_order_relaxed); }

// libpifpaf_b200: error slot, version, launch counter.
#include "common.cuh"

#include <atomic>
#include <cstdarg>

namespace pifpaf {

std::string& last_error_slot() {
    static thread_local std::string slot;
    return slot;
}

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_slot() = buf;
}

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace pifpaf

extern "C" {

const char* pifpaf_last_error(void) { return pifpaf::last_error_slot().c_str(); }
int pifpaf_abi_version(void) { return 1; }
const char* pifpaf_build_arch(void) { return "sm_100a"; }
int64_t pifpaf_launch_count(void) { return pifpaf::g_launches.load(); }

}

// Library-level entry points: ABI version and thread-local error text.
#include "common.cuh"
#include "../../include/t2s_b200.h"
#include <stdarg.h>
#include <stdlib.h>

namespace t2s {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("T2S_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}
}  // namespace t2s

extern "C" int t2s_abi_version(void) { return T2S_ABI_VERSION; }
extern "C" const char* t2s_last_error(void) { return t2s::g_err; }

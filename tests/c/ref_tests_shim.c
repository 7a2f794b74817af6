/* TEST INFRASTRUCTURE — the two internal symbols the reference's own unit tests
 * (src/tests/{filters,tone_mapping,colorspace}.c, compiled in place from /root/reference by
 * tests/test_reference_unit_tests.py) use besides the public API: pl_msg, the internal logging
 * entry point behind the PL_* macros of src/log.h. Everything else those tests call is resolved
 * by libplacebo_b200.so. */
#include <stdarg.h>
#include <stdio.h>

#include <libplacebo/log.h>

void pl_msg(pl_log log, enum pl_log_level lev, const char *fmt, ...)
{
    if (!log || !log->params.log_cb || lev > log->params.log_level)
        return;
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    log->params.log_cb(log->params.log_priv, lev, buf);
}

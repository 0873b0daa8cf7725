// common.h — status/error plumbing shared by every translation unit of libcustos_b200.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/custos_b200.h"

namespace cb {

// thread-local error text behind cb_last_error()
void set_error(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
const char *last_error();

inline int32_t fail(int32_t code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
inline int32_t fail(int32_t code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    set_error("%s", buf);
    return code;
}

inline bool is_float_dtype(int32_t dt) { return dt == CB_F32 || dt == CB_F64 || dt == CB_F16 || dt == CB_BF16; }
inline bool is_half_dtype(int32_t dt) { return dt == CB_F16 || dt == CB_BF16; }  // 16-bit storage, f32 arithmetic
inline bool is_signed_int_dtype(int32_t dt) { return dt == CB_I32 || dt == CB_I64 || dt == CB_I8 || dt == CB_I16; }
inline bool valid_dtype(int32_t dt) { return dt >= 0 && dt < CB_DTYPE_COUNT; }
size_t dtype_size(int32_t dt);
const char *dtype_name(int32_t dt);

}  // namespace cb

#define CB_CHECK_ARG(cond, msg)                                              \
    do {                                                                     \
        if (!(cond)) return ::cb::fail(CB_ERR_INVALID_ARG, "%s: %s", __func__, msg); \
    } while (0)

// propagate a non-zero status
#define CB_TRY(expr)                 \
    do {                             \
        int32_t _rc = (expr);        \
        if (_rc != CB_OK) return _rc; \
    } while (0)

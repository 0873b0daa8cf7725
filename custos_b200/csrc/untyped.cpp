// untyped.cpp — run-time typed buffer views and serde of device buffers.
//
// Reference behaviour mirrored here:
//  * src/devices/untyped/mod.rs:17-84 — `to_typed` / `as_typed` / `read_typed` succeed only when the
//    requested type is the storage's type (`matches_storage_type`, storages.rs); the accepted types are
//    the `AsType` impls (matches_type.rs:28-71);
//  * src/devices/cuda/cuda_ptr.rs:122-157 — `Serialize for CUDAPtr<T>` reads the buffer back and
//    serialises a sequence of T; `Deserialize` collects a sequence, allocates and writes it.
// The number formatting restates what the serde back ends do: serde_json prints integers in decimal and
// floats through ryu (shortest digits that round-trip, ryu's `pretty` layout rules); bincode 1.x writes a
// u64 length and fixed-width little-endian elements.  Neither crate is in the tree (Cargo.toml:45-46, no
// lock file): parity of the text is pinned only by the reference's token test (cuda_ptr.rs:170-190:
// a sequence of ten i32) and by round trips.
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"

using namespace cb;

extern "C" int32_t cbm_untyped_supports(int32_t dtype)
{
    switch (dtype) {
    case CB_U8: case CB_U32: case CB_I64: case CB_BF16: case CB_F16: case CB_F32: case CB_F64: return 1;
    default: return 0;
    }
}

extern "C" int32_t cbm_buffer_matches_type(cbm_device *d, cbm_buf b, int32_t dtype)
{
    int32_t have = -1;
    CB_TRY(cbm_buffer_dtype(d, b, &have));
    if (!valid_dtype(dtype)) return fail(CB_ERR_INVALID_ARG, "invalid dtype %d", dtype);
    if (have != dtype)
        return fail(CB_ERR_TYPE_MISMATCH, "storage type is %s, requested %s", dtype_name(have), dtype_name(dtype));
    return CB_OK;
}

extern "C" int32_t cbm_buffer_read_typed(cbm_device *d, cbm_buf b, int32_t dtype, void *host_out, size_t len)
{
    CB_TRY(cbm_buffer_matches_type(d, b, dtype));
    return cbm_buffer_read(d, b, host_out, len);
}

// ------------------------------------------------------------------ number <-> text
namespace {

bool serde_supported(int32_t dtype) { return valid_dtype(dtype) && dtype != CB_F16 && dtype != CB_BF16; }

// ryu::pretty::format32 / format64: the shortest decimal digits d1 d2 .. dn and exponent k with
// value = digits * 10^k; kk = n + k.  Layout: integers below 10^16 (10^13 for f32) as "1234000.0",
// a decimal point inside the digits when 0 < kk, "0.00ddd" down to kk = -4 (f64) / -5 (f32), otherwise
// scientific "d.ddde<exp>" with a bare "de<exp>" for one digit.
template <typename F>
void put_float(std::string &out, F v)
{
    if (!std::isfinite(v)) {  // serde_json: NaN and infinities become null
        out += "null";
        return;
    }
    if (std::signbit(v)) out.push_back('-');
    const F a = std::fabs(v);
    if (a == 0) {
        out += "0.0";
        return;
    }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf - 1, a, std::chars_format::scientific);  // shortest round-trip digits
    *r.ptr = 0;
    std::string digits;
    const char *p = buf;
    for (; p < r.ptr && *p != 'e'; p++)
        if (*p != '.') digits.push_back(*p);
    const int exp10 = std::atoi(p + 1);  // value = d.ddd * 10^exp10
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int n = (int)digits.size();
    const int kk = exp10 + 1;
    const int k = kk - n;
    const int int_limit = sizeof(F) == 8 ? 16 : 13;
    const int small_limit = sizeof(F) == 8 ? -5 : -6;
    if (0 <= k && kk <= int_limit) {
        out += digits;
        out.append((size_t)k, '0');
        out += ".0";
    } else if (0 < kk && kk <= int_limit) {
        out.append(digits, 0, (size_t)kk);
        out.push_back('.');
        out.append(digits, (size_t)kk, std::string::npos);
    } else if (small_limit < kk && kk <= 0) {
        out += "0.";
        out.append((size_t)(-kk), '0');
        out += digits;
    } else {
        out.push_back(digits[0]);
        if (n > 1) {
            out.push_back('.');
            out.append(digits, 1, std::string::npos);
        }
        out.push_back('e');
        out += std::to_string(kk - 1);
    }
}

template <typename I>
void put_int(std::string &out, I v)
{
    char buf[32];
    auto r = std::to_chars(buf, buf + sizeof buf, v);
    out.append(buf, r.ptr);
}

void put_element(std::string &out, int32_t dtype, const unsigned char *p)
{
    switch (dtype) {
    case CB_F32: { float v; std::memcpy(&v, p, 4); put_float(out, v); } break;
    case CB_F64: { double v; std::memcpy(&v, p, 8); put_float(out, v); } break;
    case CB_I8: put_int(out, (int)*(const int8_t *)p); break;
    case CB_U8: put_int(out, (unsigned)*p); break;
    case CB_I16: { int16_t v; std::memcpy(&v, p, 2); put_int(out, (int)v); } break;
    case CB_U16: { uint16_t v; std::memcpy(&v, p, 2); put_int(out, (unsigned)v); } break;
    case CB_I32: { int32_t v; std::memcpy(&v, p, 4); put_int(out, v); } break;
    case CB_U32: { uint32_t v; std::memcpy(&v, p, 4); put_int(out, v); } break;
    case CB_I64: { int64_t v; std::memcpy(&v, p, 8); put_int(out, (long long)v); } break;
    case CB_U64: { uint64_t v; std::memcpy(&v, p, 8); put_int(out, (unsigned long long)v); } break;
    default: out += *p ? "true" : "false"; break;  // CB_BOOL
    }
}

struct Parser {
    const char *p, *end;
    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool eat(char c)
    {
        ws();
        if (p < end && *p == c) {
            p++;
            return true;
        }
        return false;
    }
};

// the JSON number grammar: -?(0|[1-9][0-9]*)(\.[0-9]+)?([eE][+-]?[0-9]+)?
bool json_number(const std::string &t)
{
    size_t i = 0;
    const size_t n = t.size();
    auto digits = [&]() {
        const size_t start = i;
        while (i < n && t[i] >= '0' && t[i] <= '9') i++;
        return i - start;
    };
    if (i < n && t[i] == '-') i++;
    if (i < n && t[i] == '0') i++;
    else if (!digits()) return false;
    if (i < n && t[i] == '.') {
        i++;
        if (!digits()) return false;
    }
    if (i < n && (t[i] == 'e' || t[i] == 'E')) {
        i++;
        if (i < n && (t[i] == '+' || t[i] == '-')) i++;
        if (!digits()) return false;
    }
    return i == n;
}

// one JSON scalar into the element at `dst`; false on a token that serde_json would reject for this type
bool parse_element(Parser &ps, int32_t dtype, unsigned char *dst)
{
    ps.ws();
    if (ps.p >= ps.end) return false;
    if (dtype == CB_BOOL) {
        if (ps.end - ps.p >= 4 && !std::memcmp(ps.p, "true", 4)) { *dst = 1; ps.p += 4; return true; }
        if (ps.end - ps.p >= 5 && !std::memcmp(ps.p, "false", 5)) { *dst = 0; ps.p += 5; return true; }
        return false;
    }
    const char *tok = ps.p;
    const char *q = tok;
    auto number_char = [](char ch) { return (ch >= '0' && ch <= '9') || ch == '+' || ch == '-' || ch == '.' || ch == 'e' || ch == 'E'; };
    while (q < ps.end && number_char(*q)) q++;
    if (q == tok) return false;  // includes `null`: not a number
    const std::string text(tok, q);
    if (!json_number(text)) return false;  // "+1", "01", "1.", ".5", "1e" are not JSON
    ps.p = q;
    char *stop = nullptr;
    errno = 0;
    if (dtype == CB_F32 || dtype == CB_F64) {
        const double v = std::strtod(text.c_str(), &stop);  // serde_json parses to f64; f32 is `as f32`
        if (*stop || (errno == ERANGE && std::isinf(v))) return false;  // serde_json: "number out of range"
        if (dtype == CB_F64) std::memcpy(dst, &v, 8);
        else { const float f = (float)v; std::memcpy(dst, &f, 4); }
        return true;
    }
    if (text.find_first_of(".eE") != std::string::npos) return false;  // a float token for an integer type
    const bool is_unsigned = dtype == CB_U8 || dtype == CB_U16 || dtype == CB_U32 || dtype == CB_U64;
    if (is_unsigned) {
        if (text[0] == '-') return false;
        const unsigned long long v = std::strtoull(text.c_str(), &stop, 10);
        if (*stop || errno == ERANGE) return false;
        const size_t sz = dtype_size(dtype);
        if (sz < 8 && (v >> (8 * sz))) return false;  // out of range for the type
        std::memcpy(dst, &v, sz);  // little endian host
        return true;
    }
    const long long v = std::strtoll(text.c_str(), &stop, 10);
    if (*stop || errno == ERANGE) return false;
    const size_t sz = dtype_size(dtype);
    if (sz < 8) {
        const long long lim = 1ll << (8 * sz - 1);
        if (v < -lim || v >= lim) return false;
    }
    std::memcpy(dst, &v, sz);
    return true;
}

}  // namespace

// ------------------------------------------------------------------ host-side codec (no device needed)
extern "C" int32_t cb_serde_encode(int32_t dtype, int32_t format, const void *elems, size_t n, void *out, size_t cap,
                                   size_t *needed)
{
    CB_CHECK_ARG(needed && (elems || !n), "null argument");
    CB_CHECK_ARG(format == CB_SER_JSON || format == CB_SER_BINCODE, "unknown format");
    if (!serde_supported(dtype))
        return fail(CB_ERR_UNSUPPORTED, "%s does not implement Serialize in the reference build",
                    valid_dtype(dtype) ? dtype_name(dtype) : "this dtype");
    const size_t sz = dtype_size(dtype);
    if (format == CB_SER_BINCODE) {
        *needed = 8 + n * sz;
        if (!out || cap < *needed) return CB_OK;
        const uint64_t n64 = (uint64_t)n;
        std::memcpy(out, &n64, 8);
        if (n) std::memcpy((unsigned char *)out + 8, elems, n * sz);
        return CB_OK;
    }
    std::string text;
    text.reserve(n * 8 + 2);
    text.push_back('[');
    for (size_t i = 0; i < n; i++) {
        if (i) text.push_back(',');
        put_element(text, dtype, (const unsigned char *)elems + i * sz);
    }
    text.push_back(']');
    *needed = text.size();
    if (out && cap >= text.size()) std::memcpy(out, text.data(), text.size());
    return CB_OK;
}

static int32_t decode_to(int32_t dtype, int32_t format, const void *in, size_t len, std::vector<unsigned char> *host,
                         size_t *count)
{
    CB_CHECK_ARG(in, "null argument");
    CB_CHECK_ARG(format == CB_SER_JSON || format == CB_SER_BINCODE, "unknown format");
    if (!serde_supported(dtype))
        return fail(CB_ERR_UNSUPPORTED, "%s does not implement Deserialize in the reference build",
                    valid_dtype(dtype) ? dtype_name(dtype) : "this dtype");
    const size_t sz = dtype_size(dtype);
    host->clear();
    *count = 0;
    if (format == CB_SER_BINCODE) {
        if (len < 8) return fail(CB_ERR_PARSE, "bincode: %zu bytes, no length prefix", len);
        uint64_t n64;
        std::memcpy(&n64, in, 8);
        if (n64 > (len - 8) / sz || (size_t)n64 * sz != len - 8)
            return fail(CB_ERR_PARSE, "bincode: length prefix %llu does not match %zu payload bytes of %s",
                        (unsigned long long)n64, len - 8, dtype_name(dtype));
        const unsigned char *payload = (const unsigned char *)in + 8;
        if (dtype == CB_BOOL)
            for (size_t i = 0; i < (size_t)n64; i++)
                if (payload[i] > 1) return fail(CB_ERR_PARSE, "bincode: invalid bool byte at element %zu", i);
        host->assign(payload, payload + (size_t)n64 * sz);
        *count = (size_t)n64;
        return CB_OK;
    }
    Parser ps{(const char *)in, (const char *)in + len};
    if (!ps.eat('[')) return fail(CB_ERR_PARSE, "json: expected '['");
    size_t n = 0;
    if (!ps.eat(']')) {
        for (;;) {
            host->resize((n + 1) * sz);
            if (!parse_element(ps, dtype, host->data() + n * sz))
                return fail(CB_ERR_PARSE, "json: element %zu is not a valid %s", n, dtype_name(dtype));
            n++;
            if (ps.eat(',')) continue;
            if (ps.eat(']')) break;
            return fail(CB_ERR_PARSE, "json: expected ',' or ']' after element %zu", n - 1);
        }
    }
    ps.ws();
    if (ps.p != ps.end) return fail(CB_ERR_PARSE, "json: trailing characters");
    *count = n;
    return CB_OK;
}

extern "C" int32_t cb_serde_decode(int32_t dtype, int32_t format, const void *in, size_t len, void *elems_out,
                                   size_t cap_elems, size_t *n)
{
    CB_CHECK_ARG(n, "null argument");
    std::vector<unsigned char> host;
    CB_TRY(decode_to(dtype, format, in, len, &host, n));
    if (elems_out && cap_elems >= *n && *n) std::memcpy(elems_out, host.data(), host.size());
    return CB_OK;
}

// ------------------------------------------------------------------ device buffers
extern "C" int32_t cbm_buffer_serialize(cbm_device *d, cbm_buf b, int32_t format, void *out, size_t cap, size_t *needed)
{
    CB_CHECK_ARG(d && needed, "null argument");
    CB_CHECK_ARG(format == CB_SER_JSON || format == CB_SER_BINCODE, "unknown format");
    int32_t dtype = -1;
    size_t len = 0;
    CB_TRY(cbm_buffer_dtype(d, b, &dtype));
    CB_TRY(cbm_buffer_len(d, b, &len));
    if (!serde_supported(dtype))
        return fail(CB_ERR_UNSUPPORTED, "%s does not implement Serialize in the reference build", dtype_name(dtype));
    const size_t sz = dtype_size(dtype);
    if (format == CB_SER_BINCODE) {  // the payload IS the device bytes: D2H straight into the caller's buffer
        *needed = 8 + len * sz;
        if (!out || cap < *needed) return CB_OK;
        const uint64_t n64 = (uint64_t)len;
        std::memcpy(out, &n64, 8);
        return len ? cbm_buffer_read(d, b, (unsigned char *)out + 8, len) : CB_OK;
    }
    std::vector<unsigned char> host(len * sz);  // cu_read into a host copy, then Serialize (cuda_ptr.rs:131-139)
    if (len) CB_TRY(cbm_buffer_read(d, b, host.data(), len));
    return cb_serde_encode(dtype, format, host.data(), len, out, cap, needed);
}

extern "C" int32_t cbm_buffer_deserialize(cbm_device *d, int32_t dtype, int32_t format, const void *in, size_t len,
                                          cbm_buf *out)
{
    CB_CHECK_ARG(d && out, "null argument");
    std::vector<unsigned char> host;
    size_t n = 0;
    CB_TRY(decode_to(dtype, format, in, len, &host, &n));
    // CUDAPtr::new(len) + cu_write (cuda_ptr.rs:151-153); a zero-length sequence fails like every
    // zero-length allocation
    if (!n) return fail(CB_ERR_ZERO_LENGTH, "deserialised an empty sequence: a zero length buffer cannot be allocated");
    return cbm_buffer_from_host(d, dtype, host.data(), n, out);
}

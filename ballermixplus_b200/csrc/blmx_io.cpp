// blmx_io.cpp -- host-side text I/O of the scan at genome scale.
//
// Plain C++17, no CUDA: built by g++ into libblmx_io.so, declared in include/blmx_io.h.
//
//   blmx_io_count_rows / blmx_io_read_sites
//       the 4-column input table of InputData.readCounts / readPolyCalls
//       (/root/reference/BalLeRMix+_v1.py:80-131): header line skipped, every other line
//       `l.strip().split('\t')`, physPos = int(float(c0)), k = int(c2), n = int(c3),
//       genPos = float(c0)*Rrate (+0) with --usePhysPos else float(c1)          (v1:103,124).
//       strtod is correctly rounded like Python's float(); anything it or strtoll does not
//       consume completely (underscores, stray text, missing columns) makes the call fail
//       with BLMX_IO_ERR_FORMAT and the Python reader, which is the reference semantics by
//       construction, takes over -- so accepted files parse identically, bit for bit.
//   blmx_io_write_rows
//       the output rows of Scan._alpha/_siteBased/_fixSize_siteCenter (v1:574,591,607):
//       f'{physPos}\t{genPos}\t{T}\t{x}\t{a}\t{A}\t{nSites}'.  physPos is an integer; genPos
//       and T are numpy float64 printed by str(), i.e. the shortest digits that round-trip,
//       positional for 1e-4 <= |v| < 1e16 and d.ddde+XX otherwise (the same rule as Python's
//       repr); the grid values arrive pre-formatted by Python because their int/float type
//       is part of the output.  Rows with iA < 0 are the reference's all-zero
//       row `0.0 0.0 0.0 0.0 0.0` (v1:451).
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "blmx_io.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

struct FileBuf {
    std::vector<char> data;
    int load(const char *path) {
        FILE *fh = std::fopen(path, "rb");
        if (!fh) return fail(BLMX_IO_ERR_OPEN, std::string("cannot open ") + path + ": " + std::strerror(errno));
        std::fseek(fh, 0, SEEK_END);
        long sz = std::ftell(fh);
        std::fseek(fh, 0, SEEK_SET);
        data.resize((size_t)sz + 1);
        size_t got = sz > 0 ? std::fread(data.data(), 1, (size_t)sz, fh) : 0;
        std::fclose(fh);
        if (got != (size_t)sz) return fail(BLMX_IO_ERR_OPEN, std::string("short read on ") + path);
        data[(size_t)sz] = '\0';
        return 0;
    }
};

// One line [b, e) -> up to 8 tab-separated fields of the stripped line.
int split_stripped(const char *b, const char *e, const char *fb[8], const char *fe[8]) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    int n = 0;
    const char *s = b;
    for (const char *p = b; p <= e; ++p) {
        if (p == e || *p == '\t') {
            if (n < 8) { fb[n] = s; fe[n] = p; }
            ++n;
            s = p + 1;
        }
    }
    return n;
}

// Python float(): optional surrounding whitespace, then a full strtod match.
bool parse_float(const char *b, const char *e, double *out) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b == e || e - b > 120) return false;
    char tmp[128];
    std::memcpy(tmp, b, (size_t)(e - b));
    tmp[e - b] = '\0';
    for (const char *p = tmp; *p; ++p)      // strtod extras Python rejects: hex floats, "infinity(...)", nan(...)
        if (*p == 'x' || *p == 'X' || *p == '(' || *p == '_') return false;
    char *end = nullptr;
    errno = 0;
    double v = std::strtod(tmp, &end);
    if (end != tmp + (e - b)) return false;
    *out = v;
    return true;
}

// Python int() of a plain decimal literal (optional sign, digits, surrounding whitespace).
bool parse_int(const char *b, const char *e, int64_t *out) {
    while (b < e && is_space(*b)) ++b;
    while (e > b && is_space(e[-1])) --e;
    if (b == e) return false;
    const char *p = b;
    bool neg = false;
    if (*p == '+' || *p == '-') { neg = *p == '-'; ++p; }
    if (p == e || e - p > 18) return false;
    int64_t v = 0;
    for (; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
    }
    *out = neg ? -v : v;
    return true;
}

// str(numpy.float64) / repr(float): shortest round-trip digits, positional for decimal point
// positions in (-4, 16], else scientific with a two-digit (at least) exponent.
char *format_double(char *out, double v) {
    if (std::isnan(v)) { std::memcpy(out, "nan", 3); return out + 3; }
    if (std::isinf(v)) {
        if (v < 0) *out++ = '-';
        std::memcpy(out, "inf", 3);
        return out + 3;
    }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof(sci) - 1, v, std::chars_format::scientific);
    *r.ptr = '\0';
    // sci = [-]d[.ddd]e[+-]XX
    const char *p = sci;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24] = {0};
    int nd = 0;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    int ex = (int)std::strtol(p + 1, nullptr, 10);
    if (nd == 1 && digits[0] == '0') { std::memcpy(out, "0.0", 3); return out + 3; }
    const int decpt = ex + 1;                         // digits * 10^(decpt - nd)
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *out++ = '0'; *out++ = '.';
            for (int i = 0; i < -decpt; ++i) *out++ = '0';
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
        } else if (decpt >= nd) {
            for (int i = 0; i < nd; ++i) *out++ = digits[i];
            for (int i = nd; i < decpt; ++i) *out++ = '0';
            *out++ = '.'; *out++ = '0';
        } else {
            for (int i = 0; i < decpt; ++i) *out++ = digits[i];
            *out++ = '.';
            for (int i = decpt; i < nd; ++i) *out++ = digits[i];
        }
    } else {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            for (int i = 1; i < nd; ++i) *out++ = digits[i];
        }
        *out++ = 'e';
        *out++ = ex < 0 ? '-' : '+';
        int ax = ex < 0 ? -ex : ex;
        char eb[8];
        int ne = 0;
        do { eb[ne++] = (char)('0' + ax % 10); ax /= 10; } while (ax);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *out++ = eb[--ne];
    }
    return out;
}

char *format_int(char *out, int64_t v) {
    auto r = std::to_chars(out, out + 24, v);
    return r.ptr;
}

}  // namespace

extern "C" {

const char *blmx_io_last_error(void) { return g_err.c_str(); }

int blmx_io_count_rows(const char *path, int64_t *n_rows) {
    if (!path || !n_rows) return fail(BLMX_IO_ERR_ARG, "blmx_io_count_rows: null pointer");
    FileBuf f;
    if (int rc = f.load(path)) return rc;
    const char *p = f.data.data(), *end = p + f.data.size() - 1;
    int64_t lines = 0;
    while (p < end) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        ++lines;
        if (!nl) break;
        p = nl + 1;
    }
    *n_rows = lines > 0 ? lines - 1 : 0;               // the header line is always skipped (v1:83,116)
    return 0;
}

int blmx_io_read_sites(const char *path, int64_t n_rows, int use_phys, double rrate, int strict_columns,
                       int64_t *position, double *genpos, int64_t *count, int64_t *total) {
    if (!path || (n_rows > 0 && (!position || !genpos || !count || !total)))
        return fail(BLMX_IO_ERR_ARG, "blmx_io_read_sites: null pointer");
    FileBuf f;
    if (int rc = f.load(path)) return rc;
    const char *p = f.data.data(), *end = p + f.data.size() - 1;
    int64_t row = -1;                                   // -1 = header
    while (p < end) {
        const char *nl = (const char *)std::memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        if (row >= 0) {
            if (row >= n_rows) return fail(BLMX_IO_ERR_FORMAT, "more rows than counted");
            const char *fb[8], *fe[8];
            const int nf = split_stripped(p, le, fb, fe);
            double c0, cg;
            int64_t k, n;
            if (nf < 4 || (strict_columns && nf != 4) || !parse_float(fb[0], fe[0], &c0) || !parse_int(fb[2], fe[2], &k) ||
                !parse_int(fb[3], fe[3], &n) || !parse_float(fb[use_phys ? 0 : 1], fe[use_phys ? 0 : 1], &cg))
                return fail(BLMX_IO_ERR_FORMAT, "line " + std::to_string(row + 2) + ": not '<float>\\t<float>\\t<int>\\t<int>'");
            if (!std::isfinite(c0) || std::fabs(c0) >= 9.2e18)
                return fail(BLMX_IO_ERR_FORMAT, "line " + std::to_string(row + 2) + ": position out of range");
            position[row] = (int64_t)c0;                // int(float(c0)) truncates toward zero
            // v1:103/124: float*(1-pt)*Rrate + float*pt with pt = 0 (physical) or 1 (genetic)
            genpos[row] = use_phys ? (cg * 1 * rrate + cg * 0) : (cg * 0 * rrate + cg * 1);
            count[row] = k;
            total[row] = n;
        }
        ++row;
        if (!nl) break;
        p = nl + 1;
    }
    if (row != n_rows && !(row == -1 && n_rows == 0))
        return fail(BLMX_IO_ERR_FORMAT, "row count changed between the two passes");
    return 0;
}

int blmx_io_write_rows(const char *path, const char *header, int64_t n_rows, const int64_t *physpos,
                       const double *genpos, const double *T, const int32_t *iA, const int32_t *ix,
                       const int32_t *ia, const int32_t *nsites, const char *const *A_text, int32_t n_A,
                       const char *const *x_text, int32_t n_x, const char *const *a_text, int32_t n_a) {
    if (!path || !header || (n_rows > 0 && (!physpos || !genpos || !T || !iA || !ix || !ia || !nsites)))
        return fail(BLMX_IO_ERR_ARG, "blmx_io_write_rows: null pointer");
    FILE *fh = std::fopen(path, "wb");
    if (!fh) return fail(BLMX_IO_ERR_OPEN, std::string("cannot open ") + path + ": " + std::strerror(errno));
    std::vector<char> buf(1 << 20);
    size_t used = 0;
    auto flush = [&]() {
        bool ok = std::fwrite(buf.data(), 1, used, fh) == used;
        used = 0;
        return ok;
    };
    std::fputs(header, fh);
    for (int64_t j = 0; j < n_rows; ++j) {
        if (used + 512 > buf.size() && !flush()) { std::fclose(fh); return fail(BLMX_IO_ERR_OPEN, "write failed"); }
        char *o = buf.data() + used;
        o = format_int(o, physpos[j]); *o++ = '\t';
        o = format_double(o, genpos[j]); *o++ = '\t';
        if (iA[j] < 0) {
            static const char zero[] = "0.0\t0.0\t0.0\t0.0\t0.0\n";
            std::memcpy(o, zero, sizeof(zero) - 1);
            o += sizeof(zero) - 1;
        } else {
            if (iA[j] >= n_A || ix[j] < 0 || ix[j] >= n_x || ia[j] < 0 || ia[j] >= n_a) {
                std::fclose(fh);
                return fail(BLMX_IO_ERR_ARG, "row " + std::to_string(j) + ": grid index out of range");
            }
            o = format_double(o, T[j]); *o++ = '\t';
            size_t l = std::strlen(x_text[ix[j]]); std::memcpy(o, x_text[ix[j]], l); o += l; *o++ = '\t';
            l = std::strlen(a_text[ia[j]]); std::memcpy(o, a_text[ia[j]], l); o += l; *o++ = '\t';
            l = std::strlen(A_text[iA[j]]); std::memcpy(o, A_text[iA[j]], l); o += l; *o++ = '\t';
            o = format_int(o, nsites[j]); *o++ = '\n';
        }
        used = (size_t)(o - buf.data());
    }
    bool ok = flush();
    ok = (std::fclose(fh) == 0) && ok;
    return ok ? 0 : fail(BLMX_IO_ERR_OPEN, "write failed");
}

int blmx_io_format_doubles(const double *v, int64_t n, char *out, int64_t out_cap, int64_t *out_len) {
    if (!v || !out || !out_len) return fail(BLMX_IO_ERR_ARG, "blmx_io_format_doubles: null pointer");
    char *o = out;
    for (int64_t i = 0; i < n; ++i) {
        if ((o - out) + 40 > out_cap) return fail(BLMX_IO_ERR_ARG, "blmx_io_format_doubles: buffer too small");
        o = format_double(o, v[i]);
        *o++ = '\n';
    }
    *out_len = (int64_t)(o - out);
    return 0;
}

}  // extern "C"

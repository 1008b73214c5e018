/*
 * blmx_io.h -- C ABI of the host-side text I/O fast paths.
 *
 * These replace, at genome scale, three pure-Python loops of the reference
 * (BalLeRMix+_v1.py, "v1") without changing a byte of what they produce:
 *   blmx_io_read_sites   InputData.readCounts / readPolyCalls parsing        (v1:80-131)
 *                        and the (x, n) columns getSpect / getConfig read    (v1:650,672)
 *   blmx_io_write_rows   the per-centre scores.write(f'...') of the Scan drivers (v1:574,591,607)
 * Any input the fast reader does not accept verbatim makes it return
 * BLMX_IO_ERR_FORMAT; the caller then runs the Python reader, which is the
 * reference's semantics by construction.  No CUDA involved.
 */
#ifndef BLMX_IO_H
#define BLMX_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLMX_IO_ERR_ARG    -1
#define BLMX_IO_ERR_OPEN   -2
#define BLMX_IO_ERR_FORMAT -3

const char *blmx_io_last_error(void);

/* Data rows of a 4-column input file (lines after the header, v1:83,116). */
int blmx_io_count_rows(const char *path, int64_t *n_rows);

/* physPos = int(float(c0)), k = int(c2), n = int(c3); genPos = float(c0)*rrate if use_phys
 * else float(c1)  (v1:88,103,121,124).  Arrays have n_rows elements (from blmx_io_count_rows).
 * strict_columns != 0 declines lines that do not have exactly four fields (getSpect / getConfig
 * unpack `split('\t')[2:]` into exactly two values, v1:650,672; InputData ignores extra columns). */
int blmx_io_read_sites(const char *path, int64_t n_rows, int use_phys, double rrate, int strict_columns,
                       int64_t *position, double *genpos, int64_t *count, int64_t *total);

/* `header`, then one row per centre: physPos, genPos, T, x, a, A, nSites (tab separated).
 * iA < 0 marks the reference's all-zero row.  A_text / x_text / a_text hold the grid values
 * formatted by Python (their int/float type is visible in the output). */
int blmx_io_write_rows(const char *path, const char *header, int64_t n_rows, const int64_t *physpos,
                       const double *genpos, const double *T, const int32_t *iA, const int32_t *ix,
                       const int32_t *ia, const int32_t *nsites, const char *const *A_text, int32_t n_A,
                       const char *const *x_text, int32_t n_x, const char *const *a_text, int32_t n_a);

/* str(numpy.float64(v)) for each v, newline separated (exposed for the formatting tests). */
int blmx_io_format_doubles(const double *v, int64_t n, char *out, int64_t out_cap, int64_t *out_len);

#ifdef __cplusplus
}
#endif
#endif /* BLMX_IO_H */

/*
 * blmx_mgpu.h -- multi-GPU entry of the B200-native BalLeRMix+ CLR scan.
 *
 * The reference has no multi-process mode: one `Scan` object calls `calcBaller` for every centre
 * of one input file in a single loop (BalLeRMix+_v1.py:539,573,590,606).  Centres are independent,
 * so that loop shards by contiguous centre ranges: one process (or thread) per GPU, every rank holds
 * the read-only problem (blmx_load), scans its cost-balanced share of the centre list, and ONE
 * collective -- NCCL send/recv of the per-centre rows to rank 0 -- finishes the run.  This entry is
 * that loop for a host written in C/C++/Go/...: the Python host does the same through
 * torch.distributed (ballermixplus_b200/sharding.py).
 *
 * The entry lives in libblmx_mgpu.so, a superset of libblmx.so (every symbol of blmx.h plus the two
 * below) linked against NCCL; libblmx.so itself has no NCCL dependency.
 */
#ifndef BLMX_MGPU_H
#define BLMX_MGPU_H

#include "blmx.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * Contiguous centre ranges of equal COST for `world` ranks: the cost of a centre is the number of sites
 * within alpha-reach summed over the A grid (what the kernel's work is proportional to), evaluated on the
 * device from the loaded problem.  begin[r], end[r] (r = 0..world-1) cover [0, n_centres) in order.
 * Every rank computes the same ranges from the same inputs.  HOST buffers.
 */
int blmx_shard_ranges(blmx_handle *h, int world, int64_t n_centres, const double *t, const int64_t *lo,
                      const int64_t *hi, int64_t *begin, int64_t *end);

/*
 * Scan the whole centre list on `world` GPUs.  Called by every rank with the SAME host arrays t/lo/hi
 * (n_centres entries, meaning as in blmx_scan) and its own handle (problem already loaded on its device).
 * `nccl_comm` is an initialised ncclComm_t of `world` ranks in which this caller is `rank`.  Rank 0 receives
 * every row into `out` (host buffers, n_centres entries, centre order); on the other ranks `out` may be NULL.
 * Synchronous.  Returns 0 or a negative blmx_status (BLMX_ERR_CUDA also covers NCCL errors; message in
 * blmx_last_error()).
 */
int blmx_scan_sharded(blmx_handle *h, int rank, int world, void *nccl_comm, int64_t n_centres,
                      const double *t, const int64_t *lo, const int64_t *hi, const blmx_result *out);

#ifdef __cplusplus
}
#endif
#endif /* BLMX_MGPU_H */

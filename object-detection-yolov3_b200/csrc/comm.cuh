// comm.cuh - NCCL all-gather plumbing of the tile-sharded path (see comm.cu).
#pragma once
#include "common.cuh"

namespace y3 {

void comm_unique_id(uint8_t* id /*[128]*/);
void comm_init(y3_context* ctx, int rank, int nranks, const uint8_t* id);
void comm_destroy(y3_context* ctx);
int comm_rank(const y3_context* ctx);
int comm_size(const y3_context* ctx);
int comm_nccl_version();
// Shard sizes of the tile-sharded path.  Equal by default; with Y3_ADAPTIVE_SHARDS=1 every rank's share of the tiles follows
// its measured throughput of the previous calls (an experiment that did not pay off: the per-step times are noise, see
// comm_update_shares).  shares: nranks fractions summing to 1 (equal until comm_update_shares has been called); the update takes
// the SAME all-gathered (tiles, microseconds) pairs on every rank, so all ranks derive identical boundaries.
const double* comm_shares(y3_context* ctx);
void comm_update_shares(y3_context* ctx, const long long* tiles, const long long* micros);
// stream-ordered on ctx->stream; a plain device copy when the handle has no (or a one-rank) communicator
void comm_all_gather_i64(y3_context* ctx, const long long* send_dev, long long* recv_dev, size_t count);
void comm_all_gather_f64(y3_context* ctx, const double* send_dev, double* recv_dev, size_t count);

}  // namespace y3

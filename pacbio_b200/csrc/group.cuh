// Per-read grouping of the hits by super-read (group.cu).
#pragma once
#include "common.cuh"

struct group_sort_args {
  uint64_t* keys_in;  uint64_t* pays_in;      // hits in emission order (read-major); scratch afterwards
  uint64_t* keys_out; uint64_t* pays_out;     // every read's slice stably sorted by super-read
  uint8_t*  head;                             // head[i] = 1 where a (read, super-read) group starts
  const uint64_t* hit_off;                    // first hit of every tile (+ the total)
  const uint32_t* tile_first;                 // first tile of every read (+ the number of tiles)
  uint32_t nseq_all;                          // super-read index of the hits that belong to no super-read
  int      sr_bits;                           // nseq_all < 2^sr_bits
  unsigned long long* n_invalid_groups;       // counts the reads that have such hits (each is one group)
  // filled by launch_group_sort: a read of at most `cap` hits is sorted whole in shared memory as words of
  // (super-read << idx_bits | position); larger reads are cut into buckets first, and runs of buckets of at most
  // range_cap hits are sorted as words of (super-read - base of the run << range_idx_bits | position)
  uint32_t cap, idx_bits, range_cap, range_idx_bits;
};

// whether the per-read sort handles an index whose super-read numbers take sr_bits bits
bool group_sort_usable(int sr_bits);
int launch_group_sort(mr_context* ctx, group_sort_args A, uint32_t nreads);
